// gg_elementwise.cu — script-level glue of the hot path:
// activations, broadcasting arithmetic, reductions, softmax, layout permutes, concat/slice copies,
// one-hot/argmax and the int->float input decode.  All HBM-bound: coalesced, float4 where the
// shape allows, grid sized to a multiple of the 148 SMs for the large cases.
#include "gg_common.cuh"

namespace gg {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
thread_local int g_last_backend = 0;
int g_conv_backend = 0;
int g_null_launch = 0;
namespace { __global__ void null_kernel() {} }
void launch_null(cudaStream_t st) { null_kernel<<<1, 32, 0, st>>>(); }
int g_pdl = 0;   // programmatic dependent launch for the stream-ordered kernels (gg_set_pdl / GG_PDL=1)
}  // namespace gg

using namespace gg;

extern "C" const char* gg_last_error(void) { return g_err; }
extern "C" int gg_version(void) { return 100; }
extern "C" long long gg_launch_count(void) { return g_launches.load(); }
extern "C" void gg_reset_launch_count(void) { g_launches.store(0); }
extern "C" int gg_set_conv_backend(int mode) {
  if (mode < 0 || mode > 2) return fail(GG_ERR_BAD_ARG, "gg_set_conv_backend: mode must be 0,1,2%s");
  g_conv_backend = mode;
  return GG_OK;
}
extern "C" int gg_get_conv_backend(void) { return g_conv_backend; }
extern "C" int gg_set_pdl(int on) {
  g_pdl = on ? 1 : 0;
  return GG_OK;
}
extern "C" int gg_set_null_launch(int on) {
  g_null_launch = on ? 1 : 0;
  return GG_OK;
}
extern "C" int gg_last_backend(void) { return g_last_backend; }

// ------------------------------------------------------------------------------------------
// unary
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float unary_apply(int op, float x, float a, float b) {
  switch (op) {
    case GG_U_COPY: return x;
    case GG_U_RELU: return x > 0.f ? x : 0.f;
    case GG_U_LEAKY: return fmaxf(a * x, x);
    case GG_U_TANH: return tanhf(x);
    case GG_U_SIGMOID: return 1.f / (1.f + expf(-x));
    case GG_U_EXP: return expf(x);
    case GG_U_LOG: return logf(x);
    case GG_U_SQRT: return sqrtf(x);
    case GG_U_SQUARE: return x * x;
    case GG_U_NEG: return -x;
    case GG_U_ABS: return fabsf(x);
    case GG_U_AFFINE: return a * x + b;
    case GG_U_POW: return powf(x, a);
    case GG_U_RSQRT: return rsqrtf(x);
    case GG_U_RECIP: return 1.f / x;
    case GG_U_BCE: return fmaxf(x, 0.f) - x * a + log1pf(expf(-fabsf(x)));
    case GG_U_CLIP: return fminf(fmaxf(x, a), b);
    case GG_U_SIGN: return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
    case GG_U_SOFTSIGN: return x / (1.f + fabsf(x));
    case GG_U_DIVC: return x / a;
    case GG_U_RDIVC: return a / x;
    default: return x;
  }
}

template <int OP>
__global__ void __launch_bounds__(256) unary_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                    float a, float b) {
  GG_PDL_ENTRY();
  long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  long long stride = (long long)gridDim.x * blockDim.x * 4;
  bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  for (; i4 < n; i4 += stride) {
    if (vec && i4 + 3 < n) {
      float4 v = *reinterpret_cast<const float4*>(x + i4);
      v.x = unary_apply(OP, v.x, a, b);
      v.y = unary_apply(OP, v.y, a, b);
      v.z = unary_apply(OP, v.z, a, b);
      v.w = unary_apply(OP, v.w, a, b);
      *reinterpret_cast<float4*>(y + i4) = v;
    } else {
      for (long long j = i4; j < n && j < i4 + 4; ++j) y[j] = unary_apply(OP, x[j], a, b);
    }
  }
}

static int ew_grid(long long n, int per_thread = 4, int threads = 256) {
  long long blocks = (n + (long long)threads * per_thread - 1) / ((long long)threads * per_thread);
  long long cap = (long long)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

#define GG_UNARY_CASE(OPC) \
  case OPC: GG_LAUNCH((unary_kernel<OPC>), grid, 256, 0, st, x, y, n, a, b); break;

extern "C" int gg_unary(int op, const float* x, float* y, long long n, float a, float b, void* stream) {
  if (n <= 0) return GG_OK;
  cudaStream_t st = as_stream(stream);
  int grid = ew_grid(n);
  switch (op) {
    GG_UNARY_CASE(GG_U_COPY) GG_UNARY_CASE(GG_U_RELU) GG_UNARY_CASE(GG_U_LEAKY) GG_UNARY_CASE(GG_U_TANH)
    GG_UNARY_CASE(GG_U_SIGMOID) GG_UNARY_CASE(GG_U_EXP) GG_UNARY_CASE(GG_U_LOG) GG_UNARY_CASE(GG_U_SQRT)
    GG_UNARY_CASE(GG_U_SQUARE) GG_UNARY_CASE(GG_U_NEG) GG_UNARY_CASE(GG_U_ABS) GG_UNARY_CASE(GG_U_AFFINE)
    GG_UNARY_CASE(GG_U_POW) GG_UNARY_CASE(GG_U_RSQRT) GG_UNARY_CASE(GG_U_RECIP) GG_UNARY_CASE(GG_U_BCE)
    GG_UNARY_CASE(GG_U_CLIP) GG_UNARY_CASE(GG_U_SIGN) GG_UNARY_CASE(GG_U_SOFTSIGN)
    GG_UNARY_CASE(GG_U_DIVC) GG_UNARY_CASE(GG_U_RDIVC)
    default: return fail(GG_ERR_BAD_ARG, "gg_unary: unknown op%s");
  }
  return check_launch("gg_unary");
}

// ------------------------------------------------------------------------------------------
// binary with broadcasting
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float binary_apply(int op, float a, float b, float alpha) {
  switch (op) {
    case GG_B_ADD: return a + b;
    case GG_B_SUB: return a - b;
    case GG_B_MUL: return a * b;
    case GG_B_DIV: return a / b;
    case GG_B_MAX: return fmaxf(a, b);
    case GG_B_MIN: return fminf(a, b);
    case GG_B_RELU_GRAD: return a > 0.f ? b : 0.f;
    case GG_B_LEAKY_GRAD: return a > 0.f ? b : alpha * b;
    case GG_B_TANH_GRAD: return (1.f - a * a) * b;
    case GG_B_SIGMOID_GRAD: return a * (1.f - a) * b;
    case GG_B_BCE_GRAD: return (1.f / (1.f + expf(-a)) - alpha) * b;
    case GG_B_GE_MASK: return a >= b ? 1.f : 0.f;
    case GG_B_GT_MASK: return a > b ? 1.f : 0.f;
    case GG_B_ABS_GRAD: return (a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f)) * b;
    case GG_B_POW: return powf(a, b);
    default: return a;
  }
}

// same-shape contiguous fast path
__global__ void __launch_bounds__(256) binary_flat_kernel(int op, const float* __restrict__ a, const float* __restrict__ b,
                                                          float* __restrict__ out, long long n, float alpha) {
  GG_PDL_ENTRY();
  long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  long long stride = (long long)gridDim.x * blockDim.x * 4;
  bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  for (; i4 < n; i4 += stride) {
    if (vec && i4 + 3 < n) {
      float4 va = *reinterpret_cast<const float4*>(a + i4);
      float4 vb = *reinterpret_cast<const float4*>(b + i4);
      float4 r;
      r.x = binary_apply(op, va.x, vb.x, alpha);
      r.y = binary_apply(op, va.y, vb.y, alpha);
      r.z = binary_apply(op, va.z, vb.z, alpha);
      r.w = binary_apply(op, va.w, vb.w, alpha);
      *reinterpret_cast<float4*>(out + i4) = r;
    } else {
      for (long long j = i4; j < n && j < i4 + 4; ++j) out[j] = binary_apply(op, a[j], b[j], alpha);
    }
  }
}

struct Dims4 { int d[4]; int sa[4]; int sb[4]; };

__global__ void __launch_bounds__(256) binary_bcast_kernel(int op, const float* __restrict__ a, const float* __restrict__ b,
                                                           float* __restrict__ out, Dims4 p, long long n, float alpha) {
  GG_PDL_ENTRY();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long r = i;
    int i3 = (int)(r % p.d[3]); r /= p.d[3];
    int i2 = (int)(r % p.d[2]); r /= p.d[2];
    int i1 = (int)(r % p.d[1]); r /= p.d[1];
    int i0 = (int)r;
    long long oa = (long long)i0 * p.sa[0] + (long long)i1 * p.sa[1] + (long long)i2 * p.sa[2] + (long long)i3 * p.sa[3];
    long long ob = (long long)i0 * p.sb[0] + (long long)i1 * p.sb[1] + (long long)i2 * p.sb[2] + (long long)i3 * p.sb[3];
    out[i] = binary_apply(op, a[oa], b[ob], alpha);
  }
}

extern "C" int gg_binary(int op, const float* a, const float* b, float* out, const int* dims4, const int* sa4,
                         const int* sb4, float alpha, void* stream) {
  if (op < 0 || op > GG_B_POW) return fail(GG_ERR_BAD_ARG, "gg_binary: unknown op%s");
  long long n = 1;
  for (int i = 0; i < 4; ++i) {
    if (dims4[i] <= 0) return GG_OK;
    n *= dims4[i];
  }
  // contiguous same-shape?
  bool flat = true;
  long long expect = 1;
  for (int i = 3; i >= 0; --i) {
    if (dims4[i] != 1 && (sa4[i] != expect || sb4[i] != expect)) flat = false;
    expect *= dims4[i];
  }
  cudaStream_t st = as_stream(stream);
  if (flat) {
    GG_LAUNCH(binary_flat_kernel, ew_grid(n), 256, 0, st, op, a, b, out, n, alpha);
  } else {
    Dims4 p;
    for (int i = 0; i < 4; ++i) { p.d[i] = dims4[i]; p.sa[i] = sa4[i]; p.sb[i] = sb4[i]; }
    GG_LAUNCH(binary_bcast_kernel, ew_grid(n, 1), 256, 0, st, op, a, b, out, p, n, alpha);
  }
  return check_launch("gg_binary");
}

// ------------------------------------------------------------------------------------------
// reduce over the middle axis of [outer, red, inner]
// ------------------------------------------------------------------------------------------
// inner == 1: one warp (or block) per output row, lanes stride over `red`.
__global__ void __launch_bounds__(256) reduce_rows_kernel(int op, const float* __restrict__ x, float* __restrict__ y,
                                                          int outer, int red) {
  GG_PDL_ENTRY();
  __shared__ float sh[32];
  int o = blockIdx.x;
  if (o >= outer) return;
  const float* row = x + (long long)o * red;
  float acc = (op == 2) ? -INFINITY : 0.f;
  for (int r = threadIdx.x; r < red; r += blockDim.x) {
    float v = row[r];
    acc = (op == 2) ? fmaxf(acc, v) : acc + v;
  }
  // block reduce
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  acc = (op == 2) ? warp_max(acc) : warp_sum(acc);
  if (lane == 0) sh[w] = acc;
  __syncthreads();
  if (w == 0) {
    float r = (lane < nw) ? sh[lane] : ((op == 2) ? -INFINITY : 0.f);
    r = (op == 2) ? warp_max(r) : warp_sum(r);
    if (lane == 0) y[o] = (op == 1) ? r / (float)red : r;
  }
}

// inner > 1: thread per (outer, inner) column, coalesced over inner; rows split over threadIdx.y then smem-combined
__global__ void __launch_bounds__(256) reduce_cols_kernel(int op, const float* __restrict__ x, float* __restrict__ y,
                                                          int outer, int red, int inner) {
  GG_PDL_ENTRY();
  __shared__ float sh[8][33];
  int i = blockIdx.x * 32 + threadIdx.x;
  int o = blockIdx.y;
  float acc = (op == 2) ? -INFINITY : 0.f;
  if (i < inner) {
    const float* base = x + (long long)o * red * inner + i;
    for (int r = threadIdx.y; r < red; r += 8) {
      float v = base[(long long)r * inner];
      acc = (op == 2) ? fmaxf(acc, v) : acc + v;
    }
  }
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < inner) {
    float r = sh[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) r = (op == 2) ? fmaxf(r, sh[k][threadIdx.x]) : r + sh[k][threadIdx.x];
    y[(long long)o * inner + i] = (op == 1) ? r / (float)red : r;
  }
}

// two-stage column reduction for tall matrices (bias gradients: [B*H*W, C] -> [C]): S row slices -> partial[S][inner],
// then the same kernel folds the S partial rows.  Deterministic (fixed summation order), fills the SMs.
__global__ void __launch_bounds__(256) reduce_cols_sliced_kernel(int op, const float* __restrict__ x, float* __restrict__ part,
                                                                 int red, int inner, int rows_per) {
  GG_PDL_ENTRY();
  __shared__ float sh[8][33];
  int i = blockIdx.x * 32 + threadIdx.x;
  int s = blockIdx.y;
  int r0 = s * rows_per, r1 = min(red, r0 + rows_per);
  float acc = (op == 2) ? -INFINITY : 0.f;
  if (i < inner) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      float v = x[(long long)r * inner + i];
      acc = (op == 2) ? fmaxf(acc, v) : acc + v;
    }
  }
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < inner) {
    float r = sh[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) r = (op == 2) ? fmaxf(r, sh[k][threadIdx.x]) : r + sh[k][threadIdx.x];
    part[(long long)s * inner + i] = r;
  }
}

extern "C" int gg_reduce_ws(int op, const float* x, float* y, int outer, int red, int inner, void* workspace,
                            size_t workspace_bytes, void* stream) {
  if (op < 0 || op > 2) return fail(GG_ERR_BAD_ARG, "gg_reduce_ws: unknown op%s");
  if (outer <= 0 || inner <= 0 || red <= 0) return GG_OK;
  int ctiles = ceil_div(inner, 32);
  int S = ceil_div(2 * kNumSMs, ctiles);
  if (S > red / 16) S = red / 16;
  if (S > 256) S = 256;
  if (outer != 1 || inner == 1 || S < 2 || workspace == nullptr || workspace_bytes < (size_t)S * inner * sizeof(float))
    return gg_reduce(op, x, y, outer, red, inner, stream);
  cudaStream_t st = as_stream(stream);
  int rows_per = ceil_div(red, S);
  S = ceil_div(red, rows_per);
  float* part = reinterpret_cast<float*>(workspace);
  dim3 block(32, 8);
  GG_LAUNCH(reduce_cols_sliced_kernel, dim3(ctiles, S), block, 0, st, op == 1 ? 0 : op, x, part, red, inner, rows_per);
  int rc = check_launch("gg_reduce_ws/slices");
  if (rc) return rc;
  // fold the S partial rows; the mean divides by the true row count
  GG_LAUNCH(reduce_cols_kernel, dim3(ctiles, 1), block, 0, st, op == 1 ? 0 : op, part, y, 1, S, inner);
  rc = check_launch("gg_reduce_ws/fold");
  if (rc || op != 1) return rc;
  return gg_unary(GG_U_DIVC, y, y, inner, (float)red, 0.f, stream);
}
extern "C" size_t gg_reduce_workspace(int outer, int red, int inner) {
  if (outer != 1 || inner == 1) return 0;
  int S = ceil_div(2 * kNumSMs, ceil_div(inner, 32));
  if (S > 256) S = 256;
  return (size_t)S * inner * sizeof(float);
}

extern "C" int gg_reduce(int op, const float* x, float* y, int outer, int red, int inner, void* stream) {
  if (op < 0 || op > 2) return fail(GG_ERR_BAD_ARG, "gg_reduce: unknown op%s");
  if (outer <= 0 || inner <= 0 || red <= 0) return GG_OK;
  cudaStream_t st = as_stream(stream);
  if (inner == 1) {
    int threads = red >= 1024 ? 256 : (red >= 128 ? 128 : 32);
    GG_LAUNCH(reduce_rows_kernel, outer, threads, 0, st, op, x, y, outer, red);
  } else {
    dim3 grid(ceil_div(inner, 32), outer), block(32, 8);
    GG_LAUNCH(reduce_cols_kernel, grid, block, 0, st, op, x, y, outer, red, inner);
  }
  return check_launch("gg_reduce");
}

// ------------------------------------------------------------------------------------------
// softmax over the last axis (HyperExtractor, gmgan_inference_cifar10.py:162-163): one warp per row
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int C) {
  GG_PDL_ENTRY();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + (long long)row * C;
  float m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, xr[c]);
  m = warp_max(m);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(xr[c] - m);
  s = warp_sum(s);
  float inv = 1.f / s;
  for (int c = lane; c < C; c += 32) y[(long long)row * C + c] = expf(xr[c] - m) * inv;
}

__global__ void __launch_bounds__(128) softmax_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                          float* __restrict__ dx, int R, int C) {
  GG_PDL_ENTRY();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* yr = y + (long long)row * C;
  const float* gr = dy + (long long)row * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += yr[c] * gr[c];
  s = warp_sum(s);
  for (int c = lane; c < C; c += 32) dx[(long long)row * C + c] = yr[c] * (gr[c] - s);
}

extern "C" int gg_softmax_fwd(const float* x, float* y, int R, int C, void* stream) {
  if (R <= 0 || C <= 0) return GG_OK;
  GG_LAUNCH(softmax_fwd_kernel, ceil_div(R, 4), 128, 0, as_stream(stream), x, y, R, C);
  return check_launch("gg_softmax_fwd");
}
extern "C" int gg_softmax_bwd(const float* y, const float* dy, float* dx, int R, int C, void* stream) {
  if (R <= 0 || C <= 0) return GG_OK;
  GG_LAUNCH(softmax_bwd_kernel, ceil_div(R, 4), 128, 0, as_stream(stream), y, dy, dx, R, C);
  return check_launch("gg_softmax_bwd");
}

// ------------------------------------------------------------------------------------------
// transposes
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_b2d_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int C) {
  GG_PDL_ENTRY();
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  const float* xb = x + (long long)b * R * C;
  float* yb = y + (long long)b * R * C;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = xb[(long long)r * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) yb[(long long)c * R + r] = tile[threadIdx.x][j];
  }
}

extern "C" int gg_transpose_b2d(const float* x, float* y, int Bt, int R, int C, void* stream) {
  if (Bt <= 0 || R <= 0 || C <= 0) return GG_OK;
  GG_REQUIRE(Bt <= 65535, "gg_transpose_b2d");
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32), Bt), block(32, 8);
  GG_LAUNCH(transpose_b2d_kernel, grid, block, 0, as_stream(stream), x, y, R, C);
  return check_launch("gg_transpose_b2d");
}

// batched transpose whose batches need not be contiguous: input batch b starts at x + b*x_bs (rows of C floats), output batch b
// at y + b*y_bs ([C][R]).  Serves (1) the NHWC -> NCHW flatten written straight into the column range of a concat buffer
// (y_bs = the concat's row length) and (2) its mirror in the backward pass: the column slice of the dense layer's input
// gradient transposed back without a slice copy (x_bs = the sliced tensor's row length) — optionally multiplied by act'(m) of
// the forward activation m (same layout as the output), i.e. the activation gradient that follows.
__global__ void __launch_bounds__(256) transpose_b2d_ex_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int C,
                                                               long long x_bs, long long y_bs, const float* __restrict__ mask,
                                                               int mask_act, float mask_alpha) {
  GG_PDL_ENTRY();
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  const float* xb = x + (long long)b * x_bs;
  float* yb = y + (long long)b * y_bs;
  const float* mb = mask ? mask + (long long)b * R * C : nullptr;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = r0 + j, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[j][threadIdx.x] = xb[(long long)r * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) {
      float v = tile[threadIdx.x][j];
      if (mb) v = act_grad_from_out(mb[(long long)c * R + r], v, mask_act, mask_alpha);
      yb[(long long)c * R + r] = v;
    }
  }
}

extern "C" int gg_transpose_b2d_ex(const float* x, float* y, int Bt, int R, int C, long long x_batch_stride,
                                   long long y_batch_stride, const float* mask, int mask_act, float mask_alpha, void* stream) {
  if (Bt <= 0 || R <= 0 || C <= 0) return GG_OK;
  GG_REQUIRE(Bt <= 65535, "gg_transpose_b2d_ex");
  GG_REQUIRE(x_batch_stride >= (long long)R * C && y_batch_stride >= (long long)R * C, "gg_transpose_b2d_ex");
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32), Bt), block(32, 8);
  GG_LAUNCH(transpose_b2d_ex_kernel, grid, block, 0, as_stream(stream), x, y, R, C, x_batch_stride, y_batch_stride, mask, mask_act,
            mask_alpha);
  return check_launch("gg_transpose_b2d_ex");
}

struct Perm4 { int od[4]; long long is[4]; };
__global__ void __launch_bounds__(256) transpose4_kernel(const float* __restrict__ x, float* __restrict__ y, Perm4 p, long long n) {
  GG_PDL_ENTRY();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long r = i;
    int i3 = (int)(r % p.od[3]); r /= p.od[3];
    int i2 = (int)(r % p.od[2]); r /= p.od[2];
    int i1 = (int)(r % p.od[1]); r /= p.od[1];
    int i0 = (int)r;
    y[i] = x[i0 * p.is[0] + i1 * p.is[1] + i2 * p.is[2] + i3 * p.is[3]];
  }
}

extern "C" int gg_transpose4(const float* x, float* y, const int* dims4, const int* perm4, void* stream) {
  long long istr[4];
  long long n = 1;
  for (int i = 3; i >= 0; --i) { istr[i] = n; n *= dims4[i]; }
  if (n <= 0) return GG_OK;
  Perm4 p;
  bool seen[4] = {false, false, false, false};
  for (int i = 0; i < 4; ++i) {
    int s = perm4[i];
    if (s < 0 || s > 3 || seen[s]) return fail(GG_ERR_BAD_ARG, "gg_transpose4: bad permutation%s");
    seen[s] = true;
    p.od[i] = dims4[s];
    p.is[i] = istr[s];
  }
  GG_LAUNCH(transpose4_kernel, ew_grid(n, 1), 256, 0, as_stream(stream), x, y, p, n);
  return check_launch("gg_transpose4");
}

// ------------------------------------------------------------------------------------------
// strided copy, fill, one-hot, argmax, casts, add_n
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) copy2d_kernel(const float* __restrict__ src, long long src_ld, float* __restrict__ dst,
                                                     long long dst_ld, long long rows, long long cols, int accumulate) {
  GG_PDL_ENTRY();
  long long n = rows * cols;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long r = i / cols, c = i - r * cols;
    float v = src[r * src_ld + c];
    float* d = dst + r * dst_ld + c;
    *d = accumulate ? (*d + v) : v;
  }
}
extern "C" int gg_copy2d(const float* src, long long src_ld, float* dst, long long dst_ld, long long rows,
                         long long cols, int accumulate, void* stream) {
  if (rows <= 0 || cols <= 0) return GG_OK;
  GG_LAUNCH(copy2d_kernel, ew_grid(rows * cols, 1), 256, 0, as_stream(stream), src, src_ld, dst, dst_ld, rows, cols, accumulate);
  return check_launch("gg_copy2d");
}

__global__ void __launch_bounds__(256) fill_kernel(float* x, long long n, float v) {
  GG_PDL_ENTRY();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) x[i] = v;
}
extern "C" int gg_fill(float* x, long long n, float v, void* stream) {
  if (n <= 0) return GG_OK;
  GG_LAUNCH(fill_kernel, ew_grid(n, 1), 256, 0, as_stream(stream), x, n, v);
  return check_launch("gg_fill");
}

__global__ void __launch_bounds__(256) one_hot_kernel(const int32_t* __restrict__ idx, float* __restrict__ out, int n, int depth) {
  GG_PDL_ENTRY();
  long long total = (long long)n * depth;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int r = (int)(i / depth), c = (int)(i - (long long)r * depth);
    out[i] = (idx[r] == c) ? 1.f : 0.f;
  }
}
extern "C" int gg_one_hot(const int32_t* idx, float* out, int n, int depth, void* stream) {
  if (n <= 0 || depth <= 0) return GG_OK;
  GG_LAUNCH(one_hot_kernel, ew_grid((long long)n * depth, 1), 256, 0, as_stream(stream), idx, out, n, depth);
  return check_launch("gg_one_hot");
}

// out[m, :] = table[idx[m], :] (zeros when idx[m] is outside [0, depth)): tf.matmul(tf.one_hot(idx, depth), table) without the
// one-hot matrix and the GEMM — the mixture-prior mean lookup `tf.matmul(tf.one_hot(k, N_COMS), mu)` at the head of both
// training steps (gmgan_inference_cifar10.py:138-141).  1*w + 0*(...) of the fp32 GEMM is w: identical values.
__global__ void __launch_bounds__(256) gather_rows_kernel(const int32_t* __restrict__ idx, const float* __restrict__ table,
                                                          const float* __restrict__ addend, float* __restrict__ out, int M, int N,
                                                          int depth) {
  GG_PDL_ENTRY();
  long long total = (long long)M * N;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int m = (int)(i / N), n = (int)(i - (long long)m * N);
    int k = idx[m];
    float v = (k >= 0 && k < depth) ? table[(long long)k * N + n] : 0.f;
    out[i] = addend ? v + addend[i] : v;
  }
}
extern "C" int gg_gather_rows(const int32_t* idx, const float* table, const float* addend, float* out, int M, int N, int depth,
                              void* stream) {
  if (M <= 0 || N <= 0) return GG_OK;
  GG_REQUIRE(depth > 0, "gg_gather_rows");
  GG_LAUNCH(gather_rows_kernel, ew_grid((long long)M * N, 1), 256, 0, as_stream(stream), idx, table, addend, out, M, N, depth);
  return check_launch("gg_gather_rows");
}

// first maximal index, like tf.argmax
__global__ void __launch_bounds__(128) argmax_kernel(const float* __restrict__ x, int32_t* __restrict__ idx, int R, int C) {
  GG_PDL_ENTRY();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* xr = x + (long long)row * C;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    float v = xr[c];
    if (v > best) { best = v; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) idx[row] = (bi == 0x7fffffff) ? 0 : bi;
}
extern "C" int gg_argmax(const float* x, int32_t* idx, int R, int C, void* stream) {
  if (R <= 0 || C <= 0) return GG_OK;
  GG_LAUNCH(argmax_kernel, ceil_div(R, 4), 128, 0, as_stream(stream), x, idx, R, C);
  return check_launch("gg_argmax");
}

template <typename T>
__global__ void __launch_bounds__(256) cast_to_f32_kernel(const T* __restrict__ x, float* __restrict__ y, long long n, float a, float b) {
  GG_PDL_ENTRY();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  // TF evaluates 2*((float(x)/255.)-.5): the host passes (a, b) and we keep the same operation order
  // as the graph that produced them by applying a single fused multiply-add only when exact parity is not
  // required; exact parity is handled on the host by emitting the op chain unfused.
  for (; i < n; i += stride) y[i] = a * (float)x[i] + b;
}
extern "C" int gg_cast_i32_f32(const int32_t* x, float* y, long long n, float a, float b, void* stream) {
  if (n <= 0) return GG_OK;
  GG_LAUNCH((cast_to_f32_kernel<int32_t>), ew_grid(n, 1), 256, 0, as_stream(stream), x, y, n, a, b);
  return check_launch("gg_cast_i32_f32");
}
extern "C" int gg_cast_u8_f32(const uint8_t* x, float* y, long long n, float a, float b, void* stream) {
  if (n <= 0) return GG_OK;
  GG_LAUNCH((cast_to_f32_kernel<uint8_t>), ew_grid(n, 1), 256, 0, as_stream(stream), x, y, n, a, b);
  return check_launch("gg_cast_u8_f32");
}
// uint8 -> int32 widening of a fed image batch: the host stages and copies ONE byte per pixel (the datasets' on-disk
// dtype, tflib/cifar10.py:8-48) and the int32 placeholder of the graph (gmgan_inference_cifar10.py:341) is filled on
// the device; 16 pixels per thread (one 16-byte load, four 16-byte stores).  Bit exact.
__global__ void __launch_bounds__(256) widen_u8_i32_kernel(const uint8_t* __restrict__ x, int32_t* __restrict__ y, long long n) {
  GG_PDL_ENTRY();
  const long long n16 = n >> 4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = i; v < n16; v += stride) {
    const uint4 q = reinterpret_cast<const uint4*>(x)[v];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    int4* o = reinterpret_cast<int4*>(y) + v * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = make_int4((int)(w[j] & 0xffu), (int)((w[j] >> 8) & 0xffu), (int)((w[j] >> 16) & 0xffu), (int)(w[j] >> 24));
  }
  for (long long t = (n16 << 4) + i; t < n; t += stride) y[t] = (int32_t)x[t];
}
extern "C" int gg_widen_u8_i32(const uint8_t* x, int32_t* y, long long n, void* stream) {
  if (n <= 0) return GG_OK;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return fail(GG_ERR_BAD_ARG, "gg_widen_u8_i32: buffers must be 16-byte aligned%s");
  GG_LAUNCH(widen_u8_i32_kernel, ew_grid((n + 15) / 16, 1), 256, 0, as_stream(stream), x, y, n);
  return check_launch("gg_widen_u8_i32");
}
__global__ void __launch_bounds__(256) cast_f32_i32_kernel(const float* __restrict__ x, int32_t* __restrict__ y, long long n) {
  GG_PDL_ENTRY();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = (int32_t)x[i];  // truncation toward zero, like tf.cast / numpy astype
}
extern "C" int gg_cast_f32_i32(const float* x, int32_t* y, long long n, void* stream) {
  if (n <= 0) return GG_OK;
  GG_LAUNCH(cast_f32_i32_kernel, ew_grid(n, 1), 256, 0, as_stream(stream), x, y, n);
  return check_launch("gg_cast_f32_i32");
}

struct PtrList { const float* p[16]; };
__global__ void __launch_bounds__(256) add_n_kernel(PtrList pl, int count, float* __restrict__ out, long long n) {
  GG_PDL_ENTRY();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float acc = pl.p[0][i];
    for (int k = 1; k < count; ++k) acc += pl.p[k][i];
    out[i] = acc;
  }
}
extern "C" int gg_add_n(const float* const* ptrs, int count, float* out, long long n, void* stream) {
  if (n <= 0) return GG_OK;
  if (count < 1 || count > 16) return fail(GG_ERR_BAD_ARG, "gg_add_n: count must be in [1,16]%s");
  PtrList pl;
  for (int i = 0; i < count; ++i) pl.p[i] = ptrs[i];
  GG_LAUNCH(add_n_kernel, ew_grid(n, 1), 256, 0, as_stream(stream), pl, count, out, n);
  return check_launch("gg_add_n");
}

// ------------------------------------------------------------------------------------------
// fused element-wise programs (include/gg_b200.h: gg_ew_program)
// ------------------------------------------------------------------------------------------
// One thread evaluates the whole program for one element of the iteration space.  The register file lives in shared memory
// ([register][thread]: conflict-free, dynamically indexable); the program itself is a __grid_constant__ kernel parameter, so
// instruction fetch is a uniform constant-bank load.  Arithmetic goes through unary_apply / binary_apply — the code the
// one-op kernels run — so a fused group is bit-identical to the launches it replaces (tests/test_gpu_fusion.py).
#define GG_EW_THREADS 256

__device__ __forceinline__ float ew_load(const gg_ew_program& p, int k, long long off) {
  return p.in_is_int[k] ? (float)reinterpret_cast<const int32_t*>(p.in[k])[off] : reinterpret_cast<const float*>(p.in[k])[off];
}

__device__ __forceinline__ void ew_run_instrs(const gg_ew_program& p, float (*regs)[GG_EW_THREADS], int t) {
  for (int j = 0; j < p.n_instr; ++j) {
    const gg_ew_instr& q = p.instr[j];
    float v;
    if (q.kind == 0) v = unary_apply(q.op, regs[q.src0][t], q.a, q.b);
    else v = binary_apply(q.op, regs[q.src0][t], regs[q.src1][t], q.a);
    regs[q.dst][t] = v;
  }
}

__global__ void __launch_bounds__(GG_EW_THREADS) ew_program_kernel(const __grid_constant__ gg_ew_program p, long long n) {
  GG_PDL_ENTRY();
  __shared__ float regs[GG_EW_REGS][GG_EW_THREADS];
  const int t = threadIdx.x;
  long long i = (long long)blockIdx.x * blockDim.x + t;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    if (p.flat) {
      for (int k = 0; k < p.n_in; ++k) regs[k][t] = ew_load(p, k, i);
    } else {
      // the iteration space has < 2^31 elements (checked by the host): 32-bit divisions
      uint32_t r = (uint32_t)i;
      const uint32_t i3 = r % (uint32_t)p.dims[3]; r /= (uint32_t)p.dims[3];
      const uint32_t i2 = r % (uint32_t)p.dims[2]; r /= (uint32_t)p.dims[2];
      const uint32_t i1 = r % (uint32_t)p.dims[1]; r /= (uint32_t)p.dims[1];
      const uint32_t i0 = r;
      for (int k = 0; k < p.n_in; ++k)
        regs[k][t] = ew_load(p, k, (long long)i0 * p.in_stride[k][0] + (long long)i1 * p.in_stride[k][1] +
                                       (long long)i2 * p.in_stride[k][2] + (long long)i3 * p.in_stride[k][3]);
    }
    ew_run_instrs(p, regs, t);
    for (int k = 0; k < p.n_out; ++k) p.out[k][i] = regs[p.out_reg[k]][t];
  }
}

// output 0 reduced along the last dimension, one CTA per row, with reduce_rows_kernel's summation order (thread-strided partial
// sums, warp butterfly, per-warp partials in shared memory, one more butterfly); outputs 1.. are stored per element.  The row's
// coordinates are decomposed ONCE per CTA (thread k computes the row base offset of input k): no division in the element loop
__global__ void __launch_bounds__(GG_EW_THREADS) ew_program_reduce_kernel(const __grid_constant__ gg_ew_program p, int rows, int red) {
  GG_PDL_ENTRY();
  __shared__ float regs[GG_EW_REGS][GG_EW_THREADS];
  __shared__ float sh[32];
  __shared__ long long s_base[GG_EW_MAX_IN];
  const int t = threadIdx.x;
  const int o = blockIdx.x;
  if (o >= rows) return;
  if (t < p.n_in) {
    uint32_t r = (uint32_t)o;
    const uint32_t i2 = r % (uint32_t)p.dims[2]; r /= (uint32_t)p.dims[2];
    const uint32_t i1 = r % (uint32_t)p.dims[1]; r /= (uint32_t)p.dims[1];
    s_base[t] = (long long)r * p.in_stride[t][0] + (long long)i1 * p.in_stride[t][1] + (long long)i2 * p.in_stride[t][2];
  }
  __syncthreads();
  const bool is_max = p.reduce_op == 3;
  float acc = is_max ? -INFINITY : 0.f;
  for (int r = t; r < red; r += blockDim.x) {
    for (int k = 0; k < p.n_in; ++k) regs[k][t] = ew_load(p, k, s_base[k] + (long long)r * p.in_stride[k][3]);
    ew_run_instrs(p, regs, t);
    const long long idx = (long long)o * red + r;
    const float v = regs[p.out_reg[0]][t];
    acc = is_max ? fmaxf(acc, v) : acc + v;
    for (int k = 1; k < p.n_out; ++k) p.out[k][idx] = regs[p.out_reg[k]][t];
  }
  int lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
  acc = is_max ? warp_max(acc) : warp_sum(acc);
  if (lane == 0) sh[w] = acc;
  __syncthreads();
  if (w == 0) {
    float r = (lane < nw) ? sh[lane] : (is_max ? -INFINITY : 0.f);
    r = is_max ? warp_max(r) : warp_sum(r);
    if (lane == 0) p.out[0][o] = (p.reduce_op == 2) ? r / (float)red : r;
  }
}

extern "C" int gg_ew_program_bytes(void) { return (int)sizeof(gg_ew_program); }

extern "C" int gg_ew_run(const gg_ew_program* prog, void* stream) {
  if (prog == nullptr) return fail(GG_ERR_BAD_ARG, "gg_ew_run: null program%s");
  const gg_ew_program& p = *prog;
  GG_REQUIRE(p.n_in >= 0 && p.n_in <= GG_EW_MAX_IN && p.n_out >= 1 && p.n_out <= GG_EW_MAX_OUT && p.n_instr >= 0 &&
             p.n_instr <= GG_EW_MAX_INSTR, "gg_ew_run");
  GG_REQUIRE(p.reduce_op >= 0 && p.reduce_op <= 3, "gg_ew_run");
  long long n = 1;
  for (int i = 0; i < 4; ++i) {
    if (p.dims[i] <= 0) return GG_OK;
    n *= p.dims[i];
  }
  GG_REQUIRE(n <= 0x7fffffffLL, "gg_ew_run");
  for (int k = 0; k < p.n_in; ++k) GG_REQUIRE(p.in[k] != nullptr, "gg_ew_run");
  for (int k = 0; k < p.n_out; ++k) GG_REQUIRE(p.out[k] != nullptr && p.out_reg[k] >= 0 && p.out_reg[k] < GG_EW_REGS, "gg_ew_run");
  for (int j = 0; j < p.n_instr; ++j) {
    const gg_ew_instr& q = p.instr[j];
    GG_REQUIRE(q.kind == 0 || q.kind == 1, "gg_ew_run");
    GG_REQUIRE(q.dst >= 0 && q.dst < GG_EW_REGS && q.src0 >= 0 && q.src0 < GG_EW_REGS, "gg_ew_run");
    GG_REQUIRE(q.kind == 0 ? (q.op >= 0 && q.op <= GG_U_RDIVC) : (q.op >= 0 && q.op <= GG_B_POW && q.src1 >= 0 && q.src1 < GG_EW_REGS),
               "gg_ew_run");
  }
  cudaStream_t st = as_stream(stream);
  if (p.reduce_op == 0) {
    GG_LAUNCH(ew_program_kernel, ew_grid(n, 1, GG_EW_THREADS), GG_EW_THREADS, 0, st, p, n);
  } else {
    const int red = p.dims[3];
    const long long rows = n / red;
    GG_REQUIRE(rows <= 0x7fffffffLL, "gg_ew_run");
    int threads = red >= 1024 ? 256 : (red >= 128 ? 128 : 32);          // = gg_reduce's row kernel
    GG_LAUNCH(ew_program_reduce_kernel, (int)rows, threads, 0, st, p, (int)rows, red);
  }
  return check_launch("gg_ew_run");
}
