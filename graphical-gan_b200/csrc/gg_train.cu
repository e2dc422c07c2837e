// gg_train.cu — batch normalisation, fused loss reductions, Adam/RMSProp multi-tensor updates, Philox RNG.
// Reference call sites: tflib/ops/batchnorm.py:30,77-84; tflib/objs/gan_inference.py:85-117,28-45;
// tflib/utils/distance.py:3-7; gmgan_inference_cifar10.py:117-120,349-350.  All HBM/latency-bound.
#include "gg_common.cuh"

using namespace gg;

// ------------------------------------------------------------------------------------------
// batch norm over the rows of x[R,C]
// ------------------------------------------------------------------------------------------
static int bn_slices(int R, int C) {
  int ctiles = ceil_div(C, 32);
  int want = ceil_div(2 * kNumSMs, ctiles);  // ~2 waves of blocks
  int maxs = R / 16;                         // at least 16 rows per slice
  if (maxs < 1) maxs = 1;
  int S = want < maxs ? want : maxs;
  if (S > 64) S = 64;
  if (S < 1) S = 1;
  return S;
}
extern "C" int gg_bn_slices(int R, int C) { return bn_slices(R, C); }

__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, float* __restrict__ partial, int R, int C, int S) {
  GG_PDL_ENTRY();
  __shared__ float s1[8][33], s2[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  int s = blockIdx.y;
  int rows_per = (R + S - 1) / S;
  int r0 = s * rows_per, r1 = min(R, r0 + rows_per);
  float a = 0.f, b = 0.f;
  if (c < C) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      float v = x[(long long)r * C + c];
      a += v;
      b += v * v;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { a += s1[k][threadIdx.x]; b += s2[k][threadIdx.x]; }
    partial[((long long)s * 2 + 0) * C + c] = a;
    partial[((long long)s * 2 + 1) * C + c] = b;
  }
}

extern "C" int gg_bn_stats(const float* x, float* partial, int R, int C, void* stream) {
  if (R <= 0 || C <= 0) return GG_OK;
  int S = bn_slices(R, C);
  dim3 grid(ceil_div(C, 32), S), block(32, 8);
  GG_LAUNCH(bn_stats_kernel, grid, block, 0, as_stream(stream), x, partial, R, C, S);
  return check_launch("gg_bn_stats");
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ partial, int S,
                                                       float count, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps, float* __restrict__ y,
                                                       float* __restrict__ mean_out, float* __restrict__ rstd_out, int R, int C,
                                                       int rows_per, int act, float alpha) {
  GG_PDL_ENTRY();
  __shared__ float sc[32], sh[32];
  __shared__ double pa[8][33], pb[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  {
    // fold the S partial slices with all 8 row-lanes (the loads of a lane are independent and pipeline; a single
    // lane walking 64 slices serially cost ~10 us on the 16384-row layers)
    double a = 0.0, b = 0.0;
    if (c < C)
      for (int s = threadIdx.y; s < S; s += 8) {
        a += (double)partial[((long long)s * 2 + 0) * C + c];
        b += (double)partial[((long long)s * 2 + 1) * C + c];
      }
    pa[threadIdx.y][threadIdx.x] = a;
    pb[threadIdx.y][threadIdx.x] = b;
  }
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += pa[k][threadIdx.x]; b += pb[k][threadIdx.x]; }
    double mean = a / (double)count;
    double var = b / (double)count - mean * mean;  // biased batch variance (fused_batch_norm is_training)
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float g = gamma ? gamma[c] : 1.f, bt = beta ? beta[c] : 0.f;
    sc[threadIdx.x] = rstd * g;
    sh[threadIdx.x] = bt - (float)mean * rstd * g;
    if (blockIdx.y == 0) {
      if (mean_out) mean_out[c] = (float)mean;
      if (rstd_out) rstd_out[c] = rstd;
    }
  }
  __syncthreads();
  if (c >= C) return;
  float scale = sc[threadIdx.x], shift = sh[threadIdx.x];
  int r0 = blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
  for (int r = r0 + threadIdx.y; r < r1; r += 8) {
    long long i = (long long)r * C + c;
    y[i] = apply_act(x[i] * scale + shift, act, alpha);
  }
}

extern "C" int gg_bn_apply(const float* x, const float* partial, int S, float count, const float* gamma, const float* beta,
                           float eps, float* y, float* mean_out, float* rstd_out, int R, int C, int act, float alpha,
                           void* stream) {
  if (R <= 0 || C <= 0) return GG_OK;
  int ctiles = ceil_div(C, 32);
  int want = ceil_div(4 * kNumSMs, ctiles);
  int maxs = ceil_div(R, 8);
  int gy = want < maxs ? want : maxs;
  if (gy < 1) gy = 1;
  int rows_per = ceil_div(R, gy);
  gy = ceil_div(R, rows_per);
  dim3 grid(ctiles, gy), block(32, 8);
  GG_LAUNCH(bn_apply_kernel, grid, block, 0, as_stream(stream), x, partial, S, count, gamma, beta, eps, y, mean_out, rstd_out, R, C,
                                                         rows_per, act, alpha);
  return check_launch("gg_bn_apply");
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ y, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ partial, int R,
                                                            int C, int S, int act, float alpha) {
  GG_PDL_ENTRY();
  __shared__ float s1[8][33], s2[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  int s = blockIdx.y;
  int rows_per = (R + S - 1) / S;
  int r0 = s * rows_per, r1 = min(R, r0 + rows_per);
  float a = 0.f, b = 0.f;
  if (c < C) {
    float m = mean[c], rs = rstd[c];
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      long long i = (long long)r * C + c;
      float g = dy[i];
      if (act != GG_ACT_NONE) g = act_grad_from_out(y[i], g, act, alpha);
      float xh = (x[i] - m) * rs;
      a += g;
      b += g * xh;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { a += s1[k][threadIdx.x]; b += s2[k][threadIdx.x]; }
    partial[((long long)s * 2 + 0) * C + c] = a;
    partial[((long long)s * 2 + 1) * C + c] = b;
  }
}

extern "C" int gg_bn_bwd_reduce(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                                const float* gamma, const float* beta, float* partial, int R, int C, int act, float alpha,
                                void* stream) {
  (void)gamma; (void)beta;
  if (R <= 0 || C <= 0) return GG_OK;
  if (act != GG_ACT_NONE && y == nullptr) return fail(GG_ERR_BAD_ARG, "gg_bn_bwd_reduce: y required when act is fused%s");
  int S = bn_slices(R, C);
  dim3 grid(ceil_div(C, 32), S), block(32, 8);
  GG_LAUNCH(bn_bwd_reduce_kernel, grid, block, 0, as_stream(stream), dy, x, y, mean, rstd, partial, R, C, S, act, alpha);
  return check_launch("gg_bn_bwd_reduce");
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ partial, int S, float count,
                                                           float* __restrict__ dx, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int R, int C, int rows_per, int act,
                                                           float alpha) {
  GG_PDL_ENTRY();
  __shared__ float sg[32], sgx[32];
  __shared__ double pa[8][33], pb[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  {
    double a = 0.0, b = 0.0;
    if (c < C)
      for (int s = threadIdx.y; s < S; s += 8) {
        a += (double)partial[((long long)s * 2 + 0) * C + c];
        b += (double)partial[((long long)s * 2 + 1) * C + c];
      }
    pa[threadIdx.y][threadIdx.x] = a;
    pb[threadIdx.y][threadIdx.x] = b;
  }
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { a += pa[k][threadIdx.x]; b += pb[k][threadIdx.x]; }
    sg[threadIdx.x] = (float)(a / (double)count);
    sgx[threadIdx.x] = (float)(b / (double)count);
    if (blockIdx.y == 0) {
      if (dbeta) dbeta[c] = (float)a;
      if (dgamma) dgamma[c] = (float)b;
    }
  }
  __syncthreads();
  if (c >= C || dx == nullptr) return;
  float m = mean[c], rs = rstd[c], gm = gamma ? gamma[c] : 1.f;
  float mg = sg[threadIdx.x], mgx = sgx[threadIdx.x];
  float k = gm * rs;
  int r0 = blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
  for (int r = r0 + threadIdx.y; r < r1; r += 8) {
    long long i = (long long)r * C + c;
    float g = dy[i];
    if (act != GG_ACT_NONE) g = act_grad_from_out(y[i], g, act, alpha);
    float xh = (x[i] - m) * rs;
    dx[i] = k * (g - mg - xh * mgx);
  }
}

extern "C" int gg_bn_bwd_apply(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                               const float* gamma, const float* beta, const float* partial, int S, float count, float* dx,
                               float* dgamma, float* dbeta, int R, int C, int act, float alpha, void* stream) {
  (void)beta;
  if (R <= 0 || C <= 0) return GG_OK;
  if (act != GG_ACT_NONE && y == nullptr) return fail(GG_ERR_BAD_ARG, "gg_bn_bwd_apply: y required when act is fused%s");
  int ctiles = ceil_div(C, 32);
  int want = ceil_div(4 * kNumSMs, ctiles);
  int maxs = ceil_div(R, 8);
  int gy = want < maxs ? want : maxs;
  if (gy < 1) gy = 1;
  int rows_per = ceil_div(R, gy);
  gy = ceil_div(R, rows_per);
  dim3 grid(ctiles, gy), block(32, 8);
  GG_LAUNCH(bn_bwd_apply_kernel, grid, block, 0, as_stream(stream), dy, x, y, mean, rstd, gamma, partial, S, count, dx, dgamma, dbeta,
                                                             R, C, rows_per, act, alpha);
  return check_launch("gg_bn_bwd_apply");
}

__global__ void __launch_bounds__(256) bn_fold_kernel(const float* __restrict__ partial, int S, float* __restrict__ out, int C2) {
  GG_PDL_ENTRY();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C2) return;
  double a = 0.0;
  for (int s = 0; s < S; ++s) a += (double)partial[(long long)s * C2 + i];
  out[i] = (float)a;
}
extern "C" int gg_bn_fold_partials(const float* partial, int S, float* out, int C, void* stream) {
  if (C <= 0) return GG_OK;
  GG_LAUNCH(bn_fold_kernel, ceil_div(2 * C, 256), 256, 0, as_stream(stream), partial, S, out, 2 * C);
  return check_launch("gg_bn_fold_partials");
}

// ------------------------------------------------------------------------------------------
// fused loss reductions (single block: the logit vectors are [B] with B <= a few thousand)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_mean_kernel(const float* __restrict__ x, int n, float label, float weight,
                                                       float* __restrict__ out, int accumulate) {
  GG_PDL_ENTRY();
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = x[i];
    acc += fmaxf(v, 0.f) - v * label + log1pf(expf(-fabsf(v)));
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    float r = weight * acc / (float)n;
    out[0] = accumulate ? out[0] + r : r;
  }
}
extern "C" int gg_bce_mean(const float* x, int n, float label, float weight, float* out, int accumulate, void* stream) {
  GG_REQUIRE(n > 0, "gg_bce_mean");
  GG_LAUNCH(bce_mean_kernel, 1, 256, 0, as_stream(stream), x, n, label, weight, out, accumulate);
  return check_launch("gg_bce_mean");
}

__global__ void __launch_bounds__(256) bce_mean_grad_kernel(const float* __restrict__ x, int n, float label, float weight,
                                                            const float* __restrict__ gscale, float* __restrict__ dx,
                                                            int accumulate) {
  GG_PDL_ENTRY();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = weight / (float)n * (gscale ? gscale[0] : 1.f);
  float g = (1.f / (1.f + expf(-x[i])) - label) * s;
  dx[i] = accumulate ? dx[i] + g : g;
}
extern "C" int gg_bce_mean_grad(const float* x, int n, float label, float weight, const float* gscale, float* dx,
                                int accumulate, void* stream) {
  GG_REQUIRE(n > 0, "gg_bce_mean_grad");
  GG_LAUNCH(bce_mean_grad_kernel, ceil_div(n, 256), 256, 0, as_stream(stream), x, n, label, weight, gscale, dx, accumulate);
  return check_launch("gg_bce_mean_grad");
}

// two-stage deterministic mean of (x-y)^2 or |x-y|; stage 2 is folded into the last-arriving block
__global__ void __launch_bounds__(256) dist_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n,
                                                           int p, float* __restrict__ part) {
  GG_PDL_ENTRY();
  __shared__ float sh[32];
  float acc = 0.f;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float d = x[i] - y[i];
    acc += (p == 2) ? d * d : fabsf(d);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256) dist_final_kernel(const float* __restrict__ part, int nb, long long n, float weight,
                                                         float* __restrict__ out, int accumulate) {
  GG_PDL_ENTRY();
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += part[i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    float r = weight * acc / (float)n;
    out[0] = accumulate ? out[0] + r : r;
  }
}
static float* g_dist_scratch = nullptr;
extern "C" int gg_dist_mean(const float* x, const float* y, long long n, int p, float weight, float* out, int accumulate,
                            void* stream) {
  GG_REQUIRE(n > 0 && (p == 1 || p == 2), "gg_dist_mean");
  if (!g_dist_scratch) {
    cudaError_t e = cudaMalloc(&g_dist_scratch, sizeof(float) * 1024);
    if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "gg_dist_mean: cudaMalloc failed%s");
  }
  int nb = ceil_div(n, 256 * 8);
  if (nb > 592) nb = 592;
  GG_LAUNCH(dist_partial_kernel, nb, 256, 0, as_stream(stream), x, y, n, p, g_dist_scratch);
  int rc = check_launch("gg_dist_mean/partial");
  if (rc) return rc;
  GG_LAUNCH(dist_final_kernel, 1, 256, 0, as_stream(stream), g_dist_scratch, nb, n, weight, out, accumulate);
  return check_launch("gg_dist_mean/final");
}

// slopes[r] = ||g[r,:]||_2 ; out = weight * mean (slope-1)^2.  One warp per row, single block does the final mean.
__global__ void __launch_bounds__(256) gp_slopes_kernel(const float* __restrict__ g, int R, int C, float* __restrict__ slopes) {
  GG_PDL_ENTRY();
  int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= R) return;
  const float* gr = g + (long long)row * C;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) { float v = gr[c]; acc += v * v; }
  acc = warp_sum(acc);
  if (lane == 0) slopes[row] = sqrtf(acc);
}
__global__ void __launch_bounds__(256) gp_penalty_kernel(const float* __restrict__ slopes, int R, float weight, float* __restrict__ out) {
  GG_PDL_ENTRY();
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) { float d = slopes[i] - 1.f; acc += d * d; }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) out[0] = weight * acc / (float)R;
}
extern "C" int gg_gp_slope_penalty(const float* g, int R, int C, float weight, float* slopes, float* out, void* stream) {
  GG_REQUIRE(R > 0 && C > 0, "gg_gp_slope_penalty");
  GG_LAUNCH(gp_slopes_kernel, ceil_div(R, 8), 256, 0, as_stream(stream), g, R, C, slopes);
  int rc = check_launch("gg_gp_slope_penalty/slopes");
  if (rc) return rc;
  GG_LAUNCH(gp_penalty_kernel, 1, 256, 0, as_stream(stream), slopes, R, weight, out);
  return check_launch("gg_gp_slope_penalty/mean");
}

// ------------------------------------------------------------------------------------------
// multi-tensor Adam / RMSProp / bucket pack
// ------------------------------------------------------------------------------------------
struct AdamState { double b1t; double b2t; long long t; };

__global__ void adam_tick_kernel(AdamState* st, double b1, double b2) {
  GG_PDL_ENTRY();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (st->t == 0) { st->b1t = 1.0; st->b2t = 1.0; }
    st->t += 1;
    st->b1t *= b1;
    st->b2t *= b2;
  }
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const gg_adam_entry* __restrict__ table,
                                                         const gg_adam_chunk* __restrict__ chunks, int n_chunks,
                                                         const AdamState* __restrict__ st, float lr, float b1, float b2,
                                                         float eps, float gscale) {
  GG_PDL_ENTRY();
  int ci = blockIdx.x;
  if (ci >= n_chunks) return;
  gg_adam_chunk ch = chunks[ci];
  gg_adam_entry e = table[ch.tensor];
  // TensorFlow ApplyAdam: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), epsilon outside the bias correction
  float lr_t = (float)((double)lr * sqrt(1.0 - st->b2t) / (1.0 - st->b1t));
  long long end = min(e.n, ch.offset + (long long)GG_ADAM_CHUNK);
  const float om1 = 1.f - b1, om2 = 1.f - b2;
  const bool vec = ((reinterpret_cast<uintptr_t>(e.p) | reinterpret_cast<uintptr_t>(e.g) | reinterpret_cast<uintptr_t>(e.m) |
                     reinterpret_cast<uintptr_t>(e.v)) & 15) == 0 && ((end - ch.offset) & 3) == 0;
  if (vec) {   // 28 B/parameter of traffic: 128-bit accesses (chunk offsets are multiples of 4096 elements)
    for (long long i = ch.offset + threadIdx.x * 4; i < end; i += blockDim.x * 4) {
      float4 g = *reinterpret_cast<const float4*>(e.g + i);
      float4 m = *reinterpret_cast<float4*>(e.m + i);
      float4 v = *reinterpret_cast<float4*>(e.v + i);
      float4 p = *reinterpret_cast<float4*>(e.p + i);
      g.x *= gscale; g.y *= gscale; g.z *= gscale; g.w *= gscale;
      m.x += (g.x - m.x) * om1; m.y += (g.y - m.y) * om1; m.z += (g.z - m.z) * om1; m.w += (g.w - m.w) * om1;
      v.x += (g.x * g.x - v.x) * om2; v.y += (g.y * g.y - v.y) * om2; v.z += (g.z * g.z - v.z) * om2; v.w += (g.w * g.w - v.w) * om2;
      p.x -= lr_t * m.x / (sqrtf(v.x) + eps); p.y -= lr_t * m.y / (sqrtf(v.y) + eps);
      p.z -= lr_t * m.z / (sqrtf(v.z) + eps); p.w -= lr_t * m.w / (sqrtf(v.w) + eps);
      *reinterpret_cast<float4*>(e.m + i) = m;
      *reinterpret_cast<float4*>(e.v + i) = v;
      *reinterpret_cast<float4*>(e.p + i) = p;
    }
    return;
  }
  for (long long i = ch.offset + threadIdx.x; i < end; i += blockDim.x) {
    float g = e.g[i] * gscale;
    float m = e.m[i] + (g - e.m[i]) * om1;
    float v = e.v[i] + (g * g - e.v[i]) * om2;
    e.m[i] = m;
    e.v[i] = v;
    e.p[i] = e.p[i] - lr_t * m / (sqrtf(v) + eps);
  }
}

extern "C" int gg_adam_multi(const gg_adam_entry* table, const gg_adam_chunk* chunks, int n_chunks, void* state, float lr,
                             float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (n_chunks <= 0) return GG_OK;
  cudaStream_t st = as_stream(stream);
  GG_LAUNCH(adam_tick_kernel, 1, 32, 0, st, reinterpret_cast<AdamState*>(state), (double)beta1, (double)beta2);
  int rc = check_launch("gg_adam_multi/tick");
  if (rc) return rc;
  GG_LAUNCH(adam_multi_kernel, n_chunks, 256, 0, st, table, chunks, n_chunks, reinterpret_cast<const AdamState*>(state), lr, beta1,
                                              beta2, eps, grad_scale);
  return check_launch("gg_adam_multi");
}

// the two halves of gg_adam_multi as separate entry points: a step whose update is split by gradient readiness advances the
// bias-correction state once (gg_adam_tick) and applies the update to disjoint parameter sets at different times (gg_adam_apply)
extern "C" int gg_adam_tick(void* state, float beta1, float beta2, void* stream) {
  GG_LAUNCH(adam_tick_kernel, 1, 32, 0, as_stream(stream), reinterpret_cast<AdamState*>(state), (double)beta1, (double)beta2);
  return check_launch("gg_adam_tick");
}
extern "C" int gg_adam_apply(const gg_adam_entry* table, const gg_adam_chunk* chunks, int n_chunks, const void* state, float lr,
                             float beta1, float beta2, float eps, float grad_scale, void* stream) {
  if (n_chunks <= 0) return GG_OK;
  GG_LAUNCH(adam_multi_kernel, n_chunks, 256, 0, as_stream(stream), table, chunks, n_chunks,
            reinterpret_cast<const AdamState*>(state), lr, beta1, beta2, eps, grad_scale);
  return check_launch("gg_adam_apply");
}

__global__ void __launch_bounds__(256) rmsprop_multi_kernel(const gg_adam_entry* __restrict__ table,
                                                            const gg_adam_chunk* __restrict__ chunks, int n_chunks, float lr,
                                                            float decay, float eps, float gscale) {
  GG_PDL_ENTRY();
  int ci = blockIdx.x;
  if (ci >= n_chunks) return;
  gg_adam_chunk ch = chunks[ci];
  gg_adam_entry e = table[ch.tensor];
  long long end = min(e.n, ch.offset + (long long)GG_ADAM_CHUNK);
  for (long long i = ch.offset + threadIdx.x; i < end; i += blockDim.x) {
    float g = e.g[i] * gscale;
    // tf.train.RMSPropOptimizer (momentum 0): ms = decay*ms + (1-decay) g^2 ; p -= lr * g / sqrt(ms + eps)
    float ms = e.m[i] + (g * g - e.m[i]) * (1.f - decay);
    e.m[i] = ms;
    e.p[i] = e.p[i] - lr * g * rsqrtf(ms + eps);
  }
}
extern "C" int gg_rmsprop_multi(const gg_adam_entry* table, const gg_adam_chunk* chunks, int n_chunks, float lr, float decay,
                                float eps, float grad_scale, void* stream) {
  if (n_chunks <= 0) return GG_OK;
  GG_LAUNCH(rmsprop_multi_kernel, n_chunks, 256, 0, as_stream(stream), table, chunks, n_chunks, lr, decay, eps, grad_scale);
  return check_launch("gg_rmsprop_multi");
}

__global__ void __launch_bounds__(256) pack_kernel(const gg_adam_entry* __restrict__ table, const gg_adam_chunk* __restrict__ chunks,
                                                   int n_chunks, const long long* __restrict__ flat_offsets,
                                                   float* __restrict__ flat, int to_flat) {
  GG_PDL_ENTRY();
  int ci = blockIdx.x;
  if (ci >= n_chunks) return;
  gg_adam_chunk ch = chunks[ci];
  gg_adam_entry e = table[ch.tensor];
  long long base = flat_offsets[ch.tensor];
  long long end = min(e.n, ch.offset + (long long)GG_ADAM_CHUNK);
  float* gw = const_cast<float*>(e.g);
  for (long long i = ch.offset + threadIdx.x; i < end; i += blockDim.x) {
    if (to_flat) flat[base + i] = e.g[i];
    else gw[i] = flat[base + i];
  }
}
extern "C" int gg_pack_grads(const gg_adam_entry* table, const gg_adam_chunk* chunks, int n_chunks,
                             const long long* flat_offsets, float* flat, int to_flat, void* stream) {
  if (n_chunks <= 0) return GG_OK;
  GG_LAUNCH(pack_kernel, n_chunks, 256, 0, as_stream(stream), table, chunks, n_chunks, flat_offsets, flat, to_flat);
  return check_launch("gg_pack_grads");
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter based RNG
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u32_to_unit(uint32_t v) {  // [0,1)
  return (float)(v >> 8) * (1.0f / 16777216.0f);
}

__global__ void rng_tick_kernel(unsigned long long* tick) {
  GG_PDL_ENTRY();
  if (threadIdx.x == 0 && blockIdx.x == 0) tick[0] += 1ull;
}
extern "C" int gg_rng_tick(void* tick_counter, void* stream) {
  GG_LAUNCH(rng_tick_kernel, 1, 32, 0, as_stream(stream), reinterpret_cast<unsigned long long*>(tick_counter));
  return check_launch("gg_rng_tick");
}

// mode 0 normal (Box-Muller on pairs), 1 uniform
__global__ void __launch_bounds__(256) rng_fill_kernel(float* __restrict__ out, long long n, int mode, float a, float b,
                                                       unsigned long long seed, uint32_t stream_id,
                                                       const unsigned long long* __restrict__ tick) {
  GG_PDL_ENTRY();
  unsigned long long t = tick ? tick[0] : 0ull;
  uint2 key = make_uint2((uint32_t)seed ^ (stream_id * 0x9E3779B9u), (uint32_t)(seed >> 32) + stream_id);
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nq = (n + 3) / 4;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; q < nq; q += stride) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)t, (uint32_t)(t >> 32)), key);
    float v[4];
    if (mode == 1) {
      v[0] = a + (b - a) * u32_to_unit(r.x);
      v[1] = a + (b - a) * u32_to_unit(r.y);
      v[2] = a + (b - a) * u32_to_unit(r.z);
      v[3] = a + (b - a) * u32_to_unit(r.w);
    } else {
      float u0 = 1.0f - u32_to_unit(r.x), u1 = u32_to_unit(r.y);  // u0 in (0,1]
      float u2 = 1.0f - u32_to_unit(r.z), u3 = u32_to_unit(r.w);
      float r0 = sqrtf(-2.f * logf(u0)), r1 = sqrtf(-2.f * logf(u2));
      float s0, c0, s1, c1;
      sincospif(2.f * u1, &s0, &c0);
      sincospif(2.f * u3, &s1, &c1);
      v[0] = a + b * r0 * c0;
      v[1] = a + b * r0 * s0;
      v[2] = a + b * r1 * c1;
      v[3] = a + b * r1 * s1;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      long long i = q * 4 + k;
      if (i < n) out[i] = v[k];
    }
  }
}
static int rng_grid(long long n) {
  long long b = (n / 4 + 255) / 256;
  if (b < 1) b = 1;
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  return (int)b;
}
extern "C" int gg_rng_normal(float* out, long long n, float mean, float stddev, uint64_t seed, uint32_t stream_id,
                             const void* tick_counter, void* stream) {
  if (n <= 0) return GG_OK;
  GG_LAUNCH(rng_fill_kernel, rng_grid(n), 256, 0, as_stream(stream), out, n, 0, mean, stddev, seed, stream_id,
                                                             reinterpret_cast<const unsigned long long*>(tick_counter));
  return check_launch("gg_rng_normal");
}
extern "C" int gg_rng_uniform(float* out, long long n, float lo, float hi, uint64_t seed, uint32_t stream_id,
                              const void* tick_counter, void* stream) {
  if (n <= 0) return GG_OK;
  GG_LAUNCH(rng_fill_kernel, rng_grid(n), 256, 0, as_stream(stream), out, n, 1, lo, hi, seed, stream_id,
                                                             reinterpret_cast<const unsigned long long*>(tick_counter));
  return check_launch("gg_rng_uniform");
}

__global__ void __launch_bounds__(256) rng_categorical_kernel(int32_t* __restrict__ idx, int n, const float* __restrict__ probs,
                                                              int K, unsigned long long seed, uint32_t stream_id,
                                                              const unsigned long long* __restrict__ tick) {
  GG_PDL_ENTRY();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long t = tick ? tick[0] : 0ull;
  uint2 key = make_uint2((uint32_t)seed ^ (stream_id * 0x9E3779B9u), (uint32_t)(seed >> 32) + stream_id);
  uint4 r = philox4x32_10(make_uint4((uint32_t)i, 0u, (uint32_t)t, (uint32_t)(t >> 32)), key);
  float total = 0.f;
  for (int k = 0; k < K; ++k) total += probs[k];
  float u = u32_to_unit(r.x) * total;
  float acc = 0.f;
  int pick = K - 1;
  for (int k = 0; k < K; ++k) {
    acc += probs[k];
    if (u < acc) { pick = k; break; }
  }
  idx[i] = pick;
}
extern "C" int gg_rng_categorical(int32_t* idx, int n, const float* probs, int K, uint64_t seed, uint32_t stream_id,
                                  const void* tick_counter, void* stream) {
  if (n <= 0) return GG_OK;
  GG_REQUIRE(K > 0, "gg_rng_categorical");
  GG_LAUNCH(rng_categorical_kernel, ceil_div(n, 256), 256, 0, as_stream(stream), idx, n, probs, K, seed, stream_id,
                                                                         reinterpret_cast<const unsigned long long*>(tick_counter));
  return check_launch("gg_rng_categorical");
}
