// gg_gemm.cu — dense layers (tf.matmul + bias, tflib/ops/linear.py:132-146, and the two transposed products
// autodiff adds).  The Linear layers of the hot path have M = batch = 64: they are weight-bandwidth / latency
// bound, so this is an fp32 FFMA kernel with split-K sized to fill the 148 SMs and a
// fused bias+activation epilogue, not a tensor-core kernel.
#include "gg_common.cuh"

using namespace gg;

namespace gg {
int conv_tc_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Ci, int Co, int k,
                int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, void* ws, size_t ws_bytes,
                cudaStream_t st, bool* handled, int filt_rows = 0);
int conv_tc_dgrad(const float* dy, const float* w, const float* bias, float* dx, int B, int H, int W, int Ci, int Co, int k,
                  int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, void* ws, size_t ws_bytes,
                  cudaStream_t st, bool* handled, int filt_rows = 0);
int conv_tc_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Ci, int Co, int k, int stride,
                  int pad_t, int pad_l, int Ho, int Wo, void* ws, size_t ws_bytes, cudaStream_t st, bool* handled, int out_rows = 0);
size_t conv_tc_workspace(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo);
}  // namespace gg

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <int TA, int TB>
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                   const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K,
                                                   int k_per_split, int splits, int act, float alpha) {
  GG_PDL_ENTRY();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  int split = blockIdx.z;
  int kb = split * k_per_split, ke = min(K, kb + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kb; k0 < ke; k0 += TK) {
    // load A tile: As[kk][mm] = opA[m0+mm, k0+kk]
#pragma unroll
    for (int it = 0; it < (TM * TK) / 256; ++it) {
      int e = threadIdx.x + it * 256;
      int mm, kk;
      if (TA) { mm = e % TM; kk = e / TM; }  // stored [K,M]: m fastest
      else { kk = e % TK; mm = e / TK; }     // stored [M,K]: k fastest
      int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < ke) v = TA ? A[(long long)k * M + m] : A[(long long)m * K + k];
      As[kk][mm] = v;
    }
#pragma unroll
    for (int it = 0; it < (TN * TK) / 256; ++it) {
      int e = threadIdx.x + it * 256;
      int nn, kk;
      if (TB) { kk = e % TK; nn = e / TK; }  // stored [N,K]: k fastest
      else { nn = e % TN; kk = e / TN; }     // stored [K,N]: n fastest
      int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < N && k < ke) v = TB ? B[(long long)n * K + k] : B[(long long)k * N + n];
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = C;
  if (splits > 1) out = C + (long long)split * M * N;  // C is the workspace here
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (splits == 1) {
        if (bias) v += bias[n];
        v = apply_act(v, act, alpha);
      }
      out[(long long)m * N + n] = v;
    }
  }
}

__global__ void __launch_bounds__(256) gemm_splitk_finish(const float* __restrict__ ws, const float* __restrict__ bias,
                                                          float* __restrict__ C, long long MN, int N, int splits, int act,
                                                          float alpha) {
  GG_PDL_ENTRY();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += ws[(long long)s * MN + i];
  if (bias) v += bias[i % N];
  C[i] = apply_act(v, act, alpha);
}

// ---- the thin products of the hot path --------------------------------------------------------------------------------------
// The critic's last layer and its gradients ([128 x 512] @ [512 x 1], its outer-product dgrad with K = 1, the [512 x 1] wgrad) and
// the mixture-prior lookups ([64 x 30] @ [30 x 128] and its transpose) are 10^4..10^5 MACs: the tiled split-K kernel above spends
// 8-12 us on them (two launches, 16 CTAs, a shared-memory round trip per 16-wide K tile; profiles/launches_r2_eager.csv) and four
// of them sit on the step's critical path.  They get one flat launch each:
//   gemm_rowdot_kernel   <= 8192 outputs and K >= 32: one warp per output element, lanes stride K, shuffle reduction (a thread
//                        per output with a 64-128-long dependent-latency loop measured 13-27 us for [512 x 1] = [128 x 512]^T [128 x 1]);
//   gemm_smallk_kernel   K < 32 and <= 2^21 MACs: one thread per output element, sequential fp32 FMA over K.
template <int TA, int TB>
__global__ void __launch_bounds__(256) gemm_rowdot_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K,
                                                          int act, float alpha) {
  GG_PDL_ENTRY();
  const int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= M * N) return;
  const int m = w / N, n = w % N;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32)
    acc = fmaf(TA ? A[(long long)k * M + m] : A[(long long)m * K + k], TB ? B[(long long)n * K + k] : B[(long long)k * N + n], acc);
  acc = warp_sum(acc);
  if (lane == 0) C[(long long)m * N + n] = apply_act(acc + (bias ? bias[n] : 0.f), act, alpha);
}

template <int TA, int TB>
__global__ void __launch_bounds__(256) gemm_smallk_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K,
                                                          int act, float alpha) {
  GG_PDL_ENTRY();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float acc = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float a = TA ? A[(long long)k * M + m] : A[(long long)m * K + k];
    const float b = TB ? B[(long long)n * K + k] : B[(long long)k * N + n];
    acc = fmaf(a, b, acc);
  }
  C[i] = apply_act(acc + (bias ? bias[n] : 0.f), act, alpha);
}

int choose_splits(int M, int N, int K) {
  int tiles = ceil_div(M, TM) * ceil_div(N, TN);
  int splits = ceil_div(2 * kNumSMs, tiles);
  int maxs = K / 64;
  if (maxs < 1) maxs = 1;
  if (splits > maxs) splits = maxs;
  if (splits > 32) splits = 32;
  if (splits < 1) splits = 1;
  return splits;
}

}  // namespace

// A dense layer is a 1x1 convolution over a 1x1 image: the three products of a Linear layer (y = xW, dx = dy W^T,
// dW = x^T dy) map onto the fwd / dgrad / wgrad modes of the tcgen05 implicit-GEMM kernel (gg_conv_tc.cu).
static size_t tc_ws(int M, int N, int K) {
  size_t a = conv_tc_workspace(0, M, 1, 1, K, N, 1, 1, 1, 1);
  size_t b = conv_tc_workspace(1, M, 1, 1, N, K, 1, 1, 1, 1);
  size_t c = conv_tc_workspace(2, K, 1, 1, M, N, 1, 1, 1, 1);
  size_t m = a > b ? a : b;
  return m > c ? m : c;
}

extern "C" size_t gg_gemm_workspace(int M, int N, int K) {
  int s = choose_splits(M, N, K);
  size_t simt = s > 1 ? (size_t)s * M * N * sizeof(float) : 0;
  size_t tc = tc_ws(M, N, K);
  return simt > tc ? simt : tc;
}

extern "C" int gg_gemm(const float* A, const float* Bm, const float* bias, float* C, int M, int N, int K, int ta, int tb,
                       int act, float alpha, void* workspace, size_t workspace_bytes, void* stream) {
  if (M <= 0 || N <= 0) return GG_OK;
  GG_REQUIRE(K > 0, "gg_gemm");
  cudaStream_t st = as_stream(stream);
  if (g_conv_backend != 1 && workspace != nullptr) {
    bool handled = false;
    int rc = GG_OK;
    if (!ta && !tb) rc = conv_tc_fwd(A, Bm, bias, C, M, 1, 1, K, N, 1, 1, 0, 0, 1, 1, act, alpha, workspace, workspace_bytes, st, &handled);
    else if (!ta && tb) rc = conv_tc_dgrad(A, Bm, bias, C, M, 1, 1, N, K, 1, 1, 0, 0, 1, 1, act, alpha, workspace, workspace_bytes, st, &handled);
    else if (ta && !tb && bias == nullptr && act == GG_ACT_NONE)
      rc = conv_tc_wgrad(A, Bm, C, K, 1, 1, M, N, 1, 1, 0, 0, 1, 1, workspace, workspace_bytes, st, &handled);
    if (rc) return rc;
    if (handled) { g_last_backend = 1; return GG_OK; }
  }
  static int thin = -1;
  if (thin < 0) { const char* e = getenv("GG_GEMM_THIN"); thin = (e && e[0] == '0') ? 0 : 1; }
  if (thin && (long long)M * N <= 8192 && K >= 32) {
    g_last_backend = 0;
    const int blocks = ceil_div((long long)M * N * 32, 256);
#define GG_LAUNCH_ROWDOT(TA, TB) GG_LAUNCH((gemm_rowdot_kernel<TA, TB>), blocks, 256, 0, st, A, Bm, bias, C, M, N, K, act, alpha)
    if (!ta && !tb) GG_LAUNCH_ROWDOT(0, 0);
    else if (ta && !tb) GG_LAUNCH_ROWDOT(1, 0);
    else if (!ta && tb) GG_LAUNCH_ROWDOT(0, 1);
    else GG_LAUNCH_ROWDOT(1, 1);
#undef GG_LAUNCH_ROWDOT
    return check_launch("gg_gemm(rowdot)");
  }
  if (thin && K < 32 && (long long)M * N * K <= (1 << 21)) {
    g_last_backend = 0;
    const int blocks = ceil_div((long long)M * N, 256);
#define GG_LAUNCH_SMALLK(TA, TB) GG_LAUNCH((gemm_smallk_kernel<TA, TB>), blocks, 256, 0, st, A, Bm, bias, C, M, N, K, act, alpha)
    if (!ta && !tb) GG_LAUNCH_SMALLK(0, 0);
    else if (ta && !tb) GG_LAUNCH_SMALLK(1, 0);
    else if (!ta && tb) GG_LAUNCH_SMALLK(0, 1);
    else GG_LAUNCH_SMALLK(1, 1);
#undef GG_LAUNCH_SMALLK
    return check_launch("gg_gemm(small-K)");
  }
  int splits = choose_splits(M, N, K);
  if (splits > 1 && (workspace == nullptr || workspace_bytes < (size_t)splits * M * N * sizeof(float))) splits = 1;
  int k_per = ceil_div(ceil_div(K, splits), TK) * TK;
  splits = ceil_div(K, k_per);
  dim3 grid(ceil_div(N, TN), ceil_div(M, TM), splits);
  float* out = splits > 1 ? reinterpret_cast<float*>(workspace) : C;
  g_last_backend = 0;
#define GG_LAUNCH_GEMM(TA, TB) \
  GG_LAUNCH((gemm_kernel<TA, TB>), grid, 256, 0, st, A, Bm, bias, out, M, N, K, k_per, splits, act, alpha)
  if (!ta && !tb) GG_LAUNCH_GEMM(0, 0);
  else if (ta && !tb) GG_LAUNCH_GEMM(1, 0);
  else if (!ta && tb) GG_LAUNCH_GEMM(0, 1);
  else GG_LAUNCH_GEMM(1, 1);
#undef GG_LAUNCH_GEMM
  int rc = check_launch("gg_gemm");
  if (rc) return rc;
  if (splits > 1) {
    long long MN = (long long)M * N;
    GG_LAUNCH(gemm_splitk_finish, ceil_div(MN, 256), 256, 0, st, out, bias, C, MN, N, splits, act, alpha);
    rc = check_launch("gg_gemm/finish");
  }
  return rc;
}
