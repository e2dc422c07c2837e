// gg_conv_direct.cu — direct fp32 convolution kernels (any k / stride / padding / channel count), NHWC.
// They serve (1) the layers that can never be tensor-bound: the first conv (Cin=3, K=75) and the last
// deconv (Cout=3) of every model (SURVEY.md §7 "Tiny GEMMs"), which are HBM/latency-bound, and (2) as the
// on-device cross-check of the tcgen05 implicit-GEMM kernels in gg_conv_tc.cu.
// Reference: tf.nn.conv2d tflib/ops/conv2d.py:106-112; tf.nn.conv2d_transpose tflib/ops/deconv2d.py:101-107.
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
size_t conv_tc_wgrad_workspace(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo);
size_t conv_tc_workspace(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo);
void conv_tc_set_debug(void* p);
void conv_tc_set_max_ctas(int n);
void conv_tc_set_pending_mask(const float* y, int act, float alpha);
void conv_tc_set_pending_stats(float* stats, int ld);
int conv_tc_stats_tiles(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int pad_t, int pad_l, int Ho, int Wo);
// gg_conv_small.cu: one-launch shared-memory kernels for the 1-/3-channel first conv and last deconv of every network
int conv_small_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Ci, int Co, int k,
                   int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, cudaStream_t st, bool* handled);
int conv_small_dgrad(const float* dy, const float* w, const float* bias, float* dx, int B, int H, int W, int Ci, int Co, int k,
                     int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, cudaStream_t st, bool* handled);
int conv_small_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Ci, int Co, int k, int stride, int pad_t,
                     int pad_l, int Ho, int Wo, void* ws, size_t ws_bytes, cudaStream_t st, bool* handled);
size_t conv_small_wgrad_workspace(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo);
void conv_small_set_debug(void* p);
int conv_small_wgrad_info(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo, int* out8);
}  // namespace gg

namespace {

struct ConvP {
  int B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo;
};

// ---- forward: thread = (PIX consecutive wo, CV consecutive co) --------------------------------
template <int PIX, int CV>
__global__ void __launch_bounds__(128) conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ y, ConvP p, int act,
                                                       float alpha) {
  GG_PDL_ENTRY();
  int cog = (p.Co + CV - 1) / CV;           // co groups
  int wog = (p.Wo + PIX - 1) / PIX;         // wo groups
  long long total = (long long)p.B * p.Ho * wog * cog;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int cg = (int)(t % cog); t /= cog;
  int wg = (int)(t % wog); t /= wog;
  int ho = (int)(t % p.Ho);
  int b = (int)(t / p.Ho);
  int co0 = cg * CV, wo0 = wg * PIX;
  float acc[PIX][CV];
#pragma unroll
  for (int i = 0; i < PIX; ++i)
#pragma unroll
    for (int j = 0; j < CV; ++j) acc[i][j] = 0.f;
  for (int r = 0; r < p.k; ++r) {
    int hi = ho * p.stride + r - p.pad_t;
    if (hi < 0 || hi >= p.H) continue;
    for (int s = 0; s < p.k; ++s) {
      const float* wp = w + ((long long)(r * p.k + s) * p.Ci) * p.Co + co0;
      int wi0 = wo0 * p.stride + s - p.pad_l;
      const float* xrow = x + ((long long)(b * p.H + hi) * p.W) * p.Ci;
      for (int ci = 0; ci < p.Ci; ++ci) {
        float wv[CV];
        if (CV == 4 && co0 + 3 < p.Co) {
          float4 t4 = *reinterpret_cast<const float4*>(wp + (long long)ci * p.Co);
          wv[0] = t4.x; wv[1 % CV] = t4.y; wv[2 % CV] = t4.z; wv[3 % CV] = t4.w;
        } else {
#pragma unroll
          for (int j = 0; j < CV; ++j) wv[j] = (co0 + j < p.Co) ? wp[(long long)ci * p.Co + j] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < PIX; ++i) {
          int wi = wi0 + i * p.stride;
          float xv = (wi >= 0 && wi < p.W && wo0 + i < p.Wo) ? xrow[(long long)wi * p.Ci + ci] : 0.f;
#pragma unroll
          for (int j = 0; j < CV; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < PIX; ++i) {
    int wo = wo0 + i;
    if (wo >= p.Wo) continue;
    float* yp = y + ((long long)((b * p.Ho + ho) * p.Wo + wo)) * p.Co + co0;
#pragma unroll
    for (int j = 0; j < CV; ++j) {
      if (co0 + j < p.Co) {
        float v = acc[i][j] + (bias ? bias[co0 + j] : 0.f);
        yp[j] = apply_act(v, act, alpha);
      }
    }
  }
}

// ---- dgrad / transposed conv: thread = (input pixel, CV consecutive ci) ------------------------
template <int CV>
__global__ void __launch_bounds__(128) conv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ dx, ConvP p,
                                                         int act, float alpha) {
  GG_PDL_ENTRY();
  int cig = (p.Ci + CV - 1) / CV;
  long long total = (long long)p.B * p.H * p.W * cig;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int cg = (int)(t % cig); t /= cig;
  int wi = (int)(t % p.W); t /= p.W;
  int hi = (int)(t % p.H);
  int b = (int)(t / p.H);
  int ci0 = cg * CV;
  float acc[CV];
#pragma unroll
  for (int j = 0; j < CV; ++j) acc[j] = 0.f;
  bool vec = (p.Co % 4) == 0;
  for (int r = 0; r < p.k; ++r) {
    int hn = hi + p.pad_t - r;
    if (hn < 0 || (hn % p.stride) != 0) continue;
    int ho = hn / p.stride;
    if (ho >= p.Ho) continue;
    for (int s = 0; s < p.k; ++s) {
      int wn = wi + p.pad_l - s;
      if (wn < 0 || (wn % p.stride) != 0) continue;
      int wo = wn / p.stride;
      if (wo >= p.Wo) continue;
      const float* dyp = dy + ((long long)((b * p.Ho + ho) * p.Wo + wo)) * p.Co;
      const float* wp = w + ((long long)(r * p.k + s) * p.Ci + ci0) * p.Co;
      if (vec) {
        for (int co = 0; co < p.Co; co += 4) {
          float4 g = *reinterpret_cast<const float4*>(dyp + co);
#pragma unroll
          for (int j = 0; j < CV; ++j) {
            if (ci0 + j < p.Ci) {
              float4 wv = *reinterpret_cast<const float4*>(wp + (long long)j * p.Co + co);
              acc[j] = fmaf(g.x, wv.x, acc[j]);
              acc[j] = fmaf(g.y, wv.y, acc[j]);
              acc[j] = fmaf(g.z, wv.z, acc[j]);
              acc[j] = fmaf(g.w, wv.w, acc[j]);
            }
          }
        }
      } else {
        for (int co = 0; co < p.Co; ++co) {
          float g = dyp[co];
#pragma unroll
          for (int j = 0; j < CV; ++j)
            if (ci0 + j < p.Ci) acc[j] = fmaf(g, wp[(long long)j * p.Co + co], acc[j]);
        }
      }
    }
  }
  float* xp = dx + ((long long)((b * p.H + hi) * p.W + wi)) * p.Ci + ci0;
#pragma unroll
  for (int j = 0; j < CV; ++j) {
    if (ci0 + j < p.Ci) {
      float v = acc[j] + (bias ? bias[ci0 + j] : 0.f);
      xp[j] = apply_act(v, act, alpha);
    }
  }
}

// ---- wgrad: thread = (tap, ci, CV consecutive co), blockIdx.y = pixel slice; partial sums ---------
template <int CV>
__global__ void __launch_bounds__(128) conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ part, ConvP p, int slices) {
  GG_PDL_ENTRY();
  int cog = (p.Co + CV - 1) / CV;
  int total = p.k * p.k * p.Ci * cog;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  int cg = t % cog; int u = t / cog;
  int ci = u % p.Ci; u /= p.Ci;
  int s = u % p.k, r = u / p.k;
  int co0 = cg * CV;
  int npix = p.B * p.Ho * p.Wo;
  int per = (npix + slices - 1) / slices;
  int p0 = blockIdx.y * per, p1 = min(npix, p0 + per);
  float acc[CV];
#pragma unroll
  for (int j = 0; j < CV; ++j) acc[j] = 0.f;
  bool vec = (CV == 4) && (p.Co % 4 == 0);
  for (int pix = p0; pix < p1; ++pix) {
    int wo = pix % p.Wo; int q = pix / p.Wo;
    int ho = q % p.Ho; int b = q / p.Ho;
    int hi = ho * p.stride + r - p.pad_t, wi = wo * p.stride + s - p.pad_l;
    if (hi < 0 || hi >= p.H || wi < 0 || wi >= p.W) continue;
    float xv = x[((long long)((b * p.H + hi) * p.W + wi)) * p.Ci + ci];
    const float* dyp = dy + (long long)pix * p.Co + co0;
    if (vec) {
      float4 g = *reinterpret_cast<const float4*>(dyp);
      acc[0] = fmaf(xv, g.x, acc[0]); acc[1 % CV] = fmaf(xv, g.y, acc[1 % CV]);
      acc[2 % CV] = fmaf(xv, g.z, acc[2 % CV]); acc[3 % CV] = fmaf(xv, g.w, acc[3 % CV]);
    } else {
#pragma unroll
      for (int j = 0; j < CV; ++j)
        if (co0 + j < p.Co) acc[j] = fmaf(xv, dyp[j], acc[j]);
    }
  }
  long long wsz = (long long)p.k * p.k * p.Ci * p.Co;
  float* o = part + (long long)blockIdx.y * wsz + ((long long)(r * p.k + s) * p.Ci + ci) * p.Co + co0;
#pragma unroll
  for (int j = 0; j < CV; ++j)
    if (co0 + j < p.Co) o[j] = acc[j];
}

__global__ void __launch_bounds__(256) sum_slices_kernel(const float* __restrict__ part, float* __restrict__ out, long long n, int slices) {
  GG_PDL_ENTRY();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int s = 0; s < slices; ++s) a += part[(long long)s * n + i];
  out[i] = a;
}


// ---- 3-channel (Cin <= 4) layers on the tensor cores ----------------------------------------------------------------
// The first conv of every net (Cin = 1 or 3) and the last deconv (Cout = 1 or 3) have K = k*k*Cin = 25..100: too thin for
// a per-tap implicit GEMM (a 32-channel K block would be 90 % padding).  For these layers only, the patch matrix
// P[B*Ho*Wo, Kp] (Kp = k*k*Cin rounded up to 32; 6 MB at B=64) is materialised by one coalesced kernel and the three
// products run as plain GEMMs on the tcgen05 kernel:  y = P W,  dW = P^T dy,  dP = dy W^T followed by a col2im gather.
// The filter matrix is read in place: its rows beyond k*k*Cin are TMA out-of-bounds zeros.
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ x, float* __restrict__ P, ConvP p, int Kp) {
  GG_PDL_ENTRY();
  const int kv = Kp / 4;
  const long long total = (long long)p.B * p.Ho * p.Wo * kv;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int j4 = (int)(t % kv);
  long long m = t / kv;
  const int wo = (int)(m % p.Wo);
  const int ho = (int)((m / p.Wo) % p.Ho);
  const int b = (int)(m / ((long long)p.Wo * p.Ho));
  const int kreal = p.k * p.k * p.Ci;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int kk = j4 * 4 + e;
    float val = 0.f;
    if (kk < kreal) {
      const int tap = kk / p.Ci, c = kk - tap * p.Ci;
      const int r = tap / p.k, s = tap - r * p.k;
      const int hi = ho * p.stride + r - p.pad_t, wi = wo * p.stride + s - p.pad_l;
      if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) val = x[((long long)(b * p.H + hi) * p.W + wi) * p.Ci + c];
    }
    v[e] = val;
  }
  *reinterpret_cast<float4*>(P + m * Kp + j4 * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

__global__ void __launch_bounds__(256) col2im_kernel(const float* __restrict__ dP, const float* __restrict__ bias,
                                                     float* __restrict__ dx, ConvP p, int Kp, int act, float alpha) {
  GG_PDL_ENTRY();
  const long long total = (long long)p.B * p.H * p.W * p.Ci;
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % p.Ci);
  long long q = t / p.Ci;
  const int w = (int)(q % p.W);
  const int h = (int)((q / p.W) % p.H);
  const int b = (int)(q / ((long long)p.W * p.H));
  float acc = 0.f;
  for (int r = 0; r < p.k; ++r) {
    const int hn = h + p.pad_t - r;
    if (hn < 0 || hn % p.stride) continue;
    const int ho = hn / p.stride;
    if (ho >= p.Ho) continue;
    for (int s = 0; s < p.k; ++s) {
      const int wn = w + p.pad_l - s;
      if (wn < 0 || wn % p.stride) continue;
      const int wo = wn / p.stride;
      if (wo >= p.Wo) continue;
      acc += dP[((long long)(b * p.Ho + ho) * p.Wo + wo) * Kp + (r * p.k + s) * p.Ci + c];
    }
  }
  if (bias) acc += bias[c];
  dx[t] = apply_act(acc, act, alpha);
}

struct SmallCi {
  bool ok;
  int Kp;
  long long M;
  size_t tc_bytes, p_bytes;
};

SmallCi smallci_plan(int mode, const ConvP& p) {
  SmallCi sc{};
  sc.ok = false;
  if (g_conv_backend == 1) return sc;
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("GG_IM2COL"); enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!enabled) return sc;
  if (p.Ci > 4 || p.Co % 32 != 0) return sc;
  sc.Kp = ((p.k * p.k * p.Ci + 31) / 32) * 32;
  if (sc.Kp > 128) return sc;
  sc.M = (long long)p.B * p.Ho * p.Wo;
  if (sc.M % 32 != 0 || sc.M > (1 << 24)) return sc;
  size_t tc;
  if (mode == 0) tc = conv_tc_workspace(0, (int)sc.M, 1, 1, sc.Kp, p.Co, 1, 1, 1, 1);
  else if (mode == 1) tc = conv_tc_workspace(1, (int)sc.M, 1, 1, sc.Kp, p.Co, 1, 1, 1, 1);
  else tc = conv_tc_workspace(2, (int)sc.M, 1, 1, sc.Kp, p.Co, 1, 1, 1, 1);
  if (tc == 0) return sc;
  sc.tc_bytes = (tc + 255) & ~size_t(255);
  sc.p_bytes = (size_t)sc.M * sc.Kp * sizeof(float);
  sc.ok = true;
  return sc;
}

int direct_wgrad_slices(const ConvP& p) {
  long long threads = (long long)p.k * p.k * p.Ci * ((p.Co + 3) / 4);
  int blocks = ceil_div(threads, 128);
  int want = ceil_div(4 * kNumSMs, blocks);
  int npix = p.B * p.Ho * p.Wo;
  int maxs = npix / 32;
  if (maxs < 1) maxs = 1;
  int S = want < maxs ? want : maxs;
  if (S > 128) S = 128;
  if (S < 1) S = 1;
  return S;
}

int check_geom(const ConvP& p, const char* who) {
  if (p.B <= 0 || p.H <= 0 || p.W <= 0 || p.Ci <= 0 || p.Co <= 0 || p.k <= 0 || p.stride <= 0 || p.Ho <= 0 || p.Wo <= 0 ||
      p.pad_t < 0 || p.pad_l < 0)
    return fail(GG_ERR_BAD_ARG, "%s: non-positive geometry", who);
  return GG_OK;
}

}  // namespace

extern "C" int gg_conv2d_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Ci, int Co,
                             int k, int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, void* workspace,
                             size_t workspace_bytes, void* stream) {
  ConvP p{B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo};
  int rc = check_geom(p, "gg_conv2d_fwd");
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (g_conv_backend != 1) {
    bool handled = false;
    rc = conv_small_fwd(x, w, bias, y, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo, act, alpha, st, &handled);
    if (rc) return rc;
    if (handled) { g_last_backend = 0; return GG_OK; }
  }
  {
    SmallCi sc = smallci_plan(0, p);
    if (sc.ok && workspace != nullptr && workspace_bytes >= sc.tc_bytes + sc.p_bytes) {
      float* P = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + sc.tc_bytes);
      const long long total = sc.M * (sc.Kp / 4);
      GG_LAUNCH(im2col_kernel, ceil_div(total, 256), 256, 0, st, x, P, p, sc.Kp);
      rc = check_launch("gg_conv2d_fwd/im2col");
      if (rc) return rc;
      bool handled = false;
      rc = conv_tc_fwd(P, w, bias, y, (int)sc.M, 1, 1, sc.Kp, Co, 1, 1, 0, 0, 1, 1, act, alpha, workspace, sc.tc_bytes, st,
                       &handled, k * k * Ci);
      if (rc) return rc;
      if (handled) { g_last_backend = 1; return GG_OK; }
    }
  }
  if (g_conv_backend != 1) {
    bool handled = false;
    rc = conv_tc_fwd(x, w, bias, y, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo, act, alpha, workspace, workspace_bytes,
                     st, &handled);
    if (rc) return rc;
    if (handled) { g_last_backend = 1; return GG_OK; }
    if (g_conv_backend == 2) return fail(GG_ERR_UNSUPPORTED, "gg_conv2d_fwd: shape not supported by the tcgen05 path%s");
  }
  g_last_backend = 0;
  if (Co % 4 == 0) {
    long long total = (long long)B * Ho * ((Wo + 3) / 4) * (Co / 4);
    GG_LAUNCH((conv_fwd_kernel<4, 4>), ceil_div(total, 128), 128, 0, st, x, w, bias, y, p, act, alpha);
  } else {
    long long total = (long long)B * Ho * ((Wo + 3) / 4) * Co;
    GG_LAUNCH((conv_fwd_kernel<4, 1>), ceil_div(total, 128), 128, 0, st, x, w, bias, y, p, act, alpha);
  }
  return check_launch("gg_conv2d_fwd(direct)");
}

extern "C" int gg_conv2d_dgrad(const float* dy, const float* w, const float* bias, float* dx, int B, int H, int W, int Ci,
                               int Co, int k, int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha,
                               void* workspace, size_t workspace_bytes, void* stream) {
  ConvP p{B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo};
  int rc = check_geom(p, "gg_conv2d_dgrad");
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (g_conv_backend != 1) {
    bool handled = false;
    rc = conv_small_dgrad(dy, w, bias, dx, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo, act, alpha, st, &handled);
    if (rc) return rc;
    if (handled) { g_last_backend = 0; return GG_OK; }
  }
  {
    SmallCi sc = smallci_plan(1, p);
    if (sc.ok && workspace != nullptr && workspace_bytes >= sc.tc_bytes + sc.p_bytes) {
      float* dP = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + sc.tc_bytes);
      bool handled = false;
      rc = conv_tc_dgrad(dy, w, nullptr, dP, (int)sc.M, 1, 1, sc.Kp, Co, 1, 1, 0, 0, 1, 1, GG_ACT_NONE, 0.f, workspace,
                         sc.tc_bytes, st, &handled, k * k * Ci);
      if (rc) return rc;
      if (handled) {
        const long long total = (long long)B * H * W * Ci;
        GG_LAUNCH(col2im_kernel, ceil_div(total, 256), 256, 0, st, dP, bias, dx, p, sc.Kp, act, alpha);
        g_last_backend = 1;
        return check_launch("gg_conv2d_dgrad/col2im");
      }
    }
  }
  if (g_conv_backend != 1) {
    bool handled = false;
    rc = conv_tc_dgrad(dy, w, bias, dx, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo, act, alpha, workspace,
                       workspace_bytes, st, &handled);
    if (rc) return rc;
    if (handled) { g_last_backend = 1; return GG_OK; }
    if (g_conv_backend == 2) return fail(GG_ERR_UNSUPPORTED, "gg_conv2d_dgrad: shape not supported by the tcgen05 path%s");
  }
  g_last_backend = 0;
  if (Ci % 4 == 0) {
    long long total = (long long)B * H * W * (Ci / 4);
    GG_LAUNCH((conv_dgrad_kernel<4>), ceil_div(total, 128), 128, 0, st, dy, w, bias, dx, p, act, alpha);
  } else {
    long long total = (long long)B * H * W * Ci;
    GG_LAUNCH((conv_dgrad_kernel<1>), ceil_div(total, 128), 128, 0, st, dy, w, bias, dx, p, act, alpha);
  }
  return check_launch("gg_conv2d_dgrad(direct)");
}

// dgrad followed by the gradient of the activation that produced the dgrad's target: dx = act'(y) * (dy (*) w^T), one launch.
// Tensor-core path only (the write-out of gg_conv_tc.cu applies the mask); gg_conv2d_tc_supported tells the caller in advance.
extern "C" int gg_conv2d_tc_supported(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo) {
  if (g_conv_backend == 1) return 0;
  return conv_tc_workspace(mode, B, H, W, Ci, Co, k, stride, Ho, Wo) != 0 ? 1 : 0;
}

extern "C" int gg_conv2d_dgrad_actgrad(const float* dy, const float* w, float* dx, const float* y_fwd, int act, float alpha, int B,
                                       int H, int W, int Ci, int Co, int k, int stride, int pad_t, int pad_l, int Ho, int Wo,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  ConvP p{B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo};
  int rc = check_geom(p, "gg_conv2d_dgrad_actgrad");
  if (rc) return rc;
  if (y_fwd == nullptr) return fail(GG_ERR_BAD_ARG, "gg_conv2d_dgrad_actgrad: y_fwd is required%s");
  if (g_conv_backend == 1) return fail(GG_ERR_UNSUPPORTED, "gg_conv2d_dgrad_actgrad: tensor-core path disabled (backend 1)%s");
  bool handled = false;
  conv_tc_set_pending_mask(y_fwd, act, alpha);
  rc = conv_tc_dgrad(dy, w, nullptr, dx, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo, GG_ACT_NONE, 0.f, workspace,
                     workspace_bytes, as_stream(stream), &handled);
  conv_tc_set_pending_mask(nullptr, 0, 0.f);
  if (rc) return rc;
  if (!handled) return fail(GG_ERR_UNSUPPORTED, "gg_conv2d_dgrad_actgrad: shape not supported by the tcgen05 path%s");
  g_last_backend = 1;
  return GG_OK;
}

extern "C" size_t gg_conv2d_wgrad_workspace(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo) {
  ConvP p{B, H, W, Ci, Co, k, stride, 0, 0, Ho, Wo};
  size_t direct = (size_t)direct_wgrad_slices(p) * k * k * Ci * Co * sizeof(float);
  size_t tc = conv_tc_wgrad_workspace(B, H, W, Ci, Co, k, stride, Ho, Wo);
  SmallCi sc = smallci_plan(2, p);
  if (sc.ok && sc.tc_bytes + sc.p_bytes > tc) tc = sc.tc_bytes + sc.p_bytes;
  size_t small = conv_small_wgrad_workspace(B, H, W, Ci, Co, k, stride, Ho, Wo);
  if (small > tc) tc = small;
  return direct > tc ? direct : tc;
}

// conv / deconv / dense launch that also writes the batch-norm statistics of its output (include/gg_b200.h)
extern "C" int gg_conv2d_stats_tiles(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int pad_t, int pad_l, int Ho,
                                     int Wo) {
  if (g_conv_backend == 1) return 0;
  return conv_tc_stats_tiles(mode, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
}

extern "C" int gg_conv2d_bnstats(int mode, const float* a, const float* w, const float* bias, float* out, float* stats, int B, int H,
                                 int W, int Ci, int Co, int k, int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  ConvP p{B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo};
  int rc = check_geom(p, "gg_conv2d_bnstats");
  if (rc) return rc;
  if (mode != 0 && mode != 1) return fail(GG_ERR_BAD_ARG, "gg_conv2d_bnstats: mode must be 0 (fwd) or 1 (dgrad)%s");
  if (stats == nullptr) return fail(GG_ERR_BAD_ARG, "gg_conv2d_bnstats: stats is required%s");
  if (g_conv_backend == 1) return fail(GG_ERR_UNSUPPORTED, "gg_conv2d_bnstats: tensor-core path disabled (backend 1)%s");
  bool handled = false;
  conv_tc_set_pending_stats(stats, mode == 0 ? Co : Ci);
  if (mode == 0)
    rc = conv_tc_fwd(a, w, bias, out, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo, act, alpha, workspace, workspace_bytes,
                     as_stream(stream), &handled);
  else
    rc = conv_tc_dgrad(a, w, bias, out, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo, act, alpha, workspace, workspace_bytes,
                       as_stream(stream), &handled);
  conv_tc_set_pending_stats(nullptr, 0);
  if (rc) return rc;
  if (!handled) return fail(GG_ERR_UNSUPPORTED, "gg_conv2d_bnstats: shape not supported by the tcgen05 path%s");
  g_last_backend = 1;
  return GG_OK;
}

namespace gg { void conv_tc_set_stage_cap(int n); void conv_tc_last_info(int* out8); }
extern "C" int gg_last_tc_info(int* out8) {
  if (!out8) return fail(GG_ERR_BAD_ARG, "gg_last_tc_info: null output%s");
  gg::conv_tc_last_info(out8);
  return GG_OK;
}
extern "C" int gg_set_tc_stages(int n) {
  gg::conv_tc_set_stage_cap(n);
  return GG_OK;
}

extern "C" int gg_set_tc_max_ctas(int n) {
  conv_tc_set_max_ctas(n);
  return GG_OK;
}

extern "C" int gg_debug_set_buffer(void* device_buffer_256_int64) {
  conv_tc_set_debug(device_buffer_256_int64);
  return GG_OK;
}

extern "C" int gg_debug_set_small_buffer(void* device_buffer_int64) {
  conv_small_set_debug(device_buffer_int64);
  return GG_OK;
}

extern "C" int gg_debug_small_wgrad_info(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo, int* out8) {
  if (!out8) return fail(GG_ERR_BAD_ARG, "gg_debug_small_wgrad_info: null output%s");
  return conv_small_wgrad_info(B, H, W, Ci, Co, k, stride, Ho, Wo, out8);
}

extern "C" size_t gg_conv2d_workspace(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo) {
  if (mode == 2) return gg_conv2d_wgrad_workspace(B, H, W, Ci, Co, k, stride, Ho, Wo);
  size_t tc = conv_tc_workspace(mode, B, H, W, Ci, Co, k, stride, Ho, Wo);
  ConvP p{B, H, W, Ci, Co, k, stride, 0, 0, Ho, Wo};
  SmallCi sc = smallci_plan(mode, p);
  if (sc.ok && sc.tc_bytes + sc.p_bytes > tc) tc = sc.tc_bytes + sc.p_bytes;
  return tc > 256 ? tc : 256;
}

extern "C" int gg_conv2d_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Ci, int Co, int k,
                               int stride, int pad_t, int pad_l, int Ho, int Wo, void* workspace, size_t workspace_bytes,
                               void* stream) {
  ConvP p{B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo};
  int rc = check_geom(p, "gg_conv2d_wgrad");
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (g_conv_backend != 1) {
    bool handled = false;
    rc = conv_small_wgrad(x, dy, dw, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo, workspace, workspace_bytes, st, &handled);
    if (rc) return rc;
    if (handled) { g_last_backend = 0; return GG_OK; }
  }
  {
    SmallCi sc = smallci_plan(2, p);
    if (sc.ok && workspace != nullptr && workspace_bytes >= sc.tc_bytes + sc.p_bytes) {
      float* P = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + sc.tc_bytes);
      const long long total = sc.M * (sc.Kp / 4);
      GG_LAUNCH(im2col_kernel, ceil_div(total, 256), 256, 0, st, x, P, p, sc.Kp);
      rc = check_launch("gg_conv2d_wgrad/im2col");
      if (rc) return rc;
      bool handled = false;
      rc = conv_tc_wgrad(P, dy, dw, (int)sc.M, 1, 1, sc.Kp, Co, 1, 1, 0, 0, 1, 1, workspace, sc.tc_bytes, st, &handled,
                         k * k * Ci);
      if (rc) return rc;
      if (handled) { g_last_backend = 1; return GG_OK; }
    }
  }
  if (g_conv_backend != 1) {
    bool handled = false;
    rc = conv_tc_wgrad(x, dy, dw, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo, workspace, workspace_bytes, st, &handled);
    if (rc) return rc;
    if (handled) { g_last_backend = 1; return GG_OK; }
    if (g_conv_backend == 2) return fail(GG_ERR_UNSUPPORTED, "gg_conv2d_wgrad: shape not supported by the tcgen05 path%s");
  }
  g_last_backend = 0;
  int S = direct_wgrad_slices(p);
  long long wsz = (long long)k * k * Ci * Co;
  if (workspace == nullptr || workspace_bytes < (size_t)S * wsz * sizeof(float))
    return fail(GG_ERR_WORKSPACE, "gg_conv2d_wgrad: workspace too small (need %s%lld bytes)", "", (long long)S * wsz * 4);
  float* part = reinterpret_cast<float*>(workspace);
  if (Co % 4 == 0) {
    int threads = k * k * Ci * (Co / 4);
    dim3 grid(ceil_div(threads, 128), S);
    GG_LAUNCH((conv_wgrad_kernel<4>), grid, 128, 0, st, x, dy, part, p, S);
  } else {
    int threads = k * k * Ci * Co;
    dim3 grid(ceil_div(threads, 128), S);
    GG_LAUNCH((conv_wgrad_kernel<1>), grid, 128, 0, st, x, dy, part, p, S);
  }
  rc = check_launch("gg_conv2d_wgrad(direct)");
  if (rc) return rc;
  GG_LAUNCH(sum_slices_kernel, ceil_div(wsz, 256), 256, 0, st, part, dw, wsz, S);
  return check_launch("gg_conv2d_wgrad(direct)/sum");
}
