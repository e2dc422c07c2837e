// gg_conv_small.cu — shared-memory-tiled fp32 kernels for the 1- / 3-channel ends of every network: the first Conv2D of the
// Extractor / Discriminator (Cin = 1 or 3 -> 64; tflib/ops/conv2d.py:106) and the last Deconv2D of the Generator (64 -> Cout = 1
// or 3; tflib/ops/deconv2d.py:101, = the input-gradient of a conv with Cin = 1 or 3), plus that gradient itself (dgrad of
// Discriminator.1 towards fake_x).
//
// K = k*k*Cin = 25..100 is too thin for the per-tap implicit GEMM (a 32-channel K block would be 90 % padding).  These layers
// previously ran as patch-matrix kernel + tcgen05 GEMM (+ gather kernel): 2 launches and 25-38 us each, four to seven times
// per iteration on the step's dependency chain (profiles/timeline_gen_r1.txt).  They are 0.16-0.31 GFLOP against 3-12 MB of
// activations: HBM/L2- and latency-bound, never tensor-bound, so they get ONE CUDA-core launch each:
//   forward  block = 8x8 output pixels x 64 output channels; filter slab (<= 100 x 64 floats) and the input patch in shared
//            memory; thread = 4 pixels x 4 channels (16 accumulators), weights read as broadcast float4.
//   dgrad    block = 8x16 output pixels of one image; the dy patch (all Cout channels) and the whole filter in shared memory;
//            one warp per stride-parity class, so a warp walks one tap list and reads each weight as a broadcast.
// Arithmetic is plain fp32 FMA (exact parity class of the direct kernels in gg_conv_direct.cu: 1e-6 vs the oracle).
#include "gg_common.cuh"

using namespace gg;

namespace {

struct SmallP {
  int B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo;
};

constexpr int kTile = 8;          // forward: 8 x 8 output pixels per block
constexpr int kChunk = 64;        // forward: output channels per block

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
template <int CI>
__global__ void __launch_bounds__(256) conv_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ y, SmallP p,
                                                             int act, float alpha) {
  GG_PDL_ENTRY();
  extern __shared__ __align__(16) float smem_f[];
  const int K = p.k * p.k * CI;
  const int IW = (kTile - 1) * p.stride + p.k;          // input patch extent (square)
  float* sw = smem_f;                                   // [K][64]
  float* sx = smem_f + K * kChunk;                      // [IW][IW][CI]
  const int tiles_w = (p.Wo + kTile - 1) / kTile;
  const int ho0 = (blockIdx.x / tiles_w) * kTile, wo0 = (blockIdx.x % tiles_w) * kTile;
  const int b = blockIdx.y;
  const int co0 = blockIdx.z * kChunk;
  const int nco = min(kChunk, p.Co - co0);              // multiple of 4
  const int tid = threadIdx.x;

  // filter slab: rows kk = (r*k+s)*CI+c of Co contiguous floats -> sw[kk][0..nco)
  // (the fill loops keep 4 independent global loads in flight per thread: one load per iteration exposed an L2 round trip
  //  per iteration, ~5 us of a 10 us kernel)
  for (int base = tid; base < K * (kChunk / 4); base += 4 * 256) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * 256;
      const int kk = i / (kChunk / 4), q = i % (kChunk / 4);
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < K * (kChunk / 4) && q * 4 < nco) v[u] = *reinterpret_cast<const float4*>(w + (size_t)kk * p.Co + co0 + q * 4);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * 256;
      if (i < K * (kChunk / 4)) *reinterpret_cast<float4*>(sw + (i / (kChunk / 4)) * kChunk + (i % (kChunk / 4)) * 4) = v[u];
    }
  }
  // input patch with TF SAME padding as zeros
  const int hi0 = ho0 * p.stride - p.pad_t, wi0 = wo0 * p.stride - p.pad_l;
  const float* xb = x + (size_t)b * p.H * p.W * CI;
  for (int base = tid; base < IW * IW * CI; base += 4 * 256) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * 256;
      const int c = i % CI, iw = (i / CI) % IW, ih = i / (CI * IW);
      const int hi = hi0 + ih, wi = wi0 + iw;
      v[u] = 0.f;
      if (i < IW * IW * CI && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) v[u] = xb[((size_t)hi * p.W + wi) * CI + c];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * 256;
      if (i < IW * IW * CI) sx[i] = v[u];
    }
  }
  __syncthreads();

  const int cq = tid & 15, pl = tid >> 4;               // channel quad (4 channels), pixel lane (pixels pl, pl+16, pl+32, pl+48)
  int xo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pix = pl + 16 * j, py = pix / kTile, px = pix % kTile;
    xo[j] = ((py * p.stride) * IW + px * p.stride) * CI;
  }
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  const float* swq = sw + cq * 4;
  for (int r = 0; r < p.k; ++r) {
    for (int s = 0; s < p.k; ++s) {
      const int xoff = (r * IW + s) * CI;
      const int kk0 = (r * p.k + s) * CI;
#pragma unroll
      for (int c = 0; c < CI; ++c) {
        const float4 wv = *reinterpret_cast<const float4*>(swq + (kk0 + c) * kChunk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xv = sx[xo[j] + xoff + c];
          acc[j][0] = fmaf(xv, wv.x, acc[j][0]);
          acc[j][1] = fmaf(xv, wv.y, acc[j][1]);
          acc[j][2] = fmaf(xv, wv.z, acc[j][2]);
          acc[j][3] = fmaf(xv, wv.w, acc[j][3]);
        }
      }
    }
  }
  if (cq * 4 >= nco) return;
  float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bb = *reinterpret_cast<const float4*>(bias + co0 + cq * 4);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pix = pl + 16 * j, ho = ho0 + pix / kTile, wo = wo0 + pix % kTile;
    if (ho < p.Ho && wo < p.Wo) {
      float4 o;
      o.x = apply_act(acc[j][0] + bb.x, act, alpha);
      o.y = apply_act(acc[j][1] + bb.y, act, alpha);
      o.z = apply_act(acc[j][2] + bb.z, act, alpha);
      o.w = apply_act(acc[j][3] + bb.w, act, alpha);
      *reinterpret_cast<float4*>(y + (((size_t)b * p.Ho + ho) * p.Wo + wo) * p.Co + co0 + cq * 4) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward, variant 2 (GG_CONV_SMALL_V2=1; NOT yet validated on a GPU — written after the round's GPU budget was spent, index
// arithmetic pinned by tests/test_cpu_small_conv.py): 4 pixels x 8 channels per thread (32 accumulators) and the k x k tap
// loop unrolled for k = 5, i.e. 2 weight LDS.128 + 4 input LDS.32 per 32 FMAs (73 % FMA density instead of 46 %).
// Block = 128 threads = 8 channel octets x 16 pixel lanes over the same 8x8-pixel x 64-channel tile and smem layout.
// ---------------------------------------------------------------------------------------------------------------------
template <int CI, int KS>
__global__ void __launch_bounds__(128) conv_small_fwd_v2_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ y, SmallP p,
                                                                int act, float alpha) {
  GG_PDL_ENTRY();
  extern __shared__ __align__(16) float smem_f[];
  const int k = KS > 0 ? KS : p.k;
  const int K = k * k * CI;
  const int IW = (kTile - 1) * p.stride + k;
  float* sw = smem_f;                                   // [K][64]
  float* sx = smem_f + K * kChunk;                      // [IW][IW][CI]
  const int tiles_w = (p.Wo + kTile - 1) / kTile;
  const int ho0 = (blockIdx.x / tiles_w) * kTile, wo0 = (blockIdx.x % tiles_w) * kTile;
  const int b = blockIdx.y;
  const int co0 = blockIdx.z * kChunk;
  const int nco = min(kChunk, p.Co - co0);
  const int tid = threadIdx.x;
  for (int base = tid; base < K * (kChunk / 4); base += 4 * 128) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * 128;
      const int kk = i / (kChunk / 4), q = i % (kChunk / 4);
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < K * (kChunk / 4) && q * 4 < nco) v[u] = *reinterpret_cast<const float4*>(w + (size_t)kk * p.Co + co0 + q * 4);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * 128;
      if (i < K * (kChunk / 4)) *reinterpret_cast<float4*>(sw + (i / (kChunk / 4)) * kChunk + (i % (kChunk / 4)) * 4) = v[u];
    }
  }
  const int hi0 = ho0 * p.stride - p.pad_t, wi0 = wo0 * p.stride - p.pad_l;
  const float* xb = x + (size_t)b * p.H * p.W * CI;
  for (int i = tid; i < IW * IW * CI; i += 128) {
    const int c = i % CI, iw = (i / CI) % IW, ih = i / (CI * IW);
    const int hi = hi0 + ih, wi = wi0 + iw;
    float v = 0.f;
    if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) v = xb[((size_t)hi * p.W + wi) * CI + c];
    sx[i] = v;
  }
  __syncthreads();

  const int co8 = (tid & 7) * 8, pl = tid >> 3;         // channel octet, pixel lane (pixels pl, pl+16, pl+32, pl+48)
  int xo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pix = pl + 16 * j, py = pix / kTile, px = pix % kTile;
    xo[j] = ((py * p.stride) * IW + px * p.stride) * CI;
  }
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  const float* swq = sw + co8;
#pragma unroll
  for (int r = 0; r < (KS > 0 ? KS : 1); ++r) {
    for (int rr = (KS > 0 ? r : 0); rr < (KS > 0 ? r + 1 : k); ++rr) {        // KS == 0: runtime loop over all rows
#pragma unroll
      for (int s = 0; s < (KS > 0 ? KS : 1); ++s) {
        for (int ss = (KS > 0 ? s : 0); ss < (KS > 0 ? s + 1 : k); ++ss) {
          const int xoff = (rr * IW + ss) * CI;
          const int kk0 = (rr * k + ss) * CI;
#pragma unroll
          for (int c = 0; c < CI; ++c) {
            const float4 w0 = *reinterpret_cast<const float4*>(swq + (kk0 + c) * kChunk);
            const float4 w1 = *reinterpret_cast<const float4*>(swq + (kk0 + c) * kChunk + 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float xv = sx[xo[j] + xoff + c];
              acc[j][0] = fmaf(xv, w0.x, acc[j][0]); acc[j][1] = fmaf(xv, w0.y, acc[j][1]);
              acc[j][2] = fmaf(xv, w0.z, acc[j][2]); acc[j][3] = fmaf(xv, w0.w, acc[j][3]);
              acc[j][4] = fmaf(xv, w1.x, acc[j][4]); acc[j][5] = fmaf(xv, w1.y, acc[j][5]);
              acc[j][6] = fmaf(xv, w1.z, acc[j][6]); acc[j][7] = fmaf(xv, w1.w, acc[j][7]);
            }
          }
        }
      }
    }
  }
  if (co8 >= nco) return;                               // nco is a multiple of 4: the second quad of an octet may be outside
  const bool hi_ok = co8 + 4 < nco;
  float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
  if (bias) {
    b0 = *reinterpret_cast<const float4*>(bias + co0 + co8);
    if (hi_ok) b1 = *reinterpret_cast<const float4*>(bias + co0 + co8 + 4);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pix = pl + 16 * j, ho = ho0 + pix / kTile, wo = wo0 + pix % kTile;
    if (ho < p.Ho && wo < p.Wo) {
      float* o = y + (((size_t)b * p.Ho + ho) * p.Wo + wo) * p.Co + co0 + co8;
      float4 v0, v1;
      v0.x = apply_act(acc[j][0] + b0.x, act, alpha); v0.y = apply_act(acc[j][1] + b0.y, act, alpha);
      v0.z = apply_act(acc[j][2] + b0.z, act, alpha); v0.w = apply_act(acc[j][3] + b0.w, act, alpha);
      *reinterpret_cast<float4*>(o) = v0;
      if (hi_ok) {
        v1.x = apply_act(acc[j][4] + b1.x, act, alpha); v1.y = apply_act(acc[j][5] + b1.y, act, alpha);
        v1.z = apply_act(acc[j][6] + b1.z, act, alpha); v1.w = apply_act(acc[j][7] + b1.w, act, alpha);
        *reinterpret_cast<float4*>(o + 4) = v1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dgrad (and Deconv2D forward) towards CI <= 4 channels
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDH = 8, kDW = 16;   // output pixels per block: 8 rows x 16 columns = 128 threads

__host__ __device__ inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int CI>
__global__ void __launch_bounds__(128) conv_small_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ dx, SmallP p,
                                                               int PH, int PW, int act, float alpha) {
  GG_PDL_ENTRY();
  extern __shared__ __align__(16) float smem_f[];
  const int pitch = p.Co + 4;                           // floats per dy pixel / per filter row: +4 keeps float4 reads of
                                                        // neighbouring pixels on different banks
  const int KK = p.k * p.k * CI;
  float* sw = smem_f;                                   // [k*k*CI][pitch]
  float* sdy = smem_f + KK * pitch;                     // [PH][PW][pitch]
  const int tiles_w = (p.W + kDW - 1) / kDW;
  const int h0 = (blockIdx.x / tiles_w) * kDH, w0 = (blockIdx.x % tiles_w) * kDW;
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  const int s = p.stride;
  const int hy_min = floor_div(h0 + p.pad_t - (p.k - 1), s), wx_min = floor_div(w0 + p.pad_l - (p.k - 1), s);
  const int cq = p.Co / 4;

  const float* dyb = dy + (size_t)b * p.Ho * p.Wo * p.Co;
  if (128 % cq == 0) {
    // division-free fills: a thread keeps one float4 column q and walks rows / patch pixels with a fixed step (the generic
    // form below spends 4 integer divisions per element — as many instructions as the FMA loop itself)
    const int q4 = (tid % cq) * 4, first = tid / cq, step = 128 / cq;
    for (int row = first; row < KK; row += 4 * step) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = row + u * step;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr < KK) v[u] = *reinterpret_cast<const float4*>(w + (size_t)rr * p.Co + q4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = row + u * step;
        if (rr < KK) *reinterpret_cast<float4*>(sw + rr * pitch + q4) = v[u];
      }
    }
    const int npix = PH * PW;
    int ph_ = first / PW, pw_ = first % PW;
    for (int pix = first; pix < npix; pix += 4 * step) {
      float4 v[4];
      int pp = pix, hh = ph_, ww = pw_;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int hy = hy_min + hh, wx = wx_min + ww;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pp < npix && hy >= 0 && hy < p.Ho && wx >= 0 && wx < p.Wo)
          v[u] = *reinterpret_cast<const float4*>(dyb + ((size_t)hy * p.Wo + wx) * p.Co + q4);
        pp += step; ww += step;
        while (ww >= PW) { ww -= PW; ++hh; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pu = pix + u * step;
        if (pu < npix) *reinterpret_cast<float4*>(sdy + pu * pitch + q4) = v[u];
      }
      ph_ = hh; pw_ = ww;
    }
  } else {
    for (int i = tid; i < KK * cq; i += 128) {
      const int row = i / cq, q = i % cq;
      *reinterpret_cast<float4*>(sw + row * pitch + q * 4) = *reinterpret_cast<const float4*>(w + (size_t)row * p.Co + q * 4);
    }
    for (int i = tid; i < PH * PW * cq; i += 128) {
      const int q = i % cq, pix = i / cq, pw_ = pix % PW, ph_ = pix / PW;
      const int hy = hy_min + ph_, wx = wx_min + pw_;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hy >= 0 && hy < p.Ho && wx >= 0 && wx < p.Wo) v = *reinterpret_cast<const float4*>(dyb + ((size_t)hy * p.Wo + wx) * p.Co + q * 4);
      *reinterpret_cast<float4*>(sdy + pix * pitch + q * 4) = v;
    }
  }
  __syncthreads();

  // thread -> output pixel: the s*s stride-parity classes are laid out class-major, so that (s = 2) each warp holds ONE class
  const int ncls = s * s, per_cls = 128 / ncls;
  const int cls = tid / per_cls, idx = tid % per_cls;
  const int ch = cls / s, cw = cls % s;                 // row / column parity of the class inside the tile
  const int cols = kDW / s;                             // class columns per tile row
  const int h = h0 + (idx / cols) * s + ch, wv = w0 + (idx % cols) * s + cw;
  float acc[CI];
#pragma unroll
  for (int c = 0; c < CI; ++c) acc[c] = 0.f;
  const int r0 = ((h + p.pad_t) % s + s) % s, s0 = ((wv + p.pad_l) % s + s) % s;
  for (int r = r0; r < p.k; r += s) {
    const int hy = (h + p.pad_t - r) / s - hy_min;      // exact division (parity matched); inside the zero-filled patch
    for (int ss = s0; ss < p.k; ss += s) {
      const int wx = (wv + p.pad_l - ss) / s - wx_min;
      const float* dp = sdy + (hy * PW + wx) * pitch;
      const float* wp = sw + ((r * p.k + ss) * CI) * pitch;
#pragma unroll 4
      for (int q = 0; q < cq; ++q) {
        const float4 d = *reinterpret_cast<const float4*>(dp + q * 4);
#pragma unroll
        for (int c = 0; c < CI; ++c) {
          const float4 f = *reinterpret_cast<const float4*>(wp + c * pitch + q * 4);
          acc[c] = fmaf(d.x, f.x, acc[c]);
          acc[c] = fmaf(d.y, f.y, acc[c]);
          acc[c] = fmaf(d.z, f.z, acc[c]);
          acc[c] = fmaf(d.w, f.w, acc[c]);
        }
      }
    }
  }
  if (h < p.H && wv < p.W) {
    float* o = dx + (((size_t)b * p.H + h) * p.W + wv) * CI;
#pragma unroll
    for (int c = 0; c < CI; ++c) o[c] = apply_act(acc[c] + (bias ? bias[c] : 0.f), act, alpha);
  }
}

bool small_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GG_CONV_SMALL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

namespace gg {

// forward conv with Ci <= 4: returns GG_OK and sets *handled when the shape is served here
int conv_small_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Ci, int Co, int k,
                   int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, cudaStream_t st, bool* handled) {
  *handled = false;
  if (!small_enabled() || Ci < 1 || Ci > 4 || Co % 4 != 0 || k * k * Ci > 100 || stride < 1 || stride > 2 || k > 7) return GG_OK;
  SmallP p{B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo};
  const int IW = (kTile - 1) * stride + k;
  const size_t smem = ((size_t)k * k * Ci * kChunk + (size_t)IW * IW * Ci) * sizeof(float);
  if (smem > 48 * 1024 || B > 65535) return GG_OK;
  dim3 grid(ceil_div(Ho, kTile) * ceil_div(Wo, kTile), B, ceil_div(Co, kChunk));
  static int v2 = -1;
  if (v2 < 0) { const char* e = getenv("GG_CONV_SMALL_V2"); v2 = (e && e[0] == '1') ? 1 : 0; }
  if (v2) {
#define GG_V2(CI_)                                                                                                        \
    if (k == 5) GG_LAUNCH((conv_small_fwd_v2_kernel<CI_, 5>), grid, 128, smem, st, x, w, bias, y, p, act, alpha);                    \
    else GG_LAUNCH((conv_small_fwd_v2_kernel<CI_, 0>), grid, 128, smem, st, x, w, bias, y, p, act, alpha)
    switch (Ci) {
      case 1: GG_V2(1); break;
      case 2: GG_V2(2); break;
      case 3: GG_V2(3); break;
      default: GG_V2(4); break;
    }
#undef GG_V2
    *handled = true;
    return check_launch("gg_conv2d_fwd(small-channel v2)");
  }
  switch (Ci) {
    case 1: GG_LAUNCH((conv_small_fwd_kernel<1>), grid, 256, smem, st, x, w, bias, y, p, act, alpha); break;
    case 2: GG_LAUNCH((conv_small_fwd_kernel<2>), grid, 256, smem, st, x, w, bias, y, p, act, alpha); break;
    case 3: GG_LAUNCH((conv_small_fwd_kernel<3>), grid, 256, smem, st, x, w, bias, y, p, act, alpha); break;
    default: GG_LAUNCH((conv_small_fwd_kernel<4>), grid, 256, smem, st, x, w, bias, y, p, act, alpha); break;
  }
  *handled = true;
  return check_launch("gg_conv2d_fwd(small-channel)");
}

// dgrad (= Deconv2D forward) towards Ci <= 4 channels
int conv_small_dgrad(const float* dy, const float* w, const float* bias, float* dx, int B, int H, int W, int Ci, int Co, int k,
                     int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, cudaStream_t st, bool* handled) {
  *handled = false;
  if (!small_enabled() || Ci < 1 || Ci > 4 || Co % 4 != 0 || stride < 1 || stride > 2 || k > 7) return GG_OK;
  SmallP p{B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo};
  // dy patch extent for an 8 x 16 output tile (tile origins are multiples of the tile size; worst case over origins)
  const int PH = (kDH - 1 + k - 1) / stride + 2, PW = (kDW - 1 + k - 1) / stride + 2;
  const int pitch = Co + 4;
  const size_t smem = ((size_t)k * k * Ci * pitch + (size_t)PH * PW * pitch) * sizeof(float);
  if (smem > 48 * 1024 || B > 65535) return GG_OK;
  dim3 grid(ceil_div(H, kDH) * ceil_div(W, kDW), B);
  switch (Ci) {
    case 1: GG_LAUNCH((conv_small_dgrad_kernel<1>), grid, 128, smem, st, dy, w, bias, dx, p, PH, PW, act, alpha); break;
    case 2: GG_LAUNCH((conv_small_dgrad_kernel<2>), grid, 128, smem, st, dy, w, bias, dx, p, PH, PW, act, alpha); break;
    case 3: GG_LAUNCH((conv_small_dgrad_kernel<3>), grid, 128, smem, st, dy, w, bias, dx, p, PH, PW, act, alpha); break;
    default: GG_LAUNCH((conv_small_dgrad_kernel<4>), grid, 128, smem, st, dy, w, bias, dx, p, PH, PW, act, alpha); break;
  }
  *handled = true;
  return check_launch("gg_conv2d_dgrad(small-channel)");
}

}  // namespace gg
