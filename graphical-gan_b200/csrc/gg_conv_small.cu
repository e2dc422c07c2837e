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

#include <map>
#include <tuple>

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

// Two variants were measured on the B200 and removed (B=128 3->64, 16.2 us for the kernel above; profiles/time_conv_r2.txt):
// a 4 pixel x 8 channel register tile with the 5x5 tap loop unrolled (18.8 us), and a persistent form that keeps the filter slab
// resident and double-buffers the input patch with cp.async (16.0 us: the per-tile fills were not the limit).  The inner loop
// issues 1 LDS.128 + 4 LDS.32 per 16 FMAs — 8 shared-memory wavefronts per 16 FMA instructions: the kernel runs at the
// shared-memory pipe's rate (2.45 M wavefronts / 148 SMs = 8.4 us) plus launch, fill and store tails.

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

// ---------------------------------------------------------------------------------------------------------------------
// wgrad (filter gradient) for CI <= 4 input channels: ONE launch, no patch matrix
// ---------------------------------------------------------------------------------------------------------------------
// dw[(r*k+s)*CI+c][co] = sum over (b, ho, wo) of x[b, ho*st+r-pad_t, wo*st+s-pad_l, c] * dy[b, ho, wo, co]
// (gradient of tf.nn.conv2d w.r.t. its HWIO filter, tflib/ops/conv2d.py:106).  K = k*k*CI <= 100 rows x Co columns is a tiny
// output reduced over B*Ho*Wo = 10^4..10^5 pixels: 0.3 GFLOP against 10 MB of activations read ONCE.  It used to run as patch
// matrix + tcgen05 GEMM with M = 75 of 128 rows used, a single output tile and hence at most 8 split-K CTAs: 14 + 26 us at the
// tail of both steps (profiles/time_conv_r2.txt).  Here the PIXELS are split over the whole machine:
//   unit    = RH output rows of one image; its x patch and dy rows are staged in shared memory once;
//   thread  = 5 filter rows (tap, c) x 8 output channels = 40 fp32 accumulators held over every unit the CTA walks;
//             per pixel 5 broadcast LDS.32 (x) + 2 LDS.128 (dy) feed 40 FMAs; the pixels of a unit are split over PS thread
//             groups that are folded in a fixed order through shared memory;
//   cluster = 8 CTAs fold their [K][Co] tiles through distributed shared memory (rank r owns an eighth of the floats and
//             adds the eight copies in rank order), one partial per cluster goes to the L2 workspace, and the LAST cluster to
//             arrive (self-resetting ticket at the start of the workspace) folds the partials in cluster order into dw.
// Every sum has a fixed order: the result is bit-reproducible and independent of scheduling.
constexpr int kWgTK = 5;            // filter rows per thread
constexpr int kWgCL = 8;            // CTAs per cluster
constexpr int kWgMaxClusters = 18;  // 18 x 8 = 144 of 148 SMs, one CTA per SM

__device__ __forceinline__ long long wg_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t wg_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t wg_cluster_id() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void wg_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 wg_ld_dsmem_f4(const float* local_ptr, uint32_t cta_rank) {
  uint32_t local = (uint32_t)__cvta_generic_to_shared(local_ptr), remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta_rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t wg_ld_dsmem_u32(const uint32_t* local_ptr, uint32_t cta_rank) {
  uint32_t local = (uint32_t)__cvta_generic_to_shared(local_ptr), remote, v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta_rank));
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
  return v;
}

struct WgPlan {
  bool ok;
  int RH, PW, NT, PS, threads, clusters;
  size_t sx_floats, tile_floats, smem, ws_bytes;
};

template <int CI>
__global__ void __launch_bounds__(512) conv_small_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dw, float* __restrict__ part,
                                                               unsigned* __restrict__ ticket, SmallP p, int RH, int PW, int NT,
                                                               int PS, int sx_floats, long long* __restrict__ dbg) {
#define GG_WG_STAMP(slot) do { if (dbg && threadIdx.x == 0) dbg[blockIdx.x * 8 + (slot)] = wg_gtime(); } while (0)
  GG_WG_STAMP(0);
  GG_PDL_ENTRY();
  extern __shared__ __align__(16) float smem_f[];
  __shared__ uint32_t s_last;
  float* sx = smem_f;                                   // [(rows-1)*stride+k][PW][CI] input patch, TF SAME padding as zeros
  float* sdy = smem_f + sx_floats;                      // [rows*Wo][Co] dy rows of the unit; later the CTA's [K][Co] tile
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int K = p.k * p.k * CI, Co = p.Co, N = K * Co;
  const int nq8 = Co >> 3;                              // channel octets: quads q and q + nq8
  const bool worker = tid < NT * PS;
  const int ps = worker ? tid / NT : 0, tt = worker ? tid % NT : 0;
  const int g = tt / nq8, q = tt % nq8;
  int off[kWgTK];
#pragma unroll
  for (int j = 0; j < kWgTK; ++j) {
    const int kk = g * kWgTK + j;
    const int tap = kk / CI, c = kk % CI;
    off[j] = (kk < K) ? ((tap / p.k) * PW + (tap % p.k)) * CI + c : 0;
  }
  float acc[kWgTK][8];
#pragma unroll
  for (int j = 0; j < kWgTK; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  // the B*Ho output rows are dealt to the CTAs as contiguous, equal ranges (a 8-CTA cluster needs its SMs inside one GPC: the
  // device keeps only 15 of them resident, so a grid of "one image per CTA" ran its 16th cluster as a second wave, profiles/
  // timeline_small_wgrad_r2.txt); a range is walked in units of at most RH rows of ONE image
  const long long total_rows = (long long)p.B * p.Ho;
  const long long g_lo = total_rows * blockIdx.x / gridDim.x, g_hi = total_rows * (blockIdx.x + 1) / gridDim.x;
  for (long long g0 = g_lo; g0 < g_hi;) {
    const int b = (int)(g0 / p.Ho), ho0 = (int)(g0 % p.Ho);
    int rows = p.Ho - ho0;
    if (rows > RH) rows = RH;
    if ((long long)rows > g_hi - g0) rows = (int)(g_hi - g0);
    if (g0 != g_lo) __syncthreads();                    // previous unit's readers are done with the tiles
    g0 += rows;
    const int npix = rows * p.Wo;
    const int PP = (npix + PS - 1) / PS;
    const int p_lo = min(npix, ps * PP), p_hi = min(npix, p_lo + PP);
    // dy rows ho0 .. ho0+rows-1 of image b are one contiguous run: asynchronous 16-byte copies, all in flight at once,
    // overlapped with the index arithmetic of the x patch fill below
    const float4* dsrc = reinterpret_cast<const float4*>(dy + ((size_t)b * p.Ho + ho0) * p.Wo * Co);
    const int nvalid = npix * (Co >> 2);
    float4* sdy4 = reinterpret_cast<float4*>(sdy);
    for (int i = tid; i < nvalid; i += nthr) {
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sdy4 + i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(dsrc + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // x patch
    const int hi0 = ho0 * p.stride - p.pad_t, wi0 = -p.pad_l;
    const float* xb = x + (size_t)b * p.H * p.W * CI;
    const int nx = ((rows - 1) * p.stride + p.k) * PW * CI;
    for (int base = tid; base < nx; base += 8 * nthr) {
      float v[8];
#pragma unroll
      for (int uu = 0; uu < 8; ++uu) {
        const int i = base + uu * nthr;
        const int c = i % CI, iw = (i / CI) % PW, ih = i / (CI * PW);
        const int hi = hi0 + ih, wi = wi0 + iw;
        v[uu] = 0.f;
        if (i < nx && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) v[uu] = xb[((size_t)hi * p.W + wi) * CI + c];
      }
#pragma unroll
      for (int uu = 0; uu < 8; ++uu) {
        const int i = base + uu * nthr;
        if (i < nx) sx[i] = v[uu];
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    GG_WG_STAMP(1);                                     // tiles of the (last) unit staged
    if (worker) {
      int pr = p_lo / p.Wo, wo = p_lo % p.Wo;
      const float4* dq = reinterpret_cast<const float4*>(sdy) + q;
      const int cq4 = Co >> 2;
#pragma unroll 2
      for (int pix = p_lo; pix < p_hi; ++pix) {
        const float* xp = sx + ((pr * p.stride) * PW + wo * p.stride) * CI;
        const float4 d0 = dq[pix * cq4], d1 = dq[pix * cq4 + nq8];
        float xv[kWgTK];
#pragma unroll
        for (int j = 0; j < kWgTK; ++j) xv[j] = xp[off[j]];
#pragma unroll
        for (int j = 0; j < kWgTK; ++j) {
          acc[j][0] = fmaf(xv[j], d0.x, acc[j][0]); acc[j][1] = fmaf(xv[j], d0.y, acc[j][1]);
          acc[j][2] = fmaf(xv[j], d0.z, acc[j][2]); acc[j][3] = fmaf(xv[j], d0.w, acc[j][3]);
          acc[j][4] = fmaf(xv[j], d1.x, acc[j][4]); acc[j][5] = fmaf(xv[j], d1.y, acc[j][5]);
          acc[j][6] = fmaf(xv[j], d1.z, acc[j][6]); acc[j][7] = fmaf(xv[j], d1.w, acc[j][7]);
        }
        if (++wo == p.Wo) { wo = 0; ++pr; }
      }
    }
  }
  __syncthreads();
  GG_WG_STAMP(2);                                       // FMA loops done
  // fold the PS pixel groups into the CTA's [K][Co] tile (group order), aliased over the dy rows
  float* sacc = sdy;
  for (int i = tid; i < N; i += nthr) sacc[i] = 0.f;
  for (int s = 0; s < PS; ++s) {
    __syncthreads();
    if (worker && ps == s) {
#pragma unroll
      for (int j = 0; j < kWgTK; ++j) {
        const int kk = g * kWgTK + j;
        if (kk < K) {
          float4* r0 = reinterpret_cast<float4*>(sacc + kk * Co + q * 4);
          float4* r1 = reinterpret_cast<float4*>(sacc + kk * Co + (q + nq8) * 4);
          float4 a = *r0, c4 = *r1;
          a.x += acc[j][0]; a.y += acc[j][1]; a.z += acc[j][2]; a.w += acc[j][3];
          c4.x += acc[j][4]; c4.y += acc[j][5]; c4.z += acc[j][6]; c4.w += acc[j][7];
          *r0 = a; *r1 = c4;
        }
      }
    }
  }
  // cluster fold through distributed shared memory: rank r owns float4s [r*per, (r+1)*per)
  const uint32_t rank = wg_cluster_ctarank(), cid = wg_cluster_id();
  const int nclusters = gridDim.x / kWgCL;
  const int n4 = N >> 2, per = (n4 + kWgCL - 1) / kWgCL;
  const int lo4 = rank * per, hi4 = min(n4, lo4 + per);
  wg_cluster_sync();
  GG_WG_STAMP(3);                                       // CTA tiles folded, cluster rendezvous passed
  float4* out4 = reinterpret_cast<float4*>(nclusters == 1 ? dw : part + (size_t)cid * N);
  for (int i = lo4 + tid; i < hi4; i += nthr) {
    float4 t = wg_ld_dsmem_f4(sacc + i * 4, 0);
#pragma unroll
    for (uint32_t r = 1; r < (uint32_t)kWgCL; ++r) {
      const float4 v = wg_ld_dsmem_f4(sacc + i * 4, r);
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    out4[i] = t;
  }
  if (nclusters == 1) {
    wg_cluster_sync();                                  // peers may still be reading this CTA's tile
    return;
  }
  __threadfence();
  wg_cluster_sync();                                    // every rank's slice of the partial is written and fenced
  GG_WG_STAMP(4);                                       // cluster partial in L2
  if (rank == 0 && tid == 0) {
    __threadfence();                                    // the ranks' fenced writes are ordered before the ticket (cumulative)
    const unsigned t = atomicAdd(ticket, 1u);
    const bool last = (t == (unsigned)nclusters - 1u);
    if (last) *ticket = 0u;                             // self-resetting: the next launch starts from zero again
    __threadfence();
    s_last = last ? 1u : 0u;
  }
  wg_cluster_sync();
  const uint32_t last = wg_ld_dsmem_u32(&s_last, 0);
  wg_cluster_sync();                                    // rank 0 keeps its shared memory alive until every rank has read
  GG_WG_STAMP(5);                                       // ticket taken, flag read
  if (!last) return;
  __threadfence();
  float4* dw4 = reinterpret_cast<float4*>(dw);
  for (int i = lo4 + tid; i < hi4; i += nthr) {
    float4 v[kWgMaxClusters];
#pragma unroll
    for (int c = 0; c < kWgMaxClusters; ++c) {              // every partial of this float4 in flight together
      v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nclusters) v[c] = __ldcg(reinterpret_cast<const float4*>(part + (size_t)c * N) + i);
    }
    float4 t = v[0];
#pragma unroll
    for (int c = 1; c < kWgMaxClusters; ++c) {              // cluster order; absent clusters add +0
      t.x += v[c].x; t.y += v[c].y; t.z += v[c].z; t.w += v[c].w;
    }
    dw4[i] = t;
  }
  GG_WG_STAMP(6);                                       // last cluster: dw written
#undef GG_WG_STAMP
}

WgPlan wgrad_plan(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo) {
  WgPlan pl{};
  pl.ok = false;
  (void)H; (void)W;
  if (Ci < 1 || Ci > 4 || Co % 8 != 0 || Co > 256 || k > 7 || stride < 1 || stride > 2 || B > (1 << 20)) return pl;
  const int K = k * k * Ci;
  if (K > 100) return pl;
  const int rh_max = 16384 / (Wo * Co);                 // dy rows of a unit: at most 64 KB of shared memory
  if (rh_max < 1) return pl;
  pl.RH = Ho < rh_max ? Ho : rh_max;
  pl.PW = (Wo - 1) * stride + k;
  pl.NT = ((K + kWgTK - 1) / kWgTK) * (Co / 8);
  if (pl.NT > 512) return pl;
  pl.PS = 512 / pl.NT;
  if (pl.PS > 4) pl.PS = 4;
  pl.threads = ((pl.NT * pl.PS + 31) / 32) * 32;
  if (pl.threads < 128) pl.threads = 128;
  pl.sx_floats = (((size_t)((pl.RH - 1) * stride + k) * pl.PW * Ci + 3) / 4) * 4;
  const size_t tile = (size_t)pl.RH * Wo * Co, accs = (size_t)K * Co;
  pl.tile_floats = tile > accs ? tile : accs;
  pl.smem = (pl.sx_floats + pl.tile_floats) * sizeof(float);
  if (pl.smem > 200 * 1024) return pl;
  const long long total_rows = (long long)B * Ho;
  long long clusters = (total_rows + 2 * kWgCL - 1) / (2 * kWgCL);     // at least two output rows per CTA
  if (clusters > kWgMaxClusters) clusters = kWgMaxClusters;
  pl.clusters = (int)clusters;                          // upper bound; the launch caps it at what the device keeps resident
  pl.ws_bytes = 256 + (size_t)kWgMaxClusters * accs * sizeof(float);
  pl.ok = true;
  return pl;
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

// 8-CTA clusters of this kernel the device can keep resident at once (cudaOccupancyMaxActiveClusters; cached per configuration).
// Also opts the kernel into > 48 KB of dynamic shared memory.
int wg_resident_clusters(int Ci, int threads, size_t smem) {
  static std::map<std::tuple<int, int, size_t>, int> cache;
  const auto key = std::make_tuple(Ci, threads, smem);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(kNumSMs / kWgCL * kWgCL));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kWgCL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaSuccess;
#define GG_WG_OCC(CI_)                                                                                                      \
  e = cudaFuncSetAttribute(conv_small_wgrad_kernel<CI_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);          \
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, conv_small_wgrad_kernel<CI_>, &cfg)
  switch (Ci) {
    case 1: GG_WG_OCC(1); break;
    case 2: GG_WG_OCC(2); break;
    case 3: GG_WG_OCC(3); break;
    default: GG_WG_OCC(4); break;
  }
#undef GG_WG_OCC
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[key] = n;
  return n;
}

long long* g_small_dbg = nullptr;   // gg_debug_set_small_buffer: per-CTA %globaltimer stamps of the small-channel wgrad kernel
void conv_small_set_debug(void* p) { g_small_dbg = reinterpret_cast<long long*>(p); }

size_t conv_small_wgrad_workspace(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo) {
  if (!small_enabled()) return 0;
  const WgPlan pl = wgrad_plan(B, H, W, Ci, Co, k, stride, Ho, Wo);
  return pl.ok ? pl.ws_bytes : 0;
}

// filter gradient of a conv with Ci <= 4.  The first 256 bytes of the workspace hold the arrival ticket: they must be zero
// before the FIRST launch that uses the workspace (the kernel leaves them zero again).
int conv_small_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Ci, int Co, int k, int stride, int pad_t,
                     int pad_l, int Ho, int Wo, void* ws, size_t ws_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  if (!small_enabled()) return GG_OK;
  const WgPlan pl = wgrad_plan(B, H, W, Ci, Co, k, stride, Ho, Wo);
  if (!pl.ok || ws == nullptr || ws_bytes < pl.ws_bytes) return GG_OK;
  SmallP p{B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo};
  unsigned* ticket = reinterpret_cast<unsigned*>(ws);
  float* part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + 256);
  int clusters = pl.clusters;
  const int resident = wg_resident_clusters(Ci, pl.threads, pl.smem);
  if (resident > 0 && clusters > resident) clusters = resident;      // one wave: a second wave would double the launch
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * kWgCL));
  cfg.blockDim = dim3((unsigned)pl.threads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = kWgCL;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (g_null_launch) { launch_null(st); *handled = true; return check_launch("gg_conv2d_wgrad(small-channel, null launch)"); }
  cudaError_t e = cudaSuccess;
#define GG_WG(CI_)                                                                                                          \
  e = cudaLaunchKernelEx(&cfg, conv_small_wgrad_kernel<CI_>, x, dy, dw, part, ticket, p, pl.RH, pl.PW, pl.NT, pl.PS,        \
                         (int)pl.sx_floats, g_small_dbg)
  switch (Ci) {
    case 1: GG_WG(1); break;
    case 2: GG_WG(2); break;
    case 3: GG_WG(3); break;
    default: GG_WG(4); break;
  }
#undef GG_WG
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(GG_ERR_CUDA_BASE + (int)e, "gg_conv2d_wgrad(small-channel): launch failed%s");
  }
  *handled = true;
  return check_launch("gg_conv2d_wgrad(small-channel)");
}

// launch plan of the small-channel wgrad kernel for a geometry (development aid): out8 = {ok, max rows per unit, B*Ho rows,
// threads, pixel groups, clusters launched, dynamic shared memory bytes, clusters the device keeps resident at once}
int conv_small_wgrad_info(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo, int* out8) {
  const WgPlan pl = wgrad_plan(B, H, W, Ci, Co, k, stride, Ho, Wo);
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  if (!pl.ok) return GG_OK;
  const int resident = wg_resident_clusters(Ci, pl.threads, pl.smem);
  out8[0] = 1; out8[1] = pl.RH; out8[2] = B * Ho; out8[3] = pl.threads; out8[4] = pl.PS;
  out8[5] = (resident > 0 && pl.clusters > resident) ? resident : pl.clusters;
  out8[6] = (int)pl.smem;
  out8[7] = resident;
  return GG_OK;
}

}  // namespace gg
