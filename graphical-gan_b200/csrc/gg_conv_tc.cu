// gg_conv_tc.cu — im2col-free implicit-GEMM convolution on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// One warp-specialised kernel template serves the three GEMMs of a strided conv (tf.nn.conv2d NHWC-internal,
// tflib/ops/conv2d.py:106; its autodiff duals; tf.nn.conv2d_transpose, tflib/ops/deconv2d.py:101):
//
//   MODE 0  fwd    y [B*Ho*Wo, Co]  = sum_{tap, ci-block} X_tap[pixels, 32 ci] * W[tap][32 ci, Co]
//                  A: TMA box over x (C,W,H,B) with elementStrides (1,s,s,1) starting at (ci0, wo0*s+s'-pad_l,
//                     ho0*s+r-pad_t, b0): the stride-s gather of one filter tap, padding by TMA zero fill; K-major.
//                  B: W[tap] rows ci, Co contiguous -> MN-major (128B swizzle with 32-byte atoms, the only MN-major
//                     layout kind::tf32 accepts).
//   MODE 1  dgrad  dx[class pixels, Ci] = sum_{taps of the class, co-block} DY_shift[pixels, 32 co] * W[tap][Ci, 32 co]^T
//                  output pixels are split into stride^2 parity classes so that every class is a dense stride-1
//                  gather of dy (plain TMA box, zero fill at the borders); A K-major, B K-major.  This is also the
//                  Deconv2D forward.
//   MODE 2  wgrad  dw[(tap,ci), Co] = sum_{pixel-block} X_tap[32 px, ci]^T * DY[32 px, Co]
//                  both operands MN-major; M rows enumerate (tap, ci) so the accumulator IS the filter matrix.
//
// Operands are fp32 in HBM; the tensor maps use CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 so the TMA unit rounds to tf32
// (round-to-nearest, measured unbiased) on the way into shared memory; accumulation is fp32 in TMEM.
// Pipeline: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2-5 = epilogue (TMEM -> registers ->
// bias + activation -> global).  4-stage full/empty mbarrier ring.  The GEMMs of this workload are small
// (E.2: 4096x128x1600), so K is split across CTAs to fill the 148 SMs; partial tiles go to an L2-resident workspace
// and the last CTA to arrive on a tile (atomic ticket) sums them in split order (deterministic) and runs the epilogue.
#include "gg_tc_common.cuh"

namespace gg {

namespace {

constexpr int kStages = 4;
constexpr int kABytes = 128 * 32 * 4;          // 16 KB: 128 rows (or 4 x 32 MN-blocks) of 32 fp32
constexpr int kMaxNTile = 128;
constexpr int kThreads = 192;

struct TcParams {
  int B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo;
  int wt, ht, bt;             // pixel box of the A tile (fwd/dgrad: wt*ht*bt = 128; wgrad: = 32)
  int tw, th, tb;             // tiles along each pixel dimension
  int PH, PW;                 // pixel grid the tiles walk (fwd: Ho,Wo; dgrad: class grid; wgrad: Ho,Wo)
  int n_tile, n_tiles;        // GEMM N tiling
  int m_tiles;                // fwd/dgrad: tw*th*tb (per class); wgrad: ceil(taps*Ci/128)
  int splits;
  int act;
  float alpha;
  float* out;
  const float* bias;
  float* partial;             // [splits][tiles][128][n_tile]
  unsigned* counters;         // [tiles], zero on entry, left zero
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], accum_bar;
  __shared__ uint32_t tmem_base_sh;
  __shared__ int last_flag;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = kABytes + p.n_tile * 128;
  const int cblocks = (MODE == 1 ? p.Co : p.Ci) / 32;       // 32-channel K blocks (fwd: ci, dgrad: co)
  const int taps = p.k * p.k;

  // ---- tile decode -------------------------------------------------------------------------------------
  int tile = blockIdx.x;                 // over (class,) m-tile, n-tile
  const int split = blockIdx.y;
  const int nt = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  const int n0 = nt * p.n_tile;
  int cls = 0;
  if (MODE == 1) { cls = mt / p.m_tiles; mt = mt % p.m_tiles; }
  int w0 = 0, h0 = 0, b0 = 0;
  if (MODE != 2) {
    w0 = (mt % p.tw) * p.wt;
    h0 = ((mt / p.tw) % p.th) * p.ht;
    b0 = (mt / (p.tw * p.th)) * p.bt;
  }
  // dgrad parity class
  int a_h = 0, a_w = 0, e_h = 0, e_w = 0, d_h = 0, d_w = 0, nr = p.k, ns = p.k;
  if (MODE == 1) {
    a_h = cls / p.stride; a_w = cls % p.stride;
    e_h = ((a_h - p.pad_t) % p.stride + p.stride) % p.stride;
    e_w = ((a_w - p.pad_l) % p.stride + p.stride) % p.stride;
    d_h = (e_h + p.pad_t - a_h) / p.stride;
    d_w = (e_w + p.pad_l - a_w) / p.stride;
    nr = (p.k - a_h + p.stride - 1) / p.stride;
    ns = (p.k - a_w + p.stride - 1) / p.stride;
  }
  int kb_total;
  if (MODE == 0) kb_total = taps * cblocks;
  else if (MODE == 1) kb_total = nr * ns * cblocks;
  else kb_total = (p.B * p.Ho * p.Wo) / 32;
  const int kb_per = (kb_total + p.splits - 1) / p.splits;
  const int kb0 = split * kb_per;
  const int kb1 = min(kb_total, kb0 + kb_per);
  const int nkb = max(kb1 - kb0, 0);

  // ---- one-time setup ------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  const uint32_t tmem_cols = p.n_tile <= 32 ? 32 : (p.n_tile <= 64 ? 64 : 128);
  if (warp == 2) tmem_alloc(&tmem_base_sh, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_sh;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb0 + i;
        const int s = i % kStages;
        if (i >= kStages) mbar_wait(&empty_bar[s], ((i / kStages) - 1) & 1);
        uint8_t* sA = smem + s * stage_bytes;
        uint8_t* sB = sA + kABytes;
        mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
        if (MODE == 0) {
          const int tap = kb / cblocks, cb = kb % cblocks;
          const int r = tap / p.k, sx = tap % p.k;
          tma_load_4d(sA, &tmA, &full_bar[s], cb * 32, w0 * p.stride + sx - p.pad_l, h0 * p.stride + r - p.pad_t, b0);
          for (int nb = 0; nb < p.n_tile / 32; ++nb)
            tma_load_3d(sB + nb * 4096, &tmB, &full_bar[s], n0 + nb * 32, cb * 32, tap);
        } else if (MODE == 1) {
          const int t = kb / cblocks, cb = kb % cblocks;
          const int rq = t / ns, sq = t % ns;
          const int r = a_h + p.stride * rq, sx = a_w + p.stride * sq;
          tma_load_4d(sA, &tmA, &full_bar[s], cb * 32, w0 + d_w - sq, h0 + d_h - rq, b0);
          tma_load_3d(sB, &tmB, &full_bar[s], cb * 32, n0, r * p.k + sx);
        } else {
          // pixel block kb -> (b, ho, wo) box of 32 pixels
          const int pw0 = (kb % p.tw) * p.wt;
          const int ph0 = ((kb / p.tw) % p.th) * p.ht;
          const int pb0 = (kb / (p.tw * p.th)) * p.bt;
          const int qblocks = p.Ci / 32;
          for (int j = 0; j < 4; ++j) {
            int q = mt * 4 + j;
            if (q >= taps * qblocks) q = taps * qblocks - 1;      // padded rows: valid data, masked at the store
            const int tap = q / qblocks, cb = q % qblocks;
            const int r = tap / p.k, sx = tap % p.k;
            tma_load_4d(sA + j * 4096, &tmA, &full_bar[s], cb * 32, pw0 * p.stride + sx - p.pad_l,
                        ph0 * p.stride + r - p.pad_t, pb0);
          }
          for (int nb = 0; nb < p.n_tile / 32; ++nb)
            tma_load_2d(sB + nb * 4096, &tmB, &full_bar[s], n0 + nb * 32, kb * 32);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr int a_mn = (MODE == 2) ? 1 : 0;
      constexpr int b_mn = (MODE == 1) ? 0 : 1;
      const uint32_t idesc = make_idesc_tf32(128, p.n_tile, a_mn, b_mn);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % kStages;
        mbar_wait(&full_bar[s], (i / kStages) & 1);
        tc_fence_after();
        const uint32_t aBase = smem_u32(smem + s * stage_bytes);
        const uint32_t bBase = aBase + kABytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t ad = a_mn ? make_smem_desc(aBase + j * 1024, 4096, 512, kLayoutSW128_32B)
                                   : make_smem_desc(aBase + j * 32, 16, 1024, kLayoutSW128);
          const uint64_t bd = b_mn ? make_smem_desc(bBase + j * 1024, 4096, 512, kLayoutSW128_32B)
                                   : make_smem_desc(bBase + j * 32, 16, 1024, kLayoutSW128);
          umma_tf32(tmem_d, ad, bd, idesc, (i | j) != 0);
        }
        umma_commit(&empty_bar[s]);          // frees the smem stage once these MMAs have read it
      }
      umma_commit(&accum_bar);               // accumulator complete
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const int m = quad * 32 + lane;               // accumulator row
    const int et = threadIdx.x - 64;              // 0..127
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    bool valid;
    float* orow;
    if (MODE == 0) {
      const int wi = m % p.wt, hi = (m / p.wt) % p.ht, bi = m / (p.wt * p.ht);
      valid = (b0 + bi) < p.B;
      orow = p.out + ((size_t)(((size_t)(b0 + bi) * p.Ho + h0 + hi) * p.Wo + w0 + wi)) * p.Co + n0;
    } else if (MODE == 1) {
      const int wi = m % p.wt, hi = (m / p.wt) % p.ht, bi = m / (p.wt * p.ht);
      const int hh = (h0 + hi) * p.stride + e_h, ww = (w0 + wi) * p.stride + e_w;
      valid = (b0 + bi) < p.B && hh < p.H && ww < p.W;
      orow = p.out + ((size_t)(((size_t)(b0 + bi) * p.H + hh) * p.W + ww)) * p.Ci + n0;
    } else {
      const int row = mt * 128 + m;
      valid = row < taps * p.Ci;
      orow = p.out + (size_t)row * p.Co + n0;
    }
    const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16);
    if (p.splits == 1) {
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        float v[32];
        if (nkb > 0) tmem_ld_32x32(taddr + (uint32_t)c0, v);
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o;
            o.x = v[i], o.y = v[i + 1], o.z = v[i + 2], o.w = v[i + 3];
            if (MODE != 2) {
              if (p.bias) {
                const float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + c0 + i);
                o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
              }
              o.x = apply_act(o.x, p.act, p.alpha); o.y = apply_act(o.y, p.act, p.alpha);
              o.z = apply_act(o.z, p.act, p.alpha); o.w = apply_act(o.w, p.act, p.alpha);
            }
            *reinterpret_cast<float4*>(orow + c0 + i) = o;
          }
        }
      }
    } else {
      // split-K: park the partial tile in the (L2-resident) workspace, last CTA on this tile reduces in split order
      const size_t tile_elems = (size_t)128 * p.n_tile;
      const size_t ntiles_all = gridDim.x;
      float* prow = p.partial + ((size_t)split * ntiles_all + blockIdx.x) * tile_elems + (size_t)m * p.n_tile;
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        float v[32];
        if (nkb > 0) tmem_ld_32x32(taddr + (uint32_t)c0, v);
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(prow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        const unsigned old = atomicAdd(&p.counters[blockIdx.x], 1u);
        const int last = (old == (unsigned)(p.splits - 1));
        if (last) p.counters[blockIdx.x] = 0u;      // self-cleaning ticket
        last_flag = last;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (last_flag) {
        __threadfence();
        if (valid) {
          for (int c0 = 0; c0 < p.n_tile; c0 += 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sp = 0; sp < p.splits; ++sp) {
              const float4 t = __ldcg(reinterpret_cast<const float4*>(
                  p.partial + ((size_t)sp * ntiles_all + blockIdx.x) * tile_elems + (size_t)m * p.n_tile + c0));
              acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
            }
            if (MODE != 2) {
              if (p.bias) {
                const float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + c0);
                acc.x += bb.x; acc.y += bb.y; acc.z += bb.z; acc.w += bb.w;
              }
              acc.x = apply_act(acc.x, p.act, p.alpha); acc.y = apply_act(acc.y, p.act, p.alpha);
              acc.z = apply_act(acc.z, p.act, p.alpha); acc.w = apply_act(acc.w, p.act, p.alpha);
            }
            *reinterpret_cast<float4*>(orow + c0) = acc;
          }
        }
      }
    }
  }
  // ---- teardown -----------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_d, tmem_cols);
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
bool pixel_box(int PW, int PH, int PB, int target, int* wt, int* ht, int* bt) {
  // wt*ht*bt == target with wt | PW, ht | PH (powers of two by construction), bt | PB
  int w = PW < target ? PW : target;
  if (w <= 0 || target % w != 0 || PW % w != 0) return false;
  int h = target / w;
  if (h > PH) h = PH;
  if (h <= 0 || (target / w) % h != 0 || PH % h != 0) return false;
  int b = target / (w * h);
  if (b <= 0 || w * h * b != target || PB % b != 0) return false;
  if (w > 256 || h > 256 || b > 256) return false;
  *wt = w; *ht = h; *bt = b;
  return true;
}

int pick_n_tile(int n) {
  if (n % 128 == 0) return 128;
  if (n % 64 == 0) return 64;
  if (n % 32 == 0) return 32;
  return 0;
}

int pick_splits(int tiles, int kb_min) {
  int s = (kNumSMs + tiles - 1) / tiles;     // aim at >= one full wave of CTAs
  int cap = kb_min / 3;                      // keep >= 3 k-blocks per CTA so the pipeline fills
  if (cap < 1) cap = 1;
  if (s > cap) s = cap;
  if (s > 16) s = 16;
  if (s < 1) s = 1;
  return s;
}

struct TcPlan {
  bool ok;
  TcParams p;
  int grid_x;
  size_t partial_bytes, counter_bytes;
};

bool env_disable_tc() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GG_DISABLE_TC");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

TcPlan make_plan(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int pad_t, int pad_l, int Ho, int Wo) {
  TcPlan pl;
  pl.ok = false;
  TcParams& p = pl.p;
  p = TcParams{};
  p.B = B; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.k = k; p.stride = stride; p.pad_t = pad_t; p.pad_l = pad_l; p.Ho = Ho; p.Wo = Wo;
  if (env_disable_tc()) return pl;
  if (Ci % 32 != 0 || Co % 32 != 0 || stride < 1 || stride > 2 || k > 16) return pl;
  int kb_min;
  if (mode == 0) {
    p.PH = Ho; p.PW = Wo;
    if (!pixel_box(Wo, Ho, B, 128, &p.wt, &p.ht, &p.bt)) return pl;
    p.n_tile = pick_n_tile(Co);
    if (!p.n_tile) return pl;
    p.n_tiles = Co / p.n_tile;
    p.tw = Wo / p.wt; p.th = Ho / p.ht; p.tb = B / p.bt;
    p.m_tiles = p.tw * p.th * p.tb;
    pl.grid_x = p.m_tiles * p.n_tiles;
    kb_min = k * k * (Ci / 32);
  } else if (mode == 1) {
    if (H % stride != 0 || W % stride != 0) return pl;
    p.PH = H / stride; p.PW = W / stride;
    if (!pixel_box(p.PW, p.PH, B, 128, &p.wt, &p.ht, &p.bt)) return pl;
    p.n_tile = pick_n_tile(Ci);
    if (!p.n_tile) return pl;
    p.n_tiles = Ci / p.n_tile;
    p.tw = p.PW / p.wt; p.th = p.PH / p.ht; p.tb = B / p.bt;
    p.m_tiles = p.tw * p.th * p.tb;
    pl.grid_x = stride * stride * p.m_tiles * p.n_tiles;
    int nmin = k / stride;                    // fewest taps a class sees along one axis
    if (nmin < 1) return pl;
    kb_min = nmin * nmin * (Co / 32);
  } else {
    p.PH = Ho; p.PW = Wo;
    if (!pixel_box(Wo, Ho, B, 32, &p.wt, &p.ht, &p.bt)) return pl;
    p.n_tile = pick_n_tile(Co);
    if (!p.n_tile) return pl;
    p.n_tiles = Co / p.n_tile;
    p.tw = Wo / p.wt; p.th = Ho / p.ht; p.tb = B / p.bt;
    p.m_tiles = (k * k * Ci + 127) / 128;
    pl.grid_x = p.m_tiles * p.n_tiles;
    kb_min = (B * Ho * Wo) / 32;
  }
  if (p.wt * stride > 256 || p.ht * stride > 256) return pl;
  p.splits = pick_splits(pl.grid_x, kb_min);
  pl.partial_bytes = p.splits > 1 ? (size_t)p.splits * pl.grid_x * 128 * p.n_tile * sizeof(float) : 0;
  pl.counter_bytes = ((size_t)pl.grid_x * sizeof(unsigned) + 255) & ~size_t(255);
  pl.ok = true;
  return pl;
}

size_t plan_workspace(const TcPlan& pl) { return pl.ok ? pl.counter_bytes + pl.partial_bytes : 0; }

template <int MODE>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, TcPlan& pl, void* ws, size_t ws_bytes, cudaStream_t st) {
  TcParams& p = pl.p;
  if (p.splits > 1 && (ws == nullptr || ws_bytes < plan_workspace(pl))) {
    // not enough workspace for split-K: run unsplit (slower, still correct)
    p.splits = 1;
  }
  p.counters = reinterpret_cast<unsigned*>(ws);
  p.partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + pl.counter_bytes);
  const size_t smem = (size_t)kStages * (kABytes + p.n_tile * 128) + 1024;
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[MODE]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kStages * (kABytes + kMaxNTile * 128) + 1024);
    if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "conv_tc: cudaFuncSetAttribute failed%s");
    attr_set[MODE] = true;
  }
  dim3 grid(pl.grid_x, p.splits);
  conv_tc_kernel<MODE><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  return check_launch(MODE == 0 ? "gg_conv2d_fwd(tcgen05)" : (MODE == 1 ? "gg_conv2d_dgrad(tcgen05)" : "gg_conv2d_wgrad(tcgen05)"));
}

// filter tensor map: w [taps][Ci][Co] viewed as dims (Co, Ci, taps)
int filter_map_mn(CUtensorMap* tm, const float* w, int Ci, int Co, int taps) {   // boxes of 32 co x 32 ci, MN-major operand
  uint64_t dims[3] = {(uint64_t)Co, (uint64_t)Ci, (uint64_t)taps};
  uint64_t str[3] = {1, (uint64_t)Co, (uint64_t)Ci * Co};
  uint32_t box[3] = {32, 32, 1};
  return encode_tmap(tm, w, 3, dims, str, box, nullptr, 2, true);
}
int filter_map_k(CUtensorMap* tm, const float* w, int Ci, int Co, int taps, int n_tile) {  // 32 co x n_tile ci rows, K-major
  uint64_t dims[3] = {(uint64_t)Co, (uint64_t)Ci, (uint64_t)taps};
  uint64_t str[3] = {1, (uint64_t)Co, (uint64_t)Ci * Co};
  uint32_t box[3] = {32, (uint32_t)n_tile, 1};
  return encode_tmap(tm, w, 3, dims, str, box, nullptr, 1, true);
}
// activation tensor map over an NHWC tensor (C,W,H,B); box = 32 channels x (wt,ht,bt) pixels gathered with stride es
int act_map(CUtensorMap* tm, const float* x, int B, int H, int W, int C, int wt, int ht, int bt, int es, int swizzle) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t str[4] = {1, (uint64_t)C, (uint64_t)W * C, (uint64_t)H * W * C};
  uint32_t box[4] = {32, (uint32_t)(wt * es), (uint32_t)(ht * es), (uint32_t)bt};
  uint32_t estr[4] = {1, (uint32_t)es, (uint32_t)es, 1};
  return encode_tmap(tm, x, 4, dims, str, box, estr, swizzle, true);
}

}  // namespace

int conv_tc_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Ci, int Co, int k,
                int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, void* ws, size_t ws_bytes,
                cudaStream_t st, bool* handled) {
  *handled = false;
  TcPlan pl = make_plan(0, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
  if (!pl.ok) return GG_OK;
  if (ws == nullptr || ws_bytes < pl.counter_bytes) return GG_OK;     // needs at least the ticket area
  CUtensorMap tmA, tmB;
  int rc = act_map(&tmA, x, B, H, W, Ci, pl.p.wt, pl.p.ht, pl.p.bt, stride, 1);
  if (rc) return rc;
  rc = filter_map_mn(&tmB, w, Ci, Co, k * k);
  if (rc) return rc;
  pl.p.out = y; pl.p.bias = bias; pl.p.act = act; pl.p.alpha = alpha;
  rc = launch<0>(tmA, tmB, pl, ws, ws_bytes, st);
  if (rc) return rc;
  *handled = true;
  return GG_OK;
}

int conv_tc_dgrad(const float* dy, const float* w, const float* bias, float* dx, int B, int H, int W, int Ci, int Co, int k,
                  int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, void* ws, size_t ws_bytes,
                  cudaStream_t st, bool* handled) {
  *handled = false;
  TcPlan pl = make_plan(1, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
  if (!pl.ok) return GG_OK;
  if (ws == nullptr || ws_bytes < pl.counter_bytes) return GG_OK;
  CUtensorMap tmA, tmB;
  int rc = act_map(&tmA, dy, B, Ho, Wo, Co, pl.p.wt, pl.p.ht, pl.p.bt, 1, 1);
  if (rc) return rc;
  rc = filter_map_k(&tmB, w, Ci, Co, k * k, pl.p.n_tile);
  if (rc) return rc;
  pl.p.out = dx; pl.p.bias = bias; pl.p.act = act; pl.p.alpha = alpha;
  rc = launch<1>(tmA, tmB, pl, ws, ws_bytes, st);
  if (rc) return rc;
  *handled = true;
  return GG_OK;
}

int conv_tc_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Ci, int Co, int k, int stride,
                  int pad_t, int pad_l, int Ho, int Wo, void* ws, size_t ws_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  TcPlan pl = make_plan(2, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
  if (!pl.ok) return GG_OK;
  if (ws == nullptr || ws_bytes < pl.counter_bytes) return GG_OK;
  CUtensorMap tmA, tmB;
  int rc = act_map(&tmA, x, B, H, W, Ci, pl.p.wt, pl.p.ht, pl.p.bt, stride, 2);
  if (rc) return rc;
  {
    uint64_t dims[2] = {(uint64_t)Co, (uint64_t)B * Ho * Wo};
    uint64_t str[2] = {1, (uint64_t)Co};
    uint32_t box[2] = {32, 32};
    rc = encode_tmap(&tmB, dy, 2, dims, str, box, nullptr, 2, true);
    if (rc) return rc;
  }
  pl.p.out = dw; pl.p.bias = nullptr; pl.p.act = 0; pl.p.alpha = 0.f;
  rc = launch<2>(tmA, tmB, pl, ws, ws_bytes, st);
  if (rc) return rc;
  *handled = true;
  return GG_OK;
}

size_t conv_tc_workspace(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo) {
  TcPlan pl = make_plan(mode, B, H, W, Ci, Co, k, stride, 0, 0, Ho, Wo);
  return plan_workspace(pl);
}

size_t conv_tc_wgrad_workspace(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo) {
  return conv_tc_workspace(2, B, H, W, Ci, Co, k, stride, Ho, Wo);
}

}  // namespace gg
