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

constexpr int kMaxStages = 6;                  // 6 x 32 KB operand stages + 1 KB alignment slack fit the 227 KB of an SM
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
  int stages;                 // depth of the operand ring (as many as fit)
  int cluster;                // 1: the `splits` CTAs of a tile form a thread-block cluster and reduce through DSMEM
  int bulk;                   // 1: split-K slices come back through cp.async.bulk (copy engine) instead of per-thread L2 loads
  int act;
  float alpha;
  float* out;
  const float* bias;
  float* partial;             // [splits][tiles][128][n_tile]
  unsigned* counters;         // [tiles], zero on entry, left zero
  long long* dbg;             // optional timeline of CTA (0,0): see gg_debug_set_buffer
  int out_rows;               // wgrad: rows of the [taps*Ci, Co] result that exist in memory (im2col-padded K)
};

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 24)) __trap();
  } while (!done);
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// 1-D bulk copy global -> shared through the async proxy (copy engine), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

#ifdef GG_TIMELINE
#define GG_DBG(slot) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0) p.dbg[slot] = gtime(); } while (0)
#else
#define GG_DBG(slot) do { } while (0)
#endif

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], accum_bar, cl_bar;
  if (p.dbg && threadIdx.x == 0) {             // timeline (tools/timeline_conv.py): earliest CTA entry of the launch
    const long long t_in = gtime();
    atomicMin(reinterpret_cast<unsigned long long*>(p.dbg + 210), (unsigned long long)t_in);
    if (blockIdx.x == 0 && blockIdx.y == 0) p.dbg[211] = t_in;
  }
  const int kStages = p.stages;
  __shared__ uint32_t tmem_base_sh;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: the next kernel may start its own prologue now
  __shared__ long long s_row_off[128];                  // element offset of each accumulator row's output row (-1: masked)
  __shared__ __align__(16) float s_bias[kMaxNTile];   // bias slice of this n-tile (read from HBM/L2 once, before the accumulator is ready)

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
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    mbar_init(&cl_bar, (uint32_t)p.splits);      // cluster mode: one arrival per CTA of the tile
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  const uint32_t tmem_cols = p.n_tile <= 32 ? 32 : (p.n_tile <= 64 ? 64 : 128);
  if (warp == 2) tmem_alloc(&tmem_base_sh, tmem_cols);
  tc_fence_before();
  const bool cl = p.cluster != 0;
  if (cl) cluster_sync_all();                  // peers must not signal cl_bar before it is initialised
  else __syncthreads();
  tc_fence_after();
  // PDL: everything above (barrier init, tensor-map prefetch, TMEM allocation) is independent of the producer kernel; its
  // activations / gradients are first touched below (TMA loads, bias read), so the dependency is resolved here.  A no-op
  // for launches without a programmatic dependency.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_d = tmem_base_sh;
  if (threadIdx.x == 0) GG_DBG(0);
  if (p.dbg && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long*>(p.dbg + 201), (unsigned long long)gtime());

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // Coordinates advance incrementally (no integer divisions in the issue loop: one elected lane issues every TMA of
    // the CTA, so its instruction stream IS the load-issue rate); each operand stage is one TMA where the layout allows.
    if (lane == 0) {
      int c_cb = 0, c_s = 0, c_r = 0;          // fwd / dgrad: channel block, tap column, tap row (dgrad: class-local)
      int px_w = 0, px_h = 0, px_b = 0;        // wgrad: pixel-block origin
      int qa_c[4], qa_w[4], qa_h[4];           // wgrad: loop-invariant (ci block, tap offsets) of the 4 A boxes
      if (MODE == 2) {
        px_w = (kb0 % p.tw) * p.wt;
        px_h = ((kb0 / p.tw) % p.th) * p.ht;
        px_b = (kb0 / (p.tw * p.th)) * p.bt;
        const int qblocks = p.Ci / 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int q = mt * 4 + j;
          if (q >= taps * qblocks) q = taps * qblocks - 1;        // padded rows: valid data, masked at the store
          const int tap = q / qblocks;
          qa_c[j] = (q % qblocks) * 32;
          qa_w[j] = tap % p.k - p.pad_l;
          qa_h[j] = tap / p.k - p.pad_t;
        }
      } else {
        const int t0 = kb0 / cblocks;
        c_cb = kb0 % cblocks;
        const int row_len = (MODE == 0) ? p.k : ns;
        c_r = t0 / row_len;
        c_s = t0 % row_len;
      }
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&empty_bar[s], ph ^ 1);            // a fresh barrier passes a parity-1 wait: first lap never blocks
        uint8_t* sA = smem + s * stage_bytes;
        uint8_t* sB = sA + kABytes;
        if (i < 60) GG_DBG(1 + i);
        mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
        if (MODE == 0) {
          tma_load_4d(sA, &tmA, &full_bar[s], c_cb * 32, w0 * p.stride + c_s - p.pad_l, h0 * p.stride + c_r - p.pad_t, b0);
          tma_load_4d(sB, &tmB, &full_bar[s], 0, c_cb * 32, n0 / 32, c_r * p.k + c_s);
        } else if (MODE == 1) {
          tma_load_4d(sA, &tmA, &full_bar[s], c_cb * 32, w0 + d_w - c_s, h0 + d_h - c_r, b0);
          tma_load_3d(sB, &tmB, &full_bar[s], c_cb * 32, n0, (a_h + p.stride * c_r) * p.k + a_w + p.stride * c_s);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_4d(sA + j * 4096, &tmA, &full_bar[s], qa_c[j], px_w * p.stride + qa_w[j], px_h * p.stride + qa_h[j], px_b);
          tma_load_3d(sB, &tmB, &full_bar[s], 0, (kb0 + i) * 32, n0 / 32);
        }
        if (MODE == 2) {
          px_w += p.wt;
          if (px_w == p.PW) { px_w = 0; px_h += p.ht; if (px_h == p.PH) { px_h = 0; px_b += p.bt; } }
        } else {
          if (++c_cb == cblocks) {
            c_cb = 0;
            const int row_len = (MODE == 0) ? p.k : ns;
            if (++c_s == row_len) { c_s = 0; ++c_r; }
          }
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr int a_mn = (MODE == 2) ? 1 : 0;
      constexpr int b_mn = (MODE == 1) ? 0 : 1;
      const uint32_t idesc = make_idesc_tf32(128, p.n_tile, a_mn, b_mn);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (i < 60) GG_DBG(64 + i);
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
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
      umma_commit(&accum_bar);               // accumulator complete
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const int m = quad * 32 + lane;               // accumulator row
    const int et = threadIdx.x - 64;              // 0..127
    // While the main loop runs these warps are idle: fetch the bias slice now.  (Loading it from global memory inside the
    // read-out loop cost one exposed L2 round trip per 32-column chunk — the tcgen05.ld asm is a compiler barrier —
    // ~2 us of a ~3.4 us epilogue on the timeline.)
    const bool has_bias = (MODE != 2) && p.bias != nullptr;
    if (has_bias && et < p.n_tile) s_bias[et] = p.bias[nt * p.n_tile + et];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    if (et == 0) GG_DBG(128);
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
      valid = row < p.out_rows;
      orow = p.out + (size_t)row * p.Co + n0;
    }
    const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16);
    // The operand ring is idle once accum_bar has fired: reuse it as a [128][n_tile+4] staging tile so that the final
    // global stores are row-coalesced (a warp writes one whole output row per instruction).
    // (cluster mode keeps this CTA's partial tile in the first 64 KB for its peers to read; staging goes behind it)
    const int stage_off = cl ? 128 * p.n_tile * 4 : 0;
    float* stage_tile = reinterpret_cast<float*>(smem + stage_off);
    long long* row_off = s_row_off;
    int ld = p.n_tile + 4;                        // staging row pitch (floats)
    int col0 = 0;                                 // float4 column of the tile held in staging column 0
    row_off[m] = valid ? (long long)(orow - p.out) : -1;
    bool do_store = true;
    // activation / bias parameters hoisted into registers: none / relu / leaky are max(a*v, v) with a = 1 / 0 / alpha
    const float a_eff = act_slope(p.act, p.alpha);
    const bool slow_act = (MODE != 2) && p.act >= GG_ACT_TANH;
    const int act_code = p.act;
#define GG_FINISH4(o, col)                                                                        \
    do {                                                                                          \
      if (has_bias) {                                                                             \
        const float4 bb = *reinterpret_cast<const float4*>(s_bias + (col));                       \
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;                                       \
      }                                                                                           \
      if (MODE != 2) {                                                                            \
        if (!slow_act) {                                                                          \
          o.x = fmaxf(o.x * a_eff, o.x); o.y = fmaxf(o.y * a_eff, o.y);                           \
          o.z = fmaxf(o.z * a_eff, o.z); o.w = fmaxf(o.w * a_eff, o.w);                           \
        } else {                                                                                  \
          o.x = apply_act_slow(o.x, act_code); o.y = apply_act_slow(o.y, act_code);               \
          o.z = apply_act_slow(o.z, act_code); o.w = apply_act_slow(o.w, act_code);               \
        }                                                                                         \
      }                                                                                           \
    } while (0)
    int cbeg = 0, cend = p.n_tile / 4;          // float4 column range of the tile this CTA writes out
    if (p.splits == 1) {
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        float v[32];
        tmem_ld_32x32(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          GG_FINISH4(o, c0 + i);
          *reinterpret_cast<float4*>(stage_tile + m * ld + c0 + i) = o;
        }
      }
    } else if (cl) {
      // split-K inside a thread-block cluster: each CTA parks its partial tile in its OWN shared memory ([n_tile/4][128]
      // float4), signals every peer's mbarrier, and then sums its column slice straight out of the peers' shared memory
      // (DSMEM) in split order — no global round trip, no atomics, no cooperative launch.
      float4* ptile = reinterpret_cast<float4*>(smem);
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        float v[32];
        if (nkb > 0) tmem_ld_32x32(taddr + (uint32_t)c0, v);
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) ptile[((c0 + i) >> 2) * 128 + m] = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      asm volatile("fence.acq_rel.cluster;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < p.splits) mbar_arrive_remote(map_to_cta(smem_u32(&cl_bar), (uint32_t)et));
      mbar_wait_cluster(&cl_bar, 0);
      const int nc = p.n_tile / 4;
      cbeg = (split * nc) / p.splits;
      cend = ((split + 1) * nc) / p.splits;
      const uint32_t my_ptile = smem_u32(ptile);
      for (int c4 = cbeg; c4 < cend; c4 += 4) {
        float4 acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sp = 0; sp < p.splits; sp += 2) {
          float4 t[2][4];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t peer = map_to_cta(my_ptile, (uint32_t)min(sp + q, p.splits - 1));
#pragma unroll
            for (int u = 0; u < 4; ++u)
              t[q][u] = (sp + q < p.splits && c4 + u < cend) ? ld_dsmem_f4(peer + (uint32_t)(((c4 + u) * 128 + m) * 16))
                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc[u].x += t[q][u].x; acc[u].y += t[q][u].y; acc[u].z += t[q][u].z; acc[u].w += t[q][u].w;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (c4 + u < cend) {
            float4 o = acc[u];
            GG_FINISH4(o, (c4 + u) * 4);
            *reinterpret_cast<float4*>(stage_tile + m * ld + (c4 + u) * 4) = o;
          }
        }
      }
    } else {
      // split-K: park the partial tile in the (L2-resident) workspace in a [n_tile/4][128 rows] float4 layout (coalesced
      // for the thread-per-row TMEM readout); the last CTA to take a ticket on this tile sums the splits in split order.
      const size_t tile_elems = (size_t)128 * p.n_tile;
      const size_t ntiles_all = gridDim.x;
      float4* pme = reinterpret_cast<float4*>(p.partial + ((size_t)split * ntiles_all + blockIdx.x) * tile_elems) + m;
      for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
        float v[32];
        if (nkb > 0) tmem_ld_32x32(taddr + (uint32_t)c0, v);
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) pme[(size_t)((c0 + i) >> 2) * 128] = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      if (et == 0) GG_DBG(129);
      // All `splits` CTAs of this tile are co-resident (cooperative launch, grid <= #SMs): rendezvous on a ticket, then
      // EVERY CTA reduces its own 1/splits column slice of the tile (summing the splits in split order: deterministic)
      // instead of one CTA serially re-reading all partials — the reduction's L2 round trips run on `splits` SMs at once.
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        // arrival needs no return value: a reduction (fire and forget) instead of an atomic round trip before the poll
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.counters + blockIdx.x) : "memory");
        unsigned seen;
        unsigned spins = 0;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.counters + blockIdx.x) : "memory");
          if (seen < (unsigned)p.splits) { __nanosleep(32); if (++spins > (1u << 22)) __trap(); }
        } while (seen < (unsigned)p.splits);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) GG_DBG(212);
      const int nc = p.n_tile / 4;
      cbeg = (split * nc) / p.splits;
      cend = ((split + 1) * nc) / p.splits;
      if (p.bulk) {
        // The slice [cbeg, cend) x 128 rows of one split's partial tile is CONTIGUOUS in the [n_tile/4][128] float4 layout:
        // one elected thread asks the copy engine for the `splits` slices (cp.async.bulk -> the idle operand ring) and
        // the 128 epilogue threads then sum them out of shared memory in split order.  Replaces two dependent rounds of
        // 16 L2 loads per thread (2.8 us on the timeline) by one bulk transfer of <= 64 KB.
        const int ncols = cend - cbeg;
        const int ncols_max = (nc + p.splits - 1) / p.splits;
        const uint32_t chunk = (uint32_t)ncols * 2048u;
        if (ncols > 0) {
          if (et == 0) {
            fence_proxy_async_all();               // partials were written through the generic proxy (by other CTAs)
            mbar_expect_tx(&accum_bar, chunk * (uint32_t)p.splits);
            const uint8_t* src0 = reinterpret_cast<const uint8_t*>(p.partial + (size_t)blockIdx.x * tile_elems) + (size_t)cbeg * 2048;
            const size_t split_stride_bytes = ntiles_all * tile_elems * sizeof(float);
            for (int sp = 0; sp < p.splits; ++sp) bulk_g2s(smem + (size_t)sp * chunk, src0 + (size_t)sp * split_stride_bytes, chunk, &accum_bar);
          }
          mbar_wait(&accum_bar, 1);
        }
        stage_tile = reinterpret_cast<float*>(smem + (size_t)p.splits * ncols_max * 2048);
        ld = ncols * 4 + 4;
        col0 = cbeg;
        const float4* land = reinterpret_cast<const float4*>(smem) + m;
        for (int c = 0; c < ncols; ++c) {
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
          for (int sp = 0; sp < p.splits; ++sp) {
            const float4 t = land[((size_t)sp * ncols + c) * 128];
            o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
          }
          GG_FINISH4(o, (cbeg + c) * 4);
          *reinterpret_cast<float4*>(stage_tile + m * ld + c * 4) = o;
        }
      } else {
        const float4* pbase = reinterpret_cast<const float4*>(p.partial + (size_t)blockIdx.x * tile_elems) + m;
        const size_t split_stride4 = ntiles_all * tile_elems / 4;
        // up to 16 independent L2 loads in flight per thread (4 column groups x 4 splits)
        for (int c4 = cbeg; c4 < cend; c4 += 4) {
          float4 acc[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int sp = 0; sp < p.splits; sp += 4) {
            float4 t[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int u = 0; u < 4; ++u)
                t[q][u] = (sp + q < p.splits && c4 + u < cend)
                              ? __ldcg(pbase + (size_t)(sp + q) * split_stride4 + (size_t)(c4 + u) * 128)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                acc[u].x += t[q][u].x; acc[u].y += t[q][u].y; acc[u].z += t[q][u].z; acc[u].w += t[q][u].w;
              }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (c4 + u < cend) {
              float4 o = acc[u];
              GG_FINISH4(o, (c4 + u) * 4);
              *reinterpret_cast<float4*>(stage_tile + m * ld + (c4 + u) * 4) = o;
            }
          }
        }
      }
      // second ticket: the last CTA to finish reading the partials re-arms both counters for the next launch
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        const unsigned old = atomicAdd(&p.counters[gridDim.x + blockIdx.x], 1u);
        if (old == (unsigned)(p.splits - 1)) { p.counters[blockIdx.x] = 0u; p.counters[gridDim.x + blockIdx.x] = 0u; }
      }
    }
#undef GG_FINISH4
    (void)do_store;
    {
      if (p.dbg && et == 0) p.dbg[140] = gtime();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (p.dbg && et == 0) p.dbg[141] = gtime();
      // coalesced write-out of columns [cbeg, cend): `ncol` consecutive threads take the consecutive float4 of one row
      // (a warp stores whole 128..512-byte row segments); the row/column split is computed once, not per element
      const int ncol = cend - cbeg;
      if (ncol > 0) {
        const int rpp = 128 / ncol;                 // rows per pass of the 128 threads
        const int my_c = et % ncol, my_r = et / ncol;
        if (my_r < rpp) {
          const float* sp_ = stage_tile + (cbeg - col0 + my_c) * 4;
          float* gp_ = p.out + (size_t)(cbeg + my_c) * 4;
          for (int r = my_r; r < 128; r += 4 * rpp) {
            float4 val[4];
            long long off[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int rr = r + u * rpp;
              off[u] = rr < 128 ? row_off[rr] : -1;
              val[u] = *reinterpret_cast<const float4*>(sp_ + (rr < 128 ? rr : 0) * ld);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (off[u] >= 0) *reinterpret_cast<float4*>(gp_ + off[u]) = val[u];
          }
        }
      }
      if (p.dbg && et == 0) p.dbg[203] = gtime();
    }
  }
  if (threadIdx.x == 64) GG_DBG(131);
  if (p.dbg && threadIdx.x == 64) atomicMax(reinterpret_cast<unsigned long long*>(p.dbg + 200), (unsigned long long)gtime());
  // ---- teardown -----------------------------------------------------------------------------------------
  tc_fence_before();
  __syncwarp();
  if (cl) cluster_sync_all();                  // no CTA may exit (and free its shared memory) while a peer still reads it
  else __syncthreads();
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
  // a batch box larger than / not dividing the batch is fine: TMA zero-fills the missing images and the epilogue masks
  // their rows (this is how M = 64-row dense layers ride the 128-row MMA)
  if (b <= 0 || w * h * b != target) return false;
  if (w > 256 || h > 256 || b > 256) return false;
  *wt = w; *ht = h; *bt = b;
  return true;
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

int pick_n_tile(int n) {
  const int forced = env_int("GG_TC_NTILE", 0);   // experiment knob
  if (forced > 0 && n % forced == 0) return forced;
  if (n % 128 == 0) return 128;
  if (n % 64 == 0) return 64;
  if (n % 32 == 0) return 32;
  return 0;
}

int g_tc_max_ctas = kNumSMs;
int g_tc_stage_cap = 0;

int pick_splits(int tiles, int kb_min) {
  int forced = env_int("GG_TC_SPLITS", 0);   // experiment knob
  if (forced > 0) return forced;
  const int max_ctas = g_tc_max_ctas;       // < 148 leaves SMs for kernels running concurrently on other streams
  int s = max_ctas / tiles;                  // fill (at most) one wave of CTAs: 1 CTA per SM (shared-memory bound)
  int cap = kb_min / 3;                      // keep >= 3 k-blocks per CTA so the pipeline fills
  if (cap < 1) cap = 1;
  if (s > cap) s = cap;
  static int use_cluster = env_int("GG_TC_CLUSTER", 0);
  if (s > (use_cluster ? 8 : 16)) s = use_cluster ? 8 : 16;   // 8 = portable thread-block-cluster size
  if (s < 1) s = 1;
  return s;
}

long long* g_dbg = nullptr;
thread_local int g_last_info[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // gg_last_tc_info

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
    p.tw = Wo / p.wt; p.th = Ho / p.ht; p.tb = (B + p.bt - 1) / p.bt;
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
    p.tw = p.PW / p.wt; p.th = p.PH / p.ht; p.tb = (B + p.bt - 1) / p.bt;
    p.m_tiles = p.tw * p.th * p.tb;
    pl.grid_x = stride * stride * p.m_tiles * p.n_tiles;
    int nmin = k / stride;                    // fewest taps a class sees along one axis
    if (nmin < 1) return pl;
    kb_min = nmin * nmin * (Co / 32);
  } else {
    p.PH = Ho; p.PW = Wo;
    if (!pixel_box(Wo, Ho, B, 32, &p.wt, &p.ht, &p.bt) || B % p.bt != 0) return pl;
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
  pl.counter_bytes = ((size_t)2 * pl.grid_x * sizeof(unsigned) + 255) & ~size_t(255);   // arrival + completion tickets
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
  p.dbg = g_dbg;
  p.partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + pl.counter_bytes);
  p.stages = kMaxStages;
  if ((long long)pl.grid_x * p.splits > kNumSMs) {
    // more CTAs than SMs: keep the footprint under half an SM so two CTAs co-reside and one's epilogue overlaps the
    // other's main loop
    int fit = (int)((113 * 1024 - 1024) / (kABytes + p.n_tile * 128));
    p.stages = fit < 3 ? 3 : (fit > kMaxStages ? kMaxStages : fit);
  }
  // ring-depth cap (gg_set_tc_stages): 3 stages = 97 KB, so two CTAs — of the same or of two concurrent launches on
  // different streams — share an SM; the main loop is TMA-issue-bound, not latency-bound, and loses nothing
  if (g_tc_stage_cap >= 2 && g_tc_stage_cap < p.stages) p.stages = g_tc_stage_cap;
  // Thread-block-cluster / DSMEM reduction (GG_TC_CLUSTER=1): correct (same tests pass) but measured no faster than the
  // L2 workspace rendezvous on B200 (E.2 fwd 13.7 vs 14.2 us, E.3 fwd 26 vs 13 us: 8-CTA clusters of 197 KB CTAs place
  // badly on 16-20-SM GPCs), so it is off by default.
  p.cluster = (p.splits > 1 && p.splits <= 8 && env_int("GG_TC_CLUSTER", 0) != 0) ? 1 : 0;
  if (p.cluster) p.stages = kMaxStages;        // the cluster epilogue parks a partial tile AND a staging tile in the ring
  // the epilogue re-uses the ring as a [128][n_tile+4] staging tile; split-K with the bulk-copy reduction lands the
  // `splits` slices ([splits][ncols][128] float4, <= 64 KB + rounding) in front of a compact [128][4*ncols+4] staging tile
  // (GG_TC_BULK=1, opt-in: measured SLOWER than the per-thread L2 loads on B200 — rendezvous -> staged 3.7 us vs 2.8 us for
  // the E.2 forward conv, profiles/timeline_conv_r1.txt — the 1-D bulk copies of 16 KB slices do not beat 16 loads in flight
  // per thread here; kept as a documented negative result.)
  p.bulk = (p.splits > 1 && !p.cluster && env_int("GG_TC_BULK", 0) != 0) ? 1 : 0;
  size_t ring = (size_t)p.stages * (kABytes + p.n_tile * 128);
  size_t staging = (size_t)128 * (p.n_tile + 4) * 4;
  if (p.bulk) {
    const int nc = p.n_tile / 4, ncols_max = (nc + p.splits - 1) / p.splits;
    staging = (size_t)p.splits * ncols_max * 2048 + (size_t)128 * (ncols_max * 4 + 4) * 4;
  }
  if (ring < staging) ring = staging;
  const size_t smem = ring + 1024;
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[MODE]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kMaxStages * (kABytes + kMaxNTile * 128) + 1024);
    if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "conv_tc: cudaFuncSetAttribute failed%s");
    attr_set[MODE] = true;
  }
  dim3 grid(pl.grid_x, p.splits);
  g_last_info[0] = MODE; g_last_info[1] = pl.grid_x; g_last_info[2] = p.splits; g_last_info[3] = p.n_tile;
  g_last_info[4] = p.stages; g_last_info[5] = p.cluster; g_last_info[6] = (int)smem; g_last_info[7] = p.m_tiles;
  if (p.cluster) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = (unsigned)p.splits;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<MODE>, tmA, tmB, p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(GG_ERR_CUDA_BASE + (int)e, "conv_tc: cluster launch failed: %s", cudaGetErrorString(e)); }
  } else if (p.splits > 1) {
    // the split-K rendezvous spins on a ticket: all CTAs of the grid must be co-resident -> cooperative launch
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_pdl && env_int("GG_PDL_COOP", 0) != 0) ? 2 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<MODE>, tmA, tmB, p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(GG_ERR_CUDA_BASE + (int)e, "conv_tc: cooperative launch failed: %s", cudaGetErrorString(e)); }
  } else {
    GG_LAUNCH((conv_tc_kernel<MODE>), grid, kThreads, smem, st, tmA, tmB, p);
  }
  return check_launch(MODE == 0 ? "gg_conv2d_fwd(tcgen05)" : (MODE == 1 ? "gg_conv2d_dgrad(tcgen05)" : "gg_conv2d_wgrad(tcgen05)"));
}

// filter tensor map: w [taps][Ci][Co] viewed as dims (Co, Ci, taps)
// MN-major B stage in ONE TMA: Co is split into (32 co, Co/32 blocks) so the box (32 co, 32 ci, n_tile/32 blocks, 1 tap)
// lands as [co block][ci][32 co] = n_tile/32 swizzle atoms of 4 KB, LBO = 4096
int filter_map_mn(CUtensorMap* tm, const float* w, int Ci, int Co, int taps, int n_tile, int rows_valid = 0) {
  // rows_valid < Ci: the filter matrix in memory has fewer rows than the (im2col-padded) K the kernel walks; the
  // missing rows are TMA out-of-bounds = zeros
  uint64_t dims[4] = {32, (uint64_t)(rows_valid ? rows_valid : Ci), (uint64_t)(Co / 32), (uint64_t)taps};
  uint64_t str[4] = {1, (uint64_t)Co, 32, (uint64_t)Ci * Co};
  uint32_t box[4] = {32, 32, (uint32_t)(n_tile / 32), 1};
  return encode_tmap(tm, w, 4, dims, str, box, nullptr, 2, env_int("GG_TC_NOCVT", 0) == 0);
}
int filter_map_k(CUtensorMap* tm, const float* w, int Ci, int Co, int taps, int n_tile, int rows_valid = 0) {  // 32 co x n_tile ci rows, K-major
  uint64_t dims[3] = {(uint64_t)Co, (uint64_t)(rows_valid ? rows_valid : Ci), (uint64_t)taps};
  uint64_t str[3] = {1, (uint64_t)Co, (uint64_t)Ci * Co};
  uint32_t box[3] = {32, (uint32_t)n_tile, 1};
  return encode_tmap(tm, w, 3, dims, str, box, nullptr, 1, env_int("GG_TC_NOCVT", 0) == 0);
}
// activation tensor map over an NHWC tensor (C,W,H,B); box = 32 channels x (wt,ht,bt) pixels gathered with stride es
int act_map(CUtensorMap* tm, const float* x, int B, int H, int W, int C, int wt, int ht, int bt, int es, int swizzle) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t str[4] = {1, (uint64_t)C, (uint64_t)W * C, (uint64_t)H * W * C};
  uint32_t box[4] = {32, (uint32_t)(wt * es), (uint32_t)(ht * es), (uint32_t)bt};
  uint32_t estr[4] = {1, (uint32_t)es, (uint32_t)es, 1};
  return encode_tmap(tm, x, 4, dims, str, box, estr, swizzle, env_int("GG_TC_NOCVT", 0) == 0);
}

}  // namespace

int conv_tc_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Ci, int Co, int k,
                int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, void* ws, size_t ws_bytes,
                cudaStream_t st, bool* handled, int filt_rows) {
  *handled = false;
  TcPlan pl = make_plan(0, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
  if (!pl.ok) return GG_OK;
  if (ws == nullptr || ws_bytes < pl.counter_bytes) return GG_OK;     // needs at least the ticket area
  CUtensorMap tmA, tmB;
  int rc = act_map(&tmA, x, B, H, W, Ci, pl.p.wt, pl.p.ht, pl.p.bt, stride, 1);
  if (rc) return rc;
  rc = filter_map_mn(&tmB, w, Ci, Co, k * k, pl.p.n_tile, filt_rows);
  if (rc) return rc;
  pl.p.out = y; pl.p.bias = bias; pl.p.act = act; pl.p.alpha = alpha;
  rc = launch<0>(tmA, tmB, pl, ws, ws_bytes, st);
  if (rc) return rc;
  *handled = true;
  return GG_OK;
}

int conv_tc_dgrad(const float* dy, const float* w, const float* bias, float* dx, int B, int H, int W, int Ci, int Co, int k,
                  int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, void* ws, size_t ws_bytes,
                  cudaStream_t st, bool* handled, int filt_rows) {
  *handled = false;
  TcPlan pl = make_plan(1, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
  if (!pl.ok) return GG_OK;
  if (ws == nullptr || ws_bytes < pl.counter_bytes) return GG_OK;
  CUtensorMap tmA, tmB;
  int rc = act_map(&tmA, dy, B, Ho, Wo, Co, pl.p.wt, pl.p.ht, pl.p.bt, 1, 1);
  if (rc) return rc;
  rc = filter_map_k(&tmB, w, Ci, Co, k * k, pl.p.n_tile, filt_rows);
  if (rc) return rc;
  pl.p.out = dx; pl.p.bias = bias; pl.p.act = act; pl.p.alpha = alpha;
  rc = launch<1>(tmA, tmB, pl, ws, ws_bytes, st);
  if (rc) return rc;
  *handled = true;
  return GG_OK;
}

int conv_tc_wgrad(const float* x, const float* dy, float* dw, int B, int H, int W, int Ci, int Co, int k, int stride,
                  int pad_t, int pad_l, int Ho, int Wo, void* ws, size_t ws_bytes, cudaStream_t st, bool* handled, int out_rows) {
  *handled = false;
  TcPlan pl = make_plan(2, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
  pl.p.out_rows = out_rows ? out_rows : k * k * Ci;
  if (!pl.ok) return GG_OK;
  if (ws == nullptr || ws_bytes < pl.counter_bytes) return GG_OK;
  CUtensorMap tmA, tmB;
  int rc = act_map(&tmA, x, B, H, W, Ci, pl.p.wt, pl.p.ht, pl.p.bt, stride, 2);
  if (rc) return rc;
  {
    uint64_t dims[3] = {32, (uint64_t)B * Ho * Wo, (uint64_t)(Co / 32)};
    uint64_t str[3] = {1, (uint64_t)Co, 32};
    uint32_t box[3] = {32, 32, (uint32_t)(pl.p.n_tile / 32)};
    rc = encode_tmap(&tmB, dy, 3, dims, str, box, nullptr, 2, env_int("GG_TC_NOCVT", 0) == 0);
    if (rc) return rc;
  }
  pl.p.out = dw; pl.p.bias = nullptr; pl.p.act = 0; pl.p.alpha = 0.f;
  rc = launch<2>(tmA, tmB, pl, ws, ws_bytes, st);
  if (rc) return rc;
  *handled = true;
  return GG_OK;
}

void conv_tc_last_info(int* out8) { for (int i = 0; i < 8; ++i) out8[i] = g_last_info[i]; }

void conv_tc_set_stage_cap(int n) { g_tc_stage_cap = n; }

void conv_tc_set_max_ctas(int n) { g_tc_max_ctas = n < 1 ? 1 : (n > kNumSMs ? kNumSMs : n); }

void conv_tc_set_debug(void* p) { g_dbg = reinterpret_cast<long long*>(p); }

size_t conv_tc_workspace(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo) {
  TcPlan pl = make_plan(mode, B, H, W, Ci, Co, k, stride, 0, 0, Ho, Wo);
  return plan_workspace(pl);
}

size_t conv_tc_wgrad_workspace(int B, int H, int W, int Ci, int Co, int k, int stride, int Ho, int Wo) {
  return conv_tc_workspace(2, B, H, W, Ci, Co, k, stride, Ho, Wo);
}

}  // namespace gg
