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
//
// Roles (352 threads): warps 0 and 10 = TMA producers (even / odd K blocks: the issue loop of ONE lane — barrier wait,
// expect_tx, two tensor-map loads, ~400 cycles of dependent uniform-datapath latency — was still the pace of the main
// loop at 0.21 us per block, above the tensor pipe's 0.135 us), warp 1 = MMA issuer, warps 2-9 = epilogue.  Round 2 (profiles/
// timeline_conv_r2_baseline.txt) showed every phase of this kernel bound by instruction LATENCY of too few threads rather
// than by bytes or flops: the producer / issuer loops cost ~0.3 us per 32-wide K block regardless of tile bytes (single
// divergent lane, shared-memory addresses re-derived per instruction), the 4-warp epilogue took 1.7 us to move a 64 KB
// accumulator, and the split-K exchange through an L2 workspace (write, fence, ticket, re-read) 6-8 us.  Now:
//   * the producer / issuer loops run warp-convergent with one elected lane issuing, on precomputed 32-bit shared
//     addresses and descriptor halves;
//   * eight epilogue warps (two per TMEM lane quadrant, each taking half of the tile's columns);
//   * split-K runs inside a thread-block cluster: the `splits` CTAs of a tile park their partial tile in their own shared
//     memory and PUSH every peer's column slice with one bulk copy each (cp.async.bulk shared::cta -> shared::cluster,
//     completion counted on the owner's mbarrier); every CTA then reduces its own slice in split order (deterministic)
//     out of local shared memory.  No global workspace, no atomics, no cooperative launch.
#include "gg_tc_common.cuh"

#include <map>
#include <tuple>

namespace gg {

namespace {

constexpr int kMaxStages = 6;                  // 6 x 32 KB operand stages + slack fit the 227 KB of an SM
constexpr int kABytes = 128 * 32 * 4;          // 16 KB: 128 rows (or 4 x 32 MN-blocks) of 32 fp32
constexpr int kMaxNTile = 128;
constexpr int kEpiThreads = 256;               // warps 2..9
constexpr int kThreads = 64 + kEpiThreads + 32; // + warp 10: the second TMA producer
constexpr int kMaxSplits = 8;                  // portable thread-block-cluster size
constexpr int kMaxDynSmem = 227 * 1024 - 4096; // dynamic shared memory opt-in (the 227 KB limit includes ~3 KB of static)

struct TcParams {
  int B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo;
  int wt, ht, bt;             // pixel box of the A tile (fwd/dgrad: wt*ht*bt = 128; wgrad: = 32)
  int tw, th, tb;             // tiles along each pixel dimension
  int PH, PW;                 // pixel grid the tiles walk (fwd: Ho,Wo; dgrad: class grid; wgrad: Ho,Wo)
  int n_tile, n_tiles;        // GEMM N tiling
  int m_tiles;                // fwd/dgrad: tw*th*tb (per class); wgrad: ceil(taps*Ci/128)
  int splits;                 // K splits = cluster size along grid.y
  int stages;                 // depth of the operand ring
  int stage_off;              // split-K: byte offset of the compact staging tile inside the (idle) ring
  int land_off;               // split-K: byte offset of the landing slots (dedicated shared memory behind the ring)
  int ncols_max;              // split-K: float4 columns of the widest owner slice
  int act;
  float alpha;
  float* out;
  const float* bias;
  const float* mask;          // optional: forward activation y at the OUTPUT positions (same layout as out); the stored value
  int mask_act;               // becomes act'(y) * value — the activation gradient that follows a dgrad, fused into its write-out
  float mask_alpha;
  float* stats;               // optional: per-m-tile column sums / sums of squares of the stored values, [m-tile][2][stats_ld]:
  int stats_ld;               // the batch-norm statistics of the layer that follows (tflib/ops/batchnorm.py:29-30), taken in the
                              // epilogue of the kernel that produces its input; folded by gg_bn_apply
  long long* dbg;             // optional timeline of CTA (0,0): see gg_debug_set_buffer
  int out_rows;               // wgrad: rows of the [taps*Ci, Co] result that exist in memory (im2col-padded K)
};

// ---- PTX wrappers on 32-bit shared-window addresses (computed once per kernel, not per instruction) -----------------
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 24)) __trap();      // a mis-programmed pipeline traps instead of hanging the GPU box
  } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
// shared::cta -> shared::cluster bulk copy (the copy engine moves the slice; completion = bytes counted on the destination
// CTA's mbarrier)
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// explicit shared-space accesses on 32-bit addresses: the epilogue's pointers are derived from an integer-rounded base, so
// the compiler would otherwise emit generic LD.E / ST.E for them
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns (n_tile = 32: each of the two warps of a quadrant takes 16 columns)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// shared-memory operand descriptor halves: lo = start address | leading byte offset, hi = stride byte offset | version | layout
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

#ifdef GG_TIMELINE
#define GG_DBG(slot) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0) p.dbg[slot] = gtime(); } while (0)
#else
#define GG_DBG(slot) do { } while (0)
#endif

// min-blocks 2 for fwd / dgrad: ptxas then keeps the register count where two 352-thread CTAs fit an SM (un-split launches with
// more tiles than SMs co-reside two CTAs so that one's epilogue overlaps the other's main loop); wgrad launches are always split
template <int MODE>
__global__ void __launch_bounds__(kThreads, MODE == 2 ? 1 : 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], accum_bar, cl_bar;
  __shared__ uint32_t tmem_base_sh;
  __shared__ long long s_row_off[128];                  // element offset of each accumulator row's output row (-1: masked)
  __shared__ __align__(16) float s_bias[kMaxNTile];     // bias slice of this n-tile (fetched while the main loop runs)
#ifdef GG_TIMELINE
  if (p.dbg && threadIdx.x == 0) {             // timeline (tools/timeline_conv.py): earliest CTA entry of the launch
    const long long t_in = gtime();
    atomicMin(reinterpret_cast<unsigned long long*>(p.dbg + 210), (unsigned long long)t_in);
    if (blockIdx.x == 0 && blockIdx.y == 0) p.dbg[211] = t_in;
  }
#endif
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: the next kernel may start its own prologue now

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kStages = p.stages;
  const int stage_bytes = kABytes + p.n_tile * 128;
  const int cblocks = (MODE == 1 ? p.Co : p.Ci) / 32;       // 32-channel K blocks (fwd: ci, dgrad: co)
  const int taps = p.k * p.k;
  const uint32_t ring = smem_u32(smem);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t accum_a = smem_u32(&accum_bar), cl_a = smem_u32(&cl_bar);

  // ---- tile decode -------------------------------------------------------------------------------------
  const int tile = blockIdx.x;           // over (class,) m-tile, n-tile
  const int split = blockIdx.y;          // = rank inside the cluster (cluster dims (1, splits, 1))
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

  // ---- producer start coordinates (integer divisions: computed by every thread BEFORE the setup barrier so that they overlap
  //      barrier initialisation and the TMEM allocation instead of delaying the first TMA issue) --------------------------
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
  const int row_len = (MODE == 0) ? p.k : ns;
  const int xw0 = (MODE == 0) ? w0 * p.stride - p.pad_l : w0 + d_w;
  const int xh0 = (MODE == 0) ? h0 * p.stride - p.pad_t : h0 + d_h;

  // ---- one-time setup ------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    mbar_init(&cl_bar, 1);                       // split-K: this CTA's own arrive.expect_tx; peers only complete bytes
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  const uint32_t tmem_cols = p.n_tile <= 32 ? 32 : (p.n_tile <= 64 ? 64 : 128);
  if (warp == 2) tmem_alloc(&tmem_base_sh, tmem_cols);
  tc_fence_before();
  const bool cl = p.splits > 1;
  __syncthreads();
  // split-K: peers must not signal cl_bar before it is initialised.  Only ARRIVE here; the matching wait sits right in
  // front of the first remote access (after the main loop), so the cluster rendezvous costs nothing on the way in.
  if (cl) cluster_arrive();
  tc_fence_after();
  // PDL: everything above is independent of the producer kernel; its activations / gradients are first touched below
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_d = tmem_base_sh;
  if (threadIdx.x == 0) GG_DBG(0);
#ifdef GG_TIMELINE
  if (p.dbg && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long*>(p.dbg + 201), (unsigned long long)gtime());
#endif

  if (warp == 0 || warp == 10) {
    // =========================== TMA producers ===========================
    // The whole warp walks the loop (warp-uniform control flow and coordinates); one elected lane issues.  Coordinates
    // advance incrementally: no integer divisions in the issue loop.  Warp 0 takes the even K blocks, warp 10 the odd ones.
    const int prod = warp == 0 ? 0 : 1;
    const bool leader = elect_one();
#define GG_ADVANCE()                                                                                             \
    do {                                                                                                         \
      if (MODE == 2) {                                                                                           \
        px_w += p.wt;                                                                                            \
        if (px_w == p.PW) { px_w = 0; px_h += p.ht; if (px_h == p.PH) { px_h = 0; px_b += p.bt; } }              \
      } else {                                                                                                   \
        if (++c_cb == cblocks) {                                                                                 \
          c_cb = 0;                                                                                              \
          if (++c_s == row_len) { c_s = 0; ++c_r; }                                                              \
        }                                                                                                        \
      }                                                                                                          \
    } while (0)
    int s = prod;                                     // kStages >= 2
    uint32_t ph = 0;
    if (prod == 1) GG_ADVANCE();
    for (int i = prod; i < nkb; i += 2) {
      mbar_wait_a(empty0 + 8u * s, ph ^ 1);           // a fresh barrier passes a parity-1 wait: first lap never blocks
      if (leader) {
        const uint32_t sA = ring + (uint32_t)(s * stage_bytes), sB = sA + kABytes, fb = full0 + 8u * s;
        if (i < 60) GG_DBG(1 + i);
        mbar_expect_tx_a(fb, (uint32_t)stage_bytes);
        if (MODE == 0) {
          tma_load_4d_a(sA, &tmA, fb, c_cb * 32, xw0 + c_s, xh0 + c_r, b0);
          tma_load_4d_a(sB, &tmB, fb, 0, c_cb * 32, n0 / 32, c_r * p.k + c_s);
        } else if (MODE == 1) {
          tma_load_4d_a(sA, &tmA, fb, c_cb * 32, xw0 - c_s, xh0 - c_r, b0);
          tma_load_3d_a(sB, &tmB, fb, c_cb * 32, n0, (a_h + p.stride * c_r) * p.k + a_w + p.stride * c_s);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_4d_a(sA + j * 4096, &tmA, fb, qa_c[j], px_w * p.stride + qa_w[j], px_h * p.stride + qa_h[j], px_b);
          tma_load_3d_a(sB, &tmB, fb, 0, (kb0 + i) * 32, n0 / 32);
        }
      }
      GG_ADVANCE();
      GG_ADVANCE();
      s += 2;
      if (s >= kStages) { s -= kStages; ph ^= 1; }
    }
#undef GG_ADVANCE
    __syncwarp();
    if (cl) { cluster_wait(); cluster_arrive(); }      // phase 1 (start-up rendezvous) consumed; arrive for the exit rendezvous
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr int a_mn = (MODE == 2) ? 1 : 0;
    constexpr int b_mn = (MODE == 1) ? 0 : 1;
    const uint32_t idesc = make_idesc_tf32(128, p.n_tile, a_mn, b_mn);
    // K-major operand: 128B swizzle, LBO 16 B (unused), SBO 1024 B, k-step = +32 B inside the swizzle atom
    // MN-major operand: 128B swizzle with 32-byte atoms, LBO 4096 B (next 32-wide MN block), SBO 512 B, k-step = +1024 B
    const uint32_t a_hi = a_mn ? desc_hi(512, kLayoutSW128_32B) : desc_hi(1024, kLayoutSW128);
    const uint32_t b_hi = b_mn ? desc_hi(512, kLayoutSW128_32B) : desc_hi(1024, kLayoutSW128);
    const uint32_t a_lbo = a_mn ? 4096u : 16u, b_lbo = b_mn ? 4096u : 16u;
    const uint32_t a_step = a_mn ? (1024u >> 4) : (32u >> 4), b_step = b_mn ? (1024u >> 4) : (32u >> 4);
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait_a(full0 + 8u * s, ph);
      tc_fence_after();
      if (leader) {
        if (i < 60) GG_DBG(64 + i);
        const uint32_t aBase = ring + (uint32_t)(s * stage_bytes);
        uint32_t a_lo = desc_lo(aBase, a_lbo), b_lo = desc_lo(aBase + kABytes, b_lbo);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          umma_tf32(tmem_d, desc64(a_lo, a_hi), desc64(b_lo, b_hi), idesc, (i | j) != 0);
          a_lo += a_step;
          b_lo += b_step;
        }
        umma_commit_a(empty0 + 8u * s);       // frees the smem stage once these MMAs have read it
      }
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
    if (leader) umma_commit_a(accum_a);       // accumulator complete
    __syncwarp();
    if (cl) { cluster_wait(); cluster_arrive(); }
  } else {
    // =========================== epilogue (warps 2..9) ===========================
    const int et = threadIdx.x - 64;              // 0..255
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;             // which half of the tile's columns this warp reads out of TMEM
    const int m = quad * 32 + lane;               // accumulator row
    const int hcols = p.n_tile >> 1;              // 64 / 32 / 16 columns per warp
    const int cw0 = half * hcols;
    // While the main loop runs these warps are idle: fetch the bias slice and the output-row table now.
    const bool has_bias = (MODE != 2) && p.bias != nullptr;
    if (has_bias && et < p.n_tile) s_bias[et] = p.bias[n0 + et];
    if (half == 0) {
      bool valid;
      long long off;
      if (MODE == 0) {
        const int wi = m % p.wt, hi = (m / p.wt) % p.ht, bi = m / (p.wt * p.ht);
        valid = (b0 + bi) < p.B;
        off = ((long long)(((long long)(b0 + bi) * p.Ho + h0 + hi) * p.Wo + w0 + wi)) * p.Co + n0;
      } else if (MODE == 1) {
        const int wi = m % p.wt, hi = (m / p.wt) % p.ht, bi = m / (p.wt * p.ht);
        const int hh = (h0 + hi) * p.stride + e_h, ww = (w0 + wi) * p.stride + e_w;
        valid = (b0 + bi) < p.B && hh < p.H && ww < p.W;
        off = ((long long)(((long long)(b0 + bi) * p.H + hh) * p.W + ww)) * p.Ci + n0;
      } else {
        const int row = mt * 128 + m;
        valid = row < p.out_rows;
        off = (long long)row * p.Co + n0;
      }
      s_row_off[m] = valid ? off : -1;
    }
    epi_bar();
    mbar_wait_a(accum_a, 0);
    tc_fence_after();
    if (et == 0) GG_DBG(128);
    const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16);
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
    // The operand ring is idle once accum_bar has fired: it becomes the staging area of the epilogue.
    uint32_t stage_a;                             // shared address of the staging tile
    int ld;                                       // staging row pitch (floats)
    int cbeg = 0, cend = p.n_tile / 4;            // float4 column range of the tile this CTA writes out
    if (!cl) {
      // un-split: TMEM -> registers -> bias + activation -> [128][n_tile+4] staging tile
      stage_a = ring;
      ld = p.n_tile + 4;
      if (hcols >= 32) {
        for (int c0 = cw0; c0 < cw0 + hcols; c0 += 32) {
          float v[32];
          tmem_ld_32x32(taddr + (uint32_t)c0, v);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            GG_FINISH4(o, c0 + i);
            sts128(stage_a + (uint32_t)(m * ld + c0 + i) * 4u, o);
          }
        }
      } else {
        float v[16];
        tmem_ld_32x16(taddr + (uint32_t)cw0, v);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          GG_FINISH4(o, cw0 + i);
          sts128(stage_a + (uint32_t)(m * ld + cw0 + i) * 4u, o);
        }
      }
    } else {
      // ---- split-K inside the cluster -------------------------------------------------------------------
      // 1. park this CTA's partial tile in its own shared memory as [n_tile/4][128 rows] float4 (conflict-free for the
      //    thread-per-row TMEM read-out; an owner's column slice is one contiguous block) and
      // 2. PUSH it chunk by chunk: as soon as the four quadrant warps of a column half have parked a 32- (16-) column chunk,
      //    one of their lanes hands every peer the part of the chunk that lies in ITS slice to the copy engine, so the
      //    transfer over the SM-to-SM network runs under the rest of the TMEM read-out
      const int nc = p.n_tile / 4;
      cbeg = (split * nc) / p.splits;
      cend = ((split + 1) * nc) / p.splits;
      const int ncols = cend - cbeg;
      const uint32_t slot_bytes = (uint32_t)p.ncols_max * 2048u;
      cluster_wait();                             // start-up rendezvous: every peer's cl_bar is initialised
      if (et == 0) mbar_expect_tx_a(cl_a, (uint32_t)(p.splits - 1) * (uint32_t)ncols * 2048u);
      const bool pusher = ((warp - 2) & 3) == 0 && lane == 0;
      const int cw_chunk = hcols >= 32 ? 32 : 16;
      for (int c0 = cw0; c0 < cw0 + hcols; c0 += cw_chunk) {
        if (hcols >= 32) {
          float v[32];
          if (nkb > 0) tmem_ld_32x32(taddr + (uint32_t)c0, v);
          else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            sts128(ring + (uint32_t)(((c0 + i) >> 2) * 128 + m) * 16u, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        } else {
          float v[16];
          if (nkb > 0) tmem_ld_32x16(taddr + (uint32_t)c0, v);
          else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            sts128(ring + (uint32_t)(((c0 + i) >> 2) * 128 + m) * 16u, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        }
        fence_proxy_async();                      // generic-proxy writes -> visible to the copy engine
        // the 4 quadrant warps of this column half.  IMMEDIATE barrier ids: with a register operand ptxas reserves all 16 named
        // barriers of the SM for every CTA ("used 16 barriers"), which silently capped the kernel at ONE CTA per SM — the
        // 3-stage / 75-97 KB configurations meant to put two CTAs on an SM never did (profiles/timeline_face_r2.txt: 512 tiles
        // took 3.5 waves of 148)
        if (half == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
        else asm volatile("bar.sync 3, 128;" ::: "memory");
        if (pusher) {
          const int c4lo = c0 >> 2, c4hi = (c0 + cw_chunk) >> 2;
          for (int j = 0; j < p.splits; ++j) {
            if (j == split) continue;
            const int cb = (j * nc) / p.splits, ce = ((j + 1) * nc) / p.splits;
            const int lo = max(cb, c4lo), hi = min(ce, c4hi);
            if (lo >= hi) continue;
            const int slot = split < j ? split : split - 1;           // slot of source `split` at owner j
            const uint32_t dst = map_to_cta(ring + (uint32_t)p.land_off + (uint32_t)slot * slot_bytes + (uint32_t)(lo - cb) * 2048u, (uint32_t)j);
            bulk_s2c(dst, ring + (uint32_t)lo * 2048u, (uint32_t)(hi - lo) * 2048u, map_to_cta(cl_a, (uint32_t)j));
          }
        }
      }
      if (et == 0) GG_DBG(129);
      // 3. wait for the (splits-1) incoming slices (complete_tx on this CTA's own barrier: the same visibility contract as
      //    a TMA load, so a CTA-scope wait suffices — a cluster-scope acquire would flush the L1), tell the cluster that
      //    nothing targets this CTA any more (exit rendezvous), then reduce the own slice in split order (deterministic)
      mbar_wait_a(cl_a, 0);
      cluster_arrive();
      if (et == 0) GG_DBG(212);
      stage_a = ring + (uint32_t)p.stage_off;
      ld = ncols * 4 + 4;
      const uint32_t land_a = ring + (uint32_t)p.land_off;
      // the split count is a compile-time constant inside the loop (the generic predicated form cost ~130 instructions per
      // float4 on 8 warps: 0.3 us per item on the timeline)
#define GG_REDUCE_LOOP(NSP)                                                                                    \
      for (int idx = et; idx < ncols * 128; idx += 2 * kEpiThreads) {                                          \
        const int idx2 = idx + kEpiThreads;                                                                    \
        const bool two = idx2 < ncols * 128;                                                                   \
        const int c = idx >> 7, row = idx & 127, c2 = two ? idx2 >> 7 : c, row2 = two ? idx2 & 127 : row;      \
        const uint32_t own = ring + (uint32_t)((cbeg + c) * 128 + row) * 16u;                                  \
        const uint32_t inc = land_a + (uint32_t)(c * 128 + row) * 16u;                                         \
        const uint32_t own2 = ring + (uint32_t)((cbeg + c2) * 128 + row2) * 16u;                               \
        const uint32_t inc2 = land_a + (uint32_t)(c2 * 128 + row2) * 16u;                                      \
        float4 t[NSP], u[NSP];                                                                                 \
        _Pragma("unroll")                                                                                      \
        for (int sp = 0; sp < NSP; ++sp) {                                                                     \
          const uint32_t so = (uint32_t)(sp < split ? sp : sp - 1) * slot_bytes;                               \
          t[sp] = lds128(sp == split ? own : inc + so);                                                        \
          u[sp] = lds128(sp == split ? own2 : inc2 + so);                                                      \
        }                                                                                                      \
        float4 o = t[0], o2 = u[0];                                                                            \
        _Pragma("unroll")                                                                                      \
        for (int sp = 1; sp < NSP; ++sp) {                                                                     \
          o.x += t[sp].x; o.y += t[sp].y; o.z += t[sp].z; o.w += t[sp].w;                                      \
          o2.x += u[sp].x; o2.y += u[sp].y; o2.z += u[sp].z; o2.w += u[sp].w;                                  \
        }                                                                                                      \
        GG_FINISH4(o, (cbeg + c) * 4);                                                                         \
        GG_FINISH4(o2, (cbeg + c2) * 4);                                                                       \
        sts128(stage_a + (uint32_t)(row * ld + c * 4) * 4u, o);                                                \
        if (two) sts128(stage_a + (uint32_t)(row2 * ld + c2 * 4) * 4u, o2);                                    \
      }
      switch (p.splits) {
        case 2: GG_REDUCE_LOOP(2) break;
        case 3: GG_REDUCE_LOOP(3) break;
        case 4: GG_REDUCE_LOOP(4) break;
        case 5: GG_REDUCE_LOOP(5) break;
        case 6: GG_REDUCE_LOOP(6) break;
        case 7: GG_REDUCE_LOOP(7) break;
        default: GG_REDUCE_LOOP(8) break;
      }
#undef GG_REDUCE_LOOP
    }
#undef GG_FINISH4
    if (p.dbg && et == 0) p.dbg[140] = gtime();
    epi_bar();
    // coalesced write-out of columns [cbeg, cend): `ncol` consecutive threads take the consecutive float4 of one row
    // (a warp stores whole 128..512-byte row segments)
    {
      const int ncol = cend - cbeg;
      if (ncol > 0) {
        const int rpp = kEpiThreads / ncol;         // rows per pass of the 256 threads
        const int my_c = et % ncol, my_r = et / ncol;
        if (my_r < rpp) {
          const uint32_t sp_ = stage_a + (uint32_t)my_c * 16u;
          float* gp_ = p.out + (size_t)(cbeg + my_c) * 4;
          for (int r = my_r; r < 128; r += 4 * rpp) {
            float4 val[4];
            long long off[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int rr = r + u * rpp;
              off[u] = rr < 128 ? s_row_off[rr] : -1;
              val[u] = lds128(sp_ + (uint32_t)((rr < 128 ? rr : 0) * ld) * 4u);
            }
            if (p.mask != nullptr) {
              // fused activation gradient (gmgan_inference_cifar10.py:122-123 LeakyReLU / :179 ReLU, backward): dx *= act'(y)
              const float* mp_ = p.mask + (size_t)(cbeg + my_c) * 4;
              float4 mk[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                mk[u] = off[u] >= 0 ? *reinterpret_cast<const float4*>(mp_ + off[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                val[u].x = act_grad_from_out(mk[u].x, val[u].x, p.mask_act, p.mask_alpha);
                val[u].y = act_grad_from_out(mk[u].y, val[u].y, p.mask_act, p.mask_alpha);
                val[u].z = act_grad_from_out(mk[u].z, val[u].z, p.mask_act, p.mask_alpha);
                val[u].w = act_grad_from_out(mk[u].w, val[u].w, p.mask_act, p.mask_alpha);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (off[u] >= 0) *reinterpret_cast<float4*>(gp_ + off[u]) = val[u];
          }
        }
      }
      if (p.dbg && et == 0) p.dbg[203] = gtime();
    }
    if (MODE != 2 && p.stats != nullptr) {
      // ---- batch-norm statistics of the tile just written (fused producer of tf.nn.fused_batch_norm's moments) -----------
      // the staging tile still holds the final values (bias and activation applied).  Thread (my_c, my_r) sums its float4
      // column over rows my_r, my_r + rpp, ... (fixed order), the rpp partials of a column meet in shared memory (the staging
      // tile itself, once everybody has left it) and thread my_c writes [sum ; sum of squares] of this m-tile's rows to
      // stats[m-tile][0 / 1][channel]: every (m-tile, channel) is written by exactly one thread of the launch — no atomics,
      // bit-reproducible; rows outside the image / batch (s_row_off < 0) do not count.
      const int ncol = cend - cbeg;
      const int rpp = ncol > 0 ? kEpiThreads / ncol : 0;
      const int my_c = ncol > 0 ? et % ncol : 0, my_r = ncol > 0 ? et / ncol : 0;
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), q4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (my_r < rpp) {
        const uint32_t sp_ = stage_a + (uint32_t)my_c * 16u;
        for (int r = my_r; r < 128; r += rpp) {
          if (s_row_off[r] < 0) continue;
          const float4 v = lds128(sp_ + (uint32_t)(r * ld) * 4u);
          s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
          q4.x = fmaf(v.x, v.x, q4.x); q4.y = fmaf(v.y, v.y, q4.y); q4.z = fmaf(v.z, v.z, q4.z); q4.w = fmaf(v.w, v.w, q4.w);
        }
      }
      epi_bar();                                  // nobody reads the staging tile any more
      if (my_r < rpp) {
        sts128(stage_a + (uint32_t)et * 32u, s4);
        sts128(stage_a + (uint32_t)et * 32u + 16u, q4);
      }
      epi_bar();
      if (et < ncol) {                            // my_r == 0, my_c == et
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k2 = 0; k2 < rpp; ++k2) {
          const float4 a = lds128(stage_a + (uint32_t)(et + k2 * ncol) * 32u);
          const float4 b = lds128(stage_a + (uint32_t)(et + k2 * ncol) * 32u + 16u);
          S.x += a.x; S.y += a.y; S.z += a.z; S.w += a.w;
          Q.x += b.x; Q.y += b.y; Q.z += b.z; Q.w += b.w;
        }
        float* dst = p.stats + ((size_t)(tile / p.n_tiles) * 2) * (size_t)p.stats_ld + (size_t)n0 + (size_t)(cbeg + et) * 4;
        *reinterpret_cast<float4*>(dst) = S;
        *reinterpret_cast<float4*>(dst + p.stats_ld) = Q;
      }
    }
  }
  if (threadIdx.x == 64) GG_DBG(131);
#ifdef GG_TIMELINE
  if (p.dbg && threadIdx.x == 64) atomicMax(reinterpret_cast<unsigned long long*>(p.dbg + 200), (unsigned long long)gtime());
#endif
  // ---- teardown -----------------------------------------------------------------------------------------
  tc_fence_before();
  __syncwarp();
  // exit rendezvous: no CTA may exit (and free its shared memory) while a peer's copy still reads it.  Every thread arrived
  // above — the epilogue threads only after all slices addressed to their CTA had landed.
  if (cl) cluster_wait();
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
  if (s > kMaxSplits) s = kMaxSplits;        // the K splits of a tile form one thread-block cluster (portable size 8)
  if (s < 1) s = 1;
  return s;
}

long long* g_dbg = nullptr;
thread_local int g_last_info[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // gg_last_tc_info

struct TcPlan {
  bool ok;
  TcParams p;
  int grid_x;
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
  if (p.splits > p.n_tile / 4) p.splits = p.n_tile / 4;      // every split owns at least one float4 column of the tile
  pl.ok = true;
  return pl;
}

// split-K needs no global workspace any more (partials travel through distributed shared memory); the C-ABI keeps its
// workspace arguments for the direct / im2col paths
size_t plan_workspace(const TcPlan& pl) { return pl.ok ? 256 : 0; }

// shared-memory layout of a launch: operand ring (doubling as partial tile + staging tile in the epilogue) and, for
// split-K, the landing slots behind it
size_t smem_layout(TcParams& p) {
  const size_t stage_bytes = kABytes + (size_t)p.n_tile * 128;
  size_t ring = (size_t)p.stages * stage_bytes;
  const size_t tile = (size_t)128 * p.n_tile * 4;
  if (p.splits == 1) {
    const size_t staging = (size_t)128 * (p.n_tile + 4) * 4;
    if (ring < staging) ring = staging;
    p.stage_off = 0; p.land_off = 0; p.ncols_max = 0;
    return ring + 1024;
  }
  const int nc = p.n_tile / 4;
  p.ncols_max = (nc + p.splits - 1) / p.splits;
  size_t staging = (size_t)128 * (p.ncols_max * 4 + 4) * 4;
  if (staging < 8192) staging = 8192;             // the batch-norm statistics fold parks 256 x 32 B in the staging tile
  if (ring < tile + staging) ring = tile + staging;
  ring = (ring + 1023) & ~size_t(1023);
  p.stage_off = (int)tile;
  p.land_off = (int)ring;
  return ring + (size_t)(p.splits - 1) * p.ncols_max * 2048 + 1024;
}

// how many clusters of `splits` CTAs with `smem` bytes each can be resident at once (cached: one driver query per shape)
template <int MODE>
int max_active_clusters(int splits, size_t smem) {
  static std::map<std::pair<int, size_t>, int> cache;
  auto key = std::make_pair(splits, smem);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kNumSMs, splits);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)splits;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<MODE>, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[key] = n;
  return n;
}

int unsplit_stages(int stages, int n_tile) {
  if (g_tc_stage_cap < 2) return stages;
  int cap = g_tc_stage_cap;
  if (cap == 3) {
    const int fit = (113 * 1024 - 4096) / (kABytes + n_tile * 128);
    if (fit > cap) cap = fit;
  }
  return cap < stages ? cap : stages;
}

template <int MODE>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, TcPlan& pl, void* ws, size_t ws_bytes, cudaStream_t st) {
  TcParams& p = pl.p;
  (void)ws; (void)ws_bytes;
  p.dbg = g_dbg;
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[MODE]) {
    // opt-in limit is 227 KB per block INCLUDING the kernel's static shared memory (barriers, row table, bias: ~3 KB)
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    // largest shared-memory carveout: the driver's default choice follows the footprint of ONE block of the launch, which leaves
    // no room for a second 75-97 KB CTA on the SM
    cudaFuncSetAttribute(conv_tc_kernel<MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_set[MODE] = true;
  }
  p.stages = kMaxStages;
  if ((long long)pl.grid_x * p.splits > kNumSMs) {
    // more CTAs than SMs: keep the footprint under half an SM so two CTAs co-reside and one's epilogue overlaps the
    // other's main loop
    int fit = (int)((113 * 1024 - 4096) / (kABytes + p.n_tile * 128));
    p.stages = fit < 3 ? 3 : (fit > kMaxStages ? kMaxStages : fit);
  }
  // ring-depth cap (gg_set_tc_stages): two un-split CTAs — of the same or of two concurrent launches on different streams —
  // share an SM when each stays under ~113 KB.  Cap 3 (the multi-stream plans' setting) means "as deep as that allows":
  // 3 stages of 32 KB for 128-wide tiles, 4 of 24 KB for 64-wide, 5 of 20 KB for 32-wide ones.
  if (p.splits == 1) p.stages = unsplit_stages(p.stages, p.n_tile);
  size_t smem = smem_layout(p);
  // split-K CTAs never share an SM (ring + landing slots exceed half of it) and the main loop is bound by ring depth /
  // round-trip latency (timeline: 0.7 us from TMA issue to the freed slot), so the ring is as deep as fits beside the
  // landing slots
  while (p.splits > 1 && smem > (size_t)kMaxDynSmem && p.stages > 2) {
    --p.stages;
    smem = smem_layout(p);
  }
  if (smem > (size_t)kMaxDynSmem) return fail(GG_ERR_BAD_ARG, "conv_tc: shared-memory layout%s of %lld bytes exceeds the limit", "", (long long)smem);
  // the K splits of a tile are one cluster: shrink the split count until every tile's cluster is resident at once (a second
  // wave of clusters would serialise the launch)
  while (p.splits > 1 && max_active_clusters<MODE>(p.splits, smem) < pl.grid_x) {
    --p.splits;
    p.stages = kMaxStages;
    smem = smem_layout(p);
    while (p.splits > 1 && smem > (size_t)kMaxDynSmem && p.stages > 2) { --p.stages; smem = smem_layout(p); }
    if (p.splits == 1) { p.stages = unsplit_stages(p.stages, p.n_tile); smem = smem_layout(p); }
  }
  dim3 grid(pl.grid_x, p.splits);
  g_last_info[0] = MODE; g_last_info[1] = pl.grid_x; g_last_info[2] = p.splits; g_last_info[3] = p.n_tile;
  g_last_info[4] = p.stages; g_last_info[5] = p.splits > 1 ? 1 : 0; g_last_info[6] = (int)smem; g_last_info[7] = p.m_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p.splits > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = (unsigned)p.splits;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (g_null_launch) { launch_null(st); return check_launch("gg_conv2d(tcgen05, null launch)"); }
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<MODE>, tmA, tmB, p);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(GG_ERR_CUDA_BASE + (int)e, "conv_tc: launch failed: %s", cudaGetErrorString(e)); }
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

namespace {
struct PendingStats { float* stats; int ld; };
thread_local PendingStats g_pending_stats = {nullptr, 0};
}  // namespace
// the next conv_tc_fwd / conv_tc_dgrad of this thread also writes the per-m-tile batch-norm statistics of its output
void conv_tc_set_pending_stats(float* stats, int ld) { g_pending_stats = PendingStats{stats, ld}; }
// m-tiles (rows of the statistics buffer) of a fwd / dgrad launch; 0 when the shape does not run on the tensor-core path
int conv_tc_stats_tiles(int mode, int B, int H, int W, int Ci, int Co, int k, int stride, int pad_t, int pad_l, int Ho, int Wo) {
  if (mode != 0 && mode != 1) return 0;
  TcPlan pl = make_plan(mode, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
  if (!pl.ok) return 0;
  return pl.grid_x / pl.p.n_tiles;
}

int conv_tc_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Ci, int Co, int k,
                int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, void* ws, size_t ws_bytes,
                cudaStream_t st, bool* handled, int filt_rows) {
  *handled = false;
  const PendingStats pst = g_pending_stats;
  g_pending_stats = PendingStats{nullptr, 0};
  TcPlan pl = make_plan(0, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
  if (!pl.ok) return GG_OK;
  pl.p.stats = pst.stats; pl.p.stats_ld = pst.ld;
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

namespace {
struct PendingMask { const float* y; int act; float alpha; };
thread_local PendingMask g_pending_mask = {nullptr, 0, 0.f};
}  // namespace
// the next conv_tc_dgrad of this thread multiplies its result by act'(y) while storing it (gg_conv2d_dgrad_actgrad)
void conv_tc_set_pending_mask(const float* y, int act, float alpha) { g_pending_mask = PendingMask{y, act, alpha}; }

int conv_tc_dgrad(const float* dy, const float* w, const float* bias, float* dx, int B, int H, int W, int Ci, int Co, int k,
                  int stride, int pad_t, int pad_l, int Ho, int Wo, int act, float alpha, void* ws, size_t ws_bytes,
                  cudaStream_t st, bool* handled, int filt_rows) {
  *handled = false;
  const PendingMask mask = g_pending_mask;
  g_pending_mask = PendingMask{nullptr, 0, 0.f};
  const PendingStats pst = g_pending_stats;
  g_pending_stats = PendingStats{nullptr, 0};
  TcPlan pl = make_plan(1, B, H, W, Ci, Co, k, stride, pad_t, pad_l, Ho, Wo);
  if (!pl.ok) return GG_OK;
  pl.p.mask = mask.y; pl.p.mask_act = mask.act; pl.p.mask_alpha = mask.alpha;
  pl.p.stats = pst.stats; pl.p.stats_ld = pst.ld;
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
