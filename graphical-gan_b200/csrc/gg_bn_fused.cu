// gg_bn_fused.cu — training-mode batch normalisation (batch statistics, tflib/ops/batchnorm.py:30,77-84) and its gradient
// as ONE kernel each.
//
// The two-/three-kernel form (gg_bn_stats -> gg_bn_apply; gg_bn_bwd_reduce -> gg_bn_fold_partials -> gg_bn_bwd_apply) sits on
// the critical path of the training step five times forward and five times backward; every extra launch costs a
// kernel-to-kernel dependency latency (~2 us) on tensors that are only 1-4 MB.  Per-channel statistics are independent
// across channels, so the work is partitioned by CHANNEL GROUP: a thread-block cluster owns `cw` consecutive channels
// (32-byte row segments for cw = 8), its CTAs split the rows, the per-CTA partial sums meet in distributed shared memory
// (one cluster barrier), and every CTA then normalises its own row slice.  x is read twice (statistics, apply) — the
// second read is an L1/L2 hit for these sizes — and y written once: 12*R*C bytes algorithmic, HBM/L2-bound.
// No grid-wide synchronisation, no atomics, deterministic summation order.
//
// Data parallel (SyncBN, gg_bn_*_fused_dp): batch statistics couple the ranks, so the per-channel sums must be totalled over
// all GPUs before the normalisation.  Round 1 did that with four launches per layer (stats, fold, a one-CTA all-reduce
// kernel, apply) — 15 cross-GPU rendezvous per iteration, each with the dependent-launch latencies around it, ~0.35 ms of
// a 1.0 ms step at N=2.  Here the exchange lives INSIDE the one-launch kernel: the rank-0 CTA of every channel-group
// cluster writes its group's local sums into every peer's exchange arena over NVLink (plain stores to CUDA-IPC mapped
// memory), publishes an epoch flag per (group, source rank) with a system-scope release, acquire-spins on the flags the
// peers wrote into ITS arena, and totals the ranks' sums in rank order (identical result on every GPU); the totals reach
// the cluster's other CTAs through distributed shared memory.  Per-site flags and parity-double-buffered data slots; the
// epoch counter of a (site, group) has a single writer.  Spins are bounded (a lost peer traps instead of hanging).
#include "gg_common.cuh"

using namespace gg;

namespace {

constexpr int kThreads = 512;
constexpr int kMaxCw = 32;

constexpr int kMaxPeers = 16;
// data-parallel context of one batch-norm call site (world == 1: no exchange)
struct DpCtx {
  void* peer[kMaxPeers];     // every rank's exchange arena (own included), CUDA-IPC mapped
  long long site_off;        // byte offset of this call site's region inside every arena
  int rank, world;
  int C;                     // channels (layout of the region)
};
// region layout: [groups_max] epoch counters | [2 parities][world][C][2] 8-byte words {fp32 payload, epoch}
__host__ __device__ inline size_t dp_groups_max(int C) { return (size_t)(C + 3) / 4; }
__host__ __device__ inline size_t dp_flags_off(int C) { return ((dp_groups_max(C) * 4 + 127) / 128) * 128; }
__host__ __device__ inline size_t dp_data_off(int C, int world) { (void)world; return dp_flags_off(C); }
__host__ __device__ inline size_t dp_site_bytes(int C, int world) { return dp_data_off(C, world) + (size_t)2 * world * C * 16; }

struct BnPlan {
  bool ok;
  int cw;       // channels per cluster
  int groups;   // channel groups (clusters)
  int cs;       // CTAs per cluster (row split)
};

BnPlan bn_plan(int R, int C) {
  BnPlan p{false, 0, 0, 1};
  if (R <= 0 || C <= 0 || C % 4 != 0) return p;
  p.cw = (C >= 32 && R <= 256) ? 32 : 8;
  if (p.cw > C) p.cw = (C >= 8) ? 8 : 4;
  p.groups = ceil_div(C, p.cw);
  p.cs = 1;
  while (p.cs < 8 && p.groups * p.cs < 128 && R / (p.cs * 2) >= 128) p.cs *= 2;
  p.ok = true;
  return p;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double* local_ptr, uint32_t cta_rank) {
  uint32_t local = (uint32_t)__cvta_generic_to_shared(local_ptr), remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta_rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
  return v;
}

// Reduce per-thread (a[4], b[4]) over all threads of the CTA that share a channel quad, then over the CTAs of the
// cluster.  Thread t owns quad t % Q; on return threads 0..cw-1 hold the cluster totals of channel c0 + t in (ta, tb).
__device__ __forceinline__ double2 ld_dsmem_f64x2(const double* local_ptr, uint32_t cta_rank) {
  return make_double2(ld_dsmem_f64(local_ptr, cta_rank), ld_dsmem_f64(local_ptr + 1, cta_rank));
}

// Cross-GPU total of the cluster totals held by threads 0..cw-1 of the cluster's rank-0 CTA.
// Low-latency protocol (the idea of NCCL's LL): every 8-byte word carries 4 bytes of payload and the 4-byte epoch, and an
// aligned 8-byte store is delivered whole over NVLink — so the data IS the flag: no system-scope fence, no separate flag
// round trip (the fence + flag form measured +11.6 us per launch on 2 B200s, profiles/dp_bn_r2.txt; a kernel is only
// 3-10 us long).  Thread t pushes {sum_t, e} and {sumsq_t, e} of its channel into slot [parity][my rank] of EVERY rank's
// arena (fire and forget) and then polls the P x 2 words of its channel in its OWN arena until they carry epoch e.  The
// sums travel as fp32 (the per-rank partials are exact to 6e-8 relative; they are totalled in double).  A slot is
// rewritten two epochs later, which a peer can only reach after this rank has consumed the current one.
__device__ __forceinline__ void dp_exchange(const DpCtx& dp, int group, int c0, int cw, double& ta, double& tb) {
  const int tid = threadIdx.x;            // called by warp 0 only (cw <= 32)
  const int P = dp.world;
  uint8_t* mine = reinterpret_cast<uint8_t*>(dp.peer[dp.rank]) + dp.site_off;
  unsigned e = 0;
  if (tid == 0) {
    unsigned* ep = reinterpret_cast<unsigned*>(mine) + group;
    e = ep[0] + 1u;
    ep[0] = e;
  }
  e = __shfl_sync(0xffffffffu, e, 0);
  const size_t par = (size_t)(e & 1u);
  const int ch = c0 + tid;
  if (tid < cw && ch < dp.C) {
    const uint2 wa = make_uint2(__float_as_uint((float)ta), e), wb = make_uint2(__float_as_uint((float)tb), e);
    for (int r = 0; r < P; ++r) {
      uint2* d = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(dp.peer[r]) + dp.site_off + dp_data_off(dp.C, P)) +
                 ((par * P + dp.rank) * dp.C + ch) * 2;
      asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(d), "r"(wa.x), "r"(wa.y) : "memory");
      asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(d + 1), "r"(wb.x), "r"(wb.y) : "memory");
    }
    const uint2* src = reinterpret_cast<const uint2*>(mine + dp_data_off(dp.C, P)) + (par * P * dp.C + ch) * 2;
    ta = 0.0; tb = 0.0;
    for (int r = 0; r < P; ++r) {          // rank order: identical totals on every GPU
      const uint2* w = src + (size_t)r * dp.C * 2;
      uint2 va, vb;
      unsigned spins = 0;
      do {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(va.x), "=r"(va.y) : "l"(w) : "memory");
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(vb.x), "=r"(vb.y) : "l"(w + 1) : "memory");
        if ((va.y != e || vb.y != e) && ++spins > (1u << 27)) __trap();     // a lost peer traps instead of hanging the GPU
      } while (va.y != e || vb.y != e);
      ta += (double)__uint_as_float(va.x);
      tb += (double)__uint_as_float(vb.x);
    }
  }
  __syncwarp();
}

template <int Q>
__device__ __forceinline__ void reduce_channels(float (&a)[4], float (&b)[4], int cs, double& ta, double& tb,
                                                float (*wred)[kMaxCw][2], double (*xch)[2], const DpCtx& dp, int group, int c0,
                                                double* la, double* lb) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 16; o >= Q; o >>= 1) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      a[e] += __shfl_xor_sync(0xffffffffu, a[e], o);
      b[e] += __shfl_xor_sync(0xffffffffu, b[e], o);
    }
  }
  if (lane < Q) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { wred[warp][lane * 4 + e][0] = a[e]; wred[warp][lane * 4 + e][1] = b[e]; }
  }
  __syncthreads();
  const int cw = Q * 4;
  ta = 0.0; tb = 0.0;
  if (tid < cw) {
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) { ta += (double)wred[w][tid][0]; tb += (double)wred[w][tid][1]; }
    xch[tid][0] = ta;
    xch[tid][1] = tb;
  }
  if (cs > 1) {
    cluster_sync_all();                       // every CTA's xch is written and visible cluster-wide
    if (tid < cw) {
      ta = 0.0; tb = 0.0;
      for (int r = 0; r < cs; ++r) {          // rank order: identical, deterministic totals on every CTA
        ta += ld_dsmem_f64(&xch[tid][0], (uint32_t)r);
        tb += ld_dsmem_f64(&xch[tid][1], (uint32_t)r);
      }
    }
    cluster_sync_all();                       // peers have finished reading this CTA's xch: it may exit / reuse
  }
  *la = ta; *lb = tb;                         // local (this GPU's) totals: the parameter-gradient sums of the backward kernel
  if (dp.world > 1) {
    const bool lead = cs == 1 || cluster_ctarank() == 0;
    if (lead && tid < 32) {
      dp_exchange(dp, group, c0, cw, ta, tb);
      if (tid < cw) { xch[tid][0] = ta; xch[tid][1] = tb; }
    }
    if (cs > 1) {
      cluster_sync_all();                     // rank 0's xch now holds the cross-GPU totals
      if (!lead && tid < cw) {
        const double2 v = ld_dsmem_f64x2(&xch[tid][0], 0u);
        ta = v.x; tb = v.y;
      }
      cluster_sync_all();
    }
  }
}

// NI > 0: the rows of a thread (at most NI float4) stay in REGISTERS between the statistics pass and the normalise pass — x is
// read once, all NI loads are in flight together (the streaming form below exposes an L2 round trip per 4 rows and reads x
// twice: 9.3 us for [16384 x 64], profiles/time_bn_r2.txt).  NI = 0: streaming two-pass form for any R.
template <int Q, int NI>
__global__ void __launch_bounds__(kThreads) bn_fwd_fused_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps,
                                                                float* __restrict__ y, float* __restrict__ mean_out,
                                                                float* __restrict__ rstd_out, int R, int C, int cs, int act,
                                                                float alpha, const DpCtx dp) {
  GG_PDL_ENTRY();
  constexpr int cw = Q * 4, RL = kThreads / Q;
  __shared__ float wred[kThreads / 32][kMaxCw][2];
  __shared__ double xch[kMaxCw][2];
  __shared__ float s_scale[kMaxCw], s_shift[kMaxCw];
  const int tid = threadIdx.x;
  const int rank = cs > 1 ? (int)cluster_ctarank() : 0;
  const int group = blockIdx.x / cs;
  const int c0 = group * cw;
  const int quad = tid % Q, rl = tid / Q;
  const int c = c0 + quad * 4;
  const bool c_ok = c < C;                                   // C % 4 == 0: a quad is entirely inside or outside
  const int rows_per = (R + cs - 1) / cs;
  const int r_begin = rank * rows_per, r_end = min(R, r_begin + rows_per);

  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  float4 cache[NI > 0 ? NI : 1];
  if (c_ok) {
    if (NI > 0) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int r = r_begin + rl + i * RL;
        cache[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < r_end) cache[i] = *reinterpret_cast<const float4*>(x + (size_t)r * C + c);
      }
#pragma unroll
      for (int i = 0; i < NI; ++i) {                         // rows past r_end are zeros: they add nothing
        const float4 v = cache[i];
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
        b[0] += v.x * v.x; b[1] += v.y * v.y; b[2] += v.z * v.z; b[3] += v.w * v.w;
      }
    } else {
#pragma unroll 4
      for (int r = r_begin + rl; r < r_end; r += RL) {
        const float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * C + c);
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
        b[0] += v.x * v.x; b[1] += v.y * v.y; b[2] += v.z * v.z; b[3] += v.w * v.w;
      }
    }
  }
  double ta, tb, la, lb;
  reduce_channels<Q>(a, b, cs, ta, tb, wred, xch, dp, group, c0, &la, &lb);
  const double Rg = (double)R * (double)dp.world;            // equal shards: the global row count
  if (tid < cw) {
    const int ch = c0 + tid;
    if (ch < C) {
      const double mean = ta / Rg;
      double var = tb / Rg - mean * mean;                    // biased batch variance (fused_batch_norm, is_training)
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      const float g = gamma ? gamma[ch] : 1.f, bt = beta ? beta[ch] : 0.f;
      s_scale[tid] = rstd * g;
      s_shift[tid] = bt - (float)mean * rstd * g;
      if (rank == 0) {
        if (mean_out) mean_out[ch] = (float)mean;
        if (rstd_out) rstd_out[ch] = rstd;
      }
    }
  }
  __syncthreads();
  if (!c_ok) return;
  const float4 sc = *reinterpret_cast<const float4*>(&s_scale[quad * 4]);
  const float4 sh = *reinterpret_cast<const float4*>(&s_shift[quad * 4]);
  if (NI > 0) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int r = r_begin + rl + i * RL;
      if (r < r_end) {
        const float4 v = cache[i];
        float4 o;
        o.x = apply_act(v.x * sc.x + sh.x, act, alpha);
        o.y = apply_act(v.y * sc.y + sh.y, act, alpha);
        o.z = apply_act(v.z * sc.z + sh.z, act, alpha);
        o.w = apply_act(v.w * sc.w + sh.w, act, alpha);
        *reinterpret_cast<float4*>(y + (size_t)r * C + c) = o;
      }
    }
  } else {
#pragma unroll 4
    for (int r = r_begin + rl; r < r_end; r += RL) {
      const size_t i = (size_t)r * C + c;
      const float4 v = *reinterpret_cast<const float4*>(x + i);
      float4 o;
      o.x = apply_act(v.x * sc.x + sh.x, act, alpha);
      o.y = apply_act(v.y * sc.y + sh.y, act, alpha);
      o.z = apply_act(v.z * sc.z + sh.z, act, alpha);
      o.w = apply_act(v.w * sc.w + sh.w, act, alpha);
      *reinterpret_cast<float4*>(y + i) = o;
    }
  }
}

template <int Q, int NI>
__global__ void __launch_bounds__(kThreads) bn_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                const float* __restrict__ y, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                float* __restrict__ dx, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, int R, int C, int cs, int act,
                                                                float alpha, const DpCtx dp) {
  GG_PDL_ENTRY();
  constexpr int cw = Q * 4, RL = kThreads / Q;
  __shared__ float wred[kThreads / 32][kMaxCw][2];
  __shared__ double xch[kMaxCw][2];
  __shared__ float s_mg[kMaxCw], s_mgx[kMaxCw];
  const int tid = threadIdx.x;
  const int rank = cs > 1 ? (int)cluster_ctarank() : 0;
  const int group = blockIdx.x / cs;
  const int c0 = group * cw;
  const int quad = tid % Q, rl = tid / Q;
  const int c = c0 + quad * 4;
  const bool c_ok = c < C;
  const int rows_per = (R + cs - 1) / cs;
  const int r_begin = rank * rows_per, r_end = min(R, r_begin + rows_per);

  float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), rs4 = m4;
  if (c_ok) {
    m4 = *reinterpret_cast<const float4*>(mean + c);
    rs4 = *reinterpret_cast<const float4*>(rstd + c);
  }
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  float4 cg[NI > 0 ? NI : 1], cxh[NI > 0 ? NI : 1];         // NI > 0: activation-masked gradient and x-hat kept in registers
  if (c_ok && NI > 0) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int r = r_begin + rl + i * RL;
      cg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      cxh[i] = m4;                                           // x := mean, i.e. x-hat = 0, for rows past r_end
      if (r < r_end) {
        const size_t idx = (size_t)r * C + c;
        cg[i] = *reinterpret_cast<const float4*>(dy + idx);
        cxh[i] = *reinterpret_cast<const float4*>(x + idx);
      }
    }
    if (act != GG_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int r = r_begin + rl + i * RL;
        if (r < r_end) {
          const float4 yv = *reinterpret_cast<const float4*>(y + (size_t)r * C + c);
          float4 g = cg[i];
          g.x = act_grad_from_out(yv.x, g.x, act, alpha); g.y = act_grad_from_out(yv.y, g.y, act, alpha);
          g.z = act_grad_from_out(yv.z, g.z, act, alpha); g.w = act_grad_from_out(yv.w, g.w, act, alpha);
          cg[i] = g;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const float4 g = cg[i];
      float4 xh = cxh[i];
      xh.x = (xh.x - m4.x) * rs4.x; xh.y = (xh.y - m4.y) * rs4.y; xh.z = (xh.z - m4.z) * rs4.z; xh.w = (xh.w - m4.w) * rs4.w;
      cxh[i] = xh;
      a[0] += g.x; a[1] += g.y; a[2] += g.z; a[3] += g.w;
      b[0] += g.x * xh.x; b[1] += g.y * xh.y; b[2] += g.z * xh.z; b[3] += g.w * xh.w;
    }
  }
  if (c_ok && NI == 0) {
#pragma unroll 2
    for (int r = r_begin + rl; r < r_end; r += RL) {
      const size_t i = (size_t)r * C + c;
      float4 g = *reinterpret_cast<const float4*>(dy + i);
      const float4 xv = *reinterpret_cast<const float4*>(x + i);
      if (act != GG_ACT_NONE) {
        const float4 yv = *reinterpret_cast<const float4*>(y + i);
        g.x = act_grad_from_out(yv.x, g.x, act, alpha); g.y = act_grad_from_out(yv.y, g.y, act, alpha);
        g.z = act_grad_from_out(yv.z, g.z, act, alpha); g.w = act_grad_from_out(yv.w, g.w, act, alpha);
      }
      a[0] += g.x; a[1] += g.y; a[2] += g.z; a[3] += g.w;
      b[0] += g.x * ((xv.x - m4.x) * rs4.x); b[1] += g.y * ((xv.y - m4.y) * rs4.y);
      b[2] += g.z * ((xv.z - m4.z) * rs4.z); b[3] += g.w * ((xv.w - m4.w) * rs4.w);
    }
  }
  double ta, tb, la, lb;
  reduce_channels<Q>(a, b, cs, ta, tb, wred, xch, dp, group, c0, &la, &lb);
  const double Rg = (double)R * (double)dp.world;
  if (tid < cw) {
    const int ch = c0 + tid;
    if (ch < C) {
      s_mg[tid] = (float)(ta / Rg);
      s_mgx[tid] = (float)(tb / Rg);
      if (rank == 0) {                                       // LOCAL sums: the gradient all-reduce totals them over the ranks
        if (dbeta) dbeta[ch] = (float)la;
        if (dgamma) dgamma[ch] = (float)lb;
      }
    }
  }
  __syncthreads();
  if (!c_ok || dx == nullptr) return;
  const float4 mg = *reinterpret_cast<const float4*>(&s_mg[quad * 4]);
  const float4 mgx = *reinterpret_cast<const float4*>(&s_mgx[quad * 4]);
  float4 k4 = rs4;
  if (gamma) {
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
    k4.x *= gm.x; k4.y *= gm.y; k4.z *= gm.z; k4.w *= gm.w;
  }
  if (NI > 0) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int r = r_begin + rl + i * RL;
      if (r < r_end) {
        const float4 g = cg[i], xh = cxh[i];
        float4 o;
        o.x = k4.x * (g.x - mg.x - xh.x * mgx.x);
        o.y = k4.y * (g.y - mg.y - xh.y * mgx.y);
        o.z = k4.z * (g.z - mg.z - xh.z * mgx.z);
        o.w = k4.w * (g.w - mg.w - xh.w * mgx.w);
        *reinterpret_cast<float4*>(dx + (size_t)r * C + c) = o;
      }
    }
  } else {
#pragma unroll 2
    for (int r = r_begin + rl; r < r_end; r += RL) {
      const size_t i = (size_t)r * C + c;
      float4 g = *reinterpret_cast<const float4*>(dy + i);
      const float4 xv = *reinterpret_cast<const float4*>(x + i);
      if (act != GG_ACT_NONE) {
        const float4 yv = *reinterpret_cast<const float4*>(y + i);
        g.x = act_grad_from_out(yv.x, g.x, act, alpha); g.y = act_grad_from_out(yv.y, g.y, act, alpha);
        g.z = act_grad_from_out(yv.z, g.z, act, alpha); g.w = act_grad_from_out(yv.w, g.w, act, alpha);
      }
      float4 o;
      o.x = k4.x * (g.x - mg.x - (xv.x - m4.x) * rs4.x * mgx.x);
      o.y = k4.y * (g.y - mg.y - (xv.y - m4.y) * rs4.y * mgx.y);
      o.z = k4.z * (g.z - mg.z - (xv.z - m4.z) * rs4.z * mgx.z);
      o.w = k4.w * (g.w - mg.w - (xv.w - m4.w) * rs4.w * mgx.w);
      *reinterpret_cast<float4*>(dx + i) = o;
    }
  }
}

constexpr int kBnNI = 8;    // rows per thread the register-cached kernels hold

bool bn_cached_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GG_BN_CACHED"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

template <typename Kern, typename... Args>
int launch_clustered(Kern kern, const BnPlan& pl, cudaStream_t st, const char* what, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pl.groups * pl.cs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pl.cs > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)pl.cs;
    attr[na].val.clusterDim.y = 1;
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
  if (g_null_launch) { launch_null(st); return check_launch(what); }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(GG_ERR_CUDA_BASE + (int)e, "%s: launch failed", what);
  }
  return check_launch(what);
}

}  // namespace

extern "C" int gg_bn_fused_supported(int R, int C) { return bn_plan(R, C).ok ? 1 : 0; }

static int bn_fwd_launch(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean_out,
                         float* rstd_out, int R, int C, int act, float alpha, const DpCtx& dp, void* stream, const char* what) {
  if (R <= 0 || C <= 0) return GG_OK;
  const BnPlan pl = bn_plan(R, C);
  if (!pl.ok) return fail(GG_ERR_UNSUPPORTED, "%s: C must be a multiple of 4", what);
  cudaStream_t st = as_stream(stream);
  const int q = pl.cw / 4;
  const int rows_per = (R + pl.cs - 1) / pl.cs, rl = kThreads / q;
  const bool cached = bn_cached_enabled() && rows_per <= kBnNI * rl && rows_per >= 4 * rl;   // measured: wins from 4 rows per thread
#define GG_BN_FWD(Q_, NI_) \
  return launch_clustered(bn_fwd_fused_kernel<Q_, NI_>, pl, st, what, x, gamma, beta, eps, y, mean_out, rstd_out, R, C, pl.cs, act, alpha, dp)
  if (pl.cw == 32) { if (cached) GG_BN_FWD(8, kBnNI); GG_BN_FWD(8, 0); }
  if (pl.cw == 8) { if (cached) GG_BN_FWD(2, kBnNI); GG_BN_FWD(2, 0); }
  if (cached) GG_BN_FWD(1, kBnNI);
  GG_BN_FWD(1, 0);
#undef GG_BN_FWD
}

static int bn_bwd_launch(const float* dy, const float* x, const float* y, const float* mean, const float* rstd, const float* gamma,
                         float* dx, float* dgamma, float* dbeta, int R, int C, int act, float alpha, const DpCtx& dp, void* stream,
                         const char* what) {
  if (R <= 0 || C <= 0) return GG_OK;
  if (act != GG_ACT_NONE && y == nullptr) return fail(GG_ERR_BAD_ARG, "%s: y required when act is fused", what);
  const BnPlan pl = bn_plan(R, C);
  if (!pl.ok) return fail(GG_ERR_UNSUPPORTED, "%s: C must be a multiple of 4", what);
  cudaStream_t st = as_stream(stream);
  const int q = pl.cw / 4;
  const int rows_per = (R + pl.cs - 1) / pl.cs, rl = kThreads / q;
  const bool cached = bn_cached_enabled() && dx != nullptr && rows_per <= kBnNI * rl && rows_per >= 4 * rl;
#define GG_BN_BWD(Q_, NI_) \
  return launch_clustered(bn_bwd_fused_kernel<Q_, NI_>, pl, st, what, dy, x, y, mean, rstd, gamma, dx, dgamma, dbeta, R, C, pl.cs, act, alpha, dp)
  if (pl.cw == 32) { if (cached) GG_BN_BWD(8, kBnNI); GG_BN_BWD(8, 0); }
  if (pl.cw == 8) { if (cached) GG_BN_BWD(2, kBnNI); GG_BN_BWD(2, 0); }
  if (cached) GG_BN_BWD(1, kBnNI);
  GG_BN_BWD(1, 0);
#undef GG_BN_BWD
}

static int make_dp(DpCtx* dp, void* const* peer_arenas_host, int rank, int world, long long site_offset, int C, const char* what) {
  *dp = DpCtx{};
  dp->world = 1;
  dp->C = C;
  if (world <= 1) return GG_OK;
  if (world > kMaxPeers || rank < 0 || rank >= world || peer_arenas_host == nullptr || site_offset < 0 || (site_offset & 127))
    return fail(GG_ERR_BAD_ARG, "%s: bad world / rank / arena / site offset", what);
  for (int r = 0; r < world; ++r) dp->peer[r] = peer_arenas_host[r];
  dp->rank = rank; dp->world = world; dp->site_off = site_offset;
  return GG_OK;
}

extern "C" int gg_bn_fwd_fused(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean_out,
                               float* rstd_out, int R, int C, int act, float alpha, void* stream) {
  DpCtx dp{};
  dp.world = 1; dp.C = C;
  return bn_fwd_launch(x, gamma, beta, eps, y, mean_out, rstd_out, R, C, act, alpha, dp, stream, "gg_bn_fwd_fused");
}

extern "C" int gg_bn_bwd_fused(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                               const float* gamma, float* dx, float* dgamma, float* dbeta, int R, int C, int act, float alpha,
                               void* stream) {
  DpCtx dp{};
  dp.world = 1; dp.C = C;
  return bn_bwd_launch(dy, x, y, mean, rstd, gamma, dx, dgamma, dbeta, R, C, act, alpha, dp, stream, "gg_bn_bwd_fused");
}

extern "C" int gg_bn_fused_grid(int R, int C) {
  const BnPlan pl = bn_plan(R, C);
  return pl.ok ? pl.groups * pl.cs : 0;
}

extern "C" size_t gg_bn_dp_site_bytes(int C, int world) { return (dp_site_bytes(C, world < 1 ? 1 : world) + 127) & ~size_t(127); }

extern "C" int gg_bn_fwd_fused_dp(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean_out,
                                  float* rstd_out, int R, int C, int act, float alpha, void* const* peer_arenas_host, int rank,
                                  int world, long long site_offset, void* stream) {
  DpCtx dp;
  int rc = make_dp(&dp, peer_arenas_host, rank, world, site_offset, C, "gg_bn_fwd_fused_dp");
  if (rc) return rc;
  return bn_fwd_launch(x, gamma, beta, eps, y, mean_out, rstd_out, R, C, act, alpha, dp, stream, "gg_bn_fwd_fused_dp");
}

extern "C" int gg_bn_bwd_fused_dp(const float* dy, const float* x, const float* y, const float* mean, const float* rstd,
                                  const float* gamma, float* dx, float* dgamma, float* dbeta, int R, int C, int act, float alpha,
                                  void* const* peer_arenas_host, int rank, int world, long long site_offset, void* stream) {
  DpCtx dp;
  int rc = make_dp(&dp, peer_arenas_host, rank, world, site_offset, C, "gg_bn_bwd_fused_dp");
  if (rc) return rc;
  return bn_bwd_launch(dy, x, y, mean, rstd, gamma, dx, dgamma, dbeta, R, C, act, alpha, dp, stream, "gg_bn_bwd_fused_dp");
}
