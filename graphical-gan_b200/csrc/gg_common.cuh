// gg_common.cuh — shared helpers for the libgg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/gg_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgg_b200 is written for sm_100a (B200) only"
#endif

namespace gg {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;
extern thread_local int g_last_backend;
extern int g_conv_backend;

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return GG_ERR_CUDA_BASE + (int)e;
  }
  return GG_OK;
}

#define GG_REQUIRE(cond, msg)                                            \
  do {                                                                   \
    if (!(cond)) return gg::fail(GG_ERR_BAD_ARG, "%s: requirement failed: " #cond, msg); \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// Every kernel begins with GG_PDL_ENTRY(): it lets the NEXT kernel of the stream start launching (it will block at its own
// griddepcontrol.wait until this grid has completed and flushed) and then waits for the PREVIOUS kernel's memory.  Both
// instructions are no-ops for a launch without a programmatic dependency (measured: tools/exp/exp_pdl.cu, plain column).
// GG_LAUNCH adds the launch attribute when gg_set_pdl(1) / GG_PDL=1 is active; otherwise it is an ordinary <<<>>> launch.
// Measured on a B200 for a graph-captured chain: 1.12 -> 0.80 us per dependent launch (profiles/exp_pdl_r1.txt).
#define GG_PDL_ENTRY()                                                   \
  do {                                                                   \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      \
    asm volatile("griddepcontrol.wait;" ::: "memory");                   \
  } while (0)

extern int g_pdl;
// measurement aid (gg_set_null_launch / GG_NULL_LAUNCH=1): every kernel launch of the library is replaced by an EMPTY one-warp
// kernel on the same stream.  A captured training step then keeps its exact node / dependency structure but does no work: its
// duration is the launch + dependency-resolution cost of the graph alone (tools/exp_null_step.py).  Results are garbage.
extern int g_null_launch;
void launch_null(cudaStream_t st);

template <typename... KArgs, typename... Args>
inline void launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  if (g_null_launch) { launch_null(st); return; }
  if (!g_pdl) {
    kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
    return;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define GG_LAUNCH(kernel, grid, block, smem, stream, ...) gg::launch_ex(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)

// none / relu / leaky are all max(a*v, v) with a = 1 / 0 / alpha: one FMUL + FMNMX on the common path.  (A per-element
// `switch` over all activations made ptxas predicate the tanh/exp bodies into every element: measured 700 cycles per
// float4 in the tcgen05 epilogue.)  The transcendental activations sit behind a warp-uniform branch.
__device__ __forceinline__ float act_slope(int act, float alpha) {
  return act == GG_ACT_NONE ? 1.f : (act == GG_ACT_RELU ? 0.f : alpha);
}
static __device__ __noinline__ float apply_act_slow(float v, int act) {
  return act == GG_ACT_TANH ? tanhf(v) : 1.f / (1.f + expf(-v));
}
__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  if (act >= GG_ACT_TANH) return apply_act_slow(v, act);
  return fmaxf(v * act_slope(act, alpha), v);
}

// derivative of the activation expressed through its OUTPUT y (sign(y)==sign(x) for relu/leaky with alpha>0)
__device__ __forceinline__ float act_grad_from_out(float y, float g, int act, float alpha) {
  switch (act) {
    case GG_ACT_RELU: return y > 0.f ? g : 0.f;
    case GG_ACT_LEAKY: return y > 0.f ? g : alpha * g;
    case GG_ACT_TANH: return (1.f - y * y) * g;
    case GG_ACT_SIGMOID: return y * (1.f - y) * g;
    default: return g;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024; result valid in all threads
__device__ __forceinline__ float block_sum(float v, float* sh /* >= 32 floats */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

}  // namespace gg
