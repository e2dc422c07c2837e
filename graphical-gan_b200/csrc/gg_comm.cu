// gg_comm.cu — one-shot all-reduce for SMALL buffers over NVLink peer memory (SyncBN statistics: 2*C floats, C <= 4096).
//
// NCCL's latency for a 1-32 KB message (~25 us) plus the need to keep it outside CUDA-graph capture made the ten
// batch-norm statistic exchanges of a training step cost more than the convolutions.  Every rank owns an exchange
// buffer (cudaMalloc'ed, exported with CUDA IPC, mapped by all peers through NVLink/NVSwitch P2P).  One CTA per rank:
//   1. bump the local epoch (device-resident counter: CUDA-graph replays advance it), copy the payload into the local
//      buffer's slot (epoch & 1), publish flag[slot] = epoch with a system-scope release;
//   2. acquire-spin on every peer's flag[slot] >= epoch (bounded: a lost peer traps instead of hanging the GPU);
//   3. sum all ranks' slots straight out of peer memory in rank order (identical, deterministic result on every rank).
// Two slots suffice: a rank can only reach epoch e+2 after every peer has started e+1, i.e. finished reading slot e.
// The gradient bucket (12-16 MB) stays on NCCL — bandwidth-bound messages are what NCCL is for.
#include "gg_common.cuh"

using namespace gg;

namespace {
struct CommHeader {          // lives at the start of every rank's exchange buffer
  unsigned flag[2];
  unsigned pad[30];          // payload slots start 128-byte aligned
};
constexpr int kMaxPeers = 16;
struct PeerTable { void* buf[kMaxPeers]; };

__global__ void __launch_bounds__(512) allreduce_small_kernel(const float* __restrict__ src, float* __restrict__ dst, int n,
                                                              PeerTable peers, int rank, int P, int slot_floats,
                                                              unsigned* __restrict__ epoch_counter) {
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = ++epoch_counter[0];
  __syncthreads();
  const unsigned e = s_epoch;
  const int slot = (int)(e & 1u);
  CommHeader* mine = reinterpret_cast<CommHeader*>(peers.buf[rank]);
  float* my_slot = reinterpret_cast<float*>(mine + 1) + (size_t)slot * slot_floats;
  for (int i = threadIdx.x; i < n; i += blockDim.x) my_slot[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&mine->flag[slot]), "r"(e) : "memory");
  if (threadIdx.x < P) {
    const unsigned* pf = &reinterpret_cast<const CommHeader*>(peers.buf[threadIdx.x])->flag[slot];
    unsigned seen, spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(pf) : "memory");
      if ((int)(seen - e) < 0 && ++spins > (1u << 27)) __trap();
    } while ((int)(seen - e) < 0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < P; ++r) {
      const float* ps = reinterpret_cast<const float*>(reinterpret_cast<const CommHeader*>(peers.buf[r]) + 1) + (size_t)slot * slot_floats;
      acc += __ldcv(ps + i);          // volatile load: never served from a stale L1 line
    }
    dst[i] = acc;
  }
}
}  // namespace

extern "C" size_t gg_comm_buffer_bytes(int max_floats) { return sizeof(CommHeader) + (size_t)2 * max_floats * sizeof(float); }

extern "C" int gg_comm_alloc(int max_floats, void** buf_out, void* ipc_handle_out_64) {
  void* p = nullptr;
  size_t bytes = gg_comm_buffer_bytes(max_floats);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "gg_comm_alloc: cudaMalloc failed: %s", cudaGetErrorString(e));
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "gg_comm_alloc: memset failed: %s", cudaGetErrorString(e));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "gg_comm_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(ipc_handle_out_64, &h, 64);
  *buf_out = p;
  return GG_OK;
}

// general-purpose exchange arena (SyncBN call sites of the one-launch batch-norm kernels, gg_bn_*_fused_dp): zero-filled,
// exported with CUDA IPC like the small all-reduce buffer
extern "C" int gg_comm_alloc_bytes(size_t bytes, void** buf_out, void* ipc_handle_out_64) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "gg_comm_alloc_bytes: cudaMalloc failed: %s", cudaGetErrorString(e));
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "gg_comm_alloc_bytes: memset failed: %s", cudaGetErrorString(e));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "gg_comm_alloc_bytes: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  memcpy(ipc_handle_out_64, &h, 64);
  *buf_out = p;
  return GG_OK;
}

extern "C" int gg_comm_open(const void* ipc_handle_64, void** buf_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)e, "gg_comm_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
  *buf_out = p;
  return GG_OK;
}

extern "C" int gg_allreduce_small(const float* src, float* dst, int n, void* const* peer_bufs_host, int rank, int world,
                                  int max_floats, void* epoch_counter, void* stream) {
  if (n <= 0) return GG_OK;
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n > max_floats)
    return fail(GG_ERR_BAD_ARG, "gg_allreduce_small: bad world / rank / size%s");
  PeerTable t;
  for (int r = 0; r < kMaxPeers; ++r) t.buf[r] = r < world ? peer_bufs_host[r] : nullptr;
  allreduce_small_kernel<<<1, 512, 0, as_stream(stream)>>>(src, dst, n, t, rank, world, max_floats,
                                                           reinterpret_cast<unsigned*>(epoch_counter));
  return check_launch("gg_allreduce_small");
}
