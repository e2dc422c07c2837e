// exp_pdl.cu — standalone micro-benchmark (not part of libgg_b200): what does programmatic dependent launch (PDL) save per
// dependent kernel inside a CUDA graph on this B200?
//
// The training step is a dependency chain of ~100 launches (profiles/timeline_gen_r1.txt); each pays the gap between the
// producer's last store and the consumer's first load.  With PDL the consumer is launched early (attribute
// cudaLaunchAttributeProgrammaticStreamSerialization), runs its prologue — here a calibrated busy loop standing in for
// mbarrier init / TMEM allocation / tensor-map prefetch — and blocks at `griddepcontrol.wait` until the producer's memory is
// visible; the producer releases its dependents with `griddepcontrol.launch_dependents`.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/exp_pdl tools/exp/exp_pdl.cu && /tmp/exp_pdl
//
// Prints us per chain step for: plain launches, PDL with trigger at kernel start, PDL with trigger at kernel end, each with
// a 0 / 500 / 1500 ns prologue and with 1 CTA or 148 CTAs per launch, all replayed from a captured graph.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void step_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int prologue_ns, int trigger_early) {
  if (trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // prologue: independent of the producer's data
  if (prologue_ns > 0) {
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    long long t = t0;
    while (t - t0 < prologue_ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");     // no-op when the launch carries no programmatic dependency
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] + 1.0f;
  if (!trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

static float run_chain(int steps, int ctas, int prologue_ns, int pdl, int trigger_early) {
  const int n = ctas * 128;
  float *a, *b;
  CK(cudaMalloc(&a, n * sizeof(float)));
  CK(cudaMalloc(&b, n * sizeof(float)));
  CK(cudaMemset(a, 0, n * sizeof(float)));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int s = 0; s < steps; ++s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(128);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && s > 0) ? 1 : 0;
    const float* in = (s & 1) ? b : a;
    float* out = (s & 1) ? a : b;
    CK(cudaLaunchKernelEx(&cfg, step_kernel, in, out, n, prologue_ns, trigger_early));
  }
  CK(cudaStreamEndCapture(st, &graph));
  CK(cudaGraphInstantiate(&exec, graph, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 12; ++rep) {
    CK(cudaEventRecord(e0, st));
    CK(cudaGraphLaunch(exec, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep >= 2 && ms < best) best = ms;
  }
  // correctness: every element was incremented `steps` times
  std::vector<float> h(n);
  CK(cudaMemcpy(h.data(), (steps & 1) ? b : a, n * sizeof(float), cudaMemcpyDeviceToHost));
  const float expect = 12.0f * steps;      // 12 launches of the graph in total
  for (int i = 0; i < n; i += 97)
    if (h[i] != expect) { printf("  MISMATCH at %d: %f != %f (pdl=%d early=%d)\n", i, h[i], expect, pdl, trigger_early); break; }
  CK(cudaGraphExecDestroy(exec));
  CK(cudaGraphDestroy(graph));
  CK(cudaFree(a));
  CK(cudaFree(b));
  CK(cudaStreamDestroy(st));
  return best * 1e3f / steps;
}

int main() {
  const int steps = 64;
  printf("%-8s %-12s %10s %14s %14s\n", "CTAs", "prologue ns", "plain us", "PDL early us", "PDL late us");
  for (int ctas : {1, 148}) {
    for (int pro : {0, 500, 1500}) {
      const float plain = run_chain(steps, ctas, pro, 0, 0);
      const float early = run_chain(steps, ctas, pro, 1, 1);
      const float late = run_chain(steps, ctas, pro, 1, 0);
      printf("%-8d %-12d %10.2f %14.2f %14.2f\n", ctas, pro, plain, early, late);
    }
  }
  return 0;
}
