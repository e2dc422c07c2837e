// gg_trace.cu — timed event nodes for tools/trace_step.py.  Events recorded with cudaEventRecordExternal during stream
// capture become event-record NODES of the CUDA graph (a plain cudaEventRecord under capture is only a dependency
// marker), so cudaEventElapsedTime between two of them is valid after a replay: a per-group timeline of the captured
// training step without a profiler attached and without extra kernels in the graph.
#include "gg_common.cuh"

using namespace gg;

extern "C" int gg_trace_event_create(void** event_out) {
  cudaEvent_t e;
  cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDefault);
  if (rc != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)rc, "gg_trace_event_create: %s", cudaGetErrorString(rc));
  *event_out = reinterpret_cast<void*>(e);
  return GG_OK;
}

extern "C" int gg_trace_event_record(void* event, void* stream) {
  cudaError_t rc = cudaEventRecordWithFlags(reinterpret_cast<cudaEvent_t>(event), as_stream(stream), cudaEventRecordExternal);
  if (rc != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)rc, "gg_trace_event_record: %s", cudaGetErrorString(rc));
  return GG_OK;
}

extern "C" int gg_trace_event_elapsed_us(void* start, void* end, float* us_out) {
  float ms = 0.f;
  cudaError_t rc = cudaEventElapsedTime(&ms, reinterpret_cast<cudaEvent_t>(start), reinterpret_cast<cudaEvent_t>(end));
  if (rc != cudaSuccess) return fail(GG_ERR_CUDA_BASE + (int)rc, "gg_trace_event_elapsed_us: %s", cudaGetErrorString(rc));
  *us_out = ms * 1e3f;
  return GG_OK;
}
