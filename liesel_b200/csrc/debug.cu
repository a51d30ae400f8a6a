// Measurement hook of the product library: CUDA events recorded around the dominant row-streaming
// kernel of a compute call, so bench.py can time that kernel live without a profiler. (The pipe-peak
// micro-benchmarks and the tcgen05 probe live in the separate tools library libliesel_b200_bench.so:
// microbench.cu, tc_probe.cu.)
#include "common.cuh"

namespace lslb {
thread_local cudaEvent_t g_ev_start = nullptr;
thread_local cudaEvent_t g_ev_stop = nullptr;
thread_local int g_ev_count = 0;

void debug_record_start(cudaStream_t s) {
  if (g_ev_start) cudaEventRecord(g_ev_start, s);
}
void debug_record_stop(cudaStream_t s) {
  if (g_ev_stop) {
    cudaEventRecord(g_ev_stop, s);
    ++g_ev_count;
  }
}
}  // namespace lslb

using namespace lslb;

extern "C" int lslb_debug_set_events(void* start, void* stop) {
  g_ev_start = (cudaEvent_t)start;
  g_ev_stop = (cudaEvent_t)stop;
  g_ev_count = 0;
  return LSLB_OK;
}
extern "C" int lslb_debug_event_count(void) { return g_ev_count; }

