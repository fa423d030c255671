// Host-side helpers of the graph-replayed update path: the per-update host work of a training loop is a handful of CUDA
// runtime calls (graph launch, event record / wait); issued through torch's Python stream / graph objects each costs
// 5-20 us of interpreter time, through one C call about 1 us.
#include "ssac_common.cuh"

using namespace ssac;

#define SSAC_RT(call, what)                                                                \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ::ssac::set_error(std::string(what) + ": " + cudaGetErrorString(e__));               \
      return (int)e__;                                                                     \
    }                                                                                      \
  } while (0)

extern "C" {

int ssac_event_record(void* event, void* stream) {
  SSAC_REQUIRE(event, "ssac_event_record: null event");
  SSAC_RT(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream), "ssac_event_record");
  return 0;
}

int ssac_stream_wait_event(void* stream, void* event) {
  SSAC_REQUIRE(event, "ssac_stream_wait_event: null event");
  SSAC_RT(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0), "ssac_stream_wait_event");
  return 0;
}

int ssac_graph_launch(void* graph_exec, void* stream, void* done_event) {
  SSAC_REQUIRE(graph_exec, "ssac_graph_launch: null graph");
  SSAC_RT(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream), "ssac_graph_launch");
  if (done_event) SSAC_RT(cudaEventRecord((cudaEvent_t)done_event, (cudaStream_t)stream), "ssac_graph_launch (event)");
  return 0;
}

int ssac_pipelined_launch(void* graph_exec, void* launch_stream, void* caller_stream, void* caller_ready_event,
                          void* done_event) {
  SSAC_REQUIRE(graph_exec && caller_ready_event, "ssac_pipelined_launch: null handle");
  SSAC_RT(cudaEventRecord((cudaEvent_t)caller_ready_event, (cudaStream_t)caller_stream), "ssac_pipelined_launch (record)");
  SSAC_RT(cudaStreamWaitEvent((cudaStream_t)launch_stream, (cudaEvent_t)caller_ready_event, 0), "ssac_pipelined_launch (wait)");
  SSAC_RT(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)launch_stream), "ssac_pipelined_launch");
  if (done_event) SSAC_RT(cudaEventRecord((cudaEvent_t)done_event, (cudaStream_t)launch_stream), "ssac_pipelined_launch (event)");
  return 0;
}

}  // extern "C"
