// Host-side fast path of the serving engine (no device code): everything one batch needs from the host in ONE call --
// lay the raw text out in the pinned staging buffer, enqueue its H2D copy, launch the captured CUDA graph of the step
// (device tokeniser .. select, or the whole sharded step), enqueue the D2H copy of the results and record the completion
// event.  The interpreter then pays for one foreign call per batch instead of six (stream context, staging, copy, graph
// replay, copy, event record), which is what bounded the end-to-end rate once the GPU step dropped below ~60 us.
#include "common.cuh"

extern "C" {

int t2p_serving_submit(const char* texts, size_t total_bytes, int n_texts, void* h_stage, size_t stage_capacity, void* d_stage,
                       void* graph_exec, const void* d_out, void* h_out, size_t out_bytes, void* event, t2p_stream stream,
                       size_t* used_bytes, int* all_ascii) {
  T2P_REQUIRE(texts && h_stage && d_stage && graph_exec && d_out && h_out && all_ascii, T2P_ERR_INVALID, "serving_submit: null argument");
  int ascii = 0;
  // stages and (ASCII batches only) enqueues the H2D copy
  T2P_TRY(t2p_stage_texts(texts, total_bytes, n_texts, h_stage, stage_capacity, d_stage, stream, used_bytes, &ascii));
  *all_ascii = ascii;
  if (!ascii) return T2P_OK;  // nothing launched: the caller tokenises on the host (Unicode rules) and takes the slow path
  cudaStream_t s = t2p::as_stream(stream);
  T2P_CUDA(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_exec), s));
  T2P_CUDA(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
  if (event) T2P_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), s));
  return T2P_OK;
}

}  // extern "C"
