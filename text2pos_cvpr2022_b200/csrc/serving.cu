// Host-side fast path of the serving engine (no device code): everything one batch needs from the host in ONE call --
// lay the raw text out in the pinned staging buffer, enqueue its H2D copy, launch the captured CUDA graph of the step
// (device tokeniser .. select, or the whole sharded step), enqueue the D2H copy of the results and record the completion
// event.  The interpreter then pays for one foreign call per batch instead of six (stream context, staging, copy, graph
// replay, copy, event record), which is what bounded the end-to-end rate once the GPU step dropped below ~60 us.
#include "common.cuh"

#include <algorithm>
#include <mutex>
#include <vector>

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

int t2p_serving_replay_many(void* const* graph_execs, const t2p_stream* streams, int n, t2p_stream origin, int fork_join) {
  T2P_REQUIRE(graph_execs && streams && n >= 0, T2P_ERR_INVALID, "serving_replay_many: null argument");
  // fork_join: the distinct streams first wait for everything enqueued on `origin` so far, and `origin` waits for them at the end
  // (what a caller timing the region with events on `origin` needs), without a round trip through the interpreter per stream
  static std::mutex mu;
  static std::vector<cudaEvent_t> pool;
  std::vector<cudaStream_t> uniq;
  if (fork_join) {
    for (int i = 0; i < n; ++i) {
      cudaStream_t s = t2p::as_stream(streams[i]);
      if (std::find(uniq.begin(), uniq.end(), s) == uniq.end()) uniq.push_back(s);
    }
  }
  std::lock_guard<std::mutex> lock(mu);
  cudaStream_t o = t2p::as_stream(origin);
  if (fork_join) {
    while (pool.size() < uniq.size() + 1) {
      cudaEvent_t e;
      T2P_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      pool.push_back(e);
    }
    T2P_CUDA(cudaEventRecord(pool[0], o));
    for (cudaStream_t s : uniq)
      if (s != o) T2P_CUDA(cudaStreamWaitEvent(s, pool[0], 0));
  }
  for (int i = 0; i < n; ++i)
    T2P_CUDA(cudaGraphLaunch(static_cast<cudaGraphExec_t>(graph_execs[i]), t2p::as_stream(streams[i])));
  if (fork_join) {
    for (size_t j = 0; j < uniq.size(); ++j) {
      if (uniq[j] == o) continue;
      T2P_CUDA(cudaEventRecord(pool[j + 1], uniq[j]));
      T2P_CUDA(cudaStreamWaitEvent(o, pool[j + 1], 0));
    }
  }
  return T2P_OK;
}

}  // extern "C"
