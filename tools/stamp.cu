// Diagnostic only: a one-thread kernel that writes %globaltimer (ns) into dst -- an in-stream timestamp that can be captured
// into CUDA graphs (tools/diag_shard_shape.py uses it to draw the timeline of the batches in flight).
#include <cstdint>
#include <cuda_runtime.h>
__global__ void stamp_kernel(unsigned long long* dst) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *dst = t;
}
extern "C" int stamp(void* dst, void* stream) {
  stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)dst);
  return (int)cudaGetLastError();
}
