// DSMEM micro-benchmark: what the SM-to-SM network of THIS GPU sustains in the traffic pattern of lstm_tc_kernel
// (csrc/lstm_tc.cu), so that bench.py's `roofline_network.peak` is a measured number instead of a figure quoted from the
// B300 guide.
//
// Pattern: clusters of 8 CTAs; every step each CTA sends `bytes` (its slice of h) to each of its 7 peers and receives as
// much, gated by a double-buffered mbarrier transaction count exactly like the kernel (no CTA starts step s+1 before the
// slices of step s have arrived from all peers).
//   mode 0: per-thread st.async.v4 (16 bytes + complete_tx), the kernel's mechanism;
//   mode 1: one cp.async.bulk.shared::cluster.shared::cta per (peer, step) issued by a single thread.
// Reported per configuration: bytes per clock and SM, in + out, measured with clock64 over the steady-state loop, for one
// cluster alone and for as many clusters as are co-resident.   Prints one JSON object.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/dsmem_bench tools/dsmem_bench.cu && tools/bin/dsmem_bench
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../text2pos_cvpr2022_b200/csrc/sm100.cuh"

namespace cg = cooperative_groups;
using namespace t2p::sm100;

constexpr int CS = 8;
constexpr int THREADS = 256;
constexpr int MAX_BYTES = 8192;  // per peer and step

struct Bars {
  uint64_t bar[2];
};

__device__ __forceinline__ uint32_t map_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_cluster(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar)
               : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
dsmem_kernel(int mode, int bytes, int steps, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  // recv[2 buffers][7 sources][bytes] | send[bytes] | bars
  uint8_t* recv = raw;
  uint8_t* send = recv + 2 * 7 * MAX_BYTES;
  Bars* bars = reinterpret_cast<Bars*>(send + MAX_BYTES);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  for (int i = tid; i < MAX_BYTES / 4; i += THREADS) reinterpret_cast<uint32_t*>(send)[i] = i * 2654435761u + rank;
  const uint32_t step_bytes = 7u * (uint32_t)bytes;
  if (tid == 0) {
    mbar_init(&bars->bar[0], 1);
    mbar_init(&bars->bar[1], 1);
    mbar_fence_init();
    mbar_expect_tx(&bars->bar[0], step_bytes);
    mbar_expect_tx(&bars->bar[1], step_bytes);
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  cluster.sync();
  uint32_t rdst[7], rbar[7];
#pragma unroll
  for (int p = 0; p < 7; ++p) {
    const uint32_t peer = (uint32_t)((rank + 1 + p) & (CS - 1));
    // my slot in the peer's receive area: source index = (my rank - peer rank - 1) mod 8 in [0,7)
    const uint32_t slot = (uint32_t)((rank - (int)peer - 1) & (CS - 1));
    rdst[p] = map_rank(smem_u32(recv) + slot * MAX_BYTES, peer);
    rbar[p] = map_rank(smem_u32(&bars->bar[0]), peer);
  }
  const int chunks = bytes / 16;  // per peer
  long long t0 = 0;
  for (int step = 0; step < steps; ++step) {
    if (step == 8 && tid == 0) t0 = clock64();  // steady state
    const int buf = step & 1;
    const uint32_t doff = (uint32_t)buf * 7u * MAX_BYTES, boff = (uint32_t)buf * 8u;
    if (mode == 0) {
      for (int c = tid; c < chunks * 7; c += THREADS) {
        const int p = c % 7, ch = c / 7;
        const uint4 v = reinterpret_cast<const uint4*>(send)[ch];
#pragma unroll
        for (int pp = 0; pp < 7; ++pp)
          if (pp == p) st_async_v4(rdst[pp] + doff + ch * 16, v, rbar[pp] + boff);
      }
    } else if (mode == 1) {
      if (tid < 7) {
#pragma unroll
        for (int pp = 0; pp < 7; ++pp)
          if (pp == tid) bulk_copy_cluster(rdst[pp] + doff, smem_u32(send), (uint32_t)bytes, rbar[pp] + boff);
      }
    } else {
      // mode 2: every one of the 8 warps ships its own eighth of the slice (the block an epilogue warp of lstm_tc_kernel
      // produces) with one bulk copy per peer: 56 copies of bytes/8 per step, issued by 7 lanes of each warp
      const int w = tid >> 5, l = tid & 31;
      const uint32_t piece = (uint32_t)bytes / 8u;
      if (l < 7) {
#pragma unroll
        for (int pp = 0; pp < 7; ++pp)
          if (pp == l) bulk_copy_cluster(rdst[pp] + doff + w * piece, smem_u32(send) + w * piece, piece, rbar[pp] + boff);
      }
    }
    // wait for the 7 slices of this step, re-arm the barrier for step + 2
    mbar_wait(&bars->bar[buf], (uint32_t)((step >> 1) & 1));
    __syncthreads();
    if (tid == 0) mbar_expect_tx(&bars->bar[buf], step_bytes);
  }
  if (tid == 0 && cycles_out) cycles_out[blockIdx.x] = clock64() - t0;
  cluster.sync();
}

static double run(int mode, int bytes, int clusters, int steps, int sm_clock_khz, double* ms_out) {
  const size_t smem = 2 * 7 * MAX_BYTES + MAX_BYTES + sizeof(Bars) + 64;
  cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* d_cyc;
  cudaMalloc(&d_cyc, sizeof(long long) * clusters * CS);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CS);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, dsmem_kernel, mode, bytes, 64, (long long*)nullptr);  // warm-up
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, dsmem_kernel, mode, bytes, steps, d_cyc);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "dsmem_bench: %s (mode %d bytes %d clusters %d)\n", cudaGetErrorString(e), mode, bytes, clusters);
    exit(1);
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_out = ms;
  std::vector<long long> cyc(clusters * CS);
  cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * cyc.size(), cudaMemcpyDeviceToHost);
  cudaFree(d_cyc);
  long long worst = 0;
  for (long long c : cyc) worst = c > worst ? c : worst;
  (void)sm_clock_khz;
  // per SM and clock: 7*bytes out + 7*bytes in per step
  return 2.0 * 7.0 * bytes * (steps - 8) / (double)worst;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  int max_clusters = 0;
  {
    const size_t smem = 2 * 7 * MAX_BYTES + MAX_BYTES + sizeof(Bars) + 64;
    cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaOccupancyMaxActiveClusters(&max_clusters, dsmem_kernel, &cfg);
  }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz\": %d, \"cluster_size\": %d, \"max_active_clusters\": %d, \"rows\": [",
         prop.name, prop.multiProcessorCount, khz, CS, max_clusters);
  const int sizes[] = {256, 512, 1024, 2048, 4096, 8192};
  bool first = true;
  double best[3] = {0, 0, 0}, best_lstm[3] = {0, 0, 0};
  const char* names[3] = {"st.async.v4", "cp.async.bulk", "cp.async.bulk x 8 warps"};
  for (int mode = 0; mode < 3; ++mode)
    for (int bytes : sizes)
      for (int clusters : {1, 2, max_clusters > 2 ? max_clusters : 2}) {
        double ms;
        const double bpc = run(mode, bytes, clusters, 2008, khz, &ms);
        printf("%s{\"mode\": \"%s\", \"bytes_per_peer\": %d, \"clusters\": %d, \"bytes_per_clk_per_sm\": %.3f, \"ms\": %.4f}",
               first ? "" : ", ", names[mode], bytes, clusters, bpc, ms);
        first = false;
        if (bpc > best[mode]) best[mode] = bpc;
        if (bytes == 2048 && clusters <= 2 && bpc > best_lstm[mode]) best_lstm[mode] = bpc;  // lstm_tc throughput mode
      }
  printf("], \"peak_st_async_b_per_clk\": %.3f, \"peak_bulk_b_per_clk\": %.3f, \"peak_bulk8_b_per_clk\": %.3f, "
         "\"lstm_shape_st_async_b_per_clk\": %.3f, \"lstm_shape_bulk_b_per_clk\": %.3f, \"lstm_shape_bulk8_b_per_clk\": %.3f, "
         "\"lstm_shape\": \"2048 B per peer and step (16 sequences x 32 units x fp16 hi+lo), 7 peers, 1-2 clusters\"}\n",
         best[0], best[1], best[2], best_lstm[0], best_lstm[1], best_lstm[2]);
  return 0;
}
