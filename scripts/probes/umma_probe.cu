// Stand-alone probe (not part of the library): issue rate of back-to-back tcgen05.mma kind::f16 (bf16, K = 16)
// as a function of how many operand bytes each SM must fetch from shared memory per MMA.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rlzero_b200/csrc -o umma_probe umma_probe.cu -lcuda
//
// For each shape one elected thread per CTA (pair) issues `iters` groups of 4 MMAs (K = 64) on operands that are
// already in shared memory (contents irrelevant), cycling through several A / B tiles so no operand can be reused,
// then commits and waits.  No loads, no epilogue: the time is the tensor pipe + its operand fetch only.
// Reported: cycles per MMA, the floor M*N/(2*4096) (= MACs / 8192 per clk and SM... see below), and bytes per clk
// and SM the operands need at full rate.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "rz_tc.cuh"

namespace rz { void set_error_tmap(const char*, int) {} }

template <int kCG>
__global__ void __launch_bounds__(128, 1)
probe_kernel(int M, int N, int iters, int a_row_shift, int b_row_shift, long long* cycles_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = rz::smem_u32(smem);
  // 4 A tiles of [128 rows][64 k] (16 KB each) then 2 B tiles of [<=256 rows][64 k] (32 KB each)
  const uint32_t a0 = base, b0 = base + 4 * 16384, ctrl = base + 4 * 16384 + 2 * 32768;
  const uint32_t bar = ctrl;
  uint32_t* holder = reinterpret_cast<uint32_t*>(smem + 4 * 16384 + 2 * 32768 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kCG == 2 ? rz::cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < (4 * 16384 + 2 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { rz::mbar_init(bar, 1); rz::fence_barrier_init(); }
  if (warp == 1) {
    if (kCG == 2) { rz::tmem_alloc_pair(rz::smem_u32(holder), 256); rz::tmem_relinquish_pair(); }
    else          { rz::tmem_alloc(rz::smem_u32(holder), 256); rz::tmem_relinquish(); }
  }
  rz::fence_proxy_async();
  rz::tc_fence_before();
  if (kCG == 2) rz::cluster_sync_all(); else __syncthreads();
  rz::tc_fence_after();
  const uint32_t tmem = *holder;
  long long t0 = 0, t1 = 0;
  if (warp == 0 && lane == 0 && rank == 0) {
    const uint32_t idesc = rz::umma_idesc_bf16(M, N);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // a_row_shift / b_row_shift: start the operand tile that many 128-byte rows into the buffer (the trunk
      // convolution reads its 9 taps as row-shifted views of one halo tile: shifts of 16*dy + dx rows)
      const uint64_t ad = rz::umma_desc_sw128(a0 + (uint32_t)(it & 1) * 32768u + (uint32_t)a_row_shift * 128u);
      const uint64_t bd = rz::umma_desc_sw128(b0 + (uint32_t)(it & 1) * 16384u + (uint32_t)b_row_shift * 128u);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kCG == 2) rz::umma_bf16_pair(tmem, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, 1u);
        else          rz::umma_bf16(tmem, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, 1u);
      }
    }
    if (kCG == 2) rz::umma_commit_pair(bar, 1); else rz::umma_commit(bar);
    rz::mbar_wait(bar, 0);
    t1 = clock64();
    cycles_out[blockIdx.x / kCG] = t1 - t0;
  }
  rz::tc_fence_before();
  if (kCG == 2) rz::cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    rz::tc_fence_after();
    if (kCG == 2) rz::tmem_dealloc_pair(tmem, 256); else rz::tmem_dealloc(tmem, 256);
  }
}

template <int kCG>
static void run(int M, int N, int iters, int ctas, int a_shift = 0, int b_shift = 0) {
  const int smem = 4 * 16384 + 2 * 32768 + 1024;
  cudaFuncSetAttribute(probe_kernel<kCG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  const int workers = ctas / kCG;
  cudaMalloc(&d, workers * sizeof(long long));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel<kCG>, M, N, iters, a_shift, b_shift, d);
    if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return; }
  }
  long long* h = (long long*)malloc(workers * sizeof(long long));
  cudaMemcpy(h, d, workers * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0; long long mx = 0;
  for (int i = 0; i < workers; ++i) { mean += (double)h[i]; if (h[i] > mx) mx = h[i]; }
  mean /= workers;
  const double n_mma = 4.0 * iters;
  const int m_sm = M / kCG;                                  // rows of A each SM holds / computes
  const double floor_clk = (double)m_sm * N / 256.0;         // m_sm*N*16 MACs at 4096 bf16 MACs per clk and SM
  const double a_bytes = m_sm * 16 * 2.0, b_bytes = (double)(N / kCG) * 16 * 2.0;
  printf("{\"cta_group\": %d, \"M\": %d, \"N\": %d, \"ctas\": %d, \"a_row_shift\": %d, \"b_row_shift\": %d, \"cycles_per_mma_mean\": %.2f, \"cycles_per_mma_worst\": %.2f, "
         "\"floor_cycles\": %.1f, \"pipe_frac\": %.3f, \"smem_bytes_per_mma_per_sm\": %.0f, \"smem_B_per_clk_at_full_rate\": %.1f, "
         "\"smem_B_per_clk_achieved\": %.1f}\n",
         kCG, M, N, ctas, a_shift, b_shift, mean / n_mma, (double)mx / n_mma, floor_clk, floor_clk / (mean / n_mma), a_bytes + b_bytes,
         (a_bytes + b_bytes) / floor_clk, (a_bytes + b_bytes) / (mean / n_mma));
  free(h); cudaFree(d);
}

int main() {
  const int iters = 4096;
  // cta_group::1
  run<1>(128, 64, iters, 148);
  run<1>(128, 128, iters, 148);
  run<1>(128, 256, iters, 148);
  run<1>(64, 256, iters, 148);
  // cta_group::2
  run<2>(256, 64, iters, 148);
  run<2>(256, 128, iters, 148);     // the trunk convolution's shape
  run<2>(256, 256, iters, 148);     // the cuBLAS-style tile
  run<2>(128, 256, iters, 148);
  // one worker alone (no neighbours competing for anything)
  run<2>(256, 128, iters, 2);
  run<1>(128, 128, iters, 1);
  // row-shifted A views (the convolution's taps): 128-byte rows, 8-row swizzle atoms of 1024 B
  const int shifts[] = {1, 2, 4, 7, 8, 9, 16, 17, 33};
  for (int s : shifts) run<2>(256, 128, iters, 148, s, 0);
  for (int s : shifts) run<1>(128, 128, iters, 148, s, 0);
  run<2>(256, 128, iters, 148, 0, 1);
  run<2>(256, 128, iters, 148, 1, 1);
  return 0;
}
