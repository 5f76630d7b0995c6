// Development microbenchmark 2: how groups of 8 tcgen05.mma + tcgen05.commit behave (issue blocking, commit cost,
// effect of warps spinning on mbarriers, effect of the operand placement).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I self-paced-contrastive-learning_b200/csrc \
//        -o /tmp/mma_bench2 tools/mma_bench2.cu && /tmp/mma_bench2
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx_sm100.cuh"

using namespace spcl::ptx;

constexpr int GROUPS = 32;

// mode bits: 1 = commit after every group (two commits), 2 = 8 extra warps spin on a never-completing mbarrier,
//            4 = A operand from TMEM (TS), 8 = N = 256, 16 = wait for each group's commit before the next group
__global__ void __launch_bounds__(384, 1) bench(int mode, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[GROUPS * 2 + 2];
  __shared__ uint32_t tmem_base_s;
  __shared__ unsigned long long stamps[GROUPS + 1];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu);
  if (warp == 10) tmem_alloc<512>(&tmem_base_s);
  if (threadIdx.x == 32) {
    for (int i = 0; i < GROUPS * 2 + 2; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
    stop = 0;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 32768);
  if (warp == 9) {
    const int n = (mode & 8) ? 256 : 128;
    const uint32_t idesc = make_idesc_bf16(128, n, false, false);
    const uint32_t b_panel = (n == 256) ? 32768u : 16384u;
    __syncwarp();
    const unsigned long long t0 = clock64();
    for (int g = 0; g < GROUPS; ++g) {
      if (elect_one()) {
        const uint32_t d = tmem + ((mode & 8) ? (g & 1) * 256 : (g & 3) * 128);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off_a = (kk >> 2) * 16384u + (kk & 3) * 32u;
          const uint32_t off_b = (kk >> 2) * b_panel + (kk & 3) * 32u;
          if (mode & 4)
            mma_ts(d, tmem + 448 + kk * 8, make_smem_desc_sw128(b_base + off_b, 16, 1024), idesc, kk != 0);
          else
            mma_ss(d, make_smem_desc_sw128(a_base + off_a, 16, 1024), make_smem_desc_sw128(b_base + off_b, 16, 1024),
                   idesc, kk != 0);
        }
        if (mode & 1) {
          tc_commit(&bars[2 * g]);
          tc_commit(&bars[2 * g + 1]);
        }
      }
      __syncwarp();
      if ((mode & 16) && (mode & 1)) {
        if (lane == 0) mbar_wait(&bars[2 * g], 0);
        __syncwarp();
      }
      if (lane == 0) stamps[g] = clock64() - t0;
    }
    if (elect_one()) tc_commit(&bars[GROUPS * 2]);
    __syncwarp();
    if (lane == 0) mbar_wait(&bars[GROUPS * 2], 0);
    __syncwarp();
    const unsigned long long t1 = clock64();
    if (lane == 0) {
      stop = 1;
      if (blockIdx.x == 0) {
        out[0] = t1 - t0;
        for (int g = 0; g < GROUPS; ++g) out[1 + g] = stamps[g];
      }
    }
  } else if (warp < 8 && (mode & 2)) {
    // spin like the epilogue warps of the fused kernels do while they wait for a tile
    while (!stop) (void)mbar_try_wait(&bars[GROUPS * 2 + 1], 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 256 * sizeof(unsigned long long));
  const size_t smem = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode : {0, 1, 3, 17, 19, 4, 5, 7, 8, 9, 11, 12, 13}) {
    bench<<<148, 384, smem>>>(mode, d_out);
    bench<<<148, 384, smem>>>(mode, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
    std::vector<unsigned long long> h(1 + GROUPS);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    printf("mode %2d [%s%s%s%s%s]: total %6llu cyc = %6.1f / MMA ; issue stamps:", mode, (mode & 1) ? "commit " : "",
           (mode & 2) ? "spin " : "", (mode & 4) ? "TS " : "SS ", (mode & 8) ? "N256 " : "N128 ",
           (mode & 16) ? "waitEach" : "", h[0], double(h[0]) / (GROUPS * 8));
    for (int g = 0; g < 8; ++g) printf(" %llu", h[1 + g]);
    printf(" ... %llu\n", h[GROUPS]);
  }
  return 0;
}
