// Development microbenchmark 3: tcgen05.mma with cta_group::2 (CTA pair, M = 256) against the cta_group::1
// shapes, issued like the fused kernels do (groups of 8 K-steps, a new accumulator per group).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I self-paced-contrastive-learning_b200/csrc \
//        -o tools/mma_bench3.bin tools/mma_bench3.cu && tools/mma_bench3.bin
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx_sm100.cuh"

using namespace spcl::ptx;

constexpr int GROUPS = 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma2_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      : : "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      : : "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      : : "r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}

// variant: 0 cta1 SS N128 | 1 cta2 SS N128 | 2 cta2 SS N256 | 3 cta2 TS N128 (B K-major) | 4 cta2 TS N128 (B MN-major)
//          5 cta1 TS N128 (B MN-major) | 6 cta2 TS N256 | 7 cta1 SS N256 | 8 cta1 TS N128 (B K-major)
// dpat: 0 = two D buffers alternating, new accumulation per group | 1 = one D, always accumulate | 2 = four D (N128) rotating
__global__ void __launch_bounds__(128, 1) bench(int variant, int dpat, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool two = (variant >= 1 && variant <= 4) || variant == 6;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu);
  if (warp == 0) {
    if (two) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                   "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      tmem_alloc<512>(&tmem_base_s);
    }
  }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 32768);
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const int n = (variant == 2 || variant == 6 || variant == 7) ? 256 : 128;
    const bool mn = (variant == 4 || variant == 5);
    const uint32_t idesc = make_idesc_bf16(two ? 256 : 128, n, false, mn);
    const int n_cta = two ? n / 2 : n;                 // rows of B held by one CTA
    const uint32_t b_panel = static_cast<uint32_t>(n_cta) * 128u;
    const bool ts = (variant == 3 || variant == 4 || variant == 5 || variant == 6 || variant == 8);
    __syncwarp();
    t0 = clock64();
    if (!two || rank == 0) {
      if (elect_one()) {
        for (int g = 0; g < GROUPS; ++g) {
          const uint32_t d = tmem + (dpat == 1 ? 0 : (n == 256) ? (g & 1) * 256 : (dpat == 2 ? (g & 3) * 128 : (g & 1) * 128));
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t acc = (dpat == 1) ? ((g | kk) != 0) : (kk != 0);
            const uint32_t off_a = (kk >> 2) * 16384u + (kk & 3) * 32u;
            const uint32_t off_b = (kk >> 2) * b_panel + (kk & 3) * 32u;
            const uint64_t bd = mn ? make_smem_desc_sw128(b_base + kk * 2048, 16384, 1024)
                                   : make_smem_desc_sw128(b_base + off_b, 16, 1024);
            if (two) {
              if (ts) mma2_ts(d, tmem + 448 + kk * 8, bd, idesc, acc);
              else mma2_ss(d, make_smem_desc_sw128(a_base + off_a, 16, 1024), bd, idesc, acc);
            } else {
              if (ts) mma_ts(d, tmem + 448 + kk * 8, bd, idesc, acc);
              else mma_ss(d, make_smem_desc_sw128(a_base + off_a, 16, 1024), bd, idesc, acc);
            }
          }
        }
        if (two) commit2(&bar); else tc_commit(&bar);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_wait(&bar, 0);
    __syncwarp();
    t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (two) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    else tmem_dealloc<512>(tmem);
  }
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 2048 * sizeof(unsigned long long));
  const size_t smem = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"cta1 SS M128 N128", "cta2 SS M256 N128", "cta2 SS M256 N256", "cta2 TS M256 N128 B-K",
                         "cta2 TS M256 N128 B-MN", "cta1 TS M128 N128 B-MN", "cta2 TS M256 N256", "cta1 SS M128 N256",
                         "cta1 TS M128 N128 B-K"};
  auto launch = [&](int v, int dpat, int cl) {
    if (cl == 1) { bench<<<148, 128, smem>>>(v, dpat, d_out); return; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, bench, v, dpat, d_out);
    if (le != cudaSuccess) printf("launch failed: %s\n", cudaGetErrorString(le));
  };
  for (int i = 0; i < 500; ++i) launch(2, 0, 2);
  cudaDeviceSynchronize();
  for (int cl : {2, 1}) {
    for (int dpat : {0}) {
      cudaMemset(d_out, 0, 2048 * 8);
      for (int v = 0; v < 9; ++v) {
        const bool two = (v >= 1 && v <= 4) || v == 6;
        if (two && cl == 1) continue;
        double best = 1e30;
        for (int rep = 0; rep < 3; ++rep) {
          launch(v, dpat, cl);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("variant %d failed: %s\n", v, cudaGetErrorString(e)); return 1; }
          std::vector<unsigned long long> h(148);
          cudaMemcpy(h.data(), d_out, 148 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
          double sum = 0;
          for (int i = 0; i < 148; i += 2) sum += h[i];
          const double per = sum / 74 / (GROUPS * 8);
          best = per < best ? per : best;
        }
        const int n = (v == 2 || v == 6 || v == 7) ? 256 : 128;
        printf("cluster=%d dpat=%d %-26s : %6.1f cycles / MMA = %6.1f per 128x128x16 per SM\n", cl, dpat, names[v], best,
               best * 128.0 / n);
      }
    }
  }
  return 0;
}
