// Development microbenchmark: cycles per tcgen05.mma for the instruction shapes the fused kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I self-paced-contrastive-learning_b200/csrc \
//        -o /tmp/mma_bench tools/mma_bench.cu && /tmp/mma_bench
// One CTA per SM issues COUNT back-to-back MMAs from one elected thread and waits for the commit.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx_sm100.cuh"

using namespace spcl::ptx;

constexpr int COUNT = 256;

// variant: 0 SS N128 | 1 SS N256 | 2 TS N128 | 3 TS N256 | 4 SS N128, two D buffers alternating
//          5 TS N128 with B read MN-major | 6 SS N64 | 7 SS N128 with B read MN-major
//          8 SS N128, A read MN-major | 9 SS N128, A and B read MN-major | 10 SS N256, A read MN-major
//          11 SS N256, B read MN-major (4 atoms of 64, one panel apart)
__global__ void __launch_bounds__(128, 1) bench(int variant, int ldtm_noise, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: A 128 x 128 bf16 (2 panels of 16 KB), B 256 x 128 bf16 (2 panels of 32 KB); contents arbitrary
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i * 2654435761u & 0x00ff00ffu);
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + 32768);
  unsigned long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const int n = (variant == 1 || variant == 3 || variant == 10 || variant == 11) ? 256 : (variant == 6 ? 64 : 128);
    const uint32_t idesc = make_idesc_bf16(128, n, variant >= 8 && variant <= 10, variant == 5 || variant == 7 || variant == 9 || variant == 11);
    const uint32_t b_panel = (n == 256) ? 32768u : 16384u;
    __syncwarp();
    t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < COUNT; ++i) {
        const int kk = i & 7;
        const uint32_t off_a = (kk >> 2) * 16384u + (kk & 3) * 32u;
        const uint32_t off_b = (kk >> 2) * b_panel + (kk & 3) * 32u;
        const uint32_t d = tmem + ((variant == 4) ? (i & 1) * 128 : 0);
        if (variant == 2 || variant == 3) {
          mma_ts(d, tmem + 256 + kk * 8, make_smem_desc_sw128(b_base + off_b, 16, 1024), idesc, i != 0);
        } else if (variant == 5) {
          mma_ts(d, tmem + 256 + kk * 8, make_smem_desc_sw128(b_base + kk * 2048, 16384, 1024), idesc, i != 0);
        } else if (variant == 11) {
          mma_ss(d, make_smem_desc_sw128(a_base + off_a, 16, 1024), make_smem_desc_sw128(b_base + kk * 2048, 16384, 1024),
                 idesc, i != 0);
        } else if (variant >= 8) {
          // A read MN-major: 2 panels of 128 k-rows x 64 rows of M (128 B), K = 16 step = 16 k-rows = 2048 B
          const uint64_t ad = make_smem_desc_sw128(a_base + kk * 2048, 16384, 1024);
          const uint64_t bd = variant == 9 ? make_smem_desc_sw128(b_base + kk * 2048, 16384, 1024)
                                           : make_smem_desc_sw128(b_base + off_b, 16, 1024);
          mma_ss(d, ad, bd, idesc, i != 0);
        } else if (variant == 7) {
          mma_ss(d, make_smem_desc_sw128(a_base + off_a, 16, 1024), make_smem_desc_sw128(b_base + kk * 2048, 16384, 1024),
                 idesc, i != 0);
        } else {
          mma_ss(d, make_smem_desc_sw128(a_base + off_a, 16, 1024), make_smem_desc_sw128(b_base + off_b, 16, 1024),
                 idesc, i != 0);
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
    if (lane == 0) mbar_wait(&bar, 0);
    __syncwarp();
    t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp >= 2 && ldtm_noise) {
    // epilogue-like TMEM read traffic on another buffer while the MMAs run
    uint32_t v[32];
    uint32_t acc = 0;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 384;
    for (int i = 0; i < ldtm_noise; ++i) {
      tmem_ld_32x32b_x32(lane_base + (i & 3) * 32, v);
      tmem_wait_ld();
#pragma unroll
      for (int e = 0; e < 32; ++e) acc += v[e];
    }
    if (acc == 0x12345678u) out[1000] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 2048 * sizeof(unsigned long long));
  const size_t smem = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[] = {"SS M128 N128", "SS M128 N256", "TS M128 N128", "TS M128 N256", "SS N128 two D buffers",
                         "TS N128, B MN-major", "SS M128 N64", "SS N128, B MN-major", "SS N128, A MN-major",
                         "SS N128, A+B MN-major", "SS N256, A MN-major",
                         "SS N256, B MN-major"};
  for (int noise : {0}) {
    for (int grid : {148}) {
      for (int v = 0; v < 12; ++v) {
        bench<<<grid, 128, smem>>>(v, noise, d_out);
        bench<<<grid, 128, smem>>>(v, noise, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d failed: %s\n", v, cudaGetErrorString(e)); return 1; }
        std::vector<unsigned long long> h(grid);
        cudaMemcpy(h.data(), d_out, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        double sum = 0;
        for (auto x : h) sum += x;
        printf("ldtm_noise=%d grid=%3d %-24s : %7.1f cycles / MMA (K=16)\n", noise, grid, names[v],
               sum / grid / COUNT);
      }
    }
  }
  return 0;
}
