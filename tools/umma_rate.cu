// Microbenchmark: sustained issue rate of tcgen05.mma (kind::f16, bf16) per SM as a function of where A comes from
// (shared memory = SS, tensor memory = TS), of N, and of the operand majors.  Answers the design question behind the
// attention-backward kernels: is an M=128, N=64 SS MMA bound by the 32-cycle tensor floor or by the shared-memory
// operand fetch (A 4 KB + B 2 KB per instruction)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate tools/umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdlib>

#include "../hy-video-prfl_b200/csrc/common.cuh"

using namespace prfl;

// MODE 0: SS, MODE 1: TS.  One CTA per SM, one elected thread issues ITER groups of 8 k-steps.
template <int MODE, int N, int A_MN, int B_MN>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* cycles, int iters, int extra_smem_reader) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, N, A_MN, B_MN);
      constexpr uint32_t HI = sdesc_hi(1024);
      // A tile: 128 rows x 128 K (two 64-wide K blocks of 16 KB) at smem+0; B tile: N rows x 128 K at smem+32 KB
      const uint32_t a_lo = sdesc_lo(smem_u32(smem), A_MN ? 16384 : 16);
      const uint32_t b_lo = sdesc_lo(smem_u32(smem) + 32768, B_MN ? (uint32_t)(N <= 64 ? 8192 : 16384) : 16);
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tb + 256 + (it & 0) * N;   // accumulator (kept in columns [256, 256+N))
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t offa = A_MN ? k * (2048 >> 4) : (k >> 2) * (16384 >> 4) + (k & 3) * 2;
          const uint32_t offb = B_MN ? k * (2048 >> 4) : (k >> 2) * ((N * 128) >> 4) + (k & 3) * 2;
          if (MODE == 0) umma_ss(d, sdesc_join(a_lo + offa, HI), sdesc_join(b_lo + offb, HI), idesc, 1u);
          else umma_ts(d, tb + k * 8, sdesc_join(b_lo + offb, HI), idesc, 1u);
        }
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
  } else if (extra_smem_reader) {
    // optional: the other warps stream shared memory with ld.shared.v4 to probe contention with the operand fetch
    uint4 acc = make_uint4(0, 0, 0, 0);
    const uint4* p = reinterpret_cast<const uint4*>(smem + 96 * 1024);
    for (int it = 0; it < iters * 4; ++it) {
      const uint4 v = p[(threadIdx.x + it * 96) & 2047];
      acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    if (acc.x == 0x12345678u) cycles[1] = acc.y;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tb);
  }
}

template <int MODE, int N, int A_MN, int B_MN>
static void run(const char* name, long long* d_cycles, int extra = 0) {
  const int iters = 2000, smem = 164 * 1024;
  cudaFuncSetAttribute(rate_kernel<MODE, N, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h = 0;
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<MODE, N, A_MN, B_MN><<<148, 128, smem>>>(d_cycles, iters, extra);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-44s FAILED: %s\n", name, cudaGetErrorString(e));
      exit(1);
    }
  }
  cudaMemcpy(&h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost);
  const double per = (double)h / (iters * 8.0);
  const double floor_c = 128.0 * N / 256.0;
  const double a_bytes = MODE == 0 ? 128 * 16 * 2 : 0, b_bytes = N * 16 * 2;
  printf("%-44s %7.1f cyc/MMA  (tensor floor %5.1f, %4.0f%% of floor rate; smem operand fetch %5.1f B/cyc)\n", name, per,
         floor_c, 100.0 * floor_c / per, (a_bytes + b_bytes) / per);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  run<0, 64, 0, 0>("SS M128 N64  A K-major  B K-major", d);
  run<0, 128, 0, 0>("SS M128 N128 A K-major  B K-major", d);
  run<0, 256, 0, 0>("SS M128 N256 A K-major  B K-major", d);
  run<0, 64, 0, 1>("SS M128 N64  A K-major  B MN-major", d);
  run<0, 128, 0, 1>("SS M128 N128 A K-major  B MN-major", d);
  run<0, 64, 1, 1>("SS M128 N64  A MN-major B MN-major", d);
  run<0, 128, 1, 1>("SS M128 N128 A MN-major B MN-major", d);
  run<1, 64, 0, 0>("TS M128 N64  A tmem     B K-major", d);
  run<1, 128, 0, 0>("TS M128 N128 A tmem     B K-major", d);
  run<1, 128, 0, 1>("TS M128 N128 A tmem     B MN-major", d);
  run<1, 256, 0, 0>("TS M128 N256 A tmem     B K-major", d);
  run<0, 64, 0, 0>("SS M128 N64  + 3 warps ld.shared.v4", d, 1);
  run<0, 128, 0, 0>("SS M128 N128 + 3 warps ld.shared.v4", d, 1);
  run<1, 128, 0, 1>("TS M128 N128 + 3 warps ld.shared.v4", d, 1);
  return 0;
}
