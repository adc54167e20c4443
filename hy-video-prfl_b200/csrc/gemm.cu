// bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA
// into 128B-swizzled shared memory through a 4-stage mbarrier ring; persistent, warp-specialised:
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane)
//   warps 2..5  : epilogue (TMEM -> registers -> fused bias / GELU / gated-residual -> global)
// CTA tile 128 x 256 x 64, UMMA 128x256x16, two 256-column accumulators in TMEM so the epilogue of tile i
// overlaps the main loop of tile i+1.  Roofline: tensor pipe (2*M*N*K flops); smem operand traffic
// 48 KB / 512 MMA cycles = 96 B/clk per SM, below the 128 B/clk shared-memory port.
#include <atomic>

#include "common.cuh"

namespace prfl {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_STAGE = BM * BK * 2;  // 16 KB
constexpr int GEMM_THREADS = 192;

// CG = 1: one CTA per 128 x 256 tile, B stage 32 KB, 4 stages.
// CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x 256 tile: each CTA stages its 128 rows of A and HALF
//         of B (128 of the 256 columns), the leader issues UMMA 256x256x16 that reads both CTAs' shared memory and writes
//         both CTAs' TMEM.  Per-SM operand traffic from L2 and shared memory drops from 48 KB to 32 KB per k-block
//         (96 -> 64 B/clk), which is what a power-capped B200 needs to hold its clocks; 6 stages fit.
template <int CG> struct GemmCfg {
  static constexpr int STAGES = CG == 2 ? 6 : 4;
  static constexpr int B_ROWS = BN / CG;
  static constexpr int B_STAGE = B_ROWS * BK * 2;
  static constexpr int SMEM = STAGES * (A_STAGE + B_STAGE) + 256 + 1024;  // + barriers + alignment slack
};

struct GemmParams {
  void* out;
  int64_t ldc;
  const float* bias;
  const float* gate;
  const float* resid;   // RESIDUAL: the stream that is read (out = resid + gate * bf16(acc + bias)); == out for the in-place form
  __nv_bfloat16* aux;   // DGELU: input (pre-activation); GELU / RESIDUAL: optional output bf16(acc + bias)
  int64_t ldaux;
  int M, N, K, epi, beta;
  int tiles_m, tiles_n, num_kb;
  int group_m;   // raster: tiles are walked m-fastest inside groups of `group_m` row-tiles so a wave shares A and B panels in L2
};

// group_m > 0: groups of `group_m` row-tiles, all column tiles, walked m-fastest (the group's A panels stay in L2 while
//              B streams once per group);
// group_m < 0: the transpose — groups of `-group_m` column-tiles, all row tiles, walked n-fastest (the group's B panels
//              stay in L2 while A streams once per group): fewer bytes when A (tokens) is the larger operand.
__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int group_m, int& m_blk, int& n_blk) {
  if (group_m > 0) {
    const int group = group_m * tiles_n;
    const int g = tile / group;
    const int first_m = g * group_m;
    const int gm = min(tiles_m - first_m, group_m);
    const int r = tile - g * group;
    m_blk = first_m + r % gm;
    n_blk = r / gm;
  } else {
    const int group_n = -group_m;
    const int group = group_n * tiles_m;
    const int g = tile / group;
    const int first_n = g * group_n;
    const int gn = min(tiles_n - first_n, group_n);
    const int r = tile - g * group;
    n_blk = first_n + r % gn;
    m_blk = r / gn;
  }
}

template <bool A_T, bool B_T, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  constexpr int STAGES = GemmCfg<CG>::STAGES, B_STAGE = GemmCfg<CG>::B_STAGE, B_ROWS = GemmCfg<CG>::B_ROWS;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;   // position inside the CTA pair
  const bool leader = cta_rank == 0;
  const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;  // a "unit" = the CTA (pair) that owns whole tiles
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B tiles need 1024-B alignment
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE + B_STAGE));
  uint64_t* full = bars;                  // [STAGES]
  uint64_t* empty = bars + STAGES;        // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;    // [2]
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], CG);     // CG == 2: one arrive (+ the TMA bytes) from each CTA's producer, on the leader's barrier
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], 4 * CG);  // CG == 2: the epilogue warps of BOTH CTAs release the leader's barrier
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_cg2<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();   // peer barriers are initialised before anyone arrives on them remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      auto load = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
        if (CG == 2) tma_load_2d_cg2(dst, tm, bar, c0, c1);   // bytes are credited to the leader CTA's barrier
        else tma_load_2d(dst, tm, bar, c0, c1);
      };
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        int m_blk, n_blk;
        tile_coords(tile, p.tiles_m, p.tiles_n, p.group_m, m_blk, n_blk);
        const int m0 = (m_blk * CG + cta_rank) * BM, n0 = n_blk * BN + cta_rank * B_ROWS;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          if (CG == 1 || leader) mbar_arrive_expect_tx(&full[s], CG * (A_STAGE + B_STAGE));
          uint8_t* a = sA + s * A_STAGE;
          uint8_t* b = sB + s * B_STAGE;
          const int k0 = kb * BK;
          if (!A_T) {
            load(a, &tmA, &full[s], k0, m0);  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) load(a + c * 8192, &tmA, &full[s], m0 + 64 * c, k0);  // box {64 m, 64 k}
          }
          if (!B_T) {
            load(b, &tmB, &full[s], k0, n0);  // box {64 k, 256 / CG n}
          } else {
#pragma unroll
            for (int c = 0; c < B_ROWS / 64; ++c) load(b + c * 8192, &tmB, &full[s], n0 + 64 * c, k0);  // box {64 n, 64 k}
          }
          // the follower's arrival carries no data ordering (its bytes are counted by complete_tx): relaxed, after the loads
          if (CG == 2 && !leader) mbar_arrive_cluster_relaxed(&full[s], 0);
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN, A_T ? 1 : 0, B_T ? 1 : 0);
      uint32_t it = 0, t = 0;
      for (int tile = unit; tile < total_tiles; tile += n_units, ++t) {
        const uint32_t buf = t & 1, aph = (t >> 1) & 1;
        mbar_wait(&tempty[buf], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + s * A_STAGE);
          const uint32_t b_addr = smem_u32(sB + s * B_STAGE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_T ? make_sdesc_sw128(a_addr + k * 2048, 8192, 1024) : make_sdesc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_T ? make_sdesc_sw128(b_addr + k * 2048, 8192, 1024) : make_sdesc_sw128(b_addr + k * 32, 16, 1024);
            if (CG == 2) umma_ss_cg2(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_ss(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
          if (CG == 2) umma_commit_mc2(&empty[s], 3);
          else umma_commit(&empty[s]);
        }
        if (CG == 2) umma_commit_mc2(&tfull[buf], 3);  // accumulator complete (both CTAs' epilogues)
        else umma_commit(&tfull[buf]);
      }
    }
  } else {
    // ---------------- epilogue: thread = one accumulator row ----------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    uint32_t t = 0;
    for (int tile = unit; tile < total_tiles; tile += n_units, ++t) {
      int m_blk, n_blk;
      tile_coords(tile, p.tiles_m, p.tiles_n, p.group_m, m_blk, n_blk);
      const uint32_t buf = t & 1, aph = (t >> 1) & 1;
      const int row = (m_blk * CG + cta_rank) * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      if ((p.epi == PRFL_EPI_RESIDUAL || (p.epi == PRFL_EPI_F32 && p.beta)) && row_ok) {
        // the read-modify-write epilogue is latency-bound on the fp32 tile it updates: pull this thread's row segment
        // (1 KB = 8 lines) into L2 now, while the tile's main loop is still running
        const float* o = (p.epi == PRFL_EPI_RESIDUAL ? p.resid : reinterpret_cast<const float*>(p.out)) + (int64_t)row * p.ldc + n_blk * BN;
#pragma unroll
        for (int j = 0; j < BN / 32; ++j)
          if (n_blk * BN + j * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(o + j * 32));
      }
      mbar_wait(&tfull[buf], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * BN + ((uint32_t)(quad * 32) << 16);
      const int n0 = n_blk * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < p.N) {
              float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        }
        if (row_ok) {
        if (p.epi == PRFL_EPI_BF16 || p.epi == PRFL_EPI_BF16_GELU || p.epi == PRFL_EPI_BF16_DGELU) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (int64_t)row * p.ldc + col0;
          if (p.epi == PRFL_EPI_BF16_GELU) {
            if (p.aux) {
              __nv_bfloat16* ax = p.aux + (int64_t)row * p.ldaux + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (col0 + j < p.N) {
                  uint4 o4;
                  o4.x = pack_bf16x2(v[j], v[j + 1]); o4.y = pack_bf16x2(v[j + 2], v[j + 3]);
                  o4.z = pack_bf16x2(v[j + 4], v[j + 5]); o4.w = pack_bf16x2(v[j + 6], v[j + 7]);
                  *reinterpret_cast<uint4*>(ax + j) = o4;
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(bf16_round(v[j]));
          } else if (p.epi == PRFL_EPI_BF16_DGELU) {
            const __nv_bfloat16* ax = p.aux + (int64_t)row * p.ldaux + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < p.N) {
                uint4 a4 = *reinterpret_cast<const uint4*>(ax + j);
                const uint32_t u[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  v[j + 2 * q] *= gelu_tanh_grad(bf16lo(u[q]));
                  v[j + 2 * q + 1] *= gelu_tanh_grad(bf16hi(u[q]));
                }
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < p.N) {
              uint4 o4;
              o4.x = pack_bf16x2(v[j], v[j + 1]); o4.y = pack_bf16x2(v[j + 2], v[j + 3]);
              o4.z = pack_bf16x2(v[j + 4], v[j + 5]); o4.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = o4;
            }
          }
        } else {
          float* o = reinterpret_cast<float*>(p.out) + (int64_t)row * p.ldc + col0;
          if (p.epi == PRFL_EPI_RESIDUAL) {
            const float* rs = p.resid + (int64_t)row * p.ldc + col0;
            if (p.aux) {
              __nv_bfloat16* ax = p.aux + (int64_t)row * p.ldaux + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (col0 + j < p.N) {
                  uint4 o4;
                  o4.x = pack_bf16x2(v[j], v[j + 1]); o4.y = pack_bf16x2(v[j + 2], v[j + 3]);
                  o4.z = pack_bf16x2(v[j + 4], v[j + 5]); o4.w = pack_bf16x2(v[j + 6], v[j + 7]);
                  *reinterpret_cast<uint4*>(ax + j) = o4;
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < p.N) {
                float4 x4 = *reinterpret_cast<const float4*>(rs + j);
                float4 g4 = p.gate ? __ldg(reinterpret_cast<const float4*>(p.gate + col0 + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
                x4.x += g4.x * bf16_round(v[j]); x4.y += g4.y * bf16_round(v[j + 1]);
                x4.z += g4.z * bf16_round(v[j + 2]); x4.w += g4.w * bf16_round(v[j + 3]);
                *reinterpret_cast<float4*>(o + j) = x4;
              }
            }
          } else {  // PRFL_EPI_F32
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < p.N) {
                float4 x4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (p.beta) {
                  float4 y4 = *reinterpret_cast<const float4*>(o + j);
                  x4.x += y4.x; x4.y += y4.y; x4.z += y4.z; x4.w += y4.w;
                }
                *reinterpret_cast<float4*>(o + j) = x4;
              }
            }
          }
        }
        }  // row_ok
        __syncwarp();  // reconverge before the next warp-collective tcgen05.ld
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(&tempty[buf], 0);
        else mbar_arrive(&tempty[buf]);
      }
    }
  }

  tc_fence_before();
  __syncwarp();
  if (CG == 2) cluster_sync_all();   // the peer may still be the target of multicast commits / hold live accumulators
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

template <bool A_T, bool B_T, int CG>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
  static unsigned long long attr_mask = 0;      // per-device bit mask; forward and autograd threads may race: the call is idempotent
  auto kern = gemm_bf16_kernel<A_T, B_T, CG>;
  constexpr int SMEM = GemmCfg<CG>::SMEM;
  if (device_needs_init(&attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "gemm: cudaFuncSetAttribute");
    device_mark_init(&attr_mask);
  }
  const int total = p.tiles_m * p.tiles_n;          // tiles_m counts (128 * CG)-row tiles
  const int units = sm_count() / CG;
  const int grid = (total < units ? total : units) * CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
  count_launch();
  if (e != cudaSuccess) return cuda_fail(e, "gemm_bf16 launch");
  PRFL_LAUNCH_CHECK("gemm_bf16");
  return PRFL_OK;
}

// PRFL_GEMM_CG=1 forces the single-CTA kernel (debugging / A-B comparison); default is the CTA-pair kernel when M > 128.
static int gemm_cta_group(int M) {
  static int forced = [] {
    const char* e = getenv("PRFL_GEMM_CG");
    return e ? atoi(e) : 0;
  }();
  if (forced == 1 || forced == 2) return forced;
  return M > BM ? 2 : 1;
}

}  // namespace prfl

using namespace prfl;

extern "C" int prfl_gemm_bf16(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans, void* out,
                              int64_t ldc, const float* bias, const float* gate, const float* resid, void* aux_bf16,
                              int64_t ldaux, int M, int N, int K, int epi, int beta, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(M > 0 && N > 0 && K > 0 && N % 8 == 0, PRFL_E_SHAPE, "gemm: M=%d N=%d K=%d (need N%%8==0)", M, N, K);
  PRFL_REQUIRE(K % 8 == 0 || (a_trans && b_trans), PRFL_E_SHAPE, "gemm: K=%d must be a multiple of 8 unless both operands are transposed", K);
  PRFL_REQUIRE(epi >= PRFL_EPI_BF16 && epi <= PRFL_EPI_BF16_DGELU, PRFL_E_SHAPE, "gemm: unknown epilogue %d", epi);
  PRFL_REQUIRE(!(a_trans && M % 8 != 0), PRFL_E_SHAPE, "gemm: transposed A needs M%%8==0 (M=%d)", M);
  PRFL_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0 && lda >= (a_trans ? M : K) && ldb >= (b_trans ? N : K) && ldc >= N,
               PRFL_E_ALIGN, "gemm: leading dims lda=%lld ldb=%lld ldc=%lld", (long long)lda, (long long)ldb, (long long)ldc);
  PRFL_REQUIRE(epi != PRFL_EPI_BF16_DGELU || aux_bf16, PRFL_E_SHAPE, "gemm: DGELU needs aux");
  PRFL_REQUIRE(!aux_bf16 || (ldaux >= N && ldaux % 8 == 0), PRFL_E_ALIGN, "gemm: ldaux=%lld", (long long)ldaux);
  PRFL_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(gate) & 15) == 0 && (reinterpret_cast<uintptr_t>(aux_bf16) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(resid) & 15) == 0 &&
                   ((epi == PRFL_EPI_F32 || epi == PRFL_EPI_RESIDUAL) ? true : ldc % 8 == 0),
               PRFL_E_ALIGN, "gemm: out/bias/gate/aux must be 16-byte aligned");
  const int cg = gemm_cta_group(M);
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_trans) rc = make_tmap_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, 64, BM, 1);
  else rc = make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, 64, 1);
  if (rc != PRFL_OK) return rc;
  if (!b_trans) rc = make_tmap_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, 64, BN / cg, 1);
  else rc = make_tmap_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, 64, 1);
  if (rc != PRFL_OK) return rc;
  GemmParams p;
  p.resid = (epi == PRFL_EPI_RESIDUAL && resid) ? resid : reinterpret_cast<const float*>(out);
  p.out = out; p.ldc = ldc; p.bias = bias; p.gate = gate; p.aux = (__nv_bfloat16*)aux_bf16; p.ldaux = ldaux;
  p.M = M; p.N = N; p.K = K; p.epi = epi; p.beta = beta;
  p.tiles_m = (M + BM * cg - 1) / (BM * cg); p.tiles_n = (N + BN - 1) / BN; p.num_kb = (K + BK - 1) / BK;
  static int group_env = [] {
    const char* e = getenv("PRFL_GEMM_GROUP_M");
    return e ? atoi(e) : 0;
  }();
  // Raster default, from the round-2 sweep on B200 (profiles/r02_gemm_raster.md; 14B block shapes at M = 32 760, same box,
  // CUDA events): row-tile groups of 16 when a group's A panels fit L2 comfortably (K <= 8192: QKV +1 %, ffn.0 +6 %) and for
  // the wgrad shapes (+13 %); column-tile groups of 8 when N is short (o-proj + gated residual: 1 042 -> 1 185 TFLOP/s,
  // the fp32 read-modify-write of a row stays in one DRAM page); row-tile groups of 8 for long K (ffn.2, dgrad).
  int group_auto = 8;
  if (a_trans && b_trans) group_auto = 16;
  else if (K <= 8192) group_auto = p.tiles_n <= 24 ? -8 : 16;
  p.group_m = group_env != 0 ? group_env : group_auto;
  cudaStream_t st = (cudaStream_t)stream;
  if (cg == 2) {
    if (!a_trans && !b_trans) return launch_gemm<false, false, 2>(tmA, tmB, p, st);
    if (!a_trans && b_trans) return launch_gemm<false, true, 2>(tmA, tmB, p, st);
    if (a_trans && !b_trans) return launch_gemm<true, false, 2>(tmA, tmB, p, st);
    return launch_gemm<true, true, 2>(tmA, tmB, p, st);
  }
  if (!a_trans && !b_trans) return launch_gemm<false, false, 1>(tmA, tmB, p, st);
  if (!a_trans && b_trans) return launch_gemm<false, true, 1>(tmA, tmB, p, st);
  if (a_trans && !b_trans) return launch_gemm<true, false, 1>(tmA, tmB, p, st);
  return launch_gemm<true, true, 1>(tmA, tmB, p, st);
}
