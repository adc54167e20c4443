// Flash-attention backward for sm_100a (bf16, head_dim 128, non-causal, ragged tails), tcgen05 + TMEM + TMA.
// Two kernels from one template, both with the forward kernel's skeleton (TMA warp, MMA warp, 2 x 256 softmax-gradient
// threads — two threads per row — that ping-pong over streamed "units"), no atomics, deterministic:
//
//   dK/dV : CTA = one head x 128 keys (K_j, V_j resident), loop over query units
//        S^T = K Q^T, dP^T = V dO^T           (M = keys, N = unit queries)             -> TMEM
//        P^T = exp2(S^T c - lse), dS^T = P^T (dP^T - delta)   (thread = key row)       -> bf16 back into TMEM
//        dV += P^T dO, dK += dS^T Q           (TS MMA, A from TMEM, B = dO / Q as MN-major smem operands)
//   dQ    : CTA = one head x 128 queries (Q, dO resident), loop over 64-key units
//        S = Q K^T, dP = dO V^T               (M = queries, N = 64 keys)               -> TMEM
//        dS = P (dP - delta)                  (thread = query row)                      -> bf16 back into TMEM
//        dQ += dS K                           (TS MMA, B = K as MN-major smem operand); scaled in the epilogue
//
// The split costs 7 instead of 5 tile-GEMMs per (q, k) tile pair (S and dP are recomputed in the dQ kernel) but needs
// no cross-CTA reduction of dQ.
//
// Where the resident operand lives.  Measured on B200 (tools/umma_rate.cu -> profiles/r01_umma_rate.txt): a tcgen05.mma
// whose A operand comes from shared memory is bound by the 128 B/cycle shared-memory operand fetch, not by the tensor
// pipe — an M=128, N=64, K=16 "SS" MMA (A 4 KB + B 2 KB) takes 48 cycles instead of its 32-cycle floor — while with A
// in tensor memory ("TS") every shape runs at the floor.  Re-reading the resident 128x128 operand from shared memory
// for every streamed unit caps the S / dP products at 67 % of the pipe.  With TS = true the resident operands are written
// ONCE into tensor memory as packed-bf16 A operands (64 columns each) and all products are TS MMAs; shared memory only
// holds the streamed tiles.  TMEM budget (512 columns):
//   dQ    (TS)      : X0 X1 [0,128)  Y0 Y1 [128,256)  dQ [256,384)              Q [384,448) dO [448,512)
//   dK/dV (TS)      : X0 X1 [0,64)   Y0 Y1 [64,128)   K [128,192) V [192,256)   dV [256,384) dK [384,512)   (32-query units)
//   dK/dV (TS=false): X0 X1 [0,128)  Y0 Y1 [128,256)  dV [256,384) dK [384,512)                              (64-query units)
// Streamed tiles always arrive as 64-row TMA stages; the 32-wide variant consumes a stage as two units.
//
// A pre-pass writes -lse*log2(e) and -delta = -rowsum(dO * O) into a workspace padded to a multiple of 64 queries
// (-inf / 0 in the tail, so tail queries contribute exactly 0); the dK/dV kernel's TMA warp copies the 64 values of a
// stage next to the Q / dO tiles, so the inner loop has no global loads and no block barriers.
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace prfl {

constexpr int BWD_THREADS = 576;            // 16 softmax-gradient warps (2 threads per row) + TMA warp + MMA warp
constexpr int W_TMA = 16, W_MMA = 17;
constexpr int BIG = 128;                     // resident tile rows
constexpr int SUB = 64;                      // rows of one streamed TMA stage
constexpr int BIG_BYTES = BIG * 128 * 2;     // 32 KB
constexpr int SUB_BYTES = SUB * 128 * 2;     // 16 KB
constexpr int MAX_STAGES = 4;

template <bool TS>
struct BwdSmem {
  static constexpr int STAGES = TS ? 4 : 3;
  static constexpr int RES_BYTES = TS ? 0 : 2 * BIG_BYTES;
  static constexpr int BYTES = RES_BYTES + STAGES * (2 * SUB_BYTES + 2 * SUB * 4) + 256 + 1024;
};

struct AttnBwdParams {
  const float* nlse2;   // [H, Lp]  -lse * log2(e)   (Lp = Lq rounded up to 64; -inf in the tail)
  const float* ndelta;  // [H, Lp]  -rowsum(dO * O)  (0 in the tail)
  const __nv_bfloat16* res0;   // resident operand 0 (K | Q) and 1 (V | dO): [L_res, H, 128] strided (TS only)
  const __nv_bfloat16* res1;
  int64_t r0_ld_tok, r0_ld_head, r1_ld_tok, r1_ld_head;
  __nv_bfloat16* out0;  // dkdv: dK ; dq: dQ
  __nv_bfloat16* out1;  // dkdv: dV
  int64_t o0_ld_tok, o0_ld_head, o1_ld_tok, o1_ld_head;
  int Lq, Lk, Lp;
  float scale, scale_log2;
};

// ws[0][h][i] = -lse[h][i] * log2(e), ws[1][h][i] = -sum_d dO[i,h,d] * O[i,h,d]   — one warp per (token, head), i < Lp
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t o_ld_tok, int64_t o_ld_head,
                                  const __nv_bfloat16* __restrict__ dout, int64_t do_ld_tok, int64_t do_ld_head,
                                  const float* __restrict__ lse, float* __restrict__ ws, int Lq, int Lp, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)Lp * H) return;
  const int h = (int)(w % H);
  const int64_t i = w / H;
  float s = 0.f, l = -INFINITY;
  if (i < Lq) {
    const uint2 a = *reinterpret_cast<const uint2*>(o + i * o_ld_tok + (int64_t)h * o_ld_head + lane * 4);
    const uint2 b = *reinterpret_cast<const uint2*>(dout + i * do_ld_tok + (int64_t)h * do_ld_head + lane * 4);
    s = bf16lo(a.x) * bf16lo(b.x) + bf16hi(a.x) * bf16hi(b.x) + bf16lo(a.y) * bf16lo(b.y) + bf16hi(a.y) * bf16hi(b.y);
    s = -warp_sum(s);
    l = -lse[(int64_t)h * Lq + i] * 1.4426950408889634f;
  }
  if (lane == 0) {
    ws[(int64_t)h * Lp + i] = l;
    ws[(int64_t)(H + h) * Lp + i] = s;
  }
}

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// DKDV = true : resident = (K, V) of 128 keys, streamed = (Q, dO);  DKDV = false: resident = (Q, dO), streamed = (K, V).
// QUAD (dQ kernel only): all 16 softmax-gradient warps work on EVERY unit, four threads per row with 16 columns each, instead
// of two 8-warp groups that alternate units with two threads per row.  The stage between "S / dP computed" and "dS written"
// is latency-bound (tcgen05.ld -> exp2 -> pack -> tcgen05.st -> arrive ~ 860 cycles with 32 columns per thread), longer
// than the 768 tensor cycles of a unit, so with alternating groups the tensor pipe idles ~10-20 % (measured 81 %); with
// 16 columns per thread the stage is shorter than X(u+1) (512 cycles) and dS(u) is always ready when the pipe wants it.
template <bool DKDV, bool TS, bool QUAD = false>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                const __grid_constant__ CUtensorMap tmS0, const __grid_constant__ CUtensorMap tmS1, const AttnBwdParams p) {
  static_assert(DKDV || TS, "the dQ kernel only exists in the TS form");
  constexpr int STAGES = BwdSmem<TS>::STAGES;
  constexpr int SUBW = (DKDV && TS) ? 32 : 64;  // unit width = TMEM columns of one X / Y buffer
  constexpr int UPS = SUB / SUBW;               // units per TMA stage
  static_assert(!QUAD || !DKDV, "QUAD is the dQ kernel's variant");
  constexpr int CW = QUAD ? SUBW / 4 : SUBW / 2;   // columns per softmax-gradient thread (two or four threads per row)
  constexpr int XB = 0, YB = 2 * SUBW, ACC = 256, RES = DKDV ? 128 : 384;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sR0 = smem;                                   // resident operands (TS = false only)  [128][128] each
  uint8_t* sR1 = smem + BIG_BYTES;
  uint8_t* sS0 = smem + BwdSmem<TS>::RES_BYTES;          // streamed operand 0 (Q | K)   [stages][64][128]
  uint8_t* sS1 = sS0 + STAGES * SUB_BYTES;               // streamed operand 1 (dO | V)
  float* sLD = reinterpret_cast<float*>(sS1 + STAGES * SUB_BYTES);      // [stages][2][64]: -lse*log2e, -delta (DKDV only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sLD) + STAGES * 2 * SUB * 4);
  uint64_t* rfull = bars;                        // [1] resident operands ready
  uint64_t* sfull_ld = bars + 1;                 // [STAGES] streamed tiles landed
  uint64_t* sempty = bars + 1 + MAX_STAGES;      // [STAGES]
  uint64_t* xfull = bars + 1 + 2 * MAX_STAGES;   // [2] S/dP of buffer b computed
  uint64_t* pfull = xfull + 2;                   // [2] bf16 operands of buffer b written
  uint64_t* ofull = pfull + 2;                   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ofull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int r0 = blockIdx.x * BIG;                      // first resident row (key index | query index)
  const int L_stream = DKDV ? p.Lq : p.Lk;
  const int L_res = DKDV ? p.Lk : p.Lq;
  const int n_stage = (L_stream + SUB - 1) / SUB;
  const int n_unit = (L_stream + SUBW - 1) / SUBW;

  if (warp == W_TMA && lane == 0) {
    if (!TS) {
      tma_prefetch_desc(&tmR0);
      tma_prefetch_desc(&tmR1);
    }
    tma_prefetch_desc(&tmS0);
    tma_prefetch_desc(&tmS1);
    mbar_init(rfull, TS ? 16 : 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&sfull_ld[s], 1);
      mbar_init(&sempty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&xfull[b], 1);
      mbar_init(&pfull[b], QUAD ? 16 : 8);
    }
    mbar_init(ofull, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    // ------------------------------- TMA producer -------------------------------
    if (elect_one()) {
      if (!TS) {
        mbar_arrive_expect_tx(rfull, 2 * BIG_BYTES);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_load_3d(sR0 + c * 16384, &tmR0, rfull, c * 64, r0, head);
          tma_load_3d(sR1 + c * 16384, &tmR1, rfull, c * 64, r0, head);
        }
      }
      for (int i = 0; i < n_stage; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&sempty[s], ph ^ 1);
        mbar_arrive_expect_tx(&sfull_ld[s], 2 * SUB_BYTES + (DKDV ? 2 * SUB * 4 : 0));
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_load_3d(sS0 + s * SUB_BYTES + c * 8192, &tmS0, &sfull_ld[s], c * 64, i * SUB, head);
          tma_load_3d(sS1 + s * SUB_BYTES + c * 8192, &tmS1, &sfull_ld[s], c * 64, i * SUB, head);
        }
        if (DKDV) {
          bulk_load_1d(sLD + s * 2 * SUB, p.nlse2 + (int64_t)head * p.Lp + i * SUB, SUB * 4, &sfull_ld[s]);
          bulk_load_1d(sLD + s * 2 * SUB + SUB, p.ndelta + (int64_t)head * p.Lp + i * SUB, SUB * 4, &sfull_ld[s]);
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------- MMA issuer -------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_x = make_idesc_bf16(128, SUBW, 0, 0);    // [128 x SUBW] = resident x streamed^T (K-major)
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 128, 0, 1);   // [128 x 128] += A (tmem) x streamed (MN-major)
      const uint32_t r0a = smem_u32(sR0), r1a = smem_u32(sR1), s0a = smem_u32(sS0), s1a = smem_u32(sS1);
      constexpr uint32_t HI = sdesc_hi(1024);
      auto issue_x = [&](int u) {
        // X_b = R0 . S0_u^T ; Y_b = R1 . S1_u^T   (contraction over head_dim = 128, 8 k-steps)
        const int b = u & 1, st = (u / UPS) % STAGES, h = u % UPS;
        const uint32_t s0lo = sdesc_lo(s0a + st * SUB_BYTES + h * (SUBW * 128), 16);
        const uint32_t s1lo = sdesc_lo(s1a + st * SUB_BYTES + h * (SUBW * 128), 16);
        const uint32_t dx = tmem_base + XB + b * SUBW, dy = tmem_base + YB + b * SUBW;
        if (TS) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ts(dx, tmem_base + RES + k * 8, sdesc_join(s0lo + (k >> 2) * (8192 >> 4) + (k & 3) * 2, HI), idesc_x, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ts(dy, tmem_base + RES + 64 + k * 8, sdesc_join(s1lo + (k >> 2) * (8192 >> 4) + (k & 3) * 2, HI), idesc_x, k != 0 ? 1u : 0u);
        } else {
          const uint32_t r0lo = sdesc_lo(r0a, 16), r1lo = sdesc_lo(r1a, 16);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t offr = (k >> 2) * (16384 >> 4) + (k & 3) * 2, offs = (k >> 2) * (8192 >> 4) + (k & 3) * 2;
            umma_ss(dx, sdesc_join(r0lo + offr, HI), sdesc_join(s0lo + offs, HI), idesc_x, k != 0 ? 1u : 0u);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t offr = (k >> 2) * (16384 >> 4) + (k & 3) * 2, offs = (k >> 2) * (8192 >> 4) + (k & 3) * 2;
            umma_ss(dy, sdesc_join(r1lo + offr, HI), sdesc_join(s1lo + offs, HI), idesc_x, k != 0 ? 1u : 0u);
          }
        }
        umma_commit(&xfull[b]);
      };
      auto issue_acc = [&](int u, bool acc) {
        // contraction over the SUBW streamed rows of the unit; B = streamed tile as MN-major operand (two 64-wide d halves
        // 8192 B apart), one k-step = 16 rows = 2048 B.  The packed bf16 A columns of unit rows 16k..16k+15 were written
        // by thread-half (16k / CW) at its column (16k % CW) / 2.
        const int b = u & 1, st = (u / UPS) % STAGES, h = u % UPS;
        const uint32_t s0lo = sdesc_lo(s0a + st * SUB_BYTES + h * (SUBW * 128), 8192);
        const uint32_t s1lo = sdesc_lo(s1a + st * SUB_BYTES + h * (SUBW * 128), 8192);
        const uint32_t ax = tmem_base + XB + b * SUBW, ay = tmem_base + YB + b * SUBW;
#pragma unroll
        for (int k = 0; k < SUBW / 16; ++k) {
          const uint32_t acol = (k / (CW / 16)) * CW + (k % (CW / 16)) * 8;
          if (DKDV) umma_ts(tmem_base + ACC, ax + acol, sdesc_join(s1lo + k * (2048 >> 4), HI), idesc_acc, (acc || k != 0) ? 1u : 0u);   // dV += P^T dO
          else umma_ts(tmem_base + ACC, ay + acol, sdesc_join(s0lo + k * (2048 >> 4), HI), idesc_acc, (acc || k != 0) ? 1u : 0u);       // dQ += dS K
        }
        if (DKDV) {
#pragma unroll
          for (int k = 0; k < SUBW / 16; ++k) {
            const uint32_t acol = (k / (CW / 16)) * CW + (k % (CW / 16)) * 8;
            umma_ts(tmem_base + ACC + 128, ay + acol, sdesc_join(s0lo + k * (2048 >> 4), HI), idesc_acc, (acc || k != 0) ? 1u : 0u);      // dK += dS^T Q
          }
        }
        if (h == UPS - 1 || u == n_unit - 1) umma_commit(&sempty[st]);
      };
      mbar_wait(rfull, 0);
      mbar_wait(&sfull_ld[0], 0);
      tc_fence_after();
      issue_x(0);
      for (int u = 0; u < n_unit; ++u) {
        if (u + 1 < n_unit) {
          if ((u + 1) % UPS == 0) {
            const int i1 = (u + 1) / UPS;
            mbar_wait(&sfull_ld[i1 % STAGES], (i1 / STAGES) & 1);
            tc_fence_after();
          }
          issue_x(u + 1);
        }
        mbar_wait(&pfull[u & 1], (u >> 1) & 1);
        tc_fence_after();
        issue_acc(u, u > 0);
      }
      umma_commit(ofull);
    }
  } else {
    // ------------------------------- softmax-gradient warps + epilogue -------------------------------
    // 16 warps: buffer `wg` = warp / 8 handles units u with (u & 1) == wg; inside a buffer two warps share each TMEM lane
    // quadrant and split the unit's columns (`half`): two threads per row halve the latency of this stage, which is what
    // bounds the tensor pipe here (the MMAs of one buffer overlap the softmax-gradient of the other).  Packed bf16
    // results of a thread's fp32 columns [CW*half + 16c, +16) go to columns [CW*half + 8c, +8): they only alias fp32 columns
    // the same thread has already consumed, so the two halves never race.
    const int wg = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int quad = warp & 3;
    const int row = r0 + quad * 32 + lane;               // resident row of this thread (key | query)
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    if (TS) {
      // resident operands -> TMEM: packed bf16 = the raw row bytes; 4 threads per row, 32 elements (16 columns) each
      const int part = wg * 2 + half;
      uint32_t a[16], c[16];
      if (row < L_res) {
        const uint4* g0 = reinterpret_cast<const uint4*>(p.res0 + (int64_t)row * p.r0_ld_tok + (int64_t)head * p.r0_ld_head + part * 32);
        const uint4* g1 = reinterpret_cast<const uint4*>(p.res1 + (int64_t)row * p.r1_ld_tok + (int64_t)head * p.r1_ld_head + part * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u0 = ldg_nc_v4(g0 + j), u1 = ldg_nc_v4(g1 + j);
          a[4 * j] = u0.x; a[4 * j + 1] = u0.y; a[4 * j + 2] = u0.z; a[4 * j + 3] = u0.w;
          c[4 * j] = u1.x; c[4 * j + 1] = u1.y; c[4 * j + 2] = u1.z; c[4 * j + 3] = u1.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = c[j] = 0u;
      }
      tmem_st16(tmem_base + RES + part * 16 + lane_off, a);
      tmem_st16(tmem_base + RES + 64 + part * 16 + lane_off, c);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(rfull);
    }
    float2 nl_row = make_float2(0.f, 0.f), nd_row = make_float2(0.f, 0.f);
    if (!DKDV) {   // row < Lp always: the workspace is padded to whole 64-query blocks... but resident tiles are 128 rows
      const bool ok = row < p.Lp;
      const float l = ok ? p.nlse2[(int64_t)head * p.Lp + row] : -INFINITY;
      const float d = ok ? p.ndelta[(int64_t)head * p.Lp + row] : 0.f;
      nl_row = make_float2(l, l);
      nd_row = make_float2(d, d);
    }
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2);
    const int c0 = (QUAD ? wg * 2 + half : half) * CW;   // this thread's first fp32 column inside the unit
    for (int u = QUAD ? 0 : wg; u < n_unit; u += QUAD ? 1 : 2) {
      const int st = (u / UPS) % STAGES, h = u % UPS;
      const int xb = QUAD ? (u & 1) : wg;                // X / Y buffer of this unit
      const float* ld = sLD + st * 2 * SUB + h * SUBW + c0;      // -lse2 of this thread's columns; -delta is SUB floats further
      if (DKDV) mbar_wait(&sfull_ld[st], ((u / UPS) / STAGES) & 1);   // makes the bulk-copied -lse2 / -delta visible
      mbar_wait(&xfull[xb], (u >> 1) & 1);
      tc_fence_after();
      const uint32_t x_addr = tmem_base + XB + xb * SUBW + c0 + lane_off, y_addr = tmem_base + YB + xb * SUBW + c0 + lane_off;
#pragma unroll
      for (int c = 0; c < CW / 16; ++c) {
        uint32_t xs[16], ys[16];
        tmem_ld16(x_addr + c * 16, xs);
        tmem_ld16(y_addr + c * 16, ys);
        float4 l4[4], d4[4];
        if (DKDV) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            l4[j] = *reinterpret_cast<const float4*>(ld + c * 16 + 4 * j);
            d4[j] = *reinterpret_cast<const float4*>(ld + SUB + c * 16 + 4 * j);
          }
        }
        tmem_wait_ld();
        uint32_t pk[8], dk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float2 nl = nl_row, nd = nd_row;
          if (DKDV) {
            nl = (j & 2) ? make_float2(l4[j >> 2].z, l4[j >> 2].w) : make_float2(l4[j >> 2].x, l4[j >> 2].y);
            nd = (j & 2) ? make_float2(d4[j >> 2].z, d4[j >> 2].w) : make_float2(d4[j >> 2].x, d4[j >> 2].y);
          }
          const float2 t = ffma2(make_float2(__uint_as_float(xs[j]), __uint_as_float(xs[j + 1])), c2, nl);
          // all on MUFU: moving 25 % / 50 % of these to the FMA pipe (as the forward does) measured 23.40 / 25.04 ms vs 23.47 ms
          const float2 pr = make_float2(fast_exp2(t.x), fast_exp2(t.y));
          const float2 g = fmul2(pr, fadd2(make_float2(__uint_as_float(ys[j]), __uint_as_float(ys[j + 1])), nd));
          pk[j >> 1] = pack_bf16x2(pr.x, pr.y);
          dk[j >> 1] = pack_bf16x2(g.x, g.y);
        }
        if (DKDV) tmem_st8(x_addr + c * 8, pk);   // P^T (A operand of the dV MMA)
        tmem_st8(y_addr + c * 8, dk);             // dS^T | dS (A operand of the dK | dQ MMA)
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[xb]);
    }
    // ---- epilogue: 16 warps share the accumulator read-out ----
    mbar_wait(ofull, 0);
    tc_fence_after();
    const bool row_ok = row < L_res;
    // DKDV: buffer group 0 stores dV (ACC0), group 1 stores dK = scale * ACC1; each half takes 64 of the 128 columns.
    // DQ  : the four (group, half) pairs take 32 columns each of dQ = scale * ACC0.
    const int ncol = DKDV ? 64 : 32;
    const int col0 = DKDV ? half * 64 : (wg * 2 + half) * 32;
    const uint32_t a_addr = tmem_base + ACC + (DKDV ? wg * 128 : 0) + col0 + lane_off;
    __nv_bfloat16* dst = (DKDV && wg == 0) ? p.out1 + (int64_t)row * p.o1_ld_tok + (int64_t)head * p.o1_ld_head + col0
                                           : p.out0 + (int64_t)row * p.o0_ld_tok + (int64_t)head * p.o0_ld_head + col0;
    const float mul = (DKDV && wg == 0) ? 1.0f : p.scale;
#pragma unroll 1
    for (int c = 0; c < ncol / 16; ++c) {
      uint32_t o[16];
      tmem_ld16(a_addr + c * 16, o);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[j]) * mul, __uint_as_float(o[j + 1]) * mul);
          v.y = pack_bf16x2(__uint_as_float(o[j + 2]) * mul, __uint_as_float(o[j + 3]) * mul);
          v.z = pack_bf16x2(__uint_as_float(o[j + 4]) * mul, __uint_as_float(o[j + 5]) * mul);
          v.w = pack_bf16x2(__uint_as_float(o[j + 6]) * mul, __uint_as_float(o[j + 7]) * mul);
          *reinterpret_cast<uint4*>(dst + c * 16 + j) = v;
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace prfl

using namespace prfl;

extern "C" int64_t prfl_attn_bwd_ws_floats(int Lq, int H) { return 2LL * H * (((int64_t)Lq + 63) / 64 * 64); }

extern "C" int prfl_attn_bwd(const void* q, int64_t q_ld_tok, int64_t q_ld_head, const void* k, int64_t k_ld_tok,
                             int64_t k_ld_head, const void* v, int64_t v_ld_tok, int64_t v_ld_head, const void* o,
                             int64_t o_ld_tok, int64_t o_ld_head, const void* dout, int64_t do_ld_tok, int64_t do_ld_head,
                             const float* lse, float* ws, void* dq, int64_t dq_ld_tok, int64_t dq_ld_head, void* dk,
                             int64_t dk_ld_tok, int64_t dk_ld_head, void* dv, int64_t dv_ld_tok, int64_t dv_ld_head, int Lq,
                             int Lk, int H, float scale, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(Lq > 0 && Lk > 0 && H > 0 && lse && ws, PRFL_E_SHAPE, "attn_bwd: Lq=%d Lk=%d H=%d", Lq, Lk, H);
  const int64_t lds[] = {q_ld_tok, q_ld_head, k_ld_tok, k_ld_head, v_ld_tok, v_ld_head, o_ld_tok, o_ld_head, do_ld_tok, do_ld_head,
                         dq_ld_tok, dq_ld_head, dk_ld_tok, dk_ld_head, dv_ld_tok, dv_ld_head};
  for (int64_t l : lds) PRFL_REQUIRE(l % 8 == 0, PRFL_E_ALIGN, "attn_bwd: strides must be multiples of 8 elements");
  PRFL_REQUIRE(((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv) |
                 reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(q) |
                 reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(ws)) & 15) == 0,
               PRFL_E_ALIGN, "attn_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int Lp = (Lq + 63) / 64 * 64;
  {
    int64_t warps = (int64_t)Lp * H;
    attn_delta_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)o, o_ld_tok, o_ld_head,
                                                                   (const __nv_bfloat16*)dout, do_ld_tok, do_ld_head, lse, ws, Lq, Lp, H);
    count_launch();
    PRFL_LAUNCH_CHECK("attn_delta");
  }
  CUtensorMap tQb, tDOb, tKb, tVb, tQs, tDOs, tKs, tVs;
  int rc;
#define TM(m, ptr, L, ldt, ldh, rows)                                                                                   \
  rc = make_tmap_3d(&m, ptr, 128, (uint64_t)(L), (uint64_t)H, (uint64_t)(ldt) * 2, (uint64_t)(ldh) * 2, 64, rows, 1, 1); \
  if (rc != PRFL_OK) return rc;
  TM(tKb, k, Lk, k_ld_tok, k_ld_head, BIG)
  TM(tVb, v, Lk, v_ld_tok, v_ld_head, BIG)
  TM(tQs, q, Lq, q_ld_tok, q_ld_head, SUB)
  TM(tDOs, dout, Lq, do_ld_tok, do_ld_head, SUB)
  TM(tQb, q, Lq, q_ld_tok, q_ld_head, BIG)
  TM(tDOb, dout, Lq, do_ld_tok, do_ld_head, BIG)
  TM(tKs, k, Lk, k_ld_tok, k_ld_head, SUB)
  TM(tVs, v, Lk, v_ld_tok, v_ld_head, SUB)
#undef TM
  // dK/dV default: keys resident in shared memory, 64-query units (31.4 ms at L = 32 760 x 40 heads).  The all-TS form has
  // to shrink its units to 32 queries to fit K, V in tensor memory next to two accumulators; that halves the latency
  // budget of the softmax-gradient stage (512 tensor cycles) and measures slower (33.5 ms): PRFL_ATTN_BWD_DKDV=ts for A/B.
  static const bool dkdv_ts = [] { const char* e = getenv("PRFL_ATTN_BWD_DKDV"); return e && e[0] == 't'; }();
  static unsigned long long attr_mask = 0;      // per-device bit mask; forward and autograd threads may race: the call is idempotent
  if (device_needs_init(&attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<true>::BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<false>::BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<true>::BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem<true>::BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "attn_bwd: cudaFuncSetAttribute");
    device_mark_init(&attr_mask);
  }
  AttnBwdParams p;
  p.nlse2 = ws; p.ndelta = ws + (int64_t)H * Lp; p.Lq = Lq; p.Lk = Lk; p.Lp = Lp; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.out0 = (__nv_bfloat16*)dk; p.o0_ld_tok = dk_ld_tok; p.o0_ld_head = dk_ld_head;
  p.out1 = (__nv_bfloat16*)dv; p.o1_ld_tok = dv_ld_tok; p.o1_ld_head = dv_ld_head;
  p.res0 = (const __nv_bfloat16*)k; p.r0_ld_tok = k_ld_tok; p.r0_ld_head = k_ld_head;
  p.res1 = (const __nv_bfloat16*)v; p.r1_ld_tok = v_ld_tok; p.r1_ld_head = v_ld_head;
  if (dkdv_ts) attn_bwd_kernel<true, true><<<dim3((Lk + BIG - 1) / BIG, H), BWD_THREADS, BwdSmem<true>::BYTES, st>>>(tKb, tVb, tQs, tDOs, p);
  else attn_bwd_kernel<true, false><<<dim3((Lk + BIG - 1) / BIG, H), BWD_THREADS, BwdSmem<false>::BYTES, st>>>(tKb, tVb, tQs, tDOs, p);
  count_launch();
  PRFL_LAUNCH_CHECK("attn_bwd_dkdv");
  p.out0 = (__nv_bfloat16*)dq; p.o0_ld_tok = dq_ld_tok; p.o0_ld_head = dq_ld_head;
  p.out1 = nullptr; p.o1_ld_tok = p.o1_ld_head = 0;
  p.res0 = (const __nv_bfloat16*)q; p.r0_ld_tok = q_ld_tok; p.r0_ld_head = q_ld_head;
  p.res1 = (const __nv_bfloat16*)dout; p.r1_ld_tok = do_ld_tok; p.r1_ld_head = do_ld_head;
  // PRFL_ATTN_BWD_DQ=pair selects the round-1 form (two alternating 8-warp groups, two threads per row) for A/B runs
  static const bool dq_quad = [] { const char* e = getenv("PRFL_ATTN_BWD_DQ"); return !(e && e[0] == 'p'); }();
  if (dq_quad) attn_bwd_kernel<false, true, true><<<dim3((Lq + BIG - 1) / BIG, H), BWD_THREADS, BwdSmem<true>::BYTES, st>>>(tQb, tDOb, tKs, tVs, p);
  else attn_bwd_kernel<false, true><<<dim3((Lq + BIG - 1) / BIG, H), BWD_THREADS, BwdSmem<true>::BYTES, st>>>(tQb, tDOb, tKs, tVs, p);
  count_launch();
  PRFL_LAUNCH_CHECK("attn_bwd_dq");
  return PRFL_OK;
}
