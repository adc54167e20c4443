// Flash-attention backward for sm_100a (bf16, head_dim 128, non-causal, ragged tails), tcgen05 + TMEM + TMA.
// Two kernels, both with the forward kernel's skeleton (TMA warp, MMA warp, 2 x 256 softmax-gradient threads — two
// threads per row — that ping-pong over 64-wide sub-tiles), no atomics, deterministic:
//
//   attn_bwd_dkdv_kernel : CTA = one head x 128 keys (K_j, V_j resident in smem), loop over 64-query sub-tiles
//        S^T = K Q^T, dP^T = V dO^T           (SS MMA, M = keys, N = 64 queries)       -> TMEM
//        P^T = exp2(S^T c - lse), dS^T = P^T (dP^T - delta)   (thread = key row)       -> bf16 back into TMEM
//        dV += P^T dO, dK += dS^T Q           (TS MMA, A from TMEM, B = dO / Q as MN-major smem operands)
//   attn_bwd_dq_kernel   : CTA = one head x 128 queries (Q, dO resident), loop over 64-key sub-tiles
//        S = Q K^T, dP = dO V^T               (SS MMA, M = queries, N = 64 keys)       -> TMEM
//        dS = P (dP - delta) scale            (thread = query row)                      -> bf16 back into TMEM
//        dQ += dS K                           (TS MMA, B = K as MN-major smem operand)
//
// The split costs 7 instead of 5 tile-GEMMs per (q, k) tile pair (S and dP are recomputed in the dQ kernel) but needs
// no cross-CTA reduction of dQ.  delta = rowsum(dO * O) comes from a small pre-pass.
#include <math.h>

#include "common.cuh"

namespace prfl {

constexpr int BWD_THREADS = 576;            // 16 softmax-gradient warps (2 threads per row) + TMA warp + MMA warp
constexpr int W_TMA = 16, W_MMA = 17;
constexpr int BIG = 128;                     // resident tile rows
constexpr int SUB = 64;                      // streamed sub-tile rows
constexpr int BIG_BYTES = BIG * 128 * 2;     // 32 KB
constexpr int SUB_BYTES = SUB * 128 * 2;     // 16 KB
constexpr int BWD_STAGES = 3;
constexpr int BWD_SMEM = 2 * BIG_BYTES + BWD_STAGES * 2 * SUB_BYTES + 2 * 2 * SUB * 4 + 256 + 1024;

struct AttnBwdParams {
  const float* lse;     // [H, Lq]
  const float* delta;   // [H, Lq]
  __nv_bfloat16* out0;  // dkdv: dK ; dq: dQ
  __nv_bfloat16* out1;  // dkdv: dV
  int64_t o0_ld_tok, o0_ld_head, o1_ld_tok, o1_ld_head;
  int Lq, Lk;
  float scale, scale_log2;
};

// delta[h, i] = sum_d dO[i,h,d] * O[i,h,d]   — one warp per (token, head)
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t o_ld_tok, int64_t o_ld_head,
                                  const __nv_bfloat16* __restrict__ dout, int64_t do_ld_tok, int64_t do_ld_head,
                                  float* __restrict__ delta, int Lq, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)Lq * H) return;
  const int h = (int)(w % H);
  const int64_t i = w / H;
  const uint2 a = *reinterpret_cast<const uint2*>(o + i * o_ld_tok + (int64_t)h * o_ld_head + lane * 4);
  const uint2 b = *reinterpret_cast<const uint2*>(dout + i * do_ld_tok + (int64_t)h * do_ld_head + lane * 4);
  float s = bf16lo(a.x) * bf16lo(b.x) + bf16hi(a.x) * bf16hi(b.x) + bf16lo(a.y) * bf16lo(b.y) + bf16hi(a.y) * bf16hi(b.y);
  s = warp_sum(s);
  if (lane == 0) delta[(int64_t)h * Lq + i] = s;
}

// Shared skeleton.  DKDV = true : resident = (K, V) of 128 keys, streamed = (Q, dO) sub-tiles of 64 queries.
//                   DKDV = false: resident = (Q, dO) of 128 queries, streamed = (K, V) sub-tiles of 64 keys.
// TMEM columns: X0 [0,64) X1 [64,128) (S or S^T, double buffered), Y0 [128,192) Y1 [192,256) (dP or dP^T),
//               ACC0 [256,384), ACC1 [384,512).  bf16 operands written by the softmax threads alias X / Y.
template <bool DKDV>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                const __grid_constant__ CUtensorMap tmS0, const __grid_constant__ CUtensorMap tmS1, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sR0 = smem;                                   // resident operand 0 (K | Q)   [128][128]
  uint8_t* sR1 = smem + BIG_BYTES;                       // resident operand 1 (V | dO)
  uint8_t* sS0 = smem + 2 * BIG_BYTES;                   // streamed operand 0 (Q | K)   [stages][64][128]
  uint8_t* sS1 = sS0 + BWD_STAGES * SUB_BYTES;           // streamed operand 1 (dO | V)
  float* sLD = reinterpret_cast<float*>(sS1 + BWD_STAGES * SUB_BYTES);  // [2 bufs][2][64]: lse*log2e, delta (DKDV only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sLD) + 2 * 2 * SUB * 4);
  uint64_t* rfull = bars;                  // [1]
  uint64_t* sfull_ld = bars + 1;           // [STAGES] streamed tiles landed
  uint64_t* sempty = bars + 1 + BWD_STAGES;      // [STAGES]
  uint64_t* xfull = bars + 1 + 2 * BWD_STAGES;   // [2] S/dP of buffer b computed
  uint64_t* pfull = xfull + 2;             // [2] bf16 operands of buffer b written
  uint64_t* ofull = pfull + 2;             // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ofull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int r0 = blockIdx.x * BIG;                      // first resident row (key index | query index)
  const int L_stream = DKDV ? p.Lq : p.Lk;
  const int n_sub = (L_stream + SUB - 1) / SUB;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmR0);
    tma_prefetch_desc(&tmR1);
    tma_prefetch_desc(&tmS0);
    tma_prefetch_desc(&tmS1);
    mbar_init(rfull, 1);
    for (int s = 0; s < BWD_STAGES; ++s) {
      mbar_init(&sfull_ld[s], 1);
      mbar_init(&sempty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&xfull[b], 1);
      mbar_init(&pfull[b], 8);
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
      mbar_arrive_expect_tx(rfull, 2 * BIG_BYTES);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tma_load_3d(sR0 + c * 16384, &tmR0, rfull, c * 64, r0, head);
        tma_load_3d(sR1 + c * 16384, &tmR1, rfull, c * 64, r0, head);
      }
      for (int i = 0; i < n_sub; ++i) {
        const int s = i % BWD_STAGES;
        const uint32_t ph = (i / BWD_STAGES) & 1;
        mbar_wait(&sempty[s], ph ^ 1);
        mbar_arrive_expect_tx(&sfull_ld[s], 2 * SUB_BYTES);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_load_3d(sS0 + s * SUB_BYTES + c * 8192, &tmS0, &sfull_ld[s], c * 64, i * SUB, head);
          tma_load_3d(sS1 + s * SUB_BYTES + c * 8192, &tmS1, &sfull_ld[s], c * 64, i * SUB, head);
        }
      }
    }
  } else if (warp == W_MMA) {
    // ------------------------------- MMA issuer -------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_x = make_idesc_bf16(128, SUB, 0, 0);   // [128 x 64] = R (K-major) x S^T (K-major)
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 128, 0, 1); // [128 x 128] += tmem A x MN-major B
      const uint32_t r0a = smem_u32(sR0), r1a = smem_u32(sR1), s0a = smem_u32(sS0), s1a = smem_u32(sS1);
      constexpr uint32_t HI = sdesc_hi(1024);
      auto issue_x = [&](int b, int st) {
        // X_b = R0 . S0^T ; Y_b = R1 . S1^T   (contraction over head_dim = 128, 8 k-steps)
        const uint32_t r0lo = sdesc_lo(r0a, 16), r1lo = sdesc_lo(r1a, 16);
        const uint32_t s0lo = sdesc_lo(s0a + st * SUB_BYTES, 16), s1lo = sdesc_lo(s1a + st * SUB_BYTES, 16);
        const uint32_t dx = tmem_base + b * 64, dy = tmem_base + 128 + b * 64;
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
        umma_commit(&xfull[b]);
      };
      auto issue_acc = [&](int b, int st, bool acc) {
        // contraction over the 64 streamed rows (4 k-steps); B = streamed tile as MN-major operand (LBO = 8192).
        // packed bf16 A columns of streamed rows 16k..16k+15 live at 32*(k>>1) + 8*(k&1) of the buffer
        const uint32_t s0lo = sdesc_lo(s0a + st * SUB_BYTES, 8192), s1lo = sdesc_lo(s1a + st * SUB_BYTES, 8192);
        const uint32_t ax = tmem_base + b * 64, ay = tmem_base + 128 + b * 64;
        if (DKDV) {
          // dV (ACC0) += P^T_b . dO ; dK (ACC1) += dS^T_b . Q
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ts(tmem_base + 256, ax + (k >> 1) * 32 + (k & 1) * 8, sdesc_join(s1lo + k * (2048 >> 4), HI), idesc_acc,
                    (acc || k != 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ts(tmem_base + 384, ay + (k >> 1) * 32 + (k & 1) * 8, sdesc_join(s0lo + k * (2048 >> 4), HI), idesc_acc,
                    (acc || k != 0) ? 1u : 0u);
        } else {
          // dQ (ACC0) += dS_b . K
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ts(tmem_base + 256, ay + (k >> 1) * 32 + (k & 1) * 8, sdesc_join(s0lo + k * (2048 >> 4), HI), idesc_acc,
                    (acc || k != 0) ? 1u : 0u);
        }
      };
      mbar_wait(rfull, 0);
      mbar_wait(&sfull_ld[0], 0);
      tc_fence_after();
      issue_x(0, 0);
      for (int i = 0; i < n_sub; ++i) {
        const int b = i & 1, st = i % BWD_STAGES;
        if (i + 1 < n_sub) {
          const int st1 = (i + 1) % BWD_STAGES;
          mbar_wait(&sfull_ld[st1], ((i + 1) / BWD_STAGES) & 1);
          tc_fence_after();
          issue_x(b ^ 1, st1);
        }
        mbar_wait(&pfull[b], (i >> 1) & 1);
        tc_fence_after();
        issue_acc(b, st, i > 0);
        umma_commit(&sempty[st]);
      }
      umma_commit(ofull);
    }
  } else {
    // ------------------------------- softmax-gradient warps + epilogue -------------------------------
    // 16 warps: buffer `wg` = warp / 8 handles sub-tiles i with (i & 1) == wg; inside a buffer two warps share each TMEM
    // lane quadrant and split the 64 streamed columns (`half`): two threads per row halve the latency of this stage,
    // which is what bounds the tensor pipe here (the MMAs of one buffer overlap the softmax-gradient of the other).
    // Packed bf16 results of fp32 columns [32h + 16c, +16) go to columns [32h + 8c, +8): they only alias fp32 columns the
    // same thread has already consumed, so the two halves never race.
    const int wg = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int quad = warp & 3;
    const int row = r0 + quad * 32 + lane;               // resident row of this thread (key | query)
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float lse2_row = 0.f, delta_row = 0.f;
    if (!DKDV) {
      const bool ok = row < p.Lq;
      lse2_row = ok ? p.lse[(int64_t)head * p.Lq + row] * 1.4426950408889634f : INFINITY;
      delta_row = ok ? p.delta[(int64_t)head * p.Lq + row] : 0.f;
    }
    float* ld = sLD + wg * 2 * SUB;
    const int tid_wg = threadIdx.x & 255;
    const int c0 = half * 32;                            // this thread's 32 fp32 columns of every sub-tile
    for (int i = wg; i < n_sub; i += 2) {
      if (DKDV) {
        // stage lse*log2e and delta of the 64 streamed queries (guarding the ragged tail)
        named_bar_sync(1 + wg, 256);                       // previous sub-tile's readers are done
        if (tid_wg < 128) {
          const int qi = i * SUB + (tid_wg & 63);
          const bool ok = qi < p.Lq;
          if (tid_wg < 64) ld[tid_wg] = ok ? p.lse[(int64_t)head * p.Lq + qi] * 1.4426950408889634f : INFINITY;
          else ld[tid_wg] = ok ? p.delta[(int64_t)head * p.Lq + qi] : 0.f;
        }
        named_bar_sync(1 + wg, 256);
      }
      mbar_wait(&xfull[wg], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t x_addr = tmem_base + wg * 64 + c0 + lane_off, y_addr = tmem_base + 128 + wg * 64 + c0 + lane_off;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t xs[16], ys[16];
        tmem_ld16(x_addr + c * 16, xs);
        tmem_ld16(y_addr + c * 16, ys);
        tmem_wait_ld();
        uint32_t pk[8], dk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float l0, l1, d0, d1;
          if (DKDV) {
            const float2 l2 = *reinterpret_cast<const float2*>(ld + c0 + c * 16 + j);
            const float2 d2 = *reinterpret_cast<const float2*>(ld + SUB + c0 + c * 16 + j);
            l0 = l2.x; l1 = l2.y; d0 = d2.x; d1 = d2.y;
          } else {
            l0 = l1 = lse2_row; d0 = d1 = delta_row;
          }
          const float p0 = fast_exp2(fmaf(__uint_as_float(xs[j]), p.scale_log2, -l0));
          const float p1 = fast_exp2(fmaf(__uint_as_float(xs[j + 1]), p.scale_log2, -l1));
          float g0 = p0 * (__uint_as_float(ys[j]) - d0);
          float g1 = p1 * (__uint_as_float(ys[j + 1]) - d1);
          if (!DKDV) { g0 *= p.scale; g1 *= p.scale; }
          pk[j >> 1] = pack_bf16x2(p0, p1);
          dk[j >> 1] = pack_bf16x2(g0, g1);
        }
        if (DKDV) tmem_st8(x_addr + c * 8, pk);   // P^T (A operand of the dV MMA)
        tmem_st8(y_addr + c * 8, dk);             // dS^T | dS (A operand of the dK | dQ MMA)
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[wg]);
    }
    // ---- epilogue: 16 warps share the accumulator read-out ----
    mbar_wait(ofull, 0);
    tc_fence_after();
    const int L_res = DKDV ? p.Lk : p.Lq;
    const bool row_ok = row < L_res;
    // DKDV: buffer group 0 stores dV (ACC0), group 1 stores dK = scale * ACC1; each half takes 64 of the 128 columns.
    // DQ  : the four (group, half) pairs take 32 columns each of dQ (ACC0).
    const int ncol = DKDV ? 64 : 32;
    const int col0 = DKDV ? half * 64 : (wg * 2 + half) * 32;
    const uint32_t a_addr = tmem_base + 256 + (DKDV ? wg * 128 : 0) + col0 + lane_off;
    __nv_bfloat16* dst = (DKDV && wg == 0) ? p.out1 + (int64_t)row * p.o1_ld_tok + (int64_t)head * p.o1_ld_head + col0
                                           : p.out0 + (int64_t)row * p.o0_ld_tok + (int64_t)head * p.o0_ld_head + col0;
    const float mul = (DKDV && wg == 1) ? p.scale : 1.0f;
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

extern "C" int prfl_attn_bwd(const void* q, int64_t q_ld_tok, int64_t q_ld_head, const void* k, int64_t k_ld_tok,
                             int64_t k_ld_head, const void* v, int64_t v_ld_tok, int64_t v_ld_head, const void* o,
                             int64_t o_ld_tok, int64_t o_ld_head, const void* dout, int64_t do_ld_tok, int64_t do_ld_head,
                             const float* lse, float* delta, void* dq, int64_t dq_ld_tok, int64_t dq_ld_head, void* dk,
                             int64_t dk_ld_tok, int64_t dk_ld_head, void* dv, int64_t dv_ld_tok, int64_t dv_ld_head, int Lq,
                             int Lk, int H, float scale, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(Lq > 0 && Lk > 0 && H > 0 && lse && delta, PRFL_E_SHAPE, "attn_bwd: Lq=%d Lk=%d H=%d", Lq, Lk, H);
  const int64_t lds[] = {q_ld_tok, q_ld_head, k_ld_tok, k_ld_head, v_ld_tok, v_ld_head, o_ld_tok, o_ld_head, do_ld_tok, do_ld_head,
                         dq_ld_tok, dq_ld_head, dk_ld_tok, dk_ld_head, dv_ld_tok, dv_ld_head};
  for (int64_t l : lds) PRFL_REQUIRE(l % 8 == 0, PRFL_E_ALIGN, "attn_bwd: strides must be multiples of 8 elements");
  PRFL_REQUIRE(((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv) |
                 reinterpret_cast<uintptr_t>(o) | reinterpret_cast<uintptr_t>(dout)) & 15) == 0,
               PRFL_E_ALIGN, "attn_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  {
    int64_t warps = (int64_t)Lq * H;
    attn_delta_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)o, o_ld_tok, o_ld_head,
                                                                   (const __nv_bfloat16*)dout, do_ld_tok, do_ld_head, delta, Lq, H);
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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "attn_bwd: cudaFuncSetAttribute");
    attr_set = true;
  }
  AttnBwdParams p;
  p.lse = lse; p.delta = delta; p.Lq = Lq; p.Lk = Lk; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;
  p.out0 = (__nv_bfloat16*)dk; p.o0_ld_tok = dk_ld_tok; p.o0_ld_head = dk_ld_head;
  p.out1 = (__nv_bfloat16*)dv; p.o1_ld_tok = dv_ld_tok; p.o1_ld_head = dv_ld_head;
  attn_bwd_kernel<true><<<dim3((Lk + BIG - 1) / BIG, H), BWD_THREADS, BWD_SMEM, st>>>(tKb, tVb, tQs, tDOs, p);
  count_launch();
  PRFL_LAUNCH_CHECK("attn_bwd_dkdv");
  p.out0 = (__nv_bfloat16*)dq; p.o0_ld_tok = dq_ld_tok; p.o0_ld_head = dq_ld_head;
  p.out1 = nullptr; p.o1_ld_tok = p.o1_ld_head = 0;
  attn_bwd_kernel<false><<<dim3((Lq + BIG - 1) / BIG, H), BWD_THREADS, BWD_SMEM, st>>>(tQb, tDOb, tKs, tVs, p);
  count_launch();
  PRFL_LAUNCH_CHECK("attn_bwd_dq");
  return PRFL_OK;
}
