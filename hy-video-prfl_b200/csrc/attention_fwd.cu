// Flash-attention forward for sm_100a: bf16, head_dim 128, non-causal, ragged tails.
//   S = Q K^T   : tcgen05.mma SS (Q, K in 128B-swizzled smem via TMA), fp32 S in TMEM
//   P = softmax : 2 x 128 softmax threads (one row each) read S from TMEM, write bf16 P back over S
//   O += P V    : tcgen05.mma TS (A = P from TMEM, B = V as an MN-major smem operand), fp32 O in TMEM
// One CTA = one head x 256 queries (two 128-row Q tiles that ping-pong: while the tensor core works on
// tile A's MMAs the softmax warps of tile B run, as in FlashAttention-4).  O is rescaled lazily: the
// running max used for exp2 only moves when the true max grew by more than 8 (log2 units), which is
// exact because the row sum is kept relative to the same stale max.
// Warps: 0-7 softmax of Q tile 0, 8-15 softmax of Q tile 1 (two threads per row), 16 TMA producer, 17 MMA issuer.
// TMEM: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_i (bf16) aliases columns [0,32) and [64,96) of S_i.
// Roofline: tensor pipe, 4*Lq*Lk*128 flops per head.
//
// Wave quantisation.  A unit (one head x 256 queries over all keys) is one CTA and all units cost the same, so a grid of
// U units on 148 SMs takes ceil(U / 148) waves: 640 units (L = 32 760, 5 heads per rank at 8 GPUs) run 5 waves with the
// last one 32 % full.  When the remainder is at most half a wave, the launcher runs the full waves as usual and the
// remaining units as a second launch split S = 2..4 ways along the KEY axis (flash-decoding style): each piece writes its
// normalised fp32 partial output and log2-domain LSE to a caller-provided workspace and `attn_combine_kernel` merges
// them (and performs the peer stores of the fused Ulysses epilogue).  640 units: 4 waves + 1/3 wave instead of 5.
#include <math.h>

#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace prfl {

constexpr int ATT_THREADS = 576;   // 16 softmax warps (2 Q tiles x 2 column halves x 4 lane quadrants) + TMA warp + MMA warp
constexpr int AW_TMA = 16, AW_MMA = 17;
constexpr int QT = 128;       // rows per Q tile
constexpr int KT = 128;       // keys per KV tile
constexpr int HD = 128;
constexpr int TILE_BYTES = 128 * 128 * 2;  // 32 KB: any 128 x 128 bf16 tile (two 64-wide TMA boxes)
constexpr int KV_STAGES = 2;
constexpr int ATT_XCH = 2 * 2 * 4 * 128 * 4;   // [slot][Q tile][column part][row] floats exchanged between the threads of a row
constexpr int ATT_SMEM = 2 * TILE_BYTES + 2 * KV_STAGES * TILE_BYTES + ATT_XCH + 256 + 1024;

struct AttnFwdParams {
  __nv_bfloat16* o;
  int64_t o_ld_tok, o_ld_head;
  float* lse;
  int Lq, Lk;
  float scale_log2;
  // Ulysses output exchange fused into the epilogue: row i belongs to rank i / L_loc and is stored straight into that
  // rank's [L_loc, H_total, 128] buffer over NVLink (peer-mapped pointers), at head head_off + h.  n_peer == 0: off.
  __nv_bfloat16* o_peer[8];
  int n_peer, L_loc, head_off;
  // unit addressing + key-axis split (see "Wave quantisation" above)
  int nq;            // 256-query blocks per head
  int unit0;         // first unit of this launch; unit = unit0 + blockIdx.x = head * nq + q block
  int n_split;       // pieces per unit along the key axis (blockIdx.y); 1 = write the final output directly
  float* part_o;     // [blockIdx.x][n_split][256][128] fp32, normalised partial outputs
  float* part_lse;   // [blockIdx.x][n_split][256]      log2-domain LSE of each partial
  int bar_wide;      // QUAD softmax: 1 = one 512-thread barrier per (tile, KV tile) for the row-max exchange (A/B), 0 = one
                     // 128-thread barrier per TMEM lane quadrant (the 4 warps that share rows sit on the same SM sub-partition)
};

// FMA_MASK: which of the 8 (i = 0, 4, ..., 28) second pairs of each 32-column chunk take the FMA-pipe exp2 (bit i/4):
// 0xAA = 4 of 16 pairs (25 %), 0xEE = 37.5 %, 0xFF = 50 %, 0 = all on MUFU.
// QUAD: all 16 softmax warps work on Q tile 0, then Q tile 1, of every KV tile — four threads per row with 32 keys / 32
// output columns each — instead of one 8-warp group per Q tile with two threads per row.  The S -> softmax -> P stage is
// what the tensor pipe waits on (P_t(j) gates P.V of tile t and, behind it, S_t(j+1)); with 64 columns per thread it
// lasts ~1 500 cycles against the 1 024 tensor cycles available to hide it.  Halving the columns per thread shortens the
// serial part (tcgen05.ld latency, pack, tcgen05.st, fences) while the MUFU work per tile stays the same.
// Measured and dropped (profiles/r02_attn_fwd_quad_and_prefetch_experiment.patch): fetching S in 16-column chunks one chunk
// ahead of the exponentials (tcgen05.ld latency hidden behind math / exchange / P store) — 20.0 ms against 17.8 ms for this
// form at L = 32 760 x 40 heads: the extra tcgen05.ld / wait / probe instructions cost more than the latency they hide.
template <int FMA_MASK, bool QUAD = false>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                                  // [2][32 KB]
  uint8_t* sK = smem + 2 * TILE_BYTES;                 // [KV_STAGES][32 KB]
  uint8_t* sV = sK + KV_STAGES * TILE_BYTES;           // [KV_STAGES][32 KB]
  float* sX = reinterpret_cast<float*>(sV + KV_STAGES * TILE_BYTES);   // [2 slots][2 tiles][2 halves][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sX) + ATT_XCH);
  uint64_t* qfull = bars;          // [1]
  uint64_t* kfull = bars + 1;      // [2]
  uint64_t* kempty = bars + 3;     // [2]
  uint64_t* vfull = bars + 5;      // [2]
  uint64_t* vempty = bars + 7;     // [2]
  uint64_t* sfull = bars + 9;      // [2]  S_i written by the tensor core
  uint64_t* pfull = bars + 11;     // [2]  P_i written by the softmax warps
  uint64_t* ofull = bars + 13;     // [1]  all MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = p.unit0 + blockIdx.x;
  const int head = unit / p.nq;
  const int q0 = (unit % p.nq) * (2 * QT);
  const int n_kv_all = (p.Lk + KT - 1) / KT;
  const int kv_per = (n_kv_all + p.n_split - 1) / p.n_split;
  const int kv0 = blockIdx.y * kv_per;                         // first KV tile of this piece
  const int n_kv = min(kv_per, n_kv_all - kv0);                // >= 1 (the launcher keeps n_split <= n_kv_all / 2)

  if (warp == AW_TMA && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(qfull, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kfull[s], 1);
      mbar_init(&kempty[s], 1);
      mbar_init(&vfull[s], 1);
      mbar_init(&vempty[s], 1);
      mbar_init(&sfull[s], 1);
      mbar_init(&pfull[s], QUAD ? 16 : 8);
    }
    mbar_init(ofull, 1);
    fence_barrier_init();
  }
  if (warp == AW_MMA) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == AW_TMA) {
    // ------------------------------- TMA producer -------------------------------
    if (elect_one()) {
      mbar_arrive_expect_tx(qfull, 2 * TILE_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int c = 0; c < 2; ++c) tma_load_3d(sQ + t * TILE_BYTES + c * 16384, &tmQ, qfull, c * 64, q0 + t * QT, head);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&kempty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kfull[s], TILE_BYTES);
#pragma unroll
        for (int c = 0; c < 2; ++c) tma_load_3d(sK + s * TILE_BYTES + c * 16384, &tmK, &kfull[s], c * 64, (kv0 + j) * KT, head);
        mbar_wait(&vempty[s], ph ^ 1);
        mbar_arrive_expect_tx(&vfull[s], TILE_BYTES);
#pragma unroll
        for (int c = 0; c < 2; ++c) tma_load_3d(sV + s * TILE_BYTES + c * 16384, &tmV, &vfull[s], c * 64, (kv0 + j) * KT, head);
      }
    }
  } else if (warp == AW_MMA) {
    // ------------------------------- MMA issuer -------------------------------
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // A = P (TMEM, K-major), B = V (MN-major)
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      constexpr uint32_t HI = sdesc_hi(1024);
      auto issue_qk = [&](int t, int ks) {
        const uint32_t qlo = sdesc_lo(q_addr + t * TILE_BYTES, 16), klo = sdesc_lo(k_addr + ks * TILE_BYTES, 16);
        const uint32_t d = tmem_base + t * 128;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t off = (k >> 2) * (16384 >> 4) + (k & 3) * (32 >> 4);   // in 16-byte units
          umma_ss(d, sdesc_join(qlo + off, HI), sdesc_join(klo + off, HI), idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&sfull[t]);
      };
      auto issue_pv = [&](int t, int vs, bool acc) {
        const uint32_t vlo = sdesc_lo(v_addr + vs * TILE_BYTES, 16384);
        const uint32_t d = tmem_base + 256 + t * 128, a = tmem_base + t * 128;
        // packed bf16 P of keys [16k, 16k + 16): written by the thread that owns those keys at the start of its own fp32 columns
#pragma unroll
        for (int k = 0; k < KT / 16; ++k)
          umma_ts(d, a + (QUAD ? (k >> 1) * 32 + (k & 1) * 8 : (k >> 2) * 64 + (k & 3) * 8), sdesc_join(vlo + k * (2048 >> 4), HI),
                  idesc_pv, (acc || k != 0) ? 1u : 0u);
      };
      mbar_wait(qfull, 0);
      mbar_wait(&kfull[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      umma_commit(&kempty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        const bool more = j + 1 < n_kv;
        const int s1 = (j + 1) % KV_STAGES;
        const uint32_t ph1 = ((j + 1) / KV_STAGES) & 1;
        mbar_wait(&vfull[s], ph);
        mbar_wait(&pfull[0], j & 1);
        tc_fence_after();
        issue_pv(0, s, j > 0);
        if (more) {
          mbar_wait(&kfull[s1], ph1);
          tc_fence_after();
          issue_qk(0, s1);
        }
        mbar_wait(&pfull[1], j & 1);
        tc_fence_after();
        issue_pv(1, s, j > 0);
        umma_commit(&vempty[s]);
        if (more) {
          issue_qk(1, s1);
          umma_commit(&kempty[s1]);
        }
      }
      umma_commit(ofull);
    }
  } else if (QUAD) {
    // ------------------------------- softmax + epilogue, four threads per row -------------------------------
    const int part = warp >> 2;              // which 32 keys / 32 output columns
    const int quad = warp & 3;               // TMEM lane quadrant
    const int r = quad * 32 + lane;          // row inside a Q tile
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float m_used[2] = {-INFINITY, -INFINITY};   // per Q tile: max the exponentials are relative to (identical in the 4 threads of a row)
    float l_sum[2] = {0.f, 0.f};                // per Q tile: this thread's share of the row sum
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    auto xslot = [&](int slot, int t, int pt) { return sX + ((slot * 2 + t) * 4 + pt) * 128 + r; };
    auto row_sync = [&]() {
      if (p.bar_wide) named_bar_sync(1, 512);
      else named_bar_sync(2 + quad, 128);     // warps quad, quad + 4, quad + 8, quad + 12: the four parts of the same 32 rows
    };
    auto row_max4 = [&](int slot, int t, float mine) {
      *xslot(slot, t, part) = mine;
      row_sync();
      return fmaxf(fmaxf(*xslot(slot, t, 0), *xslot(slot, t, 1)), fmaxf(*xslot(slot, t, 2), *xslot(slot, t, 3)));
    };
    for (int j = 0; j < n_kv; ++j) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t s_addr = tmem_base + t * 128 + part * 32 + lane_off;
        const uint32_t o_addr = tmem_base + 256 + t * 128 + part * 32 + lane_off;
        mbar_wait(&sfull[t], j & 1);
        tc_fence_after();
        const int valid = p.Lk - (kv0 + j) * KT - part * 32;   // my columns >= valid are padding (only the last tile is ragged)
        uint32_t pk[16];
        float tile_sum = 0.f, tile_max = -INFINITY;
        auto exp_pass = [&](const float m_ref) {
          const float2 neg_m2 = make_float2(-m_ref, -m_ref);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
          float mx0 = -INFINITY, mx1 = -INFINITY;
          uint32_t rr[32];
          tmem_ld32(s_addr, rr);
          tmem_wait_ld();
          if (valid < 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i >= valid) rr[i] = 0xff800000u;  // -inf
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmax3(mx0, __uint_as_float(rr[i]), __uint_as_float(rr[i + 1]));
            mx1 = fmax3(mx1, __uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3]));
            float2 xa = ffma2(make_float2(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1])), sc2, neg_m2);
            float2 xb = ffma2(make_float2(__uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3])), sc2, neg_m2);
            float2 pa, pb;
            pa = make_float2(fast_exp2(xa.x), fast_exp2(xa.y));
            if ((FMA_MASK >> (i >> 2)) & 1) pb = exp2_fma2(xb);   // FMA pipe instead of MUFU (compile-time pattern)
            else pb = make_float2(fast_exp2(xb.x), fast_exp2(xb.y));
            acc0 = fadd2(acc0, pa);
            acc1 = fadd2(acc1, pb);
            pk[i >> 1] = pack_bf16x2(pa.x, pa.y);
            pk[(i >> 1) + 1] = pack_bf16x2(pb.x, pb.y);
          }
          tile_sum = (acc0.x + acc0.y) + (acc1.x + acc1.y);
          tile_max = fmaxf(mx0, mx1);
        };
        if (j == 0) {
          // first tile: the reference max must be the true row max over all 128 keys (one extra read of S from TMEM)
          float mx = -INFINITY;
          {
            uint32_t rr[32];
            tmem_ld32(s_addr, rr);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < valid) mx = fmaxf(mx, __uint_as_float(rr[i]));
          }
          m_used[t] = row_max4(0, t, mx) * p.scale_log2;
          exp_pass(m_used[t]);
        } else {
          // optimistic: exponentiate against the stale max; exact unless the row max grew by more than 8 (log2 units)
          exp_pass(m_used[t]);
          const float m_new = fmaxf(m_used[t], row_max4(j & 1, t, tile_max) * p.scale_log2);
          // same rows in the same lanes of the three partner warps => all four take the same (warp-uniform) branch
          if (__any_sync(0xffffffffu, m_new > m_used[t] + 8.0f)) {
            const float alpha = fast_exp2(m_used[t] - m_new);
            l_sum[t] *= alpha;
            m_used[t] = m_new;
            uint32_t o[32];
            tmem_ld32(o_addr, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(o_addr, o);
            exp_pass(m_used[t]);
          }
        }
        l_sum[t] += tile_sum;
        tmem_st16(s_addr, pk);          // packed columns [32 part, 32 part + 16) = my keys [32 part, 32 part + 32)
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pfull[t]);
      }
    }
    // epilogue: row sum = the four parts' shares; O / l -> bf16 -> global (32 columns per thread); LSE by part 0
    float l_tot[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      *xslot(n_kv & 1, t, part) = l_sum[t];
    }
    row_sync();
#pragma unroll
    for (int t = 0; t < 2; ++t)
      l_tot[t] = (*xslot(n_kv & 1, t, 0) + *xslot(n_kv & 1, t, 1)) + (*xslot(n_kv & 1, t, 2) + *xslot(n_kv & 1, t, 3));
    mbar_wait(ofull, 0);
    tc_fence_after();
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const uint32_t o_addr = tmem_base + 256 + t * 128 + part * 32 + lane_off;
      const int row = q0 + t * QT + r;
      const float inv_l = 1.0f / l_tot[t];
      const bool row_ok = row < p.Lq;
      uint32_t o[32];
      tmem_ld32(o_addr, o);
      tmem_wait_ld();
      if (p.n_split > 1) {
        const int64_t prow = ((int64_t)blockIdx.x * p.n_split + blockIdx.y) * (2 * QT) + t * QT + r;
        float* dst = p.part_o + prow * HD + part * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l,
                                                            __uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
        if (part == 0) p.part_lse[prow] = m_used[t] + log2f(l_tot[t]);
      } else {
        __nv_bfloat16* orow;
        if (p.n_peer > 0) {
          const int rk = row_ok ? row / p.L_loc : 0;
          orow = p.o_peer[rk] + (int64_t)(row - rk * p.L_loc) * p.o_ld_tok + (int64_t)(p.head_off + head) * p.o_ld_head;
        } else {
          orow = p.o + (int64_t)row * p.o_ld_tok + (int64_t)head * p.o_ld_head;
        }
        orow += part * 32;
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
            v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
            v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
            v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + i) = v;
          }
          if (part == 0 && p.lse) p.lse[(int64_t)head * p.Lq + row] = (m_used[t] + log2f(l_tot[t])) * 0.6931471805599453f;
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------- softmax + epilogue -------------------------------
    // Two threads per query row: warp = (Q tile t, column half h, lane quadrant); thread (h, row) owns keys
    // [64h, 64h + 64) of every KV tile and columns [64h, 64h + 64) of O.  The tensor pipe waits on exactly this stage
    // (P_t(j) must exist 1024 tensor-cycles after S_t(j)), so its latency, not its throughput, is what matters.
    // The two threads of a row agree on the running max through shared memory (one 256-thread named barrier per tile).
    // Packed bf16 P of keys [64h, 64h+64) goes to TMEM columns [64h, 64h+32): it only aliases fp32 S columns that the
    // same thread has already consumed.
    const int t = warp >> 3;                 // Q tile
    const int half = (warp >> 2) & 1;        // which 64 keys / 64 output columns
    const int quad = warp & 3;               // TMEM lane quadrant
    const int r = quad * 32 + lane;          // row inside the Q tile
    const int row = q0 + t * QT + r;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + t * 128 + half * 64 + lane_off;
    const uint32_t o_addr = tmem_base + 256 + t * 128 + half * 64 + lane_off;
    float* x_mine = sX + (t * 2 + half) * 128 + r;          // + slot * 512: exchange slots alternate per tile
    float* x_peer = sX + (t * 2 + (half ^ 1)) * 128 + r;
    float m_used = -INFINITY;  // max the exponentials are currently relative to (log2 domain), identical in both threads
    float l_sum = 0.f;         // this thread's share of the row sum
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&sfull[t], j & 1);
      tc_fence_after();
      const int valid = p.Lk - (kv0 + j) * KT - half * 64;   // my columns >= valid are padding (only the last tile is ragged)
      uint32_t pk[32];                                // my 64 keys as packed bf16x2, stored only after the max check
      float tile_sum = 0.f, tile_max = -INFINITY;
      auto exp_pass = [&](const float m_ref) {
        const float2 neg_m2 = make_float2(-m_ref, -m_ref);
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t rr[32];
          tmem_ld32(s_addr + c * 32, rr);
          tmem_wait_ld();
          if (valid < 64) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= valid) rr[i] = 0xff800000u;  // -inf
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmax3(mx0, __uint_as_float(rr[i]), __uint_as_float(rr[i + 1]));
            mx1 = fmax3(mx1, __uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3]));
            float2 xa = ffma2(make_float2(__uint_as_float(rr[i]), __uint_as_float(rr[i + 1])), sc2, neg_m2);
            float2 xb = ffma2(make_float2(__uint_as_float(rr[i + 2]), __uint_as_float(rr[i + 3])), sc2, neg_m2);
            float2 pa, pb;
            pa = make_float2(fast_exp2(xa.x), fast_exp2(xa.y));
            if ((FMA_MASK >> (i >> 2)) & 1) pb = exp2_fma2(xb);   // FMA pipe instead of MUFU (compile-time pattern)
            else pb = make_float2(fast_exp2(xb.x), fast_exp2(xb.y));
            acc0 = fadd2(acc0, pa);
            acc1 = fadd2(acc1, pb);
            pk[c * 16 + (i >> 1)] = pack_bf16x2(pa.x, pa.y);
            pk[c * 16 + (i >> 1) + 1] = pack_bf16x2(pb.x, pb.y);
          }
        }
        tile_sum = (acc0.x + acc0.y) + (acc1.x + acc1.y);
        tile_max = fmaxf(mx0, mx1);
      };
      if (j == 0) {
        // first tile: the reference max must be the true row max over all 128 keys (one extra read of S from TMEM)
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t rr[32];
          tmem_ld32(s_addr + c * 32, rr);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(rr[i]));
        }
        x_mine[0] = mx;                                      // slot 0 (j & 1 == 0)
        named_bar_sync(1 + t, 256);
        m_used = fmaxf(mx, x_peer[0]) * p.scale_log2;
        exp_pass(m_used);
      } else {
        // optimistic: exponentiate against the stale max while tracking this tile's max; exact unless the row max grew
        // by more than 8 (log2 units): then O and l are rescaled and the tile is redone (rare after the first tiles)
        exp_pass(m_used);
        x_mine[(j & 1) * 512] = tile_max;
        named_bar_sync(1 + t, 256);
        const float m_new = fmaxf(m_used, fmaxf(tile_max, x_peer[(j & 1) * 512]) * p.scale_log2);
        // same rows in the same lanes of the partner warp => both warps take the same (warp-uniform) branch
        if (__any_sync(0xffffffffu, m_new > m_used + 8.0f)) {
          const float alpha = fast_exp2(m_used - m_new);
          l_sum *= alpha;
          m_used = m_new;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(o_addr + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(o_addr + c * 32, o);
          }
          exp_pass(m_used);
        }
      }
      l_sum += tile_sum;
      tmem_st32(s_addr, pk);          // packed columns [64h, 64h + 32) = my keys [64h, 64h + 64)
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[t]);
    }
    // epilogue: row sum = both halves' shares; O / l -> bf16 -> global (64 columns per thread); LSE by half 0
    x_mine[(n_kv & 1) * 512] = l_sum;
    named_bar_sync(1 + t, 256);
    const float l_tot = l_sum + x_peer[(n_kv & 1) * 512];
    mbar_wait(ofull, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l_tot;
    const bool row_ok = row < p.Lq;
    if (p.n_split > 1) {
      // one piece of a key-split unit: normalised fp32 partial + its LSE (log2 domain) for attn_combine_kernel
      const int64_t prow = ((int64_t)blockIdx.x * p.n_split + blockIdx.y) * (2 * QT) + t * QT + r;
      float* dst = p.part_o + prow * HD + half * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(o_addr + c * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(dst + c * 32 + i) = make_float4(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l,
                                                                      __uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
        __syncwarp();
      }
      if (half == 0) p.part_lse[prow] = m_used + log2f(l_tot);
    } else {
      __nv_bfloat16* orow;
      if (p.n_peer > 0) {
        const int rk = row_ok ? row / p.L_loc : 0;
        orow = p.o_peer[rk] + (int64_t)(row - rk * p.L_loc) * p.o_ld_tok + (int64_t)(p.head_off + head) * p.o_ld_head;
      } else {
        orow = p.o + (int64_t)row * p.o_ld_tok + (int64_t)head * p.o_ld_head;
      }
      orow += half * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(o_addr + c * 32, o);
        tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
            v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
            v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
            v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c * 32 + i) = v;
          }
        }
        __syncwarp();
      }
      if (row_ok && half == 0 && p.lse) p.lse[(int64_t)head * p.Lq + row] = (m_used + log2f(l_tot)) * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == AW_MMA) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Merge the key-split pieces of the tail units: one warp per query row, lane = 4 output columns.
//   w_s = 2^(lse_s - max lse), O = sum_s w_s O_s / sum_s w_s, LSE = (max + log2 sum_s w_s) ln 2
__global__ void __launch_bounds__(256) attn_combine_kernel(const AttnFwdParams p, int n_units) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)n_units * (2 * QT)) return;
  const int ui = (int)(w / (2 * QT)), rr = (int)(w % (2 * QT));
  const int unit = p.unit0 + ui;
  const int head = unit / p.nq;
  const int row = (unit % p.nq) * (2 * QT) + rr;
  if (row >= p.Lq) return;
  float m = -INFINITY;
  for (int s = 0; s < p.n_split; ++s) m = fmaxf(m, p.part_lse[((int64_t)ui * p.n_split + s) * (2 * QT) + rr]);
  float den = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < p.n_split; ++s) {
    const int64_t prow = ((int64_t)ui * p.n_split + s) * (2 * QT) + rr;
    const float ws = fast_exp2(p.part_lse[prow] - m);
    const float4 v = *reinterpret_cast<const float4*>(p.part_o + prow * HD + lane * 4);
    den += ws;
    acc.x += ws * v.x; acc.y += ws * v.y; acc.z += ws * v.z; acc.w += ws * v.w;
  }
  const float inv = 1.0f / den;
  __nv_bfloat16* orow;
  if (p.n_peer > 0) {
    const int rk = row / p.L_loc;
    orow = p.o_peer[rk] + (int64_t)(row - rk * p.L_loc) * p.o_ld_tok + (int64_t)(p.head_off + head) * p.o_ld_head;
  } else {
    orow = p.o + (int64_t)row * p.o_ld_tok + (int64_t)head * p.o_ld_head;
  }
  uint2 o;
  o.x = pack_bf16x2(acc.x * inv, acc.y * inv);
  o.y = pack_bf16x2(acc.z * inv, acc.w * inv);
  *reinterpret_cast<uint2*>(orow + lane * 4) = o;
  if (lane == 0 && p.lse) p.lse[(int64_t)head * p.Lq + row] = (m + log2f(den)) * 0.6931471805599453f;
}

// Ring / context-parallel merge (xdit_context_parallel.py:190-233 + xfuser's ring attention): combine the
// running result over the key blocks seen so far with the attention over one more key block.  One warp per (token, head),
// lane = 4 output columns.  LSEs in natural log (what prfl_attn_fwd writes).
//   m = max(lse_a, lse_n); w_a = e^(lse_a - m), w_n = e^(lse_n - m); acc = (w_a acc + w_n o_n) / (w_a + w_n); lse_a = m + ln(w_a + w_n)
__global__ void __launch_bounds__(256) attn_merge_kernel(float* __restrict__ acc, float* __restrict__ lse_acc,
                                                         const __nv_bfloat16* __restrict__ o_new, int64_t n_ld_tok, int64_t n_ld_head,
                                                         const float* __restrict__ lse_new, int first,
                                                         __nv_bfloat16* __restrict__ out, int64_t o_ld_tok, int64_t o_ld_head, int L, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)L * H) return;
  const int h = (int)(w % H);
  const int64_t i = w / H;
  const uint2 nv = *reinterpret_cast<const uint2*>(o_new + i * n_ld_tok + (int64_t)h * n_ld_head + lane * 4);
  float4 r = make_float4(bf16lo(nv.x), bf16hi(nv.x), bf16lo(nv.y), bf16hi(nv.y));
  float* a = acc + (i * H + h) * HD + lane * 4;
  const float ln = lse_new[(int64_t)h * L + i];
  float lse = ln;
  if (!first) {
    const float la = lse_acc[(int64_t)h * L + i];
    const float m = fmaxf(la, ln);
    const float wa = __expf(la - m), wn = __expf(ln - m);
    const float inv = 1.0f / (wa + wn);
    const float4 av = *reinterpret_cast<const float4*>(a);
    r.x = (wa * av.x + wn * r.x) * inv; r.y = (wa * av.y + wn * r.y) * inv;
    r.z = (wa * av.z + wn * r.z) * inv; r.w = (wa * av.w + wn * r.w) * inv;
    lse = m + __logf(wa + wn);
  }
  *reinterpret_cast<float4*>(a) = r;
  if (lane == 0) lse_acc[(int64_t)h * L + i] = lse;
  if (out) {
    uint2 o;
    o.x = pack_bf16x2(r.x, r.y);
    o.y = pack_bf16x2(r.z, r.w);
    *reinterpret_cast<uint2*>(out + i * o_ld_tok + (int64_t)h * o_ld_head + lane * 4) = o;
  }
}

}  // namespace prfl

using namespace prfl;

extern "C" int prfl_attn_merge(float* o_acc, float* lse_acc, const void* o_new_bf16, int64_t n_ld_tok, int64_t n_ld_head,
                               const float* lse_new, int first, void* out_bf16, int64_t o_ld_tok, int64_t o_ld_head, int L, int H,
                               prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(L > 0 && H > 0 && o_acc && lse_acc && o_new_bf16 && lse_new, PRFL_E_SHAPE, "attn_merge: L=%d H=%d", L, H);
  PRFL_REQUIRE(n_ld_tok % 4 == 0 && n_ld_head % 4 == 0 && o_ld_tok % 4 == 0 && o_ld_head % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(o_acc) & 15) == 0 && (reinterpret_cast<uintptr_t>(o_new_bf16) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0,
               PRFL_E_ALIGN, "attn_merge: alignment");
  const int64_t warps = (int64_t)L * H;
  attn_merge_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      o_acc, lse_acc, (const __nv_bfloat16*)o_new_bf16, n_ld_tok, n_ld_head, lse_new, first, (__nv_bfloat16*)out_bf16, o_ld_tok,
      o_ld_head, L, H);
  count_launch();
  PRFL_LAUNCH_CHECK("attn_merge");
  return PRFL_OK;
}

// Tail plan: units beyond the last full wave are split `n_split` ways along the key axis when that shortens the kernel.
struct SplitPlan {
  int n_units, n_main, n_tail, n_split;
};
static SplitPlan plan_split_for(int Lq, int Lk, int H, int sms, int force) {
  SplitPlan sp;
  const int nq = (Lq + 2 * QT - 1) / (2 * QT);
  sp.n_units = nq * H;
  sp.n_main = sp.n_units;
  sp.n_tail = 0;
  sp.n_split = 1;
  const int n_kv = (Lk + KT - 1) / KT;
  if (force == 0 || sms <= 0) return sp;
  int rem = sp.n_units % sms, s = 1;
  if (force >= 2) {
    s = force;
    if (rem == 0) rem = sp.n_units < sms ? sp.n_units : sms;
  } else if (sp.n_units > sms && rem > 0 && 2 * rem <= sms && n_kv >= 32) {
    s = sms / rem;
    if (s > 4) s = 4;
  }
  while (s >= 2 && (s - 1) * ((n_kv + s - 1) / s) >= n_kv) --s;     // every piece must own at least one KV tile
  if (s < 2 || n_kv < 2 * s) return sp;
  sp.n_tail = rem;
  sp.n_main = sp.n_units - rem;
  sp.n_split = s;
  return sp;
}
static SplitPlan plan_split(int Lq, int Lk, int H) {
  static const int force = [] { const char* e = getenv("PRFL_ATTN_SPLIT"); return e ? atoi(e) : -1; }();   // 0: off, 2..4: always
  return plan_split_for(Lq, Lk, H, sm_count(), force);
}

// The split policy as a pure host function (no device needed): out = {units, units in the plain launch, tail units, pieces}
extern "C" void prfl_attn_fwd_split_plan(int Lq, int Lk, int H, int n_sms, int* out4) {
  const SplitPlan sp = plan_split_for(Lq, Lk, H, n_sms, -1);
  out4[0] = sp.n_units; out4[1] = sp.n_main; out4[2] = sp.n_tail; out4[3] = sp.n_split;
}

extern "C" int64_t prfl_attn_fwd_ws_bytes(int Lq, int Lk, int H) {
  const SplitPlan sp = plan_split(Lq, Lk, H);
  return sp.n_tail == 0 ? 0 : (int64_t)sp.n_tail * sp.n_split * (2 * QT) * (HD + 1) * (int64_t)sizeof(float);
}

static int attn_fwd_launch(const void* q, int64_t q_ld_tok, int64_t q_ld_head, const void* k, int64_t k_ld_tok, int64_t k_ld_head,
                           const void* v, int64_t v_ld_tok, int64_t v_ld_head, void* o, int64_t o_ld_tok, int64_t o_ld_head,
                           float* lse, int Lq, int Lk, int H, float scale, void* const* o_peers, int n_peer, int L_loc,
                           int head_off, void* ws, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(Lq > 0 && Lk > 0 && H > 0, PRFL_E_SHAPE, "attn_fwd: Lq=%d Lk=%d H=%d", Lq, Lk, H);
  PRFL_REQUIRE(q_ld_tok % 8 == 0 && q_ld_head % 8 == 0 && k_ld_tok % 8 == 0 && k_ld_head % 8 == 0 && v_ld_tok % 8 == 0 &&
                   v_ld_head % 8 == 0 && o_ld_tok % 8 == 0 && o_ld_head % 8 == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0,
               PRFL_E_ALIGN, "attn_fwd: strides must be multiples of 8 elements, pointers 16-byte aligned");
  PRFL_REQUIRE(n_peer >= 0 && n_peer <= 8 && (n_peer == 0 || (o_peers && L_loc > 0 && (int64_t)L_loc * n_peer >= Lq)), PRFL_E_SHAPE,
               "attn_fwd: n_peer=%d L_loc=%d Lq=%d", n_peer, L_loc, Lq);
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_3d(&tmQ, q, HD, (uint64_t)Lq, (uint64_t)H, (uint64_t)q_ld_tok * 2, (uint64_t)q_ld_head * 2, 64, QT, 1, 1);
  if (rc != PRFL_OK) return rc;
  rc = make_tmap_3d(&tmK, k, HD, (uint64_t)Lk, (uint64_t)H, (uint64_t)k_ld_tok * 2, (uint64_t)k_ld_head * 2, 64, KT, 1, 1);
  if (rc != PRFL_OK) return rc;
  rc = make_tmap_3d(&tmV, v, HD, (uint64_t)Lk, (uint64_t)H, (uint64_t)v_ld_tok * 2, (uint64_t)v_ld_head * 2, 64, KT, 1, 1);
  if (rc != PRFL_OK) return rc;
  // measured at L = 32 760 x 40 heads: 0 % -> 20.98 ms, 25 % -> 18.34 ms, 37.5 % -> 19.40 ms, 50 % -> 19.96 ms
  // Default: four threads per row, 12.5 % of the exponentials on the FMA pipe, one 128-thread row barrier per lane quadrant.
  // Measured on one box, back to back, L = 32 760 x 40 heads (CUDA events, sustained): two threads per row (round 1) 19.03 ms;
  // four threads per row with a 512-thread barrier 18.5; with per-quadrant barriers and 0 / 12.5 / 25 / 37.5 % FMA-pipe
  // exponentials 18.3 / 17.8 / 18.3 / 17.85 ms.  PRFL_ATTN_FWD=pair | quad25 | quad37 | quad0 | quadw select the others for A/B.
  static const int variant = [] {
    const char* e = getenv("PRFL_ATTN_FWD");
    if (!e) return 3;
    if (e[0] == 'p' || e[0] == 'b') return 0;         // pair / base: the round-1 kernel
    if (e[0] != 'q') return 3;
    if (e[4] == '2') return 1;
    if (e[4] == '3') return 2;
    if (e[4] == '0') return 4;
    return 3;
  }();
  static const int bar_wide = [] { const char* e = getenv("PRFL_ATTN_FWD"); return (e && e[0] == 'q' && e[strlen(e) - 1] == 'w') ? 1 : 0; }();
  void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, AttnFwdParams) =
      variant == 1 ? attn_fwd_kernel<0xAA, true> : variant == 2 ? attn_fwd_kernel<0xEE, true> : variant == 3 ? attn_fwd_kernel<0x88, true>
      : variant == 4 ? attn_fwd_kernel<0x00, true> : attn_fwd_kernel<0xAA, false>;
  static unsigned long long attr_mask = 0;      // per-device bit mask; forward and autograd threads may race: the call is idempotent
  if (device_needs_init(&attr_mask)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<0xAA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<0xAA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<0xEE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<0x88, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fwd_kernel<0x00, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "attn_fwd: cudaFuncSetAttribute");
    device_mark_init(&attr_mask);
  }
  AttnFwdParams p;
  p.o = (__nv_bfloat16*)o; p.o_ld_tok = o_ld_tok; p.o_ld_head = o_ld_head; p.lse = lse; p.Lq = Lq; p.Lk = Lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.n_peer = n_peer; p.L_loc = L_loc; p.head_off = head_off;
  p.bar_wide = bar_wide;
  for (int i = 0; i < 8; ++i) p.o_peer[i] = i < n_peer ? (__nv_bfloat16*)o_peers[i] : nullptr;
  SplitPlan sp = plan_split(Lq, Lk, H);
  if (ws == nullptr) {                      // no workspace from the caller: one launch over all units
    sp.n_main = sp.n_units;
    sp.n_tail = 0;
  }
  PRFL_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, PRFL_E_ALIGN, "attn_fwd: workspace must be 16-byte aligned");
  p.nq = (Lq + 2 * QT - 1) / (2 * QT);
  p.unit0 = 0;
  p.n_split = 1;
  p.part_o = p.part_lse = nullptr;
  if (sp.n_main > 0) {
    kern<<<dim3(sp.n_main, 1), ATT_THREADS, ATT_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
    count_launch();
    PRFL_LAUNCH_CHECK("attn_fwd");
  }
  if (sp.n_tail > 0) {
    p.unit0 = sp.n_main;
    p.n_split = sp.n_split;
    p.part_o = static_cast<float*>(ws);
    p.part_lse = p.part_o + (int64_t)sp.n_tail * sp.n_split * (2 * QT) * HD;
    kern<<<dim3(sp.n_tail, sp.n_split), ATT_THREADS, ATT_SMEM, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
    count_launch();
    PRFL_LAUNCH_CHECK("attn_fwd (key-split tail)");
    const int64_t rows = (int64_t)sp.n_tail * (2 * QT);
    attn_combine_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p, sp.n_tail);
    count_launch();
    PRFL_LAUNCH_CHECK("attn_combine");
  }
  return PRFL_OK;
}

extern "C" int prfl_attn_fwd(const void* q, int64_t q_ld_tok, int64_t q_ld_head, const void* k, int64_t k_ld_tok,
                             int64_t k_ld_head, const void* v, int64_t v_ld_tok, int64_t v_ld_head, void* o, int64_t o_ld_tok,
                             int64_t o_ld_head, float* lse, int Lq, int Lk, int H, float scale, void* ws, prfl_stream_t stream) {
  return attn_fwd_launch(q, q_ld_tok, q_ld_head, k, k_ld_tok, k_ld_head, v, v_ld_tok, v_ld_head, o, o_ld_tok, o_ld_head, lse, Lq, Lk,
                         H, scale, nullptr, 0, 0, 0, ws, stream);
}

extern "C" int prfl_attn_fwd_p2p(const void* q, int64_t q_ld_tok, int64_t q_ld_head, const void* k, int64_t k_ld_tok,
                                 int64_t k_ld_head, const void* v, int64_t v_ld_tok, int64_t v_ld_head, void* const* o_peers,
                                 int n_peer, int L_loc, int head_off, int64_t o_ld_tok, int64_t o_ld_head, float* lse, int Lq,
                                 int Lk, int H, float scale, void* ws, prfl_stream_t stream) {
  PRFL_REQUIRE(n_peer >= 1 && o_peers, PRFL_E_SHAPE, "attn_fwd_p2p: needs peer pointers");
  return attn_fwd_launch(q, q_ld_tok, q_ld_head, k, k_ld_tok, k_ld_head, v, v_ld_tok, v_ld_head, o_peers[0], o_ld_tok, o_ld_head, lse,
                         Lq, Lk, H, scale, o_peers, n_peer, L_loc, head_off, ws, stream);
}
