// Memory-bound kernels of the Wan-DiT block: LayerNorm+AdaLN modulate, RMSNorm+3-D RoPE, casts,
// patchify / unpatchify.  All are HBM-bound: one 4-warp CTA owns one row, the whole row lives in
// registers between its single 128-bit-vectorised read and its single write, reductions are
// warp shuffles + one shared-memory exchange.  Roofline per row (C channels): ln_mod fwd 4C read + 2C written; rmsnorm_rope
// 2C + 2C (+ 512 B of cos/sin).
#include "common.cuh"

namespace prfl {

// Row kernels: one CTA (4 warps) per row, thread t owns the 8-channel pieces t, t+128, ... (PER = ceil(C/1024)).  The row
// payload is a few dozen registers per thread (the one-warp-per-row form needed 226-255 registers at C = 5120, i.e. 8
// warps per SM), every load/store instruction of a warp touches 32 consecutive 16- or 32-byte pieces, and a row costs
// one or two block reductions through shared memory.
__device__ __forceinline__ float block_sum(float v, float* red /*[2][4]*/, int phase) {
  v = warp_sum(v);
  float* r = red + (phase & 1) * 4;
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  return (r[0] + r[1]) + (r[2] + r[3]);
}

// =============================================================================================
// LayerNorm + modulate   (model.py:125-135, 345, 352, 353)
// =============================================================================================
template <int NCH>
__global__ void __launch_bounds__(128) ln_mod_fwd_kernel(const float* __restrict__ x, const float* __restrict__ shift,
                                                         const float* __restrict__ scale, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                                         float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                         int64_t rows, float eps, int round_bf16,
                                                         __nv_bfloat16* __restrict__ out_lo = nullptr) {
  constexpr int C = NCH * 256, NPIECE = NCH * 32, PER = (NPIECE + 127) / 128;
  __shared__ float red[8];
  int phase = 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float* xr = x + row * C;
    float v[PER][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int pc = threadIdx.x + 128 * i;
      const bool ok = NPIECE % 128 == 0 || pc < NPIECE;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (ok) {
        const float4* p = reinterpret_cast<const float4*>(xr + 8 * pc);
        a = __ldcs(p);
        b = __ldcs(p + 1);
      }
      v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
      v[i][4] = b.x; v[i][5] = b.y; v[i][6] = b.z; v[i][7] = b.w;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
    const float mean = block_sum(s, red, phase++) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const bool ok = NPIECE % 128 == 0 || threadIdx.x + 128 * i < NPIECE;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q += ok ? d * d : 0.f;
      }
    }
    const float rstd = rsqrtf(block_sum(q, red, phase++) * (1.0f / C) + eps);
    if (threadIdx.x == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    __nv_bfloat16* orow = out + row * C;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int pc = threadIdx.x + 128 * i;
      if (NPIECE % 128 != 0 && pc >= NPIECE) continue;
      const int c0 = 8 * pc;
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        y[j] = (v[i][j] - mean) * rstd;
        if (round_bf16) y[j] = bf16_round(y[j]);
      }
      if (gamma) {
        float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
        float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
        y[0] = y[0] * g0.x + b0.x; y[1] = y[1] * g0.y + b0.y; y[2] = y[2] * g0.z + b0.z; y[3] = y[3] * g0.w + b0.w;
        y[4] = y[4] * g1.x + b1.x; y[5] = y[5] * g1.y + b1.y; y[6] = y[6] * g1.z + b1.z; y[7] = y[7] * g1.w + b1.w;
      }
      if (scale) {
        float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c0) + 1);
        float4 t0 = __ldg(reinterpret_cast<const float4*>(shift + c0)), t1 = __ldg(reinterpret_cast<const float4*>(shift + c0) + 1);
        y[0] = y[0] * (1.f + s0.x) + t0.x; y[1] = y[1] * (1.f + s0.y) + t0.y;
        y[2] = y[2] * (1.f + s0.z) + t0.z; y[3] = y[3] * (1.f + s0.w) + t0.w;
        y[4] = y[4] * (1.f + s1.x) + t1.x; y[5] = y[5] * (1.f + s1.y) + t1.y;
        y[6] = y[6] * (1.f + s1.z) + t1.z; y[7] = y[7] * (1.f + s1.w) + t1.w;
      }
      uint4 o;
      o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
      o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
      stg_v4(orow + c0, o);
      if (out_lo) {  // residual of the bf16 rounding: y = hi + lo to ~16 mantissa bits (fp32-grade head GEMM)
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = y[j] - bf16_round(y[j]);
        o.x = pack_bf16x2(r[0], r[1]); o.y = pack_bf16x2(r[2], r[3]);
        o.z = pack_bf16x2(r[4], r[5]); o.w = pack_bf16x2(r[6], r[7]);
        stg_v4(out_lo + row * C + c0, o);
      }
    }
  }
}

// =============================================================================================
// RMSNorm (+ RoPE)   (model.py:106-122, 60-103)
// =============================================================================================
template <int NCH>
__global__ void __launch_bounds__(128) rmsnorm_rope_fwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                               const float* __restrict__ w, const float* __restrict__ cos_tab,
                                                               const float* __restrict__ sin_tab, __nv_bfloat16* __restrict__ out,
                                                               int64_t ldo, float* __restrict__ rstd_out, int64_t rows,
                                                               int64_t n_rot, int64_t pos0, float eps) {
  constexpr int C = NCH * 256, NPIECE = NCH * 32, PER = (NPIECE + 127) / 128;
  __shared__ float red[8];
  int phase = 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const __nv_bfloat16* xr = x + row * ldx;
    uint4 raw[PER];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int pc = threadIdx.x + 128 * i;
      raw[i] = (NPIECE % 128 == 0 || pc < NPIECE) ? ldg_nc_v4(xr + 8 * pc) : make_uint4(0u, 0u, 0u, 0u);
      const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = bf16lo(u[j]), b = bf16hi(u[j]);
        ss += a * a + b * b;
      }
    }
    const float rstd = rsqrtf(block_sum(ss, red, phase++) * (1.0f / C) + eps);
    if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
    const bool rot = cos_tab != nullptr && row < n_rot;
    const float* cr = rot ? cos_tab + (pos0 + row) * 64 : nullptr;
    const float* sr = rot ? sin_tab + (pos0 + row) * 64 : nullptr;
    __nv_bfloat16* orow = out + row * ldo;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int pc = threadIdx.x + 128 * i;
      if (NPIECE % 128 != 0 && pc >= NPIECE) continue;
      const int c0 = 8 * pc;
      const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
      float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c0)), w1 = __ldg(reinterpret_cast<const float4*>(w + c0) + 1);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float y[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        y[2 * j] = bf16_round(bf16lo(u[j]) * rstd) * wv[2 * j];
        y[2 * j + 1] = bf16_round(bf16hi(u[j]) * rstd) * wv[2 * j + 1];
      }
      if (rot) {
        const int j0 = (c0 & 127) >> 1;  // first complex pair of this 8-channel piece inside its head
        float4 cv = __ldg(reinterpret_cast<const float4*>(cr + j0));
        float4 sv = __ldg(reinterpret_cast<const float4*>(sr + j0));
        const float cc[4] = {cv.x, cv.y, cv.z, cv.w}, sn[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float re = y[2 * j], im = y[2 * j + 1];
          y[2 * j] = re * cc[j] - im * sn[j];
          y[2 * j + 1] = re * sn[j] + im * cc[j];
        }
      }
      uint4 o;
      o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
      o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
      stg_v4(orow + c0, o);
    }
  }
}

// =============================================================================================
// cast, patchify, unpatchify
// =============================================================================================
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n8 = n >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i), b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  for (int64_t i = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16_rn(src[i]);
}

// one thread = one (channel, f, h2, w2) 2x2 patch: coalesced float2 reads along w, 8-byte writes
__global__ void patchify_kernel(const float* __restrict__ x, int Cx, const float* __restrict__ y, int Cy,
                                __nv_bfloat16* __restrict__ patches, int F, int H, int W) {
  const int h2 = H >> 1, w2 = W >> 1;
  const int Ct = Cx + Cy;
  const int64_t total = (int64_t)Ct * F * h2 * w2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int wi = (int)(idx % w2);
    int64_t t = idx / w2;
    int hi = (int)(t % h2);
    t /= h2;
    int f = (int)(t % F);
    int c = (int)(t / F);
    const float* src = (c < Cx) ? x + (((int64_t)c * F + f) * H + 2 * hi) * W + 2 * wi
                                : y + (((int64_t)(c - Cx) * F + f) * H + 2 * hi) * W + 2 * wi;
    float2 r0 = *reinterpret_cast<const float2*>(src);
    float2 r1 = *reinterpret_cast<const float2*>(src + W);
    int64_t l = ((int64_t)f * h2 + hi) * w2 + wi;
    uint2 o;
    o.x = pack_bf16x2(r0.x, r0.y);
    o.y = pack_bf16x2(r1.x, r1.y);
    *reinterpret_cast<uint2*>(patches + l * (Ct * 4) + c * 4) = o;
  }
}

__global__ void patchify_bwd_kernel(const float* __restrict__ dp, int Cx, int Ct, float* __restrict__ dx, int F, int H, int W) {
  const int h2 = H >> 1, w2 = W >> 1;
  const int64_t total = (int64_t)Cx * F * h2 * w2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int wi = (int)(idx % w2);
    int64_t t = idx / w2;
    int hi = (int)(t % h2);
    t /= h2;
    int f = (int)(t % F);
    int c = (int)(t / F);
    int64_t l = ((int64_t)f * h2 + hi) * w2 + wi;
    float4 g = *reinterpret_cast<const float4*>(dp + l * (Ct * 4) + c * 4);
    float* dst = dx + (((int64_t)c * F + f) * H + 2 * hi) * W + 2 * wi;
    *reinterpret_cast<float2*>(dst) = make_float2(g.x, g.y);
    *reinterpret_cast<float2*>(dst + W) = make_float2(g.z, g.w);
  }
}

// tokens [F*h*w, (p=1,q=2,r=2,c)] -> video [c, F, 2h, 2w]   ('fhwpqrc->cfphqwr', model.py:700-703)
__global__ void unpatchify_kernel(const float* __restrict__ tok, float* __restrict__ vid, int c, int F, int h, int w, int inverse) {
  const int64_t total = (int64_t)c * F * h * w;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int wi = (int)(idx % w);
    int64_t t = idx / w;
    int hi = (int)(t % h);
    t /= h;
    int f = (int)(t % F);
    int ci = (int)(t / F);
    int64_t l = ((int64_t)f * h + hi) * w + wi;
    const int64_t tb = l * (4 * c) + ci;  // + (q*2 + r) * c
    float* v0 = vid + (((int64_t)ci * F + f) * (2 * h) + 2 * hi) * (2 * w) + 2 * wi;
    if (!inverse) {
      *reinterpret_cast<float2*>(v0) = make_float2(tok[tb], tok[tb + c]);
      *reinterpret_cast<float2*>(v0 + 2 * w) = make_float2(tok[tb + 2 * c], tok[tb + 3 * c]);
    } else {
      float2 a = *reinterpret_cast<const float2*>(v0), b = *reinterpret_cast<const float2*>(v0 + 2 * w);
      float* tw = const_cast<float*>(tok);
      tw[tb] = a.x; tw[tb + c] = a.y; tw[tb + 2 * c] = b.x; tw[tb + 3 * c] = b.y;
    }
  }
}

// Ulysses staging copies, 16 bytes per thread.  `strided` holds [L_loc, H, 128] with element (t, h, d) at
// t*ld_tok + h*ld_head + d; `packed` is the contiguous all-to-all buffer [P][L_loc][H/P][128].
//   mode 0 (pack, before the q/k/v exchange)      : packed[p][t][hl][:] = strided[t][p*Hl + hl][:]
//   mode 1 (unpack, after the attention-out exchange): strided[t][p*Hl + hl][:] = packed[p][t][hl][:]
// (After the q/k/v exchange the receive buffer already IS [L, H/P, 128] in global token order, and before the
// output exchange the attention output [L, H/P, 128] already IS the send buffer, so those two need no copy.)
__global__ void a2a_pack_kernel(__nv_bfloat16* __restrict__ strided, int64_t ld_tok, int64_t ld_head,
                                __nv_bfloat16* __restrict__ packed, int L_loc, int H, int P, int mode) {
  const int Hl = H / P;
  const int64_t total = (int64_t)P * L_loc * Hl * 16;  // 16-byte pieces (16 per 128-wide head)
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int piece = (int)(idx & 15);
    int64_t t = idx >> 4;
    int hl = (int)(t % Hl);
    t /= Hl;
    int tok = (int)(t % L_loc);
    int p = (int)(t / L_loc);
    __nv_bfloat16* s = strided + (int64_t)tok * ld_tok + (int64_t)(p * Hl + hl) * ld_head + piece * 8;
    if (mode == 0)
      reinterpret_cast<uint4*>(packed)[idx] = *reinterpret_cast<const uint4*>(s);
    else
      *reinterpret_cast<uint4*>(s) = reinterpret_cast<const uint4*>(packed)[idx];
  }
}

// Ulysses q/k/v exchange without NCCL: every rank stores its token chunk straight into the peers' receive buffers
// (peer-mapped NVLink pointers): peer[p][(rank*L_loc + t), hl, :] = strided[t][p*Hl + hl][:]
struct PeerPtrs {
  __nv_bfloat16* p[8];
};
__global__ void a2a_scatter_p2p_kernel(const __nv_bfloat16* __restrict__ strided, int64_t ld_tok, int64_t ld_head, PeerPtrs peers,
                                       int L_loc, int H, int P, int rank) {
  const int Hl = H / P;
  const int64_t total = (int64_t)P * L_loc * Hl * 16;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int piece = (int)(idx & 15);
    int64_t t = idx >> 4;
    int hl = (int)(t % Hl);
    t /= Hl;
    int tok = (int)(t % L_loc);
    int p = (int)(t / L_loc);
    const uint4 v = *reinterpret_cast<const uint4*>(strided + (int64_t)tok * ld_tok + (int64_t)(p * Hl + hl) * ld_head + piece * 8);
    *reinterpret_cast<uint4*>(peers.p[p] + (((int64_t)rank * L_loc + tok) * Hl + hl) * 128 + piece * 8) = v;
  }
}

// The reverse exchange (attention layout -> token-chunk layout) as direct peer stores: this rank holds all L tokens of its
// Hl heads; token chunk p goes to rank p, at this rank's head offset:
//   peer[p][t * dst_ld_tok + (rank*Hl + hl) * dst_ld_head + :] = src[(p*L_loc + t) * src_ld_tok + hl * src_ld_head + :]
__global__ void a2a_gather_p2p_kernel(const __nv_bfloat16* __restrict__ src, int64_t src_ld_tok, int64_t src_ld_head, PeerPtrs peers,
                                      int64_t dst_ld_tok, int64_t dst_ld_head, int L_loc, int Hl, int P, int rank) {
  const int64_t total = (int64_t)P * L_loc * Hl * 16;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int piece = (int)(idx & 15);
    int64_t t = idx >> 4;
    int hl = (int)(t % Hl);
    t /= Hl;                                   // global token index p * L_loc + tok
    int p = (int)(t / L_loc);
    int tok = (int)(t - (int64_t)p * L_loc);
    const uint4 v = *reinterpret_cast<const uint4*>(src + t * src_ld_tok + (int64_t)hl * src_ld_head + piece * 8);
    *reinterpret_cast<uint4*>(peers.p[p] + (int64_t)tok * dst_ld_tok + (int64_t)(rank * Hl + hl) * dst_ld_head + piece * 8) = v;
  }
}

template <typename F>
static int dispatch_nch(int C, F&& f) {
  switch (C / 256) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    case 3: return f(std::integral_constant<int, 3>{});
    case 4: return f(std::integral_constant<int, 4>{});
    case 6: return f(std::integral_constant<int, 6>{});
    case 8: return f(std::integral_constant<int, 8>{});
    case 12: return f(std::integral_constant<int, 12>{});
    case 16: return f(std::integral_constant<int, 16>{});
    case 20: return f(std::integral_constant<int, 20>{});
    default:
      set_error("unsupported channel count C=%d (need C/256 in {1,2,3,4,6,8,12,16,20})", C);
      return PRFL_E_SHAPE;
  }
}

static inline int row_grid(int64_t rows) {
  // one 4-warp CTA per row, grid-stride; 16 resident CTAs per SM is the hardware ceiling at 128 threads
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(rows < cap ? (rows > 0 ? rows : 1) : cap);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace prfl

using namespace prfl;

extern "C" {

int prfl_ln_mod_fwd(const float* x, const float* shift, const float* scale, const float* gamma, const float* beta,
                    void* out_bf16, float* mean, float* rstd, int64_t rows, int C, float eps, int round_bf16,
                    prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(rows >= 0 && C > 0 && C % 256 == 0, PRFL_E_SHAPE, "ln_mod_fwd: rows=%lld C=%d", (long long)rows, C);
  PRFL_REQUIRE((shift == nullptr) == (scale == nullptr) && (gamma == nullptr) == (beta == nullptr), PRFL_E_SHAPE,
               "ln_mod_fwd: shift/scale and gamma/beta must be given in pairs");
  PRFL_REQUIRE(aligned16(x) && aligned16(out_bf16) && aligned16(shift) && aligned16(scale) && aligned16(gamma) && aligned16(beta),
               PRFL_E_ALIGN, "ln_mod_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return PRFL_OK;
  return dispatch_nch(C, [&](auto nch) {
    constexpr int NCH = decltype(nch)::value;
    ln_mod_fwd_kernel<NCH><<<row_grid(rows), 128, 0, (cudaStream_t)stream>>>(
        x, shift, scale, gamma, beta, (__nv_bfloat16*)out_bf16, mean, rstd, rows, eps, round_bf16);
    count_launch();
    PRFL_LAUNCH_CHECK("ln_mod_fwd");
    return PRFL_OK;
  });
}

int prfl_ln_mod_split_fwd(const float* x, const float* shift, const float* scale, void* out_hi_bf16, void* out_lo_bf16,
                          int64_t rows, int C, float eps, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(rows >= 0 && C > 0 && C % 256 == 0 && out_hi_bf16 && out_lo_bf16, PRFL_E_SHAPE, "ln_mod_split_fwd: rows=%lld C=%d",
               (long long)rows, C);
  PRFL_REQUIRE((shift == nullptr) == (scale == nullptr), PRFL_E_SHAPE, "ln_mod_split_fwd: shift/scale come in pairs");
  PRFL_REQUIRE(aligned16(x) && aligned16(out_hi_bf16) && aligned16(out_lo_bf16) && aligned16(shift) && aligned16(scale), PRFL_E_ALIGN,
               "ln_mod_split_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return PRFL_OK;
  return dispatch_nch(C, [&](auto nch) {
    constexpr int NCH = decltype(nch)::value;
    ln_mod_fwd_kernel<NCH><<<row_grid(rows), 128, 0, (cudaStream_t)stream>>>(
        x, shift, scale, nullptr, nullptr, (__nv_bfloat16*)out_hi_bf16, nullptr, nullptr, rows, eps, 0, (__nv_bfloat16*)out_lo_bf16);
    count_launch();
    PRFL_LAUNCH_CHECK("ln_mod_split_fwd");
    return PRFL_OK;
  });
}

int prfl_rmsnorm_rope_fwd(const void* x_bf16, int64_t ldx, const float* w, const float* cos_tab, const float* sin_tab,
                          void* out_bf16, int64_t ldo, float* rstd, int64_t rows, int C, int64_t n_rot, int64_t pos0,
                          float eps, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(rows >= 0 && C > 0 && C % 256 == 0 && ldx >= C && ldo >= C, PRFL_E_SHAPE, "rmsnorm_rope_fwd: rows=%lld C=%d ldx=%lld ldo=%lld",
               (long long)rows, C, (long long)ldx, (long long)ldo);
  PRFL_REQUIRE((cos_tab == nullptr) == (sin_tab == nullptr), PRFL_E_SHAPE, "rmsnorm_rope_fwd: cos/sin must both be given");
  PRFL_REQUIRE(aligned16(x_bf16) && aligned16(out_bf16) && aligned16(w) && aligned16(cos_tab) && aligned16(sin_tab) &&
                   ldx % 8 == 0 && ldo % 8 == 0,
               PRFL_E_ALIGN, "rmsnorm_rope_fwd: pointers / leading dims must be 16-byte aligned");
  if (rows == 0) return PRFL_OK;
  return dispatch_nch(C, [&](auto nch) {
    constexpr int NCH = decltype(nch)::value;
    rmsnorm_rope_fwd_kernel<NCH><<<row_grid(rows), 128, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x_bf16, ldx, w, cos_tab, sin_tab, (__nv_bfloat16*)out_bf16, ldo, rstd, rows, n_rot, pos0, eps);
    count_launch();
    PRFL_LAUNCH_CHECK("rmsnorm_rope_fwd");
    return PRFL_OK;
  });
}

int prfl_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(n >= 0, PRFL_E_SHAPE, "cast: n=%lld", (long long)n);
  PRFL_REQUIRE(aligned16(src) && aligned16(dst_bf16), PRFL_E_ALIGN, "cast: pointers must be 16-byte aligned");
  if (n == 0) return PRFL_OK;
  int64_t blocks = (n / 8 + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 16;
  cast_f32_bf16_kernel<<<(int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap)), 256, 0, (cudaStream_t)stream>>>(
      src, (__nv_bfloat16*)dst_bf16, n);
  count_launch();
  PRFL_LAUNCH_CHECK("cast_f32_bf16");
  return PRFL_OK;
}

int prfl_patchify(const float* x, int Cx, const float* y, int Cy, void* patches_bf16, int F, int H, int W,
                  prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(Cx > 0 && Cy >= 0 && F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, PRFL_E_SHAPE,
               "patchify: Cx=%d Cy=%d F=%d H=%d W=%d", Cx, Cy, F, H, W);
  PRFL_REQUIRE((Cy == 0) == (y == nullptr), PRFL_E_SHAPE, "patchify: y and Cy disagree");
  PRFL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0 && aligned16(patches_bf16),
               PRFL_E_ALIGN, "patchify: alignment");
  int64_t total = (int64_t)(Cx + Cy) * F * (H / 2) * (W / 2);
  int64_t blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  patchify_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(x, Cx, y, Cy, (__nv_bfloat16*)patches_bf16,
                                                                                        F, H, W);
  count_launch();
  PRFL_LAUNCH_CHECK("patchify");
  return PRFL_OK;
}

int prfl_patchify_bwd(const float* dpatches, int Cx, int Ctot, float* dx, int F, int H, int W, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(Cx > 0 && Ctot >= Cx && F > 0 && H % 2 == 0 && W % 2 == 0, PRFL_E_SHAPE, "patchify_bwd: shapes");
  PRFL_REQUIRE(aligned16(dpatches) && (reinterpret_cast<uintptr_t>(dx) & 7) == 0, PRFL_E_ALIGN, "patchify_bwd: alignment");
  int64_t total = (int64_t)Cx * F * (H / 2) * (W / 2);
  int64_t blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  patchify_bwd_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(dpatches, Cx, Ctot, dx, F, H, W);
  count_launch();
  PRFL_LAUNCH_CHECK("patchify_bwd");
  return PRFL_OK;
}

int prfl_unpatchify(const float* tokens, float* video, int c, int F, int h, int w, int inverse, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(c > 0 && F > 0 && h > 0 && w > 0, PRFL_E_SHAPE, "unpatchify: shapes");
  PRFL_REQUIRE((reinterpret_cast<uintptr_t>(video) & 7) == 0, PRFL_E_ALIGN, "unpatchify: alignment");
  int64_t total = (int64_t)c * F * h * w;
  int64_t blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  unpatchify_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(tokens, video, c, F, h, w, inverse);
  count_launch();
  PRFL_LAUNCH_CHECK("unpatchify");
  return PRFL_OK;
}

int prfl_a2a_pack(void* strided, int64_t ld_tok, int64_t ld_head, void* packed, int L_loc, int H, int P, int mode,
                  prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(L_loc > 0 && H > 0 && P > 0 && H % P == 0 && (mode == 0 || mode == 1), PRFL_E_SHAPE,
               "a2a_pack: L_loc=%d H=%d P=%d mode=%d", L_loc, H, P, mode);
  PRFL_REQUIRE(aligned16(strided) && aligned16(packed) && ld_tok % 8 == 0 && ld_head % 8 == 0, PRFL_E_ALIGN, "a2a_pack: alignment");
  int64_t total = (int64_t)L_loc * H * 16;
  int64_t blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  a2a_pack_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)strided, ld_tok, ld_head,
                                                                                        (__nv_bfloat16*)packed, L_loc, H, P, mode);
  count_launch();
  PRFL_LAUNCH_CHECK("a2a_pack");
  return PRFL_OK;
}

int prfl_a2a_scatter_p2p(const void* strided, int64_t ld_tok, int64_t ld_head, void* const* peer_recv, int L_loc, int H, int P,
                         int rank, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(L_loc > 0 && H > 0 && P > 0 && P <= 8 && H % P == 0 && rank >= 0 && rank < P && peer_recv, PRFL_E_SHAPE,
               "a2a_scatter_p2p: L_loc=%d H=%d P=%d rank=%d", L_loc, H, P, rank);
  PRFL_REQUIRE(aligned16(strided) && ld_tok % 8 == 0 && ld_head % 8 == 0, PRFL_E_ALIGN, "a2a_scatter_p2p: alignment");
  PeerPtrs pp;
  for (int i = 0; i < 8; ++i) pp.p[i] = i < P ? (__nv_bfloat16*)peer_recv[i] : nullptr;
  int64_t total = (int64_t)L_loc * H * 16;
  int64_t blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  a2a_scatter_p2p_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)strided, ld_tok,
                                                                                               ld_head, pp, L_loc, H, P, rank);
  count_launch();
  PRFL_LAUNCH_CHECK("a2a_scatter_p2p");
  return PRFL_OK;
}

int prfl_a2a_gather_p2p(const void* src, int64_t src_ld_tok, int64_t src_ld_head, void* const* peer_dst, int64_t dst_ld_tok,
                        int64_t dst_ld_head, int L_loc, int Hl, int P, int rank, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(L_loc > 0 && Hl > 0 && P >= 1 && P <= 8 && rank >= 0 && rank < P && peer_dst, PRFL_E_SHAPE,
               "a2a_gather_p2p: L_loc=%d Hl=%d P=%d rank=%d", L_loc, Hl, P, rank);
  PRFL_REQUIRE(aligned16(src) && src_ld_tok % 8 == 0 && src_ld_head % 8 == 0 && dst_ld_tok % 8 == 0 && dst_ld_head % 8 == 0, PRFL_E_ALIGN,
               "a2a_gather_p2p: alignment");
  PeerPtrs pp;
  for (int i = 0; i < 8; ++i) {
    pp.p[i] = i < P ? (__nv_bfloat16*)peer_dst[i] : nullptr;
    PRFL_REQUIRE(i >= P || aligned16(peer_dst[i]), PRFL_E_ALIGN, "a2a_gather_p2p: peer pointer %d not 16-byte aligned", i);
  }
  int64_t total = (int64_t)P * L_loc * Hl * 16;
  int64_t blocks = (total + 255) / 256, cap = (int64_t)sm_count() * 16;
  a2a_gather_p2p_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)src, src_ld_tok, src_ld_head, pp, dst_ld_tok, dst_ld_head, L_loc, Hl, P, rank);
  count_launch();
  PRFL_LAUNCH_CHECK("a2a_gather_p2p");
  return PRFL_OK;
}

}  // extern "C"
