// Memory-bound backward kernels of the Wan-DiT block: LayerNorm(+modulate/affine) backward, RMSNorm(+RoPE)
// backward (row kernels: one 4-warp CTA per row, shuffle + shared-memory reductions) and the token-dimension reductions that produce
// bias / modulation / gate / norm-weight gradients (column kernels: a CTA owns 256 columns x a chunk of rows and
// writes one partial row; the [nparts, N] partials are summed by the caller).
#include "common.cuh"

namespace prfl {

constexpr int COL_ROWS = 256;   // rows per partial

// Row kernels here: one CTA (4 warps) per row, thread t owns the 8-channel pieces t, t+128, ... (PER = ceil(C/1024) of
// them).  The whole row payload then fits in a few dozen registers per thread (the earlier one-warp-per-row form needed
// 160+ registers at C = 5120 and ptxas spilled the row to local memory: 570 GB/s), every global access is one coalesced
// 16/32-byte piece per thread, and each row costs two block reductions through shared memory.
__device__ __forceinline__ float2 block_sum2(float2 v, float2* red /*[2][4]*/, int phase) {
  v.x = warp_sum(v.x);
  v.y = warp_sum(v.y);
  float2* r = red + (phase & 1) * 4;
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  const float2 a = r[0], b = r[1], c = r[2], d = r[3];
  return make_float2((a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y));
}

// =============================================================================================
// LayerNorm backward:  y = ((x-mean)*rstd [*gamma + beta]) [* (1+scale) + shift]
//   g = dy * (1+scale) * gamma ; dx = rstd * (g - mean(g) - xhat * mean(g*xhat)) ; dx_accum += dx
// =============================================================================================
template <int NCH>
__global__ void __launch_bounds__(128) ln_mod_bwd_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                         const float* __restrict__ scale, const float* __restrict__ gamma,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         float* __restrict__ dx, int64_t rows) {
  constexpr int C = NCH * 256, NPIECE = NCH * 32, PER = (NPIECE + 127) / 128;
  __shared__ float2 red[8];
  int phase = 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mu = mean[row], rs = rstd[row];
    const float* xr = x + row * C;
    const __nv_bfloat16* dyr = dy + row * C;
    float g[PER][8], xh[PER][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int pc = threadIdx.x + 128 * i;
      if (NPIECE % 128 != 0 && pc >= NPIECE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[i][j] = xh[i][j] = 0.f;
        continue;
      }
      const int c0 = 8 * pc;
      const uint4 gr = ldg_nc_v4(dyr + c0);
      const float4 a = __ldcs(reinterpret_cast<const float4*>(xr + c0)), b = __ldcs(reinterpret_cast<const float4*>(xr + c0) + 1);
      const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      const uint32_t u[4] = {gr.x, gr.y, gr.z, gr.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { g[i][2 * j] = bf16lo(u[j]); g[i][2 * j + 1] = bf16hi(u[j]); }
      if (scale) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(scale + c0)), q = __ldg(reinterpret_cast<const float4*>(scale + c0) + 1);
        const float sv[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) g[i][j] *= 1.f + sv[j];
      }
      if (gamma) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(gamma + c0)), q = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
        const float gv[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) g[i][j] *= gv[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xv[j] - mu) * rs;
        s1 += g[i][j];
        s2 += g[i][j] * xh[i][j];
      }
    }
    const float2 tot = block_sum2(make_float2(s1, s2), red, phase++);
    const float c1 = tot.x * (1.0f / C), c2 = tot.y * (1.0f / C);
    float* dxr = dx + row * C;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int pc = threadIdx.x + 128 * i;
      if (NPIECE % 128 != 0 && pc >= NPIECE) continue;
      const int c0 = 8 * pc;
      float4 d0 = *reinterpret_cast<const float4*>(dxr + c0), d1 = *(reinterpret_cast<const float4*>(dxr + c0) + 1);
      float o[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += rs * (g[i][j] - c1 - xh[i][j] * c2);
      *reinterpret_cast<float4*>(dxr + c0) = make_float4(o[0], o[1], o[2], o[3]);
      *(reinterpret_cast<float4*>(dxr + c0) + 1) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// =============================================================================================
// RMSNorm(+RoPE) backward.  forward: n = bf16(x*rstd); t = n*w; y = rope(t)
//   dt = rope^T(dy); gw = dt*n (-> column sum = dw); dn = dt*w; dx = rstd*(dn - xhat*mean(dn*xhat)), xhat = x*rstd
// =============================================================================================
template <int NCH>
__global__ void __launch_bounds__(128) rmsnorm_rope_bwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                               const float* __restrict__ w, const float* __restrict__ cos_tab,
                                                               const float* __restrict__ sin_tab, const __nv_bfloat16* __restrict__ dy,
                                                               int64_t lddy, const float* __restrict__ rstd,
                                                               __nv_bfloat16* __restrict__ dx, int64_t lddx,
                                                               __nv_bfloat16* __restrict__ gw, int64_t ldgw, int64_t rows,
                                                               int64_t n_rot, int64_t pos0) {
  constexpr int C = NCH * 256, NPIECE = NCH * 32, PER = (NPIECE + 127) / 128;
  __shared__ float2 red[8];
  int phase = 0;
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float rs = rstd[row];
    const bool rot = cos_tab != nullptr && row < n_rot;
    const float* cr = rot ? cos_tab + (pos0 + row) * 64 : nullptr;
    const float* sr = rot ? sin_tab + (pos0 + row) * 64 : nullptr;
    float dn[PER][8], xh[PER][8];      // dn = un-rotated dy (dt); xh = x * rstd
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int pc = threadIdx.x + 128 * i;
      if (NPIECE % 128 != 0 && pc >= NPIECE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dn[i][j] = xh[i][j] = 0.f;
        continue;
      }
      const int c0 = 8 * pc;
      const uint4 xraw = ldg_nc_v4(x + row * ldx + c0), graw = ldg_nc_v4(dy + row * lddy + c0);
      const uint32_t ux[4] = {xraw.x, xraw.y, xraw.z, xraw.w};
      const uint32_t ug[4] = {graw.x, graw.y, graw.z, graw.w};
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c0)), w1 = __ldg(reinterpret_cast<const float4*>(w + c0) + 1);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float cc[4] = {1.f, 1.f, 1.f, 1.f}, sn[4] = {0.f, 0.f, 0.f, 0.f};
      if (rot) {
        const int j0 = (c0 & 127) >> 1;
        const float4 cv = __ldg(reinterpret_cast<const float4*>(cr + j0)), sv = __ldg(reinterpret_cast<const float4*>(sr + j0));
        cc[0] = cv.x; cc[1] = cv.y; cc[2] = cv.z; cc[3] = cv.w;
        sn[0] = sv.x; sn[1] = sv.y; sn[2] = sv.z; sn[3] = sv.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gr = bf16lo(ug[j]), gi = bf16hi(ug[j]);
        dn[i][2 * j] = gr * cc[j] + gi * sn[j];
        dn[i][2 * j + 1] = -gr * sn[j] + gi * cc[j];
        xh[i][2 * j] = bf16lo(ux[j]) * rs;
        xh[i][2 * j + 1] = bf16hi(ux[j]) * rs;
        acc += dn[i][2 * j] * wv[2 * j] * xh[i][2 * j] + dn[i][2 * j + 1] * wv[2 * j + 1] * xh[i][2 * j + 1];
      }
    }
    const float c2 = block_sum2(make_float2(acc, 0.f), red, phase++).x * (1.0f / C);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int pc = threadIdx.x + 128 * i;
      if (NPIECE % 128 != 0 && pc >= NPIECE) continue;
      const int c0 = 8 * pc;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c0)), w1 = __ldg(reinterpret_cast<const float4*>(w + c0) + 1);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float od[8], og[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        og[j] = dn[i][j] * bf16_round(xh[i][j]);
        od[j] = rs * (dn[i][j] * wv[j] - xh[i][j] * c2);
      }
      uint4 o;
      o.x = pack_bf16x2(od[0], od[1]); o.y = pack_bf16x2(od[2], od[3]);
      o.z = pack_bf16x2(od[4], od[5]); o.w = pack_bf16x2(od[6], od[7]);
      stg_v4(dx + row * lddx + c0, o);
      if (gw) {
        o.x = pack_bf16x2(og[0], og[1]); o.y = pack_bf16x2(og[2], og[3]);
        o.z = pack_bf16x2(og[4], og[5]); o.w = pack_bf16x2(og[6], og[7]);
        stg_v4(gw + row * ldgw + c0, o);
      }
    }
  }
}

// =============================================================================================
// column (token-dimension) reductions.  grid = (ceil(N/256), nparts); thread = 2 adjacent columns.
// =============================================================================================
// part[chunk, n] = sum_rows a[row, n]
__global__ void __launch_bounds__(128) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda, float* __restrict__ part,
                                                          int64_t rows, int N) {
  const int col = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (col >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * COL_ROWS, r1 = min(rows, r0 + COL_ROWS);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(a + r * lda + col);
    s0 += bf16lo(u);
    s1 += bf16hi(u);
  }
  *reinterpret_cast<float2*>(part + (int64_t)blockIdx.y * N + col) = make_float2(s0, s1);
}

// part1 = sum dy ; part2 = sum dy * (x - mean) * rstd     (modulation shift/scale or affine beta/gamma grads)
__global__ void __launch_bounds__(128) colsum_ln_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        float* __restrict__ part1, float* __restrict__ part2, int64_t rows, int N) {
  const int col = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (col >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * COL_ROWS, r1 = min(rows, r0 + COL_ROWS);
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(dy + r * N + col);
    const float2 xv = *reinterpret_cast<const float2*>(x + r * N + col);
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    const float g0 = bf16lo(u), g1 = bf16hi(u);
    a0 += g0; a1 += g1;
    b0 += g0 * (xv.x - mu) * rs; b1 += g1 * (xv.y - mu) * rs;
  }
  *reinterpret_cast<float2*>(part1 + (int64_t)blockIdx.y * N + col) = make_float2(a0, a1);
  *reinterpret_cast<float2*>(part2 + (int64_t)blockIdx.y * N + col) = make_float2(b0, b1);
}

// gated-residual backward: dy_out = bf16(dx * gate) ; part = sum dx * y      (x_out = x_in + gate * y)
__global__ void __launch_bounds__(128) gate_bwd_kernel(const float* __restrict__ dx, const __nv_bfloat16* __restrict__ y,
                                                       const float* __restrict__ gate, __nv_bfloat16* __restrict__ dy_out,
                                                       float* __restrict__ part, int64_t rows, int N) {
  const int col = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (col >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * COL_ROWS, r1 = min(rows, r0 + COL_ROWS);
  const float2 gt = gate ? *reinterpret_cast<const float2*>(gate + col) : make_float2(1.f, 1.f);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
  for (int64_t r = r0; r < r1; ++r) {
    const float2 d = *reinterpret_cast<const float2*>(dx + r * N + col);
    if (y) {
      const uint32_t u = *reinterpret_cast<const uint32_t*>(y + r * N + col);
      s0 += d.x * bf16lo(u);
      s1 += d.y * bf16hi(u);
    }
    *reinterpret_cast<uint32_t*>(dy_out + r * N + col) = pack_bf16x2(d.x * gt.x, d.y * gt.y);
  }
  if (part) *reinterpret_cast<float2*>(part + (int64_t)blockIdx.y * N + col) = make_float2(s0, s1);
}

template <typename F>
static int dispatch_nch_b(int C, F&& f) {
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

static inline int row_grid_b(int64_t rows) {   // one CTA per row, grid-stride; up to 16 resident CTAs of 4 warps per SM
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(rows < cap ? (rows > 0 ? rows : 1) : cap);
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace prfl

using namespace prfl;

extern "C" {

int prfl_colsum_parts(int64_t rows) { return (int)((rows + COL_ROWS - 1) / COL_ROWS); }

int prfl_ln_mod_bwd(const float* x, const void* dy_bf16, const float* scale, const float* gamma, const float* mean,
                    const float* rstd, float* dx_accum, float* dshift_part, float* dscale_part, int64_t rows, int C,
                    prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(rows > 0 && C > 0 && C % 256 == 0 && mean && rstd && dx_accum, PRFL_E_SHAPE, "ln_mod_bwd: rows=%lld C=%d", (long long)rows, C);
  PRFL_REQUIRE((dshift_part == nullptr) == (dscale_part == nullptr), PRFL_E_SHAPE, "ln_mod_bwd: partial buffers come in pairs");
  PRFL_REQUIRE(al16(x) && al16(dy_bf16) && al16(scale) && al16(gamma) && al16(dx_accum), PRFL_E_ALIGN, "ln_mod_bwd: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  if (dshift_part) {
    dim3 grid((C / 2 + 127) / 128, prfl_colsum_parts(rows));
    colsum_ln_kernel<<<grid, 128, 0, st>>>((const __nv_bfloat16*)dy_bf16, x, mean, rstd, dshift_part, dscale_part, rows, C);
    count_launch();
  }
  return dispatch_nch_b(C, [&](auto nch) {
    constexpr int NCH = decltype(nch)::value;
    ln_mod_bwd_kernel<NCH><<<row_grid_b(rows), 128, 0, st>>>(x, (const __nv_bfloat16*)dy_bf16, scale, gamma, mean, rstd, dx_accum, rows);
    count_launch();
    PRFL_LAUNCH_CHECK("ln_mod_bwd");
    return PRFL_OK;
  });
}

int prfl_rmsnorm_rope_bwd(const void* x_bf16, int64_t ldx, const float* w, const float* cos_tab, const float* sin_tab,
                          const void* dy_bf16, int64_t lddy, const float* rstd, void* dx_bf16, int64_t lddx, void* gw_bf16,
                          int64_t ldgw, int64_t rows, int C, int64_t n_rot, int64_t pos0, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(rows > 0 && C > 0 && C % 256 == 0 && rstd, PRFL_E_SHAPE, "rmsnorm_rope_bwd: rows=%lld C=%d", (long long)rows, C);
  PRFL_REQUIRE(al16(x_bf16) && al16(dy_bf16) && al16(dx_bf16) && al16(gw_bf16) && al16(w) && ldx % 8 == 0 && lddy % 8 == 0 &&
                   lddx % 8 == 0 && ldgw % 8 == 0,
               PRFL_E_ALIGN, "rmsnorm_rope_bwd: alignment");
  return dispatch_nch_b(C, [&](auto nch) {
    constexpr int NCH = decltype(nch)::value;
    rmsnorm_rope_bwd_kernel<NCH><<<row_grid_b(rows), 128, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x_bf16, ldx, w, cos_tab, sin_tab, (const __nv_bfloat16*)dy_bf16, lddy, rstd, (__nv_bfloat16*)dx_bf16,
        lddx, (__nv_bfloat16*)gw_bf16, ldgw, rows, n_rot, pos0);
    count_launch();
    PRFL_LAUNCH_CHECK("rmsnorm_rope_bwd");
    return PRFL_OK;
  });
}

int prfl_colsum_bf16(const void* a_bf16, int64_t lda, float* part, int64_t rows, int N, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(rows > 0 && N > 0 && N % 2 == 0 && lda >= N && lda % 2 == 0, PRFL_E_SHAPE, "colsum_bf16: rows=%lld N=%d", (long long)rows, N);
  dim3 grid((N / 2 + 127) / 128, prfl_colsum_parts(rows));
  colsum_bf16_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a_bf16, lda, part, rows, N);
  count_launch();
  PRFL_LAUNCH_CHECK("colsum_bf16");
  return PRFL_OK;
}

int prfl_gate_bwd(const float* dx, const void* y_bf16, const float* gate, void* dy_bf16, float* dgate_part, int64_t rows, int N,
                  prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(rows > 0 && N > 0 && N % 2 == 0 && dx && dy_bf16, PRFL_E_SHAPE, "gate_bwd: rows=%lld N=%d", (long long)rows, N);
  PRFL_REQUIRE((dgate_part == nullptr) || (y_bf16 != nullptr), PRFL_E_SHAPE, "gate_bwd: dgate needs y");
  dim3 grid((N / 2 + 127) / 128, prfl_colsum_parts(rows));
  gate_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(dx, (const __nv_bfloat16*)y_bf16, gate, (__nv_bfloat16*)dy_bf16, dgate_part,
                                                          rows, N);
  count_launch();
  PRFL_LAUNCH_CHECK("gate_bwd");
  return PRFL_OK;
}

}  // extern "C"
