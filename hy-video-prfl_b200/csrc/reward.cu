// PAVRM reward head, single learnable query (network.py:44-110) as two streaming passes over the
// fp32 features instead of the reference's [L, C] x [C, 2C] in-projection GEMM + hd-640 attention:
//   pass 1  scores[l, h] = x[l, :] . wk_eff[h, :]        (wk_eff = Wk_h^T q_h / sqrt(hd), staged in smem)
//   pass 2  pooled[h, :] = sum_l softmax_l(scores[:, h]) x[l, :]
// Both are HBM-bound: 4*C bytes per token per pass (SURVEY.md §8d "reward single-query attention").
#include <math.h>

#include "common.cuh"

namespace prfl {

constexpr int SQ_MAX_HEADS = 8;

template <int NCH, int NH>
__global__ void __launch_bounds__(256, 1)
sq_scores_kernel(const float* __restrict__ x, const float* __restrict__ wk, float* __restrict__ scores, int64_t L) {
  constexpr int C = NCH * 256;
  extern __shared__ float s_wk[];  // [NH][C]
  for (int i = threadIdx.x; i < NH * C / 4; i += blockDim.x)
    reinterpret_cast<float4*>(s_wk)[i] = __ldg(reinterpret_cast<const float4*>(wk) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp_global; row < L; row += nwarps) {
    const float* xr = x + row * C;
    float4 v[NCH][2];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float4* p = reinterpret_cast<const float4*>(xr + 8 * (lane + 32 * i));
      v[i][0] = __ldg(p);
      v[i][1] = __ldg(p + 1);
    }
    float acc[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const float4* wp = reinterpret_cast<const float4*>(s_wk + h * C + 8 * (lane + 32 * i));
        float4 w0 = wp[0], w1 = wp[1];
        a += v[i][0].x * w0.x + v[i][0].y * w0.y + v[i][0].z * w0.z + v[i][0].w * w0.w;
        a += v[i][1].x * w1.x + v[i][1].y * w1.y + v[i][1].z * w1.z + v[i][1].w * w1.w;
      }
      acc[h] = warp_sum(a);
    }
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < NH; ++h) scores[row * NH + h] = acc[h];
    }
  }
}

// one CTA per head: max and sum(exp(s - max)) over L
__global__ void sq_stats_kernel(const float* __restrict__ scores, float* __restrict__ stats, int64_t L, int NH) {
  __shared__ float red[32];
  const int h = blockIdx.x;
  float m = -INFINITY;
  for (int64_t l = threadIdx.x; l < L; l += blockDim.x) m = fmaxf(m, scores[l * NH + h]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    t = warp_max(t);
    if (threadIdx.x == 0) red[0] = t;
  }
  __syncthreads();
  m = red[0];
  __syncthreads();
  float s = 0.f;
  for (int64_t l = threadIdx.x; l < L; l += blockDim.x) s += __expf(scores[l * NH + h] - m);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      stats[h] = m;
      stats[NH + h] = t;
    }
  }
}

// pooled[h, c] += sum over this CTA's rows of p[l, h] * x[l, c]; thread owns columns tid + 256*i
template <int NCH, int NH>
__global__ void __launch_bounds__(256)
sq_pool_kernel(const float* __restrict__ x, const float* __restrict__ scores, const float* __restrict__ stats,
               float* __restrict__ pooled, int64_t L, int rows_per_cta) {
  constexpr int C = NCH * 256;
  __shared__ float s_p[64][NH];
  float mx[NH], inv[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    mx[h] = stats[h];
    inv[h] = 1.0f / stats[NH + h];
  }
  float acc[NH][NCH];
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[h][i] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(L, r0 + rows_per_cta);
  for (int64_t base = r0; base < r1; base += 64) {
    const int nrow = (int)min((int64_t)64, r1 - base);
    __syncthreads();
    for (int i = threadIdx.x; i < nrow * NH; i += blockDim.x) {
      const int h = i % NH;
      s_p[i / NH][h] = __expf(scores[base * NH + i] - mx[h]) * inv[h];
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < nrow; ++r) {
      const float* xr = x + (base + r) * C + threadIdx.x;
      float xv[NCH];
#pragma unroll
      for (int i = 0; i < NCH; ++i) xv[i] = __ldcs(xr + 256 * i);
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float pw = s_p[r][h];
#pragma unroll
        for (int i = 0; i < NCH; ++i) acc[h][i] = fmaf(pw, xv[i], acc[h][i]);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int i = 0; i < NCH; ++i) atomicAdd(pooled + h * C + threadIdx.x + 256 * i, acc[h][i]);
}

// dx[l, c] (+)= sum_h p[l,h] * dpooled[h,c] + ds[l,h] * wk[h,c],  ds[l,h] = p[l,h] * (x[l,:].dpooled[h,:] - pd[h])
// warp per row; dpooled and wk staged in smem ([2][NH][C] fp32 can exceed smem for C=5120, so they are read
// through L1/L2 instead: 2*NH*C*4 B = 320 KB, L2-resident).
template <int NCH, int NH>
__global__ void __launch_bounds__(128)
sq_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wk, const float* __restrict__ scores,
                   const float* __restrict__ stats, const float* __restrict__ pooled, const float* __restrict__ dpooled,
                   float* __restrict__ dx, float* __restrict__ ds_out, int64_t L, int accumulate) {
  constexpr int C = NCH * 256;
  const int lane = threadIdx.x & 31;
  // pd[h] = pooled[h,:] . dpooled[h,:]
  float pd[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a += pooled[h * C + c] * dpooled[h * C + c];
    pd[h] = warp_sum(a);
  }
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp_global; row < L; row += nwarps) {
    const float* xr = x + row * C;
    float4 v[NCH][2];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float4* p = reinterpret_cast<const float4*>(xr + 8 * (lane + 32 * i));
      v[i][0] = __ldg(p);
      v[i][1] = __ldg(p + 1);
    }
    float pr[NH], ds[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const float4* gp = reinterpret_cast<const float4*>(dpooled + h * C + 8 * (lane + 32 * i));
        float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
        a += v[i][0].x * g0.x + v[i][0].y * g0.y + v[i][0].z * g0.z + v[i][0].w * g0.w;
        a += v[i][1].x * g1.x + v[i][1].y * g1.y + v[i][1].z * g1.z + v[i][1].w * g1.w;
      }
      a = warp_sum(a);
      pr[h] = __expf(scores[row * NH + h] - stats[h]) / stats[NH + h];
      ds[h] = pr[h] * (a - pd[h]);
    }
    if (lane == 0 && ds_out) {
#pragma unroll
      for (int h = 0; h < NH; ++h) ds_out[row * NH + h] = ds[h];
    }
    float* dr = dx + row * C;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
      float4* op = reinterpret_cast<float4*>(dr + 8 * (lane + 32 * i));
      if (accumulate) {
        o0 = op[0];
        o1 = op[1];
      }
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float4* gp = reinterpret_cast<const float4*>(dpooled + h * C + 8 * (lane + 32 * i));
        const float4* wp = reinterpret_cast<const float4*>(wk + h * C + 8 * (lane + 32 * i));
        float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), w0 = __ldg(wp), w1 = __ldg(wp + 1);
        o0.x += pr[h] * g0.x + ds[h] * w0.x; o0.y += pr[h] * g0.y + ds[h] * w0.y;
        o0.z += pr[h] * g0.z + ds[h] * w0.z; o0.w += pr[h] * g0.w + ds[h] * w0.w;
        o1.x += pr[h] * g1.x + ds[h] * w1.x; o1.y += pr[h] * g1.y + ds[h] * w1.y;
        o1.z += pr[h] * g1.z + ds[h] * w1.z; o1.w += pr[h] * g1.w + ds[h] * w1.w;
      }
      op[0] = o0;
      op[1] = o1;
    }
  }
}

template <typename F>
static int dispatch_sq(int C, int NH, F&& f) {
  if (NH != 8) {
    set_error("sq_pool: only NH == 8 heads compiled (got %d)", NH);
    return PRFL_E_SHAPE;
  }
  switch (C / 256) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    case 4: return f(std::integral_constant<int, 4>{});
    case 6: return f(std::integral_constant<int, 6>{});
    case 8: return f(std::integral_constant<int, 8>{});
    case 20: return f(std::integral_constant<int, 20>{});
    default:
      set_error("sq_pool: unsupported C=%d", C);
      return PRFL_E_SHAPE;
  }
}

}  // namespace prfl

using namespace prfl;

extern "C" int prfl_sq_pool_fwd(const float* x, const float* wk_eff, float* scores, float* stats, float* pooled, int64_t L,
                                int C, int NH, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(L > 0 && C > 0 && C % 256 == 0, PRFL_E_SHAPE, "sq_pool_fwd: L=%lld C=%d", (long long)L, C);
  PRFL_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wk_eff)) & 15) == 0, PRFL_E_ALIGN, "sq_pool_fwd: alignment");
  cudaStream_t st = (cudaStream_t)stream;
  return dispatch_sq(C, NH, [&](auto nch) {
    constexpr int NCH = decltype(nch)::value;
    constexpr int NHc = 8;
    const int smem = NHc * NCH * 256 * 4;
    auto k1 = sq_scores_kernel<NCH, NHc>;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "sq_pool_fwd: smem attr");
    int64_t need = (L + 7) / 8;
    int grid1 = (int)(need < sm_count() ? need : sm_count());
    k1<<<grid1, 256, smem, st>>>(x, wk_eff, scores, L);
    sq_stats_kernel<<<NHc, 1024, 0, st>>>(scores, stats, L, NHc);
    e = cudaMemsetAsync(pooled, 0, sizeof(float) * NHc * C, st);
    if (e != cudaSuccess) return cuda_fail(e, "sq_pool_fwd: memset");
    int ctas = sm_count() * 2;
    int rows_per_cta = (int)((L + ctas - 1) / ctas);
    rows_per_cta = ((rows_per_cta + 63) / 64) * 64;
    int grid3 = (int)((L + rows_per_cta - 1) / rows_per_cta);
    sq_pool_kernel<NCH, NHc><<<grid3, 256, 0, st>>>(x, scores, stats, pooled, L, rows_per_cta);
    count_launch(3);
    PRFL_LAUNCH_CHECK("sq_pool_fwd");
    return PRFL_OK;
  });
}

extern "C" int prfl_sq_pool_bwd(const float* x, const float* wk_eff, const float* scores, const float* stats,
                                const float* pooled, const float* dpooled, float* dx, float* ds, int64_t L, int C,
                                int NH, int accumulate, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(L > 0 && C > 0 && C % 256 == 0, PRFL_E_SHAPE, "sq_pool_bwd: L=%lld C=%d", (long long)L, C);
  cudaStream_t st = (cudaStream_t)stream;
  return dispatch_sq(C, NH, [&](auto nch) {
    constexpr int NCH = decltype(nch)::value;
    int64_t need = (L + 3) / 4;
    int64_t cap = (int64_t)sm_count() * 4;
    sq_pool_bwd_kernel<NCH, 8><<<(int)(need < cap ? need : cap), 128, 0, st>>>(x, wk_eff, scores, stats, pooled, dpooled, dx, ds, L,
                                                                               accumulate);
    count_launch();
    PRFL_LAUNCH_CHECK("sq_pool_bwd");
    return PRFL_OK;
  });
}
