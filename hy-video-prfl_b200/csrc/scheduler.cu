// PRFL chain glue: one FlowUniPC multistep update as ONE memory-bound kernel.
// The reference (diffusers_lite/wan/utils/fm_solvers_unipc.py:655-739) runs convert_model_output, the UniC corrector and
// the UniP predictor as ~25 ATen elementwise launches over latent-sized fp32 tensors with several temporaries; all three
// are linear in (sample, model_output, last_sample, previous x0 predictions), so the host folds the step's scalars into
// two coefficient vectors and this kernel reads each input once and writes x0, the corrected sample and the next sample:
//   v         = v_cond, or v_uncond + guide * (v_cond - v_uncond) when a second model output is given: classifier-free
//               guidance of the sampling loop (text2video.py:295-296) folded in
//   x0        = sample - sigma * v                                              (:318-321)
//   corrected = c[0] last + c[1] x0 + c[2] h0 + c[3] h1 + c[4] h2               (:486-626, only when the corrector runs)
//   prev      = p[0] (corrected | sample) + p[1] x0 + p[2] h0 + p[3] h1 + p[4] h2   (:350-484)
// Algorithmic bytes: 4 B x (inputs read + 3 outputs) per latent element.
#include "common.cuh"

namespace prfl {

struct UniPCCoef {
  float sigma;
  float guide;
  float c[5];
  float p[5];
  int use_corrector;
};

template <bool VEC>
__global__ void __launch_bounds__(256) unipc_step_kernel(const float* __restrict__ sample, const float* __restrict__ v,
                                                         const float* __restrict__ v_uncond, const float* __restrict__ last, const float* __restrict__ h0,
                                                         const float* __restrict__ h1, const float* __restrict__ h2,
                                                         float* __restrict__ x0_out, float* __restrict__ corr_out,
                                                         float* __restrict__ prev_out, int64_t n, const UniPCCoef k) {
  using T = typename std::conditional<VEC, float4, float>::type;
  constexpr int W = VEC ? 4 : 1;
  const int64_t nv = n / W;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float s[4], m[4], l[4] = {0, 0, 0, 0}, a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
    auto ld = [&](const float* p, float* r) {
      if (VEC) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(p) + i);
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
      } else {
        r[0] = p[i];
      }
    };
    auto st = [&](float* p, const float* r) {
      if (VEC) reinterpret_cast<float4*>(p)[i] = make_float4(r[0], r[1], r[2], r[3]);
      else p[i] = r[0];
    };
    ld(sample, s);
    ld(v, m);
    if (v_uncond) {
      float u[4];
      ld(v_uncond, u);
#pragma unroll
      for (int j = 0; j < W; ++j) m[j] = u[j] + k.guide * (m[j] - u[j]);
    }
    if (k.use_corrector) ld(last, l);
    if (h0) ld(h0, a0);
    if (h1) ld(h1, a1);
    if (h2) ld(h2, a2);
    float x0[4], xc[4], xp[4];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      x0[j] = fmaf(-k.sigma, m[j], s[j]);
      xc[j] = s[j];
      if (k.use_corrector) xc[j] = k.c[0] * l[j] + k.c[1] * x0[j] + k.c[2] * a0[j] + k.c[3] * a1[j] + k.c[4] * a2[j];
      xp[j] = k.p[0] * xc[j] + k.p[1] * x0[j] + k.p[2] * a0[j] + k.p[3] * a1[j] + k.p[4] * a2[j];
    }
    st(x0_out, x0);
    if (k.use_corrector) st(corr_out, xc);
    st(prev_out, xp);
  }
}

// ya = a * g ; yb = b * g (yb optional): the backward of the step above with respect to (model_output, sample)
__global__ void __launch_bounds__(256) scale2_kernel(const float* __restrict__ g, float a, float* __restrict__ ya, float b,
                                                     float* __restrict__ yb, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float t = g[i];
    ya[i] = a * t;
    if (yb) yb[i] = b * t;
  }
}

}  // namespace prfl

using namespace prfl;

extern "C" int prfl_unipc_step(const float* sample, const float* model_output, const float* model_output_uncond, float guide_scale,
                               const float* last_sample, const float* hist0, const float* hist1, const float* hist2, float sigma,
                               const float* corr_coef,
                               const float* pred_coef, float* x0_out, float* corrected_out, float* prev_out, int64_t n,
                               prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(n > 0 && sample && model_output && pred_coef && x0_out && prev_out, PRFL_E_SHAPE, "unipc_step: n=%lld", (long long)n);
  PRFL_REQUIRE(!corr_coef || (last_sample && corrected_out), PRFL_E_SHAPE, "unipc_step: corrector needs last_sample and corrected_out");
  UniPCCoef k;
  k.sigma = sigma;
  k.guide = guide_scale;
  k.use_corrector = corr_coef != nullptr;
  for (int i = 0; i < 5; ++i) {
    k.c[i] = corr_coef ? corr_coef[i] : 0.f;
    k.p[i] = pred_coef[i];
  }
  // a missing history tensor must carry a zero coefficient
  const float* hs[3] = {hist0, hist1, hist2};
  for (int i = 0; i < 3; ++i)
    PRFL_REQUIRE(hs[i] || (k.c[2 + i] == 0.f && k.p[2 + i] == 0.f), PRFL_E_SHAPE, "unipc_step: hist%d is NULL but has a coefficient", i);
  uintptr_t al = reinterpret_cast<uintptr_t>(sample) | reinterpret_cast<uintptr_t>(model_output) | reinterpret_cast<uintptr_t>(model_output_uncond) |
                 reinterpret_cast<uintptr_t>(last_sample) |
                 reinterpret_cast<uintptr_t>(hist0) | reinterpret_cast<uintptr_t>(hist1) | reinterpret_cast<uintptr_t>(hist2) |
                 reinterpret_cast<uintptr_t>(x0_out) | reinterpret_cast<uintptr_t>(corrected_out) | reinterpret_cast<uintptr_t>(prev_out);
  const bool vec = (al & 15) == 0 && (n & 3) == 0;
  const int64_t work = vec ? n / 4 : n;
  int64_t blocks = (work + 255) / 256, cap = (int64_t)sm_count() * 8;
  const int grid = (int)(blocks < cap ? blocks : cap);
  if (vec)
    unipc_step_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(sample, model_output, model_output_uncond, last_sample, hist0, hist1, hist2, x0_out,
                                                                   corrected_out, prev_out, n, k);
  else
    unipc_step_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(sample, model_output, model_output_uncond, last_sample, hist0, hist1, hist2, x0_out,
                                                                    corrected_out, prev_out, n, k);
  count_launch();
  PRFL_LAUNCH_CHECK("unipc_step");
  return PRFL_OK;
}

extern "C" int prfl_scale2_f32(const float* g, float a, float* ya, float b, float* yb, int64_t n, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(n > 0 && g && ya, PRFL_E_SHAPE, "scale2: n=%lld", (long long)n);
  int64_t blocks = (n + 255) / 256, cap = (int64_t)sm_count() * 8;
  scale2_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(g, a, ya, b, yb, n);
  count_launch();
  PRFL_LAUNCH_CHECK("scale2");
  return PRFL_OK;
}
