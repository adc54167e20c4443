// Sharded AdamW update on fp32 master shards (the optimizer.step of train_prfl.py:825-830 under the
// FSDP role of §8 a17) as one memory-bound kernel per FSDP unit instead of ~10 ATen elementwise launches:
//   g = clip * grad ; w *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
//   w -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)                       (torch.optim.AdamW, decoupled decay)
// reads g, w, m, v and writes w, m, v once: 28 B per parameter (+2 B when the bf16 compute copy of the updated weight is
// written in the same pass: the resident operand the GEMMs read, all-gathered in bf16 afterwards).  `clip` is a DEVICE scalar (the clip_grad_norm_
// coefficient) so the step needs no host synchronisation.  A second kernel accumulates sum(g^2) for that norm.
#include "common.cuh"

namespace prfl {

struct AdamWArgs {
  float lr, b1, b2, eps, wd, bc1, bc2_sqrt;   // bc1 = 1 - b1^t, bc2_sqrt = sqrt(1 - b2^t)
};

__global__ void __launch_bounds__(256) adamw_kernel(const float* __restrict__ g, float* __restrict__ w, float* __restrict__ m,
                                                    float* __restrict__ v, const float* __restrict__ clip, __nv_bfloat16* __restrict__ wb,
                                                    int64_t n, const AdamWArgs a) {
  const float c = clip ? *clip : 1.0f;
  const float decay = 1.0f - a.lr * a.wd, step = a.lr / a.bc1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g4 = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 w4 = reinterpret_cast<float4*>(w)[i], m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    float gg[4] = {g4.x * c, g4.y * c, g4.z * c, g4.w * c};
    float ww[4] = {w4.x, w4.y, w4.z, w4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ww[j] *= decay;
      mm[j] = a.b1 * mm[j] + (1.0f - a.b1) * gg[j];
      vv[j] = a.b2 * vv[j] + (1.0f - a.b2) * gg[j] * gg[j];
      ww[j] -= step * mm[j] / (sqrtf(vv[j]) / a.bc2_sqrt + a.eps);
    }
    reinterpret_cast<float4*>(w)[i] = make_float4(ww[0], ww[1], ww[2], ww[3]);
    if (wb) reinterpret_cast<uint2*>(wb)[i] = make_uint2(pack_bf16x2(ww[0], ww[1]), pack_bf16x2(ww[2], ww[3]));
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gj = g[i] * c;
    float wj = w[i] * decay;
    const float mj = a.b1 * m[i] + (1.0f - a.b1) * gj, vj = a.b2 * v[i] + (1.0f - a.b2) * gj * gj;
    wj -= step * mj / (sqrtf(vj) / a.bc2_sqrt + a.eps);
    w[i] = wj; m[i] = mj; v[i] = vj;
    if (wb) wb[i] = __float2bfloat16_rn(wj);
  }
}

// acc[0] += sum x^2 (double accumulation across CTAs keeps the 1.7e9-element norm accurate)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ acc) {
  __shared__ float red[8];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float s = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(x) + i);
    s += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += x[i] * x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(acc, (double)t);
  }
}

}  // namespace prfl

using namespace prfl;

extern "C" int prfl_adamw_step(const float* grad, float* master, float* exp_avg, float* exp_avg_sq, const float* clip_coef_dev,
                               void* master_bf16_out, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                               int step, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(n > 0 && grad && master && exp_avg && exp_avg_sq && step >= 1, PRFL_E_SHAPE, "adamw_step: n=%lld step=%d", (long long)n, step);
  PRFL_REQUIRE(((reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(master) | reinterpret_cast<uintptr_t>(exp_avg) |
                 reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 && (reinterpret_cast<uintptr_t>(master_bf16_out) & 7) == 0,
               PRFL_E_ALIGN, "adamw_step: fp32 pointers must be 16-byte aligned (bf16 output 8-byte)");
  AdamWArgs a;
  a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  int64_t blocks = (n / 4 + 255) / 256, cap = (int64_t)sm_count() * 8;
  blocks = blocks < 1 ? 1 : blocks;
  adamw_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(grad, master, exp_avg, exp_avg_sq, clip_coef_dev,
                                                                                      (__nv_bfloat16*)master_bf16_out, n, a);
  count_launch();
  PRFL_LAUNCH_CHECK("adamw_step");
  return PRFL_OK;
}

extern "C" int prfl_sumsq_f32(const float* x, int64_t n, double* acc, prfl_stream_t stream) {
  PRFL_CHECK_ARCH();
  PRFL_REQUIRE(n > 0 && x && acc && (reinterpret_cast<uintptr_t>(x) & 15) == 0, PRFL_E_SHAPE, "sumsq: n=%lld", (long long)n);
  int64_t blocks = (n / 4 + 255) / 256, cap = (int64_t)sm_count() * 8;
  blocks = blocks < 1 ? 1 : blocks;
  sumsq_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(x, n, acc);
  count_launch();
  PRFL_LAUNCH_CHECK("sumsq");
  return PRFL_OK;
}
