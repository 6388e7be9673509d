// KA -- gradient post-processing and Adam (train_boxpose.py:262-288; flax.optim.Adam).
// Two multi-tensor passes over the flat parameter blob: sanitise + clip + sum of squares, then the global-norm
// scale folded into the Adam update.  HBM-bound: 8 B + 28 B per parameter.
#include "common.cuh"

namespace durf {

// jnp.nan_to_num(g, posinf=0.0): NaN -> 0, +inf -> 0, -inf -> -FLT_MAX; then clip to +-max_val (if > 0).
__global__ void __launch_bounds__(256)
grad_sanitize_kernel(int64_t n, float* __restrict__ g, float max_val, float scale, float* __restrict__ sumsq) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = g[i] * scale;
    if (v != v) v = 0.f;
    else if (isinf(v)) v = v > 0.f ? 0.f : -3.402823466e+38f;
    if (max_val > 0.f) v = fminf(fmaxf(v, -max_val), max_val);
    g[i] = v;
    acc += v * v;
  }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(sumsq, t);
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            const float* __restrict__ sumsq, float max_norm, float lr, float b1, float b2, float omb1, float omb2, float eps,
            float c1, float c2) {
  float mult = 1.f;
  if (max_norm > 0.f) mult = fminf(1.f, max_norm / (1e-7f + sqrtf(*sumsq)));   // train_boxpose.py:283-285
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = mult * g[i];
    const float mi = omb1 * gi + b1 * m[i];
    const float vi = omb2 * (gi * gi) + b2 * v[i];
    m[i] = mi;
    v[i] = vi;
    const float mhat = mi / c1;
    const float denom = sqrtf(vi / c2) + eps;
    p[i] = p[i] - lr * mhat / denom;
  }
}

// Device-scalar variant: lr and the step counter live in device memory.
__global__ void __launch_bounds__(256)
adam_dev_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                const float* __restrict__ sumsq, float max_norm, const float* __restrict__ lr_dev, const int32_t* __restrict__ step_dev,
                double beta1, double beta2, float eps) {
  float mult = 1.f;
  if (max_norm > 0.f) mult = fminf(1.f, max_norm / (1e-7f + sqrtf(*sumsq)));
  const float lr = *lr_dev;
  const double t = (double)(*step_dev) + 1.0;
  const float c1 = (float)(1.0 - pow(beta1, t)), c2 = (float)(1.0 - pow(beta2, t));
  const float b1 = (float)beta1, b2 = (float)beta2;
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = mult * g[i];
    const float mi = omb1 * gi + b1 * m[i];
    const float vi = omb2 * (gi * gi) + b2 * v[i];
    m[i] = mi;
    v[i] = vi;
    const float mhat = mi / c1;
    const float denom = sqrtf(vi / c2) + eps;
    p[i] = p[i] - lr * mhat / denom;
  }
}

__global__ void advance_step_kernel(int32_t* step_dev) { *step_dev += 1; }

}  // namespace durf

using namespace durf;

extern "C" int durf_adam_step_dev(durf_stream_t stream, int64_t n, float* params, const float* grad, float* m, float* v,
                                  const float* sumsq, float max_norm, const float* lr_dev, int32_t* step_dev, int32_t advance_step,
                                  double beta1, double beta2, double eps) {
  DURF_REQUIRE(n >= 0 && params && grad && m && v && sumsq && lr_dev && step_dev, DURF_E_INVALID, "durf_adam_step_dev: bad argument");
  if (n > 0) {
    const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    adam_dev_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, params, grad, m, v, sumsq, max_norm, lr_dev, step_dev, beta1, beta2, (float)eps);
    DURF_CHECK_LAUNCH("durf_adam_step_dev");
  }
  if (advance_step) {
    advance_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    DURF_CHECK_LAUNCH("durf_adam_step_dev(advance)");
  }
  return DURF_OK;
}

extern "C" int durf_grad_sanitize(durf_stream_t stream, int64_t n, float* grad, float max_val, float grad_scale, float* sumsq) {
  DURF_REQUIRE(n >= 0 && grad && sumsq, DURF_E_INVALID, "durf_grad_sanitize: bad argument");
  if (n == 0) return DURF_OK;
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  grad_sanitize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, grad, max_val, grad_scale, sumsq);
  DURF_CHECK_LAUNCH("durf_grad_sanitize");
  return DURF_OK;
}

extern "C" int durf_adam_step(durf_stream_t stream, int64_t n, float* params, const float* grad, float* m, float* v,
                              const float* sumsq, float max_norm, float lr, double beta1, double beta2, double eps, int32_t step) {
  DURF_REQUIRE(n >= 0 && params && grad && m && v && sumsq && step >= 0, DURF_E_INVALID, "durf_adam_step: bad argument");
  if (n == 0) return DURF_OK;
  const double t = (double)step + 1.0;
  const float c1 = (float)(1.0 - pow(beta1, t)), c2 = (float)(1.0 - pow(beta2, t));
  // (1. - beta) is a Python double in flax.optim.Adam and only then meets the fp32 arrays
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, params, grad, m, v, sumsq, max_norm, lr, (float)beta1, (float)beta2, omb1, omb2, (float)eps, c1, c2);
  DURF_CHECK_LAUNCH("durf_adam_step");
  return DURF_OK;
}
