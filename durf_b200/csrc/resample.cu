// K4 -- hierarchical resampling: blur-pool of the coarse weights + inverse-CDF sampling.  One warp per ray.
// Replaces mip.resample_along_rays (mip.py:393-412) and math.sorted_piecewise_constant_pdf (math.py:222-284).
// The reference finds the interval of each sample with a dense [B, S, S] mask + max/min (math.py:270-280,
// O(N^2) per ray); here the CDF sits in shared memory and each sample does an O(log N) search for the last
// knot <= u, which selects the same interval because the CDF is non-decreasing.
// HBM-bound: 512 + 516 B in, 516 B out per ray.
#include "common.cuh"

namespace durf {

struct ResampleParams {
  int B, N, S_out, blur;
  const float* t_vals;
  const float* weights;
  const float* u_rand;
  float padding;
  float s_step;      // fl32(1 / (N+1))
  float s_jit;       // fl32(1/(N+1) - eps32): scale of the uniform jitter (math.py:257-260)
  float u_max;       // fl32(1 - eps32)
  float* out;
};

__global__ void __launch_bounds__(128)
resample_kernel(const ResampleParams p) {
  extern __shared__ float smem[];
  const int N = p.N, S = N + 1, SO = p.S_out;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * (blockDim.x >> 5) + warp;
  if (ray >= p.B) return;
  float* s_bins = smem + warp * (3 * S + 1);   // [S]
  float* s_cdf = s_bins + S;                   // [S]
  float* s_w = s_cdf + S;                      // [N] raw weights, then pdf
  const float* w = p.weights + (size_t)ray * N;
  for (int i = lane; i < S; i += 32) s_bins[i] = p.t_vals[(size_t)ray * S + i];
  for (int i = lane; i < N; i += 32) s_w[i] = w[i];
  __syncwarp();

  // blur-pool (mip.py:394-401): wmax_j = max(wp_j, wp_{j+1}) on the edge-padded row, wblur_i = (wmax_i + wmax_{i+1})/2
  float wb[4];
  float part = 0.f;
  {
    int c = 0;
    for (int i = lane; i < N; i += 32, ++c) {
      if (p.blur) {
        const float wl = s_w[max(i - 1, 0)], wc = s_w[i], wr = s_w[min(i + 1, N - 1)];
        const float m0 = fmaxf(wl, wc), m1 = fmaxf(wc, wr);
        wb[c] = 0.5f * (m0 + m1) + p.padding;
      } else {
        wb[c] = s_w[i];
      }
      part += wb[c];
    }
  }
  // math.py:237-245
  float wsum = warp_sum(part);
  const float pad = fmaxf(0.f, 1e-5f - wsum);
  wsum += pad;
  __syncwarp();
  {
    int c = 0;
    for (int i = lane; i < N; i += 32, ++c) s_w[i] = (wb[c] + pad / (float)N) / wsum;
  }
  __syncwarp();
  // cdf = [0, min(1, cumsum(pdf[:-1])), 1] (math.py:246-251).  Lane-contiguous blocks of ceil(N/32) entries.
  {
    const int per = (N + 31) / 32;
    const int lo = lane * per, hi = min(lo + per, N);
    float run = 0.f;
    for (int i = lo; i < hi; ++i) run += s_w[i];
    float acc = warp_scan_excl(run, lane);
    for (int i = lo; i < hi; ++i) {
      acc += s_w[i];
      if (i + 1 < N) s_cdf[i + 1] = fminf(1.f, acc);
    }
    if (lane == 0) { s_cdf[0] = 0.f; s_cdf[N] = 1.f; }
  }
  __syncwarp();

  for (int i = lane; i < SO; i += 32) {
    float u;
    if (p.u_rand) {
      u = (float)i * p.s_step + p.u_rand[(size_t)ray * SO + i] * p.s_jit;
      u = fminf(u, p.u_max);
    } else {
      u = (i == SO - 1) ? p.u_max : p.u_max * ((float)i / (float)(SO - 1));   // linspace(0, 1-eps, num_samples)
    }
    // last knot j with cdf[j] <= u  (cdf[0] = 0 <= u always)
    int lo = 0, hi = N;           // invariant: cdf[lo] <= u; answer in [lo, hi]
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_cdf[mid] <= u) lo = mid; else hi = mid - 1;
    }
    const int j1 = min(lo + 1, N);
    const float c0 = s_cdf[lo], c1 = s_cdf[j1];
    const float b0 = s_bins[lo], b1 = s_bins[j1];
    float t = nan_to_num((u - c0) / (c1 - c0));
    t = fminf(fmaxf(t, 0.f), 1.f);
    p.out[(size_t)ray * SO + i] = b0 + t * (b1 - b0);
  }
}

// Specialisation for the model's shape (N = 128 bins -> 129 samples): every lane owns 4 consecutive bins, all loops have
// fixed trip counts and the interval search is a branch-free 7-step bisection, ~3x fewer instructions than the generic
// kernel above (which stays for other N, e.g. math.sorted_piecewise_constant_pdf on arbitrary histograms).
__global__ void __launch_bounds__(128)
resample128_kernel(const ResampleParams p) {
  __shared__ float s_all[4][2 * 129 + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * 4 + warp;
  if (ray >= p.B) return;
  constexpr int N = 128, S = 129;
  float* s_bins = s_all[warp];
  float* s_cdf = s_bins + S;
  const float* tv = p.t_vals + (size_t)ray * S;
#pragma unroll
  for (int k = 0; k < 4; ++k) s_bins[lane + 32 * k] = tv[lane + 32 * k];      // coalesced: 128 contiguous bytes per instruction
  if (lane == 0) s_bins[N] = tv[N];
  const float4 w4 = *reinterpret_cast<const float4*>(p.weights + (size_t)ray * N + 4 * lane);
  float w[4] = {w4.x, w4.y, w4.z, w4.w};
  float wb[4];
  if (p.blur) {
    // blur-pool (mip.py:394-401) on the edge-padded row: neighbours across lanes by shuffle
    float wl = __shfl_up_sync(kFull, w[3], 1), wr = __shfl_down_sync(kFull, w[0], 1);
    if (lane == 0) wl = w[0];
    if (lane == 31) wr = w[3];
    const float e[6] = {wl, w[0], w[1], w[2], w[3], wr};
#pragma unroll
    for (int k = 0; k < 4; ++k) wb[k] = 0.5f * (fmaxf(e[k], e[k + 1]) + fmaxf(e[k + 1], e[k + 2])) + p.padding;
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) wb[k] = w[k];
  }
  // math.py:237-245
  float wsum = warp_sum((wb[0] + wb[1]) + (wb[2] + wb[3]));
  const float pad = fmaxf(0.f, 1e-5f - wsum);
  wsum += pad;
  float pdf[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) pdf[k] = (wb[k] + pad / (float)N) / wsum;
  // cdf = [0, min(1, cumsum(pdf[:-1])), 1] (math.py:246-251)
  float acc = warp_scan_excl((pdf[0] + pdf[1]) + (pdf[2] + pdf[3]), lane);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    acc += pdf[k];
    if (4 * lane + k + 1 < N) s_cdf[4 * lane + k + 1] = fminf(1.f, acc);
  }
  if (lane == 0) { s_cdf[0] = 0.f; s_cdf[N] = 1.f; }
  __syncwarp();
  float* out = p.out + (size_t)ray * S;
  const float* ur = p.u_rand ? p.u_rand + (size_t)ray * S : nullptr;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int i = lane + 32 * k;                    // strided ownership: loads of u_rand and stores of the samples coalesce
    if (k == 4 && lane != 0) break;                 // the 129th sample goes to lane 0
    float u;
    if (ur) u = fminf((float)i * p.s_step + ur[i] * p.s_jit, p.u_max);
    else u = (i == S - 1) ? p.u_max : p.u_max * ((float)i * 0.0078125f);      // i / 128: a power of two, the product is the exact quotient
    // last knot j in [0,127] with cdf[j] <= u (cdf[0] = 0 <= u, cdf[128] = 1 > u): branch-free bisection
    int j = 0;
#pragma unroll
    for (int step = 64; step >= 1; step >>= 1) j += (s_cdf[j + step] <= u) ? step : 0;
    const float c0 = s_cdf[j], c1 = s_cdf[j + 1], b0 = s_bins[j], b1 = s_bins[j + 1];
    float t = nan_to_num((u - c0) / (c1 - c0));
    t = fminf(fmaxf(t, 0.f), 1.f);
    out[i] = b0 + t * (b1 - b0);
  }
}

}  // namespace durf

using namespace durf;

extern "C" int durf_resample_fwd(durf_stream_t stream, int32_t B, int32_t N, const float* t_vals, const float* weights,
                                 const float* u_rand, float resample_padding, int32_t blurpool, int32_t num_samples,
                                 float* new_t_vals) {
  DURF_REQUIRE(B >= 0 && N >= 1 && N <= 128, DURF_E_INVALID, "durf_resample_fwd: need 1 <= N <= 128 (got %d)", N);
  DURF_REQUIRE(num_samples >= 2, DURF_E_INVALID, "durf_resample_fwd: num_samples < 2");
  if (B == 0) return DURF_OK;
  DURF_REQUIRE(t_vals && weights && new_t_vals, DURF_E_INVALID, "durf_resample_fwd: null buffer");
  ResampleParams p;
  p.B = B; p.N = N; p.S_out = num_samples; p.blur = blurpool; p.t_vals = t_vals; p.weights = weights; p.u_rand = u_rand; p.padding = blurpool ? resample_padding : 0.f;
  const double s = 1.0 / (double)num_samples;
  const double eps32 = 1.1920928955078125e-07;
  p.s_step = (float)s;
  p.s_jit = (float)(s - eps32);
  p.u_max = (float)(1.0 - eps32);
  p.out = new_t_vals;
  if (N == 128 && num_samples == 129 && aligned16(weights)) {       // the specialisation reads the weights with 16-byte loads
    resample128_kernel<<<ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(p);
  } else {
    const size_t smem = 4 * (3 * (N + 1) + 1) * sizeof(float);
    resample_kernel<<<ceil_div(B, 4), 128, smem, (cudaStream_t)stream>>>(p);
  }
  DURF_CHECK_LAUNCH("durf_resample_fwd");
  return DURF_OK;
}
