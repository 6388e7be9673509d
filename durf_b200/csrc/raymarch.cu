// K1 -- fused ray-march: fenceposts -> conical-frustum Gaussian -> (mask) -> mip360 contraction -> IPE features.
// Replaces mip.sample_along_rays / cast_rays / conical_frustum_to_gaussian / lift_gaussian (mip.py:330-370,
// 155-179, 99-130, 76-96), mip360.new_space (mip360.py:63-79), mip.integrated_pos_enc (mip.py:226-282) and
// mip.weighted_ipe (mip.py:182-223).
//
// One warp per ray.  The ray's fenceposts and the per-sample (mean, covariance diagonal) live in shared memory;
// only the covariance DIAGONAL is formed because the reference's encoding basis is a stack of scaled identities
// (mip.py:273-278), so no other entry reaches the output.  Features leave either as fp32 rows (coalesced:
// consecutive lanes write consecutive floats of the flattened [N,F] block) or as bf16 128x64 SWIZZLE_128B tile
// images that the tcgen05 MLP kernel pulls with one bulk copy per tile.
// HBM-bound by design: 48 B + 516 B in, N*F*4 B out per ray-level; compiled with -fmad=false so that the
// arithmetic is the reference's operation sequence.
#include "common.cuh"
#include "raymarch_device.cuh"

namespace durf {

struct RayMarchParams {
  DurfRaymarchArgs a;
  int M;          // rows when no device count is given
  int F;          // features per sample
  int D;          // number of degrees
};

// Fills s_t[0..N] with the ray's fenceposts (mip.py:351-368) or loads them.
__device__ __forceinline__ void load_or_sample_t(const DurfRaymarchArgs& a, int ray, int lane, float* s_t) {
  const int N = a.N;
  if (a.flags & DURF_RM_SAMPLE) {
    const float nr = a.near[ray], fr = a.far[ray];
    for (int i = lane; i <= N; i += 32) {
      const float s = (float)i / (float)N;                 // jnp.linspace(0,1,N+1)
      s_t[i] = nr * (1.f - s) + fr * s;
    }
    __syncwarp();
    if (a.flags & DURF_RM_RANDOMIZED) {
      float nt[5];                                         // N <= 128 -> at most 5 per lane
      int c = 0;
      for (int i = lane; i <= N; i += 32, ++c) {
        const float t = s_t[i];
        const float lower = (i > 0) ? 0.5f * (t + s_t[i - 1]) : t;
        const float upper = (i < N) ? 0.5f * (s_t[i + 1] + t) : t;
        nt[c] = lower + (upper - lower) * a.t_rand[(size_t)ray * (N + 1) + i];
      }
      __syncwarp();
      c = 0;
      for (int i = lane; i <= N; i += 32, ++c) s_t[i] = nt[c];
      __syncwarp();
    }
    for (int i = lane; i <= N; i += 32) a.t_vals[(size_t)ray * (N + 1) + i] = s_t[i];
  } else {
    for (int i = lane; i <= N; i += 32) s_t[i] = a.t_vals[(size_t)ray * (N + 1) + i];
    __syncwarp();
  }
}

// Feature f (0 <= f < F) of a sample.  Layout: [sin(2^l x_d)]_{l,d}, then the same shifted by pi/2 (mip.py:280-282);
// weighted variant: [mean, w[i/6] * enc_i] (mip.py:215-222).
__device__ __forceinline__ float feature_value(const float* __restrict__ g6, int f, int D, int min_deg, bool weighted,
                                               const float* __restrict__ s_w) {
  if (weighted) {
    if (f < 3) return g6[f];
    f -= 3;
  }
  const int half = 3 * D;
  const bool shifted = f >= half;
  const int ff = shifted ? f - half : f;
  const int l = ff / 3, dd = ff - 3 * l;
  const float sc = pow2i(min_deg + l);
  float y = g6[dd] * sc;
  const float yv = g6[3 + dd] * (sc * sc);
  if (shifted) y = y + kHalfPi;
  float e = expf(-0.5f * yv) * safe_sinf(y);
  if (weighted) e = s_w[f / 6] * e;
  return e;
}

__global__ void __launch_bounds__(128, 5)
raymarch_fwd_kernel(const RayMarchParams p) {
  extern __shared__ float smem[];
  const DurfRaymarchArgs& a = p.a;
  const int N = a.N, F = p.F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_w = smem;                                  // [16] BARF weights
  float* s_t = smem + 16 + warp * (7 * N + 1);        // [N+1]
  float* s_g = s_t + (N + 1);                         // [N][6]
  const bool weighted = (a.flags & DURF_RM_WEIGHTED) != 0;
  if (threadIdx.x < 16) {
    // mip.py:217-218: w_k = (1 - cos(clip(alpha - k, 0, 1) * pi)) / 2
    const float c = fminf(fmaxf((a.alpha_dev ? *a.alpha_dev : a.alpha) - (float)threadIdx.x, 0.f), 1.f);
    s_w[threadIdx.x] = (1.f - cosf(c * 3.14159265358979324f)) / 2.f;
  }
  __syncthreads();
  const int M = a.count ? min(*a.count, p.M) : p.M;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int m = blockIdx.x * (blockDim.x >> 5) + warp; m < M; m += warps_total) {
    const int ray = a.ray_index ? a.ray_index[m] : m;
    load_or_sample_t(a, ray, lane, s_t);
    const float o[3] = {a.origins[3 * ray], a.origins[3 * ray + 1], a.origins[3 * ray + 2]};
    const float d[3] = {a.dirs[3 * ray], a.dirs[3 * ray + 1], a.dirs[3 * ray + 2]};
    const float radius = a.radii[ray];
    const bool has_mult = a.ray_mult != nullptr;
    const float mult = !has_mult ? 1.f : ((a.flags & DURF_RM_MULT_IS_NHIT) ? 1.f - a.ray_mult[ray] : a.ray_mult[ray]);
    const bool fast_tiles = (a.flags & DURF_RM_OUT_BF16_TILE) && p.D == 10;
    for (int n = lane; n < N; n += 32) {
      const Gauss g = sample_gaussian(a.flags, o, d, radius, mult, has_mult, s_t[n], s_t[n + 1]);
      if (fast_tiles) {
        uint8_t* tile_base = reinterpret_cast<uint8_t*>(a.features) + ((size_t)m * N / 128) * (128 * 128);
        const int row = (int)(((size_t)m * N) % 128) + n;
        if (weighted) encode_row_bf16<true>(g, a.min_deg, s_w, tile_base, row);
        else encode_row_bf16<false>(g, a.min_deg, s_w, tile_base, row);
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) { s_g[6 * n + i] = g.mean[i]; s_g[6 * n + 3 + i] = g.var[i]; }
      if (a.means) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          a.means[((size_t)m * N + n) * 3 + i] = g.mean[i];
          a.cov_diag[((size_t)m * N + n) * 3 + i] = g.var[i];
        }
      }
    }
    __syncwarp();
    if (fast_tiles) {
      // rows already written by encode_row_bf16
    } else if (a.flags & DURF_RM_OUT_BF16_TILE) {
      // 16-byte chunks: item i -> (sample i/8, chunk i%8); a warp store covers 4 samples = 512 contiguous bytes.
      uint8_t* tile_base = reinterpret_cast<uint8_t*>(a.features) + ((size_t)m * N / 128) * (128 * 128);
      const int row0 = (int)(((size_t)m * N) % 128);
      for (int i = lane; i < N * 8; i += 32) {
        const int n = i >> 3, c = i & 7;
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int f0 = c * 8 + 2 * j;
          const float v0 = f0 < F ? feature_value(s_g + 6 * n, f0, p.D, a.min_deg, weighted, s_w) : 0.f;
          const float v1 = f0 + 1 < F ? feature_value(s_g + 6 * n, f0 + 1, p.D, a.min_deg, weighted, s_w) : 0.f;
          w[j] = pack_bf16x2(v0, v1);
        }
        *reinterpret_cast<uint4*>(tile_base + sw128_offset(row0 + n, c)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    } else {
      float* out = reinterpret_cast<float*>(a.features) + (size_t)m * N * F;
      const int total = N * F;
      for (int i = lane; i < total; i += 32) {
        const int n = i / F, f = i - n * F;
        out[i] = feature_value(s_g + 6 * n, f, p.D, a.min_deg, weighted, s_w);
      }
    }
    __syncwarp();
  }
}

// ---- backward of the weighted (object) encoding into origins_s / dirs_s -----------------------------------
// features = [mean, w * exp(-var_y/2) * sin(y [+pi/2])], mean = o + d t_mean, var = t_var d_i^2 + r_var (1 - d_i^2/|d|^2).
// One warp per ray: lanes own samples, reduce over samples with shuffles.
__global__ void __launch_bounds__(128)
raymarch_bwd_kernel(const RayMarchParams p, const float* __restrict__ d_features,
                    float* __restrict__ d_origins, float* __restrict__ d_dirs) {
  extern __shared__ float smem[];
  const DurfRaymarchArgs& a = p.a;
  const int N = a.N, F = p.F, D = p.D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_w = smem;
  float* s_t = smem + 16 + warp * (7 * N + 1);
  if (threadIdx.x < 16) {
    const float c = fminf(fmaxf((a.alpha_dev ? *a.alpha_dev : a.alpha) - (float)threadIdx.x, 0.f), 1.f);
    s_w[threadIdx.x] = (1.f - cosf(c * 3.14159265358979324f)) / 2.f;
  }
  __syncthreads();
  const bool weighted = (a.flags & DURF_RM_WEIGHTED) != 0;
  const int M = a.count ? min(*a.count, p.M) : p.M;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  for (int m = blockIdx.x * (blockDim.x >> 5) + warp; m < M; m += warps_total) {
    const int ray = a.ray_index ? a.ray_index[m] : m;
    for (int i = lane; i <= N; i += 32) s_t[i] = a.t_vals[(size_t)ray * (N + 1) + i];
    __syncwarp();
    const float o[3] = {a.origins[3 * ray], a.origins[3 * ray + 1], a.origins[3 * ray + 2]};
    const float d[3] = {a.dirs[3 * ray], a.dirs[3 * ray + 1], a.dirs[3 * ray + 2]};
    const float radius = a.radii[ray];
    const float mult = !a.ray_mult ? 1.f : ((a.flags & DURF_RM_MULT_IS_NHIT) ? 1.f - a.ray_mult[ray] : a.ray_mult[ray]);
    float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
    for (int n = lane; n < N; n += 32) {
      const float t0 = s_t[n], t1 = s_t[n + 1];
      const float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
      const float mu2 = mu * mu, hw2 = hw * hw, den = 3.f * mu2 + hw2, hw4 = hw2 * hw2;
      const float t_mean = mu + (2.f * mu * hw2) / den;
      const float t_var = hw2 / 3.f - (4.f / 15.f) * ((hw4 * (12.f * mu2 - hw2)) / (den * den));
      const float r_var = (radius * radius) * (mu2 / 4.f + (5.f / 12.f) * hw2 - (4.f / 15.f) * hw4 / den);
      const float dsq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
      const bool clampd = dsq < 1e-10f;
      const float dmag = clampd ? 1e-10f : dsq;
      const float* gf = d_features + ((size_t)m * N + n) * F;
      float gmean[3], gvar[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float mean = mult * (d[i] * t_mean + o[i]);
        const float var = mult * (t_var * (d[i] * d[i]) + r_var * (1.f - d[i] * (d[i] / dmag)));
        float gm = weighted ? gf[i] : 0.f, gv = 0.f;
        for (int l = 0; l < D; ++l) {
          const float sc = pow2i(a.min_deg + l);
          const float y = mean * sc, yv = var * (sc * sc);
          const float ex = expf(-0.5f * yv);
          const int f_s = 3 * l + i, f_c = 3 * D + 3 * l + i;
          const float w_s = weighted ? s_w[f_s / 6] : 1.f, w_c = weighted ? s_w[f_c / 6] : 1.f;
          const float g_s = gf[(weighted ? 3 : 0) + f_s] * w_s, g_c = gf[(weighted ? 3 : 0) + f_c] * w_c;
          const float ys = y + kHalfPi;
          const float sn = safe_sinf(y), cn = safe_cosf(y), sn2 = safe_sinf(ys), cn2 = safe_cosf(ys);
          gm += (g_s * ex * cn + g_c * ex * cn2) * sc;
          gv += (g_s * sn + g_c * sn2) * ex * (-0.5f) * (sc * sc);
        }
        gmean[i] = gm * mult;
        gvar[i] = gv * mult;
      }
      float qsum = 0.f;   // sum_i gvar_i * r_var * d_i^2 / dmag^2 : gradient through 1/|d|^2
#pragma unroll
      for (int i = 0; i < 3; ++i) qsum += gvar[i] * r_var * (d[i] * d[i]);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        go[i] += gmean[i];
        gd[i] += gmean[i] * t_mean + gvar[i] * (2.f * t_var * d[i] - 2.f * r_var * d[i] / dmag);
        if (!clampd) gd[i] += qsum * 2.f * d[i] / (dmag * dmag);
      }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { go[i] = warp_sum(go[i]); gd[i] = warp_sum(gd[i]); }
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { d_origins[3 * ray + i] = go[i]; d_dirs[3 * ray + i] = gd[i]; }
    }
    __syncwarp();
  }
}

__global__ void viewdir_enc_kernel(int B, int deg, const float* __restrict__ v, float* __restrict__ enc) {
  // mip.py:36-45 with append_identity: [x, sin(2^l x_d), sin(2^l x_d + pi/2)], plain sin (no safe wrapper).
  const int F = 3 + 6 * deg;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * F) return;
  const int b = (int)(i / F), f = (int)(i - (int64_t)b * F);
  float out;
  if (f < 3) out = v[3 * b + f];
  else {
    int ff = f - 3;
    const bool shifted = ff >= 3 * deg;
    if (shifted) ff -= 3 * deg;
    const int l = ff / 3, d = ff - 3 * l;
    float y = v[3 * b + d] * pow2i(l);
    if (shifted) y = y + kHalfPi;
    out = sinf(y);
  }
  enc[i] = out;
}

static int check_args(const DurfRaymarchArgs* a, RayMarchParams& p, const char* who) {
  DURF_REQUIRE(a != nullptr, DURF_E_INVALID, "%s: null args", who);
  DURF_REQUIRE(a->B >= 0 && a->N >= 1 && a->N <= 128, DURF_E_INVALID, "%s: need 1 <= N <= 128 (got %d)", who, a->N);
  const int D = a->max_deg - a->min_deg;
  DURF_REQUIRE(D >= 1 && D <= 16 && a->min_deg >= -60 && a->max_deg <= 60, DURF_E_INVALID, "%s: bad degree range [%d,%d)", who,
               a->min_deg, a->max_deg);
  p.a = *a;
  p.M = a->B;
  p.D = D;
  p.F = 6 * D + ((a->flags & DURF_RM_WEIGHTED) ? 3 : 0);
  if (a->B == 0) return DURF_OK;     // empty batch: nothing is dereferenced, null buffers are fine
  DURF_REQUIRE(a->origins && a->dirs && a->radii && a->t_vals && a->features, DURF_E_INVALID, "%s: null buffer", who);
  if (a->flags & DURF_RM_SAMPLE) DURF_REQUIRE(a->near && a->far, DURF_E_INVALID, "%s: DURF_RM_SAMPLE needs near/far", who);
  if (a->flags & DURF_RM_RANDOMIZED) DURF_REQUIRE(a->t_rand, DURF_E_INVALID, "%s: DURF_RM_RANDOMIZED needs t_rand", who);
  DURF_REQUIRE((a->means == nullptr) == (a->cov_diag == nullptr), DURF_E_INVALID, "%s: means and cov_diag go together", who);
  if (a->flags & DURF_RM_OUT_BF16_TILE)
    DURF_REQUIRE(p.F <= 64 && 128 % a->N == 0, DURF_E_UNSUPPORTED, "%s: bf16 tile output needs F <= 64 and N | 128", who);
  return DURF_OK;
}

}  // namespace durf

using namespace durf;

extern "C" int durf_raymarch_fwd(durf_stream_t stream, const DurfRaymarchArgs* args) {
  RayMarchParams p;
  int rc = check_args(args, p, "durf_raymarch_fwd");
  if (rc != DURF_OK) return rc;
  if (p.M == 0) return DURF_OK;
  const size_t smem = (16 + 4 * (7 * args->N + 1)) * sizeof(float);
  const int grid = min(ceil_div(p.M, 4), 148 * 16);
  raymarch_fwd_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(p);
  DURF_CHECK_LAUNCH("durf_raymarch_fwd");
  return DURF_OK;
}

extern "C" int durf_raymarch_bwd(durf_stream_t stream, const DurfRaymarchArgs* args, const float* d_features,
                                 float* d_origins_s, float* d_dirs_s) {
  RayMarchParams p;
  int rc = check_args(args, p, "durf_raymarch_bwd");
  if (rc != DURF_OK) return rc;
  DURF_REQUIRE(d_features && d_origins_s && d_dirs_s, DURF_E_INVALID, "durf_raymarch_bwd: null gradient buffer");
  DURF_REQUIRE(!(args->flags & (DURF_RM_CONTRACT | DURF_RM_CYLINDER | DURF_RM_NO_INTEGRATE | DURF_RM_OUT_BF16_TILE)),
               DURF_E_UNSUPPORTED, "durf_raymarch_bwd: only the object (weighted / plain IPE, cone, fp32) variant has a backward");
  if (p.M == 0) return DURF_OK;
  const size_t smem = (16 + 4 * (7 * args->N + 1)) * sizeof(float);
  const int grid = min(ceil_div(p.M, 4), 148 * 16);
  raymarch_bwd_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(p, d_features, d_origins_s, d_dirs_s);
  DURF_CHECK_LAUNCH("durf_raymarch_bwd");
  return DURF_OK;
}

extern "C" int durf_viewdir_enc_fwd(durf_stream_t stream, int32_t B, int32_t deg, const float* viewdirs, float* enc) {
  DURF_REQUIRE(B >= 0 && deg >= 0 && deg <= 16, DURF_E_INVALID, "durf_viewdir_enc_fwd: bad argument");
  if (B == 0) return DURF_OK;
  DURF_REQUIRE(viewdirs && enc, DURF_E_INVALID, "durf_viewdir_enc_fwd: null buffer");
  const int64_t total = (int64_t)B * (3 + 6 * deg);
  viewdir_enc_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(B, deg, viewdirs, enc);
  DURF_CHECK_LAUNCH("durf_viewdir_enc_fwd");
  return DURF_OK;
}
