// Device functions of the ray-march shared by raymarch.cu (K1 as its own kernel) and mlp_tc.cu (K1 fused into the tcgen05
// MLP kernel): per-sample Gaussian of mip.py / mip360.py and the bf16 feature row of the tensor-core path.
#pragma once

#include "common.cuh"

namespace durf {

__device__ __forceinline__ float pow2i(int l) { return __int_as_float((127 + l) << 23); }

struct Gauss {
  float mean[3];
  float var[3];
};

// Per-sample Gaussian: mip.py:117-124 (cone) / 149-151 (cylinder), lift (76-96, diagonal only),
// ray multiplier (obbpose_model.py:179-180 / 207-208), contraction (mip360.py:47-79).
__device__ __forceinline__ Gauss sample_gaussian(uint32_t flags, const float o[3], const float d[3],
                                                 float radius, float mult, bool has_mult, float t0, float t1) {
  float t_mean, t_var, r_var;
  if (flags & DURF_RM_CYLINDER) {
    t_mean = (t0 + t1) / 2.f;
    r_var = radius * radius / 4.f;
    t_var = (t1 - t0) * (t1 - t0) / 12.f;
  } else {
    const float mu = (t0 + t1) / 2.f;
    const float hw = (t1 - t0) / 2.f;
    const float mu2 = mu * mu, hw2 = hw * hw;
    const float den = 3.f * mu2 + hw2;
    const float hw4 = hw2 * hw2;
    t_mean = mu + (2.f * mu * hw2) / den;
    t_var = hw2 / 3.f - (4.f / 15.f) * ((hw4 * (12.f * mu2 - hw2)) / (den * den));
    r_var = (radius * radius) * (mu2 / 4.f + (5.f / 12.f) * hw2 - (4.f / 15.f) * hw4 / den);
  }
  const float dmag = fmaxf(1e-10f, d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  Gauss g;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    g.mean[i] = d[i] * t_mean + o[i];
    const float outer = d[i] * d[i];
    const float null_outer = 1.f - d[i] * (d[i] / dmag);
    g.var[i] = t_var * outer + r_var * null_outer;
  }
  if (flags & DURF_RM_NO_INTEGRATE) g.var[0] = g.var[1] = g.var[2] = 0.f;
  if (has_mult) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { g.mean[i] = mult * g.mean[i]; g.var[i] = mult * g.var[i]; }
  }
  if (flags & DURF_RM_CONTRACT) {
    const float x0 = g.mean[0], x1 = g.mean[1], x2 = g.mean[2];
    float sq = x0 * x0 + x1 * x1 + x2 * x2;
    const bool floor_hit = sq < 1e-12f;
    sq = floor_hit ? 1e-12f : sq;
    const float n = sqrtf(sq);
    if (n > 0.1f) {
      const float inv = 1.f / n;
      const float A = 2.f - inv;
      const float dn = floor_hit ? 0.f : (x0 + x1 + x2) / n;     // JVP of the norm along the all-ones tangent
      const float dA = dn / (n * n);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float Bv = g.mean[i] / n;
        const float dB = inv - g.mean[i] * dn / (n * n);
        const float v = dA * Bv + A * dB;
        g.mean[i] = A * Bv;
        g.var[i] = (g.var[i] * v) * v;                           // cov @ diag(v)^2, diagonal entry
      }
    }
  }
  return g;
}

// Tensor-core path (features leave as bf16, half-ulp 2^-9): all 2*3*10 encodings of a sample from THREE accurate sincosf
// calls, the higher octaves by angle doubling (sin 2y = 2 s c, cos 2y = 1 - 2 s^2; the error doubles per octave and stays
// below 1e-4 at 2^9, 20x under the bf16 rounding of the stored value) and one ex2 per (octave, axis).  The fp32 output
// path of raymarch.cu keeps the reference's exact sequence (sin(y + fl32(pi/2)) with the 100*pi wrap, math.py:35-36) instead.
// Produces the sample's 128-byte row (64 bf16, columns past the feature count are zero) as 32 packed words.
template <bool weighted>
__device__ __forceinline__ void encode_feats_bf16(const Gauss& g, int min_deg, const float* __restrict__ s_w, uint32_t (&w)[32]) {
  constexpr int D = 10;
  float feat[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) feat[i] = 0.f;
  constexpr int o = weighted ? 3 : 0;
  const float sc0 = pow2i(min_deg);
  float sn[3], cs[3], yv[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (weighted) feat[d] = g.mean[d];
    sincosf(g.mean[d] * sc0, &sn[d], &cs[d]);
    yv[d] = g.var[d] * (sc0 * sc0) * (-0.5f * 1.44269504088896341f);
  }
  // octave-major order: a pair of neighbouring features is complete within two octaves, so it can be packed early
#pragma unroll
  for (int l = 0; l < D; ++l) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float e = exp2f(yv[d]);
      feat[o + 3 * l + d] = e * sn[d];
      feat[o + 3 * D + 3 * l + d] = e * cs[d];
      const float s2 = 2.f * sn[d] * cs[d], c2 = 1.f - 2.f * (sn[d] * sn[d]);
      sn[d] = s2; cs[d] = c2;
      yv[d] = yv[d] * 4.f;
    }
  }
  if (weighted) {
#pragma unroll
    for (int i = 0; i < 6 * D; ++i) feat[3 + i] = s_w[i / 6] * feat[3 + i];     // mip.py:220: weight index i // 6
  }
#pragma unroll
  for (int e = 0; e < 32; ++e) w[e] = pack_bf16x2(feat[2 * e], feat[2 * e + 1]);
}

// The row written to a SWIZZLE_128B tile image in GLOBAL memory: 32-byte stores (STG.256), every store fills a whole sector.
template <bool weighted>
__device__ __forceinline__ void encode_row_bf16(const Gauss& g, int min_deg, const float* __restrict__ s_w,
                                                uint8_t* __restrict__ tile_base, int row) {
  uint32_t w[32];
  encode_feats_bf16<weighted>(g, min_deg, s_w, w);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t w8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w8[e] = w[8 * k + e];
    st_sw128_pair(tile_base, (uint32_t)row, 2 * k, w8);
  }
}

// The same row written to a tile image in SHARED memory (the tcgen05 MLP kernel generating its own input tile, N1).
template <bool weighted>
__device__ __forceinline__ void encode_row_bf16_smem(const Gauss& g, int min_deg, const float* __restrict__ s_w,
                                                     uint32_t tile_smem, int row) {
  uint32_t w[32];
  encode_feats_bf16<weighted>(g, min_deg, s_w, w);
#pragma unroll
  for (int c = 0; c < 8; ++c)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_smem + sw128_offset((uint32_t)row, (uint32_t)c)),
                 "r"(w[4 * c]), "r"(w[4 * c + 1]), "r"(w[4 * c + 2]), "r"(w[4 * c + 3]) : "memory");
}

// Fenceposts i and i+1 of a ray computed in place (mip.py:351-368): linspace in depth, optional stratified jitter.
__device__ __forceinline__ float sample_fencepost(float nr, float fr, int i, int N, const float* __restrict__ t_rand_row) {
  auto base = [&](int k) { const float s = (float)k / (float)N; return nr * (1.f - s) + fr * s; };
  const float t = base(i);
  if (!t_rand_row) return t;
  const float lower = (i > 0) ? 0.5f * (t + base(i - 1)) : t;
  const float upper = (i < N) ? 0.5f * (base(i + 1) + t) : t;
  return lower + (upper - lower) * t_rand_row[i];
}

}  // namespace durf
