// K3 -- activations + alpha compositing, forward and backward.  One warp per ray, 4 consecutive samples per lane
// (float4 loads of raw density / weights, 3 x float4 of raw rgb), exclusive transmittance prefix by warp scan.
// Replaces obbpose_model.py:243-245 (sigmoid / softplus(x + density_bias)) and mip.volumetric_rendering
// (mip.py:285-327).  HBM-bound: 16 B/sample in, 4..12 B/sample out.
#include "common.cuh"

namespace durf {

constexpr int kSPL = 4;  // samples per lane

__device__ __forceinline__ float softplusf(float x) {   // jax.nn.softplus = logaddexp(x, 0)
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

struct RayLocal {
  float t[kSPL + 1];
  float dens[kSPL];
  float rgb[kSPL][3];
  float raw_d[kSPL];
};

__device__ __forceinline__ void load_ray(const DurfCompositeArgs& a, int ray, int lane, RayLocal& r) {
  const int N = a.N;
  const int n0 = lane * kSPL;
  const float* tv = a.t_vals + (size_t)ray * (N + 1);
  const float* rd = a.raw_density + (size_t)ray * N;
  const float* rc = a.raw_rgb + (size_t)ray * N * 3;
  if (N == 128) {
    const float4 d4 = *reinterpret_cast<const float4*>(rd + n0);
    r.raw_d[0] = d4.x; r.raw_d[1] = d4.y; r.raw_d[2] = d4.z; r.raw_d[3] = d4.w;
    const float4* c4 = reinterpret_cast<const float4*>(rc + 3 * n0);
    const float4 c0 = c4[0], c1 = c4[1], c2 = c4[2];
    const float cc[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
#pragma unroll
    for (int q = 0; q < kSPL; ++q)
#pragma unroll
      for (int c = 0; c < 3; ++c) r.rgb[q][c] = cc[3 * q + c];
#pragma unroll
    for (int q = 0; q <= kSPL; ++q) r.t[q] = tv[n0 + q];      // (N+1)-float rows are only 4-byte aligned
  } else {
#pragma unroll
    for (int q = 0; q < kSPL; ++q) {
      const bool ok = n0 + q < N;
      r.raw_d[q] = ok ? rd[n0 + q] : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) r.rgb[q][c] = ok ? rc[3 * (n0 + q) + c] : 0.f;
    }
#pragma unroll
    for (int q = 0; q <= kSPL; ++q) r.t[q] = (n0 + q <= N) ? tv[n0 + q] : 0.f;
  }
}

__global__ void __launch_bounds__(128)
composite_fwd_kernel(const DurfCompositeArgs a) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= a.B) return;
  const int N = a.N, n0 = lane * kSPL;
  RayLocal r;
  load_ray(a, ray, lane, r);
  const float dx = a.dirs[3 * ray], dy = a.dirs[3 * ray + 1], dz = a.dirs[3 * ray + 2];
  const float dn = sqrtf(dx * dx + dy * dy + dz * dz);

  float dd[kSPL], tmid[kSPL], tdist[kSPL];
  float run = 0.f;
#pragma unroll
  for (int q = 0; q < kSPL; ++q) {
    const bool ok = n0 + q < N;
    tmid[q] = 0.5f * (r.t[q] + r.t[q + 1]);
    tdist[q] = r.t[q + 1] - r.t[q];
    const float dens = a.activated ? r.raw_d[q] : softplusf(r.raw_d[q] + a.density_bias);
    dd[q] = ok ? dens * (tdist[q] * dn) : 0.f;
    run += dd[q];
  }
  float before = warp_scan_excl(run, lane);                     // sum of density*delta over earlier lanes
  float w[kSPL];
  float acc = 0.f, dep = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
#pragma unroll
  for (int q = 0; q < kSPL; ++q) {
    const bool ok = n0 + q < N;
    const float alpha = 1.f - expf(-dd[q]);
    const float trans = expf(-before);
    w[q] = ok ? nan_to_num(alpha * trans) : 0.f;
    before += dd[q];
    acc += w[q];
    dep += w[q] * tmid[q];
    cr += w[q] * (a.activated ? r.rgb[q][0] : sigmoidf(r.rgb[q][0]));
    cg += w[q] * (a.activated ? r.rgb[q][1] : sigmoidf(r.rgb[q][1]));
    cb += w[q] * (a.activated ? r.rgb[q][2] : sigmoidf(r.rgb[q][2]));
  }
  acc = warp_sum(acc); dep = warp_sum(dep); cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);

  if (N == 128) {
    *reinterpret_cast<float4*>(a.weights + (size_t)ray * N + n0) = make_float4(w[0], w[1], w[2], w[3]);
    if (a.t_mids) *reinterpret_cast<float4*>(a.t_mids + (size_t)ray * N + n0) = make_float4(tmid[0], tmid[1], tmid[2], tmid[3]);
    if (a.t_dists) *reinterpret_cast<float4*>(a.t_dists + (size_t)ray * N + n0) = make_float4(tdist[0], tdist[1], tdist[2], tdist[3]);
  } else {
#pragma unroll
    for (int q = 0; q < kSPL; ++q) if (n0 + q < N) {
      a.weights[(size_t)ray * N + n0 + q] = w[q];
      if (a.t_mids) a.t_mids[(size_t)ray * N + n0 + q] = tmid[q];
      if (a.t_dists) a.t_dists[(size_t)ray * N + n0 + q] = tdist[q];
    }
  }
  if (lane == 0) {
    float bg = 0.f, c[3] = {cr, cg, cb};
    // mip.py:321-326: white adds (1-acc); rand_bkgd adds randint(0,1)==0 times (1-acc); otherwise grey 0.5.
    if (a.white_bkgd) { bg = 1.f - acc; c[0] += bg; c[1] += bg; c[2] += bg; }
    if (a.rand_bkgd) { const float z = 0.f * (1.f - acc); c[0] += z; c[1] += z; c[2] += z; }
    else if (!a.white_bkgd) { bg = 0.5f * (1.f - acc); c[0] += bg; c[1] += bg; c[2] += bg; }
    a.comp_rgb[3 * ray] = c[0]; a.comp_rgb[3 * ray + 1] = c[1]; a.comp_rgb[3 * ray + 2] = c[2];
    a.depth[ray] = dep;
    a.acc[ray] = acc;
  }
}

// Backward.  With g_i = dL/dw_i (all paths), s_j = density_j * delta_j:
//   dL/ds_j = g_j T_j (1 - alpha_j) - sum_{i>j} g_i w_i        (reverse exclusive scan)
__global__ void __launch_bounds__(128)
composite_bwd_kernel(const DurfCompositeArgs a, const float* __restrict__ d_comp_rgb, const float* __restrict__ d_depth,
                     const float* __restrict__ d_acc, const float* __restrict__ d_weights,
                     float* __restrict__ d_raw_rgb, float* __restrict__ d_raw_density, float* __restrict__ d_dirs) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= a.B) return;
  const int N = a.N, n0 = lane * kSPL;
  RayLocal r;
  load_ray(a, ray, lane, r);
  const float dx = a.dirs[3 * ray], dy = a.dirs[3 * ray + 1], dz = a.dirs[3 * ray + 2];
  const float dn = sqrtf(dx * dx + dy * dy + dz * dz);
  const float gc[3] = {d_comp_rgb[3 * ray], d_comp_rgb[3 * ray + 1], d_comp_rgb[3 * ray + 2]};
  const float gdep = d_depth[ray];
  float gacc = d_acc ? d_acc[ray] : 0.f;
  if (a.white_bkgd) gacc -= (gc[0] + gc[1] + gc[2]);
  if (!a.rand_bkgd && !a.white_bkgd) gacc -= 0.5f * (gc[0] + gc[1] + gc[2]);

  float dd[kSPL], dens[kSPL], tdist[kSPL], tmid[kSPL];
  float run = 0.f;
#pragma unroll
  for (int q = 0; q < kSPL; ++q) {
    const bool ok = n0 + q < N;
    tmid[q] = 0.5f * (r.t[q] + r.t[q + 1]);
    tdist[q] = r.t[q + 1] - r.t[q];
    dens[q] = a.activated ? r.raw_d[q] : softplusf(r.raw_d[q] + a.density_bias);
    dd[q] = ok ? dens[q] * (tdist[q] * dn) : 0.f;
    run += dd[q];
  }
  float before = warp_scan_excl(run, lane);
  float w[kSPL], trans[kSPL], alpha[kSPL], gw[kSPL], sig[kSPL][3];
  float gwsum = 0.f;
  float gin[kSPL];                                                 // dL/dweights of this lane's samples (one 16-byte load at N = 128)
  if (N == 128) {
    const float4 g4 = *reinterpret_cast<const float4*>(d_weights + (size_t)ray * N + n0);
    gin[0] = g4.x; gin[1] = g4.y; gin[2] = g4.z; gin[3] = g4.w;
  } else {
#pragma unroll
    for (int q = 0; q < kSPL; ++q) gin[q] = (n0 + q < N) ? d_weights[(size_t)ray * N + n0 + q] : 0.f;
  }
#pragma unroll
  for (int q = 0; q < kSPL; ++q) {
    const bool ok = n0 + q < N;
    alpha[q] = 1.f - expf(-dd[q]);
    trans[q] = expf(-before);
    before += dd[q];
    const float raw_w = alpha[q] * trans[q];
    const bool finite = (raw_w == raw_w) && !isinf(raw_w);
    w[q] = ok ? nan_to_num(raw_w) : 0.f;
    float g = gin[q];
#pragma unroll
    for (int c = 0; c < 3; ++c) { sig[q][c] = a.activated ? r.rgb[q][c] : sigmoidf(r.rgb[q][c]); g += gc[c] * sig[q][c]; }
    g += gdep * tmid[q] + gacc;
    gw[q] = (ok && finite) ? g : 0.f;                            // nan_to_num blocks the gradient of non-finite entries
    gwsum += gw[q] * w[q];
  }
  // reverse exclusive scan of gw*w over samples
  float after = warp_rscan_excl(gwsum, lane);                    // sum over later lanes
  float gnorm = 0.f;
  float od[kSPL], oc[kSPL * 3];                                    // this lane's 4 + 12 output floats (contiguous in both arrays)
#pragma unroll
  for (int q = kSPL - 1; q >= 0; --q) {
    const bool ok = n0 + q < N;
    const float gs = gw[q] * trans[q] * (1.f - alpha[q]) - after;
    after += gw[q] * w[q];
    const float gdens = gs * (tdist[q] * dn);
    od[q] = a.activated ? gdens : gdens * sigmoidf(r.raw_d[q] + a.density_bias);
    if (ok) gnorm += gs * dens[q] * tdist[q];
#pragma unroll
    for (int c = 0; c < 3; ++c) oc[3 * q + c] = a.activated ? gc[c] * w[q] : gc[c] * w[q] * sig[q][c] * (1.f - sig[q][c]);
  }
  if (N == 128) {      // whole 16-byte stores: 1 + 3 per lane instead of 16 strided 4-byte stores
    *reinterpret_cast<float4*>(d_raw_density + (size_t)ray * N + n0) = make_float4(od[0], od[1], od[2], od[3]);
    float4* o4 = reinterpret_cast<float4*>(d_raw_rgb + ((size_t)ray * N + n0) * 3);
    o4[0] = make_float4(oc[0], oc[1], oc[2], oc[3]);
    o4[1] = make_float4(oc[4], oc[5], oc[6], oc[7]);
    o4[2] = make_float4(oc[8], oc[9], oc[10], oc[11]);
  } else {
#pragma unroll
    for (int q = 0; q < kSPL; ++q)
      if (n0 + q < N) {
        d_raw_density[(size_t)ray * N + n0 + q] = od[q];
#pragma unroll
        for (int c = 0; c < 3; ++c) d_raw_rgb[((size_t)ray * N + n0 + q) * 3 + c] = oc[3 * q + c];
      }
  }
  if (d_dirs) {
    gnorm = warp_sum(gnorm);
    if (lane == 0) {
      d_dirs[3 * ray] = gnorm * dx / dn;
      d_dirs[3 * ray + 1] = gnorm * dy / dn;
      d_dirs[3 * ray + 2] = gnorm * dz / dn;
    }
  }
}

static int check(const DurfCompositeArgs* a, const char* who) {
  DURF_REQUIRE(a != nullptr, DURF_E_INVALID, "%s: null args", who);
  DURF_REQUIRE(a->B >= 0 && a->N >= 1 && a->N <= 128, DURF_E_INVALID, "%s: need 1 <= N <= 128 (got %d)", who, a->N);
  DURF_REQUIRE(a->B == 0 || (a->raw_rgb && a->raw_density && a->t_vals && a->dirs), DURF_E_INVALID, "%s: null input", who);
  return DURF_OK;
}

}  // namespace durf

using namespace durf;

extern "C" int durf_composite_fwd(durf_stream_t stream, const DurfCompositeArgs* args) {
  int rc = check(args, "durf_composite_fwd");
  if (rc != DURF_OK) return rc;
  if (args->B == 0) return DURF_OK;
  DURF_REQUIRE(args->comp_rgb && args->depth && args->acc && args->weights, DURF_E_INVALID, "durf_composite_fwd: null output");
  // N = 128 rows are moved with 16-byte loads / stores
  DURF_REQUIRE(args->N != 128 || aligned16(args->raw_rgb, args->raw_density, args->weights, args->t_mids, args->t_dists), DURF_E_INVALID,
               "durf_composite_fwd: raw_rgb / raw_density / weights / t_mids / t_dists must be 16-byte aligned when N = 128");
  composite_fwd_kernel<<<ceil_div(args->B, 4), 128, 0, (cudaStream_t)stream>>>(*args);
  DURF_CHECK_LAUNCH("durf_composite_fwd");
  return DURF_OK;
}

extern "C" int durf_composite_bwd(durf_stream_t stream, const DurfCompositeArgs* args, const float* d_comp_rgb,
                                  const float* d_depth, const float* d_acc, const float* d_weights,
                                  float* d_raw_rgb, float* d_raw_density, float* d_dirs) {
  int rc = check(args, "durf_composite_bwd");
  if (rc != DURF_OK) return rc;
  if (args->B == 0) return DURF_OK;
  DURF_REQUIRE(d_comp_rgb && d_depth && d_weights && d_raw_rgb && d_raw_density, DURF_E_INVALID,
               "durf_composite_bwd: null gradient buffer");
  DURF_REQUIRE(args->N != 128 || aligned16(args->raw_rgb, args->raw_density, d_weights, d_raw_rgb, d_raw_density), DURF_E_INVALID,
               "durf_composite_bwd: raw_rgb / raw_density / d_weights / d_raw_rgb / d_raw_density must be 16-byte aligned when N = 128");
  composite_bwd_kernel<<<ceil_div(args->B, 4), 128, 0, (cudaStream_t)stream>>>(*args, d_comp_rgb, d_depth, d_acc, d_weights,
                                                                                d_raw_rgb, d_raw_density, d_dirs);
  DURF_CHECK_LAUNCH("durf_composite_bwd");
  return DURF_OK;
}
