// XLA typed-FFI handlers over the C ABI of libdurf_b200.so (include/durf_b200.h).
//
// This is the reference-side binding a DURF maintainer adds to reach the CUDA hot path from JAX (north_star: "the hot path
// reached through jax.ffi custom calls over a thin C-ABI").  Each handler takes the stream XLA executes on and the device
// buffers XLA owns and forwards them, unchanged, to ONE C-ABI entry point: no allocation, no synchronisation, no stream of
// its own - legal inside jit / pmap and CUDA-graph capture (command-buffer compatible).  Forward AND backward handlers are
// here; integration/jax_ffi/durf_jax.py wires them into jax.custom_vjp rules so that jax.value_and_grad(loss_fn)
// (train_boxpose.py:251) differentiates through them.
//
// Build (any jaxlib that ships xla/ffi/api/ffi.h):
//     g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -I../../include
//         durf_ffi.cc -L../../durf_b200 -ldurf_b200 -o libdurf_jax_ffi.so        (one command line)
// The build image of this repository has no jaxlib; tests/test_ffi_shim.py compiles this file against the real header when
// it can find one and otherwise against tests/_ffi_stub (an API stand-in that type-checks the handler bodies against
// include/durf_b200.h) and says which.
#include <cstdint>

#include "durf_b200.h"
#include "xla/ffi/api/ffi.h"

#ifndef DURF_FFI_STUB_HEADER
#include <cuda_runtime_api.h>
#else
typedef struct CUstream_st* cudaStream_t;      // the stub build has no CUDA toolkit dependency
#endif

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;
using U8 = ffi::Buffer<ffi::U8>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

static ffi::Error Check(int rc) {
  if (rc == DURF_OK) return ffi::Error::Success();
  return ffi::Error(rc == DURF_E_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, durf_last_error());
}
template <class B>
static int32_t Dim(const B& b, int i) { return static_cast<int32_t>(b.dimensions()[i]); }
template <class B>
static auto* OrNull(B& b) { return b.element_count() ? b.typed_data() : nullptr; }      // zero-sized operand = "not given"

// ---- K0: world2object_rpy + ray_box_intersection + scene-graph merge (box_helpers.py:286-341, 59-106; obbpose_model.py:99-131)
static ffi::Error ObbFrontendFwdImpl(cudaStream_t stream, F32 origins, F32 directions, F32 box, F32 ext, RF32 origins_s, RF32 dirs_s,
                                     RS32 hit, RF32 zi, RF32 zo, RF32 zo_ret, RF32 nhit) {
  return Check(durf_obb_frontend_fwd(stream, Dim(origins, 0), Dim(box, 0), origins.typed_data(), directions.typed_data(),
                                     box.typed_data(), ext.typed_data(), origins_s->typed_data(), dirs_s->typed_data(),
                                     hit->typed_data(), zi->typed_data(), zo->typed_data(), zo_ret->typed_data(), nhit->typed_data(),
                                     nullptr, nullptr));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfObbFrontendFwd, ObbFrontendFwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Ret<F32>().Ret<F32>().Ret<S32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

// d box_centers[ts] from dL/d origins_s, dL/d dirs_s (obbpose_model.py:99-122 under jax.value_and_grad).  d_box is ACCUMULATED
// by the library: the rule below passes a zero-initialised operand aliased to the result (input_output_aliases={6: 0}).
static ffi::Error ObbFrontendBwdImpl(cudaStream_t stream, F32 origins, F32 directions, F32 box, S32 hit, F32 d_origins_s, F32 d_dirs_s,
                                     F32 d_box_init, int32_t pose_grad, int32_t rot_grad, RF32 d_box) {
  (void)d_box_init;                                  // same buffer as d_box
  return Check(durf_obb_frontend_bwd(stream, Dim(origins, 0), Dim(box, 0), origins.typed_data(), directions.typed_data(), box.typed_data(),
                                     hit.typed_data(), d_origins_s.typed_data(), d_dirs_s.typed_data(), pose_grad, rot_grad,
                                     d_box->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfObbFrontendBwd, ObbFrontendBwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Attr<int32_t>("pose_grad").Attr<int32_t>("rot_grad").Ret<F32>());

// rays that hit object k: ray_index [B] (unordered), count [1]
static ffi::Error CompactHitsImpl(cudaStream_t stream, S32 hit, int32_t k, RS32 ray_index, RS32 count) {
  return Check(durf_compact_hits(stream, Dim(hit, 0), Dim(hit, 1), k, hit.typed_data(), ray_index->typed_data(), count->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfCompactHits, CompactHitsImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<S32>().Attr<int32_t>("k").Ret<S32>().Ret<S32>());

static ffi::Error CompactHitsAllImpl(cudaStream_t stream, S32 hit, RS32 ray_index, RS32 count) {
  return Check(durf_compact_hits_all(stream, Dim(hit, 0), Dim(hit, 1), hit.typed_data(), ray_index->typed_data(), count->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfCompactHitsAll, CompactHitsAllImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<S32>().Ret<S32>().Ret<S32>());

// raw[ray_index[m]] += src[m] for an object network evaluated into compact rows (DurfMlpArgs.accumulate == 2): the per-ray
// buffers are operands aliased to the results (input_output_aliases={4: 0, 5: 1}).
static ffi::Error MlpMergeRawImpl(cudaStream_t stream, F32 src_rgb, F32 src_density, S32 ray_index, S32 count, F32 raw_rgb_in,
                                  F32 raw_density_in, RF32 raw_rgb, RF32 raw_density) {
  (void)raw_rgb_in; (void)raw_density_in;            // same buffers as the results
  return Check(durf_mlp_merge_raw(stream, Dim(src_density, 0), Dim(src_density, 1), ray_index.typed_data(), OrNull(count),
                                  src_rgb.typed_data(), src_density.typed_data(), raw_rgb->typed_data(), raw_density->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfMlpMergeRaw, MlpMergeRawImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<S32>()
                                  .Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>());

// ---- K1: mip.sample_along_rays | resampled t_vals -> cast_rays -> mip360.new_space -> integrated_pos_enc | weighted_ipe
// (mip.py:330-370, 155-179, 226-282, 182-223; mip360.py:63-79).  t_vals is an OPERAND aliased to the first result
// (input_output_aliases={7: 0}): read when DURF_RM_SAMPLE is clear, written when it is set.  `features` is fp32 [M,N,F] or,
// with DURF_RM_OUT_BF16_TILE, the bf16 tile images [M, 128*64] the tensor-core MLP consumes - hence an untyped result.
static void FillRaymarch(DurfRaymarchArgs& a, F32& origins, F32& dirs, F32& radii, F32& near, F32& far, F32& t_rand, F32& ray_mult,
                         S32& ray_index, S32& count, int32_t num_samples, int32_t min_deg, int32_t max_deg, int32_t flags, float alpha) {
  a.B = Dim(origins, 0);
  a.N = num_samples; a.min_deg = min_deg; a.max_deg = max_deg; a.flags = static_cast<uint32_t>(flags); a.alpha = alpha;
  a.origins = origins.typed_data(); a.dirs = dirs.typed_data(); a.radii = radii.typed_data();
  a.near = OrNull(near); a.far = OrNull(far); a.t_rand = OrNull(t_rand); a.ray_mult = OrNull(ray_mult);
  a.ray_index = OrNull(ray_index); a.count = OrNull(count);
}
static ffi::Error RaymarchFwdImpl(cudaStream_t stream, F32 origins, F32 dirs, F32 radii, F32 near, F32 far, F32 t_rand, F32 ray_mult,
                                  F32 t_vals_in, S32 ray_index, S32 count, int32_t num_samples, int32_t min_deg, int32_t max_deg,
                                  int32_t flags, float alpha, RF32 t_vals, ffi::Result<ffi::AnyBuffer> features) {
  (void)t_vals_in;                                   // same buffer as t_vals
  DurfRaymarchArgs a{};
  FillRaymarch(a, origins, dirs, radii, near, far, t_rand, ray_mult, ray_index, count, num_samples, min_deg, max_deg, flags, alpha);
  a.t_vals = t_vals->typed_data();
  a.features = features->untyped_data();
  return Check(durf_raymarch_fwd(stream, &a));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfRaymarchFwd, RaymarchFwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<S32>()
                                  .Attr<int32_t>("num_samples").Attr<int32_t>("min_deg").Attr<int32_t>("max_deg")
                                  .Attr<int32_t>("flags").Attr<float>("alpha").Ret<F32>().Ret<ffi::AnyBuffer>());

// d features [M,N,63] -> d origins_s, d dirs_s [B,3] (object-frame weighted IPE: the box-pose path); both results are
// zero-initialised operands aliased to the results (rows of rays outside ray_index stay zero)
static ffi::Error RaymarchBwdImpl(cudaStream_t stream, F32 origins, F32 dirs, F32 radii, F32 t_vals, S32 ray_index, S32 count,
                                  F32 d_features, F32 d_origins_init, F32 d_dirs_init, int32_t num_samples, int32_t min_deg,
                                  int32_t max_deg, int32_t flags, float alpha, RF32 d_origins_s, RF32 d_dirs_s) {
  (void)d_origins_init; (void)d_dirs_init;
  DurfRaymarchArgs a{};
  a.B = Dim(origins, 0);
  a.N = num_samples; a.min_deg = min_deg; a.max_deg = max_deg; a.flags = static_cast<uint32_t>(flags); a.alpha = alpha;
  a.origins = origins.typed_data(); a.dirs = dirs.typed_data(); a.radii = radii.typed_data();
  a.t_vals = const_cast<float*>(t_vals.typed_data());
  a.ray_index = OrNull(ray_index); a.count = OrNull(count);
  a.features = const_cast<float*>(d_features.typed_data());      // unused by the backward; must be non-null
  return Check(durf_raymarch_bwd(stream, &a, d_features.typed_data(), d_origins_s->typed_data(), d_dirs_s->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfRaymarchBwd, RaymarchBwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<S32>().Arg<S32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("num_samples").Attr<int32_t>("min_deg").Attr<int32_t>("max_deg")
                                  .Attr<int32_t>("flags").Attr<float>("alpha").Ret<F32>().Ret<F32>());

// mip.pos_enc(viewdirs, 0, deg, append_identity=True) (mip.py:36-45)
static ffi::Error ViewdirEncImpl(cudaStream_t stream, F32 viewdirs, int32_t deg, RF32 enc) {
  return Check(durf_viewdir_enc_fwd(stream, Dim(viewdirs, 0), deg, viewdirs.typed_data(), enc->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfViewdirEnc, ViewdirEncImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Attr<int32_t>("deg").Ret<F32>());

// ---- K2: MLP.__call__ / BoxMLP.__call__ (obbpose_model.py:294-354, 358-418) -----------------------------------------------
static DurfMlpTopology Topo(int32_t in_dim, int32_t width, int32_t depth, int32_t skip, int32_t cond_dim, int32_t cond_width) {
  return DurfMlpTopology{in_dim, width, depth, skip, cond_dim, cond_width};
}
// fp32 parameter blob -> tensor-core weight image (call again after every optimizer step)
static ffi::Error MlpPackImpl(cudaStream_t stream, F32 params, int32_t in_dim, int32_t width, int32_t depth, int32_t skip,
                              int32_t cond_dim, int32_t cond_width, RU8 packed) {
  const DurfMlpTopology t = Topo(in_dim, width, depth, skip, cond_dim, cond_width);
  return Check(durf_mlp_pack_weights(stream, &t, params.typed_data(), packed->untyped_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfMlpPack, MlpPackImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>()
                                  .Attr<int32_t>("in_dim").Attr<int32_t>("width").Attr<int32_t>("depth").Attr<int32_t>("skip")
                                  .Attr<int32_t>("cond_dim").Attr<int32_t>("cond_width").Ret<U8>());

// Forward.  `features` = bf16 tile images (precision 1) or fp32 rows (precision 0).  raw_rgb / raw_density are operands aliased
// to the results (input_output_aliases={6: 0, 7: 1}) so that an object MLP ACCUMULATES into the background's buffers
// (obbpose_model.py:203-204, 233-234).  `saved` (size durf_mlp_saved_bytes, 0 bytes = inference) and `workspace`
// (durf_mlp_workspace_bytes) are extra results XLA allocates; `saved` feeds the backward handler.
static ffi::Error MlpFwdImpl(cudaStream_t stream, ffi::AnyBuffer features, F32 cond, F32 params, U8 packed, S32 ray_index, S32 count,
                             F32 raw_rgb_in, F32 raw_density_in, int32_t in_dim, int32_t width, int32_t depth, int32_t skip,
                             int32_t cond_dim, int32_t cond_width, int32_t precision, int32_t num_rays, int32_t num_samples,
                             int32_t accumulate, RF32 raw_rgb, RF32 raw_density, RU8 saved, RU8 workspace) {
  (void)raw_rgb_in; (void)raw_density_in;
  DurfMlpArgs a{};
  a.topo = Topo(in_dim, width, depth, skip, cond_dim, cond_width);
  a.precision = precision; a.M = num_rays; a.N = num_samples;
  a.features = features.untyped_data(); a.cond = cond.typed_data(); a.params = params.typed_data();
  a.packed = packed.element_count() ? packed.untyped_data() : nullptr;
  a.ray_index = OrNull(ray_index); a.count = OrNull(count); a.accumulate = accumulate;
  a.raw_rgb = raw_rgb->typed_data(); a.raw_density = raw_density->typed_data();
  a.saved = saved->element_count() ? saved->untyped_data() : nullptr;
  a.workspace = workspace->untyped_data(); a.workspace_bytes = workspace->size_bytes();
  return Check(durf_mlp_fwd(stream, &a));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfMlpFwd, MlpFwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::AnyBuffer>().Arg<F32>().Arg<F32>()
                                  .Arg<U8>().Arg<S32>().Arg<S32>().Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("in_dim").Attr<int32_t>("width").Attr<int32_t>("depth").Attr<int32_t>("skip")
                                  .Attr<int32_t>("cond_dim").Attr<int32_t>("cond_width").Attr<int32_t>("precision")
                                  .Attr<int32_t>("num_rays").Attr<int32_t>("num_samples").Attr<int32_t>("accumulate")
                                  .Ret<F32>().Ret<F32>().Ret<U8>().Ret<U8>());

// The same forward with the ray-march FUSED into the kernel (DurfMlpArgs.fused_raymarch, BF16 only): the MLP kernel
// generates its input tiles from the rays (mip.py:155-282, mip360.py:47-79 inside K2), so no feature tensor exists between the
// two reference calls.  t_vals is an operand aliased to the third result (written when DURF_RM_SAMPLE, else read);
// `features_out` (0 elements = not wanted) receives the generated bf16 tiles for the backward handler.
static ffi::Error MlpFwdFusedImpl(cudaStream_t stream, F32 origins, F32 dirs, F32 radii, F32 near, F32 far, F32 t_rand, F32 ray_mult,
                                  F32 t_vals_in, F32 cond, F32 params, U8 packed, S32 ray_index, S32 count, F32 raw_rgb_in,
                                  F32 raw_density_in, int32_t in_dim, int32_t width, int32_t depth, int32_t skip, int32_t cond_dim,
                                  int32_t cond_width, int32_t num_rays, int32_t min_deg, int32_t max_deg, int32_t flags, float alpha,
                                  int32_t accumulate, RF32 raw_rgb, RF32 raw_density, RF32 t_vals, ffi::Result<ffi::AnyBuffer> features_out,
                                  RU8 saved) {
  (void)t_vals_in; (void)raw_rgb_in; (void)raw_density_in;
  DurfRaymarchArgs r{};
  FillRaymarch(r, origins, dirs, radii, near, far, t_rand, ray_mult, ray_index, count, 128, min_deg, max_deg, flags, alpha);
  r.t_vals = t_vals->typed_data();
  DurfMlpArgs a{};
  a.topo = Topo(in_dim, width, depth, skip, cond_dim, cond_width);
  a.precision = DURF_PREC_BF16; a.M = num_rays; a.N = 128;
  a.features = features_out->element_count() ? features_out->untyped_data() : nullptr;
  a.cond = cond.typed_data(); a.params = params.typed_data(); a.packed = packed.untyped_data();
  a.ray_index = OrNull(ray_index); a.count = OrNull(count); a.accumulate = accumulate;
  a.raw_rgb = raw_rgb->typed_data(); a.raw_density = raw_density->typed_data();
  a.saved = saved->element_count() ? saved->untyped_data() : nullptr;
  a.fused_raymarch = &r;
  return Check(durf_mlp_fwd(stream, &a));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfMlpFwdFused, MlpFwdFusedImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<U8>().Arg<S32>().Arg<S32>()
                                  .Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("in_dim").Attr<int32_t>("width").Attr<int32_t>("depth").Attr<int32_t>("skip")
                                  .Attr<int32_t>("cond_dim").Attr<int32_t>("cond_width").Attr<int32_t>("num_rays")
                                  .Attr<int32_t>("min_deg").Attr<int32_t>("max_deg").Attr<int32_t>("flags").Attr<float>("alpha")
                                  .Attr<int32_t>("accumulate")
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<ffi::AnyBuffer>().Ret<U8>());

// Backward: d raw_rgb, d raw_density -> d params (zero-initialised operand aliased to the result: the library accumulates)
// and, for the box-pose path, d features fp32 [M*N, in_dim] (0-sized result = not wanted).
static ffi::Error MlpBwdImpl(cudaStream_t stream, ffi::AnyBuffer features, F32 cond, F32 params, U8 packed, S32 ray_index, S32 count,
                             U8 saved, F32 d_raw_rgb, F32 d_raw_density, F32 d_params_init, int32_t in_dim, int32_t width, int32_t depth,
                             int32_t skip, int32_t cond_dim, int32_t cond_width, int32_t precision, int32_t num_rays,
                             int32_t num_samples, RF32 d_params, RF32 d_features, RU8 workspace) {
  (void)d_params_init;
  DurfMlpArgs a{};
  a.topo = Topo(in_dim, width, depth, skip, cond_dim, cond_width);
  a.precision = precision; a.M = num_rays; a.N = num_samples;
  a.features = features.untyped_data(); a.cond = cond.typed_data(); a.params = params.typed_data();
  a.packed = packed.element_count() ? packed.untyped_data() : nullptr;
  a.ray_index = OrNull(ray_index); a.count = OrNull(count);
  a.raw_rgb = const_cast<float*>(d_raw_rgb.typed_data());          // not written by the backward; must be non-null
  a.raw_density = const_cast<float*>(d_raw_density.typed_data());
  a.saved = const_cast<void*>(saved.untyped_data());
  a.workspace = workspace->untyped_data(); a.workspace_bytes = workspace->size_bytes();
  return Check(durf_mlp_bwd(stream, &a, d_raw_rgb.typed_data(), d_raw_density.typed_data(), d_params->typed_data(),
                            d_features->element_count() ? d_features->typed_data() : nullptr));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfMlpBwd, MlpBwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::AnyBuffer>().Arg<F32>().Arg<F32>()
                                  .Arg<U8>().Arg<S32>().Arg<S32>().Arg<U8>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("in_dim").Attr<int32_t>("width").Attr<int32_t>("depth").Attr<int32_t>("skip")
                                  .Attr<int32_t>("cond_dim").Attr<int32_t>("cond_width").Attr<int32_t>("precision")
                                  .Attr<int32_t>("num_rays").Attr<int32_t>("num_samples").Ret<F32>().Ret<F32>().Ret<U8>());

// ---- K3: activations + mip.volumetric_rendering (obbpose_model.py:243-245; mip.py:285-327) --------------------------------
static void FillComposite(DurfCompositeArgs& a, F32& raw_rgb, F32& raw_density, F32& t_vals, F32& dirs, int32_t white_bkgd,
                          int32_t rand_bkgd, float density_bias) {
  a.B = Dim(raw_density, 0); a.N = Dim(raw_density, 1);
  a.white_bkgd = white_bkgd; a.rand_bkgd = rand_bkgd; a.activated = 0; a.density_bias = density_bias;
  a.raw_rgb = raw_rgb.typed_data(); a.raw_density = raw_density.typed_data(); a.t_vals = t_vals.typed_data(); a.dirs = dirs.typed_data();
}
static ffi::Error CompositeFwdImpl(cudaStream_t stream, F32 raw_rgb, F32 raw_density, F32 t_vals, F32 dirs, int32_t white_bkgd,
                                   int32_t rand_bkgd, float density_bias, RF32 comp_rgb, RF32 depth, RF32 acc, RF32 weights,
                                   RF32 t_mids, RF32 t_dists) {
  DurfCompositeArgs a{};
  FillComposite(a, raw_rgb, raw_density, t_vals, dirs, white_bkgd, rand_bkgd, density_bias);
  a.comp_rgb = comp_rgb->typed_data(); a.depth = depth->typed_data(); a.acc = acc->typed_data();
  a.weights = weights->typed_data(); a.t_mids = t_mids->typed_data(); a.t_dists = t_dists->typed_data();
  return Check(durf_composite_fwd(stream, &a));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfCompositeFwd, CompositeFwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("white_bkgd").Attr<int32_t>("rand_bkgd").Attr<float>("density_bias")
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());

// cotangents of (comp_rgb, depth, acc, weights) -> d raw_rgb, d raw_density, d dirs_s (t_vals are stop_gradient'ed, mip.py:413-414)
static ffi::Error CompositeBwdImpl(cudaStream_t stream, F32 raw_rgb, F32 raw_density, F32 t_vals, F32 dirs, F32 d_comp_rgb, F32 d_depth,
                                   F32 d_acc, F32 d_weights, int32_t white_bkgd, int32_t rand_bkgd, float density_bias,
                                   RF32 d_raw_rgb, RF32 d_raw_density, RF32 d_dirs) {
  DurfCompositeArgs a{};
  FillComposite(a, raw_rgb, raw_density, t_vals, dirs, white_bkgd, rand_bkgd, density_bias);
  return Check(durf_composite_bwd(stream, &a, d_comp_rgb.typed_data(), d_depth.typed_data(), OrNull(d_acc), d_weights.typed_data(),
                                  d_raw_rgb->typed_data(), d_raw_density->typed_data(), d_dirs->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfCompositeBwd, CompositeBwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<int32_t>("white_bkgd").Attr<int32_t>("rand_bkgd").Attr<float>("density_bias")
                                  .Ret<F32>().Ret<F32>().Ret<F32>());

// ---- K4: mip.resample_along_rays + math.sorted_piecewise_constant_pdf (mip.py:393-412; math.py:222-284) --------------------
static ffi::Error ResampleFwdImpl(cudaStream_t stream, F32 t_vals, F32 weights, F32 u_rand, float resample_padding, int32_t blurpool,
                                  RF32 new_t_vals) {
  return Check(durf_resample_fwd(stream, Dim(weights, 0), Dim(weights, 1), t_vals.typed_data(), weights.typed_data(), OrNull(u_rand),
                                 resample_padding, blurpool, static_cast<int32_t>(new_t_vals->dimensions()[1]),
                                 new_t_vals->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfResampleFwd, ResampleFwdImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<float>("resample_padding").Attr<int32_t>("blurpool").Ret<F32>());
