// XLA typed-FFI shims over the C ABI of libdurf_b200.so (include/durf_b200.h).
//
// This is the reference-side binding a DURF maintainer adds to reach the CUDA hot path from JAX: each handler takes the
// stream XLA executes on and the device buffers XLA owns, and forwards them, unchanged, to one C-ABI entry point.
// It cannot be compiled in the build image of this repository (no jaxlib, hence no xla/ffi/api/ffi.h), so the whole
// file is guarded; with jaxlib installed:
//     g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//         -I../../include durf_ffi.cc -L../../durf_b200 -ldurf_b200 -o libdurf_jax_ffi.so
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define DURF_HAVE_XLA_FFI 1
#endif
#endif

#ifdef DURF_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include "durf_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Check(int rc) {
  if (rc == DURF_OK) return ffi::Error::Success();
  return ffi::Error(rc == DURF_E_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, durf_last_error());
}

// mip.sample_along_rays / resample-output -> cast_rays -> mip360.new_space -> integrated_pos_enc | weighted_ipe
// (internal/mip.py:330-370,155-179,226-282,182-223; internal/mip360.py:63-79).  flags = DURF_RM_* bits.
static ffi::Error RaymarchFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> origins, ffi::Buffer<ffi::F32> dirs,
                                  ffi::Buffer<ffi::F32> radii, ffi::Buffer<ffi::F32> near, ffi::Buffer<ffi::F32> far,
                                  ffi::Buffer<ffi::F32> t_rand, ffi::Buffer<ffi::F32> ray_mult, int32_t num_samples,
                                  int32_t min_deg, int32_t max_deg, int32_t flags, float alpha,
                                  ffi::ResultBuffer<ffi::F32> t_vals, ffi::ResultBuffer<ffi::F32> features) {
  DurfRaymarchArgs a{};
  a.B = static_cast<int32_t>(origins.dimensions()[0]);
  a.N = num_samples; a.min_deg = min_deg; a.max_deg = max_deg; a.flags = static_cast<uint32_t>(flags); a.alpha = alpha;
  a.origins = origins.typed_data(); a.dirs = dirs.typed_data(); a.radii = radii.typed_data();
  a.near = near.typed_data(); a.far = far.typed_data(); a.t_rand = t_rand.typed_data();
  a.ray_mult = ray_mult.element_count() ? ray_mult.typed_data() : nullptr;
  a.t_vals = t_vals->typed_data(); a.features = features->typed_data();
  return Check(durf_raymarch_fwd(stream, &a));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfRaymarchFwd, RaymarchFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("num_samples").Attr<int32_t>("min_deg").Attr<int32_t>("max_deg")
                                  .Attr<int32_t>("flags").Attr<float>("alpha")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());

// obbpose_model.py:243-245 + mip.volumetric_rendering (internal/mip.py:285-327)
static ffi::Error CompositeFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> raw_rgb, ffi::Buffer<ffi::F32> raw_density,
                                   ffi::Buffer<ffi::F32> t_vals, ffi::Buffer<ffi::F32> dirs, int32_t white_bkgd,
                                   int32_t rand_bkgd, float density_bias, ffi::ResultBuffer<ffi::F32> comp_rgb,
                                   ffi::ResultBuffer<ffi::F32> depth, ffi::ResultBuffer<ffi::F32> acc,
                                   ffi::ResultBuffer<ffi::F32> weights, ffi::ResultBuffer<ffi::F32> t_mids,
                                   ffi::ResultBuffer<ffi::F32> t_dists) {
  DurfCompositeArgs a{};
  a.B = static_cast<int32_t>(raw_density.dimensions()[0]);
  a.N = static_cast<int32_t>(raw_density.dimensions()[1]);
  a.white_bkgd = white_bkgd; a.rand_bkgd = rand_bkgd; a.activated = 0; a.density_bias = density_bias;
  a.raw_rgb = raw_rgb.typed_data(); a.raw_density = raw_density.typed_data(); a.t_vals = t_vals.typed_data();
  a.dirs = dirs.typed_data(); a.comp_rgb = comp_rgb->typed_data(); a.depth = depth->typed_data(); a.acc = acc->typed_data();
  a.weights = weights->typed_data(); a.t_mids = t_mids->typed_data(); a.t_dists = t_dists->typed_data();
  return Check(durf_composite_fwd(stream, &a));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfCompositeFwd, CompositeFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("white_bkgd").Attr<int32_t>("rand_bkgd").Attr<float>("density_bias")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());

// mip.resample_along_rays + math.sorted_piecewise_constant_pdf (internal/mip.py:393-412; internal/math.py:222-284)
static ffi::Error ResampleFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> t_vals, ffi::Buffer<ffi::F32> weights,
                                  ffi::Buffer<ffi::F32> u_rand, float resample_padding, int32_t blurpool,
                                  ffi::ResultBuffer<ffi::F32> new_t_vals) {
  const int32_t B = static_cast<int32_t>(weights.dimensions()[0]), N = static_cast<int32_t>(weights.dimensions()[1]);
  const int32_t S = static_cast<int32_t>(new_t_vals->dimensions()[1]);
  return Check(durf_resample_fwd(stream, B, N, t_vals.typed_data(), weights.typed_data(),
                                 u_rand.element_count() ? u_rand.typed_data() : nullptr, resample_padding, blurpool, S,
                                 new_t_vals->typed_data()));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfResampleFwd, ResampleFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("resample_padding").Attr<int32_t>("blurpool")
                                  .Ret<ffi::Buffer<ffi::F32>>());

// MLP.__call__ / BoxMLP.__call__ (internal/obbpose_model.py:294-354, 358-418).  XLA supplies the workspace as an extra
// result buffer (sized with durf_mlp_workspace_bytes at trace time); `packed` is the tensor-core weight image.
static ffi::Error MlpFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::BF16> feature_tiles, ffi::Buffer<ffi::F32> cond,
                             ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::U8> packed, int32_t in_dim, int32_t width,
                             int32_t depth, int32_t skip, int32_t cond_dim, int32_t cond_width, int32_t num_rays,
                             ffi::ResultBuffer<ffi::F32> raw_rgb, ffi::ResultBuffer<ffi::F32> raw_density,
                             ffi::ResultBuffer<ffi::U8> workspace) {
  DurfMlpArgs a{};
  a.topo = DurfMlpTopology{in_dim, width, depth, skip, cond_dim, cond_width};
  a.precision = DURF_PREC_BF16; a.M = num_rays; a.N = 128;
  a.features = feature_tiles.untyped_data(); a.cond = cond.typed_data(); a.params = params.typed_data();
  a.packed = packed.untyped_data(); a.raw_rgb = raw_rgb->typed_data(); a.raw_density = raw_density->typed_data();
  a.workspace = workspace->untyped_data(); a.workspace_bytes = workspace->size_bytes();
  return Check(durf_mlp_fwd(stream, &a));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(DurfMlpFwd, MlpFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Attr<int32_t>("in_dim").Attr<int32_t>("width").Attr<int32_t>("depth").Attr<int32_t>("skip")
                                  .Attr<int32_t>("cond_dim").Attr<int32_t>("cond_width").Attr<int32_t>("num_rays")
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>());
#endif  // DURF_HAVE_XLA_FFI
