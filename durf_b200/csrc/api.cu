// C-ABI glue: library info, error state, MLP parameter layout and the precision dispatch of durf_mlp_*.
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "mlp_topology.h"

namespace durf {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

// implemented in mlp_fp32.cu / mlp_tc.cu
size_t mlp_fp32_saved_floats(const DurfMlpTopology& t, int64_t R);
size_t mlp_fp32_infer_floats(const DurfMlpTopology& t, int64_t R);
size_t mlp_fp32_bwd_floats(const DurfMlpTopology& t, int64_t R);
int mlp_fp32_forward(cudaStream_t st, const DurfMlpArgs& a);
int mlp_fp32_backward(cudaStream_t st, const DurfMlpArgs& a, const float* d_raw_rgb, const float* d_raw_density,
                      float* d_params, float* d_features);
int64_t mlp_tc_packed_bytes(const DurfMlpTopology& t);
int mlp_tc_pack_blocks(const DurfMlpTopology& t, const float* params, void* packed, std::vector<PackBlock>& out);
int mlp_tc_forward(cudaStream_t st, const DurfMlpArgs& a);
size_t mlp_tc_workspace_bytes(const DurfMlpTopology& t, int64_t M);
bool mlp_tc_bwd_supported(const DurfMlpTopology& t);
int mlp_tc_saved_blocks(const DurfMlpTopology& t);
int64_t mlp_tc_packed_t_bytes(const DurfMlpTopology& t);
int mlp_tc_pack_t_blocks(const DurfMlpTopology& t, const float* params, void* packed_t, std::vector<PackBlock>& out);
int mlp_tc_backward(cudaStream_t st, const DurfMlpArgs& a, const float* d_raw_rgb, const float* d_raw_density, float* d_params,
                    float* d_features);
int mlp_tc_backward_data(cudaStream_t st, const DurfMlpArgs& a, const float* d_raw_rgb, const float* d_raw_density,
                         float* d_features, int32_t* tile_done, int max_ctas);
int mlp_tc_backward_weights(cudaStream_t st, const DurfMlpArgs& a, const float* d_raw_rgb, const float* d_raw_density,
                            float* d_params, const int32_t* tile_done, int max_ctas);

}  // namespace durf

using namespace durf;

extern "C" const char* durf_version(void) { return "durf_b200 0.1.0 (sm_100a)"; }
extern "C" const char* durf_last_error(void) { return g_err; }
extern "C" int64_t durf_launch_count(void) { return g_launches; }
extern "C" void durf_reset_launch_count(void) { g_launches = 0; }

extern "C" int64_t durf_mlp_param_count(const DurfMlpTopology* topo) {
  if (!topology_ok(topo)) { set_error("durf_mlp_param_count: bad topology"); return DURF_E_INVALID; }
  return MlpLayout(*topo).total;
}

extern "C" int64_t durf_mlp_param_offset(const DurfMlpTopology* topo, int32_t layer, int32_t* in_dim, int32_t* out_dim) {
  if (!topology_ok(topo)) { set_error("durf_mlp_param_offset: bad topology"); return DURF_E_INVALID; }
  MlpLayout L(*topo);
  if (layer < 0 || layer >= L.n_layers) { set_error("durf_mlp_param_offset: layer %d out of range", layer); return DURF_E_INVALID; }
  if (in_dim) *in_dim = L.in_dim[layer];
  if (out_dim) *out_dim = L.out_dim[layer];
  return L.w_off[layer];
}

extern "C" int64_t durf_mlp_packed_bytes(const DurfMlpTopology* topo) {
  if (!topology_ok(topo)) { set_error("durf_mlp_packed_bytes: bad topology"); return DURF_E_INVALID; }
  return mlp_tc_packed_bytes(*topo) + mlp_tc_packed_t_bytes(*topo);   // forward image, then the transposed image of dgrad
}

extern "C" int durf_mlp_pack_weights_multi(durf_stream_t stream, int32_t n, const DurfMlpTopology* topos, const float* const* params,
                                           void* const* packed) {
  DURF_REQUIRE(n >= 0 && (n == 0 || (topos && params && packed)), DURF_E_INVALID, "durf_mlp_pack_weights: bad argument");
  std::vector<PackBlock> blocks;
  for (int i = 0; i < n; ++i) {
    DURF_REQUIRE(topology_ok(&topos[i]) && params[i] && packed[i], DURF_E_INVALID, "durf_mlp_pack_weights: bad argument (network %d)", i);
    int rc = mlp_tc_pack_blocks(topos[i], params[i], packed[i], blocks);
    if (rc < 0) return rc;
    if (!mlp_tc_bwd_supported(topos[i])) continue;
    rc = mlp_tc_pack_t_blocks(topos[i], params[i], (uint8_t*)packed[i] + mlp_tc_packed_bytes(topos[i]), blocks);
    if (rc < 0) return rc;
  }
  return pack_blocks_launch((cudaStream_t)stream, blocks.data(), (int)blocks.size());
}

extern "C" int durf_mlp_pack_weights(durf_stream_t stream, const DurfMlpTopology* topo, const float* params, void* packed) {
  return durf_mlp_pack_weights_multi(stream, 1, topo, &params, &packed);
}

extern "C" size_t durf_mlp_workspace_bytes(const DurfMlpTopology* topo, int32_t precision, int32_t M, int32_t N, int32_t training) {
  if (!topology_ok(topo) || M < 0 || N < 1) return 0;
  const int64_t R = (int64_t)M * N;
  if (precision == DURF_PREC_FP32)
    return sizeof(float) * (training ? mlp_fp32_bwd_floats(*topo, R) : mlp_fp32_infer_floats(*topo, R));
  // forward: nothing (the per-tile view bias of the condition layer is formed inside the kernel); backward: the dz tile
  // records read by wgrad
  if (training) return mlp_tc_bwd_supported(*topo) ? (size_t)M * mlp_tc_saved_blocks(*topo) * 16384 : 0;
  return mlp_tc_workspace_bytes(*topo, M);
}

extern "C" size_t durf_mlp_saved_bytes(const DurfMlpTopology* topo, int32_t precision, int32_t M, int32_t N) {
  if (!topology_ok(topo) || M < 0 || N < 1) return 0;
  if (precision == DURF_PREC_FP32) return sizeof(float) * mlp_fp32_saved_floats(*topo, (int64_t)M * N);
  // bf16 activations of every layer as tile images, then the 1-bit ReLU masks of the trunk layers and the condition layer (4 bytes per row and
  // 32-column group)
  return mlp_tc_bwd_supported(*topo)
             ? (size_t)M * mlp_tc_saved_blocks(*topo) * 16384 + (size_t)M * (topo->depth + 1) * (topo->width / 32) * 128 * 4
             : 0;
}

static int check_mlp(const DurfMlpArgs* a, const char* who) {
  DURF_REQUIRE(a != nullptr, DURF_E_INVALID, "%s: null args", who);
  DURF_REQUIRE(topology_ok(&a->topo), DURF_E_INVALID, "%s: bad topology", who);
  DURF_REQUIRE(a->M >= 0 && a->N >= 1, DURF_E_INVALID, "%s: bad shape M=%d N=%d", who, a->M, a->N);
  DURF_REQUIRE(a->fused_raymarch == nullptr || a->precision == DURF_PREC_BF16, DURF_E_UNSUPPORTED,
               "%s: the fused ray-march (in-kernel feature generation) exists on the tensor-core path only", who);
  DURF_REQUIRE((a->features || a->fused_raymarch) && a->cond && a->params && a->raw_rgb && a->raw_density, DURF_E_INVALID,
               "%s: null buffer", who);
  DURF_REQUIRE(a->precision == DURF_PREC_FP32 || a->precision == DURF_PREC_BF16, DURF_E_INVALID, "%s: unknown precision %d",
               who, a->precision);
  DURF_REQUIRE(a->accumulate >= 0 && a->accumulate <= 2, DURF_E_INVALID, "%s: accumulate must be 0, 1 or 2 (got %d)", who, a->accumulate);
  return DURF_OK;
}

extern "C" int durf_mlp_fwd(durf_stream_t stream, const DurfMlpArgs* args) {
  int rc = check_mlp(args, "durf_mlp_fwd");
  if (rc != DURF_OK) return rc;
  if (args->M == 0) return DURF_OK;
  if (args->precision == DURF_PREC_FP32) return mlp_fp32_forward((cudaStream_t)stream, *args);
  return mlp_tc_forward((cudaStream_t)stream, *args);
}

extern "C" int durf_mlp_bwd(durf_stream_t stream, const DurfMlpArgs* args, const float* d_raw_rgb,
                            const float* d_raw_density, float* d_params, float* d_features) {
  int rc = check_mlp(args, "durf_mlp_bwd");
  if (rc != DURF_OK) return rc;
  DURF_REQUIRE(d_raw_rgb && d_raw_density && d_params, DURF_E_INVALID, "durf_mlp_bwd: null gradient buffer");
  if (args->M == 0) return DURF_OK;
  if (args->precision == DURF_PREC_FP32)
    return mlp_fp32_backward((cudaStream_t)stream, *args, d_raw_rgb, d_raw_density, d_params, d_features);
  return mlp_tc_backward((cudaStream_t)stream, *args, d_raw_rgb, d_raw_density, d_params, d_features);
}

extern "C" size_t durf_mlp_bwd_flags_bytes(const DurfMlpTopology* topo, int32_t M) {
  if (!topology_ok(topo) || M < 0) return 0;
  return (size_t)M * (topo->depth + 2) * sizeof(int32_t);
}

extern "C" int durf_mlp_bwd_data(durf_stream_t stream, const DurfMlpArgs* args, const float* d_raw_rgb, const float* d_raw_density,
                                 float* d_features, int32_t* tile_done, int32_t max_ctas) {
  int rc = check_mlp(args, "durf_mlp_bwd_data");
  if (rc != DURF_OK) return rc;
  DURF_REQUIRE(args->precision == DURF_PREC_BF16, DURF_E_UNSUPPORTED, "durf_mlp_bwd_data: tensor-core (BF16) path only");
  DURF_REQUIRE(d_raw_rgb && d_raw_density, DURF_E_INVALID, "durf_mlp_bwd_data: null gradient buffer");
  if (args->M == 0) return DURF_OK;
  return mlp_tc_backward_data((cudaStream_t)stream, *args, d_raw_rgb, d_raw_density, d_features, tile_done, max_ctas);
}

extern "C" int durf_mlp_bwd_weights(durf_stream_t stream, const DurfMlpArgs* args, const float* d_raw_rgb, const float* d_raw_density,
                                    float* d_params, const int32_t* tile_done, int32_t max_ctas) {
  int rc = check_mlp(args, "durf_mlp_bwd_weights");
  if (rc != DURF_OK) return rc;
  DURF_REQUIRE(args->precision == DURF_PREC_BF16, DURF_E_UNSUPPORTED, "durf_mlp_bwd_weights: tensor-core (BF16) path only");
  DURF_REQUIRE(d_raw_rgb && d_raw_density && d_params, DURF_E_INVALID, "durf_mlp_bwd_weights: null gradient buffer");
  if (args->M == 0) return DURF_OK;
  return mlp_tc_backward_weights((cudaStream_t)stream, *args, d_raw_rgb, d_raw_density, d_params, tile_done, max_ctas);
}
