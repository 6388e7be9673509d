// K2 -- the radiance/density MLP as a fused tcgen05 GEMM chain (sm_100a), forward.
// Replaces MLP.__call__ / BoxMLP.__call__ (obbpose_model.py:305-354, 369-418) for width 256 / 128, cond_width 128.
//
// One persistent CTA per SM; one tile = the 128 samples of one ray-level = UMMA M.  Everything between the input
// features and the raw outputs stays on chip:
//   * accumulators (fp32, 128 lanes x W columns) AND the activations live in TENSOR MEMORY: the epilogue packs
//     relu(acc + bias) to bf16 and writes it back to TMEM with tcgen05.st, and the next layer's tcgen05.mma takes its
//     A operand straight from TMEM (".ts" form).  Two activation buffers alternate by layer, so shared memory is free
//     for a deep ring of weight chunks;
//   * every layer is issued as two N-halves.  The epilogue of half 0 runs while the tensor core computes half 1, and the
//     epilogue hands the next layer's A operand over per 64-column K block, so the next layer starts on the blocks of half 0
//     while the epilogue of half 1 is still packing;
//   * the MMA issuer never executes a wait between two groups of MMAs (a wait there drains the tensor queue): every K block
//     probes the barriers of the NEXT K block before its own MMAs and consumes the outcome after them (umma_kblock_conv),
//     and the same asm block carries the tcgen05.commit's (accumulator ready, ring stage free).
// Warp roles (384 threads): warp 0 = weight producer (cp.async.bulk of pre-tiled, pre-swizzled bf16 chunks into an
// mbarrier ring; in a CTA pair each CTA fetches half of a stage and multicasts it), warp 1 = MMA issuer (whole warp walks
// the schedule, one elected lane issues), warps 2-3 = the input side: warp 2 allocates TMEM; with DurfMlpArgs.fused_raymarch
// the 64 threads GENERATE the next tile's A operand in shared memory (sampling, conical-frustum Gaussian, contraction, IPE:
// raymarch_device.cuh), otherwise one of them loads the tile image; in both cases they form the next tile's view bias,
// warps 4-11 = epilogue (TMEM lane quarter = warp % 4; the two warps of a quarter split the columns of a half).
// The skip connection (obbpose_model.py:332-333) is an extra K block read from the still-resident input tile (smem,
// ".ss" form); the view direction (constant along a ray) enters the condition layer as a per-tile fp32 bias
// (b + W_view^T enc, view_bias below), so its 27 input columns never occupy tensor-core K; the density head (N=1) and the
// rgb head (N=3) are dot products inside the epilogues.  Training (SAVE) additionally stores every layer's bf16 activations,
// 1-bit ReLU masks and the generated tile for the backward kernels (save_piece).
#include <stdlib.h>

#include <vector>

#include "tc_common.cuh"
#include "mlp_topology.h"
#include "raymarch_device.cuh"

namespace durf {

constexpr int kInpBytes = 16384;      // input tile image: 128 rows x 64 bf16, K-major SWIZZLE_128B
constexpr int kMaxG = 12;
constexpr int kMaxStages = 8;
constexpr int kMaxSteps = 112;
constexpr int kMaxStageUses = 64;

struct LayerSched {
  int n_halves;     // output columns / (W/2)
  int n_act_kb;     // 64-wide K blocks taken from the activation buffer (TMEM)
  int uses_inp;     // +1 K block from the input-feature tile (layer 0, skip layer)
  int kind;         // 0 relu->act, 1 relu->act + density head, 2 linear->act (bottleneck), 3 condition + rgb head + output
  int bias_off;     // offset (floats) of this layer's bias inside the parameter blob
  int last_inp_use; // 1 if no later layer of the tile reads the input-feature tile
  int drain_sig;    // 1: the epilogue of half 1 signals "accumulator half 1 read out" (the next layer has two halves)
};

// One K block (four K=16 tcgen05.mma of M=128, N=128) of a tile's schedule, in ISSUE ORDER.  A layer with two N-halves is
// issued as  [h0: k0 k1] [h1: k0] [h0: k2 k3] [h1: k1] [h1: k2 k3]  (+ the input-tile block of layer 0 / the skip layer after
// each half's last K block).  K blocks 0-1 of the A operand come from the epilogue of the previous layer's half 0 (long done),
// blocks 2-3 from its half 1, which only STARTS when this layer starts: [h1: k0] gives the tensor core 4 more MMAs of
// independent work before the first step that needs block 2, and half 0 still completes early (after 20 of the 32 MMAs), so
// its epilogue finishes blocks 0-1 of the next layer well before that layer begins.  Every barrier is then complete >= 500
// cycles before the step that needs it, more than the one-step lead of the software-pipelined probe (umma_kblock_conv); the
// [h0: k0..k3][h1: k0..k3] order of round 1 left ~250 cycles for blocks 2-3 and lost ~230 cycles per layer.
struct StepSched {
  int8_t g, nh;         // layer, N-half
  int8_t kb;            // K block of the layer's A operand (activation buffer), unused for inp
  int8_t inp;           // 1: A operand = input-feature tile (shared memory), 0: activation buffer (TMEM)
  int8_t acc0;          // 1: first K block of (g, nh): overwrite the accumulator
  int8_t commit;        // 1: last K block of (g, nh): commit acc_full[nh]
  int8_t stage_first;   // 1: first K block read from its ring stage
  int8_t stage_last;    // 1: last K block read from its ring stage (release it)
  int8_t slot;          // block index inside the ring stage
  int8_t wait_a;        // 0..3: before this step, a_ready(k) must have completed (K block k of the A operand is in TMEM);
                        // 4: the previous layer's epilogue has read accumulator half 1 out (this step overwrites it); -1: nothing
  int8_t pad[2];
};
// One use of a weight-ring stage: `nkb` consecutive 16 KB blocks of the packed image starting at `block0`.
struct StageUse {
  int16_t block0, nkb;
};

struct TcParams {
  const uint8_t* feat;       // bf16 tile images, 16 KB per tile
  const float* cond;         // [B, cond_dim]
  int off_bcond;             // bias of the condition layer inside the parameter blob (its view part is added per tile, see view_bias)
  const float* params;       // fp32 blob
  const uint8_t* packed;     // weight image
  const int32_t* ray_index;
  const int32_t* count;
  uint8_t* saved;            // [opt] training: every layer's bf16 activations as tile images (see mlp_tc_saved_bytes)
  int saved_blocks_per_tile;
  uint32_t* masks;           // [opt] training: 1-bit ReLU masks of the trunk layers and (layer index `depth`) the condition layer,
                             // [tile][depth + 1][W / 32 column groups][row] words
  int M;
  int accumulate;
  float* raw_rgb;
  float* raw_density;
  int G;                     // GEMM layers per tile: depth + 2
  int depth;
  int cond_dim;
  int n_steps;               // K-block steps per tile
  int n_uses;                // ring-stage uses per tile
  int off_wden, off_bden, off_wrgb, off_brgb, off_wview;
  int trace;                 // DURF_TC_TRACE=1: block 0 prints where its MMA thread and one epilogue thread spent their cycles
  // N1: the input tile is GENERATED inside the kernel by warps 2-3 (fenceposts -> conical-frustum Gaussian ->
  // mask -> contraction -> IPE, the arithmetic of raymarch.cu's bf16 path) instead of loaded from `feat`
  int gen;
  uint32_t rm_flags;
  int min_deg;
  float alpha;
  const float* alpha_dev;
  const float* g_origins;    // [B,3] origins_s
  const float* g_dirs;       // [B,3] dirs_s
  const float* g_radii;      // [B]
  const float* g_near;       // [B] (DURF_RM_SAMPLE)
  const float* g_far;
  const float* g_t_rand;     // [B,129] (DURF_RM_RANDOMIZED)
  const float* g_ray_mult;   // [opt] [B]
  float* g_t_vals;           // [B,129]: written when DURF_RM_SAMPLE, else read
  uint8_t* feat_out;         // [opt] the generated tiles are also stored here (training: wgrad reads them)
  LayerSched sched[kMaxG];
  StepSched steps[kMaxSteps];   // host side (pack kernel, debugging)
  uint32_t step_w[kMaxSteps];   // the same, one packed word per step: what the MMA issuer reads (see step_word)
  StageUse uses[kMaxStageUses];
};

// The K-block steps of ONE layer in issue order, as a compile-time table: the host builds the packed weight image and the
// ring-stage list from it (build_sched), the MMA issuer unrolls it (issue_layer) - so that every operand of a tcgen05
// instruction is a compile-time function of a few loop-carried uniform values.  (A fully table-driven issuer, one decoded
// descriptor per step, was measured at 580 cycles per 4 MMAs instead of 316: the issuing thread has no latency hiding, and
// a ~250-cycle dependent decode chain per step is not hidden behind 4 MMAs.)
struct LayerSteps {
  int n;
  int8_t nh[12], kb[12], inp[12], acc0[12], commit[12], wait[12];
};
__host__ __device__ constexpr LayerSteps layer_steps(int nh, int kbs, bool inp, bool drained_wait) {
  LayerSteps L{};
  auto add = [&](int h, int kb, bool is_inp, bool first, bool last, int wait) {
    L.nh[L.n] = (int8_t)h; L.kb[L.n] = (int8_t)kb; L.inp[L.n] = is_inp ? 1 : 0; L.acc0[L.n] = first ? 1 : 0;
    L.commit[L.n] = last ? 1 : 0; L.wait[L.n] = (int8_t)wait; ++L.n;
  };
  if (nh == 2 && kbs == 4) {
    add(0, 0, false, true, false, 0); add(0, 1, false, false, false, 1);
    add(1, 0, false, true, false, drained_wait ? 4 : -1);
    add(0, 2, false, false, false, 2); add(0, 3, false, false, !inp, 3);
    if (inp) add(0, 0, true, false, true, -1);
    add(1, 1, false, false, false, -1); add(1, 2, false, false, false, -1); add(1, 3, false, false, !inp, -1);
    if (inp) add(1, 0, true, false, true, -1);
  } else {
    // one N-half (width 128, condition layer) or no activation input (layer 0): half after half, K blocks in order
    for (int h = 0; h < nh; ++h) {
      for (int kb = 0; kb < kbs; ++kb) add(h, kb, false, kb == 0, kb == kbs - 1 && !inp, h == 0 ? kb : -1);
      if (inp) add(h, 0, true, kbs == 0, true, -1);
    }
  }
  return L;
}

// bits: g 0-3 | nh 4 | kb 5-6 | inp 7 | acc0 8 | commit 9 | stage_first 10 | stage_last 11 | slot 12-13 | wait_a + 1 14-16
__host__ __device__ __forceinline__ uint32_t step_word(const StepSched& sp) {
  return (uint32_t)sp.g | ((uint32_t)sp.nh << 4) | ((uint32_t)sp.kb << 5) | ((uint32_t)sp.inp << 7) | ((uint32_t)sp.acc0 << 8) |
         ((uint32_t)sp.commit << 9) | ((uint32_t)sp.stage_first << 10) | ((uint32_t)sp.stage_last << 11) | ((uint32_t)sp.slot << 12) |
         ((uint32_t)(sp.wait_a + 1) << 14);
}

// Epilogue arithmetic of one 32-column group, specialised per layer kind so that the unrolled body has no branches:
// KIND 0: relu(acc + bias) -> bf16; KIND 1: the same + fp32 partial dot product with the density head; KIND 2: linear.
template <int KIND>
__device__ __forceinline__ void epi_pack(const uint32_t (&v)[32], const float4 (&b4)[8], uint32_t wden_addr, float& den,
                                         uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 a = add2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), make_float2(b4[j].x, b4[j].y));
    const float2 b = add2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]), make_float2(b4[j].z, b4[j].w));
    if (KIND == 0) {
      pk[2 * j] = cvt_bf16x2_relu(a.x, a.y);
      pk[2 * j + 1] = cvt_bf16x2_relu(b.x, b.y);
    } else if (KIND == 1) {
      const float4 w4 = lds128(wden_addr + 16 * j);
      const float r0 = fmaxf(a.x, 0.f), r1 = fmaxf(a.y, 0.f), r2 = fmaxf(b.x, 0.f), r3 = fmaxf(b.y, 0.f);
      den = fmaf(r0, w4.x, den); den = fmaf(r1, w4.y, den); den = fmaf(r2, w4.z, den); den = fmaf(r3, w4.w, den);
      pk[2 * j] = cvt_bf16x2(r0, r1);
      pk[2 * j + 1] = cvt_bf16x2(r2, r3);
    } else {
      pk[2 * j] = cvt_bf16x2(a.x, a.y);
      pk[2 * j + 1] = cvt_bf16x2(b.x, b.y);
    }
  }
}

template <int W, bool SAVE>
struct TcCfg {
  static constexpr int NHALF = W / 128;                   // N-halves per trunk layer (every tcgen05.mma is M=128, N=128)
  static constexpr int KB = W / 64;                       // 64-wide K blocks of an activation row
  static constexpr int CPW = 64;                          // columns one epilogue warp owns inside a half
  // Weight ring.  Inference: one stage per (layer, N-half) chunk (64 KB at W = 256, three of them).  Training (SAVE) gives
  // 64 KB to the per-warp staging of the activation stores, which leaves 128 KB: two whole chunks would mean a chunk's
  // weights can only be requested when the chunk before it has completed (measured: -17 % MMA rate), so the ring is cut into
  // four 32 KB stages of two K blocks, each released by the MMA issuer's commit as soon as its own MMAs are done.
  // Weight ring: a stage holds up to SKB consecutive K blocks of the issue order (never across a layer boundary).  Inference:
  // 64 KB stages at W = 256 (two per layer, three in the ring), i.e. two stage releases and two weight probes per layer - every
  // tcgen05.commit costs the issuing thread ~190 cycles and every barrier probe ~140, and that thread has only the ~2500
  // cycles of a layer's MMAs to spend.  Training (SAVE) gives 64 KB to the per-warp staging of the activation stores, which
  // leaves 128 KB: four 32 KB stages.
  static constexpr int SKB = (W == 256 && !SAVE) ? 4 : 2;
  static constexpr int STAGE_BYTES = SKB * kBlockBytes;
  static constexpr int STAGES = SAVE ? 4 : ((W == 256) ? 3 : 6);
  static constexpr int TMEM_COLS = 2 * W;                 // W accumulator columns + 2 x W/2 activation columns
  static constexpr int ACC_COL = 0;
  static constexpr int ACT_COL = W;                       // buffer b at ACT_COL + b * W/2 (bf16 pairs)
  static constexpr int MAX_BIAS_LAYERS = 10;              // trunk (depth <= 9) + bottleneck
  // shared memory map (bytes, from a 1024-aligned base)
  static constexpr int OFF_INP = 0;
  static constexpr int OFF_RING = OFF_INP + kInpBytes;
  static constexpr int OFF_STG = OFF_RING + STAGES * STAGE_BYTES;     // [2 buffers][8 epilogue warps][4 KB]
  static constexpr int STG_BYTES = SAVE ? 2 * 8 * 4096 : 0;
  static constexpr int OFF_BIAS = OFF_STG + STG_BYTES;                // fp32 [MAX_BIAS_LAYERS][W]
  static constexpr int OFF_WDEN = OFF_BIAS + MAX_BIAS_LAYERS * W * 4; // fp32 [W]
  static constexpr int OFF_WRGB = OFF_WDEN + W * 4;                   // fp32 [3][128]
  static constexpr int OFF_VBIAS = OFF_WRGB + 3 * 128 * 4;            // fp32 [128] this tile's | [2][128] the coming tiles' view bias
  static constexpr int OFF_PART = OFF_VBIAS + 3 * 128 * 4;            // fp32 [4][128] partial density / rgb of the upper column warps
  static constexpr int OFF_MISC = OFF_PART + 4 * 128 * 4;             // head biases [4] + tmem ptr + barriers
  static constexpr int MISC_BYTES = 512;
  static constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;     // + alignment slack
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// State the MMA issuer carries from step to step (all warp-uniform).
struct IssueState {
  uint32_t stage, phase;   // weight-ring position
  uint32_t ar_bits;        // bit k: parity of a_ready(k) (k = 4: "accumulator half 1 drained")
};

// All K-block steps of one layer, unrolled from the compile-time table layer_steps(NH, KBS, INP, DRAINED).  Every step
// issues its four MMAs and, inside the same asm block, probes / waits for what the NEXT step needs (umma_kblock_conv);
// for the layer's last step that is the first step of the next layer: its weights (need_w_after) and a_ready(0)
// (need_a0_after).  `blk0` = index of the layer's first K block inside its ring stage sequence is always 0: stages never
// cross a layer boundary.
template <class C, int W, int NH, int KBS, bool INP, bool DRAINED>
__device__ __forceinline__ void issue_layer(IssueState& st, uint32_t sbase, uint32_t bar0, uint32_t tmem_u, uint32_t inp_lo, int g,
                                            uint32_t need_a0_after, uint32_t need_w_after, uint32_t rel_mask) {
  constexpr LayerSteps LS = layer_steps(NH, KBS, INP, DRAINED);
  constexpr uint32_t idesc = umma_idesc(128, 128);
  constexpr uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);     // SBO | version | SWIZZLE_128B
  auto bar_full = [&](uint32_t s_) { return bar0 + 8 * s_; };
  auto bar_empty = [&](uint32_t s_) { return bar0 + 8 * (kMaxStages + s_); };
  auto bar_acc_full = [&](int h) { return bar0 + 8 * (2 * kMaxStages + 2 + h); };
  auto bar_a_ready = [&](int kb) { return bar0 + 8 * (2 * kMaxStages + 4 + kb); };
  const uint32_t a_buf = tmem_u + C::ACT_COL + (g & 1) * (W / 2);       // written by the epilogue of layer g-1
#pragma unroll
  for (int i = 0; i < LS.n; ++i) {
    constexpr int SKB = C::SKB;
    const int slot = i % SKB;
    const bool stage_last = (slot == SKB - 1) || (i == LS.n - 1);
    const bool last = i == LS.n - 1;
    if (LS.wait[i] >= 0) st.ar_bits ^= 1u << LS.wait[i];     // this step's a_ready was waited for inside the previous step
    tc_fence_after();
    // what the next step needs
    const int nwait = last ? 0 : (LS.wait[i + 1] >= 0 ? LS.wait[i + 1] : 0);
    const uint32_t need_a = last ? need_a0_after : (LS.wait[i + 1] >= 0 ? 1u : 0u);
    const uint32_t need_w = last ? need_w_after : (stage_last ? 1u : 0u);
    const uint32_t next_stage = (st.stage + 1 == (uint32_t)C::STAGES) ? 0 : st.stage + 1;
    const uint32_t next_phase = (st.stage + 1 == (uint32_t)C::STAGES) ? st.phase ^ 1 : st.phase;
    const uint32_t d_addr = tmem_u + C::ACC_COL + LS.nh[i] * 128;
    const uint32_t b_lo = (((sbase + C::OFF_RING + st.stage * C::STAGE_BYTES + slot * kBlockBytes) & 0x3FFFF) >> 4) | (1u << 16);
    if (!LS.inp[i])
      umma_kblock_conv<true>(d_addr, a_buf + LS.kb[i] * 32, b_lo, desc_hi, idesc, LS.acc0[i] ? 0u : 1u,
                             bar_a_ready(nwait), (st.ar_bits >> nwait) & 1u, need_a, bar_full(next_stage), next_phase, need_w,
                             bar_acc_full(LS.nh[i]), LS.commit[i] ? 1u : 0u, bar_empty(st.stage), stage_last ? 1u : 0u, rel_mask);
    else
      umma_kblock_conv<false>(d_addr, inp_lo, b_lo, desc_hi, idesc, LS.acc0[i] ? 0u : 1u,
                              bar_a_ready(nwait), (st.ar_bits >> nwait) & 1u, need_a, bar_full(next_stage), next_phase, need_w,
                              bar_acc_full(LS.nh[i]), LS.commit[i] ? 1u : 0u, bar_empty(st.stage), stage_last ? 1u : 0u, rel_mask);
    if (stage_last) { st.stage = next_stage; st.phase = next_phase; }
  }
}

// Per-tile bias of the condition layer: vb[j] = b_cond[j] + sum_i enc(viewdir of the tile's ray)[i] * W_cond[width + i][j]
// (obbpose_model.py:343-350).  The view direction is constant along a ray, so these 27 input columns never occupy
// tensor-core K; the 64 threads of warps 2-3 form the row for the NEXT tile (two columns each, the same summation order as a
// ray-at-a-time loop) while the current one executes, into the half of a double buffer the epilogue copies from at layer 1.
// The weights of a thread's two columns stay in registers for the whole kernel (re-reading the 13.8 KB from L2 for every tile
// cost 0.8 % of the render rate: the run is power-capped).
struct ViewWeights {          // this thread's two columns (gt, gt + 64) of W_cond[width:, :] and of b_cond, loaded once per CTA
  float w0[32], w1[32], b0, b1;
};
__device__ __forceinline__ void view_weights_load(const TcParams& p, int gt, ViewWeights& vw) {
  const float* __restrict__ w = p.params + p.off_wview;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    vw.w0[i] = i < p.cond_dim ? __ldg(w + i * 128 + gt) : 0.f;
    vw.w1[i] = i < p.cond_dim ? __ldg(w + i * 128 + gt + 64) : 0.f;
  }
  vw.b0 = __ldg(p.params + p.off_bcond + gt);
  vw.b1 = __ldg(p.params + p.off_bcond + gt + 64);
}
__device__ __forceinline__ void view_bias(const TcParams& p, const ViewWeights& vw, int ray, int gt, float* __restrict__ dst) {
  const float* __restrict__ enc = p.cond + (size_t)ray * p.cond_dim;
  float v0 = vw.b0, v1 = vw.b1;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < p.cond_dim) {
      const float c = __ldg(enc + i);
      v0 = fmaf(c, vw.w0[i], v0);
      v1 = fmaf(c, vw.w1[i], v1);
    }
  dst[gt] = v0;
  dst[gt + 64] = v1;
}

template <int W, bool SAVE>
__global__ void __launch_bounds__(384, 1)
mlp_tc_fwd_kernel(const __grid_constant__ TcParams p) {
  using C = TcCfg<W, SAVE>;
  extern __shared__ uint8_t smem_raw[];
  // (aligned through the generic address: with an offset into the __shared__ array the per-tile housekeeping would use
  // ld/st.shared instead of generic accesses, but the inference kernel measured 0.4 % slower that way - kept in dgrad / wgrad)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;      // warp-uniform for the compiler (uniform registers, uniform branches)
  // CTAs of a cluster (1 or 2, chosen at launch) walk their tiles in lock step and share every weight fetch: each loads
  // 1/nct of a ring stage and multicasts it, so the L2 -> SM weight traffic per tile drops by nct.
  const uint32_t nct = cluster_nctarank(), crank = cluster_ctarank();

  float* s_bias = reinterpret_cast<float*>(smem + C::OFF_BIAS);
  float* s_wden = reinterpret_cast<float*>(smem + C::OFF_WDEN);
  float* s_wrgb = reinterpret_cast<float*>(smem + C::OFF_WRGB);
  float* s_vbias = reinterpret_cast<float*>(smem + C::OFF_VBIAS);
  float* s_part = reinterpret_cast<float*>(smem + C::OFF_PART);
  float* s_hb = reinterpret_cast<float*>(smem + C::OFF_MISC);             // [0]=b_den, [1..3]=b_rgb
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + C::OFF_MISC + 16);
  const uint32_t bar0 = sbase + C::OFF_MISC + 32;
  // barrier map (8 bytes each)
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (kMaxStages + s); };
  const uint32_t bar_inp_full = bar0 + 8 * (2 * kMaxStages), bar_inp_empty = bar0 + 8 * (2 * kMaxStages + 1);
  auto bar_acc_full = [&](int h) { return bar0 + 8 * (2 * kMaxStages + 2 + h); };
  auto bar_a_ready = [&](int kb) { return bar0 + 8 * (2 * kMaxStages + 4 + kb); };   // 0..3: one per 64-column K block of the next layer's A; 4: accumulator half 1 drained
  static_assert(32 + 8 * (2 * kMaxStages + 9) + 64 <= C::MISC_BYTES, "barrier area + BARF weights");
  float* s_barf = reinterpret_cast<float*>(smem + C::OFF_MISC + 32 + 8 * (2 * kMaxStages + 9));     // [16] (weighted IPE)

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), nct); }
    mbar_init(bar_inp_full, p.gen ? 2 : 1); mbar_init(bar_inp_empty, 1);   // generated tiles: one arrival per generator warp
    for (int h = 0; h < 2; ++h) mbar_init(bar_acc_full(h), 1);
    for (int kb = 0; kb < 5; ++kb) mbar_init(bar_a_ready(kb), 8);   // one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // biases, head weights -> shared memory
  for (int g = 0; g < p.G - 1; ++g)      // trunk layers and bottleneck (the condition layer's bias goes into s_vbias)
    for (int i = threadIdx.x; i < W; i += blockDim.x) s_bias[g * W + i] = p.params[p.sched[g].bias_off + i];
  for (int i = threadIdx.x; i < W; i += blockDim.x) s_wden[i] = p.params[p.off_wden + i];
  for (int i = threadIdx.x; i < 3 * 128; i += blockDim.x) {
    const int j = i / 128, c = i % 128;
    s_wrgb[j * 128 + c] = p.params[p.off_wrgb + c * 3 + j];
  }
  if (threadIdx.x < 4) s_hb[threadIdx.x] = threadIdx.x == 0 ? p.params[p.off_bden] : p.params[p.off_brgb + threadIdx.x - 1];
  if (p.gen && threadIdx.x >= 32 && threadIdx.x < 48) {      // mip.py:217-218: w_k = (1 - cos(clip(alpha - k, 0, 1) * pi)) / 2
    const float c = fminf(fmaxf((p.alpha_dev ? *p.alpha_dev : p.alpha) - (float)(threadIdx.x - 32), 0.f), 1.f);
    s_barf[threadIdx.x - 32] = (1.f - cosf(c * 3.14159265358979324f)) / 2.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (nct > 1) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them
  const uint32_t tmem_base = *s_tmem;

  // Every CTA of a cluster runs the same number of tile iterations (ring stages are filled and released jointly); a CTA
  // whose tile index is past the end recomputes the last tile and stores nothing.
  const int num_tiles = p.count ? min(*p.count, p.M) : p.M;
  auto more = [&](int tile) { return tile - (int)crank < num_tiles; };
  const int halves_last = p.sched[p.G - 1].n_halves;

  if (warp == 0) {
    // ===== weight producer: one ring stage per (layer, N-half [, input block]) chunk =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; more(tile); tile += gridDim.x) {
        for (int u = 0; u < p.n_uses; ++u) {
          const StageUse su = p.uses[u];
          const uint32_t bytes = (uint32_t)su.nkb * kBlockBytes;
          mbar_wait(bar_empty(stage), phase ^ 1);          // released by the MMA issuer of every CTA of the cluster
          mbar_arrive_expect_tx(bar_full(stage), bytes);
          const uint32_t dst = sbase + C::OFF_RING + stage * C::STAGE_BYTES;
          const uint8_t* src = p.packed + (size_t)su.block0 * kBlockBytes;
          if (nct == 1) {
            bulk_g2s(dst, src, bytes, bar_full(stage));
          } else if ((uint32_t)u % nct == crank) {
            // the CTAs of a pair take turns: one fetches the WHOLE stage and multicasts it to both (a 32 KB copy streams at
            // ~120 B/cycle/SM, two 16 KB halves at ~67: profiles/r01_ubench_stream.log); every CTA armed its own barrier above
            bulk_g2s_multicast(dst, src, bytes, bar_full(stage), (uint16_t)((1u << nct) - 1));
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the schedule (uniform control flow), one elected lane issues =====
    {
      IssueState st{0u, 0u, 0u};
      uint32_t inp_par = 0;
      int it = 0;
      const bool tr = DURF_TRACE && p.trace && blockIdx.x == 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      long long t_start = 0, t_begin = clock64(), tq = 0;
      const uint32_t inp_lo = ((sbase + C::OFF_INP) & 0x3FFFF) >> 4 | (1u << 16);
      const uint32_t rel_mask = nct > 1 ? (1u << nct) - 1 : 0u;
      // Every K block of MMAs waits - inside umma_kblock_conv, after its MMAs are queued - for the barriers of the NEXT
      // K block, so the thread itself never sits in a wait with an empty tensor queue behind it.  Only the first weights
      // and the per-tile start conditions are waited for up front.
      if (more(blockIdx.x)) mbar_wait(bar_full(0), 0);
      for (int tile = blockIdx.x; more(tile); tile += gridDim.x, ++it) {
        const bool more_tiles = more(tile + (int)gridDim.x);
        if (tr) tq = clock64();
        mbar_wait(bar_inp_full, inp_par); inp_par ^= 1;
        if (it > 0) {    // accumulators of the previous tile's last layer must have been drained
          mbar_wait(bar_a_ready(0), st.ar_bits & 1u); st.ar_bits ^= 1u;
          if (halves_last > 1) { mbar_wait(bar_a_ready(2), (st.ar_bits >> 2) & 1u); st.ar_bits ^= 4u; }
        }
        if (tr) t_start += clock64() - tq;
        for (int g = 0; g < p.G; ++g) {
          const LayerSched& Ls = p.sched[g];
          const bool last_layer = g + 1 == p.G;
          // the first step of the next layer: a_ready(0) unless it reads the input tile (layer 0 of the next tile, whose
          // start conditions are waited for explicitly above); its weights unless this was the last tile
          const uint32_t na0 = last_layer ? 0u : 1u;
          const uint32_t nw = (!last_layer || more_tiles) ? 1u : 0u;
          const int kbs = Ls.n_act_kb;
          const bool inp = Ls.uses_inp != 0;
          if (W == 256) {
            if (kbs == 0) issue_layer<C, W, 2, 0, true, false>(st, sbase, bar0, tmem_u, inp_lo, g, na0, nw, rel_mask);
            else if (Ls.n_halves == 2 && !inp) issue_layer<C, W, 2, 4, false, true>(st, sbase, bar0, tmem_u, inp_lo, g, na0, nw, rel_mask);
            else if (Ls.n_halves == 2) issue_layer<C, W, 2, 4, true, true>(st, sbase, bar0, tmem_u, inp_lo, g, na0, nw, rel_mask);
            else issue_layer<C, W, 1, 4, false, false>(st, sbase, bar0, tmem_u, inp_lo, g, na0, nw, rel_mask);
          } else {
            if (kbs == 0) issue_layer<C, W, 1, 0, true, false>(st, sbase, bar0, tmem_u, inp_lo, g, na0, nw, rel_mask);
            else if (!inp) issue_layer<C, W, 1, 2, false, false>(st, sbase, bar0, tmem_u, inp_lo, g, na0, nw, rel_mask);
            else issue_layer<C, W, 1, 2, true, false>(st, sbase, bar0, tmem_u, inp_lo, g, na0, nw, rel_mask);
          }
        }
      }
      if (tr && lane == 0) printf("durf mlp_tc trace: MMA thread: %d tiles, total %lld cyc; tile start (features, drained accumulators) %lld\n",
                     it, clock64() - t_begin, t_start);
    }
  } else if (warp == 2 || warp == 3) {
    if (!p.gen) {
      // ===== feature-tile loader: the next tile's features arrive while the layers after the skip layer run; the 64 threads
      // also form the next tile's view bias =====
      uint32_t par = 0;
      int it2 = 0;
      ViewWeights vw;
      view_weights_load(p, threadIdx.x - 64, vw);
      for (int tile = blockIdx.x; more(tile); tile += gridDim.x, ++it2) {
        const int tcl = min(tile, num_tiles - 1);
        // every thread waits for the input tile to be free: that also says the tile before the previous one is past its
        // layer 1, i.e. its half of the view-bias double buffer has been read
        mbar_wait(bar_inp_empty, par ^ 1);
        view_bias(p, vw, p.ray_index ? p.ray_index[tcl] : tcl, threadIdx.x - 64, s_vbias + 128 + (it2 & 1) * 128);
        asm volatile("bar.sync 3, 64;" ::: "memory");        // the row is complete before the tile is announced
        if (threadIdx.x == 64) {
          mbar_arrive_expect_tx(bar_inp_full, kInpBytes);
          bulk_g2s(sbase + C::OFF_INP, p.feat + (size_t)tcl * kInpBytes, kInpBytes, bar_inp_full);
        }
        par ^= 1;
      }
    } else {
      // ===== feature-tile GENERATOR (N1): 64 threads, two samples each; the ray-march of raymarch.cu's bf16 path, written
      // straight into the SWIZZLE_128B A-operand image in shared memory.  It runs while the layers after the skip layer of the
      // previous tile execute (the tile is ~30 k cycles of MMAs, a row costs ~400 instructions). =====
      const int gt = threadIdx.x - 64;                       // 0..63
      const bool weighted = (p.rm_flags & DURF_RM_WEIGHTED) != 0;
      const bool sample = (p.rm_flags & DURF_RM_SAMPLE) != 0;
      uint32_t par = 0;
      int it2 = 0;
      ViewWeights vw;
      view_weights_load(p, gt, vw);
      for (int tile = blockIdx.x; more(tile); tile += gridDim.x, ++it2) {
        const bool valid = tile < num_tiles;
        const int tcl = min(tile, num_tiles - 1);
        const int ray = p.ray_index ? p.ray_index[tcl] : tcl;
        view_bias(p, vw, ray, gt, s_vbias + 128 + (it2 & 1) * 128);   // announced with the tile (inp_full below)
        const float o[3] = {p.g_origins[3 * ray], p.g_origins[3 * ray + 1], p.g_origins[3 * ray + 2]};
        const float d[3] = {p.g_dirs[3 * ray], p.g_dirs[3 * ray + 1], p.g_dirs[3 * ray + 2]};
        const float radius = p.g_radii[ray];
        const bool has_mult = p.g_ray_mult != nullptr;
        const float mult = !has_mult ? 1.f : ((p.rm_flags & DURF_RM_MULT_IS_NHIT) ? 1.f - p.g_ray_mult[ray] : p.g_ray_mult[ray]);
        float tf[2][2];                                      // fenceposts (r, r+1) of this thread's rows gt and gt + 64
        if (sample) {
          const float nr = p.g_near[ray], fr = p.g_far[ray];
          const float* tr_row = (p.rm_flags & DURF_RM_RANDOMIZED) ? p.g_t_rand + (size_t)ray * (kTileM + 1) : nullptr;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int r = gt + 64 * k;
            tf[k][0] = sample_fencepost(nr, fr, r, kTileM, tr_row);
            tf[k][1] = sample_fencepost(nr, fr, r + 1, kTileM, tr_row);
            if (valid && !(p.rm_flags & DURF_RM_NO_TVALS_OUT)) {
              p.g_t_vals[(size_t)ray * (kTileM + 1) + r] = tf[k][0];
              if (r == kTileM - 1) p.g_t_vals[(size_t)ray * (kTileM + 1) + kTileM] = tf[k][1];
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int r = gt + 64 * k;
            tf[k][0] = p.g_t_vals[(size_t)ray * (kTileM + 1) + r];
            tf[k][1] = p.g_t_vals[(size_t)ray * (kTileM + 1) + r + 1];
          }
        }
        if (p.feat_out) {     // the previous tile's image must have left shared memory before it is overwritten
          if (gt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 3, 64;" ::: "memory");
        }
        mbar_wait(bar_inp_empty, par ^ 1);                   // every MMA reading the previous tile's image has retired
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const Gauss g = sample_gaussian(p.rm_flags, o, d, radius, mult, has_mult, tf[k][0], tf[k][1]);
          if (weighted) encode_row_bf16_smem<true>(g, p.min_deg, s_barf, sbase + C::OFF_INP, gt + 64 * k);
          else encode_row_bf16_smem<false>(g, p.min_deg, s_barf, sbase + C::OFF_INP, gt + 64 * k);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to tcgen05.mma / bulk copies
        if (p.feat_out) {
          asm volatile("bar.sync 3, 64;" ::: "memory");
          if (gt == 0) {
            if (valid)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.feat_out + (size_t)tile * kInpBytes),
                           "r"(sbase + C::OFF_INP), "n"(kInpBytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_inp_full);
        par ^= 1;
      }
      if (p.feat_out && gt == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = accumulator row = sample; warps q and q+4 share TMEM lane quarter q =====
    const int q = warp & 3;
    const int ch = (warp - 4) >> 2;               // which 64-column slice of a half this warp owns
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t af_par = 0;                          // bit h: parity of acc_full(h)
    constexpr int NG = C::CPW / 32;               // 32-column groups per warp per half
    // The first epilogue thread frees the input tile once the MMAs of the last layer reading it have retired.  (It must be a
    // thread whose progress the MMA issuer depends on: a passive observer of acc_full could fall two completions behind and
    // miss a phase.)  Weight ring stages are released by the MMA issuer's own commits.
    const bool releaser = threadIdx.x == 128;
    uint32_t stg_buf = 0;                         // SAVE: which of this warp's two 4 KB staging pieces the next epilogue fills
    const bool tr = DURF_TRACE && p.trace && blockIdx.x == 0 && threadIdx.x == 128;
    long long e_acc = 0, e_ld = 0, e_math = 0, e_st = 0, e_begin = clock64(), eq = 0, e_m0 = 0, e_ld1 = 0, e_m1 = 0;
    const bool trs = DURF_TRACE_DETAIL && tr;
    long long sv_wg = 0, sv_fill = 0, sv_fence = 0, sv_issue = 0, sq = 0;
    // Per-tile housekeeping (raw outputs of the PREVIOUS tile, view bias of this one) is deferred until after layer 0's
    // epilogues, when the tensor core has a whole layer of MMAs queued: the tile boundary costs the issuer nothing.
    float den_prev = 0.f, rgb_prev[3] = {0.f, 0.f, 0.f};
    int ray_prev = -1, tile_prev = 0;
    auto flush_prev = [&]() {      // ch == 0 threads: combine the two column slices of every row, write raw outputs
      const size_t o = (size_t)(p.accumulate == 2 ? tile_prev : ray_prev) * kTileM + row;   // 2: compact rows (durf_mlp_merge_raw)
      const float dv = den_prev + s_part[row] + s_hb[0];
      const float r0 = rgb_prev[0] + s_part[128 + row] + s_hb[1];
      const float r1 = rgb_prev[1] + s_part[256 + row] + s_hb[2];
      const float r2 = rgb_prev[2] + s_part[384 + row] + s_hb[3];
      if (p.accumulate == 1) {
        p.raw_density[o] += dv;
        p.raw_rgb[o * 3 + 0] += r0; p.raw_rgb[o * 3 + 1] += r1; p.raw_rgb[o * 3 + 2] += r2;
      } else {
        p.raw_density[o] = dv;
        p.raw_rgb[o * 3 + 0] = r0; p.raw_rgb[o * 3 + 1] = r1; p.raw_rgb[o * 3 + 2] = r2;
      }
    };
    // per-ray bias of the condition layer (b + W_view^T enc(viewdir), obbpose_model.py:343-350): fetched a tile ahead
    int it_e = 0;
    for (int tile = blockIdx.x; more(tile); tile += gridDim.x, ++it_e) {
      const bool valid = tile < num_tiles;
      const int ray = !valid ? -1 : (p.ray_index ? p.ray_index[tile] : tile);
      float den = 0.f;
      float rgb[3] = {0.f, 0.f, 0.f};
      for (int g = 0; g < p.G; ++g) {
        if (g == 1) {
          asm volatile("bar.sync 1, 256;" ::: "memory");   // s_part of the previous tile is complete; its s_vbias readers are done
          if (ch == 0) {
            if (ray_prev >= 0) flush_prev();
            // this tile's view bias: formed by warps 2-3 before they announced the tile's input (inp_full -> MMAs of layer 0 ->
            // acc_full, which this thread has waited on)
            s_vbias[row] = s_vbias[128 + (it_e & 1) * 128 + row];
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        // the layer's schedule entry is fetched before the wait: nothing between the accumulator barrier and the first
        // TMEM load may depend on a constant-bank round trip (this gap is on the layer-to-layer critical path)
        const LayerSched& L = p.sched[g];
        const int kind = L.kind, n_halves = L.n_halves;
        const bool rel_inp = L.last_inp_use != 0;
        const uint32_t o_buf = t_lane + C::ACT_COL + ((g + 1) & 1) * (W / 2);
        for (int h = 0; h < n_halves; ++h) {
          const int col0 = h * 128 + ch * C::CPW;
          if (tr) eq = clock64();
          mbar_wait(bar_acc_full(h), (af_par >> h) & 1u); af_par ^= 1u << h;
          if (tr) { e_acc += clock64() - eq; eq = clock64(); }
          tc_fence_after();
          auto release = [&]() {      // the input tile is free once the last layer reading it has completed
            if (releaser && rel_inp && h == n_halves - 1) mbar_arrive(bar_inp_empty);
          };
          // training: one warp's 32 rows x 64 columns of a saved activation (and their 1-bit ReLU mask words)
          auto save_piece = [&](const uint32_t (&pk)[NG][16], bool with_mask) {
            if (trs) sq = clock64();
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the piece filled two epilogues ago was read
            __syncwarp();
            if (trs) { sv_wg += clock64() - sq; sq = clock64(); }
            const uint32_t sdst = sbase + C::OFF_STG + ((stg_buf * 8 + (warp - 4)) << 12) + (lane >> 3) * 1024 + (lane & 7) * 128;
#pragma unroll
            for (int i = 0; i < NG; ++i) {
              if (with_mask) {
                uint32_t mw = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) {     // HSET2: 0xFFFF per half that is > 0, then one LOP3 picks the word's two bits
                  uint32_t gt;
                  asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(gt) : "r"(pk[i][k]), "r"(0u));
                  mw |= gt & (0x80008000u >> k);
                }
                if (valid) p.masks[(((size_t)tile * (p.depth + 1) + (kind == 3 ? p.depth : g)) * (W / 32) + ((col0 + i * 32) >> 5)) * 128 + row] = mw;
              }
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdst + (((uint32_t)(i * 4 + q4) ^ (lane & 7)) << 4)),
                             "r"(pk[i][4 * q4]), "r"(pk[i][4 * q4 + 1]), "r"(pk[i][4 * q4 + 2]), "r"(pk[i][4 * q4 + 3]) : "memory");
            }
            __syncwarp();
            if (trs) { sv_fill += clock64() - sq; sq = clock64(); }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (trs) { sv_fence += clock64() - sq; sq = clock64(); }
            if (lane == 0) {
              // record of layer g: [sample half][64-column block][64 rows x 128 B], so that the weight-gradient kernel fetches the
              // 64 samples of ALL the layer's blocks with one contiguous bulk copy; this warp's 32 rows are 4 KB of it
              const int nb = 2 * n_halves;            // 64-column blocks of this layer's output
              uint8_t* gdst = p.saved + ((size_t)tile * p.saved_blocks_per_tile + g * C::KB) * kBlockBytes +
                              (size_t)(q >> 1) * nb * 8192 + (col0 >> 6) * 8192 + (q & 1) * 4096;
              if (valid)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(gdst),
                             "r"(sbase + C::OFF_STG + ((stg_buf * 8 + (warp - 4)) << 12)) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (trs) sv_issue += clock64() - sq;
            stg_buf ^= 1;
          };
          uint32_t v[NG][32];
          if (kind != 3) {
            // software pipeline over two 32-column groups: the TMEM load of group 1 and the bias fetches run
            // under the arithmetic of group 0 (TMEM reads are the epilogue's floor: 32 B/cycle per lane quarter).
            // Inference: group i of every warp lies in K block 2h+i of the next layer's A operand, so the eight warps
            // finish block 2h first and the issuer can start on it while block 2h+1 is still being packed.  Training keeps
            // each warp on 64 adjacent columns (its 4 KB piece of the saved activation image).
            auto gcol = [&](int i) { return SAVE ? col0 + i * 32 : h * 128 + i * 64 + ch * 32; };
            tmem_ld32_issue(t_lane + C::ACC_COL + gcol(0), v[0]);
            release();
            uint32_t pk[NG][16];
#pragma unroll
            for (int i = 0; i < NG; ++i) {
              const int cg = gcol(i);
              const uint32_t sb = sbase + C::OFF_BIAS + (g * W + cg) * 4;
              float4 b4[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) b4[j] = lds128(sb + 16 * j);
              long long tq1 = 0;
              if (tr && i == 1) tq1 = clock64();
              tmem_ld_wait();
              tmem_ld_pin(v[i]);
              if (tr && i == 1) e_ld1 += clock64() - tq1;
              if (i + 1 < NG) tmem_ld32_issue(t_lane + C::ACC_COL + gcol(i + 1), v[i + 1]);
              if (i == 0 && h == 1 && L.drain_sig) {
                // accumulator half 1 is in registers: the next layer's [h1: k0] (issued before anything that depends on this
                // epilogue's output) may overwrite it.  Off the critical path: every consumer of this half's output has slack.
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_a_ready(4));
              }
              if (tr && i == 0) { e_ld += clock64() - eq; eq = clock64(); }
              if (tr) tq1 = clock64();
              const uint32_t wden_addr = sbase + C::OFF_WDEN + cg * 4;
              if (kind == 0) epi_pack<0>(v[i], b4, wden_addr, den, pk[i]);
              else if (kind == 1) epi_pack<1>(v[i], b4, wden_addr, den, pk[i]);
              else epi_pack<2>(v[i], b4, wden_addr, den, pk[i]);
              tmem_st16(o_buf + cg / 2, pk[i]);
              if constexpr (!SAVE) {
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_a_ready(2 * h + i));   // K block 2h+i of the next layer's A operand is in TMEM
              }
              if (tr) { if (i == 0) e_m0 += clock64() - tq1; else e_m1 += clock64() - tq1; }
            }
            if (tr) { e_math += clock64() - eq; eq = clock64(); }
            if constexpr (SAVE) {
              // hand the half over first: nothing of the bookkeeping below may sit between "accumulator complete" and "A operand
              // of the next layer ready"
              tmem_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                mbar_arrive(bar_a_ready(2 * h));   // half h (K blocks 2h, 2h+1) of the next layer's A operand is in TMEM
                mbar_arrive(bar_a_ready(2 * h + 1));
              }
              if (tr) { e_st += clock64() - eq; eq = clock64(); }
              // training, off the critical path: (1) the ReLU mask dgrad needs, 1 bit per element (word k of a group holds
              // elements 2k, 2k+1 in its halves; a non-negative bf16 half is non-zero iff bit 15 of half + 0x7FFF is set):
              // element 2k -> bit 15-k, element 2k+1 -> bit 31-k; dgrad reads 4 bytes per thread and group instead of a
              // 64-byte activation row.  (2) the activation itself (A operand of wgrad): the warp's 32 rows x 64 columns are a
              // contiguous 4 KB piece of the global block image, staged in shared memory in image order and handed to the
              // bulk-copy engine, so the store to HBM never blocks the epilogue.
              save_piece(pk, kind != 2);
              if (tr) e_math += clock64() - eq;
            }
          } else {
            // condition layer: + per-ray bias, ReLU, partial rgb head over this warp's columns
#pragma unroll
            for (int i = 0; i < NG; ++i) tmem_ld32_issue(t_lane + C::ACC_COL + col0 + i * 32, v[i]);
            release();
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < NG; ++i) tmem_ld_pin(v[i]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a_ready(2 * h));   // accumulators drained: the next tile's first layer may overwrite them
            const uint32_t svb = sbase + C::OFF_VBIAS + col0 * 4, swr = sbase + C::OFF_WRGB + col0 * 4;
            uint32_t pkc[NG][16];
#pragma unroll
            for (int i = 0; i < NG; ++i)
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = i * 32 + 4 * j;
                const float4 vb = lds128_volatile(svb + c * 4);   // rewritten per tile
                const float4 w0 = lds128(swr + c * 4), w1 = lds128(swr + (128 + c) * 4), w2 = lds128(swr + (256 + c) * 4);
                const float a0 = fmaxf(__uint_as_float(v[i][4 * j]) + vb.x, 0.f), a1 = fmaxf(__uint_as_float(v[i][4 * j + 1]) + vb.y, 0.f);
                const float a2 = fmaxf(__uint_as_float(v[i][4 * j + 2]) + vb.z, 0.f), a3 = fmaxf(__uint_as_float(v[i][4 * j + 3]) + vb.w, 0.f);
                rgb[0] = fmaf(a0, w0.x, rgb[0]); rgb[0] = fmaf(a1, w0.y, rgb[0]); rgb[0] = fmaf(a2, w0.z, rgb[0]); rgb[0] = fmaf(a3, w0.w, rgb[0]);
                rgb[1] = fmaf(a0, w1.x, rgb[1]); rgb[1] = fmaf(a1, w1.y, rgb[1]); rgb[1] = fmaf(a2, w1.z, rgb[1]); rgb[1] = fmaf(a3, w1.w, rgb[1]);
                rgb[2] = fmaf(a0, w2.x, rgb[2]); rgb[2] = fmaf(a1, w2.y, rgb[2]); rgb[2] = fmaf(a2, w2.z, rgb[2]); rgb[2] = fmaf(a3, w2.w, rgb[2]);
                if (SAVE) { pkc[i][2 * j] = cvt_bf16x2(a0, a1); pkc[i][2 * j + 1] = cvt_bf16x2(a2, a3); }
              }
            if constexpr (SAVE) save_piece(pkc, true);     // the condition layer's activation (wgrad operand) and its 1-bit mask (mask layer `depth`)
          }
        }
      }
      if (tr && !more(tile + (int)gridDim.x))
        printf("durf mlp_tc trace: epilogue thread: total %lld cyc; waiting acc_full %lld, tmem ld %lld, math+st issue %lld (g0 %lld, ld1 wait %lld, g1 %lld), st wait %lld\n",
               clock64() - e_begin, e_acc, e_ld, e_math, e_m0, e_ld1, e_m1, e_st);
      if (SAVE && trs && !more(tile + (int)gridDim.x))
        printf("durf mlp_tc trace: save_piece: wait_group %lld, masks + st.shared %lld, fence.proxy.async %lld, bulk-store issue %lld\n", sv_wg, sv_fill, sv_fence, sv_issue);
      // stash this tile's head results: they are combined and written while the next tile's layer 1 runs
      if (ch == 1) {
        s_part[row] = den; s_part[128 + row] = rgb[0]; s_part[256 + row] = rgb[1]; s_part[384 + row] = rgb[2];
      } else {
        den_prev = den; rgb_prev[0] = rgb[0]; rgb_prev[1] = rgb[1]; rgb_prev[2] = rgb[2];
      }
      ray_prev = ray; tile_prev = tile;
    }
    if (ray_prev >= 0) {             // the last tile's outputs
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ch == 0) flush_prev();
    }
  }
  if (SAVE && warp >= 4 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // staged activations are in HBM
  tc_fence_before();
  __syncthreads();
  if (nct > 1) cluster_sync_all();      // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
  }
}

// fp32 parameter blob -> bf16 weight images: for every GEMM layer, N half and K block (the order the kernels consume them)
// one 128 x 64 K-major SWIZZLE_128B block of 16 KB.  One thread per 16-byte piece; ONE launch covers the forward and the
// transposed images of every network handed to durf_mlp_pack_weights_multi.
struct PackMultiParams {
  int n_blocks;
  int pad;
  PackBlock b[kPackMaxBlocks];
};
__global__ void pack_blocks_kernel(const __grid_constant__ PackMultiParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)p.n_blocks * 1024) return;
  const int blk = (int)(i / 1024), r = (int)(i % 1024) / 8, c = (int)(i % 8);
  const PackBlock& B = p.b[blk];
  const float* src = B.src + (size_t)r * B.sr;
  const bool row_ok = r < B.r_avail;
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = c * 8 + 2 * e;
    w[e] = pack_bf16x2(row_ok && k < B.k_avail ? src[(size_t)k * B.sk] : 0.f, row_ok && k + 1 < B.k_avail ? src[(size_t)(k + 1) * B.sk] : 0.f);
  }
  *reinterpret_cast<uint4*>(B.dst + sw128_offset(r, c)) = make_uint4(w[0], w[1], w[2], w[3]);
}

int pack_blocks_launch(cudaStream_t st, const PackBlock* blocks, int n) {
  for (int first = 0; first < n; first += kPackMaxBlocks) {
    PackMultiParams pp;
    pp.n_blocks = n - first < kPackMaxBlocks ? n - first : kPackMaxBlocks; pp.pad = 0;
    for (int i = 0; i < pp.n_blocks; ++i) pp.b[i] = blocks[first + i];
    pack_blocks_kernel<<<ceil_div((int64_t)pp.n_blocks * 1024, 256), 256, 0, st>>>(pp);
    DURF_CHECK_LAUNCH("durf_mlp_pack_weights");
  }
  return DURF_OK;
}

static bool tc_supported(const DurfMlpTopology& t) {
  return (t.width == 256 || t.width == 128) && t.cond_width == 128 && t.in_dim <= 64 && t.depth >= 2 && t.depth <= 9 &&
         t.cond_dim <= 64 && !(((t.depth - 1) % t.skip == 0) && t.depth - 1 > 0);
}

// Builds the per-layer schedule and the flat list of ring-stage uses shared by the pack kernel and the MLP kernel.
// Returns the number of 16 KB weight blocks.
static int build_sched(const DurfMlpTopology& t, TcParams& P, int skb = 2) {
  MlpLayout L(t);
  const int KB = t.width / 64;
  int last_inp = 0;
  P.G = t.depth + 2;
  P.depth = t.depth;
  P.cond_dim = t.cond_dim;
  for (int g = 0; g < P.G; ++g) {
    LayerSched& s = P.sched[g];
    if (g < t.depth) {
      const bool skip_in = (g >= 1) && ((g - 1) % t.skip == 0) && (g - 1 > 0);
      s.n_halves = t.width / 128;
      s.n_act_kb = (g == 0) ? 0 : KB;
      s.uses_inp = (g == 0 || skip_in) ? 1 : 0;
      s.kind = (g == t.depth - 1) ? 1 : 0;
      s.bias_off = (int)L.b_off[g];
      if (s.uses_inp) last_inp = g;
    } else if (g == t.depth) {       // bottleneck = Dense_{depth+1}
      s.n_halves = t.width / 128; s.n_act_kb = KB; s.uses_inp = 0; s.kind = 2; s.bias_off = (int)L.b_off[t.depth + 1];
    } else {                         // condition layer = Dense_{depth+2}
      s.n_halves = t.cond_width / 128; s.n_act_kb = KB; s.uses_inp = 0; s.kind = 3; s.bias_off = (int)L.b_off[t.depth + 2];
    }
    s.last_inp_use = 0;
  }
  P.sched[last_inp].last_inp_use = 1;
  // Issue order (see StepSched).  Every group of <= 2 K blocks of one N-half is one ring-stage use; the packed image stores
  // the blocks in exactly this order.
  int blocks = 0, ns = 0, nu = 0;
  for (int g = 0; g < P.G; ++g)
    P.sched[g].drain_sig = (g + 1 < P.G && P.sched[g].n_halves == 2 && P.sched[g + 1].n_halves == 2 && P.sched[g + 1].n_act_kb == 4) ? 1 : 0;
  for (int g = 0; g < P.G; ++g) {
    const LayerSched& s = P.sched[g];
    const bool drained = g >= 1 && P.sched[g - 1].drain_sig;
    const LayerSteps LS = layer_steps(s.n_halves, s.n_act_kb, s.uses_inp != 0, drained);
    for (int i = 0; i < LS.n; ++i) {
      StepSched& sp = P.steps[ns++];
      sp.g = (int8_t)g; sp.nh = LS.nh[i]; sp.kb = LS.kb[i]; sp.inp = LS.inp[i]; sp.acc0 = LS.acc0[i]; sp.commit = LS.commit[i];
      sp.stage_first = sp.stage_last = sp.slot = 0; sp.wait_a = LS.wait[i]; sp.pad[0] = sp.pad[1] = 0;
    }
    blocks += LS.n;
  }
  // ring-stage uses: up to `skb` consecutive K blocks of the issue order, never across a layer boundary
  for (int c = 0; c < ns;) {
    int n = 1;
    while (n < skb && c + n < ns && P.steps[c + n].g == P.steps[c].g) ++n;
    P.uses[nu].block0 = (int16_t)c; P.uses[nu].nkb = (int16_t)n; ++nu;
    for (int i = 0; i < n; ++i) {
      P.steps[c + i].slot = (int8_t)i;
      P.steps[c + i].stage_first = i == 0 ? 1 : 0;
      P.steps[c + i].stage_last = i == n - 1 ? 1 : 0;
    }
    c += n;
  }
  for (int c = 0; c < ns; ++c) P.step_w[c] = step_word(P.steps[c]);
  P.n_steps = ns; P.n_uses = nu;
  P.off_wden = (int)L.w_off[t.depth]; P.off_bden = (int)L.b_off[t.depth];
  P.off_wrgb = (int)L.w_off[t.depth + 3]; P.off_brgb = (int)L.b_off[t.depth + 3];
  P.off_wview = (int)L.w_off[t.depth + 2] + t.width * t.cond_width;
  return blocks;
}

int64_t mlp_tc_packed_bytes(const DurfMlpTopology& t) {
  if (!tc_supported(t)) return 0;
  TcParams P;
  return (int64_t)build_sched(t, P) * kBlockBytes;
}

// Appends the block descriptions of the forward image to `out`; returns their number (or a negative error code).
int mlp_tc_pack_blocks(const DurfMlpTopology& t, const float* params, void* packed, std::vector<PackBlock>& out) {
  DURF_REQUIRE(tc_supported(t), DURF_E_UNSUPPORTED,
               "durf_mlp_pack_weights: tensor-core path needs width 128/256, cond_width 128, in_dim <= 64, depth <= 9");
  TcParams P;
  const int blocks = build_sched(t, P);
  DURF_REQUIRE(P.n_steps <= kMaxSteps && P.n_uses <= kMaxStageUses, DURF_E_UNSUPPORTED,
               "durf_mlp_pack_weights: too many weight blocks (%d)", blocks);
  MlpLayout L(t);
  for (int c = 0; c < P.n_steps; ++c) {          // block c of the image = K block of step c (same order as the ring-stage uses)
    const StepSched& sp = P.steps[c];
    const int g = sp.g;
    const int layer = (g < t.depth) ? g : (g == t.depth ? t.depth + 1 : t.depth + 2);
    const int n_out = L.out_dim[layer];
    const int n_first = sp.nh * 128;
    // K rows of the block: a trunk K block, or the input features (for the skip layer they follow the trunk rows)
    const int k_first = !sp.inp ? sp.kb * 64 : (g == 0 ? 0 : t.width);
    PackBlock b;
    b.src = params + L.w_off[layer] + (size_t)k_first * n_out + n_first;
    b.dst = (uint8_t*)packed + (size_t)c * kBlockBytes;
    b.sr = 1; b.sk = n_out;                      // block[r][k] = kernel[k_first + k][n_first + r]
    b.r_avail = n_out - n_first < 128 ? n_out - n_first : 128;
    b.k_avail = !sp.inp ? 64 : t.in_dim;
    out.push_back(b);
  }
  return blocks;
}

bool mlp_tc_bwd_supported(const DurfMlpTopology& t);
int mlp_tc_saved_blocks(const DurfMlpTopology& t);

size_t mlp_tc_workspace_bytes(const DurfMlpTopology&, int64_t) { return 0; }      // the tensor-core forward needs no workspace

int mlp_tc_forward(cudaStream_t st, const DurfMlpArgs& a) {
  const DurfMlpTopology& t = a.topo;
  DURF_REQUIRE(tc_supported(t), DURF_E_UNSUPPORTED,
               "durf_mlp_fwd(bf16): tensor-core path needs width 128/256, cond_width 128, in_dim <= 64, depth <= 9");
  DURF_REQUIRE(a.N == kTileM, DURF_E_UNSUPPORTED, "durf_mlp_fwd(bf16): needs 128 samples per ray (got %d)", a.N);
  const DurfRaymarchArgs* rm = a.fused_raymarch;
  DURF_REQUIRE(a.packed && a.params && (a.features || rm) && a.cond, DURF_E_INVALID, "durf_mlp_fwd(bf16): null buffer");
  if (rm) {
    const bool weighted = (rm->flags & DURF_RM_WEIGHTED) != 0;
    DURF_REQUIRE(rm->N == kTileM && rm->max_deg - rm->min_deg == 10 && t.in_dim == 60 + (weighted ? 3 : 0), DURF_E_UNSUPPORTED,
                 "durf_mlp_fwd(bf16): fused ray-march needs N = 128, 10 degrees and in_dim = 60 (IPE) / 63 (weighted IPE)");
    DURF_REQUIRE(rm->origins && rm->dirs && rm->radii, DURF_E_INVALID, "durf_mlp_fwd(bf16): fused ray-march: null ray buffer");
    DURF_REQUIRE(rm->t_vals || (rm->flags & (DURF_RM_SAMPLE | DURF_RM_NO_TVALS_OUT)) == (DURF_RM_SAMPLE | DURF_RM_NO_TVALS_OUT), DURF_E_INVALID,
                 "durf_mlp_fwd(bf16): fused ray-march: t_vals missing");
    DURF_REQUIRE(!(rm->flags & DURF_RM_SAMPLE) || (rm->near && rm->far), DURF_E_INVALID, "durf_mlp_fwd(bf16): fused ray-march: near/far missing");
    DURF_REQUIRE(!(rm->flags & DURF_RM_RANDOMIZED) || rm->t_rand, DURF_E_INVALID, "durf_mlp_fwd(bf16): fused ray-march: t_rand missing");
  }
  DURF_REQUIRE(a.saved == nullptr || mlp_tc_bwd_supported(t), DURF_E_UNSUPPORTED,
               "durf_mlp_fwd(bf16): activation saving needs a topology with a tensor-core backward");
  TcParams P;
  build_sched(t, P, (t.width == 256 && a.saved == nullptr) ? 4 : 2);      // = TcCfg<W, SAVE>::SKB
  // `saved` may be a buffer for more ray-levels than this call's (several levels, one backward): records first, then masks
  const int total = a.saved_total_tiles > 0 ? a.saved_total_tiles : a.M, off = a.saved_total_tiles > 0 ? a.saved_tile_offset : 0;
  DURF_REQUIRE(off >= 0 && off + a.M <= total, DURF_E_INVALID, "durf_mlp_fwd(bf16): saved_tile_offset %d + M %d > saved_total_tiles %d",
               off, a.M, total);
  P.saved_blocks_per_tile = mlp_tc_saved_blocks(t);
  P.saved = a.saved ? (uint8_t*)a.saved + (size_t)off * P.saved_blocks_per_tile * kBlockBytes : nullptr;
  P.masks = a.saved ? reinterpret_cast<uint32_t*>((uint8_t*)a.saved + (size_t)total * P.saved_blocks_per_tile * kBlockBytes) +
                          (size_t)off * (t.depth + 1) * (t.width / 32) * 128
                    : nullptr;
  {
    MlpLayout L(t);
    P.off_bcond = (int)L.b_off[t.depth + 2];      // the per-tile view bias is formed inside the kernel (view_bias): no workspace
  }
  P.gen = rm ? 1 : 0;
  P.feat_out = nullptr;
  if (rm) {
    P.rm_flags = rm->flags; P.min_deg = rm->min_deg; P.alpha = rm->alpha; P.alpha_dev = rm->alpha_dev;
    P.g_origins = rm->origins; P.g_dirs = rm->dirs; P.g_radii = rm->radii; P.g_near = rm->near; P.g_far = rm->far;
    P.g_t_rand = rm->t_rand; P.g_ray_mult = rm->ray_mult; P.g_t_vals = rm->t_vals;
    P.feat_out = (uint8_t*)a.features;            // NULL: the tiles exist only in shared memory
  }
  P.feat = (const uint8_t*)a.features; P.cond = a.cond; P.params = a.params; P.packed = (const uint8_t*)a.packed;
  P.ray_index = a.ray_index; P.count = a.count; P.M = a.M; P.accumulate = a.accumulate;
  P.raw_rgb = a.raw_rgb; P.raw_density = a.raw_density;
  static const int trace_env = getenv("DURF_TC_TRACE") ? atoi(getenv("DURF_TC_TRACE")) : 0;
  P.trace = trace_env;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // CTA pairs (one cluster per TPC) share each weight fetch by multicast; DURF_TC_CLUSTER=1 launches unpaired CTAs.
  static const int cluster_env = getenv("DURF_TC_CLUSTER") ? atoi(getenv("DURF_TC_CLUSTER")) : 2;
  const int nct = (cluster_env == 1 || sms < 2) ? 1 : 2;
  const int grid_cap = sms / nct * nct;
  const int grid_want = (a.M + nct - 1) / nct * nct;
  const int grid = grid_want < grid_cap ? grid_want : grid_cap;
  cudaError_t e = cudaSuccess;
  auto launch = [&](auto kernel, int smem) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nct; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kernel, P);
  };
  if (t.width == 256) {
    if (P.saved) launch(mlp_tc_fwd_kernel<256, true>, TcCfg<256, true>::SMEM_BYTES);
    else launch(mlp_tc_fwd_kernel<256, false>, TcCfg<256, false>::SMEM_BYTES);
  } else {
    if (P.saved) launch(mlp_tc_fwd_kernel<128, true>, TcCfg<128, true>::SMEM_BYTES);
    else launch(mlp_tc_fwd_kernel<128, false>, TcCfg<128, false>::SMEM_BYTES);
  }
  DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_fwd(bf16): launch: %s", cudaGetErrorString(e));
  DURF_CHECK_LAUNCH("durf_mlp_fwd(bf16)");
  return DURF_OK;
}

}  // namespace durf
