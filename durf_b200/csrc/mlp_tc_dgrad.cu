// K2 backward, data gradients as a fused tcgen05 GEMM chain (sm_100a): the reverse of mlp_tc.cu.
// Per tile (the 128 samples of one ray-level), from dL/d raw_rgb [128,3] and dL/d raw_density [128]:
//   prologue : dZ_cond = (d_rgb W_rgb^T) * [cond_act > 0]                              (CUDA cores, N = 3 head)
//   stage 0  : dBott   = dZ_cond W_cond[:width]^T                                      (K = 128)
//   stage 1  : dZ_last = (dBott W_bott^T + d_density (x) w_density) * [a_last > 0]
//   stage 2+ : dZ_{g-1} = (dZ_g W_g[:width]^T) * [a_{g-1} > 0]      g = depth-1 .. 1
// Same machinery as the forward: accumulators AND the bf16 dZ operands live in tensor memory (tcgen05.mma .ts), every
// stage is issued as N-halves so the masking epilogue of one half overlaps the MMAs of the other, the transposed weight
// image streams through a ring of 32 KB stages.  Every dZ (and dBott, dZ_cond) is also written to HBM as block images
// (per-warp 4 KB pieces: shared-memory staging + cp.async.bulk stores): they are the B operands of the weight-gradient
// kernel (mlp_tc_wgrad.cu).  Every ReLU mask is a 1-bit word written by the forward pass (the saved activations themselves
// are not read here); the next tile's upstream gradients and mask words are prefetched under the last stage.  With
// `tile_done` counters (durf_mlp_bwd_data) warp 3 publishes every completed dZ piece for a weight-gradient kernel running
// concurrently on other SMs.  For width 128 (BoxMLP) an optional last stage forms the input gradient
//   dX = dZ_0 W_0^T + dZ_skip W_skip[width:]^T   (64 columns, fp32 rows)
// that the box-pose path needs; dZ_skip waits in a third TMEM buffer.  The background branch needs no input gradient
// (its samples depend on no parameter).
#include <stdlib.h>

#include <vector>

#include "tc_common.cuh"
#include "mlp_topology.h"

namespace durf {

constexpr int kDgMaxStages = 12;
constexpr int kMailSlots = 32;        // per epilogue warp: dZ pieces whose completion is known but not yet published

struct DgStage {
  int n_halves;     // output columns / 128 (1 for the 64-column input-gradient stage)
  int n_kb;         // 64-wide K blocks of the incoming dZ
  int kind;         // 0: linear (dBott); 1: + density term, mask; 2: mask; 3: input gradient (fp32 rows to d_features)
  int mask_slot;    // block offset (inside a saved tile record) of the activation whose sign masks this stage's output
  int out_slot;     // block offset (inside a dz tile record) where this stage's output is stored
  int block0;       // first 16 KB block of this stage inside the transposed weight image
  int a_sel;        // TMEM buffer holding the A operand: 0 / 1 = the alternating dZ buffers, 2 = the kept dZ of the skip layer
  int o_sel;        // buffer the epilogue writes the next A operand to (-1: none)
  int first_part;   // first MMA overwrites the accumulator
  int last_part;    // commit acc_full after this entry (an epilogue follows); entries with 0 only accumulate
  int parts;        // ring stages the epilogue releases (entries since the previous epilogue)
  int to_skip;      // the epilogue also keeps its dZ in the skip buffer (it multiplies the input rows of the skip layer later)
  int ncols;        // MMA N: 128, or 64 for the input gradient
};

struct DgParams {
  const uint8_t* saved;      // forward activations (tile records of saved_blocks blocks)
  const uint32_t* masks;     // 1-bit ReLU masks written by the forward: [tile][trunk layer][32-column group][row]
  int depth;
  uint8_t* dz;               // output: dz tile records (same record shape)
  const uint8_t* packed_t;   // transposed weight image
  const float* params;       // fp32 blob (density / rgb head weights)
  const float* d_raw_rgb;    // [B,128,3]
  const float* d_raw_density;// [B,128]
  const int32_t* ray_index;
  const int32_t* count;
  int M, saved_blocks;
  int n_stages;
  int cond_slot;             // slot of the condition layer (its activation in `saved`, dZ_cond in `dz`)
  int off_wden, off_wrgb;
  float* d_features;         // [opt] [M*128, in_dim] fp32 gradient w.r.t. the input features (object MLPs, box-pose path)
  int in_dim;
  int trace;
  // [opt] overlap with the weight-gradient kernel running on other SMs: tile_done[tile * flag_stride + layer slot] counts the
  // pieces of that layer's dZ that are complete in global memory (one release-increment per epilogue warp and N-half); the
  // consumer acquires it before it fetches the blocks (mlp_tc_wgrad.cu)
  int32_t* flags;
  int flag_stride;
  int max_ctas;              // 0: one CTA per SM
  DgStage st[kDgMaxStages];
};

template <int W>
struct DgCfg {
  static constexpr int KB = W / 64;
  // weight ring: 32 KB stages of two K blocks, released by the MMA issuer's commit as soon as their own MMAs are done
  // (with whole 64 KB chunks only two fit at W = 256 and the next chunk's weights could not be requested early enough)
  static constexpr int SKB = 2;
  static constexpr int STAGE_BYTES = SKB * kBlockBytes;
  static constexpr int STAGES = (W == 256) ? 5 : 4;
  static constexpr int TMEM_COLS = (W == 128) ? 512 : 2 * W;   // W = 128 also keeps the skip layer's dZ (64 columns at ACT_COL + W)
  static constexpr int ACC_COL = 0;
  static constexpr int ACT_COL = W;
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_OUT = OFF_RING + STAGES * STAGE_BYTES;     // [8 epilogue warps][4 KB] dZ pieces staged for bulk stores
  static constexpr int OFF_WDEN = OFF_OUT + 8 * 4096;                 // fp32 [W]
  static constexpr int OFF_WRGB = OFF_WDEN + W * 4;                   // fp32 [128][4] (rgb head kernel rows, padded)
  static constexpr int OFF_MISC = OFF_WRGB + 128 * 4 * 4;
  static constexpr int MISC_BYTES = 2048;                             // tmem ptr + barriers (256 B) | publisher mailboxes
  static constexpr int OFF_MAIL = OFF_MISC + 256;                     // [8] heads (128 B) | [8][kMailSlots] counter offsets
  static constexpr int OFF_EPI = OFF_MAIL + 128 + 8 * kMailSlots * 4;  // [kDgMaxStages + 4] stage table of the epilogue warps
  static constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

// Masking / packing of one 32-column group, specialised per stage kind (0: linear, 1: + density term and mask, 2: mask)
// so the unrolled body is branch-free.  Writes the group's 64 bytes of the dZ tile image row and returns the packed words.
template <int KIND>
__device__ __forceinline__ void dg_pack(const uint32_t (&v)[32], uint32_t mw, float gden, uint32_t wden_addr,
                                        uint32_t stage_row, uint32_t r7, int chunk0, uint32_t (&pk)[16]);

// keep a packed bf16 pair where the matching activation halfwords are non-zero (ReLU outputs are +0 or positive)
__device__ __forceinline__ uint32_t mask_pair(uint32_t packed, uint32_t act) {
  const uint32_t m = ((act & 0xFFFFu) ? 0xFFFFu : 0u) | ((act & 0xFFFF0000u) ? 0xFFFF0000u : 0u);
  return packed & m;
}

// keep a packed bf16 pair (word k of a 32-column group) where the forward's mask word has the pair's bits set: element 2k
// is bit 15-k, element 2k+1 bit 31-k; after the shift they are the sign bits of bytes 1 and 3, which PRMT replicates.
__device__ __forceinline__ uint32_t mask_pair_bits(uint32_t packed, uint32_t mw, int k) {
  uint32_t m;      // prmt default mode: selector nibble 8+b = byte b's sign bit replicated over the output byte
  asm("prmt.b32 %0, %1, %2, 0xBB99;" : "=r"(m) : "r"(mw << k), "r"(0u));
  return packed & m;
}

template <int KIND>
__device__ __forceinline__ void dg_pack(const uint32_t (&v)[32], uint32_t mw, float gden, uint32_t wden_addr,
                                        uint32_t stage_row, uint32_t r7, int chunk0, uint32_t (&pk)[16]) {
#pragma unroll
  for (int c8 = 0; c8 < 4; ++c8) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v0 = __uint_as_float(v[c8 * 8 + 2 * e]), v1 = __uint_as_float(v[c8 * 8 + 2 * e + 1]);
      if (KIND == 1) {
        const float2 wd = lds64(wden_addr + (c8 * 8 + 2 * e) * 4);
        v0 = fmaf(gden, wd.x, v0); v1 = fmaf(gden, wd.y, v1);
      }
      uint32_t pr = cvt_bf16x2(v0, v1);
      if (KIND != 0) pr = mask_pair_bits(pr, mw, c8 * 4 + e);
      pk[c8 * 4 + e] = pr;
    }
  }
  // the row's 64 bytes go to the warp's staging piece (image order), from where one bulk store takes them to HBM
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_row + (((uint32_t)(chunk0 + q4) ^ r7) << 4)), "r"(pk[4 * q4]),
                 "r"(pk[4 * q4 + 1]), "r"(pk[4 * q4 + 2]), "r"(pk[4 * q4 + 3]) : "memory");
}

template <int W>
__global__ void __launch_bounds__(384, 1)
mlp_tc_dgrad_kernel(const __grid_constant__ DgParams p) {
  using C = DgCfg<W>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by an OFFSET into the __shared__ array: the pointer keeps its address space (ld/st.shared, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;      // warp-uniform for the compiler
  float* s_wden = reinterpret_cast<float*>(smem + C::OFF_WDEN);   // filled below, read with ld.shared
  float* s_wrgb = reinterpret_cast<float*>(smem + C::OFF_WRGB);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + C::OFF_MISC);
  const uint32_t bar0 = sbase + C::OFF_MISC + 16;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (8 + s); };
  auto bar_acc_full = [&](int h) { return bar0 + 8 * (16 + h); };
  auto bar_a_ready = [&](int h) { return bar0 + 8 * (18 + h); };
  const uint32_t bar_p_ready = bar0 + 8 * 20;
  static_assert(16 + 8 * 21 <= 256 && 256 + 128 + 8 * kMailSlots * 4 + (kDgMaxStages + 4) * 4 <= C::MISC_BYTES, "barrier area + mailboxes + stage table");

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int h = 0; h < 2; ++h) { mbar_init(bar_acc_full(h), 1); mbar_init(bar_a_ready(h), 8); }   // one arrival per epilogue warp
    mbar_init(bar_p_ready, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) reinterpret_cast<uint32_t*>(smem + C::OFF_MAIL)[threadIdx.x] = 0;      // mailbox heads
  if (threadIdx.x >= 32 && threadIdx.x < 32 + kDgMaxStages + 4) {
    // what the epilogue warps need of a stage, one word (a constant-bank lookup with a run-time index costs them ~250 cycles
    // per epilogue): kind 0-1 | n_halves 2-3 | last_part 4 | feeds_next 5 | o_sel 6 | to_skip 7 | out_slot 8-15 | mask layer 16-20
    const int si = threadIdx.x - 32;
    uint32_t w = 0;
    if (si < p.n_stages) {
      const DgStage& S = p.st[si];
      const int mg = (S.kind == 1 || S.kind == 2) ? S.mask_slot / C::KB : 31;
      w = (uint32_t)S.kind | ((uint32_t)S.n_halves << 2) | (S.last_part ? 16u : 0u) | (S.o_sel >= 0 ? 32u : 0u) |
          ((S.o_sel > 0 ? 1u : 0u) << 6) | (S.to_skip ? 128u : 0u) | ((uint32_t)S.out_slot << 8) | ((uint32_t)mg << 16);
    }
    reinterpret_cast<uint32_t*>(smem + C::OFF_EPI)[si] = w;
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = threadIdx.x; i < W; i += blockDim.x) s_wden[i] = p.params[p.off_wden + i];
  for (int i = threadIdx.x; i < 128 * 4; i += blockDim.x) s_wrgb[i] = (i & 3) < 3 ? p.params[p.off_wrgb + (i >> 2) * 3 + (i & 3)] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int num_tiles = p.count ? min(*p.count, p.M) : p.M;
  const int last_halves = p.st[p.n_stages - 1].n_halves;

  if (warp == 0) {
    // ===== weight producer: one ring stage per (stage, N-half) =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
        for (int s = 0; s < p.n_stages; ++s) {
          const DgStage S = p.st[s];
          for (int nh = 0; nh < S.n_halves; ++nh)
            for (int kb0 = 0; kb0 < S.n_kb; kb0 += C::SKB) {
              const uint32_t bytes = (uint32_t)min(C::SKB, S.n_kb - kb0) * kBlockBytes;
              mbar_wait(bar_empty(stage), phase ^ 1);
              mbar_arrive_expect_tx(bar_full(stage), bytes);
              bulk_g2s(sbase + C::OFF_RING + stage * C::STAGE_BYTES, p.packed_t + (size_t)(S.block0 + nh * S.n_kb + kb0) * kBlockBytes,
                       bytes, bar_full(stage));
              if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the schedule (uniform control flow), one elected lane issues =====
    {
      constexpr uint32_t idesc128 = umma_idesc(128, 128), idesc64 = umma_idesc(128, 64);
      constexpr uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      uint32_t stage = 0, phase = 0, ar_par[2] = {0, 0}, pr_par = 0;
      int it = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const bool tr = DURF_TRACE_DETAIL && p.trace && blockIdx.x == 0;
      long long t_begin = clock64(), t_start = 0, tq = 0;
      // Barrier waits are software-pipelined as in the forward kernel (umma_kblock_conv): each K block of MMAs probes the
      // barriers of the NEXT K block before its MMAs and consumes the outcome after them, because a wait executed between
      // two groups of MMAs returns only once the tensor queue has drained.
      if (blockIdx.x < num_tiles) mbar_wait(bar_full(0), 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const bool more_tiles = tile + (int)gridDim.x < num_tiles;
        for (int s = 0; s < p.n_stages; ++s) {
          const DgStage S = p.st[s];
          const uint32_t a_buf = tmem_u + C::ACT_COL + S.a_sel * (W / 2);
          const uint32_t idesc = S.ncols == 64 ? idesc64 : idesc128;
          const bool waits_a = s > 0 && S.a_sel != 2;       // A operand written by the previous stage's epilogue, half by half
          for (int nh = 0; nh < S.n_halves; ++nh) {
            const uint32_t d_addr = tmem_u + C::ACC_COL + nh * 128;
            if (s == 0 && nh == 0) {
              if (tr) tq = clock64();
              if (it > 0)      // accumulators of the previous tile's last stage must have been drained
                for (int h = 0; h < last_halves; ++h) { mbar_wait(bar_a_ready(h), ar_par[h]); ar_par[h] ^= 1; }
              mbar_wait(bar_p_ready, pr_par); pr_par ^= 1;           // dZ_cond is in TMEM
              if (tr) t_start += clock64() - tq;
            }
            tc_fence_after();      // this chunk's weights (and its first K block) were waited for by the previous K block
            // the chunk after this one: (s, nh+1), else (s+1, 0), else the next tile's first
            const bool last_in_stage = nh + 1 == S.n_halves;
            const bool last_chunk = last_in_stage && s + 1 == p.n_stages;
            const uint32_t need_w = (!last_chunk || more_tiles) ? 1u : 0u;
            const uint32_t need_a0 = (last_in_stage && !last_chunk && p.st[s + 1].a_sel != 2) ? 1u : 0u;
#pragma unroll
            for (int kb = 0; kb < C::KB; ++kb) {
              if (kb < S.n_kb) {
                if (waits_a && nh == 0 && (kb & 1) == 0) { ar_par[kb >> 1] ^= 1; tc_fence_after(); }   // half kb/2 of the A operand is there
                const uint32_t acc0 = (S.first_part && kb == 0) ? 0u : 1u;
                const uint32_t b_lo = (((sbase + C::OFF_RING + stage * C::STAGE_BYTES + (kb % C::SKB) * kBlockBytes) & 0x3FFFF) >> 4) | (1u << 16);
                const bool chunk_end = kb + 1 == S.n_kb;
                const bool stage_end = chunk_end || (kb % C::SKB) == C::SKB - 1;      // last K block read from this ring stage
                const uint32_t next_stage = (stage + 1 == C::STAGES) ? 0 : stage + 1;
                const uint32_t next_phase = (stage + 1 == C::STAGES) ? phase ^ 1 : phase;
                if (!chunk_end)
                  umma_kblock_conv<true>(d_addr, a_buf + kb * 32, b_lo, desc_hi, idesc, acc0,
                                         bar_a_ready(((kb + 1) >> 1) & 1), ar_par[((kb + 1) >> 1) & 1],
                                         (waits_a && nh == 0 && ((kb + 1) & 1) == 0) ? 1u : 0u,
                                         bar_full(next_stage), next_phase, stage_end ? 1u : 0u, bar_acc_full(nh), 0u,
                                         bar_empty(stage), stage_end ? 1u : 0u, 0u);
                else
                  umma_kblock_conv<true>(d_addr, a_buf + kb * 32, b_lo, desc_hi, idesc, acc0,
                                         bar_a_ready(0), ar_par[0], need_a0, bar_full(next_stage), next_phase, need_w,
                                         bar_acc_full(nh), S.last_part ? 1u : 0u, bar_empty(stage), 1u, 0u);
                if (stage_end) { stage = next_stage; phase = next_phase; }
              }
            }
          }
        }
      }
      if (tr && lane == 0) printf("durf dgrad trace: MMA thread: %d tiles, total %lld cyc; tile start (drained accumulators, dZ_cond) %lld\n",
                                  it, clock64() - t_begin, t_start);
    }
  } else if (warp == 3) {
    // ===== publisher: lane l < 8 drains the mailbox of epilogue warp l.  One gpu-scope fence per batch (it orders the dZ
    // pieces, whose completion the epilogue warps observed before they posted, before the increments), then one relaxed
    // increment per piece: fence + relaxed atomic = release.  The fence's latency is hidden in this warp. =====
    if (p.flags) {
      const uint32_t head_addr = smem_u32(smem + C::OFF_MAIL) + (lane & 7) * 16;
      const int* ring = reinterpret_cast<const int*>(smem + C::OFF_MAIL + 128) + (lane & 7) * kMailSlots;
      uint32_t tail = 0;
      long long t0 = 0;
      while (true) {
        uint32_t h = 0x80000000u;                   // lanes >= 8: nothing to do, done
        if (lane < 8) asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(h) : "r"(head_addr) : "memory");
        const bool fin = (h >> 31) != 0;
        h &= 0x7FFFFFFFu;
        if (lane >= 8) h = tail;
        const bool fresh = h != tail;
        if (__any_sync(kFull, fresh)) {
          __threadfence();
          for (; tail != h; ++tail)
            asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(p.flags + ring[tail & (kMailSlots - 1)]) : "memory");
          t0 = 0;
        } else {
          __nanosleep(200);
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 8000000000LL) { printf("durf dgrad kernel: publisher of block %d timed out\n", blockIdx.x); __trap(); }
        }
        if (__all_sync(kFull, fin && tail == h)) break;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = sample row; warps q and q+4 share TMEM lane quarter q and split a half's 128 columns =====
    const int q = warp & 3, ch = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t af_par[2] = {0, 0};
    uint32_t mw_next[2] = {0, 0};                 // mask words of the next masked epilogue, fetched one epilogue ahead
    // counters (offsets into p.flags, -1 = none) of this warp's kPubDepth most recent dZ pieces, oldest first: a piece is
    // published once at most kPubDepth - 1 younger bulk stores are still pending (cp.async.bulk.wait_group), i.e. kPubDepth
    // epilogues (~15 k cycles) after it was staged - a store's completion in L2 takes several microseconds under load, and a
    // shorter queue stalled the epilogue (the critical path of this kernel) on it
    constexpr int kPubDepth = 6;
    // The release itself (a gpu-scope fence: ~1.5 k cycles, it waits for the SM's outstanding stores) is NOT executed here:
    // lane 0 drops the counter's offset into this warp's mailbox and the otherwise idle warp 3 publishes it.
    uint32_t* mail_head = reinterpret_cast<uint32_t*>(smem + C::OFF_MAIL) + (warp - 4) * 4;
    int* mail_ring = reinterpret_cast<int*>(smem + C::OFF_MAIL + 128) + (warp - 4) * kMailSlots;
    uint32_t mail_n = 0;
    auto mail_push = [&](int off) {
      mail_ring[mail_n & (kMailSlots - 1)] = off;
      ++mail_n;
      asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(mail_head)), "r"(mail_n) : "memory");
    };
    int pq[kPubDepth];
#pragma unroll
    for (int i = 0; i < kPubDepth; ++i) pq[i] = -1;
    // stage table (see the word's layout above): read with ld.shared and broadcast from lane 0, so that the compiler knows the
    // word - and the stage kind, slot and flags decoded from it - is warp-uniform (uniform branches, uniform store addresses)
    auto s_epi = [&](int i) {
      uint32_t v;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + C::OFF_EPI + 4u * (uint32_t)i));
      return __shfl_sync(kFull, v, 0);
    };
    const uint32_t my_out = sbase + C::OFF_OUT + ((warp - 4) << 12);
    const uint32_t row_off = (lane >> 3) * 1024 + (lane & 7) * 128, r7 = lane & 7;
    // this warp's 4 KB piece inside a layer record [sample half][64-column block][64 rows x 128 B]: `nb` blocks per half
    const uint32_t piece_w = (uint32_t)(q >> 1) * (C::KB * 8192) + ch * 8192 + (q & 1) * 4096;      // + h * 16384: a W-wide layer
    const uint32_t piece_c = (uint32_t)(q >> 1) * (2 * 8192) + ch * 8192 + (q & 1) * 4096;          // the condition layer (2 blocks)
    const size_t mask_tile_words = (size_t)(p.depth + 1) * (W / 32) * 128;
    // The staging piece is free once the previous bulk store has read it; with tile_done counters the oldest queued piece is
    // handed to the publisher once it is complete in global memory.
    auto staging_acquire = [&](int flag_off) {
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (p.flags) {
          asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kPubDepth - 1) : "memory");
          if (pq[0] >= 0) mail_push(pq[0]);
        }
      }
#pragma unroll
      for (int i = 0; i + 1 < kPubDepth; ++i) pq[i] = pq[i + 1];
      pq[kPubDepth - 1] = p.flags ? flag_off : -1;
      __syncwarp();
    };
    auto staging_store = [&](uint8_t* gdst) {      // generic-proxy writes of the whole warp -> one 4 KB bulk store
      __syncwarp();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(gdst), "r"(my_out) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };
    // inputs of a tile's prologue (upstream gradients of this row, mask words of the condition layer): fetched while the
    // previous tile's last stage runs, so that no global-memory latency sits at the tile boundary
    float pre_g[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t pre_mw[2] = {0u, 0u};
    auto prefetch_tile = [&](int t2) {
      const int ray2 = p.ray_index ? p.ray_index[t2] : t2;
      const float* g3 = p.d_raw_rgb + ((size_t)ray2 * kTileM + row) * 3;
      pre_g[0] = __ldg(g3); pre_g[1] = __ldg(g3 + 1); pre_g[2] = __ldg(g3 + 2);
      pre_g[3] = __ldg(p.d_raw_density + (size_t)ray2 * kTileM + row);
      const uint32_t* src = p.masks + (size_t)t2 * mask_tile_words + ((size_t)p.depth * (W / 32) + ch * 2) * 128 + row;
#pragma unroll
      for (int i = 0; i < 2; ++i) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(pre_mw[i]) : "l"(src + i * 128));
    };
    if ((int)blockIdx.x < num_tiles) prefetch_tile(blockIdx.x);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      uint8_t* dzt = p.dz + (size_t)tile * p.saved_blocks * kBlockBytes;
      const uint32_t* mrow = p.masks + (size_t)tile * mask_tile_words + (size_t)(ch * 2) * 128 + row;
      const float gden = pre_g[3];
      // ---- prologue: dZ_cond[row, c] = (sum_j d_rgb[row, j] W_rgb[c, j]) * [cond_act[row, c] > 0], c in this warp's 64 columns;
      // the ReLU mask is the forward's 1-bit word, the row goes to TMEM (A operand of stage 0) and, through the staging piece,
      // to the dZ record (B operand of the weight-gradient kernel)
      {
        const float g0 = pre_g[0], g1 = pre_g[1], g2 = pre_g[2];
        const uint32_t mwc[2] = {pre_mw[0], pre_mw[1]};
        const int c0 = ch * 64;
        staging_acquire(tile * p.flag_stride + p.cond_slot / C::KB);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int c = c0 + i * 32 + 2 * e;
            const float4 w0 = lds128(sbase + C::OFF_WRGB + c * 16), w1 = lds128(sbase + C::OFF_WRGB + (c + 1) * 16);
            const float v0 = fmaf(g2, w0.z, fmaf(g1, w0.y, g0 * w0.x)), v1 = fmaf(g2, w1.z, fmaf(g1, w1.y, g0 * w1.x));
            pk[e] = mask_pair_bits(cvt_bf16x2(v0, v1), mwc[i], e);
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_out + row_off + (((uint32_t)(i * 4 + q4) ^ r7) << 4)),
                         "r"(pk[4 * q4]), "r"(pk[4 * q4 + 1]), "r"(pk[4 * q4 + 2]), "r"(pk[4 * q4 + 3]) : "memory");
          tmem_st16(t_lane + C::ACT_COL + (c0 + i * 32) / 2, pk);     // A operand of stage 0 = activation buffer 0
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p_ready);
        staging_store(dzt + (size_t)p.cond_slot * kBlockBytes + piece_c);
      }
      uint32_t ew = s_epi(0);
      for (int s = 0; s < p.n_stages; ++s) {
        const uint32_t w = ew;
        ew = s_epi(s + 1);                             // the next stage's entry (table is padded): off the post-barrier path
        if (s + 1 == p.n_stages && tile + (int)gridDim.x < num_tiles) prefetch_tile(tile + (int)gridDim.x);
        if (!(w & 16u)) continue;                      // entries that only accumulate have no epilogue
        const int kind = w & 3, n_halves = (w >> 2) & 3, out_slot = (w >> 8) & 0xFF;
        const bool feeds_next = (w & 32u) != 0, to_skip = (w & 128u) != 0;
        const uint32_t o_buf = t_lane + C::ACT_COL + ((w >> 6) & 1u) * (W / 2);
        for (int h = 0; h < n_halves; ++h) {
          const int col0 = h * 128 + ch * 64;
          mbar_wait(bar_acc_full(h), af_par[h]); af_par[h] ^= 1;
          tc_fence_after();
          if (kind == 3) {
            // input gradient: 64 accumulator columns (the in_dim features), fp32 rows straight to d_features
            uint32_t v[32];
            tmem_ld32_issue(t_lane + C::ACC_COL + ch * 32, v);
            tmem_ld_wait();
            tmem_ld_pin(v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a_ready(0));               // accumulators drained
            float* dx = p.d_features + ((size_t)tile * kTileM + row) * p.in_dim;
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (ch * 32 + e < p.in_dim) dx[ch * 32 + e] = __uint_as_float(v[e]);
            continue;
          }
          uint32_t v[32];
          tmem_ld32_issue(t_lane + C::ACC_COL + col0, v);
          const uint32_t mw[2] = {mw_next[0], mw_next[1]};
          staging_acquire(tile * p.flag_stride + out_slot / C::KB);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            tmem_ld_wait();
            tmem_ld_pin(v);
            uint32_t pk[16];
            const uint32_t wden_addr = sbase + C::OFF_WDEN + (col0 + i * 32) * 4;
            if (kind == 0) dg_pack<0>(v, mw[i], gden, wden_addr, my_out + row_off, r7, i * 4, pk);
            else if (kind == 1) dg_pack<1>(v, mw[i], gden, wden_addr, my_out + row_off, r7, i * 4, pk);
            else dg_pack<2>(v, mw[i], gden, wden_addr, my_out + row_off, r7, i * 4, pk);
            if (i == 0) tmem_ld32_issue(t_lane + C::ACC_COL + col0 + 32, v);
            if (feeds_next) tmem_st16(o_buf + (col0 + i * 32) / 2, pk);
            if (W == 128 && to_skip) tmem_st16(t_lane + C::ACT_COL + 2 * (W / 2) + (col0 + i * 32) / 2, pk);
          }
          if (feeds_next || to_skip) tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_a_ready(h));     // next stage's A operand half is in TMEM (last stage: accumulators drained)
          // off the critical path: the dZ piece leaves by one bulk store, the next masked epilogue's mask words are requested
          staging_store(dzt + (size_t)out_slot * kBlockBytes + h * 16384 + piece_w);
          uint32_t wn = w;
          int h2 = h + 1;
          if (h2 >= n_halves) {                          // the next entry with an epilogue (entries are padded with zeros)
            h2 = 0; wn = ew;
            for (int s2 = s + 2; !(wn & 16u) && s2 <= p.n_stages; ++s2) wn = s_epi(s2);
          }
          const int mg = (wn >> 16) & 31;                // trunk layer whose ReLU mask gates that epilogue's output (31: none)
          if ((wn & 16u) && mg != 31) {
            const uint32_t* src = mrow + ((size_t)mg * (W / 32) + h2 * 4) * 128;
#pragma unroll
            for (int i = 0; i < 2; ++i) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(mw_next[i]) : "l"(src + i * 128));
          }
        }
      }
    }
    if (lane == 0) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every staged dZ piece is in HBM
#pragma unroll
      for (int i = 0; i < kPubDepth; ++i)
        if (pq[i] >= 0) mail_push(pq[i]);
      if (p.flags) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(mail_head)), "r"(mail_n | 0x80000000u) : "memory");   // done
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
// Transposed weight image: for every backward stage, N half and K block one 128 x 64 block with
// block[r][k] = kernel[in = n_first + r][out = k_first + k] (the contraction runs over the forward layer's outputs); written
// by pack_blocks_kernel (mlp_tc.cu) from the PackBlock list built here.
bool mlp_tc_bwd_supported(const DurfMlpTopology& t) {
  return (t.width == 256 || t.width == 128) && t.cond_width == 128 && t.in_dim <= 64 && t.depth >= 2 && t.depth <= 9 &&
         t.cond_dim <= 32 && !(((t.depth - 1) % t.skip == 0) && t.depth - 1 > 0);
}

int mlp_tc_saved_blocks(const DurfMlpTopology& t) { return (t.depth + 1) * (t.width / 64) + t.cond_width / 64; }

// stage list + per-block source description; returns the number of blocks of the transposed image.  The image always
// carries the input-gradient blocks (width 128 only); `want_dx` decides whether the kernel walks those last two entries.
static int build_dg(const DurfMlpTopology& t, DgParams& P, std::vector<PackBlock>* pp, bool want_dx = false,
                    const float* params = nullptr, uint8_t* packed_t = nullptr) {
  MlpLayout L(t);
  const int KB = t.width / 64, NH = t.width / 128;
  auto slot = [&](int g) { return g * KB; };
  int ns = 0, blocks = 0;
  // entry: `n_out_blocks` x n_kb blocks with block[r][k] = kernel(layer)[in_first + 128 nh + r][64 kb + k]
  auto add = [&](int layer, int n_kb, int kind, int mask_slot, int out_slot, int in_first = 0, int in_avail = 1 << 30) -> DgStage& {
    DgStage& S = P.st[ns];
    S.n_halves = (kind == 3) ? 1 : NH; S.n_kb = n_kb; S.kind = kind; S.mask_slot = mask_slot; S.out_slot = out_slot; S.block0 = blocks;
    S.a_sel = ns & 1; S.o_sel = (ns + 1) & 1; S.first_part = 1; S.last_part = 1; S.parts = 1; S.to_skip = 0; S.ncols = 128;
    ++ns;
    for (int nh = 0; nh < S.n_halves; ++nh)
      for (int kb = 0; kb < n_kb; ++kb, ++blocks)
        if (pp) {
          const int ld = L.out_dim[layer], row0 = in_first + nh * 128, left = in_avail - nh * 128;
          PackBlock b;
          b.src = params + L.w_off[layer] + (size_t)row0 * ld + kb * 64;
          b.dst = packed_t + (size_t)blocks * kBlockBytes;
          b.sr = ld; b.sk = 1;
          b.r_avail = left < 128 ? (left < 0 ? 0 : left) : 128;
          b.k_avail = ld - kb * 64 < 64 ? ld - kb * 64 : 64;
          pp->push_back(b);
        }
    return S;
  };
  // stage 0: dBott = dZ_cond W_cond[:width]^T   (contraction over the 128 condition outputs)
  add(t.depth + 2, t.cond_width / 64, 0, 0, slot(t.depth));
  // stage 1: dZ_{depth-1} = (dBott W_bott^T + d_den (x) w_den) * [a_{depth-1} > 0]
  add(t.depth + 1, KB, 1, slot(t.depth - 1), slot(t.depth - 1));
  // stages 2..: dZ_{g-1} = (dZ_g W_g[:width]^T) * [a_{g-1} > 0]
  int skip_layer = -1;                                       // trunk layer whose input is [a_{g-1} | x]
  for (int g = 1; g < t.depth; ++g)
    if ((g - 1) % t.skip == 0 && g - 1 > 0) skip_layer = g;
  for (int g = t.depth - 1; g >= 1; --g) {
    DgStage& S = add(g, KB, 2, slot(g - 1), slot(g - 1));
    if (g - 1 == skip_layer) S.to_skip = 1;                  // this stage's output is dZ of the skip layer: keep it
  }
  const int n_core = ns;
  P.st[n_core - 1].o_sel = -1;
  // input gradient (width 128: the kept dZ fits in TMEM): dX = dZ_0 W_0^T + dZ_skip W_skip[width:]^T, 64 output columns
  if (t.width == 128) {
    DgStage& F0 = add(0, KB, 3, 0, 0, 0, t.in_dim);
    F0.a_sel = n_core & 1; F0.o_sel = -1; F0.last_part = (skip_layer < 0) ? 1 : 0; F0.ncols = 64;
    if (skip_layer >= 0) {
      DgStage& F1 = add(skip_layer, KB, 3, 0, 0, t.width, t.in_dim);
      F1.a_sel = 2; F1.o_sel = -1; F1.first_part = 0; F1.parts = 2; F1.ncols = 64;
    }
    if (want_dx) P.st[n_core - 1].o_sel = n_core & 1;        // dZ_0 must reach TMEM as the A operand of the first part
  }
  P.n_stages = (want_dx && t.width == 128) ? ns : n_core;
  if (!want_dx)
    for (int i = 0; i < n_core; ++i) P.st[i].to_skip = 0;
  P.cond_slot = slot(t.depth + 1);
  P.off_wden = (int)L.w_off[t.depth];
  P.off_wrgb = (int)L.w_off[t.depth + 3];
  return blocks;
}

int64_t mlp_tc_packed_t_bytes(const DurfMlpTopology& t) {
  if (!mlp_tc_bwd_supported(t)) return 0;
  DgParams P;
  return (int64_t)build_dg(t, P, nullptr) * kBlockBytes;
}

// Appends the block descriptions of the transposed image to `out`; returns their number (or a negative error code).
int mlp_tc_pack_t_blocks(const DurfMlpTopology& t, const float* params, void* packed_t, std::vector<PackBlock>& out) {
  DURF_REQUIRE(mlp_tc_bwd_supported(t), DURF_E_UNSUPPORTED, "durf_mlp_pack_weights: no tensor-core backward for this topology");
  DgParams P;
  return build_dg(t, P, &out, false, params, (uint8_t*)packed_t);
}

int mlp_tc_dgrad_launch(cudaStream_t st, const DurfMlpTopology& t, const DgParams& base) {
  DgParams P = base;
  build_dg(t, P, nullptr, base.d_features != nullptr);
  P.d_features = base.d_features; P.in_dim = t.in_dim;
  P.saved_blocks = mlp_tc_saved_blocks(t);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (P.max_ctas > 0 && P.max_ctas < sms) sms = P.max_ctas;
  int grid = P.M < sms ? P.M : sms;
  cudaError_t e;
  if (t.width == 256) {
    e = cudaFuncSetAttribute(mlp_tc_dgrad_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, DgCfg<256>::SMEM_BYTES);
    DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_bwd(bf16): smem attribute: %s", cudaGetErrorString(e));
    if (P.flags && grid >= 2) {
      // sharing the GPU with the weight-gradient kernel: launched as CTA pairs so that the two SMs of a TPC run the SAME kernel
      // (the CTAs do not communicate)
      grid &= ~1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = DgCfg<256>::SMEM_BYTES; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, mlp_tc_dgrad_kernel<256>, P);
      DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_bwd(bf16): dgrad launch: %s", cudaGetErrorString(e));
    } else
    mlp_tc_dgrad_kernel<256><<<grid, 384, DgCfg<256>::SMEM_BYTES, st>>>(P);
  } else {
    e = cudaFuncSetAttribute(mlp_tc_dgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, DgCfg<128>::SMEM_BYTES);
    DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_bwd(bf16): smem attribute: %s", cudaGetErrorString(e));
    mlp_tc_dgrad_kernel<128><<<grid, 384, DgCfg<128>::SMEM_BYTES, st>>>(P);
  }
  DURF_CHECK_LAUNCH("durf_mlp_bwd(bf16): dgrad");
  return DURF_OK;
}

}  // namespace durf

// ---- durf_mlp_bwd(DURF_PREC_BF16): dgrad chain, then the weight-gradient kernel ---------------------------------------
namespace durf {

struct WgradParams;
int64_t mlp_tc_packed_bytes(const DurfMlpTopology& t);
int mlp_tc_wgrad_run(cudaStream_t st, const DurfMlpTopology& t, const uint8_t* saved, const uint8_t* feat, const uint8_t* dz,
                     const float* d_raw_rgb, const float* d_raw_density, const float* cond, const int32_t* ray_index,
                     const int32_t* count, int M, float* d_params, const int32_t* tile_done, int max_ctas);

// The two halves of the tensor-core backward.  `tile_done` (may be NULL) is the [M, depth + 2] counter array through which
// the data-gradient kernel tells a concurrently running weight-gradient kernel which dZ blocks are complete.
static int tc_bwd_check(const DurfMlpArgs& a, const char* who) {
  const DurfMlpTopology& t = a.topo;
  DURF_REQUIRE(mlp_tc_bwd_supported(t), DURF_E_UNSUPPORTED, "%s: no tensor-core backward for this topology", who);
  DURF_REQUIRE(a.N == kTileM, DURF_E_UNSUPPORTED, "%s: needs 128 samples per ray (got %d)", who, a.N);
  DURF_REQUIRE(a.saved && a.packed && a.features, DURF_E_INVALID, "%s: needs saved activations, weight images, feature tiles", who);
  const size_t need = (size_t)a.M * mlp_tc_saved_blocks(t) * kBlockBytes;
  DURF_REQUIRE(a.workspace && a.workspace_bytes >= need, DURF_E_WORKSPACE, "%s: workspace %zu < %zu bytes", who, a.workspace_bytes, need);
  return DURF_OK;
}

int mlp_tc_backward_data(cudaStream_t st, const DurfMlpArgs& a, const float* d_raw_rgb, const float* d_raw_density,
                         float* d_features, int32_t* tile_done, int max_ctas) {
  DURF_REQUIRE(d_features == nullptr || a.topo.width == 128, DURF_E_UNSUPPORTED,
               "durf_mlp_bwd(bf16): the input gradient is available for width-128 networks (BoxMLP) only");
  int rc = tc_bwd_check(a, "durf_mlp_bwd(bf16)");
  if (rc != DURF_OK) return rc;
  const DurfMlpTopology& t = a.topo;
  DgParams P{};
  P.saved = (const uint8_t*)a.saved; P.dz = (uint8_t*)a.workspace;
  P.masks = reinterpret_cast<const uint32_t*>(P.saved + (size_t)a.M * mlp_tc_saved_blocks(t) * kBlockBytes);
  P.depth = t.depth;
  P.packed_t = (const uint8_t*)a.packed + mlp_tc_packed_bytes(t);
  P.params = a.params; P.d_raw_rgb = d_raw_rgb; P.d_raw_density = d_raw_density;
  P.ray_index = a.ray_index; P.count = a.count; P.M = a.M; P.d_features = d_features;
  P.flags = tile_done; P.flag_stride = t.depth + 2; P.max_ctas = max_ctas;
  static const int trace_env = getenv("DURF_TC_TRACE") ? atoi(getenv("DURF_TC_TRACE")) : 0;
  P.trace = trace_env;
  return mlp_tc_dgrad_launch(st, t, P);
}

int mlp_tc_backward_weights(cudaStream_t st, const DurfMlpArgs& a, const float* d_raw_rgb, const float* d_raw_density,
                            float* d_params, const int32_t* tile_done, int max_ctas) {
  int rc = tc_bwd_check(a, "durf_mlp_bwd(bf16)");
  if (rc != DURF_OK) return rc;
  return mlp_tc_wgrad_run(st, a.topo, (const uint8_t*)a.saved, (const uint8_t*)a.features, (const uint8_t*)a.workspace, d_raw_rgb,
                          d_raw_density, a.cond, a.ray_index, a.count, a.M, d_params, tile_done, max_ctas);
}

int mlp_tc_backward(cudaStream_t st, const DurfMlpArgs& a, const float* d_raw_rgb, const float* d_raw_density, float* d_params,
                    float* d_features) {
  int rc = mlp_tc_backward_data(st, a, d_raw_rgb, d_raw_density, d_features, nullptr, 0);
  if (rc != DURF_OK) return rc;
  return mlp_tc_backward_weights(st, a, d_raw_rgb, d_raw_density, d_params, nullptr, 0);
}

}  // namespace durf
