// K2 backward, weight gradients on the tensor cores (sm_100a): for every Dense layer of MLP / BoxMLP
// (obbpose_model.py:326-353, 390-417 under jax.value_and_grad, train_boxpose.py:251)
//     dW[in, out] += A^T dZ        db[out] += column sums of dZ
// summed over all samples.  A (the layer's bf16 input activations, saved by the forward kernel) and dZ (the bf16
// pre-activation gradients written by the dgrad kernel) live in HBM as per-tile block images: 64-feature blocks of
// 128-byte rows, SWIZZLE_128B, a layer's record ordered [sample half][block][64 rows] so that the 64 samples of all its
// blocks are one contiguous 32 KB piece.  For wgrad the SAMPLE index is the contraction index, so the very same images are
// MN-major tcgen05 operands (the 128-byte rows run along M / N, the 8-row groups along K): no transpose is ever made.
//
// The work is a list of jobs (one per weight matrix).  Every CTA is bound to ONE job and takes that job's tiles round-robin
// (the host splits the SMs over the jobs in proportion to the bytes a job streams per tile): the job's [in <= 256, out <= 256]
// fp32 accumulator sits in TMEM (<= 512 columns) for the whole kernel while the CTA streams its tiles through a 3-stage
// ring (64 samples of A and dZ per stage, cp.async.bulk + mbarrier), and is added to the global gradient ONCE at the end
// with 16-byte fp32 reductions.  (Round 1 gave every CTA a tile range and ALL jobs: 11 accumulator read-outs of 64 K scalar
// reductions per CTA and launch = ~0.5 ms per launch whatever the batch, 74 % of a 512-ray step.)  The kernel is HBM-bound (64 B/cycle/SM of operand
// bytes at full tensor rate), so the narrow heads (density N=1, rgb N=3, the 27 view inputs of the condition layer) and
// all bias gradients are computed by the otherwise idle warps from the stages already in shared memory.
#include "tc_common.cuh"
#include "mlp_topology.h"
#include <stdlib.h>

namespace durf {

constexpr int kWgStages = 3;
constexpr int kHalfBlock = 8192;          // 64 samples of one block image
constexpr int kWgStageBytes = 8 * kHalfBlock;
constexpr int kMaxJobs = 16;
constexpr int kMaxCtas = 160;

struct WgradJob {
  int a_src;        // 0: saved activations, 1: input-feature tiles
  int a_off;        // first block of A inside the tile record
  int a_blocks;     // 64-feature blocks of A (1, 2 or 4)
  int z_off;        // first block of dZ inside the dz tile record
  int z_blocks;     // 64-column blocks of dZ (0 = no tensor-core work: rgb-head pseudo job)
  int dw_off;       // float offset of kernel[k_first][0] inside the gradient blob
  int ld;           // out dim of the kernel
  int k_valid;      // valid input rows (60 / 63 for the input-feature block, else 64 * a_blocks)
  int n_valid;      // valid output columns
  int db_off;       // float offset of the bias gradient, -1: none (second K part of the skip layer)
  int extra;        // 0 none; 1: density head (vec = d_raw_density); 2: rgb head (vec = d_raw_rgb); 3: view part of the condition layer
  int ex_off;       // float offset of the extra's kernel gradient
  int ex_b_off;     // float offset of the extra's bias gradient (-1: none)
  int flag_idx;     // slot of this job's dZ inside a tile's row of completion counters (-1: the job reads no dZ)
  int flag_need;    // counter value at which the dZ blocks of a tile are complete
};

struct WgradParams {
  const uint8_t* saved;       // forward activations, saved_blocks blocks per tile
  const uint8_t* feat;        // input-feature tiles, 1 block per tile
  const uint8_t* dz;          // dgrad outputs, dz_blocks blocks per tile
  int saved_blocks, dz_blocks;
  const float* d_raw_rgb;     // [B,128,3]
  const float* d_raw_density; // [B,128]
  const float* cond;          // [B,cond_dim]
  const int32_t* ray_index;
  const int32_t* count;
  int M, cond_dim;
  float* d_params;
  int n_jobs;
  WgradJob jobs[kMaxJobs];
  uint8_t cta_job[kMaxCtas];   // job of CTA b
  uint8_t cta_part[kMaxCtas];  // which share of the job's tiles
  uint8_t job_parts[kMaxJobs]; // CTAs bound to the job
  long long* trace;            // [opt] debugging (DURF_WGRAD_TRACE=1): cycles of every CTA
  int trace_detail;            // DURF_WGRAD_TRACE=2: job 1's first CTA prints where its producer / MMA / aux threads waited
  // [opt] running CONCURRENTLY with the data-gradient kernel (on other SMs): flags[tile * flag_stride + job.flag_idx] reaches
  // job.flag_need when the tile's dZ blocks of that layer are complete in global memory (mlp_tc_dgrad.cu, publish_piece)
  const int32_t* flags;
  int flag_stride;
};

__device__ __forceinline__ void red_add(float* addr, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {      // addr 16-byte aligned
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(384, 1)
mlp_tc_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by an OFFSET into the __shared__ array: the pointer keeps its address space (ld/st.shared, not generic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;      // warp-uniform for the compiler
  constexpr int OFF_MISC = kWgStages * kWgStageBytes;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + OFF_MISC);
  const uint32_t bar0 = sbase + OFF_MISC + 16;
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (4 + s); };
  const uint32_t bar_acc_done = bar0 + 8 * 8, bar_acc_free = bar0 + 8 * 9;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1 + 8); }   // MMA commit + 8 aux warps
    mbar_init(bar_acc_done, 1); mbar_init(bar_acc_free, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int num_tiles = p.count ? min(*p.count, p.M) : p.M;
  // The CTAs of a job take the tiles round-robin (tile = part, part + parts, ...): together they walk the tiles in index
  // order, which is the order the data-gradient kernel produces them in when the two kernels overlap.
  const int my_job = p.cta_job[blockIdx.x], parts = p.job_parts[my_job];
  const int t_begin = (int)p.cta_part[blockIdx.x], t_end = num_tiles, t_step = parts;
  const int my_tiles = t_begin < num_tiles ? (num_tiles - t_begin + parts - 1) / parts : 0;
  const long long trace_t0 = clock64();

  if (warp == 0) {
    // ===== producer: per (job, tile, sample half) one stage: a_blocks + z_blocks half blocks of 8 KB =====
    if (lane == 0 && my_tiles > 0) {
      uint32_t stage = 0, phase = 0;
      const bool trp = DURF_TRACE_DETAIL && p.trace_detail && my_job == 1 && p.cta_part[blockIdx.x] == 0;
      long long w_flag = 0, w_empty = 0, tq = 0, t_all = clock64();
      for (int j = my_job; j == my_job; ++j) {
        const WgradJob jb = p.jobs[j];
        const uint8_t* a_base = jb.a_src ? p.feat : p.saved;
        const size_t a_stride = (size_t)(jb.a_src ? 1 : p.saved_blocks) * kBlockBytes;
        for (int tile = t_begin; tile < t_end; tile += t_step)
          for (int half = 0; half < 2; ++half) {
            if (trp) tq = clock64();
            if (half == 0 && p.flags && jb.flag_idx >= 0) {
              // overlap with dgrad: wait until this tile's dZ blocks are complete (acquire), then order the generic-proxy
              // acquire before the async-proxy reads of the bulk copies below.  A protocol bug traps instead of hanging.
              const int32_t* f = p.flags + (size_t)tile * p.flag_stride + jb.flag_idx;
              int v;
              long long t0 = 0;
              while (true) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if (v >= jb.flag_need) break;
                __nanosleep(200);
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000LL) {
                  printf("durf wgrad kernel: tile %d of job %d never completed (counter %d of %d)\n", tile, my_job, v, jb.flag_need);
                  __trap();
                }
              }
              asm volatile("fence.proxy.async.global;" ::: "memory");
            }
            if (trp) { w_flag += clock64() - tq; tq = clock64(); }
            mbar_wait(bar_empty(stage), phase ^ 1);
            if (trp) w_empty += clock64() - tq;
            mbar_arrive_expect_tx(bar_full(stage), (jb.a_blocks + jb.z_blocks) * kHalfBlock);
            const uint32_t dst = sbase + stage * kWgStageBytes;
            // layer records are [sample half][64-column block][64 rows x 128 B]: the 64 samples of all blocks of a layer are ONE
            // contiguous piece of up to 32 KB (8 KB pieces stream at a third of the per-SM rate of 32 KB ones:
            // profiles/r01_ubench_stream.log); an input-feature tile is a single 128-row block image
            if (jb.a_src)
              bulk_g2s(dst, a_base + (size_t)tile * a_stride + half * kHalfBlock, kHalfBlock, bar_full(stage));
            else
              bulk_g2s(dst, a_base + (size_t)tile * a_stride + (size_t)jb.a_off * kBlockBytes + (size_t)half * jb.a_blocks * kHalfBlock,
                       jb.a_blocks * kHalfBlock, bar_full(stage));
            if (jb.z_blocks > 0)
              bulk_g2s(dst + 4 * kHalfBlock,
                       p.dz + ((size_t)tile * p.dz_blocks + jb.z_off) * kBlockBytes + (size_t)half * jb.z_blocks * kHalfBlock,
                       jb.z_blocks * kHalfBlock, bar_full(stage));
            if (++stage == kWgStages) { stage = 0; phase ^= 1; }
          }
      }
      if (trp) printf("durf wgrad trace: producer of job 1 part 0: %d tiles, total %lld cyc; waiting for tile flags %lld, for empty stages %lld\n",
                      my_tiles, clock64() - t_all, w_flag, w_empty);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer: D[features of A, columns of dZ] += A^T dZ over the 64 samples of a stage (4 x K=16 per M tile) =====
    if (my_tiles > 0) {      // the whole warp walks the schedule (uniform control flow), one elected lane issues
      constexpr uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
      uint32_t stage = 0, phase = 0;
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const WgradJob jb = p.jobs[my_job];
      // one job per CTA: the accumulator is never recycled, no wait on bar_acc_free.  The wait for a stage's operands is
      // software-pipelined (umma_stage_mn): a blocking wait between two groups of MMAs would drain the tensor queue, which
      // capped a CTA at ~30 B/cycle of operands - invisible while 148 CTAs saturate HBM, fatal when the kernel shares the GPU
      // with the data-gradient kernel and runs on a third of the SMs.
      const uint32_t m2 = jb.a_blocks > 2 ? 1u : 0u;
      const uint32_t a_lbo = jb.a_blocks >= 2 ? (uint32_t)(kHalfBlock >> 4) : 0u;    // a single block is read twice (rows 64..127 unused)
      const uint32_t N = jb.z_blocks * 64;
      const uint32_t idesc = umma_idesc_mn(128, (int)N);
      mbar_wait(bar_full(0), 0);
      const bool trm = DURF_TRACE_DETAIL && p.trace_detail && my_job == 1 && p.cta_part[blockIdx.x] == 0;
      const long long m_all = clock64();
      for (int tile = t_begin; tile < t_end; tile += t_step)
        for (int half = 0; half < 2; ++half) {
          tc_fence_after();
          const bool last = half == 1 && tile + t_step >= t_end;
          const uint32_t next_stage = stage + 1 == kWgStages ? 0 : stage + 1;
          const uint32_t next_phase = stage + 1 == kWgStages ? phase ^ 1 : phase;
          if (jb.z_blocks > 0) {
            const uint32_t st = sbase + stage * kWgStageBytes;
            const uint32_t b_lo = (((st + 4 * kHalfBlock) & 0x3FFFF) >> 4) | ((uint32_t)(kHalfBlock >> 4) << 16);
            const uint32_t a0_lo = ((st & 0x3FFFF) >> 4) | (a_lbo << 16);
            const uint32_t a1_lo = (((st + 2 * kHalfBlock) & 0x3FFFF) >> 4) | (a_lbo << 16);
            umma_stage_mn(tmem_u, tmem_u + N, a0_lo, a1_lo, b_lo, desc_hi, idesc, (tile == t_begin && half == 0) ? 0u : 1u, m2,
                          bar_empty(stage), bar_full(next_stage), next_phase, last ? 0u : 1u);
          } else {
            tc_commit_conv(bar_empty(stage));
            if (!last) mbar_wait(bar_full(next_stage), next_phase);
          }
          stage = next_stage; phase = next_phase;
        }
      if (jb.z_blocks > 0) tc_commit_conv(bar_acc_done);
      if (trm && lane == 0) printf("durf wgrad trace: MMA warp of job 1 part 0: %d stages in %lld cyc\n", 2 * my_tiles, clock64() - m_all);
    }
  } else if (warp >= 4) {
    // ===== auxiliary warps: bias / narrow-head gradients from the stages in shared memory, then the accumulator read-out.
    // Thread (rg, ck) of the 256 owns the 16-byte piece ck (8 consecutive columns) of the rows rg, rg+8, ... of a stage, so
    // a stage costs it 8 ld.shared.v4 per operand; the 8 row groups of a column meet in the final reductions.
    const int t = threadIdx.x - 128;
    const int q = warp & 3, ch = (warp - 4) >> 2;
    const int ck = t & 31, rg = t >> 5;           // piece (0..31: block ck>>3, 16-byte chunk ck&7) and row group (0..7)
    float* s_cs = reinterpret_cast<float*>(smem + OFF_MISC + 128);     // [8][128] per-tile column sums (view part)
    uint32_t stage = 0, phase = 0, done_par = 0;
    if (my_tiles > 0)
      for (int j = my_job; j == my_job; ++j) {
        const WgradJob jb = p.jobs[j];
        float bsum[8], ex[8][3], exb[3] = {0.f, 0.f, 0.f};
        float cv[32];                              // view part: d kernel[width + i][t], threads t < 128
#pragma unroll
        for (int e = 0; e < 8; ++e) { bsum[e] = 0.f; ex[e][0] = ex[e][1] = ex[e][2] = 0.f; }
#pragma unroll
        for (int i = 0; i < 32; ++i) cv[i] = 0.f;
        const int nvec = jb.extra == 1 ? 1 : (jb.extra == 2 ? 3 : 0);
        const bool do_bias = (jb.db_off >= 0 || jb.extra == 3) && ck < jb.z_blocks * 8;
        const bool do_ex = nvec > 0 && ck < jb.a_blocks * 8;
        const bool tra = DURF_TRACE_DETAIL && p.trace_detail && my_job == 1 && p.cta_part[blockIdx.x] == 0 && t == 0;
        long long a_wait = 0, a_work = 0, aq = 0;
        for (int tile = t_begin; tile < t_end; tile += t_step) {
          const int ray = p.ray_index ? p.ray_index[tile] : tile;
          float tsum[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) tsum[e] = 0.f;
          for (int half = 0; half < 2; ++half) {
            if (tra) aq = clock64();
            mbar_wait(bar_full(stage), phase);
            if (tra) { a_wait += clock64() - aq; aq = clock64(); }
            const uint32_t st = sbase + stage * kWgStageBytes;
            if (do_bias) {
              const uint32_t zb = st + (4 + (ck >> 3)) * kHalfBlock;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = rg + 8 * i;          // r & 7 == rg
                const float4 w = lds128_volatile(zb + (r >> 3) * 1024 + rg * 128 + (((ck & 7) ^ rg) << 4));
                const uint32_t u[4] = {__float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(w.z), __float_as_uint(w.w)};
#pragma unroll
                for (int e = 0; e < 4; ++e) { tsum[2 * e] += __uint_as_float(u[e] << 16); tsum[2 * e + 1] += __uint_as_float(u[e] & 0xFFFF0000u); }
              }
            }
            if (do_ex) {
              const uint32_t ab = st + (ck >> 3) * kHalfBlock;
              const float* vec = (jb.extra == 1 ? p.d_raw_density + (size_t)ray * kTileM : p.d_raw_rgb + (size_t)ray * kTileM * 3) +
                                 (size_t)half * 64 * nvec;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = rg + 8 * i;
                const float4 w = lds128_volatile(ab + (r >> 3) * 1024 + rg * 128 + (((ck & 7) ^ rg) << 4));
                const uint32_t u[4] = {__float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(w.z), __float_as_uint(w.w)};
                float g[3];
                for (int jv = 0; jv < 3; ++jv) g[jv] = jv < nvec ? __ldg(vec + r * nvec + jv) : 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float a0 = __uint_as_float(u[e] << 16), a1 = __uint_as_float(u[e] & 0xFFFF0000u);
#pragma unroll
                  for (int jv = 0; jv < 3; ++jv) { ex[2 * e][jv] = fmaf(a0, g[jv], ex[2 * e][jv]); ex[2 * e + 1][jv] = fmaf(a1, g[jv], ex[2 * e + 1][jv]); }
                }
                if (ck == 0)
                  for (int jv = 0; jv < 3; ++jv) exb[jv] += g[jv];
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty(stage));
            if (tra) a_work += clock64() - aq;
            if (++stage == kWgStages) { stage = 0; phase ^= 1; }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) bsum[e] += tsum[e];
          if (jb.extra == 3) {
            // view part: d kernel[width + i][c] += enc_i(ray) * (column sum of dZ_cond over this tile's 128 samples)
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (ck < 16) {
#pragma unroll
              for (int e = 0; e < 8; ++e) s_cs[rg * 128 + ck * 8 + e] = tsum[e];
            }
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (t < 128) {
              float tot = 0.f;
#pragma unroll
              for (int g8 = 0; g8 < 8; ++g8) tot += s_cs[g8 * 128 + t];
              const float* ve = p.cond + (size_t)ray * p.cond_dim;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < p.cond_dim) cv[i] = fmaf(__ldg(ve + i), tot, cv[i]);
            }
          }
        }
        if (tra) printf("durf wgrad trace: aux thread of job 1 part 0: waiting for full stages %lld cyc, column sums %lld cyc\n", a_wait, a_work);
        // per-thread partial sums -> global gradient (8 row groups x 148 CTAs add into every address)
        if (jb.db_off >= 0 && do_bias) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (ck * 8 + e < jb.n_valid) red_add(p.d_params + jb.db_off + ck * 8 + e, bsum[e]);
        }
        if (do_ex) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            for (int jv = 0; jv < nvec; ++jv) red_add(p.d_params + jb.ex_off + (ck * 8 + e) * nvec + jv, ex[e][jv]);
          if (ck == 0 && jb.ex_b_off >= 0)
            for (int jv = 0; jv < nvec; ++jv) red_add(p.d_params + jb.ex_b_off + jv, exb[jv]);
        }
        if (jb.extra == 3 && t < 128) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < p.cond_dim) red_add(p.d_params + jb.ex_off + i * jb.ld + t, cv[i]);
        }
        // accumulator read-out: lane = input feature, columns = output features
        if (jb.z_blocks > 0) {
          mbar_wait(bar_acc_done, done_par); done_par ^= 1;
          tc_fence_after();
          const int m_tiles = jb.a_blocks > 2 ? 2 : 1;
          const int N = jb.z_blocks * 64;
          for (int mt = 0; mt < m_tiles; ++mt) {
            const int m = mt * 128 + q * 32 + lane;               // input feature (row of the kernel)
            for (int c0 = ch * (N / 2); c0 < (ch + 1) * (N / 2); c0 += 32) {
              uint32_t v[32];
              tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + mt * N + c0, v);
              tmem_ld_wait();
              tmem_ld_pin(v);
              if (m < jb.k_valid) {
                float* dst = p.d_params + jb.dw_off + (size_t)m * jb.ld + c0;
                // ld is a multiple of 4, so every row of the kernel has the same 16-byte phase (the blob offsets after the
                // 257-float density head are odd): `lead` scalar reductions, then 16-byte ones, then the scalar tail
                const int lead = (int)((4u - ((uint32_t)(reinterpret_cast<uintptr_t>(dst) >> 2) & 3u)) & 3u);
                const int nv = min(32, jb.n_valid - c0);
                int e = 0;
                for (; e < lead && e < nv; ++e) red_add(dst + e, __uint_as_float(v[e]));
                if (lead == 0) {
#pragma unroll
                  for (int q4 = 0; q4 < 8; ++q4)
                    if (4 * q4 + 4 <= nv) red_add4(dst + 4 * q4, __uint_as_float(v[4 * q4]), __uint_as_float(v[4 * q4 + 1]),
                                                   __uint_as_float(v[4 * q4 + 2]), __uint_as_float(v[4 * q4 + 3]));
                  e = nv & ~3;
                } else if (lead == 1) {
#pragma unroll
                  for (int q4 = 0; q4 < 7; ++q4)
                    if (1 + 4 * q4 + 4 <= nv) red_add4(dst + 1 + 4 * q4, __uint_as_float(v[1 + 4 * q4]), __uint_as_float(v[2 + 4 * q4]),
                                                       __uint_as_float(v[3 + 4 * q4]), __uint_as_float(v[4 + 4 * q4]));
                  e = nv >= 1 ? 1 + ((nv - 1) & ~3) : nv;
                } else if (lead == 2) {
#pragma unroll
                  for (int q4 = 0; q4 < 7; ++q4)
                    if (2 + 4 * q4 + 4 <= nv) red_add4(dst + 2 + 4 * q4, __uint_as_float(v[2 + 4 * q4]), __uint_as_float(v[3 + 4 * q4]),
                                                       __uint_as_float(v[4 + 4 * q4]), __uint_as_float(v[5 + 4 * q4]));
                  e = nv >= 2 ? 2 + ((nv - 2) & ~3) : nv;
                } else {
#pragma unroll
                  for (int q4 = 0; q4 < 7; ++q4)
                    if (3 + 4 * q4 + 4 <= nv) red_add4(dst + 3 + 4 * q4, __uint_as_float(v[3 + 4 * q4]), __uint_as_float(v[4 + 4 * q4]),
                                                       __uint_as_float(v[5 + 4 * q4]), __uint_as_float(v[6 + 4 * q4]));
                  e = nv >= 3 ? 3 + ((nv - 3) & ~3) : nv;
                }
                for (; e < nv; ++e) red_add(dst + e, __uint_as_float(v[e]));
              }
            }
          }
          tc_fence_before();
          mbar_arrive(bar_acc_free);
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (p.trace && threadIdx.x == 0) p.trace[blockIdx.x] = clock64() - trace_t0;
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 128 + 8 * 128 * 4 + 1024;

// Builds the job list for a topology.  Tile records: saved activations = [layer g][W/64 blocks] for the depth+1 trunk /
// bottleneck layers, then 2 blocks of the condition layer's activation; dz has the same record shape (dz of layer g at
// the slot of its output activation).
int mlp_tc_wgrad_launch(cudaStream_t st, const DurfMlpTopology& t, int saved_blocks, const WgradParams& base, int max_ctas) {
  MlpLayout L(t);
  WgradParams P = base;
  const int KB = t.width / 64, G = t.depth + 2;
  int nj = 0;
  auto slot = [&](int g) { return g * KB; };       // block offset of layer g's output inside a tile record
  for (int g = 0; g < G; ++g) {
    const int layer = (g < t.depth) ? g : (g == t.depth ? t.depth + 1 : t.depth + 2);
    const bool skip_in = (g >= 1) && (g < t.depth) && ((g - 1) % t.skip == 0) && (g - 1 > 0);
    const int n_out = L.out_dim[layer];
    const int zb = n_out / 64;
    WgradJob jb{};
    jb.z_off = slot(g); jb.z_blocks = zb; jb.ld = n_out; jb.n_valid = n_out; jb.db_off = (int)L.b_off[layer];
    jb.extra = 0; jb.ex_off = 0; jb.ex_b_off = -1;
    // dZ of layer g: one increment per epilogue warp (8) and N-half of the data-gradient kernel; dZ_cond: one per warp
    jb.flag_idx = g; jb.flag_need = (g == G - 1) ? 8 : 8 * (t.width / 128);
    if (g == 0) { jb.a_src = 1; jb.a_off = 0; jb.a_blocks = 1; jb.dw_off = (int)L.w_off[layer]; jb.k_valid = t.in_dim; }
    else {
      // input of trunk layer g / bottleneck = activation g-1 (bottleneck: the last trunk activation); condition layer = bottleneck output
      const int src = (g == t.depth) ? t.depth - 1 : g - 1;
      jb.a_src = 0; jb.a_off = slot(src); jb.a_blocks = KB; jb.dw_off = (int)L.w_off[layer]; jb.k_valid = t.width;
    }
    if (g == t.depth) {       // bottleneck job also carries the density head: both read the last trunk activation
      jb.extra = 1; jb.ex_off = (int)L.w_off[t.depth]; jb.ex_b_off = (int)L.b_off[t.depth];
    }
    if (g == t.depth + 1) {   // condition layer: the view inputs' rows follow the `width` bottleneck rows
      jb.extra = 3; jb.ex_off = (int)L.w_off[layer] + t.width * t.cond_width;
    }
    P.jobs[nj++] = jb;
    if (skip_in) {            // second K part of the skip layer: the re-concatenated input features (rows width .. width+in_dim)
      WgradJob js = jb;
      js.a_src = 1; js.a_off = 0; js.a_blocks = 1; js.dw_off = (int)L.w_off[layer] + t.width * n_out; js.k_valid = t.in_dim;
      js.db_off = -1; js.extra = 0;
      P.jobs[nj++] = js;
    }
  }
  {                           // rgb head: A = condition-layer activation, vec = d_raw_rgb; no tensor-core work
    WgradJob jr{};
    jr.a_src = 0; jr.a_off = slot(G - 1); jr.a_blocks = t.cond_width / 64; jr.z_blocks = 0; jr.db_off = -1;
    jr.extra = 2; jr.ex_off = (int)L.w_off[t.depth + 3]; jr.ex_b_off = (int)L.b_off[t.depth + 3];
    jr.flag_idx = -1; jr.flag_need = 0;
    P.jobs[nj++] = jr;
  }
  DURF_REQUIRE(nj <= kMaxJobs, DURF_E_UNSUPPORTED, "durf_mlp_bwd(bf16): too many wgrad jobs (%d)", nj);
  P.n_jobs = nj;
  P.saved_blocks = saved_blocks; P.dz_blocks = saved_blocks;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  sms = sms > kMaxCtas ? kMaxCtas : sms;
  if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;      // the rest of the GPU runs the data-gradient kernel
  // CTAs per job in proportion to the job's cost per tile, at least one each, never more CTAs than tiles.  Cost model
  // (cycles per tile, measured with DURF_WGRAD_TRACE on 16,384 tiles): 1500 (two ring stages' worth of latency) + 15 per KB
  // streamed + the auxiliary warps' work for the narrow heads (density head 2000, rgb head 2560, view part 1160).
  {
    int bytes[kMaxJobs], total = 0, parts[kMaxJobs], used = 0;
    static const int kExtraCost[4] = {0, 2000, 2560, 1160};
    for (int j = 0; j < nj; ++j) {
      bytes[j] = 1500 + 15 * 16 * (P.jobs[j].a_blocks + P.jobs[j].z_blocks) + kExtraCost[P.jobs[j].extra & 3];
      total += bytes[j];
    }
    const int budget = sms < nj ? nj : sms;
    const int max_parts = P.M < 1 ? 1 : (P.M > 255 ? 255 : P.M);
    for (int j = 0; j < nj; ++j) {
      int q = (int)((long long)bytes[j] * budget / total);
      q = q < 1 ? 1 : (q > max_parts ? max_parts : q);
      parts[j] = q; used += q;
    }
    for (bool grew = true; used < budget && grew;) {        // hand the remaining SMs to the jobs with the most bytes per CTA
      grew = false;
      int best = -1;
      for (int j = 0; j < nj; ++j)
        if (parts[j] < max_parts && (best < 0 || bytes[j] * parts[best] > bytes[best] * parts[j])) best = j;
      if (best >= 0) { ++parts[best]; ++used; grew = true; }
    }
    DURF_REQUIRE(used <= kMaxCtas, DURF_E_UNSUPPORTED, "durf_mlp_bwd(bf16): wgrad needs %d CTAs", used);
    int b = 0;
    for (int j = 0; j < nj; ++j) {
      P.job_parts[j] = (uint8_t)parts[j];
      for (int q = 0; q < parts[j]; ++q, ++b) { P.cta_job[b] = (uint8_t)j; P.cta_part[b] = (uint8_t)q; }
    }
    // interleave would spread a job over both dies; the order only matters for L2 locality of A / dZ, which stream once
  }
  int grid = 0;
  for (int j = 0; j < nj; ++j) grid += P.job_parts[j];
  cudaError_t e = cudaFuncSetAttribute(mlp_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes);
  DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_bwd(bf16): smem attribute: %s", cudaGetErrorString(e));
  static long long* trace_buf = nullptr;
  const bool tracing = getenv("DURF_WGRAD_TRACE") != nullptr && atoi(getenv("DURF_WGRAD_TRACE")) == 1;
  P.trace_detail = (getenv("DURF_WGRAD_TRACE") != nullptr && atoi(getenv("DURF_WGRAD_TRACE")) == 2) ? 1 : 0;
  if (tracing && !trace_buf) cudaMalloc(&trace_buf, kMaxCtas * sizeof(long long));
  P.trace = tracing ? trace_buf : nullptr;
  if (P.flags && (grid & 1) == 0) {
    // sharing the GPU with the data-gradient kernel: CTA pairs, so that the two SMs of a TPC run the SAME kernel
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = kWgSmemBytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, mlp_tc_wgrad_kernel, P);
    DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_bwd(bf16): wgrad launch: %s", cudaGetErrorString(e));
  } else
  mlp_tc_wgrad_kernel<<<grid, 384, kWgSmemBytes, st>>>(P);
  DURF_CHECK_LAUNCH("durf_mlp_bwd(bf16): wgrad");
  if (tracing) {                                             // debugging aid only: synchronises
    long long h[kMaxCtas];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, trace_buf, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    fprintf(stderr, "wgrad trace: M=%d width=%d grid=%d\n", P.M, t.width, grid);
    for (int j = 0, b = 0; j < nj; ++j) {
      long long mx = 0;
      for (int q = 0; q < P.job_parts[j]; ++q, ++b) mx = h[b] > mx ? h[b] : mx;
      fprintf(stderr, "  job %2d a=%d z=%d extra=%d parts=%3d max_cycles=%lld\n", j, P.jobs[j].a_blocks, P.jobs[j].z_blocks, P.jobs[j].extra,
              P.job_parts[j], mx);
    }
  }
  return DURF_OK;
}

}  // namespace durf

namespace durf {

int mlp_tc_saved_blocks(const DurfMlpTopology& t);

int mlp_tc_wgrad_run(cudaStream_t st, const DurfMlpTopology& t, const uint8_t* saved, const uint8_t* feat, const uint8_t* dz,
                     const float* d_raw_rgb, const float* d_raw_density, const float* cond, const int32_t* ray_index,
                     const int32_t* count, int M, float* d_params, const int32_t* tile_done, int max_ctas) {
  WgradParams P{};
  P.saved = saved; P.feat = feat; P.dz = dz; P.d_raw_rgb = d_raw_rgb; P.d_raw_density = d_raw_density; P.cond = cond;
  P.ray_index = ray_index; P.count = count; P.M = M; P.cond_dim = t.cond_dim; P.d_params = d_params;
  P.flags = tile_done; P.flag_stride = t.depth + 2;
  return mlp_tc_wgrad_launch(st, t, mlp_tc_saved_blocks(t), P, max_ctas);
}

}  // namespace durf
