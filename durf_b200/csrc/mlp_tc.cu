// K2 -- the radiance/density MLP as a fused tcgen05 GEMM chain (sm_100a), forward.
// Replaces MLP.__call__ / BoxMLP.__call__ (obbpose_model.py:305-354, 369-418) for width 256 / 128, cond_width 128.
//
// One persistent CTA per SM works on PAIRS of 128-row tiles (one tile = the 128 samples of one ray-level), so
// every 16 KB weight chunk pulled from L2 feeds 256 rows.  Per layer:
//   TMA producer (warp 0)  : streams the pre-tiled, pre-swizzled bf16 weight image chunk by chunk
//                            (cp.async.bulk -> mbarrier) through a ring of shared-memory stages;
//   MMA issuer  (warp 1)   : one thread issues tcgen05.mma (M=128, N=128, K=16, bf16 x bf16 -> fp32) with the
//                            tile's activations (A, K-major SWIZZLE_128B in shared memory) against the staged
//                            weight chunk (B); accumulators live in TMEM (128 lanes x width columns per tile);
//   feature loader (warp 2): bulk-copies the bf16 feature tile image written by the ray-march kernel;
//   epilogue (warps 4-7 for tile 0, 8-11 for tile 1): tcgen05.ld the accumulators, + bias, ReLU, bf16 pack and
//                            write the next layer's A operand in place; the last trunk layer also forms the density
//                            head, the condition layer's epilogue adds the per-ray view term as a bias, forms the
//                            rgb head and writes raw_rgb / raw_density.
// The skip connection (obbpose_model.py:332-333) is a fifth K block taken from the still-resident input tile;
// the view direction (constant along a ray) enters the condition layer as a per-tile fp32 bias
// (b + W_view^T enc), so its 27 input columns never occupy tensor-core K.
#include "common.cuh"
#include "mlp_topology.h"

namespace durf {

constexpr int kTileM = 128;
constexpr int kChunkBytes = 16384;    // 128 (n) x 64 (k) bf16, K-major SWIZZLE_128B
constexpr int kMaxG = 16;

struct LayerSched {
  int n_halves;     // output columns / 128
  int n_act_kb;     // 64-wide K blocks taken from the activation tile
  int uses_inp;     // +1 K block from the input-feature tile (layer 0, skip layer)
  int kind;         // 0 relu->act, 1 relu->act + density head, 2 linear->act (bottleneck), 3 condition + rgb head + output
  int bias_off;     // offset (floats) of this layer's bias inside the parameter blob
  int last_inp_use; // 1 if no later layer of the tile reads the input-feature tile
};

struct TcParams {
  const uint8_t* feat;       // bf16 tile images, 16 KB per tile
  const float* cond;         // [B, cond_dim]
  const float* params;       // fp32 blob
  const uint8_t* packed;     // weight image
  const int32_t* ray_index;
  const int32_t* count;
  int M;
  int accumulate;
  float* raw_rgb;
  float* raw_density;
  int G;                     // GEMM layers per tile: depth + 2
  int depth;
  int cond_dim;
  int chunks_per_pair;
  int off_wden, off_bden, off_wrgb, off_brgb, off_wview;
  LayerSched sched[kMaxG];
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spin on try_wait (which itself blocks for a bounded hardware interval).  A watchdog turns a protocol bug into a
// trapped launch (reported by the next CUDA call) instead of a hung GPU: no legitimate wait lasts 4e9 cycles.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 4000000000LL) {
      printf("durf mlp_tc: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: start>>4 | LBO(=1, unused for swizzled K-major)<<16 |
// SBO (1024 B between 8-row groups)>>4 <<32 | version 1 <<46 | layout 2 <<61   (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10), K-major both,
// N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int W>
struct TcCfg {
  static constexpr int KB = W / 64;                       // K blocks of the activation tile
  static constexpr int STAGES = (W == 256) ? 3 : 4;
  static constexpr int ACT_BYTES = kTileM * W * 2;
  static constexpr int TMEM_COLS = 2 * W;                 // two tiles; power of two >= 32
  // shared memory map (bytes, from a 1024-aligned base)
  static constexpr int OFF_ACT = 0;
  static constexpr int OFF_INP = OFF_ACT + 2 * ACT_BYTES;
  static constexpr int OFF_WST = OFF_INP + 2 * kChunkBytes;
  static constexpr int OFF_BIAS = OFF_WST + STAGES * kChunkBytes;     // fp32 [kMaxBiasLayers][W] trunk + bottleneck
  static constexpr int N_BIAS = 13 * W;                               // depth <= 12 trunk layers + bottleneck
  static constexpr int OFF_WDEN = OFF_BIAS + N_BIAS * 4;              // fp32 [W]
  static constexpr int OFF_WRGB = OFF_WDEN + W * 4;                   // fp32 [3][128]
  static constexpr int OFF_VBIAS = OFF_WRGB + 3 * 128 * 4;            // fp32 [2][128]
  static constexpr int OFF_MISC = OFF_VBIAS + 2 * 128 * 4;            // head biases [4] + tmem ptr + barriers
  static constexpr int MISC_BYTES = 256;
  static constexpr int SMEM_BYTES = OFF_MISC + MISC_BYTES + 1024;     // + alignment slack
};

template <int W>
__global__ void __launch_bounds__(384, 1)
mlp_tc_fwd_kernel(const TcParams p) {
  using C = TcCfg<W>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  float* s_bias = reinterpret_cast<float*>(smem + C::OFF_BIAS);
  float* s_wden = reinterpret_cast<float*>(smem + C::OFF_WDEN);
  float* s_wrgb = reinterpret_cast<float*>(smem + C::OFF_WRGB);
  float* s_vbias = reinterpret_cast<float*>(smem + C::OFF_VBIAS);
  float* s_hb = reinterpret_cast<float*>(smem + C::OFF_MISC);             // [0]=b_den, [1..3]=b_rgb
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + C::OFF_MISC + 16);
  const uint32_t bar0 = sbase + C::OFF_MISC + 32;
  // barrier map (8 bytes each)
  auto bar_full = [&](int s) { return bar0 + 8 * s; };
  auto bar_empty = [&](int s) { return bar0 + 8 * (4 + s); };
  auto bar_inp_full = [&](int t) { return bar0 + 8 * (8 + t); };
  auto bar_inp_empty = [&](int t) { return bar0 + 8 * (10 + t); };
  auto bar_acc_full = [&](int t) { return bar0 + 8 * (12 + t); };
  auto bar_act_ready = [&](int t) { return bar0 + 8 * (14 + t); };

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar_inp_full(t), 1); mbar_init(bar_inp_empty(t), 1);
      mbar_init(bar_acc_full(t), 1); mbar_init(bar_act_ready(t), 128);
    }
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int num_tiles = p.count ? min(*p.count, p.M) : p.M;
  const int num_pairs = (num_tiles + 1) / 2;

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
        for (int c = 0; c < p.chunks_per_pair; ++c) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          mbar_arrive_expect_tx(bar_full(stage), kChunkBytes);
          bulk_g2s(sbase + C::OFF_WST + stage * kChunkBytes, p.packed + (size_t)c * kChunkBytes, kChunkBytes, bar_full(stage));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(128, 128);
      uint32_t stage = 0, phase = 0;
      uint32_t act_par[2] = {1, 1}, inp_par[2] = {0, 0};
      for (int pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
        const bool active[2] = {true, 2 * pair + 1 < num_tiles};
        for (int g = 0; g < p.G; ++g) {
          const LayerSched& L = p.sched[g];
          const int nkb = L.n_act_kb + L.uses_inp;
          for (int nh = 0; nh < L.n_halves; ++nh) {
            for (int kc = 0; kc < nkb; ++kc) {
              mbar_wait(bar_full(stage), phase);
              tc_fence_after();
              const uint32_t b_addr = sbase + C::OFF_WST + stage * kChunkBytes;
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                if (!active[t]) continue;
                if (nh == 0 && kc == 0) {
                  mbar_wait(bar_act_ready(t), act_par[t]); act_par[t] ^= 1;   // A operand written, accumulators drained
                  if (g == 0) { mbar_wait(bar_inp_full(t), inp_par[t]); inp_par[t] ^= 1; }
                  tc_fence_after();
                }
                const uint32_t a_addr = (kc < L.n_act_kb) ? sbase + C::OFF_ACT + t * C::ACT_BYTES + kc * kChunkBytes
                                                          : sbase + C::OFF_INP + t * kChunkBytes;
                const uint32_t d_addr = tmem_base + t * W + nh * 128;
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16)
                  umma_bf16(d_addr, umma_desc_sw128(a_addr + k16 * 32), umma_desc_sw128(b_addr + k16 * 32), idesc,
                            (kc > 0 || k16 > 0) ? 1u : 0u);
              }
              tc_commit(bar_empty(stage));                                     // frees the weight stage when the MMAs retire
              if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
          }
          for (int t = 0; t < 2; ++t) {
            if (!active[t]) continue;
            if (L.last_inp_use) tc_commit(bar_inp_empty(t));
            tc_commit(bar_acc_full(t));
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===== feature-tile loader =====
    if (lane == 0) {
      uint32_t par[2] = {0, 0};
      for (int pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
        for (int t = 0; t < 2; ++t) {
          const int tile = 2 * pair + t;
          if (tile >= num_tiles) continue;
          mbar_wait(bar_inp_empty(t), par[t] ^ 1);
          mbar_arrive_expect_tx(bar_inp_full(t), kChunkBytes);
          bulk_g2s(sbase + C::OFF_INP + t * kChunkBytes, p.feat + (size_t)tile * kChunkBytes, kChunkBytes, bar_inp_full(t));
          par[t] ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===== epilogue: warps 4-7 own tile 0, warps 8-11 own tile 1; thread = accumulator row = sample =====
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int tid_in_tile = row;
    uint8_t* act = smem + C::OFF_ACT + t * C::ACT_BYTES;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + t * W;
    uint32_t acc_par = 0;
    for (int pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
      const int tile = 2 * pair + t;
      if (tile >= num_tiles) break;
      const int ray = p.ray_index ? p.ray_index[tile] : tile;
      // per-ray bias of the condition layer: b + W_view^T enc(viewdir)   (obbpose_model.py:343-350)
      {
        const LayerSched& Lc = p.sched[p.G - 1];
        asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");   // previous tile's readers of s_vbias are done
        float vb = p.params[Lc.bias_off + tid_in_tile];
        const float* cv = p.cond + (size_t)ray * p.cond_dim;
        for (int i = 0; i < p.cond_dim; ++i) vb = fmaf(cv[i], p.params[p.off_wview + i * 128 + tid_in_tile], vb);
        s_vbias[t * 128 + tid_in_tile] = vb;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + t) : "memory");
      }
      float den = 0.f;
      for (int g = 0; g < p.G; ++g) {
        const LayerSched& L = p.sched[g];
        mbar_wait(bar_acc_full(t), acc_par); acc_par ^= 1;
        tc_fence_after();
        if (L.kind != 3) {
          const float* sb = s_bias + g * W;
#pragma unroll 1
          for (int cb = 0; cb < W / 32; ++cb) {
            uint32_t v[32];
            tmem_ld32(t_lane + cb * 32, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t wv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int col = cb * 32 + j * 8 + 2 * e;
                float a0 = __uint_as_float(v[j * 8 + 2 * e]) + sb[col];
                float a1 = __uint_as_float(v[j * 8 + 2 * e + 1]) + sb[col + 1];
                if (L.kind != 2) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
                if (L.kind == 1) { den = fmaf(a0, s_wden[col], den); den = fmaf(a1, s_wden[col + 1], den); }
                wv[e] = pack_bf16x2(a0, a1);
              }
              const int col0 = cb * 32 + j * 8;
              *reinterpret_cast<uint4*>(act + (col0 >> 6) * kChunkBytes + sw128_offset(row, (col0 & 63) >> 3)) =
                  make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
          }
          tc_fence_before();
          fence_async_smem();            // generic-proxy stores -> visible to the tensor core's async proxy
          mbar_arrive(bar_act_ready(t));
        } else {
          // condition layer (128 columns): + per-ray bias, ReLU, rgb head, write raw outputs
          float rgb[3] = {0.f, 0.f, 0.f};
          const float* vb = s_vbias + t * 128;
#pragma unroll 1
          for (int cb = 0; cb < 4; ++cb) {
            uint32_t v[32];
            tmem_ld32(t_lane + cb * 32, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int col = cb * 32 + e;
              const float a = fmaxf(__uint_as_float(v[e]) + vb[col], 0.f);
              rgb[0] = fmaf(a, s_wrgb[col], rgb[0]);
              rgb[1] = fmaf(a, s_wrgb[128 + col], rgb[1]);
              rgb[2] = fmaf(a, s_wrgb[256 + col], rgb[2]);
            }
          }
          tc_fence_before();
          mbar_arrive(bar_act_ready(t));   // tile finished: accumulators drained, activation buffer free
          const size_t o = (size_t)ray * kTileM + row;
          const float dv = den + s_hb[0];
          if (p.accumulate) {
            p.raw_density[o] += dv;
            p.raw_rgb[o * 3 + 0] += rgb[0] + s_hb[1];
            p.raw_rgb[o * 3 + 1] += rgb[1] + s_hb[2];
            p.raw_rgb[o * 3 + 2] += rgb[2] + s_hb[3];
          } else {
            p.raw_density[o] = dv;
            p.raw_rgb[o * 3 + 0] = rgb[0] + s_hb[1];
            p.raw_rgb[o * 3 + 1] = rgb[1] + s_hb[2];
            p.raw_rgb[o * 3 + 2] = rgb[2] + s_hb[3];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS));
  }
}

// fp32 parameter blob -> bf16 weight image: for every GEMM layer, N half and K block (the order the kernel
// consumes them) one 128x64 K-major SWIZZLE_128B chunk.  One thread per 16-byte piece.
struct PackParams {
  const float* params;
  uint8_t* packed;
  int n_chunks;
  int W, in_dim;
  // per chunk: source kernel offset, its leading dimension (= out dim), first source row, rows available, first column
  int w_off[160], ld[160], k_first[160], k_avail[160], n_first[160], n_avail[160];
};
__global__ void pack_weights_kernel(const __grid_constant__ PackParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)p.n_chunks * 128 * 8) return;
  const int chunk = (int)(i / 1024), r = (int)(i % 1024) / 8, c = (int)(i % 8);
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = c * 8 + 2 * e + h;
      v[h] = (k < p.k_avail[chunk] && r < p.n_avail[chunk])
                 ? p.params[p.w_off[chunk] + (size_t)(p.k_first[chunk] + k) * p.ld[chunk] + p.n_first[chunk] + r]
                 : 0.f;
    }
    w[e] = pack_bf16x2(v[0], v[1]);
  }
  *reinterpret_cast<uint4*>(p.packed + (size_t)chunk * kChunkBytes + sw128_offset(r, c)) = make_uint4(w[0], w[1], w[2], w[3]);
}

static bool tc_supported(const DurfMlpTopology& t) {
  return (t.width == 256 || t.width == 128) && t.cond_width == 128 && t.in_dim <= 64 && t.depth <= 12 && t.cond_dim <= 64 &&
         !(((t.depth - 1) % t.skip == 0) && t.depth - 1 > 0);
}

// Builds the per-layer schedule shared by the pack kernel and the MLP kernel.
static int build_sched(const DurfMlpTopology& t, TcParams& P) {
  MlpLayout L(t);
  const int KB = t.width / 64;
  int last_inp = 0;
  P.G = t.depth + 2;
  P.depth = t.depth;
  P.cond_dim = t.cond_dim;
  int chunks = 0;
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
      s.n_halves = 1; s.n_act_kb = KB; s.uses_inp = 0; s.kind = 3; s.bias_off = (int)L.b_off[t.depth + 2];
    }
    s.last_inp_use = 0;
    chunks += s.n_halves * (s.n_act_kb + s.uses_inp);
  }
  P.sched[last_inp].last_inp_use = 1;
  P.chunks_per_pair = chunks;
  P.off_wden = (int)L.w_off[t.depth]; P.off_bden = (int)L.b_off[t.depth];
  P.off_wrgb = (int)L.w_off[t.depth + 3]; P.off_brgb = (int)L.b_off[t.depth + 3];
  P.off_wview = (int)L.w_off[t.depth + 2] + t.width * t.cond_width;
  return chunks;
}

int64_t mlp_tc_packed_bytes(const DurfMlpTopology& t) {
  if (!tc_supported(t)) return 0;
  TcParams P;
  return (int64_t)build_sched(t, P) * kChunkBytes;
}

int mlp_tc_pack(cudaStream_t st, const DurfMlpTopology& t, const float* params, void* packed) {
  DURF_REQUIRE(tc_supported(t), DURF_E_UNSUPPORTED,
               "durf_mlp_pack_weights: tensor-core path needs width 128/256, cond_width 128, in_dim <= 64");
  TcParams P;
  const int chunks = build_sched(t, P);
  DURF_REQUIRE(chunks <= 160, DURF_E_UNSUPPORTED, "durf_mlp_pack_weights: too many weight chunks (%d)", chunks);
  MlpLayout L(t);
  PackParams pp;
  pp.params = params; pp.packed = (uint8_t*)packed; pp.n_chunks = chunks; pp.W = t.width; pp.in_dim = t.in_dim;
  int c = 0;
  for (int g = 0; g < P.G; ++g) {
    const LayerSched& s = P.sched[g];
    const int layer = (g < t.depth) ? g : (g == t.depth ? t.depth + 1 : t.depth + 2);
    const int n_out = L.out_dim[layer];
    for (int nh = 0; nh < s.n_halves; ++nh)
      for (int kc = 0; kc < s.n_act_kb + s.uses_inp; ++kc, ++c) {
        pp.w_off[c] = (int)L.w_off[layer];
        pp.ld[c] = n_out;
        pp.n_first[c] = nh * 128;
        pp.n_avail[c] = n_out - nh * 128 < 128 ? n_out - nh * 128 : 128;
        if (kc < s.n_act_kb) { pp.k_first[c] = kc * 64; pp.k_avail[c] = 64; }
        else { pp.k_first[c] = (g == 0) ? 0 : t.width; pp.k_avail[c] = t.in_dim; }   // input-feature block (skip rows follow the trunk rows)
      }
  }
  const int64_t total = (int64_t)chunks * 1024;
  pack_weights_kernel<<<ceil_div(total, 256), 256, 0, st>>>(pp);
  DURF_CHECK_LAUNCH("durf_mlp_pack_weights");
  return DURF_OK;
}

int mlp_tc_forward(cudaStream_t st, const DurfMlpArgs& a) {
  const DurfMlpTopology& t = a.topo;
  DURF_REQUIRE(tc_supported(t), DURF_E_UNSUPPORTED,
               "durf_mlp_fwd(bf16): tensor-core path needs width 128/256, cond_width 128, in_dim <= 64");
  DURF_REQUIRE(a.N == kTileM, DURF_E_UNSUPPORTED, "durf_mlp_fwd(bf16): needs 128 samples per ray (got %d)", a.N);
  DURF_REQUIRE(a.packed && a.params && a.features && a.cond, DURF_E_INVALID, "durf_mlp_fwd(bf16): null buffer");
  DURF_REQUIRE(a.saved == nullptr, DURF_E_UNSUPPORTED, "durf_mlp_fwd(bf16): activation saving is not available on this path");
  TcParams P;
  build_sched(t, P);
  P.feat = (const uint8_t*)a.features; P.cond = a.cond; P.params = a.params; P.packed = (const uint8_t*)a.packed;
  P.ray_index = a.ray_index; P.count = a.count; P.M = a.M; P.accumulate = a.accumulate;
  P.raw_rgb = a.raw_rgb; P.raw_density = a.raw_density;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int pairs = (a.M + 1) / 2;
  const int grid = pairs < sms ? pairs : sms;
  cudaError_t e;
  if (t.width == 256) {
    e = cudaFuncSetAttribute(mlp_tc_fwd_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256>::SMEM_BYTES);
    DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_fwd(bf16): smem attribute: %s", cudaGetErrorString(e));
    mlp_tc_fwd_kernel<256><<<grid, 384, TcCfg<256>::SMEM_BYTES, st>>>(P);
  } else {
    e = cudaFuncSetAttribute(mlp_tc_fwd_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM_BYTES);
    DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_fwd(bf16): smem attribute: %s", cudaGetErrorString(e));
    mlp_tc_fwd_kernel<128><<<grid, 384, TcCfg<128>::SMEM_BYTES, st>>>(P);
  }
  DURF_CHECK_LAUNCH("durf_mlp_fwd(bf16)");
  return DURF_OK;
}

}  // namespace durf
