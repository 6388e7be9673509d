// Micro-benchmarks behind the design of durf_b200/csrc/mlp_tc.cu (run on a B200: tools/ubench_tc).
//   1. L2 -> shared-memory streaming rate of cp.async.bulk when every CTA streams the SAME weight image (1.2 MB)
//      through an mbarrier ring, for several chunk sizes / ring depths / grid sizes.
//   2. tcgen05.mma issue rate from one thread: cycles per MMA for N = 64/128/256, A from TMEM or shared memory,
//      with and without a tcgen05.commit every few instructions.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = clock64();
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 2000000000LL) { printf("timeout bar %x\n", bar); __trap(); }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- 1. streaming ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* img, int img_bytes, int chunk, int stages, int reps, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + stages * chunk;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bar0 + 8 * s, 1); mbar_init(bar0 + 8 * (32 + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nchunks = img_bytes / chunk;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {          // producer
    uint32_t stage = 0, phase = 0;
    for (int r = 0; r < reps; ++r)
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bar0 + 8 * (32 + stage), phase ^ 1);
        mbar_arrive_expect_tx(bar0 + 8 * stage, chunk);
        bulk_g2s(sbase + stage * chunk, img + (size_t)c * chunk, chunk, bar0 + 8 * stage);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
  } else if (threadIdx.x == 32) {  // consumer: frees the stage as soon as it has landed
    uint32_t stage = 0, phase = 0;
    for (int r = 0; r < reps; ++r)
      for (int c = 0; c < nchunks; ++c) {
        mbar_wait(bar0 + 8 * stage, phase);
        mbar_arrive(bar0 + 8 * (32 + stage));
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

// ---- 2. MMA issue rate ---------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred;
}

// Same measurement with the issuing warp CONVERGED: all 32 lanes run the loop (waits included), only the tcgen05
// instructions are guarded by elect.sync, so descriptors stay in uniform registers.
template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) mma_kernel_conv(int iters, int commit_every, int with_wait, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[2];
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&s_bar[0]), 1); mbar_init(smem_u32(&s_bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tmem;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = idesc_bf16(128, N);
    const uint32_t bar = smem_u32(&s_bar[0]), bar_end = smem_u32(&s_bar[1]);
    const long long t0 = clock64();
    const int group = commit_every > 0 ? commit_every : iters;
    long long t_wait = 0;
    const uint64_t b_base = desc_sw128(sbase + 32768);
    for (int i0 = 0; i0 < iters; i0 += group) {
      if (with_wait) {
        const long long tq = clock64();
        mbar_wait(bar_end, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        t_wait += clock64() - tq;
      }
      const uint32_t leader = elect_one();
      for (int i = i0; i < i0 + group; ++i) {
        const uint64_t b_desc = b_base + (uint64_t)((i & 3) * (8192 >> 4) + ((i >> 2) & 1) * 2);
        if (leader) {
          if (TS) umma_ts(tm, tm + 256 + (i & 15) * 8, b_desc, idesc, 1u);
          else umma_ss(tm, desc_sw128(sbase) + (uint64_t)((i & 3) * 2), b_desc, idesc, 1u);
        }
      }
      if (commit_every > 0 && leader) tc_commit(bar);
      __syncwarp();
    }
    if (elect_one()) { tc_commit(bar_end); }
    __syncwarp();
    mbar_wait(bar_end, 0);
    if (threadIdx.x == 0) { cycles[blockIdx.x] = clock64() - t0; if (with_wait) cycles[200 + blockIdx.x] = t_wait; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) mma_kernel(int iters, int commit_every, int with_wait, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[2];
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&s_bar[0]), 1); mbar_init(smem_u32(&s_bar[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = s_tmem;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = idesc_bf16(128, N);
    const uint32_t bar = smem_u32(&s_bar[0]), bar_end = smem_u32(&s_bar[1]);
    const long long t0 = clock64();
    const int group = commit_every > 0 ? commit_every : iters;
    long long t_wait = 0;
    for (int i0 = 0; i0 < iters; i0 += group) {
      if (with_wait) {            // an already-completed mbarrier wait + fence per group, like the real issue loop
        const long long tq = clock64();
        mbar_wait(bar_end, 1);    // fresh barrier: the phase with parity 1 counts as complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        t_wait += clock64() - tq;
      }
      for (int i = i0; i < i0 + group; ++i) {
        const uint32_t b_addr = sbase + 32768 + (i & 3) * 8192 + ((i >> 2) & 1) * 32;
        if (TS) umma_ts(tm, tm + 256 + (i & 15) * 8, desc_sw128(b_addr), idesc, 1u);
        else umma_ss(tm, desc_sw128(sbase + (i & 3) * 32), desc_sw128(b_addr), idesc, 1u);
      }
      if (commit_every > 0) tc_commit(bar);        // nobody waits on it: only the cost of issuing the commit is measured
    }
    if (with_wait) cycles[200 + blockIdx.x] = t_wait;
    tc_commit(bar_end);
    mbar_wait(bar_end, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int N, bool TS>
static void run_mma(const char* name, int commit_every, int grid, int with_wait = 0, int conv = 0) {
  long long* d; CK(cudaMalloc(&d, 400 * 8));
  const int iters = 4096, smem = 96 * 1024 + 2048;
  CK(cudaFuncSetAttribute(mma_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(mma_kernel_conv<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) {
    if (conv) mma_kernel_conv<N, TS><<<grid, 128, smem>>>(iters, commit_every, with_wait, d);
    else mma_kernel<N, TS><<<grid, 128, smem>>>(iters, commit_every, with_wait, d);
    CK(cudaDeviceSynchronize());
  }
  long long h[400]; CK(cudaMemcpy(h, d, 400 * 8, cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
  printf("mma %s %-10s N=%3d commit_every=%2d wait=%d grid=%3d : %7.1f cyc/MMA  (ideal %d) -> %.0f%% of 4096 MAC/cyc/SM", conv ? "converged" : "lane0    ", name, N, commit_every,
         with_wait, grid, avg / iters, N / 2, 100.0 * (N / 2) / (avg / iters));
  if (with_wait) printf("   [completed try_wait+fence: %.0f cyc each]", (double)h[200] / (iters / commit_every));
  printf("\n");
  cudaFree(d);
}

int main() {
  // 1. streaming
  const int img_bytes = 76 * 16384;
  uint8_t* img; CK(cudaMalloc(&img, img_bytes)); CK(cudaMemset(img, 1, img_bytes));
  long long* d; CK(cudaMalloc(&d, 148 * 8));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int cfgs[][3] = {{16384, 3, 148}, {16384, 11, 148}, {32768, 5, 148}, {8192, 22, 148}, {16384, 11, 74}, {16384, 11, 37}, {16384, 11, 8}, {16384, 11, 1}, {65536, 2, 148}, {65536, 3, 148}};
  for (auto& c : cfgs) {
    const int chunk = c[0], stages = c[1], grid = c[2], reps = 50;
    const int smem = stages * chunk + 1024 + 1024;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    stream_kernel<<<grid, 64, smem>>>(img, img_bytes, chunk, stages, 2, d);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    stream_kernel<<<grid, 64, smem>>>(img, img_bytes, chunk, stages, reps, d);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; CK(cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    const double bytes = (double)img_bytes * reps;
    printf("stream chunk=%6d stages=%2d grid=%3d : %6.2f B/cyc/SM, aggregate %7.1f GB/s (%.3f ms)\n", chunk, stages, grid, bytes / avg,
           bytes * grid / (ms * 1e-3) / 1e9, ms);
  }
  // 2. MMA issue
  for (int ce : {0, 4, 8}) {
    run_mma<128, true>("A=TMEM", ce, 148);
    run_mma<256, true>("A=TMEM", ce, 148);
    run_mma<128, false>("A=SMEM", ce, 148);
  }
  run_mma<64, true>("A=TMEM", 4, 148);
  run_mma<128, true>("A=TMEM", 4, 148, 1);
  run_mma<128, true>("A=TMEM", 8, 148, 1);
  run_mma<256, true>("A=TMEM", 4, 148, 1);
  run_mma<128, true>("A=TMEM", 2, 148, 1);
  for (int ce : {0, 2, 4, 8, 16}) {
    run_mma<128, true>("A=TMEM", ce, 148, 0, 1);
    if (ce) run_mma<128, true>("A=TMEM", ce, 148, 1, 1);
  }
  run_mma<64, true>("A=TMEM", 4, 148, 1, 1);
  run_mma<256, true>("A=TMEM", 4, 148, 1, 1);
  run_mma<128, false>("A=SMEM", 4, 148, 1, 1);
  return 0;
}
