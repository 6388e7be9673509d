// Does a second tcgen05.mma-issuing thread raise the M=128,N=128 MMA rate?  One thread issues ~79-109 cycles per MMA against 64
// cycles of pipe time (tools/ubench_tc.cu); if that is a per-thread issue limit, two threads (two warps, two accumulator
// halves) should approach 64 cycles per MMA in aggregate; if it is a property of the tensor pipe, nothing changes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_two_issuers tools/ubench_two_issuers.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = clock64();
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 2000000000LL) { printf("timeout bar %x\n", bar); __trap(); }
  }
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
// four K=16 MMAs in one asm block, A from TMEM (+8 columns per step), B descriptor +2 per step: the production issue pattern
__device__ __forceinline__ void umma4_ts(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 db;\n.reg .b32 al, bl;\nsetp.eq.u32 p, %0, %0;\n"
      "mov.b64 db, {%2, %3};\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
      "add.u32 bl, %2, 2;\n add.u32 al, %1, 8;\n mov.b64 db, {bl, %3};\n tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, p;\n"
      "add.u32 bl, %2, 4;\n add.u32 al, %1, 16;\n mov.b64 db, {bl, %3};\n tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, p;\n"
      "add.u32 bl, %2, 6;\n add.u32 al, %1, 24;\n mov.b64 db, {bl, %3};\n tcgen05.mma.cta_group::1.kind::f16 [%0], [al], db, %4, p;\n"
      "}\n" ::"r"(d), "r"(a), "r"(b_lo), "r"(b_hi), "r"(idesc) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

template <int N>
__global__ void __launch_bounds__(128, 1) k(int iters, int issuers, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar[2];
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
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
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < issuers) {
    constexpr uint32_t idesc = idesc_bf16(128, N);
    const uint32_t bar = smem_u32(&s_bar[w]);
    const uint64_t bd = desc_sw128(sbase + w * 65536);            // each issuer reads its own 64 KB of "weights"
    const uint32_t b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4)
      umma4_ts(tm + w * N, tm + 256 + ((i >> 2) & 3) * 32, b_lo + ((i >> 2) & 3) * (16384 >> 4), b_hi, idesc);
    tc_commit(bar);
    mbar_wait(bar, 0);
    cycles[blockIdx.x * 2 + w] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int N>
static void run(int issuers, int grid) {
  long long* d; CK(cudaMalloc(&d, 400 * 8)); CK(cudaMemset(d, 0, 400 * 8));
  const int iters = 8192, smem = 128 * 1024 + 2048;
  CK(cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int rep = 0; rep < 2; ++rep) { k<N><<<grid, 128, smem>>>(iters, issuers, d); CK(cudaDeviceSynchronize()); }
  long long h[400]; CK(cudaMemcpy(h, d, 400 * 8, cudaMemcpyDeviceToHost));
  double mx = 0; for (int i = 0; i < grid; ++i) { double m = h[2 * i] > h[2 * i + 1] ? h[2 * i] : h[2 * i + 1]; mx += m; } mx /= grid;
  printf("N=%3d issuers=%d grid=%3d : %7.1f cycles per MMA in aggregate (pipe time %d) -> %.0f%% of the tensor pipe\n", N, issuers, grid,
         mx / (iters * issuers), N / 2, 100.0 * (N / 2) / (mx / (iters * issuers)));
  cudaFree(d);
}
int main() {
  for (int g : {1, 148}) { run<128>(1, g); run<128>(2, g); run<64>(1, g); run<64>(2, g); run<256>(1, g); }
  return 0;
}
