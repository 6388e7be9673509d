// Shared device/host helpers for libdurf_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/durf_b200.h"

namespace durf {

// ---- host side: error reporting and launch accounting -------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DURF_REQUIRE(cond, code, ...)                 \
  do {                                                \
    if (!(cond)) {                                    \
      ::durf::set_error(__VA_ARGS__);                 \
      return (code);                                  \
    }                                                 \
  } while (0)

// Checks the launch itself (no synchronisation): configuration errors surface here.
#define DURF_CHECK_LAUNCH(name)                                                          \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      ::durf::set_error("%s: launch failed: %s", (name), cudaGetErrorString(e__));       \
      return DURF_E_LAUNCH;                                                              \
    }                                                                                    \
    ::durf::count_launch();                                                              \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// true if every non-null pointer is 16-byte aligned (rows moved with 128-bit loads / stores)
template <class... P>
static inline bool aligned16(const P*... p) { return ((... | reinterpret_cast<uintptr_t>(p)) & 15u) == 0; }

// ---- weight images of the tensor-core kernels (mlp_tc*.cu) ------------------------------------------
// One 16 KB block of a weight image: block[r][k] = src[r * sr + k * sk] for r < r_avail, k < k_avail, else 0 (the forward image
// takes rows = output columns of a kernel, the transposed image of the data-gradient kernel rows = input rows).
struct PackBlock {
  const float* src;
  uint8_t* dst;
  int32_t sr, sk, r_avail, k_avail;
};
constexpr int kPackMaxBlocks = 448;     // blocks per launch of pack_blocks_kernel (14 KB of kernel parameters)
int pack_blocks_launch(cudaStream_t st, const PackBlock* blocks, int n);

// ---- device helpers ---------------------------------------------------------------------------
constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
// Inclusive scan across the warp.
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Exclusive prefix (earlier lanes) and exclusive suffix (later lanes) sums, without subtracting totals.
__device__ __forceinline__ float warp_scan_excl(float v, int lane) {
  const float incl = warp_scan_incl(v, lane);
  const float up = __shfl_up_sync(kFull, incl, 1);
  return lane == 0 ? 0.f : up;
}
__device__ __forceinline__ float warp_rscan_excl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(kFull, v, o);
    if (lane + o < 32) v += t;
  }
  const float dn = __shfl_down_sync(kFull, v, 1);
  return lane == 31 ? 0.f : dn;
}

// jnp.nan_to_num with default fills: NaN -> 0, +-inf -> +-FLT_MAX (mip.py:313, math.py:282).
__device__ __forceinline__ float nan_to_num(float x) {
  if (x != x) return 0.f;
  if (isinf(x)) return x > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
  return x;
}

// reference internal/math.py:35-36: sin(where(|x| < 100*pi, x, x % (100*pi))); `%` is floored remainder
// on top of an exact fmod.  Accurate sinf (never __sinf): IPE evaluates sin(2^l x) at large arguments.
__device__ __forceinline__ float safe_arg(float x) {
  const float t = 314.15926535897932f;  // fl32(100*pi)
  if (fabsf(x) < t) return x;
  float r = fmodf(x, t);
  if (r != 0.f && r < 0.f) r += t;
  return r;
}
__device__ __forceinline__ float safe_sinf(float x) { return sinf(safe_arg(x)); }
__device__ __forceinline__ float safe_cosf(float x) { return cosf(safe_arg(x)); }

constexpr float kHalfPi = 1.57079632679489662f;  // fl32(0.5*pi), added in fp32 like `y + 0.5*jnp.pi`

// Byte offset of 16-byte chunk `c` (0..7) of row `r` inside a K-major SWIZZLE_128B operand tile whose rows
// are 128 bytes (64 bf16): 8-row groups of 1024 B, chunk index XORed with (row & 7).  This is the canonical
// UMMA / TMA 128B-swizzle layout (cute Swizzle<3,4,3>), used for activations, features and weights alike.
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t c) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((c ^ (r & 7u)) << 4);
}

// 32-byte store of the logical 16-byte chunks (c, c+1), c even, of row r of a SWIZZLE_128B block image: they occupy one
// aligned physical pair (swapped on odd rows), so a single STG.256 fills a whole 32-byte sector.
__device__ __forceinline__ void st_sw128_pair(uint8_t* blk, uint32_t r, uint32_t c, const uint32_t (&w)[8]) {
  const uint32_t r7 = r & 7u;
  uint8_t* dst = blk + (r >> 3) * 1024u + r7 * 128u + (((c ^ r7) & ~1u) << 4);
  if (r7 & 1u)
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[0]),
                 "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  else
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
                 "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace durf
