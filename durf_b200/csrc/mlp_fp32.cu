// K2 (parity mode) -- the radiance/density MLP in fp32 on CUDA cores, forward and backward, layer by layer.
// Replaces MLP.__call__ / BoxMLP.__call__ (obbpose_model.py:305-354, 369-418) with flax nn.Dense semantics
// (y = x @ kernel[in,out] + bias).  This path exists for <=1e-5 parity against the fp32 oracle and for gradient
// parity; the throughput path is the tcgen05 chain in mlp_tc.cu.  Inner products use explicit fmaf (the library is
// compiled with -fmad=false so that the ray math keeps the reference's operation order).
#include "common.cuh"
#include "mlp_topology.h"

namespace durf {

// Row source made of up to two column blocks: A = [A1 | A2].  A2 rows may be shared by `a2_div` consecutive rows
// (the view encoding is per ray: obbpose_model.py:343-347) and routed through an index (compacted object rays).
struct RowSrc {
  const float* a1; int lda1; int k1;
  const float* a2; int lda2; int k2; int a2_div; const int32_t* a2_index;
  __device__ __forceinline__ float at(int r, int k) const {
    if (k < k1) return a1[(size_t)r * lda1 + k];
    int rr = r / a2_div;
    if (a2_index) rr = a2_index[rr];
    return a2[(size_t)rr * lda2 + (k - k1)];
  }
};

constexpr int BM = 64, BN = 64, BK = 16;

// C[M,Nn] = act(A[M,K] W[K,Nn] + b)
template <bool RELU>
__global__ void __launch_bounds__(256)
gemm_nn_kernel(int M, int Nn, RowSrc A, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
               float* __restrict__ C, int ldc) {
  __shared__ float sA[BK][BM + 1];
  __shared__ float sW[BK][BN];
  const int K = A.k1 + A.k2;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += BK) {
    for (int i = threadIdx.x; i < BM * BK; i += 256) {
      const int m = i / BK, k = i % BK;
      sA[k][m] = (m0 + m < M && k0 + k < K) ? A.at(m0 + m, k0 + k) : 0.f;
    }
    for (int i = threadIdx.x; i < BK * BN; i += 256) {
      const int k = i / BN, n = i % BN;
      sW[k][n] = (k0 + k < K && n0 + n < Nn) ? W[(size_t)(k0 + k) * ldw + n0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; w[i] = sW[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Nn) continue;
      float v = acc[i][j] + bias[n];
      if (RELU) v = (v < 0.f) ? 0.f : v;          // jnp.maximum(x, 0) propagates NaN (fmaxf would return 0)
      C[(size_t)m * ldc + n] = v;
    }
  }
}

// dA[M,K] (+)= (dZ[M,Nn] W[K,Nn]^T) * mask, mask = (H[M,K] > 0) when H given (ReLU of the producing layer).
__global__ void __launch_bounds__(256)
gemm_nt_kernel(int M, int K, int Nn, const float* __restrict__ dZ, int ldz, const float* __restrict__ W, int ldw,
               const float* __restrict__ H, int ldh, float* __restrict__ dA, int lda, int accumulate) {
  __shared__ float sZ[BK][BM + 1];
  __shared__ float sW[BK][BN + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * BM, k0 = blockIdx.x * BN;
  float acc[4][4] = {};
  for (int n0 = 0; n0 < Nn; n0 += BK) {
    for (int i = threadIdx.x; i < BM * BK; i += 256) {
      const int m = i / BK, n = i % BK;
      sZ[n][m] = (m0 + m < M && n0 + n < Nn) ? dZ[(size_t)(m0 + m) * ldz + n0 + n] : 0.f;
    }
    for (int i = threadIdx.x; i < BN * BK; i += 256) {
      const int k = i / BK, n = i % BK;
      sW[n][k] = (k0 + k < K && n0 + n < Nn) ? W[(size_t)(k0 + k) * ldw + n0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < BK; ++n) {
      float z[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { z[i] = sZ[n][ty * 4 + i]; w[i] = sW[n][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(z[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= K) continue;
      float v = acc[i][j];
      if (accumulate) v += dA[(size_t)m * lda + k];
      if (H && !(H[(size_t)m * ldh + k] > 0.f)) v = 0.f;
      dA[(size_t)m * lda + k] = v;
    }
  }
}

// dW[K,Nn] += A[M,K]^T dZ[M,Nn]: the row range is split over blockIdx.z, partial tiles are added with fp32 atomics.
constexpr int kRowsPerSplit = 2048;
__global__ void __launch_bounds__(256)
gemm_tn_kernel(int M, int Nn, RowSrc A, int k_begin, int k_count, const float* __restrict__ dZ, int ldz,
               float* __restrict__ dW, int ldw) {
  __shared__ float sA[BK][BM + 1];   // [row][k]
  __shared__ float sZ[BK][BN + 1];   // [row][n]
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int k0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int r_begin = blockIdx.z * kRowsPerSplit, r_end = min(M, r_begin + kRowsPerSplit);
  float acc[4][4] = {};
  for (int r0 = r_begin; r0 < r_end; r0 += BK) {
    for (int i = threadIdx.x; i < BK * BM; i += 256) {
      const int r = i / BM, k = i % BM;
      sA[r][k] = (r0 + r < r_end && k0 + k < k_count) ? A.at(r0 + r, k_begin + k0 + k) : 0.f;
    }
    for (int i = threadIdx.x; i < BK * BN; i += 256) {
      const int r = i / BN, n = i % BN;
      sZ[r][n] = (r0 + r < r_end && n0 + n < Nn) ? dZ[(size_t)(r0 + r) * ldz + n0 + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < BK; ++r) {
      float a[4], z[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[r][ty * 4 + i]; z[i] = sZ[r][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], z[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= k_count) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < Nn) atomicAdd(&dW[(size_t)(k_begin + k) * ldw + n], acc[i][j]);
    }
  }
}

// db[Nn] += column sums of dZ[M,Nn]
__global__ void __launch_bounds__(256)
colsum_kernel(int M, int Nn, const float* __restrict__ dZ, int ldz, float* __restrict__ db) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rows_per_block = 4096;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < Nn)
    for (int r = r0 + (threadIdx.x >> 5); r < r1; r += 8) s += dZ[(size_t)r * ldz + n];
  __shared__ float red[8][33];
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && n < Nn) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(&db[n], t);
  }
}

// Narrow heads (density: 1 output, rgb: 3 outputs): one warp per row, output routed to the ray's slot.
template <int NOUT>
__global__ void __launch_bounds__(256)
head_fwd_kernel(int R, int N, RowSrc A, const float* __restrict__ W, const float* __restrict__ bias,
                const int32_t* __restrict__ ray_index, int accumulate, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  const int K = A.k1 + A.k2;
  float s[NOUT] = {};
  for (int k = lane; k < K; k += 32) {
    const float a = A.at(r, k);
#pragma unroll
    for (int j = 0; j < NOUT; ++j) s[j] = fmaf(a, W[(size_t)k * NOUT + j], s[j]);
  }
#pragma unroll
  for (int j = 0; j < NOUT; ++j) s[j] = warp_sum(s[j]);
  if (lane == 0) {
    const int m = r / N, n = r - m * N;
    const size_t row = (size_t)(ray_index ? ray_index[m] : m) * N + n;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      const float v = s[j] + bias[j];
      if (accumulate) out[row * NOUT + j] += v; else out[row * NOUT + j] = v;
    }
  }
}

// Gathers the head gradients of the (possibly compacted) rows into a dense [R, NOUT] buffer.
template <int NOUT>
__global__ void gather_head_grad_kernel(int R, int N, const float* __restrict__ g, const int32_t* __restrict__ ray_index,
                                        float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * NOUT) return;
  const int r = (int)(i / NOUT), j = (int)(i % NOUT);
  const int m = r / N, n = r - m * N;
  out[i] = g[((size_t)(ray_index ? ray_index[m] : m) * N + n) * NOUT + j];
}

__global__ void read_count_kernel(const int32_t* count, int32_t* out) { *out = *count; }

static void launch_nn(cudaStream_t st, bool relu, int M, int Nn, const RowSrc& A, const float* W, int ldw, const float* b,
                      float* C, int ldc) {
  dim3 grid(ceil_div(Nn, BN), ceil_div(M, BM));
  if (relu) gemm_nn_kernel<true><<<grid, 256, 0, st>>>(M, Nn, A, W, ldw, b, C, ldc);
  else gemm_nn_kernel<false><<<grid, 256, 0, st>>>(M, Nn, A, W, ldw, b, C, ldc);
  count_launch();
}

static RowSrc plain(const float* a, int lda, int k) { return RowSrc{a, lda, k, nullptr, 0, 0, 1, nullptr}; }

// Workspace / saved-activation layout (floats), R = rows:
//   h[i]   i = 0..depth-1 : [R, width]  post-ReLU trunk activations
//   bott                  : [R, width]
//   cond                  : [R, cond_width]
// Inference keeps only two trunk buffers (ping-pong); training keeps all of them in `saved`.
size_t mlp_fp32_saved_floats(const DurfMlpTopology& t, int64_t R) {
  return (size_t)R * ((size_t)t.depth * t.width + t.width + t.cond_width);
}
size_t mlp_fp32_infer_floats(const DurfMlpTopology& t, int64_t R) {
  return (size_t)R * (3 * (size_t)t.width + t.cond_width);
}
size_t mlp_fp32_bwd_floats(const DurfMlpTopology& t, int64_t R) {
  // dZ ping-pong [R,width] x2, d_cond [R,cond_width], head grads [R,4]
  return (size_t)R * (2 * (size_t)t.width + t.cond_width + 4);
}

int mlp_fp32_forward(cudaStream_t st, const DurfMlpArgs& a) {
  const DurfMlpTopology& t = a.topo;
  const int64_t R = (int64_t)a.M * a.N;
  DURF_REQUIRE(a.count == nullptr, DURF_E_UNSUPPORTED,
               "durf_mlp_fwd(fp32): the parity path takes a host-known row count (pass M, not a device count)");
  const bool training = a.saved != nullptr;
  const size_t need = training ? 0 : mlp_fp32_infer_floats(t, R) * sizeof(float);
  DURF_REQUIRE(training || (a.workspace && a.workspace_bytes >= need), DURF_E_WORKSPACE,
               "durf_mlp_fwd(fp32): workspace %zu < %zu bytes", a.workspace_bytes, need);
  float* base = training ? (float*)a.saved : (float*)a.workspace;
  const size_t RW = (size_t)R * t.width;
  auto hbuf = [&](int i) { return base + (training ? (size_t)i : (size_t)(i & 1)) * RW; };
  float* bott = base + (training ? (size_t)t.depth : 2) * RW;
  float* cond = bott + RW;
  const float* x = (const float*)a.features;
  MlpLayout L(t);

  RowSrc in = plain(x, t.in_dim, t.in_dim);
  for (int i = 0; i < t.depth; ++i) {
    launch_nn(st, true, (int)R, t.width, in, a.params + L.w_off[i], t.width, a.params + L.b_off[i], hbuf(i), t.width);
    in = plain(hbuf(i), t.width, t.width);
    if (i % t.skip == 0 && i > 0) { in.a2 = x; in.lda2 = t.in_dim; in.k2 = t.in_dim; in.a2_div = 1; }
  }
  const int iD = t.depth, iB = t.depth + 1, iC = t.depth + 2, iR = t.depth + 3;
  const int32_t* out_index = a.accumulate == 2 ? nullptr : a.ray_index;      // 2: compact output rows (durf_mlp_merge_raw)
  head_fwd_kernel<1><<<ceil_div(R, 8), 256, 0, st>>>((int)R, a.N, in, a.params + L.w_off[iD], a.params + L.b_off[iD],
                                                      out_index, a.accumulate == 1, a.raw_density);
  count_launch();
  launch_nn(st, false, (int)R, t.width, in, a.params + L.w_off[iB], t.width, a.params + L.b_off[iB], bott, t.width);
  RowSrc cin = RowSrc{bott, t.width, t.width, a.cond, t.cond_dim, t.cond_dim, a.N, a.ray_index};
  launch_nn(st, true, (int)R, t.cond_width, cin, a.params + L.w_off[iC], t.cond_width, a.params + L.b_off[iC], cond, t.cond_width);
  head_fwd_kernel<3><<<ceil_div(R, 8), 256, 0, st>>>((int)R, a.N, plain(cond, t.cond_width, t.cond_width), a.params + L.w_off[iR],
                                                      a.params + L.b_off[iR], out_index, a.accumulate == 1, a.raw_rgb);
  count_launch();
  cudaError_t e = cudaGetLastError();
  DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_fwd(fp32): %s", cudaGetErrorString(e));
  return DURF_OK;
}

static void launch_tn(cudaStream_t st, int M, int Nn, const RowSrc& A, int k_begin, int k_count, const float* dZ, int ldz,
                      float* dW, int ldw) {
  dim3 grid(ceil_div(Nn, BN), ceil_div(k_count, BM), ceil_div(M, kRowsPerSplit));
  gemm_tn_kernel<<<grid, 256, 0, st>>>(M, Nn, A, k_begin, k_count, dZ, ldz, dW, ldw);
  count_launch();
}
static void launch_nt(cudaStream_t st, int M, int K, int Nn, const float* dZ, int ldz, const float* W, int ldw, const float* H,
                      int ldh, float* dA, int lda, int accumulate) {
  dim3 grid(ceil_div(K, BN), ceil_div(M, BM));
  gemm_nt_kernel<<<grid, 256, 0, st>>>(M, K, Nn, dZ, ldz, W, ldw, H, ldh, dA, lda, accumulate);
  count_launch();
}
static void launch_colsum(cudaStream_t st, int M, int Nn, const float* dZ, int ldz, float* db) {
  dim3 grid(ceil_div(Nn, 32), ceil_div(M, 4096));
  colsum_kernel<<<grid, 256, 0, st>>>(M, Nn, dZ, ldz, db);
  count_launch();
}

int mlp_fp32_backward(cudaStream_t st, const DurfMlpArgs& a, const float* d_raw_rgb, const float* d_raw_density,
                      float* d_params, float* d_features) {
  const DurfMlpTopology& t = a.topo;
  const int R = a.M * a.N;
  DURF_REQUIRE(a.saved, DURF_E_INVALID, "durf_mlp_bwd(fp32): needs the activations saved by durf_mlp_fwd");
  DURF_REQUIRE(a.count == nullptr, DURF_E_UNSUPPORTED, "durf_mlp_bwd(fp32): pass a host-known row count");
  const size_t need = mlp_fp32_bwd_floats(t, R) * sizeof(float);
  DURF_REQUIRE(a.workspace && a.workspace_bytes >= need, DURF_E_WORKSPACE, "durf_mlp_bwd(fp32): workspace %zu < %zu bytes",
               a.workspace_bytes, need);
  const float* sv = (const float*)a.saved;
  const size_t RW = (size_t)R * t.width;
  auto h = [&](int i) { return sv + (size_t)i * RW; };
  const float* bott = sv + (size_t)t.depth * RW;
  const float* cond = bott + RW;
  float* ws = (float*)a.workspace;
  float* g0 = ws; float* g1 = ws + RW; float* gcond = ws + 2 * RW; float* ghead = gcond + (size_t)R * t.cond_width;
  const float* x = (const float*)a.features;
  MlpLayout L(t);
  const int iD = t.depth, iB = t.depth + 1, iC = t.depth + 2, iR = t.depth + 3;
  auto W = [&](int i) { return a.params + L.w_off[i]; };
  auto dWp = [&](int i) { return d_params + L.w_off[i]; };
  auto dbp = [&](int i) { return d_params + L.b_off[i]; };

  // rgb head
  gather_head_grad_kernel<3><<<ceil_div((int64_t)R * 3, 256), 256, 0, st>>>(R, a.N, d_raw_rgb, a.ray_index, ghead);
  count_launch();
  launch_tn(st, R, 3, plain(cond, t.cond_width, t.cond_width), 0, t.cond_width, ghead, 3, dWp(iR), 3);
  launch_colsum(st, R, 3, ghead, 3, dbp(iR));
  launch_nt(st, R, t.cond_width, 3, ghead, 3, W(iR), 3, cond, t.cond_width, gcond, t.cond_width, 0);   // masked by ReLU(cond)
  // condition layer: input [bott | viewenc]
  RowSrc cin = RowSrc{bott, t.width, t.width, a.cond, t.cond_dim, t.cond_dim, a.N, a.ray_index};
  launch_tn(st, R, t.cond_width, cin, 0, t.width + t.cond_dim, gcond, t.cond_width, dWp(iC), t.cond_width);
  launch_colsum(st, R, t.cond_width, gcond, t.cond_width, dbp(iC));
  launch_nt(st, R, t.width, t.cond_width, gcond, t.cond_width, W(iC), t.cond_width, nullptr, 0, g0, t.width, 0);  // d bottleneck
  // heads on the last trunk activation
  const bool last_skip = ((t.depth - 1) % t.skip == 0) && (t.depth - 1 > 0);
  DURF_REQUIRE(!last_skip, DURF_E_UNSUPPORTED, "durf_mlp_bwd(fp32): a skip concat on the last trunk layer is not supported");
  RowSrc hin = plain(h(t.depth - 1), t.width, t.width);
  launch_tn(st, R, t.width, hin, 0, t.width, g0, t.width, dWp(iB), t.width);
  launch_colsum(st, R, t.width, g0, t.width, dbp(iB));
  gather_head_grad_kernel<1><<<ceil_div((int64_t)R, 256), 256, 0, st>>>(R, a.N, d_raw_density, a.ray_index, ghead);
  count_launch();
  launch_tn(st, R, 1, hin, 0, t.width, ghead, 1, dWp(iD), 1);
  launch_colsum(st, R, 1, ghead, 1, dbp(iD));
  // dh = g_bott W_B^T, then += g_den W_D^T, masked by ReLU(h_last)
  launch_nt(st, R, t.width, t.width, g0, t.width, W(iB), t.width, nullptr, 0, g1, t.width, 0);
  launch_nt(st, R, t.width, 1, ghead, 1, W(iD), 1, h(t.depth - 1), t.width, g1, t.width, 1);
  float* gz = g1; float* gnext = g0;
  bool feat_written = false;
  for (int i = t.depth - 1; i >= 0; --i) {
    const bool has_skip_in = (i >= 1) && ((i - 1) % t.skip == 0) && (i - 1 > 0);
    RowSrc in = (i == 0) ? plain(x, t.in_dim, t.in_dim) : plain(h(i - 1), t.width, t.width);
    if (has_skip_in) { in.a2 = x; in.lda2 = t.in_dim; in.k2 = t.in_dim; in.a2_div = 1; }
    const int K = in.k1 + in.k2;
    launch_tn(st, R, t.width, in, 0, K, gz, t.width, dWp(i), t.width);
    launch_colsum(st, R, t.width, gz, t.width, dbp(i));
    if (i > 0) {
      launch_nt(st, R, t.width, t.width, gz, t.width, W(i), t.width, h(i - 1), t.width, gnext, t.width, 0);
      if (has_skip_in && d_features) {  // gradient into the re-concatenated input block (rows width.. of W_i)
        launch_nt(st, R, t.in_dim, t.width, gz, t.width, W(i) + (size_t)t.width * t.width, t.width, nullptr, 0, d_features,
                  t.in_dim, feat_written ? 1 : 0);
        feat_written = true;
      }
      float* tmp = gz; gz = gnext; gnext = tmp;
    } else if (d_features) {
      launch_nt(st, R, t.in_dim, t.width, gz, t.width, W(0), t.width, nullptr, 0, d_features, t.in_dim, feat_written ? 1 : 0);
    }
  }
  cudaError_t e = cudaGetLastError();
  DURF_REQUIRE(e == cudaSuccess, DURF_E_LAUNCH, "durf_mlp_bwd(fp32): %s", cudaGetErrorString(e));
  return DURF_OK;
}

}  // namespace durf
