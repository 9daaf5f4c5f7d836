// gemm.cu -- fp32 SIMT GEMM with fused BN+ReLU operand prologues, bias epilogue and deterministic split-K.
//
// Exact-fp32 reference kernel, selected with YOLAT_GEMM=simt (debugging aid; the default path is the tcgen05
// 3xTF32 kernel in gemm_tc.cu):  y = x W^T + b  (gcn_lib/sparse/torch_nn.py:58),
// dx = dy W and dW = dy^T x (autograd of the same).  Tiles: 64x64x16, 256 threads, 4x4 per thread.
#include "common.cuh"

namespace yolat {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int GEMM_THREADS = 256;

template <int MODE>
__device__ __forceinline__ float load_a(const GemmArgs& g, int m, int64_t k) {
  if (m >= g.M || k >= g.K) return 0.f;
  float v;
  if (MODE == GEMM_TN) {
    v = g.A[k * g.lda + m];
    if (g.a_sc) v = fmaxf(fmaf(v, g.a_sc[m], g.a_sh[m]), 0.f);
  } else {
    v = g.A[(int64_t)m * g.lda + k];
    if (g.a_sc) v = fmaxf(fmaf(v, g.a_sc[k], g.a_sh[k]), 0.f);
  }
  return v;
}

template <int MODE>
__device__ __forceinline__ float load_b(const GemmArgs& g, int64_t k, int n) {
  if (n >= g.N || k >= g.K) return 0.f;
  float v;
  if (MODE == GEMM_NT) {
    v = g.B[(int64_t)n * g.ldb + k];
  } else {
    v = g.B[k * g.ldb + n];
    if (g.b_sc) v = fmaxf(fmaf(v, g.b_sc[n], g.b_sh[n]), 0.f);
  }
  return v;
}

// grid: (ceil(N/BN), ceil(M/BM), ksplit).  ksplit > 1: raw partial tiles go to part[z][M][N].
template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS) k_gemm(GemmArgs g, int64_t k_chunk, float* __restrict__ part) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * k_chunk;
  const int64_t k_end = min(g.K, k_begin + k_chunk);
  const int tx = tid & 15, ty = tid >> 4;  // thread tile origin: rows ty*4.., cols tx*4..

  // loader mapping: each thread moves 4 elements of A and 4 of B per k-tile.
  // contiguous-in-k operands: thread -> (row = tid/4, k4 = (tid%4)*4); contiguous-in-m/n: (k = tid/16, col4 = (tid%16)*4)
  float ra[4], rb[4];
  auto fetch = [&](int64_t kt) {
    if (MODE == GEMM_TN) {
      const int kk = tid >> 4, c4 = (tid & 15) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) ra[q] = load_a<MODE>(g, m0 + c4 + q, kt + kk);
    } else {
      const int r = tid >> 2, k4 = (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) ra[q] = load_a<MODE>(g, m0 + r, kt + k4 + q);
    }
    if (MODE == GEMM_NT) {
      const int r = tid >> 2, k4 = (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) rb[q] = load_b<MODE>(g, kt + k4 + q, n0 + r);
    } else {
      const int kk = tid >> 4, c4 = (tid & 15) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) rb[q] = load_b<MODE>(g, kt + kk, n0 + c4 + q);
    }
  };
  auto stash = [&](int buf) {
    if (MODE == GEMM_TN) {
      const int kk = tid >> 4, c4 = (tid & 15) * 4;
      *reinterpret_cast<float4*>(&As[buf][kk][c4]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    } else {
      const int r = tid >> 2, k4 = (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) As[buf][k4 + q][r] = ra[q];
    }
    if (MODE == GEMM_NT) {
      const int r = tid >> 2, k4 = (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) Bs[buf][k4 + q][r] = rb[q];
    } else {
      const int kk = tid >> 4, c4 = (tid & 15) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][kk][c4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  int buf = 0;
  if (k_begin < k_end) {
    fetch(k_begin);
    stash(0);
  }
  __syncthreads();
  for (int64_t kt = k_begin; kt < k_end; kt += BK) {
    const bool has_next = kt + BK < k_end;
    if (has_next) fetch(kt + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      if (split) {
        part[((int64_t)blockIdx.z * g.M + m) * g.N + n] = acc[i][j];
      } else {
        float v = acc[i][j];
        if (g.bias) v += g.bias[n];
        float* c = g.C + (int64_t)m * g.ldc + n;
        *c = g.accumulate ? (*c + v) : v;
      }
    }
  }
}

__global__ void k_splitk_reduce(const float* __restrict__ part, int ksplit, int M, int N, const float* __restrict__ bias,
                                float* __restrict__ C, int64_t ldc, int accumulate) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * N) return;
  int m = (int)(idx / N), n = (int)(idx % N);
  float s = 0.f;
  for (int z = 0; z < ksplit; ++z) s += part[(int64_t)z * M * N + idx];
  if (bias) s += bias[n];
  float* c = C + (int64_t)m * ldc + n;
  *c = accumulate ? (*c + s) : s;
}

__global__ void k_bias_fill(int M, int N, const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int accumulate) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * N) return;
  int m = (int)(idx / N), n = (int)(idx % N);
  float v = bias ? bias[n] : 0.f;
  float* c = C + (int64_t)m * ldc + n;
  if (accumulate) *c += v; else *c = v;
}

int gemm_simt(const GemmArgs& a, GemmMode mode, Arena& ws, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) return YOLAT_OK;
  const int gm = (int)cdiv(a.M, BM), gn = (int)cdiv(a.N, BN);
  // split-K: enough CTAs to cover the machine ~2x, chunks multiple of BK, at least 8 k-tiles per chunk
  int ksplit = 1;
  const int64_t tiles = (int64_t)gm * gn;
  const int64_t ktiles = cdiv(a.K, BK);
  if (tiles < 2 * kNumSMs && ktiles >= 16) {
    const int64_t want = cdiv(2 * kNumSMs, tiles), cap = ktiles / 8;
    ksplit = (int)(want < cap ? want : cap);
    if (ksplit < 1) ksplit = 1;
    if (ksplit > 512) ksplit = 512;
  }
  int64_t k_chunk = align_up(cdiv(a.K > 0 ? a.K : 1, ksplit), BK);
  ksplit = (int)cdiv(a.K > 0 ? a.K : 1, k_chunk);
  float* part = nullptr;
  if (ksplit > 1) part = ws.take((int64_t)ksplit * a.M * a.N);
  if (ws.dry()) return YOLAT_OK;
  if (ws.overflow) return YOLAT_ERR_WORKSPACE;
  if (a.K <= 0) {
    int64_t tot = (int64_t)a.M * a.N;
    k_bias_fill<<<(unsigned)cdiv(tot, 256), 256, 0, st>>>(a.M, a.N, a.bias, a.C, a.ldc, a.accumulate);
    YOLAT_CHECK_LAUNCH();
    return YOLAT_OK;
  }
  dim3 grid(gn, gm, ksplit);
  switch (mode) {
    case GEMM_NT: k_gemm<GEMM_NT><<<grid, GEMM_THREADS, 0, st>>>(a, k_chunk, part); break;
    case GEMM_NN: k_gemm<GEMM_NN><<<grid, GEMM_THREADS, 0, st>>>(a, k_chunk, part); break;
    case GEMM_TN: k_gemm<GEMM_TN><<<grid, GEMM_THREADS, 0, st>>>(a, k_chunk, part); break;
  }
  YOLAT_CHECK_LAUNCH();
  if (ksplit > 1) {
    int64_t tot = (int64_t)a.M * a.N;
    k_splitk_reduce<<<(unsigned)cdiv(tot, 256), 256, 0, st>>>(part, ksplit, a.M, a.N, a.bias, a.C, a.ldc, a.accumulate);
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

}  // namespace yolat
