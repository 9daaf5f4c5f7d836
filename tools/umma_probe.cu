// umma_probe.cu -- hardware probe of the tcgen05 kind::f16 (bf16) shared-memory descriptor conventions the fused K-EDGE
// backward relies on.  Stand-alone test program (not part of libyolat_b200.so):
//     nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I yolat_vectorgraphicsrecognition_b200/csrc tools/umma_probe.cu -o gpurun_out/umma_probe
// One tile pair in shared memory, both stored as [row][64 bf16] (128-byte rows, 8-row SWIZZLE_128B atoms); the same
// bytes are read
//   T1  K-major A (128 x 64) x K-major B (64 x 64)                 D[m][n] = sum_k X[m][k] Y[n][k]
//   T2  K-major A            x MN-major B (rows = k, n contiguous) D[m][n] = sum_k X[m][k] Y[k][n]
//   T3  MN-major A, M = 128 stacked from two tiles through LBO, MN-major B, K = 128 rows
//                                                                  D[m][n] = sum_s Z[m / 64][s][m % 64] X[s][n]
// for a few (LBO, SBO, k-step advance) candidates; prints the max abs error of each against a double reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "tc.cuh"

using namespace yolat::tc;

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

struct Variant {
  int a_mn, b_mn;            // major-ness flags of the instruction descriptor
  uint32_t a_lbo, a_sbo, a_adv, a_base;   // bytes; a_base: 0 = tile X, 1 = tile Z0 (stack of Z0 | Z1)
  uint32_t b_lbo, b_sbo, b_adv, b_base;   // b_base: 0 = tile Y, 1 = tile X
  int ksteps;
};

// tiles: X [128][64], Y [64][64] (padded to 128 rows), Z0 [128][64], Z1 [128][64]: 16 KB each
__global__ void __launch_bounds__(128, 1) k_probe(const float* __restrict__ X, const float* __restrict__ Y,
                                                 const float* __restrict__ Z0, const float* __restrict__ Z1,
                                                 Variant v, float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_u32 = raw + pad;
  const int tid = threadIdx.x, warp = tid >> 5;
  const float* srcs[4] = {X, Y, Z0, Z1};
  for (int t = 0; t < 4; ++t) {
    for (int idx = tid; idx < 128 * 8; idx += 128) {
      const int row = idx >> 3, ch = idx & 7;
      __nv_bfloat16 b[8];
      for (int j = 0; j < 8; ++j) b[j] = __float2bfloat16(srcs[t][row * 64 + ch * 8 + j]);
      const uint32_t off = (uint32_t)t * 16384u + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
                           (uint32_t)((ch ^ (row & 7)) << 4);
      *reinterpret_cast<uint4*>(sm + off) = *reinterpret_cast<const uint4*>(b);
    }
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = idesc_bf16(128, 64, v.a_mn, v.b_mn);
      const uint32_t a0 = sm_u32 + (v.a_base ? 2u * 16384u : 0u);
      const uint32_t b0 = sm_u32 + (v.b_base ? 0u : 16384u);
      for (int ks = 0; ks < v.ksteps; ++ks) {
        const uint64_t ad = make_desc(a0 + (uint32_t)ks * v.a_adv, v.a_lbo, v.a_sbo, LAYOUT_SW128);
        const uint64_t bd = make_desc(b0 + (uint32_t)ks * v.b_adv, v.b_lbo, v.b_sbo, LAYOUT_SW128);
        umma_bf16(tmem_d, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0u);
  tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    float r[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * 32), r);
    for (int i = 0; i < 32; ++i) D[tid * 64 + h * 32 + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<float> X(128 * 64), Y(128 * 64, 0.f), Z0(128 * 64), Z1(128 * 64);
  srand(1);
  auto rnd = [] { return bf((float)rand() / RAND_MAX * 2.f - 1.f); };
  for (auto& v : X) v = rnd();
  for (int i = 0; i < 64 * 64; ++i) Y[i] = rnd();
  for (auto& v : Z0) v = rnd();
  for (auto& v : Z1) v = rnd();
  float *dX, *dY, *dZ0, *dZ1, *dD;
  cudaMalloc(&dX, 128 * 64 * 4); cudaMalloc(&dY, 128 * 64 * 4); cudaMalloc(&dZ0, 128 * 64 * 4); cudaMalloc(&dZ1, 128 * 64 * 4);
  cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dX, X.data(), 128 * 64 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dY, Y.data(), 128 * 64 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dZ0, Z0.data(), 128 * 64 * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dZ1, Z1.data(), 128 * 64 * 4, cudaMemcpyHostToDevice);
  const int SMEM = 4 * 16384 + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);

  std::vector<double> R1(128 * 64), R2(128 * 64), R3(128 * 64);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double s1 = 0, s2 = 0, s3 = 0;
      for (int k = 0; k < 64; ++k) { s1 += (double)X[m * 64 + k] * Y[n * 64 + k]; s2 += (double)X[m * 64 + k] * Y[k * 64 + n]; }
      const float* Z = m < 64 ? Z0.data() : Z1.data();
      for (int s = 0; s < 128; ++s) s3 += (double)Z[s * 64 + (m & 63)] * X[s * 64 + n];
      R1[m * 64 + n] = s1; R2[m * 64 + n] = s2; R3[m * 64 + n] = s3;
    }
  struct Case { const char* name; Variant v; const std::vector<double>* ref; };
  std::vector<Case> cases = {
      {"T1 K-major x K-major (lbo 16, sbo 1024, adv 32)", {0, 0, 16, 1024, 32, 0, 16, 1024, 32, 0, 4}, &R1},
      {"T2 B MN-major (lbo 8192, sbo 1024, adv 2048)", {0, 1, 16, 1024, 32, 0, 8192, 1024, 2048, 0, 4}, &R2},
      {"T2 B MN-major (lbo 1024, sbo 8192, adv 2048)", {0, 1, 16, 1024, 32, 0, 1024, 8192, 2048, 0, 4}, &R2},
      {"T2 B MN-major (lbo 16, sbo 1024, adv 2048)", {0, 1, 16, 1024, 32, 0, 16, 1024, 2048, 0, 4}, &R2},
      {"T3 A MN-major stacked (lbo 16384, sbo 1024, adv 2048), B MN (lbo 8192, sbo 1024)", {1, 1, 16384, 1024, 2048, 1, 8192, 1024, 2048, 1, 8}, &R3},
      {"T3 A MN-major stacked (lbo 1024, sbo 16384, adv 2048), B MN (lbo 1024, sbo 8192)", {1, 1, 1024, 16384, 2048, 1, 1024, 8192, 2048, 1, 8}, &R3},
      {"T3 A MN-major stacked (lbo 16384, sbo 1024), B MN (lbo 16, sbo 1024)", {1, 1, 16384, 1024, 2048, 1, 16, 1024, 2048, 1, 8}, &R3},
  };
  std::vector<float> D(128 * 64);
  int rc = 0;
  for (auto& c : cases) {
    cudaMemset(dD, 0, 128 * 64 * 4);
    k_probe<<<1, 128, SMEM>>>(dX, dY, dZ0, dZ1, c.v, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-90s CUDA error: %s\n", c.name, cudaGetErrorString(e)); rc = 1; break; }
    cudaMemcpy(D.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    double err = 0, err_lo = 0, err_hi = 0;
    for (int i = 0; i < 128 * 64; ++i) {
      const double d = fabs((double)D[i] - (*c.ref)[i]);
      err = fmax(err, d);
      if (i < 64 * 64) err_lo = fmax(err_lo, d); else err_hi = fmax(err_hi, d);
    }
    printf("%-90s max|err| = %.3e (rows 0-63: %.3e, rows 64-127: %.3e) %s\n", c.name, err, err_lo, err_hi, err < 1e-3 ? "OK" : "MISMATCH");
  }
  return rc;
}
