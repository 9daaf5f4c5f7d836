// umma_probe_m64.cu -- where do the 64 rows of an M = 64 tcgen05.mma (cta_group::1) accumulator live in TMEM?
//     D[m][n] = sum_k Y[m][k] X[n][k],  M = 64, N = 128, K = 64, kind::tf32 AND kind::f16 (bf16)
// All 128 TMEM lanes x 128 columns are dumped; the host matches every lane against the reference rows.
//     nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I yolat_vectorgraphicsrecognition_b200/csrc tools/umma_probe_m64.cu -o tools/build/umma_probe_m64
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "tc.cuh"

using namespace yolat::tc;

// X [128][64], Y [64][64]; TF32 = 1: fp32 tiles of two k-blocks (32 floats = 128 B rows), else bf16 rows of 64
template <int TF32>
__global__ void __launch_bounds__(128, 1) k_probe(const float* __restrict__ X, const float* __restrict__ Y, float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_u32 = raw + pad;
  const int tid = threadIdx.x, warp = tid >> 5;
  // X tile at 0 (32 KB as fp32: 2 k-blocks of 16 KB; 16 KB as bf16), Y tile at 32 KB
  for (int idx = tid; idx < 128 * 16; idx += 128) {
    const int row = idx >> 4, c4 = idx & 15;   // 4 floats
    for (int t = 0; t < 2; ++t) {
      if (t == 1 && row >= 64) continue;
      const float* src = (t ? Y : X) + row * 64 + c4 * 4;
      uint8_t* base = sm + (t ? 32768u : 0u);
      if (TF32) {
        const int kb = c4 >> 3, c = c4 & 7;
        const uint32_t off = (uint32_t)kb * (t ? 8192u : 16384u) + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
        *reinterpret_cast<float4*>(base + off) = *reinterpret_cast<const float4*>(src);
      } else {
        __nv_bfloat16 b[4];
        for (int j = 0; j < 4; ++j) b[j] = __float2bfloat16(src[j]);
        const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)(((c4 >> 1) ^ (row & 7)) << 4) + (uint32_t)(c4 & 1) * 8u;
        *reinterpret_cast<uint2*>(base + off) = *reinterpret_cast<const uint2*>(b);
      }
    }
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 128);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  // clear the accumulator region first so that untouched lanes read as a sentinel
  {
    for (int h = 0; h < 4; ++h) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                   ::"r"(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * 32)), "r"(__float_as_uint(-12345.f)) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = TF32 ? make_idesc(64, 128, 0, 0) : make_idesc_bf16(64, 128, 0, 0);
      const uint32_t x0 = sm_u32, y0 = sm_u32 + 32768u;
      if (TF32) {
        for (int kb = 0; kb < 2; ++kb)
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_desc(y0 + kb * 8192u + ks * 32u, 16, 1024, LAYOUT_SW128);
            const uint64_t bd = make_desc(x0 + kb * 16384u + ks * 32u, 16, 1024, LAYOUT_SW128);
            umma_tf32(tmem_d, ad, bd, idesc, (kb | ks) ? 1u : 0u);
          }
      } else {
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = make_desc(y0 + ks * 32u, 16, 1024, LAYOUT_SW128);
          const uint64_t bd = make_desc(x0 + ks * 32u, 16, 1024, LAYOUT_SW128);
          umma_bf16(tmem_d, ad, bd, idesc, ks ? 1u : 0u);
        }
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), 0u);
  tc_fence_after();
  for (int h = 0; h < 4; ++h) {
    float r[32];
    tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(h * 32), r);
    for (int i = 0; i < 32; ++i) D[tid * 128 + h * 32 + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 128);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  std::vector<float> X(128 * 64), Y(64 * 64);
  srand(2);
  for (auto& v : X) v = bf((float)rand() / RAND_MAX * 2.f - 1.f);
  for (auto& v : Y) v = bf((float)rand() / RAND_MAX * 2.f - 1.f);
  float *dX, *dY, *dD;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dY, Y.size() * 4); cudaMalloc(&dD, 128 * 128 * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dY, Y.data(), Y.size() * 4, cudaMemcpyHostToDevice);
  std::vector<double> R(64 * 128);
  for (int m = 0; m < 64; ++m)
    for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int k = 0; k < 64; ++k) s += (double)Y[m * 64 + k] * X[n * 64 + k];
      R[m * 128 + n] = s;
    }
  const int SMEM = 65536 + 1024;
  cudaFuncSetAttribute(k_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  std::vector<float> D(128 * 128);
  for (int tf = 1; tf >= 0; --tf) {
    cudaMemset(dD, 0, 128 * 128 * 4);
    if (tf) k_probe<1><<<1, 128, SMEM>>>(dX, dY, dD); else k_probe<0><<<1, 128, SMEM>>>(dX, dY, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("== kind::%s M=64 N=128: %s\n", tf ? "tf32" : "f16(bf16)", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(D.data(), dD, 128 * 128 * 4, cudaMemcpyDeviceToHost);
    int mapped = 0;
    for (int lane = 0; lane < 128; ++lane) {
      int best = -1; double best_err = 1e30;
      for (int m = 0; m < 64; ++m) {
        double err = 0;
        for (int n = 0; n < 128; ++n) err = fmax(err, fabs((double)D[lane * 128 + n] - R[m * 128 + n]));
        if (err < best_err) { best_err = err; best = m; }
      }
      const bool untouched = D[lane * 128] == -12345.f && D[lane * 128 + 127] == -12345.f;
      if (best_err < 1e-2) { printf("lane %3d <- row %2d (err %.1e)\n", lane, best, best_err); ++mapped; }
      else if (!untouched) printf("lane %3d: no row matches (best %d err %.2e), D[0..3] = %g %g %g %g\n", lane, best, best_err, D[lane * 128], D[lane * 128 + 1], D[lane * 128 + 2], D[lane * 128 + 3]);
    }
    printf("   %d lanes hold rows\n", mapped);
  }
  return 0;
}
