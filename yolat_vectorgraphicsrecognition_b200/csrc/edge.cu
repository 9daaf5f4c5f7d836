// edge.cu -- per-edge kernels of GraphConv('attr_edge_gp2') over the CSR-by-target slot order.
//
// Reference formulation (gcn_lib/sparse/torch_vertex.py:330-337):
//     f_e = [x_i, x_j - x_i, attr_e];  z1_e = W1 f_e + b1
// W1 = [W1a | W1b | W1c] along that concat, so z1_e = P[i_e] + Q[j_e] + W1c attr_e + b1 with
//     P = x (W1a - W1b)^T,  Q = x W1b^T          (node-level [N,Cin]x[Cin,2C] GEMM instead of an edge-level one)
// One warp owns one target row; lane l owns channels l*CPL .. l*CPL+CPL-1 (C = 32*CPL), so every row
// access is one fully coalesced 128*CPL-byte segment and the mean aggregation needs no atomics.
#include "common.cuh"

namespace yolat {

constexpr int EDGE_WARPS = 8;

template <int CPL> struct VecT;
template <> struct VecT<1> { using T = float; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<4> { using T = float4; };

template <int CPL>
__device__ __forceinline__ void ldv(const float* p, float (&v)[CPL]) {
  using V = typename VecT<CPL>::T;
  V t = *reinterpret_cast<const V*>(p);
  const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
  for (int q = 0; q < CPL; ++q) v[q] = f[q];
}
template <int CPL>
__device__ __forceinline__ void stv(float* p, const float (&v)[CPL]) {
  using V = typename VecT<CPL>::T;
  V t;
  float* f = reinterpret_cast<float*>(&t);
#pragma unroll
  for (int q = 0; q < CPL; ++q) f[q] = v[q];
  *reinterpret_cast<V*>(p) = t;
}

// Wpq [2C, Cin]: rows 0..C-1 = W1a - W1b, rows C..2C-1 = W1b.  With wr != null also rows 2C..3C-1 = Wr (lin_r) and
// bias3 [3C] = (0, 0, br): one GEMM then produces P | Q | lin_r(x).
__global__ void k_prep_wpq(const float* __restrict__ w1, int Cin, int C, float* __restrict__ wpq,
                           const float* __restrict__ wr, const float* __restrict__ br, float* __restrict__ bias3) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * Cin) return;
  const int c = idx / Cin, k = idx % Cin;
  const int ld = 2 * Cin + 4;
  const float a = w1[c * ld + k], b = w1[c * ld + Cin + k];
  wpq[c * Cin + k] = a - b;
  wpq[(C + c) * Cin + k] = b;
  if (wr) {
    wpq[(2 * C + c) * Cin + k] = wr[c * Cin + k];
    if (k == 0) {
      bias3[c] = 0.f;
      bias3[C + c] = 0.f;
      bias3[2 * C + c] = br ? br[c] : 0.f;
    }
  }
}

// dW1 [C, 2Cin+4] from dWpq [2C, Cin] and dW1c [C,4]:  dW1a = dWp, dW1b = dWq - dWp
__global__ void k_assemble_dw1(const float* __restrict__ dwpq, const float* __restrict__ dw1c, int Cin, int C,
                               float* __restrict__ dw1) {
  const int ld = 2 * Cin + 4;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * ld) return;
  const int c = idx / ld, k = idx % ld;
  float v;
  if (k < Cin) v = dwpq[c * Cin + k];
  else if (k < 2 * Cin) v = dwpq[(C + c) * Cin + (k - Cin)] - dwpq[c * Cin + (k - Cin)];
  else v = dw1c[c * 4 + (k - 2 * Cin)];
  dw1[idx] = v;
}

// z1[slot] = P[dst] + Q[src] + W1c attr[eid] + b1; per-CTA partial (sum, sumsq) per channel for BN1.
template <int CPL>
__global__ void __launch_bounds__(EDGE_WARPS * 32)
k_edge_z1(GraphView g, int64_t N, const float* __restrict__ pq, const float* __restrict__ attr,
          const float* __restrict__ w1, int Cin, const float* __restrict__ b1, float* __restrict__ z1,
          float* __restrict__ part) {
  constexpr int C = 32 * CPL;
  __shared__ float red[2][EDGE_WARPS][C];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c0 = lane * CPL;
  float w[CPL][4], bias[CPL], s[CPL], ss[CPL];
  const int ld = 2 * Cin + 4;
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
#pragma unroll
    for (int k = 0; k < 4; ++k) w[q][k] = w1[(c0 + q) * ld + 2 * Cin + k];
    bias[q] = b1 ? b1[c0 + q] : 0.f;
    s[q] = 0.f; ss[q] = 0.f;
  }
  for (int64_t row = (int64_t)blockIdx.x * EDGE_WARPS + wid; row < N; row += (int64_t)gridDim.x * EDGE_WARPS) {
    const int b = g.rowptr_t[row], e = g.rowptr_t[row + 1];
    if (b == e) continue;
    float p[CPL];
    ldv<CPL>(pq + row * (2 * C) + c0, p);
#pragma unroll
    for (int q = 0; q < CPL; ++q) p[q] += bias[q];
    for (int slot = b; slot < e; ++slot) {
      const int src = g.src_t[slot];
      const int eid = g.eid_t[slot];
      float qv[CPL];
      ldv<CPL>(pq + (int64_t)src * (2 * C) + C + c0, qv);
      const float4 at = *reinterpret_cast<const float4*>(attr + (int64_t)eid * 4);
      float z[CPL];
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        float v = p[q] + qv[q];
        v = fmaf(at.x, w[q][0], v);
        v = fmaf(at.y, w[q][1], v);
        v = fmaf(at.z, w[q][2], v);
        v = fmaf(at.w, w[q][3], v);
        z[q] = v;
        s[q] += v;
        ss[q] = fmaf(v, v, ss[q]);
      }
      if (z1) stv<CPL>(z1 + (int64_t)slot * C + c0, z);
    }
  }
  if (part == nullptr) return;
#pragma unroll
  for (int q = 0; q < CPL; ++q) { red[0][wid][c0 + q] = s[q]; red[1][wid][c0 + q] = ss[q]; }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * C; t += blockDim.x) {
    const int which = t / C, c = t % C;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < EDGE_WARPS; ++k) a += red[which][k][c];
    part[((int64_t)blockIdx.x * 2 + which) * C + c] = a;
  }
}

// out[v] = base[v] + deg_inv[v] * sum_{slot in row v} relu(bn2(z2[slot])) (* edge_weight[eid]); base may alias out
template <int CPL>
__global__ void __launch_bounds__(EDGE_WARPS * 32)
k_edge_agg(GraphView g, int64_t N, const float* __restrict__ z2, const float* __restrict__ stat2,
           const float* __restrict__ ew, const float* base, int64_t ldb, float* out, int64_t ldo) {
  constexpr int C = 32 * CPL;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c0 = lane * CPL;
  float sc[CPL], sh[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) { sc[q] = stat2[c0 + q]; sh[q] = stat2[C + c0 + q]; }
  for (int64_t row = (int64_t)blockIdx.x * EDGE_WARPS + wid; row < N; row += (int64_t)gridDim.x * EDGE_WARPS) {
    const int b = g.rowptr_t[row], e = g.rowptr_t[row + 1];
    if (b == e && base == out) continue;
    float acc[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) acc[q] = 0.f;
    for (int slot = b; slot < e; ++slot) {
      float z[CPL];
      ldv<CPL>(z2 + (int64_t)slot * C + c0, z);
      const float wgt = ew ? ew[g.eid_t[slot]] : 1.f;
#pragma unroll
      for (int q = 0; q < CPL; ++q) acc[q] += wgt * fmaxf(fmaf(z[q], sc[q], sh[q]), 0.f);
    }
    const float di = g.deg_inv[row];
    float bv[CPL];
    ldv<CPL>(base + row * ldb + c0, bv);
#pragma unroll
    for (int q = 0; q < CPL; ++q) bv[q] += acc[q] * di;
    stv<CPL>(out + row * ldo + c0, bv);
  }
}

// dPQ[v, 0:C] = sum over the target row of dz1; dPQ[v, C:2C] = sum over the source row of dz1;
// per-CTA partial of dW1c[c][k] = sum_slot dz1[slot,c] * attr[eid,k]   (part: [nblocks][C][4])
template <int CPL>
__global__ void __launch_bounds__(EDGE_WARPS * 32)
k_edge_bwd_scatter(GraphView g, int64_t N, const float* __restrict__ dz1, const float* __restrict__ attr,
                   float* __restrict__ dpq, float* __restrict__ part) {
  constexpr int C = 32 * CPL;
  __shared__ float red[EDGE_WARPS][C][4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c0 = lane * CPL;
  float wacc[CPL][4];
#pragma unroll
  for (int q = 0; q < CPL; ++q)
#pragma unroll
    for (int k = 0; k < 4; ++k) wacc[q][k] = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * EDGE_WARPS + wid; row < N; row += (int64_t)gridDim.x * EDGE_WARPS) {
    float dp[CPL], dq[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) { dp[q] = 0.f; dq[q] = 0.f; }
    for (int slot = g.rowptr_t[row]; slot < g.rowptr_t[row + 1]; ++slot) {
      float d[CPL];
      ldv<CPL>(dz1 + (int64_t)slot * C + c0, d);
      const float4 at = *reinterpret_cast<const float4*>(attr + (int64_t)g.eid_t[slot] * 4);
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        dp[q] += d[q];
        wacc[q][0] = fmaf(d[q], at.x, wacc[q][0]);
        wacc[q][1] = fmaf(d[q], at.y, wacc[q][1]);
        wacc[q][2] = fmaf(d[q], at.z, wacc[q][2]);
        wacc[q][3] = fmaf(d[q], at.w, wacc[q][3]);
      }
    }
    for (int k = g.rowptr_s[row]; k < g.rowptr_s[row + 1]; ++k) {
      float d[CPL];
      ldv<CPL>(dz1 + (int64_t)g.slot_s[k] * C + c0, d);
#pragma unroll
      for (int q = 0; q < CPL; ++q) dq[q] += d[q];
    }
    stv<CPL>(dpq + row * (2 * C) + c0, dp);
    stv<CPL>(dpq + row * (2 * C) + C + c0, dq);
  }
#pragma unroll
  for (int q = 0; q < CPL; ++q)
#pragma unroll
    for (int k = 0; k < 4; ++k) red[wid][c0 + q][k] = wacc[q][k];
  __syncthreads();
  for (int t = threadIdx.x; t < C * 4; t += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < EDGE_WARPS; ++k) a += (&red[k][0][0])[t];
    part[(int64_t)blockIdx.x * C * 4 + t] = a;
  }
}

// out[t] = sum_p part[p][t]   (fp64 combine); 32 x 32 block: tx = element, ty strides over the partials
__global__ void __launch_bounds__(1024) k_reduce_parts(const float* __restrict__ part, int nparts, int L,
                                                       float* __restrict__ out) {
  __shared__ double sm[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int t = blockIdx.x * 32 + tx;
  double s = 0.0;
  if (t < L)
    for (int p = ty; p < nparts; p += 32) s += (double)part[(int64_t)p * L + t];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty != 0 || t >= L) return;
  for (int q = 1; q < 32; ++q) s += sm[q][tx];
  out[t] = (float)s;
}

static int edge_grid(int64_t N) {
  int64_t need = cdiv(N, EDGE_WARPS);
  int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

int edge_prep_wpq(const float* w1, int Cin, int C, float* wpq, const float* wr, const float* br, float* bias3,
                  cudaStream_t st) {
  k_prep_wpq<<<(unsigned)cdiv(C * Cin, 256), 256, 0, st>>>(w1, Cin, C, wpq, wr, br, bias3);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int edge_assemble_dw1(const float* dwpq, const float* dw1c, int Cin, int C, float* dw1, cudaStream_t st) {
  k_assemble_dw1<<<(unsigned)cdiv(C * (2 * Cin + 4), 256), 256, 0, st>>>(dwpq, dw1c, Cin, C, dw1);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int edge_z1_nparts(int64_t N) { return edge_grid(N); }

int edge_z1(const GraphView& g, int64_t N, int C, const float* pq, const float* attr, const float* w1, int Cin,
            const float* b1, float* z1, float* part, cudaStream_t st) {
  const int grid = edge_grid(N);
  switch (C) {
    case 32: k_edge_z1<1><<<grid, EDGE_WARPS * 32, 0, st>>>(g, N, pq, attr, w1, Cin, b1, z1, part); break;
    case 64: k_edge_z1<2><<<grid, EDGE_WARPS * 32, 0, st>>>(g, N, pq, attr, w1, Cin, b1, z1, part); break;
    case 128: k_edge_z1<4><<<grid, EDGE_WARPS * 32, 0, st>>>(g, N, pq, attr, w1, Cin, b1, z1, part); break;
    default: return YOLAT_ERR_UNSUPPORTED;
  }
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int edge_agg(const GraphView& g, int64_t N, int C, const float* z2, const float* stat2, const float* ew,
             const float* base, int64_t ldb, float* out, int64_t ldo, cudaStream_t st) {
  const int grid = edge_grid(N);
  switch (C) {
    case 32: k_edge_agg<1><<<grid, EDGE_WARPS * 32, 0, st>>>(g, N, z2, stat2, ew, base, ldb, out, ldo); break;
    case 64: k_edge_agg<2><<<grid, EDGE_WARPS * 32, 0, st>>>(g, N, z2, stat2, ew, base, ldb, out, ldo); break;
    case 128: k_edge_agg<4><<<grid, EDGE_WARPS * 32, 0, st>>>(g, N, z2, stat2, ew, base, ldb, out, ldo); break;
    default: return YOLAT_ERR_UNSUPPORTED;
  }
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int edge_bwd_scatter(const GraphView& g, int64_t N, int C, const float* dz1, const float* attr, float* dpq,
                     float* part, float* dw1c, cudaStream_t st) {
  const int grid = edge_grid(N);
  switch (C) {
    case 32: k_edge_bwd_scatter<1><<<grid, EDGE_WARPS * 32, 0, st>>>(g, N, dz1, attr, dpq, part); break;
    case 64: k_edge_bwd_scatter<2><<<grid, EDGE_WARPS * 32, 0, st>>>(g, N, dz1, attr, dpq, part); break;
    case 128: k_edge_bwd_scatter<4><<<grid, EDGE_WARPS * 32, 0, st>>>(g, N, dz1, attr, dpq, part); break;
    default: return YOLAT_ERR_UNSUPPORTED;
  }
  YOLAT_CHECK_LAUNCH();
  k_reduce_parts<<<(unsigned)cdiv(C * 4, 32), dim3(32, 32), 0, st>>>(part, grid, C * 4, dw1c);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

}  // namespace yolat
