// bn.cu -- training/eval BatchNorm1d (+ReLU) forward statistics, apply, and backward over [M, C] row-major.
//
// Semantics: nn.BatchNorm1d(nc, affine=True) as built by gcn_lib/sparse/torch_nn.py:23-34 -- batch mean and
// BIASED variance normalise in training mode, running buffers are updated with momentum 0.1 and the
// UNBIASED variance; eval mode uses the running buffers.  Reductions are two-stage and deterministic:
// fp32 partial sums per CTA (<= 128 addends per thread), combined in fp64.
#include "common.cuh"

namespace yolat {

constexpr int ST_COLS = 32;   // columns per CTA (one warp-width => 128-byte coalesced row segments)
constexpr int ST_ROWS = 8;    // row lanes per CTA
constexpr int ST_ROWS_PER_CTA = 256;   // short per-CTA row loops: these reductions are latency-bound, not bandwidth-bound

// part layout: [nparts][2][C]
__global__ void __launch_bounds__(ST_COLS* ST_ROWS) k_colstats(const float* __restrict__ z, int64_t ldz, int64_t M, int C,
                                                                float* __restrict__ part) {
  __shared__ float s1[ST_ROWS][ST_COLS], s2[ST_ROWS][ST_COLS];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * ST_COLS + tx;
  const int64_t r0 = (int64_t)blockIdx.y * ST_ROWS_PER_CTA;
  const int64_t r1 = min(M, r0 + ST_ROWS_PER_CTA);
  float a = 0.f, b = 0.f;
  if (c < C) {
#pragma unroll 8
    for (int64_t r = r0 + ty; r < r1; r += ST_ROWS) {
      float v = z[r * ldz + c];
      a += v;
      b = fmaf(v, v, b);
    }
  }
  s1[ty][tx] = a; s2[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int q = 1; q < ST_ROWS; ++q) { a += s1[q][tx]; b += s2[q][tx]; }
    part[((int64_t)blockIdx.y * 2 + 0) * C + c] = a;
    part[((int64_t)blockIdx.y * 2 + 1) * C + c] = b;
  }
}

// Sum part[p][w][C] over p (fp64) for w = 0,1 with a 32 x 32 thread block: tx = channel (coalesced 128-byte rows),
// ty strides over the partials.  The result is valid in the ty == 0 threads.
__device__ __forceinline__ void reduce_parts2(const float* __restrict__ part, int nparts, int C, int c, double& s, double& t) {
  __shared__ double sm[2][32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  double a = 0.0, b = 0.0;
  if (c < C) {
    double a1 = 0.0, b1 = 0.0, a2 = 0.0, b2 = 0.0, a3 = 0.0, b3 = 0.0;
    int p = ty;
    for (; p + 96 < nparts; p += 128) {     // 8 independent loads in flight per thread
      const float x0 = part[((int64_t)p * 2 + 0) * C + c], y0 = part[((int64_t)p * 2 + 1) * C + c];
      const float x1 = part[((int64_t)(p + 32) * 2 + 0) * C + c], y1 = part[((int64_t)(p + 32) * 2 + 1) * C + c];
      const float x2 = part[((int64_t)(p + 64) * 2 + 0) * C + c], y2 = part[((int64_t)(p + 64) * 2 + 1) * C + c];
      const float x3 = part[((int64_t)(p + 96) * 2 + 0) * C + c], y3 = part[((int64_t)(p + 96) * 2 + 1) * C + c];
      a += (double)x0; b += (double)y0; a1 += (double)x1; b1 += (double)y1;
      a2 += (double)x2; b2 += (double)y2; a3 += (double)x3; b3 += (double)y3;
    }
    for (; p < nparts; p += 32) {
      a += (double)part[((int64_t)p * 2 + 0) * C + c];
      b += (double)part[((int64_t)p * 2 + 1) * C + c];
    }
    a = (a + a1) + (a2 + a3);
    b = (b + b1) + (b2 + b3);
  }
  sm[0][ty][tx] = a; sm[1][ty][tx] = b;
  __syncthreads();
  if (ty == 0) {      // fixed order, four independent chains per sum (the loads pipeline instead of serialising)
    double a1 = 0.0, a2 = 0.0, a3 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
    a = sm[0][0][tx]; b = sm[1][0][tx];
#pragma unroll
    for (int q = 1; q < 8; ++q) { a += sm[0][q][tx]; b += sm[1][q][tx]; }
#pragma unroll
    for (int q = 8; q < 16; ++q) { a1 += sm[0][q][tx]; b1 += sm[1][q][tx]; }
#pragma unroll
    for (int q = 16; q < 24; ++q) { a2 += sm[0][q][tx]; b2 += sm[1][q][tx]; }
#pragma unroll
    for (int q = 24; q < 32; ++q) { a3 += sm[0][q][tx]; b3 += sm[1][q][tx]; }
    a = (a + a1) + (a2 + a3);
    b = (b + b1) + (b2 + b3);
  }
  s = a; t = b;
}

__global__ void __launch_bounds__(1024) k_bn_finalize(const float* __restrict__ part, int nparts, int64_t M, int C,
                                                      yolat_bn bn, int training, float* __restrict__ stat) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0 && training && bn.num_batches_tracked)
    *bn.num_batches_tracked += 1;
  // the affine parameters and running buffers do not depend on the reduction: request them first so that their latency
  // hides behind the partial sums (this kernel is a chain of dependent round trips, not a bandwidth problem)
  const bool fin = threadIdx.y == 0 && c < C;
  float w = 0.f, b = 0.f, rm = 0.f, rv = 0.f;
  if (fin) {
    w = bn.w[c];
    b = bn.b[c];
    if (bn.running_mean) rm = bn.running_mean[c];
    if (bn.running_var) rv = bn.running_var[c];
  }
  double s = 0.0, ss = 0.0;
  if (training) reduce_parts2(part, nparts, C, c, s, ss);
  if (!fin) return;
  float mean, var;
  if (training) {
    const double mu = s / (double)M;
    double v = ss / (double)M - mu * mu;
    if (v < 0.0) v = 0.0;
    mean = (float)mu;
    var = (float)v;
    if (bn.running_mean) bn.running_mean[c] = (1.f - kBnMomentum) * rm + kBnMomentum * mean;
    if (bn.running_var) {
      const double unbiased = M > 1 ? v * (double)M / (double)(M - 1) : v;
      bn.running_var[c] = (1.f - kBnMomentum) * rv + kBnMomentum * (float)unbiased;
    }
  } else {
    mean = rm;
    var = rv;
  }
  const float invstd = 1.0f / sqrtf(var + kBnEps);
  const float sc = w * invstd;
  stat[c] = sc;
  stat[C + c] = b - mean * sc;
  stat[2 * C + c] = mean;
  stat[3 * C + c] = invstd;
}

int bn_finalize_from_partials(const float* part, int nparts, int64_t M, int C, const yolat_bn* bn, int training,
                              float* stat, cudaStream_t st) {
  k_bn_finalize<<<(unsigned)cdiv(C, 32), dim3(32, 32), 0, st>>>(part, nparts, M, C, *bn, training, stat);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

// z = x W^T + b through the tensor-core GEMM with the BatchNorm statistics taken from the GEMM's own epilogue
// (per-tile column sums) when the kernel produced them, otherwise from a separate pass over z.
int linear_bn_stats(const GemmArgs& a, Arena& ws, const yolat_bn* bn, int training, float* stat, cudaStream_t st) {
  float* part = nullptr;
  int np = 0;
  if (training) {
    YOLAT_TRY(gemm_stats(a, GEMM_NT, ws, &part, &np, st));
  } else {
    YOLAT_TRY(gemm(a, GEMM_NT, ws, st));
  }
  if (training && np > 0) {
    if (ws.dry()) return YOLAT_OK;
    return bn_finalize_from_partials(part, np, a.M, a.N, bn, training, stat, st);
  }
  return bn_forward_stats(a.C, a.ldc, a.M, a.N, bn, training, stat, ws, st);
}

int bn_forward_stats(const float* z, int64_t ldz, int64_t M, int C, const yolat_bn* bn, int training, float* stat,
                     Arena& ws, cudaStream_t st) {
  const int nparts = training ? (int)cdiv(M > 0 ? M : 1, ST_ROWS_PER_CTA) : 0;
  float* part = training ? ws.take((int64_t)nparts * 2 * C) : nullptr;
  if (ws.dry()) return YOLAT_OK;
  if (ws.overflow) return YOLAT_ERR_WORKSPACE;
  if (training) {
    dim3 grid((unsigned)cdiv(C, ST_COLS), nparts);
    k_colstats<<<grid, ST_COLS * ST_ROWS, 0, st>>>(z, ldz, M, C, part);
    YOLAT_CHECK_LAUNCH();
  }
  return bn_finalize_from_partials(part, nparts, M, C, bn, training, stat, st);
}

// ---- apply -------------------------------------------------------------------------------------
__global__ void k_bn_apply(const float* __restrict__ z, int64_t ldz, int64_t M, int C, const float* __restrict__ stat,
                           int relu, float* __restrict__ y, int64_t ldy) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= M * C) return;
  const int64_t r = idx / C;
  const int c = (int)(idx % C);
  float v = fmaf(z[r * ldz + c], stat[c], stat[C + c]);
  if (relu) v = fmaxf(v, 0.f);
  y[r * ldy + c] = v;
}

__global__ void k_bn_apply4(const float* __restrict__ z, int64_t ldz, int64_t M, int C4, const float* __restrict__ stat,
                            int relu, float* __restrict__ y, int64_t ldy) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= M * C4) return;
  const int64_t r = idx / C4;
  const int c = (int)(idx % C4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(z + r * ldz + c);
  const float4 sc = __ldg(reinterpret_cast<const float4*>(stat + c));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(stat + C4 * 4 + c));
  float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  *reinterpret_cast<float4*>(y + r * ldy + c) = o;
}

int bn_apply(const float* z, int64_t ldz, int64_t M, int C, const float* stat, int relu, float* y, int64_t ldy,
             cudaStream_t st) {
  if (M * C <= 0) return YOLAT_OK;
  const bool vec = (C % 4 == 0) && (ldz % 4 == 0) && (ldy % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(stat)) & 15u) == 0;
  if (vec) {
    k_bn_apply4<<<(unsigned)cdiv(M * (C / 4), 256), 256, 0, st>>>(z, ldz, M, C / 4, stat, relu, y, ldy);
  } else {
    k_bn_apply<<<(unsigned)cdiv(M * C, 256), 256, 0, st>>>(z, ldz, M, C, stat, relu, y, ldy);
  }
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

// ---- backward ----------------------------------------------------------------------------------
__device__ __forceinline__ float bwd_load_dy(const BnBwdArgs& a, int64_t r, int c) {
  int64_t row = a.row_idx ? (int64_t)a.row_idx[r] : r;
  float v = a.gy[row * a.ldgy + c];
  if (a.row_scale) v *= a.row_scale[row];
  if (a.slot_scale) v *= a.slot_scale[a.slot_idx ? a.slot_idx[r] : r];
  return v;
}

// partial sums of dy' and dy'*xhat per column.  part: [nparts][2][C]
__global__ void __launch_bounds__(ST_COLS* ST_ROWS) k_bn_bwd_partial(BnBwdArgs a, float* __restrict__ part) {
  __shared__ float s1[ST_ROWS][ST_COLS], s2[ST_ROWS][ST_COLS];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * ST_COLS + tx;
  const int64_t r0 = (int64_t)blockIdx.y * ST_ROWS_PER_CTA;
  const int64_t r1 = min(a.M, r0 + ST_ROWS_PER_CTA);
  float p = 0.f, q = 0.f;
  if (c < a.C) {
    const float sc = a.stat[c], sh = a.stat[a.C + c], mean = a.stat[2 * a.C + c], invstd = a.stat[3 * a.C + c];
#pragma unroll 8
    for (int64_t r = r0 + ty; r < r1; r += ST_ROWS) {
      const float zz = a.z[r * a.ldz + c];
      float dy = bwd_load_dy(a, r, c);
      if (a.relu && !(fmaf(zz, sc, sh) > 0.f)) dy = 0.f;
      p += dy;
      q = fmaf(dy, (zz - mean) * invstd, q);
    }
  }
  s1[ty][tx] = p; s2[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && c < a.C) {
#pragma unroll
    for (int k = 1; k < ST_ROWS; ++k) { p += s1[k][tx]; q += s2[k][tx]; }
    part[((int64_t)blockIdx.y * 2 + 0) * a.C + c] = p;
    part[((int64_t)blockIdx.y * 2 + 1) * a.C + c] = q;
  }
}

// bstat: [0] m1 = sum(dy')/M, [1] m2 = sum(dy'*xhat)/M  (both 0 in eval mode: statistics are constants)
__global__ void __launch_bounds__(1024) k_bn_bwd_finalize(const float* __restrict__ part, int nparts, int64_t M, int C,
                                                          const float* __restrict__ stat, const float* __restrict__ gamma,
                                                          int training, float* __restrict__ bstat,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                          float* __restrict__ dbias) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s, t;
  reduce_parts2(part, nparts, C, c, s, t);
  if (threadIdx.y != 0 || c >= C) return;
  if (dgamma) dgamma[c] = (float)t;
  if (dbeta) dbeta[c] = (float)s;
  if (training) {
    bstat[c] = (float)(s / (double)M);
    bstat[C + c] = (float)(t / (double)M);
    // sum_m dz[m,c] = gamma*invstd*(S - M*m1 - m2*sum(xhat)) == 0 exactly: a bias feeding a training-mode BN
    if (dbias) dbias[c] = 0.f;
  } else {
    bstat[c] = 0.f;
    bstat[C + c] = 0.f;
    if (dbias) dbias[c] = (float)(s * (double)stat[c]);   // dz = dy' * sc
  }
}

__global__ void k_bn_bwd_apply(BnBwdArgs a, const float* __restrict__ bstat) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= a.M * a.C) return;
  const int64_t r = idx / a.C;
  const int c = (int)(idx % a.C);
  const float sc = a.stat[c], sh = a.stat[a.C + c], mean = a.stat[2 * a.C + c], invstd = a.stat[3 * a.C + c];
  const float zz = a.z[r * a.ldz + c];
  float dy = bwd_load_dy(a, r, c);
  if (a.relu && !(fmaf(zz, sc, sh) > 0.f)) dy = 0.f;
  const float xhat = (zz - mean) * invstd;
  a.dz[r * a.lddz + c] = sc * (dy - bstat[c] - xhat * bstat[a.C + c]);
}

// ---- four columns per thread (C % 4 == 0, 16-byte aligned rows): 128-bit accesses, one row
//      look-up per thread instead of per element ----------------------------------------------------------------------
template <bool NC>   // NC: gy is read-only for the whole kernel (the apply pass may run in place on gy)
__device__ __forceinline__ float4 bwd_load_dy4(const BnBwdArgs& a, int64_t r, int c) {
  const int64_t row = a.row_idx ? (int64_t)__ldg(a.row_idx + r) : r;
  const float4* gp = reinterpret_cast<const float4*>(a.gy + row * a.ldgy + c);
  float4 v = NC ? __ldg(gp) : *gp;
  float f = 1.f;
  if (a.row_scale) f = __ldg(a.row_scale + row);
  if (a.slot_scale) f *= __ldg(a.slot_scale + (a.slot_idx ? __ldg(a.slot_idx + r) : r));
  if (a.row_scale || a.slot_scale) { v.x *= f; v.y *= f; v.z *= f; v.w *= f; }
  return v;
}

// The per-thread chain (row index -> gathered gradient row) is latency-bound: four rows per thread, many CTAs.
constexpr int BWD4_ROWS_PER_CTA = 64;
// part: [nparts][2][C]; 256 threads = 16 row lanes x 16 column groups (64 columns), grid (ceil(C4 / 16), nparts)
__global__ void __launch_bounds__(256) k_bn_bwd_partial4(BnBwdArgs a, int C4, float* __restrict__ part) {
  __shared__ float4 s1[256], s2[256];
  const int cg = blockIdx.x * 16 + (threadIdx.x & 15), rl = threadIdx.x >> 4;
  const int c = cg * 4;
  const int64_t r0 = (int64_t)blockIdx.y * BWD4_ROWS_PER_CTA;
  const int64_t r1 = min(a.M, r0 + BWD4_ROWS_PER_CTA);
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), q = p;
  if (cg < C4) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(a.stat + c)), sh = __ldg(reinterpret_cast<const float4*>(a.stat + a.C + c));
    const float4 mean = __ldg(reinterpret_cast<const float4*>(a.stat + 2 * a.C + c));
    const float4 invstd = __ldg(reinterpret_cast<const float4*>(a.stat + 3 * a.C + c));
#pragma unroll 4
    for (int64_t r = r0 + rl; r < r1; r += 16) {
      const float4 zz = __ldg(reinterpret_cast<const float4*>(a.z + r * a.ldz + c));
      float4 dy = bwd_load_dy4<true>(a, r, c);
      if (a.relu) {
        if (!(fmaf(zz.x, sc.x, sh.x) > 0.f)) dy.x = 0.f;
        if (!(fmaf(zz.y, sc.y, sh.y) > 0.f)) dy.y = 0.f;
        if (!(fmaf(zz.z, sc.z, sh.z) > 0.f)) dy.z = 0.f;
        if (!(fmaf(zz.w, sc.w, sh.w) > 0.f)) dy.w = 0.f;
      }
      p.x += dy.x; p.y += dy.y; p.z += dy.z; p.w += dy.w;
      q.x = fmaf(dy.x, (zz.x - mean.x) * invstd.x, q.x);
      q.y = fmaf(dy.y, (zz.y - mean.y) * invstd.y, q.y);
      q.z = fmaf(dy.z, (zz.z - mean.z) * invstd.z, q.z);
      q.w = fmaf(dy.w, (zz.w - mean.w) * invstd.w, q.w);
    }
  }
  s1[threadIdx.x] = p; s2[threadIdx.x] = q;
  __syncthreads();
  if (rl == 0 && cg < C4) {
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      const float4 u = s1[k * 16 + (threadIdx.x & 15)], v = s2[k * 16 + (threadIdx.x & 15)];
      p.x += u.x; p.y += u.y; p.z += u.z; p.w += u.w;
      q.x += v.x; q.y += v.y; q.z += v.z; q.w += v.w;
    }
    *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.y * 2 + 0) * a.C + c) = p;
    *reinterpret_cast<float4*>(part + ((int64_t)blockIdx.y * 2 + 1) * a.C + c) = q;
  }
}

__global__ void k_bn_bwd_apply4(BnBwdArgs a, int C4, const float* __restrict__ bstat) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= a.M * C4) return;
  const int64_t r = idx / C4;
  const int c = (int)(idx - r * C4) * 4;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(a.stat + c)), sh = __ldg(reinterpret_cast<const float4*>(a.stat + a.C + c));
  const float4 mean = __ldg(reinterpret_cast<const float4*>(a.stat + 2 * a.C + c));
  const float4 invstd = __ldg(reinterpret_cast<const float4*>(a.stat + 3 * a.C + c));
  const float4 m1 = __ldg(reinterpret_cast<const float4*>(bstat + c)), m2 = __ldg(reinterpret_cast<const float4*>(bstat + a.C + c));
  const float4 zz = *reinterpret_cast<const float4*>(a.z + r * a.ldz + c);     // may alias dz: plain load
  float4 dy = bwd_load_dy4<false>(a, r, c);
  if (a.relu) {
    if (!(fmaf(zz.x, sc.x, sh.x) > 0.f)) dy.x = 0.f;
    if (!(fmaf(zz.y, sc.y, sh.y) > 0.f)) dy.y = 0.f;
    if (!(fmaf(zz.z, sc.z, sh.z) > 0.f)) dy.z = 0.f;
    if (!(fmaf(zz.w, sc.w, sh.w) > 0.f)) dy.w = 0.f;
  }
  *reinterpret_cast<float4*>(a.dz + r * a.lddz + c) =
      make_float4(sc.x * (dy.x - m1.x - (zz.x - mean.x) * invstd.x * m2.x), sc.y * (dy.y - m1.y - (zz.y - mean.y) * invstd.y * m2.y),
                  sc.z * (dy.z - m1.z - (zz.z - mean.z) * invstd.z * m2.z), sc.w * (dy.w - m1.w - (zz.w - mean.w) * invstd.w * m2.w));
}

static bool bn_bwd_vec_ok(const BnBwdArgs& a) {
  if (a.C % 4 != 0) return false;
  if ((a.ldz | a.ldgy | (a.dz ? a.lddz : 0)) % 4 != 0) return false;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(a.z) | reinterpret_cast<uintptr_t>(a.gy) | reinterpret_cast<uintptr_t>(a.stat) |
                         reinterpret_cast<uintptr_t>(a.dz);
  return (bits & 15u) == 0;
}

int bn_bwd_finalize(const float* part, int nparts, int64_t M, int C, const float* stat, const float* gamma, int training,
                    float* bstat, float* dgamma, float* dbeta, float* dbias, cudaStream_t st) {
  k_bn_bwd_finalize<<<(unsigned)cdiv(C, 32), dim3(32, 32), 0, st>>>(part, nparts, M, C, stat, gamma, training, bstat, dgamma,
                                                            dbeta, dbias);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int bn_backward(const BnBwdArgs& a, Arena& ws, cudaStream_t st) {
  // (the workspace is sized for the finer of the two partitions so that the dry run needs no pointers)
  const int nparts_max = (int)cdiv(a.M > 0 ? a.M : 1, BWD4_ROWS_PER_CTA);
  float* part = ws.take((int64_t)nparts_max * 2 * a.C);
  float* bstat = ws.take(2 * a.C);
  if (ws.dry()) return YOLAT_OK;
  if (ws.overflow) return YOLAT_ERR_WORKSPACE;
  const bool vec = bn_bwd_vec_ok(a) && nparts_max <= 65535 &&
                   ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(bstat)) & 15u) == 0;
  const int nparts = vec ? nparts_max : (int)cdiv(a.M > 0 ? a.M : 1, ST_ROWS_PER_CTA);
  if (vec) {
    k_bn_bwd_partial4<<<dim3((unsigned)cdiv(a.C / 4, 16), nparts), 256, 0, st>>>(a, a.C / 4, part);
  } else {
    dim3 grid((unsigned)cdiv(a.C, ST_COLS), nparts);
    k_bn_bwd_partial<<<grid, ST_COLS * ST_ROWS, 0, st>>>(a, part);
  }
  YOLAT_CHECK_LAUNCH();
  YOLAT_TRY(bn_bwd_finalize(part, nparts, a.M, a.C, a.stat, a.gamma, a.training, bstat, a.dgamma, a.dbeta, a.dbias, st));
  if (a.M * a.C > 0 && a.dz) {
    if (vec) k_bn_bwd_apply4<<<(unsigned)cdiv(a.M * (a.C / 4), 256), 256, 0, st>>>(a, a.C / 4, bstat);
    else k_bn_bwd_apply<<<(unsigned)cdiv(a.M * a.C, 256), 256, 0, st>>>(a, bstat);
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

// ---- misc --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_COLS* ST_ROWS) k_colsum_partial(const float* __restrict__ z, int64_t ldz, int64_t M,
                                                                      int C, float* __restrict__ part) {
  __shared__ float s1[ST_ROWS][ST_COLS];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * ST_COLS + tx;
  const int64_t r0 = (int64_t)blockIdx.y * ST_ROWS_PER_CTA;
  const int64_t r1 = min(M, r0 + ST_ROWS_PER_CTA);
  float a = 0.f;
  if (c < C) {
#pragma unroll 8
    for (int64_t r = r0 + ty; r < r1; r += ST_ROWS) a += z[r * ldz + c];
  }
  s1[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int q = 1; q < ST_ROWS; ++q) a += s1[q][tx];
    part[(int64_t)blockIdx.y * C + c] = a;
  }
}
__global__ void __launch_bounds__(1024) k_colsum_final(const float* __restrict__ part, int nparts, int C,
                                                       float* __restrict__ out) {
  __shared__ double sm[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  double s = 0.0;
  if (c < C)
    for (int p = ty; p < nparts; p += 32) s += (double)part[(int64_t)p * C + c];
  sm[ty][tx] = s;
  __syncthreads();
  if (ty != 0 || c >= C) return;
  for (int q = 1; q < 32; ++q) s += sm[q][tx];
  out[c] = (float)s;
}

int colsum(const float* a, int64_t lda, int64_t M, int C, float* out, Arena& ws, cudaStream_t st) {
  const int nparts = (int)cdiv(M > 0 ? M : 1, ST_ROWS_PER_CTA);
  float* part = ws.take((int64_t)nparts * C);
  if (ws.dry()) return YOLAT_OK;
  if (ws.overflow) return YOLAT_ERR_WORKSPACE;
  dim3 grid((unsigned)cdiv(C, ST_COLS), nparts);
  k_colsum_partial<<<grid, ST_COLS * ST_ROWS, 0, st>>>(a, lda, M, C, part);
  YOLAT_CHECK_LAUNCH();
  k_colsum_final<<<(unsigned)cdiv(C, 32), dim3(32, 32), 0, st>>>(part, nparts, C, out);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

__global__ void k_fill_zero(float* __restrict__ p, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}
int fill_zero(float* p, int64_t n, cudaStream_t st) {
  if (n <= 0) return YOLAT_OK;
  cudaMemsetAsync(p, 0, n * sizeof(float), st);
  return YOLAT_OK;
}

}  // namespace yolat
