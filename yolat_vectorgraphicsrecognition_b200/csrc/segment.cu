// segment.cu -- torch_scatter.scatter(dim=0, reduce='mean'|'max') over prepared segments, the fused
// fusion_block -> scatter-max pooling, and CrossEntropyLoss.
//
// Reference call sites: cad_recognition/architecture3cc_rpn_gp_iter2.py:67 (mean over bbox_idx),
// :122 (max over bbox_idx), :363/:376 (CrossEntropyLoss).  Segments are contiguous row runs for the
// reference's data (bbox_idx is sorted, Datasets/graph_dict3.py:732) but the kernels go through
// perm[] so unsorted indices work too.  No atomics: one thread owns one (segment, column).
#include "common.cuh"

namespace yolat {

constexpr int SEG_T = 128;

__global__ void __launch_bounds__(SEG_T) k_segment_mean(const float* __restrict__ src, int64_t lds, int C, SegView sv,
                                                        float* __restrict__ out, int64_t ldo) {
  const int s = blockIdx.x;
  const int c = blockIdx.y * SEG_T + threadIdx.x;
  if (c >= C) return;
  const int b = sv.segptr[s], e = sv.segptr[s + 1];
  float acc = 0.f;
  for (int r = b; r < e; ++r) acc += src[(int64_t)sv.perm[r] * lds + c];
  out[(int64_t)s * ldo + c] = acc / (float)max(e - b, 1);
}

__global__ void __launch_bounds__(SEG_T) k_segment_mean_bwd(const float* __restrict__ g, int64_t ldg, int C, SegView sv,
                                                            float* __restrict__ dsrc, int64_t ldd, int accumulate) {
  const int s = blockIdx.x;
  const int c = blockIdx.y * SEG_T + threadIdx.x;
  if (c >= C) return;
  const int b = sv.segptr[s], e = sv.segptr[s + 1];
  const float v = g[(int64_t)s * ldg + c] / (float)max(e - b, 1);
  for (int r = b; r < e; ++r) {
    float* d = dsrc + (int64_t)sv.perm[r] * ldd + c;
    *d = accumulate ? (*d + v) : v;
  }
}

// out = max over the segment of t(src) where t = identity or relu(src*sc+sh); arg = first row attaining it
__global__ void __launch_bounds__(SEG_T) k_segment_max(const float* __restrict__ src, int64_t lds, int C, SegView sv,
                                                       const float* __restrict__ stat, float* __restrict__ out,
                                                       int64_t ldo, int32_t* __restrict__ arg, int64_t lda) {
  const int s = blockIdx.x;
  const int c = blockIdx.y * SEG_T + threadIdx.x;
  if (c >= C) return;
  const int b = sv.segptr[s], e = sv.segptr[s + 1];
  float best = 0.f;
  int32_t bi = -1;
  float sc = 1.f, sh = 0.f;
  if (stat) { sc = stat[c]; sh = stat[C + c]; }
  for (int r = b; r < e; ++r) {
    const int32_t row = sv.perm[r];
    float v = src[(int64_t)row * lds + c];
    if (stat) v = fmaxf(fmaf(v, sc, sh), 0.f);
    if (bi < 0 || v > best) { best = v; bi = row; }
  }
  out[(int64_t)s * ldo + c] = best;
  arg[(int64_t)s * lda + c] = bi;
}

// dsrc[arg[s,c], c] (+)= g[s,c]; rows of different segments are disjoint, so no atomics are needed.
__global__ void k_segment_max_bwd(const float* __restrict__ g, int64_t ldg, int C, int64_t S,
                                  const int32_t* __restrict__ arg, int64_t lda, float* __restrict__ dsrc, int64_t ldd) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= S * C) return;
  const int64_t s = idx / C;
  const int c = (int)(idx % C);
  const int32_t row = arg[s * lda + c];
  if (row >= 0) dsrc[(int64_t)row * ldd + c] += g[s * ldg + c];
}

__global__ void k_zero_strided(float* __restrict__ p, int64_t ld, int64_t M, int C) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= M * C) return;
  p[(idx / C) * ld + (idx % C)] = 0.f;
}

// ---- fusion backward: the post-ReLU gradient is sparse (one row per (segment, column)) --------------
// partial sums over segments of dy' and dy'*xhat; part [nparts][2][F]; FM_SEGS_PER_CTA segments per CTA
constexpr int FM_SEGS_PER_CTA = 32;    // 4 dependent (arg -> z) look-ups per thread: the chain is latency-bound, so spread it
__global__ void __launch_bounds__(256) k_fusemax_bwd_partial(const float* __restrict__ gp, int64_t ldg, int F, int64_t S,
                                                             const int32_t* __restrict__ arg, int64_t lda,
                                                             const float* __restrict__ z, const float* __restrict__ stat,
                                                             float* __restrict__ part) {
  __shared__ float s1[8][32], s2[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * FM_SEGS_PER_CTA, r1 = min(S, r0 + FM_SEGS_PER_CTA);
  float p = 0.f, q = 0.f;
  if (c < F) {
    const float sc = stat[c], sh = stat[F + c], mean = stat[2 * F + c], invstd = stat[3 * F + c];
    for (int64_t s = r0 + ty; s < r1; s += 8) {
      const int32_t row = arg[s * lda + c];
      if (row < 0) continue;
      const float zz = z[(int64_t)row * F + c];
      if (!(fmaf(zz, sc, sh) > 0.f)) continue;
      const float dy = gp[s * ldg + c];
      p += dy;
      q = fmaf(dy, (zz - mean) * invstd, q);
    }
  }
  s1[ty][tx] = p; s2[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && c < F) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { p += s1[k][tx]; q += s2[k][tx]; }
    part[((int64_t)blockIdx.y * 2 + 0) * F + c] = p;
    part[((int64_t)blockIdx.y * 2 + 1) * F + c] = q;
  }
}

// z[n,c] <- dz[n,c] = sc*(dy' - m1 - xhat*m2), dy' = g_pooled[seg(n),c] iff n is the arg row and the ReLU is active
__global__ void k_fusemax_bwd_apply(float* __restrict__ z, int64_t M, int F, const int32_t* __restrict__ seg_of_row,
                                    const float* __restrict__ gp, int64_t ldg, const int32_t* __restrict__ arg,
                                    int64_t lda, const float* __restrict__ stat, const float* __restrict__ bstat) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= M * F) return;
  const int64_t n = idx / F;
  const int c = (int)(idx % F);
  const float sc = stat[c], sh = stat[F + c], mean = stat[2 * F + c], invstd = stat[3 * F + c];
  const float zz = z[idx];
  const int32_t s = seg_of_row[n];
  float dy = 0.f;
  if (s >= 0 && arg[(int64_t)s * lda + c] == (int32_t)n && fmaf(zz, sc, sh) > 0.f) dy = gp[(int64_t)s * ldg + c];
  z[idx] = sc * (dy - bstat[c] - (zz - mean) * invstd * bstat[F + c]);
}

// Same, four columns x four rows per thread (F % 4 == 0, 16-byte aligned rows): 128-bit loads / stores, the six
// per-column parameter vectors are loaded once per thread and the four rows' loads are all in flight together.
constexpr int FMB_ROWS = 4;
__global__ void __launch_bounds__(256) k_fusemax_bwd_apply4(float* __restrict__ z, int64_t M, int F4,
                                                            const int32_t* __restrict__ seg_of_row,
                                                            const float* __restrict__ gp, int64_t ldg,
                                                            const int32_t* __restrict__ arg, int64_t lda,
                                                            const float* __restrict__ stat, const float* __restrict__ bstat) {
  const int cg = blockIdx.y * 64 + (threadIdx.x & 63);              // column group (4 columns)
  const int64_t n0 = ((int64_t)blockIdx.x * 4 + (threadIdx.x >> 6)) * FMB_ROWS;
  if (cg >= F4 || n0 >= M) return;
  const int c = cg * 4;
  const int F = F4 * 4;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(stat + c)), sh = __ldg(reinterpret_cast<const float4*>(stat + F + c));
  const float4 mean = __ldg(reinterpret_cast<const float4*>(stat + 2 * F + c));
  const float4 invstd = __ldg(reinterpret_cast<const float4*>(stat + 3 * F + c));
  const float4 m1 = __ldg(reinterpret_cast<const float4*>(bstat + c)), m2 = __ldg(reinterpret_cast<const float4*>(bstat + F + c));
  float4 zz[FMB_ROWS];
  int32_t sg[FMB_ROWS];
  int4 a[FMB_ROWS];
#pragma unroll
  for (int i = 0; i < FMB_ROWS; ++i) {
    const int64_t n = n0 + i;
    const bool in = n < M;
    zz[i] = in ? *reinterpret_cast<const float4*>(z + n * F + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    sg[i] = in ? __ldg(seg_of_row + n) : -1;
  }
#pragma unroll
  for (int i = 0; i < FMB_ROWS; ++i)
    a[i] = sg[i] >= 0 ? __ldg(reinterpret_cast<const int4*>(arg + (int64_t)sg[i] * lda + c)) : make_int4(-1, -1, -1, -1);
#pragma unroll
  for (int i = 0; i < FMB_ROWS; ++i) {
    const int64_t n = n0 + i;
    if (n >= M) break;
    const int32_t nn = (int32_t)n;
    float4 dy = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a[i].x == nn || a[i].y == nn || a[i].z == nn || a[i].w == nn) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gp + (int64_t)sg[i] * ldg + c));
      if (a[i].x == nn && fmaf(zz[i].x, sc.x, sh.x) > 0.f) dy.x = g.x;
      if (a[i].y == nn && fmaf(zz[i].y, sc.y, sh.y) > 0.f) dy.y = g.y;
      if (a[i].z == nn && fmaf(zz[i].z, sc.z, sh.z) > 0.f) dy.z = g.z;
      if (a[i].w == nn && fmaf(zz[i].w, sc.w, sh.w) > 0.f) dy.w = g.w;
    }
    *reinterpret_cast<float4*>(z + n * F + c) =
        make_float4(sc.x * (dy.x - m1.x - (zz[i].x - mean.x) * invstd.x * m2.x), sc.y * (dy.y - m1.y - (zz[i].y - mean.y) * invstd.y * m2.y),
                    sc.z * (dy.z - m1.z - (zz[i].z - mean.z) * invstd.z * m2.z), sc.w * (dy.w - m1.w - (zz[i].w - mean.w) * invstd.w * m2.w));
  }
}

// ---- CrossEntropyLoss (mean) -----------------------------------------------------------------------
// one warp per row: prob = softmax(logits), row_loss = logsumexp - logit[label]
__global__ void k_xent_rows(const float* __restrict__ logits, int64_t ldl, int64_t B, int ncls,
                            const int64_t* __restrict__ labels, float* __restrict__ prob, float* __restrict__ row_loss) {
  const int lane = threadIdx.x & 31;
  const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= B) return;
  const float* l = logits + r * ldl;
  float mx = -INFINITY;
  for (int c = lane; c < ncls; c += 32) mx = fmaxf(mx, l[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
  for (int c = lane; c < ncls; c += 32) se += expf(l[c] - mx);
  se = warp_sum(se);
  const float lse = logf(se) + mx;
  for (int c = lane; c < ncls; c += 32) prob[r * ncls + c] = expf(l[c] - lse);
  if (lane == 0) {
    const int64_t y = labels[r];
    row_loss[r] = (y >= 0 && y < ncls) ? (lse - l[y]) : 0.f;
  }
}

__global__ void __launch_bounds__(1024) k_mean_reduce(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
  __shared__ double sm[32];
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += (double)v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (threadIdx.x == 0) *out = (float)(a / (double)(n > 0 ? n : 1));
  }
}

__global__ void k_xent_bwd(const float* __restrict__ prob, int64_t B, int ncls, const int64_t* __restrict__ labels,
                           const float* __restrict__ g_loss, float* __restrict__ dl, int64_t ldd) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= B * ncls) return;
  const int64_t r = idx / ncls;
  const int c = (int)(idx % ncls);
  const float scale = (g_loss ? *g_loss : 1.f) / (float)B;
  const int64_t y = labels[r];
  float v = 0.f;
  if (y >= 0 && y < ncls) v = (prob[idx] - (c == y ? 1.f : 0.f)) * scale;
  dl[r * ldd + c] = v;
}

__global__ void k_relu_fwd(float* __restrict__ y, int64_t ldy, int64_t M, int C) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= M * C) return;
  float* p = y + (idx / C) * ldy + (idx % C);
  *p = fmaxf(*p, 0.f);
}
__global__ void k_relu_bwd(const float* __restrict__ gy, int64_t ldgy, const float* __restrict__ z, int64_t ldz, int64_t M,
                           int C, float* __restrict__ dz, int64_t lddz) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= M * C) return;
  const int64_t r = idx / C, c = idx % C;
  dz[r * lddz + c] = z[r * ldz + c] > 0.f ? gy[r * ldgy + c] : 0.f;
}

int relu_fwd(float* y, int64_t ldy, int64_t M, int C, cudaStream_t st) {
  if (M * C <= 0) return YOLAT_OK;
  k_relu_fwd<<<(unsigned)cdiv(M * C, 256), 256, 0, st>>>(y, ldy, M, C);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}
int relu_bwd(const float* gy, int64_t ldgy, const float* z, int64_t ldz, int64_t M, int C, float* dz, int64_t lddz,
             cudaStream_t st) {
  if (M * C <= 0) return YOLAT_OK;
  k_relu_bwd<<<(unsigned)cdiv(M * C, 256), 256, 0, st>>>(gy, ldgy, z, ldz, M, C, dz, lddz);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int segmax_launch(const float* src, int64_t lds, int C, const SegView& sv, int64_t S, const float* stat, float* out,
                  int64_t ldo, int32_t* arg, int64_t lda, cudaStream_t st) {
  if (S <= 0 || C <= 0) return YOLAT_OK;
  dim3 grid((unsigned)S, (unsigned)cdiv(C, SEG_T));
  k_segment_max<<<grid, SEG_T, 0, st>>>(src, lds, C, sv, stat, out, ldo, arg, lda);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}
int segmax_bwd_add_launch(const float* g, int64_t ldg, int C, int64_t S, const int32_t* arg, int64_t lda, float* dsrc,
                          int64_t ldd, cudaStream_t st) {
  if (S * C <= 0) return YOLAT_OK;
  k_segment_max_bwd<<<(unsigned)cdiv(S * C, 256), 256, 0, st>>>(g, ldg, C, S, arg, lda, dsrc, ldd);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}
int fusemax_bwd_nparts(int64_t S) { return (int)cdiv(S > 0 ? S : 1, FM_SEGS_PER_CTA); }
int fusemax_bwd_partial_launch(const float* gp, int64_t ldg, int F, int64_t S, const int32_t* arg, int64_t lda,
                               const float* z, const float* stat, float* part, int* nparts, cudaStream_t st) {
  *nparts = fusemax_bwd_nparts(S);
  dim3 grid((unsigned)cdiv(F, 32), *nparts);
  k_fusemax_bwd_partial<<<grid, 256, 0, st>>>(gp, ldg, F, S, arg, lda, z, stat, part);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}
int fusemax_bwd_apply_launch(float* z, int64_t M, int F, const int32_t* seg_of_row, const float* gp, int64_t ldg,
                             const int32_t* arg, int64_t lda, const float* stat, const float* bstat, cudaStream_t st) {
  if (M * F <= 0) return YOLAT_OK;
  const bool vec = (F % 4 == 0) && (ldg % 4 == 0) && (lda % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(gp) | reinterpret_cast<uintptr_t>(arg) |
                     reinterpret_cast<uintptr_t>(stat) | reinterpret_cast<uintptr_t>(bstat)) & 15u) == 0;
  if (vec) {
    k_fusemax_bwd_apply4<<<dim3((unsigned)cdiv(M, 4 * FMB_ROWS), (unsigned)cdiv(F / 4, 64)), 256, 0, st>>>(z, M, F / 4, seg_of_row, gp,
                                                                                                           ldg, arg, lda, stat, bstat);
    YOLAT_CHECK_LAUNCH();
    return YOLAT_OK;
  }
  k_fusemax_bwd_apply<<<(unsigned)cdiv(M * F, 256), 256, 0, st>>>(z, M, F, seg_of_row, gp, ldg, arg, lda, stat, bstat);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

}  // namespace yolat

using namespace yolat;

static inline bool seg_ok(const int32_t* seg, int64_t M, int64_t S, SegView* v) {
  if (!seg) return false;
  seg_layout(M, S, seg, v);
  return true;
}

extern "C" {

int yolat_segment_mean_fwd(const float* src, int64_t lds, int64_t M, int C, const int32_t* seg, int64_t S, float* out,
                           int64_t ldo, void* stream) {
  SegView sv;
  if (!seg_ok(seg, M, S, &sv) || !out || (M > 0 && !src)) return YOLAT_ERR_INVALID;
  if (S <= 0 || C <= 0) return YOLAT_OK;
  dim3 grid((unsigned)S, (unsigned)cdiv(C, SEG_T));
  k_segment_mean<<<grid, SEG_T, 0, (cudaStream_t)stream>>>(src, lds, C, sv, out, ldo);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int yolat_segment_mean_bwd(const float* g, int64_t ldg, int64_t M, int C, const int32_t* seg, int64_t S, float* dsrc,
                           int64_t ldd, int accumulate, void* stream) {
  SegView sv;
  if (!seg_ok(seg, M, S, &sv) || !dsrc || !g) return YOLAT_ERR_INVALID;
  if (S <= 0 || C <= 0) return YOLAT_OK;
  dim3 grid((unsigned)S, (unsigned)cdiv(C, SEG_T));
  k_segment_mean_bwd<<<grid, SEG_T, 0, (cudaStream_t)stream>>>(g, ldg, C, sv, dsrc, ldd, accumulate);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int yolat_segment_max_fwd(const float* src, int64_t lds, int64_t M, int C, const int32_t* seg, int64_t S, float* out,
                          int64_t ldo, int32_t* arg, void* stream) {
  SegView sv;
  if (!seg_ok(seg, M, S, &sv) || !out || !arg || (M > 0 && !src)) return YOLAT_ERR_INVALID;
  if (S <= 0 || C <= 0) return YOLAT_OK;
  dim3 grid((unsigned)S, (unsigned)cdiv(C, SEG_T));
  k_segment_max<<<grid, SEG_T, 0, (cudaStream_t)stream>>>(src, lds, C, sv, nullptr, out, ldo, arg, C);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int yolat_segment_max_bwd(const float* g, int64_t ldg, int64_t M, int C, int64_t S, const int32_t* arg, float* dsrc,
                          int64_t ldd, int accumulate, void* stream) {
  if (!g || !arg || !dsrc) return YOLAT_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate && M * C > 0) {
    k_zero_strided<<<(unsigned)cdiv(M * C, 256), 256, 0, st>>>(dsrc, ldd, M, C);
    YOLAT_CHECK_LAUNCH();
  }
  if (S * C > 0) {
    k_segment_max_bwd<<<(unsigned)cdiv(S * C, 256), 256, 0, st>>>(g, ldg, C, S, arg, C, dsrc, ldd);
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

int yolat_softmax_xent_fwd(const float* logits, int64_t ldl, int64_t B, int ncls, const int64_t* labels, float* loss,
                           float* prob, float* ws, int64_t ws_floats, void* stream) {
  if (!logits || !labels || !loss || !prob || B <= 0 || ncls <= 0) return YOLAT_ERR_INVALID;
  if (!ws || ws_floats < B) return YOLAT_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  k_xent_rows<<<(unsigned)cdiv(B, 8), 256, 0, st>>>(logits, ldl, B, ncls, labels, prob, ws);
  YOLAT_CHECK_LAUNCH();
  k_mean_reduce<<<1, 1024, 0, st>>>(ws, B, loss);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

int yolat_softmax_xent_bwd(const float* prob, int64_t B, int ncls, const int64_t* labels, const float* g_loss,
                           float* dlogits, int64_t ldd, void* stream) {
  if (!prob || !labels || !dlogits || B <= 0 || ncls <= 0) return YOLAT_ERR_INVALID;
  k_xent_bwd<<<(unsigned)cdiv(B * ncls, 256), 256, 0, (cudaStream_t)stream>>>(prob, B, ncls, labels, g_loss, dlogits, ldd);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

}  // extern "C"
