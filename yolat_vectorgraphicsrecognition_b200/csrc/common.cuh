// common.cuh -- shared helpers for the yolat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/yolat_b200.h"

#ifndef __CUDACC__
#error "CUDA only"
#endif

namespace yolat {

constexpr float kBnEps = 1e-5f;       // nn.BatchNorm1d default (gcn_lib/sparse/torch_nn.py:27)
constexpr float kBnMomentum = 0.1f;
constexpr int kNumSMs = 148;          // B200

void set_last_error(cudaError_t e);
void count_launch();   // process-wide kernel-launch counter behind yolat_launch_count()

// Opt-in device timing of one launch (prof.cu); ids are the YOLAT_PROF_* constants of the public header.
struct ProfScope {
  ProfScope(int id, cudaStream_t st);
  ~ProfScope();
  cudaStream_t st_; bool active_; size_t idx_;
};

#define YOLAT_CHECK_LAUNCH()                                  \
  do {                                                        \
    cudaError_t _e = cudaGetLastError();                      \
    ::yolat::count_launch();                                  \
    if (_e != cudaSuccess) {                                  \
      ::yolat::set_last_error(_e);                            \
      return YOLAT_ERR_LAUNCH;                                \
    }                                                         \
  } while (0)

#define YOLAT_TRY(expr)            \
  do {                             \
    int _s = (expr);               \
    if (_s != YOLAT_OK) return _s; \
  } while (0)

__host__ __device__ inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t align_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// Bump allocator over a caller-provided float workspace.  In "dry" mode (base == nullptr) it only
// measures, so the *_floats query functions run exactly the same planning code as the real call.
struct Arena {
  float* base;
  int64_t cap;
  int64_t off;
  bool overflow;
  Arena(float* b, int64_t c) : base(b), cap(c), off(0), overflow(false) {}
  bool dry() const { return base == nullptr; }
  template <typename T = float>
  T* take(int64_t n_elems) {
    int64_t bytes = align_up(n_elems * (int64_t)sizeof(T), 256);
    int64_t fl = bytes / 4;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += fl;
    if (base && off > cap) overflow = true;
    return p;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- per-BatchNorm statistics block: 4 rows of C floats -------------------------------------
//   [0] sc = gamma * invstd     [1] sh = beta - mean * sc     [2] mean     [3] invstd
// backward statistics block: 2 rows of C floats:   [0] m1 = mean(dy')   [1] m2 = mean(dy' * xhat)
struct BnStat {
  const float* sc;
  const float* sh;
  const float* mean;
  const float* invstd;
  __host__ __device__ BnStat() : sc(nullptr), sh(nullptr), mean(nullptr), invstd(nullptr) {}
  __host__ __device__ BnStat(const float* base, int C) : sc(base), sh(base + C), mean(base + 2 * C), invstd(base + 3 * C) {}
};

// ---- graph buffer layout (int32 units) ----------------------------------------------------------
struct GraphView {
  const int32_t* rowptr_t;  // [N+1]  CSR by target
  const int32_t* rowptr_s;  // [N+1]  CSR by source
  const int32_t* src_t;     // [E]    source node of slot
  const int32_t* dst_t;     // [E]    target node of slot
  const int32_t* eid_t;     // [E]    original edge id of slot
  const int32_t* slot_s;    // [E]    target-CSR slot of the k-th entry of the source-CSR
  const float* deg_inv;     // [N]    1 / max(in-degree, 1)
  const int32_t* err;       // [1]
  // scratch used only while building
  int32_t* cursor_t;        // [N]
  int32_t* cursor_s;        // [N]
  int32_t* slot_of_edge;    // [E]
};

inline int64_t pad4(int64_t n) { return align_up(n, 4); }

inline int64_t graph_layout(int64_t N, int64_t E, const int32_t* base, GraphView* v) {
  int64_t o = 0;
  auto take = [&](int64_t n) {
    const int32_t* p = base ? base + o : nullptr;
    o += pad4(n);
    return p;
  };
  const int32_t* rowptr_t = take(N + 1);
  const int32_t* rowptr_s = take(N + 1);
  const int32_t* src_t = take(E);
  const int32_t* dst_t = take(E);
  const int32_t* eid_t = take(E);
  const int32_t* slot_s = take(E);
  const int32_t* deg_inv = take(N);
  const int32_t* err = take(4);
  const int32_t* cursor_t = take(N);
  const int32_t* cursor_s = take(N);
  const int32_t* slot_of_edge = take(E);
  if (v) {
    v->rowptr_t = rowptr_t; v->rowptr_s = rowptr_s; v->src_t = src_t; v->dst_t = dst_t; v->eid_t = eid_t;
    v->slot_s = slot_s; v->deg_inv = reinterpret_cast<const float*>(deg_inv); v->err = err;
    v->cursor_t = const_cast<int32_t*>(cursor_t); v->cursor_s = const_cast<int32_t*>(cursor_s);
    v->slot_of_edge = const_cast<int32_t*>(slot_of_edge);
  }
  return o;
}

struct SegView {
  const int32_t* segptr;  // [S+1]
  const int32_t* perm;    // [M]
  const int32_t* seg_of_row;  // [M]  segment id of each row (-1: index out of range)
  int32_t* cursor;        // [S] scratch
};
inline int64_t seg_layout(int64_t M, int64_t S, const int32_t* base, SegView* v) {
  int64_t o = 0;
  auto take = [&](int64_t n) {
    const int32_t* p = base ? base + o : nullptr;
    o += pad4(n);
    return p;
  };
  const int32_t* segptr = take(S + 1);
  const int32_t* perm = take(M);
  const int32_t* seg_of_row = take(M);
  const int32_t* cursor = take(S);
  if (v) { v->segptr = segptr; v->perm = perm; v->seg_of_row = seg_of_row; v->cursor = const_cast<int32_t*>(cursor); }
  return o;
}

// ---- internal launchers shared between translation units --------------------------------------
enum GemmMode { GEMM_NT = 0, GEMM_NN = 1, GEMM_TN = 2 };

struct GemmArgs {
  // C[M,N] (row-major, ldc) (+)= op(A)[M,K] * op(B)[K,N] (+ bias[n])
  //  NT: A[m*lda+k], B[n*ldb+k]      (y = x W^T)
  //  NN: A[m*lda+k], B[k*ldb+n]      (dx = dy W)
  //  TN: A[k*lda+m], B[k*ldb+n]      (dW = dy^T x), K is the long (row) dimension
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int M, N; int64_t K;
  const float* bias;       // [N] or null
  const float* a_sc;       // optional prologue on A: relu(a * a_sc[j] + a_sh[j]); j = k (NT/NN) or m (TN)
  const float* a_sh;
  const float* b_sc;       // optional prologue on B (TN only): relu(b * b_sc[n] + b_sh[n])
  const float* b_sh;
  int accumulate;          // C += result
};
// ws: split-K partials; returns needed floats when arena is dry.  Default: tcgen05 3xTF32 kernel (gemm_tc.cu).
int gemm(const GemmArgs& a, GemmMode mode, Arena& ws, cudaStream_t st);
int gemm_simt(const GemmArgs& a, GemmMode mode, Arena& ws, cudaStream_t st);
// Same as gemm(), and additionally leaves per-row-tile column statistics of C (sum, sum of squares) in
// *stat_part [*nparts][2][N] (taken from ws).  *nparts == 0: not produced (split-K or SIMT path) -- the caller
// then runs the separate statistics pass over C.
int gemm_stats(const GemmArgs& a, GemmMode mode, Arena& ws, float** stat_part, int* nparts, cudaStream_t st);

// column statistics of z [M,C] -> BN stat block (+ running-stat update when training)
int bn_forward_stats(const float* z, int64_t ldz, int64_t M, int C, const yolat_bn* bn, int training,
                     float* stat /*[4C]*/, Arena& ws, cudaStream_t st);
// z = x W^T + b (a.C = z) followed by the BN statistic block of z; statistics come from the GEMM epilogue when possible
int linear_bn_stats(const GemmArgs& a, Arena& ws, const yolat_bn* bn, int training, float* stat, cudaStream_t st);
// finalize from externally produced partial sums: part [nparts][2][C] (sum, sumsq)
int bn_finalize_from_partials(const float* part, int nparts, int64_t M, int C, const yolat_bn* bn, int training,
                              float* stat, cudaStream_t st);
// y = act(z*sc+sh)
int bn_apply(const float* z, int64_t ldz, int64_t M, int C, const float* stat, int relu, float* y, int64_t ldy,
             cudaStream_t st);
// BN(+ReLU) backward.  dy source: dense gy [M,C] (ldgy), or gathered rows gy[row_idx[m]] * row_scale[row_idx[m]]
// (* slot_scale[slot_idx[m]] if given).  Writes dz [M,C] (may alias gy when dense), dgamma/dbeta/dbias (nullable).
struct BnBwdArgs {
  const float* gy; int64_t ldgy;
  const int32_t* row_idx;      // nullable
  const float* row_scale;      // nullable (indexed by row_idx[m])
  const int32_t* slot_idx;     // nullable
  const float* slot_scale;     // nullable (indexed by slot_idx[m])
  const float* z; int64_t ldz;
  int64_t M; int C;
  const float* stat;           // forward stat block [4C]
  const float* gamma;
  int relu; int training;
  float* dz; int64_t lddz;
  float* dgamma; float* dbeta; float* dbias;
};
int bn_backward(const BnBwdArgs& a, Arena& ws, cudaStream_t st);
int bn_bwd_finalize(const float* part, int nparts, int64_t M, int C, const float* stat, const float* gamma,
                    int training, float* bstat /*[2C]*/, float* dgamma, float* dbeta, float* dbias, cudaStream_t st);

// segment.cu launchers
int segmax_launch(const float* src, int64_t lds, int C, const SegView& sv, int64_t S, const float* stat, float* out,
                  int64_t ldo, int32_t* arg, int64_t lda, cudaStream_t st);
int segmax_bwd_add_launch(const float* g, int64_t ldg, int C, int64_t S, const int32_t* arg, int64_t lda, float* dsrc,
                          int64_t ldd, cudaStream_t st);
int fusemax_bwd_partial_launch(const float* gp, int64_t ldg, int F, int64_t S, const int32_t* arg, int64_t lda,
                               const float* z, const float* stat, float* part, int* nparts, cudaStream_t st);
int fusemax_bwd_nparts(int64_t S);
int fusemax_bwd_apply_launch(float* z, int64_t M, int F, const int32_t* seg_of_row, const float* gp, int64_t ldg,
                             const int32_t* arg, int64_t lda, const float* stat, const float* bstat, cudaStream_t st);
int relu_fwd(float* y, int64_t ldy, int64_t M, int C, cudaStream_t st);
int relu_bwd(const float* gy, int64_t ldgy, const float* z, int64_t ldz, int64_t M, int C, float* dz, int64_t lddz,
             cudaStream_t st);
// edge.cu launchers
// wr / br / bias3 optional (null): append the lin_r rows so that one GEMM yields P | Q | lin_r(x)
int edge_prep_wpq(const float* w1, int Cin, int C, float* wpq, const float* wr, const float* br, float* bias3,
                  cudaStream_t st);
int edge_assemble_dw1(const float* dwpq, const float* dw1c, int Cin, int C, float* dw1, cudaStream_t st);
int edge_z1_nparts(int64_t N);
int edge_z1(const GraphView& g, int64_t N, int C, const float* pq, const float* attr, const float* w1, int Cin,
            const float* b1, float* z1, float* part, cudaStream_t st);
int edge_agg(const GraphView& g, int64_t N, int C, const float* z2, const float* stat2, const float* ew,
             const float* base, int64_t ldb, float* out, int64_t ldo, cudaStream_t st);
int edge_bwd_scatter(const GraphView& g, int64_t N, int C, const float* dz1, const float* attr, float* dpq,
                     float* part, float* dw1c, cudaStream_t st);

// edge_fused.cu: the fused gather -> edge MLP (tcgen05) -> statistics / segmented mean kernel (C == 64)
enum { EF_TAPE = 1, EF_STATS = 2, EF_AGG = 4, EF_BSTAT = 16 };
bool edge_fused_supported(int C);
bool edge_fused_fits(int64_t N, int64_t ldpq);
int edge_fused_grid(int64_t E);
int edge_stats1_grid(int64_t E);
// slot-ordered records (P / Q byte offsets, eid | attribute row) of one layer call: the TMA-fed input of the kernels below
int64_t edge_records_floats(int64_t E);
int edge_records(const GraphView& g, int64_t E, const float* attr, int64_t ldpq, float* rec, cudaStream_t st);
int edge_stats1(const GraphView& g, int64_t N, int64_t E, const float* pq, int64_t ldpq, const float* rec,
                const float* w1, int Cin, const float* b1, float* part, cudaStream_t st);
// EF_AGG: out = base + mean (base may alias out or be null); every row of out is written exactly once
int edge_fused(const GraphView& g, int64_t N, int64_t E, int flags, const float* pq, int64_t ldpq, const float* rec,
               const float* w1, int Cin, const float* b1, const float* stat1, const float* w2, const float* b2,
               const float* stat2, const float* ew, float* z1, float* z2, float* part, const float* base, int64_t ldb,
               float* out, int64_t ldo, cudaStream_t st);

// edge_bwd.cu: the tape-free backward of the fused edge path (training mode, C == 64): recompute passes D1 / D2T / D2S
struct EdgeBwdWs {
  int grid;
  float *rec_t, *rec_s, *rec5, *Ut, *Xt, *Us, *Xs, *part1, *part_tx, *part2, *part_T, *part_w2, *bstat2, *bstat1;
};
int edge_bwd_grid(int64_t E);
void edge_bwd_layout(Arena& ws, int64_t N, int64_t E, EdgeBwdWs* o);
int edge_bwd_fused(const GraphView& g, int64_t N, int64_t E, const float* pq, int64_t ldpq, const float* attr,
                   const float* w1, int Cin, const float* b1, const float* stat1, const float* gamma1, const float* w2,
                   const float* b2, const float* stat2, const float* gamma2, const float* ew, const float* g_out,
                   int64_t ldgo, const EdgeBwdWs& w, float* dpq, float* dw1c, float* dw2, float* db1, float* dg1,
                   float* dbe1, float* db2, float* dg2, float* dbe2, cudaStream_t st);

// edge_fused2.cu: K-EDGE v6 (channel-major accumulator); same arguments as edge_fused, `rows` = base (EF_AGG, must alias
// out) or g_out (EF_BSTAT: BN2 backward statistics -> part [grid][2][C])
bool edge_fused2_enabled();
int edge_fused2(const GraphView& g, int64_t N, int64_t E, int flags, const float* pq, int64_t ldpq, const float* rec,
                const float* w1, int Cin, const float* b1, const float* stat1, const float* w2, const float* b2,
                const float* stat2, const float* ew, float* z1, float* z2, float* part, const float* rows, int64_t ldr,
                float* out, int64_t ldo, cudaStream_t st);

int colsum(const float* a, int64_t lda, int64_t M, int C, float* out, Arena& ws, cudaStream_t st);
int fill_zero(float* p, int64_t n, cudaStream_t st);

}  // namespace yolat
