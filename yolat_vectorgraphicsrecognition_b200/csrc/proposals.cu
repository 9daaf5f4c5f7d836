// proposals.cu -- proposal enumeration of the reference Dataset on the device.
//
// Reference: Datasets/graph_dict3.py:309-789 (`SESYDFloorPlan._get_proposal`, do_mixup off).  Per image it
//   (1) drops Bezier control points and renumbers nodes / edges / connected components through a python dict (:325-352),
//   (2) per connected component ranks the distinct x / y coordinates, lays a sampling grid of `bbox_sampling_step`
//       cells over the component's box and walks four nested python loops over grid lines (:386-556); every
//       window (x0,y0,x1,y1) of coordinate ranks yields the set of nodes inside it (python sets built from 2-D
//       prefix unions, :436-465, :544-555); the sets are de-duplicated through `list(set(...))` (:557),
//   (3) per distinct set gathers the induced shape / super edges through a dense python adjacency (:560-609), drops the
//       set when it has no shape edge, a degenerate box or no angle (:594-596, :616-617, :674-675), labels it by IoU /
//       IoS against the ground-truth boxes that touch the component (:572-637) and computes 13 statistics (:640-702),
//   (4) appends nodes / edges / attributes renumbered per proposal and picks the largest box per component as the
//       root of its idxTree (:704-768).
// Here:
//   k_prop_o2n        control-point compaction (the o2n dict) by a block scan
//   k_prop_remap      edges / component members renumbered through o2n
//   k_prop_cc_nodes   one CTA per component: members sorted by id, distinct-coordinate ranks, sorted coordinate values
//   k_prop_edge_*     edges bucketed by component (count, scan, fill)
//   k_prop_cc_main    one CTA per component: edges sorted by (lo, hi, id) -- the order the reference's pair loop
//                     (:586-591) visits them --, directed neighbour list, the window walk restated over precomputed
//                     lower / upper bounds, de-duplication by tight rank box, evaluation of every candidate
//   k_prop_cc_scan    exclusive offsets of the components' surviving proposals
//   k_prop_fill       one CTA per component writes the proposals' nodes, edges, attributes, labels, boxes, statistics
// A window's node set equals the set inside its TIGHT rank box, so two windows give the same set iff their tight boxes
// are equal: de-duplication never materialises a set.  Proposals of one component come out in first-occurrence order
// of the window walk; the reference's order is CPython's hash-table iteration order of `set` of int tuples
// (implementation-defined), every other output is identical -- integer / index outputs and all double arithmetic
// that decides something (grid lines, box, IoU / IoS thresholds, dot-product classes) bit for bit (-fmad=false),
// means / standard deviations up to summation order.
//
// Candidates are independent, so k_prop_cc_main / k_prop_fill give each one to a GROUP (a warp: 32 consecutive threads,
// reductions and scans through shared memory ordered by __syncwarp()); CTA barriers only separate the per-component
// phases.  (Profile of the first version, one CTA-wide loop over candidates: 71 % of the warp stalls at CTA barriers
// with one warp's worth of work per candidate -- profiles/r2_prop_cc_main_stalls.txt.)
//
// Every kernel is written for any power-of-two blockDim: the CPU suite compiles this very file with
// -DYOLAT_HOST_EMU against tests/emu/cuda_emu.h (1 / 8 / 64 host threads per CTA, CTAs in sequence) to check the
// logic and the barriers without a GPU; that build is test infrastructure and is never part of libyolat_b200.so.
#ifdef YOLAT_HOST_EMU
#include "cuda_emu.h"
#else
#include "common.cuh"
#define PROP_LAUNCH(kern, grid, block, st, ...) kern<<<(unsigned)(grid), (block), 0, (st)>>>(__VA_ARGS__)
#endif
#include <math.h>

namespace yolat {
namespace prop {

constexpr int kThreads = 256;
constexpr int kMaxLines = 64;               // grid lines per axis: bbox_sampling_step + 2 at most
constexpr int kGroup = 32;                  // threads that share one candidate (a warp)
constexpr int kStats = 13;                  // stat_feats columns (graph_dict3.py:690-691)
constexpr int kLocBits = 20, kIdBits = 24;  // sort keys: (lo:20 | hi:20 | id:24)
constexpr unsigned long long kLocMask = (1ull << kLocBits) - 1, kIdMask = (1ull << kIdBits) - 1;

typedef YolatProposalIn In;
typedef YolatProposalOut Out;

struct Ws {
  int32_t *o2n, *node_cc, *node_loc, *cc_new, *cs, *xi, *yi, *t0, *t1, *nxy, *lrank;
  double *pos, *xv, *yv;
  uint8_t* issup;
  int32_t *edge_n, *sup_n;
  int32_t *ecnt, *eptr, *ecur, *scnt, *sptr, *scur;
  int32_t *elist, *esort, *ea, *eb, *slist, *ssort, *sa, *sb;
  unsigned long long *key_e, *key_s, *key_d;
  int32_t *da, *db, *dcnt;
  int32_t *win, *tight, *wcnt, *cbox, *ccnt, *cm, *cms, *csurv, *clabel, *chas, *cgt, *coff;
  double* cstats;
  int32_t *ncand, *cc_root;
  int64_t *cc_tot, *cc_off;
  int64_t* totals;
};

static int64_t slots_per_cc(int S) {
  const int64_t g = S + 2;
  const int64_t pairs = g * (g - 1) / 2;
  return pairs * pairs;
}

// Carves the byte workspace; with base == nullptr it only measures.
static int64_t carve(const In& in, char* base, Ws* w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) -> char* {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes > 0 ? bytes : 1, 256);
    return p;
  };
  const int64_t n = in.n_all, ct = in.cc_total, ncc = in.ncc, E = in.E, Es = in.Es;
  const int64_t slots = ncc * slots_per_cc(in.sampling_step);
  Ws t;
  t.totals = (int64_t*)take(8 * YOLAT_PROP_TOTALS);
  t.o2n = (int32_t*)take(4 * n);
  t.node_cc = (int32_t*)take(4 * n);
  t.node_loc = (int32_t*)take(4 * n);
  t.pos = (double*)take(16 * n);
  t.issup = (uint8_t*)take(n);
  t.cc_new = (int32_t*)take(4 * ct);
  t.cs = (int32_t*)take(4 * ct);
  t.xi = (int32_t*)take(4 * ct);
  t.yi = (int32_t*)take(4 * ct);
  t.t0 = (int32_t*)take(4 * ct);
  t.t1 = (int32_t*)take(4 * ct);
  t.lrank = (int32_t*)take(4 * ct * (kThreads / kGroup));     // one copy per group of k_prop_fill
  t.xv = (double*)take(8 * ct);
  t.yv = (double*)take(8 * ct);
  t.nxy = (int32_t*)take(8 * ncc);
  t.edge_n = (int32_t*)take(8 * E);
  t.sup_n = (int32_t*)take(8 * Es);
  t.ecnt = (int32_t*)take(4 * (ncc + 1));
  t.eptr = (int32_t*)take(4 * (ncc + 1));
  t.ecur = (int32_t*)take(4 * (ncc + 1));
  t.scnt = (int32_t*)take(4 * (ncc + 1));
  t.sptr = (int32_t*)take(4 * (ncc + 1));
  t.scur = (int32_t*)take(4 * (ncc + 1));
  t.elist = (int32_t*)take(4 * E);
  t.esort = (int32_t*)take(4 * E);
  t.ea = (int32_t*)take(4 * E);
  t.eb = (int32_t*)take(4 * E);
  t.slist = (int32_t*)take(4 * Es);
  t.ssort = (int32_t*)take(4 * Es);
  t.sa = (int32_t*)take(4 * Es);
  t.sb = (int32_t*)take(4 * Es);
  t.key_e = (unsigned long long*)take(8 * E);
  t.key_s = (unsigned long long*)take(8 * Es);
  t.key_d = (unsigned long long*)take(16 * E);
  t.da = (int32_t*)take(8 * E);
  t.db = (int32_t*)take(8 * E);
  t.dcnt = (int32_t*)take(4 * ncc);
  t.win = (int32_t*)take(16 * slots);
  t.tight = (int32_t*)take(16 * slots);
  t.wcnt = (int32_t*)take(4 * slots);
  t.cbox = (int32_t*)take(16 * slots);
  t.ccnt = (int32_t*)take(4 * slots);
  t.cm = (int32_t*)take(4 * slots);
  t.cms = (int32_t*)take(4 * slots);
  t.csurv = (int32_t*)take(4 * slots);
  t.clabel = (int32_t*)take(4 * slots);
  t.chas = (int32_t*)take(4 * slots);
  t.cgt = (int32_t*)take(4 * slots);
  t.coff = (int32_t*)take(16 * slots);
  t.cstats = (double*)take(8 * kStats * slots);
  t.ncand = (int32_t*)take(4 * ncc);
  t.cc_root = (int32_t*)take(4 * ncc);
  t.cc_tot = (int64_t*)take(32 * ncc);
  t.cc_off = (int64_t*)take(32 * ncc);
  if (w) *w = t;
  return off;
}

// ---- block-level helpers (any power-of-two blockDim <= kThreads; scratch = kThreads shared entries) ----------------
template <typename T>
__device__ T block_sum(T v, T* sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (tid < s) sh[tid] += sh[tid + s];
    __syncthreads();
  }
  const T r = sh[0];
  __syncthreads();
  return r;
}
// exclusive scan of one value per thread; *total = sum over the block
template <typename T>
__device__ T block_excl_scan(T v, T* sh, T* total) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int off = 1; off < (int)blockDim.x; off <<= 1) {
    const T t = tid >= off ? sh[tid - off] : 0;
    __syncthreads();
    sh[tid] += t;
    __syncthreads();
  }
  const T incl = sh[tid];
  *total = sh[blockDim.x - 1];
  __syncthreads();
  return incl - v;
}

// One reduction tree for everything a candidate needs at once: four integer sums, two double sums, a maximum, a
// minimum and an arg-max (largest key; the smallest arg among equal keys = numpy's first maximum).
struct Acc {
  long long l[4];
  double s[2];
  double mx, mn, key;
  long long arg;
};
__device__ inline void acc_init(Acc& a) {
  for (int i = 0; i < 4; ++i) a.l[i] = 0;
  a.s[0] = a.s[1] = 0.0;
  a.mx = -INFINITY; a.mn = INFINITY; a.key = -INFINITY;
  a.arg = 0x7fffffffffffffffll;
}
__device__ inline void acc_merge(Acc& a, const Acc& b) {
  for (int i = 0; i < 4; ++i) a.l[i] += b.l[i];
  a.s[0] += b.s[0]; a.s[1] += b.s[1];
  if (b.mx > a.mx) a.mx = b.mx;
  if (b.mn < a.mn) a.mn = b.mn;
  if (b.key > a.key || (b.key == a.key && b.arg < a.arg)) { a.key = b.key; a.arg = b.arg; }
}
__device__ Acc block_reduce(const Acc& v, Acc* sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (tid < s) acc_merge(sh[tid], sh[tid + s]);
    __syncthreads();
  }
  const Acc r = sh[0];
  __syncthreads();
  return r;
}

// ---- group-level helpers: a group = min(kGroup, blockDim) consecutive threads (one warp on the device); `sh` points at
// the group's own segment of a shared array, __syncwarp() orders the group's shared-memory accesses
struct Group { int size, lane, id, count; };
__device__ inline Group this_group() {
  Group g;
  g.size = (int)blockDim.x < kGroup ? (int)blockDim.x : kGroup;
  g.lane = (int)threadIdx.x % g.size;
  g.id = (int)threadIdx.x / g.size;
  g.count = (int)blockDim.x / g.size;
  return g;
}
__device__ Acc group_reduce(const Acc& v, Acc* sh, const Group& g) {
  sh[g.lane] = v;
  __syncwarp();
  for (int s = g.size >> 1; s > 0; s >>= 1) {
    if (g.lane < s) acc_merge(sh[g.lane], sh[g.lane + s]);
    __syncwarp();
  }
  const Acc r = sh[0];
  __syncwarp();
  return r;
}
__device__ int group_excl_scan(int v, int* sh, const Group& g, int* total) {
  sh[g.lane] = v;
  __syncwarp();
  for (int off = 1; off < g.size; off <<= 1) {
    const int t = g.lane >= off ? sh[g.lane - off] : 0;
    __syncwarp();
    sh[g.lane] += t;
    __syncwarp();
  }
  const int incl = sh[g.lane];
  *total = sh[g.size - 1];
  __syncwarp();
  return incl - v;
}
// four running sums scanned together (nodes, shape edges, super edges, proposals); plain aggregate (shared arrays of it)
struct I4 { int v[4]; };
__device__ inline I4 i4_zero() { I4 r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0; return r; }
__device__ I4 block_excl_scan_i4(const I4& v, I4* sh, I4* total) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int off = 1; off < (int)blockDim.x; off <<= 1) {
    I4 t = i4_zero();
    if (tid >= off) t = sh[tid - off];
    __syncthreads();
    for (int i = 0; i < 4; ++i) sh[tid].v[i] += t.v[i];
    __syncthreads();
  }
  I4 ex = sh[tid];
  for (int i = 0; i < 4; ++i) ex.v[i] -= v.v[i];
  *total = sh[blockDim.x - 1];
  __syncthreads();
  return ex;
}

__device__ inline void raise_err(int64_t* totals, unsigned long long bit, int cc) {
  atomicOr((unsigned long long*)(totals + YOLAT_PROP_T_ERR), bit);
  atomicMin((unsigned long long*)(totals + YOLAT_PROP_T_ERR_CC), (unsigned long long)cc);
}

// ---- (1) control-point compaction: o2n[i] = number of non-control nodes before i (graph_dict3.py:325-330) ---------
__global__ void __launch_bounds__(kThreads) k_prop_o2n(In in, Ws w) {
  __shared__ int sh[kThreads];
  constexpr int kPer = 8;                                  // consecutive nodes per thread and chunk
  const int tid = threadIdx.x, bd = blockDim.x;
  int carry = 0;
  for (int64_t base = 0; base < in.n_all; base += (int64_t)bd * kPer) {
    const int64_t i0 = base + (int64_t)tid * kPer;
    int mine = 0;
    for (int t = 0; t < kPer; ++t) mine += (i0 + t < in.n_all && in.is_control[i0 + t] == 0) ? 1 : 0;
    int tot;
    int q = carry + block_excl_scan<int>(mine, sh, &tot);
    for (int t = 0; t < kPer; ++t) {
      const int64_t i = i0 + t;
      if (i >= in.n_all) break;
      if (in.is_control[i] == 0) {
        w.o2n[i] = q;
        w.pos[2 * q] = in.pos[2 * i];
        w.pos[2 * q + 1] = in.pos[2 * i + 1];
        w.issup[q] = in.is_super[i];
        w.node_cc[q] = -1;
        w.node_loc[q] = 0;
        ++q;
      } else {
        w.o2n[i] = -1;
      }
    }
    carry += tot;
  }
  if (tid == 0) w.totals[YOLAT_PROP_T_NODES_IN] = carry;
}

// ---- (1b) renumber edges and component members (graph_dict3.py:332-348) -------------------------------------------
__device__ inline int remap_one(const In& in, const Ws& w, int64_t v) {
  if (v < 0 || v >= in.n_all) {
    raise_err(w.totals, YOLAT_PROP_ERR_INDEX, 0);
    return -1;
  }
  const int q = w.o2n[v];
  if (q < 0) raise_err(w.totals, YOLAT_PROP_ERR_CONTROL_REF, 0);   // the reference's dict lookup raises KeyError
  return q;
}
__global__ void __launch_bounds__(kThreads) k_prop_remap(In in, Ws w) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < 2 * in.E; i += stride) w.edge_n[i] = remap_one(in, w, in.edge[i]);
  for (int64_t i = t0; i < 2 * in.Es; i += stride) w.sup_n[i] = remap_one(in, w, in.edge_super[i]);
  for (int64_t i = t0; i < in.cc_total; i += stride) w.cc_new[i] = remap_one(in, w, in.cc_idx[i]);
  for (int64_t i = t0; i <= in.ncc; i += stride) {
    w.ecnt[i] = 0; w.scnt[i] = 0;
  }
}

// ---- (2a) per component: members sorted by id, coordinate ranks (graph_dict3.py:386-413) --------------------------
// ranks of the distinct values of coordinate `axis` among the component's members (sorted-member order r):
//   rank_out[r] = number of distinct values < v_r;  vals[rank] = value;  returns the number of distinct values
__device__ int cc_axis_ranks(const Ws& w, int64_t nb, int nc, int axis, int32_t* rank_out, double* vals, int* sh) {
  const int tid = threadIdx.x, bd = blockDim.x;
  for (int r = tid; r < nc; r += bd) {                 // first occurrence of its value?
    const double v = w.pos[2 * w.cs[nb + r] + axis];
    int first = 1;
    for (int k = 0; k < r; ++k)
      if (w.pos[2 * w.cs[nb + k] + axis] == v) { first = 0; break; }
    w.t0[nb + r] = first;
  }
  __syncthreads();
  int mine = 0;
  for (int r = tid; r < nc; r += bd) {
    const double v = w.pos[2 * w.cs[nb + r] + axis];
    int rk = 0;
    for (int k = 0; k < nc; ++k)
      if (w.t0[nb + k] && w.pos[2 * w.cs[nb + k] + axis] < v) ++rk;
    rank_out[nb + r] = rk;
    if (w.t0[nb + r]) {
      vals[nb + rk] = v;
      ++mine;
    }
  }
  const int n = block_sum<int>(mine, sh);
  return n;
}

__global__ void __launch_bounds__(kThreads) k_prop_cc_nodes(In in, Ws w) {
  __shared__ int sh[kThreads];
  const int c = blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const int64_t nb = in.cc_ptr[c];
  const int nc = (int)(in.cc_ptr[c + 1] - nb);
  if (nc > (int)kLocMask - 1) {
    if (tid == 0) { raise_err(w.totals, YOLAT_PROP_ERR_LIMIT, c); w.nxy[2 * c] = 0; w.nxy[2 * c + 1] = 0; }
    return;
  }
  bool bad = false;
  for (int j = tid; j < nc; j += bd) {
    const int v = w.cc_new[nb + j];
    if (v < 0) { bad = true; continue; }
    int r = 0;
    for (int k = 0; k < nc; ++k) {
      const int u = w.cc_new[nb + k];
      if (u < v || (u == v && k < j)) ++r;
    }
    w.cs[nb + r] = v;
  }
  const int nbad = block_sum<int>(bad ? 1 : 0, sh);   // a control point inside a component: flagged by k_prop_remap
  if (nbad) {
    if (tid == 0) { w.nxy[2 * c] = 0; w.nxy[2 * c + 1] = 0; }
    return;
  }
  for (int r = tid; r < nc; r += bd) {
    const int v = w.cs[nb + r];
    const int prev = atomicExch(&w.node_cc[v], c);
    if (prev != -1) raise_err(w.totals, YOLAT_PROP_ERR_CC_OVERLAP, c);
    w.node_loc[v] = r;
  }
  __syncthreads();
  const int nx = cc_axis_ranks(w, nb, nc, 0, w.xi, w.xv, sh);
  __syncthreads();
  const int ny = cc_axis_ranks(w, nb, nc, 1, w.yi, w.yv, sh);
  if (tid == 0) { w.nxy[2 * c] = nx; w.nxy[2 * c + 1] = ny; }
}

// ---- (2b) edges bucketed by component -----------------------------------------------------------------------------
__device__ inline int edge_cc(const Ws& w, const int32_t* en, int64_t e) {
  const int a = en[2 * e], b = en[2 * e + 1];
  if (a < 0 || b < 0) return -1;
  const int ca = w.node_cc[a];
  return (ca >= 0 && ca == w.node_cc[b]) ? ca : -1;   // an edge across components is in no proposal
}
__global__ void __launch_bounds__(kThreads) k_prop_edge_count(In in, Ws w) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t e = t0; e < in.E; e += stride) {
    const int c = edge_cc(w, w.edge_n, e);
    if (c >= 0) atomicAdd(&w.ecnt[c], 1);
  }
  for (int64_t e = t0; e < in.Es; e += stride) {
    const int c = edge_cc(w, w.sup_n, e);
    if (c >= 0) atomicAdd(&w.scnt[c], 1);
  }
}
__global__ void __launch_bounds__(kThreads) k_prop_edge_scan(In in, Ws w) {
  __shared__ int sh[kThreads];
  const int tid = threadIdx.x, bd = blockDim.x;
  for (int which = 0; which < 2; ++which) {
    const int32_t* cnt = which ? w.scnt : w.ecnt;
    int32_t* ptr = which ? w.sptr : w.eptr;
    int32_t* cur = which ? w.scur : w.ecur;
    int carry = 0;
    for (int64_t base = 0; base < in.ncc; base += bd) {
      const int64_t i = base + tid;
      const int v = i < in.ncc ? cnt[i] : 0;
      int tot;
      const int ex = block_excl_scan<int>(v, sh, &tot);
      if (i < in.ncc) { ptr[i] = carry + ex; cur[i] = 0; }
      carry += tot;
    }
    if (tid == 0) ptr[in.ncc] = carry;
  }
}
__global__ void __launch_bounds__(kThreads) k_prop_edge_fill(In in, Ws w) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t e = t0; e < in.E; e += stride) {
    const int c = edge_cc(w, w.edge_n, e);
    if (c >= 0) w.elist[w.eptr[c] + atomicAdd(&w.ecur[c], 1)] = (int)e;
  }
  for (int64_t e = t0; e < in.Es; e += stride) {
    const int c = edge_cc(w, w.sup_n, e);
    if (c >= 0) w.slist[w.sptr[c] + atomicAdd(&w.scur[c], 1)] = (int)e;
  }
}

// ---- (2c) sort a component's edges by (lo, hi, id): the order of the reference's `for i: for j > i: A[i][j]` walk ---
// list/sorted/la/lb are offset to the component's segment already; la/lb = member ranks of the edge's two ends as given
__device__ void cc_sort_edges(const Ws& w, const int32_t* en, const int32_t* list, int m, unsigned long long* key,
                              int32_t* sorted, int32_t* la, int32_t* lb) {
  const int tid = threadIdx.x, bd = blockDim.x;
  for (int j = tid; j < m; j += bd) {
    const int e = list[j];
    const unsigned long long a = (unsigned)w.node_loc[en[2 * e]], b = (unsigned)w.node_loc[en[2 * e + 1]];
    const unsigned long long lo = a < b ? a : b, hi = a < b ? b : a;
    key[j] = (lo << (kLocBits + kIdBits)) | (hi << kIdBits) | (unsigned long long)e;
  }
  __syncthreads();
  for (int j = tid; j < m; j += bd) {
    const unsigned long long kj = key[j];
    int r = 0;
    for (int k = 0; k < m; ++k) r += key[k] < kj ? 1 : 0;
    const int e = (int)(kj & kIdMask);
    sorted[r] = e;
    la[r] = w.node_loc[en[2 * e]];
    lb[r] = w.node_loc[en[2 * e + 1]];
  }
  __syncthreads();
}

struct Box { int x0, y0, x1, y1; };
__device__ inline bool inside(const Ws& w, int64_t nb, int r, const Box& b) {
  const int x = w.xi[nb + r], y = w.yi[nb + r];
  return x >= b.x0 && x <= b.x1 && y >= b.y0 && y <= b.y1;
}
// number of sorted distinct values v[0..n) that are < bound (strict) or <= bound
__device__ inline int count_below(const double* v, int n, double bound, bool inclusive) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const bool below = inclusive ? (v[mid] <= bound) : (v[mid] < bound);
    if (below) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// np.arange(lo, hi, (hi - lo) / S) followed by np.append(., hi)  (graph_dict3.py:470-479), value for value
__device__ int grid_lines(double lo, double hi, int S, double* out, bool* ok) {
  const double step = (hi - lo) / (double)S;
  *ok = true;
  if (!(step != 0.0)) { *ok = false; return 0; }          // numpy: 0/0 -> 'arange: cannot compute length'
  const double flen = ceil((hi - lo) / step);
  if (!(flen >= 0.0) || flen > (double)(kMaxLines - 1)) { *ok = false; return -1; }
  const int L = (int)flen;
  const double second = lo + step;
  const double delta = second - lo;                          // numpy fills element i >= 2 as start + i * (a[1] - a[0])
  for (int i = 0; i < L; ++i) out[i] = i == 0 ? lo : (i == 1 ? second : lo + (double)i * delta);
  out[L] = hi;
  return L + 1;
}
// move_endpoint / move_endpoint_close (graph_dict3.py:482-500) over sorted distinct values: `below` = number of values
// <= bound (resp. < bound), n = len(values)
__device__ inline int move_end(int x, int n, int below) {
  if (x >= n) return x - 1;
  return (x > below ? x : below) - 1;
}

struct Iou { double iou, ios; };
// utils/det_util.py:326-340 with box1 = the proposal, box2 = one ground-truth box, operation for operation
__device__ inline Iou iou_ios(const double* p, const double* g) {
  const double ix1 = p[0] > g[0] ? p[0] : g[0], iy1 = p[1] > g[1] ? p[1] : g[1];
  const double ix2 = p[2] < g[2] ? p[2] : g[2], iy2 = p[3] < g[3] ? p[3] : g[3];
  const double dx = ix2 - ix1, dy = iy2 - iy1;
  const double inter = (dx > 0.0 ? dx : 0.0) * (dy > 0.0 ? dy : 0.0);
  const double a1 = (p[2] - p[0]) * (p[3] - p[1]);
  const double a2 = (g[2] - g[0]) * (g[3] - g[1]);
  Iou r;
  r.iou = inter / (a1 + a2 - inter + 1e-16);
  r.ios = inter / a2;
  return r;
}
// utils/det_util.py:355-362: strict overlap of the component's box with a ground-truth box
__device__ inline bool gt_touches(const double* ccb, const double* g) {
  const double ix1 = ccb[0] > g[0] ? ccb[0] : g[0], iy1 = ccb[1] > g[1] ? ccb[1] : g[1];
  const double ix2 = ccb[2] < g[2] ? ccb[2] : g[2], iy2 = ccb[3] < g[3] ? ccb[3] : g[3];
  return ix2 > ix1 && iy2 > iy1;
}

// dot products of every pair of distinct neighbours of every anchor inside the box (graph_dict3.py:646-669).
// pass 0: l[0..3] = number of angles / obtuse / acute / right, s[0] = sum, mx / mn; pass 1: s[0] += (dot - mean)^2
__device__ void cc_angles(const Ws& w, int64_t nb, int nc, int d0, int nd, const Box& box, int pass, double mean, Acc& a,
                          const Group& g) {
  const int32_t* da = w.da + d0;
  const int32_t* db = w.db + d0;
  for (int r = g.lane; r < nc; r += g.size) {
    if (!inside(w, nb, r, box)) continue;
    int lo = 0, hi = nd;                                   // first directed entry of anchor r
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (da[mid] < r) lo = mid + 1; else hi = mid;
    }
    const int s0 = lo;
    const double ax = w.pos[2 * w.cs[nb + r]], ay = w.pos[2 * w.cs[nb + r] + 1];
    for (int i = s0; i < nd && da[i] == r; ++i) {
      if (i > s0 && db[i] == db[i - 1]) continue;          // parallel edges: the reference's adjacency is a set
      if (!inside(w, nb, db[i], box)) continue;
      const double v0x = w.pos[2 * w.cs[nb + db[i]]] - ax, v0y = w.pos[2 * w.cs[nb + db[i]] + 1] - ay;
      for (int j = i + 1; j < nd && da[j] == r; ++j) {
        if (db[j] == db[j - 1]) continue;
        if (!inside(w, nb, db[j], box)) continue;
        const double v1x = w.pos[2 * w.cs[nb + db[j]]] - ax, v1y = w.pos[2 * w.cs[nb + db[j]] + 1] - ay;
        const double dot = v0x * v1x + v0y * v1y;
        if (pass == 0) {
          if (dot <= -1e-2) ++a.l[1];
          else if (dot >= 1e-2) ++a.l[2];
          else if (fabs(dot) < 1e-2) ++a.l[3];
          ++a.l[0];
          a.s[0] += dot;
          if (dot > a.mx) a.mx = dot;
          if (dot < a.mn) a.mn = dot;
        } else {
          const double d = dot - mean;
          a.s[0] += d * d;
        }
      }
    }
  }
}

// ---- (2d, 3) one CTA per component: sort, window walk, de-duplication, evaluation ---------------------------------
__global__ void __launch_bounds__(kThreads) k_prop_cc_main(In in, Ws w, int wmax) {
  __shared__ Acc sh_acc[kThreads];
  __shared__ I4 sh_i4[kThreads];
  __shared__ int sh_i[kThreads];
  __shared__ double s_xg[kMaxLines], s_yg[kMaxLines];
  __shared__ int s_lbx[kMaxLines], s_ubx[kMaxLines], s_lby[kMaxLines], s_uby[kMaxLines];
  __shared__ int s_gx, s_gy, s_nw, s_flag;

  const int c = blockIdx.x, tid = threadIdx.x, bd = blockDim.x;
  const int64_t nb = in.cc_ptr[c];
  const int nc = (int)(in.cc_ptr[c + 1] - nb);
  const int nx = w.nxy[2 * c], ny = w.nxy[2 * c + 1];
  const int e0 = w.eptr[c], m = w.eptr[c + 1] - e0;
  const int s0 = w.sptr[c], ms = w.sptr[c + 1] - s0;
  const int64_t slot0 = (int64_t)c * wmax;
  const int S = in.sampling_step;

  if (tid == 0) {
    w.ncand[c] = 0; w.cc_root[c] = -1; w.dcnt[c] = 0;
    for (int q = 0; q < 4; ++q) w.cc_tot[4 * c + q] = 0;
  }
  if (nc == 0 || nx == 0 || ny == 0) {                     // an empty component: the reference's .max() of nothing raises
    if (tid == 0) raise_err(w.totals, YOLAT_PROP_ERR_NO_PROPOSAL, c);
    return;
  }
  if (m >= (1 << (kIdBits - 1)) || in.E > (int64_t)kIdMask || in.Es > (int64_t)kIdMask) {
    if (tid == 0) raise_err(w.totals, YOLAT_PROP_ERR_LIMIT, c);
    return;
  }

  // -- edges in the reference's visiting order; directed neighbour entries sorted by (anchor, neighbour)
  cc_sort_edges(w, w.edge_n, w.elist + e0, m, w.key_e + e0, w.esort + e0, w.ea + e0, w.eb + e0);
  cc_sort_edges(w, w.sup_n, w.slist + s0, ms, w.key_s + s0, w.ssort + s0, w.sa + s0, w.sb + s0);
  const int d0 = 2 * e0;
  int valid = 0;
  for (int j = tid; j < m; j += bd) {
    const unsigned long long a = (unsigned)w.ea[e0 + j], b = (unsigned)w.eb[e0 + j];
    if (a != b) {
      w.key_d[d0 + 2 * j] = (a << (kLocBits + kIdBits)) | (b << kIdBits) | (unsigned long long)(2 * j);
      w.key_d[d0 + 2 * j + 1] = (b << (kLocBits + kIdBits)) | (a << kIdBits) | (unsigned long long)(2 * j + 1);
      valid += 2;
    } else {                                               // self loop: never in a pair i < j of the reference's walk
      w.key_d[d0 + 2 * j] = (kLocMask << (kLocBits + kIdBits)) | (unsigned long long)(2 * j);
      w.key_d[d0 + 2 * j + 1] = (kLocMask << (kLocBits + kIdBits)) | (unsigned long long)(2 * j + 1);
    }
  }
  const int nd = block_sum<int>(valid, sh_i);              // (the sync inside also publishes key_d)
  for (int t = tid; t < 2 * m; t += bd) {
    const unsigned long long kt = w.key_d[d0 + t];
    int r = 0;
    for (int k = 0; k < 2 * m; ++k) r += w.key_d[d0 + k] < kt ? 1 : 0;
    w.da[d0 + r] = (int)(kt >> (kLocBits + kIdBits));
    w.db[d0 + r] = (int)((kt >> kIdBits) & kLocMask);
  }
  if (tid == 0) w.dcnt[c] = nd;
  __syncthreads();

  // -- sampling grid and the bounds of every grid line among the distinct coordinate values
  const double* xv = w.xv + nb;
  const double* yv = w.yv + nb;
  const double ccb[4] = {xv[0], yv[0], xv[nx - 1], yv[ny - 1]};
  if (tid == 0) {
    bool okx, oky;
    s_gx = grid_lines(ccb[0], ccb[2], S, s_xg, &okx);
    s_gy = grid_lines(ccb[1], ccb[3], S, s_yg, &oky);
    s_flag = 0;
    if (!okx || !oky) {
      s_flag = 1;
      raise_err(w.totals, (s_gx < 0 || s_gy < 0) ? YOLAT_PROP_ERR_LIMIT : YOLAT_PROP_ERR_ZERO_STEP, c);
    }
  }
  __syncthreads();
  if (s_flag) return;
  const int gx = s_gx, gy = s_gy;
  for (int g = tid; g < gx; g += bd) {
    s_lbx[g] = count_below(xv, nx, s_xg[g], false);
    s_ubx[g] = count_below(xv, nx, s_xg[g], true);
  }
  for (int g = tid; g < gy; g += bd) {
    s_lby[g] = count_below(yv, ny, s_yg[g], false);
    s_uby[g] = count_below(yv, ny, s_yg[g], true);
  }
  __syncthreads();

  // -- the four nested loops of graph_dict3.py:502-523 (stateful `prev` skipping), one thread
  if (tid == 0) {
    int nw = 0;
    bool over = false;
    int prev_y0 = -1;
    for (int iy0 = 0; iy0 < gy; ++iy0) {
      int y0 = move_end(prev_y0 + 1, ny, s_lby[iy0]);
      if (y0 != ny) ++y0;
      if (y0 == prev_y0) continue;
      prev_y0 = y0;
      int prev_x0 = -1;
      for (int ix0 = 0; ix0 < gx; ++ix0) {
        int x0 = move_end(prev_x0 + 1, nx, s_lbx[ix0]);
        if (x0 != ny) ++x0;                               // sic: the reference compares with len(y_values) (:516)
        if (x0 == prev_x0) continue;
        prev_x0 = x0;
        int prev_y1 = y0;
        for (int iy1 = iy0 + 1; iy1 < gy; ++iy1) {
          const int y1 = move_end(prev_y1 + 1, ny, s_uby[iy1]);
          if (y1 == prev_y1) continue;
          prev_y1 = y1;
          int prev_x1 = x0;
          for (int ix1 = ix0 + 1; ix1 < gx; ++ix1) {
            const int x1 = move_end(prev_x1 + 1, nx, s_ubx[ix1]);
            if (x1 == prev_x1) continue;
            prev_x1 = x1;
            if (nw >= wmax) { over = true; continue; }
            int32_t* q = w.win + 4 * (slot0 + nw);
            q[0] = x0; q[1] = y0; q[2] = x1; q[3] = y1;
            ++nw;
          }
        }
      }
    }
    s_nw = nw;
    if (over) { s_flag = 1; raise_err(w.totals, YOLAT_PROP_ERR_LIMIT, c); }
  }
  __syncthreads();
  if (s_flag) return;
  const int nw = s_nw;

  // -- node count and tight rank box of every window
  for (int wi = tid; wi < nw; wi += bd) {
    const int32_t* q = w.win + 4 * (slot0 + wi);
    const Box b = {q[0], q[1], q[2], q[3]};
    int cnt = 0, tx0 = 1 << 30, ty0 = 1 << 30, tx1 = -1, ty1 = -1;
    for (int r = 0; r < nc; ++r) {
      const int x = w.xi[nb + r], y = w.yi[nb + r];
      if (x >= b.x0 && x <= b.x1 && y >= b.y0 && y <= b.y1) {
        ++cnt;
        tx0 = x < tx0 ? x : tx0; tx1 = x > tx1 ? x : tx1;
        ty0 = y < ty0 ? y : ty0; ty1 = y > ty1 ? y : ty1;
      }
    }
    int32_t* t = w.tight + 4 * (slot0 + wi);
    t[0] = tx0; t[1] = ty0; t[2] = tx1; t[3] = ty1;
    w.wcnt[slot0 + wi] = cnt;
  }
  __syncthreads();

  // -- distinct non-empty sets, first occurrence kept (list(set(sub_clusters)), :557; the empty tuple has no edges)
  int ncand = 0;
  for (int base = 0; base < nw; base += bd) {
    const int wi = base + tid;
    int keep = 0;
    if (wi < nw && w.wcnt[slot0 + wi] > 0) {
      keep = 1;
      const int32_t* t = w.tight + 4 * (slot0 + wi);
      for (int k = 0; k < wi; ++k) {
        const int32_t* u = w.tight + 4 * (slot0 + k);
        if (w.wcnt[slot0 + k] > 0 && u[0] == t[0] && u[1] == t[1] && u[2] == t[2] && u[3] == t[3]) { keep = 0; break; }
      }
    }
    int tot;
    const int ex = block_excl_scan<int>(keep, sh_i, &tot);
    if (keep) {
      const int k = ncand + ex;
      const int32_t* t = w.tight + 4 * (slot0 + wi);
      int32_t* q = w.cbox + 4 * (slot0 + k);
      q[0] = t[0]; q[1] = t[1]; q[2] = t[2]; q[3] = t[3];
      w.ccnt[slot0 + k] = w.wcnt[slot0 + wi];
    }
    ncand += tot;
  }
  __syncthreads();
  if (tid == 0) w.ncand[c] = ncand;

  // -- ground-truth boxes that touch the component (:572-576)
  int touching = 0;
  for (int64_t g = tid; g < in.G; g += bd) touching += gt_touches(ccb, in.gt_bbox + 4 * g) ? 1 : 0;
  if (block_sum<int>(touching, sh_i) == 0) {
    if (tid == 0) raise_err(w.totals, YOLAT_PROP_ERR_NO_GT, c);
    return;
  }

  // -- every candidate: one group each, three group reductions per candidate, no CTA barrier inside the loop
  const Group g = this_group();
  Acc* gsh = sh_acc + g.id * g.size;
  const int acol = in.A - 1;
  for (int k = g.id; k < ncand; k += g.count) {
    const int64_t slot = slot0 + k;
    const int32_t* q = w.cbox + 4 * slot;
    const Box box = {q[0], q[1], q[2], q[3]};
    const double pb[4] = {xv[box.x0], yv[box.y0], xv[box.x1], yv[box.y1]};
    const double width = pb[2] - pb[0], height = pb[3] - pb[1];
    if (g.lane == 0) w.csurv[slot] = 0;
    // (i) induced shape / super edges (both ends inside, no self loop), the sum of the shape edges' distance attribute,
    //     and the best-overlapping touching ground-truth box (:619-637: first maximum of IoU)
    Acc a;
    acc_init(a);
    for (int j = g.lane; j < m; j += g.size) {
      const int ja = w.ea[e0 + j], jb = w.eb[e0 + j];
      if (ja != jb && inside(w, nb, ja, box) && inside(w, nb, jb, box)) {
        ++a.l[0];
        a.s[0] += in.e_attr[(int64_t)w.esort[e0 + j] * in.A + acol];
      }
    }
    for (int j = g.lane; j < ms; j += g.size) {
      const int ja = w.sa[s0 + j], jb = w.sb[s0 + j];
      if (ja != jb && inside(w, nb, ja, box) && inside(w, nb, jb, box)) ++a.l[1];
    }
    for (int64_t t = g.lane; t < in.G; t += g.size) {
      const double* gb = in.gt_bbox + 4 * t;
      if (!gt_touches(ccb, gb)) continue;
      const double v = iou_ios(pb, gb).iou;
      if (v > a.key) { a.key = v; a.arg = t; }
    }
    a = group_reduce(a, gsh, g);
    const long long mk = a.l[0], msk = a.l[1];
    if (mk == 0) continue;                                 // :594-596
    if (width < 1e-4 || height < 1e-4) continue;           // :616-617
    const double dmean = a.s[0] / (double)mk;
    long long gsel = a.arg;
    if (gsel >= in.G) gsel = 0;                            // only with NaN boxes (np.argmax would pick the first NaN)
    const Iou sel = iou_ios(pb, in.gt_bbox + 4 * gsel);

    // (ii) angles (:640-675)
    acc_init(a);
    cc_angles(w, nb, nc, d0, nd, box, 0, 0.0, a, g);
    a = group_reduce(a, gsh, g);
    const long long acnt = a.l[0], n_more = a.l[1], n_less = a.l[2], n_eq = a.l[3];
    if (acnt == 0) continue;                               // :674-675
    const double amean = a.s[0] / (double)acnt, amax = a.mx, amin = a.mn;

    // (iii) squared deviations of the angles and of the distance attribute (np.std: two-pass, population)
    acc_init(a);
    cc_angles(w, nb, nc, d0, nd, box, 1, amean, a, g);
    for (int j = g.lane; j < m; j += g.size) {
      const int ja = w.ea[e0 + j], jb = w.eb[e0 + j];
      if (ja != jb && inside(w, nb, ja, box) && inside(w, nb, jb, box)) {
        const double d = in.e_attr[(int64_t)w.esort[e0 + j] * in.A + acol] - dmean;
        a.s[1] += d * d;
      }
    }
    a = group_reduce(a, gsh, g);
    const double avar = a.s[0] / (double)acnt, dvar = a.s[1] / (double)mk;

    if (g.lane == 0) {
      w.csurv[slot] = 1;
      w.cm[slot] = (int)mk;
      w.cms[slot] = (int)msk;
      const bool hit = sel.iou > 0.7;
      w.cgt[slot] = hit ? (int)gsel : -1;
      w.clabel[slot] = hit ? (int)in.gt_labels[gsel] : in.n_classes - 1;
      w.chas[slot] = sel.ios > 0.7 ? 1 : 0;
      double* st = w.cstats + kStats * slot;
      st[0] = (double)w.ccnt[slot]; st[1] = (double)mk; st[2] = (double)n_eq; st[3] = (double)n_less;
      st[4] = (double)n_more; st[5] = width; st[6] = height; st[7] = amean; st[8] = amax; st[9] = amin;
      st[10] = sqrt(avar); st[11] = dmean; st[12] = sqrt(dvar);
    }
  }
  __syncthreads();

  // -- the component's survivors in candidate order: offsets inside the component's block, totals, and the root =
  //    first maximum of the box area (:726-728)
  I4 carry = i4_zero();
  Acc best;
  acc_init(best);
  for (int base = 0; base < ncand; base += bd) {
    const int k = base + tid;
    I4 mine = i4_zero();
    if (k < ncand && w.csurv[slot0 + k]) {
      const int64_t slot = slot0 + k;
      mine.v[0] = w.ccnt[slot]; mine.v[1] = w.cm[slot]; mine.v[2] = w.cms[slot]; mine.v[3] = 1;
      const int32_t* q = w.cbox + 4 * slot;
      const double area = (xv[q[2]] - xv[q[0]]) * (yv[q[3]] - yv[q[1]]);
      if (area > best.key) { best.key = area; best.arg = k; }   // k ascends per thread: keeps the first maximum
    }
    I4 tot;
    const I4 ex = block_excl_scan_i4(mine, sh_i4, &tot);
    if (k < ncand)
      for (int t = 0; t < 4; ++t) w.coff[4 * (slot0 + k) + t] = carry.v[t] + ex.v[t];
    for (int t = 0; t < 4; ++t) carry.v[t] += tot.v[t];
  }
  best = block_reduce(best, sh_acc);                       // (its barriers also publish coff)
  if (tid == 0) {
    for (int t = 0; t < 4; ++t) w.cc_tot[4 * c + t] = carry.v[t];
    w.cc_root[c] = carry.v[3] > 0 ? w.coff[4 * (slot0 + best.arg) + 3] : -1;   // ordinal of the root among the survivors
    if (carry.v[3] == 0) raise_err(w.totals, YOLAT_PROP_ERR_NO_PROPOSAL, c);   // np.argmax of an empty area list raises (:728)
  }
}

// ---- (4a) offsets of every component's block of proposals ---------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_prop_cc_scan(In in, Ws w) {
  __shared__ long long sh[kThreads];
  const int tid = threadIdx.x, bd = blockDim.x;
  for (int q = 0; q < 4; ++q) {                            // nodes, shape edges, super edges, proposals
    long long carry = 0;
    for (int64_t base = 0; base < in.ncc; base += bd) {
      const int64_t c = base + tid;
      const long long v = c < in.ncc ? (long long)w.cc_tot[4 * c + q] : 0;
      long long tot;
      const long long ex = block_excl_scan<long long>(v, sh, &tot);
      if (c < in.ncc) w.cc_off[4 * c + q] = carry + ex;
      carry += tot;
    }
    if (tid == 0) w.totals[YOLAT_PROP_T_NODES + q] = carry;   // T_NODES, T_EDGES, T_SUPER, T_BOXES are consecutive
  }
}

// ---- (4b) one CTA per component writes its proposals (graph_dict3.py:598-603, :693-722) ---------------------------
__device__ void fill_edges(const Ws& w, int64_t nb, const Box& box, int m, const int32_t* sorted, const int32_t* la,
                           const int32_t* lb, const double* attr_in, int A, int64_t node_off, int64_t edge_off,
                           const int32_t* lrank, int64_t* edge_out, double* attr_out, int* sh, const Group& g) {
  int done = 0;
  for (int base = 0; base < m; base += g.size) {
    const int j = base + g.lane;
    int keep = 0, a = 0, b = 0;
    if (j < m) {
      a = la[j]; b = lb[j];
      keep = (a != b && inside(w, nb, a, box) && inside(w, nb, b, box)) ? 1 : 0;
    }
    int tot;
    const int ex = group_excl_scan(keep, sh, g, &tot);
    if (keep) {
      const int64_t o = edge_off + done + ex;
      edge_out[2 * o] = node_off + lrank[a];
      edge_out[2 * o + 1] = node_off + lrank[b];
      const double* src = attr_in + (int64_t)sorted[j] * A;
      double* dst = attr_out + o * A;
      for (int q = 0; q < A; ++q) dst[q] = src[q];
    }
    done += tot;
  }
}

__global__ void __launch_bounds__(kThreads) k_prop_fill(In in, Ws w, Out out, int wmax) {
  __shared__ int sh_all[kThreads];
  const int c = blockIdx.x, tid = threadIdx.x;
  const Group g = this_group();
  int* sh = sh_all + g.id * g.size;
  const int64_t nb = in.cc_ptr[c];
  const int nc = (int)(in.cc_ptr[c + 1] - nb);
  const int e0 = w.eptr[c], m = w.eptr[c + 1] - e0;
  const int s0 = w.sptr[c], ms = w.sptr[c + 1] - s0;
  const int64_t slot0 = (int64_t)c * wmax;
  const int ncand = w.ncand[c];
  const int64_t cc_node = w.cc_off[4 * c], cc_edge = w.cc_off[4 * c + 1], cc_sup = w.cc_off[4 * c + 2];
  const int64_t cc_box = w.cc_off[4 * c + 3];
  if (tid == 0) {
    out.cc_table[3 * c] = cc_box;                          // first proposal of the component
    out.cc_table[3 * c + 1] = w.cc_tot[4 * c + 3];         // how many
    out.cc_table[3 * c + 2] = cc_box + (w.cc_root[c] < 0 ? 0 : w.cc_root[c]);   // its root (largest box)
    if (c == 0) { out.slice_pos[0] = 0; out.slice_edge[0] = 0; out.slice_super[0] = 0; out.slice_bbox[0] = 0; }
  }
  const double* xv = w.xv + nb;
  const double* yv = w.yv + nb;
  int32_t* lrank = w.lrank + (int64_t)g.id * in.cc_total + nb;      // the group's own copy
  for (int k = g.id; k < ncand; k += g.count) {              // one group per surviving candidate
    const int64_t slot = slot0 + k;
    if (!w.csurv[slot]) continue;
    const int32_t* q = w.cbox + 4 * slot;
    const Box box = {q[0], q[1], q[2], q[3]};
    const double pb[4] = {xv[box.x0], yv[box.y0], xv[box.x1], yv[box.y1]};
    const double width = pb[2] - pb[0], height = pb[3] - pb[1];
    const int32_t* co = w.coff + 4 * slot;
    const int64_t node_off = cc_node + co[0], edge_off = cc_edge + co[1], sup_off = cc_sup + co[2];
    const int64_t box_off = cc_box + co[3];
    // nodes in ascending id order (the sorted tuple of :555), rank inside the proposal = the local o2n of :579-581
    int done = 0;
    for (int base = 0; base < nc; base += g.size) {
      const int r = base + g.lane;
      const int keep = (r < nc && inside(w, nb, r, box)) ? 1 : 0;
      int tot;
      const int ex = group_excl_scan(keep, sh, g, &tot);
      if (keep) {
        const int lr = done + ex;
        lrank[r] = lr;
        const int v = w.cs[nb + r];
        const int64_t o = node_off + lr;
        double px = w.pos[2 * v], py = w.pos[2 * v + 1];
        if (in.normalize_bbox) {                           // :693-701
          px = (px - pb[0]) / width;
          py = (py - pb[1]) / height;
        }
        out.pos[2 * o] = px;
        out.pos[2 * o + 1] = py;
        out.is_super[o] = w.issup[v];
        out.bbox_idx[o] = box_off;
      }
      done += tot;
    }
    __syncwarp();                                          // lrank visible to the group's edge writers
    fill_edges(w, nb, box, m, w.esort + e0, w.ea + e0, w.eb + e0, in.e_attr, in.A, node_off, edge_off, lrank, out.edge,
               out.e_attr, sh, g);
    fill_edges(w, nb, box, ms, w.ssort + s0, w.sa + s0, w.sb + s0, in.e_attr_super, in.As, node_off, sup_off, lrank,
               out.edge_super, out.e_attr_super, sh, g);
    if (g.lane == 0) {
      const int gt = w.cgt[slot];
      out.labels[box_off] = w.clabel[slot];
      out.has_obj[box_off] = w.chas[slot];
      for (int t = 0; t < 4; ++t) {
        out.bbox[4 * box_off + t] = pb[t];
        out.bbox_targets[4 * box_off + t] = gt >= 0 ? in.gt_bbox[4 * (int64_t)gt + t] : 0.0;
      }
      for (int t = 0; t < kStats; ++t) out.stat_feats[kStats * box_off + t] = w.cstats[kStats * slot + t];
      out.slice_pos[box_off + 1] = node_off + w.ccnt[slot];
      out.slice_edge[box_off + 1] = edge_off + w.cm[slot];
      out.slice_super[box_off + 1] = sup_off + w.cms[slot];
      out.slice_bbox[box_off + 1] = box_off + 1;
    }
    __syncwarp();                                          // lrank is rewritten by the group's next proposal
  }
}

static int check_in(const In* in) {
  if (!in) return YOLAT_ERR_INVALID;
  if (in->n_all < 0 || in->ncc < 0 || in->cc_total < 0 || in->E < 0 || in->Es < 0 || in->G < 0) return YOLAT_ERR_INVALID;
  if (in->sampling_step < 1 || in->sampling_step + 2 > kMaxLines) return YOLAT_ERR_UNSUPPORTED;
  if (in->A < 1 || in->As < 1 || in->n_classes < 1) return YOLAT_ERR_INVALID;
  if (in->n_all >= (1ll << 31) || in->cc_total >= (1ll << 31) || in->E > (int64_t)kIdMask || in->Es > (int64_t)kIdMask)
    return YOLAT_ERR_UNSUPPORTED;
  if (in->ncc * slots_per_cc(in->sampling_step) >= (1ll << 40)) return YOLAT_ERR_UNSUPPORTED;
  return YOLAT_OK;
}
static bool null_inputs(const In* in) {
  return (in->n_all && (!in->pos || !in->is_control || !in->is_super)) || !in->cc_ptr || (in->cc_total && !in->cc_idx) ||
         (in->E && (!in->edge || !in->e_attr)) || (in->Es && (!in->edge_super || !in->e_attr_super)) ||
         (in->G && (!in->gt_bbox || !in->gt_labels));
}

}  // namespace prop
}  // namespace yolat

using namespace yolat;
using namespace yolat::prop;

extern "C" int64_t yolat_proposals_ws_bytes(const YolatProposalIn* in) {
  if (check_in(in) != YOLAT_OK) return -1;
  return carve(*in, nullptr, nullptr);
}

extern "C" int yolat_proposals_count(const YolatProposalIn* in, void* ws, int64_t ws_bytes, int64_t* totals,
                                     void* stream) {
  YOLAT_TRY(check_in(in));
  if (!ws || !totals || null_inputs(in)) return YOLAT_ERR_INVALID;
  Ws w;
  if (carve(*in, (char*)ws, &w) > ws_bytes) return YOLAT_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int wmax = (int)slots_per_cc(in->sampling_step);
  if (cudaMemsetAsync(w.totals, 0, 8 * YOLAT_PROP_TOTALS, st) != cudaSuccess) return YOLAT_ERR_LAUNCH;
  if (cudaMemsetAsync(w.totals + YOLAT_PROP_T_ERR_CC, 0x7f, 8, st) != cudaSuccess) return YOLAT_ERR_LAUNCH;
  int64_t work = 2 * (in->E > in->Es ? in->E : in->Es);
  if (in->cc_total > work) work = in->cc_total;
  int64_t gs = cdiv(work + 1, kThreads);                  // grid-stride kernels: at most 4 CTAs per SM
  if (gs > 4 * kNumSMs) gs = 4 * kNumSMs;
  PROP_LAUNCH(k_prop_o2n, 1, kThreads, st, *in, w);
  YOLAT_CHECK_LAUNCH();
  PROP_LAUNCH(k_prop_remap, gs, kThreads, st, *in, w);
  YOLAT_CHECK_LAUNCH();
  if (in->ncc > 0) {
    PROP_LAUNCH(k_prop_cc_nodes, in->ncc, kThreads, st, *in, w);
    YOLAT_CHECK_LAUNCH();
  }
  PROP_LAUNCH(k_prop_edge_count, gs, kThreads, st, *in, w);
  YOLAT_CHECK_LAUNCH();
  PROP_LAUNCH(k_prop_edge_scan, 1, kThreads, st, *in, w);
  YOLAT_CHECK_LAUNCH();
  PROP_LAUNCH(k_prop_edge_fill, gs, kThreads, st, *in, w);
  YOLAT_CHECK_LAUNCH();
  if (in->ncc > 0) {
    PROP_LAUNCH(k_prop_cc_main, in->ncc, kThreads, st, *in, w, wmax);
    YOLAT_CHECK_LAUNCH();
  }
  PROP_LAUNCH(k_prop_cc_scan, 1, kThreads, st, *in, w);
  YOLAT_CHECK_LAUNCH();
  if (cudaMemcpyAsync(totals, w.totals, 8 * YOLAT_PROP_TOTALS, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return YOLAT_ERR_LAUNCH;
  return YOLAT_OK;
}

extern "C" int yolat_proposals_fill(const YolatProposalIn* in, void* ws, int64_t ws_bytes, const YolatProposalOut* out,
                                    void* stream) {
  YOLAT_TRY(check_in(in));
  if (!ws || !out || null_inputs(in)) return YOLAT_ERR_INVALID;
  if (!out->slice_pos || !out->slice_edge || !out->slice_super || !out->slice_bbox || !out->cc_table)
    return YOLAT_ERR_INVALID;
  Ws w;
  if (carve(*in, (char*)ws, &w) > ws_bytes) return YOLAT_ERR_WORKSPACE;
  if (in->ncc == 0) return YOLAT_OK;
  const int wmax = (int)slots_per_cc(in->sampling_step);
  cudaStream_t st = (cudaStream_t)stream;
  PROP_LAUNCH(k_prop_fill, in->ncc, kThreads, st, *in, w, *out, wmax);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}
