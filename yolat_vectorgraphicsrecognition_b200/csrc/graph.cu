// graph.cu -- once-per-batch graph preparation: CSR by target, CSR by source, in-degree, segments.
//
// Replaces the per-call gather/scatter bookkeeping of PyG's MessagePassing.propagate
// (gcn_lib/sparse/torch_vertex.py:324) and of torch_scatter.scatter
// (cad_recognition/architecture3cc_rpn_gp_iter2.py:67,122).  The result is deterministic: inside a
// row, slots are ordered by original edge id (rows are sorted after the atomic fill), so every
// segmented reduction downstream has a fixed summation order.
#include <atomic>
#include "common.cuh"

namespace yolat {

static thread_local cudaError_t g_last = cudaSuccess;
void set_last_error(cudaError_t e) { g_last = e; }
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- counting --------------------------------------------------------------------------------
__global__ void k_count_edges(const int64_t* __restrict__ edge, int64_t se, int64_t sc, int64_t E, int64_t N,
                              int32_t* __restrict__ cnt_t, int32_t* __restrict__ cnt_s, int32_t* __restrict__ err) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t j = edge[e * se], i = edge[e * se + sc];
  if (j < 0 || j >= N || i < 0 || i >= N) { atomicAdd(err, 1); return; }
  atomicAdd(cnt_t + i, 1);
  atomicAdd(cnt_s + j, 1);
}

__global__ void k_count_index(const int64_t* __restrict__ index, int64_t M, int64_t S, int32_t* __restrict__ cnt) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= M) return;
  int64_t s = index[r];
  if (s >= 0 && s < S) atomicAdd(cnt + s, 1);
}

// ---- exclusive scan, one CTA per array (blockIdx.x selects the array) ------------------------------
// cnt arrives in out[0..n) (counts were accumulated in place), result: out[0..n] exclusive prefix.
// Optionally writes deg_inv = 1/max(cnt,1).
template <int IPT>
__global__ void __launch_bounds__(1024) k_scan_inplace(int32_t* a0, int32_t* a1, int64_t n, float* deg_inv0) {
  int32_t* a = blockIdx.x == 0 ? a0 : a1;
  float* deg_inv = blockIdx.x == 0 ? deg_inv0 : nullptr;
  __shared__ int32_t warp_off[32];
  __shared__ int32_t tile_total;
  __shared__ int32_t carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  // Tiles of 1024 * IPT elements.  Warp w owns the contiguous range [w * 32 * IPT, (w + 1) * 32 * IPT) of the tile and
  // walks it in IPT coalesced rows of 32 (all loads issued up front); IPT is chosen so that the graphs of the timed
  // configurations (N <= 32768) are ONE tile: one round of loads, two barriers.
  for (int64_t base = 0; base < n; base += 1024 * IPT) {
    const int64_t w0 = base + (int64_t)wid * (32 * IPT);
    int32_t v[IPT];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
      const int64_t i = w0 + q * 32 + lane;
      v[q] = i < n ? a[i] : 0;
    }
    if (deg_inv) {
#pragma unroll
      for (int q = 0; q < IPT; ++q) {
        const int64_t i = w0 + q * 32 + lane;
        if (i < n) deg_inv[i] = 1.0f / (float)max(v[q], 1);
      }
    }
    // inclusive scan of every row of 32, rows chained through a running total
    int32_t inc[IPT];
    int32_t run = 0;
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
      int32_t x = v[q];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
      }
      inc[q] = run + x;
      run += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) warp_off[wid] = run;             // this warp's total
    __syncthreads();
    if (wid == 0) {
      const int32_t w = warp_off[lane];
      int32_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_off[lane] = winc - w;  // exclusive offset of each warp inside the tile
      if (lane == 31) tile_total = winc;
    }
    __syncthreads();
    const int32_t off = carry + warp_off[wid];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
      const int64_t i = w0 + q * 32 + lane;
      if (i < n) a[i] = off + inc[q] - v[q];
    }
    __syncthreads();
    if (tid == 0) carry += tile_total;
    __syncthreads();
  }
  if (tid == 0) a[n] = carry;
}
static void launch_scan(int arrays, int32_t* a0, int32_t* a1, int64_t n, float* deg_inv0, cudaStream_t st) {
  if (n <= 1024 * 4) k_scan_inplace<4><<<arrays, 1024, 0, st>>>(a0, a1, n, deg_inv0);
  else if (n <= 1024 * 8) k_scan_inplace<8><<<arrays, 1024, 0, st>>>(a0, a1, n, deg_inv0);
  else if (n <= 1024 * 16) k_scan_inplace<16><<<arrays, 1024, 0, st>>>(a0, a1, n, deg_inv0);
  else k_scan_inplace<32><<<arrays, 1024, 0, st>>>(a0, a1, n, deg_inv0);
}

// ---- fill + per-row sort -------------------------------------------------------------------------
__global__ void k_fill_target(const int64_t* __restrict__ edge, int64_t se, int64_t sc, int64_t E, int64_t N,
                              const int32_t* __restrict__ rowptr_t, int32_t* __restrict__ cursor_t,
                              int32_t* __restrict__ eid_t) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t j = edge[e * se], i = edge[e * se + sc];
  if (j < 0 || j >= N || i < 0 || i >= N) return;
  int32_t pos = rowptr_t[i] + atomicAdd(cursor_t + i, 1);
  eid_t[pos] = (int32_t)e;
}

// Rows are short (Bezier-curve graphs: 4 neighbours, proposals: ~16 nodes): up to 8 values are sorted in registers with
// one round of loads and one of stores; longer rows fall back to the in-memory insertion sort.
__device__ __forceinline__ void insertion_sort(int32_t* a, int n) {
  if (n <= 1) return;
  if (n <= 8) {
    int32_t r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = i < n ? a[i] : 0x7fffffff;
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {          // odd-even transposition network, fully unrolled
#pragma unroll
      for (int i = pass & 1; i + 1 < 8; i += 2) {
        const int32_t lo = min(r[i], r[i + 1]), hi = max(r[i], r[i + 1]);
        r[i] = lo; r[i + 1] = hi;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < n) a[i] = r[i];
    return;
  }
  for (int i = 1; i < n; ++i) {
    int32_t key = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > key) { a[j + 1] = a[j]; --j; }
    a[j + 1] = key;
  }
}

__global__ void k_sort_target_rows(const int64_t* __restrict__ edge, int64_t se, int64_t N,
                                   const int32_t* __restrict__ rowptr_t, int32_t* __restrict__ eid_t,
                                   int32_t* __restrict__ src_t, int32_t* __restrict__ dst_t,
                                   int32_t* __restrict__ slot_of_edge) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= N) return;
  int32_t b = rowptr_t[v], e = rowptr_t[v + 1];
  insertion_sort(eid_t + b, e - b);
  for (int32_t s = b; s < e; ++s) {
    int32_t id = eid_t[s];
    src_t[s] = (int32_t)edge[(int64_t)id * se];
    dst_t[s] = (int32_t)v;
    slot_of_edge[id] = s;
  }
}

__global__ void k_fill_source(const int64_t* __restrict__ edge, int64_t se, int64_t sc, int64_t E, int64_t N,
                              const int32_t* __restrict__ rowptr_s, int32_t* __restrict__ cursor_s,
                              const int32_t* __restrict__ slot_of_edge, int32_t* __restrict__ slot_s) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t j = edge[e * se], i = edge[e * se + sc];
  if (j < 0 || j >= N || i < 0 || i >= N) return;
  int32_t pos = rowptr_s[j] + atomicAdd(cursor_s + j, 1);
  slot_s[pos] = slot_of_edge[e];
}

__global__ void k_sort_rows(int64_t N, const int32_t* __restrict__ rowptr, int32_t* __restrict__ vals) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= N) return;
  int32_t b = rowptr[v];
  insertion_sort(vals + b, rowptr[v + 1] - b);
}

__global__ void k_fill_index(const int64_t* __restrict__ index, int64_t M, int64_t S,
                             const int32_t* __restrict__ segptr, int32_t* __restrict__ cursor, int32_t* __restrict__ perm,
                             int32_t* __restrict__ seg_of_row) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= M) return;
  int64_t s = index[r];
  if (s < 0 || s >= S) { seg_of_row[r] = -1; return; }
  seg_of_row[r] = (int32_t)s;
  perm[segptr[s] + atomicAdd(cursor + s, 1)] = (int32_t)r;
}

}  // namespace yolat

using namespace yolat;

extern "C" {

int yolat_abi_version(void) { return 1; }

int64_t yolat_launch_count(void) { return (int64_t)yolat::g_launches.load(std::memory_order_relaxed); }

const char* yolat_status_string(int s) {
  switch (s) {
    case YOLAT_OK: return "ok";
    case YOLAT_ERR_INVALID: return "invalid argument";
    case YOLAT_ERR_WORKSPACE: return "workspace or tape too small";
    case YOLAT_ERR_LAUNCH: return "CUDA kernel launch failed";
    case YOLAT_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

const char* yolat_last_cuda_error(void) { return cudaGetErrorString(yolat::g_last); }

int64_t yolat_graph_ints(int64_t N, int64_t E) { return graph_layout(N, E, nullptr, nullptr); }

const int32_t* yolat_graph_error_ptr(const int32_t* graph, int64_t N, int64_t E) {
  GraphView v; graph_layout(N, E, graph, &v); return v.err;
}
const int32_t* yolat_graph_rowptr(const int32_t* graph, int64_t N, int64_t E) {
  GraphView v; graph_layout(N, E, graph, &v); return v.rowptr_t;
}
const int32_t* yolat_graph_src(const int32_t* graph, int64_t N, int64_t E) {
  GraphView v; graph_layout(N, E, graph, &v); return v.src_t;
}
const int32_t* yolat_graph_eid(const int32_t* graph, int64_t N, int64_t E) {
  GraphView v; graph_layout(N, E, graph, &v); return v.eid_t;
}

int yolat_graph_build(const int64_t* edge, int64_t se, int64_t sc, int64_t E, int64_t N, int32_t* graph, void* stream) {
  if (!graph || N < 0 || E < 0 || (E > 0 && !edge)) return YOLAT_ERR_INVALID;
  if (N >= (1ll << 31) - 8 || E >= (1ll << 31) - 8) return YOLAT_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  GraphView v;
  int64_t total = graph_layout(N, E, graph, &v);
  int32_t* rowptr_t = const_cast<int32_t*>(v.rowptr_t);
  int32_t* rowptr_s = const_cast<int32_t*>(v.rowptr_s);
  // zero everything that is accumulated into (counts live in rowptr arrays, cursors, err)
  cudaMemsetAsync(graph, 0, total * sizeof(int32_t), st);
  const int T = 256;
  if (E > 0) {
    k_count_edges<<<(unsigned)cdiv(E, T), T, 0, st>>>(edge, se, sc, E, N, rowptr_t, rowptr_s, const_cast<int32_t*>(v.err));
    YOLAT_CHECK_LAUNCH();
  }
  launch_scan(2, rowptr_t, rowptr_s, N, const_cast<float*>(v.deg_inv), st);
  YOLAT_CHECK_LAUNCH();
  if (E > 0 && N > 0) {
    k_fill_target<<<(unsigned)cdiv(E, T), T, 0, st>>>(edge, se, sc, E, N, rowptr_t, v.cursor_t, const_cast<int32_t*>(v.eid_t));
    YOLAT_CHECK_LAUNCH();
    k_sort_target_rows<<<(unsigned)cdiv(N, 64), 64, 0, st>>>(edge, se, N, rowptr_t, const_cast<int32_t*>(v.eid_t),
                                                           const_cast<int32_t*>(v.src_t), const_cast<int32_t*>(v.dst_t),
                                                           v.slot_of_edge);
    YOLAT_CHECK_LAUNCH();
    k_fill_source<<<(unsigned)cdiv(E, T), T, 0, st>>>(edge, se, sc, E, N, rowptr_s, v.cursor_s, v.slot_of_edge,
                                                      const_cast<int32_t*>(v.slot_s));
    YOLAT_CHECK_LAUNCH();
    k_sort_rows<<<(unsigned)cdiv(N, 64), 64, 0, st>>>(N, rowptr_s, const_cast<int32_t*>(v.slot_s));
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

int64_t yolat_segments_ints(int64_t M, int64_t S) { return seg_layout(M, S, nullptr, nullptr); }

int yolat_segments_build(const int64_t* index, int64_t M, int64_t S, int32_t* seg, void* stream) {
  if (!seg || M < 0 || S < 0 || (M > 0 && !index)) return YOLAT_ERR_INVALID;
  if (M >= (1ll << 31) - 8 || S >= (1ll << 31) - 8) return YOLAT_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  SegView v;
  int64_t total = seg_layout(M, S, seg, &v);
  int32_t* segptr = const_cast<int32_t*>(v.segptr);
  cudaMemsetAsync(seg, 0, total * sizeof(int32_t), st);
  const int T = 256;
  if (M > 0) {
    k_count_index<<<(unsigned)cdiv(M, T), T, 0, st>>>(index, M, S, segptr);
    YOLAT_CHECK_LAUNCH();
  }
  launch_scan(1, segptr, segptr, S, nullptr, st);
  YOLAT_CHECK_LAUNCH();
  if (M > 0 && S > 0) {
    k_fill_index<<<(unsigned)cdiv(M, T), T, 0, st>>>(index, M, S, segptr, v.cursor, const_cast<int32_t*>(v.perm),
                                                     const_cast<int32_t*>(v.seg_of_row));
    YOLAT_CHECK_LAUNCH();
    k_sort_rows<<<(unsigned)cdiv(S, 64), 64, 0, st>>>(S, segptr, const_cast<int32_t*>(v.perm));
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

}  // extern "C"
