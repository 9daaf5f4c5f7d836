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
__global__ void __launch_bounds__(1024) k_scan_inplace(int32_t* a0, int32_t* a1, int64_t n, float* deg_inv0) {
  int32_t* a = blockIdx.x == 0 ? a0 : a1;
  float* deg_inv = blockIdx.x == 0 ? deg_inv0 : nullptr;
  __shared__ int32_t warp_off[32];
  __shared__ int32_t tile_total;
  __shared__ int32_t carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  // tiles of 1024*4 elements; each thread owns 4 consecutive elements
  for (int64_t base = 0; base < n; base += 4096) {
    int64_t i0 = base + tid * 4;
    int32_t v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = (i0 + q < n) ? a[i0 + q] : 0;
    if (deg_inv) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (i0 + q < n) deg_inv[i0 + q] = 1.0f / (float)max(v[q], 1);
    }
    const int32_t tsum = v[0] + v[1] + v[2] + v[3];
    int32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_off[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const int32_t w = warp_off[lane];
      int32_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_off[lane] = winc - w;  // exclusive offset of each warp inside the tile
      if (lane == 31) tile_total = winc;
    }
    __syncthreads();
    int32_t excl = carry + warp_off[wid] + (inc - tsum);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (i0 + q < n) a[i0 + q] = excl;
      excl += v[q];
    }
    __syncthreads();
    if (tid == 0) carry += tile_total;
    __syncthreads();
  }
  if (tid == 0) a[n] = carry;
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

__device__ __forceinline__ void insertion_sort(int32_t* a, int n) {
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
  k_scan_inplace<<<2, 1024, 0, st>>>(rowptr_t, rowptr_s, N, const_cast<float*>(v.deg_inv));
  YOLAT_CHECK_LAUNCH();
  if (E > 0 && N > 0) {
    k_fill_target<<<(unsigned)cdiv(E, T), T, 0, st>>>(edge, se, sc, E, N, rowptr_t, v.cursor_t, const_cast<int32_t*>(v.eid_t));
    YOLAT_CHECK_LAUNCH();
    k_sort_target_rows<<<(unsigned)cdiv(N, T), T, 0, st>>>(edge, se, N, rowptr_t, const_cast<int32_t*>(v.eid_t),
                                                           const_cast<int32_t*>(v.src_t), const_cast<int32_t*>(v.dst_t),
                                                           v.slot_of_edge);
    YOLAT_CHECK_LAUNCH();
    k_fill_source<<<(unsigned)cdiv(E, T), T, 0, st>>>(edge, se, sc, E, N, rowptr_s, v.cursor_s, v.slot_of_edge,
                                                      const_cast<int32_t*>(v.slot_s));
    YOLAT_CHECK_LAUNCH();
    k_sort_rows<<<(unsigned)cdiv(N, T), T, 0, st>>>(N, rowptr_s, const_cast<int32_t*>(v.slot_s));
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
  k_scan_inplace<<<1, 1024, 0, st>>>(segptr, segptr, S, nullptr);
  YOLAT_CHECK_LAUNCH();
  if (M > 0 && S > 0) {
    k_fill_index<<<(unsigned)cdiv(M, T), T, 0, st>>>(index, M, S, segptr, v.cursor, const_cast<int32_t*>(v.perm),
                                                     const_cast<int32_t*>(v.seg_of_row));
    YOLAT_CHECK_LAUNCH();
    k_sort_rows<<<(unsigned)cdiv(S, T), T, 0, st>>>(S, segptr, const_cast<int32_t*>(v.perm));
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

}  // extern "C"
