// slicing.cu -- the index arithmetic of SparseCADGCN.predict on the device.
//
// Reference: cad_recognition/architecture3cc_rpn_gp_iter2.py:153-162 / :276-290 concatenate python `range` lists of
// every selected proposal's node and edge ranges, then `build_data` (:167-234) renumbers nodes through a python dict,
// re-indexes every edge in a python loop and renumbers bbox_idx node by node with tensor compares.  Here:
//   k_expand_ranges   out[t] = start[k] + (t - prefix[k]) for the range k that holds t   (binary search per element)
//   k_o2n_scatter     o2n[pos_idx[t]] = t                                                 (the dict)
//   k_edge_renumber   edge_out[e] = (o2n[edge[edge_idx[e]][0]], o2n[edge[edge_idx[e]][1]]) (the per-edge loop)
//   k_bbox_flags + scan + k_widen: bbox_idx_out[t] = #changes of bbox_idx[pos_idx[.]] up to t (the per-node loop)
// All launches are stream-ordered and allocation-free; gathers of the float tensors stay torch index_selects.
#include "common.cuh"

namespace yolat {

__global__ void k_expand_ranges(const int64_t* __restrict__ start, const int64_t* __restrict__ prefix, int64_t K,
                                int64_t total, int64_t* __restrict__ out) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total) return;
  int64_t lo = 0, hi = K;                    // largest k with prefix[k] <= t
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= t) lo = mid; else hi = mid;
  }
  out[t] = __ldg(start + lo) + (t - __ldg(prefix + lo));
}

__global__ void k_o2n_scatter(const int64_t* __restrict__ pos_idx, int64_t Np, int64_t N_all, int32_t* __restrict__ o2n) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= Np) return;
  const int64_t o = pos_idx[t];
  if (o >= 0 && o < N_all) o2n[o] = (int32_t)t;
}

__global__ void k_edge_renumber(const int64_t* __restrict__ edge, const int64_t* __restrict__ edge_idx, int64_t Ep,
                                int64_t E_all, int64_t N_all, const int32_t* __restrict__ o2n, int64_t* __restrict__ out) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= Ep) return;
  const int64_t s = edge_idx[e];
  int64_t a = -1, b = -1;
  if (s >= 0 && s < E_all) {
    const int64_t j = edge[2 * s], i = edge[2 * s + 1];
    if (j >= 0 && j < N_all) a = o2n[j];
    if (i >= 0 && i < N_all) b = o2n[i];
  }
  out[2 * e] = a;           // -1: the edge leaves the selection (the reference's dict lookup would raise KeyError)
  out[2 * e + 1] = b;
}

__global__ void k_bbox_flags(const int64_t* __restrict__ bbox_idx, const int64_t* __restrict__ pos_idx, int64_t Np,
                             int32_t* __restrict__ flag) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= Np) return;
  flag[t] = (t > 0 && bbox_idx[pos_idx[t]] != bbox_idx[pos_idx[t - 1]]) ? 1 : 0;
}

// inclusive scan of int32 flags into int64, one CTA (the selections are 1e3..1e5 nodes)
__global__ void __launch_bounds__(1024) k_scan_widen(const int32_t* __restrict__ flag, int64_t n, int64_t* __restrict__ out) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + tid;
    const int32_t v = i < n ? flag[i] : 0;
    int32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    int32_t off = carry;
    for (int w = 0; w < wid; ++w) off += warp_tot[w];
    if (i < n) out[i] = (int64_t)(off + inc);
    __syncthreads();
    if (tid == 1023) carry = off + inc;
    __syncthreads();
  }
}


// Per-image offsets of a collated batch (cad_recognition/train.py:238-258: a python loop over the batch items adds the
// image's node offset to its edge rows and its proposal offset to its bbox_idx rows, in place, on the host).  Here the
// four slice tables travel with the batch and one launch applies both offsets on the device.
//   tab: [4][G + 1] int64 = edge slices | pos slices | bbox_idx slices | labels slices (prefix sums, tab[.][0] = 0)
__device__ __forceinline__ int64_t slice_of(const int64_t* __restrict__ bounds, int64_t G, int64_t i) {
  int64_t lo = 0, hi = G;            // largest g with bounds[g] <= i
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (bounds[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}
__global__ void k_batch_offsets(int64_t* __restrict__ edge, int64_t E, int64_t* __restrict__ bbox_idx, int64_t N,
                                const int64_t* __restrict__ tab, int64_t G) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t* t_edge = tab, *t_pos = tab + (G + 1), *t_bidx = tab + 2 * (G + 1), *t_lab = tab + 3 * (G + 1);
  if (i < E) {
    const int64_t off = t_pos[slice_of(t_edge, G, i)];
    edge[2 * i] += off;
    edge[2 * i + 1] += off;
  }
  if (i < N) bbox_idx[i] += t_lab[slice_of(t_bidx, G, i)];
}

}  // namespace yolat

using namespace yolat;

extern "C" {

int yolat_expand_ranges(const int64_t* start, const int64_t* prefix, int64_t K, int64_t total, int64_t* out, void* stream) {
  if (K < 0 || total < 0 || (total > 0 && (!start || !prefix || !out || K == 0))) return YOLAT_ERR_INVALID;
  if (total == 0) return YOLAT_OK;
  k_expand_ranges<<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(start, prefix, K, total, out);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

// ws: N_all + Np int32.  edge: [E_all, 2] contiguous int64.  edge_out: [Ep, 2], bbox_idx_out: [Np].
int64_t yolat_slice_graph_ints(int64_t N_all, int64_t Np) { return (N_all > 0 ? N_all : 0) + (Np > 0 ? Np : 0) + 8; }

int yolat_slice_graph(const int64_t* pos_idx, int64_t Np, const int64_t* edge_idx, int64_t Ep, const int64_t* edge,
                      int64_t E_all, int64_t N_all, const int64_t* bbox_idx, int32_t* ws, int64_t* edge_out,
                      int64_t* bbox_idx_out, void* stream) {
  if (Np < 0 || Ep < 0 || N_all < 0 || E_all < 0 || !ws) return YOLAT_ERR_INVALID;
  if ((Np > 0 && (!pos_idx || !bbox_idx || !bbox_idx_out)) || (Ep > 0 && (!edge_idx || !edge || !edge_out))) return YOLAT_ERR_INVALID;
  if (N_all >= (1ll << 31) - 8 || Np >= (1ll << 31) - 8) return YOLAT_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* o2n = ws;
  int32_t* flag = ws + N_all;
  if (N_all > 0) cudaMemsetAsync(o2n, 0xff, (size_t)N_all * sizeof(int32_t), st);      // -1
  if (Np > 0) {
    k_o2n_scatter<<<(unsigned)cdiv(Np, 256), 256, 0, st>>>(pos_idx, Np, N_all, o2n);
    YOLAT_CHECK_LAUNCH();
    k_bbox_flags<<<(unsigned)cdiv(Np, 256), 256, 0, st>>>(bbox_idx, pos_idx, Np, flag);
    YOLAT_CHECK_LAUNCH();
    k_scan_widen<<<1, 1024, 0, st>>>(flag, Np, bbox_idx_out);
    YOLAT_CHECK_LAUNCH();
  }
  if (Ep > 0) {
    k_edge_renumber<<<(unsigned)cdiv(Ep, 256), 256, 0, st>>>(edge, edge_idx, Ep, E_all, N_all, o2n, edge_out);
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

// edge [E, 2] and bbox_idx [N] are updated in place; tab [4][G + 1] as described at k_batch_offsets.
int yolat_batch_offsets(int64_t* edge, int64_t E, int64_t* bbox_idx, int64_t N, const int64_t* tab, int64_t G, void* stream) {
  if (E < 0 || N < 0 || G < 1 || !tab || (E > 0 && !edge) || (N > 0 && !bbox_idx)) return YOLAT_ERR_INVALID;
  const int64_t n = E > N ? E : N;
  if (n == 0) return YOLAT_OK;
  k_batch_offsets<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(edge, E, bbox_idx, N, tab, G);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

}  // extern "C"
