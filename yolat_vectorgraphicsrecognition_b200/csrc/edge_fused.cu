// edge_fused.cu -- K-EDGE: the fused gather -> edge-MLP -> scatter kernel of GraphConv('attr_edge_gp2').
//
// Reference chain (gcn_lib/sparse/torch_vertex.py:324-337 + PyG propagate + torch_nn.py:58-68 + scatter-mean):
//     x_i, x_j = index_select ; f = cat(x_i, x_j - x_i, attr) ; z1 = Lin1(f) ; a1 = relu(bn1(z1)) ;
//     z2 = Lin2(a1) ; m = relu(bn2(z2)) ; out[i] = mean_{e -> i} m_e
// which materialises eight [E, *] tensors.  Here one persistent CTA owns a contiguous range of target rows
// (CSR slots are sorted by target), and per tile of 128 slots:
//   1. gathers P[dst] + Q[src] + W1c attr + b1 (Lin1 pre-reduced to node level: P = x (W1a-W1b)^T, Q = x W1b^T),
//      applies BN1 + ReLU in registers, splits to 3xTF32 hi/lo and writes the 128 x 64 a1 tile straight into the
//      SWIZZLE_128B K-major shared-memory layout of the tensor core -- a1 never exists in global memory;
//   2. one thread issues 24 tcgen05.mma.kind::tf32 (128 x 64 x 64, W2 hi/lo resident in shared memory for the whole
//      kernel), accumulating z2 in TMEM;
//   3. tcgen05.ld brings z2 back; depending on the pass the epilogue
//        F_STATS: accumulates the BatchNorm-2 batch statistics (training needs them before any output exists),
//        F_AGG:   applies BN2 + ReLU (+ edge weight) and does the segmented mean over the target rows with a
//                 windowed run-length reduction -- no atomics, fixed summation order,
//        F_TAPE:  writes z1 / z2 for the backward pass (only when autograd needs them).
// Training forward = pass A (BN1 statistics, edge.cu) + this kernel with F_STATS + this kernel with F_AGG;
// nothing of size [E, C] touches HBM unless F_TAPE is set.
#include "common.cuh"
#include "tc.cuh"

namespace yolat {
namespace ef {

using namespace tc;

constexpr int C = 64;             // channels (n_filters of the README configs)
constexpr int TILE = 128;         // CSR slots per MMA tile = TMEM lanes
constexpr int THREADS = 256;
constexpr uint32_t A_KB = TILE * 128;      // one k-block (32 channels) of the a1 tile: 16 KB
constexpr uint32_t A_HI = 2 * A_KB;        // hi part (2 k-blocks): 32 KB; lo follows
constexpr uint32_t W_KB = C * 128;         // one k-block of W2: 8 KB
constexpr uint32_t W_HI = 2 * W_KB;        // 16 KB; lo follows
constexpr int LDS = C + 4;                 // padded row of the z2 staging tile (aliases the a1 tile)
constexpr uint32_t SMEM_BYTES = 2 * A_HI + 2 * W_HI + 1024;

enum { F_TAPE = 1, F_STATS = 2, F_AGG = 4 };

struct Params {
  const int32_t* rowptr; const int32_t* src; const int32_t* dst; const int32_t* eid; const float* deg_inv;
  int64_t N, E;
  const float* pq;            // [N, 2C]: P | Q
  const float* attr;          // [E, 4] original edge order
  const float* w1c; int ld1;  // W1[:, 2Cin:2Cin+4], row stride ld1
  const float* b1;            // [C] or null
  const float* stat1;         // BN1 (sc | sh)
  const float* w2;            // [C, C]
  const float* b2;            // [C] or null
  const float* stat2;         // BN2 (sc | sh), F_AGG only
  const float* ew;            // [E] or null
  float* z1; float* z2;       // tape [E, C] in slot order, F_TAPE only
  float* part;                // [gridDim.x][2][C], F_STATS only
  float* out; int64_t ldo;    // [N, C] += mean, F_AGG only
};

// first row whose slots start at or after slot s (rows never straddle CTAs)
__device__ __forceinline__ int row_at_or_after(const Params& p, int64_t s) {
  if (s <= 0) return 0;
  if (s >= p.E) return (int)p.N;
  const int v = p.dst[s];
  return (p.rowptr[v] == (int32_t)s) ? v : v + 1;
}

template <int FLAGS>
__global__ void __launch_bounds__(THREADS, 2) k_edge_fused(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  __shared__ int32_t dst_s[TILE];
  __shared__ float ew_s[TILE];
  __shared__ float red[2][THREADS];
  __shared__ int32_t bnd_node[4][2];
  __shared__ float bnd_val[4][2][C];

  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = ((raw_u32 + 1023u) & ~1023u) - raw_u32;
  uint8_t* a_tile = smem_raw + pad;                 // a1 hi | lo, later the z2 staging tile
  uint8_t* w_tile = a_tile + 2 * A_HI;              // W2 hi | lo
  const uint32_t a_u32 = raw_u32 + pad, w_u32 = a_u32 + 2 * A_HI;
  float* stage = reinterpret_cast<float*>(a_tile);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(smem_u32(&mbar), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), C);

  // W2 (B operand, K-major: row n = output channel, 64 k) -> hi/lo, resident for the whole kernel
  for (int idx = tid; idx < C * 16; idx += THREADS) {
    const int n = idx >> 4, c = idx & 15;
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.w2 + n * C + c * 4));
    const uint32_t off = (uint32_t)(c >> 3) * W_KB + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u +
                         (uint32_t)(((c & 7) ^ (n & 7)) << 4);
    store_split(w_tile, w_tile + W_HI, off, v);
  }

  // per-thread constants of the gather phase: this thread always produces channels 4*gc .. 4*gc+3
  const int gc = tid & 15;
  float w1c[4][4], bias1[4], sc1[4], sh1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int ch = gc * 4 + q;
#pragma unroll
    for (int k = 0; k < 4; ++k) w1c[q][k] = __ldg(p.w1c + ch * p.ld1 + k);
    bias1[q] = p.b1 ? __ldg(p.b1 + ch) : 0.f;
    sc1[q] = __ldg(p.stat1 + ch);
    sh1[q] = __ldg(p.stat1 + C + ch);
  }
  // per-thread constants of the epilogue: TMEM lane quarter q, column half h -> columns h*32 .. h*32+31
  const int eq = warp & 3, eh = warp >> 2;
  // per-thread constants of the consumers: channel cc, slot window cg
  const int cc = tid & 63, cg = tid >> 6;
  const float b2c = p.b2 ? __ldg(p.b2 + cc) : 0.f;
  float sc2 = 0.f, sh2 = 0.f;
  if (FLAGS & F_AGG) { sc2 = __ldg(p.stat2 + cc); sh2 = __ldg(p.stat2 + C + cc); }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  constexpr uint32_t IDESC = make_idesc(TILE, C, 0, 0);

  // slot range of this CTA, snapped to row boundaries
  const int r_begin = row_at_or_after(p, p.E * (int64_t)blockIdx.x / gridDim.x);
  const int r_end = row_at_or_after(p, p.E * (int64_t)(blockIdx.x + 1) / gridDim.x);
  const int64_t s_begin = r_begin < p.N ? p.rowptr[r_begin] : p.E;
  const int64_t s_end = r_end < p.N ? p.rowptr[r_end] : p.E;

  float st_s = 0.f, st_ss = 0.f;          // F_STATS accumulators (channel cc, window cg)
  int carry_node = -1;                    // F_AGG: row whose slots continue across windows / tiles (threads < C)
  float carry_acc = 0.f;
  uint32_t phase = 0;

  for (int64_t s0 = s_begin; s0 < s_end; s0 += TILE) {
    const int nvalid = (int)min((int64_t)TILE, s_end - s0);

    // ---- 1. gather + Lin1 (node-level P/Q) + BN1 + ReLU -> a1 tile in UMMA layout ---------------------
#pragma unroll
    for (int half = 0; half < 2; ++half) {     // two rounds of 4 slots per thread: 12 gathers in flight, <= 128 registers
      int dsts[4], srcs[4], eids[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int slot = (tid >> 4) + 16 * (half * 4 + i);
        const bool ok = slot < nvalid;
        dsts[i] = ok ? __ldg(p.dst + s0 + slot) : -1;
        srcs[i] = ok ? __ldg(p.src + s0 + slot) : 0;
        eids[i] = ok ? __ldg(p.eid + s0 + slot) : 0;
      }
      float4 pv[4], qv[4], av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pv[i] = qv[i] = av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dsts[i] >= 0) {
          pv[i] = __ldg(reinterpret_cast<const float4*>(p.pq + (int64_t)dsts[i] * (2 * C) + gc * 4));
          qv[i] = __ldg(reinterpret_cast<const float4*>(p.pq + (int64_t)srcs[i] * (2 * C) + C + gc * 4));
          av[i] = __ldg(reinterpret_cast<const float4*>(p.attr + (int64_t)eids[i] * 4));
        }
      }
      if (gc == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int slot = (tid >> 4) + 16 * (half * 4 + i);
          dst_s[slot] = dsts[i];
          if (FLAGS & F_AGG) ew_s[slot] = (p.ew && dsts[i] >= 0) ? __ldg(p.ew + eids[i]) : 1.f;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int slot = (tid >> 4) + 16 * (half * 4 + i);
        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (dsts[i] >= 0) {
          float z[4];
          const float pp[4] = {pv[i].x, pv[i].y, pv[i].z, pv[i].w};
          const float qq[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v = (pp[q] + bias1[q]) + qq[q];       // same association as edge.cu::k_edge_z1 (pass A statistics)
            v = fmaf(av[i].x, w1c[q][0], v);
            v = fmaf(av[i].y, w1c[q][1], v);
            v = fmaf(av[i].z, w1c[q][2], v);
            v = fmaf(av[i].w, w1c[q][3], v);
            z[q] = v;
          }
          if (FLAGS & F_TAPE)
            *reinterpret_cast<float4*>(p.z1 + (s0 + slot) * C + gc * 4) = make_float4(z[0], z[1], z[2], z[3]);
          a1.x = fmaxf(fmaf(z[0], sc1[0], sh1[0]), 0.f);
          a1.y = fmaxf(fmaf(z[1], sc1[1], sh1[1]), 0.f);
          a1.z = fmaxf(fmaf(z[2], sc1[2], sh1[2]), 0.f);
          a1.w = fmaxf(fmaf(z[3], sc1[3], sh1[3]), 0.f);
        }
        const uint32_t off = (uint32_t)(gc >> 3) * A_KB + (uint32_t)(slot >> 3) * 1024u + (uint32_t)(slot & 7) * 128u +
                             (uint32_t)(((gc & 7) ^ (slot & 7)) << 4);
        store_split(a_tile, a_tile + A_HI, off, a1);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    // ---- 2. z2 = a1 W2^T on the tensor core (3xTF32), accumulator in TMEM -----------------------------
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ao = (uint32_t)kb * A_KB + (uint32_t)ks * 32u;
          const uint32_t bo = (uint32_t)kb * W_KB + (uint32_t)ks * 32u;
          const uint64_t a_hi = make_desc(a_u32 + ao, 16, 1024, LAYOUT_SW128);
          const uint64_t a_lo = make_desc(a_u32 + A_HI + ao, 16, 1024, LAYOUT_SW128);
          const uint64_t b_hi = make_desc(w_u32 + bo, 16, 1024, LAYOUT_SW128);
          const uint64_t b_lo = make_desc(w_u32 + W_HI + bo, 16, 1024, LAYOUT_SW128);
          umma_tf32(tmem_d, a_lo, b_hi, IDESC, (kb > 0 || ks > 0) ? 1u : 0u);
          umma_tf32(tmem_d, a_hi, b_lo, IDESC, 1u);
          umma_tf32(tmem_d, a_hi, b_hi, IDESC, 1u);
        }
      }
      umma_commit(smem_u32(&mbar));
    }
    mbar_wait(smem_u32(&mbar), phase);
    phase ^= 1u;
    tc_fence_after();

    // ---- 3. TMEM -> registers -> staging tile (aliases the a1 tile: every MMA that read it has completed) ----
    {
      float v[32];
      tmem_ld32(tmem_d + ((uint32_t)(eq * 32) << 16) + (uint32_t)(eh * 32), v);
      float* d = stage + (eq * 32 + lane) * LDS + eh * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(d + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    tc_fence_before();
    __syncthreads();

    // ---- 4. consumers of the z2 tile ---------------------------------------------------------------------
    if (FLAGS & F_TAPE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int slot = (tid >> 4) + 16 * i;
        if (slot < nvalid) {
          float4 v = *reinterpret_cast<const float4*>(stage + slot * LDS + gc * 4);
          if (p.b2) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.b2 + gc * 4));
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          }
          *reinterpret_cast<float4*>(p.z2 + (s0 + slot) * C + gc * 4) = v;
        }
      }
    }
    if (FLAGS & F_STATS) {
      const int j_end = min(nvalid, cg * 32 + 32);
#pragma unroll 8
      for (int j = cg * 32; j < j_end; ++j) {
        const float z = stage[j * LDS + cc] + b2c;
        st_s += z;
        st_ss = fmaf(z, z, st_ss);
      }
    }
    if (FLAGS & F_AGG) {
      // windowed run-length reduction: window cg = slots [32cg, 32cg+32); complete rows inside a window are
      // written directly, the first / last run of every window go through the boundary records below.
      const int j_end = min(nvalid, cg * 32 + 32);
      int cur = -1, first_node = -1;
      float acc = 0.f, first_acc = 0.f;
      bool have_first = false;
      for (int j = cg * 32; j < j_end; ++j) {
        const int d = dst_s[j];
        const float m = fmaxf(fmaf(stage[j * LDS + cc] + b2c, sc2, sh2), 0.f) * ew_s[j];
        if (d != cur) {
          if (cur >= 0) {
            if (!have_first) { first_node = cur; first_acc = acc; have_first = true; }
            else p.out[(int64_t)cur * p.ldo + cc] += acc * __ldg(p.deg_inv + cur);
          }
          cur = d;
          acc = 0.f;
        }
        acc += m;
      }
      int last_node = -1;
      float last_acc = 0.f;
      if (cur >= 0) {
        if (!have_first) { first_node = cur; first_acc = acc; }
        else { last_node = cur; last_acc = acc; }
      }
      if (cc == 0) { bnd_node[cg][0] = first_node; bnd_node[cg][1] = last_node; }
      bnd_val[cg][0][cc] = first_acc;
      bnd_val[cg][1][cc] = last_acc;
      __syncthreads();
      if (tid < C) {   // merge the boundary records in slot order with the carry of the previous windows / tiles
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int node = bnd_node[r >> 1][r & 1];
          if (node < 0) continue;
          const float v = bnd_val[r >> 1][r & 1][tid];
          if (node != carry_node) {
            if (carry_node >= 0) p.out[(int64_t)carry_node * p.ldo + tid] += carry_acc * __ldg(p.deg_inv + carry_node);
            carry_node = node;
            carry_acc = v;
          } else {
            carry_acc += v;
          }
        }
      }
    }
    __syncthreads();   // staging tile / dst_s are rewritten by the next tile
  }

  if (FLAGS & F_AGG) {
    if (tid < C && carry_node >= 0) p.out[(int64_t)carry_node * p.ldo + tid] += carry_acc * __ldg(p.deg_inv + carry_node);
  }
  if (FLAGS & F_STATS) {
    red[0][tid] = st_s; red[1][tid] = st_ss;
    __syncthreads();
    if (tid < C) {
#pragma unroll
      for (int g = 1; g < 4; ++g) { st_s += red[0][g * 64 + tid]; st_ss += red[1][g * 64 + tid]; }
      p.part[((int64_t)blockIdx.x * 2 + 0) * C + tid] = st_s;
      p.part[((int64_t)blockIdx.x * 2 + 1) * C + tid] = st_ss;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, C);
}

// ---- pass A: BatchNorm-1 batch statistics of z1 (gather only, nothing stored) ------------------------------
// thread = (16-byte channel chunk gc, slot lane); 4 slots per thread in flight; per-thread channel sums are
// combined per CTA and written as one partial [2][C] (fp64 combine in k_bn_finalize).
constexpr int S1_THREADS = 256;
__global__ void __launch_bounds__(S1_THREADS, 3) k_edge_stats1(const Params p) {
  __shared__ float red[2][S1_THREADS / 16][C];
  const int tid = threadIdx.x, gc = tid & 15, sl = tid >> 4;
  float w1c[4][4], bias1[4], s[4], ss[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int ch = gc * 4 + q;
#pragma unroll
    for (int k = 0; k < 4; ++k) w1c[q][k] = __ldg(p.w1c + ch * p.ld1 + k);
    bias1[q] = p.b1 ? __ldg(p.b1 + ch) : 0.f;
    s[q] = 0.f; ss[q] = 0.f;
  }
  const int64_t s_begin = p.E * (int64_t)blockIdx.x / gridDim.x;
  const int64_t s_end = p.E * (int64_t)(blockIdx.x + 1) / gridDim.x;
  for (int64_t s0 = s_begin; s0 < s_end; s0 += 64) {
    int dsts[4], srcs[4], eids[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t slot = s0 + sl + 16 * i;
      const bool ok = slot < s_end;
      dsts[i] = ok ? __ldg(p.dst + slot) : -1;
      srcs[i] = ok ? __ldg(p.src + slot) : 0;
      eids[i] = ok ? __ldg(p.eid + slot) : 0;
    }
    float4 pv[4], qv[4], av[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      pv[i] = qv[i] = av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (dsts[i] >= 0) {
        pv[i] = __ldg(reinterpret_cast<const float4*>(p.pq + (int64_t)dsts[i] * (2 * C) + gc * 4));
        qv[i] = __ldg(reinterpret_cast<const float4*>(p.pq + (int64_t)srcs[i] * (2 * C) + C + gc * 4));
        av[i] = __ldg(reinterpret_cast<const float4*>(p.attr + (int64_t)eids[i] * 4));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (dsts[i] < 0) continue;
      const float pp[4] = {pv[i].x, pv[i].y, pv[i].z, pv[i].w};
      const float qq[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = (pp[q] + bias1[q]) + qq[q];
        v = fmaf(av[i].x, w1c[q][0], v);
        v = fmaf(av[i].y, w1c[q][1], v);
        v = fmaf(av[i].z, w1c[q][2], v);
        v = fmaf(av[i].w, w1c[q][3], v);
        s[q] += v;
        ss[q] = fmaf(v, v, ss[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) { red[0][sl][gc * 4 + q] = s[q]; red[1][sl][gc * 4 + q] = ss[q]; }
  __syncthreads();
  if (tid < 2 * C) {
    const int which = tid / C, c = tid % C;
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < S1_THREADS / 16; ++k) a += red[which][k][c];
    p.part[((int64_t)blockIdx.x * 2 + which) * C + c] = a;
  }
}

template <int FLAGS>
static cudaError_t launch(const Params& p, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_fused<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_edge_fused<FLAGS><<<grid, THREADS, SMEM_BYTES, st>>>(p);
  return cudaSuccess;
}

}  // namespace ef

bool edge_fused_supported(int C) { return C == ef::C; }

int edge_fused_grid(int64_t E) {
  const int64_t tiles = cdiv(E > 0 ? E : 1, ef::TILE);
  const int64_t cap = 2 * (int64_t)kNumSMs;       // two resident CTAs per SM (98 KB of shared memory each)
  return (int)(tiles < cap ? tiles : cap);
}

int edge_stats1_grid(int64_t E) {
  const int64_t need = cdiv(E > 0 ? E : 1, 64);
  const int64_t cap = 3 * (int64_t)kNumSMs;
  return (int)(need < cap ? need : cap);
}

// part: [edge_stats1_grid(E)][2][C] sums / sums of squares of z1 over all edges
int edge_stats1(const GraphView& g, int64_t N, int64_t E, const float* pq, const float* attr, const float* w1, int Cin,
                const float* b1, float* part, cudaStream_t st) {
  if (E <= 0) return YOLAT_OK;
  ef::Params p{};
  p.rowptr = g.rowptr_t; p.src = g.src_t; p.dst = g.dst_t; p.eid = g.eid_t; p.deg_inv = g.deg_inv;
  p.N = N; p.E = E; p.pq = pq; p.attr = attr;
  p.w1c = w1 + 2 * Cin; p.ld1 = 2 * Cin + 4; p.b1 = b1; p.part = part;
  ProfScope prof(YOLAT_PROF_EDGE_STATS1, st);
  ef::k_edge_stats1<<<edge_stats1_grid(E), ef::S1_THREADS, 0, st>>>(p);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

// flags: EF_TAPE | EF_STATS | EF_AGG (common.cuh).  part: [edge_fused_grid(E)][2][C] when EF_STATS.
int edge_fused(const GraphView& g, int64_t N, int64_t E, int flags, const float* pq, const float* attr, const float* w1,
               int Cin, const float* b1, const float* stat1, const float* w2, const float* b2, const float* stat2,
               const float* ew, float* z1, float* z2, float* part, float* out, int64_t ldo, cudaStream_t st) {
  if (E <= 0) return YOLAT_OK;
  ef::Params p{};
  p.rowptr = g.rowptr_t; p.src = g.src_t; p.dst = g.dst_t; p.eid = g.eid_t; p.deg_inv = g.deg_inv;
  p.N = N; p.E = E; p.pq = pq; p.attr = attr;
  p.w1c = w1 + 2 * Cin; p.ld1 = 2 * Cin + 4;
  p.b1 = b1; p.stat1 = stat1; p.w2 = w2; p.b2 = b2; p.stat2 = stat2; p.ew = ew;
  p.z1 = z1; p.z2 = z2; p.part = part; p.out = out; p.ldo = ldo;
  const int grid = edge_fused_grid(E);
  cudaError_t e;
  ProfScope prof((flags & EF_AGG) ? YOLAT_PROF_EDGE_FUSED_AGG : YOLAT_PROF_EDGE_FUSED_STATS, st);
  switch (flags) {
    case EF_STATS: e = ef::launch<ef::F_STATS>(p, grid, st); break;
    case EF_STATS | EF_TAPE: e = ef::launch<ef::F_STATS | ef::F_TAPE>(p, grid, st); break;
    case EF_AGG: e = ef::launch<ef::F_AGG>(p, grid, st); break;
    case EF_AGG | EF_TAPE: e = ef::launch<ef::F_AGG | ef::F_TAPE>(p, grid, st); break;
    default: return YOLAT_ERR_INVALID;
  }
  if (e != cudaSuccess) { set_last_error(e); return YOLAT_ERR_LAUNCH; }
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

}  // namespace yolat
