// edge_fused.cu -- K-EDGE: the fused gather -> edge-MLP -> scatter kernel of GraphConv('attr_edge_gp2').
//
// Reference chain (gcn_lib/sparse/torch_vertex.py:324-337 + PyG propagate + torch_nn.py:58-68 + scatter-mean):
//     x_i, x_j = index_select ; f = cat(x_i, x_j - x_i, attr) ; z1 = Lin1(f) ; a1 = relu(bn1(z1)) ;
//     z2 = Lin2(a1) ; m = relu(bn2(z2)) ; out[i] = mean_{e -> i} m_e
// which materialises eight [E, *] tensors.  Here one persistent CTA per SM (512 threads, 128 registers each, ~220 KB
// of shared memory) owns a contiguous range of target rows (CSR slots are sorted by target) and streams it in tiles
// of 128 slots through mbarrier-connected warp roles -- no CTA-wide barrier inside the loop:
//   ring fill   (epilogue warps, ahead of their own work): dst / src byte offsets and the attribute row of every slot
//               of the next tiles -> 8-stage shared-memory ring (3-step register pipeline over the eid -> attr chain);
//   gather      (8 warps): P[dst] + Q[src] + W1c attr + b1 (Lin1 pre-reduced to node level: P = x (W1a-W1b)^T,
//               Q = x W1b^T), BN1 + ReLU in registers, 3xTF32 hi/lo split written straight into the SWIZZLE_128B
//               K-major a1 stage of the tensor core (2 stages); the P / Q rows of tile t+1 are requested slot by slot
//               while tile t is computed (8 slots x 2 rows per thread always in flight); a1 never exists in HBM;
//   MMA         (one elected lane of gather warp t % 8, polling a_full between slots): 24 tcgen05.mma.kind::tf32
//               (128 x 64 x 8; W2 hi/lo resident in shared memory, BN2 scale folded into its rows for F_AGG),
//               z2 accumulates in TMEM (2 x 64 columns);
//   epilogue    (8 warps, tcgen05.ld 32 lanes x 32 columns each):
//        F_STATS: BatchNorm-2 batch statistics in registers (training needs them before any output exists),
//        F_AGG:   BN2 + ReLU (+ edge weight) -> staging tile -> segmented mean per target row (4 threads x 16
//                 channels per row, rows finishing in the tile are written once as base + mean, the straddling row
//                 goes through a carry) -- no atomics, fixed summation order, no read-modify-write of `out`,
//        F_TAPE:  z1 / z2 for the backward pass (only when autograd needs them),
//        F_Z1:    pass A -- only the gather half runs and accumulates the BatchNorm-1 statistics of z1.
// Training forward = F_Z1 + F_STATS + F_AGG launches; nothing of size [E, C] touches HBM unless F_TAPE is set.
// Measured bound (profiles/): the SM <-> L2 path of the row gathers plus the shared-memory bandwidth of the 3xTF32
// operand reads (A and B are each read three times per k-step); see DESIGN.md section 3.
#include <cstdlib>
#include "common.cuh"
#include "tc.cuh"

namespace yolat {
namespace ef {

using namespace tc;

constexpr int C = 64;             // channels (n_filters of the README configs)
constexpr int TILE = 128;         // CSR slots per MMA tile = TMEM lanes
constexpr int E_WARPS = 8;        // warps 0..7: epilogue (TMEM lane quarter = warp & 3, column half = warp >> 2);
                                  // warps 4..7 also fill the index ring, lane 0 of warp 0 issues the tcgen05.mma
constexpr int G_WARPS = 8;        // warps 8..15: gather + Lin1 + BN1 + ReLU + 3xTF32 split
constexpr int E_THREADS = E_WARPS * 32, G_THREADS = G_WARPS * 32;
constexpr int THREADS = E_THREADS + G_THREADS;          // 512 threads x 128 registers = the whole register file
constexpr uint32_t A_KB = TILE * 128;      // one k-block (32 channels) of the a1 tile: 16 KB
constexpr uint32_t A_HI = 2 * A_KB;        // hi part (2 k-blocks): 32 KB; lo follows
constexpr uint32_t A_STAGE = 2 * A_HI;     // one a1 stage: 64 KB; two stages
constexpr uint32_t W_KB = C * 128;         // one k-block of W2: 8 KB
constexpr uint32_t W_HI = 2 * W_KB;        // 16 KB; lo follows
constexpr int LDS = C + 4;                 // padded row of the message staging tile
constexpr int RING = 8;                    // index ring stages (tiles)
constexpr int PF = 4;                      // ring steps (of two tiles) the fill runs ahead of the epilogue
constexpr int RING_BYTES = TILE * 8 + TILE * 16;          // per stage: (dst, src) int2 per slot | attr float4 per slot
constexpr uint32_t OFF_W = 2 * A_STAGE;
constexpr uint32_t OFF_STAGE = OFF_W + 2 * W_HI;
constexpr uint32_t OFF_RING = OFF_STAGE + TILE * LDS * 4;
constexpr uint32_t OFF_CARRY = OFF_RING + RING * RING_BYTES;
constexpr uint32_t OFF_BN2 = OFF_CARRY + 2 * C * 4;
constexpr uint32_t SMEM_BYTES = OFF_BN2 + 2 * C * 4 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(2 * TILE * (C + 1) * 4 <= 2 * A_STAGE, "final statistics reduction aliases the a1 ring");

enum { F_TAPE = 1, F_STATS = 2, F_AGG = 4, F_Z1 = 8 };   // F_Z1: pass A -- BatchNorm-1 statistics of z1 only (no MMA)

struct Params {
  const int32_t* rowptr; const int32_t* src; const int32_t* dst; const int32_t* eid; const float* deg_inv;
  int64_t N, E;
  const float* pq; uint32_t ldpq;  // [N, ldpq]: P at column 0, Q at column C (N * ldpq < 2^30)
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
  const float* base; int64_t ldb;   // [N, C] added to the mean (lin_r(x)); may alias out; F_AGG only
  float* out; int64_t ldo;    // [N, C] = base + mean, F_AGG only: every row of the CTA's range is written once
};

// first row whose slots start at or after slot s (rows never straddle CTAs)
__device__ __forceinline__ int row_at_or_after(const Params& p, int64_t s) {
  if (s <= 0) return 0;
  if (s >= p.E) return (int)p.N;
  const int v = p.dst[s];
  return (p.rowptr[v] == (int32_t)s) ? v : v + 1;
}

// Persistent, warp-specialised: one CTA per SM owns a contiguous, row-aligned range of CSR slots and walks it in
// tiles of 128 slots through a pipeline of mbarrier-connected roles (no CTA-wide barrier inside the loop):
//     ring fill (idx, attr; 3 tiles ahead)  ->  gather warps  --a1 stage (smem, UMMA layout, 2 stages)-->
//     tcgen05.mma  --z2 accumulator (TMEM, 2 x 64 columns)-->  epilogue warps (statistics | segmented mean | tape)
template <int FLAGS>
__global__ void __launch_bounds__(THREADS, 1) k_edge_fused(const Params p) {
  constexpr bool FOLD = (FLAGS & F_AGG) && !(FLAGS & F_TAPE);   // the tape needs the unscaled z2
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_a_full[2], bar_a_empty[2], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint64_t bar_ring_full[RING], bar_ring_empty[RING];
  __shared__ uint32_t tmem_slot;

  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = ((raw_u32 + 1023u) & ~1023u) - raw_u32;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_u32 = raw_u32 + pad;
  uint8_t* w_tile = sm + OFF_W;
  float* stage = reinterpret_cast<float*>(sm + OFF_STAGE);
  uint8_t* ring = sm + OFF_RING;
  float* carry = reinterpret_cast<float*>(sm + OFF_CARRY);       // [2][C]
  float* bn2_s = reinterpret_cast<float*>(sm + OFF_BN2);         // [2][C]: sc2 | b2*sc2 + sh2

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_a_full[i]), G_THREADS);
      mbar_init(smem_u32(&bar_a_empty[i]), 1);
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), E_THREADS);
    }
#pragma unroll
    for (int i = 0; i < RING; ++i) {
      mbar_init(smem_u32(&bar_ring_full[i]), TILE);
      mbar_init(smem_u32(&bar_ring_empty[i]), G_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 2 * C);

  // W2 (B operand, K-major: row n = output channel, 64 k) -> hi/lo, resident for the whole kernel
  if (!(FLAGS & F_Z1))
  for (int idx = tid; idx < C * 16; idx += THREADS) {
    const int n = idx >> 4, c = idx & 15;
    float4 v = __ldg(reinterpret_cast<const float4*>(p.w2 + n * C + c * 4));
    if (FOLD) {                                    // BN2 scale folded into the weight rows: acc = sc2 * (a1 W2^T)
      const float sc = __ldg(p.stat2 + n);
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    }
    const uint32_t off = (uint32_t)(c >> 3) * W_KB + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u +
                         (uint32_t)(((c & 7) ^ (n & 7)) << 4);
    store_split(w_tile, w_tile + W_HI, off, v);
  }
  if ((FLAGS & F_AGG) && tid < C) {
    const float sc = __ldg(p.stat2 + tid), sh = __ldg(p.stat2 + C + tid);
    bn2_s[tid] = sc;
    bn2_s[C + tid] = fmaf(p.b2 ? __ldg(p.b2 + tid) : 0.f, sc, sh);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  // slot range of this CTA, snapped to row boundaries
  const int r_begin = row_at_or_after(p, p.E * (int64_t)blockIdx.x / gridDim.x);
  const int r_end = row_at_or_after(p, p.E * (int64_t)(blockIdx.x + 1) / gridDim.x);
  const int64_t s_begin = r_begin < p.N ? p.rowptr[r_begin] : p.E;
  const int64_t s_end = r_end < p.N ? p.rowptr[r_end] : p.E;
  const int ntiles = (int)((s_end - s_begin + TILE - 1) / TILE);

  // ---- tcgen05.mma issue: z2 = a1 W2^T, 3xTF32, accumulator (tile & 1) in TMEM.  Called by a whole (converged) warp
  //      once a_full[t & 1] has completed; one elected lane issues.  The duty rotates over the gather warps: warp t % 8
  //      polls a_full (non-blocking) between the slots of tile t+1, so the product of tile t starts as soon as its
  //      last gather warp has arrived and overlaps both the gathers of tile t+1 and the epilogue of tile t-1.
  auto issue_mma = [&](int t) {
    constexpr uint32_t IDESC = make_idesc(TILE, C, 0, 0);
    const int s = t & 1;
    const uint32_t a_u32 = sm_u32 + (uint32_t)s * A_STAGE, w_u32 = sm_u32 + OFF_W;
    if (t >= 2) mbar_wait(smem_u32(&bar_acc_empty[s]), (uint32_t)(((t >> 1) - 1) & 1));
    tc_fence_after();
    const uint32_t d = tmem_d + (uint32_t)(s * C);
    if (elect_one_sync()) {
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
        umma_tf32(d, a_lo, b_hi, IDESC, (kb > 0 || ks > 0) ? 1u : 0u);
        umma_tf32(d, a_hi, b_lo, IDESC, 1u);
        umma_tf32(d, a_hi, b_hi, IDESC, 1u);
      }
    }
    umma_commit(smem_u32(&bar_acc_full[s]));    // z2 of this tile is complete ...
    umma_commit(smem_u32(&bar_a_empty[s]));     // ... and its a1 stage may be refilled
    }
    __syncwarp();
  };

  if (warp < E_WARPS) {
    // =========================== epilogue warps ==========================================================
    const int q = warp & 3, h = warp >> 2, slot = q * 32 + lane, et = tid;

    // ---- index ring fill (one slot per thread; warps 0..3 serve the even tiles, warps 4..7 the odd ones), a 3-step
    //      register pipeline so that neither of the two dependent global loads (edge id -> attribute row) is waited for.
    //      With T(j) = 2j + h:  step j: store tile T(j-2) | load attr of T(j-1) (its eid arrived during the previous
    //      step) | load idx of T(j)
    const uint32_t ldpq_b = p.ldpq * 4u;             // the ring carries byte offsets of the P / Q rows
    int2 ds0 = make_int2(0, 0), ds1 = make_int2(0, 0);
    int e1 = 0;
    float4 at0 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ring_step = [&](int j) {
      // (a) store tile T(j - 2)
      const int ts = 2 * (j - 2) + h;
      if (ts >= 0 && ts < ntiles) {
        const int st = ts & (RING - 1);
        if (ts >= RING) mbar_wait(smem_u32(&bar_ring_empty[st]), (uint32_t)(((ts / RING) - 1) & 1));
        uint8_t* rg = ring + st * RING_BYTES;
        reinterpret_cast<int2*>(rg)[slot] = ds0;
        reinterpret_cast<float4*>(rg + TILE * 8)[slot] = at0;
        mbar_arrive(smem_u32(&bar_ring_full[st]));
      }
      // (b) attribute row of tile T(j - 1)
      ds0 = ds1;
      at0 = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        const int64_t s = s_begin + (int64_t)(2 * (j - 1) + h) * TILE + slot;
        if (j >= 1 && s < s_end) at0 = __ldg(reinterpret_cast<const float4*>(p.attr + (int64_t)e1 * 4));
      }
      // (c) indices of tile T(j) (padding slots read node 0 / edge 0 and are masked downstream)
      {
        const int64_t s = s_begin + (int64_t)(2 * j + h) * TILE + slot;
        ds1 = make_int2(0, 0);
        e1 = 0;
        if (s < s_end) {
          ds1 = make_int2((int)((uint32_t)__ldg(p.dst + s) * ldpq_b), (int)((uint32_t)__ldg(p.src + s) * ldpq_b));
          e1 = __ldg(p.eid + s);
        }
      }
    };
    for (int j = 0; j < PF; ++j) ring_step(j);       // tiles 0 .. 2 PF - 5 are in the ring, two more pairs in flight

    if (FLAGS & F_Z1) {                              // pass A: no accumulator to drain, only keep the ring filled
      for (int j = PF; 2 * (j - 2) + h < ntiles; ++j) ring_step(j);
    } else {
    float st_s[32], st_ss[32];
    if (FLAGS & F_STATS) {
#pragma unroll
      for (int i = 0; i < 32; ++i) { st_s[i] = 0.f; st_ss[i] = 0.f; }
    }
    // F_AGG: row bookkeeping of the next tile is fetched one tile ahead
    const int ch = (et & 3) * 4;                     // 4 threads per row, channels ch + 16k (k = 0..3): 64 rows in flight,
                                                     // 64-byte contiguous per k across the 4 threads (no bank conflicts)
    int R_prev = r_begin;
    int R_cur = 0;
    int nb = 0, ne_ = 0;                             // rowptr[r], rowptr[r + 1] of this thread's first row of the tile
    float ndi = 0.f;
    float4 nbase[4];
    auto row_meta = [&](int r, int& b, int& e, float& di, float4 (&bs)[4], bool want) {
      b = 0; e = 0; di = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) bs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (want && r < p.N) {
        b = __ldg(p.rowptr + r);
        e = __ldg(p.rowptr + r + 1);
        di = __ldg(p.deg_inv + r);
        if (p.base) {
#pragma unroll
          for (int k = 0; k < 4; ++k) bs[k] = __ldg(reinterpret_cast<const float4*>(p.base + (int64_t)r * p.ldb + ch + 16 * k));
        }
      }
    };
    auto tile_rcur = [&](int t) {                    // first row that is not finished by the end of tile t
      const int64_t s1 = s_begin + (int64_t)(t + 1) * TILE;
      return (s1 >= s_end) ? r_end : __ldg(p.dst + s1);
    };
    if ((FLAGS & F_AGG) && ntiles > 0) {
      R_cur = tile_rcur(0);
      row_meta(R_prev + (et >> 2), nb, ne_, ndi, nbase, true);
    }

    for (int it = 0; it < ntiles; ++it) {
      const int a = it & 1;
      const int64_t s0 = s_begin + (int64_t)it * TILE;
      const int nvalid = (int)min((int64_t)TILE, s_end - s0);
      if ((it & 1) == h) ring_step((it >> 1) + PF);   // stores tile it + 2 PF - 4
      mbar_wait_sleep(smem_u32(&bar_acc_full[a]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_d + (uint32_t)(a * C) + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32), v);
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_acc_empty[a]));     // the accumulator is free for tile it + 2

      if (FLAGS & F_TAPE) {
        if (slot < nvalid) {
          float* d = p.z2 + (s0 + slot) * C + h * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.b2) b = __ldg(reinterpret_cast<const float4*>(p.b2 + h * 32 + i));
            *reinterpret_cast<float4*>(d + i) = make_float4(v[i] + b.x, v[i + 1] + b.y, v[i + 2] + b.z, v[i + 3] + b.w);
          }
        }
      }
      if (FLAGS & F_STATS) {
        // raw accumulator sums over the valid slots; the bias enters analytically at the end
        if (slot < nvalid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { st_s[i] += v[i]; st_ss[i] = fmaf(v[i], v[i], st_ss[i]); }
        }
      }
      if (FLAGS & F_AGG) {
        float wgt = 1.f;
        if (p.ew && slot < nvalid) wgt = __ldg(p.ew + __ldg(p.eid + s0 + slot));
        named_bar_sync(1, E_THREADS);               // the row loop of the previous tile has finished reading `stage`
        float* d = stage + slot * LDS + h * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 sh = *reinterpret_cast<const float4*>(bn2_s + C + h * 32 + i);
          float4 m;
          if (FOLD) {
            m.x = fmaxf(v[i] + sh.x, 0.f);
            m.y = fmaxf(v[i + 1] + sh.y, 0.f);
            m.z = fmaxf(v[i + 2] + sh.z, 0.f);
            m.w = fmaxf(v[i + 3] + sh.w, 0.f);
          } else {
            const float4 sc = *reinterpret_cast<const float4*>(bn2_s + h * 32 + i);
            m.x = fmaxf(fmaf(v[i], sc.x, sh.x), 0.f);
            m.y = fmaxf(fmaf(v[i + 1], sc.y, sh.y), 0.f);
            m.z = fmaxf(fmaf(v[i + 2], sc.z, sh.z), 0.f);
            m.w = fmaxf(fmaf(v[i + 3], sc.w, sh.w), 0.f);
          }
          if (p.ew) { m.x *= wgt; m.y *= wgt; m.z *= wgt; m.w *= wgt; }
          *reinterpret_cast<float4*>(d + i) = m;
        }
        named_bar_sync(1, E_THREADS);
        // Segmented mean by target row.  This tile finishes rows [R_prev, R_cur); row R_cur (if it has slots here)
        // continues in the next tile and goes to the carry.  4 threads (16 channels each) per row, 64 rows in flight;
        // the bookkeeping of each thread's first row was loaded during the previous tile.
        const int64_t s1 = s0 + nvalid;
        const bool last = s1 >= s_end;
        const float* cin = carry + (it & 1) * C;
        float* cout = carry + ((it + 1) & 1) * C;
        int b = nb, e = ne_;
        float di = ndi;
        float4 bs[4] = {nbase[0], nbase[1], nbase[2], nbase[3]};
        const int R_next = (it + 1 < ntiles) ? tile_rcur(it + 1) : r_end;     // used one tile later
        row_meta(R_cur + (et >> 2), nb, ne_, ndi, nbase, it + 1 < ntiles);    // first row of the next tile
        for (int r = R_prev + (et >> 2); r <= R_cur; r += E_THREADS / 4) {
          if (r == R_cur && last) break;
          if (r != R_prev + (et >> 2)) row_meta(r, b, e, di, bs, true);
          const bool done = r < R_cur;
          if (!done && (int64_t)b >= s1) break;     // the next row starts exactly at the tile boundary: nothing to carry
          float4 acc[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((int64_t)b < s0) {                    // the row started in an earlier tile
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] = *reinterpret_cast<const float4*>(cin + ch + 16 * k);
          }
          const int lo = (int)(max((int64_t)b, s0) - s0), hi = (int)(min((int64_t)e, s1) - s0);
          for (int j = lo; j < hi; ++j) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 m = *reinterpret_cast<const float4*>(stage + j * LDS + ch + 16 * k);
              acc[k].x += m.x; acc[k].y += m.y; acc[k].z += m.z; acc[k].w += m.w;
            }
          }
          if (done) {
            float* o = p.out + (int64_t)r * p.ldo + ch;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<float4*>(o + 16 * k) = make_float4(fmaf(acc[k].x, di, bs[k].x), fmaf(acc[k].y, di, bs[k].y),
                                                                  fmaf(acc[k].z, di, bs[k].z), fmaf(acc[k].w, di, bs[k].w));
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(cout + ch + 16 * k) = acc[k];
          }
        }
        R_prev = R_cur;
        R_cur = R_next;
      }
    }
    if ((FLAGS & F_AGG) && ntiles == 0) {           // a range of rows without a single slot: out = base
      for (int r = r_begin + (et >> 2); r < r_end; r += E_THREADS / 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.base) b0 = __ldg(reinterpret_cast<const float4*>(p.base + (int64_t)r * p.ldb + ch + 16 * k));
          *reinterpret_cast<float4*>(p.out + (int64_t)r * p.ldo + ch + 16 * k) = b0;
        }
      }
    }
    if (FLAGS & F_STATS) {
      // every MMA of this CTA has completed (the last accumulator was read above): the a1 ring is free
      float* S = reinterpret_cast<float*>(sm);
      float* SS = S + TILE * (C + 1);
      named_bar_sync(1, E_THREADS);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        S[slot * (C + 1) + h * 32 + i] = st_s[i];
        SS[slot * (C + 1) + h * 32 + i] = st_ss[i];
      }
      named_bar_sync(1, E_THREADS);
      if (et < C) {
        float s = 0.f, ss = 0.f;
#pragma unroll 8
        for (int r = 0; r < TILE; ++r) { s += S[r * (C + 1) + et]; ss += SS[r * (C + 1) + et]; }
        const float b = p.b2 ? __ldg(p.b2 + et) : 0.f;
        const float cnt = (float)(s_end - s_begin);
        // sum (acc + b) and sum (acc + b)^2 from the raw accumulator sums
        p.part[((int64_t)blockIdx.x * 2 + 0) * C + et] = fmaf(cnt, b, s);
        p.part[((int64_t)blockIdx.x * 2 + 1) * C + et] = fmaf(b, fmaf(cnt, b, 2.f * s), ss);
      }
    }
    }   // !F_Z1
  } else {
    // =========================== gather warps ============================================================
    // Thread (gc, sl): channels 4gc .. 4gc+3 of slots sl*8 .. sl*8+7 of every tile.  The P / Q rows of tile t+1 are
    // requested slot by slot while tile t is being computed ("rolling" register prefetch: 8 slots per thread are
    // always in flight, across tile boundaries).  Straight-line code, 32-bit row offsets.
    const int g = tid - E_THREADS, gc = g & 15, sl = g >> 4;
    float w1c[4][4], bias1[4], sc1[4], sh1[4];
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      const int c = gc * 4 + qq;
#pragma unroll
      for (int k = 0; k < 4; ++k) w1c[qq][k] = __ldg(p.w1c + c * p.ld1 + k);
      bias1[qq] = p.b1 ? __ldg(p.b1 + c) : 0.f;
      sc1[qq] = (FLAGS & F_Z1) ? 0.f : __ldg(p.stat1 + c);
      sh1[qq] = (FLAGS & F_Z1) ? 0.f : __ldg(p.stat1 + C + c);
    }
    float z_s[4] = {0.f, 0.f, 0.f, 0.f}, z_ss[4] = {0.f, 0.f, 0.f, 0.f};     // F_Z1 accumulators
    const char* pbase = reinterpret_cast<const char*>(p.pq + gc * 4);
    const char* qbase = reinterpret_cast<const char*>(p.pq + C + gc * 4);
    float4 qv[8], pv[8];
    auto issue = [&](int st, int i) {
      const int2 ds = reinterpret_cast<const int2*>(ring + st * RING_BYTES)[sl * 8 + i];
      // (re-loading P[dst] for every slot of a run was measured faster than de-duplicating it with a select chain)
      pv[i] = __ldg(reinterpret_cast<const float4*>(pbase + (uint32_t)ds.x));
      qv[i] = __ldg(reinterpret_cast<const float4*>(qbase + (uint32_t)ds.y));
    };
    if (ntiles > 0) {
      mbar_wait(smem_u32(&bar_ring_full[0]), 0u);
#pragma unroll
      for (int i = 0; i < 8; ++i) issue(0, i);
    }
    const int gw = g >> 5;       // gather warp index
    int pending = -1;            // tile whose MMA this warp still has to issue
    auto poll_issue = [&]() {    // warp-uniform
      if (pending >= 0 && mbar_try_wait(smem_u32(&bar_a_full[pending & 1]), (uint32_t)((pending >> 1) & 1))) {
        issue_mma(pending);
        pending = -1;
      }
    };
    // byte offset of (slot sl*8, chunk gc) in the SWIZZLE_128B K-major a1 tile; slot sl*8+i adds i*128 and flips
    // the chunk index by i
    const uint32_t off0 = (uint32_t)(gc >> 3) * A_KB + (uint32_t)sl * 1024u;
    for (int it = 0; it < ntiles; ++it) {
      const int s = it & 1;
      const int64_t s0 = s_begin + (int64_t)it * TILE;
      const int nvalid = (int)min((int64_t)TILE, s_end - s0);
      const int st_cur = it & (RING - 1), st_next = (it + 1) & (RING - 1);
      const bool have_next = it + 1 < ntiles;
      if (have_next) mbar_wait(smem_u32(&bar_ring_full[st_next]), (uint32_t)(((it + 1) / RING) & 1));
      if (!(FLAGS & F_Z1) && it >= 2) mbar_wait(smem_u32(&bar_a_empty[s]), (uint32_t)(((it >> 1) - 1) & 1));
      uint8_t* a_tile = sm + (uint32_t)s * A_STAGE;
      const float4* attr_s = reinterpret_cast<const float4*>(ring + st_cur * RING_BYTES + TILE * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int slot = sl * 8 + i;
        const float4 at = attr_s[slot];
        const float pp[4] = {pv[i].x, pv[i].y, pv[i].z, pv[i].w};
        const float qq4[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w};
        float z[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          float v = fmaf(at.x, w1c[qq][0], bias1[qq]);      // same association as pass A (k_edge_stats1)
          v = fmaf(at.y, w1c[qq][1], v);
          v = fmaf(at.z, w1c[qq][2], v);
          v = fmaf(at.w, w1c[qq][3], v);
          z[qq] = (v + pp[qq]) + qq4[qq];
        }
        if (FLAGS & F_Z1) {
          const float m = slot < nvalid ? 1.f : 0.f;      // padding slots carry node 0's rows
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const float v = z[qq] * m;
            z_s[qq] += v;
            z_ss[qq] = fmaf(v, v, z_ss[qq]);
          }
          if (have_next) issue(st_next, i);
          continue;
        }
        if ((FLAGS & F_TAPE) && slot < nvalid)
          *reinterpret_cast<float4*>(p.z1 + (s0 + slot) * C + gc * 4) = make_float4(z[0], z[1], z[2], z[3]);
        float4 a1;
        a1.x = fmaxf(fmaf(z[0], sc1[0], sh1[0]), 0.f);
        a1.y = fmaxf(fmaf(z[1], sc1[1], sh1[1]), 0.f);
        a1.z = fmaxf(fmaf(z[2], sc1[2], sh1[2]), 0.f);
        a1.w = fmaxf(fmaf(z[3], sc1[3], sh1[3]), 0.f);
        const uint32_t off = off0 + (uint32_t)i * 128u + (uint32_t)(((gc & 7) ^ i) << 4);
        store_split_fast(a_tile, a_tile + A_HI, off, a1);   // a1 >= 0 and finite
        if (have_next) issue(st_next, i);                 // refill this register slot for the next tile
        if (i & 1) poll_issue();
      }
      if (FLAGS & F_Z1) {
        mbar_arrive(smem_u32(&bar_ring_empty[st_cur]));
        continue;
      }
      fence_proxy_async_smem();                     // generic-proxy smem writes -> visible to the tensor core
      mbar_arrive(smem_u32(&bar_a_full[s]));
      mbar_arrive(smem_u32(&bar_ring_empty[st_cur]));     // idx (read one tile ago) and attr of this tile are consumed
      if (pending >= 0) {                                 // (only if this warp ran a whole tile ahead of the others)
        mbar_wait(smem_u32(&bar_a_full[pending & 1]), (uint32_t)((pending >> 1) & 1));
        issue_mma(pending);
        pending = -1;
      }
      if (gw == (it & (G_WARPS - 1))) pending = it;
    }
    if (pending >= 0) {                                   // the last tile
      mbar_wait(smem_u32(&bar_a_full[pending & 1]), (uint32_t)((pending >> 1) & 1));
      issue_mma(pending);
    }
    if (FLAGS & F_Z1) {                                   // combine the 16 slot groups; part = [sum | sum of squares]
      float* red = reinterpret_cast<float*>(sm);          // [2][16][C], the a1 ring is unused in this pass
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        red[(0 * 16 + sl) * C + gc * 4 + qq] = z_s[qq];
        red[(1 * 16 + sl) * C + gc * 4 + qq] = z_ss[qq];
      }
      named_bar_sync(2, G_THREADS);
      if (g < 2 * C) {
        const int which = g / C, c = g % C;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += red[(which * 16 + k) * C + c];
        p.part[((int64_t)blockIdx.x * 2 + which) * C + c] = acc;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 2 * C);
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
// the kernels index P / Q rows with 32-bit element offsets
bool edge_fused_fits(int64_t N, int64_t ldpq) { return N * ldpq < (1ll << 30); }   // 32-bit byte offsets

int edge_fused_grid(int64_t E) {
  const int64_t tiles = cdiv(E > 0 ? E : 1, ef::TILE);
  const int64_t cap = (int64_t)kNumSMs;           // persistent: one CTA (17 warps, ~200 KB of shared memory) per SM
  return (int)(tiles < cap ? tiles : cap);
}

int edge_stats1_grid(int64_t E) { return edge_fused_grid(E); }

// part: [edge_stats1_grid(E)][2][C] sums / sums of squares of z1 over all edges
int edge_stats1(const GraphView& g, int64_t N, int64_t E, const float* pq, int64_t ldpq, const float* attr,
                const float* w1, int Cin, const float* b1, float* part, cudaStream_t st) {
  if (E <= 0) return YOLAT_OK;
  ef::Params p{};
  p.rowptr = g.rowptr_t; p.src = g.src_t; p.dst = g.dst_t; p.eid = g.eid_t; p.deg_inv = g.deg_inv;
  p.N = N; p.E = E; p.pq = pq; p.ldpq = (uint32_t)ldpq; p.attr = attr;
  p.w1c = w1 + 2 * Cin; p.ld1 = 2 * Cin + 4; p.b1 = b1; p.part = part;
  ProfScope prof(YOLAT_PROF_EDGE_STATS1, st);
  cudaError_t e = ef::launch<ef::F_Z1>(p, edge_fused_grid(E), st);
  if (e != cudaSuccess) { set_last_error(e); return YOLAT_ERR_LAUNCH; }
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

// flags: EF_TAPE | EF_STATS | EF_AGG (common.cuh).  part: [edge_fused_grid(E)][2][C] when EF_STATS.
int edge_fused(const GraphView& g, int64_t N, int64_t E, int flags, const float* pq, int64_t ldpq, const float* attr,
               const float* w1, int Cin, const float* b1, const float* stat1, const float* w2, const float* b2,
               const float* stat2, const float* ew, float* z1, float* z2, float* part, const float* base, int64_t ldb,
               float* out, int64_t ldo, cudaStream_t st) {
  if (E <= 0) return YOLAT_OK;
  ef::Params p{};
  p.rowptr = g.rowptr_t; p.src = g.src_t; p.dst = g.dst_t; p.eid = g.eid_t; p.deg_inv = g.deg_inv;
  p.N = N; p.E = E; p.pq = pq; p.ldpq = (uint32_t)ldpq; p.attr = attr;
  p.w1c = w1 + 2 * Cin; p.ld1 = 2 * Cin + 4;
  p.b1 = b1; p.stat1 = stat1; p.w2 = w2; p.b2 = b2; p.stat2 = stat2; p.ew = ew;
  p.z1 = z1; p.z2 = z2; p.part = part; p.base = base; p.ldb = ldb; p.out = out; p.ldo = ldo;
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
