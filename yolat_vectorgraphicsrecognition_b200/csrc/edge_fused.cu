// edge_fused.cu -- K-EDGE: the fused gather -> edge-MLP -> scatter kernel of GraphConv('attr_edge_gp2').
//
// Reference chain (gcn_lib/sparse/torch_vertex.py:324-337 + PyG propagate + torch_nn.py:58-68 + scatter-mean):
//     x_i, x_j = index_select ; f = cat(x_i, x_j - x_i, attr) ; z1 = Lin1(f) ; a1 = relu(bn1(z1)) ;
//     z2 = Lin2(a1) ; m = relu(bn2(z2)) ; out[i] = mean_{e -> i} m_e
// which materialises eight [E, *] tensors.  Here one persistent CTA per SM (640 threads, ~220 KB of shared memory) owns
// a contiguous range of target rows (CSR slots are sorted by target) and streams it in tiles of 128 slots through
// mbarrier-connected warp roles -- no CTA-wide barrier inside the loop:
//   control     (one warp of a fourth warp group; setmaxnreg gives its registers to the gather warps):
//        ring:  slot-ordered records (k_edge_records, once per layer call: P / Q byte offsets | attribute row, 16 B
//               each) reach a 6-stage shared-memory ring as two TMA bulk copies per tile (cp.async.bulk + mbarrier
//               complete_tx), five tiles ahead -- no thread touches an index on the way;
//        MMA:   waits for the a1 stage of tile t and issues 24 tcgen05.mma.kind::tf32 (128 x 64 x 8; W2 hi/lo resident
//               in shared memory, BN2 scale folded into its rows for F_AGG); z2 accumulates in TMEM (2 x 64 columns).
//               The issue blocks for the ~1.3k cycles the tensor core needs (its operand reads are shared-memory
//               bound), which is why no working warp does it (profiles/r1_t_edge_fused_agg_role_trace.txt);
//   gather      (8 warps; thread = 4 channels x 8 slots of every tile): P[dst] + Q[src] + W1c attr + b1 (Lin1
//               pre-reduced to node level: P = x (W1a-W1b)^T, Q = x W1b^T), BN1 + ReLU in packed fp32 pairs
//               (FFMA2 / FADD2), truncating 3xTF32 hi/lo split written straight into the SWIZZLE_128B K-major a1
//               stage of the tensor core (2 stages); the P / Q rows of tile t+1 are requested slot by slot while
//               tile t is computed (8 slots x 2 rows per thread always in flight = 64 KB per SM) and prefetched into
//               L2 three tiles ahead; straight-line code with no polling inside; a1 never exists in HBM;
//   epilogue    (8 warps; thread = one slot = one TMEM lane, one column half):
//        drain: tcgen05.ld (32 lanes x 32 columns) -> staging tile [slot][channel] in shared memory, then
//        F_STATS: BatchNorm-2 batch statistics as column sums of the staging tile (training needs them before any
//                 output exists),
//        F_AGG:   BN2 + ReLU (+ edge weight) -> segmented mean per target row (8 threads x 8 channels per row, rows
//                 finishing in the tile are written once as base + mean, the straddling row goes through a carry)
//                 -- no atomics, fixed summation order, no read-modify-write of `out`,
//        F_TAPE:  z1 / z2 for the backward pass (only when autograd needs them),
//        F_Z1:    pass A -- only the gather half runs and accumulates the BatchNorm-1 statistics of z1.
// Training forward = F_Z1 + F_STATS + F_AGG launches; nothing of size [E, C] touches HBM unless F_TAPE is set.
// Other role splits (4 epilogue x 16 gather warps; MMAs issued by epilogue warp 0 or piece-wise by the gather warps)
// are template parameters kept for experiments (YOLAT_EF_ROLES, YOLAT_EF_MG).  Measured bound: DESIGN.md section 3.1.
#include <cstdlib>
#include "common.cuh"
#include "tc.cuh"

namespace yolat {
namespace ef {

using namespace tc;

constexpr int C = 64;             // channels (n_filters of the README configs)
constexpr int TILE = 128;         // CSR slots per MMA tile = TMEM lanes
// Warp roles are template parameters <EW, GW>: EW epilogue warps (4: one thread per slot draining both column halves;
// 8: TMEM lane quarter = warp & 3, column half = warp >> 2) and GW gather warps (16 threads per slot, 64 / GW slots per
// thread and tile).  Warps 0..3 also fill the index ring, warp 0 issues the MMAs.
constexpr uint32_t A_KB = TILE * 128;      // one k-block (32 channels) of the a1 tile: 16 KB
constexpr uint32_t A_HI = 2 * A_KB;        // hi part (2 k-blocks): 32 KB; lo follows
constexpr uint32_t A_STAGE = 2 * A_HI;     // one a1 stage: 64 KB; two stages
constexpr uint32_t W_KB = C * 128;         // one k-block of W2: 8 KB
constexpr uint32_t W_HI = 2 * W_KB;        // 16 KB; lo follows
constexpr int LDS = C + 4;                 // padded row of the staging tile
constexpr int RING = 6;                    // record ring stages (tiles)
constexpr int PF = RING - 1;               // tiles the TMA producer runs ahead of the tile being drained
constexpr int PFL2 = 3;                    // tiles the L2 prefetch of the P / Q rows runs ahead of the gather
constexpr int RING_BYTES = TILE * 16 + TILE * 16;         // per stage: (P off, Q off, eid, 0) int4 per slot | attr float4 per slot
constexpr uint32_t OFF_W = 2 * A_STAGE;
constexpr uint32_t OFF_STAGE = OFF_W + 2 * W_HI;
constexpr uint32_t OFF_RING = OFF_STAGE + TILE * LDS * 4;
constexpr uint32_t OFF_CARRY = OFF_RING + RING * RING_BYTES;
constexpr uint32_t OFF_BN2 = OFF_CARRY + 2 * C * 4;
constexpr uint32_t SMEM_BYTES = OFF_BN2 + 2 * C * 4 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

enum { F_TAPE = 1, F_STATS = 2, F_AGG = 4, F_Z1 = 8 };

// -DYOLAT_EF_TRACE: per-tile clock64 timestamps of the roles of one CTA (tools/ef_trace.py); never in the product build.
#ifdef YOLAT_EF_TRACE
__device__ long long g_ef_trace[128 * 16];
#define EF_TRACE(tile, ev) do { if (blockIdx.x == 74 && (tile) < 128) g_ef_trace[(tile) * 16 + (ev)] = clock64(); } while (0)
#else
#define EF_TRACE(tile, ev) do { } while (0)
#endif   // F_Z1: pass A -- BatchNorm-1 statistics of z1 only (no MMA)

struct Params {
  const int32_t* rowptr; const int32_t* src; const int32_t* dst; const int32_t* eid; const float* deg_inv;
  int64_t N, E;
  const float* pq; uint32_t ldpq;  // [N, ldpq]: P at column 0, Q at column C (N * ldpq < 2^30)
  const int4* rec_idx;        // [E] slot order: (byte offset of P[dst], byte offset of Q[src] relative to P, eid, 0)
  const float4* rec_attr;     // [E] slot order: attribute row of the slot's edge
  const float* w1c; int ld1;  // W1[:, 2Cin:2Cin+4], row stride ld1
  const float* b1;            // [C] or null
  const float* stat1;         // BN1 (sc | sh)
  const float* w2;            // [C, C]
  const float* b2;            // [C] or null
  const float* stat2;         // BN2 (sc | sh), F_AGG only
  const float* ew;            // [E] or null
  float* z1; float* z2;       // tape [E, C] in slot order, F_TAPE only
  float* part;                // [gridDim.x][2][C], F_STATS / F_Z1
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

// one arrival per warp: every lane's earlier shared-memory accesses are ordered before it by the warp barrier
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// Persistent, warp-specialised: one CTA per SM owns a contiguous, row-aligned range of CSR slots and walks it in
// tiles of 128 slots through a pipeline of mbarrier-connected roles (no CTA-wide barrier inside the loop):
//     ring fill (idx, attr; PF - 4 tiles ahead)  ->  gather warps  --a1 stage (smem, UMMA layout, 2 stages)-->
//     tcgen05.mma  --z2 accumulator (TMEM, 2 x 64 columns)-->  epilogue warps (statistics | segmented mean | tape)
// MMA_MODE: who issues the tcgen05.mma of a tile (the issue loop blocks for the ~1.2k cycles the tensor core needs for a
// tile's 24 MMAs -- their operand reads are shared-memory bound -- so whoever issues cannot do anything else meanwhile):
//   0 epilogue warp 0, 1 a rotating gather warp, 2 a dedicated control warp (a fourth warp group of which one warp
//   works; register budgets are rebalanced with setmaxnreg: control 96 -> 32, gather 96 -> 128 per thread; the pool only
//   holds what the control group released, so the sum after must not exceed 20 warps x 96).
template <int FLAGS, int E_WARPS, int G_WARPS, int MMA_MODE>
__global__ void __launch_bounds__((E_WARPS + G_WARPS + (MMA_MODE == 2 ? 4 : 0)) * 32, 1) k_edge_fused(const Params p) {
  constexpr bool MMA_G = MMA_MODE == 1, MMA_E = MMA_MODE == 0, CTRL = MMA_MODE == 2;
  constexpr int E_THREADS = E_WARPS * 32, G_THREADS = G_WARPS * 32, THREADS = E_THREADS + G_THREADS + (CTRL ? 128 : 0);
  constexpr int SPT = TILE * 16 / G_THREADS;              // slots per gather thread and tile (16 threads per slot)
  constexpr int NSLG = TILE / SPT;                        // slot groups: slot = i * NSLG + sl
  constexpr int NSG = E_THREADS / 16;                     // F_STATS: slot groups of the column sums
  constexpr int TPR = E_THREADS / 32;                     // F_AGG: threads per target row
  constexpr int RIF = 32;                                 // F_AGG: rows per sweep of the epilogue threads
  constexpr int NPF = CTRL ? 1 : 2;                       // F_AGG: sweeps whose row bookkeeping is prefetched a tile ahead
                                                          // (one in control-warp mode: the epilogue has 96 registers there)
  constexpr int CPT = 16 / TPR;                           // F_AGG: 16-byte chunks per thread (chunk k at ch + 4 TPR k)
  static_assert(E_WARPS == 4 || E_WARPS == 8, "epilogue mapping");
  static_assert(SPT * G_THREADS == TILE * 16 && (SPT == 4 || SPT == 8), "gather mapping");
  static_assert(2 * (G_THREADS / 16) * C * 4 <= 2 * A_STAGE, "pass A reduction aliases the a1 ring");
  static_assert(2 * NSG * C * 4 <= TILE * LDS * 4, "statistics reduction aliases the staging tile");
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
      mbar_init(smem_u32(&bar_a_full[i]), G_WARPS);
      mbar_init(smem_u32(&bar_a_empty[i]), 1);
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), E_THREADS);
    }
#pragma unroll
    for (int i = 0; i < RING; ++i) {
      mbar_init(smem_u32(&bar_ring_full[i]), 1);
      mbar_init(smem_u32(&bar_ring_empty[i]), G_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 2 * C);

  // The slots of a partial last tile keep whatever the stage held before: zeros (node 0) until the ring has wrapped.
  for (int idx = tid; idx < RING * RING_BYTES / 16; idx += THREADS)
    reinterpret_cast<int4*>(ring)[idx] = make_int4(0, 0, 0, 0);

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
  fence_proxy_async_smem();                        // generic-proxy writes (ring zeros, W2) before async-proxy accesses
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

  // ---- tcgen05.mma issue: z2 = a1 W2^T, 3xTF32, accumulator (tile & 1) in TMEM.  Called by one whole warp, one
  //      elected lane issues the whole tile and blocks until the tensor core has taken all 24 MMAs (~1.3k cycles).
  //      Used by epilogue warp 0 (MMA_MODE 0, one tile ahead of its drain), by the control warp (MODE 2) and for the
  //      last tile of MODE 1; otherwise MODE 1 issues piece-wise (issue_kstep below).
  auto issue_mma = [&](int t) {
    constexpr uint32_t IDESC = make_idesc(TILE, C, 0, 0);
    const int s = t & 1;
    const uint32_t a_u32 = sm_u32 + (uint32_t)s * A_STAGE, w_u32 = sm_u32 + OFF_W;
    mbar_wait_park(smem_u32(&bar_a_full[s]), (uint32_t)((t >> 1) & 1));
    if (lane == 0) EF_TRACE(t, 11);
    if (t >= 2) mbar_wait_park(smem_u32(&bar_acc_empty[s]), (uint32_t)(((t >> 1) - 1) & 1));
    if (lane == 0) EF_TRACE(t, 12);
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

  // One k-step (8 of the 64 reduction indices: three tcgen05.mma) of tile t -- the piece-wise form of issue_mma used by the
  // gather warp on MMA duty: issued between its slots of the next tile, the tensor core's queue never fills and the
  // issuing lane never blocks (a whole tile issued at once blocks for the ~1.3k cycles the MMAs take).
  auto issue_kstep = [&](int t, int k) {
    constexpr uint32_t IDESC = make_idesc(TILE, C, 0, 0);
    const int s = t & 1, kb = k >> 2, ks = k & 3;
    const uint32_t a_u32 = sm_u32 + (uint32_t)s * A_STAGE, w_u32 = sm_u32 + OFF_W;
    const uint32_t d = tmem_d + (uint32_t)(s * C);
    if (elect_one_sync()) {
      const uint32_t ao = (uint32_t)kb * A_KB + (uint32_t)ks * 32u;
      const uint32_t bo = (uint32_t)kb * W_KB + (uint32_t)ks * 32u;
      const uint64_t a_hi = make_desc(a_u32 + ao, 16, 1024, LAYOUT_SW128);
      const uint64_t a_lo = make_desc(a_u32 + A_HI + ao, 16, 1024, LAYOUT_SW128);
      const uint64_t b_hi = make_desc(w_u32 + bo, 16, 1024, LAYOUT_SW128);
      const uint64_t b_lo = make_desc(w_u32 + W_HI + bo, 16, 1024, LAYOUT_SW128);
      umma_tf32(d, a_lo, b_hi, IDESC, k > 0 ? 1u : 0u);
      umma_tf32(d, a_hi, b_lo, IDESC, 1u);
      umma_tf32(d, a_hi, b_hi, IDESC, 1u);
      if (k == 7) {
        umma_commit(smem_u32(&bar_acc_full[s]));    // z2 of this tile is complete ...
        umma_commit(smem_u32(&bar_a_empty[s]));     // ... and its a1 stage may be refilled
      }
    }
    __syncwarp();
  };

  if (warp < E_WARPS) {
    // =========================== epilogue warps ==========================================================
    const int q = warp & 3, slot = q * 32 + lane, et = tid;
    const int h_begin = E_WARPS == 8 ? (warp >> 2) : 0, h_end = E_WARPS == 8 ? h_begin + 1 : 2;

    // ---- record ring producer (thread 0): two TMA bulk copies per tile, completion counted in bytes on
    //      ring_full[stage]; a stage is refilled once every gather warp has released it.
    auto fill = [&](int j) {
      if (j >= ntiles) return;
      const int st = j % RING;
      if (j >= RING) mbar_wait_park(smem_u32(&bar_ring_empty[st]), (uint32_t)(((j / RING) - 1) & 1));
      const int64_t s0 = s_begin + (int64_t)j * TILE;
      const uint32_t bytes = (uint32_t)min((int64_t)TILE, s_end - s0) * 16u;
      const uint32_t bar = smem_u32(&bar_ring_full[st]);
      const uint32_t dst = smem_u32(ring + st * RING_BYTES);
      mbar_expect_tx(bar, 2u * bytes);
      bulk_g2s(dst, p.rec_idx + s0, bytes, bar);
      bulk_g2s(dst + TILE * 16, p.rec_attr + s0, bytes, bar);
    };
    const bool producer = !CTRL && (tid == 0);
    if (producer) {
      for (int j = 0; j < PF; ++j) fill(j);
    }
    if (FLAGS & F_Z1) {                              // pass A: no accumulator to drain, only keep the ring filled
      if (producer) {
        for (int j = PF; j < ntiles; ++j) fill(j);
      }
    } else {
    float2 st_s[2], st_ss[2];                        // F_STATS: channels 4 cg .. 4 cg + 3 of TILE / NSG slots
    st_s[0] = st_s[1] = st_ss[0] = st_ss[1] = make_float2(0.f, 0.f);
    const int cg = et & 15, sp = et >> 4;
    // F_AGG: TPR threads per row, this thread's chunks are channels ch + 4 TPR k (k < CPT): 16 TPR contiguous bytes per
    // k across the threads of a row (no bank conflicts); row bookkeeping of the next tile is fetched one tile ahead
    const int ch = (et % TPR) * 4;
    const int rl = et / TPR;                         // row lane: 0 .. RIF - 1
    int R_prev = r_begin;
    int R_cur = 0;
    int nb[NPF], ne_[NPF];                           // rowptr[r], rowptr[r + 1] of this thread's first NPF rows of the tile
    float ndi[NPF];
    float4 nbase[NPF][CPT];
#pragma unroll
    for (int q2 = 0; q2 < NPF; ++q2) { nb[q2] = 0; ne_[q2] = 0; ndi[q2] = 0.f; }
    auto row_meta = [&](int r, int& b, int& e, float& di, float4 (&bs)[CPT], bool want) {
      b = 0; e = 0; di = 0.f;
#pragma unroll
      for (int k = 0; k < CPT; ++k) bs[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (want && r < p.N) {
        b = __ldg(p.rowptr + r);
        e = __ldg(p.rowptr + r + 1);
        di = __ldg(p.deg_inv + r);
        if (p.base) {
#pragma unroll
          for (int k = 0; k < CPT; ++k)
            bs[k] = __ldg(reinterpret_cast<const float4*>(p.base + (int64_t)r * p.ldb + ch + 4 * TPR * k));
        }
      }
    };
    auto tile_rcur = [&](int t) {                    // first row that is not finished by the end of tile t
      const int64_t s1 = s_begin + (int64_t)(t + 1) * TILE;
      return (s1 >= s_end) ? r_end : __ldg(p.dst + s1);
    };
    if ((FLAGS & F_AGG) && ntiles > 0) {
      R_cur = tile_rcur(0);
#pragma unroll
      for (int q2 = 0; q2 < NPF; ++q2) row_meta(R_prev + rl + q2 * RIF, nb[q2], ne_[q2], ndi[q2], nbase[q2], true);
    }

    if (MMA_E && warp == 0 && ntiles > 0) issue_mma(0);
    constexpr bool SH_REG = FOLD && E_WARPS == 8 && !CTRL;   // one column half per thread: its BN2 shifts live in registers
    float shr[SH_REG ? 32 : 1];
    if (SH_REG) {
#pragma unroll
      for (int i = 0; i < 32; ++i) shr[SH_REG ? i : 0] = bn2_s[C + h_begin * 32 + i];
    }
    for (int t = 0; t < ntiles; ++t) {
      if (producer) fill(t + PF);                    // never blocks in practice: the gather finished tile t - 1 long ago
      if (MMA_E && warp == 0 && t + 1 < ntiles) issue_mma(t + 1);
      const int a = t & 1;
      const int64_t s0 = s_begin + (int64_t)t * TILE;
      const int nvalid = (int)min((int64_t)TILE, s_end - s0);
      float wgt = 1.f;
      if ((FLAGS & F_AGG) && p.ew && slot < nvalid) wgt = __ldg(p.ew + __ldg(p.eid + s0 + slot));
      // F_AGG: the row bookkeeping of the next tile is requested here, a whole drain ahead of its first use (the
      // boundary row R_next of tile t + 1 is needed when this iteration ends: requested after the row sums it used to
      // cost a global-memory round trip per tile).  The current tile's copy moves to c* first.
      int cb[NPF], ce[NPF], R_next = r_end;
      float cdi[NPF];
      float4 cbs[NPF][CPT];
      if (FLAGS & F_AGG) {
#pragma unroll
        for (int q2 = 0; q2 < NPF; ++q2) {
          cb[q2] = nb[q2]; ce[q2] = ne_[q2]; cdi[q2] = ndi[q2];
#pragma unroll
          for (int k = 0; k < CPT; ++k) cbs[q2][k] = nbase[q2][k];
        }
        if (t + 1 < ntiles) R_next = tile_rcur(t + 1);
#pragma unroll
        for (int q2 = 0; q2 < NPF; ++q2)
          row_meta(R_cur + rl + q2 * RIF, nb[q2], ne_[q2], ndi[q2], nbase[q2], t + 1 < ntiles);
      }
      if (tid == 0) EF_TRACE(t, 0);
      mbar_wait_park(smem_u32(&bar_acc_full[a]), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      if (tid == 0) EF_TRACE(t, 1);
      named_bar_sync(1, E_THREADS);                 // the readers of the previous tile's staging rows have finished
      if (tid == 0) EF_TRACE(t, 2);
      for (int h = h_begin; h < h_end; ++h) {
        float v[32];
        tmem_ld32(tmem_d + (uint32_t)(a * C) + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 32), v);
        if (h == h_end - 1) {
          tc_fence_before();
          mbar_arrive(smem_u32(&bar_acc_empty[a]));   // the accumulator is free for tile t + 2
        }
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
        float* d = stage + slot * LDS + h * 32;
        if (FLAGS & F_STATS) {
          // raw accumulator values of the valid slots (zeros elsewhere); the bias enters analytically at the end
          const float m = slot < nvalid ? 1.f : 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(d + i) = make_float4(v[i] * m, v[i + 1] * m, v[i + 2] * m, v[i + 3] * m);
        }
        if (FLAGS & F_AGG) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 sh;
            if (SH_REG) sh = make_float4(shr[SH_REG ? i : 0], shr[SH_REG ? i + 1 : 0], shr[SH_REG ? i + 2 : 0], shr[SH_REG ? i + 3 : 0]);
            else sh = *reinterpret_cast<const float4*>(bn2_s + C + h * 32 + i);
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
        }
      }
      if (tid == 0) EF_TRACE(t, 3);
      named_bar_sync(1, E_THREADS);
      if (tid == 0) EF_TRACE(t, 4);
      if (FLAGS & F_STATS) {
        // column sums of the staging tile: this thread owns 4 channels of TILE / NSG slots
#pragma unroll 4
        for (int j = 0; j < TILE / NSG; ++j) {
          const float4 m = *reinterpret_cast<const float4*>(stage + (sp * (TILE / NSG) + j) * LDS + cg * 4);
          const float2 m0 = make_float2(m.x, m.y), m1 = make_float2(m.z, m.w);
          st_s[0] = fadd2(st_s[0], m0);
          st_s[1] = fadd2(st_s[1], m1);
          st_ss[0] = ffma2(m0, m0, st_ss[0]);
          st_ss[1] = ffma2(m1, m1, st_ss[1]);
        }
      }
      if (FLAGS & F_AGG) {
        // Segmented mean by target row.  This tile finishes rows [R_prev, R_cur); row R_cur (if it has slots here)
        // continues in the next tile and goes to the carry.  TPR threads per row, RIF rows in flight; the bookkeeping of
        // each thread's first row was loaded during the previous tile.
        const int64_t s1 = s0 + nvalid;
        const bool last = s1 >= s_end;
        const float* cin = carry + (t & 1) * C;
        float* cout = carry + ((t + 1) & 1) * C;
        auto reduce_row = [&](int r, int b, int e, float di, const float4 (&bs)[CPT]) {
          if (r > R_cur || (r == R_cur && last)) return;
          const bool done = r < R_cur;
          if (!done && (int64_t)b >= s1) return;    // the next row starts exactly at the tile boundary: nothing to carry
          float2 acc[2 * CPT];
#pragma unroll
          for (int k = 0; k < 2 * CPT; ++k) acc[k] = make_float2(0.f, 0.f);
          if ((int64_t)b < s0) {                    // the row started in an earlier tile
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
              const float4 m = *reinterpret_cast<const float4*>(cin + ch + 4 * TPR * k);
              acc[2 * k] = make_float2(m.x, m.y);
              acc[2 * k + 1] = make_float2(m.z, m.w);
            }
          }
          const int lo = (int)(max((int64_t)b, s0) - s0), hi = (int)(min((int64_t)e, s1) - s0);
          for (int j = lo; j < hi; ++j) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
              const float4 m = *reinterpret_cast<const float4*>(stage + j * LDS + ch + 4 * TPR * k);
              acc[2 * k] = fadd2(acc[2 * k], make_float2(m.x, m.y));
              acc[2 * k + 1] = fadd2(acc[2 * k + 1], make_float2(m.z, m.w));
            }
          }
          if (done) {
            float* o = p.out + (int64_t)r * p.ldo + ch;
#pragma unroll
            for (int k = 0; k < CPT; ++k)
              *reinterpret_cast<float4*>(o + 4 * TPR * k) =
                  make_float4(fmaf(acc[2 * k].x, di, bs[k].x), fmaf(acc[2 * k].y, di, bs[k].y),
                              fmaf(acc[2 * k + 1].x, di, bs[k].z), fmaf(acc[2 * k + 1].y, di, bs[k].w));
          } else {
#pragma unroll
            for (int k = 0; k < CPT; ++k)
              *reinterpret_cast<float4*>(cout + ch + 4 * TPR * k) =
                  make_float4(acc[2 * k].x, acc[2 * k].y, acc[2 * k + 1].x, acc[2 * k + 1].y);
          }
        };
#pragma unroll
        for (int q2 = 0; q2 < NPF; ++q2) reduce_row(R_prev + rl + q2 * RIF, cb[q2], ce[q2], cdi[q2], cbs[q2]);
        for (int r = R_prev + rl + NPF * RIF; r <= R_cur; r += RIF) {        // tiles of many short rows (rare)
          int b, e;
          float di;
          float4 bs[CPT];
          row_meta(r, b, e, di, bs, true);
          reduce_row(r, b, e, di, bs);
        }
        R_prev = R_cur;
        R_cur = R_next;
      }
      if (tid == 0) EF_TRACE(t, 5);
    }
    if ((FLAGS & F_AGG) && ntiles == 0) {           // a range of rows without a single slot: out = base
      for (int r = r_begin + rl; r < r_end; r += RIF) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.base) b0 = __ldg(reinterpret_cast<const float4*>(p.base + (int64_t)r * p.ldb + ch + 4 * TPR * k));
          *reinterpret_cast<float4*>(p.out + (int64_t)r * p.ldo + ch + 4 * TPR * k) = b0;
        }
      }
    }
    if (FLAGS & F_STATS) {
      // combine the NSG slot groups through the staging tile: red[2][NSG][C]
      float* red = stage;
      named_bar_sync(1, E_THREADS);
      *reinterpret_cast<float4*>(red + (0 * NSG + sp) * C + cg * 4) = make_float4(st_s[0].x, st_s[0].y, st_s[1].x, st_s[1].y);
      *reinterpret_cast<float4*>(red + (1 * NSG + sp) * C + cg * 4) = make_float4(st_ss[0].x, st_ss[0].y, st_ss[1].x, st_ss[1].y);
      named_bar_sync(1, E_THREADS);
      if (et < C) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int k = 0; k < NSG; ++k) { s += red[(0 * NSG + k) * C + et]; ss += red[(1 * NSG + k) * C + et]; }
        const float b = p.b2 ? __ldg(p.b2 + et) : 0.f;
        const float cnt = (float)(s_end - s_begin);
        // sum (acc + b) and sum (acc + b)^2 from the raw accumulator sums
        p.part[((int64_t)blockIdx.x * 2 + 0) * C + et] = fmaf(cnt, b, s);
        p.part[((int64_t)blockIdx.x * 2 + 1) * C + et] = fmaf(b, fmaf(cnt, b, 2.f * s), ss);
      }
    }
    }   // !F_Z1
  } else if (CTRL && warp >= E_WARPS + G_WARPS) {
    // =========================== control warp group: TMA producer + MMA issuer (first warp only) =========
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == E_WARPS + G_WARPS) {
      auto fillc = [&](int j) {
        if (j >= ntiles) return;
        const int st = j % RING;
        if (j >= RING) mbar_wait_park(smem_u32(&bar_ring_empty[st]), (uint32_t)(((j / RING) - 1) & 1));
        const int64_t s0 = s_begin + (int64_t)j * TILE;
        const uint32_t bytes = (uint32_t)min((int64_t)TILE, s_end - s0) * 16u;
        const uint32_t bar = smem_u32(&bar_ring_full[st]);
        const uint32_t dst = smem_u32(ring + st * RING_BYTES);
        mbar_expect_tx(bar, 2u * bytes);
        bulk_g2s(dst, p.rec_idx + s0, bytes, bar);
        bulk_g2s(dst + TILE * 16, p.rec_attr + s0, bytes, bar);
      };
      if (lane == 0) {
        for (int j = 0; j < PF; ++j) fillc(j);
      }
      for (int t = 0; t < ntiles; ++t) {
        if (lane == 0) fillc(t + PF);
        __syncwarp();
        if (!(FLAGS & F_Z1)) issue_mma(t);
      }
    }
  } else {
    // =========================== gather warps ============================================================
    if (CTRL) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    // Thread (gc, sl): channels 4gc .. 4gc+3 of slots i * NSLG + sl (i < SPT) of every tile -- the two half-warps work
    // on adjacent slots, so their record reads share a wavefront and their a1 rows share an 8-row swizzle atom.  The
    // P / Q rows of tile t+1 are requested slot by slot while tile t is being computed ("rolling" register prefetch:
    // SPT slots per thread are always in flight, across tile boundaries).  Straight-line code, packed fp32 pairs.
    const int g = tid - E_THREADS, gc = g & 15, sl = g >> 4;
    float2 w[4][2], bias1[2], sc1[2], sh1[2];
    {
      float wt[4][4], bt[4], sct[4], sht[4];
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const int c = gc * 4 + qq;
#pragma unroll
        for (int k = 0; k < 4; ++k) wt[qq][k] = __ldg(p.w1c + c * p.ld1 + k);
        bt[qq] = p.b1 ? __ldg(p.b1 + c) : 0.f;
        sct[qq] = (FLAGS & F_Z1) ? 0.f : __ldg(p.stat1 + c);
        sht[qq] = (FLAGS & F_Z1) ? 0.f : __ldg(p.stat1 + C + c);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        w[k][0] = make_float2(wt[0][k], wt[1][k]);
        w[k][1] = make_float2(wt[2][k], wt[3][k]);
      }
      bias1[0] = make_float2(bt[0], bt[1]); bias1[1] = make_float2(bt[2], bt[3]);
      sc1[0] = make_float2(sct[0], sct[1]); sc1[1] = make_float2(sct[2], sct[3]);
      sh1[0] = make_float2(sht[0], sht[1]); sh1[1] = make_float2(sht[2], sht[3]);
    }
    float2 z_s[2], z_ss[2];                                 // F_Z1 accumulators
    z_s[0] = z_s[1] = z_ss[0] = z_ss[1] = make_float2(0.f, 0.f);
    const char* pbase = reinterpret_cast<const char*>(p.pq + gc * 4);
    float4 qv[SPT], pv[SPT];
    auto issue = [&](int st, int i) {
      const int2 ds = *reinterpret_cast<const int2*>(ring + st * RING_BYTES + (i * NSLG + sl) * 16);
      // (re-loading P[dst] for every slot of a run was measured faster than de-duplicating it with a select chain)
      pv[i] = __ldg(reinterpret_cast<const float4*>(pbase + (uint32_t)ds.x));
      qv[i] = __ldg(reinterpret_cast<const float4*>(pbase + (uint32_t)ds.y));
    };
    const int gw = g >> 5;                                  // gather warp index
    if (ntiles > 0) {
      mbar_wait_park(smem_u32(&bar_ring_full[0]), 0u);
#pragma unroll
      for (int i = 0; i < SPT; ++i) issue(0, i);
    }
    // byte offset of (row sl, chunk gc) in the SWIZZLE_128B K-major a1 tile: row r = i * NSLG + sl lives in the 8-row
    // atom r >> 3 at row r & 7 (= sl & 7, since NSLG is a multiple of 8), its 16-byte chunk index is flipped by r & 7
    const uint32_t off0 = (uint32_t)(gc >> 3) * A_KB + (uint32_t)(sl >> 3) * 1024u + (uint32_t)(sl & 7) * 128u +
                          ((((uint32_t)gc & 7u) ^ ((uint32_t)sl & 7u)) << 4);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it & 1;
      const int64_t s0 = s_begin + (int64_t)it * TILE;
      const int nvalid = (int)min((int64_t)TILE, s_end - s0);
      const int st_cur = it % RING, st_next = (it + 1) % RING;
      const bool have_next = it + 1 < ntiles;
      if (g == 0) EF_TRACE(it, 6);
      if (have_next) mbar_wait_park(smem_u32(&bar_ring_full[st_next]), (uint32_t)(((it + 1) / RING) & 1));
      if (g == 0) EF_TRACE(it, 7);
      if (!(FLAGS & F_Z1) && it >= 2) mbar_wait_park(smem_u32(&bar_a_empty[s]), (uint32_t)(((it >> 1) - 1) & 1));
      if (g == 0) EF_TRACE(it, 8);
      uint8_t* a_tile = sm + (uint32_t)s * A_STAGE;
      const float4* attr_s = reinterpret_cast<const float4*>(ring + st_cur * RING_BYTES + TILE * 16);
      // MMA duty (MMA_G): gather warp (it - 1) % G_WARPS issues the product of tile it - 1 while it gathers tile it, one
      // k-step after each of its slots.  Every warp arrived for tile it - 1 before starting this one, so the waits below
      // are short; the duty rotates, so no warp falls behind the others.
      const bool duty = MMA_G && !(FLAGS & F_Z1) && it >= 1 && gw == (it - 1) % G_WARPS;
      if (duty) {
        const int tp = it - 1;
        mbar_wait_park(smem_u32(&bar_a_full[tp & 1]), (uint32_t)((tp >> 1) & 1));
        if (tp >= 2) mbar_wait_park(smem_u32(&bar_acc_empty[tp & 1]), (uint32_t)(((tp >> 1) - 1) & 1));
        tc_fence_after();
      }
      if (it + PFL2 < ntiles) {
        // L2 prefetch of the P / Q rows of tile it + PFL2 (its records are already in the ring): the register loads
        // above run only one tile ahead of their use, which hides an L2 hit but not a DRAM miss.  Four 128-byte lines
        // per slot, one per lane gc < 4 (P line 0 / 1, Q line 0 / 1).
        const int st_pf = (it + PFL2) % RING;
        mbar_wait_park(smem_u32(&bar_ring_full[st_pf]), (uint32_t)(((it + PFL2) / RING) & 1));
        if (gc < 4) {
          const char* pq_bytes = reinterpret_cast<const char*>(p.pq);
#pragma unroll
          for (int i = 0; i < SPT; ++i) {
            const int2 ds = *reinterpret_cast<const int2*>(ring + st_pf * RING_BYTES + (i * NSLG + sl) * 16);
            const uint32_t off = (uint32_t)((gc & 2) ? ds.y : ds.x) + (uint32_t)(gc & 1) * 128u;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pq_bytes + off));
          }
        }
      }
#pragma unroll
      for (int i = 0; i < SPT; ++i) {
        const int slot = i * NSLG + sl;
        const float4 at = attr_s[slot];
        // same association in every pass: ((W1c attr + b1) + P) + Q
        float2 v0 = ffma2s(at.x, w[0][0], bias1[0]), v1 = ffma2s(at.x, w[0][1], bias1[1]);
        v0 = ffma2s(at.y, w[1][0], v0); v1 = ffma2s(at.y, w[1][1], v1);
        v0 = ffma2s(at.z, w[2][0], v0); v1 = ffma2s(at.z, w[2][1], v1);
        v0 = ffma2s(at.w, w[3][0], v0); v1 = ffma2s(at.w, w[3][1], v1);
        v0 = fadd2(fadd2(v0, make_float2(pv[i].x, pv[i].y)), make_float2(qv[i].x, qv[i].y));
        v1 = fadd2(fadd2(v1, make_float2(pv[i].z, pv[i].w)), make_float2(qv[i].z, qv[i].w));
        if (have_next) issue(st_next, i);                 // refill this register slot for the next tile
        if (FLAGS & F_Z1) {
          if (slot < nvalid) {                            // padding slots carry stale rows
            z_s[0] = fadd2(z_s[0], v0); z_s[1] = fadd2(z_s[1], v1);
            z_ss[0] = ffma2(v0, v0, z_ss[0]); z_ss[1] = ffma2(v1, v1, z_ss[1]);
          }
          continue;
        }
        if ((FLAGS & F_TAPE) && slot < nvalid)
          *reinterpret_cast<float4*>(p.z1 + (s0 + slot) * C + gc * 4) = make_float4(v0.x, v0.y, v1.x, v1.y);
        float2 a0 = ffma2(v0, sc1[0], sh1[0]), a1 = ffma2(v1, sc1[1], sh1[1]);
        a0.x = fmaxf(a0.x, 0.f); a0.y = fmaxf(a0.y, 0.f); a1.x = fmaxf(a1.x, 0.f); a1.y = fmaxf(a1.y, 0.f);
        store_split_trunc(a_tile, a_tile + A_HI, off0 + (uint32_t)(i * (NSLG / 8)) * 1024u, a0, a1);   // a1 >= 0, finite
        if (duty) {
#pragma unroll
          for (int k = i * (8 / SPT); k < (i + 1) * (8 / SPT); ++k) issue_kstep(it - 1, k);
        }
      }
      if (FLAGS & F_Z1) {
        warp_arrive(smem_u32(&bar_ring_empty[st_cur]), lane);
        continue;
      }
      fence_proxy_async_smem();                     // generic-proxy smem writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar_a_full[s]));
        mbar_arrive(smem_u32(&bar_ring_empty[st_cur]));   // idx (read one tile ago) and attr of this tile are consumed
      }
      if (g == 0) EF_TRACE(it, 9);
    }
    if (MMA_G && !(FLAGS & F_Z1) && ntiles > 0 && gw == (ntiles - 1) % G_WARPS) {     // the last tile: nothing left to overlap
      __syncwarp();
      issue_mma(ntiles - 1);
    }
    if (FLAGS & F_Z1) {                                   // combine the slot groups; part = [sum | sum of squares]
      float* red = reinterpret_cast<float*>(sm);          // [2][NSLG][C], the a1 ring is unused in this pass
      *reinterpret_cast<float4*>(red + (0 * NSLG + sl) * C + gc * 4) = make_float4(z_s[0].x, z_s[0].y, z_s[1].x, z_s[1].y);
      *reinterpret_cast<float4*>(red + (1 * NSLG + sl) * C + gc * 4) = make_float4(z_ss[0].x, z_ss[0].y, z_ss[1].x, z_ss[1].y);
      named_bar_sync(2, G_THREADS);
      if (g < 2 * C) {
        const int which = g / C, c = g % C;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < NSLG; ++k) acc += red[(which * NSLG + k) * C + c];
        p.part[((int64_t)blockIdx.x * 2 + which) * C + c] = acc;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 2 * C);
}

// Slot-ordered records of one layer call: P / Q byte offsets (Q relative to the P base, i.e. + C floats) | eid, and the
// attribute row of the slot's edge.  16 bytes each so that any slot range is a legal TMA bulk copy.
__global__ void k_edge_records(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                               const int32_t* __restrict__ eid, const float4* __restrict__ attr, int64_t E,
                               uint32_t ldpq_b, int4* __restrict__ rec_idx, float4* __restrict__ rec_attr) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= E) return;
  const int e = __ldg(eid + s);
  rec_idx[s] = make_int4((int)((uint32_t)__ldg(dst + s) * ldpq_b), (int)((uint32_t)__ldg(src + s) * ldpq_b + (uint32_t)C * 4u), e, 0);
  rec_attr[s] = __ldg(attr + e);
}

// Role split: YOLAT_EF_ROLES = "8x8" (default) or "4x16" epilogue x gather warps (a tuning knob, all CUDA).
static int roles() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YOLAT_EF_ROLES");
    v = 1;
    if (e && e[0] == '4') v = 0;
  }
  return v;
}

template <int FLAGS, int EW, int GW, int MG>
static cudaError_t launch_cfg(const Params& p, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_fused<FLAGS, EW, GW, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_edge_fused<FLAGS, EW, GW, MG><<<grid, (EW + GW + (MG == 2 ? 4 : 0)) * 32, SMEM_BYTES, st>>>(p);
  return cudaSuccess;
}

// Who issues the MMAs (MMA_MODE of the kernel).  Measured at N = 320 000, E = 1 280 000 (8x8 warps):
//   F_STATS  144 us (0: epilogue warp 0)   155 us (1: gather warps, piece-wise)   127 us (2: control warp)
//   F_AGG    237 us                        180 us                                 169 us
// Pass A has no MMAs and runs in mode 0.  YOLAT_EF_MG = 0 / 1 / 2 forces one mode for the two MMA passes.
static int mma_by_gather(int flags) {
  static int v = -2;
  if (v == -2) { const char* e = getenv("YOLAT_EF_MG"); v = e ? atoi(e) : -1; }
  if (flags & F_Z1) return 0;
  return v >= 0 ? v : 2;
}

template <int FLAGS>
static cudaError_t launch(const Params& p, int grid, cudaStream_t st) {
  const int mg = mma_by_gather(FLAGS);
  switch (roles()) {
    case 1:
      return mg == 2 ? launch_cfg<FLAGS, 8, 8, 2>(p, grid, st)
                     : (mg == 1 ? launch_cfg<FLAGS, 8, 8, 1>(p, grid, st) : launch_cfg<FLAGS, 8, 8, 0>(p, grid, st));
    default: return mg ? launch_cfg<FLAGS, 4, 16, 1>(p, grid, st) : launch_cfg<FLAGS, 4, 16, 0>(p, grid, st);
  }
}

}  // namespace ef

bool edge_fused_supported(int C) { return C == ef::C; }
// the kernels index P / Q rows with 32-bit element offsets
bool edge_fused_fits(int64_t N, int64_t ldpq) { return N * ldpq < (1ll << 30); }   // 32-bit byte offsets

int edge_fused_grid(int64_t E) {
  const int64_t tiles = cdiv(E > 0 ? E : 1, ef::TILE);
  const int64_t cap = (int64_t)kNumSMs;           // persistent: one CTA (~220 KB of shared memory) per SM
  return (int)(tiles < cap ? tiles : cap);
}

int edge_stats1_grid(int64_t E) { return edge_fused_grid(E); }

// floats of workspace for the slot-ordered records of one layer call (8 per slot)
int64_t edge_records_floats(int64_t E) { return 8 * (E > 0 ? E : 0) + 8; }

// rec: [edge_records_floats(E)] floats, 16-byte aligned (the arena is): int4 records first, float4 attribute rows after
int edge_records(const GraphView& g, int64_t E, const float* attr, int64_t ldpq, float* rec, cudaStream_t st) {
  if (E <= 0) return YOLAT_OK;
  int4* ri = reinterpret_cast<int4*>(rec);
  float4* ra = reinterpret_cast<float4*>(rec + 4 * E);
  ef::k_edge_records<<<(unsigned)cdiv(E, 256), 256, 0, st>>>(g.src_t, g.dst_t, g.eid_t, reinterpret_cast<const float4*>(attr), E,
                                                             (uint32_t)ldpq * 4u, ri, ra);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

static void set_records(ef::Params& p, const float* rec, int64_t E) {
  p.rec_idx = reinterpret_cast<const int4*>(rec);
  p.rec_attr = reinterpret_cast<const float4*>(rec + 4 * E);
}

// part: [edge_stats1_grid(E)][2][C] sums / sums of squares of z1 over all edges
int edge_stats1(const GraphView& g, int64_t N, int64_t E, const float* pq, int64_t ldpq, const float* rec,
                const float* w1, int Cin, const float* b1, float* part, cudaStream_t st) {
  if (E <= 0) return YOLAT_OK;
  ef::Params p{};
  p.rowptr = g.rowptr_t; p.src = g.src_t; p.dst = g.dst_t; p.eid = g.eid_t; p.deg_inv = g.deg_inv;
  p.N = N; p.E = E; p.pq = pq; p.ldpq = (uint32_t)ldpq; set_records(p, rec, E);
  p.w1c = w1 + 2 * Cin; p.ld1 = 2 * Cin + 4; p.b1 = b1; p.part = part;
  ProfScope prof(YOLAT_PROF_EDGE_STATS1, st);
  cudaError_t e = ef::launch<ef::F_Z1>(p, edge_fused_grid(E), st);
  if (e != cudaSuccess) { set_last_error(e); return YOLAT_ERR_LAUNCH; }
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

// flags: EF_TAPE | EF_STATS | EF_AGG (common.cuh).  part: [edge_fused_grid(E)][2][C] when EF_STATS.
int edge_fused(const GraphView& g, int64_t N, int64_t E, int flags, const float* pq, int64_t ldpq, const float* rec,
               const float* w1, int Cin, const float* b1, const float* stat1, const float* w2, const float* b2,
               const float* stat2, const float* ew, float* z1, float* z2, float* part, const float* base, int64_t ldb,
               float* out, int64_t ldo, cudaStream_t st) {
  if (E <= 0) return YOLAT_OK;
  ef::Params p{};
  p.rowptr = g.rowptr_t; p.src = g.src_t; p.dst = g.dst_t; p.eid = g.eid_t; p.deg_inv = g.deg_inv;
  p.N = N; p.E = E; p.pq = pq; p.ldpq = (uint32_t)ldpq; set_records(p, rec, E);
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

#ifdef YOLAT_EF_TRACE
extern "C" int yolat_debug_ef_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, yolat::ef::g_ef_trace, sizeof(long long) * 128 * 16);
}
#endif
