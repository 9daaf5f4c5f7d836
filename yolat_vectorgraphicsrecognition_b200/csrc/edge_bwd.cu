// edge_bwd.cu -- K-EDGE backward: the tape-free backward of the fused gather -> edge-MLP -> scatter-mean path of
// GraphConv('attr_edge_gp2') (autograd of gcn_lib/sparse/torch_vertex.py:330-337 + PyG propagate + scatter-mean).
//
// The reference's autograd keeps z1, bn1, a1, z2, bn2, m (six [E, C] tensors) alive and walks them backwards.  Here
// nothing of size [E, C] exists in HBM in either direction: every pass RECOMPUTES z1 -> a1 -> z2 per 128-slot tile from
// the node-level P | Q rows (P = x (W1a - W1b)^T, Q = x W1b^T), exactly like the forward passes of edge_fused.cu, and
// keeps the per-edge gradients on chip.  With g = dL/dout, G_e = g[dst_e] / max(deg_in, 1) * w_e:
//
//   D1   (CSR-by-target order)   z2 -> mask2, xhat2;  BN2 backward statistics  sum G mask2, sum G mask2 xhat2
//                                (segment sums per target row x the row's g: g is row-constant in this order);
//                                also sum xhat1 (x) attr and sum attr (data-only terms of dW1c).
//   D2T  (CSR-by-target order)   dz2 = sc2 (G mask2 - m1 - xhat2 m2)            [epilogue 1, -> shared memory, bf16 hi/lo]
//                                da1 = dz2 W2                                     [tcgen05, B = W2 read MN-major]
//                                dW2 += dz2^T a1                                  [tcgen05, both operands MN-major, TMEM
//                                                                                  accumulator lives across the CTA's tiles]
//                                dy1 = da1 mask1                                  [epilogue 2]
//                                U_t[r] = sum_{e -> r} dy1_e,  X_t[r] = sum_{e -> r} xhat1_e   (segmented, no atomics)
//                                sum dy1, sum dy1 xhat1 (BN1 backward statistics), T = sum dy1 (x) attr.
//   D2S  (CSR-by-source order)   the same recompute over the source-sorted slot order: U_s[v], X_s[v] (no atomics: the
//                                scatter to the source endpoint is a segmented sum in this order).
//
// BN1's backward is linear in its statistics, so no third pass is needed: with (m1, m2) = BN1 backward means,
//   dP[r] = sc1 (U_t[r] - deg_in(r) m1 - m2 X_t[r]),   dQ[v] = sc1 (U_s[v] - deg_out(v) m1 - m2 X_s[v]),
//   dW1c  = sc1 (T - m1 (x) sum attr - m2 sum xhat1 (x) attr)                     (k_edge_bwd_combine, node level).
//
// Tensor-core products are split-bf16: every fp32 operand is split into bf16 terms (round to nearest).  The gradient
// products (da1, dW2) use two terms, hi*hi + hi*lo + lo*hi (~2^-16 per product, unbiased); the recomputed z2 -- whose
// sign decides the ReLU masks -- uses three terms and six products (~2^-24, the accuracy of the forward's 3xTF32).  16-bit operands use the same
// SWIZZLE_128B shared-memory layout for K-major and MN-major reads (tools/umma_probe.cu), so the a1 / dz2 / W2 tiles are
// written once and read by all three products; dW2 stacks [dz2_hi ; dz2_lo] as M = 128 through the descriptor's LBO.
//
// One persistent CTA (512 threads) per SM owns a row-aligned slot range and walks it tile by tile in bulk-synchronous
// phases (gather | MMA | epilogue 1 | MMA | epilogue 2 | row sweep); records arrive through a 3-stage TMA bulk-copy ring.
#include <cstdlib>
#include "common.cuh"
#include "tc.cuh"

namespace yolat {
namespace eb {

using namespace tc;

constexpr int C = 64;
constexpr int RING = 3;
constexpr uint32_t W_BF = 64 * 128;                   // one bf16 part of W2 [64 out][64 in]: 8 KB
// Tile geometry.  TILE = 128: one CTA of 512 threads per SM.  TILE = 64: 256 threads, TWO CTAs per SM -- the phases of a
// tile are bulk-synchronous (every phase starts with loads nothing overlaps, both MMA waits idle the whole CTA), so a
// second resident CTA is what fills those gaps; the M = 64 accumulators of z2 / da1 then keep rows 16 q .. 16 q + 15 in
// TMEM lanes 32 q .. 32 q + 15 (tools/umma_probe_m64.cu): half of each epilogue warp idles in the two epilogue phases.
template <int TILE>
struct L {
  static constexpr int THREADS = TILE * 4;
  static constexpr int SLG = THREADS / 16;             // slot groups of the gather: slot = sl + SLG * i, i < 4
  static constexpr int NCOL = TILE == 128 ? 16 : 32;   // accumulator columns per epilogue thread
  static constexpr uint32_t T_BF = TILE * 128;         // one bf16 tile [TILE slots][64 channels]
  static constexpr uint32_t OFF_A1 = 0;                // a1 hi | mid | lo
  static constexpr uint32_t OFF_DZ = OFF_A1 + 3 * T_BF;   // dz2 hi | lo
  static constexpr uint32_t OFF_W2 = OFF_DZ + 2 * T_BF;   // W2 hi | mid | lo
  static constexpr uint32_t OFF_ZT = OFF_W2 + 3 * W_BF;   // z1 tile, fp32 [TILE][64], chunk-swizzled
  static constexpr uint32_t OFF_ST = OFF_ZT + TILE * 256; // G -> dy1 (D2) | xhat2 with the mask in the lowest mantissa bit (D1)
  static constexpr uint32_t OFF_REC = OFF_ST + TILE * 256;   // record ring: RING x (TILE x int4 | TILE x float4)
  static constexpr uint32_t REC_BYTES = TILE * 32;
  static constexpr uint32_t OFF_TAB = OFF_REC + RING * REC_BYTES;   // per-channel tables, 16 x 64 floats
  static constexpr uint32_t OFF_EW = OFF_TAB + 16 * 256;  // edge weight per slot (D1)
  static constexpr uint32_t OFF_CARRY = OFF_EW + TILE * 4;   // [2 parities][U | X][64]
  static constexpr uint32_t SMEM_BYTES = OFF_CARRY + 2 * 2 * 256 + 1024;
  static_assert(SMEM_BYTES * (TILE == 64 ? 2 : 1) <= 226 * 1024, "shared memory budget");
  static_assert(OFF_REC >= (THREADS / 8) * 8 * 48 * 4 && OFF_REC >= (SLG * 16 * 20) * 4, "the final reduction aliases the tiles");
};

enum { M_D1 = 0, M_D2T = 1, M_D2S = 2 };
// per-channel tables
enum { TB_SC1 = 0, TB_SH1, TB_IS1, TB_XM1, TB_SC2, TB_SH2B, TB_IS2, TB_XM2, TB_C0, TB_C1, TB_B1, TB_W1C /* 4 rows: k */ };

struct Params {
  const int32_t* rowptr;      // [N + 1] of this pass's slot order
  const int4* rec_idx;        // [E] (dst, src, eid, 0) in this pass's slot order
  const float4* rec_attr;     // [E]
  const float* deg_inv;       // [N] 1 / max(in-degree, 1)
  int64_t N, E;
  const float* pq; uint32_t ldpq;
  const float* w1c; int ld1;
  const float* b1;
  const float* stat1;         // [4C] sc | sh | mean | invstd of BN1
  const float* w2; const float* b2;
  const float* stat2;         // [4C] of BN2
  const float* bstat2;        // [2C] BN2 backward means m1 | m2 (D2)
  const float* ew;            // [E] by edge id, or null
  const float* g; int64_t ldg;
  float* U; float* X;         // [N, C] (D2)
  float* part;                // [grid][2][C]
  float* part_t;              // D2S: [grid][C * 4 + 4]  (sum xhat1 (x) attr | sum attr);  D2T: [grid][C * 4]
  float* part_w2;             // D2T: [grid][128][64]
};

__device__ __forceinline__ uint32_t off_bf(int slot, int c) {     // c: channel, multiple of 4
  return (uint32_t)(slot >> 3) * 1024u + (uint32_t)(slot & 7) * 128u + ((((uint32_t)c >> 3) ^ ((uint32_t)slot & 7u)) << 4) +
         ((uint32_t)c & 7u) * 2u;
}
__device__ __forceinline__ uint32_t off_f(int slot, int chunk) {  // chunk: 16-byte chunk (4 channels) of the fp32 row
  return (uint32_t)slot * 256u + ((((uint32_t)chunk) ^ ((uint32_t)slot & 15u)) << 4);
}

__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }

template <int MODE>
__device__ __forceinline__ int slot_row(const int4& r) { return MODE == M_D2S ? r.y : r.x; }

// first row whose slots start at or after slot s
template <int MODE>
__device__ __forceinline__ int row_at_or_after(const Params& p, int64_t s) {
  if (s <= 0) return 0;
  if (s >= p.E) return (int)p.N;
  const int4 r = __ldg(p.rec_idx + s);
  const int v = slot_row<MODE>(r);
  return (p.rowptr[v] == (int32_t)s) ? v : v + 1;
}

template <int MODE, int TILE>
__global__ void __launch_bounds__(L<TILE>::THREADS, TILE == 64 ? 2 : 1) k_edge_bwd(const Params p) {
  using LT = L<TILE>;
  constexpr int THREADS = LT::THREADS, SLG = LT::SLG, NCOL = LT::NCOL;
  constexpr uint32_t T_BF = LT::T_BF, REC_BYTES = LT::REC_BYTES;
  constexpr bool D1 = MODE == M_D1, D2 = MODE != M_D1, D2T = MODE == M_D2T, TXA = MODE == M_D2S;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_ring[RING];
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_slot;

  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = ((raw_u32 + 1023u) & ~1023u) - raw_u32;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_u32 = raw_u32 + pad;
  uint8_t* a1_t = sm + LT::OFF_A1;
  uint8_t* dz_t = sm + LT::OFF_DZ;
  uint8_t* w2_t = sm + LT::OFF_W2;
  uint8_t* z_t = sm + LT::OFF_ZT;
  uint8_t* s_t = sm + LT::OFF_ST;
  uint8_t* ring = sm + LT::OFF_REC;
  float* tab = reinterpret_cast<float*>(sm + LT::OFF_TAB);
  float* ew_s = reinterpret_cast<float*>(sm + LT::OFF_EW);
  float* carry = reinterpret_cast<float*>(sm + LT::OFF_CARRY);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < RING; ++i) mbar_init(smem_u32(&bar_ring[i]), 1);
    mbar_init(smem_u32(&bar_mma), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 256);
  for (int idx = tid; idx < (int)(RING * REC_BYTES / 16); idx += THREADS) reinterpret_cast<int4*>(ring)[idx] = make_int4(0, 0, 0, 0);
  // W2 -> bf16 hi / lo, [out][in], SWIZZLE_128B rows of 64 in-channels
  for (int idx = tid; idx < C * 16; idx += THREADS) {
    const int n = idx >> 4, c4 = idx & 15;
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.w2 + n * C + c4 * 4));
    uint2 hi, mid, lo;
    split3_bf16x4(v, hi, mid, lo);
    const uint32_t off = off_bf(n, c4 * 4);
    *reinterpret_cast<uint2*>(w2_t + off) = hi;
    *reinterpret_cast<uint2*>(w2_t + W_BF + off) = mid;
    *reinterpret_cast<uint2*>(w2_t + 2 * W_BF + off) = lo;
  }
  if (tid < C) {
    const int c = tid;
    const float sc1 = __ldg(p.stat1 + c), sh1 = __ldg(p.stat1 + C + c), mu1 = __ldg(p.stat1 + 2 * C + c), is1 = __ldg(p.stat1 + 3 * C + c);
    const float sc2 = __ldg(p.stat2 + c), sh2 = __ldg(p.stat2 + C + c), mu2 = __ldg(p.stat2 + 2 * C + c), is2 = __ldg(p.stat2 + 3 * C + c);
    const float b2 = p.b2 ? __ldg(p.b2 + c) : 0.f;
    tab[TB_SC1 * C + c] = sc1; tab[TB_SH1 * C + c] = sh1; tab[TB_IS1 * C + c] = is1; tab[TB_XM1 * C + c] = -mu1 * is1;
    tab[TB_B1 * C + c] = p.b1 ? __ldg(p.b1 + c) : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) tab[(TB_W1C + k) * C + c] = __ldg(p.w1c + c * p.ld1 + k);
    tab[TB_SC2 * C + c] = sc2; tab[TB_SH2B * C + c] = fmaf(b2, sc2, sh2);
    const float xm2 = (b2 - mu2) * is2;
    tab[TB_IS2 * C + c] = is2; tab[TB_XM2 * C + c] = xm2;
    if (D2) {
      // dz2 = sc2 (G mask - m1 - xhat2 m2),  xhat2 = acc is2 + xm2   =>   dz2 = mask (sc2 G) - c0 - c1 acc
      const float m1 = __ldg(p.bstat2 + c), m2 = __ldg(p.bstat2 + C + c);
      tab[TB_C0 * C + c] = sc2 * fmaf(m2, xm2, m1);
      tab[TB_C1 * C + c] = sc2 * m2 * is2;
    }
  }
  if (tid < 2 * 2 * C) carry[tid] = 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t T_Z2 = tmem_d, T_DA = tmem_d + 64, T_DW = tmem_d + 128;

  // slot range of this CTA, snapped to row boundaries
  const int r_begin = row_at_or_after<MODE>(p, p.E * (int64_t)blockIdx.x / gridDim.x);
  const int r_end = row_at_or_after<MODE>(p, p.E * (int64_t)(blockIdx.x + 1) / gridDim.x);
  const int64_t s_begin = r_begin < p.N ? p.rowptr[r_begin] : p.E;
  const int64_t s_end = r_end < p.N ? p.rowptr[r_end] : p.E;
  const int ntiles = (int)((s_end - s_begin + TILE - 1) / TILE);

  auto fill = [&](int j) {      // thread 0: records of tile j -> ring stage j % RING
    if (j >= ntiles) return;
    const int st = j % RING;
    const int64_t s0 = s_begin + (int64_t)j * TILE;
    const uint32_t bytes = (uint32_t)min((int64_t)TILE, s_end - s0) * 16u;
    const uint32_t bar = smem_u32(&bar_ring[st]);
    const uint32_t dst = smem_u32(ring + st * REC_BYTES);
    mbar_expect_tx(bar, 2u * bytes);
    bulk_g2s(dst, p.rec_idx + s0, bytes, bar);
    bulk_g2s(dst + TILE * 16, p.rec_attr + s0, bytes, bar);
  };
  if (tid == 0) { fill(0); fill(1); }

  // ---- gather mapping: thread (gc, sl) = channels 4 gc .. 4 gc + 3 of slots sl + SLG i (constants come from the tables) ----
  const int gc = tid & 15, sl = tid >> 4;
  // ---- epilogue mapping: warp quarter eq <-> TMEM lanes 32 eq .., NCOL columns from NCOL * ecg.  M = 128 accumulators: thread =
  //      slot 32 eq + lane; M = 64: slot 16 eq + lane for lane < 16 (the other half of the warp has no row) ----
  const int eq = warp & 3, ecg = warp >> 2;
  const int eslot = TILE == 128 ? eq * 32 + lane : eq * 16 + (lane & 15);
  const bool eact = TILE == 128 || lane < 16;
  // ---- sweep mapping: 8 threads per row, thread = chunks k8 and k8 + 8 of the row (channels 4 k8 .. and 32 + 4 k8 ..) ----
  const int rl = tid >> 3, k8 = tid & 7;
  // persistent accumulators
  float acc_a[8], acc_b[8];            // D1: sum G S0 | sum G S1;  D2T: sum dy1 | sum dy1 xhat1
  float acc_t[D2T ? 32 : 1];           // D2T: T[channel i][k]
  float acc_tx[TXA ? 16 : 1], acc_sa[TXA ? 4 : 1];   // D2S (gather mapping): sum xhat1 (x) attr, sum attr
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc_a[i] = 0.f; acc_b[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < (D2T ? 32 : 1); ++i) acc_t[i] = 0.f;
#pragma unroll
  for (int i = 0; i < (TXA ? 16 : 1); ++i) acc_tx[i] = 0.f;
#pragma unroll
  for (int i = 0; i < (TXA ? 4 : 1); ++i) acc_sa[i] = 0.f;

  uint32_t mma_phase = 0;
  int row_lo = r_begin;
  const char* pq_b = reinterpret_cast<const char*>(p.pq);
  const uint32_t ldpq_b = p.ldpq * 4u;

  for (int t = 0; t < ntiles; ++t) {
    const int st = t % RING;
    const int64_t s0 = s_begin + (int64_t)t * TILE;
    const int nvalid = (int)min((int64_t)TILE, s_end - s0);
    const int64_t s1 = s0 + nvalid;
    if (tid == 0) fill(t + 2);       // stage (t + 2) % RING was last read in tile t - 1, before its closing barrier
    mbar_wait_bounded(smem_u32(&bar_ring[st]), (uint32_t)((t / RING) & 1));
    const int4* rec_i = reinterpret_cast<const int4*>(ring + st * REC_BYTES);
    const float4* rec_a = reinterpret_cast<const float4*>(ring + st * REC_BYTES + TILE * 16);
    // row bookkeeping of the sweep, requested now: the loads complete under the gather / MMA / epilogue phases
    const int r_last = slot_row<MODE>(rec_i[nvalid - 1]);
    const int pre_r = row_lo + rl;
    int pre_b = 0, pre_e = 0;
    if (pre_r < p.N) { pre_b = __ldg(p.rowptr + pre_r); pre_e = __ldg(p.rowptr + pre_r + 1); }
    const int next_ptr = __ldg(p.rowptr + r_last + 1);

    // ================= gather: z1, a1 (bf16 hi / lo), G ==================================================
    {
      const float4 b1q = *reinterpret_cast<const float4*>(tab + TB_B1 * C + gc * 4);
      const float4 sc1q = *reinterpret_cast<const float4*>(tab + TB_SC1 * C + gc * 4), sh1q = *reinterpret_cast<const float4*>(tab + TB_SH1 * C + gc * 4);
      const float4 w0q = *reinterpret_cast<const float4*>(tab + (TB_W1C + 0) * C + gc * 4), w1q = *reinterpret_cast<const float4*>(tab + (TB_W1C + 1) * C + gc * 4);
      const float4 w2q = *reinterpret_cast<const float4*>(tab + (TB_W1C + 2) * C + gc * 4), w3q = *reinterpret_cast<const float4*>(tab + (TB_W1C + 3) * C + gc * 4);
      const float2 bp[2] = {make_float2(b1q.x, b1q.y), make_float2(b1q.z, b1q.w)};
      const float2 scp[2] = {make_float2(sc1q.x, sc1q.y), make_float2(sc1q.z, sc1q.w)};
      const float2 shp[2] = {make_float2(sh1q.x, sh1q.y), make_float2(sh1q.z, sh1q.w)};
      const float2 wp[4][2] = {{make_float2(w0q.x, w0q.y), make_float2(w0q.z, w0q.w)}, {make_float2(w1q.x, w1q.y), make_float2(w1q.z, w1q.w)},
                               {make_float2(w2q.x, w2q.y), make_float2(w2q.z, w2q.w)}, {make_float2(w3q.x, w3q.y), make_float2(w3q.z, w3q.w)}};
      // store offsets of slot sl + SLG i (SLG is a multiple of 16): the swizzle terms depend on sl only
      const uint32_t fbase = (uint32_t)sl * 256u + ((((uint32_t)gc) ^ ((uint32_t)sl & 15u)) << 4);
      const uint32_t bbase = (uint32_t)(sl >> 3) * 1024u + (uint32_t)(sl & 7) * 128u + (((((uint32_t)gc * 4u) >> 3) ^ ((uint32_t)sl & 7u)) << 4) +
                             (((uint32_t)gc * 4u) & 7u) * 2u;
      int4 rc[4];
      float4 pv[4], qv[4], gv[4];
      float gs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = sl + SLG * i;
        rc[i] = rec_i[j];
        pv[i] = __ldg(reinterpret_cast<const float4*>(pq_b + (uint32_t)rc[i].x * ldpq_b + (uint32_t)gc * 16u));
        qv[i] = __ldg(reinterpret_cast<const float4*>(pq_b + (uint32_t)rc[i].y * ldpq_b + (uint32_t)(C * 4 + gc * 16)));
        if (D2) {
          gv[i] = __ldg(reinterpret_cast<const float4*>(p.g + (int64_t)rc[i].x * p.ldg + gc * 4));
          gs[i] = __ldg(p.deg_inv + rc[i].x);
          if (p.ew) gs[i] *= __ldg(p.ew + rc[i].z);
        }
      }
      if (t + 1 < ntiles && gc < (D2 ? 6 : 4)) {
        // L2 prefetch of the rows tile t + 1 will gather (its records landed a tile ago)
        const int stn = (t + 1) % RING;
        mbar_wait_bounded(smem_u32(&bar_ring[stn]), (uint32_t)(((t + 1) / RING) & 1));
        const int4* rn = reinterpret_cast<const int4*>(ring + stn * REC_BYTES);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int4 r = rn[sl + SLG * i];
          const char* a = gc < 2 ? pq_b + (uint32_t)r.x * ldpq_b + (uint32_t)gc * 128u
                        : gc < 4 ? pq_b + (uint32_t)r.y * ldpq_b + (uint32_t)(C * 4) + (uint32_t)(gc - 2) * 128u
                                 : reinterpret_cast<const char*>(p.g + (int64_t)r.x * p.ldg) + (gc - 4) * 128;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = sl + SLG * i;
        const bool valid = j < nvalid;
        const float4 at = rec_a[j];
        // same association as the forward passes: ((W1c attr + b1) + P) + Q, in packed fp32 pairs
        float2 v0 = ffma2s(at.x, wp[0][0], bp[0]), v1 = ffma2s(at.x, wp[0][1], bp[1]);
        v0 = ffma2s(at.y, wp[1][0], v0); v1 = ffma2s(at.y, wp[1][1], v1);
        v0 = ffma2s(at.z, wp[2][0], v0); v1 = ffma2s(at.z, wp[2][1], v1);
        v0 = ffma2s(at.w, wp[3][0], v0); v1 = ffma2s(at.w, wp[3][1], v1);
        v0 = fadd2(fadd2(v0, make_float2(pv[i].x, pv[i].y)), make_float2(qv[i].x, qv[i].y));
        v1 = fadd2(fadd2(v1, make_float2(pv[i].z, pv[i].w)), make_float2(qv[i].z, qv[i].w));
        const float z[4] = {v0.x, v0.y, v1.x, v1.y};
        float2 a0 = ffma2(v0, scp[0], shp[0]), a1 = ffma2(v1, scp[1], shp[1]);
        float4 a;
        a.x = valid ? fmaxf(a0.x, 0.f) : 0.f; a.y = valid ? fmaxf(a0.y, 0.f) : 0.f;
        a.z = valid ? fmaxf(a1.x, 0.f) : 0.f; a.w = valid ? fmaxf(a1.y, 0.f) : 0.f;
        *reinterpret_cast<float4*>(z_t + fbase + (uint32_t)i * (SLG * 256u)) = make_float4(z[0], z[1], z[2], z[3]);
        uint2 hi, mid, lo;
        split3_bf16x4(a, hi, mid, lo);
        const uint32_t ob = bbase + (uint32_t)i * (SLG / 8 * 1024u);
        *reinterpret_cast<uint2*>(a1_t + ob) = hi;
        *reinterpret_cast<uint2*>(a1_t + T_BF + ob) = mid;
        *reinterpret_cast<uint2*>(a1_t + 2 * T_BF + ob) = lo;
        if (D2) {
          const float s = valid ? gs[i] : 0.f;
          *reinterpret_cast<float4*>(s_t + fbase + (uint32_t)i * (SLG * 256u)) = make_float4(gv[i].x * s, gv[i].y * s, gv[i].z * s, gv[i].w * s);
        }
        if (D1 && gc == 0) ew_s[j] = (valid && p.ew) ? __ldg(p.ew + rc[i].z) : 1.f;
        if (TXA && valid) {
          const float atv[4] = {at.x, at.y, at.z, at.w};
          const float4 is1q = *reinterpret_cast<const float4*>(tab + TB_IS1 * C + gc * 4), xm1q = *reinterpret_cast<const float4*>(tab + TB_XM1 * C + gc * 4);
          const float is1v[4] = {is1q.x, is1q.y, is1q.z, is1q.w}, xm1v[4] = {xm1q.x, xm1q.y, xm1q.z, xm1q.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float xh = fmaf(z[q], is1v[q], xm1v[q]);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc_tx[TXA ? q * 4 + k : 0] = fmaf(xh, atv[k], acc_tx[TXA ? q * 4 + k : 0]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) acc_sa[TXA ? k : 0] += atv[k];
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    // ================= MMA 1: z2 = a1 W2^T ==============================================================
    if (warp == 0) {
      tc_fence_after();
      if (elect_one_sync()) {
        constexpr uint32_t IDESC = make_idesc_bf16(TILE, C, 0, 0);
        const uint32_t a_u32 = sm_u32 + LT::OFF_A1, w_u32 = sm_u32 + LT::OFF_W2;
        // z2 decides the ReLU masks of the backward: three-term splits, six products (everything above 2^-24 |a w|), so
        // the recomputed masks flip no more often than an fp32 product's would
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t a_h = make_desc(a_u32 + ks * 32u, 16, 1024, LAYOUT_SW128);
          const uint64_t a_m = make_desc(a_u32 + T_BF + ks * 32u, 16, 1024, LAYOUT_SW128);
          const uint64_t a_l = make_desc(a_u32 + 2 * T_BF + ks * 32u, 16, 1024, LAYOUT_SW128);
          const uint64_t b_h = make_desc(w_u32 + ks * 32u, 16, 1024, LAYOUT_SW128);
          const uint64_t b_m = make_desc(w_u32 + W_BF + ks * 32u, 16, 1024, LAYOUT_SW128);
          const uint64_t b_l = make_desc(w_u32 + 2 * W_BF + ks * 32u, 16, 1024, LAYOUT_SW128);
          umma_bf16(T_Z2, a_l, b_h, IDESC, ks > 0 ? 1u : 0u);     // small terms first
          umma_bf16(T_Z2, a_h, b_l, IDESC, 1u);
          umma_bf16(T_Z2, a_m, b_m, IDESC, 1u);
          umma_bf16(T_Z2, a_m, b_h, IDESC, 1u);
          umma_bf16(T_Z2, a_h, b_m, IDESC, 1u);
          umma_bf16(T_Z2, a_h, b_h, IDESC, 1u);
        }
        umma_commit(smem_u32(&bar_mma));
      }
      __syncwarp();
    }
    mbar_wait_bounded(smem_u32(&bar_mma), mma_phase);
    mma_phase ^= 1u;
    tc_fence_after();

    // ================= epilogue 1 =======================================================================
    {
      float v[NCOL];
      tmem_ldn(T_Z2 + ((uint32_t)(eq * 32) << 16) + (uint32_t)(ecg * NCOL), v);
      const bool valid = eslot < nvalid;
      if (!eact) {
      } else if (D1) {
        // xhat2 with the ReLU mask in the lowest mantissa bit (a 1-ulp perturbation of a statistic's summand)
#pragma unroll
        for (int k = 0; k < NCOL / 4; ++k) {
          const int c = ecg * NCOL + k * 4;
          const float4 is2 = *reinterpret_cast<const float4*>(tab + TB_IS2 * C + c), xm2 = *reinterpret_cast<const float4*>(tab + TB_XM2 * C + c);
          const float4 sc2 = *reinterpret_cast<const float4*>(tab + TB_SC2 * C + c), sh2 = *reinterpret_cast<const float4*>(tab + TB_SH2B * C + c);
          const float isv[4] = {is2.x, is2.y, is2.z, is2.w}, xmv[4] = {xm2.x, xm2.y, xm2.z, xm2.w};
          const float scv[4] = {sc2.x, sc2.y, sc2.z, sc2.w}, shv[4] = {sh2.x, sh2.y, sh2.z, sh2.w};
          float o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float a = v[k * 4 + q];
            const float xh = fmaf(a, isv[q], xmv[q]);
            const uint32_t m = fmaf(a, scv[q], shv[q]) > 0.f ? 1u : 0u;
            o[q] = __uint_as_float((__float_as_uint(xh) & ~1u) | m);
          }
          *reinterpret_cast<float4*>(s_t + off_f(eslot, ecg * (NCOL / 4) + k)) = make_float4(o[0], o[1], o[2], o[3]);
        }
      } else {
        float dz[NCOL];
#pragma unroll
        for (int k = 0; k < NCOL / 4; ++k) {
          const int c = ecg * NCOL + k * 4;
          const float4 sc2 = *reinterpret_cast<const float4*>(tab + TB_SC2 * C + c), sh2 = *reinterpret_cast<const float4*>(tab + TB_SH2B * C + c);
          const float4 c0 = *reinterpret_cast<const float4*>(tab + TB_C0 * C + c), c1 = *reinterpret_cast<const float4*>(tab + TB_C1 * C + c);
          const float4 gq = *reinterpret_cast<const float4*>(s_t + off_f(eslot, ecg * (NCOL / 4) + k));
          const float scv[4] = {sc2.x, sc2.y, sc2.z, sc2.w}, shv[4] = {sh2.x, sh2.y, sh2.z, sh2.w};
          const float c0v[4] = {c0.x, c0.y, c0.z, c0.w}, c1v[4] = {c1.x, c1.y, c1.z, c1.w};
          const float gvv[4] = {gq.x, gq.y, gq.z, gq.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float a = v[k * 4 + q];
            const float dy = fmaf(a, scv[q], shv[q]) > 0.f ? scv[q] * gvv[q] : 0.f;
            dz[k * 4 + q] = valid ? fmaf(-c1v[q], a, dy - c0v[q]) : 0.f;
          }
        }
#pragma unroll
        for (int k = 0; k < NCOL / 8; ++k) {
          uint2 h0, l0, h1, l1;
          split_bf16x4(make_float4(dz[k * 8], dz[k * 8 + 1], dz[k * 8 + 2], dz[k * 8 + 3]), h0, l0);
          split_bf16x4(make_float4(dz[k * 8 + 4], dz[k * 8 + 5], dz[k * 8 + 6], dz[k * 8 + 7]), h1, l1);
          const uint32_t ob = off_bf(eslot, ecg * NCOL + k * 8);
          *reinterpret_cast<uint4*>(dz_t + ob) = make_uint4(h0.x, h0.y, h1.x, h1.y);
          *reinterpret_cast<uint4*>(dz_t + T_BF + ob) = make_uint4(l0.x, l0.y, l1.x, l1.y);
        }
      }
    }
    if (D2) fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (D2) {
      // ================= MMA 2: da1 = dz2 W2;  dW2 += [dz2_hi ; dz2_lo]^T a1 ==============================
      if (warp == 0) {
        tc_fence_after();
        if (elect_one_sync()) {
          constexpr uint32_t IDESC_NN = make_idesc_bf16(TILE, C, 0, 1);     // A K-major, B = W2 read MN-major
          constexpr uint32_t IDESC_TN = make_idesc_bf16(128, C, 1, 1);      // M = 128 = [dz2_hi ; dz2_lo]^T, both MN-major: contraction over the slots
          const uint32_t a_u32 = sm_u32 + LT::OFF_A1, d_u32 = sm_u32 + LT::OFF_DZ, w_u32 = sm_u32 + LT::OFF_W2;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = make_desc(d_u32 + ks * 32u, 16, 1024, LAYOUT_SW128);
            const uint64_t a_lo = make_desc(d_u32 + T_BF + ks * 32u, 16, 1024, LAYOUT_SW128);
            const uint64_t b_hi = make_desc(w_u32 + ks * 2048u, 8192, 1024, LAYOUT_SW128);
            const uint64_t b_lo = make_desc(w_u32 + W_BF + ks * 2048u, 8192, 1024, LAYOUT_SW128);   // the mid part of W2
            umma_bf16(T_DA, a_lo, b_hi, IDESC_NN, ks > 0 ? 1u : 0u);
            umma_bf16(T_DA, a_hi, b_lo, IDESC_NN, 1u);
            umma_bf16(T_DA, a_hi, b_hi, IDESC_NN, 1u);
          }
          if (D2T) {
#pragma unroll
            for (int ks = 0; ks < TILE / 16; ++ks) {
              const uint64_t a_st = make_desc(d_u32 + ks * 2048u, T_BF, 1024, LAYOUT_SW128);      // rows 0-63 hi, 64-127 lo
              const uint64_t b_hi = make_desc(a_u32 + ks * 2048u, 8192, 1024, LAYOUT_SW128);
              const uint64_t b_lo = make_desc(a_u32 + T_BF + ks * 2048u, 8192, 1024, LAYOUT_SW128);   // the mid part of a1
              umma_bf16(T_DW, a_st, b_lo, IDESC_TN, (t > 0 || ks > 0) ? 1u : 0u);
              umma_bf16(T_DW, a_st, b_hi, IDESC_TN, 1u);
            }
          }
          umma_commit(smem_u32(&bar_mma));
        }
        __syncwarp();
      }
      mbar_wait_bounded(smem_u32(&bar_mma), mma_phase);
      mma_phase ^= 1u;
      tc_fence_after();

      // ================= epilogue 2: dy1 = da1 mask1 (overwrites the thread's own G entries) ==============
      {
        float v[NCOL];
        tmem_ldn(T_DA + ((uint32_t)(eq * 32) << 16) + (uint32_t)(ecg * NCOL), v);
        const bool valid = eslot < nvalid;
        if (eact) {
#pragma unroll
        for (int k = 0; k < NCOL / 4; ++k) {
          const int c = ecg * NCOL + k * 4;
          const float4 sc1 = *reinterpret_cast<const float4*>(tab + TB_SC1 * C + c), sh1 = *reinterpret_cast<const float4*>(tab + TB_SH1 * C + c);
          const float4 z = *reinterpret_cast<const float4*>(z_t + off_f(eslot, ecg * (NCOL / 4) + k));
          float4 o;
          o.x = (valid && fmaf(z.x, sc1.x, sh1.x) > 0.f) ? v[k * 4 + 0] : 0.f;
          o.y = (valid && fmaf(z.y, sc1.y, sh1.y) > 0.f) ? v[k * 4 + 1] : 0.f;
          o.z = (valid && fmaf(z.z, sc1.z, sh1.z) > 0.f) ? v[k * 4 + 2] : 0.f;
          o.w = (valid && fmaf(z.w, sc1.w, sh1.w) > 0.f) ? v[k * 4 + 3] : 0.f;
          *reinterpret_cast<float4*>(s_t + off_f(eslot, ecg * (NCOL / 4) + k)) = o;
        }
        }
      }
      tc_fence_before();
      __syncthreads();
    }

    // ================= row sweep: segmented sums over the rows of this tile ================================
    {
      const bool last_tile = t + 1 == ntiles;
      const int r_hi = last_tile ? r_end - 1 : r_last;
      const float* cin = carry + (t & 1) * 2 * C;
      float* cout = carry + ((t + 1) & 1) * 2 * C;
      float is1s[8], xm1s[8];
      if (D2) {
        const float4 i0 = *reinterpret_cast<const float4*>(tab + TB_IS1 * C + 4 * k8), i1 = *reinterpret_cast<const float4*>(tab + TB_IS1 * C + 32 + 4 * k8);
        const float4 m0 = *reinterpret_cast<const float4*>(tab + TB_XM1 * C + 4 * k8), m1 = *reinterpret_cast<const float4*>(tab + TB_XM1 * C + 32 + 4 * k8);
        is1s[0] = i0.x; is1s[1] = i0.y; is1s[2] = i0.z; is1s[3] = i0.w; is1s[4] = i1.x; is1s[5] = i1.y; is1s[6] = i1.z; is1s[7] = i1.w;
        xm1s[0] = m0.x; xm1s[1] = m0.y; xm1s[2] = m0.z; xm1s[3] = m0.w; xm1s[4] = m1.x; xm1s[5] = m1.y; xm1s[6] = m1.z; xm1s[7] = m1.w;
      }
      for (int r = row_lo + rl; r <= r_hi; r += THREADS / 8) {
        const int b = r == pre_r ? pre_b : __ldg(p.rowptr + r), e = r == pre_r ? pre_e : __ldg(p.rowptr + r + 1);
        const int lo = (int)(max((int64_t)b, s0) - s0), hi = (int)(min((int64_t)e, s1) - s0);
        float u[8], x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { u[i] = 0.f; x[i] = 0.f; }
        if (D1) {
          // S0 = sum w mask, S1 = sum w mask xhat2 over the row's slots of this tile; linear in the row: no carry
          for (int j = lo; j < hi; ++j) {
            const float4 w0 = *reinterpret_cast<const float4*>(s_t + off_f(j, k8)), w1 = *reinterpret_cast<const float4*>(s_t + off_f(j, k8 + 8));
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            const float ewj = ew_s[j];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float m = (__float_as_uint(wv[i]) & 1u) ? ewj : 0.f;
              u[i] += m;
              x[i] = fmaf(m, wv[i], x[i]);
            }
          }
          if (hi > lo) {
            const float di = __ldg(p.deg_inv + r);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.g + (int64_t)r * p.ldg + 4 * k8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.g + (int64_t)r * p.ldg + 32 + 4 * k8));
            const float gr[8] = {g0.x * di, g0.y * di, g0.z * di, g0.w * di, g1.x * di, g1.y * di, g1.z * di, g1.w * di};
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc_a[i] = fmaf(gr[i], u[i], acc_a[i]); acc_b[i] = fmaf(gr[i], x[i], acc_b[i]); }
          }
        } else {
          float lu[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) lu[i] = 0.f;
          for (int j = lo; j < hi; ++j) {
            const float4 d0 = *reinterpret_cast<const float4*>(s_t + off_f(j, k8)), d1 = *reinterpret_cast<const float4*>(s_t + off_f(j, k8 + 8));
            const float4 z0 = *reinterpret_cast<const float4*>(z_t + off_f(j, k8)), z1 = *reinterpret_cast<const float4*>(z_t + off_f(j, k8 + 8));
            const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            float4 at = make_float4(0.f, 0.f, 0.f, 0.f);
            if (D2T) at = rec_a[j];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float xh = fmaf(zv[i], is1s[i], xm1s[i]);
              lu[i] += dv[i];
              x[i] += xh;
              if (D2T) {
                acc_b[i] = fmaf(dv[i], xh, acc_b[i]);
                acc_t[D2T ? i * 4 + 0 : 0] = fmaf(dv[i], at.x, acc_t[D2T ? i * 4 + 0 : 0]);
                acc_t[D2T ? i * 4 + 1 : 0] = fmaf(dv[i], at.y, acc_t[D2T ? i * 4 + 1 : 0]);
                acc_t[D2T ? i * 4 + 2 : 0] = fmaf(dv[i], at.z, acc_t[D2T ? i * 4 + 2 : 0]);
                acc_t[D2T ? i * 4 + 3 : 0] = fmaf(dv[i], at.w, acc_t[D2T ? i * 4 + 3 : 0]);
              }
            }
          }
          if (D2T) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc_a[i] += lu[i];
          }
          if ((int64_t)b < s0) {      // the row started in the previous tile
            const float4 cu0 = *reinterpret_cast<const float4*>(cin + 4 * k8), cu1 = *reinterpret_cast<const float4*>(cin + 32 + 4 * k8);
            const float4 cx0 = *reinterpret_cast<const float4*>(cin + C + 4 * k8), cx1 = *reinterpret_cast<const float4*>(cin + C + 32 + 4 * k8);
            u[0] = cu0.x; u[1] = cu0.y; u[2] = cu0.z; u[3] = cu0.w; u[4] = cu1.x; u[5] = cu1.y; u[6] = cu1.z; u[7] = cu1.w;
            x[0] += cx0.x; x[1] += cx0.y; x[2] += cx0.z; x[3] += cx0.w; x[4] += cx1.x; x[5] += cx1.y; x[6] += cx1.z; x[7] += cx1.w;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) u[i] += lu[i];
          if ((int64_t)e <= s1) {     // the row ends here (empty rows: zeros)
            float* uo = p.U + (int64_t)r * C;
            float* xo = p.X + (int64_t)r * C;
            *reinterpret_cast<float4*>(uo + 4 * k8) = make_float4(u[0], u[1], u[2], u[3]);
            *reinterpret_cast<float4*>(uo + 32 + 4 * k8) = make_float4(u[4], u[5], u[6], u[7]);
            *reinterpret_cast<float4*>(xo + 4 * k8) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(xo + 32 + 4 * k8) = make_float4(x[4], x[5], x[6], x[7]);
          } else {                    // continues in the next tile
            *reinterpret_cast<float4*>(cout + 4 * k8) = make_float4(u[0], u[1], u[2], u[3]);
            *reinterpret_cast<float4*>(cout + 32 + 4 * k8) = make_float4(u[4], u[5], u[6], u[7]);
            *reinterpret_cast<float4*>(cout + C + 4 * k8) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(cout + C + 32 + 4 * k8) = make_float4(x[4], x[5], x[6], x[7]);
          }
        }
      }
      row_lo = ((int64_t)next_ptr > s1) ? r_last : r_last + 1;
    }
    __syncthreads();
  }
  if (D2 && ntiles == 0) {            // a range of rows without a single slot
    for (int r = r_begin + rl; r < r_end; r += THREADS / 8) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(p.U + (int64_t)r * C + 4 * k8) = z; *reinterpret_cast<float4*>(p.U + (int64_t)r * C + 32 + 4 * k8) = z;
      *reinterpret_cast<float4*>(p.X + (int64_t)r * C + 4 * k8) = z; *reinterpret_cast<float4*>(p.X + (int64_t)r * C + 32 + 4 * k8) = z;
    }
  }

  // ================= per-CTA partials (fixed order, no atomics) ============================================
  if (D2T) {
    // dW2 accumulator [128 = hi | lo rows of dz2^T][64]
    float v[NCOL];
    if (ntiles > 0) tmem_ldn(T_DW + ((uint32_t)(eq * 32) << 16) + (uint32_t)(ecg * NCOL), v);
    else {
#pragma unroll
      for (int i = 0; i < NCOL; ++i) v[i] = 0.f;
    }
    float* o = p.part_w2 + ((int64_t)blockIdx.x * 128 + eq * 32 + lane) * C + ecg * NCOL;
#pragma unroll
    for (int i = 0; i < NCOL; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  if (MODE != M_D2S) {
    constexpr int NV = D2T ? 48 : 16;
    float* red = reinterpret_cast<float*>(sm);                  // [THREADS / 8 rl][8 k8][NV], aliases the operand tiles
    float* mine = red + (size_t)(rl * 8 + k8) * NV;
#pragma unroll
    for (int i = 0; i < 8; ++i) { mine[i] = acc_a[i]; mine[8 + i] = acc_b[i]; }
    if (D2T) {
#pragma unroll
      for (int i = 0; i < 32; ++i) mine[16 + i] = acc_t[D2T ? i : 0];
    }
    __syncthreads();
    for (int tt = tid; tt < 8 * NV; tt += THREADS) {
      const int kk = tt / NV, vv = tt % NV;
      float s = 0.f;
      for (int q = 0; q < THREADS / 8; ++q) s += red[(size_t)(q * 8 + kk) * NV + vv];
      const int i = vv < 16 ? (vv & 7) : (vv - 16) >> 2;
      const int c = (i < 4) ? 4 * kk + i : 32 + 4 * kk + (i - 4);
      if (vv < 16) p.part[((int64_t)blockIdx.x * 2 + (vv >> 3)) * C + c] = s;
      else p.part_t[(int64_t)blockIdx.x * (C * 4) + c * 4 + ((vv - 16) & 3)] = s;
    }
  } else {
    float* red2 = reinterpret_cast<float*>(sm);                 // [SLG sl][16 gc][20]
    float* m2 = red2 + (size_t)(sl * 16 + gc) * 20;
#pragma unroll
    for (int i = 0; i < 16; ++i) m2[i] = acc_tx[TXA ? i : 0];
#pragma unroll
    for (int k = 0; k < 4; ++k) m2[16 + k] = acc_sa[TXA ? k : 0];
    __syncthreads();
    for (int tt = tid; tt < 16 * 20; tt += THREADS) {
      const int g2 = tt / 20, vv = tt % 20;
      float s = 0.f;
      for (int q = 0; q < SLG; ++q) s += red2[(size_t)(q * 16 + g2) * 20 + vv];
      float* o = p.part_t + (int64_t)blockIdx.x * (C * 4 + 4);
      if (vv < 16) o[(4 * g2 + (vv >> 2)) * 4 + (vv & 3)] = s;
      else if (g2 == 0) o[C * 4 + (vv - 16)] = s;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 256);
}

// slot-ordered records of the passes: (dst, src, eid, 0) | attribute row, once in CSR-by-target order (D1, D2T) and once
// in CSR-by-source order (D2S).  One launch builds both: the two dependent-load chains of a thread overlap.
__global__ void k_edge_bwd_records(const int32_t* __restrict__ src_t, const int32_t* __restrict__ dst_t,
                                   const int32_t* __restrict__ eid_t, const int32_t* __restrict__ slot_s,
                                   const float4* __restrict__ attr, int64_t E, int4* __restrict__ rec_idx_t,
                                   float4* __restrict__ rec_attr_t, int4* __restrict__ rec_idx_s,
                                   float4* __restrict__ rec_attr_s) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= E) return;
  const int64_t s = (int64_t)__ldg(slot_s + k);
  const int et = __ldg(eid_t + k), es = __ldg(eid_t + s);
  const int4 it = make_int4(__ldg(dst_t + k), __ldg(src_t + k), et, 0), is = make_int4(__ldg(dst_t + s), __ldg(src_t + s), es, 0);
  const float4 at = __ldg(attr + et), as = __ldg(attr + es);
  rec_idx_t[k] = it;
  rec_attr_t[k] = at;
  rec_idx_s[k] = is;
  rec_attr_s[k] = as;
}

// Node-level tail of the backward (see the header): dW2, dW1c and dP | dQ from the per-CTA partials and row sums.
//   blocks [0, nb_pq): rows of dpq;  then 64 blocks: dW2;  last block: dW1c.
__global__ void __launch_bounds__(256) k_edge_bwd_combine(
    int64_t N, int nb_pq, const float* __restrict__ Ut, const float* __restrict__ Xt, const float* __restrict__ Us,
    const float* __restrict__ Xs, const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ rowptr_s,
    const float* __restrict__ stat1, const float* __restrict__ bstat1, const float* __restrict__ part_w2, int grid2,
    const float* __restrict__ part_T, const float* __restrict__ part_tx, int grid1, float* __restrict__ dpq,
    float* __restrict__ dw2, float* __restrict__ dw1c) {
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < nb_pq) {
    // thread = one float4 of a dpq row: 32 chunks per row (16 of dP, 16 of dQ), 8 rows per block pass
    const int ch = tid & 31;
    const bool isq = ch >= 16;
    const int c = (ch & 15) * 4;
    const float4 sc = *reinterpret_cast<const float4*>(stat1 + c);
    const float4 m1 = *reinterpret_cast<const float4*>(bstat1 + c), m2 = *reinterpret_cast<const float4*>(bstat1 + C + c);
    const float* U = isq ? Us : Ut;
    const float* X = isq ? Xs : Xt;
    const int32_t* rp = isq ? rowptr_s : rowptr_t;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (tid >> 5); r < N; r += (int64_t)nb_pq * 8) {
      const float deg = (float)(__ldg(rp + r + 1) - __ldg(rp + r));
      const float4 u = __ldg(reinterpret_cast<const float4*>(U + r * C + c));
      const float4 x = __ldg(reinterpret_cast<const float4*>(X + r * C + c));
      float4 o;
      o.x = sc.x * (u.x - deg * m1.x - m2.x * x.x);
      o.y = sc.y * (u.y - deg * m1.y - m2.y * x.y);
      o.z = sc.z * (u.z - deg * m1.z - m2.z * x.z);
      o.w = sc.w * (u.w - deg * m1.w - m2.w * x.w);
      *reinterpret_cast<float4*>(dpq + r * (2 * C) + (isq ? C : 0) + c) = o;
    }
    return;
  }
  const int b = (int)blockIdx.x - nb_pq;
  if (b < 64) {
    // dW2[o][i] = sum over CTAs of (hi rows + lo rows): block = 64 outputs x 4 slices of the CTA range, fixed summation order
    if (!dw2) return;
    __shared__ float red[4][64];
    const int lo = tid & 63, sl = tid >> 6;
    const int idx = b * 64 + lo, o = idx >> 6, i = idx & 63;
    const int per = (grid2 + 3) / 4, k0 = sl * per, k1 = min(grid2, k0 + per);
    float s = 0.f;
    for (int k = k0; k < k1; ++k)
      s += part_w2[((int64_t)k * 128 + o) * C + i] + part_w2[((int64_t)k * 128 + 64 + o) * C + i];
    red[sl][lo] = s;
    __syncthreads();
    if (sl == 0) dw2[idx] = (red[0][lo] + red[1][lo]) + (red[2][lo] + red[3][lo]);
    return;
  }
  if (!dw1c) return;
  {
    const int c = tid >> 2, k = tid & 3;
    float T = 0.f, TX = 0.f, SA = 0.f;
    for (int q = 0; q < grid2; ++q) T += part_T[(int64_t)q * (C * 4) + c * 4 + k];
    for (int q = 0; q < grid1; ++q) {
      TX += part_tx[(int64_t)q * (C * 4 + 4) + c * 4 + k];
      SA += part_tx[(int64_t)q * (C * 4 + 4) + C * 4 + k];
    }
    dw1c[c * 4 + k] = stat1[c] * (T - bstat1[c] * SA - bstat1[C + c] * TX);
  }
}

template <int MODE, int TILE>
static cudaError_t launch_t(const Params& p, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_bwd<MODE, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L<TILE>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_edge_bwd<MODE, TILE><<<grid, L<TILE>::THREADS, L<TILE>::SMEM_BYTES, st>>>(p);
  return cudaSuccess;
}

// Default: one 512-thread CTA per SM over 128-slot tiles.  YOLAT_EB_TILE=64 selects two 256-thread CTAs per SM over 64-slot
// tiles -- measured no faster (N = 320 000: D1 321 / D2T 603 / D2S 480 us vs 330 / 585 / 456; step 1.50 vs 1.46 ms): the
// passes are bound by their instruction count (296 M warp instructions for D2T, profiles/r2_c_bwd_*), not by the
// latencies a second resident CTA would hide, and the M = 64 accumulators idle half of every epilogue warp.
static int tile_slots() {
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("YOLAT_EB_TILE");
    v = (e && atoi(e) == 64) ? 64 : 128;
  }
  return v;
}

template <int MODE>
static cudaError_t launch(const Params& p, int grid, cudaStream_t st) {
  return tile_slots() == 128 ? launch_t<MODE, 128>(p, grid, st) : launch_t<MODE, 64>(p, grid, st);
}

}  // namespace eb

int edge_bwd_grid(int64_t E) {
  const int tile = eb::tile_slots();
  const int64_t tiles = cdiv(E > 0 ? E : 1, tile);
  const int64_t cap = (int64_t)kNumSMs * (tile == 64 ? 2 : 1);       // persistent: every resident CTA owns one slot range
  return (int)(tiles < cap ? tiles : cap);
}

// Workspace of the fused edge backward, in floats (all taken from `ws` in this order).
void edge_bwd_layout(Arena& ws, int64_t N, int64_t E, EdgeBwdWs* o) {
  const int grid = edge_bwd_grid(E);
  o->grid = grid;
  o->rec_t = ws.take(8 * E + 8);
  o->rec_s = ws.take(8 * E + 8);
  o->rec5 = ws.take(edge_records_floats(E));
  o->Ut = ws.take(N * eb::C); o->Xt = ws.take(N * eb::C);
  o->Us = ws.take(N * eb::C); o->Xs = ws.take(N * eb::C);
  o->part1 = ws.take((int64_t)grid * 2 * eb::C);
  o->part_tx = ws.take((int64_t)grid * (eb::C * 4 + 4));
  o->part2 = ws.take((int64_t)grid * 2 * eb::C);
  o->part_T = ws.take((int64_t)grid * eb::C * 4);
  o->part_w2 = ws.take((int64_t)grid * 128 * eb::C);
  o->bstat2 = ws.take(2 * eb::C);
  o->bstat1 = ws.take(2 * eb::C);
}

// The whole edge path of the GraphConv backward (training mode, C = 64): grads of W2, b2, BN2, BN1, b1, W1c and
// dpq [N, 2C] = dP | dQ; the caller turns dpq into dW1a / dW1b / dx with two node-level GEMMs.
int edge_bwd_fused(const GraphView& g, int64_t N, int64_t E, const float* pq, int64_t ldpq, const float* attr,
                   const float* w1, int Cin, const float* b1, const float* stat1, const float* gamma1, const float* w2,
                   const float* b2, const float* stat2, const float* gamma2, const float* ew, const float* g_out,
                   int64_t ldgo, const EdgeBwdWs& w, float* dpq, float* dw1c, float* dw2, float* db1, float* dg1,
                   float* dbe1, float* db2, float* dg2, float* dbe2, cudaStream_t st) {
  if (E <= 0) return YOLAT_ERR_INVALID;
  using namespace eb;
  const int grid = w.grid;
  int4* ri_t = reinterpret_cast<int4*>(w.rec_t);
  float4* ra_t = reinterpret_cast<float4*>(w.rec_t + 4 * E);
  int4* ri_s = reinterpret_cast<int4*>(w.rec_s);
  float4* ra_s = reinterpret_cast<float4*>(w.rec_s + 4 * E);
  const unsigned nb = (unsigned)cdiv(E, 256);
  k_edge_bwd_records<<<nb, 256, 0, st>>>(g.src_t, g.dst_t, g.eid_t, g.slot_s, reinterpret_cast<const float4*>(attr), E, ri_t, ra_t,
                                         ri_s, ra_s);
  YOLAT_CHECK_LAUNCH();

  Params p{};
  p.deg_inv = g.deg_inv; p.N = N; p.E = E; p.pq = pq; p.ldpq = (uint32_t)ldpq;
  p.w1c = w1 + 2 * Cin; p.ld1 = 2 * Cin + 4; p.b1 = b1; p.stat1 = stat1; p.w2 = w2; p.b2 = b2; p.stat2 = stat2;
  p.ew = ew; p.g = g_out; p.ldg = ldgo;
  cudaError_t e;
  if (edge_fused2_enabled()) {
    // D1 on the channel-major forward kernel (edge_fused2.cu, F_BSTAT): its record format is the forward's
    YOLAT_TRY(edge_records(g, E, attr, ldpq, w.rec5, st));
    YOLAT_TRY(edge_fused2(g, N, E, EF_BSTAT, pq, ldpq, w.rec5, w1, Cin, b1, stat1, w2, b2, stat2, ew, nullptr, nullptr, w.part1,
                          g_out, ldgo, nullptr, 0, st));
  } else {   // D1: BN2 backward statistics
    Params q = p;
    q.rowptr = g.rowptr_t; q.rec_idx = ri_t; q.rec_attr = ra_t; q.part = w.part1;
    ProfScope prof(YOLAT_PROF_EDGE_BWD_D1, st);
    e = launch<M_D1>(q, grid, st);
    if (e != cudaSuccess) { set_last_error(e); return YOLAT_ERR_LAUNCH; }
    YOLAT_CHECK_LAUNCH();
  }
  YOLAT_TRY(bn_bwd_finalize(w.part1, grid, E, C, stat2, gamma2, 1, w.bstat2, dg2, dbe2, db2, st));
  {   // D2T: dW2, BN1 backward statistics, row sums by target
    Params q = p;
    q.rowptr = g.rowptr_t; q.rec_idx = ri_t; q.rec_attr = ra_t; q.bstat2 = w.bstat2;
    q.U = w.Ut; q.X = w.Xt; q.part = w.part2; q.part_t = w.part_T; q.part_w2 = w.part_w2;
    ProfScope prof(YOLAT_PROF_EDGE_BWD_D2T, st);
    e = launch<M_D2T>(q, grid, st);
    if (e != cudaSuccess) { set_last_error(e); return YOLAT_ERR_LAUNCH; }
    YOLAT_CHECK_LAUNCH();
  }
  {   // D2S: row sums by source
    Params q = p;
    q.rowptr = g.rowptr_s; q.rec_idx = ri_s; q.rec_attr = ra_s; q.bstat2 = w.bstat2; q.U = w.Us; q.X = w.Xs; q.part_t = w.part_tx;
    ProfScope prof(YOLAT_PROF_EDGE_BWD_D2S, st);
    e = launch<M_D2S>(q, grid, st);
    if (e != cudaSuccess) { set_last_error(e); return YOLAT_ERR_LAUNCH; }
    YOLAT_CHECK_LAUNCH();
  }
  YOLAT_TRY(bn_bwd_finalize(w.part2, grid, E, C, stat1, gamma1, 1, w.bstat1, dg1, dbe1, db1, st));
  const int nb_pq = (int)(cdiv(N, 8) < 592 ? cdiv(N, 8) : 592);
  k_edge_bwd_combine<<<nb_pq + 65, 256, 0, st>>>(N, nb_pq, w.Ut, w.Xt, w.Us, w.Xs, g.rowptr_t, g.rowptr_s, stat1, w.bstat1,
                                                 w.part_w2, grid, w.part_T, w.part_tx, grid, dpq, dw2, dw1c);
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

}  // namespace yolat
