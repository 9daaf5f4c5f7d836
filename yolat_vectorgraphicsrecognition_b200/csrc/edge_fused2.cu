// edge_fused2.cu -- K-EDGE v6: the fused gather -> edge-MLP -> scatter kernel with a CHANNEL-MAJOR accumulator.
//
// Same contract, same gather side and same record ring as edge_fused.cu (v5; kept for pass A and as YOLAT_EF=v5), but the
// tensor-core product is transposed:  z2^T [64 channels x 128 slots] = W2 [64 x 64] . a1^T, i.e. W2 is the A operand
// (M = 64) and the a1 stage -- unchanged in shared memory, rows = slots, K-major -- is read as the B operand (N = 128).
// The accumulator then lives in TMEM with one CHANNEL per lane and one SLOT per column, which turns everything the v5
// epilogue did through a shared-memory staging tile (17 M shared wavefronts, 28 % barrier stalls) into
// a sequential in-register sweep of one thread per channel:
//   F_STATS  sum / sum of squares of the channel over the tile's columns                      (BN2 batch statistics)
//   F_AGG    BN2 + ReLU (+ edge weight), running sum of the current target row, flushed at the row's last slot as
//            out[r][c] = base[r][c] + sum / deg  -- the segmented mean needs no staging tile, no barrier, no carry
//            buffer (the running sum simply stays in its register across tile boundaries), no atomics
//   F_BSTAT  (backward pass D1) BN2 backward statistics: per row S0 = sum w mask, S1 = sum w mask xhat2, folded with the
//            row's upstream gradient g[r][c] / deg at the row's last slot
//   F_TAPE   z1 / z2 for the tape backward (eval-mode autograd only).
// Cost model: an M = 64 tcgen05.mma occupies the tensor core like an M = 128 one (64 cycles for N = 128, K = 8), so the
// 24 MMAs of a tile take 1536 cycles instead of ~1150 shared-memory-bound ones -- but that is now the ONLY per-tile cost
// outside the gather: 4 epilogue warps (one per TMEM lane quarter; an M = 64 accumulator keeps rows 16 q .. 16 q + 15 in
// lanes 32 q .. 32 q + 15, tools/umma_probe_m64.cu) spend ~5 instructions per slot, and the 8 warps the v5 epilogue needed
// go to the gather (16 gather warps, 4 slots per thread and tile).
// The rows a tile finishes form a contiguous node range, so their `base` (or g) rows are one contiguous block of global
// memory: the control warp stages it with ONE cp.async.bulk per tile (double-buffered, 64 rows; wider ranges fall back
// to direct loads).
#include <cstdlib>
#include "common.cuh"
#include "tc.cuh"

namespace yolat {
namespace ef2 {

using namespace tc;

constexpr int C = 64;
constexpr int TILE = 128;
constexpr int E_WARPS = 4, G_WARPS = 16, C_WARPS = 4;
constexpr int E_THREADS = E_WARPS * 32, G_THREADS = G_WARPS * 32, THREADS = (E_WARPS + G_WARPS + C_WARPS) * 32;
constexpr int SPT = TILE * 16 / G_THREADS;      // 4 slots per gather thread and tile
constexpr int NSLG = TILE / SPT;                // 32 slot groups
constexpr uint32_t A_KB = TILE * 128;           // one k-block (32 channels) of the a1 tile: 16 KB
constexpr uint32_t A_HI = 2 * A_KB;
constexpr uint32_t A_STAGE = 2 * A_HI;          // hi | lo: 64 KB; two stages
constexpr uint32_t W_KB = C * 128;
constexpr uint32_t W_HI = 2 * W_KB;
constexpr int RING = 6;
constexpr int PF = RING - 1;
constexpr int PFL2 = 3;
constexpr int RING_BYTES = TILE * 32;
constexpr int BASE_ROWS = 64;                   // rows of `base` / g staged per tile
constexpr uint32_t OFF_W = 2 * A_STAGE;
constexpr uint32_t OFF_RING = OFF_W + 2 * W_HI;
constexpr uint32_t OFF_BASE = OFF_RING + RING * RING_BYTES;
constexpr uint32_t OFF_EW = OFF_BASE + 2 * BASE_ROWS * C * 4;
constexpr uint32_t OFF_TAB = OFF_EW + 4 * TILE * 4;
constexpr uint32_t SMEM_BYTES = OFF_TAB + 6 * C * 4 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

enum { F_TAPE = 1, F_STATS = 2, F_AGG = 4, F_BSTAT = 16 };

struct Params {
  const int32_t* rowptr; const int32_t* dst; const int32_t* eid;
  int64_t N, E;
  const float* pq; uint32_t ldpq;
  const int4* rec_idx; const float4* rec_attr;
  const float* w1c; int ld1; const float* b1;
  const float* stat1;
  const float* w2; const float* b2;
  const float* stat2;
  const float* ew;
  float* z1; float* z2;
  float* part;                       // [grid][2][C]: F_STATS / F_BSTAT
  const float* rows; int64_t ldr;    // F_AGG: base (may alias out); F_BSTAT: g = dL/dout
  float* out; int64_t ldo;           // F_AGG
};

__device__ __forceinline__ int row_at_or_after(const Params& p, int64_t s) {
  if (s <= 0) return 0;
  if (s >= p.E) return (int)p.N;
  const int v = p.dst[s];
  return (p.rowptr[v] == (int32_t)s) ? v : v + 1;
}

template <int FLAGS, bool EW>
__global__ void __launch_bounds__(THREADS, 1) k_edge_fused2(const Params p) {
  constexpr bool ROWS = (FLAGS & (F_AGG | F_BSTAT)) != 0;
  constexpr bool FOLD = (FLAGS & F_AGG) && !(FLAGS & F_TAPE);     // BN2 scale folded into the rows of W2
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_a_full[2], bar_a_empty[2], bar_acc_full[2], bar_acc_empty[2], bar_base_full[2], bar_base_empty[2];
  __shared__ uint64_t bar_ring_full[RING], bar_ring_empty[RING];
  __shared__ uint32_t tmem_slot;

  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = ((raw_u32 + 1023u) & ~1023u) - raw_u32;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_u32 = raw_u32 + pad;
  uint8_t* w_tile = sm + OFF_W;
  uint8_t* ring = sm + OFF_RING;
  float* base_s = reinterpret_cast<float*>(sm + OFF_BASE);       // [2][BASE_ROWS][C]
  float* ew_s = reinterpret_cast<float*>(sm + OFF_EW);           // [4][TILE]: written by gather(t), read by epilogue(t)
  float* tab = reinterpret_cast<float*>(sm + OFF_TAB);           // sc2 | sh2b | is2 | xm2 | b2

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_a_full[i]), G_WARPS);
      mbar_init(smem_u32(&bar_a_empty[i]), 1);
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), E_WARPS);
      mbar_init(smem_u32(&bar_base_full[i]), 1);
      mbar_init(smem_u32(&bar_base_empty[i]), E_WARPS);
    }
#pragma unroll
    for (int i = 0; i < RING; ++i) {
      mbar_init(smem_u32(&bar_ring_full[i]), 1);
      mbar_init(smem_u32(&bar_ring_empty[i]), G_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 256);
  for (int idx = tid; idx < RING * RING_BYTES / 16; idx += THREADS) reinterpret_cast<int4*>(ring)[idx] = make_int4(0, 0, 0, 0);
  // W2 (A operand, K-major: row m = output channel, 64 k) -> hi / lo, resident for the whole kernel
  for (int idx = tid; idx < C * 16; idx += THREADS) {
    const int n = idx >> 4, c = idx & 15;
    float4 v = __ldg(reinterpret_cast<const float4*>(p.w2 + n * C + c * 4));
    if (FOLD) {
      const float sc = __ldg(p.stat2 + n);
      v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    }
    const uint32_t off = (uint32_t)(c >> 3) * W_KB + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u +
                         (uint32_t)(((c & 7) ^ (n & 7)) << 4);
    store_split(w_tile, w_tile + W_HI, off, v);
  }
  if (tid < C) {
    const float b2 = p.b2 ? __ldg(p.b2 + tid) : 0.f;
    tab[4 * C + tid] = b2;
    if (FLAGS & (F_AGG | F_BSTAT)) {
      const float sc = __ldg(p.stat2 + tid), sh = __ldg(p.stat2 + C + tid), mu = __ldg(p.stat2 + 2 * C + tid), is = __ldg(p.stat2 + 3 * C + tid);
      tab[tid] = sc;
      tab[C + tid] = fmaf(b2, sc, sh);
      tab[2 * C + tid] = is;
      tab[3 * C + tid] = (b2 - mu) * is;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;

  const int r_begin = row_at_or_after(p, p.E * (int64_t)blockIdx.x / gridDim.x);
  const int r_end = row_at_or_after(p, p.E * (int64_t)(blockIdx.x + 1) / gridDim.x);
  const int64_t s_begin = r_begin < p.N ? p.rowptr[r_begin] : p.E;
  const int64_t s_end = r_end < p.N ? p.rowptr[r_end] : p.E;
  const int ntiles = (int)((s_end - s_begin + TILE - 1) / TILE);

  if (warp < E_WARPS) {
    // =========================== epilogue: one thread per channel, one sweep over the tile's slots ================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    const int q = warp, c = 16 * q + (lane & 15);
    const bool active = lane < 16;
    const bool staged = ROWS && p.rows && p.ldr == C;      // the control warp stages the tile's rows in shared memory
    const float sh = (FLAGS & (F_AGG | F_BSTAT)) ? tab[C + c] : 0.f;
    const float sc = (FLAGS & (F_AGG | F_BSTAT)) ? tab[c] : 0.f;
    const float is2 = (FLAGS & F_BSTAT) ? tab[2 * C + c] : 0.f, xm2 = (FLAGS & F_BSTAT) ? tab[3 * C + c] : 0.f;
    const float b2 = tab[4 * C + c];
    float st_a = 0.f, st_b = 0.f;          // F_STATS: sum, sum of squares;  F_BSTAT: sum G S0, sum G S1
    float acc = 0.f, acc2 = 0.f;           // running sums of the current row
    int cnt = 0;                           // slots of the current row seen so far (= its in-degree at the flush)
    for (int t = 0; t < ntiles; ++t) {
      const int a = t & 1;
      const int64_t s0 = s_begin + (int64_t)t * TILE;
      const int nvalid = (int)min((int64_t)TILE, s_end - s0);
      // row bookkeeping: target row of this lane's slots and "last slot of its row" masks (requested before the wait)
      int dcur[4];
      uint32_t em[4];
      if (ROWS) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int s = 32 * k + lane;
          int dn = -1;
          dcur[k] = -2;
          if (s < nvalid) {
            dcur[k] = __ldg(p.dst + s0 + s);
            if (s0 + s + 1 < s_end) dn = __ldg(p.dst + s0 + s + 1);
          }
          em[k] = __ballot_sync(0xffffffffu, s < nvalid && dcur[k] != dn);
        }
      }
      const int r0 = ROWS ? __shfl_sync(0xffffffffu, dcur[0], 0) : 0;
      mbar_wait_bounded(smem_u32(&bar_acc_full[a]), (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      if (ROWS) mbar_wait_bounded(smem_u32(&bar_base_full[a]), (uint32_t)((t >> 1) & 1));
      const float* rows_s = base_s + a * BASE_ROWS * C;
      const float* ew_t = ew_s + (t & 3) * TILE;
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        if (32 * k >= nvalid) break;
        float v[32];
        tmem_ld32(tmem_d + (uint32_t)(a * TILE) + ((uint32_t)(q * 32) << 16) + (uint32_t)(k * 32), v);
        const uint32_t emk = !ROWS ? 0u : (k == 0 ? em[0] : k == 1 ? em[1] : k == 2 ? em[2] : em[3]);
        const int dk = !ROWS ? 0 : (k == 0 ? dcur[0] : k == 1 ? dcur[1] : k == 2 ? dcur[2] : dcur[3]);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int s = 32 * k + j;
          if (s < nvalid) {
            if ((FLAGS & F_TAPE) && active) p.z2[(s0 + s) * C + c] = v[j] + b2;
            if (FLAGS & F_STATS) {
              st_a += v[j];
              st_b = fmaf(v[j], v[j], st_b);
            }
            if (FLAGS & F_AGG) {
              float m = FOLD ? fmaxf(v[j] + sh, 0.f) : fmaxf(fmaf(v[j], sc, sh), 0.f);
              if (EW) m *= ew_t[s];
              acc += m;
            }
            if (FLAGS & F_BSTAT) {
              const float wv = fmaf(v[j], sc, sh) > 0.f ? (EW ? ew_t[s] : 1.f) : 0.f;
              acc += wv;
              acc2 = fmaf(wv, fmaf(v[j], is2, xm2), acc2);
            }
            if (ROWS) {
              ++cnt;
              if ((emk >> j) & 1u) {          // warp-uniform: the row of slot s ends here
                const int r = __shfl_sync(0xffffffffu, dk, j);
                const int idx = r - r0;
                float rv = 0.f;
                if (active) {
                  if (staged && idx < BASE_ROWS) rv = rows_s[idx * C + c];
                  else if (p.rows) rv = __ldg(p.rows + (int64_t)r * p.ldr + c);
                }
                const float di = __frcp_rn((float)cnt);
                if (FLAGS & F_AGG) {
                  if (active) p.out[(int64_t)r * p.ldo + c] = fmaf(acc, di, rv);
                } else {
                  const float gd = rv * di;
                  st_a = fmaf(gd, acc, st_a);
                  st_b = fmaf(gd, acc2, st_b);
                }
                acc = 0.f; acc2 = 0.f; cnt = 0;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar_acc_empty[a]));
        if (ROWS) mbar_arrive(smem_u32(&bar_base_empty[a]));
      }
    }
    if ((FLAGS & (F_STATS | F_BSTAT)) && active) {
      if (FLAGS & F_STATS) {
        // sum (acc + b2) and sum (acc + b2)^2 from the raw accumulator sums
        const float n = (float)(s_end - s_begin);
        p.part[((int64_t)blockIdx.x * 2 + 0) * C + c] = fmaf(n, b2, st_a);
        p.part[((int64_t)blockIdx.x * 2 + 1) * C + c] = fmaf(b2, fmaf(n, b2, 2.f * st_a), st_b);
      } else {
        p.part[((int64_t)blockIdx.x * 2 + 0) * C + c] = st_a;
        p.part[((int64_t)blockIdx.x * 2 + 1) * C + c] = st_b;
      }
    }
  } else if (warp >= E_WARPS + G_WARPS) {
    // =========================== control warp: record ring + row staging (TMA) + MMA issue ======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == E_WARPS + G_WARPS) {
      auto fillc = [&](int j) {
        if (j >= ntiles) return;
        const int st = j % RING;
        if (j >= RING) mbar_wait_bounded(smem_u32(&bar_ring_empty[st]), (uint32_t)(((j / RING) - 1) & 1));
        const int64_t s0 = s_begin + (int64_t)j * TILE;
        const uint32_t bytes = (uint32_t)min((int64_t)TILE, s_end - s0) * 16u;
        const uint32_t bar = smem_u32(&bar_ring_full[st]);
        const uint32_t dst = smem_u32(ring + st * RING_BYTES);
        mbar_expect_tx(bar, 2u * bytes);
        bulk_g2s(dst, p.rec_idx + s0, bytes, bar);
        bulk_g2s(dst + TILE * 16, p.rec_attr + s0, bytes, bar);
      };
      // rows (base / g) of tile j: the contiguous node range [dst of its first slot, dst of its last slot], up to BASE_ROWS
      auto rows_range = [&](int j, int& r0, int& nr) {
        r0 = 0; nr = 0;
        if (!ROWS || j >= ntiles) return;
        const int64_t s0 = s_begin + (int64_t)j * TILE;
        const int nv = (int)min((int64_t)TILE, s_end - s0);
        r0 = __ldg(p.dst + s0);
        nr = __ldg(p.dst + s0 + nv - 1) - r0 + 1;
        if (nr > BASE_ROWS) nr = BASE_ROWS;
      };
      auto fill_rows = [&](int j, int r0, int nr) {
        if (!ROWS || j >= ntiles) return;
        const int a = j & 1;
        if (j >= 2) mbar_wait_bounded(smem_u32(&bar_base_empty[a]), (uint32_t)(((j >> 1) - 1) & 1));
        const uint32_t bar = smem_u32(&bar_base_full[a]);
        if (p.rows && p.ldr == C) {
          mbar_expect_tx(bar, (uint32_t)nr * C * 4u);
          bulk_g2s(smem_u32(base_s + a * BASE_ROWS * C), p.rows + (int64_t)r0 * C, (uint32_t)nr * C * 4u, bar);
        } else {
          mbar_arrive(bar);                    // nothing staged: the epilogue reads the rows from global memory
        }
      };
      int nr0 = 0, nnr = 0;
      if (lane == 0) {
        for (int j = 0; j < PF; ++j) fillc(j);
        int r0, nr;
        rows_range(0, r0, nr);
        fill_rows(0, r0, nr);
        rows_range(1, nr0, nnr);
      }
      constexpr uint32_t IDESC = make_idesc(C, TILE, 0, 0);      // M = 64 channels, N = 128 slots
      for (int t = 0; t < ntiles; ++t) {
        int cr0 = 0, cnr = 0;
        if (lane == 0) {
          fillc(t + PF);
          cr0 = nr0; cnr = nnr;              // range of tile t + 1 (loaded one iteration ago)
          rows_range(t + 2, nr0, nnr);       // its loads complete under this tile's MMA issue
        }
        __syncwarp();
        const int s = t & 1;
        const uint32_t a_u32 = sm_u32 + (uint32_t)s * A_STAGE, w_u32 = sm_u32 + OFF_W;
        mbar_wait_bounded(smem_u32(&bar_a_full[s]), (uint32_t)((t >> 1) & 1));
        if (t >= 2) mbar_wait_bounded(smem_u32(&bar_acc_empty[s]), (uint32_t)(((t >> 1) - 1) & 1));
        tc_fence_after();
        const uint32_t d = tmem_d + (uint32_t)(s * TILE);
        if (elect_one_sync()) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t wo = (uint32_t)kb * W_KB + (uint32_t)ks * 32u;
              const uint32_t ao = (uint32_t)kb * A_KB + (uint32_t)ks * 32u;
              const uint64_t w_hi = make_desc(w_u32 + wo, 16, 1024, LAYOUT_SW128);
              const uint64_t w_lo = make_desc(w_u32 + W_HI + wo, 16, 1024, LAYOUT_SW128);
              const uint64_t a_hi = make_desc(a_u32 + ao, 16, 1024, LAYOUT_SW128);
              const uint64_t a_lo = make_desc(a_u32 + A_HI + ao, 16, 1024, LAYOUT_SW128);
              umma_tf32(d, w_hi, a_lo, IDESC, (kb > 0 || ks > 0) ? 1u : 0u);
              umma_tf32(d, w_lo, a_hi, IDESC, 1u);
              umma_tf32(d, w_hi, a_hi, IDESC, 1u);
            }
          }
          umma_commit(smem_u32(&bar_acc_full[s]));
          umma_commit(smem_u32(&bar_a_empty[s]));
        }
        __syncwarp();
        // rows of tile t + 1: after the issue, so that waiting for the epilogue of tile t - 1 (stage reuse) cannot delay it
        if (lane == 0) fill_rows(t + 1, cr0, cnr);
      }
    }
  } else {
    // =========================== gather warps (16): thread = 4 channels x 4 slots of every tile ====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
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
        sct[qq] = __ldg(p.stat1 + c);
        sht[qq] = __ldg(p.stat1 + C + c);
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
    const char* pbase = reinterpret_cast<const char*>(p.pq + gc * 4);
    float4 qv[SPT], pv[SPT];
    auto issue = [&](int st, int i) {
      const int2 ds = *reinterpret_cast<const int2*>(ring + st * RING_BYTES + (i * NSLG + sl) * 16);
      pv[i] = __ldg(reinterpret_cast<const float4*>(pbase + (uint32_t)ds.x));
      qv[i] = __ldg(reinterpret_cast<const float4*>(pbase + (uint32_t)ds.y));
    };
    if (ntiles > 0) {
      mbar_wait_bounded(smem_u32(&bar_ring_full[0]), 0u);
#pragma unroll
      for (int i = 0; i < SPT; ++i) issue(0, i);
    }
    const uint32_t off0 = (uint32_t)(gc >> 3) * A_KB + (uint32_t)(sl >> 3) * 1024u + (uint32_t)(sl & 7) * 128u +
                          ((((uint32_t)gc & 7u) ^ ((uint32_t)sl & 7u)) << 4);
    for (int it = 0; it < ntiles; ++it) {
      const int s = it & 1;
      const int64_t s0 = s_begin + (int64_t)it * TILE;
      const int nvalid = (int)min((int64_t)TILE, s_end - s0);
      const int st_cur = it % RING, st_next = (it + 1) % RING;
      const bool have_next = it + 1 < ntiles;
      if (have_next) mbar_wait_bounded(smem_u32(&bar_ring_full[st_next]), (uint32_t)(((it + 1) / RING) & 1));
      if (it >= 2) mbar_wait_bounded(smem_u32(&bar_a_empty[s]), (uint32_t)(((it >> 1) - 1) & 1));
      uint8_t* a_tile = sm + (uint32_t)s * A_STAGE;
      const float4* attr_s = reinterpret_cast<const float4*>(ring + st_cur * RING_BYTES + TILE * 16);
      if (it + PFL2 < ntiles) {
        const int st_pf = (it + PFL2) % RING;
        mbar_wait_bounded(smem_u32(&bar_ring_full[st_pf]), (uint32_t)(((it + PFL2) / RING) & 1));
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
        if (EW && gc == 0) {
          const int e = reinterpret_cast<const int4*>(ring + st_cur * RING_BYTES)[slot].z;
          ew_s[(it & 3) * TILE + slot] = slot < nvalid ? __ldg(p.ew + e) : 0.f;
        }
        float2 v0 = ffma2s(at.x, w[0][0], bias1[0]), v1 = ffma2s(at.x, w[0][1], bias1[1]);
        v0 = ffma2s(at.y, w[1][0], v0); v1 = ffma2s(at.y, w[1][1], v1);
        v0 = ffma2s(at.z, w[2][0], v0); v1 = ffma2s(at.z, w[2][1], v1);
        v0 = ffma2s(at.w, w[3][0], v0); v1 = ffma2s(at.w, w[3][1], v1);
        v0 = fadd2(fadd2(v0, make_float2(pv[i].x, pv[i].y)), make_float2(qv[i].x, qv[i].y));
        v1 = fadd2(fadd2(v1, make_float2(pv[i].z, pv[i].w)), make_float2(qv[i].z, qv[i].w));
        if (have_next) issue(st_next, i);
        if ((FLAGS & F_TAPE) && slot < nvalid)
          *reinterpret_cast<float4*>(p.z1 + (s0 + slot) * C + gc * 4) = make_float4(v0.x, v0.y, v1.x, v1.y);
        float2 a0 = ffma2(v0, sc1[0], sh1[0]), a1 = ffma2(v1, sc1[1], sh1[1]);
        a0.x = fmaxf(a0.x, 0.f); a0.y = fmaxf(a0.y, 0.f); a1.x = fmaxf(a1.x, 0.f); a1.y = fmaxf(a1.y, 0.f);
        store_split_trunc(a_tile, a_tile + A_HI, off0 + (uint32_t)(i * (NSLG / 8)) * 1024u, a0, a1);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar_a_full[s]));
        mbar_arrive(smem_u32(&bar_ring_empty[st_cur]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, 256);
}

template <int FLAGS, bool EW>
static cudaError_t launch_cfg(const Params& p, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_fused2<FLAGS, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_edge_fused2<FLAGS, EW><<<grid, THREADS, SMEM_BYTES, st>>>(p);
  return cudaSuccess;
}
template <int FLAGS>
static cudaError_t launch(const Params& p, int grid, cudaStream_t st) {
  return p.ew ? launch_cfg<FLAGS, true>(p, grid, st) : launch_cfg<FLAGS, false>(p, grid, st);
}

}  // namespace ef2

// Off by default: YOLAT_EF=v6 selects this kernel for the MMA passes and the backward statistics pass D1.
// Measured at N = 320 000, E = 1 280 000 (profiles/r2_g_ef2_*.txt): F_STATS 144 us (v5: 130), F_AGG 821 us (v5: 172),
// F_BSTAT 715 us (eb::k_edge_bwd<D1>: 383).  Shared-memory wavefronts drop from 17 M to 10-11 M as designed, but the
// per-channel sweep is ONE warp per TMEM lane quarter executing ~2.4 k dependent instructions per tile next to four
// gather warps on the same scheduler: ~10 cycles per instruction = 23 k cycles per tile.  The thread-per-slot epilogue
// of v5 (many threads, short chains, one staging round trip) is the better balance on this machine; splitting the sweep
// over more warps needs more than the 1024 threads a CTA can have next to 16 gather warps.
bool edge_fused2_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YOLAT_EF");
    v = (e && e[0] == 'v' && e[1] == '6') ? 1 : 0;
  }
  return v == 1;
}

// flags: EF_TAPE | EF_STATS | EF_AGG (common.cuh) or EF_BSTAT (backward pass D1: rows = g_out, part = [grid][2][C])
int edge_fused2(const GraphView& g, int64_t N, int64_t E, int flags, const float* pq, int64_t ldpq, const float* rec,
                const float* w1, int Cin, const float* b1, const float* stat1, const float* w2, const float* b2,
                const float* stat2, const float* ew, float* z1, float* z2, float* part, const float* rows, int64_t ldr,
                float* out, int64_t ldo, cudaStream_t st) {
  if (E <= 0) return YOLAT_OK;
  ef2::Params p{};
  p.rowptr = g.rowptr_t; p.dst = g.dst_t; p.eid = g.eid_t; p.N = N; p.E = E; p.pq = pq; p.ldpq = (uint32_t)ldpq;
  p.rec_idx = reinterpret_cast<const int4*>(rec);
  p.rec_attr = reinterpret_cast<const float4*>(rec + 4 * E);
  p.w1c = w1 + 2 * Cin; p.ld1 = 2 * Cin + 4; p.b1 = b1; p.stat1 = stat1; p.w2 = w2; p.b2 = b2; p.stat2 = stat2; p.ew = ew;
  p.z1 = z1; p.z2 = z2; p.part = part; p.rows = rows; p.ldr = ldr; p.out = out; p.ldo = ldo;
  if ((flags & EF_AGG) && rows != out) return YOLAT_ERR_UNSUPPORTED;   // rows without slots keep `out` = base as it is
  const int grid = edge_fused_grid(E);
  cudaError_t e;
  ProfScope prof((flags & EF_BSTAT) ? YOLAT_PROF_EDGE_BWD_D1 : (flags & EF_AGG) ? YOLAT_PROF_EDGE_FUSED_AGG : YOLAT_PROF_EDGE_FUSED_STATS, st);
  switch (flags) {
    case EF_STATS: e = ef2::launch<ef2::F_STATS>(p, grid, st); break;
    case EF_STATS | EF_TAPE: e = ef2::launch<ef2::F_STATS | ef2::F_TAPE>(p, grid, st); break;
    case EF_AGG: e = ef2::launch<ef2::F_AGG>(p, grid, st); break;
    case EF_AGG | EF_TAPE: e = ef2::launch<ef2::F_AGG | ef2::F_TAPE>(p, grid, st); break;
    case EF_BSTAT: e = ef2::launch<ef2::F_BSTAT>(p, grid, st); break;
    default: return YOLAT_ERR_INVALID;
  }
  if (e != cudaSuccess) { set_last_error(e); return YOLAT_ERR_LAUNCH; }
  YOLAT_CHECK_LAUNCH();
  return YOLAT_OK;
}

}  // namespace yolat
