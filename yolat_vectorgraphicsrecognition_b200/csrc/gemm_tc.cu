// gemm_tc.cu -- tcgen05 (5th-generation tensor core) GEMM with TMEM accumulators and fp32-accurate
// 3xTF32 operand splitting, for every Linear on the hot path:
//     y  = x W^T + b   (gcn_lib/sparse/torch_nn.py:58)        NT
//     dx = dy W                                               NN
//     dW = dy^T x                                             TN   (reduction over rows, deterministic split-K)
//
// Why 3xTF32: the parity bar is 1e-4 against an fp32 reference (BASELINE.json), kind::tf32 alone gives
// ~1e-3.  Each operand is split in registers into hi = rn_tf32(v) and lo = v - hi (exact in fp32) and the
// product is accumulated as  Ahi*Bhi + Ahi*Blo + Alo*Bhi  in the fp32 TMEM accumulator -- error ~2^-21.
//
// Structure of one CTA (256 threads, one 128 x BN output tile, BN = 64 | 128):
//   * all 8 warps load the next k-block (32 fp32 = one 128-byte swizzle row per tile row) of A and B from
//     global memory into registers, apply the optional fused BN+ReLU operand prologue, split hi/lo and
//     store into shared memory in the canonical UMMA SWIZZLE_128B layout (K-major or MN-major, so the
//     transposed operands of NN / TN need no transposition pass);
//   * one thread issues 12 tcgen05.mma.kind::tf32 (4 k-steps x 3 split products) per k-block and commits
//     them to the stage's mbarrier; the tensor core runs asynchronously while the next k-block is loaded
//     (double-buffered shared memory);
//   * epilogue: tcgen05.ld TMEM -> registers -> padded shared tile -> coalesced stores (+bias, +accumulate),
//     optional per-tile column sum / sum-of-squares partials (BatchNorm statistics without re-reading z).
#include <cuda.h>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "tc.cuh"

namespace yolat {
namespace tc {

constexpr int BM = 128;        // output rows per CTA = TMEM lanes
constexpr int BK = 32;         // fp32 per k-block = 128 bytes = one swizzle row
constexpr int THREADS = 256;
constexpr int STAGES = 2;
constexpr uint32_t A_BYTES = BM * 128;

struct Params {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int M, N; int64_t K;
  const float* bias;
  const float* a_sc; const float* a_sh;
  const float* b_sc; const float* b_sh;
  int accumulate;
  int64_t k_chunk;       // reduction length per split (multiple of BK)
  float* part;           // split-K partial tiles [splits][M][N] (null when gridDim.z == 1)
  float* stat_part;      // [gridDim.y][2][N] column sum / sumsq of the stored tile (null = off)
  int a_vec, b_vec;      // 16-byte vector loads are legal for A / B
  int c_vec;             // 16-byte vector stores are legal for C
  int tma_store;         // TMA variant: C leaves through cp.async.bulk.tensor stores (maps.c) instead of per-thread STG
  int dbg;               // YOLAT_TC_DBG (timing experiments only): 1 = no hi/lo conversion, 2 = no C stores, 4 = no MMA
};

// ---- operand loaders -------------------------------------------------------------------------------
// K-major operand: element (row, k) at base[row*ld + k].  ROWS x 32 tile = ROWS*8 16-byte chunks;
// thread t owns chunk c = t&7 of rows (t>>3) + 32*i.
template <int ROWS>
struct KMajor {
  static constexpr int PER_THREAD = ROWS / 32;
  float4 v[PER_THREAD];
  __device__ __forceinline__ void fetch(const float* __restrict__ base, int64_t ld, int row0, int rows_total, int64_t k0,
                                        int64_t k_end, int vec, const float* __restrict__ sc,
                                        const float* __restrict__ sh) {
    const int t = threadIdx.x, c = t & 7;
    const int64_t k = k0 + c * 4;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int row = row0 + (t >> 3) + 32 * i;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows_total && k < k_end) {
        const float* p = base + (int64_t)row * ld + k;
        if (vec && k + 3 < k_end) {
          x = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          x.x = __ldg(p);
          if (k + 1 < k_end) x.y = __ldg(p + 1);
          if (k + 2 < k_end) x.z = __ldg(p + 2);
          if (k + 3 < k_end) x.w = __ldg(p + 3);
        }
        if (sc) {   // fused BatchNorm + ReLU on the operand, per reduction index k
          x.x = fmaxf(fmaf(x.x, __ldg(sc + k), __ldg(sh + k)), 0.f);
          x.y = (k + 1 < k_end) ? fmaxf(fmaf(x.y, __ldg(sc + k + 1), __ldg(sh + k + 1)), 0.f) : 0.f;
          x.z = (k + 2 < k_end) ? fmaxf(fmaf(x.z, __ldg(sc + k + 2), __ldg(sh + k + 2)), 0.f) : 0.f;
          x.w = (k + 3 < k_end) ? fmaxf(fmaf(x.w, __ldg(sc + k + 3), __ldg(sh + k + 3)), 0.f) : 0.f;
        }
      }
      v[i] = x;
    }
  }
  __device__ __forceinline__ void stash(uint8_t* hi, uint8_t* lo) const {
    const int t = threadIdx.x, c = t & 7;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int row = (t >> 3) + 32 * i;
      const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
      store_split(hi, lo, off, v[i]);
    }
  }
  // Fast path (16-byte aligned operand, K % 4 == 0): unconditional vector loads from clamped addresses -- no branch
  // between the load and its first use, so the loads really stay in flight until stash_fast() two k-blocks later.
  // Out-of-range chunks are zeroed and the BN+ReLU prologue is applied at stash time.
  uint32_t ok;
  __device__ __forceinline__ void fetch_fast(const float* __restrict__ base, int64_t ld, int row0, int rows_total,
                                             int64_t k0, int64_t k_end) {
    const int t = threadIdx.x, c = t & 7;
    const int64_t k = k0 + c * 4;
    const bool kin = k < k_end;
    const int64_t kc = kin ? k : k0;
    ok = 0;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int row = row0 + (t >> 3) + 32 * i;
      const bool in = kin && row < rows_total;
      const int rc = row < rows_total ? row : rows_total - 1;
      v[i] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)rc * ld + kc));
      ok |= in ? (1u << i) : 0u;
    }
  }
  __device__ __forceinline__ void stash_fast(uint8_t* hi, uint8_t* lo, int64_t k0, const float* __restrict__ sc,
                                             const float* __restrict__ sh) const {
    const int t = threadIdx.x, c = t & 7;
    float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sc && ok) {      // per reduction index k (uniform branch on sc; ok != 0 implies k in range)
      const int64_t k = k0 + c * 4;
      s4 = make_float4(__ldg(sc + k), __ldg(sc + k + 1), __ldg(sc + k + 2), __ldg(sc + k + 3));
      h4 = make_float4(__ldg(sh + k), __ldg(sh + k + 1), __ldg(sh + k + 2), __ldg(sh + k + 3));
    }
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int row = (t >> 3) + 32 * i;
      float4 x = v[i];
      if (sc) {
        x.x = fmaxf(fmaf(x.x, s4.x, h4.x), 0.f); x.y = fmaxf(fmaf(x.y, s4.y, h4.y), 0.f);
        x.z = fmaxf(fmaf(x.z, s4.z, h4.z), 0.f); x.w = fmaxf(fmaf(x.w, s4.w, h4.w), 0.f);
      }
      if (!((ok >> i) & 1u)) x = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint32_t off = (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((c ^ (row & 7)) << 4);
      store_split_fast(hi, lo, off, x);
    }
  }
};

// MN-major operand: element (mn, k) at base[k*ld + mn].  MN x 32 tile; one 128-byte smem row holds 32
// consecutive mn at one k; thread t owns chunk cm = t % (MN/4) of k-rows t/(MN/4) + (1024/MN)*i.
template <int MN>
struct MNMajor {
  static constexpr int CPR = MN / 4;                 // 16-byte chunks per k-row
  static constexpr int KR_PER_PASS = THREADS / CPR;  // k-rows covered per pass
  static constexpr int PER_THREAD = BK / KR_PER_PASS;
  float4 v[PER_THREAD];
  __device__ __forceinline__ void fetch(const float* __restrict__ base, int64_t ld, int mn0, int mn_total, int64_t k0,
                                        int64_t k_end, int vec, const float* __restrict__ sc,
                                        const float* __restrict__ sh) {
    const int t = threadIdx.x, cm = t % CPR;
    const int mn = mn0 + cm * 4;
    float4 s = make_float4(1.f, 1.f, 1.f, 1.f), h = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sc && mn < mn_total) {
      s.x = __ldg(sc + mn); h.x = __ldg(sh + mn);
      if (mn + 1 < mn_total) { s.y = __ldg(sc + mn + 1); h.y = __ldg(sh + mn + 1); }
      if (mn + 2 < mn_total) { s.z = __ldg(sc + mn + 2); h.z = __ldg(sh + mn + 2); }
      if (mn + 3 < mn_total) { s.w = __ldg(sc + mn + 3); h.w = __ldg(sh + mn + 3); }
    }
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int64_t k = k0 + t / CPR + KR_PER_PASS * i;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < k_end && mn < mn_total) {
        const float* p = base + k * ld + mn;
        if (vec && mn + 3 < mn_total) {
          x = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          x.x = __ldg(p);
          if (mn + 1 < mn_total) x.y = __ldg(p + 1);
          if (mn + 2 < mn_total) x.z = __ldg(p + 2);
          if (mn + 3 < mn_total) x.w = __ldg(p + 3);
        }
        if (sc) {   // fused BatchNorm + ReLU on the operand, per output index mn
          x.x = fmaxf(fmaf(x.x, s.x, h.x), 0.f);
          x.y = (mn + 1 < mn_total) ? fmaxf(fmaf(x.y, s.y, h.y), 0.f) : 0.f;
          x.z = (mn + 2 < mn_total) ? fmaxf(fmaf(x.z, s.z, h.z), 0.f) : 0.f;
          x.w = (mn + 3 < mn_total) ? fmaxf(fmaf(x.w, s.w, h.w), 0.f) : 0.f;
        }
      }
      v[i] = x;
    }
  }
  // 32-bit MN-major operands only exist in the SWIZZLE_128B_BASE32B layout: atoms of 32 mn x 4 k (4 rows of 128
  // bytes), the 32-byte chunk index of a row XORed with the row index (Swizzle<2,5,2> on byte addresses).
  // Tile = [k_atom (8)][mn_atom (MN/32)][4 k-rows][128 B]  =>  LBO = 512 B, SBO = MN/32 * 512 B.
  __device__ __forceinline__ void stash(uint8_t* hi, uint8_t* lo) const {
    const int t = threadIdx.x, cm = t % CPR;
    const int c16 = cm & 7;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int kr = t / CPR + KR_PER_PASS * i;
      const int kr4 = kr & 3;
      const uint32_t off = (uint32_t)(kr >> 2) * (uint32_t)(MN / 32 * 512) + (uint32_t)(cm >> 3) * 512u +
                           (uint32_t)kr4 * 128u + (uint32_t)((((c16 >> 1) ^ kr4) << 5) | ((c16 & 1) << 4));
      store_split(hi, lo, off, v[i]);
    }
  }
  // Fast path: see KMajor::fetch_fast (requires mn_total % 4 == 0 and a 16-byte aligned operand).
  uint32_t ok;
  __device__ __forceinline__ void fetch_fast(const float* __restrict__ base, int64_t ld, int mn0, int mn_total,
                                             int64_t k0, int64_t k_end) {
    const int t = threadIdx.x, cm = t % CPR;
    const int mn = mn0 + cm * 4;
    const bool min_ = mn < mn_total;
    const int mnc = min_ ? mn : 0;
    ok = 0;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int64_t k = k0 + t / CPR + KR_PER_PASS * i;
      const bool in = min_ && k < k_end;
      const int64_t kc = k < k_end ? k : k0;
      v[i] = __ldg(reinterpret_cast<const float4*>(base + kc * ld + mnc));
      ok |= in ? (1u << i) : 0u;
    }
  }
  __device__ __forceinline__ void stash_fast(uint8_t* hi, uint8_t* lo, int mn0, const float* __restrict__ sc,
                                             const float* __restrict__ sh) const {
    const int t = threadIdx.x, cm = t % CPR;
    const int c16 = cm & 7;
    float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sc && ok) {      // per output index mn
      const int mn = mn0 + cm * 4;
      s4 = make_float4(__ldg(sc + mn), __ldg(sc + mn + 1), __ldg(sc + mn + 2), __ldg(sc + mn + 3));
      h4 = make_float4(__ldg(sh + mn), __ldg(sh + mn + 1), __ldg(sh + mn + 2), __ldg(sh + mn + 3));
    }
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int kr = t / CPR + KR_PER_PASS * i;
      const int kr4 = kr & 3;
      float4 x = v[i];
      if (sc) {
        x.x = fmaxf(fmaf(x.x, s4.x, h4.x), 0.f); x.y = fmaxf(fmaf(x.y, s4.y, h4.y), 0.f);
        x.z = fmaxf(fmaf(x.z, s4.z, h4.z), 0.f); x.w = fmaxf(fmaf(x.w, s4.w, h4.w), 0.f);
      }
      if (!((ok >> i) & 1u)) x = make_float4(0.f, 0.f, 0.f, 0.f);
      const uint32_t off = (uint32_t)(kr >> 2) * (uint32_t)(MN / 32 * 512) + (uint32_t)(cm >> 3) * 512u +
                           (uint32_t)kr4 * 128u + (uint32_t)((((c16 >> 1) ^ kr4) << 5) | ((c16 & 1) << 4));
      store_split_fast(hi, lo, off, x);
    }
  }
};

template <int MODE, int BN>
struct Loaders {
  using ALoad = typename std::conditional<MODE == GEMM_TN, MNMajor<BM>, KMajor<BM>>::type;
  using BLoad = typename std::conditional<MODE == GEMM_NT, KMajor<BN>, MNMajor<BN>>::type;
};

// ---- the kernel ------------------------------------------------------------------------------------
template <int MODE, int BN>
__global__ void __launch_bounds__(THREADS, 1) k_tc_gemm(const Params p) {
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr bool A_MN = (MODE == GEMM_TN);
  constexpr bool B_MN = (MODE != GEMM_NT);
  constexpr uint32_t IDESC = make_idesc(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
  constexpr int LDS = BN + 4;   // padded epilogue tile row (floats)

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t mbar_empty[STAGES];
  __shared__ uint64_t mbar_done;
  __shared__ uint32_t tmem_base_slot;

  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = ((raw_u32 + 1023u) & ~1023u) - raw_u32;
  uint8_t* tiles = smem_raw + pad;            // 1024-byte aligned (SWIZZLE_128B atoms)
  const uint32_t tiles_u32 = raw_u32 + pad;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * p.k_chunk;
  const int64_t k_end = min(p.K, k_begin + p.k_chunk);
  const int nkb = (int)((k_end - k_begin + BK - 1) / BK);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&mbar_empty[s]), 1);
    mbar_init(smem_u32(&mbar_done), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_slot), BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_slot;

  // Two k-blocks of global loads are kept in flight per thread (register sets 0/1 <-> smem stages 0/1):
  // the loop is latency-bound on the loads, not on the tensor core.
  typename Loaders<MODE, BN>::ALoad la[2];
  typename Loaders<MODE, BN>::BLoad lb[2];
  auto fetch = [&](int kb, typename Loaders<MODE, BN>::ALoad& ra, typename Loaders<MODE, BN>::BLoad& rb) {
    const int64_t k0 = k_begin + (int64_t)kb * BK;
    ra.fetch(p.A, p.lda, m0, p.M, k0, k_end, p.a_vec, p.a_sc, p.a_sh);
    rb.fetch(p.B, p.ldb, n0, p.N, k0, k_end, p.b_vec, MODE == GEMM_TN ? p.b_sc : nullptr, p.b_sh);
  };
  auto step = [&](int kb, int s, typename Loaders<MODE, BN>::ALoad& ra, typename Loaders<MODE, BN>::BLoad& rb) {
    const int use = kb / STAGES;
    if (use > 0) mbar_wait(smem_u32(&mbar_empty[s]), (uint32_t)((use - 1) & 1));   // MMAs that read stage s are done
    uint8_t* st = tiles + (uint32_t)s * STAGE_BYTES;
    ra.stash(st, st + A_BYTES);
    rb.stash(st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES);
    if (kb + STAGES < nkb) fetch(kb + STAGES, ra, rb);
    fence_proxy_async_smem();                  // generic-proxy smem writes -> visible to the tensor core
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one_sync()) {   // converged warp, one elected lane: descriptors stay in uniform registers
      const uint32_t sa = tiles_u32 + (uint32_t)s * STAGE_BYTES;
      const uint32_t sb = sa + 2 * A_BYTES;
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {
        const uint32_t a_off = A_MN ? (uint32_t)ks * (BM / 32 * 1024) : (uint32_t)ks * 32u;
        const uint32_t b_off = B_MN ? (uint32_t)ks * (BN / 32 * 1024) : (uint32_t)ks * 32u;
        const uint64_t a_hi = A_MN ? make_desc(sa + a_off, 512, BM / 32 * 512, LAYOUT_SW128_BASE32B)
                                   : make_desc(sa + a_off, 16, 1024, LAYOUT_SW128);
        const uint64_t a_lo = A_MN ? make_desc(sa + A_BYTES + a_off, 512, BM / 32 * 512, LAYOUT_SW128_BASE32B)
                                   : make_desc(sa + A_BYTES + a_off, 16, 1024, LAYOUT_SW128);
        const uint64_t b_hi = B_MN ? make_desc(sb + b_off, 512, BN / 32 * 512, LAYOUT_SW128_BASE32B)
                                   : make_desc(sb + b_off, 16, 1024, LAYOUT_SW128);
        const uint64_t b_lo = B_MN ? make_desc(sb + B_BYTES + b_off, 512, BN / 32 * 512, LAYOUT_SW128_BASE32B)
                                   : make_desc(sb + B_BYTES + b_off, 16, 1024, LAYOUT_SW128);
        umma_tf32(tmem_d, a_lo, b_hi, IDESC, (kb > 0 || ks > 0) ? 1u : 0u);   // small terms first
        umma_tf32(tmem_d, a_hi, b_lo, IDESC, 1u);
        umma_tf32(tmem_d, a_hi, b_hi, IDESC, 1u);
      }
      umma_commit(smem_u32(&mbar_empty[s]));
      if (kb + 1 == nkb) umma_commit(smem_u32(&mbar_done));
      }
      __syncwarp();
    }
  };

  static_assert(STAGES == 2, "the main loop is unrolled by the stage count");
  if (nkb > 0) fetch(0, la[0], lb[0]);
  if (nkb > 1) fetch(1, la[1], lb[1]);
  for (int kb = 0; kb < nkb; kb += 2) {
    step(kb, 0, la[0], lb[0]);
    if (kb + 1 < nkb) step(kb + 1, 1, la[1], lb[1]);
  }

  // ---- epilogue ----------------------------------------------------------------------------------
  float* cs = reinterpret_cast<float*>(tiles);   // [BM][LDS], aliases the (now idle) operand stages
  if (nkb > 0) {
    mbar_wait(smem_u32(&mbar_done), 0u);
    tc_fence_after();
    const int q = warp & 3, h = warp >> 2;       // TMEM lane quarter / column half of this warp
    const int r = q * 32 + lane;
#pragma unroll
    for (int j = 0; j < BN / 64; ++j) {
      float v[32];
      const int col = h * (BN / 2) + j * 32;
      tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
      float* dst = cs + r * LDS + col;
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  } else {
    for (int i = tid; i < BM * LDS; i += THREADS) cs[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();

  const bool split = gridDim.z > 1;
  const int rows = min(BM, p.M - m0), cols = min(BN, p.N - n0);
  if (split) {
    float* dst = p.part + ((int64_t)blockIdx.z * p.M + m0) * p.N + n0;
    for (int idx = tid; idx < BM * (BN / 4); idx += THREADS) {
      const int rr = idx / (BN / 4), c4 = (idx % (BN / 4)) * 4;
      if (rr >= rows) continue;
      const float4 v = *reinterpret_cast<const float4*>(cs + rr * LDS + c4);
      float* o = dst + (int64_t)rr * p.N + c4;
      if ((p.N & 3) == 0 && c4 + 3 < cols) {
        *reinterpret_cast<float4*>(o) = v;
      } else {
        if (c4 < cols) o[0] = v.x;
        if (c4 + 1 < cols) o[1] = v.y;
        if (c4 + 2 < cols) o[2] = v.z;
        if (c4 + 3 < cols) o[3] = v.w;
      }
    }
  } else {
    for (int idx = tid; idx < BM * (BN / 4); idx += THREADS) {
      const int rr = idx / (BN / 4), c4 = (idx % (BN / 4)) * 4;
      if (rr >= rows || c4 >= cols) continue;
      float4 v = *reinterpret_cast<const float4*>(cs + rr * LDS + c4);
      if (p.bias) {
        v.x += __ldg(p.bias + n0 + c4);
        if (c4 + 1 < cols) v.y += __ldg(p.bias + n0 + c4 + 1);
        if (c4 + 2 < cols) v.z += __ldg(p.bias + n0 + c4 + 2);
        if (c4 + 3 < cols) v.w += __ldg(p.bias + n0 + c4 + 3);
      }
      float* o = p.C + (int64_t)(m0 + rr) * p.ldc + n0 + c4;
      if (p.c_vec && c4 + 3 < cols) {
        if (p.accumulate) {
          const float4 old = *reinterpret_cast<const float4*>(o);
          v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
        }
        *reinterpret_cast<float4*>(o) = v;
      } else {
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c4 + e < cols) o[e] = p.accumulate ? o[e] + vv[e] : vv[e];
      }
    }
    if (p.stat_part) {   // column statistics of z = acc + bias over this tile's valid rows (block-uniform branch)
      constexpr int GROUPS = THREADS / BN;          // row groups working on the same column
      constexpr int RPG = BM / GROUPS;
      __shared__ float red[2][THREADS];
      const int c = tid % BN, g = tid / BN;
      const float b = (p.bias && c < cols) ? __ldg(p.bias + n0 + c) : 0.f;
      float s = 0.f, ss = 0.f;
      const int r_end = min(rows, (g + 1) * RPG);
#pragma unroll 8
      for (int rr = g * RPG; rr < r_end; ++rr) {
        const float z = cs[rr * LDS + c] + b;
        s += z;
        ss = fmaf(z, z, ss);
      }
      red[0][tid] = s; red[1][tid] = ss;
      __syncthreads();
      if (tid < BN && c < cols) {
#pragma unroll
        for (int q = 1; q < GROUPS; ++q) { s += red[0][q * BN + c]; ss += red[1][q * BN + c]; }
        p.stat_part[((int64_t)blockIdx.y * 2 + 0) * p.N + n0 + c] = s;
        p.stat_part[((int64_t)blockIdx.y * 2 + 1) * p.N + n0 + c] = ss;
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, BN);
}

// ---- persistent, warp-specialised variant ---------------------------------------------------------------------------
// One CTA per SM walks the (n-tile, m-tile, k-split) list round-robin.  Roles, connected by mbarriers only:
//   warps 0..7  producers: global -> registers (2 k-blocks in flight) -> BN+ReLU prologue -> 3xTF32 split -> smem ring
//               (3 stages of 64 KB for BN = 128, 4 stages of 48 KB for BN = 64); they run ahead across tile boundaries;
//   warp  8     MMA issue: one elected lane, 12 tcgen05.mma per k-block into one of two TMEM accumulators;
//   warps 9..12 epilogue: tcgen05.ld (lane quarter = warp & 3) -> bias / accumulate -> direct 128-byte row stores, or
//               split-K partial tiles; per-warp column statistics through a private 32 x 33 transpose tile.
// The epilogue of tile i overlaps the main loop of tile i+1, which the one-tile-per-CTA kernel above cannot do.
//
// TMA variant (template flag TMA): a 14th warp issues cp.async.bulk.tensor loads of the raw fp32 k-blocks straight into
// the stage's `hi` buffers (the tensor maps are encoded with the swizzle mode the UMMA descriptors expect: SWIZZLE_128B
// for K-major operands, SWIZZLE_128B_ATOM_32B for 32-bit MN-major ones; out-of-range rows / reduction tails arrive as
// zeros), and warps 0..7 become converters: they wait for the bytes, split every 16-byte chunk IN PLACE (hi back to
// where it was, lo into the neighbouring buffer -- elementwise, so the swizzle never has to be undone) and hand the
// stage to the MMA warp.  No global load passes through registers; the loads of NST k-blocks are in flight per CTA.
constexpr int WS_PRODUCERS = 8, WS_THREADS = (WS_PRODUCERS + 1 + 4) * 32, WS_THREADS_TMA = WS_THREADS + 32;
struct alignas(64) TmaMaps { CUtensorMap a, b, c; };
template <int BN> struct WsCfg {
  static constexpr int NST = (BN == 128) ? 3 : 4;
  static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * BN * 128;
  // epilogue scratch per warp: a 32 x 33 transpose tile for the column statistics, or (TMA stores) two 32 x 32 fp32
  // boxes in the SWIZZLE_128B layout of the output tensor map
  static constexpr uint32_t XPOSE_BYTES = 4 * 2 * 4096;
  static constexpr size_t SMEM = (size_t)NST * STAGE_BYTES + XPOSE_BYTES + 1024;
};

template <int MODE, int BN, bool FAST, bool TMA>
__global__ void __launch_bounds__(TMA ? WS_THREADS_TMA : WS_THREADS, 1)
k_tc_gemm_ws(const Params p, const int gn, const int gm, const int ksplit, const __grid_constant__ TmaMaps maps) {
  using Cfg = WsCfg<BN>;
  constexpr int NST = Cfg::NST;
  constexpr uint32_t B_BYTES = BN * 128;
  constexpr uint32_t STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr bool A_MN = (MODE == GEMM_TN);
  constexpr bool B_MN = (MODE != GEMM_NT);
  constexpr uint32_t IDESC = make_idesc(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[NST], bar_empty[NST], bar_acc_full[2], bar_acc_empty[2], bar_tma[NST];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = ((raw_u32 + 1023u) & ~1023u) - raw_u32;
  uint8_t* tiles = smem_raw + pad;
  const uint32_t tiles_u32 = raw_u32 + pad;
  float* xpose = reinterpret_cast<float*>(tiles + NST * STAGE_BYTES);
  // MN-major operands staged by TMA: one 32 x 32 box per 32 output indices => [mn_atom][k_atom (8)][4 k-rows][128 B]
  // (the register loaders keep the k-atom outermost); only the two strides of the descriptor differ.
  constexpr uint32_t A_LBO = TMA ? 4096u : 512u, A_SBO = TMA ? 512u : (uint32_t)(BM / 32 * 512);
  constexpr uint32_t B_LBO = TMA ? 4096u : 512u, B_SBO = TMA ? 512u : (uint32_t)(BN / 32 * 512);
  constexpr uint32_t A_KSTEP = TMA ? 1024u : (uint32_t)(BM / 32 * 1024), B_KSTEP = TMA ? 1024u : (uint32_t)(BN / 32 * 1024);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = gn * gm * ksplit;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NST; ++i) {
      mbar_init(smem_u32(&bar_full[i]), WS_PRODUCERS * 32);
      mbar_init(smem_u32(&bar_empty[i]), 1);
      mbar_init(smem_u32(&bar_tma[i]), 1);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), 4 * 32);
    }
    fence_barrier_init();
  }
  if (warp == WS_PRODUCERS) tmem_alloc(smem_u32(&tmem_base_slot), 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_slot;

  // tile id -> (n-tile fastest: neighbouring CTAs share their A rows in L2)
  struct Cur { int tile, kb, nkb, m0, n0, z; int64_t kbeg, kend; };
  auto load_tile = [&](int tile, Cur& c) {
    c.tile = tile; c.kb = 0; c.nkb = 0;
    if (tile >= total) return;
    const int nx = tile % gn, rest = tile / gn;
    c.n0 = nx * BN; c.m0 = (rest % gm) * BM; c.z = rest / gm;
    c.kbeg = (int64_t)c.z * p.k_chunk;
    c.kend = min(p.K, c.kbeg + p.k_chunk);
    c.nkb = (int)((c.kend - c.kbeg + BK - 1) / BK);
  };

  if (TMA && warp < WS_PRODUCERS) {
    // =========================== converters (TMA variant) ================================================
    Cur s;
    load_tile(blockIdx.x, s);
    int st = 0;
    uint32_t use = 0;
    while (s.tile < total) {
      mbar_wait(smem_u32(&bar_tma[st]), use & 1u);            // the raw fp32 k-block of A and B has landed
      uint8_t* sp = tiles + (uint32_t)st * STAGE_BYTES;
      if (!(p.dbg & 1)) {
#pragma unroll
      for (int i = 0; i < (int)(A_BYTES / 16) / (WS_PRODUCERS * 32); ++i) {
        const uint32_t off = (uint32_t)(tid + i * WS_PRODUCERS * 32) * 16u;
        store_split_fast(sp, sp + A_BYTES, off, *reinterpret_cast<const float4*>(sp + off));
      }
#pragma unroll
      for (int i = 0; i < (int)(B_BYTES / 16) / (WS_PRODUCERS * 32); ++i) {
        const uint32_t off = (uint32_t)(tid + i * WS_PRODUCERS * 32) * 16u;
        store_split_fast(sp + 2 * A_BYTES, sp + 2 * A_BYTES + B_BYTES, off,
                         *reinterpret_cast<const float4*>(sp + 2 * A_BYTES + off));
      }
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bar_full[st]));
      if (++st == NST) { st = 0; ++use; }
      if (++s.kb >= s.nkb) load_tile(s.tile + gridDim.x, s);
    }
  } else if (TMA && warp == WS_PRODUCERS + 5) {
    // =========================== TMA issue (TMA variant) =================================================
    if (lane == 0) { tma_prefetch_desc(&maps.a); tma_prefetch_desc(&maps.b); if (p.tma_store) tma_prefetch_desc(&maps.c); }
    Cur f;
    load_tile(blockIdx.x, f);
    int st = 0;
    uint32_t use = 0;
    while (f.tile < total) {
      if (use > 0) mbar_wait(smem_u32(&bar_empty[st]), (use - 1) & 1u);     // the MMAs that read this stage are done
      if (elect_one_sync()) {
        const uint32_t bar = smem_u32(&bar_tma[st]);
        const uint32_t sa = tiles_u32 + (uint32_t)st * STAGE_BYTES, sb = sa + 2 * A_BYTES;
        const int32_t k0 = (int32_t)(f.kbeg + (int64_t)f.kb * BK);
        mbar_expect_tx(bar, A_BYTES + B_BYTES);
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + (uint32_t)j * 4096u, &maps.a, f.m0 + 32 * j, k0, bar);
        } else {
          tma_load_2d(sa, &maps.a, k0, f.m0, bar);
        }
        if (B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + (uint32_t)j * 4096u, &maps.b, f.n0 + 32 * j, k0, bar);
        } else {
          tma_load_2d(sb, &maps.b, k0, f.n0, bar);
        }
      }
      __syncwarp();
      if (++st == NST) { st = 0; ++use; }
      if (++f.kb >= f.nkb) load_tile(f.tile + gridDim.x, f);
    }
  } else if (warp < WS_PRODUCERS) {
    // =========================== producers ===============================================================
    typename Loaders<MODE, BN>::ALoad la[2];
    typename Loaders<MODE, BN>::BLoad lb[2];
    Cur f, s;
    load_tile(blockIdx.x, f);
    s = f;
    auto next = [&](Cur& c) { if (++c.kb >= c.nkb) load_tile(c.tile + gridDim.x, c); };
    auto fetch = [&](Cur& c, typename Loaders<MODE, BN>::ALoad& ra, typename Loaders<MODE, BN>::BLoad& rb) {
      if (c.tile >= total) return;
      const int64_t k0 = c.kbeg + (int64_t)c.kb * BK;
      if (FAST) {
        ra.fetch_fast(p.A, p.lda, c.m0, p.M, k0, c.kend);
        rb.fetch_fast(p.B, p.ldb, c.n0, p.N, k0, c.kend);
      } else {
        ra.fetch(p.A, p.lda, c.m0, p.M, k0, c.kend, p.a_vec, p.a_sc, p.a_sh);
        rb.fetch(p.B, p.ldb, c.n0, p.N, k0, c.kend, p.b_vec, MODE == GEMM_TN ? p.b_sc : nullptr, p.b_sh);
      }
      next(c);
    };
    fetch(f, la[0], lb[0]);
    fetch(f, la[1], lb[1]);
    int st = 0;
    uint32_t use = 0;     // how many times the ring has wrapped
    auto step = [&](typename Loaders<MODE, BN>::ALoad& ra, typename Loaders<MODE, BN>::BLoad& rb) {
      if (use > 0) mbar_wait(smem_u32(&bar_empty[st]), (use - 1) & 1u);     // the MMAs that read this stage are done
      uint8_t* sp = tiles + (uint32_t)st * STAGE_BYTES;
      if (FAST) {
        const int64_t k0 = s.kbeg + (int64_t)s.kb * BK;
        // K-major operands carry the prologue per k, MN-major ones per m / n
        if (MODE == GEMM_TN) ra.stash_fast(sp, sp + A_BYTES, s.m0, p.a_sc, p.a_sh);
        else ra.stash_fast(sp, sp + A_BYTES, k0, p.a_sc, p.a_sh);
        if (MODE == GEMM_NT) rb.stash_fast(sp + 2 * A_BYTES, sp + 2 * A_BYTES + B_BYTES, k0, nullptr, nullptr);
        else rb.stash_fast(sp + 2 * A_BYTES, sp + 2 * A_BYTES + B_BYTES, s.n0, MODE == GEMM_TN ? p.b_sc : nullptr, p.b_sh);
      } else {
        ra.stash(sp, sp + A_BYTES);
        rb.stash(sp + 2 * A_BYTES, sp + 2 * A_BYTES + B_BYTES);
      }
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&bar_full[st]));
      fetch(f, ra, rb);
      if (++st == NST) { st = 0; ++use; }
      next(s);
    };
    while (s.tile < total) {
      step(la[0], lb[0]);
      if (s.tile >= total) break;
      step(la[1], lb[1]);
    }
  } else if (warp == WS_PRODUCERS) {
    // =========================== MMA issue ===============================================================
    int st = 0;
    uint32_t use = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      Cur c;
      load_tile(tile, c);
      const int a = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&bar_acc_empty[a]), (uint32_t)(((it >> 1) - 1) & 1));
      const uint32_t d = tmem_d + (uint32_t)(a * BN);
      for (int kb = 0; kb < c.nkb; ++kb) {
        mbar_wait(smem_u32(&bar_full[st]), use & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t sa = tiles_u32 + (uint32_t)st * STAGE_BYTES;
          const uint32_t sb = sa + 2 * A_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint32_t a_off = A_MN ? (uint32_t)ks * A_KSTEP : (uint32_t)ks * 32u;
            const uint32_t b_off = B_MN ? (uint32_t)ks * B_KSTEP : (uint32_t)ks * 32u;
            const uint64_t a_hi = A_MN ? make_desc(sa + a_off, A_LBO, A_SBO, LAYOUT_SW128_BASE32B)
                                       : make_desc(sa + a_off, 16, 1024, LAYOUT_SW128);
            const uint64_t a_lo = A_MN ? make_desc(sa + A_BYTES + a_off, A_LBO, A_SBO, LAYOUT_SW128_BASE32B)
                                       : make_desc(sa + A_BYTES + a_off, 16, 1024, LAYOUT_SW128);
            const uint64_t b_hi = B_MN ? make_desc(sb + b_off, B_LBO, B_SBO, LAYOUT_SW128_BASE32B)
                                       : make_desc(sb + b_off, 16, 1024, LAYOUT_SW128);
            const uint64_t b_lo = B_MN ? make_desc(sb + B_BYTES + b_off, B_LBO, B_SBO, LAYOUT_SW128_BASE32B)
                                       : make_desc(sb + B_BYTES + b_off, 16, 1024, LAYOUT_SW128);
            if (p.dbg & 4) continue;
            umma_tf32(d, a_lo, b_hi, IDESC, (kb > 0 || ks > 0) ? 1u : 0u);   // small terms first
            umma_tf32(d, a_hi, b_lo, IDESC, 1u);
            umma_tf32(d, a_hi, b_hi, IDESC, 1u);
          }
          umma_commit(smem_u32(&bar_empty[st]));
          if (kb + 1 == c.nkb) umma_commit(smem_u32(&bar_acc_full[a]));
        }
        __syncwarp();
        if (++st == NST) { st = 0; ++use; }
      }
    }
  } else if (warp <= WS_PRODUCERS + 4) {
    // =========================== epilogue ================================================================
    const int q4 = warp & 3;                       // TMEM lane quarter this warp may read
    float* xp = xpose + (warp - WS_PRODUCERS - 1) * 2048;
    const uint32_t xp_u32 = tiles_u32 + NST * STAGE_BYTES + (uint32_t)(warp - WS_PRODUCERS - 1) * 8192u;
    const bool tstore = TMA && p.tma_store;
    uint32_t nbox = 0;                              // boxes this warp has handed to the TMA store engine
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      Cur c;
      load_tile(tile, c);
      const int a = it & 1;
      mbar_wait_sleep(smem_u32(&bar_acc_full[a]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const int row = c.m0 + q4 * 32 + lane;
      const bool row_ok = row < p.M;
      const int rows_w = min(32, p.M - (c.m0 + q4 * 32));     // valid rows of this warp's quarter (may be <= 0)
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        float v[32];
        tmem_ld32(tmem_d + (uint32_t)(a * BN + j * 32) + ((uint32_t)(q4 * 32) << 16), v);
        if (j == BN / 32 - 1) {
          tc_fence_before();
          mbar_arrive(smem_u32(&bar_acc_empty[a]));           // the accumulator may be overwritten by tile it + 2
        }
        const int col0 = c.n0 + j * 32;
        if (col0 >= p.N) continue;
        const int cols = min(32, p.N - col0);
        if (ksplit > 1) {
          if (row_ok) {
            float* o = p.part + ((int64_t)c.z * p.M + row) * p.N + col0;
            if ((p.N & 3) == 0 && cols == 32) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < cols) o[i] = v[i];
            }
          }
          continue;
        }
        if (p.bias) {
          if (cols == 32 && ((reinterpret_cast<uintptr_t>(p.bias + col0) & 15u) == 0)) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + i));
              v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < cols) v[i] += __ldg(p.bias + col0 + i);
          }
        }
        if (tstore) {
          // registers -> swizzled 32 x 32 box (row = lane: 8 conflict-free 16-byte stores) -> one TMA store per box; the
          // store engine clips rows >= M and columns >= N.  Two boxes per warp alternate, so the store of box j drains
          // while box j + 1 is filled.
          uint8_t* box = reinterpret_cast<uint8_t*>(xp) + (nbox & 1u) * 4096u;
          if (nbox >= 2) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(box + lane * 128 + ((i ^ (lane & 7)) << 4)) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (p.stat_part) {
            float sm_ = 0.f, ss_ = 0.f;
            for (int r = 0; r < rows_w; ++r) {
              const float z = *reinterpret_cast<const float*>(box + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
              sm_ += z;
              ss_ = fmaf(z, z, ss_);
            }
            if (lane < cols) {
              const int64_t pi = (int64_t)(c.m0 / BM) * 4 + q4;
              p.stat_part[(pi * 2 + 0) * p.N + col0 + lane] = sm_;
              p.stat_part[(pi * 2 + 1) * p.N + col0 + lane] = ss_;
            }
          }
          if (lane == 0 && rows_w > 0 && !(p.dbg & 2)) {
            tma_store_2d(&maps.c, xp_u32 + (nbox & 1u) * 4096u, col0, c.m0 + q4 * 32);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++nbox;
          continue;
        }
        if (p.stat_part) {      // column statistics of z = acc + bias (before any accumulate: stats GEMMs never accumulate)
#pragma unroll
          for (int i = 0; i < 32; ++i) xp[lane * 33 + i] = v[i];
          __syncwarp();
          float sm_ = 0.f, ss_ = 0.f;
          for (int r = 0; r < rows_w; ++r) {
            const float z = xp[r * 33 + lane];
            sm_ += z;
            ss_ = fmaf(z, z, ss_);
          }
          if (lane < cols) {
            const int64_t pi = (int64_t)(c.m0 / BM) * 4 + q4;
            p.stat_part[(pi * 2 + 0) * p.N + col0 + lane] = sm_;
            p.stat_part[(pi * 2 + 1) * p.N + col0 + lane] = ss_;
          }
          __syncwarp();
        }
        if (row_ok && !(p.dbg & 2)) {
          float* o = p.C + (int64_t)row * p.ldc + col0;
          if (p.c_vec && cols == 32) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 w = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
              if (p.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(o + i);
                w.x += old.x; w.y += old.y; w.z += old.z; w.w += old.w;
              }
              *reinterpret_cast<float4*>(o + i) = w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < cols) o[i] = p.accumulate ? o[i] + v[i] : v[i];
          }
        }
      }
    }
    if (tstore && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every box has reached memory
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WS_PRODUCERS) tmem_dealloc(tmem_d, 2 * BN);
}

template <int MODE, int BN, bool FAST, bool TMA>
static cudaError_t launch_ws2(const Params& p, int gn, int gm, int ksplit, const TmaMaps& maps, cudaStream_t st) {
  constexpr size_t smem = WsCfg<BN>::SMEM;
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(k_tc_gemm_ws<MODE, BN, FAST, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int total = gn * gm * ksplit;
  const int grid = total < kNumSMs ? total : kNumSMs;
  k_tc_gemm_ws<MODE, BN, FAST, TMA><<<grid, TMA ? WS_THREADS_TMA : WS_THREADS, smem, st>>>(p, gn, gm, ksplit, maps);
  return cudaSuccess;
}

// ---- tensor maps ----------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is looked up through the runtime so the library keeps linking
// against cudart only.  One 2-D fp32 map per operand: dim0 = the contiguous index, dim1 = the strided one.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
// `base[outer * ld + inner]`, inner in [0, n_inner), outer in [0, n_outer); box = 32 inner x box_outer.
static bool encode_map(CUtensorMap* m, const float* base, int64_t ld, int64_t n_inner, int64_t n_outer, int box_outer,
                       bool mn_major) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)n_inner, (cuuint64_t)n_outer};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4u};
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// YOLAT_TC_TMA: unset / "all" = every mode, "nt" | "nn" | "tn" (comma-separated) = those modes only, "0" = off.
static bool tma_mode_enabled(int mode) {
  static const int mask = [] {
    const char* e = getenv("YOLAT_TC_TMA");
    if (!e || !e[0] || e[0] == 'a') return 7;
    if (e[0] == '0') return 0;
    int m = 0;
    for (const char* c = e; *c; ++c) {
      if (c[0] == 'n' && c[1] == 't') m |= 1 << GEMM_NT;
      if (c[0] == 'n' && c[1] == 'n') m |= 1 << GEMM_NN;
      if (c[0] == 't' && c[1] == 'n') m |= 1 << GEMM_TN;
    }
    return m;
  }();
  return (mask >> mode) & 1;
}

template <int MODE, int BN>
static cudaError_t launch_ws(const Params& p, int gn, int gm, int ksplit, cudaStream_t st) {
  // fast loaders: aligned vector loads on both operands and no partial 16-byte chunk along the contiguous dimension
  const bool a_k = (MODE != GEMM_TN), b_k = (MODE == GEMM_NT);      // operand contiguous along K?
  TmaMaps maps{};
  // TMA-fed variant: no operand prologue, 16-byte aligned bases and row pitches (the box handles every ragged edge)
  if (tma_mode_enabled(MODE) && p.a_vec && p.b_vec && !p.a_sc && !p.b_sc && p.K < (1ll << 31)) {
    const bool ok_a = a_k ? encode_map(&maps.a, p.A, p.lda, p.K, p.M, BM, false) : encode_map(&maps.a, p.A, p.lda, p.M, p.K, 32, true);
    const bool ok_b = b_k ? encode_map(&maps.b, p.B, p.ldb, p.K, p.N, BN, false) : encode_map(&maps.b, p.B, p.ldb, p.N, p.K, 32, true);
    if (ok_a && ok_b) {
      Params q = p;
      static const bool store_on = !(getenv("YOLAT_TC_TMA_STORE") && getenv("YOLAT_TC_TMA_STORE")[0] == '0');
      // C through TMA stores: whole-K tiles written once (no split-K partials, no read-modify-write)
      q.tma_store = store_on && ksplit == 1 && !p.accumulate && p.c_vec && encode_map(&maps.c, p.C, p.ldc, p.N, p.M, 32, false);
      return launch_ws2<MODE, BN, true, true>(q, gn, gm, ksplit, maps, st);
    }
  }
  const bool fast = p.a_vec && p.b_vec && ((a_k ? p.K : (int64_t)p.M) % 4 == 0) && ((b_k ? p.K : (int64_t)p.N) % 4 == 0);
  return fast ? launch_ws2<MODE, BN, true, false>(p, gn, gm, ksplit, maps, st)
              : launch_ws2<MODE, BN, false, false>(p, gn, gm, ksplit, maps, st);
}

template <int MODE, int BN>
static cudaError_t launch(const Params& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (2 * A_BYTES + 2 * BN * 128) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_tc_gemm<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_tc_gemm<MODE, BN><<<grid, THREADS, smem, st>>>(p);
  return cudaSuccess;
}

}  // namespace tc

__global__ void k_splitk_sum(const float* __restrict__ part, int ksplit, int M, int N, const float* __restrict__ bias,
                             float* __restrict__ C, int64_t ldc, int accumulate);
__global__ void k_splitk_sum_deep(const float* __restrict__ part, int ksplit, int M, int N, const float* __restrict__ bias,
                                  float* __restrict__ C, int64_t ldc, int accumulate);

static bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- skinny dW ------------------------------------------------------------------------------------------------------
// dW = dy^T x with a handful of input columns (the head GraphConv reads the 5 raw curve features): a 128 x 64 tensor-core
// tile would be 92 % padding.  One thread per output row m keeps the N <= 8 accumulators of its row; a CTA walks
// SK_ROWS reduction rows (dy row segment coalesced across the threads, the x row is a broadcast) and writes one partial
// [M, N]; the partials are summed in fixed order by k_splitk_sum_deep.
constexpr int SK_ROWS = 128, SK_MAXN = 8, SK_GROUPS = 4;
__global__ void __launch_bounds__(64 * SK_GROUPS) k_tn_skinny(const float* __restrict__ A, int64_t lda,
                                                              const float* __restrict__ B, int64_t ldb, int M, int N,
                                                              int64_t K, float* __restrict__ part) {
  const int mi = threadIdx.x & 63, grp = threadIdx.x >> 6;          // 64 output rows per CTA, SK_GROUPS row lanes
  const int m = blockIdx.y * 64 + mi;
  const int64_t k0 = (int64_t)blockIdx.x * SK_ROWS;
  const int rows = (int)min((int64_t)SK_ROWS, K - k0);
  __shared__ float xs[SK_ROWS * SK_MAXN];
  __shared__ float red[SK_GROUPS][64][SK_MAXN + 1];
  for (int i = threadIdx.x; i < SK_ROWS * SK_MAXN; i += 64 * SK_GROUPS) {
    const int r = i / SK_MAXN, j = i % SK_MAXN;
    xs[i] = (r < rows && j < N) ? B[(k0 + r) * ldb + j] : 0.f;
  }
  __syncthreads();
  float acc[SK_MAXN];
#pragma unroll
  for (int j = 0; j < SK_MAXN; ++j) acc[j] = 0.f;
  const int mc = m < M ? m : M - 1;
#pragma unroll 1
  for (int i0 = 0; i0 < SK_ROWS / SK_GROUPS; i0 += 8) {
    float a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {                                    // eight independent loads in flight
      const int r = (i0 + u) * SK_GROUPS + grp;
      a[u] = r < rows ? A[(k0 + r) * lda + mc] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float* x = xs + ((i0 + u) * SK_GROUPS + grp) * SK_MAXN;
#pragma unroll
      for (int j = 0; j < SK_MAXN; ++j) acc[j] = fmaf(a[u], x[j], acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < SK_MAXN; ++j) red[grp][mi][j] = acc[j];
  __syncthreads();
  if (grp != 0 || m >= M) return;
  float* o = part + ((int64_t)blockIdx.x * M + m) * N;
#pragma unroll
  for (int j = 0; j < SK_MAXN; ++j) {
    if (j < N) {
      float v = red[0][mi][j];
#pragma unroll
      for (int g2 = 1; g2 < SK_GROUPS; ++g2) v += red[g2][mi][j];
      o[j] = v;
    }
  }
}

// ---- skinny reduction: y = x W^T + b with K <= 8 (the head GraphConv's Linear layers read 5 input features) ------------
// A 128 x 64 x 32 tensor tile would be 84 % padding along K and its operands cannot be TMA-loaded (20-byte rows).  SIMT:
// thread = 4 output columns of one row, W^T staged in shared memory as [K][N] (conflict-free float4 reads), the x row
// is a broadcast.  A CTA walks SKN_ROWS rows and, on request, leaves the column statistics of its rows as one partial
// [2][N] -- the same contract as the tensor-core epilogue.
constexpr int SKN_ROWS = 128;
__global__ void __launch_bounds__(256) k_nt_skinny(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                                                   int64_t ldb, const float* __restrict__ bias, float* __restrict__ C,
                                                   int64_t ldc, int M, int N, int K, int accumulate, int c_vec,
                                                   float* __restrict__ stat_part) {
  __shared__ __align__(16) float wt[SK_MAXN * 256];       // [K][N]
  __shared__ __align__(16) float red[2 * 1024];           // [2][lanes][N], lanes * N = 1024
  const int tid = threadIdx.x;
  const int cgs = N >> 2, lanes = 256 / cgs;
  const int cg = tid % cgs, rl = tid / cgs;
  const int c = cg * 4;
  for (int i = tid; i < N * K; i += 256) wt[(i % K) * N + i / K] = B[(int64_t)(i / K) * ldb + (i % K)];
  __syncthreads();
  const int r0 = blockIdx.x * SKN_ROWS, r1 = min(M, r0 + SKN_ROWS);
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) b4 = make_float4(bias[c], bias[c + 1], bias[c + 2], bias[c + 3]);
  float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), q4 = s4;
  if (rl < lanes) {
    for (int r = r0 + rl; r < r1; r += lanes) {
      float4 v = b4;
      const float* x = A + (int64_t)r * lda;
#pragma unroll
      for (int k = 0; k < SK_MAXN; ++k) {
        if (k < K) {
          const float xv = __ldg(x + k);
          const float4 w = *reinterpret_cast<const float4*>(wt + k * N + c);
          v.x = fmaf(xv, w.x, v.x); v.y = fmaf(xv, w.y, v.y); v.z = fmaf(xv, w.z, v.z); v.w = fmaf(xv, w.w, v.w);
        }
      }
      s4.x += v.x; s4.y += v.y; s4.z += v.z; s4.w += v.w;
      q4.x = fmaf(v.x, v.x, q4.x); q4.y = fmaf(v.y, v.y, q4.y); q4.z = fmaf(v.z, v.z, q4.z); q4.w = fmaf(v.w, v.w, q4.w);
      float* o = C + (int64_t)r * ldc + c;
      if (c_vec) {
        if (accumulate) { const float4 old = *reinterpret_cast<const float4*>(o); v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
        *reinterpret_cast<float4*>(o) = v;
      } else {
        o[0] = accumulate ? o[0] + v.x : v.x; o[1] = accumulate ? o[1] + v.y : v.y;
        o[2] = accumulate ? o[2] + v.z : v.z; o[3] = accumulate ? o[3] + v.w : v.w;
      }
    }
  }
  if (!stat_part) return;
  if (rl < lanes) {
    *reinterpret_cast<float4*>(red + (0 * lanes + rl) * N + c) = s4;
    *reinterpret_cast<float4*>(red + (1 * lanes + rl) * N + c) = q4;
  }
  __syncthreads();
  if (tid < N) {
    float s = 0.f, q = 0.f;
    for (int l = 0; l < lanes; ++l) { s += red[(0 * lanes + l) * N + tid]; q += red[(1 * lanes + l) * N + tid]; }
    stat_part[((int64_t)blockIdx.x * 2 + 0) * N + tid] = s;
    stat_part[((int64_t)blockIdx.x * 2 + 1) * N + tid] = q;
  }
}

// Planning shared by the dry (workspace query) and the real call.
struct TcPlan { int bn, gm, gn, ksplit; int64_t k_chunk; };
static TcPlan tc_plan(int M, int N, int64_t K, GemmMode mode, bool want_stats) {
  TcPlan pl;
  static const int force_bn = getenv("YOLAT_TC_BN") ? atoi(getenv("YOLAT_TC_BN")) : 0;   // tuning knob
  pl.bn = force_bn == 64 ? 64 : (N > 64 ? 128 : 64);
  pl.gm = (int)cdiv(M, tc::BM);
  pl.gn = (int)cdiv(N, pl.bn);
  const int64_t tiles = (int64_t)pl.gm * pl.gn;
  const int64_t nkb = cdiv(K > 0 ? K : 1, tc::BK);
  // Split-K by a small cost model (microseconds): waves x (fixed CTA cost + k-blocks x per-block cost) + the
  // reduction pass.  Small GEMMs on 148 SMs are dominated by wave quantisation, not by flops.
  const int64_t slots = (int64_t)kNumSMs * (pl.bn == 64 ? 2 : 1);
  // the fp32 TMEM accumulator is rounded once per MMA: keep every split's reduction short so the rounding
  // stays far below the 1e-4 parity bar on the long row reductions of dW = dy^T x.
  const int64_t max_chunk_kb = (mode == GEMM_TN) ? 64 : 128;
  const int64_t s_min = cdiv(nkb, max_chunk_kb);
  int64_t s_max = nkb / 2 > 0 ? nkb / 2 : 1;       // at least 2 k-blocks per split
  if (s_max > 1024) s_max = 1024;
  if (s_max < s_min) s_max = s_min;
  const double mn_bytes = 4.0 * (double)M * (double)N;
  double best = 1e30;
  int64_t best_s = s_min;
  for (int64_t s = s_min; s <= s_max; s = (s < 16 ? s + 1 : s + s / 8)) {
    const double waves = (double)cdiv(tiles * s, slots);
    const double per_cta = 3.0 + 0.8 * (double)cdiv(nkb, s);
    double cost = waves * per_cta;
    if (s > 1) cost += 4.0 + (double)(s + 1) * mn_bytes / 3.0e6;   // ~3 TB/s through L2
    if (s > 1 && want_stats) cost += 12.0;                        // separate column-statistics pass
    if (cost < best) { best = cost; best_s = s; }
  }
  pl.k_chunk = align_up(cdiv(K > 0 ? K : 1, best_s), tc::BK);
  pl.ksplit = (int)cdiv(K > 0 ? K : 1, pl.k_chunk);
  return pl;
}

static int gemm_tc(const GemmArgs& a, GemmMode mode, Arena& ws, float* stat_part, int* stat_nparts, cudaStream_t st) {
  static const bool skinny_on = !(getenv("YOLAT_TC_SKINNY") && getenv("YOLAT_TC_SKINNY")[0] == '0');
  if (skinny_on && mode == GEMM_TN && a.N <= SK_MAXN && a.N > 0 && a.K > 0 && !a.a_sc && !a.b_sc && !stat_nparts) {
    const int nsplit = (int)cdiv(a.K, SK_ROWS);
    float* part = ws.take((int64_t)nsplit * a.M * a.N);
    if (ws.dry()) return YOLAT_OK;
    if (ws.overflow) return YOLAT_ERR_WORKSPACE;
    k_tn_skinny<<<dim3((unsigned)nsplit, (unsigned)cdiv(a.M, 64)), 64 * SK_GROUPS, 0, st>>>(a.A, a.lda, a.B, a.ldb, a.M, a.N, a.K, part);
    YOLAT_CHECK_LAUNCH();
    k_splitk_sum_deep<<<(unsigned)cdiv((int64_t)a.M * a.N, 32), dim3(32, 32), 0, st>>>(part, nsplit, a.M, a.N, a.bias, a.C, a.ldc,
                                                                                  a.accumulate);
    YOLAT_CHECK_LAUNCH();
    return YOLAT_OK;
  }
  if (skinny_on && mode == GEMM_NT && a.K > 0 && a.K <= SK_MAXN && a.N >= 16 && a.N <= 256 && a.N % 4 == 0 && !a.a_sc &&
      !a.b_sc && a.M > 0) {
    const int nb = (int)cdiv(a.M, SKN_ROWS);
    if (stat_nparts) *stat_nparts = nb;                 // one partial per CTA (fits the cdiv(M, 128) * 4 the caller reserved)
    if (ws.dry()) return YOLAT_OK;
    const int c_vec = aligned16p(a.C) && (a.ldc % 4 == 0);
    k_nt_skinny<<<nb, 256, 0, st>>>(a.A, a.lda, a.B, a.ldb, a.bias, a.C, a.ldc, a.M, a.N, (int)a.K, a.accumulate, c_vec, stat_part);
    YOLAT_CHECK_LAUNCH();
    return YOLAT_OK;
  }
  const TcPlan pl = tc_plan(a.M, a.N, a.K, mode, stat_nparts != nullptr);
  // Wave quantisation: one CTA per SM walks the tile list, so 157 row tiles cost two full passes for 1.06 passes of work
  // (dx = dz W of the fusion block: M = 20000, N = 128).  When a short tail spills over the last full wave, the rows of
  // the full waves run as one call and the tail rows as a second, much smaller problem that the planner splits along K.
  // Row-disjoint outputs: bias / accumulate keep their meaning; not for dW (rows are the reduction) nor when the
  // epilogue has to deliver column statistics of the whole matrix.
  static const bool tail_split = !(getenv("YOLAT_TC_TAIL") && getenv("YOLAT_TC_TAIL")[0] == '0');
  if (tail_split && mode != GEMM_TN && !stat_nparts && pl.ksplit == 1 && a.K >= 32 * tc::BK) {   // long reductions only
    const int64_t tiles = (int64_t)pl.gm * pl.gn;
    const int64_t full = tiles / kNumSMs * kNumSMs;
    const int gm_main = (int)(full / pl.gn);
    if (full > 0 && tiles - full > 0 && (tiles - full) * 8 <= kNumSMs && gm_main > 0 && gm_main < pl.gm) {
      GemmArgs head = a, tail = a;
      head.M = gm_main * tc::BM;
      tail.M = a.M - head.M;
      tail.A = a.A ? a.A + (int64_t)head.M * a.lda : nullptr;
      tail.C = a.C ? a.C + (int64_t)head.M * a.ldc : nullptr;
      YOLAT_TRY(gemm_tc(head, mode, ws, nullptr, nullptr, st));
      return gemm_tc(tail, mode, ws, nullptr, nullptr, st);
    }
  }
  float* part = nullptr;
  if (pl.ksplit > 1) part = ws.take((int64_t)pl.ksplit * a.M * a.N);
  if (stat_nparts) *stat_nparts = pl.ksplit > 1 ? 0 : pl.gm * 4;    // one partial per epilogue warp (32 rows)
  if (ws.dry()) return YOLAT_OK;
  if (ws.overflow) return YOLAT_ERR_WORKSPACE;
  tc::Params p{};
  p.A = a.A; p.lda = a.lda; p.B = a.B; p.ldb = a.ldb; p.C = a.C; p.ldc = a.ldc;
  p.M = a.M; p.N = a.N; p.K = a.K; p.bias = a.bias;
  p.a_sc = a.a_sc; p.a_sh = a.a_sh; p.b_sc = a.b_sc; p.b_sh = a.b_sh;
  p.accumulate = a.accumulate;
  p.k_chunk = pl.k_chunk;
  p.part = part;
  p.stat_part = pl.ksplit > 1 ? nullptr : stat_part;
  p.a_vec = aligned16p(a.A) && (a.lda % 4 == 0);
  p.b_vec = aligned16p(a.B) && (a.ldb % 4 == 0);
  p.c_vec = aligned16p(a.C) && (a.ldc % 4 == 0);
  static const int dbg = getenv("YOLAT_TC_DBG") ? atoi(getenv("YOLAT_TC_DBG")) : 0;
  p.dbg = dbg;
  dim3 grid(pl.gn, pl.gm, pl.ksplit);
  cudaError_t e = cudaSuccess;
  ProfScope prof(YOLAT_PROF_GEMM, st);
  static const bool use_ws = !(getenv("YOLAT_TC_KERNEL") && getenv("YOLAT_TC_KERNEL")[0] == 'o');   // 'old' = one tile per CTA
#define YOLAT_TC_CASE(MODE_)                                                    \
  case MODE_:                                                                   \
    if (use_ws && a.K > 0)                                                      \
      e = pl.bn == 128 ? tc::launch_ws<MODE_, 128>(p, pl.gn, pl.gm, pl.ksplit, st) \
                       : tc::launch_ws<MODE_, 64>(p, pl.gn, pl.gm, pl.ksplit, st); \
    else                                                                        \
      e = pl.bn == 128 ? tc::launch<MODE_, 128>(p, grid, st) : tc::launch<MODE_, 64>(p, grid, st); \
    break;
  switch (mode) {
    YOLAT_TC_CASE(GEMM_NT)
    YOLAT_TC_CASE(GEMM_NN)
    YOLAT_TC_CASE(GEMM_TN)
  }
#undef YOLAT_TC_CASE
  if (e != cudaSuccess) { set_last_error(e); return YOLAT_ERR_LAUNCH; }
  YOLAT_CHECK_LAUNCH();
  if (pl.ksplit > 1) {
    const int64_t tot = (int64_t)a.M * a.N;
    if (pl.ksplit >= 16 && tot * 4 <= (int64_t)pl.ksplit * 16384) {   // few outputs per split: one block per 32 outputs
      k_splitk_sum_deep<<<(unsigned)cdiv(tot, 32), dim3(32, 32), 0, st>>>(part, pl.ksplit, a.M, a.N, a.bias, a.C, a.ldc,
                                                                          a.accumulate);
    } else {
      k_splitk_sum<<<(unsigned)cdiv(tot, 256), 256, 0, st>>>(part, pl.ksplit, a.M, a.N, a.bias, a.C, a.ldc, a.accumulate);
    }
    YOLAT_CHECK_LAUNCH();
  }
  return YOLAT_OK;
}

// C = sum_z part[z] (+ bias), fixed summation order (deterministic).  Wide variant: many output elements, few splits.
__global__ void k_splitk_sum(const float* __restrict__ part, int ksplit, int M, int N,
                             const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int accumulate) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * N) return;
  const int m = (int)(idx / N), n = (int)(idx % N);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int z = 0;
  const int64_t stride = (int64_t)M * N;
  for (; z + 3 < ksplit; z += 4) {
    s0 += part[(int64_t)z * stride + idx];
    s1 += part[(int64_t)(z + 1) * stride + idx];
    s2 += part[(int64_t)(z + 2) * stride + idx];
    s3 += part[(int64_t)(z + 3) * stride + idx];
  }
  for (; z < ksplit; ++z) s0 += part[(int64_t)z * stride + idx];
  float s = (s0 + s1) + (s2 + s3);
  if (bias) s += bias[n];
  float* c = C + (int64_t)m * ldc + n;
  *c = accumulate ? (*c + s) : s;
}

// Deep variant: few output elements, hundreds of splits (dW of the edge / node layers).  32 x 32 block:
// tx = output element (coalesced), ty strides over the splits with 4 loads in flight, fp64 combine.
__global__ void __launch_bounds__(1024) k_splitk_sum_deep(const float* __restrict__ part, int ksplit, int M, int N,
                                                          const float* __restrict__ bias, float* __restrict__ C,
                                                          int64_t ldc, int accumulate) {
  __shared__ double sm[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t idx = blockIdx.x * 32 + (int64_t)tx;
  const int64_t stride = (int64_t)M * N;
  double s = 0.0;
  if (idx < stride) {
    double s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int z = ty;
    for (; z + 96 < ksplit; z += 128) {
      const float x0 = part[(int64_t)z * stride + idx], x1 = part[(int64_t)(z + 32) * stride + idx];
      const float x2 = part[(int64_t)(z + 64) * stride + idx], x3 = part[(int64_t)(z + 96) * stride + idx];
      s += (double)x0; s1 += (double)x1; s2 += (double)x2; s3 += (double)x3;
    }
    for (; z < ksplit; z += 32) s += (double)part[(int64_t)z * stride + idx];
    s = (s + s1) + (s2 + s3);
  }
  sm[ty][tx] = s;
  __syncthreads();
  if (ty != 0 || idx >= stride) return;
#pragma unroll 4
  for (int q = 1; q < 32; ++q) s += sm[q][tx];
  const int m = (int)(idx / N), n = (int)(idx % N);
  float v = (float)s;
  if (bias) v += bias[n];
  float* c = C + (int64_t)m * ldc + n;
  *c = accumulate ? (*c + v) : v;
}

static bool gemm_use_tc();

int gemm_stats(const GemmArgs& a, GemmMode mode, Arena& ws, float** stat_part, int* nparts, cudaStream_t st) {
  if (a.M <= 0 || a.N <= 0) { if (nparts) *nparts = 0; return YOLAT_OK; }
  if (!gemm_use_tc()) {
    if (nparts) *nparts = 0;
    return gemm_simt(a, mode, ws, st);
  }
  float* sp = nullptr;
  if (stat_part) {
    sp = ws.take((int64_t)cdiv(a.M, tc::BM) * 4 * 2 * a.N);
    *stat_part = sp;
  }
  return gemm_tc(a, mode, ws, sp, nparts, st);
}

int gemm(const GemmArgs& a, GemmMode mode, Arena& ws, cudaStream_t st) {
  return gemm_stats(a, mode, ws, nullptr, nullptr, st);
}

static bool gemm_use_tc() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YOLAT_GEMM");
    v = (e && e[0] == 's') ? 0 : 1;   // YOLAT_GEMM=simt selects the exact-fp32 SIMT kernel (debugging aid)
  }
  return v == 1;
}

}  // namespace yolat
