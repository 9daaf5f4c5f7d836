// tc.cuh -- tcgen05 / TMEM / mbarrier inline-PTX helpers shared by the tensor-core kernels (sm_100a only).
// Bit layouts follow the CUTLASS sm_100 descriptors (cute/arch/mma_sm100_desc.hpp): shared-memory matrix
// descriptor (version 1) and the kind::tf32 instruction descriptor.
#pragma once
#include "common.cuh"

namespace yolat {
namespace tc {

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a descriptor / protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  if (mbar_try_wait(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(addr, parity)) {
    if (clock64() - t0 > 4000000000LL) asm volatile("trap;");
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
// named barrier among `nthreads` threads (multiple of 32) of the CTA; id 0 is __syncthreads
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Long waits (a whole pipeline stage): back off with nanosleep so the spinning warp does not steal issue slots
// from the warps doing the work.  Same bounded-time trap as mbar_wait.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t addr, uint32_t parity) {
  if (mbar_try_wait(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(addr, parity)) {
    __nanosleep(40);
    if (clock64() - t0 > 4000000000LL) asm volatile("trap;");
  }
}
// Waits that leave the issue slots to the working warps: try_wait with a suspend-time hint parks the warp in hardware
// until the phase completes (wake-up ~60 cycles after the arrive) or the hint expires.  Bounded like mbar_wait.
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t addr, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(addr), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait_park(uint32_t addr, uint32_t parity) {
  if (mbar_try_wait(addr, parity)) return;
  for (int spins = 0; !mbar_try_wait_hint(addr, parity, 1000000u); ++spins) {
    if (spins > 200000000) asm volatile("trap;");
  }
}
// Same parking wait with a wall-clock bound (~2 s of SM cycles): a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_bounded(uint32_t addr, uint32_t parity) {
  if (mbar_try_wait(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_hint(addr, parity, 20000u)) {
    if (clock64() - t0 > 4000000000LL) asm volatile("trap;");
  }
}
// ---- TMA bulk copy (cp.async.bulk, 1-D): global -> shared, completion counted in bytes on an mbarrier -------------
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
// dst / src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
// ---- TMA tensor copy (cp.async.bulk.tensor, 2-D tiled): one box of a tensor map -> shared, bytes counted on an mbarrier.
// The box lands in the swizzle mode the map was encoded with; elements outside the tensor arrive as zeros and still count.
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void* tmap, int32_t c0, int32_t c1, uint32_t mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst_smem), "l"(tmap), "r"(mbar), "r"(c0), "r"(c1)
               : "memory");
}
// shared -> global: one box of the tensor map, clipped to the tensor's extents; completion through bulk async-groups
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src_smem, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// one lane of a fully converged warp; keeps the guarded region in the uniform datapath (descriptors in UR registers)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- kind::f16 with bf16 operands (fp32 accumulate): the "bf16x3" products of the fused K-EDGE backward ------------
// 16-bit operands use the plain SWIZZLE_128B layout for K-major AND MN-major reads, so ONE shared-memory tile
// [row][64 bf16] serves a product that contracts over its columns (K-major) and one that contracts over its rows
// (MN-major) -- verified on hardware by tools/umma_probe.cu (profiles/r2_a_umma_probe.txt).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// two fp32 -> packed bf16x2 (round to nearest even), low half = a
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// Two-term bf16 split of four fp32 values: hi = rn_bf16(v), lo = rn_bf16(v - hi); v - hi - lo is below 2^-17 |v| and of
// either sign, so hi*hi + hi*lo + lo*hi reproduces an fp32 product to ~2^-16 with unbiased rounding.
__device__ __forceinline__ void split_bf16x4(const float4& v, uint2& hi, uint2& lo) {
  hi.x = pack_bf16x2(v.x, v.y);
  hi.y = pack_bf16x2(v.z, v.w);
  const float rx = v.x - __uint_as_float(hi.x << 16), ry = v.y - __uint_as_float(hi.x & 0xffff0000u);
  const float rz = v.z - __uint_as_float(hi.y << 16), rw = v.w - __uint_as_float(hi.y & 0xffff0000u);
  lo.x = pack_bf16x2(rx, ry);
  lo.y = pack_bf16x2(rz, rw);
}

// Three-term split: v = hi + mid + lo up to 2^-25 |v|.
__device__ __forceinline__ void split3_bf16x4(const float4& v, uint2& hi, uint2& mid, uint2& lo) {
  hi.x = pack_bf16x2(v.x, v.y);
  hi.y = pack_bf16x2(v.z, v.w);
  const float rx = v.x - __uint_as_float(hi.x << 16), ry = v.y - __uint_as_float(hi.x & 0xffff0000u);
  const float rz = v.z - __uint_as_float(hi.y << 16), rw = v.w - __uint_as_float(hi.y & 0xffff0000u);
  mid.x = pack_bf16x2(rx, ry);
  mid.y = pack_bf16x2(rz, rw);
  lo.x = pack_bf16x2(rx - __uint_as_float(mid.x << 16), ry - __uint_as_float(mid.x & 0xffff0000u));
  lo.y = pack_bf16x2(rz - __uint_as_float(mid.y << 16), rw - __uint_as_float(mid.y & 0xffff0000u));
}

// Shared-memory matrix descriptor, sm_100 version field = 1
// (bit layout: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version [46,48) | layout [61,64)).
constexpr uint32_t LAYOUT_SW128 = 2;          // K-major operands
constexpr uint32_t LAYOUT_SW128_BASE32B = 1;  // MN-major 32-bit operands
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// Instruction descriptor for kind::tf32, fp32 accumulate
// (c_format [4,6)=1 F32 | a_format [7,10)=2 TF32 | b_format [10,13)=2 | a_major 15 | b_major 16 | N>>3 [17,23) | M>>4 [24,29)).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}
// Same split for finite inputs in 2 + 1 instructions per value (cvt.rna.tf32 compiles to 4): adding half an ulp of
// the 10-bit mantissa and masking the low 13 bits is round-to-nearest, ties away from zero, on the magnitude.
__device__ __forceinline__ void split_tf32_fast(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}
// Finite inputs, 2.5 instructions per value: hi = x rounded to nearest at 10 mantissa bits (add half an ulp of the kept
// field to the bit pattern, clear the low 13 bits: ties away from zero on the magnitude), lo = x - hi exactly (either
// sign, <= 2^-12 |x|) through the packed FADD2.  The tensor core reads lo truncated to tf32: what 3xTF32 then drops is
// <= 2^-23 |a b| per product and sign-symmetric.  (Round 1 truncated hi: one instruction cheaper, but lo was one-sided and
// its truncation a systematic 2^-21 bias -- enough to flip ~3x more ReLU masks / arg-max picks against an fp64 run than
// an fp32 product does, tools/e2e_noise.py.)
__device__ __forceinline__ void store_split_fast(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, const float4& v) {
  float4 h;
  h.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xffffe000u);
  h.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xffffe000u);
  h.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xffffe000u);
  h.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xffffe000u);
  float2 l0, l1;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(l0.x), "=f"(l0.y) : "f"(v.x), "f"(v.y), "f"(h.x), "f"(h.y));
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(l1.x), "=f"(l1.y) : "f"(v.z), "f"(v.w), "f"(h.z), "f"(h.w));
  *reinterpret_cast<float4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = make_float4(l0.x, l0.y, l1.x, l1.y);
}
__device__ __forceinline__ void store_split(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, const float4& v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x);
  split_tf32(v.y, h.y, l.y);
  split_tf32(v.z, h.z, l.z);
  split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = l;
}


// ---- packed fp32 pairs (FFMA2 / FADD2 on sm_100: two fp32 lanes per instruction, each lane IEEE round-to-nearest) ----
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 ffma2s(float s, float2 b, float2 c) { return ffma2(make_float2(s, s), b, c); }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "sub.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// Round-to-nearest split of a finite value for 3xTF32: hi = x rounded at 10 mantissa bits, lo = x - hi exactly (either
// sign).  See store_split_fast.
__device__ __forceinline__ void store_split_trunc(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float2 a, float2 b) {
  float4 h;
  h.x = __uint_as_float((__float_as_uint(a.x) + 0x1000u) & 0xffffe000u);
  h.y = __uint_as_float((__float_as_uint(a.y) + 0x1000u) & 0xffffe000u);
  h.z = __uint_as_float((__float_as_uint(b.x) + 0x1000u) & 0xffffe000u);
  h.w = __uint_as_float((__float_as_uint(b.y) + 0x1000u) & 0xffffe000u);
  const float2 l0 = fsub2(a, make_float2(h.x, h.y)), l1 = fsub2(b, make_float2(h.z, h.w));
  *reinterpret_cast<float4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = make_float4(l0.x, l0.y, l1.x, l1.y);
}

}  // namespace tc
}  // namespace yolat
