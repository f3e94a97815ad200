// Inline-PTX wrappers shared by the tcgen05 kernels (mbarrier, TMA, TMEM, UMMA descriptors).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <cstdio>

#include "hfg_internal.h"

namespace hfg {
namespace ptx {

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.  The bound is wall time on %globaltimer
// (20 s), not a poll count: under time-slicing / MPS a healthy kernel may be descheduled for long stretches between polls, and a
// trap is a sticky context error for the whole host process.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    for (uint32_t i = 1;; ++i) {
        if (mbar_try_wait(bar, parity)) return;
        if ((i & 0x3fffu) == 0) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 20000000000ull) break;
        }
    }
    printf("hfg umma: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
    __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Descriptor low word (start address >> 4 | LBO) and the constant high word (SBO, version, swizzle): advancing the
// operand by `bytes` (a multiple of 16) is  lo += bytes >> 4  -- one add per MMA in the issue loop.
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t row_bytes) {
    const uint32_t layout = row_bytes == 128 ? 2u : 4u;   // SWIZZLE_128B : SWIZZLE_64B
    return ((row_bytes * 8u) >> 4) | (1u << 14) | (layout << 29);
}
__device__ __forceinline__ void umma_bf16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
        : "memory");
}
// KSTEPS consecutive K=16 slices of one (tap, pass, 128-row subtile): operands advance by 32 bytes per slice.
template <int KSTEPS>
__device__ __forceinline__ void umma_ksteps(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc0) {
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k) umma_bf16_lh(d_tmem, a_lo + 2u * k, b_lo + 2u * k, hi, idesc, k == 0 ? acc0 : 1u);
}
// tcgen05.commit implies tcgen05.fence::before_thread_sync
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
template <int W> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[W]);
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (sm_100 format, version 1).
//   128-byte rows: SWIZZLE_128B, 8-row groups 1024 B apart ; 64-byte rows: SWIZZLE_64B, 512 B apart.
// The swizzle XOR acts on absolute smem address bits, so a start address advanced by whole
// rows (a conv tap) or by 32 B (a K=16 step) addresses the TMA-written tile consistently as
// long as the tile base is aligned to the swizzle pattern (1024 B here).
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : 4ull;   // SWIZZLE_128B : SWIZZLE_64B
    const uint64_t sbo = (row_bytes * 8u) >> 4;
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : v * kLreluSlope; }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}


// ---- packed fp32 pairs (FADD2 / FMUL2 on sm_100): the epilogues are issue-bound, a pair costs one instruction ----
typedef uint64_t f2;   // two fp32 in one 64-bit register pair
__device__ __forceinline__ f2 f2_pack(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 f2_bits(uint32_t lo, uint32_t hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// leaky_relu(v, 0.1) = max(v, 0.1 v) and its inverse p > 0 ? p : 10 p = min(p, 10 p): the same values as the select forms
__device__ __forceinline__ f2 f2_lrelu(f2 v) {
    float a, b, sa, sb;
    f2_unpack(v, a, b);
    f2_unpack(f2_mul(v, f2_pack(kLreluSlope, kLreluSlope)), sa, sb);
    return f2_pack(fmaxf(a, sa), fmaxf(b, sb));
}
__device__ __forceinline__ f2 f2_inv_lrelu(f2 p) {
    float a, b, ta, tb;
    f2_unpack(p, a, b);
    f2_unpack(f2_mul(p, f2_pack(1.0f / kLreluSlope, 1.0f / kLreluSlope)), ta, tb);
    return f2_pack(fminf(a, ta), fminf(b, tb));
}
// two bf16 (channels 2i, 2i+1 of one 32-bit word) -> fp32 pair
__device__ __forceinline__ f2 f2_from_bf16x2(uint32_t w) { return f2_bits(w << 16, w & 0xffff0000u); }
__device__ __forceinline__ uint32_t f2_to_bf16x2(f2 v) {
    float a, b;
    f2_unpack(v, a, b);
    return pack_bf16(a, b);
}

// ---- 16-bit operand format of the single-plane modes: bf16 (8-bit significand) or fp16 (11-bit significand; HFG_PREC_FP16) ----
// fp16 stores saturate to +-65504 instead of overflowing to inf (F2FP.SATFINITE): an activation that large is outside anything a
// vocoder produces, but an inf would turn into NaN in the next MMA.
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
template <bool F16> __device__ __forceinline__ uint32_t pack2(float a, float b) { return F16 ? pack_f16(a, b) : pack_bf16(a, b); }
template <bool F16> __device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b) {
    if (F16) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        a = f.x; b = f.y;
    } else {
        a = __uint_as_float(w << 16); b = __uint_as_float(w & 0xffff0000u);
    }
}
template <bool F16> __device__ __forceinline__ f2 f2_from_h2(uint32_t w) {
    float a, b;
    unpack2<F16>(w, a, b);
    return f2_pack(a, b);
}
template <bool F16> __device__ __forceinline__ uint32_t f2_to_h2(f2 v) {
    float a, b;
    f2_unpack(v, a, b);
    return pack2<F16>(a, b);
}
// tcgen05 kind::f16 instruction descriptor: D = f32 (bit 4), A / B format at bits 7 / 10 (0 = f16, 1 = bf16), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t make_idesc(uint32_t n, bool f16) {
    return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---- programmatic dependent launch: the prologue (barriers, TMEM, weights) of kernel N+1 overlaps the tail of kernel N ----
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- TMA stores / bulk groups ---------------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace ptx
}  // namespace hfg
