// Device helpers shared by the tcgen05 kernels (tc_gemm.cu, skinny.cu): mbarriers, TMA loads,
// UMMA descriptors / issue, TMEM loads, the fp16 power-of-two scaling, coalesced staged stores.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "tnc_internal.h"

namespace tnc {
namespace {

// Power-of-two scaling of an fp16-split operand.  e = biased exponent of the operand's largest
// magnitude (clamped): scale = 2^(141 - e) maps it into [2^14, 2^15); inv_scale undoes it.
__device__ __forceinline__ uint32_t amax_exponent(uint32_t amax_bits) {
    const uint32_t e = (amax_bits >> 23) & 0xffu;
    return e < 15u ? 15u : (e > 254u ? 254u : e);
}
__device__ __forceinline__ float f16_scale(uint32_t amax_bits) { return __uint_as_float((268u - amax_exponent(amax_bits)) << 23); }
__device__ __forceinline__ float f16_inv_scale(uint32_t amax_bits) { return __uint_as_float((amax_exponent(amax_bits) - 14u) << 23); }

template <int PREC>
struct Prec;
template <>
struct Prec<TNC_TC_3XTF32> {
    static constexpr int ELEM = 4, PANELS = 2;
    static constexpr uint32_t FMT = 2;      // instruction-descriptor operand format: TF32
    static constexpr bool F16 = false;
};
template <>
struct Prec<TNC_TC_3XF16> {
    static constexpr int ELEM = 2, PANELS = 2;
    static constexpr uint32_t FMT = 0;      // F16
    static constexpr bool F16 = true;
};
template <>
struct Prec<TNC_TC_F16> {
    static constexpr int ELEM = 2, PANELS = 1;
    static constexpr uint32_t FMT = 0;
    static constexpr bool F16 = true;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// try_wait with a suspend-time hint: the warp is descheduled until the phase completes or the hint
// (nanoseconds) runs out, instead of re-issuing the poll back to back.
__device__ __forceinline__ bool mbar_try_wait_suspend(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// A wait that cannot complete is a protocol bug: trap after ~2 s instead of hanging the GPU.
// Waiting warps must not eat issue slots (the streaming kernels are issue-bound: ncu showed 20 % of
// the executed instructions in this loop when it polled back to back and read the clock on every
// iteration) nor power (the GEMM's epilogue warps wait most of the time, under a power cap).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    long long t0 = 0;
    while (!mbar_try_wait_suspend(bar, parity, 2000u)) {
        if ((++polls & 255u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// One lane of a converged warp (the warp must execute this together).  Code guarded by it is known to
// ptxas to run in a single thread, so descriptors and barrier addresses stay in uniform registers; a
// `lane == 0` branch instead makes every tcgen05 / TMA instruction pay a warp-uniformisation loop
// (ELECT + PLOP3 + BRA.U.ANY, ~30 issue slots per MMA: measured, the MMA thread became the bottleneck
// of the 3M kernel at 59 % tensor-pipe utilisation).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// warp index as a warp-uniform value (threadIdx-derived values are divergent to the compiler)
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 128-byte rows, SWIZZLE_128B: rows 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, 16-byte units      bits [0,14)
    d |= (uint64_t)1 << 16;                     // leading byte offset (unused here)  bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset = 1024 B        bits [32,46)
    d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)        bits [46,48)
    d |= (uint64_t)2 << 61;                     // SWIZZLE_128B                       bits [61,64)
    return d;
}
// instruction descriptor: D fp32, A/B in the precision's format, both K-major
template <int PREC>
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (Prec<PREC>::FMT << 7) | (Prec<PREC>::FMT << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T; CG = cta_group
template <int PREC, int CG>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    if constexpr (Prec<PREC>::F16) {
        if constexpr (CG == 1) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                : "memory");
        } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                : "memory");
        }
    } else {
        if constexpr (CG == 1) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                : "memory");
        } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                : "memory");
        }
    }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


// Coalesced store of a warp's 32 rows x (4 * P) floats (lane = row, `v` = the lane's floats) to
// global rows `pitch` floats apart, through a per-warp shared-memory buffer of 32 * P * 16 bytes:
// rows go in with one 16-byte store per piece (piece index XOR-swizzled by the row: conflict-free),
// and come out as 16-byte pieces in global-address order, so that 8 consecutive lanes write one
// full 128-byte line instead of every lane touching a line of its own.
template <int P>
__device__ __forceinline__ void store_rows_coalesced(uint32_t stage_smem, const float* v, float* gbase, int64_t pitch, int lane) {
    static_assert(P == 1 || P == 2 || P == 4 || P == 8, "pieces per row");
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const uint32_t addr = stage_smem + (uint32_t)lane * (P * 16) + (uint32_t)((p ^ (lane & (P - 1))) << 4);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * p]), "f"(v[4 * p + 1]), "f"(v[4 * p + 2]),
                     "f"(v[4 * p + 3])
                     : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int x = i * 32 + lane;          // piece number in global-address order
        const int row = x / P, p = x % P;
        const uint32_t addr = stage_smem + (uint32_t)row * (P * 16) + (uint32_t)((p ^ (row & (P - 1))) << 4);
        float4 o;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w) : "r"(addr) : "memory");
        *(float4*)(gbase + (int64_t)row * pitch + 4 * p) = o;
    }
    __syncwarp();
}

// Largest output magnitude as a by-product of a producing kernel (the consumer is a tensor-core step that scales
// its fp16 operands by it: no separate pass over the tensor).  Non-negative floats order like unsigned integers.
__device__ __forceinline__ void amax_fold(float& m, const float x) {
    m = fmaxf(m, fabsf(x));
}
// whole warp, converged
__device__ __forceinline__ void amax_commit(uint32_t* word, const float m) {
    const uint32_t bits = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
    if (bits && (threadIdx.x & 31) == 0) atomicMax(word, bits);
}

}  // namespace
}  // namespace tnc
