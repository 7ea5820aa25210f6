// Streaming tensor-core kernel for the "stem" steps of a contraction tree
// (artensor/contraction.py:70, :147-190 call sites): a huge left operand meets a small right
// operand (K <= 64, N <= 128 complex), 4..32 flop per byte.  Such a step is bound by HBM bandwidth
// as long as the math is cheap: on the CUDA cores it is not (stem.cu turns FMA-bound above
// ~8 flop/byte), on the tensor cores it is.  So this kernel streams A and C exactly once, like
// stem.cu, but multiplies on tcgen05:
//
//   * the right operand is expanded to its real form B'[2n + c'][2k + c], split into fp16 hi/lo
//     and written ONCE per CTA into shared memory in the K-major 128B-swizzled layout
//     tcgen05.mma reads;
//   * producer warps (one thread per output row) read the K amplitudes of their row straight
//     from A's own bit layout -- the permutation torch.einsum would materialise as a copy is
//     folded into the address computation, there is no pack pass --, scale the row by a power of
//     two of its own (fp16 range; exact), split it into fp16 hi/lo and store it into a
//     shared-memory slot in the same swizzled layout;
//     For short rows (K < 32) a thread owns R rows 128 apart, so that it always has ~256 bytes of
//     loads in flight: the R sub-tiles sit side by side along the K axis of the same 128-byte
//     swizzled rows, and the MMA of sub-tile i simply starts its A descriptor i * K bytes in;
//   * one thread of the producer group issues the 3 x K/16 MMAs of every 128-row sub-tile (lo*hi + hi*lo + hi*hi, or hi*hi
//     alone in the complex-half mode) into the slot's accumulator in tensor memory;
//   * epilogue warps read the accumulator (tcgen05.ld), undo the row and operand scales, and write
//     C[rows][m][n] through a shared-memory staging buffer so that every global store instruction
//     covers full 128-byte lines.
//
// The K sum of a tile is a single accumulation chunk (K <= 64 real), so the tensor core's
// round-toward-zero accumulator contributes the same small coherent bias as one chunk of the big
// GEMM (measured -3.6e-8 / -5.4e-8 / -8.6e-8 for 1 / 2 / 4 MMAs per product); the epilogue
// removes its mean (`debias`).
#include <algorithm>

#include "tc_common.cuh"

namespace tnc {

namespace {

constexpr int kGroups = 3;             // producer groups of 4 warps, taking tiles round-robin
constexpr int kProducerWarps = 4 * kGroups;
constexpr int kSkinnyThreads = (kProducerWarps + 4) * 32;   // + 4 epilogue warps (TMEM lane quarter = warp & 3); 16 warps: 128 registers each
constexpr int kAllocWarp = 0;          // owns the TMEM allocation
constexpr int kTileRows = 128;
constexpr int kTileBytes = kTileRows * 128;      // one K-major tile: 128 rows x 128 bytes
constexpr int kMaxSlots = 4;

struct SkinnyParams {
    const float2* a;
    const float2* b;
    float2* c;
    const int32_t* rows_a;
    const int32_t* rows_b;
    int32_t rows_mode_a, rows_mode_b;
    int32_t nbatch;
    int32_t rank_a, rank_b;
    int32_t mb, kb, nb;
    int32_t a_vec;                 // k bit 0 sits at A position 0: two k neighbours form one 16-byte load
    int32_t n_runs;                // row index -> A offset: runs of consecutive bits
    int32_t n_mma;                 // MMA N: max(16, 2 << nb)
    int32_t k_steps;               // MMAs per product: max(1, (2 << kb) / 16)
    int32_t slots;
    int32_t sub;                   // R: 128-row sub-tiles per slot
    float debias;
    int64_t tiles;
    uint32_t run_mask[TNC_MAX_BITS];
    int8_t run_src[TNC_MAX_BITS], run_dst[TNC_MAX_BITS];
    int8_t k_a[8], k_b[8], n_b[8];
    int32_t k_blocks;              // 1, or 2 for K = 64
    int32_t groups;                // active producer groups (<= kGroups, <= slots)
    int32_t fold;                  // rows of B folded into N (1 = none): accumulator columns [f * n_real, (f+1) * n_real)
                                   // of a tile are the outputs against row f of B and go to row block (batch * fold + f) of C
    uint64_t koff_c[32];           // A BYTE offset of contracted index k inside one k-block: kernel parameters live in the
                                   // constant bank, so with k unrolled the offset is an instruction operand (the shared-
                                   // memory table cost an LDS + a 64-bit IMAD per 8-byte load: 6 instructions per load)
    int64_t kblock_off;            // A offset of the second k-block (K = 64)
    uint32_t* amax_out;            // not null: the largest |component| written goes here (atomicMax of float bits)
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// KC = complex k per row and k-block, KB = k-blocks (K = KC * KB; KB == 2 only with KC == 32: the
// two producer groups then convert the two k-blocks of the SAME tile, each with its own row scale,
// into two accumulators that the epilogue adds in fp32 registers, round-to-nearest), R = 128-row
// sub-tiles per slot, PANELS = 2 (hi + lo) or 1 (hi only)
template <int KC, int KB, int R, int PANELS>
__global__ void __launch_bounds__(kSkinnyThreads, 1) skinny_kernel(const SkinnyParams p) {
    static_assert(KB == 1 || (KB == 2 && KC == 32 && R == 1), "two k-blocks: full rows, one sub-tile");
    constexpr int SUBS = KB * R;                          // accumulators (and row-scale vectors) per slot
    constexpr int PREC = PANELS == 2 ? TNC_TC_3XF16 : TNC_TC_F16;
    constexpr int KSUB = KC * 4 < 32 ? 32 : KC * 4;       // bytes of one sub-tile's K run inside a 128-byte row
    constexpr int CHUNKS = (KC + 3) / 4;                  // 16-byte chunks a row really carries
    static_assert(R * KSUB <= 128, "sub-tiles must fit one swizzle atom row");
    extern __shared__ unsigned char skinny_smem_raw[];
    const uint32_t raw = smem_u32(skinny_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* aligned = skinny_smem_raw + (base - raw);
    // layout: B' hi | B' lo | slots x (A hi | A lo) | slots x row scales | 4 staging buffers | barriers | misc
    const uint32_t b_bytes = (uint32_t)p.n_mma * 128u;                    // one part of one k-block; multiple of 2048
    const uint32_t b_hi = base, b_lo = base + b_bytes;                    // k-block kb: + kb * 2 * b_bytes
    const uint32_t a0 = base + KB * 2u * b_bytes;
    const uint32_t slot_bytes = KB * PANELS * kTileBytes;                 // k-block kb: + kb * PANELS * kTileBytes
    const uint32_t scales_off = KB * 2u * b_bytes + (uint32_t)p.slots * slot_bytes;
    float* row_scale = (float*)(aligned + scales_off);                    // [slots][SUBS][128]
    const uint32_t stage0 = base + scales_off + (uint32_t)p.slots * (SUBS * 512u); // 4 x 4 KB
    const uint32_t bars = stage0 + 4u * 4096u;
    uint32_t* misc = (uint32_t*)(aligned + (bars - base) + 3 * kMaxSlots * 8);   // [0] TMEM base, [1] amax(B) bits
    uint32_t* koff = misc + 4;                                            // [KC] A offset of contracted index k
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto done_bar = [&](int s) { return bars + 8u * (kMaxSlots + s); };
    auto free_bar = [&](int s) { return bars + 8u * (2 * kMaxSlots + s); };

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int K = KC * KB, N = 1 << p.nb;

    // ---------------------------------------------------------------- set-up (all threads)
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.slots; ++s) {
            mbar_init(full_bar(s), 4 * KB);         // one arrival per producer warp (its lane 0, behind a __syncwarp)
            mbar_init(done_bar(s), 1);
            mbar_init(free_bar(s), 4);
        }
        misc[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kAllocWarp) {
        uint32_t cols = (uint32_t)(p.slots * SUBS * p.n_mma);
        cols = cols < 32u ? 32u : cols;                                   // power of two by construction
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(misc)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int k = threadIdx.x; k < K; k += kSkinnyThreads) {
        uint32_t o = 0;
        for (int i = 0; i < p.kb; ++i) o |= ((uint32_t)(k >> i) & 1u) << p.k_a[i];
        koff[k] = o;
    }
    // zero B' and every A slot once: the padding (rows >= 2N, columns >= 2K) is never written again
    {
        const uint32_t total16 = (KB * 2u * b_bytes + (uint32_t)p.slots * slot_bytes) >> 4;
        for (uint32_t i = threadIdx.x; i < total16; i += kSkinnyThreads) sts128(base + (i << 4), 0u, 0u, 0u, 0u);
    }
    __syncthreads();
    // the right operand: one block (B has no rows, or a single output row block), or -- folded -- rows
    // 0 .. fold-1 of B stacked along N
    int64_t rb0 = 0;
    if (p.fold == 1 && p.rows_mode_b >= 0) rb0 = p.rows_b[0];
    {
        float m = 0.f;
        const float2* __restrict__ ball = p.b + (rb0 << p.rank_b);
        for (int e = threadIdx.x; e < p.fold * K * N; e += kSkinnyThreads) {
            const float2 x = ball[e];                                     // amax does not care about the order
            m = fmaxf(m, fmaxf(fabsf(x.x), fabsf(x.y)));
        }
        uint32_t bits = __reduce_max_sync(0xffffffffu, __float_as_uint(m));
        if (lane == 0 && bits) atomicMax(misc + 1, bits);
    }
    __syncthreads();
    const uint32_t b_amax = misc[1];
    {
        const float sc = f16_scale(b_amax);
        for (int ef = threadIdx.x; ef < p.fold * K * N; ef += kSkinnyThreads) {
            const int f = ef / (K * N), e = ef - f * (K * N);
            const float2* __restrict__ bsrc = p.b + ((rb0 + f) << p.rank_b);
            const int k = e & (K - 1), n = e >> p.kb;
            uint32_t o = 0;
            for (int i = 0; i < p.kb; ++i) o |= ((uint32_t)(k >> i) & 1u) << p.k_b[i];
            for (int i = 0; i < p.nb; ++i) o |= ((uint32_t)(n >> i) & 1u) << p.n_b[i];
            const float2 x = bsrc[o];
            const float xr = x.x * sc, xi = x.y * sc;
            const __half hr = __float2half_rn(xr), hi = __float2half_rn(xi);
            // B'[2n][2k] = Br, B'[2n][2k+1] = -Bi, B'[2n+1][2k] = Bi, B'[2n+1][2k+1] = Br
            // element (row, col) lives at row * 128 + ((col >> 3) ^ (row & 7)) * 16 + (col & 7) * 2
            const uint32_t col = 2u * (k & (KC - 1)), r0 = 2u * (f * N + n), r1 = r0 + 1u;
            const uint32_t kbo = (uint32_t)(k / KC) * 2u * b_bytes;       // this k's k-block
            const uint32_t off0 = r0 * 128u + ((((col >> 3) ^ (r0 & 7u)) << 4) | ((col & 7u) << 1));
            const uint32_t off1 = r1 * 128u + ((((col >> 3) ^ (r1 & 7u)) << 4) | ((col & 7u) << 1));
            const __half2 v0 = __halves2half2(hr, __hneg(hi)), v1 = __halves2half2(hi, hr);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(b_hi + kbo + off0), "r"(*(const uint32_t*)&v0) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(b_hi + kbo + off1), "r"(*(const uint32_t*)&v1) : "memory");
            if constexpr (PANELS == 2) {
                const __half lr = __float2half_rn(xr - __half2float(hr)), li = __float2half_rn(xi - __half2float(hi));
                const __half2 w0 = __halves2half2(lr, __hneg(li)), w1 = __halves2half2(li, lr);
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(b_lo + kbo + off0), "r"(*(const uint32_t*)&w0) : "memory");
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(b_lo + kbo + off1), "r"(*(const uint32_t*)&w1) : "memory");
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic-proxy writes -> visible to tcgen05.mma
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = misc[0];

    // a tile = R sub-tiles of 128 consecutive rows; mb >= 7 + log2(R) (checked by the launcher)
    const int64_t tiles_per_batch = ((int64_t)1 << (p.mb - 7)) / R;

    // the MMAs of one tile: every sub-tile / k-block into its own accumulator of the slot
    const uint32_t idesc = umma_idesc<PREC>(kTileRows, p.n_mma);
    const uint64_t db_hi = umma_desc(b_hi), db_lo = umma_desc(b_lo);
    // called by a whole converged warp: every lane waits, one elected lane issues (elect_one, tc_common.cuh)
    auto issue_mma = [&](int slot, uint32_t use) {
        mbar_wait(full_bar(slot), use & 1u);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            const uint32_t abase = a0 + (uint32_t)slot * slot_bytes + (uint32_t)kb * (PANELS * kTileBytes);
            const uint64_t da_hi = umma_desc(abase), da_lo = umma_desc(abase + kTileBytes);
            const uint64_t bh = db_hi + kb * ((2u * b_bytes) >> 4), bl = db_lo + kb * ((2u * b_bytes) >> 4);
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const uint32_t tacc = tmem + (uint32_t)((slot * SUBS + kb * R + i) * p.n_mma);
                const uint64_t ah = da_hi + i * (KSUB / 16), al = da_lo + i * (KSUB / 16);   // K offset of sub-tile i
                uint32_t acc = 0;
                if constexpr (PANELS == 2) {
                    for (int s = 0; s < p.k_steps; ++s) {
                        umma<PREC, 1>(tacc, al + 2 * s, bh + 2 * s, idesc, acc);
                        acc = 1;
                    }
                    for (int s = 0; s < p.k_steps; ++s) umma<PREC, 1>(tacc, ah + 2 * s, bl + 2 * s, idesc, 1);
                }
                for (int s = 0; s < p.k_steps; ++s) {
                    umma<PREC, 1>(tacc, ah + 2 * s, bh + 2 * s, idesc, acc);
                    acc = 1;
                }
            }
        }
        umma_commit(done_bar(slot));
        }
        __syncwarp();
    };

    if (warp < kProducerWarps) {
        // ------------------------------------------------------------ producers
        const int group = warp >> 2;
        const int row = ((warp & 3) << 5) | lane;                         // row inside the tile = TMEM lane
        // KB == 1: the groups take tiles round-robin; KB == 2: groups 0 and 1 convert k-blocks 0 and 1
        // of every tile (a third group has nothing to do)
        const int kb_mine = KB == 2 ? group : 0;
        // a group must never run two phases ahead of a slot's `free` barrier (parity waits alias):
        // with the epilogue retiring tiles in order that holds iff groups <= slots
        const bool idle = group >= p.groups;
        if (KB == 2 && group == 2) {
            // two k-blocks: groups 0 and 1 convert, the first warp of group 2 issues the MMAs
            if ((warp & 3) == 0)
                for (int64_t j = 0;; ++j) {
                    if ((int64_t)blockIdx.x + j * gridDim.x >= p.tiles) break;
                    issue_mma((int)(j % p.slots), (uint32_t)(j / p.slots));
                }
            __syncwarp();
        }
        int64_t row_off[R];                                               // A offset of this thread's row bits (sub-tile i)
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int64_t rl = (int64_t)(i << 7) | row;
            int64_t o = 0;
            for (int t = 0; t < p.n_runs; ++t) o |= ((rl >> p.run_src[t]) & (int64_t)p.run_mask[t]) << p.run_dst[t];
            row_off[i] = o;
        }
        // slot / use of tile number j without a division per tile: j advances by `step`
        const int step = KB == 2 ? 1 : p.groups;
        int slot = (KB == 2 ? 0 : group) % p.slots;
        uint32_t use = (uint32_t)((KB == 2 ? 0 : group) / p.slots);
        for (int64_t j = (KB == 2 ? 0 : group);; j += step) {
            const int64_t tile = (int64_t)blockIdx.x + j * gridDim.x;
            if (idle || tile >= p.tiles) break;
            const int64_t bidx = tile / tiles_per_batch;
            int64_t ra = 0;
            if (p.rows_mode_a == TNC_ROWS_IDENTITY) ra = bidx;
            else if (p.rows_mode_a >= 0) ra = p.rows_a[bidx];
            float2 av[R][KC];
            // the tile's part of the A offset is the same for the R rows of a thread; the thread's own
            // row bits (row_off) never change
            const int64_t r0 = ((tile - bidx * tiles_per_batch) * R) << 7;
            int64_t toff = (ra << p.rank_a) + (KB == 2 ? kb_mine * p.kblock_off : 0);
            for (int t = 0; t < p.n_runs; ++t) toff |= ((r0 >> p.run_src[t]) & (int64_t)p.run_mask[t]) << p.run_dst[t];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                // byte pointer of the row + a 32-bit byte offset per contracted index: one 64-bit add per load
                const char* __restrict__ ap = (const char*)(p.a + (toff | row_off[i]));
                if (KC >= 2 && p.a_vec) {
#pragma unroll
                    for (int k = 0; k < KC; k += 2) {
                        const float4 x = TNC_LDG((const float4*)(ap + p.koff_c[k]));
                        av[i][k] = make_float2(x.x, x.y);
                        av[i][k + 1] = make_float2(x.z, x.w);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KC; ++k) av[i][k] = TNC_LDG((const float2*)(ap + p.koff_c[k]));
                }
            }
            // the slot (shared memory tile, scales, TMEM accumulators) is free once the epilogue of
            // its previous tile is done
            mbar_wait(free_bar(slot), (use & 1u) ^ 1u);
            const uint32_t ahi = a0 + (uint32_t)slot * slot_bytes + (uint32_t)kb_mine * (PANELS * kTileBytes) + (uint32_t)row * 128u;
            const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
            for (int i = 0; i < R; ++i) {
                float m = 0.f;
#pragma unroll
                for (int k = 0; k < KC; ++k) m = fmaxf(m, fmaxf(fabsf(av[i][k].x), fabsf(av[i][k].y)));
                const uint32_t mbits = __float_as_uint(m);
                const float sc = f16_scale(mbits);                       // this row's own power of two
#pragma unroll
                for (int c = 0; c < CHUNKS; ++c) {                       // 16-byte chunk = 4 amplitudes = 8 halves
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int k = 4 * c + q;
                        if (k < KC) {
                            // packed fp32x2 math (sm_100): scale, and residual x * sc - hi, one instruction each
                            const float2 xs = __fmul2_rn(av[i][k], make_float2(sc, sc));
                            const __half2 hh = __float22half2_rn(xs);
                            h[q] = *(const uint32_t*)&hh;
                            if constexpr (PANELS == 2) {
                                const float2 lo = __ffma2_rn(__half22float2(hh), make_float2(-1.f, -1.f), xs);
                                const __half2 ll = __float22half2_rn(lo);
                                l[q] = *(const uint32_t*)&ll;
                            }
                        } else {
                            h[q] = 0u;
                            l[q] = 0u;
                        }
                    }
                    const uint32_t off = (((uint32_t)(i * (KSUB / 16) + c) ^ sw) << 4);
                    sts128(ahi + off, h[0], h[1], h[2], h[3]);
                    if constexpr (PANELS == 2) sts128(ahi + kTileBytes + off, l[0], l[1], l[2], l[3]);
                }
                row_scale[(slot * SUBS + kb_mine * R + i) * 128 + row] = f16_inv_scale(mbits);
            }
            // every thread orders its own tile stores before the tensor core's (async-proxy) reads; ONE arrival per
            // warp then: the waiting warps (MMA issuer, epilogue) are woken by every arrival on a barrier they sleep
            // on, and 128 arrivals per tile were ~15 % of all issued instructions
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(full_bar(slot));
            // KB == 1: one thread of the group issues the tile's MMAs once every producer of the tile
            // has arrived (KB == 2: the idle third group does, see below)
            if (KB == 1 && (warp & 3) == 0) issue_mma(slot, use);
            __syncwarp();
            slot += step;                                                 // step <= slots (groups <= slots)
            if (slot >= p.slots) {
                slot -= p.slots;
                ++use;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;
        const uint32_t stage = stage0 + (uint32_t)q * 4096u;
        const float b_inv = f16_inv_scale(b_amax);
        const int n_real = 2 << p.nb;                                     // floats per output row
        int slot = 0;
        uint32_t use = 0;
        float am = 0.f;                                                   // largest |component| this thread wrote (amax_out)
        for (int64_t j = 0;; ++j, ++slot) {
            const int64_t tile = (int64_t)blockIdx.x + j * gridDim.x;
            if (tile >= p.tiles) break;
            if (slot == p.slots) {                                        // slot = j % slots, use = j / slots
                slot = 0;
                ++use;
            }
            mbar_wait(done_bar(slot), use & 1u);                          // the tile's MMAs are done (sleeps on ONE arrival)
            mbar_wait(full_bar(slot), use & 1u);                          // complete by then: acquires the producers' row scales
            tc_fence_after();
            // one accumulator (n_real columns from TMEM column `tcol`) -> n_real floats per row at `cptr`;
            // `release`: this is the last read of the slot's accumulators
            auto emit = [&](uint32_t tcol, float* cptr, float s, float s1, bool release) {
                if (n_real >= 32) {
                    for (int c0 = 0; c0 < n_real; c0 += 32) {
                        uint32_t v[32], w[KB == 2 ? 32 : 1];
                        tmem_ld16(tcol + c0, v);
                        tmem_ld16(tcol + c0 + 16, v + 16);
                        if constexpr (KB == 2) {
                            tmem_ld16(tcol + p.n_mma + c0, w);
                            tmem_ld16(tcol + p.n_mma + c0 + 16, w + 16);
                        }
                        tmem_ld_wait();
                        if (release && c0 + 32 >= n_real) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(free_bar(slot));
                        }
                        float o[32];
#pragma unroll
                        for (int x = 0; x < 32; ++x) {
                            float y = __uint_as_float(v[x]) * s;
                            if constexpr (KB == 2) y = fmaf(__uint_as_float(w[x]), s1, y);
                            o[x] = fmaf(y, p.debias, y);
                        }
                        if (p.amax_out) {
#pragma unroll
                            for (int x = 0; x < 32; ++x) amax_fold(am, o[x]);
                        }
                        store_rows_coalesced<8>(stage, o, cptr + c0, n_real, lane);
                    }
                } else {
                    uint32_t v[16] = {}, w[KB == 2 ? 16 : 1] = {};
                    if (n_real == 16) tmem_ld16(tcol, v);
                    else tmem_ld8(tcol, v);
                    if constexpr (KB == 2) {
                        if (n_real == 16) tmem_ld16(tcol + p.n_mma, w);
                        else tmem_ld8(tcol + p.n_mma, w);
                    }
                    tmem_ld_wait();
                    if (release) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(free_bar(slot));
                    }
                    float o[16];
#pragma unroll
                    for (int x = 0; x < 16; ++x) {
                        float y = __uint_as_float(v[x]) * s;
                        if constexpr (KB == 2) y = fmaf(__uint_as_float(w[x]), s1, y);
                        o[x] = fmaf(y, p.debias, y);
                    }
                    if (p.amax_out) {
#pragma unroll
                        for (int x = 0; x < 16; ++x) amax_fold(am, o[x]);
                    }
                    if (n_real == 16) store_rows_coalesced<4>(stage, o, cptr, n_real, lane);
                    else if (n_real == 8) store_rows_coalesced<2>(stage, o, cptr, n_real, lane);
                    else store_rows_coalesced<1>(stage, o, cptr, n_real, lane);
                }
            };
            const int64_t bidx = tile / tiles_per_batch;
#pragma unroll 1
            for (int i = 0; i < R; ++i) {
                const float s = row_scale[(slot * SUBS + i) * 128 + q * 32 + lane] * b_inv;
                float s1 = 0.f;                                           // second k-block's scale (KB == 2)
                if constexpr (KB == 2) s1 = row_scale[(slot * SUBS + 1) * 128 + q * 32 + lane] * b_inv;
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((slot * SUBS + i) * p.n_mma);
                // C[rows][m][n]: the sub-tile's rows are consecutive, n_real floats each
                const int64_t row_in_batch = ((((tile - bidx * tiles_per_batch) * R + i) << 7) + q * 32);
                const bool last_sub = i == R - 1;
#pragma unroll 1
                for (int f = 0; f < p.fold; ++f) {
                    float* cptr = (float*)p.c + ((((bidx * p.fold + f) << p.mb) + row_in_batch) * (int64_t)n_real);
                    emit(taddr + (uint32_t)(f * n_real), cptr, s, s1, last_sub && f == p.fold - 1);
                }
            }
        }
        if (p.amax_out) amax_commit(p.amax_out, am);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp) {
        uint32_t cols = (uint32_t)(p.slots * SUBS * p.n_mma);
        cols = cols < 32u ? 32u : cols;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
    }
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int KC, int KB, int R, int PANELS>
int launch(const SkinnyParams& p, size_t smem, int grid, cudaStream_t s) {
    static bool configured_on[kMaxDevices] = {};       // the attribute is per device
    bool& configured = configured_on[current_device()];
    if (!configured) {
        TNC_CUDA(cudaFuncSetAttribute(skinny_kernel<KC, KB, R, PANELS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    skinny_kernel<KC, KB, R, PANELS><<<grid, kSkinnyThreads, smem, s>>>(p);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

template <int KC, int PANELS>
int launch_r(const SkinnyParams& p, size_t smem, int grid, cudaStream_t s) {
    if constexpr (KC <= 8) {
        if (p.sub == 4) return launch<KC, 1, 4, PANELS>(p, smem, grid, s);
    }
    if constexpr (KC <= 16) {
        if (p.sub == 2) return launch<KC, 1, 2, PANELS>(p, smem, grid, s);
    }
    return launch<KC, 1, 1, PANELS>(p, smem, grid, s);
}

template <int PANELS>
int launch_k(const SkinnyParams& p, size_t smem, int grid, cudaStream_t s) {
    switch (p.kb) {
        case 2: return launch_r<4, PANELS>(p, smem, grid, s);
        case 3: return launch_r<8, PANELS>(p, smem, grid, s);
        case 4: return launch_r<16, PANELS>(p, smem, grid, s);
        case 5: return launch_r<32, PANELS>(p, smem, grid, s);
        default: return launch<32, 2, 1, PANELS>(p, smem, grid, s);
    }
}

}  // namespace

// Rows of B handled by one launch: 1 (B has no rows / a single row block), the row count when they
// can be folded into N -- a plain step whose rows come from B alone, or a full outer step (all row
// pairs, A-major) --, 0 when the step needs a different right operand per row block.
int skinny_fold(const tnc_einsum& e) {
    if (e.rows_b == TNC_ROWS_NONE || e.nb == 1) return (e.rows_a == TNC_ROWS_NONE && e.nb != 1) ? 0 : 1;
    int fold = 0;
    if (e.rows_a == TNC_ROWS_NONE && e.rows_b == TNC_ROWS_IDENTITY) fold = e.nb;
    else if (e.flags & TNC_EINSUM_OUTER_ROWS) fold = e.b.rows;
    if (fold < 1 || e.n_k > 5 || e.n_n < 2 || ((int64_t)fold << (e.n_n + 1)) > 256) return 0;
    return fold;
}

bool skinny_supported(const tnc_einsum& e, int dtype, int precision) {
    if (dtype != TNC_C64 || e.n_h != 0) return false;
    if (precision != TNC_TC_3XF16 && precision != TNC_TC_F16) return false;
    if (e.n_k < 2 || e.n_k > 6 || e.n_n < 1 || e.n_n > 7 || e.n_m < 7) return false;
    if (e.n_k == 6 && e.n_n > 6) return false;                           // two k-blocks of B' and two slots must fit shared memory
    if (skinny_fold(e) == 0) return false;                               // one right operand (possibly folded) per launch
    for (int i = 0; i < e.n_n; ++i)
        if (e.n_c[i] >= e.n_n) return false;                             // output must be C[rows][m][n]
    return true;
}

int launch_skinny(const tnc_einsum& e, int precision, const void* a, const void* b, void* c, const int32_t* dev_rows_a,
                  const int32_t* dev_rows_b, cudaStream_t s, uint32_t* amax_out) {
    if (!skinny_supported(e, TNC_C64, precision)) {
        set_error("skinny einsum: unsupported shape, rows or output layout (m=%d k=%d n=%d h=%d)", e.n_m, e.n_k, e.n_n, e.n_h);
        return TNC_ERR_UNSUPPORTED;
    }
    SkinnyParams p{};
    p.amax_out = amax_out;
    p.a = (const float2*)a;
    p.b = (const float2*)b;
    p.c = (float2*)c;
    p.rows_a = dev_rows_a;
    p.rows_b = dev_rows_b;
    p.rows_mode_a = e.rows_a;
    p.rows_mode_b = e.rows_b;
    p.nbatch = e.nb;
    p.rank_a = e.a.rank;
    p.rank_b = e.b.rank;
    p.mb = e.n_m;
    p.kb = e.n_k;
    p.nb = e.n_n;
    for (int i = 0; i < e.n_k; ++i) {
        p.k_a[i] = e.k_a[i];
        p.k_b[i] = e.k_b[i];
    }
    for (int i = 0; i < e.n_n; ++i) p.n_b[e.n_c[i]] = e.n_b[i];
    p.a_vec = e.k_a[0] == 0;
    for (int k = 0; k < 32; ++k) {
        uint64_t o = 0;
        for (int i = 0; i < e.n_k && i < 5; ++i) o |= (uint64_t)((k >> i) & 1) << e.k_a[i];
        p.koff_c[k] = o << 3;                                            // bytes
    }
    p.kblock_off = e.n_k == 6 ? ((int64_t)1 << e.k_a[5]) : 0;
    // row bit j (output position n_n + j) -> A position; merge consecutive bits into runs
    int8_t pa[TNC_MAX_BITS];
    for (int i = 0; i < e.n_m; ++i) pa[e.m_c[i] - e.n_n] = e.m_a[i];
    p.n_runs = 0;
    for (int j = 0; j < e.n_m;) {
        int len = 1;
        while (j + len < e.n_m && pa[j + len] == pa[j] + len) ++len;
        p.run_src[p.n_runs] = (int8_t)j;
        p.run_dst[p.n_runs] = pa[j];
        p.run_mask[p.n_runs] = len >= 32 ? 0xffffffffu : ((1u << len) - 1u);
        ++p.n_runs;
        j += len;
    }
    const int n_real = 2 << e.n_n, k_real = 2 << e.n_k;
    p.fold = skinny_fold(e);
    p.n_mma = 16;
    while (p.n_mma < p.fold * n_real) p.n_mma <<= 1;                     // power of two: TMEM allocations are
    p.k_blocks = e.n_k == 6 ? 2 : 1;
    p.k_steps = std::max(1, (k_real / p.k_blocks) / 16);
    // rows per producer thread: ~256 bytes of loads in flight, and >= 2 slots of accumulators in TMEM
    p.sub = std::max(1, std::min(e.n_k <= 3 ? 4 : e.n_k == 4 ? 2 : 1, 128 / p.n_mma));   // 4 slots of accumulators in TMEM when possible
    while (p.sub > 1 && e.n_m < 7 + (p.sub == 4 ? 2 : 1)) p.sub >>= 1;
    p.slots = std::min(kMaxSlots, 512 / (p.sub * p.k_blocks * p.n_mma));
    if (p.k_blocks == 2) p.slots = 2;                                    // 64 KB of A per slot
    p.groups = p.k_blocks == 2 ? 2 : std::min(kGroups, p.slots);
    // mean of the accumulator's round-toward-zero bias, by MMAs per product (tools/tc_calibrate.py)
    p.debias = p.k_steps >= 4 ? 8.6e-8f : p.k_steps == 2 ? 5.4e-8f : 3.6e-8f;
    if (precision == TNC_TC_F16) p.debias = 0.f;
    if (p.fold > 1) {
        p.nbatch = (e.flags & TNC_EINSUM_OUTER_ROWS) ? e.a.rows : 1;     // outer batch = row of A
        if (e.flags & TNC_EINSUM_OUTER_ROWS) p.rows_mode_a = TNC_ROWS_IDENTITY;
    }
    p.tiles = (((int64_t)p.nbatch << e.n_m) >> 7) / p.sub;
    const int panels = precision == TNC_TC_F16 ? 1 : 2;
    const size_t smem = 1024 + p.k_blocks * 2 * (size_t)p.n_mma * 128 + (size_t)p.slots * p.k_blocks * panels * kTileBytes +
                        (size_t)p.slots * p.sub * p.k_blocks * 512 +
                        4 * 4096 + 3 * kMaxSlots * 8 + 16 + 64 * 4 + 64;
    const int grid = (int)std::min<int64_t>(p.tiles, sm_count());
    return panels == 2 ? launch_k<2>(p, smem, grid, s) : launch_k<1>(p, smem, grid, s);
}

}  // namespace tnc
