// Streaming kernel for the "stem" steps of a contraction tree: a huge left operand (2^20..2^30
// amplitudes) meets a tiny right operand (K, N <= 64), arithmetic intensity 2..20 flop/byte, so
// the step is bound by HBM bandwidth, not by math.  It is NOT reshaped into a tensor-core GEMM:
//
//   * one thread owns one output row: it reads the K amplitudes of its row of A straight from
//     A's own bit layout (the row index is deposited into A's address bits -- the permutation
//     torch.einsum would materialise as a copy is folded into the address computation), keeps
//     them in registers, and produces the N outputs of the row with fp32 FMAs (round to nearest,
//     the same arithmetic as the reference's cgemm);
//   * the right operand is gathered once per CTA into shared memory as B[k][n] and read with
//     warp-uniform (broadcast) 128-bit loads: 8 FMAs per shared-memory instruction;
//   * consecutive threads own consecutive rows, rows are numbered by A's lowest address bits, and
//     the output is written as C[rows][m][n] through a per-warp shared-memory staging buffer, so
//     that a store instruction covers full 128-byte lines instead of one line per lane.
//
// HBM traffic is the algorithmic minimum (A read once, C written once).  fp32 FMA throughput
// (~70 TFLOP/s) becomes the limit above ~11 flop/byte; steps above ~24 flop/byte go to the
// tcgen05 path instead.
#include <algorithm>
#include <cstdlib>

#include "tc_common.cuh"

namespace tnc {

namespace {

constexpr int kStemThreads = 256;

struct StemParams {
    const float2* a;
    const float2* b;
    float2* c;
    const int32_t* rows_a;
    const int32_t* rows_b;
    int32_t rows_mode_a, rows_mode_b;
    int32_t nbatch;                // outer batches (each reloads the right operand(s))
    int32_t fold;                  // right-operand rows handled per row of A read (1 = none), see launch_stem
    int32_t rank_a, rank_b;
    int32_t mb, kb, nb;
    int32_t a_vec;                 // k bit 0 sits at A position 0: two k neighbours form one 16-byte load
    int32_t n_runs;                // row index -> A offset: runs of consecutive bits
    uint32_t run_mask[TNC_MAX_BITS];
    int8_t run_src[TNC_MAX_BITS], run_dst[TNC_MAX_BITS];
    int8_t k_a[TNC_MAX_BITS], k_b[TNC_MAX_BITS], n_b[TNC_MAX_BITS];
    uint32_t* amax_out;            // not null: the largest |component| written goes here (atomicMax of float bits)
};

__device__ __forceinline__ void cfma(float2& acc, const float2 a, const float2 b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}

// Packed form (fma.rn.f32x2, sm_100): TWO outputs n, n+1 per instruction.  The right operand is kept
// PLANAR in shared memory (real plane, imaginary plane), the accumulators planar in registers:
//   accR(n, n+1) += a.x * bR(n, n+1) - a.y * bI(n, n+1)      accI(n, n+1) += a.x * bI(n, n+1) + a.y * bR(n, n+1)
// -- the same four fused multiply-adds per output in the same order as cfma (bit-identical results), in
// half the instructions: these kernels are bound by issue slots, not by the FMA count, once the SM
// clock sits at the power-capped ~1.2 GHz of a long slice.
struct RowAmp {
    float2 x, y, ny;          // (a.x, a.x), (a.y, a.y), (-a.y, -a.y)
};
__device__ __forceinline__ RowAmp row_amp(const float2 a) {
    RowAmp r;
    r.x = make_float2(a.x, a.x);
    r.y = make_float2(a.y, a.y);
    r.ny = make_float2(-a.y, -a.y);
    return r;
}
__device__ __forceinline__ void cfma2(float2& accR, float2& accI, const RowAmp& a, const float2 bR, const float2 bI) {
#ifdef TNC_STEM_SCALAR      // A/B build: the same planar data flow with scalar FFMAs
    accR.x = fmaf(a.x.x, bR.x, accR.x), accR.y = fmaf(a.x.x, bR.y, accR.y);
    accR.x = fmaf(a.ny.x, bI.x, accR.x), accR.y = fmaf(a.ny.x, bI.y, accR.y);
    accI.x = fmaf(a.x.x, bI.x, accI.x), accI.y = fmaf(a.x.x, bI.y, accI.y);
    accI.x = fmaf(a.y.x, bR.x, accI.x), accI.y = fmaf(a.y.x, bR.y, accI.y);
#else
    accR = __ffma2_rn(a.x, bR, accR);
    accR = __ffma2_rn(a.ny, bI, accR);
    accI = __ffma2_rn(a.x, bI, accI);
    accI = __ffma2_rn(a.y, bR, accI);
#endif
}
// One row amplitude against NCH outputs: bR / bI point at the NCH reals / imaginaries of B[k][n0 ..]
// (warp-uniform shared-memory addresses: broadcast loads).
template <int NCH>
__device__ __forceinline__ void row_fma(float2* accR, float2* accI, const float2 a, const float* bR, const float* bI) {
    static_assert(NCH >= 2 && NCH % 2 == 0, "packed path: pairs of outputs");
    const RowAmp ra = row_amp(a);
    if constexpr (NCH == 2) {
        cfma2(accR[0], accI[0], ra, *(const float2*)bR, *(const float2*)bI);
    } else {
#pragma unroll
        for (int i = 0; i < NCH; i += 4) {
            const float4 r = *(const float4*)(bR + i), m = *(const float4*)(bI + i);
            cfma2(accR[i / 2], accI[i / 2], ra, make_float2(r.x, r.y), make_float2(m.x, m.y));
            cfma2(accR[i / 2 + 1], accI[i / 2 + 1], ra, make_float2(r.z, r.w), make_float2(m.z, m.w));
        }
    }
}
// planar accumulators -> interleaved (re, im) outputs
template <int NCH>
__device__ __forceinline__ void interleave(const float2* accR, const float2* accI, float2* out) {
#pragma unroll
    for (int i = 0; i < NCH / 2; ++i) {
        out[2 * i] = make_float2(accR[i].x, accI[i].x);
        out[2 * i + 1] = make_float2(accR[i].y, accI[i].y);
    }
}

// NCH: outputs (complex) a thread produces per pass over its row; KCH: row amplitudes per chunk.
// Light variants (few accumulators) are compiled for 4 resident CTAs per SM: their rows are short,
// and the bytes in flight per SM are what bounds them.
template <int NCH, int KCH>
__global__ void __launch_bounds__(kStemThreads, (NCH * KCH <= 32 && NCH <= 8 ? 4 : 2)) stem_kernel(const StemParams p) {
    extern __shared__ __align__(16) unsigned char stem_smem[];
    const int K = 1 << p.kb, N = 1 << p.nb;
    float* BsR = (float*)stem_smem;                     // [fold][K][N] real plane, then the imaginary plane
    float* BsI = BsR + (size_t)p.fold * K * N;
    uint32_t* koff = (uint32_t*)(BsI + (size_t)p.fold * K * N);   // A offset of contracted index k
    const uint32_t stage = ((smem_u32(koff + K) + 15u) & ~15u) + (threadIdx.x >> 5) * 4096u;   // this warp's staging buffer
    const bool staged = NCH >= 2 && p.mb >= 5;           // whole warps only
    const int lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < K; k += kStemThreads) {
        uint32_t o = 0;
        for (int i = 0; i < p.kb; ++i) o |= ((uint32_t)(k >> i) & 1u) << p.k_a[i];
        koff[k] = o;
    }
    const int64_t rows = (int64_t)1 << p.mb;
    const int64_t tiles_per_batch = (rows + kStemThreads - 1) / kStemThreads;
    const int64_t tiles = tiles_per_batch * p.nbatch;
    int cur_batch = -1;
    float am = 0.f;                                      // largest |component| this thread wrote (amax_out)
    // Each CTA walks a contiguous range of tiles: with many small batches (outer steps: thousands of
    // row pairs of a few hundred tiles each) a strided walk would change batch -- reload B, two
    // barriers -- on every tile.
    const int64_t t_begin = tiles * blockIdx.x / gridDim.x, t_end = tiles * (blockIdx.x + 1) / gridDim.x;
    for (int64_t t = t_begin; t < t_end; ++t) {
        const int b = (int)(t / tiles_per_batch);
        if (b != cur_batch) {
            __syncthreads();
            for (int f = 0; f < p.fold; ++f) {
                int64_t rb = 0;
                if (p.fold > 1) rb = f;                                   // folded: rows 0 .. fold-1 of B
                else if (p.rows_mode_b == TNC_ROWS_IDENTITY) rb = b;
                else if (p.rows_mode_b >= 0) rb = p.rows_b[b];
                const float2* __restrict__ bsrc = p.b + (rb << p.rank_b);
                for (int e = threadIdx.x; e < K * N; e += kStemThreads) {
                    const int k = e >> p.nb, n = e & (N - 1);
                    uint32_t o = 0;
                    for (int i = 0; i < p.kb; ++i) o |= ((uint32_t)(k >> i) & 1u) << p.k_b[i];
                    for (int i = 0; i < p.nb; ++i) o |= ((uint32_t)(n >> i) & 1u) << p.n_b[i];
                    const float2 x = bsrc[o];
                    BsR[(size_t)f * K * N + e] = x.x;
                    BsI[(size_t)f * K * N + e] = x.y;
                }
            }
            __syncthreads();
            cur_batch = b;
        }
        const int64_t r = (t - (int64_t)b * tiles_per_batch) * kStemThreads + threadIdx.x;
        if (r >= rows) continue;
        int64_t ra = 0;
        if (p.rows_mode_a == TNC_ROWS_IDENTITY) ra = b;
        else if (p.rows_mode_a >= 0) ra = p.rows_a[b];
        int64_t aoff = ra << p.rank_a;
        for (int i = 0; i < p.n_runs; ++i) aoff |= ((r >> p.run_src[i]) & (int64_t)p.run_mask[i]) << p.run_dst[i];
        const float2* __restrict__ ap = p.a + aoff;
        // folded: the row of A is read from HBM once and meets every row of B (the repeats hit L1/L2)
#pragma unroll 1
        for (int f = 0; f < p.fold; ++f) {
        const float* __restrict__ BfR = BsR + (size_t)f * K * N;
        const float* __restrict__ BfI = BsI + (size_t)f * K * N;
        float2* __restrict__ cp = p.c + ((((int64_t)b * p.fold + f) << p.mb) + r) * N;
#pragma unroll 1
        for (int n0 = 0; n0 < N; n0 += NCH) {
            float2 acc[NCH];                             // NCH == 1: the output; else (accR, accI) pairs of outputs
#pragma unroll
            for (int i = 0; i < NCH; ++i) acc[i] = make_float2(0.f, 0.f);
            float2* accR = acc;
            float2* accI = acc + (NCH >= 2 ? NCH / 2 : 0);
#pragma unroll 1
            for (int k0 = 0; k0 < K; k0 += KCH) {
                float2 av[KCH];
                if (KCH >= 2 && p.a_vec) {
#pragma unroll
                    for (int j = 0; j < KCH; j += 2) {
                        const float4 x = *(const float4*)(ap + koff[k0 + j]);
                        av[j] = make_float2(x.x, x.y);
                        av[j + 1] = make_float2(x.z, x.w);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < KCH; ++j) av[j] = ap[koff[k0 + j]];
                }
#pragma unroll
                for (int j = 0; j < KCH; ++j) {
                    const size_t bo = (size_t)(k0 + j) * N + n0;
                    if constexpr (NCH == 1) cfma(acc[0], av[j], make_float2(BfR[bo], BfI[bo]));
                    else row_fma<NCH>(accR, accI, av[j], BfR + bo, BfI + bo);
                }
            }
            if (p.amax_out) {
#pragma unroll
                for (int i = 0; i < NCH; ++i) {
                    amax_fold(am, acc[i].x);
                    amax_fold(am, acc[i].y);
                }
            }
            if constexpr (NCH == 1) {
                cp[n0] = acc[0];
            } else {
                float2 out[NCH];
                interleave<NCH>(accR, accI, out);
                if (staged) {
                    store_rows_coalesced<NCH / 2>(stage, (const float*)out, (float*)(cp - (int64_t)lane * N + n0), 2 * (int64_t)N, lane);
                } else {
#pragma unroll
                    for (int i = 0; i < NCH; i += 2)
                        *(float4*)(cp + n0 + i) = make_float4(out[i].x, out[i].y, out[i + 1].x, out[i + 1].y);
                }
            }
        }
        }
    }
    if (p.amax_out) amax_commit(p.amax_out, am);
}

// -------------------------------------------------------------------------------------
// Bulk-copy variant for K <= 16 (the n <= 4, k <= 4 steps with 2..8 flop/byte, where bytes in flight
// per SM are what bounds the kernel above).  A tile is 256 consecutive rows; because the row index is
// numbered by A's address bits in ascending order, the tile's amplitudes are the 2^T lowest
// address bits of A (8 row bits + the kl contracted bits that sit below them) times the 2^(kb-kl)
// combinations of the higher contracted bits: 2^(kb-kl) contiguous chunks of 2^T amplitudes.
//   * a producer warp moves whole chunks HBM -> shared memory with cp.async.bulk into a ring of
//     stages (mbarrier full / empty), so 48-96 KB per CTA are in flight whatever the consumers do;
//   * 256 consumer threads (one per row) copy their K amplitudes from the stage into registers,
//     release the stage at once (behind a fence.proxy.async: the next write of the stage comes
//     through the async proxy; without it concurrent kernels on the same SM exposed a race), and keep them for every row of the right operand that meets this
//     row of A ("segment": the batches of the step that share the A row -- all of them for a
//     plain step with rows on B, b.rows for a full outer step, a run of the A-row table for an
//     outer step with a row subset).  A is therefore read from HBM once;
//   * every row of the right operand stays resident in shared memory as B[row][k][n].
struct BulkParams {
    const float2* a;
    const float2* b;
    float2* c;
    const int32_t* rows_a;
    const int32_t* rows_b;
    const int32_t* seg_begin;      // n_seg + 1 batch indices, or nullptr (see launch_stem)
    int32_t rows_mode_a, rows_mode_b;
    int32_t n_seg, nbatch, b_rows;
    int32_t rank_a, rank_b, mb, kb, nb;
    int32_t T;                     // address bits of one chunk
    int32_t stages;
    uint32_t off_b, off_koff, off_bars, off_staging;   // shared-memory byte offsets
    int32_t n_runs;                // tile index -> A offset: runs of consecutive bits
    uint32_t run_mask[TNC_MAX_BITS];
    int8_t run_src[TNC_MAX_BITS], run_dst[TNC_MAX_BITS];
    int8_t row_lo[8];              // A position (< T) of the tile's row bit j
    int8_t k_a[TNC_MAX_BITS], k_b[TNC_MAX_BITS], n_b[TNC_MAX_BITS];
    uint32_t* amax_out;            // as in StemParams
};

constexpr int kBulkConsumers = 256;
constexpr int kBulkThreads = kBulkConsumers + 32;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

template <int NCH, int KB, int PER_SM>
__global__ void __launch_bounds__(kBulkThreads, PER_SM) stem_bulk_kernel(const BulkParams p) {
    constexpr int K = 1 << KB;
    extern __shared__ __align__(128) unsigned char bulk_smem[];
    const int N = NCH < 16 ? NCH : (1 << p.nb);        // the launcher picks NCH = min(N, 16)
    const uint32_t stage_bytes = (uint32_t)(K << 8) * 8u;
    float* BsR = (float*)(bulk_smem + p.off_b);                  // [b_rows][K][N] real plane, then the imaginary plane
    float* BsI = BsR + (size_t)p.b_rows * K * N;
    uint32_t* koff_s = (uint32_t*)(bulk_smem + p.off_koff);      // stage offset (amplitudes) of contracted index k
    const uint32_t bar_full = smem_u32(bulk_smem + p.off_bars), bar_empty = bar_full + 8u * (uint32_t)p.stages;
    const uint32_t stage0 = smem_u32(bulk_smem);
    const int tid = threadIdx.x;
    for (int e = tid; e < p.b_rows * K * N; e += kBulkThreads) {
        const int row = e / (K * N), rem = e - row * (K * N);
        const int k = rem >> p.nb, n = rem & (N - 1);
        uint32_t o = 0;
        for (int i = 0; i < KB; ++i) o |= ((uint32_t)(k >> i) & 1u) << p.k_b[i];
        for (int i = 0; i < p.nb; ++i) o |= ((uint32_t)(n >> i) & 1u) << p.n_b[i];
        const float2 x = p.b[((int64_t)row << p.rank_b) + o];
        BsR[e] = x.x;
        BsI[e] = x.y;
    }
    if (tid < K) {
        uint32_t lo = 0, c = 0;
        int ci = 0;
        for (int i = 0; i < KB; ++i) {
            const uint32_t bit = ((uint32_t)tid >> i) & 1u;
            if (p.k_a[i] < p.T) lo |= bit << p.k_a[i];
            else c |= bit << ci++;
        }
        koff_s[tid] = (c << p.T) | lo;
    }
    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(bar_full + 8u * s, 1);
            mbar_init(bar_empty + 8u * s, kBulkConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tile_bits = p.mb - 8;
    const int64_t items = (int64_t)p.n_seg << tile_bits;
    const int64_t i_begin = items * blockIdx.x / gridDim.x, i_end = items * (blockIdx.x + 1) / gridDim.x;
    const int64_t tile_mask = ((int64_t)1 << tile_bits) - 1;
    int s = 0;
    uint32_t ph = 0;
    if (uniform_warp_idx() == kBulkConsumers / 32) {
        // ---- producer warp: every lane walks the tiles and waits, one elected lane issues the bulk
        // copies of each tile (elect_one, tc_common.cuh: descriptors stay in uniform registers)
        int n_hi = 0;
        int8_t hi_pos[4];
        for (int i = 0; i < KB; ++i)
            if (p.k_a[i] >= p.T) hi_pos[n_hi++] = p.k_a[i];
        const uint32_t chunk_bytes = 8u << p.T;
        int64_t cur_seg = -1, ra = 0;
        for (int64_t it = i_begin; it < i_end; ++it) {
            const int64_t seg = it >> tile_bits, tile = it & tile_mask;
            if (seg != cur_seg) {
                cur_seg = seg;
                if (p.rows_mode_a == TNC_ROWS_NONE) ra = 0;
                else if (p.seg_begin) {
                    const int32_t b0 = p.seg_begin[seg];
                    ra = p.rows_mode_a >= 0 ? p.rows_a[b0] : b0;
                } else ra = seg;
            }
            int64_t aoff = ra << p.rank_a;
            for (int i = 0; i < p.n_runs; ++i) aoff |= ((tile >> p.run_src[i]) & (int64_t)p.run_mask[i]) << p.run_dst[i];
            mbar_wait(bar_empty + 8u * s, ph ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(bar_full + 8u * s, stage_bytes);
                const uint32_t dst = stage0 + (uint32_t)s * stage_bytes;
                for (int c = 0; c < (1 << n_hi); ++c) {
                    int64_t o = aoff;
                    for (int i = 0; i < n_hi; ++i) o |= (int64_t)((c >> i) & 1) << hi_pos[i];
                    bulk_load(dst + (uint32_t)c * chunk_bytes, p.a + o, chunk_bytes, bar_full + 8u * s);
                }
            }
            __syncwarp();
            if (++s == p.stages) {
                s = 0;
                ph ^= 1u;
            }
        }
        // stay until the last copy has landed: the issuing warp outlives its bulk copies
        if (i_end > i_begin) {
            const int last = s == 0 ? p.stages - 1 : s - 1;
            mbar_wait(bar_full + 8u * last, s == 0 ? ph ^ 1u : ph);
        }
        return;
    }
    // ---- consumers: one thread per row of the tile
    const int lane = tid & 31;
    const uint32_t stage_w = smem_u32(bulk_smem + p.off_staging) + (uint32_t)(tid >> 5) * (256u * NCH);
    uint32_t roff = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) roff |= (((uint32_t)tid >> j) & 1u) << p.row_lo[j];
    int64_t cur_seg = -1;
    int32_t b0 = 0, b1 = 0;
    float am = 0.f;                                      // largest |component| this thread wrote (amax_out)
    for (int64_t it = i_begin; it < i_end; ++it) {
        const int64_t seg = it >> tile_bits, tile = it & tile_mask;
        if (seg != cur_seg) {
            cur_seg = seg;
            if (p.seg_begin) {
                b0 = p.seg_begin[seg];
                b1 = p.seg_begin[seg + 1];
            } else if (p.rows_mode_a == TNC_ROWS_NONE) {
                b0 = 0;
                b1 = p.nbatch;
            } else {
                b0 = (int32_t)seg;
                b1 = b0 + 1;
            }
        }
        mbar_wait(bar_full + 8u * s, ph);
        const float2* st = (const float2*)(bulk_smem + (size_t)s * stage_bytes) + roff;
        float2 av[K];
#pragma unroll
        for (int k = 0; k < K; ++k) av[k] = st[koff_s[k]];
        // the stage is next written through the async proxy: order these generic-proxy reads before it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8u * s);
        if (++s == p.stages) {
            s = 0;
            ph ^= 1u;
        }
        const int64_t r = (tile << 8) + tid;
#pragma unroll 1
        for (int32_t b = b0; b < b1; ++b) {
            int32_t rb = 0;
            if (p.rows_mode_b == TNC_ROWS_IDENTITY) rb = b;
            else if (p.rows_mode_b >= 0) rb = p.rows_b[b];
            const float* __restrict__ BfR = BsR + (size_t)rb * K * N;
            const float* __restrict__ BfI = BsI + (size_t)rb * K * N;
            float2* __restrict__ cp = p.c + ((((int64_t)b) << p.mb) + r) * N;
#pragma unroll 1
            for (int n0 = 0; n0 < N; n0 += NCH) {
                float2 acc[NCH];                         // NCH == 1: the output; else (accR, accI) pairs of outputs
#pragma unroll
                for (int i = 0; i < NCH; ++i) acc[i] = make_float2(0.f, 0.f);
                float2* accR = acc;
                float2* accI = acc + (NCH >= 2 ? NCH / 2 : 0);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const size_t bo = (size_t)k * N + n0;
                    if constexpr (NCH == 1) cfma(acc[0], av[k], make_float2(BfR[bo], BfI[bo]));
                    else row_fma<NCH>(accR, accI, av[k], BfR + bo, BfI + bo);
                }
                if (p.amax_out) {
#pragma unroll
                    for (int i = 0; i < NCH; ++i) {
                        amax_fold(am, acc[i].x);
                        amax_fold(am, acc[i].y);
                    }
                }
                if constexpr (NCH == 1) {
                    cp[n0] = acc[0];
                } else {
                    float2 out[NCH];
                    interleave<NCH>(accR, accI, out);
                    store_rows_coalesced<NCH / 2>(stage_w, (const float*)out, (float*)(cp - (int64_t)lane * N + n0), 2 * (int64_t)N, lane);
                }
            }
        }
    }
    if (p.amax_out) amax_commit(p.amax_out, am);
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

template <int NCH, int KCH>
int launch(const StemParams& p, size_t smem, int grid, cudaStream_t s) {
    static bool configured_on[kMaxDevices] = {};       // the attribute is per device
    bool& configured = configured_on[current_device()];
    if (!configured) {
        TNC_CUDA(cudaFuncSetAttribute(stem_kernel<NCH, KCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured = true;
    }
    stem_kernel<NCH, KCH><<<grid, kStemThreads, smem, s>>>(p);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

template <int NCH>
int launch_k(const StemParams& p, size_t smem, int grid, cudaStream_t s) {
    switch (p.kb) {
        case 0: return launch<NCH, 1>(p, smem, grid, s);
        case 1: return launch<NCH, 2>(p, smem, grid, s);
        case 2: return launch<NCH, 4>(p, smem, grid, s);
        default: return launch<NCH, 8>(p, smem, grid, s);
    }
}

template <int NCH, int KB, int PER_SM>
int launch_bulk(const BulkParams& p, size_t smem, int grid, cudaStream_t s) {
    static bool configured_on[kMaxDevices] = {};
    bool& configured = configured_on[current_device()];
    if (!configured) {
        TNC_CUDA(cudaFuncSetAttribute(stem_bulk_kernel<NCH, KB, PER_SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
        configured = true;
    }
    stem_bulk_kernel<NCH, KB, PER_SM><<<grid, kBulkThreads, smem, s>>>(p);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

template <int NCH>
int launch_bulk_k(const BulkParams& p, size_t smem, int grid, int per_sm, cudaStream_t s) {
    if (per_sm >= 3) {
        switch (p.kb) {
            case 0: return launch_bulk<NCH, 0, 3>(p, smem, grid, s);
            case 1: return launch_bulk<NCH, 1, 3>(p, smem, grid, s);
            case 2: return launch_bulk<NCH, 2, 3>(p, smem, grid, s);
            case 3: return launch_bulk<NCH, 3, 3>(p, smem, grid, s);
            default: return launch_bulk<NCH, 4, 3>(p, smem, grid, s);
        }
    }
    switch (p.kb) {
        case 0: return launch_bulk<NCH, 0, 2>(p, smem, grid, s);
        case 1: return launch_bulk<NCH, 1, 2>(p, smem, grid, s);
        case 2: return launch_bulk<NCH, 2, 2>(p, smem, grid, s);
        case 3: return launch_bulk<NCH, 3, 2>(p, smem, grid, s);
        default: return launch_bulk<NCH, 4, 2>(p, smem, grid, s);
    }
}

// The bulk-copy variant applies when a row's K <= 16 amplitudes fit registers, tiles are whole
// (>= 256 rows) and every row of the right operand fits shared memory together.
bool bulk_applies(const tnc_einsum& e) {
    const size_t b_bytes = ((size_t)8 << (e.n_k + e.n_n)) * (size_t)e.b.rows;
    return e.n_k <= 4 && e.n_m >= 8 && b_bytes <= 32 * 1024;
}

int launch_stem_bulk(const tnc_einsum& e, const void* a, const void* b, void* c, const int32_t* dev_rows_a,
                     const int32_t* dev_rows_b, const int32_t* dev_seg_begin, int n_seg, uint32_t* amax_out, cudaStream_t s) {
    BulkParams p{};
    p.amax_out = amax_out;
    p.a = (const float2*)a;
    p.b = (const float2*)b;
    p.c = (float2*)c;
    p.rows_a = dev_rows_a;
    p.rows_b = dev_rows_b;
    p.seg_begin = dev_seg_begin;
    p.rows_mode_a = e.rows_a;
    p.rows_mode_b = e.rows_b;
    p.nbatch = e.nb;
    p.n_seg = dev_seg_begin ? n_seg : (e.rows_a == TNC_ROWS_NONE ? 1 : e.nb);
    p.b_rows = e.b.rows;
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
    // row bit j (output position n_n + j) -> A position, ascending (the planner numbers the rows by
    // A's address bits); a layout that does not is left to the per-thread kernel
    int8_t pa[TNC_MAX_BITS];
    for (int i = 0; i < e.n_m; ++i) pa[e.m_c[i] - e.n_n] = e.m_a[i];
    for (int j = 1; j < e.n_m; ++j)
        if (pa[j] < pa[j - 1]) return TNC_ERR_UNSUPPORTED;
    p.T = pa[7] + 1;
    for (int j = 0; j < 8; ++j) p.row_lo[j] = pa[j];
    p.n_runs = 0;
    for (int j = 8; j < e.n_m;) {
        int len = 1;
        while (j + len < e.n_m && pa[j + len] == pa[j] + len) ++len;
        p.run_src[p.n_runs] = (int8_t)(j - 8);
        p.run_dst[p.n_runs] = pa[j];
        p.run_mask[p.n_runs] = len >= 32 ? 0xffffffffu : ((1u << len) - 1u);
        ++p.n_runs;
        j += len;
    }
    const int nch = std::min(1 << e.n_n, 16);
    const size_t stage_bytes = (size_t)8 << (8 + e.n_k);
    const size_t b_bytes = ((size_t)8 << (e.n_k + e.n_n)) * (size_t)e.b.rows;
    const size_t warp_staging = nch >= 2 ? (size_t)256 * nch : 0;      // 32 rows x nch outputs
    const size_t staging = (size_t)(kBulkConsumers / 32) * warp_staging;
    const size_t fixed = ((b_bytes + 15) & ~(size_t)15) + 64 + 16 * 8 + staging + 128;
    // Three resident CTAs hide the consumers' latencies better at the power-capped clock of a
    // long run, but the variants with >= 8 outputs x >= 8 amplitudes per row need more than the 72
    // registers that leaves them (measured inside n53 / n30 slices: 1.18 vs 1.38 ms for n = k = 3
    // at two CTAs, 2.15 vs 3.2 ms for n = 2, k = 4 at three).
    static const int forced_per_sm = knob("TNC_STEM_BULK_CTAS") ? atoi(knob("TNC_STEM_BULK_CTAS")) : 0;
    int per_sm = forced_per_sm ? (forced_per_sm >= 3 ? 3 : 2) : ((nch >= 8 && e.n_k >= 3) || nch >= 16 ? 2 : 3);
    int stages = (int)(((per_sm == 3 ? 74 : 112) * 1024 - fixed) / stage_bytes);
    if (stages < 2 && per_sm == 3) {
        per_sm = 2;
        stages = (int)((112 * 1024 - fixed) / stage_bytes);
    }
    if (stages < 2) {
        per_sm = 1;
        stages = (int)((220 * 1024 - fixed) / stage_bytes);
    }
    stages = std::min(stages, 8);
    if (stages < 2) return TNC_ERR_UNSUPPORTED;
    p.stages = stages;
    size_t off = (size_t)stages * stage_bytes;
    p.off_b = (uint32_t)off;
    off += (b_bytes + 15) & ~(size_t)15;
    p.off_koff = (uint32_t)off;
    off += 64;
    p.off_bars = (uint32_t)off;
    off += 16 * 8;
    p.off_staging = (uint32_t)off;
    off += staging;
    const int64_t items = (int64_t)p.n_seg << (e.n_m - 8);
    const int grid = (int)std::min<int64_t>(items, (int64_t)sm_count() * per_sm);
    switch (e.n_n) {
        case 0: return launch_bulk_k<1>(p, off, grid, per_sm, s);
        case 1: return launch_bulk_k<2>(p, off, grid, per_sm, s);
        case 2: return launch_bulk_k<4>(p, off, grid, per_sm, s);
        case 3: return launch_bulk_k<8>(p, off, grid, per_sm, s);
        default: return launch_bulk_k<16>(p, off, grid, per_sm, s);
    }
}

}  // namespace

bool stem_supported(const tnc_einsum& e, int dtype) {
    if (dtype != TNC_C64 || e.n_h != 0) return false;
    if (e.n_k > 12 || e.n_n > 12) return false;
    if (((size_t)8 << (e.n_k + e.n_n)) + ((size_t)4 << e.n_k) > 60 * 1024) return false;   // B[k][n] + k offsets in smem
    for (int i = 0; i < e.n_n; ++i)
        if (e.n_c[i] >= e.n_n) return false;           // output must be C[rows][m][n]
    return true;
}

int launch_stem(const tnc_einsum& e, const void* a, const void* b, void* c, const int32_t* dev_rows_a,
                const int32_t* dev_rows_b, const int32_t* dev_seg_begin, int n_seg, cudaStream_t s, uint32_t* amax_out) {
    if (!stem_supported(e, TNC_C64)) {
        set_error("stem einsum: unsupported shape or output layout (k=%d n=%d h=%d)", e.n_k, e.n_n, e.n_h);
        return TNC_ERR_UNSUPPORTED;
    }
    static const bool no_bulk = knob("TNC_STEM_NO_BULK") != nullptr;      // measurement aid
    if (!no_bulk && bulk_applies(e)) {
        const int rc = launch_stem_bulk(e, a, b, c, dev_rows_a, dev_rows_b, dev_seg_begin, n_seg, amax_out, s);
        if (rc != TNC_ERR_UNSUPPORTED) return rc;
    }
    StemParams p{};
    p.amax_out = amax_out;
    p.a = (const float2*)a;
    p.b = (const float2*)b;
    p.c = (float2*)c;
    p.rows_a = dev_rows_a;
    p.rows_b = dev_rows_b;
    p.rows_mode_a = e.rows_a;
    p.rows_mode_b = e.rows_b;
    p.nbatch = e.nb;
    p.fold = 1;
    // Right-operand rows folded into the row loop (A is then read from HBM once, not once per row
    // of B): a plain step whose rows come from B alone, or a full outer step (all row pairs,
    // A-major: validated when the operation was added).
    const size_t b_bytes = (size_t)8 << (e.n_k + e.n_n);
    if (e.nb > 1 && e.rows_a == TNC_ROWS_NONE && e.rows_b == TNC_ROWS_IDENTITY && e.nb * b_bytes <= 32 * 1024) {
        p.fold = e.nb;
        p.nbatch = 1;
    } else if ((e.flags & TNC_EINSUM_OUTER_ROWS) && e.b.rows > 1 && e.b.rows * b_bytes <= 32 * 1024) {
        p.fold = e.b.rows;
        p.nbatch = e.a.rows;
        p.rows_mode_a = TNC_ROWS_IDENTITY;                                // outer batch = row of A
    }
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
    p.a_vec = e.n_k >= 1 && e.k_a[0] == 0;
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
    const size_t smem = p.fold * b_bytes + ((size_t)4 << e.n_k) + 16 + (kStemThreads / 32) * 4096;
    const int64_t tiles = ((((int64_t)1 << e.n_m) + kStemThreads - 1) / kStemThreads) * p.nbatch;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count() * 4 * 4);
    switch (e.n_n) {
        case 0: return launch_k<1>(p, smem, grid, s);
        case 1: return launch_k<2>(p, smem, grid, s);
        case 2: return launch_k<4>(p, smem, grid, s);
        case 3: return launch_k<8>(p, smem, grid, s);
        default: return launch_k<16>(p, smem, grid, s);
    }
}

}  // namespace tnc
