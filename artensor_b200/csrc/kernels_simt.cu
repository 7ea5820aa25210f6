// CUDA-core kernels of tnc_b200: the generic bit-einsum (any shape; used for the many tiny
// steps of a scheme and as the always-available algorithm), leaf slicing, the bit-permutation
// copy and the slice accumulator.  HBM-bound byte movers: coalesced, vectorised where the
// permutation allows, grids sized in multiples of the SM count.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "tnc_internal.h"

namespace tnc {

namespace {

constexpr int kThreads = 256;

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

inline int grid_for(int64_t work_items, int per_sm = 8) {
    int64_t blocks = (work_items + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;          // grid-stride beyond: a multiple of the SM count
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T> struct Cplx;
template <> struct Cplx<float2> {
    static __device__ __forceinline__ float2 load(const float2* p) { return *p; }
    static __device__ __forceinline__ void store(float2* p, float2 v) { *p = v; }
};

// ------------------------------------------------------------------ generic einsum
// One thread per output element.  The output index is split into its bits; every bit that
// also addresses A (m/h modes) or B (n/h modes) is moved to its position there.  The
// contracted index walks two precomputed offset tables (deposit of k into A / B positions).
template <typename T>
__global__ void __launch_bounds__(kThreads) simt_einsum_kernel(SimtEinsumParams p) {
    const T* __restrict__ A = (const T*)p.a;
    const T* __restrict__ B = (const T*)p.b;
    T* __restrict__ C = (T*)p.c;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint32_t cmask = p.rank_c >= 32 ? 0xffffffffu : ((1u << p.rank_c) - 1u);
    const uint32_t nk = 1u << p.kb;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.total; e += stride) {
        const int64_t row = e >> p.rank_c;
        const uint32_t cb = (uint32_t)e & cmask;
        uint32_t oa = 0, ob = 0;
        for (int q = 0; q < p.rank_c; ++q) {
            const uint32_t bit = (cb >> q) & 1u;
            const int pa = p.c2a[q], pb = p.c2b[q];
            if (pa >= 0) oa |= bit << pa;
            if (pb >= 0) ob |= bit << pb;
        }
        int64_t ra = 0, rb = 0;
        if (p.rows_mode_a == TNC_ROWS_IDENTITY) ra = row;
        else if (p.rows_mode_a >= 0) ra = p.rows_a[row];
        if (p.rows_mode_b == TNC_ROWS_IDENTITY) rb = row;
        else if (p.rows_mode_b >= 0) rb = p.rows_b[row];
        const T* __restrict__ a = A + (ra << p.rank_a) + oa;
        const T* __restrict__ b = B + (rb << p.rank_b) + ob;
        float cr = 0.f, ci = 0.f;
        for (uint32_t k = 0; k < nk; ++k) {
            const float2 x = Cplx<T>::load(a + p.koff_a[k]);
            const float2 y = Cplx<T>::load(b + p.koff_b[k]);
            cr = fmaf(x.x, y.x, cr);
            cr = fmaf(-x.y, y.y, cr);
            ci = fmaf(x.x, y.y, ci);
            ci = fmaf(x.y, y.x, ci);
        }
        Cplx<T>::store(C + e, make_float2(cr, ci));
    }
}

// ------------------------------------------------------------------ long contraction, few outputs per row
// The tail of a sparse scheme: batched steps [rows][m][k] x [rows][k][n] with a long contraction
// (k ~ 2^7..2^13) and at most 16 outputs per row.  One thread per output (the generic kernel) walks
// K alone with uncoalesced loads; here one CTA owns an output row, its threads split K (consecutive
// threads take consecutive k: the contracted bits are ordered by their position in A, so the loads
// of A are coalesced), every thread keeps all 2^RC partial sums, and the CTA reduces them.
// THREADS per CTA: 128 when there are many rows (many CTAs per SM stream together); 256 for a few hundred rows with a
// very long contraction -- one 128-thread CTA per row would leave ~7 warps per SM, far too few loads in flight for
// HBM (n53 m20 sc31_s20 tree: [256 rows][3 bits][19 bits] x [256][19 bits][2 bits], 12.9 GB: 2.0 -> 3.4-3.7 TB/s
// inside a slice; 512 threads measured the same to -8 %, and 64 outputs per row do not fit their 128 registers).
// NM / NN: left-only / right-only output bits (NM + NN <= 4, or NM, NN <= 3: up to 64 outputs per row --
// the batched tail of a deep sparse scheme, e.g. [512 rows][3 bits][19 bits] x [512][19 bits][3 bits] in the
// sc_target-32 n53 tree: 34 GB streamed once; the generic kernel ran it at 0.2 TB/s); no shared kept modes
template <int NM, int NN, int THREADS>
__global__ void __launch_bounds__(THREADS) simt_rowdot_kernel(SimtEinsumParams p) {
    constexpr int M = 1 << NM, NQ = 1 << NN, RC = NM + NN;
    constexpr int kRowdotThreads = THREADS;
    __shared__ float2 part[kRowdotThreads / 32][M * NQ];
    const float2* __restrict__ A = (const float2*)p.a;
    const float2* __restrict__ B = (const float2*)p.b;
    float2* __restrict__ C = (float2*)p.c;
    // offsets of the M left-only / NQ right-only index values in A / B, and of output (mi, ni) in C
    uint32_t oa[M], ob[NQ], cm[M], cn[NQ];
#pragma unroll
    for (int i = 0; i < M; ++i) oa[i] = cm[i] = 0;
#pragma unroll
    for (int i = 0; i < NQ; ++i) ob[i] = cn[i] = 0;
    {
        int im = 0, in = 0;
#pragma unroll
        for (int q = 0; q < RC; ++q) {
            if (p.c2a[q] >= 0) {
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    oa[i] |= (((uint32_t)i >> im) & 1u) << p.c2a[q];
                    cm[i] |= (((uint32_t)i >> im) & 1u) << q;
                }
                ++im;
            } else {
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    ob[i] |= (((uint32_t)i >> in) & 1u) << p.c2b[q];
                    cn[i] |= (((uint32_t)i >> in) & 1u) << q;
                }
                ++in;
            }
        }
    }
    const uint32_t nk = 1u << p.kb;
    const int64_t rows = p.total >> RC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        int64_t ra = 0, rb = 0;
        if (p.rows_mode_a == TNC_ROWS_IDENTITY) ra = row;
        else if (p.rows_mode_a >= 0) ra = p.rows_a[row];
        if (p.rows_mode_b == TNC_ROWS_IDENTITY) rb = row;
        else if (p.rows_mode_b >= 0) rb = p.rows_b[row];
        const float2* __restrict__ a = A + (ra << p.rank_a);
        const float2* __restrict__ b = B + (rb << p.rank_b);
        float2 acc[M][NQ];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < NQ; ++j) acc[i][j] = make_float2(0.f, 0.f);
        if constexpr (NQ >= 2) {
            // packed fp32x2 (sm_100): two right-only outputs per instruction, planar (re, re) / (im, im)
            // accumulators, the same four fused multiply-adds per output in the same order
            float2 accR[M][NQ / 2], accI[M][NQ / 2];
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j < NQ / 2; ++j) accR[i][j] = accI[i][j] = make_float2(0.f, 0.f);
            for (uint32_t k = threadIdx.x; k < nk; k += kRowdotThreads) {
                const uint32_t ka = p.koff_a[k], kb = p.koff_b[k];
                float2 x[M], y[NQ];
#pragma unroll
                for (int i = 0; i < M; ++i) x[i] = a[oa[i] + ka];
#pragma unroll
                for (int j = 0; j < NQ; ++j) y[j] = b[ob[j] + kb];
#pragma unroll
                for (int j = 0; j < NQ / 2; ++j) {
                    const float2 yr = make_float2(y[2 * j].x, y[2 * j + 1].x), yi = make_float2(y[2 * j].y, y[2 * j + 1].y);
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        accR[i][j] = __ffma2_rn(make_float2(x[i].x, x[i].x), yr, accR[i][j]);
                        accR[i][j] = __ffma2_rn(make_float2(-x[i].y, -x[i].y), yi, accR[i][j]);
                        accI[i][j] = __ffma2_rn(make_float2(x[i].x, x[i].x), yi, accI[i][j]);
                        accI[i][j] = __ffma2_rn(make_float2(x[i].y, x[i].y), yr, accI[i][j]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < M; ++i)
#pragma unroll
                for (int j = 0; j < NQ / 2; ++j) {
                    acc[i][2 * j] = make_float2(accR[i][j].x, accI[i][j].x);
                    acc[i][2 * j + 1] = make_float2(accR[i][j].y, accI[i][j].y);
                }
        } else {
            for (uint32_t k = threadIdx.x; k < nk; k += kRowdotThreads) {
                const uint32_t ka = p.koff_a[k], kb = p.koff_b[k];
                float2 x[M], y[NQ];
#pragma unroll
                for (int i = 0; i < M; ++i) x[i] = a[oa[i] + ka];
#pragma unroll
                for (int j = 0; j < NQ; ++j) y[j] = b[ob[j] + kb];
#pragma unroll
                for (int i = 0; i < M; ++i)
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        acc[i][j].x = fmaf(x[i].x, y[j].x, acc[i][j].x);
                        acc[i][j].x = fmaf(-x[i].y, y[j].y, acc[i][j].x);
                        acc[i][j].y = fmaf(x[i].x, y[j].y, acc[i][j].y);
                        acc[i][j].y = fmaf(x[i].y, y[j].x, acc[i][j].y);
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    acc[i][j].x += __shfl_xor_sync(0xffffffffu, acc[i][j].x, d);
                    acc[i][j].y += __shfl_xor_sync(0xffffffffu, acc[i][j].y, d);
                }
                if (lane == 0) part[warp][i * NQ + j] = acc[i][j];
            }
        __syncthreads();
        if (threadIdx.x < M * NQ) {
            float2 o = part[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < kRowdotThreads / 32; ++w) {
                o.x += part[w][threadIdx.x].x;
                o.y += part[w][threadIdx.x].y;
            }
            // thread t holds output (mi, ni) = (t / NQ, t % NQ); its place in C comes from the bit tables
            uint32_t off = 0;
#pragma unroll
            for (int i = 0; i < M; ++i)
                if ((int)threadIdx.x / NQ == i) off |= cm[i];
#pragma unroll
            for (int j = 0; j < NQ; ++j)
                if ((int)threadIdx.x % NQ == j) off |= cn[j];
            C[(row << RC) + off] = o;
        }
        __syncthreads();
    }
}

// (nm, nn) combinations compiled: nm + nn <= 4, and every nm, nn <= 3
constexpr bool rowdot_shape(int nm, int nn) { return nm >= 0 && nn >= 0 && (nm + nn <= 4 || (nm <= 3 && nn <= 3)); }

template <int NM, int THREADS>
void launch_rowdot_n(const SimtEinsumParams& p, int nn, int grid, cudaStream_t s) {
    if constexpr (rowdot_shape(NM, 0) && (THREADS < 512 || NM + 0 <= 5)) {
        if (nn == 0) simt_rowdot_kernel<NM, 0, THREADS><<<grid, THREADS, 0, s>>>(p);
    }
    if constexpr (rowdot_shape(NM, 1) && (THREADS < 512 || NM + 1 <= 5)) {
        if (nn == 1) simt_rowdot_kernel<NM, 1, THREADS><<<grid, THREADS, 0, s>>>(p);
    }
    if constexpr (rowdot_shape(NM, 2) && (THREADS < 512 || NM + 2 <= 5)) {
        if (nn == 2) simt_rowdot_kernel<NM, 2, THREADS><<<grid, THREADS, 0, s>>>(p);
    }
    if constexpr (rowdot_shape(NM, 3) && (THREADS < 512 || NM + 3 <= 5)) {
        if (nn == 3) simt_rowdot_kernel<NM, 3, THREADS><<<grid, THREADS, 0, s>>>(p);
    }
    if constexpr (rowdot_shape(NM, 4) && (THREADS < 512 || NM + 4 <= 5)) {
        if (nn == 4) simt_rowdot_kernel<NM, 4, THREADS><<<grid, THREADS, 0, s>>>(p);
    }
}
template <int THREADS>
void launch_rowdot(const SimtEinsumParams& p, int nm, int nn, int grid, cudaStream_t s) {
    switch (nm) {
        case 0: launch_rowdot_n<0, THREADS>(p, nn, grid, s); break;
        case 1: launch_rowdot_n<1, THREADS>(p, nn, grid, s); break;
        case 2: launch_rowdot_n<2, THREADS>(p, nn, grid, s); break;
        case 3: launch_rowdot_n<3, THREADS>(p, nn, grid, s); break;
        default: launch_rowdot_n<4, THREADS>(p, nn, grid, s); break;
    }
}

// ------------------------------------------------------------------ chain of tiny einsums
// The same arithmetic as simt_einsum_kernel for a run of consecutive tiny steps, in one launch:
// one CTA of 1024 threads walks the run, `__syncthreads()` between steps makes a step's outputs
// (global memory, written by this CTA) visible to the next one.  Operands are read with ld.cg:
// a step may read what an earlier step of the same launch wrote, so the non-coherent path is out.
constexpr int kChainThreads = 1024;
__global__ void __launch_bounds__(kChainThreads) simt_chain_kernel(const ChainStep* __restrict__ steps, int n, char* ws,
                                                                   const char* __restrict__ blob) {
    __shared__ ChainStep recs[2];              // double buffer: step s + 1 is fetched while step s computes
    static_assert(sizeof(ChainStep) % 4 == 0, "copied word by word");
    constexpr int kWords = (int)(sizeof(ChainStep) / 4);
    if (threadIdx.x < kWords) ((uint32_t*)&recs[0])[threadIdx.x] = ((const uint32_t*)steps)[threadIdx.x];
    __syncthreads();
    for (int s = 0; s < n; ++s) {
        const ChainStep& p = recs[s & 1];
        // the last warp fetches the next record; the barrier at the end of the step publishes it together
        // with this step's outputs
        if (s + 1 < n && threadIdx.x >= kChainThreads - kWords)
            ((uint32_t*)&recs[(s + 1) & 1])[threadIdx.x - (kChainThreads - kWords)] =
                ((const uint32_t*)(steps + s + 1))[threadIdx.x - (kChainThreads - kWords)];
        const float2* A = (const float2*)(ws + p.a_off);
        const float2* B = (const float2*)(ws + p.b_off);
        float2* C = (float2*)(ws + p.c_off);
        const int32_t* rows_a = (const int32_t*)(blob + p.rows_a_off);
        const int32_t* rows_b = (const int32_t*)(blob + p.rows_b_off);
        const uint32_t* koff_a = (const uint32_t*)(blob + p.koff_a_off);
        const uint32_t* koff_b = (const uint32_t*)(blob + p.koff_b_off);
        const uint32_t cmask = p.rank_c >= 32 ? 0xffffffffu : ((1u << p.rank_c) - 1u);
        const uint32_t nk = 1u << p.kb;
        for (int64_t e = threadIdx.x; e < p.total; e += kChainThreads) {
            const int64_t row = e >> p.rank_c;
            const uint32_t cb = (uint32_t)e & cmask;
            uint32_t oa = 0, ob = 0;
            for (int q = 0; q < p.rank_c; ++q) {
                const uint32_t bit = (cb >> q) & 1u;
                const int pa = p.c2a[q], pb = p.c2b[q];
                if (pa >= 0) oa |= bit << pa;
                if (pb >= 0) ob |= bit << pb;
            }
            int64_t ra = 0, rb = 0;
            if (p.rows_mode_a == TNC_ROWS_IDENTITY) ra = row;
            else if (p.rows_mode_a >= 0) ra = TNC_LDG(rows_a + row);          // the plan's tables never change: cached path
            if (p.rows_mode_b == TNC_ROWS_IDENTITY) rb = row;
            else if (p.rows_mode_b >= 0) rb = TNC_LDG(rows_b + row);
            const float2* a = A + (ra << p.rank_a) + oa;
            const float2* b = B + (rb << p.rank_b) + ob;
            float cr = 0.f, ci = 0.f;
#pragma unroll 4
            for (uint32_t k = 0; k < nk; ++k) {
                const float2 x = __ldcg(a + TNC_LDG(koff_a + k));
                const float2 y = __ldcg(b + TNC_LDG(koff_b + k));
                cr = fmaf(x.x, y.x, cr);
                cr = fmaf(-x.y, y.y, cr);
                ci = fmaf(x.x, y.y, ci);
                ci = fmaf(x.y, y.x, ci);
            }
            C[e] = make_float2(cr, ci);
        }
        __syncthreads();                       // step s is complete (and record s + 1 is in place) for everybody
    }
}

// ------------------------------------------------------------------ leaf slicing
// One block per leaf; leaves are tiny (rank <= ~6).  dst[r][e] = src[r][deposit(e) | fixed].
template <typename T>
__global__ void __launch_bounds__(128) leaf_gather_kernel(const LeafDev* __restrict__ leaves,
                                                          const T* __restrict__ blob, char* arena,
                                                          uint64_t slice_id, const uint64_t* slice_word) {
    // graph-replayed slices read their id from a word in the workspace that the graph's last node bumps
    if (slice_word) slice_id = *slice_word;
    const LeafDev L = leaves[blockIdx.x];
    T* dst = (T*)(arena + L.dst_offset);
    uint32_t fixed = 0;
    for (int s = 0; s < L.n_sliced; ++s)
        fixed |= (uint32_t)((slice_id >> L.sliced_shift[s]) & 1ull) << L.sliced_pos[s];
    const int64_t total = (int64_t)L.dst_rows << L.dst_rank;
    const uint32_t mask = (1u << L.dst_rank) - 1u;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
        const int64_t row = e >> L.dst_rank;
        const uint32_t eb = (uint32_t)e & mask;
        uint32_t so = fixed;
        for (int q = 0; q < L.dst_rank; ++q) so |= ((eb >> q) & 1u) << L.keep_pos[q];
        dst[e] = blob[L.src_offset + (row << L.src_rank) + so];
    }
}

// ------------------------------------------------------------------ bit permutation (v1)
// Thread per destination element, grid-stride; destination writes are fully coalesced, the
// source side is coalesced for as many low positions as the permutation keeps in place.
template <typename T>
__global__ void __launch_bounds__(kThreads) permute_kernel(PermuteParams p) {
    const T* __restrict__ src = (const T*)p.src;
    T* __restrict__ dst = (T*)p.dst;
    const int64_t total = p.rows << p.rank;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t mask = (1ull << p.rank) - 1ull;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint64_t q = (uint64_t)e & mask;
        uint64_t s = 0;
        for (int i = 0; i < p.rank; ++i) s |= ((q >> i) & 1ull) << p.perm[i];
        dst[e] = src[((e >> p.rank) << p.rank) + (int64_t)s];
    }
}

// ------------------------------------------------------------------ accumulate
template <typename T>
__global__ void __launch_bounds__(kThreads) accum_kernel(AccumParams p) {
    const T* __restrict__ src = (const T*)p.src;
    float2* __restrict__ out = (float2*)p.out;
    const int64_t total = p.rows << p.rank;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t mask = (1ull << p.rank) - 1ull;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const uint64_t q = (uint64_t)e & mask;
        uint64_t d = 0;
        for (int i = 0; i < p.rank; ++i) d |= ((q >> i) & 1ull) << p.out_pos[i];
        const float2 v = Cplx<T>::load(src + e);
        float2* o = out + ((e >> p.rank) << p.rank) + (int64_t)d;
        float2 cur = *o;
        cur.x += v.x;
        cur.y += v.y;
        *o = cur;
    }
}

}  // namespace

// Whether the generic path runs a step with the row-dot kernel (many rows, a long contraction, <= 64 outputs per
// row, no shared kept modes: one CTA per row whose threads split K).  Such a step is never made part of a chain
// launch (tnc_plan_finalize): it is summed in the same order however the steps around it are grouped.
bool simt_uses_rowdot(int rank_c, int kb, int64_t total, int n_m, int n_n, int n_h) {
    static const bool no_rowdot = knob("TNC_NO_ROWDOT") != nullptr;      // measurement aid
    return !no_rowdot && rank_c <= 6 && kb >= 7 && (total >> rank_c) >= 32 && n_h == 0 && rowdot_shape(n_m, n_n);
}

int launch_simt_einsum(const SimtEinsumParams& p, int dtype, cudaStream_t s) {
    if (p.total <= 0) return TNC_OK;
    {
        int nm = 0, nn = 0, nh = 0;
        for (int q = 0; q < p.rank_c; ++q) {
            if (p.c2a[q] >= 0 && p.c2b[q] >= 0) ++nh;
            else if (p.c2a[q] >= 0) ++nm;
            else ++nn;
        }
        if (dtype == TNC_C64 && simt_uses_rowdot(p.rank_c, p.kb, p.total, nm, nn, nh)) {
            const int64_t rows = p.total >> p.rank_c;
            const int grid = (int)std::min<int64_t>(rows, (int64_t)sm_count() * 16);
            // few rows x a very long contraction: wide CTAs (the number of CTAs is bounded by the rows)
            // (64 outputs per row: 128 accumulator registers -- 256 threads, which may use up to 255 registers each)
            static const int forced = knob("TNC_ROWDOT_THREADS") ? atoi(knob("TNC_ROWDOT_THREADS")) : 0;   // experiment knob
            if (rows < (int64_t)sm_count() * 8 && p.kb >= 12 && forced != 128) launch_rowdot<256>(p, nm, nn, grid, s);
            else launch_rowdot<128>(p, nm, nn, grid, s);
            TNC_CUDA(cudaGetLastError());
            return TNC_OK;
        }
    }
    const int grid = grid_for(p.total);
    if (dtype != TNC_C64) {
        set_error("einsum: only complex64 tensors are supported");
        return TNC_ERR_UNSUPPORTED;
    }
    simt_einsum_kernel<float2><<<grid, kThreads, 0, s>>>(p);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

int launch_simt_chain(const ChainStep* dev_steps, int n, void* workspace, const void* dev_blob, cudaStream_t s) {
    if (n <= 0) return TNC_OK;
    simt_chain_kernel<<<1, kChainThreads, 0, s>>>(dev_steps, n, (char*)workspace, (const char*)dev_blob);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

__global__ void slice_word_kernel(uint64_t* word, uint64_t value, int add) {
    *word = add ? *word + value : value;
}

int launch_slice_word(uint64_t* word, uint64_t value, bool add, cudaStream_t s) {
    slice_word_kernel<<<1, 1, 0, s>>>(word, value, add ? 1 : 0);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

int launch_leaf_gather(const LeafDev* dev_leaves, int n, int max_elems, const void* blob,
                       void* arena, uint64_t slice_id, const uint64_t* slice_word, int dtype, cudaStream_t s) {
    (void)max_elems;
    if (n <= 0) return TNC_OK;
    if (dtype != TNC_C64) {
        set_error("leaves: only complex64 tensors are supported");
        return TNC_ERR_UNSUPPORTED;
    }
    leaf_gather_kernel<float2><<<n, 128, 0, s>>>(dev_leaves, (const float2*)blob, (char*)arena, slice_id, slice_word);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

int launch_permute(const PermuteParams& p, int elem_bytes, cudaStream_t s) {
    const int64_t total = p.rows << p.rank;
    if (total <= 0) return TNC_OK;
    const int grid = grid_for(total);
    if (elem_bytes == 8) permute_kernel<float2><<<grid, kThreads, 0, s>>>(p);
    else if (elem_bytes == 4) permute_kernel<float><<<grid, kThreads, 0, s>>>(p);
    else {
        set_error("permute: elem_bytes must be 4 or 8, got %d", elem_bytes);
        return TNC_ERR_INVALID;
    }
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

int launch_accum(const AccumParams& p, int dtype, cudaStream_t s) {
    const int64_t total = p.rows << p.rank;
    if (total <= 0) return TNC_OK;
    const int grid = grid_for(total);
    if (dtype != TNC_C64) {
        set_error("accumulate: only complex64 tensors are supported");
        return TNC_ERR_UNSUPPORTED;
    }
    accum_kernel<float2><<<grid, kThreads, 0, s>>>(p);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

}  // namespace tnc
