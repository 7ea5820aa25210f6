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
};

__device__ __forceinline__ void cfma(float2& acc, const float2 a, const float2 b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}

// NCH: outputs (complex) a thread produces per pass over its row; KCH: row amplitudes per chunk.
// Light variants (few accumulators) are compiled for 4 resident CTAs per SM: their rows are short,
// and the bytes in flight per SM are what bounds them.
template <int NCH, int KCH>
__global__ void __launch_bounds__(kStemThreads, (NCH * KCH <= 32 && NCH <= 8 ? 4 : 2)) stem_kernel(const StemParams p) {
    extern __shared__ __align__(16) unsigned char stem_smem[];
    const int K = 1 << p.kb, N = 1 << p.nb;
    float2* Bs = (float2*)stem_smem;                    // [fold][K][N]
    uint32_t* koff = (uint32_t*)(Bs + (size_t)p.fold * K * N);   // A offset of contracted index k
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
                    Bs[(size_t)f * K * N + e] = bsrc[o];
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
        const float2* __restrict__ Bf = Bs + (size_t)f * K * N;
        float2* __restrict__ cp = p.c + ((((int64_t)b * p.fold + f) << p.mb) + r) * N;
#pragma unroll 1
        for (int n0 = 0; n0 < N; n0 += NCH) {
            float2 acc[NCH];
#pragma unroll
            for (int i = 0; i < NCH; ++i) acc[i] = make_float2(0.f, 0.f);
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
                    const float2* brow = Bf + (size_t)(k0 + j) * N + n0;
                    if constexpr (NCH == 1) {
                        cfma(acc[0], av[j], brow[0]);
                    } else {
#pragma unroll
                        for (int i = 0; i < NCH; i += 2) {
                            const float4 bb = *(const float4*)(brow + i);      // warp-uniform: broadcast
                            cfma(acc[i], av[j], make_float2(bb.x, bb.y));
                            cfma(acc[i + 1], av[j], make_float2(bb.z, bb.w));
                        }
                    }
                }
            }
            if constexpr (NCH == 1) {
                cp[n0] = acc[0];
            } else {
                if (staged) {
                    store_rows_coalesced<NCH / 2>(stage, (const float*)acc, (float*)(cp - (int64_t)lane * N + n0), 2 * (int64_t)N, lane);
                } else {
#pragma unroll
                    for (int i = 0; i < NCH; i += 2)
                        *(float4*)(cp + n0 + i) = make_float4(acc[i].x, acc[i].y, acc[i + 1].x, acc[i + 1].y);
                }
            }
        }
        }
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
                const int32_t* dev_rows_b, cudaStream_t s) {
    if (!stem_supported(e, TNC_C64)) {
        set_error("stem einsum: unsupported shape or output layout (k=%d n=%d h=%d)", e.n_k, e.n_n, e.n_h);
        return TNC_ERR_UNSUPPORTED;
    }
    StemParams p{};
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
