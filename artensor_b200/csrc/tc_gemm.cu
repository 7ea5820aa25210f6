// Tensor-core lowering of one scheme step (artensor/contraction.py:70, :147-190 call sites):
//
//   C[b][m, n] = sum_k A[ra[b]][m, k] * B[rb[b]][k, n]            (complex64)
//
// is executed as ONE real GEMM on the 5th-generation tensor cores.  complex64 storage is
// interleaved, so a K-major complex panel A[M][K] *is* the real matrix A'[M][2K]; with
//   B'[2n  ][2k] =  Br(k,n)   B'[2n  ][2k+1] = -Bi(k,n)
//   B'[2n+1][2k] =  Bi(k,n)   B'[2n+1][2k+1] =  Br(k,n)
// the product A' * B'^T is C in interleaved complex form (the "4M" formulation, no extra flops).
// fp32 accuracy comes from the 3xTF32 split: x = hi + lo (hi = tf32(x), lo = tf32(x - hi)) and
//   A'B' ~= lo*hi + hi*lo + hi*hi.
// The tensor core's fp32 accumulator rounds toward zero on every tcgen05.mma (measured on B200:
// the error of a K = 32768 product grows linearly with K, 7e-4 of the rms), so tensor memory
// only ever holds the sum of a short chunk (one k-block = 12 MMAs); the epilogue warps add the
// chunks into fp32 registers with round-to-nearest while the next chunk is being computed in the
// other half of tensor memory.
//
//
// Operand precisions (TcPrecision, tc_gemm.h) -- all accumulate in fp32:
//   3xTF32   hi/lo are TF32 (11 + 11 significant bits), kind::tf32, 3 MMAs per useful one.
//   3xF16    hi/lo are fp16 (11 + 11 significant bits, the same operand precision), kind::f16:
//            the tensor pipe runs fp16 at twice the TF32 rate and the panels are half as large.
//            fp16 has a 5-bit exponent, so each operand is first multiplied by a power of two
//            that brings its largest magnitude into [2^14, 2^15) (amax_kernel finds it, the pack
//            kernel applies it, the GEMM epilogue undoes it -- all exact); elements more than
//            2^17 below the maximum lose low bits to fp16 subnormals, an absolute error below
//            2^-39 of the maximum.
//   F16      hi only: the reduced-precision complex-half mode (Pan et al. 2023), 1 MMA per
//            useful one.
//
// Kernels in this file
//   amax_kernel          largest |re|, |im| of both operands (one launch), fp16 precisions only.
//   pack_kernel          bit-permutation of an operand into its K-major panel(s), tiled through
//                        shared memory with an XOR swizzle (coalesced reads AND writes, no bank
//                        conflicts); also does the hi/lo split and the B' expansion.  HBM bound.
//   gemm_kernel          warp-specialised: warp 0 = TMA producer (cp.async.bulk.tensor, 128B
//   gemm_2cta_kernel     swizzle), warp 1 = tcgen05.mma issuer (two accumulators in TMEM),
//                        warps 2-9 = chunk accumulation + epilogue (tcgen05.ld -> fp32
//                        registers -> global).  Tensor bound.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "tc_gemm.h"
#include "tc_common.cuh"

namespace tnc {

namespace {

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

// =====================================================================================
// pack kernel
// =====================================================================================
constexpr int kPackThreads = 256;
constexpr int kPackMaxTileBits = 10;

struct PackParams {
    const float2* src;
    float2* dst_hi;
    float2* dst_lo;
    const int32_t* rows;
    const uint32_t* amax;                    // fp16 modes: bits of the operand's largest magnitude
    int32_t rows_mode;
    int32_t rank, tbits, mode, inner_bits, nb, n_outer, n_xor, blocked, bn_log2, kb_log2;
    int64_t n_tiles;
    int8_t tile_src_pos[kPackMaxTileBits];   // source position of tile bit j in SOURCE order (ascending)
    int8_t tile_u2v[kPackMaxTileBits];       // destination-order tile bit that source-order bit j is
    int8_t tile_dst_pos[kPackMaxTileBits];   // destination position of tile bit j in DESTINATION order
    int8_t xor_from[kPackMaxTileBits], xor_to[kPackMaxTileBits];
    int8_t outer_dst_pos[TNC_MAX_BITS], outer_src_pos[TNC_MAX_BITS];
};

__device__ __forceinline__ float tf32_round(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Largest |component| of two tensors in one launch: blocks [0, split) reduce `a`, the rest `b`;
// out[0] / out[1] (zeroed before the launch) receive the float bits (non-negative floats order
// like unsigned integers).
__global__ void __launch_bounds__(256) amax_kernel(const float4* __restrict__ a, int64_t n4_a, const float4* __restrict__ b,
                                                   int64_t n4_b, int split, uint32_t* out) {
    const bool second = (int)blockIdx.x >= split;
    const float4* __restrict__ src = second ? b : a;
    const int64_t n4 = second ? n4_b : n4_a;
    const int64_t nblk = second ? (int64_t)gridDim.x - split : split;
    const int64_t blk = second ? (int64_t)blockIdx.x - split : blockIdx.x;
    float m0 = 0.f, m1 = 0.f;
    int64_t i = blk * 256 + threadIdx.x;
    const int64_t stride = nblk * 256;
    for (; i + stride < n4; i += 2 * stride) {          // two independent loads in flight
        const float4 x = src[i], y = src[i + stride];
        m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
        m1 = fmaxf(m1, fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w))));
    }
    if (i < n4) {
        const float4 x = src[i];
        m0 = fmaxf(m0, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
    }
    uint32_t bits = __float_as_uint(fmaxf(m0, m1));
    bits = __reduce_max_sync(0xffffffffu, bits);
    __shared__ uint32_t warp_max[8];
    if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = bits;
    __syncthreads();
    if (threadIdx.x < 8) {
        bits = __reduce_max_sync(0xffu, warp_max[threadIdx.x]);
        if (threadIdx.x == 0 && bits) atomicMax(out + (second ? 1 : 0), bits);
    }
}

__global__ void __launch_bounds__(kPackThreads) pack_kernel(const PackParams p) {
    extern __shared__ __align__(16) unsigned char pack_smem[];
    const int tile = 1 << p.tbits;
    float2* data = (float2*)pack_smem;
    uint32_t* src_off = (uint32_t*)(data + tile);
    uint32_t* dst_off = src_off + tile;
    uint16_t* slot_u = (uint16_t*)(dst_off + tile);   // swizzled smem slot of source-order element u
    uint16_t* slot_v = slot_u + tile;                 // swizzled smem slot of destination-order element v
    for (int x = threadIdx.x; x < tile; x += kPackThreads) {
        uint32_t so = 0, dofs = 0, v = 0;
        for (int j = 0; j < p.tbits; ++j) {
            const uint32_t bit = (x >> j) & 1u;
            so |= bit << p.tile_src_pos[j];
            v |= bit << p.tile_u2v[j];
            dofs |= bit << p.tile_dst_pos[j];
        }
        uint32_t sv = v, sx = x;
        for (int j = 0; j < p.n_xor; ++j) {
            sv ^= ((v >> p.xor_from[j]) & 1u) << p.xor_to[j];
            sx ^= (((uint32_t)x >> p.xor_from[j]) & 1u) << p.xor_to[j];
        }
        src_off[x] = so;
        dst_off[x] = dofs;
        slot_u[x] = (uint16_t)sv;
        slot_v[x] = (uint16_t)sx;
    }
    __syncthreads();
    const int obits = p.rank - p.tbits;
    const int64_t omask = (obits >= 63) ? -1 : (((int64_t)1 << obits) - 1);
    const int64_t kmask = ((int64_t)1 << p.inner_bits) - 1;
    for (int64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const int64_t blk = t >> obits;
        const int64_t o = t & omask;
        int64_t sbase = 0, dbase = 0;
        for (int j = 0; j < p.n_outer; ++j) {
            const int64_t bit = (o >> j) & 1;
            sbase |= bit << p.outer_src_pos[j];
            dbase |= bit << p.outer_dst_pos[j];
        }
        int64_t row = 0;
        if (p.rows_mode == TNC_ROWS_IDENTITY) row = blk;
        else if (p.rows_mode >= 0) row = p.rows[blk];
        const float2* __restrict__ src = p.src + (row << p.rank) + sbase;
        for (int u = threadIdx.x; u < tile; u += kPackThreads) data[slot_u[u]] = src[src_off[u]];
        __syncthreads();
        if (p.mode == PACK_COPY) {
            float2* __restrict__ d = p.dst_hi + (blk << p.rank) + dbase;
            for (int v = threadIdx.x; v < tile; v += kPackThreads) d[dst_off[v]] = data[slot_v[v]];
        } else if (p.mode == PACK_SPLIT) {
            float2* __restrict__ dh = p.dst_hi + (blk << p.rank) + dbase;
            float2* __restrict__ dl = p.dst_lo + (blk << p.rank) + dbase;
            for (int v = threadIdx.x; v < tile; v += kPackThreads) {
                const float2 x = data[slot_v[v]];
                const float2 h = make_float2(tf32_round(x.x), tf32_round(x.y));
                dh[dst_off[v]] = h;
                dl[dst_off[v]] = make_float2(tf32_round(x.x - h.x), tf32_round(x.y - h.y));
            }
        } else if (p.mode == PACK_SPLIT_F16) {
            // A panel in fp16: one __half2 (re, im) per amplitude, scaled by a power of two
            const float sc = f16_scale(*p.amax);
            __half2* __restrict__ dh = (__half2*)p.dst_hi + (blk << p.rank) + dbase;
            __half2* __restrict__ dl = p.dst_lo ? (__half2*)p.dst_lo + (blk << p.rank) + dbase : nullptr;
            for (int v = threadIdx.x; v < tile; v += kPackThreads) {
                const float2 x = data[slot_v[v]];
                const float xr = x.x * sc, xi = x.y * sc;
                const __half2 h = __floats2half2_rn(xr, xi);
                dh[dst_off[v]] = h;
                if (dl) {
                    const float2 hf = __half22float2(h);
                    dl[dst_off[v]] = __floats2half2_rn(xr - hf.x, xi - hf.y);
                }
            }
        } else {
            // B'[2n + c'][2k + c]: in (re, im)-pair units the index is k | c' << inner | n << (inner + 1)
            const bool f16 = p.mode == PACK_EXPAND_SPLIT_F16;
            const int64_t K = (int64_t)1 << p.inner_bits;
            const int kbl = p.kb_log2;                       // complex k per k-block: 16 (tf32) or 32 (fp16)
            const int64_t kbm = ((int64_t)1 << kbl) - 1;
            const float sc = f16 ? f16_scale(*p.amax) : 1.f;
            float2* __restrict__ dh = p.dst_hi + (blk << (p.rank + 1));
            float2* __restrict__ dl = p.dst_lo + (blk << (p.rank + 1));
            __half2* __restrict__ hh = (__half2*)p.dst_hi + (blk << (p.rank + 1));
            __half2* __restrict__ hl = p.dst_lo ? (__half2*)p.dst_lo + (blk << (p.rank + 1)) : nullptr;
            for (int v = threadIdx.x; v < tile; v += kPackThreads) {
                const float2 x = data[slot_v[v]];
                const int64_t q = dbase + dst_off[v];
                int64_t e, second = K;
                if (p.blocked) {
                    // [n tile][k block][2n + c' within the tile][k within the block]: one contiguous block per TMA box
                    const int64_t k = q & kmask, n = q >> p.inner_bits;
                    const int hb = p.bn_log2 - 1;
                    const int64_t row = (n & (((int64_t)1 << hb) - 1)) << 1;
                    const int64_t nkb = K >> kbl;
                    e = (k & kbm) + ((row + ((int64_t)1 << p.bn_log2) * ((k >> kbl) + nkb * (n >> hb))) << kbl);
                    second = kbm + 1;
                } else {
                    e = (q & kmask) | ((q >> p.inner_bits) << (p.inner_bits + 1));
                }
                if (f16) {
                    const float xr = x.x * sc, xi = x.y * sc;
                    const __half hr = __float2half_rn(xr), hi = __float2half_rn(xi);
                    hh[e] = __halves2half2(hr, __hneg(hi));
                    hh[e + second] = __halves2half2(hi, hr);
                    if (hl) {
                        const __half lr = __float2half_rn(xr - __half2float(hr)), li = __float2half_rn(xi - __half2float(hi));
                        hl[e] = __halves2half2(lr, __hneg(li));
                        hl[e + second] = __halves2half2(li, lr);
                    }
                } else {
                    const float hr = tf32_round(x.x), hi = tf32_round(x.y);
                    const float lr = tf32_round(x.x - hr), li = tf32_round(x.y - hi);
                    dh[e] = make_float2(hr, -hi);
                    dh[e + second] = make_float2(hi, hr);
                    dl[e] = make_float2(lr, -li);
                    dl[e + second] = make_float2(li, lr);
                }
            }
        }
        __syncthreads();
    }
}


// -------------------------------------------------------------------------------------
// Fast path of the pack kernel (copy / hi-lo split, no B' expansion).  Same tiling idea -- the
// tile is the union of the lowest destination bits and of the destination bits fed by the lowest
// source bits, so that both the global reads and the global writes are long contiguous runs --
// but built for instruction economy: every thread keeps the offsets and shared-memory slots of
// its 8 elements in registers (no index tables), reads two source-adjacent amplitudes per
// 16-byte load, writes four destination-adjacent amplitudes per 16/32-byte store, tiles are 2^11
// amplitudes, the shared-memory tile is double buffered and the next tile's loads are in flight
// while the current one is written out.  Shared-memory slot of tile element v (destination
// order): the low nibble of v XOR a fold of its high bits, chosen on the host so that the 16
// lanes of a half-warp hit 16 different 8-byte bank pairs on the write side AND on the read side.
constexpr int kPack2TileBits = 11;
constexpr int kPack2Threads = 256;

struct Pack2Params {
    const float2* src;
    void* dst_hi;
    void* dst_lo;
    const int32_t* rows;
    const uint32_t* amax;
    int32_t rows_mode, rank, tbits, mode, n_outer;
    int32_t inner_bits, blocked, bn_log2, kb_log2;   // PACK_EXPAND_SPLIT_F16 (see PackDesc)
    int32_t planes;                                  // PACK_PLANAR3_F16 (see PackDesc)
    int64_t n_tiles;
    int8_t tile_src_pos[kPack2TileBits];   // source position of tile bit j in SOURCE order
    int8_t u2v[kPack2TileBits];            // destination-order bit that source-order bit j is
    int8_t tile_dst_pos[kPack2TileBits];   // destination position of tile bit j in DESTINATION order
    uint8_t fold[kPack2TileBits];          // nibble XORed into the slot when destination-order bit j is set (j >= 4)
    int8_t outer_dst_pos[TNC_MAX_BITS], outer_src_pos[TNC_MAX_BITS];
};

__device__ __forceinline__ uint32_t pack2_slot(const Pack2Params& p, uint32_t v) {
    uint32_t s = v;
    for (int j = 4; j < p.tbits; ++j)
        if ((v >> j) & 1u) s ^= p.fold[j];
    return s;
}

__global__ void __launch_bounds__(kPack2Threads, 4) pack2_kernel(const Pack2Params p) {
    extern __shared__ __align__(16) unsigned char pack2_smem[];
    const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(pack2_smem);
    const uint32_t tile = 1u << p.tbits;
    const uint32_t buf_bytes = tile * 8u;
    const uint32_t t = threadIdx.x;
    const uint32_t n_pairs = tile >> 1, n_quads = tile >> 2;
    // this thread's elements: pairs (source order) t + 256 i, quads (destination order) t + 256 i
    uint32_t soff[4], doff[2];
    uint32_t ls[4][2], ss[2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t u = 2u * (t + kPack2Threads * i);
        uint32_t so = 0, v = 0;
        for (int j = 0; j < p.tbits; ++j) {
            const uint32_t bit = (u >> j) & 1u;
            so |= bit << p.tile_src_pos[j];
            v |= bit << p.u2v[j];
        }
        soff[i] = so;
        ls[i][0] = pack2_slot(p, v) << 3;
        ls[i][1] = pack2_slot(p, v | (1u << p.u2v[0])) << 3;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const uint32_t v0 = 4u * (t + kPack2Threads * i);
        uint32_t dofs = 0;
        for (int j = 0; j < p.tbits; ++j) dofs |= ((v0 >> j) & 1u) << p.tile_dst_pos[j];
        doff[i] = dofs;
#pragma unroll
        for (int e = 0; e < 4; ++e) ss[i][e] = pack2_slot(p, v0 + e) << 3;
    }
    const int obits = p.rank - p.tbits;
    const int64_t omask = (((int64_t)1 << obits) - 1);
    const int lane = t & 31;
    float sc = 1.f;
    if (p.mode == PACK_SPLIT_F16) sc = f16_scale(*p.amax);
    if (p.mode == PACK_PLANAR3_F16) sc = 0.5f * f16_scale(*p.amax);

    auto bases = [&](int64_t tl, int64_t& sb, int64_t& db) {
        const int64_t blk = tl >> obits;
        const int64_t o = tl & omask;
        // lane j contributes outer bit j; OR-reduce the two 32-bit halves of both offsets
        uint32_t slo = 0, shi = 0, dlo = 0, dhi = 0;
        if (lane < p.n_outer && ((o >> lane) & 1)) {
            const int sp = p.outer_src_pos[lane], dp = p.outer_dst_pos[lane];
            if (sp < 32) slo = 1u << sp; else shi = 1u << (sp - 32);
            if (dp < 32) dlo = 1u << dp; else dhi = 1u << (dp - 32);
        }
        slo = __reduce_or_sync(0xffffffffu, slo);
        shi = __reduce_or_sync(0xffffffffu, shi);
        dlo = __reduce_or_sync(0xffffffffu, dlo);
        dhi = __reduce_or_sync(0xffffffffu, dhi);
        int64_t row = 0;
        if (p.rows_mode == TNC_ROWS_IDENTITY) row = blk;
        else if (p.rows_mode >= 0) row = p.rows[blk];
        sb = (row << p.rank) + (int64_t)(((uint64_t)shi << 32) | slo);
        db = (blk << p.rank) + (int64_t)(((uint64_t)dhi << 32) | dlo);
    };
    const bool expand = p.mode == PACK_EXPAND_SPLIT_F16;
    if (expand) sc = f16_scale(*p.amax);
    const int64_t blk_mask = ((int64_t)1 << p.rank) - 1;

    float4 r[4];
    int64_t sb, db;
    int64_t tl = blockIdx.x;
    if (tl >= p.n_tiles) return;
    bases(tl, sb, db);
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (t + kPack2Threads * i < n_pairs) r[i] = *(const float4*)(p.src + sb + soff[i]);
    uint32_t buf = 0;
    for (;;) {
        const uint32_t sbuf = sm0 + buf * buf_bytes;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (t + kPack2Threads * i < n_pairs) {
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sbuf + ls[i][0]), "f"(r[i].x), "f"(r[i].y) : "memory");
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sbuf + ls[i][1]), "f"(r[i].z), "f"(r[i].w) : "memory");
            }
        __syncthreads();
        const int64_t cur_db = db;
        const int64_t next = tl + gridDim.x;
        const bool more = next < p.n_tiles;
        if (more) {                                   // next tile's loads fly while this one is written out
            bases(next, sb, db);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (t + kPack2Threads * i < n_pairs) r[i] = *(const float4*)(p.src + sb + soff[i]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (t + kPack2Threads * i >= n_quads) continue;
            float2 x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x[e].x), "=f"(x[e].y) : "r"(sbuf + ss[i][e]) : "memory");
            const int64_t d = cur_db + doff[i];
            if (p.mode == PACK_COPY) {
                float4* o = (float4*)((float2*)p.dst_hi + d);
                o[0] = make_float4(x[0].x, x[0].y, x[1].x, x[1].y);
                o[1] = make_float4(x[2].x, x[2].y, x[3].x, x[3].y);
            } else if (p.mode == PACK_ACCUM) {
                float4* o = (float4*)((float2*)p.dst_hi + d);
                float4 c0 = o[0], c1 = o[1];
                c0.x += x[0].x, c0.y += x[0].y, c0.z += x[1].x, c0.w += x[1].y;
                c1.x += x[2].x, c1.y += x[2].y, c1.z += x[3].x, c1.w += x[3].y;
                o[0] = c0;
                o[1] = c1;
            } else if (p.mode == PACK_SPLIT) {
                float2 h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    h[e] = make_float2(tf32_round(x[e].x), tf32_round(x[e].y));
                    l[e] = make_float2(tf32_round(x[e].x - h[e].x), tf32_round(x[e].y - h[e].y));
                }
                float4* oh = (float4*)((float2*)p.dst_hi + d);
                float4* ol = (float4*)((float2*)p.dst_lo + d);
                oh[0] = make_float4(h[0].x, h[0].y, h[1].x, h[1].y);
                oh[1] = make_float4(h[2].x, h[2].y, h[3].x, h[3].y);
                ol[0] = make_float4(l[0].x, l[0].y, l[1].x, l[1].y);
                ol[1] = make_float4(l[2].x, l[2].y, l[3].x, l[3].y);
            } else if (expand) {
                // B'[2n + c'][2k + c], fp16: the four amplitudes are four consecutive k of one n, so
                // each of the two rows (c' = 0, 1) gets one 16-byte store per part
                const int64_t blk = d >> p.rank, q = d & blk_mask;
                const int64_t K = (int64_t)1 << p.inner_bits;
                const int64_t k = q & (K - 1), n = q >> p.inner_bits;
                int64_t e, second;
                if (p.blocked) {
                    const int hb = p.bn_log2 - 1, kbl = p.kb_log2;
                    const int64_t row = (n & (((int64_t)1 << hb) - 1)) << 1;
                    e = (k & (((int64_t)1 << kbl) - 1)) +
                        ((row + ((int64_t)1 << p.bn_log2) * ((k >> kbl) + (K >> kbl) * (n >> hb))) << kbl);
                    second = (int64_t)1 << kbl;
                } else {
                    e = k | (n << (p.inner_bits + 1));
                    second = K;
                }
                uint32_t h0[4], h1[4], l0[4], l1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float xr = x[j].x * sc, xi = x[j].y * sc;
                    const __half hr = __float2half_rn(xr), hi = __float2half_rn(xi);
                    const __half2 a0 = __halves2half2(hr, __hneg(hi)), a1 = __halves2half2(hi, hr);
                    h0[j] = *(const uint32_t*)&a0;
                    h1[j] = *(const uint32_t*)&a1;
                    const __half lr = __float2half_rn(xr - __half2float(hr)), li = __float2half_rn(xi - __half2float(hi));
                    const __half2 b0 = __halves2half2(lr, __hneg(li)), b1 = __halves2half2(li, lr);
                    l0[j] = *(const uint32_t*)&b0;
                    l1[j] = *(const uint32_t*)&b1;
                }
                __half2* hh = (__half2*)p.dst_hi + (blk << (p.rank + 1)) + e;
                *(uint4*)hh = make_uint4(h0[0], h0[1], h0[2], h0[3]);
                *(uint4*)(hh + second) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
                if (p.dst_lo) {
                    __half2* hl = (__half2*)p.dst_lo + (blk << (p.rank + 1)) + e;
                    *(uint4*)hl = make_uint4(l0[0], l0[1], l0[2], l0[3]);
                    *(uint4*)(hl + second) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
                }
            } else if (p.mode == PACK_PLANAR3_F16) {
                // 3M panels: planes re, im, re + im (hi then lo each) of 2^13 halves per 2^13 amplitudes; the
                // four amplitudes are four consecutive k of one row: one 8-byte store per plane
                const int planes = p.dst_lo ? 6 : 3;
                __half* base = (__half*)p.dst_hi + ((d >> 13) * planes << 13) + (d & 8191);
                float v[3][4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float xr = x[e].x * sc, xi = x[e].y * sc;
                    if (p.planes == 1) v[0][e] = xr + xi, v[1][e] = xr, v[2][e] = xi;
                    else if (p.planes == 2) v[0][e] = xr, v[1][e] = xi - xr, v[2][e] = xr + xi;
                    else v[0][e] = xr, v[1][e] = xi, v[2][e] = xr + xi;
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    __half h[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) h[e] = __float2half_rn(v[c][e]);
                    const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
                    __half* o = base + ((p.dst_lo ? 2 * c : c) << 13);
                    *(uint2*)o = make_uint2(*(const uint32_t*)&h01, *(const uint32_t*)&h23);
                    if (p.dst_lo) {
                        const __half2 l01 = __floats2half2_rn(v[c][0] - __half2float(h[0]), v[c][1] - __half2float(h[1]));
                        const __half2 l23 = __floats2half2_rn(v[c][2] - __half2float(h[2]), v[c][3] - __half2float(h[3]));
                        *(uint2*)(o + 8192) = make_uint2(*(const uint32_t*)&l01, *(const uint32_t*)&l23);
                    }
                }
            } else {                                  // PACK_SPLIT_F16
                uint32_t h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float xr = x[e].x * sc, xi = x[e].y * sc;
                    const __half2 hh = __floats2half2_rn(xr, xi);
                    h[e] = *(const uint32_t*)&hh;
                    const float2 hf = __half22float2(hh);
                    const __half2 ll = __floats2half2_rn(xr - hf.x, xi - hf.y);
                    l[e] = *(const uint32_t*)&ll;
                }
                *(uint4*)((__half2*)p.dst_hi + d) = make_uint4(h[0], h[1], h[2], h[3]);
                if (p.dst_lo) *(uint4*)((__half2*)p.dst_lo + d) = make_uint4(l[0], l[1], l[2], l[3]);
            }
        }
        if (!more) break;
        tl = next;
        buf ^= 1u;
    }
}

// GF(2): are the n 4-bit vectors linearly independent?
bool nibbles_independent(const uint8_t* v, int n) {
    uint8_t basis[4] = {0, 0, 0, 0};
    int rank = 0;
    for (int i = 0; i < n; ++i) {
        uint8_t x = v[i] & 15;
        for (int b = 3; b >= 0 && x; --b) {
            if (!((x >> b) & 1)) continue;
            if (basis[b]) {
                x ^= basis[b];
            } else {
                basis[b] = x;
                ++rank;
                x = 0;
            }
        }
    }
    return rank == n;
}

// tile geometry and slot folds of a bit permutation (everything in Pack2Params that does not
// depend on pointers, rows or mode); cached per permutation: a plan launches the same few
// permutations for every slice
void pack2_geometry(const PackDesc& d, Pack2Params& p);

int launch_pack2(const PackDesc& d, const void* src, void* dst_hi, void* dst_lo, cudaStream_t s) {
    static std::mutex mu;
    static std::map<std::string, Pack2Params> cache;
    Pack2Params p{};
    {
        std::string key((const char*)d.src_pos, (size_t)d.rank);
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it == cache.end()) {
            Pack2Params g{};
            pack2_geometry(d, g);
            it = cache.emplace(key, g).first;
        }
        p = it->second;
    }
    p.src = (const float2*)src;
    p.dst_hi = dst_hi;
    p.dst_lo = dst_lo;
    p.rows = d.rows;
    p.amax = d.amax;
    p.rows_mode = d.rows_mode;
    p.rank = d.rank;
    p.mode = d.mode;
    p.inner_bits = d.inner_bits;
    p.planes = d.planes;
    p.blocked = d.blocked;
    p.bn_log2 = d.bn_log2;
    p.kb_log2 = d.kb_log2 ? d.kb_log2 : 4;
    p.n_tiles = (int64_t)d.nb << (d.rank - p.tbits);
    const size_t smem = 2 * ((size_t)8 << p.tbits);
    const int64_t grid = std::min<int64_t>(p.n_tiles, (int64_t)sm_count() * 4);
    pack2_kernel<<<(unsigned)grid, kPack2Threads, smem, s>>>(p);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

void pack2_geometry(const PackDesc& d, Pack2Params& p) {
    const int r = d.rank;
    const int lo = 5;                              // contiguous run wanted on both sides: 2^5 amplitudes
    std::vector<int> in_tile(r, 0);
    for (int i = 0; i < lo; ++i) in_tile[i] = 1;
    for (int i = 0; i < r; ++i)
        if (d.src_pos[i] < lo) in_tile[i] = 1;
    int have = 0;
    for (int i = 0; i < r; ++i) have += in_tile[i];
    for (int i = 0; i < r && have < kPack2TileBits; ++i)
        if (!in_tile[i]) {
            in_tile[i] = 1;
            ++have;
        }
    std::vector<int> tdst;                          // destination order
    for (int i = 0; i < r; ++i)
        if (in_tile[i]) tdst.push_back(i);
    const int t = (int)tdst.size();
    std::vector<int> order(t);                      // source order: tile bits sorted by source position
    for (int j = 0; j < t; ++j) order[j] = j;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return d.src_pos[tdst[x]] < d.src_pos[tdst[y]]; });
    p.tbits = t;
    for (int j = 0; j < t; ++j) {
        p.tile_dst_pos[j] = (int8_t)tdst[j];
        p.tile_src_pos[j] = d.src_pos[tdst[order[j]]];
        p.u2v[j] = (int8_t)order[j];
    }
    // bank-conflict-free slots: M(e_j) = e_j for j < 4, fold[j] for j >= 4; the images of the four
    // bits a half-warp walks on the read side (destination-order bits 2..5) and on the write side
    // (the destination-order bits of source-order bits 1..4) must each be independent
    std::vector<int> S, L;
    for (int j = 2; j <= 5 && j < t; ++j) S.push_back(j);
    for (int j = 1; j <= 4 && j < t; ++j) L.push_back(order[j]);
    auto image = [&](int j) -> uint8_t { return j < 4 ? (uint8_t)(1u << j) : p.fold[j]; };
    auto ok = [&]() {
        uint8_t a[4], b[4];
        for (size_t i = 0; i < S.size(); ++i) a[i] = image(S[i]);
        for (size_t i = 0; i < L.size(); ++i) b[i] = image(L[i]);
        return nibbles_independent(a, (int)S.size()) && nibbles_independent(b, (int)L.size());
    };
    for (int j = 4; j < t; ++j) p.fold[j] = (uint8_t)(1u << (j & 3));
    if (!ok()) {
        uint32_t rng = 12345u;
        bool found = false;
        for (int trial = 0; trial < 200000 && !found; ++trial) {
            for (int j = 4; j < t; ++j) {
                rng = rng * 1664525u + 1013904223u;
                p.fold[j] = (uint8_t)((rng >> 24) & 15u);
            }
            found = ok();
        }
        if (!found)
            for (int j = 4; j < t; ++j) p.fold[j] = (uint8_t)(1u << (j & 3));     // correct, merely conflicted
    }
    p.n_outer = 0;
    for (int i = 0; i < r; ++i)
        if (!in_tile[i]) {
            p.outer_dst_pos[p.n_outer] = (int8_t)i;
            p.outer_src_pos[p.n_outer] = d.src_pos[i];
            ++p.n_outer;
        }
}

}  // namespace

int launch_pack(const PackDesc& d, const void* src, void* dst_hi, void* dst_lo, cudaStream_t s) {
    if (d.rank < 0 || d.rank >= TNC_MAX_BITS || d.nb < 1) {
        set_error("pack: bad descriptor (rank %d, blocks %d)", d.rank, d.nb);
        return TNC_ERR_INVALID;
    }
    if (d.mode == PACK_PLANAR3_F16 && d.rank < 13) {
        set_error("pack: the 3M panels need rank >= 13 (got %d)", d.rank);
        return TNC_ERR_UNSUPPORTED;
    }
    if ((d.mode == PACK_SPLIT_F16 || d.mode == PACK_EXPAND_SPLIT_F16 || d.mode == PACK_PLANAR3_F16) && !d.amax) {
        set_error("pack: the fp16 modes need the operand's amax word");
        return TNC_ERR_INVALID;
    }
    static const bool fast = !(knob("TNC_PACK_FAST") && atoi(knob("TNC_PACK_FAST")) == 0);
    if (fast && d.rank >= 8 && d.rank - 8 < 32 &&
        (d.mode == PACK_COPY || d.mode == PACK_SPLIT || d.mode == PACK_SPLIT_F16 || d.mode == PACK_ACCUM ||
         d.mode == PACK_PLANAR3_F16 ||
         (d.mode == PACK_EXPAND_SPLIT_F16 && d.inner_bits >= 2)))
        return launch_pack2(d, src, dst_hi, dst_lo, s);
    if (d.mode == PACK_ACCUM || d.mode == PACK_PLANAR3_F16) {
        set_error("pack: this mode needs the fast kernel (rank >= 8, got %d)", d.rank);
        return TNC_ERR_UNSUPPORTED;
    }
    PackParams p{};
    p.src = (const float2*)src;
    p.dst_hi = (float2*)dst_hi;
    p.dst_lo = (float2*)dst_lo;
    p.rows = d.rows;
    p.amax = d.amax;
    p.kb_log2 = d.kb_log2 ? d.kb_log2 : 4;
    p.rows_mode = d.rows_mode;
    p.rank = d.rank;
    p.mode = d.mode;
    p.inner_bits = d.inner_bits;
    p.nb = d.nb;
    p.blocked = d.blocked;
    p.bn_log2 = d.bn_log2;
    const int r = d.rank;
    const int lo = std::min(5, r);                 // contiguous run wanted on both sides: 2^5 * 8 B
    // destination positions in the tile: the `lo` lowest destination bits and the destination
    // bits fed by the `lo` lowest source bits
    std::vector<int> in_tile(r, 0);
    for (int i = 0; i < lo; ++i) in_tile[i] = 1;
    for (int i = 0; i < r; ++i)
        if (d.src_pos[i] < lo) in_tile[i] = 1;
    // a tile should carry enough elements to amortise its setup: grow it with the next lowest
    // destination bits (keeps both sides' runs contiguous) up to 2^10 elements
    {
        int have = 0;
        for (int i = 0; i < r; ++i) have += in_tile[i];
        for (int i = 0; i < r && have < kPackMaxTileBits; ++i)
            if (!in_tile[i]) {
                in_tile[i] = 1;
                ++have;
            }
    }
    std::vector<int> tdst;                          // destination order (ascending destination position)
    for (int i = 0; i < r; ++i)
        if (in_tile[i]) tdst.push_back(i);
    const int t = (int)tdst.size();
    if (t > kPackMaxTileBits) {
        set_error("pack: tile of %d bits exceeds %d", t, kPackMaxTileBits);
        return TNC_ERR_UNSUPPORTED;
    }
    std::vector<int> order(t);                      // source order: tile bits sorted by source position
    for (int j = 0; j < t; ++j) order[j] = j;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return d.src_pos[tdst[x]] < d.src_pos[tdst[y]]; });
    p.tbits = t;
    for (int j = 0; j < t; ++j) {
        p.tile_dst_pos[j] = (int8_t)tdst[j];
        p.tile_src_pos[j] = d.src_pos[tdst[order[j]]];
        p.tile_u2v[j] = (int8_t)order[j];
    }
    // XOR swizzle: every high tile bit driven by a low source bit is folded onto a low tile bit
    // that no low source bit drives, so a warp walking the source order and a warp walking the
    // destination order both touch 32 distinct 8-byte slots
    std::vector<int> from, to;
    std::vector<int> driven(t, 0);
    for (int j = 0; j < std::min(lo, t); ++j) driven[order[j]] = 1;
    for (int j = 0; j < std::min(lo, t); ++j)
        if (order[j] >= lo) from.push_back(order[j]);
    for (int j = 0; j < std::min(lo, t); ++j)
        if (!driven[j]) to.push_back(j);
    p.n_xor = (int)std::min(from.size(), to.size());
    for (int j = 0; j < p.n_xor; ++j) {
        p.xor_from[j] = (int8_t)from[j];
        p.xor_to[j] = (int8_t)to[j];
    }
    p.n_outer = 0;
    for (int i = 0; i < r; ++i)
        if (!in_tile[i]) {
            p.outer_dst_pos[p.n_outer] = (int8_t)i;
            p.outer_src_pos[p.n_outer] = d.src_pos[i];
            ++p.n_outer;
        }
    p.n_tiles = (int64_t)d.nb << (r - t);
    const size_t smem = ((size_t)1 << t) * (sizeof(float2) + 2 * sizeof(uint32_t) + 2 * sizeof(uint16_t));
    int64_t grid = std::min<int64_t>(p.n_tiles, (int64_t)sm_count() * 8);
    pack_kernel<<<(unsigned)grid, kPackThreads, smem, s>>>(p);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

// =====================================================================================
// tcgen05 GEMM
// =====================================================================================
namespace {

constexpr int BM = 128;            // rows of A' per CTA tile (UMMA M)
constexpr int BKB = 128;           // bytes per operand row per stage: one 128B-swizzle atom
constexpr int kMmaPerKb = BKB / 32;   // every tcgen05.mma consumes 32 bytes of K per row (8 tf32 / 16 fp16)
constexpr int kEpiWarps = 8;       // two warps per TMEM lane quarter, each owning half of the columns
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2-9 epilogue
constexpr int kSmemBudget = 192 * 1024;
constexpr int kStageBytes = kEpiWarps * 4096;   // per-warp staging buffers of the coalesced epilogue stores

// k-blocks accumulated inside tensor memory before the fp32 register add.  Measured on B200
// (tools/tc_calibrate.py, 3xTF32): coherent shrink of the result per GEMM of -8.6e-8 at 1,
// -2.8e-7 at 2, -6.6e-7 at 4; the fat GEMM runs within 7% of the same speed at 1.  The
// reduced-precision mode tolerates the bias and drains less often.
constexpr int kDefaultKC[3] = {1, 1, 4};

// elements of K per k-block
constexpr int bk_of(int prec) { return prec == TNC_TC_3XTF32 ? BKB / 4 : BKB / 2; }

template <int BN, int PREC>
struct Cfg {
    static constexpr int PANELS = Prec<PREC>::PANELS;
    static constexpr int BK = BKB / Prec<PREC>::ELEM;
    static constexpr int A_TILE = BM * BKB;
    static constexpr int B_TILE = BN * BKB;
    static constexpr int A_LO = A_TILE, B_HI = PANELS * A_TILE, B_LO = PANELS * A_TILE + B_TILE;   // offsets in a stage
    static constexpr int STAGE = PANELS * (A_TILE + B_TILE);
    static constexpr int STAGES = (kSmemBudget / STAGE) > 8 ? 8 : (kSmemBudget / STAGE);
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // two accumulators (chunk double buffer)
    static constexpr int CPT = BN / 2;                            // columns per epilogue thread
    static constexpr int SMEM = STAGES * STAGE + 1024 /*alignment*/ + 256 /*barriers*/ + kStageBytes;
    static_assert(STAGES >= 2, "need a double buffer");
};

struct GemmArgs {
    float* c;
    int64_t c_batch_stride;   // floats
    int32_t ldc;              // floats
    int32_t M, N, K;          // real sizes: rows of A', rows of B' (= columns of C'), columns of A'/B'
    int32_t m_tiles, n_tiles, group_m;
    int32_t a_batched, b_batched;
    int32_t kc;               // k-blocks per TMEM accumulation chunk
    int32_t tiles, rounds;    // persistent grid: CTA c runs tiles c, c + grid, ... (rounds of them)
    int32_t blocked;          // A panel is tile-contiguous: [tile][k-block][rows][128 bytes]
    int32_t blocked_b;        // B' panel likewise
    int32_t n_inner;          // outer-rows step with B's rows folded into N: real columns per row of B (0 = not folded)
    int32_t outer_rj, outer_mb;   // outer-rows step: batch = row of B, GEMM row = (row of A, m): see c_row()
    const int32_t* pair_rows;     // OUTER_PAIRS: C row block of the pair (ra * outer_rj + rb); nullptr: the pair index itself
    int32_t sync_every;       // k-blocks between grid-wide lockstep barriers (0 = none)
    uint32_t* sync_counter;   // zeroed before the launch
    const uint32_t *amax_a, *amax_b;   // fp16 precisions: amax words of A and B (the epilogue undoes their scaling)
    uint32_t* amax_out;       // not null: the largest |component| of C goes here (atomicMax of float bits), see tc_gemm_run
    float out_scale;          // extra factor of the fp16 epilogue: 4 for the 3M panels (scaled one bit lower), else 1
};

// Grid-wide barrier among the co-resident persistent CTAs (one thread per CTA calls it).  It only
// keeps the CTAs within a few k-blocks of each other so that the operand panels they share are
// still in L2 when the next CTA asks for them; the spin is bounded, a missing CTA costs time, not
// correctness.
__device__ __forceinline__ bool lockstep_barrier(uint32_t* counter, uint32_t target, bool wait) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    if (!wait) return false;
    for (int spin = 0; spin < (1 << 15); ++spin) {
        uint32_t v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if (v >= target) return true;
        __nanosleep(64);
    }
    return false;      // a peer is missing (not co-resident): keep arriving but stop waiting
}

// Output row of C for GEMM row `row` of batch `batch`.  Outer-rows steps fold A's rows into the
// GEMM rows and loop B's rows as the batch, while C wants the pair (ra, rb) outermost.
__device__ __forceinline__ int64_t c_row(const GemmArgs& g, int batch, int row) {
    if (g.outer_rj == 0) return (int64_t)batch * g.M + row;
    const int64_t ra = row >> g.outer_mb;
    int64_t pair = ra * g.outer_rj + batch;
    if (g.pair_rows) pair = g.pair_rows[pair];
    return (pair << g.outer_mb) + (row & ((1 << g.outer_mb) - 1));
}

struct TileCoord {
    int batch, m0, n0, m_tile, n_tile;
};
template <int BN>
__device__ __forceinline__ TileCoord tile_coord(const GemmArgs& g, int tile) {
    // groups of `group_m` row tiles sweep all column tiles, so that the CTAs resident at one time
    // share a few A' panels and a few B' panels in L2
    const int tiles_per_batch = g.m_tiles * g.n_tiles;
    TileCoord t;
    t.batch = tile / tiles_per_batch;
    const int r = tile - t.batch * tiles_per_batch;
    const int grp = r / (g.group_m * g.n_tiles);
    const int first_m = grp * g.group_m;
    const int gm = min(g.group_m, g.m_tiles - first_m);
    const int within = r - grp * g.group_m * g.n_tiles;
    t.m_tile = first_m + within % gm;
    t.n_tile = within / gm;
    t.m0 = t.m_tile * BM;
    t.n0 = t.n_tile * BN;
    return t;
}

// One k-block of the split product into the accumulator at `tacc`: small cross terms first, then
// the leading term (hi part only in the reduced-precision mode); +2 = 32 bytes = one MMA K step.
template <int PREC, int CG>
__device__ __forceinline__ void umma_kblock(uint32_t tacc, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                            uint32_t idesc, uint32_t acc) {
    if constexpr (Prec<PREC>::PANELS == 2) {
#pragma unroll
        for (int j = 0; j < kMmaPerKb; ++j) {
            umma<PREC, CG>(tacc, a_lo + 2 * j, b_hi + 2 * j, idesc, acc);
            acc = 1;
        }
#pragma unroll
        for (int j = 0; j < kMmaPerKb; ++j) umma<PREC, CG>(tacc, a_hi + 2 * j, b_lo + 2 * j, idesc, 1);
    }
#pragma unroll
    for (int j = 0; j < kMmaPerKb; ++j) {
        umma<PREC, CG>(tacc, a_hi + 2 * j, b_hi + 2 * j, idesc, acc);
        acc = 1;
    }
}
// Adds one finished TMEM chunk (CPT columns of this thread's lane) into the fp32 registers,
// round-to-nearest.
template <int CPT>
__device__ __forceinline__ void drain_chunk(uint32_t taddr, float* acc) {
    if constexpr (CPT >= 32) {
#pragma unroll
        for (int c = 0; c < CPT; c += 32) {
            uint32_t v[32];
            tmem_ld16(taddr + c, v);
            tmem_ld16(taddr + c + 16, v + 16);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 r = __fadd2_rn(make_float2(acc[c + j], acc[c + j + 1]),
                                            make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
                acc[c + j] = r.x;
                acc[c + j + 1] = r.y;
            }
        }
    } else if constexpr (CPT == 16) {
        uint32_t v[16];
        tmem_ld16(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(v[j]);
    } else {
        uint32_t v[8];
        tmem_ld8(taddr, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += __uint_as_float(v[j]);
    }
}

// Stores a warp's 32 finished rows x CPT columns (lane = row); fp16 precisions undo the operand
// scaling.  Whole tiles go through the warp's shared-memory staging buffer so that every global
// store instruction covers full 128-byte lines (a thread-per-row store touches 32 lines per
// instruction and caps a large-C step near 1 TB/s); ragged tiles keep the guarded per-row path.
template <int CPT, bool F16>
__device__ __forceinline__ void store_tile_rows(const GemmArgs& g, uint32_t stage, int batch, int row0, int col0, float* acc,
                                                int lane) {
    float am = 0.f;            // g.amax_out: largest |component| of this warp's part of the tile, committed per tile (a
                               // running maximum would stay live across the chunk loop of the next tile)
    float sab = 1.f;
    if constexpr (F16) sab = f16_inv_scale(*g.amax_a) * f16_inv_scale(*g.amax_b) * g.out_scale;
    // Output address of GEMM element (row, col).  Folded outer-rows steps (n_inner != 0): the GEMM
    // columns are (row of B, n), and C wants the row pair outermost, so every n_inner columns
    // belong to another row block of C; n_inner >= 32, so a 32-column slab never straddles two.
    auto cptr = [&](int row, int col) -> float* {
        if (g.n_inner == 0) return g.c + c_row(g, batch, row) * g.ldc + col;
        const int rb = col / g.n_inner;
        return g.c + c_row(g, rb, row) * g.ldc + (col - rb * g.n_inner);
    };
    const bool whole = row0 + 32 <= g.M && col0 + CPT <= g.N && (g.outer_rj == 0 || g.outer_mb >= 5);
    if (whole) {
        if constexpr (CPT >= 32) {
#pragma unroll
            for (int c = 0; c < CPT; c += 32) {
                if constexpr (F16) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c + j] *= sab;
                }
                if (g.amax_out) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) amax_fold(am, acc[c + j]);
                }
                store_rows_coalesced<8>(stage, acc + c, cptr(row0, col0 + c), g.ldc, lane);
            }
        } else {
            if constexpr (F16) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) acc[j] *= sab;
            }
            if (g.amax_out) {
#pragma unroll
                for (int j = 0; j < CPT; ++j) amax_fold(am, acc[j]);
            }
            store_rows_coalesced<CPT / 4>(stage, acc, cptr(row0, col0), g.ldc, lane);
        }
        if (g.amax_out) amax_commit(g.amax_out, am);
        return;
    }
    const int row = row0 + lane;
    if (row < g.M) {
#pragma unroll
        for (int j = 0; j < CPT; j += 4)
            if (col0 + j < g.N) {
                float4 o = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                if constexpr (F16) o = make_float4(o.x * sab, o.y * sab, o.z * sab, o.w * sab);
                if (g.amax_out) amax_fold(am, o.x), amax_fold(am, o.y), amax_fold(am, o.z), amax_fold(am, o.w);
                *(float4*)cptr(row, col0 + j) = o;
            }
    }
    __syncwarp();
    if (g.amax_out) amax_commit(g.amax_out, am);
}

template <int BN, int PREC>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
            const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const GemmArgs g) {
    using C = Cfg<BN, PREC>;
    constexpr int BK = C::BK;
    extern __shared__ unsigned char gemm_smem_raw[];
    const uint32_t raw = smem_u32(gemm_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                 // 128B swizzle wants 1024-byte aligned tiles
    unsigned char* aligned = gemm_smem_raw + (base - raw);
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base slot
    const uint32_t bars = base + C::STAGES * C::STAGE;
    uint32_t* tmem_slot = (uint32_t*)(aligned + C::STAGES * C::STAGE + (2 * C::STAGES + 4) * 8);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tmem_full_bar = [&](int b) { return bars + 8u * (2 * C::STAGES + b); };
    auto tmem_empty_bar = [&](int b) { return bars + 8u * (2 * C::STAGES + 2 + b); };

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;

    const int nkb = (g.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full_bar(b), 1);
            mbar_init(tmem_empty_bar(b), kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        {
            uint32_t it = 0, epoch = 0;          // k-blocks loaded so far (stage ring position), barrier epochs
            bool wait_peers = true;
            for (int round = 0; round < g.rounds; ++round) {
                const int tile = round * (int)gridDim.x + (int)blockIdx.x;
                const bool valid = tile < g.tiles;
                TileCoord t{};
                if (valid) t = tile_coord<BN>(g, tile);
                const int ba = g.a_batched ? t.batch : 0, bb = g.b_batched ? t.batch : 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (g.sync_every && kb % g.sync_every == 0) {
                        ++epoch;
                        if (lane == 0) wait_peers = lockstep_barrier(g.sync_counter, epoch * gridDim.x, wait_peers);
                        __syncwarp();
                    }
                    if (!valid) continue;
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    ++it;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (elect_one()) {
                    mbar_expect_tx(full_bar(s), C::STAGE);
                    const uint32_t st = base + s * C::STAGE;
                    // blocked panels: one contiguous [rows][128 bytes] block per (tile, k-block)
                    const int a0 = g.blocked ? 0 : kb * BK, a1 = g.blocked ? 0 : t.m0;
                    const int a2 = g.blocked ? kb + nkb * (t.m_tile + g.m_tiles * ba) : ba;
                    const int b0 = g.blocked_b ? 0 : kb * BK, b1 = g.blocked_b ? 0 : t.n0;
                    const int b2 = g.blocked_b ? kb + nkb * (t.n_tile + g.n_tiles * bb) : bb;
                    tma_load_3d(st, &map_a_hi, full_bar(s), a0, a1, a2);
                    tma_load_3d(st + C::B_HI, &map_b_hi, full_bar(s), b0, b1, b2);
                    if constexpr (C::PANELS == 2) {
                        tma_load_3d(st + C::A_LO, &map_a_lo, full_bar(s), a0, a1, a2);
                        tma_load_3d(st + C::B_LO, &map_b_lo, full_bar(s), b0, b1, b2);
                    }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        {
            constexpr uint32_t idesc = umma_idesc<PREC>(BM, BN);
            const int KC = g.kc;
            uint32_t it = 0, gc = 0;             // k-blocks consumed, accumulation chunks produced
            for (int round = 0; round < g.rounds; ++round) {
                if (round * (int)gridDim.x + (int)blockIdx.x >= g.tiles) break;
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t tacc = tmem + (uint32_t)((gc & 1u) * BN);
                    uint32_t acc = 1;
                    if (kb % KC == 0) {            // a new chunk starts: its TMEM half must have been drained
                        mbar_wait(tmem_empty_bar(gc & 1u), ((gc >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        acc = 0;
                    }
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    ++it;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t st = base + s * C::STAGE;
                    const bool last = kb % KC == KC - 1 || kb == nkb - 1;
                    if (elect_one()) {
                        umma_kblock<PREC, 1>(tacc, umma_desc(st), umma_desc(st + C::A_LO), umma_desc(st + C::B_HI),
                                             umma_desc(st + C::B_LO), idesc, acc);
                        umma_commit(empty_bar(s));     // frees the stage when these MMAs have read it
                        if (last) umma_commit(tmem_full_bar(gc & 1u));
                    }
                    __syncwarp();
                    if (last) ++gc;
                }
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                // TMEM lane quarter this warp is allowed to read
        const int half = (warp - 2) >> 2;      // which half of the tile's columns this warp owns
        constexpr int CPT = C::CPT;
        const int nchunks = (nkb + g.kc - 1) / g.kc;
        uint32_t gc = 0;
        for (int round = 0; round < g.rounds; ++round) {
            const int tile = round * (int)gridDim.x + (int)blockIdx.x;
            if (tile >= g.tiles) break;
            const TileCoord t = tile_coord<BN>(g, tile);
            float acc[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
#pragma unroll 1
            for (int chunk = 0; chunk < nchunks; ++chunk, ++gc) {
                mbar_wait(tmem_full_bar(gc & 1u), (gc >> 1) & 1u);
                tc_fence_after();
                drain_chunk<CPT>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((gc & 1u) * BN + half * CPT), acc);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar(gc & 1u));
            }
            store_tile_rows<CPT, Prec<PREC>::F16>(g, base + C::STAGES * C::STAGE + 256 + (uint32_t)(warp - 2) * 4096u, t.batch,
                                                  t.m0 + q * 32, t.n0 + half * CPT, acc, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}


// =====================================================================================
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x 256 tile.  Each CTA stages
// its own 128 rows of A' and only HALF of the B' tile (128 of its 256 rows); the tensor cores of
// both SMs read both halves.  Per k-block a CTA pulls 64 KB through L2 instead of 96 KB, and
// three stages fit.  Rank 0 of the pair issues the MMAs; TMA completions of both CTAs signal its
// `full` barrier; tcgen05.commit multicasts `empty` / `tmem_full` to both CTAs; the epilogue
// warps of both CTAs release a TMEM half by arriving on rank 0's `tmem_empty` barrier.
// =====================================================================================
template <int PREC>
struct Cfg2 {
    static constexpr int BN = 256;
    static constexpr int PANELS = Prec<PREC>::PANELS;
    static constexpr int BK = BKB / Prec<PREC>::ELEM;
    static constexpr int A_TILE = BM * BKB;             // this CTA's 128 rows of A'
    static constexpr int B_TILE = (BN / 2) * BKB;       // this CTA's half of the B' tile
    static constexpr int A_LO = A_TILE, B_HI = PANELS * A_TILE, B_LO = PANELS * A_TILE + B_TILE;
    static constexpr int STAGE = PANELS * (A_TILE + B_TILE);
    static constexpr int STAGES = (kSmemBudget / STAGE) > 6 ? 6 : (kSmemBudget / STAGE);
    static constexpr int TMEM_COLS = 512;
    static constexpr int CPT = BN / 2;
    static constexpr int SMEM = STAGES * STAGE + 1024 + 256 + kStageBytes;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the barrier at the same offset in CTA rank 0
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    const uint32_t leader_bar = bar & 0xFEFFFFFFu;
    const uint64_t evict_normal = 0x1000000000000000ull;
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "l"(evict_normal)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    const uint16_t both = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(both)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_cta(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
        "r"(cta)
        : "memory");
}

template <int PREC>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, const GemmArgs g) {
    using C = Cfg2<PREC>;
    constexpr int BN = C::BN;
    constexpr int BK = C::BK;
    extern __shared__ unsigned char gemm_smem_raw[];
    const uint32_t raw = smem_u32(gemm_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* aligned = gemm_smem_raw + (base - raw);
    const uint32_t bars = base + C::STAGES * C::STAGE;
    uint32_t* tmem_slot = (uint32_t*)(aligned + C::STAGES * C::STAGE + (2 * C::STAGES + 4) * 8);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tmem_full_bar = [&](int b) { return bars + 8u * (2 * C::STAGES + b); };
    auto tmem_empty_bar = [&](int b) { return bars + 8u * (2 * C::STAGES + 2 + b); };

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)blockIdx.x >> 1, pairs = (int)gridDim.x >> 1;
    const int nkb = (g.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tmem_full_bar(b), 1);
            mbar_init(tmem_empty_bar(b), 2 * kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's barriers and TMEM are ready before anything targets them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // pair tiles: g.m_tiles counts 128-row tiles; a pair covers two of them
    GemmArgs gp = g;
    gp.m_tiles = g.m_tiles / 2;
    const int pair_tiles = g.tiles / 2;

    if (warp == 0) {
        {
            uint32_t it = 0, epoch = 0;
            bool wait_peers = true;
            for (int round = 0; round < g.rounds; ++round) {
                const int tile = round * pairs + pair;
                const bool valid = tile < pair_tiles;
                TileCoord t{};
                if (valid) t = tile_coord<BN>(gp, tile);      // m_tile counts 256-row pair tiles here
                const int ba = g.a_batched ? t.batch : 0, bb = g.b_batched ? t.batch : 0;
                const int m_tile128 = 2 * t.m_tile + (int)rank;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (g.sync_every && kb % g.sync_every == 0) {
                        ++epoch;
                        if (lane == 0) wait_peers = lockstep_barrier(g.sync_counter, epoch * gridDim.x, wait_peers);
                        __syncwarp();
                    }
                    if (!valid) continue;
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    ++it;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(full_bar(s), 2 * C::STAGE);     // both CTAs' bytes land on rank 0's barrier
                    const uint32_t st = base + s * C::STAGE;
                    const int a0 = g.blocked ? 0 : kb * BK, a1 = g.blocked ? 0 : m_tile128 * BM;
                    const int a2 = g.blocked ? kb + nkb * (m_tile128 + g.m_tiles * ba) : ba;
                    const int b0 = g.blocked_b ? 0 : kb * BK, b1 = (g.blocked_b ? 0 : t.n0) + 128 * (int)rank;
                    const int b2 = g.blocked_b ? kb + nkb * (t.n_tile + g.n_tiles * bb) : bb;
                    tma_load_3d_2sm(st, &map_a_hi, full_bar(s), a0, a1, a2);
                    tma_load_3d_2sm(st + C::B_HI, &map_b_hi, full_bar(s), b0, b1, b2);
                    if constexpr (C::PANELS == 2) {
                        tma_load_3d_2sm(st + C::A_LO, &map_a_lo, full_bar(s), a0, a1, a2);
                        tma_load_3d_2sm(st + C::B_LO, &map_b_lo, full_bar(s), b0, b1, b2);
                    }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // N = 256, M = 256 (128 rows in each CTA of the pair)
            constexpr uint32_t idesc = umma_idesc<PREC>(256, BN);
            const int KC = g.kc;
            uint32_t it = 0, gc = 0;
            for (int round = 0; round < g.rounds; ++round) {
                if (round * pairs + pair >= pair_tiles) break;
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t tacc = tmem + (uint32_t)((gc & 1u) * BN);
                    uint32_t acc = 1;
                    if (kb % KC == 0) {
                        mbar_wait(tmem_empty_bar(gc & 1u), ((gc >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        acc = 0;
                    }
                    const int s = it % C::STAGES;
                    const uint32_t ph = (it / C::STAGES) & 1;
                    ++it;
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t st = base + s * C::STAGE;
                    const bool last = kb % KC == KC - 1 || kb == nkb - 1;
                    if (elect_one()) {
                        umma_kblock<PREC, 2>(tacc, umma_desc(st), umma_desc(st + C::A_LO), umma_desc(st + C::B_HI),
                                             umma_desc(st + C::B_LO), idesc, acc);
                        umma_commit_2sm(empty_bar(s));
                        if (last) umma_commit_2sm(tmem_full_bar(gc & 1u));
                    }
                    __syncwarp();
                    if (last) ++gc;
                }
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int CPT = C::CPT;
        const int nchunks = (nkb + g.kc - 1) / g.kc;
        uint32_t gc = 0;
        for (int round = 0; round < g.rounds; ++round) {
            const int tile = round * pairs + pair;
            if (tile >= pair_tiles) break;
            const TileCoord t = tile_coord<BN>(gp, tile);
            float acc[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
#pragma unroll 1
            for (int chunk = 0; chunk < nchunks; ++chunk, ++gc) {
                mbar_wait(tmem_full_bar(gc & 1u), (gc >> 1) & 1u);
                tc_fence_after();
                drain_chunk<CPT>(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((gc & 1u) * BN + half * CPT), acc);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_on_cta(tmem_empty_bar(gc & 1u), 0);
            }
            store_tile_rows<CPT, Prec<PREC>::F16>(g, base + C::STAGES * C::STAGE + 256 + (uint32_t)(warp - 2) * 4096u, t.batch,
                                                  (2 * t.m_tile + (int)rank) * BM + q * 32, t.n0 + half * CPT, acc, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // nobody leaves while the pair's MMAs may still read its shared memory
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}


// =====================================================================================
// 3M (Karatsuba) complex product on CTA pairs.  With planar fp16 panels re / im / (re + im) of both
// operands,
//     P1 = Ar Br,   P2 = Ai Bi,   P3 = (Ar + Ai)(Br + Bi)       (three REAL M x N x K products)
//     Cr = P1 - P2,               Ci = P3 - P1 - P2
// is 6 M N K real flops instead of the 8 of the interleaved 4M form above: 25 % fewer tensor-core
// instructions per useful flop, which is the only lever left on a GEMM that already keeps the
// tensor pipe busy under the power cap.  Same output tile as gemm_2cta_kernel (a pair owns 256
// rows x 128 complex columns of C, interleaved fp32 accumulators in the epilogue warps'
// registers), same chunked accumulation: tensor memory only ever holds the sum of `kc` k-blocks
// of ONE product.  TMEM is a ring of four 128-column slots that the products of consecutive
// chunks walk through (P1, P2, P3, P1, ...); the epilogue warps fold each finished slot into the
// registers with the product's signs (no temporaries: Cr += P1, Ci -= P1; Cr -= P2, Ci -= P2;
// Ci += P3, every add rounded to nearest) and free it at once.  The pipeline unit is one product
// of one k-block: a stage holds A_p hi/lo (this CTA's 128 rows x 64 k) and B_p hi/lo (this CTA's
// 64 of the tile's 128 columns) = 48 KB, four stages.
// =====================================================================================
template <int PREC>
struct Cfg3 {
    static constexpr int BN = 256;                      // real (interleaved) columns of the output tile
    static constexpr int BNC = 128;                     // complex columns = N of every MMA
    static constexpr int PANELS = Prec<PREC>::PANELS;   // hi (+ lo)
    static constexpr int BK = BKB / 2;                  // 64 complex k per k-block (planar fp16 rows of 128 bytes)
    static constexpr int A_TILE = BM * BKB;             // one part of this CTA's 128 rows
    static constexpr int B_TILE = (BNC / 2) * BKB;      // one part of this CTA's half of the B tile
    static constexpr int A_LO = A_TILE, B_HI = PANELS * A_TILE, B_LO = PANELS * A_TILE + B_TILE;
    static constexpr int STAGE = PANELS * (A_TILE + B_TILE);
    static constexpr int STAGES = (kSmemBudget / STAGE) > 8 ? 8 : (kSmemBudget / STAGE);
    static constexpr int SLOTS = 4;                     // TMEM ring: 4 x 128 columns
    static constexpr int TMEM_COLS = 512;
    static constexpr int CPT = BN / 2;                  // interleaved floats per epilogue thread (64 complex)
    static constexpr int SMEM = STAGES * STAGE + 1024 + 256 + kStageBytes;
    static_assert(Prec<PREC>::F16, "the 3M kernel is fp16 only");
};

// Folds 64 columns of product P (0: Ar Br, 1: Ai Bi, 2: (Ar + Ai)(Br + Bi)) into the planar accumulators
// (re[j], im[j] = columns 2j, 2j + 1) with packed fp32x2 adds: 2.5 instructions per complex output and chunk.
// `release` runs as soon as the slot's last columns sit in registers (before their adds).
// GAUSS: the products are X = (Ar + Ai) Br, Y = Ar (Bi - Br), Z = Ai (Br + Bi) with Cr = X - Z, Ci = X + Y: four
// packed adds per column pair and chunk instead of Karatsuba's five.
template <int P, int DB, bool GAUSS, typename Release>
__device__ __forceinline__ void drain_3m(uint32_t taddr, float2* re, float2* im, Release release) {
    const float2 neg = make_float2(-1.f, -1.f);
#pragma unroll
    for (int c = 0; c < 64; c += DB) {
        uint32_t v[DB];
        tmem_ld16(taddr + c, v);
        if constexpr (DB == 32) tmem_ld16(taddr + c + 16, v + 16);
        tmem_ld_wait();
        if (c + DB == 64) release();
#pragma unroll
        for (int j = 0; j < DB; j += 2) {
            const float2 x = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
            const int o = (c + j) >> 1;
            if constexpr (GAUSS) {
                if (P == 0) {
                    re[o] = __fadd2_rn(re[o], x);
                    im[o] = __fadd2_rn(im[o], x);
                } else if (P == 1) {
                    im[o] = __fadd2_rn(im[o], x);
                } else {
                    re[o] = __ffma2_rn(x, neg, re[o]);
                }
            } else if (P == 0) {
                re[o] = __fadd2_rn(re[o], x);
                im[o] = __ffma2_rn(x, neg, im[o]);
            } else if (P == 1) {
                re[o] = __ffma2_rn(x, neg, re[o]);
                im[o] = __ffma2_rn(x, neg, im[o]);
            } else {
                im[o] = __fadd2_rn(im[o], x);
            }
        }
    }
}

template <int PREC, int DB, bool GAUSS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm3m_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmArgs g) {
    using C = Cfg3<PREC>;
    constexpr int BN = C::BN;
    constexpr int BK = C::BK;
    constexpr int HL = C::PANELS;
    extern __shared__ unsigned char gemm_smem_raw[];
    const uint32_t raw = smem_u32(gemm_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* aligned = gemm_smem_raw + (base - raw);
    // barriers: full[STAGES], empty[STAGES], tmem_full[SLOTS], tmem_empty[SLOTS]; then the TMEM base slot
    const uint32_t bars = base + C::STAGES * C::STAGE;
    uint32_t* tmem_slot = (uint32_t*)(aligned + C::STAGES * C::STAGE + (2 * C::STAGES + 2 * C::SLOTS) * 8);
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
    auto tmem_full_bar = [&](int b) { return bars + 8u * (2 * C::STAGES + b); };
    auto tmem_empty_bar = [&](int b) { return bars + 8u * (2 * C::STAGES + C::SLOTS + b); };

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)blockIdx.x >> 1, pairs = (int)gridDim.x >> 1;
    const int nkb = (g.K / 2) / BK;                     // g.K counts interleaved reals: K / 2 complex k
    const int KC = g.kc;
    const int nchunks = (nkb + KC - 1) / KC;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < C::SLOTS; ++b) {
            mbar_init(tmem_full_bar(b), 1);
            mbar_init(tmem_empty_bar(b), 2 * kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    GemmArgs gp = g;
    gp.m_tiles = g.m_tiles / 2;                         // tile_coord counts 256-row pair tiles
    const int pair_tiles = g.tiles / 2;

    if (warp == 0) {
        {
            uint32_t it = 0, epoch = 0;
            bool wait_peers = true;
            for (int round = 0; round < g.rounds; ++round) {
                const int tile = round * pairs + pair;
                const bool valid = tile < pair_tiles;
                TileCoord t{};
                if (valid) t = tile_coord<BN>(gp, tile);
                const int ba = g.a_batched ? t.batch : 0, bb = g.b_batched ? t.batch : 0;
                const int m_tile128 = 2 * t.m_tile + (int)rank;
                const int a_blk = nkb * (m_tile128 + g.m_tiles * ba), b_blk = nkb * (t.n_tile + g.n_tiles * bb);
                for (int chunk = 0; chunk < nchunks; ++chunk) {
                    const int kb0 = chunk * KC, kb1 = min(nkb, kb0 + KC);
                    if (g.sync_every && kb0 % g.sync_every == 0) {
                        ++epoch;
                        if (lane == 0) wait_peers = lockstep_barrier(g.sync_counter, epoch * gridDim.x, wait_peers);
                        __syncwarp();
                    }
                    if (!valid) continue;
                    for (int p = 0; p < 3; ++p)
                        for (int kb = kb0; kb < kb1; ++kb) {
                            const int s = it % C::STAGES;
                            const uint32_t ph = (it / C::STAGES) & 1;
                            ++it;
                            mbar_wait(empty_bar(s), ph ^ 1u);
                            if (elect_one()) {
                                if (rank == 0) mbar_expect_tx(full_bar(s), 2 * C::STAGE);
                                const uint32_t st = base + s * C::STAGE;
                                // A: one box of HL x 128 rows (hi then lo of part p); B: this CTA's 64 rows of hi, of lo
                                tma_load_3d_2sm(st, &map_a, full_bar(s), 0, 0, (a_blk + kb) * 3 + p);
                                tma_load_3d_2sm(st + C::B_HI, &map_b, full_bar(s), 0, 64 * (int)rank,
                                                ((b_blk + kb) * 3 + p) * HL);
                                if constexpr (HL == 2)
                                    tma_load_3d_2sm(st + C::B_LO, &map_b, full_bar(s), 0, 64 * (int)rank,
                                                    ((b_blk + kb) * 3 + p) * HL + 1);
                            }
                            __syncwarp();
                        }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc<PREC>(256, C::BNC);
            uint32_t it = 0, gs = 0;                     // stages consumed, TMEM slots produced
            for (int round = 0; round < g.rounds; ++round) {
                if (round * pairs + pair >= pair_tiles) break;
                for (int chunk = 0; chunk < nchunks; ++chunk) {
                    const int kb0 = chunk * KC, kb1 = min(nkb, kb0 + KC);
                    for (int p = 0; p < 3; ++p, ++gs) {
                        const uint32_t slot = gs & 3u;
                        const uint32_t tacc = tmem + slot * (uint32_t)C::BNC;
                        mbar_wait(tmem_empty_bar(slot), ((gs >> 2) & 1u) ^ 1u);
                        tc_fence_after();
                        uint32_t acc = 0;
                        for (int kb = kb0; kb < kb1; ++kb) {
                            const int s = it % C::STAGES;
                            const uint32_t ph = (it / C::STAGES) & 1;
                            ++it;
                            mbar_wait(full_bar(s), ph);
                            tc_fence_after();
                            const uint32_t st = base + s * C::STAGE;
                            if (elect_one()) {
                                umma_kblock<PREC, 2>(tacc, umma_desc(st), umma_desc(st + C::A_LO), umma_desc(st + C::B_HI),
                                                     umma_desc(st + C::B_LO), idesc, acc);
                                umma_commit_2sm(empty_bar(s));
                                if (kb == kb1 - 1) umma_commit_2sm(tmem_full_bar(slot));
                            }
                            __syncwarp();
                            acc = 1;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int CPT = C::CPT;
        uint32_t gs = 0;
        for (int round = 0; round < g.rounds; ++round) {
            const int tile = round * pairs + pair;
            if (tile >= pair_tiles) break;
            const TileCoord t = tile_coord<BN>(gp, tile);
            float2 re[CPT / 4], im[CPT / 4];
#pragma unroll
            for (int j = 0; j < CPT / 4; ++j) re[j] = im[j] = make_float2(0.f, 0.f);
            const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
            auto slot_wait = [&]() {
                mbar_wait(tmem_full_bar(gs & 3u), (gs >> 2) & 1u);
                tc_fence_after();
                return lane_base + (gs & 3u) * (uint32_t)C::BNC;
            };
            auto slot_free = [&]() {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_on_cta(tmem_empty_bar(gs & 3u), 0);
                ++gs;
            };
#pragma unroll 1
            for (int chunk = 0; chunk < nchunks; ++chunk) {
                drain_3m<0, DB, GAUSS>(slot_wait(), re, im, slot_free);
                drain_3m<1, DB, GAUSS>(slot_wait(), re, im, slot_free);
                drain_3m<2, DB, GAUSS>(slot_wait(), re, im, slot_free);
            }
            // whole tiles by construction (M % 256 == 0, N % 256 == 0, no folded right-operand rows): 32
            // interleaved floats (16 complex columns) at a time through the warp's staging buffer, straight
            // from the planar accumulators (an interleaved copy of all 128 made the chunk loop spill)
            {
                const float sab = f16_inv_scale(*g.amax_a) * f16_inv_scale(*g.amax_b) * g.out_scale;
                const uint32_t stage = base + C::STAGES * C::STAGE + 256 + (uint32_t)(warp - 2) * 4096u;
                const int row0 = (2 * t.m_tile + (int)rank) * BM + q * 32, col0 = t.n0 + half * CPT;
                float* crow = g.c + c_row(g, t.batch, row0) * g.ldc + col0;
                float am = 0.f;                // g.amax_out, committed per tile as in store_tile_rows
#pragma unroll
                for (int c = 0; c < CPT; c += 32) {
                    float o[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        o[4 * j] = re[c / 4 + j].x * sab;
                        o[4 * j + 1] = im[c / 4 + j].x * sab;
                        o[4 * j + 2] = re[c / 4 + j].y * sab;
                        o[4 * j + 3] = im[c / 4 + j].y * sab;
                    }
                    if (g.amax_out) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) amax_fold(am, o[j]);
                    }
                    store_rows_coalesced<8>(stage, o, crow + c, g.ldc, lane);
                }
                if (g.amax_out) amax_commit(g.amax_out, am);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 3-D map over elements of `elem` bytes (4: fp32/tf32, 2: fp16): dims {d0, d1, d2}, d0 contiguous;
// box = one 128-byte swizzle atom of d0 x box_rows of d1
int make_map(CUtensorMap* map, void* addr, int elem, int64_t d0, int64_t d1, int64_t d2, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return TNC_ERR_CUDA;
    }
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)d0 * elem, (cuuint64_t)d0 * elem * (cuuint64_t)d1};
    cuuint32_t box[3] = {(cuuint32_t)(BKB / elem), (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = fn(map, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, addr, dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d (elem=%d dims=%lld x %lld x %lld box_rows=%d)", (int)rc, elem,
                  (long long)d0, (long long)d1, (long long)d2, box_rows);
        return TNC_ERR_CUDA;
    }
    return TNC_OK;
}

template <int BN, int PREC>
int launch_gemm(const CUtensorMap* maps, const GemmArgs& g, int64_t grid, cudaStream_t s) {
    static bool configured_on[kMaxDevices] = {};       // the attribute is per device
    bool& configured = configured_on[current_device()];
    if (!configured) {
        TNC_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN, PREC>::SMEM));
        configured = true;
    }
    gemm_kernel<BN, PREC><<<(unsigned)grid, kGemmThreads, Cfg<BN, PREC>::SMEM, s>>>(maps[0], maps[1], maps[2], maps[3], g);
    TNC_CUDA(cudaGetLastError());
    return TNC_OK;
}

template <int PREC>
int launch_gemm_bn(int bn, const CUtensorMap* maps, const GemmArgs& g, int64_t grid, cudaStream_t s) {
    switch (bn) {
        case 256: return launch_gemm<256, PREC>(maps, g, grid, s);
        case 128: return launch_gemm<128, PREC>(maps, g, grid, s);
        case 64: return launch_gemm<64, PREC>(maps, g, grid, s);
        case 32: return launch_gemm<32, PREC>(maps, g, grid, s);
        default: return launch_gemm<16, PREC>(maps, g, grid, s);
    }
}

template <int PREC>
int launch_gemm_2cta(const CUtensorMap* maps, const GemmArgs& g, int64_t grid, cudaStream_t s) {
    static bool configured_on[kMaxDevices] = {};       // the attribute is per device
    bool& configured = configured_on[current_device()];
    if (!configured) {
        TNC_CUDA(cudaFuncSetAttribute(gemm_2cta_kernel<PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<PREC>::SMEM));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg2<PREC>::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TNC_CUDA(cudaLaunchKernelEx(&cfg, gemm_2cta_kernel<PREC>, maps[0], maps[1], maps[2], maps[3], g));
    return TNC_OK;
}

template <int PREC, int DB, bool GAUSS>
int launch_gemm3m(const CUtensorMap* maps, const GemmArgs& g, int64_t grid, cudaStream_t s) {
    static bool configured_on[kMaxDevices] = {};       // the attribute is per device
    bool& configured = configured_on[current_device()];
    if (!configured) {
        TNC_CUDA(cudaFuncSetAttribute(gemm3m_2cta_kernel<PREC, DB, GAUSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg3<PREC>::SMEM));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = Cfg3<PREC>::SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TNC_CUDA(cudaLaunchKernelEx(&cfg, gemm3m_2cta_kernel<PREC, DB, GAUSS>, maps[0], maps[2], g));
    return TNC_OK;
}

int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace

struct TcGemmOp {
    PackDesc pa{}, pb{};
    int precision = TNC_TC_3XTF32;
    int64_t a_off = 0, b_off = 0, c_off = 0;          // operand offsets in the workspace (bytes)
    int64_t a_elems = 0, b_elems = 0;                 // complex elements of the whole source tensors (amax range)
    int64_t ahi_off = 0, alo_off = 0, bhi_off = 0, blo_off = 0;
    int64_t M = 0, N = 0, K = 0;                      // real GEMM sizes
    int64_t batch = 1;
    int a_batched = 0, b_batched = 0;
    int n_inner = 0;
    int bn = 256;
    GemmArgs args{};
    int64_t tiles = 0;
    // tensor maps per workspace base (they embed device addresses); the only state a run touches, under `mu`:
    // a plan may be executed from several host threads, each with its own workspace and stream
    struct Maps {
        CUtensorMap m[4];
    };
    std::mutex mu;
    std::map<char*, Maps> maps_by_ws;
    int blocked = 0, blocked_b = 0;                   // tile-contiguous panels (A, B')
    int two_cta = 0;                                  // CTA pairs (cta_group::2) on 256 x 256 tiles
    int use_3m = 0;                                   // 3M complex product on planar panels (gemm3m_2cta_kernel)
    int gauss = 0;                                    //   in Gauss's form (planes s, r, i / r, i - r, s) instead of Karatsuba's
    int grid = 1;
    int64_t words_off = 0;                            // 256 bytes at the end of the step's scratch region (in the WORKSPACE, so that
                                                      // executions with different workspaces never share them): [0], [1] amax bits
                                                      // of A, B; [32] lockstep barrier counter
};

namespace {

struct Shape {
    int64_t nb_a, nb_b, batch, M, N, K;
    int fold_rows;   // A's identity rows folded into M
    int outer;       // TNC_EINSUM_OUTER_ROWS lowering
    int n_inner;     // outer lowering with B's rows folded into N: real columns per row of B (0: B's rows are the batch)
};

int shape_of(const tnc_einsum& e, Shape* sh) {
    if (e.n_h != 0) {
        set_error("tensor-core einsum: shared kept modes are not supported");
        return TNC_ERR_UNSUPPORTED;
    }
    if (e.n_k < 1 || e.n_n < 1) {
        set_error("tensor-core einsum: needs at least one contracted and one right-only bit (k=%d, n=%d)", e.n_k, e.n_n);
        return TNC_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < e.n_n; ++i)
        if (e.n_c[i] >= e.n_n) {
            set_error("tensor-core einsum: the right-only modes must occupy the %d lowest positions of the output", e.n_n);
            return TNC_ERR_UNSUPPORTED;
        }
    // without rows on the right operand the gathered rows of A simply extend M
    sh->fold_rows = e.rows_b == TNC_ROWS_NONE || e.nb == 1;
    sh->nb_a = e.rows_a == TNC_ROWS_NONE ? 1 : e.nb;
    sh->nb_b = e.rows_b == TNC_ROWS_NONE ? 1 : e.nb;
    if (e.rows_a == TNC_ROWS_NONE && e.rows_b == TNC_ROWS_NONE && e.nb != 1) {
        set_error("tensor-core einsum: %d output rows but neither operand has rows", e.nb);
        return TNC_ERR_INVALID;
    }
    sh->batch = sh->fold_rows ? 1 : e.nb;
    sh->outer = (e.flags & (TNC_EINSUM_OUTER_ROWS | TNC_EINSUM_OUTER_PAIRS)) && e.nb > 1 && !sh->fold_rows;
    sh->n_inner = 0;
    if (sh->outer) {          // A's rows extend M, B's rows are the batch: nothing is gathered twice
        sh->nb_a = e.a.rows;
        sh->nb_b = e.b.rows;
        sh->batch = e.b.rows;
        // ... or, when a row of B spans whole 32-column slabs, B's rows extend N instead: ONE GEMM
        // that reads the (large) left panel once per 256 columns instead of once per row of B
        if (((int64_t)2 << e.n_n) >= 32 && !(knob("TNC_TC_FOLDN") && atoi(knob("TNC_TC_FOLDN")) == 0)) {
            sh->n_inner = 2 << e.n_n;
            sh->batch = 1;
        }
    }
    sh->M = ((int64_t)1 << e.n_m) * ((sh->fold_rows || sh->outer) ? sh->nb_a : 1);
    sh->N = ((int64_t)2 << e.n_n) * (sh->n_inner ? sh->nb_b : 1);
    sh->K = (int64_t)2 << e.n_k;
    if (sh->M >= ((int64_t)1 << 31) || sh->N >= ((int64_t)1 << 31) || sh->K >= ((int64_t)1 << 31) ||
        sh->batch >= ((int64_t)1 << 31)) {
        set_error("tensor-core einsum: a GEMM dimension exceeds 2^31");
        return TNC_ERR_UNSUPPORTED;
    }
    return TNC_OK;
}

}  // namespace

// Sized for the largest panels of any precision (3xTF32: 8 bytes per amplitude of A per part, 16
// per amplitude of B'); the fp16 precisions use half of it or less.
int64_t tc_gemm_scratch_bytes(const tnc_einsum& e, int dtype) {
    if (dtype != TNC_C64) {
        set_error("tensor-core einsum: only complex64 is implemented");
        return -1;
    }
    Shape sh;
    if (shape_of(e, &sh) != TNC_OK) return -1;
    const int64_t a_panel = align_up((sh.nb_a << (e.n_m + e.n_k)) * 8, 1024);
    const int64_t b_panel = align_up((sh.nb_b << (e.n_n + e.n_k)) * 16, 1024);
    return 2 * a_panel + 2 * b_panel + 1024;          // + the amax / barrier words
}

int tc_gemm_create(const tnc_einsum& e, int dtype, int precision, const int32_t* dev_rows_a, const int32_t* dev_rows_b,
                   const int32_t* dev_pair_rows, TcGemmOp** out) {
    const int64_t need = tc_gemm_scratch_bytes(e, dtype);
    if (need < 0) return TNC_ERR_UNSUPPORTED;
    if (e.scratch_bytes < need || (e.scratch_offset & 1023)) {
        set_error("tensor-core einsum: scratch region too small or misaligned (%lld < %lld)", (long long)e.scratch_bytes,
                  (long long)need);
        return TNC_ERR_NOMEM;
    }
    if (precision != TNC_TC_3XTF32 && precision != TNC_TC_3XF16 && precision != TNC_TC_F16) {
        set_error("tensor-core einsum: unknown precision %d", precision);
        return TNC_ERR_INVALID;
    }
    const bool f16 = precision != TNC_TC_3XTF32;
    if (f16 && e.n_k < 2) {
        // a row of the fp16 panel must be a multiple of 16 bytes for TMA: K >= 4 complex
        set_error("tensor-core einsum: the fp16 precisions need at least 2 contracted bits (k=%d)", e.n_k);
        return TNC_ERR_UNSUPPORTED;
    }
    Shape sh;
    shape_of(e, &sh);
    for (int i = 0; i < e.n_m; ++i)
        if (e.m_c[i] < e.n_n) {
            set_error("tensor-core einsum: a left-only mode sits below the right-only modes in the output");
            return TNC_ERR_UNSUPPORTED;
        }
    TcGemmOp* op = new TcGemmOp();
    op->precision = precision;
    const int kbl = f16 ? 5 : 4;                       // log2 of the complex k per k-block (128 bytes per row)
    // A panel: [rows][m (output order)][k]  -- destination bits: k first, then m by output position
    op->pa.rank = e.a.rank;
    op->pa.nb = (int32_t)sh.nb_a;
    op->pa.rows_mode = sh.outer ? TNC_ROWS_IDENTITY : e.rows_a;
    op->pa.rows = dev_rows_a;
    op->pa.mode = f16 ? PACK_SPLIT_F16 : PACK_SPLIT;
    op->pa.inner_bits = e.n_k;
    op->pa.kb_log2 = kbl;
    const int bn = sh.N >= 256 ? 256 : sh.N >= 128 ? 128 : sh.N >= 64 ? 64 : sh.N >= 32 ? 32 : 16;
    // tile-contiguous panels need whole tiles: K a multiple of one k-block, every row block a
    // multiple of 128 rows, N a multiple of the tile width (a folded B' panel keeps the plain
    // [row of B][n][k] order: its row count need not be a multiple of the tile)
    op->blocked = e.n_k >= kbl && e.n_m >= 7 && sh.N >= bn;
    if (const char* env = knob("TNC_TC_BLOCKED"))
        if (atoi(env) == 0) op->blocked = 0;
    op->blocked_b = op->blocked && sh.n_inner == 0;
    // CTA pairs on 256 x 256 tiles
    op->two_cta = bn == 256 && sh.M % 256 == 0 && sh.M / BM >= 2;
    if (const char* env = knob("TNC_TC_2CTA"))
        if (atoi(env) == 0) op->two_cta = 0;
    // 3M complex product (gemm3m_2cta_kernel): fp16 precisions, whole 64-k blocks and pair tiles
    // (3xF16 only: in the single-product complex-half mode the 4M kernel is as fast -- both sit at the same
    // power / shared-memory ceiling, 52.3 vs 53.4 ms on the fat step -- and 1.4x more accurate, the imaginary
    // part of 3M being a difference of larger products)
    op->use_3m = precision == TNC_TC_3XF16 && op->two_cta && op->blocked_b && e.n_k >= 6 && sh.N % 256 == 0;
    if (const char* env = knob("TNC_TC_3M"))
        if (atoi(env) == 0) op->use_3m = 0;
    if (op->use_3m) {
        // planar panels, both operands alike: [tile of 128 rows][block of 64 k][re, im, re + im][hi, lo][128 rows][64 k]
        op->pa.mode = op->pb.mode = PACK_PLANAR3_F16;
        // Gauss's form saves one of five packed adds per column pair and chunk; measured on one box, three
        // alternations: 123.4 vs 123.9 ms on the fat step -- no difference, so Karatsuba's stays the default
        op->gauss = 0;
        if (const char* env = knob("TNC_TC_GAUSS")) op->gauss = atoi(env) != 0;
        op->pa.planes = op->gauss ? 1 : 0;
        op->pb.planes = op->gauss ? 2 : 0;
        int8_t pm[TNC_MAX_BITS], pn[TNC_MAX_BITS];      // operand position of row bit j (output order)
        for (int i = 0; i < e.n_m; ++i) pm[e.m_c[i] - e.n_n] = e.m_a[i];
        for (int i = 0; i < e.n_n; ++i) pn[e.n_c[i]] = e.n_b[i];
        int d = 0;
        for (int i = 0; i < 6; ++i) op->pa.src_pos[d++] = e.k_a[i];
        for (int j = 0; j < 7; ++j) op->pa.src_pos[d++] = pm[j];
        for (int i = 6; i < e.n_k; ++i) op->pa.src_pos[d++] = e.k_a[i];
        for (int j = 7; j < e.n_m; ++j) op->pa.src_pos[d++] = pm[j];
        d = 0;
        for (int i = 0; i < 6; ++i) op->pb.src_pos[d++] = e.k_b[i];
        for (int j = 0; j < 7; ++j) op->pb.src_pos[d++] = pn[j];
        for (int i = 6; i < e.n_k; ++i) op->pb.src_pos[d++] = e.k_b[i];
        for (int j = 7; j < e.n_n; ++j) op->pb.src_pos[d++] = pn[j];
    } else {
        int8_t pm[TNC_MAX_BITS];                       // A position of row bit j (output order)
        for (int i = 0; i < e.n_m; ++i) pm[e.m_c[i] - e.n_n] = e.m_a[i];
        int d = 0;
        if (op->blocked) {
            // [m tile][k block][128 rows][k within the block]
            for (int i = 0; i < kbl; ++i) op->pa.src_pos[d++] = e.k_a[i];
            for (int j = 0; j < 7; ++j) op->pa.src_pos[d++] = pm[j];
            for (int i = kbl; i < e.n_k; ++i) op->pa.src_pos[d++] = e.k_a[i];
            for (int j = 7; j < e.n_m; ++j) op->pa.src_pos[d++] = pm[j];
        } else {
            for (int i = 0; i < e.n_k; ++i) op->pa.src_pos[d++] = e.k_a[i];
            for (int j = 0; j < e.n_m; ++j) op->pa.src_pos[d++] = pm[j];
        }
    }
    // B' panel: [rows][n (output order)][c'][k][c]
    op->pb.rank = e.b.rank;
    op->pb.nb = (int32_t)sh.nb_b;
    op->pb.rows_mode = sh.outer ? TNC_ROWS_IDENTITY : e.rows_b;
    op->pb.rows = dev_rows_b;
    op->pb.inner_bits = e.n_k;
    op->pb.kb_log2 = kbl;
    op->pb.blocked = op->blocked_b;
    for (int l = 0; (1 << l) <= bn; ++l) op->pb.bn_log2 = l;
    if (!op->use_3m) {
        op->pb.mode = f16 ? PACK_EXPAND_SPLIT_F16 : PACK_EXPAND_SPLIT;
        for (int i = 0; i < e.n_k; ++i) op->pb.src_pos[i] = e.k_b[i];
        for (int i = 0; i < e.n_n; ++i) op->pb.src_pos[e.n_k + e.n_c[i]] = e.n_b[i];
    }
    op->a_off = e.a.offset;
    op->b_off = e.b.offset;
    op->c_off = e.c.offset;
    op->a_elems = (int64_t)e.a.rows << e.a.rank;
    op->b_elems = (int64_t)e.b.rows << e.b.rank;
    const int64_t a_panel = align_up((sh.nb_a << (e.n_m + e.n_k)) * 8, 1024);
    const int64_t b_panel = align_up((sh.nb_b << (e.n_n + e.n_k)) * 16, 1024);
    op->ahi_off = e.scratch_offset;
    op->alo_off = op->ahi_off + a_panel;
    op->bhi_off = op->alo_off + a_panel;
    op->blo_off = op->bhi_off + b_panel;
    op->M = sh.M;
    op->N = sh.N;
    op->K = sh.K;
    op->batch = sh.batch;
    op->a_batched = sh.batch > 1 && e.rows_a != TNC_ROWS_NONE && !sh.outer;
    op->b_batched = sh.batch > 1 && e.rows_b != TNC_ROWS_NONE;
    op->n_inner = sh.n_inner;
    op->bn = bn;
    GemmArgs& g = op->args;
    g.c_batch_stride = sh.M * sh.N;
    g.outer_rj = sh.outer ? (int32_t)sh.nb_b : 0;
    g.pair_rows = (sh.outer && (e.flags & TNC_EINSUM_OUTER_PAIRS)) ? dev_pair_rows : nullptr;
    g.outer_mb = e.n_m;
    g.ldc = (int32_t)(sh.n_inner ? sh.n_inner : sh.N);
    g.n_inner = sh.n_inner;
    g.M = (int32_t)sh.M;
    g.N = (int32_t)sh.N;
    g.K = (int32_t)sh.K;
    g.m_tiles = (int32_t)((sh.M + BM - 1) / BM);
    g.n_tiles = (int32_t)((sh.N + op->bn - 1) / op->bn);
    // row tiles per sweep group: the ~74 resident tile pairs then cover 8 row panels x ~9 column panels
    // (measured on the fat GEMM of n53 m20, ncu: 143.6 GB of DRAM traffic against 168.3 GB at 16 and
    // 181.4 GB at 4, same time; the floor for 74 resident 256x256 tiles is ~64 GB)
    g.group_m = 8;
    if (const char* env = knob("TNC_TC_GROUP_M")) {  // experiment knob: row tiles per sweep group
        const int v = atoi(env);
        if (v >= 1 && v <= 1024) g.group_m = v;
    }
    g.a_batched = op->a_batched;
    g.b_batched = op->b_batched;
    g.kc = kDefaultKC[precision];
    if (const char* env = knob("TNC_TC_KC")) {       // experiment knob
        const int v = atoi(env);
        if (v >= 1 && v <= 64) g.kc = v;
    }
    op->tiles = (int64_t)g.m_tiles * g.n_tiles * sh.batch;
    if (op->tiles >= ((int64_t)1 << 31)) {
        delete op;
        set_error("tensor-core einsum: too many tiles");
        return TNC_ERR_UNSUPPORTED;
    }
    g.tiles = (int32_t)op->tiles;
    g.blocked = op->blocked;
    g.blocked_b = op->blocked_b;
    g.out_scale = op->use_3m ? 4.f : 1.f;
    // lockstep only pays off for long k loops shared by many tiles
    const int bk = bk_of(precision);
    const int nkb = (int)((sh.K + bk - 1) / bk);
    g.sync_every = nkb >= 64 ? 16 : 0;
    if (const char* env = knob("TNC_TC_SYNC")) g.sync_every = nkb >= 64 ? atoi(env) : 0;
    if (op->use_3m) {
        // k-blocks of 64 complex k; the barrier spacing is rounded to whole accumulation chunks
        g.sync_every = (g.sync_every / 2 + g.kc - 1) / g.kc * g.kc;
    }
    op->words_off = e.scratch_offset + need - 1024;
    *out = op;
    return TNC_OK;
}

int tc_gemm_run(TcGemmOp* op, char* ws, cudaStream_t s, LaunchHook hook, void* ctx, int* launches, const TcAmaxWords& ext) {
    const bool f16 = op->precision != TNC_TC_3XTF32;
    const bool lo = op->precision != TNC_TC_F16;
    const int elem = f16 ? 2 : 4;
    int n_launch = 3;
    uint32_t* words = (uint32_t*)(ws + op->words_off);
    PackDesc pa = op->pa, pb = op->pb;                  // the op itself stays read-only: a run only fills local copies
    // amax words: the step's own (scratch) unless the operand's producer already reduced one elsewhere
    const uint32_t* amax_a = ext.a >= 0 ? (const uint32_t*)(ws + ext.a) : words;
    const uint32_t* amax_b = ext.b >= 0 ? (const uint32_t*)(ws + ext.b) : words + 1;
    pa.amax = f16 ? amax_a : nullptr;
    pb.amax = f16 ? amax_b : nullptr;
    if (f16 || op->args.sync_every) TNC_CUDA(cudaMemsetAsync(words, 0, 256, s));
    if (f16 && (ext.a < 0 || ext.b < 0)) {
        // one launch finds the largest magnitude of both operands (whole source tensors: an upper
        // bound of the gathered rows is all the scaling needs); an operand whose word is external gets no blocks
        const int64_t n4_a = op->a_elems / 2, n4_b = op->b_elems / 2;       // float4 = two amplitudes; ranks >= 2
        const int64_t cap = (int64_t)sm_count() * 8;
        const int ga = ext.a >= 0 ? 0 : (int)std::max<int64_t>(1, std::min<int64_t>(cap, (n4_a + 511) / 512));
        const int gb = ext.b >= 0 ? 0 : (int)std::max<int64_t>(1, std::min<int64_t>(cap, (n4_b + 511) / 512));
        amax_kernel<<<ga + gb, 256, 0, s>>>((const float4*)(ws + op->a_off), n4_a, (const float4*)(ws + op->b_off), n4_b, ga,
                                            words);
        TNC_CUDA(cudaGetLastError());
        n_launch = 4;
    }
    int rc = launch_pack(pa, ws + op->a_off, ws + op->ahi_off, lo ? ws + op->alo_off : nullptr, s);
    if (rc != TNC_OK) return rc;
    if (hook) hook(ctx);
    rc = launch_pack(pb, ws + op->b_off, ws + op->bhi_off, lo ? ws + op->blo_off : nullptr, s);
    if (rc != TNC_OK) return rc;
    if (hook) hook(ctx);
    TcGemmOp::Maps maps;
    std::unique_lock<std::mutex> lock(op->mu);
    auto cached = op->maps_by_ws.find(ws);
    if (cached != op->maps_by_ws.end()) {
        maps = cached->second;
        lock.unlock();
    } else {
        lock.unlock();
        CUtensorMap* const om = maps.m;
        const int64_t batch_a = op->a_batched ? op->batch : 1, batch_b = op->b_batched ? op->batch : 1;
        const int box_b = op->two_cta ? 128 : op->bn;
        const int bk = bk_of(op->precision);
        char* lo_a = ws + (lo ? op->alo_off : op->ahi_off);
        char* lo_b = ws + (lo ? op->blo_off : op->bhi_off);
        const int64_t nkb = op->K / bk;
        if (op->use_3m) {
            // planar panels: blocks of [hi, lo][128 rows][64 halves], three parts per (tile, 64-k block)
            const int hl = lo ? 2 : 1;
            const int64_t nkb3 = op->K / 2 / 64;
            const int64_t blocks_a = nkb3 * op->args.m_tiles * batch_a, blocks_b = nkb3 * op->args.n_tiles * batch_b;
            if ((rc = make_map(&om[0], ws + op->ahi_off, 2, 64, 128 * hl, blocks_a * 3, 128 * hl)) != TNC_OK) return rc;
            if ((rc = make_map(&om[2], ws + op->bhi_off, 2, 64, 128, blocks_b * 3 * hl, 64)) != TNC_OK) return rc;
            om[1] = om[0];
            om[3] = om[2];
        } else if (op->blocked) {
            // `blocks` blocks of [rows][128 bytes], one per (tile, k-block)
            const int64_t blocks_a = nkb * op->args.m_tiles * batch_a;
            if ((rc = make_map(&om[0], ws + op->ahi_off, elem, bk, BM, blocks_a, BM)) != TNC_OK) return rc;
            if ((rc = make_map(&om[1], lo_a, elem, bk, BM, blocks_a, BM)) != TNC_OK) return rc;
        } else {
            // K-major panel [batch][rows][K]; folded rows are part of M
            if ((rc = make_map(&om[0], ws + op->ahi_off, elem, op->K, op->M, batch_a, BM)) != TNC_OK) return rc;
            if ((rc = make_map(&om[1], lo_a, elem, op->K, op->M, batch_a, BM)) != TNC_OK) return rc;
        }
        if (op->use_3m) {
        } else if (op->blocked_b) {
            const int64_t blocks_b = nkb * op->args.n_tiles * batch_b;
            if ((rc = make_map(&om[2], ws + op->bhi_off, elem, bk, op->bn, blocks_b, box_b)) != TNC_OK) return rc;
            if ((rc = make_map(&om[3], lo_b, elem, bk, op->bn, blocks_b, box_b)) != TNC_OK) return rc;
        } else {
            // [batch][rows][K]; a folded panel ([row of B][n][k], no batch) is one matrix of N rows
            if ((rc = make_map(&om[2], ws + op->bhi_off, elem, op->K, op->N, batch_b, box_b)) != TNC_OK) return rc;
            if ((rc = make_map(&om[3], lo_b, elem, op->K, op->N, batch_b, box_b)) != TNC_OK) return rc;
        }
        lock.lock();
        op->maps_by_ws.emplace(ws, maps);
        lock.unlock();
    }
    GemmArgs g = op->args;
    g.c = (float*)(ws + op->c_off);
    g.amax_a = f16 ? amax_a : nullptr;
    g.amax_b = f16 ? amax_b : nullptr;
    g.amax_out = ext.out >= 0 ? (uint32_t*)(ws + ext.out) : nullptr;
    g.sync_counter = words + 32;
    // persistent grid: one CTA per SM (the shared-memory footprint allows no more), every CTA
    // walks tiles c, c + grid, ...
    int64_t grid = std::min<int64_t>(op->tiles, sm_count());
    if (op->two_cta) grid &= ~(int64_t)1;               // whole pairs; tiles is even here
    g.rounds = (int32_t)((op->tiles + grid - 1) / grid);
    if (op->use_3m) {
        // 16 TMEM columns per tcgen05.ld batch: 32 spill inside the chunk loop at the kernel's 168 registers
        rc = op->gauss ? launch_gemm3m<TNC_TC_3XF16, 16, true>(maps.m, g, grid, s)
                       : launch_gemm3m<TNC_TC_3XF16, 16, false>(maps.m, g, grid, s);
    } else if (op->two_cta) {
        switch (op->precision) {
            case TNC_TC_3XF16: rc = launch_gemm_2cta<TNC_TC_3XF16>(maps.m, g, grid, s); break;
            case TNC_TC_F16: rc = launch_gemm_2cta<TNC_TC_F16>(maps.m, g, grid, s); break;
            default: rc = launch_gemm_2cta<TNC_TC_3XTF32>(maps.m, g, grid, s); break;
        }
    } else {
        switch (op->precision) {
            case TNC_TC_3XF16: rc = launch_gemm_bn<TNC_TC_3XF16>(op->bn, maps.m, g, grid, s); break;
            case TNC_TC_F16: rc = launch_gemm_bn<TNC_TC_F16>(op->bn, maps.m, g, grid, s); break;
            default: rc = launch_gemm_bn<TNC_TC_3XTF32>(op->bn, maps.m, g, grid, s); break;
        }
    }
    if (rc != TNC_OK) return rc;
    if (hook) hook(ctx);
    if (launches) *launches = n_launch;
    return TNC_OK;
}

bool tc_gemm_emits_amax(const TcGemmOp* op) {
    return op != nullptr;      // every GEMM kernel's epilogue can
}

void tc_gemm_destroy(TcGemmOp* op) {
    delete op;
}

}  // namespace tnc
