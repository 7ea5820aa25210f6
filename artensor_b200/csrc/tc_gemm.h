// Tensor-core (tcgen05) lowering of one einsum step: two pack launches (bit-permutation of
// each operand into a K-major panel, split into hi/lo parts; the right operand is also
// expanded to its real 2N x 2K form) followed by one TMA-fed tcgen05 GEMM.  The fp16
// precisions add one amax launch in front (power-of-two operand scaling) unless the kernels that
// produced the operands already reduced it (TcAmaxWords).
#pragma once
#include "tnc_internal.h"

namespace tnc {

struct TcGemmOp;

// Called after every kernel launch of an operation (profiling); may be null.
typedef void (*LaunchHook)(void* ctx);

// Bytes of scratch the lowering needs inside the workspace; <0 if the step cannot be lowered
// (reason in tnc_last_error()).
int64_t tc_gemm_scratch_bytes(const tnc_einsum& e, int dtype);

// `dev_rows_a` / `dev_rows_b`: device copies of the row tables (nullptr unless rows_* >= 0).
// `precision` is a tnc_tc_precision.
// `dev_pair_rows` (TNC_EINSUM_OUTER_PAIRS only): device table, C row block of the pair (ra * b.rows + rb)
int tc_gemm_create(const tnc_einsum& e, int dtype, int precision, const int32_t* dev_rows_a, const int32_t* dev_rows_b,
                   const int32_t* dev_pair_rows,
                   TcGemmOp** out);
// Workspace byte offsets of amax words kept OUTSIDE the step's scratch (-1: none).  `a` / `b`: the operand's
// largest magnitude was already reduced there by the kernel that produced the operand, the step skips its own
// pass over it; `out`: the GEMM epilogue reduces the largest magnitude of C there for the step that consumes C.
struct TcAmaxWords {
    int64_t a = -1, b = -1, out = -1;
};
int tc_gemm_run(TcGemmOp* op, char* workspace, cudaStream_t s, LaunchHook hook, void* ctx, int* launches,
                const TcAmaxWords& ext = TcAmaxWords());
// whether the step's GEMM kernel can reduce the amax of its output (TcAmaxWords::out)
bool tc_gemm_emits_amax(const TcGemmOp* op);
void tc_gemm_destroy(TcGemmOp* op);

// ---------------------------------------------------------------- pack (bit-permutation) kernel
// dst[b][q] = f(src[row(b)][p]) where bit i of q is bit src_pos[i] of p; tiled through shared
// memory so that both the global reads and the global writes are contiguous runs.
enum PackMode { PACK_COPY = 0, PACK_SPLIT = 1, PACK_EXPAND_SPLIT = 2, PACK_SPLIT_F16 = 3, PACK_EXPAND_SPLIT_F16 = 4,
                PACK_ACCUM = 5 /* dst += permuted src (fast kernel only: rank >= 8) */,
                PACK_PLANAR3_F16 = 6 /* panels of the 3M complex product (fast kernel only: rank >= 13): per 2^13
                                        amplitudes of the destination order, the fp16 planes re, im, re + im -- each
                                        as hi then lo when dst_lo is non-null -- of 2^13 halves each, scaled by half
                                        the fp16 operand scale (re + im must stay in range) */ };
struct PackDesc {
    int32_t rank;                 // bits per block (source and destination)
    int32_t nb;                   // destination blocks
    int32_t rows_mode;            // TNC_ROWS_NONE / TNC_ROWS_IDENTITY / >= 0 (table)
    const int32_t* rows;          // device table when rows_mode >= 0
    int32_t mode;                 // PackMode
    int32_t inner_bits;           // PACK_EXPAND_SPLIT: number of low destination bits that are k
    int32_t blocked;              // PACK_EXPAND_SPLIT: write tile-contiguous blocks [n tile][k block][rows][16 k]
    int32_t bn_log2;              //   log2 of the rows (2n + c') per n tile
    int32_t kb_log2;              //   log2 of the complex k per k-block: 4 (tf32, default) or 5 (fp16)
    const uint32_t* amax;         // fp16 modes: device word with the bits of the operand's largest magnitude;
                                  //   dst_lo may be null (hi part only)
    int32_t planes;               // PACK_PLANAR3_F16: which three planes -- 0: re, im, re + im (Karatsuba, both operands);
                                  //   1: re + im, re, im (Gauss, left operand); 2: re, im - re, re + im (Gauss, right operand)
    int8_t src_pos[TNC_MAX_BITS]; // source position feeding destination position i
};
int launch_pack(const PackDesc& d, const void* src, void* dst_hi, void* dst_lo, cudaStream_t s);

}  // namespace tnc
