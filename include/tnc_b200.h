/* tnc_b200 -- C ABI of the B200-native tensor-network contraction executor.
 *
 * This library replaces what artensor's numerical executor dispatches to: the chain of
 * `torch.einsum` calls in `tensor_contraction` (artensor/contraction.py:62-76) and
 * `tensor_contraction_sparse` (artensor/contraction.py:132-205), and the slice loop of
 * `TensorNetworkSimulation.contraction` (artensor/simulation.py:103-117; copy at :198-213).
 * The reference is pure Python and has no FFI of its own; the binding a maintainer would
 * add is the ctypes stub shown in INTEGRATION.md (it is `artensor_b200/_native.py`).
 *
 * Model: a *plan* is an immutable list of operations over one device workspace ("arena").
 * The host-side plan compiler (artensor_b200/plan.py, backend.py) lowers a reference-format
 * scheme into operations; the library runs them for a range of slice ids and accumulates
 * the per-slice results into a caller-owned accumulator.
 *
 * Every tensor is "bits": `rows` blocks (the bitstring batch mode of sparse schemes,
 * contraction.py:219-220; 1 if absent) of 2^rank elements; bit p of the in-block element
 * index is "position p".  A mode order is an assignment of modes to positions.
 *
 * All pointers are plain device pointers; no torch types cross this boundary.
 * All functions return 0 on success, a tnc_status otherwise; tnc_last_error() describes the
 * last failure on the calling thread.  The reference's error convention for this path is
 * print + sys.exit(1) (contraction.py:71-74, :192-195); here errors are returned.
 */
#ifndef TNC_B200_H
#define TNC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNC_ABI_VERSION 7
#define TNC_MAX_BITS 40          /* max bit modes per group / per tensor */
#define TNC_MAX_SLICED 8         /* max sliced bonds on one leaf */
#define TNC_WORKSPACE_TAIL_BYTES 4352   /* the library's own words behind the arena: tnc_plan_workspace_bytes() =
                                           arena rounded up to 256 + this (amax words reduced by producing kernels,
                                           the slice-id word of CUDA-graph replay) */

typedef enum tnc_status {
    TNC_OK = 0,
    TNC_ERR_INVALID = 1,         /* malformed operation / argument */
    TNC_ERR_CUDA = 2,            /* a CUDA call failed */
    TNC_ERR_NOMEM = 3,           /* workspace too small */
    TNC_ERR_UNSUPPORTED = 4,     /* valid but not implemented for these sizes */
    TNC_ERR_STATE = 5            /* plan not finalized / already finalized */
} tnc_status;

/* Storage type of the tensors in the arena.  complex64 only: the reduced-precision complex-half
 * mode (torch.complex32 at the Python boundary) is an OPERAND precision of the tensor-core steps
 * (TNC_TC_F16 below) -- intermediates stay complex64 because n53 amplitudes (~1e-8) are below the
 * fp16 range (the reference needs its per-step rescaling for them, contraction.py:197-200).  ABI 5
 * declared a TNC_C32 storage type that no tensor-core path ever implemented; it was removed. */
typedef enum tnc_dtype {
    TNC_C64 = 0                  /* complex64 (interleaved fp32 pairs) */
} tnc_dtype;

typedef enum tnc_phase {
    TNC_PHASE_ONCE = 0,          /* slice-invariant: runs once per execute call */
    TNC_PHASE_SLICE = 1          /* runs for every slice id */
} tnc_phase;

/* Operand precision of the tensor-core steps of a complex64 plan (fp32 accumulation always). */
typedef enum tnc_tc_precision {
    TNC_TC_3XTF32 = 0,           /* fp32-accurate: TF32 hi/lo split, 3 MMAs per useful one */
    TNC_TC_3XF16 = 1,            /* fp32-accurate: fp16 hi/lo split with power-of-two operand scaling,
                                    3 MMAs per useful one at twice the TF32 rate (default) */
    TNC_TC_F16 = 2               /* reduced precision: fp16 operands, 1 MMA per useful one -- the
                                    complex-half tensor-core mode of Pan et al. 2023 */
} tnc_tc_precision;

typedef enum tnc_option {
    TNC_OPT_TC_PRECISION = 0,    /* value: tnc_tc_precision */
    TNC_OPT_CUDA_GRAPH = 1,      /* value 1: tnc_plan_execute replays the slice phase as ONE CUDA graph per slice
                                    (captured on first use per workspace / leaf blob / accumulator) instead of
                                    ~100 stream launches -- for slices of a few ms, where launch gaps are a
                                    measurable share.  The slice-id word lives in the workspace tail. */
    TNC_OPT_FUSE_AMAX = 2,       /* value 1 (default): when the operand of a tensor-core step (fp16 precisions) was
                                    written by a streaming or GEMM kernel of the same phase, that kernel reduces the
                                    operand's largest magnitude (its power-of-two scale) into a word of the workspace
                                    tail while it stores the tensor, and the step's own amax pass skips the operand.
                                    Results are bit-identical either way; 0 keeps the separate pass (A/B aid). */
    TNC_OPT_SLICE_REUSE = 3      /* value 1: inside one tnc_plan_execute call, after its first slice, a SLICE-phase
                                    operation runs again only when a sliced bond its operands depend on changed from
                                    the previous slice id (the library derives the dependencies from the leaf records).
                                    The reference's loop (simulation.py:107-114) recomputes the whole tree per slice;
                                    most of a deep tree depends on few of the sliced bonds.  Results are bit-identical.
                                    Slice ids are walked in ascending order, so going from s - 1 to s changes the
                                    slice-id bits 0 .. ctz(s): an operation runs when the LOWEST bit behind it is among
                                    them (or, with TNC_EINSUM_RUN_WITH_READER, whenever its reader runs).
                                    Asks of the caller's layout, checked at finalize (TNC_ERR_INVALID): the result of
                                    an operation whose reader runs on other slices than it does (their lowest bits
                                    differ) is read again in later slices, so no other operation of the phase may write
                                    over it (tensor or scratch).
                                    With TNC_OPT_CUDA_GRAPH the first slice of a call runs as plain launches and every
                                    later slice replays the graph of its class (the operations that run when the bits
                                    0 .. ctz(s) changed; at most one graph per sliced bond, captured on first use).
                                    Default 0. */
} tnc_option;

typedef enum tnc_algo {
    TNC_ALGO_SIMT = 0,           /* generic CUDA-core kernel, any shape */
    TNC_ALGO_TC = 1,             /* tcgen05 tensor-core kernel (split-precision for c64): compute-bound steps */
    TNC_ALGO_STEM = 2,           /* streaming fp32 kernel for HBM-bound steps (tiny right operand) */
    TNC_ALGO_SKINNY = 3          /* streaming tcgen05 kernel: A read in place (no pack pass), small right
                                    operand resident in shared memory; 4..32 flop/byte steps */
} tnc_algo;

/* tnc_einsum.flags.  OUTER_ROWS: the output rows enumerate ALL pairs of operand rows, A-major
 * (row b = ra * b.rows + rb; the reference's two-batch-label einsum + reshape,
 * artensor/contraction.py:180-185).  The row tables must say the same; the flag lets the
 * tensor-core path fold A's rows into M and loop over B's rows instead of gathering copies. */
#define TNC_EINSUM_OUTER_ROWS 1
/* OUTER_PAIRS: the output rows enumerate ALL pairs of operand rows exactly once, in the order the
 * row tables give (the reference's chunked batched steps whose wanted bitstrings are the full
 * product of the operands' rows, artensor/contraction.py:272-300: sorted by the row of the larger
 * operand, not A-major).  The tensor-core path then packs every operand row ONCE, folds A's rows
 * into M and B's rows into N, and scatters the row blocks of C through the inverse of the tables;
 * without the flag the same step gathers a.rows * b.rows copies of the operand rows. */
#define TNC_EINSUM_OUTER_PAIRS 2
/* RUN_WITH_READER (only read with TNC_OPT_SLICE_REUSE): the operation's result is not preserved across slices --
 * it runs whenever the operation that reads its result runs (instead of only when a sliced bond behind its own
 * operands changed), so its result needs no memory of its own.  Trades recomputation for workspace. */
#define TNC_EINSUM_RUN_WITH_READER 4

/* Row table ids: a plan-owned int32 table (tnc_plan_add_table) or one of these. */
#define TNC_ROWS_NONE (-1)       /* operand has no row mode: always block 0 */
#define TNC_ROWS_IDENTITY (-2)   /* source row == output row */

typedef struct tnc_plan tnc_plan;

/* A tensor living in the arena. */
typedef struct tnc_tensor {
    int64_t offset;              /* byte offset into the workspace; multiple of 256 */
    int32_t rank;                /* number of bit modes: a row block has 2^rank elements */
    int32_t rows;                /* number of row blocks (>= 1) */
} tnc_tensor;

/* C[b][m,n,h] = sum_k A[ra[b]][m,k,h] * B[rb[b]][k,n,h]   (one step of a scheme;
 * artensor/contraction.py:70 and :147-190 are the einsum call sites this replaces).
 * Each mode is given by its bit position in the operands that carry it. */
typedef struct tnc_einsum {
    tnc_tensor a, b, c;
    int32_t nb;                  /* output rows == c.rows */
    int32_t rows_a;              /* row table id for A (nb entries) or TNC_ROWS_* */
    int32_t rows_b;
    int32_t n_m, n_n, n_k, n_h;
    int8_t m_a[TNC_MAX_BITS], m_c[TNC_MAX_BITS];
    int8_t n_b[TNC_MAX_BITS], n_c[TNC_MAX_BITS];
    int8_t k_a[TNC_MAX_BITS], k_b[TNC_MAX_BITS];
    int8_t h_a[TNC_MAX_BITS], h_b[TNC_MAX_BITS], h_c[TNC_MAX_BITS];
    int32_t algo;                /* tnc_algo */
    int32_t flags;               /* TNC_EINSUM_* bits */
    int64_t scratch_offset;      /* TNC_ALGO_TC: byte offset of a scratch region of              */
    int64_t scratch_bytes;       /* tnc_einsum_tc_scratch_bytes() bytes inside the workspace      */
} tnc_einsum;

/* dst[r][q] = src[r][p] where bit i of q equals bit perm[i] of p (a bit permutation of the
 * in-block index; replaces the permute+reshape copies torch.einsum makes). */
typedef struct tnc_permute {
    tnc_tensor src, dst;
    int8_t perm[TNC_MAX_BITS];   /* perm[i] = source position feeding destination position i */
} tnc_permute;

/* One leaf tensor copied from the caller's packed leaf blob into the arena, with its sliced
 * bonds fixed by the slice id (artensor/simulation.py:108-113: select(ind, bit).clone()).
 * Slice-id bit numbering follows np.binary_repr(s, S): bond x is bit (S-1-x). */
typedef struct tnc_leaf {
    int64_t src_offset;          /* element offset of the leaf inside the leaf blob */
    tnc_tensor dst;              /* rank = un-sliced rank - n_sliced */
    int32_t src_rank;            /* bit modes of the stored leaf (rows excluded) */
    int32_t n_sliced;
    int8_t sliced_pos[TNC_MAX_SLICED];   /* source bit position of each sliced mode */
    int8_t sliced_bond[TNC_MAX_SLICED];  /* which slice-id bit (x, MSB-first index) drives it */
    int8_t keep_pos[TNC_MAX_BITS];       /* source position of destination position i */
} tnc_leaf;

/* out[dst_index(e)] += src[e]: adds a per-slice result into the caller's accumulator
 * (artensor/simulation.py:114 `collect_tensor += ...`); out_pos lets the accumulator use
 * a different mode order (e.g. qubit order, simulation.py:115-116). */
typedef struct tnc_accum {
    tnc_tensor src;
    int8_t out_pos[TNC_MAX_BITS];        /* accumulator position of source position i */
} tnc_accum;

/* Scratch the tensor-core lowering of `e` needs (packed hi/lo operand panels); 0 on error. */
int64_t tnc_einsum_tc_scratch_bytes(int32_t dtype, const tnc_einsum* e);

/* ---- plan construction (host only, no CUDA calls until finalize) ---- */
int tnc_abi_version(void);
int tnc_plan_create(int32_t dtype, int32_t n_sliced_bonds, tnc_plan** out);
/* Before finalize.  Unknown options / values return TNC_ERR_INVALID. */
int tnc_plan_set_option(tnc_plan* plan, int32_t option, int64_t value);
int tnc_plan_add_table(tnc_plan* plan, const int32_t* data, int64_t n, int32_t* table_id);
int tnc_plan_add_leaves(tnc_plan* plan, int32_t phase, const tnc_leaf* leaves, int32_t n);
int tnc_plan_add_einsum(tnc_plan* plan, int32_t phase, const tnc_einsum* op);
int tnc_plan_add_permute(tnc_plan* plan, int32_t phase, const tnc_permute* op);
int tnc_plan_add_accum(tnc_plan* plan, int32_t phase, const tnc_accum* op);
/* Declares the arena size the operations were laid out for (every tensor and scratch region must fit) and
 * uploads tables.  The workspace a caller passes to execute is larger: tnc_plan_workspace_bytes() = the arena
 * rounded up to 256 bytes + TNC_WORKSPACE_TAIL_BYTES. */
int tnc_plan_finalize(tnc_plan* plan, int64_t arena_bytes);
int64_t tnc_plan_workspace_bytes(const tnc_plan* plan);
int64_t tnc_plan_num_ops(const tnc_plan* plan, int32_t phase);
/* kernels launched by the last tnc_plan_execute on this plan */
int64_t tnc_plan_last_launches(const tnc_plan* plan);
/* tensor-core operands of `phase` whose amax is reduced by the kernel that produces them (TNC_OPT_FUSE_AMAX) */
int64_t tnc_plan_num_fused_amax(const tnc_plan* plan, int32_t phase);
void tnc_plan_destroy(tnc_plan* plan);

/* ---- execution ----
 * Runs ONCE-phase operations, then SLICE-phase operations for every slice id in
 * [slice_begin, slice_end), on `stream` (a cudaStream_t passed as void*), without host
 * synchronisation.  `leaf_blob`, `accum_out` and `workspace` are device pointers;
 * `accum_out` is read-modify-written (the caller zeroes it, simulation.py:101-105).
 * A finalized plan is immutable on the device and every word a launch writes lives in the caller's
 * workspace, so a plan may be used with any number of workspaces and, from one host thread, on
 * several streams with a workspace (and an accum_out) each. */
int tnc_plan_execute(tnc_plan* plan, const void* leaf_blob, uint64_t slice_begin,
                     uint64_t slice_end, void* accum_out, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* Same work as tnc_plan_execute for ONE slice id, but every kernel launch is bracketed by CUDA
 * events on `stream` and the call synchronises.  ms_once / ms_slice hold TNC_PROFILE_SLOTS
 * floats per operation (arrays sized TNC_PROFILE_SLOTS * tnc_plan_num_ops): slot 0 is the
 * whole operation, slots 1.. are its individual launches in order (for a TNC_ALGO_TC einsum:
 * pack A, pack B, tcgen05 GEMM), unused slots are 0.  Measurement aid for bench.py's roofline;
 * not used on the product path. */
#define TNC_PROFILE_SLOTS 4
int tnc_plan_profile(tnc_plan* plan, const void* leaf_blob, uint64_t slice_id, void* accum_out,
                     void* workspace, int64_t workspace_bytes, void* stream,
                     float* ms_once, float* ms_slice);

/* Stand-alone bit permutation (same kernel the plan uses); elem_bytes is 8 (c64) or 4. */
int tnc_permute_bits(const void* src, void* dst, int32_t rank, int64_t rows,
                     const int8_t* perm, int32_t elem_bytes, void* stream);

const char* tnc_last_error(void);

/* ---- environment ----
 * TNC_NVTX=1          every operation runs inside an NVTX range named after its class and shape
 *                     ("tc m15 n13 k15 rows1", "skinny ...", "chain x13", "leaves", "accum").
 *
 * Experiment knobs: kernel variants for A/B measurements (tools/, profiles/).  They are honoured
 * ONLY together with TNC_EXPERIMENTS=1 and are not part of the contract; some change rounding
 * (never beyond the documented tolerances' order of magnitude) -- the product path runs without.
 *   TNC_TC_3M=0          interleaved 4M complex product instead of the 3M (Karatsuba) kernel
 *   TNC_TC_2CTA=0        single-CTA tiles instead of cta_group::2 pairs (also disables 3M)
 *   TNC_TC_KC=<n>        k-blocks accumulated in tensor memory before the fp32 register add
 *                        (default 1; more = faster, larger round-toward-zero bias)
 *   TNC_TC_GAUSS=1       3M product in Gauss's form (planes s, r, i / r, i - r, s) instead of Karatsuba's
 *   TNC_TC_SYNC=<n>      k-blocks between the grid-wide lockstep barriers of a GEMM (0 = none)
 *   TNC_TC_GROUP_M=<n>   row tiles per sweep group of the persistent tile order
 *   TNC_TC_BLOCKED=0     row-major instead of tile-contiguous packed panels
 *   TNC_TC_FOLDN=0       outer-row steps loop over B's rows instead of folding them into N
 *   TNC_PACK_FAST=0      generic pack kernel instead of the register-resident one
 *   TNC_STEM_NO_BULK=1   per-thread streaming kernel instead of the bulk-copy one
 *   TNC_STEM_BULK_CTAS=<2|3>  resident CTAs per SM of the bulk-copy streaming kernel
 *   TNC_NO_CHAIN=1       one launch per tiny generic step instead of chained launches
 *   TNC_NO_ROWDOT=1      generic kernel instead of the row-dot kernel for long contractions
 *   TNC_ROWDOT_THREADS=128  narrow CTAs of the row-dot kernel also for few rows x a very long contraction
 */

#ifdef __cplusplus
}
#endif
#endif /* TNC_B200_H */
