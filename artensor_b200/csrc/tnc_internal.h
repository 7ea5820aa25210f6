// Internal declarations shared by the tnc_b200 translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "tnc_b200.h"

// Read-only loads of operands: the non-coherent path by default.  -DTNC_COHERENT_LOADS (together
// with -D__restrict__=) builds the library with L2-coherent loads instead: the experiment behind
// DESIGN.md's note on concurrent executions of one plan.
#ifdef TNC_COHERENT_LOADS
#define TNC_LDG(p) __ldcg(p)
#else
#define TNC_LDG(p) __ldg(p)
#endif

namespace tnc {

void set_error(const char* fmt, ...);

// Experiment knobs (listed in include/tnc_b200.h): environment variables that select kernel variants
// for A/B measurements.  They are read ONLY when TNC_EXPERIMENTS=1 is set as well, so that a stray
// variable can never change the numerics or the speed of the product path silently.
inline const char* knob(const char* name) {
    const char* on = getenv("TNC_EXPERIMENTS");
    if (!on || on[0] == '\0' || on[0] == '0') return nullptr;
    return getenv(name);
}

constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
int cuda_fail(cudaError_t e, const char* what);

#define TNC_CUDA(call)                                              \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return ::tnc::cuda_fail(_e, #call);  \
    } while (0)

// ---------------------------------------------------------------- generic (SIMT) einsum
struct SimtEinsumParams {
    const void* a;
    const void* b;
    void* c;
    const int32_t* rows_a;   // table or nullptr
    const int32_t* rows_b;
    int32_t rows_mode_a;     // TNC_ROWS_NONE / TNC_ROWS_IDENTITY / >=0 (table)
    int32_t rows_mode_b;
    int32_t rank_a, rank_b, rank_c;
    int32_t kb;
    int64_t total;           // nb << rank_c
    const uint32_t* koff_a;  // 2^kb entries
    const uint32_t* koff_b;
    int8_t c2a[TNC_MAX_BITS];  // for C position p: position in A, or -1
    int8_t c2b[TNC_MAX_BITS];
};
int launch_simt_einsum(const SimtEinsumParams& p, int dtype, cudaStream_t s);
bool simt_uses_rowdot(int rank_c, int kb, int64_t total, int n_m, int n_n, int n_h);

// A run of consecutive tiny generic steps executed by ONE launch (one CTA walks the run, a block
// barrier between steps) instead of one ~6 us launch per step.  Records live in the plan's device
// blob; operand addresses are offsets into the workspace given at execute time.
struct ChainStep {
    int64_t a_off, b_off, c_off;               // bytes into the workspace
    int64_t rows_a_off, rows_b_off;            // bytes into the device blob (unused for NONE / IDENTITY)
    int64_t koff_a_off, koff_b_off;            // bytes into the device blob
    int64_t total;                             // nb << rank_c
    int32_t rows_mode_a, rows_mode_b;
    int32_t rank_a, rank_b, rank_c, kb;
    int8_t c2a[TNC_MAX_BITS];
    int8_t c2b[TNC_MAX_BITS];
};
constexpr int64_t kChainMaxOutputs = 1 << 13;  // per step
constexpr int64_t kChainMaxMacs = 1 << 19;     // per step (outputs << contracted bits)
int launch_simt_chain(const ChainStep* dev_steps, int n, void* workspace, const void* dev_blob, cudaStream_t s);

// ---------------------------------------------------------------- streaming ("stem") einsum
// HBM-bound steps: huge left operand, tiny right operand; output written as C[rows][m][n].
bool stem_supported(const tnc_einsum& e, int dtype);
// `dev_seg_begin` (n_seg + 1 batch indices, device) splits the batches into runs that share their
// row of A (built by tnc_plan_finalize from the A-row table); nullptr: one run per batch, or a
// single run when A has no rows.
// `amax_out` (device word, zeroed by the caller) != nullptr: the kernel also leaves the bits of the largest
// |component| it wrote there -- the operand scale of a consuming tensor-core step, for free.
int launch_stem(const tnc_einsum& e, const void* a, const void* b, void* c, const int32_t* dev_rows_a,
                const int32_t* dev_rows_b, const int32_t* dev_seg_begin, int n_seg, cudaStream_t s,
                uint32_t* amax_out = nullptr);

// ---------------------------------------------------------------- streaming tensor-core ("skinny") einsum
// Same operand shapes and output layout as the streaming kernel, multiplied on tcgen05 (skinny.cu).
bool skinny_supported(const tnc_einsum& e, int dtype, int precision);
int launch_skinny(const tnc_einsum& e, int precision, const void* a, const void* b, void* c, const int32_t* dev_rows_a,
                  const int32_t* dev_rows_b, cudaStream_t s, uint32_t* amax_out = nullptr);

// ---------------------------------------------------------------- leaves
struct LeafDev {
    int64_t src_offset;      // elements
    int64_t dst_offset;      // bytes
    int32_t dst_rank, dst_rows, src_rank, n_sliced;
    int8_t sliced_pos[TNC_MAX_SLICED];
    int8_t sliced_shift[TNC_MAX_SLICED];   // slice-id bit index (LSB-based)
    int8_t keep_pos[TNC_MAX_BITS];
};
// `slice_word` (nullable): device word holding the slice id instead of the argument (CUDA-graph replay)
int launch_leaf_gather(const LeafDev* dev_leaves, int n, int max_elems, const void* blob,
                       void* arena, uint64_t slice_id, const uint64_t* slice_word, int dtype, cudaStream_t s);
// *word = value, or *word += value (one thread)
int launch_slice_word(uint64_t* word, uint64_t value, bool add, cudaStream_t s);

// ---------------------------------------------------------------- permute / accumulate
struct PermuteParams {
    const void* src;
    void* dst;
    int32_t rank;
    int64_t rows;
    int8_t perm[TNC_MAX_BITS];   // perm[i]: source position feeding destination position i
};
int launch_permute(const PermuteParams& p, int elem_bytes, cudaStream_t s);

struct AccumParams {
    const void* src;
    void* out;
    int32_t rank;
    int64_t rows;
    int8_t out_pos[TNC_MAX_BITS];
};
int launch_accum(const AccumParams& p, int dtype, cudaStream_t s);

}  // namespace tnc
