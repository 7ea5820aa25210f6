// Tensor-core (tcgen05) lowering of one einsum step.
#pragma once
#include "tnc_internal.h"

namespace tnc {

struct TcGemmOp;

// Builds the device-side tables for running `e` on the tcgen05 path.  `rows_a` / `rows_b`
// are the host copies of the row tables (nullptr for TNC_ROWS_NONE / TNC_ROWS_IDENTITY).
int tc_gemm_create(const tnc_einsum& e, int dtype, const int32_t* rows_a, const int32_t* rows_b, TcGemmOp** out);
int tc_gemm_run(TcGemmOp* op, const void* a, const void* b, void* c, cudaStream_t s, int* launches);
void tc_gemm_destroy(TcGemmOp* op);

}  // namespace tnc
