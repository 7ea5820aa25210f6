// Tensor-core (tcgen05) lowering of one einsum step -- placeholder until the kernel lands.
#include "tc_gemm.h"

namespace tnc {

struct TcGemmOp {};

int tc_gemm_create(const tnc_einsum&, int, const int32_t*, const int32_t*, TcGemmOp**) {
    set_error("einsum: TNC_ALGO_TC is not available in this build");
    return TNC_ERR_UNSUPPORTED;
}
int tc_gemm_run(TcGemmOp*, const void*, const void*, void*, cudaStream_t, int*) { return TNC_ERR_UNSUPPORTED; }
void tc_gemm_destroy(TcGemmOp* op) { delete op; }

}  // namespace tnc
