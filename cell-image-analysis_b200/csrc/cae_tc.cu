// cae_tc.cu -- K3 tensor-core path (tcgen05 implicit GEMM).  Placeholder until the
// UMMA kernels land: refuses loudly instead of falling back.
#include "common.cuh"

int k_cae_tc_prepare(cia_ctx* h, int which) { (void)h; (void)which; return CIA_OK; }

int k_cae_forward_tc(cia_ctx* h, const float*, int, const int32_t*, float*, float*, float*, cudaStream_t) {
    h->err = "cia_cae_forward: tensor-core path not built in this library version";
    return CIA_E_STATE;
}
