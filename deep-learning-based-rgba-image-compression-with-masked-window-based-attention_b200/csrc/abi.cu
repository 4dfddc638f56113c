// C-ABI bookkeeping: version, status strings, last CUDA error.
#include <string.h>
#include <stdio.h>
#include "status.cuh"

namespace b200 {
static thread_local char g_last_error[256] = "";
void record_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}
}  // namespace b200

extern "C" {

int mwa_b200_abi_version(void) { return MWA_B200_ABI_VERSION; }

const char* mwa_b200_status_string(int status) {
    switch (status) {
        case MWA_OK: return "ok";
        case MWA_ERR_INVALID: return "invalid argument";
        case MWA_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        case MWA_ERR_ALIGNMENT: return "pointer not 16-byte aligned";
        case MWA_ERR_WORKSPACE: return "workspace / parameter block too small";
        case MWA_ERR_CUDA: return "CUDA runtime error (see mwa_b200_last_cuda_error)";
        default: return "unknown status";
    }
}

const char* mwa_b200_last_cuda_error(void) { return b200::g_last_error; }

}  // extern "C"
