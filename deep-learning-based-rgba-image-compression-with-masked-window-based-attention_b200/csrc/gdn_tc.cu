// placeholder until the tcgen05 GDN kernel lands
#include "common.cuh"
#include "params.cuh"
#include "status.cuh"
namespace b200 {
bool gdn_tc_supported(int, int64_t, int) { return false; }
int gdn_forward_tc(const float*, float*, const void*, int64_t, int, int64_t, int, int, cudaStream_t) { return MWA_ERR_UNSUPPORTED; }
}
