// placeholder until the tcgen05 attention kernel lands
#include "common.cuh"
#include "params.cuh"
#include "status.cuh"
namespace b200 {
bool mwa_tc_supported(int, int, int, int, int, int, int) { return false; }
int mwa_forward_tc(const float*, const float*, float*, const void*, int, int, int, int, int, int, int, int, int32_t*, cudaStream_t) { return MWA_ERR_UNSUPPORTED; }
void mwa_tc_prepare_images(const float*, const float*, const float*, const float*, int, int, int, float, uint8_t*, cudaStream_t) {}
}
