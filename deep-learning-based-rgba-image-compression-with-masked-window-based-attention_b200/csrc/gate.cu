// Gate + residual of the attention wrapper: out = a * sigmoid(b) + x in ONE pass.
// Reference semantics: layers/Masked_Attention.py:186-188 (Win_noShift_Attention.forward:  out = a * torch.sigmoid(b);
// out += identity), which the reference runs as three elementwise kernels (sigmoid, mul, add_: 8 tensor passes); here
// 3 reads + 1 write, 128-bit vectorised, grid sized in multiples of the SM count.  HBM-bound: 16 B per element.
// Backward: grad_a = g * s, grad_b = g * a * s * (1 - s), grad_x = g (passed through by the caller), s = sigmoid(b).
#include "common.cuh"
#include "status.cuh"

namespace b200 {
namespace {

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void __launch_bounds__(256)
gate_fwd_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ x,
                float4* __restrict__ out, int64_t n4) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
        const float4 va = __ldcs(a + i), vb = __ldcs(b + i), vx = __ldcs(x + i);
        float4 o;
        o.x = va.x * sigmoidf_(vb.x) + vx.x;
        o.y = va.y * sigmoidf_(vb.y) + vx.y;
        o.z = va.z * sigmoidf_(vb.z) + vx.z;
        o.w = va.w * sigmoidf_(vb.w) + vx.w;
        __stcs(out + i, o);
    }
}
__global__ void __launch_bounds__(256)
gate_fwd_tail_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ x,
                     float* __restrict__ out, int64_t lo, int64_t n) {
    const int64_t i = lo + blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = a[i] * sigmoidf_(b[i]) + x[i];
}
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ g,
                float* __restrict__ ga, float* __restrict__ gb, int64_t n) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const float s = sigmoidf_(__ldcs(b + i)), gv = __ldcs(g + i);
        ga[i] = gv * s;
        gb[i] = gv * __ldcs(a + i) * s * (1.f - s);
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int gate_residual_forward(const float* a, const float* b, const float* x, float* out, int64_t n, void* stream) {
    if (!a || !b || !x || !out || n < 0) return MWA_ERR_INVALID;
    if (n == 0) return MWA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = aligned16(a) && aligned16(b) && aligned16(x) && aligned16(out);
    const int64_t n4 = vec ? n / 4 : 0;
    if (n4 > 0) {
        gate_fwd_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                     reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), n4);
        int rc = check_launch("gate_residual_forward");
        if (rc != MWA_OK) return rc;
    }
    if (n4 * 4 < n) {
        const int64_t rest = n - n4 * 4;
        gate_fwd_tail_kernel<<<static_cast<unsigned>((rest + 255) / 256), 256, 0, st>>>(a, b, x, out, n4 * 4, n);
        return check_launch("gate_residual_forward(tail)");
    }
    return MWA_OK;
}

int gate_residual_backward(const float* a, const float* b, const float* grad_out, float* grad_a, float* grad_b, int64_t n,
                           void* stream) {
    if (!a || !b || !grad_out || !grad_a || !grad_b || n < 0) return MWA_ERR_INVALID;
    if (n == 0) return MWA_OK;
    gate_bwd_kernel<<<kNumSMs * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, grad_out, grad_a, grad_b, n);
    return check_launch("gate_residual_backward");
}

}  // extern "C"
