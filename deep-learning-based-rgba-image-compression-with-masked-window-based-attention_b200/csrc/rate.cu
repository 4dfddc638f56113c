// Rate and distortion terms of the codec's forward (the tail of models/AutoEncoderRGB_Journal.py:203-297): the masked squared
// error (reconstruct_error, :36-64), the bits of y under the Gaussian conditional and the bits of z under the factorised
// prior (:283-291; CompressAI's GaussianConditional._likelihood and EntropyBottleneck._likelihood / _logits_cumulative,
// filters (3, 3, 3, 3)).  The reference (and this package with autograd recording) evaluates them as ~90 small elementwise /
// reduction / batched-GEMM launches; in inference each term is ONE pass over its tensors here, accumulating into a double
// workspace that rate_finalize turns into [mse, y bpp, z bpp, total bpp].  The arithmetic per element follows the torch
// expressions operation by operation in fp32 (this file is compiled without FMA contraction); only the order of the final
// summation differs (double accumulation of per-block fp32 partial sums).
#include "common.cuh"
#include "status.cuh"

namespace b200 {
namespace {

constexpr int kRateThreads = 256;

__device__ __forceinline__ void block_add(double v, double* dst) {
    // warp shuffle tree, one atomic per warp (double atomics: the result is independent of the arrival order to ~1e-16)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(dst, v);
}

// clamp(-log(lik + 1e-10) / ln 2, 0, 50)      (_bits of the codec: torch.clamp(-1.0 * torch.log(lik + 1e-10) / math.log(2.0), 0, 50))
__device__ __forceinline__ float bits_of(float lik) {
    const float b = (-1.0f * logf(lik + 1e-10f)) / 0.6931471805599453f;
    return fminf(fmaxf(b, 0.f), 50.f);
}

// squared error over the pixels whose alpha is > 0: acc[2 b] += sum_c (x m - y m)^2, acc[2 b + 1] += C * #(alpha > 0)
__global__ void __launch_bounds__(kRateThreads)
masked_sq_error_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ mask, int C,
                       int64_t hw, double* __restrict__ acc) {
    const int b = blockIdx.y;
    const float* xb = x + int64_t(b) * C * hw;
    const float* yb = y + int64_t(b) * C * hw;
    const float* mb = mask + int64_t(b) * hw;
    float se = 0.f, cnt = 0.f;
    for (int64_t i = blockIdx.x * int64_t(kRateThreads) + threadIdx.x; i < hw; i += int64_t(gridDim.x) * kRateThreads) {
        const float m = mb[i] > 0.f ? 1.f : 0.f;
        cnt += m;
        for (int c = 0; c < C; ++c) {
            const float d = xb[c * hw + i] * m - yb[c * hw + i] * m;
            se += d * d;
        }
    }
    block_add(double(se), acc + 2 * b);
    block_add(double(cnt) * C, acc + 2 * b + 1);
}

// Gaussian conditional (scale_bound 0.11, likelihood_bound 1e-9):  v = |y - mu|, s = max(scale, 0.11),
// lik = max(0.5 erfc(-c (0.5 - v) / s) - 0.5 erfc(-c (-0.5 - v) / s), 1e-9),  c = 2^-1/2
__global__ void __launch_bounds__(kRateThreads)
gaussian_bits_kernel(const float* __restrict__ y, const float* __restrict__ scales, const float* __restrict__ means, int64_t n,
                     double* __restrict__ acc) {
    const float c = 0.70710678118654752440f;
    float sum = 0.f;
    for (int64_t i = blockIdx.x * int64_t(kRateThreads) + threadIdx.x; i < n; i += int64_t(gridDim.x) * kRateThreads) {
        const float v = fabsf(y[i] - means[i]);
        const float s = fmaxf(scales[i], 0.11f);
        const float upper = 0.5f * erfcf((-c * (0.5f - v)) / s);
        const float lower = 0.5f * erfcf((-c * (-0.5f - v)) / s);
        sum += bits_of(fmaxf(upper - lower, 1e-9f));
    }
    block_add(double(sum), acc);
}

// Factorised prior with filters (3, 3, 3, 3): the cumulative logits are a 1 -> 3 -> 3 -> 3 -> 3 -> 1 network per channel,
//   v <- softplus(M_i) v + b_i ;  v <- v + tanh(f_i) tanh(v)   (all but the last layer)
// lik = max(|sigmoid(sg up) - sigmoid(sg lo)|, 1e-9), lo / up = logits(z -/+ 0.5), sg = -sign(lo + up).
struct FactorisedParams {
    const float* m[5];      // _matrix0 (C, 3, 1), _matrix1..3 (C, 3, 3), _matrix4 (C, 1, 3)
    const float* b[5];      // _bias0..3 (C, 3, 1), _bias4 (C, 1, 1)
    const float* f[4];      // _factor0..3 (C, 3, 1)
};

__device__ __forceinline__ float softplus_(float v) { return v > 20.f ? v : log1pf(expf(v)); }
__device__ __forceinline__ float sigmoid_(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ float logits_cumulative(float v, const float* __restrict__ p) {
    // p: this channel's parameters in shared memory: M0[3] b0[3] f0[3] | (M[9] b[3] f[3]) x 3 | M4[3] b4[1]
    float h[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float t = p[r] * v + p[3 + r];
        h[r] = t + p[6 + r] * tanhf(t);
    }
    p += 9;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        float g[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float t = (p[3 * r] * h[0] + p[3 * r + 1] * h[1] + p[3 * r + 2] * h[2]) + p[9 + r];
            g[r] = t + p[12 + r] * tanhf(t);
        }
        h[0] = g[0]; h[1] = g[1]; h[2] = g[2];
        p += 15;
    }
    return (p[0] * h[0] + p[1] * h[1] + p[2] * h[2]) + p[3];
}

__global__ void __launch_bounds__(kRateThreads)
factorised_bits_kernel(const float* __restrict__ z, const FactorisedParams P, int B, int C, int64_t hw, double* __restrict__ acc) {
    __shared__ float sp[9 + 3 * 15 + 4];
    const int ch = blockIdx.x;
    if (threadIdx.x < 58) {
        const int t = threadIdx.x;
        float v;
        if (t < 9) {                                   // layer 0: M (3), b (3), f (3)
            const int k = t / 3, r = t % 3;
            v = k == 0 ? softplus_(P.m[0][ch * 3 + r]) : k == 1 ? P.b[0][ch * 3 + r] : tanhf(P.f[0][ch * 3 + r]);
        } else if (t < 54) {                           // layers 1..3: M (9), b (3), f (3)
            const int l = (t - 9) / 15 + 1, o = (t - 9) % 15;
            v = o < 9 ? softplus_(P.m[l][ch * 9 + o]) : o < 12 ? P.b[l][ch * 3 + o - 9] : tanhf(P.f[l][ch * 3 + o - 12]);
        } else {                                       // layer 4: M (3), b (1)
            const int o = t - 54;
            v = o < 3 ? softplus_(P.m[4][ch * 3 + o]) : P.b[4][ch];
        }
        sp[t] = v;
    }
    __syncthreads();
    float sum = 0.f;
    const int64_t n = int64_t(B) * hw;
    for (int64_t i = blockIdx.y * int64_t(kRateThreads) + threadIdx.x; i < n; i += int64_t(gridDim.y) * kRateThreads) {
        const int64_t b = i / hw, px = i - b * hw;
        const float v = z[(b * C + ch) * hw + px];
        const float lo = logits_cumulative(v - 0.5f, sp), up = logits_cumulative(v + 0.5f, sp);
        const float t = lo + up;
        const float sg = t > 0.f ? -1.f : t < 0.f ? 1.f : 0.f;       // -sign(lo + up)
        sum += bits_of(fmaxf(fabsf(sigmoid_(sg * up) - sigmoid_(sg * lo)), 1e-9f));
    }
    block_add(double(sum), acc);
}

// acc: [2 b] squared error, [2 b + 1] masked element count of image b; [2 B] y bits; [2 B + 1] z bits
__global__ void rate_finalize_kernel(const double* __restrict__ acc, int B, double npix, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float mse = 0.f;
    for (int b = 0; b < B; ++b) mse += float(acc[2 * b]) / fmaxf(float(acc[2 * b + 1]), 1.f);
    const float ybpp = float(acc[2 * B] / npix), zbpp = float(acc[2 * B + 1] / npix);
    out[0] = mse / float(B);
    out[1] = ybpp;
    out[2] = zbpp;
    out[3] = ybpp + zbpp;
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int64_t rate_workspace_bytes(int B) { return B < 0 ? MWA_ERR_INVALID : int64_t(2 * B + 2) * 8; }

int rate_forward(const float* input, const float* x_hat, const float* mask, int B, int C, int H, int W, const float* y,
                 const float* scales, const float* means, int64_t n_y, const float* z_hat, const float* const* eb_params,
                 int Cz, int64_t hw_z, void* workspace, int64_t workspace_bytes, float* out4, void* stream) {
    if (B < 0 || C <= 0 || H < 0 || W < 0 || n_y < 0 || Cz < 0 || hw_z < 0) return MWA_ERR_INVALID;
    if (!out4 || !workspace || !eb_params) return MWA_ERR_INVALID;
    if (workspace_bytes < rate_workspace_bytes(B)) return MWA_ERR_WORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 7u) != 0) return MWA_ERR_ALIGNMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* acc = static_cast<double*>(workspace);
    MWA_TRY_CUDA(cudaMemsetAsync(acc, 0, size_t(2 * B + 2) * 8, st), "rate_forward(memset)");
    const int64_t hw = int64_t(H) * W;
    if (B > 0 && hw > 0) {
        if (!input || !x_hat || !mask) return MWA_ERR_INVALID;
        const int per_img = int((hw + kRateThreads * 4 - 1) / (kRateThreads * 4));
        const int bx = per_img < 1 ? 1 : per_img > 4 * kNumSMs ? 4 * kNumSMs : per_img;
        masked_sq_error_kernel<<<dim3(bx, B), kRateThreads, 0, st>>>(input, x_hat, mask, C, hw, acc);
        int rc = check_launch("rate_forward(masked squared error)");
        if (rc != MWA_OK) return rc;
    }
    if (n_y > 0) {
        if (!y || !scales || !means) return MWA_ERR_INVALID;
        const int64_t want = (n_y + kRateThreads * 4 - 1) / (kRateThreads * 4);
        const int bx = want > 8 * kNumSMs ? 8 * kNumSMs : int(want);
        gaussian_bits_kernel<<<bx, kRateThreads, 0, st>>>(y, scales, means, n_y, acc + 2 * B);
        int rc = check_launch("rate_forward(gaussian bits)");
        if (rc != MWA_OK) return rc;
    }
    if (Cz > 0 && hw_z > 0 && B > 0) {
        if (!z_hat) return MWA_ERR_INVALID;
        FactorisedParams P;
        for (int i = 0; i < 5; ++i) { P.m[i] = eb_params[i]; P.b[i] = eb_params[5 + i]; }
        for (int i = 0; i < 4; ++i) P.f[i] = eb_params[10 + i];
        for (int i = 0; i < 14; ++i)
            if (!eb_params[i]) return MWA_ERR_INVALID;
        const int64_t per_ch = int64_t(B) * hw_z;
        int by = int((per_ch + kRateThreads * 2 - 1) / (kRateThreads * 2));
        by = by < 1 ? 1 : by > 64 ? 64 : by;
        factorised_bits_kernel<<<dim3(Cz, by), kRateThreads, 0, st>>>(z_hat, P, B, Cz, hw_z, acc + 2 * B + 1);
        int rc = check_launch("rate_forward(factorised bits)");
        if (rc != MWA_OK) return rc;
    }
    rate_finalize_kernel<<<1, 32, 0, st>>>(acc, B, double(B) * double(hw), out4);
    return check_launch("rate_forward(finalize)");
}

}  // extern "C"
