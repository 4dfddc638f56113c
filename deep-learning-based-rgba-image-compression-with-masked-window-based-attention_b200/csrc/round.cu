// Latent rounding kernels (HBM-bound elementwise; 128-bit vectorised, grid sized in multiples of the SM count).
// Reference semantics: models/AutoEncoderRGB_Journal.py:31-32 (ste_round), :227-229, :257, :262-264, :212-214.
// Compiled WITHOUT fast-math and with -fmad=false so that sub / rint / add / div round exactly like
// torch's separate fp32 kernels (bit-exact rounded latents).
#include "common.cuh"
#include "status.cuh"

namespace b200 {
namespace {

constexpr int kThreads = 256;

enum Op { kRound = 0, kQuantSame = 1, kQuantChan = 2, kLrp = 3 };

template <int OP>
__device__ __forceinline__ float apply(float a, float b) {
    // ste_round(v) = (round(v) - v) + v evaluated literally (models/AutoEncoderRGB_Journal.py:31-32): the value is
    // rint(v) except that a zero result is always +0.0 (e.g. v = -0.3: (-0.0 + 0.3) - 0.3 = +0.0), like the reference.
    if (OP == kRound) return __fadd_rn(__fsub_rn(rintf(a), a), a);
    if (OP == kQuantSame || OP == kQuantChan) {
        const float d = __fsub_rn(a, b);
        return __fadd_rn(__fadd_rn(__fsub_rn(rintf(d), d), d), b);
    }
    return __fadd_rn(a, __fmul_rn(0.5f, tanhf(b)));
}

// VEC == 4: every row start and row_len are multiples of 4 floats and pointers 16-byte aligned.
template <int OP, int VEC>
__global__ void __launch_bounds__(kThreads)
rowwise_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t rows,
               int64_t row_len, int64_t a_stride, int64_t b_stride, int64_t o_stride, int mu_channels, int64_t hw) {
    const int64_t per_row = row_len / VEC;
    const int64_t total = rows * per_row;
    for (int64_t i = blockIdx.x * int64_t(kThreads) + threadIdx.x; i < total; i += int64_t(gridDim.x) * kThreads) {
        const int64_t r = i / per_row, c = (i - r * per_row) * VEC;
        if (VEC == 4) {
            const float4 va = __ldcs(reinterpret_cast<const float4*>(a + r * a_stride + c));
            float4 vb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (OP == kQuantSame || OP == kLrp) vb = __ldcs(reinterpret_cast<const float4*>(b + r * b_stride + c));
            if (OP == kQuantChan) {
                vb.x = __ldg(b + ((c + 0) / hw) % mu_channels);
                vb.y = __ldg(b + ((c + 1) / hw) % mu_channels);
                vb.z = __ldg(b + ((c + 2) / hw) % mu_channels);
                vb.w = __ldg(b + ((c + 3) / hw) % mu_channels);
            }
            float4 vo;
            vo.x = apply<OP>(va.x, vb.x);
            vo.y = apply<OP>(va.y, vb.y);
            vo.z = apply<OP>(va.z, vb.z);
            vo.w = apply<OP>(va.w, vb.w);
            __stcs(reinterpret_cast<float4*>(out + r * o_stride + c), vo);
        } else {
            const float va = a[r * a_stride + c];
            float vb = 0.f;
            if (OP == kQuantSame || OP == kLrp) vb = b[r * b_stride + c];
            if (OP == kQuantChan) vb = __ldg(b + (c / hw) % mu_channels);
            out[r * o_stride + c] = apply<OP>(va, vb);
        }
    }
}

__device__ __forceinline__ float to_level(float v, float levels) { return __fdiv_rn(rintf(__fmul_rn(v, levels)), levels); }
// VEC = 4: 128-bit streaming loads / stores (n % 4 == 0, 16-byte aligned); VEC = 1: any n
template <int VEC>
__global__ void __launch_bounds__(kThreads)
levels_kernel(const float* __restrict__ m, float* __restrict__ out, int64_t n, float levels) {
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    if constexpr (VEC == 4) {
        const float4* m4 = reinterpret_cast<const float4*>(m);
        float4* o4 = reinterpret_cast<float4*>(out);
        for (int64_t i = blockIdx.x * int64_t(kThreads) + threadIdx.x; i < n / 4; i += stride) {
            float4 v = __ldcs(m4 + i);
            v.x = to_level(v.x, levels);
            v.y = to_level(v.y, levels);
            v.z = to_level(v.z, levels);
            v.w = to_level(v.w, levels);
            __stcs(o4 + i, v);
        }
    } else {
        for (int64_t i = blockIdx.x * int64_t(kThreads) + threadIdx.x; i < n; i += stride) out[i] = to_level(m[i], levels);
    }
}

inline int grid_for(int64_t work_items) {
    int64_t blocks = (work_items + kThreads - 1) / kThreads;
    const int64_t cap = int64_t(kNumSMs) * 8;          // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

template <int OP>
int launch_rowwise(const float* a, const float* b, float* out, int64_t rows, int64_t row_len, int64_t a_stride,
                   int64_t b_stride, int64_t o_stride, int mu_channels, int64_t hw, cudaStream_t st,
                   const char* where) {
    if (rows < 0 || row_len < 0) return MWA_ERR_INVALID;
    if (rows == 0 || row_len == 0) return MWA_OK;            // empty tensors carry null data pointers
    if (!a || !out) return MWA_ERR_INVALID;
    if ((OP != kRound) && !b) return MWA_ERR_INVALID;
    if (OP == kQuantChan && (mu_channels <= 0 || hw <= 0)) return MWA_ERR_INVALID;
    bool vec = aligned16(a) && aligned16(out) && row_len % 4 == 0 && a_stride % 4 == 0 && o_stride % 4 == 0;
    if (OP == kQuantSame || OP == kLrp) vec = vec && aligned16(b) && b_stride % 4 == 0;
    if (vec) {
        rowwise_kernel<OP, 4><<<grid_for(rows * row_len / 4), kThreads, 0, st>>>(a, b, out, rows, row_len, a_stride,
                                                                                b_stride, o_stride, mu_channels, hw);
    } else {
        rowwise_kernel<OP, 1><<<grid_for(rows * row_len), kThreads, 0, st>>>(a, b, out, rows, row_len, a_stride,
                                                                            b_stride, o_stride, mu_channels, hw);
    }
    return check_launch(where);
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int round_ste_forward(const float* x, float* out, int64_t rows, int64_t row_len, int64_t x_row_stride,
                      int64_t out_row_stride, void* stream) {
    return launch_rowwise<kRound>(x, nullptr, out, rows, row_len, x_row_stride, 0, out_row_stride, 0, 1,
                                  static_cast<cudaStream_t>(stream), "round_ste_forward");
}

int quantize_offset_forward(const float* x, const float* mu, float* out, int64_t rows, int64_t row_len,
                            int64_t x_row_stride, int64_t mu_row_stride, int64_t out_row_stride, int mu_channels,
                            int64_t hw, void* stream) {
    if (mu_channels > 0)
        return launch_rowwise<kQuantChan>(x, mu, out, rows, row_len, x_row_stride, 0, out_row_stride, mu_channels, hw,
                                          static_cast<cudaStream_t>(stream), "quantize_offset_forward");
    return launch_rowwise<kQuantSame>(x, mu, out, rows, row_len, x_row_stride, mu_row_stride, out_row_stride, 0, 1,
                                      static_cast<cudaStream_t>(stream), "quantize_offset_forward");
}

int lrp_add_forward(const float* y_hat, const float* lrp, float* out, int64_t rows, int64_t row_len,
                    int64_t y_row_stride, int64_t lrp_row_stride, int64_t out_row_stride, void* stream) {
    return launch_rowwise<kLrp>(y_hat, lrp, out, rows, row_len, y_row_stride, lrp_row_stride, out_row_stride, 0, 1,
                                static_cast<cudaStream_t>(stream), "lrp_add_forward");
}

int quantize_levels_forward(const float* m, float* out, int64_t n, float levels, void* stream) {
    if (n < 0 || !(levels > 0.f)) return MWA_ERR_INVALID;
    if (n == 0) return MWA_OK;
    if (!m || !out) return MWA_ERR_INVALID;
    if (n % 4 == 0 && aligned16(m) && aligned16(out))
        levels_kernel<4><<<grid_for(n / 4), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(m, out, n, levels);
    else
        levels_kernel<1><<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(m, out, n, levels);
    return check_launch("quantize_levels_forward");
}

}  // extern "C"
