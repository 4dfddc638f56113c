// Masked MS-SSIM on the device (the evaluation metric of the RGBA codec).
// Reference semantics: metrics/masked_ms_ssim_torch.py:27-55 (separable valid-mode Gaussian, 11 taps, sigma 1.5, rows then
// columns), :58-121 (_ssim: SSIM / CS maps averaged over the positions where the nearest-resized mask is non-zero),
// :181-265 (ms_ssim: five levels; at each level the mask is binarised and multiplied into both images, then images and
// mask are 2x2 average-pooled with padding = size % 2).  The reference runs ~40 torch kernels per level (five grouped
// convolution pairs, elementwise maps, boolean reductions, three poolings); here a level is TWO launches:
//   ms_ssim_level : one CTA per 16 x 32 tile of the valid region of one (image, channel) plane: the masked X, Y tile with
//                   its 10-pixel apron goes to shared memory once, the five moments are blurred along H into shared
//                   memory, then along W in registers, maps and masked sums follow in the same thread; one atomicAdd pair
//                   per CTA.  HBM traffic = the two planes + the mask, once.
//   ms_ssim_pool  : (X * m, Y * m, m) -> 2x2 average pools of all three in one pass.
// fp32 like the reference; the sums are accumulated in a different order, so values agree to ~1e-6, not bit for bit.
#include "common.cuh"
#include "status.cuh"

namespace b200 {
namespace {

constexpr int kWin = 11, kApron = kWin - 1;
constexpr int kTH = 16, kTW = 32;                 // output tile (valid-region positions)
constexpr int kIH = kTH + kApron, kIW = kTW + kApron;

struct Gauss {
    float w[kWin];
};

__global__ void __launch_bounds__(256)
ms_ssim_level_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ mask, int C, int H,
                     int W, float C1, float C2, float scale_h, float scale_w, Gauss g, float* __restrict__ sums,
                     float* __restrict__ counts) {
    __shared__ float sx[kIH][kIW + 1], sy[kIH][kIW + 1];
    __shared__ float v[5][kTH][kIW + 1];           // moments after the blur along H: x, y, xx, yy, xy
    __shared__ float red[3][8];
    const int oh = H - kApron, ow = W - kApron;
    const int tiles_x = (ow + kTW - 1) / kTW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int c = blockIdx.y, b = blockIdx.z;
    const int y0 = ty * kTH, x0 = tx * kTW;
    const float* xp = X + (int64_t(b) * C + c) * H * W;
    const float* yp = Y + (int64_t(b) * C + c) * H * W;
    const float* mp = mask + int64_t(b) * H * W;
    for (int i = threadIdx.x; i < kIH * kIW; i += 256) {
        const int r = i / kIW, q = i % kIW, yy = y0 + r, xx = x0 + q;
        float a = 0.f, d = 0.f;
        if (yy < H && xx < W) {
            const float m = mp[int64_t(yy) * W + xx] > 0.f ? 1.f : 0.f;        // (mask > 0) multiplied into both images (:236-238)
            a = xp[int64_t(yy) * W + xx] * m;
            d = yp[int64_t(yy) * W + xx] * m;
        }
        sx[r][q] = a;
        sy[r][q] = d;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kTH * kIW; i += 256) {
        const int r = i / kIW, q = i % kIW;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const float a = sx[r + k][q], d = sy[r + k][q], w = g.w[k];
            a0 = fmaf(w, a, a0);
            a1 = fmaf(w, d, a1);
            a2 = fmaf(w, a * a, a2);
            a3 = fmaf(w, d * d, a3);
            a4 = fmaf(w, a * d, a4);
        }
        v[0][r][q] = a0; v[1][r][q] = a1; v[2][r][q] = a2; v[3][r][q] = a3; v[4][r][q] = a4;
    }
    __syncthreads();
    float s_ssim = 0.f, s_cs = 0.f, s_n = 0.f;
    for (int i = threadIdx.x; i < kTH * kTW; i += 256) {
        const int r = i / kTW, q = i % kTW, yy = y0 + r, xx = x0 + q;
        if (yy < oh && xx < ow) {
            float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < kWin; ++k) {
                const float w = g.w[k];
#pragma unroll
                for (int f = 0; f < 5; ++f) m[f] = fmaf(w, v[f][r][q + k], m[f]);
            }
            const float mu1 = m[0], mu2 = m[1];
            const float s1 = m[2] - mu1 * mu1, s2 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
            const float cs = (2.f * s12 + C2) / (s1 + s2 + C2);
            const float ss = ((2.f * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs;
            // nearest-resized mask of the valid region (:104-108): source index floor(dst * in / out)
            const int my = min(int(floorf(float(yy) * scale_h)), H - 1), mx = min(int(floorf(float(xx) * scale_w)), W - 1);
            if (mp[int64_t(my) * W + mx] > 0.f) {
                s_ssim += ss;
                s_cs += cs;
                s_n += 1.f;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_ssim += __shfl_xor_sync(0xffffffffu, s_ssim, o);
        s_cs += __shfl_xor_sync(0xffffffffu, s_cs, o);
        s_n += __shfl_xor_sync(0xffffffffu, s_n, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s_ssim;
        red[1][threadIdx.x >> 5] = s_cs;
        red[2][threadIdx.x >> 5] = s_n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, d = 0.f, n = 0.f;
        for (int i = 0; i < 8; ++i) {
            a += red[0][i];
            d += red[1][i];
            n += red[2][i];
        }
        atomicAdd(sums + (int64_t(b) * C + c) * 2, a);
        atomicAdd(sums + (int64_t(b) * C + c) * 2 + 1, d);
        if (c == 0) atomicAdd(counts + b, n);
    }
}

// out sizes: Hp = (H + 2 ph - 2) / 2 + 1 with ph = H % 2 (F.avg_pool2d(kernel 2, padding = size % 2), count_include_pad)
__global__ void __launch_bounds__(256)
ms_ssim_pool_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ mask, int C, int H,
                    int W, int Hp, int Wp, float* __restrict__ Xo, float* __restrict__ Yo, float* __restrict__ Mo, int64_t n) {
    const int ph = H % 2, pw = W % 2;
    for (int64_t i = blockIdx.x * int64_t(256) + threadIdx.x; i < n; i += int64_t(gridDim.x) * 256) {
        const int xo = int(i % Wp), yo = int((i / Wp) % Hp), b = int(i / (int64_t(Wp) * Hp));
        float mv[4], ms = 0.f;
        int64_t off[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int yy = 2 * yo - ph + (t >> 1), xx = 2 * xo - pw + (t & 1);
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
            off[t] = in ? int64_t(yy) * W + xx : -1;
            mv[t] = in && mask[int64_t(b) * H * W + off[t]] > 0.f ? 1.f : 0.f;
            ms += mv[t];
        }
        Mo[i] = ms * 0.25f;
        for (int c = 0; c < C; ++c) {
            const float* xp = X + (int64_t(b) * C + c) * H * W;
            const float* yp = Y + (int64_t(b) * C + c) * H * W;
            float a = 0.f, d = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (off[t] >= 0) {
                    a += xp[off[t]] * mv[t];
                    d += yp[off[t]] * mv[t];
                }
            const int64_t o = ((int64_t(b) * C + c) * Hp + yo) * Wp + xo;
            Xo[o] = a * 0.25f;
            Yo[o] = d * 0.25f;
        }
    }
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int ms_ssim_level_forward(const float* X, const float* Y, const float* mask, int B, int C, int H, int W, float data_range,
                          float* sums, float* counts, void* stream) {
    if (!X || !Y || !mask || !sums || !counts || B < 0 || C <= 0) return MWA_ERR_INVALID;
    if (H < kWin || W < kWin) return MWA_ERR_UNSUPPORTED;      // the reference leaves a short dimension unblurred; not needed above 160 px
    if (B == 0) return MWA_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MWA_TRY_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * B * C, st), "ms_ssim_level(memset)");
    MWA_TRY_CUDA(cudaMemsetAsync(counts, 0, sizeof(float) * B, st), "ms_ssim_level(memset)");
    Gauss g;
    {
        // metrics/masked_ms_ssim_torch.py:13-24 (_fspecial_gauss_1d) in fp32
        float s = 0.f;
        for (int i = 0; i < kWin; ++i) {
            const float c = float(i - kWin / 2);
            g.w[i] = expf(-(c * c) / (2.f * 1.5f * 1.5f));
            s += g.w[i];
        }
        for (int i = 0; i < kWin; ++i) g.w[i] /= s;
    }
    const int oh = H - kApron, ow = W - kApron;
    const float C1 = (0.01f * data_range) * (0.01f * data_range), C2 = (0.03f * data_range) * (0.03f * data_range);
    dim3 grid(unsigned(((ow + kTW - 1) / kTW) * ((oh + kTH - 1) / kTH)), unsigned(C), unsigned(B));
    ms_ssim_level_kernel<<<grid, 256, 0, st>>>(X, Y, mask, C, H, W, C1, C2, float(double(H) / oh), float(double(W) / ow), g, sums,
                                              counts);
    return check_launch("ms_ssim_level_forward");
}

int ms_ssim_pool_forward(const float* X, const float* Y, const float* mask, int B, int C, int H, int W, float* Xo, float* Yo,
                         float* Mo, void* stream) {
    if (!X || !Y || !mask || !Xo || !Yo || !Mo || B < 0 || C <= 0 || H <= 0 || W <= 0) return MWA_ERR_INVALID;
    if (B == 0) return MWA_OK;
    const int Hp = (H + 2 * (H % 2) - 2) / 2 + 1, Wp = (W + 2 * (W % 2) - 2) / 2 + 1;
    const int64_t n = int64_t(B) * Hp * Wp;
    int64_t blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    ms_ssim_pool_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(X, Y, mask, C, H, W, Hp, Wp, Xo, Yo, Mo, n);
    return check_launch("ms_ssim_pool_forward");
}

}  // extern "C"
