// GDN / IGDN: parameter preparation, the general-shape SIMT forward kernel and the backward kernels.
// Reference semantics: layers/GDN.py:64-94 (forward), :9-23 (LowerBound), :46-62 (constants).
// The tcgen05 forward for C == 192 lives in gdn_tc.cu.
#include "common.cuh"
#include "params.cuh"
#include "status.cuh"

namespace b200 {

int gdn_forward_planes_tc(const float* x, void* out_hi, void* out_lo, int ps, int cstride, const void* params, int64_t n_img,
                          int C, int H, int W, int inverse, cudaStream_t st);
int gdn_forward_tc(const float* x, float* y, const void* params, int64_t n_img, int C, int64_t hw, int inverse,
                   int channels_last, cudaStream_t st);   // gdn_tc.cu
bool gdn_tc_supported(int C, int64_t hw, int channels_last);

namespace {

// ------------------------------------------------------------------ prepare
// beta = max(beta_p, bb)^2 - ped ; gamma = max(gamma_p, gb)^2 - ped   (layers/GDN.py:74-80), in fp32 like torch.
__global__ void gdn_prepare_kernel(const float* __restrict__ beta_p, const float* __restrict__ gamma_p, int C,
                                   float beta_bound, float gamma_bound, float pedestal, uint8_t* __restrict__ blk) {
    const GdnParamLayout L(C);
    float* beta = reinterpret_cast<float*>(blk + L.beta);
    float* gamma = reinterpret_cast<float*>(blk + L.gamma);
    float* gammaT = reinterpret_cast<float*>(blk + L.gammaT);
    uint8_t* mb = blk + L.mask_beta;
    uint8_t* mg = blk + L.mask_gamma;
    uint8_t* hi = blk + L.img_hi;
    uint8_t* lo = blk + L.img_lo;
    const int64_t rows = align_up(C, 8);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int nthreads = gridDim.x * blockDim.x;
    for (int i = tid; i < C; i += nthreads) {
        const float p = beta_p[i];
        const float b = fmaxf(p, beta_bound);
        beta[i] = __fsub_rn(__fmul_rn(b, b), pedestal);
        mb[i] = p >= beta_bound;
    }
    for (int e = tid; e < C * C; e += nthreads) {
        const int i = e / C, j = e % C;        // i = output channel, j = input channel
        const float p = gamma_p[e];
        const float g0 = fmaxf(p, gamma_bound);
        const float g = __fsub_rn(__fmul_rn(g0, g0), pedestal);
        gamma[e] = g;
        gammaT[j * C + i] = g;
        mg[e] = p >= gamma_bound;
        // bf16 hi/lo split for the tensor-core path: B operand rows = output channel i, K = input channel j
        const float gh = bf16_round(g);
        const float gl = bf16_round(g - gh);
        const int64_t off = int64_t(j / 64) * rows * 128 + sw128_offset(i, j % 64);
        *reinterpret_cast<uint16_t*>(hi + off) = static_cast<uint16_t>(__float_as_uint(gh) >> 16);
        *reinterpret_cast<uint16_t*>(lo + off) = static_cast<uint16_t>(__float_as_uint(gl) >> 16);
    }
}

__global__ void zero_bytes_kernel(uint4* p, int64_t n16) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n16; i += int64_t(gridDim.x) * blockDim.x)
        p[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------ SIMT forward (any C <= 512)
// One CTA = one tile of P pixels of one image.  x tile staged in smem as s[c][p] (row stride P+1);
// thread (p = tid % P, g = tid / P) accumulates CPT output channels in registers:
//   n_i = beta_i + sum_j gammaT[j][i] * x_j^2      (gammaT rows are contiguous in i -> broadcast float4 loads)
constexpr int kTileP = 64;
constexpr int kSimtThreads = 256;

template <int CPT>
__global__ void __launch_bounds__(kSimtThreads)
gdn_simt_kernel(const float* __restrict__ x, float* __restrict__ y, const uint8_t* __restrict__ blk, int C,
                int64_t hw, int64_t tiles_per_img, int inverse, int channels_last) {
    extern __shared__ float s[];                     // [C][kTileP + 1]
    constexpr int LD = kTileP + 1;
    const GdnParamLayout L(C);
    const float* __restrict__ beta = reinterpret_cast<const float*>(blk + L.beta);
    const float* __restrict__ gT = reinterpret_cast<const float*>(blk + L.gammaT);
    const int tid = threadIdx.x;
    const int64_t img = blockIdx.x / tiles_per_img;
    const int64_t p0 = (blockIdx.x % tiles_per_img) * kTileP;
    const int np = (hw - p0 < kTileP) ? static_cast<int>(hw - p0) : kTileP;
    const float* xi = x + img * C * hw;
    float* yi = y + img * C * hw;

    if (!channels_last) {
        for (int e = tid; e < C * kTileP; e += kSimtThreads) {
            const int c = e / kTileP, p = e % kTileP;
            s[c * LD + p] = p < np ? __ldg(xi + int64_t(c) * hw + p0 + p) : 0.f;
        }
    } else {
        for (int e = tid; e < C * kTileP; e += kSimtThreads) {
            const int p = e / C, c = e % C;
            s[c * LD + p] = p < np ? __ldg(xi + (p0 + p) * C + c) : 0.f;
        }
    }
    __syncthreads();

    const int p = tid % kTileP, g = tid / kTileP;    // 4 channel groups
    const int c0 = g * CPT;
    float acc[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) acc[i] = (c0 + i < C) ? beta[c0 + i] : 1.f;
    if (c0 < C) {
        for (int j = 0; j < C; ++j) {
            const float v = s[j * LD + p];
            const float v2 = v * v;
            const float* grow = gT + int64_t(j) * C + c0;
#pragma unroll
            for (int i = 0; i < CPT; ++i)
                if (c0 + i < C) acc[i] = fmaf(__ldg(grow + i), v2, acc[i]);
        }
    }
    float res[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const float xv = (c0 + i < C) ? s[(c0 + i) * LD + p] : 0.f;
        const float r = sqrtf(acc[i]);
        res[i] = inverse ? xv * r : xv / r;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CPT; ++i)
        if (c0 + i < C) s[(c0 + i) * LD + p] = res[i];
    __syncthreads();

    if (!channels_last) {
        for (int e = tid; e < C * kTileP; e += kSimtThreads) {
            const int c = e / kTileP, pp = e % kTileP;
            if (pp < np) yi[int64_t(c) * hw + p0 + pp] = s[c * LD + pp];
        }
    } else {
        for (int e = tid; e < C * kTileP; e += kSimtThreads) {
            const int pp = e / C, c = e % C;
            if (pp < np) yi[(p0 + pp) * C + c] = s[c * LD + pp];
        }
    }
}

template <int CPT>
int launch_gdn_simt(const float* x, float* y, const void* params, int64_t n_img, int C, int64_t hw, int inverse,
                    int channels_last, cudaStream_t st) {
    const int64_t tiles = (hw + kTileP - 1) / kTileP;
    const int64_t grid = n_img * tiles;
    if (grid > 0x7fffffffll) return MWA_ERR_UNSUPPORTED;
    const int smem = C * (kTileP + 1) * 4;
    MWA_TRY_CUDA(cudaFuncSetAttribute(gdn_simt_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                 "gdn_forward(simt attr)");
    gdn_simt_kernel<CPT><<<static_cast<unsigned>(grid), kSimtThreads, smem, st>>>(
        x, y, static_cast<const uint8_t*>(params), C, hw, tiles, inverse, channels_last);
    return check_launch("gdn_forward(simt)");
}


// ------------------------------------------------------------------ SIMT backward
// Per pixel:  n_i = beta_i + sum_j gamma_ij x_j^2
//   GDN : y = x n^-1/2 ; dn_i = -1/2 g_i x_i n_i^-3/2 ;  IGDN: y = x n^1/2 ; dn_i = 1/2 g_i x_i n_i^-1/2
//   dx_j = g_j n_j^(-/+1/2) + 2 x_j sum_i gamma_ij dn_i ;  dgamma_ij = sum_px dn_i x_j^2 ;  dbeta_i = sum_px dn_i
// Kernel A (per 64-pixel tile): dx, dn -> workspace (always [img][c][p]), dbeta partial sums (atomics).
// Kernel B: dgamma = dn * (x^2)^T  (64x64 output tiles, split over pixels, fp32 atomics).
// Kernel C: chain through gamma = max(gamma_p, bound)^2 - pedestal incl. LowerBound's pass-through rule.
template <int CPT>
__global__ void __launch_bounds__(kSimtThreads)
gdn_bwd_tile_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                    float* __restrict__ dn_ws, float* __restrict__ dbeta, const uint8_t* __restrict__ blk, int C,
                    int64_t hw, int64_t tiles_per_img, int inverse, int channels_last) {
    extern __shared__ float s[];                     // sx[C][LD], sg[C][LD]
    constexpr int LD = kTileP + 1;
    float* sx = s;
    float* sg = s + C * LD;
    const GdnParamLayout L(C);
    const float* __restrict__ beta = reinterpret_cast<const float*>(blk + L.beta);
    const float* __restrict__ gT = reinterpret_cast<const float*>(blk + L.gammaT);   // [j][i]
    const float* __restrict__ gm = reinterpret_cast<const float*>(blk + L.gamma);    // [i][j]
    const int tid = threadIdx.x;
    const int64_t img = blockIdx.x / tiles_per_img;
    const int64_t p0 = (blockIdx.x % tiles_per_img) * kTileP;
    const int np = (hw - p0 < kTileP) ? static_cast<int>(hw - p0) : kTileP;
    const float* xi = x + img * C * hw;
    const float* gi = gy + img * C * hw;
    float* oi = gx + img * C * hw;

    for (int e = tid; e < C * kTileP; e += kSimtThreads) {
        int c, p;
        int64_t off;
        if (!channels_last) { c = e / kTileP; p = e % kTileP; off = int64_t(c) * hw + p0 + p; }
        else                { p = e / C;      c = e % C;      off = (p0 + p) * C + c; }
        const bool ok = p < np;
        sx[c * LD + p] = ok ? __ldg(xi + off) : 0.f;
        sg[c * LD + p] = ok ? __ldg(gi + off) : 0.f;
    }
    __syncthreads();

    const int p = tid % kTileP, g = tid / kTileP;
    const int c0 = g * CPT;
    float acc[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) acc[i] = (c0 + i < C) ? beta[c0 + i] : 1.f;
    if (c0 < C) {
        for (int j = 0; j < C; ++j) {
            const float v = sx[j * LD + p];
            const float v2 = v * v;
            const float* grow = gT + int64_t(j) * C + c0;
#pragma unroll
            for (int i = 0; i < CPT; ++i)
                if (c0 + i < C) acc[i] = fmaf(__ldg(grow + i), v2, acc[i]);
        }
    }
    float term1[CPT], dn[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const bool ok = c0 + i < C;
        const float xv = ok ? sx[(c0 + i) * LD + p] : 0.f;
        const float gv = ok ? sg[(c0 + i) * LD + p] : 0.f;
        const float n = acc[i];
        const float r = sqrtf(n);
        if (inverse) { term1[i] = gv * r;  dn[i] = 0.5f * gv * xv / r; }
        else         { term1[i] = gv / r;  dn[i] = -0.5f * gv * xv / (n * r); }
        if (p >= np) dn[i] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CPT; ++i)
        if (c0 + i < C) sg[(c0 + i) * LD + p] = dn[i];          // sg now holds dn
    __syncthreads();

    // dn -> workspace ([img][c][p]) and dbeta partials
    float* dni = dn_ws + img * C * hw;
    for (int e = tid; e < C * kTileP; e += kSimtThreads) {
        const int c = e / kTileP, pp = e % kTileP;
        if (pp < np) dni[int64_t(c) * hw + p0 + pp] = sg[c * LD + pp];
    }
    for (int c = tid >> 5; c < C; c += kSimtThreads / 32) {
        float v = sg[c * LD + (tid & 31)] + sg[c * LD + 32 + (tid & 31)];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) atomicAdd(dbeta + c, v);
    }

    // t_j = sum_i gamma_ij dn_i  for this thread's channels j
    float t[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) t[i] = 0.f;
    if (c0 < C) {
        for (int i = 0; i < C; ++i) {
            const float d = sg[i * LD + p];
            const float* grow = gm + int64_t(i) * C + c0;
#pragma unroll
            for (int j = 0; j < CPT; ++j)
                if (c0 + j < C) t[j] = fmaf(__ldg(grow + j), d, t[j]);
        }
    }
    float res[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const float xv = (c0 + j < C) ? sx[(c0 + j) * LD + p] : 0.f;
        res[j] = fmaf(2.f * xv, t[j], term1[j]);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < CPT; ++j)
        if (c0 + j < C) sx[(c0 + j) * LD + p] = res[j];
    __syncthreads();
    for (int e = tid; e < C * kTileP; e += kSimtThreads) {
        int c, pp;
        int64_t off;
        if (!channels_last) { c = e / kTileP; pp = e % kTileP; off = int64_t(c) * hw + p0 + pp; }
        else                { pp = e / C;     c = e % C;       off = (p0 + pp) * C + c; }
        if (pp < np) oi[off] = sx[c * LD + pp];
    }
}

// dgamma[i][j] += sum over this block's pixel range of dn[i][p] * x[j][p]^2 ; 64x64 tile, 256 threads x (4x4)
constexpr int kGT = 64, kGK = 32;
__global__ void __launch_bounds__(256)
gdn_bwd_dgamma_kernel(const float* __restrict__ x, const float* __restrict__ dn_ws, float* __restrict__ dgamma, int C,
                      int64_t hw, int64_t n_img, int64_t px_per_block, int channels_last) {
    __shared__ float sd[kGK][kGT + 1];     // dn  [p][i]
    __shared__ float sq[kGK][kGT + 1];     // x^2 [p][j]
    const int tiles = (C + kGT - 1) / kGT;
    const int i0 = (blockIdx.x / tiles) * kGT, j0 = (blockIdx.x % tiles) * kGT;
    const int64_t total_px = n_img * hw;
    const int64_t pbeg = int64_t(blockIdx.y) * px_per_block;
    const int64_t pend = (pbeg + px_per_block < total_px) ? pbeg + px_per_block : total_px;
    const int tid = threadIdx.x, ti = tid / 16, tj = tid % 16;
    float acc[4][4] = {};
    for (int64_t pb = pbeg; pb < pend; pb += kGK) {
        for (int e = tid; e < kGK * kGT; e += 256) {
            const int cc = e / kGK, pp = e % kGK;              // pixel fastest: coalesced for [c][p] layouts
            const int64_t gp = pb + pp;
            float dv = 0.f, xv = 0.f;
            if (gp < pend) {
                const int64_t img = gp / hw, pix = gp % hw;
                if (i0 + cc < C) dv = __ldg(dn_ws + (img * C + i0 + cc) * hw + pix);
                if (j0 + cc < C)
                    xv = channels_last ? __ldg(x + (img * hw + pix) * C + j0 + cc)
                                       : __ldg(x + (img * C + j0 + cc) * hw + pix);
            }
            sd[pp][cc] = dv;
            sq[pp][cc] = xv * xv;
        }
        __syncthreads();
#pragma unroll 8
        for (int pp = 0; pp < kGK; ++pp) {
            float a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = sd[pp][ti * 4 + u]; b[u] = sq[pp][tj * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int i = i0 + ti * 4 + u, j = j0 + tj * 4 + v;
            if (i < C && j < C) atomicAdd(dgamma + int64_t(i) * C + j, acc[u][v]);
        }
}

// ------------------------------------------------------------------ GEMM-composed backward: elementwise stages
// The three C x C contractions of the backward (n = gamma x^2, t = gamma^T dn, dgamma = dn (x^2)^T) are plain GEMMs over
// the pixel dimension; the caller runs them with a library GEMM on the fp32 gamma / gamma^T of the parameter block and
// uses these kernels for everything in between (128-bit vectorised, grid-stride, one pass each).
__global__ void __launch_bounds__(256) gdn_bwd_square_kernel(const float4* __restrict__ x, float4* __restrict__ x2, int64_t n4) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
        const float4 v = __ldg(x + i);
        x2[i] = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
    }
}
// in: nb = gamma x^2 (without beta).  out: dn, and nb overwritten with term1 = g * n^(-/+ 1/2).  Same arithmetic as
// gdn_bwd_tile_kernel.  Channel of flat element i: (i / hw) % C (NCHW) or i % C (NHWC).
__global__ void __launch_bounds__(256)
gdn_bwd_dn_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ nb,
                  const float* __restrict__ beta, float* __restrict__ dn, int64_t n, int C, int64_t hw, int inverse,
                  int channels_last) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const int c = channels_last ? int(i % C) : int((i / hw) % C);
        const float nn = nb[i] + __ldg(beta + c), gv = __ldg(g + i), xv = __ldg(x + i);
        const float r = sqrtf(nn);
        float t1, d;
        if (inverse) { t1 = gv * r;  d = 0.5f * gv * xv / r; }
        else         { t1 = gv / r;  d = -0.5f * gv * xv / (nn * r); }
        nb[i] = t1;
        dn[i] = d;
    }
}
__global__ void __launch_bounds__(256)
gdn_bwd_dx_kernel(const float4* __restrict__ term1, const float4* __restrict__ x, const float4* __restrict__ t,
                  float4* __restrict__ dx, int64_t n4) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
        const float4 a = __ldg(term1 + i), b = __ldg(x + i), c = __ldg(t + i);
        dx[i] = make_float4(fmaf(2.f * b.x, c.x, a.x), fmaf(2.f * b.y, c.y, a.y), fmaf(2.f * b.z, c.z, a.z),
                            fmaf(2.f * b.w, c.w, a.w));
    }
}
// dbeta[c] = sum over images and pixels of dn.  NCHW: one warp per (image, channel) row; NHWC: block-strided columns.
__global__ void __launch_bounds__(256)
gdn_bwd_dbeta_kernel(const float* __restrict__ dn, float* __restrict__ dbeta, int64_t n_img, int C, int64_t hw,
                     int channels_last) {
    const int lane = threadIdx.x & 31;
    if (!channels_last) {
        const int64_t rows = n_img * C;
        for (int64_t row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += int64_t(gridDim.x) * 8) {
            const float* p = dn + row * hw;
            float a = 0.f;
            for (int64_t i = lane; i < hw; i += 32) a += __ldg(p + i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) atomicAdd(dbeta + row % C, a);
        }
    } else {
        const int64_t px = n_img * hw;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float a = 0.f;
            for (int64_t p = blockIdx.x; p < px; p += gridDim.x) a += __ldg(dn + p * C + c);
            atomicAdd(dbeta + c, a);
        }
    }
}

// LowerBound backward (layers/GDN.py:17-23): pass where param >= bound or the incoming gradient is negative
__global__ void gdn_bwd_finalize_kernel(const float* __restrict__ beta_p, const float* __restrict__ gamma_p,
                                        const float* __restrict__ dbeta_eff, const float* __restrict__ dgamma_eff,
                                        float* __restrict__ dbeta_p, float* __restrict__ dgamma_p, int C,
                                        float beta_bound, float gamma_bound) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < C; i += nth) {
        const float p = beta_p[i];
        const float g = dbeta_eff[i] * (2.f * fmaxf(p, beta_bound));
        dbeta_p[i] = (p >= beta_bound || g < 0.f) ? g : 0.f;
    }
    for (int e = tid; e < C * C; e += nth) {
        const float p = gamma_p[e];
        const float g = dgamma_eff[e] * (2.f * fmaxf(p, gamma_bound));
        dgamma_p[e] = (p >= gamma_bound || g < 0.f) ? g : 0.f;
    }
}

template <int CPT>
int launch_gdn_bwd_tile(const float* x, const float* gy, float* gx, float* dn_ws, float* dbeta, const void* params,
                        int64_t n_img, int C, int64_t hw, int inverse, int channels_last, cudaStream_t st) {
    const int64_t tiles = (hw + kTileP - 1) / kTileP;
    const int64_t grid = n_img * tiles;
    if (grid > 0x7fffffffll) return MWA_ERR_UNSUPPORTED;
    const int smem = 2 * C * (kTileP + 1) * 4;
    if (smem > 227 * 1024) return MWA_ERR_UNSUPPORTED;
    MWA_TRY_CUDA(cudaFuncSetAttribute(gdn_bwd_tile_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                 "gdn_backward(attr)");
    gdn_bwd_tile_kernel<CPT><<<static_cast<unsigned>(grid), kSimtThreads, smem, st>>>(
        x, gy, gx, dn_ws, dbeta, static_cast<const uint8_t*>(params), C, hw, tiles, inverse, channels_last);
    return check_launch("gdn_backward(tile)");
}

struct GdnBwdWorkspace {
    int64_t dn, dbeta, dgamma, total;
    GdnBwdWorkspace(int64_t n_img, int C, int64_t hw) {
        int64_t o = 0;
        dbeta = o;  o = align_up(o + 4ll * C, 256);
        dgamma = o; o = align_up(o + 4ll * C * C, 256);
        dn = o;     o = align_up(o + 4ll * n_img * C * hw, 256);
        total = o;
    }
};

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int64_t gdn_param_bytes(int C) {
    if (C <= 0) return MWA_ERR_INVALID;
    return GdnParamLayout(C).total;
}

int gdn_prepare(const float* beta_p, const float* gamma_p, int C, float beta_bound, float gamma_bound,
                float pedestal, void* params, int64_t params_bytes, void* stream) {
    if (!beta_p || !gamma_p || !params || C <= 0) return MWA_ERR_INVALID;
    if (!aligned16(params)) return MWA_ERR_ALIGNMENT;
    const GdnParamLayout L(C);
    if (params_bytes < L.total) return MWA_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* blk = static_cast<uint8_t*>(params);
    // operand images: padded rows / partial K blocks must read as zero
    zero_bytes_kernel<<<64, 256, 0, st>>>(reinterpret_cast<uint4*>(blk + L.img_hi), (L.total - L.img_hi) / 16);
    const int blocks = (C * C + 255) / 256;
    gdn_prepare_kernel<<<blocks, 256, 0, st>>>(beta_p, gamma_p, C, beta_bound, gamma_bound, pedestal, blk);
    return check_launch("gdn_prepare");
}

int gdn_forward(const float* x, float* y, const void* params, int64_t n_img, int C, int64_t hw, int inverse,
                int channels_last, int algo, void* stream) {
    if (n_img < 0 || hw < 0 || C <= 0) return MWA_ERR_INVALID;
    if (n_img == 0 || hw == 0) return MWA_OK;               // empty tensors carry null data pointers
    if (!x || !y || !params) return MWA_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (algo == MWA_ALGO_TCGEN05 || (algo == MWA_ALGO_AUTO && gdn_tc_supported(C, hw, channels_last))) {
        if (!gdn_tc_supported(C, hw, channels_last)) return MWA_ERR_UNSUPPORTED;
        if (!aligned16(x) || !aligned16(y)) return MWA_ERR_ALIGNMENT;
        return gdn_forward_tc(x, y, params, n_img, C, hw, inverse, channels_last, st);
    }
    if (C <= 64) return launch_gdn_simt<16>(x, y, params, n_img, C, hw, inverse, channels_last, st);
    if (C <= 128) return launch_gdn_simt<32>(x, y, params, n_img, C, hw, inverse, channels_last, st);
    if (C <= 192) return launch_gdn_simt<48>(x, y, params, n_img, C, hw, inverse, channels_last, st);
    if (C <= 256) return launch_gdn_simt<64>(x, y, params, n_img, C, hw, inverse, channels_last, st);
    if (C <= 512) return launch_gdn_simt<128>(x, y, params, n_img, C, hw, inverse, channels_last, st);
    return MWA_ERR_UNSUPPORTED;
}

int gdn_forward_planes(const float* x, void* out_hi, void* out_lo, int ps, int out_cstride, const void* params, int64_t n_img,
                       int C, int H, int W, int inverse, void* stream) {
    if (n_img < 0 || H < 0 || W < 0 || C <= 0) return MWA_ERR_INVALID;
    if (n_img == 0 || H == 0 || W == 0) return MWA_OK;
    if (!x || !out_hi || !out_lo || !params) return MWA_ERR_INVALID;
    if (!aligned16(x) || !aligned16(out_hi) || !aligned16(out_lo)) return MWA_ERR_ALIGNMENT;
    return gdn_forward_planes_tc(x, out_hi, out_lo, ps, out_cstride, params, n_img, C, H, W, inverse,
                                 static_cast<cudaStream_t>(stream));
}

int64_t gdn_backward_workspace_bytes(int64_t n_img, int C, int64_t hw) {
    if (n_img < 0 || C <= 0 || hw < 0) return MWA_ERR_INVALID;
    return GdnBwdWorkspace(n_img, C, hw).total;
}

int gdn_backward(const float* x, const float* grad_y, const float* beta_p, const float* gamma_p, const void* params,
                 float beta_bound, float gamma_bound, float* grad_x, float* grad_beta_p, float* grad_gamma_p,
                 void* workspace, int64_t workspace_bytes, int64_t n_img, int C, int64_t hw, int inverse,
                 int channels_last, void* stream) {
    if (!x || !grad_y || !beta_p || !gamma_p || !params || !grad_x || !grad_beta_p || !grad_gamma_p || !workspace)
        return MWA_ERR_INVALID;
    if (n_img < 0 || hw < 0 || C <= 0) return MWA_ERR_INVALID;
    if (C > 256) return MWA_ERR_UNSUPPORTED;
    const GdnBwdWorkspace Wk(n_img, C, hw);
    if (workspace_bytes < Wk.total) return MWA_ERR_WORKSPACE;
    if (!aligned16(workspace)) return MWA_ERR_ALIGNMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* wsp = static_cast<uint8_t*>(workspace);
    float* dbeta = reinterpret_cast<float*>(wsp + Wk.dbeta);
    float* dgamma = reinterpret_cast<float*>(wsp + Wk.dgamma);
    float* dn = reinterpret_cast<float*>(wsp + Wk.dn);
    MWA_TRY_CUDA(cudaMemsetAsync(wsp, 0, Wk.dn, st), "gdn_backward(memset)");
    if (n_img > 0 && hw > 0) {
        int rc;
        if (C <= 64) rc = launch_gdn_bwd_tile<16>(x, grad_y, grad_x, dn, dbeta, params, n_img, C, hw, inverse, channels_last, st);
        else if (C <= 128) rc = launch_gdn_bwd_tile<32>(x, grad_y, grad_x, dn, dbeta, params, n_img, C, hw, inverse, channels_last, st);
        else if (C <= 192) rc = launch_gdn_bwd_tile<48>(x, grad_y, grad_x, dn, dbeta, params, n_img, C, hw, inverse, channels_last, st);
        else rc = launch_gdn_bwd_tile<64>(x, grad_y, grad_x, dn, dbeta, params, n_img, C, hw, inverse, channels_last, st);
        if (rc != MWA_OK) return rc;
        const int tiles = (C + kGT - 1) / kGT;
        const int64_t total_px = n_img * hw;
        int64_t splits = (int64_t(kNumSMs) * 4 + tiles * tiles - 1) / (tiles * tiles);
        int64_t per = (total_px + splits - 1) / splits;
        per = (per + kGK - 1) / kGK * kGK;
        if (per < kGK) per = kGK;
        splits = (total_px + per - 1) / per;
        dim3 grid(tiles * tiles, static_cast<unsigned>(splits));
        gdn_bwd_dgamma_kernel<<<grid, 256, 0, st>>>(x, dn, dgamma, C, hw, n_img, per, channels_last);
        rc = check_launch("gdn_backward(dgamma)");
        if (rc != MWA_OK) return rc;
    }
    gdn_bwd_finalize_kernel<<<(C * C + 255) / 256, 256, 0, st>>>(beta_p, gamma_p, dbeta, dgamma, grad_beta_p,
                                                                 grad_gamma_p, C, beta_bound, gamma_bound);
    return check_launch("gdn_backward(finalize)");
}

int64_t gdn_param_offset(int C, int which) {
    if (C <= 0 || which < 0 || which > 2) return MWA_ERR_INVALID;
    const GdnParamLayout L(C);
    return which == 0 ? L.beta : which == 1 ? L.gamma : L.gammaT;
}

int gdn_bwd_square(const float* x, float* x2, int64_t n, void* stream) {
    if (!x || !x2 || n < 0 || n % 4 != 0) return MWA_ERR_INVALID;
    if (!aligned16(x) || !aligned16(x2)) return MWA_ERR_ALIGNMENT;
    if (n == 0) return MWA_OK;
    gdn_bwd_square_kernel<<<kNumSMs * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(x2), n / 4);
    return check_launch("gdn_bwd_square");
}

int gdn_bwd_dn(const float* grad_y, const float* x, float* n_to_term1, const void* params, float* dn, int64_t n_img,
               int C, int64_t hw, int inverse, int channels_last, void* stream) {
    if (!grad_y || !x || !n_to_term1 || !params || !dn || n_img < 0 || C <= 0 || hw < 0) return MWA_ERR_INVALID;
    const int64_t n = n_img * C * hw;
    if (n == 0) return MWA_OK;
    const GdnParamLayout L(C);
    const float* beta = reinterpret_cast<const float*>(static_cast<const uint8_t*>(params) + L.beta);
    gdn_bwd_dn_kernel<<<kNumSMs * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_y, x, n_to_term1, beta, dn, n, C,
                                                                                 hw, inverse, channels_last);
    return check_launch("gdn_bwd_dn");
}

int gdn_bwd_dx(const float* term1, const float* x, const float* t, float* grad_x, int64_t n, void* stream) {
    if (!term1 || !x || !t || !grad_x || n < 0 || n % 4 != 0) return MWA_ERR_INVALID;
    if (!aligned16(term1) || !aligned16(x) || !aligned16(t) || !aligned16(grad_x)) return MWA_ERR_ALIGNMENT;
    if (n == 0) return MWA_OK;
    gdn_bwd_dx_kernel<<<kNumSMs * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(term1), reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(t),
        reinterpret_cast<float4*>(grad_x), n / 4);
    return check_launch("gdn_bwd_dx");
}

int gdn_bwd_finalize(const float* dn, const float* dgamma_eff, const float* beta_p, const float* gamma_p,
                     float beta_bound, float gamma_bound, float* grad_beta_p, float* grad_gamma_p, float* dbeta_scratch,
                     int64_t n_img, int C, int64_t hw, int channels_last, void* stream) {
    if (!dn || !dgamma_eff || !beta_p || !gamma_p || !grad_beta_p || !grad_gamma_p || !dbeta_scratch) return MWA_ERR_INVALID;
    if (n_img < 0 || C <= 0 || hw < 0) return MWA_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MWA_TRY_CUDA(cudaMemsetAsync(dbeta_scratch, 0, sizeof(float) * C, st), "gdn_bwd_finalize(memset)");
    if (n_img > 0 && hw > 0) {
        gdn_bwd_dbeta_kernel<<<kNumSMs * 4, 256, 0, st>>>(dn, dbeta_scratch, n_img, C, hw, channels_last);
        int rc = check_launch("gdn_bwd_finalize(dbeta)");
        if (rc != MWA_OK) return rc;
    }
    gdn_bwd_finalize_kernel<<<(C * C + 255) / 256, 256, 0, st>>>(beta_p, gamma_p, dbeta_scratch, dgamma_eff, grad_beta_p,
                                                                 grad_gamma_p, C, beta_bound, gamma_bound);
    return check_launch("gdn_bwd_finalize");
}

}  // extern "C"
