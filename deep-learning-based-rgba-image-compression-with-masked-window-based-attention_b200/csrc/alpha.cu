// Alpha pyramid: the cascade of AvgPool2d(3, stride 2, padding 1, count_include_pad) that turns the alpha plane into the
// per-scale masks the masked window attention consumes (reference layers/SupplyMask.py:7-18; wiring
// layers/TransformRGB.py:68,72,91,95), optionally preceded by the decoder's mask quantisation round(255 a) / 255
// (models/AutoEncoderRGB_Journal.py:212-214).
//
// The reference runs six pooling launches (+ three elementwise ones for the quantisation).  Here THREE levels are produced
// per launch: a CTA owns an 8 x 8 block of the coarsest of its three levels and recomputes the halo it needs of the two
// finer ones in shared memory (source tile 71 x 71 -> 35 x 35 -> 17 x 17 -> 8 x 8), so the source is read 1.23 x and every
// level is written once.  Six levels = two launches of the same kernel (the second on the 1/8-scale plane).
// Bit-exact with torch: each output is the row-major fp32 sum of its (up to) nine in-range taps divided by 9.
#include "common.cuh"
#include "status.cuh"

namespace b200 {
namespace {

constexpr int kOwn3 = 8;                      // owned block of the third level
constexpr int kT2 = 2 * kOwn3 + 1;            // 17: second level incl. the one-row halo above / left
constexpr int kT1 = 2 * kT2 + 1;              // 35
constexpr int kT0 = 2 * kT1 + 1;              // 71

__host__ __device__ inline int pooled(int n) { return (n - 1) / 2 + 1; }      // floor((n + 2 - 3) / 2) + 1

// dst[r][c] (tile coordinates, origin (oy, ox) in the level's plane) = avg of the 3 x 3 source taps around (2r+1, 2c+1)
// of the source tile.  Positions outside the level's plane are the NEXT pooling's zero padding: they must be zero.
template <int TS, int TD>
__device__ __forceinline__ void pool_tile(const float* __restrict__ src, float* __restrict__ dst, int oy, int ox, int Hd,
                                          int Wd, int tid, int nthreads) {
    for (int e = tid; e < TD * TD; e += nthreads) {
        const int r = e / TD, c = e % TD;
        const int y = oy + r, x = ox + c;
        float v = 0.f;
        if (y >= 0 && y < Hd && x >= 0 && x < Wd) {
            float s = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) s += src[(2 * r + dy) * TS + 2 * c + dx];
            v = __fdiv_rn(s, 9.0f);
        }
        dst[e] = v;
    }
}

// write the owned part of a tile (tile rows / cols >= 1: row / col 0 is halo) to the level's plane
template <int TD>
__device__ __forceinline__ void store_owned(const float* __restrict__ tile, float* __restrict__ plane, int oy, int ox, int Hd,
                                            int Wd, int tid, int nthreads) {
    constexpr int OWN = TD - 1;
    for (int e = tid; e < OWN * OWN; e += nthreads) {
        const int r = e / OWN + 1, c = e % OWN + 1;
        const int y = oy + r, x = ox + c;
        if (y < Hd && x < Wd) plane[int64_t(y) * Wd + x] = tile[r * TD + c];
    }
}

// src: (B, Hs, Ws) plane.  l1 / l2 / l3: the next three levels (l2, l3 may be null when fewer levels are asked for).
// quant > 0: the source is first quantised to round(quant * a) / quant, and written to `recon` (owned 64 x 64 part).
__global__ void __launch_bounds__(256)
alpha_pyramid3_kernel(const float* __restrict__ src, float* __restrict__ recon, float* __restrict__ l1, float* __restrict__ l2,
                      float* __restrict__ l3, int Hs, int Ws, float quant) {
    __shared__ float t0[kT0 * kT0];
    __shared__ float t1[kT1 * kT1];
    __shared__ float t2[kT2 * kT2];
    const int tid = threadIdx.x, b = blockIdx.z;
    const int H1 = pooled(Hs), W1 = pooled(Ws), H2 = pooled(H1), W2 = pooled(W1), H3 = pooled(H2), W3 = pooled(W2);
    const int y3 = blockIdx.y * kOwn3, x3 = blockIdx.x * kOwn3;         // owned block origin at level 3
    // tile origins (may be negative: halo above / left of the plane)
    const int oy2 = 2 * y3 - 1, ox2 = 2 * x3 - 1;
    const int oy1 = 2 * oy2 - 1, ox1 = 2 * ox2 - 1;
    const int oy0 = 2 * oy1 - 1, ox0 = 2 * ox1 - 1;
    const float* sp = src + int64_t(b) * Hs * Ws;
    float* rp = recon ? recon + int64_t(b) * Hs * Ws : nullptr;
    // (batching these loads per thread was measured: no gain at 124 registers, slower at 64 -- the kernel is bound by CTA
    //  turnover on an L2-resident plane, not by load latency)
    for (int e = tid; e < kT0 * kT0; e += 256) {
        const int r = e / kT0, c = e % kT0;
        const int y = oy0 + r, x = ox0 + c;
        float v = 0.f;
        if (y >= 0 && y < Hs && x >= 0 && x < Ws) {
            v = __ldg(sp + int64_t(y) * Ws + x);
            if (quant > 0.f) {
                v = __fdiv_rn(rintf(__fmul_rn(v, quant)), quant);
                // the owned source region of this CTA: tile rows / cols >= 7 (64 x 64)
                if (rp && r >= 7 && c >= 7) rp[int64_t(y) * Ws + x] = v;
            }
        }
        t0[e] = v;
    }
    __syncthreads();
    pool_tile<kT0, kT1>(t0, t1, oy1, ox1, H1, W1, tid, 256);
    __syncthreads();
    // owned level-1 part: tile rows / cols >= 3 (32 x 32)
    {
        float* p1 = l1 + int64_t(b) * H1 * W1;
        for (int e = tid; e < 32 * 32; e += 256) {
            const int r = e / 32 + 3, c = e % 32 + 3;
            const int y = oy1 + r, x = ox1 + c;
            if (y < H1 && x < W1) p1[int64_t(y) * W1 + x] = t1[r * kT1 + c];
        }
    }
    if (l2 == nullptr) return;
    pool_tile<kT1, kT2>(t1, t2, oy2, ox2, H2, W2, tid, 256);
    __syncthreads();
    store_owned<kT2>(t2, l2 + int64_t(b) * H2 * W2, oy2, ox2, H2, W2, tid, 256);
    if (l3 == nullptr) return;
    // third level: no halo needed (tile origin = owned origin)
    for (int e = tid; e < kOwn3 * kOwn3; e += 256) {
        const int r = e / kOwn3, c = e % kOwn3;
        const int y = y3 + r, x = x3 + c;
        if (y < H3 && x < W3) {
            float s = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) s += t2[(2 * r + dy) * kT2 + 2 * c + dx];
            l3[(int64_t(b) * H3 + y) * W3 + x] = __fdiv_rn(s, 9.0f);
        }
    }
}

// Isolated-pixel clean-up of the decoded mask (reference trainRGB.py:98-111 = trainmask.py:133-146 `constraint`), with the
// clamp + quantisation that precedes it at its call site (trainRGB.py:285-287) folded in when quant > 0:
//   m = round(clamp(m, 0, 1) * quant) / quant ;  s = sum of the 8 neighbours (zero padding)
//   out = 1 if (m == 0 and s == 8);  0 if (m > 0 and s == 0);  m otherwise
// The reference does it with a convolution, two boolean masks and two masked assignments (host-synchronising
// index_put); both masks are taken from the tensor BEFORE either assignment, so it is a pure 3 x 3 stencil.
__global__ void __launch_bounds__(256)
mask_constraint_kernel(const float* __restrict__ m, float* __restrict__ out, int H, int W, float quant, int64_t total) {
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int x = int(idx % W), y = int((idx / W) % H);
    const float* plane = m + (idx - int64_t(y) * W - x);
    auto at = [&](int yy, int xx) -> float {
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) return 0.f;
        float v = __ldg(plane + int64_t(yy) * W + xx);
        if (quant > 0.f) v = __fdiv_rn(rintf(__fmul_rn(fminf(fmaxf(v, 0.f), 1.f), quant)), quant);
        return v;
    };
    float s = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
            if (dy != 0 || dx != 0) s += at(y + dy, x + dx);
    const float c = at(y, x);
    out[idx] = (c == 0.f && s == 8.f) ? 1.f : ((c > 0.f && s == 0.f) ? 0.f : c);
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

int64_t alpha_pyramid_level_offset(int B, int H, int W, int level) {
    if (B < 0 || H < 1 || W < 1 || level < 0 || level > 6) return -1;
    int64_t off = 0;
    int h = H, w = W;
    for (int k = 0; k < level; ++k) {
        h = pooled(h);
        w = pooled(w);
        off += int64_t(B) * h * w;
    }
    return off;
}

int alpha_pyramid_forward(const float* alpha, float* recon, float* levels, int B, int H, int W, int nlevels,
                                     int quant_levels, void* stream) {
    if (B < 0 || H < 1 || W < 1 || nlevels < 1 || nlevels > 6 || quant_levels < 0) return MWA_ERR_INVALID;
    if (B == 0) return MWA_OK;
    if (!alpha || !levels) return MWA_ERR_INVALID;
    if (B > 65535) return MWA_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float* src = alpha;
    int hs = H, ws = W;
    float* out = levels;
    for (int k = 0; k < nlevels; k += 3) {
        const int h1 = pooled(hs), w1 = pooled(ws), h2 = pooled(h1), w2 = pooled(w1), h3 = pooled(h2), w3 = pooled(w2);
        float* l1 = out;
        float* l2 = (k + 1 < nlevels) ? l1 + int64_t(B) * h1 * w1 : nullptr;
        float* l3 = (k + 2 < nlevels) ? l2 + int64_t(B) * h2 * w2 : nullptr;
        const dim3 grid((w3 + kOwn3 - 1) / kOwn3, (h3 + kOwn3 - 1) / kOwn3, B);
        if (grid.y > 65535) return MWA_ERR_UNSUPPORTED;
        alpha_pyramid3_kernel<<<grid, 256, 0, st>>>(src, k == 0 ? recon : nullptr, l1, l2, l3, hs, ws,
                                                    k == 0 ? float(quant_levels) : 0.f);
        const int rc = check_launch("alpha_pyramid_forward");
        if (rc != MWA_OK) return rc;
        if (l3 == nullptr) break;
        src = l3;
        hs = h3;
        ws = w3;
        out = l3 + int64_t(B) * h3 * w3;
    }
    return MWA_OK;
}

int mask_constraint_forward(const float* mask, float* out, int B, int H, int W, int quant_levels, void* stream) {
    if (B < 0 || H < 1 || W < 1 || quant_levels < 0) return MWA_ERR_INVALID;
    if (B == 0) return MWA_OK;
    if (!mask || !out || mask == out) return MWA_ERR_INVALID;          // a stencil cannot run in place
    const int64_t total = int64_t(B) * H * W;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffll) return MWA_ERR_UNSUPPORTED;
    mask_constraint_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        mask, out, H, W, float(quant_levels), total);
    return check_launch("mask_constraint_forward");
}

}  // extern "C"
