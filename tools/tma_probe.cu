// TMA probe (development tool, B200): can cp.async.bulk.tensor gather / scatter the 8x8-pixel windows of an NCHW fp32
// tensor (32-byte inner extent, 16-byte-offset box origins for shift 4) at a useful rate?  Decides whether the attention
// kernel's global traffic can move from the LSU (2.3 cycles per request x 128-byte line, tools/sm_probe.cu) to the TMA
// engine.  Measures, for all 6144 windows of a (16,192,128,192) tensor:
//   copy   : tiled load of a [CB ch][8][8] box per window -> smem -> tiled store to `out`            (read + write)
//   reduce : same, but the store is cp.reduce.async.bulk.tensor ... .add.f32 onto a pre-filled `out`  (the epilogue)
// and checks the results (shift 0 and 4, interior windows and windows that hang over the right / bottom border:
// out-of-bound elements are zero-filled on load and skipped on store).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../deep-learning-based-rgba-image-compression-with-masked-window-based-attention_b200/csrc/common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

using namespace b200;
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { mbar_arrive_expect_tx(b, bytes); }
__device__ __forceinline__ void tma_load4(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    tma_load_4d(dst, map, c0, c1, c2, c3, bar);
}
__device__ __forceinline__ void tma_store4(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    tma_store_4d(map, src, c0, c1, c2, c3);
}
__device__ __forceinline__ void tma_red4(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
    tma_reduce_add_4d(map, src, c0, c1, c2, c3);
}

// One CTA walks over windows (grid-stride).  Per window: C / CB boxes of [CB][8][8] floats through a ring of STAGES
// buffers.  Thread 0 issues the loads (STAGES - 1 ahead) and the stores; the other threads add 1.0 to the staged values
// in the `reduce` mode (so the result is checkable) -- a stand-in for the epilogue's writes into the staging buffer.
template <int CB, int STAGES>
__global__ void __launch_bounds__(128) tma_window_kernel(const __grid_constant__ CUtensorMap in_map,
                                                         const __grid_constant__ CUtensorMap out_map, int nwx, int nwy,
                                                         int nwin, int C, int shift, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* stage = reinterpret_cast<float*>(smem);                       // STAGES x [CB][64]
    __shared__ uint64_t full[STAGES];
    const int tid = threadIdx.x;
    constexpr uint32_t kBoxBytes = CB * 64 * 4;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int chunks = C / CB;
    int my = 0;
    for (int w = blockIdx.x; w < nwin; w += gridDim.x) ++my;
    const int total = my * chunks;
    auto coords = [&](int job, int& x0, int& y0, int& c0, int& b) {
        const int win = blockIdx.x + (job / chunks) * gridDim.x;
        b = win / (nwx * nwy);
        const int r = win % (nwx * nwy);
        y0 = (r / nwx) * 8 + shift;
        x0 = (r % nwx) * 8 + shift;
        c0 = (job % chunks) * CB;
    };
    if (tid == 0) {
        for (int j = 0; j < STAGES - 1 && j < total; ++j) {
            int x0, y0, c0, b;
            coords(j, x0, y0, c0, b);
            mbar_expect(full + j % STAGES, kBoxBytes);
            tma_load4(stage + (j % STAGES) * CB * 64, &in_map, x0, y0, c0, b, full + j % STAGES);
        }
    }
    for (int j = 0; j < total; ++j) {
        const int s = j % STAGES;
        // refill the stage that job j - 1 used: its store must have finished READING the buffer
        if (tid == 0) {
            const int jn = j + STAGES - 1;
            if (jn < total) {
                bulk_wait_group_read<0>();
                int x0, y0, c0, b;
                coords(jn, x0, y0, c0, b);
                mbar_expect(full + jn % STAGES, kBoxBytes);
                tma_load4(stage + (jn % STAGES) * CB * 64, &in_map, x0, y0, c0, b, full + jn % STAGES);
            }
        }
        mbar_wait(full + s, (j / STAGES) & 1);
        float* buf = stage + s * CB * 64;
        if (mode == 1) {
            for (int e = tid; e < CB * 64; e += 128) buf[e] = 1.0f;      // "projection" contribution
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0 && mode != 2) {                                     // mode 2: loads only (the gather alone)
            int x0, y0, c0, b;
            coords(j, x0, y0, c0, b);
            if (mode == 0) tma_store4(&out_map, buf, x0, y0, c0, b);
            else tma_red4(&out_map, buf, x0, y0, c0, b);
            bulk_commit_group();
        }
    }
    if (tid == 0) bulk_wait_group<0>();
}

// Same gather (loads only), but NW warps of ONE CTA per SM each run their own ring: does the per-SM rate scale with the
// number of issuing warps inside a CTA the way it does with the number of CTAs?
template <int CB, int STAGES, int NW>
__global__ void __launch_bounds__(NW * 32) tma_gather_warps_kernel(const __grid_constant__ CUtensorMap in_map, int nwx, int nwy,
                                                                   int nwin, int C, int shift) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[NW][STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* stage = reinterpret_cast<float*>(smem) + warp * STAGES * CB * 64;
    constexpr uint32_t kBoxBytes = CB * 64 * 4;
    if (threadIdx.x == 0) {
        for (int w = 0; w < NW; ++w)
            for (int s = 0; s < STAGES; ++s) mbar_init(&full[w][s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int chunks = C / CB;
    int my = 0;
    for (int w = blockIdx.x; w < nwin; w += gridDim.x) ++my;
    const int total_cta = my * chunks;
    const int total = (total_cta - warp + NW - 1) / NW;                 // jobs warp, warp + NW, ...
    auto issue = [&](int jw) {
        const int job = warp + jw * NW;
        const int win = blockIdx.x + (job / chunks) * gridDim.x;
        const int b = win / (nwx * nwy), r = win % (nwx * nwy);
        mbar_expect(&full[warp][jw % STAGES], kBoxBytes);
        tma_load4(stage + (jw % STAGES) * CB * 64, &in_map, (r % nwx) * 8 + shift, (r / nwx) * 8 + shift, (job % chunks) * CB, b,
                  &full[warp][jw % STAGES]);
    };
    if (lane == 0)
        for (int j = 0; j < STAGES - 1 && j < total; ++j) issue(j);
    for (int j = 0; j < total; ++j) {
        if (lane == 0 && j + STAGES - 1 < total) issue(j + STAGES - 1);
        mbar_wait(&full[warp][j % STAGES], (j / STAGES) & 1);
        __syncwarp();
    }
}

// Are NEGATIVE box coordinates legal for a tiled LOAD (they are an illegal instruction for the reduce-add store)?  One box at
// (x0, y0) = (-4, -4): the out-of-bound part must arrive as zeros -- what a wrapped window of the cyclic shift needs.
__global__ void tma_negative_load_kernel(const __grid_constant__ CUtensorMap in_map, float* dbg, int x0, int y0) {
    __shared__ __align__(128) float box[16 * 64];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect(&bar, sizeof(box));
        tma_load4(box, &in_map, x0, y0, 0, 0, &bar);
    }
    mbar_wait(&bar, 0);
    for (int e = threadIdx.x; e < 16 * 64; e += blockDim.x) dbg[e] = box[e];
}

int main() {
    const int B = 16, C = 192, H = 128, W = 192;
    const size_t n = (size_t)B * C * H * W;
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    float *x, *y, *flush;
    CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&y, n * 4)); CK(cudaMalloc(&flush, 512u << 20));
    std::vector<float> hx(n);
    for (size_t i = 0; i < n; ++i) hx[i] = float((i * 2654435761u) % 1000) * 0.001f;
    CK(cudaMemcpy(x, hx.data(), n * 4, cudaMemcpyHostToDevice));
    auto make_map = [&](float* base, int cb, CUtensorMap* m) -> bool {
        cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t box[4] = {8, 8, (cuuint32_t)cb, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return false; }
        return true;
    };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int nwx = W / 8, nwy = H / 8, nwin = B * nwx * nwy;
    std::vector<float> hy(n);
    bool all_ok = true;
    // deeper pipelines, loads only: what the gather alone can reach (8 stages of [16 ch][8][8] = 32 KB)
    {
        CUtensorMap in_map, out_map;
        if (!make_map(x, 16, &in_map) || !make_map(y, 16, &out_map)) return 1;
        // more CTAs per SM (4 stages each): does the per-SM rate scale with the number of independent issue streams?
        for (int shift : {0, 4}) for (int grid : {296, 592, 1184}) {
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemsetAsync(flush, rep, 512u << 20));
                cudaEventRecord(e0);
                const int smem = 4 * 16 * 64 * 4;
                cudaFuncSetAttribute(tma_window_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                tma_window_kernel<16, 4><<<grid, 128, smem>>>(in_map, out_map, nwx, nwy, nwin, C, shift, 2);
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("TMA window gather only  box [16 ch][8][8]  shift %d   4 stages  grid %4d (%d CTAs per SM): %.3f ms  %.0f GB/s (read)\n",
                   shift, grid, grid / 148, best, 1.0 * n * 4 / best / 1e6);
        }
        for (int shift : {0, 4}) for (int nw : {1, 2, 4, 8}) {
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemsetAsync(flush, rep, 512u << 20));
                cudaEventRecord(e0);
                const int smem = nw * 4 * 16 * 64 * 4;
#define LAUNCH_W(NW_)                                                                                                   \
    cudaFuncSetAttribute(tma_gather_warps_kernel<16, 4, NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);      \
    tma_gather_warps_kernel<16, 4, NW_><<<148, NW_ * 32, smem>>>(in_map, nwx, nwy, nwin, C, shift);
                if (nw == 1) { LAUNCH_W(1) } else if (nw == 2) { LAUNCH_W(2) } else if (nw == 4) { LAUNCH_W(4) } else { LAUNCH_W(8) }
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("TMA window gather only  box [16 ch][8][8]  shift %d   ONE CTA per SM, %d issuing warps x 4 stages: %.3f ms  %.0f GB/s "
                   "(read) = %.0f cycles per 128-token tile at 1.9 GHz\n", shift, nw, best, 1.0 * n * 4 / best / 1e6,
                   best * 1e-3 * 1.9e9 / (6144.0 / 2 / 148));
        }
        for (int shift : {0, 4}) for (int stages : {4, 8, 12}) {
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemsetAsync(flush, rep, 512u << 20));
                cudaEventRecord(e0);
                const int smem = stages * 16 * 64 * 4;
                if (stages == 4) {
                    cudaFuncSetAttribute(tma_window_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    tma_window_kernel<16, 4><<<148, 128, smem>>>(in_map, out_map, nwx, nwy, nwin, C, shift, 2);
                } else if (stages == 8) {
                    cudaFuncSetAttribute(tma_window_kernel<16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    tma_window_kernel<16, 8><<<148, 128, smem>>>(in_map, out_map, nwx, nwy, nwin, C, shift, 2);
                } else {
                    cudaFuncSetAttribute(tma_window_kernel<16, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    tma_window_kernel<16, 12><<<148, 128, smem>>>(in_map, out_map, nwx, nwy, nwin, C, shift, 2);
                }
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("TMA window gather only  box [16 ch][8][8]  shift %d  %2d stages  grid 148: %.3f ms  %.0f GB/s (read)  = %.0f cycles "
                   "per 128-token tile at 1.9 GHz\n", shift, stages, best, 1.0 * n * 4 / best / 1e6, best * 1e-3 * 1.9e9 / (6144.0 / 2 / 148));
        }
    }
    for (int cb : {16, 32, 64}) {
        CUtensorMap in_map, out_map;
        if (!make_map(x, cb, &in_map) || !make_map(y, cb, &out_map)) return 1;
        for (int shift : {0, 4}) for (int mode : {0, 1}) for (int grid : {148, 296}) {
            float best = 1e9;
            for (int rep = 0; rep < 3; ++rep) {
                if (mode == 1) CK(cudaMemcpy(y, x, n * 4, cudaMemcpyDeviceToDevice));
                else CK(cudaMemset(y, 0, n * 4));
                CK(cudaMemsetAsync(flush, rep, 512u << 20));
                cudaEventRecord(e0);
                const int stages = 4;
                const int smem = stages * cb * 64 * 4;
                if (cb == 16) {
                    cudaFuncSetAttribute(tma_window_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    tma_window_kernel<16, 4><<<grid, 128, smem>>>(in_map, out_map, nwx, nwy, nwin, C, shift, mode);
                } else if (cb == 32) {
                    cudaFuncSetAttribute(tma_window_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    tma_window_kernel<32, 4><<<grid, 128, smem>>>(in_map, out_map, nwx, nwy, nwin, C, shift, mode);
                } else {
                    cudaFuncSetAttribute(tma_window_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                    tma_window_kernel<64, 4><<<grid, 128, smem>>>(in_map, out_map, nwx, nwy, nwin, C, shift, mode);
                }
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            // check: windows cover [shift, H) x [shift, W) (the parts hanging over the border are clipped)
            CK(cudaMemcpy(hy.data(), y, n * 4, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t i = 0; i < n; i += 97) {
                const int px = i % W, py = (i / W) % H;
                const bool covered = px >= shift && py >= shift;
                const float want = mode == 0 ? (covered ? hx[i] : 0.f) : (covered ? hx[i] + 1.0f : hx[i]);
                if (hy[i] != want) ++bad;
            }
            if (bad) all_ok = false;
            printf("TMA window %s  box [%2d ch][8][8]  shift %d  grid %3d: %.3f ms  %.0f GB/s (read+write)  %s\n",
                   mode == 0 ? "copy  " : "reduce", cb, shift, grid, best, 2.0 * n * 4 / best / 1e6, bad ? "MISMATCH" : "ok");
        }
    }
    {   // negative coordinates on the load path (run last: a fault here ends the process)
        CUtensorMap in_map;
        if (!make_map(x, 16, &in_map)) return 1;
        float* dbg;
        CK(cudaMalloc(&dbg, 16 * 64 * 4));
        for (int neg : {0, 1}) {
            const int x0 = neg ? -4 : W - 4, y0 = neg ? -4 : H - 4;
            tma_negative_load_kernel<<<1, 128>>>(in_map, dbg, x0, y0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("TMA load at (%d, %d): %s\n", x0, y0, cudaGetErrorString(e)); all_ok = false; break; }
            std::vector<float> hb(16 * 64);
            CK(cudaMemcpy(hb.data(), dbg, hb.size() * 4, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (int c = 0; c < 16; ++c) for (int yy = 0; yy < 8; ++yy) for (int xx = 0; xx < 8; ++xx) {
                const int gx = x0 + xx, gy = y0 + yy;
                const float want = (gx < 0 || gy < 0 || gx >= W || gy >= H) ? 0.f : hx[((size_t)c * H + gy) * W + gx];
                if (hb[c * 64 + yy * 8 + xx] != want) ++bad;
            }
            printf("TMA load of one box at (%d, %d): %s (out-of-bound part zero-filled, %zu mismatches)\n", x0, y0,
                   bad ? "MISMATCH" : "ok", bad);
            if (bad) all_ok = false;
        }
    }
    printf(all_ok ? "PROBE PASSED\n" : "PROBE FAILED\n");
    return 0;
}
