// Micro-probes (development tool, B200): issue rates that bound the per-window attention core and the NCHW gather.
//   1. legacy mma.sync m16n8k16 f16 -> f32 throughput per SM for 1..16 warps (independent accumulators)
//   2. ldmatrix.x2 / x4 throughput
//   3. window gather from an NCHW tensor: thread = token (LDG.32, 4 window rows per request) against
//      lane = 4 tokens (LDG.128), for shift 0 and 4
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/sm_probe tools/sm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void hmma_kernel(float* out, int iters, long long* cyc) {
    float acc[8][4];
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, 3, 4}, b[2] = {5, threadIdx.x};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void hmma_dep_kernel(float* out, int iters, long long* cyc) {      // one dependent chain: latency
    float acc[4] = {0, 0, 0, 0};
    uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, 3, 4}, b[2] = {5, threadIdx.x};
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    const long long t1 = clock64();
    out[threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int X>
__global__ void ldm_kernel(float* out, int iters, long long* cyc) {
    __shared__ __align__(128) uint8_t buf[16384];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint32_t*>(buf)[i] = i;
    __syncthreads();
    uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(buf)) + (threadIdx.x & 31) * 128 + ((threadIdx.x >> 5) & 7) * 16;
    uint32_t s = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t r0, r1, r2 = 0, r3 = 0;
            const uint32_t addr = base ^ (((i + it) & 7) << 4);
            if (X == 4)
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
            else
                asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
            s += r0 ^ r1 ^ r2 ^ r3;
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// gather of 8x8 windows from NCHW (C channels, H x W), shift s; one CTA (256 threads) per pair of windows, copies
// x -> y (read + write) like the attention kernel's load / epilogue.  MODE 0: thread = token, LDG.32 / STG.32;
// MODE 1: lane = 4 tokens (float4), warp = channel group.
template <int MODE>
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int H, int W, int s,
                                                     int npairs) {
    const int nwx = W / 8;
    const int hw = H * W;
    for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (MODE == 0) {
            const int q = warp & 3, half = warp >> 2, r = q * 32 + lane, wslot = r >> 6, tok = r & 63;
            const int win = pair * 2 + wslot, b = win / ((H / 8) * nwx), rem = win % ((H / 8) * nwx);
            int py = (rem / nwx) * 8 + tok / 8 + s, px = (rem % nwx) * 8 + tok % 8 + s;
            if (py >= H) py -= H;
            if (px >= W) px -= W;
            const float* p = x + (size_t)b * C * hw + py * W + px + (size_t)half * (C / 2) * hw;
            float* o = y + (p - x);
            for (int c0 = 0; c0 < C / 2; c0 += 16) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __ldg(p + (size_t)(c0 + j) * hw);
#pragma unroll
                for (int j = 0; j < 16; ++j) o[(size_t)(c0 + j) * hw] = v[j] + 1.f;
            }
        } else {
            const int r0 = 4 * lane, wslot = r0 >> 6, tok0 = r0 & 63;
            const int win = pair * 2 + wslot, b = win / ((H / 8) * nwx), rem = win % ((H / 8) * nwx);
            int py = (rem / nwx) * 8 + tok0 / 8 + s, px = (rem % nwx) * 8 + tok0 % 8 + s;
            if (py >= H) py -= H;
            if (px >= W) px -= W;
            const float* p = x + (size_t)b * C * hw + py * W + px;
            float* o = y + (p - x);
            for (int c0 = warp; c0 < C; c0 += 8 * 8) {
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) if (c0 + 8 * j < C) v[j] = __ldg(reinterpret_cast<const float4*>(p + (size_t)(c0 + 8 * j) * hw));
#pragma unroll
                for (int j = 0; j < 8; ++j) if (c0 + 8 * j < C) {
                    v[j].x += 1.f; v[j].y += 1.f; v[j].z += 1.f; v[j].w += 1.f;
                    *reinterpret_cast<float4*>(o + (size_t)(c0 + 8 * j) * hw) = v[j];
                }
            }
        }
    }
}

int main() {
    float* out; long long* cyc; CK(cudaMalloc(&out, 1 << 24)); CK(cudaMalloc(&cyc, 8));
    long long h;
    const int iters = 2000;
    for (int warps : {1, 2, 4, 8, 12, 16}) {
        hmma_kernel<<<1, warps * 32>>>(out, iters, cyc);
        CK(cudaDeviceSynchronize());
        hmma_kernel<<<1, warps * 32>>>(out, iters, cyc);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("mma.sync m16n8k16 f16->f32: %2d warps  %.2f cycles per HMMA per warp, %.2f cycles per HMMA per SM (%.0f dense FMA/clk/SM)\n",
               warps, double(h) / (iters * 8), double(h) / (iters * 8 * warps), 2048.0 * iters * 8 * warps / double(h));
    }
    hmma_dep_kernel<<<1, 32>>>(out, iters, cyc);
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    printf("mma.sync dependent chain latency: %.1f cycles\n", double(h) / iters);
    for (int warps : {1, 4, 8, 16}) {
        ldm_kernel<2><<<1, warps * 32>>>(out, iters, cyc);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double c2 = double(h) / (iters * 8 * warps);
        ldm_kernel<4><<<1, warps * 32>>>(out, iters, cyc);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("ldmatrix: %2d warps  x2 %.2f cycles per instr per SM, x4 %.2f\n", warps, c2, double(h) / (iters * 8 * warps));
    }
    // gather
    const int B = 16, C = 192, H = 128, W = 192;
    const size_t n = (size_t)B * C * H * W;
    float *x, *y; CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&y, n * 4)); CK(cudaMemset(x, 0, n * 4));
    float* flush; CK(cudaMalloc(&flush, 512u << 20));
    const int npairs = B * (H / 8) * (W / 8) / 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int s : {0, 4}) for (int mode : {0, 1}) for (int grid : {148, 296, 3072}) {
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaMemsetAsync(flush, rep, 512u << 20));
            cudaEventRecord(e0);
            if (mode == 0) gather_kernel<0><<<grid, 256>>>(x, y, C, H, W, s, npairs);
            else gather_kernel<1><<<grid, 256>>>(x, y, C, H, W, s, npairs);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("gather NCHW 8x8 windows shift %d mode %s grid %4d: %.3f ms  %.0f GB/s (read+write)\n", s,
               mode == 0 ? "thread=token LDG.32 " : "lane=4 tokens LDG.128", grid, best, 2.0 * n * 4 / best / 1e6);
    }
    printf("PROBE DONE\n");
    return 0;
}
