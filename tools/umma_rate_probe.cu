// Hardware probe: cost per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N and of the operand
// sources, measured as clock64 cycles from the first issue to the completion of a chain of `reps` MMAs (one thread
// issues, whole warp converged, elect_one).  Garbage operands (zeros); only timing matters.
#include <cstdio>
#include <cstdlib>
#include "../deep-learning-based-rgba-image-compression-with-masked-window-based-attention_b200/csrc/common.cuh"
using namespace b200;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

// MODE 0: SS (A, B from smem)   1: TS (A from TMEM)   2: TS + MN-major B   4: SS lane-masked (half the lanes disabled)
// 5: an earlier QKV pattern (SS N, TS 32, SS 32 per k step)   6: the P V pattern of csrc/mwa_sp.cu (TS, MN-major B,
// lane-masked halves alternating, second B start advanced inside the atom)   7: the S pattern (TS, lane-masked halves)
// 8: the QKV pattern (SS N, TS N, SS N per k step)
template <int N, int MODE, bool LANE0>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;
    if (warp == 0) {
        constexpr uint32_t idesc = umma_idesc(kFmtF16, kFmtF16, 128, N) | ((MODE == 2 || MODE == 6) ? kUmmaBMajorMN : 0u);
        constexpr uint32_t id32 = umma_idesc(kFmtF16, kFmtF16, 128, 32);
        const uint64_t ad = umma_desc_k_sw128(smem_u32(smem)), bd = umma_desc_k_sw128(smem_u32(smem + 32768));
        long long t0 = 0, t1 = 0, t2 = 0;
        bool me;
        if constexpr (LANE0) me = (tid == 0);
        else me = elect_one();
        t0 = clock64();
        if (me) {
            for (int r = 0; r < reps; r += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t ks = u & 3;
                    if constexpr (MODE == 0) umma_f16_ss(tm, ad + ks * 2, bd + ks * 2, idesc, 1);
                    else if constexpr (MODE == 1) umma_f16_ts(tm, tm + 256 + ks * 8, bd + ks * 2, idesc, 1);
                    else if constexpr (MODE == 2) umma_f16_ts(tm, tm + 256 + ks * 8, bd + ks * 128, idesc, 1);
                    else if constexpr (MODE == 4) {
                        if (u & 1) umma_f16_ss_lanes(tm, ad + ks * 2, bd + ks * 2, idesc, 1, ~0u, ~0u, 0u, 0u);
                        else umma_f16_ss_lanes(tm, ad + ks * 2, bd + ks * 2, idesc, 1, 0u, 0u, ~0u, ~0u);
                    } else if constexpr (MODE == 6) {
                        if (u & 1) umma_f16_ts_lanes(tm, tm + 256 + ks * 8, bd + ks * 128 + 4, idesc, 1, ~0u, ~0u, 0u, 0u);
                        else umma_f16_ts_lanes(tm, tm + 256 + ks * 8, bd + ks * 128, idesc, 1, 0u, 0u, ~0u, ~0u);
                    } else if constexpr (MODE == 7) {
                        if (u & 1) umma_f16_ts_lanes(tm, tm + 256 + ks * 8, bd + 512 + ks * 2, idesc, 1, ~0u, ~0u, 0u, 0u);
                        else umma_f16_ts_lanes(tm, tm + 256 + ks * 8, bd + ks * 2, idesc, 1, 0u, 0u, ~0u, ~0u);
                    } else if constexpr (MODE == 8) {
                        umma_f16_ss(tm, ad + ks * 2, bd + ks * 2, idesc, 1);
                        umma_f16_ts(tm, tm + 256 + ks * 8, bd + ks * 2, idesc, 1);
                        umma_f16_ss(tm, ad + ks * 2, bd + 1024 + ks * 2, idesc, 1);
                    } else {
                        umma_f16_ss(tm, ad + ks * 2, bd + ks * 2, idesc, 1);
                        umma_f16_ts(tm + 48, tm + 256 + ks * 8, bd + 48 * 8 + ks * 2, id32, 1);
                        umma_f16_ss(tm + 48, ad + ks * 2, bd + 80 * 8 + ks * 2, id32, 1);
                    }
                }
            }
            t1 = clock64();
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t2 = clock64();
        if (me) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N, int MODE, bool LANE0>
void run(long long* d) {
    const char* names[] = {"SS", "TS", "TS+MN", "", "SS masked", "QKV old", "PV masked", "S masked", "QKV mix"};
    const int reps = 512;
    CK(cudaFuncSetAttribute(rate_kernel<N, MODE, LANE0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    for (int w = 0; w < 2; ++w) {
        rate_kernel<N, MODE, LANE0><<<1, 128, 96 * 1024>>>(reps, d);
        CK(cudaDeviceSynchronize());
    }
    long long h[2];
    CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
    const int n_mma = (MODE == 5 || MODE == 8) ? 3 * reps : reps;
    printf("%-9s N=%3d %s: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", names[MODE], N,
           LANE0 ? "lane0-branch" : "elect_one", double(h[0]) / n_mma, double(h[1]) / n_mma, N / 2);
}

int main() {
    long long* d;
    CK(cudaMalloc(&d, 16));
    run<16, 0, false>(d); run<32, 0, false>(d); run<64, 0, false>(d); run<80, 0, false>(d); run<96, 0, false>(d);
    run<128, 0, false>(d); run<192, 0, false>(d); run<256, 0, false>(d);
    run<16, 1, false>(d); run<32, 1, false>(d); run<96, 1, false>(d); run<192, 1, false>(d);
    run<48, 2, false>(d); run<64, 2, false>(d);
    run<64, 4, false>(d);
    run<80, 5, false>(d);
    run<32, 6, false>(d); run<32, 2, false>(d); run<64, 7, false>(d); run<80, 8, false>(d); run<96, 8, false>(d);
    run<16, 2, false>(d); run<16, 6, false>(d);
    run<32, 0, true>(d); run<96, 0, true>(d); run<192, 0, true>(d); run<32, 1, true>(d); run<80, 5, true>(d);
    return 0;
}
