// Hardware probe for the CTA-pair tensor-core path (tcgen05 cta_group::2), groundwork for splitting the attention /
// GDN weight operands across two SMs (run on the B200 box under `timeout`):
//   D[256 x N] = A[256 x K] * B[N x K]^T,   cluster of 2 CTAs, each CTA holds ITS 128 rows of A and HALF of the N rows
//   of B in shared memory (K-major SWIZZLE_128B), the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), the
//   accumulator lands in each CTA's own TMEM (128 lanes x N columns), completion is multicast to both CTAs' mbarriers.
// Prints max |err| against a host reference.  Exit code 0 = all pass.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../deep-learning-based-rgba-image-compression-with-masked-window-based-attention_b200/csrc/common.cuh"

using namespace b200;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, float* __restrict__ D, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int kblocks = (K + 63) / 64, NH = N / 2;
    uint8_t* sA = smem;                                  // kblocks x [128 rows x 128 B]   rows 128*rank .. +127 of A
    uint8_t* sB = smem + kblocks * 128 * 128;            // kblocks x [NH rows x 128 B]    rows NH*rank .. of B
    for (int i = tid; i < (kblocks * (128 + NH) * 128) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<uint16_t*>(sA + (k / 64) * (128 * 128) + sw128_offset(r, k % 64)) = A[(size_t)(128 * rank + r) * K + k];
    }
    for (int i = tid; i < NH * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<uint16_t*>(sB + (k / 64) * (NH * 128) + sw128_offset(r, k % 64)) = B[(size_t)(NH * rank + r) * K + k];
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {      // both CTAs of the pair allocate (collective over the pair)
        tmem_alloc_pair<256>(&tmem_base_s);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                  // both CTAs' operands + barriers + TMEM are ready
    tc_fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;

    if (rank == 0 && tid == 0) {
        // instruction descriptor: M = 256 (the pair), N columns
        const uint32_t idesc = umma_idesc(kFmtF16, kFmtF16, 256, N);
        for (int kb = 0; kb < kblocks; ++kb) {
            const uint64_t a0 = umma_desc_k_sw128(smem_u32(sA + kb * 128 * 128));
            const uint64_t b0 = umma_desc_k_sw128(smem_u32(sB + kb * NH * 128));
            const int nks = (K - kb * 64 >= 64) ? 4 : (K - kb * 64 + 15) / 16;
            for (int ks = 0; ks < nks; ++ks) {
                umma_f16_ss_pair(tmem_d, a0 + ks * 2, b0 + ks * 2, idesc, (kb | ks) != 0);
            }
        }
        // completion -> the mbarrier at this smem offset in BOTH CTAs
        umma_commit_pair(&bar, 3);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    // each CTA reads its 128 rows (TMEM lanes) x N columns
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t acc[8];
        tmem_ld_x8(tmem_d + lane_addr + c0, acc);
        tmem_wait_ld();
        for (int j = 0; j < 8; ++j) D[(size_t)(128 * rank + tid) * N + c0 + j] = __uint_as_float(acc[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc_pair<256>(tmem_d);
}

int main() {
    bool all = true;
    for (int N : {64, 144, 192, 256}) for (int K : {64, 192}) {
        std::vector<uint16_t> hA(256 * K), hB(N * K);
        std::vector<float> fA(256 * K), fB(N * K), ref(256 * N), hD(256 * N);
        srand(N * 131 + K);
        auto h16 = [](float v, float& back) { __half h = __float2half(v); back = __half2float(h); uint16_t u; memcpy(&u, &h, 2); return u; };
        for (int i = 0; i < 256 * K; ++i) hA[i] = h16((rand() % 2001 - 1000) * 0.001f, fA[i]);
        for (int i = 0; i < N * K; ++i) hB[i] = h16((rand() % 2001 - 1000) * 0.001f, fB[i]);
        for (int m = 0; m < 256; ++m) for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += double(fA[m * K + k]) * fB[n * K + k];
            ref[m * N + n] = float(s);
        }
        uint16_t *dA, *dB; float* dD;
        CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, hD.size() * 4));
        CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0xff, hD.size() * 4));
        const int kblocks = (K + 63) / 64;
        const int smem = kblocks * (128 + N / 2) * 128;
        CK(cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe2_kernel<<<2, 128, smem>>>(dA, dB, dD, N, K);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
        double err = 0;
        for (size_t i = 0; i < hD.size(); ++i) err = fmax(err, fabs(double(hD[i]) - ref[i]));
        const bool ok = err < 1e-3;
        all = all && ok;
        printf("cta_group::2 probe  M=256 N=%3d K=%3d (B split %d + %d rows)  max|err| = %.3e  %s\n", N, K, N / 2, N / 2, err,
               ok ? "OK" : "FAIL");
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    printf(all ? "PROBE2 PASSED\n" : "PROBE2 FAILED\n");
    return all ? 0 : 1;
}
