// Hardware probe for the halo-tile trick of the 3x3 convolutions (run on the B200 box under `timeout`):
// the A operand of a tcgen05.mma is read out of a LARGER K-major SWIZZLE_128B tile (rows of 128 B written linearly, chunk
// index XORed with bits [7,10) of the row's shared-memory address -- what a tiled TMA box with 128-byte rows produces)
// through a descriptor whose start address is shifted by a whole number of rows (not a multiple of 8) and whose
// stride-byte-offset is 10 rows (1280 B) instead of 1024:   A[m] = tile[start + (m / 8) * sbo_rows + m % 8].
// Variants of the descriptor's base_offset field are tried; prints which combinations reproduce the host reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
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

constexpr int kRows = 18 * 10 + 8;      // halo tile of a 16 x 8 pixel tile (+ slack)
constexpr int kN = 32;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(base_offset & 7u) << 49;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// T: [kRows][64] fp16 halo tile, B: [kN][64] fp16, D: [128][kN] fp32
__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint16_t* __restrict__ T, const uint16_t* __restrict__ B, float* __restrict__ D, int start_row,
             int sbo_rows, int base_offset_mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sT = smem;                         // kRows x 128 B, linear rows, address-swizzled
    uint8_t* sB = smem + 24 * 1024;             // canonical [kN rows x 128 B]
    for (int i = tid; i < kRows * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        const uint32_t off = r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + (k & 7) * 2;
        *reinterpret_cast<uint16_t*>(sT + off) = T[i];
    }
    for (int i = tid; i < kN * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        *reinterpret_cast<uint16_t*>(sB + sw128_offset(r, k)) = B[i];
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<32>(&tmem_base_s);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = umma_idesc(kFmtF16, kFmtF16, 128, kN);
        const uint32_t a_addr = smem_u32(sT) + start_row * 128;
        const uint32_t bo = base_offset_mode == 0 ? 0u : base_offset_mode == 1 ? uint32_t(start_row & 7) : uint32_t((8 - (start_row & 7)) & 7);
        for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = desc_sw128(a_addr + ks * 32, sbo_rows * 128, bo);
            const uint64_t bd = umma_desc_k_sw128(smem_u32(sB) + ks * 32);
            umma_f16_ss(tmem_d, ad, bd, idesc, ks > 0);
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < kN; c0 += 8) {
        uint32_t v[8];
        tmem_ld_x8(tmem_d + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_wait_ld();
        for (int j = 0; j < 8; ++j) D[tid * kN + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<32>(tmem_d);
}

int main() {
    std::vector<uint16_t> hT(kRows * 64), hB(kN * 64);
    std::vector<float> fT(kRows * 64), fB(kN * 64), hD(128 * kN);
    srand(99);
    for (size_t i = 0; i < hT.size(); ++i) {
        __half h = __float2half_rn((rand() % 2001 - 1000) / 500.f);
        hT[i] = *reinterpret_cast<uint16_t*>(&h);
        fT[i] = __half2float(h);
    }
    for (size_t i = 0; i < hB.size(); ++i) {
        __half h = __float2half_rn((rand() % 2001 - 1000) / 500.f);
        hB[i] = *reinterpret_cast<uint16_t*>(&h);
        fB[i] = __half2float(h);
    }
    uint16_t *dT, *dB;
    float* dD;
    CK(cudaMalloc(&dT, hT.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, hD.size() * 4));
    CK(cudaMemcpy(dT, hT.data(), hT.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    const int smem = 24 * 1024 + kN * 128;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int pass[3] = {0, 0, 0}, total = 0;
    const int sbos[] = {8, 10, 16, 18};
    for (int sbo : sbos)
        for (int start = 0; start < 24; ++start) {
            if (start + 15 * sbo + 8 > kRows) continue;
            ++total;
            for (int mode = 0; mode < 3; ++mode) {
                CK(cudaMemset(dD, 0xff, hD.size() * 4));
                probe_kernel<<<1, 128, smem>>>(dT, dB, dD, start, sbo, mode);
                CK(cudaGetLastError());
                CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
                double maxerr = 0;
                for (int m = 0; m < 128; ++m) {
                    const int r = start + (m / 8) * sbo + m % 8;
                    for (int n = 0; n < kN; ++n) {
                        double acc = 0;
                        for (int k = 0; k < 64; ++k) acc += double(fT[r * 64 + k]) * fB[n * 64 + k];
                        const double e = fabs(double(hD[m * kN + n]) - acc);
                        if (!(e <= maxerr)) maxerr = e;
                    }
                }
                const bool ok = maxerr < 1e-2;
                pass[mode] += ok;
                printf("sbo_rows=%2d start=%2d base_offset_mode=%d  max|err|=%.3e %s\n", sbo, start, mode, maxerr, ok ? "OK" : "FAIL");
            }
        }
    printf("cases %d: base_offset 0 -> %d ok, (start&7) -> %d ok, (-start&7) -> %d ok\n", total, pass[0], pass[1], pass[2]);
    return 0;
}
