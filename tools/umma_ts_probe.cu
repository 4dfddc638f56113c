// Hardware probe for the tcgen05 forms the split-precision attention kernel (csrc/mwa_sp.cu) relies on, beyond what
// tools/umma_probe.cu covers (run on the B200 box under `timeout`; exit code 0 = all pass):
//   ts     A operand read from TMEM (packed fp16, lane = row, 32-bit column = two consecutive k), B K-major from smem
//   mn     B operand MN-major (stored [k][n], n contiguous, SWIZZLE_128B) -- the V operand of P V
//   tsmn   both at once (the actual P V configuration: P in TMEM, V [key][d] in smem)
//   acc    a second, narrower MMA accumulating onto a column sub-range of the first one's accumulator, with the B rows
//          taken at a row offset inside the same slab (the V-column correction passes of the QKV GEMM)
//   koff   A and B taken from the two 64-byte halves of one [128 x 64] K-major SW128 buffer (Q | K side by side), plus
//          the zero-block trick: D = [A_top; 0] B0^T + [0; A_bot] B1^T
//   half   P V with one accumulator half per window: A from TMEM, B MN-major [key][w0 n | w1 n] (n = N / 2 columns per
//          window inside one 128-byte row), two lane-masked MMAs per k step, the second with its B start address advanced
//          by n * 2 bytes INSIDE the swizzle atom: D[0:64] = A[0:64] B[:, 0:n], D[64:128] = A[64:128] B[:, n:2n]
//   war    an MMA that reads its A operand from TMEM columns which the NEXT MMA (issued right behind it) overwrites
//          as its accumulator: the first result must be unaffected (in-order execution of the tensor pipe)
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

constexpr uint32_t kBMajorMN = 1u << 16;

enum Mode { TS = 0, MN = 1, TSMN = 2, ACC = 3, KOFF = 4, WAR = 5, HALF = 6 };

// A [128][K] fp16 row-major; B: K-major modes [N][K], MN-major modes [K][N]; D [128][N] fp32
__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, float* __restrict__ D, int N, int K, int mode,
             int reps, const float* __restrict__ Ref, int* __restrict__ nbad) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sA = smem;                       // up to 2 x [128 x 128 B]
    uint8_t* sB = smem + 2 * 16384;           // up to 32 KB
    uint8_t* sZ = smem + 4 * 16384;           // 8 KB of zeros
    for (int i = tid; i < (4 * 16384 + 8192) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const bool a_tmem = (mode == TS || mode == TSMN || mode == WAR || mode == HALF);
    const bool b_mn = (mode == MN || mode == TSMN || mode == HALF);
    if (!a_tmem && mode != KOFF)
        for (int i = tid; i < 128 * K; i += 128) {
            int r = i / K, k = i % K;
            *reinterpret_cast<uint16_t*>(sA + (k / 64) * 16384 + sw128_offset(r, k % 64)) = A[i];
        }
    if (mode == KOFF) {
        // layout [W0: rows 0-63][Z: 64 zero rows][W1: rows 64-127]; A in columns 0..31, B in columns 32..63 (K = 32)
        for (int i = tid; i < 128 * K; i += 128) {
            int r = i / K, k = i % K;
            uint8_t* blk = smem + (r < 64 ? 0 : 16384);
            *reinterpret_cast<uint16_t*>(blk + sw128_offset(r % 64, k)) = A[i];
        }
        for (int i = tid; i < N * 2 * K; i += 128) {          // B: [2N][K]: rows 0..N-1 pair with A rows 0-63, N..2N-1 with 64-127
            int r = i / K, k = i % K;
            uint8_t* blk = smem + (r < N ? 0 : 16384);
            *reinterpret_cast<uint16_t*>(blk + sw128_offset(r % N, 32 + k)) = B[i];
        }
    } else if (b_mn) {
        for (int i = tid; i < K * N; i += 128) {              // B given as [K][N]; row k = 128 bytes (N <= 64)
            int k = i / N, n = i % N;
            *reinterpret_cast<uint16_t*>(sB + sw128_offset(k, n)) = B[i];
        }
    } else {
        for (int i = tid; i < N * K; i += 128) {
            int r = i / K, k = i % K;
            *reinterpret_cast<uint16_t*>(sB + (k / 64) * (N * 128) + sw128_offset(r, k % 64)) = B[i];
        }
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tA = 256, tD = 0, tD2 = 256;     // WAR: the second MMA's accumulator covers the A columns
    if (a_tmem) {
        // row = tid; column c of the A region holds k = 2c (low half) and 2c + 1 (high half)
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t v[8];
            for (int j = 0; j < 8; ++j) {
                const int k = 2 * (c0 + j);
                v[j] = uint32_t(A[tid * K + k]) | (uint32_t(A[tid * K + k + 1]) << 16);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tm + lane_addr + tA + c0),
                         "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                         : "memory");
        }
        tmem_wait_st();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }
    uint32_t phase = 0;
    for (int rep = 0; rep < reps; ++rep) {
        if (warp == 0 && elect_one()) {
            const int ksteps = K / 16;
            if (mode == TS || mode == MN || mode == TSMN || mode == WAR) {
                const uint32_t idesc = umma_idesc(kFmtF16, kFmtF16, 128, N) | (b_mn ? kBMajorMN : 0u);
                for (int ks = 0; ks < ksteps; ++ks) {
                    uint64_t bd;
                    if (b_mn) bd = umma_desc_k_sw128(smem_u32(sB) + ks * 2048);          // 16 k-rows of 128 B
                    else bd = umma_desc_k_sw128(smem_u32(sB + (ks / 4) * N * 128) + (ks % 4) * 32);
                    if (a_tmem) umma_f16_ts(tm + tD, tm + tA + ks * 8, bd, idesc, ks > 0);
                    else umma_f16_ss(tm + tD, umma_desc_k_sw128(smem_u32(sA + (ks / 4) * 16384) + (ks % 4) * 32), bd, idesc, ks > 0);
                }
                if (mode == WAR) {
                    // second MMA: accumulator over the columns that hold A (and beyond); operands from smem (zeros)
                    const uint32_t idesc2 = umma_idesc(kFmtF16, kFmtF16, 128, 128);
                    umma_f16_ss(tm + tD2, umma_desc_k_sw128(smem_u32(sA)), umma_desc_k_sw128(smem_u32(sA)), idesc2, 0);   // sA is all zeros in this mode
                }
            } else if (mode == HALF) {
                const int n = N / 2;
                const uint32_t idesc = umma_idesc(kFmtF16, kFmtF16, 128, n) | kBMajorMN;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t b0 = smem_u32(sB) + ks * 2048;
                    umma_f16_ts_lanes(tm + tD, tm + tA + ks * 8, umma_desc_k_sw128(b0), idesc, ks > 0, 0u, 0u, ~0u, ~0u);
                    umma_f16_ts_lanes(tm + tD, tm + tA + ks * 8, umma_desc_k_sw128(b0 + n * 2), idesc, ks > 0, ~0u, ~0u, 0u, 0u);
                }
            } else if (mode == ACC) {
                // first N columns from all B rows; then columns [N-32, N) get a second helping from B rows [N-32, N)
                const uint32_t idesc = umma_idesc(kFmtF16, kFmtF16, 128, N), idesc32 = umma_idesc(kFmtF16, kFmtF16, 128, 32);
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t ad = umma_desc_k_sw128(smem_u32(sA + (ks / 4) * 16384) + (ks % 4) * 32);
                    const uint32_t bbase = smem_u32(sB + (ks / 4) * N * 128) + (ks % 4) * 32;
                    umma_f16_ss(tm + tD, ad, umma_desc_k_sw128(bbase), idesc, ks > 0);
                    umma_f16_ss(tm + tD + (N - 32), ad, umma_desc_k_sw128(bbase + (N - 32) * 128), idesc32, 1);
                }
            } else if (mode == KOFF) {
                const uint32_t idesc = umma_idesc(kFmtF16, kFmtF16, 128, N);
                const uint32_t w0 = smem_u32(smem), z = w0 + 8192, w1 = w0 + 16384;
                for (int ks = 0; ks < ksteps; ++ks) {
                    umma_f16_ss(tm + tD, umma_desc_k_sw128(w0 + ks * 32), umma_desc_k_sw128(w0 + 64 + ks * 32), idesc, ks > 0);
                    umma_f16_ss(tm + tD, umma_desc_k_sw128(z + ks * 32), umma_desc_k_sw128(w1 + 64 + ks * 32), idesc, 1);
                }
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after_sync();
        if (mode == WAR && rep + 1 < reps) {
            // re-write A for the next repetition (the second MMA has zeroed it)
            for (int c0 = 0; c0 < K / 2; c0 += 8) {
                uint32_t v[8];
                for (int j = 0; j < 8; ++j) {
                    const int k = 2 * (c0 + j);
                    v[j] = uint32_t(A[tid * K + k]) | (uint32_t(A[tid * K + k + 1]) << 16);
                }
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tm + lane_addr + tA + c0),
                             "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                             : "memory");
            }
            tmem_wait_st();
            // check this repetition's result on the fly: every repetition exercises the hazard
            for (int c0 = 0; c0 < N; c0 += 8) {
                uint32_t v[8];
                tmem_ld_x8(tm + lane_addr + tD + c0, v);
                tmem_wait_ld();
                for (int j = 0; j < 8; ++j)
                    if (!(fabsf(__uint_as_float(v[j]) - Ref[tid * N + c0 + j]) < 2e-2f)) atomicAdd(nbad, 1);
            }
            tc_fence_before_sync();
            __syncthreads();
            tc_fence_after_sync();
        }
    }
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tmem_ld_x8(tm + lane_addr + tD + c0, v);
        tmem_wait_ld();
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

static uint16_t f2h(float f) {
    __half h = __float2half_rn(f);
    return *reinterpret_cast<uint16_t*>(&h);
}
static float h2f(uint16_t u) {
    __half h = *reinterpret_cast<__half*>(&u);
    return __half2float(h);
}

static int run_case(const char* name, int mode, int N, int K, int reps = 1) {
    const int brows = (mode == KOFF) ? 2 * N : N;
    std::vector<uint16_t> hA(128 * K), hB(brows * K);
    std::vector<float> fA(128 * K), fB(brows * K), ref(128 * N), hD(128 * N);
    srand(99 + N * 7 + K + mode);
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = f2h((rand() % 2001 - 1000) / 500.f); fA[i] = h2f(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = f2h((rand() % 2001 - 1000) / 500.f); fB[i] = h2f(hB[i]); }
    const bool b_mn = (mode == MN || mode == TSMN || mode == HALF);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double acc = 0;
            if (mode == HALF && n >= N / 2) { ref[m * N + n] = 0.f; continue; }
            for (int k = 0; k < K; ++k) {
                double b;
                if (mode == HALF) b = fB[k * N + (m < 64 ? 0 : N / 2) + n];
                else if (mode == KOFF) b = fB[((m < 64 ? 0 : N) + n) * K + k];
                else b = b_mn ? fB[k * N + n] : fB[n * K + k];
                acc += double(fA[m * K + k]) * b;
            }
            if (mode == ACC && n >= N - 32) acc *= 2.0;
            ref[m * N + n] = float(acc);
        }
    uint16_t *dA, *dB;
    float *dD, *dRef;
    int* dBad;
    CK(cudaMalloc(&dRef, ref.size() * 4));
    CK(cudaMalloc(&dBad, 4));
    CK(cudaMemset(dBad, 0, 4));
    CK(cudaMemcpy(dRef, ref.data(), ref.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, hD.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, hD.size() * 4));
    const int smem = 4 * 16384 + 8192;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, mode, reps, dRef, dBad);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (size_t i = 0; i < hD.size(); ++i) {
        if (mode == HALF && int(i % N) >= N / 2) continue;
        double e = fabs(double(hD[i]) - ref[i]);
        if (!(e <= maxerr)) maxerr = e;
    }
    int nbad = 0;
    CK(cudaMemcpy(&nbad, dBad, 4, cudaMemcpyDeviceToHost));
    const bool ok = maxerr < 2e-2 && nbad == 0;
    if (nbad) printf("   %d bad values in earlier repetitions\n", nbad);
    printf("probe %-5s N=%3d K=%3d reps=%4d  max|err|=%.3e  %s\n", name, N, K, reps, maxerr, ok ? "OK" : "FAIL");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return ok ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run_case("ts", TS, 64, 32);
    bad += run_case("ts", TS, 96, 32);
    bad += run_case("ts", TS, 192, 64);
    bad += run_case("mn", MN, 48, 64);
    bad += run_case("mn", MN, 64, 64);
    bad += run_case("mn", MN, 16, 64);
    bad += run_case("tsmn", TSMN, 48, 64);
    bad += run_case("tsmn", TSMN, 64, 64);
    bad += run_case("tsmn", TSMN, 16, 128);
    bad += run_case("half", HALF, 64, 64);
    bad += run_case("half", HALF, 48, 64);
    bad += run_case("half", HALF, 32, 64);
    bad += run_case("acc", ACC, 80, 64);
    bad += run_case("acc", ACC, 96, 128);
    bad += run_case("koff", KOFF, 64, 32);
    bad += run_case("war", WAR, 48, 64, 1);
    bad += run_case("war", WAR, 48, 64, 2000);
    bad += run_case("war", WAR, 64, 32, 2000);
    printf(bad ? "PROBE FAILED (%d cases)\n" : "PROBE PASSED\n", bad);
    return bad ? 1 : 0;
}
