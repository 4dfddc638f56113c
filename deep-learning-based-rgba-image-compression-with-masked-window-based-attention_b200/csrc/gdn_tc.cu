// GDN / IGDN forward for C == 192 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
// Reference semantics: layers/GDN.py:83-90   y_i = x_i * (beta_i + sum_j gamma_ij x_j^2)^(-/+ 1/2).
//
// Roofline: HBM-bound (1536 B and 73.7 kFLOP per pixel, AI = 48 FLOP/B): the contraction must run on the tensor
// pipe (the fp32 FMA pipe tops out near 12 FLOP/B) and everything else must hide behind the HBM stream.
//
// Design (one persistent CTA per SM, 16 warps; lane 0 of warp 0 doubles as the MMA issuer):
//   * tile = 128 consecutive pixels of one image x 192 channels.  x is read ONCE from HBM with coalesced
//     128-byte warp loads straight into registers (48 per thread: pixel = TMEM lane, 3 x 16 channels).
//     After the x^2 operand has been produced the fp32 values are parked in spare TMEM columns (tcgen05.st;
//     same lane = pixel / column = channel layout as the accumulator) until the epilogue, so the registers are
//     free to receive the NEXT tile while the tensor core works: loads overlap the MMAs, no raw copy in shared
//     memory, HBM traffic = algorithmic bytes.
//   * the tile after next is pulled into L2 with prefetch.global.L2.
//   * precision: x^2 and gamma are split into bf16 hi + lo (16 significant bits each) and the product is taken as
//     hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (relative error ~2^-16 on the norm pool) -- three
//     kind::f16 passes cost 3456 tensor cycles per tile against ~8500 cycles of HBM time per tile.
//   * gamma (hi and lo UMMA B-operand images, 144 KB, prepared by gdn_prepare in the K-major SWIZZLE_128B layout)
//     is bulk-copied (TMA engine, cp.async.bulk) into shared memory once per CTA and reused for every tile.
//   * A operand (x^2 hi / lo) is produced in 64-channel chunks into a 2-slot shared-memory ring
//     (mbarrier full/empty, tcgen05.commit frees a slot); one elected thread issues the 36 MMAs per tile.
//   * epilogue: tcgen05.ld of the thread's 48 accumulators, rsqrt / sqrt, multiply with the register copy of x,
//     coalesced stores.
#include "common.cuh"
#include "params.cuh"
#include "status.cuh"

namespace b200 {
namespace {

constexpr int kC = 192;
constexpr int kTileM = 128;                 // pixels per tile = UMMA M = TMEM lanes
constexpr int kComputeWarps = 16;
constexpr int kThreads = kComputeWarps * 32;   // 4 warps per SM sub-partition -> 128 registers per thread
constexpr int kChunkK = 64;                 // channels per A chunk = one SWIZZLE_128B K block of bf16
constexpr int kChunks = kC / kChunkK;       // 3
constexpr int kSlots = 2;
constexpr uint32_t kGammaBlockBytes = kC * 128;             // one K block of the gamma image: 192 rows x 128 B
constexpr uint32_t kGammaImgBytes = kChunks * kGammaBlockBytes;   // 73,728
constexpr uint32_t kAHalfBytes = kTileM * 128;              // hi (or lo) part of one A chunk: 16 KB
constexpr uint32_t kSlotBytes = 2 * kAHalfBytes;            // 32 KB
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kXColOffset = 256;           // TMEM columns [256,448) hold the fp32 x tile

struct Smem {
    // offsets from the 1024-aligned dynamic smem base
    static constexpr uint32_t gamma_hi = 0;
    static constexpr uint32_t gamma_lo = gamma_hi + kGammaImgBytes;
    static constexpr uint32_t a_ring = gamma_lo + kGammaImgBytes;              // 147,456
    static constexpr uint32_t beta = a_ring + kSlots * kSlotBytes;             // 212,992
    static constexpr uint32_t bars = beta + kC * 4;
    static constexpr uint32_t tmem_ptr = bars + 8 * 8;
    static constexpr uint32_t total = tmem_ptr + 16;
};
static_assert(Smem::total + 1024 <= 227 * 1024, "shared memory budget");

__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Where the result goes when the consumer is a convolution of csrc/conv_tc.cu (gdn_forward_planes): the fp16 hi / lo planes its
// TMA reads, [n_img][ps * ps][H / ps][W / ps][cstride] channels last -- the same values conv_act_split would produce from y.
struct GdnPlanes {
    uint16_t* hi;
    uint16_t* lo;
    int W, hp, wp, sh, cstride;          // sh = ps - 1 (ps is 1 or 2), hp = H / ps, wp = W / ps
};

template <bool kInverse, bool kNHWC, bool kPlanes = false>
__global__ void __launch_bounds__(kThreads, 1)
gdn_tc_kernel(const float* __restrict__ x, float* __restrict__ y, const uint8_t* __restrict__ blk, int64_t hw,
              int64_t tiles_per_img, int64_t num_tiles, const GdnPlanes sp) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
    uint64_t* bar_gamma = bars + 0;
    uint64_t* bar_full = bars + 1;          // [kSlots]  A chunk written by the compute warps
    uint64_t* bar_empty = bars + 3;         // [kSlots]  MMAs reading the slot have completed
    uint64_t* bar_dfull = bars + 5;         // accumulator of the current tile complete
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::tmem_ptr);
    const uint32_t s_beta = sbase + Smem::beta;          // shared-window address: beta is read with ld.shared

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const GdnParamLayout L(kC);

    if (tid == 0) {
        mbar_init(bar_gamma, 1);
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(bar_full + s, kComputeWarps);
            mbar_init(bar_empty + s, 1);
        }
        mbar_init(bar_dfull, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_ptr);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = *tmem_ptr;                 // columns [0,192): accumulator D
    const uint32_t tmem_x = tmem_d + kXColOffset;      // columns [256,448): fp32 copy of the tile's x

    if (tid == 0) {
        // parameters: gamma hi / lo operand images + beta, one bulk copy each (TMA engine)
        mbar_arrive_expect_tx(bar_gamma, 2 * kGammaImgBytes + kC * 4);
        bulk_g2s(smem + Smem::gamma_hi, blk + L.img_hi, kGammaImgBytes, bar_gamma);
        bulk_g2s(smem + Smem::gamma_lo, blk + L.img_lo, kGammaImgBytes, bar_gamma);
        bulk_g2s(smem + Smem::beta, blk + L.beta, kC * 4, bar_gamma);
    }
    {
        // ===================== compute warps =====================
        const int q = warp & 3;               // TMEM lane quadrant this warp may access
        const int cg = warp >> 2;             // 16-channel group inside every 64-channel chunk
        const int p = q * 32 + lane;          // pixel within the tile == TMEM lane == A row
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;

        float xr[kChunks * 16];               // this thread's 48 values of the tile being converted
        const int64_t chunk_stride = int64_t(kChunkK) * hw;
        auto load_tile = [&](int64_t tile) {
            const int64_t img = tile / tiles_per_img;
            const int64_t p0 = (tile - img * tiles_per_img) * kTileM;
            // rows past the end of the image re-read its last pixel (never stored): no predication on the loads
            const int64_t pp = (p0 + p < hw) ? p0 + p : hw - 1;
            if constexpr (kNHWC) {
                // NHWC: a pixel is 192 contiguous floats.  Coalesced mapping, independent of the TMEM lane mapping:
                // for chunk kc, piece idx = tid + 512 j (j < 4) is float4 #(idx & 15) of the chunk's 64 channels of
                // pixel idx >> 4  ->  a warp reads two 256-byte runs per instruction.
                const float* tb = x + (img * hw) * kC;
#pragma unroll
                for (int kc = 0; kc < kChunks; ++kc)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int idx = tid + 512 * j;
                        const int64_t px = p0 + (idx >> 4);
                        const float4 t4 = ld_stream4(tb + (px < hw ? px : hw - 1) * kC + kc * kChunkK + 4 * (idx & 15));
                        xr[kc * 16 + 4 * j + 0] = t4.x;
                        xr[kc * 16 + 4 * j + 1] = t4.y;
                        xr[kc * 16 + 4 * j + 2] = t4.z;
                        xr[kc * 16 + 4 * j + 3] = t4.w;
                    }
            } else {
                const float* pc = x + (img * kC + cg * 16) * hw + pp;
#pragma unroll
                for (int kc = 0; kc < kChunks; ++kc) {
                    const float* pj = pc;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        xr[kc * 16 + j] = ld_stream(pj);
                        pj += hw;
                    }
                    pc += chunk_stride;
                }
            }
        };
        auto prefetch_tile = [&](int64_t nt) {                 // 768 lines of 128 B, 512 threads
            if (nt >= num_tiles) return;
            const int64_t nimg = nt / tiles_per_img;
            const int64_t np0 = (nt - nimg * tiles_per_img) * kTileM;
            if constexpr (kNHWC) {
                const float* nb = x + (nimg * hw + np0) * kC;
                const int64_t lines = ((hw - np0 < kTileM ? hw - np0 : kTileM) * kC * 4 + 127) / 128;
                for (int i = tid; i < lines; i += kComputeWarps * 32) prefetch_l2(nb + int64_t(i) * 32);
            } else {
                const float* nb = x + (nimg * kC) * hw + np0;
                for (int i = tid; i < kC * 4; i += kComputeWarps * 32)
                    if (np0 + (i & 3) * 32 < hw) prefetch_l2(nb + int64_t(i >> 2) * hw + (i & 3) * 32);
            }
        };

        // first tile's HBM loads go out BEFORE waiting for the 147 KB parameter copy: the two latencies overlap
        if (int64_t(blockIdx.x) < num_tiles) load_tile(blockIdx.x);
        prefetch_tile(int64_t(blockIdx.x) + gridDim.x);
        mbar_wait(bar_gamma, 0);              // beta in smem (also orders the first MMA after the gamma copy)

        uint32_t g = 0, it = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int64_t img = tile / tiles_per_img;
            const int64_t p0 = (tile - img * tiles_per_img) * kTileM;
            const bool valid = p0 + p < hw;

            // (1) A operand: x^2 split into bf16 hi + lo, K-major SWIZZLE_128B rows (row = pixel);
            //     the fp32 x values are parked in TMEM until the epilogue so that the registers are free again.
#pragma unroll
            for (int kc = 0; kc < kChunks; ++kc, ++g) {
                const uint32_t slot = g % kSlots, use = g / kSlots;
                if (use > 0) mbar_wait(bar_empty + slot, (use - 1) & 1);
                uint32_t hi[8], lo[8], raw[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float v0 = xr[kc * 16 + 2 * j], v1 = xr[kc * 16 + 2 * j + 1];
                    raw[2 * j] = __float_as_uint(v0);
                    raw[2 * j + 1] = __float_as_uint(v1);
                    const float s0 = v0 * v0, s1 = v1 * v1;
                    hi[j] = pack_bf16x2(s0, s1);
                    const float h0 = __uint_as_float(hi[j] << 16), h1 = __uint_as_float(hi[j] & 0xffff0000u);
                    lo[j] = pack_bf16x2(s0 - h0, s1 - h1);
                }
                // (TMEM columns [256,448) of this thread's lane are its private scratch for the fp32 values)
                tmem_st_x16(tmem_x + lane_addr + kc * kChunkK + cg * 16, raw);
                if constexpr (kNHWC) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {          // piece j: pixel (tid + 512 j) >> 4, channels 4*(idx & 15)..+3
                        const int idx = tid + 512 * j, px = idx >> 4, c4 = idx & 15;
                        const uint32_t a = sbase + Smem::a_ring + slot * kSlotBytes + (px >> 3) * 1024 + (px & 7) * 128 +
                                           (((c4 >> 1) ^ (px & 7)) << 4) + (c4 & 1) * 8;
                        st_shared_v2(a, hi[2 * j], hi[2 * j + 1]);
                        st_shared_v2(a + kAHalfBytes, lo[2 * j], lo[2 * j + 1]);
                    }
                } else {
                    const uint32_t row = sbase + Smem::a_ring + slot * kSlotBytes + (p >> 3) * 1024 + (p & 7) * 128;
                    const uint32_t c0 = ((2 * cg) ^ (p & 7)) * 16, c1 = ((2 * cg + 1) ^ (p & 7)) * 16;
                    st_shared_v4(row + c0, hi[0], hi[1], hi[2], hi[3]);
                    st_shared_v4(row + c1, hi[4], hi[5], hi[6], hi[7]);
                    st_shared_v4(row + kAHalfBytes + c0, lo[0], lo[1], lo[2], lo[3]);
                    st_shared_v4(row + kAHalfBytes + c1, lo[4], lo[5], lo[6], lo[7]);
                }
                fence_proxy_async_smem();            // generic-proxy writes -> visible to the tensor core
                tc_fence_before_sync();              // (also orders this thread's earlier tcgen05.ld of D)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + slot);
                if (warp == 0) {
                    // MMA issue: one thread, as soon as all 16 warps have delivered this chunk
                    if (lane == 0) {
                        mbar_wait(bar_full + slot, use & 1);
                        tc_fence_after_sync();
                        constexpr uint32_t idesc = umma_idesc(kFmtBF16, kFmtBF16, kTileM, kC);
                        const uint32_t a_hi = sbase + Smem::a_ring + slot * kSlotBytes;
                        const uint32_t a_lo = a_hi + kAHalfBytes;
                        const uint32_t b_hi = sbase + Smem::gamma_hi + kc * kGammaBlockBytes;
                        const uint32_t b_lo = sbase + Smem::gamma_lo + kc * kGammaBlockBytes;
#pragma unroll
                        for (int ks = 0; ks < kChunkK / 16; ++ks) {
                            const uint32_t ko = ks * 32;
                            umma_f16_ss(tmem_d, umma_desc_k_sw128(a_hi + ko), umma_desc_k_sw128(b_hi + ko), idesc,
                                        (kc | ks) != 0);
                            umma_f16_ss(tmem_d, umma_desc_k_sw128(a_lo + ko), umma_desc_k_sw128(b_hi + ko), idesc, 1);
                            umma_f16_ss(tmem_d, umma_desc_k_sw128(a_hi + ko), umma_desc_k_sw128(b_lo + ko), idesc, 1);
                        }
                        umma_commit(bar_empty + slot);
                        if (kc == kChunks - 1) umma_commit(bar_dfull);
                    }
                    __syncwarp();
                }
            }
            tmem_wait_st();

            // (2) next tile: loads in flight while the tensor core works on this one; the one after -> L2
            const int64_t nt = tile + gridDim.x;
            if (nt < num_tiles) load_tile(nt);
            prefetch_tile(nt + gridDim.x);

            // (3) epilogue: n = D + beta ; y = x * n^(-/+ 1/2)
            mbar_wait(bar_dfull, it & 1);
            tc_fence_after_sync();
            if constexpr (kNHWC) {
                // NHWC epilogue: the lane = pixel role turns the accumulator into r = n^(-/+ 1/2) and hands it, through
                // a swizzled fp32 staging tile in a (now idle) A-ring slot, to the coalesced role that owns x.
                float* yt = y + (img * hw) * kC;
#pragma unroll
                for (int kc = 0; kc < kChunks; ++kc) {
                    const uint32_t stage = sbase + Smem::a_ring + (kc & 1) * kSlotBytes;   // [128 px][64 ch] fp32
                    uint32_t acc[16], xv[16];
                    const int ch0 = kc * kChunkK + cg * 16;
                    tmem_ld_x16(tmem_d + lane_addr + ch0, acc);
                    tmem_ld_x16(tmem_x + lane_addr + ch0, xv);
                    float bt[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(bt[4 * j]), "=f"(bt[4 * j + 1]), "=f"(bt[4 * j + 2]), "=f"(bt[4 * j + 3])
                                     : "r"(s_beta + 4 * (ch0 + 4 * j)));
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float rr[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float n = __uint_as_float(acc[4 * j + i]) + bt[4 * j + i];
                            if (kInverse) asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rr[i]) : "f"(n));
                            else asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rr[i]) : "f"(n));
                        }
                        st_shared_v4(stage + p * 256 + (((cg * 4 + j) ^ (p & 15)) << 4), __float_as_uint(rr[0]),
                                     __float_as_uint(rr[1]), __float_as_uint(rr[2]), __float_as_uint(rr[3]));
                    }
                    __syncthreads();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int idx = tid + 512 * j, px = idx >> 4, c4 = idx & 15;
                        float4 r4;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(r4.x), "=f"(r4.y), "=f"(r4.z), "=f"(r4.w)
                                     : "r"(stage + px * 256 + ((c4 ^ (px & 15)) << 4)));
                        if (p0 + px < hw)
                            *reinterpret_cast<float4*>(yt + (p0 + px) * kC + kc * kChunkK + 4 * c4) =
                                make_float4(__uint_as_float(xv[4 * j]) * r4.x, __uint_as_float(xv[4 * j + 1]) * r4.y,
                                            __uint_as_float(xv[4 * j + 2]) * r4.z, __uint_as_float(xv[4 * j + 3]) * r4.w);
                    }
                }
                __syncthreads();       // staging lives in the A ring: the next tile's operand stores must wait
            } else {
            float* yc = y + (img * kC + cg * 16) * hw + p0 + p;
            int64_t spo = 0;
            if constexpr (kPlanes) {
                // this pixel's row of the planes: 32 contiguous bytes of hi and of lo per 16-channel group
                const int64_t pp = valid ? p0 + p : 0;
                const int oy = int(pp / sp.W), ox = int(pp - int64_t(oy) * sp.W);
                const int plane = ((oy & sp.sh) << sp.sh) + (ox & sp.sh);
                spo = ((((img << (2 * sp.sh)) + plane) * sp.hp + (oy >> sp.sh)) * sp.wp + (ox >> sp.sh)) * sp.cstride + cg * 16;
            }
#pragma unroll
            for (int kc = 0; kc < kChunks; ++kc) {
                uint32_t acc[16], xv[16];
                const int ch0 = kc * kChunkK + cg * 16;
                tmem_ld_x16(tmem_d + lane_addr + ch0, acc);
                tmem_ld_x16(tmem_x + lane_addr + ch0, xv);
                float bt[16];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(bt[4 * j]), "=f"(bt[4 * j + 1]), "=f"(bt[4 * j + 2]), "=f"(bt[4 * j + 3])
                                     : "r"(s_beta + 4 * (ch0 + 4 * j)));
                tmem_wait_ld();
                if constexpr (kPlanes) {
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float n = __uint_as_float(acc[2 * j + i]) + bt[2 * j + i];
                            float r;
                            if (kInverse) asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n));
                            else asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n));
                            v[i] = __uint_as_float(xv[2 * j + i]) * r;
                        }
                        hi[j] = pack_f16x2(v[0], v[1]);
                        const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[j]));
                        lo[j] = pack_f16x2(v[0] - hf.x, v[1] - hf.y);
                    }
                    if (valid) {
                        uint16_t* ph = sp.hi + spo + kc * kChunkK;
                        uint16_t* pl = sp.lo + spo + kc * kChunkK;
                        st_global_v8(ph, hi);         // one full 32-byte sector per lane and instruction (STG.256)
                        st_global_v8(pl, lo);
                    }
                } else {
                float* yj = yc;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float n = __uint_as_float(acc[j]) + bt[j];
                    float r;
                    if (kInverse) asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n));
                    else asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n));
                    if (valid) *yj = __uint_as_float(xv[j]) * r;
                    yj += hw;
                }
                yc += chunk_stride;
                }
            }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem_d);
}

}  // namespace

bool gdn_tc_supported(int C, int64_t hw, int channels_last) {
    (void)channels_last;
    return C == kC && hw >= 1;
}

template <bool kInverse, bool kNHWC, bool kPlanes = false>
static int launch_gdn_tc(const float* x, float* y, const uint8_t* blk, int64_t hw, int64_t tiles_per_img,
                         int64_t num_tiles, int grid, int smem, cudaStream_t st, const GdnPlanes& sp = GdnPlanes{}) {
    MWA_TRY_CUDA(cudaFuncSetAttribute(gdn_tc_kernel<kInverse, kNHWC, kPlanes>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                 "gdn_forward(tc attr)");
    gdn_tc_kernel<kInverse, kNHWC, kPlanes><<<grid, kThreads, smem, st>>>(x, y, blk, hw, tiles_per_img, num_tiles, sp);
    return check_launch("gdn_forward(tcgen05)");
}

int gdn_forward_tc(const float* x, float* y, const void* params, int64_t n_img, int C, int64_t hw, int inverse,
                   int channels_last, cudaStream_t st) {
    if (!gdn_tc_supported(C, hw, channels_last)) return MWA_ERR_UNSUPPORTED;
    const int64_t tiles_per_img = (hw + kTileM - 1) / kTileM;
    const int64_t num_tiles = n_img * tiles_per_img;
    const int grid = static_cast<int>(num_tiles < kNumSMs ? num_tiles : kNumSMs);
    const int smem = Smem::total + 1024;
    const uint8_t* blk = static_cast<const uint8_t*>(params);
    if (inverse)
        return channels_last ? launch_gdn_tc<true, true>(x, y, blk, hw, tiles_per_img, num_tiles, grid, smem, st)
                             : launch_gdn_tc<true, false>(x, y, blk, hw, tiles_per_img, num_tiles, grid, smem, st);
    return channels_last ? launch_gdn_tc<false, true>(x, y, blk, hw, tiles_per_img, num_tiles, grid, smem, st)
                         : launch_gdn_tc<false, false>(x, y, blk, hw, tiles_per_img, num_tiles, grid, smem, st);
}

// GDN whose consumer is a convolution: the result as that convolution's fp16 hi / lo input planes (no fp32 tensor, no
// conv_act_split launch in between).  NCHW fp32 input, C == 192.
int gdn_forward_planes_tc(const float* x, void* out_hi, void* out_lo, int ps, int cstride, const void* params, int64_t n_img,
                          int C, int H, int W, int inverse, cudaStream_t st) {
    const int64_t hw = int64_t(H) * W;
    if (!gdn_tc_supported(C, hw, 0)) return MWA_ERR_UNSUPPORTED;
    if ((ps != 1 && ps != 2) || H % ps != 0 || W % ps != 0 || cstride < C || cstride % 16 != 0) return MWA_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo)) & 31) return MWA_ERR_ALIGNMENT;
    const int64_t tiles_per_img = (hw + kTileM - 1) / kTileM;
    const int64_t num_tiles = n_img * tiles_per_img;
    const int grid = static_cast<int>(num_tiles < kNumSMs ? num_tiles : kNumSMs);
    const int smem = Smem::total + 1024;
    GdnPlanes sp;
    sp.hi = static_cast<uint16_t*>(out_hi); sp.lo = static_cast<uint16_t*>(out_lo);
    sp.W = W; sp.sh = ps - 1; sp.hp = H / ps; sp.wp = W / ps; sp.cstride = cstride;
    const uint8_t* blk = static_cast<const uint8_t*>(params);
    if (inverse) return launch_gdn_tc<true, false, true>(x, nullptr, blk, hw, tiles_per_img, num_tiles, grid, smem, st, sp);
    return launch_gdn_tc<false, false, true>(x, nullptr, blk, hw, tiles_per_img, num_tiles, grid, smem, st, sp);
}

}  // namespace b200
