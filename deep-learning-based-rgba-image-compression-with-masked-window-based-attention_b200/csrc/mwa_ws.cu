// Fused masked window attention forward, warp-specialised and software-pipelined (tcgen05 + TMEM), sm_100a only.
// Reference semantics: layers/masked_win_attention.py:169-251 (block) and :96-131 (window attention).
//
// Same three launches as mwa_tc.cu (scan -> compact -> persistent main kernel, one CTA per SM, tile = 128 tokens =
// 2 kept 8x8 windows / 8 kept 4x4 windows), same parameter images.  The main kernel is re-organised around roles
// that run concurrently behind mbarrier pipelines instead of CTA-wide phases:
//
//   warp 0        MMA issuer.  Stream of head groups G = tile * NG + g:  QKV(G): D_qkv[G % 2] = X * Wqkv_g^T,
//                 PROJ(G - 2): D_proj += O_g * Wproj_g^T.  Runs up to two groups ahead of the attention warps.
//   warp 1        weight feeder: bulk copies (TMA engine) of the fp16 weight slabs from L2 into a 3-slot ring
//                 (+ 1 projection slot), in exactly the order the issuer consumes them.
//   warps 4-11    x producer + epilogue (PE): gather x of tile i+2 into registers (fp32 -> packed fp16) while tile i
//                 is computed, store it as the swizzled A operand as soon as the QKV MMAs of tile i+1 ... i.e. the
//                 previous user of the X buffer ... have completed; then the epilogue of tile i
//                 (TMEM -> + bias + residual -> out), thread = token.
//   warps 12-19   attention (A): per head group drain D_qkv (+ bias -> fp16 Q / K / V in smem), then the per-window
//                 core on warp-level HMMA tiles with the softmax in registers (bias comes in as the accumulator
//                 initialiser through immediate-offset LDS, the SW-MSA region mask only on windows that touch the
//                 wrapped border), O_g -> smem as the projection's A operand.
//
// Registers are re-balanced with setmaxnreg (issuer / feeder 32, PE 96, A 128).
// TMEM: two D_qkv buffers + the projection accumulator.  Shared memory: X, Q/K/V, 1-2 O buffers, weight ring.
// Operand precision: fp16 x fp16 -> fp32 accumulate; softmax, bias, residual in fp32.
#include <cstring>
#include <cuda.h>          // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint)
#include "mwa_tc_shared.cuh"

// tuning knobs (defaults = the measured best; tools/build_variants.py builds side-by-side variants)
#ifndef MWA_WS_LB
#define MWA_WS_LB 4
#endif
#ifndef MWA_WS_TMA_DEBUG
#define MWA_WS_TMA_DEBUG 0      // 1: stage but do not issue the reduce-adds (bring-up aid)
#endif
// 8 x 8 windows: epilogue through cp.reduce.async.bulk.tensor (TMA reduce-add of [16 ch][8][8] boxes staged in shared
// memory) instead of red.global.  Correct (the GPU suite passes with it) but measured SLOWER: 0.464 vs 0.425 ms for 6144
// windows -- the reduce-add of 32-byte rows at a 16-byte offset runs at ~7 B/clk per SM (13.6 k cycles per tile against
// 9.6 k for the LSU reductions) and its L2 traffic slows the gather of the next tile (14.5 k -> 19.9 k cycles).  Kept
// behind this switch, default off, as the scaffolding (tensor map, staging ring, rotating issuer warps) of the TMA gather.
#ifndef MWA_WS_TMA_OUT
#define MWA_WS_TMA_OUT 0
#endif
// 8 x 8 windows: gather x with cp.async.bulk.tensor boxes [16 ch][8][8] issued by the two otherwise idle control warps (one
// ring of three 4 KB slots per window slot of the tile, paid for by the third weight slab); the producer warps read the
// staged fp32 values from shared memory instead of gathering them from global memory through the LSU.  Tiles with a window
// on the wrapped image border keep the LSU gather (whole tile: the choice must be warp-uniform in the producer warps).
// Measured A/B on one box (tools/build_variants.py): 0.415 vs 0.435 ms for 6144 kept windows, 0.353 vs 0.378 ms at 50 %
// kept, attention share of the bench step 0.912 vs 0.961 ms; +4 % only at 25 % kept.  Needs MWA_WS_SKEW (border tiles spread
// over the CTAs): without it every fourth CTA owns all border tiles and sets the kernel time (0.466 ms).
#ifndef MWA_WS_TMA_IN
#define MWA_WS_TMA_IN 1
#endif
#ifndef MWA_WS_SLOTS
#define MWA_WS_SLOTS ((MWA_WS_TMA_OUT || MWA_WS_TMA_IN) ? 2 : 3)   // 2 slabs in flight measured as fast as 3 (0.425 vs 0.425 ms) and free 18 KB
#endif
#ifndef MWA_WS_OBUFS_MAX
#define MWA_WS_OBUFS_MAX 2
#endif
#ifndef MWA_WS_SKEW
#define MWA_WS_SKEW 1u          // 0: tile t -> CTA t % grid (every round the same)
#endif
#ifndef MWA_WS_MASK_PREFETCH
#define MWA_WS_MASK_PREFETCH 0
#endif


namespace b200 {
namespace {

// Register budget: the file is 4 x 16 K registers, one quarter per SM sub-partition, and a CTA's pool is what its launch
// bound allocates: 5 warps per sub-partition x 32 x 96 = 15,360 of 16,384 (the next multiple of 8 above 96 does not fit).
// setmaxnreg only moves registers inside that pool: 32 (control) + 2 x 96 (producer) + 2 x 128 (attention) per
// sub-partition uses it up exactly.  A 21st "donor" warp does not help: it lands as a 6th warp on one sub-partition and
// ptxas lowers the launch allocation to 80 (checked), shrinking the pool.
constexpr int kWsWarps = 20;
constexpr int kWsThreads = kWsWarps * 32;
constexpr int kMmaWarp = 0, kWgtWarp = 1, kAllocWarp = 2;
constexpr int kPeWarp0 = 4, kNumPe = 8;
constexpr int kAtWarp0 = 12, kNumAt = 8;
constexpr int kPeThreads = kNumPe * 32, kAtThreads = kNumAt * 32;
constexpr int kRegsCtl = 32, kRegsPe = 96, kRegsAt = 128;

template <class CF>
struct WsMap {
    static constexpr int kSlots = MWA_WS_SLOTS;                           // QKV weight slabs in flight
    static constexpr uint32_t kDqStride = (CF::NQKV + 31) / 32 * 32;      // TMEM columns per D_qkv buffer
    static constexpr int kDqBufs = (2 * kDqStride + CF::C <= 512) ? 2 : 1;
    static constexpr uint32_t tDq = 0;
    static constexpr uint32_t tP = kDqBufs * kDqStride;                   // projection accumulator, C columns
    static_assert(tP + CF::C <= 512, "TMEM budget");
    // shared memory map (offsets from a 1024-aligned base)
    static constexpr uint32_t oX = 0;                                     // KB x [128 x 64] fp16
    static constexpr uint32_t oQ = oX + CF::KB * 16384;
    static constexpr uint32_t oK = oQ + 16384;
    static constexpr uint32_t oV = oK + 16384;
    static constexpr uint32_t oO = oV + 16384;                            // kOBufs x [128 x 64]
    static constexpr uint32_t kFixed = CF::KB * 16384 + 3 * 16384 + kSlots * CF::kQkvSlabBytes + CF::kProjSlabBytes;
    static constexpr uint32_t kSmall = ((CF::HEADS * CF::TBL * 4 + 15) / 16) * 16 + CF::NG * CF::NQKV * 4 + CF::C * 4 + 512 +
                                       ((CF::WS == 8 && (MWA_WS_TMA_OUT != 0 || MWA_WS_TMA_IN != 0)) ? 3 * 8192 : 0);
    static constexpr int kOBufs = (MWA_WS_OBUFS_MAX >= 3 && kFixed + 3 * 16384 + kSmall <= 227 * 1024) ? 3
                                  : (kFixed + 2 * 16384 + kSmall <= 227 * 1024) ? 2 : 1;
    static constexpr uint32_t oRing = oO + kOBufs * 16384;
    static constexpr uint32_t oRingP = oRing + kSlots * CF::kQkvSlabBytes;
    // epilogue staging for the TMA reduce-add: kStageBufs x [2 windows][16 ch][8][8] fp32 (box order of the tensor map)
    static constexpr bool kTmaOut = (CF::WS == 8) && (MWA_WS_TMA_OUT != 0);
    static constexpr int kStageBufs = 3;
    static constexpr uint32_t kStageBytes = 2 * 16 * 64 * 4;
    static constexpr uint32_t oStage = oRingP + CF::kProjSlabBytes;
    static constexpr bool kTmaIn = (CF::WS == 8) && (MWA_WS_TMA_IN != 0);      // staging = 3 slots x [2 windows] x 4 KB boxes
    static_assert(!(kTmaOut && kTmaIn), "one user of the staging ring at a time");
    static constexpr uint32_t oTbl = oStage + ((kTmaOut || kTmaIn) ? kStageBufs * kStageBytes : 0);   // fp32 [HEADS][TBL]
    static constexpr uint32_t oBqkv = oTbl + ((CF::HEADS * CF::TBL * 4 + 15) / 16) * 16;   // fp32 [NG][NQKV]
    static constexpr uint32_t oBproj = oBqkv + CF::NG * CF::NQKV * 4;
    static constexpr uint32_t oBars = (oBproj + CF::C * 4 + 15) / 16 * 16;
    static constexpr uint32_t oTmem = oBars + 40 * 8;
    static constexpr uint32_t oTotal = oTmem + 16;
    static_assert(oTotal <= 227 * 1024, "shared memory budget");
    // barrier indices
    static constexpr int bXFull = 0, bXEmpty = 1, bPjFull = 2, bPjEmpty = 3, bPFull = 4, bPEmpty = 5;
    static constexpr int bWFull = 6, bWEmpty = bWFull + kSlots;
    static constexpr int bDqFull = bWEmpty + kSlots, bDqEmpty = bDqFull + 2, bOFull = bDqEmpty + 2, bOEmpty = bOFull + 3;
    static constexpr int bSFull = bOEmpty + 3, bSEmpty = bSFull + 6;       // staging ring of the TMA gather: [window][slot]
    static_assert(bSEmpty + 6 <= 40, "barrier slots");
};

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void named_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
// out = x wholesale -- unless every window is kept (then the main kernel pre-stores the residual itself, see load_x).
// The decision is taken on the device from the compacted count, so the forward stays free of host synchronisation.
__global__ void __launch_bounds__(256)
residual_copy_kernel(const float4* __restrict__ x, float4* __restrict__ out, int64_t n4,
                     const int32_t* __restrict__ count_p, int nwin) {
    if (*count_p == nwin) return;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    constexpr int U = 8;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldcs(x + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) out[i + u * stride] = v[u];
    }
    for (; i < n4; i += stride) out[i] = __ldcs(x + i);
}
__device__ __forceinline__ float rcp_approx(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// Tile row order.  A warp's 32 TMEM lanes / tile rows hold, for 8x8 windows, image rows 2q and 2q+1 of BOTH windows of
// the tile (for 4x4 windows: image row q of all 8 windows) instead of 4 rows of one window: when the two windows are
// horizontal neighbours (the common case in the kept-window list) a warp-wide access of one channel then touches
// 2 x 64 contiguous bytes instead of 4 x 32 -- the NCHW gather is bound by the number of 128-byte lines a request
// touches.  The tensor-core side does not care about the row order; the attention warps address rows through
// ldmatrix's per-lane row pointers.
template <int WS>
__device__ __forceinline__ int tile_row(int w, int t) {          // window slot w, token t -> tile row
    if constexpr (WS == 8) return (t >> 4) * 32 + ((t >> 3) & 1) * 16 + w * 8 + (t & 7);
    else return (t >> 2) * 32 + w * 4 + (t & 3);
}
template <int WS>
__device__ __forceinline__ void tile_row_inv(int r, int& w, int& t) {
    if constexpr (WS == 8) {
        w = (r >> 3) & 1;
        t = (r >> 5) * 16 + ((r >> 4) & 1) * 8 + (r & 7);
    } else {
        w = (r >> 2) & 7;
        t = (r >> 5) * 4 + (r & 3);
    }
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(b0));
}

// Per-lane shared-memory offsets of the operand rows an attention task touches (relative to the Q / K / V / O buffer
// and to window slot 0); for 8x8 windows every other address of the task is base + w * 1024 + a compile-time term,
// optionally XORed with a compile-time chunk bit -- nothing is recomputed (or spilled) per task.
template <class CF>
struct TaskAddr {
    uint32_t q, k, v, o;
    int rbk, cb;
    __device__ __forceinline__ TaskAddr(int rbk_, int hh, int lane) : rbk(rbk_), cb(hh * (CF::DPAD / 8)) {
        if constexpr (CF::WS == 8) {
            const uint32_t m = lane & 7;
            q = (rbk * 4 + ((lane >> 3) & 1) * 2) * 1024 + m * 128 + (((cb + (lane >> 4)) ^ m) << 4);
            k = m * 128 + (((cb + (lane >> 3)) ^ m) << 4);
            v = ((lane >> 3) & 1) * 2048 + m * 128 + (((cb + (lane >> 4)) ^ m) << 4);
            const uint32_t mo = lane >> 2;
            o = rbk * 4096 + mo * 128 + ((cb ^ mo) << 4) + (lane & 3) * 4;
        } else {
            q = k = v = o = 0;
        }
    }
};

// One attention task: the 16 query tokens rbk*16 .. rbk*16+15 of window slot w against the window's NTOK keys, for the
// head at 16-byte-chunk offset cb of the group buffers.  S = Q K^T + bias (accumulator initialiser, `bias` holds this
// lane's values: index (h2 - n + NT - 1) * 2 + e) + region mask; softmax in registers; O = P V; O / rowsum -> fp16.
template <class CF>
__device__ __forceinline__ void attention_task_ws(uint32_t sQ, uint32_t sK, uint32_t sV, uint32_t sO,
                                                  const TaskAddr<CF>& ad, const float (&bias)[2 * (CF::NTOK / 8) + 2],
                                                  int w, bool has_mask, const uint32_t (&rowmask)[2], int lane) {
    constexpr int WS = CF::WS, NTOK = CF::NTOK, DPAD = CF::DPAD, D = CF::D;
    constexpr int KS = DPAD / 16, NT = NTOK / 8, PK = NTOK / 16, ON = (D + 7) / 8;
    constexpr bool kHalfStep = (D % 16) != 0 && (D % 16) <= 8;      // last k step of Q K^T only needs 8 columns
    constexpr bool kFast = (WS == 8) && (DPAD == 32);               // closed-form addresses (see TaskAddr)
    const int rbk = ad.rbk, cb = ad.cb;
    const uint32_t wo = w * 1024;
    uint32_t qa[KS][4];
    if constexpr (kFast) {
        ldmatrix_x4(qa[0], sQ + ad.q + wo);
        ldmatrix_x4(qa[1], (sQ + ad.q + wo) ^ 32u);
    } else {
        const uint32_t row = tile_row<WS>(w, rbk * 16 + (lane & 15));
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) ldmatrix_x4(qa[ks], sQ + swz(row, cb + 2 * ks + (lane >> 4)));
    }
    float sc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        // accumulator initialised with the relative-position bias (layers/masked_win_attention.py:109-112)
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
            for (int e = 0; e < 2; ++e) sc[n][2 * h2 + e] = bias[(h2 - n + NT - 1) * 2 + e];
    }
    if constexpr (KS == 2) {
#pragma unroll
        for (int n = 0; n < NT; ++n) {        // one x4 load: both k steps of the 8 keys of n-tile n
            uint32_t kb[4];
            if constexpr (kFast) ldmatrix_x4(kb, sK + ad.k + wo + ((n >> 1) * 4 + (n & 1) * 2) * 1024);
            else ldmatrix_x4(kb, sK + swz(tile_row<WS>(w, 8 * n + (lane & 7)), cb + (lane >> 3)));
            const uint32_t k0[2] = {kb[0], kb[1]}, k1[2] = {kb[2], kb[3]};
            mma16816(sc[n], qa[0], k0);
            if constexpr (kHalfStep) mma1688(sc[n], qa[1][0], qa[1][1], kb[2]);
            else mma16816(sc[n], qa[1], k1);
        }
    } else {
#pragma unroll
        for (int n = 0; n < NT; n += 2) {     // one x4 load: the single k step of n-tiles n and n+1
            uint32_t kb[4];
            ldmatrix_x4(kb, sK + swz(tile_row<WS>(w, 8 * (n + (lane >> 4)) + (lane & 7)), cb + ((lane >> 3) & 1)));
            const uint32_t k0[2] = {kb[0], kb[1]}, k1[2] = {kb[2], kb[3]};
            mma16816(sc[n], qa[0], k0);
            mma16816(sc[n + 1], qa[0], k1);
        }
    }
    if (has_mask) {      // warp-uniform: SW-MSA region mask, only for windows on the wrapped border (:194-216)
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 8 * n + 2 * (lane & 3) + e, yj = j / WS, xj = j % WS;
                if (((rowmask[0] >> yj) | (rowmask[0] >> (8 + xj))) & 1u) sc[n][e] += kNegMask;
                if (((rowmask[1] >> yj) | (rowmask[1] >> (8 + xj))) & 1u) sc[n][2 + e] += kNegMask;
            }
    }
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        ma = fmaxf(ma, fmaxf(sc[n][0], sc[n][1]));
        mb = fmaxf(mb, fmaxf(sc[n][2], sc[n][3]));
    }
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1));
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    const float mla = ma * kLog2e, mlb = mb * kLog2e;
    float suma = 0.f, sumb = 0.f;
    uint32_t pa[PK][4];                                   // un-normalised probabilities as A fragments of P V
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const float p0 = ex2(fmaf(sc[n][0], kLog2e, -mla)), p1 = ex2(fmaf(sc[n][1], kLog2e, -mla));
        const float p2 = ex2(fmaf(sc[n][2], kLog2e, -mlb)), p3 = ex2(fmaf(sc[n][3], kLog2e, -mlb));
        suma += p0 + p1;
        sumb += p2 + p3;
        pa[n >> 1][(n & 1) * 2 + 0] = pack_f16x2(p0, p1);
        pa[n >> 1][(n & 1) * 2 + 1] = pack_f16x2(p2, p3);
    }
    suma += __shfl_xor_sync(0xffffffffu, suma, 1);
    suma += __shfl_xor_sync(0xffffffffu, suma, 2);
    sumb += __shfl_xor_sync(0xffffffffu, sumb, 1);
    sumb += __shfl_xor_sync(0xffffffffu, sumb, 2);
    float oc[ON][4];
#pragma unroll
    for (int nt = 0; nt < ON; ++nt) oc[nt][0] = oc[nt][1] = oc[nt][2] = oc[nt][3] = 0.f;
#pragma unroll
    for (int j = 0; j < PK; ++j) {
        // keys 16j .. 16j+15: lanes 0-15 address the two 8-key halves of chunk nt, lanes 16-31 those of chunk nt + 1
        uint32_t vaddr;
        if constexpr (kFast) vaddr = sV + ad.v + wo + j * 4096;
        else vaddr = 0;
        const uint32_t row = kFast ? 0u : uint32_t(tile_row<WS>(w, 16 * j + (lane & 7) + 8 * ((lane >> 3) & 1)));
#pragma unroll
        for (int nt = 0; nt < ON; nt += 2) {
            if (nt + 1 < ON) {
                uint32_t vb[4];
                if constexpr (kFast) ldmatrix_x4_trans(vb, vaddr ^ (uint32_t(nt) << 4));
                else ldmatrix_x4_trans(vb, sV + swz(row, cb + nt + (lane >> 4)));
                const uint32_t v0[2] = {vb[0], vb[1]}, v1[2] = {vb[2], vb[3]};
                mma16816(oc[nt], pa[j], v0);
                mma16816(oc[nt + 1], pa[j], v1);
            } else {
                uint32_t vb[2];
                if constexpr (kFast) ldmatrix_x2_trans(vb, vaddr ^ (uint32_t(nt) << 4));
                else ldmatrix_x2_trans(vb, sV + swz(row, cb + nt));
                mma16816(oc[nt], pa[j], vb);
            }
        }
    }
    const float inva = rcp_approx(suma), invb = rcp_approx(sumb);
    if constexpr (kFast) {
        const uint32_t oa = sO + ad.o + wo;
#pragma unroll
        for (int nt = 0; nt < ON; ++nt) {
            st_shared_b32(oa ^ (uint32_t(nt) << 4), pack_f16x2(oc[nt][0] * inva, oc[nt][1] * inva));
            st_shared_b32((oa + 2048) ^ (uint32_t(nt) << 4), pack_f16x2(oc[nt][2] * invb, oc[nt][3] * invb));
        }
    } else {
        const uint32_t ra = tile_row<WS>(w, rbk * 16 + (lane >> 2)), rb = tile_row<WS>(w, rbk * 16 + 8 + (lane >> 2));
#pragma unroll
        for (int nt = 0; nt < ON; ++nt) {
            const uint32_t coff = (lane & 3) * 4;         // byte offset of the column pair inside its 16-byte chunk
            st_shared_b32(sO + swz(ra, cb + nt) + coff, pack_f16x2(oc[nt][0] * inva, oc[nt][1] * inva));
            st_shared_b32(sO + swz(rb, cb + nt) + coff, pack_f16x2(oc[nt][2] * invb, oc[nt][3] * invb));
        }
    }
}

// D_qkv columns [part * NQ + COL0, + NCOLS) of this thread's row + bias -> fp16 -> the head-padded chunks of heads
// [H_LO, H_HI) of the destination operand buffer (pad columns written as zeros, all-pad chunks left untouched:
// they were zeroed once at kernel start).
template <class CF, int COL0, int NCOLS, int H_LO, int H_HI, bool kBias>
__device__ __forceinline__ void drain_part(uint32_t taddr, const float* bias, uint32_t rowaddr, int r) {
    constexpr int D = CF::D, DPAD = CF::DPAD;
    static_assert(COL0 % 8 == 0 && NCOLS % 8 == 0 && H_LO * D >= COL0 && H_HI * D <= COL0 + NCOLS, "drain split");
    float val[NCOLS];
    {
        uint32_t acc[NCOLS / 8][8];
#pragma unroll
        for (int cc = 0; cc < NCOLS / 8; ++cc) tmem_ld_x8(taddr + COL0 + cc * 8, acc[cc]);
        tmem_wait_ld();
#pragma unroll
        for (int cc = 0; cc < NCOLS / 8; ++cc)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                val[cc * 8 + j] = kBias ? __uint_as_float(acc[cc][j]) + bias[COL0 + cc * 8 + j] : __uint_as_float(acc[cc][j]);
    }
#pragma unroll
    for (int ch = H_LO * DPAD / 8; ch < H_HI * DPAD / 8; ++ch) {
        if ((8 * ch) % DPAD >= D) continue;                // chunk made of pad columns only
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = 8 * ch + j;
            v[j] = ((col % DPAD) < D) ? val[(col / DPAD) * D + (col % DPAD) - COL0] : 0.f;
        }
        st_shared_v4(rowaddr + ((ch ^ (r & 7)) << 4), pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]),
                     pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
    }
}

template <class CF, bool kTiming>
__global__ void __launch_bounds__(kWsThreads, 1)
mwa_ws_kernel(const float* __restrict__ x, float* __restrict__ out, const uint8_t* __restrict__ blk,
              const uint8_t* __restrict__ tcp, const int32_t* __restrict__ list, const int32_t* __restrict__ count_p,
              Geom geo, unsigned long long* __restrict__ timing, const __grid_constant__ CUtensorMap out_map,
              const __grid_constant__ CUtensorMap x_map) {
    using MP = WsMap<CF>;
    constexpr int C = CF::C, WS = CF::WS, NTOK = CF::NTOK, DPAD = CF::DPAD, HPG = CF::HPG, NG = CF::NG;
    constexpr int LOOK = MP::kDqBufs;                   // QKV groups issued ahead of the projection stream
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MP::oBars);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + MP::oTmem);
    float* s_tbl = reinterpret_cast<float*>(smem + MP::oTbl);
    float* s_bqkv = reinterpret_cast<float*>(smem + MP::oBqkv);
    float* s_bproj = reinterpret_cast<float*>(smem + MP::oBproj);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const MwaParamLayout L(C, CF::HEADS, WS);

    // ---- one-time setup (all warps)
    if (tid == 0) {
        if (sb & 1023u) __trap();
        mbar_init(bars + MP::bXFull, kPeThreads);
        mbar_init(bars + MP::bXEmpty, 1);
        mbar_init(bars + MP::bPjFull, 1);
        mbar_init(bars + MP::bPjEmpty, kPeThreads);
        mbar_init(bars + MP::bPFull, 1);
        mbar_init(bars + MP::bPEmpty, 1);
        for (int i = 0; i < MP::kSlots; ++i) {
            mbar_init(bars + MP::bWFull + i, 1);
            mbar_init(bars + MP::bWEmpty + i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bars + MP::bDqFull + i, 1);
            mbar_init(bars + MP::bDqEmpty + i, kAtThreads);
        }
        for (int i = 0; i < 3; ++i) {
            mbar_init(bars + MP::bOFull + i, kAtThreads);
            mbar_init(bars + MP::bOEmpty + i, 1);
        }
        for (int i = 0; i < 6; ++i) {
            mbar_init(bars + MP::bSFull + i, 1);
            mbar_init(bars + MP::bSEmpty + i, 128);      // the 128 producer threads of a window slot (both channel halves)
        }
        fence_mbar_init();
    }
    if (warp == kAllocWarp) tmem_alloc<512>(tmem_ptr);
    {
        const float* bq = reinterpret_cast<const float*>(tcp + TcParams<CF>::bq);
        for (int i = tid; i < NG * CF::NQKV; i += kWsThreads) s_bqkv[i] = bq[i];
        const float* bp = reinterpret_cast<const float*>(tcp + TcParams<CF>::bpf);      // proj.bias + Wproj * b_v
        for (int i = tid; i < C; i += kWsThreads) s_bproj[i] = bp[i];
        // operand buffers: padding columns must read as exact zeros for the whole kernel
        for (int i = tid; i < (MP::oRing - MP::oX) / 16; i += kWsThreads)
            reinterpret_cast<uint4*>(smem + MP::oX)[i] = make_uint4(0, 0, 0, 0);
        // compact relative-position table s_tbl[h][idx], idx = (yi-yj+WS-1)*(2WS-1) + (xi-xj+WS-1)
        const float* bexp = reinterpret_cast<const float*>(blk + L.bias);
        for (int e = tid; e < CF::HEADS * CF::TBL; e += kWsThreads) {
            const int h = e / CF::TBL, idx = e % CF::TBL;
            const int dy = idx / (2 * WS - 1) - (WS - 1), dx = idx % (2 * WS - 1) - (WS - 1);
            const int yi = dy >= 0 ? dy : 0, yj = dy >= 0 ? 0 : -dy;
            const int xi = dx >= 0 ? dx : 0, xj = dx >= 0 ? 0 : -dx;
            s_tbl[e] = bexp[(int64_t(h) * NTOK + (yi * WS + xi)) * NTOK + (yj * WS + xj)];
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = *tmem_ptr;

    const int count = *count_p;
    const bool prestore = count == geo.B * geo.nwx * geo.nwy;      // every window kept: no residual copy pass ran
    const int num_tiles = (count + CF::WPT - 1) / CF::WPT;
    const int64_t hw = int64_t(geo.H) * geo.W;
    int my_tiles = 0;
    // tile of this CTA in round `it`: the rounds are skewed by one CTA each, so that the tiles with a window on the wrapped
    // image border (every 12th tile of a 24-window row: they would all land on every fourth CTA) spread over all CTAs
    auto tile_of = [&](int it) -> int { return it * int(gridDim.x) + int((blockIdx.x + unsigned(it) * MWA_WS_SKEW) % gridDim.x); };
    {
        const int full = num_tiles / int(gridDim.x), rem = num_tiles - full * int(gridDim.x);
        my_tiles = full + ((int((blockIdx.x + unsigned(full) * MWA_WS_SKEW) % gridDim.x) < rem) ? 1 : 0);
    }
    const int total = my_tiles * NG;                       // head groups this CTA processes
    const uint8_t* wimg = tcp + TcParams<CF>::img;

    // development aid (kTiming instantiation only): lane 0 of the first warp of each role of CTA 0 accumulates
    // clock64() deltas per pipeline stage
    const bool do_time = kTiming && timing != nullptr && blockIdx.x == 0 && lane == 0 &&
                         (warp == kMmaWarp || warp == kPeWarp0 || warp == kAtWarp0);
    const long long t_cta0 = (kTiming && timing != nullptr) ? clock64() : 0;
    // event trace of CTA 0: [1024 + role * 1024 + i] = stage << 48 | cycles since kernel start at the END of the stage.
    // Plain stores only (fire and forget): a read-modify-write per stage costs the single-warp roles ~1k cycles per
    // event and distorts the very pipeline it measures.
    int n_ev = 0;
    auto tick = [&](int slot) {
        if constexpr (kTiming) {
            if (do_time && n_ev < 1024)
                timing[1024 + (slot >> 3) * 1024 + n_ev++] =
                    (static_cast<unsigned long long>(slot) << 48) | static_cast<unsigned long long>(clock64() - t_cta0);
        }
    };

    if (warp < 4) {
        // =========================================================================================== control warps
        reg_dec<kRegsCtl>();
        if (warp == kMmaWarp) {
            constexpr uint32_t idesc_q = umma_idesc(kFmtF16, kFmtF16, kTileM, CF::NQKV);
            constexpr uint32_t idesc_p = umma_idesc(kFmtF16, kFmtF16, kTileM, C);
            uint32_t q_used = 0, p_used = 0;
            for (int G = 0; G < total + LOOK; ++G) {
                if (G < total) {
                    const int it = G / NG, g = G % NG, b = G % MP::kDqBufs;
                    if (g == 0) mbar_wait(bars + MP::bXFull, it & 1);
                    tick(0);                                                     // 0: wait X
                    if (G >= MP::kDqBufs) mbar_wait(bars + MP::bDqEmpty + b, ((G / MP::kDqBufs) - 1) & 1);
                    tick(1);                                                     // 1: wait D_qkv drained
                    tc_fence_after_sync();
#pragma unroll
                    for (int kb = 0; kb < CF::KB; ++kb) {
                        const uint32_t slot = q_used % MP::kSlots;
                        mbar_wait(bars + MP::bWFull + slot, (q_used / MP::kSlots) & 1);
                        tc_fence_after_sync();
                        const uint64_t a0 = umma_desc_k_sw128(sb + MP::oX + kb * 16384);
                        const uint64_t b0 = umma_desc_k_sw128(sb + MP::oRing + slot * CF::kQkvSlabBytes);
                        const int nks = (kb == CF::KB - 1) ? (CF::KSTEPS - 4 * (CF::KB - 1)) : 4;
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                if (ks < nks)
                                    umma_f16_ss(tm + MP::tDq + b * MP::kDqStride, a0 + ks * 2, b0 + ks * 2, idesc_q,
                                                (kb | ks) != 0);
                            umma_commit(bars + MP::bWEmpty + slot);
                        }
                        __syncwarp();
                        ++q_used;
                    }
                    if (elect_one()) {
                        umma_commit(bars + MP::bDqFull + b);
                        if (g == NG - 1) umma_commit(bars + MP::bXEmpty);
                    }
                    __syncwarp();
                    tick(2);                                                     // 2: QKV issue incl. slab waits
                }
                if (G >= LOOK) {
                    const int Gp = G - LOOK, itp = Gp / NG, gp = Gp % NG, ob = Gp % MP::kOBufs;
                    mbar_wait(bars + MP::bOFull + ob, (Gp / MP::kOBufs) & 1);
                    tick(3);                                                     // 3: wait O_g
                    if (gp == 0 && itp > 0) mbar_wait(bars + MP::bPjEmpty, (itp - 1) & 1);
                    tick(4);                                                     // 4: wait projection accumulator free
                    mbar_wait(bars + MP::bPFull, p_used & 1);
                    tc_fence_after_sync();
                    const uint64_t a0 = umma_desc_k_sw128(sb + MP::oO + ob * 16384);
                    const uint64_t b0 = umma_desc_k_sw128(sb + MP::oRingP);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_f16_ss(tm + MP::tP, a0 + ks * 2, b0 + ks * 2, idesc_p, (gp | ks) != 0);
                        umma_commit(bars + MP::bPEmpty);
                        umma_commit(bars + MP::bOEmpty + ob);
                        if (gp == NG - 1) umma_commit(bars + MP::bPjFull);
                    }
                    __syncwarp();
                    ++p_used;
                    tick(5);                                                     // 5: projection issue incl. slab wait
                }
            }
        } else if (warp == kWgtWarp) {
            uint32_t q_issued = 0, p_issued = 0;
            for (int G = 0; G < total + LOOK; ++G) {
                if (G < total) {
                    const int g = G % NG;
#pragma unroll
                    for (int kb = 0; kb < CF::KB; ++kb) {
                        const uint32_t slot = q_issued % MP::kSlots;
                        if (q_issued >= uint32_t(MP::kSlots))
                            mbar_wait(bars + MP::bWEmpty + slot, ((q_issued / MP::kSlots) - 1) & 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(bars + MP::bWFull + slot, CF::kQkvSlabBytes);
                            bulk_g2s(smem + MP::oRing + slot * CF::kQkvSlabBytes,
                                     wimg + int64_t(g) * CF::kGroupBytes + int64_t(kb) * CF::kQkvSlabBytes,
                                     CF::kQkvSlabBytes, bars + MP::bWFull + slot);
                        }
                        __syncwarp();
                        ++q_issued;
                    }
                }
                if (G >= LOOK) {
                    const int gp = (G - LOOK) % NG;
                    if (p_issued >= 1) mbar_wait(bars + MP::bPEmpty, (p_issued - 1) & 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bars + MP::bPFull, CF::kProjSlabBytes);
                        bulk_g2s(smem + MP::oRingP,
                                 wimg + int64_t(gp) * CF::kGroupBytes + int64_t(CF::KB) * CF::kQkvSlabBytes,
                                 CF::kProjSlabBytes, bars + MP::bPFull);
                    }
                    __syncwarp();
                    ++p_issued;
                }
            }
        }
        if constexpr (MP::kTmaIn) {
            if (warp == 2 || warp == 3) {
                // ---- TMA gather issuer of window slot w: boxes of tile i in the order the producer warps consume them
                //      (chunk j of channel half 0, chunk j of half 1, ...), three boxes in flight
                const int w = warp - 2;
                uint32_t n = 0;
                for (int i = 0; i < my_tiles; ++i) {
                    // a tile is staged only if BOTH of its windows exist and neither wraps round the image border (the
                    // choice must be warp-uniform in the producer warps, whose lanes cover both windows)
                    const int lidx0 = tile_of(i) * CF::WPT;
                    if (lidx0 + 1 >= count) continue;
                    int b_, wy_, wx_, bo_, wyo_, wxo_;
                    window_coords(geo, list[lidx0 + w], b_, wy_, wx_);
                    window_coords(geo, list[lidx0 + 1 - w], bo_, wyo_, wxo_);
                    const int x0 = wx_ * WS + geo.shift, y0 = wy_ * WS + geo.shift;
                    if (x0 + WS > geo.W || y0 + WS > geo.H || wxo_ * WS + geo.shift + WS > geo.W ||
                        wyo_ * WS + geo.shift + WS > geo.H)
                        continue;                                           // LSU path for the whole tile
                    for (int j = 0; j < C / 16; ++j, ++n) {
                        const int k = (j & 1) * (C / 32) + (j >> 1);
                        const uint32_t slot = n % 3;
                        if (n >= 3) mbar_wait(bars + MP::bSEmpty + w * 3 + slot, ((n / 3) - 1) & 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(bars + MP::bSFull + w * 3 + slot, 4096);
                            tma_load_4d(smem + MP::oStage + (slot * 2 + w) * 4096, &x_map, x0, y0, k * 16, b_,
                                        bars + MP::bSFull + w * 3 + slot);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp < kAtWarp0) {
        // =========================================================================================== x producer + epilogue
        reg_dec<kRegsPe>();
        const int pw = warp - kPeWarp0, q = pw & 3, half = pw >> 2;
        const int r = q * 32 + lane;                        // token row of the tile == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        int wslot, tok;
        tile_row_inv<WS>(r, wslot, tok);
        constexpr int CPH = CF::NCHUNK / 2;                 // 8-channel chunks per thread (half of the row)
        static_assert(CF::NCHUNK % 2 == 0, "C must be a multiple of 16");
        constexpr int LB = MWA_WS_LB;                       // chunks per load batch (8 * LB loads in flight per thread)
        uint32_t pk[CPH][4];
        auto row_base = [&](int tile, bool& valid) -> int64_t {      // NCHW element offset of (b, c = 0, py, px)
            const int lidx = tile * CF::WPT + wslot;
            valid = lidx < count;
            const int win = list[valid ? lidx : (count - 1)];
            int b, wy, wx, py, px;
            window_coords(geo, win, b, wy, wx);
            token_pixel<WS>(geo, wy, wx, tok, py, px);
            return int64_t(b) * C * hw + int64_t(py) * geo.W + px;
        };
        // x of one tile -> packed fp16 in registers.  The epilogue only ADDS the projection onto `out` with
        // fire-and-forget reductions, so `out` must already hold the residual x.  Two ways (measured):
        //   some window dropped: residual_copy_kernel has copied x -> out wholesale before this kernel, which also
        //                settles the dropped windows; x is read exactly once here.
        //   every window kept (prestore): the fp32 values just loaded go straight to `out`; no copy pass at all.
        //                With dropped neighbours these half-sector stores are slow, hence the switch.  (An L2
        //                evict-last hint on these stores changed nothing measurable and would leave persisting lines
        //                behind for the next kernel, so they are plain stores.)
        uint32_t stage_n = 0;                                // boxes of this thread's window slot issued so far (kTmaIn)
        auto load_x = [&](int tile) {
            bool valid;
            const int64_t off = row_base(tile, valid) + int64_t(half * CPH * 8) * hw;
            const float* p = x + off;
            float* po = out + off;
            if constexpr (MP::kTmaIn) {
                bool staged = tile * CF::WPT + 1 < count;
                if (staged) {
#pragma unroll
                    for (int w2 = 0; w2 < 2; ++w2) {
                        int b_, wy_, wx_;
                        window_coords(geo, list[tile * CF::WPT + w2], b_, wy_, wx_);
                        staged = staged && (wx_ * WS + geo.shift + WS <= geo.W) && (wy_ * WS + geo.shift + WS <= geo.H);
                    }
                }
                if (staged) {
                    // this thread's 96 channels arrive as six [16 ch][64 tok] boxes (every second box of the window's
                    // stream: the other channel half's boxes are interleaved with them)
                    // (every thread of the window observes EVERY box of the stream, also the other half's: a parity wait
                    //  may lag the barrier by at most one phase)
#pragma unroll
                    for (int jj = 0; jj < CPH; ++jj) {
                        const uint32_t nb = stage_n + jj, slot = nb % 3;
                        mbar_wait(bars + MP::bSFull + wslot * 3 + slot, (nb / 3) & 1);
                        if ((jj & 1) != half) {
                            mbar_arrive(bars + MP::bSEmpty + wslot * 3 + slot);
                            continue;
                        }
                        const int i = jj >> 1;
                        const uint32_t base = sb + MP::oStage + (slot * 2 + wslot) * 4096 + tok * 4;
                        float v[16];
#pragma unroll
                        for (int c = 0; c < 16; ++c) v[c] = ld_shared_f32(base + c * 256);
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            if (prestore) *po = v[c];
                            po += hw;
                        }
#pragma unroll
                        for (int jp = 0; jp < 4; ++jp) {
                            pk[2 * i][jp] = pack_f16x2(v[2 * jp], v[2 * jp + 1]);
                            pk[2 * i + 1][jp] = pack_f16x2(v[8 + 2 * jp], v[8 + 2 * jp + 1]);
                        }
                        // Release the slot only after the loaded values have ARRIVED: an mbarrier arrive does not wait for
                        // ld.shared instructions still queued in the LSU (behind global stores they can be hundreds of cycles
                        // late; found with csrc/mwa_sp.cu in round 2, where the TMA overwrote boxes under their readers).
                        // The arrive is predicated on a value computed from all 16 loads (cvt.satfinite never produces the
                        // all-ones pattern, so the minimum of the packed words never equals it and the predicate is always
                        // true -- but the hardware has to wait for the data).
                        {
                            uint32_t all = 0xffffffffu;
#pragma unroll
                            for (int jp = 0; jp < 4; ++jp) all = min(all, min(pk[2 * i][jp], pk[2 * i + 1][jp]));
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.u32 p, %1, 0xffffffff;\n\t"
                                "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}\n" ::"r"(
                                    smem_u32(bars + MP::bSEmpty + wslot * 3 + slot)),
                                "r"(all)
                                : "memory");
                        }
                    }
                    stage_n += C / 16;
                    return;
                }
            }
#pragma unroll
            for (int c0 = 0; c0 < CPH; c0 += LB) {
                float v[LB][8];
#pragma unroll
                for (int i = 0; i < LB; ++i)
                    if (c0 + i < CPH) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[i][j] = valid ? __ldg(p) : 0.f;
                            p += hw;
                        }
                    }
#pragma unroll
                for (int i = 0; i < LB; ++i)
                    if (c0 + i < CPH) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (valid && prestore) *po = v[i][j];
                            po += hw;
                        }
#pragma unroll
                        for (int jp = 0; jp < 4; ++jp) pk[c0 + i][jp] = pack_f16x2(v[i][2 * jp], v[i][2 * jp + 1]);
                    }
            }
        };
        auto store_x = [&]() {
#pragma unroll
            for (int i = 0; i < CPH; ++i) {
                const int ci = half * CPH + i;
                const uint32_t addr = sb + MP::oX + (ci >> 3) * 16384 + (r >> 3) * 1024 + (r & 7) * 128 +
                                      (((ci & 7) ^ (r & 7)) << 4);
                st_shared_v4(addr, pk[i][0], pk[i][1], pk[i][2], pk[i][3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(bars + MP::bXFull);
        };
        if (my_tiles > 0) {
            load_x(blockIdx.x);
            store_x();
            if (my_tiles > 1) load_x(tile_of(1));
        }
        tick(8);                                                                 // 8: PE prologue
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = tile_of(it);
            if (it + 1 < my_tiles) {
                mbar_wait(bars + MP::bXEmpty, it & 1);       // QKV MMAs of tile `it` have consumed X
                tick(9);                                                         // 9: wait X free
                store_x();
                tick(10);                                                        // 10: store X
                if (it + 2 < my_tiles) load_x(tile_of(it + 2));
                tick(11);                                                        // 11: load x (tile + 2)
            }
            // ---- epilogue of tile `it`: out (= x, stored by this very thread when it loaded the tile) += proj + bias;
            //      thread = token, half of the channels
            bool valid;
            float* orow = out + row_base(tile, valid) + int64_t(half * CPH * 8) * hw;
            mbar_wait(bars + MP::bPjFull, it & 1);
            tc_fence_after_sync();
            tick(12);                                                            // 12: wait projection complete
            constexpr int EG = 4;                            // chunks per TMEM round trip
            if constexpr (MP::kTmaOut) {
                // ---- TMA epilogue: projection + bias -> fp32 staging boxes [window][16 ch][8][8] -> reduce-add onto
                //      `out` inside the L2 (cp.reduce.async.bulk.tensor): no global request touches the LSU, the
                //      accumulator is released after ~3 k cycles of shared-memory stores instead of ~10 k of reductions.
                //      A window that hangs over the right / bottom border (cyclic shift): the tensor map clips the
                //      out-of-bound part of its box; the tokens that wrap round to the left / top edge (a few per cent)
                //      go through red.global as before -- negative box coordinates are an illegal instruction for the
                //      reduce (probed), so the wrapped copies cannot be issued as shifted boxes.
                //      Chunk e of the kernel (12 per tile) uses staging buffer e % 3; its two boxes are issued by lane 0 of
                //      producer warps (2e) & 7 and (2e + 1) & 7 (several issuing warps: the TMA rate scales with them,
                //      tools/tma_probe.cu), which also wait for their own reads before the buffer comes round again.
                constexpr int NCK = C / 16;
                int bwin[2] = {0, 0};
                bool bvalid[2] = {false, false};
                if (lane == 0) {
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const int lidx = tile * CF::WPT + w;
                        bvalid[w] = lidx < count;
                        bwin[w] = list[bvalid[w] ? lidx : count - 1];
                    }
                }
                const uint32_t my_off = wslot * (MP::kStageBytes / 2) + (half * 8) * 256 + tok * 4;
                bool wraps = false;                          // this token lies beyond the right / bottom image border
                {
                    const int lidx = tile * CF::WPT + wslot;
                    if (lidx < count) {
                        int b_, wy_, wx_;
                        window_coords(geo, list[lidx], b_, wy_, wx_);
                        wraps = (wx_ * WS + tok % WS + geo.shift >= geo.W) || (wy_ * WS + tok / WS + geo.shift >= geo.H);
                    }
                }
                float* owrap = orow - int64_t(half * CPH * 8) * hw + int64_t(half * 8) * hw;     // channel half * 8 of chunk 0
#pragma unroll
                for (int k0 = 0; k0 < NCK; k0 += EG) {
                    uint32_t acc[EG][8];
#pragma unroll
                    for (int i = 0; i < EG; ++i) tmem_ld_x8(tm + MP::tP + lane_addr + (k0 + i) * 16 + half * 8, acc[i]);
                    tmem_wait_ld();
                    if (k0 + EG >= NCK) {                    // last TMEM read of the tile: hand the accumulator back
                        tc_fence_before_sync();
                        mbar_arrive(bars + MP::bPjEmpty);
                    }
#pragma unroll
                    for (int i = 0; i < EG; ++i) {
                        const int k = k0 + i;
                        const uint32_t e = uint32_t(it) * NCK + k;
                        const uint32_t sbuf = sb + MP::oStage + (e % MP::kStageBufs) * MP::kStageBytes;
                        if (!wraps) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                st_shared_f32(sbuf + my_off + j * 256,
                                              __uint_as_float(acc[i][j]) + s_bproj[k * 16 + half * 8 + j]);
                        } else if (valid) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                red_add_f32(owrap + int64_t(k * 16 + j) * hw,
                                            __uint_as_float(acc[i][j]) + s_bproj[k * 16 + half * 8 + j]);
                        }
                        fence_proxy_async_smem();
                        // the buffer chunk e + 1 will use was read by the boxes of chunk e - 2: their issuers wait here
                        if (lane == 0 && e >= 2 && (pw == int((2 * (e - 2)) & 7) || pw == int((2 * (e - 2) + 1) & 7)))
                            bulk_wait_group_read<0>();
                        named_sync(3, kPeThreads);
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
                            if (MWA_WS_TMA_DEBUG != 1 && lane == 0 && pw == int((2 * e + w) & 7) && bvalid[w]) {
                                int b_, wy_, wx_;
                                window_coords(geo, bwin[w], b_, wy_, wx_);
                                const int x0 = wx_ * WS + geo.shift, y0 = wy_ * WS + geo.shift;
                                const void* src = smem + MP::oStage + (e % MP::kStageBufs) * MP::kStageBytes +
                                                  w * (MP::kStageBytes / 2);
                                tma_reduce_add_4d(&out_map, src, x0, y0, k * 16, b_);
                                bulk_commit_group();
                            }
                        }
                    }
                }
                tick(13);
                continue;
            }
#pragma unroll
            for (int c0 = 0; c0 < CPH; c0 += EG) {
                uint32_t acc[EG][8];
#pragma unroll
                for (int i = 0; i < EG; ++i)
                    if (c0 + i < CPH) tmem_ld_x8(tm + MP::tP + lane_addr + (half * CPH + c0 + i) * 8, acc[i]);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < EG; ++i)
                    if (c0 + i < CPH) {
                        const int cc = (half * CPH + c0 + i) * 8;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (valid) red_add_f32(orow, __uint_as_float(acc[i][j]) + s_bproj[cc + j]);
                            orow += hw;
                        }
                    }
            }
            tc_fence_before_sync();
            mbar_arrive(bars + MP::bPjEmpty);
            tick(13);                                                            // 13: epilogue
        }
        if constexpr (MP::kTmaOut) {
            if (lane == 0) bulk_wait_group<0>();             // this warp's reduce-adds have left shared memory and landed
        }
    } else {
        // =========================================================================================== attention warps
        reg_inc<kRegsAt>();
        const int a = warp - kAtWarp0, q = a & 3, half = a >> 2;
        const int r = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        // task decomposition: a warp keeps one (16-row block, head of the group) and walks over window slots, so the
        // bias values it holds in registers serve all of its tasks of a head group
        constexpr int NSLOT = CF::RB * HPG;                 // (row block, head) combinations
        static_assert(kNumAt % NSLOT == 0 && CF::WPT % (kNumAt / NSLOT) == 0, "task decomposition");
        constexpr int GW = kNumAt / NSLOT;                  // warps sharing a combination (split the window slots)
        constexpr int TPW = CF::WPT / GW;                   // tasks (window slots) per warp and head group
        constexpr int NT = NTOK / 8, NB = 2 * NT + 2;
        constexpr int kRowStep = (8 / WS) * (2 * WS - 1);   // table index step per 8 tokens (rows: +, keys: -)
        static_assert(8 % WS == 0, "window size must divide 8");
        const int rbk = (a % NSLOT) % CF::RB, hh = (a % NSLOT) / CF::RB, w0 = (a / NSLOT) * TPW;
        // this lane's entry of the relative-position table for (h2 = 0, n = 0, e = 0), head 0
        int tbl_idx;
        {
            const int ti = rbk * 16 + (lane >> 2), tj = 2 * (lane & 3);
            tbl_idx = (ti / WS - tj / WS + WS - 1) * (2 * WS - 1) + (ti % WS - tj % WS + WS - 1);
        }
        const uint32_t tbl0 = sb + MP::oTbl + 4 * (tbl_idx + hh * CF::TBL);
        const TaskAddr<CF> taddr(rbk, hh, lane);
        // window ids of this warp's tasks, fetched one tile ahead: the list read (L2 latency on a busy LSU) plus the
        // mask set-up sat on the attention warps' critical path once per tile (~2.5 k cycles, tools/phase_times.py)
        int win_next[TPW];
        auto fetch_wins = [&](int tile) {
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                const int lidx = tile * CF::WPT + w0 + i;
                win_next[i] = (geo.shift > 0) ? __ldg(list + (lidx < count ? lidx : count - 1)) : 0;
            }
        };
        if (my_tiles > 0) fetch_wins(blockIdx.x);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = tile_of(it);
            // SW-MSA region mask bits of this warp's rows (:194-216); zero unless the window touches the wrapped border
            uint32_t rowmask[TPW][2];
            uint32_t mask_any = 0;
            int win_cur[TPW];
#pragma unroll
            for (int i = 0; i < TPW; ++i) win_cur[i] = win_next[i];
            if (MWA_WS_MASK_PREFETCH && it + 1 < my_tiles) fetch_wins(tile_of(it + 1));
#pragma unroll
            for (int i = 0; i < TPW; ++i) {
                rowmask[i][0] = rowmask[i][1] = 0u;
                if (geo.shift > 0) {
                    int tb_, twy, twx;
                    if (!MWA_WS_MASK_PREFETCH) {
                        const int lidx = tile * CF::WPT + w0 + i;
                        win_cur[i] = list[lidx < count ? lidx : count - 1];
                    }
                    window_coords(geo, win_cur[i], tb_, twy, twx);
                    if ((twy == geo.nwy - 1) || (twx == geo.nwx - 1)) {
                        mask_any |= 1u << i;
                        const int ys0 = twy * WS, xs0 = twx * WS;
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const int tk = rbk * 16 + (lane >> 2) + 8 * h2;
                            const int by = (ys0 + tk / WS >= geo.H - WS) + (ys0 + tk / WS >= geo.H - geo.shift);
                            const int bx = (xs0 + tk % WS >= geo.W - WS) + (xs0 + tk % WS >= geo.W - geo.shift);
                            uint32_t mbits = 0;
#pragma unroll
                            for (int j = 0; j < WS; ++j) {
                                const int byj = (ys0 + j >= geo.H - WS) + (ys0 + j >= geo.H - geo.shift);
                                const int bxj = (xs0 + j >= geo.W - WS) + (xs0 + j >= geo.W - geo.shift);
                                mbits |= uint32_t(byj != by) << j;
                                mbits |= uint32_t(bxj != bx) << (8 + j);
                            }
                            rowmask[i][h2] = mbits;
                        }
                    }
                }
            }
            tick(20);                                                            // 20: per-tile set-up (mask bits)
            for (int g = 0; g < NG; ++g) {
                const int G = it * NG + g, b = G % MP::kDqBufs, ob = G % MP::kOBufs;
                mbar_wait(bars + MP::bDqFull + b, (G / MP::kDqBufs) & 1);
                tc_fence_after_sync();
                tick(16);                                                        // 16: wait D_qkv
                {   // ---- drain: half 0: q (all heads, + bias) + k (first half of the heads); half 1: rest of k + v.
                    //      k and v biases are not added here (TcParams<>::bpf explains why that is exact)
                    constexpr int NQ = CF::NQ, D = CF::D;
                    constexpr bool kSplitK = (HPG % 2 == 0) && ((HPG / 2) * D % 8 == 0);
                    constexpr int KLO = kSplitK ? (HPG / 2) * D : NQ;     // k columns loaded by half 0
                    const uint32_t ta = tm + MP::tDq + b * MP::kDqStride + lane_addr;
                    const float* bqkv = s_bqkv + g * CF::NQKV;
                    const uint32_t rowoff = (r >> 3) * 1024 + (r & 7) * 128;
                    if (half == 0) {
                        drain_part<CF, 0, NQ, 0, HPG, true>(ta, bqkv, sb + MP::oQ + rowoff, r);
                        drain_part<CF, 0, KLO, 0, HPG / 2, false>(ta + NQ, bqkv + NQ, sb + MP::oK + rowoff, r);
                    } else {
                        drain_part<CF, kSplitK ? KLO : 0, kSplitK ? NQ - KLO : NQ, HPG / 2, HPG, false>(
                            ta + NQ, bqkv + NQ, sb + MP::oK + rowoff, r);
                        drain_part<CF, 0, NQ, 0, HPG, false>(ta + 2 * NQ, bqkv + 2 * NQ, sb + MP::oV + rowoff, r);
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(bars + MP::bDqEmpty + b);
                named_sync(1, kAtThreads);                   // Q / K / V of the group visible to the 8 warps
                tick(17);                                                        // 17: drain
                if (G >= MP::kOBufs) mbar_wait(bars + MP::bOEmpty + ob, ((G / MP::kOBufs) - 1) & 1);
                tick(18);                                                        // 18: wait O buffer free
                // relative-position bias of this lane for head g * HPG + hh: value (h2 - n, e) at index (h2-n+NT-1)*2+e
                float bias[NB];
#pragma unroll
                for (int dq = 0; dq < NT + 1; ++dq)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        bias[dq * 2 + e] = ld_shared_f32(tbl0 + 4 * (g * HPG * CF::TBL + (dq - (NT - 1)) * kRowStep - e));
#pragma unroll
                for (int i = 0; i < TPW; ++i)
                    attention_task_ws<CF>(sb + MP::oQ, sb + MP::oK, sb + MP::oV, sb + MP::oO + ob * 16384, taddr, bias,
                                          w0 + i, (mask_any >> i) & 1u, rowmask[i], lane);
                fence_proxy_async_smem();                    // O_g is read by the projection MMA (async proxy)
                mbar_arrive(bars + MP::bOFull + ob);
                named_sync(2, kAtThreads);                   // all reads of Q / K / V done before the next drain
                tick(19);                                                        // 19: attention core
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == kAllocWarp) tmem_dealloc<512>(tm);
    if constexpr (kTiming) {      // per-CTA totals: [64 + cta] cycles, [64 + 256 + cta] tiles (buffer of 4096 u64)
        if (timing != nullptr && tid == 0) {
            timing[64 + blockIdx.x] = static_cast<unsigned long long>(clock64() - t_cta0);
            timing[64 + 256 + blockIdx.x] = my_tiles;
        }
    }
}

unsigned long long* g_ws_timing = nullptr;

// 4-D tiled tensor map over an NCHW fp32 tensor (x for the gather, out for the experimental reduce-add epilogue),
// box = [16 ch][8][8]: one 16-channel slice of an 8 x 8 window
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_window_box_map(float* out, int B, int C, int H, int W, CUtensorMap* map) {
    static EncodeTiledFn encode = nullptr;
    if (encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        MWA_TRY_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres), "mwa_forward(tensor map)");
        if (fn == nullptr || qres != cudaDriverEntryPointSuccess) return MWA_ERR_UNSUPPORTED;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t dims[4] = {cuuint64_t(W), cuuint64_t(H), cuuint64_t(C), cuuint64_t(B)};
    const cuuint64_t strides[3] = {cuuint64_t(W) * 4, cuuint64_t(H) * W * 4, cuuint64_t(C) * H * W * 4};
    const cuuint32_t box[4] = {8, 8, 16, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MWA_OK : MWA_ERR_UNSUPPORTED;
}

template <class CF>
int launch_ws(const float* x, const float* alpha, float* out, const void* params, int B, int H, int W, int shift,
              int32_t* kept_count, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    const Geom geo{B, H, W, shift, W / CF::WS, H / CF::WS, 0};
    const int64_t nwin64 = int64_t(B) * geo.nwx * geo.nwy;
    if (nwin64 > 0x3fffffffll) return MWA_ERR_UNSUPPORTED;
    const int nwin = static_cast<int>(nwin64);
    const ScanWs ws(nwin);
    if (!workspace || workspace_bytes < ws.total) return MWA_ERR_WORKSPACE;
    if (!aligned16(workspace)) return MWA_ERR_ALIGNMENT;
    uint8_t* wsp = static_cast<uint8_t*>(workspace);
    int32_t* count = reinterpret_cast<int32_t*>(wsp + ws.count);
    uint8_t* flags = wsp + ws.flags;
    int32_t* list = reinterpret_cast<int32_t*>(wsp + ws.list);
    const uint8_t* blk = static_cast<const uint8_t*>(params);
    const MwaParamLayout L(CF::C, CF::HEADS, CF::WS);
    if (alpha != nullptr) {
        mwa_scan_kernel<CF::WS, 1><<<(nwin + 7) / 8, 256, 0, st>>>(x, alpha, out, geo, CF::C, nwin, flags, 0);
        int rc = check_launch("mwa_forward(scan)");
        if (rc != MWA_OK) return rc;
    }
    mwa_compact_kernel<<<1, 1024, 0, st>>>(alpha ? flags : nullptr, nwin, list, count);
    int rc = check_launch("mwa_forward(compact)");
    if (rc != MWA_OK) return rc;
    // out = x wholesale if any window was dropped (the block is the identity there; kept windows get the projection
    // added on top); the kernel returns at once when every window is kept
    if (alpha != nullptr && out != x) {
        const int64_t n = int64_t(B) * CF::C * H * W;            // multiple of 4: C % 16 == 0
        residual_copy_kernel<<<kNumSMs * 8, 256, 0, st>>>(reinterpret_cast<const float4*>(x),
                                                           reinterpret_cast<float4*>(out), n / 4, count, nwin);
        rc = check_launch("mwa_forward(residual copy)");
        if (rc != MWA_OK) return rc;
    }
    CUtensorMap out_map, x_map;
    memset(&out_map, 0, sizeof(out_map));
    memset(&x_map, 0, sizeof(x_map));
    if constexpr (WsMap<CF>::kTmaOut) {
        rc = make_window_box_map(out, B, CF::C, H, W, &out_map);
        if (rc != MWA_OK) return rc;
    }
    if constexpr (WsMap<CF>::kTmaIn) {
        rc = make_window_box_map(const_cast<float*>(x), B, CF::C, H, W, &x_map);
        if (rc != MWA_OK) return rc;
    }
    const int smem = WsMap<CF>::oTotal;
    const int max_tiles = (nwin + CF::WPT - 1) / CF::WPT;
    const int grid = max_tiles < kNumSMs ? max_tiles : kNumSMs;
    if (g_ws_timing != nullptr) {
        MWA_TRY_CUDA(cudaFuncSetAttribute(mwa_ws_kernel<CF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                     "mwa_forward(ws attr)");
        mwa_ws_kernel<CF, true><<<grid, kWsThreads, smem, st>>>(x, out, blk, blk + L.img_wqkv, list, count, geo,
                                                                g_ws_timing, out_map, x_map);
    } else {
        MWA_TRY_CUDA(cudaFuncSetAttribute(mwa_ws_kernel<CF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem),
                     "mwa_forward(ws attr)");
        mwa_ws_kernel<CF, false><<<grid, kWsThreads, smem, st>>>(x, out, blk, blk + L.img_wqkv, list, count, geo,
                                                                 nullptr, out_map, x_map);
    }
    rc = check_launch("mwa_forward(tcgen05 ws)");
    if (rc != MWA_OK) return rc;
    if (kept_count)
        MWA_TRY_CUDA(cudaMemcpyAsync(kept_count, count, sizeof(int32_t), cudaMemcpyDeviceToDevice, st),
                     "mwa_forward(kept_count)");
    return MWA_OK;
}

using Cfg192h8 = Cfg<192, 8, 8>;
using Cfg192h6 = Cfg<192, 6, 8>;
using Cfg80h8 = Cfg<80, 8, 4>;

}  // namespace

bool mwa_ws_supported(int C, int heads, int ws, int channels_last) {
    if (channels_last) return false;
    return (C == 192 && heads == 8 && ws == 8) || (C == 192 && heads == 6 && ws == 8) || (C == 80 && heads == 8 && ws == 4);
}
void mwa_ws_set_timing_buffer(void* p) { g_ws_timing = static_cast<unsigned long long*>(p); }

int mwa_forward_ws(const float* x, const float* alpha, float* out, const void* params, int B, int C, int H, int W,
                   int heads, int ws, int shift, int32_t* kept_count, void* workspace, int64_t workspace_bytes,
                   cudaStream_t st) {
    if (C == 192 && heads == 8 && ws == 8)
        return launch_ws<Cfg192h8>(x, alpha, out, params, B, H, W, shift, kept_count, workspace, workspace_bytes, st);
    if (C == 192 && heads == 6 && ws == 8)
        return launch_ws<Cfg192h6>(x, alpha, out, params, B, H, W, shift, kept_count, workspace, workspace_bytes, st);
    if (C == 80 && heads == 8 && ws == 4)
        return launch_ws<Cfg80h8>(x, alpha, out, params, B, H, W, shift, kept_count, workspace, workspace_bytes, st);
    return MWA_ERR_UNSUPPORTED;
}

}  // namespace b200
