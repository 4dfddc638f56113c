// Pieces shared by the two tcgen05 attention kernels (mwa_tc.cu: phase-serial v1, mwa_ws.cu: warp-specialised v2):
// tile / head-group geometry, the layout of the tcgen05 section of the parameter block, the keep-flag scan and the
// compaction kernel, and the warp-level tensor-core primitives of the per-window attention core.
// Reference semantics: layers/masked_win_attention.py:6-47 (partition, keep predicate), :96-131, :169-251.
#pragma once
#include "common.cuh"
#include "params.cuh"
#include "status.cuh"

namespace b200 {
namespace {

constexpr int kTileM = 128;
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kNegMask = -100.0f;                 // layers/masked_win_attention.py:214

struct Geom {
    int B, H, W, shift, nwx, nwy, channels_last;
};

// ------------------------------------------------------------------------------------------------ configuration
template <int C_, int HEADS_, int WS_>
struct Cfg {
    static constexpr int C = C_, HEADS = HEADS_, WS = WS_;
    static constexpr int D = C / HEADS;
    static constexpr int DPAD = (D + 15) / 16 * 16;
    static constexpr int HPG = 64 / DPAD;                 // heads per group
    static constexpr int NG = HEADS / HPG;                // head groups
    static constexpr int NTOK = WS * WS;                  // tokens per window
    static constexpr int WPT = kTileM / NTOK;             // windows per tile
    static constexpr int KB = (C + 63) / 64;              // K blocks of the x operand
    static constexpr int KSTEPS = C / 16;
    static constexpr int NCHUNK = C / 8;                  // 16-byte fp16 chunks per token row
    static constexpr int NCOLG = C / 16;                  // 16-column groups of the projection output
    static constexpr int TBL = (2 * WS - 1) * (2 * WS - 1);
    static constexpr int NQ = HPG * D;                    // un-padded q (k, v) columns of a head group
    static constexpr int NQKV = (3 * NQ + 15) / 16 * 16;  // rows of a QKV weight slab = columns of D_qkv
    static constexpr uint32_t kQkvSlabBytes = NQKV * 128;
    static constexpr uint32_t kProjSlabBytes = C * 128;
    static constexpr uint32_t kGroupBytes = KB * kQkvSlabBytes + kProjSlabBytes;
    static_assert(NQ % 8 == 0 && kQkvSlabBytes % 1024 == 0 && kProjSlabBytes % 1024 == 0, "slab geometry");
    // QKV slab ring: a whole group (KB slabs) when it fits the 227 KB budget, else 2 slots
    static constexpr int kQSlots = (3 * kQkvSlabBytes + kProjSlabBytes <= 80 * 1024) ? 3 : 2;
    static_assert(C % 16 == 0 && HEADS % HPG == 0 && DPAD * HPG == 64, "unsupported head geometry");
    static_assert(kTileM % NTOK == 0 && NTOK % 16 == 0, "unsupported window size");
    // warp-level attention tasks per head group: (window of the tile) x (head of the group) x (16-row block)
    static constexpr int RB = NTOK / 16;
    static constexpr int TASKS = WPT * HPG * RB;
    static_assert(TASKS % kWarps == 0, "tasks must tile the 16 warps");
    // shared memory map (offsets from a 1024-aligned base)
    static constexpr uint32_t oX = 0;                                   // KB x [128 x 64] fp16
    static constexpr uint32_t oQ = oX + KB * 16384;
    static constexpr uint32_t oK = oQ + 16384;
    static constexpr uint32_t oV = oK + 16384;                          // [128 keys x 64] like Q / K (not transposed)
    static constexpr uint32_t oO = oV + 16384;                          // 2 x [128 x 64] head outputs (A operand of proj)
    static constexpr uint32_t oRing = oO + 32768;                        // kQSlots QKV slabs, then 1 projection slab
    static constexpr uint32_t oRingP = oRing + kQSlots * kQkvSlabBytes;
    static constexpr uint32_t oTbl = oRingP + kProjSlabBytes;           // fp32 [HEADS][TBL]
    static constexpr uint32_t oBqkv = oTbl + ((HEADS * TBL * 4 + 15) / 16) * 16;   // fp32 [NG][NQKV]
    static constexpr uint32_t oBproj = oBqkv + NG * NQKV * 4;
    static constexpr uint32_t oBars = (oBproj + C * 4 + 15) / 16 * 16;
    static constexpr uint32_t oTmem = oBars + 16 * 8;
    static constexpr uint32_t oTotal = oTmem + 16;
    static_assert(oTotal <= 227 * 1024, "shared memory budget");
    // fp32 [C][128] output staging of the NCHW epilogue, over operand buffers that are dead (and fully rewritten
    // by the next tile) at that point
    static constexpr uint32_t oStage = (C * 512 <= 3 * 16384) ? oQ : oX;
    static_assert(oStage + C * 512 <= oO, "output staging must fit the dead operand buffers");
    // TMEM columns
    static constexpr uint32_t tA = 0;        // D_qkv of the current head group, NQKV <= 192 columns
    static constexpr uint32_t tP = 256;      // projection accumulator, C columns
};

// layout of the tcgen05 section of the parameter block (offsets from MwaParamLayout::img_wqkv)
template <class CF>
struct TcParams {
    static constexpr int64_t img = 0;                                          // NG x kGroupBytes
    static constexpr int64_t bq = img + int64_t(CF::NG) * CF::kGroupBytes;     // fp32 [NG][192] padded order
    // proj.bias + Wproj * b_v: the v bias adds the same vector to every row of P V (rows of P sum to 1), so it can be
    // folded into the projection bias; the k bias shifts every logit of a row by the same amount and drops out of the
    // softmax altogether.  The warp-specialised kernel therefore only adds the q bias when it drains D_qkv.
    static constexpr int64_t bpf = bq + CF::NG * CF::NQKV * 4;                  // fp32 [C]
    static constexpr int64_t total = bpf + CF::C * 4;
};

// ------------------------------------------------------------------------------------------------ prepare
template <class CF>
__global__ void mwa_tc_prepare_kernel(const float* __restrict__ qkv_w, const float* __restrict__ qkv_b,
                                      const float* __restrict__ proj_w, const float* __restrict__ proj_b, float scale,
                                      uint8_t* __restrict__ out) {
    constexpr int C = CF::C, D = CF::D, DPAD = CF::DPAD, HPG = CF::HPG;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // qkv slabs: row n of group g = part (q|k|v) * NQ + head-in-group * D + c   (no padding inside the slab; rows
    // 3*NQ .. NQKV-1 stay zero); q rows and the q bias are pre-multiplied by the softmax scale
    constexpr int NQ = CF::NQ, NQKV = CF::NQKV;
    for (int e = tid; e < CF::NG * 3 * NQ * C; e += nth) {
        const int g = e / (3 * NQ * C), n = (e / C) % (3 * NQ), k = e % C;
        const int part = n / NQ, hh = (n % NQ) / D, c = (n % NQ) % D;
        float v = qkv_w[int64_t(part * C + (g * HPG + hh) * D + c) * C + k];
        if (part == 0) v *= scale;
        const int64_t off = int64_t(g) * CF::kGroupBytes + int64_t(k / 64) * CF::kQkvSlabBytes + sw128_offset(n, k % 64);
        *reinterpret_cast<__half*>(out + TcParams<CF>::img + off) = __float2half_rn(v);
    }
    // projection slabs: rows = output channel, K = this group's 64 head-padded O columns
    for (int e = tid; e < CF::NG * C * 64; e += nth) {
        const int g = e / (C * 64), n = (e / 64) % C, kk = e % 64;
        const int hh = kk / DPAD, c = kk % DPAD;
        const float v = (c < D) ? proj_w[int64_t(n) * C + (g * HPG + hh) * D + c] : 0.f;
        const int64_t off = int64_t(g) * CF::kGroupBytes + int64_t(CF::KB) * CF::kQkvSlabBytes + sw128_offset(n, kk);
        *reinterpret_cast<__half*>(out + TcParams<CF>::img + off) = __float2half_rn(v);
    }
    for (int e = tid; e < CF::NG * NQKV; e += nth) {
        const int g = e / NQKV, n = e % NQKV;
        float v = 0.f;
        if (n < 3 * NQ && qkv_b != nullptr) {
            const int part = n / NQ, hh = (n % NQ) / D, c = (n % NQ) % D;
            v = qkv_b[part * C + (g * HPG + hh) * D + c];
            if (part == 0) v *= scale;
        }
        reinterpret_cast<float*>(out + TcParams<CF>::bq)[e] = v;
    }
    for (int o = tid; o < C; o += nth) {
        float v = proj_b[o];
        if (qkv_b != nullptr)
            for (int c = 0; c < C; ++c) v = fmaf(proj_w[int64_t(o) * C + c], qkv_b[2 * C + c], v);
        reinterpret_cast<float*>(out + TcParams<CF>::bpf)[o] = v;
    }
}

__global__ void zero16_kernel(uint4* p, int64_t n16) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n16; i += int64_t(gridDim.x) * blockDim.x)
        p[i] = make_uint4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------ scan + compaction
struct ScanWs {           // workspace layout
    int64_t count, flags, list, total;
    __host__ __device__ explicit ScanWs(int64_t nwin) {
        count = 0;
        flags = 16;
        list = align_up(flags + nwin, 16);
        total = align_up(list + 4 * (nwin + 16), 256);
    }
};

__device__ __forceinline__ void window_coords(const Geom& g, int win, int& b, int& wy, int& wx) {
    b = win / (g.nwy * g.nwx);
    const int r = win - b * g.nwy * g.nwx;
    wy = r / g.nwx;
    wx = r - wy * g.nwx;
}
// token t of window (wy, wx): original (un-shifted) pixel
template <int WS>
__device__ __forceinline__ void token_pixel(const Geom& g, int wy, int wx, int t, int& y, int& x) {
    y = wy * WS + t / WS + g.shift;
    if (y >= g.H) y -= g.H;
    x = wx * WS + t % WS + g.shift;
    if (x >= g.W) x -= g.W;
}

// one warp per window: keep = (sum alpha != 0); dropped windows are copied through (the block is the identity
// there).  NCHW copy: lane = 4 consecutive tokens (VEC-wide pieces, as in the main kernel), 8 channels in flight.
// copy_dropped == 0: flags only (the caller has already copied x -> out wholesale).
template <int WS, int VEC>
__global__ void __launch_bounds__(256)
mwa_scan_kernel(const float* __restrict__ x, const float* __restrict__ alpha, float* __restrict__ out, Geom g, int C,
                int nwin, uint8_t* __restrict__ flags, int copy_dropped = 1) {
    constexpr int NTOK = WS * WS;
    const int win = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (win >= nwin) return;
    int b, wy, wx;
    window_coords(g, win, b, wy, wx);
    float a = 0.f;
    for (int t = lane; t < NTOK; t += 32) {
        int y, xx;
        token_pixel<WS>(g, wy, wx, t, y, xx);
        a += __ldg(alpha + (int64_t(b) * g.H + y) * g.W + xx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    const bool keep = a != 0.f;
    if (lane == 0) flags[win] = keep;
    if (keep || !copy_dropped) return;
    const int64_t hw = int64_t(g.H) * g.W;
    if (!g.channels_last) {
        constexpr int GROUPS = NTOK / 4;                  // token groups of 4 (16 for 8x8, 4 for 4x4 windows)
        constexpr int CPI = 32 / GROUPS;                  // channels covered by one warp iteration (2 / 8)
        const int grp = lane % GROUPS, csub = lane / GROUPS;
        const int tok0 = grp * 4;
        int py = wy * WS + tok0 / WS + g.shift;
        if (py >= g.H) py -= g.H;
        int64_t off[4 / VEC];
#pragma unroll
        for (int pc = 0; pc < 4 / VEC; ++pc) {
            int px = wx * WS + tok0 % WS + g.shift + pc * VEC;
            if (px >= g.W) px -= g.W;
            off[pc] = int64_t(b) * C * hw + int64_t(py) * g.W + px;
        }
        constexpr int U = 4;                              // iterations in flight
        for (int c0 = csub; c0 < C; c0 += CPI * U) {
            float v[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u * CPI;
                if (c < C) {
#pragma unroll
                    for (int pc = 0; pc < 4 / VEC; ++pc) {
                        const float* src = x + int64_t(c) * hw + off[pc];
                        if constexpr (VEC == 4) {
                            const float4 t4 = __ldg(reinterpret_cast<const float4*>(src));
                            v[u][0] = t4.x; v[u][1] = t4.y; v[u][2] = t4.z; v[u][3] = t4.w;
                        } else if constexpr (VEC == 2) {
                            const float2 t2 = __ldg(reinterpret_cast<const float2*>(src));
                            v[u][2 * pc] = t2.x; v[u][2 * pc + 1] = t2.y;
                        } else {
                            v[u][pc] = __ldg(src);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u * CPI;
                if (c < C) {
#pragma unroll
                    for (int pc = 0; pc < 4 / VEC; ++pc) {
                        float* dst = out + int64_t(c) * hw + off[pc];
                        if constexpr (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                        else if constexpr (VEC == 2) *reinterpret_cast<float2*>(dst) = make_float2(v[u][2 * pc], v[u][2 * pc + 1]);
                        else *dst = v[u][pc];
                    }
                }
            }
        }
    } else {
        for (int t = 0; t < NTOK; ++t) {                  // NHWC: a token is C contiguous floats
            int y, xx;
            token_pixel<WS>(g, wy, wx, t, y, xx);
            const int64_t o = ((int64_t(b) * g.H + y) * g.W + xx) * C;
            for (int c = lane * 4; c < C; c += 128)
                *reinterpret_cast<float4*>(out + o + c) = __ldg(reinterpret_cast<const float4*>(x + o + c));
        }
    }
}

// single block: ordered list of kept windows (flags == nullptr: every window kept)
__global__ void __launch_bounds__(1024)
mwa_compact_kernel(const uint8_t* __restrict__ flags, int nwin, int32_t* __restrict__ list,
                   int32_t* __restrict__ count) {
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    const int per = (nwin + 1023) / 1024;
    const int beg = tid * per, end = min(beg + per, nwin);
    int n = 0;
    for (int i = beg; i < end; ++i) n += flags ? flags[i] : 1;
    part[tid] = n;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {            // Hillis-Steele inclusive scan
        const int v = (tid >= o) ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int pos = part[tid] - n;
    for (int i = beg; i < end; ++i)
        if (!flags || flags[i]) list[pos++] = i;
    if (tid == 1023) *count = part[1023];
}

// ------------------------------------------------------------------------------------------------ main kernel
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ float ex2(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// VEC: widest vector (floats) that tiles a window row in memory: 4 when shift % 4 == 0 and W % 4 == 0, else 2 / 1
// ---- warp-level tensor-core primitives for the per-window attention core (legacy HMMA path: tiles of 16x8x16 keep
//      the whole softmax of a (window, head, 16-row block) inside one warp's registers -- no TMEM round trip, no
//      CTA barrier, no single-thread MMA issue for these tiny contractions)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// byte offset of 16-byte chunk `chunk` of row `row` in a [128 rows x 64 fp16] K-major SWIZZLE_128B buffer
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t chunk) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}

}  // namespace
}  // namespace b200
