// Layout of the kernel-ready parameter blocks written by gdn_prepare / mwa_prepare.
// Host and device agree on the offsets through these structs (all offsets in bytes, 1024-byte aligned
// sections so that UMMA operand images can be bulk-copied straight into swizzled shared memory).
#pragma once
#include <stdint.h>

namespace b200 {

__host__ __device__ inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- GDN
struct GdnParamLayout {
    int C, kblocks;
    int64_t beta;       // fp32 [C]            effective beta
    int64_t gamma;      // fp32 [C][C]         effective gamma, [out i][in j]
    int64_t gammaT;     // fp32 [C][C]         transposed, [in j][out i]
    int64_t mask_beta;  // uint8 [C]           1 where beta_p  >= beta_bound  (LowerBound pass-through, backward)
    int64_t mask_gamma; // uint8 [C][C]        1 where gamma_p >= gamma_bound
    int64_t img_hi;     // bf16 UMMA B-operand image of gamma (hi part): kblocks x [C rows x 64 k] K-major SW128
    int64_t img_lo;     // same for the lo part (gamma - hi)
    int64_t total;
    __host__ __device__ explicit GdnParamLayout(int C_) : C(C_), kblocks((C_ + 63) / 64) {
        int64_t o = 0;
        beta = o;       o = align_up(o + 4ll * C, 1024);
        gamma = o;      o = align_up(o + 4ll * C * C, 1024);
        gammaT = o;     o = align_up(o + 4ll * C * C, 1024);
        mask_beta = o;  o = align_up(o + C, 1024);
        mask_gamma = o; o = align_up(o + 1ll * C * C, 1024);
        const int64_t rows = align_up(C, 8);
        img_hi = o;     o = align_up(o + kblocks * rows * 128, 1024);
        img_lo = o;     o = align_up(o + kblocks * rows * 128, 1024);
        total = o;
    }
    __host__ __device__ int64_t img_bytes() const { return int64_t(kblocks) * align_up(C, 8) * 128; }
};

// ---------------------------------------------------------------- masked window attention
struct MwaParamLayout {
    int C, heads, ws, N, d, dpad, kblocks;
    int64_t header;     // float[4]: scale, -, -, -
    int64_t wqkvT;      // fp32 [C][3C]   qkv.weight transposed  ([in][out])
    int64_t bqkv;       // fp32 [3C]      (zeros when qkv_bias=False)
    int64_t wprojT;     // fp32 [C][C]    proj.weight transposed
    int64_t bproj;      // fp32 [C]
    int64_t bias;       // fp32 [heads][N][N]   expanded relative position bias
    // fp16 UMMA B-operand images (K-major SW128), used by the tcgen05 kernel
    int64_t img_wqkv;   // 3 sections (q,k,v); each kblocks x [rows_qkv x 64 k]; q,k rows are head-padded (heads*dpad),
                        // q rows pre-multiplied by scale; v rows = C
    int64_t img_wproj;  // kblocks x [C rows x 64 k]
    int64_t bq_pad;     // fp32 [3][heads*dpad]  bias in the padded q/k row order (q part pre-scaled), v part [C]
    int64_t total;
    __host__ __device__ MwaParamLayout(int C_, int heads_, int ws_)
        : C(C_), heads(heads_), ws(ws_), N(ws_ * ws_), d(C_ / heads_), dpad(int(align_up(C_ / heads_, 16))),
          kblocks((C_ + 63) / 64) {
        int64_t o = 0;
        header = o;    o = align_up(o + 16, 1024);
        wqkvT = o;     o = align_up(o + 4ll * C * 3 * C, 1024);
        bqkv = o;      o = align_up(o + 4ll * 3 * C, 1024);
        wprojT = o;    o = align_up(o + 4ll * C * C, 1024);
        bproj = o;     o = align_up(o + 4ll * C, 1024);
        bias = o;      o = align_up(o + 4ll * heads * N * N, 1024);
        img_wqkv = o;  o = align_up(o + 3 * qkv_section_bytes(), 1024);
        img_wproj = o; o = align_up(o + int64_t(kblocks) * align_up(C, 8) * 128, 1024);
        bq_pad = o;    o = align_up(o + 4ll * 3 * hp(), 1024);
        total = o;
    }
    __host__ __device__ int hp() const { return heads * dpad; }          // padded q/k width (>= C)
    __host__ __device__ int64_t qkv_section_bytes() const { return int64_t(kblocks) * hp() * 128; }
};

}  // namespace b200
