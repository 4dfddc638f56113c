// Fused masked window attention forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
// Reference semantics: layers/masked_win_attention.py:169-251 (block) and :96-131 (window attention).
//
// Three launches per forward, no host synchronisation:
//   1. mwa_scan_kernel      keep flag per window (sum of the cyclically shifted alpha != 0, :35-47); dropped
//                           windows are copied x -> out right here (the block is the identity there).
//   2. mwa_compact_kernel   ordered compaction of the kept windows into a work list + count.
//   3. mwa_tc_kernel        persistent, one CTA per SM; a tile = 128 tokens = 2 kept 8x8 windows (8 kept 4x4
//                           windows); everything between the x load and the out store stays on chip.
//
// Tile pipeline (16 warps; thread 0 issues every MMA and feeds the weight ring):
//   load     x (fp32, NCHW or NHWC, cyclic shift + window partition by index arithmetic) -> fp16 A operand in smem
//   per head group g (64 padded q/k/v columns = 2 heads of d<=32 or 4 heads of d<=16):
//     QKV    D[128 x 192] = X[128 x C] * Wqkv_g^T   (weights streamed from L2 through a 2-slot bulk-copy ring;
//            q rows pre-scaled by d^-1/2 and zero-padded per head at prepare time)
//     drain  + bias -> fp16 -> Q_g / K_g (K-major) and V_g^T (keys contiguous) operands in smem
//     per head: S = Q_h K_h^T over all 128 tokens of the tile (block diagonal: a token only uses the 64 / 16
//            columns of its own window), + relative-position bias + SW-MSA region mask (from coordinates),
//            row softmax split over 4 threads per row (max exchanged through smem, sums deferred),
//            P (fp16, exact zeros off the diagonal blocks) -> smem, O_h = P V_h accumulated in TMEM
//     drain  O / rowsum -> fp16 A operand;  proj: Dproj[128 x C] += O_g * Wproj_g^T  (accumulated over groups)
//   epilogue out = x + Dproj + bproj, scattered back to the un-shifted pixel positions.
// Operand precision: fp16 x fp16 -> fp32 accumulate (kind::f16); softmax, bias, residual in fp32.
#include "mwa_tc_shared.cuh"

namespace b200 {
namespace {


// One attention task: rows row0..row0+15 of the tile (all inside one window whose keys are rows key0..key0+NTOK-1)
// for the head at 16-byte-chunk offset cb of the 64-column group buffers.  S = Q K^T + bias + mask; softmax; O = P V;
// O / rowsum -> fp16 into the projection operand buffer.  rowmask[i] (i = 0: row lane/4, 1: row lane/4 + 8) holds the
// SW-MSA band bits of that row: bit yj = its row band differs from key row yj, bit 8 + xj likewise for columns.
template <class CF>
__device__ __forceinline__ void attention_task(uint32_t sQ, uint32_t sK, uint32_t sV, uint32_t sO, const float* tb,
                                               int row0, int key0, int cb, const uint32_t (&rowmask)[2], int lane) {
    constexpr int WS = CF::WS, NTOK = CF::NTOK, DPAD = CF::DPAD;
    constexpr int KS = DPAD / 16, NT = NTOK / 8, PK = NTOK / 16, ON = DPAD / 8;
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) ldmatrix_x4(qa[ks], sQ + swz(row0 + (lane & 15), cb + 2 * ks + (lane >> 4)));
    float sc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t kb[2];
            ldmatrix_x2(kb, sK + swz(key0 + 8 * n + (lane & 7), cb + 2 * ks + ((lane >> 3) & 1)));
            mma16816(sc[n], qa[ks], kb);
        }
    }
    // bias + region mask, row maxima (rows lane/4 and lane/4 + 8; a row is spread over the 4 lanes of a quad)
    const int ta = (row0 + (lane >> 2)) % NTOK, tb2 = ta + 8;
    const int tya = ta / WS, txa = ta % WS, tyb = tb2 / WS, txb = tb2 % WS;
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int j = 8 * n + 2 * (lane & 3) + e, yj = j / WS, xj = j % WS;
            float va = sc[n][e] + tb[(tya - yj + WS - 1) * (2 * WS - 1) + (txa - xj + WS - 1)];
            float vb = sc[n][2 + e] + tb[(tyb - yj + WS - 1) * (2 * WS - 1) + (txb - xj + WS - 1)];
            if (((rowmask[0] >> yj) | (rowmask[0] >> (8 + xj))) & 1u) va += kNegMask;
            if (((rowmask[1] >> yj) | (rowmask[1] >> (8 + xj))) & 1u) vb += kNegMask;
            sc[n][e] = va;
            sc[n][2 + e] = vb;
            ma = fmaxf(ma, va);
            mb = fmaxf(mb, vb);
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
    for (int j = 0; j < PK; ++j)
#pragma unroll
        for (int nt = 0; nt < ON; ++nt) {
            uint32_t vb[2];
            ldmatrix_x2_trans(vb, sV + swz(key0 + 16 * j + (lane & 7) + 8 * ((lane >> 3) & 1), cb + nt));
            mma16816(oc[nt], pa[j], vb);
        }
    const float inva = 1.f / suma, invb = 1.f / sumb;
    const uint32_t ra = row0 + (lane >> 2), rb = ra + 8;
#pragma unroll
    for (int nt = 0; nt < ON; ++nt) {
        const uint32_t coff = (lane & 3) * 4;             // byte offset of the column pair inside its 16-byte chunk
        st_shared_b32(sO + swz(ra, cb + nt) + coff, pack_f16x2(oc[nt][0] * inva, oc[nt][1] * inva));
        st_shared_b32(sO + swz(rb, cb + nt) + coff, pack_f16x2(oc[nt][2] * invb, oc[nt][3] * invb));
    }
}

template <class CF, int VEC>
__global__ void __launch_bounds__(kThreads, 1)
mwa_tc_kernel(const float* __restrict__ x, float* __restrict__ out, const uint8_t* __restrict__ blk,
              const uint8_t* __restrict__ tcp, const int32_t* __restrict__ list, const int32_t* __restrict__ count_p,
              Geom geo, unsigned long long* __restrict__ timing) {
    constexpr int C = CF::C, WS = CF::WS, NTOK = CF::NTOK, DPAD = CF::DPAD, HPG = CF::HPG, NG = CF::NG;
    // optional phase timing (development aid): CTA 0 / thread 0 accumulates clock64() deltas per phase
    const bool do_time = timing != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    long long t_prev = do_time ? clock64() : 0;
    auto tick = [&](int phase) {
        if (do_time) {
            const long long now = clock64();
            timing[phase] += static_cast<unsigned long long>(now - t_prev);
            t_prev = now;
        }
    };
    constexpr int XI = (CF::NCHUNK + 3) / 4;            // 16-byte chunks of the x row handled per thread
    extern __shared__ __align__(1024) uint8_t smem[];   // base is 1024-aligned (checked once below)
    const uint32_t sb = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CF::oBars);
    uint64_t* bar_full = bars;                           // [CF::kQSlots + 1] weight slab landed (last: projection slab)
    uint64_t* bar_empty = bars + (CF::kQSlots + 1);          // [CF::kQSlots + 1] MMAs reading the slab done
    uint64_t* bar_mma = bars + 2 * (CF::kQSlots + 1);        // [2], used alternately: "MMAs issued so far are complete"
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + CF::oTmem);
    float* s_tbl = reinterpret_cast<float*>(smem + CF::oTbl);
    float* s_bqkv = reinterpret_cast<float*>(smem + CF::oBqkv);
    float* s_bproj = reinterpret_cast<float*>(smem + CF::oBproj);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cg = warp >> 2;
    const int r = q * 32 + lane;                        // token row of the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const MwaParamLayout L(C, CF::HEADS, WS);

    // ---- one-time setup
    if (tid == 0) {
        if (sb & 1023u) __trap();
        for (int i = 0; i < CF::kQSlots + 1; ++i) {
            mbar_init(bar_full + i, 1);
            mbar_init(bar_empty + i, 1);
        }
        mbar_init(bar_mma + 0, 1);
        mbar_init(bar_mma + 1, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_ptr);
    {   // padded-order biases; zero the operand buffers whose padding / off-diagonal blocks must read as exact
        // zeros for the whole kernel (Q/K/V pad columns come from zero weight rows)
        const float* bq = reinterpret_cast<const float*>(tcp + TcParams<CF>::bq);
        for (int i = tid; i < NG * CF::NQKV; i += kThreads) s_bqkv[i] = bq[i];
        const float* bp = reinterpret_cast<const float*>(blk + L.bproj);
        for (int i = tid; i < C; i += kThreads) s_bproj[i] = bp[i];
        for (int i = tid; i < (CF::oRing - CF::oX) / 16; i += kThreads)
            reinterpret_cast<uint4*>(smem + CF::oX)[i] = make_uint4(0, 0, 0, 0);
        // compact relative-position table s_tbl[h][idx], idx = (yi-yj+WS-1)*(2WS-1) + (xi-xj+WS-1), recovered from
        // the expanded bias[h][i][j] of the parameter block (any token pair with that offset carries the value)
        const float* bexp = reinterpret_cast<const float*>(blk + L.bias);
        for (int e = tid; e < CF::HEADS * CF::TBL; e += kThreads) {
            const int h = e / CF::TBL, idx = e % CF::TBL;
            const int dy = idx / (2 * WS - 1) - (WS - 1), dx = idx % (2 * WS - 1) - (WS - 1);   // yi - yj, xi - xj
            const int yi = dy >= 0 ? dy : 0, yj = dy >= 0 ? 0 : -dy;
            const int xi = dx >= 0 ? dx : 0, xj = dx >= 0 ? 0 : -dx;
            s_tbl[e] = bexp[(int64_t(h) * NTOK + (yi * WS + xi)) * NTOK + (yj * WS + xj)];
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = *tmem_ptr;

    const int count = *count_p;
    const int num_tiles = (count + CF::WPT - 1) / CF::WPT;
    const int64_t hw = int64_t(geo.H) * geo.W;
    int my_tiles = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) ++my_tiles;
    // ---- thread 0: two weight rings fed with bulk copies (L2 -> smem, TMA engine):
    //   QKV ring : CF::kQSlots slabs of [NQKV rows x 64 k]; slab i (running index) = K block i % KB of group (i / KB) % NG.
    //              All slabs of the NEXT group are requested as soon as the MMAs of the current group have
    //              released the slots, i.e. they travel during the whole attention phase.
    //   proj slot: one slab [C rows x 64 k] of group i % NG, requested when the previous projection MMA is done.
    const uint32_t my_qslabs = uint32_t(my_tiles) * NG * CF::KB, my_pslabs = uint32_t(my_tiles) * NG;
    uint32_t q_issued = 0, q_used = 0, p_issued = 0, p_used = 0, mma_waits = 0, mma_commits = 0;
    const uint8_t* wimg = tcp + TcParams<CF>::img;
    uint64_t* bar_pfull = bar_full + CF::kQSlots;
    uint64_t* bar_pempty = bar_empty + CF::kQSlots;
    // All ring / MMA-issue code below is executed by the WHOLE warp 0 in convergence (uniform state in every lane);
    // the instructions with side effects are issued by one elected lane.  (A lone thread 0 issuing while its 31
    // siblings spin in mbarrier.try_wait shares the warp's issue slot with that spin loop and is stalled by its
    // hardware suspend: measured ~200 cycles per MMA issue and ~500 per barrier probe.)
    auto issue_q = [&]() {
        const uint32_t slot = q_issued % CF::kQSlots, g = (q_issued / CF::KB) % NG, kb = q_issued % CF::KB;
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_full + slot, CF::kQkvSlabBytes);
            bulk_g2s(smem + CF::oRing + slot * CF::kQkvSlabBytes,
                     wimg + int64_t(g) * CF::kGroupBytes + int64_t(kb) * CF::kQkvSlabBytes, CF::kQkvSlabBytes,
                     bar_full + slot);
        }
        __syncwarp();
        ++q_issued;
    };
    auto issue_p = [&]() {
        const uint32_t g = p_issued % NG;
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_pfull, CF::kProjSlabBytes);
            bulk_g2s(smem + CF::oRingP, wimg + int64_t(g) * CF::kGroupBytes + int64_t(CF::KB) * CF::kQkvSlabBytes,
                     CF::kProjSlabBytes, bar_pfull);
        }
        __syncwarp();
        ++p_issued;
    };
    // keep both rings as full as the completed MMAs allow; never blocks
    auto feed_try = [&]() {
        while (q_issued < my_qslabs && q_issued < q_used + CF::kQSlots) {
            const uint32_t slot = q_issued % CF::kQSlots, use = q_issued / CF::kQSlots;
            if (use > 0 && !mbar_test_wait(bar_empty + slot, (use - 1) & 1)) break;
            issue_q();
        }
        if (p_issued < my_pslabs && p_issued == p_used &&
            (p_issued == 0 || mbar_test_wait(bar_pempty, (p_issued - 1) & 1)))
            issue_p();
    };
    auto acquire_q = [&]() -> uint32_t {                 // slab `q_used` is needed now
        while (q_issued <= q_used) {
            const uint32_t slot = q_issued % CF::kQSlots, use = q_issued / CF::kQSlots;
            if (use > 0) mbar_wait(bar_empty + slot, (use - 1) & 1);
            issue_q();
        }
        const uint32_t slot = q_used % CF::kQSlots, use = q_used / CF::kQSlots;
        mbar_wait(bar_full + slot, use & 1);
        tc_fence_after_sync();
        return slot;
    };
    auto acquire_p = [&]() {
        if (p_issued <= p_used) {
            if (p_issued > 0) mbar_wait(bar_pempty, (p_issued - 1) & 1);
            issue_p();
        }
        mbar_wait(bar_pfull, p_used & 1);
        tc_fence_after_sync();
    };
    // thread 0: commit k goes to barrier k & 1; every thread: wait k on the same barrier with parity (k >> 1) & 1.
    // At most two commits are ever issued between two CTA-wide barriers, so a slow waiter can never be lapped.
    auto commit_mma = [&]() {                             // warp 0
        if (elect_one()) umma_commit(bar_mma + (mma_commits & 1));
        __syncwarp();
        ++mma_commits;
    };
    auto wait_mma = [&]() {
        mbar_wait_sleep(bar_mma + (mma_waits & 1), (mma_waits >> 1) & 1);
        ++mma_waits;
        __syncwarp();
        tc_fence_after_sync();
        if (warp == 0) feed_try();
    };

    // geometry of this thread's token row in a tile
    const int wslot = r / NTOK, tok = r % NTOK;
    auto row_base = [&](int tile, bool& valid) -> int64_t {
        const int lidx = tile * CF::WPT + wslot;
        valid = lidx < count;
        const int win = list[valid ? lidx : (count - 1)];
        int b, wy, wx, py, px;
        window_coords(geo, win, b, wy, wx);
        token_pixel<WS>(geo, wy, wx, tok, py, px);
        const int64_t pix = int64_t(py) * geo.W + px;
        return geo.channels_last ? (int64_t(b) * hw + pix) * C : int64_t(b) * C * hw + pix;
    };

    if (warp == 0) feed_try();
    tick(0);                                             // 0: prologue

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        bool row_valid;
        const int64_t base = row_base(tile, row_valid);
        int wy, wx;
        {
            int b;
            window_coords(geo, list[row_valid ? tile * CF::WPT + wslot : count - 1], b, wy, wx);
        }

        // ---------------- load x -> fp16 A operand.
        // NCHW: lane l owns the 4 consecutive tokens 4l..4l+3 of the tile (one window row or half of it, contiguous
        // pixels), warp w owns the 8-channel chunks w, w+16: every load instruction moves 512 contiguous-by-row
        // bytes of one channel.  NHWC: thread (row r, cg) reads its own 32-byte chunks.
        int64_t gbase[4 / VEC];                 // NCHW: element offset of each VEC-pixel piece of this lane (channel 0)
        bool lane_valid = false;
        if (!geo.channels_last) {
            const int r0 = 4 * lane, ws_l = r0 / NTOK, tok0 = r0 % NTOK;
            const int lidx = tile * CF::WPT + ws_l;
            lane_valid = lidx < count;
            int b, lwy, lwx;
            window_coords(geo, list[lane_valid ? lidx : count - 1], b, lwy, lwx);
            int py = lwy * WS + tok0 / WS + geo.shift;
            if (py >= geo.H) py -= geo.H;
#pragma unroll
            for (int pc = 0; pc < 4 / VEC; ++pc) {
                int px = lwx * WS + tok0 % WS + geo.shift + pc * VEC;
                if (px >= geo.W) px -= geo.W;
                gbase[pc] = int64_t(b) * C * hw + int64_t(py) * geo.W + px;
            }
            constexpr int NCI = (CF::NCHUNK + kWarps - 1) / kWarps;
            float v[NCI][8][4];
#pragma unroll
            for (int i = 0; i < NCI; ++i) {
                const int ci = warp + kWarps * i;
                if (ci < CF::NCHUNK) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float* pcn = x + int64_t(ci * 8 + j) * hw;
#pragma unroll
                        for (int pc = 0; pc < 4 / VEC; ++pc) {
                            if constexpr (VEC == 4) {
                                const float4 t4 = __ldg(reinterpret_cast<const float4*>(pcn + gbase[pc]));
                                v[i][j][0] = t4.x; v[i][j][1] = t4.y; v[i][j][2] = t4.z; v[i][j][3] = t4.w;
                            } else if constexpr (VEC == 2) {
                                const float2 t2 = __ldg(reinterpret_cast<const float2*>(pcn + gbase[pc]));
                                v[i][j][2 * pc] = t2.x; v[i][j][2 * pc + 1] = t2.y;
                            } else {
                                v[i][j][pc] = __ldg(pcn + gbase[pc]);
                            }
                        }
                    }
                }
            }
            // Lane l stores its rows in the rotated order 4l + ((k + l/2) & 3): the 8 lanes of every 128-bit store
            // phase then hit 8 different (row & 7) values, i.e. 8 different 16-byte bank groups of the swizzled
            // layout (storing row 4l + k from every lane would be a 16-way bank conflict).
            const int rot = (lane >> 1) & 3;
#pragma unroll
            for (int i = 0; i < NCI; ++i) {
                const int ci = warp + kWarps * i;
                if (ci < CF::NCHUNK) {
                    uint32_t pk[4][4];                       // [token k][channel pair]
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int jp = 0; jp < 4; ++jp) pk[k][jp] = pack_f16x2(v[i][2 * jp][k], v[i][2 * jp + 1][k]);
                    // rotate the token index by `rot` with two conditional stages (no dynamic register indexing)
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {
                        uint32_t a0 = pk[0][jp], a1 = pk[1][jp], a2 = pk[2][jp], a3 = pk[3][jp];
                        if (rot & 1) { const uint32_t t0 = a0; a0 = a1; a1 = a2; a2 = a3; a3 = t0; }
                        if (rot & 2) { const uint32_t t0 = a0, t1 = a1; a0 = a2; a1 = a3; a2 = t0; a3 = t1; }
                        pk[0][jp] = a0; pk[1][jp] = a1; pk[2][jp] = a2; pk[3][jp] = a3;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int row = 4 * lane + ((k + rot) & 3);
                        const uint32_t addr = sb + CF::oX + (ci >> 3) * 16384 + (row >> 3) * 1024 + (row & 7) * 128 +
                                              (((ci & 7) ^ (row & 7)) << 4);
                        st_shared_v4(addr, pk[k][0], pk[k][1], pk[k][2], pk[k][3]);
                    }
                }
            }
        } else {
            const float* xb = x + base;
            float4 v[XI][2];
#pragma unroll
            for (int i = 0; i < XI; ++i) {
                const int ci = cg + 4 * i;
                if (ci < CF::NCHUNK) {
                    v[i][0] = __ldg(reinterpret_cast<const float4*>(xb + ci * 8));
                    v[i][1] = __ldg(reinterpret_cast<const float4*>(xb + ci * 8 + 4));
                }
            }
#pragma unroll
            for (int i = 0; i < XI; ++i) {
                const int ci = cg + 4 * i;
                if (ci < CF::NCHUNK) {
                    const uint32_t addr = sb + CF::oX + (ci >> 3) * 16384 + (r >> 3) * 1024 + (r & 7) * 128 +
                                          (((ci & 7) ^ (r & 7)) << 4);
                    st_shared_v4(addr, pack_f16x2(v[i][0].x, v[i][0].y), pack_f16x2(v[i][0].z, v[i][0].w),
                                 pack_f16x2(v[i][1].x, v[i][1].y), pack_f16x2(v[i][1].z, v[i][1].w));
                }
            }
        }
        // the x rows of this CTA's next tile -> L2 (its loads and the residual re-read of the epilogue then hit L2)
        if (tile + int(gridDim.x) < num_tiles) {
            bool nv;
            const float* nb = x + row_base(tile + gridDim.x, nv);
            if (geo.channels_last) {
                for (int i = cg; i < (C * 4 + 127) / 128; i += 4) prefetch_l2(nb + i * 32);
            } else if ((tok % WS) == 0) {            // one sector (or half) per window row and channel
                for (int c = cg; c < C; c += 4) prefetch_l2(nb + int64_t(c) * hw);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        tick(1);                                         // 1: x load + convert

        // ---------------- per-warp attention tasks of this tile: geometry + SW-MSA region mask bits (:194-216).
        // task t = warp + 16 i  ->  (16-row block, window slot, head of the group); its two rows per lane are
        // lane/4 and lane/4 + 8 of the block.
        constexpr int TPW = CF::TASKS / kWarps;
        uint32_t task_mask[TPW][2];
#pragma unroll
        for (int i = 0; i < TPW; ++i) {
            const int tsk = warp + kWarps * i;
            const int rbk = tsk % CF::RB, tws = (tsk / CF::RB) % CF::WPT;
            task_mask[i][0] = task_mask[i][1] = 0;
            if (geo.shift > 0) {
                const int lidx = tile * CF::WPT + tws;
                int tb_, twy, twx;
                window_coords(geo, list[lidx < count ? lidx : count - 1], tb_, twy, twx);
                const int ys0 = twy * WS, xs0 = twx * WS;
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int tk = rbk * 16 + (lane >> 2) + 8 * h2;          // token of this lane's row in its window
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
                    task_mask[i][h2] = mbits;
                }
            }
        }

        // QKV GEMM of head group g: D_qkv[128 x NQKV] = X * Wqkv_g^T   (warp 0; weights from the slab ring)
        auto issue_qkv = [&]() {
            tc_fence_after_sync();
            constexpr uint32_t idesc = umma_idesc(kFmtF16, kFmtF16, kTileM, CF::NQKV);
#pragma unroll
            for (int kb = 0; kb < CF::KB; ++kb) {
                long long tq0 = do_time ? clock64() : 0;
                const uint32_t slot = acquire_q();
                if (do_time) timing[16 + kb] += static_cast<unsigned long long>(clock64() - tq0);
                const uint64_t a0 = umma_desc_k_sw128(sb + CF::oX + kb * 16384);
                const uint64_t b0 = umma_desc_k_sw128(sb + CF::oRing + slot * CF::kQkvSlabBytes);
                const int nks = (kb == CF::KB - 1) ? (CF::KSTEPS - 4 * (CF::KB - 1)) : 4;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        if (ks < nks) umma_f16_ss(tm + CF::tA, a0 + ks * 2, b0 + ks * 2, idesc, (kb | ks) != 0);
                    umma_commit(bar_empty + slot);
                }
                __syncwarp();
                ++q_used;
            }
            commit_mma();
        };

        if (warp == 0) issue_qkv();
        tick(2);                                         // 2: QKV issue of group 0 (incl. slab waits)

        for (int g = 0; g < NG; ++g) {
            wait_mma();                                  // D_qkv of group g complete
            tick(3);                                     // 3: QKV MMA completion wait

            // ---------------- drain: q (cg 0) / k (cg 1) / v (cg 2): NQ un-padded accumulator columns + bias -> fp16,
            //                  re-spaced to the head-padded [128 x 64] operand layout (pad columns written as zeros);
            //                  the three buffers share one layout (K-major SWIZZLE_128B), read back with ldmatrix
            if (cg < 3) {
                constexpr int NQ = CF::NQ, D = CF::D;
                const float* bias = s_bqkv + g * CF::NQKV + cg * NQ;
                const uint32_t rowaddr = sb + (cg == 0 ? CF::oQ : cg == 1 ? CF::oK : CF::oV) + (r >> 3) * 1024 + (r & 7) * 128;
                auto put8 = [&](int ch, const float (&v)[8]) {   // 8 padded columns 8*ch .. 8*ch+7 of this row
                    st_shared_v4(rowaddr + ((ch ^ (r & 7)) << 4), pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]),
                                 pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
                };
                if constexpr (D % 8 == 0) {
#pragma unroll
                    for (int hh = 0; hh < HPG; ++hh)
#pragma unroll
                        for (int ci = 0; ci < DPAD / 8; ++ci) {
                            float v[8];
                            if (ci * 8 < D) {
                                uint32_t acc[8];
                                tmem_ld_x8(tm + CF::tA + lane_addr + cg * NQ + hh * D + ci * 8, acc);
                                tmem_wait_ld();
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[j]) + bias[hh * D + ci * 8 + j];
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = 0.f;
                            }
                            put8(hh * (DPAD / 8) + ci, v);
                        }
                } else {
                    float val[NQ];
#pragma unroll
                    for (int cc = 0; cc < NQ / 8; ++cc) {
                        uint32_t acc[8];
                        tmem_ld_x8(tm + CF::tA + lane_addr + cg * NQ + cc * 8, acc);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 8; ++j) val[cc * 8 + j] = __uint_as_float(acc[j]) + bias[cc * 8 + j];
                    }
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int col = 8 * ch + j;
                            v[j] = ((col % DPAD) < D) ? val[(col / DPAD) * D + (col % DPAD)] : 0.f;
                        }
                        put8(ch, v);
                    }
                }
            }
            tc_fence_before_sync();
            __syncthreads();
            tick(4);                                     // 4: q/k/v drain

            // ---------------- the next group's QKV GEMM runs on the tensor core while the warps do the attention
            //                  core of this group (D_qkv has just been drained; X is still resident)
            if (warp == 0 && g + 1 < NG) issue_qkv();
            tick(5);                                     // 5: QKV issue of the next group

            // ---------------- attention core: one (window, head, 16-row block) task per warp, entirely in registers
            {
                const uint32_t sO = sb + CF::oO + (g & 1) * 16384;
#pragma unroll
                for (int i = 0; i < TPW; ++i) {
                    const int tsk = warp + kWarps * i;
                    const int rbk = tsk % CF::RB, tws = (tsk / CF::RB) % CF::WPT, hh = tsk / (CF::RB * CF::WPT);
                    attention_task<CF>(sb + CF::oQ, sb + CF::oK, sb + CF::oV, sO, s_tbl + (g * HPG + hh) * CF::TBL,
                                       tws * NTOK + rbk * 16, tws * NTOK, hh * (DPAD / 8), task_mask[i], lane);
                }
            }
            fence_proxy_async_smem();                    // O_g is read by the projection MMA (async proxy)
            tc_fence_before_sync();
            __syncthreads();
            tick(6);                                     // 6: attention core

            // ---------------- projection partial sum over this group's columns
            if (warp == 0) {
                tc_fence_after_sync();
                constexpr uint32_t idesc_p = umma_idesc(kFmtF16, kFmtF16, kTileM, C);
                acquire_p();
                const uint64_t a0 = umma_desc_k_sw128(sb + CF::oO + (g & 1) * 16384), b0 = umma_desc_k_sw128(sb + CF::oRingP);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_f16_ss(tm + CF::tP, a0 + ks * 2, b0 + ks * 2, idesc_p, (g | ks) != 0);
                    umma_commit(bar_pempty);
                }
                __syncwarp();
                ++p_used;
                if (g == NG - 1) commit_mma();
            }
            tick(11);                                    // 11: projection issue (incl. slab wait)
        }

        // ---------------- epilogue: out = x + proj + bias at the un-shifted pixel.
        // NCHW: the projection tile is staged in shared memory as [channel][token] (the operand buffers are dead now)
        // and written with the same lane = 4 tokens / warp = channel mapping as the load (512-byte row groups per
        // store instruction); the residual x loads are issued before waiting for the last projection MMA.
        if (!geo.channels_last) {
            constexpr int NCH = (C + kWarps - 1) / kWarps;         // channels per warp
            float xres[NCH][4];
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int c = warp + kWarps * i;
                if (c < C) {
                    const float* pcn = x + int64_t(c) * hw;
#pragma unroll
                    for (int pc = 0; pc < 4 / VEC; ++pc) {
                        if constexpr (VEC == 4) {
                            const float4 t4 = __ldg(reinterpret_cast<const float4*>(pcn + gbase[pc]));
                            xres[i][0] = t4.x; xres[i][1] = t4.y; xres[i][2] = t4.z; xres[i][3] = t4.w;
                        } else if constexpr (VEC == 2) {
                            const float2 t2 = __ldg(reinterpret_cast<const float2*>(pcn + gbase[pc]));
                            xres[i][2 * pc] = t2.x; xres[i][2 * pc + 1] = t2.y;
                        } else {
                            xres[i][pc] = __ldg(pcn + gbase[pc]);
                        }
                    }
                }
            }
            tick(12);                                    // 12: residual loads issued
            wait_mma();
            tick(13);                                    // 13: wait last projection
            float* stage = reinterpret_cast<float*>(smem + CF::oStage);
            constexpr int EG = (CF::NCOLG + 3) / 4;
#pragma unroll
            for (int i = 0; i < EG; ++i) {
                const int gi = cg + 4 * i;
                if (gi < CF::NCOLG) {                   // warp-uniform (cg): tcgen05.ld stays warp-collective
                    uint32_t acc[16];
                    tmem_ld_x16(tm + CF::tP + lane_addr + gi * 16, acc);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        stage[(gi * 16 + j) * kTileM + r] = __uint_as_float(acc[j]) + s_bproj[gi * 16 + j];
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int c = warp + kWarps * i;
                if (c < C && lane_valid) {
                    const float4 sv = *reinterpret_cast<const float4*>(stage + c * kTileM + 4 * lane);
                    float* pon = out + int64_t(c) * hw;
                    const float o0 = xres[i][0] + sv.x, o1 = xres[i][1] + sv.y, o2 = xres[i][2] + sv.z,
                                o3 = xres[i][3] + sv.w;
                    if constexpr (VEC == 4) {
                        *reinterpret_cast<float4*>(pon + gbase[0]) = make_float4(o0, o1, o2, o3);
                    } else if constexpr (VEC == 2) {
                        *reinterpret_cast<float2*>(pon + gbase[0]) = make_float2(o0, o1);
                        *reinterpret_cast<float2*>(pon + gbase[1]) = make_float2(o2, o3);
                    } else {
                        pon[gbase[0]] = o0; pon[gbase[1]] = o1; pon[gbase[2]] = o2; pon[gbase[3]] = o3;
                    }
                }
            }
        } else {
            constexpr int EG = (CF::NCOLG + 3) / 4;
            const float* xb = x + base;
            float* ob = out + base;
            float4 xres[EG][4];
#pragma unroll
            for (int i = 0; i < EG; ++i) {
                const int gi = cg + 4 * i;
                if (gi < CF::NCOLG) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) xres[i][j] = __ldg(reinterpret_cast<const float4*>(xb + gi * 16 + 4 * j));
                }
            }
            tick(12);
            wait_mma();
            tick(13);
#pragma unroll
            for (int i = 0; i < EG; ++i) {
                const int gi = cg + 4 * i;
                if (gi < CF::NCOLG) {
                    uint32_t acc[16];
                    tmem_ld_x16(tm + CF::tP + lane_addr + gi * 16, acc);
                    tmem_wait_ld();
                    if (row_valid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float* bp = s_bproj + gi * 16 + 4 * j;
                            float4 o;
                            o.x = xres[i][j].x + (__uint_as_float(acc[4 * j + 0]) + bp[0]);
                            o.y = xres[i][j].y + (__uint_as_float(acc[4 * j + 1]) + bp[1]);
                            o.z = xres[i][j].z + (__uint_as_float(acc[4 * j + 2]) + bp[2]);
                            o.w = xres[i][j].w + (__uint_as_float(acc[4 * j + 3]) + bp[3]);
                            *reinterpret_cast<float4*>(ob + gi * 16 + 4 * j) = o;
                        }
                    }
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();
        tick(14);                                        // 14: epilogue stores
        if (do_time) timing[15] += 1;                    // 15: tiles
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

// ------------------------------------------------------------------------------------------------ host side
unsigned long long* g_timing = nullptr;     // development aid, see mwa_debug_set_timing_buffer

template <class CF>
int launch_tc(const float* x, const float* alpha, float* out, const void* params, int B, int H, int W, int shift,
              int channels_last, int32_t* kept_count, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    const Geom geo{B, H, W, shift, W / CF::WS, H / CF::WS, channels_last};
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
    const int vec = (shift % 4 == 0 && W % 4 == 0) ? 4 : (shift % 2 == 0 && W % 2 == 0) ? 2 : 1;
    if (alpha != nullptr) {
        if (vec == 4) mwa_scan_kernel<CF::WS, 4><<<(nwin + 7) / 8, 256, 0, st>>>(x, alpha, out, geo, CF::C, nwin, flags);
        else if (vec == 2) mwa_scan_kernel<CF::WS, 2><<<(nwin + 7) / 8, 256, 0, st>>>(x, alpha, out, geo, CF::C, nwin, flags);
        else mwa_scan_kernel<CF::WS, 1><<<(nwin + 7) / 8, 256, 0, st>>>(x, alpha, out, geo, CF::C, nwin, flags);
        int rc = check_launch("mwa_forward(scan)");
        if (rc != MWA_OK) return rc;
    }
    mwa_compact_kernel<<<1, 1024, 0, st>>>(alpha ? flags : nullptr, nwin, list, count);
    int rc = check_launch("mwa_forward(compact)");
    if (rc != MWA_OK) return rc;
    const int smem = CF::oTotal;
    const int max_tiles = (nwin + CF::WPT - 1) / CF::WPT;
    const int grid = max_tiles < kNumSMs ? max_tiles : kNumSMs;
#define MWA_LAUNCH_TC(V)                                                                                            \
    do {                                                                                                            \
        MWA_TRY_CUDA(cudaFuncSetAttribute(mwa_tc_kernel<CF, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), \
                     "mwa_forward(tc attr)");                                                                       \
        mwa_tc_kernel<CF, V><<<grid, kThreads, smem, st>>>(x, out, blk, blk + L.img_wqkv, list, count, geo,         \
                                                          g_timing);                                                \
    } while (0)
    if (vec == 4) MWA_LAUNCH_TC(4);
    else if (vec == 2) MWA_LAUNCH_TC(2);
    else MWA_LAUNCH_TC(1);
#undef MWA_LAUNCH_TC
    rc = check_launch("mwa_forward(tcgen05)");
    if (rc != MWA_OK) return rc;
    if (kept_count)
        MWA_TRY_CUDA(cudaMemcpyAsync(kept_count, count, sizeof(int32_t), cudaMemcpyDeviceToDevice, st),
                     "mwa_forward(kept_count)");
    return MWA_OK;
}

template <class CF>
void prepare_tc(const float* qkv_w, const float* qkv_b, const float* proj_w, const float* proj_b, float scale,
                uint8_t* dst, cudaStream_t st) {
    zero16_kernel<<<64, 256, 0, st>>>(reinterpret_cast<uint4*>(dst), TcParams<CF>::total / 16);
    mwa_tc_prepare_kernel<CF><<<148, 256, 0, st>>>(qkv_w, qkv_b, proj_w, proj_b, scale, dst);
}

using Cfg192h8 = Cfg<192, 8, 8>;
using Cfg192h6 = Cfg<192, 6, 8>;
using Cfg80h8 = Cfg<80, 8, 4>;

}  // namespace

int64_t mwa_tc_param_bytes(int C, int heads, int ws) {
    if (C == 192 && heads == 8 && ws == 8) return TcParams<Cfg192h8>::total;
    if (C == 192 && heads == 6 && ws == 8) return TcParams<Cfg192h6>::total;
    if (C == 80 && heads == 8 && ws == 4) return TcParams<Cfg80h8>::total;
    return 0;
}

bool mwa_tc_supported(int C, int heads, int ws, int H, int W, int shift, int channels_last) {
    (void)H; (void)W; (void)shift; (void)channels_last;
    return mwa_tc_param_bytes(C, heads, ws) > 0;
}

int64_t mwa_tc_workspace_bytes(int64_t nwin) { return ScanWs(nwin).total; }
void mwa_tc_set_timing_buffer(void* p) { g_timing = static_cast<unsigned long long*>(p); }

void mwa_tc_prepare_images(const float* qkv_w, const float* qkv_b, const float* proj_w, const float* proj_b, int C,
                           int heads, int ws, float scale, uint8_t* blk, cudaStream_t st) {
    const MwaParamLayout L(C, heads, ws);
    uint8_t* dst = blk + L.img_wqkv;
    if (C == 192 && heads == 8 && ws == 8) prepare_tc<Cfg192h8>(qkv_w, qkv_b, proj_w, proj_b, scale, dst, st);
    else if (C == 192 && heads == 6 && ws == 8) prepare_tc<Cfg192h6>(qkv_w, qkv_b, proj_w, proj_b, scale, dst, st);
    else if (C == 80 && heads == 8 && ws == 4) prepare_tc<Cfg80h8>(qkv_w, qkv_b, proj_w, proj_b, scale, dst, st);
}

int mwa_forward_tc(const float* x, const float* alpha, float* out, const void* params, int B, int C, int H, int W,
                   int heads, int ws, int shift, int channels_last, int32_t* kept_count, void* workspace,
                   int64_t workspace_bytes, cudaStream_t st) {
    if (C == 192 && heads == 8 && ws == 8)
        return launch_tc<Cfg192h8>(x, alpha, out, params, B, H, W, shift, channels_last, kept_count, workspace,
                                   workspace_bytes, st);
    if (C == 192 && heads == 6 && ws == 8)
        return launch_tc<Cfg192h6>(x, alpha, out, params, B, H, W, shift, channels_last, kept_count, workspace,
                                   workspace_bytes, st);
    if (C == 80 && heads == 8 && ws == 4)
        return launch_tc<Cfg80h8>(x, alpha, out, params, B, H, W, shift, channels_last, kept_count, workspace,
                                  workspace_bytes, st);
    return MWA_ERR_UNSUPPORTED;
}

}  // namespace b200
