// Masked window attention: backward (training), fp32 SIMT, one window per CTA iteration, persistent grid.
// Reference semantics: autograd through layers/masked_win_attention.py:169-251 (block) and :96-131 (window attention);
// layers/win_attention.py:153-207 (alpha == NULL); token mode = WindowAttention.forward with an additive mask.
//
// Per kept window the kernel re-computes the forward per head (q, k, v, P) and produces
//   grad_x    = grad_out + dXw            (dropped windows: grad_x = grad_out, the block is the identity there)
//   grad_table (accumulated over windows in shared memory, flushed once per CTA with atomics)
// and writes four token-major scratch tensors from which the remaining parameter gradients are PLAIN GEMMs /
// column sums over all tokens (done by the caller with a library GEMM):
//   xw_tok  (nwin, N, C)   gathered input tokens          dWqkv = dqkv_tok^T xw_tok      dbqkv = sum dqkv_tok
//   ao_tok  (nwin, N, C)   head-concatenated P V          dWproj = dy_tok^T ao_tok       dbproj = sum dy_tok
//   dy_tok  (nwin, N, C)   gathered grad_out tokens
//   dqkv_tok(nwin, N, 3C)  gradient wrt the qkv Linear output (q part already multiplied by the softmax scale)
// Rows of dropped windows are written as zeros, so the GEMMs need no compaction.
//
// Math per head (q' = scale * q):  S = q' k^T + bias + mask,  P = softmax(S),  AO = P v,  y = AO Wproj^T + b
//   dAO = dy Wproj          dP = dAO v^T        dv = P^T dAO        dS = P * (dP - rowsum(dP * P))
//   dq' = dS k              dk = dS^T q'        dbias += dS         dXw = dqkv Wqkv
#include "common.cuh"
#include "params.cuh"
#include "status.cuh"

namespace b200 {
namespace {

constexpr int kBT = 1024;                           // threads per CTA (32 warps: the kernel is latency-bound)
constexpr float kNegMaskB = -100.0f;                // layers/masked_win_attention.py:214

struct BGeo {
    int B, C, H, W, ws, shift, nwx, nwy, N;
    int channels_last;
    int tokens;          // 1: tensors are (K, N, C) window tokens, no geometry / residual / keep predicate
    int mask_nw;         // token mode: windows in the external additive mask (0 = none)
    __device__ __forceinline__ void token_pixel(int wy, int wx, int t, int& y, int& x) const {
        y = wy * ws + t / ws + shift; if (y >= H) y -= H;
        x = wx * ws + t % ws + shift; if (x >= W) x -= W;
    }
    __device__ __forceinline__ int region(int wy, int wx, int t) const {
        const int ys = wy * ws + t / ws, xs = wx * ws + t % ws;
        return 3 * ((ys >= H - ws) + (ys >= H - shift)) + (xs >= W - ws) + (xs >= W - shift);
    }
    __device__ __forceinline__ int64_t elem(int win, int b, int wy, int wx, int t, int c) const {
        if (tokens) return (int64_t(win) * N + t) * C + c;
        int y, xx;
        token_pixel(wy, wx, t, y, xx);
        return channels_last ? ((int64_t(b) * H + y) * W + xx) * C + c : ((int64_t(b) * C + c) * H + y) * W + xx;
    }
};

// out[n][o] = sum_k A[n][k] * Wm[k * ldw + o] (+ bias[o]);  A in smem (row stride lda); lanes run over o.
template <int TN, class Epilogue>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ A, int lda, const float* __restrict__ Wm, int ldw,
                                          const float* __restrict__ bias, int N, int O, int K, Epilogue epi) {
    const int nchunks = (N + TN - 1) / TN;
    for (int item = threadIdx.x; item < O * nchunks; item += kBT) {
        const int o = item % O, n0 = (item / O) * TN;
        float acc[TN];
        const float b = bias ? bias[o] : 0.f;
#pragma unroll
        for (int i = 0; i < TN; ++i) acc[i] = b;
        for (int k = 0; k < K; k += 4) {
            float w[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) w[kk] = (k + kk < K) ? __ldg(Wm + int64_t(k + kk) * ldw + o) : 0.f;
#pragma unroll
            for (int i = 0; i < TN; ++i) {
                if (n0 + i < N) {
                    const float* a = A + (n0 + i) * lda + k;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        if (k + kk < K) acc[i] = fmaf(a[kk], w[kk], acc[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < TN; ++i)
            if (n0 + i < N) epi(n0 + i, o, acc[i]);
    }
}

// shared memory (floats):  R1 [N][C]  dy, then xw | R2 [N][C] dAO | R3 [N][3d+1] q' k v of the head |
//                          R4 [N][N+1] S / P | R5 [N][N+1] dP / dS | R6 [N][3d+1] dq dk dv | tacc [heads][TBL]
// after the head loop R1..R6 are re-used as one [N][3C+1] buffer for dqkv.
__host__ __device__ inline int64_t bwd_smem_floats(int C, int heads, int ws) {
    const int64_t N = ws * ws, d = C / heads, tbl = (2 * ws - 1) * (2 * ws - 1);
    const int64_t work = 2 * N * C + 2 * N * (3 * d + 1) + 2 * N * (N + 1);
    const int64_t dq = N * (3 * C + 1);
    return (work > dq ? work : dq) + heads * tbl + 8;
}

__global__ void __launch_bounds__(kBT)
mwa_bwd_kernel(const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ gout,
               const float* __restrict__ qkv_w, const float* __restrict__ proj_w, const uint8_t* __restrict__ blk,
               const float* __restrict__ ext_mask, float* __restrict__ gx, float* __restrict__ gtable,
               float* __restrict__ xw_tok, float* __restrict__ ao_tok, float* __restrict__ dy_tok,
               float* __restrict__ dqkv_tok, BGeo g, int heads, int nwin) {
    extern __shared__ float smem[];
    const int C = g.C, N = g.N, ws = g.ws, d = C / heads;
    const int ldh = 3 * d + 1, lds = N + 1, ldq = 3 * C + 1, TBL = (2 * ws - 1) * (2 * ws - 1);
    float* R1 = smem;
    float* R2 = R1 + N * C;
    float* R3 = R2 + N * C;
    float* R4 = R3 + N * ldh;
    float* R5 = R4 + N * lds;
    float* R6 = R5 + N * lds;
    const int64_t work = 2ll * N * C + 2ll * N * ldh + 2ll * N * lds, dqf = int64_t(N) * ldq;
    float* tacc = smem + (work > dqf ? work : dqf);
    __shared__ float red[kBT / 32];
    __shared__ int keep_s;

    const MwaParamLayout L(C, heads, ws);
    const float scale = reinterpret_cast<const float*>(blk + L.header)[0];
    const float* wqkvT = reinterpret_cast<const float*>(blk + L.wqkvT);     // [C][3C]
    const float* bqkv = reinterpret_cast<const float*>(blk + L.bqkv);
    const float* bias = reinterpret_cast<const float*>(blk + L.bias);       // [heads][N][N]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool c_fast = g.channels_last || g.tokens;

    for (int i = tid; i < heads * TBL; i += kBT) tacc[i] = 0.f;
    __syncthreads();

    for (int win = blockIdx.x; win < nwin; win += gridDim.x) {
        const int b = win / (g.nwy * g.nwx), wy = (win / g.nwx) % g.nwy, wx = win % g.nwx;
        bool keep = true;
        if (alpha != nullptr) {
            float a = 0.f;
            for (int t = tid; t < N; t += kBT) {
                int y, xx;
                g.token_pixel(wy, wx, t, y, xx);
                a += __ldg(alpha + (int64_t(b) * g.H + y) * g.W + xx);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) red[warp] = a;
            __syncthreads();
            if (tid == 0) {
                float tot = 0.f;
                for (int i = 0; i < kBT / 32; ++i) tot += red[i];
                keep_s = (tot != 0.f);
            }
            __syncthreads();
            keep = keep_s != 0;
        }
        const int64_t tok0 = int64_t(win) * N;
        if (!keep) {      // identity block: grad passes through; scratch rows are zeros
            for (int e = tid; e < N * C; e += kBT) {
                const int c = c_fast ? e % C : e / N, t = c_fast ? e / C : e % N;
                const int64_t off = g.elem(win, b, wy, wx, t, c);
                gx[off] = __ldg(gout + off);
                if (xw_tok) xw_tok[tok0 * C + e] = 0.f;
                ao_tok[tok0 * C + e] = 0.f;
                if (dy_tok) dy_tok[tok0 * C + e] = 0.f;
            }
            for (int e = tid; e < N * 3 * C; e += kBT) dqkv_tok[tok0 * 3 * C + e] = 0.f;
            __syncthreads();
            continue;
        }

        // ---- dy -> R1 (+ scratch), dAO = dy Wproj -> R2
        for (int e = tid; e < N * C; e += kBT) {
            const int c = c_fast ? e % C : e / N, t = c_fast ? e / C : e % N;
            const float v = __ldg(gout + g.elem(win, b, wy, wx, t, c));
            R1[t * C + c] = v;
            if (dy_tok) dy_tok[(tok0 + t) * C + c] = v;
        }
        __syncthreads();
        gemm_rows<8>(R1, C, proj_w, C, nullptr, N, C, C, [&](int n, int o, float v) { R2[n * C + o] = v; });
        __syncthreads();
        // ---- xw -> R1 (+ scratch)
        for (int e = tid; e < N * C; e += kBT) {
            const int c = c_fast ? e % C : e / N, t = c_fast ? e / C : e % N;
            const float v = __ldg(x + g.elem(win, b, wy, wx, t, c));
            R1[t * C + c] = v;
            if (xw_tok) xw_tok[(tok0 + t) * C + c] = v;
        }
        __syncthreads();

        for (int h = 0; h < heads; ++h) {
            // q' k v of this head: R3[n][part * d + c] = xw[n] . Wqkv[part*C + h*d + c] + b   (q' pre-scaled)
            for (int item = tid; item < 3 * d * (N / 4); item += kBT) {
                const int j = item % (3 * d), n0 = (item / (3 * d)) * 4;
                const int col = (j / d) * C + h * d + (j % d);
                float acc[4];
                const float bb = bqkv[col];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = bb;
                for (int k = 0; k < C; ++k) {
                    const float w = __ldg(wqkvT + int64_t(k) * 3 * C + col);
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] = fmaf(R1[(n0 + i) * C + k], w, acc[i]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) R3[(n0 + i) * ldh + j] = (j < d) ? acc[i] * scale : acc[i];
            }
            __syncthreads();
            // S = q' k^T + bias + mask -> R4
            const float* bh = bias + int64_t(h) * N * N;
            for (int e = tid; e < N * N; e += kBT) {
                const int i = e / N, j = e % N;
                const float* q = R3 + i * ldh;
                const float* k = R3 + j * ldh + d;
                float acc = 0.f;
                for (int c = 0; c < d; ++c) acc = fmaf(q[c], k[c], acc);
                acc += __ldg(bh + e);
                if (g.tokens) {
                    if (g.mask_nw > 0) acc += __ldg(ext_mask + (int64_t(win % g.mask_nw) * N + i) * N + j);
                } else if (g.shift > 0 && g.region(wy, wx, i) != g.region(wy, wx, j)) {
                    acc += kNegMaskB;
                }
                R4[i * lds + j] = acc;
            }
            __syncthreads();
            for (int i = warp; i < N; i += kBT / 32) {                 // softmax, warp per row
                float* row = R4 + i * lds;
                float m = -INFINITY;
                for (int j = lane; j < N; j += 32) m = fmaxf(m, row[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                float sum = 0.f;
                for (int j = lane; j < N; j += 32) {
                    const float e = expf(row[j] - m);
                    row[j] = e;
                    sum += e;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float inv = 1.f / sum;
                for (int j = lane; j < N; j += 32) row[j] *= inv;
            }
            __syncthreads();
            // AO_h = P v -> scratch ;  dP = dAO_h v^T -> R5 ;  dv = P^T dAO_h -> R6[:, 2d..3d)
            for (int e = tid; e < N * d; e += kBT) {
                const int i = e / d, c = e % d;
                const float* p = R4 + i * lds;
                float acc = 0.f;
                for (int j = 0; j < N; ++j) acc = fmaf(p[j], R3[j * ldh + 2 * d + c], acc);
                ao_tok[(tok0 + i) * C + h * d + c] = acc;
                float dv = 0.f;                                       // row i acts as key index here
                for (int n = 0; n < N; ++n) dv = fmaf(R4[n * lds + i], R2[n * C + h * d + c], dv);
                R6[i * ldh + 2 * d + c] = dv;
            }
            for (int e = tid; e < N * N; e += kBT) {
                const int i = e / N, j = e % N;
                const float* da = R2 + i * C + h * d;
                const float* v = R3 + j * ldh + 2 * d;
                float acc = 0.f;
                for (int c = 0; c < d; ++c) acc = fmaf(da[c], v[c], acc);
                R5[i * lds + j] = acc;
            }
            __syncthreads();
            for (int i = warp; i < N; i += kBT / 32) {                 // dS = P * (dP - sum_j dP P)
                float dot = 0.f;
                for (int j = lane; j < N; j += 32) dot = fmaf(R5[i * lds + j], R4[i * lds + j], dot);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                for (int j = lane; j < N; j += 32) R5[i * lds + j] = R4[i * lds + j] * (R5[i * lds + j] - dot);
            }
            __syncthreads();
            // dq' = dS k ; dk = dS^T q'  -> R6 ; relative-position table gradient: one thread per table entry
            for (int e = tid; e < N * d; e += kBT) {
                const int i = e / d, c = e % d;
                float dq = 0.f, dk = 0.f;
                for (int j = 0; j < N; ++j) {
                    dq = fmaf(R5[i * lds + j], R3[j * ldh + d + c], dq);
                    dk = fmaf(R5[j * lds + i], R3[j * ldh + c], dk);
                }
                R6[i * ldh + c] = dq * scale;                         // gradient wrt the un-scaled q
                R6[i * ldh + d + c] = dk;
            }
            for (int idx = tid; idx < TBL; idx += kBT) {
                const int dy = idx / (2 * ws - 1) - (ws - 1), dx = idx % (2 * ws - 1) - (ws - 1);   // yi - yj, xi - xj
                float acc = 0.f;
                for (int yj = max(0, -dy); yj < min(ws, ws - dy); ++yj)
                    for (int xj = max(0, -dx); xj < min(ws, ws - dx); ++xj)
                        acc += R5[((yj + dy) * ws + xj + dx) * lds + yj * ws + xj];
                tacc[h * TBL + idx] += acc;
            }
            __syncthreads();
            for (int e = tid; e < N * 3 * d; e += kBT) {               // R6 -> scratch columns (part, h, c)
                const int n = e / (3 * d), j = e % (3 * d);
                dqkv_tok[(tok0 + n) * 3 * C + (j / d) * C + h * d + (j % d)] = R6[n * ldh + j];
            }
            __syncthreads();
        }

        // ---- dXw = dqkv Wqkv ; grad_x = grad_out + dXw at the un-shifted position
        __threadfence_block();
        float* DQ = smem;                                              // [N][3C+1] over R1..R6
        for (int e = tid; e < N * 3 * C; e += kBT) DQ[(e / (3 * C)) * ldq + e % (3 * C)] = dqkv_tok[tok0 * 3 * C + e];
        __syncthreads();
        // results staged in registers per work item and written straight to global (token rows are disjoint)
        gemm_rows<8>(DQ, ldq, qkv_w, C, nullptr, N, C, 3 * C, [&](int n, int o, float v) {
            const int64_t off = g.elem(win, b, wy, wx, n, o);
            gx[off] = g.tokens ? v : __ldg(gout + off) + v;
        });
        __syncthreads();
    }
    for (int i = tid; i < heads * TBL; i += kBT) {                     // table layout: [(2ws-1)^2][heads]
        const int h = i / TBL, idx = i % TBL;
        atomicAdd(gtable + idx * heads + h, tacc[i]);
    }
}


// ================================================================================================================
// GEMM-composed backward: the three token GEMMs of the backward (qkv = xw Wqkv^T + b, dAO = dy Wproj, dXw = dqkv Wqkv)
// are plain GEMMs over ALL tokens and run in a library GEMM on the caller's side; these kernels do the rest:
//   mwa_bwd_gather_kernel   x, grad_out (image layout) -> xw_tok, dy_tok (token-major, zeros for dropped windows), flags
//   mwa_bwd_core_kernel     per window and head, from qkv_tok and dao_tok: P, AO (-> ao_tok), dS, dq / dk / dv
//                           (-> dqkv_tok) and the relative-position table gradient
//   mwa_bwd_scatter_kernel  grad_x = grad_out + dxw_tok at the un-shifted pixel (dropped windows carry zeros)
// ================================================================================================================
__global__ void __launch_bounds__(256)
mwa_bwd_gather_kernel(const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ gout,
                      float* __restrict__ xw_tok, float* __restrict__ dy_tok, uint8_t* __restrict__ flags, BGeo g,
                      int nwin) {
    __shared__ float red[8];
    __shared__ int keep_s;
    const int C = g.C, N = g.N, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int win = blockIdx.x; win < nwin; win += gridDim.x) {
        const int b = win / (g.nwy * g.nwx), wy = (win / g.nwx) % g.nwy, wx = win % g.nwx;
        bool keep = true;
        if (alpha != nullptr) {
            float a = 0.f;
            for (int t = tid; t < N; t += 256) {
                int y, xx;
                g.token_pixel(wy, wx, t, y, xx);
                a += __ldg(alpha + (int64_t(b) * g.H + y) * g.W + xx);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) red[warp] = a;
            __syncthreads();
            if (tid == 0) {
                float tot = 0.f;
                for (int i = 0; i < 8; ++i) tot += red[i];
                keep_s = (tot != 0.f);
            }
            __syncthreads();
            keep = keep_s != 0;
            __syncthreads();
        }
        if (tid == 0) flags[win] = keep;
        const int64_t tok0 = int64_t(win) * N;
        for (int e = tid; e < N * C; e += 256) {
            // NCHW: pixel-fastest iteration coalesces the reads; NHWC: channel-fastest coalesces both sides
            const int c = g.channels_last ? e % C : e / N, t = g.channels_last ? e / C : e % N;
            float xv = 0.f, gv = 0.f;
            if (keep) {
                const int64_t off = g.elem(win, b, wy, wx, t, c);
                xv = __ldg(x + off);
                gv = __ldg(gout + off);
            }
            xw_tok[(tok0 + t) * C + c] = xv;
            dy_tok[(tok0 + t) * C + c] = gv;
        }
    }
}

__global__ void __launch_bounds__(256)
mwa_bwd_scatter_kernel(const float* __restrict__ gout, const float* __restrict__ dxw_tok, float* __restrict__ gx, BGeo g,
                       int nwin) {
    const int C = g.C, N = g.N, tid = threadIdx.x;
    for (int win = blockIdx.x; win < nwin; win += gridDim.x) {
        const int b = win / (g.nwy * g.nwx), wy = (win / g.nwx) % g.nwy, wx = win % g.nwx;
        const int64_t tok0 = int64_t(win) * N;
        for (int e = tid; e < N * C; e += 256) {
            const int c = g.channels_last ? e % C : e / N, t = g.channels_last ? e / C : e % N;
            const int64_t off = g.elem(win, b, wy, wx, t, c);
            gx[off] = __ldg(gout + off) + dxw_tok[(tok0 + t) * C + c];
        }
    }
}

// shared memory (floats): R3 [N][3d+1] q' k v | R2 [N][d+1] dAO_h | R4, R5 [N][N+1] | R6 [N][3d+1] | tacc [heads][TBL]
__host__ __device__ inline int64_t bwd_core_smem_floats(int C, int heads, int ws) {
    const int64_t N = ws * ws, d = C / heads, tbl = (2 * ws - 1) * (2 * ws - 1);
    return 2 * N * (3 * d + 1) + N * (d + 1) + 2 * N * (N + 1) + heads * tbl + 8;
}

// ---- register-tiled products for the 8 x 8-window case (N = 64), 256 threads.  The plain loops below issue two shared
//      loads per FMA and are LDS-bound; these tiles issue 0.5 (N x N outputs, 4 x 4 per thread) / 0.58 (N x D outputs,
//      4 x D/8 per thread on half of the CTA, so two independent products run side by side).
// out(i, j) = sum_c A[i][c] * B[j][c], i, j < 64: thread (ti, tj) owns rows ti + 16 a, columns tj + 16 b (rows of B at a
// stride of 3D+1 or D+1 floats fall into distinct banks; A is a broadcast)
template <int D, class Epi>
__device__ __forceinline__ void tile_nn64(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int tid,
                                          Epi epi) {
    const int tj = tid & 15, ti = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
        float av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = A[(ti + 16 * a) * lda + c];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = B[(tj + 16 * b) * ldb + c];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) epi(ti + 16 * a, tj + 16 * b, acc[a][b]);
}
// out(r, c) = sum_n P(r, n) * G[n][c]  (kTrans = false)   or   sum_n P(n, r) * G[n][c]  (kTrans = true),  r < 64, c < D;
// P has a row stride of 65 floats.  t = thread index inside a 128-thread half: rows rg + 16 a, columns cg + 8 cc.
template <int D, bool kTrans, class Epi>
__device__ __forceinline__ void tile_nd64(const float* __restrict__ P, const float* __restrict__ G, int ldg, int t, Epi epi) {
    constexpr int CT = D / 8, LDS_ = 65;
    const int cg = t & 7, rg = t >> 3;
    float acc[4][CT];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int cc = 0; cc < CT; ++cc) acc[a][cc] = 0.f;
#pragma unroll 4
    for (int n = 0; n < 64; ++n) {
        float pv[4], gv[CT];
#pragma unroll
        for (int a = 0; a < 4; ++a) pv[a] = kTrans ? P[n * LDS_ + rg + 16 * a] : P[(rg + 16 * a) * LDS_ + n];
#pragma unroll
        for (int cc = 0; cc < CT; ++cc) gv[cc] = G[n * ldg + cg + 8 * cc];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int cc = 0; cc < CT; ++cc) acc[a][cc] = fmaf(pv[a], gv[cc], acc[a][cc]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int cc = 0; cc < CT; ++cc) epi(rg + 16 * a, cg + 8 * cc, acc[a][cc]);
}

__global__ void __launch_bounds__(256)
mwa_bwd_core_kernel(const float* __restrict__ qkv_tok, const float* __restrict__ dao_tok, const uint8_t* __restrict__ blk,
                    const float* __restrict__ ext_mask, const uint8_t* __restrict__ flags, float* __restrict__ ao_tok,
                    float* __restrict__ dqkv_tok, float* __restrict__ gtable, BGeo g, int heads, int nwin) {
    extern __shared__ float smem[];
    constexpr int kT = 256;
    const int C = g.C, N = g.N, ws = g.ws, d = C / heads;
    const int ldh = 3 * d + 1, ld2 = d + 1, lds = N + 1, TBL = (2 * ws - 1) * (2 * ws - 1);
    float* R3 = smem;
    float* R2 = R3 + N * ldh;
    float* R4 = R2 + N * ld2;
    float* R5 = R4 + N * lds;
    float* R6 = R5 + N * lds;
    float* tacc = R6 + N * ldh;
    const MwaParamLayout L(C, heads, ws);
    const float scale = reinterpret_cast<const float*>(blk + L.header)[0];
    const float* bias = reinterpret_cast<const float*>(blk + L.bias);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fast = (N == 64 && (d == 24 || d == 32)) ? d : 0;        // register-tiled products (tile_nn64 / tile_nd64)
    for (int i = tid; i < heads * TBL; i += kT) tacc[i] = 0.f;
    __syncthreads();
    // work unit = (window, head): 8 x finer than a window, so that a few hundred windows still balance over the CTAs
    for (int64_t unit = blockIdx.x; unit < int64_t(nwin) * heads; unit += gridDim.x) {
        const int win = static_cast<int>(unit / heads), h = static_cast<int>(unit % heads);
        const int wy = (win / g.nwx) % g.nwy, wx = win % g.nwx;
        const int64_t tok0 = int64_t(win) * N;
        if (flags != nullptr && !flags[win]) {
            if (h == 0) {
                for (int e = tid; e < N * C; e += kT) ao_tok[tok0 * C + e] = 0.f;
                for (int e = tid; e < N * 3 * C; e += kT) dqkv_tok[tok0 * 3 * C + e] = 0.f;
            }
            continue;
        }
        {
            for (int e = tid; e < N * 3 * d; e += kT) {
                const int n = e / (3 * d), j = e % (3 * d);
                const float v = __ldg(qkv_tok + (tok0 + n) * 3 * C + (j / d) * C + h * d + (j % d));
                R3[n * ldh + j] = (j < d) ? v * scale : v;
            }
            for (int e = tid; e < N * d; e += kT) R2[(e / d) * ld2 + e % d] = __ldg(dao_tok + (tok0 + e / d) * C + h * d + e % d);
            __syncthreads();
            const float* bh = bias + int64_t(h) * N * N;
            auto logit = [&](int i, int j, float acc) {
                acc += __ldg(bh + i * N + j);
                if (g.tokens) {
                    if (g.mask_nw > 0) acc += __ldg(ext_mask + (int64_t(win % g.mask_nw) * N + i) * N + j);
                } else if (g.shift > 0 && g.region(wy, wx, i) != g.region(wy, wx, j)) {
                    acc += kNegMaskB;
                }
                R4[i * lds + j] = acc;
            };
            if (fast == 24) tile_nn64<24>(R3, ldh, R3 + d, ldh, tid, logit);
            else if (fast == 32) tile_nn64<32>(R3, ldh, R3 + d, ldh, tid, logit);
            else
                for (int e = tid; e < N * N; e += kT) {
                    const int i = e / N, j = e % N;
                    const float* q = R3 + i * ldh;
                    const float* k = R3 + j * ldh + d;
                    float acc = 0.f;
                    for (int c = 0; c < d; ++c) acc = fmaf(q[c], k[c], acc);
                    logit(i, j, acc);
                }
            __syncthreads();
            for (int i = warp; i < N; i += kT / 32) {
                float* row = R4 + i * lds;
                float m = -INFINITY;
                for (int j = lane; j < N; j += 32) m = fmaxf(m, row[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                float sum = 0.f;
                for (int j = lane; j < N; j += 32) {
                    const float e = expf(row[j] - m);
                    row[j] = e;
                    sum += e;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float inv = 1.f / sum;
                for (int j = lane; j < N; j += 32) row[j] *= inv;
            }
            __syncthreads();
            auto put_ao = [&](int i, int c, float v) { ao_tok[(tok0 + i) * C + h * d + c] = v; };
            auto put_dv = [&](int i, int c, float v) { R6[i * ldh + 2 * d + c] = v; };
            if (fast) {                 // warps 0-3: AO = P v, warps 4-7: dv = P^T dAO (independent, both read P)
                if (tid < 128) {
                    if (fast == 24) tile_nd64<24, false>(R4, R3 + 2 * d, ldh, tid, put_ao);
                    else tile_nd64<32, false>(R4, R3 + 2 * d, ldh, tid, put_ao);
                } else {
                    if (fast == 24) tile_nd64<24, true>(R4, R2, ld2, tid - 128, put_dv);
                    else tile_nd64<32, true>(R4, R2, ld2, tid - 128, put_dv);
                }
            } else
            for (int e = tid; e < N * d; e += kT) {
                const int i = e / d, c = e % d;
                const float* p = R4 + i * lds;
                float acc = 0.f;
                for (int j = 0; j < N; ++j) acc = fmaf(p[j], R3[j * ldh + 2 * d + c], acc);
                put_ao(i, c, acc);
                float dv = 0.f;
                for (int n = 0; n < N; ++n) dv = fmaf(R4[n * lds + i], R2[n * ld2 + c], dv);
                put_dv(i, c, dv);
            }
            auto put_dp = [&](int i, int j, float v) { R5[i * lds + j] = v; };
            if (fast == 24) tile_nn64<24>(R2, ld2, R3 + 2 * d, ldh, tid, put_dp);
            else if (fast == 32) tile_nn64<32>(R2, ld2, R3 + 2 * d, ldh, tid, put_dp);
            else
            for (int e = tid; e < N * N; e += kT) {
                const int i = e / N, j = e % N;
                const float* da = R2 + i * ld2;
                const float* v = R3 + j * ldh + 2 * d;
                float acc = 0.f;
                for (int c = 0; c < d; ++c) acc = fmaf(da[c], v[c], acc);
                put_dp(i, j, acc);
            }
            __syncthreads();
            for (int i = warp; i < N; i += kT / 32) {
                float dot = 0.f;
                for (int j = lane; j < N; j += 32) dot = fmaf(R5[i * lds + j], R4[i * lds + j], dot);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                for (int j = lane; j < N; j += 32) R5[i * lds + j] = R4[i * lds + j] * (R5[i * lds + j] - dot);
            }
            __syncthreads();
            auto put_dq = [&](int i, int c, float v) { R6[i * ldh + c] = v * scale; };
            auto put_dk = [&](int i, int c, float v) { R6[i * ldh + d + c] = v; };
            if (fast) {                 // warps 0-3: dq = dS k, warps 4-7: dk = dS^T q'
                if (tid < 128) {
                    if (fast == 24) tile_nd64<24, false>(R5, R3 + d, ldh, tid, put_dq);
                    else tile_nd64<32, false>(R5, R3 + d, ldh, tid, put_dq);
                } else {
                    if (fast == 24) tile_nd64<24, true>(R5, R3, ldh, tid - 128, put_dk);
                    else tile_nd64<32, true>(R5, R3, ldh, tid - 128, put_dk);
                }
            } else
            for (int e = tid; e < N * d; e += kT) {
                const int i = e / d, c = e % d;
                float dq = 0.f, dk = 0.f;
                for (int j = 0; j < N; ++j) {
                    dq = fmaf(R5[i * lds + j], R3[j * ldh + d + c], dq);
                    dk = fmaf(R5[j * lds + i], R3[j * ldh + c], dk);
                }
                put_dq(i, c, dq);
                put_dk(i, c, dk);
            }
            for (int idx = tid; idx < TBL; idx += kT) {
                const int dy = idx / (2 * ws - 1) - (ws - 1), dx = idx % (2 * ws - 1) - (ws - 1);
                float acc = 0.f;
                for (int yj = max(0, -dy); yj < min(ws, ws - dy); ++yj)
                    for (int xj = max(0, -dx); xj < min(ws, ws - dx); ++xj)
                        acc += R5[((yj + dy) * ws + xj + dx) * lds + yj * ws + xj];
                tacc[h * TBL + idx] += acc;
            }
            __syncthreads();
            for (int e = tid; e < N * 3 * d; e += kT) {
                const int n = e / (3 * d), j = e % (3 * d);
                dqkv_tok[(tok0 + n) * 3 * C + (j / d) * C + h * d + (j % d)] = R6[n * ldh + j];
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < heads * TBL; i += kT) atomicAdd(gtable + (i % TBL) * heads + i / TBL, tacc[i]);
}

}  // namespace
}  // namespace b200

using namespace b200;

extern "C" {

static int launch_bwd(const float* x, const float* alpha, const float* gout, const float* qkv_w, const float* proj_w,
                      const void* params, const float* ext_mask, float* gx, float* gtable, float* xw_tok,
                      float* ao_tok, float* dy_tok, float* dqkv_tok, BGeo g, int heads, int64_t nwin64,
                      cudaStream_t st) {
    if (nwin64 > 0x7fffffffll) return MWA_ERR_UNSUPPORTED;
    const int nwin = static_cast<int>(nwin64);
    const int TBL = (2 * g.ws - 1) * (2 * g.ws - 1);
    MWA_TRY_CUDA(cudaMemsetAsync(gtable, 0, sizeof(float) * TBL * heads, st), "mwa_backward(memset)");
    if (nwin == 0) return MWA_OK;
    const int64_t smem = 4 * bwd_smem_floats(g.C, heads, g.ws);
    if (g.N > 64 || g.N % 4 != 0 || smem > 227 * 1024 - 64) return MWA_ERR_UNSUPPORTED;
    MWA_TRY_CUDA(cudaFuncSetAttribute(mwa_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)),
                 "mwa_backward(attr)");
    const int grid = nwin < kNumSMs ? nwin : kNumSMs;
    mwa_bwd_kernel<<<grid, kBT, smem, st>>>(x, alpha, gout, qkv_w, proj_w, static_cast<const uint8_t*>(params),
                                            ext_mask, gx, gtable, xw_tok, ao_tok, dy_tok, dqkv_tok, g, heads, nwin);
    return check_launch("mwa_backward");
}

int mwa_backward(const float* x, const float* alpha, const float* grad_out, const float* qkv_w, const float* proj_w,
                 const void* params, float* grad_x, float* grad_table, float* xw_tok, float* ao_tok, float* dy_tok,
                 float* dqkv_tok, int B, int C, int H, int W, int heads, int ws, int shift, int channels_last,
                 void* stream) {
    if (!x || !grad_out || !qkv_w || !proj_w || !params || !grad_x || !grad_table || !xw_tok || !ao_tok || !dy_tok ||
        !dqkv_tok)
        return MWA_ERR_INVALID;
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || heads <= 0 || ws <= 0 || C % heads != 0) return MWA_ERR_INVALID;
    if (shift < 0 || shift >= ws || H % ws != 0 || W % ws != 0) return MWA_ERR_INVALID;
    BGeo g{B, C, H, W, ws, shift, W / ws, H / ws, ws * ws, channels_last, 0, 0};
    return launch_bwd(x, alpha, grad_out, qkv_w, proj_w, params, nullptr, grad_x, grad_table, xw_tok, ao_tok, dy_tok,
                      dqkv_tok, g, heads, int64_t(B) * g.nwx * g.nwy, static_cast<cudaStream_t>(stream));
}

int window_attention_backward(const float* xw, const float* mask, const float* grad_out, const float* qkv_w,
                              const float* proj_w, const void* params, float* grad_xw, float* grad_table,
                              float* ao_tok, float* dqkv_tok, int64_t K, int C, int heads, int ws, int mask_windows,
                              void* stream) {
    if (!xw || !grad_out || !qkv_w || !proj_w || !params || !grad_xw || !grad_table || !ao_tok || !dqkv_tok)
        return MWA_ERR_INVALID;
    if (K < 0 || C <= 0 || heads <= 0 || ws <= 0 || C % heads != 0 || mask_windows < 0) return MWA_ERR_INVALID;
    if (mask_windows > 0 && (!mask || K % mask_windows != 0)) return MWA_ERR_INVALID;
    BGeo g{static_cast<int>(K), C, ws, ws, ws, 0, 1, 1, ws * ws, 1, 1, mask_windows};
    // token mode: xw / grad_out already ARE the token-major tensors the caller's GEMMs need -> no copies of them
    return launch_bwd(xw, nullptr, grad_out, qkv_w, proj_w, params, mask, grad_xw, grad_table, /*xw_tok=*/nullptr,
                      ao_tok, /*dy_tok=*/nullptr, dqkv_tok, g, heads, K, static_cast<cudaStream_t>(stream));
}

int mwa_bwd_gather(const float* x, const float* alpha, const float* grad_out, float* xw_tok, float* dy_tok,
                   uint8_t* keep_flags, int B, int C, int H, int W, int ws, int shift, int channels_last, void* stream) {
    if (!x || !grad_out || !xw_tok || !dy_tok || !keep_flags) return MWA_ERR_INVALID;
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || ws <= 0 || shift < 0 || shift >= ws || H % ws || W % ws) return MWA_ERR_INVALID;
    BGeo g{B, C, H, W, ws, shift, W / ws, H / ws, ws * ws, channels_last, 0, 0};
    const int64_t nwin = int64_t(B) * g.nwx * g.nwy;
    if (nwin > 0x7fffffffll) return MWA_ERR_UNSUPPORTED;
    if (nwin == 0) return MWA_OK;
    const int grid = static_cast<int>(nwin < kNumSMs * 8 ? nwin : kNumSMs * 8);
    mwa_bwd_gather_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, alpha, grad_out, xw_tok, dy_tok,
                                                                               keep_flags, g, static_cast<int>(nwin));
    return check_launch("mwa_bwd_gather");
}

int mwa_bwd_core(const float* qkv_tok, const float* dao_tok, const void* params, const float* mask,
                 const uint8_t* keep_flags, float* ao_tok, float* dqkv_tok, float* grad_table, int64_t nwin, int C,
                 int H, int W, int heads, int ws, int shift, int mask_windows, void* stream) {
    if (!qkv_tok || !dao_tok || !params || !ao_tok || !dqkv_tok || !grad_table) return MWA_ERR_INVALID;
    if (nwin < 0 || C <= 0 || heads <= 0 || ws <= 0 || C % heads != 0 || mask_windows < 0) return MWA_ERR_INVALID;
    if (nwin > 0x7fffffffll) return MWA_ERR_UNSUPPORTED;
    const bool tokens = H <= 0;                                   // token mode: no image geometry, optional mask
    if (tokens && mask_windows > 0 && (!mask || nwin % mask_windows != 0)) return MWA_ERR_INVALID;
    if (!tokens && (W <= 0 || shift < 0 || shift >= ws || H % ws || W % ws)) return MWA_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int TBL = (2 * ws - 1) * (2 * ws - 1);
    MWA_TRY_CUDA(cudaMemsetAsync(grad_table, 0, sizeof(float) * TBL * heads, st), "mwa_bwd_core(memset)");
    if (nwin == 0) return MWA_OK;
    BGeo g = tokens ? BGeo{static_cast<int>(nwin), C, ws, ws, ws, 0, 1, 1, ws * ws, 1, 1, mask_windows}
                    : BGeo{0, C, H, W, ws, shift, W / ws, H / ws, ws * ws, 0, 0, 0};
    const int64_t smem = 4 * bwd_core_smem_floats(C, heads, ws);
    if (g.N > 64 || smem > 227 * 1024 - 64) return MWA_ERR_UNSUPPORTED;
    MWA_TRY_CUDA(cudaFuncSetAttribute(mwa_bwd_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)),
                 "mwa_bwd_core(attr)");
    const int per_sm = smem > 0 ? static_cast<int>((220 * 1024) / smem) : 1;
    const int64_t cap = int64_t(kNumSMs) * (per_sm < 1 ? 1 : per_sm > 6 ? 6 : per_sm);
    const int64_t units = nwin * heads;
    const int grid = static_cast<int>(units < cap ? units : cap);
    mwa_bwd_core_kernel<<<grid, 256, smem, st>>>(qkv_tok, dao_tok, static_cast<const uint8_t*>(params), mask, keep_flags,
                                                 ao_tok, dqkv_tok, grad_table, g, heads, static_cast<int>(nwin));
    return check_launch("mwa_bwd_core");
}

int mwa_bwd_scatter(const float* grad_out, const float* dxw_tok, float* grad_x, int B, int C, int H, int W, int ws,
                    int shift, int channels_last, void* stream) {
    if (!grad_out || !dxw_tok || !grad_x) return MWA_ERR_INVALID;
    if (B < 0 || C <= 0 || H <= 0 || W <= 0 || ws <= 0 || shift < 0 || shift >= ws || H % ws || W % ws) return MWA_ERR_INVALID;
    BGeo g{B, C, H, W, ws, shift, W / ws, H / ws, ws * ws, channels_last, 0, 0};
    const int64_t nwin = int64_t(B) * g.nwx * g.nwy;
    if (nwin > 0x7fffffffll) return MWA_ERR_UNSUPPORTED;
    if (nwin == 0) return MWA_OK;
    const int grid = static_cast<int>(nwin < kNumSMs * 8 ? nwin : kNumSMs * 8);
    mwa_bwd_scatter_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, dxw_tok, grad_x, g,
                                                                                static_cast<int>(nwin));
    return check_launch("mwa_bwd_scatter");
}

}  // extern "C"
