// inbatch_ce.cuh — K4/K5: in-batch-negative score matrix + cross entropy, forward and backward.
//
// Replaces `logits = x @ y.T; F.cross_entropy(logits, target)` of SimpleContrastiveLoss.forward
// (DRT/trainer/losses.py:16-17) and the loss block of DRModel.forward
// (DRT/model/biencoder.py:107-116), plus their autograd.
//
// The reference computes this in fp32 (TF32 is off by default for torch.matmul), and the loss
// must match to 1e-4 relative, so the contraction runs as fp32 FFMA (a bf16 tensor-core pass
// would put ~5e-4 relative error on the loss).  At the named shape (128 x 1024 x 768,
// 0.2 GFLOP, 3.5 MB of operands) the step is launch-latency bound, not FLOP bound; the win
// over the eager path is one launch for scores + log-sum-exp + NLL (and no HBM round trip of
// the score matrix unless the caller asks for it), and two launches for the backward.
#pragma once
#include <cfloat>
#include <cstdint>

namespace drt {

constexpr int kCeTM = 32;    // logits tile rows  (queries)
constexpr int kCeTN = 32;    // logits tile cols  (passages)
constexpr int kCeTK = 32;    // k-step
constexpr int kCeThreads = 256;

// C tile [32x32] = X[m0:m0+32, :] · Y[n0:n0+32, :]^T, fp32.  Thread t computes a 1x4 strip:
// row = t / 8, cols = 4*(t % 8) .. +3.  Results returned in acc[4].
__device__ __forceinline__ void ce_tile_logits(const float* __restrict__ x, const float* __restrict__ y,
                                               long long B, long long P, int dim, long long m0,
                                               long long n0, float (&sx)[kCeTK][kCeTM + 1],
                                               float (&sy)[kCeTK][kCeTN + 1], float (&acc)[4]) {
    const int t = threadIdx.x;
    const int r = t >> 3, c4 = (t & 7) * 4;
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    // loader mapping: 256 threads load 32 rows x 32 k of X and of Y (one float4 each)
    const int lr = t >> 3, lk = (t & 7) * 4;
    for (int k0 = 0; k0 < dim; k0 += kCeTK) {
        float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vy = vx;
        if (m0 + lr < B) vx = *reinterpret_cast<const float4*>(x + (m0 + lr) * dim + k0 + lk);
        if (n0 + lr < P) vy = *reinterpret_cast<const float4*>(y + (n0 + lr) * dim + k0 + lk);
        __syncthreads();
        sx[lk + 0][lr] = vx.x; sx[lk + 1][lr] = vx.y; sx[lk + 2][lr] = vx.z; sx[lk + 3][lr] = vx.w;
        sy[lk + 0][lr] = vy.x; sy[lk + 1][lr] = vy.y; sy[lk + 2][lr] = vy.z; sy[lk + 3][lr] = vy.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kCeTK; ++kk) {
            const float a = sx[kk][r];
            acc[0] = fmaf(a, sy[kk][c4 + 0], acc[0]);
            acc[1] = fmaf(a, sy[kk][c4 + 1], acc[1]);
            acc[2] = fmaf(a, sy[kk][c4 + 2], acc[2]);
            acc[3] = fmaf(a, sy[kk][c4 + 3], acc[3]);
        }
    }
}

// Forward: grid (ceil(P/32), ceil(B/32)).  Each CTA produces per-row partial (max, sum-exp) over
// its 32 columns; the last CTA to finish (atomic ticket) folds the partials into lse / per-row
// loss / the scaled total in a fixed order, so the result is deterministic.
__global__ void __launch_bounds__(kCeThreads)
inbatch_ce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, long long B,
                      long long P, int dim, const long long* __restrict__ target,
                      long long target_stride, float loss_scale, float* logits_out,
                      float* part_max, float* part_sum, float* tgt_logit, unsigned int* ticket,
                      float* lse_out, float* loss_rows, float* loss_out) {
    __shared__ float sx[kCeTK][kCeTM + 1];
    __shared__ float sy[kCeTK][kCeTN + 1];
    __shared__ bool s_last;
    const long long m0 = static_cast<long long>(blockIdx.y) * kCeTM;
    const long long n0 = static_cast<long long>(blockIdx.x) * kCeTN;
    float acc[4];
    ce_tile_logits(x, y, B, P, dim, m0, n0, sx, sy, acc);

    const int t = threadIdx.x, r = t >> 3, c4 = (t & 7) * 4;
    const long long row = m0 + r;
    const long long tcol = (row < B) ? (target ? target[row] : row * target_stride) : -1;
    float mx = -FLT_MAX;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long col = n0 + c4 + j;
        if (row < B && col < P) {
            if (logits_out) logits_out[row * P + col] = acc[j];
            if (col == tcol) tgt_logit[row] = acc[j];
            mx = fmaxf(mx, acc[j]);
        }
    }
    // reduce over the 8 threads that share a row (consecutive lanes)
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (row < B && n0 + c4 + j < P) se += expf(acc[j] - mx);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if ((t & 7) == 0 && row < B) {
        part_max[row * gridDim.x + blockIdx.x] = mx;
        part_sum[row * gridDim.x + blockIdx.x] = se;
    }
    // ---- last-CTA reduction ----
    __threadfence();
    __syncthreads();
    if (t == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        s_last = (atomicAdd(ticket, 1u) == total - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ double s_red[kCeThreads];
    double local = 0.0;
    const int ncol = gridDim.x;
    for (long long i = t; i < B; i += kCeThreads) {
        float m = -FLT_MAX;
        for (int c = 0; c < ncol; ++c) m = fmaxf(m, __ldcg(part_max + i * ncol + c));
        float s = 0.f;
        for (int c = 0; c < ncol; ++c) s += __ldcg(part_sum + i * ncol + c) * expf(__ldcg(part_max + i * ncol + c) - m);
        const float lse = m + logf(s);
        const float li = lse - __ldcg(tgt_logit + i);
        lse_out[i] = lse;
        loss_rows[i] = li;
        local += static_cast<double>(li);
    }
    s_red[t] = local;
    __syncthreads();
    for (int o = kCeThreads / 2; o > 0; o >>= 1) {
        if (t < o) s_red[t] += s_red[t + o];
        __syncthreads();
    }
    if (t == 0) {
        *loss_out = static_cast<float>(s_red[0] * static_cast<double>(loss_scale));
        *ticket = 0u;    // re-arm for the next launch on this stream
    }
}

// Backward step 1: dlogits[i,j] = g_i * (exp(logit_ij - lse_i) - [j == target_i]); logits are
// recomputed tile by tile (they were never stored).
__global__ void __launch_bounds__(kCeThreads)
inbatch_ce_dlogits_kernel(const float* __restrict__ x, const float* __restrict__ y, long long B,
                          long long P, int dim, const long long* __restrict__ target,
                          long long target_stride, const float* __restrict__ lse,
                          const float* __restrict__ grad_rows, float* __restrict__ dlogits) {
    __shared__ float sx[kCeTK][kCeTM + 1];
    __shared__ float sy[kCeTK][kCeTN + 1];
    const long long m0 = static_cast<long long>(blockIdx.y) * kCeTM;
    const long long n0 = static_cast<long long>(blockIdx.x) * kCeTN;
    float acc[4];
    ce_tile_logits(x, y, B, P, dim, m0, n0, sx, sy, acc);
    const int t = threadIdx.x, r = t >> 3, c4 = (t & 7) * 4;
    const long long row = m0 + r;
    if (row >= B) return;
    const long long tcol = target ? target[row] : row * target_stride;
    const float l = lse[row], g = grad_rows[row];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const long long col = n0 + c4 + j;
        if (col < P) dlogits[row * P + col] = g * (expf(acc[j] - l) - (col == tcol ? 1.f : 0.f));
    }
}

// Backward step 2: generic fp32 C[M,N] = sum_k A(m,k) * Bm[k,N] with A(m,k) = A[m*sam + k*sak].
//   dx = dlogits · y      : A = dlogits (sam = P, sak = 1), Bm = y, M = B, K = P
//   dy = dlogitsᵀ · x     : A = dlogits (sam = 1, sak = P), Bm = x, M = P, K = B
// 64x64 tiles, 256 threads, 4x4 outputs per thread.
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, long long sam, long long sak,
                     const float* __restrict__ Bm, long long M, long long N, long long K,
                     float* __restrict__ C) {
    __shared__ float sa[16][64 + 1];
    __shared__ float sb[16][64 + 4];
    const int t = threadIdx.x;
    const long long m0 = static_cast<long long>(blockIdx.y) * 64, n0 = static_cast<long long>(blockIdx.x) * 64;
    const int tr = (t >> 4) * 4, tc = (t & 15) * 4;
    float acc[4][4] = {};
    for (long long k0 = 0; k0 < K; k0 += 16) {
        __syncthreads();
        // A tile: 64 rows x 16 k.  Pick the loader orientation that is contiguous in memory.
        for (int i = t; i < 64 * 16; i += 256) {
            int mm, kk;
            if (sak == 1) { mm = i >> 4; kk = i & 15; } else { kk = i >> 6; mm = i & 63; }
            const long long m = m0 + mm, k = k0 + kk;
            sa[kk][mm] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f;
        }
        for (int i = t; i < 16 * 64; i += 256) {
            const int kk = i >> 6, nn = i & 63;
            const long long k = k0 + kk, n = n0 + nn;
            sb[kk][nn] = (k < K && n < N) ? Bm[k * N + n] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sa[kk][tr + i]; b[i] = sb[kk][tc + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long m = m0 + tr + i, n = n0 + tc + j;
            if (m < M && n < N) C[m * N + n] = acc[i][j];
        }
}

}  // namespace drt
