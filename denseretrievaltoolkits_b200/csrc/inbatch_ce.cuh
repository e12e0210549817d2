// inbatch_ce.cuh — the fp32 SIMT form of K4/K5: in-batch-negative score matrix + cross entropy,
// forward and backward, for the shapes the tensor-core kernels do not take (a contraction length
// that is not a multiple of 4: the split operand's row pitch must be 16-byte aligned), and the
// register-tiled SIMT GEMM core that the last-resort exact search pass (exact_filter_kernel)
// shares.  The reference's own shapes run on tcgen05: gemm_tc_small.cuh / gemm_tc.cuh.
//
// Replaces `logits = x @ y.T; F.cross_entropy(logits, target)` of SimpleContrastiveLoss.forward
// (DRT/trainer/losses.py:16-17) and the loss block of DRModel.forward
// (DRT/model/biencoder.py:107-116), plus their autograd.
//
// One register-tiled SIMT GEMM core serves all three contractions:
//   NT  logits  = x · yᵀ          A(m,k)=x[m*d+k]    B(k,n)=y[n*d+k]
//   NN  dx      = dlogits · y      A(m,k)=dL[m*P+k]   B(k,n)=y[k*d+n]
//   TN  dy      = dlogitsᵀ · x     A(m,k)=dL[k*P+m]   B(k,n)=x[k*d+n]
// with three tile shapes chosen by how many tiles the output has (every SM should get a CTA):
// 32x32 / 128 threads, 64x64 / 256 threads, 128x128 / 256 threads with 8x8 register tiles.
#pragma once
#include <cfloat>
#include <cstdint>

namespace drt {

struct GemmOperand {
    const float* p;
    long long s_outer;   // stride of the non-k index (m for A, n for B)
    long long s_k;       // stride of k
};

// Tile config: BM x BN outputs, TM x TN per thread, threads = (BM/TM) * (BN/TN), k-slab depth BK.
// The slab must be deep enough that its FMAs cover the L2 latency of the next slab's loads
// (with BK = 16 a 32x32 tile spent ~85 % of its time waiting for them).
template <int BM_, int BN_, int TM_, int TN_, int BK_>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, TM = TM_, TN = TN_, BK = BK_;
    static constexpr int TX = BN / TN, TY = BM / TM, THREADS = TX * TY;
    static constexpr int PAD = 4;
    static constexpr int VA = BM * BK / 4 / THREADS, VB = BN * BK / 4 / THREADS;   // float4 per thread per slab
    static_assert(VA * 4 * THREADS == BM * BK && VB * 4 * THREADS == BN * BK, "slab must split into whole float4s");
    // Tile-local row / column of the thread's i-th / j-th output.  With 8-wide register tiles the
    // eight outputs are two groups of four, BM/2 (BN/2) apart, so that the lanes of a quarter
    // warp read CONSECUTIVE float4s of the smem fragment (a contiguous 8-float strip per lane
    // would make every LDS.128 a 2-way bank conflict).
    __host__ __device__ static constexpr int row_of(int ty, int i) {
        return TM == 8 ? (i >> 2) * (BM / 2) + ty * 4 + (i & 3) : ty * TM + i;
    }
    __host__ __device__ static constexpr int col_of(int tx, int j) {
        return TN == 8 ? (j >> 2) * (BN / 2) + tx * 4 + (j & 3) : tx * TN + j;
    }
};
using GemmSmall = GemmCfg<32, 32, 4, 2, 64>;    // 128 threads, 36 KB smem
using GemmLarge = GemmCfg<64, 64, 4, 4, 32>;    // 256 threads, 34 KB smem
using GemmHuge = GemmCfg<128, 128, 8, 8, 16>;   // 256 threads, 34 KB smem: 64 FMAs per 4 LDS.128

// Vector v (of NV per thread) of a [ROWS x BK] operand tile.  Element (r, k) lives at
// p[(r0 + r) * s_outer + (k0 + k) * s_k]; out-of-range -> 0.  fetch issues the global load into
// registers; store writes it into smem laid out [k][row] (row contiguous) -- split so the load
// latency overlaps the FMAs of the current slab.
template <int ROWS, int BK, int THREADS>
__device__ __forceinline__ void fetch_vec(float (&v)[4], int vi, const GemmOperand& op, long long r0, long long nrows,
                                          long long k0, long long K, bool vec_ok) {
    const int i = vi * THREADS + threadIdx.x;
    v[0] = v[1] = v[2] = v[3] = 0.f;
    if (op.s_k == 1) {            // k contiguous: 4 consecutive k of one row
        constexpr int VEC_PER_ROW = BK / 4;
        const int r = i / VEC_PER_ROW, kk = (i % VEC_PER_ROW) * 4;
        const long long gr = r0 + r, gk = k0 + kk;
        if (gr < nrows) {
            const float* src = op.p + gr * op.s_outer + gk;
            if (vec_ok && gk + 3 < K) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(src));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (gk + j < K) v[j] = src[j];
            }
        }
    } else {                      // row index contiguous: 4 consecutive rows of one k
        constexpr int VEC_PER_K = ROWS / 4;
        const int kk = i / VEC_PER_K, r = (i % VEC_PER_K) * 4;
        const long long gr = r0 + r, gk = k0 + kk;
        if (gk < K) {
            const float* src = op.p + gk * op.s_k + gr;
            if (vec_ok && gr + 3 < nrows) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(src));
                v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (gr + j < nrows) v[j] = src[j];
            }
        }
    }
}
template <int ROWS, int BK, int THREADS, int LD>
__device__ __forceinline__ void store_vec(float (*dst)[LD], const float (&v)[4], int vi, bool k_contiguous) {
    const int i = vi * THREADS + threadIdx.x;
    if (k_contiguous) {
        constexpr int VEC_PER_ROW = BK / 4;
        const int r = i / VEC_PER_ROW, kk = (i % VEC_PER_ROW) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[kk + j][r] = v[j];
    } else {
        constexpr int VEC_PER_K = ROWS / 4;
        const int kk = i / VEC_PER_K, r = (i % VEC_PER_K) * 4;
        *reinterpret_cast<float4*>(&dst[kk][r]) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// acc[TM][TN] = tile (m0.., n0..) of sum_k A(m,k) B(k,n).  Thread (ty, tx) owns rows
// m0 + ty*TM .. and columns n0 + tx*TN ..; lanes of a warp vary tx fastest, so the A fragment is
// a smem broadcast and the B fragment one contiguous wavefront.
template <class Cfg>
__device__ __forceinline__ void gemm_tile(const GemmOperand& A, const GemmOperand& B, long long M, long long N,
                                          long long K, long long m0, long long n0, bool vecA, bool vecB,
                                          float (&acc)[Cfg::TM][Cfg::TN]) {
    constexpr int BM = Cfg::BM, BN = Cfg::BN, TM = Cfg::TM, TN = Cfg::TN, BK = Cfg::BK, T = Cfg::THREADS;
    __shared__ __align__(16) float sa[2][BK][BM + Cfg::PAD];
    __shared__ __align__(16) float sb[2][BK][BN + Cfg::PAD];
    const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
    const bool a_kc = A.s_k == 1, b_kc = B.s_k == 1;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float ra[Cfg::VA][4], rb[Cfg::VB][4];
#pragma unroll
    for (int v = 0; v < Cfg::VA; ++v) { fetch_vec<BM, BK, T>(ra[v], v, A, m0, M, 0, K, vecA); store_vec<BM, BK, T>(sa[0], ra[v], v, a_kc); }
#pragma unroll
    for (int v = 0; v < Cfg::VB; ++v) { fetch_vec<BN, BK, T>(rb[v], v, B, n0, N, 0, K, vecB); store_vec<BN, BK, T>(sb[0], rb[v], v, b_kc); }
    __syncthreads();
    int buf = 0;
    for (long long k0 = 0; k0 < K; k0 += BK) {
        const bool more = k0 + BK < K;
        if (more) {               // global loads of the next slab fly while this slab is multiplied
#pragma unroll
            for (int v = 0; v < Cfg::VA; ++v) fetch_vec<BM, BK, T>(ra[v], v, A, m0, M, k0 + BK, K, vecA);
#pragma unroll
            for (int v = 0; v < Cfg::VB; ++v) fetch_vec<BN, BK, T>(rb[v], v, B, n0, N, k0 + BK, K, vecB);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int v = 0; v < TM / 4; ++v) {
                const float4 a4 = *reinterpret_cast<const float4*>(&sa[buf][kk][Cfg::row_of(ty, 4 * v)]);
                a[4 * v] = a4.x; a[4 * v + 1] = a4.y; a[4 * v + 2] = a4.z; a[4 * v + 3] = a4.w;
            }
            if constexpr (TN % 4 == 0) {
#pragma unroll
                for (int v = 0; v < TN / 4; ++v) {
                    const float4 b4 = *reinterpret_cast<const float4*>(&sb[buf][kk][Cfg::col_of(tx, 4 * v)]);
                    b[4 * v] = b4.x; b[4 * v + 1] = b4.y; b[4 * v + 2] = b4.z; b[4 * v + 3] = b4.w;
                }
            } else {
                const float2 b2 = *reinterpret_cast<const float2*>(&sb[buf][kk][tx * TN]);
                b[0] = b2.x; b[1] = b2.y;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
#pragma unroll
            for (int v = 0; v < Cfg::VA; ++v) store_vec<BM, BK, T>(sa[buf ^ 1], ra[v], v, a_kc);
#pragma unroll
            for (int v = 0; v < Cfg::VB; ++v) store_vec<BN, BK, T>(sb[buf ^ 1], rb[v], v, b_kc);
        }
        __syncthreads();
        buf ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------
// Forward: grid (ceil(P/BN), ceil(B/BM)).  Each CTA produces per-row partial (max, sum-exp) over
// its BN columns; the last CTA to finish (atomic ticket) folds the partials into lse / per-row
// loss / the scaled total in a fixed order, so the result is deterministic.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS)
inbatch_ce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, long long B,
                      long long P, int dim, const long long* __restrict__ target,
                      long long target_stride, float loss_scale, float* logits_out,
                      float* part_max, float* part_sum, float* tgt_logit, unsigned int* ticket,
                      float* lse_out, float* loss_rows, float* loss_out, int vec_ok) {
    constexpr int TM = Cfg::TM, TN = Cfg::TN;
    __shared__ bool s_last;
    const long long m0 = static_cast<long long>(blockIdx.y) * Cfg::BM;
    const long long n0 = static_cast<long long>(blockIdx.x) * Cfg::BN;
    float acc[TM][TN];
    gemm_tile<Cfg>(GemmOperand{x, dim, 1}, GemmOperand{y, dim, 1}, B, P, dim, m0, n0, vec_ok, vec_ok, acc);

    const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long row = m0 + Cfg::row_of(ty, i);
        const long long tcol = (row < B) ? (target ? target[row] : row * target_stride) : -1;
        float mx = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const long long col = n0 + Cfg::col_of(tx, j);
            if (row < B && col < P) {
                if (logits_out) logits_out[row * P + col] = acc[i][j];
                if (col == tcol) tgt_logit[row] = acc[i][j];
                mx = fmaxf(mx, acc[i][j]);
            }
        }
        // the TX threads sharing this row are consecutive lanes (TX = 16 divides the warp)
#pragma unroll
        for (int o = Cfg::TX / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < TN; ++j)
            if (row < B && n0 + Cfg::col_of(tx, j) < P) se += expf(acc[i][j] - mx);
#pragma unroll
        for (int o = Cfg::TX / 2; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        if (tx == 0 && row < B) {
            part_max[row * gridDim.x + blockIdx.x] = mx;
            part_sum[row * gridDim.x + blockIdx.x] = se;
        }
    }
    // ---- last-CTA reduction ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        s_last = (atomicAdd(ticket, 1u) == total - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ double s_red[Cfg::THREADS];
    double local = 0.0;
    const int ncol = gridDim.x;
    for (long long i = threadIdx.x; i < B; i += Cfg::THREADS) {
        float m = -FLT_MAX;
        for (int c0 = 0; c0 < ncol; c0 += 8) {      // 8 independent loads in flight per step
            float pm[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) pm[u] = (c0 + u < ncol) ? __ldcg(part_max + i * ncol + c0 + u) : -FLT_MAX;
#pragma unroll
            for (int u = 0; u < 8; ++u) m = fmaxf(m, pm[u]);
        }
        float s = 0.f;
        for (int c0 = 0; c0 < ncol; c0 += 8) {
            float pm[8], ps[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool in = c0 + u < ncol;
                pm[u] = in ? __ldcg(part_max + i * ncol + c0 + u) : -FLT_MAX;
                ps[u] = in ? __ldcg(part_sum + i * ncol + c0 + u) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) s += ps[u] * expf(pm[u] - m);
        }
        const float lse = m + logf(s);
        const long long tc = target ? target[i] : i * target_stride;
        // an out-of-range target poisons the loss instead of reading a stale logit
        const float li = (tc >= 0 && tc < P) ? lse - __ldcg(tgt_logit + i) : __int_as_float(0x7fc00000);
        lse_out[i] = lse;
        loss_rows[i] = li;
        local += static_cast<double>(li);
    }
    s_red[threadIdx.x] = local;
    __syncthreads();
    for (int o = Cfg::THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *loss_out = static_cast<float>(s_red[0] * static_cast<double>(loss_scale));
        *ticket = 0u;    // re-arm for the next launch on this stream
    }
}

// Backward step 1: dlogits[i,j] = g_i * (exp(logit_ij - lse_i) - [j == target_i]); logits are
// recomputed tile by tile (they were never stored).
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS)
inbatch_ce_dlogits_kernel(const float* __restrict__ x, const float* __restrict__ y, long long B,
                          long long P, int dim, const long long* __restrict__ target,
                          long long target_stride, const float* __restrict__ lse,
                          const float* __restrict__ grad_rows, int grad_stride, float grad_scale,
                          float* __restrict__ dlogits, int vec_ok) {
    constexpr int TM = Cfg::TM, TN = Cfg::TN;
    const long long m0 = static_cast<long long>(blockIdx.y) * Cfg::BM;
    const long long n0 = static_cast<long long>(blockIdx.x) * Cfg::BN;
    float acc[TM][TN];
    gemm_tile<Cfg>(GemmOperand{x, dim, 1}, GemmOperand{y, dim, 1}, B, P, dim, m0, n0, vec_ok, vec_ok, acc);
    const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long row = m0 + Cfg::row_of(ty, i);
        if (row >= B) continue;
        const long long tcol = target ? target[row] : row * target_stride;
        const float l = lse[row], g = grad_scale * grad_rows[row * grad_stride];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const long long col = n0 + Cfg::col_of(tx, j);
            if (col < P) dlogits[row * P + col] = g * (expf(acc[i][j] - l) - (col == tcol ? 1.f : 0.f));
        }
    }
}

// Backward step 1, when the forward kept its logits: one elementwise pass.
__global__ void ce_dlogits_from_logits_kernel(const float* logits, long long B, long long P,
                                              const long long* __restrict__ target, long long target_stride,
                                              const float* __restrict__ lse, const float* __restrict__ grad_rows,
                                              int grad_stride, float grad_scale, float* dlogits) {
    const long long total = B * P;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / P, col = i - row * P;
        const long long tcol = target ? target[row] : row * target_stride;
        const float g = grad_scale * grad_rows[row * grad_stride];
        dlogits[i] = g * (expf(logits[i] - lse[row]) - (col == tcol ? 1.f : 0.f));
    }
}

// Backward step 2: C[M,N] = sum_k A(m,k) B(k,n) with arbitrary operand strides (see header).
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS)
sgemm_kernel(GemmOperand A, GemmOperand Bm, long long M, long long N, long long K, float* __restrict__ C,
             int vecA, int vecB) {
    constexpr int TM = Cfg::TM, TN = Cfg::TN;
    const long long m0 = static_cast<long long>(blockIdx.y) * Cfg::BM;
    const long long n0 = static_cast<long long>(blockIdx.x) * Cfg::BN;
    float acc[TM][TN];
    gemm_tile<Cfg>(A, Bm, M, N, K, m0, n0, vecA, vecB, acc);
    const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const long long m = m0 + Cfg::row_of(ty, i), n = n0 + Cfg::col_of(tx, j);
            if (m < M && n < N) C[m * N + n] = acc[i][j];
        }
}

}  // namespace drt
