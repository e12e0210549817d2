// gemm_tc.cuh — fp32-accurate NT GEMM on tcgen05 for the LARGE in-batch-loss shapes
// (cross-device negatives: [B·W, P·W] logits, DRT/model/biencoder.py:103-107).
//
// C[M,N] (fp32, row-major) = A[M,K] · B[N,K]ᵀ with fp32 operands, computed on the bf16 tensor
// pipe by splitting every fp32 value exactly into three bf16 pieces (hi + mid + lo = x: 8+8+8
// mantissa bits) and contracting the six products that matter,
//     hi·hi + hi·mid + mid·hi + mid·mid + hi·lo + lo·hi        (dropped: ≤ 2^-24 relative),
// as ONE bf16 GEMM over a K axis of length 6K: the split kernels below lay the pieces out as
//     A' = [ mid | hi | lo | hi  | mid | hi ]      B' = [ mid | lo | hi | mid | hi | hi ]
// so A'·B'ᵀ is the sum above, accumulated in fp32 in TMEM.  bf16 x bf16 products are exact in
// fp32; the tensor core's fp32 accumulation truncates (~0.5 ulp of the current accumulator per
// MMA step), which is why the segments are ordered by ascending product magnitude: the
// accumulator stays ~2^-8 of its final size until the last (hi·hi) segment, so only K/16 steps
// truncate at full scale.  Measured on a 768-d contraction: 1.45e-4 abs at score scale 138
// (1e-6 relative; cuBLAS sgemm 1.9e-4; with hi·hi first 1.4e-3; one plain bf16 pass ~0.5).
//
// The kernel is K1's pipeline (TMA producer warp, single-thread tcgen05.mma issuer, two TMEM
// accumulator stages, 8 epilogue warps, optional CTA pairs) with a store epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cfloat>
#include "mips_filter.cuh"

namespace drt {

struct GemmTcParams {
    int num_m_tiles;      // ceil(M / (128 * kCtas))
    int num_n_tiles;      // ceil(N / 256)
    int unit_tiles;       // n-tiles per work unit
    int ksplit;           // split-K factor: unit = (m tile, n group, k chunk); chunk s writes C + s*M*ldc
    int num_k_blocks;     // K' / 64 (all chunks)
    long long M, N;
    long long ldc;
    float* C;             // output (may be NULL in CE mode: the logits are then never stored)
    int* err;
    // CE mode (ce_part_max != NULL, ksplit == 1): the epilogue thread that owns a row keeps an
    // online (max, sum-exp) over its 128 columns of the tile and picks the target logit, so the
    // cross entropy needs no second pass over the logits (ce_fold_kernel finishes per row)
    float* ce_part_max;   // [M][2 * num_n_tiles]
    float* ce_part_sum;
    float* ce_tgt_logit;  // [M]
    const long long* ce_target;
    long long ce_target_stride;
};

template <int kCtas>
struct GemmTcCfg {
    static constexpr int kStages = (kCtas == 1) ? 4 : 6;
    static constexpr uint32_t kABytes = kTileM * kBlockK * 2;
    static constexpr uint32_t kBRows = kTileN / kCtas;
    static constexpr uint32_t kBBytes = kBRows * kBlockK * 2;
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr uint32_t kBarBytes = 256;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
};

template <int kCtas>
__global__ void __launch_bounds__(kFilterThreads, 1)
gemm_tc_nt_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const GemmTcParams p) {
    using Cfg = GemmTcCfg<kCtas>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rank = (kCtas == 2) ? ptx::cluster_ctarank() : 0u;

    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + kStages * Cfg::kStageBytes;
    auto smem_a = [&](int s) { return base + s * Cfg::kStageBytes; };
    auto smem_b = [&](int s) { return base + s * Cfg::kStageBytes + Cfg::kABytes; };
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_holder = bars + 8u * (2 * kStages + 4);

    if constexpr (kCtas == 2) ptx::cluster_sync_all();
    if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmap_a); ptx::prefetch_tmap(&tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { ptx::mbar_init(full_bar(s), kCtas); ptx::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(tfull_bar(a), 1); ptx::mbar_init(tempty_bar(a), kEpilogueWarps * kCtas); }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<kCtas>(tmem_holder, 512);
    ptx::tc_fence_before();
    if constexpr (kCtas == 2) ptx::cluster_sync_all(); else __syncthreads();
    ptx::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_holder));

    const int cluster_id = blockIdx.x / kCtas;
    const int num_clusters = gridDim.x / kCtas;
    const int n_groups = (p.num_n_tiles + p.unit_tiles - 1) / p.unit_tiles;
    const int units_per_split = p.num_m_tiles * n_groups;
    const int num_units = units_per_split * p.ksplit;
    auto kb_begin_of = [&](int s) { return static_cast<int>(static_cast<long long>(p.num_k_blocks) * s / p.ksplit); };

    if (warp == 0) {
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int uu = cluster_id; uu < num_units; uu += num_clusters) {
                const int sp = uu / units_per_split, u = uu - sp * units_per_split;
                const int kb0 = kb_begin_of(sp), nkb = kb_begin_of(sp + 1) - kb0;
                const int m_tile = u % p.num_m_tiles;
                const int nt0 = (u / p.num_m_tiles) * p.unit_tiles;
                const int nt1 = min(nt0 + p.unit_tiles, p.num_n_tiles);
                const int a_row = (m_tile * kCtas + static_cast<int>(rank)) * kTileM;
                for (int nt = nt0; nt < nt1; ++nt) {
                    const int b_row = nt * kTileN + static_cast<int>(rank) * Cfg::kBRows;
                    for (int kb = 0; kb < nkb; ++kb) {
                        ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.err, 201);
                        if constexpr (kCtas == 1) {
                            ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
                            ptx::tma_load_2d(smem_a(stage), &tmap_a, full_bar(stage), (kb0 + kb) * kBlockK, a_row, ptx::kEvictNormal);
                            ptx::tma_load_2d(smem_b(stage), &tmap_b, full_bar(stage), (kb0 + kb) * kBlockK, b_row, ptx::kEvictNormal);
                        } else {
                            const uint32_t lead_full = ptx::mapa(full_bar(stage), 0);
                            ptx::tma_load_2d_pair(smem_a(stage), &tmap_a, lead_full, (kb0 + kb) * kBlockK, a_row, ptx::kEvictNormal);
                            ptx::tma_load_2d_pair(smem_b(stage), &tmap_b, lead_full, (kb0 + kb) * kBlockK, b_row, ptx::kEvictNormal);
                            if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
                            else           ptx::mbar_arrive_cluster(lead_full);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && ptx::elect_one()) {
            const uint32_t idesc = ptx::make_idesc_bf16(kTileM * kCtas, kTileN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int uu = cluster_id; uu < num_units; uu += num_clusters) {
                const int sp = uu / units_per_split, u = uu - sp * units_per_split;
                const int nkb = kb_begin_of(sp + 1) - kb_begin_of(sp);
                const int nt0 = (u / p.num_m_tiles) * p.unit_tiles;
                const int nt1 = min(nt0 + p.unit_tiles, p.num_n_tiles);
                for (int nt = nt0; nt < nt1; ++nt, ++it) {
                    const int acc = it & 1;
                    const uint32_t acc_phase = (it >> 1) & 1u;
                    ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err, 202);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * kTileN;
                    for (int kb = 0; kb < nkb; ++kb) {
                        ptx::mbar_wait(full_bar(stage), phase, p.err, 203);
                        ptx::tc_fence_after();
                        const uint64_t a_desc = ptx::make_kmajor_sw128_desc(smem_a(stage));
                        const uint64_t b_desc = ptx::make_kmajor_sw128_desc(smem_b(stage));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            ptx::umma_bf16<kCtas>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        ptx::umma_commit<kCtas>(empty_bar(stage));
                        if (kb == nkb - 1) ptx::umma_commit<kCtas>(tfull_bar(acc));
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else {
        const uint32_t quarter = warp & 3u;
        const uint32_t half = (warp - 2u) >> 2;
        constexpr int kColsPerWarp = kTileN / (kEpilogueWarps / 4);
        const uint32_t lead_tempty0 = (kCtas == 2) ? ptx::mapa(tempty_bar(0), 0) : tempty_bar(0);
        int it = 0;
        for (int uu = cluster_id; uu < num_units; uu += num_clusters) {
            const int sp = uu / units_per_split, u = uu - sp * units_per_split;
            const int m_tile = u % p.num_m_tiles;
            const int nt0 = (u / p.num_m_tiles) * p.unit_tiles;
            const int nt1 = min(nt0 + p.unit_tiles, p.num_n_tiles);
            const long long row = static_cast<long long>(m_tile * kCtas + static_cast<int>(rank)) * kTileM + quarter * 32 + lane;
            for (int nt = nt0; nt < nt1; ++nt, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1u;
                ptx::mbar_wait(tfull_bar(acc), acc_phase, p.err, 204);
                ptx::tc_fence_after();
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kTileN + half * kColsPerWarp;
                const long long col0 = static_cast<long long>(nt) * kTileN + half * kColsPerWarp;
                float* crow = p.C + (static_cast<long long>(sp) * p.M + row) * p.ldc;
                const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
                const bool ce = p.ce_part_max != nullptr;
                const long long tcol = (ce && row < p.M) ? (p.ce_target ? p.ce_target[row] : row * p.ce_target_stride) : -1;
                float run_m = -FLT_MAX, run_s = 0.f;
#pragma unroll 1
                for (int c = 0; c < kColsPerWarp / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr + c * 32, v);
                    tmem_ld_wait_regs(v);
                    const long long cc = col0 + c * 32;
                    if (ce && row < p.M) {
                        float cm = -FLT_MAX;
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (cc + j < p.N) cm = fmaxf(cm, __uint_as_float(v[j]));
                        const float nm = fmaxf(run_m, cm);
                        float cs = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (cc + j < p.N) cs += expf(__uint_as_float(v[j]) - nm);
                        run_s = run_s * expf(run_m - nm) + cs;
                        run_m = nm;
                        if (tcol >= cc && tcol < cc + 32) {
                            float t = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (cc + j == tcol) t = __uint_as_float(v[j]);
                            p.ce_tgt_logit[row] = t;
                        }
                    }
                    if (row < p.M && p.C) {
                        if (vec_ok && cc + 32 <= p.N) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(crow + cc + j) =
                                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (cc + j < p.N) crow[cc + j] = __uint_as_float(v[j]);
                        }
                    }
                }
                if (ce && row < p.M) {
                    const long long slot = row * (2ll * p.num_n_tiles) + 2 * nt + half;
                    p.ce_part_max[slot] = run_m;
                    p.ce_part_sum[slot] = run_s;
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kCtas == 1) ptx::mbar_arrive(tempty_bar(acc));
                    else if (rank == 0) ptx::mbar_arrive(tempty_bar(acc));
                    else ptx::mbar_arrive_cluster(lead_tempty0 + 8u * acc);
                }
            }
        }
    }
    __syncwarp();
    ptx::tc_fence_before();
    if constexpr (kCtas == 2) ptx::cluster_sync_all(); else __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<kCtas>(tmem_base, 512);
}

// ---- exact 3-way bf16 split of fp32 operands into the 6-segment K layout -------------------
// The tensor core's fp32 accumulation truncates, so every MMA step costs ~0.5 ulp of the CURRENT
// accumulator magnitude.  The six partial products are therefore laid out along K' in ascending
// magnitude (mid*mid, hi*lo, lo*hi, hi*mid, mid*hi, hi*hi): the accumulator stays ~2^-8 of its
// final size until the last segment, and only those K/16 steps truncate at full scale.
__device__ __forceinline__ void split3(float x, __nv_bfloat16& hi, __nv_bfloat16& mid, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);
    mid = __float2bfloat16_rn(r1);
    lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}

// src [R, C] fp32 row-major (row stride lds) -> dst [R, 6C] bf16.  is_b selects B' ordering.
__global__ void split3_rows_kernel(const float* __restrict__ src, long long R, long long C, long long lds,
                                   __nv_bfloat16* __restrict__ dst, int is_b) {
    const long long total = R * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / C, c = i - r * C;
        __nv_bfloat16 hi, mid, lo;
        split3(src[r * lds + c], hi, mid, lo);
        __nv_bfloat16* d = dst + r * 6 * C + c;
        // segment order = ascending product magnitude: mid*mid, hi*lo, lo*hi, hi*mid, mid*hi, hi*hi
        if (is_b) { d[0] = mid; d[C] = lo; d[2 * C] = hi; d[3 * C] = mid; d[4 * C] = hi;  d[5 * C] = hi; }
        else      { d[0] = mid; d[C] = hi; d[2 * C] = lo; d[3 * C] = hi;  d[4 * C] = mid; d[5 * C] = hi; }
    }
}

// src [R, C] fp32 row-major -> dst [C, 6R] bf16 (split of the TRANSPOSE), via a 32x32 smem tile.
__global__ void split3_transpose_kernel(const float* __restrict__ src, long long R, long long C,
                                        __nv_bfloat16* __restrict__ dst, int is_b) {
    __shared__ float tile[32][33];
    const long long r0 = static_cast<long long>(blockIdx.y) * 32, c0 = static_cast<long long>(blockIdx.x) * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long r = r0 + j, c = c0 + threadIdx.x;
        tile[j][threadIdx.x] = (r < R && c < C) ? src[r * C + c] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const long long c = c0 + j, r = r0 + threadIdx.x;      // output row = c, output col = r
        if (c < C && r < R) {
            __nv_bfloat16 hi, mid, lo;
            split3(tile[threadIdx.x][j], hi, mid, lo);
            __nv_bfloat16* d = dst + c * 6 * R + r;
            if (is_b) { d[0] = mid; d[R] = lo; d[2 * R] = hi; d[3 * R] = mid; d[4 * R] = hi;  d[5 * R] = hi; }
            else      { d[0] = mid; d[R] = hi; d[2 * R] = lo; d[3 * R] = hi;  d[4 * R] = mid; d[5 * R] = hi; }
        }
    }
}

// Finishes the fused cross entropy of gemm_tc_nt_kernel's CE mode: folds the per-(row, column
// block) partials (max, sum-exp) into lse / per-row loss — one warp per row — and the last CTA
// (atomic ticket) adds the per-row losses in a fixed order into the scaled total.
__global__ void __launch_bounds__(256)
ce_fold_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum, const float* __restrict__ tgt_logit,
               long long B, long long P, int ncol, const long long* __restrict__ target, long long target_stride,
               float loss_scale, float* lse_out, float* loss_rows, float* loss_out, unsigned int* ticket) {
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    const long long i = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (i < B) {
        float m = -FLT_MAX;
        for (int c = lane; c < ncol; c += 32) m = fmaxf(m, part_max[i * ncol + c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int c = lane; c < ncol; c += 32) s += part_sum[i * ncol + c] * expf(part_max[i * ncol + c] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const float lse = m + logf(s);
            const long long tc = target ? target[i] : i * target_stride;
            lse_out[i] = lse;
            loss_rows[i] = (tc >= 0 && tc < P) ? lse - tgt_logit[i] : __int_as_float(0x7fc00000);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ double s_red[256];
    double local = 0.0;
    for (long long r = threadIdx.x; r < B; r += blockDim.x) local += static_cast<double>(__ldcg(loss_rows + r));
    s_red[threadIdx.x] = local;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) { *loss_out = static_cast<float>(s_red[0] * static_cast<double>(loss_scale)); *ticket = 0u; }
}

// Row-wise cross entropy over stored logits (large-shape forward): one CTA per row, online
// max / sum-exp, per-row loss; the scaled total is accumulated in double by the last CTA.
__global__ void __launch_bounds__(256)
ce_rows_from_logits_kernel(const float* __restrict__ logits, long long B, long long P,
                           const long long* __restrict__ target, long long target_stride, float loss_scale,
                           float* lse_out, float* loss_rows, float* loss_out, unsigned int* ticket) {
    __shared__ float s_m[8], s_s[8];
    __shared__ bool s_last;
    const long long row = blockIdx.x;
    const float* l = logits + row * P;
    float m = -FLT_MAX, s = 0.f;
    for (long long j = threadIdx.x; j < P; j += blockDim.x) {
        const float v = l[j];
        const float mn = fmaxf(m, v);
        s = s * expf(m - mn) + expf(v - mn);
        m = mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, o), so = __shfl_xor_sync(0xffffffffu, s, o);
        const float mn = fmaxf(m, mo);
        s = s * expf(m - mn) + so * expf(mo - mn);
        m = mn;
    }
    if ((threadIdx.x & 31) == 0) { s_m[threadIdx.x >> 5] = m; s_s[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = s_m[0], ss = s_s[0];
        for (int w = 1; w < 8; ++w) {
            const float mn = fmaxf(mm, s_m[w]);
            ss = ss * expf(mm - mn) + s_s[w] * expf(s_m[w] - mn);
            mm = mn;
        }
        const float lse = mm + logf(ss);
        const long long tc = target ? target[row] : row * target_stride;
        lse_out[row] = lse;
        loss_rows[row] = (tc >= 0 && tc < P) ? lse - l[tc] : __int_as_float(0x7fc00000);
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ double s_red[256];
    double local = 0.0;
    for (long long i = threadIdx.x; i < B; i += blockDim.x) local += static_cast<double>(__ldcg(loss_rows + i));
    s_red[threadIdx.x] = local;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) { *loss_out = static_cast<float>(s_red[0] * static_cast<double>(loss_scale)); *ticket = 0u; }
}

}  // namespace drt
