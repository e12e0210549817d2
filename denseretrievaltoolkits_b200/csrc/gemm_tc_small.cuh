// gemm_tc_small.cuh — the in-batch loss at the reference's OWN shapes (BASELINE cfg3: 128 queries
// x 1024 passages x 768 dims, DRT/model/biencoder.py:107-116, DRT/trainer/losses.py:11-17):
// fp32-accurate contractions on tcgen05 for problems that are far too small to fill the GPU with
// 128x256 tiles, and the cross entropy folded into the accumulator read-out.
//
// Same arithmetic as gemm_tc.cuh — every fp32 operand split exactly into three bf16 pieces, the
// six significant partial products contracted as one bf16 GEMM over K' = 6K with fp32
// accumulation in TMEM — but laid out for latency instead of throughput:
//   * 128 x 64 output tiles and split-K, so that ~148 CTAs work on a 128 x 1024 x 4608 problem
//     (16 tiles x 9 K-chunks) instead of 4; one tile per CTA, no persistence;
//   * the whole K-chunk of a CTA is in flight at once (8 x 24 KB TMA stages);
//   * split-K partials go to an L2-resident workspace, the last CTA of a tile (atomic ticket) sums
//     them in a fixed order (deterministic) and runs the epilogue:
//       kStore — write the fp32 tile (backward: dx = dL·y, dy = dLᵀ·x, both problems in ONE launch);
//       kCe    — per-row max / sum-exp over the tile's 64 columns, the target logit, optionally the
//                logits themselves (only when the caller wants DROutput.scores or a backward
//                follows); the last tile of the launch folds the per-tile partials into
//                lse / per-row loss / the scaled total.  The score matrix never has to exist.
// Roles: warp 0 TMA producer, warp 1 tcgen05.mma issuer (one elected thread), warps 2-5 epilogue
// (one thread per accumulator row, 64 columns each, tcgen05.ld 32x32b.x32 twice).
#pragma once
#include <cuda_bf16.h>
#include <cfloat>
#include "gemm_tc.cuh"

namespace drt {

constexpr int kSmallTileN = 64;
constexpr int kSmallStages = 8;
constexpr int kSmallThreads = 192;
constexpr uint32_t kSmallABytes = kTileM * kBlockK * 2;            // 16 KB
constexpr uint32_t kSmallBBytes = kSmallTileN * kBlockK * 2;       // 8 KB
constexpr uint32_t kSmallStageBytes = kSmallABytes + kSmallBBytes;
constexpr uint32_t kSmallSmemBytes = kSmallStages * kSmallStageBytes + 256 + 1024;

struct SmallProblem {
    int m_tiles, n_tiles;      // 128-row / 64-column tiles of the output
    int ksplit;                // K-chunks per tile (CTAs per tile)
    int num_k_blocks;          // K' / 64
    long long M, N;
    float* C;                  // kStore: output [M, N] row-major; kCe: logits or NULL
    float* partials;           // [tiles][ksplit][128][64] fp32 (ksplit > 1)
    unsigned int* tile_ticket; // [tiles] zero before the launch, re-armed by the kernel
};

struct SmallCe {               // kCe only (problem 0)
    const long long* target;   // [M] or NULL -> row * target_stride
    long long target_stride;
    float loss_scale;
    float* part_max;           // [M][n_tiles]
    float* part_sum;           // [M][n_tiles]
    float* tgt_logit;          // [M]
    unsigned int* ticket;      // launch-wide: counts finished tiles
    float* lse_out;            // [M]
    float* loss_rows;          // [M]
    float* loss_out;           // [1]
};

struct SmallParams {
    SmallProblem prob[2];
    int ctas0;                 // CTAs [0, ctas0) work on problem 0, the rest on problem 1
    int* err;
    unsigned long long* dbg;   // dev tool (DRT_B200_CE_TRACE): [cta][8] globaltimer stamps of the phase boundaries
    SmallCe ce;
};

enum { kStore = 0, kCe = 1 };

__device__ __forceinline__ void trace_stamp(const SmallParams& p, int slot) {
    if (p.dbg) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.dbg[static_cast<size_t>(blockIdx.x) * 8 + slot] = t;
    }
}

// Everything after a CTA has its K-chunk's accumulator rows in registers (v = the 64 columns of row
// `et` of the tile): split-K publication / ticket / deterministic sum, then the store or the
// cross-entropy epilogue and the launch-wide fold.  `tile_scratch`: >= 128 x 65 floats of shared
// memory that no asynchronous operation touches any more; `s_last`: a shared int.
template <int kMode>
__device__ __forceinline__ void small_tile_finish(float (&v)[kSmallTileN], const SmallParams& p, const SmallProblem& pb,
                                                  int tile, int sp, int m_tile, int n_tile, int et, long long row,
                                                  float* tile_scratch, int* s_last) {
    bool finisher = true;
    if (pb.ksplit > 1) {
        float4* mine = reinterpret_cast<float4*>(pb.partials + ((static_cast<size_t>(tile) * pb.ksplit + sp) * kTileM + et) * kSmallTileN);
#pragma unroll
        for (int j = 0; j < kSmallTileN / 4; ++j) mine[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) *s_last = (atomicAdd(pb.tile_ticket + tile, 1u) == static_cast<unsigned int>(pb.ksplit) - 1u) ? 1 : 0;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        finisher = *s_last != 0;
        if (et == 0) trace_stamp(p, 3);                                 // partial published, ticket taken
        if (finisher) {
            __threadfence();
            // Sum the K-chunks in a fixed order (deterministic whichever CTA arrives last).  The
            // partial tiles are read as flat arrays — thread t takes 16-byte slots t, t + 128, ...
            // so a warp reads 512 contiguous bytes per instruction (row-wise reads, one 256-byte
            // row per thread, cost 1.3 us per K-chunk: request-rate bound) — and the sums go
            // through shared memory (the TMA stages are free: every MMA has retired) back to
            // one-row-per-thread for the epilogue.
            constexpr int kSlots = kTileM * kSmallTileN / 4 / 128;       // 16
            float4 acc[kSlots];
#pragma unroll
            for (int k = 0; k < kSlots; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4* pbase = reinterpret_cast<const float4*>(pb.partials + static_cast<size_t>(tile) * pb.ksplit * kTileM * kSmallTileN) + et;
            constexpr size_t pstride = static_cast<size_t>(kTileM) * kSmallTileN / 4;
            int s2 = 0;
            for (; s2 + 1 < pb.ksplit; s2 += 2) {
                float4 ta[kSlots], tb[kSlots];
#pragma unroll
                for (int k = 0; k < kSlots; ++k) { ta[k] = __ldcg(pbase + s2 * pstride + k * 128); tb[k] = __ldcg(pbase + (s2 + 1) * pstride + k * 128); }
#pragma unroll
                for (int k = 0; k < kSlots; ++k) {       // order: chunk s2, then s2 + 1
                    acc[k].x += ta[k].x; acc[k].y += ta[k].y; acc[k].z += ta[k].z; acc[k].w += ta[k].w;
                    acc[k].x += tb[k].x; acc[k].y += tb[k].y; acc[k].z += tb[k].z; acc[k].w += tb[k].w;
                }
            }
            if (s2 < pb.ksplit) {
#pragma unroll
                for (int k = 0; k < kSlots; ++k) {
                    const float4 t = __ldcg(pbase + s2 * pstride + k * 128);
                    acc[k].x += t.x; acc[k].y += t.y; acc[k].z += t.z; acc[k].w += t.w;
                }
            }
            float* tile_s = tile_scratch;   // [128][65]
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                const int slot = k * 128 + et, r = slot >> 4, c = (slot & 15) * 4;
                float* d = tile_s + r * (kSmallTileN + 1) + c;
                d[0] = acc[k].x; d[1] = acc[k].y; d[2] = acc[k].z; d[3] = acc[k].w;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int j = 0; j < kSmallTileN; ++j) v[j] = tile_s[et * (kSmallTileN + 1) + j];
            if (et == 0) pb.tile_ticket[tile] = 0u;       // re-arm for the next launch on this stream
            if (et == 0) trace_stamp(p, 4);                             // K-chunks summed
        }
    }
    if (finisher) {
        const long long col0 = static_cast<long long>(n_tile) * kSmallTileN;
        const bool row_ok = row < pb.M;
        const bool vec_ok = (pb.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(pb.C) & 15u) == 0) && col0 + kSmallTileN <= pb.N;
        auto store_tile = [&]() {
            if (pb.C && row_ok) {
                float* crow = pb.C + row * pb.N + col0;
                if (vec_ok) {
#pragma unroll
                    for (int j = 0; j < kSmallTileN; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < kSmallTileN; ++j) if (col0 + j < pb.N) crow[j] = v[j];
                }
            }
        };
        if constexpr (kMode == kStore) store_tile();
        if constexpr (kMode == kCe) {
            const SmallCe& ce = p.ce;
            if (row_ok) {
                const long long tcol = ce.target ? ce.target[row] : row * ce.target_stride;
                float mx = -FLT_MAX;
#pragma unroll
                for (int j = 0; j < kSmallTileN; ++j) if (col0 + j < pb.N) mx = fmaxf(mx, v[j]);
                float se = 0.f;
#pragma unroll
                for (int j = 0; j < kSmallTileN; ++j) if (col0 + j < pb.N) se += expf(v[j] - mx);
                ce.part_max[row * pb.n_tiles + n_tile] = mx;
                ce.part_sum[row * pb.n_tiles + n_tile] = se;
                if (tcol >= col0 && tcol < col0 + kSmallTileN && tcol < pb.N) {
                    float t = 0.f;
#pragma unroll
                    for (int j = 0; j < kSmallTileN; ++j) if (col0 + j == tcol) t = v[j];
                    ce.tgt_logit[row] = t;
                }
            }
            // ---- the last tile of the launch folds the per-tile partials ----
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) *s_last = (atomicAdd(ce.ticket, 1u) == static_cast<unsigned int>(pb.m_tiles * pb.n_tiles) - 1u) ? 1 : 0;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) trace_stamp(p, 5);                             // tile epilogue done, launch ticket taken
            store_tile();        // the logits (if wanted) after the ticket: the fold does not read them
            if (*s_last) {
                __threadfence();
                __shared__ double s_red[128];
                double local_sum = 0.0;
                const int nt = pb.n_tiles;
                for (long long i = et; i < pb.M; i += 128) {
                    // all loads of a step are independent and in flight together: a
                    // one-load-per-iteration loop is a chain of L2 latencies
                    float m = -FLT_MAX, s = 0.f;
                    for (int c0 = 0; c0 < nt; c0 += 16) {
                        float pm[16], ps[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const bool in = c0 + u < nt;
                            pm[u] = in ? __ldcg(ce.part_max + i * nt + c0 + u) : -FLT_MAX;
                            ps[u] = in ? __ldcg(ce.part_sum + i * nt + c0 + u) : 0.f;
                        }
                        float cm = m;
#pragma unroll
                        for (int u = 0; u < 16; ++u) cm = fmaxf(cm, pm[u]);
                        s *= expf(m - cm);
#pragma unroll
                        for (int u = 0; u < 16; ++u) s += ps[u] * expf(pm[u] - cm);
                        m = cm;
                    }
                    const float lse = m + logf(s);
                    const long long tc = ce.target ? ce.target[i] : i * ce.target_stride;
                    // an out-of-range target poisons the loss instead of reading a stale logit
                    const float li = (tc >= 0 && tc < pb.N) ? lse - __ldcg(ce.tgt_logit + i) : __int_as_float(0x7fc00000);
                    ce.lse_out[i] = lse;
                    ce.loss_rows[i] = li;
                    local_sum += static_cast<double>(li);
                }
                s_red[et] = local_sum;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                for (int o = 64; o > 0; o >>= 1) {
                    if (et < o) s_red[et] += s_red[et + o];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                if (et == 0) {
                    *ce.loss_out = static_cast<float>(s_red[0] * static_cast<double>(ce.loss_scale));
                    *ce.ticket = 0u;
                    trace_stamp(p, 6);                                  // fold done
                }
            }
        }
    }
}

template <int kMode>
__global__ void __launch_bounds__(kSmallThreads, 1)
gemm_tc_small_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_b0,
                     const __grid_constant__ CUtensorMap tmap_a1, const __grid_constant__ CUtensorMap tmap_b1,
                     const SmallParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ int s_last;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    if (threadIdx.x == 64) trace_stamp(p, 0);                               // CTA started
    const int which = (static_cast<int>(blockIdx.x) < p.ctas0) ? 0 : 1;
    const SmallProblem& pb = p.prob[which];
    const CUtensorMap* tmap_a = which == 0 ? &tmap_a0 : &tmap_a1;
    const CUtensorMap* tmap_b = which == 0 ? &tmap_b0 : &tmap_b1;
    const int local = static_cast<int>(blockIdx.x) - (which == 0 ? 0 : p.ctas0);
    const int tile = local / pb.ksplit, sp = local - tile * pb.ksplit;
    const int m_tile = tile % pb.m_tiles, n_tile = tile / pb.m_tiles;
    const int kb0 = static_cast<int>(static_cast<long long>(pb.num_k_blocks) * sp / pb.ksplit);
    const int nkb = static_cast<int>(static_cast<long long>(pb.num_k_blocks) * (sp + 1) / pb.ksplit) - kb0;

    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + kSmallStages * kSmallStageBytes;
    auto smem_a = [&](int s) { return base + s * kSmallStageBytes; };
    auto smem_b = [&](int s) { return base + s * kSmallStageBytes + kSmallABytes; };
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kSmallStages + s); };
    const uint32_t tfull_bar = bars + 8u * (2 * kSmallStages);
    const uint32_t tmem_holder = bars + 8u * (2 * kSmallStages + 1);

    if (warp == 0 && lane == 0) { ptx::prefetch_tmap(tmap_a); ptx::prefetch_tmap(tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kSmallStages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 1); }
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<1>(tmem_holder, kSmallTileN);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_holder));

    if (warp == 0) {
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.err, 301);
                ptx::mbar_arrive_expect_tx(full_bar(stage), kSmallStageBytes);
                ptx::tma_load_2d(smem_a(stage), tmap_a, full_bar(stage), (kb0 + kb) * kBlockK, m_tile * kTileM, ptx::kEvictNormal);
                ptx::tma_load_2d(smem_b(stage), tmap_b, full_bar(stage), (kb0 + kb) * kBlockK, n_tile * kSmallTileN, ptx::kEvictNormal);
                if (++stage == kSmallStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (ptx::elect_one()) {
            const uint32_t idesc = ptx::make_idesc_bf16(kTileM, kSmallTileN);
            int stage = 0;
            uint32_t phase = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase, p.err, 303);
                ptx::tc_fence_after();
                const uint64_t a_desc = ptx::make_kmajor_sw128_desc(smem_a(stage));
                const uint64_t b_desc = ptx::make_kmajor_sw128_desc(smem_b(stage));
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                    ptx::umma_bf16<1>(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                ptx::umma_commit<1>(empty_bar(stage));
                if (kb == nkb - 1) ptx::umma_commit<1>(tfull_bar);
                if (++stage == kSmallStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ================================ epilogue (128 threads) ==============================
        const uint32_t quarter = warp & 3u;                  // TMEM lanes [32 * quarter, +32)
        const int et = static_cast<int>(quarter * 32u + lane);   // accumulator row of this thread
        const long long row = static_cast<long long>(m_tile) * kTileM + et;
        float v[kSmallTileN];
        if (et == 0) trace_stamp(p, 1);                                     // setup done, waiting for the accumulator
        ptx::mbar_wait(tfull_bar, 0u, p.err, 304);
        ptx::tc_fence_after();
        if (et == 0) trace_stamp(p, 2);                                     // MMAs of this K-chunk retired
        {
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16);
            ptx::tmem_ld_32x32(taddr, r0);
            ptx::tmem_ld_32x32(taddr + 32, r1);
            tmem_ld_wait_regs(r0);
            tmem_ld_wait_regs(r1);
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
        }
        small_tile_finish<kMode>(v, p, pb, tile, sp, m_tile, n_tile, et, row,
                                 reinterpret_cast<float*>(smem_raw + (base - ptx::smem_u32(smem_raw))), &s_last);
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, kSmallTileN);
    if (threadIdx.x == 64) trace_stamp(p, 7);                               // CTA done
}

// pieces of 8 consecutive values as three 16-byte vectors (hi, mid, lo)
__device__ __forceinline__ void split3x8(const float (&x)[8], uint4& hi, uint4& mid, uint4& lo) {
    __nv_bfloat16 h[8], m[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split3(x[i], h[i], m[i], l[i]);
    hi = *reinterpret_cast<const uint4*>(h);
    mid = *reinterpret_cast<const uint4*>(m);
    lo = *reinterpret_cast<const uint4*>(l);
}

// ---- forward with the operand split fused in --------------------------------------------------
// gemm_tc_fused_kernel: same tiles, split-K and epilogue as gemm_tc_small_kernel<kCe>, but the CTA
// reads the fp32 operands themselves (TMA, no swizzle) and forms the three bf16 pieces of every
// value in shared memory, in the 128-byte-swizzled K-major layout the MMA descriptors expect — no
// preparation launch, no 6x-expanded operand copy in HBM/L2.  Per 64-wide k-block: the producer
// lands x[128x64] + y[64x64] fp32 (48 KB) in a 2-stage ring; the four epilogue warps convert them
// (hi/mid/lo tiles: 3 x 16 KB for x, 3 x 8 KB for y); fence.proxy.async; the MMA thread issues the
// six partial products as 24 K=16 MMAs — the five small ones into one TMEM accumulator, hi*hi into
// a second one, so the tensor core's truncating fp32 accumulation never adds small terms to a
// full-scale running sum (gemm_tc.cuh orders the K' segments for the same reason); the two are
// added when the accumulators are read.  Needs dim % 64 == 0 and 16-byte aligned fp32 operands.
constexpr int kFusedStages = 2;
constexpr uint32_t kFusedXBytes = kTileM * kBlockK * 4;             // 32 KB fp32 staging
constexpr uint32_t kFusedYBytes = kSmallTileN * kBlockK * 4;        // 16 KB
constexpr uint32_t kFusedStageBytes = kFusedXBytes + kFusedYBytes;
constexpr uint32_t kFusedOpsBytes = 3 * kSmallABytes + 3 * kSmallBBytes;   // 72 KB of bf16 piece tiles
constexpr uint32_t kFusedSmemBytes = kFusedStages * kFusedStageBytes + kFusedOpsBytes + 256 + 1024;

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 8 consecutive k of one row: fp32 staging -> one 16-byte chunk in each of the hi / mid / lo tiles
__device__ __forceinline__ void convert_chunk(const uint8_t* stage_rows, uint8_t* hi_t, uint8_t* mid_t, uint8_t* lo_t, int r, int c) {
    const float4 a = *reinterpret_cast<const float4*>(stage_rows + r * 256 + c * 32);
    const float4 b = *reinterpret_cast<const float4*>(stage_rows + r * 256 + c * 32 + 16);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 hi, mid, lo;
    split3x8(x, hi, mid, lo);
    const int off = r * 128 + ((c ^ (r & 7)) << 4);               // 128-byte swizzle: chunk ^= row % 8
    *reinterpret_cast<uint4*>(hi_t + off) = hi;
    *reinterpret_cast<uint4*>(mid_t + off) = mid;
    *reinterpret_cast<uint4*>(lo_t + off) = lo;
}

__global__ void __launch_bounds__(kSmallThreads, 1)
gemm_tc_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y, const SmallParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ int s_last;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (threadIdx.x == 64) trace_stamp(p, 0);
    const SmallProblem& pb = p.prob[0];
    const int tile = static_cast<int>(blockIdx.x) / pb.ksplit, sp = static_cast<int>(blockIdx.x) - tile * pb.ksplit;
    const int m_tile = tile % pb.m_tiles, n_tile = tile / pb.m_tiles;
    const int kb0 = static_cast<int>(static_cast<long long>(pb.num_k_blocks) * sp / pb.ksplit);      // SOURCE k-blocks (dim / 64)
    const int nkb = static_cast<int>(static_cast<long long>(pb.num_k_blocks) * (sp + 1) / pb.ksplit) - kb0;

    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - ptx::smem_u32(smem_raw));
    const uint32_t ops = base + kFusedStages * kFusedStageBytes;       // A hi | A mid | A lo | B hi | B mid | B lo
    const uint32_t bars = ops + kFusedOpsBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };              // fp32 stage landed (TMA tx)
    auto empty_bar = [&](int s) { return bars + 8u * (kFusedStages + s); };   // stage converted (128 arrivals)
    const uint32_t ready_bar = bars + 8u * (2 * kFusedStages);         // piece tiles written (128 arrivals)
    const uint32_t free_bar = bars + 8u * (2 * kFusedStages + 1);      // MMAs that read them retired (commit)
    const uint32_t tfull_bar = bars + 8u * (2 * kFusedStages + 2);
    const uint32_t tmem_holder = bars + 8u * (2 * kFusedStages + 3);

    if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmap_x); ptx::prefetch_tmap(&tmap_y); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kFusedStages; ++s) { ptx::mbar_init(full_bar(s), 1); ptx::mbar_init(empty_bar(s), 128); }
        ptx::mbar_init(ready_bar, 128);
        ptx::mbar_init(free_bar, 1);
        ptx::mbar_init(tfull_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<1>(tmem_holder, 2 * kSmallTileN);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_holder));

    if (warp == 0) {
        if (ptx::elect_one()) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int stage = kb % kFusedStages;
                const uint32_t phase = static_cast<uint32_t>(kb / kFusedStages) & 1u;
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.err, 311);
                ptx::mbar_arrive_expect_tx(full_bar(stage), kFusedStageBytes);
                ptx::tma_load_2d(base + stage * kFusedStageBytes, &tmap_x, full_bar(stage), (kb0 + kb) * kBlockK, m_tile * kTileM, ptx::kEvictNormal);
                ptx::tma_load_2d(base + stage * kFusedStageBytes + kFusedXBytes, &tmap_y, full_bar(stage), (kb0 + kb) * kBlockK,
                                 n_tile * kSmallTileN, ptx::kEvictNormal);
            }
        }
    } else if (warp == 1) {
        if (ptx::elect_one()) {
            const uint32_t idesc = ptx::make_idesc_bf16(kTileM, kSmallTileN);
            const uint32_t a_hi = ops, a_mid = ops + kSmallABytes, a_lo = ops + 2 * kSmallABytes;
            const uint32_t b_hi = ops + 3 * kSmallABytes, b_mid = b_hi + kSmallBBytes, b_lo = b_hi + 2 * kSmallBBytes;
            // ascending product magnitude inside the small accumulator; hi*hi alone in the second one
            const uint32_t pa[6] = {a_mid, a_hi, a_lo, a_hi, a_mid, a_hi};
            const uint32_t pbt[6] = {b_mid, b_lo, b_hi, b_mid, b_hi, b_hi};
            for (int kb = 0; kb < nkb; ++kb) {
                ptx::mbar_wait(ready_bar, static_cast<uint32_t>(kb) & 1u, p.err, 313);
                ptx::tc_fence_after();
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const uint64_t a_desc = ptx::make_kmajor_sw128_desc(pa[c]);
                    const uint64_t b_desc = ptx::make_kmajor_sw128_desc(pbt[c]);
                    const uint32_t d_tmem = tmem_base + (c == 5 ? kSmallTileN : 0);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        const uint32_t accumulate = (c == 5) ? ((kb | k) != 0 ? 1u : 0u) : ((kb | c | k) != 0 ? 1u : 0u);
                        ptx::umma_bf16<1>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, accumulate);
                    }
                }
                ptx::umma_commit<1>(free_bar);
                if (kb == nkb - 1) ptx::umma_commit<1>(tfull_bar);
            }
        }
    } else {
        const uint32_t quarter = warp & 3u;
        const int et = static_cast<int>(quarter * 32u + lane);
        const int ct = static_cast<int>((warp - 2u) * 32u + lane);           // converter thread id 0..127
        const long long row = static_cast<long long>(m_tile) * kTileM + et;
        uint8_t* g_ops = gbase + kFusedStages * kFusedStageBytes;
        uint8_t *a_hi = g_ops, *a_mid = g_ops + kSmallABytes, *a_lo = g_ops + 2 * kSmallABytes;
        uint8_t *b_hi = g_ops + 3 * kSmallABytes, *b_mid = b_hi + kSmallBBytes, *b_lo = b_hi + 2 * kSmallBBytes;
        if (et == 0) trace_stamp(p, 1);
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % kFusedStages;
            const uint32_t phase = static_cast<uint32_t>(kb / kFusedStages) & 1u;
            ptx::mbar_wait(full_bar(stage), phase, p.err, 314);
            if (kb > 0) ptx::mbar_wait(free_bar, static_cast<uint32_t>(kb - 1) & 1u, p.err, 315);   // previous k-block's MMAs read the piece tiles
            const uint8_t* sx = gbase + stage * kFusedStageBytes;
            const uint8_t* sy = sx + kFusedXBytes;
#pragma unroll
            for (int i = 0; i < (kTileM * 8) / 128; ++i) {           // 1024 chunks of x
                const int id = ct + 128 * i;
                convert_chunk(sx, a_hi, a_mid, a_lo, id >> 3, id & 7);
            }
#pragma unroll
            for (int i = 0; i < (kSmallTileN * 8) / 128; ++i) {      // 512 chunks of y
                const int id = ct + 128 * i;
                convert_chunk(sy, b_hi, b_mid, b_lo, id >> 3, id & 7);
            }
            fence_proxy_async_smem();                                 // generic-proxy writes -> visible to the MMA's async proxy
            ptx::mbar_arrive(ready_bar);
            ptx::mbar_arrive(empty_bar(stage));                       // the fp32 stage may be refilled
        }
        float v[kSmallTileN];
        if (nkb > 0) {
            ptx::mbar_wait(tfull_bar, 0u, p.err, 316);
            ptx::tc_fence_after();
            if (et == 0) trace_stamp(p, 2);
            uint32_t r0[32], r1[32];
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                ptx::tmem_ld_32x32(taddr + h * 32, r0);                      // small terms
                ptx::tmem_ld_32x32(taddr + kSmallTileN + h * 32, r1);        // hi * hi
                tmem_ld_wait_regs(r0);
                tmem_ld_wait_regs(r1);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[h * 32 + j] = __uint_as_float(r1[j]) + __uint_as_float(r0[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kSmallTileN; ++j) v[j] = 0.f;
        }
        small_tile_finish<kCe>(v, p, pb, tile, sp, m_tile, n_tile, et, row, reinterpret_cast<float*>(gbase), &s_last);
    }
    __syncwarp();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, 2 * kSmallTileN);
    if (threadIdx.x == 64) trace_stamp(p, 7);
}

// ---- operand preparation: every split / transpose / dlogits job of one pass in ONE launch ------
// Job: dst[out_r][6 * out_c layout] = split3(element(r, c)) of a logical [R, C] fp32 source,
// optionally transposed (dst rows = source columns).  `kind` 0: element = src[r * lds + c];
// kind 1: element = dlogits(r, c) = g_r * (exp(logits[r, c] - lse[r]) - [c == target_r]),
// g_r = grad_scale * grad_rows[r * grad_stride] — the backward's first step, never materialised
// in fp32 (drt_inbatch_ce_bwd's `work` buffer stays untouched on this path).
struct SplitJob {
    const float* src;          // source matrix (kind 0) or logits (kind 1), row-major [R, C]
    long long R, C, lds;
    __nv_bfloat16* dst;        // [R, 6C] (transpose = 0) or [C, 6R] (transpose = 1)
    int is_b;                  // segment order of the B operand (gemm_tc.cuh)
    int transpose;
    int kind;
    int vec;                   // row job with C % 8 == 0 and 16-byte aligned rows: 32 x 64 blocks, 16-byte accesses
    int tiles_x, tiles_y;      // blocks covering [R, C] (32 x 32, or 32 x 64 when vec)
    int tile_begin;            // first blockIdx.x of this job
};
struct SplitJobs {
    SplitJob job[4];
    int n_jobs;
    // kind-1 inputs
    const float* lse;
    const long long* target;
    long long target_stride;
    const float* grad_rows;
    int grad_stride;
    float grad_scale;
};

__device__ __forceinline__ void store_split(__nv_bfloat16* d, long long seg, float x, int is_b) {
    __nv_bfloat16 hi, mid, lo;
    split3(x, hi, mid, lo);
    // segment order = ascending product magnitude: mid*mid, hi*lo, lo*hi, hi*mid, mid*hi, hi*hi
    if (is_b) { d[0] = mid; d[seg] = lo; d[2 * seg] = hi; d[3 * seg] = mid; d[4 * seg] = hi;  d[5 * seg] = hi; }
    else      { d[0] = mid; d[seg] = hi; d[2 * seg] = lo; d[3 * seg] = hi;  d[4 * seg] = mid; d[5 * seg] = hi; }
}

__global__ void __launch_bounds__(256)
split3_jobs_kernel(const SplitJobs js) {
    __shared__ float tile[32][33];
    int j = 0;
#pragma unroll
    for (int t = 1; t < 4; ++t) if (t < js.n_jobs && static_cast<int>(blockIdx.x) >= js.job[t].tile_begin) j = t;
    const SplitJob& jb = js.job[j];
    const int tl = static_cast<int>(blockIdx.x) - jb.tile_begin;
    if (!jb.transpose && jb.vec) {
        // row job, C % 8 == 0 and 16-byte aligned rows: a block covers 32 rows x 64 columns, a
        // thread 8 consecutive columns -> two 16-byte loads, six 16-byte stores
        const long long r = static_cast<long long>(tl / jb.tiles_x) * 32 + (threadIdx.x >> 3);
        const long long c = static_cast<long long>(tl % jb.tiles_x) * 64 + (threadIdx.x & 7) * 8;
        if (r >= jb.R || c >= jb.C) return;
        const float4 a = *reinterpret_cast<const float4*>(jb.src + r * jb.lds + c);
        const float4 b = *reinterpret_cast<const float4*>(jb.src + r * jb.lds + c + 4);
        float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        if (jb.kind == 1) {
            const long long tcol = js.target ? js.target[r] : r * js.target_stride;
            const float g = js.grad_scale * js.grad_rows[r * js.grad_stride], l = js.lse[r];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = g * (expf(x[i] - l) - (c + i == tcol ? 1.f : 0.f));
        }
        uint4 hi, mid, lo;
        split3x8(x, hi, mid, lo);
        uint4* d = reinterpret_cast<uint4*>(jb.dst + r * 6 * jb.C + c);
        const long long seg = jb.C / 8;                      // one segment, in 16-byte units
        // segment order = ascending product magnitude: mid*mid, hi*lo, lo*hi, hi*mid, mid*hi, hi*hi
        if (jb.is_b) { d[0] = mid; d[seg] = lo; d[2 * seg] = hi; d[3 * seg] = mid; d[4 * seg] = hi;  d[5 * seg] = hi; }
        else         { d[0] = mid; d[seg] = hi; d[2 * seg] = lo; d[3 * seg] = hi;  d[4 * seg] = mid; d[5 * seg] = hi; }
        return;
    }
    const long long r0 = static_cast<long long>(tl / jb.tiles_x) * 32, c0 = static_cast<long long>(tl % jb.tiles_x) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const long long r = r0 + i, c = c0 + tx;
        float val = 0.f;
        if (r < jb.R && c < jb.C) {
            val = jb.src[r * jb.lds + c];
            if (jb.kind == 1) {
                const long long tcol = js.target ? js.target[r] : r * js.target_stride;
                const float g = js.grad_scale * js.grad_rows[r * js.grad_stride];
                val = g * (expf(val - js.lse[r]) - (c == tcol ? 1.f : 0.f));
            }
        }
        tile[i][tx] = val;
    }
    __syncthreads();
    if (!jb.transpose) {
        for (int i = ty; i < 32; i += 8) {
            const long long r = r0 + i, c = c0 + tx;
            if (r < jb.R && c < jb.C) store_split(jb.dst + r * 6 * jb.C + c, jb.C, tile[i][tx], jb.is_b);
        }
    } else {
        for (int i = ty; i < 32; i += 8) {
            const long long c = c0 + i, r = r0 + tx;              // output row = c, output column = r
            if (c < jb.C && r < jb.R) store_split(jb.dst + c * 6 * jb.R + r, jb.R, tile[tx][i], jb.is_b);
        }
    }
}

}  // namespace drt
