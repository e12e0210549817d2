// mips_filter.cuh — K1: bf16 Q·Dᵀ on tcgen05 with a fused per-query threshold filter.
//
// Replaces the arithmetic of faiss.IndexFlatIP.search as called from
// DRT/evaluator/index.py:32 (sgemm blocks + heap k-selection on the CPU).
//
// Shape of the problem: scores[q, r] = <query q, corpus row r>, q in [0,nq), r in one corpus
// segment.  The kernel never writes the score matrix.  Queries sit on the MMA M dimension, so an
// accumulator row (= one TMEM lane) belongs to one query and ONE epilogue thread owns it: that
// thread keeps the query's current admission threshold in a register, streams the 256 scores of
// its lane out of TMEM with tcgen05.ld, and only when a score beats the threshold stages
// (score,row) in shared memory; staged candidates are appended to the query's list in global
// memory with one atomicAdd per flush (rare: O(k log N) candidates per query per search).
//
// Roles inside a CTA (320 threads):
//   warp 0    TMA producer: streams 128x64 query k-blocks and 256x64 (128x64 per CTA in pair
//             mode) corpus k-blocks into a 128B-swizzled smem ring.
//   warp 1    MMA issuer: one thread issues tcgen05.mma (M=128*kCtas, N=256, K=16) four times per
//             k-block into one of two 256-column TMEM accumulator stages; tcgen05.commit frees
//             smem slots and publishes finished accumulators.
//   warps 2-9 epilogue: two warps per TMEM lane quarter (each takes 128 of the tile's 256
//             columns); overlaps with the MMA of the next tile through the second accumulator
//             stage.
// kCtas = 2 runs the same protocol on a CTA pair (cta_group::2): each CTA owns 128 queries and
// loads half of every corpus tile, halving the corpus-operand smem/L2 traffic per FLOP.
//
// Work decomposition: a UNIT is one query tile x `unit_tiles` consecutive corpus tiles; units
// are numbered query-tile-fastest and dealt round-robin to the persistent CTAs.  The CTAs
// resident at any moment therefore work on the same few groups of corpus tiles (each tile leaves
// HBM once and is served from L2 to the other query tiles) while the bf16 query matrix stays L2
// resident, and an epilogue thread keeps one query (threshold in a register, staged candidates
// in smem) for a whole unit.
#pragma once
#include "ptx.cuh"

namespace drt {

constexpr int kTileM = 128;     // queries per CTA
constexpr int kTileN = 256;     // corpus rows per tile (per CTA pair in pair mode)
constexpr int kBlockK = 64;     // bf16 elements per k-block = one 128-byte swizzle span
constexpr int kEpilogueWarps = 8;
constexpr int kEpilogueThreads = kEpilogueWarps * 32;
constexpr int kFilterThreads = 64 + kEpilogueThreads;   // TMA warp + MMA warp + epilogue
constexpr int kStageSlots = 16; // staged candidates per epilogue thread before a flush

struct FilterParams {
    int num_m_tiles;        // ceil(nq / (128 * kCtas))
    int n_tile_begin;       // first 256-row tile of this launch, relative to the segment
    int n_tile_count;
    int unit_tiles;         // corpus tiles per work unit
    int num_k_blocks;       // dim / 64
    int nq;
    int f16;                // the 16-bit planes (queries and corpus) are IEEE fp16, not bf16
    uint32_t rows_valid;    // rows of this segment that may be admitted by this launch
    uint32_t row_base;      // store row id of the segment's first row
    uint32_t cap;           // candidate slots per query
    const float* thr;       // [nq] admission threshold on the UPPER-BOUND score (strict >)
    const float4* qbound;   // [nq] per-query error-bound coefficients (A, B, C, -), see below
    const float4* tile_bound; // this segment's per-256-row-tile maxima of the row bounds (r, Dx, Dt, 1/min|d|)
    const float4* row_bound;  // this segment's per-row bounds (read only for tiles that mix very different norms)
    uint32_t* cnt;          // [nq] candidates appended so far (may exceed cap: overflow)
    uint64_t* cand;         // [nq][cap] packed (ordered UPPER-BOUND score << 32 | ~row)
    int* err;               // host-mapped watchdog flag
};

// Monotone map float -> uint32 (ascending), so packed keys sort like (score, then lower row id
// first when sorted descending).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t b = __float_as_uint(f);
    return b ^ ((b & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t b = o ^ ((o & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu);
    return __uint_as_float(b);
}
__device__ __forceinline__ uint64_t pack_key(float score, uint32_t row) {
    return (static_cast<uint64_t>(float_to_ordered(score)) << 32) | (0xFFFFFFFFu - row);
}

// ---- exactness certificate: a rigorous bound on |first-pass score - fp32 rescoring score| -----
// Row j stores rb = (r, Dx, Dt): r = |d - image(d)|_2 over all dims, Dx = |d[0:split)|_2,
// Dt = |d[split:)|_2 (split = dim unless the caller declares an exactly-representable tail);
// query q carries qb = (A, B, C) = (|q~|(1+c_acc), e_x + c, e_t + c) with q~ = the 16-bit image of q
// the tensor core sees: bf16 by default, IEEE fp16 on both sides with DRT_B200_FIRST_PASS=f16 (three
// more significand bits: the rounding terms shrink 8x, but the MMA draws more power — see
// drt_b200.cu).  fp16's narrow range costs nothing in rigour: r is the EXACT residual norm of whatever image was stored (row
// values beyond +-65504 saturate and simply get a large r), and a query whose magnitudes fall
// outside [2^-6, 2^14] is scaled by a power of two sigma (kept in qb.w) that scales that
// query's whole first-pass domain: s~, ub, thr,
// e_x / e_t = |q - q~|_2 over the head / tail dims and c = c_acc |q~| + c_32 |q| covering the
// tensor core's fp32 accumulation (worst case: operands aligned to the largest exponent and
// TRUNCATED, 18 ulp of the running magnitude per K=16 step) and the rounding of the fp32
// rescoring dot.  By Cauchy-Schwarz, term by term,
//     s_fp32(q,j)  <=  ub(q,j) := s~(q,j) + A r + B Dx + C Dt            (all roundings upward)
// The first pass keeps, per query, the k' rows with the largest ub; every other row has
// ub <= thr (the k'-th largest ub), so once the exact k-th score exceeds thr no other row can
// belong to the top-k: a proof, not an estimate.  K1 works with the 256-row tile maxima
// (ub_tile >= ub_row: still a bound, a superset is admitted) and stores ub_tile as the
// candidate's score; a tile that mixes very different row norms re-tests its hits against the
// rows' own entries before admitting them (filter_chunk<true>), and in segments with such tiles
// select_kernel tightens the stored bound to the row's own entry.
constexpr float kBoundHuge = 1e30f;     // stands in for non-finite norms (keeps 0 * x finite)
__device__ __forceinline__ float bound_term(const float4& qb, const float4& rb) {
    return __fmaf_ru(qb.x, rb.x, __fmaf_ru(qb.y, rb.y, __fmul_ru(qb.z, rb.z)));
}

template <int kCtas>
struct FilterCfg {
    static constexpr int kStages = (kCtas == 1) ? 4 : 6;
    static constexpr uint32_t kABytes = kTileM * kBlockK * 2;             // 16 KB
    static constexpr uint32_t kBRows = kTileN / kCtas;                    // rows this CTA loads
    static constexpr uint32_t kBBytes = kBRows * kBlockK * 2;             // 32 KB | 16 KB
    static constexpr uint32_t kStageBytes = kABytes + kBBytes;
    static constexpr uint32_t kBarBytes = 256;                            // (2*kStages+4) mbarriers + tmem holder
    static constexpr uint32_t kStagingBytes = kStageSlots * kEpilogueThreads * 8;   // 32 KB candidate staging
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kBarBytes + kStagingBytes + 1024;
    static_assert((2 * kStages + 4) * 8 + 8 <= kBarBytes, "barrier block too small");
};

// wait for all outstanding tcgen05.ld; the register operands tie later uses to this point
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]),
                   "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]),
                   "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]),
                   "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// Append this thread's `n` staged candidates (smem slots my_stage + i*kSlotStride) to query q's list.
constexpr uint32_t kSlotStride = kEpilogueThreads * 8;
__device__ __noinline__ void flush_staging(uint32_t my_stage, uint32_t n, const FilterParams& p, int q) {
    const uint32_t base = atomicAdd(p.cnt + q, n);
    uint64_t* dst = p.cand + static_cast<size_t>(q) * p.cap;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t key;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(key) : "r"(my_stage + i * kSlotStride));
        if (base + i < p.cap) dst[base + i] = key;
    }
}

// Filter 32 consecutive scores of one query (registers v) against thr.  `row_id0` is the store
// row id of v[0]; only the first `nvalid` columns exist in the corpus.
// A hit is staged with its tile-level upper bound s~ + et (rounded up) as the key's score.
// `heavy` (warp-uniform): the tile mixes row norms more than 1.5x apart — typically one outlier row
// whose bound makes et huge, so that EVERY row of the tile passes the tile-level test (200 outliers
// per 2^20 rows would admit 51k rows per query and overflow the candidate buffers).  Such tiles
// re-test each hit against the row's own bound (`rb`, a warp-uniform 16-byte load per row).
template <bool kHeavy>
__device__ __forceinline__ void filter_chunk(const uint32_t (&v)[32], float thr, float et, uint32_t row_id0,
                                             int nvalid, uint32_t my_stage, uint32_t& scnt,
                                             const FilterParams& p, int q, float thr_pub,
                                             const float4& qb, const float4* rb) {
    // maxima of the four 8-column sub-blocks: the warp only walks a sub-block in which some
    // lane has a hit, so the admission cost stays proportional to the (rare) hits
    float mb[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float m = __uint_as_float(v[8 * b]);
#pragma unroll
        for (int j = 1; j < 8; ++j) m = fmaxf(m, __uint_as_float(v[8 * b + j]));
        mb[b] = m;
    }
    const bool any_hit = fmaxf(fmaxf(mb[0], mb[1]), fmaxf(mb[2], mb[3])) > thr;
    if (__any_sync(0xffffffffu, any_hit)) {  // warp-uniform; rare once the threshold warmed up
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            // warp-uniform entry, then branch-free (predicated) staging: every lane's hits of
            // this sub-block are handled by the same instruction stream, so dense admissions
            // (the threshold warm-up chunks, large k) do not serialise lane by lane
            if (__any_sync(0xffffffffu, mb[b] > thr)) {
#pragma unroll
                for (int j = 8 * b; j < 8 * b + 8; ++j) {
                    const float f = __uint_as_float(v[j]);
                    bool hit = (f > thr) && (j < nvalid);
                    if constexpr (kHeavy) hit = hit && (__fadd_ru(f, bound_term(qb, __ldg(rb + j))) > thr_pub);
                    const uint64_t key = pack_key(__fadd_ru(f, et), row_id0 + j);
                    if (hit) asm volatile("st.shared.u64 [%0], %1;" ::"r"(my_stage + scnt * kSlotStride), "l"(key) : "memory");
                    scnt += hit ? 1u : 0u;
                }
            }
            // A sub-block adds at most 8 entries, so keeping every lane at <= 8 staged entries
            // here bounds the staging at 16.  When one lane runs full the WHOLE warp flushes:
            // the lanes' atomics (different queries, different addresses) are in flight together
            // instead of each lane stalling the warp for its own round trip.
            if (__any_sync(0xffffffffu, scnt > kStageSlots - 8)) {
                if (scnt) flush_staging(my_stage, scnt, p, q);
                scnt = 0;
            }
        }
    }
    __syncwarp();   // the caller continues with warp-aligned tcgen05 instructions
}

template <int kCtas>
__global__ void __launch_bounds__(kFilterThreads, 1)
mips_filter_kernel(const __grid_constant__ CUtensorMap tmap_q,
                   const __grid_constant__ CUtensorMap tmap_d, const FilterParams p) {
    using Cfg = FilterCfg<kCtas>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rank = (kCtas == 2) ? ptx::cluster_ctarank() : 0u;

    // ---- shared memory carve-up (shared::cta addresses; tiles 1024-byte aligned) ----
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + kStages * Cfg::kStageBytes;
    const uint32_t staging = bars + Cfg::kBarBytes;
    auto smem_a = [&](int s) { return base + s * Cfg::kStageBytes; };
    auto smem_b = [&](int s) { return base + s * Cfg::kStageBytes + Cfg::kABytes; };
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
    const uint32_t tmem_holder = bars + 8u * (2 * kStages + 4);

    if constexpr (kCtas == 2) ptx::cluster_sync_all();

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_q);
        ptx::prefetch_tmap(&tmap_d);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(full_bar(s), kCtas);   // pair mode: leader + peer producer arrive
            ptx::mbar_init(empty_bar(s), 1);      // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(tfull_bar(a), 1);           // one tcgen05.commit
            ptx::mbar_init(tempty_bar(a), kEpilogueWarps * kCtas);  // one arrive per epilogue warp
        }
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
    const int n_groups = (p.n_tile_count + p.unit_tiles - 1) / p.unit_tiles;
    const int num_units = p.num_m_tiles * n_groups;
    const int nkb = p.num_k_blocks;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters) {
                const int m_tile = u % p.num_m_tiles;
                const int nt0 = (u / p.num_m_tiles) * p.unit_tiles;
                const int nt1 = min(nt0 + p.unit_tiles, p.n_tile_count);
                const int q_row = (m_tile * kCtas + static_cast<int>(rank)) * kTileM;
                for (int nt = nt0; nt < nt1; ++nt) {
                    const int d_row = (p.n_tile_begin + nt) * kTileN + static_cast<int>(rank) * Cfg::kBRows;
                    for (int kb = 0; kb < nkb; ++kb) {
                        ptx::mbar_wait(empty_bar(stage), phase ^ 1u, p.err, 101);
                        if constexpr (kCtas == 1) {
                            ptx::mbar_arrive_expect_tx(full_bar(stage), Cfg::kStageBytes);
                            ptx::tma_load_2d(smem_a(stage), &tmap_q, full_bar(stage), kb * kBlockK,
                                             q_row, ptx::kEvictLast);
                            ptx::tma_load_2d(smem_b(stage), &tmap_d, full_bar(stage), kb * kBlockK,
                                             d_row, ptx::kEvictNormal);
                        } else {
                            // both CTAs signal the LEADER's full barrier
                            const uint32_t lead_full = ptx::mapa(full_bar(stage), 0);
                            ptx::tma_load_2d_pair(smem_a(stage), &tmap_q, lead_full, kb * kBlockK,
                                                  q_row, ptx::kEvictLast);
                            ptx::tma_load_2d_pair(smem_b(stage), &tmap_d, lead_full, kb * kBlockK,
                                                  d_row, ptx::kEvictNormal);
                            if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
                            else           ptx::mbar_arrive_cluster(lead_full);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ==================================
        if (rank == 0 && ptx::elect_one()) {
            const uint32_t idesc = p.f16 ? ptx::make_idesc_f16(kTileM * kCtas, kTileN) : ptx::make_idesc_bf16(kTileM * kCtas, kTileN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int u = cluster_id; u < num_units; u += num_clusters) {
                const int nt0 = (u / p.num_m_tiles) * p.unit_tiles;
                const int nt1 = min(nt0 + p.unit_tiles, p.n_tile_count);
                for (int nt = nt0; nt < nt1; ++nt, ++it) {
                    const int acc = it & 1;
                    const uint32_t acc_phase = (it >> 1) & 1u;
                    ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u, p.err, 102);
                    ptx::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * kTileN;
                    for (int kb = 0; kb < nkb; ++kb) {
                        ptx::mbar_wait(full_bar(stage), phase, p.err, 103);
                        ptx::tc_fence_after();
                        const uint64_t a_desc = ptx::make_kmajor_sw128_desc(smem_a(stage));
                        const uint64_t b_desc = ptx::make_kmajor_sw128_desc(smem_b(stage));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            // +32 bytes (= 16 bf16) along K inside the swizzle span: +2 in the
                            // 16-byte-granular start-address field
                            ptx::umma_bf16<kCtas>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc,
                                                  (kb | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit<kCtas>(empty_bar(stage));      // smem slot free when read
                        if (kb == nkb - 1) ptx::umma_commit<kCtas>(tfull_bar(acc));  // accumulator done
                        if (++stage == kStages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else {
        // ================================ epilogue ====================================
        const uint32_t quarter = warp & 3u;   // TMEM lanes [32*quarter, +32) belong to this warp
        const uint32_t half = (warp - 2u) >> 2;   // which 128 columns of the tile this warp filters
        constexpr int kColsPerWarp = kTileN / (kEpilogueWarps / 4);
        const uint32_t lead_tempty0 = (kCtas == 2) ? ptx::mapa(tempty_bar(0), 0) : tempty_bar(0);
        const uint32_t my_stage = staging + ((warp - 2u) * 32u + lane) * 8u;
        uint32_t scnt = 0;
        int it = 0;
        for (int u = cluster_id; u < num_units; u += num_clusters) {
            const int m_tile = u % p.num_m_tiles;
            const int nt0 = (u / p.num_m_tiles) * p.unit_tiles;
            const int nt1 = min(nt0 + p.unit_tiles, p.n_tile_count);
            const int q = (m_tile * kCtas + static_cast<int>(rank)) * kTileM + quarter * 32 + lane;
            const float thr = (q < p.nq) ? p.thr[q] : __int_as_float(0x7f800000);
            const int qc = (q < p.nq) ? q : 0;
            const float4 qb = __ldg(p.qbound + qc);
            for (int nt = nt0; nt < nt1; ++nt, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1u;
                const uint32_t row0 = static_cast<uint32_t>(p.n_tile_begin + nt) * kTileN;
                const int valid_cols = static_cast<int>(min(p.rows_valid - row0, static_cast<uint32_t>(kTileN)));
                // the tile's error bound (warp-uniform 16-byte load, in flight while the MMA finishes):
                // a row is admitted when its UPPER-BOUND score s~ + E can exceed the threshold
                const float4 tb = __ldg(p.tile_bound + p.n_tile_begin + nt);
                ptx::mbar_wait(tfull_bar(acc), acc_phase, p.err, 104);
                ptx::tc_fence_after();
                const float et = bound_term(qb, tb);
                const float thr_eff = __fsub_rd(thr, et);
                const bool heavy = !(tb.y * tb.w <= 1.5f);        // max|d| / min|d| of the tile (NaN / inf: heavy)

                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kTileN + half * kColsPerWarp;
                const uint32_t col0 = half * kColsPerWarp;
                const float4* rb_tile = p.row_bound + row0 + col0;
                uint32_t va[32], vb[32];
                ptx::tmem_ld_32x32(taddr, va);
                const uint32_t id0 = p.row_base + row0 + col0;
                const int nv0 = valid_cols - static_cast<int>(col0);
                if (!heavy) {
#pragma unroll 1
                    for (int c = 0; c < kColsPerWarp / 32; c += 2) {
                        tmem_ld_wait_regs(va);
                        ptx::tmem_ld_32x32(taddr + (c + 1) * 32, vb);
                        filter_chunk<false>(va, thr_eff, et, id0 + c * 32, nv0 - c * 32, my_stage, scnt, p, qc, thr, qb, rb_tile);
                        tmem_ld_wait_regs(vb);
                        if (c + 2 < kColsPerWarp / 32) ptx::tmem_ld_32x32(taddr + (c + 2) * 32, va);
                        filter_chunk<false>(vb, thr_eff, et, id0 + (c + 1) * 32, nv0 - (c + 1) * 32, my_stage, scnt, p, qc, thr, qb, rb_tile);
                    }
                } else {
#pragma unroll 1
                    for (int c = 0; c < kColsPerWarp / 32; ++c) {
                        tmem_ld_wait_regs(va);
                        filter_chunk<true>(va, thr_eff, et, id0 + c * 32, nv0 - c * 32, my_stage, scnt, p, qc, thr, qb, rb_tile + c * 32);
                        if (c + 1 < kColsPerWarp / 32) ptx::tmem_ld_32x32(taddr + (c + 1) * 32, va);
                    }
                }
                // accumulator stage drained: hand it back to the MMA issuer (leader CTA's barrier)
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kCtas == 1) {
                        ptx::mbar_arrive(tempty_bar(acc));
                    } else {
                        if (rank == 0) ptx::mbar_arrive(tempty_bar(acc));
                        else           ptx::mbar_arrive_cluster(lead_tempty0 + 8u * acc);
                    }
                }
            }
            if (scnt) {   // the next unit belongs to another query: publish what is staged
                flush_staging(my_stage, scnt, p, qc);
                scnt = 0;
            }
            __syncwarp();
        }
    }

    // ---- teardown ----
    __syncwarp();
    ptx::tc_fence_before();
    if constexpr (kCtas == 2) ptx::cluster_sync_all(); else __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<kCtas>(tmem_base, 512);
}

}  // namespace drt
