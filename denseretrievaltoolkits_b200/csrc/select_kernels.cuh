// select_kernels.cuh — the integer/index side of the search path: query preparation (16-bit image
// + error-bound coefficients), candidate-list trimming by upper bound and threshold publication
// (between corpus chunks), exact fp32 rescoring + final ordering + the exactness certificate, the
// ingest kernel (16-bit plane + per-row / per-tile error-bound tables), the cross-shard merge
// (incl. the peer-memory exchange form) and the mining filter.
//
// These are HBM/L2-bound byte and index kernels: coalesced loads, shared-memory staging, no
// tensor cores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cfloat>
#include "inbatch_ce.cuh"
#include "mips_filter.cuh"

namespace drt {

__device__ __forceinline__ int next_pow2_dev(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// In-place descending bitonic sort of P (power of two) 64-bit keys in shared memory.
__device__ __forceinline__ void bitonic_sort_desc_u64(uint64_t* keys, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// Query preparation, one warp per query: fp32 -> the 16-bit image the tensor core reads, the
// query's error-bound coefficients qb = (A, B, C, sigma) (mips_filter.cuh), and the initial search
// state (threshold -FLT_MAX like faiss' heap, empty candidate list).
//   f16 = 1: image = fp16(sigma q) (round to nearest even), sigma a power of two that brings
//     max|q_i| to ~1 when it lies outside [2^-6, 2^14] (fp16 overflows at 65504 and loses
//     precision below 6e-5), else 1.  The query's first-pass scores, upper bounds and threshold
//     all live in the sigma-scaled domain; K2 multiplies exact scores by sigma to compare.
//   f16 = 0: image = bf16(q), sigma = 1.
//   exact_mode = 1 (fp32 SIMT first pass of the last-resort refinement): A = B = C = 0, rows are
//   selected by their fp32 scores as they are (exact ties keep the canonical id order).
__device__ __forceinline__ float image16(float x, int f16) {
    return f16 ? __half2float(__float2half_rn(x)) : __bfloat162float(__float2bfloat16_rn(x));
}
// row values: fp16 saturates at +-65504 instead of overflowing to inf (the residual norm then
// carries the difference, so the bound stays valid and the row is simply always a candidate)
__device__ __forceinline__ float sat16(float x, int f16) {
    return (f16 && x == x) ? fminf(fmaxf(x, -65504.f), 65504.f) : x;
}

__global__ void __launch_bounds__(256)
prep_queries_kernel(const float* __restrict__ q, int nq, int dim, int split, uint2* __restrict__ q16,
                    float4* __restrict__ qbound, float* __restrict__ thr, uint32_t* __restrict__ cnt,
                    float c_acc, float c_32, int exact_mode, int f16) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float up = 1.f + 0x1p-10f;      // covers the rounding of the fp32 sums / sqrt below
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nq; i += warps) {
        const float4* src = reinterpret_cast<const float4*>(q + static_cast<size_t>(i) * dim);
        uint2* dst = q16 ? q16 + static_cast<size_t>(i) * (dim >> 2) : nullptr;
        float sigma = 1.f;
        if (f16 && !exact_mode) {
            float mx = 0.f;
            for (int j = lane; j < (dim >> 2); j += 32) {
                const float4 v = src[j];
                mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (mx > 0.f && mx < 3.0e38f && (mx >= 16384.f || mx < 0x1p-6f)) {
                int e;
                frexpf(mx, &e);                       // mx = m 2^e, m in [0.5, 1)
                sigma = ldexpf(1.f, -e);              // sigma mx in [0.5, 1)
            }
        }
        float s_t = 0.f, s_ex = 0.f, s_et = 0.f, s_n = 0.f;
        for (int j = lane; j < (dim >> 2); j += 32) {
            const float4 v = src[j];
            const float x[4] = {v.x * sigma, v.y * sigma, v.z * sigma, v.w * sigma};     // exact: power of two
            float xt[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                xt[c] = image16(x[c], f16);
                const float e = x[c] - xt[c];                  // exact in fp32
                s_t = fmaf(xt[c], xt[c], s_t);
                s_n = fmaf(x[c], x[c], s_n);
                if (4 * j + c < split) s_ex = fmaf(e, e, s_ex); else s_et = fmaf(e, e, s_et);
            }
            if (dst) {
                uint2 o;
                if (f16) {
                    const __half2 lo = __floats2half2_rn(x[0], x[1]);
                    const __half2 hi = __floats2half2_rn(x[2], x[3]);
                    o.x = *reinterpret_cast<const uint32_t*>(&lo);
                    o.y = *reinterpret_cast<const uint32_t*>(&hi);
                } else {
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(x[0], x[1]);
                    const __nv_bfloat162 hi = __floats2bfloat162_rn(x[2], x[3]);
                    o.x = *reinterpret_cast<const uint32_t*>(&lo);
                    o.y = *reinterpret_cast<const uint32_t*>(&hi);
                }
                dst[j] = o;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s_t += __shfl_xor_sync(0xffffffffu, s_t, o);
            s_n += __shfl_xor_sync(0xffffffffu, s_n, o);
            s_ex += __shfl_xor_sync(0xffffffffu, s_ex, o);
            s_et += __shfl_xor_sync(0xffffffffu, s_et, o);
        }
        if (lane == 0) {
            const float nt = sqrtf(s_t) * up, nn = sqrtf(s_n) * up;
            float A, B, C;
            if (exact_mode) {
                A = B = C = 0.f;
            } else {
                const float c = (c_acc * nt + c_32 * nn) * up;
                A = nt * (1.f + c_acc) * up;
                B = (sqrtf(s_ex) * up + c) * up;
                C = (sqrtf(s_et) * up + c) * up;
            }
            if (!(A < kBoundHuge)) A = kBoundHuge;             // also catches NaN
            if (!(B < kBoundHuge)) B = kBoundHuge;
            if (!(C < kBoundHuge)) C = kBoundHuge;
            qbound[i] = make_float4(A, B, C, sigma);
            thr[i] = -FLT_MAX;                                  // faiss: heap starts at -FLT_MAX
            cnt[i] = 0u;
        }
    }
}

// K-select: one CTA per query.  If the query gathered more than `keep` candidates, keep the
// `keep` with the largest UPPER-BOUND score (ties: row asc) and publish the keep-th ub as the new
// admission threshold.  Every row that was never admitted, or is dropped here, has ub <= that
// threshold, hence an exact score <= it: the invariant K2's exactness certificate rests on.
// Chunks are visited in ascending row order and admission is strict (>).
//
// K1 stored every candidate with its TILE-level bound.  In segments flagged `heavy` (some tile
// mixes row norms more than 1.5x apart, e.g. one huge row among ordinary ones) the bound is
// tightened here to the row's own entry — ub_row = RU(RU(ub_tile - E_tile) + E_row), still an
// upper bound — so one outlier cannot fill the list with its 255 tile neighbours.  That costs a
// 16-byte gather per candidate, which homogeneous corpora never pay.
//
// Selection is an MSD radix select on the packed 64-bit keys: 8-bit digits (256-bin smem
// histogram, one bin per thread, block-wide suffix scan to locate the bin holding the keep-th
// largest key), starting at the highest bit in which the keys differ at all (scores of one query
// share their sign / exponent bits: the leading digits would land in one bin).  O(n) work per
// query; keys are unique (the row id is part of the key), so exactly `keep` keys are >= the
// pivot.  Survivors are written back unordered: neither the filter kernel nor later selects need
// order, and K2 sorts its own output.
//
// Fast path: up to `smem_keys` candidates are staged in shared memory.  A longer list
// (thresholds frozen too early, adversarial row order) is selected straight from global memory
// and compacted through a scratch row; the host sizes chunks so that this is the exception.
__device__ __forceinline__ uint64_t tighten_key(uint64_t key, const float4& qb, const float4* const* seg_bound,
                                                const float4* const* seg_tile, const unsigned char* seg_heavy,
                                                uint32_t seg_rows) {
    const uint32_t lo = static_cast<uint32_t>(key);
    const uint32_t r = 0xFFFFFFFFu - lo;
    const uint32_t seg = r / seg_rows;
    if (!seg_heavy[seg]) return key;
    const uint32_t local = r - seg * seg_rows;
    const float et = bound_term(qb, __ldg(seg_tile[seg] + (local >> 8)));
    const float er = bound_term(qb, __ldg(seg_bound[seg] + local));
    const float ub = __fadd_ru(__fsub_ru(ordered_to_float(static_cast<uint32_t>(key >> 32)), et), er);
    return (static_cast<uint64_t>(float_to_ordered(ub)) << 32) | lo;
}

// kTighten = false: no segment of the store is heavy, the stored keys are final (8 B per staged
// key instead of 12, no per-candidate segment lookup).
template <bool kTighten>
__global__ void __launch_bounds__(256)
select_kernel(uint64_t* cand, uint32_t* cnt, float* thr, uint32_t cap, uint32_t keep,
              int* overflow, const float4* __restrict__ qbound, const float4* const* __restrict__ seg_bound,
              const float4* const* __restrict__ seg_tile, const unsigned char* __restrict__ seg_heavy,
              uint32_t seg_rows, uint32_t smem_keys, uint64_t* scratch, uint32_t max_keep) {
    // max_keep >= keep: the select may stop as soon as a digit boundary leaves between keep and
    // max_keep candidates (threshold = that boundary, a valid if slightly lower bound; usually
    // after ONE pass).  Between chunks the list only has to shrink and the threshold to rise; the
    // final select before K2 passes max_keep = keep and is exact.
    extern __shared__ uint64_t s_keys[];             // [smem_keys] tightened keys, then [smem_keys] stored scores
    uint32_t* s_orig = reinterpret_cast<uint32_t*>(s_keys + smem_keys);
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_bin, s_need, s_bucket, s_out, s_wsum[8];
    __shared__ uint64_t s_pivot, s_or[8], s_and[8];
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    uint32_t n = cnt[q];
    if (n > cap) {                       // appends were dropped: the host must redo the search
        if (tid == 0) *overflow = 1;
        n = cap;
    }
    if (n <= keep) return;
    uint64_t* row = cand + static_cast<size_t>(q) * cap;
    const float4 qb = qbound[q];
    const uint32_t old_thr_o = float_to_ordered(thr[q]);
    const bool in_smem = n <= smem_keys;
    // the lists keep the STORED (tile-level) keys, so tightening is applied afresh by every select
    auto tight = [&](uint64_t stored) -> uint64_t {
        return kTighten ? tighten_key(stored, qb, seg_bound, seg_tile, seg_heavy, seg_rows) : stored;
    };
    auto tkey = [&](uint32_t i) -> uint64_t { return in_smem ? s_keys[i] : tight(row[i]); };
    // ---- stage / tighten, and find the bits in which the keys differ ----
    uint64_t k_or = 0, k_and = ~0ull;
    for (uint32_t i = tid; i < n; i += blockDim.x) {
        const uint64_t stored = row[i];
        const uint64_t key = tight(stored);
        if (in_smem) { s_keys[i] = key; if (kTighten) s_orig[i] = static_cast<uint32_t>(stored >> 32); }
        k_or |= key; k_and &= key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        k_or |= __shfl_xor_sync(0xffffffffu, k_or, o);
        k_and &= __shfl_xor_sync(0xffffffffu, k_and, o);
    }
    if ((tid & 31) == 0) { s_or[tid >> 5] = k_or; s_and[tid >> 5] = k_and; }
    if (tid == 0) { s_need = keep; s_out = 0; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) { k_or |= s_or[w]; k_and &= s_and[w]; }
    const uint64_t differ = k_or ^ k_and;              // n > keep >= 1 distinct keys: never 0
    const int top = 63 - __clzll(static_cast<long long>(differ));
    int shift = max(0, top - 7);                        // first digit = bits [shift, shift + 8)
    uint64_t mask = (shift + 8 >= 64) ? 0ull : ~((1ull << (shift + 8)) - 1ull);
    uint64_t prefix = k_and & mask;                     // bits above the first digit are common to all keys
    bool found = false;
    while (!found) {
        s_hist[tid] = 0;                 // blockDim.x == 256
        __syncthreads();
        for (uint32_t i = tid; i < n; i += blockDim.x) {
            const uint64_t key = tkey(i);
            if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        {
            // thread t owns bin t; suffix-scan the counts to find the bin (from the top) at which
            // the running count reaches `need`
            const uint32_t need = s_need;
            const uint32_t h = s_hist[tid];
            uint32_t v = h;                          // inclusive suffix sum inside the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_down_sync(0xffffffffu, v, o);
                if ((tid & 31) + o < 32) v += t;
            }
            if ((tid & 31) == 0) s_wsum[tid >> 5] = v;
            __syncthreads();
            uint32_t above = v - h;                  // keys in higher bins of this warp ...
            for (int w = (tid >> 5) + 1; w < 8; ++w) above += s_wsum[w];   // ... and of higher warps
            if (above < need && need <= above + h) { s_bin = tid; s_need = need - above; s_bucket = h; }
        }
        __syncthreads();
        prefix |= static_cast<uint64_t>(s_bin) << shift;
        mask |= 0xFFull << shift;
        if (s_bucket == 1 || shift == 0) {   // one key carries this prefix (always true at shift 0): the pivot
            for (uint32_t i = tid; i < n; i += blockDim.x) {
                const uint64_t key = tkey(i);
                if ((key & mask) == prefix) s_pivot = key;
            }
            found = true;
        } else if (keep - s_need + s_bucket <= max_keep && static_cast<uint32_t>(prefix >> 32) >= old_thr_o) {
            // everything from this digit's lower edge upwards: keep .. max_keep candidates.  Only
            // if that edge does not fall below the threshold already published (the exact
            // keep-th never does): with a wide score range the first digits are coarse, and a
            // lowered threshold would flood the next chunk with admissions.
            if (tid == 0) s_pivot = prefix;
            found = true;
        }
        __syncthreads();
        // the next digit may overlap bits already fixed when shift < 8: harmless, the overlapping
        // bits are equal for every key that still matches the prefix
        shift = max(0, shift - 8);
    }
    const uint64_t pivot = s_pivot;
    if (in_smem) {
        for (uint32_t i = tid; i < n; i += blockDim.x) {
            const uint64_t key = s_keys[i];
            if (key >= pivot)
                row[atomicAdd(&s_out, 1u)] = kTighten ? (static_cast<uint64_t>(s_orig[i]) << 32) | (key & 0xFFFFFFFFull) : key;
        }
    } else {
        // compact through the scratch row (the list cannot be rewritten in place while it is read)
        uint64_t* tmp = scratch + static_cast<size_t>(q) * max_keep;
        for (uint32_t i = tid; i < n; i += blockDim.x) {
            const uint64_t stored = row[i];
            if (tight(stored) >= pivot) tmp[atomicAdd(&s_out, 1u)] = stored;
        }
        __syncthreads();
        const uint32_t kept = s_out;
        for (uint32_t i = tid; i < kept; i += blockDim.x) row[i] = tmp[i];
    }
    __syncthreads();
    if (tid == 0) {
        cnt[q] = s_out;                   // keep, or up to max_keep after an early stop
        thr[q] = ordered_to_float(static_cast<uint32_t>(pivot >> 32));
    }
}

// Per-segment `heavy` flags for select_kernel: one CTA per segment scans the segment's tile table
// (max residual, max norm, max tail norm, max 1/norm per 256-row tile).
__global__ void __launch_bounds__(256)
segment_heavy_kernel(const float4* const* __restrict__ seg_tile, int tiles_per_seg, const long long* __restrict__ seg_valid_rows,
                     unsigned char* __restrict__ seg_heavy) {
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    const int seg = blockIdx.x;
    const int tiles = static_cast<int>((seg_valid_rows[seg] + 255) >> 8);
    int any = 0;
    for (int t = threadIdx.x; t < tiles && t < tiles_per_seg; t += blockDim.x) {
        const float4 tb = seg_tile[seg][t];
        // tb.y = max |d|, tb.w = max 1/|d| over the tile: their product is max|d| / min|d|
        if (!(tb.y * tb.w <= 1.5f)) any = 1;           // also true for NaN / inf (zero or non-finite rows)
    }
    if (any) atomicOr(&s_any, 1);
    __syncthreads();
    if (threadIdx.x == 0) seg_heavy[seg] = static_cast<unsigned char>(s_any);
}

// K2: exact fp32 rescoring of the first-pass candidates + final ordering + output.
// One CTA per query; each warp computes whole dot products (coalesced float4 row reads).
// Replaces the fp32 exactness of faiss' sgemm for the rows that can still matter.
//
// The candidates are first ordered by their upper bound ub.  The best k of them (by ub) are
// rescored; the smallest exact score among those is a lower bound tau_lb of the final k-th
// score, so a remaining candidate is rescored only if its ub reaches tau_lb — one whose ub lies
// below cannot enter the top-k.  For 768-d Gaussian-like data that skips ~1/4 of the fp32 row
// gathers at k' = 2k (the kernel is a pure HBM gather: bytes are time).
__device__ __forceinline__ float rescore_dot(const float4* __restrict__ x4, const float4* q4, int dim4, int lane) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int j = lane; j < dim4; j += 32) {
        const float4 x = __ldg(x4 + j);
        const float4 y = q4[j];
        a0 = fmaf(x.x, y.x, a0); a1 = fmaf(x.y, y.y, a1);
        a2 = fmaf(x.z, y.z, a2); a3 = fmaf(x.w, y.w, a3);
    }
    float acc = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}

// kThreads: 256 (six CTAs per SM when there are thousands of queries) or 1024 for small batches
// (at most two queries per SM: the row gathers of one query are then spread over 32 warps instead
// of 8, which is what bounds the kernel there — 63 -> ~25 us at 128 queries).
template <int kThreads>
__global__ void __launch_bounds__(kThreads, kThreads == 256 ? 6 : 1)
rescore_kernel(const uint64_t* cand, const uint32_t* cnt, uint32_t cap, uint32_t keep,
               const float* q_f32, int dim, const float* const* seg_f32, uint32_t seg_rows,
               int k, long long id_offset, float* out_scores, long long* out_ids,
               int do_rescore, unsigned long long* flagged, unsigned char* qflag, int check,
               const float* __restrict__ thr, const float4* __restrict__ qbound) {
    extern __shared__ uint64_t s_keys[];            // [P] then dim floats
    __shared__ unsigned int s_taulb;                // min exact score of the best-k-by-ub (ordered encoding)
    __shared__ unsigned int s_done;                 // rows actually rescored (statistics)
    const int q = blockIdx.x;
    const int n = static_cast<int>(min(min(cnt[q], cap), keep));
    const int P = next_pow2_dev(max(n, 1));
    float* s_q = reinterpret_cast<float*>(s_keys + next_pow2_dev(static_cast<int>(keep)));
    for (int j = threadIdx.x; j < dim; j += blockDim.x) s_q[j] = q_f32[static_cast<size_t>(q) * dim + j];
    if (threadIdx.x == 0) { s_taulb = 0xFFFFFFFFu; s_done = 0u; }
    const float sigma = qbound[q].w;    // the first-pass domain (ub, thr) of this query is scaled by sigma (a power of two)

    const uint64_t* row = cand + static_cast<size_t>(q) * cap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_keys[i] = row[i];
    for (int i = n + threadIdx.x; i < P; i += blockDim.x) s_keys[i] = 0ull;
    bitonic_sort_desc_u64(s_keys, P);               // by the stored upper bound

    if (do_rescore) {
        const float4* q4 = reinterpret_cast<const float4*>(s_q);
        const int kk = min(k, n);
        for (int pass = 0; pass < 2; ++pass) {
            const int lo = pass == 0 ? 0 : kk, hi = pass == 0 ? kk : n;
            const unsigned int taulb = (pass == 1 && kk == k) ? s_taulb : 0u;   // no skipping without k exact scores
            for (int i = lo + warp; i < hi; i += nwarps) {
                const uint64_t key = s_keys[i];
                if (static_cast<uint32_t>(key >> 32) < taulb) {      // ub below k exact scores: cannot enter
                    if (lane == 0) s_keys[i] = 0ull;
                    continue;
                }
                const uint32_t r = 0xFFFFFFFFu - static_cast<uint32_t>(key);
                const uint32_t seg = r / seg_rows, local = r - seg * seg_rows;
                const float acc = rescore_dot(reinterpret_cast<const float4*>(seg_f32[seg] + static_cast<size_t>(local) * dim),
                                              q4, dim >> 2, lane);
                if (lane == 0) {
                    const bool ok = acc > -FLT_MAX;    // also false for NaN
                    s_keys[i] = ok ? pack_key(acc, r) : 0ull;
                    if (pass == 0) atomicMin(&s_taulb, ok ? float_to_ordered(acc * sigma) : 0u);
                    atomicAdd(&s_done, 1u);
                }
            }
            __syncthreads();
        }
        bitonic_sort_desc_u64(s_keys, P);
    }

    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = (i < P) ? s_keys[i] : 0ull;
        const size_t o = static_cast<size_t>(q) * k + i;
        if (key != 0ull) {
            out_scores[o] = ordered_to_float(static_cast<uint32_t>(key >> 32));
            out_ids[o] = id_offset + static_cast<long long>(0xFFFFFFFFu - static_cast<uint32_t>(key));
        } else {
            out_scores[o] = -FLT_MAX;
            out_ids[o] = -1;
        }
    }
    // Exactness certificate.  Every row that is not in this list has an upper-bound score
    // ub <= thr[q] (select_kernel's invariant; thr is still -FLT_MAX when nothing was ever
    // dropped), and ub bounds the value THIS kernel would compute for that row.  So if the exact
    // k-th score lies strictly above thr[q], no other row can belong to the top-k: proven exact.
    // Otherwise the query is flagged and the host searches it again (larger k', then fp32).
    if (threadIdx.x == 0) {
        bool flag = false;
        if (check && do_rescore) {
            const float bound = thr[q];
            if (bound > -FLT_MAX) {
                const bool have_k = n >= k && s_keys[k - 1] != 0ull;
                flag = !have_k || !(ordered_to_float(static_cast<uint32_t>(s_keys[k - 1] >> 32)) * sigma > bound);
            }
        }
        if (flag) atomicAdd(flagged, 1ull);
        if (do_rescore) atomicAdd(flagged + 1, static_cast<unsigned long long>(s_done));
        if (qflag) qflag[q] = flag ? 1 : 0;
    }
}

// Fallback plumbing: pull the flagged queries into a dense sub-batch / put their results back.
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                   float* __restrict__ dst, int dim, int n) {
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<size_t>(idx[r]) * dim);
        float4* d4 = reinterpret_cast<float4*>(dst + static_cast<size_t>(r) * dim);
        for (int j = threadIdx.x; j < (dim >> 2); j += blockDim.x) d4[j] = s4[j];
    }
}
__global__ void scatter_results_kernel(const float* __restrict__ ss, const long long* __restrict__ si,
                                       const unsigned char* __restrict__ sf, const int* __restrict__ idx,
                                       float* os, long long* oi, unsigned char* qf, int k, int n) {
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
        const size_t d = static_cast<size_t>(idx[r]) * k, s = static_cast<size_t>(r) * k;
        for (int j = threadIdx.x; j < k; j += blockDim.x) { os[d + j] = ss[s + j]; oi[d + j] = si[s + j]; }
        if (threadIdx.x == 0) qf[idx[r]] = sf[r];
    }
}

// Exact first pass (last-resort refinement): fp32 FFMA scores of a few queries against a chunk of
// the fp32 plane, fed into the same threshold / candidate-list machinery as K1.  Used only for
// queries whose exactness check still fails after the bf16 pass was repeated with larger k'
// (e.g. a corpus of near-duplicates whose score gaps are below bf16 resolution), so every
// returned result is either margin-checked or computed from exact scores.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS)
exact_filter_kernel(const float* __restrict__ q, long long nq, const float* __restrict__ rows, long long nrows,
                    int dim, uint32_t row_id0, const float* __restrict__ thr, uint32_t* cnt, uint64_t* cand,
                    uint32_t cap, int vec_ok, const float4* __restrict__ qbound, const float4* __restrict__ row_bound) {
    constexpr int TM = Cfg::TM, TN = Cfg::TN;
    const long long m0 = static_cast<long long>(blockIdx.y) * Cfg::BM;
    const long long n0 = static_cast<long long>(blockIdx.x) * Cfg::BN;
    float acc[TM][TN];
    gemm_tile<Cfg>(GemmOperand{q, dim, 1}, GemmOperand{rows, dim, 1}, nq, nrows, dim, m0, n0, vec_ok, vec_ok, acc);
    const int tx = threadIdx.x % Cfg::TX, ty = threadIdx.x / Cfg::TX;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long qi = m0 + Cfg::row_of(ty, i);
        if (qi >= nq) continue;
        const float t = thr[qi];
        const float4 qb = qbound[qi];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const long long col = n0 + Cfg::col_of(tx, j);
            // same rule as K1: admit when the upper bound (here only fp32 summation-order slack) beats thr
            if (col < nrows && acc[i][j] > __fsub_rd(t, bound_term(qb, __ldg(row_bound + col)))) {
                const uint32_t slot = atomicAdd(cnt + qi, 1u);
                if (slot < cap) cand[static_cast<size_t>(qi) * cap + slot] = pack_key(acc[i][j], row_id0 + static_cast<uint32_t>(col));
            }
        }
    }
}

// K6 ingest, one warp per row: fp32 row -> bf16 plane (round to nearest even) plus the row's entry
// of the error-bound table rb = (|d - bf16(d)|, |d[0:split)|, |d[split:)|, 0), each rounded up,
// and the component-wise maxima of the row's 256-row tile (atomicMax on the bit patterns of
// non-negative floats; the fourth component tracks max 1/|d|, i.e. the tile's smallest norm).  Non-finite norms are stored as kBoundHuge.  HBM-bound: 4 B in, 2 B out
// per element, 16 B per row.
__global__ void __launch_bounds__(256)
ingest_rows_kernel(const float* __restrict__ plane, long long n, int dim, int split, uint2* __restrict__ bf16,
                   float4* __restrict__ row_bound, unsigned int* __restrict__ tile_bound, long long row0, int f16) {
    const int lane = threadIdx.x & 31;
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const float up = 1.f + 0x1p-10f;
    for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5; i < n; i += warps) {
        const float4* src = reinterpret_cast<const float4*>(plane + static_cast<size_t>(i) * dim);
        uint2* dst = bf16 + static_cast<size_t>(i) * (dim >> 2);
        float s_r = 0.f, s_x = 0.f, s_t = 0.f;
        for (int j = lane; j < (dim >> 2); j += 32) {
            const float4 v = src[j];
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float e = x[c] - image16(sat16(x[c], f16), f16);                // exact in fp32 (inf when x is)
                s_r = fmaf(e, e, s_r);
                if (4 * j + c < split) s_x = fmaf(x[c], x[c], s_x); else s_t = fmaf(x[c], x[c], s_t);
            }
            uint2 o;
            if (f16) {
                const __half2 lo = __floats2half2_rn(sat16(v.x, 1), sat16(v.y, 1));
                const __half2 hi = __floats2half2_rn(sat16(v.z, 1), sat16(v.w, 1));
                o.x = *reinterpret_cast<const uint32_t*>(&lo);
                o.y = *reinterpret_cast<const uint32_t*>(&hi);
            } else {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
                const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
                o.x = *reinterpret_cast<const uint32_t*>(&lo);
                o.y = *reinterpret_cast<const uint32_t*>(&hi);
            }
            dst[j] = o;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s_r += __shfl_xor_sync(0xffffffffu, s_r, o);
            s_x += __shfl_xor_sync(0xffffffffu, s_x, o);
            s_t += __shfl_xor_sync(0xffffffffu, s_t, o);
        }
        if (lane == 0) {
            float r = sqrtf(s_r) * up, dx = sqrtf(s_x) * up, dt = sqrtf(s_t) * up;
            if (!(r < kBoundHuge)) r = kBoundHuge;
            if (!(dx < kBoundHuge)) dx = kBoundHuge;
            if (!(dt < kBoundHuge)) dt = kBoundHuge;
            row_bound[row0 + i] = make_float4(r, dx, dt, 0.f);
            unsigned int* tb = tile_bound + 4 * ((row0 + i) >> 8);
            atomicMax(tb + 0, __float_as_uint(r));
            atomicMax(tb + 1, __float_as_uint(dx));
            atomicMax(tb + 2, __float_as_uint(dt));
            const float nrm = sqrtf(s_x + s_t);
            atomicMax(tb + 3, __float_as_uint(nrm > 0.f ? 1.f / nrm : __int_as_float(0x7f800000)));   // 1 / min norm
        }
    }
}

// fp32 -> bf16 plane conversion only (16 B in / 8 B out per thread step)
__global__ void f32_to_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t n4) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4 v = in[i];
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&lo);
        o.y = *reinterpret_cast<const uint32_t*>(&hi);
        out[i] = o;
    }
}

// drt_search_async: "this result is not final" (candidate overflow, or the certificate flagged a
// query) as ONE device byte, so that the caller's merge step — not the host — can look at it.
__global__ void publish_status_kernel(const int* overflow, const unsigned long long* flagged, unsigned char* status) {
    if (threadIdx.x == 0) *status = (*overflow != 0 || *flagged != 0ull) ? 1 : 0;
}

__global__ void fill_outputs_kernel(float* scores, long long* ids, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        scores[i] = -FLT_MAX;
        ids[i] = -1;
    }
}

// ---------------------------------------------------------------------------------------------
// K3 cross-shard merge (merge_retrieval_results_by_score, DRT/model/utils.py:215-229).
struct MergeEnt { float s; int src; long long id; };

template <class Less>
__device__ __forceinline__ void bitonic_sort_ents(MergeEnt* e, int P, Less less) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool fwd = (lo & size) == 0;
                const MergeEnt a = e[lo], b = e[hi];
                if (less(b, a) == fwd && (less(b, a) || less(a, b))) { e[lo] = b; e[hi] = a; }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
merge_topk_kernel(int G, const float* scores, const long long* ids, long long nq, int k_in,
                  int k_out, float* out_scores, long long* out_ids) {
    extern __shared__ uint64_t s_raw[];
    MergeEnt* e = reinterpret_cast<MergeEnt*>(s_raw);
    const long long q = blockIdx.x;
    const int n = G * k_in;
    const int P = next_pow2_dev(n);
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        MergeEnt m;
        if (i < n) {
            const int g = i / k_in, j = i - g * k_in;
            const size_t o = (static_cast<size_t>(g) * nq + q) * k_in + j;
            m.s = scores[o]; m.id = ids[o]; m.src = i;
            if (m.id < 0 || !(m.s > -FLT_MAX)) { m.id = -1; m.s = -FLT_MAX; }
        } else { m.s = -FLT_MAX; m.id = -1; m.src = i; }
        e[i] = m;
    }
    // pass 1: by (id asc, src asc), padding last -> duplicates of an id become adjacent with the
    // first partition's copy in front (utils.py:224-226: first partition that mentions it wins)
    bitonic_sort_ents(e, P, [](const MergeEnt& a, const MergeEnt& b) {
        const unsigned long long ia = static_cast<unsigned long long>(a.id), ib = static_cast<unsigned long long>(b.id);
        return ia < ib || (ia == ib && a.src < b.src);   // id -1 -> 0xFFFF... sorts last
    });
    bool dup[32];  // P <= 8192, 256 threads -> <= 32 entries per thread
    int c = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x, ++c) dup[c] = (i > 0 && e[i].id >= 0 && e[i].id == e[i - 1].id);
    __syncthreads();
    c = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x, ++c) if (dup[c]) { e[i].id = -1; e[i].s = -FLT_MAX; }
    // pass 2: canonical result order (score desc, id asc), padding last
    bitonic_sort_ents(e, P, [](const MergeEnt& a, const MergeEnt& b) {
        if (a.s != b.s) return a.s > b.s;
        return static_cast<unsigned long long>(a.id) < static_cast<unsigned long long>(b.id);
    });
    for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
        const size_t o = static_cast<size_t>(q) * k_out + i;
        if (i < P && e[i].id >= 0) { out_scores[o] = e[i].s; out_ids[o] = e[i].id; }
        else { out_scores[o] = -FLT_MAX; out_ids[o] = -1; }
    }
}

// K3, fast path for the row-sharded store: every input list is already in canonical order
// (score desc, id asc) and the shards hold disjoint id ranges, so no id can repeat.  Each entry's
// output position is then its merge rank = (position in its own list) + sum over the other lists
// of (entries that precede it), found by binary search: O(G*k*G*log k) per query, no sort.
__device__ __forceinline__ bool merge_precedes(float sa, long long ia, float sb, long long ib) {
    return sa > sb || (sa == sb && static_cast<unsigned long long>(ia) < static_cast<unsigned long long>(ib));
}

__global__ void __launch_bounds__(256)
merge_sorted_kernel(int G, const float* scores, const long long* ids, long long nq, int k_in,
                    int k_out, float* out_scores, long long* out_ids) {
    extern __shared__ uint64_t s_raw[];
    const long long q = blockIdx.x;
    const int n = G * k_in;
    long long* s_id = reinterpret_cast<long long*>(s_raw);
    float* s_sc = reinterpret_cast<float*>(s_id + n);
    __shared__ int s_valid;
    if (threadIdx.x == 0) s_valid = 0;
    __syncthreads();
    int local_valid = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int g = i / k_in, j = i - g * k_in;
        const size_t o = (static_cast<size_t>(g) * nq + q) * k_in + j;
        float sc = scores[o];
        long long id = ids[o];
        if (id < 0 || !(sc > -FLT_MAX)) { id = -1; sc = -FLT_MAX; } else ++local_valid;
        s_id[i] = id; s_sc[i] = sc;
    }
    if (local_valid) atomicAdd(&s_valid, local_valid);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const long long id = s_id[i];
        if (id < 0) continue;                      // padding sits at list tails and is never placed
        const float sc = s_sc[i];
        const int g = i / k_in;
        int rank = i - g * k_in;
        for (int h = 0; h < G; ++h) {
            if (h == g) continue;
            int lo = 0, hi = k_in;                 // first position in list h that does not precede us
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const long long im = s_id[h * k_in + mid];
                if (im >= 0 && merge_precedes(s_sc[h * k_in + mid], im, sc, id)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k_out) {
            const size_t o = static_cast<size_t>(q) * k_out + rank;
            out_scores[o] = sc;
            out_ids[o] = id;
        }
    }
    const int valid = s_valid;
    for (int i = valid + threadIdx.x; i < k_out; i += blockDim.x) {
        const size_t o = static_cast<size_t>(q) * k_out + i;
        out_scores[o] = -FLT_MAX;
        out_ids[o] = -1;
    }
}

// K3 over peer memory: the candidate exchange and the merge as ONE kernel.  Every rank keeps its
// [Q, k_in] lists in a buffer that is mapped into all peers (NVLink / NVSwitch).  The CTA of query
// q reads the G lists of q straight from the G peers (P2P loads), merges them by rank as above,
// and stores the merged row — and the "a shard list was cut short" flag of the reduced-depth
// check — into EVERY peer's result buffer (P2P stores), so no all-gather precedes or follows.
// Rank r launches the queries of its own slice; the caller brackets the launch with two
// cross-rank barriers (lists complete before / results complete after).
struct PeerPtrs {
    const float* scores[16];
    const long long* ids[16];
    float* out_scores[16];
    long long* out_ids[16];
    unsigned char* truncated[16];
    const unsigned char* status[16];   // per-rank "local result not final" bytes (NULL: not used)
    unsigned char* redo;               // this rank's OR of all of them
};

__global__ void __launch_bounds__(256)
merge_sorted_peers_kernel(PeerPtrs p, int G, long long q_begin, int k_in, int k_out) {
    extern __shared__ uint64_t s_raw[];
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.redo) {
        // every rank reads the same G bytes (written before the first barrier): all ranks agree
        unsigned char any = 0;
        for (int g = 0; g < G; ++g) if (p.status[g]) any |= *p.status[g];
        *p.redo = any;
    }
    const long long q = q_begin + blockIdx.x;
    const int n = G * k_in;
    long long* s_id = reinterpret_cast<long long*>(s_raw);
    long long* s_oid = s_id + n;
    float* s_sc = reinterpret_cast<float*>(s_oid + k_out);
    float* s_osc = s_sc + n;
    __shared__ int s_valid, s_trunc;
    if (threadIdx.x == 0) { s_valid = 0; s_trunc = 0; }
    for (int i = threadIdx.x; i < k_out; i += blockDim.x) { s_osc[i] = -FLT_MAX; s_oid[i] = -1; }
    __syncthreads();
    int local_valid = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int g = i / k_in, j = i - g * k_in;
        const size_t o = static_cast<size_t>(q) * k_in + j;
        float sc = p.scores[g][o];
        long long id = p.ids[g][o];
        if (id < 0 || !(sc > -FLT_MAX)) { id = -1; sc = -FLT_MAX; } else ++local_valid;
        s_id[i] = id; s_sc[i] = sc;
    }
    if (local_valid) atomicAdd(&s_valid, local_valid);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const long long id = s_id[i];
        if (id < 0) continue;
        const float sc = s_sc[i];
        const int g = i / k_in;
        int rank = i - g * k_in;
        for (int h = 0; h < G; ++h) {
            if (h == g) continue;
            int lo = 0, hi = k_in;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const long long im = s_id[h * k_in + mid];
                if (im >= 0 && merge_precedes(s_sc[h * k_in + mid], im, sc, id)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k_out) { s_osc[rank] = sc; s_oid[rank] = id; }
    }
    __syncthreads();
    // a full list whose last score is not strictly below the merged k_out-th may hold more rows
    const float kth = s_osc[k_out - 1];
    if (threadIdx.x < G) {
        const float last = s_sc[threadIdx.x * k_in + k_in - 1];
        if (s_id[threadIdx.x * k_in + k_in - 1] >= 0 && last >= kth) s_trunc = 1;
    }
    __syncthreads();
    const unsigned char tr = static_cast<unsigned char>(s_trunc);
    for (int g = 0; g < G; ++g) {
        if (p.out_scores[g]) {          // NULL: that rank does not want this query's row (rank-local results)
            float* os = p.out_scores[g] + static_cast<size_t>(q) * k_out;
            long long* oi = p.out_ids[g] + static_cast<size_t>(q) * k_out;
            for (int i = threadIdx.x; i < k_out; i += blockDim.x) { os[i] = s_osc[i]; oi[i] = s_oid[i]; }
        }
        if (threadIdx.x == 0) p.truncated[g][q] = tr;
    }
}

__global__ void or_status_kernel(PeerPtrs p, int G) {
    if (threadIdx.x == 0 && p.redo) {
        unsigned char any = 0;
        for (int g = 0; g < G; ++g) if (p.status[g]) any |= *p.status[g];
        *p.redo = any;
    }
}

// Mining filter (process_sample, DRT/trainer/sampler.py:69-80): one warp per query, ballot
// compaction keeps rank order.
__global__ void filter_negatives_kernel(const long long* ids, long long nq, int k,
                                        const long long* pos_begin, const long long* pos_end,
                                        int num_negative, long long* out) {
    const long long q = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const long long b = pos_begin[q], e = pos_end[q];
    int kept = 0;
    for (int j0 = 0; j0 < k && kept < num_negative; j0 += 32) {
        const int j = j0 + lane;
        const long long d = (j < k) ? ids[q * k + j] : -1;
        const bool keep = d >= 0 && !(d >= b && d < e);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pos = kept + __popc(m & ((1u << lane) - 1u));
        if (keep && pos < num_negative) out[q * num_negative + pos] = d;
        kept += __popc(m);
    }
    kept = min(kept, num_negative);
    for (int j = kept + lane; j < num_negative; j += 32) out[q * num_negative + j] = -1;
}

}  // namespace drt
