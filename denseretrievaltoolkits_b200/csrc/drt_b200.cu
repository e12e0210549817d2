// drt_b200.cu — C ABI (include/drt_b200.h) and host-side orchestration of the B200 exact-MIPS
// hot path.  Device code lives in the .cuh files next to this one.  sm_100a only; every entry
// point fails with an error (never a CPU fallback) when no such device is present.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/drt_b200.h"
#include "inbatch_ce.cuh"
#include "mips_filter.cuh"
#include "gemm_tc.cuh"
#include "gemm_tc_small.cuh"
#include "select_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            (void)cudaGetLastError();                                                       \
            return fail(_e == cudaErrorMemoryAllocation ? DRT_E_OOM : DRT_E_CUDA,           \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                        __LINE__);                                                          \
        }                                                                                   \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; (void)cudaGetLastError(); }
        ok = cudaSetDevice(dev) == cudaSuccess;
        if (!ok) (void)cudaGetLastError();
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int check_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        return fail(DRT_E_NO_DEVICE, "no CUDA device available: this library has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(DRT_E_INVALID, "device %d out of range [0,%d)", device, n);
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10)
        return fail(DRT_E_NO_DEVICE, "device %d is sm_%d0, need sm_100 (B200): no fallback path", device, major);
    return DRT_OK;
}

// ---- TMA descriptor creation through the driver entry point (no link-time libcuda dep) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    });
    return fn;
}

// bf16 row-major [rows, dim] matrix, box = [64 k-elements x box_rows rows], 128-byte swizzle
int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t dim, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(DRT_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {dim, rows};
    cuuint64_t gstride[1] = {dim * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DRT_E_CUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
    return DRT_OK;
}

// fp32 row-major [rows, dim] matrix, box = [64 k-elements x box_rows rows], no swizzle (the fused
// loss kernel converts it to bf16 pieces itself)
int make_tmap_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t dim, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(DRT_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {dim, rows};
    cuuint64_t gstride[1] = {dim * 4};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DRT_E_CUDA, "cuTensorMapEncodeTiled (fp32) failed: CUresult %d", (int)r);
    return DRT_OK;
}

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return DRT_OK;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        size_t want = need + need / 8;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            e = cudaMalloc(&p, need);
            want = need;
        }
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            p = nullptr;
            return fail(DRT_E_OOM, "cudaMalloc of %zu bytes failed: %s", need, cudaGetErrorString(e));
        }
        bytes = want;
        return DRT_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }
inline bool aligned16_ptr(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// =============================================================================================
// One corpus segment: row r of the store lives in segment r / seg_rows.  Every segment but the
// last holds exactly seg_rows rows; the last one is allocated for the rows it has to hold and
// grows geometrically up to seg_rows (a 10-row index does not pin a 2^20-row segment).
struct Segment {
    float* f32 = nullptr;           // [cap][dim] exact plane (rescoring, reconstruct, write_index)
    void* bf16 = nullptr;           // [cap][dim] tensor-core plane
    float4* bound = nullptr;        // [cap] per-row error-bound entries (r, Dx, Dt, 0)
    unsigned int* tile = nullptr;   // [seg_rows/256][4] per-tile maxima of the above (float bits)
    int64_t cap = 0;                // rows allocated (multiple of 256, <= seg_rows)
};

struct drt_store {
    int dim = 0;          // row pitch in elements: the caller's dim rounded up to a multiple of 64, zero padded
    int dim_user = 0;     // the caller's embedding dim (rows / queries / reconstruct use this pitch)
    int split = 0;        // dims [split, dim) are the caller-declared exactly-representable tail (default: none)
    int f16 = 0;          // 16-bit planes are bf16 (default) or IEEE fp16 (DRT_B200_FIRST_PASS=f16), fixed at creation
    int device = 0;
    int64_t seg_rows = 0;
    int64_t ntotal = 0;
    int sm_count = 148;
    std::vector<Segment> segs;
    std::vector<float*> seg_f32;    // device pointer tables as the kernels take them (mirrors of segs)
    std::vector<float4*> seg_bound;
    std::vector<float4*> seg_tile;
    bool heavy_valid = false;       // per-segment `heavy` flags (select_kernel) are current
    bool any_heavy = true;          // host copy: some segment is heavy (select tightens bounds per row)
    double margin_scale = 1.0;      // widens k' after searches in which the certificate flagged many queries
    // search workspace (grow-only)
    DevBuf q_bf16, q_f32, thr, cnt, cand, seg_table, bound_table, tile_table, heavy, seg_valid, qbound, sel_scratch, out_scores, out_ids, misc;
    DevBuf qflag, sub_idx, sub_q, sub_os, sub_oi, sub_flag;   // exactness-check fallback
    int* err_host = nullptr;     // pinned + mapped: kernel watchdog code
    int* err_dev = nullptr;
    int64_t* misc_host = nullptr;  // pinned: [0] overflow flag [1] flagged count
    int64_t stats[12] = {0};
    bool attrs_set = false;
    std::vector<cudaEvent_t> ev;   // per-launch timing events (DRT_SEARCH_TIME_KERNELS)
    int64_t exact_queries = 0;     // queries of the last search that needed the exact fp32 first pass
    unsigned char* async_status = nullptr;   // drt_search_async: device byte that receives "redo needed"
    cudaEvent_t async_done = nullptr;         // recorded behind an asynchronous search
    size_t async_timed = 0;                   // K1 launches of it that were bracketed by timing events
    bool async_pending = false;               // its counters have not been collected yet
    mutable std::mutex mu;         // add / search / reset / reconstruct on one store are serialised
};

namespace {

// k' = candidates kept from the bf16 pass.  The margin k' - k must make the score gap between
// exact rank k and bf16 rank k' exceed ~6 sigma of the bf16 score error; for a Gaussian-like
// score tail the gap of m ranks at rank k is ~ m / k of the tail scale, so the margin grows
// with k (mid-range k needs relatively more because the gap of few ranks fluctuates more).
// Queries for which the a-posteriori check still fails are refined with a doubled k'.
//
// With the rigorous certificate (mips_filter.cuh) the margin has to cover the worst-case error
// bound E instead of the typical error (sqrt(dim) smaller): the gap between the exact k-th
// score and the k'-th upper bound behaves like sigma_tail * ln(k'/k) with a spread of
// sigma_tail * sqrt((k'-k) / (k k')) (top order statistics of a light tail), and must exceed E.
// rho = E / sigma_tail is ~0.4 for 768-d Gaussian-like embeddings at N ~ 1e7 (E ~ 2.5 at score
// scale 27.7: 2 |q| |d| 2^-8 / sqrt(6) + accumulation); k' is the smallest count with
// ln(k'/k) >= rho + 4 sqrt((k'-k)/(k k')).  `scale` widens rho for stores whose searches flagged.
bool first_pass_f16_default() {
    // bf16 by default: fp16 images give an 8x tighter certificate (k' 140 instead of 200 at k=100)
    // but the wider multipliers draw more power, and under the 1000 W cap the tensor-bound pass ran
    // 4.6 % slower (SM clock 1.22 vs 1.29 GHz, same box, profiles/r2_results.md) — a net loss at
    // large Q.  DRT_B200_FIRST_PASS=f16 selects fp16 (a gain for HBM-bound small-Q calls).
    static const bool f16 = [] { const char* e = getenv("DRT_B200_FIRST_PASS"); return e && !strcmp(e, "f16"); }();
    return f16;
}

int kprime_for(int k, double scale = 1.0, int f16 = -1) {
    static const double rho_env = [] { const char* e = getenv("DRT_B200_KPRIME_RHO"); return e ? atof(e) : -1.0; }();
    if (f16 < 0) f16 = first_pass_f16_default() ? 1 : 0;
    // E / sigma_tail for 768-d Gaussian-like data: ~0.4 with bf16 images (E ~ 2.6 at score scale
    // 27.7), ~0.07 with fp16 images (E ~ 0.45); fp16 keeps a wider safety factor for real data
    const double rho0 = rho_env >= 0.0 ? rho_env : (f16 ? 0.12 : 0.40);
    const double rho = rho0 * scale;
    int kp = k + std::max(28, k / 5);
    for (; kp < 8192; kp += 4) {
        const double m = kp - k;
        if (std::log((double)kp / k) >= rho + 4.0 * std::sqrt(m / ((double)k * kp))) break;
    }
    return std::min(8192, (kp + 3) & ~3);
}

// CTA-pair tiles (M=256) halve the corpus-operand smem/L2 traffic per FLOP and run ~10 % faster
// per padded query row than M=128 tiles (profiles/r1_q_sweep.md), but pad the query count to a
// multiple of 256: pick the variant with the smaller padded cost.  With <= 128 queries half of
// a pair tile would be padding and the pass is HBM-bound anyway.
// Candidate buffer slots per query (8 B each) and the part of it select_kernel handles in shared
// memory: 12 B per key, so 4096 keys leave room for four selecting CTAs per SM.
constexpr int kCandCap = 16384;
int select_capacity(int keep) { return keep <= 1024 ? 4096 : keep <= 2730 ? 8192 : 16384; }

int default_ctas(int64_t nq) {
    const int64_t pad1 = (nq + drt::kTileM - 1) / drt::kTileM * drt::kTileM;
    const int64_t pad2 = (nq + 2 * drt::kTileM - 1) / (2 * drt::kTileM) * (2 * drt::kTileM);
    return (10 * pad2 <= 11 * pad1) ? 2 : 1;
}

int set_kernel_attrs(drt_store* s) {
    if (s->attrs_set) return DRT_OK;
    CUDA_TRY(cudaFuncSetAttribute(drt::mips_filter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)drt::FilterCfg<1>::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(drt::mips_filter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)drt::FilterCfg<2>::kSmemBytes));
    CUDA_TRY(cudaFuncSetAttribute(drt::select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(drt::select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(drt::rescore_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));   // 8192 keys + 8192 dims
    CUDA_TRY(cudaFuncSetAttribute(drt::rescore_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    s->attrs_set = true;
    return DRT_OK;
}

template <int kCtas>
int launch_filter(const CUtensorMap& tq, const CUtensorMap& td, const drt::FilterParams& p, int sm_count,
                  cudaStream_t st) {
    const int n_groups = (p.n_tile_count + p.unit_tiles - 1) / p.unit_tiles;
    const int total = p.num_m_tiles * n_groups;                   // work units
    int clusters = std::min(sm_count / kCtas, total);
    if (clusters < 1) clusters = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * kCtas);
    cfg.blockDim = dim3(drt::kFilterThreads);
    cfg.dynamicSmemBytes = drt::FilterCfg<kCtas>::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, drt::mips_filter_kernel<kCtas>, tq, td, p));
    return DRT_OK;
}

void free_segment(Segment& g) {
    if (g.f32) cudaFree(g.f32);
    if (g.bf16) cudaFree(g.bf16);
    if (g.bound) cudaFree(g.bound);
    if (g.tile) cudaFree(g.tile);
    g = Segment();
}

int alloc_segment(const drt_store* s, Segment& g, int64_t cap) {
    cudaError_t e = cudaMalloc((void**)&g.f32, (size_t)cap * s->dim * 4);
    if (e == cudaSuccess) e = cudaMalloc(&g.bf16, (size_t)cap * s->dim * 2);
    if (e == cudaSuccess) e = cudaMalloc((void**)&g.bound, (size_t)cap * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc((void**)&g.tile, (size_t)(s->seg_rows / 256) * 16);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        free_segment(g);
        return fail(DRT_E_OOM, "allocating a corpus segment of %lld rows x %d failed: %s", (long long)cap, s->dim,
                    cudaGetErrorString(e));
    }
    g.cap = cap;
    return DRT_OK;
}

// Make segment `seg` able to hold rows [0, need) (need <= seg_rows); rows [0, used) are live.
// New segments are sized for what they must hold (at least 4096 rows, whole segments from a
// quarter segment up); an under-sized last segment grows geometrically, moving its live rows.
int ensure_segment(drt_store* s, int64_t seg, int64_t used, int64_t need, cudaStream_t st) {
    auto round_cap = [&](int64_t rows) {
        int64_t c = std::max<int64_t>(4096, (rows + 255) / 256 * 256);
        if (4 * c >= s->seg_rows) c = s->seg_rows;
        return std::min(c, s->seg_rows);
    };
    if ((size_t)seg == s->segs.size()) {
        Segment g;
        int rc = alloc_segment(s, g, round_cap(need));
        if (rc != DRT_OK) return rc;
        s->segs.push_back(g);
        s->seg_f32.push_back(g.f32);
        s->seg_bound.push_back(g.bound);
        s->seg_tile.push_back((float4*)g.tile);
    }
    Segment& cur = s->segs[seg];
    if (cur.cap < need) {
        Segment g;
        int rc = alloc_segment(s, g, round_cap(std::max(need, 2 * cur.cap)));
        if (rc != DRT_OK) return rc;
        if (used > 0) {
            CUDA_TRY(cudaMemcpyAsync(g.f32, cur.f32, (size_t)used * s->dim * 4, cudaMemcpyDeviceToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(g.bf16, cur.bf16, (size_t)used * s->dim * 2, cudaMemcpyDeviceToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(g.bound, cur.bound, (size_t)used * sizeof(float4), cudaMemcpyDeviceToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(g.tile, cur.tile, (size_t)(s->seg_rows / 256) * 16, cudaMemcpyDeviceToDevice, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));     // the old buffers may still be read by queued work
        free_segment(cur);
        cur = g;
        s->seg_f32[seg] = g.f32;
        s->seg_bound[seg] = g.bound;
        s->seg_tile[seg] = (float4*)g.tile;
    }
    // a segment that starts (again) at row 0 -- new, or reused after reset -- has no tile maxima yet
    if (used == 0) CUDA_TRY(cudaMemsetAsync(cur.tile, 0, (size_t)(s->seg_rows / 256) * 16, st));
    return DRT_OK;
}

struct Chunk { int seg; int64_t row0, row1; };   // rows relative to the segment

// Corpus chunk schedule.  Admission thresholds are refreshed between chunks, so chunk sizes
// grow geometrically: once `seen` rows have been scanned the threshold sits at rank ~keep of
// `seen`, and a chunk of (growth-1)*seen further rows is expected to admit ~(growth-1)*keep
// candidates per query, which must fit the candidate buffer.  attempt 2 uses fixed chunks of
// (cap - keep) rows, which cannot overflow whatever the data order.
// `sel` = candidates select_kernel handles in shared memory (its fast path), `cap` = slots of a
// query's candidate buffer (>= sel; the rest is headroom against overflow retries).
std::vector<Chunk> plan_chunks(int64_t ntotal, int64_t seg_rows, int sel, int cap, int keep, int attempt) {
    std::vector<Chunk> out;
    std::vector<int64_t> bounds;
    // Every row of the first chunk is admitted (thresholds start at -FLT_MAX), which costs the
    // epilogue's slow path per row: keep it at ~4 k' rows (enough for a k'-th score to exist).
    // Later chunks grow by up to 16x: expected admissions (growth-1) k' stay below half of `sel`.
    int64_t first = std::max<int64_t>(256, (sel / 2) / 256 * 256);
    int64_t growth = std::max<int64_t>(2, std::min<int64_t>(16, 1 + (sel - keep) / (2 * (int64_t)keep)));
    if (attempt == 0) {
        // smallest first chunk (>= ~4 k') that still reaches the end of the first segment in as
        // few growth steps as the largest admissible one (cap / 2) would
        const int64_t target = std::min<int64_t>(ntotal, seg_rows);
        int64_t reach = first, gpow = 1;
        while (reach < target) { reach *= growth; gpow *= growth; }
        const int64_t need = ((target + gpow - 1) / gpow + 255) / 256 * 256;
        const int64_t lower = (4 * (int64_t)keep + 255) / 256 * 256;
        first = std::min<int64_t>(first, std::max<int64_t>(need, lower));
    }
    if (attempt == 1) growth = 2;
    const int64_t fixed = std::max<int64_t>(256, ((int64_t)(cap - keep)) / 256 * 256);
    int64_t b = 0;
    while (b < ntotal) {
        int64_t nb;
        if (b == 0) nb = first;
        else if (attempt >= 2) nb = b + fixed;
        else nb = b * growth;
        nb = std::min(nb, ntotal);
        // a chunk that starts inside a segment ends at that segment's end at the latest (no
        // sliver launches behind the boundary); later chunks are whole segments
        if (b % seg_rows != 0) nb = std::min(nb, (b / seg_rows + 1) * seg_rows);
        // split at segment boundaries
        int64_t a = b;
        while (a < nb) {
            const int64_t seg = a / seg_rows;
            const int64_t e = std::min(nb, (seg + 1) * seg_rows);
            out.push_back({(int)seg, a - seg * seg_rows, e - seg * seg_rows});
            a = e;
        }
        b = nb;
    }
    return out;
}

// One query batch, all on device.  Returns DRT_OK, an error, or +1 = candidate overflow (retry).
int search_batch(drt_store* s, const float* q_dev, int64_t nq, int k, float* out_s, int64_t* out_i,
                 int64_t id_offset, uint32_t flags, cudaStream_t st, int attempt, int kctas,
                 int keep_override, unsigned char* qflag, int64_t* flagged_out) {
    const int dim = s->dim;
    const int keep = keep_override > 0 ? keep_override : kprime_for(k, s->margin_scale, s->f16);
    const int sel = select_capacity(keep);
    const int cap = kCandCap;
    if (cap < 2 * keep) return fail(DRT_E_UNSUPPORTED, "k=%d too large for the candidate buffer", k);

    int rc;
    // the 16-bit query plane is padded with zero rows to whole M tiles: a TMA box that hangs over
    // the end of the tensor is filled with zeros by the copy engine, but measurably slower than a
    // box it reads (K1 over 8.8M rows: 2.66-2.77 ms at 16 queries against 2.18 ms at 128)
    const int64_t m_tile_rows = (int64_t)drt::kTileM * std::max(kctas, 1);
    const int64_t nq_pad = (nq + m_tile_rows - 1) / m_tile_rows * m_tile_rows;
    if ((rc = s->q_bf16.ensure((size_t)nq_pad * dim * 2)) != DRT_OK) return rc;
    if ((rc = s->thr.ensure((size_t)nq * 4)) != DRT_OK) return rc;
    if ((rc = s->cnt.ensure((size_t)nq * 4)) != DRT_OK) return rc;
    if ((rc = s->cand.ensure((size_t)nq * cap * 8)) != DRT_OK) return rc;
    if ((rc = s->misc.ensure(64)) != DRT_OK) return rc;
    if ((rc = s->seg_table.ensure(std::max<size_t>(8, s->seg_f32.size() * sizeof(float*)))) != DRT_OK) return rc;
    if ((rc = s->bound_table.ensure(std::max<size_t>(8, s->seg_bound.size() * sizeof(float4*)))) != DRT_OK) return rc;
    if ((rc = s->tile_table.ensure(std::max<size_t>(8, s->seg_tile.size() * sizeof(float4*)))) != DRT_OK) return rc;
    if ((rc = s->heavy.ensure(std::max<size_t>(8, s->segs.size()))) != DRT_OK) return rc;
    if ((rc = s->seg_valid.ensure(std::max<size_t>(8, s->segs.size() * 8))) != DRT_OK) return rc;
    if ((rc = s->qbound.ensure((size_t)nq * sizeof(float4))) != DRT_OK) return rc;
    if ((rc = s->sel_scratch.ensure((size_t)nq * 2 * keep * 8)) != DRT_OK) return rc;

    float* thr = (float*)s->thr.p;
    uint32_t* cnt = (uint32_t*)s->cnt.p;
    uint64_t* cand = (uint64_t*)s->cand.p;
    int* overflow = (int*)s->misc.p;
    unsigned long long* flagged = (unsigned long long*)((char*)s->misc.p + 8);

    CUDA_TRY(cudaMemsetAsync(s->misc.p, 0, 64, st));
    CUDA_TRY(cudaMemcpyAsync(s->seg_table.p, s->seg_f32.data(), s->seg_f32.size() * sizeof(float*),
                             cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->bound_table.p, s->seg_bound.data(), s->seg_bound.size() * sizeof(float4*),
                             cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->tile_table.p, s->seg_tile.data(), s->seg_tile.size() * sizeof(float4*),
                             cudaMemcpyHostToDevice, st));
    const bool exact_pass = (kctas == 0);    // fp32 SIMT first pass (last-resort refinement)
    const int q_f16 = s->f16;
    const float4* qbound = (const float4*)s->qbound.p;
    const float4* const* bound_table = (const float4* const*)s->bound_table.p;
    const float4* const* tile_table = (const float4* const*)s->tile_table.p;
    if (!s->heavy_valid) {     // rows were added since the last search: refresh the per-segment flags
        std::vector<long long> valid(s->segs.size());
        for (size_t g = 0; g < valid.size(); ++g)
            valid[g] = std::max<int64_t>(0, std::min<int64_t>(s->seg_rows, s->ntotal - (int64_t)g * s->seg_rows));
        CUDA_TRY(cudaMemcpyAsync(s->seg_valid.p, valid.data(), valid.size() * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));    // `valid` is a stack-lived pageable source
        drt::segment_heavy_kernel<<<(int)s->segs.size(), 256, 0, st>>>(tile_table, (int)(s->seg_rows / 256),
                                                                      (const long long*)s->seg_valid.p, (unsigned char*)s->heavy.p);
        std::vector<unsigned char> hv(s->segs.size());
        CUDA_TRY(cudaMemcpyAsync(hv.data(), s->heavy.p, hv.size(), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        s->any_heavy = false;
        for (unsigned char h : hv) s->any_heavy = s->any_heavy || h != 0;
        s->heavy_valid = true;
        s->stats[0] += 1;
    }
    {
        // error-model constants of the certificate (mips_filter.cuh): the tensor core accumulates
        // dim/16 K=16 steps, each at worst 18 truncations of 2^-23 relative to the running
        // magnitude; the rescoring dot (K2) rounds dim/128 FMAs + 7 adds per lane chain.
        const float c_acc = (float)(dim / 16) * 18.f * 0x1p-23f;
        const float c_k2 = (float)(dim / 128 + 8) * 0x1p-23f;
        const int blocks = (int)std::min<int64_t>((nq + 7) / 8, (int64_t)s->sm_count * 8);
        drt::prep_queries_kernel<<<blocks, 256, 0, st>>>(q_dev, (int)nq, dim, s->split, exact_pass ? nullptr : (uint2*)s->q_bf16.p,
                                                        (float4*)s->qbound.p, thr, cnt, c_acc, c_k2, exact_pass ? 1 : 0, q_f16);
        s->stats[0] += 1;
    }
    CUtensorMap tmap_q;
    if (!exact_pass) {
        if (nq_pad > nq)
            CUDA_TRY(cudaMemsetAsync((char*)s->q_bf16.p + (size_t)nq * dim * 2, 0, (size_t)(nq_pad - nq) * dim * 2, st));
        if ((rc = make_tmap_bf16(&tmap_q, s->q_bf16.p, (uint64_t)nq_pad, (uint64_t)dim, drt::kTileM)) != DRT_OK) return rc;
    }

    const std::vector<Chunk> chunks = plan_chunks(s->ntotal, s->seg_rows, sel, cap, keep, attempt);
    if (keep_override == 0 && !exact_pass) s->stats[6] = (int64_t)chunks.size();
    // `expect` = candidates a query is expected to hold at this select.  The shared-memory staging
    // area is sized for ~1.5x that (not for the worst case `sel`): more selecting CTAs fit an SM,
    // and a query that gathered more takes the kernel's global-memory path.
    auto launch_select = [&](double expect, bool final_select) {
        // between chunks the select may keep up to 2 k' candidates (it then stops after one digit
        // pass); the last one before K2 is exact
        // (and so is every select of the overflow-retry attempts: their chunk sizes count on k' survivors)
        const uint32_t max_keep = (final_select || attempt > 0) ? (uint32_t)keep : (uint32_t)(2 * keep);
        int keys = 1024;
        while (keys < sel && keys < 1.5 * expect + 256) keys *= 2;
        keys = std::min(keys, sel);
        if (s->any_heavy)
            drt::select_kernel<true><<<(int)nq, 256, (size_t)keys * 12, st>>>(
                cand, cnt, thr, (uint32_t)cap, (uint32_t)keep, overflow, qbound, bound_table, tile_table,
                (const unsigned char*)s->heavy.p, (uint32_t)s->seg_rows, (uint32_t)keys, (uint64_t*)s->sel_scratch.p, max_keep);
        else
            drt::select_kernel<false><<<(int)nq, 256, (size_t)keys * 8, st>>>(
                cand, cnt, thr, (uint32_t)cap, (uint32_t)keep, overflow, qbound, bound_table, tile_table,
                (const unsigned char*)s->heavy.p, (uint32_t)s->seg_rows, (uint32_t)keys, (uint64_t*)s->sel_scratch.p, max_keep);
        s->stats[0] += 1;
    };
    int64_t seen_at_select = 0;      // rows scanned when the thresholds were last refreshed
    CUtensorMap tmap_d;
    int tmap_seg = -1;
    size_t n_timed = 0;
    bool frozen = false;
    for (const Chunk& c : chunks) {
        const int64_t seg_valid = std::min<int64_t>(s->seg_rows, s->ntotal - (int64_t)c.seg * s->seg_rows);
        if (exact_pass) {
            using C = drt::GemmLarge;
            const int64_t nrows = c.row1 - c.row0;
            const float* rows = s->seg_f32[c.seg] + (size_t)c.row0 * dim;
            dim3 grid((unsigned)((nrows + C::BN - 1) / C::BN), (unsigned)((nq + C::BM - 1) / C::BM));
            drt::exact_filter_kernel<C><<<grid, C::THREADS, 0, st>>>(
                q_dev, (long long)nq, rows, (long long)nrows, dim, (uint32_t)((int64_t)c.seg * s->seg_rows + c.row0), thr,
                cnt, cand, (uint32_t)cap, aligned16_ptr(q_dev) ? 1 : 0, qbound, s->seg_bound[c.seg] + c.row0);
            launch_select((double)sel, &c == &chunks.back());
            s->stats[0] += 1;
            continue;
        }
        if (c.seg != tmap_seg) {
            if ((rc = make_tmap_bf16(&tmap_d, s->segs[c.seg].bf16, (uint64_t)seg_valid, (uint64_t)dim,
                                     drt::kTileN / kctas)) != DRT_OK) return rc;
            tmap_seg = c.seg;
        }
        drt::FilterParams p;
        p.num_m_tiles = (int)((nq + drt::kTileM * kctas - 1) / (drt::kTileM * kctas));
        p.n_tile_begin = (int)(c.row0 / drt::kTileN);
        p.n_tile_count = (int)((c.row1 + drt::kTileN - 1) / drt::kTileN) - p.n_tile_begin;
        // corpus tiles per work unit: 16 when the launch is large, fewer for the small early
        // chunks so that every CTA (pair) still gets at least ~4 units
        p.unit_tiles = 16;
        if (const char* e = getenv("DRT_B200_UNIT_TILES")) { const int v = atoi(e); if (v >= 1 && v <= 256) p.unit_tiles = v; }
        while (p.unit_tiles > 1 &&
               (int64_t)p.num_m_tiles * ((p.n_tile_count + p.unit_tiles - 1) / p.unit_tiles) < 4ll * (s->sm_count / kctas))
            p.unit_tiles /= 2;
        p.num_k_blocks = dim / drt::kBlockK;
        p.nq = (int)nq;
        p.f16 = s->f16;
        p.rows_valid = (uint32_t)c.row1;    // rows past this chunk's end are not admitted yet
        p.row_base = (uint32_t)((int64_t)c.seg * s->seg_rows);
        p.cap = (uint32_t)cap;
        p.thr = thr; p.cnt = cnt; p.cand = cand; p.err = s->err_dev;
        p.qbound = qbound;
        p.tile_bound = (const float4*)s->segs[c.seg].tile;
        p.row_bound = s->segs[c.seg].bound;
        const bool timed = (flags & DRT_SEARCH_TIME_KERNELS) != 0;
        if (timed) {
            while (s->ev.size() < 2 * (n_timed + 1)) {
                cudaEvent_t e;
                CUDA_TRY(cudaEventCreate(&e));
                s->ev.push_back(e);
            }
            CUDA_TRY(cudaEventRecord(s->ev[2 * n_timed], st));
        }
        rc = (kctas == 2) ? launch_filter<2>(tmap_q, tmap_d, p, s->sm_count, st)
                          : launch_filter<1>(tmap_q, tmap_d, p, s->sm_count, st);
        if (rc != DRT_OK) return rc;
        if (timed) { CUDA_TRY(cudaEventRecord(s->ev[2 * n_timed + 1], st)); ++n_timed; }
        // Trim + publish thresholds after the chunk -- unless the thresholds are already so tight
        // that everything the REST of the corpus is expected to admit (keep * remaining / seen
        // per query, thresholds frozen) fits the buffer many times over; the last chunk always
        // trims, because K2 reads at most `keep` candidates.
        const int64_t seen = (int64_t)c.seg * s->seg_rows + c.row1;
        const bool last = (&c == &chunks.back());
        if (!frozen || last) {
            // candidates held now: all rows of the first chunk; later the k' survivors plus what the
            // rows since the last refresh admitted against that (by now stale) threshold,
            // ~k' (seen - seen_then) / seen_then
            // (x1.25 + k': an early-stopped select leaves up to 2 k' candidates and a slightly lower threshold)
            const double expect = seen_at_select == 0 ? (double)std::min<int64_t>(seen, cap)
                                                      : 1.25 * (double)keep * (double)seen / (double)seen_at_select + keep;
            launch_select(attempt == 0 ? expect : (double)sel, last);
            seen_at_select = seen;
            // thresholds frozen: the final select must still find its candidates in shared memory
            if (attempt == 0 && (double)keep * (double)(s->ntotal - seen) / (double)seen < (double)(sel - keep) / 2.0) frozen = true;
        }
        s->stats[0] += 1;
        s->stats[1] += 1;
    }
    {
        const size_t smem = (size_t)next_pow2(keep) * 8 + (size_t)dim * 4;
        // few queries: one CTA of 32 warps per query (latency-bound gathers); many: 8 warps, 6 CTAs per SM
        if (nq <= 2 * (int64_t)s->sm_count)
            drt::rescore_kernel<1024><<<(int)nq, 1024, smem, st>>>(cand, cnt, (uint32_t)cap, (uint32_t)keep, q_dev, dim,
                                                       (const float* const*)s->seg_table.p, (uint32_t)s->seg_rows, k,
                                                       (long long)id_offset, out_s, (long long*)out_i,
                                                       (flags & DRT_SEARCH_NO_RESCORE) ? 0 : 1, flagged, qflag,
                                                       exact_pass ? 0 : 1, thr, qbound);
        else
            drt::rescore_kernel<256><<<(int)nq, 256, smem, st>>>(cand, cnt, (uint32_t)cap, (uint32_t)keep, q_dev, dim,
                                                       (const float* const*)s->seg_table.p, (uint32_t)s->seg_rows, k,
                                                       (long long)id_offset, out_s, (long long*)out_i,
                                                       (flags & DRT_SEARCH_NO_RESCORE) ? 0 : 1, flagged, qflag,
                                                       exact_pass ? 0 : 1, thr, qbound);
        s->stats[0] += 1;
    }
    CUDA_TRY(cudaGetLastError());
    if (s->async_status) {
        // asynchronous search: no host round trip.  "Result not final" (candidate overflow, or a
        // query the certificate flagged) is published on the device for the caller's merge step.
        drt::publish_status_kernel<<<1, 32, 0, st>>>((const int*)s->misc.p, (const unsigned long long*)((char*)s->misc.p + 8), s->async_status);
        CUDA_TRY(cudaMemcpyAsync(s->misc_host, s->misc.p, 24, cudaMemcpyDeviceToHost, st));
        if (!s->async_done) CUDA_TRY(cudaEventCreateWithFlags(&s->async_done, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(s->async_done, st));
        if (keep_override == 0 && !exact_pass) { s->stats[3] = keep; s->stats[5] = kctas; }
        s->async_timed = n_timed;
        s->async_pending = true;
        s->stats[0] += 1;
        return DRT_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(s->misc_host, s->misc.p, 24, cudaMemcpyDeviceToHost, st));
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        const int code = s->err_host ? *s->err_host : 0;
        return fail(code ? DRT_E_INTERNAL : DRT_E_CUDA, "search kernels failed: %s (watchdog code %d)",
                    cudaGetErrorString(e), code);
    }
    if (keep_override == 0 && !exact_pass) { s->stats[3] = keep; s->stats[5] = kctas; }
    for (size_t i = 0; i < n_timed; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s->ev[2 * i], s->ev[2 * i + 1]) == cudaSuccess) s->stats[7] += (int64_t)(ms * 1e6);
    }
    if ((int)(s->misc_host[0] & 0xffffffff) != 0) return 1;   // overflow -> caller retries
    if (flagged_out) *flagged_out = (int64_t)s->misc_host[1];
    s->stats[10] += (int64_t)s->misc_host[2];
    return DRT_OK;
}

// search_batch + the overflow retry ladder (bigger buffer / growth 2, then fixed chunks)
int search_retrying(drt_store* s, const float* q_dev, int64_t nq, int k, float* out_s, int64_t* out_i,
                    int64_t id_offset, uint32_t flags, cudaStream_t st, int kctas, int keep_override,
                    unsigned char* qflag, int64_t* flagged_out) {
    for (int attempt = 0;; ++attempt) {
        const int rc = search_batch(s, q_dev, nq, k, out_s, out_i, id_offset, flags, st, attempt, kctas,
                                    keep_override, qflag, flagged_out);
        if (rc <= 0) return rc;
        s->stats[2] += 1;
        if (attempt >= 2) return fail(DRT_E_INTERNAL, "candidate buffer overflow persisted after retries");
    }
}

// Queries whose exactness check flagged are searched again with a doubled candidate count k'
// (up to 3 rounds); what is still flagged afterwards is reported in stats[4].
int refine_flagged(drt_store* s, const float* q_dev, int64_t nq, int k, float* out_s, int64_t* out_i,
                   int64_t id_offset, uint32_t flags, cudaStream_t st, unsigned char* qflag, int64_t* flagged) {
    int keep = kprime_for(k, s->margin_scale, s->f16);
    std::vector<unsigned char> hflag;
    std::vector<int> idx;
    for (int round = 0; *flagged > 0 && round < 3 && keep * 2 <= 8192; ++round) {
        keep *= 2;
        hflag.resize((size_t)nq);
        CUDA_TRY(cudaMemcpyAsync(hflag.data(), qflag, (size_t)nq, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        idx.clear();
        for (int64_t i = 0; i < nq; ++i) if (hflag[i]) idx.push_back((int)i);
        const int64_t nf = (int64_t)idx.size();
        if (nf == 0) break;
        int rc;
        if ((rc = s->sub_idx.ensure((size_t)nf * 4)) != DRT_OK) return rc;
        if ((rc = s->sub_q.ensure((size_t)nf * s->dim * 4)) != DRT_OK) return rc;
        if ((rc = s->sub_os.ensure((size_t)nf * k * 4)) != DRT_OK) return rc;
        if ((rc = s->sub_oi.ensure((size_t)nf * k * 8)) != DRT_OK) return rc;
        if ((rc = s->sub_flag.ensure((size_t)nf)) != DRT_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(s->sub_idx.p, idx.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
        drt::gather_rows_kernel<<<(int)std::min<int64_t>(nf, 4096), 192, 0, st>>>(
            q_dev, (const int*)s->sub_idx.p, (float*)s->sub_q.p, s->dim, (int)nf);
        int64_t still = 0;
        const int kctas = (flags & DRT_SEARCH_FORCE_1CTA) ? 1 : (flags & DRT_SEARCH_FORCE_2CTA) ? 2 : default_ctas(nf);
        rc = search_retrying(s, (const float*)s->sub_q.p, nf, k, (float*)s->sub_os.p, (int64_t*)s->sub_oi.p, id_offset,
                             flags, st, kctas, keep, (unsigned char*)s->sub_flag.p, &still);
        if (rc != DRT_OK) return rc;
        drt::scatter_results_kernel<<<(int)std::min<int64_t>(nf, 4096), 128, 0, st>>>(
            (const float*)s->sub_os.p, (const long long*)s->sub_oi.p, (const unsigned char*)s->sub_flag.p,
            (const int*)s->sub_idx.p, out_s, (long long*)out_i, qflag, k, (int)nf);
        CUDA_TRY(cudaGetLastError());
        s->stats[0] += 2;
        *flagged = still;
    }
    return DRT_OK;
}

// Last resort for queries the bf16 pass cannot certify: first pass in exact fp32 (SIMT).
int exact_flagged(drt_store* s, const float* q_dev, int64_t nq, int k, float* out_s, int64_t* out_i,
                  int64_t id_offset, uint32_t flags, cudaStream_t st, unsigned char* qflag, int64_t* flagged) {
    std::vector<unsigned char> hflag((size_t)nq);
    CUDA_TRY(cudaMemcpyAsync(hflag.data(), qflag, (size_t)nq, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::vector<int> idx;
    for (int64_t i = 0; i < nq; ++i) if (hflag[i]) idx.push_back((int)i);
    const int64_t nf = (int64_t)idx.size();
    if (nf == 0) { *flagged = 0; return DRT_OK; }
    int rc;
    if ((rc = s->sub_idx.ensure((size_t)nf * 4)) != DRT_OK) return rc;
    if ((rc = s->sub_q.ensure((size_t)nf * s->dim * 4)) != DRT_OK) return rc;
    if ((rc = s->sub_os.ensure((size_t)nf * k * 4)) != DRT_OK) return rc;
    if ((rc = s->sub_oi.ensure((size_t)nf * k * 8)) != DRT_OK) return rc;
    if ((rc = s->sub_flag.ensure((size_t)nf)) != DRT_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->sub_idx.p, idx.data(), (size_t)nf * 4, cudaMemcpyHostToDevice, st));
    drt::gather_rows_kernel<<<(int)std::min<int64_t>(nf, 4096), 192, 0, st>>>(
        q_dev, (const int*)s->sub_idx.p, (float*)s->sub_q.p, s->dim, (int)nf);
    int64_t still = 0;
    rc = search_retrying(s, (const float*)s->sub_q.p, nf, k, (float*)s->sub_os.p, (int64_t*)s->sub_oi.p, id_offset,
                         flags, st, /*kctas=*/0, 0, (unsigned char*)s->sub_flag.p, &still);
    if (rc != DRT_OK) return rc;
    drt::scatter_results_kernel<<<(int)std::min<int64_t>(nf, 4096), 128, 0, st>>>(
        (const float*)s->sub_os.p, (const long long*)s->sub_oi.p, (const unsigned char*)s->sub_flag.p,
        (const int*)s->sub_idx.p, out_s, (long long*)out_i, qflag, k, (int)nf);
    CUDA_TRY(cudaGetLastError());
    s->stats[0] += 2;
    s->exact_queries += nf;
    *flagged = still;
    return DRT_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int drt_abi_version(void) { return DRT_B200_ABI_VERSION; }
const char* drt_last_error(void) { return g_err.c_str(); }

int drt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

int drt_store_create(drt_store** out, int dim, int device, int64_t seg_rows) {
    if (!out) return fail(DRT_E_INVALID, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return fail(DRT_E_INVALID, "dim must be positive, got %d", dim);
    if (dim > 8192) return fail(DRT_E_UNSUPPORTED, "dim %d exceeds 8192", dim);
    // one TMA box spans 64 bf16 (a 128-byte swizzle row): other dims are stored zero-padded to
    // the next multiple of 64, which changes no inner product
    const int dim_pad = (dim + 63) / 64 * 64;
    if (seg_rows == 0) seg_rows = 1 << 20;
    if (seg_rows < 256 || seg_rows % 256 != 0) return fail(DRT_E_INVALID, "seg_rows must be a positive multiple of 256");
    int rc = check_device(device);
    if (rc != DRT_OK) return rc;
    DeviceGuard g(device);
    if (!g.ok) return fail(DRT_E_CUDA, "cudaSetDevice(%d) failed", device);
    drt_store* s = new (std::nothrow) drt_store();
    if (!s) return fail(DRT_E_OOM, "host allocation failed");
    s->dim = dim_pad; s->dim_user = dim; s->split = dim_pad; s->device = device; s->seg_rows = seg_rows;
    s->f16 = first_pass_f16_default() ? 1 : 0;
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (cudaHostAlloc((void**)&s->err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&s->err_dev, s->err_host, 0) != cudaSuccess ||
        cudaHostAlloc((void**)&s->misc_host, 64, cudaHostAllocDefault) != cudaSuccess) {
        (void)cudaGetLastError();
        delete s;
        return fail(DRT_E_CUDA, "pinned host allocation failed");
    }
    *s->err_host = 0;
    *out = s;
    return DRT_OK;
}

int drt_store_destroy(drt_store* s) {
    if (!s) return DRT_OK;
    DeviceGuard g(s->device);
    for (cudaEvent_t e : s->ev) cudaEventDestroy(e);
    if (s->async_done) cudaEventDestroy(s->async_done);
    for (Segment& g : s->segs) free_segment(g);
    s->q_bf16.release(); s->q_f32.release(); s->thr.release(); s->cnt.release(); s->cand.release();
    s->seg_table.release(); s->bound_table.release(); s->tile_table.release(); s->heavy.release(); s->seg_valid.release(); s->qbound.release(); s->sel_scratch.release(); s->out_scores.release(); s->out_ids.release(); s->misc.release();
    s->qflag.release(); s->sub_idx.release(); s->sub_q.release(); s->sub_os.release(); s->sub_oi.release(); s->sub_flag.release();
    if (s->err_host) cudaFreeHost(s->err_host);
    if (s->misc_host) cudaFreeHost(s->misc_host);
    (void)cudaGetLastError();
    delete s;
    return DRT_OK;
}

int drt_store_add(drt_store* s, const float* rows, int64_t n, int rows_on_device, void* stream) {
    if (!s) return fail(DRT_E_INVALID, "store is NULL");
    if (n < 0 || (n > 0 && !rows)) return fail(DRT_E_INVALID, "bad rows/n");
    if (n == 0) return DRT_OK;
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->ntotal + n > 0xFFFFFF00ll) return fail(DRT_E_UNSUPPORTED, "a shard holds at most 2^32-256 rows");
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t row_f32 = (size_t)s->dim * 4, row_bf16 = (size_t)s->dim * 2;
    int64_t done = 0;
    while (done < n) {
        const int64_t seg = s->ntotal / s->seg_rows, off = s->ntotal % s->seg_rows;
        const int64_t take = std::min(n - done, s->seg_rows - off);
        int rc = ensure_segment(s, seg, off, off + take, st);
        if (rc != DRT_OK) return rc;
        Segment& sg = s->segs[seg];
        float* dst = sg.f32 + (size_t)off * s->dim;
        const cudaMemcpyKind kind = rows_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        if (s->dim_user == s->dim) {
            CUDA_TRY(cudaMemcpyAsync(dst, rows + (size_t)done * s->dim, (size_t)take * row_f32, kind, st));
        } else {   // rows are stored zero-padded to the pitch
            CUDA_TRY(cudaMemsetAsync(dst, 0, (size_t)take * row_f32, st));
            CUDA_TRY(cudaMemcpy2DAsync(dst, row_f32, rows + (size_t)done * s->dim_user, (size_t)s->dim_user * 4,
                                       (size_t)s->dim_user * 4, (size_t)take, kind, st));
        }
        const int blocks = (int)std::min<int64_t>((take + 7) / 8, (int64_t)s->sm_count * 16);
        drt::ingest_rows_kernel<<<blocks, 256, 0, st>>>(dst, (long long)take, s->dim, s->split,
                                                       (uint2*)((char*)sg.bf16 + (size_t)off * row_bf16), sg.bound, sg.tile,
                                                       (long long)off, s->f16);
        CUDA_TRY(cudaGetLastError());
        s->ntotal += take;
        s->heavy_valid = false;
        done += take;
    }
    if (!rows_on_device) CUDA_TRY(cudaStreamSynchronize(st));
    return DRT_OK;
}

int drt_store_set_exact_tail(drt_store* s, int tail_dims) {
    if (!s) return fail(DRT_E_INVALID, "store is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    if (tail_dims < 0 || tail_dims > s->dim_user) return fail(DRT_E_INVALID, "tail_dims %d out of range [0,%d]", tail_dims, s->dim_user);
    if (s->ntotal != 0) return fail(DRT_E_INVALID, "the exact tail must be declared before the first add");
    s->split = tail_dims == 0 ? s->dim : s->dim_user - tail_dims;
    return DRT_OK;
}

int64_t drt_store_ntotal(const drt_store* s) { return s ? s->ntotal : -1; }
int drt_store_dim(const drt_store* s) { return s ? s->dim_user : -1; }
int drt_store_device(const drt_store* s) { return s ? s->device : -1; }

int drt_store_reset(drt_store* s) {
    if (!s) return fail(DRT_E_INVALID, "store is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    s->ntotal = 0;
    s->heavy_valid = false;
    return DRT_OK;
}

int drt_store_reconstruct(const drt_store* s, int64_t row0, int64_t n, float* out, int out_on_device, void* stream) {
    if (!s || (n > 0 && !out)) return fail(DRT_E_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(s->mu);
    if (row0 < 0 || n < 0 || row0 + n > s->ntotal) return fail(DRT_E_INVALID, "rows [%lld,%lld) out of range", (long long)row0, (long long)(row0 + n));
    DeviceGuard g(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    int64_t done = 0;
    while (done < n) {
        const int64_t r = row0 + done, seg = r / s->seg_rows, off = r % s->seg_rows;
        const int64_t take = std::min(n - done, s->seg_rows - off);
        CUDA_TRY(cudaMemcpy2DAsync(out + (size_t)done * s->dim_user, (size_t)s->dim_user * 4,
                                   s->seg_f32[seg] + (size_t)off * s->dim, (size_t)s->dim * 4, (size_t)s->dim_user * 4,
                                   (size_t)take, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        done += take;
    }
    if (!out_on_device) CUDA_TRY(cudaStreamSynchronize(st));
    return DRT_OK;
}

int drt_search(drt_store* s, const float* q, int64_t nq, int k, float* out_scores, int64_t* out_ids,
               int io_on_device, int64_t id_offset, uint32_t flags, void* stream) {
    if (!s) return fail(DRT_E_INVALID, "store is NULL");
    if (nq < 0 || k <= 0) return fail(DRT_E_INVALID, "need nq >= 0 and k > 0 (nq=%lld k=%d)", (long long)nq, k);
    if (k > DRT_MAX_K) return fail(DRT_E_UNSUPPORTED, "k=%d exceeds DRT_MAX_K=%d", k, DRT_MAX_K);
    if (nq == 0) return DRT_OK;
    if (!q || !out_scores || !out_ids) return fail(DRT_E_INVALID, "NULL query/output pointer");
    std::lock_guard<std::mutex> lk(s->mu);
    int rc = check_device(s->device);
    if (rc != DRT_OK) return rc;
    DeviceGuard g(s->device);
    if ((rc = set_kernel_attrs(s)) != DRT_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < 12; ++i) s->stats[i] = 0;
    s->exact_queries = 0;

    int kctas = default_ctas(nq);
    if (const char* e = getenv("DRT_B200_CTAS")) kctas = (atoi(e) == 2) ? 2 : 1;
    if (flags & DRT_SEARCH_FORCE_1CTA) kctas = 1;
    if (flags & DRT_SEARCH_FORCE_2CTA) kctas = 2;

    // Query batches bound the candidate workspace (nq_batch * cap * 8 bytes).
    const int64_t max_batch = 16384;
    for (int64_t q0 = 0; q0 < nq; q0 += max_batch) {
        const int64_t nb = std::min(max_batch, nq - q0);
        const float* q_dev;
        float* os_dev;
        int64_t* oi_dev;
        const bool padded = s->dim_user != s->dim;
        if (io_on_device && !padded && aligned16_ptr(q)) {
            q_dev = q + (size_t)q0 * s->dim;
        } else {
            // host queries, a dim that is stored zero-padded, or a device pointer the 16-byte
            // vector loads cannot take: stage into the (padded) workspace
            if ((rc = s->q_f32.ensure((size_t)nb * s->dim * 4)) != DRT_OK) return rc;
            const cudaMemcpyKind kind = io_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
            if (padded) {
                CUDA_TRY(cudaMemsetAsync(s->q_f32.p, 0, (size_t)nb * s->dim * 4, st));
                CUDA_TRY(cudaMemcpy2DAsync(s->q_f32.p, (size_t)s->dim * 4, q + (size_t)q0 * s->dim_user, (size_t)s->dim_user * 4,
                                           (size_t)s->dim_user * 4, (size_t)nb, kind, st));
            } else {
                CUDA_TRY(cudaMemcpyAsync(s->q_f32.p, q + (size_t)q0 * s->dim, (size_t)nb * s->dim * 4, kind, st));
            }
            q_dev = (const float*)s->q_f32.p;
        }
        if (io_on_device) {
            os_dev = out_scores + (size_t)q0 * k;
            oi_dev = out_ids + (size_t)q0 * k;
        } else {
            if ((rc = s->out_scores.ensure((size_t)nb * k * 4)) != DRT_OK) return rc;
            if ((rc = s->out_ids.ensure((size_t)nb * k * 8)) != DRT_OK) return rc;
            os_dev = (float*)s->out_scores.p;
            oi_dev = (int64_t*)s->out_ids.p;
        }
        if (s->ntotal == 0) {
            drt::fill_outputs_kernel<<<std::max(1, (int)std::min<int64_t>((nb * k + 255) / 256, 4096)), 256, 0, st>>>(
                os_dev, (long long*)oi_dev, (size_t)nb * k);
            CUDA_TRY(cudaGetLastError());
        } else {
            if ((rc = s->qflag.ensure((size_t)nb)) != DRT_OK) return rc;
            int64_t flagged = 0;
            rc = search_retrying(s, q_dev, nb, k, os_dev, oi_dev, id_offset, flags, st, kctas, 0,
                                 (unsigned char*)s->qflag.p, &flagged);
            s->stats[9] += flagged;
            // a store whose searches keep flagging (its score gaps are small against the error
            // bound) gets a wider first-pass margin from now on instead of paying the ladder
            if (flagged > std::max<int64_t>(1, nb / 50) && s->margin_scale < 8.0) s->margin_scale *= 1.5;
            if (rc == DRT_OK && flagged > 0 && !(flags & DRT_SEARCH_NO_RESCORE)) {
                rc = refine_flagged(s, q_dev, nb, k, os_dev, oi_dev, id_offset, flags, st, (unsigned char*)s->qflag.p, &flagged);
                if (rc == DRT_OK && flagged > 0)
                    rc = exact_flagged(s, q_dev, nb, k, os_dev, oi_dev, id_offset, flags, st, (unsigned char*)s->qflag.p, &flagged);
            }
            s->stats[4] += flagged;
            if (rc != DRT_OK) return rc;
        }
        if (!io_on_device) {
            CUDA_TRY(cudaMemcpyAsync(out_scores + (size_t)q0 * k, os_dev, (size_t)nb * k * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(out_ids + (size_t)q0 * k, oi_dev, (size_t)nb * k * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        }
    }
    return DRT_OK;
}

int drt_search_async(drt_store* s, const float* q, int64_t nq, int k, float* out_scores, int64_t* out_ids,
                     int64_t id_offset, uint32_t flags, uint8_t* status_out, void* stream) {
    if (!s) return fail(DRT_E_INVALID, "store is NULL");
    if (nq <= 0 || k <= 0 || k > DRT_MAX_K) return fail(DRT_E_INVALID, "need nq > 0 and 0 < k <= DRT_MAX_K (nq=%lld k=%d)", (long long)nq, k);
    if (!q || !out_scores || !out_ids || !status_out) return fail(DRT_E_INVALID, "NULL query/output/status pointer");
    std::lock_guard<std::mutex> lk(s->mu);
    if (nq > 16384 || s->dim_user != s->dim || !aligned16_ptr(q) || s->ntotal == 0)
        return fail(DRT_E_UNSUPPORTED, "asynchronous search needs 1..16384 queries, a dim that is a multiple of 64, 16-byte aligned queries and a non-empty store");
    int rc = check_device(s->device);
    if (rc != DRT_OK) return rc;
    DeviceGuard g(s->device);
    if ((rc = set_kernel_attrs(s)) != DRT_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (s->async_pending) { CUDA_TRY(cudaEventSynchronize(s->async_done)); s->async_pending = false; }   // misc_host is reused
    for (int i = 0; i < 12; ++i) s->stats[i] = 0;
    s->exact_queries = 0;
    int kctas = default_ctas(nq);
    if (const char* e = getenv("DRT_B200_CTAS")) kctas = (atoi(e) == 2) ? 2 : 1;
    if (flags & DRT_SEARCH_FORCE_1CTA) kctas = 1;
    if (flags & DRT_SEARCH_FORCE_2CTA) kctas = 2;
    if ((rc = s->qflag.ensure((size_t)nq)) != DRT_OK) return rc;
    s->async_status = status_out;
    int64_t flagged = 0;
    rc = search_batch(s, q, nq, k, out_scores, out_ids, id_offset, flags, st, /*attempt=*/0, kctas, 0, (unsigned char*)s->qflag.p, &flagged);
    s->async_status = nullptr;
    return rc;
}

int drt_plan_chunks(int64_t ntotal, int64_t seg_rows, int k, int attempt, int64_t* out, int max_chunks) {
    if (ntotal < 0 || seg_rows < 256 || seg_rows % 256 != 0 || k <= 0 || k > DRT_MAX_K || attempt < 0)
        return fail(DRT_E_INVALID, "bad plan arguments");
    const int keep = kprime_for(k);
    const std::vector<Chunk> chunks = plan_chunks(ntotal, seg_rows, select_capacity(keep), kCandCap, keep, attempt);
    for (size_t i = 0; i < chunks.size() && (int)i < max_chunks && out; ++i) {
        out[3 * i] = chunks[i].seg; out[3 * i + 1] = chunks[i].row0; out[3 * i + 2] = chunks[i].row1;
    }
    return (int)chunks.size();
}

int drt_plan_params(int k, int attempt, int* kprime, int* cap_out) {
    if (k <= 0 || k > DRT_MAX_K) return fail(DRT_E_INVALID, "bad k");
    const int keep = kprime_for(k);
    if (kprime) *kprime = keep;
    if (cap_out) *cap_out = kCandCap;
    return DRT_OK;
}

int drt_search_stats(const drt_store* cs, int64_t out[12]) {
    if (!cs || !out) return fail(DRT_E_INVALID, "bad arguments");
    drt_store* s = const_cast<drt_store*>(cs);
    {
        std::lock_guard<std::mutex> lk(s->mu);
        if (s->async_pending) {      // counters of an asynchronous search: collect them now (waits for it)
            DeviceGuard g(s->device);
            CUDA_TRY(cudaEventSynchronize(s->async_done));
            for (size_t i = 0; i < s->async_timed; ++i) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, s->ev[2 * i], s->ev[2 * i + 1]) == cudaSuccess) s->stats[7] += (int64_t)(ms * 1e6);
            }
            if ((int)(s->misc_host[0] & 0xffffffff) != 0) s->stats[2] += 1;
            s->stats[9] += (int64_t)s->misc_host[1];
            s->stats[4] += (int64_t)s->misc_host[1];     // flagged and NOT refined here: the caller redoes the search
            s->stats[10] += (int64_t)s->misc_host[2];
            s->async_pending = false;
        }
    }
    for (int i = 0; i < 12; ++i) out[i] = s->stats[i];
    out[8] = s->exact_queries;
    return DRT_OK;
}

int drt_merge_topk(int n_lists, const float* scores, const int64_t* ids, int64_t nq, int k_in, int k_out,
                   float* out_scores, int64_t* out_ids, uint32_t flags, int device, void* stream) {
    if (n_lists <= 0 || nq < 0 || k_in <= 0 || k_out <= 0) return fail(DRT_E_INVALID, "bad merge shape");
    if (nq == 0) return DRT_OK;
    if (!scores || !ids || !out_scores || !out_ids) return fail(DRT_E_INVALID, "NULL pointer");
    const int P = next_pow2(n_lists * k_in);
    if (P > 8192) return fail(DRT_E_UNSUPPORTED, "merge of %d x %d entries exceeds 8192 per query; merge hierarchically", n_lists, k_in);
    if (k_out > P) return fail(DRT_E_INVALID, "k_out=%d exceeds the %d merged entries", k_out, P);
    int rc = check_device(device);
    if (rc != DRT_OK) return rc;
    DeviceGuard g(device);
    if (flags & DRT_MERGE_SORTED_UNIQUE) {
        CUDA_TRY(cudaFuncSetAttribute(drt::merge_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12));
        drt::merge_sorted_kernel<<<(int)nq, 256, (size_t)n_lists * k_in * 12, (cudaStream_t)stream>>>(
            n_lists, scores, (const long long*)ids, (long long)nq, k_in, k_out, out_scores, (long long*)out_ids);
        CUDA_TRY(cudaGetLastError());
        return DRT_OK;
    }
    CUDA_TRY(cudaFuncSetAttribute(drt::merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * (int)sizeof(drt::MergeEnt)));
    drt::merge_topk_kernel<<<(int)nq, 256, (size_t)P * sizeof(drt::MergeEnt), (cudaStream_t)stream>>>(
        n_lists, scores, (const long long*)ids, (long long)nq, k_in, k_out, out_scores, (long long*)out_ids);
    CUDA_TRY(cudaGetLastError());
    return DRT_OK;
}

int drt_merge_topk_peers(int n_lists, const float* const* scores, const int64_t* const* ids, int64_t q_begin,
                         int64_t q_count, int k_in, int k_out, float* const* out_scores, int64_t* const* out_ids,
                         uint8_t* const* truncated, int device, void* stream) {
    return drt_merge_topk_peers2(n_lists, scores, ids, q_begin, q_count, k_in, k_out, out_scores, out_ids, truncated, nullptr,
                                 nullptr, device, stream);
}

int drt_merge_topk_peers2(int n_lists, const float* const* scores, const int64_t* const* ids, int64_t q_begin,
                          int64_t q_count, int k_in, int k_out, float* const* out_scores, int64_t* const* out_ids,
                          uint8_t* const* truncated, const uint8_t* const* status, uint8_t* redo, int device, void* stream) {
    if (n_lists <= 0 || n_lists > 16 || q_begin < 0 || q_count < 0 || k_in <= 0 || k_out <= 0)
        return fail(DRT_E_INVALID, "bad peer-merge shape");
    if (!scores || !ids || !out_scores || !out_ids || !truncated) return fail(DRT_E_INVALID, "NULL pointer table");
    if ((int64_t)n_lists * k_in > 8192 || k_out > 4096) return fail(DRT_E_UNSUPPORTED, "peer merge of %d x %d entries is too large", n_lists, k_in);
    if (k_out > n_lists * k_in) return fail(DRT_E_INVALID, "k_out=%d exceeds the %d merged entries", k_out, n_lists * k_in);
    if (q_count == 0 && !status) return DRT_OK;
    int rc = check_device(device);
    if (rc != DRT_OK) return rc;
    DeviceGuard g(device);
    drt::PeerPtrs p;
    for (int i = 0; i < n_lists; ++i) {
        if (!scores[i] || !ids[i] || !truncated[i] || (!out_scores[i]) != (!out_ids[i])) return fail(DRT_E_INVALID, "NULL peer pointer %d", i);
        p.scores[i] = scores[i]; p.ids[i] = (const long long*)ids[i];
        p.out_scores[i] = out_scores[i]; p.out_ids[i] = (long long*)out_ids[i]; p.truncated[i] = truncated[i];
    }
    for (int i = n_lists; i < 16; ++i) { p.scores[i] = nullptr; p.ids[i] = nullptr; p.out_scores[i] = nullptr; p.out_ids[i] = nullptr; p.truncated[i] = nullptr; }
    for (int i = 0; i < 16; ++i) p.status[i] = (status && i < n_lists) ? status[i] : nullptr;
    p.redo = status ? redo : nullptr;
    const size_t smem = ((size_t)n_lists * k_in + (size_t)k_out) * 12;
    CUDA_TRY(cudaFuncSetAttribute(drt::merge_sorted_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (8192 + 4096) * 12));
    if (q_count == 0) {       // nothing to merge on this rank, but it still has to agree on `redo`
        drt::or_status_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, n_lists);
        CUDA_TRY(cudaGetLastError());
        return DRT_OK;
    }
    drt::merge_sorted_peers_kernel<<<(unsigned)q_count, 256, smem, (cudaStream_t)stream>>>(p, n_lists, (long long)q_begin, k_in, k_out);
    CUDA_TRY(cudaGetLastError());
    return DRT_OK;
}

// ---- in-batch CE ----------------------------------------------------------------------------
namespace {
struct CeWorkspace {
    DevBuf part_max, part_sum, tgt, ticket;
    DevBuf a_split, b_split, a2_split, b2_split, partials, partials2, logits;   // tensor-core paths
    bool ticket_init = false;
    int* err_host = nullptr;
    int* err_dev = nullptr;
    bool tc_attrs = false;
};
// One workspace per (device, stream): two losses in flight on different streams of one device
// (side streams, forward of step n+1 overlapping the backward of step n) never share tickets,
// partial buffers or split operands.  Work on ONE stream is ordered by the stream itself.
std::mutex g_ce_mu;
std::map<std::pair<int, void*>, CeWorkspace> g_ce_ws;
constexpr int kTicketSlots = 4096;       // [0] launch-wide ticket, [64 ..] per-tile tickets (two problems)
}  // namespace

extern "C++" {
namespace {
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cached per-device capability check: the CE entry points are latency sensitive
int check_device_cached(int device) {
    static int ok[64] = {0};
    if (device >= 0 && device < 64 && ok[device]) return DRT_OK;
    const int rc = check_device(device);
    if (rc == DRT_OK && device >= 0 && device < 64) ok[device] = 1;
    return rc;
}

template <class Cfg>
void launch_sgemm(const drt::GemmOperand& A, const drt::GemmOperand& Bm, long long M, long long N, long long K,
                  float* C, cudaStream_t st) {
    const int vecA = aligned16(A.p) && ((A.s_k == 1 ? A.s_outer : A.s_k) % 4 == 0);
    const int vecB = aligned16(Bm.p) && ((Bm.s_k == 1 ? Bm.s_outer : Bm.s_k) % 4 == 0);
    dim3 grid((unsigned)((N + Cfg::BN - 1) / Cfg::BN), (unsigned)((M + Cfg::BM - 1) / Cfg::BM));
    drt::sgemm_kernel<Cfg><<<grid, Cfg::THREADS, 0, st>>>(A, Bm, M, N, K, C, vecA, vecB);
}
// Tile tier for an M x N output: the largest tile that still gives every SM a CTA
// (0 = 32x32, 1 = 64x64, 2 = 128x128 with 8x8 register tiles for FMA-bound problems).
inline int gemm_tier(long long M, long long N, int sm_count) {
    if (((M + 127) / 128) * ((N + 127) / 128) >= 1ll * sm_count) return 2;
    if (((M + 63) / 64) * ((N + 63) / 64) >= 1ll * sm_count) return 1;
    return 0;
}
void launch_sgemm_tiered(const drt::GemmOperand& A, const drt::GemmOperand& Bm, long long M, long long N, long long K,
                         float* C, cudaStream_t st) {
    switch (gemm_tier(M, N, 148)) {
        case 2: launch_sgemm<drt::GemmHuge>(A, Bm, M, N, K, C, st); break;
        case 1: launch_sgemm<drt::GemmLarge>(A, Bm, M, N, K, C, st); break;
        default: launch_sgemm<drt::GemmSmall>(A, Bm, M, N, K, C, st); break;
    }
}
template <class C>
void launch_ce_fwd(const float* x, const float* y, long long B, long long P, int dim, const long long* target,
                   float loss_scale, float* logits_out, CeWorkspace& w, float* lse_out, float* loss_rows, float* loss_out,
                   int vec_ok, cudaStream_t st) {
    dim3 grid((unsigned)((P + C::BN - 1) / C::BN), (unsigned)((B + C::BM - 1) / C::BM));
    drt::inbatch_ce_fwd_kernel<C><<<grid, C::THREADS, 0, st>>>(
        x, y, B, P, dim, target, P / B, loss_scale, logits_out, (float*)w.part_max.p, (float*)w.part_sum.p,
        (float*)w.tgt.p, (unsigned int*)w.ticket.p, lse_out, loss_rows, loss_out, vec_ok);
}
template <class C>
void launch_ce_dlogits(const float* x, const float* y, long long B, long long P, int dim, const long long* target,
                       const float* lse, const float* grad_rows, int grad_stride, float grad_scale, float* work, int vec_ok,
                       cudaStream_t st) {
    dim3 grid((unsigned)((P + C::BN - 1) / C::BN), (unsigned)((B + C::BM - 1) / C::BM));
    drt::inbatch_ce_dlogits_kernel<C><<<grid, C::THREADS, 0, st>>>(x, y, B, P, dim, target, P / B, lse, grad_rows, grad_stride,
                                                                   grad_scale, work, vec_ok);
}

// ---- tensor-core (bf16x3 split) paths --------------------------------------------------------
__global__ void sum_partials_kernel(const float* __restrict__ part, int S, long long n, float* __restrict__ C) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int s = 0; s < S; ++s) a += part[(long long)s * n + i];
        C[i] = a;
    }
}

// Which path serves an [M, N] output contracted over K (all fp32-accurate):
//   kPathSimt   odd shapes (K not a multiple of 4: the split operand's row pitch must be 16-byte aligned)
//   kPathSmall  128 x 64 tiles + split-K, epilogue fused (gemm_tc_small.cuh): everything up to ~1e9 MACs
//   kPathBig    128/256 x 256 persistent tiles (gemm_tc.cuh): cross-device negatives and larger
enum { kPathSimt = 0, kPathSmall = 1, kPathBig = 2 };
inline int ce_path(long long M, long long N, long long K) {
    static const int forced = [] {
        if (getenv("DRT_B200_CE_SIMT")) return (int)kPathSimt;
        const char* e = getenv("DRT_B200_CE_PATH");        // tuning / test knob: simt | small | big
        if (!e) return -1;
        return !strcmp(e, "simt") ? (int)kPathSimt : !strcmp(e, "small") ? (int)kPathSmall : !strcmp(e, "big") ? (int)kPathBig : -1;
    }();
    if (K % 4 != 0 || K < 16) return kPathSimt;
    if (forced == kPathSimt) return kPathSimt;
    const bool big_ok = K % 32 == 0 && M >= 128 && N >= 256;
    if (forced == kPathBig && big_ok) return kPathBig;
    if (forced == kPathSmall) return kPathSmall;
    static const double big_min = [] { const char* e = getenv("DRT_B200_CE_BIG_MIN"); return e ? atof(e) : 1.0e9; }();
    return (big_ok && (double)M * (double)N * (double)K >= big_min) ? kPathBig : kPathSmall;
}

int ce_tc_setup(CeWorkspace& w, cudaStream_t st) {
    if (!w.tc_attrs) {
        CUDA_TRY(cudaFuncSetAttribute(drt::gemm_tc_nt_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)drt::GemmTcCfg<1>::kSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(drt::gemm_tc_nt_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)drt::GemmTcCfg<2>::kSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(drt::gemm_tc_small_kernel<drt::kStore>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)drt::kSmallSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(drt::gemm_tc_small_kernel<drt::kCe>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)drt::kSmallSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(drt::gemm_tc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)drt::kFusedSmemBytes));
        if (cudaHostAlloc((void**)&w.err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void**)&w.err_dev, w.err_host, 0) != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(DRT_E_CUDA, "pinned host allocation failed");
        }
        *w.err_host = 0;
        w.tc_attrs = true;
    }
    int rc;
    if ((rc = w.ticket.ensure(kTicketSlots * 4)) != DRT_OK) return rc;
    if (!w.ticket_init) { CUDA_TRY(cudaMemsetAsync(w.ticket.p, 0, kTicketSlots * 4, st)); w.ticket_init = true; }
    return DRT_OK;
}

// C[M,N] = A'[M,Kp] · B'[N,Kp]^T, A'/B' bf16 K-major (Kp = 6K, multiple of 64).  Split-K over
// `ksplit` chunks when the output has too few tiles to fill the GPU; partials are summed.
// `ce`: fused cross-entropy epilogue (no split-K; C may be NULL).
struct BigCe { float* part_max; float* part_sum; float* tgt; const long long* target; long long target_stride; };
int gemm_tc_nt(CeWorkspace& w, const void* a_split, const void* b_split, long long M, long long N, long long Kp,
               float* C, cudaStream_t st, const BigCe* ce = nullptr) {
    const int kctas = M > drt::kTileM ? 2 : 1;
    const int clusters_max = 148 / kctas;
    const int m_tiles = (int)((M + drt::kTileM * kctas - 1) / (drt::kTileM * kctas));
    const int n_tiles = (int)((N + drt::kTileN - 1) / drt::kTileN);
    const int nkb = (int)(Kp / drt::kBlockK);
    int ksplit = 1;
    if (!ce && m_tiles * n_tiles < clusters_max)
        ksplit = std::max(1, std::min(nkb / 16, (2 * clusters_max + m_tiles * n_tiles - 1) / (m_tiles * n_tiles)));
    CUtensorMap ta, tb;
    int rc;
    if ((rc = make_tmap_bf16(&ta, a_split, (uint64_t)M, (uint64_t)Kp, drt::kTileM)) != DRT_OK) return rc;
    if ((rc = make_tmap_bf16(&tb, b_split, (uint64_t)N, (uint64_t)Kp, drt::kTileN / kctas)) != DRT_OK) return rc;
    float* out = C;
    if (ksplit > 1) {
        if ((rc = w.partials.ensure((size_t)ksplit * M * N * 4)) != DRT_OK) return rc;
        out = (float*)w.partials.p;
    }
    {
        drt::GemmTcParams p;
        p.num_m_tiles = m_tiles; p.num_n_tiles = n_tiles; p.unit_tiles = 1;
        p.ksplit = ksplit; p.num_k_blocks = nkb;
        p.M = M; p.N = N; p.ldc = N; p.C = out; p.err = w.err_dev;
        p.ce_part_max = ce ? ce->part_max : nullptr; p.ce_part_sum = ce ? ce->part_sum : nullptr;
        p.ce_tgt_logit = ce ? ce->tgt : nullptr; p.ce_target = ce ? ce->target : nullptr;
        p.ce_target_stride = ce ? ce->target_stride : 0;
        const int total = m_tiles * n_tiles * ksplit;
        const int clusters = std::max(1, std::min(clusters_max, total));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(clusters * kctas);
        cfg.blockDim = dim3(drt::kFilterThreads);
        cfg.dynamicSmemBytes = kctas == 2 ? drt::GemmTcCfg<2>::kSmemBytes : drt::GemmTcCfg<1>::kSmemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (kctas == 2) CUDA_TRY(cudaLaunchKernelEx(&cfg, drt::gemm_tc_nt_kernel<2>, ta, tb, p));
        else CUDA_TRY(cudaLaunchKernelEx(&cfg, drt::gemm_tc_nt_kernel<1>, ta, tb, p));
    }
    if (ksplit > 1) {
        const long long n = M * N;
        sum_partials_kernel<<<(int)std::min<long long>((n + 255) / 256, 148 * 8), 256, 0, st>>>((const float*)w.partials.p, ksplit, n, C);
    }
    CUDA_TRY(cudaGetLastError());
    return DRT_OK;
}

inline int split_blocks(long long total) { return (int)std::min<long long>((total + 255) / 256, 148ll * 16); }

// ---- small-shape path: split jobs + gemm_tc_small_kernel --------------------------------------
drt::SplitJob make_job(const float* src, long long R, long long C, long long lds, void* dst, int is_b, int transpose,
                       int kind, int& tile_cursor) {
    drt::SplitJob j;
    j.src = src; j.R = R; j.C = C; j.lds = lds; j.dst = (__nv_bfloat16*)dst; j.is_b = is_b; j.transpose = transpose; j.kind = kind;
    j.vec = (!transpose && C % 8 == 0 && lds % 4 == 0 && aligned16(src) && aligned16(dst)) ? 1 : 0;
    j.tiles_x = (int)((C + (j.vec ? 63 : 31)) / (j.vec ? 64 : 32)); j.tiles_y = (int)((R + 31) / 32);
    j.tile_begin = tile_cursor;
    tile_cursor += j.tiles_x * j.tiles_y;
    return j;
}

// Describes C[M,N] = A'[M,6K]·B'[N,6K]^T for gemm_tc_small_kernel; `budget` = CTAs this problem may use.
int small_problem(drt::SmallProblem& pb, CUtensorMap* ta, CUtensorMap* tb, const void* a_split, const void* b_split,
                  long long M, long long N, long long K, float* C, DevBuf& partials, unsigned int* tickets, int budget) {
    const long long Kp = 6 * K;
    pb.m_tiles = (int)((M + drt::kTileM - 1) / drt::kTileM);
    pb.n_tiles = (int)((N + drt::kSmallTileN - 1) / drt::kSmallTileN);
    pb.num_k_blocks = (int)((Kp + drt::kBlockK - 1) / drt::kBlockK);
    const int tiles = pb.m_tiles * pb.n_tiles;
    // K-chunks per tile: fill the CTA budget, at least `min_kb` k-blocks (x 24 KB in flight) per chunk
    static const int max_split = [] { const char* e = getenv("DRT_B200_CE_KSPLIT_MAX"); return e ? std::max(1, atoi(e)) : 32; }();
    static const int min_kb = [] { const char* e = getenv("DRT_B200_CE_MIN_KB"); return e ? std::max(1, atoi(e)) : 4; }();
    pb.ksplit = std::max(1, std::min(std::min(budget / std::max(1, tiles), pb.num_k_blocks / min_kb), max_split));
    pb.M = M; pb.N = N; pb.C = C;
    pb.tile_ticket = tickets;
    pb.partials = nullptr;
    int rc;
    if (pb.ksplit > 1) {
        if ((rc = partials.ensure((size_t)tiles * pb.ksplit * drt::kTileM * drt::kSmallTileN * 4)) != DRT_OK) return rc;
        pb.partials = (float*)partials.p;
    }
    if ((rc = make_tmap_bf16(ta, a_split, (uint64_t)M, (uint64_t)Kp, drt::kTileM)) != DRT_OK) return rc;
    if ((rc = make_tmap_bf16(tb, b_split, (uint64_t)N, (uint64_t)Kp, drt::kSmallTileN)) != DRT_OK) return rc;
    return DRT_OK;
}
}  // namespace
}  // extern "C++"

int drt_inbatch_ce_fwd(const float* x, const float* y, int64_t B, int64_t P, int dim, const int64_t* target,
                       float loss_scale, float* logits_out, float* lse_out, float* loss_rows, float* loss_out,
                       int device, void* stream) {
    if (B <= 0 || P <= 0 || dim <= 0) return fail(DRT_E_INVALID, "bad shape B=%lld P=%lld d=%d", (long long)B, (long long)P, dim);
    if (!x || !y || !lse_out || !loss_rows || !loss_out) return fail(DRT_E_INVALID, "NULL pointer");
    if (!target && P / B == 0) return fail(DRT_E_INVALID, "default targets need P >= B");
    int rc = check_device_cached(device);
    if (rc != DRT_OK) return rc;
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lk(g_ce_mu);
    CeWorkspace& w = g_ce_ws[std::make_pair(device, stream)];
    const int path = ce_path(B, P, dim);
    const long long* tg = (const long long*)target;
    static const bool no_fuse = getenv("DRT_B200_CE_NOFUSE") != nullptr;
    if (path == kPathSmall && B <= drt::kTileM && dim % 64 == 0 && aligned16(x) && aligned16(y) && !no_fuse) {
        // ONE launch: the kernel reads the fp32 operands, splits them into bf16 pieces in shared
        // memory, contracts, reduces the K-chunks and finishes the cross entropy.  Only with a
        // single row tile (B <= 128, the reference's per-GPU batches): every y tile is then
        // converted exactly once; with more row tiles the conversion work would be repeated per
        // row tile and the separate preparation launch is cheaper (measured at 256 x 2048)
        if ((rc = ce_tc_setup(w, st)) != DRT_OK) return rc;
        drt::SmallParams sp = {};
        drt::SmallProblem& pb = sp.prob[0];
        pb.m_tiles = (int)((B + drt::kTileM - 1) / drt::kTileM);
        pb.n_tiles = (int)((P + drt::kSmallTileN - 1) / drt::kSmallTileN);
        pb.num_k_blocks = dim / 64;                    // SOURCE k-blocks: each becomes 6 piece products
        const int tiles = pb.m_tiles * pb.n_tiles;
        if (tiles > kTicketSlots - 64) return fail(DRT_E_UNSUPPORTED, "loss shape needs %d tiles on the small-tile path", tiles);
        static const int max_split = [] { const char* e = getenv("DRT_B200_CE_KSPLIT_MAX"); return e ? std::max(1, atoi(e)) : 32; }();
        // K-chunks of EQUAL length (a divisor of the k-block count): the tile's finisher waits for its
        // slowest chunk, so 9 chunks of 1-2 k-blocks were slower than 6 chunks of 2
        const int limit = std::max(1, std::min(std::min(148 / std::max(1, tiles), pb.num_k_blocks), max_split));
        pb.ksplit = 1;
        for (int dv = 1; dv <= limit; ++dv) if (pb.num_k_blocks % dv == 0) pb.ksplit = dv;
        pb.M = B; pb.N = P; pb.C = logits_out;
        unsigned int* tickets = (unsigned int*)w.ticket.p;
        pb.tile_ticket = tickets + 64;
        pb.partials = nullptr;
        if (pb.ksplit > 1) {
            if ((rc = w.partials.ensure((size_t)tiles * pb.ksplit * drt::kTileM * drt::kSmallTileN * 4)) != DRT_OK) return rc;
            pb.partials = (float*)w.partials.p;
        }
        if ((rc = w.part_max.ensure((size_t)B * pb.n_tiles * 4)) != DRT_OK) return rc;
        if ((rc = w.part_sum.ensure((size_t)B * pb.n_tiles * 4)) != DRT_OK) return rc;
        if ((rc = w.tgt.ensure((size_t)B * 4)) != DRT_OK) return rc;
        CUtensorMap tx, ty;
        if ((rc = make_tmap_f32(&tx, x, (uint64_t)B, (uint64_t)dim, drt::kTileM)) != DRT_OK) return rc;
        if ((rc = make_tmap_f32(&ty, y, (uint64_t)P, (uint64_t)dim, drt::kSmallTileN)) != DRT_OK) return rc;
        sp.prob[1] = sp.prob[0];
        sp.ctas0 = tiles * pb.ksplit;
        sp.err = w.err_dev;
        sp.ce.target = tg; sp.ce.target_stride = P / B; sp.ce.loss_scale = loss_scale;
        sp.ce.part_max = (float*)w.part_max.p; sp.ce.part_sum = (float*)w.part_sum.p; sp.ce.tgt_logit = (float*)w.tgt.p;
        sp.ce.ticket = tickets; sp.ce.lse_out = lse_out; sp.ce.loss_rows = loss_rows; sp.ce.loss_out = loss_out;
        if (const char* e = getenv("DRT_B200_CE_TRACE")) sp.dbg = (unsigned long long*)strtoull(e, nullptr, 0);   // dev: device buffer address
        drt::gemm_tc_fused_kernel<<<sp.ctas0, drt::kSmallThreads, drt::kFusedSmemBytes, st>>>(tx, ty, sp);
        CUDA_TRY(cudaGetLastError());
        return DRT_OK;
    }
    if (path == kPathSmall) {
        // one launch prepares both operands (exact bf16x3 split), one launch does the contraction,
        // the split-K reduction and the cross entropy; logits are stored only when asked for
        if ((rc = ce_tc_setup(w, st)) != DRT_OK) return rc;
        if ((rc = w.a_split.ensure((size_t)B * 6 * dim * 2)) != DRT_OK) return rc;
        if ((rc = w.b_split.ensure((size_t)P * 6 * dim * 2)) != DRT_OK) return rc;
        drt::SmallParams sp = {};
        CUtensorMap ta, tb;
        unsigned int* tickets = (unsigned int*)w.ticket.p;
        if ((rc = small_problem(sp.prob[0], &ta, &tb, w.a_split.p, w.b_split.p, B, P, dim, logits_out, w.partials, tickets + 64, 148)) != DRT_OK) return rc;
        const int tiles = sp.prob[0].m_tiles * sp.prob[0].n_tiles;
        if (tiles > kTicketSlots - 64) return fail(DRT_E_UNSUPPORTED, "loss shape needs %d tiles on the small-tile path", tiles);
        if ((rc = w.part_max.ensure((size_t)B * sp.prob[0].n_tiles * 4)) != DRT_OK) return rc;
        if ((rc = w.part_sum.ensure((size_t)B * sp.prob[0].n_tiles * 4)) != DRT_OK) return rc;
        if ((rc = w.tgt.ensure((size_t)B * 4)) != DRT_OK) return rc;
        drt::SplitJobs js = {};
        int cursor = 0;
        js.job[0] = make_job(x, B, dim, dim, w.a_split.p, 0, 0, 0, cursor);
        js.job[1] = make_job(y, P, dim, dim, w.b_split.p, 1, 0, 0, cursor);
        js.n_jobs = 2;
        drt::split3_jobs_kernel<<<cursor, 256, 0, st>>>(js);
        sp.prob[1] = sp.prob[0];
        sp.ctas0 = tiles * sp.prob[0].ksplit;
        sp.err = w.err_dev;
        sp.ce.target = tg; sp.ce.target_stride = P / B; sp.ce.loss_scale = loss_scale;
        sp.ce.part_max = (float*)w.part_max.p; sp.ce.part_sum = (float*)w.part_sum.p; sp.ce.tgt_logit = (float*)w.tgt.p;
        sp.ce.ticket = tickets; sp.ce.lse_out = lse_out; sp.ce.loss_rows = loss_rows; sp.ce.loss_out = loss_out;
        if (const char* e = getenv("DRT_B200_CE_TRACE")) sp.dbg = (unsigned long long*)strtoull(e, nullptr, 0);   // dev: device buffer address
        drt::gemm_tc_small_kernel<drt::kCe><<<sp.ctas0, drt::kSmallThreads, drt::kSmallSmemBytes, st>>>(ta, tb, ta, tb, sp);
        CUDA_TRY(cudaGetLastError());
        return DRT_OK;
    }
    if (path == kPathBig) {
        // large shape: 128/256 x 256 persistent tiles; every epilogue thread keeps the online
        // (max, sum-exp) of its row over its 128 columns, so no pass over the logits follows
        if ((rc = ce_tc_setup(w, st)) != DRT_OK) return rc;
        if ((rc = w.a_split.ensure((size_t)B * 6 * dim * 2)) != DRT_OK) return rc;
        if ((rc = w.b_split.ensure((size_t)P * 6 * dim * 2)) != DRT_OK) return rc;
        const int ncol = 2 * (int)((P + drt::kTileN - 1) / drt::kTileN);
        if ((rc = w.part_max.ensure((size_t)B * ncol * 4)) != DRT_OK) return rc;
        if ((rc = w.part_sum.ensure((size_t)B * ncol * 4)) != DRT_OK) return rc;
        if ((rc = w.tgt.ensure((size_t)B * 4)) != DRT_OK) return rc;
        drt::SplitJobs js = {};
        int cursor = 0;
        js.job[0] = make_job(x, B, dim, dim, w.a_split.p, 0, 0, 0, cursor);
        js.job[1] = make_job(y, P, dim, dim, w.b_split.p, 1, 0, 0, cursor);
        js.n_jobs = 2;
        drt::split3_jobs_kernel<<<cursor, 256, 0, st>>>(js);
        BigCe ce{(float*)w.part_max.p, (float*)w.part_sum.p, (float*)w.tgt.p, tg, (long long)(P / B)};
        if ((rc = gemm_tc_nt(w, w.a_split.p, w.b_split.p, B, P, 6ll * dim, logits_out, st, &ce)) != DRT_OK) return rc;
        drt::ce_fold_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>((const float*)w.part_max.p, (const float*)w.part_sum.p,
                                                                     (const float*)w.tgt.p, B, P, ncol, tg, (long long)(P / B), loss_scale,
                                                                     lse_out, loss_rows, loss_out, (unsigned int*)w.ticket.p);
        CUDA_TRY(cudaGetLastError());
        return DRT_OK;
    }
    const int tier = gemm_tier(B, P, 148);
    const int bn = tier == 2 ? drt::GemmHuge::BN : tier == 1 ? drt::GemmLarge::BN : drt::GemmSmall::BN;
    const int ncol = (int)((P + bn - 1) / bn);
    if ((rc = w.part_max.ensure((size_t)B * ncol * 4)) != DRT_OK) return rc;
    if ((rc = w.part_sum.ensure((size_t)B * ncol * 4)) != DRT_OK) return rc;
    if ((rc = w.tgt.ensure((size_t)B * 4)) != DRT_OK) return rc;
    if ((rc = w.ticket.ensure(kTicketSlots * 4)) != DRT_OK) return rc;
    if (!w.ticket_init) { CUDA_TRY(cudaMemsetAsync(w.ticket.p, 0, kTicketSlots * 4, st)); w.ticket_init = true; }
    const int vec_ok = aligned16(x) && aligned16(y) && dim % 4 == 0;
    if (tier == 2) launch_ce_fwd<drt::GemmHuge>(x, y, B, P, dim, tg, loss_scale, logits_out, w, lse_out, loss_rows, loss_out, vec_ok, st);
    else if (tier == 1) launch_ce_fwd<drt::GemmLarge>(x, y, B, P, dim, tg, loss_scale, logits_out, w, lse_out, loss_rows, loss_out, vec_ok, st);
    else launch_ce_fwd<drt::GemmSmall>(x, y, B, P, dim, tg, loss_scale, logits_out, w, lse_out, loss_rows, loss_out, vec_ok, st);
    CUDA_TRY(cudaGetLastError());
    return DRT_OK;
}

int drt_inbatch_ce_bwd_needs_work(int64_t B, int64_t P, int dim, int have_logits) {
    if (B <= 0 || P <= 0 || dim <= 0) return 1;
    const bool small = have_logits && ce_path(B, P, dim) == kPathSmall && ce_path(B, dim, P) != kPathSimt &&
                       ce_path(P, dim, B) != kPathSimt;
    return small ? 0 : 1;
}

int drt_inbatch_ce_bwd(const float* x, const float* y, int64_t B, int64_t P, int dim, const int64_t* target,
                       const float* lse, const float* logits, const float* grad_rows, int grad_stride, float grad_scale,
                       float* work, float* dx, float* dy, int device, void* stream) {
    if (B <= 0 || P <= 0 || dim <= 0) return fail(DRT_E_INVALID, "bad shape");
    if (!x || !y || !lse || !grad_rows) return fail(DRT_E_INVALID, "NULL pointer");
    if (grad_stride != 0 && grad_stride != 1) return fail(DRT_E_INVALID, "grad_stride must be 0 (scalar) or 1 (per row)");
    int rc = check_device_cached(device);
    if (rc != DRT_OK) return rc;
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int vec_ok = aligned16(x) && aligned16(y) && dim % 4 == 0;
    const long long* tg = (const long long*)target;
    // dx[B,d] contracts over P, dy[P,d] over B: both must be splittable for a tensor-core path
    const int pdx = ce_path(B, dim, P), pdy = ce_path(P, dim, B), pfw = ce_path(B, P, dim);
    const bool small = logits && pfw == kPathSmall && pdx != kPathSimt && pdy != kPathSimt;
    if (small) {
        // Launch 1: dlogits from the kept logits, written directly as the split bf16 operands of
        // both contractions (row-major for dx, transposed for dy), plus the splits of y^T and x^T.
        // Launch 2: both GEMMs (two problems, one grid).  Nothing is written to `work` or `logits`.
        std::lock_guard<std::mutex> lk(g_ce_mu);
        CeWorkspace& w = g_ce_ws[std::make_pair(device, stream)];
        if ((rc = ce_tc_setup(w, st)) != DRT_OK) return rc;
        drt::SplitJobs js = {};
        int cursor = 0, nj = 0;
        if (dx) {
            if ((rc = w.a_split.ensure((size_t)B * 6 * P * 2)) != DRT_OK) return rc;       // dL'  [B, 6P]
            if ((rc = w.b_split.ensure((size_t)dim * 6 * P * 2)) != DRT_OK) return rc;     // yT'  [d, 6P]
            js.job[nj++] = make_job(logits, B, P, P, w.a_split.p, 0, 0, 1, cursor);
            js.job[nj++] = make_job(y, P, dim, dim, w.b_split.p, 1, 1, 0, cursor);
        }
        if (dy) {
            if ((rc = w.a2_split.ensure((size_t)P * 6 * B * 2)) != DRT_OK) return rc;      // dLT' [P, 6B]
            if ((rc = w.b2_split.ensure((size_t)dim * 6 * B * 2)) != DRT_OK) return rc;    // xT'  [d, 6B]
            js.job[nj++] = make_job(logits, B, P, P, w.a2_split.p, 0, 1, 1, cursor);
            js.job[nj++] = make_job(x, B, dim, dim, w.b2_split.p, 1, 1, 0, cursor);
        }
        if (nj == 0) return DRT_OK;
        js.n_jobs = nj;
        js.lse = lse; js.target = tg; js.target_stride = P / B; js.grad_rows = grad_rows; js.grad_stride = grad_stride; js.grad_scale = grad_scale;
        drt::split3_jobs_kernel<<<cursor, 256, 0, st>>>(js);
        drt::SmallParams sp = {};
        CUtensorMap t[4];
        unsigned int* tickets = (unsigned int*)w.ticket.p;
        // CTA budget: dy (M = P rows) has many tiles and a short K, dx few tiles and a long K
        const int dy_tiles = dy ? (int)(((P + 127) / 128) * ((dim + 63) / 64)) : 0;
        const int dx_tiles = dx ? (int)(((B + 127) / 128) * ((dim + 63) / 64)) : 0;
        if (dx_tiles + dy_tiles > kTicketSlots - 64) return fail(DRT_E_UNSUPPORTED, "loss shape needs too many tiles on the small-tile path");
        int np = 0, ctas[2] = {0, 0};
        if (dx) {
            const int budget = std::max(dx_tiles, 148 - std::min(dy_tiles, 100));
            if ((rc = small_problem(sp.prob[np], &t[2 * np], &t[2 * np + 1], w.a_split.p, w.b_split.p, B, dim, P, dx, w.partials, tickets + 64, budget)) != DRT_OK) return rc;
            ctas[np] = dx_tiles * sp.prob[np].ksplit; ++np;
        }
        if (dy) {
            const int budget = std::max(dy_tiles, 148 - ctas[0]);
            if ((rc = small_problem(sp.prob[np], &t[2 * np], &t[2 * np + 1], w.a2_split.p, w.b2_split.p, P, dim, B, dy, w.partials2, tickets + 64 + dx_tiles, budget)) != DRT_OK) return rc;
            ctas[np] = dy_tiles * sp.prob[np].ksplit; ++np;
        }
        if (np == 1) { sp.prob[1] = sp.prob[0]; t[2] = t[0]; t[3] = t[1]; }
        sp.ctas0 = ctas[0];
        sp.err = w.err_dev;
        drt::gemm_tc_small_kernel<drt::kStore><<<ctas[0] + ctas[1], drt::kSmallThreads, drt::kSmallSmemBytes, st>>>(t[0], t[1], t[2], t[3], sp);
        CUDA_TRY(cudaGetLastError());
        return DRT_OK;
    }
    if (!work) return fail(DRT_E_INVALID, "this shape needs the B*P work buffer");
    const int tier = gemm_tier(B, P, 148);
    if (logits) {    // forward kept the logits: dlogits is one elementwise pass into `work`
        const long long total = (long long)B * P;
        const int blocks = (int)std::min<long long>((total + 255) / 256, 148ll * 16);
        drt::ce_dlogits_from_logits_kernel<<<blocks, 256, 0, st>>>(logits, (long long)B, (long long)P, tg, (long long)(P / B), lse,
                                                                  grad_rows, grad_stride, grad_scale, work);
    } else if (tier == 2) launch_ce_dlogits<drt::GemmHuge>(x, y, B, P, dim, tg, lse, grad_rows, grad_stride, grad_scale, work, vec_ok, st);
    else if (tier == 1) launch_ce_dlogits<drt::GemmLarge>(x, y, B, P, dim, tg, lse, grad_rows, grad_stride, grad_scale, work, vec_ok, st);
    else launch_ce_dlogits<drt::GemmSmall>(x, y, B, P, dim, tg, lse, grad_rows, grad_stride, grad_scale, work, vec_ok, st);
    const bool tc = pdx == kPathBig && pdy == kPathBig && dim >= 256;
    if (tc) {
        std::lock_guard<std::mutex> lk(g_ce_mu);
        CeWorkspace& w = g_ce_ws[std::make_pair(device, stream)];
        if ((rc = ce_tc_setup(w, st)) != DRT_OK) return rc;
        if ((rc = w.a_split.ensure((size_t)B * 6 * P * 2)) != DRT_OK) return rc;      // dL' / dLT' : B*6P == P*6B elements
        if ((rc = w.b_split.ensure((size_t)dim * 6 * std::max(B, P) * 2)) != DRT_OK) return rc;
        dim3 tb(32, 8);
        if (dx) {   // dx[B,d] = dL[B,P] · (yT)[d,P]^T
            drt::split3_rows_kernel<<<split_blocks(B * P), 256, 0, st>>>(work, B, P, P, (__nv_bfloat16*)w.a_split.p, 0);
            drt::split3_transpose_kernel<<<dim3((unsigned)((dim + 31) / 32), (unsigned)((P + 31) / 32)), tb, 0, st>>>(
                y, P, dim, (__nv_bfloat16*)w.b_split.p, 1);
            if ((rc = gemm_tc_nt(w, w.a_split.p, w.b_split.p, B, dim, 6ll * P, dx, st)) != DRT_OK) return rc;
        }
        if (dy) {   // dy[P,d] = (dLT)[P,B] · (xT)[d,B]^T
            drt::split3_transpose_kernel<<<dim3((unsigned)((P + 31) / 32), (unsigned)((B + 31) / 32)), tb, 0, st>>>(
                work, B, P, (__nv_bfloat16*)w.a_split.p, 0);
            drt::split3_transpose_kernel<<<dim3((unsigned)((dim + 31) / 32), (unsigned)((B + 31) / 32)), tb, 0, st>>>(
                x, B, dim, (__nv_bfloat16*)w.b_split.p, 1);
            if ((rc = gemm_tc_nt(w, w.a_split.p, w.b_split.p, P, dim, 6ll * B, dy, st)) != DRT_OK) return rc;
        }
        CUDA_TRY(cudaGetLastError());
        return DRT_OK;
    }
    // dx[B,d] = dlogits[B,P] · y[P,d]   (A k-contiguous, B n-contiguous)
    if (dx) launch_sgemm_tiered(drt::GemmOperand{work, (long long)P, 1}, drt::GemmOperand{y, 1, (long long)dim}, B, dim, P, dx, st);
    // dy[P,d] = dlogitsᵀ[P,B] · x[B,d]  (A m-contiguous, B n-contiguous)
    if (dy) launch_sgemm_tiered(drt::GemmOperand{work, 1, (long long)P}, drt::GemmOperand{x, 1, (long long)dim}, P, dim, B, dy, st);
    CUDA_TRY(cudaGetLastError());
    return DRT_OK;
}

int drt_filter_negatives(const int64_t* ids, int64_t nq, int k, const int64_t* pos_begin, const int64_t* pos_end,
                         int num_negative, int64_t* out_ids, int device, void* stream) {
    if (nq < 0 || k <= 0 || num_negative <= 0) return fail(DRT_E_INVALID, "bad shape");
    if (nq == 0) return DRT_OK;
    if (!ids || !pos_begin || !pos_end || !out_ids) return fail(DRT_E_INVALID, "NULL pointer");
    int rc = check_device(device);
    if (rc != DRT_OK) return rc;
    DeviceGuard g(device);
    const int64_t threads = nq * 32;
    drt::filter_negatives_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)ids, (long long)nq, k, (const long long*)pos_begin, (const long long*)pos_end, num_negative,
        (long long*)out_ids);
    CUDA_TRY(cudaGetLastError());
    return DRT_OK;
}

}  // extern "C"
