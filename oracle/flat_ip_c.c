/*
 * flat_ip_c.c — plain-C CPU restatement of faiss.IndexFlatIP.search.  TEST INFRASTRUCTURE ONLY:
 * linked/loaded by tests/, __graft_entry__.smoke() and bench.py's baseline legs, never by the
 * product library.
 *
 * Follows the call the reference makes at DRT/evaluator/index.py:32 (`self.index.search(q, k)`
 * on a `faiss.IndexFlatIP`, index.py:19).  faiss is an un-vendored, unpinned dependency of the
 * reference; its published algorithm for IndexFlatIP is: for every query, inner product with
 * every stored row in fp32, k-selection through a min-heap whose root is the current threshold,
 * admission test `threshold < score`, heap initialised to (-FLT_MAX, -1), final ordering by
 * decreasing score.  This file implements exactly that with scalar fp32 accumulation (no BLAS),
 * breaking ties by ascending id so the output is canonical: (score desc, id asc).
 *
 * Independent of oracle/flat_ip.py (numpy sgemm + partition); tests check the two agree.
 * Parity against the faiss binary itself is unpinned (faiss cannot be installed here).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { float s; int64_t id; } ent_t;

/* "a ranks worse than b": lower score, or equal score and larger id */
static inline int worse(const ent_t a, const ent_t b) {
    return (a.s < b.s) || (a.s == b.s && a.id > b.id);
}

static void sift_down(ent_t* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && worse(h[l], h[m])) m = l;
        if (r < n && worse(h[r], h[m])) m = r;
        if (m == i) return;
        ent_t t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}

static int cmp_desc(const void* pa, const void* pb) {
    const ent_t* a = (const ent_t*)pa; const ent_t* b = (const ent_t*)pb;
    if (worse(*b, *a)) return -1;
    if (worse(*a, *b)) return 1;
    return 0;
}

/* corpus [n,d], q [nq,d] row-major fp32; out_d [nq,k], out_i [nq,k].  Returns 0. */
int oracle_flat_ip_search(const float* corpus, int64_t n, int d, const float* q, int64_t nq,
                          int k, float* out_d, int64_t* out_i) {
    if (k <= 0 || d <= 0) return -1;
    /* single-threaded: this image's gcc has no libgomp */
    for (int64_t qi = 0; qi < nq; ++qi) {
        ent_t* heap = (ent_t*)malloc(sizeof(ent_t) * (size_t)k);
        int filled = 0;
        const float* qv = q + qi * d;
        for (int64_t r = 0; r < n; ++r) {
            const float* xv = corpus + r * d;
            float acc = 0.f;
            for (int j = 0; j < d; ++j) acc += qv[j] * xv[j];
            if (!(acc > -FLT_MAX)) continue;           /* NaN / -inf / -FLT_MAX never enter */
            ent_t e = { acc, r };
            if (filled < k) {                          /* heap not full: threshold is -FLT_MAX */
                heap[filled++] = e;
                if (filled == k) for (int i = k / 2 - 1; i >= 0; --i) sift_down(heap, k, i);
            } else if (heap[0].s < acc) {              /* faiss: C::cmp(threshold, score)      */
                heap[0] = e;                           /* rows arrive in id order, so a tie at */
                sift_down(heap, k, 0);                 /* the threshold keeps the lower id     */
            }
        }
        qsort(heap, (size_t)filled, sizeof(ent_t), cmp_desc);
        for (int j = 0; j < k; ++j) {
            out_d[qi * k + j] = j < filled ? heap[j].s : -FLT_MAX;
            out_i[qi * k + j] = j < filled ? heap[j].id : -1;
        }
        free(heap);
    }
    return 0;
}
