/*
 * oracle/flat_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, scalar) of the exhaustive flat search the reference
 * performs through faiss-cpu==1.7.4 (requirements.txt:9) at
 *   src/retrieval.py:102            faiss_index.search(query_embedding, top_k)
 *   src/create_embeddings.py:291    index.search(test_vector, 1)
 *   scripts/phase3_pdf_chunking.py:430,441
 * faiss itself is a third-party wheel that is NOT vendored under /root/reference and
 * is not installable here (no network), so this file restates the published
 * algorithm of faiss 1.7.4 IndexFlat::search:
 *   - METRIC_L2 (IndexFlatL2, fourcc IxF2): squared L2; per query a max-heap of
 *     size k; a candidate is admitted only if dis < heap top (strict); the heap is
 *     then re-ordered ascending; unfilled slots keep id -1 / distance FLT_MAX.
 *       nq <  20  -> per-pair sum_i (q_i - x_i)^2            ("direct" form)
 *       nq >= 20  -> ||q||^2 + ||x||^2 - 2 q.x, clamped at 0 ("expanded" form)
 *   - METRIC_INNER_PRODUCT (IndexFlatIP, fourcc IxFI): min-heap, admitted if
 *     ip > heap top (strict), re-ordered descending; unfilled: id -1 / -FLT_MAX.
 *   Rows are visited in ascending id order, so at equal value the lower id is
 *   kept.  The canonical total order used everywhere in this repo is therefore
 *     L2: (distance asc, id asc)      IP: (score desc, id asc).
 *
 * PARITY UNPINNED: the reference ships no tests or golden vectors for this call
 * (SURVEY.md section 8c).  The restatement is checked against the structural
 * invariants the reference's recorded outputs do pin (tests/test_oracle.py):
 * ascending distances, similarity == 1/(1+distance), self-search distance 0 on
 * all 14 shipped indices, and numpy float64 brute force.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm
 * may call this.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_METRIC_IP 0
#define ORACLE_METRIC_L2 1

/* ---- scalar kernels (faiss fvec_L2sqr_ref / fvec_inner_product_ref) ---- */
static float l2sqr(const float* a, const float* b, int d) {
    float s = 0.f;
    for (int i = 0; i < d; i++) { float t = a[i] - b[i]; s += t * t; }
    return s;
}
static float inner(const float* a, const float* b, int d) {
    float s = 0.f;
    for (int i = 0; i < d; i++) s += a[i] * b[i];
    return s;
}

/* "a is worse than b" under the canonical order; worse elements sit at the heap top */
static int worse_l2(float va, int64_t ia, float vb, int64_t ib) { return va > vb || (va == vb && ia > ib); }
static int worse_ip(float va, int64_t ia, float vb, int64_t ib) { return va < vb || (va == vb && ia > ib); }

typedef int (*worse_fn)(float, int64_t, float, int64_t);

static void sift_down(float* v, int64_t* id, int k, int i, worse_fn worse) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < k && worse(v[l], id[l], v[m], id[m])) m = l;
        if (r < k && worse(v[r], id[r], v[m], id[m])) m = r;
        if (m == i) return;
        float tv = v[i]; v[i] = v[m]; v[m] = tv;
        int64_t ti = id[i]; id[i] = id[m]; id[m] = ti;
        i = m;
    }
}

/* heap of the k best so far, worst at index 0 (faiss heap_replace_top + heap_reorder) */
static void heap_finish(float* v, int64_t* id, int k, int filled, worse_fn worse, float sentinel) {
    /* pop worst repeatedly to the back -> best first */
    int n = filled;
    /* compact: unfilled slots are sentinels which are "worst"; they pop first */
    for (int end = k; end > 1; end--) {
        float tv = v[0]; v[0] = v[end - 1]; v[end - 1] = tv;
        int64_t ti = id[0]; id[0] = id[end - 1]; id[end - 1] = ti;
        sift_down(v, id, end - 1, 0, worse);
    }
    (void)n; (void)sentinel;
}

/*
 * x: [n, d] row-major fp32 corpus; q: [nq, d] fp32 queries.
 * form: 0 = what faiss would pick (direct if nq < 20 else expanded), 1 = direct, 2 = expanded.
 * D: [nq, k] float32, I: [nq, k] int64.  Returns 0.
 */
int oracle_flat_search(const float* x, int64_t n, int d, const float* q, int64_t nq, int k,
                       int metric, int form, float* D, int64_t* I) {
    if (k <= 0 || d <= 0) return -1;
    int expanded = (form == 2) || (form == 0 && nq >= 20);
    float* xn = NULL;
    if (metric == ORACLE_METRIC_L2 && expanded) {
        xn = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        for (int64_t j = 0; j < n; j++) xn[j] = inner(x + j * d, x + j * d, d);
    }
    worse_fn worse = metric == ORACLE_METRIC_L2 ? worse_l2 : worse_ip;
    float sentinel = metric == ORACLE_METRIC_L2 ? FLT_MAX : -FLT_MAX;
    for (int64_t i = 0; i < nq; i++) {
        const float* qi = q + i * d;
        float* v = D + i * k;
        int64_t* id = I + i * k;
        for (int s = 0; s < k; s++) { v[s] = sentinel; id[s] = -1; }
        /* ids of sentinels: -1 compares as "better id" than any real id at equal value, but a
           real value is never equal to +-FLT_MAX for finite inputs */
        float qn = (xn != NULL) ? inner(qi, qi, d) : 0.f;
        int filled = 0;
        for (int64_t j = 0; j < n; j++) {
            float val;
            if (metric == ORACLE_METRIC_L2) {
                if (expanded) {
                    float ip = inner(qi, x + j * d, d);
                    val = qn + xn[j] - 2.f * ip;
                    if (val < 0.f) val = 0.f;
                } else {
                    val = l2sqr(qi, x + j * d, d);
                }
                if (!(val < v[0])) continue;          /* strict admission */
            } else {
                val = inner(qi, x + j * d, d);
                if (!(val > v[0])) continue;
            }
            v[0] = val; id[0] = j;
            sift_down(v, id, k, 0, worse);
            if (filled < k) filled++;
        }
        heap_finish(v, id, k, filled, worse, sentinel);
    }
    free(xn);
    return 0;
}

/* Strict total-order reference: sort ALL n candidates by the canonical order and keep k.
 * O(n log n) per query; used to cross-check the heap restatement on small inputs. */
typedef struct { float v; int64_t id; } pair_t;
static int cmp_l2(const void* a, const void* b) {
    const pair_t* x = (const pair_t*)a; const pair_t* y = (const pair_t*)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}
static int cmp_ip(const void* a, const void* b) {
    const pair_t* x = (const pair_t*)a; const pair_t* y = (const pair_t*)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}
int oracle_flat_search_sort(const float* x, int64_t n, int d, const float* q, int64_t nq, int k,
                            int metric, float* D, int64_t* I) {
    pair_t* p = (pair_t*)malloc(sizeof(pair_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < nq; i++) {
        for (int64_t j = 0; j < n; j++) {
            p[j].id = j;
            p[j].v = metric == ORACLE_METRIC_L2 ? l2sqr(q + i * d, x + j * d, d) : inner(q + i * d, x + j * d, d);
        }
        qsort(p, (size_t)n, sizeof(pair_t), metric == ORACLE_METRIC_L2 ? cmp_l2 : cmp_ip);
        for (int s = 0; s < k; s++) {
            if (s < n) { D[i * k + s] = p[s].v; I[i * k + s] = p[s].id; }
            else { D[i * k + s] = metric == ORACLE_METRIC_L2 ? FLT_MAX : -FLT_MAX; I[i * k + s] = -1; }
        }
    }
    free(p);
    return 0;
}
