/*
 * vm_oracle.c -- CPU restatement of the reference's embedding-similarity path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library, and only as
 * the checker (or the timed CPU baseline).  The product path (libvidmem.so) never links,
 * loads or calls anything in oracle/.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout).  Parity pinning: the reference ships no tests and no golden vectors
 * for this path, so this file is pinned against outputs of the
 * reference's own unmodified functions executed in the authoring container
 * (oracle/gen_golden.py -> tests/golden/ npz files; tests/test_oracle_golden.py).
 * The Neo4j `vector.similarity.cosine` restatement (vo_vector_search) could not be run
 * against a Neo4j server -> that one function is "parity unpinned" (SURVEY.md 9.3).
 *
 * Arithmetic notes
 *  - The reference computes on Python floats (IEEE binary64).  `sum()` over a generator
 *    of floats is a plain left-to-right sum on CPython < 3.12 and a Neumaier-compensated
 *    sum on CPython >= 3.12 (Python/bltinmodule.c, builtin_sum float fast path).  Both are
 *    restated (sum_mode); the authoring container runs 3.12.3, so goldens use Neumaier.
 *  - Build with -O2 -ffp-contract=off and without -ffast-math: no FMA contraction, no
 *    reassociation, so every operation rounds exactly like CPython's C doubles.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VO_SUM_NAIVE 0    /* CPython <  3.12 */
#define VO_SUM_NEUMAIER 1 /* CPython >= 3.12 */

/* ---- builtin sum() over floats, start=0 -------------------------------------------- */
/* CPython: result starts as int 0; the first float item leaves the int fast path through
 * PyNumber_Add(0, x0) == x0 (as a double: 0.0 + x0), the rest run in the float fast path. */
typedef struct { double s, c; int n; int mode; } vo_acc;

static inline void acc_init(vo_acc *a, int mode) { a->s = 0.0; a->c = 0.0; a->n = 0; a->mode = mode; }

static inline void acc_add(vo_acc *a, double x)
{
    if (a->n == 0) { a->s = 0.0 + x; a->n = 1; return; }
    if (a->mode == VO_SUM_NAIVE) { a->s = a->s + x; return; }
    double t = a->s + x;
    if (fabs(a->s) >= fabs(x)) a->c += (a->s - t) + x;
    else a->c += (x - t) + a->s;
    a->s = t;
}

static inline double acc_result(const vo_acc *a)
{
    double r = a->s;
    if (a->mode == VO_SUM_NEUMAIER && a->n > 0 && a->c != 0.0 && isfinite(a->c)) r += a->c;
    return r;
}

static double dot_seq(const double *a, const double *b, int64_t n, int mode)
{
    vo_acc acc; acc_init(&acc, mode);
    for (int64_t i = 0; i < n; ++i) acc_add(&acc, a[i] * b[i]);
    return acc_result(&acc);
}

/* ---- scalar cosine variants ---------------------------------------- */

/* PreLLMInjector._cosine_similarity: src/components/pre_llm_injector.py:374-388
 * length mismatch -> 0.0; norm1 == 0 or norm2 == 0 -> 0.0; dot / (norm1 * norm2). */
double vo_cosine_injector(const double *v1, int64_t n1, const double *v2, int64_t n2, int sum_mode)
{
    if (n1 != n2) return 0.0;
    double dot = dot_seq(v1, v2, n1, sum_mode);
    double norm1 = sqrt(dot_seq(v1, v1, n1, sum_mode));
    double norm2 = sqrt(dot_seq(v2, v2, n2, sum_mode));
    if (norm1 == 0 || norm2 == 0) return 0.0;
    return dot / (norm1 * norm2);
}

/* HybridRetriever._cosine_similarity: src/pipeline/retriever_hybrid.py:655-664
 * zip() truncates the dot to the shorter vector, the magnitudes use the full vectors;
 * mag1 * mag2 == 0 -> 0.0 (also fires when the product underflows). */
double vo_cosine_retriever(const double *v1, int64_t n1, const double *v2, int64_t n2, int sum_mode)
{
    int64_t n = n1 < n2 ? n1 : n2;
    double dot = dot_seq(v1, v2, n, sum_mode);
    double mag1 = sqrt(dot_seq(v1, v1, n1, sum_mode));
    double mag2 = sqrt(dot_seq(v2, v2, n2, sum_mode));
    if (mag1 * mag2 == 0) return 0.0;
    return dot / (mag1 * mag2);
}

/* EmbeddingUtils.cosine_similarity: src/utils/embedding_utils.py:29-39
 * magnitudes via `** 0.5` (libm pow), either magnitude == 0 -> 0.0. */
double vo_cosine_utils(const double *v1, int64_t n1, const double *v2, int64_t n2, int sum_mode)
{
    int64_t n = n1 < n2 ? n1 : n2;
    double dot = dot_seq(v1, v2, n, sum_mode);
    double m1 = pow(dot_seq(v1, v1, n1, sum_mode), 0.5);
    double m2 = pow(dot_seq(v2, v2, n2, sum_mode), 0.5);
    if (m1 == 0 || m2 == 0) return 0.0;
    return dot / (m1 * m2);
}

/* ---- stable descending sort helper -------------------------------------------------- */
/* list.sort(key=score, reverse=True) is stable: equal scores keep original order
 * (src/components/pre_llm_injector.py:369). */
typedef struct { double score; int64_t idx; } vo_pair;

static int cmp_desc_stable(const void *pa, const void *pb)
{
    const vo_pair *a = (const vo_pair *)pa, *b = (const vo_pair *)pb;
    if (a->score > b->score) return -1;
    if (a->score < b->score) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx);
}

/* PreLLMInjector._calculate_batch_similarities: src/components/pre_llm_injector.py:346-372
 *  queries  [q][d]   row-major doubles; query_ok[q]==0 models an Exception-valued embedding
 *                    (`isinstance(chunk_emb, Exception)` -> [] , :357-359)
 *  store    [n][d]   row-major doubles in store (dict insertion) order; row_ok[i]==0 models a
 *                    falsy embedding that is skipped (`if existing_emb:` , :363)
 *  top_k             embedder_config.top_k_chunk_with_batch_similarity (:370)
 *  out_idx/out_score [q][top_k], out_count[q] = number of valid entries (no padding, :370)
 * Every (query,row) score uses vo_cosine_injector on equal lengths. */
int vo_batch_similarities(const double *queries, const uint8_t *query_ok, int64_t q, const double *store,
                          const uint8_t *row_ok, int64_t n, int64_t d, int64_t top_k, int sum_mode,
                          int64_t *out_idx, double *out_score, int64_t *out_count)
{
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < q; ++qi) {
        out_count[qi] = 0;
        if (query_ok && !query_ok[qi]) continue;
        vo_pair *all = (vo_pair *)malloc(sizeof(vo_pair) * (size_t)(n > 0 ? n : 1));
        if (!all) { rc = -1; continue; }
        const double *qv = queries + qi * d;
        double qq = dot_seq(qv, qv, d, sum_mode);
        double norm1 = sqrt(qq);
        int64_t m = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (row_ok && !row_ok[i]) continue;
            const double *rv = store + i * d;
            double dot = dot_seq(qv, rv, d, sum_mode);
            double norm2 = sqrt(dot_seq(rv, rv, d, sum_mode));
            double s = (norm1 == 0 || norm2 == 0) ? 0.0 : dot / (norm1 * norm2);
            all[m].score = s; all[m].idx = i; ++m;
        }
        qsort(all, (size_t)m, sizeof(vo_pair), cmp_desc_stable);
        int64_t keep = m < top_k ? m : top_k;
        for (int64_t j = 0; j < keep; ++j) {
            out_idx[qi * top_k + j] = all[j].idx;
            out_score[qi * top_k + j] = all[j].score;
        }
        out_count[qi] = keep;
        free(all);
    }
    return rc;
}

/* Cross-query merge in _parallel_chunk_extraction_with_similarity:
 * src/components/pre_llm_injector.py:235-249 -- final[id] = max score (strict `>` replace,
 * first-seen insertion order kept), sorted(..., reverse=True) stable, [:top_k_similar_batch]. */
int64_t vo_merge_max_by_id(const int64_t *idx, const double *score, const int64_t *count, int64_t q, int64_t top_k,
                           int64_t top_k2, int64_t *out_idx, double *out_score)
{
    int64_t cap = q * top_k, m = 0;
    vo_pair *uniq = (vo_pair *)malloc(sizeof(vo_pair) * (size_t)(cap > 0 ? cap : 1));
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cap > 0 ? cap : 1));
    for (int64_t qi = 0; qi < q; ++qi)
        for (int64_t j = 0; j < count[qi]; ++j) {
            int64_t id = idx[qi * top_k + j];
            double s = score[qi * top_k + j];
            int64_t f = -1;
            for (int64_t u = 0; u < m; ++u) if (uniq[u].idx == id) { f = u; break; }
            if (f < 0) { uniq[m].idx = id; uniq[m].score = s; ++m; }
            else if (s > uniq[f].score) uniq[f].score = s;
        }
    /* stable sort by score desc over first-seen order: sort positions, tie -> position */
    for (int64_t u = 0; u < m; ++u) order[u] = u;
    for (int64_t a = 1; a < m; ++a) { /* insertion sort: stable, m is tiny */
        int64_t o = order[a]; int64_t b = a;
        while (b > 0 && uniq[order[b - 1]].score < uniq[o].score) { order[b] = order[b - 1]; --b; }
        order[b] = o;
    }
    int64_t keep = m < top_k2 ? m : top_k2;
    for (int64_t j = 0; j < keep; ++j) { out_idx[j] = uniq[order[j]].idx; out_score[j] = uniq[order[j]].score; }
    free(uniq); free(order);
    return keep;
}

/* HybridRetriever._vector_search_chunks Cypher: src/pipeline/retriever_hybrid.py:293-306
 *   WHERE c.embedding IS NOT NULL ; similarity = vector.similarity.cosine(c.embedding, $q)
 *   WHERE similarity > 0.3 ; ORDER BY score DESC ; LIMIT k
 * PARITY UNPINNED: the arithmetic lives in the Neo4j 5 server (not vendored, no JVM here).
 * Restated per SURVEY.md 9.3: score = (1 + cos)/2, strict threshold on that value, tie order
 * (unspecified by Cypher) = lowest row; zero-norm -> cos 0.0.  Raw cosine is computed with the
 * binary64 injector formula; the normalisation is then a single binary64 operation. */
int vo_vector_search(const double *query, const double *store, const uint8_t *row_ok, int64_t n, int64_t d,
                     int64_t top_k, double min_score, int sum_mode, int64_t *out_idx, double *out_score,
                     int64_t *out_count)
{
    vo_pair *all = (vo_pair *)malloc(sizeof(vo_pair) * (size_t)(n > 0 ? n : 1));
    if (!all) return -1;
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (row_ok && !row_ok[i]) continue;
        double c = vo_cosine_injector(query, d, store + i * d, d, sum_mode);
        double s = (1.0 + c) / 2.0;
        if (s > min_score) { all[m].score = s; all[m].idx = i; ++m; }
    }
    qsort(all, (size_t)m, sizeof(vo_pair), cmp_desc_stable);
    int64_t keep = m < top_k ? m : top_k;
    for (int64_t j = 0; j < keep; ++j) { out_idx[j] = all[j].idx; out_score[j] = all[j].score; }
    *out_count = keep;
    free(all);
    return 0;
}

/* HybridRetriever._post_compress_chunks filter: src/pipeline/retriever_hybrid.py:492-504
 * keep segment i iff retriever-cosine(query, segment_i) >= threshold (inclusive), original
 * order, then [:top_k] (:509).  Returns number kept; out_idx/out_score hold the kept ones. */
int64_t vo_threshold_filter_ge(const double *query, const double *segs, int64_t n, int64_t d, double threshold,
                               int64_t top_k, int sum_mode, int64_t *out_idx, double *out_score)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n && m < top_k; ++i) {
        double s = vo_cosine_retriever(query, d, segs + i * d, d, sum_mode);
        if (s >= threshold) { out_idx[m] = i; out_score[m] = s; ++m; }
    }
    return m;
}

/* ---- prune.py all-pairs (float32, sklearn semantics) -------------------------------- */
/* sklearn.metrics.pairwise.cosine_similarity (third-party, sklearn >= 1.3 per
 * requirements.txt:22; 1.9.0 installed): X_n = normalize(X) -- row / ||row||_2 in the input
 * dtype (float32 here), zero rows left untouched -- then X_n @ X_n.T.  The BLAS summation
 * order is not specified, so this restatement is compared with a tolerance, not bit-exactly:
 * it accumulates the float32-normalised rows in binary64 and rounds once to float32. */
static void normalize_rows_f32(const float *x, int64_t n, int64_t d, float *xn)
{
    for (int64_t i = 0; i < n; ++i) {
        double ss = 0.0;
        for (int64_t k = 0; k < d; ++k) ss += (double)x[i * d + k] * (double)x[i * d + k];
        float nrm = (float)sqrt(ss);
        if (nrm == 0.0f) nrm = 1.0f; /* sklearn _handle_zeros_in_scale */
        for (int64_t k = 0; k < d; ++k) xn[i * d + k] = x[i * d + k] / nrm;
    }
}

/* Graph._are_same_context generalised: src/pipeline/prune.py:67-79
 * S = cosine_similarity(E); fill_diagonal(S, 0); hits = S > threshold (strict).
 * Emits every (i, j, S_ij) with i < j (S is symmetric) in (i, j) order; returns the total
 * number of hits (may exceed cap; only the first cap are written). `any(hits)` is count>0. */
int64_t vo_pairs_above(const float *x, int64_t n, int64_t d, float threshold, int64_t cap, int64_t *out_i,
                       int64_t *out_j, float *out_score)
{
    if (n <= 1) return 0; /* prune.py:73-74 */
    float *xn = (float *)malloc(sizeof(float) * (size_t)(n * d));
    if (!xn) return -1;
    normalize_rows_f32(x, n, d, xn);
    int64_t total = 0;
    int64_t *row_cnt = (int64_t *)calloc((size_t)n, sizeof(int64_t));
    /* pass 1: per-row counts (parallel) ; pass 2: write in (i,j) order */
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        int64_t c = 0;
        for (int64_t j = i + 1; j < n; ++j) {
            double acc = 0.0;
            for (int64_t k = 0; k < d; ++k) acc += (double)xn[i * d + k] * (double)xn[j * d + k];
            if ((float)acc > threshold) ++c;
        }
        row_cnt[i] = c;
    }
    int64_t *row_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) { row_off[i] = total; total += row_cnt[i]; }
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n; ++i) {
        if (!row_cnt[i]) continue;
        int64_t w = row_off[i];
        for (int64_t j = i + 1; j < n; ++j) {
            double acc = 0.0;
            for (int64_t k = 0; k < d; ++k) acc += (double)xn[i * d + k] * (double)xn[j * d + k];
            float s = (float)acc;
            if (s > threshold) {
                if (w < cap) { out_i[w] = i; out_j[w] = j; out_score[w] = s; }
                ++w;
            }
        }
    }
    free(xn); free(row_cnt); free(row_off);
    return total;
}

/* Helpers of the streamed all-pairs tier (oracle.pairs_above_streamed): the SAME arithmetic as vo_pairs_above, applied
 * to a candidate list instead of every pair.  vo_normalize_rows_f32 exposes the normalisation above;
 * vo_pairs_rescore writes (float)sum_k (double)xn_i[k] * (double)xn_j[k] for each candidate (i, j). */
void vo_normalize_rows_f32(const float *x, int64_t n, int64_t d, float *xn)
{
    normalize_rows_f32(x, n, d, xn);
}

void vo_pairs_rescore(const float *xn, int64_t d, const int64_t *ci, const int64_t *cj, int64_t m, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < m; ++p) {
        const float *a = xn + ci[p] * d, *b = xn + cj[p] * d;
        double acc = 0.0;
        for (int64_t k = 0; k < d; ++k) acc += (double)a[k] * (double)b[k];
        out[p] = (float)acc;
    }
}

/* Graph._get_representative_relation: src/pipeline/prune.py:56-65
 * centroid = mean(E, axis=0) (float32); sims = cosine_similarity([centroid], E)[0];
 * argmax -> first maximal index.  Writes sims (float32) if out_sims != NULL. */
int64_t vo_representative(const float *x, int64_t n, int64_t d, float *out_sims)
{
    if (n <= 0) return -1;
    float *c = (float *)malloc(sizeof(float) * (size_t)d);
    float *xn = (float *)malloc(sizeof(float) * (size_t)(n * d));
    float *cn = (float *)malloc(sizeof(float) * (size_t)d);
    for (int64_t k = 0; k < d; ++k) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += (double)x[i * d + k];
        c[k] = (float)(s / (double)n);
    }
    normalize_rows_f32(x, n, d, xn);
    normalize_rows_f32(c, 1, d, cn);
    int64_t best = 0; float bests = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int64_t k = 0; k < d; ++k) acc += (double)cn[k] * (double)xn[i * d + k];
        float s = (float)acc;
        if (out_sims) out_sims[i] = s;
        if (i == 0 || s > bests) { best = i; bests = s; }
    }
    free(c); free(xn); free(cn);
    return best;
}

/* ---- exact rescoring helper used by the blocked large-N oracle (oracle.py) ---------- */
/* scores[j] = vo_cosine_injector(query, store_row[rows[j]]) for a list of rows: lets the
 * numpy tier pre-rank with a BLAS float64 matmul and then fix order/score bit-exactly. */
void vo_rescore_rows_f32(const float *query, const float *store, int64_t d, const int64_t *rows, int64_t m,
                         int sum_mode, double *out)
{
    double *qd = (double *)malloc(sizeof(double) * (size_t)d);
    for (int64_t k = 0; k < d; ++k) qd[k] = (double)query[k];
#pragma omp parallel
    {
        double *rd = (double *)malloc(sizeof(double) * (size_t)d);
#pragma omp for schedule(static)
        for (int64_t j = 0; j < m; ++j) {
            const float *r = store + rows[j] * d;
            for (int64_t k = 0; k < d; ++k) rd[k] = (double)r[k];
            out[j] = vo_cosine_injector(qd, d, rd, d, sum_mode);
        }
        free(rd);
    }
    free(qd);
}

/* ---- synthetic generator: counter hash -> integer in [-127,127] / 128 ---- */
/* Shared definition with the device generator (csrc/synth.cuh) and oracle/synth.py. */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline float synth_value(uint64_t seed, uint64_t row, uint64_t col)
{
    uint64_t rowkey = splitmix64(seed * 0x9E3779B97F4A7C15ULL + row);
    uint64_t h = splitmix64(rowkey + (col >> 3));
    uint32_t b = (uint32_t)((h >> (8 * (col & 7))) & 0xFF);
    int v = (int)((b * 255u) >> 8) - 127;
    return (float)v / 128.0f;
}

/* Planted near-duplicates: with dup_period > 0, row r > 0 is a
 * near-copy of an earlier row `parent` when hash(r) % dup_period == 0: 15/16 of its columns
 * repeat the parent's base values, the rest keep its own -> cosine ~ 0.94. */
static inline float synth_value_dup(uint64_t seed, uint64_t row, uint64_t col, uint64_t dup_period)
{
    if (dup_period > 0 && row > 0) {
        uint64_t hr = splitmix64(splitmix64(seed ^ 0xD6E8FEB86659FD93ULL) + row);
        if (hr % dup_period == 0) {
            uint64_t parent = splitmix64(hr) % row;
            uint64_t hc = splitmix64(hr + 0x632BE59BD9B4E019ULL + (col >> 3));
            uint32_t b = (uint32_t)((hc >> (8 * (col & 7))) & 0xFF);
            if ((b & 15u) != 0u) return synth_value(seed, parent, col);
        }
    }
    return synth_value(seed, row, col);
}

void vo_synth_rows(uint64_t seed, int64_t row0, int64_t n, int64_t d, uint64_t dup_period, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        for (int64_t k = 0; k < d; ++k)
            out[i * d + k] = synth_value_dup(seed, (uint64_t)(row0 + i), (uint64_t)k, dup_period);
}

int vo_version(void) { return 1; }
