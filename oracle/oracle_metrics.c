/* CPU oracle (plain C) for the OOD metric stage: AUROC / AP / FPR@95TPR.
 *
 * TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by, or called from the
 * product path (multishiftseg_b200/).  Built by oracle/Makefile into
 * oracle/liboracle_metrics.so and used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker / timed CPU baseline.
 *
 * Parity status: PINNED -- tests/test_oracle_metrics.py checks this library
 * against tests/golden/metrics_golden.json, which tests/golden/make_golden.py
 * produced by executing the reference's own lib/utils/metric.py.
 *
 * What it restates (all citations relative to /root/reference or the image's
 * site-packages):
 *   lib/utils/metric.py:170-180  eval_ood_measure     label==0 / label==1 selection, None on empty class
 *   lib/utils/metric.py:130-153  get_measures         roc_auc_score / average_precision_score / fpr_and_fdr_at_recall
 *   lib/utils/metric.py:87-127   fpr_and_fdr_at_recall
 *   sklearn/metrics/_ranking.py:878-921   sort descending, thresholds where the float32 score changes
 *   sklearn/metrics/_ranking.py:1023-1045 tps = cumsum(y)[idx] (float64), fps = 1 + idx - tps
 *   sklearn/metrics/_ranking.py:1331-1378 roc_curve, drop_intermediate=True, prepend (0,0), fpr/tpr
 *   sklearn/metrics/_ranking.py:53-116    auc -> scipy trapezoid: sum(d * (y1 + y0) / 2.0)
 *   sklearn/metrics/_ranking.py:1160-1208 precision_recall_curve (reversed, (1,0) appended)
 *   sklearn/metrics/_ranking.py:243-260   AP = max(0, -sum(diff(recall) * precision[:-1]))
 *   numpy/_core/src/umath/loops_utils.h.src DOUBLE_pairwise_sum (np.sum's summation order)
 *
 * Compile with -ffp-contract=off: numpy evaluates every product and sum as a
 * separate rounded float64 operation (no FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PW_BLOCKSIZE 128

/* numpy's pairwise summation, float64, unit stride. */
double oracle_pairwise_sum(const double *a, int64_t n)
{
    if (n < 8) {
        double res = -0.0;
        for (int64_t i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= PW_BLOCKSIZE) {
        double r[8], res;
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8) {
            r[0] += a[i + 0]; r[1] += a[i + 1]; r[2] += a[i + 2]; r[3] += a[i + 3];
            r[4] += a[i + 4]; r[5] += a[i + 5]; r[6] += a[i + 6]; r[7] += a[i + 7];
        }
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return oracle_pairwise_sum(a, n2) + oracle_pairwise_sum(a + n2, n - n2);
    }
}

static inline uint32_t key_desc(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if (u == 0x80000000u) u = 0;                       /* -0.0 == +0.0: one threshold */
    uint32_t asc = (u >> 31) ? ~u : (u | 0x80000000u); /* ascending-order key */
    return ~asc;                                       /* ascending key == descending score */
}

static inline int64_t load_label(const void *labels, int label_bytes, int64_t i)
{
    switch (label_bytes) {
    case 1: return ((const uint8_t *)labels)[i];
    case 4: return ((const int32_t *)labels)[i];
    default: return ((const int64_t *)labels)[i];
    }
}

/* LSD radix sort of (key, label) pairs, 4 x 8-bit digits, stable. */
static void radix_sort_pairs(uint32_t *k, uint8_t *v, uint32_t *k2, uint8_t *v2, int64_t n)
{
    for (int pass = 0; pass < 4; pass++) {
        int64_t hist[256] = {0};
        int shift = 8 * pass;
        for (int64_t i = 0; i < n; i++) hist[(k[i] >> shift) & 255]++;
        int64_t sum = 0;
        for (int b = 0; b < 256; b++) { int64_t c = hist[b]; hist[b] = sum; sum += c; }
        for (int64_t i = 0; i < n; i++) {
            int64_t d = hist[(k[i] >> shift) & 255]++;
            k2[d] = k[i]; v2[d] = v[i];
        }
        uint32_t *tk = k; k = k2; k2 = tk;
        uint8_t *tv = v; v = v2; v2 = tv;
    }
    /* 4 passes: result is back in the original (k, v) buffers */
}

/* Integer stage.  On success returns 0 and mallocs *tps_out / *fps_out (T entries
 * each, caller frees with oracle_free).  1 = a class is empty (reference returns
 * None), -1 = NaN among valid scores, -2 = infinity among valid scores. */
int oracle_ood_counts(const float *conf, const void *labels, int label_bytes, int64_t n,
                      int64_t id_in, int64_t id_out,
                      int64_t **tps_out, int64_t **fps_out, int64_t *T_out)
{
    int64_t m = 0, P = 0;
    int has_nan = 0, has_inf = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t l = load_label(labels, label_bytes, i);
        if (l == id_in || l == id_out) { m++; P += (l == id_out); }
    }
    if (P == 0 || P == m) return 1;
    uint32_t *k = malloc(sizeof(uint32_t) * m), *k2 = malloc(sizeof(uint32_t) * m);
    uint8_t *v = malloc(m), *v2 = malloc(m);
    int64_t j = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t l = load_label(labels, label_bytes, i);
        if (l == id_in || l == id_out) {
            float f = conf[i];
            if (isnan(f)) has_nan = 1;
            if (isinf(f)) has_inf = 1;
            k[j] = key_desc(f);
            v[j] = (l == id_out);
            j++;
        }
    }
    if (has_nan || has_inf) { free(k); free(k2); free(v); free(v2); return has_nan ? -1 : -2; }
    radix_sort_pairs(k, v, k2, v2, m);
    int64_t T = 1;
    for (int64_t i = 1; i < m; i++) T += (k[i] != k[i - 1]);
    int64_t *tps = malloc(sizeof(int64_t) * T), *fps = malloc(sizeof(int64_t) * T);
    int64_t t = 0, cp = 0;
    for (int64_t i = 0; i < m; i++) {
        cp += v[i];
        if (i == m - 1 || k[i] != k[i + 1]) { tps[t] = cp; fps[t] = i + 1 - cp; t++; }
    }
    free(k); free(k2); free(v); free(v2);
    *tps_out = tps; *fps_out = fps; *T_out = T;
    return 0;
}

void oracle_free(void *p) { free(p); }

/* float64 tail over (tps, fps); out = {AUROC, AP, FPR95}; aux = {T_kept, k_cutoff}. */
void oracle_metrics_from_counts(const int64_t *tps, const int64_t *fps, int64_t T,
                                double recall_level, double out[3], int64_t aux[2])
{
    const double P = (double)tps[T - 1], N = (double)fps[T - 1];

    /* AUROC: drop collinear points, prepend the origin, trapezoid */
    double *terms = calloc((size_t)(T > 0 ? T : 1), sizeof(double));
    int64_t nt = 0;
    double pf = 0.0 / N, pt = 0.0 / P;     /* the prepended (0, 0) point */
    for (int64_t k = 0; k < T; k++) {
        int keep = 1;
        if (T > 2 && k > 0 && k < T - 1) {
            int64_t d2f = fps[k + 1] - 2 * fps[k] + fps[k - 1];
            int64_t d2t = tps[k + 1] - 2 * tps[k] + tps[k - 1];
            keep = (d2f != 0) || (d2t != 0);
        }
        if (!keep) continue;
        double f = (double)fps[k] / N, t = (double)tps[k] / P;
        double d = f - pf;
        double s = t + pt;
        terms[nt++] = (d * s) / 2.0;
        pf = f; pt = t;
    }
    out[0] = oracle_pairwise_sum(terms, nt);
    aux[0] = nt;

    /* AP over reversed arrays: j = 0..T-1 <-> k = T-1-j; recall[-1] := 0, precision appended 1 unused */
    for (int64_t j = 0; j < T; j++) {
        int64_t k = T - 1 - j;
        double rec_k = (double)tps[k] / P;
        double rec_prev = (k > 0) ? (double)tps[k - 1] / P : 0.0;
        double prec_k = (double)tps[k] / (double)(tps[k] + fps[k]); /* f64(tps)+f64(fps) is exact */
        double d = rec_prev - rec_k;
        terms[j] = d * prec_k;
    }
    double ap = -oracle_pairwise_sum(terms, T);
    out[1] = ap > 0.0 ? ap : 0.0;
    free(terms);

    /* FPR@95: first k with tps == P bounds the search; ties -> largest k */
    int64_t last = 0;
    while (tps[last] != tps[T - 1]) last++;
    double best = INFINITY;
    int64_t kbest = 0;
    for (int64_t k = 0; k <= last; k++) {
        double d = fabs((double)tps[k] / P - recall_level);
        if (d <= best) { best = d; kbest = k; }
    }
    out[2] = (double)fps[kbest] / N;
    aux[1] = kbest;
}

/* eval_ood_measure end to end.  counts = {P, N, T, T_kept}. */
int oracle_ood_metrics(const float *conf, const void *labels, int label_bytes, int64_t n,
                       int64_t id_in, int64_t id_out, double out[3], int64_t counts[4])
{
    int64_t *tps, *fps, T, aux[2];
    int rc = oracle_ood_counts(conf, labels, label_bytes, n, id_in, id_out, &tps, &fps, &T);
    if (rc) return rc;
    oracle_metrics_from_counts(tps, fps, T, 0.95, out, aux);
    counts[0] = tps[T - 1]; counts[1] = fps[T - 1]; counts[2] = T; counts[3] = aux[0];
    free(tps); free(fps);
    return 0;
}
