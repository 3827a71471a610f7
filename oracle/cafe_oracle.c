/* TEST INFRASTRUCTURE — see cafe_oracle.h.  CPU restatement of the reference's likelihood path;
 * parity pinned against the compiled reference in tests/test_oracle_golden.py. */
#include "cafe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define LG_TABLE 1024

static double g_lgamma[LG_TABLE];
static int g_ready = 0;

/* src/probability.cpp:66-72 — table of lgamma(i), i = 0..1023 (entry 0 is +inf). */
void orc_init(void)
{
    if (g_ready) return;
    for (int i = 0; i < LG_TABLE; ++i) g_lgamma[i] = lgamma((double)i);
    g_ready = 1;
}

/* src/probability.cpp:58-64 — table lookup for (near-)integers in range, libm otherwise. */
static double lgamma_cached(double n)
{
    if (n >= 0 && n < LG_TABLE && (n - (int)n < 0.00000000001)) return g_lgamma[(int)n];
    return lgamma(n);
}

/* src/probability.cpp:79-88.  The reference's 100x100 chooseln_cache holds exactly the same
 * expression evaluated from the same table (:74-76), so one formula covers both branches. */
double orc_chooseln(double n, double r)
{
    if (r == 0) return 0.0;
    if (n <= 0 || r <= 0) return log(0.0);
    return lgamma_cached(n + 1) - lgamma_cached(r + 1) - lgamma_cached(n - r + 1);
}

/* src/probability.cpp:101-128,144 — the vectorised branch (m < 10000 always holds here):
 * terms exp(t_j) * pow(coeff, j), summed in ascending j from 0.0, clamped to [0,1]. */
double orc_birthdeath_rate_with_log_alpha(int s, int c, double log_alpha, double coeff)
{
    int m = s < c ? s : c;
    double total = 0.0;
    for (int j = 0; j <= m; ++j) {
        double t = orc_chooseln(s, j) + orc_chooseln(s + c - 1 - j, s - 1) + (s + c - 2 * j) * log_alpha;
        double term = exp(t) * pow(coeff, (double)j);
        total += term;
    }
    if (total > 1.0) total = 1.0;
    if (total < 0.0) total = 0.0;
    return total;
}

/* src/probability.cpp:147-164 */
double orc_bd_probability(double lambda, double branch_length, int parent_size, int size)
{
    double alpha = lambda * branch_length / (1 + lambda * branch_length);
    double coeff = 1 - 2 * alpha;
    if (coeff > 0 && coeff != 1) return orc_birthdeath_rate_with_log_alpha(parent_size, size, log(alpha), coeff);
    return 0.0;
}

/* src/matrix_cache.h:49,55-57 */
double orc_quantise_lambda(double lambda) { return (double)((long)(lambda * 1000000000)) / 1000000000.0; }
/* src/matrix_cache.h:50,58-60 */
double orc_quantise_branch(double t) { return (double)((long)(t * 1000)) / 1000.0; }

/* src/matrix_cache.cpp:115-119 */
int orc_is_saturated(double branch_length, double lambda)
{
    double alpha = lambda * branch_length / (1 + lambda * branch_length);
    return (1 - 2 * alpha) < 0;
}

/* src/matrix_cache.cpp:121-171 with :70-77.  The matrix is computed from the QUANTISED key
 * values (:148-149).  Entry (0,0) is always 1; a saturated key leaves every other entry 0. */
void orc_build_matrix(int n, double lambda_raw, double t_raw, double* out)
{
    orc_init();
    double lambda = orc_quantise_lambda(lambda_raw);
    double t = orc_quantise_branch(t_raw);
    memset(out, 0, sizeof(double) * (size_t)n * n);
    out[0] = 1.0;
    if (orc_is_saturated(t, lambda)) return;
    for (int s = 1; s < n; ++s)
        for (int c = 0; c < n; ++c)
            out[(size_t)s * n + c] = orc_bd_probability(lambda, t, s, c);
}

/* ------------------------------------------------------------------------------------------ */
/* Discrete gamma (src/gamma.cpp).  Published algorithms AS 32, AS 70, AS 91 as the reference   */
/* carries them, with its tolerances, constants and quirks (see comments).                      */

/* src/gamma.cpp:66-116 — AS 32.  Series when x <= 1 or x < alpha, continued fraction otherwise.
 * Quirk kept: on convergence of the continued fraction the PREVIOUS convergent is used (:98-112). */
double orc_incomplete_gamma(double x, double alpha, double ln_gamma_alpha)
{
    const double accurate = 1e-8, overflow = 1e30;
    if (x == 0) return 0;
    if (x < 0 || alpha <= 0) return -1;
    double factor = exp(alpha * log(x) - x - ln_gamma_alpha);
    if (!(x > 1 && x >= alpha)) {
        double sum = 1, term = 1, rn = alpha;
        do {
            rn += 1;
            term *= x / rn;
            sum += term;
        } while (term > accurate);
        return sum * (factor / alpha);
    }
    double a = 1 - alpha, b = a + x + 1, count = 0;
    double p0 = 1, p1 = x, p2 = x + 1, p3 = x * b;
    double gin = p2 / p3;
    for (;;) {
        a += 1; b += 2; count += 1;
        double an = a * count;
        double p4 = b * p2 - an * p0;
        double p5 = b * p3 - an * p1;
        if (p5 != 0) {
            double rn = p4 / p5;
            double dif = fabs(gin - rn);
            if (dif <= accurate && dif <= accurate * rn) return 1 - factor * gin;
            gin = rn;
        }
        p0 = p2; p1 = p3; p2 = p4; p3 = p5;
        if (fabs(p4) >= overflow) { p0 /= overflow; p1 /= overflow; p2 /= overflow; p3 /= overflow; }
    }
}

/* src/gamma.cpp:203-215 — AS 70 rational approximation. */
double orc_point_normal(double prob)
{
    static const double a[5] = { -.322232431088, -1, -.342242088547, -.0204231210245, -.453642210148e-4 };
    static const double b[5] = { .0993484626060, .588581570495, .531103462366, .103537752850, .0038560700634 };
    double tail = prob < 0.5 ? prob : 1 - prob;
    if (tail < 1e-20) return -9999;
    double y = sqrt(log(1 / (tail * tail)));
    double z = y + ((((y * a[4] + a[3]) * y + a[2]) * y + a[1]) * y + a[0]) / ((((y * b[4] + b[3]) * y + b[2]) * y + b[1]) * y + b[0]);
    return prob < 0.5 ? -z : z;
}

/* src/gamma.cpp:129-186 — AS 91.  e = .5e-6, ln2 truncated to .6931471805 as in the reference. */
double orc_point_chi2(double prob, double v)
{
    const double e = .5e-6, aa = .6931471805;
    double p = prob;
    if (p < .000002 || p > .999998 || v <= 0) return -1;
    double g = lgamma(v / 2);
    double xx = v / 2;
    double c = xx - 1;
    double ch;
    if (v < -1.24 * log(p)) {
        ch = pow(p * xx * exp(g + xx * aa), 1 / xx);
        if (ch - e < 0) return ch;
    }
    else if (v <= .32) {
        ch = 0.4;
        double a = log(1 - p);
        double q;
        do {
            q = ch;
            double p1 = 1 + ch * (4.67 + ch);
            double p2 = ch * (6.73 + ch * (6.66 + ch));
            double t = -0.5 + (4.67 + 2 * ch) / p1 - (6.73 + ch * (13.32 + 3 * ch)) / p2;
            ch -= (1 - exp(a + g + .5 * ch + c * aa) * p2 / p1) / t;
        } while (fabs(q / ch - 1) - .01 > 0);
    }
    else {
        double x = orc_point_normal(p);
        double p1 = 0.222222 / v;
        ch = v * pow(x * sqrt(p1) + 1 - p1, 3.0);
        if (ch > 2.2 * v + 6) ch = -2 * (log(1 - p) - c * log(.5 * ch) + g);
    }
    double q;
    do {
        q = ch;
        double p1 = .5 * ch;
        double t = orc_incomplete_gamma(p1, xx, g);
        if (t < 0) return -1;
        double p2 = p - t;
        t = p2 * exp(xx * aa + g + p1 - c * log(ch));
        double b = t / ch;
        double a = 0.5 * t - b * c;
        double s1 = (210 + a * (140 + a * (105 + a * (84 + a * (70 + 60 * a))))) / 420;
        double s2 = (420 + a * (735 + a * (966 + a * (1141 + 1278 * a)))) / 2520;
        double s3 = (210 + a * (462 + a * (707 + 932 * a))) / 2520;
        double s4 = (252 + a * (672 + 1182 * a) + c * (294 + a * (889 + 1740 * a))) / 5040;
        double s5 = (84 + 264 * a + c * (175 + 606 * a)) / 2520;
        double s6 = (120 + c * (346 + 127 * c)) / 5040;
        ch += t * (1 + 0.5 * t * s1 - b * c * (s1 - b * (s2 - b * (s3 - b * (s4 - b * (s5 - b * s6))))));
    } while (fabs(q / ch - 1) > e);
    return ch;
}

/* src/gamma.cpp:15-52 (median == 0 branch) and :225-241 with alpha == beta. */
void orc_get_gamma(int k, double alpha, double* freq, double* rate)
{
    double beta = alpha;
    double factor = alpha / beta * k;
    double lnga1 = lgamma(alpha + 1);
    for (int i = 0; i < k - 1; ++i) freq[i] = orc_point_chi2((i + 1.0) / k, 2.0 * (alpha)) / (2.0 * (beta));
    for (int i = 0; i < k - 1; ++i) freq[i] = orc_incomplete_gamma(freq[i] * beta, alpha + 1, lnga1);
    rate[0] = freq[0] * factor;
    rate[k - 1] = (1 - freq[k - 2]) * factor;
    for (int i = 1; i < k - 1; ++i) rate[i] = (freq[i] - freq[i - 1]) * factor;
    for (int i = 0; i < k; ++i) freq[i] = 1.0 / k;
}

/* ------------------------------------------------------------------------------------------ */
/* Root priors                                                                                  */

/* src/root_distribution.cpp:15-31: expand the map into a list (key repeated count times), or a
 * list of max ones.  Returns malloc'ed list, length in *n. */
static int* expand_rootdist(const int* sizes, const int* counts, int n_entries, int mrf, int* n)
{
    int total = 0;
    if (n_entries == 0) {
        int* v = (int*)malloc(sizeof(int) * (size_t)(mrf > 0 ? mrf : 1));
        for (int i = 0; i < mrf; ++i) v[i] = 1;
        *n = mrf;
        return v;
    }
    for (int i = 0; i < n_entries; ++i) total += counts[i] > 0 ? counts[i] : 0;
    int* v = (int*)malloc(sizeof(int) * (size_t)(total > 0 ? total : 1));
    int pos = 0;
    for (int i = 0; i < n_entries; ++i)
        for (int c = 0; c < counts[i]; ++c) v[pos++] = sizes[i];
    *n = total;
    return v;
}

/* src/root_equilibrium_distribution.cpp:20-32: float(list[val]) / float(sum), 0 past the end. */
void orc_prior_uniform(const int* rootdist_sizes, const int* rootdist_counts, int n_entries, int max_root_family_size,
                       double* out, int n_out)
{
    int n = 0;
    int* list = expand_rootdist(rootdist_sizes, rootdist_counts, n_entries, max_root_family_size, &n);
    int sum = 0;
    for (int i = 0; i < n; ++i) sum += list[i];
    for (int v = 0; v < n_out; ++v) {
        float f = 0;
        if (v < n) f = (float)list[v] / (float)sum;
        out[v] = (double)f;
    }
    free(list);
}

/* src/poisson.cpp:19-36 and src/root_equilibrium_distribution.h:45-51: table length is the size
 * of the root distribution list; value returned through a float. */
void orc_prior_poisson(double poisson_lambda, const int* rootdist_sizes, const int* rootdist_counts, int n_entries,
                       int max_root_family_size, double* out, int n_out)
{
    int n = 0;
    int* list = expand_rootdist(rootdist_sizes, rootdist_counts, n_entries, max_root_family_size, &n);
    free(list);
    for (int v = 0; v < n_out; ++v) {
        float f = 0;
        if (v < n) f = (float)exp(v * log(poisson_lambda) - lgamma((double)(v + 1)) - poisson_lambda);
        out[v] = (double)f;
    }
}

/* src/error_model.cpp:25-29 */
static int nearly_equal(double x, double y) { return fabs(x - y) <= 0.01 * fabs(x); }

/* src/error_model.cpp:79-109 with the checks of set_probabilities (:31-40). */
int orc_error_model_replace_epsilon(double* probs, int rows, double old_eps, double new_eps)
{
    for (int r = 0; r < rows; ++r) {
        double* v = probs + 3 * r;
        if (!nearly_equal(old_eps, v[2])) continue;
        if (r == 0) { v[2] = new_eps; v[1] = 1 - new_eps; if (!nearly_equal(v[0], 0.0)) return -1; }
        else { v[2] = new_eps; v[0] = new_eps; v[1] = 1 - (new_eps * 2); }
        if (!nearly_equal(v[0] + v[1] + v[2], 1.0)) return -1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Transition matrices for one category: one per non-root node, shared between equal keys.      */

typedef struct matset {
    int n;                 /* matrix dimension max(mrf, mf) + 1 (src/base_model.cpp:77) */
    int n_nodes;
    double** of_node;      /* [n_nodes] -> matrix (root: NULL) */
    double** owned;
    int n_owned;
} matset;

static void matset_build(matset* ms, const orc_tree* tree, const double* lambdas, int n)
{
    ms->n = n;
    ms->n_nodes = tree->n_nodes;
    ms->of_node = (double**)calloc((size_t)tree->n_nodes, sizeof(double*));
    ms->owned = (double**)calloc((size_t)tree->n_nodes, sizeof(double*));
    ms->n_owned = 0;
    long* key_l = (long*)malloc(sizeof(long) * (size_t)tree->n_nodes);
    long* key_t = (long*)malloc(sizeof(long) * (size_t)tree->n_nodes);
    int* todo = (int*)malloc(sizeof(int) * (size_t)tree->n_nodes);
    int n_todo = 0;
    for (int v = 0; v < tree->n_nodes; ++v) {
        if (tree->parent[v] < 0) continue;
        double lam = lambdas[tree->lambda_index[v]];
        long kl = (long)(lam * 1000000000), kt = (long)(tree->branch[v] * 1000);
        int found = -1;
        for (int j = 0; j < n_todo; ++j) if (key_l[j] == kl && key_t[j] == kt) { found = j; break; }
        if (found < 0) {
            key_l[n_todo] = kl; key_t[n_todo] = kt; todo[n_todo] = v;
            ms->owned[n_todo] = (double*)malloc(sizeof(double) * (size_t)n * n);
            found = n_todo++;
        }
        ms->of_node[v] = ms->owned[found];
    }
    ms->n_owned = n_todo;
#pragma omp parallel for schedule(dynamic)
    for (int j = 0; j < n_todo; ++j) {
        int v = todo[j];
        orc_build_matrix(n, lambdas[tree->lambda_index[v]], tree->branch[v], ms->owned[j]);
    }
    free(key_l); free(key_t); free(todo);
}

static void matset_free(matset* ms)
{
    for (int j = 0; j < ms->n_owned; ++j) free(ms->owned[j]);
    free(ms->owned); free(ms->of_node);
}

/* src/matrix_cache.cpp:28-57 (scalar branch): y[s - s_min] = sum_{c = c_min..c_max} M[s][c] v[c - c_min],
 * ascending c, starting from 0. */
static void matrix_multiply(const double* m, int n, const double* v, int s_min, int s_max, int c_min, int c_max, double* y)
{
    for (int s = s_min; s <= s_max; ++s) {
        double acc = 0;
        const double* row = m + (size_t)s * n;
        for (int c = c_min; c <= c_max; ++c) acc += row[c] * v[c - c_min];
        y[s - s_min] = acc;
    }
}

/* src/probability.cpp:173-242 over src/clade.cpp:255-280 order.  work: n_nodes * (mf+1) doubles
 * + (mf+1) scratch. */
static int prune_with(const orc_tree* tree, const matset* ms, const int32_t* counts_row, const double* err, int err_rows, int err_ndev,
                      int mf, int mrf, double* work, double* out_root)
{
    int width = mf + 1;
    int rootw = mrf > width ? mrf : width;
    double* factor = work + (size_t)tree->n_nodes * rootw;
    for (int v = 0; v < tree->n_nodes; ++v) {
        double* probs = work + (size_t)v * rootw;
        int first = tree->child_offset[v], last = tree->child_offset[v + 1];
        if (first == last) {
            int obs = counts_row[tree->leaf_col[v]];
            memset(probs, 0, sizeof(double) * (size_t)width);
            if (obs < 0 || obs > mf) return -1;
            if (err) {
                /* :182-193 — stencil of deviation probabilities centred on the observed count */
                if (obs >= err_rows) return -1;
                int offset = obs - (err_ndev - 1) / 2;
                for (int i = 0; i < err_ndev; ++i) {
                    if (offset + i < 0) continue;
                    if (offset + i > mf) return -1;
                    probs[offset + i] = err[(size_t)obs * err_ndev + i];
                }
            }
            else probs[obs] = 1.0;
            continue;
        }
        int is_root = tree->parent[v] < 0;
        int s_min = is_root ? 1 : 0, s_max = is_root ? mrf : mf;
        int len = s_max - s_min + 1;
        for (int i = 0; i < len; ++i) probs[i] = 1;
        for (int e = first; e < last; ++e) {
            int ch = tree->child_list[e];
            matrix_multiply(ms->of_node[ch], ms->n, work + (size_t)ch * rootw, s_min, s_max, 0, mf, factor);
            for (int i = 0; i < len; ++i) probs[i] *= factor[i];
        }
    }
    memcpy(out_root, work + (size_t)(tree->n_nodes - 1) * rootw, sizeof(double) * (size_t)mrf);
    return 0;
}

static size_t prune_work_doubles(const orc_tree* tree, int mf, int mrf)
{
    int rootw = mrf > mf + 1 ? mrf : mf + 1;
    return (size_t)(tree->n_nodes + 1) * rootw;
}

int orc_inference_prune(const orc_tree* tree, const int32_t* counts_row, const double* lambdas, int n_lambdas,
                        const double* err, int err_rows, int err_ndev, int max_family_size, int max_root_family_size,
                        double* out_root)
{
    (void)n_lambdas;
    orc_init();
    matset ms;
    int n = (max_root_family_size > max_family_size ? max_root_family_size : max_family_size) + 1;
    matset_build(&ms, tree, lambdas, n);
    double* work = (double*)malloc(sizeof(double) * prune_work_doubles(tree, max_family_size, max_root_family_size));
    int rc = prune_with(tree, &ms, counts_row, err, err_rows, err_ndev, max_family_size, max_root_family_size, work, out_root);
    free(work);
    matset_free(&ms);
    return rc;
}

double orc_infer_family_likelihoods(const orc_tree* tree, const int32_t* counts, int64_t n_families, int n_leaves,
                                    const double* lambdas, int n_lambdas, const double* cat_probs, int k,
                                    const double* prior, const double* err, int err_rows, int err_ndev,
                                    int max_family_size, int max_root_family_size, int mode,
                                    double* family_lnl, double* cat_lk, uint8_t* failed, int64_t* n_failed)
{
    orc_init();
    int mf = max_family_size, mrf = max_root_family_size;
    int n = (mrf > mf ? mrf : mf) + 1;
    matset* ms = (matset*)malloc(sizeof(matset) * (size_t)k);
    for (int c = 0; c < k; ++c) matset_build(&ms[c], tree, lambdas + (size_t)c * n_lambdas, n);
    double* lnl = family_lnl ? family_lnl : (double*)malloc(sizeof(double) * (size_t)(n_families > 0 ? n_families : 1));
    uint8_t* bad = failed ? failed : (uint8_t*)calloc((size_t)(n_families > 0 ? n_families : 1), 1);
    size_t work_n = prune_work_doubles(tree, mf, mrf);
    int error = 0;
#pragma omp parallel
    {
        double* work = (double*)malloc(sizeof(double) * work_n);
        double* root = (double*)malloc(sizeof(double) * (size_t)mrf);
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < n_families; ++i) {
            const int32_t* row = counts + (size_t)i * n_leaves;
            bad[i] = 0;
            if (mode == ORC_BASE_LOGMAX) {
                /* src/base_model.cpp:89-106 */
                if (prune_with(tree, &ms[0], row, err, err_rows, err_ndev, mf, mrf, work, root)) { error = 1; lnl[i] = NAN; continue; }
                double best = 0;
                for (int j = 0; j < mrf; ++j) {
                    double full = log(root[j]) + log(prior[j]);
                    if (j == 0 || best < full) best = full;   /* std::max_element: first of the largest */
                }
                lnl[i] = best;
            }
            else {
                /* src/gamma_core.cpp:144-166,203-219 */
                double family = 0;
                int ok = 1;
                for (int c = 0; c < k && ok; ++c) {
                    if (prune_with(tree, &ms[c], row, err, err_rows, err_ndev, mf, mrf, work, root)) { error = 1; ok = 0; break; }
                    double sum = 0;
                    for (int j = 0; j < mrf; ++j) sum += root[j];
                    if (sum == 0.0) { ok = 0; break; }
                    double best = 0;
                    for (int j = 0; j < mrf; ++j) {
                        double full = root[j] * prior[j];
                        if (j == 0 || best < full) best = full;
                    }
                    double lk = best * cat_probs[c];
                    if (cat_lk) cat_lk[(size_t)i * k + c] = lk;
                    family += lk;
                }
                if (ok) lnl[i] = log(family);
                else { bad[i] = 1; lnl[i] = NAN; if (cat_lk) for (int c = 0; c < k; ++c) cat_lk[(size_t)i * k + c] = NAN; }
            }
        }
        free(work); free(root);
    }
    int64_t nf = 0;
    double total = 0.0;
    for (int64_t i = 0; i < n_families; ++i) { if (bad[i]) ++nf; else total += lnl[i]; }   /* std::accumulate, family order */
    if (n_failed) *n_failed = nf;
    for (int c = 0; c < k; ++c) matset_free(&ms[c]);
    free(ms);
    if (!family_lnl) free(lnl);
    if (!failed) free(bad);
    if (error) return NAN;
    if (nf > 0) return -log(0.0);     /* src/gamma_core.cpp:227-236 */
    return -total;
}

/* src/gene_family_reconstructor.cpp:13-165 for one family and one category. */
static void reconstruct_one(const orc_tree* tree, const matset* ms, const int32_t* counts_row, const double* prior,
                            int mf, int mrf, double* L, int* C, int* state, int32_t* out)
{
    int width = mf + 1;
    int n = ms->n;
    int root = tree->n_nodes - 1;
    for (int v = 0; v < tree->n_nodes; ++v) {
        double* Lv = L + (size_t)v * width;
        int* Cv = C + (size_t)v * width;
        int first = tree->child_offset[v], last = tree->child_offset[v + 1];
        if (first == last) {
            /* :13-33 — error model is NOT applied; L[0] stays 0 */
            int obs = counts_row[tree->leaf_col[v]];
            const double* m = ms->of_node[v];
            Lv[0] = 0;
            for (int i = 1; i < width; ++i) Lv[i] = m[(size_t)i * n + obs];
            for (int i = 0; i < width; ++i) Cv[i] = obs;
        }
        else if (v == root) {
            /* :35-72 — sizes 1..min(mf,mrf); prior indexed by the size itself */
            int lim = (mf < mrf ? mf : mrf) + 1;
            double max_val = -1;
            int arg = 0;
            for (int j = 1; j < lim; ++j) {
                double value = 1.0;
                for (int e = first; e < last; ++e) value *= L[(size_t)tree->child_list[e] * width + j];
                double val = value * prior[j];
                if (val > max_val) { max_val = val; arg = j; }
            }
            Cv[0] = arg;
        }
        else {
            /* :74-112 — strict '>' from -1: first maximum wins */
            const double* m = ms->of_node[v];
            for (int i = 0; i < width; ++i) {
                int max_j = 0;
                double max_val = -1;
                for (int j = 0; j < width; ++j) {
                    double value = 1.0;
                    for (int e = first; e < last; ++e) value *= L[(size_t)tree->child_list[e] * width + j];
                    double val = value * m[(size_t)i * n + j];
                    if (val > max_val) { max_j = j; max_val = val; }
                }
                Lv[i] = max_val;
                Cv[i] = max_j;
            }
        }
    }
    /* :148-163 — traceback from the root */
    state[root] = C[(size_t)root * width];
    for (int v = tree->n_nodes - 2; v >= 0; --v) {
        if (tree->child_offset[v] == tree->child_offset[v + 1]) continue;
        state[v] = C[(size_t)v * width + state[tree->parent[v]]];
    }
    int pos = 0;
    for (int v = 0; v < tree->n_nodes; ++v)
        if (tree->child_offset[v] != tree->child_offset[v + 1]) out[pos++] = state[v];
}

int orc_reconstruct(const orc_tree* tree, const int32_t* counts, int64_t n_families, int n_leaves,
                    const double* lambdas, int n_lambdas, int k, const double* prior,
                    int max_family_size, int max_root_family_size, int32_t* states)
{
    orc_init();
    int mf = max_family_size, mrf = max_root_family_size;
    int n = (mrf > mf ? mrf : mf) + 1;
    int n_internal = 0;
    for (int v = 0; v < tree->n_nodes; ++v) if (tree->child_offset[v] != tree->child_offset[v + 1]) ++n_internal;
    matset* ms = (matset*)malloc(sizeof(matset) * (size_t)k);
    for (int c = 0; c < k; ++c) matset_build(&ms[c], tree, lambdas + (size_t)c * n_lambdas, n);
#pragma omp parallel
    {
        double* L = (double*)malloc(sizeof(double) * (size_t)tree->n_nodes * (mf + 1));
        int* C = (int*)malloc(sizeof(int) * (size_t)tree->n_nodes * (mf + 1));
        int* state = (int*)malloc(sizeof(int) * (size_t)tree->n_nodes);
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = 0; i < n_families; ++i)
            for (int c = 0; c < k; ++c)
                reconstruct_one(tree, &ms[c], counts + (size_t)i * n_leaves, prior, mf, mrf, L, C, state,
                                states + ((size_t)i * k + c) * n_internal);
        free(L); free(C); free(state);
    }
    for (int c = 0; c < k; ++c) matset_free(&ms[c]);
    free(ms);
    return 0;
}

/* src/gamma_core.cpp:282-299 — sum_k catprob[k] * double(state_k), ascending k from 0.0 */
void orc_weighted_averages(const int32_t* states, int k, int n_internal, const double* cat_probs, double* out)
{
    for (int v = 0; v < n_internal; ++v) {
        double val = 0.0;
        for (int c = 0; c < k; ++c) val += cat_probs[c] * (double)states[(size_t)c * n_internal + v];
        out[v] = val;
    }
}

/* src/probability.cpp:301-308 (simulated families) and :396-399 (observed families): the likelihood both halves of
 * compute_pvalues use is the plain maximum of the root vector — one lambda set, no categories, no prior. */
int orc_root_max(const orc_tree* tree, const int32_t* counts, int64_t n_families, int n_leaves, const double* lambdas, int n_lambdas,
                 int max_family_size, int max_root_family_size, double* out)
{
    orc_init();
    int mf = max_family_size, mrf = max_root_family_size;
    int n = (mrf > mf ? mrf : mf) + 1;
    matset ms;
    matset_build(&ms, tree, lambdas, n);
    (void)n_lambdas;
    size_t work_n = prune_work_doubles(tree, mf, mrf);
    int error = 0;
#pragma omp parallel
    {
        double* work = (double*)malloc(sizeof(double) * work_n);
        double* root = (double*)malloc(sizeof(double) * (size_t)mrf);
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < n_families; ++i) {
            if (prune_with(tree, &ms, counts + (size_t)i * n_leaves, NULL, 0, 0, mf, mrf, work, root)) { error = 1; out[i] = NAN; continue; }
            double best = root[0];
            for (int j = 1; j < mrf; ++j) if (best < root[j]) best = root[j];      /* std::max_element */
            out[i] = best;
        }
        free(work); free(root);
    }
    matset_free(&ms);
    return error ? -1 : 0;
}

static int cmp_double(const void* a, const void* b)
{
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* src/probability.cpp:379-389: idx = upper_bound(conddist, v) - begin, or size-1 when nothing is greater; idx / size */
double orc_pvalue(double v, const double* sorted, int n)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (sorted[mid] > v) hi = mid; else lo = mid + 1;
    }
    int idx = lo < n ? lo : n - 1;
    return idx / (double)n;
}

/* src/probability.cpp:310 (sort of each conditional distribution) + :391-409 (max over root sizes) for every family.
 * cond: [n_root_sizes][n_sim] unsorted, sorted IN PLACE. */
void orc_pvalues(double* cond, int n_root_sizes, int n_sim, const double* observed, int64_t n_families, double* pvalues)
{
    for (int s = 0; s < n_root_sizes; ++s) qsort(cond + (size_t)s * n_sim, (size_t)n_sim, sizeof(double), cmp_double);
    for (int64_t i = 0; i < n_families; ++i) {
        double best = 0;
        for (int s = 0; s < n_root_sizes; ++s) {
            double pv = orc_pvalue(observed[i], cond + (size_t)s * n_sim, n_sim);
            if (s == 0 || best < pv) best = pv;
        }
        pvalues[i] = best;
    }
}

/* src/gene_family_reconstructor.cpp:361-400 for every (family, node); -1 where the reference returns invalid(). */
void orc_branch_probabilities(const orc_tree* tree, const int32_t* node_sizes, const uint8_t* selected, int64_t n_families,
                              const double* lambdas, int n_lambdas, int max_family_size, int max_root_family_size, double* out)
{
    orc_init();
    int mf = max_family_size, mrf = max_root_family_size;
    int n = (mrf > mf ? mrf : mf) + 1;
    int nn = tree->n_nodes;
    matset ms;
    matset_build(&ms, tree, lambdas, n);
    (void)n_lambdas;
    for (int64_t f = 0; f < n_families; ++f) {
        for (int v = 0; v < nn; ++v) {
            double result = -1.0;
            int par = tree->parent[v];
            if (par >= 0 && (!selected || selected[f])) {
                int ps = node_sizes[f * nn + par], cs = node_sizes[f * nn + v];
                if (ps != cs) {
                    const double* m = ms.of_node[v];
                    double pstar = m[(size_t)ps * n + cs];
                    double acc = 0;
                    for (int c = 0; c < mf; ++c) {
                        double pm = m[(size_t)ps * n + c];
                        if (pm == pstar) acc += pm / 2.0;
                        else if (pm < pstar) acc += pm;
                    }
                    result = acc;
                }
            }
            out[f * nn + v] = result;
        }
    }
    matset_free(&ms);
}

