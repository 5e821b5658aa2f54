/*
 * ldpc_oracle.c -- CPU restatement of the reference's flooding decoders.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or call
 * this file; it is used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
 * reference arm as the checker and the timed CPU port.
 *
 * Parity pinning: the reference has no golden vectors.  This restatement is
 * pinned against the reference's own OpenCL kernel sources executed on the CPU through
 * oracle/build_ref.py (oracle/_ref/libref_kernels.so); the comparison lives in
 * tests/test_oracle_vs_ref.py and the outputs are frozen in tests/golden/*.npz.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference repository root).  Data types are the reference's: int32 messages for the
 * IB decoders, float64 for min-sum / BP, buffers laid out [row * B + frame].
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC ldpc_oracle.c -o libldpc_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n_var, n_chk, n_edge;
    const int *sc, *dc, *tc; /* inbox_memory_start_checknodes, degree_checknode_nr, target_memorycells_checknodes */
    const int *sv, *dv, *tv; /* same for variable nodes */
} oracle_graph;

/* send_channel_values_to_checknode_inbox: kernels_template_irreg.cl:13-31 (int32),
 * kernels_min_and_BP.cl:12-29 (float64). */
#define DEFINE_SEND(NAME, TYPE)                                                            \
    static void NAME(const oracle_graph *g, const TYPE *ch, TYPE *cin, int B)              \
    {                                                                                      \
        _Pragma("omp parallel for schedule(static)")                                       \
        for (int v = 0; v < g->n_var; v++)                                                 \
            for (int w = 0; w < g->dv[v]; w++) {                                           \
                int cell = g->tv[g->sv[v] + w];                                            \
                memcpy(cin + (size_t)cell * B, ch + (size_t)v * B, sizeof(TYPE) * (size_t)B); \
            }                                                                              \
    }
DEFINE_SEND(send_i32, int)
DEFINE_SEND(send_f64, double)

/* checknode_update_iter0: kernels_template_irreg.cl:33-99 (regular twin kernels_template.cl:32-90).
 * Degree-2 checks: the reference reads an uninitialised second operand
 * (kernels_template_irreg.cl:72); defined here as pass-through of the single other input. */
static void ib_cn_iter0(const oracle_graph *g, const int *cin, int *vin, int B, int Tc, int T,
                        const int *C, int match, const int *MC)
{
#pragma omp parallel for schedule(static)
    for (int c = 0; c < g->n_chk; c++) {
        int d = g->dc[c], s = g->sc[c];
        int m[64];
        for (int f = 0; f < B; f++)
            for (int w = 0; w < d; w++) {
                int pos = 0;
                for (int v = 0; v < d; v++)
                    if (v != w) m[pos++] = cin[(size_t)(s + v) * B + f];
                int t = (d >= 3) ? C[m[0] * Tc + m[1]] : m[0];
                for (int l = 1; l < d - 2; l++)
                    t = C[t * T + m[l + 1] + Tc * Tc + (l - 1) * Tc * T];
                if (match) t = MC[(d - 1) * T + t];
                vin[(size_t)g->tc[s + w] * B + f] = t;
            }
    }
}

/* varnode_update: kernels_template_irreg.cl:103-179 (regular twin kernels_template.cl:94-171). */
static void ib_vn_update(const oracle_graph *g, const int *ch, const int *vin, int *cin, int B,
                         int Tc, int T, int DV, int it, const int *V, int match, const int *MV)
{
    const int off = it * (Tc * T + (DV - 1) * T * T);
#pragma omp parallel for schedule(static)
    for (int v = 0; v < g->n_var; v++) {
        int d = g->dv[v], s = g->sv[v];
        int m[65];
        for (int f = 0; f < B; f++) {
            m[0] = ch[(size_t)v * B + f];
            if (d == 1) { /* :132-136 raw channel value, no LUT, no matching */
                cin[(size_t)g->tv[s] * B + f] = m[0];
                continue;
            }
            for (int w = 0; w < d; w++) {
                int pos = 1;
                for (int u = 0; u < d; u++)
                    if (u != w) m[pos++] = vin[(size_t)(s + u) * B + f];
                int t = V[off + m[0] * T + m[1]];
                for (int l = 1; l <= d - 2; l++)
                    t = V[off + t * T + m[l + 1] + Tc * T + (l - 1) * T * T];
                if (match) t = MV[it * T * DV + (d - 1) * T + t];
                cin[(size_t)g->tv[s + w] * B + f] = t;
            }
        }
    }
}

/* checknode_update (iterations >= 1): kernels_template_irreg.cl:181-246. */
static void ib_cn_update(const oracle_graph *g, const int *cin, int *vin, int B, int Tc, int T,
                         int DC, int it, const int *C, int match, const int *MC)
{
    const int off = Tc * Tc + (DC - 3) * Tc * T + it * ((DC - 2) * T * T);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < g->n_chk; c++) {
        int d = g->dc[c], s = g->sc[c];
        int m[64];
        for (int f = 0; f < B; f++)
            for (int w = 0; w < d; w++) {
                int pos = 0;
                for (int v = 0; v < d; v++)
                    if (v != w) m[pos++] = cin[(size_t)(s + v) * B + f];
                int t = m[0];
                for (int l = 0; l < d - 2; l++)
                    t = C[off + t * T + m[l + 1] + l * T * T];
                if (match) t = MC[(it + 1) * T * DC + (d - 1) * T + t];
                vin[(size_t)g->tc[s + w] * B + f] = t;
            }
    }
}

/* calc_syndrome + cl_array.sum: kernels_template_irreg.cl:304-325, discrete_LDPC_decoder_irreg.py:319.
 * Returns the batch-wide sum of per-check parities. */
static long ib_syndrome_sum(const oracle_graph *g, const int *cin, int B, int T)
{
    long total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int c = 0; c < g->n_chk; c++)
        for (int f = 0; f < B; f++) {
            int p = 0;
            for (int w = 0; w < g->dc[c]; w++)
                p ^= (cin[(size_t)(g->sc[c] + w) * B + f] < T / 2);
            total += p;
        }
    return total;
}

/* calc_varnode_output: kernels_template_irreg.cl:249-302 (no matching on the output). */
static void ib_vn_output(const oracle_graph *g, const int *ch, const int *vin, int *out, int B,
                         int Tc, int T, int DV, int it, const int *V)
{
    const int off = it * (Tc * T + (DV - 1) * T * T);
#pragma omp parallel for schedule(static)
    for (int v = 0; v < g->n_var; v++) {
        int d = g->dv[v], s = g->sv[v];
        for (int f = 0; f < B; f++) {
            int t = V[off + ch[(size_t)v * B + f] * T + vin[(size_t)s * B + f]];
            for (int l = 1; l < d; l++)
                t = V[off + t * T + vin[(size_t)(s + l) * B + f] + Tc * T + (l - 1) * T * T];
            out[(size_t)v * B + f] = t;
        }
    }
}

/* Host loop: discrete_LDPC_decoder.py:202-295 / discrete_LDPC_decoder_irreg.py:245-341.
 * early != 0 reproduces the batch-granular early termination (one sum over the whole
 * syndrome buffer); early == 0 always runs imax-1 passes (the timed mode).
 * Returns i_num (the number the reference loop ends with). */
int oracle_ib_decode(int n_var, int n_chk, int n_edge, const int *sc, const int *dc, const int *tc,
                     const int *sv, const int *dv, const int *tv, int DC, int DV, int Tc, int T,
                     int imax, int match, const int *C, const int *V, const int *MC, const int *MV,
                     const int *ch, int B, int early, int *out)
{
    oracle_graph g = {n_var, n_chk, n_edge, sc, dc, tc, sv, dv, tv};
    int *cin = (int *)malloc(sizeof(int) * (size_t)n_edge * B);
    int *vin = (int *)malloc(sizeof(int) * (size_t)n_edge * B);
    if (!cin || !vin) { free(cin); free(vin); return -1; }
    send_i32(&g, ch, cin, B);
    ib_cn_iter0(&g, cin, vin, B, Tc, T, C, match, MC);
    int i_num = 1, zero = 0;
    while (i_num < imax && !zero) {
        ib_vn_update(&g, ch, vin, cin, B, Tc, T, DV, i_num - 1, V, match, MV);
        ib_cn_update(&g, cin, vin, B, Tc, T, DC, i_num - 1, C, match, MC);
        if (early && ib_syndrome_sum(&g, cin, B, T) == 0) zero = 1;
        i_num++;
    }
    ib_vn_output(&g, ch, vin, out, B, Tc, T, DV, i_num - 1, V);
    free(cin);
    free(vin);
    return i_num;
}

/* ---------------------------------------------------------------- min-sum / BP (float64) */

static inline double sgn(double x) { return (x > 0.0) - (x < 0.0); } /* OpenCL sign(): sign(0)=0 */
static inline double clip150(double x) { return sgn(x) * fmin(150.0, sgn(x) * x); }

/* boxplus: kernels_min_and_BP.cl:5-9 */
static inline double boxplus(double a, double b)
{
    double bp = log((1 + exp(a + b)) / (exp(a) + exp(b)));
    return clip150(bp);
}

/* checknode_update_minsum (:126-167) and checknode_update BP (:32-71) */
static void llr_cn_update(const oracle_graph *g, const double *cin, double *vin, int B, int bp)
{
#pragma omp parallel for schedule(static)
    for (int c = 0; c < g->n_chk; c++) {
        int d = g->dc[c], s = g->sc[c];
        double m[64];
        for (int f = 0; f < B; f++)
            for (int w = 0; w < d; w++) {
                int pos = 0;
                for (int v = 0; v < d; v++)
                    if (v != w) m[pos++] = cin[(size_t)(s + v) * B + f];
                double t = m[0];
                if (bp) {
                    for (int l = 0; l < d - 2; l++) t = boxplus(m[l + 1], t);
                    t = clip150(t);
                } else {
                    for (int l = 0; l < d - 2; l++)
                        t = sgn(m[l + 1] * t) * fmin(sgn(t) * t, sgn(m[l + 1]) * m[l + 1]);
                }
                vin[(size_t)g->tc[s + w] * B + f] = t;
            }
    }
}

/* varnode_update (LLR): kernels_min_and_BP.cl:76-123, sequential sum in slot order, clip +-150 */
static void llr_vn_update(const oracle_graph *g, const double *ch, const double *vin, double *cin, int B)
{
#pragma omp parallel for schedule(static)
    for (int v = 0; v < g->n_var; v++) {
        int d = g->dv[v], s = g->sv[v];
        double m[65];
        for (int f = 0; f < B; f++) {
            m[0] = ch[(size_t)v * B + f];
            for (int w = 0; w < d; w++) {
                int pos = 1;
                for (int u = 0; u < d; u++)
                    if (u != w) m[pos++] = vin[(size_t)(s + u) * B + f];
                /* degree-1 node: the reference reads one uninitialised addend (:113);
                 * defined here as the channel value alone. */
                double t = (d >= 2) ? m[0] + m[1] : m[0];
                for (int l = 1; l <= d - 2; l++) t = t + m[l + 1];
                cin[(size_t)g->tv[s + w] * B + f] = clip150(t);
            }
        }
    }
}

/* calc_syndrome (:206-227) + sum */
static long llr_syndrome_sum(const oracle_graph *g, const double *cin, int B)
{
    long total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int c = 0; c < g->n_chk; c++)
        for (int f = 0; f < B; f++) {
            int p = 0;
            for (int w = 0; w < g->dc[c]; w++) p ^= (cin[(size_t)(g->sc[c] + w) * B + f] < 0);
            total += p;
        }
    return total;
}

/* calc_varnode_output (:170-204): unclipped sum of channel + all inbox values */
static void llr_vn_output(const oracle_graph *g, const double *ch, const double *vin, double *out, int B)
{
#pragma omp parallel for schedule(static)
    for (int v = 0; v < g->n_var; v++) {
        int d = g->dv[v], s = g->sv[v];
        for (int f = 0; f < B; f++) {
            double t = ch[(size_t)v * B + f] + vin[(size_t)s * B + f];
            for (int l = 1; l < d; l++) t = t + vin[(size_t)(s + l) * B + f];
            out[(size_t)v * B + f] = t;
        }
    }
}

/* Host loop: min_sum_decoder_irreg.py:221-287, bp_decoder_irreg.py:221-286 (CN update first,
 * at most imax-1 passes).  algo: 0 = min-sum, 1 = BP. */
int oracle_llr_decode(int n_var, int n_chk, int n_edge, const int *sc, const int *dc, const int *tc,
                      const int *sv, const int *dv, const int *tv, int algo, int imax,
                      const double *ch, int B, int early, double *out)
{
    oracle_graph g = {n_var, n_chk, n_edge, sc, dc, tc, sv, dv, tv};
    double *cin = (double *)malloc(sizeof(double) * (size_t)n_edge * B);
    double *vin = (double *)malloc(sizeof(double) * (size_t)n_edge * B);
    if (!cin || !vin) { free(cin); free(vin); return -1; }
    send_f64(&g, ch, cin, B);
    int i_num = 1, zero = 0;
    while (i_num < imax && !zero) {
        llr_cn_update(&g, cin, vin, B, algo);
        llr_vn_update(&g, ch, vin, cin, B);
        if (early && llr_syndrome_sum(&g, cin, B) == 0) zero = 1;
        i_num++;
    }
    llr_vn_output(&g, ch, vin, out, B);
    free(cin);
    free(vin);
    return i_num;
}

/* ---------------------------------------------------------------- quantizer */

/* quantize: kernels_quanti_template.cl:2-27; cluster = #{w in [1,card) : x - limits[w] > 0} */
void oracle_quantize(int card, const double *x, const double *limits, long n, int *clusters)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        int c = 0;
        for (int w = 1; w != card; w++)
            if ((x[i] - limits[w]) > 0) c++;
        clusters[i] = c;
    }
}

/* quantize_LLR: kernels_quanti_template.cl:29-52 */
void oracle_quantize_llr(int card, const double *x, const double *limits, const double *llr_vector,
                         long n, double *llrs)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) {
        int c = 0;
        for (int w = 1; w != card; w++)
            if ((x[i] - limits[w]) > 0) c++;
        llrs[i] = llr_vector[c];
    }
}

/* return_errors_all_zero: discrete_LDPC_decoder_irreg.py:343-349 (rows [:data_len], out < T/2)
 * and min_sum_decoder_irreg.py:290-295 (LLR < 0). */
long oracle_count_errors_i32(const int *out, long rows, long B, int threshold)
{
    long e = 0;
    for (long i = 0; i < rows * B; i++) e += out[i] < threshold;
    return e;
}
long oracle_count_errors_f64(const double *out, long rows, long B)
{
    long e = 0;
    for (long i = 0; i < rows * B; i++) e += out[i] < 0;
    return e;
}
