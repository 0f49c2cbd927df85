/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle for the snapshot-POD sparse-sensing path).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (openmeasure_b200/) never links, imports or calls it.
 *
 * Plain-C restatements of the third-party numerics the reference reaches on its hot path
 * (the reference itself is pure Python: /root/reference/src/openmeasure/sparse_sensing.py):
 *
 *   omo_pairwise_sum / omo_block_stats
 *       numpy's pairwise add.reduce as used by np.average / np.std / np.var
 *       (sparse_sensing.py:112, :115, :121, :137), numpy >= 1.24.2 per pyproject.toml:13;
 *       pinned against numpy 2.3.5 itself in tests/test_oracle_cpu.py.
 *   omo_qrcp_dlaqp2
 *       LAPACK dgeqp3 -> dlaqp2 (+ dlarfg, dlarf), the routine behind
 *       scipy.linalg.qr(self.Ur.T, pivoting=True, mode='economic') at sparse_sensing.py:739
 *       (scipy >= 1.10.1 per pyproject.toml:14).  Restated from the published LAPACK 3.x
 *       algorithm; pinned against scipy 1.18.1 itself and against golden pivots produced by the
 *       unmodified reference (tests/golden/, oracle/make_golden.py).
 *   omo_synth_fill
 *       the deterministic synthetic snapshot generator of DESIGN.md (not a reference function;
 *       the CPU twin of the CUDA generator so both sides see bit-identical inputs).
 *
 * Arithmetic is IEEE double with explicit fma() where stated; compile with -ffp-contract=off so
 * nothing else is contracted.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* numpy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum).    */
/* ------------------------------------------------------------------------------------------ */
typedef double (*omo_map_fn)(double, double);
static double map_id(double x, double p) { (void)p; return x; }
static double map_sqdev(double x, double p) { double d = x - p; return d * d; }

static double pw(const double *a, int64_t n, omo_map_fn f, double p)
{
    if (n < 8) {
        double res = -0.0;
        for (int64_t i = 0; i < n; ++i) res += f(a[i], p);
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = f(a[k], p);
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += f(a[i + k], p);
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += f(a[i], p);
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return pw(a, n2, f, p) + pw(a + n2, n - n2, f, p);
}

double omo_pairwise_sum(const double *a, int64_t n) { return pw(a, n, map_id, 0.0); }

/* sum, sum of squared deviations about `mean`, min, max of a contiguous block (np.std's two
 * passes: arrmean = add.reduce(x)/N ; add.reduce((x-arrmean)**2)/N ; sparse_sensing.py:115). */
void omo_block_stats(const double *a, int64_t n, double *sum, double *mean, double *sqdev,
                     double *mn, double *mx)
{
    double s = pw(a, n, map_id, 0.0);
    double mu = s / (double)n;
    double q = pw(a, n, map_sqdev, mu);
    double lo = a[0], hi = a[0];
    for (int64_t i = 1; i < n; ++i) {
        if (a[i] < lo) lo = a[i];
        if (a[i] > hi) hi = a[i];
    }
    *sum = s; *mean = mu; *sqdev = q; *mn = lo; *mx = hi;
}

/* row means of a C-order (rows x m) matrix: np.average(x, axis=1), sparse_sensing.py:112. */
void omo_row_means(const double *x, int64_t rows, int64_t m, double *out)
{
    for (int64_t i = 0; i < rows; ++i) out[i] = pw(x + i * m, m, map_id, 0.0) / (double)m;
}

/* ------------------------------------------------------------------------------------------ */
/* LAPACK dgeqp3 / dlaqp2 column-pivoted QR of the r x n matrix A = Ur^T.                      */
/* A is column-major with leading dimension r, i.e. exactly the memory of a C-order (n, r) Ur. */
/* ------------------------------------------------------------------------------------------ */
static double nrm2_seq(const double *x, int64_t n)
{
    double s = 0.0;
    for (int64_t k = 0; k < n; ++k) s = fma(x[k], x[k], s);
    return sqrt(s);
}

static double lapy2(double x, double y)
{
    double xa = fabs(x), ya = fabs(y);
    double w = xa > ya ? xa : ya, z = xa > ya ? ya : xa;
    if (z == 0.0) return w;
    double q = z / w;
    return w * sqrt(1.0 + q * q);
}

/* dlarfg on (alpha, x[0..n-2]); returns tau, overwrites x with v[1:], alpha with beta.
 * The safmin rescaling loop of LAPACK is kept for fidelity. */
static double larfg(int64_t n, double *alpha, double *x)
{
    if (n <= 1) return 0.0;
    double xnorm = nrm2_seq(x, n - 1);
    if (xnorm == 0.0) return 0.0;
    double a = *alpha;
    double beta = -copysign(lapy2(a, xnorm), a);
    const double safmin = 2.2250738585072014e-308 / 1.1102230246251565e-16;
    const double rsafmn = 1.0 / safmin;
    int knt = 0;
    if (fabs(beta) < safmin) {
        do {
            ++knt;
            for (int64_t k = 0; k < n - 1; ++k) x[k] *= rsafmn;
            beta *= rsafmn;
            a *= rsafmn;
        } while (fabs(beta) < safmin && knt < 20);
        xnorm = nrm2_seq(x, n - 1);
        beta = -copysign(lapy2(a, xnorm), a);
    }
    double tau = (beta - a) / beta;
    double sc = 1.0 / (a - beta);
    for (int64_t k = 0; k < n - 1; ++k) x[k] *= sc;
    for (int j = 0; j < knt; ++j) beta *= safmin;
    *alpha = beta;
    return tau;
}

/*
 * In:  A (r x n, col-major, ld = r) is overwritten (R in the upper triangle of the permuted
 *      matrix, reflectors below), nsteps <= min(r, n) pivot steps are taken.
 * Out: jpvt[0..n-1]  0-based column permutation (jpvt[k] = original index of the column now at
 *      position k), rdiag[k] = R[k,k], gap[k] = (best - second best)/best partial column norm at
 *      the moment pivot k was chosen (degeneracy meter; 1.0 when only one candidate is left),
 *      nrecomp[k] = number of columns whose norm was recomputed after step k.
 */
int omo_qrcp_dlaqp2(double *A, int64_t r, int64_t n, int64_t nsteps, int64_t *jpvt,
                    double *rdiag, double *gap, int64_t *nrecomp)
{
    const double tol3z = sqrt(1.1102230246251565e-16);
    double *vn1 = (double *)malloc(sizeof(double) * (size_t)n);
    double *vn2 = (double *)malloc(sizeof(double) * (size_t)n);
    double *colbuf = (double *)malloc(sizeof(double) * (size_t)r);
    if (!vn1 || !vn2 || !colbuf) return -1;
    for (int64_t j = 0; j < n; ++j) {
        jpvt[j] = j;
        vn1[j] = vn2[j] = nrm2_seq(A + j * r, r);
    }
    int64_t mn = r < n ? r : n;
    if (nsteps > mn) nsteps = mn;
    for (int64_t i = 0; i < nsteps; ++i) {
        /* idamax over vn1[i:], first maximum in the CURRENT (permuted) order */
        int64_t pvt = i;
        double best = vn1[i], second = -1.0;
        for (int64_t j = i + 1; j < n; ++j) {
            if (vn1[j] > best) { second = best; best = vn1[j]; pvt = j; }
            else if (vn1[j] > second) second = vn1[j];
        }
        if (gap) gap[i] = (second < 0.0 || best == 0.0) ? 1.0 : (best - second) / best;
        if (pvt != i) {
            double *ci = A + i * r, *cp = A + pvt * r;
            memcpy(colbuf, ci, sizeof(double) * (size_t)r);
            memcpy(ci, cp, sizeof(double) * (size_t)r);
            memcpy(cp, colbuf, sizeof(double) * (size_t)r);
            int64_t t = jpvt[pvt]; jpvt[pvt] = jpvt[i]; jpvt[i] = t;
            vn1[pvt] = vn1[i];
            vn2[pvt] = vn2[i];
        }
        double *ci = A + i * r;
        double tau = larfg(r - i, ci + i, ci + i + 1);
        if (rdiag) rdiag[i] = ci[i];
        const double *v = ci;           /* v[i] is implicitly 1 */
        int64_t nre = 0;
        for (int64_t j = i + 1; j < n; ++j) {
            double *c = A + j * r;
            if (tau != 0.0) {
                /* dlarf: w = v^T c (dgemv), c -= tau * w * v (dger) */
                double w = c[i];
                for (int64_t k = i + 1; k < r; ++k) w = fma(v[k], c[k], w);
                double tw = tau * w;
                c[i] -= tw;
                for (int64_t k = i + 1; k < r; ++k) c[k] = fma(-tw, v[k], c[k]);
            }
            if (vn1[j] != 0.0) {
                double q = fabs(c[i]) / vn1[j];
                double temp = 1.0 - q * q;
                if (temp < 0.0) temp = 0.0;
                double q2 = vn1[j] / vn2[j];
                double temp2 = temp * (q2 * q2);
                if (temp2 <= tol3z) {
                    if (i < r - 1) {
                        vn1[j] = nrm2_seq(c + i + 1, r - i - 1);
                        vn2[j] = vn1[j];
                    } else {
                        vn1[j] = 0.0;
                        vn2[j] = 0.0;
                    }
                    ++nre;
                } else {
                    vn1[j] *= sqrt(temp);
                }
            }
        }
        if (nrecomp) nrecomp[i] = nre;
    }
    free(vn1); free(vn2); free(colbuf);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Deterministic synthetic snapshots (DESIGN.md "Synthetic workload"); CPU twin of the CUDA    */
/* generator in openmeasure_b200/csrc/synth.cu.  Every operation is a single IEEE rounding.    */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline double u01(uint64_t seed, uint64_t i, uint64_t j)
{
    return (double)(splitmix64(seed ^ (i * 0x9E3779B97F4A7C15ULL + j)) >> 11) * 0x1.0p-53;
}
double omo_synth_u01(uint64_t seed, uint64_t i, uint64_t j) { return u01(seed, i, j); }

/*
 * Fill rows [cell0, cell0+ncell_loc) of every feature of the global F x n_cells x m snapshot
 * matrix into out, laid out as (F * ncell_loc) x m C-order (feature-major, like the reference's
 * X with n_points = ncell_loc).  amp[k] = rho^k (k < K), dec[j] = delta^j (j < m) are supplied by
 * the caller so host libm pow() is evaluated in exactly one place (synth.py).
 */
void omo_synth_fill(double *out, int64_t F, int64_t n_cells, int64_t cell0, int64_t ncell_loc,
                    int64_t m, int64_t K, uint64_t seed, const double *amp, const double *dec,
                    double eps)
{
    double *H = (double *)malloc(sizeof(double) * (size_t)(K * m));
    double *g = (double *)malloc(sizeof(double) * (size_t)K);
    for (int64_t k = 0; k < K; ++k)
        for (int64_t j = 0; j < m; ++j)
            H[k * m + j] = 2.0 * u01(seed + 1, (uint64_t)k, (uint64_t)j) - 1.0;
    for (int64_t f = 0; f < F; ++f) {
        double mu = ldexp(1.0, (int)f) * (1.0 + (double)f / 8.0);
        for (int64_t cl = 0; cl < ncell_loc; ++cl) {
            int64_t c = cell0 + cl;
            double sc = ((double)c + 0.5) / (double)n_cells;
            for (int64_t k = 0; k < K; ++k) {
                double omega = (double)(k + 1) * 0.6180339887498949 + 0.5;
                double theta = u01(seed + 2, (uint64_t)k, (uint64_t)f);
                double t = omega * sc;
                t = t + theta;
                t = t - floor(t);
                double tri = 4.0 * fabs(t - 0.5) - 1.0;
                g[k] = amp[k] * tri;
            }
            uint64_t irow = (uint64_t)(f * n_cells + c);
            double *row = out + (f * ncell_loc + cl) * m;
            for (int64_t j = 0; j < m; ++j) {
                double acc = 0.0;
                for (int64_t k = 0; k < K; ++k) {
                    double p = g[k] * H[k * m + j];
                    acc = acc + p;
                }
                double nz = 2.0 * u01(seed + 3, irow, (uint64_t)j) - 1.0;
                double e = eps * dec[j];
                e = e * nz;
                double v = 0.25 * acc;
                v = v + e;
                v = 1.0 + v;
                row[j] = mu * v;
            }
        }
    }
    free(H); free(g);
}
