/*
 * pmf_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY), plain C, float64.
 *
 * A restatement of the reference's Gamma-Poisson CAVI path for sizes where the NumPy row loop of
 * oracle/pmf_oracle.py is too slow, and the CPU baseline timed by bench.py ("port", OpenMP over rows:
 * rows of one pass are independent).  Pinned against the reference through
 * the tests/golden npz files (tests/test_oracle_golden.py runs every entry point against them).
 * Never linked into or called from the product library.
 *
 * Reference lines (paths relative to the reference root):
 *   orc_group        _build_index_lists          poisson_mf_cavi.py:73-84
 *   orc_gamma_pass   user / item row loops       poisson_mf_cavi.py:135-164, :173-194; hpf_cavi.py:126-151, :162-185
 *   orc_poisson_sweeps                           poisson_mf_cavi.py:104-197
 *   orc_hpf_sweeps                               hpf_cavi.py:120-193
 *   orc_predict                                  poisson_mf_cavi.py:221-241
 */
#define _POSIX_C_SOURCE 199309L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RATE_FLOOR 1e-10

/* Stable grouping: perm lists observation indices row by row, original order kept inside a row. */
void orc_group(const int32_t* key, int64_t nnz, int32_t n_rows, int64_t* row_ptr, int64_t* perm) {
    memset(row_ptr, 0, sizeof(int64_t) * ((size_t)n_rows + 1));
    for (int64_t t = 0; t < nnz; ++t) row_ptr[key[t] + 1]++;
    for (int32_t r = 0; r < n_rows; ++r) row_ptr[r + 1] += row_ptr[r];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_rows);
    memcpy(cur, row_ptr, sizeof(int64_t) * (size_t)n_rows);
    for (int64_t t = 0; t < nnz; ++t) perm[cur[key[t]]++] = t;   /* scanning t upward = append order */
    free(cur);
}

/* One Jacobi pass over the rows of one side; sums run over a row's observations in original order. */
void orc_gamma_pass(const int64_t* row_ptr, const int64_t* perm, const int32_t* other, const double* x,
                    int32_t n_rows, int32_t K, const double* E_self, const double* E_oth, double shape_prior,
                    double rate_prior, const double* rate_prior_vec, double* shp, double* rte, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
    {
        double* sa = (double*)malloc(sizeof(double) * 2 * (size_t)K);
        double* sb = sa + K;
#pragma omp for schedule(dynamic, 64)
        for (int32_t r = 0; r < n_rows; ++r) {
            const double rp = rate_prior_vec ? rate_prior_vec[r] : rate_prior;
            const double* own = E_self + (size_t)r * K;
            for (int k = 0; k < K; ++k) sa[k] = sb[k] = 0.0;
            for (int64_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
                const int64_t t = perm[p];
                const double* o = E_oth + (size_t)other[t] * K;
                double rate = 0.0;
#pragma omp simd reduction(+ : rate)
                for (int k = 0; k < K; ++k) rate += o[k] * own[k];
                if (rate < RATE_FLOOR) rate = RATE_FLOOR;
                const double w = x[t] / rate;
#pragma omp simd
                for (int k = 0; k < K; ++k) {
                    sa[k] += w * o[k] * own[k];
                    sb[k] += o[k];
                }
            }
            for (int k = 0; k < K; ++k) {
                shp[(size_t)r * K + k] = shape_prior + sa[k];
                rte[(size_t)r * K + k] = rp + sb[k];
            }
        }
        free(sa);
    }
}

static double now_seconds(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void divide(const double* a, const double* b, double* e, size_t n) {
    for (size_t k = 0; k < n; ++k) e[k] = a[k] / b[k];
}

/* In/out: E_theta (N,K), E_beta (M,K) hold the initial expectations and receive the final ones. */
void orc_poisson_sweeps(const int32_t* u, const int32_t* i, const double* x, int64_t nnz, int32_t N, int32_t M,
                        int32_t K, double a0, double b0, int32_t sweeps, double* E_theta, double* E_beta,
                        double* a_theta, double* b_theta, double* a_beta, double* b_beta, int threads,
                        double* sweep_seconds /* out, may be NULL: time of the sweep loop only */) {
    int64_t* rp_u = (int64_t*)malloc(sizeof(int64_t) * ((size_t)N + 1));
    int64_t* rp_i = (int64_t*)malloc(sizeof(int64_t) * ((size_t)M + 1));
    int64_t* pm_u = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    int64_t* pm_i = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    orc_group(u, nnz, N, rp_u, pm_u);
    orc_group(i, nnz, M, rp_i, pm_i);
    const double t0 = now_seconds();
    for (int32_t s = 0; s < sweeps; ++s) {
        orc_gamma_pass(rp_u, pm_u, i, x, N, K, E_theta, E_beta, a0, b0, NULL, a_theta, b_theta, threads);
        divide(a_theta, b_theta, E_theta, (size_t)N * K);
        orc_gamma_pass(rp_i, pm_i, u, x, M, K, E_beta, E_theta, a0, b0, NULL, a_beta, b_beta, threads);
        divide(a_beta, b_beta, E_beta, (size_t)M * K);
    }
    if (sweep_seconds) *sweep_seconds = now_seconds() - t0;
    free(rp_u); free(rp_i); free(pm_u); free(pm_i);
}

/* HPF: E_xi (N) / E_eta (M) in/out; b_xi / b_eta receive the rates; a_xi, a_eta are the constant shapes. */
void orc_hpf_sweeps(const int32_t* u, const int32_t* i, const double* x, int64_t nnz, int32_t N, int32_t M,
                    int32_t K, double a, double c, double b_prime, double d_prime, double a_xi, double a_eta,
                    int32_t sweeps, double* E_theta, double* E_beta, double* E_xi, double* E_eta, double* a_theta,
                    double* b_theta, double* a_beta, double* b_beta, double* b_xi, double* b_eta, int threads,
                    double* sweep_seconds /* out, may be NULL */) {
    int64_t* rp_u = (int64_t*)malloc(sizeof(int64_t) * ((size_t)N + 1));
    int64_t* rp_i = (int64_t*)malloc(sizeof(int64_t) * ((size_t)M + 1));
    int64_t* pm_u = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    int64_t* pm_i = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    orc_group(u, nnz, N, rp_u, pm_u);
    orc_group(i, nnz, M, rp_i, pm_i);
    const double t0 = now_seconds();
    for (int32_t s = 0; s < sweeps; ++s) {
        orc_gamma_pass(rp_u, pm_u, i, x, N, K, E_theta, E_beta, a, 0.0, E_xi, a_theta, b_theta, threads);
        divide(a_theta, b_theta, E_theta, (size_t)N * K);
        for (int32_t r = 0; r < N; ++r) {
            double sum = 0.0;
            for (int k = 0; k < K; ++k) sum += E_theta[(size_t)r * K + k];
            b_xi[r] = b_prime + sum;
            E_xi[r] = a_xi / b_xi[r];
        }
        orc_gamma_pass(rp_i, pm_i, u, x, M, K, E_beta, E_theta, c, 0.0, E_eta, a_beta, b_beta, threads);
        divide(a_beta, b_beta, E_beta, (size_t)M * K);
        for (int32_t r = 0; r < M; ++r) {
            double sum = 0.0;
            for (int k = 0; k < K; ++k) sum += E_beta[(size_t)r * K + k];
            b_eta[r] = d_prime + sum;
            E_eta[r] = a_eta / b_eta[r];
        }
    }
    if (sweep_seconds) *sweep_seconds = now_seconds() - t0;
    free(rp_u); free(rp_i); free(pm_u); free(pm_i);
}

void orc_predict(const int64_t* users, const int64_t* items, int64_t n, const double* F_user, int32_t N,
                 const double* F_item, int32_t M, int32_t K, double* out) {
    for (int64_t t = 0; t < n; ++t) {
        double acc = 0.0;
        if (users[t] < N && items[t] < M) {
            const double* a = F_user + (size_t)users[t] * K;
            const double* b = F_item + (size_t)items[t] * K;
            for (int k = 0; k < K; ++k) acc += a[k] * b[k];
        }
        out[t] = acc;
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
