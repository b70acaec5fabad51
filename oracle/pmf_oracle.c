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
 *   orc_gauss_sweeps                             gaussian_mf_cavi_bias.py:125-263, gaussian_mf_cavi.py:114-178
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

/* ---- Gaussian MF CAVI (gaussian_mf_cavi_bias.py:125-263; no-bias variant gaussian_mf_cavi.py:114-178) -------------
 * orc_gauss_sweeps follows oracle/pmf_oracle.py::gauss_sweeps (the NumPy restatement pinned on the golden files):
 *   factor pass  S = sum_t (V_oth[c_t] + m_t m_t^T);  V[r] = inv(I/eta2 + S/sigma2)                    :148-158
 *                m[r] = V[r] (sum_t (x_t - b_self[r] - b_oth[c_t]) m_t) / sigma2                        :160-162
 *   bias pass    b[r] = (sum_t (x_t - b_oth[c_t] - m_oth[c_t].m_self[r]) / sigma2) / (1/eta_b2 + n_r/sigma2)  :206-232
 * rows without observations keep their state (:134-135).  np.linalg.inv is LAPACK getrf/getri; here: Gauss-Jordan
 * elimination with partial pivoting (the matrices are SPD with condition numbers <= ~1e4, so the two agree to ~1e-13
 * relative; tests/test_oracle_golden.py pins this port on the reference's own outputs at 1e-10). */
static void invert_kxk(double* A, double* inv, int K) {   /* A is destroyed */
    for (int r = 0; r < K; ++r)
        for (int c = 0; c < K; ++c) inv[r * K + c] = r == c ? 1.0 : 0.0;
    for (int col = 0; col < K; ++col) {
        int piv = col;
        double best = A[col * K + col] < 0 ? -A[col * K + col] : A[col * K + col];
        for (int r = col + 1; r < K; ++r) {
            const double v = A[r * K + col] < 0 ? -A[r * K + col] : A[r * K + col];
            if (v > best) { best = v; piv = r; }
        }
        if (piv != col)
            for (int c = 0; c < K; ++c) {
                double t = A[col * K + c]; A[col * K + c] = A[piv * K + c]; A[piv * K + c] = t;
                t = inv[col * K + c]; inv[col * K + c] = inv[piv * K + c]; inv[piv * K + c] = t;
            }
        const double d = 1.0 / A[col * K + col];
        for (int c = 0; c < K; ++c) { A[col * K + c] *= d; inv[col * K + c] *= d; }
        for (int r = 0; r < K; ++r) {
            if (r == col) continue;
            const double f = A[r * K + col];
            if (f == 0.0) continue;
            for (int c = 0; c < K; ++c) { A[r * K + c] -= f * A[col * K + c]; inv[r * K + c] -= f * inv[col * K + c]; }
        }
    }
}

static void gauss_factor_pass(const int64_t* row_ptr, const int64_t* perm, const int32_t* other, const double* x,
                              int32_t n_rows, int K, double* m_self, double* V_self, const double* m_oth,
                              const double* V_oth, const double* b_self, const double* b_oth, double sigma2, double eta2) {
    const size_t KK = (size_t)K * K;
#pragma omp parallel
    {
        double* S = (double*)malloc(sizeof(double) * (2 * KK + 2 * (size_t)K));
        double* Vn = S + KK;
        double* rhs = Vn + KK;
        double* mn = rhs + K;
#pragma omp for schedule(dynamic, 64)
        for (int32_t r = 0; r < n_rows; ++r) {
            if (row_ptr[r + 1] == row_ptr[r]) continue;
            for (size_t q = 0; q < KK; ++q) S[q] = 0.0;
            for (int k = 0; k < K; ++k) rhs[k] = 0.0;
            for (int64_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
                const int64_t t = perm[p];
                const int32_t c = other[t];
                const double* mo = m_oth + (size_t)c * K;
                const double* Vo = V_oth + (size_t)c * KK;
                const double res = x[t] - b_self[r] - b_oth[c];
                for (int a = 0; a < K; ++a) {
                    for (int b = 0; b < K; ++b) S[a * K + b] += Vo[a * K + b] + mo[a] * mo[b];
                    rhs[a] += mo[a] * res;
                }
            }
            for (int a = 0; a < K; ++a)
                for (int b = 0; b < K; ++b) S[a * K + b] = (a == b ? 1.0 / eta2 : 0.0) + S[a * K + b] / sigma2;
            invert_kxk(S, Vn, K);
            for (int a = 0; a < K; ++a) {
                double acc = 0.0;
                for (int b = 0; b < K; ++b) acc += Vn[a * K + b] * rhs[b];
                mn[a] = (1.0 / sigma2) * acc;
            }
            /* Jacobi inside a pass: row r's own state is read by no other row of this side */
            memcpy(m_self + (size_t)r * K, mn, sizeof(double) * (size_t)K);
            memcpy(V_self + (size_t)r * KK, Vn, sizeof(double) * KK);
        }
        free(S);
    }
}

static void gauss_bias_pass(const int64_t* row_ptr, const int64_t* perm, const int32_t* other, const double* x,
                            int32_t n_rows, int K, const double* m_self, const double* m_oth, double* b_self,
                            const double* b_oth, double sigma2, double eta_b2) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t r = 0; r < n_rows; ++r) {
        const int64_t n = row_ptr[r + 1] - row_ptr[r];
        if (n == 0) continue;
        const double* ms = m_self + (size_t)r * K;
        double sum = 0.0;
        for (int64_t p = row_ptr[r]; p < row_ptr[r + 1]; ++p) {
            const int64_t t = perm[p];
            const int32_t c = other[t];
            const double* mo = m_oth + (size_t)c * K;
            double dot = 0.0;
            for (int k = 0; k < K; ++k) dot += mo[k] * ms[k];
            sum += x[t] - b_oth[c] - dot;
        }
        const double var = 1.0 / (1.0 / eta_b2 + (double)n / sigma2);
        b_self[r] = (var / sigma2) * sum;
    }
}

/* State in/out: m_theta (N,K), m_beta (M,K), V_theta (N,K,K), V_beta (M,K,K), b_user (N), b_item (M). */
void orc_gauss_sweeps(const int32_t* u, const int32_t* i, const double* x, int64_t nnz, int32_t n_users, int32_t n_items,
                      int32_t K, double sigma2, double eta_theta2, double eta_beta2, double eta_bias2, int32_t bias,
                      int32_t sweeps, double* m_theta, double* m_beta, double* V_theta, double* V_beta, double* b_user,
                      double* b_item, int threads, double* sweep_seconds) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int64_t* rp_u = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n_users + 1));
    int64_t* rp_i = (int64_t*)malloc(sizeof(int64_t) * ((size_t)n_items + 1));
    int64_t* pm_u = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    int64_t* pm_i = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz > 0 ? nnz : 1));
    orc_group(u, nnz, n_users, rp_u, pm_u);
    orc_group(i, nnz, n_items, rp_i, pm_i);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int32_t s = 0; s < sweeps; ++s) {
        gauss_factor_pass(rp_u, pm_u, i, x, n_users, K, m_theta, V_theta, m_beta, V_beta, b_user, b_item, sigma2, eta_theta2);
        gauss_factor_pass(rp_i, pm_i, u, x, n_items, K, m_beta, V_beta, m_theta, V_theta, b_item, b_user, sigma2, eta_beta2);
        if (bias) {
            gauss_bias_pass(rp_u, pm_u, i, x, n_users, K, m_theta, m_beta, b_user, b_item, sigma2, eta_bias2);
            gauss_bias_pass(rp_i, pm_i, u, x, n_items, K, m_beta, m_theta, b_item, b_user, sigma2, eta_bias2);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (sweep_seconds) *sweep_seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    free(rp_u); free(rp_i); free(pm_u); free(pm_i);
}
