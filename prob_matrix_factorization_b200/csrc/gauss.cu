// a5: Gaussian MF CAVI passes (with and without biases).
//
// Reference loops replaced: gaussian_mf_cavi_bias.py:132-165 (users), :170-201 (items), :206-232 and
// :237-263 (biases); gaussian_mf_cavi.py:121-178 (no-bias variant = NULL bias vectors, no bias passes).
//
// Data layout.  Per side: means m[R, ld] (ld = K rounded up to 8, padding zero), covariances as packed
// lower triangles V[R, ldq] (ldq = K(K+1)/2 rounded up to 8; element (i,j), i >= j, at i(i+1)/2 + j) and
// second moments Q = V + m m^T in the same packed layout.  What the other side gathers per rating is
// exactly the reference's E[b b^T] = V + m m^T (:151), so Q is stored ready-made by the row update that
// produces V and m; V itself is output only.  Biases b[R].
//
// Factor pass = two kernels:
//   gauss_accumulate_kernel  one CTA per segment; thread s owns float4 slot s of the concatenated row
//                            [Q (ldq/4 slots) | m (ld/4 slots)] and sums it over the segment's ratings
//                            (Q unweighted, m weighted by the residual x - b_self - b_oth): pure
//                            gather + elementwise accumulate, HBM-bound, 4(ldq+ld)+12 bytes per rating.
//   gauss_solve_kernel       one CTA per row: adds the row's segment sums in order, forms
//                            P = I/eta2 + S/sigma2 in float64 shared memory, inverts it by Cholesky
//                            (np.linalg.inv in the reference, :158), m = V rhs / sigma2, writes m, V, Q.
// Rows without ratings are skipped (state kept, :134-135).
#include "common.cuh"

namespace pmf {

struct GaussArgs {
    const int32_t *seg_row, *seg_start, *seg_order, *row_ptr, *row_seg, *col;
    const int4* seg_desc;
    const float* val;
    int32_t n_seg, n_rows, seg_len, row_offset, K, ld, ldq, nq4, nm4;
    const float *m_oth, *Q_oth, *b_oth, *b_self;
    float *m_self, *V_self, *Q_self;
    float sigma2, eta2;
    float* scratch;  // [n_seg][ldq + ld]
    // sharded form (several GPUs): the segment sums are first added per row into row_sums[n_rows][ldq + ld], summed over
    // the ranks by the caller (NCCL all-reduce: the per-iteration "sufficient statistics combine" of SURVEY.md §8e) and
    // the solve then reads them; counts[n_rows] = ratings of the row over ALL ranks (the skip rule :134-135)
    float* row_sums;
    const int32_t* counts;
};

// row_sums[row] = sum of the row's segment sums, in segment order (one thread per float4 slot)
__global__ void gauss_row_sums_kernel(const GaussArgs a) {
    const int W4 = a.nq4 + a.nm4;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)a.n_rows * W4) return;
    const int row = (int)(e / W4), slot = (int)(e % W4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sg = a.row_seg[row]; sg < a.row_seg[row + 1]; ++sg) {
        if (a.row_ptr[row + 1] == a.row_ptr[row]) break;     // an empty row's single segment was never written
        const float4 v = *reinterpret_cast<const float4*>(a.scratch + (size_t)sg * (a.ldq + a.ld) + 4 * slot);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(a.row_sums + (size_t)row * (a.ldq + a.ld) + 4 * slot) = acc;
}

template <int V>
__global__ void gauss_accumulate_kernel(const GaussArgs a) {
    const int sidx = a.seg_order[blockIdx.x];            // scratch is indexed by segment id, not by processing order
    const int4 d = __ldg(a.seg_desc + blockIdx.x);       // {row, start, end, .}
    const int row = d.x, p0 = d.y, p1 = d.z;
    const int T = blockDim.x;
    const int nslots = a.nq4 + a.nm4;
    const float bs = a.b_self ? a.b_self[a.row_offset + row] : 0.f;
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int p = p0; p < p1; ++p) {
        const int c = __ldg(a.col + p);
        const float res = __ldg(a.val + p) - bs - (a.b_oth ? __ldg(a.b_oth + c) : 0.f);
        const float* qrow = a.Q_oth + (size_t)c * a.ldq;
        const float* mrow = a.m_oth + (size_t)c * a.ld;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int s = threadIdx.x + v * T;
            if (s < nslots) {
                const bool isq = s < a.nq4;
                const float4 r = isq ? ldg_f4(qrow + 4 * s) : ldg_f4(mrow + 4 * (s - a.nq4));
                const float w = isq ? 1.f : res;
                acc[v].x = fmaf(w, r.x, acc[v].x);
                acc[v].y = fmaf(w, r.y, acc[v].y);
                acc[v].z = fmaf(w, r.z, acc[v].z);
                acc[v].w = fmaf(w, r.w, acc[v].w);
            }
        }
    }
    float* dst = a.scratch + (size_t)sidx * (a.ldq + a.ld);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int s = threadIdx.x + v * T;
        if (s < nslots) *reinterpret_cast<float4*>(dst + 4 * s) = acc[v];
    }
}

// packed lower-triangle index e -> (i, j), j <= i
__device__ __forceinline__ void unpack_tri(int e, int& i, int& j) {
    i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= e) ++i;
    while (i * (i + 1) / 2 > e) --i;
    j = e - i * (i + 1) / 2;
}

// 1/sqrt(d) in float64: float estimate + two Newton steps (relative error ~1e-16; the results are stored as float32)
__device__ __forceinline__ double rsqrt64(double d) {
    double y = (double)rsqrtf((float)d);
    y = y * (1.5 - 0.5 * d * y * y);
    y = y * (1.5 - 0.5 * d * y * y);
    return y;
}

// One WARP per row (several rows per CTA, no block-wide barriers).  Shared, per warp, float64:
// A[K][K+1] (P, then L below / L^-1 above the diagonal, then V below), rinv[K] (1/L_kk, later diag V), rhs[K], mean[K].
__global__ void __launch_bounds__(256) gauss_solve_kernel(const GaussArgs a) {
    extern __shared__ double sm[];
    const int K = a.K, LD = K + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= a.n_rows) return;
    if (a.counts ? a.counts[row] == 0 : a.row_ptr[row + 1] == a.row_ptr[row]) return;  // no ratings: keep the state (:134-135)
    double* A = sm + (size_t)warp * ((size_t)K * LD + 3 * K);
    double* rinv = A + (size_t)K * LD;
    double* rhs = rinv + K;
    double* mean = rhs + K;
    const int W = a.ldq + a.ld;
    const int s0 = a.row_sums ? row : a.row_seg[row], s1 = a.row_sums ? row + 1 : a.row_seg[row + 1];
    const float* sums = a.row_sums ? a.row_sums : a.scratch;
    const int npk = K * (K + 1) / 2;
    const double inv_sigma2 = 1.0 / (double)a.sigma2, inv_eta2 = 1.0 / (double)a.eta2;
    // P = I/eta2 + S/sigma2 (lower triangle), rhs = sum res*m; the segments' sums are added in segment order
    for (int e = lane; e < npk + K; e += 32) {
        const int off = e < npk ? e : a.ldq + (e - npk);
        double s = 0.0;
        for (int sg = s0; sg < s1; ++sg) s += (double)sums[(size_t)sg * W + off];
        if (e < npk) {
            int i, j;
            unpack_tri(e, i, j);
            A[i * LD + j] = s * inv_sigma2 + (i == j ? inv_eta2 : 0.0);
        } else {
            rhs[e - npk] = s;
        }
    }
    __syncwarp();
    // Cholesky P = L L^T, left-looking by columns: every lane recomputes the pivot (k multiply-adds, no exchange),
    // lane i finishes L[i][k] for its rows i > k.  L[k][j], j < k, was completed in earlier columns.
    for (int k = 0; k < K; ++k) {
        double d = A[k * LD + k];
        for (int j = 0; j < k; ++j) d -= A[k * LD + j] * A[k * LD + j];
        const double rk = rsqrt64(d);
        for (int i = k + 1 + lane; i < K; i += 32) {
            double t = A[i * LD + k];
            for (int j = 0; j < k; ++j) t -= A[i * LD + j] * A[k * LD + j];
            A[i * LD + k] = t * rk;
        }
        __syncwarp();              // everybody has read the old A[k][k]
        if (lane == 0) { A[k * LD + k] = d * rk; rinv[k] = rk; }
        __syncwarp();
    }
    // L^-1 column by column (lane c owns column c): x_c = 1/L_cc, x_i = -(sum_{c<=j<i} L_ij x_j)/L_ii, stored in row c of
    // the UPPER triangle (A[c][i] = Linv[i][c], i > c); the diagonal follows after everybody is done reading L.
    for (int c = lane; c < K; c += 32) {
        double* x = A + (size_t)c * LD;
        const double xc = rinv[c];
        for (int i = c + 1; i < K; ++i) {
            double s = A[i * LD + c] * xc;
            for (int j = c + 1; j < i; ++j) s += A[i * LD + j] * x[j];
            x[i] = -s * rinv[i];
        }
        mean[c] = xc;
    }
    __syncwarp();
    for (int c = lane; c < K; c += 32) A[c * LD + c] = mean[c];
    __syncwarp();
    // V = Linv^T Linv: V[i][j] = sum_{k >= i} Linv[k][i] Linv[k][j] (j <= i) = sum_k A[i][k] A[j][k]; strictly-lower entries
    // overwrite L (no longer needed), the diagonal goes to rinv[]
    for (int e = lane; e < npk; e += 32) {
        int i, j;
        unpack_tri(e, i, j);
        double s = 0.0;
        for (int k = i; k < K; ++k) s += A[i * LD + k] * A[j * LD + k];
        if (i != j) A[i * LD + j] = s; else rinv[i] = s;
    }
    __syncwarp();
    // m = V rhs / sigma2
    const size_t R = (size_t)(a.row_offset + row);
    for (int i = lane; i < K; i += 32) {
        double s = rinv[i] * rhs[i];
        for (int j = 0; j < K; ++j) {
            if (j == i) continue;
            s += (j < i ? A[i * LD + j] : A[j * LD + i]) * rhs[j];
        }
        const double mi = s * inv_sigma2;
        a.m_self[R * a.ld + i] = (float)mi;
        mean[i] = mi;
    }
    __syncwarp();
    for (int e = lane; e < npk; e += 32) {
        int i, j;
        unpack_tri(e, i, j);
        const double v = (i == j) ? rinv[i] : A[i * LD + j];
        a.V_self[R * a.ldq + e] = (float)v;
        a.Q_self[R * a.ldq + e] = (float)(v + mean[i] * mean[j]);   // E[th th^T], :151 / :187
    }
}

// Bias pass (gaussian_mf_cavi_bias.py:206-263), two kernels so that a row with thousands of ratings is not walked by one
// warp: (1) one warp per SEGMENT (<= seg_len ratings), 8 lanes per rating compute <m_self[row], m_oth[col]> and the
// segment's residual sum in float64; (2) one thread per row adds its segments' sums in segment order (deterministic)
// and applies the update.
struct BiasArgs {
    const int32_t *row_ptr, *col, *seg_row, *seg_start, *row_seg;
    const float* val;
    int32_t n_rows, n_seg, seg_len, row_offset, ld, nvec;
    const float *m_self, *m_oth, *b_oth;
    float* b_self;
    float sigma2, eta_b2;
    double* partial;   // [n_seg]
    double* row_resid;         // sharded form: per-row residual sums (summed over the ranks by the caller) ...
    const int32_t* counts;     // ... and the rows' rating counts over all ranks
    int phase;                 // 0 whole pass, 1 up to the row sums, 2 from the (all-reduced) row sums
};

__global__ void __launch_bounds__(256) gauss_bias_partial_kernel(const BiasArgs a) {
    const int lane = threadIdx.x & 31, gl = lane & 7, grp = lane >> 3;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= a.n_seg) return;
    const int row = a.seg_row[wid];
    const int p0 = a.seg_start[wid];
    const int p1 = min(p0 + a.seg_len, a.row_ptr[row + 1]);
    const float* own = a.m_self + (size_t)(a.row_offset + row) * a.ld;
    double acc = 0.0;
    for (int base = p0; base < p1; base += 4) {
        const int p = base + grp;
        const bool ok = p < p1;
        const int c = ok ? __ldg(a.col + p) : 0;
        double d = 0.0;
        if (ok) {
            const float* oth = a.m_oth + (size_t)c * a.ld;
            for (int idx = gl; idx < a.nvec; idx += 8) {
                const float4 x = ldg_f4(own + idx * 4), y = ldg_f4(oth + idx * 4);
                d += (double)x.x * y.x + (double)x.y * y.y + (double)x.z * y.z + (double)x.w * y.w;
            }
        }
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        if (ok && gl == 0) acc += (double)__ldg(a.val + p) - (double)__ldg(a.b_oth + c) - d;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (lane == 0) a.partial[wid] = acc;
}

__global__ void __launch_bounds__(256) gauss_bias_finish_kernel(const BiasArgs a) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.n_rows) return;
    const int n_local = a.row_ptr ? a.row_ptr[row + 1] - a.row_ptr[row] : 0;
    double acc = 0.0;
    if (a.phase != 2) {
        if (n_local > 0)
            for (int sg = a.row_seg[row]; sg < a.row_seg[row + 1]; ++sg) acc += a.partial[sg];
        if (a.phase == 1) { a.row_resid[row] = acc; return; }
    } else {
        acc = a.row_resid[row];
    }
    const int n = a.counts ? a.counts[row] : n_local;
    if (n == 0) return;  // gaussian_mf_cavi_bias.py:208-209
    const double prec = 1.0 / (double)a.eta_b2 + (double)n / (double)a.sigma2;          // :226
    a.b_self[a.row_offset + row] = (float)((1.0 / prec) / (double)a.sigma2 * acc);     // :230
}

}  // namespace pmf

using namespace pmf;

extern "C" {

int pmf_gauss_packed_stride(int K) { return K <= 0 ? 0 : ((K * (K + 1) / 2 + 7) / 8) * 8; }

int64_t pmf_gauss_workspace_bytes(const pmf_csr* csr, int32_t K) {
    if (!csr || K <= 0) return -1;
    const CsrView c = csr_view(csr);
    const int64_t w = (int64_t)pmf_gauss_packed_stride(K) + pmf_row_stride(K);
    const int64_t b = (int64_t)c.n_seg * w * (int64_t)sizeof(float);
    return b > 0 ? b : 16;
}

static int gauss_factor_impl(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_Q_oth,
                             const float* d_b_oth, float* d_m_self, float* d_V_self, float* d_Q_self,
                             const float* d_b_self, float sigma2, float eta2, void* d_workspace, float* d_row_sums,
                             const int32_t* d_counts, int phase, void* stream);

int pmf_gauss_factor_pass(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_Q_oth,
                          const float* d_b_oth, float* d_m_self, float* d_V_self, float* d_Q_self,
                          const float* d_b_self, float sigma2, float eta2, void* d_workspace, void* stream) {
    return gauss_factor_impl(csr, K, d_m_oth, d_Q_oth, d_b_oth, d_m_self, d_V_self, d_Q_self, d_b_self, sigma2, eta2,
                             d_workspace, nullptr, nullptr, 0, stream);
}

int pmf_gauss_factor_pass_sharded(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_Q_oth,
                                  const float* d_b_oth, float* d_m_self, float* d_V_self, float* d_Q_self,
                                  const float* d_b_self, float sigma2, float eta2, void* d_workspace, float* d_row_sums,
                                  const int32_t* d_counts, int32_t phase, void* stream) {
    PMF_REQUIRE(d_row_sums != nullptr && d_counts != nullptr && (phase == 1 || phase == 2), "bad sharded-pass arguments");
    return gauss_factor_impl(csr, K, d_m_oth, d_Q_oth, d_b_oth, d_m_self, d_V_self, d_Q_self, d_b_self, sigma2, eta2,
                             d_workspace, d_row_sums, d_counts, phase, stream);
}

static int gauss_factor_impl(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_Q_oth,
                             const float* d_b_oth, float* d_m_self, float* d_V_self, float* d_Q_self,
                             const float* d_b_self, float sigma2, float eta2, void* d_workspace, float* d_row_sums,
                             const int32_t* d_counts, int phase, void* stream) {
    PMF_REQUIRE(csr != nullptr, "csr is NULL");
    PMF_REQUIRE(K >= 1 && K <= 96, "K=%d outside [1, 96] for the Gaussian model", K);
    PMF_REQUIRE(d_m_oth && d_Q_oth && d_m_self && d_V_self && d_Q_self && d_workspace, "NULL table");
    PMF_REQUIRE((d_b_oth == nullptr) == (d_b_self == nullptr), "bias vectors go together");
    PMF_REQUIRE(sigma2 > 0.f && eta2 > 0.f, "variances must be positive");
    const CsrView c = csr_view(csr);
    GaussArgs a;
    a.seg_row = c.seg_row; a.seg_start = c.seg_start; a.seg_order = c.seg_order; a.row_ptr = c.row_ptr;
    a.row_seg = c.row_seg; a.col = c.col; a.val = c.val; a.seg_desc = c.seg_desc;
    a.n_seg = c.n_seg; a.n_rows = c.n_rows; a.seg_len = c.seg_len; a.row_offset = c.row_offset;
    a.K = K; a.ld = pmf_row_stride(K); a.ldq = pmf_gauss_packed_stride(K); a.nq4 = a.ldq / 4; a.nm4 = a.ld / 4;
    a.m_oth = d_m_oth; a.Q_oth = d_Q_oth; a.b_oth = d_b_oth; a.b_self = d_b_self;
    a.m_self = d_m_self; a.V_self = d_V_self; a.Q_self = d_Q_self;
    a.sigma2 = sigma2; a.eta2 = eta2; a.scratch = (float*)d_workspace;
    a.row_sums = d_row_sums; a.counts = d_counts;
    cudaStream_t s = (cudaStream_t)stream;
    if (c.n_seg > 0 && phase != 2) {
        const int nslots = a.nq4 + a.nm4;
        // threads per segment: a multiple of 32 covering the slots with at most 4 per thread
        int V = 1;
        while (V < 4 && (nslots + V - 1) / V > 256) V *= 2;
        int T = (((nslots + V - 1) / V) + 31) / 32 * 32;
        if (T > 1024) { set_error("K=%d needs %d slots: too wide", K, nslots); return PMF_EUNSUPPORTED; }
        if (V == 1) gauss_accumulate_kernel<1><<<c.n_seg, T, 0, s>>>(a);
        else if (V == 2) gauss_accumulate_kernel<2><<<c.n_seg, T, 0, s>>>(a);
        else gauss_accumulate_kernel<4><<<c.n_seg, T, 0, s>>>(a);
        PMF_LAUNCH_CHECK();
    }
    if (phase == 1) {
        if (c.n_rows > 0) {
            gauss_row_sums_kernel<<<(unsigned)cdiv((int64_t)c.n_rows * (a.nq4 + a.nm4), 256), 256, 0, s>>>(a);
            PMF_LAUNCH_CHECK();
        }
        return PMF_OK;
    }
    // (a one-thread-per-row register Cholesky for K <= 12 was tried and dropped: 252 registers and a ~3000-long dependent
    // float64 chain per thread made the C1 sweep 0.31 ms instead of 0.21 ms, profiles/README.md)
    if (c.n_rows > 0) {
        const size_t per_warp = ((size_t)K * (K + 1) + 3 * (size_t)K) * sizeof(double);
        int warps = (int)((size_t)(96 * 1024) / per_warp);       // rows per CTA: as many as fit in 96 KB, at most 8
        warps = warps > 8 ? 8 : (warps < 1 ? 1 : warps);
        const size_t smem = per_warp * warps;
        if (smem > 48 * 1024)
            PMF_CUDA(cudaFuncSetAttribute(gauss_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gauss_solve_kernel<<<(unsigned)cdiv(c.n_rows, warps), 32 * warps, smem, s>>>(a);
        PMF_LAUNCH_CHECK();
    }
    return PMF_OK;
}

static int gauss_bias_impl(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_m_self, const float* d_b_oth,
                           float* d_b_self, float sigma2, float eta_b2, void* d_workspace, double* d_row_resid,
                           const int32_t* d_counts, int phase, void* stream);

int pmf_gauss_bias_pass(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_m_self,
                        const float* d_b_oth, float* d_b_self, float sigma2, float eta_b2, void* d_workspace, void* stream) {
    return gauss_bias_impl(csr, K, d_m_oth, d_m_self, d_b_oth, d_b_self, sigma2, eta_b2, d_workspace, nullptr, nullptr, 0, stream);
}

int pmf_gauss_bias_pass_sharded(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_m_self,
                                const float* d_b_oth, float* d_b_self, float sigma2, float eta_b2, void* d_workspace,
                                double* d_row_resid, const int32_t* d_counts, int32_t phase, void* stream) {
    PMF_REQUIRE(d_row_resid != nullptr && d_counts != nullptr && (phase == 1 || phase == 2), "bad sharded-pass arguments");
    return gauss_bias_impl(csr, K, d_m_oth, d_m_self, d_b_oth, d_b_self, sigma2, eta_b2, d_workspace, d_row_resid, d_counts, phase, stream);
}

static int gauss_bias_impl(const pmf_csr* csr, int32_t K, const float* d_m_oth, const float* d_m_self, const float* d_b_oth,
                           float* d_b_self, float sigma2, float eta_b2, void* d_workspace, double* d_row_resid,
                           const int32_t* d_counts, int phase, void* stream) {
    PMF_REQUIRE(csr != nullptr, "csr is NULL");
    PMF_REQUIRE(K >= 1, "K must be positive");
    PMF_REQUIRE(d_m_oth && d_m_self && d_b_oth && d_b_self && d_workspace, "NULL table");
    PMF_REQUIRE(((uintptr_t)d_workspace & 7) == 0, "workspace must be 8-byte aligned");
    PMF_REQUIRE(sigma2 > 0.f && eta_b2 > 0.f, "variances must be positive");
    const CsrView c = csr_view(csr);
    BiasArgs a;
    a.row_ptr = c.row_ptr; a.col = c.col; a.val = c.val; a.seg_row = c.seg_row; a.seg_start = c.seg_start; a.row_seg = c.row_seg;
    a.n_rows = c.n_rows; a.n_seg = c.n_seg; a.seg_len = c.seg_len; a.row_offset = c.row_offset;
    a.ld = pmf_row_stride(K); a.nvec = a.ld / 4;
    a.m_self = d_m_self; a.m_oth = d_m_oth; a.b_oth = d_b_oth; a.b_self = d_b_self;
    a.sigma2 = sigma2; a.eta_b2 = eta_b2;
    a.partial = (double*)d_workspace;   // n_seg doubles <= pmf_gauss_workspace_bytes (>= 8 floats per segment)
    a.row_resid = d_row_resid; a.counts = d_counts; a.phase = phase;
    cudaStream_t s = (cudaStream_t)stream;
    if (c.n_seg > 0 && phase != 2) {
        gauss_bias_partial_kernel<<<(unsigned)cdiv((int64_t)c.n_seg * 32, 256), 256, 0, s>>>(a);
        PMF_LAUNCH_CHECK();
    }
    if (c.n_rows > 0) {
        gauss_bias_finish_kernel<<<(unsigned)cdiv(c.n_rows, 256), 256, 0, s>>>(a);
        PMF_LAUNCH_CHECK();
    }
    return PMF_OK;
}

}  // extern "C"
