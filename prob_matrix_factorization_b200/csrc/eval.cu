// a8-a10: gather-dot prediction and fused evaluation statistics.
//
// predict: poisson_mf_cavi.py:221-241, hpf_cavi.py:215-231, gaussian_mf_cavi_bias.py:291-316,
//          hpf_pytorch.py:66-69,186-195.
// stats:   evaluate_rmse / evaluate_macro_mae (poisson_mf_cavi.py:243-251,
//          gaussian_mf_cavi_bias.py:318-347) with metrics.py:6-16, :37-51, :53-66 fused behind the
//          prediction so a validation pass costs one kernel and one small D2H.
//
// HBM-bound gathers: 8 lanes per (user,item) pair, 128-bit loads, float64 accumulation (cheap; the
// factor rows are float32).
#include <stddef.h>

#include "common.cuh"

namespace pmf {

constexpr int kEvalGroup = 8;
constexpr int kMaxLabels = 64;

__device__ __forceinline__ float softplus_f(float z) { return z > 20.f ? z : log1pf(expf(z)); }

struct PredictArgs {
    const int32_t *users, *items;
    int64_t n;
    const float *F_user, *F_item, *b_user, *b_item;
    int32_t n_users, n_items, K, ld, nvec, softplus;
    float global_mean;
};

// Returns the prediction WITHOUT global_mean; *valid tells whether both ids are known.
__device__ __forceinline__ double pair_predict(const PredictArgs& a, int64_t t, int gl, bool* valid) {
    const int32_t u = a.users[t], i = a.items[t];
    const bool ok = (u >= 0) && (i >= 0) && (u < a.n_users) && (i < a.n_items);
    *valid = ok;
    double acc = 0.0;
    if (ok) {
        const float* fu = a.F_user + (size_t)u * a.ld;
        const float* fi = a.F_item + (size_t)i * a.ld;
        for (int idx = gl; idx < a.nvec; idx += kEvalGroup) {
            float4 x = ldg_f4(fu + idx * 4), y = ldg_f4(fi + idx * 4);
            if (a.softplus) {
                const int k0 = idx * 4;  // padding columns hold 0 -> softplus(0) != 0, so mask them
                x.x = k0 + 0 < a.K ? softplus_f(x.x) : 0.f;  y.x = softplus_f(y.x);
                x.y = k0 + 1 < a.K ? softplus_f(x.y) : 0.f;  y.y = softplus_f(y.y);
                x.z = k0 + 2 < a.K ? softplus_f(x.z) : 0.f;  y.z = softplus_f(y.z);
                x.w = k0 + 3 < a.K ? softplus_f(x.w) : 0.f;  y.w = softplus_f(y.w);
            }
            acc += (double)x.x * y.x + (double)x.y * y.y + (double)x.z * y.z + (double)x.w * y.w;
        }
    }
#pragma unroll
    for (int o = kEvalGroup / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (ok && a.b_user) acc = (double)a.b_user[u] + (double)a.b_item[i] + acc;
    return acc;
}

__global__ void __launch_bounds__(256) predict_kernel(const PredictArgs a, double* __restrict__ pred) {
    const int gl = threadIdx.x & (kEvalGroup - 1);
    const int64_t groups = (int64_t)gridDim.x * blockDim.x / kEvalGroup;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kEvalGroup;
    // warp-uniform trip count (full-mask shuffles inside pair_predict)
    const int64_t iters = (a.n + groups - 1) / groups;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t t = g0 + it * groups;
        const bool in = t < a.n;
        bool valid = false;
        const double p = pair_predict(a, in ? t : 0, gl, &valid);
        if (in && gl == 0) pred[t] = (valid ? p : 0.0) + (double)a.global_mean;
    }
}

// Scratch of pmf_eval_stats: per-block partial sums of the four scalars, fixed-point per-label |error| sums, a counter.
constexpr int kEvalMaxBlocks = kNumSMs * 8;
struct EvalScratch {
    double part[kEvalMaxBlocks][4];
    unsigned long long lab_abs[kMaxLabels];   // sum |y - p| * 2^32, rounded per term
    unsigned long long lab_cnt[kMaxLabels];
    unsigned int done;
};
constexpr double kLabScale = 4294967296.0;    // 2^32

// DETERMINISTIC reduction (the early-stopping rule compares RMSEs of consecutive sweeps against a tolerance, on the host or
// -- pmf_loop_decide -- on the device, and sharded fits must take the same decision on every rank): the element -> thread
// mapping is fixed, warp and block sums use fixed trees, every block parks its four sums in scratch and the LAST block to
// finish adds them in block order; per-label sums are integer atomics (counts, and |error| in 2^-32 fixed point), whose
// result does not depend on the order of arrival.
__global__ void __launch_bounds__(256) eval_stats_kernel(const PredictArgs a, const float* __restrict__ y,
                                                         const int32_t* __restrict__ label, int n_labels,
                                                         int drop_invalid, double* __restrict__ out,
                                                         EvalScratch* __restrict__ sc) {
    __shared__ double s_red[4][8];
    __shared__ bool s_last;
    const int gl = threadIdx.x & (kEvalGroup - 1);
    const int64_t groups = (int64_t)gridDim.x * blockDim.x / kEvalGroup;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kEvalGroup;
    const int64_t iters = (a.n + groups - 1) / groups;
    double cnt = 0, sse = 0, sae = 0, lpl = 0;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t t = g0 + it * groups;
        const bool in = t < a.n;
        bool valid = false;
        double p = pair_predict(a, in ? t : 0, gl, &valid);
        if (in && gl == 0 && (valid || !drop_invalid)) {
            if (!valid) p = 0.0;
            const double yt = (double)y[t];
            const double e = yt - p;  // global_mean cancels: y_true + mean - (pred + mean)
            cnt += 1.0;
            sse += e * e;
            sae += fabs(e);
            lpl += yt * log(fmax(p, 1e-10)) - p - lgamma(yt + 1.0);  // metrics.py:53-66
            if (label) {
                const int lb = label[t];
                if (lb >= 0 && lb < n_labels) {
                    atomicAdd(&sc->lab_abs[lb], (unsigned long long)llrint(fmin(fabs(e), 1048576.0) * kLabScale));
                    atomicAdd(&sc->lab_cnt[lb], 1ull);
                }
            }
        }
    }
    // block reduction of the four scalars (fixed shuffle tree, then warps in order)
    double v[4] = {cnt, sse, sae, lpl};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if (lane == 0) s_red[k][warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
        sc->part[blockIdx.x][threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&sc->done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < 4) {
        double t = 0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += sc->part[b][threadIdx.x];
        out[threadIdx.x] = t;
    }
    for (int k = threadIdx.x; k < n_labels; k += blockDim.x) {
        out[4 + k] = (double)sc->lab_abs[k] / kLabScale;
        out[4 + n_labels + k] = (double)sc->lab_cnt[k];
    }
    if (threadIdx.x == 0) sc->done = 0;
}

static int fill_args(PredictArgs& a, const int32_t* d_users, const int32_t* d_items, int64_t n, const float* d_F_user,
                     int32_t n_users, const float* d_F_item, int32_t n_items, int32_t K, int32_t ld,
                     const float* d_b_user, const float* d_b_item, float global_mean, int32_t softplus) {
    PMF_REQUIRE(n >= 0, "n < 0");
    PMF_REQUIRE(n == 0 || (d_users && d_items), "id arrays are NULL");
    PMF_REQUIRE(d_F_user && d_F_item, "factor tables are NULL");
    PMF_REQUIRE(K >= 1 && ld >= K && ld % 8 == 0, "need 1 <= K <= ld and ld %% 8 == 0 (K=%d ld=%d)", K, ld);
    PMF_REQUIRE((d_b_user == nullptr) == (d_b_item == nullptr), "bias vectors go together");
    a.users = d_users; a.items = d_items; a.n = n; a.F_user = d_F_user; a.F_item = d_F_item;
    a.b_user = d_b_user; a.b_item = d_b_item; a.n_users = n_users; a.n_items = n_items;
    a.K = K; a.ld = ld; a.nvec = ld / 4; a.softplus = softplus; a.global_mean = global_mean;
    return PMF_OK;
}

static unsigned eval_grid(int64_t n) {
    const int64_t per_block = 256 / kEvalGroup;
    int64_t blocks = cdiv(n > 0 ? n : 1, per_block);
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (unsigned)(blocks < cap ? blocks : cap);   // <= kEvalMaxBlocks
}

__global__ void scale_rows_kernel(const float* __restrict__ F, const float* __restrict__ scale, int64_t n4, int ld4,
                                  float* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n4) return;
    const float sc = scale[e / ld4];
    float4 v = reinterpret_cast<const float4*>(F)[e];
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    reinterpret_cast<float4*>(out)[e] = v;
}

}  // namespace pmf

using namespace pmf;

extern "C" {

int pmf_predict(const int32_t* d_users, const int32_t* d_items, int64_t n, const float* d_F_user, int32_t n_users,
                const float* d_F_item, int32_t n_items, int32_t K, int32_t ld, const float* d_b_user,
                const float* d_b_item, float global_mean, int32_t softplus, double* d_pred, void* stream) {
    PredictArgs a;
    PMF_TRY(fill_args(a, d_users, d_items, n, d_F_user, n_users, d_F_item, n_items, K, ld, d_b_user, d_b_item,
                      global_mean, softplus));
    if (n == 0) return PMF_OK;
    PMF_REQUIRE(d_pred != nullptr, "d_pred is NULL");
    predict_kernel<<<eval_grid(n), 256, 0, (cudaStream_t)stream>>>(a, d_pred);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

int64_t pmf_eval_stats_scratch_bytes(void) { return (int64_t)sizeof(EvalScratch); }

int pmf_eval_stats(const int32_t* d_users, const int32_t* d_items, const float* d_y, const int32_t* d_label,
                   int32_t n_labels, int64_t n, const float* d_F_user, int32_t n_users, const float* d_F_item,
                   int32_t n_items, int32_t K, int32_t ld, const float* d_b_user, const float* d_b_item,
                   float global_mean, int32_t drop_invalid, double* d_out, void* d_scratch, void* stream) {
    PredictArgs a;
    PMF_TRY(fill_args(a, d_users, d_items, n, d_F_user, n_users, d_F_item, n_items, K, ld, d_b_user, d_b_item,
                      global_mean, 0));
    PMF_REQUIRE(d_out != nullptr, "d_out is NULL");
    PMF_REQUIRE(n_labels >= 0 && n_labels <= kMaxLabels, "n_labels=%d exceeds %d", n_labels, kMaxLabels);
    PMF_REQUIRE(n == 0 || d_y, "d_y is NULL");
    PMF_REQUIRE(n == 0 || (d_scratch != nullptr && ((uintptr_t)d_scratch & 7) == 0), "d_scratch is NULL or misaligned");
    cudaStream_t s = (cudaStream_t)stream;
    PMF_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (4 + 2 * (size_t)n_labels), s));
    if (n == 0) return PMF_OK;
    EvalScratch* sc = (EvalScratch*)d_scratch;
    PMF_CUDA(cudaMemsetAsync(sc->lab_abs, 0, sizeof(EvalScratch) - offsetof(EvalScratch, lab_abs), s));
    eval_stats_kernel<<<eval_grid(n), 256, 0, s>>>(a, d_y, n_labels > 0 ? d_label : nullptr, n_labels, drop_invalid, d_out, sc);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

int pmf_scale_rows(const float* d_F, const float* d_scale, int64_t rows, int32_t ld, float* d_out, void* stream) {
    PMF_REQUIRE(rows >= 0 && ld > 0 && ld % 4 == 0, "bad shape");
    if (rows == 0) return PMF_OK;
    PMF_REQUIRE(d_F && d_scale && d_out, "NULL argument");
    const int64_t n4 = rows * (ld / 4);
    scale_rows_kernel<<<(unsigned)cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>(d_F, d_scale, n4, ld / 4, d_out);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

}  // extern "C"
