// a6/a7: gradient-based HPF ("HPF_PyTorch"): fused loss + analytic backward, fused Adam, predict.
//
// Reference replaced: HPF_PyTorch.loss / forward / predict (hpf_pytorch.py:66-69, :71-184, :186-195) and the
// autograd backward + torch.optim.Adam step of the training loops (compare_models.py:305-313,
// train_hpf_pytorch_full.py:99-108).  The reference evaluates softplus over the WHOLE theta/beta tables
// several times per mini-batch and back-propagates dense zero-filled gradients; here softplus, the loss
// terms and the gradient of every term are evaluated only on the rows the batch touches, in one kernel,
// and duplicate ids inside a batch are combined with float atomics (RED.ADD.F32).
//
// Parameters keep the reference's shapes: theta_raw (N,K), beta_raw (M,K) row-major with row stride K
// (no padding -- they are torch nn.Parameters), xi_raw (N), eta_raw (M); everything float32.
#include "common.cuh"

namespace pmf {

__device__ __forceinline__ float softplus_t(float z) { return z > 20.f ? z : log1pf(expf(z)); }        // F.softplus, threshold 20
__device__ __forceinline__ float softplus_grad_t(float z) { return z > 20.f ? 1.f : 1.f / (1.f + expf(-z)); }

struct MapArgs {
    const void *users, *items;   // int64 or int32
    const float* ratings;
    int64_t B;
    const float *theta, *beta, *xi, *eta, *user_scale, *item_scale;
    int32_t N, M, K;
    float a, a_prime, b_prime, c, c_prime, d_prime;
    float *g_theta, *g_beta, *g_xi, *g_eta;
    double* loss;
    int32_t* bad;
};

// One group of G lanes per batch element; lane l owns factors k = l, l+G, ...
template <int G, typename IdT>
__global__ void __launch_bounds__(256) hpf_map_loss_grad_kernel(const MapArgs a) {
    constexpr int MAXV = 8;   // K <= G * MAXV
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    double loss = 0.0;
    if (gid < a.B) {
        const int64_t u = (int64_t)((const IdT*)a.users)[gid], it = (int64_t)((const IdT*)a.items)[gid];
        if (u < 0 || u >= a.N || it < 0 || it >= a.M) {
            if (gl == 0) atomicOr(a.bad, 1);   // torch indexing would raise IndexError
        } else {
            const float r = a.ratings[gid];
            const float s = a.user_scale[u], t = a.item_scale[it];
            const float* tr = a.theta + (size_t)u * a.K;
            const float* br = a.beta + (size_t)it * a.K;
            float th[MAXV], be[MAXV], traw[MAXV], braw[MAXV];
            float dot = 0.f, sum_th = 0.f, sum_be = 0.f, sum_lth = 0.f, sum_lbe = 0.f;
#pragma unroll
            for (int v = 0; v < MAXV; ++v) {
                const int k = gl + v * G;
                if (k < a.K) {
                    traw[v] = tr[k]; braw[v] = br[k];
                    th[v] = softplus_t(traw[v]); be[v] = softplus_t(braw[v]);
                    dot = fmaf(th[v], be[v], dot);
                    sum_th += th[v]; sum_be += be[v];
                    sum_lth += logf(th[v]); sum_lbe += logf(be[v]);
                }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                dot += __shfl_xor_sync(gmask, dot, o);
                sum_th += __shfl_xor_sync(gmask, sum_th, o);
                sum_be += __shfl_xor_sync(gmask, sum_be, o);
                sum_lth += __shfl_xor_sync(gmask, sum_lth, o);
                sum_lbe += __shfl_xor_sync(gmask, sum_lbe, o);
            }
            const float xr = a.xi[u], er = a.eta[it];
            const float xi = softplus_t(xr), eta = softplus_t(er);
            const float lam = fmaxf(dot, 1e-6f);                         // hpf_pytorch.py:80
            const float g = dot >= 1e-6f ? 1.f - r / lam : 0.f;          // clamp passes gradient only inside its range
#pragma unroll
            for (int v = 0; v < MAXV; ++v) {
                const int k = gl + v * G;
                if (k < a.K) {
                    const float dth = g * be[v] + s * (xi - (a.a - 1.f) / th[v]);
                    const float dbe = g * th[v] + t * (eta - (a.c - 1.f) / be[v]);
                    atomicAdd(a.g_theta + (size_t)u * a.K + k, dth * softplus_grad_t(traw[v]));
                    atomicAdd(a.g_beta + (size_t)it * a.K + k, dbe * softplus_grad_t(braw[v]));
                }
            }
            if (gl == 0) {
                const float Kf = (float)a.K;
                const float lxi = logf(xi), leta = logf(eta);
                const float dxi = s * (-Kf * a.a / xi + sum_th - (a.a_prime - 1.f) / xi + a.b_prime);
                const float deta = t * (-Kf * a.c / eta + sum_be - (a.c_prime - 1.f) / eta + a.d_prime);
                atomicAdd(a.g_xi + u, dxi * softplus_grad_t(xr));
                atomicAdd(a.g_eta + it, deta * softplus_grad_t(er));
                // loss terms (hpf_pytorch.py:83, :145-152, :158-165, :169-173, :176-180)
                const float nll = lam - r * logf(lam);
                const float p_th = s * (-a.a * Kf * lxi + xi * sum_th - (a.a - 1.f) * sum_lth);
                const float p_be = t * (-a.c * Kf * leta + eta * sum_be - (a.c - 1.f) * sum_lbe);
                const float p_xi = s * (-(a.a_prime - 1.f) * lxi + a.b_prime * xi);
                const float p_eta = t * (-(a.c_prime - 1.f) * leta + a.d_prime * eta);
                loss = (double)nll + (double)p_th + (double)p_be + (double)p_xi + (double)p_eta;
            }
        }
    }
    // block reduction of the loss -> one float64 atomic per block
    __shared__ double s_red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) s_red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
        atomicAdd(a.loss, tot);
    }
}

// torch.optim.Adam (single tensor, defaults: amsgrad off, weight_decay 0, maximize off), one launch for
// a whole parameter tensor.  step_size = lr / (1 - beta1^t) and bc2_sqrt = sqrt(1 - beta2^t) are computed
// by the host in float64 exactly as torch does and passed as floats.
__global__ void adam_dense_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, int64_t n, float beta1, float beta2, float eps,
                                  float step_size, float bc2_sqrt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - beta1);       // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * beta2 + (1.f - beta2) * gi * gi;   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - step_size * (mi / denom);                    // param.addcdiv_(exp_avg, denom, value=-step_size)
}

template <typename IdT>
__global__ void __launch_bounds__(256) hpf_map_predict_kernel(const void* users, const void* items, int64_t n,
                                                              const float* __restrict__ theta,
                                                              const float* __restrict__ beta, int N, int M, int K,
                                                              float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= n) return;
    const int64_t u = (int64_t)((const IdT*)users)[wid], it = (int64_t)((const IdT*)items)[wid];
    float acc = 0.f;
    if (u >= 0 && u < N && it >= 0 && it < M) {
        const float* tr = theta + (size_t)u * K;
        const float* br = beta + (size_t)it * K;
        for (int k = lane; k < K; k += 32) acc = fmaf(softplus_t(tr[k]), softplus_t(br[k]), acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[wid] = acc;
}

}  // namespace pmf

using namespace pmf;

extern "C" {

int pmf_hpf_map_loss_grad(const void* d_users, const void* d_items, int32_t id_bytes, const float* d_ratings,
                          int64_t B, const float* d_theta_raw, const float* d_beta_raw, const float* d_xi_raw,
                          const float* d_eta_raw, const float* d_user_scale, const float* d_item_scale, int32_t N,
                          int32_t M, int32_t K, float a, float a_prime, float b_prime, float c, float c_prime,
                          float d_prime, float* d_g_theta, float* d_g_beta, float* d_g_xi, float* d_g_eta,
                          double* d_loss, int32_t* d_bad, void* stream) {
    PMF_REQUIRE(B >= 0, "B < 0");
    PMF_REQUIRE(id_bytes == 4 || id_bytes == 8, "id_bytes must be 4 or 8");
    PMF_REQUIRE(K >= 1 && K <= 256, "K=%d outside [1, 256]", K);
    PMF_REQUIRE(d_theta_raw && d_beta_raw && d_xi_raw && d_eta_raw && d_user_scale && d_item_scale, "NULL parameter");
    PMF_REQUIRE(d_g_theta && d_g_beta && d_g_xi && d_g_eta && d_loss && d_bad, "NULL output");
    if (B == 0) return PMF_OK;
    PMF_REQUIRE(d_users && d_items && d_ratings, "NULL batch");
    MapArgs m;
    m.users = d_users; m.items = d_items; m.ratings = d_ratings; m.B = B;
    m.theta = d_theta_raw; m.beta = d_beta_raw; m.xi = d_xi_raw; m.eta = d_eta_raw;
    m.user_scale = d_user_scale; m.item_scale = d_item_scale; m.N = N; m.M = M; m.K = K;
    m.a = a; m.a_prime = a_prime; m.b_prime = b_prime; m.c = c; m.c_prime = c_prime; m.d_prime = d_prime;
    m.g_theta = d_g_theta; m.g_beta = d_g_beta; m.g_xi = d_g_xi; m.g_eta = d_g_eta; m.loss = d_loss; m.bad = d_bad;
    cudaStream_t s = (cudaStream_t)stream;
    if (K <= 64) {
        const unsigned grid = (unsigned)cdiv(B * 8, 256);
        if (id_bytes == 8) hpf_map_loss_grad_kernel<8, int64_t><<<grid, 256, 0, s>>>(m);
        else hpf_map_loss_grad_kernel<8, int32_t><<<grid, 256, 0, s>>>(m);
    } else {
        const unsigned grid = (unsigned)cdiv(B * 32, 256);
        if (id_bytes == 8) hpf_map_loss_grad_kernel<32, int64_t><<<grid, 256, 0, s>>>(m);
        else hpf_map_loss_grad_kernel<32, int32_t><<<grid, 256, 0, s>>>(m);
    }
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

int pmf_adam_dense_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                        float beta1, float beta2, float eps, float step_size, float bias_correction2_sqrt,
                        void* stream) {
    PMF_REQUIRE(n >= 0 && (n == 0 || (d_param && d_grad && d_exp_avg && d_exp_avg_sq)), "bad argument");
    if (n == 0) return PMF_OK;
    adam_dense_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n,
                                                                                beta1, beta2, eps, step_size,
                                                                                bias_correction2_sqrt);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

int pmf_hpf_map_predict(const void* d_users, const void* d_items, int32_t id_bytes, int64_t n,
                        const float* d_theta_raw, const float* d_beta_raw, int32_t N, int32_t M, int32_t K,
                        float* d_out, void* stream) {
    PMF_REQUIRE(n >= 0 && (id_bytes == 4 || id_bytes == 8), "bad argument");
    if (n == 0) return PMF_OK;
    PMF_REQUIRE(d_users && d_items && d_theta_raw && d_beta_raw && d_out, "NULL argument");
    const unsigned grid = (unsigned)cdiv(n * 32, 256);
    if (id_bytes == 8) hpf_map_predict_kernel<int64_t><<<grid, 256, 0, (cudaStream_t)stream>>>(d_users, d_items, n, d_theta_raw, d_beta_raw, N, M, K, d_out);
    else hpf_map_predict_kernel<int32_t><<<grid, 256, 0, (cudaStream_t)stream>>>(d_users, d_items, n, d_theta_raw, d_beta_raw, N, M, K, d_out);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

}  // extern "C"
