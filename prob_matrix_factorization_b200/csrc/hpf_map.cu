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
// both at once, one exponential: softplus(z) = log1p(e^z), softplus'(z) = e^z / (1 + e^z)
__device__ __forceinline__ void softplus_both(float z, float& sp, float& dsp) {
    if (z > 20.f) { sp = z; dsp = 1.f; return; }
    const float e = expf(z);
    sp = log1pf(e);
    dsp = __fdividef(e, 1.f + e);   // 2 ulp: the gradients are compared at 1e-5
}

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

// Loss terms and gradients of ONE batch element (u, it, r), computed by the G lanes of a group (lane l owns factors
// k = l, l+G, ...); gradients are added to the dense gradient tensors with float atomics (duplicate ids in a batch).
// Returns the element's loss (valid in lane 0 of the group).
template <int G, int MAXV>   // K <= G * MAXV
__device__ __forceinline__ double map_element(const MapArgs& a, int64_t u, int64_t it, float r, int gl, unsigned gmask) {
    const float s = a.user_scale[u], t = a.item_scale[it];
    const float* tr = a.theta + (size_t)u * a.K;
    const float* br = a.beta + (size_t)it * a.K;
    float th[MAXV], be[MAXV], dth_raw[MAXV], dbe_raw[MAXV];   // values and d softplus / d raw
    float dot = 0.f, sum_th = 0.f, sum_be = 0.f, sum_lth = 0.f, sum_lbe = 0.f;
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
        const int k = gl + v * G;
        if (k < a.K) {
            // L2 loads: the rows may have been settled by another SM earlier in this kernel
            softplus_both(__ldcg(tr + k), th[v], dth_raw[v]);
            softplus_both(__ldcg(br + k), be[v], dbe_raw[v]);
            dot = fmaf(th[v], be[v], dot);
            sum_th += th[v]; sum_be += be[v];
            sum_lth += __logf(th[v]); sum_lbe += __logf(be[v]);   // loss only (reported per epoch, compared at 1e-5)
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
    float xi, eta, dxi_raw, deta_raw;
    softplus_both(__ldcg(a.xi + u), xi, dxi_raw);
    softplus_both(__ldcg(a.eta + it), eta, deta_raw);
    const float lam = fmaxf(dot, 1e-6f);                         // hpf_pytorch.py:80
    const float g = dot >= 1e-6f ? 1.f - __fdividef(r, lam) : 0.f;   // clamp passes gradient only inside its range
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
        const int k = gl + v * G;
        if (k < a.K) {
            const float dth = g * be[v] + s * (xi - __fdividef(a.a - 1.f, th[v]));
            const float dbe = g * th[v] + t * (eta - __fdividef(a.c - 1.f, be[v]));
            atomicAdd(a.g_theta + (size_t)u * a.K + k, dth * dth_raw[v]);
            atomicAdd(a.g_beta + (size_t)it * a.K + k, dbe * dbe_raw[v]);
        }
    }
    double loss = 0.0;
    if (gl == 0) {
        const float Kf = (float)a.K;
        const float lxi = logf(xi), leta = logf(eta);
        const float dxi = s * (-Kf * a.a / xi + sum_th - (a.a_prime - 1.f) / xi + a.b_prime);
        const float deta = t * (-Kf * a.c / eta + sum_be - (a.c_prime - 1.f) / eta + a.d_prime);
        atomicAdd(a.g_xi + u, dxi * dxi_raw);
        atomicAdd(a.g_eta + it, deta * deta_raw);
        // loss terms (hpf_pytorch.py:83, :145-152, :158-165, :169-173, :176-180)
        const float nll = lam - r * logf(lam);
        const float p_th = s * (-a.a * Kf * lxi + xi * sum_th - (a.a - 1.f) * sum_lth);
        const float p_be = t * (-a.c * Kf * leta + eta * sum_be - (a.c - 1.f) * sum_lbe);
        const float p_xi = s * (-(a.a_prime - 1.f) * lxi + a.b_prime * xi);
        const float p_eta = t * (-(a.c_prime - 1.f) * leta + a.d_prime * eta);
        loss = (double)nll + (double)p_th + (double)p_be + (double)p_xi + (double)p_eta;
    }
    return loss;
}

// block reduction of the loss -> one float64 atomic per block
__device__ __forceinline__ void block_add_loss(double loss, double* out) {
    __shared__ double s_red[8];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if (lane == 0) s_red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
        atomicAdd(out, tot);
    }
}

// One group of G lanes per batch element.
template <int G, int V, typename IdT>
__global__ void __launch_bounds__(256) hpf_map_loss_grad_kernel(const MapArgs a) {
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    double loss = 0.0;
    if (gid < a.B) {
        const int64_t u = (int64_t)((const IdT*)a.users)[gid], it = (int64_t)((const IdT*)a.items)[gid];
        if (u < 0 || u >= a.N || it < 0 || it >= a.M) {
            if (gl == 0) atomicOr(a.bad, 1);   // torch indexing would raise IndexError
        } else {
            loss = map_element<G, V>(a, u, it, a.ratings[gid], gl, gmask);
        }
    }
    block_add_loss(loss, a.loss);
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

// The same update on four consecutive elements per thread (128-bit accesses: a quarter of the memory instructions for
// the 28 bytes per parameter this kernel moves); element-wise arithmetic identical to adam_dense_kernel.
__global__ void __launch_bounds__(256) adam_dense_vec4_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                              float4* __restrict__ m, float4* __restrict__ v, int64_t n4,
                                                              float beta1, float beta2, float eps, float step_size,
                                                              float bc2_sqrt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 g4 = g[i];
    float4 m4 = m[i], v4 = v[i], p4 = p[i];
    const float gs[4] = {g4.x, g4.y, g4.z, g4.w};
    float ms[4] = {m4.x, m4.y, m4.z, m4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float mi = ms[k] + (gs[k] - ms[k]) * (1.f - beta1);
        const float vi = vs[k] * beta2 + (1.f - beta2) * gs[k] * gs[k];
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        ms[k] = mi;
        vs[k] = vi;
        ps[k] = ps[k] - step_size * (mi / denom);
    }
    m[i] = make_float4(ms[0], ms[1], ms[2], ms[3]);
    v[i] = make_float4(vs[0], vs[1], vs[2], vs[3]);
    p[i] = make_float4(ps[0], ps[1], ps[2], ps[3]);
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


// ------------------------------------------------------------------------------------------------
// Lazy ("touch-only") Adam, equivalent to torch's dense Adam.
// torch.optim.Adam updates EVERY row at EVERY step: a row without gradient still decays m and v and moves
// by -step_size*m/(sqrt(v)/bc2+eps).  Those zero-gradient steps depend only on the row's own (p, m, v) and on
// per-step scalars, so they can be applied when the row is next referenced: each row remembers the last step it is up
// to date with (`last`, see settle_row); when a step references a row, the row first takes its deferred real step and
// the run of zero-gradient steps it skipped (closed form, RowCatchUp), then collects this step's gradient.  Traffic
// per step drops from (N+M)(K+1)*28 bytes to the referenced rows.
// ------------------------------------------------------------------------------------------------
struct LazyState {   // mirrors pmf_lazy_adam in pmf_b200.h
    float *p[4], *m[4], *v[4], *g[4];          // theta (N,K), beta (M,K), xi (N), eta (M)
    int32_t *last_user, *last_item, *claim_user, *claim_item;
    const float *step_size, *bc2_sqrt;          // indexed by step number (1-based)
    float beta1, beta2, eps;
    const double *tail1, *tail2;                // closed-form catch-up tables (see RowCatchUp), or NULL: replay step by step
    const double* pow5;                         // [5][n_pow]: beta1^J, beta2^J, rho^J, r2^J, beta2^(-J/2) for J = 0 .. n_pow-1
    int32_t n_pow;
};

// Closed form of a run of zero-gradient steps.  After the row's last real step s0, step s = s0 + j does
//     m_s = beta1^j m_0,   v_s = beta2^j v_0,   p_s = p_{s-1} - ss_s m_s / (sqrt(v_s)/bc_s + eps)
// (ss_s = lr/(1-beta1^s), bc_s = sqrt(1-beta2^s)).  With q = m_0/sqrt(v_0), rho = beta1/sqrt(beta2), r2 = beta1/beta2 and
// d_s = eps bc_s/sqrt(v_s) (<= 1e-3 enforced, else the steps are replayed one by one):
//     sum_j ss_s m_s/(sqrt(v_s)/bc_s + eps) = q sum_j ss_s bc_s rho^j/(1+d_s) = q [W1 - (eps/sqrt(v_0)) W2] + O(d^2),
//     W1 = sum_j ss_s bc_s rho^j,  W2 = sum_j ss_s bc_s^2 r2^j
// -- per-ROW scalars, obtained from the host-built float64 tables tail1[s] = sum_{t>s} ss_t bc_t rho^(t-s) (tail2 alike with
// bc_t^2, r2) as W1 = tail1[s0] - rho^J tail1[s0+J].  A catch-up then costs one sqrt and one divide per ELEMENT instead of
// one of each per element per skipped step: the replay was 97 % of the lazy epoch (profiles/README.md).  Real-arithmetic
// identical to the dense update up to O(d^2) <= 1e-6 of an update that is itself <= ~1e-2 of the parameter; the float32
// rounding differs from the step-by-step sequence by a few ulp (tests: 1e-5 max-norm relative on p, m, v).
struct RowCatchUp {
    int J;                  // number of zero-gradient steps to apply
    int from, to;           // first / last step of the run (fallback replay)
    float d1, d2;           // beta1^J, beta2^J
    float W1, W2;
    float dmax_num;         // eps * bc_to / sqrt(beta2^J): divided by sqrt(v_0) this is the largest d_s of the run
    bool closed;            // tables available for this run (else: step-by-step replay)
};

__device__ __forceinline__ RowCatchUp row_catch_up(const LazyState& L, int last, int to) {
    RowCatchUp c;
    c.from = last + 1; c.to = to; c.J = to - last;
    c.closed = L.tail1 != nullptr && c.J < L.n_pow;
    if (c.J <= 0 || !c.closed) { c.d1 = c.d2 = 1.f; c.W1 = c.W2 = 0.f; c.dmax_num = 0.f; return c; }
    // powers of the decay factors come from host-built float64 tables (a run is never longer than the call's step count):
    // five loads instead of float64 exp / log / sqrt / divide sequences, which were ~1300 of the ~4600 warp instructions
    // a batch element cost (ncu, profiles/README.md)
    const double* T = L.pow5 + c.J;
    const size_t n = (size_t)L.n_pow;
    c.d1 = (float)T[0]; c.d2 = (float)T[n];
    c.W1 = (float)(L.tail1[last] - T[2 * n] * L.tail1[to]);
    c.W2 = (float)(L.tail2[last] - T[3 * n] * L.tail2[to]);
    c.dmax_num = (float)((double)L.eps * (double)L.bc2_sqrt[to] * T[4 * n]);
    return c;
}

// step-by-step replay of zero-gradient steps (same expressions as adam_dense_kernel with grad = 0): the fallback of the
// closed form; by value and not inlined, so that it costs the hot path neither registers nor local memory
__device__ __noinline__ float3 adam_zero_grad_replay(float p, float m, float v, int from, int to, const float* __restrict__ step_size,
                                                     const float* __restrict__ bc2_sqrt, float beta1, float beta2, float eps) {
    for (int s = from; s <= to; ++s) {
        m = m + (0.f - m) * (1.f - beta1);
        v = v * beta2 + (1.f - beta2) * 0.f * 0.f;
        const float denom = sqrtf(v) / bc2_sqrt[s] + eps;
        p = p - step_size[s] * (m / denom);
    }
    return make_float3(p, m, v);
}
__device__ __forceinline__ void adam_zero_grad_steps(float& p, float& m, float& v, int from, int to, const LazyState& L) {
    const float3 r = adam_zero_grad_replay(p, m, v, from, to, L.step_size, L.bc2_sqrt, L.beta1, L.beta2, L.eps);
    p = r.x; m = r.y; v = r.z;
}

__device__ __forceinline__ void adam_one_step(float& p, float& m, float& v, float g, int s, const LazyState& L) {
    m = m + (g - m) * (1.f - L.beta1);
    v = v * L.beta2 + (1.f - L.beta2) * g * g;
    const float denom = sqrtf(v) / L.bc2_sqrt[s] + L.eps;
    p = p - L.step_size[s] * (m / denom);
}

__device__ __forceinline__ void apply_catch_up(float& p, float& m, float& v, const RowCatchUp& c, const LazyState& L) {
    if (c.J <= 0) return;
    if (!c.closed) { adam_zero_grad_steps(p, m, v, c.from, c.to, L); return; }
    if (m != 0.f) {
        const float rs = rsqrtf(v);          // 1/sqrt(v) (2 ulp; the result is compared at 1e-5): no division, no IEEE sqrt
        if (!(c.dmax_num * rs <= 1e-3f)) {   // eps is not negligible against sqrt(v) somewhere in the run (or v == 0): replay
            adam_zero_grad_steps(p, m, v, c.from, c.to, L);
            return;
        }
        p = p - (m * rs) * (c.W1 - (L.eps * rs) * c.W2);
    }
    m *= c.d1;
    v *= c.d2;
}

// Row bookkeeping of the lazy scheme.  last[row] = s > 0: (p, m, v) are up to date with step s-1 and g[row] holds the
// gradient of step s, NOT yet applied;  last[row] = -s <= 0: up to date with step s, nothing pending.
// settle_row brings the row to "up to date with step t-1, gradient zeroed, pending at t": it applies the pending real
// step (torch's update with the accumulated gradient), then the run of zero-gradient steps in closed form.
template <int G, int V>   // K + 1 <= G * (V + 1)
__device__ __forceinline__ void settle_row(const LazyState& L, int mat, int64_t row, int K, int t, int gl, unsigned gmask,
                                           int32_t* last_arr) {
    const int last = last_arr[row];
    const int pending = last > 0 ? last : 0;
    const RowCatchUp c = row_catch_up(L, last > 0 ? last : -last, t - 1);
    // elements 0..K-1 are the factor row, element K is the row's scalar (xi / eta): one pass, no single-lane tail.
    // All loads of the row are issued before the first dependent instruction (one memory round trip per row, not one per
    // element: the kernel is latency-bound -- ncu: 39 % warps active, 38 % of stalls on the long scoreboard).
    constexpr int VS = V + 1;
    float p[VS], m[VS], v[VS], g[VS];
    const bool work = pending || c.J > 0;
#pragma unroll
    for (int q = 0; q < VS; ++q) {
        const int k = gl + q * G;
        if (work && k <= K) {
            const bool vec = k < K;
            const size_t e = vec ? (size_t)row * K + k : (size_t)row;
            const int tb = vec ? mat : mat + 2;
            p[q] = L.p[tb][e]; m[q] = L.m[tb][e]; v[q] = L.v[tb][e];
            g[q] = pending ? L.g[tb][e] : 0.f;
        }
    }
#pragma unroll
    for (int q = 0; q < VS; ++q) {
        const int k = gl + q * G;
        if (k <= K) {
            const bool vec = k < K;
            const size_t e = vec ? (size_t)row * K + k : (size_t)row;
            const int tb = vec ? mat : mat + 2;
            if (work) {
                if (pending) adam_one_step(p[q], m[q], v[q], g[q], pending, L);
                apply_catch_up(p[q], m[q], v[q], c, L);
                L.p[tb][e] = p[q]; L.m[tb][e] = m[q]; L.v[tb][e] = v[q];
            }
            L.g[tb][e] = 0.f;
        }
    }
    // publish: the group's stores happen-before lane 0's fence (warp barrier), the fence makes them visible device-wide
    // before the row is marked ready for step t (release is cumulative)
    __syncwarp(gmask);
    if (gl == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(last_arr + row), "r"(t) : "memory");
    }
}

__device__ __forceinline__ void wait_row_ready(const int32_t* last_arr, int64_t row, int t) {
    int v;
    do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(last_arr + row) : "memory");
    } while (v != t);
}

// ONE kernel per mini-batch step t (the reference does loss.backward() + a dense Adam step over every parameter,
// compare_models.py:305-313).  A group of G lanes per batch element:
//   1. claim: the first group of this step to see a row (atomicMax on claim[row]) settles it -- applies the row's pending
//      Adam step, catches it up over the zero-gradient steps it skipped, zeroes its gradient -- and marks it ready;
//   2. the other groups that reference the row spin until it is ready (the claimant is a RUNNING group, and it settles
//      both of its rows before it waits for anything, so there is no cycle);
//   3. loss terms + gradients of the element, float atomics into the rows' gradients.
// The Adam step with the accumulated gradient is deferred to the row's next settle (or the final flush): kernel
// boundaries order "all gradients of step t" before "the next settle of the row", so no grid-wide barrier is needed.
template <int G, int V, typename IdT>
__global__ void __launch_bounds__(256, 4) lazy_step_kernel(const LazyState L, const MapArgs a, int t) {
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    double loss = 0.0;
    if (gid < a.B) {
        const int64_t u = (int64_t)((const IdT*)a.users)[gid], it = (int64_t)((const IdT*)a.items)[gid];
        if (u < 0 || u >= a.N || it < 0 || it >= a.M) {
            if (gl == 0) atomicOr(a.bad, 1);   // torch indexing would raise IndexError
        } else {
            int first_u = 0, first_i = 0;
            if (gl == 0) {
                first_u = atomicMax(L.claim_user + u, t) < t;
                first_i = atomicMax(L.claim_item + it, t) < t;
            }
            first_u = __shfl_sync(gmask, first_u, lane & ~(G - 1));
            first_i = __shfl_sync(gmask, first_i, lane & ~(G - 1));
            if (first_u) settle_row<G, V>(L, 0, u, a.K, t, gl, gmask, L.last_user);
            if (first_i) settle_row<G, V>(L, 1, it, a.K, t, gl, gmask, L.last_item);
            if (gl == 0) {
                if (!first_u) wait_row_ready(L.last_user, u, t);
                if (!first_i) wait_row_ready(L.last_item, it, t);
            }
            __syncwarp(gmask);
            loss = map_element<G, V>(a, u, it, a.ratings[gid], gl, gmask);
        }
    }
    block_add_loss(loss, a.loss);
}

// bring every row up to step t, nothing pending (end of training / before the parameters are read from outside)
__global__ void __launch_bounds__(256) lazy_flush_kernel(const LazyState L, int N, int M, int K, int t) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (int64_t)N + M) return;
    const int mat = wid < N ? 0 : 1;
    const int64_t row = mat == 0 ? wid : wid - N;
    int32_t* last_arr = mat == 0 ? L.last_user : L.last_item;
    const int last = last_arr[row];
    const int pending = last > 0 ? last : 0;
    const RowCatchUp c = row_catch_up(L, last > 0 ? last : -last, t);
    if (!pending && c.J <= 0) return;
    for (int k = lane; k < K; k += 32) {
        const size_t e = (size_t)row * K + k;
        float p = L.p[mat][e], m = L.m[mat][e], v = L.v[mat][e];
        if (pending) adam_one_step(p, m, v, L.g[mat][e], pending, L);
        apply_catch_up(p, m, v, c, L);
        L.p[mat][e] = p; L.m[mat][e] = m; L.v[mat][e] = v;
    }
    __syncwarp();
    if (lane == 0) {
        float p = L.p[mat + 2][row], m = L.m[mat + 2][row], v = L.v[mat + 2][row];
        if (pending) adam_one_step(p, m, v, L.g[mat + 2][row], pending, L);
        apply_catch_up(p, m, v, c, L);
        L.p[mat + 2][row] = p; L.m[mat + 2][row] = m; L.v[mat + 2][row] = v;
        last_arr[row] = -t;
    }
}

}  // namespace pmf

using namespace pmf;

// kernel<G, V, IdT>: V = factors per lane rounded up to 1, 2, 4 or 8 (the per-lane loops are unrolled V times)
#define PMF_DISPATCH_V(kernel, G_, v_, args)                                                             \
    do {                                                                                                 \
        const int vv_ = (v_);                                                                            \
        if (id_bytes == 8) {                                                                             \
            if (vv_ <= 1) kernel<G_, 1, int64_t><<<grid, 256, 0, s>>> args;                              \
            else if (vv_ <= 2) kernel<G_, 2, int64_t><<<grid, 256, 0, s>>> args;                         \
            else if (vv_ <= 4) kernel<G_, 4, int64_t><<<grid, 256, 0, s>>> args;                         \
            else kernel<G_, 8, int64_t><<<grid, 256, 0, s>>> args;                                       \
        } else {                                                                                         \
            if (vv_ <= 1) kernel<G_, 1, int32_t><<<grid, 256, 0, s>>> args;                              \
            else if (vv_ <= 2) kernel<G_, 2, int32_t><<<grid, 256, 0, s>>> args;                         \
            else if (vv_ <= 4) kernel<G_, 4, int32_t><<<grid, 256, 0, s>>> args;                         \
            else kernel<G_, 8, int32_t><<<grid, 256, 0, s>>> args;                                       \
        }                                                                                                \
    } while (0)

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
        PMF_DISPATCH_V(hpf_map_loss_grad_kernel, 8, (K + 7) / 8, (m));
    } else {
        const unsigned grid = (unsigned)cdiv(B * 32, 256);
        PMF_DISPATCH_V(hpf_map_loss_grad_kernel, 32, (K + 31) / 32, (m));
    }
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

int pmf_adam_dense_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, int64_t n,
                        float beta1, float beta2, float eps, float step_size, float bias_correction2_sqrt,
                        void* stream) {
    PMF_REQUIRE(n >= 0 && (n == 0 || (d_param && d_grad && d_exp_avg && d_exp_avg_sq)), "bad argument");
    if (n == 0) return PMF_OK;
    const uintptr_t align = (uintptr_t)d_param | (uintptr_t)d_grad | (uintptr_t)d_exp_avg | (uintptr_t)d_exp_avg_sq;
    if (n % 4 == 0 && (align & 15) == 0) {
        adam_dense_vec4_kernel<<<(unsigned)cdiv(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
            (float4*)d_param, (const float4*)d_grad, (float4*)d_exp_avg, (float4*)d_exp_avg_sq, n / 4, beta1, beta2, eps,
            step_size, bias_correction2_sqrt);
    } else {
        adam_dense_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n,
                                                                                    beta1, beta2, eps, step_size,
                                                                                    bias_correction2_sqrt);
    }
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


static int lazy_state_from(const pmf_lazy_adam* st, LazyState& L) {
    PMF_REQUIRE(st != nullptr, "state is NULL");
    float* const P[4] = {st->theta, st->beta, st->xi, st->eta};
    float* const Mo[4] = {st->m_theta, st->m_beta, st->m_xi, st->m_eta};
    float* const Vo[4] = {st->v_theta, st->v_beta, st->v_xi, st->v_eta};
    float* const Gr[4] = {st->g_theta, st->g_beta, st->g_xi, st->g_eta};
    for (int k = 0; k < 4; ++k) {
        PMF_REQUIRE(P[k] && Mo[k] && Vo[k] && Gr[k], "NULL parameter / moment / gradient tensor");
        L.p[k] = P[k]; L.m[k] = Mo[k]; L.v[k] = Vo[k]; L.g[k] = Gr[k];
    }
    PMF_REQUIRE(st->last_user && st->last_item && st->claim_user && st->claim_item && st->step_size && st->bc2_sqrt,
                "NULL bookkeeping array");
    L.last_user = st->last_user; L.last_item = st->last_item; L.claim_user = st->claim_user; L.claim_item = st->claim_item;
    L.step_size = st->step_size; L.bc2_sqrt = st->bc2_sqrt;
    L.beta1 = st->beta1; L.beta2 = st->beta2; L.eps = st->eps;
    PMF_REQUIRE((st->tail1 == nullptr) == (st->tail2 == nullptr) && (st->tail1 == nullptr) == (st->pow5 == nullptr),
                "tail1, tail2 and pow5 go together");
    PMF_REQUIRE(st->pow5 == nullptr || st->n_pow >= 1, "n_pow must be positive");
    L.tail1 = st->tail1; L.tail2 = st->tail2; L.pow5 = st->pow5; L.n_pow = st->n_pow;
    return PMF_OK;
}

int pmf_hpf_map_lazy_epoch(const pmf_lazy_adam* st, const void* d_users, const void* d_items, int32_t id_bytes,
                           const float* d_ratings, int64_t n, int64_t batch, int64_t step0, const float* d_user_scale,
                           const float* d_item_scale, int32_t N, int32_t M, int32_t K, float a, float a_prime,
                           float b_prime, float c, float c_prime, float d_prime, double* d_loss, int32_t* d_bad,
                           void* stream) {
    LazyState L;
    PMF_TRY(lazy_state_from(st, L));
    PMF_REQUIRE(n >= 0 && batch >= 1 && step0 >= 0, "bad sizes");
    PMF_REQUIRE(id_bytes == 4 || id_bytes == 8, "id_bytes must be 4 or 8");
    PMF_REQUIRE(K >= 1 && K <= 256, "K=%d outside [1, 256]", K);
    PMF_REQUIRE(d_user_scale && d_item_scale && d_loss && d_bad, "NULL argument");
    if (n == 0) return PMF_OK;
    PMF_REQUIRE(d_users && d_items && d_ratings, "NULL batch");
    cudaStream_t s = (cudaStream_t)stream;
    const int G = K <= 64 ? 8 : 32;
    MapArgs m;
    m.theta = st->theta; m.beta = st->beta; m.xi = st->xi; m.eta = st->eta;
    m.user_scale = d_user_scale; m.item_scale = d_item_scale; m.N = N; m.M = M; m.K = K;
    m.a = a; m.a_prime = a_prime; m.b_prime = b_prime; m.c = c; m.c_prime = c_prime; m.d_prime = d_prime;
    m.g_theta = st->g_theta; m.g_beta = st->g_beta; m.g_xi = st->g_xi; m.g_eta = st->g_eta;
    m.loss = d_loss; m.bad = d_bad;
    int64_t t = step0;
    for (int64_t off = 0; off < n; off += batch) {
        const int64_t B = (n - off) < batch ? (n - off) : batch;
        ++t;
        PMF_REQUIRE(t < INT32_MAX, "too many steps");
        m.users = (const char*)d_users + off * id_bytes;
        m.items = (const char*)d_items + off * id_bytes;
        m.ratings = d_ratings + off;
        m.B = B;
        const unsigned grid = (unsigned)cdiv(B * G, 256);
        if (G == 8) { PMF_DISPATCH_V(lazy_step_kernel, 8, (K + 7) / 8, (L, m, (int)t)); }
        else { PMF_DISPATCH_V(lazy_step_kernel, 32, (K + 31) / 32, (L, m, (int)t)); }
        PMF_LAUNCH_CHECK();
    }
    return PMF_OK;
}

int pmf_hpf_map_lazy_flush(const pmf_lazy_adam* st, int32_t N, int32_t M, int32_t K, int64_t step_now, void* stream) {
    LazyState L;
    PMF_TRY(lazy_state_from(st, L));
    PMF_REQUIRE(step_now >= 0 && step_now < INT32_MAX, "bad step");
    lazy_flush_kernel<<<(unsigned)cdiv(((int64_t)N + M) * 32, 256), 256, 0, (cudaStream_t)stream>>>(L, N, M, K, (int)step_now);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

}  // extern "C"
