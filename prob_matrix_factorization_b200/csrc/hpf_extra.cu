// a11: textbook-HPF extras that exist only in the reference's documentation, not in its code
// (PARITY UNPINNED -- checked against oracle/pmf_oracle.py's restatement only):
//   * the digamma ("multinomial") allocation of docs/Models.tex:652-664: phi_k ∝ exp(E log th_k + E log be_k)
//   * the evidence lower bound (docs/Models.tex:583-726; Gopalan, Hofman & Blei 2015), rate term over
//     observed pairs only, matching hpf_cavi.py:149-151.
// The geometric-mean tables G = exp(psi(shape)) / rate turn the multinomial step into the same SDDMM as the
// mean-based pass: phi_k = G_self_k G_oth_k / sum_k G_self_k G_oth_k.
#include "common.cuh"

namespace pmf {

// psi(x), x > 0: recurrence up to x >= 6 then the asymptotic series.  T = float or double.
template <typename T>
__device__ __forceinline__ T digamma_pos(T x) {
    T r = T(0);
#pragma unroll 1
    while (x < T(6)) { r -= T(1) / x; x += T(1); }
    const T inv = T(1) / x, inv2 = inv * inv;
    const T series = inv2 * (T(1.0 / 12) - inv2 * (T(1.0 / 120) - inv2 * (T(1.0 / 252) - inv2 * (T(1.0 / 240) - inv2 * T(1.0 / 132)))));
    return r + log(x) - T(0.5) * inv - series;
}

// lgamma(s) and psi(s) together, s > 0, float64: shift to x = s + n >= 8 with the running product den = s (s+1) ... (s+n-1)
// and num / den = sum 1/(s+i) kept as a fraction (one division at the end, no division per shift), then the Stirling /
// asymptotic series (truncation error < 1e-12 at x = 8):
//   lgamma(s) = lgamma(x) - log(den),  psi(s) = psi(x) - num/den.
// Three logarithms and two divisions per call instead of libdevice's lgamma plus a division per shift: the prior /
// entropy terms of the ELBO evaluate this pair (N + M) K times per sweep and were ~85 % of the ELBO's time.
__host__ __device__ __forceinline__ void lgamma_digamma(double s, double& lg, double& psi) {
    double x = s, num = 0.0, den = 1.0;
#pragma unroll 1
    while (x < 8.0) { num = num * x + den; den *= x; x += 1.0; }
    const double lx = log(x), inv = 1.0 / x, inv2 = inv * inv;
    lg = (x - 0.5) * lx - x + 0.91893853320467274178 +
         inv * (1.0 / 12 - inv2 * (1.0 / 360 - inv2 * (1.0 / 1260 - inv2 * (1.0 / 1680 - inv2 * (1.0 / 1188)))));
    psi = lx - 0.5 * inv -
          inv2 * (1.0 / 12 - inv2 * (1.0 / 120 - inv2 * (1.0 / 252 - inv2 * (1.0 / 240 - inv2 * (1.0 / 132 - inv2 * (691.0 / 32760))))));
    if (den != 1.0) { lg -= log(den); psi -= num / den; }
}

// float32 version for the per-element terms of the ELBO: inputs are float32 tables, each term is accurate to ~1e-7 relative
// and the (N + M) K terms are accumulated in float64 -- measured 1e-7 relative on the total against the float64 form
// (DESIGN.md §6), well inside the 1e-5 the ELBO is compared at, at a fifth of the instructions (no float64 log / divide).
__device__ __forceinline__ void lgamma_digamma_f(float s, float& lg, float& psi) {
    float x = s, num = 0.f, den = 1.f;
#pragma unroll 1
    while (x < 8.f) { num = fmaf(num, x, den); den *= x; x += 1.f; }
    const float lx = logf(x), inv = __fdividef(1.f, x), inv2 = inv * inv;
    lg = (x - 0.5f) * lx - x + 0.91893853320467274178f +
         inv * (1.f / 12 - inv2 * (1.f / 360 - inv2 * (1.f / 1260 - inv2 * (1.f / 1680))));
    psi = lx - 0.5f * inv - inv2 * (1.f / 12 - inv2 * (1.f / 120 - inv2 * (1.f / 252 - inv2 * (1.f / 240 - inv2 * (1.f / 132)))));
    if (den != 1.f) { lg -= logf(den); psi -= __fdividef(num, den); }
}

__global__ void geomean_kernel(const float* __restrict__ shp, const float* __restrict__ rte, int64_t rows, int K,
                               int ld, float* __restrict__ G) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * ld) return;
    const int k = (int)(e % ld);
    G[e] = k < K ? expf(digamma_pos<float>(shp[e])) / rte[e] : 0.f;
}

struct DigammaArgs {
    const int4* seg_desc;
    const int32_t *col, *multi_row, *multi_first;
    const float* val;
    int32_t n_seg, n_multi, seg_len, row_offset, K, ld, nvec;
    const float *G_oth, *E_oth;
    float *G_self, *E_self, *shp, *rte;
    float shape_prior, rate_prior;
    const float* rate_prior_vec;
    float *hyper_rate, *hyper_mean;
    float hyper_shape, hyper_rate_prior;
    float* partial;
};

template <int V>
__device__ __forceinline__ void digamma_row_update(const DigammaArgs& a, int R, int gl, unsigned gmask,
                                                   const float4 (&self)[V], const float4 (&sa)[V], const float4 (&sb)[V]) {
    const float rp = a.rate_prior_vec ? a.rate_prior_vec[R] : a.rate_prior;
    float esum = 0.f;
    const size_t rowoff = (size_t)R * a.ld;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int idx = gl + v * 8;
        if (idx < a.nvec) {
            const int k0 = idx * 4;
            const float sv[4] = {self[v].x, self[v].y, self[v].z, self[v].w};
            const float av[4] = {sa[v].x, sa[v].y, sa[v].z, sa[v].w};
            const float bv[4] = {sb[v].x, sb[v].y, sb[v].z, sb[v].w};
            float s[4], r[4], e[4], g[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (k0 + c < a.K) {
                    s[c] = a.shape_prior + sv[c] * av[c];
                    r[c] = rp + bv[c];
                    e[c] = s[c] / r[c];
                    g[c] = expf(digamma_pos<float>(s[c])) / r[c];
                } else { s[c] = 0.f; r[c] = 1.f; e[c] = 0.f; g[c] = 0.f; }
                esum += e[c];
            }
            if (a.shp) *reinterpret_cast<float4*>(a.shp + rowoff + k0) = make_float4(s[0], s[1], s[2], s[3]);
            if (a.rte) *reinterpret_cast<float4*>(a.rte + rowoff + k0) = make_float4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<float4*>(a.E_self + rowoff + k0) = make_float4(e[0], e[1], e[2], e[3]);
            *reinterpret_cast<float4*>(a.G_self + rowoff + k0) = make_float4(g[0], g[1], g[2], g[3]);
        }
    }
    if (a.hyper_rate) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) esum += __shfl_xor_sync(gmask, esum, o);
        if (gl == 0) {
            const float hr = a.hyper_rate_prior + esum;
            a.hyper_rate[R] = hr;
            a.hyper_mean[R] = a.hyper_shape / hr;
        }
    }
}

// Same mapping as gamma_pass_kernel (8 lanes per segment, V float4 slices per lane), two gathers per rating.
template <int V>
__global__ void __launch_bounds__(256) digamma_pass_kernel(const DigammaArgs a) {
    const int lane = threadIdx.x & 31, gl = lane & 7;
    const unsigned gmask = 0xffu << (lane & ~7);
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool has = gid < a.n_seg;
    int row = 0, p = 0, end = 0, pidx = -1;
    if (has) {
        const int4 d = __ldg(a.seg_desc + gid);   // {row, start, end, partial slot}: one load, not a chain of three
        row = d.x; p = d.y; end = d.z; pidx = d.w;
    }
    const int R = a.row_offset + row;
    float4 self[V], sa[V], sb[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int idx = gl + v * 8;
        self[v] = (has && idx < a.nvec) ? *reinterpret_cast<const float4*>(a.G_self + (size_t)R * a.ld + idx * 4)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        sa[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        sb[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int maxlen = __reduce_max_sync(0xffffffffu, end - p);
    for (int base = 0; base < maxlen; base += 8) {
        const int q = p + base + gl;
        const bool okq = q < end;
        const int c_l = okq ? __ldg(a.col + q) : 0;
        const float x_l = okq ? __ldg(a.val + q) : 0.f;
        const int rem = end - (p + base);
#pragma unroll 2
        for (int j = 0; j < 8; ++j) {
            const int c = __shfl_sync(0xffffffffu, c_l, j, 8);
            const float x = __shfl_sync(0xffffffffu, x_l, j, 8);
            float4 og[V], oe[V];
            float d = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int idx = gl + v * 8;
                const bool ok = j < rem && idx < a.nvec;
                og[v] = ok ? ldg_f4(a.G_oth + (size_t)c * a.ld + idx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                oe[v] = ok ? ldg_f4(a.E_oth + (size_t)c * a.ld + idx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                d = fmaf(self[v].x, og[v].x, d); d = fmaf(self[v].y, og[v].y, d);
                d = fmaf(self[v].z, og[v].z, d); d = fmaf(self[v].w, og[v].w, d);
            }
            d += __shfl_xor_sync(0xffffffffu, d, 4);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            const float w = x / fmaxf(d, 1e-10f);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                sa[v].x = fmaf(w, og[v].x, sa[v].x); sb[v].x += oe[v].x;
                sa[v].y = fmaf(w, og[v].y, sa[v].y); sb[v].y += oe[v].y;
                sa[v].z = fmaf(w, og[v].z, sa[v].z); sb[v].z += oe[v].z;
                sa[v].w = fmaf(w, og[v].w, sa[v].w); sb[v].w += oe[v].w;
            }
        }
    }
    if (!has) return;
    if (pidx < 0) {
        digamma_row_update<V>(a, R, gl, gmask, self, sa, sb);
    } else {
        float* dst = a.partial + (size_t)pidx * 2 * a.ld;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int idx = gl + v * 8;
            if (idx < a.nvec) {
                *reinterpret_cast<float4*>(dst + idx * 4) = sa[v];
                *reinterpret_cast<float4*>(dst + a.ld + idx * 4) = sb[v];
            }
        }
    }
}

template <int V>
__global__ void __launch_bounds__(32) digamma_multi_kernel(const DigammaArgs a) {
    // one warp per multi-segment row; group 0 sums the partials in order (this path is not tuned)
    const int lane = threadIdx.x & 31, gl = lane & 7;
    if (lane >= 8) return;
    const int row = a.multi_row[blockIdx.x];
    const int first = a.multi_first[blockIdx.x], last = a.multi_first[blockIdx.x + 1];
    const int R = a.row_offset + row;
    float4 self[V], sa[V], sb[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int idx = gl + v * 8;
        self[v] = idx < a.nvec ? *reinterpret_cast<const float4*>(a.G_self + (size_t)R * a.ld + idx * 4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        sa[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        sb[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int q = first; q < last; ++q) {
        const float* src = a.partial + (size_t)q * 2 * a.ld;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int idx = gl + v * 8;
            if (idx < a.nvec) {
                const float4 pa = *reinterpret_cast<const float4*>(src + idx * 4);
                const float4 pb = *reinterpret_cast<const float4*>(src + a.ld + idx * 4);
                sa[v].x += pa.x; sa[v].y += pa.y; sa[v].z += pa.z; sa[v].w += pa.w;
                sb[v].x += pb.x; sb[v].y += pb.y; sb[v].z += pb.z; sb[v].w += pb.w;
            }
        }
    }
    digamma_row_update<V>(a, R, gl, 0xffu, self, sa, sb);
}

// ---- ELBO ------------------------------------------------------------------------------------------
struct ElboArgs {
    const int32_t *row_ptr, *col;
    const float* val;
    int32_t n_rows, row_offset, K, ld, nvec;
    const float *E_theta, *E_beta, *G_theta, *G_beta;
    double* out;
};

// log(n!) for n = 0..20 (ratings are small counts: lgamma(x + 1) is a table lookup for them)
__constant__ double kLogFactorial[21] = {
    0.0, 0.0, 0.69314718055994530942, 1.79175946922805500081, 3.17805383034794561965, 4.78749174278204599425,
    6.57925121201010099506, 8.52516136106541430017, 10.60460290274525022842, 12.80182748008146961121,
    15.10441257307551529523, 17.50230784587388583929, 19.98721449566188614952, 22.55216385312342288557,
    25.19122118273868150009, 27.89927138384089156609, 30.67186010608067280376, 33.50507345013688888401,
    36.39544520803305357622, 39.33988418719949403622, 42.33561646075348502966};

__device__ __forceinline__ double log_gamma_x_plus_1(double x) {
    const int n = (int)x;
    if (x >= 0.0 && x <= 20.0 && (double)n == x) return kLogFactorial[n];
    return lgamma(x + 1.0);
}

// likelihood part: one 8-lane group per SEGMENT of the rating list (segments longest first, as in the pass kernels: the
// row is known from the segment descriptor -- no search per rating -- and the four groups of a warp run similar lengths)
__global__ void __launch_bounds__(256) elbo_like_kernel(const ElboArgs a, const int4* __restrict__ seg_desc, int n_seg) {
    const int lane = threadIdx.x & 31, gl = lane & 7;
    const unsigned gmask = 0xffu << (lane & ~7);
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8;
    double acc = 0.0;
    if (gid < n_seg) {
        const int4 d = __ldg(seg_desc + gid);   // {row, start, end, .}
        const size_t ru = (size_t)(a.row_offset + d.x) * a.ld;
        for (int t = d.y; t < d.z; ++t) {
            const size_t rc = (size_t)__ldg(a.col + t) * a.ld;
            float dg = 0.f, de = 0.f;
            for (int idx = gl; idx < a.nvec; idx += 8) {
                const float4 gt = ldg_f4(a.G_theta + ru + idx * 4), gb = ldg_f4(a.G_beta + rc + idx * 4);
                const float4 et = ldg_f4(a.E_theta + ru + idx * 4), eb = ldg_f4(a.E_beta + rc + idx * 4);
                dg += gt.x * gb.x + gt.y * gb.y + gt.z * gb.z + gt.w * gb.w;
                de += et.x * eb.x + et.y * eb.y + et.z * eb.z + et.w * eb.w;
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                dg += __shfl_xor_sync(gmask, dg, o);
                de += __shfl_xor_sync(gmask, de, o);
            }
            if (gl == 0) {
                const double x = (double)__ldg(a.val + t);
                acc += x * log(fmax((double)dg, 1e-10)) - log_gamma_x_plus_1(x) - (double)de;
            }
        }
    }
    __shared__ double s_red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        atomicAdd(a.out + 0, t);
    }
}

// prior + entropy terms of one side: rows [row_begin, row_end).  out[slot_p] += E log p(factor | hyper),
// out[slot_h] += E log p(hyper), out[5] += entropies.
struct ElboRowConsts {   // lgamma / digamma of the three shape constants of a side, evaluated once on the host
    double lg_hs, psi_hs, lg_sp, lg_hp;
};

__global__ void __launch_bounds__(256) elbo_rows_kernel(const float* __restrict__ shp, const float* __restrict__ rte,
                                                        const float* __restrict__ hyper_rate, int row_begin, int row_end,
                                                        int K, int ld, double shape_prior, double hyper_shape,
                                                        double hyper_prior_shape, double hyper_prior_rate, double* out,
                                                        int slot_p, int slot_h, const ElboRowConsts cst) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double p_fac = 0.0, p_hyp = 0.0, ent = 0.0;
    const int row = row_begin + (int)wid;
    if (row < row_end) {
        const double lg_hs = cst.lg_hs, psi_hs = cst.psi_hs, lg_sp = cst.lg_sp;
        const double hr = (double)hyper_rate[row];
        const double log_hr = log(hr);
        const double Lh = psi_hs - log_hr;   // E log xi
        const double Eh = hyper_shape / hr;
        const double row_const = shape_prior * Lh - lg_sp;
        const float sp1 = (float)(shape_prior - 1.0), Ehf = (float)Eh;
        for (int k = lane; k < K; k += 32) {
            const float s = shp[(size_t)row * ld + k], r = rte[(size_t)row * ld + k];
            float lg, ps;
            lgamma_digamma_f(s, lg, ps);
            const float lr = logf(r);
            p_fac += row_const + (double)(sp1 * (ps - lr) - Ehf * __fdividef(s, r));
            ent += (double)(s - lr + lg + (1.f - s) * ps);
        }
        if (lane == 0) {
            const double lg_hp = cst.lg_hp;
            p_hyp = hyper_prior_shape * log(hyper_prior_rate) - lg_hp + (hyper_prior_shape - 1.0) * Lh - hyper_prior_rate * Eh;
            ent += hyper_shape - log_hr + lg_hs + (1.0 - hyper_shape) * psi_hs;
        }
    }
    __shared__ double s_red[3][8];
    double v[3] = {p_fac, p_hyp, ent};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
        if (lane == 0) s_red[c][threadIdx.x >> 5] = v[c];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
        atomicAdd(out + (threadIdx.x == 0 ? slot_p : threadIdx.x == 1 ? slot_h : 5), t);
    }
}

}  // namespace pmf

using namespace pmf;

extern "C" {

int pmf_gamma_geomean(const float* d_shp, const float* d_rte, int64_t rows, int32_t K, int32_t ld, float* d_G,
                      void* stream) {
    PMF_REQUIRE(rows >= 0 && K >= 1 && ld >= K && ld % 8 == 0, "bad shape");
    if (rows == 0) return PMF_OK;
    PMF_REQUIRE(d_shp && d_rte && d_G, "NULL table");
    geomean_kernel<<<(unsigned)cdiv(rows * ld, 256), 256, 0, (cudaStream_t)stream>>>(d_shp, d_rte, rows, K, ld, d_G);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

int pmf_gamma_pass_digamma(const pmf_csr* csr, int32_t K, int32_t ld, const float* d_G_oth, const float* d_E_oth,
                           float* d_G_self, float* d_E_self, float* d_shp, float* d_rte, float shape_prior,
                           float rate_prior, const float* d_rate_prior_vec, float* d_hyper_rate, float* d_hyper_mean,
                           float hyper_shape, float hyper_rate_prior, void* d_workspace, void* stream) {
    PMF_REQUIRE(csr != nullptr, "csr is NULL");
    PMF_REQUIRE(K >= 1 && ld >= K && ld % 8 == 0 && ld <= 128, "digamma pass supports K <= 128 (K=%d ld=%d)", K, ld);
    PMF_REQUIRE(d_G_oth && d_E_oth && d_G_self && d_E_self, "NULL table");
    PMF_REQUIRE((d_hyper_rate == nullptr) == (d_hyper_mean == nullptr), "hyper_rate and hyper_mean go together");
    const CsrView c = csr_view(csr);
    PMF_REQUIRE(c.n_partial == 0 || d_workspace != nullptr, "workspace is NULL");
    DigammaArgs a;
    a.seg_desc = c.seg_desc; a.col = c.col; a.val = c.val; a.multi_row = c.multi_row; a.multi_first = c.multi_first;
    a.n_seg = c.n_seg; a.n_multi = c.n_multi; a.seg_len = c.seg_len; a.row_offset = c.row_offset;
    a.K = K; a.ld = ld; a.nvec = ld / 4;
    a.G_oth = d_G_oth; a.E_oth = d_E_oth; a.G_self = d_G_self; a.E_self = d_E_self; a.shp = d_shp; a.rte = d_rte;
    a.shape_prior = shape_prior; a.rate_prior = rate_prior; a.rate_prior_vec = d_rate_prior_vec;
    a.hyper_rate = d_hyper_rate; a.hyper_mean = d_hyper_mean; a.hyper_shape = hyper_shape;
    a.hyper_rate_prior = hyper_rate_prior; a.partial = (float*)d_workspace;
    cudaStream_t s = (cudaStream_t)stream;
    const int V = (a.nvec + 7) / 8;
    if (c.n_seg > 0) {
        const unsigned grid = (unsigned)cdiv((int64_t)c.n_seg * 8, 256);
        switch (V) {
            case 1: digamma_pass_kernel<1><<<grid, 256, 0, s>>>(a); break;
            case 2: digamma_pass_kernel<2><<<grid, 256, 0, s>>>(a); break;
            case 3: digamma_pass_kernel<3><<<grid, 256, 0, s>>>(a); break;
            default: digamma_pass_kernel<4><<<grid, 256, 0, s>>>(a); break;
        }
        PMF_LAUNCH_CHECK();
    }
    if (c.n_multi > 0) {
        switch (V) {
            case 1: digamma_multi_kernel<1><<<c.n_multi, 32, 0, s>>>(a); break;
            case 2: digamma_multi_kernel<2><<<c.n_multi, 32, 0, s>>>(a); break;
            case 3: digamma_multi_kernel<3><<<c.n_multi, 32, 0, s>>>(a); break;
            default: digamma_multi_kernel<4><<<c.n_multi, 32, 0, s>>>(a); break;
        }
        PMF_LAUNCH_CHECK();
    }
    return PMF_OK;
}

int pmf_hpf_elbo(const pmf_csr* by_user, int32_t K, int32_t ld, const float* d_E_theta, const float* d_E_beta,
                 const float* d_G_theta, const float* d_G_beta, const float* d_shp_theta, const float* d_rte_theta,
                 const float* d_shp_beta, const float* d_rte_beta, const float* d_rate_xi, const float* d_rate_eta,
                 int32_t user_begin, int32_t user_end, int32_t item_begin, int32_t item_end, float a, float a_prime,
                 float b_prime, float c, float c_prime, float d_prime, double* d_out6, void* stream) {
    PMF_REQUIRE(by_user != nullptr && d_out6 != nullptr, "NULL argument");
    PMF_REQUIRE(K >= 1 && ld >= K && ld % 8 == 0, "bad shape");
    PMF_REQUIRE(d_E_theta && d_E_beta && d_G_theta && d_G_beta && d_shp_theta && d_rte_theta && d_shp_beta &&
                    d_rte_beta && d_rate_xi && d_rate_eta, "NULL table");
    const CsrView cv = csr_view(by_user);
    cudaStream_t s = (cudaStream_t)stream;
    PMF_CUDA(cudaMemsetAsync(d_out6, 0, 6 * sizeof(double), s));
    ElboArgs e;
    e.row_ptr = cv.row_ptr; e.col = cv.col; e.val = cv.val; e.n_rows = cv.n_rows; e.row_offset = cv.row_offset;
    e.K = K; e.ld = ld; e.nvec = ld / 4;
    e.E_theta = d_E_theta; e.E_beta = d_E_beta; e.G_theta = d_G_theta; e.G_beta = d_G_beta; e.out = d_out6;
    if (cv.nnz > 0 && cv.n_seg > 0) {
        elbo_like_kernel<<<(unsigned)cdiv((int64_t)cv.n_seg * 8, 256), 256, 0, s>>>(e, cv.seg_desc, cv.n_seg);
        PMF_LAUNCH_CHECK();
    }
    auto consts = [](double shape_prior, double hyper_shape, double hyper_prior_shape) {
        ElboRowConsts k;
        double unused;
        lgamma_digamma(hyper_shape, k.lg_hs, k.psi_hs);
        lgamma_digamma(shape_prior, k.lg_sp, unused);
        lgamma_digamma(hyper_prior_shape, k.lg_hp, unused);
        return k;
    };
    if (user_end > user_begin) {
        const double hs = (double)a_prime + K * (double)a;
        elbo_rows_kernel<<<(unsigned)cdiv((int64_t)(user_end - user_begin) * 32, 256), 256, 0, s>>>(
            d_shp_theta, d_rte_theta, d_rate_xi, user_begin, user_end, K, ld, (double)a, hs,
            (double)a_prime, (double)b_prime, d_out6, 1, 3, consts((double)a, hs, (double)a_prime));
        PMF_LAUNCH_CHECK();
    }
    if (item_end > item_begin) {
        const double hs = (double)c_prime + K * (double)c;
        elbo_rows_kernel<<<(unsigned)cdiv((int64_t)(item_end - item_begin) * 32, 256), 256, 0, s>>>(
            d_shp_beta, d_rte_beta, d_rate_eta, item_begin, item_end, K, ld, (double)c, hs,
            (double)c_prime, (double)d_prime, d_out6, 2, 4, consts((double)c, hs, (double)c_prime));
        PMF_LAUNCH_CHECK();
    }
    return PMF_OK;
}

}  // extern "C"
