// a3/a4: the fused Gamma-Poisson row pass shared by Poisson MF and HPF-CAVI.
//
// Reference loops replaced: poisson_mf_cavi.py:135-164, :173-194 (+ :167, :197) and
// hpf_cavi.py:126-151, :162-185 (+ :153, :158-159, :187, :192-193).
//
// One kernel does, per observed rating, the SDDMM  rate = <E_self[row], E_oth[col]>, the
// allocation  (x / rate) * E_oth[col] * E_self[row]  and the segmented row sums of the allocation
// and of E_oth[col]; the Gamma shape/rate update, E = shape/rate and the HPF hyper-rate update run in
// the same kernel's epilogue.  Nothing per-observation is written to HBM.
//
// Mapping: a GROUP of G lanes owns one segment (<= seg_len observations of one row).  Each lane keeps
// V float4 slices of the K-wide rows in registers (slice index = lane + v*G, so the G lanes of a load
// instruction read 16*G contiguous bytes).  Per G observations the group reads col/val with one
// coalesced load per lane and broadcasts them with shuffles; U gathered rows are kept in flight per
// lane.  The dot product is reduced over the G lanes with log2(G) xor-shuffles.  Rows cut into several
// segments write their partial sums to a scratch buffer and a second, tiny kernel combines them in
// segment order (deterministic, no atomics).
//
// HBM-bound: 4*ld + 8 bytes per observation, 16*ld + 4 per row.
//
// MODE (template): 0 plain Poisson MF, 1 HPF (rate prior per row + hyper-rate update in the epilogue), 2 extended
// Poisson MF (poisson_mf_extended_cavi.py:110-216): per-row scalars phi / psi scale the rate sums, the allocation
// divides by the UNCLAMPED dot product (the reference's clamped rate_est :137-138 is never used), the scalar's rate
// uses the row's NEW mean (in-row Gauss-Seidel :160-164), and rows without observations get prior shape/rate but keep
// their expectations (:112-118 `continue`).
#include "common.cuh"

namespace pmf {

extern thread_local int g_tune_topn_growth;   // topn_fused.cu

struct GammaArgs {
    const int4* seg_desc;   // segments in processing order: {row, start, end, partial slot or -1}
    const int32_t* col;
    const float* val;
    const int32_t *multi_row, *multi_first;
    int32_t n_seg, n_multi, seg_len, row_offset, K, ld, nvec;
    int64_t nnz_hint;   // observations of this rating list (dispatch heuristics only)
    int64_t gather_bytes;   // size of the part of E_oth this pass can touch (dispatch heuristics only)
    const float* E_oth;
    float* E_self;
    float* shp;
    float* rte;
    float shape_prior, rate_prior;
    const float* rate_prior_vec;
    float* hyper_rate;
    float* hyper_mean;
    float hyper_shape, hyper_rate_prior;
    float* partial;  // [n_partial][2*ld]: sum (x/rate) E_oth | sum E_oth
    // MODE 2: scale_oth = the other side's scalar means (psi for the user pass); the row's own scalar goes to scale_shp
    // (shape), hyper_rate (rate) and hyper_mean (mean); partial_x[n_partial] holds the segments' rating sums
    const float* scale_oth;
    float* scale_shp;
    float* partial_x;
    // block -> chunk of the longest-first segment list: chunk = (blockIdx.x * block_stride) % gridDim.x with
    // gcd(block_stride, gridDim.x) = 1.  stride 1 = longest first.
    uint32_t block_stride;
    // Tiled passes (the gathered table is visited one L2-sized row range -- "tile" -- at a time, and, on several GPUs,
    // one user range per GPU): the row sums [sum (x/rate) E_oth | sum E_oth] accumulate across the tiles in
    // acc[(R - acc_row_base)][2*ld].  acc_in: add the sums accumulated so far; acc_out: store the running sums instead of
    // finishing the row (the last tile -- or, across GPUs, gamma_combine_kernel -- applies the Gamma update).
    float* acc;
    int32_t acc_in, acc_out, acc_row_base;
    float* mc_E;   // NVSwitch multicast alias of E_self (all replicas at once, multimem.st) or NULL
};

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// Blackwell issues two FP32 multiply-adds per instruction on register pairs (FFMA2 / FADD2): the three 4-wide updates of
// a rating (dot product, weighted sum, plain sum) take 6 instructions instead of 12.  Each half is an ordinary IEEE
// operation, so the results are those of the scalar forms (the dot product pairs its terms differently).
#ifndef PMF_PACKED_FP32
#define PMF_PACKED_FP32 1
#endif
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& o) {   // acc += w * o
#if PMF_PACKED_FP32
    const float2 ww = make_float2(w, w);
    const float2 lo = __ffma2_rn(ww, make_float2(o.x, o.y), make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(ww, make_float2(o.z, o.w), make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
#else
    acc.x = fmaf(w, o.x, acc.x); acc.y = fmaf(w, o.y, acc.y); acc.z = fmaf(w, o.z, acc.z); acc.w = fmaf(w, o.w, acc.w);
#endif
}
__device__ __forceinline__ void add4(float4& acc, const float4& o) {            // acc += o
#if PMF_PACKED_FP32
    const float2 lo = __fadd2_rn(make_float2(acc.x, acc.y), make_float2(o.x, o.y));
    const float2 hi = __fadd2_rn(make_float2(acc.z, acc.w), make_float2(o.z, o.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
#else
    acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
#endif
}
template <int V>
__device__ __forceinline__ float dot4(const float4 (&s)[V], const float4 (&o)[V]) {   // sum_v <s[v], o[v]>
#if PMF_PACKED_FP32
    float2 t = make_float2(0.f, 0.f);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        t = __ffma2_rn(make_float2(s[v].x, s[v].y), make_float2(o[v].x, o[v].y), t);
        t = __ffma2_rn(make_float2(s[v].z, s[v].w), make_float2(o[v].z, o[v].w), t);
    }
    return t.x + t.y;
#else
    float d = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        d = fmaf(s[v].x, o[v].x, d); d = fmaf(s[v].y, o[v].y, d); d = fmaf(s[v].z, o[v].z, d); d = fmaf(s[v].w, o[v].w, d);
    }
    return d;
#endif
}

// Sum over the G lanes of a group.  `mask` names the participating lanes: the full warp inside the
// warp-uniform main loop, only the group's own lanes in the (group-divergent) epilogues.
template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned mask = 0xffffffffu) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

template <int G>
__device__ __forceinline__ unsigned group_mask(int lane) {
    return G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
}

// Gamma update for global row R from the row's complete sums.  Executed by the G lanes of a group.
template <int G, int V, int MODE>
__device__ __forceinline__ void gamma_row_update(const GammaArgs& a, int R, int gl, unsigned gmask,
                                                 const float4 (&self)[V], const float4 (&sa)[V],
                                                 const float4 (&sb)[V], float sx = 0.f, bool empty = false) {
    const float rp = a.rate_prior_vec ? a.rate_prior_vec[R] : a.rate_prior;
    float esum = 0.f;
    const size_t rowoff = (size_t)R * a.ld;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int idx = gl + v * G;
        if (idx < a.nvec) {
            const int k0 = idx * 4;
            float4 s, r, e;
            s.x = a.shape_prior + self[v].x * sa[v].x;  r.x = rp + sb[v].x;  e.x = s.x / r.x;
            s.y = a.shape_prior + self[v].y * sa[v].y;  r.y = rp + sb[v].y;  e.y = s.y / r.y;
            s.z = a.shape_prior + self[v].z * sa[v].z;  r.z = rp + sb[v].z;  e.z = s.z / r.z;
            s.w = a.shape_prior + self[v].w * sa[v].w;  r.w = rp + sb[v].w;  e.w = s.w / r.w;
            // padding columns K..ld-1 stay (shape 0, rate 1, mean 0) so they never reach a dot product
            if (k0 + 0 >= a.K) { s.x = 0.f; r.x = 1.f; e.x = 0.f; }
            if (k0 + 1 >= a.K) { s.y = 0.f; r.y = 1.f; e.y = 0.f; }
            if (k0 + 2 >= a.K) { s.z = 0.f; r.z = 1.f; e.z = 0.f; }
            if (k0 + 3 >= a.K) { s.w = 0.f; r.w = 1.f; e.w = 0.f; }
            if (a.shp) *reinterpret_cast<float4*>(a.shp + rowoff + k0) = s;
            if (a.rte) *reinterpret_cast<float4*>(a.rte + rowoff + k0) = r;
            if constexpr (MODE == 2) {
                // scalar rate: sum_t s_oth[c_t] (E_oth[c_t] . E_new) = E_new . sum_t s_oth[c_t] E_oth[c_t] = E_new . sb
                esum += (e.x * sb[v].x + e.y * sb[v].y) + (e.z * sb[v].z + e.w * sb[v].w);
                if (empty) continue;   // no observations: prior shape/rate, expectations untouched (:112-118)
            }
            *reinterpret_cast<float4*>(a.E_self + rowoff + k0) = e;
            if (a.mc_E)   // one store, replicated to every GPU of the multicast group by the NVSwitch
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                             ::"l"(a.mc_E + rowoff + k0), "f"(e.x), "f"(e.y), "f"(e.z), "f"(e.w) : "memory");
            if constexpr (MODE != 2) esum += (e.x + e.y) + (e.z + e.w);
        }
    }
    if constexpr (MODE == 2) {
        esum = group_sum<G>(esum, gmask);
        if (gl == 0) {
            const float s_shp = a.shape_prior + sx;      // :153  a_phi = a0 + sum_t x_t
            const float s_rte = a.rate_prior + esum;     // :164  b_phi = b0 + sum_t psi_t (beta_t . theta_new)
            a.scale_shp[R] = s_shp;
            a.hyper_rate[R] = s_rte;
            if (!empty) a.hyper_mean[R] = s_shp / s_rte;
        }
    }
    if constexpr (MODE == 1) {
        esum = group_sum<G>(esum, gmask);
        if (gl == 0) {
            const float hr = a.hyper_rate_prior + esum;   // hpf_cavi.py:158 / :192
            const float hm = a.hyper_shape / hr;          // hpf_cavi.py:94-95
            a.hyper_rate[R] = hr;
            a.hyper_mean[R] = hm;   // only ever read by the row's owner: not replicated
        }
    }
}

// A row's sums over the current tile are complete (held by the G lanes of a group): fold them into the sums of the
// earlier tiles and either park the running sums (acc_out) or finish the row.
template <int G, int V, int MODE>
__device__ __forceinline__ void finish_row(const GammaArgs& a, int R, int gl, unsigned gmask, const float4 (&self)[V],
                                           float4 (&sa)[V], float4 (&sb)[V], float sx, bool empty) {
    if constexpr (MODE != 2) {
        if (a.acc) {
            if (a.acc_in && a.acc_out && empty) return;   // nothing of this row in this tile: running sums unchanged
            float* pr = a.acc + (size_t)(R - a.acc_row_base) * 2 * a.ld;
            if (a.acc_in) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int idx = gl + v * G;
                    if (idx < a.nvec) {
                        const float4 pa = *reinterpret_cast<const float4*>(pr + idx * 4);
                        const float4 pb = *reinterpret_cast<const float4*>(pr + a.ld + idx * 4);
                        sa[v].x = pa.x + sa[v].x; sa[v].y = pa.y + sa[v].y; sa[v].z = pa.z + sa[v].z; sa[v].w = pa.w + sa[v].w;
                        sb[v].x = pb.x + sb[v].x; sb[v].y = pb.y + sb[v].y; sb[v].z = pb.z + sb[v].z; sb[v].w = pb.w + sb[v].w;
                    }
                }
            }
            if (a.acc_out) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int idx = gl + v * G;
                    if (idx < a.nvec) {
                        *reinterpret_cast<float4*>(pr + idx * 4) = sa[v];
                        *reinterpret_cast<float4*>(pr + a.ld + idx * 4) = sb[v];
                    }
                }
                return;
            }
        }
    }
    gamma_row_update<G, V, MODE>(a, R, gl, gmask, self, sa, sb, sx, empty);
}

template <int G, int V, int U, int MODE, bool CHUNK_REDUCE>
__global__ void __launch_bounds__(256) gamma_pass_kernel(const GammaArgs a) {
    static_assert(G == 4 || G == 8 || G == 16 || G == 32, "group size");
    static_assert(G % U == 0, "U must divide G");
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const uint32_t chunk = (uint32_t)(((uint64_t)blockIdx.x * a.block_stride) % gridDim.x);
    const int64_t gid = ((int64_t)chunk * blockDim.x + threadIdx.x) / G;
    const bool has = gid < a.n_seg;
    int row = 0, p = 0, end = 0, pidx = -1;
    if (has) {
        // segments are visited longest first, so the groups sharing a warp run (nearly) equal trip counts
        const int4 d = __ldg(a.seg_desc + gid);
        row = d.x; p = d.y; end = d.z; pidx = d.w;
    }
    const int R = a.row_offset + row;
    float4 self[V], sa[V], sb[V];
    float sx = 0.f;   // MODE 2: sum of the segment's ratings (every lane of the group holds the same value)
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int idx = gl + v * G;
        self[v] = (has && idx < a.nvec) ? *reinterpret_cast<const float4*>(a.E_self + (size_t)R * a.ld + idx * 4) : f4_zero();
        sa[v] = f4_zero();
        sb[v] = f4_zero();
    }
    const bool empty = end == p;
    // warp-uniform trip count so that full-mask shuffles are legal; short groups idle on predicates
    const int maxlen = __reduce_max_sync(0xffffffffu, end - p);
    // column ids / ratings are fetched one chunk ahead, so the dependent chain per chunk is one memory
    // latency (the gathered rows), not two (ids, then rows)
    int c_nxt = (p + gl < end) ? __ldg(a.col + p + gl) : 0;
    float x_nxt = (p + gl < end) ? __ldg(a.val + p + gl) : 0.f;
    for (int base = 0; base < maxlen; base += G) {
        const int c_l = c_nxt;
        const float x_l = x_nxt;
        const int qn = p + base + G + gl;
        c_nxt = qn < end ? __ldg(a.col + qn) : 0;
        x_nxt = qn < end ? __ldg(a.val + qn) : 0.f;
        const int rem = end - (p + base);
#pragma unroll
        for (int j0 = 0; j0 < G; j0 += U) {
            if (U < G && j0 >= maxlen - base) break;   // warp-uniform: nobody in the warp has observations left
            float4 o[U][V];
            float so[U];   // MODE 2: the gathered rows' scalars
#pragma unroll
            for (int jj = 0; jj < U; ++jj) {
                const int j = j0 + jj;
                const int c = __shfl_sync(0xffffffffu, c_l, j, G);
                const float* rowp = a.E_oth + (size_t)c * a.ld;
                if constexpr (MODE == 2) so[jj] = j < rem ? __ldg(a.scale_oth + c) : 0.f;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const int idx = gl + v * G;
                    o[jj][v] = (j < rem && idx < a.nvec) ? ldg_f4(rowp + idx * 4) : f4_zero();
                }
            }
            if constexpr (CHUNK_REDUCE) {
                // The whole chunk is in registers: reduce its G dot products TOGETHER.  A butterfly that halves the number
                // of values per step leaves lane j with the complete dot product of rating j after G-1 shuffles (instead
                // of log2(G) per rating), lane j -- which also loaded rating j's value -- does the ONE division, and the
                // weights are handed back with one shuffle per rating.
                float d[G];
#pragma unroll
                for (int jj = 0; jj < G; ++jj) d[jj] = dot4<V>(self, o[jj]);
#pragma unroll
                for (int h = G / 2; h > 0; h >>= 1) {
                    const bool up = (gl & h) != 0;
#pragma unroll
                    for (int i = 0; i < h; ++i) {
                        const float send = up ? d[i] : d[i + h];
                        const float keep = up ? d[i + h] : d[i];
                        d[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
                    }
                }
                float w_l;   // weight of rating gl of the chunk (x_l = 0 and all-zero rows past the segment end give 0)
                if constexpr (MODE == 2) {
                    w_l = gl < rem ? x_l / d[0] : 0.f;   // poisson_mf_extended_cavi.py:142: the raw dot product
                    sx += x_l;                           // per-lane share; summed over the group after the loop
                } else {
                    w_l = x_l / fmaxf(d[0], 1e-10f);     // poisson_mf_cavi.py:153,157
                }
#pragma unroll
                for (int jj = 0; jj < G; ++jj) {
                    const float w = __shfl_sync(0xffffffffu, w_l, jj, G);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        fma4(sa[v], w, o[jj][v]);
                        if constexpr (MODE == 2) fma4(sb[v], so[jj], o[jj][v]);
                        else add4(sb[v], o[jj][v]);
                    }
                }
            } else {
#pragma unroll
            for (int jj = 0; jj < U; ++jj) {
                const int j = j0 + jj;
                const float x = __shfl_sync(0xffffffffu, x_l, j, G);
                const float d = group_sum<G>(dot4<V>(self, o[jj]));
                if constexpr (MODE == 2) {
                    // poisson_mf_extended_cavi.py:142 divides by the raw dot product; lanes past the segment end add 0
                    const float w = j < rem ? x / d : 0.f;
                    sx += x;
#pragma unroll
                    for (int v = 0; v < V; ++v) { fma4(sa[v], w, o[jj][v]); fma4(sb[v], so[jj], o[jj][v]); }
                    continue;
                }
                const float w = x / fmaxf(d, 1e-10f);  // poisson_mf_cavi.py:153,157 (x = 0 past the segment end)
#pragma unroll
                for (int v = 0; v < V; ++v) { fma4(sa[v], w, o[jj][v]); add4(sb[v], o[jj][v]); }
            }
            }   // per-rating reduction
        }
    }
    if constexpr (MODE == 2 && CHUNK_REDUCE) sx = group_sum<G>(sx);   // full mask: the warp is still converged here
    if (!has) return;
    if (pidx < 0) {  // the whole row lives in this segment: finish it here
        finish_row<G, V, MODE>(a, R, gl, group_mask<G>(lane), self, sa, sb, sx, empty);
    } else {
        if constexpr (MODE == 2) if (gl == 0) a.partial_x[pidx] = sx;
        float* dst = a.partial + (size_t)pidx * 2 * a.ld;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int idx = gl + v * G;
            if (idx < a.nvec) {
                *reinterpret_cast<float4*>(dst + idx * 4) = sa[v];
                *reinterpret_cast<float4*>(dst + a.ld + idx * 4) = sb[v];
            }
        }
    }
}

// Rows cut into several segments: one CTA per row.  Its 256/G lane groups each sum every (256/G)-th
// partial (independent loads in flight, so a row with hundreds of partials costs a few round trips, not
// hundreds), park their sums in shared memory, and group 0 folds them in group order and applies the
// same row update.  Deterministic; no atomics.
template <int G, int V, int MODE>
__global__ void __launch_bounds__(256) gamma_multi_kernel(const GammaArgs a) {
    constexpr int NG = 256 / G;
    extern __shared__ float s_part[];   // [NG][2*ld] (+ [NG] rating sums, MODE 2)
    const int lane = threadIdx.x & 31;
    const int gl = threadIdx.x & (G - 1);
    const int grp = threadIdx.x / G;
    const int row = a.multi_row[blockIdx.x];
    const int first = a.multi_first[blockIdx.x], last = a.multi_first[blockIdx.x + 1];
    const int R = a.row_offset + row;
    float4 sa[V], sb[V];
    float sx = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) { sa[v] = f4_zero(); sb[v] = f4_zero(); }
    for (int q = first + grp; q < last; q += NG) {
        if constexpr (MODE == 2) sx += a.partial_x[q];
        const float* src = a.partial + (size_t)q * 2 * a.ld;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int idx = gl + v * G;
            if (idx < a.nvec) {
                const float4 pa = ld_stream_f4(src + idx * 4);
                const float4 pb = ld_stream_f4(src + a.ld + idx * 4);
                sa[v].x += pa.x; sa[v].y += pa.y; sa[v].z += pa.z; sa[v].w += pa.w;
                sb[v].x += pb.x; sb[v].y += pb.y; sb[v].z += pb.z; sb[v].w += pb.w;
            }
        }
    }
    const int used = min(NG, last - first);   // groups that saw at least one partial
    if (grp > 0 && grp < used) {
        if constexpr (MODE == 2) if (gl == 0) s_part[(size_t)NG * 2 * a.ld + grp] = sx;
        float* dst = s_part + (size_t)grp * 2 * a.ld;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int idx = gl + v * G;
            if (idx < a.nvec) {
                *reinterpret_cast<float4*>(dst + idx * 4) = sa[v];
                *reinterpret_cast<float4*>(dst + a.ld + idx * 4) = sb[v];
            }
        }
    }
    __syncthreads();
    if (grp != 0) return;
    float4 self[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int idx = gl + v * G;
        self[v] = (idx < a.nvec) ? *reinterpret_cast<const float4*>(a.E_self + (size_t)R * a.ld + idx * 4) : f4_zero();
    }
    for (int g = 1; g < used; ++g) {
        if constexpr (MODE == 2) sx += s_part[(size_t)NG * 2 * a.ld + g];
        const float* src = s_part + (size_t)g * 2 * a.ld;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int idx = gl + v * G;
            if (idx < a.nvec) {
                const float4 pa = *reinterpret_cast<const float4*>(src + idx * 4);
                const float4 pb = *reinterpret_cast<const float4*>(src + a.ld + idx * 4);
                sa[v].x += pa.x; sa[v].y += pa.y; sa[v].z += pa.z; sa[v].w += pa.w;
                sb[v].x += pb.x; sb[v].y += pb.y; sb[v].z += pb.z; sb[v].w += pb.w;
            }
        }
    }
    finish_row<G, V, MODE>(a, R, gl, group_mask<G>(lane), self, sa, sb, sx, false);
}

// Multi-GPU item pass, last step: every rank has parked its user range's row sums in its own `acc` table; the owner of
// row R adds the ranks' sums -- ONE multimem.ld_reduce per 16 bytes, the addition is done by the NVSwitch -- applies the
// Gamma update and pushes the new row to every replica with multimem.st.  With mc_acc == NULL the sums in `acc` are
// already complete (NCCL all-reduce, or a single GPU) and are read with plain loads.  One lane group per row.
// (staged form) n_src > 0: the ranks' sums of the row were copied by the copy engines into `stage`
// ([n_src][src_stride floats], slot s = source rank s, row q of a slot = this owner's q-th row); they are added in rank
// order with the local sums in slot `self_rank`'s place, so the result does not depend on who owns the row.
struct StageArgs {
    const float* stage;
    int64_t src_stride;   // floats between two sources' slots
    int32_t n_src, self_rank, row0;   // row0: slot row of row_begin
};

template <int G, int V, int MODE>
__global__ void __launch_bounds__(256) gamma_combine_kernel(const GammaArgs a, const float* mc_acc, int row_begin, int row_end,
                                                            const StageArgs st) {
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G - 1);
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int R = row_begin + (int)gid;
    if (gid >= row_end - row_begin) return;
    float4 self[V], sa[V], sb[V];
    const size_t poff = (size_t)(R - a.acc_row_base) * 2 * a.ld;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int idx = gl + v * G;
        self[v] = sa[v] = sb[v] = f4_zero();
        if (idx < a.nvec) {
            self[v] = *reinterpret_cast<const float4*>(a.E_self + (size_t)R * a.ld + idx * 4);
            if (st.n_src > 0) {
                const size_t q = (size_t)(st.row0 + (R - row_begin)) * 2 * a.ld;
                for (int src = 0; src < st.n_src; ++src) {
                    const float* base = src == st.self_rank ? a.acc + poff : st.stage + (size_t)src * st.src_stride + q;
                    const float4 pa = ld_stream_f4(base + idx * 4), pb = ld_stream_f4(base + a.ld + idx * 4);
                    sa[v].x += pa.x; sa[v].y += pa.y; sa[v].z += pa.z; sa[v].w += pa.w;
                    sb[v].x += pb.x; sb[v].y += pb.y; sb[v].z += pb.z; sb[v].w += pb.w;
                }
            } else if (mc_acc) {
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(sa[v].x), "=f"(sa[v].y), "=f"(sa[v].z), "=f"(sa[v].w) : "l"(mc_acc + poff + idx * 4) : "memory");
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(sb[v].x), "=f"(sb[v].y), "=f"(sb[v].z), "=f"(sb[v].w) : "l"(mc_acc + poff + a.ld + idx * 4) : "memory");
            } else {
                sa[v] = ld_stream_f4(a.acc + poff + idx * 4);
                sb[v] = ld_stream_f4(a.acc + poff + a.ld + idx * 4);
            }
        }
    }
    gamma_row_update<G, V, MODE>(a, R, gl, group_mask<G>(lane), self, sa, sb, 0.f, false);
}

static thread_local int g_tune_interleave = 0;   // 1 = golden-ratio block interleave of the longest-first order (experiment)
static thread_local int g_tune_group = 0;   // 0 = auto; else forced G for nvec <= 16 (8 or 16)
static thread_local int g_tune_unroll = 0;  // 0 = auto; else forced U
static thread_local int g_tune_chunk_reduce = -1;  // -1 = auto (by gathered-table size), 0 = per-rating dot reduction, 1 = chunk-wide butterfly

static uint32_t gcd_u32(uint32_t x, uint32_t y) {
    while (y) { const uint32_t t = x % y; x = y; y = t; }
    return x;
}

template <int G, int V, int U>
static int launch_gamma(const GammaArgs& a_in, int mode, cudaStream_t s) {
    GammaArgs a = a_in;
    if (a.n_seg > 0) {
        const unsigned grid = (unsigned)cdiv((int64_t)a.n_seg * G, 256);
        const bool interleave = g_tune_interleave > 0;   // measured slower at N=1 and N=8 (profiles/README.md): off
        a.block_stride = 1;
        if (g_tune_interleave == 2 && grid > 2) {
            a.block_stride = grid - 1;   // shortest segments first
        } else if (interleave && grid > 2) {
            uint32_t st = (uint32_t)(0.6180339887 * grid) | 1u;   // golden-ratio stride: even spread of every length class
            while (gcd_u32(st, grid) != 1) st += 2;
            a.block_stride = st % grid;
        }
        // chunk-wide dot reduction needs the whole chunk in registers (U == G); gamma_chunk_reduce=0 keeps the
        // per-rating reduction for comparison
        // Measured on C5 (profiles/README.md): with the gathered table around L2 size (user pass, 128 MB) the pass is
        // bound by instruction issue and gather latency and the chunk-wide reduction is 21 % faster; with a table far
        // beyond L2 (item pass, 512 MB) the pass is DRAM-bound and prefers the per-rating form, whose accumulation of
        // early rows overlaps the wait for late ones (12 % faster).  gamma_chunk_reduce: -1 auto, 0 off, 1 on.
        constexpr bool kCan = U == G;
        const bool cr = kCan && (g_tune_chunk_reduce < 0 ? a.gather_bytes <= (int64_t)256 << 20 : g_tune_chunk_reduce != 0);
        if (cr) {
            if (mode == 2) gamma_pass_kernel<G, V, U, 2, kCan><<<grid, 256, 0, s>>>(a);
            else if (mode == 1) gamma_pass_kernel<G, V, U, 1, kCan><<<grid, 256, 0, s>>>(a);
            else gamma_pass_kernel<G, V, U, 0, kCan><<<grid, 256, 0, s>>>(a);
        } else {
            if (mode == 2) gamma_pass_kernel<G, V, U, 2, false><<<grid, 256, 0, s>>>(a);
            else if (mode == 1) gamma_pass_kernel<G, V, U, 1, false><<<grid, 256, 0, s>>>(a);
            else gamma_pass_kernel<G, V, U, 0, false><<<grid, 256, 0, s>>>(a);
        }
        PMF_LAUNCH_CHECK();
    }
    if (a.n_multi > 0) {
        const unsigned grid = (unsigned)a.n_multi;
        const size_t smem = (size_t)(256 / G) * (2 * a.ld + 1) * sizeof(float);   // <= 33 KB for every (G, ld) dispatched
        if (mode == 2) gamma_multi_kernel<G, V, 2><<<grid, 256, smem, s>>>(a);
        else if (mode == 1) gamma_multi_kernel<G, V, 1><<<grid, 256, smem, s>>>(a);
        else gamma_multi_kernel<G, V, 0><<<grid, 256, smem, s>>>(a);
        PMF_LAUNCH_CHECK();
    }
    return PMF_OK;
}

static void fill_csr_args(GammaArgs& a, const CsrView& c, int32_t K, int32_t ld) {
    a.seg_desc = c.seg_desc;
    a.col = c.col; a.val = c.val; a.multi_row = c.multi_row; a.multi_first = c.multi_first;
    a.n_seg = c.n_seg; a.n_multi = c.n_multi; a.seg_len = c.seg_len; a.row_offset = c.row_offset;
    a.K = K; a.ld = ld; a.nvec = ld / 4; a.nnz_hint = c.nnz;
    a.gather_bytes = (int64_t)(c.n_cols - c.col_lo) * ld * 4;   // the span of E_oth this list's ratings point into
    a.scale_oth = nullptr; a.scale_shp = nullptr; a.partial_x = nullptr;
    a.acc = nullptr; a.acc_in = a.acc_out = 0; a.acc_row_base = 0;
}

static int dispatch_gamma(const GammaArgs& a, int mode, cudaStream_t s) {
    const int nv = a.nvec;
    // U = gathered rows in flight per lane.  Long segments (C5: 50-200 ratings per row) want 8; when the average segment
    // is shorter than a lane group's chunk of 8, U = 2 wins: 40 fewer registers (3 CTAs per SM instead of 2, the pass is
    // latency-bound on short rows) and the chunk loop stops as soon as no group of the warp has ratings left
    // (C2/C3: 0.31 -> 0.27 ms per sweep, profiles/README.md).
    const bool short_segments = g_tune_unroll == 0 && a.n_seg > 0 && a.nnz_hint / a.n_seg < 8;
    if (nv <= 4) return launch_gamma<4, 1, 4>(a, mode, s);
    if (nv <= 8) {
        if (g_tune_unroll == 2 || short_segments) return launch_gamma<8, 1, 2>(a, mode, s);
        return g_tune_unroll == 4 ? launch_gamma<8, 1, 4>(a, mode, s) : launch_gamma<8, 1, 8>(a, mode, s);
    }
    if (nv <= 16) {
        if (g_tune_group == 16) return g_tune_unroll == 8 ? launch_gamma<16, 1, 8>(a, mode, s) : launch_gamma<16, 1, 4>(a, mode, s);
        if (g_tune_unroll == 4) return launch_gamma<8, 2, 4>(a, mode, s);
        if (g_tune_unroll == 2 || short_segments) return launch_gamma<8, 2, 2>(a, mode, s);
        return launch_gamma<8, 2, 8>(a, mode, s);   // measured best on C5 (profiles/README.md)
    }
    if (nv <= 24) return launch_gamma<8, 3, 2>(a, mode, s);
    if (nv <= 32) return launch_gamma<8, 4, 2>(a, mode, s);
    return launch_gamma<16, 4, 2>(a, mode, s);
}

template <int G, int V>
static int launch_combine(const GammaArgs& a, const float* mc_acc, int row_begin, int row_end, int mode, cudaStream_t s,
                          const StageArgs& st) {
    const unsigned grid = (unsigned)cdiv((int64_t)(row_end - row_begin) * G, 256);
    if (mode == 1) gamma_combine_kernel<G, V, 1><<<grid, 256, 0, s>>>(a, mc_acc, row_begin, row_end, st);
    else gamma_combine_kernel<G, V, 0><<<grid, 256, 0, s>>>(a, mc_acc, row_begin, row_end, st);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

}  // namespace pmf

using namespace pmf;

extern "C" {

int pmf_tune(const char* key, int value) {
    PMF_REQUIRE(key != nullptr, "key is NULL");
    if (!strcmp(key, "gamma_group")) g_tune_group = value;
    else if (!strcmp(key, "gamma_interleave")) g_tune_interleave = value;
    else if (!strcmp(key, "gamma_unroll")) g_tune_unroll = value;
    else if (!strcmp(key, "gamma_chunk_reduce")) g_tune_chunk_reduce = value;
    else if (!strcmp(key, "topn_growth")) g_tune_topn_growth = value;
    else { set_error("unknown tuning key '%s'", key); return PMF_EINVAL; }
    return PMF_OK;
}

int64_t pmf_gamma_pass_workspace_bytes(const pmf_csr* csr, int32_t ld) {
    if (!csr || ld <= 0) return -1;
    const CsrView c = csr_view(csr);
    // per partial: 2*ld floats of row sums + one rating sum (extended model only)
    const int64_t b = (int64_t)c.n_partial * (2 * ld + 1) * (int64_t)sizeof(float);
    return b > 0 ? b : 16;
}

int pmf_gamma_pass(const pmf_csr* csr, int32_t K, int32_t ld, const float* d_E_oth, float* d_E_self, float* d_shp,
                   float* d_rte, float shape_prior, float rate_prior, const float* d_rate_prior_vec,
                   float* d_hyper_rate, float* d_hyper_mean, float hyper_shape, float hyper_rate_prior,
                   void* d_workspace, void* stream) {
    return pmf_gamma_pass_acc(csr, K, ld, d_E_oth, d_E_self, d_shp, d_rte, shape_prior, rate_prior, d_rate_prior_vec,
                              d_hyper_rate, d_hyper_mean, hyper_shape, hyper_rate_prior, d_workspace, nullptr, 0, 0,
                              stream);
}

static int check_gamma_tables(int32_t K, int32_t ld, const void* d_E_oth, const void* d_E_self, const void* d_hyper_rate,
                              const void* d_hyper_mean) {
    PMF_REQUIRE(K >= 1 && ld >= K && ld % 8 == 0 && ld <= 256, "need 1 <= K <= ld <= 256 and ld %% 8 == 0 (K=%d ld=%d)", K, ld);
    PMF_REQUIRE(d_E_oth && d_E_self, "factor tables are NULL");
    PMF_REQUIRE((d_hyper_rate == nullptr) == (d_hyper_mean == nullptr), "hyper_rate and hyper_mean go together");
    return PMF_OK;
}

int pmf_gamma_pass_acc(const pmf_csr* csr, int32_t K, int32_t ld, const float* d_E_oth, float* d_E_self, float* d_shp,
                       float* d_rte, float shape_prior, float rate_prior, const float* d_rate_prior_vec,
                       float* d_hyper_rate, float* d_hyper_mean, float hyper_shape, float hyper_rate_prior,
                       void* d_workspace, float* d_acc, int32_t acc_row_base, int32_t acc_flags, void* stream) {
    PMF_REQUIRE(csr != nullptr, "csr is NULL");
    PMF_TRY(check_gamma_tables(K, ld, d_E_oth, d_E_self, d_hyper_rate, d_hyper_mean));
    PMF_REQUIRE((acc_flags & ~3) == 0 && (acc_flags == 0 || d_acc != nullptr), "acc_flags=%d needs d_acc", acc_flags);
    const CsrView c = csr_view(csr);
    PMF_REQUIRE(c.n_partial == 0 || d_workspace != nullptr, "workspace is NULL but %d partial sums are needed", c.n_partial);
    PMF_REQUIRE(acc_flags == 0 || acc_row_base <= c.row_offset, "acc_row_base=%d beyond the first row %d", acc_row_base, c.row_offset);
    GammaArgs a;
    fill_csr_args(a, c, K, ld);
    a.E_oth = d_E_oth; a.E_self = d_E_self; a.shp = d_shp; a.rte = d_rte;
    a.shape_prior = shape_prior; a.rate_prior = rate_prior; a.rate_prior_vec = d_rate_prior_vec;
    a.hyper_rate = d_hyper_rate; a.hyper_mean = d_hyper_mean; a.hyper_shape = hyper_shape;
    a.hyper_rate_prior = hyper_rate_prior; a.partial = (float*)d_workspace;
    a.acc = acc_flags ? d_acc : nullptr;
    a.acc_in = (acc_flags & PMF_ACC_IN) != 0; a.acc_out = (acc_flags & PMF_ACC_OUT) != 0; a.acc_row_base = acc_row_base;
    a.mc_E = nullptr;
    return dispatch_gamma(a, d_hyper_rate != nullptr ? 1 : 0, (cudaStream_t)stream);
}

static int combine_impl(int32_t row_begin, int32_t row_end, int32_t K, int32_t ld, const float* d_acc,
                        const float* d_mc_acc, int32_t acc_row_base, float* d_E_self, float* d_mc_E_self, float* d_shp,
                        float* d_rte, float shape_prior, float rate_prior, const float* d_rate_prior_vec,
                        float* d_hyper_rate, float* d_hyper_mean, float hyper_shape, float hyper_rate_prior,
                        const StageArgs& st, void* stream) {
    PMF_TRY(check_gamma_tables(K, ld, d_acc, d_E_self, d_hyper_rate, d_hyper_mean));
    PMF_REQUIRE(0 <= acc_row_base && acc_row_base <= row_begin && row_begin <= row_end, "bad row range [%d,%d) base %d",
                row_begin, row_end, acc_row_base);
    if (row_end == row_begin) return PMF_OK;
    GammaArgs a = {};
    a.K = K; a.ld = ld; a.nvec = ld / 4;
    a.E_self = d_E_self; a.shp = d_shp; a.rte = d_rte;
    a.shape_prior = shape_prior; a.rate_prior = rate_prior; a.rate_prior_vec = d_rate_prior_vec;
    a.hyper_rate = d_hyper_rate; a.hyper_mean = d_hyper_mean; a.hyper_shape = hyper_shape;
    a.hyper_rate_prior = hyper_rate_prior;
    a.acc = const_cast<float*>(d_acc); a.acc_row_base = acc_row_base;
    a.mc_E = d_mc_E_self;
    const int mode = d_hyper_rate != nullptr ? 1 : 0;
    cudaStream_t s = (cudaStream_t)stream;
    const int nv = a.nvec;
    if (nv <= 4) return launch_combine<4, 1>(a, d_mc_acc, row_begin, row_end, mode, s, st);
    if (nv <= 8) return launch_combine<8, 1>(a, d_mc_acc, row_begin, row_end, mode, s, st);
    if (nv <= 16) return launch_combine<8, 2>(a, d_mc_acc, row_begin, row_end, mode, s, st);
    if (nv <= 24) return launch_combine<8, 3>(a, d_mc_acc, row_begin, row_end, mode, s, st);
    if (nv <= 32) return launch_combine<8, 4>(a, d_mc_acc, row_begin, row_end, mode, s, st);
    return launch_combine<16, 4>(a, d_mc_acc, row_begin, row_end, mode, s, st);
}

int pmf_gamma_combine(int32_t row_begin, int32_t row_end, int32_t K, int32_t ld, const float* d_acc,
                      const float* d_mc_acc, int32_t acc_row_base, float* d_E_self, float* d_mc_E_self, float* d_shp,
                      float* d_rte, float shape_prior, float rate_prior, const float* d_rate_prior_vec,
                      float* d_hyper_rate, float* d_hyper_mean, float hyper_shape, float hyper_rate_prior, void* stream) {
    const StageArgs st = {nullptr, 0, 0, 0, 0};
    return combine_impl(row_begin, row_end, K, ld, d_acc, d_mc_acc, acc_row_base, d_E_self, d_mc_E_self, d_shp, d_rte,
                        shape_prior, rate_prior, d_rate_prior_vec, d_hyper_rate, d_hyper_mean, hyper_shape,
                        hyper_rate_prior, st, stream);
}

int pmf_gamma_combine_staged(int32_t row_begin, int32_t row_end, int32_t K, int32_t ld, const float* d_acc,
                             int32_t acc_row_base, const float* d_stage, int32_t n_src, int64_t src_stride,
                             int32_t self_rank, int32_t stage_row0, float* d_E_self, float* d_mc_E_self, float* d_shp,
                             float* d_rte, float shape_prior, float rate_prior, const float* d_rate_prior_vec,
                             float* d_hyper_rate, float* d_hyper_mean, float hyper_shape, float hyper_rate_prior,
                             void* stream) {
    PMF_REQUIRE(d_stage != nullptr && n_src >= 1 && n_src <= 64 && self_rank >= 0 && self_rank < n_src && stage_row0 >= 0 &&
                    src_stride >= 0, "bad staging arguments");
    const StageArgs st = {d_stage, src_stride, n_src, self_rank, stage_row0};
    return combine_impl(row_begin, row_end, K, ld, d_acc, nullptr, acc_row_base, d_E_self, d_mc_E_self, d_shp, d_rte,
                        shape_prior, rate_prior, d_rate_prior_vec, d_hyper_rate, d_hyper_mean, hyper_shape,
                        hyper_rate_prior, st, stream);
}

int pmf_gamma_pass_ext(const pmf_csr* csr, int32_t K, int32_t ld, const float* d_E_oth, const float* d_scale_oth,
                       float* d_E_self, float* d_shp, float* d_rte, float* d_scale_shp, float* d_scale_rte,
                       float* d_scale_mean, float a0, float b0, void* d_workspace, void* stream) {
    PMF_REQUIRE(csr != nullptr, "csr is NULL");
    PMF_REQUIRE(K >= 1 && ld >= K && ld % 8 == 0 && ld <= 256, "need 1 <= K <= ld <= 256 and ld %% 8 == 0 (K=%d ld=%d)", K, ld);
    PMF_REQUIRE(d_E_oth && d_scale_oth && d_E_self && d_scale_shp && d_scale_rte && d_scale_mean, "NULL table");
    const CsrView c = csr_view(csr);
    PMF_REQUIRE(c.n_partial == 0 || d_workspace != nullptr, "workspace is NULL but %d partial sums are needed", c.n_partial);
    GammaArgs a;
    fill_csr_args(a, c, K, ld);
    a.E_oth = d_E_oth; a.E_self = d_E_self; a.shp = d_shp; a.rte = d_rte;
    a.shape_prior = a0; a.rate_prior = b0; a.rate_prior_vec = nullptr;
    a.hyper_rate = d_scale_rte; a.hyper_mean = d_scale_mean; a.hyper_shape = 0.f; a.hyper_rate_prior = 0.f;
    a.partial = (float*)d_workspace;
    a.partial_x = a.partial ? a.partial + (size_t)c.n_partial * 2 * ld : nullptr;
    a.scale_oth = d_scale_oth; a.scale_shp = d_scale_shp;
    a.mc_E = nullptr;
    return dispatch_gamma(a, 2, (cudaStream_t)stream);
}

}  // extern "C"
