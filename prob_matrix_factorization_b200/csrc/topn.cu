// a11: dense U V^T top-n scoring (evaluation only) -- the ONE GEMM-shaped piece of the path, so the one
// place tensor cores are used.  No reference code exists (nearest: per-dimension nlargest in
// analyze_top_dimensions.py:52-63); semantics fixed by oracle/pmf_oracle.py::topn (PARITY UNPINNED):
//   score[b][j] = float32 chain over k = 0..K-1 of  s = s + u[k]*v[k]  (separate multiply and add),
//   ranking by (score descending, item index ascending), no item excluded.
//
// Pipeline per batch of user rows:
//   tensor path   pack (fp32 -> bf16 UMMA canonical tiles)  ->  topn_mma_kernel: tcgen05.mma 128x128x16,
//                 operands brought to shared memory with cp.async.bulk (1-D bulk TMA) and accumulated in TMEM,
//                 approximate scores S~ to HBM  ->  topn_select_kernel: per row radix-select the approximate
//                 n-th score, keep every item with S~ >= that - 2*margin (margin bounds the bf16 error by
//                 Cauchy-Schwarz, so the true top-n is provably inside), re-score those EXACTLY in fp32 and sort.
//   exact path    topn_score_exact_kernel (CUDA cores, the exact chain) -> same select kernel.  Used when
//                 tensor_cores == 0, and per row as fallback when a row's candidate set overflows.
// Indices are therefore bit-exact against the oracle in both paths.
#include "topn.cuh"

namespace pmf {

// ---------------------------------------------------------------------------------------------------
// packing: fp32 rows -> bf16, UMMA K-major no-swizzle canonical tiles
//   tile t = rows [128 t, 128 t + 128); bytes laid out [k_chunk = kp16/8][row_group = 16][8 rows][8 bf16]
//   => core matrix (8 rows x 16 B) contiguous; SBO (row-group stride) = 128 B; LBO (k-chunk stride) = 2048 B
// ---------------------------------------------------------------------------------------------------
__global__ void topn_pack_kernel(const float* __restrict__ F, const int32_t* __restrict__ rows, int64_t n_rows,
                                 int64_t n_rows_padded, int K, int ld, int kp16, __nv_bfloat16* __restrict__ out) {
    const int kcs = kp16 / 8;
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows_padded * kcs) return;
    const int64_t r = e / kcs;
    const int kc = (int)(e % kcs);
    __align__(16) __nv_bfloat16 v[8];
    const bool live = r < n_rows;
    const float* src = live ? F + (size_t)(rows ? rows[r] : r) * ld : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = kc * 8 + j;
        v[j] = __float2bfloat16_rn((live && k < K) ? src[k] : 0.f);
    }
    const int64_t tile = r / kTile;
    const int rr = (int)(r % kTile);
    const size_t off = (((size_t)tile * kcs + kc) * 16 + (rr >> 3)) * 64 + (size_t)(rr & 7) * 8;
    *reinterpret_cast<uint4*>(out + off) = *reinterpret_cast<const uint4*>(v);
}

__global__ void __launch_bounds__(256) topn_maxnorm_kernel(const float* __restrict__ F, int64_t n_rows, int K, int ld,
                                                          unsigned* __restrict__ out_bits) {
    __shared__ float warp_max[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float best = 0.f;
    // one warp per row, grid-stride; non-negative floats order like their bit patterns, so one atomicMax per CTA
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n_rows; row += (int64_t)gridDim.x * 8) {
        float s = 0.f;
        for (int k = lane; k < K; k += 32) { const float v = F[(size_t)row * ld + k]; s = fmaf(v, v, s); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        best = fmaxf(best, s);
    }
    if (lane == 0) warp_max[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = fmaxf(best, warp_max[w]);
        atomicMax(out_bits, __float_as_uint(best));
    }
}

// ---------------------------------------------------------------------------------------------------
// tcgen05 scoring kernel
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) topn_mma_kernel(const __nv_bfloat16* __restrict__ A_pack,
                                                       const __nv_bfloat16* __restrict__ B_pack, int kp16,
                                                       float* __restrict__ S, int64_t m_padded) {
    extern __shared__ __align__(128) uint8_t tiles[];   // A tile | B tile, each 128 x kp16 bf16
    __shared__ __align__(8) uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t tile_bytes = (uint32_t)kTile * kp16 * 2;
    uint8_t* sA = tiles;
    uint8_t* sB = tiles + tile_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // one warp allocates 128 TMEM columns (the 128x128 fp32 accumulator) and frees them at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_slot;

    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_load, 2 * tile_bytes);
        bulk_g2s(sA, reinterpret_cast<const uint8_t*>(A_pack) + (size_t)blockIdx.y * tile_bytes, tile_bytes, &bar_load);
        bulk_g2s(sB, reinterpret_cast<const uint8_t*>(B_pack) + (size_t)blockIdx.x * tile_bytes, tile_bytes, &bar_load);
        mbar_wait(&bar_load, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 128, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int k = 0; k < kp16 / 16; ++k) {   // one MMA consumes K = 16 = two 8-wide k-chunks (LBO = 2048 B apart)
            const uint64_t adesc = umma_desc(a0 + (uint32_t)k * 4096u, 2048u, 128u);
            const uint64_t bdesc = umma_desc(b0 + (uint32_t)k * 4096u, 2048u, 128u);
            const uint32_t accumulate = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    mbar_wait(&bar_mma, 0);
    __syncwarp();   // the .sync.aligned TMEM loads below need the whole warp converged again
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: warp w owns TMEM lanes 32w..32w+31 = tile rows; each thread drains its row 32 columns at a time
    const int64_t row = (int64_t)blockIdx.y * kTile + warp * 32 + lane;
    float* dst = S + (size_t)row * m_padded + (size_t)blockIdx.x * kTile;
#pragma unroll 1
    for (int c0 = 0; c0 < kTile; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
            "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                   __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// exact scoring on CUDA cores: 128 items x 8 user rows per CTA
// ---------------------------------------------------------------------------------------------------
constexpr int kExactRows = 8;
__global__ void __launch_bounds__(128) topn_score_exact_kernel(const float* __restrict__ F_user, const int32_t* __restrict__ rows,
                                                               int64_t n_rows, const float* __restrict__ F_item, int n_items,
                                                               int K, int ld, float* __restrict__ S, int64_t m_padded) {
    extern __shared__ float su[];   // [kExactRows][K]
    const int64_t r0 = (int64_t)blockIdx.y * kExactRows;
    for (int e = threadIdx.x; e < kExactRows * K; e += blockDim.x) {
        const int64_t r = r0 + e / K;
        su[e] = r < n_rows ? F_user[(size_t)(rows ? rows[r] : r) * ld + e % K] : 0.f;
    }
    __syncthreads();
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= n_items) return;
    const float* v = F_item + (size_t)j * ld;
    float acc[kExactRows];
#pragma unroll
    for (int r = 0; r < kExactRows; ++r) acc[r] = 0.f;
    for (int k = 0; k < K; ++k) {
        const float vk = v[k];
#pragma unroll
        for (int r = 0; r < kExactRows; ++r) acc[r] = __fadd_rn(acc[r], __fmul_rn(su[r * K + k], vk));
    }
#pragma unroll
    for (int r = 0; r < kExactRows; ++r)
        if (r0 + r < n_rows) S[(size_t)(r0 + r) * m_padded + j] = acc[r];
}

// ---------------------------------------------------------------------------------------------------
// selection: one CTA per user row
// ---------------------------------------------------------------------------------------------------
// Top n of one row from its score array s[0..n_items): approximate scores (approx: candidates re-scored exactly) or
// exact ones.  score_first: fill s with exact scores before selecting (the fused path's fallback rows).
__device__ void select_row(const SelArgs& a, int64_t row, float* __restrict__ s, bool approx, bool score_first) {
    __shared__ unsigned hist[256];
    __shared__ unsigned sh[4];
    __shared__ float c_score[kCandCap];
    __shared__ int c_idx[kCandCap];
    __shared__ int s_count, s_eq_taken;
    __shared__ int warp_cnt[kSelThreads / 32];
    extern __shared__ float s_user[];   // [K]
    const int M = a.n_items, n = a.n;
    const float* urow = a.F_user + (size_t)(a.rows ? a.rows[row] : row) * a.ld;
    for (int k = threadIdx.x; k < a.K; k += blockDim.x) s_user[k] = urow[k];
    __syncthreads();
    if (score_first) {
        for (int j = threadIdx.x; j < M; j += blockDim.x) s[j] = exact_dot(s_user, a.F_item + (size_t)j * a.ld, a.K);
        if (threadIdx.x == 0 && a.stats) atomicAdd(a.stats + 0, 1);
        __syncthreads();
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        unsigned key_n;
        int need_eq;
        radix_select(s, M, n, hist, sh, &key_n, &need_eq);
        if (threadIdx.x == 0) { s_count = 0; s_eq_taken = 0; }
        __syncthreads();
        if (approx) {
            // S~ >= (n-th largest S~) - 2*margin contains the exact top n; margin = 2^-7 |u| max|v| bounds the bf16 error
            row_norm_warp0(s_user, a.K, &sh[2]);
            __syncthreads();
            const float margin = 0.0078125f * __uint_as_float(sh[2]) * sqrtf(__uint_as_float(*a.item_maxnorm2_bits));
            const unsigned kn = key_n;
            const float vn = __uint_as_float((kn & 0x80000000u) ? (kn & 0x7FFFFFFFu) : ~kn);
            const float thr = vn - 2.f * margin - 1e-30f;
            for (int j = threadIdx.x; j < M; j += blockDim.x) {
                const float v = s[j];
                if (v >= thr) {
                    const int slot = atomicAdd(&s_count, 1);
                    if (slot < kCandCap) c_idx[slot] = j;
                }
            }
            __syncthreads();
            const int c = s_count;
            if (c > kCandCap) {
                // too many near-ties for the candidate buffer: score this row exactly in place and select exactly
                for (int j = threadIdx.x; j < M; j += blockDim.x) s[j] = exact_dot(s_user, a.F_item + (size_t)j * a.ld, a.K);
                if (threadIdx.x == 0 && a.stats) atomicAdd(a.stats + 0, 1);
                approx = false;
                __syncthreads();
                continue;
            }
            for (int t = threadIdx.x; t < c; t += blockDim.x) c_score[t] = exact_dot(s_user, a.F_item + (size_t)c_idx[t] * a.ld, a.K);
            if (threadIdx.x == 0 && a.stats) atomicAdd(a.stats + 1, c);
            int p2 = 1;
            while (p2 < c) p2 <<= 1;
            for (int t = c + threadIdx.x; t < p2; t += blockDim.x) { c_score[t] = -INFINITY; c_idx[t] = 0x7FFFFFFF; }
            __syncthreads();
            bitonic_sort(c_score, c_idx, p2);
        } else {
            // exact scores: everything above the n-th value, plus the lowest-index ties at the n-th value
            const int count_gt = n - need_eq;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            for (int base = 0; base < M; base += blockDim.x) {   // index order: ties are taken lowest index first
                const int j = base + threadIdx.x;
                const unsigned k = j < M ? order_key(s[j]) : 0u;
                const bool gt = j < M && k > key_n;
                const bool eq = j < M && k == key_n;
                if (gt) {
                    const int slot = atomicAdd(&s_count, 1);
                    c_idx[slot] = j;
                    c_score[slot] = s[j];
                }
                const unsigned bal = __ballot_sync(0xffffffffu, eq);
                if (lane == 0) warp_cnt[warp] = __popc(bal);
                __syncthreads();
                int before = s_eq_taken;
                for (int w = 0; w < warp; ++w) before += warp_cnt[w];
                const int pos = before + __popc(bal & ((1u << lane) - 1u));
                if (eq && pos < need_eq) {
                    c_idx[count_gt + pos] = j;
                    c_score[count_gt + pos] = s[j];
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    int tot = 0;
                    for (int w = 0; w < kSelThreads / 32; ++w) tot += warp_cnt[w];
                    s_eq_taken += tot;
                }
                __syncthreads();
            }
            int p2 = 1;
            while (p2 < n) p2 <<= 1;
            for (int t = n + threadIdx.x; t < p2; t += blockDim.x) { c_score[t] = -INFINITY; c_idx[t] = 0x7FFFFFFF; }
            __syncthreads();
            bitonic_sort(c_score, c_idx, p2);
        }
        break;
    }
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        a.idx_out[(size_t)row * n + t] = c_idx[t];
        a.score_out[(size_t)row * n + t] = c_score[t];
    }
    __syncthreads();   // the shared buffers are reused by the caller's next row
}

__global__ void __launch_bounds__(kSelThreads) topn_select_kernel(const SelArgs a) {
    const int64_t row = blockIdx.x;
    select_row(a, row, a.S + (size_t)row * a.m_padded, a.approx != 0, false);
}

__global__ void __launch_bounds__(kSelThreads) topn_fallback_kernel(const SelArgs a, const int32_t* __restrict__ row_list,
                                                                    const int32_t* __restrict__ row_count,
                                                                    float* __restrict__ scratch) {
    const int count = *row_count;
    for (int i = blockIdx.x; i < count; i += gridDim.x)
        select_row(a, row_list[i], scratch + (size_t)blockIdx.x * a.m_padded, false, true);
}

// host-side launchers used by topn_fused.cu
int topn_launch_pack(const float* F, const int32_t* rows, int64_t n_rows, int64_t n_rows_padded, int K, int ld, int kp16,
                     __nv_bfloat16* out, cudaStream_t s) {
    const int kcs = kp16 / 8;
    topn_pack_kernel<<<(unsigned)cdiv(n_rows_padded * kcs, 256), 256, 0, s>>>(F, rows, n_rows, n_rows_padded, K, ld, kp16, out);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}
int topn_launch_maxnorm(const float* F, int64_t n_rows, int K, int ld, unsigned* out_bits, cudaStream_t s) {
    PMF_CUDA(cudaMemsetAsync(out_bits, 0, sizeof(unsigned), s));
    const int64_t ctas = cdiv(n_rows, 8);
    topn_maxnorm_kernel<<<(unsigned)(ctas < 8 * kNumSMs ? ctas : 8 * kNumSMs), 256, 0, s>>>(F, n_rows, K, ld, out_bits);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}
int topn_launch_fallback(const SelArgs& a, const int32_t* row_list, const int32_t* row_count, float* scratch, int n_ctas,
                         cudaStream_t s) {
    topn_fallback_kernel<<<(unsigned)n_ctas, kSelThreads, (size_t)a.K * sizeof(float), s>>>(a, row_list, row_count, scratch);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

}  // namespace pmf

using namespace pmf;

extern "C" {

static int64_t pad_to(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

static int64_t unfused_workspace_bytes(int64_t batch_rows, int32_t n_items, int32_t K) {
    const int64_t bp = pad_to(batch_rows, kTile), mp = pad_to(n_items, kTile), kp16 = pad_to(K, 16);
    return bp * mp * 4 + (bp + mp) * kp16 * 2 + 256;
}

// tensor_cores: 0 exact CUDA-core scoring, 1 tcgen05 (fused filter when the shape allows it), 2 tcgen05 unfused
// (score matrix through HBM; kept as the comparison point)
static bool use_fused(int32_t K, int32_t n, int32_t tensor_cores) { return tensor_cores == 1 && topn_fused_supported(K, n); }

int64_t pmf_topn_workspace_bytes_ex(int64_t batch_rows, int32_t n_items, int32_t K, int32_t n, int32_t tensor_cores) {
    if (batch_rows <= 0 || n_items <= 0 || K <= 0 || n <= 0) return -1;
    if (use_fused(K, n, tensor_cores)) return topn_fused_workspace_bytes(batch_rows, n_items, K);
    return unfused_workspace_bytes(batch_rows, n_items, K);
}

int64_t pmf_topn_workspace_bytes(int64_t batch_rows, int32_t n_items, int32_t K) {   // enough for every mode and n
    if (batch_rows <= 0 || n_items <= 0 || K <= 0) return -1;
    const int64_t a = unfused_workspace_bytes(batch_rows, n_items, K), b = topn_fused_workspace_bytes(batch_rows, n_items, K);
    return a > b ? a : b;
}

int pmf_topn(const float* d_F_user, const int32_t* d_user_rows, int64_t batch_rows, const float* d_F_item,
             int32_t n_items, int32_t K, int32_t ld, int32_t n, int32_t tensor_cores, int32_t* d_idx, float* d_score,
             void* d_workspace, int64_t workspace_bytes, int32_t* d_stats, void* stream) {
    PMF_REQUIRE(batch_rows >= 0 && n_items > 0 && K >= 1 && ld >= K, "bad shape");
    PMF_REQUIRE(n >= 1 && n <= n_items && n <= kCandCap / 2, "n=%d must lie in [1, min(n_items, %d)]", n, kCandCap / 2);
    PMF_REQUIRE(tensor_cores >= 0 && tensor_cores <= 2, "tensor_cores must be 0, 1 or 2");
    if (batch_rows == 0) return PMF_OK;
    PMF_REQUIRE(d_F_user && d_F_item && d_idx && d_score && d_workspace, "NULL argument");
    PMF_REQUIRE(workspace_bytes >= pmf_topn_workspace_bytes_ex(batch_rows, n_items, K, n, tensor_cores),
                "workspace too small: %lld < %lld bytes", (long long)workspace_bytes,
                (long long)pmf_topn_workspace_bytes_ex(batch_rows, n_items, K, n, tensor_cores));
    PMF_REQUIRE(K <= 512, "K=%d too wide for top-n scoring", K);
    cudaStream_t s = (cudaStream_t)stream;
    if (d_stats) PMF_CUDA(cudaMemsetAsync(d_stats, 0, 2 * sizeof(int32_t), s));
    if (use_fused(K, n, tensor_cores)) {
        TopnProblem p{d_F_user, d_user_rows, batch_rows, d_F_item, n_items, K, ld, n, d_idx, d_score, d_stats};
        return topn_fused_run(p, d_workspace, workspace_bytes, s);
    }
    const int64_t bp = pad_to(batch_rows, kTile), mp = pad_to(n_items, kTile);
    const int kp16 = (int)pad_to(K, 16);
    uint8_t* ws = (uint8_t*)d_workspace;
    float* S = (float*)ws;
    __nv_bfloat16* A_pack = (__nv_bfloat16*)(ws + bp * mp * 4);
    __nv_bfloat16* B_pack = A_pack + bp * kp16;
    unsigned* maxnorm = (unsigned*)(B_pack + mp * kp16);
    if (tensor_cores) {
        PMF_TRY(topn_launch_maxnorm(d_F_item, n_items, K, ld, maxnorm, s));
        PMF_TRY(topn_launch_pack(d_F_user, d_user_rows, batch_rows, bp, K, ld, kp16, A_pack, s));
        PMF_TRY(topn_launch_pack(d_F_item, nullptr, n_items, mp, K, ld, kp16, B_pack, s));
        const size_t smem = (size_t)2 * kTile * kp16 * 2;
        PMF_REQUIRE(smem <= 200 * 1024, "K=%d needs %zu bytes of shared memory per tile pair", K, smem);
        PMF_CUDA(cudaFuncSetAttribute(topn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)(mp / kTile), (unsigned)(bp / kTile));
        topn_mma_kernel<<<grid, 128, smem, s>>>(A_pack, B_pack, kp16, S, mp);
        PMF_LAUNCH_CHECK();
    } else {
        dim3 grid((unsigned)cdiv(n_items, 128), (unsigned)cdiv(batch_rows, kExactRows));
        topn_score_exact_kernel<<<grid, 128, (size_t)kExactRows * K * sizeof(float), s>>>(d_F_user, d_user_rows, batch_rows, d_F_item,
                                                                                           n_items, K, ld, S, mp);
        PMF_LAUNCH_CHECK();
    }
    SelArgs a;
    a.S = S; a.m_padded = mp; a.n_items = n_items; a.n = n; a.approx = tensor_cores ? 1 : 0;
    a.F_user = d_F_user; a.F_item = d_F_item; a.rows = d_user_rows; a.K = K; a.ld = ld;
    a.item_maxnorm2_bits = maxnorm; a.idx_out = d_idx; a.score_out = d_score; a.stats = d_stats;
    topn_select_kernel<<<(unsigned)batch_rows, kSelThreads, (size_t)K * sizeof(float), s>>>(a);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

}  // extern "C"
