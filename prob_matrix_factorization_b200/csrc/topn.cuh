// Shared pieces of the top-n scoring kernels.
//   topn.cu        packing, the unfused tcgen05 / exact scoring kernels, the per-row selection, pmf_topn dispatch
//   topn_fused.cu  the persistent tcgen05 filter kernel: scores never leave the SM (TMEM -> registers -> threshold test)
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace pmf {

constexpr int kTile = 128;          // UMMA M = N = 128
constexpr int kCandCap = 2048;      // candidates kept per row in shared memory by the selection kernels
constexpr int kSelThreads = 256;

// The oracle's float32 chain  s = s + u[k]*v[k], k = 0..K-1 (no FMA contraction).  The chain is sequential, so the item row
// is fetched 32 floats at a time (eight 128-bit loads in flight) instead of one dependent load per step.
__device__ __forceinline__ float exact_dot(const float* __restrict__ u, const float* __restrict__ v, int K) {
    float s = 0.f;
    int k = 0;
    if ((reinterpret_cast<uintptr_t>(v) & 15u) == 0) {
        for (; k + 32 <= K; k += 32) {
            float4 x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = __ldg(reinterpret_cast<const float4*>(v + k) + i);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s = __fadd_rn(s, __fmul_rn(u[k + 4 * i], x[i].x));
                s = __fadd_rn(s, __fmul_rn(u[k + 4 * i + 1], x[i].y));
                s = __fadd_rn(s, __fmul_rn(u[k + 4 * i + 2], x[i].z));
                s = __fadd_rn(s, __fmul_rn(u[k + 4 * i + 3], x[i].w));
            }
        }
        for (; k + 4 <= K; k += 4) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(v + k));
            s = __fadd_rn(s, __fmul_rn(u[k], x.x));
            s = __fadd_rn(s, __fmul_rn(u[k + 1], x.y));
            s = __fadd_rn(s, __fmul_rn(u[k + 2], x.z));
            s = __fadd_rn(s, __fmul_rn(u[k + 3], x.w));
        }
    }
    for (; k < K; ++k) s = __fadd_rn(s, __fmul_rn(u[k], v[k]));
    return s;
}

// |u|^2 of the row staged in shared memory, by warp 0; the result is published through sh_out after the caller's
// next __syncthreads().
__device__ __forceinline__ void row_norm_warp0(const float* s_user, int K, unsigned* sh_out) {
    if (threadIdx.x < 32) {
        float un2 = 0.f;
        for (int k = threadIdx.x; k < K; k += 32) un2 = fmaf(s_user[k], s_user[k], un2);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) un2 += __shfl_xor_sync(0xffffffffu, un2, o);
        if (threadIdx.x == 0) *sh_out = __float_as_uint(sqrtf(un2));
    }
}

// ---------------------------------------------------------------------------------------------------
// mbarrier / bulk copy / tcgen05 wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();   // a lost arrival must surface as an error, never as a hung GPU
    } while (!ok);
}
// Same, for the long waits of a persistent pipeline: the hardware may keep the warp suspended for up to ~1 us per try
// instead of returning at once, so waiting warps do not eat the issue slots of the warps they are waiting for.
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(1000u) : "memory");
        if (!ok && ++spins > (1u << 21)) __trap();
    } while (!ok);
}
__device__ __forceinline__ bool elect_one() {   // true in exactly one lane of a converged warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, N = 128, M = 128
__device__ __forceinline__ uint32_t umma_idesc_bf16_128x128() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on `bar` once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 consecutive accumulator columns of this thread's TMEM lane (warp-collective, returns after the data has landed)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// selection helpers (one CTA of kSelThreads threads works on one user row)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned order_key(float v) {   // larger float <=> larger unsigned
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct SelArgs {
    float* S;
    int64_t m_padded;
    int32_t n_items, n, approx;
    const float *F_user, *F_item;
    const int32_t* rows;
    int32_t K, ld;
    const unsigned* item_maxnorm2_bits;
    int32_t* idx_out;
    float* score_out;
    int32_t* stats;   // [0] rows re-scored exactly in full, [1] candidates re-scored (approx path)
};

// key of the n-th largest element of s[0..M) and how many of the elements equal to it belong to the top n.
// hist: 256 counters, sh: 2 words, both shared; blockDim.x >= 32.
__device__ inline void radix_select(const float* __restrict__ s, int M, int n, unsigned* hist, unsigned* sh, unsigned* key_out,
                                    int* need_eq_out) {
    unsigned prefix = 0, mask = 0;
    int remaining = n;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        for (int j0 = 0; j0 < M; j0 += blockDim.x) {   // uniform trip count: the warp votes below
            const int j = j0 + threadIdx.x;
            unsigned digit = 256u;                      // 256 = not a candidate for this pass
            if (j < M) {
                const unsigned k = order_key(s[j]);
                if ((k & mask) == prefix) digit = (k >> shift) & 255u;
            }
            // scores of one row share their leading bytes, so most lanes want the same counter: one atomic per distinct
            // digit per warp instead of a 32-way same-address conflict
            const unsigned peers = __match_any_sync(0xffffffffu, digit);
            if (digit != 256u && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // digit d with  count(digits > d) < remaining <= count(digits >= d):  lane l owns digits 255-8l .. 248-8l
            // (descending), a warp scan gives the count above each lane's block, the owning lane walks its 8 digits
            const int lane = threadIdx.x, top = 255 - 8 * lane;
            int mine = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) mine += (int)hist[top - i];
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int above = incl - mine;
            if (above < remaining && remaining <= incl) {   // exactly one lane (counts are cumulative and total >= remaining)
                int d = top;
                for (;; --d) {
                    const int h = (int)hist[d];
                    if (above + h >= remaining) break;
                    above += h;
                }
                sh[0] = (unsigned)d;
                sh[1] = (unsigned)(remaining - above);
            }
        }
        __syncthreads();
        prefix |= sh[0] << shift;
        mask |= 255u << shift;
        remaining = (int)sh[1];
        __syncthreads();
    }
    *key_out = prefix;
    *need_eq_out = remaining;
}

__device__ __forceinline__ bool ranks_before(float sa, int ia, float sb, int ib) { return sa > sb || (sa == sb && ia < ib); }

__device__ inline void bitonic_sort(float* sc, int* ix, int n_pow2) {
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n_pow2; t += blockDim.x) {
                const int p = t ^ j;
                if (p > t) {
                    const bool up = (t & k) == 0;
                    const bool swap = up ? ranks_before(sc[p], ix[p], sc[t], ix[t]) : ranks_before(sc[t], ix[t], sc[p], ix[p]);
                    if (swap) {
                        const float ts = sc[t]; sc[t] = sc[p]; sc[p] = ts;
                        const int ti = ix[t]; ix[t] = ix[p]; ix[p] = ti;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// entry points shared between the two translation units (host side)
// ---------------------------------------------------------------------------------------------------
struct TopnProblem {
    const float* F_user;
    const int32_t* user_rows;
    int64_t batch_rows;
    const float* F_item;
    int32_t n_items, K, ld, n;
    int32_t* idx_out;
    float* score_out;
    int32_t* stats;
};
extern thread_local int g_tune_topn_growth;   // 0 = auto (256 / n); else items covered per level = (1 + growth) x what was covered (pmf_tune)
bool topn_fused_supported(int32_t K, int32_t n);
int64_t topn_fused_workspace_bytes(int64_t batch_rows, int32_t n_items, int32_t K);
int topn_fused_run(const TopnProblem& p, void* workspace, int64_t workspace_bytes, cudaStream_t s);

// host-side launchers of kernels defined in topn.cu (no relocatable device code: each kernel is launched from its own file)
int topn_launch_pack(const float* F, const int32_t* rows, int64_t n_rows, int64_t n_rows_padded, int K, int ld, int kp16,
                     __nv_bfloat16* out, cudaStream_t s);
int topn_launch_maxnorm(const float* F, int64_t n_rows, int K, int ld, unsigned* out_bits, cudaStream_t s);
// exact scoring of the whole rows listed in row_list[0 .. *row_count) into per-CTA scratch (n_ctas x m_padded floats),
// then exact selection; the list and its length live on the device (no host synchronisation)
int topn_launch_fallback(const SelArgs& a, const int32_t* row_list, const int32_t* row_count, float* scratch, int n_ctas,
                         cudaStream_t s);

}  // namespace pmf
