// a11: dense U V^T top-n scoring, fused form -- the score matrix never reaches HBM.
//
// The unfused pipeline in topn.cu writes batch x n_items approximate scores (1.8 GB for 2048 x 230k) and reads them
// back five times to select; its tensor pipe idles waiting for those stores.  Here the MMA epilogue tests every score
// against a per-row threshold while it is still in registers and only the survivors (a few hundred per row) are kept:
//
//   level 0   the first kLevel0Tiles item tiles are scored with no threshold ("dense": every score kept, <= kFuseCap/2)
//   refine    per row: t = n-th largest approximate score so far.  The n-th largest over ALL items can only be larger,
//             so an item scoring below t - 2*margin can never be needed (margin bounds the bf16 error, see topn.cu);
//             the list is compacted to the entries above that threshold
//   level l   the next item range (g times what has been covered, g ~ 256/n) is scored against the thresholds;
//             expected survivors per row ~ n*g, far below the list capacity
//   finish    after the last level the list holds every item with S~ >= t - 2*margin >= (true n-th S~) - 2*margin, i.e.
//             exactly the candidate set the unfused path proves sufficient: exact fp32 re-score, sort, write the top n
//   fallback  rows whose list overflowed (massive ties) are scored exactly in full by topn_fallback_kernel; the list of
//             such rows lives on the device, so there is no host synchronisation anywhere
//
// topn_filter_kernel is a persistent warp-specialised tcgen05 kernel, one CTA per SM:
//   warp 0        producer: cp.async.bulk of the two 128-row user tiles of a work unit, then a ring of item tiles
//   warp 1        one thread issues tcgen05.mma 128x128x16 (bf16 -> fp32 in TMEM): each item tile is multiplied with
//                 BOTH user tiles, so every byte of B fetched from L2 feeds 256 rows (L2 read rate 32 B/clk/SM, under
//                 the ~42 B/clk/SM the L2 sustains chip-wide); two accumulator stages x two row blocks = 512 TMEM columns
//   warps 2..17   epilogue, two sets of 8 warps; set s drains accumulator stage s (every other item tile): a warp reads
//                 TMEM lanes 32*(w%4).. of one row block with tcgen05.ld.x32 (thread = user row), reduces its 32 scores to
//                 maxima of 4 and of 32 and votes; only groups of 4 in which some lane beats its row's threshold are
//                 walked, with predicated stores into a shared-memory staging area that the warp flushes to the rows'
//                 global lists with one list-space reservation per row
#include "topn.cuh"

namespace pmf {

constexpr int kFuseCap = 4096;                 // list capacity per row
constexpr int kLevel0Tiles = kFuseCap / 2 / kTile;   // 16 tiles = 2048 items scored densely
constexpr int kFuseRows = 2 * kTile;           // user rows per CTA (two UMMA M = 128 row blocks)
constexpr int kEpiWarps = 16;                  // warps 2..17: two sets of 8, set s drains accumulator stage s
constexpr int kFuseThreads = 64 + 32 * kEpiWarps;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kStage = 12;                     // survivors staged per epilogue thread before list space is reserved
constexpr size_t kExtraBytes = (size_t)kEpiThreads * kStage * 8;
constexpr int kFuseMaxN = 256;                 // level 0 must hold several times n items
constexpr int kMaxStages = 4;
constexpr int kFallbackCtas = 148;
constexpr size_t kFuseSmemBudget = 220 * 1024;

struct FilterArgs {
    const __nv_bfloat16 *A_pack, *B_pack;
    int32_t kp16, stages;
    int32_t n_rowpairs;                  // 256-row blocks of the batch
    int32_t tile_begin, tile_end;        // item tiles of this level
    int32_t steps_per_cta;               // the (row block, item tile) steps of the level are dealt out in contiguous runs
    int32_t n_items, dense;
    const float* thr;                    // [rows padded to 256]; +inf for padding rows and rows already overflowed
    int32_t* cand_idx;
    float* cand_score;                   // [rows][kFuseCap]
    int32_t* cand_cnt;
};

// Survivors are staged in shared memory, kStage per epilogue thread ([slot][thread]: the lanes of a warp hit distinct
// banks), and moved to the rows' global lists by the whole warp at once: every lane reserves space in its row's list with
// ONE returning atomic (all lanes' atomics in flight together) and copies its own entries.  s_*: this thread's staging
// column; g_*: this thread's row.
__device__ __noinline__ void stage_flush_all(const int* s_ix, const float* s_sc, int count, int32_t* g_cnt, int32_t* g_ix, float* g_sc) {
    if (count > 0) {
        const int base = atomicAdd(g_cnt, count);   // other CTAs (other item ranges) append to the same row
        for (int i = 0; i < count; ++i) {
            const int slot = base + i;
            if (slot < kFuseCap) { g_ix[slot] = s_ix[i * kEpiThreads]; g_sc[slot] = s_sc[i * kEpiThreads]; }
        }
    }
}

// Four scores of one thread against its row's threshold: straight-line, predicated, no register indexing.
template <bool CHECK>
__device__ __forceinline__ void push_group(const uint32_t (&r)[32], int i, float thr, int item0, int limit, int* s_ix, float* s_sc,
                                           int& count) {
#pragma unroll
    for (int j = 4 * i; j < 4 * i + 4; ++j) {
        const float v = __uint_as_float(r[j]);
        bool p = v >= thr;
        if (CHECK) p = p && j < limit;       // the last item tile is padded with zero rows
        if (p) {
            s_ix[count * kEpiThreads] = item0 + j;
            s_sc[count * kEpiThreads] = v;
            ++count;
        }
    }
}

// Work of a level = n_rowpairs x tiles steps, linearised row block major; CTA b owns steps [b*L, (b+1)*L).  A work unit is
// the part of that run inside one row block (the A tiles change between units).  Every warp role walks the same units.
struct UnitWalk {
    int64_t lin, end;
    int tiles, tile_begin;
    __device__ UnitWalk(const FilterArgs& a) {
        tiles = a.tile_end - a.tile_begin;
        tile_begin = a.tile_begin;
        const int64_t total = (int64_t)a.n_rowpairs * tiles;
        lin = (int64_t)blockIdx.x * a.steps_per_cta;
        end = lin + a.steps_per_cta < total ? lin + a.steps_per_cta : total;
    }
    __device__ bool next(int& rp, int& t0, int& t1) {
        if (lin >= end) return false;
        rp = (int)(lin / tiles);
        const int off = (int)(lin - (int64_t)rp * tiles);
        const int64_t run = end - lin < tiles - off ? end - lin : tiles - off;
        t0 = tile_begin + off;
        t1 = t0 + (int)run;
        lin += run;
        return true;
    }
};

__global__ void __launch_bounds__(kFuseThreads, 1) topn_filter_kernel(const FilterArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];   // A0 | A1 | B[stages] | staging indices | staging scores
    __shared__ __align__(8) uint64_t bar_a_full, bar_a_empty, bar_b_full[kMaxStages], bar_b_empty[kMaxStages], bar_acc_full[4],
        bar_acc_empty[4];   // accumulator slot = stage * 2 + row block
    __shared__ uint32_t tmem_base_slot;
    const uint32_t tile_bytes = (uint32_t)kTile * a.kp16 * 2;
    uint8_t* sA = smem;
    uint8_t* sB = smem + 2 * tile_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = a.stages;
    if (threadIdx.x == 0) {
        mbar_init(&bar_a_full, 1);
        mbar_init(&bar_a_empty, 1);
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&bar_b_full[i], 1); mbar_init(&bar_b_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // all 512 TMEM columns: [stage][row block][128]
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t bs = 0, bphase = 0, ucount = 0;
            UnitWalk walk(a);
            for (int rp, t0, t1; walk.next(rp, t0, t1); ++ucount) {
                mbar_wait_sleepy(&bar_a_empty, (ucount & 1u) ^ 1u);   // the previous unit's MMAs no longer read A
                mbar_expect_tx(&bar_a_full, 2 * tile_bytes);
                bulk_g2s(sA, reinterpret_cast<const uint8_t*>(a.A_pack) + (size_t)rp * 2 * tile_bytes, 2 * tile_bytes, &bar_a_full);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait_sleepy(&bar_b_empty[bs], bphase ^ 1u);
                    mbar_expect_tx(&bar_b_full[bs], tile_bytes);
                    bulk_g2s(sB + (size_t)bs * tile_bytes, reinterpret_cast<const uint8_t*>(a.B_pack) + (size_t)t * tile_bytes,
                             tile_bytes, &bar_b_full[bs]);
                    if (++bs == (uint32_t)stages) { bs = 0; bphase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // The whole warp walks the loops (uniform control flow keeps the descriptor arithmetic on the uniform datapath);
        // one elected lane issues.  Descriptor low word = address/16 | LBO/16 << 16, high word = SBO/16 | version.
        const uint32_t idesc = umma_idesc_bf16_128x128();
        const uint32_t tile16 = tile_bytes >> 4;
        const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (128u << 16);
        const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (128u << 16);
        const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;
        const int ksteps = a.kp16 / 16;
        uint32_t bs = 0, bphase = 0, as = 0, aphase = 0, ucount = 0;
        UnitWalk walk(a);
        for (int rp, t0, t1; walk.next(rp, t0, t1); ++ucount) {
            mbar_wait_sleepy(&bar_a_full, ucount & 1u);
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&bar_b_full[bs], bphase);
                const uint32_t b_lo = b_lo0 + bs * tile16;
#pragma unroll
                for (int rb = 0; rb < 2; ++rb) {
                    const uint32_t slot = as * 2u + (uint32_t)rb;
                    mbar_wait(&bar_acc_empty[slot], aphase ^ 1u);   // the epilogue has drained this accumulator
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d = tmem + slot * 128u;
                        const uint32_t a_lo = a_lo0 + (uint32_t)rb * tile16;
                        for (int k = 0; k < ksteps; ++k)   // one MMA consumes K = 16 = two 8-wide k-chunks = 4096 B of each tile
                            umma_bf16(d, desc_hi | (a_lo + (uint32_t)k * 256u), desc_hi | (b_lo + (uint32_t)k * 256u), idesc, k > 0);
                        umma_commit(&bar_acc_full[slot]);                  // this row block's scores are complete -> epilogue
                        if (rb == 1) umma_commit(&bar_b_empty[bs]);        // item tile consumed -> producer may refill the slot
                    }
                    __syncwarp();
                }
                if (++bs == (uint32_t)stages) { bs = 0; bphase ^= 1u; }
                if (++as == 2u) { as = 0; aphase ^= 1u; }
            }
            if (elect_one()) umma_commit(&bar_a_empty);                    // every MMA of the unit done -> A tiles may be replaced
            __syncwarp();
        }
    } else {
        const int ew = warp - 2;                        // 0..15
        const uint32_t set = (uint32_t)ew >> 3;         // the accumulator stage this warp drains (every other item tile)
        const int rb = (ew & 7) >> 2, q = warp & 3;     // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        int* s_ix = reinterpret_cast<int*>(sB + (size_t)stages * tile_bytes) + (threadIdx.x - 64);   // [kStage][kEpiThreads]
        float* s_sc = reinterpret_cast<float*>(sB + (size_t)stages * tile_bytes) + kEpiThreads * kStage + (threadIdx.x - 64);
        int count = 0;                                  // entries staged by this thread
        uint32_t aphase = 0, gt = 0;                    // gt: item tiles this CTA has gone through (the MMA warp's sequence)
        UnitWalk walk(a);
        for (int rp, t0, t1; walk.next(rp, t0, t1);) {
            const int64_t row0 = (int64_t)rp * kFuseRows + rb * kTile + q * 32;
            const float thr = a.thr[row0 + lane];
            int32_t* g_cnt = a.cand_cnt + row0 + lane;
            int32_t* g_ix = a.cand_idx + (size_t)(row0 + lane) * kFuseCap;
            float* g_sc = a.cand_score + (size_t)(row0 + lane) * kFuseCap;
            for (int t = t0; t < t1; ++t, ++gt) {
                if ((gt & 1u) != set) continue;
                mbar_wait(&bar_acc_full[set * 2u + (uint32_t)rb], aphase);
                __syncwarp();
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + set * 256u + (uint32_t)rb * 128u;
#pragma unroll 1
                for (int c0 = 0; c0 < kTile; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + (uint32_t)c0, r);
                    const int item0 = t * kTile + c0;
                    if (a.dense) {   // level 0: every score is kept, slot = position in the level
                        const int slot0 = item0 - a.tile_begin * kTile;
                        if (thr < INFINITY) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)   // the item index is the slot: the first refine fills it in
                                *reinterpret_cast<uint4*>(g_sc + slot0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                        }
                        continue;
                    }
                    // maxima of groups of 4, then of all 32: ~25 instructions when no lane of the warp has a hit.  Otherwise
                    // only the groups in which some lane has a hit are walked, with predicated stores into the staging area.
                    float g[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        g[i] = fmaxf(fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])),
                                     fmaxf(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
                    const float mx = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
                    if (__any_sync(0xffffffffu, mx >= thr)) {
                        const int limit = a.n_items - item0;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (__any_sync(0xffffffffu, g[i] >= thr)) {
                                if (__any_sync(0xffffffffu, count > kStage - 4)) {   // a group adds at most 4 entries
                                    stage_flush_all(s_ix, s_sc, count, g_cnt, g_ix, g_sc);
                                    count = 0;
                                }
                                if (limit >= 32) push_group<false>(r, i, thr, item0, limit, s_ix, s_sc, count);
                                else push_group<true>(r, i, thr, item0, limit, s_ix, s_sc, count);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_acc_empty[set * 2u + (uint32_t)rb]);
                aphase ^= 1u;
            }
            if (__any_sync(0xffffffffu, count > 0)) {   // the next unit works on other rows
                stage_flush_all(s_ix, s_sc, count, g_cnt, g_ix, g_sc);
                count = 0;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// per-row list maintenance
// ---------------------------------------------------------------------------------------------------
__global__ void topn_fused_init_kernel(float* thr, int32_t* cnt, int32_t* ovf_count, int64_t rows, int64_t rows_padded, int32_t cnt0) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) *ovf_count = 0;
    if (r >= rows_padded) return;
    thr[r] = r < rows ? -INFINITY : INFINITY;
    cnt[r] = r < rows ? cnt0 : 0;
}

struct RefineArgs {
    int32_t* cand_idx;
    float* cand_score;
    int32_t* cand_cnt;
    float* thr;
    int32_t n, final_level;
    int32_t dense_input;      // the list is level 0's output: scores only, item index = slot
    const float *F_user, *F_item;
    const int32_t* rows;
    int32_t K, ld;
    const unsigned* item_maxnorm2_bits;
    int32_t* idx_out;
    float* score_out;
    int32_t* stats;
    int32_t *ovf_list, *ovf_count;
};
constexpr int kRefineThreads = 128;

// A lower bound of the n-th largest of s[0..c) (c >= n), tight to 2^-16 of the value range.  Thresholds only have to be
// lower bounds, so instead of four exact radix passes over float keys whose leading bytes all agree, the keys are
// rebased to the row's minimum and shifted so the top 16 bits carry the spread, then two 8-bit passes pick the bucket
// of the n-th largest; its lower edge is returned.  hist: 256 counters, sh: 4 words (shared).  All threads must call.
__device__ __forceinline__ float nth_largest_lower_bound(const float* s, int c, int n, unsigned* hist, unsigned* sh) {
    unsigned kmin = 0xFFFFFFFFu, kmax = 0u;
    for (int j = threadIdx.x; j < c; j += blockDim.x) {
        const unsigned k = order_key(s[j]);
        kmin = min(kmin, k);
        kmax = max(kmax, k);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (threadIdx.x == 0) { sh[2] = 0xFFFFFFFFu; sh[3] = 0u; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { atomicMin(&sh[2], kmin); atomicMax(&sh[3], kmax); }
    __syncthreads();
    kmin = sh[2];
    kmax = sh[3];
    if (kmax == kmin) return key_to_float(kmin);
    const int lz = __clz(kmax - kmin);
    unsigned prefix = 0, mask = 0;
    int remaining = n;
    for (int shift = 24; shift >= 16; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        for (int j0 = 0; j0 < c; j0 += blockDim.x) {   // uniform trip count: the warp votes below
            const int j = j0 + threadIdx.x;
            unsigned digit = 256u;
            if (j < c) {
                const unsigned d = (order_key(s[j]) - kmin) << lz;
                if ((d & mask) == prefix) digit = (d >> shift) & 255u;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, digit);
            if (digit != 256u && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (threadIdx.x < 32) {   // as in radix_select: lane l owns digits 255-8l .. 248-8l, descending
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
            if (above < remaining && remaining <= incl) {
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
    return key_to_float(kmin + (prefix >> lz));   // every key of the chosen bucket is >= its lower edge
}

// One CTA per user row.  Dynamic shared memory: user row | list scores | list indices [| kept scores | kept indices].
__global__ void __launch_bounds__(kRefineThreads) topn_refine_kernel(const RefineArgs a) {
    extern __shared__ __align__(16) uint8_t dyn[];
    float* s_user = reinterpret_cast<float*>(dyn);                    // [K rounded up to 4]
    float* l_sc = s_user + ((a.K + 3) & ~3);
    int* l_ix = reinterpret_cast<int*>(l_sc + kFuseCap);
    float* k_sc = reinterpret_cast<float*>(l_ix + kFuseCap);         // kept scores / indices: final level only
    int* k_ix = reinterpret_cast<int*>(k_sc + kCandCap);
    __shared__ unsigned hist[256];
    __shared__ unsigned sh[4];
    __shared__ int s_count;
    const int64_t row = blockIdx.x;
    const int n = a.n;
    if (a.thr[row] == INFINITY) return;     // overflowed at an earlier level: already on the fallback list
    const int c = a.cand_cnt[row];
    int32_t* g_ix = a.cand_idx + (size_t)row * kFuseCap;
    float* g_sc = a.cand_score + (size_t)row * kFuseCap;
    // a list that overflowed, or (final level) one that cannot hold n entries, sends the row to exact scoring
    bool bad = c > kFuseCap || (a.final_level && c < n);
    if (!bad) {
        const float* urow = a.F_user + (size_t)(a.rows ? a.rows[row] : row) * a.ld;
        for (int k = threadIdx.x; k < a.K; k += blockDim.x) s_user[k] = urow[k];
        for (int t = threadIdx.x; t < c; t += blockDim.x) { l_sc[t] = g_sc[t]; l_ix[t] = a.dense_input ? t : g_ix[t]; }
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        float thr_new = -INFINITY;
        if (c >= n) {
            const float nth = nth_largest_lower_bound(l_sc, c, n, hist, sh);
            __syncthreads();
            row_norm_warp0(s_user, a.K, &sh[2]);
            __syncthreads();
            // margin = 2^-7 |u| max|v| bounds |S~ - S| (bf16 operands, Cauchy-Schwarz); see topn.cu::select_row
            const float margin = 0.0078125f * __uint_as_float(sh[2]) * sqrtf(__uint_as_float(*a.item_maxnorm2_bits));
            thr_new = nth - 2.f * margin - 1e-30f;
        }
        if (!a.final_level) {
            for (int t = threadIdx.x; t < c; t += blockDim.x) {
                if (l_sc[t] >= thr_new) {
                    const int slot = atomicAdd(&s_count, 1);
                    g_sc[slot] = l_sc[t];
                    g_ix[slot] = l_ix[t];
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) { a.cand_cnt[row] = s_count; a.thr[row] = thr_new; }
            return;
        }
        for (int t = threadIdx.x; t < c; t += blockDim.x) {
            if (l_sc[t] >= thr_new) {
                const int slot = atomicAdd(&s_count, 1);
                if (slot < kCandCap) k_ix[slot] = l_ix[t];
            }
        }
        __syncthreads();
        const int kept = s_count;
        bad = kept > kCandCap || kept < n;
        if (!bad) {
            for (int t = threadIdx.x; t < kept; t += blockDim.x) k_sc[t] = exact_dot(s_user, a.F_item + (size_t)k_ix[t] * a.ld, a.K);
            if (threadIdx.x == 0 && a.stats) atomicAdd(a.stats + 1, kept);
            int p2 = 1;
            while (p2 < kept) p2 <<= 1;
            for (int t = kept + threadIdx.x; t < p2; t += blockDim.x) { k_sc[t] = -INFINITY; k_ix[t] = 0x7FFFFFFF; }
            __syncthreads();
            bitonic_sort(k_sc, k_ix, p2);
            for (int t = threadIdx.x; t < n; t += blockDim.x) {
                a.idx_out[(size_t)row * n + t] = k_ix[t];
                a.score_out[(size_t)row * n + t] = k_sc[t];
            }
            return;
        }
    }
    if (threadIdx.x == 0) {
        a.ovf_list[atomicAdd(a.ovf_count, 1)] = (int32_t)row;
        a.thr[row] = INFINITY;   // later levels append nothing for this row
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
thread_local int g_tune_topn_growth = 0;

static int64_t pad_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

static int fused_stages(int kp16) {
    const size_t tile = (size_t)kTile * kp16 * 2;
    int s = (int)((kFuseSmemBudget - kExtraBytes) / tile) - 2;
    return s > kMaxStages ? kMaxStages : s;
}

bool topn_fused_supported(int32_t K, int32_t n) { return n <= kFuseMaxN && fused_stages((int)pad_up(K, 16)) >= 2; }

struct FusedLayout {
    int64_t bp, mp, kp16;
    size_t off_A, off_B, off_idx, off_score, off_cnt, off_thr, off_ovf, off_misc, off_scratch, total;
    int fallback_ctas;
};

static FusedLayout fused_layout(int64_t batch_rows, int32_t n_items, int32_t K) {
    FusedLayout L;
    L.bp = pad_up(batch_rows, kFuseRows);
    L.mp = pad_up(n_items, kTile);
    L.kp16 = pad_up(K, 16);
    L.fallback_ctas = (int)(batch_rows < kFallbackCtas ? batch_rows : kFallbackCtas);
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) / 256 * 256; return at; };
    L.off_A = take((size_t)L.bp * L.kp16 * 2);
    L.off_B = take((size_t)L.mp * L.kp16 * 2);
    L.off_idx = take((size_t)L.bp * kFuseCap * 4);
    L.off_score = take((size_t)L.bp * kFuseCap * 4);
    L.off_cnt = take((size_t)L.bp * 4);
    L.off_thr = take((size_t)L.bp * 4);
    L.off_ovf = take((size_t)L.bp * 4);
    L.off_misc = take(256);                                     // [0] item max |v|^2 bits, [1] overflow row count
    L.off_scratch = take((size_t)L.fallback_ctas * L.mp * 4);   // exact score rows of the fallback kernel
    L.total = o;
    return L;
}

int64_t topn_fused_workspace_bytes(int64_t batch_rows, int32_t n_items, int32_t K) {
    return (int64_t)fused_layout(batch_rows, n_items, K).total;
}

static int launch_filter(FilterArgs fa, int tile_begin, int tile_end, int dense, size_t smem, cudaStream_t s) {
    fa.tile_begin = tile_begin;
    fa.tile_end = tile_end;
    fa.dense = dense;
    const int64_t total = (int64_t)fa.n_rowpairs * (tile_end - tile_begin);   // (row block, item tile) steps
    const int64_t per_cta = cdiv(total, kNumSMs);
    fa.steps_per_cta = (int32_t)per_cta;
    topn_filter_kernel<<<(unsigned)cdiv(total, per_cta), kFuseThreads, smem, s>>>(fa);
    PMF_LAUNCH_CHECK();
    return PMF_OK;
}

int topn_fused_run(const TopnProblem& p, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
    const FusedLayout L = fused_layout(p.batch_rows, p.n_items, p.K);
    PMF_REQUIRE((int64_t)L.total <= workspace_bytes, "workspace too small for the fused top-n path");
    PMF_REQUIRE(((uintptr_t)workspace & 15) == 0, "top-n workspace must be 16-byte aligned");
    uint8_t* ws = (uint8_t*)workspace;
    __nv_bfloat16* A_pack = (__nv_bfloat16*)(ws + L.off_A);
    __nv_bfloat16* B_pack = (__nv_bfloat16*)(ws + L.off_B);
    int32_t* cand_idx = (int32_t*)(ws + L.off_idx);
    float* cand_score = (float*)(ws + L.off_score);
    int32_t* cand_cnt = (int32_t*)(ws + L.off_cnt);
    float* thr = (float*)(ws + L.off_thr);
    int32_t* ovf_list = (int32_t*)(ws + L.off_ovf);
    unsigned* maxnorm = (unsigned*)(ws + L.off_misc);
    int32_t* ovf_count = (int32_t*)(ws + L.off_misc) + 1;
    float* scratch = (float*)(ws + L.off_scratch);
    const int kp16 = (int)L.kp16;
    const int tiles = (int)(L.mp / kTile);

    PMF_TRY(topn_launch_maxnorm(p.F_item, p.n_items, p.K, p.ld, maxnorm, s));
    PMF_TRY(topn_launch_pack(p.F_user, p.user_rows, p.batch_rows, L.bp, p.K, p.ld, kp16, A_pack, s));
    PMF_TRY(topn_launch_pack(p.F_item, nullptr, p.n_items, L.mp, p.K, p.ld, kp16, B_pack, s));
    const int level0_tiles = tiles < kLevel0Tiles ? tiles : kLevel0Tiles;
    const int32_t cnt0 = (int32_t)((int64_t)level0_tiles * kTile < p.n_items ? (int64_t)level0_tiles * kTile : p.n_items);
    topn_fused_init_kernel<<<(unsigned)cdiv(L.bp, 256), 256, 0, s>>>(thr, cand_cnt, ovf_count, p.batch_rows, L.bp, cnt0);
    PMF_LAUNCH_CHECK();

    const int stages = fused_stages(kp16);
    const size_t smem = (size_t)(2 + stages) * kTile * kp16 * 2 + kExtraBytes;
    PMF_CUDA(cudaFuncSetAttribute(topn_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t refine_smem = (size_t)kFuseCap * 8 + (size_t)((p.K + 3) & ~3) * 4;      // + kept arrays at the final level
    const size_t final_smem = refine_smem + (size_t)kCandCap * 8;
    PMF_CUDA(cudaFuncSetAttribute(topn_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)final_smem));

    FilterArgs fa;
    fa.A_pack = A_pack; fa.B_pack = B_pack; fa.kp16 = kp16; fa.stages = stages;
    fa.n_rowpairs = (int32_t)(L.bp / kFuseRows); fa.n_items = p.n_items;
    fa.thr = thr; fa.cand_idx = cand_idx; fa.cand_score = cand_score; fa.cand_cnt = cand_cnt;
    RefineArgs ra;
    ra.cand_idx = cand_idx; ra.cand_score = cand_score; ra.cand_cnt = cand_cnt; ra.thr = thr; ra.n = p.n;
    ra.F_user = p.F_user; ra.F_item = p.F_item; ra.rows = p.user_rows; ra.K = p.K; ra.ld = p.ld;
    ra.item_maxnorm2_bits = maxnorm; ra.idx_out = p.idx_out; ra.score_out = p.score_out; ra.stats = p.stats;
    ra.ovf_list = ovf_list; ra.ovf_count = ovf_count;

    PMF_TRY(launch_filter(fa, 0, level0_tiles, 1, smem, s));
    int covered = level0_tiles;
    int growth = 256 / p.n < 2 ? 2 : 256 / p.n;           // a level adds ~ n * growth entries per row
    if (g_tune_topn_growth >= 1) growth = g_tune_topn_growth;
    ra.dense_input = 1;
    while (covered < tiles) {
        ra.final_level = 0;
        topn_refine_kernel<<<(unsigned)p.batch_rows, kRefineThreads, refine_smem, s>>>(ra);
        PMF_LAUNCH_CHECK();
        ra.dense_input = 0;
        const int64_t want = (int64_t)covered * (1 + growth);
        const int next = (int)(want < tiles ? want : tiles);
        PMF_TRY(launch_filter(fa, covered, next, 0, smem, s));
        covered = next;
    }
    ra.final_level = 1;
    topn_refine_kernel<<<(unsigned)p.batch_rows, kRefineThreads, final_smem, s>>>(ra);
    PMF_LAUNCH_CHECK();

    SelArgs sa;
    sa.S = nullptr; sa.m_padded = L.mp; sa.n_items = p.n_items; sa.n = p.n; sa.approx = 0;
    sa.F_user = p.F_user; sa.F_item = p.F_item; sa.rows = p.user_rows; sa.K = p.K; sa.ld = p.ld;
    sa.item_maxnorm2_bits = maxnorm; sa.idx_out = p.idx_out; sa.score_out = p.score_out; sa.stats = p.stats;
    return topn_launch_fallback(sa, ovf_list, ovf_count, scratch, L.fallback_ctas, s);
}

}  // namespace pmf
