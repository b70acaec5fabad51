// a1: observation grouping on the device -- stable LSD radix sort + row pointers + the
// segment decomposition consumed by the pass kernels.
//
// Replaces the reference's _build_index_lists (poisson_mf_cavi.py:73-84 and its copies): a
// Python loop that appends observation t to the list of row key[t].  The result is the stable
// sort of arange(nnz) by key, which is what the radix sort below produces bit for bit.
//
// HBM-bound integer work: every pass streams (key, index) pairs once in and once out; ranks
// are computed with warp match + shared-memory counters, no global atomics.
#include <stdarg.h>

#include <stdlib.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace pmf {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------
// exclusive scan (int32), hierarchical: 4096 items per block
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_tile_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                                 int64_t n, int32_t* __restrict__ tile_sums) {
    __shared__ int32_t warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int32_t v[kScanItems];
    int32_t sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    int32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_tot[lane];
        int32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_tot[lane] = wi - w;  // exclusive over warps
        if (lane == 31 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    int32_t run = warp_tot[warp] + incl - sum;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

__global__ void scan_add_kernel(int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ tile_offsets) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_offsets[i / kScanTile];
}

__global__ void scan_total_kernel(const int32_t* __restrict__ excl, const int32_t* __restrict__ last_in,
                                  int32_t* __restrict__ total) {
    *total = *excl + *last_in;
}

int exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, int64_t n, int32_t* d_total, cudaStream_t s) {
    if (n <= 0) {
        if (d_total) PMF_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int32_t), s));
        return PMF_OK;
    }
    // the total needs the last input element, which an in-place scan overwrites: save it first
    int32_t* d_last = nullptr;
    if (d_total) {
        PMF_TRY(alloc_async(&d_last, 1, s));
        PMF_CUDA(cudaMemcpyAsync(d_last, d_in + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    }
    const int64_t tiles = cdiv(n, kScanTile);
    int32_t* d_sums = nullptr;
    if (tiles > 1) PMF_TRY(alloc_async(&d_sums, tiles, s));
    scan_tile_kernel<<<(unsigned)tiles, kScanThreads, 0, s>>>(d_in, d_out, n, d_sums);
    PMF_LAUNCH_CHECK();
    if (tiles > 1) {
        PMF_TRY(exclusive_scan_i32(d_sums, d_sums, tiles, nullptr, s));
        scan_add_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(d_out, n, d_sums);
        PMF_LAUNCH_CHECK();
        free_async(d_sums, s);
    }
    if (d_total) {
        scan_total_kernel<<<1, 1, 0, s>>>(d_out + (n - 1), d_last, d_total);
        PMF_LAUNCH_CHECK();
        free_async(d_last, s);
    }
    return PMF_OK;
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort of (key, original index), 8-bit digits, 4096-item tiles
// ------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 16;                       // items per thread
constexpr int kSortTile = kSortThreads * kSortRounds;  // 4096
constexpr int kWarpChunk = 32 * kSortRounds;           // each warp ranks a contiguous 512-item chunk

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int shift,
                                                                  int32_t* __restrict__ hist, int n_tiles) {
    __shared__ int32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t idx = base + r * kSortThreads + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255], 1);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];  // digit-major
}

template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const int32_t* __restrict__ keys_in,
                                                                     const int32_t* __restrict__ vals_in,
                                                                     int32_t* __restrict__ keys_out,
                                                                     int32_t* __restrict__ vals_out, int64_t n, int shift,
                                                                     const int32_t* __restrict__ offsets, int n_tiles) {
    __shared__ int32_t cnt[kSortThreads / 32][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int d = threadIdx.x; d < (kSortThreads / 32) * 256; d += kSortThreads) (&cnt[0][0])[d] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * kWarpChunk;
    int32_t key[kSortRounds], val[kSortRounds], rank[kSortRounds];
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        const bool ok = idx < n;
        key[r] = ok ? keys_in[idx] : 0;
        val[r] = FIRST ? (int32_t)idx : (ok ? vals_in[idx] : 0);
    }
    // rank inside this warp's chunk, in input order (round-major, then lane): stable
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const bool ok = (wbase + r * 32 + lane) < n;
        const int d = ok ? ((key[r] >> shift) & 255) : (256 + lane);  // inactive lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int below = __popc(peers & ((1u << lane) - 1u));
        const int old = ok ? cnt[warp][d] : 0;
        __syncwarp();
        if (ok && below == 0) cnt[warp][d] = old + __popc(peers);
        __syncwarp();
        rank[r] = old + below;
    }
    __syncthreads();
    {   // digit d: global offset of (digit, tile) then exclusive prefix over the tile's warps
        const int d = threadIdx.x;
        int32_t run = offsets[(int64_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) {
            const int32_t t = cnt[w][d];
            cnt[w][d] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        if ((wbase + r * 32 + lane) < n) {
            const int32_t pos = cnt[warp][(key[r] >> shift) & 255] + rank[r];
            keys_out[pos] = key[r];
            vals_out[pos] = val[r];
        }
    }
}

// Stable sort of (key, index) pairs by key in [0, 2^bits): d_sorted_keys / d_perm receive the result.
// perm[p] = original position of the p-th smallest key (ties in input order).
static int stable_sort_by_key(const int32_t* d_key, int64_t n, int bits, int32_t* d_sorted_keys, int32_t* d_perm,
                              cudaStream_t s) {
    const int passes = (bits + 7) / 8 > 0 ? (bits + 7) / 8 : 1;
    const int n_tiles = (int)cdiv(n, kSortTile);
    int32_t *k1 = nullptr, *v1 = nullptr, *hist = nullptr;
    PMF_TRY(alloc_async(&hist, (int64_t)256 * n_tiles, s));
    if (passes > 1) {
        PMF_TRY(alloc_async(&k1, n, s));
        PMF_TRY(alloc_async(&v1, n, s));
    }
    // ping-pong so that the LAST pass lands in (d_sorted_keys, d_perm)
    const int32_t* kin = d_key;
    const int32_t* vin = nullptr;
    for (int p = 0; p < passes; ++p) {
        const bool to_final = ((passes - 1 - p) % 2) == 0;
        int32_t* kout = to_final ? d_sorted_keys : k1;
        int32_t* vout = to_final ? d_perm : v1;
        radix_hist_kernel<<<n_tiles, kSortThreads, 0, s>>>(kin, n, 8 * p, hist, n_tiles);
        PMF_LAUNCH_CHECK();
        PMF_TRY(exclusive_scan_i32(hist, hist, (int64_t)256 * n_tiles, nullptr, s));
        if (p == 0)
            radix_scatter_kernel<true><<<n_tiles, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, 8 * p, hist, n_tiles);
        else
            radix_scatter_kernel<false><<<n_tiles, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, 8 * p, hist, n_tiles);
        PMF_LAUNCH_CHECK();
        kin = kout;
        vin = vout;
    }
    free_async(k1, s);
    free_async(v1, s);
    free_async(hist, s);
    return PMF_OK;
}

static int bit_length(int64_t max_value) {
    int bits = 1;
    while (bits < 31 && (1ll << bits) <= max_value) ++bits;
    return bits;
}

// row_ptr from sorted keys: the thread at position p closes every row in (key[p-1], key[p]]
__global__ void row_ptr_kernel(const int32_t* __restrict__ sorted_keys, int64_t n, int32_t n_rows,
                               int32_t* __restrict__ row_ptr) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t k = sorted_keys[p];
    const int32_t kp = (p == 0) ? -1 : sorted_keys[p - 1];
    for (int32_t r = kp + 1; r <= k; ++r) row_ptr[r] = (int32_t)p;
    if (p == n - 1)
        for (int32_t r = k + 1; r <= n_rows; ++r) row_ptr[r] = (int32_t)n;
}

__global__ void gather_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ other,
                              const float* __restrict__ x, int64_t n, int32_t* __restrict__ col, float* __restrict__ val) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t t = perm[p];
    col[p] = other[t];
    val[p] = x[t];
}

__global__ void check_keys_kernel(const int32_t* __restrict__ key, const int32_t* __restrict__ other, int64_t n,
                                  int32_t n_rows, int32_t* __restrict__ bad) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int oth = -1;
    if (p < n) {
        oth = other[p];
        if (key[p] < 0 || key[p] >= n_rows || oth < 0) atomicOr(bad, 1);
    }
    // bad[1] / bad[2] = largest / smallest id of the other side: the span of the other side's table the passes can touch
    const int wmax = __reduce_max_sync(0xffffffffu, oth);
    const int wmin = __reduce_min_sync(0xffffffffu, oth < 0 ? INT32_MAX : oth);
    if ((threadIdx.x & 31) == 0 && wmax >= 0) {
        atomicMax(bad + 1, wmax);
        atomicMin(bad + 2, wmin);
    }
}

// ------------------------------------------------------------------------------------------
// routing helpers for sharded / tiled rating lists
// ------------------------------------------------------------------------------------------
__global__ void count_keys_kernel(const int32_t* __restrict__ key, int64_t n, int32_t n_bins, int32_t* __restrict__ counts,
                                  int32_t* __restrict__ bad) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t k = key[p];
    if (k < 0 || k >= n_bins) { atomicOr(bad, 1); return; }
    atomicAdd(counts + k, 1);
}

// bucket[p] = b with bounds[b] <= key[p] < bounds[b+1]  (bounds ascending, bounds[0] <= every key < bounds[nb])
__global__ void bucket_kernel(const int32_t* __restrict__ key, int64_t n, const int32_t* __restrict__ bounds, int32_t nb,
                              int32_t* __restrict__ bucket) {
    __shared__ int32_t sb[257];
    for (int t = threadIdx.x; t <= nb; t += blockDim.x) sb[t] = bounds[t];
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t k = key[p];
    int lo = 0, hi = nb;   // invariant: sb[lo] <= k < sb[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (sb[mid] <= k) lo = mid; else hi = mid;
    }
    bucket[p] = lo;
}

__global__ void gather3_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                               const float* __restrict__ x, int64_t n, int32_t* __restrict__ a_out, int32_t* __restrict__ b_out,
                               float* __restrict__ x_out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t t = perm[p];
    a_out[p] = a[t];
    b_out[p] = b[t];
    x_out[p] = x[t];
}

// ------------------------------------------------------------------------------------------
// segment decomposition
// ------------------------------------------------------------------------------------------
__global__ void seg_count_kernel(const int32_t* __restrict__ row_ptr, int32_t n_rows, int32_t seg_len,
                                 int32_t* __restrict__ seg_cnt, int32_t* __restrict__ multi_flag,
                                 int32_t* __restrict__ part_cnt) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int32_t len = row_ptr[r + 1] - row_ptr[r];
    const int32_t ns = len <= seg_len ? 1 : (len + seg_len - 1) / seg_len;
    seg_cnt[r] = ns;
    multi_flag[r] = ns > 1;
    part_cnt[r] = ns > 1 ? ns : 0;
}

__global__ void seg_fill_kernel(const int32_t* __restrict__ row_ptr, int32_t n_rows, int32_t seg_len,
                                const int32_t* __restrict__ seg_off, const int32_t* __restrict__ multi_off,
                                const int32_t* __restrict__ part_off, int32_t n_multi, int32_t n_partial,
                                int32_t* __restrict__ seg_row, int32_t* __restrict__ seg_start,
                                int32_t* __restrict__ seg_partial, int32_t* __restrict__ multi_row,
                                int32_t* __restrict__ multi_first) {
    const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) multi_first[n_multi] = n_partial;
    if (r >= n_rows) return;
    const int32_t b = row_ptr[r], e = row_ptr[r + 1];
    const int32_t len = e - b;
    const int32_t ns = len <= seg_len ? 1 : (len + seg_len - 1) / seg_len;
    const int32_t s0 = seg_off[r];
    if (ns == 1) {
        seg_row[s0] = r;
        seg_start[s0] = b;
        seg_partial[s0] = -1;
        return;
    }
    const int32_t p0 = part_off[r];
    multi_row[multi_off[r]] = r;
    multi_first[multi_off[r]] = p0;
    for (int32_t j = 0; j < ns; ++j) {
        seg_row[s0 + j] = r;
        seg_start[s0 + j] = b + j * seg_len;
        seg_partial[s0 + j] = p0 + j;
    }
}

// sort key of a segment: seg_len - (its length), so that an ascending stable sort lists long segments first
__global__ void seg_key_kernel(const int32_t* __restrict__ seg_row, const int32_t* __restrict__ seg_start,
                               const int32_t* __restrict__ row_ptr, int32_t n_seg, int32_t seg_len,
                               int32_t* __restrict__ key) {
    const int32_t sidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= n_seg) return;
    const int32_t b = seg_start[sidx];
    const int32_t e = min(b + seg_len, row_ptr[seg_row[sidx] + 1]);
    key[sidx] = seg_len - (e - b);
}

__global__ void rebase_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t n, int32_t base) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] - base;
}

}  // namespace pmf

using namespace pmf;

struct pmf_csr {
    int64_t nnz = 0;
    int32_t n_rows = 0, row_offset = 0, seg_len = 0;
    int32_t n_cols = 0;            // 1 + largest id of the other side seen in `col`
    int32_t col_lo = 0;            // smallest id of the other side seen in `col`
    int32_t n_seg = 0, n_multi = 0, n_partial = 0;
    int32_t *row_ptr = nullptr, *perm = nullptr, *col = nullptr;
    float* val = nullptr;
    int32_t *seg_row = nullptr, *seg_start = nullptr, *seg_partial = nullptr;
    int32_t* seg_order = nullptr;  // segment ids, longest first: a warp's groups get equally long segments
    cudaStream_t alloc_stream = nullptr;   // stream the buffers were allocated on (the build stream; synchronised at the end)
    int4* seg_desc = nullptr;      // [n_seg] {row, start, end, partial slot} in seg_order order
    int32_t* row_seg = nullptr;    // [n_rows+1] first segment of each row
    int32_t *multi_row = nullptr, *multi_first = nullptr;
    int64_t bytes = 0;
};

// The rating list's buffers (and the library's scratch) come from a LIBRARY-OWNED stream-ordered memory pool, one per
// device, so a model that is fitted again (tuning loops fit dozens of times; bench.py's timed fit follows an untimed one)
// gets the previous fit's memory back without paying the driver for ~30 fresh allocations -- on a fresh box those
// cudaMalloc / cudaFree calls were measured at up to 1 s per fit of the 100 M-rating config.  The pool keeps at most
// PMF_POOL_KEEP_MB (default 8192) MB of freed memory; the process-wide default pool is never modified, and pmf_trim()
// returns everything that is not in use to the driver (for callers that go on to allocate through other allocators).
static std::mutex g_pool_mutex;
static cudaMemPool_t g_pools[64] = {};
static bool g_pool_failed[64] = {};

static cudaMemPool_t library_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (g_pools[dev] || g_pool_failed[dev]) return g_pools[dev];
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) {
        cudaGetLastError();   // the pool is an optimisation: fall back to the default pool
        g_pool_failed[dev] = true;
        return nullptr;
    }
    uint64_t keep_mb = 8192;
    if (const char* e = getenv("PMF_POOL_KEEP_MB")) keep_mb = strtoull(e, nullptr, 10);
    uint64_t threshold = keep_mb << 20;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    cudaGetLastError();
    g_pools[dev] = pool;
    return pool;
}

cudaError_t pmf::pool_alloc(void** p, size_t bytes, cudaStream_t s) {
    if (cudaMemPool_t pool = library_pool()) {
        const cudaError_t e = cudaMallocFromPoolAsync(p, bytes, pool, s);
        if (e == cudaSuccess) return e;
        cudaGetLastError();
    }
    return cudaMallocAsync(p, bytes, s);
}

static int dev_alloc(void** p, int64_t bytes, pmf_csr* c) {
    if (bytes <= 0) bytes = 4;
    cudaError_t e = pool_alloc(p, (size_t)bytes, c->alloc_stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(p, (size_t)bytes);   // e.g. a pool that cannot grow: fall back to a plain allocation
    }
    if (e != cudaSuccess) {
        set_error("device allocation of %lld bytes failed: %s", (long long)bytes, cudaGetErrorString(e));
        return PMF_ENOMEM;
    }
    c->bytes += bytes;
    return PMF_OK;
}

__global__ void seg_pack_kernel(const int32_t* __restrict__ seg_order, const int32_t* __restrict__ seg_row,
                                const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_partial,
                                const int32_t* __restrict__ row_ptr, int32_t n_seg, int32_t seg_len, int4* __restrict__ out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_seg) return;
    const int sidx = seg_order[g];
    const int row = seg_row[sidx], p = seg_start[sidx];
    out[g] = make_int4(row, p, min(p + seg_len, row_ptr[row + 1]), seg_partial[sidx]);
}

// Builds seg_* / multi_* for c (row_ptr must be final).  Synchronises the stream.
static int build_segments(pmf_csr* c, cudaStream_t s) {
    const int32_t R = c->n_rows;
    int32_t *seg_cnt = nullptr, *multi_flag = nullptr, *part_cnt = nullptr, *totals = nullptr;
    PMF_TRY(alloc_async(&seg_cnt, R, s));
    PMF_TRY(alloc_async(&multi_flag, R, s));
    PMF_TRY(alloc_async(&part_cnt, R, s));
    PMF_TRY(alloc_async(&totals, 3, s));
    const unsigned grid = (unsigned)cdiv(R > 0 ? R : 1, 256);
    seg_count_kernel<<<grid, 256, 0, s>>>(c->row_ptr, R, c->seg_len, seg_cnt, multi_flag, part_cnt);
    PMF_LAUNCH_CHECK();
    PMF_TRY(exclusive_scan_i32(seg_cnt, seg_cnt, R, totals + 0, s));
    PMF_TRY(exclusive_scan_i32(multi_flag, multi_flag, R, totals + 1, s));
    PMF_TRY(exclusive_scan_i32(part_cnt, part_cnt, R, totals + 2, s));
    int32_t h_tot[3] = {0, 0, 0};
    PMF_CUDA(cudaMemcpyAsync(h_tot, totals, sizeof(h_tot), cudaMemcpyDeviceToHost, s));
    PMF_CUDA(cudaStreamSynchronize(s));
    c->n_seg = h_tot[0];
    c->n_multi = h_tot[1];
    c->n_partial = h_tot[2];
    PMF_TRY(dev_alloc((void**)&c->seg_row, (int64_t)c->n_seg * 4, c));
    PMF_TRY(dev_alloc((void**)&c->seg_start, (int64_t)c->n_seg * 4, c));
    PMF_TRY(dev_alloc((void**)&c->seg_partial, (int64_t)c->n_seg * 4, c));
    PMF_TRY(dev_alloc((void**)&c->multi_row, (int64_t)c->n_multi * 4, c));
    PMF_TRY(dev_alloc((void**)&c->multi_first, ((int64_t)c->n_multi + 1) * 4, c));
    seg_fill_kernel<<<grid, 256, 0, s>>>(c->row_ptr, R, c->seg_len, seg_cnt, multi_flag, part_cnt, c->n_multi,
                                         c->n_partial, c->seg_row, c->seg_start, c->seg_partial, c->multi_row,
                                         c->multi_first);
    PMF_LAUNCH_CHECK();
    PMF_TRY(dev_alloc((void**)&c->row_seg, ((int64_t)R + 1) * 4, c));
    PMF_CUDA(cudaMemcpyAsync(c->row_seg, seg_cnt, (size_t)R * 4, cudaMemcpyDeviceToDevice, s));
    PMF_CUDA(cudaMemcpyAsync(c->row_seg + R, &c->n_seg, 4, cudaMemcpyHostToDevice, s));
    free_async(seg_cnt, s);
    free_async(multi_flag, s);
    free_async(part_cnt, s);
    free_async(totals, s);
    // processing order: segments sorted by length, longest first (stable -> deterministic)
    PMF_TRY(dev_alloc((void**)&c->seg_order, (int64_t)c->n_seg * 4, c));
    if (c->n_seg > 0) {
        int32_t *key = nullptr, *sorted = nullptr;
        PMF_TRY(alloc_async(&key, c->n_seg, s));
        PMF_TRY(alloc_async(&sorted, c->n_seg, s));
        seg_key_kernel<<<(unsigned)cdiv(c->n_seg, 256), 256, 0, s>>>(c->seg_row, c->seg_start, c->row_ptr, c->n_seg,
                                                                     c->seg_len, key);
        PMF_LAUNCH_CHECK();
        PMF_TRY(stable_sort_by_key(key, c->n_seg, bit_length(c->seg_len), sorted, c->seg_order, s));
        free_async(key, s);
        free_async(sorted, s);
    }
    PMF_TRY(dev_alloc((void**)&c->seg_desc, (int64_t)c->n_seg * 16, c));
    if (c->n_seg > 0) {
        seg_pack_kernel<<<(unsigned)cdiv(c->n_seg, 256), 256, 0, s>>>(c->seg_order, c->seg_row, c->seg_start, c->seg_partial,
                                                                      c->row_ptr, c->n_seg, c->seg_len, c->seg_desc);
        PMF_LAUNCH_CHECK();
    }
    PMF_CUDA(cudaStreamSynchronize(s));
    return PMF_OK;
}

extern "C" {

int pmf_version(void) { return 100; }
const char* pmf_last_error(void) { return g_err; }

int pmf_device_count(int* count) {
    PMF_REQUIRE(count != nullptr, "count is NULL");
    PMF_CUDA(cudaGetDeviceCount(count));
    return PMF_OK;
}

int pmf_copy_to_host(void* h_dst, const void* d_src, int64_t bytes, void* stream) {
    PMF_REQUIRE(bytes >= 0 && (bytes == 0 || (h_dst && d_src)), "bad argument");
    if (bytes == 0) return PMF_OK;
    PMF_CUDA(cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PMF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PMF_OK;
}

int pmf_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream) {
    PMF_REQUIRE(bytes >= 0 && (bytes == 0 || (dst && src)), "bad argument");
    if (bytes == 0) return PMF_OK;
    // unified addressing: either pointer may be a peer GPU's memory mapped into this process; the copy engines move it
    PMF_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return PMF_OK;
}

int pmf_trim(void) {
    PMF_CUDA(cudaDeviceSynchronize());
    if (cudaMemPool_t pool = library_pool()) PMF_CUDA(cudaMemPoolTrimTo(pool, 0));
    return PMF_OK;
}

int pmf_row_stride(int K) { return K <= 0 ? 0 : ((K + 7) / 8) * 8; }

int pmf_csr_free(pmf_csr* c) {
    if (!c) return PMF_OK;
    void* ptrs[] = {c->seg_desc, c->row_ptr, c->perm, c->col, c->val, c->seg_row, c->seg_start, c->seg_partial, c->seg_order,
                    c->row_seg, c->multi_row, c->multi_first};
    // One device-wide synchronisation (no kernel on any stream can still be reading the list), then stream-ordered frees
    // that hand the buffers back to the pool without a driver round trip each.
    cudaDeviceSynchronize();
    for (void* p : ptrs) {
        if (!p) continue;
        if (cudaFreeAsync(p, c->alloc_stream) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(p);
        }
    }
    delete c;
    return PMF_OK;
}

int pmf_csr_build(const int32_t* d_key, const int32_t* d_other, const float* d_val, int64_t nnz, int32_t n_rows,
                  int32_t seg_len, void* stream, pmf_csr** out) {
    PMF_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    PMF_REQUIRE(nnz >= 0 && nnz < (int64_t)INT32_MAX, "nnz=%lld out of range", (long long)nnz);
    PMF_REQUIRE(n_rows > 0, "n_rows=%d must be positive", n_rows);
    PMF_REQUIRE(seg_len >= 8 && seg_len % 8 == 0 && seg_len <= 65536, "seg_len=%d must be a multiple of 8 in [8, 65536]", seg_len);
    PMF_REQUIRE(nnz == 0 || (d_key && d_other && d_val), "NULL input with nnz > 0");
    cudaStream_t s = (cudaStream_t)stream;
    pmf_csr* c = new pmf_csr();
    c->alloc_stream = s;
    c->nnz = nnz;
    c->n_rows = n_rows;
    c->seg_len = seg_len;
    int st = PMF_OK;
    auto fail = [&](int code) {
        pmf_csr_free(c);
        return code;
    };
    if ((st = dev_alloc((void**)&c->row_ptr, ((int64_t)n_rows + 1) * 4, c)) != PMF_OK) return fail(st);
    if ((st = dev_alloc((void**)&c->perm, nnz * 4, c)) != PMF_OK) return fail(st);
    if ((st = dev_alloc((void**)&c->col, nnz * 4, c)) != PMF_OK) return fail(st);
    if ((st = dev_alloc((void**)&c->val, nnz * 4, c)) != PMF_OK) return fail(st);

    if (nnz == 0) {
        if (cudaMemsetAsync(c->row_ptr, 0, ((size_t)n_rows + 1) * 4, s) != cudaSuccess) return fail(PMF_ECUDA);
    } else {
        auto body = [&]() -> int {
            int32_t* bad = nullptr;
            PMF_TRY(alloc_async(&bad, 3, s));
            int32_t h_bad[3] = {0, 0, INT32_MAX};
            PMF_CUDA(cudaMemcpyAsync(bad, h_bad, 12, cudaMemcpyHostToDevice, s));
            check_keys_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(d_key, d_other, nnz, n_rows, bad);
            PMF_LAUNCH_CHECK();
            PMF_CUDA(cudaMemcpyAsync(h_bad, bad, 12, cudaMemcpyDeviceToHost, s));
            PMF_CUDA(cudaStreamSynchronize(s));
            free_async(bad, s);
            PMF_REQUIRE(h_bad[0] == 0, "ids out of range: keys must lie in [0, %d), other ids must be >= 0", n_rows);
            c->n_cols = h_bad[1] + 1;
            c->col_lo = h_bad[2] <= h_bad[1] ? h_bad[2] : 0;

            int32_t* k0 = nullptr;
            PMF_TRY(alloc_async(&k0, nnz, s));
            PMF_TRY(stable_sort_by_key(d_key, nnz, bit_length((int64_t)n_rows - 1), k0, c->perm, s));
            row_ptr_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(k0, nnz, n_rows, c->row_ptr);
            PMF_LAUNCH_CHECK();
            gather_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(c->perm, d_other, d_val, nnz, c->col, c->val);
            PMF_LAUNCH_CHECK();
            free_async(k0, s);
            return PMF_OK;
        };
        if ((st = body()) != PMF_OK) return fail(st);
    }
    if ((st = build_segments(c, s)) != PMF_OK) return fail(st);
    *out = c;
    return PMF_OK;
}

int pmf_csr_slice(const pmf_csr* src, int32_t row_begin, int32_t row_end, void* stream, pmf_csr** out) {
    PMF_REQUIRE(src && out, "NULL argument");
    *out = nullptr;
    PMF_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= src->n_rows, "bad row range [%d,%d) of %d",
                row_begin, row_end, src->n_rows);
    PMF_REQUIRE(row_end > row_begin, "empty row range");
    cudaStream_t s = (cudaStream_t)stream;
    int32_t h_b = 0, h_e = 0;
    PMF_CUDA(cudaMemcpyAsync(&h_b, src->row_ptr + row_begin, 4, cudaMemcpyDeviceToHost, s));
    PMF_CUDA(cudaMemcpyAsync(&h_e, src->row_ptr + row_end, 4, cudaMemcpyDeviceToHost, s));
    PMF_CUDA(cudaStreamSynchronize(s));
    pmf_csr* c = new pmf_csr();
    c->alloc_stream = s;
    c->nnz = h_e - h_b;
    c->n_rows = row_end - row_begin;
    c->row_offset = src->row_offset + row_begin;
    c->seg_len = src->seg_len;
    c->n_cols = src->n_cols;
    c->col_lo = src->col_lo;
    auto body = [&]() -> int {
        PMF_TRY(dev_alloc((void**)&c->row_ptr, ((int64_t)c->n_rows + 1) * 4, c));
        PMF_TRY(dev_alloc((void**)&c->perm, c->nnz * 4, c));
        PMF_TRY(dev_alloc((void**)&c->col, c->nnz * 4, c));
        PMF_TRY(dev_alloc((void**)&c->val, c->nnz * 4, c));
        rebase_kernel<<<(unsigned)cdiv(c->n_rows + 1, 256), 256, 0, s>>>(src->row_ptr + row_begin, c->row_ptr,
                                                                         c->n_rows + 1, h_b);
        PMF_LAUNCH_CHECK();
        if (c->nnz > 0) {
            PMF_CUDA(cudaMemcpyAsync(c->perm, src->perm + h_b, c->nnz * 4, cudaMemcpyDeviceToDevice, s));
            PMF_CUDA(cudaMemcpyAsync(c->col, src->col + h_b, c->nnz * 4, cudaMemcpyDeviceToDevice, s));
            PMF_CUDA(cudaMemcpyAsync(c->val, src->val + h_b, c->nnz * 4, cudaMemcpyDeviceToDevice, s));
        }
        return build_segments(c, s);
    };
    int st = body();
    if (st != PMF_OK) {
        pmf_csr_free(c);
        return st;
    }
    *out = c;
    return PMF_OK;
}

int pmf_csr_set_row_offset(pmf_csr* c, int32_t row_offset) {
    PMF_REQUIRE(c != nullptr && row_offset >= 0, "bad argument");
    c->row_offset = row_offset;
    return PMF_OK;
}

int pmf_count_keys(const int32_t* d_key, int64_t n, int32_t n_bins, int32_t* d_counts, void* stream) {
    PMF_REQUIRE(n >= 0 && n_bins > 0 && d_counts && (n == 0 || d_key), "bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    PMF_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)n_bins * 4, s));
    if (n == 0) return PMF_OK;
    int32_t* bad = nullptr;
    PMF_TRY(alloc_async(&bad, 1, s));
    PMF_CUDA(cudaMemsetAsync(bad, 0, 4, s));
    count_keys_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(d_key, n, n_bins, d_counts, bad);
    PMF_LAUNCH_CHECK();
    int32_t h_bad = 0;
    PMF_CUDA(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, s));
    PMF_CUDA(cudaStreamSynchronize(s));
    free_async(bad, s);
    PMF_REQUIRE(h_bad == 0, "ids out of range: keys must lie in [0, %d)", n_bins);
    return PMF_OK;
}

int pmf_coo_partition(const int32_t* d_u, const int32_t* d_i, const float* d_x, int64_t n, int32_t by_item,
                      const int32_t* h_bounds, int32_t n_buckets, int32_t* d_u_out, int32_t* d_i_out, float* d_x_out,
                      int64_t* h_offsets, void* stream) {
    PMF_REQUIRE(n >= 0 && n < (int64_t)INT32_MAX && h_bounds && h_offsets && n_buckets >= 1 && n_buckets <= 256,
                "bad argument (n=%lld, n_buckets=%d)", (long long)n, n_buckets);
    PMF_REQUIRE(n == 0 || (d_u && d_i && d_x && d_u_out && d_i_out && d_x_out), "NULL array with n > 0");
    for (int b = 0; b < n_buckets; ++b) PMF_REQUIRE(h_bounds[b] <= h_bounds[b + 1], "bounds must ascend");
    cudaStream_t s = (cudaStream_t)stream;
    for (int b = 0; b <= n_buckets; ++b) h_offsets[b] = 0;
    if (n == 0) return PMF_OK;
    int32_t *bounds = nullptr, *bucket = nullptr, *sorted = nullptr, *perm = nullptr, *offs = nullptr;
    PMF_TRY(alloc_async(&bounds, n_buckets + 1, s));
    PMF_TRY(alloc_async(&bucket, n, s));
    PMF_TRY(alloc_async(&sorted, n, s));
    PMF_TRY(alloc_async(&perm, n, s));
    PMF_TRY(alloc_async(&offs, n_buckets + 1, s));
    PMF_CUDA(cudaMemcpyAsync(bounds, h_bounds, (size_t)(n_buckets + 1) * 4, cudaMemcpyHostToDevice, s));
    const unsigned grid = (unsigned)cdiv(n, 256);
    bucket_kernel<<<grid, 256, 0, s>>>(by_item ? d_i : d_u, n, bounds, n_buckets, bucket);
    PMF_LAUNCH_CHECK();
    PMF_TRY(stable_sort_by_key(bucket, n, 8, sorted, perm, s));
    row_ptr_kernel<<<grid, 256, 0, s>>>(sorted, n, n_buckets, offs);
    PMF_LAUNCH_CHECK();
    gather3_kernel<<<grid, 256, 0, s>>>(perm, d_u, d_i, d_x, n, d_u_out, d_i_out, d_x_out);
    PMF_LAUNCH_CHECK();
    std::vector<int32_t> h((size_t)n_buckets + 1);
    PMF_CUDA(cudaMemcpyAsync(h.data(), offs, h.size() * 4, cudaMemcpyDeviceToHost, s));
    PMF_CUDA(cudaStreamSynchronize(s));
    for (int b = 0; b <= n_buckets; ++b) h_offsets[b] = h[b];
    free_async(bounds, s); free_async(bucket, s); free_async(sorted, s); free_async(perm, s); free_async(offs, s);
    return PMF_OK;
}

int64_t pmf_csr_nnz(const pmf_csr* c) { return c ? c->nnz : -1; }
int32_t pmf_csr_rows(const pmf_csr* c) { return c ? c->n_rows : -1; }
int32_t pmf_csr_row_offset(const pmf_csr* c) { return c ? c->row_offset : -1; }
int32_t pmf_csr_segments(const pmf_csr* c) { return c ? c->n_seg : -1; }
int32_t pmf_csr_multi_rows(const pmf_csr* c) { return c ? c->n_multi : -1; }
int32_t pmf_csr_seg_len(const pmf_csr* c) { return c ? c->seg_len : -1; }
const int32_t* pmf_csr_row_ptr(const pmf_csr* c) { return c ? c->row_ptr : nullptr; }
const int32_t* pmf_csr_perm(const pmf_csr* c) { return c ? c->perm : nullptr; }
const int32_t* pmf_csr_col(const pmf_csr* c) { return c ? c->col : nullptr; }
const float* pmf_csr_val(const pmf_csr* c) { return c ? c->val : nullptr; }
int64_t pmf_csr_device_bytes(const pmf_csr* c) { return c ? c->bytes : -1; }

int pmf_csr_partition(const pmf_csr* c, int32_t parts, int32_t* h_bounds) {
    PMF_REQUIRE(c && h_bounds && parts >= 1, "bad argument");
    std::vector<int32_t> rp((size_t)c->n_rows + 1);
    PMF_CUDA(cudaMemcpy(rp.data(), c->row_ptr, rp.size() * 4, cudaMemcpyDeviceToHost));
    h_bounds[0] = 0;
    for (int32_t p = 1; p < parts; ++p) {
        const int64_t target = (c->nnz * p) / parts;
        // first row whose start is >= target, never before the previous boundary
        int32_t lo = h_bounds[p - 1], hi = c->n_rows;
        while (lo < hi) {
            const int32_t mid = lo + (hi - lo) / 2;
            if ((int64_t)rp[mid] < target) lo = mid + 1; else hi = mid;
        }
        h_bounds[p] = lo;
    }
    h_bounds[parts] = c->n_rows;
    return PMF_OK;
}

}  // extern "C"

// Internal accessors for the other translation units.
namespace pmf {
CsrView csr_view(const pmf_csr* c) {
    return CsrView{c->nnz, c->n_rows, c->row_offset, c->seg_len, c->n_seg, c->n_multi, c->n_partial, c->row_ptr,
                   c->col, c->val, c->seg_row, c->seg_start, c->seg_partial, c->seg_order, c->row_seg,
                   c->multi_row, c->multi_first, c->seg_desc, c->n_cols, c->col_lo};
}
}  // namespace pmf
