// Shared helpers for libpmf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "pmf_b200.h"

namespace pmf {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);

#define PMF_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            pmf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PMF_ECUDA;                                                               \
        }                                                                                   \
    } while (0)

#define PMF_REQUIRE(cond, ...)             \
    do {                                   \
        if (!(cond)) {                     \
            pmf::set_error(__VA_ARGS__);   \
            return PMF_EINVAL;             \
        }                                  \
    } while (0)

#define PMF_TRY(expr)              \
    do {                           \
        int _s = (expr);           \
        if (_s != PMF_OK) return _s; \
    } while (0)

#define PMF_LAUNCH_CHECK() PMF_CUDA(cudaGetLastError())

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Stream-ordered allocation from the LIBRARY's own memory pool of the current device (csr_build.cu): a bounded release
// threshold keeps a re-fit's buffers cached without touching the process-wide default pool (pmf_trim() empties it).
cudaError_t pool_alloc(void** p, size_t bytes, cudaStream_t s);

// Stream-ordered scratch allocation (freed on the same stream).
template <typename T>
inline int alloc_async(T** p, int64_t count, cudaStream_t s) {
    *p = nullptr;
    if (count <= 0) count = 1;
    PMF_CUDA(pool_alloc(reinterpret_cast<void**>(p), static_cast<size_t>(count) * sizeof(T), s));
    return PMF_OK;
}
template <typename T>
inline void free_async(T* p, cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
}

// Exclusive prefix sum of int32 (n up to 2^31-1), d_out may alias d_in.  If d_total != nullptr
// the grand total is written there.  Defined in scan_sort.cu.
int exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, int64_t n, int32_t* d_total, cudaStream_t s);

// Read-only view of a pmf_csr for the kernels in other translation units (csr_build.cu).
struct CsrView {
    int64_t nnz;
    int32_t n_rows, row_offset, seg_len, n_seg, n_multi, n_partial;
    const int32_t *row_ptr, *col;
    const float* val;
    const int32_t *seg_row, *seg_start, *seg_partial, *seg_order, *row_seg, *multi_row, *multi_first;
    // segments in processing order (longest first), one 16-byte record each: {row, first observation, one past the last
    // observation, partial-sum slot or -1} -- what a lane group needs to start, in ONE load instead of a chain of three
    const int4* seg_desc;
    int32_t n_cols;   // 1 + largest id of the other side in `col` (extent of the gathered table that can be touched)
    int32_t col_lo;   // smallest id of the other side in `col`
};
CsrView csr_view(const pmf_csr* c);

// 128-bit read-only streaming loads.
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace pmf
