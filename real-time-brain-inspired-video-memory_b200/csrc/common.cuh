// common.cuh -- shared device/host helpers for libvidmem (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/vidmem.h"

namespace vm {

// ---- error plumbing --------------------------------------------------------------------
void set_error(const char *fmt, ...);
#define VM_CUDA_CHECK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            vm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return VM_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)
#define VM_REQUIRE(cond, code, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            vm::set_error(__VA_ARGS__);      \
            return (code);                   \
        }                                    \
    } while (0)

// ---- candidate keys --------------------------------------------------------------------
// A candidate is one u64: high word = order-preserving image of the fp32 score, low word =
// ~row.  Larger key == better candidate (higher score, then LOWER row), keys are unique per
// row, and 0 is reserved for "empty slot" (every real score maps to a high word >= 0x007FFFFF).
__host__ __device__ __forceinline__ uint32_t f32_ordered(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_ordered(uint32_t o)
{
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row)
{
    return ((uint64_t)f32_ordered(score) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFu); }
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return f32_from_ordered((uint32_t)(k >> 32)); }

// ---- binary64 reference arithmetic -----------------------------------------------------
// Python's builtin sum() over float products (see oracle/vm_oracle.c): the first item enters
// through 0.0 + x, the rest through a naive or Neumaier recurrence.  Explicit _rn intrinsics
// keep nvcc from contracting mul+add into FMA, so every step rounds like CPython's doubles.
struct RefSum {
    double s, c;
    __device__ __forceinline__ void init() { s = 0.0; c = 0.0; }
    // Branch-free: with s = c = 0 the general step reproduces CPython's first step (0.0 + x)
    // exactly, so no "first item" case is needed; selects instead of an if keep the loop
    // unrollable and the loads pipelined.
    template <bool NEUMAIER>
    __device__ __forceinline__ void add(double x)
    {
        double t = __dadd_rn(s, x);
        if (NEUMAIER) {
            const bool big = fabs(s) >= fabs(x);
            const double hi = big ? s : x, lo = big ? x : s;
            c = __dadd_rn(c, __dadd_rn(__dadd_rn(hi, -t), lo));
        }
        s = t;
    }
    // Eight consecutive items at once.  Exactly the same operations in the same order as eight add()
    // calls, but laid out so the three dependency chains are visible to the scheduler: the running
    // sums t[u] (8 dependent DADDs, 8.2 cycles each on sm_100), the per-item error terms (independent
    // of each other) and the compensation chain, which overlaps the next block's running sums.
    // Measured: ~11 cycles/item instead of 34 for the item-at-a-time loop.
    template <bool NEUMAIER>
    __device__ __forceinline__ void add_block(const double (&x)[8])
    {
        double t[8];
        t[0] = __dadd_rn(s, x[0]);
#pragma unroll
        for (int u = 1; u < 8; ++u) t[u] = __dadd_rn(t[u - 1], x[u]);
        if (NEUMAIER) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const double prev = u ? t[u - 1] : s;
                const bool big = fabs(prev) >= fabs(x[u]);
                const double hi = big ? prev : x[u], lo = big ? x[u] : prev;
                c = __dadd_rn(c, __dadd_rn(__dadd_rn(hi, -t[u]), lo));
            }
        }
        s = t[7];
    }
    template <bool NEUMAIER>
    __device__ __forceinline__ double result() const
    {
        double r = s;
        if (NEUMAIER && c != 0.0 && isfinite(c)) r = __dadd_rn(r, c);
        return r;
    }
};

// acc += sum over i in [lo, hi) of prod(i), in index order (blocks of 8 + tail)
template <bool NEUMAIER, typename F>
__device__ __forceinline__ void ref_sum_range(RefSum &acc, int lo, int hi, F prod)
{
    int i = lo;
    for (; i + 8 <= hi; i += 8) {
        double x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = prod(i + u);
        acc.add_block<NEUMAIER>(x);
    }
    for (; i < hi; ++i) acc.add<NEUMAIER>(prod(i));
}

__device__ __forceinline__ float load_as_float(const float *p, int64_t i) { return p[i]; }
__device__ __forceinline__ float load_as_float(const __nv_bfloat16 *p, int64_t i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ double load_as_double(const void *p, int dtype, int64_t i)
{
    if (dtype == VM_F32) return (double)((const float *)p)[i];
    if (dtype == VM_BF16) return (double)__bfloat162float(((const __nv_bfloat16 *)p)[i]);
    return ((const double *)p)[i];
}

static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int ld_for_dim(int dim) { return (dim + 7) & ~7; }
static inline size_t dtype_size(int dt) { return dt == VM_F32 ? 4 : dt == VM_BF16 ? 2 : 8; }

// ---- kernel launch interfaces (defined in the .cu files) -------------------------------
struct ScanArgs {
    const void *rows;        // [n][ld] store dtype
    const float *inv_norms;  // [n] 1/||row||; 0 = zero-norm row (scores 0.0), < 0 = skipped row (never returned)
    int dtype;               // store dtype
    int64_t n;
    int dim, ld;
    const float *queries;    // [nq_pad][ld] fp32, L2-normalised (zero query stays zero), device
    int nq;                  // queries in this launch
    int kp;                  // candidate list length per query (<= 64)
    uint64_t *cand;          // [ctas][nq][kp] keys out
    int ctas;                // number of CTAs to launch (lists produced)
    cudaStream_t stream;
    int dump = 0;            // tcgen05 scan only: small store -- emit EVERY row's key, cand = [tiles][nq][128] (see scan_tc.cu)
};
static constexpr int SCAN_DUMP_TILE = 128;     // rows (= keys per query) per dumped tile
static constexpr int SCAN_DUMP_MAX_KEYS = 9472;  // per query: what the fused select kernel ranks in shared memory (148 x 64)
int launch_scan_simt(const ScanArgs &a);
int scan_simt_max_queries();
// second pass of the tcgen05 scan for uncertified queries (see scan_tc.cu "collect mode")
struct ScanCollect {
    const float *thr;    // [nq] thresholds written by the rescoring kernel (+inf = query is done)
    uint64_t *buf;       // [nq][cap]
    int *cnt;            // [nq], zeroed before the launch
    int cap;
    const int *pending;  // uncertified-query counter
};
// tcgen05 path; seed_tab [ctas][nq_pad] u32 + seed_ctr must be zeroed before the launch (NULL = no seeding);
// sc != NULL runs the collect pass instead of the top-kp pass
struct ScanInfo { int stages = 0, variant = 0; };  // what the launch chose (reported in vm_topk_stats): pipeline depth; 0 lists, 1 dump, 2 lists + threshold warp
int launch_scan_tc(const ScanArgs &a, const void *queries_store_dtype, uint32_t *seed_tab, int *seed_ctr,
                   const ScanCollect *sc = nullptr, ScanInfo *info = nullptr);
bool scan_tc_supported(int dtype, int dim, int nq, int kp);


// ---- argument blocks shared by api.cu and select.cu ------------------------------------
struct FinalizeArgs {
    int k;
    double min_score;
    int score_mode;
    int64_t row_offset;  // added to local rows in the output (sharded stores)
    int64_t *out_idx;    // [nq][k]
    double *out_score;   // [nq][k]
    int32_t *out_count;  // [nq]
};
struct RescoreArgs {
    const uint64_t *merged;
    int kp;
    const void *rows;
    const float *inv_norms;
    int dtype, ld, dim;
    int64_t n_rows;
    const void *queries;
    int q_dtype, nq;
    double eps;
    int sum_mode;
    FinalizeArgs fin;
    int32_t *flags, *uncertified_count;
    const int *extreme;  // store-level count of rows outside the scans' numeric range (forces the exact pass)
    float *collect_thr;  // [nq] out: threshold of the collect pass for uncertified queries (+inf otherwise); may be NULL
    unsigned long long *cum = nullptr;  // store-lifetime counters (vm_store_read_counters): [0] uncertified, [1] settled from the
                                        // band, [2] settled by the collect pass, [3] redone by the binary64 scan of every row
};
struct ExactArgs {
    const void *rows;
    const float *inv_norms;
    int dtype, ld, dim;
    int64_t n;
    const void *queries;
    int q_dtype, nq, k, sum_mode;
    const int32_t *flags;  // NULL = all queries
    double *xlist_score;   // [ctas][nq][k]
    uint32_t *xlist_row;
    int32_t *xlist_cnt;    // [ctas][nq]
    uint8_t *taken;        // [nq][ctas*k]
    int ctas;
    FinalizeArgs fin;
    unsigned long long *cum = nullptr;  // see RescoreArgs::cum
};

}  // namespace vm
