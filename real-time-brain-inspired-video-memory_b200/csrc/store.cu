// store.cu -- store-side kernels: dtype conversion on append, cached inverse norms,
// query normalisation, synthetic fill.  None of these is on the per-query critical path.
#include "common.cuh"
#include "synth.cuh"

namespace vm {

// src [n][dim] (f32/bf16/f64) -> dst [n][ld] (f32/bf16, or f64: the exact copy of a binary64 store), zero padded
// columns [dim, ld).
template <typename DST>
__global__ void convert_rows_kernel(const void *__restrict__ src, int src_dtype, DST *__restrict__ dst, int64_t n,
                                    int dim, int ld)
{
    int64_t total = n * (int64_t)ld;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e / ld;
        int c = (int)(e - r * ld);
        double dv = 0.0;
        if (c < dim) dv = load_as_double(src, src_dtype, r * dim + c);
        if constexpr (sizeof(DST) == 8) dst[e] = dv;  // widening is exact
        else {
            const float v = (float)dv;                // binary64 -> binary32, round to nearest
            if constexpr (sizeof(DST) == 4) dst[e] = v;
            else dst[e] = __float2bfloat16_rn(v);
        }
    }
}

// One warp per row: 1/||row|| of the STORED values, accumulated in binary64.
// 0 for a zero-norm row (its score is 0.0: pre_llm_injector.py:385-386).
// `extreme` (may be NULL) counts rows whose squared norm lies outside [2^-120, 2^120]: the fast scans'
// error bound assumes normal-range arithmetic (tensor cores flush fp32 denormals, huge rows overflow the
// fp32 accumulator), so any such row makes the exact pass re-do every query of this store.
// Rows with a non-finite norm (NaN / Inf elements) are marked skipped.
// exact (binary64 store only, else NULL): the binary64 originals of the rows.  The scan works on `rows`, their rounded
// shadow: a row whose shadow lost it (all elements flushed to zero, or pushed to infinity, by the rounding) while the
// original is an ordinary vector is counted as extreme as well, and keeps inverse norm 0 instead of "skipped" -- the
// binary64 pass then scores it from the original.
template <typename T>
__global__ void row_inv_norms_kernel(const T *__restrict__ rows, float *__restrict__ inv_norms, int64_t row0,
                                     int64_t n, int ld, int *__restrict__ extreme, const double *__restrict__ exact)
{
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = row0 + warp; r < n; r += nwarps) {
        const T *p = rows + r * (int64_t)ld;
        double ss = 0.0;
        for (int c = lane; c < ld; c += 32) {
            double v = (double)load_as_float(p, c);
            ss += v * v;
        }
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        double se = 0.0;   // exact rows: scaled squared norm of the original (scaled so that it cannot overflow)
        if (exact != nullptr) {
            for (int c = lane; c < ld; c += 32) {
                const double v = exact[r * (int64_t)ld + c] * 0x1p-600;
                se += v * v;
            }
            for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        }
        if (lane == 0) {
            float inv = 0.0f;
            if (!isfinite(ss)) {
                inv = -1.0f;
                if (exact != nullptr && isfinite(se)) { inv = 0.0f; if (extreme) atomicAdd(extreme, 1); }  // finite original
            } else if (ss > 0.0) {
                inv = (float)(1.0 / sqrt(ss));
                if (extreme && (ss < 7.52316384526264e-37 || ss > 1.329227995784916e+36)) atomicAdd(extreme, 1);
            } else if (exact != nullptr && se != 0.0) {
                if (extreme) atomicAdd(extreme, 1);                                                     // shadow flushed to zero
            }
            inv_norms[r] = inv;
        }
    }
}

__global__ void invalidate_rows_kernel(float *inv_norms, const int64_t *rows, int64_t n, int64_t size)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && rows[i] >= 0 && rows[i] < size) inv_norms[rows[i]] = -1.0f;
}

// Queries (any dtype) -> fp32 [nq_pad][ld], each scaled by 1/||q|| (binary64 norm, zero query
// stays zero), plus an optional bf16 copy for the tcgen05 bf16 path.  One warp per query.
//   tf32_round: the fp32 copy is rounded to nearest tf32 (10 mantissa bits) -- the tensor core would otherwise
//               truncate it, and rounding halves the query's share of the scan's error bound;
//   split:      the bf16 copy is [2][nq_pad][ld]: hi = bf16(v), lo = bf16(v - hi) -- 16 mantissa bits in two terms.
__global__ void normalize_queries_kernel(const void *__restrict__ q, int q_dtype, int nq, int nq_pad, int dim, int ld,
                                         float *__restrict__ out_f32, __nv_bfloat16 *__restrict__ out_bf16,
                                         int32_t *__restrict__ zero_me, uint32_t *__restrict__ zero_tab, int zero_tab_n,
                                         int tf32_round, int split)
{
    pdl_launch_dependents();
    pdl_wait();  // first kernel of the chain: the previous call's kernels may still be reading what is zeroed below
    // per-batch device state: zero_me[0] = uncertified-query counter, zero_me[1] = seed counter,
    // zero_tab = the threshold-seeding table, published bounds and union-buffer counters of the tcgen05 scan
    if (zero_me && blockIdx.x == 0 && threadIdx.x < 2) zero_me[threadIdx.x] = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zero_tab_n; i += gridDim.x * blockDim.x) zero_tab[i] = 0u;
    int lane = threadIdx.x & 31;
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nq_pad) return;
    double ss = 0.0;
    if (w < nq)
        for (int c = lane; c < dim; c += 32) {
            double v = load_as_double(q, q_dtype, (int64_t)w * dim + c);
            ss += v * v;
        }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    double inv = (ss > 0.0 && isfinite(ss)) ? 1.0 / sqrt(ss) : 0.0;
    for (int c = lane; c < ld; c += 32) {
        float v = 0.0f;
        if (w < nq && c < dim) v = (float)(load_as_double(q, q_dtype, (int64_t)w * dim + c) * inv);
        float vf = v;
        if (tf32_round) vf = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
        out_f32[(int64_t)w * ld + c] = vf;
        if (out_bf16) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            out_bf16[(int64_t)w * ld + c] = hi;
            if (split) out_bf16[((int64_t)nq_pad + w) * ld + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
    }
}

template <typename T>
__global__ void synth_fill_kernel(T *__restrict__ rows, uint64_t seed, int64_t row0, int64_t n, int dim, int ld,
                                  uint64_t dup_period)
{
    // one thread per 8-column group: one splitmix64 word yields 8 values
    int groups = ld >> 3;
    int64_t total = n * (int64_t)groups;
    uint64_t hk = splitmix64(seed ^ 0xD6E8FEB86659FD93ULL);
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = e / groups;
        int g = (int)(e - i * groups);
        uint64_t row = (uint64_t)(row0 + i);
        uint64_t own = splitmix64(splitmix64(seed * 0x9E3779B97F4A7C15ULL + row) + (uint64_t)g);
        uint64_t par = 0, sel = 0;
        bool planted = false;
        if (dup_period > 0 && row > 0) {
            uint64_t hr = splitmix64(hk + row);
            if (hr % dup_period == 0) {
                planted = true;
                uint64_t parent = splitmix64(hr) % row;
                par = splitmix64(splitmix64(seed * 0x9E3779B97F4A7C15ULL + parent) + (uint64_t)g);
                sel = splitmix64(hr + 0x632BE59BD9B4E019ULL + (uint64_t)g);
            }
        }
        T *dst = rows + i * (int64_t)ld + g * 8;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            int col = g * 8 + b;
            int v = synth_int(own, b);
            if (planted && (((sel >> (8 * b)) & 15ULL) != 0ULL)) v = synth_int(par, b);
            float f = col < dim ? (float)v * (1.0f / 128.0f) : 0.0f;
            if constexpr (sizeof(T) == 4) dst[b] = f;
            else dst[b] = __float2bfloat16_rn(f);
        }
    }
}

// ---- host wrappers ---------------------------------------------------------------------
int k_convert_rows(const void *src, int src_dtype, void *dst, int dst_dtype, int64_t n, int dim, int ld, cudaStream_t st)
{
    if (n <= 0) return VM_OK;
    int64_t total = n * (int64_t)ld;
    int grid = (int)vm::imin64((total + 255) / 256, 148 * 16);
    if (dst_dtype == VM_F32) convert_rows_kernel<float><<<grid, 256, 0, st>>>(src, src_dtype, (float *)dst, n, dim, ld);
    else if (dst_dtype == VM_F64) convert_rows_kernel<double><<<grid, 256, 0, st>>>(src, src_dtype, (double *)dst, n, dim, ld);
    else convert_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(src, src_dtype, (__nv_bfloat16 *)dst, n, dim, ld);
    VM_CUDA_CHECK(cudaGetLastError());
    return VM_OK;
}

int k_row_inv_norms(const void *rows, int dtype, float *inv_norms, int64_t row0, int64_t n, int ld, int *extreme, cudaStream_t st,
                    const double *exact)
{
    if (n <= row0) return VM_OK;
    int64_t warps = n - row0;
    int grid = (int)vm::imin64((warps * 32 + 255) / 256, 148 * 16);
    if (dtype == VM_F32) row_inv_norms_kernel<float><<<grid, 256, 0, st>>>((const float *)rows, inv_norms, row0, n, ld, extreme, exact);
    else row_inv_norms_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)rows, inv_norms, row0, n, ld, extreme, exact);
    VM_CUDA_CHECK(cudaGetLastError());
    return VM_OK;
}

int k_invalidate_rows(float *inv_norms, const int64_t *rows_dev, int64_t n, int64_t size, cudaStream_t st)
{
    if (n <= 0) return VM_OK;
    invalidate_rows_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(inv_norms, rows_dev, n, size);
    VM_CUDA_CHECK(cudaGetLastError());
    return VM_OK;
}

int k_normalize_queries(const void *q, int q_dtype, int nq, int nq_pad, int dim, int ld, float *out_f32,
                        void *out_bf16, int32_t *zero_me, uint32_t *zero_tab, int zero_tab_n, int tf32_round, int split, bool pdl,
                        cudaStream_t st)
{
    int threads = 128;
    int grid = (nq_pad * 32 + threads - 1) / threads;
    VM_CUDA_CHECK(launch_pdl(normalize_queries_kernel, dim3(grid), dim3(threads), 0, st, pdl, q, q_dtype, nq, nq_pad, dim, ld, out_f32,
                             (__nv_bfloat16 *)out_bf16, zero_me, zero_tab, zero_tab_n, tf32_round, split));
    VM_CUDA_CHECK(cudaGetLastError());
    return VM_OK;
}

int k_synth_fill(void *rows, int dtype, uint64_t seed, int64_t row0, int64_t n, int dim, int ld, uint64_t dup_period,
                 cudaStream_t st)
{
    if (n <= 0) return VM_OK;
    int64_t total = n * (int64_t)(ld >> 3);
    int grid = (int)vm::imin64((total + 255) / 256, 148 * 32);
    if (dtype == VM_F32) synth_fill_kernel<float><<<grid, 256, 0, st>>>((float *)rows, seed, row0, n, dim, ld, dup_period);
    else synth_fill_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16 *)rows, seed, row0, n, dim, ld, dup_period);
    VM_CUDA_CHECK(cudaGetLastError());
    return VM_OK;
}

}  // namespace vm
