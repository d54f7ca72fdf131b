// kb_scan_util.cu -- generic exclusive prefix sums (three small launches: block sums, scan of the sums by
// one CTA, apply).  Used for group/watcher/victim offsets; the record-level scans of the hot path live
// in kb_scan.cu.
#include "kb_internal.cuh"

namespace {

constexpr int SB = 256;      // threads
constexpr int SI = 4;        // items per thread
constexpr int SE = SB * SI;  // items per block

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v)
{
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

// exclusive scan of one value per thread over a block of SB threads; returns exclusive prefix, sets total
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T *warp_sums /* SB/32 + 1 */, T &total)
{
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = warp_incl_scan(v);
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        T s = (lane < SB / 32) ? warp_sums[lane] : (T)0;
        T si = warp_incl_scan(s);
        if (lane < SB / 32) warp_sums[lane] = si - s;
        if (lane == SB / 32 - 1) warp_sums[SB / 32] = si;
    }
    __syncthreads();
    T res = inc - v + warp_sums[w];
    total = warp_sums[SB / 32];
    __syncthreads();
    return res;
}

template <typename T>
__global__ void __launch_bounds__(SB) k_scan_block_sums(const T *__restrict__ in, T *__restrict__ sums, uint32_t n)
{
    __shared__ T ws[SB / 32 + 1];
    uint32_t base = blockIdx.x * SE + threadIdx.x * SI;
    T s = 0;
#pragma unroll
    for (int k = 0; k < SI; k++)
        if (base + k < n) s += in[base + k];
    T total;
    block_excl_scan(s, ws, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(SB) k_scan_sums(T *__restrict__ sums, uint32_t nb, T *__restrict__ total_out)
{
    __shared__ T ws[SB / 32 + 1];
    __shared__ T carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t c0 = 0; c0 < nb; c0 += SE) {
        uint32_t base = c0 + threadIdx.x * SI;
        T v[SI];
        T s = 0;
#pragma unroll
        for (int k = 0; k < SI; k++) {
            v[k] = (base + k < nb) ? sums[base + k] : (T)0;
            s += v[k];
        }
        T total;
        T ex = block_excl_scan(s, ws, total);
        T carry = carry_s;
        T run = carry + ex;
#pragma unroll
        for (int k = 0; k < SI; k++) {
            if (base + k < nb) sums[base + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

template <typename T>
__global__ void __launch_bounds__(SB) k_scan_apply(const T *__restrict__ in, T *__restrict__ out,
                                                   const T *__restrict__ sums, uint32_t n)
{
    __shared__ T ws[SB / 32 + 1];
    uint32_t base = blockIdx.x * SE + threadIdx.x * SI;
    T v[SI];
    T s = 0;
#pragma unroll
    for (int k = 0; k < SI; k++) {
        v[k] = (base + k < n) ? in[base + k] : (T)0;
        s += v[k];
    }
    T total;
    T ex = block_excl_scan(s, ws, total);
    T run = sums[blockIdx.x] + ex;
#pragma unroll
    for (int k = 0; k < SI; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

template <typename T>
int scan_impl(kb_ctx *ctx, const T *in, T *out, uint32_t n, T *total_dev)
{
    if (n == 0) {
        if (total_dev) KB_CUDA(ctx, cudaMemsetAsync(total_dev, 0, sizeof(T), ctx->stream));
        return KB_OK;
    }
    uint32_t nb = (n + SE - 1) / SE;
    KB_TRY(dbuf_ensure(ctx, ctx->d_scan_tmp, (size_t)nb * sizeof(T)));
    T *sums = (T *)ctx->d_scan_tmp.p;
    KB_LAUNCH(ctx, "util_scan", 0, (k_scan_block_sums<T><<<nb, SB, 0, ctx->stream>>>(in, sums, n)));
    KB_LAUNCH(ctx, "util_scan", 0, (k_scan_sums<T><<<1, SB, 0, ctx->stream>>>(sums, nb, total_dev)));
    KB_LAUNCH(ctx, "util_scan", 0, (k_scan_apply<T><<<nb, SB, 0, ctx->stream>>>(in, out, sums, n)));
    KB_CUDA(ctx, cudaGetLastError());
    return KB_OK;
}

}  // namespace

int scan_exclusive_u32(kb_ctx *ctx, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *total_dev)
{
    return scan_impl<uint32_t>(ctx, in, out, n, total_dev);
}

int scan_exclusive_u64(kb_ctx *ctx, const uint64_t *in, uint64_t *out, uint32_t n, uint64_t *total_dev)
{
    return scan_impl<uint64_t>(ctx, in, out, n, total_dev);
}
