// hll_count.cu -- K3: stand-alone HyperLogLog++ cardinality, and the K5 row helpers.
//
// Replaces ElphHashes.hll_count (/root/reference/src/hashing.py:212-232) with _linearcounting (:194-195),
// _estimate_bias (:197-204), _refine_hll_count_estimate (:206-210); jaccard (:247-256); _hll_merge (:234-237).
// The reference materialises a [n, m] float32 tensor of 2^-reg and an [n, T] argsort; here one warp reads a
// register row once, sums 2^-reg exactly in fixed point, and finishes with a scalar tail.
#include "common.cuh"

namespace ss {

__global__ void __launch_bounds__(256) hll_count_kernel(const uint8_t *__restrict__ regs, int64_t row_stride, int64_t n,
                                                         HllDev h, float *__restrict__ out, int64_t out_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int units = h.m >> 3;
    for (int64_t i = gwarp; i < n; i += n_warps) {
        const uint8_t *row = regs + i * row_stride;
        RegSum s;
        s.lo = s.hi = 0;
        s.zeros = 0;
        for (int u = lane; u < units; u += 32) {
            const uint2 v = ld_nc_u2(row + (int64_t)u * 8);
            regsum_add_word(s, v.x);
            regsum_add_word(s, v.y);
        }
        int zeros;
        unsigned __int128 t = regsum_warp_total(s, zeros);
        float c = hll_estimate(h, zeros, t);
        if (lane == 0) out[i * out_stride] = c;
    }
}

__global__ void __launch_bounds__(256) estimate_bias_kernel(const float *__restrict__ e, int64_t n, HllDev h,
                                                             float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = bias_6nn(h, e[i]);
}

__global__ void __launch_bounds__(256) jaccard_i64_kernel(const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                           int64_t n, int64_t width, int64_t denom, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = gwarp; i < n; i += n_warps) {
        uint32_t eq = 0;
        for (int64_t c = lane; c < width; c += 32) eq += (src[i * width + c] == dst[i * width + c]) ? 1u : 0u;
        eq = __reduce_add_sync(FULL, eq);
        if (lane == 0) out[i] = __fdiv_rn((float)eq, (float)denom);
    }
}

__global__ void __launch_bounds__(256) max_i8_kernel(const int8_t *__restrict__ a, const int8_t *__restrict__ b,
                                                      int64_t count, int8_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a[i] > b[i] ? a[i] : b[i];
}

static int grid_for(int64_t threads, int block, int per_sm) {
    int64_t blocks = (threads + block - 1) / block;
    int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace ss

extern "C" {

int ss_hll_count(const void *regs, int64_t row_stride, int64_t n, const ss_hll_consts *hc, float *out,
                 int64_t out_stride, ss_stream_t stream) {
    SS_REQUIRE(n >= 0, "n must be >= 0");
    SS_REQUIRE(hc != nullptr, "hll constants are required");
    int rc = ss::check_hll_consts(hc, hc->p);
    if (rc != SS_OK) return rc;
    SS_REQUIRE(hc->p >= 4 && hc->p <= 18, "unsupported hll_p=%d", hc->p);
    if (n == 0) return SS_OK;
    SS_REQUIRE(regs && out, "null pointer passed to ss_hll_count");
    SS_REQUIRE(((uintptr_t)regs & 7) == 0 && (row_stride & 7) == 0 && row_stride >= (1 << hc->p),
               "register rows must be 8-byte aligned and at least m bytes apart");
    ss::hll_count_kernel<<<ss::grid_for(n * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        (const uint8_t *)regs, row_stride, n, ss::to_dev(hc), out, out_stride);
    SS_LAUNCH_CHECK("hll_count_kernel");
    return SS_OK;
}

int ss_estimate_bias(const float *e, int64_t n, const ss_hll_consts *hc, float *out, ss_stream_t stream) {
    SS_REQUIRE(n >= 0, "n must be >= 0");
    SS_REQUIRE(hc != nullptr, "hll constants are required");
    int rc = ss::check_hll_consts(hc, hc->p);
    if (rc != SS_OK) return rc;
    if (n == 0) return SS_OK;
    SS_REQUIRE(e && out, "null pointer passed to ss_estimate_bias");
    ss::estimate_bias_kernel<<<ss::grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(e, n, ss::to_dev(hc), out);
    SS_LAUNCH_CHECK("estimate_bias_kernel");
    return SS_OK;
}

int ss_jaccard_i64(const int64_t *src, const int64_t *dst, int64_t n, int64_t width, int64_t denom, float *out,
                   ss_stream_t stream) {
    SS_REQUIRE(n >= 0 && width >= 1 && denom >= 1, "bad size passed to ss_jaccard_i64");
    if (n == 0) return SS_OK;
    SS_REQUIRE(src && dst && out, "null pointer passed to ss_jaccard_i64");
    ss::jaccard_i64_kernel<<<ss::grid_for(n * 32, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, dst, n, width, denom, out);
    SS_LAUNCH_CHECK("jaccard_i64_kernel");
    return SS_OK;
}

int ss_max_i8(const int8_t *a, const int8_t *b, int64_t count, int8_t *out, ss_stream_t stream) {
    SS_REQUIRE(count >= 0, "count must be >= 0");
    if (count == 0) return SS_OK;
    SS_REQUIRE(a && b && out, "null pointer passed to ss_max_i8");
    ss::max_i8_kernel<<<ss::grid_for(count, 256, 8), 256, 0, (cudaStream_t)stream>>>(a, b, count, out);
    SS_LAUNCH_CHECK("max_i8_kernel");
    return SS_OK;
}

}  // extern "C"
