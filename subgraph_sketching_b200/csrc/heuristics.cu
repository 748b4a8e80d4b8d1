// heuristics.cu -- common-neighbour link heuristics on a sorted CSR adjacency.
//
// Replaces CN / AA / RA of /root/reference/src/heuristics.py:11-71 (scipy: per batch A[src].multiply(A_[dst]) and
// a row sum, in 100k-2M link batches on the CPU; HashDataset calls RA when --use_RA, datasets/elph.py:76-77):
//     score(u, v) = sum over common neighbours w of  A[u,w] * (A[v,w] * mult[w])
// with mult = 1 (CN), 1 / log(colsum) (AA), 1 / colsum (RA), everything in float64 like scipy, result cast to
// float32.  One warp per link: the lanes walk the shorter adjacency row and binary-search the longer one.
#include "common.cuh"

namespace ss {

__global__ void __launch_bounds__(256) col_sums_kernel(const int32_t *__restrict__ colidx, const double *__restrict__ w,
                                                        int64_t nnz, double *__restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(out + colidx[e], w ? w[e] : 1.0);
}

__global__ void __launch_bounds__(256) cn_scores_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                         const double *__restrict__ w, const double *__restrict__ mult,
                                                         int64_t n_nodes, const int64_t *__restrict__ links, int64_t n_links,
                                                         float *__restrict__ out, int *err) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = gwarp; i < n_links; i += n_warps) {
        int64_t u = __ldg(links + 2 * i), v = __ldg(links + 2 * i + 1);
        if ((uint64_t)u >= (uint64_t)n_nodes || (uint64_t)v >= (uint64_t)n_nodes) {
            if (lane == 0) {
                if (err) atomicExch(err, 1);
                out[i] = 0.f;
            }
            continue;
        }
        int64_t ub = __ldg(rowptr + u), ue = __ldg(rowptr + u + 1);
        int64_t vb = __ldg(rowptr + v), ve = __ldg(rowptr + v + 1);
        // walk the shorter row (a), search the longer one (b); remember which one is u for the product order
        const bool swap = (ue - ub) > (ve - vb);
        const int64_t ab = swap ? vb : ub, ae = swap ? ve : ue, bb = swap ? ub : vb, be = swap ? ue : ve;
        double acc = 0.0;
        for (int64_t p = ab + lane; p < ae; p += 32) {
            const int32_t c = __ldg(colidx + p);
            int64_t lo = bb, hi = be;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(colidx + mid) < c) lo = mid + 1; else hi = mid;
            }
            if (lo < be && __ldg(colidx + lo) == c) {
                const double wa = w ? w[p] : 1.0, wb = w ? w[lo] : 1.0;
                const double wu = swap ? wb : wa, wv = swap ? wa : wb;
                acc += wu * (wv * __ldg(mult + c));  // A[u,w] * (A[v,w] * mult[w]), as A[src].multiply(A_[dst])
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == 0) out[i] = (float)acc;
    }
}

}  // namespace ss

extern "C" {

int ss_col_sums(const int32_t *colidx, const double *weights, int64_t nnz, int64_t n_cols, double *out, ss_stream_t stream) {
    SS_REQUIRE(nnz >= 0 && n_cols >= 0, "negative size passed to ss_col_sums");
    SS_REQUIRE(n_cols == 0 || out, "out is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_cols > 0) SS_CUDA(cudaMemsetAsync(out, 0, (size_t)n_cols * 8, st));
    if (nnz == 0) return SS_OK;
    SS_REQUIRE(colidx, "colidx is null");
    int64_t blocks = (nnz + 255) / 256, cap = (int64_t)ss::sm_count() * 16;
    ss::col_sums_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(colidx, weights, nnz, out);
    SS_LAUNCH_CHECK("col_sums_kernel");
    return SS_OK;
}

int ss_common_neighbour_scores(const int64_t *rowptr, const int32_t *colidx, const double *weights, const double *mult,
                               int64_t n_nodes, const int64_t *links, int64_t n_links, float *out, int32_t *error_flag,
                               ss_stream_t stream) {
    SS_REQUIRE(n_nodes >= 0 && n_links >= 0, "negative size passed to ss_common_neighbour_scores");
    if (n_links == 0) return SS_OK;
    SS_REQUIRE(rowptr && mult && links && out, "null pointer passed to ss_common_neighbour_scores");
    int64_t blocks = (n_links + 7) / 8, cap = (int64_t)ss::sm_count() * 8;
    ss::cn_scores_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        rowptr, colidx, weights, mult, n_nodes, links, n_links, out, error_flag);
    SS_LAUNCH_CHECK("cn_scores_kernel");
    return SS_OK;
}

}  // extern "C"
