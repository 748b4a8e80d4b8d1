// sign.cu -- next row: SIGN node-feature pre-propagation.
//
// Replaces HashDataset._generate_sign_features (/root/reference/src/datasets/elph.py:87-110):
//     edge_index, w = gcn_norm(edge_index, edge_weight.float(), num_nodes)        (PyG, third party)
//     x' = torch_sparse.spmm(edge_index, w, N, N, x)                              (third party)
// gcn_norm = add_remaining_self_loops (fill 1, existing self-loop weights kept) ; deg = scatter-sum of the
// weights at edge_index[1] ; w = deg^-1/2[row] * w * deg^-1/2[col] with inf -> 0.   spmm: out[row] += w * x[col].
//
// The normalised edge list is never materialised: the kernels keep deg^-1/2 per node and the self-loop weight
// per node, and the SpMM forms each coefficient on the fly in the reference's left-to-right float32 order.
// The adjacency is a CSR keyed by edge_index[0] (the spmm row) whose entries are EDGE POSITIONS, so arbitrary
// edge weights ride along without a permuted copy.  Float32, no FMA contraction (-fmad=false): every product
// and sum is rounded like the reference's.  When the edge list is sorted by row (what coalesce / to_undirected
// produce) the CSR IS the edge list: no fill, and every row is summed in edge order exactly like the
// reference's sequential scatter-add -- bit-identical results.  Otherwise the CSR is filled with atomics, the
// ORDER of the per-row sum is unspecified and results agree to float32 summation-order tolerance.
#include <string.h>

#include "common.cuh"

namespace ss {

static int64_t sg_align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// workspace: int32 loop_eid[N] | uint32 cursor[N]
static int64_t sign_ws_bytes(int64_t n) { return 2 * sg_align_up(n * 4, 256); }

static int sg_grid(int64_t items) {
    int64_t blocks = (items + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 32;
    if (blocks < 1) blocks = 1;
    return (int)(blocks < cap ? blocks : cap);
}

// position of the LAST self-loop edge of every node (add_remaining_self_loops keeps that weight: the index_put
// over duplicate indices is sequential on the CPU); the same pass validates the ids and notes whether the list
// is sorted by row.  flags: bit 0 = row[] is not non-decreasing, bit 1 = some id is outside [0, n_nodes)
__global__ void __launch_bounds__(256) sign_loops_kernel(const int64_t *__restrict__ row, const int64_t *__restrict__ col,
                                                          int64_t n_edges, int64_t n_nodes, int32_t *loop_eid, int32_t *flags) {
    int f = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row[e], c = col[e];
        if ((uint64_t)r >= (uint64_t)n_nodes || (uint64_t)c >= (uint64_t)n_nodes) { f |= 2; continue; }
        if (e > 0 && row[e - 1] > r) f |= 1;
        if (r == c) atomicMax(loop_eid + r, (int32_t)e);
    }
    if (f && flags) atomicOr(flags, f);
}

// deg starts at the self-loop weight of the node (its own, or the fill value 1)
__global__ void __launch_bounds__(256) sign_loop_weight_kernel(const int32_t *__restrict__ loop_eid, const float *__restrict__ ew,
                                                                int64_t n_nodes, float *__restrict__ loop_w,
                                                                float *__restrict__ deg) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t e = loop_eid[i];
        const float w = (e >= 0 && ew) ? ew[e] : 1.0f;
        loop_w[i] = w;
        deg[i] = w;
    }
}

// deg[col] += w over the non-self-loop edges (float32 atomics: exact, hence order independent, for the
// integer-valued weights the reference's datasets carry; summation-order tolerance otherwise)
__global__ void __launch_bounds__(256) sign_degree_kernel(const int64_t *__restrict__ row, const int64_t *__restrict__ col,
                                                           const float *__restrict__ ew, int64_t n_edges, int64_t n_nodes,
                                                           float *deg) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row[e], c = col[e];
        if (r == c || (uint64_t)c >= (uint64_t)n_nodes) continue;
        atomicAdd(deg + c, ew ? ew[e] : 1.0f);
    }
}

// deg -> deg^-1/2 in place (torch: pow(-0.5) = 1 / sqrt, both correctly rounded), inf -> 0
__global__ void __launch_bounds__(256) sign_dinv_kernel(float *__restrict__ deg, int64_t n_nodes) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = __fdiv_rn(1.0f, __fsqrt_rn(deg[i]));
        deg[i] = isinf(d) ? 0.0f : d;
    }
}

__global__ void __launch_bounds__(256) sign_fill_kernel(const int64_t *__restrict__ row, int64_t n_edges, int64_t n_nodes,
                                                         const int64_t *__restrict__ rowptr, uint32_t *cursor,
                                                         int32_t *__restrict__ perm) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row[e];
        if ((uint64_t)r >= (uint64_t)n_nodes) continue;
        const uint32_t k = atomicAdd(cursor + r, 1u);
        perm[rowptr[r] + k] = (int32_t)e;
    }
}

struct SpmmArgs {
    const int64_t *rowptr;
    const int32_t *perm;
    const int64_t *col;
    const float *ew;
    const float *dinv;
    const float *loop_w;
    const float *x;
    int64_t x_stride;   // floats between consecutive rows of x
    int64_t n_nodes;
    int64_t F;
    float *out;
    int64_t out_stride; // floats between consecutive rows of out
    int copies;         // the result row is written `copies` times, F floats apart
};

// One warp per output row; the feature axis is processed in tiles of 128 columns (lane l owns columns
// 4l..4l+3 of the tile as one float4 when VEC4, else columns l, l+32, l+64, l+96).  The 32 lanes fetch the
// metadata of 32 entries at once (edge position -> source node, coefficient) and broadcast them with shuffles;
// four source rows are in flight per lane.
template <bool VEC4>
__global__ void __launch_bounds__(256) sign_spmm_kernel(const SpmmArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = gwarp; i < a.n_nodes; i += n_warps) {
        const int64_t b = a.rowptr[i], e = a.rowptr[i + 1];
        const float di = a.dinv[i];
        // self loop appended last by add_remaining_self_loops: (dinv[i] * w_loop) * dinv[i]
        const float self_coef = __fmul_rn(__fmul_rn(di, a.loop_w[i]), di);
        for (int64_t c0 = 0; c0 < a.F; c0 += 128) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            int64_t cols[4];
            bool live[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                cols[q] = VEC4 ? c0 + 4 * lane + q : c0 + lane + 32 * q;
                live[q] = cols[q] < a.F;
            }
            for (int64_t base = b; base < e; base += 32) {
                const int64_t j = base + lane;
                int64_t src = 0;
                float coef = 0.f;
                bool use = false;
                if (j < e) {
                    const int64_t eid = a.perm ? (int64_t)__ldg(a.perm + j) : j;  // no perm: the edge list is sorted by row
                    src = __ldg(a.col + eid);
                    use = (src != i) && (uint64_t)src < (uint64_t)a.n_nodes;  // self-loop edges are replaced by the loop term
                    if (use) coef = __fmul_rn(__fmul_rn(di, a.ew ? __ldg(a.ew + eid) : 1.0f), __ldg(a.dinv + src));
                }
                const unsigned mask = __ballot_sync(FULL, use);
                const int cnt = (int)min((int64_t)32, e - base);
                for (int t0 = 0; t0 < cnt; t0 += 4) {
                    float4 v[4];
                    float cf[4];
                    bool on[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = t0 + u;
                        on[u] = t < cnt && ((mask >> t) & 1u);
                        const int64_t s = __shfl_sync(FULL, src, t & 31);
                        cf[u] = __shfl_sync(FULL, coef, t & 31);
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (on[u]) {
                            const float *xr = a.x + s * a.x_stride;
                            if (VEC4) {
                                if (live[0]) v[u] = __ldg(reinterpret_cast<const float4 *>(xr + cols[0]));
                            } else {
                                if (live[0]) v[u].x = __ldg(xr + cols[0]);
                                if (live[1]) v[u].y = __ldg(xr + cols[1]);
                                if (live[2]) v[u].z = __ldg(xr + cols[2]);
                                if (live[3]) v[u].w = __ldg(xr + cols[3]);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (on[u]) {
                            acc[0] = __fadd_rn(acc[0], __fmul_rn(cf[u], v[u].x));
                            acc[1] = __fadd_rn(acc[1], __fmul_rn(cf[u], v[u].y));
                            acc[2] = __fadd_rn(acc[2], __fmul_rn(cf[u], v[u].z));
                            acc[3] = __fadd_rn(acc[3], __fmul_rn(cf[u], v[u].w));
                        }
                    }
                }
            }
            // the self-loop term, then the stores
            const float *xi = a.x + i * a.x_stride;
            float *orow = a.out + i * a.out_stride;
            if (VEC4) {
                if (live[0]) {
                    const float4 s = __ldg(reinterpret_cast<const float4 *>(xi + cols[0]));
                    float4 r;
                    r.x = __fadd_rn(acc[0], __fmul_rn(self_coef, s.x));
                    r.y = __fadd_rn(acc[1], __fmul_rn(self_coef, s.y));
                    r.z = __fadd_rn(acc[2], __fmul_rn(self_coef, s.z));
                    r.w = __fadd_rn(acc[3], __fmul_rn(self_coef, s.w));
                    for (int k = 0; k < a.copies; ++k) *reinterpret_cast<float4 *>(orow + k * a.F + cols[0]) = r;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (live[q]) {
                        const float r = __fadd_rn(acc[q], __fmul_rn(self_coef, __ldg(xi + cols[q])));
                        for (int k = 0; k < a.copies; ++k) orow[k * a.F + cols[q]] = r;
                    }
                }
            }
        }
    }
}

}  // namespace ss

extern "C" {

int64_t ss_sign_workspace_bytes(int64_t n_nodes) {
    if (n_nodes < 0) return SS_ERR_INVALID;
    return ss::sign_ws_bytes(n_nodes);
}

int ss_gcn_norm(const int64_t *row, const int64_t *col, const float *edge_weight, int64_t n_edges, int64_t n_nodes,
                float *dinv_out, float *loop_weight_out, int32_t *flags_out, void *workspace, int64_t workspace_bytes,
                ss_stream_t stream) {
    SS_REQUIRE(n_edges >= 0 && n_nodes >= 0, "negative size passed to ss_gcn_norm");
    SS_REQUIRE(n_edges < (1ll << 31), "at most 2^31-1 edges");
    if (n_nodes == 0) {
        if (flags_out) SS_CUDA(cudaMemsetAsync(flags_out, n_edges > 0 ? 2 : 0, 1, (cudaStream_t)stream));
        return SS_OK;
    }
    SS_REQUIRE(dinv_out && loop_weight_out && workspace, "null pointer passed to ss_gcn_norm");
    SS_REQUIRE(n_edges == 0 || (row && col), "row / col is null");
    SS_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    if (workspace_bytes < ss::sign_ws_bytes(n_nodes)) {
        ss::set_error("sign workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)ss::sign_ws_bytes(n_nodes));
        return SS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *loop_eid = (int32_t *)workspace;
    SS_CUDA(cudaMemsetAsync(loop_eid, 0xff, (size_t)n_nodes * 4, st));  // -1
    if (flags_out) SS_CUDA(cudaMemsetAsync(flags_out, 0, 4, st));
    if (n_edges > 0) {
        ss::sign_loops_kernel<<<ss::sg_grid(n_edges), 256, 0, st>>>(row, col, n_edges, n_nodes, loop_eid, flags_out);
        SS_LAUNCH_CHECK("sign_loops_kernel");
    }
    ss::sign_loop_weight_kernel<<<ss::sg_grid(n_nodes), 256, 0, st>>>(loop_eid, edge_weight, n_nodes, loop_weight_out, dinv_out);
    SS_LAUNCH_CHECK("sign_loop_weight_kernel");
    if (n_edges > 0) {
        ss::sign_degree_kernel<<<ss::sg_grid(n_edges), 256, 0, st>>>(row, col, edge_weight, n_edges, n_nodes, dinv_out);
        SS_LAUNCH_CHECK("sign_degree_kernel");
    }
    ss::sign_dinv_kernel<<<ss::sg_grid(n_nodes), 256, 0, st>>>(dinv_out, n_nodes);
    SS_LAUNCH_CHECK("sign_dinv_kernel");
    return SS_OK;
}

int ss_sign_fill(const int64_t *row, int64_t n_edges, int64_t n_nodes, const int64_t *rowptr, int32_t *perm_out,
                 void *workspace, int64_t workspace_bytes, ss_stream_t stream) {
    SS_REQUIRE(n_edges >= 0 && n_nodes >= 0, "negative size passed to ss_sign_fill");
    SS_REQUIRE(n_edges < (1ll << 31), "at most 2^31-1 edges");
    if (n_nodes == 0 || n_edges == 0) return SS_OK;
    SS_REQUIRE(row && rowptr && perm_out && workspace, "null pointer passed to ss_sign_fill");
    SS_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    if (workspace_bytes < ss::sign_ws_bytes(n_nodes)) {
        ss::set_error("sign workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)ss::sign_ws_bytes(n_nodes));
        return SS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t *cursor = (uint32_t *)((char *)workspace + ss::sg_align_up(n_nodes * 4, 256));
    SS_CUDA(cudaMemsetAsync(cursor, 0, (size_t)n_nodes * 4, st));
    ss::sign_fill_kernel<<<ss::sg_grid(n_edges), 256, 0, st>>>(row, n_edges, n_nodes, rowptr, cursor, perm_out);
    SS_LAUNCH_CHECK("sign_fill_kernel");
    return SS_OK;
}

int ss_sign_spmm(const int64_t *rowptr, const int32_t *perm, const int64_t *col, const float *edge_weight,
                 const float *dinv, const float *loop_weight, const float *x, int64_t x_stride, int64_t n_nodes,
                 int64_t n_features, float *out, int64_t out_stride, int copies, ss_stream_t stream) {
    SS_REQUIRE(n_nodes >= 0 && n_features >= 0 && copies >= 1, "bad size passed to ss_sign_spmm");
    if (n_nodes == 0 || n_features == 0) return SS_OK;
    SS_REQUIRE(rowptr && dinv && loop_weight && x && out, "null pointer passed to ss_sign_spmm");
    SS_REQUIRE(x_stride >= n_features && out_stride >= (int64_t)copies * n_features, "row strides are too small");
    ss::SpmmArgs a;
    a.rowptr = rowptr; a.perm = perm; a.col = col; a.ew = edge_weight; a.dinv = dinv; a.loop_w = loop_weight;
    a.x = x; a.x_stride = x_stride; a.n_nodes = n_nodes; a.F = n_features;
    a.out = out; a.out_stride = out_stride; a.copies = copies;
    const bool vec4 = (n_features % 4 == 0) && (x_stride % 4 == 0) && (out_stride % 4 == 0) &&
                      (((uintptr_t)x | (uintptr_t)out) & 15) == 0;
    int64_t blocks = (n_nodes + 7) / 8;
    int64_t cap = (int64_t)ss::sm_count() * 16;
    const int grid = (int)(blocks < cap ? blocks : cap);
    cudaStream_t st = (cudaStream_t)stream;
    if (vec4) ss::sign_spmm_kernel<true><<<grid, 256, 0, st>>>(a);
    else ss::sign_spmm_kernel<false><<<grid, 256, 0, st>>>(a);
    SS_LAUNCH_CHECK("sign_spmm_kernel");
    return SS_OK;
}

}  // extern "C"
