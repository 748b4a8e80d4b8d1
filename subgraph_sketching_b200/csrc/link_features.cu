// link_features.cu -- K4: pairwise structural features over a batch of candidate links.
//
// Replaces ElphHashes.get_subgraph_features (/root/reference/src/hashing.py:258-323) including
// _get_intersections (:167-189), jaccard (:247-256), _hll_merge (:234-237) and the per-combination
// hll_count (:186).  The reference gathers 4 rows for each of the K^2 hop combinations and materialises
// [n, m] float and [n, T] argsort temporaries; here one warp owns one link, loads the 2K records of (u, v)
// once (each lane keeps 16 B of MinHash + 8 B of HLL per record in registers), evaluates all K^2
// combinations from registers and finishes with the inclusion-exclusion algebra in the reference's order.
#include "common.cuh"

namespace ss {

constexpr int REC_MH = 512;

struct LinkArgs {
    const int64_t *links;
    int64_t n_links;
    const uint8_t *hop[4];  // records by hop (index 0 unused)
    int64_t stride[4];
    const float *cards;
    int64_t cards_stride;
    HllDev h;
    int flags;
    float *features;
    float *inter;
    RecordShape s;
};

// ---- per-combination statistics of the union sketch, default shape ---------------------------------
// Fast path: every register <= 28, so sum 2^-r fits a 32-bit fixed-point lane accumulator in units of 2^-28
// (8 registers per lane, each <= 2^28).  Otherwise the exact 64-bit accumulator is used.  Both give the
// exactly rounded float32 of the true sum.
__device__ __forceinline__ uint32_t sum_pow_word28(uint32_t w) {
    return (0x10000000u >> (w & 0xffu)) + (0x10000000u >> ((w >> 8) & 0xffu)) + (0x10000000u >> ((w >> 16) & 0xffu)) +
           (0x10000000u >> (w >> 24));
}

// scalar tail: (zeros, S = sum 2^-r as float, matches) -> jaccard * union cardinality (hashing.py:184-187)
__device__ __forceinline__ float intersection_tail(const HllDev &h, int zeros, float S, uint32_t matches, int P) {
    float val = __fadd_rn(h.threshold, 1.0f);
    if (zeros > 0) val = __ldg(h.lc + zeros);
    if (val > h.threshold) {
        float e = __fmul_rn(__frcp_rn(S), h.alpha_m2);
        if (e <= h.five_m) e = __fsub_rn(e, bias_6nn(h, e));
        val = e;
    }
    const float jac = __fdiv_rn((float)matches, (float)P);
    return __fmul_rn(jac, val);
}

// inclusion-exclusion algebra in the reference's left-to-right order (hashing.py:276-307);
// I[(k1-1)*K + (k2-1)], cu/cv = cards of u / v.  torch.sum over a column slice is evaluated sequentially.
template <int K>
__device__ __forceinline__ void feature_algebra(const float *I, const float *cu, const float *cv, float *f) {
#define SUB(a, b) __fsub_rn(a, b)
#define ADD(a, b) __fadd_rn(a, b)
    f[0] = I[0];
    if (K == 1) {
        f[1] = SUB(cv[0], f[0]);
        f[2] = SUB(cu[0], f[0]);
    } else if (K == 2) {
        f[1] = SUB(I[1 * 2 + 0], f[0]);                                   // (2,1)
        f[2] = SUB(I[0 * 2 + 1], f[0]);                                   // (1,2)
        f[3] = SUB(SUB(SUB(I[1 * 2 + 1], f[0]), f[1]), f[2]);             // (2,2)
        f[4] = SUB(cv[0], ADD(f[0], f[1]));                               // (0,1)
        f[5] = SUB(SUB(cu[0], f[0]), f[2]);                               // (1,0)
        float s5 = ADD(ADD(ADD(ADD(f[0], f[1]), f[2]), f[3]), f[4]);
        f[6] = SUB(cv[1], s5);                                            // (0,2)
        float s4 = ADD(ADD(ADD(f[0], f[1]), f[2]), f[3]);
        f[7] = SUB(SUB(SUB(cu[1], f[0]), s4), f[5]);                      // (2,0) -- f0 twice, as the reference
    } else {
        f[1] = SUB(I[1 * 3 + 0], f[0]);                                   // (2,1)
        f[2] = SUB(I[0 * 3 + 1], f[0]);                                   // (1,2)
        f[3] = SUB(SUB(SUB(I[1 * 3 + 1], f[0]), f[1]), f[2]);             // (2,2)
        f[4] = SUB(SUB(I[2 * 3 + 0], f[0]), f[1]);                        // (3,1)
        f[5] = SUB(SUB(I[0 * 3 + 2], f[0]), f[2]);                        // (1,3)
        float s4 = ADD(ADD(ADD(f[0], f[1]), f[2]), f[3]);
        f[6] = SUB(SUB(I[2 * 3 + 1], s4), f[4]);                          // (3,2)
        f[7] = SUB(SUB(I[1 * 3 + 2], s4), f[5]);                          // (2,3)
        float s8 = ADD(ADD(ADD(ADD(s4, f[4]), f[5]), f[6]), f[7]);
        f[8] = SUB(I[2 * 3 + 2], s8);                                     // (3,3)
        f[9] = SUB(SUB(SUB(cv[0], f[0]), f[1]), f[4]);                    // (0,1)
        f[10] = SUB(SUB(SUB(cu[0], f[0]), f[2]), f[5]);                   // (1,0)
        float s5 = ADD(s4, f[4]);
        f[11] = SUB(SUB(SUB(cv[1], s5), f[6]), f[9]);                     // (0,2)
        f[12] = SUB(SUB(SUB(cu[1], s5), f[7]), f[10]);                    // (2,0) -- s5 holds (3,1), as the reference
        float s9 = ADD(s8, f[8]);
        f[13] = SUB(SUB(SUB(cv[2], s9), f[9]), f[11]);                    // (0,3)
        f[14] = SUB(SUB(SUB(cu[2], s9), f[10]), f[12]);                   // (3,0)
    }
#undef SUB
#undef ADD
}

template <int K>
__device__ __forceinline__ void knockout_and_floor(float *f, int flags) {
    if (!(flags & SS_FLAG_USE_ZERO_ONE)) {  // hashing.py:310-318
        if (K == 2) { f[4] = 0.f; f[5] = 0.f; }
        if (K == 3) { f[4] = 0.f; f[5] = 0.f; f[11] = 0.f; f[12] = 0.f; }
    }
    if (flags & SS_FLAG_FLOOR) {  // hashing.py:319-320
#pragma unroll
        for (int i = 0; i < K * (K + 2); ++i)
            if (f[i] < 0.f) f[i] = 0.f;
    }
}

// ---- default shape (P=128, p=8): one warp per link, records in registers ----------------------------
template <int K>
__global__ void __launch_bounds__(256) link_features_kernel(const LinkArgs a) {
    constexpr int F = K * (K + 2);
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = gwarp; i < a.n_links; i += n_warps) {
        const int64_t u = __ldg(a.links + 2 * i), v = __ldg(a.links + 2 * i + 1);
        uint4 mu[K], mv[K];
        uint2 hu[K], hv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint8_t *ru = a.hop[k + 1] + u * a.stride[k + 1];
            const uint8_t *rv = a.hop[k + 1] + v * a.stride[k + 1];
            mu[k] = ld_nc_u4(ru + lane * 16);
            hu[k] = ld_nc_u2(ru + REC_MH + lane * 8);
            mv[k] = ld_nc_u4(rv + lane * 16);
            hv[k] = ld_nc_u2(rv + REC_MH + lane * 8);
        }
        float cu[K], cv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            cu[k] = __ldg(a.cards + u * a.cards_stride + k);
            cv[k] = __ldg(a.cards + v * a.cards_stride + k);
        }
        // combination c = (k1-1)*K + (k2-1) ends up on lane c: (zeros, S, matches)
        int my_zeros = 0;
        float my_S = 1.f;
        uint32_t my_match = 0;
#pragma unroll
        for (int k1 = 0; k1 < K; ++k1) {
#pragma unroll
            for (int k2 = 0; k2 < K; ++k2) {
                uint32_t eq = (mu[k1].x == mv[k2].x) + (mu[k1].y == mv[k2].y) + (mu[k1].z == mv[k2].z) +
                              (mu[k1].w == mv[k2].w);
                const uint32_t wx = __vmaxu4(hu[k1].x, hv[k2].x), wy = __vmaxu4(hu[k1].y, hv[k2].y);
                const uint32_t big = (((wx + 0x63636363u) | wx) | ((wy + 0x63636363u) | wy)) & 0x80808080u;  // any register > 28
                int zeros;
                float S;
                if (!__any_sync(FULL, big != 0u)) {
                    const uint32_t nzx = (wx + 0x7f7f7f7fu) & 0x80808080u, nzy = (wy + 0x7f7f7f7fu) & 0x80808080u;
                    const uint32_t acc = sum_pow_word28(wx) + sum_pow_word28(wy);  // <= 2^31
                    // pack: zeros (<= 8 per lane) ride in the low-half reduction's spare bits
                    const uint32_t lo = __reduce_add_sync(FULL, acc & 0xffffu);
                    const uint32_t hi = __reduce_add_sync(FULL, acc >> 16);
                    const uint32_t nz = __reduce_add_sync(FULL, (uint32_t)(__popc(nzx) + __popc(nzy)));
                    zeros = 256 - (int)nz;
                    const uint64_t total = (uint64_t)lo + ((uint64_t)hi << 16);
                    S = __fmul_rn(__ull2float_rn(total), 3.7252902984619140625e-09f);  // * 2^-28, exact
                } else {
                    uint64_t acc = 0;
                    int nz = 0;
                    acc_regs_word(wx, acc, nz);
                    acc_regs_word(wy, acc, nz);
                    unsigned __int128 t = warp_total_units(acc, nz, zeros);
                    S = units_to_f32(t);
                }
                eq = __reduce_add_sync(FULL, eq);
                if (lane == k1 * K + k2) {
                    my_zeros = zeros;
                    my_S = S;
                    my_match = eq;
                }
            }
        }
        float my_inter = 0.f;
        if (lane < K * K) my_inter = intersection_tail(a.h, my_zeros, my_S, my_match, 128);
        float I[K * K];
#pragma unroll
        for (int c = 0; c < K * K; ++c) I[c] = __shfl_sync(FULL, my_inter, c);
        if (a.inter && lane < K * K) a.inter[i * (K * K) + lane] = my_inter;
        if (a.features) {
            float f[F];
            feature_algebra<K>(I, cu, cv, f);
            knockout_and_floor<K>(f, a.flags);
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < F; ++j)
                if (lane == j) mine = f[j];
            if (lane < F) a.features[i * F + lane] = mine;
        }
    }
}

// ---- generic shape: one warp per link, rows re-read per combination (L1/L2 resident) ---------------
template <int K>
__global__ void __launch_bounds__(256) link_features_generic_kernel(const LinkArgs a) {
    constexpr int F = K * (K + 2);
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = gwarp; i < a.n_links; i += n_warps) {
        const int64_t u = __ldg(a.links + 2 * i), v = __ldg(a.links + 2 * i + 1);
        float cu[K], cv[K], I[K * K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            cu[k] = __ldg(a.cards + u * a.cards_stride + k);
            cv[k] = __ldg(a.cards + v * a.cards_stride + k);
        }
#pragma unroll
        for (int k1 = 0; k1 < K; ++k1) {
#pragma unroll
            for (int k2 = 0; k2 < K; ++k2) {
                const uint8_t *ru = a.hop[k1 + 1] + u * a.stride[k1 + 1];
                const uint8_t *rv = a.hop[k2 + 1] + v * a.stride[k2 + 1];
                uint32_t eq = 0;
                for (int j = lane; j < a.s.P; j += 32)
                    eq += (*reinterpret_cast<const uint32_t *>(ru + 4 * j) == *reinterpret_cast<const uint32_t *>(rv + 4 * j)) ? 1u : 0u;
                eq = __reduce_add_sync(FULL, eq);
                RegSum rs;
                rs.lo = rs.hi = 0;
                rs.zeros = 0;
                for (int x = lane; x < (a.s.m >> 3); x += 32) {
                    const uint2 p = *reinterpret_cast<const uint2 *>(ru + a.s.mh_bytes + 8 * x);
                    const uint2 q = *reinterpret_cast<const uint2 *>(rv + a.s.mh_bytes + 8 * x);
                    regsum_add_word(rs, __vmaxu4(p.x, q.x));
                    regsum_add_word(rs, __vmaxu4(p.y, q.y));
                }
                int zeros;
                unsigned __int128 t = regsum_warp_total(rs, zeros);
                I[k1 * K + k2] = intersection_tail(a.h, zeros, units_to_f32(t), eq, a.s.P);
            }
        }
        if (a.inter && lane < K * K) {
            float mine = 0.f;
#pragma unroll
            for (int c = 0; c < K * K; ++c)
                if (lane == c) mine = I[c];
            a.inter[i * (K * K) + lane] = mine;
        }
        if (a.features) {
            float f[F];
            feature_algebra<K>(I, cu, cv, f);
            knockout_and_floor<K>(f, a.flags);
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < F; ++j)
                if (lane == j) mine = f[j];
            if (lane < F) a.features[i * F + lane] = mine;
        }
    }
}

template <int K>
static int launch_links(const LinkArgs &a, bool fast, cudaStream_t st) {
    int64_t blocks = (a.n_links + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (fast) {
        link_features_kernel<K><<<grid, 256, 0, st>>>(a);
        SS_LAUNCH_CHECK("link_features_kernel");
    } else {
        link_features_generic_kernel<K><<<grid, 256, 0, st>>>(a);
        SS_LAUNCH_CHECK("link_features_generic_kernel");
    }
    return SS_OK;
}

}  // namespace ss

extern "C" {

int ss_link_features(const int64_t *links, int64_t n_links, const ss_hop_view *hops, int max_hops, int num_perm,
                     int hll_p, const float *cards, int64_t cards_stride, const ss_hll_consts *hc, int flags,
                     float *features_out, float *inter_out, ss_stream_t stream) {
    ss::RecordShape s;
    SS_REQUIRE(ss::make_shape(num_perm, hll_p, &s), "unsupported sketch shape num_perm=%d hll_p=%d", num_perm, hll_p);
    SS_REQUIRE(max_hops >= 1 && max_hops <= 3, "Only 1, 2 and 3 hop hashes are implemented (got %d)", max_hops);
    SS_REQUIRE(n_links >= 0, "n_links must be >= 0");
    if (n_links == 0) return SS_OK;
    SS_REQUIRE(links && hops, "null pointer passed to ss_link_features");
    SS_REQUIRE(features_out || inter_out, "no output requested");
    SS_REQUIRE(!features_out || cards, "cards are required for features");
    int rc = ss::check_hll_consts(hc, hll_p);
    if (rc != SS_OK) return rc;
    ss::LinkArgs a;
    memset(&a, 0, sizeof(a));
    a.links = links;
    a.n_links = n_links;
    for (int k = 1; k <= max_hops; ++k) {
        SS_REQUIRE(hops[k].records && ((uintptr_t)hops[k].records & 15) == 0, "hop %d records must be a 16-byte aligned device pointer", k);
        SS_REQUIRE(hops[k].row_stride >= s.bytes && (hops[k].row_stride & 15) == 0, "hop %d has a bad row stride", k);
        a.hop[k] = (const uint8_t *)hops[k].records;
        a.stride[k] = hops[k].row_stride;
    }
    // inter-only calls read no cards: point at a valid dummy (the hll lc table) with stride 0
    a.cards = cards ? cards : hc->lc_table;
    a.cards_stride = cards ? cards_stride : 0;
    a.h = ss::to_dev(hc);
    a.flags = flags;
    a.features = features_out;
    a.inter = inter_out;
    a.s = s;
    const bool fast = (num_perm == 128 && hll_p == 8);
    cudaStream_t st = (cudaStream_t)stream;
    switch (max_hops) {
        case 1: return ss::launch_links<1>(a, fast, st);
        case 2: return ss::launch_links<2>(a, fast, st);
        default: return ss::launch_links<3>(a, fast, st);
    }
}

}  // extern "C"
