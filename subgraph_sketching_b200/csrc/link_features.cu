// link_features.cu -- K4: pairwise structural features over a batch of candidate links.
//
// Replaces ElphHashes.get_subgraph_features (/root/reference/src/hashing.py:258-323) including
// _get_intersections (:167-189), jaccard (:247-256), _hll_merge (:234-237) and the per-combination
// hll_count (:186).  The reference gathers 4 rows for each of the K^2 hop combinations and materialises
// [n, m] float and [n, T] argsort temporaries; here one warp owns one link, loads the 2K records of (u, v)
// once (each lane keeps 16 B of MinHash + 8 B of HLL per record in registers), evaluates all K^2
// combinations from registers and finishes with the inclusion-exclusion algebra in the reference's order.
#include <cuda.h>
#include <stdlib.h>

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
    int64_t n_nodes;  // rows of the smallest hop table: link endpoints must be in [0, n_nodes)
    int *err;         // set to 1 if any endpoint is out of range (such links are evaluated on node 0)
    // node-sharded tables (multi-GPU, batched kernel only): this rank's copies hold its own row block plus the
    // halo rows it gathered during the build; any other row is read from its OWNER's copy over NVLink
    int n_ranks, rank, last_hop_own_only;
    int64_t bounds[SS_MAX_PEERS + 2];
    const uint8_t *peer_hop[4][SS_MAX_PEERS + 1];
    const uint8_t *local_rows;
};

// bounds check of a link endpoint (the reference would raise IndexError from torch indexing)
__device__ __forceinline__ int64_t checked_node(const LinkArgs &a, int64_t id) {
    if ((uint64_t)id < (uint64_t)a.n_nodes) return id;
    if (a.err) atomicExch(a.err, 1);
    return 0;
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
// Per record a lane keeps: 4 MinHash slots, its 8 HLL registers split into an even-byte and an odd-byte plane
// (two 16-bit lanes per word, so the union max is the native VIMNMX.U16x2 -- a 4 x uint8 max costs 7
// instructions), and an 8-bit non-zero mask.  Per combination: 4 compares, 4 maxes, 8 funnel shifts for
// sum 2^-r (shf.r.wrap needs no byte extraction for the low byte of a 16-bit lane), one popc for the zero
// count.  Match counts and zero counts of all K^2 combinations are packed into bit fields and reduced across
// the warp once per word instead of once per combination.
struct RowRegs {
    uint4 mh;
    uint2 he, ho;   // even / odd byte planes of the lane's 8 HLL registers
    uint32_t nz;    // bit set per non-zero register (bits 7,15,23,31 from word x; 6,14,22,30 from word y)
};

__device__ __forceinline__ void prep_row(RowRegs &r, const uint4 &m, const uint2 &h, uint32_t &big) {
    r.mh = m;
    r.he = make_uint2(h.x & 0x00ff00ffu, h.y & 0x00ff00ffu);
    r.ho = make_uint2(h.x & 0xff00ff00u, h.y & 0xff00ff00u);
    const uint32_t nzx = ((h.x + 0x7f7f7f7fu) | h.x) & 0x80808080u;
    const uint32_t nzy = ((h.y + 0x7f7f7f7fu) | h.y) & 0x80808080u;
    r.nz = nzx | (nzy >> 1);
    big |= (((h.x + 0x63636363u) | h.x) | ((h.y + 0x63636363u) | h.y)) & 0x80808080u;  // any register > 28
}

// sum over the two 16-bit lanes of an even-plane word (register value in the LOW byte of each lane) of
// 2^(28 - r); shf.r.wrap uses only the low 5 bits of the shift amount
__device__ __forceinline__ uint32_t pow_sum_even(uint32_t w) {
    return __funnelshift_r(0x10000000u, 0u, w) + __funnelshift_r(0x10000000u, 0u, w >> 16);
}
// same for an odd-plane word (register value in the HIGH byte of each lane)
__device__ __forceinline__ uint32_t pow_sum_odd(uint32_t w) {
    return __funnelshift_r(0x10000000u, 0u, w >> 8) + __funnelshift_r(0x10000000u, 0u, w >> 24);
}

// all K^2 combinations, tails, algebra and stores of ONE link whose 2K records are in registers
template <int K>
__device__ __forceinline__ void process_link(const LinkArgs &a, int64_t i, const uint4 *mu, const uint2 *hu,
                                             const uint4 *mv, const uint2 *hv, const float *cu, const float *cv,
                                             int lane) {
    constexpr int F = K * (K + 2);
    constexpr int C = K * K;
    RowRegs U[K], V[K];
    uint32_t big = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        prep_row(U[k], mu[k], hu[k], big);
        prep_row(V[k], mv[k], hv[k], big);
    }
    const bool any_big = __any_sync(FULL, big != 0u);  // rare: some register > 28 -> exact 128-bit path

    // per-lane partial statistics of every combination c = (k1-1)*K + (k2-1)
    uint32_t eq_pack[(C + 3) / 4] = {0};   // 8-bit fields: matches per lane <= 4, per warp <= 128
    uint32_t nz_pack[(C + 2) / 3] = {0};   // 10-bit fields: non-zero registers per lane <= 8, per warp <= 256
    int my_zeros = 0;
    float my_S = 1.f;
#pragma unroll
    for (int k1 = 0; k1 < K; ++k1) {
#pragma unroll
        for (int k2 = 0; k2 < K; ++k2) {
            const int c = k1 * K + k2;
            const uint32_t eq = (U[k1].mh.x == V[k2].mh.x) + (U[k1].mh.y == V[k2].mh.y) +
                                (U[k1].mh.z == V[k2].mh.z) + (U[k1].mh.w == V[k2].mh.w);
            eq_pack[c / 4] += eq << (8 * (c % 4));
            nz_pack[c / 3] += (uint32_t)__popc(U[k1].nz | V[k2].nz) << (10 * (c % 3));
        }
    }
    // the (rare) exact path is hoisted out of the combination loop so that the common path is straight-line
    // code: the K^2 independent reduction chains (REDUX -> I2F -> FMUL) can then overlap instead of each one
    // waiting behind a branch
    if (!any_big) {
#pragma unroll
        for (int k1 = 0; k1 < K; ++k1) {
#pragma unroll
            for (int k2 = 0; k2 < K; ++k2) {
                const int c = k1 * K + k2;
                const uint32_t ex = __vmaxu2(U[k1].he.x, V[k2].he.x), ey = __vmaxu2(U[k1].he.y, V[k2].he.y);
                const uint32_t ox = __vmaxu2(U[k1].ho.x, V[k2].ho.x), oy = __vmaxu2(U[k1].ho.y, V[k2].ho.y);
                const uint32_t acc = pow_sum_even(ex) + pow_sum_even(ey) + pow_sum_odd(ox) + pow_sum_odd(oy);  // <= 2^31
                const uint32_t lo = __reduce_add_sync(FULL, acc & 0xffffu);
                const uint32_t hi = __reduce_add_sync(FULL, acc >> 16);
                // lo < 2^21 and hi < 2^20 are exact in float32 and so is hi * 2^16: one rounding in the add gives
                // the correctly rounded total
                const float S = __fmul_rn(__fadd_rn(__fmul_rn((float)hi, 65536.f), (float)lo), 3.7252902984619140625e-09f);
                if (lane == c) my_S = S;
            }
        }
    } else {
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            const int k1 = c / K, k2 = c % K;
            uint2 ue = make_uint2(0u, 0u), uo = ue, ve = ue, vo = ue;
#pragma unroll
            for (int k = 0; k < K; ++k) {  // select without dynamic register indexing
                if (k == k1) { ue = U[k].he; uo = U[k].ho; }
                if (k == k2) { ve = V[k].he; vo = V[k].ho; }
            }
            uint64_t acc = 0;
            int nz = 0, zeros;
            acc_regs_word(__vmaxu2(ue.x, ve.x) | __vmaxu2(uo.x, vo.x), acc, nz);
            acc_regs_word(__vmaxu2(ue.y, ve.y) | __vmaxu2(uo.y, vo.y), acc, nz);
            unsigned __int128 t = warp_total_units(acc, nz, zeros);
            const float S = units_to_f32(t);
            if (lane == c) {
                my_S = S;
                my_zeros = zeros;
            }
        }
    }
    uint32_t my_match = 0;
#pragma unroll
    for (int w = 0; w < (C + 3) / 4; ++w) {
        const uint32_t r = __reduce_add_sync(FULL, eq_pack[w]);
        if (lane / 4 == w) my_match = (r >> (8 * (lane % 4))) & 0xffu;
    }
#pragma unroll
    for (int w = 0; w < (C + 2) / 3; ++w) {
        const uint32_t r = __reduce_add_sync(FULL, nz_pack[w]);
        if (lane / 3 == w && !any_big) my_zeros = 256 - (int)((r >> (10 * (lane % 3))) & 0x3ffu);
    }
    float my_inter = 0.f;
    if (lane < C) my_inter = intersection_tail(a.h, my_zeros, my_S, my_match, 128);
    float I[C];
#pragma unroll
    for (int c = 0; c < C; ++c) I[c] = __shfl_sync(FULL, my_inter, c);
    if (a.inter && lane < C) a.inter[i * C + lane] = my_inter;
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

// ---- LDG front end: the 2K records are loaded straight into registers -----------------------------------
template <int K>
__global__ void __launch_bounds__(256, 3) link_features_kernel(const LinkArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // the endpoints of the NEXT link are fetched while the current one is evaluated: the id load (a dependent
    // round trip in front of every record load -- over PCIe when the link list is a pinned host buffer) leaves the
    // critical path
    int64_t u_next = 0, v_next = 0;
    if (gwarp < a.n_links) {
        u_next = __ldg(a.links + 2 * gwarp);
        v_next = __ldg(a.links + 2 * gwarp + 1);
    }
    for (int64_t i = gwarp; i < a.n_links; i += n_warps) {
        const int64_t u = checked_node(a, u_next), v = checked_node(a, v_next);
        if (i + n_warps < a.n_links) {
            u_next = __ldg(a.links + 2 * (i + n_warps));
            v_next = __ldg(a.links + 2 * (i + n_warps) + 1);
        }
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
        process_link<K>(a, i, mu, hu, mv, hv, cu, cv, lane);
    }
}

// ---- batched front end (default): tiles of consecutive links, tails and algebra evaluated for a whole batch --------
// The per-link kernel above spends ~280 of its ~970 warp instructions per link (K = 3) in code that only 9 or 15
// lanes execute: the K^2 scalar tails (linear counting / raw estimate / 6-NN bias search) and the inclusion-
// exclusion algebra.  Here a warp owns a TILE of consecutive links and walks it in batches of B = 32 / K^2 links:
// the elementwise part of each link leaves its K^2 raw statistics (matches, zero count, fixed-point sum) in the
// slot lanes j*K^2 .. j*K^2 + K^2 - 1 of the batch, then ONE pass of the tail code serves all B*K^2 combinations and
// one pass of the algebra (lane j = link j of the batch) serves all B links; features leave through a shared-memory
// transpose as one contiguous store.  Consecutive links with the same source (the reference's ranking evaluation:
// 1 positive + 1000 negatives per source, data.py:226-230) reuse u's K prepared records instead of re-loading them.
constexpr int LK_TILE_MAX = 96;  // links per tile: a multiple of B for K = 1, 2, 3 (B = 32, 8, 3)

struct RowRegsB {
    uint4 mh;
    uint2 he, ho;   // even-byte plane, odd-byte plane shifted down: register value in the LOW byte of each 16-bit lane
    uint32_t nz;
};

__device__ __forceinline__ void prep_row_b(RowRegsB &r, const uint4 &m, const uint2 &h, uint32_t &big) {
    r.mh = m;
    r.he = make_uint2(h.x & 0x00ff00ffu, h.y & 0x00ff00ffu);
    r.ho = make_uint2((h.x >> 8) & 0x00ff00ffu, (h.y >> 8) & 0x00ff00ffu);
    const uint32_t nzx = ((h.x + 0x7f7f7f7fu) | h.x) & 0x80808080u;
    const uint32_t nzy = ((h.y + 0x7f7f7f7fu) | h.y) & 0x80808080u;
    r.nz = nzx | (nzy >> 1);
    big |= (((h.x + 0x63636363u) | h.x) | ((h.y + 0x63636363u) | h.y)) & 0x80808080u;  // any register > 28
}

// 1 if x != 0 else 0 as ONE min instruction.  Written as opaque PTX: left to the compiler, min(x ^ y, 1) is turned back
// into a compare + predicated moves (3 instructions per MinHash slot instead of 2).
__device__ __forceinline__ uint32_t lk_nonzero(uint32_t x) {
    uint32_t r;
    asm("min.u32 %0, %1, 1;" : "=r"(r) : "r"(x));
    return r;
}

__device__ __forceinline__ int lk_select3(int idx, int a0, int a1, int a2) { return idx == 0 ? a0 : (idx == 1 ? a1 : a2); }

__device__ __forceinline__ void lk_cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void lk_cp_async4(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void lk_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// which rank owns row `node` (row blocks are contiguous: bounds[q] <= node < bounds[q + 1])
__device__ __forceinline__ int lk_owner(const LinkArgs &a, int node) {
    int o = 0;
    for (int q = 1; q < a.n_ranks; ++q) o += ((int64_t)node >= a.bounds[q]) ? 1 : 0;
    return o;
}
// base pointer of the copy of hop table k1 (1-based) that holds a valid record of a node owned by `owner`;
// `have` = this rank's copy of the replicated-by-halo hops holds the row
__device__ __forceinline__ const uint8_t *lk_table(const LinkArgs &a, int k1, int K, int owner, bool have) {
    if (a.n_ranks <= 1 || owner == a.rank) return a.hop[k1];
    if (have && !(k1 == K && a.last_hop_own_only)) return a.hop[k1];
    return a.peer_hop[k1][owner];
}

__device__ __forceinline__ void lk_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void lk_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lk_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void lk_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint4 lk_lds_u4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint2 lk_lds_u2(uint32_t addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}

// Per-lane slot of the batched tail: the raw statistics of ONE (link, combination) of the running batch
struct LkSlot {
    uint32_t lo, hi, match;   // fixed-point sum of 2^-r (two halves), matching MinHash slots
    int zeros;                // empty registers of the union
    float Sx;                 // >= 0: exact-path sum (some register > 28)
};

// all K^2 combinations of one link whose prepared records are in registers; lane (j, c) keeps combination c
template <int K>
__device__ __forceinline__ void lk_combine(const RowRegsB *U, const RowRegsB *V, uint32_t big, int j, int my_j, int my_c,
                                           LkSlot &sl) {
    constexpr int C = K * K;
    constexpr int EQW = (C + 3) / 4, NZW = (C + 2) / 3;
    const bool any_big = __any_sync(FULL, big != 0u);
    const bool mine = (my_j == j);
    const int my_slot = mine ? my_c : -1;  // combination this lane keeps for this link
    uint32_t eq_pack[EQW] = {0}, nz_pack[NZW] = {0};
#pragma unroll
    for (int k1 = 0; k1 < K; ++k1) {
#pragma unroll
        for (int k2 = 0; k2 < K; ++k2) {
            const int c = k1 * K + k2;
            // mismatching slots: min(x ^ y, 1) summed (XOR + VIMNMX per slot, IADD3 for the sums)
            const uint32_t ne = lk_nonzero(U[k1].mh.x ^ V[k2].mh.x) + lk_nonzero(U[k1].mh.y ^ V[k2].mh.y) +
                                lk_nonzero(U[k1].mh.z ^ V[k2].mh.z) + lk_nonzero(U[k1].mh.w ^ V[k2].mh.w);
            eq_pack[c / 4] += ne << (8 * (c % 4));  // MISmatches; turned into matches at extraction
            nz_pack[c / 3] += (uint32_t)__popc(U[k1].nz | V[k2].nz) << (10 * (c % 3));
        }
    }
    if (!any_big) {
#pragma unroll
        for (int k1 = 0; k1 < K; ++k1) {
#pragma unroll
            for (int k2 = 0; k2 < K; ++k2) {
                const int c = k1 * K + k2;
                const uint32_t ex = __vmaxu2(U[k1].he.x, V[k2].he.x), ey = __vmaxu2(U[k1].he.y, V[k2].he.y);
                const uint32_t ox = __vmaxu2(U[k1].ho.x, V[k2].ho.x), oy = __vmaxu2(U[k1].ho.y, V[k2].ho.y);
                const uint32_t acc = pow_sum_even(ex) + pow_sum_even(ey) + pow_sum_even(ox) + pow_sum_even(oy);
                const uint32_t lo = __reduce_add_sync(FULL, acc & 0xffffu);
                const uint32_t hi = __reduce_add_sync(FULL, acc >> 16);
                if (my_slot == c) { sl.lo = lo; sl.hi = hi; }
            }
        }
    } else {  // rare: exact 128-bit path, combination by combination
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
            const int k1 = c / K, k2 = c % K;
            uint2 ue = make_uint2(0u, 0u), uo = ue, ve = ue, vo = ue;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k == k1) { ue = U[k].he; uo = U[k].ho; }
                if (k == k2) { ve = V[k].he; vo = V[k].ho; }
            }
            uint64_t acc = 0;
            int nz = 0, zeros;
            acc_regs_word(__vmaxu2(ue.x, ve.x) | (__vmaxu2(uo.x, vo.x) << 8), acc, nz);
            acc_regs_word(__vmaxu2(ue.y, ve.y) | (__vmaxu2(uo.y, vo.y) << 8), acc, nz);
            unsigned __int128 tot = warp_total_units(acc, nz, zeros);
            const float S = units_to_f32(tot);
            if (my_slot == c) sl.Sx = S;
        }
    }
    uint32_t eq_r[3] = {0, 0, 0}, nz_r[3] = {0, 0, 0};
#pragma unroll
    for (int w = 0; w < EQW; ++w) eq_r[w] = __reduce_add_sync(FULL, eq_pack[w]);
#pragma unroll
    for (int w = 0; w < NZW; ++w) nz_r[w] = __reduce_add_sync(FULL, nz_pack[w]);
    if (mine) {
        const int nz_w = my_c / 3;
        sl.match = 128u - (((uint32_t)lk_select3(my_c >> 2, (int)eq_r[0], (int)eq_r[1], (int)eq_r[2]) >> (8 * (my_c & 3))) & 0xffu);
        sl.zeros = 256 - (int)(((uint32_t)lk_select3(nz_w, (int)nz_r[0], (int)nz_r[1], (int)nz_r[2]) >>
                                (10 * (my_c - 3 * nz_w))) & 0x3ffu);
    }
}

// batched tails + algebra + stores of the nb links [i0, i0 + nb); cards: the batch's cardinalities in shared memory,
// link j at cards[j * 2K .. j * 2K + 2K) (u's K then v's K)
template <int K>
__device__ __forceinline__ void lk_finish_batch(const LinkArgs &a, const LkSlot &sl, int nb, int64_t i0, int lane,
                                                const float *cards, float *stage) {
    constexpr int C = K * K;
    constexpr int F = K * (K + 2);
    // ---- batched tails: lane (j, c) finishes combination c of link j
    float my_inter = 0.f;
    if (lane < nb * C) {
        // lo < 2^21 and hi < 2^20 are exact in float32 and so is hi * 2^16: one rounding in the add gives the
        // correctly rounded total
        float S = sl.Sx;
        if (S < 0.f) S = __fmul_rn(__fadd_rn(__fmul_rn((float)sl.hi, 65536.f), (float)sl.lo), 3.7252902984619140625e-09f);
        my_inter = intersection_tail(a.h, sl.zeros, S, sl.match, 128);
    }
    if (a.inter && lane < nb * C) a.inter[i0 * C + lane] = my_inter;  // contiguous: consecutive links
    if (a.features) {
        // ---- batched algebra: lane j = link j of the batch
        const int jj = lane < nb ? lane : 0;
        float I[C];
#pragma unroll
        for (int c = 0; c < C; ++c) I[c] = __shfl_sync(FULL, my_inter, jj * C + c);
        float cu[K], cv[K], f[F];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            cu[k] = cards[jj * 2 * K + k];
            cv[k] = cards[jj * 2 * K + K + k];
        }
        feature_algebra<K>(I, cu, cv, f);
        knockout_and_floor<K>(f, a.flags);
        __syncwarp();
        if (lane < nb) {
#pragma unroll
            for (int x = 0; x < F; ++x) stage[lane * F + x] = f[x];
        }
        __syncwarp();
        for (int x = lane; x < nb * F; x += 32) a.features[i0 * F + x] = stage[x];
    }
}

// Shared memory of one warp (dynamic, carved by lk_smem_per_warp):
//   ids    2 x tile x 16 B   endpoints of the current and the next tile (cp.async double buffer)
//   recs   2K x 768 B        the records of the NEXT link, in flight while the current link is evaluated
//   cards  2 x B x 2K floats cardinalities of the endpoints of the current and the next batch (arrive with the records)
//   stage  B x F floats      feature transpose
//   bar    one mbarrier      transaction barrier of the record copies
template <int K> struct LkSmem {
    static constexpr int C = K * K, B = 32 / C, F = K * (K + 2);
    static constexpr int IDS = 2 * LK_TILE_MAX * 16, RECS = 2 * K * 768, CARDS = 2 * B * 2 * K * 4, STAGE = B * F * 4;
    static constexpr int BAR = 16;  // one mbarrier: completion of the bulk copies of a link's records
    static constexpr int PER_WARP = (IDS + RECS + CARDS + STAGE + BAR + 15) & ~15;
};

// One warp owns a tile of consecutive links.  Software pipeline per link (no registers spent on it):
//     wait for the link's records in shared memory -> LDS + prepare -> request the NEXT link's records (cp.async)
//     -> K^2 combinations from registers
// so the DRAM (or NVLink, for rows held by another GPU) round trip of link j + 1 overlaps the ~800 instructions of
// link j instead of stalling the warp in front of its first use (17 % of all stall samples before).  Along a run of
// links with the same source only v's records move.
template <int K>
__global__ void __launch_bounds__(256, 3) link_features_batched_kernel(const LinkArgs a, const int tile) {
    typedef LkSmem<K> SM;
    constexpr int C = SM::C, B = SM::B;
    extern __shared__ __align__(16) uint8_t lk_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *mine_smem = lk_smem + warp * SM::PER_WARP;
    longlong2 *ids_buf = reinterpret_cast<longlong2 *>(mine_smem);
    const uint32_t recs = smem_u32(mine_smem + SM::IDS);
    float *cards = reinterpret_cast<float *>(mine_smem + SM::IDS + SM::RECS);
    float *stage = reinterpret_cast<float *>(mine_smem + SM::IDS + SM::RECS + SM::CARDS);
    const uint32_t bar = smem_u32(mine_smem + ((SM::IDS + SM::RECS + SM::CARDS + SM::STAGE + 15) & ~15));
    if (lane == 0) {
        lk_mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t bar_parity = 0;
    const int gwarp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int n_warps = (int)((gridDim.x * blockDim.x) >> 5);
    const int n_tiles = (int)((a.n_links + tile - 1) / tile);  // the host keeps n_links / tile below 2^31
    const int my_j = lane / C;              // link of the batch this lane's tail slot belongs to (>= B: idle)
    const int my_c = lane - my_j * C;
    const bool sharded = a.n_ranks > 1;

    auto fetch_ids = [&](int t, int buf) {
        const int64_t base = (int64_t)t * tile;
        const int cnt = (int)min((int64_t)tile, a.n_links - base);
        for (int x = lane; x < cnt; x += 32)
            lk_cp_async16(smem_u32(ids_buf + buf * LK_TILE_MAX + x), reinterpret_cast<const longlong2 *>(a.links) + base + x);
    };
    // request the records (and cardinalities) of link `li` of the tile; u's only when it differs from `u_prev`
    auto request = [&](const longlong2 *ids, int li, int slot, int u_prev) {
        const int2 e = reinterpret_cast<const int2 *>(ids + li)[0];
        const int ow_u = sharded ? lk_owner(a, e.x) : 0, ow_v = sharded ? lk_owner(a, e.y) : 0;
        const bool have_u = sharded ? (__ldg(a.local_rows + e.x) != 0) : true;
        const bool have_v = sharded ? (__ldg(a.local_rows + e.y) != 0) : true;
        // ONE bulk copy (cp.async.bulk, SASS UBLKCP) per 768-byte record, issued by lane 0, completion counted in bytes on
        // the warp's mbarrier: a record held by another GPU crosses NVLink as a few large read requests instead of 48
        // sector-sized ones (per-lane 16-byte loads reached only ~260 GB/s of NVLink read rate)
        if (lane == 0) {
            lk_mbar_expect(bar, (uint32_t)((e.x != u_prev ? 2 * K : K) * 768));
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (e.x != u_prev) {
                    const uint8_t *ru = lk_table(a, k + 1, K, ow_u, have_u) + (int64_t)e.x * a.stride[k + 1];
                    lk_bulk_g2s(recs + k * 768, ru, 768, bar);
                }
                const uint8_t *rv = lk_table(a, k + 1, K, ow_v, have_v) + (int64_t)e.y * a.stride[k + 1];
                lk_bulk_g2s(recs + (K + k) * 768, rv, 768, bar);
            }
        }
        if (a.features && lane < 2 * K) {
            const int node = lane < K ? e.x : e.y;
            lk_cp_async4(smem_u32(cards + slot * 2 * K + lane), a.cards + (int64_t)node * a.cards_stride + (lane < K ? lane : lane - K));
        }
    };

    int buf = 0;
    if (gwarp < n_tiles) fetch_ids(gwarp, 0);
    for (int t = gwarp; t < n_tiles; t += n_warps) {
        lk_cp_async_wait();
        __syncwarp();
        if (t + n_warps < n_tiles) fetch_ids(t + n_warps, buf ^ 1);
        longlong2 *ids = ids_buf + buf * LK_TILE_MAX;
        buf ^= 1;
        const int cnt = (int)min((int64_t)tile, a.n_links - (int64_t)t * tile);
        // bounds-check the tile's endpoints once (the reference would raise IndexError); from here on they are int32
        for (int x = lane; x < cnt; x += 32) {
            const longlong2 e = ids[x];
            reinterpret_cast<int2 *>(ids + x)[0] = make_int2((int)checked_node(a, e.x), (int)checked_node(a, e.y));
        }
        __syncwarp();
        RowRegsB U[K];
        uint32_t big_u = 0;
        int u_cur = -1;
        request(ids, 0, 0, -1);
        int half = 0;  // which half of `cards` the running batch uses
        for (int b0 = 0; b0 < cnt; b0 += B, half ^= 1) {
            const int nb = min(B, cnt - b0);
            LkSlot sl;
            sl.lo = sl.hi = sl.match = 0u;
            sl.zeros = 0;
            sl.Sx = -1.f;
#pragma unroll 1
            for (int j = 0; j < nb; ++j) {
                const int u = reinterpret_cast<const int2 *>(ids + b0 + j)[0].x;
                lk_cp_async_wait();   // this link's cardinalities (and, long since, the next tile's ids)
                lk_mbar_wait(bar, bar_parity);  // ... and its records
                bar_parity ^= 1u;
                if (u != u_cur) {  // warp-uniform: a run of links with the same source keeps u's prepared records
                    u_cur = u;
                    big_u = 0;
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        prep_row_b(U[k], lk_lds_u4(recs + k * 768 + lane * 16), lk_lds_u2(recs + k * 768 + REC_MH + lane * 8), big_u);
                }
                // v's raw words leave shared memory first (18 registers), then the buffer is free and the NEXT link's
                // request goes out before anything is computed
                uint4 mv[K];
                uint2 hv[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    mv[k] = lk_lds_u4(recs + (K + k) * 768 + lane * 16);
                    hv[k] = lk_lds_u2(recs + (K + k) * 768 + REC_MH + lane * 8);
                }
                __syncwarp();  // every lane has its copy: the buffer may be refilled
                if (b0 + j + 1 < cnt) request(ids, b0 + j + 1, (j + 1 < nb) ? half * B + j + 1 : (half ^ 1) * B, u);
                RowRegsB V[K];
                uint32_t big = big_u;
#pragma unroll
                for (int k = 0; k < K; ++k) prep_row_b(V[k], mv[k], hv[k], big);
                lk_combine<K>(U, V, big, j, my_j, my_c, sl);
            }
            // (the cardinalities of the NEXT batch's first link are already landing in the other half of `cards`)
            lk_finish_batch<K>(a, sl, nb, (int64_t)t * tile + b0, lane, cards + half * B * 2 * K, stage);
        }
    }
}

// ---- TMA front end: links are processed in PAIRS; for each hop ONE gather4 (UTMALDG.2D.GATHER4) fetches
// the four records (uA, vA, uB, vB) of the pair into the warp's shared-memory stage, two stages deep, so the
// 12 records of the next pair are in flight while the current pair is being evaluated -- the load latency
// that the LDG front end exposes (one link per warp at a time, no room for a register double buffer) is
// hidden, and the 64-bit address arithmetic disappears (the TMA coordinates are the node ids).
struct LinkMaps {
    CUtensorMap m[3];
};

__device__ __forceinline__ void lk_gather4(uint32_t dst, const CUtensorMap *tmap, uint32_t bar, int r0, int r1, int r2,
                                           int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}

constexpr int LK_WARPS = 4;

template <int K>
__global__ void __launch_bounds__(LK_WARPS * 32) link_features_tma_kernel(const LinkArgs a,
                                                                          const __grid_constant__ LinkMaps maps) {
    constexpr uint32_t HOP_BYTES = 4 * 768;            // (uA, vA, uB, vB) of one hop
    constexpr uint32_t STAGE = K * HOP_BYTES;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = smem_u32(smem) + (uint32_t)warp * (2 * STAGE);
    const uint32_t bars = smem_u32(smem) + (uint32_t)LK_WARPS * (2 * STAGE) + (uint32_t)warp * 16;
    if (lane == 0) {
        lk_mbar_init(bars, 1);
        lk_mbar_init(bars + 8, 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int64_t n_pairs = (a.n_links + 1) >> 1;
    const int64_t gwarp = (int64_t)blockIdx.x * LK_WARPS + warp;
    const int64_t n_warps = (int64_t)gridDim.x * LK_WARPS;

    // lanes 0..3 hold the node ids (uA, vA, uB, vB) of a pair; lanes < 4K its cards; lane 0 issues the TMA
    auto fetch = [&](int64_t q, uint32_t stage, float &card) {
        int64_t li = 2 * q + (lane >> 1);               // link of this lane's id (lanes 0,1 -> A; 2,3 -> B)
        if (li >= a.n_links) li = a.n_links - 1;        // odd tail: B repeats A
        int id = 0;
        if (lane < 4) id = (int)checked_node(a, __ldg(a.links + 2 * li + (lane & 1)));
        const int r0 = __shfl_sync(FULL, id, 0), r1 = __shfl_sync(FULL, id, 1);
        const int r2 = __shfl_sync(FULL, id, 2), r3 = __shfl_sync(FULL, id, 3);
        const int node = __shfl_sync(FULL, id, (lane / K) & 3);
        card = (lane < 4 * K) ? __ldg(a.cards + (int64_t)node * a.cards_stride + (lane % K)) : 0.f;
        if (lane == 0) {
            const uint32_t bar = bars + 8 * stage;
            lk_mbar_expect(bar, STAGE);
#pragma unroll
            for (int k = 0; k < K; ++k) lk_gather4(base + stage * STAGE + k * HOP_BYTES, &maps.m[k], bar, r0, r1, r2, r3);
        }
    };

    float card_next = 0.f;
    uint32_t stage = 0, parity = 0;
    if (gwarp < n_pairs) fetch(gwarp, 0, card_next);
    for (int64_t q = gwarp; q < n_pairs; q += n_warps) {
        const float card_cur = card_next;
        if (q + n_warps < n_pairs) fetch(q + n_warps, stage ^ 1u, card_next);   // next pair into the other stage
        lk_mbar_wait(bars + 8 * stage, parity);
        const uint32_t rows = base + stage * STAGE;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int64_t i = 2 * q + half;
            if (i >= a.n_links) break;
            uint4 mu[K], mv[K];
            uint2 hu[K], hv[K];
            float cu[K], cv[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const uint32_t ru = rows + k * HOP_BYTES + (2 * half) * 768;
                const uint32_t rv = ru + 768;
                mu[k] = lk_lds_u4(ru + lane * 16);
                hu[k] = lk_lds_u2(ru + REC_MH + lane * 8);
                mv[k] = lk_lds_u4(rv + lane * 16);
                hv[k] = lk_lds_u2(rv + REC_MH + lane * 8);
                cu[k] = __shfl_sync(FULL, card_cur, (2 * half) * K + k);
                cv[k] = __shfl_sync(FULL, card_cur, (2 * half + 1) * K + k);
            }
            process_link<K>(a, i, mu, hu, mv, hv, cu, cv, lane);
        }
        __syncwarp();  // all lanes are done with this stage before it is refilled two iterations later
        if (stage == 1) parity ^= 1u;
        stage ^= 1u;
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
        const int64_t u = checked_node(a, __ldg(a.links + 2 * i)), v = checked_node(a, __ldg(a.links + 2 * i + 1));
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn lk_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// which front end.  Measured on B200 (R-MAT 24, K=3, 20 M links): LDG 27.8 ms, TMA pairs 37.2 ms -- the kernel is
// instruction-bound (~900 warp instructions per link) and the TMA stages cap residency at 12 warps / SM, so
// hiding the load latency does not pay for the lost issue slots.  LDG is the default; SS_B200_LINKS=tma opts in.
// which kernel evaluates unsharded links (measured on B200, R-MAT 24, 20 M links -- profiles/r02_link_features.txt):
//   K = 3: the per-link kernel (22.3 ms random) beats the batched one (24.0-25.3 ms) although the latter executes
//          16 % fewer instructions: both are bound by the ALU pipe (LOP3 / SHF / IADD3 / VIMNMX issue every other
//          cycle per scheduler) and the batched kernel's extra shared-memory traffic costs more than its shorter
//          tails save; on source-grouped lists the two are equal (22.2 vs 22.5 ms)
//   K <= 2: the batched kernel wins (13.9 vs 14.8 ms random, 11.7 vs 14.8 ms grouped)
// SS_B200_LINKS=ldg / batched / tma overrides; sharded tables are always read by the batched kernel (its software
// pipeline also hides the NVLink round trip of records held by another GPU).
static bool want_per_link_kernel(int K, bool sharded) {
    const char *e = getenv("SS_B200_LINKS");
    if (e && e[0] == 'l') return true;
    if (e && e[0] == 'b') return false;
    return K == 3 && !sharded;
}

static bool want_tma_links() {
    const char *e = getenv("SS_B200_LINKS");
    if (!(e && e[0] == 't')) return false;
    return lk_encode_fn() != nullptr;
}

template <int K>
static int launch_links(const LinkArgs &a, const int64_t *hop_rows, bool fast, cudaStream_t st) {
    int64_t blocks = (a.n_links + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 16;
    int grid = (int)(blocks < cap ? blocks : cap);
    if (fast && want_tma_links()) {
        LinkMaps maps;
        memset(&maps, 0, sizeof(maps));
        for (int k = 0; k < 3; ++k) {
            const int kk = k < K ? k + 1 : 1;  // unused maps repeat hop 1 (must still be valid descriptors)
            cuuint64_t dims[2] = {192, (cuuint64_t)hop_rows[kk]};
            cuuint64_t strides[1] = {(cuuint64_t)a.stride[kk]};
            cuuint32_t box[2] = {192, 1};
            cuuint32_t elem[2] = {1, 1};
            CUresult r = lk_encode_fn()(&maps.m[k], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t *>(a.hop[kk]), dims,
                                        strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled failed with CUresult %d for hop %d", (int)r, kk);
                return SS_ERR_CUDA;
            }
        }
        const size_t smem = (size_t)LK_WARPS * 2 * K * 4 * 768 + LK_WARPS * 16;
        auto k = link_features_tma_kernel<K>;
        SS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        SS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, LK_WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
        int64_t want = ((a.n_links + 1) / 2 + LK_WARPS - 1) / LK_WARPS;
        int64_t resident = (int64_t)per_sm * sm_count();
        k<<<(int)(want < resident ? want : resident), LK_WARPS * 32, smem, st>>>(a, maps);
        SS_LAUNCH_CHECK("link_features_tma_kernel");
    } else if (fast && want_per_link_kernel(K, a.n_ranks > 1)) {
        link_features_kernel<K><<<grid, 256, 0, st>>>(a);
        SS_LAUNCH_CHECK("link_features_kernel");
    } else if (fast) {
        // tile: a multiple of B = 32 / K^2, at most LK_TILE_MAX, small enough that every resident warp gets ~2 tiles
        constexpr int B = 32 / (K * K);
        const int64_t resident_warps = (int64_t)sm_count() * 3 * 8;
        int64_t tile = a.n_links / (2 * resident_warps);
        tile = tile / B * B;
        if (tile < B) tile = B;
        if (tile > LK_TILE_MAX) tile = LK_TILE_MAX;
        if (const char *e = getenv("SS_B200_LINK_TILE")) {  // tuning knob
            int64_t v = atoi(e) / B * B;
            if (v >= B && v <= LK_TILE_MAX) tile = v;
        }
        const int64_t n_tiles = (a.n_links + tile - 1) / tile;
        int64_t bl = (n_tiles + 7) / 8;
        int64_t cap_b = (int64_t)sm_count() * 3;
        auto k = link_features_batched_kernel<K>;
        const size_t smem = 8 * (size_t)LkSmem<K>::PER_WARP;
        SS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<(int)(bl < cap_b ? bl : cap_b), 256, smem, st>>>(a, (int)tile);
        SS_LAUNCH_CHECK("link_features_batched_kernel");
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
                     float *features_out, float *inter_out, int32_t *error_flag, ss_stream_t stream) {
    return ss_link_features_sharded(links, n_links, hops, max_hops, num_perm, hll_p, cards, cards_stride, hc, flags,
                                    features_out, inter_out, error_flag, nullptr, stream);
}

int ss_link_features_sharded(const int64_t *links, int64_t n_links, const ss_hop_view *hops, int max_hops, int num_perm,
                             int hll_p, const float *cards, int64_t cards_stride, const ss_hll_consts *hc, int flags,
                             float *features_out, float *inter_out, int32_t *error_flag, const ss_shard_view *shard,
                             ss_stream_t stream) {
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
    int64_t hop_rows[4] = {0, 0, 0, 0};
    a.links = links;
    a.n_links = n_links;
    for (int k = 1; k <= max_hops; ++k) {
        SS_REQUIRE(hops[k].records && ((uintptr_t)hops[k].records & 15) == 0, "hop %d records must be a 16-byte aligned device pointer", k);
        SS_REQUIRE(hops[k].row_stride >= s.bytes && (hops[k].row_stride & 15) == 0, "hop %d has a bad row stride", k);
        a.hop[k] = (const uint8_t *)hops[k].records;
        a.stride[k] = hops[k].row_stride;
        hop_rows[k] = hops[k].num_rows;
        SS_REQUIRE(hops[k].num_rows > 0 && hops[k].num_rows < (1ll << 31), "hop %d: num_rows must be in (0, 2^31)", k);
    }
    // inter-only calls read no cards: point at a valid dummy (the hll lc table) with stride 0
    a.cards = cards ? cards : hc->lc_table;
    a.cards_stride = cards ? cards_stride : 0;
    a.h = ss::to_dev(hc);
    a.flags = flags;
    a.features = features_out;
    a.inter = inter_out;
    a.s = s;
    a.err = error_flag;
    a.n_nodes = hop_rows[1];
    for (int k = 2; k <= max_hops; ++k)
        if (hop_rows[k] < a.n_nodes) a.n_nodes = hop_rows[k];
    const bool fast = (num_perm == 128 && hll_p == 8);
    a.n_ranks = 1;
    if (shard && shard->n_ranks > 1) {
        SS_REQUIRE(fast, "sharded tables need num_perm=128, hll_p=8");
        SS_REQUIRE(shard->n_ranks <= SS_MAX_PEERS + 1 && shard->rank >= 0 && shard->rank < shard->n_ranks,
                   "bad rank / world size in ss_shard_view");
        SS_REQUIRE(shard->local_rows, "ss_shard_view.local_rows is null");
        SS_REQUIRE(!ss::want_per_link_kernel(max_hops, true) && !ss::want_tma_links(), "sharded tables are read by the batched kernel only");
        a.n_ranks = shard->n_ranks;
        a.rank = shard->rank;
        a.last_hop_own_only = shard->last_hop_own_only;
        a.local_rows = shard->local_rows;
        for (int q = 0; q <= shard->n_ranks; ++q) a.bounds[q] = shard->bounds[q];
        SS_REQUIRE(a.bounds[0] == 0 && a.bounds[shard->n_ranks] <= a.n_nodes, "row blocks must start at 0 and end within the tables");
        for (int k = 1; k <= max_hops; ++k)
            for (int q = 0; q < shard->n_ranks; ++q) {
                a.peer_hop[k][q] = q == shard->rank ? a.hop[k] : (const uint8_t *)shard->peer_records[k][q];
                SS_REQUIRE(a.peer_hop[k][q] && ((uintptr_t)a.peer_hop[k][q] & 15) == 0, "peer table of hop %d / rank %d is null or misaligned", k, q);
            }
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (max_hops) {
        case 1: return ss::launch_links<1>(a, hop_rows, fast, st);
        case 2: return ss::launch_links<2>(a, hop_rows, fast, st);
        default: return ss::launch_links<3>(a, hop_rows, fast, st);
    }
}

}  // extern "C"
