// merge.cu -- K2: one hop of sketch propagation over the destination-keyed CSR.
//
// Replaces MinhashPropagation.forward + HllPropagation.forward
// (/root/reference/src/hashing.py:28-45; called per hop at :160-162) and the per-hop
// `cards[:, k-1] = hll_count(...)` (:163).  The reference materialises nnz x (8P + m) bytes of messages and
// scatter-maxes them; here every destination row pulls its in-neighbours' 768-byte records
// (min over the 128 uint32 MinHash slots, max over the 256 uint8 HLL registers).
//
// Work decomposition (default shape P=128, p=8): the colidx array is cut into equal RANGES of Q neighbours
// (nnz-split, like merge-path SpMV) and a warp streams one range at a time, whatever rows it covers:
//   * a row that lies completely inside the range is reduced in registers and written once (two coalesced
//     stores per lane: 16 B of MinHash + 8 B of HLL) together with its HLL++ cardinality;
//   * a row cut by a range boundary leaves a partial record in scratch slot 2w (row started before range w)
//     or 2w+1 (row continues after range w); the fix-up kernel folds the partials of each cut row.
// This makes the load perfectly balanced on power-law graphs (a hub of 700k neighbours is simply 342
// ranges) and lets the row pipeline run across row boundaries on low-degree graphs.  min/max are
// idempotent and commutative, so any split / order is bit-exact.
//
// Two streaming engines, same decomposition:
//   TMA: lanes issue one `cp.async.bulk` (SASS UBLKCP) per neighbour row into a per-warp shared-memory ring;
//        completion is tracked with one mbarrier per stage; the warp then reads the rows conflict-free.
//   LDG: each lane issues LDG.128 + LDG.64 per neighbour row, U rows in flight in registers.
// The generic kernel (any P, p) is row-per-warp with a column-chunk outer loop.
#include "common.cuh"

namespace ss {

constexpr int REC = 768;       // default record: 512 B MinHash + 256 B HLL
constexpr int REC_MH = 512;

struct MergeArgs {
    const int64_t *rowptr;
    const int32_t *colidx;
    int64_t n_rows;
    int64_t nnz;
    const uint8_t *in;
    int64_t in_stride;
    uint8_t *out;
    int64_t out_stride;
    uint8_t *scratch;  // 2 * n_ranges records
    int64_t n_ranges;
    int quantum;       // neighbours per range
    float *cards;
    int64_t cards_stride;
    HllDev h;
};

// HLL++ estimate of a row held as 8 registers per lane.  Not inlined: it is called once per output row
// from several places in the unrolled stream loop.
__device__ __noinline__ float row_cardinality(uint2 hl, int m, int T, int monotone, float threshold, float alpha_m2,
                                              float five_m, const float *lc, const float *est, const float *bias) {
    HllDev h;
    h.m = m; h.T = T; h.monotone = monotone; h.threshold = threshold; h.alpha_m2 = alpha_m2; h.five_m = five_m;
    h.lc = lc; h.est = est; h.bias = bias;
    uint64_t acc = 0;
    int nz = 0;
    acc_regs_word(hl.x, acc, nz);
    acc_regs_word(hl.y, acc, nz);
    int zeros;
    unsigned __int128 t = warp_total_units(acc, nz, zeros);
    return hll_estimate(h, zeros, t);
}

#define ROW_CARD(a, hl) \
    row_cardinality(hl, (a).h.m, (a).h.T, (a).h.monotone, (a).h.threshold, (a).h.alpha_m2, (a).h.five_m, (a).h.lc, \
                    (a).h.est, (a).h.bias)

// first row whose neighbour range contains position `pos` (0 <= pos < nnz): largest r with rowptr[r] <= pos
__device__ __forceinline__ int64_t row_of_position(const int64_t *__restrict__ rowptr, int64_t n_rows, int64_t pos) {
    int64_t lo = 0, hi = n_rows;  // invariant: rowptr[lo] <= pos < rowptr[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(rowptr + mid) <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

// state of the row currently being reduced by a warp
struct RowState {
    uint4 mh;
    uint2 hl;
    int64_t cur;      // row index
    int64_t rs, re;   // its neighbour range [rs, re)
    int64_t re_next;  // rowptr[cur + 2] (prefetched)
};

__device__ __forceinline__ void acc_reset(RowState &st) {
    st.mh = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    st.hl = make_uint2(0u, 0u);
}
__device__ __forceinline__ void acc_merge(RowState &st, const uint4 &m, const uint2 &h) {
    st.mh.x = min(st.mh.x, m.x);
    st.mh.y = min(st.mh.y, m.y);
    st.mh.z = min(st.mh.z, m.z);
    st.mh.w = min(st.mh.w, m.w);
    st.hl.x = __vmaxu4(st.hl.x, h.x);
    st.hl.y = __vmaxu4(st.hl.y, h.y);
}

// write the finished (or partial) current row; s/e = range bounds, w = range index
__device__ __forceinline__ void flush_row(const MergeArgs &a, const RowState &st, int64_t w, int64_t s, int64_t e,
                                          int lane) {
    if (st.rs == st.re) return;  // empty rows are zero-filled by the fix-up kernel
    const bool whole = st.rs >= s && st.re <= e;
    uint8_t *dst = whole ? a.out + st.cur * a.out_stride : a.scratch + (2 * w + (st.rs < s ? 0 : 1)) * (int64_t)REC;
    st_na_u4(dst + lane * 16, st.mh);
    st_na_u2(dst + REC_MH + lane * 8, st.hl);
    if (whole && a.cards) {
        float c = ROW_CARD(a, st.hl);
        if (lane == 0) a.cards[st.cur * a.cards_stride] = c;
    }
}

// move to the next row (the one starting at st.re)
__device__ __forceinline__ void advance_row(const MergeArgs &a, RowState &st) {
    st.cur += 1;
    st.rs = st.re;
    st.re = st.re_next;
    st.re_next = (st.cur + 2 <= a.n_rows) ? __ldg(a.rowptr + st.cur + 2) : st.re;
    acc_reset(st);
}

__device__ __forceinline__ void begin_range(const MergeArgs &a, RowState &st, int64_t s) {
    st.cur = row_of_position(a.rowptr, a.n_rows, s);
    st.rs = __ldg(a.rowptr + st.cur);
    st.re = __ldg(a.rowptr + st.cur + 1);
    st.re_next = (st.cur + 2 <= a.n_rows) ? __ldg(a.rowptr + st.cur + 2) : st.re;
    acc_reset(st);
}

// ------------------------------------------------------------------------------------------------
// LDG engine
// ------------------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(256) merge_ldg_kernel(const MergeArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint8_t *__restrict__ in = a.in;
    for (int64_t w = gwarp; w < a.n_ranges; w += n_warps) {
        const int64_t s = w * a.quantum;
        const int64_t e = min(s + (int64_t)a.quantum, a.nnz);
        if (s >= e) continue;  // nnz == 0
        RowState st;
        begin_range(a, st, s);
        int32_t next_ids = (s + lane < e) ? __ldg(a.colidx + s + lane) : 0;
        for (int64_t base = s; base < e; base += 32) {
            const int32_t ids = next_ids;
            next_ids = (base + 32 + lane < e) ? __ldg(a.colidx + base + 32 + lane) : 0;
            const int cnt = (int)min((int64_t)32, e - base);
            for (int j = 0; j < cnt; j += U) {
                uint4 m[U];
                uint2 h[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = __shfl_sync(FULL, ids, (j + u) & 31);
                    if (j + u < cnt) {
                        const uint8_t *row = in + (int64_t)c * a.in_stride;
                        m[u] = ld_nc_u4(row + lane * 16);
                        h[u] = ld_nc_u2(row + REC_MH + lane * 8);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (j + u < cnt) {
                        const int64_t pos = base + j + u;
                        while (pos == st.re) {
                            flush_row(a, st, w, s, e, lane);
                            advance_row(a, st);
                        }
                        acc_merge(st, m[u], h[u]);
                    }
                }
            }
        }
        flush_row(a, st, w, s, e, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// TMA engine: per-warp ring of S stages x G rows, one mbarrier per stage
// ------------------------------------------------------------------------------------------------
template <int G, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) merge_tma_kernel(const MergeArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint8_t *ring = smem + (size_t)warp * (S * G * REC);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)WARPS * S * G * REC) + warp * S;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < S; ++i) mbar_init(bars + i, 1);
        mbar_fence_init();
    }
    __syncwarp();

    const int64_t gwarp = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t n_warps = (int64_t)gridDim.x * WARPS;
    const uint8_t *__restrict__ in = a.in;
    uint32_t gcount = 0;  // groups consumed so far by this warp (ring position / phase bookkeeping)

    for (int64_t w = gwarp; w < a.n_ranges; w += n_warps) {
        const int64_t s = w * a.quantum;
        const int64_t e = min(s + (int64_t)a.quantum, a.nnz);
        if (s >= e) continue;
        const int n_groups = (int)((e - s + G - 1) / G);
        RowState st;

        // producer state: ids of the 32-position block the next group to issue falls into
        int issued = 0;
        int32_t p_ids = (s + lane < e) ? __ldg(a.colidx + s + lane) : 0;
        int32_t p_next = (s + 32 + lane < e) ? __ldg(a.colidx + s + 32 + lane) : 0;

        auto issue = [&]() {
            const int g = issued;
            const int off = g * G;  // position offset inside the range
            if (g > 0 && (off & 31) == 0) {
                p_ids = p_next;
                p_next = (s + off + 32 + lane < e) ? __ldg(a.colidx + s + off + 32 + lane) : 0;
            }
            const int cnt = (int)min((int64_t)G, e - s - off);
            const uint32_t slot = (gcount + (uint32_t)g) % S;
            const int c = __shfl_sync(FULL, p_ids, (off + (lane & (G - 1))) & 31);
            if (lane == 0) mbar_arrive_expect_tx(bars + slot, (uint32_t)cnt * REC);
            __syncwarp();
            if (lane < cnt) bulk_g2s(ring + ((size_t)slot * G + lane) * REC, in + (int64_t)c * a.in_stride, REC, bars + slot);
            issued = g + 1;
        };

        const int prologue = n_groups < S ? n_groups : S;
        for (int i = 0; i < prologue; ++i) issue();
        begin_range(a, st, s);

        for (int g = 0; g < n_groups; ++g) {
            const uint32_t gi = gcount + (uint32_t)g;
            const uint32_t slot = gi % S;
            mbar_wait(bars + slot, (gi / S) & 1u);
            const int cnt = (int)min((int64_t)G, e - s - (int64_t)g * G);
            const uint8_t *rows = ring + (size_t)slot * G * REC;
#pragma unroll
            for (int l = 0; l < G; ++l) {
                if (l < cnt) {
                    const int64_t pos = s + (int64_t)g * G + l;
                    const uint4 m = *reinterpret_cast<const uint4 *>(rows + l * REC + lane * 16);
                    const uint2 h = *reinterpret_cast<const uint2 *>(rows + l * REC + REC_MH + lane * 8);
                    while (pos == st.re) {
                        flush_row(a, st, w, s, e, lane);
                        advance_row(a, st);
                    }
                    acc_merge(st, m, h);
                }
            }
            __syncwarp();  // every lane has finished reading the slot before it is refilled
            if (issued < n_groups) issue();
        }
        flush_row(a, st, w, s, e, lane);
        gcount += (uint32_t)n_groups;
    }
}

// ------------------------------------------------------------------------------------------------
// fix-up: fold the partial records of rows cut by range boundaries; zero-fill rows with no in-edge
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_fixup_kernel(const MergeArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // (A) one warp per range: the range in which a cut row STARTS folds all of its pieces
    for (int64_t w = gwarp; w < a.n_ranges; w += n_warps) {
        const int64_t s = w * a.quantum;
        const int64_t e = min(s + (int64_t)a.quantum, a.nnz);
        if (s >= e) continue;
        const int64_t last = row_of_position(a.rowptr, a.n_rows, e - 1);
        const int64_t rs = __ldg(a.rowptr + last), re = __ldg(a.rowptr + last + 1);
        if (rs < s || re <= e) continue;  // started earlier (someone else folds) or not cut
        const int64_t w_end = (re - 1) / a.quantum;  // range holding the last neighbour
        RowState st;
        acc_reset(st);
        {
            const uint8_t *p = a.scratch + (2 * w + 1) * (int64_t)REC;
            acc_merge(st, ld_nc_u4(p + lane * 16), ld_nc_u2(p + REC_MH + lane * 8));
        }
        for (int64_t x = w + 1; x <= w_end; x += 4) {
            uint4 m[4];
            uint2 h[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (x + u <= w_end) {
                    const uint8_t *p = a.scratch + (2 * (x + u)) * (int64_t)REC;
                    m[u] = ld_nc_u4(p + lane * 16);
                    h[u] = ld_nc_u2(p + REC_MH + lane * 8);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (x + u <= w_end) acc_merge(st, m[u], h[u]);
        }
        uint8_t *dst = a.out + last * a.out_stride;
        st_na_u4(dst + lane * 16, st.mh);
        st_na_u2(dst + REC_MH + lane * 8, st.hl);
        if (a.cards) {
            float c = ROW_CARD(a, st.hl);
            if (lane == 0) a.cards[last * a.cards_stride] = c;
        }
    }
    // (B) rows without any in-edge: all-zero record (scatter-max fill value), cardinality of an empty sketch
    for (int64_t r0 = gwarp * 32; r0 < a.n_rows; r0 += n_warps * 32) {
        const int64_t r = r0 + lane;
        const bool empty = r < a.n_rows && __ldg(a.rowptr + r) == __ldg(a.rowptr + r + 1);
        unsigned mask = __ballot_sync(FULL, empty);
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            uint8_t *dst = a.out + (r0 + b) * a.out_stride;
            st_na_u4(dst + lane * 16, make_uint4(0u, 0u, 0u, 0u));
            st_na_u2(dst + REC_MH + lane * 8, make_uint2(0u, 0u));
            if (a.cards) {
                float c = ROW_CARD(a, make_uint2(0u, 0u));
                if (lane == 0) a.cards[(r0 + b) * a.cards_stride] = c;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// generic shape: one warp per row, columns processed in chunks of 32 lanes x 4 units
// ------------------------------------------------------------------------------------------------
struct GenericArgs {
    const int64_t *rowptr;
    const int32_t *colidx;
    int64_t n_rows;
    const uint8_t *in;
    int64_t in_stride;
    uint8_t *out;
    int64_t out_stride;
    float *cards;
    int64_t cards_stride;
    RecordShape s;
    HllDev h;
};

__global__ void __launch_bounds__(256) merge_generic_kernel(const GenericArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int UPL = 4;  // units per lane per chunk
    for (int64_t r = gwarp; r < a.n_rows; r += n_warps) {
        const int64_t b = a.rowptr[r], e = a.rowptr[r + 1];
        uint8_t *dst = a.out + r * a.out_stride;
        RegSum rsum;
        rsum.lo = rsum.hi = 0;
        rsum.zeros = 0;
        for (int u0 = 0; u0 < a.s.units; u0 += 32 * UPL) {
            uint2 acc[UPL];
            bool is_mh[UPL], live[UPL];
#pragma unroll
            for (int i = 0; i < UPL; ++i) {
                const int u = u0 + i * 32 + lane;
                live[i] = u < a.s.units;
                is_mh[i] = u < a.s.mh_units;
                acc[i] = (is_mh[i] && e > b) ? make_uint2(0xffffffffu, 0xffffffffu) : make_uint2(0u, 0u);
            }
            for (int64_t k = b; k < e; ++k) {
                const uint8_t *row = a.in + (int64_t)__ldg(a.colidx + k) * a.in_stride;
#pragma unroll
                for (int i = 0; i < UPL; ++i) {
                    if (live[i]) {
                        const uint2 v = ld_nc_u2(row + (int64_t)(u0 + i * 32 + lane) * 8);
                        acc[i] = is_mh[i] ? unit_min_u32(acc[i], v) : unit_max_u8(acc[i], v);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < UPL; ++i) {
                if (live[i]) {
                    st_na_u2(dst + (int64_t)(u0 + i * 32 + lane) * 8, acc[i]);
                    if (!is_mh[i]) {
                        regsum_add_word(rsum, acc[i].x);
                        regsum_add_word(rsum, acc[i].y);
                    }
                }
            }
        }
        if (a.cards) {
            int zeros;
            unsigned __int128 t = regsum_warp_total(rsum, zeros);
            float c = hll_estimate(a.h, zeros, t);
            if (lane == 0) a.cards[r * a.cards_stride] = c;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// operator forms on the reference layouts (int64 MinHash / int8 HLL tensors)
// ------------------------------------------------------------------------------------------------
template <typename T, bool IS_MIN>
__global__ void __launch_bounds__(256) prop_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                    int64_t n_rows, const T *__restrict__ x, T *__restrict__ out,
                                                    int64_t width) {
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = gwarp; r < n_rows; r += n_warps) {
        const int64_t b = rowptr[r], e = rowptr[r + 1];
        for (int64_t c0 = 0; c0 < width; c0 += 128) {
            T acc[4];
            bool have = false;
            for (int64_t k = b; k < e; ++k) {
                const int64_t src = colidx ? (int64_t)__ldg(colidx + k) : k;
                const T *row = x + src * width;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t c = c0 + i * 32 + lane;
                    if (c < width) {
                        const T v = __ldg(row + c);
                        acc[i] = !have ? v : (IS_MIN ? (v < acc[i] ? v : acc[i]) : (v > acc[i] ? v : acc[i]));
                    }
                }
                have = true;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t c = c0 + i * 32 + lane;
                if (c < width) out[r * width + c] = have ? acc[i] : (T)0;
            }
        }
    }
}

// range size: a power of two in [32, 2048] giving every resident warp several ranges
static int pick_quantum(int64_t nnz) {
    const int64_t want_ranges = (int64_t)sm_count() * 32 * 8;
    int64_t q = nnz / (want_ranges > 0 ? want_ranges : 1);
    int p2 = 32;
    while (p2 * 2 <= q && p2 < 2048) p2 *= 2;
    return p2;
}

static int64_t n_ranges_for(int64_t nnz, int quantum) {
    int64_t n = (nnz + quantum - 1) / quantum;
    return n < 1 ? 1 : n;
}

template <typename K>
static int resident_grid(K kernel, int block, size_t smem, int *out_grid) {
    int per_sm = 0;
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
    if (err != cudaSuccess || per_sm < 1) {
        set_error("kernel cannot be resident (block=%d smem=%zu): %s", block, smem, cudaGetErrorString(err));
        return SS_ERR_CUDA;
    }
    *out_grid = per_sm * sm_count();
    return SS_OK;
}

constexpr int TMA_G = 4, TMA_S = 4, TMA_WARPS = 8;
constexpr size_t TMA_SMEM = (size_t)TMA_WARPS * TMA_S * TMA_G * REC + TMA_WARPS * TMA_S * 8;

}  // namespace ss

extern "C" {

int64_t ss_merge_workspace_bytes(int64_t nnz, int num_perm, int hll_p) {
    ss::RecordShape s;
    if (nnz < 0 || !ss::make_shape(num_perm, hll_p, &s)) {
        ss::set_error("bad arguments to ss_merge_workspace_bytes");
        return SS_ERR_INVALID;
    }
    if (!(num_perm == 128 && hll_p == 8)) return 16;  // generic kernel needs no scratch
    return 2 * ss::n_ranges_for(nnz, ss::pick_quantum(nnz)) * (int64_t)ss::REC;
}

int ss_khop_merge(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, int64_t nnz, const void *rec_in,
                  int64_t in_stride, void *rec_out, int64_t out_stride, int num_perm, int hll_p, void *workspace,
                  int64_t workspace_bytes, float *cards_out, int64_t cards_stride, const ss_hll_consts *hc, int variant,
                  ss_stream_t stream) {
    ss::RecordShape s;
    SS_REQUIRE(ss::make_shape(num_perm, hll_p, &s), "unsupported sketch shape num_perm=%d hll_p=%d", num_perm, hll_p);
    SS_REQUIRE(n_rows >= 0 && nnz >= 0, "negative size passed to ss_khop_merge");
    if (n_rows == 0) return SS_OK;
    SS_REQUIRE(n_rows < (1ll << 31), "at most 2^31-1 rows per call");
    SS_REQUIRE(rowptr && rec_in && rec_out, "null pointer passed to ss_khop_merge");
    SS_REQUIRE(nnz == 0 || colidx, "colidx is null");
    SS_REQUIRE((((uintptr_t)rec_in | (uintptr_t)rec_out) & 15) == 0, "record tables must be 16-byte aligned");
    SS_REQUIRE(in_stride >= s.bytes && out_stride >= s.bytes && ((in_stride | out_stride) & 15) == 0,
               "record strides must be >= %d and multiples of 16", s.bytes);
    ss::HllDev hd;
    memset(&hd, 0, sizeof(hd));
    if (cards_out) {
        int rc = ss::check_hll_consts(hc, hll_p);
        if (rc != SS_OK) return rc;
        hd = ss::to_dev(hc);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool fast_shape = (num_perm == 128 && hll_p == 8);
    if (variant == SS_MERGE_AUTO) variant = fast_shape ? SS_MERGE_TMA : SS_MERGE_GENERIC;
    SS_REQUIRE(variant == SS_MERGE_GENERIC || fast_shape, "TMA/LDG merge kernels need num_perm=128, hll_p=8");

    if (variant == SS_MERGE_GENERIC) {
        ss::GenericArgs g;
        g.rowptr = rowptr; g.colidx = colidx; g.n_rows = n_rows;
        g.in = (const uint8_t *)rec_in; g.in_stride = in_stride;
        g.out = (uint8_t *)rec_out; g.out_stride = out_stride;
        g.cards = cards_out; g.cards_stride = cards_stride; g.s = s; g.h = hd;
        int64_t blocks = (n_rows + 7) / 8;
        int64_t cap = (int64_t)ss::sm_count() * 8;
        ss::merge_generic_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(g);
        SS_LAUNCH_CHECK("merge_generic_kernel");
        return SS_OK;
    }

    ss::MergeArgs a;
    a.rowptr = rowptr; a.colidx = colidx; a.n_rows = n_rows; a.nnz = nnz;
    a.in = (const uint8_t *)rec_in; a.in_stride = in_stride;
    a.out = (uint8_t *)rec_out; a.out_stride = out_stride;
    a.quantum = ss::pick_quantum(nnz);
    a.n_ranges = ss::n_ranges_for(nnz, a.quantum);
    a.scratch = (uint8_t *)workspace;
    a.cards = cards_out; a.cards_stride = cards_stride; a.h = hd;
    const int64_t need = 2 * a.n_ranges * (int64_t)ss::REC;
    SS_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0, "merge workspace must be a 16-byte aligned device buffer");
    if (workspace_bytes < need) {
        ss::set_error("merge workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
        return SS_ERR_WORKSPACE;
    }
    int grid = 0, rc;
    if (nnz > 0) {
        if (variant == SS_MERGE_TMA) {
            auto k = ss::merge_tma_kernel<ss::TMA_G, ss::TMA_S, ss::TMA_WARPS>;
            SS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss::TMA_SMEM));
            if ((rc = ss::resident_grid(k, ss::TMA_WARPS * 32, ss::TMA_SMEM, &grid)) != SS_OK) return rc;
            int64_t blocks = (a.n_ranges + ss::TMA_WARPS - 1) / ss::TMA_WARPS;
            if (blocks < grid) grid = (int)blocks;
            k<<<grid, ss::TMA_WARPS * 32, ss::TMA_SMEM, st>>>(a);
            SS_LAUNCH_CHECK("merge_tma_kernel");
        } else if (variant == SS_MERGE_LDG) {
            auto k = ss::merge_ldg_kernel<8>;
            if ((rc = ss::resident_grid(k, 256, 0, &grid)) != SS_OK) return rc;
            int64_t blocks = (a.n_ranges + 7) / 8;
            if (blocks < grid) grid = (int)blocks;
            k<<<grid, 256, 0, st>>>(a);
            SS_LAUNCH_CHECK("merge_ldg_kernel");
        } else {
            ss::set_error("unknown merge variant %d", variant);
            return SS_ERR_INVALID;
        }
    }
    {
        int64_t warps = a.n_ranges > (n_rows + 31) / 32 ? a.n_ranges : (n_rows + 31) / 32;
        int64_t blocks = (warps + 7) / 8;
        int64_t cap = (int64_t)ss::sm_count() * 8;
        ss::merge_fixup_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(a);
        SS_LAUNCH_CHECK("merge_fixup_kernel");
    }
    return SS_OK;
}

static int prop_grid(int64_t n_rows) {
    int64_t blocks = (n_rows + 7) / 8;
    int64_t cap = (int64_t)ss::sm_count() * 8;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

int ss_prop_min_i64(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, const int64_t *x, int64_t *out,
                    int64_t width, ss_stream_t stream) {
    SS_REQUIRE(n_rows >= 0 && width >= 0, "negative size passed to ss_prop_min_i64");
    if (n_rows == 0 || width == 0) return SS_OK;
    SS_REQUIRE(rowptr && x && out, "null pointer passed to ss_prop_min_i64");
    ss::prop_kernel<int64_t, true><<<prop_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, n_rows, x, out, width);
    SS_LAUNCH_CHECK("prop_kernel<int64,min>");
    return SS_OK;
}

int ss_prop_max_i8(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, const int8_t *x, int8_t *out,
                   int64_t width, ss_stream_t stream) {
    SS_REQUIRE(n_rows >= 0 && width >= 0, "negative size passed to ss_prop_max_i8");
    if (n_rows == 0 || width == 0) return SS_OK;
    SS_REQUIRE(rowptr && x && out, "null pointer passed to ss_prop_max_i8");
    ss::prop_kernel<int8_t, false><<<prop_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, n_rows, x, out, width);
    SS_LAUNCH_CHECK("prop_kernel<int8,max>");
    return SS_OK;
}

}  // extern "C"
