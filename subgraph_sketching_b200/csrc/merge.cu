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
// Streaming engines, same decomposition:
//   TMA : lane 0 issues ONE `cp.async.bulk.tensor.2d ... tile::gather4` (SASS UTMALDG.2D.GATHER4) per group of 4
//         neighbour rows into a per-warp shared-memory ring; completion is tracked with one mbarrier per stage;
//         the warp then reads the rows conflict-free.  The default.
//   BULK: same ring, four 1-D `cp.async.bulk` copies (SASS UBLKCP) per group.
//   LDG : each lane issues LDG.128 + LDG.64 per neighbour row, U rows in flight in registers.
// The generic kernel (any P, p) is row-per-warp with a column-chunk outer loop.
#include <stdlib.h>
#include <string.h>

#include <cuda.h>

#include "common.cuh"

namespace ss {

constexpr int REC = 768;       // default record: 512 B MinHash + 256 B HLL

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
    // fused exchange (multi-GPU): peer copies of `out` / `cards`, mapped into this process (NVLink P2P);
    // every finished row is also stored there, so the next-hop table is replicated when the launch ends
    int n_peers;
    uint8_t *peer_out[SS_MAX_PEERS];
    float *peer_cards[SS_MAX_PEERS];
    // NVSwitch multicast alternative: ONE multimem.st per store lands in every GPU's copy (own copy included)
    uint8_t *mc_out;
    float *mc_cards;
    // halo push (optional): bit p of peer_mask[row] says whether peer p ever reads output row `row` (it owns a
    // destination with that row as in-neighbour, or a link endpoint); rows nobody else reads stay local
    const uint8_t *peer_mask;
    // memoised pipelines: skip the launch when *guard == 0 (see guarded_skip)
    const int *guard;
    // BLOCKED launches (hop 1 under the ingest stream): the row block [row_begin, row_end) and its neighbour positions
    // [pos_begin, pos_end) are read from DEVICE memory at kernel start -- the host enqueues the launch before it knows
    // them.  rowptr / colidx / out / cards / scratch are then the WHOLE graph's arrays (absolute positions).
    const long long *block;
};

// the part of the problem a launch works on: everything (host-sized) or a device-described row block
struct MergeSpan {
    int64_t row0, n_rows;   // rows [row0, row0 + n_rows)
    int64_t pos0, pos1;     // neighbour positions [pos0, pos1)
    int64_t win0, n_ranges; // first quantum-aligned window and number of windows touching [pos0, pos1)
};
template <bool BLOCKED>
__device__ __forceinline__ MergeSpan merge_span(const MergeArgs &a) {
    MergeSpan sp;
    if (BLOCKED) {
        sp.row0 = __ldg(a.block + 0);
        sp.n_rows = __ldg(a.block + 1) - sp.row0;
        sp.pos0 = __ldg(a.block + 2);
        sp.pos1 = __ldg(a.block + 3);
        sp.win0 = sp.pos0 / a.quantum;
        sp.n_ranges = sp.pos1 > sp.pos0 ? (sp.pos1 - 1) / a.quantum - sp.win0 + 1 : 0;
    } else {
        sp.row0 = 0; sp.n_rows = a.n_rows; sp.pos0 = 0; sp.pos1 = a.nnz; sp.win0 = 0; sp.n_ranges = a.n_ranges;
    }
    return sp;
}
// arguments as the row-level helpers see them: in a blocked launch rowptr / out / cards are rebased to the block's
// first row (rowptr VALUES stay absolute positions; range starts are absolute too)
template <bool BLOCKED>
__device__ __forceinline__ MergeArgs merge_rebased(const MergeArgs &a, const MergeSpan &sp) {
    MergeArgs b = a;
    if (BLOCKED) {
        b.rowptr = a.rowptr + sp.row0;
        b.n_rows = sp.n_rows;
        b.out = a.out + sp.row0 * a.out_stride;
        if (a.cards) b.cards = a.cards + sp.row0 * a.cards_stride;
    }
    return b;
}

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

__device__ __forceinline__ void mc_st_u4(void *p, const uint4 &v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}
__device__ __forceinline__ void mc_st_u2(void *p, const uint2 &v) {
    asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y))
                 : "memory");
}
__device__ __forceinline__ void mc_st_f32(float *p, float v) {
    asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- row layouts ---------------------------------------------------------------------------------------
// What one row of the gathered table holds; a lane owns MW MinHash words and HW HLL words of it:
//   FULL (768 B)  128 x u32 MinHash + 256 x u8 registers   lane: 16 B + 8 B   the hop tables of build_hash_tables
//   MH   (512 B)  128 x u32 MinHash                        lane: 16 B         MinhashPropagation called singly
//   HLL  (256 B)  256 x u8 registers                       lane:        8 B   HllPropagation called singly: the
//                                                                             reference's int8 [N, 256] tensor AS IS
//   HALF (384 B)  64 x u32 MinHash + 128 x u8 registers    lane:  8 B + 4 B   one column half of a record
//                                                                             (column-sharded multi-GPU variants)
constexpr int LAY_FULL = 0, LAY_MH = 1, LAY_HLL = 2, LAY_HALF = 3;
template <int MODE> struct Lay;
template <int MW_, int HW_, uint32_t BIAS_> struct LayBase {
    static constexpr int MW = MW_, HW = HW_;
    static constexpr int MH_BYTES = MW_ * 128, BYTES = (MW_ + HW_) * 128;
    static constexpr uint32_t BIAS = BIAS_;  // see acc_merge
};
template <> struct Lay<LAY_FULL> : LayBase<4, 2, 0u> {};
template <> struct Lay<LAY_MH> : LayBase<4, 0, 0u> {};
template <> struct Lay<LAY_HLL> : LayBase<0, 2, 0x80808080u> {};
template <> struct Lay<LAY_HALF> : LayBase<2, 1, 0u> {};

// state of the row currently being reduced by a warp.  Positions are relative to the range start s
// (32-bit: one compare per neighbour), clamped so that "started before" / "continues after" stay visible.
// HLL registers are accumulated in two planes (even / odd bytes, each in its own 16-bit lane) so that the
// register-wise max is the native 16x2 max (VIMNMX.U16x2 / VIMNMX3) -- a 4 x uint8 max does not exist in
// hardware and costs 7 instructions when emulated.
template <int MODE>
struct RowStateT {
    uint32_t mh[Lay<MODE>::MW > 0 ? Lay<MODE>::MW : 1];
    uint32_t he[Lay<MODE>::HW > 0 ? Lay<MODE>::HW : 1], ho[Lay<MODE>::HW > 0 ? Lay<MODE>::HW : 1];
    int cur;        // row index
    int rs, re;     // neighbour range of the row relative to s: rs = -1 if it started before the range
    int re_next;    // relative end of row cur + 1 (prefetched)
};
// one gathered (or finished) row as a lane sees it
template <int MODE>
struct RowVec {
    uint32_t m[Lay<MODE>::MW > 0 ? Lay<MODE>::MW : 1];
    uint32_t h[Lay<MODE>::HW > 0 ? Lay<MODE>::HW : 1];
};
constexpr int REL_CAP = 1 << 30;
constexpr uint32_t EVEN = 0x00ff00ffu, ODD = 0xff00ff00u;

__device__ __forceinline__ int rel_pos(int64_t abs_pos, int64_t s) {
    const int64_t d = abs_pos - s;
    return d < 0 ? -1 : (d > REL_CAP ? REL_CAP : (int)d);
}
template <int MODE>
__device__ __forceinline__ void acc_reset(RowStateT<MODE> &st) {
#pragma unroll
    for (int i = 0; i < Lay<MODE>::MW; ++i) st.mh[i] = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < Lay<MODE>::HW; ++i) { st.he[i] = 0u; st.ho[i] = 0u; }
}
// The HLL-only layout runs on the reference's int8 tensor as it is, and HllPropagation is a SIGNED max there
// (hashing.py:38-45): bytes are biased by 0x80 inside the accumulator (signed order == unsigned order of x ^ 0x80;
// the XOR folds into the plane-extracting LOP3), so any int8 content is merged exactly.  Records hold registers
// in [0, 64] and need no bias.

template <int MODE>
__device__ __forceinline__ void acc_merge(RowStateT<MODE> &st, const RowVec<MODE> &r) {
#pragma unroll
    for (int i = 0; i < Lay<MODE>::MW; ++i) st.mh[i] = min(st.mh[i], r.m[i]);
#pragma unroll
    for (int i = 0; i < Lay<MODE>::HW; ++i) {
        st.he[i] = __vmaxu2(st.he[i], (r.h[i] ^ Lay<MODE>::BIAS) & EVEN);
        st.ho[i] = __vmaxu2(st.ho[i], (r.h[i] ^ Lay<MODE>::BIAS) & ODD);
    }
}
// two neighbour rows at once: three-input min / max (VIMNMX3)
template <int MODE>
__device__ __forceinline__ void acc_merge2(RowStateT<MODE> &st, const RowVec<MODE> &r1, const RowVec<MODE> &r2) {
#pragma unroll
    for (int i = 0; i < Lay<MODE>::MW; ++i) st.mh[i] = __vimin3_u32(st.mh[i], r1.m[i], r2.m[i]);
#pragma unroll
    for (int i = 0; i < Lay<MODE>::HW; ++i) {
        st.he[i] = __vimax3_u16x2(st.he[i], (r1.h[i] ^ Lay<MODE>::BIAS) & EVEN, (r2.h[i] ^ Lay<MODE>::BIAS) & EVEN);
        st.ho[i] = __vimax3_u16x2(st.ho[i], (r1.h[i] ^ Lay<MODE>::BIAS) & ODD, (r2.h[i] ^ Lay<MODE>::BIAS) & ODD);
    }
}
template <int MODE>
__device__ __forceinline__ RowVec<MODE> acc_row(const RowStateT<MODE> &st) {
    RowVec<MODE> r;
#pragma unroll
    for (int i = 0; i < Lay<MODE>::MW; ++i) r.m[i] = st.mh[i];
#pragma unroll
    for (int i = 0; i < Lay<MODE>::HW; ++i) r.h[i] = (st.he[i] | st.ho[i]) ^ Lay<MODE>::BIAS;
    return r;
}

// vector accesses of a lane's slice of a row: MinHash words at lane * 4 MW, HLL words at mh_bytes + lane * 4 HW
template <int W>
__device__ __forceinline__ void ldg_words(const uint8_t *p, uint32_t *w) {
    if (W == 4) { const uint4 v = ld_nc_u4(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
    if (W == 2) { const uint2 v = ld_nc_u2(p); w[0] = v.x; w[1] = v.y; }
    if (W == 1) { w[0] = __ldg(reinterpret_cast<const uint32_t *>(p)); }
}
template <int W>
__device__ __forceinline__ void stg_words(uint8_t *p, const uint32_t *w) {
    if (W == 4) st_na_u4(p, make_uint4(w[0], w[1], w[2], w[3]));
    if (W == 2) st_na_u2(p, make_uint2(w[0], w[1]));
    if (W == 1) *reinterpret_cast<uint32_t *>(p) = w[0];
}
template <int W>
__device__ __forceinline__ void mc_words(uint8_t *p, const uint32_t *w) {
    if (W == 4) mc_st_u4(p, make_uint4(w[0], w[1], w[2], w[3]));
    if (W == 2) mc_st_u2(p, make_uint2(w[0], w[1]));
    if (W == 1) mc_st_f32(reinterpret_cast<float *>(p), __uint_as_float(w[0]));
}
template <int MODE>
__device__ __forceinline__ RowVec<MODE> ldg_row(const uint8_t *row, int lane) {
    RowVec<MODE> r;
    ldg_words<Lay<MODE>::MW>(row + lane * (4 * Lay<MODE>::MW), r.m);
    ldg_words<Lay<MODE>::HW>(row + Lay<MODE>::MH_BYTES + lane * (4 * Lay<MODE>::HW), r.h);
    return r;
}
template <int MODE>
__device__ __forceinline__ void stg_row(uint8_t *row, const RowVec<MODE> &r, int lane) {
    stg_words<Lay<MODE>::MW>(row + lane * (4 * Lay<MODE>::MW), r.m);
    stg_words<Lay<MODE>::HW>(row + Lay<MODE>::MH_BYTES + lane * (4 * Lay<MODE>::HW), r.h);
}

// final store of output row `row`: local table, its cardinality, and the same to every peer table
// (one store per peer over NVLink, or one multicast store that the NVSwitch replicates to every GPU)
template <int MODE>
__device__ __forceinline__ void store_row(const MergeArgs &a, int64_t row, const RowVec<MODE> &r, int lane) {
    const int64_t off = row * a.out_stride;
    if (a.mc_out) {
        mc_words<Lay<MODE>::MW>(a.mc_out + off + lane * (4 * Lay<MODE>::MW), r.m);
        mc_words<Lay<MODE>::HW>(a.mc_out + off + Lay<MODE>::MH_BYTES + lane * (4 * Lay<MODE>::HW), r.h);
    } else {
        stg_row<MODE>(a.out + off, r, lane);
        for (int p = 0; p < a.n_peers; ++p) {
            if (a.peer_mask && !((__ldg(a.peer_mask + row) >> p) & 1)) continue;  // halo push: peer p never reads this row
            stg_row<MODE>(a.peer_out[p] + off, r, lane);
        }
    }
    if (MODE == LAY_FULL || MODE == LAY_HLL) {
        if (a.cards) {
            float c = ROW_CARD(a, make_uint2(r.h[0], r.h[Lay<MODE>::HW > 1 ? 1 : 0]));
            if (lane == 0) {
                if (a.mc_cards) {
                    mc_st_f32(a.mc_cards + row * a.cards_stride, c);
                } else {
                    a.cards[row * a.cards_stride] = c;
                    for (int p = 0; p < a.n_peers; ++p) a.peer_cards[p][row * a.cards_stride] = c;
                }
            }
        }
    }
}

// first row whose neighbour range contains position `pos` (0 <= pos < nnz): largest r with rowptr[r] <= pos
__device__ __forceinline__ int64_t row_of_position(const int64_t *__restrict__ rowptr, int64_t n_rows, int64_t pos) {
    int64_t lo = 0, hi = n_rows;  // invariant: rowptr[lo] <= pos < rowptr[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(rowptr + mid) <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

// write the finished (or partial) current row; n_pos = length of the range, w = range index
template <int MODE>
__device__ __forceinline__ void flush_row(const MergeArgs &a, const RowStateT<MODE> &st, int64_t w, int n_pos, int lane) {
    if (st.rs == st.re) return;  // empty rows are zero-filled by the fix-up kernel
    const RowVec<MODE> r = acc_row<MODE>(st);
    if (st.rs >= 0 && st.re <= n_pos) {
        store_row<MODE>(a, st.cur, r, lane);
    } else {  // cut by a range boundary: partial record for the fix-up kernel
        stg_row<MODE>(a.scratch + (2 * w + (st.rs < 0 ? 0 : 1)) * (int64_t)Lay<MODE>::BYTES, r, lane);
    }
}

// move to the next row (the one starting at st.re)
template <int MODE>
__device__ __forceinline__ void advance_row(const MergeArgs &a, RowStateT<MODE> &st, int64_t s) {
    st.cur += 1;
    st.rs = st.re;
    st.re = st.re_next;
    st.re_next = ((int64_t)st.cur + 2 <= a.n_rows) ? rel_pos(__ldg(a.rowptr + st.cur + 2), s) : st.re;
    acc_reset<MODE>(st);
}

template <int MODE>
__device__ __forceinline__ void begin_range(const MergeArgs &a, RowStateT<MODE> &st, int64_t s) {
    st.cur = (int)row_of_position(a.rowptr, a.n_rows, s);
    st.rs = rel_pos(__ldg(a.rowptr + st.cur), s);  // rowptr[cur] <= s: 0 if the row starts here, else -1
    st.re = rel_pos(__ldg(a.rowptr + st.cur + 1), s);
    st.re_next = ((int64_t)st.cur + 2 <= a.n_rows) ? rel_pos(__ldg(a.rowptr + st.cur + 2), s) : st.re;
    acc_reset<MODE>(st);
}

// kernels that belong to a memoised pipeline (the ELPH per-batch path) are launched with a GUARD: a device word that
// says whether their cached result is still valid; they return at once when it is (no host synchronisation needed
// to decide)
__device__ __forceinline__ bool guarded_skip(const int *guard) { return guard && *reinterpret_cast<const volatile int *>(guard) == 0; }

// ------------------------------------------------------------------------------------------------
// LDG engine
// ------------------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(256) merge_ldg_kernel(const MergeArgs a) {
    constexpr int MODE = LAY_FULL;
    if (guarded_skip(a.guard)) return;
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const uint8_t *__restrict__ in = a.in;
    for (int64_t w = gwarp; w < a.n_ranges; w += n_warps) {
        const int64_t s = w * a.quantum;
        const int n_pos = (int)min((int64_t)a.quantum, a.nnz - s);
        if (n_pos <= 0) continue;  // nnz == 0
        RowStateT<MODE> st;
        begin_range<MODE>(a, st, s);
        const int32_t *__restrict__ ids_ptr = a.colidx + s;
        int32_t next_ids = (lane < n_pos) ? __ldg(ids_ptr + lane) : 0;
        for (int base = 0; base < n_pos; base += 32) {
            const int32_t ids = next_ids;
            next_ids = (base + 32 + lane < n_pos) ? __ldg(ids_ptr + base + 32 + lane) : 0;
            const int cnt = min(32, n_pos - base);
            for (int j = 0; j < cnt; j += U) {
                RowVec<MODE> rows[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = __shfl_sync(FULL, ids, (j + u) & 31);
                    if (j + u < cnt) rows[u] = ldg_row<MODE>(in + (int64_t)c * a.in_stride, lane);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (j + u < cnt) {
                        const int pos = base + j + u;
                        while (pos == st.re) {
                            flush_row<MODE>(a, st, w, n_pos, lane);
                            advance_row<MODE>(a, st, s);
                        }
                        acc_merge<MODE>(st, rows[u]);
                    }
                }
            }
        }
        flush_row<MODE>(a, st, w, n_pos, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// TMA engine: per-warp ring of S stages x 4 rows in shared memory, one mbarrier per stage.
// Lane 0 is the producer: it reads the 4 neighbour ids of a group with one 16-byte load (prefetched one
// group ahead), posts the expected byte count on the stage's mbarrier and issues
//   GATHER4: ONE `cp.async.bulk.tensor.2d ... tile::gather4` (SASS UTMALDG) -- Blackwell's row-gather TMA:
//            the 4 row indices go straight into the instruction, no address arithmetic; the previous-hop
//            table is described by a 2-D tensor map [N rows x 192 uint32] built per call on the host;
//   else   : four 1-D `cp.async.bulk` copies (SASS UBLKCP), one per row.
// All 32 lanes then wait on the mbarrier and reduce the rows out of shared memory (conflict-free: lane l
// reads bytes [16 l, 16 l + 16) of the MinHash part and [8 l, 8 l + 8) of the HLL part), two rows per step
// where the output row does not change (three-input min/max).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init32(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait32(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bulk_g2s32(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *tmap, uint32_t bar, int col, int r0, int r1,
                                            int r2, int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
    return r;
}

__device__ __forceinline__ uint32_t lds_u1(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
template <int W>
__device__ __forceinline__ void lds_words(uint32_t addr, uint32_t *w) {
    if (W == 4) { const uint4 v = lds_u4(addr); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
    if (W == 2) { const uint2 v = lds_u2(addr); w[0] = v.x; w[1] = v.y; }
    if (W == 1) { w[0] = lds_u1(addr); }
}

template <int S, int WARPS, int MIN_CTAS, bool GATHER4, int MODE, bool BLOCKED = false>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS) merge_tma_kernel(const MergeArgs a_in,
                                                                           const __grid_constant__ CUtensorMap tmap) {
    constexpr int G = 4;
    constexpr int RB = Lay<MODE>::BYTES;
    constexpr int MW = Lay<MODE>::MW, HW = Lay<MODE>::HW;
    constexpr uint32_t STAGE = G * RB;
    extern __shared__ __align__(1024) uint8_t smem[];
    if (guarded_skip(a_in.guard)) return;
    const MergeSpan sp = merge_span<BLOCKED>(a_in);
    const MergeArgs blocked_args = merge_rebased<BLOCKED>(a_in, sp);
    const MergeArgs &a = BLOCKED ? blocked_args : a_in;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t ring = smem_u32(smem) + (uint32_t)warp * (S * STAGE);
    const uint32_t bars = smem_u32(smem) + (uint32_t)WARPS * (S * STAGE) + (uint32_t)warp * (S * 8);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < S; ++i) mbar_init32(bars + 8 * i, 1);
        mbar_fence_init();
    }
    __syncwarp();

    const int64_t gwarp = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t n_warps = (int64_t)gridDim.x * WARPS;
    const uint8_t *__restrict__ in = a.in;
    const uint32_t in_stride = (uint32_t)a.in_stride;
    // ring positions persist across ranges (every range drains its pipeline, so both end up equal)
    uint32_t slot_p = 0, slot_c = 0, par_c = 0;

    for (int64_t wi = gwarp; wi < sp.n_ranges; wi += n_warps) {
        const int64_t w = sp.win0 + wi;  // absolute window: also the index of the range's scratch slots
        const int64_t s = BLOCKED ? max(w * a.quantum, sp.pos0) : w * a.quantum;
        const int n_pos = (int)(min((w + 1) * (int64_t)a.quantum, sp.pos1) - s);
        if (n_pos <= 0) continue;
        const int n_groups = (n_pos + G - 1) / G;
        // s is a multiple of 32 and colidx is 16-byte aligned: the ids of a full group are one aligned int4
        // (not in a blocked launch, whose first range starts at the block's first position: scalar loads there)
        const int32_t *__restrict__ ids_ptr = a.colidx + s;
        auto load_ids = [&](int g) {
            if (!BLOCKED && g * G + G <= n_pos) return __ldg(reinterpret_cast<const int4 *>(ids_ptr) + g);
            int4 r;  // ragged tail of the last range: never read past nnz; missing ids repeat the first
            r.x = __ldg(ids_ptr + g * G);
            r.y = (g * G + 1 < n_pos) ? __ldg(ids_ptr + g * G + 1) : r.x;
            r.z = (g * G + 2 < n_pos) ? __ldg(ids_ptr + g * G + 2) : r.x;
            r.w = (BLOCKED && g * G + 3 < n_pos) ? __ldg(ids_ptr + g * G + 3) : r.x;
            return r;
        };

        int issued = 0;
        int4 ids_next = make_int4(0, 0, 0, 0);
        if (lane == 0) ids_next = load_ids(0);

        auto issue = [&]() {
            if (lane == 0) {
                const int4 ids = ids_next;
                if (issued + 1 < n_groups) ids_next = load_ids(issued + 1);
                const uint32_t bar = bars + 8 * slot_p;
                const uint32_t dst = ring + slot_p * STAGE;
                mbar_expect32(bar, STAGE);  // always 4 rows: a ragged group re-reads its first row
                if (GATHER4) {
                    tma_gather4(dst, &tmap, bar, 0, ids.x, ids.y, ids.z, ids.w);
                } else {
                    bulk_g2s32(dst, in + (uint64_t)(uint32_t)ids.x * in_stride, RB, bar);
                    bulk_g2s32(dst + RB, in + (uint64_t)(uint32_t)ids.y * in_stride, RB, bar);
                    bulk_g2s32(dst + 2 * RB, in + (uint64_t)(uint32_t)ids.z * in_stride, RB, bar);
                    bulk_g2s32(dst + 3 * RB, in + (uint64_t)(uint32_t)ids.w * in_stride, RB, bar);
                }
            }
            issued += 1;
            slot_p = (slot_p + 1 == S) ? 0 : slot_p + 1;
        };

        const int prologue = n_groups < S ? n_groups : S;
        for (int i = 0; i < prologue; ++i) issue();
        RowStateT<MODE> st;
        begin_range<MODE>(a, st, s);

        int pos = 0;
        for (int g = 0; g < n_groups; ++g) {
            mbar_wait32(bars + 8 * slot_c, par_c);
            const int cnt = min(G, n_pos - g * G);
            const uint32_t rows = ring + slot_c * STAGE + lane * (4 * MW);
            const uint32_t rows_h = ring + slot_c * STAGE + Lay<MODE>::MH_BYTES + lane * (4 * HW);
            int l = 0;
            while (l < cnt) {
                while (pos == st.re) {
                    flush_row<MODE>(a, st, w, n_pos, lane);
                    advance_row<MODE>(a, st, s);
                }
                if (l + 1 < cnt && pos + 1 < st.re) {
                    RowVec<MODE> r1, r2;
                    lds_words<MW>(rows + l * RB, r1.m);
                    lds_words<HW>(rows_h + l * RB, r1.h);
                    lds_words<MW>(rows + (l + 1) * RB, r2.m);
                    lds_words<HW>(rows_h + (l + 1) * RB, r2.h);
                    acc_merge2<MODE>(st, r1, r2);
                    l += 2;
                    pos += 2;
                } else {
                    RowVec<MODE> r1;
                    lds_words<MW>(rows + l * RB, r1.m);
                    lds_words<HW>(rows_h + l * RB, r1.h);
                    acc_merge<MODE>(st, r1);
                    l += 1;
                    pos += 1;
                }
            }
            slot_c += 1;
            if (slot_c == S) {
                slot_c = 0;
                par_c ^= 1u;
            }
            __syncwarp();  // every lane has finished reading the slot before it is refilled
            if (issued < n_groups) issue();
        }
        flush_row<MODE>(a, st, w, n_pos, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// fix-up: fold the partial records of rows cut by range boundaries; zero-fill rows with no in-edge
// ------------------------------------------------------------------------------------------------
template <int MODE, bool BLOCKED = false>
__global__ void __launch_bounds__(256) merge_fixup_kernel(const MergeArgs a_in) {
    constexpr int RB = Lay<MODE>::BYTES;
    if (guarded_skip(a_in.guard)) return;
    const MergeSpan sp = merge_span<BLOCKED>(a_in);
    const MergeArgs blocked_args = merge_rebased<BLOCKED>(a_in, sp);
    const MergeArgs &a = BLOCKED ? blocked_args : a_in;
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    // (A) one warp per range: the range in which a cut row STARTS folds all of its pieces
    for (int64_t wi = gwarp; wi < sp.n_ranges; wi += n_warps) {
        const int64_t w = sp.win0 + wi;
        const int64_t s = BLOCKED ? max(w * a.quantum, sp.pos0) : w * a.quantum;
        const int64_t e = min((w + 1) * (int64_t)a.quantum, sp.pos1);
        if (s >= e) continue;
        const int64_t last = row_of_position(a.rowptr, a.n_rows, e - 1);
        const int64_t rs = __ldg(a.rowptr + last), re = __ldg(a.rowptr + last + 1);
        if (rs < s || re <= e) continue;  // started earlier (someone else folds) or not cut
        const int64_t w_end = (re - 1) / a.quantum;  // range holding the last neighbour
        RowStateT<MODE> st;
        acc_reset<MODE>(st);
        acc_merge<MODE>(st, ldg_row<MODE>(a.scratch + (2 * w + 1) * (int64_t)RB, lane));
        for (int64_t x = w + 1; x <= w_end; x += 4) {
            RowVec<MODE> r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (x + u <= w_end) r[u] = ldg_row<MODE>(a.scratch + (2 * (x + u)) * (int64_t)RB, lane);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (x + u <= w_end) acc_merge<MODE>(st, r[u]);
        }
        store_row<MODE>(a, last, acc_row<MODE>(st), lane);
    }
    // (B) rows without any in-edge: all-zero record (scatter-max fill value), cardinality of an empty sketch
    RowVec<MODE> zero;
#pragma unroll
    for (int i = 0; i < (Lay<MODE>::MW > 0 ? Lay<MODE>::MW : 1); ++i) zero.m[i] = 0u;
#pragma unroll
    for (int i = 0; i < (Lay<MODE>::HW > 0 ? Lay<MODE>::HW : 1); ++i) zero.h[i] = 0u;
    for (int64_t r0 = gwarp * 32; r0 < a.n_rows; r0 += n_warps * 32) {
        const int64_t r = r0 + lane;
        const bool empty = r < a.n_rows && __ldg(a.rowptr + r) == __ldg(a.rowptr + r + 1);
        unsigned mask = __ballot_sync(FULL, empty);
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            store_row<MODE>(a, r0 + b, zero, lane);
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
                                                    int64_t width, const int *guard = nullptr) {
    if (guarded_skip(guard)) return;
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

// ---- tensor map of the previous-hop table for the gather4 engine ---------------------------------------
// cuTensorMapEncodeTiled is a driver entry point; it is resolved at run time so the library links against
// the CUDA runtime only (and still builds on a machine without a driver).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
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

// 2-D map over [n_rows x row_bytes / 4 uint32] with row pitch `stride` bytes; box = one row (gather4 fetches 4 boxes)
static int make_row_gather_map(CUtensorMap *map, const void *base, int64_t n_rows, int64_t stride, int row_bytes) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SS_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)(row_bytes / 4), (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)stride};
    cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 4), 1};
    cuuint32_t elem[2] = {1, 1};
    // L2 promotion = granularity of the L2 fills behind the gather (tuning knob SS_B200_TMA_L2PROMO = 0 / 64 / 128 / 256)
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (const char *e = getenv("SS_B200_TMA_L2PROMO")) {
        const int v = atoi(e);
        promo = v >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                         : (v >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                     : (v >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE));
    }
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(base), dims, strides, box, elem,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld stride=%lld)", (int)r, (long long)n_rows,
                  (long long)stride);
        return SS_ERR_CUDA;
    }
    return SS_OK;
}

// TMA engine configurations: (stages, warps per CTA, CTAs per SM).  Shared memory per warp = stages * 4 rows.
template <int S, int WARPS, int MIN_CTAS, bool GATHER4, int MODE>
static int launch_tma(const MergeArgs &a, const CUtensorMap &tmap, cudaStream_t st) {
    constexpr size_t smem = (size_t)WARPS * S * 4 * Lay<MODE>::BYTES + WARPS * S * 8;
    auto k = merge_tma_kernel<S, WARPS, MIN_CTAS, GATHER4, MODE>;
    SS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = 0, rc;
    if ((rc = resident_grid(k, WARPS * 32, smem, &grid)) != SS_OK) return rc;
    int64_t blocks = (a.n_ranges + WARPS - 1) / WARPS;
    if (blocks < grid) grid = (int)blocks;
    k<<<grid, WARPS * 32, smem, st>>>(a, tmap);
    SS_LAUNCH_CHECK("merge_tma_kernel");
    return SS_OK;
}

// measured on B200 (tools/tune_merge.py, profiles/r01_merge_tuning.txt): 24 warps x 3 stages wins while the
// tables are a few GB; on the largest graphs 16 warps x 4 stages (deeper per-warp pipeline) is ahead
static int tma_config(int64_t nnz) {
    const char *e = getenv("SS_B200_TMA_CFG");  // tuning knob
    if (e) return atoi(e);
    return nnz >= (1ll << 28) ? 0 : 1;
}

template <bool GATHER4>
static int launch_tma_cfg(const MergeArgs &a, const CUtensorMap &tmap, cudaStream_t st) {
    switch (tma_config(a.nnz)) {
        case 0: return launch_tma<4, 4, 4, GATHER4, LAY_FULL>(a, tmap, st);   // 16 warps / SM, 4 stages
        case 2: return launch_tma<2, 4, 8, GATHER4, LAY_FULL>(a, tmap, st);   // 32 warps / SM, 2 stages
        case 3: return launch_tma<6, 4, 3, GATHER4, LAY_FULL>(a, tmap, st);   // 12 warps / SM, 6 stages
        case 4: return launch_tma<3, 8, 3, GATHER4, LAY_FULL>(a, tmap, st);   // 24 warps / SM, 3 stages, 8-warp CTAs
        default: return launch_tma<3, 4, 6, GATHER4, LAY_FULL>(a, tmap, st);  // 24 warps / SM, 3 stages
    }
}

// blocked launches: the grid cannot be trimmed to the work (only the device knows it): always the resident grid
template <int S, int WARPS, int MIN_CTAS>
static int launch_tma_blocked(const MergeArgs &a, const CUtensorMap &tmap, cudaStream_t st) {
    constexpr size_t smem = (size_t)WARPS * S * 4 * Lay<LAY_FULL>::BYTES + WARPS * S * 8;
    auto k = merge_tma_kernel<S, WARPS, MIN_CTAS, true, LAY_FULL, true>;
    SS_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = 0, rc;
    if ((rc = resident_grid(k, WARPS * 32, smem, &grid)) != SS_OK) return rc;
    k<<<grid, WARPS * 32, smem, st>>>(a, tmap);
    SS_LAUNCH_CHECK("merge_tma_kernel (blocked)");
    merge_fixup_kernel<LAY_FULL, true><<<sm_count() * 8, 256, 0, st>>>(a);
    SS_LAUNCH_CHECK("merge_fixup_kernel (blocked)");
    return SS_OK;
}

// narrower rows (one half of a record, or a column half): gather4 only; the ring is smaller, so more warps fit
template <int MODE>
static int launch_tma_narrow(const MergeArgs &a, const CUtensorMap &tmap, cudaStream_t st) {
    switch (tma_config(a.nnz)) {
        case 0: return launch_tma<4, 4, 6, true, MODE>(a, tmap, st);    // 24 warps / SM, 4 stages
        case 2: return launch_tma<3, 4, 8, true, MODE>(a, tmap, st);    // 32 warps / SM, 3 stages
        default: return launch_tma<4, 4, 8, true, MODE>(a, tmap, st);   // 32 warps / SM, 4 stages
    }
}

template <int MODE>
static int launch_fixup(const MergeArgs &a, cudaStream_t st) {
    int64_t warps = a.n_ranges > (a.n_rows + 31) / 32 ? a.n_ranges : (a.n_rows + 31) / 32;
    int64_t blocks = (warps + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 8;
    merge_fixup_kernel<MODE><<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(a);
    SS_LAUNCH_CHECK("merge_fixup_kernel");
    return SS_OK;
}

static int layout_bytes(int layout) {
    switch (layout) {
        case SS_LAYOUT_FULL: return Lay<LAY_FULL>::BYTES;
        case SS_LAYOUT_MINHASH: return Lay<LAY_MH>::BYTES;
        case SS_LAYOUT_HLL: return Lay<LAY_HLL>::BYTES;
        case SS_LAYOUT_HALF: return Lay<LAY_HALF>::BYTES;
        default: return -1;
    }
}

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
                  int64_t in_rows, int64_t in_stride, void *rec_out, int64_t out_stride, int num_perm, int hll_p, void *workspace,
                  int64_t workspace_bytes, float *cards_out, int64_t cards_stride, const ss_hll_consts *hc, int variant,
                  ss_stream_t stream) {
    return ss_khop_merge_peers(rowptr, colidx, n_rows, nnz, rec_in, in_rows, in_stride, rec_out, out_stride, num_perm, hll_p,
                               workspace, workspace_bytes, cards_out, cards_stride, hc, variant, 0, nullptr, nullptr, nullptr,
                               nullptr, stream);
}

int ss_khop_merge_peers(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, int64_t nnz, const void *rec_in,
                        int64_t in_rows, int64_t in_stride, void *rec_out, int64_t out_stride, int num_perm, int hll_p,
                        void *workspace, int64_t workspace_bytes, float *cards_out, int64_t cards_stride,
                        const ss_hll_consts *hc, int variant, int n_peers, void *const *peer_rec_out,
                        float *const *peer_cards_out, void *mc_rec_out, float *mc_cards_out, ss_stream_t stream) {
    ss_merge_desc d;
    memset(&d, 0, sizeof(d));
    d.rowptr = rowptr; d.colidx = colidx; d.n_rows = n_rows; d.nnz = nnz;
    d.rec_in = rec_in; d.in_rows = in_rows; d.in_stride = in_stride; d.rec_out = rec_out; d.out_stride = out_stride;
    d.num_perm = num_perm; d.hll_p = hll_p; d.layout = SS_LAYOUT_FULL;
    d.workspace = workspace; d.workspace_bytes = workspace_bytes;
    d.cards_out = cards_out; d.cards_stride = cards_stride; d.hc = hc; d.variant = variant;
    d.n_peers = n_peers; d.peer_rec_out = peer_rec_out; d.peer_cards_out = peer_cards_out;
    d.mc_rec_out = mc_rec_out; d.mc_cards_out = mc_cards_out;
    return ss_khop_merge_ex(&d, stream);
}

int ss_khop_merge_ex(const ss_merge_desc *d, ss_stream_t stream) {
    SS_REQUIRE(d, "null descriptor passed to ss_khop_merge_ex");
    const int num_perm = d->num_perm, hll_p = d->hll_p;
    const int64_t n_rows = d->n_rows, nnz = d->nnz, in_rows = d->in_rows, in_stride = d->in_stride, out_stride = d->out_stride;
    int variant = d->variant;
    ss::RecordShape s;
    SS_REQUIRE(ss::make_shape(num_perm, hll_p, &s), "unsupported sketch shape num_perm=%d hll_p=%d", num_perm, hll_p);
    SS_REQUIRE(n_rows >= 0 && nnz >= 0 && in_rows >= 0, "negative size passed to ss_khop_merge");
    SS_REQUIRE(in_rows < (1ll << 31), "at most 2^31-1 rows in the previous-hop table");
    if (n_rows == 0) return SS_OK;
    SS_REQUIRE(n_rows < (1ll << 31), "at most 2^31-1 rows per call");
    SS_REQUIRE(d->rowptr && d->rec_in && d->rec_out, "null pointer passed to ss_khop_merge");
    SS_REQUIRE(nnz == 0 || d->colidx, "colidx is null");
    SS_REQUIRE((((uintptr_t)d->rec_in | (uintptr_t)d->rec_out) & 15) == 0, "record tables must be 16-byte aligned");
    SS_REQUIRE(((uintptr_t)d->colidx & 15) == 0, "colidx must be 16-byte aligned");
    const bool fast_shape = (num_perm == 128 && hll_p == 8);
    const int row_bytes = d->layout == SS_LAYOUT_FULL ? s.bytes : ss::layout_bytes(d->layout);
    SS_REQUIRE(row_bytes > 0, "unknown row layout %d", d->layout);
    SS_REQUIRE(d->layout == SS_LAYOUT_FULL || fast_shape, "partial row layouts need num_perm=128, hll_p=8");
    SS_REQUIRE(in_stride >= row_bytes && out_stride >= row_bytes && ((in_stride | out_stride) & 15) == 0,
               "row strides must be >= %d and multiples of 16", row_bytes);
    const bool want_cards = d->cards_out && (d->layout == SS_LAYOUT_FULL || d->layout == SS_LAYOUT_HLL);
    SS_REQUIRE(!d->cards_out || want_cards, "cardinalities need all 256 registers of a row (FULL or HLL layout)");
    ss::HllDev hd;
    memset(&hd, 0, sizeof(hd));
    if (want_cards) {
        int rc = ss::check_hll_consts(d->hc, hll_p);
        if (rc != SS_OK) return rc;
        hd = ss::to_dev(d->hc);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == SS_MERGE_AUTO) variant = fast_shape ? SS_MERGE_TMA : SS_MERGE_GENERIC;
    SS_REQUIRE(variant == SS_MERGE_GENERIC || fast_shape, "TMA/LDG merge kernels need num_perm=128, hll_p=8");
    SS_REQUIRE(d->layout == SS_LAYOUT_FULL || variant == SS_MERGE_TMA, "partial row layouts run on the TMA engine only");

    const int n_peers = d->n_peers;
    SS_REQUIRE(n_peers >= 0 && n_peers <= SS_MAX_PEERS, "n_peers must be in [0, %d]", SS_MAX_PEERS);
    SS_REQUIRE(n_peers == 0 || (variant != SS_MERGE_GENERIC && d->peer_rec_out && (!want_cards || d->peer_cards_out)),
               "peer stores need the P=128/p=8 engines and one pointer per peer");
    SS_REQUIRE(!d->mc_rec_out || variant != SS_MERGE_GENERIC, "multicast stores need the P=128/p=8 engines");
    SS_REQUIRE(!d->peer_mask || (n_peers > 0 && !d->mc_rec_out), "peer_mask selects among peer stores (not multicast)");
    if (variant == SS_MERGE_GENERIC) {
        SS_REQUIRE(!d->guard, "guarded launches need the P=128/p=8 engines");
        ss::GenericArgs g;
        g.rowptr = d->rowptr; g.colidx = d->colidx; g.n_rows = n_rows;
        g.in = (const uint8_t *)d->rec_in; g.in_stride = in_stride;
        g.out = (uint8_t *)d->rec_out; g.out_stride = out_stride;
        g.cards = d->cards_out; g.cards_stride = d->cards_stride; g.s = s; g.h = hd;
        int64_t blocks = (n_rows + 7) / 8;
        int64_t cap = (int64_t)ss::sm_count() * 8;
        ss::merge_generic_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(g);
        SS_LAUNCH_CHECK("merge_generic_kernel");
        return SS_OK;
    }

    ss::MergeArgs a;
    memset(&a, 0, sizeof(a));
    a.rowptr = d->rowptr; a.colidx = d->colidx; a.n_rows = n_rows; a.nnz = nnz;
    a.in = (const uint8_t *)d->rec_in; a.in_stride = in_stride;
    a.out = (uint8_t *)d->rec_out; a.out_stride = out_stride;
    a.quantum = ss::pick_quantum(nnz);
    a.n_ranges = ss::n_ranges_for(nnz, a.quantum);
    a.scratch = (uint8_t *)d->workspace;
    a.cards = want_cards ? d->cards_out : nullptr; a.cards_stride = d->cards_stride; a.h = hd;
    SS_REQUIRE(!d->mc_rec_out || (((uintptr_t)d->mc_rec_out & 15) == 0 && (!want_cards || d->mc_cards_out)),
               "multicast table must be 16-byte aligned and come with a multicast cards pointer");
    a.mc_out = (uint8_t *)d->mc_rec_out;
    a.mc_cards = d->mc_rec_out ? d->mc_cards_out : nullptr;
    a.n_peers = n_peers;
    a.peer_mask = d->peer_mask;
    a.guard = d->guard;
    a.block = (const long long *)d->block;
    if (d->block) {
        SS_REQUIRE(d->layout == SS_LAYOUT_FULL && variant == SS_MERGE_TMA && n_peers == 0 && !d->mc_rec_out,
                   "blocked launches: full records, TMA engine, single GPU");
        // two extra windows: a block's first and last range share their window with the neighbouring blocks
        a.n_ranges += 2;
    }
    for (int p = 0; p < SS_MAX_PEERS; ++p) {
        a.peer_out[p] = p < n_peers ? (uint8_t *)d->peer_rec_out[p] : nullptr;
        a.peer_cards[p] = (p < n_peers && want_cards) ? d->peer_cards_out[p] : nullptr;
        SS_REQUIRE(p >= n_peers || (a.peer_out[p] && ((uintptr_t)a.peer_out[p] & 15) == 0), "peer table %d is null or misaligned", p);
    }
    const int64_t need = 2 * a.n_ranges * (int64_t)row_bytes;
    SS_REQUIRE(d->workspace && ((uintptr_t)d->workspace & 15) == 0, "merge workspace must be a 16-byte aligned device buffer");
    if (d->workspace_bytes < need) {
        ss::set_error("merge workspace too small: %lld < %lld", (long long)d->workspace_bytes, (long long)need);
        return SS_ERR_WORKSPACE;
    }
    int grid = 0, rc;
    if (d->block) {
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof(tmap));
        if ((rc = ss::make_row_gather_map(&tmap, d->rec_in, in_rows, in_stride, row_bytes)) != SS_OK) return rc;
        return ss::tma_config(nnz) == 0 ? ss::launch_tma_blocked<4, 4, 4>(a, tmap, st) : ss::launch_tma_blocked<3, 4, 6>(a, tmap, st);
    }
    if (nnz > 0) {
        if (variant == SS_MERGE_TMA || variant == SS_MERGE_BULK) {
            CUtensorMap tmap;
            memset(&tmap, 0, sizeof(tmap));
            if (variant == SS_MERGE_TMA) {
                // rows addressed by colidx are < 2^31; the map covers every row the caller's table can hold
                if ((rc = ss::make_row_gather_map(&tmap, d->rec_in, in_rows, in_stride, row_bytes)) != SS_OK) return rc;
                switch (d->layout) {
                    case SS_LAYOUT_FULL: rc = ss::launch_tma_cfg<true>(a, tmap, st); break;
                    case SS_LAYOUT_MINHASH: rc = ss::launch_tma_narrow<ss::LAY_MH>(a, tmap, st); break;
                    case SS_LAYOUT_HLL: rc = ss::launch_tma_narrow<ss::LAY_HLL>(a, tmap, st); break;
                    default: rc = ss::launch_tma_narrow<ss::LAY_HALF>(a, tmap, st); break;
                }
            } else {
                rc = ss::launch_tma_cfg<false>(a, tmap, st);
            }
            if (rc != SS_OK) return rc;
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
    switch (d->layout) {
        case SS_LAYOUT_FULL: return ss::launch_fixup<ss::LAY_FULL>(a, st);
        case SS_LAYOUT_MINHASH: return ss::launch_fixup<ss::LAY_MH>(a, st);
        case SS_LAYOUT_HLL: return ss::launch_fixup<ss::LAY_HLL>(a, st);
        default: return ss::launch_fixup<ss::LAY_HALF>(a, st);
    }
}

static int prop_grid(int64_t n_rows) {
    int64_t blocks = (n_rows + 7) / 8;
    int64_t cap = (int64_t)ss::sm_count() * 8;
    return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

int ss_prop_min_i64(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, const int64_t *x, int64_t *out,
                    int64_t width, ss_stream_t stream) {
    return ss_prop_min_i64_guarded(rowptr, colidx, n_rows, x, out, width, nullptr, stream);
}

int ss_prop_min_i64_guarded(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, const int64_t *x, int64_t *out,
                            int64_t width, const int32_t *guard, ss_stream_t stream) {
    SS_REQUIRE(n_rows >= 0 && width >= 0, "negative size passed to ss_prop_min_i64");
    if (n_rows == 0 || width == 0) return SS_OK;
    SS_REQUIRE(rowptr && x && out, "null pointer passed to ss_prop_min_i64");
    ss::prop_kernel<int64_t, true><<<prop_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, n_rows, x, out, width, guard);
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
