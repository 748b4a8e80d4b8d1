// csr.cu -- K6: destination-keyed CSR from the reference's COO edge_index (+ implicit self loops).
//
// The reference never builds a CSR: PyG scatters over the COO list after add_self_loops(edge_index)
// (/root/reference/src/hashing.py:148; flow source -> target, reduction at edge_index[1]).  A pull-style
// merge needs in-neighbours per destination, so: histogram of destinations -> exclusive scan -> fill with
// per-row cursors.  Order inside a row is arbitrary (min/max are order independent).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ss {

constexpr int SCAN_ITEMS = 4;                   // per thread
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_TILE = SCAN_ITEMS * SCAN_BLOCK;  // rows per block

struct CsrWorkspace {
    uint32_t *deg;       // [n_rows]  in-degree histogram, later the fill cursors
    int64_t *tile_sum;   // [n_tiles] per-tile totals, then exclusive tile offsets
    int64_t n_tiles;
};

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static int64_t n_tiles_for(int64_t n_rows) { return (n_rows + SCAN_TILE - 1) / SCAN_TILE; }

static int64_t workspace_bytes(int64_t n_rows) {
    return align_up(n_rows * 4, 256) + align_up((n_tiles_for(n_rows) + 1) * 8, 256);
}

static CsrWorkspace carve(void *ws, int64_t n_rows) {
    CsrWorkspace w;
    w.deg = (uint32_t *)ws;
    w.tile_sum = (int64_t *)((char *)ws + align_up(n_rows * 4, 256));
    w.n_tiles = n_tiles_for(n_rows);
    return w;
}

// pass 1 over the COO list: in-degree histogram of the owned rows, max / min node id, and optional 32-bit device
// copies of both endpoint arrays.  The loads are fully coalesced, so src / dst may live in pinned HOST memory and
// are then consumed at PCIe rate with no staging copy (the 32-bit copies keep pass 2 off the bus).
// kernels of a memoised pipeline take a GUARD (device int32 or NULL): they return at once when *guard == 0
__device__ __forceinline__ bool csr_skip(const int *guard) { return guard && *reinterpret_cast<const volatile int *>(guard) == 0; }

__global__ void __launch_bounds__(256) degree_kernel(const int64_t *__restrict__ src, const int64_t *__restrict__ dst,
                                                      int64_t n_edges, int64_t row_begin, int64_t n_rows, uint32_t *deg,
                                                      int32_t *__restrict__ src32, int32_t *__restrict__ dst32,
                                                      long long *stats, const int *guard = nullptr) {
    if (csr_skip(guard)) return;
    long long mx = -1, mn = 0x7fffffffffffffffll;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t dv = dst[e];
        const int64_t d = dv - row_begin;
        if (d >= 0 && d < n_rows) atomicAdd(deg + d, 1u);
        if (stats) {
            const int64_t sv = src[e];
            mx = max(mx, (long long)max(sv, dv));
            mn = min(mn, (long long)min(sv, dv));
            if (src32) src32[e] = (int32_t)sv;
            if (dst32) dst32[e] = (int32_t)dv;
        }
    }
    if (stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = max(mx, __shfl_xor_sync(FULL, mx, o));
            mn = min(mn, __shfl_xor_sync(FULL, mn, o));
        }
        if ((threadIdx.x & 31) == 0) {
            if (mx >= 0) atomicMax(stats + 0, mx);
            if (mn != 0x7fffffffffffffffll) atomicMin(stats + 3, mn);
        }
    }
}

// number of implicit self loops: given by the host, or max(edge_index) + 1 computed by pass 1
__device__ __forceinline__ int64_t resolve_loops(int64_t n_self_loops, const long long *stats) {
    return n_self_loops >= 0 ? n_self_loops : (int64_t)stats[0] + 1;
}

__device__ __forceinline__ int64_t row_degree(const uint32_t *deg, int64_t r, int64_t n_rows, int64_t row_begin,
                                              int64_t n_self_loops) {
    if (r >= n_rows) return 0;
    return (int64_t)deg[r] + ((row_begin + r) < n_self_loops ? 1 : 0);
}

__device__ __forceinline__ int64_t block_sum_i64(int64_t v, int64_t *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    int64_t t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += smem[w];
    __syncthreads();
    return t;
}

// phase 1: total of each tile of SCAN_TILE rows
__global__ void __launch_bounds__(SCAN_BLOCK) tile_sum_kernel(const uint32_t *__restrict__ deg, int64_t n_rows,
                                                               int64_t row_begin, int64_t n_self_loops_arg,
                                                               const long long *stats, int64_t *__restrict__ tile_sum,
                                                               const int *guard = nullptr) {
    __shared__ int64_t sm[SCAN_BLOCK / 32];
    if (csr_skip(guard)) return;
    const int64_t n_self_loops = resolve_loops(n_self_loops_arg, stats);
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) v += row_degree(deg, base + k, n_rows, row_begin, n_self_loops);
    int64_t t = block_sum_i64(v, sm);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = t;
}

// phase 2: one block turns the tile totals into exclusive offsets (serial over chunks of blockDim tiles)
__global__ void __launch_bounds__(1024) tile_scan_kernel(int64_t *tile_sum, int64_t n_tiles, const int *guard = nullptr) {
    __shared__ int64_t warp_tot[32];
    __shared__ int64_t carry_s;
    if (csr_skip(guard)) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const int64_t x = (i < n_tiles) ? tile_sum[i] : 0;
        int64_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int64_t y = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_tot[lane];
            int64_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int64_t y = __shfl_up_sync(FULL, winc, o);
                if (lane >= o) winc += y;
            }
            warp_tot[lane] = winc - w;  // exclusive prefix of warp totals
        }
        __syncthreads();
        const int64_t carry = carry_s;
        const int64_t excl = carry + warp_tot[warp] + inc - x;
        if (i < n_tiles) tile_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sum[n_tiles] = carry_s;  // grand total
}

// phase 3: exclusive scan inside each tile + tile offset -> rowptr
__global__ void __launch_bounds__(SCAN_BLOCK) rowptr_kernel(const uint32_t *__restrict__ deg, int64_t n_rows,
                                                             int64_t row_begin, int64_t n_self_loops_arg,
                                                             long long *stats, const int64_t *__restrict__ tile_off,
                                                             int64_t *__restrict__ rowptr, const int *guard = nullptr) {
    __shared__ int64_t warp_tot[SCAN_BLOCK / 32];
    if (csr_skip(guard)) return;
    const int64_t n_self_loops = resolve_loops(n_self_loops_arg, stats);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t d[SCAN_ITEMS];
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        d[k] = row_degree(deg, base + k, n_rows, row_begin, n_self_loops);
        v += d[k];
    }
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int64_t off = tile_off[blockIdx.x];
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    int64_t run = off + inc - v;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n_rows) rowptr[base + k] = run;
        run += d[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        rowptr[n_rows] = tile_off[gridDim.x];
        if (stats) {
            stats[1] = tile_off[gridDim.x];  // nnz
            stats[2] = n_self_loops;
        }
    }
}

// fill cursors that already hold the absolute write position: cursor[r] = low 32 bits of rowptr[r]
__global__ void __launch_bounds__(256) cursor_init_kernel(const int64_t *__restrict__ rowptr, int64_t n_rows,
                                                           uint32_t *__restrict__ cursor, const int *guard = nullptr) {
    if (csr_skip(guard)) return;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x)
        cursor[r] = (uint32_t)rowptr[r];
}

// ABS_CURSOR: the cursors start at rowptr[d] (mod 2^32), so one atomic returns the write position and the random
// 8-byte rowptr read per edge disappears (one third of the kernel's sector traffic).  When the local nnz does not
// fit 32 bits the high part is recovered from rowptr[d] (uniform branch, whole grid takes the same side).
template <typename IdT, bool ABS_CURSOR>
__global__ void __launch_bounds__(256) fill_kernel(const IdT *__restrict__ src, const IdT *__restrict__ dst,
                                                    int64_t n_edges, int64_t n_self_loops_arg, const long long *stats,
                                                    int64_t row_begin, int64_t n_rows, const int64_t *__restrict__ rowptr,
                                                    uint32_t *cursor, int32_t *__restrict__ colidx,
                                                    const int *guard = nullptr, int64_t clamp_cols = 0) {
    if (csr_skip(guard)) return;
    const int64_t n_self_loops = resolve_loops(n_self_loops_arg, stats);
    const int64_t total = n_edges + n_rows;
    const bool fits32 = ABS_CURSOR && (uint64_t)__ldg(rowptr + n_rows) < (1ull << 32);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t d, s;
        if (t < n_edges) {
            d = (int64_t)dst[t] - row_begin;
            s = (int64_t)src[t];
            if (d < 0 || d >= n_rows) continue;
            // deferred validation (sync-free callers): an out-of-range source must never become a gather index;
            // the id statistics of the degree pass carry the error to the host
            if (clamp_cols > 0 && (s < 0 || s >= clamp_cols)) s = 0;
        } else {
            d = t - n_edges;
            s = row_begin + d;
            if (s >= n_self_loops) continue;
        }
        const uint32_t k = atomicAdd(cursor + d, 1u);
        if (ABS_CURSOR) {
            if (fits32) {
                colidx[k] = (int32_t)s;
            } else {
                const int64_t base = rowptr[d];
                colidx[base + (uint32_t)(k - (uint32_t)base)] = (int32_t)s;
            }
        } else {
            colidx[rowptr[d] + k] = (int32_t)s;
        }
    }
}

// ---- EXPERIMENTAL (opt-in, not on the default path; unmeasured): coarse binning of the edge list by destination
// block before the fill.  The fill's cost is its scattered 4-byte stores: an edge list sorted by source hits a
// random row of the 2 GB colidx array per edge, so every 32-byte sector is written ~8 times at unrelated moments
// and travels to DRAM half empty each time.  One extra streaming pass that groups the edges by dst >> shift
// (<= 2048 buckets) makes the fill walk colidx window by window (1-2 MB at a time, L2 resident), so sectors are
// completed in L2.  The bucket capacities are free: they are differences of the rowptr the scan just produced.
constexpr int BIN_MAX_BUCKETS = 2048;
constexpr int BIN_EPT = 32;                       // edges per thread per tile
constexpr int BIN_TILE = 256 * BIN_EPT;

static int bin_shift_for(int64_t n_rows) {
    int s = 0;
    while (((n_rows + ((int64_t)1 << s) - 1) >> s) > BIN_MAX_BUCKETS) ++s;
    return s;
}

// edges (self loops excluded) whose local destination row is < r
__device__ __forceinline__ int64_t edge_prefix(const int64_t *__restrict__ rowptr, int64_t r, int64_t row_begin,
                                               int64_t n_self_loops) {
    int64_t loops = n_self_loops - row_begin;
    loops = loops < 0 ? 0 : (loops > r ? r : loops);
    return rowptr[r] - loops;
}

template <typename IdT>
__global__ void __launch_bounds__(256) bin_edges_kernel(const IdT *__restrict__ src, const IdT *__restrict__ dst,
                                                         int64_t n_edges, int64_t n_self_loops_arg, const long long *stats,
                                                         int64_t row_begin, int64_t n_rows, int shift, int n_buckets,
                                                         const int64_t *__restrict__ rowptr, unsigned long long *cursors,
                                                         int32_t *__restrict__ src_out, int32_t *__restrict__ dst_out) {
    __shared__ uint32_t count[BIN_MAX_BUCKETS];          // entries of this tile per bucket, then the running offset
    __shared__ unsigned long long base[BIN_MAX_BUCKETS]; // where this tile's run starts in each bucket
    const int64_t n_self_loops = resolve_loops(n_self_loops_arg, stats);
    const int64_t n_tiles = (n_edges + BIN_TILE - 1) / BIN_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int b = threadIdx.x; b < n_buckets; b += blockDim.x) count[b] = 0u;
        __syncthreads();
        int32_t s32[BIN_EPT], d32[BIN_EPT];  // d32 = -1 marks an edge that is not binned
        const int64_t e0 = tile * BIN_TILE + threadIdx.x;
#pragma unroll
        for (int i = 0; i < BIN_EPT; ++i) {
            const int64_t e = e0 + (int64_t)i * 256;
            d32[i] = -1;
            s32[i] = 0;
            if (e < n_edges) {
                const int64_t dv = (int64_t)dst[e];
                const int64_t d = dv - row_begin;
                if (d >= 0 && d < n_rows) {
                    d32[i] = (int32_t)dv;
                    s32[i] = (int32_t)src[e];
                    atomicAdd(&count[(int)(d >> shift)], 1u);
                }
            }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < n_buckets; b += blockDim.x) {
            const uint32_t c = count[b];
            if (c) {
                const unsigned long long start = (unsigned long long)edge_prefix(rowptr, (int64_t)b << shift, row_begin, n_self_loops);
                base[b] = start + atomicAdd(cursors + b, (unsigned long long)c);
                count[b] = 0u;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < BIN_EPT; ++i) {
            if (d32[i] >= 0) {
                const int b = (int)(((int64_t)d32[i] - row_begin) >> shift);
                const unsigned long long pos = base[b] + atomicAdd(&count[b], 1u);
                src_out[pos] = s32[i];
                dst_out[pos] = d32[i];
            }
        }
        __syncthreads();
    }
}

// ---- streaming CSR for edge lists that are already ordered by the CSR key ------------------------------------
// PyG's coalesce / to_undirected (what the reference's datasets hand to build_hash_tables, datasets/elph.py:63-66)
// emit edge lists sorted by (row, col).  For such a list the CSR keyed by `key` needs no histogram, no atomics on
// rows and no scattered stores: rowptr[r] is the position of the first edge with key >= r, colidx is the other
// endpoint in list order -- ONE coalesced pass (read 16 B, write 4 B per edge).
//   key = edge_index[1] (dst) : exact destination-keyed CSR
//   key = edge_index[0] (src) : the SOURCE-keyed CSR, which equals the destination-keyed one exactly when the edge
//                               multiset is symmetric.  The pass accumulates keyed 2 x 64-bit multiset fingerprints of
//                               (key, val) and (val, key); the host accepts the result only if they agree (a false
//                               accept needs a collision of a 128-bit fingerprint under a per-process random key)
// With self loops (add_self_loops without num_nodes: one per id <= max id, hashing.py:148) the loop of row r is
// stored FIRST in its row: every key that occurs is <= max id, so the position of edge e is e + key[e] + 1 and
// rowptr[r] = (#edges with key < r) + r for r <= last key; ss_csr_sorted_finish completes the rows above it.
// The pass is speculative: order violations / range errors are counted, never acted upon, and every write is
// bounds-guarded; the host falls back to the histogram path when stats say so.
struct SortedArgs {
    const int64_t *key;
    const int64_t *val;
    int64_t n;              // edges in this chunk
    int64_t e_base;         // index of key[0] in the whole list
    int64_t e_origin;       // index (in the whole list) of the first edge of this CSR: positions are relative to it
    int64_t row_begin;      // first row of this CSR (node-sharded builds); rows are [row_begin, row_begin + n_rows)
    int64_t n_rows;
    int64_t capacity;       // entries colidx can hold
    int loops;              // 1: implicit self loops (first in row)
    unsigned long long ka, kb;  // fingerprint keys
    int64_t *rowptr;
    int32_t *colidx;
    long long *stats;       // [12]
    long long *carry;       // [2]: last key / last val of the previous chunk
};

// 32-bit keyed mixes (murmur3-style finalisers) of an ORDERED id pair; four independent 32-bit hashes per direction
// would be overkill: two, accumulated into 64-bit sums, give a multiset fingerprint whose false-accept probability
// is ~2^-46 per accumulator (sum of E random 32-bit values spreads over 2^32 sqrt(E) values) and ~2^-92 for the pair
__device__ __forceinline__ uint32_t fp_mix_a(uint32_t x, uint32_t y, uint32_t key) {
    uint32_t h = (x * 0x9e3779b1u) ^ y ^ key;
    h ^= h >> 15; h *= 0x85ebca77u; h ^= h >> 13; h *= 0xc2b2ae3du; h ^= h >> 16;
    return h;
}
__device__ __forceinline__ uint32_t fp_mix_b(uint32_t x, uint32_t y, uint32_t key) {
    uint32_t h = (y * 0x27d4eb2fu) + (x ^ key) * 0x165667b1u;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256) sorted_csr_kernel(const SortedArgs a) {
    constexpr int UNR = 4;  // 32-edge windows per trip: their loads are issued together (memory-level parallelism)
    const int lane = threadIdx.x & 31;
    unsigned long long fa = 0, fb = 0, ra = 0, rb = 0;
    int mx = -1, mn = 0x7fffffff;
    unsigned bad_val = 0, bad_range = 0;
    bool dead = false;  // an order violation was seen (by this warp or any other): the result will be discarded
    unsigned long long *st = (unsigned long long *)a.stats;
    const uint32_t ka = (uint32_t)a.ka, kb = (uint32_t)a.kb;
    const int row_lo = (int)a.row_begin, row_hi = (int)(a.row_begin + a.n_rows);
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // a warp owns UNR consecutive windows per trip: [w0, w0 + 32 UNR); whole warps iterate together so that the
    // shuffles are well defined (the tail is padded with inactive lanes)
    for (int64_t w0 = gwarp * (32 * UNR); w0 < a.n; w0 += n_warps * (32 * UNR)) {
        int64_t k64[UNR], v64[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t i = w0 + u * 32 + lane;
            k64[u] = i < a.n ? a.key[i] : 0;
            v64[u] = i < a.n ? a.val[i] : 0;
        }
        // the edge in front of this trip (lane 0 of the first window needs it)
        int64_t p0 = -1, p1 = -1;
        bool have_prev = false;
        if (lane == 0) {
            if (w0 > 0) { p0 = a.key[w0 - 1]; p1 = a.val[w0 - 1]; have_prev = true; }
            else if (a.e_base > a.e_origin) { p0 = a.carry[0]; p1 = a.carry[1]; have_prev = true; }
        }
        // The pass is speculative.  On an unordered list the "rows that start here" ranges below are garbage and
        // could add up to E * n_rows writes, so all structural work stops as soon as ANY warp has seen a violation
        // (every window is checked before it is acted upon; the global counter is polled once per trip).
        if (!dead) dead = __any_sync(FULL, *(volatile unsigned long long *)(st + 8) != 0ull);
        int kp_carry = 0, vp_carry = 0;  // last edge of the previous window (lane 31's values)
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int64_t i = w0 + u * 32 + lane;
            if (w0 + u * 32 >= a.n) break;  // warp-uniform
            const bool live = i < a.n;
            // ids outside [0, 2^31) are range errors; inside the loop everything is 32-bit (-2 marks an invalid id)
            const bool id_ok = ((uint64_t)k64[u] | (uint64_t)v64[u]) < (1ull << 31);
            const int k = id_ok ? (int)k64[u] : -2, v = id_ok ? (int)v64[u] : -2;
            int kp = __shfl_up_sync(FULL, k, 1);
            int vp = __shfl_up_sync(FULL, v, 1);
            if (lane == 0) {
                if (u > 0) { kp = kp_carry; vp = vp_carry; }
                else if (have_prev) {
                    const bool okp = ((uint64_t)p0 | (uint64_t)p1) < (1ull << 31);
                    kp = okp ? (int)p0 : -2; vp = okp ? (int)p1 : -2;
                } else { kp = row_lo - 1; vp = -1; }
            }
            kp_carry = __shfl_sync(FULL, k, 31);
            vp_carry = __shfl_sync(FULL, v, 31);
            const int64_t e = a.e_base + i - a.e_origin;   // position among the edges of this CSR
            const bool ok = live && id_ok && k >= row_lo && k < row_hi && kp >= row_lo - 1;
            if (live) {
                if (!id_ok || k < row_lo || k >= row_hi) bad_range = 1;
                if (id_ok) { mx = max(mx, max(k, v)); mn = min(mn, min(k, v)); }
                else { mn = (k64[u] < 0 || v64[u] < 0) ? -1 : mn; mx = (k64[u] >= (1ll << 31) || v64[u] >= (1ll << 31)) ? 0x7fffffff : mx; }
                if (v < vp) bad_val = 1;   // (only meaningful for the caller's "is the OTHER row ordered" question)
                if (i == a.n - 1) { a.carry[0] = k64[u]; a.carry[1] = v64[u]; }  // read by the next chunk (stream order)
            }
            const bool viol = __any_sync(FULL, live && k < kp);
            if (viol && !dead && lane == 0) atomicAdd(st + 8, 1ull);
            dead = dead || viol;
            if (dead) continue;
            if (live) {
                const uint32_t hf = fp_mix_a((uint32_t)k, (uint32_t)v, ka), hr = fp_mix_a((uint32_t)v, (uint32_t)k, ka);
                fa += hf; ra += hr;
                fb += fp_mix_b((uint32_t)k, (uint32_t)v, kb); rb += fp_mix_b((uint32_t)v, (uint32_t)k, kb);
            }
            // rows (kp, k] start at this edge.  Short gaps inline; long ones (runs of rows without edges) warp-wide
            const int gap = (ok && k > kp) ? (k - kp) : 0;
            if (gap > 0 && gap <= 4) {
                for (int r = kp + 1; r <= k; ++r) {
                    const int64_t pos = e + (a.loops ? r - row_lo : 0);
                    a.rowptr[r - row_lo] = pos;
                    if (a.loops && pos < a.capacity) a.colidx[pos] = r;
                }
            }
            unsigned longs = __ballot_sync(FULL, gap > 4);
            while (longs) {
                const int b = __ffs(longs) - 1;
                longs &= longs - 1;
                const int r0 = __shfl_sync(FULL, kp, b) + 1, r1 = __shfl_sync(FULL, k, b);
                const int64_t eb = a.e_base + (w0 + u * 32 + b) - a.e_origin;
                for (int r = r0 + lane; r <= r1; r += 32) {
                    const int64_t pos = eb + (a.loops ? r - row_lo : 0);
                    a.rowptr[r - row_lo] = pos;
                    if (a.loops && pos < a.capacity) a.colidx[pos] = r;
                }
            }
            if (ok) {
                const int64_t pos = e + (a.loops ? k - row_lo + 1 : 0);
                if (pos < a.capacity) a.colidx[pos] = v;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        fa += __shfl_xor_sync(FULL, fa, o); fb += __shfl_xor_sync(FULL, fb, o);
        ra += __shfl_xor_sync(FULL, ra, o); rb += __shfl_xor_sync(FULL, rb, o);
        mx = max(mx, __shfl_xor_sync(FULL, mx, o)); mn = min(mn, __shfl_xor_sync(FULL, mn, o));
    }
    bad_val = __any_sync(FULL, bad_val); bad_range = __any_sync(FULL, bad_range);
    if (lane == 0) {
        if (mx >= 0) atomicMax(a.stats + 0, (long long)mx);
        if (mn != 0x7fffffff) atomicMin(a.stats + 3, (long long)mn);
        atomicAdd(st + 4, fa); atomicAdd(st + 5, fb); atomicAdd(st + 6, ra); atomicAdd(st + 7, rb);
        if (bad_val) atomicAdd(st + 9, 1ull);
        if (bad_range) atomicAdd(st + 10, 1ull);
    }
}

// rows above the last key of this CSR: r in (last, row_begin + n_rows]: rowptr = E + (self loops below r inside the
// block), self loops of the rows below L = max id + 1; nnz and loop count into stats
__global__ void __launch_bounds__(256) sorted_csr_finish_kernel(int64_t n_edges, int64_t row_begin, int64_t n_rows, int64_t capacity,
                                                                 int loops, int64_t *rowptr, int32_t *colidx, long long *stats,
                                                                 const long long *carry) {
    int64_t last = n_edges > 0 ? carry[0] : row_begin - 1;
    last = last < row_begin - 1 ? row_begin - 1 : (last > row_begin + n_rows ? row_begin + n_rows : last);
    const int64_t L = loops ? (int64_t)stats[0] + 1 : 0;  // max id + 1 (global)
    auto loops_below = [&](int64_t r) {  // self loops of rows [row_begin, r)
        int64_t x = (r < L ? r : L) - row_begin;
        return x < 0 ? 0 : x;
    };
    for (int64_t r = last + 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= row_begin + n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pos = n_edges + loops_below(r);
        rowptr[r - row_begin] = pos;
        if (r < L && r < row_begin + n_rows && pos < capacity) colidx[pos] = (int32_t)r;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        stats[1] = n_edges + loops_below(row_begin + n_rows);  // nnz (only meaningful when max id < number of nodes)
        stats[2] = L;
    }
}

// the rows that are COMPLETE after a chunk of a key-ordered list has been absorbed by sorted_csr_kernel: every row below
// the chunk's last key (the row of the last key itself may continue in the next chunk).  Writes the descriptor of the
// block [previous block's end, that row) for a blocked merge launch; `final`: everything up to n_rows (after
// sorted_csr_finish_kernel).  If the pass has seen an order violation or a range error so far, the block is EMPTY:
// rowptr may be garbage then and the speculative result will be discarded anyway.
__global__ void sorted_block_kernel(const long long *carry, const int64_t *rowptr, int64_t n_rows, int64_t capacity,
                                    const long long *stats, int final, const long long *prev, long long *block) {
    const long long row_begin = prev ? prev[1] : 0, pos_begin = prev ? prev[3] : 0;
    long long row_end = final ? n_rows : carry[0];
    const bool bad = stats[8] != 0 || stats[10] != 0;
    if (bad || row_end < row_begin) row_end = row_begin;
    if (row_end > n_rows) row_end = n_rows;
    long long pos_end = row_end > row_begin ? rowptr[row_end] : pos_begin;
    if (pos_end < pos_begin || pos_end > capacity) { row_end = row_begin; pos_end = pos_begin; }
    block[0] = row_begin; block[1] = row_end; block[2] = pos_begin; block[3] = pos_end;
}

// cost-balanced row blocks of a key-ordered list without a histogram: cost(r) = (#edges with key < r) + row_cost * r,
// cut at the cumulative shares cum[q]; one WARP per cut, binary search over rows with a 32-ary search over the list
// inside (key may be a pinned host pointer: ~150 dependent round trips per cut instead of ~700)
__global__ void __launch_bounds__(256) sorted_bounds_kernel(const int64_t *__restrict__ key, int64_t n_edges, int64_t n_rows,
                                                             double row_cost, const double *__restrict__ cum, int n_cuts,
                                                             int64_t *bounds, int64_t *edge_off) {
    const int lane = threadIdx.x & 31;
    const int q = threadIdx.x >> 5;
    auto lower = [&](int64_t r) {  // first e with key[e] >= r  (warp-cooperative; identical result on every lane)
        int64_t lo = 0, hi = n_edges;  // invariant: key[lo - 1] < r <= key[hi]
        while (hi - lo > 0) {
            const int64_t span = hi - lo;
            const int64_t step = (span + 32) / 33;  // 32 probes cut the span into 33 pieces
            const int64_t pos = lo + (int64_t)(lane + 1) * step - 1;
            const bool lt = pos < hi ? (key[pos] < r) : false;
            const unsigned m = __ballot_sync(FULL, lt);
            const int n_lt = __popc(m);  // probes are ordered: the first n_lt of them are < r
            const int64_t new_lo = n_lt ? min(lo + (int64_t)n_lt * step, hi) : lo;
            const int64_t new_hi = n_lt < 32 ? min(lo + (int64_t)(n_lt + 1) * step - 1, hi) : hi;
            lo = new_lo;
            hi = new_hi;
        }
        return lo;
    };
    if (threadIdx.x == 0) { bounds[0] = 0; edge_off[0] = 0; bounds[n_cuts + 1] = n_rows; edge_off[n_cuts + 1] = n_edges; }
    if (q >= n_cuts) return;
    const double total = (double)n_edges + row_cost * (double)n_rows;
    const double target = cum[q] * total;
    int64_t lo = 0, hi = n_rows;  // smallest r with cost(r) >= target
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((double)lower(mid) + row_cost * (double)mid < target) lo = mid + 1; else hi = mid;
    }
    const int64_t e = lower(lo);
    if (lane == 0) {
        bounds[q + 1] = lo;
        edge_off[q + 1] = e;
    }
}

// halo of a SYMMETRIC graph from the owner's own CSR rows (no exchange): row r is read by rank q exactly when r has an
// in-neighbour owned by q, so peer_mask[r] = OR over r's neighbours c of bit(position of owner(c) among the other
// ranks); and every neighbour c is a row this rank reads: mark[c] = 1.  Element-parallel over the neighbour list (a
// hub row of 700k neighbours is spread over the whole grid, not walked by one warp): every 32-position window finds
// the row of its first position by binary search, its lanes walk forward from there, lanes of one row combine their
// bits and one of them issues the atomic OR.  peer_mask must be zeroed by the caller and padded to a multiple of 4 bytes.
__global__ void __launch_bounds__(256) halo_from_csr_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                                             int64_t n_rows, int64_t nnz, const int64_t *__restrict__ bounds,
                                                             int n_ranks, int rank, uint32_t *__restrict__ peer_mask_words,
                                                             uint8_t *__restrict__ mark) {
    const int lane = threadIdx.x & 31;
    int b[SS_MAX_PEERS];  // b[q - 1] = first row of rank q (registers: the loops below are fully unrolled)
#pragma unroll
    for (int q = 1; q <= SS_MAX_PEERS; ++q) b[q - 1] = q < n_ranks ? (int)bounds[q] : 0x7fffffff;
    // a warp owns CHUNKS of 32 consecutive windows: one binary search per chunk (row of its first position), after that
    // the running row only moves forward
    constexpr int WINDOWS = 32;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_chunks = (nnz + 32 * WINDOWS - 1) / (32 * WINDOWS);
    for (int64_t ch = gwarp; ch < n_chunks; ch += n_warps) {
      const int64_t c0 = ch * (32 * WINDOWS);
      int64_t lo = 0, hi = n_rows;  // largest r with rowptr[r] <= c0
      while (hi - lo > 1) {
          const int64_t mid = (lo + hi) >> 1;
          if (__ldg(rowptr + mid) <= c0) lo = mid; else hi = mid;
      }
      int64_t base_row = lo;  // row of the current window's first position (identical on every lane)
      for (int w = 0; w < WINDOWS; ++w) {
        const int64_t x0 = c0 + (int64_t)w * 32;
        if (x0 >= nnz) break;  // warp-uniform
        const int64_t x = x0 + lane;
        const bool live = x < nnz;
        int64_t row = base_row;
        unsigned bits = 0;
        if (live) {
            while (row + 1 < n_rows && __ldg(rowptr + row + 1) <= x) ++row;   // short walks: rows are consecutive
            const int c = __ldg(colidx + x);
            if (mark[c] == 0) mark[c] = 1;  // test first: a row is read by many lists but marked once
            int o = 0;
#pragma unroll
            for (int q = 0; q < SS_MAX_PEERS; ++q) o += (c >= b[q]) ? 1 : 0;
            if (o != rank) bits = 1u << (o < rank ? o : o - 1);
        }
        // the next window starts in the row of this window's last live position (or one further: the walk finds out)
        base_row = __shfl_sync(FULL, row, min(31, (int)min((int64_t)31, nnz - 1 - x0)));
        // lanes of one row are CONSECUTIVE (positions and rows both ascend): segmented OR-scan in 5 uniform steps (every
        // lane executes every shuffle), then the last lane of each segment publishes its row's bits
        const int64_t key = live ? row : (int64_t)-1 - lane;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, bits, o);
            const int64_t k2 = __shfl_up_sync(FULL, key, o);
            if (lane >= o && k2 == key) bits |= t;
        }
        const int64_t k_next = __shfl_down_sync(FULL, key, 1);
        if (live && bits && (lane == 31 || k_next != key)) atomicOr(peer_mask_words + (row >> 2), bits << (8 * (row & 3)));
      }
    }
}

// mark[colidx[e]] = 1: the rows a rank's neighbour lists read (its halo + own rows), for the halo push of the
// node-sharded build
__global__ void __launch_bounds__(256) mark_rows_kernel(const int32_t *__restrict__ colidx, int64_t nnz, uint8_t *mark) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x)
        mark[__ldg(colidx + e)] = 1;
}

// SS_B200_CSR_FILL=legacy keeps zero-based cursors + a rowptr read per edge (the first version)
static bool abs_cursor_fill() {
    const char *e = getenv("SS_B200_CSR_FILL");
    return !(e && e[0] == 'l');
}

}  // namespace ss

extern "C" {

int64_t ss_csr_workspace_bytes(int64_t n_rows) {
    if (n_rows < 0) return SS_ERR_INVALID;
    return ss::workspace_bytes(n_rows);
}

// pass 1 over ONE CHUNK of the COO list: accumulates the in-degree histogram and the id statistics, writes the
// int32 copies of the chunk.  `first` zeroes the histogram and initialises the statistics.
int ss_csr_degree_chunk(const int64_t *src, const int64_t *dst, int64_t n_edges, int64_t row_begin, int64_t n_rows,
                        int32_t *src32_out, int32_t *dst32_out, int64_t *stats_io, void *workspace, int64_t workspace_bytes,
                        int first, ss_stream_t stream) {
    SS_REQUIRE(n_edges >= 0 && n_rows >= 0 && row_begin >= 0, "negative size passed to ss_csr_degree_chunk");
    SS_REQUIRE(workspace, "null workspace passed to ss_csr_degree_chunk");
    SS_REQUIRE(n_edges == 0 || dst, "dst is null");
    SS_REQUIRE(!stats_io || n_edges == 0 || src, "src is required for the id statistics");
    SS_REQUIRE(!(src32_out || dst32_out) || stats_io, "32-bit copies are written together with the statistics");
    SS_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    if (workspace_bytes < ss::workspace_bytes(n_rows)) {
        ss::set_error("csr workspace too small: %lld < %lld", (long long)workspace_bytes,
                      (long long)ss::workspace_bytes(n_rows));
        return SS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long *stats = (long long *)stats_io;
    ss::CsrWorkspace w = ss::carve(workspace, n_rows);
    if (first) {
        if (stats) {
            const long long init[4] = {-1, 0, 0, 0x7fffffffffffffffll};
            SS_CUDA(cudaMemcpyAsync(stats, init, sizeof(init), cudaMemcpyHostToDevice, st));  // staged before returning
        }
        if (n_rows > 0) SS_CUDA(cudaMemsetAsync(w.deg, 0, (size_t)n_rows * 4, st));
    }
    if (n_edges > 0) {
        int64_t blocks = (n_edges + 255) / 256;
        int64_t cap = (int64_t)ss::sm_count() * 32;
        ss::degree_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(src, dst, n_edges, row_begin, n_rows, w.deg,
                                                                         src32_out, dst32_out, stats);
        SS_LAUNCH_CHECK("degree_kernel");
    }
    return SS_OK;
}

// exclusive scan of the accumulated in-degrees (+ implicit self loops) -> rowptr; completes stats[1..2]
int ss_csr_rowptr_finish(int64_t n_self_loops, int64_t row_begin, int64_t n_rows, int64_t *rowptr, int64_t *stats_io,
                         void *workspace, int64_t workspace_bytes, ss_stream_t stream) {
    SS_REQUIRE(n_rows >= 0 && row_begin >= 0, "negative size passed to ss_csr_rowptr_finish");
    SS_REQUIRE(rowptr && workspace, "null pointer passed to ss_csr_rowptr_finish");
    SS_REQUIRE(n_self_loops >= 0 || stats_io, "n_self_loops < 0 (= max id + 1) needs the statistics");
    SS_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    if (workspace_bytes < ss::workspace_bytes(n_rows)) {
        ss::set_error("csr workspace too small: %lld < %lld", (long long)workspace_bytes,
                      (long long)ss::workspace_bytes(n_rows));
        return SS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long *stats = (long long *)stats_io;
    if (n_rows == 0) {
        SS_CUDA(cudaMemsetAsync(rowptr, 0, 8, st));
        return SS_OK;
    }
    ss::CsrWorkspace w = ss::carve(workspace, n_rows);
    ss::tile_sum_kernel<<<(int)w.n_tiles, ss::SCAN_BLOCK, 0, st>>>(w.deg, n_rows, row_begin, n_self_loops, stats, w.tile_sum);
    SS_LAUNCH_CHECK("tile_sum_kernel");
    ss::tile_scan_kernel<<<1, 1024, 0, st>>>(w.tile_sum, w.n_tiles);
    SS_LAUNCH_CHECK("tile_scan_kernel");
    ss::rowptr_kernel<<<(int)w.n_tiles, ss::SCAN_BLOCK, 0, st>>>(w.deg, n_rows, row_begin, n_self_loops, stats, w.tile_sum,
                                                                rowptr);
    SS_LAUNCH_CHECK("rowptr_kernel");
    return SS_OK;
}

int ss_csr_rowptr(const int64_t *src, const int64_t *dst, int64_t n_edges, int64_t n_self_loops, int64_t row_begin,
                  int64_t n_rows, int64_t *rowptr, int32_t *src32_out, int32_t *dst32_out, int64_t *stats_out,
                  void *workspace, int64_t workspace_bytes, ss_stream_t stream) {
    SS_REQUIRE(rowptr, "null pointer passed to ss_csr_rowptr");
    SS_REQUIRE(n_self_loops >= 0 || stats_out, "n_self_loops < 0 (= max id + 1) needs stats_out");
    int rc = ss_csr_degree_chunk(src, dst, n_edges, row_begin, n_rows, src32_out, dst32_out, stats_out, workspace,
                                 workspace_bytes, 1, stream);
    if (rc != SS_OK) return rc;
    return ss_csr_rowptr_finish(n_self_loops, row_begin, n_rows, rowptr, stats_out, workspace, workspace_bytes, stream);
}

// EXPERIMENTAL (see bin_edges_kernel): groups the edges whose destination lies in [row_begin, row_begin + n_rows)
// by destination block into src32_out / dst32_out (int32, rowptr[n_rows] - self loops entries), ready for
// ss_csr_fill(src32 = src32_out, dst32 = dst32_out, n_edges = that count).  workspace: ss_csr_bin_workspace_bytes()
// bytes, ZEROED by the caller.
int64_t ss_csr_bin_workspace_bytes(void) { return (int64_t)ss::BIN_MAX_BUCKETS * 8; }

int ss_csr_bin_edges(const int64_t *src, const int64_t *dst, const int32_t *src32, const int32_t *dst32, int64_t n_edges,
                     int64_t n_self_loops, const int64_t *stats, int64_t row_begin, int64_t n_rows, const int64_t *rowptr,
                     int32_t *src32_out, int32_t *dst32_out, void *workspace, int64_t workspace_bytes, ss_stream_t stream) {
    SS_REQUIRE(n_edges >= 0 && n_rows >= 0 && row_begin >= 0, "negative size passed to ss_csr_bin_edges");
    if (n_edges == 0 || n_rows == 0) return SS_OK;
    SS_REQUIRE(rowptr && workspace && src32_out && dst32_out, "null pointer passed to ss_csr_bin_edges");
    SS_REQUIRE((src && dst) || (src32 && dst32), "src/dst is null");
    SS_REQUIRE(n_self_loops >= 0 || stats, "n_self_loops < 0 (= max id + 1) needs the statistics of ss_csr_rowptr");
    SS_REQUIRE(src32_out != src32 && dst32_out != dst32, "binning is not in place");
    SS_REQUIRE(((uintptr_t)workspace & 7) == 0 && workspace_bytes >= ss_csr_bin_workspace_bytes(),
               "bin workspace must be 8-byte aligned and ss_csr_bin_workspace_bytes() long");
    const int shift = ss::bin_shift_for(n_rows);
    const int n_buckets = (int)((n_rows + ((int64_t)1 << shift) - 1) >> shift);
    int64_t tiles = (n_edges + ss::BIN_TILE - 1) / ss::BIN_TILE;
    int64_t cap = (int64_t)ss::sm_count() * 4;
    const int grid = (int)(tiles < cap ? tiles : cap);
    cudaStream_t st = (cudaStream_t)stream;
    const long long *stp = (const long long *)stats;
    unsigned long long *cur = (unsigned long long *)workspace;
    if (src32 && dst32)
        ss::bin_edges_kernel<int32_t><<<grid, 256, 0, st>>>(src32, dst32, n_edges, n_self_loops, stp, row_begin, n_rows, shift,
                                                            n_buckets, rowptr, cur, src32_out, dst32_out);
    else
        ss::bin_edges_kernel<int64_t><<<grid, 256, 0, st>>>(src, dst, n_edges, n_self_loops, stp, row_begin, n_rows, shift,
                                                            n_buckets, rowptr, cur, src32_out, dst32_out);
    SS_LAUNCH_CHECK("bin_edges_kernel");
    return SS_OK;
}

// ---- streaming CSR of a key-ordered edge list (see sorted_csr_kernel) --------------------------------------------
int ss_csr_sorted_chunk(const int64_t *key, const int64_t *val, int64_t n_edges, int64_t e_base, int64_t n_rows,
                        int add_self_loops, int64_t colidx_capacity, uint64_t fp_key_a, uint64_t fp_key_b, int64_t *rowptr,
                        int32_t *colidx, int64_t *stats_io, int64_t *carry_io, ss_stream_t stream) {
    return ss_csr_sorted_chunk_rows(key, val, n_edges, e_base, 0, 0, n_rows, add_self_loops, colidx_capacity, fp_key_a,
                                    fp_key_b, rowptr, colidx, stats_io, carry_io, stream);
}

int ss_csr_sorted_chunk_rows(const int64_t *key, const int64_t *val, int64_t n_edges, int64_t e_base, int64_t e_origin,
                             int64_t row_begin, int64_t n_rows, int add_self_loops, int64_t colidx_capacity,
                             uint64_t fp_key_a, uint64_t fp_key_b, int64_t *rowptr, int32_t *colidx, int64_t *stats_io,
                             int64_t *carry_io, ss_stream_t stream) {
    SS_REQUIRE(n_edges >= 0 && e_base >= e_origin && e_origin >= 0 && row_begin >= 0 && n_rows >= 0 && colidx_capacity >= 0,
               "bad sizes passed to ss_csr_sorted_chunk");
    SS_REQUIRE(row_begin + n_rows < (1ll << 31), "node ids must be < 2^31");
    SS_REQUIRE(rowptr && colidx && stats_io && carry_io, "null pointer passed to ss_csr_sorted_chunk");
    SS_REQUIRE(n_edges == 0 || (key && val), "key/val is null");
    cudaStream_t st = (cudaStream_t)stream;
    if (e_base == e_origin) {
        const long long init[12] = {-1, 0, 0, 0x7fffffffffffffffll, 0, 0, 0, 0, 0, 0, 0, 0};
        SS_CUDA(cudaMemcpyAsync(stats_io, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    if (n_edges == 0) return SS_OK;
    ss::SortedArgs a;
    a.key = key; a.val = val; a.n = n_edges; a.e_base = e_base; a.e_origin = e_origin; a.row_begin = row_begin;
    a.n_rows = n_rows; a.capacity = colidx_capacity;
    a.loops = add_self_loops ? 1 : 0; a.ka = fp_key_a; a.kb = fp_key_b; a.rowptr = rowptr; a.colidx = colidx;
    a.stats = (long long *)stats_io; a.carry = (long long *)carry_io;
    int64_t blocks = (n_edges + 255) / 256;
    int64_t cap = (int64_t)ss::sm_count() * 8;
    ss::sorted_csr_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(a);
    SS_LAUNCH_CHECK("sorted_csr_kernel");
    return SS_OK;
}

int ss_csr_sorted_finish(int64_t n_edges_total, int64_t n_rows, int add_self_loops, int64_t colidx_capacity, int64_t *rowptr,
                         int32_t *colidx, int64_t *stats_io, const int64_t *carry, ss_stream_t stream) {
    return ss_csr_sorted_finish_rows(n_edges_total, 0, n_rows, add_self_loops, colidx_capacity, rowptr, colidx, stats_io, carry,
                                     stream);
}

int ss_csr_sorted_finish_rows(int64_t n_edges_total, int64_t row_begin, int64_t n_rows, int add_self_loops,
                              int64_t colidx_capacity, int64_t *rowptr, int32_t *colidx, int64_t *stats_io,
                              const int64_t *carry, ss_stream_t stream) {
    SS_REQUIRE(n_edges_total >= 0 && n_rows >= 0 && row_begin >= 0, "negative size passed to ss_csr_sorted_finish");
    SS_REQUIRE(rowptr && colidx && stats_io && carry, "null pointer passed to ss_csr_sorted_finish");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n_rows + 256) / 256;
    int64_t cap = (int64_t)ss::sm_count() * 8;
    ss::sorted_csr_finish_kernel<<<(int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap), 256, 0, st>>>(
        n_edges_total, row_begin, n_rows, colidx_capacity, add_self_loops ? 1 : 0, rowptr, colidx, (long long *)stats_io,
        (const long long *)carry);
    SS_LAUNCH_CHECK("sorted_csr_finish_kernel");
    return SS_OK;
}

int ss_csr_sorted_block(const int64_t *carry, const int64_t *rowptr, int64_t n_rows, int64_t colidx_capacity, const int64_t *stats,
                        int final, const int64_t *prev_block, int64_t *block_out, ss_stream_t stream) {
    SS_REQUIRE(carry && rowptr && stats && block_out && n_rows >= 0, "bad arguments to ss_csr_sorted_block");
    ss::sorted_block_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const long long *)carry, rowptr, n_rows, colidx_capacity,
                                                              (const long long *)stats, final, (const long long *)prev_block,
                                                              (long long *)block_out);
    SS_LAUNCH_CHECK("sorted_block_kernel");
    return SS_OK;
}

int ss_csr_sorted_bounds(const int64_t *key, int64_t n_edges, int64_t n_rows, double row_cost, const double *cum_shares,
                         int n_cuts, int64_t *bounds_out, int64_t *edge_offsets_out, ss_stream_t stream) {
    SS_REQUIRE(n_edges >= 0 && n_rows >= 0 && n_cuts >= 0 && n_cuts <= SS_MAX_PEERS, "bad sizes passed to ss_csr_sorted_bounds");
    SS_REQUIRE(bounds_out && edge_offsets_out && (n_cuts == 0 || cum_shares) && (n_edges == 0 || key),
               "null pointer passed to ss_csr_sorted_bounds");
    ss::sorted_bounds_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(key, n_edges, n_rows, row_cost, cum_shares, n_cuts, bounds_out,
                                                                edge_offsets_out);
    SS_LAUNCH_CHECK("sorted_bounds_kernel");
    return SS_OK;
}

int ss_halo_from_csr(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, int64_t nnz, const int64_t *bounds, int n_ranks,
                     int rank, uint8_t *peer_mask_out, uint8_t *mark_out, ss_stream_t stream) {
    SS_REQUIRE(n_rows >= 0 && n_ranks >= 1 && n_ranks <= SS_MAX_PEERS + 1 && rank >= 0 && rank < n_ranks,
               "bad sizes passed to ss_halo_from_csr");
    if (n_rows == 0) return SS_OK;
    SS_REQUIRE(rowptr && colidx && bounds && peer_mask_out && mark_out, "null pointer passed to ss_halo_from_csr");
    SS_REQUIRE(((uintptr_t)peer_mask_out & 3) == 0, "peer_mask_out must be 4-byte aligned (and padded to a multiple of 4 bytes)");
    SS_REQUIRE(nnz >= 0, "negative nnz");
    if (nnz == 0) return SS_OK;
    int64_t blocks = (nnz + 8 * 1024 - 1) / (8 * 1024);  // 8 warps x one chunk of 1024 positions
    int64_t cap = (int64_t)ss::sm_count() * 8;
    ss::halo_from_csr_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        rowptr, colidx, n_rows, nnz, bounds, n_ranks, rank, (uint32_t *)peer_mask_out, mark_out);
    SS_LAUNCH_CHECK("halo_from_csr_kernel");
    return SS_OK;
}

int ss_mark_rows(const int32_t *colidx, int64_t nnz, uint8_t *mark, ss_stream_t stream) {
    SS_REQUIRE(nnz >= 0, "negative size passed to ss_mark_rows");
    if (nnz == 0) return SS_OK;
    SS_REQUIRE(colidx && mark, "null pointer passed to ss_mark_rows");
    int64_t blocks = (nnz + 255) / 256;
    int64_t cap = (int64_t)ss::sm_count() * 16;
    ss::mark_rows_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(colidx, nnz, mark);
    SS_LAUNCH_CHECK("mark_rows_kernel");
    return SS_OK;
}

// ---- synchronisation-free CSR of an edge list that already holds its self loops (operator forms) -------------------
// What MinhashPropagation / HllPropagation get from ELPH.forward (models/elph.py:186: add_self_loops is applied by the
// caller): nnz = n_edges is known to the host, so nothing has to be read back.  Ids are validated on the device:
// stats_out = { max id, nnz of valid rows, 0, min id }, checked by the host whenever it next synchronises anyway; sources
// outside [0, n_rows) are clamped to 0 and edges whose destination is outside are dropped, colidx is zero-filled first,
// so a bad list can produce wrong sketches (the reference's CUDA scatter would hit a device-side assert) but never an
// out-of-bounds access.  guard (device int32 or NULL): every kernel returns at once when *guard == 0 -- the caller
// re-enqueues the build each time the edge tensor OBJECT changes and lets ss_i64_differs decide on the device whether
// the content did.
namespace ss {
__global__ void __launch_bounds__(256) guarded_zero_kernel(uint32_t *p, int64_t n_words, long long *stats, const int *guard) {
    if (csr_skip(guard)) return;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0u;
    if (stats && blockIdx.x == 0 && threadIdx.x == 0) {
        stats[0] = -1; stats[1] = 0; stats[2] = 0; stats[3] = 0x7fffffffffffffffll;
    }
}

__global__ void __launch_bounds__(256) i64_differs_kernel(const int64_t *__restrict__ a, const int64_t *__restrict__ b,
                                                           int64_t n, int *flag) {
    bool diff = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        diff |= (a[i] != b[i]);
    if (__any_sync(FULL, diff) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}
}  // namespace ss

int ss_i64_differs(const int64_t *a, const int64_t *b, int64_t count, int32_t *flag_out, ss_stream_t stream) {
    SS_REQUIRE(count >= 0 && flag_out, "bad arguments to ss_i64_differs");
    cudaStream_t st = (cudaStream_t)stream;
    SS_CUDA(cudaMemsetAsync(flag_out, 0, sizeof(int32_t), st));
    if (count == 0) return SS_OK;
    SS_REQUIRE(a && b, "null pointer passed to ss_i64_differs");
    int64_t blocks = (count + 1023) / 1024;
    int64_t cap = (int64_t)ss::sm_count() * 8;
    ss::i64_differs_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(a, b, count, flag_out);
    SS_LAUNCH_CHECK("i64_differs_kernel");
    return SS_OK;
}

int ss_csr_build_nosync(const int64_t *src, const int64_t *dst, int64_t n_edges, int64_t n_rows, int64_t *rowptr,
                        int32_t *colidx, int64_t *stats_out, void *workspace, int64_t workspace_bytes, const int32_t *guard,
                        ss_stream_t stream) {
    SS_REQUIRE(n_edges >= 0 && n_rows >= 0, "negative size passed to ss_csr_build_nosync");
    SS_REQUIRE(rowptr && stats_out && workspace && (n_edges == 0 || (src && dst && colidx)), "null pointer passed to ss_csr_build_nosync");
    SS_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    if (workspace_bytes < ss::workspace_bytes(n_rows)) {
        ss::set_error("csr workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)ss::workspace_bytes(n_rows));
        return SS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    long long *stats = (long long *)stats_out;
    ss::CsrWorkspace w = ss::carve(workspace, n_rows);
    const int64_t cap = (int64_t)ss::sm_count() * 32;
    auto grid_for = [&](int64_t items) { int64_t b = (items + 255) / 256; return (int)(b < cap ? (b < 1 ? 1 : b) : cap); };
    ss::guarded_zero_kernel<<<grid_for(n_rows), 256, 0, st>>>(w.deg, n_rows, stats, guard);
    SS_LAUNCH_CHECK("guarded_zero_kernel");
    if (n_edges > 0) {
        ss::guarded_zero_kernel<<<grid_for(n_edges), 256, 0, st>>>((uint32_t *)colidx, n_edges, nullptr, guard);
        SS_LAUNCH_CHECK("guarded_zero_kernel");
        ss::degree_kernel<<<grid_for(n_edges), 256, 0, st>>>(src, dst, n_edges, 0, n_rows, w.deg, nullptr, nullptr, stats, guard);
        SS_LAUNCH_CHECK("degree_kernel");
    }
    if (n_rows == 0) {
        SS_CUDA(cudaMemsetAsync(rowptr, 0, 8, st));
        return SS_OK;
    }
    ss::tile_sum_kernel<<<(int)w.n_tiles, ss::SCAN_BLOCK, 0, st>>>(w.deg, n_rows, 0, 0, stats, w.tile_sum, guard);
    SS_LAUNCH_CHECK("tile_sum_kernel");
    ss::tile_scan_kernel<<<1, 1024, 0, st>>>(w.tile_sum, w.n_tiles, guard);
    SS_LAUNCH_CHECK("tile_scan_kernel");
    ss::rowptr_kernel<<<(int)w.n_tiles, ss::SCAN_BLOCK, 0, st>>>(w.deg, n_rows, 0, 0, stats, w.tile_sum, rowptr, guard);
    SS_LAUNCH_CHECK("rowptr_kernel");
    ss::cursor_init_kernel<<<grid_for(n_rows), 256, 0, st>>>(rowptr, n_rows, w.deg, guard);
    SS_LAUNCH_CHECK("cursor_init_kernel");
    if (n_edges > 0) {
        ss::fill_kernel<int64_t, true><<<grid_for(n_edges + n_rows), 256, 0, st>>>(src, dst, n_edges, 0, stats, 0, n_rows, rowptr,
                                                                                w.deg, colidx, guard, n_rows);
        SS_LAUNCH_CHECK("fill_kernel");
    }
    return SS_OK;
}

int ss_csr_fill(const int64_t *src, const int64_t *dst, const int32_t *src32, const int32_t *dst32, int64_t n_edges,
                int64_t n_self_loops, const int64_t *stats, int64_t row_begin, int64_t n_rows, const int64_t *rowptr,
                int32_t *colidx, void *workspace, int64_t workspace_bytes, ss_stream_t stream) {
    SS_REQUIRE(n_edges >= 0 && n_rows >= 0 && row_begin >= 0, "negative size passed to ss_csr_fill");
    SS_REQUIRE(rowptr && workspace, "null pointer passed to ss_csr_fill");
    SS_REQUIRE(n_edges == 0 || (src && dst) || (src32 && dst32), "src/dst is null");
    SS_REQUIRE(n_self_loops >= 0 || stats, "n_self_loops < 0 (= max id + 1) needs the statistics of ss_csr_rowptr");
    SS_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    if (workspace_bytes < ss::workspace_bytes(n_rows)) {
        ss::set_error("csr workspace too small: %lld < %lld", (long long)workspace_bytes,
                      (long long)ss::workspace_bytes(n_rows));
        return SS_ERR_WORKSPACE;
    }
    if (n_rows == 0) return SS_OK;
    SS_REQUIRE(colidx, "colidx is null");
    cudaStream_t st = (cudaStream_t)stream;
    ss::CsrWorkspace w = ss::carve(workspace, n_rows);
    int64_t total = n_edges + n_rows;
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)ss::sm_count() * 32;
    int grid = (int)(blocks < cap ? blocks : cap);
    const bool abs_cursor = ss::abs_cursor_fill();
    if (abs_cursor) {
        int64_t cb = (n_rows + 255) / 256;
        ss::cursor_init_kernel<<<(int)(cb < cap ? cb : cap), 256, 0, st>>>(rowptr, n_rows, w.deg);
        SS_LAUNCH_CHECK("cursor_init_kernel");
    } else {
        SS_CUDA(cudaMemsetAsync(w.deg, 0, (size_t)n_rows * 4, st));
    }
    const long long *stp = (const long long *)stats;
    if (src32 && dst32) {
        if (abs_cursor)
            ss::fill_kernel<int32_t, true><<<grid, 256, 0, st>>>(src32, dst32, n_edges, n_self_loops, stp, row_begin, n_rows,
                                                                 rowptr, w.deg, colidx);
        else
            ss::fill_kernel<int32_t, false><<<grid, 256, 0, st>>>(src32, dst32, n_edges, n_self_loops, stp, row_begin, n_rows,
                                                                  rowptr, w.deg, colidx);
    } else {
        if (abs_cursor)
            ss::fill_kernel<int64_t, true><<<grid, 256, 0, st>>>(src, dst, n_edges, n_self_loops, stp, row_begin, n_rows,
                                                                 rowptr, w.deg, colidx);
        else
            ss::fill_kernel<int64_t, false><<<grid, 256, 0, st>>>(src, dst, n_edges, n_self_loops, stp, row_begin, n_rows,
                                                                  rowptr, w.deg, colidx);
    }
    SS_LAUNCH_CHECK("fill_kernel");
    return SS_OK;
}

}  // extern "C"
