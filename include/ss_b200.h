/*
 * ss_b200.h -- C ABI of the B200-native subgraph-sketch engine (libss_b200.so, sm_100a).
 *
 * The reference (melifluos/subgraph-sketching) is pure Python; it has no FFI.  The boundary this
 * library serves is the operator surface of `ElphHashes` in /root/reference/src/hashing.py: each entry
 * point below names the reference function(s) it replaces.  A Python host binds these with `ctypes`
 * (see INTEGRATION.md and subgraph_sketching_b200/_lib.py).
 *
 * Conventions
 *   - every function is `extern "C"`, takes plain pointers / sizes, returns 0 on success, <0 on error;
 *     `ss_last_error()` returns a thread-local message for the last failure on the calling thread
 *   - all array pointers are DEVICE pointers (CUDA unified addressing) unless the name says `host`
 *   - the caller owns every buffer, including workspaces; nothing is allocated or freed inside
 *   - `stream` is a `cudaStream_t` passed as void*; calls only enqueue work, they never synchronise
 *   - node ids are int64 in link lists / COO edges (as in the reference) and int32 inside the CSR
 *
 * Sketch record (the HBM layout, one per node per hop):
 *     [ P_pad x uint32 MinHash | 2^p x uint8 HLL registers ]   P_pad = P rounded up to 4
 *   = `ss_record_bytes(P, p)` bytes (768 B at the reference defaults P=128, p=8), 16-byte aligned.
 *   The reference stores MinHash as int64 (values are < 2^32, hashing.py:59,122) and HLL as int8:
 *   1280 B per node per hop.  `ss_pack_records` / `ss_unpack_records` convert between the two.
 *   Record tables are addressed as (base pointer, row stride in bytes); stride >= record bytes and a
 *   multiple of 16, so per-hop tables and node-major interleaved tables use the same entry points.
 */
#ifndef SS_B200_H
#define SS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SS_ABI_VERSION 14

#define SS_OK 0
#define SS_ERR_INVALID (-1)     /* bad argument (null pointer, unsupported P/p/K, misaligned buffer) */
#define SS_ERR_CUDA (-2)        /* a CUDA runtime call or kernel launch failed */
#define SS_ERR_WORKSPACE (-3)   /* caller-provided workspace too small */

/* flags for ss_link_features */
#define SS_FLAG_USE_ZERO_ONE 1  /* keep the (0,1)/(1,0) [and K=3: (0,2)/(2,0)] columns; hashing.py:310-318 */
#define SS_FLAG_FLOOR 2         /* clamp negative features to 0; hashing.py:319-320 */

/* merge kernel variants (ss_khop_merge `variant`) */
#define SS_MERGE_AUTO 0
#define SS_MERGE_TMA 1          /* TMA row gather (cp.async.bulk.tensor tile::gather4) into a shared-memory ring (P=128, p=8) */
#define SS_MERGE_LDG 2          /* direct 128-bit global loads, register accumulators (P=128, p=8) */
#define SS_MERGE_GENERIC 3      /* any (P, p): column-chunk outer loop, no staging */
#define SS_MERGE_BULK 4         /* one 1-D cp.async.bulk per row into the same ring (P=128, p=8) */

typedef void *ss_stream_t;
#define SS_MAX_PEERS 7   /* peers of one GPU in the fused multi-GPU exchange (8 GPUs per box) */

/*
 * HyperLogLog++ constants for one precision p.  They are INPUTS (the reference takes them from
 * datasketch, hashing.py:69-80).  The struct lives in host memory; the three table pointers are device
 * pointers.  `lc_table[z]` = linear-counting estimate for z empty registers, z = 0..m (entry 0 unused),
 * computed by the host with the same float32 ops as hashing.py:194-195 so the LC regime is bit-exact.
 */
typedef struct ss_hll_consts {
    int32_t p;             /* precision; m = 1 << p registers */
    int32_t table_len;     /* T = length of raw_estimate / bias (>= 6) */
    int32_t monotone;      /* 1 if raw_estimate is non-decreasing (enables the binary-search 6-NN) */
    float threshold;       /* hashing.py:77 (float32 of the integer threshold) */
    float alpha_m2;        /* float32(alpha * m * m), hashing.py:228 */
    float five_m;          /* float32(5 * m), hashing.py:207 */
    const float *lc_table; /* device, [m + 1] */
    const float *raw_estimate; /* device, [T]  (hashing.py:80) */
    const float *bias;         /* device, [T]  (hashing.py:79) */
} ss_hll_consts;

/* one hop of a sketch table as seen by the pairwise kernel */
typedef struct ss_hop_view {
    const void *records;   /* device; compact records of all nodes for this hop */
    int64_t row_stride;    /* bytes between consecutive nodes */
    int64_t num_rows;      /* nodes in the table (every link endpoint must be < num_rows) */
} ss_hop_view;

/* ---- library ------------------------------------------------------------------------------- */
int ss_version(void);
const char *ss_last_error(void);
/* SM count and compute capability of the current device */
int ss_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* bytes of one compact record; <0 if (num_perm, hll_p) is unsupported (need 1<=P<=4096, 4<=p<=18) */
int64_t ss_record_bytes(int num_perm, int hll_p);

/* ---- K1: hop-0 sketches ---------------------------------------------------------------------
 * Replaces ElphHashes.initialise_minhash (hashing.py:118-124) and initialise_hll (hashing.py:126-137)
 * for node ids first_id .. first_id+n-1 (the reference hashes ids 1..N, so first_id = 1 + row offset).
 * perm_a/perm_b: device uint64 [P], the host-side RandomState(1) draws (hashing.py:106-116).
 * log2_window:   device int32 [64]; entry k = largest j >= 0 such that the reference's float64
 *                ceil(log2(2^k + j)) still evaluates to k (hashing.py:83-89), computed on the host with
 *                numpy so the device reproduces the reference's bit_length quirk exactly.
 * Returns SS_ERR_INVALID semantics of hashing.py:101-103 are impossible here: rank >= 1 always holds for
 * 64-bit hashes, so no overflow error exists.
 */
int ss_init_records(int64_t n, int64_t first_id, int num_perm, int hll_p, const uint64_t *perm_a,
                    const uint64_t *perm_b, const int32_t *log2_window, void *rec_out, int64_t out_stride,
                    ss_stream_t stream);

/* reference layout (int64 [n,P] MinHash, int8 [n,m] HLL)  <->  compact records; either side of the
 * pair may be NULL to skip it.  These back `hash_table[k]['minhash']` / `['hll']` (hashing.py:153-154). */
int ss_pack_records(const int64_t *minhash, const int8_t *hll, int64_t n, int num_perm, int hll_p,
                    void *rec_out, int64_t out_stride, ss_stream_t stream);
int ss_unpack_records(const void *rec, int64_t rec_stride, int64_t n, int num_perm, int hll_p,
                      int64_t *minhash_out, int8_t *hll_out, ss_stream_t stream);
/* the same with
 *   error_flag (device int32[2] initialised to {0, 1} by the caller, or NULL): when a value does not fit a record (MinHash
 *              outside [0, 2^32), register outside [0, 127]) [0] is set to 1 and [1] cleared to 0.  A host that did not
 *              synchronise to inspect the tensor first enqueues the record engine guarded by &flag[1] and the plain
 *              int64 kernel (ss_prop_min_i64_guarded) guarded by &flag[0]: exactly one of them runs (the reference's
 *              sketches always fit: hashing.py:59,122, :132-136)
 *   guard      (device int32 or NULL): the kernel returns at once when *guard == 0 (see ss_khop_merge_ex)
 * and, when only one side of the pair is given, a row stride as narrow as that side (512-byte MinHash-only tables). */
int ss_pack_records_ex(const int64_t *minhash, const int8_t *hll, int64_t n, int num_perm, int hll_p, void *rec_out,
                       int64_t out_stride, int32_t *error_flag, const int32_t *guard, ss_stream_t stream);
int ss_unpack_records_ex(const void *rec, int64_t rec_stride, int64_t n, int num_perm, int hll_p, int64_t *minhash_out,
                         int8_t *hll_out, const int32_t *guard, ss_stream_t stream);

/* ---- K6: COO -> CSR keyed by destination ------------------------------------------------------
 * The reference scatters over the COO edge_index with self loops appended by
 * add_self_loops(edge_index) (hashing.py:148; PyG flow source -> target, aggregation at edge_index[1]).
 * Rows are destinations in [row_begin, row_begin + n_rows); a self loop (i, i) is added for every
 * i < n_self_loops that falls in the row range; n_self_loops < 0 means max(edge_index) + 1, computed on the
 * device by ss_csr_rowptr -- exactly add_self_loops without num_nodes.  Two steps because nnz is data dependent:
 *   ss_csr_rowptr : rowptr[0..n_rows] (int64, exclusive scan of in-degrees incl. self loops) and, when
 *                   stats_out != NULL, stats_out[0..3] = { max id (-1 if no edges), nnz, self loops used, min id }
 *                   (device int64[4]) plus optional int32 device copies src32_out / dst32_out of the endpoints
 *   ss_csr_fill   : colidx[0..nnz) (int32 global source ids), order inside a row is unspecified
 *                   (min/max merges are order independent); reads the int32 copies when given
 * src / dst are read with coalesced loads only, so they may be PINNED HOST pointers (UVA): the edge list is
 * then consumed at PCIe rate with no staging copy, and with the int32 copies it crosses the bus once.
 * ss_csr_rowptr = ss_csr_degree_chunk(first = 1) over the whole list + ss_csr_rowptr_finish.  The two halves are
 * exported for STREAMED ingestion: a host edge list is copied in chunks by the DMA engine (55 GB/s measured, against
 * 43-48 GB/s for in-place reads by the SMs) into a small device staging ring and each chunk is histogrammed as it
 * lands; src32_out / dst32_out then point at the chunk's slice of the int32 copies that ss_csr_fill reads.
 * workspace: ss_csr_workspace_bytes(n_rows) bytes, 256-byte aligned, the same buffer for all calls.
 */
int64_t ss_csr_workspace_bytes(int64_t n_rows);
int ss_csr_degree_chunk(const int64_t *src, const int64_t *dst, int64_t n_edges, int64_t row_begin, int64_t n_rows,
                        int32_t *src32_out, int32_t *dst32_out, int64_t *stats_io, void *workspace, int64_t workspace_bytes,
                        int first, ss_stream_t stream);
int ss_csr_rowptr_finish(int64_t n_self_loops, int64_t row_begin, int64_t n_rows, int64_t *rowptr, int64_t *stats_io,
                         void *workspace, int64_t workspace_bytes, ss_stream_t stream);
int ss_csr_rowptr(const int64_t *src, const int64_t *dst, int64_t n_edges, int64_t n_self_loops,
                  int64_t row_begin, int64_t n_rows, int64_t *rowptr, int32_t *src32_out, int32_t *dst32_out,
                  int64_t *stats_out, void *workspace, int64_t workspace_bytes, ss_stream_t stream);
int ss_csr_fill(const int64_t *src, const int64_t *dst, const int32_t *src32, const int32_t *dst32, int64_t n_edges,
                int64_t n_self_loops, const int64_t *stats, int64_t row_begin, int64_t n_rows, const int64_t *rowptr,
                int32_t *colidx, void *workspace, int64_t workspace_bytes, ss_stream_t stream);
/* Streaming CSR for edge lists that are ALREADY ORDERED by the CSR key -- what PyG's coalesce / to_undirected hand the
 * reference (datasets/elph.py:63-66, hashing.py:148): one coalesced pass, no histogram, no atomics on rows, no scattered
 * stores.  key[] is expected non-decreasing:
 *   key = edge_index[1], val = edge_index[0] : the exact destination-keyed CSR
 *   key = edge_index[0], val = edge_index[1] : the source-keyed CSR, equal to the destination-keyed one iff the edge
 *       multiset is symmetric; the pass accumulates keyed 2 x 64-bit multiset fingerprints of (key, val) and (val, key)
 * The pass is SPECULATIVE: it never fails, it reports.  stats_io (device int64[12]) =
 *   { max id, nnz, self loops, min id, fp(key,val) a, b, fp(val,key) a, b, key-order violations (0 = ordered),
 *     val-order violations, range errors (id < 0, key >= n_rows, val >= 2^31), reserved }
 * and the caller accepts rowptr / colidx only if [8] == 0, [10] == 0, max id < n_rows and (for key = edge_index[0])
 * [4..5] == [6..7]; otherwise it falls back to ss_csr_rowptr + ss_csr_fill.  Every write is bounds-guarded
 * (colidx_capacity entries, n_rows + 1 rowptr entries) and structural work stops at the first violation anywhere.
 * With add_self_loops the loop of node r (every r <= max id, add_self_loops without num_nodes) is stored FIRST in
 * row r.  A list may be fed in consecutive chunks (e_base = global index of key[0]; carry_io, device int64[2], carries
 * the last edge across chunks; the chunk with e_base == 0 initialises stats_io); ss_csr_sorted_finish(total) completes
 * rowptr above the last key and writes nnz / self loops into stats_io[1..2]. */
int ss_csr_sorted_chunk(const int64_t *key, const int64_t *val, int64_t n_edges, int64_t e_base, int64_t n_rows,
                        int add_self_loops, int64_t colidx_capacity, uint64_t fp_key_a, uint64_t fp_key_b, int64_t *rowptr,
                        int32_t *colidx, int64_t *stats_io, int64_t *carry_io, ss_stream_t stream);
int ss_csr_sorted_finish(int64_t n_edges_total, int64_t n_rows, int add_self_loops, int64_t colidx_capacity, int64_t *rowptr,
                         int32_t *colidx, int64_t *stats_io, const int64_t *carry, ss_stream_t stream);
/* Row-block forms for the node-sharded build: in a list ordered by its key the edges of the rows
 * [row_begin, row_begin + n_rows) are ONE contiguous slice [e_origin, e_origin + n_edges_total) of the list, so every rank
 * streams only its own slice (also over PCIe when the list is in pinned host memory) -- no histogram all-reduce, no
 * pass over the whole list.  rowptr (n_rows + 1 entries) and colidx positions are local to the block; stats_io[0] (max id)
 * must hold the GLOBAL maximum when ss_csr_sorted_finish_rows runs (all-reduce it between the two calls), and the
 * fingerprints / violation counters are summed over ranks before the verdict.
 * ss_csr_sorted_bounds cuts the rows into cost-balanced blocks WITHOUT a histogram: cost(r) = (#edges with key < r) +
 * row_cost * r, cut q at cum_shares[q] (device double [n_cuts], increasing, in (0, 1)) of the total; it writes
 * bounds_out[0..n_cuts+1] (rows) and edge_offsets_out[0..n_cuts+1] (first edge of each block), device int64.
 * ss_mark_rows: mark[colidx[e]] = 1 -- the rows a rank's neighbour lists read, from which the owners derive the
 * peer_mask of ss_khop_merge_ex. */
int ss_csr_sorted_chunk_rows(const int64_t *key, const int64_t *val, int64_t n_edges, int64_t e_base, int64_t e_origin,
                             int64_t row_begin, int64_t n_rows, int add_self_loops, int64_t colidx_capacity,
                             uint64_t fp_key_a, uint64_t fp_key_b, int64_t *rowptr, int32_t *colidx, int64_t *stats_io,
                             int64_t *carry_io, ss_stream_t stream);
int ss_csr_sorted_finish_rows(int64_t n_edges_total, int64_t row_begin, int64_t n_rows, int add_self_loops,
                              int64_t colidx_capacity, int64_t *rowptr, int32_t *colidx, int64_t *stats_io,
                              const int64_t *carry, ss_stream_t stream);
int ss_csr_sorted_bounds(const int64_t *key, int64_t n_edges, int64_t n_rows, double row_cost, const double *cum_shares,
                         int n_cuts, int64_t *bounds_out, int64_t *edge_offsets_out, ss_stream_t stream);
/* ss_csr_sorted_block: after a chunk has been absorbed, block_out (device int64[4]) = { row_begin, row_end, pos_begin, pos_end }
 * of the rows that are complete now and were not in prev_block (NULL for the first): every row below the chunk's last key;
 * final != 0: everything up to n_rows (call it after ss_csr_sorted_finish).  The block is empty once the pass has seen an
 * order violation or a range error.  Feeds the `block` of ss_khop_merge_ex: hop 1 runs under the ingest stream. */
int ss_csr_sorted_block(const int64_t *carry, const int64_t *rowptr, int64_t n_rows, int64_t colidx_capacity, const int64_t *stats,
                        int final, const int64_t *prev_block, int64_t *block_out, ss_stream_t stream);
int ss_mark_rows(const int32_t *colidx, int64_t nnz, uint8_t *mark, ss_stream_t stream);
/* The halo of a SYMMETRIC graph needs no exchange: row r of this rank is read by rank q exactly when r has an in-neighbour
 * owned by q.  From the rank's own CSR rows: peer_mask_out[r] (bit i = the i-th OTHER rank in ascending order reads row r)
 * and mark_out[c] = 1 for every neighbour c (the rows this rank reads).  bounds: device int64 [n_ranks + 1]; peer_mask_out:
 * ZEROED by the caller, 4-byte aligned and padded to a multiple of 4 bytes (rows of one 32-bit word are updated atomically). */
int ss_halo_from_csr(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, int64_t nnz, const int64_t *bounds, int n_ranks,
                     int rank, uint8_t *peer_mask_out, uint8_t *mark_out, ss_stream_t stream);
/* Synchronisation-free CSR for the operator forms (the edge list already holds its self loops: ELPH.forward applies
 * add_self_loops itself, models/elph.py:186, and calls hll_prop / minhash_prop 2K times per training batch): nnz = n_edges
 * is known to the host, so nothing is read back.  Ids are validated on the device -- stats_out (device int64[4]) =
 * { max id, entries kept, 0, min id }, to be inspected whenever the host next synchronises; sources outside [0, n_rows)
 * are clamped to 0, edges with a destination outside are dropped, colidx (n_edges entries) is zero-filled first: a bad
 * list gives wrong sketches (the reference's CUDA scatter would trip a device-side assert) but never an out-of-bounds
 * access.  guard (device int32 or NULL): every kernel returns at once when *guard == 0.
 * ss_i64_differs: flag_out = (a[0..count) != b[0..count)) -- lets the device decide whether a new edge tensor OBJECT also
 * has new CONTENT, so a cached CSR (and everything memoised on it) is revalidated without a host round trip. */
int ss_csr_build_nosync(const int64_t *src, const int64_t *dst, int64_t n_edges, int64_t n_rows, int64_t *rowptr,
                        int32_t *colidx, int64_t *stats_out, void *workspace, int64_t workspace_bytes, const int32_t *guard,
                        ss_stream_t stream);
int ss_i64_differs(const int64_t *a, const int64_t *b, int64_t count, int32_t *flag_out, ss_stream_t stream);
/* EXPERIMENTAL, opt-in (SS_B200_CSR_BIN=1 in the Python host), not on the default path, MEASURED SLOWER (profiles/r02_experiments_call_a.txt: 29.3 vs 22.5 ms): one streaming pass that groups
 * the edges by destination block (dst >> shift, at most 2048 blocks; capacities = differences of rowptr) into
 * src32_out / dst32_out so that the fill walks colidx window by window and completes its 32-byte sectors in L2.
 * Writes rowptr[n_rows] - (self loops in range) entries; then call ss_csr_fill with src32 = src32_out,
 * dst32 = dst32_out and that count as n_edges.  workspace: ss_csr_bin_workspace_bytes() bytes, zeroed by the caller. */
int64_t ss_csr_bin_workspace_bytes(void);
int ss_csr_bin_edges(const int64_t *src, const int64_t *dst, const int32_t *src32, const int32_t *dst32, int64_t n_edges,
                     int64_t n_self_loops, const int64_t *stats, int64_t row_begin, int64_t n_rows, const int64_t *rowptr,
                     int32_t *src32_out, int32_t *dst32_out, void *workspace, int64_t workspace_bytes, ss_stream_t stream);

/* ---- K2: one hop of sketch propagation ----------------------------------------------------------
 * Replaces MinhashPropagation.forward + HllPropagation.forward (hashing.py:28-45, called at :160-162)
 * and, when cards_out != NULL, the hll_count of the merged rows (hashing.py:163).
 *   rec_out[r] = ( min over c in colidx[rowptr[r]..rowptr[r+1]) of rec_in[c].minhash,
 *                  max ...                                      of rec_in[c].hll )
 *   rows with no in-edge are all-zero (scatter-max fill).
 * rec_in is the full previous-hop table (in_rows records indexed by global id; every colidx entry must be
 * < in_rows), rec_out holds the n_rows owned rows
 * (rowptr is local to them: rowptr[0] = 0, rowptr[n_rows] = nnz); rec_in and rec_out must not overlap.
 * cards_out[r * cards_stride] receives the float32 HLL++ estimate of row r (stride in elements).
 * workspace: ss_merge_workspace_bytes(nnz, P, p) bytes, 16-byte aligned.  The neighbour list is cut into
 * equal ranges; rows cut by a range boundary leave partial records there and a fix-up launch folds them,
 * so power-law hubs cost no more per neighbour than any other row.
 */
int64_t ss_merge_workspace_bytes(int64_t nnz, int num_perm, int hll_p);
int ss_khop_merge(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, int64_t nnz, const void *rec_in,
                  int64_t in_rows, int64_t in_stride, void *rec_out, int64_t out_stride, int num_perm, int hll_p, void *workspace,
                  int64_t workspace_bytes, float *cards_out, int64_t cards_stride, const ss_hll_consts *hc,
                  int variant, ss_stream_t stream);

/* Fused merge + exchange for the node-sharded multi-GPU build: the same launch, and every
 * finished row (and its cardinality) is additionally stored into `n_peers` peer copies of rec_out / cards_out
 * mapped into this process over NVLink (CUDA IPC / symmetric memory).  peer_rec_out[i] / peer_cards_out[i]
 * are HOST arrays of device pointers addressed exactly like rec_out / cards_out (row 0 = first owned row).
 * When every rank's launch has completed (stream-ordered barrier between ranks), every rank holds the whole
 * next-hop table: the per-hop all-gather of the sketches happens inside the kernel, overlapped row by row.
 * mc_rec_out / mc_cards_out (optional, else NULL): NVSwitch MULTICAST addresses of the same buffers
 * (cuMulticast / symmetric memory `multicast_ptr`), addressed like rec_out / cards_out; when given, every
 * store is ONE `multimem.st` that the switch replicates into all GPUs' copies (the writer's included) and the
 * peer arrays are ignored.  P=128 / p=8 engines only. */
int ss_khop_merge_peers(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, int64_t nnz, const void *rec_in,
                        int64_t in_rows, int64_t in_stride, void *rec_out, int64_t out_stride, int num_perm, int hll_p,
                        void *workspace, int64_t workspace_bytes, float *cards_out, int64_t cards_stride,
                        const ss_hll_consts *hc, int variant, int n_peers, void *const *peer_rec_out,
                        float *const *peer_cards_out, void *mc_rec_out, float *mc_cards_out, ss_stream_t stream);

/* Layout-general form of the merge (everything ss_khop_merge / ss_khop_merge_peers do, plus):
 *   layout      what a row of rec_in / rec_out holds (P=128, p=8 engines, TMA variant):
 *                 SS_LAYOUT_FULL     768 B  [128 x u32 MinHash | 256 x u8 HLL]        the hop tables
 *                 SS_LAYOUT_MINHASH  512 B  [128 x u32 MinHash]                       MinhashPropagation called alone
 *                 SS_LAYOUT_HLL      256 B  [256 x i8 registers]  = the reference's int8 [N, 256] tensor AS IS
 *                                           (signed max, exact for any int8 content)  HllPropagation called alone
 *                 SS_LAYOUT_HALF     384 B  [64 x u32 MinHash | 128 x u8 HLL]          one column half of a record
 *               cards_out needs all 256 registers of a row: FULL and HLL only
 *   peer_mask   device uint8 [n_rows] or NULL: bit i set = peer i reads output row r at some point (it owns a
 *               destination with r as in-neighbour); rows whose bit is clear are not stored to that peer ("halo push":
 *               on R-MAT-24 over 8 GPUs only ~30 % of the (row, peer) pairs are ever read).  Cards are always replicated.
 *   guard       device int32 or NULL: when *guard == 0 at launch time the kernels return at once.  Lets a memoised
 *               pipeline (ELPH re-propagating an unchanged graph every batch, models/elph.py:180-218) be re-enqueued
 *               without a host synchronisation deciding whether the cached result is still valid.
 *   block       "blocked launch": device int64[4] = { row_begin, row_end, pos_begin, pos_end } read at kernel start, so the
 *               host can enqueue the merge of a row block BEFORE it knows the block (hop 1 of the rows whose edges have
 *               already arrived, under the PCIe stream of a host edge list: ss_csr_sorted_block writes the descriptor).
 *               rowptr / colidx / rec_out / cards_out are then the WHOLE graph's arrays (n_rows rows in total, nnz = an
 *               upper bound of the neighbour positions, e.g. the colidx capacity), rowptr[row_begin] == pos_begin,
 *               rowptr[row_end] == pos_end.  Full records, TMA engine, no peers; the workspace must hold
 *               ss_merge_workspace_bytes(nnz) + 4 records. */
#define SS_LAYOUT_FULL 0
#define SS_LAYOUT_MINHASH 1
#define SS_LAYOUT_HLL 2
#define SS_LAYOUT_HALF 3
typedef struct ss_merge_desc {
    const int64_t *rowptr;
    const int32_t *colidx;
    int64_t n_rows, nnz;
    const void *rec_in;
    int64_t in_rows, in_stride;
    void *rec_out;
    int64_t out_stride;
    int32_t num_perm, hll_p, layout, variant;
    void *workspace;
    int64_t workspace_bytes;
    float *cards_out;
    int64_t cards_stride;
    const ss_hll_consts *hc;
    int32_t n_peers, reserved;
    void *const *peer_rec_out;
    float *const *peer_cards_out;
    const uint8_t *peer_mask;
    void *mc_rec_out;
    float *mc_cards_out;
    const int32_t *guard;
    const int64_t *block;   /* see "blocked launches" above; NULL = the whole problem, sized by n_rows / nnz */
} ss_merge_desc;
int ss_khop_merge_ex(const ss_merge_desc *desc, ss_stream_t stream);

/* operator forms on the reference's own tensor layouts (ELPH calls these per batch,
 * /root/reference/src/models/elph.py:209-212): element-wise signed min (int64) / max (int8) over
 * in-neighbours; rows with no in-edge are 0.  colidx == NULL means neighbour j of row r is row
 * rowptr[r] + j of x (used for hll_neighbour_merge / minhash_neighbour_merge, hashing.py:239-245). */
int ss_prop_min_i64(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, const int64_t *x,
                    int64_t *out, int64_t width, ss_stream_t stream);
int ss_prop_max_i8(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, const int8_t *x, int8_t *out,
                   int64_t width, ss_stream_t stream);
/* ss_prop_min_i64 that returns at once when *guard == 0 (guard: device int32 or NULL; see ss_pack_records_ex) */
int ss_prop_min_i64_guarded(const int64_t *rowptr, const int32_t *colidx, int64_t n_rows, const int64_t *x, int64_t *out,
                            int64_t width, const int32_t *guard, ss_stream_t stream);

/* ---- K3: HyperLogLog++ cardinality ---------------------------------------------------------------
 * Replaces ElphHashes.hll_count (hashing.py:212-232) with _linearcounting (:194-195),
 * _estimate_bias (:197-204) and _refine_hll_count_estimate (:206-210).
 * regs: uint8 rows of m registers, `row_stride` bytes apart (8-byte aligned); out[i * out_stride] float32.
 */
int ss_hll_count(const void *regs, int64_t row_stride, int64_t n, const ss_hll_consts *hc, float *out,
                 int64_t out_stride, ss_stream_t stream);
/* _estimate_bias alone (hashing.py:197-204): out[i] = mean bias of the 6 nearest raw estimates to e[i] */
int ss_estimate_bias(const float *e, int64_t n, const ss_hll_consts *hc, float *out, ss_stream_t stream);

/* ---- K5: small row-wise helpers ---------------------------------------------------------------------
 * jaccard (hashing.py:247-256): out[i] = count(src[i,:] == dst[i,:]) / denom, int64 rows of `width` slots
 *   (the reference divides by num_perm whatever the width).
 * hll merge (hashing.py:234-237): element-wise max of two int8 arrays of `count` bytes. */
int ss_jaccard_i64(const int64_t *src, const int64_t *dst, int64_t n, int64_t width, int64_t denom, float *out,
                   ss_stream_t stream);
int ss_max_i8(const int8_t *a, const int8_t *b, int64_t count, int8_t *out, ss_stream_t stream);

/* ---- K4: pairwise structural features ---------------------------------------------------------------
 * Replaces ElphHashes.get_subgraph_features (hashing.py:258-323) incl. _get_intersections (:167-189),
 * jaccard (:247-256) and _hll_merge (:234-237).
 *   links        int64 [L, 2] (u, v) global node ids
 *   hops         HOST array of K+1 views indexed by hop (entry 0 is ignored)
 *   cards        float32, cards[node * cards_stride + (k-1)] = k-hop cardinality (hashing.py:163)
 *   features_out float32 [L, K(K+2)] or NULL;  inter_out float32 [L, K*K] (row-major (k1-1)*K+(k2-1)) or NULL
 *   error_flag   device int32 or NULL: set to 1 when a link endpoint is outside [0, num_rows) (torch indexing in the
 *                reference raises IndexError); such links are evaluated on node 0, never out of bounds.
 *   links may be a PINNED HOST pointer (coalesced 16-byte reads over PCIe, 0.3 % of the kernel's traffic).
 */
int ss_link_features(const int64_t *links, int64_t n_links, const ss_hop_view *hops, int max_hops,
                     int num_perm, int hll_p, const float *cards, int64_t cards_stride,
                     const ss_hll_consts *hc, int flags, float *features_out, float *inter_out,
                     int32_t *error_flag, ss_stream_t stream);

/* The same over NODE-SHARDED tables (multi-GPU): this rank's hop tables hold its own row block plus the halo
 * rows it gathered while building (local_rows[x] != 0); every other record is read from its OWNER's copy, mapped into
 * this process over NVLink (symmetric memory / CUDA IPC), so the last hop -- which no later hop gathers from -- is never
 * replicated at all and hops 1..K-1 only where a neighbour list needed them.  All copies share the local row stride.
 * shard == NULL or n_ranks == 1: exactly ss_link_features. */
typedef struct ss_shard_view {
    int32_t n_ranks, rank;
    int32_t last_hop_own_only;                        /* 1: hop K is valid for the own row block only */
    int32_t reserved;
    int64_t bounds[SS_MAX_PEERS + 2];                 /* rank q owns rows [bounds[q], bounds[q+1]) */
    const void *peer_records[4][SS_MAX_PEERS + 1];    /* [hop][rank] -> that rank's copy of the hop table (own entry ignored) */
    const uint8_t *local_rows;                        /* device uint8 [num_rows]: 1 = hops 1..K-1 of the row are valid here */
} ss_shard_view;
int ss_link_features_sharded(const int64_t *links, int64_t n_links, const ss_hop_view *hops, int max_hops,
                             int num_perm, int hll_p, const float *cards, int64_t cards_stride,
                             const ss_hll_consts *hc, int flags, float *features_out, float *inter_out,
                             int32_t *error_flag, const ss_shard_view *shard, ss_stream_t stream);

/* ---- next row: common-neighbour heuristics on a SORTED CSR adjacency ------------
 * Replaces CN / AA / RA of /root/reference/src/heuristics.py:11-71 (scipy A[src].multiply(A_[dst]) row sums;
 * HashDataset computes RA with it when --use_RA, datasets/elph.py:76-77):
 *     out[i] = float32( sum over w in N(u) & N(v) of  A[u,w] * (A[v,w] * mult[w]) ),  (u, v) = links[i]
 * rowptr int64 [n_nodes+1], colidx int32 sorted and unique inside each row, weights float64 [nnz] or NULL (= 1),
 * mult float64 [n_nodes] (1 for CN, 1/log(colsum) for AA, 1/colsum for RA; infinities replaced by 0 as the
 * reference does).  ss_col_sums gives colsum (= A.sum(axis=0), also the `degrees` of datasets/elph.py:74).
 */
int ss_col_sums(const int32_t *colidx, const double *weights, int64_t nnz, int64_t n_cols, double *out, ss_stream_t stream);
int ss_common_neighbour_scores(const int64_t *rowptr, const int32_t *colidx, const double *weights, const double *mult,
                               int64_t n_nodes, const int64_t *links, int64_t n_links, float *out, int32_t *error_flag,
                               ss_stream_t stream);

/* ---- next row: SIGN node-feature pre-propagation --------------------------------
 * Replaces HashDataset._generate_sign_features (/root/reference/src/datasets/elph.py:87-110):
 *     edge_index, w = gcn_norm(edge_index, edge_weight.float(), num_nodes)  ;  x' = torch_sparse.spmm(edge_index, w, N, N, x)
 * (PyG gcn_norm: add_remaining_self_loops with fill 1 -- an existing self loop keeps its weight, the last one
 * if there are several --, deg = sum of weights at edge_index[1], w = deg^-1/2[row] * w * deg^-1/2[col], inf -> 0;
 * torch_sparse.spmm: out[row] += w * x[col].)  The normalised edge list is never materialised:
 *   ss_gcn_norm   : dinv_out[i] = deg^-1/2 (0 where deg = 0), loop_weight_out[i] = weight of node i's self loop;
 *                   flags_out (device int32 or NULL): bit 0 = row[] is not sorted (non-decreasing), bit 1 = an id is
 *                   outside [0, n_nodes) (such edges are ignored; the reference would fail in torch indexing)
 *   ss_sign_fill  : perm_out[rowptr[r] .. rowptr[r+1]) = positions e of the edges with row[e] = r, where rowptr is
 *                   ss_csr_rowptr(src = col, dst = row, n_self_loops = 0) -- a CSR of EDGE POSITIONS keyed by the
 *                   spmm row, so arbitrary edge weights need no permuted copy.  Not needed when row[] is sorted:
 *                   pass perm = NULL to ss_sign_spmm, the CSR is the edge list itself and every row is summed in
 *                   edge order like the reference's sequential scatter-add (bit-identical results)
 *   ss_sign_spmm  : out[i, k*F + f] = sum_e (dinv[i] * w_e) * dinv[col_e] * x[col_e, f] + (dinv[i] * loop_w[i]) * dinv[i] * x[i, f]
 *                   for k = 0..copies-1 (the reference concatenates sign_k identical blocks: it re-propagates
 *                   data.x every time, elph.py:104-107), float32, every product and sum rounded separately
 * row / col int64 [n_edges] = edge_index[0] / edge_index[1] (ids in [0, n_nodes): see flags_out),
 * edge_weight float32 [n_edges] or NULL (= 1), x float32 rows of x_stride floats, out rows of out_stride floats.
 * workspace: ss_sign_workspace_bytes(n_nodes) bytes, 256-byte aligned, shared by ss_gcn_norm and ss_sign_fill.
 */
int64_t ss_sign_workspace_bytes(int64_t n_nodes);
int ss_gcn_norm(const int64_t *row, const int64_t *col, const float *edge_weight, int64_t n_edges, int64_t n_nodes,
                float *dinv_out, float *loop_weight_out, int32_t *flags_out, void *workspace, int64_t workspace_bytes,
                ss_stream_t stream);
int ss_sign_fill(const int64_t *row, int64_t n_edges, int64_t n_nodes, const int64_t *rowptr, int32_t *perm_out,
                 void *workspace, int64_t workspace_bytes, ss_stream_t stream);
int ss_sign_spmm(const int64_t *rowptr, const int32_t *perm, const int64_t *col, const float *edge_weight,
                 const float *dinv, const float *loop_weight, const float *x, int64_t x_stride, int64_t n_nodes,
                 int64_t n_features, float *out, int64_t out_stride, int copies, ss_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SS_B200_H */
