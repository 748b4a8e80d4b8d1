// init.cu -- K1: hop-0 sketch records, and conversion between compact records and the reference's
// int64/int8 tensors.
//
// Replaces ElphHashes.initialise_minhash (/root/reference/src/hashing.py:118-124) and initialise_hll
// (:126-137).  One warp writes one 768-byte record (coalesced 8-byte units); nothing is read from HBM but
// the permutation parameters (L1-resident).
#include "common.cuh"

namespace ss {

// pandas.util.hash_array on an int64 id: splitmix64-style finaliser (hashing.py:121,128 call it on
// arange(1, n+1)).
__device__ __forceinline__ uint64_t node_hash64(uint64_t v) {
    v ^= v >> 30;
    v *= 0xBF58476D1CE4E5B9ull;
    v ^= v >> 27;
    v *= 0x94D049BB133111EBull;
    v ^= v >> 31;
    return v;
}

// ((a*h + b) mod 2^64) mod (2^61 - 1), low 32 bits (hashing.py:122; the product wraps before the modulus)
__device__ __forceinline__ uint32_t permuted_slot(uint64_t h, uint64_t a, uint64_t b) {
    const uint64_t M = (1ull << 61) - 1;
    uint64_t x = a * h + b;
    uint64_t r = (x & M) + (x >> 61);
    if (r >= M) r -= M;
    return (uint32_t)r;
}

// The reference computes bit_length(bits) as ceil(log2(bits + 1)) in float64 (hashing.py:83-89).  That is
// the true bit length except for v = bits + 1 in (2^k, 2^k + window[k]] where the float64 result rounds
// down to k.  window[] is measured on the host with numpy, so the quirk is reproduced exactly.
__device__ __forceinline__ int ref_bit_length(uint64_t bits, const int32_t *__restrict__ window) {
    uint64_t v = bits + 1;
    int k = 63 - __clzll((long long)v);  // floor(log2 v), v >= 1
    uint64_t excess = v - (1ull << k);
    if (excess == 0) return k;
    return (excess <= (uint64_t)(uint32_t)__ldg(window + k)) ? k : k + 1;
}

__global__ void __launch_bounds__(256) init_records_kernel(int64_t n, uint64_t first_id, RecordShape s,
                                                            const uint64_t *__restrict__ pa,
                                                            const uint64_t *__restrict__ pb,
                                                            const int32_t *__restrict__ window,
                                                            uint8_t *__restrict__ out, int64_t out_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += n_warps) {
        const uint64_t h = node_hash64(first_id + (uint64_t)i);
        const uint32_t slot = (uint32_t)(h & (uint64_t)(s.m - 1));
        const int rank = (64 - s.p) - ref_bit_length(h >> s.p, window) + 1;
        uint8_t *row = out + i * out_stride;
        for (int u = lane; u < s.units; u += 32) {
            uint2 v;
            if (u < s.mh_units) {
                int j = 2 * u;
                v.x = (j < s.P) ? permuted_slot(h, __ldg(pa + j), __ldg(pb + j)) : 0u;
                v.y = (j + 1 < s.P) ? permuted_slot(h, __ldg(pa + j + 1), __ldg(pb + j + 1)) : 0u;
            } else {
                uint32_t byte0 = (uint32_t)(u - s.mh_units) * 8u;
                uint64_t w = (slot >= byte0 && slot < byte0 + 8u) ? ((uint64_t)(uint32_t)rank << (8u * (slot - byte0))) : 0ull;
                v.x = (uint32_t)w;
                v.y = (uint32_t)(w >> 32);
            }
            st_na_u2(row + (int64_t)u * 8, v);
        }
    }
}

// reference tensors -> records.  One thread per 8-byte unit.
// error_flag (optional, int32[2] initialised to {0, 1} by the caller): [0] is set and [1] cleared when a value cannot be represented in a record (MinHash outside [0, 2^32), register
// outside [0, 127]) -- the host then knows, without having synchronised up front, that the record engine was not
// applicable to this tensor.  guard: see ss_khop_merge_ex.
__global__ void __launch_bounds__(256) pack_kernel(const int64_t *__restrict__ mh, const int8_t *__restrict__ hll,
                                                    int64_t n, RecordShape s, uint8_t *__restrict__ out,
                                                    int64_t out_stride, int *error_flag = nullptr,
                                                    const int *guard = nullptr) {
    if (guard && *reinterpret_cast<const volatile int *>(guard) == 0) return;
    bool bad = false;
    const int64_t total = n * (int64_t)s.units;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / s.units;
        const int u = (int)(t - i * s.units);
        uint2 v;
        if (u < s.mh_units) {
            if (!mh) continue;
            int j = 2 * u;
            const int64_t x = (j < s.P) ? mh[i * s.P + j] : 0, y = (j + 1 < s.P) ? mh[i * s.P + j + 1] : 0;
            bad |= ((uint64_t)x | (uint64_t)y) >> 32 != 0;
            v.x = (uint32_t)x;
            v.y = (uint32_t)y;
        } else {
            if (!hll) continue;
            v = *reinterpret_cast<const uint2 *>(hll + i * (int64_t)s.m + (int64_t)(u - s.mh_units) * 8);
            bad |= ((v.x | v.y) & 0x80808080u) != 0u;
        }
        *reinterpret_cast<uint2 *>(out + i * out_stride + (int64_t)u * 8) = v;
    }
    if (error_flag && bad) {  // int32[2]: [0] = 'does not fit' (set), [1] = 'fits' (cleared) -- either can serve as a guard
        atomicExch(error_flag, 1);
        atomicExch(error_flag + 1, 0);
    }
}

__global__ void __launch_bounds__(256) unpack_kernel(const uint8_t *__restrict__ rec, int64_t rec_stride, int64_t n,
                                                      RecordShape s,
                                                      int64_t *__restrict__ mh, int8_t *__restrict__ hll,
                                                      const int *guard = nullptr) {
    if (guard && *reinterpret_cast<const volatile int *>(guard) == 0) return;
    const int64_t total = n * (int64_t)s.units;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / s.units;
        const int u = (int)(t - i * s.units);
        const uint2 v = *reinterpret_cast<const uint2 *>(rec + i * rec_stride + (int64_t)u * 8);
        if (u < s.mh_units) {
            if (!mh) continue;
            int j = 2 * u;
            if (j < s.P) mh[i * s.P + j] = (int64_t)v.x;
            if (j + 1 < s.P) mh[i * s.P + j + 1] = (int64_t)v.y;
        } else {
            if (!hll) continue;
            *reinterpret_cast<uint2 *>(hll + i * (int64_t)s.m + (int64_t)(u - s.mh_units) * 8) = v;
        }
    }
}

static int grid_for(int64_t threads_needed, int block) {
    int64_t blocks = (threads_needed + block - 1) / block;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace ss

extern "C" {

int ss_init_records(int64_t n, int64_t first_id, int num_perm, int hll_p, const uint64_t *perm_a,
                    const uint64_t *perm_b, const int32_t *log2_window, void *rec_out, int64_t out_stride,
                    ss_stream_t stream) {
    ss::RecordShape s;
    SS_REQUIRE(ss::make_shape(num_perm, hll_p, &s), "unsupported sketch shape num_perm=%d hll_p=%d", num_perm, hll_p);
    SS_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return SS_OK;
    SS_REQUIRE(perm_a && perm_b && log2_window && rec_out, "null pointer passed to ss_init_records");
    SS_REQUIRE(((uintptr_t)rec_out & 15) == 0, "record table must be 16-byte aligned");
    SS_REQUIRE(out_stride >= s.bytes && (out_stride & 15) == 0, "bad record stride %lld", (long long)out_stride);
    int grid = ss::grid_for(n * 32, 256);
    ss::init_records_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, (uint64_t)first_id, s, perm_a, perm_b,
                                                                   log2_window, (uint8_t *)rec_out, out_stride);
    SS_LAUNCH_CHECK("init_records_kernel");
    return SS_OK;
}

int ss_pack_records(const int64_t *minhash, const int8_t *hll, int64_t n, int num_perm, int hll_p, void *rec_out,
                    int64_t out_stride, ss_stream_t stream) {
    return ss_pack_records_ex(minhash, hll, n, num_perm, hll_p, rec_out, out_stride, nullptr, nullptr, stream);
}

int ss_pack_records_ex(const int64_t *minhash, const int8_t *hll, int64_t n, int num_perm, int hll_p, void *rec_out,
                       int64_t out_stride, int32_t *error_flag, const int32_t *guard, ss_stream_t stream) {
    ss::RecordShape s;
    SS_REQUIRE(ss::make_shape(num_perm, hll_p, &s), "unsupported sketch shape num_perm=%d hll_p=%d", num_perm, hll_p);
    SS_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return SS_OK;
    SS_REQUIRE(rec_out && (minhash || hll), "null pointer passed to ss_pack_records");
    SS_REQUIRE(((uintptr_t)rec_out & 15) == 0, "record table must be 16-byte aligned");
    SS_REQUIRE(((uintptr_t)hll & 7) == 0, "hll tensor must be 8-byte aligned");
    // a MinHash-only table (hll == NULL) may be as narrow as the MinHash part of a record (SS_LAYOUT_MINHASH rows)
    SS_REQUIRE(out_stride >= (hll ? s.bytes : s.mh_bytes) && (out_stride & 15) == 0, "bad record stride %lld", (long long)out_stride);
    int grid = ss::grid_for(n * s.units, 256);
    ss::pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(minhash, hll, n, s, (uint8_t *)rec_out, out_stride, error_flag, guard);
    SS_LAUNCH_CHECK("pack_kernel");
    return SS_OK;
}

int ss_unpack_records(const void *rec, int64_t rec_stride, int64_t n, int num_perm, int hll_p, int64_t *minhash_out,
                      int8_t *hll_out, ss_stream_t stream) {
    return ss_unpack_records_ex(rec, rec_stride, n, num_perm, hll_p, minhash_out, hll_out, nullptr, stream);
}

int ss_unpack_records_ex(const void *rec, int64_t rec_stride, int64_t n, int num_perm, int hll_p, int64_t *minhash_out,
                         int8_t *hll_out, const int32_t *guard, ss_stream_t stream) {
    ss::RecordShape s;
    SS_REQUIRE(ss::make_shape(num_perm, hll_p, &s), "unsupported sketch shape num_perm=%d hll_p=%d", num_perm, hll_p);
    SS_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return SS_OK;
    SS_REQUIRE(rec && (minhash_out || hll_out), "null pointer passed to ss_unpack_records");
    SS_REQUIRE(((uintptr_t)hll_out & 7) == 0, "hll tensor must be 8-byte aligned");
    SS_REQUIRE(rec_stride >= (hll_out ? s.bytes : s.mh_bytes) && (rec_stride & 15) == 0, "bad record stride %lld", (long long)rec_stride);
    int grid = ss::grid_for(n * s.units, 256);
    ss::unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint8_t *)rec, rec_stride, n, s, minhash_out, hll_out, guard);
    SS_LAUNCH_CHECK("unpack_kernel");
    return SS_OK;
}

}  // extern "C"
