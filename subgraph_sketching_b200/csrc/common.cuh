// common.cuh -- shared device/host helpers for libss_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ss_b200.h"

namespace ss {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char *fmt, ...);

#define SS_REQUIRE(cond, ...)              \
    do {                                   \
        if (!(cond)) {                     \
            ss::set_error(__VA_ARGS__);    \
            return SS_ERR_INVALID;         \
        }                                  \
    } while (0)

#define SS_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (call);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ss::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__,  \
                          __LINE__);                                                         \
            return SS_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define SS_LAUNCH_CHECK(name)                                                                \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            ss::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));          \
            return SS_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

int sm_count();  // cached SM count of the current device

// record geometry ----------------------------------------------------------------------------
struct RecordShape {
    int P;        // permutations
    int P_pad;    // rounded up to 4 (16-byte boundary)
    int p;        // hll precision
    int m;        // registers
    int mh_bytes; // 4 * P_pad
    int bytes;    // mh_bytes + m
    int units;    // bytes / 8      (a "unit" is 8 bytes = 2 MinHash slots or 8 HLL registers)
    int mh_units; // mh_bytes / 8
};

inline bool make_shape(int P, int p, RecordShape *s) {
    if (P < 1 || P > 4096 || p < 4 || p > 18) return false;
    s->P = P;
    s->P_pad = (P + 3) & ~3;
    s->p = p;
    s->m = 1 << p;
    s->mh_bytes = 4 * s->P_pad;
    s->bytes = s->mh_bytes + s->m;
    s->units = s->bytes / 8;
    s->mh_units = s->mh_bytes / 8;
    return true;
}

// device view of ss_hll_consts
struct HllDev {
    int m;
    int T;
    int monotone;
    float threshold;
    float alpha_m2;
    float five_m;
    const float *lc;
    const float *est;
    const float *bias;
};

inline HllDev to_dev(const ss_hll_consts *hc) {
    HllDev d;
    d.m = 1 << hc->p;
    d.T = hc->table_len;
    d.monotone = hc->monotone;
    d.threshold = hc->threshold;
    d.alpha_m2 = hc->alpha_m2;
    d.five_m = hc->five_m;
    d.lc = hc->lc_table;
    d.est = hc->raw_estimate;
    d.bias = hc->bias;
    return d;
}

int check_hll_consts(const ss_hll_consts *hc, int p);

#ifdef __CUDACC__
// ---------------------------------------------------------------- small device helpers
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint2 ld_nc_u2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_nc_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_na_u2(void *p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_na_u4(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared::cta bulk copy, completion signalled on `bar` as transaction bytes.
// size and both addresses must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- the two merge operators on one 8-byte unit ---------------------------------------------------
__device__ __forceinline__ uint2 unit_min_u32(uint2 a, uint2 b) { return make_uint2(min(a.x, b.x), min(a.y, b.y)); }
__device__ __forceinline__ uint2 unit_max_u8(uint2 a, uint2 b) {
    return make_uint2(__vmaxu4(a.x, b.x), __vmaxu4(a.y, b.y));
}

// ---- exact sum of 2^-r over registers -------------------------------------------------------------
// Registers r in [1, 61] contribute 2^(61-r) to a fixed-point accumulator in units of 2^-61; r == 0 is
// counted separately (it contributes 1 = 2^61 units).  Registers > 61 cannot occur for 4 <= p <= 18.
__device__ __forceinline__ void acc_regs_word(uint32_t w, uint64_t &acc, int &nz) {
    nz += __popc(__vcmpeq4(w, 0u)) >> 3;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t r = (w >> (8 * b)) & 0xffu;
        uint64_t t = 1ull << ((61u - r) & 63u);
        acc += (r != 0u && r <= 61u) ? t : 0ull;
    }
}

// warp total of per-lane (acc, nz); acc per lane must be < 2^63.  Result identical on every lane:
// `zeros` and the 128-bit sum S61 = sum_j 2^(61 - r_j) over ALL registers (zeros included).
__device__ __forceinline__ unsigned __int128 warp_total_units(uint64_t acc, int nz, int &zeros) {
    uint32_t p0 = (uint32_t)(acc & 0x3fffffu);
    uint32_t p1 = (uint32_t)((acc >> 22) & 0x3fffffu);
    uint32_t p2 = (uint32_t)(acc >> 44);
    p0 = __reduce_add_sync(FULL, p0);
    p1 = __reduce_add_sync(FULL, p1);
    p2 = __reduce_add_sync(FULL, p2);
    zeros = (int)__reduce_add_sync(FULL, (uint32_t)nz);
    unsigned __int128 t = (unsigned __int128)p0 + ((unsigned __int128)p1 << 22) + ((unsigned __int128)p2 << 44) +
                          ((unsigned __int128)(uint32_t)zeros << 61);
    return t;
}

// Generic-width variant for any m (K3, generic shapes): two 64-bit lanes so that up to 2^18 registers per
// lane cannot overflow.  lo: registers 31..61 in units of 2^-61; hi: registers 1..30 in units of 2^-30.
struct RegSum {
    uint64_t lo, hi;
    uint32_t zeros;
};
__device__ __forceinline__ void regsum_add_word(RegSum &s, uint32_t w) {
    s.zeros += __popc(__vcmpeq4(w, 0u)) >> 3;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t r = (w >> (8 * b)) & 0xffu;
        if (r >= 31u) s.lo += (r <= 61u) ? (1ull << (61u - r)) : 0ull;
        else if (r != 0u) s.hi += 1ull << (30u - r);
    }
}
__device__ __forceinline__ unsigned __int128 regsum_warp_total(RegSum s, int &zeros) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s.lo += __shfl_xor_sync(FULL, s.lo, o);
        s.hi += __shfl_xor_sync(FULL, s.hi, o);
        s.zeros += __shfl_xor_sync(FULL, s.zeros, o);
    }
    zeros = (int)s.zeros;
    return (unsigned __int128)s.lo + ((unsigned __int128)s.hi << 31) + ((unsigned __int128)s.zeros << 61);
}

// round-to-nearest-even conversion of t * 2^-61 to float32 (t > 0)
__device__ __forceinline__ float units_to_f32(unsigned __int128 t) {
    uint64_t hi = (uint64_t)(t >> 64), lo = (uint64_t)t;
    int msb = hi ? (127 - __clzll((long long)hi)) : (63 - __clzll((long long)lo));
    if (msb <= 23) return scalbnf((float)(uint32_t)lo, -61);
    int shift = msb - 23;
    unsigned __int128 mant = t >> shift;
    unsigned __int128 rem = t & ((((unsigned __int128)1) << shift) - 1);
    unsigned __int128 half = ((unsigned __int128)1) << (shift - 1);
    uint32_t mnt = (uint32_t)mant;
    if (rem > half || (rem == half && (mnt & 1u))) mnt += 1u;
    return scalbnf((float)mnt, shift - 61);
}

// ---- HyperLogLog++ estimate from (zero count, exact register sum) -----------------------------------
// Mirrors hashing.py:212-232 in float32 with the reference's operation order:
//   zeros > 0 and LC <= threshold            -> LC (table built on the host with torch, bit-exact)
//   otherwise e = f32(alpha m^2) * rcp(S);   if e <= 5m: e -= mean(bias of the 6 nearest raw estimates)
// 6-NN: squared float32 distances as in hashing.py:203; ties are unspecified in the reference (unstable
// argsort).  mean = (sum of the six) / 6 in float32.
__device__ __forceinline__ float bias_6nn(const HllDev &h, float e) {
    const float *est = h.est;
    int idx[6];
    if (h.monotone) {
        // lower bound: first j with est[j] >= e
        int lo = 0, hi = h.T;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (__ldg(est + mid) < e) lo = mid + 1; else hi = mid;
        }
        int l = lo - 1, r = lo;
        // two-pointer selection outwards from e.  est is monotone and float32 subtraction is monotone, so the
        // distances on each side are non-decreasing: the picks come out in ascending (distance, index) order --
        // exact ties between the sides go to the lower index, as a stable ascending sort would do.  The only
        // way to leave that order is a run of equal distances on the LEFT side (duplicate table entries), which
        // is detected and repaired by the insertion sort below (never taken with the packaged tables).
        float dd[6];
        bool left_run = false, prev_left = false;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            float dl = 3.4e38f, dr = 3.4e38f;
            if (l >= 0) { float d = __fsub_rn(e, __ldg(est + l)); dl = __fmul_rn(d, d); }
            if (r < h.T) { float d = __fsub_rn(e, __ldg(est + r)); dr = __fmul_rn(d, d); }
            if (l >= 0 && (r >= h.T || dl <= dr)) {
                if (k > 0 && prev_left && dl == dd[k - 1]) left_run = true;
                idx[k] = l; dd[k] = dl; --l; prev_left = true;
            } else {
                idx[k] = r; dd[k] = dr; ++r; prev_left = false;
            }
        }
        if (left_run) {
#pragma unroll
            for (int a = 1; a < 6; ++a) {
#pragma unroll
                for (int b = a; b > 0; --b) {
                    bool sw = dd[b] < dd[b - 1] || (dd[b] == dd[b - 1] && idx[b] < idx[b - 1]);
                    if (sw) { float td = dd[b]; dd[b] = dd[b - 1]; dd[b - 1] = td; int ti = idx[b]; idx[b] = idx[b - 1]; idx[b - 1] = ti; }
                }
            }
        }
    } else {
        float dd[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) { dd[k] = 3.4e38f; idx[k] = 0x7fffffff; }
        for (int j = 0; j < h.T; ++j) {
            float d = __fsub_rn(e, __ldg(est + j));
            float d2 = __fmul_rn(d, d);
            if (d2 < dd[5]) {
                dd[5] = d2; idx[5] = j;
#pragma unroll
                for (int b = 5; b > 0; --b) {
                    if (dd[b] < dd[b - 1]) { float td = dd[b]; dd[b] = dd[b - 1]; dd[b - 1] = td; int ti = idx[b]; idx[b] = idx[b - 1]; idx[b - 1] = ti; }
                }
            }
        }
    }
    // float32 sum in the order ATen's vectorised inner reduction uses for a length-6 row on AVX-512/AVX2
    // hosts (probed: ((((b0+b4)+b5)+b1)+b2)+b3); any order is within the stated tolerance.
    float b0 = __ldg(h.bias + idx[0]), b1 = __ldg(h.bias + idx[1]), b2 = __ldg(h.bias + idx[2]);
    float b3 = __ldg(h.bias + idx[3]), b4 = __ldg(h.bias + idx[4]), b5 = __ldg(h.bias + idx[5]);
    float s = __fadd_rn(b0, b4);
    s = __fadd_rn(s, b5);
    s = __fadd_rn(s, b1);
    s = __fadd_rn(s, b2);
    s = __fadd_rn(s, b3);
    return __fdiv_rn(s, 6.0f);
}

__device__ __forceinline__ float hll_estimate(const HllDev &h, int zeros, unsigned __int128 units) {
    float val = __fadd_rn(h.threshold, 1.0f);
    if (zeros > 0) val = __ldg(h.lc + zeros);
    if (val > h.threshold) {
        float S = units_to_f32(units);
        float e = __fmul_rn(__frcp_rn(S), h.alpha_m2);
        if (e <= h.five_m) e = __fsub_rn(e, bias_6nn(h, e));
        val = e;
    }
    return val;
}
#endif  // __CUDACC__

}  // namespace ss
