// lib.cu -- library plumbing of libss_b200: error reporting, device info, record geometry.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ss {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

int check_hll_consts(const ss_hll_consts *hc, int p) {
    SS_REQUIRE(hc != nullptr, "hll constants are required");
    SS_REQUIRE(hc->p == p, "hll constants are for p=%d, tables are p=%d", hc->p, p);
    SS_REQUIRE(hc->table_len >= 6, "bias table needs at least 6 entries (got %d)", hc->table_len);
    SS_REQUIRE(hc->lc_table && hc->raw_estimate && hc->bias, "hll constant tables must be device pointers");
    return SS_OK;
}

}  // namespace ss

extern "C" {

int ss_version(void) { return SS_ABI_VERSION; }

const char *ss_last_error(void) { return ss::g_err; }

int ss_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    SS_CUDA(cudaGetDevice(&dev));
    int sms = 0, maj = 0, min = 0;
    SS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SS_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    SS_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    return SS_OK;
}

int64_t ss_record_bytes(int num_perm, int hll_p) {
    ss::RecordShape s;
    if (!ss::make_shape(num_perm, hll_p, &s)) {
        ss::set_error("unsupported sketch shape num_perm=%d hll_p=%d (need 1<=P<=4096, 4<=p<=18)", num_perm, hll_p);
        return SS_ERR_INVALID;
    }
    return s.bytes;
}

}  // extern "C"
