// orbx_capi_common.cuh — what the C-ABI translation units (orbx_capi.cu, orbx_capi_match.cu) share: the thread-local error
// text behind orbx_last_error(), the CUDA check macro and the stream-ordered scratch pool setting.
#pragma once
#include "../../include/orbx.h"
#include "orbx_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

extern thread_local std::string g_orbx_err;
static inline int fail(int code, const std::string& msg) { g_orbx_err = msg; return code; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ORBX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

// stream-ordered allocations are used for per-call scratch: keep freed blocks in the pool instead of returning them to the
// driver at every synchronisation (the default release threshold is 0)
void orbx_keep_mempool(int device);

// upload helper of the one-shot host forms: arrays are packed into one stream-ordered allocation
struct HostPack {
    std::vector<size_t> off; size_t tot = 0; uint8_t* pool = nullptr;
    size_t add(size_t bytes) { off.push_back(tot); tot += (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255; return off.size() - 1; }
    uint8_t* at(size_t i) const { return pool + off[i]; }
};

// where the device results of an extractor's last HOST call live (orbx_capi.cu): frame `frame_index` of it, the stream that
// produced them and the device; used by the frame handle (orbx_capi_frame.cu)
int orbx_internal_results(orbx_extractor* h, int frame_index, const OrbxKp28** d_kps, const uint8_t** d_desc, cudaStream_t* st, int* device);
