// tvl1_internal.h -- declarations shared by the translation units of libtvl1_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tvl1_b200.h"

namespace tvl1 {

extern thread_local char g_err[512];
int fail(int code, const char* fmt, ...);

// Small transfers between the device and PINNED host memory without the copy engines: with unified
// addressing a kernel reads / writes pinned host memory directly.  The solver's 32-byte stop-flag
// read-backs (one per outer iteration) and the sampler's few hundred bytes must not queue behind the
// bulk flow downloads on the D2H engine -- measured: a pair solved while the previous pair's 512 MB of
// flow drains lost ~7 ms to exactly that.  n32: number of 32-bit words.
int copy_words(const void* src, void* dst, int n32, cudaStream_t st);

// sampler scratch owned by a handle (tvl1_sampler.cu)
void sampler_release(void* scratch);
int handle_device(const tvl1_handle* h);
void** handle_sampler_slot(tvl1_handle* h);
// feature scratch owned by a handle (tvl1_features.cu)
void features_release(void* scratch);
void** handle_feature_slot(tvl1_handle* h);

}  // namespace tvl1
