// tvl1_internal.h -- declarations shared by the translation units of libtvl1_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tvl1_b200.h"

namespace tvl1 {

extern thread_local char g_err[512];
int fail(int code, const char* fmt, ...);

// sampler scratch owned by a handle (tvl1_sampler.cu)
void sampler_release(void* scratch);
int handle_device(const tvl1_handle* h);
void** handle_sampler_slot(tvl1_handle* h);

}  // namespace tvl1
