// Host-side helpers shared by the C-ABI translation units: error reporting, TMA descriptor encode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/fgt_b200.h"

namespace fgt {

// thread-local last-error text (returned by fgt_last_error()).
char* err_buf();
int set_err(int code, const char* fmt, ...);

#define FGT_REQUIRE(cond, code, ...)                   \
  do {                                                 \
    if (!(cond)) return fgt::set_err((code), __VA_ARGS__); \
  } while (0)

#define FGT_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return fgt::set_err(FGT_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                 \
                          cudaGetErrorString(e__), __FILE__, __LINE__);                 \
  } while (0)

// Encode a bf16 tiled tensor map (rank <= 5) with the 128B swizzle and zero OOB fill.
// dims[0] is the contiguous dimension; strides_bytes has rank-1 entries (dims 1..rank-1).
int encode_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box);

int num_sms();

}  // namespace fgt
