// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "../../include/cer_b200.h"

namespace cer {
int set_error(int code, const std::string& msg);   // records msg for cer_last_error(), returns code
}

#define CER_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::cer::set_error(CER_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
  } while (0)
