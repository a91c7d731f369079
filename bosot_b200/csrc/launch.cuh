// Error reporting and launch bookkeeping shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/bosot_b200.h"

namespace bosot {

int set_error(int code, const char* msg);             // stores a thread-local message, returns code
int after_launch(const char* kernel_name);            // counts the launch, maps cudaGetLastError()
int check_grammar(const bosot_grammar* g);            // BOSOT_OK or BOSOT_EINVAL
int cuda_check(cudaError_t e, const char* what);      // BOSOT_OK or BOSOT_ECUDA

}  // namespace bosot
