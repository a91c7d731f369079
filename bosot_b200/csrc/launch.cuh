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

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// The hot path is a chain of dependent launches on one stream (scan -> pairs -> posterior).  A kernel launched with
// launch_pdl() may START while its predecessor on the stream is still draining: its CTAs are scheduled as soon as
// every CTA of the predecessor has executed pdl_launch_dependents() (or exited), run their prologue (barrier
// set-up, staging of tables that older kernels produced, zero fills) and then block in pdl_wait() until the
// predecessor has completed and its writes are visible.  Nothing the predecessor writes may be read before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();  // c_api.cu: false when the environment variable BOSOT_NO_PDL is set (measurement switch)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool allow_early_start, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (allow_early_start && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace bosot
