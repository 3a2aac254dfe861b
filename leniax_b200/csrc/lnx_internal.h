// Host-side glue shared by the translation units of libleniax_b200.so (not part of the ABI).
//   lnx_kernels.cu     C ABI (include/leniax_b200.h): plans, argument checks, dispatch; multi-pass engines for worlds that are
//                      not 128x128 (generic tiled passes, 64^3 lines, 2048^2 four-step); direct convolution; statistics summary
//   lnx_tu_tm.cu       lnx_world128_tm      fused 1-channel 1-kernel scan, state + multipliers in tensor memory (default)
//   lnx_tu_generic.cu  lnx_world128_gen_tm / lnx_world128_generic (several channels / kernels), kernel-table builder, rfft2,
//                      FP32 probe
//   lnx_tu_gen2.cu     lnx_world128_gen2    several channels / kernels, two worlds per SM (default when the kernel graph allows)
//   lnx_tu_setup.cu    kernel rasterisation, exact kernel spectrum, random numbers, initial states
// Every kernel family compiles in its own translation unit so that the library builds in parallel.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdint>

#include <cuda_runtime.h>

#include "../../include/leniax_b200.h"

namespace lnx {
struct RunArgs;
}

int lnx_fail(int code, const char* fmt, ...);  // sets the thread-local message, returns code
#define LNX_CUDA(call)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) return lnx_fail(LNX_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

namespace lnx {
namespace host {

// ---- lnx_tu_tm.cu ----
int tm_setup_device();  // once per device: twiddles, shared-memory opt-in of every instantiation
bool tm_kernel_exists(int gf, int sf);
int tm_launch(int gf, int sf, bool nan_propagating, int grid, const RunArgs& a, cudaStream_t st);

// ---- lnx_tu_generic.cu ----
int generic_setup_device();
bool gen_tm_supports(int C);  // lnx_world128_gen_tm (field accumulators in tensor memory); otherwise lnx_world128_generic
int generic_launch(bool gen_tm, int grid, const RunArgs& a, cudaStream_t st);
int prepare_launch(const lnx_desc& d, int n_sols, const void* K_fft, void* table, cudaStream_t st);
int rfft2_launch(int n_images, const float* images, void* spectra, cudaStream_t st);
int fp32_peak_launch(int grid, int block, float* out, int iters, cudaStream_t st);

// ---- lnx_tu_gen2.cu ----
int gen2_setup_device();
bool gen2_plan(const lnx_desc& d, RunArgs& a);  // fills the schedule fields of `a`; false: this kernel graph runs in gen_tm instead
size_t gen2_scratch_planes(int C);              // 64 KB planes of L2 scratch per CTA
int gen2_launch(int grid, const RunArgs& a, cudaStream_t st);

}  // namespace host
}  // namespace lnx
