// Translation unit of lnx_world128_gen2 (lnx_kernel_gen2.cuh): several channels / kernels, two worlds per SM.
#include "lnx_internal.h"
#include "lnx_kernel_gen2.cuh"

namespace lnx {
namespace host {

int gen2_setup_device() {
    float2 tw[128];
    for (int k = 0; k < 128; ++k) tw[k] = make_float2(Tw128::c[k], Tw128::s[k]);
    LNX_CUDA(cudaMemcpyToSymbol(c_tw128, tw, sizeof(tw)));
    LNX_CUDA(cudaFuncSetAttribute(lnx_world128_gen2, cudaFuncAttributeMaxDynamicSharedMemorySize, G3_SMEM));
    LNX_CUDA(cudaFuncSetAttribute(lnx_world128_gen2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    return LNX_OK;
}

bool gen2_plan(const lnx_desc& d, RunArgs& a) { return gen2_schedule(d.nb_channels, d.nb_kernels, d.c_in, d.c_out, a); }

size_t gen2_scratch_planes(int C) { return (size_t)(C + 1); }

int gen2_launch(int grid, const RunArgs& a, cudaStream_t st) {
    lnx_world128_gen2<<<grid, NT, G3_SMEM, st>>>(a);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

}  // namespace host
}  // namespace lnx
