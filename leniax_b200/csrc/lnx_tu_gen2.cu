// Translation unit of lnx_world128_gen2 (lnx_kernel_gen2.cuh): several channels / kernels, two worlds per SM.
#include "lnx_internal.h"
#include "lnx_kernel_gen2.cuh"

namespace lnx {
namespace host {

// the specialised instantiation covers what the reference's multi-channel configurations use (conf/config_qd_cmame_3c6k.yaml,
// conf/species/2d/*: poly_quad4 growth, v1 update); everything else selects per kernel at run time
static const void* gen2_kernel(int variant) {  // 0: run-time selection, 1: poly_quad4 + v1, NaN-propagating clamps, 2: poly_quad4 + v1, plain clamps
    if (variant == 1) return reinterpret_cast<const void*>(&lnx_world128_gen2<GF_POLY_QUAD4, SF_V1, true>);
    if (variant == 2) return reinterpret_cast<const void*>(&lnx_world128_gen2<GF_POLY_QUAD4, SF_V1, false>);
    return reinterpret_cast<const void*>(&lnx_world128_gen2<-1, -1, true>);
}

int gen2_setup_device() {
    float2 tw[128];
    for (int k = 0; k < 128; ++k) tw[k] = make_float2(Tw128::c[k], Tw128::s[k]);
    LNX_CUDA(cudaMemcpyToSymbol(c_tw128, tw, sizeof(tw)));
    for (const void* fn : {gen2_kernel(0), gen2_kernel(1), gen2_kernel(2)}) {
        LNX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, G3_SMEM));
        LNX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    return LNX_OK;
}

bool gen2_plan(const lnx_desc& d, RunArgs& a) { return gen2_schedule(d.nb_channels, d.nb_kernels, d.c_in, d.c_out, a); }

size_t gen2_scratch_planes(int C) { return (size_t)(C + 1); }

int gen2_launch(int grid, const RunArgs& a, cudaStream_t st) {
    bool fast = a.state_fn == SF_V1;
    for (int k = 0; k < a.K; ++k) fast = fast && a.gf_id[k] == GF_POLY_QUAD4;
    RunArgs args = a;
    void* kargs[] = {&args};
    LNX_CUDA(cudaLaunchKernel(gen2_kernel(fast ? ((a.flags & LNX_RUN_ASSUME_FINITE) ? 2 : 1) : 0), dim3(grid), dim3(NT), kargs, G3_SMEM, st));
    return LNX_OK;
}

}  // namespace host
}  // namespace lnx
