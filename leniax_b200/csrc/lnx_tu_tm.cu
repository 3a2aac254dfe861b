// Translation unit of the default fused kernel (lnx_world128_tm, lnx_kernel_tm.cuh): one instantiation per growth function
// (state function v1) plus gaussian_target / v2, each in a NaN-propagating and a min/max variant.
#include "lnx_internal.h"
#include "lnx_kernel_tm.cuh"

namespace lnx {
namespace host {

bool tm_kernel_exists(int gf, int sf) { return sf == SF_V1 || (sf == SF_V2 && gf == GF_GAUSSIAN_TARGET); }

static const void* tm_kernel_for(int gf, int sf, bool np) {
    if (sf == SF_V2)  // the asymptotic update of conf/config_qd_cmame_v2.yaml and species/2d/1c-1k-v2 (gaussian_target growth)
        return np ? reinterpret_cast<const void*>(&lnx_world128_tm<GF_GAUSSIAN_TARGET, SF_V2, true>)
                  : reinterpret_cast<const void*>(&lnx_world128_tm<GF_GAUSSIAN_TARGET, SF_V2, false>);
#define LNX_TM_CASE(G) \
    case G: return np ? reinterpret_cast<const void*>(&lnx_world128_tm<G, SF_V1, true>) : reinterpret_cast<const void*>(&lnx_world128_tm<G, SF_V1, false>);
    switch (gf) {
        LNX_TM_CASE(GF_POLY_QUAD4)
        LNX_TM_CASE(GF_GAUSSIAN)
        LNX_TM_CASE(GF_GAUSSIAN_TARGET)
        LNX_TM_CASE(GF_STEP)
        LNX_TM_CASE(GF_STAIRCASE)
        LNX_TM_CASE(GF_TRIANGLE)
        default: break;
    }
    return np ? reinterpret_cast<const void*>(&lnx_world128_tm<GF_IDENTITY, SF_V1, true>)
              : reinterpret_cast<const void*>(&lnx_world128_tm<GF_IDENTITY, SF_V1, false>);
#undef LNX_TM_CASE
}

int tm_setup_device() {
    float2 tw[128];
    for (int k = 0; k < 128; ++k) tw[k] = make_float2(Tw128::c[k], Tw128::s[k]);
    LNX_CUDA(cudaMemcpyToSymbol(c_tw128, tw, sizeof(tw)));
    for (int gf = 0; gf <= GF_COUNT; ++gf)  // (the extra round sets up the v2 instantiation)
        for (int np = 0; np < 2; ++np) {
            const void* fn = gf < GF_COUNT ? tm_kernel_for(gf, SF_V1, np != 0) : tm_kernel_for(GF_GAUSSIAN_TARGET, SF_V2, np != 0);
            LNX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM));
            LNX_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
    return LNX_OK;
}

int tm_launch(int gf, int sf, bool nan_propagating, int grid, const RunArgs& a, cudaStream_t st) {
    RunArgs args = a;
    void* kargs[] = {&args};
    LNX_CUDA(cudaLaunchKernel(tm_kernel_for(gf, sf, nan_propagating), dim3(grid), dim3(NT), kargs, TM_SMEM, st));
    return LNX_OK;
}

}  // namespace host
}  // namespace lnx
