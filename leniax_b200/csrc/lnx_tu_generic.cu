// Translation unit of the resident kernels for several channels / kernels (lnx_world128_gen_tm, lnx_world128_generic,
// lnx_kernel_generic.cuh) and of the small stand-alone kernels of the resident path (lnx_aux_kernels.cuh).
#include "lnx_internal.h"
#include "lnx_kernel_generic.cuh"
#include "lnx_aux_kernels.cuh"

namespace lnx {
namespace host {

int generic_setup_device() {
    float2 tw[128];
    for (int k = 0; k < 128; ++k) tw[k] = make_float2(Tw128::c[k], Tw128::s[k]);
    LNX_CUDA(cudaMemcpyToSymbol(c_tw128, tw, sizeof(tw)));
    LNX_CUDA(cudaFuncSetAttribute(lnx_world128_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, GENERIC_SMEM));
    LNX_CUDA(cudaFuncSetAttribute(lnx_world128_gen_tm, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
    LNX_CUDA(cudaFuncSetAttribute(lnx_rfft2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + TW_BYTES));
    return LNX_OK;
}

bool gen_tm_supports(int C) { return C <= G2_MAX_C; }

int generic_launch(bool gen_tm, int grid, const RunArgs& a, cudaStream_t st) {
    if (gen_tm)
        lnx_world128_gen_tm<<<grid, NT, G2_SMEM, st>>>(a);
    else
        lnx_world128_generic<<<grid, NTHREADS, GENERIC_SMEM, st>>>(a);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int prepare_launch(const lnx_desc& d, int n_sols, const void* K_fft, void* table, cudaStream_t st) {
    PrepArgs a;
    a.K_fft = static_cast<const float2*>(K_fft);
    a.table = static_cast<float4*>(table);
    a.K = d.nb_kernels;
    a.nb_slots = d.nb_slots;
    for (int k = 0; k < a.K; ++k) a.slot[k] = d.slot[k];
    lnx_prepare_kernel<<<n_sols * a.K, NT, 0, st>>>(a);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int rfft2_launch(int n_images, const float* images, void* spectra, cudaStream_t st) {
    lnx_rfft2_kernel<<<n_images, NT, 65536 + TW_BYTES, st>>>(images, static_cast<float2*>(spectra));
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

int fp32_peak_launch(int grid, int block, float* out, int iters, cudaStream_t st) {
    lnx_fp32_peak_kernel<<<grid, block, 0, st>>>(out, iters, 0.999f, 0.001f);
    LNX_CUDA(cudaGetLastError());
    return LNX_OK;
}

}  // namespace host
}  // namespace lnx
