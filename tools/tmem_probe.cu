// TMEM as a thread-private backing store: correctness of the 32x32b lane/column addressing for 8 warps x 2 CTAs per SM and
// the tcgen05.ld / tcgen05.st throughput such a use sees (no MMA involved).  Build + run (B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/tmem_probe tools/tmem_probe.cu && gpurun_out/tmem_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t a, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t a, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(a), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t a, float* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
                   "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]) : "r"(a) : "memory");
}

constexpr int NCOLS = 256;

// mode 0: correctness.  mode 1: ld x8 loop (wait per chunk).  mode 2: ld x16 loop.  mode 3: st x8 loop.  mode 4: ld x8, one wait per 64 cols.
__global__ void __launch_bounds__(256, 2) probe(int mode, int iters, int* errors, float* sink) {
    extern __shared__ unsigned char smem[];
    __shared__ uint32_t base_s;
    const int tid = threadIdx.x, w = tid >> 5;
    if (w == 0) tmem_alloc(&base_s, NCOLS);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = base_s;
    // warp w owns lanes 32*(w%4).. and columns (w/4)*64 .. +63 (and the same again at +128)
    const uint32_t mine = base + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)((w >> 2) * 64);
    float v[16];
    if (mode == 0) {
        for (int half = 0; half < 2; ++half)
            for (int c = 0; c < 8; ++c) {
                for (int e = 0; e < 8; ++e) v[e] = (float)(blockIdx.x * 1000003 + tid * 131 + half * 64 + c * 8 + e);
                tmem_st8(mine + half * 128 + c * 8, v);
            }
        tmem_wait_st();
        int bad = 0;
        for (int half = 0; half < 2; ++half)
            for (int c = 0; c < 8; ++c) {
                tmem_ld8(mine + half * 128 + c * 8, v);
                tmem_wait_ld();
                for (int e = 0; e < 8; ++e) bad += v[e] != (float)(blockIdx.x * 1000003 + tid * 131 + half * 64 + c * 8 + e);
            }
        // x16 read of the same data
        for (int c = 0; c < 4; ++c) {
            tmem_ld16(mine + c * 16, v);
            tmem_wait_ld();
            for (int e = 0; e < 16; ++e) bad += v[e] != (float)(blockIdx.x * 1000003 + tid * 131 + c * 16 + e);
        }
        if (bad) atomicAdd(errors, bad);
    } else {
        for (int e = 0; e < 16; ++e) v[e] = (float)(tid + e);
        for (int c = 0; c < 8; ++c) tmem_st8(mine + c * 8, v);
        tmem_wait_st();
        float acc = 0.f;
        for (int it = 0; it < iters; ++it) {
            if (mode == 1) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    tmem_ld8(mine + c * 8, v);
                    tmem_wait_ld();
                    acc += v[0] + v[7];
                }
            } else if (mode == 2) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    tmem_ld16(mine + c * 16, v);
                    tmem_wait_ld();
                    acc += v[0] + v[15];
                }
            } else if (mode == 3) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    v[0] = acc + (float)it;
                    tmem_st8(mine + c * 8, v);
                }
                tmem_wait_st();
                acc += 1.f;
            } else {
                float u[8][8];
#pragma unroll
                for (int c = 0; c < 8; ++c) tmem_ld8(mine + c * 8, u[c]);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 8; ++c) acc += u[c][0] + u[c][7];
            }
        }
        if (acc == 123.456f) sink[0] = acc;
    }
    __syncthreads();
    if (w == 0) tmem_dealloc(base, NCOLS);
}

int main() {
    int* errors;
    float* sink;
    CK(cudaMalloc(&errors, 4));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(errors, 0, 4));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int smem = 80 * 1024;  // like the Lenia kernel: two CTAs per SM
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe, 256, smem));
    const int grid = 2 * prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz, occupancy %d CTAs/SM, grid %d\n", prop.name, prop.multiProcessorCount, prop.clockRate, occ, grid);
    probe<<<grid, 256, smem>>>(0, 0, errors, sink);
    CK(cudaDeviceSynchronize());
    int h = -1;
    CK(cudaMemcpy(&h, errors, 4, cudaMemcpyDeviceToHost));
    printf("correctness: %d mismatches (8 warps x 2 CTAs/SM, 128 columns per thread pair)\n", h);
    const char* names[5] = {"", "ld x8, wait per chunk", "ld x16, wait per chunk", "st x8, wait per 64 cols", "ld x8 x8, one wait"};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int mode = 1; mode <= 4; ++mode) {
        const int iters = 20000;
        probe<<<grid, 256, smem>>>(mode, 100, errors, sink);
        CK(cudaEventRecord(e0));
        probe<<<grid, 256, smem>>>(mode, iters, errors, sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes_per_sm = 2.0 * 256 * 64 * 4 * iters;  // two CTAs per SM
        const double cycles = ms * 1e-3 * 1.965e9;              // assumes the boost clock the Lenia bench ran at
        printf("mode %d (%s): %.3f ms, %.1f B/cycle/SM at 1.965 GHz, %.2f TB/s chip\n", mode, names[mode], ms, bytes_per_sm / cycles,
               bytes_per_sm * prop.multiProcessorCount / (ms * 1e-3) / 1e12);
    }
    return h != 0;
}
