// Probe: achievable HBM write bandwidth for the phase-O store pattern: each warp-store writes 32 consecutive
// floats (128 B) of one row, 16 rows per step (row pitch P floats), 128 rows x P per CTA tile.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e_=(x); if(e_!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e_),__LINE__);exit(1);} }while(0)
template <int MODE> __device__ __forceinline__ void st(float* p, float v) {
    if (MODE == 0) asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    else if (MODE == 1) asm volatile("st.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    else if (MODE == 2) asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
    else asm volatile("st.global.cg.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
// 16 warps: quad = w&3 (32-pixel column group), grp = w>>2 (16 rows); loop pt over pixel tiles, h over halves
template <int MODE, int ORDER>
__global__ void __launch_bounds__(512, 1) probe(float* __restrict__ c, float* __restrict__ u, int P, int B) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, quad = w & 3, grp = w >> 2;
    const int ntiles = B / 128, npt = (P + 127) / 128;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (ORDER == 0) {
            for (int pt = 0; pt < npt; ++pt)
                for (int h = 0; h < 2; ++h) {
                    const int i = pt * 128 + quad * 32 + lane;
                    const size_t o0 = (size_t)(tile * 128 + h * 64 + grp * 16) * P + i;
                    if (i < P) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) { st<MODE>(c + o0 + (size_t)r * P, 1.0f); st<MODE>(u + o0 + (size_t)r * P, 2.0f); }
                    }
                }
        } else {   // row-major: each warp writes whole rows (16 warps x 8 rows), 128 B per store instr, sequential
            for (int r = 0; r < 8; ++r) {
                const size_t o0 = (size_t)(tile * 128 + w * 8 + r) * P;
                for (int i = lane; i < P; i += 32) { st<MODE>(c + o0 + i, 1.0f); st<MODE>(u + o0 + i, 2.0f); }
            }
        }
    }
}
template <int MODE, int ORDER>
void run(float* c, float* u, int P, int B, int grid, int smem, const char* name) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    if (smem) CK(cudaFuncSetAttribute(probe<MODE, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int i = 0; i < 2; ++i) probe<MODE, ORDER><<<grid, 512, smem>>>(c, u, P, B);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; ++i) probe<MODE, ORDER><<<grid, 512, smem>>>(c, u, P, B);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
    printf("%-28s P=%d grid=%d smem=%3dKB: %.3f ms  %.0f GB/s\n", name, P, grid, smem / 1024, ms, 2.0 * B * P * 4 / ms / 1e6);
}
int main() {
    const int B = 148 * 128 * 4;
    float *c, *u;
    CK(cudaMalloc(&c, (size_t)B * 1920 * 4)); CK(cudaMalloc(&u, (size_t)B * 1920 * 4));
    for (int P : {1913, 1920}) {
        run<0, 0>(c, u, P, B, 148, 120 * 1024, "tile-order no_allocate");
        run<1, 0>(c, u, P, B, 148, 120 * 1024, "tile-order default");
        run<2, 0>(c, u, P, B, 148, 120 * 1024, "tile-order .cs");
        run<3, 0>(c, u, P, B, 148, 120 * 1024, "tile-order .cg");
        run<0, 1>(c, u, P, B, 148, 120 * 1024, "row-order no_allocate");
        run<1, 1>(c, u, P, B, 148, 120 * 1024, "row-order default");
        run<0, 0>(c, u, P, B, 148, 200 * 1024, "tile-order no_alloc bigsmem");
        run<0, 0>(c, u, P, B, 296, 100 * 1024, "tile-order no_alloc 2cta");
    }
    return 0;
}
