// Probe: achievable HBM read bandwidth for the tile access pattern of k_tc_gram:
// persistent CTA owns 128 rows (row pitch P floats, P odd), visits them in runs of RUNK*32 floats.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e_=(x); if(e_!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e_),__LINE__);exit(1);} }while(0)

__device__ __forceinline__ float ldg_stream(const float* p) {
    float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
// NW warps; each warp owns 128/NW rows; per visit: RUNK consecutive 32-float blocks of RB rows at a time
template <int NW, int RB, int RUNK>
__global__ void __launch_bounds__(NW * 32, 1) probe(const float* __restrict__ x, const float* __restrict__ e, int P, int B, float* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int RPW = 128 / NW;
    float acc = 0.f;
    const int ntiles = B / 128;
    const int nkb = P / 32;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * 128 + warp * RPW;
        for (int kb = 0; kb + RUNK <= nkb; kb += RUNK) {
            for (int r0 = 0; r0 < RPW; r0 += RB) {
                float v[RB][RUNK][2];
#pragma unroll
                for (int r = 0; r < RB; ++r)
#pragma unroll
                    for (int c = 0; c < RUNK; ++c) {
                        size_t o = (size_t)(row0 + r0 + r) * P + (kb + c) * 32 + lane;
                        v[r][c][0] = ldg_stream(x + o);
                        v[r][c][1] = ldg_stream(e + o);
                    }
#pragma unroll
                for (int r = 0; r < RB; ++r)
#pragma unroll
                    for (int c = 0; c < RUNK; ++c) acc += v[r][c][0] * v[r][c][1];
            }
        }
    }
    if (acc == 12345.678f) out[0] = acc;
}

template <int NW, int RB, int RUNK>
void run(const float* x, const float* e, int P, int B, float* out, int grid, int occ_smem) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    if (occ_smem) CK(cudaFuncSetAttribute(probe<NW, RB, RUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, occ_smem));
    for (int i = 0; i < 2; ++i) probe<NW, RB, RUNK><<<grid, NW * 32, occ_smem>>>(x, e, P, B, out);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; ++i) probe<NW, RB, RUNK><<<grid, NW * 32, occ_smem>>>(x, e, P, B, out);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
    double bytes = 2.0 * (double)(B / 128 * 128) * (P / 32 / RUNK * RUNK * 32) * 4;
    printf("NW=%2d RB=%d RUNK=%d grid=%d inflight/thread=%2d : %.3f ms  %.0f GB/s\n", NW, RB, RUNK, grid, RB * RUNK * 2, ms, bytes / ms / 1e6);
}

int main() {
    const int P = 1913, B = 148 * 128 * 4;
    float *x, *e, *out;
    CK(cudaMalloc(&x, (size_t)B * P * 4)); CK(cudaMalloc(&e, (size_t)B * P * 4)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(x, 0, (size_t)B * P * 4)); CK(cudaMemset(e, 0, (size_t)B * P * 4));
    printf("-- smem sweep, 1 CTA/SM (grid 148), 16 warps, RB=4 RUNK=4\n");
    for (int kb : {0, 32, 64, 100, 132, 164, 200}) { printf("smem=%3d KB: ", kb); run<16, 4, 4>(x, e, P, B, out, 148, kb * 1024); }
    printf("-- smem sweep, grid 296, 16 warps\n");
    for (int kb : {0, 32, 64, 100}) { printf("smem=%3d KB: ", kb); run<16, 4, 4>(x, e, P, B, out, 296, kb * 1024); }
    printf("-- grid 444 / 592, smem 0 / 64\n");
    run<16, 4, 4>(x, e, P, B, out, 444, 0);
    run<16, 4, 4>(x, e, P, B, out, 592, 0);
    run<16, 4, 4>(x, e, P, B, out, 444, 64 * 1024);
    run<8, 4, 4>(x, e, P, B, out, 592, 50 * 1024);
    run<8, 4, 4>(x, e, P, B, out, 296, 100 * 1024);
    run<32, 4, 1>(x, e, P, B, out, 148, 0);
    run<32, 4, 4>(x, e, P, B, out, 148, 0);
    return 0;
}
