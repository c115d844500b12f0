// Probe: what does fence.proxy.async (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC) cost with STS / LDG in flight?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e_=(x); if(e_!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e_),__LINE__);exit(1);} }while(0)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int NSTS, int NLDG, int GAP, int KIND>
__global__ void __launch_bounds__(512, 1) probe(const float* __restrict__ src, float* out, long long* clk, size_t stride) {
    extern __shared__ float sm[];
    const int tid = threadIdx.x;
    float v[NLDG > 0 ? NLDG : 1];
    float acc = 0.f;
    long long tsum = 0;
    for (int it = 0; it < 16; ++it) {
        const float* p = src + ((size_t)blockIdx.x * 64 + it * 4) * stride + tid;
#pragma unroll
        for (int j = 0; j < NLDG; ++j) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[j]) : "l"(p + (size_t)j * 16384));
#pragma unroll
        for (int j = 0; j < NSTS; ++j) sm[j * 512 + tid] = (float)(it + j);
        // GAP independent ALU instructions
        float g = (float)tid;
#pragma unroll
        for (int j = 0; j < GAP; ++j) g = fmaf(g, 1.0001f, 0.5f);
        long long t0 = clock64();
        if (KIND == 0) fence_proxy_async();
        else if (KIND == 1) __threadfence_block();
        else if (KIND == 2) asm volatile("fence.acq_rel.cta;" ::: "memory");
        long long t1 = clock64();
        tsum += t1 - t0;
        acc += g;
#pragma unroll
        for (int j = 0; j < NLDG; ++j) acc += v[j];
        __syncthreads();
    }
    out[blockIdx.x * 512 + tid] = acc + sm[tid];
    if ((tid & 31) == 0) clk[blockIdx.x * 16 + (tid >> 5)] = tsum / 16;
}
template <int NSTS, int NLDG, int GAP, int KIND>
void run(const float* src, float* out, long long* clk, const char* name) {
    CK(cudaFuncSetAttribute(probe<NSTS, NLDG, GAP, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 512 * 4));
    probe<NSTS, NLDG, GAP, KIND><<<148, 512, 64 * 512 * 4>>>(src, out, clk, 1 << 20);
    CK(cudaDeviceSynchronize());
    long long h[148 * 16];
    CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
    double s = 0; long long mx = 0;
    for (int i = 0; i < 148 * 16; ++i) { s += h[i]; if (h[i] > mx) mx = h[i]; }
    printf("%-34s NSTS=%2d NLDG=%2d GAP=%4d : mean %7.0f cycles  max %lld\n", name, NSTS, NLDG, GAP, s / (148 * 16), mx);
}
int main() {
    float *src, *out; long long* clk;
    size_t n = (size_t)148 * 64 * (1 << 20) + (1 << 24);
    CK(cudaMalloc(&src, n * 4)); CK(cudaMemset(src, 0, n * 4));
    CK(cudaMalloc(&out, 148 * 512 * 4)); CK(cudaMalloc(&clk, 148 * 16 * 8));
    run<0, 0, 0, 0>(src, out, clk, "fence.proxy.async idle");
    run<16, 0, 0, 0>(src, out, clk, "fence.proxy.async");
    run<32, 0, 0, 0>(src, out, clk, "fence.proxy.async");
    run<16, 0, 600, 0>(src, out, clk, "fence.proxy.async");
    run<0, 32, 0, 0>(src, out, clk, "fence.proxy.async");
    run<16, 32, 0, 0>(src, out, clk, "fence.proxy.async");
    run<0, 32, 600, 0>(src, out, clk, "fence.proxy.async");
    run<16, 0, 0, 1>(src, out, clk, "__threadfence_block");
    run<0, 32, 0, 1>(src, out, clk, "__threadfence_block");
    run<16, 0, 0, 2>(src, out, clk, "fence.acq_rel.cta");
    run<0, 32, 0, 2>(src, out, clk, "fence.acq_rel.cta");
    return 0;
}
