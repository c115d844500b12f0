// Shared device helpers for the QFA hot-path kernels (sm_100a).
//
// Per-pixel physics follows reference QFA/model.py:125-131 and QFA/utils.py:72,91-92,
// 105-141, evaluated on the FULL pixel grid with masked pixels given weight 0
// (SURVEY.md section 7.1), instead of the reference's boolean gathers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qfa {

constexpr double kLog2Pi = 1.8378770664093453;  // reference model.py:20

template <typename T> struct Mth;
template <> struct Mth<float> {
    static __device__ __forceinline__ float exp(float x) { return expf(x); }
    static __device__ __forceinline__ float log(float x) { return logf(x); }
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rcp(float x) { return 1.0f / x; }
};
template <> struct Mth<double> {
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
    static __device__ __forceinline__ double log(double x) { return ::log(x); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
};

// tau(z) = t0 * ((1+z)/zn)^be + C      reference utils.py:105,119,133,141
struct LawConst { double t0, be, C, zn; };
inline LawConst law_constants(int law) {
    switch (law) {
        case 0: return {0.751, 2.90, -0.132, 4.5};
        case 1: return {0.0018, 3.92, 0.0, 1.0};
        case 2: return {5.54 * 1e-3, 3.182, 0.0, 1.0};
        default: return {0.2231435513142097, 3.2, 0.0, 3.25};
    }
}

// Everything a kernel needs to evaluate one (spectrum, pixel) cell.
template <typename T>
struct Field {
    const float* x;        // delta (train) or flux (predict), (B, P)
    const float* err;      // (B, P)
    const float* zabs;     // (B, Nb)
    const uint8_t* mask;   // (B, P)
    const float* F;        // (P, Nh)
    const float* Psi;      // (P)
    const float* omega;    // (Nb)
    const float* scal;     // tau0, c0, beta (device)
    const float* mu;       // (P) or nullptr
    int Nb, P, Nh;
    T lt0, lbe, lC, llogzn;  // optical-depth law
};

template <typename T>
struct Cell {
    T A, zdep, om, w, logD, r, powb, logopz;
    bool mk;
};

// PREDICT: residual is flux - mu*A (model.py:166); otherwise x already is delta.
template <typename T, bool PREDICT>
__device__ __forceinline__ Cell<T> eval_cell(const Field<T>& f, size_t b, int i, T tau0, T c0, T beta) {
    Cell<T> c;
    c.A = T(1); c.zdep = T(0); c.om = T(0); c.powb = T(0); c.logopz = T(0);
    if (i < f.Nb) {
        T z = (T)__ldg(f.zabs + b * (size_t)f.Nb + i);
        T L = Mth<T>::log(T(1) + z);
        T tau = f.lt0 * Mth<T>::exp(f.lbe * (L - f.llogzn)) + f.lC;   // utils.py:106 etc.
        c.A = Mth<T>::exp(-tau);                                       // model.py:125
        c.powb = Mth<T>::exp(beta * L);                                // (1+z)^beta, utils.py:72
        T root = T(1) - c0 - Mth<T>::exp(-tau0 * c.powb);              // utils.py:91
        c.zdep = root * root;
        c.om = (T)__ldg(f.omega + i);
        c.logopz = L;
    }
    size_t o = b * (size_t)f.P + i;
    T e = (T)__ldg(f.err + o);
    c.mk = __ldg(f.mask + o) != 0;
    T D = c.A * c.A * (T)__ldg(f.Psi + i) + c.om * c.zdep + e * e;     // model.py:128-131
    c.w = c.mk ? Mth<T>::rcp(D) : T(0);
    c.logD = c.mk ? Mth<T>::log(D) : T(0);
    T xv = (T)__ldg(f.x + o);
    if (PREDICT) xv = xv - (T)__ldg(f.mu + i) * c.A;                   // model.py:166
    c.r = c.mk ? xv : T(0);
    return c;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV values per thread; result valid in every thread. `red` holds NV*32 T.
template <typename T, int NV, int NT>
__device__ __forceinline__ void block_sum(T (&v)[NV], T* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[k * 32 + wid] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        T s = T(0);
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s += red[k * 32 + w];   // fixed order: deterministic
        v[k] = s;
    }
}

}  // namespace qfa
