// Kernels either side of the hot path, sm_100a:
//   k_prepare          zabs from zqso, delta = flux - mu * exp(-tau_total)              dataloader.py:102,135-136, utils.py:174-203
//   k_gather_prepare   the same for a SHUFFLED batch: rows are gathered through a device-resident permutation, the batch
//                      start is read from a device cursor (so a CUDA graph can replay the step without host arguments)
//   k_tau_weight_sums  the two column sums of the mean spectrum                          dataloader.py:110-111
//   k_ood_select       out-of-distribution scoring on the per-spectrum NLL: threshold count / compaction and exact top-k
//   k_sample_posterior h ~ N(hmean, hcov) (+ continuum samples mu + F h)                 nb/predict.ipynb cell 11
// All of them are HBM-bound byte work: coalesced row-major accesses, one pass over the data.
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#include "qfa_common.cuh"

namespace qfa {
namespace aux {

// Lyman series (oscillator strength f, wavelength in Angstrom): the atomic data the reference reads from
// QFA/Lyman_series.csv (utils.py:144-147); coefficient of line s = lambda_s f_s / (lambda_1 f_1).
constexpr int kNSeries = 30;
__constant__ float c_ly_lambda[kNSeries] = {
    1215.6701f, 1025.7222f, 972.5367f, 949.7430f, 937.8034f, 930.7482f, 926.2256f, 923.1503f, 920.9630f, 919.3513f,
    918.1293f,  917.1805f,  916.4291f, 915.8238f, 915.3289f, 914.9192f, 914.5762f, 914.2861f, 914.0385f, 913.8256f,
    913.6411f,  913.4803f,  913.3391f, 913.2146f, 913.1042f, 913.0059f, 912.9179f, 912.8389f, 912.7676f, 912.7032f};
__constant__ float c_ly_f[kNSeries] = {
    4.1620e-01f, 7.9140e-02f, 2.9010e-02f, 1.3950e-02f, 7.8030e-03f, 4.8160e-03f, 3.1850e-03f, 2.2170e-03f, 1.6060e-03f, 1.2010e-03f,
    9.2190e-04f, 7.2310e-04f, 5.7770e-04f, 4.6890e-04f, 3.8580e-04f, 3.2120e-04f, 2.7030e-04f, 2.2970e-04f, 1.9680e-04f, 1.6990e-04f,
    1.4770e-04f, 1.2930e-04f, 1.1370e-04f, 1.0060e-04f, 8.9360e-05f, 7.9780e-05f, 7.1480e-05f, 6.4350e-05f, 5.8120e-05f, 5.2640e-05f};

struct Law { float t0, be, C, zn; };

// total mean optical depth of rest-frame pixel `wav` of a quasar at zq: sum over the Lyman lines redward of the pixel
// (utils.py:186-201; max_series = 1 restricts it to Ly-alpha, which is what the model itself uses, model.py:125)
__device__ __forceinline__ float tau_total_px(float wav, float opz, const Law& lw, int max_series) {
    float tau = 0.0f;
    const float lf0 = c_ly_lambda[0] * c_ly_f[0];
#pragma unroll 1
    for (int s = 0; s < max_series; ++s) {
        const float lam = c_ly_lambda[s];
        if (!(wav < lam)) break;                                 // lines are sorted by decreasing wavelength
        const float z1 = opz * wav / lam;                        // 1 + zabs of this line        utils.py:199
        const float t = lw.t0 * powf(z1 / lw.zn, lw.be) + lw.C;  // utils.py:105,119,133,141
        tau += t * (lam * c_ly_f[s] / lf0);                      // utils.py:147,160
    }
    return tau;
}

// idx == nullptr: rows b0.. of the arrays as they lie; otherwise row perm[cursor + b].  cursor (device, may be null = 0).
struct PrepArgs {
    const float* flux; const float* error; const uint8_t* mask; const float* zq;   // the resident data set (N rows)
    const float* wav; const float* mu;
    const int64_t* perm; const int64_t* cursor;
    int B, Nb, P, max_series;
    Law lw;
    float* zabs_out; float* delta_out; float* error_out; uint8_t* mask_out;        // (B, ...) batch buffers; any may be null
};

// One CTA streams whole rows (grid-stride over the batch rows): the row index / redshift of the NEXT row is fetched while the
// current one streams, and every thread keeps four independent pixels in flight.  (The first version used one 256-pixel CTA
// per row segment: 65 536 tiny CTAs for a batch of 8 192, each starting with a dependent perm -> zq -> data load chain:
// 260 us for 400 MB, latency-bound.)
//
// TABLE = true (whenever 3 Nb floats fit in shared memory): the power-law optical depth SEPARATES into a per-row and a per-pixel
// factor,  tau_total(i, row) = (1 + zq_row)^be * T1[i] + T0[i]  with  T1[i] = t0 * sum_s c_s (wav_i / (lambda_s zn))^be  and
// T0[i] = C * sum_s c_s  over the lines s redward of pixel i (which lines those are depends on the pixel only).  Every CTA builds
// T1, T0 and wav_i / 1215.67 once in shared memory; per cell that leaves one FMA and one exp where the direct form evaluates
// a powf and two IEEE divisions per line (ncu, round 2: 150 instructions per pixel on average, issue slots 74 % busy, 2.7 TB/s;
// the row-gather itself is 305 MB for a batch of 8 192).  (1 + zq)^be is computed by ONE thread per row, one row ahead, and
// broadcast through shared memory.  The two forms agree to a few ulp of tau (two accurate powf instead of one).
template <bool TABLE>
__global__ void __launch_bounds__(256) k_gather_prepare(const PrepArgs a) {
    extern __shared__ float s_tab[];                 // TABLE: T1[Nb] | T0[Nb] | T2[Nb]
    __shared__ float s_g[2];
    const int P = a.P, Nb = a.Nb;
    const float* __restrict__ flux = a.flux; const float* __restrict__ error = a.error; const uint8_t* __restrict__ mask = a.mask;
    const float* __restrict__ mu = a.mu;
    float* __restrict__ delta_out = a.delta_out; float* __restrict__ error_out = a.error_out;
    uint8_t* __restrict__ mask_out = a.mask_out; float* __restrict__ zabs_out = a.zabs_out;
    const float* T1 = s_tab; const float* T0 = s_tab + Nb; const float* T2 = s_tab + 2 * Nb;
    const int64_t cur = a.cursor ? *a.cursor : 0;
    int b = blockIdx.x;
    if (b >= a.B) return;
    int64_t src = a.perm ? a.perm[cur + b] : cur + b;
    float opz = 1.0f + a.zq[src];
    if (TABLE) {
        const float lf0 = c_ly_lambda[0] * c_ly_f[0];
        for (int i = threadIdx.x; i < Nb; i += blockDim.x) {
            const float wv = a.wav[i];
            float t1 = 0.0f, t0 = 0.0f;
#pragma unroll 1
            for (int s = 0; s < a.max_series; ++s) {
                const float lam = c_ly_lambda[s];
                if (!(wv < lam)) break;
                const float c = lam * c_ly_f[s] / lf0;
                t1 += c * powf(wv / (lam * a.lw.zn), a.lw.be);
                t0 += c;
            }
            s_tab[i] = a.lw.t0 * t1; s_tab[Nb + i] = a.lw.C * t0; s_tab[2 * Nb + i] = wv / 1215.67f;
        }
        if (threadIdx.x == 0) s_g[0] = powf(opz, a.lw.be);
        __syncthreads();
    }
    for (int k = 0; b < a.B; b += gridDim.x, ++k) {
        const int bn = b + gridDim.x;
        int64_t src_n = src; float opz_n = opz;
        if (bn < a.B) { src_n = a.perm ? a.perm[cur + bn] : cur + bn; opz_n = 1.0f + a.zq[src_n]; }
        const float g = TABLE ? s_g[k & 1] : 0.0f;
        // next row's factor: one thread, at the head of the row, so that only its own warp's loads wait for the powf
        if (TABLE && threadIdx.x == 0 && bn < a.B) s_g[(k + 1) & 1] = powf(opz_n, a.lw.be);
        const size_t so = (size_t)src * P, dofs = (size_t)b * P;
        for (int i0 = threadIdx.x; i0 < P; i0 += 4 * blockDim.x) {
            float fl[4], er[4]; uint8_t mk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * blockDim.x;
                if (i < P) {
                    if (delta_out) fl[q] = flux[so + i];
                    if (error_out) er[q] = error[so + i];
                    if (mask_out) mk[q] = mask[so + i];
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * blockDim.x;
                if (i < P) {
                    float A = 1.0f;
                    if (i < Nb) {
                        if (TABLE) {
                            if (zabs_out) zabs_out[(size_t)b * Nb + i] = fmaf(opz, T2[i], -1.0f);             // dataloader.py:102
                            A = expf(-fmaf(g, T1[i], T0[i]));
                        } else {
                            const float wv = a.wav[i];
                            if (zabs_out) zabs_out[(size_t)b * Nb + i] = opz * wv / 1215.67f - 1.0f;          // dataloader.py:102
                            A = expf(-tau_total_px(wv, opz, a.lw, a.max_series));
                        }
                    }
                    if (delta_out) delta_out[dofs + i] = fl[q] - mu[i] * A;                                    // dataloader.py:135-136
                    if (error_out) error_out[dofs + i] = er[q];
                    if (mask_out) mask_out[dofs + i] = mk[q];
                }
            }
        }
        if (TABLE) __syncthreads();                 // s_g[(k + 1) & 1] is visible, s_g[k & 1] may be rewritten
        src = src_n; opz = opz_n;
    }
}

// sums[0][i] = sum_b flux * exp(+tau_total) * mask ; sums[1][i] = #{b : flux != -999}      dataloader.py:110-111
// one thread per pixel column, a slab of rows per blockIdx.y; double atomics (the order does not matter at 1e-16)
__global__ void __launch_bounds__(128) k_tau_weight_sums(const float* flux, const uint8_t* mask, const float* zq, const float* wav,
                                                         int N, int Nb, int P, int max_series, Law lw, double* sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int rows_per = (N + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(N, r0 + rows_per);
    const float wv = wav[i];
    double num = 0.0, den = 0.0;
    for (int b = r0; b < r1; ++b) {
        const float fl = flux[(size_t)b * P + i];
        const bool mk = mask[(size_t)b * P + i] != 0;
        if (fl != -999.0f) den += 1.0;
        if (mk) {
            float s = 1.0f;
            if (i < Nb) s = expf(tau_total_px(wv, 1.0f + zq[b], lw, max_series));
            num += (double)fl * (double)s;
        }
    }
    atomicAdd(sums + i, num);
    atomicAdd(sums + P + i, den);
}

// ---------------------------------------------------------------------------------------------------------------------
// OOD scoring on a vector of per-spectrum NLLs (larger = less likely under the model).
//   count_out[0] = #{b : nll[b] > threshold}; the first min(count, cap) of them (ascending index order is NOT guaranteed by the
//   atomic compaction, so they are sorted by index afterwards by the caller if it cares) go to thr_idx.
//   top-k: EXACT k largest values by a 4-pass radix select on the order-preserving integer image of the floats, then a
//   bitonic sort of the k winners in shared memory (descending value, ties by ascending index).  One CTA: B <= ~1e7 is a few
//   passes over an L2-resident vector.  NaNs sort above everything (they are the most suspicious scores).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kTopKMax = 2048;
__device__ __forceinline__ uint32_t f2key(float x) {
    uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // monotone: larger float -> larger key; NaN (positive) on top
}

__global__ void __launch_bounds__(1024) k_ood_select(const float* nll, int B, float threshold, int k, int thr_cap,
                                                     int* count_out, int* thr_idx, int* top_idx, float* top_val) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_want, s_n_gt, s_n_eq;
    __shared__ uint64_t skey[kTopKMax];           // (key << 32) | ~index : descending sort gives value desc, index asc
    const int tid = threadIdx.x, NT = blockDim.x;
    // ---- threshold pass
    if (tid == 0) { s_n_gt = 0; }
    __syncthreads();
    if (count_out) {
        for (int b0 = 0; b0 < B; b0 += NT) {                 // warp-aggregated compaction: one shared atomic per warp and trip
            const int b = b0 + tid;
            const float v = b < B ? nll[b] : 0.f;
            const bool hit = b < B && (v > threshold || v != v);
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (m) {
                uint32_t base = 0;
                if ((tid & 31) == (__ffs(m) - 1)) base = atomicAdd(&s_n_gt, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                const uint32_t slot = base + __popc(m & ((1u << (tid & 31)) - 1u));
                if (hit && thr_idx && (int)slot < thr_cap) thr_idx[slot] = b;
            }
        }
        __syncthreads();
        if (tid == 0) count_out[0] = (int)s_n_gt;
    }
    if (k <= 0 || !top_idx) return;
    if (k > B) k = B;
    if (k > kTopKMax) k = kTopKMax;
    // ---- radix select: find the key T of the k-th largest element
    if (tid == 0) { s_prefix = 0; s_want = (uint32_t)k; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int q = tid; q < 256; q += NT) hist[q] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        // NLLs of one data set share their leading bits: almost every element hits the SAME bin in the first passes, so the
        // counts are aggregated per warp before they touch shared memory (one atomic per distinct digit and warp)
        for (int b0 = 0; b0 < B; b0 += NT) {
            const int b = b0 + tid;
            const uint32_t key = b < B ? f2key(nll[b]) : 0u;
            const bool in = b < B && (key & pmask) == prefix;
            const uint32_t digit = (key >> shift) & 255u;
            const uint32_t act = __ballot_sync(0xffffffffu, in);
            if (in) {
                const uint32_t peers = __match_any_sync(act, digit);
                if ((tid & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[digit], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t want = s_want, d = 255;
            for (;; --d) {                         // walk down from the largest digit
                if (hist[d] >= want) break;
                want -= hist[d];
                if (d == 0) break;
            }
            s_prefix = prefix | (d << shift);
            s_want = want;                          // rank inside the chosen digit
        }
        __syncthreads();
    }
    const uint32_t T = s_prefix;                    // key of the k-th largest
    const uint32_t need_eq = s_want;                // how many elements equal to T belong to the top k
    if (tid == 0) { s_n_gt = 0; s_n_eq = 0; }
    __syncthreads();
    // elements with key > T: all of them (k - need_eq); elements == T: the need_eq with the smallest indices -- the
    // compaction takes ANY need_eq of them here and the tie rule is restored by scanning in index order per thread block
    // stripe: simple and exact because ties are resolved by a second ordered pass below.
    for (int b = tid; b < B; b += NT) {
        const uint32_t key = f2key(nll[b]);
        if (key > T) {
            const uint32_t slot = atomicAdd(&s_n_gt, 1u);
            skey[slot] = ((uint64_t)key << 32) | (uint32_t)(~(uint32_t)b);
        }
    }
    __syncthreads();
    // ties at the k-th value: the need_eq SMALLEST indices among the elements equal to T.  Rounds of "block-wide minimum index
    // above the last one taken" (need_eq is almost always 1: one round)
    __shared__ uint32_t s_min;
    {
        uint32_t last = 0xFFFFFFFFu;                // index taken in the previous round (as +1 offset logic below)
        int lower = 0;                              // candidates must have index >= lower
        for (uint32_t got = 0; got < need_eq; ++got) {
            if (tid == 0) s_min = 0xFFFFFFFFu;
            __syncthreads();
            uint32_t mine = 0xFFFFFFFFu;
            for (int b = lower + tid; b < B; b += NT)
                if (f2key(nll[b]) == T) { mine = (uint32_t)b; break; }          // this thread's smallest candidate (strided order)
            if (mine != 0xFFFFFFFFu) atomicMin(&s_min, mine);
            __syncthreads();
            last = s_min;
            if (tid == 0) skey[s_n_gt + got] = ((uint64_t)T << 32) | (uint32_t)(~last);
            lower = (int)last + 1;
            __syncthreads();
        }
        (void)last;
    }
    __syncthreads();
    // ---- bitonic sort (descending) of the k entries, padded with zeros to a power of two
    int n2 = 1;
    while (n2 < k) n2 <<= 1;
    for (int q = k + tid; q < n2; q += NT) skey[q] = 0ull;
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int q = tid; q < n2; q += NT) {
                const int p = q ^ stride;
                if (p > q) {
                    const bool desc = (q & size) == 0;
                    const uint64_t a = skey[q], c = skey[p];
                    if (desc ? a < c : a > c) { skey[q] = c; skey[p] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int q = tid; q < k; q += NT) {
        const int b = (int)(~(uint32_t)(skey[q] & 0xFFFFFFFFull));
        top_idx[q] = b;
        if (top_val) top_val[q] = nll[b];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// posterior samples: h = hmean + L z, L L^T = hcov (lower Cholesky, float), z ~ N(0, I) from Philox4x32-10 keyed by
// (seed, spectrum, sample); optionally the continuum sample mu + F h on the full grid (nb/predict.ipynb cell 11).
// One CTA per spectrum; Nh <= 32.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sample_posterior(const float* hmean, const float* hcov, const float* F, const float* mu,
                                                          int B, int Nh, int P, int S, unsigned long long seed,
                                                          float* z_out, float* h_out, float* cont_out) {
    __shared__ float sL[32 * 33];
    __shared__ float sh[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (b >= B) return;
    for (int q = tid; q < Nh * Nh; q += blockDim.x) sL[(q / Nh) * 33 + (q % Nh)] = hcov[(size_t)b * Nh * Nh + q];
    __syncthreads();
    if (tid < 32) {                                    // warp Cholesky, lane = row (in place, lower triangle)
        const int r = tid;
        for (int j = 0; j < Nh; ++j) {
            float d = sL[j * 33 + j];
            d = d > 0.f ? sqrtf(d) : 0.f;
            __syncwarp();
            if (r == j) sL[j * 33 + j] = d;
            const float inv = d > 0.f ? 1.0f / d : 0.f;
            if (r > j && r < Nh) sL[r * 33 + j] *= inv;
            __syncwarp();
            if (r > j && r < Nh) {
                const float lrj = sL[r * 33 + j];
                for (int c = j + 1; c <= r; ++c) sL[r * 33 + c] -= lrj * sL[c * 33 + j];
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int s = 0; s < S; ++s) {
        if (tid < Nh) {
            curandStatePhilox4_32_10_t st;
            curand_init(seed, (unsigned long long)((size_t)b * S + s) * 32ull + tid, 0ull, &st);
            const float z = curand_normal(&st);
            if (z_out) z_out[((size_t)b * S + s) * Nh + tid] = z;
            sh[tid] = z;
        }
        __syncthreads();
        float hval = 0.f;
        if (tid < Nh) {
            hval = hmean[(size_t)b * Nh + tid];
            for (int c = 0; c <= tid; ++c) hval = fmaf(sL[tid * 33 + c], sh[c], hval);
        }
        __syncthreads();                               // every row has read z before sh is overwritten with h
        if (tid < Nh) {
            if (h_out) h_out[((size_t)b * S + s) * Nh + tid] = hval;
            sh[tid] = hval;
        }
        __syncthreads();
        if (cont_out) {
            for (int i = tid; i < P; i += blockDim.x) {
                float c = mu[i];
                for (int k = 0; k < Nh; ++k) c = fmaf(F[(size_t)i * Nh + k], sh[k], c);
                cont_out[((size_t)b * S + s) * P + i] = c;
            }
        }
        __syncthreads();
    }
}

}  // namespace aux
}  // namespace qfa
