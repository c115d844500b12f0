// CUDA-core kernels of the QFA hot path (float / double), sm_100a.
//
//   k_gram_solve  (spectrum-major)  per spectrum: weights -> masked weighted Grams
//                 M = I + F^T diag(w A^2) F, M2 = F^T diag(w A^3) F, b, b2 -> Cholesky ->
//                 NLL, hmean, hcov [, continuum, sigma]            (model.py:121-135, 161-180)
//   k_grad        (pixel-major)     per (spectrum, pixel): Sigma^-1 delta, diag Sigma^-1 and
//                 the reference-parity partials, accumulated over spectra in registers
//                 with a fixed thread<->pixel ownership (no atomics)  (model.py:136-150)
//   k_reduce      folds the per-split partials + per-spectrum NLL into the `acc` buffer
//                 (model.py:98-103)
//
// Algebra: SURVEY.md section 7.1.  All summation orders are fixed => deterministic.
#pragma once
#include "qfa_common.cuh"

namespace qfa {

enum { MODE_TRAIN = 0, MODE_PREDICT = 1, MODE_NLL = 2 };

template <int HP> struct SmallLayout {
    // per-spectrum hand-off from k_gram_solve to k_grad, in T units
    static constexpr int a = 0;                 // hmean = M^-1 b            [HP]
    static constexpr int c = HP;                // c = b2 - M2 a             [HP]
    static constexpr int Linv = 2 * HP;         // L^-1 (lower, row-major)   [HP*HP]
    static constexpr int K = 2 * HP + HP * HP;  // K = M^-1 M2 (row-major)   [HP*HP]
    static constexpr int len = 2 * HP + 2 * HP * HP;
};

// ---------------------------------------------------------------------------------------
// Warp-level dense algebra on an HP x HP SPD matrix held in shared memory (one warp).
// In : sM = M (full, symmetric), sM2 = M2 (TRAIN), sb, sb2.
// Out: sL = L^-1 (lower), sM = M^-1, sM2 = K = M^-1 M2 (TRAIN), sa = M^-1 b, sc = b2 - M2 a,
//      sout[0] = log det M, sout[1] = b^T M^-1 b.
// ---------------------------------------------------------------------------------------
template <typename T, int HP, bool TRAIN>
__device__ __noinline__ void small_algebra(T* sM, T* sM2, T* sL, T* sb, T* sb2, T* sa, T* sc, T* sout) {
    constexpr int LD = HP + 1;
    const int lane = threadIdx.x & 31;
    T logdet = T(0);
    for (int j = 0; j < HP; ++j) {                       // right-looking Cholesky, lane <-> row
        T d = Mth<T>::sqrt(sM[j * LD + j]);
        logdet += Mth<T>::log(d);
        T inv = Mth<T>::rcp(d);
        __syncwarp();
        if (lane == j) sM[j * LD + j] = d;
        if (lane > j && lane < HP) sM[lane * LD + j] *= inv;
        __syncwarp();
        if (lane > j && lane < HP) {
            T lij = sM[lane * LD + j];
            for (int k = j + 1; k <= lane; ++k) sM[lane * LD + k] -= lij * sM[k * LD + j];
        }
        __syncwarp();
    }
    logdet *= T(2);
    T x[HP];                                             // column `lane` of L^-1
#pragma unroll
    for (int r = 0; r < HP; ++r) {
        T s = (r == lane) ? T(1) : T(0);
#pragma unroll
        for (int k = 0; k < r; ++k) s -= sM[r * LD + k] * x[k];
        x[r] = s * Mth<T>::rcp(sM[r * LD + r]);
    }
    if (lane < HP) {
#pragma unroll
        for (int r = 0; r < HP; ++r) sL[r * LD + lane] = x[r];
    }
    __syncwarp();
    {                                                    // M^-1 = L^-T L^-1, lane <-> column
        T col[HP];
#pragma unroll
        for (int r = 0; r < HP; ++r) {
            T s = T(0);
#pragma unroll
            for (int k = r; k < HP; ++k) s += sL[k * LD + r] * x[k];
            col[r] = s;
        }
        if (lane < HP) {
#pragma unroll
            for (int r = 0; r < HP; ++r) sM[r * LD + lane] = col[r];
        }
    }
    __syncwarp();
    T av = T(0);
    if (lane < HP) {
        for (int k = 0; k < HP; ++k) av += sM[lane * LD + k] * sb[k];
        sa[lane] = av;
    }
    T quad = warp_sum((lane < HP) ? av * sb[lane] : T(0));
    __syncwarp();
    if (TRAIN) {
        T m2c[HP];
#pragma unroll
        for (int k = 0; k < HP; ++k) m2c[k] = (lane < HP) ? sM2[k * LD + lane] : T(0);
        if (lane < HP) {
            T cv = sb2[lane];
            for (int k = 0; k < HP; ++k) cv -= sM2[lane * LD + k] * sa[k];
            sc[lane] = cv;
        }
        __syncwarp();
        T kc[HP];
#pragma unroll
        for (int r = 0; r < HP; ++r) {
            T s = T(0);
#pragma unroll
            for (int k = 0; k < HP; ++k) s += sM[r * LD + k] * m2c[k];
            kc[r] = s;
        }
        if (lane < HP) {
#pragma unroll
            for (int r = 0; r < HP; ++r) sM2[r * LD + lane] = kc[r];
        }
    }
    if (lane == 0) { sout[0] = logdet; sout[1] = quad; }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------
// k_gram_solve
// ---------------------------------------------------------------------------------------
template <typename T>
struct GramArgs {
    Field<T> f;
    int B;
    T* small;        // TRAIN: [B][SmallLayout::len]
    T* nll;          // [B]
    float* hasblue;  // TRAIN: [B] 1.0 if the spectrum has >= 1 unmasked blue pixel
    T* hmean;        // PREDICT (optional): [B][Nh]
    T* hcov;         // PREDICT (optional): [B][Nh][Nh]
    T* cont;         // PREDICT (optional): [B][P]
    T* unc;          // PREDICT (optional): [B][P]
};

template <typename T, int HP, int MODE>
struct GramCfg {
    static constexpr int NT = 256;
    static constexpr int PC = (HP >= 32) ? 128 : 256;          // pixels per chunk
    static constexpr int TB = 4;                               // register tile edge
    static constexpr int NB1 = HP / TB;
    static constexpr int NBK = NB1 * (NB1 + 1) / 2;            // upper-triangular 4x4 blocks
    static constexpr int NSL = NT / NBK;                       // pixel slices
    static constexpr bool TRAIN = (MODE == MODE_TRAIN);
    static constexpr int NW = TRAIN ? 4 : 2;                   // weight arrays per chunk
    static constexpr int NACC = TRAIN ? 40 : 20;               // accumulators per thread
    static constexpr int LD = HP + 1;
    // streaming part in T, the per-spectrum H x H algebra always in double (its cost is negligible and it removes the
    // factorisation error: what is left in float mode is the rounding of the Gram entries themselves)
    static constexpr size_t smem_stream = ((size_t)PC * HP + (size_t)NW * PC + (size_t)NACC * NT + 4 * 32) * sizeof(T);
    static constexpr size_t smem_alg = (3 * (size_t)HP * LD + 4 * HP + 8) * sizeof(double);
    static constexpr size_t smem_bytes = ((smem_stream + 15) / 16) * 16 + smem_alg;
};

template <typename T, int HP, int MODE>
#ifndef QFA_GRAMSOLVE_MINB
#define QFA_GRAMSOLVE_MINB 4
#endif
// HP <= 8, float: four CTAs per SM (64 registers per thread instead of 80, 4 x 56.8 KB of shared memory = the 227 KB an SM
// offers, see launch_gram).  Measured on one B200: a batch of 500 spectra -- the reference's default -- fits ONE wave of
// 592 CTAs instead of 1.13 waves of 444 (captured train step 140 -> 96 us), and the float mode as a whole gains
// (predict 9.2 -> 12.2, train step 4.5 -> 7.8 M spectra/s).
__global__ void __launch_bounds__(256, (HP <= 8 && sizeof(T) == 4) ? QFA_GRAMSOLVE_MINB : 1) k_gram_solve(GramArgs<T> g) {
    using C = GramCfg<T, HP, MODE>;
    constexpr int NT = C::NT, PC = C::PC, NBK = C::NBK, NSL = C::NSL, NACC = C::NACC, LD = C::LD;
    constexpr bool TRAIN = C::TRAIN;
    constexpr bool PREDICT = (MODE == MODE_PREDICT);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using TA = double;                          // algebra type
    T* sF = reinterpret_cast<T*>(smem_raw);     // [PC][HP]
    T* sW = sF + PC * HP;                       // [NW][PC]
    T* sPart = sW + C::NW * PC;                 // [NACC][NT]
    T* sred = sPart + NACC * NT;                // [4*32]
    TA* sM = reinterpret_cast<TA*>(smem_raw + ((C::smem_stream + 15) / 16) * 16);   // [HP][LD]
    TA* sM2 = sM + HP * LD;
    TA* sL = sM2 + HP * LD;
    TA* sb = sL + HP * LD;
    TA* sb2 = sb + HP;
    TA* sa = sb2 + HP;
    TA* sc = sa + HP;
    TA* sout = sc + HP;                         // [8]

    const Field<T>& f = g.f;
    const int tid = threadIdx.x;
    const int P = f.P, Nh = f.Nh;
    const T tau0 = (T)__ldg(f.scal + 0), c0 = (T)__ldg(f.scal + 1), beta = (T)__ldg(f.scal + 2);

    // tile assignment of this thread for the Gram accumulation
    const bool tile_on = tid < NSL * NBK;
    const int slice = tid / NBK;
    int bk = 0, bl = 0;
    {
        int rem = tid % NBK;
        while (rem >= C::NB1 - bk) { rem -= C::NB1 - bk; ++bk; }
        bl = bk + rem;
    }
    const bool diag = (bk == bl);

    for (int b = blockIdx.x; b < g.B; b += gridDim.x) {
        T G[4][4], G2[4][4], vb[4], vb2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            vb[i] = vb2[i] = T(0);
#pragma unroll
            for (int j = 0; j < 4; ++j) G[i][j] = G2[i][j] = T(0);
        }
        T sums[4] = {T(0), T(0), T(0), T(0)};   // sum w r^2, sum log D, n, n_blue

        for (int p0 = 0; p0 < P; p0 += PC) {
            // ---- elementwise: one pixel per thread -> weights in smem
            if (tid < PC) {
                const int i = p0 + tid;
                T s2 = T(0), s3 = T(0), wb = T(0), wb2 = T(0);
                if (i < P) {
                    Cell<T> c = eval_cell<T, PREDICT>(f, (size_t)b, i, tau0, c0, beta);
                    s2 = c.w * c.A * c.A;
                    wb = c.w * c.A * c.r;
                    if (TRAIN) { s3 = s2 * c.A; wb2 = s2 * c.r; }
                    sums[0] += c.w * c.r * c.r;
                    sums[1] += c.logD;
                    sums[2] += c.mk ? T(1) : T(0);
                    sums[3] += (c.mk && i < f.Nb) ? T(1) : T(0);
                }
                sW[0 * PC + tid] = s2;
                sW[1 * PC + tid] = wb;
                if (TRAIN) { sW[2 * PC + tid] = s3; sW[3 * PC + tid] = wb2; }
            }
            // ---- stage the F rows of this chunk (zero padded to HP columns / past P)
            for (int e = tid; e < PC * HP; e += NT) {
                const int row = e / HP, col = e % HP;
                const int i = p0 + row;
                sF[e] = (i < P && col < Nh) ? (T)__ldg(f.F + (size_t)i * Nh + col) : T(0);
            }
            __syncthreads();
            // ---- register-tiled weighted Gram: thread <-> (4x4 block, pixel slice)
            if (tile_on) {
                for (int p = slice; p < PC; p += NSL) {
                    const T* fr = sF + p * HP;
                    T fk[4], fl[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { fk[q] = fr[4 * bk + q]; fl[q] = fr[4 * bl + q]; }
                    const T s2 = sW[0 * PC + p];
                    T xk[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) xk[q] = s2 * fk[q];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int r = 0; r < 4; ++r) G[q][r] += xk[q] * fl[r];
                    if (TRAIN) {
                        const T s3 = sW[2 * PC + p];
#pragma unroll
                        for (int q = 0; q < 4; ++q) xk[q] = s3 * fk[q];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int r = 0; r < 4; ++r) G2[q][r] += xk[q] * fl[r];
                    }
                    if (diag) {
                        const T wb = sW[1 * PC + p];
#pragma unroll
                        for (int q = 0; q < 4; ++q) vb[q] += wb * fk[q];
                        if (TRAIN) {
                            const T wb2 = sW[3 * PC + p];
#pragma unroll
                            for (int q = 0; q < 4; ++q) vb2[q] += wb2 * fk[q];
                        }
                    }
                }
            }
            __syncthreads();
        }

        // ---- cross-slice reduction through shared memory (fixed order)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                sPart[(q * 4 + r) * NT + tid] = tile_on ? G[q][r] : T(0);
                if (TRAIN) sPart[(20 + q * 4 + r) * NT + tid] = tile_on ? G2[q][r] : T(0);
            }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sPart[(16 + q) * NT + tid] = tile_on ? vb[q] : T(0);
            if (TRAIN) sPart[(36 + q) * NT + tid] = tile_on ? vb2[q] : T(0);
        }
        block_sum<T, 4, NT>(sums, sred);   // contains the __syncthreads that publishes sPart
        for (int e = tid; e < NBK * NACC; e += NT) {
            const int j = e / NBK, blk = e % NBK;
            T s = T(0);
            for (int sl = 0; sl < NSL; ++sl) s += sPart[j * NT + sl * NBK + blk];
            int kb = 0, rem = blk;
            while (rem >= C::NB1 - kb) { rem -= C::NB1 - kb; ++kb; }
            const int lb = kb + rem;
            const int jj = (j >= 20) ? j - 20 : j;
            TA* dstM = (j >= 20) ? sM2 : sM;
            TA* dstb = (j >= 20) ? sb2 : sb;
            if (jj < 16) {
                const int q = jj / 4, r = jj % 4;
                const int row = 4 * kb + q, col = 4 * lb + r;
                if (kb != lb || q <= r) {
                    const TA v = (TA)s + ((j < 20 && row == col) ? TA(1) : TA(0));   // M = I + Gram
                    dstM[row * LD + col] = v;
                    dstM[col * LD + row] = v;
                }
            } else if (kb == lb) {
                dstb[4 * kb + (jj - 16)] = s;
            }
        }
        __syncthreads();
        if (tid < 32) small_algebra<TA, HP, TRAIN>(sM, sM2, sL, sb, sb2, sa, sc, sout);
        __syncthreads();

        // ---- per-spectrum outputs
        if (tid == 0) {
            // model.py:135 / 176; n*log(2 pi) kept in T (quirk Q5 is below the 1e-5 bar)
            TA nll = TA(0.5) * ((TA)sums[0] - sout[1] + (TA)sums[2] * TA(kLog2Pi) + (TA)sums[1] + sout[0]);
            g.nll[b] = (T)nll;
            if (TRAIN) g.hasblue[b] = sums[3] > T(0) ? 1.0f : 0.0f;
        }
        if (TRAIN) {
            using SL = SmallLayout<HP>;
            T* dst = g.small + (size_t)b * SL::len;
            for (int e = tid; e < SL::len; e += NT) {
                TA v;
                if (e < SL::c) v = sa[e];
                else if (e < SL::Linv) v = sc[e - SL::c];
                else if (e < SL::K) { int t = e - SL::Linv; v = sL[(t / HP) * LD + (t % HP)]; }
                else { int t = e - SL::K; v = sM2[(t / HP) * LD + (t % HP)]; }
                dst[e] = (T)v;
            }
        }
        if (PREDICT) {
            if (g.hmean && tid < Nh) g.hmean[(size_t)b * Nh + tid] = (T)sa[tid];
            if (g.hcov)
                for (int e = tid; e < Nh * Nh; e += NT)
                    g.hcov[(size_t)b * Nh * Nh + e] = (T)sM[(e / Nh) * LD + (e % Nh)];
            if (g.cont || g.unc) {
                // model.py:180: mu + F hmean and sqrt(diag(F hcov F^T)) on the FULL grid
                for (int i = tid; i < P; i += NT) {
                    T fr[HP];
#pragma unroll
                    for (int k = 0; k < HP; ++k) fr[k] = (k < Nh) ? (T)__ldg(f.F + (size_t)i * Nh + k) : T(0);
                    T fa = T(0), q = T(0);
#pragma unroll
                    for (int k = 0; k < HP; ++k) {
                        fa += fr[k] * (T)sa[k];
                        T y = T(0);
#pragma unroll
                        for (int l = 0; l <= k; ++l) y += (T)sL[k * LD + l] * fr[l];   // y = L^-1 f
                        q += y * y;
                    }
                    if (g.cont) g.cont[(size_t)b * P + i] = (T)__ldg(f.mu + i) + fa;
                    if (g.unc) g.unc[(size_t)b * P + i] = Mth<T>::sqrt(q);
                }
            }
        }
        __syncthreads();   // smem is reused by the next spectrum
    }
}

// ---------------------------------------------------------------------------------------
// k_grad: pixel-major gradient accumulation.  grid = (pixel tiles, spectrum splits)
// ---------------------------------------------------------------------------------------
template <typename T>
struct GradArgs {
    Field<T> f;
    int B;              // spectra in this launch
    int nsplit;
    const T* small;     // [B][SmallLayout::len]
    T* part;            // [nsplit][part_len]   per-pixel partial sums
    T* spart;           // [nsplit][ntiles][3]  scalar partial sums (tau0, c0, beta)
    int accumulate;     // 0: overwrite part/spart, 1: add (later sub-batches)
};

// per-split partial layout (T units): [F (P*Nh) | Psi (P) | omega (Nb) | cnt (P) | dmu (P)]
__host__ __device__ inline size_t part_len(int P, int Nb, int Nh) { return (size_t)P * Nh + 3 * (size_t)P + Nb; }

template <typename T, int HP>
__global__ void __launch_bounds__(128) k_grad(GradArgs<T> g) {
    constexpr int NT = 128;
    using SL = SmallLayout<HP>;
    __shared__ __align__(16) T ssm[2][SL::len];
    __shared__ T sred[3 * 32];
    const Field<T>& f = g.f;
    const int tid = threadIdx.x;
    const int P = f.P, Nh = f.Nh, Nb = f.Nb;
    const int i = blockIdx.x * NT + tid;
    const bool on = i < P;
    const T tau0 = (T)__ldg(f.scal + 0), c0 = (T)__ldg(f.scal + 1), beta = (T)__ldg(f.scal + 2);

    const int per = (g.B + g.nsplit - 1) / g.nsplit;
    const int b0 = blockIdx.y * per;
    const int b1 = min(g.B, b0 + per);

    T fr[HP];
#pragma unroll
    for (int k = 0; k < HP; ++k) fr[k] = (on && k < Nh) ? (T)__ldg(f.F + (size_t)i * Nh + k) : T(0);

    T gF[HP];
#pragma unroll
    for (int k = 0; k < HP; ++k) gF[k] = T(0);
    T gPsi = T(0), gOm = T(0), cnt = T(0), dmu = T(0);
    T sc3[3] = {T(0), T(0), T(0)};   // per-pixel sums for d tau0, d c0, d beta

    if (b0 < b1) {
        for (int e = tid; e < SL::len; e += NT) ssm[0][e] = g.small[(size_t)b0 * SL::len + e];
    }
    __syncthreads();
    for (int b = b0; b < b1; ++b) {
        const int cur = (b - b0) & 1;
        if (b + 1 < b1) {   // prefetch the next spectrum's hand-off into the other buffer
            for (int e = tid; e < SL::len; e += NT) ssm[cur ^ 1][e] = g.small[(size_t)(b + 1) * SL::len + e];
        }
        const T* sm = ssm[cur];
        if (on) {
            Cell<T> c = eval_cell<T, false>(f, (size_t)b, i, tau0, c0, beta);
            if (c.mk) {
                T fa = T(0), q = T(0);
#pragma unroll
                for (int k = 0; k < HP; ++k) {
                    fa += fr[k] * sm[SL::a + k];
                    T y = T(0);
#pragma unroll
                    for (int l = 0; l <= k; ++l) y += sm[SL::Linv + k * HP + l] * fr[l];
                    q += y * y;                                  // f^T M^-1 f = |L^-1 f|^2
                }
                const T A = c.A, w = c.w;
                const T u = w * (c.r - A * fa);                  // (Sigma^-1 delta)_i
                const T s2 = w * A * A, s3 = s2 * A;
                const T gd = T(0.5) * (w - w * s2 * q - u * u);  // model.py:136,138
                const T Au = A * u;
                T fK[HP];
#pragma unroll
                for (int k = 0; k < HP; ++k) fK[k] = T(0);
#pragma unroll
                for (int l = 0; l < HP; ++l)
#pragma unroll
                    for (int k = 0; k < HP; ++k) fK[k] += fr[l] * sm[SL::K + l * HP + k];
#pragma unroll
                for (int k = 0; k < HP; ++k)                     // model.py:137 (quirk Q2)
                    gF[k] += s3 * fr[k] - s2 * fK[k] - Au * sm[SL::c + k];
                gPsi += A * A * gd;                              // model.py:139
                cnt += T(1);
                dmu -= Au;
                if (i < Nb) {
                    gOm += gd * c.zdep;                          // model.py:140
                    const T root = T(1) - tau0 * c.powb - c0;    // model.py:141 (quirk Q3)
                    const T t = gd * (c.om * c.zdep) * c.zdep * T(2) * root;
                    sc3[0] -= t * c.powb;                        // model.py:142
                    sc3[1] -= t;                                 // model.py:144
                    sc3[2] -= t * (tau0 * c.powb * c.logopz);    // model.py:143
                }
            }
        }
        __syncthreads();
    }

    T* part = g.part + (size_t)blockIdx.y * part_len(P, Nb, Nh);
    if (on) {
        T* pF = part + (size_t)i * Nh;
        T* pPsi = part + (size_t)P * Nh + i;
        T* pOm = part + (size_t)P * Nh + P + i;
        T* pCnt = part + (size_t)P * Nh + P + Nb + i;
        T* pMu = part + (size_t)P * Nh + 2 * (size_t)P + Nb + i;
        if (g.accumulate) {
#pragma unroll
            for (int k = 0; k < HP; ++k) if (k < Nh) pF[k] += gF[k];
            *pPsi += gPsi; *pCnt += cnt; *pMu += dmu;
            if (i < Nb) *pOm += gOm;
        } else {
#pragma unroll
            for (int k = 0; k < HP; ++k) if (k < Nh) pF[k] = gF[k];
            *pPsi = gPsi; *pCnt = cnt; *pMu = dmu;
            if (i < Nb) *pOm = gOm;
        }
    }
    if ((int)(blockIdx.x * NT) < Nb) {   // tiles that contain blue pixels
        block_sum<T, 3, NT>(sc3, sred);
        if (tid < 3) {
            T* sp = g.spart + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 + tid;
            if (g.accumulate) *sp += sc3[tid]; else *sp = sc3[tid];
        }
    }
}

// ---------------------------------------------------------------------------------------
// k_reduce: acc += fold(partials).  One thread per acc element, fixed order.
// acc layout: see include/qfa_b200.h
// ---------------------------------------------------------------------------------------
template <typename T>
struct ReduceArgs {
    const T* part; const T* spart; const T* nll; const float* hasblue; const float* scal;
    T* acc;
    int P, Nb, Nh, B, nsplit, ntiles_blue, ntiles;   // B = length of nll / hasblue (per spectrum, or pre-folded partial sums)
    int stride;                                       // element stride of nll / hasblue
    int nsplit_red, tile_px;                          // pixels of tiles >= ntiles_blue have nsplit_red partials (tile_px pixels per tile)
    double nsp;                                       // number of spectra these sums stand for
};

template <typename T>
__global__ void __launch_bounds__(256) k_reduce(ReduceArgs<T> r) {
    const size_t PH = (size_t)r.P * r.Nh;
    const size_t plen = part_len(r.P, r.Nb, r.Nh);
    const size_t o_psi = PH, o_om = PH + r.P, o_sc = o_om + r.Nb, o_cnt = o_sc + 3,
                 o_scnt = o_cnt + r.P, o_nll = o_scnt + 3, o_nsp = o_nll + 1, o_dmu = o_nsp + 1;
    const size_t n_pix_elems = PH + 3 * (size_t)r.P + r.Nb;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n_pix_elems) {
        // map to (partial index, acc index)
        size_t pi, ai;
        if (gid < PH + r.P + r.Nb) { pi = gid; ai = gid; }                         // F, Psi, omega
        else if (gid < PH + 2 * (size_t)r.P + r.Nb) { pi = gid; ai = o_cnt + (gid - (PH + r.P + r.Nb)); }
        else { pi = gid; ai = o_dmu + (gid - (PH + 2 * (size_t)r.P + r.Nb)); }
        // pixel this element belongs to -> how many partials exist for it
        size_t px;
        if (gid < PH) px = gid / r.Nh;
        else if (gid < PH + r.P) px = gid - PH;
        else if (gid < PH + r.P + r.Nb) px = gid - PH - r.P;
        else if (gid < PH + 2 * (size_t)r.P + r.Nb) px = gid - (PH + r.P + r.Nb);
        else px = gid - (PH + 2 * (size_t)r.P + r.Nb);
        const int ns = ((int)(px / r.tile_px) < r.ntiles_blue) ? r.nsplit : r.nsplit_red;
        T s = T(0);
        for (int k = 0; k < ns; ++k) s += r.part[(size_t)k * plen + pi];
        r.acc[ai] += s;
    }
    if (blockIdx.x == gridDim.x - 1) {
        // last block additionally folds the per-spectrum scalars (block-strided, fixed tree)
        __shared__ T red[4 * 32];
        T v[2] = {T(0), T(0)};
        // contiguous per-spectrum arrays (65 536 values for one block): four elements per thread and trip, trips unrolled, so
        // that 16 loads per array are in flight instead of one (the fold was 0.13 ms of the Nh 32 step)
        int b_lo = 0;
        if (r.stride == 1) {
            const int B4 = r.B & ~3;
#pragma unroll 4
            for (int b = threadIdx.x * 4; b < B4; b += blockDim.x * 4) {
                T n4[4]; float h4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { n4[q] = r.nll[b + q]; h4[q] = r.hasblue[b + q]; }
                v[0] += (n4[0] + n4[1]) + (n4[2] + n4[3]);
                v[1] += (T)((h4[0] + h4[1]) + (h4[2] + h4[3]));
            }
            b_lo = B4;
        }
        for (int b = b_lo + threadIdx.x; b < r.B; b += blockDim.x) {
            v[0] += r.nll[(size_t)b * r.stride];
            v[1] += (T)r.hasblue[(size_t)b * r.stride];
        }
        block_sum<T, 2, 256>(v, red);
        T s3[3] = {T(0), T(0), T(0)};
        const int nsp = r.nsplit * r.ntiles;
        for (int e = threadIdx.x; e < nsp; e += blockDim.x) {
            if ((e % r.ntiles) < r.ntiles_blue) {
                s3[0] += r.spart[(size_t)e * 3 + 0];
                s3[1] += r.spart[(size_t)e * 3 + 1];
                s3[2] += r.spart[(size_t)e * 3 + 2];
            }
        }
        block_sum<T, 3, 256>(s3, red);
        if (threadIdx.x == 0) {
            r.acc[o_sc + 0] += s3[0];
            r.acc[o_sc + 1] += s3[1];
            r.acc[o_sc + 2] += s3[2];
            const T tau0 = (T)r.scal[0];
            r.acc[o_scnt + 0] += v[1];                                  // tau0
            r.acc[o_scnt + 1] += v[1];                                  // c0
            r.acc[o_scnt + 2] += (tau0 != T(0)) ? v[1] : T(0);          // beta: every term carries tau0
            r.acc[o_nll] += v[0];
            r.acc[o_nsp] += (T)r.nsp;
        }
    }
}

}  // namespace qfa
