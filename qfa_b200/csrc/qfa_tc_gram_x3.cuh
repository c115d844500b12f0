// 3xTF32 variant of the spectrum-major tensor-core kernel (qfa_tc_gram.cuh) for Nh <= 8: QFA_FLAG_TF32X3, sm_100a.
//
// Same algebra, same Khatri-Rao GEMM formulation, same CTA organisation (15 worker warps generate the operand tiles,
// one control warp issues every bulk copy and every tcgen05.mma, one thread per spectrum solves the 8 x 8 system), but
// every product  sum_k A_k B_k  is evaluated as  A_hi B_hi + A_lo B_hi + A_hi B_lo  with  x_hi = tf32(x),
// x_lo = tf32(x - x_hi): the operand rounding error drops from 2^-11 to ~2^-21 and the result is a float-accurate
// contraction from the tensor cores (fp32 accumulation in TMEM).  The per-spectrum 8 x 8 algebra runs in double, like
// in the float CUDA-core kernels.
//
// How the split is laid out (no extra shared memory for the generated operands): a K-block holds 16 pixels instead of
// 32, and the 32 columns of a 128-byte operand row are  [ hi(16 pixels) | lo(16 pixels) ].  The static image of a
// K-block comes as two images,  I1 = [ hi | hi ]  and  I2 = [ lo | 0 ]:
//        A(k-steps 0..3) x I1(k-steps 0..3)  =  A_hi B_hi + A_lo B_hi
//        A(k-steps 0..1) x I2(k-steps 0..1)  =  A_hi B_lo                      -> 6 MMAs (K = 8) per 16 pixels.
// Lane mapping of a worker warp: lanes 0..15 / 16..31 process two DIFFERENT rows of the warp's eight (pixel = lane & 15).
// Phase O (continuum / sigma GEMM, PREDICT) uses the same trick on its K axis (44 entries = 8 a + 36 Minv, three blocks
// of 16): static pixel image [ hi | lo ], per-spectrum operand as I1 = [ hi | hi ] and I2 = [ lo | 0 ].
// TRAIN: the Grams (M, M2, b, b2) come from this kernel; the hand-off is the float record of the CUDA-core gradient
// kernel k_grad<float, 8> (qfa_kernels.cuh), because single-pass TF32 errors of the gradient GEMMs do not average out
// over the handful of spectra of a small batch.
#pragma once
#include "qfa_tc_gram.cuh"
#include "qfa_kernels.cuh"

namespace qfa {
namespace tcx {

using namespace tc;
using namespace tcg;

constexpr int XKB = 16;                          // pixels per K-block
constexpr int XPB_TILE = 2 * PB_TILE;            // I1 | I2 of one K-block (12 KB)
constexpr int XNPB = 4;                          // shared-memory slots of the static image ring
constexpr int XNSTAGE = 2;
constexpr int XQA_BLK = 3;                       // phase O: K blocks of 16 entries
constexpr int XQA_TILE = XQA_BLK * PT * 128;     // 48 KB static pixel operand (single slot)
constexpr int XB2_IMG = XQA_BLK * A_TILE;        // 48 KB per per-spectrum image
constexpr int XB2_TILE = 2 * XB2_IMG;            // I1 | I2

template <int MODE> struct XCfg {
    static constexpr bool TRAIN = (MODE == TC_TRAIN);
    static constexpr int NOPS = TRAIN ? 4 : 2;
    static constexpr int STAGE_BYTES = NOPS * A_TILE;
    static constexpr int PB_OFF = XNSTAGE * STAGE_BYTES;
    static constexpr int RING_BYTES = PB_OFF + XNPB * XPB_TILE;
    static constexpr int OVERLAY_BYTES = XB2_TILE + XQA_TILE + 2 * SO_BYTES;
    static constexpr int SMEM_BYTES = (RING_BYTES > OVERLAY_BYTES ? RING_BYTES : OVERLAY_BYTES) + 1024;
    // TMEM: the fp32 accumulation of the tensor core TRUNCATES (round-toward-zero), so every accumulate step biases the sum by
    // ~2^-25 of its value: measured 2e-5 on M^-1 after the 720 steps of a 1913-pixel spectrum.  Two counter-measures: the
    // small cross terms (A_lo B_hi + A_hi B_lo) go to accumulators of their own ("LO": their truncation is 2^-11 smaller in
    // absolute terms), and the K-blocks are dealt round-robin to NSET independent accumulator sets that phase S adds up in
    // double -- the big accumulators see 1 / (3 NSET) of the steps.
    static constexpr int NSET = TRAIN ? 2 : 3;
    static constexpr int HALF = TRAIN ? 128 : 64;                 // columns of the main (resp. LO) accumulators of a set
    static constexpr int SETW = 2 * HALF;
    static constexpr int D1A = 0, D1B = 48, D1C = 64, D1D = 112;  // inside a half: M 48 | b 16 | (TRAIN: M2 48 | b2 16)
    static constexpr int D2 = TRAIN ? 0 : NSET * SETW;            // phase O: 2 x (fa 32 | q 32)
    static constexpr int TMEM_COLS = 512;
};

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = tf32_rna(v);
    lo = tf32_rna(v - hi);
}

// entry e of the phase-O K axis for pixel i: 0..7 -> F_ie, 8..43 -> F_ik F_il, 44..47 -> 0
__device__ __forceinline__ float qa_entry(const float* __restrict__ F, int P, int Nh, int i, int e) {
    if (i >= P) return 0.f;
    if (e < HP) return e < Nh ? __ldg(F + (size_t)i * Nh + e) : 0.f;
    if (e < HP + NP2) {
        int k, l; kl_of(e - HP, k, l);
        return (k < Nh && l < Nh) ? __ldg(F + (size_t)i * Nh + k) * __ldg(F + (size_t)i * Nh + l) : 0.f;
    }
    return 0.f;
}

// PBX[kb]: I1 then I2, each 48 rows x 32 columns (rows as in k_tc_build_images); QAX[pt]: 3 blocks x 128 pixel rows x 32
__global__ void k_tc_build_images_x3(const float* __restrict__ F, int P, int Nh, float* __restrict__ PBX, int nkb,
                                     float* __restrict__ QAX, int npt) {
    const size_t n_pb = (size_t)nkb * PB_ROWS * XKB;
    const size_t n_qa = QAX ? (size_t)npt * XQA_BLK * PT * XKB : 0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_pb + n_qa; e += (size_t)gridDim.x * blockDim.x) {
        if (e < n_pb) {
            const int c = (int)(e % XKB), row = (int)((e / XKB) % PB_ROWS), kb = (int)(e / (XKB * PB_ROWS));
            const int i = kb * XKB + c;
            float v = 0.f;
            if (i < P) {
                if (row < NP2) {
                    int k, l; kl_of(row, k, l);
                    if (k < Nh && l < Nh) v = __ldg(F + (size_t)i * Nh + k) * __ldg(F + (size_t)i * Nh + l);
                } else if (row >= PB_FROW && row < PB_FROW + HP) {
                    if (row - PB_FROW < Nh) v = __ldg(F + (size_t)i * Nh + (row - PB_FROW));
                }
            }
            float hi, lo; split_tf32(v, hi, lo);
            float* img = PBX + (size_t)kb * (XPB_TILE / 4);
            img[sw128_offset(row, c) / 4] = hi;
            img[sw128_offset(row, 16 + c) / 4] = hi;
            img[PB_TILE / 4 + sw128_offset(row, c) / 4] = lo;
            img[PB_TILE / 4 + sw128_offset(row, 16 + c) / 4] = 0.f;
        } else {
            const size_t q = e - n_pb;
            const int c = (int)(q % XKB), row = (int)((q / XKB) % PT), blk = (int)((q / (XKB * PT)) % XQA_BLK),
                      pt = (int)(q / (XKB * PT * XQA_BLK));
            float hi, lo; split_tf32(qa_entry(F, P, Nh, pt * PT + row, blk * XKB + c), hi, lo);
            float* img = QAX + (size_t)pt * (XQA_TILE / 4) + (size_t)blk * (PT * 128 / 4);
            img[sw128_offset(row, c) / 4] = hi;
            img[sw128_offset(row, 16 + c) / 4] = lo;
        }
    }
}

// ---------------------------------------------------------------------------------------
// 8 x 8 SPD algebra of one spectrum in double registers (one thread)
// ---------------------------------------------------------------------------------------
struct SolvedD {
    double a[HP];
    double Minv[NP2];       // packed upper (k <= l)
    double Li[HP][HP];      // L^-1 (lower)
    double logdet, quad;
};

__device__ __noinline__ void solve_spd8_d(const double (&G)[NP2], const double (&bv)[HP], SolvedD& s) {
    double L[HP][HP];
#pragma unroll
    for (int r = 0; r < HP; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) L[r][c] = G[tri(c, r)] + (r == c ? 1.0 : 0.0);
    double invd[HP];
    double ld = 0.0;
#pragma unroll
    for (int j = 0; j < HP; ++j) {
        const double d2 = L[j][j];
        ld += log(d2);
        const double inv = rsqrt(d2);
        invd[j] = inv;
#pragma unroll
        for (int r = j + 1; r < HP; ++r) L[r][j] *= inv;
#pragma unroll
        for (int r = j + 1; r < HP; ++r)
#pragma unroll
            for (int c = j + 1; c <= r; ++c) L[r][c] -= L[r][j] * L[c][j];
    }
    s.logdet = ld;
#pragma unroll
    for (int c = 0; c < HP; ++c) {
#pragma unroll
        for (int r = 0; r < c; ++r) s.Li[r][c] = 0.0;
        s.Li[c][c] = invd[c];
#pragma unroll
        for (int r = c + 1; r < HP; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int k = c; k < r; ++k) acc -= L[r][k] * s.Li[k][c];
            s.Li[r][c] = acc * invd[r];
        }
    }
    double y[HP];
    double quad = 0.0;
#pragma unroll
    for (int r = 0; r < HP; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c <= r; ++c) acc += s.Li[r][c] * bv[c];
        y[r] = acc;
        quad += acc * acc;
    }
    s.quad = quad;
#pragma unroll
    for (int c = 0; c < HP; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int r = c; r < HP; ++r) acc += s.Li[r][c] * y[r];
        s.a[c] = acc;
    }
#pragma unroll
    for (int k = 0; k < HP; ++k)
#pragma unroll
        for (int l = k; l < HP; ++l) {
            double acc = 0.0;
#pragma unroll
            for (int r = l; r < HP; ++r) acc += s.Li[r][k] * s.Li[r][l];
            s.Minv[tri(k, l)] = acc;
        }
}

struct TcGramX3Args {
    Field<float> f;
    int B;
    TileSched ts;
    int ntiles, nkb, npt;          // nkb = ceil(P / 16)
    const float* PB;               // [nkb][XPB_TILE/4]
    const float* QA;               // [npt][XQA_TILE/4]   (PREDICT with continuum / sigma)
    float* nll;                    // [B]
    float* hmean; float* hcov; float* cont; float* unc;      // PREDICT (optional)
    float* small;                  // TRAIN: [B][SmallLayout<8>::len] hand-off record of k_grad<float, 8>
    float* hasblue;                // TRAIN: [B]
};

struct XRow { float x, e, z; unsigned m; };
struct XBuf { XRow r[4]; float psi, mu, om; };

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) k_tc_gram_x3(const TcGramX3Args g) {
    using C = XCfg<MODE>;
    constexpr bool TRAIN = C::TRAIN;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* ring = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full[XNSTAGE], bar_empty[XNSTAGE], bar_pb[XNPB], bar_gram, bar_qa, bar_tm_full[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float sE[TS];
    __shared__ float sNb[TS];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const Field<float>& f = g.f;
    const int P = f.P, Nb = f.Nb, Nh = f.Nh;

    if (tid == 0) {
        for (int s = 0; s < XNSTAGE; ++s) { mbar_init(&bar_empty[s], 1); mbar_init(&bar_full[s], NWW); }
        for (int s = 0; s < XNPB; ++s) mbar_init(&bar_pb[s], 1);
        mbar_init(&bar_gram, 1);
        mbar_init(&bar_qa, 1);
        for (int s = 0; s < 2; ++s) mbar_init(&bar_tm_full[s], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<C::TMEM_COLS>(&tmem_base_s);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    PhysConst pc;
    {
        const float tau0 = __ldg(f.scal + 0), c0 = __ldg(f.scal + 1);
        pc.beta = __ldg(f.scal + 2);
        pc.one_m_c0 = 1.0f - c0;
        pc.nt0l2e = -tau0 * kLog2e;
        pc.l2zn = f.llogzn * kLog2e;
        pc.lt0 = f.lt0; pc.lbe = f.lbe; pc.lC = f.lC;
    }
    const bool want_o = (!TRAIN) && (g.cont != nullptr || g.unc != nullptr);
    const int nkb = g.nkb, npt = g.npt;
    constexpr int LEAD = XNPB - XNSTAGE;

    auto issue_pb = [&](uint32_t git, int kb) {
        const int slot = git % XNPB;
        mbar_expect_tx(&bar_pb[slot], XPB_TILE);
        bulk_g2s(ring + C::PB_OFF + (size_t)slot * XPB_TILE, g.PB + (size_t)kb * (XPB_TILE / 4), XPB_TILE, &bar_pb[slot]);
    };

    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++tile_iter) {
        const int b0 = g.ts.first(tile);
        const int nrows = g.B - b0 < g.ts.rows(tile) ? g.B - b0 : g.ts.rows(tile);
        const int nr = warp < NWW ? (nrows - RPW * warp < 0 ? 0 : (nrows - RPW * warp > RPW ? RPW : nrows - RPW * warp)) : 0;
        const uint32_t git0 = tile_iter * (uint32_t)nkb;
        // =========================================================== phase G
        if (warp == NWW) {
            if (elect_one()) {
                for (int kb = 0; kb < LEAD && kb < nkb; ++kb) issue_pb(git0 + kb, kb);
                const uint32_t idA = idesc_tf32(128, 48), idB = idesc_tf32(128, 16);
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t git = git0 + kb;
                    const int s = git % XNSTAGE, slot = git % XNPB;
                    const uint32_t sb = smem_u32(ring) + (uint32_t)s * (uint32_t)C::STAGE_BYTES;
                    const uint32_t pb = smem_u32(ring) + (uint32_t)C::PB_OFF + (uint32_t)slot * (uint32_t)XPB_TILE;
                    mbar_wait_or_trap(&bar_full[s], (git / XNSTAGE) & 1);
                    fence_proxy_async_issuer();
                    mbar_wait_or_trap(&bar_pb[slot], (git / XNPB) & 1);
                    fence_after_sync();
                    const uint64_t dS2 = desc_sw128_kmajor(sb), dWb = desc_sw128_kmajor(sb + A_TILE);
                    const bool red = TRAIN && kb * XKB >= Nb;
                    const uint64_t dS3 = red ? dS2 : desc_sw128_kmajor(sb + 2 * A_TILE);
                    const uint64_t dW2 = red ? dWb : desc_sw128_kmajor(sb + 3 * A_TILE);
                    const uint32_t tset = tmem + (uint32_t)(kb % C::NSET) * C::SETW;
                    const bool first_of_set = kb < C::NSET;
#pragma unroll
                    for (int img = 0; img < 2; ++img) {
                        const uint64_t dP = desc_sw128_kmajor(pb + img * PB_TILE), dF = desc_sw128_kmajor(pb + img * PB_TILE + 32 * 128);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (img == 1 && kk >= 2) continue;               // I2 = [ lo | 0 ]
                            const uint64_t ko = (uint64_t)(2 * kk);
                            const bool main_acc = img == 0 && kk < 2;         // A_hi B_hi ; everything else is a small cross term
                            const uint32_t t = tset + (main_acc ? 0 : C::HALF);
                            const bool acc = !(first_of_set && (main_acc ? kk == 0 : (img == 0 && kk == 2)));
                            umma_tf32(t + C::D1A, dS2 + ko, dP + ko, idA, acc);
                            umma_tf32(t + C::D1B, dWb + ko, dF + ko, idB, acc);
                            if (TRAIN) {
                                umma_tf32(t + C::D1C, dS3 + ko, dP + ko, idA, acc);
                                umma_tf32(t + C::D1D, dW2 + ko, dF + ko, idB, acc);
                            }
                        }
                    }
                    umma_commit(&bar_empty[s]);
                    if (kb == nkb - 1) umma_commit(&bar_gram);
                    if (kb + LEAD < nkb) issue_pb(git + LEAD, kb + LEAD);
                }
            }
            __syncwarp();
            named_bar_sync(1, NTHREADS);
        } else {
            const int half = lane >> 4, px = lane & 15;
            float E[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t nbm = 0u;
            const size_t row0 = (size_t)b0 + (size_t)(RPW * warp);
            auto load_kb = [&](int kb, XBuf& k) {
                const int i = kb * XKB + px;
                const bool inr = i < P;
                const int ic = inr ? i : P - 1;
                k.psi = __ldg(f.Psi + ic);
                k.mu = TRAIN ? 0.0f : __ldg(f.mu + ic);
                k.om = (i < Nb) ? __ldg(f.omega + i) : 0.0f;
                const bool kb_blue = kb * XKB < Nb;
                const int iz = i < Nb ? i : (Nb > 0 ? Nb - 1 : 0);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int row = 2 * jj + half;
                    if (row < nr && inr) {
                        const size_t o = (row0 + row) * (size_t)P + i;
                        k.r[jj].m = ldg_stream_u8(f.mask + o);
                        k.r[jj].x = ldg_stream(f.x + o);
                        k.r[jj].e = ldg_stream(f.err + o);
                        k.r[jj].z = kb_blue ? ldg_stream(f.zabs + (row0 + row) * (size_t)Nb + iz) : 0.0f;
                    } else { k.r[jj].m = 0u; k.r[jj].x = 0.f; k.r[jj].e = 1.f; k.r[jj].z = 0.f; }
                }
            };
            XBuf kA, kB;
            load_kb(0, kA);
            if (nkb > 1) load_kb(1, kB);
            // (row = 8*warp, column px, chunk not yet swizzled) of operand tile 0 of stage 0
            const uint32_t base_sa = smem_u32(ring) + (uint32_t)warp * 1024u + (uint32_t)px * 4u;
            auto do_kblock = [&](int kb, XBuf& k) {
                const uint32_t git = git0 + kb;
                const int s = git % XNSTAGE;
                const uint32_t u = git / XNSTAGE;
                const bool blue = kb * XKB + px < Nb;
                const bool kb_blue = kb * XKB < Nb;
                if (u > 0) mbar_wait_or_trap(&bar_empty[s], (u - 1) & 1);
                const uint32_t stage_sa = base_sa + (uint32_t)s * (uint32_t)C::STAGE_BYTES;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int row = 2 * jj + half;
                    const XRow& rb = k.r[jj];
                    const bool mk = rb.m != 0u;
                    float A = 1.0f, oz = 0.0f;
                    if (kb_blue) {
                        // full-accuracy transcendentals in this mode (the single-pass kernels use the .approx MUFU forms)
                        const float L2 = log2f(1.0f + rb.z);
                        const float tau = fmaf(pc.lt0, exp2f(pc.lbe * (L2 - pc.l2zn)), pc.lC);    // utils.py:106 etc.
                        const float Ab = exp2f(-kLog2e * tau);                                      // model.py:125
                        const float powb = exp2f(pc.beta * L2);                                     // utils.py:72
                        const float root = pc.one_m_c0 - exp2f(pc.nt0l2e * powb);                   // utils.py:91
                        A = blue ? Ab : 1.0f;
                        oz = k.om * (root * root);
                    }
                    const float A2 = A * A;
                    const float D = fmaf(A2, k.psi, fmaf(rb.e, rb.e, oz));                          // model.py:128-131
                    const float w = 1.0f / D;
                    const float r = TRAIN ? rb.x : fmaf(-k.mu, A, rb.x);                            // model.py:166
                    const float wA = w * A;
                    const float s2 = wA * A, wb = wA * r;
                    const float et = fmaf(w * r, r, logf(D) + kLn2Pi);
                    // address of (row, column px): swizzle phase = row & 7; the lo half sits 16 columns further = chunk ^ 4
                    const uint32_t sa = (stage_sa ^ ((uint32_t)row << 4)) + (uint32_t)row * 128u;
                    float hi, lo;
                    split_tf32(mk ? s2 : 0.f, hi, lo);  sts_f32(sa, hi);               sts_f32(sa ^ 64u, lo);
                    split_tf32(mk ? wb : 0.f, hi, lo);  sts_f32(sa + A_TILE, hi);      sts_f32((sa ^ 64u) + A_TILE, lo);
                    if (TRAIN && kb_blue) {
                        split_tf32(mk ? s2 * A : 0.f, hi, lo);  sts_f32(sa + 2 * A_TILE, hi);  sts_f32((sa ^ 64u) + 2 * A_TILE, lo);
                        split_tf32(mk ? s2 * r : 0.f, hi, lo);  sts_f32(sa + 3 * A_TILE, hi);  sts_f32((sa ^ 64u) + 3 * A_TILE, lo);
                        nbm |= (mk && blue) ? (1u << jj) : 0u;
                    }
                    E[jj] += mk ? et : 0.0f;
                }
                fence_proxy_async_writer();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_full[s]);
                if (kb + 2 < nkb) load_kb(kb + 2, k);
            };
            for (int kb = 0; kb < nkb; kb += 2) {
                do_kblock(kb, kA);
                if (kb + 1 < nkb) do_kblock(kb + 1, kB);
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float e = E[jj];
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);     // within each 16-lane half
                const uint32_t bal = __ballot_sync(0xffffffffu, (nbm >> jj) & 1u);
                if (px == 0) {
                    sE[RPW * warp + 2 * jj + half] = e;
                    sNb[RPW * warp + 2 * jj + half] = (bal & (half ? 0xFFFF0000u : 0x0000FFFFu)) ? 1.0f : 0.0f;
                }
            }
            named_bar_sync(1, NTHREADS);
        }

        // =========================================================== phase S (warps 0..3: lane = spectrum row)
        if (warp < 4) {
            mbar_wait_or_trap(&bar_gram, tile_iter & 1);
            fence_after_sync();
            if (want_o && tid == 0) {      // ring is free: static operand of the first phase-O pixel tile
                mbar_expect_tx(&bar_qa, XQA_TILE);
                bulk_g2s(ring + XB2_TILE, g.QA, XQA_TILE, &bar_qa);
            }
            const int row = warp * 32 + lane;
            const int b = b0 + row;
            const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
            // sum of the accumulator sets (main + LO), in double
            const int nset = nkb < C::NSET ? nkb : C::NSET;
            auto read_gram = [&](int colM, int colB, double (&Gd)[NP2], double (&bd)[HP]) {
#pragma unroll
                for (int q = 0; q < NP2; ++q) Gd[q] = 0.0;
#pragma unroll
                for (int q = 0; q < HP; ++q) bd[q] = 0.0;
                for (int st = 0; st < 2 * nset; ++st) {
                    const uint32_t t = ta + (uint32_t)(st >> 1) * C::SETW + (uint32_t)(st & 1) * C::HALF;
                    float v[16];
                    tmem_ld16(t + colM + 0, v);  tmem_wait_ld();
#pragma unroll
                    for (int q = 0; q < 16; ++q) Gd[q] += (double)v[q];
                    tmem_ld16(t + colM + 16, v); tmem_wait_ld();
#pragma unroll
                    for (int q = 0; q < 16; ++q) Gd[16 + q] += (double)v[q];
                    tmem_ld16(t + colM + 32, v); tmem_wait_ld();
#pragma unroll
                    for (int q = 0; q < 4; ++q) Gd[32 + q] += (double)v[q];
                    float w8[8];
                    tmem_ld8(t + colB + 8, w8); tmem_wait_ld();
#pragma unroll
                    for (int q = 0; q < HP; ++q) bd[q] += (double)w8[q];
                }
            };
            double G[NP2], bv[HP];
            read_gram(C::D1A, C::D1B, G, bv);
            const bool row_ok = row < nrows;
            if (!row_ok) {
#pragma unroll
                for (int q = 0; q < NP2; ++q) G[q] = 0.0;
#pragma unroll
                for (int q = 0; q < HP; ++q) bv[q] = 0.0;
            }
            SolvedD sv;
            solve_spd8_d(G, bv, sv);
            const float nll = (float)(0.5 * ((double)sE[row] - sv.quad + sv.logdet));               // model.py:135
            if (row_ok) {
                if (g.nll) g.nll[b] = nll;
                if (g.hmean) {
#pragma unroll
                    for (int k = 0; k < HP; ++k) if (k < Nh) g.hmean[(size_t)b * Nh + k] = (float)sv.a[k];
                }
                if (g.hcov) {
#pragma unroll
                    for (int k = 0; k < HP; ++k)
#pragma unroll
                        for (int l = 0; l < HP; ++l)
                            if (k < Nh && l < Nh) g.hcov[((size_t)b * Nh + k) * Nh + l] = (float)sv.Minv[k <= l ? tri(k, l) : tri(l, k)];
                }
            }
            if (want_o) {
                // per-spectrum operand of phase O, entries e = 0..47 (8 a | 36 Minv, off-diagonals doubled | 0): block e / 16,
                // I1 = [ hi | hi ], I2 = [ lo | 0 ]
                const uint32_t b2 = smem_u32(ring);
#pragma unroll
                for (int c4 = 0; c4 < 12; ++c4) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int e = c4 * 4 + q;
                        double v = 0.0;
                        if (e < HP) v = sv.a[e];
                        else if (e < HP + NP2) {
                            int k = 0, n = e - HP;
                            while (n >= HP - k) { n -= HP - k; ++k; }
                            v = (n == 0 ? 1.0 : 2.0) * sv.Minv[e - HP];
                        }
                        split_tf32((float)v, hi[q], lo[q]);
                    }
                    const int blk = c4 >> 2, col = (c4 & 3) * 4;
                    const uint32_t t1 = b2 + (uint32_t)blk * A_TILE, t2 = b2 + XB2_IMG + (uint32_t)blk * A_TILE;
                    sts_v4(t1 + sw128_offset(row, col), hi[0], hi[1], hi[2], hi[3]);
                    sts_v4(t1 + sw128_offset(row, 16 + col), hi[0], hi[1], hi[2], hi[3]);
                    sts_v4(t2 + sw128_offset(row, col), lo[0], lo[1], lo[2], lo[3]);
                    sts_v4(t2 + sw128_offset(row, 16 + col), 0.f, 0.f, 0.f, 0.f);
                }
                fence_proxy_async();
            }
            if (TRAIN) {
                // second Gram (quirk Q2) -> K = M^-1 M2, c = b2 - M2 a ; hand-off record of k_grad<float, 8>
                double G2[NP2], b2v[HP];
                read_gram(C::D1C, C::D1D, G2, b2v);
                if (row_ok) {
                    using SL = SmallLayout<HP>;
                    float* dst = g.small + (size_t)b * SL::len;
                    if (g.hasblue) g.hasblue[b] = sNb[row] > 0.f ? 1.0f : 0.0f;
#pragma unroll
                    for (int k = 0; k < HP; ++k) dst[SL::a + k] = (float)sv.a[k];
#pragma unroll
                    for (int k = 0; k < HP; ++k) {
                        double acc = b2v[k];
#pragma unroll
                        for (int m = 0; m < HP; ++m) acc -= G2[m <= k ? tri(m, k) : tri(k, m)] * sv.a[m];
                        dst[SL::c + k] = (float)acc;
                    }
#pragma unroll
                    for (int r = 0; r < HP; ++r)
#pragma unroll
                        for (int c = 0; c < HP; ++c) dst[SL::Linv + r * HP + c] = (float)sv.Li[r][c];
#pragma unroll
                    for (int l = 0; l < HP; ++l)
#pragma unroll
                        for (int k = 0; k < HP; ++k) {
                            double acc = 0.0;
#pragma unroll
                            for (int m = 0; m < HP; ++m)
                                acc += sv.Minv[l <= m ? tri(l, m) : tri(m, l)] * G2[m <= k ? tri(m, k) : tri(k, m)];
                            dst[SL::K + l * HP + k] = (float)acc;
                        }
                }
            }
            fence_before_sync();
        }

        // =========================================================== phase O
        if (want_o) {
            const uint32_t nsteps = (uint32_t)npt * 4u;
            const uint32_t gd0 = tile_iter * nsteps;
            const uint32_t oi0 = tile_iter * (uint32_t)npt;
            const uint32_t id32 = idesc_tf32(128, 32);
            const uint32_t b2a = smem_u32(ring);
            const uint32_t qa = smem_u32(ring + XB2_TILE);
            auto issue_step = [&](uint32_t d) {               // ONE thread (control warp)
                const uint32_t pt = d >> 2, h4 = d & 3;
                if (h4 == 0) {
                    if (pt > 0) {
                        // the single static-operand slot is free once the MMAs of step d-1 (last step of pixel tile pt-1) retired
                        mbar_wait_or_trap(&bar_tm_full[(gd0 + d - 1) & 1], ((gd0 + d - 1) >> 1) & 1);
                        mbar_expect_tx(&bar_qa, XQA_TILE);
                        bulk_g2s(ring + XB2_TILE, g.QA + (size_t)pt * (XQA_TILE / 4), XQA_TILE, &bar_qa);
                    }
                    mbar_wait_or_trap(&bar_qa, (oi0 + pt) & 1);
                }
                fence_after_sync();
                const uint32_t dcol = tmem + C::D2 + ((gd0 + d) & 1) * 64;
                bool fa_acc = false, q_acc = false;
#pragma unroll
                for (int img = 0; img < 2; ++img) {
#pragma unroll
                    for (int blk = 0; blk < XQA_BLK; ++blk) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            if (img == 1 && kk >= 2) continue;
                            const uint64_t dA = desc_sw128_kmajor(qa + blk * (PT * 128)) + (uint64_t)(2 * kk);
                            const uint64_t dB = desc_sw128_kmajor(b2a + img * XB2_IMG + blk * A_TILE + h4 * 32 * 128) + (uint64_t)(2 * kk);
                            const bool is_fa = blk == 0 && (kk == 0 || kk == 2);     // entries 0..7 (hi: k-step 0, lo: k-step 2)
                            if (is_fa) { umma_tf32(dcol, dA, dB, id32, fa_acc); fa_acc = true; }
                            else { umma_tf32(dcol + 32, dA, dB, id32, q_acc); q_acc = true; }
                        }
                    }
                }
                umma_commit(&bar_tm_full[(gd0 + d) & 1]);
            };
            named_bar_sync(1, NTHREADS);                      // B2 operand written (phase S), sE consumed
            if (warp == NWW) { if (elect_one()) { issue_step(0); if (nsteps > 1) issue_step(1); } __syncwarp(); }
            const int quad = warp & 3, grp = warp >> 2;
            const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16) + C::D2 + grp * 8;
            const uint32_t so_sa = smem_u32(ring + XB2_TILE + XQA_TILE);           // [2][2 outputs][32 spectra][128 pixels]
            const int pi = quad * 32 + lane;
            const uint32_t so_w = so_sa + (uint32_t)(grp * 8 * PT + pi) * 4u;
            const uint32_t so_r = so_sa + (uint32_t)(warp * 4 * PT + lane) * 4u;
            for (int pt = 0; pt < npt; ++pt) {
                const int i = pt * PT + pi;
                const float mu = i < P ? __ldg(f.mu + i) : 0.f;
                const int npx = P - pt * PT < PT ? P - pt * PT : PT;
#pragma unroll 1
                for (int h4 = 0; h4 < 4; ++h4) {
                    const uint32_t d = (uint32_t)pt * 4u + h4;
                    const uint32_t gd = gd0 + d;
                    const int buf = gd & 1;
                    mbar_wait_or_trap(&bar_tm_full[buf], (gd >> 1) & 1);
                    fence_after_sync();
                    float fa[8], qq[8];
                    tmem_ld8(ta + buf * 64, fa);
                    tmem_ld8(ta + buf * 64 + 32, qq);
                    tmem_wait_ld();
                    fence_before_sync();
                    const uint32_t sob = (uint32_t)buf * (uint32_t)SO_BYTES;
                    const uint32_t so_ws = so_w + sob, so_rs = so_r + sob;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        sts_f32(so_ws + (uint32_t)(j * PT * 4), mu + fa[j]);                               // model.py:180
                        sts_f32(so_ws + (uint32_t)((32 + j) * PT * 4), sqrtf(fmaxf(qq[j], 0.f)));
                    }
                    named_bar_sync(2, NTHREADS);      // D2[buf] drained by every warp and this step's rows are staged
                    if (warp == NWW) { if (elect_one()) { if (d + 2 < nsteps) issue_step(d + 2); } __syncwarp(); }
                    {
                        float* dst0 = (warp < 8 ? g.cont : g.unc);
                        const int sp0 = (warp & 7) * 4;
                        const int rfirst = h4 * 32 + sp0;
                        const int bfirst = b0 + rfirst;
                        if (dst0 && rfirst < nrows) {
                            dst0 += (size_t)bfirst * P + (size_t)pt * PT + lane;
                            for (int t = 0; t < 4; ++t) {
                                if (rfirst + t < nrows) {
#pragma unroll
                                    for (int c4 = 0; c4 < 4; ++c4) {
                                        const int ii = c4 * 32 + lane;
                                        if (ii < npx) {
                                            float v;
                                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(so_rs + (uint32_t)(t * PT + c4 * 32) * 4u));
                                            st_stream(dst0 + (size_t)t * P + c4 * 32, v);
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
    }
    if (warp == 0) tmem_dealloc<C::TMEM_COLS>(tmem);
}

}  // namespace tcx
}  // namespace qfa
