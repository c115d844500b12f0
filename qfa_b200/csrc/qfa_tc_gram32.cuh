// Tensor-core (tcgen05 / TMEM) masked weighted Grams for 8 < Nh <= 32 (train step), sm_100a.
//
//   k_tc_gram32   persistent, one CTA per SM, tile = up to 120 spectra x all pixels, same worker / control-warp
//                 organisation and the same generated-operand pipeline as k_tc_gram<TRAIN> (qfa_tc_gram.cuh):
//                    D[b,(k,l)] = sum_i s2[b,i] (F_ik F_il)   -> M - I      (model.py:126-132, utils.py:29-32)
//                    D[b,(k,l)] = sum_i s3[b,i] (F_ik F_il)   -> M2         (quirk Q2 of model.py:137)
//                    D[b,k]     = sum_i wb[b,i] F_ik , sum_i wb2[b,i] F_ik  -> b, b2
//                 With Nh = 32 the Khatri-Rao image has 528 columns per operand: 2 x 528 + 64 accumulator columns do not
//                 fit the 512 TMEM columns, so the columns are covered in THREE passes over the tile (224 + 224 + 80 columns
//                 for each of M and M2; b / b2 ride in pass 0).  The operands are re-generated in every pass (the tile's
//                 inputs are re-read, mostly from L2); after each pass the accumulators are drained to a per-spectrum
//                 scratch row in global memory.
//   k_solve32     one warp per spectrum: scratch row -> Cholesky / inverse / K = M^-1 M2 / c in shared memory (double),
//                 NLL, and the hand-off record that the pixel-major gradient kernel k_grad<float,32> reads.
#pragma once
#include "qfa_tc_gram.cuh"
#include "qfa_kernels.cuh"

namespace qfa {
namespace tcg32 {

using namespace tc;
using namespace tcg;

constexpr int H32 = 32;
constexpr int NP2_32 = H32 * (H32 + 1) / 2;        // 528 Khatri-Rao columns (k <= l)
constexpr int PB32_ROWS = NP2_32 + H32;            // 560 image rows per K-block: 528 products, then F
constexpr int PB32_KB_BYTES = PB32_ROWS * 128;     // 71 680 B per K-block
constexpr int NPASS = 3;
constexpr int SLICE = 224;                         // Khatri-Rao columns per pass (last pass: 80)
constexpr int NSTAGE32 = 2;
constexpr int NPB32 = 3;
constexpr int STAGE32_BYTES = 4 * A_TILE;          // s2 | wb | s3 | wb2
constexpr int PB32_SLOT = (SLICE + H32) * 128;     // 32 KB: a slice of the product rows + the F rows (pass 0)
constexpr int PB32_OFF = NSTAGE32 * STAGE32_BYTES;
constexpr int SMEM32_BYTES = PB32_OFF + NPB32 * PB32_SLOT + 1024;
constexpr int PB32_LEAD = NPB32 - NSTAGE32;
// passes 1, 2 REPLAY the s2 / s3 operand tiles that pass 0 generated (stored pre-swizzled to a per-CTA global buffer, mostly
// L2-resident) with bulk copies: no worker math at all.  The pass-0 ring is re-cut into RP_NST stages of [s2 | s3].
constexpr int RP_NST = 4;
constexpr int RP_LEAD = 3;
constexpr int RP_STAGE = 2 * A_TILE;
constexpr int RP_TILE = TSH * 128;                  // 120 live rows of a tile (15 360 B)
constexpr int RP_KB_FLOATS = 2 * A_TILE / 4;        // replay buffer per K-block: [s2 tile | s3 tile]
static_assert(RP_NST * RP_STAGE <= NSTAGE32 * STAGE32_BYTES, "replay stages live in the pass-0 ring");
// TMEM columns (all 512): M slice | M2 slice | b | b2
constexpr int T_A = 0, T_C = SLICE, T_B = 2 * SLICE, T_D = 2 * SLICE + H32;
// scratch row per spectrum (floats): [M - I packed (528) | M2 packed (528) | b (32) | b2 (32) | E | n_blue>0 | pad]
constexpr int G32_M = 0, G32_M2 = NP2_32, G32_B = 2 * NP2_32, G32_B2 = 2 * NP2_32 + H32, G32_E = 2 * NP2_32 + 2 * H32;
constexpr int G32_STRIDE = 1128;                      // (+ 8: the four partial E / has-blue values of the cluster kernel, qfa_tc_gram32c.cuh)

// ---- k_tc_grad32 / k_solve32 image geometry
constexpr int G32_ROWS = 80;                          // image rows per spectrum (UMMA N)
constexpr int G32_IMG = G32_ROWS * 128;               // 10 240 B
#ifndef QFA_G32_LIVE_ROWS
#define QFA_G32_LIVE_ROWS (2 * H32 + 2)
#endif
#ifndef QFA_G32_CELL_LEAD
#define QFA_G32_CELL_LEAD 4
#endif
constexpr int G32_CELL_LEAD = QFA_G32_CELL_LEAD;      // even (TMEM buffer parity is compile-time per unrolled step)
constexpr int G32_LIVE = QFA_G32_LIVE_ROWS * 128;         // rows 0..65 carry data (8 448 B): only these are written and copied
constexpr int G32_SPS = 3;                            // spectra per step
constexpr int G32_STAGE = G32_SPS * G32_IMG;          // 30 720 B
constexpr int G32_NST = 5;                            // image ring stages (see k_tc_grad32: copy of step n+2 overwrites step n-3)
constexpr int G32_W = 12;                             // worker warps
constexpr int G32_THREADS = (G32_W + 1) * 32;            // k_out32, and k_tc_grad32 without the chained MMA
constexpr int G32_THREADS_CHAIN = (G32_W + 2) * 32;      // chained form: a second control warp issues the TMEM-A MMAs
constexpr int G32_A_OFF = 0;                          // 16 KB: F rows of the pixel tile
constexpr int G32_B_OFF = 16384;
constexpr int G32_RED_OFF = G32_B_OFF;                // epilogue [3 groups][128 pixels][40] floats: over the (then idle) ring
// EXPERIMENT, measured and switched off (-DQFA_G32_CHAIN=1 builds it; parity tests pass with it): k_solve32 stops at
// W = L^-1 M2 and k_tc_grad32 finishes f^T K = (L^-1 f)^T W with a SECOND, chained MMA whose A operand is the first MMA's
// result read straight from tensor memory (tcgen05.mma with A in TMEM), so that the 32 x 32 x 32 triangular product
// K = L^-T W leaves the CUDA cores.  Result on one B200, 65 536 spectra: k_solve32 575 -> 495 us, but k_tc_grad32 865 ->
// 1 083 us (1 113 with the chained MMAs on a control warp of their own): the twelve N = 32, K = 8 TMEM-A MMAs per step cost
// ~100 cycles each in the tensor pipe, which makes the kernel tensor-bound at ~1 750 cycles per step.  Net 2.16 -> 2.30 ms.
#ifndef QFA_G32_CHAIN
#define QFA_G32_CHAIN 0
#endif
#ifndef QFA_G32_CPASYNC
#define QFA_G32_CPASYNC 0
#endif
// EXPERIMENT, measured and switched off (-DQFA_G32_CPASYNC=1 builds it): k_tc_grad32 cell prefetch through shared memory
// (cp.async, 4 bytes per array and thread, G32_CL steps deep) instead of registers.  Motivation: ncu (round 2) puts 15 % of the
// kernel's stall samples on the branch at the head of the unrolled step loop, long scoreboard -- the 16 register loads in flight
// share six hardware scoreboards, so waiting for the oldest cell also waits for the newest.  Result on one B200, 65 536 spectra:
// 865 -> 938 us (the LDGSTS / LDS / wait_group instructions cost more than the stall they remove); issuing the loads of a whole
// group of four steps at once into a second register set was worse still (1 228 us: 24 bytes of spills).
constexpr int G32_CL = 8;
constexpr int G32_CELL_OFF = G32_B_OFF + G32_NST * G32_STAGE;
constexpr int G32_CELL_BYTES = QFA_G32_CPASYNC ? G32_CL * 4 * (G32_W * 32) * 4 : 0;
constexpr int G32_SMEM = G32_B_OFF + G32_NST * G32_STAGE + G32_CELL_BYTES + 1024;
static_assert(3 * 128 * 40 * 4 <= G32_NST * G32_STAGE, "epilogue staging fits the ring");
constexpr int G32_TBUF = 256;                         // TMEM columns per buffer (3 x 80 used)


__host__ __device__ constexpr int tri32(int k, int l) { return k * H32 - k * (k - 1) / 2 + (l - k); }   // k <= l
__host__ __device__ constexpr int slice_cols(int p) { return p < 2 ? SLICE : NP2_32 - 2 * SLICE; }

// static image: PB32[kb] = 560 rows x 32 pixels (SWIZZLE_128B rows, TF32-rounded); row n < 528 -> F_ik F_il, n = tri32(k,l)
// part = 0: TF32 (rna) image; part = 1: the residual image tf32(v - tf32(v)) for the 3xTF32 prediction Grams
__global__ void k_tc_build_images32(const float* __restrict__ F, int P, int Nh, float* __restrict__ PB, int nkb, int part = 0) {
    const size_t n_el = (size_t)nkb * PB32_ROWS * KB;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (size_t)gridDim.x * blockDim.x) {
        const int kappa = (int)(e % KB), row = (int)((e / KB) % PB32_ROWS), kb = (int)(e / (KB * PB32_ROWS));
        const int i = kb * KB + kappa;
        float v = 0.f;
        if (i < P) {
            if (row < NP2_32) {
                int k = 0, n = row;
                while (n >= H32 - k) { n -= H32 - k; ++k; }
                const int l = k + n;
                if (k < Nh && l < Nh) v = __ldg(F + (size_t)i * Nh + k) * __ldg(F + (size_t)i * Nh + l);
            } else if (row - NP2_32 < Nh) {
                v = __ldg(F + (size_t)i * Nh + (row - NP2_32));
            }
        }
        const float hi = tf32_rna(v);
        PB[(size_t)kb * (PB32_KB_BYTES / 4) + sw128_offset(row, kappa) / 4] = part == 0 ? hi : tf32_rna(v - hi);
    }
    if (blockIdx.x == 0 && threadIdx.x < 4) PB[(size_t)nkb * (PB32_KB_BYTES / 4) + threadIdx.x] = 0.f;   // 16 zero bytes (dummy mask)
}

// per-cell physics + operand generation; PASS0 also produces the wb / wb2 operands and the scalar sums
// PRED: prediction flavour (model.py:160-180): residual flux - mu A, only the s2 / wb operands (no second Gram)
// PART (PRED only): 0 = the TF32 part of the generated operands, 1 = their residual tf32(v - tf32(v)) (3xTF32 prediction Grams)
__device__ __forceinline__ float tf32_part(float v, int part) {
    const float hi = tf32_rna(v);
    return part == 0 ? hi : tf32_rna(v - hi);
}
template <int BLUE, bool PASS0, bool PRED = false, int PART = 0>
__device__ __forceinline__ void compute_row32(const PhysConst& pc, const PixConst& px, const RowBuf& rb, bool blue, uint32_t sa,
                                              float& E, uint32_t& nbm, int j) {
    const bool mk = rb.m != 0u;
    float A = 1.0f, oz = 0.0f;
    if (BLUE != 0) {
        const float L2 = lg2f(1.0f + rb.z);
        const float tau = fmaf(pc.lt0, ex2f(pc.lbe * (L2 - pc.l2zn)), pc.lC);     // utils.py:106 etc.
        const float Ab = ex2f(-kLog2e * tau);                                       // model.py:125
        const float powb = ex2f(pc.beta * L2);                                      // utils.py:72
        const float root = pc.one_m_c0 - ex2f(pc.nt0l2e * powb);                    // utils.py:91
        A = (BLUE == 1 || blue) ? Ab : 1.0f;
        oz = px.om * (root * root);
    }
    const float A2 = A * A;
    const float D = fmaf(A2, px.psi, fmaf(rb.e, rb.e, oz));                         // model.py:128-131
    const float w = rcpf(D);
    const float wA = w * A;
    const float s2 = wA * A;
    sts_f32_imm<0>(sa, mk ? (PRED ? tf32_part(s2, PART) : tf32_round(s2)) : 0.0f);
    if (BLUE != 0 && !PRED) sts_f32_imm<2 * A_TILE>(sa, mk ? tf32_round(s2 * A) : 0.0f);     // s3 (= s2 on red K-blocks)
    if (PASS0) {
        const float r = PRED ? fmaf(-px.mu, A, rb.x) : rb.x;                        // model.py:166
        const float wb = wA * r;
        sts_f32_imm<A_TILE>(sa, mk ? (PRED ? tf32_part(wb, PART) : tf32_round(wb)) : 0.0f);
        if (BLUE != 0 && !PRED) {
            sts_f32_imm<3 * A_TILE>(sa, mk ? tf32_round(s2 * r) : 0.0f);            // wb2 (= wb on red K-blocks)
            nbm |= (mk && (BLUE == 1 || blue)) ? (1u << j) : 0u;
        }
        const float et = fmaf(w * r, r, fmaf(kLn2, lg2f(D), kLn2Pi));               // w r^2 + ln(2 pi D)
        E += mk ? et : 0.0f;
    }
}

template <int BLUE, bool PASS0, int NR, bool PRED = false, int PART = 0>
__device__ __forceinline__ void kblock_consume32(const PhysConst& pc, const KBuf& kb, bool blue, uint32_t warp_sa, int nr,
                                                 float (&E)[RPW], uint32_t& nbm) {
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        if (NR > 0 ? j < NR : j < nr)
            compute_row32<BLUE, PASS0, PRED, PART>(pc, kb.px, kb.r[j], blue, (warp_sa ^ ((uint32_t)j << 4)) + (uint32_t)j * 128u, E[j], nbm, j);
    }
}

// Pass 0 of one tile for one worker warp: generate the operand tiles K-block by K-block.  A separate NON-INLINED function:
// inlined into the kernel, the tile / pass loop state around it pushed the two prefetch buffers (64 registers) into local
// memory, and a load whose result is spilled must be waited for on the spot -- every K-block then paid the full memory
// latency (period 6 200 cycles, 3 600 of them in the load phase).
struct G32Worker {
    const float* x; const float* err; const float* zabs; const uint8_t* mask;   // row 0 of the tile
    const float* Psi; const float* omega; const uint8_t* zero;
    int P, Nb, nkb, nr, warp, lane;
    uint32_t git0, ring_sa;
    uint64_t* bar_full; uint64_t* bar_empty;
    float* sE; float* sNb;
    long long* trace;
};

template <bool PRED, int PART>
__device__ __noinline__ void gram32_pass0_worker(const G32Worker w, const PhysConst pc) {
    const int P = w.P, Nb = w.Nb, nkb = w.nkb, nr = w.nr, warp = w.warp, lane = w.lane;
    Field<float> f;
    // prediction flavour: the mean spectrum travels in the (debug-only, otherwise null) trace slot -- one more by-value member
    // of this struct changed the register allocation of the TRAIN instance (+10 % kernel time, measured)
    f.P = P; f.Nb = Nb; f.Psi = w.Psi; f.omega = w.omega; f.mu = PRED ? reinterpret_cast<const float*>(w.trace) : nullptr;
    float E[RPW];
    uint32_t nbm = 0u;
#pragma unroll
    for (int j = 0; j < RPW; ++j) E[j] = 0.f;
    TileView tv;
    tv.zero = w.zero;
    RowCursor rc;
    rc.x = reinterpret_cast<const unsigned char*>(w.x + (size_t)(RPW * warp) * P + lane);
    rc.e = reinterpret_cast<const unsigned char*>(w.err + (size_t)(RPW * warp) * P + lane);
    rc.m = w.mask + (size_t)(RPW * warp) * P + lane;
    rc.z = reinterpret_cast<const unsigned char*>(w.zabs + (size_t)(RPW * warp) * Nb + lane);
    KBuf kA, kB;
#pragma unroll
    for (int j = 0; j < RPW; ++j) { kA.r[j].z = 0.f; kB.r[j].z = 0.f; }
    const uint32_t ring_sa = w.ring_sa, git0 = w.git0;
    uint64_t* const bar_full = w.bar_full;
    uint64_t* const bar_empty = w.bar_empty;
    auto run_kblocks = [&](auto nr_tag) {
        constexpr int NR = decltype(nr_tag)::value;
        constexpr bool PASS0 = true;
        auto load_kb = [&](int kb, KBuf& k) {
            const int p0 = kb * KB;
            if (p0 + KB <= P) load_kblock<!PRED, true, NR, true, true>(f, tv, rc, p0, lane, nr, p0 < Nb, k);
            else load_kblock<!PRED, false, NR, true, true>(f, tv, rc, p0, lane, nr, p0 < Nb, k);
        };
        load_kb(0, kA);
        if (nkb > 1) load_kb(1, kB);
        auto do_kblock = [&](int kb, KBuf& k) {
            const uint32_t git = git0 + kb;
            const int s = git % NSTAGE32;
            const uint32_t u = git / NSTAGE32;
            const int p0 = kb * KB;
            const bool blue = p0 + lane < Nb;
            long long* tr = (kTrace && !PRED && w.trace) ? w.trace + (size_t)kb * NPW * 4 : nullptr;
            if (tr) tr[0] = clock64();
            if (u > 0) mbar_wait_or_trap(&bar_empty[s], (u - 1) & 1);
            if (tr) tr[1] = clock64();
            const uint32_t stage_sa = ring_sa + (uint32_t)s * (uint32_t)STAGE32_BYTES;
            const int bm = (p0 + KB <= Nb) ? 1 : (p0 >= Nb ? 0 : 2);
            if (bm == 1) kblock_consume32<1, PASS0, NR, PRED, PART>(pc, k, blue, stage_sa, nr, E, nbm);
            else if (bm == 0) kblock_consume32<0, PASS0, NR, PRED, PART>(pc, k, blue, stage_sa, nr, E, nbm);
            else kblock_consume32<2, PASS0, NR, PRED, PART>(pc, k, blue, stage_sa, nr, E, nbm);
            if (tr) tr[2] = clock64();
            fence_proxy_async_writer();     // MEMBAR.ALL.CTA: before the prefetch below, never after it
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[s]);
            if (kb + 2 < nkb) load_kb(kb + 2, k);
            if (tr) tr[3] = clock64();
        };
        // (tried: a loop of their own for the all-red K-blocks, where the 16 `z` registers are dead -- still 10 spilled
        //  prefetch registers per K-block there, and any spilled load result makes the load phase wait: no gain)
        for (int kb = 0; kb < nkb; kb += 2) {
            do_kblock(kb, kA);
            if (kb + 1 < nkb) do_kblock(kb + 1, kB);
        }
    };
    if (nr == RPW) run_kblocks(std::integral_constant<int, RPW>{});
    else run_kblocks(std::integral_constant<int, 0>{});
#pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const float e = warp_sum(E[j]);
        const bool any = __any_sync(0xffffffffu, (nbm >> j) & 1u);
        if (lane == 0) { w.sE[j] = e; w.sNb[j] = any ? 1.0f : 0.0f; }
    }
}

struct TcGram32Args {
    Field<float> f;
    int B;
    TileSched ts;
    int ntiles, nkb;
    const float* PB;       // [nkb][PB32_KB_BYTES/4]
    float* gram;           // [B][G32_STRIDE]
    float* replay;         // [gridDim.x][nkb][RP_KB_FLOATS]  s2 / s3 operand tiles of the CTA's current tile
    long long* trace;      // debug (-DQFA_ENABLE_TRACE): CTA 0, first tile, pass 0: [kb][warp][4] clock64 stamps
};

// PRED = true: prediction flavour -- M and b only (the s3 / wb2 operands, their MMAs, replay copies and drains are skipped)
template <bool PRED, int PART = 0>
__global__ void __launch_bounds__(NTHREADS, 1) k_tc_gram32(const TcGram32Args g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* ring = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full[NSTAGE32], bar_empty[NSTAGE32], bar_pb[NPB32], bar_gram, bar_rp_full[RP_NST], bar_rp_empty[RP_NST];
    __shared__ uint32_t tmem_base_s;
    __shared__ float sE[TS];
    __shared__ float sNb[TS];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const Field<float>& f = g.f;
    const int P = f.P, Nb = f.Nb;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE32; ++s) { mbar_init(&bar_empty[s], 2); mbar_init(&bar_full[s], NWW); }   // empty: MMAs retired + replay copy read
        for (int s = 0; s < NPB32; ++s) mbar_init(&bar_pb[s], 1);
        for (int s = 0; s < RP_NST; ++s) { mbar_init(&bar_rp_full[s], 1); mbar_init(&bar_rp_empty[s], 1); }
        mbar_init(&bar_gram, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
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
    const int nkb = g.nkb;

    uint32_t pass_iter = 0;          // counts (tile, pass) pairs of this CTA: image-ring / accumulator-barrier parities
    uint32_t tile_iter = 0;          // counts tiles of this CTA: the pass-0 operand ring; 2 replay passes per tile
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x, ++tile_iter) {
        const int b0 = g.ts.first(tile);
        const int nrows = g.B - b0 < g.ts.rows(tile) ? g.B - b0 : g.ts.rows(tile);
        const int nr = warp < NWW ? (nrows - RPW * warp < 0 ? 0 : (nrows - RPW * warp > RPW ? RPW : nrows - RPW * warp)) : 0;
        for (int pass = 0; pass < NPASS; ++pass, ++pass_iter) {
            const uint32_t git0 = tile_iter * (uint32_t)nkb;          // pass-0 operand ring
            const uint32_t pit0 = pass_iter * (uint32_t)nkb;          // image ring (every pass)
            const int ncol = slice_cols(pass);
            if (warp == NWW) {
                // ----------------------------------------------------------- CONTROL warp
                if (elect_one()) {
                    float* const rbase = g.replay + (size_t)blockIdx.x * (size_t)nkb * RP_KB_FLOATS;
                    auto issue_pb = [&](uint32_t pit, int kb) {
                        const int slot = pit % NPB32;
                        unsigned char* dst = ring + PB32_OFF + (size_t)slot * PB32_SLOT;
                        const float* src = g.PB + (size_t)kb * (PB32_KB_BYTES / 4);
                        const uint32_t bytes = (uint32_t)ncol * 128u + (pass == 0 ? (uint32_t)H32 * 128u : 0u);
                        mbar_expect_tx(&bar_pb[slot], bytes);
                        bulk_g2s(dst, src + (size_t)pass * SLICE * 32, (uint32_t)ncol * 128u, &bar_pb[slot]);
                        if (pass == 0) bulk_g2s(dst + SLICE * 128, src + (size_t)NP2_32 * 32, H32 * 128, &bar_pb[slot]);
                    };
                    const uint32_t idA = idesc_tf32(128, ncol), idB = idesc_tf32(128, H32);
                    if (pass == 0) {
                        for (int kb = 0; kb < PB32_LEAD && kb < nkb; ++kb) issue_pb(pit0 + kb, kb);
                        for (int kb = 0; kb < nkb; ++kb) {
                            const uint32_t git = git0 + kb, pit = pit0 + kb;
                            const int s = git % NSTAGE32, slot = pit % NPB32;
                            const uint32_t sb = smem_u32(ring) + (uint32_t)s * (uint32_t)STAGE32_BYTES;
                            const uint32_t pb = smem_u32(ring) + (uint32_t)PB32_OFF + (uint32_t)slot * (uint32_t)PB32_SLOT;
                            mbar_wait_or_trap(&bar_full[s], (git / NSTAGE32) & 1);
                            fence_proxy_async_issuer();
                            mbar_wait_or_trap(&bar_pb[slot], (pit / NPB32) & 1);
                            fence_after_sync();
                            const bool red = kb * KB >= Nb;        // all-red K-block: s3 = s2, wb2 = wb (two operand tiles only)
                            const uint64_t dS2 = desc_sw128_kmajor(sb), dWb = desc_sw128_kmajor(sb + A_TILE);
                            const uint64_t dS3 = red ? dS2 : desc_sw128_kmajor(sb + 2 * A_TILE);
                            const uint64_t dW2 = red ? dWb : desc_sw128_kmajor(sb + 3 * A_TILE);
                            const uint64_t dP = desc_sw128_kmajor(pb), dF = desc_sw128_kmajor(pb + SLICE * 128);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const uint64_t ko = (uint64_t)(2 * kk);
                                const bool acc = (kb | kk) != 0;
                                umma_tf32(tmem + T_A, dS2 + ko, dP + ko, idA, acc);
                                if (!PRED) umma_tf32(tmem + T_C, dS3 + ko, dP + ko, idA, acc);
                                umma_tf32(tmem + T_B, dWb + ko, dF + ko, idB, acc);
                                if (!PRED) umma_tf32(tmem + T_D, dW2 + ko, dF + ko, idB, acc);
                            }
                            umma_commit(&bar_empty[s]);
                            if (kb == nkb - 1) umma_commit(&bar_gram);
                            if (kb + PB32_LEAD < nkb) issue_pb(pit + PB32_LEAD, kb + PB32_LEAD);
                            // replay copy of the s2 / s3 tiles, shared -> global through the async proxy (the workers' stores
                            // are already fenced for it); the stage goes back to the workers when the MMAs have retired AND
                            // this copy has read it
                            float* rp = rbase + (size_t)kb * RP_KB_FLOATS;
                            bulk_s2g(rp, sb, RP_TILE);
                            if (!red && !PRED) bulk_s2g(rp + A_TILE / 4, sb + 2 * A_TILE, RP_TILE);
                            bulk_commit_group();
                            bulk_wait_group_read0();
                            mbar_arrive(&bar_empty[s]);
                        }
                        bulk_wait_group0();              // replay tiles are in global memory before pass 1 reads them back
                    } else {
                        // replay: [s2 | s3] tiles of K-block kb -> ring stage, RP_LEAD K-blocks ahead; images 2 ahead
                        const uint32_t rit0 = (tile_iter * 2u + (uint32_t)(pass - 1)) * (uint32_t)nkb;
                        auto issue_rp = [&](uint32_t rit, int kb) {
                            const int s = rit % RP_NST;
                            const uint32_t u = rit / RP_NST;
                            if (u > 0) mbar_wait_or_trap(&bar_rp_empty[s], (u - 1) & 1);     // the MMAs that read the stage retired
                            const bool red = PRED || kb * KB >= Nb;
                            unsigned char* dst = ring + (size_t)s * RP_STAGE;
                            const float* src = rbase + (size_t)kb * RP_KB_FLOATS;
                            mbar_expect_tx(&bar_rp_full[s], red ? RP_TILE : 2 * RP_TILE);
                            bulk_g2s(dst, src, RP_TILE, &bar_rp_full[s]);
                            if (!red) bulk_g2s(dst + A_TILE, src + A_TILE / 4, RP_TILE, &bar_rp_full[s]);
                        };
                        for (int kb = 0; kb < 2 && kb < nkb; ++kb) issue_pb(pit0 + kb, kb);
                        for (int kb = 0; kb < RP_LEAD && kb < nkb; ++kb) issue_rp(rit0 + kb, kb);
                        for (int kb = 0; kb < nkb; ++kb) {
                            const uint32_t rit = rit0 + kb, pit = pit0 + kb;
                            const int s = rit % RP_NST, slot = pit % NPB32;
                            const uint32_t sb = smem_u32(ring) + (uint32_t)s * (uint32_t)RP_STAGE;
                            const uint32_t pb = smem_u32(ring) + (uint32_t)PB32_OFF + (uint32_t)slot * (uint32_t)PB32_SLOT;
                            long long* tr = (kTrace && g.trace && blockIdx.x == 0 && tile == (int)blockIdx.x)
                                                ? g.trace + (size_t)nkb * NPW * 4 + ((size_t)(pass - 1) * nkb + kb) * 8 : nullptr;
                            if (tr) tr[0] = clock64();
                            mbar_wait_or_trap(&bar_rp_full[s], (rit / RP_NST) & 1);
                            if (tr) tr[1] = clock64();
                            mbar_wait_or_trap(&bar_pb[slot], (pit / NPB32) & 1);
                            fence_after_sync();
                            if (tr) tr[2] = clock64();
                            const bool red = kb * KB >= Nb;
                            const uint64_t dS2 = desc_sw128_kmajor(sb);
                            const uint64_t dS3 = red ? dS2 : desc_sw128_kmajor(sb + A_TILE);
                            const uint64_t dP = desc_sw128_kmajor(pb);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const uint64_t ko = (uint64_t)(2 * kk);
                                const bool acc = (kb | kk) != 0;
                                umma_tf32(tmem + T_A, dS2 + ko, dP + ko, idA, acc);
                                if (!PRED) umma_tf32(tmem + T_C, dS3 + ko, dP + ko, idA, acc);
                            }
                            umma_commit(&bar_rp_empty[s]);
                            if (kb == nkb - 1) umma_commit(&bar_gram);
                            if (tr) tr[3] = clock64();
                            // stage (kb + 3) % 4 = stage of K-block kb - 1: waiting for its MMAs also frees image slot (kb + 2) % 3
                            if (kb + RP_LEAD < nkb) issue_rp(rit + RP_LEAD, kb + RP_LEAD);
                            else if (kb >= 1) mbar_wait_or_trap(&bar_rp_empty[(rit - 1) % RP_NST], ((rit - 1) / RP_NST) & 1);
                            if (kb + 2 < nkb) issue_pb(pit + 2, kb + 2);
                            if (tr) tr[4] = clock64();
                        }
                    }
                }
                __syncwarp();
            } else if (pass == 0) {
                // ----------------------------------------------------------- WORKER warps (pass 0 only)
                G32Worker wa;
                wa.x = f.x + (size_t)b0 * P; wa.err = f.err + (size_t)b0 * P; wa.mask = f.mask + (size_t)b0 * P;
                wa.zabs = f.zabs + (size_t)b0 * Nb;
                wa.Psi = f.Psi; wa.omega = f.omega;
                wa.zero = reinterpret_cast<const uint8_t*>(g.PB + (size_t)nkb * (PB32_KB_BYTES / 4));   // 16 zero bytes after the image
                wa.P = P; wa.Nb = Nb; wa.nkb = nkb; wa.nr = nr; wa.warp = warp; wa.lane = lane;
                wa.git0 = git0;
                wa.ring_sa = smem_u32(ring) + (uint32_t)warp * 1024u + (uint32_t)lane * 4u;
                wa.bar_full = bar_full; wa.bar_empty = bar_empty;
                wa.sE = sE + RPW * warp; wa.sNb = sNb + RPW * warp;
                wa.trace = PRED ? reinterpret_cast<long long*>(const_cast<float*>(f.mu))
                                : ((kTrace && g.trace && blockIdx.x == 0 && tile == (int)blockIdx.x && lane == 0) ? g.trace + warp * 4 : nullptr);
                gram32_pass0_worker<PRED, PART>(wa, pc);
            }
            named_bar_sync(1, NTHREADS);            // sE / sNb visible; every worker is done with the ring
            // ----------------------------------------------------------- drain the accumulators (warps 0..3: lane = spectrum row)
            if (warp < 4) {
                mbar_wait_or_trap(&bar_gram, pass_iter & 1);
                fence_after_sync();
                const int row = warp * 32 + lane;
                const bool row_ok = row < nrows;
                float* dst = g.gram + (size_t)(b0 + (row_ok ? row : 0)) * G32_STRIDE;
                const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
                for (int c = 0; c < ncol; c += 16) {
                    float va[16], vc[16];
                    tmem_ld16(ta + T_A + c, va);
                    if (!PRED) tmem_ld16(ta + T_C + c, vc);
                    tmem_wait_ld();
                    if (row_ok) {
                        float4* da = reinterpret_cast<float4*>(dst + G32_M + pass * SLICE + c);
                        float4* dc = reinterpret_cast<float4*>(dst + G32_M2 + pass * SLICE + c);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            da[q] = make_float4(va[4 * q], va[4 * q + 1], va[4 * q + 2], va[4 * q + 3]);
                            if (!PRED) dc[q] = make_float4(vc[4 * q], vc[4 * q + 1], vc[4 * q + 2], vc[4 * q + 3]);
                        }
                    }
                }
                if (pass == 0) {
                    for (int c = 0; c < H32; c += 16) {
                        float vb[16], vd[16];
                        tmem_ld16(ta + T_B + c, vb);
                        if (!PRED) tmem_ld16(ta + T_D + c, vd);
                        tmem_wait_ld();
                        if (row_ok) {
                            float4* db = reinterpret_cast<float4*>(dst + G32_B + c);
                            float4* dd = reinterpret_cast<float4*>(dst + G32_B2 + c);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                db[q] = make_float4(vb[4 * q], vb[4 * q + 1], vb[4 * q + 2], vb[4 * q + 3]);
                                if (!PRED) dd[q] = make_float4(vd[4 * q], vd[4 * q + 1], vd[4 * q + 2], vd[4 * q + 3]);
                            }
                        }
                    }
                    if (row_ok) { dst[G32_E] = sE[row]; dst[G32_E + 1] = sNb[row]; }
                }
                fence_before_sync();
            }
            // pass boundary: the accumulators are drained and every MMA has retired before the next pass overwrites them
            fence_before_sync();
            __syncthreads();
            fence_after_sync();
        }
    }
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------
// k_solve32: one warp per spectrum.  In: scratch row of k_tc_gram32.  Out: NLL, has-blue flag and the hand-off record
// [a | c | L^-1 | K] (SmallLayout<32>, float) that k_grad<float,32> consumes.           (model.py:132-135, utils.py:29-54)
// ---------------------------------------------------------------------------------------
constexpr int SOLVE32_WARPS = 4;
constexpr int SOLVE32_LD = 36;                                                  // float rows of 36: 16-byte aligned, LDS.128
constexpr int SOLVE32_WARP_FLOATS = 2 * H32 * SOLVE32_LD + 4 * H32;             // L -> L^-T, L^-1, b, y, a, spare
constexpr int SOLVE32_SMEM = SOLVE32_WARPS * SOLVE32_WARP_FLOATS * 4;           // 9.5 KB per warp
#ifndef QFA_SOLVE32_CTAS
#define QFA_SOLVE32_CTAS 4
#endif
#ifndef QFA_SOLVE32_STAGE
#define QFA_SOLVE32_STAGE 1
#endif

__device__ __forceinline__ float solve32_rsqrt(float x) {
    const float y = rsqrtf(x);
    return y * fmaf(-0.5f * x * y, y, 1.5f);             // one Newton step: full float accuracy
}
__device__ __forceinline__ double solve32_rsqrt(double x) { return rsqrt(x); }

// Everything after the factorisation is float and uses the TRIANGULAR factors only: y = L^-1 b, a = L^-T y,
// b^T M^-1 b = |y|^2, K = L^-T (L^-1 M2).  Each step then loses eps x cond(L) = eps x sqrt(cond(M)), where forming M^-1
// first and multiplying (eps x cond(M), cond(M) up to ~1e6 for high signal-to-noise spectra) would not do in float.
// TC = scalar type of the Cholesky factorisation itself: float by default (identical parity figures on every test case,
// including the badly conditioned 96-pixel ones; 7 % faster step), double with QFA_FLAG_SOLVE_FP64.
// PRED = true (prediction, model.py:176-180): no second Gram; outputs NLL, hmean = a, hcov = M^-1 = L^-T L^-1 and the image
// rows [L^-1 | a] that k_out32 turns into the continuum and its 1-sigma.
template <typename TC, bool PRED = false>
__global__ void __launch_bounds__(SOLVE32_WARPS * 32, QFA_SOLVE32_CTAS)
k_solve32(const float* __restrict__ gram, int B, float* __restrict__ img, float* __restrict__ nll, float* __restrict__ hasblue,
          float* __restrict__ hmean = nullptr, float* __restrict__ hcov = nullptr, int Nh = H32, int nparts = 1,
          int diff = 0) {      // diff = 1: scratch rows of k_tc_gram32c (M2, b2 stored as differences; four partial E / has-blue values)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int LD = SOLVE32_LD;
    float* sA = reinterpret_cast<float*>(smem_raw) + (size_t)w * SOLVE32_WARP_FLOATS;   // rows of L (diagonal = 1/L_jj), then L^-T
    float* sLi = sA + H32 * LD;                     // L^-1, row-major
    float* sb = sLi + H32 * LD;
    float* sy = sb + H32;
    float* sa = sy + H32;
    for (int b = blockIdx.x * SOLVE32_WARPS + w; b < B; b += gridDim.x * SOLVE32_WARPS) {
        const float* src = gram + (size_t)b * G32_STRIDE;
        float* dst = img + (size_t)b * (G32_IMG / 4);
        // ---- row `lane` of M = I + Gram in registers (the scratch holds the packed upper triangle), then a right-looking
        //      Cholesky entirely in registers: pivots and columns travel by warp shuffles (no shared-memory round trips, no
        //      lane-serial inner loop); afterwards a[k] (k <= lane) = L[lane][k]
        float logdet;
        {
            // The scratch row holds the PACKED upper triangles: fetched with coalesced 16-byte loads into shared memory (the
            // L^-1 area, free until the substitution), then unpacked from there.  Reading element tri32(min(k,lane),
            // max(k,lane)) straight from global memory (one instruction per k) touched ~17 different 32-byte sectors per
            // instruction: 35 KB of L2 -> SM sector traffic for 4.5 KB of data, 64 serialised scattered loads per spectrum.
#if QFA_SOLVE32_STAGE
            {
                const float4* s4 = reinterpret_cast<const float4*>(src + G32_M);
                float4* d4 = reinterpret_cast<float4*>(sLi);
#pragma unroll
                for (int q = 0; q < (NP2_32 / 4 + 31) / 32; ++q)
                    if (q * 32 + lane < NP2_32 / 4) {
                        float4 v = __ldg(s4 + q * 32 + lane);
                        if (PRED && nparts == 3) {   // 3xTF32 prediction Grams: hi*hi + lo*hi + hi*lo from three launches (small parts first)
                            const float4 u1 = __ldg(s4 + (size_t)1 * B * (G32_STRIDE / 4) + q * 32 + lane);
                            const float4 u2 = __ldg(s4 + (size_t)2 * B * (G32_STRIDE / 4) + q * 32 + lane);
                            v.x += u1.x + u2.x; v.y += u1.y + u2.y; v.z += u1.z + u2.z; v.w += u1.w + u2.w;
                        }
                        d4[q * 32 + lane] = v;
                    }
                __syncwarp();
            }
#endif
            TC a[H32];
#pragma unroll
            for (int k = 0; k < H32; ++k) {
                const int lo = k < lane ? k : lane, hi = k < lane ? lane : k;
#if QFA_SOLVE32_STAGE
                a[k] = (TC)sLi[tri32(lo, hi)] + (k == lane ? TC(1) : TC(0));
#else
                a[k] = (TC)__ldg(src + G32_M + tri32(lo, hi)) + (k == lane ? TC(1) : TC(0));
#endif
            }
            {
                float bsum = 0.f;
                if (PRED && nparts == 3)
                    bsum = __ldg(src + (size_t)1 * B * G32_STRIDE + G32_B + lane) + __ldg(src + (size_t)2 * B * G32_STRIDE + G32_B + lane);
                sb[lane] = __ldg(src + G32_B + lane) + bsum;
            }
#if QFA_SOLVE32_STAGE
            __syncwarp();                                  // every lane has unpacked M: the staging area takes M2 now
            if (!PRED) {
                const float4* s4 = reinterpret_cast<const float4*>(src + G32_M2);
                const float4* g4 = reinterpret_cast<const float4*>(src + G32_M);
                float4* d4 = reinterpret_cast<float4*>(sLi);
#pragma unroll
                for (int q = 0; q < (NP2_32 / 4 + 31) / 32; ++q)
                    if (q * 32 + lane < NP2_32 / 4) {
                        float4 v = __ldg(s4 + q * 32 + lane);
                        if (diff) {                               // M2 = (M - I) - Md
                            const float4 m = __ldg(g4 + q * 32 + lane);
                            v = make_float4(m.x - v.x, m.y - v.y, m.z - v.z, m.w - v.w);
                        }
                        d4[q * 32 + lane] = v;
                    }
            }
#endif
            TC myinv = TC(1), mydiag = TC(1);
#pragma unroll
            for (int j = 0; j < H32; ++j) {
                const TC djj = __shfl_sync(0xffffffffu, a[j], j);           // pivot, already reduced by the previous steps
                const TC inv = solve32_rsqrt(djj);
                if (lane == j) { myinv = inv; mydiag = djj; }
                a[j] *= inv;                                                // column j of L (rows >= j)
#pragma unroll
                for (int k = j + 1; k < H32; ++k) {
                    const TC lkj = __shfl_sync(0xffffffffu, a[j], k);       // L[k][j]
                    a[k] = fma(-a[j], lkj, a[k]);                           // rows < k hold junk there (never read)
                }
            }
#pragma unroll
            for (int k4 = 0; k4 < H32; k4 += 4) {
                float4 v;
                v.x = (k4 == lane) ? (float)myinv : (float)a[k4];
                v.y = (k4 + 1 == lane) ? (float)myinv : (float)a[k4 + 1];
                v.z = (k4 + 2 == lane) ? (float)myinv : (float)a[k4 + 2];
                v.w = (k4 + 3 == lane) ? (float)myinv : (float)a[k4 + 3];
                *reinterpret_cast<float4*>(sA + lane * LD + k4) = v;
            }
            logdet = logf((float)mydiag);                                   // log det M = sum_j log(pivot_j): one log per lane
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) logdet += __shfl_xor_sync(0xffffffffu, logdet, o);
        }
        // column `lane` of M2 (symmetric): issued here so that the scattered loads fly under the substitution
        float m2c[H32];
#if QFA_SOLVE32_STAGE
        __syncwarp();                                      // M2 staged (its loads flew under the factorisation)
#endif
        float cv = 0.f;
        if (!PRED) {
#pragma unroll
            for (int k = 0; k < H32; ++k) {
                const int lo = k < lane ? k : lane, hi = k < lane ? lane : k;
#if QFA_SOLVE32_STAGE
                m2c[k] = sLi[tri32(lo, hi)];
#else
                m2c[k] = __ldg(src + G32_M2 + tri32(lo, hi));
#endif
            }
            cv = __ldg(src + G32_B2 + lane);
            if (diff) cv = __ldg(src + G32_B + lane) - cv;       // b2 = b - b2d
        }
        __syncwarp();
        // ---- column `lane` of L^-1 in registers (forward substitution, rows of L broadcast from shared memory)
        float x[H32];
#pragma unroll
        for (int r = 0; r < H32; ++r) {
            float sacc = (r == lane) ? 1.f : 0.f;
            float dinv = 0.f;
#pragma unroll
            for (int k4 = 0; k4 <= r; k4 += 4) {
                const float4 lv = *reinterpret_cast<const float4*>(sA + r * LD + k4);
                if (k4 < r) sacc = fmaf(-lv.x, x[k4], sacc); else if (k4 == r) dinv = lv.x;
                if (k4 + 1 < r) sacc = fmaf(-lv.y, x[k4 + 1], sacc); else if (k4 + 1 == r) dinv = lv.y;
                if (k4 + 2 < r) sacc = fmaf(-lv.z, x[k4 + 2], sacc); else if (k4 + 2 == r) dinv = lv.z;
                if (k4 + 3 < r) sacc = fmaf(-lv.w, x[k4 + 3], sacc); else if (k4 + 3 == r) dinv = lv.w;
            }
            x[r] = sacc * dinv;
        }
        // image rows 32..63 (row n, K index = lane) straight from the registers; L^-1 row-major and (over L) transposed
#pragma unroll
        for (int r = 0; r < H32; ++r) {
            sLi[r * LD + lane] = x[r];
            dst[sw128_offset(H32 + r, lane) / 4] = tf32_rna(x[r]);
        }
        __syncwarp();                                                        // every lane is done reading sA (as L)
#pragma unroll
        for (int r4 = 0; r4 < H32; r4 += 4)
            *reinterpret_cast<float4*>(sA + lane * LD + r4) = make_float4(x[r4], x[r4 + 1], x[r4 + 2], x[r4 + 3]);
        // ---- y = L^-1 b (lane = row), quad = |y|^2 = b^T M^-1 b
        float y = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < H32; c4 += 4) {
            const float4 lv = *reinterpret_cast<const float4*>(sLi + lane * LD + c4);
            const float4 bv = *reinterpret_cast<const float4*>(sb + c4);
            y = fmaf(lv.x, bv.x, y); y = fmaf(lv.y, bv.y, y); y = fmaf(lv.z, bv.z, y); y = fmaf(lv.w, bv.w, y);
        }
        sy[lane] = y;
        float quad = y * y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) quad += __shfl_xor_sync(0xffffffffu, quad, o);
        __syncwarp();
        // ---- a = L^-T y (lane = column of L^-1, in registers)
        float av = 0.f;
#pragma unroll
        for (int r4 = 0; r4 < H32; r4 += 4) {
            const float4 yv = *reinterpret_cast<const float4*>(sy + r4);
            av = fmaf(x[r4], yv.x, av); av = fmaf(x[r4 + 1], yv.y, av); av = fmaf(x[r4 + 2], yv.z, av); av = fmaf(x[r4 + 3], yv.w, av);
        }
        sa[lane] = av;
        dst[sw128_offset(2 * H32, lane) / 4] = tf32_rna(av);
        __syncwarp();
        if (PRED) {
            // ---- hmean, hcov = L^-T L^-1: column `lane` = sum_j Linv[j][r] Linv[j][lane]; sA row r = Linv[.][r] (broadcast)
            if (hmean && lane < Nh) hmean[(size_t)b * Nh + lane] = av;
            if (hcov) {
#pragma unroll
                for (int r = 0; r < H32; ++r) {
                    float sacc = 0.f;
#pragma unroll
                    for (int j4 = (r & ~3); j4 < H32; j4 += 4) {
                        const float4 lv = *reinterpret_cast<const float4*>(sA + r * LD + j4);   // Linv[j4..j4+3][r] (0 for j < r)
                        sacc = fmaf(lv.x, x[j4], sacc); sacc = fmaf(lv.y, x[j4 + 1], sacc);
                        sacc = fmaf(lv.z, x[j4 + 2], sacc); sacc = fmaf(lv.w, x[j4 + 3], sacc);
                    }
                    if (r < Nh && lane < Nh) hcov[((size_t)b * Nh + r) * Nh + lane] = sacc;
                }
            }
        } else {
        // ---- c = b2 - M2 a  (M2 symmetric: row `lane` = column `lane`)
#pragma unroll
        for (int k4 = 0; k4 < H32; k4 += 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(sa + k4);
            cv = fmaf(-m2c[k4], a4.x, cv); cv = fmaf(-m2c[k4 + 1], a4.y, cv);
            cv = fmaf(-m2c[k4 + 2], a4.z, cv); cv = fmaf(-m2c[k4 + 3], a4.w, cv);
        }
        dst[sw128_offset(2 * H32 + 1, lane) / 4] = tf32_rna(cv);
        // ---- W = L^-1 M2, column `lane`:  W[r] = sum_{c <= r} Linv[r][c] M2[c][lane]   (rows of L^-1 broadcast)
        float wv[H32];
#pragma unroll
        for (int r = 0; r < H32; ++r) {
            float sacc = 0.f;
#pragma unroll
            for (int c4 = 0; c4 <= r; c4 += 4) {
                const float4 lv = *reinterpret_cast<const float4*>(sLi + r * LD + c4);
                sacc = fmaf(lv.x, m2c[c4], sacc);
                if (c4 + 1 <= r) sacc = fmaf(lv.y, m2c[c4 + 1], sacc);
                if (c4 + 2 <= r) sacc = fmaf(lv.z, m2c[c4 + 2], sacc);
                if (c4 + 3 <= r) sacc = fmaf(lv.w, m2c[c4 + 3], sacc);
            }
            wv[r] = sacc;
        }
#if QFA_G32_CHAIN
        // ---- image rows 0..31 = W^T: row `lane`, column r = W[r][lane] (this lane's registers).  k_tc_grad32 forms
        //      (f^T K)_n = sum_r (L^-1 f)_r W[r][n] with a chained MMA (A = L^-1 f from tensor memory)
#pragma unroll
        for (int r4 = 0; r4 < H32; r4 += 4)
            *reinterpret_cast<float4*>(dst + sw128_offset(lane, r4) / 4) =
                make_float4(tf32_rna(wv[r4]), tf32_rna(wv[r4 + 1]), tf32_rna(wv[r4 + 2]), tf32_rna(wv[r4 + 3]));
#else
        // ---- K = L^-T W, column `lane`:  K[k] = sum_{r >= k} Linv[r][k] W[r]   (rows of L^-T broadcast).  Image rows 0..31:
        //      B[n][k] = K[k][n] -> this lane owns image row `lane`, four consecutive k per 16-byte store
#pragma unroll
        for (int k4 = 0; k4 < H32; k4 += 4) {
            float kk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k4 + q;
                float sacc = 0.f;
#pragma unroll
                for (int r4 = k4; r4 < H32; r4 += 4) {
                    const float4 lv = *reinterpret_cast<const float4*>(sA + k * LD + r4);
                    if (r4 >= k) sacc = fmaf(lv.x, wv[r4], sacc);
                    if (r4 + 1 >= k) sacc = fmaf(lv.y, wv[r4 + 1], sacc);
                    if (r4 + 2 >= k) sacc = fmaf(lv.z, wv[r4 + 2], sacc);
                    sacc = fmaf(lv.w, wv[r4 + 3], sacc);
                }
                kk[q] = tf32_rna(sacc);
            }
            *reinterpret_cast<float4*>(dst + sw128_offset(lane, k4) / 4) = make_float4(kk[0], kk[1], kk[2], kk[3]);
        }
#endif
        }
        if (lane == 0) {
            double E = (double)__ldg(src + G32_E);
            float hb = PRED ? 0.f : __ldg(src + G32_E + 1);
            if (diff) {                                           // four partial sums, one per CTA of the cluster
                E = ((double)__ldg(src + G32_E) + (double)__ldg(src + G32_E + 1)) + ((double)__ldg(src + G32_E + 2) + (double)__ldg(src + G32_E + 3));
                hb = (__ldg(src + G32_E + 4) + __ldg(src + G32_E + 5) + __ldg(src + G32_E + 6) + __ldg(src + G32_E + 7)) > 0.f ? 1.0f : 0.0f;
            }
            nll[b] = (float)(0.5 * (E - (double)quad + (double)logdet));                       // model.py:135
            if (!PRED) hasblue[b] = hb;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// k_tc_grad32: pixel-major gradient for 16 < Nh <= 32 with the two big per-cell contractions on the tensor cores
// ("per-spectrum stacked MMA").  For spectrum b and the CTA's 128-pixel tile:
//       D_b[i, 0..31]  = sum_l F_il K_b[l][n]        = (f_i^T K_b)_n                 (model.py:137, quirk Q2)
//       D_b[i, 32..63] = sum_l F_il Linv_b[n][l]     = (L_b^-1 f_i)_n   ->  q_i = |.|^2 = f_i^T M_b^-1 f_i   (model.py:136)
//       D_b[i, 64]     = sum_l F_il a_b[l]           = f_i^T hmean_b                 (model.py:136)
//   A = the tile's rows of F (one 128 x 32 K-major tile, built once per CTA), B = the 80-row image of spectrum b that
//   k_solve32 writes (rows 0..31 K^T, 32..63 L^-1, 64 a, 65 c, rest 0), accumulators in TMEM: 2 buffers x 3 spectra x 80
//   columns.  13 warps: 12 workers = 4 TMEM lane quadrants x 3 groups, group g owns spectrum 3n+g of step n (one cell per
//   thread and step), + the control warp (bulk copies of the images, every tcgen05.mma).
//   gradF_ik = f_ik sum_b s3_bi - sum_b [ s2_bi (f_i^T K_b)_k + (A u)_bi c_bk ] accumulates in 32 registers per thread.
// ---------------------------------------------------------------------------------------
struct TcGrad32Args {
    Field<float> f;
    int B;
    int nsplit;            // CTAs per pixel tile; grid = (npix_tiles, nsplit)
    const float* img;      // [B (+2 pad)][G32_IMG/4]  images written by k_solve32
    long long* trace;      // debug (-DQFA_ENABLE_TRACE): CTA (0,0), first 256 steps: [step][8] clock64 stamps
    float* part;           // [nsplit][part_len]
    float* spart;          // [nsplit][npix_tiles][3]
};

__global__ void __launch_bounds__(QFA_G32_CHAIN ? G32_THREADS_CHAIN : G32_THREADS, 1) k_tc_grad32(const TcGrad32Args g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_b[G32_NST], bar_tm_full[2], bar_tm_empty[2], bar_m1[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float sred2[3 * 32];
#if QFA_G32_CHAIN
    // ring stage: [3 compact blocks of 48 rows: L^-1 (32) | a | c | zero padding] [3 blocks of 32 rows: W^T];  TMEM buffer:
    // [3 x 48 columns: L^-1 f (32) | f.a | f.c | -] [3 x 32 columns: f^T K]
    constexpr int CROWS = 48, CBLK = CROWS * 128, WOFF = G32_SPS * CBLK, WBLK = H32 * 128, CLIVE = (H32 + 2) * 128;
    static_assert(WOFF + G32_SPS * WBLK <= G32_STAGE, "chained stage layout fits the ring stage");
#endif

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int quad = warp & 3, grp = warp >> 2;                   // workers: grp 0..2
    const Field<float>& f = g.f;
    const int P = f.P, Nb = f.Nb, Nh = f.Nh;
    const int pt = blockIdx.x;
    // spectra range of this CTA, in whole steps of 3
    const int nsteps_all = (g.B + G32_SPS - 1) / G32_SPS;
    const int st0 = (int)((long long)blockIdx.y * nsteps_all / g.nsplit);
    const int st1 = (int)((long long)(blockIdx.y + 1) * nsteps_all / g.nsplit);
    const int nst = st1 - st0;

    if (tid == 0) {
        for (int s = 0; s < G32_NST; ++s) mbar_init(&bar_b[s], 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&bar_tm_full[s], 1); mbar_init(&bar_tm_empty[s], G32_W); mbar_init(&bar_m1[s], 1); }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    const int pi = quad * 32 + lane;                 // pixel row inside the tile = TMEM lane
    const int i = pt * 128 + pi;
    const bool pix_ok = i < P;
    // A operand: this tile's rows of F, TF32, K-major SWIZZLE_128B (row = pixel); written by the 4 warps of group 0
    if (warp < 4) {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = c4 * 4 + q;
                v[q] = (pix_ok && k < Nh) ? tf32_rna(__ldg(f.F + (size_t)i * Nh + k)) : 0.0f;
            }
            sts_v4(smem_u32(sm) + G32_A_OFF + sw128_offset(pi, c4 * 4), v[0], v[1], v[2], v[3]);
        }
    }
#if QFA_G32_CHAIN
    // the padding rows of the compact blocks are zero and stay zero (the bulk copies only bring the live rows): zero the ring once
    for (int q = tid; q < G32_NST * G32_STAGE / 16; q += (int)blockDim.x) sts_v4(smem_u32(sm) + G32_B_OFF + q * 16, 0.f, 0.f, 0.f, 0.f);
#else
    // rows 66..79 of every ring image are zero and stay zero: the bulk copies only bring the 66 live rows
    for (int q = tid; q < G32_NST * G32_SPS * ((G32_IMG - G32_LIVE) / 16); q += G32_THREADS) {
        const int im = q / ((G32_IMG - G32_LIVE) / 16), o = q % ((G32_IMG - G32_LIVE) / 16);
        sts_v4(smem_u32(sm) + G32_B_OFF + im * G32_IMG + G32_LIVE + o * 16, 0.f, 0.f, 0.f, 0.f);
    }
#endif
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t sm_sa = smem_u32(sm);
    const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16);

    float gF[H32];
#pragma unroll
    for (int k = 0; k < H32; ++k) gF[k] = 0.f;
    float gPsi = 0.f, cnt = 0.f, dmu = 0.f, s3sum = 0.f, gOm = 0.f, sc0 = 0.f, sc1 = 0.f, sc2 = 0.f;

    if (warp == G32_W) {
        // =============================================================== CONTROL warp
        if (nst > 0 && elect_one()) {
#if QFA_G32_CHAIN
            const uint32_t idN1 = idesc_tf32(128, G32_SPS * CROWS), id32 = idesc_tf32(128, H32);
            auto issue_b = [&](int n) {           // step n: image rows 32..65 -> compact blocks, rows 0..31 (W^T) -> W blocks
                const int s = n % G32_NST;
                const size_t b0 = (size_t)(st0 + n) * G32_SPS;
                unsigned char* st = sm + G32_B_OFF + s * G32_STAGE;
                mbar_expect_tx(&bar_b[s], G32_SPS * (CLIVE + WBLK));
#pragma unroll
                for (int sp = 0; sp < G32_SPS; ++sp) {
                    const float* im = g.img + (b0 + sp) * (G32_IMG / 4);
                    bulk_g2s(st + sp * CBLK, im + H32 * 32, CLIVE, &bar_b[s]);
                    bulk_g2s(st + WOFF + sp * WBLK, im, WBLK, &bar_b[s]);
                }
            };
            for (int n = 0; n < 2 && n < nst; ++n) issue_b(n);
            const uint64_t dA = desc_sw128_kmajor(sm_sa + G32_A_OFF);
            for (int n = 0; n < nst; ++n) {
                const int buf = n & 1, s = n % G32_NST;
                if (n >= 2) mbar_wait_or_trap(&bar_tm_empty[buf], ((n >> 1) - 1) & 1);
                mbar_wait_or_trap(&bar_b[s], (n / G32_NST) & 1);
                fence_after_sync();
                {
                    const uint64_t dB = desc_sw128_kmajor(sm_sa + G32_B_OFF + s * G32_STAGE);
                    const uint32_t dcol = tmem + buf * G32_TBUF;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_tf32(dcol, dA + (uint64_t)(2 * kk), dB + (uint64_t)(2 * kk), idN1, kk > 0);
                }
                umma_commit(&bar_m1[buf]);
                if (n + 2 < nst) issue_b(n + 2);
            }
        }
#else
            const uint32_t idN = idesc_tf32(128, G32_SPS * G32_ROWS);     // the 3 images of a stage are ONE 240-row B operand
            auto issue_b = [&](int n) {           // images of step n -> ring stage n % 4
                const int s = n % G32_NST;              // (single thread, off the critical chain)
                const size_t b0 = (size_t)(st0 + n) * G32_SPS;
                mbar_expect_tx(&bar_b[s], G32_SPS * G32_LIVE);
#pragma unroll
                for (int sp = 0; sp < G32_SPS; ++sp)
                    bulk_g2s(sm + G32_B_OFF + s * G32_STAGE + sp * G32_IMG, g.img + (b0 + sp) * (G32_IMG / 4), G32_LIVE, &bar_b[s]);
            };
            for (int n = 0; n < 2 && n < nst; ++n) issue_b(n);
            const uint64_t dA = desc_sw128_kmajor(sm_sa + G32_A_OFF);
            for (int n = 0; n < nst; ++n) {
                const int buf = n & 1, s = n % G32_NST;
                // workers have drained TMEM buffer `buf` (step n-2): its MMAs retired long ago, so ring stage (n+2) % 4 = (n-2) % 4 is free
                long long* tr = (kTrace && g.trace && blockIdx.x == 0 && blockIdx.y == 0 && n < 256) ? g.trace + n * 8 : nullptr;
                if (tr) tr[0] = clock64();
                if (n >= 2) mbar_wait_or_trap(&bar_tm_empty[buf], ((n >> 1) - 1) & 1);
                if (tr) tr[1] = clock64();
                mbar_wait_or_trap(&bar_b[s], (n / G32_NST) & 1);
                fence_after_sync();
                if (tr) tr[2] = clock64();
                {
                    const uint64_t dB = desc_sw128_kmajor(sm_sa + G32_B_OFF + s * G32_STAGE);
                    const uint32_t dcol = tmem + buf * G32_TBUF;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_tf32(dcol, dA + (uint64_t)(2 * kk), dB + (uint64_t)(2 * kk), idN, kk > 0);
                }
                umma_commit(&bar_tm_full[buf]);
                if (n + 2 < nst) issue_b(n + 2);          // after the MMAs: off the empty -> full critical chain
                if (tr) tr[3] = clock64();
            }
        }
#endif
        __syncwarp();
#if QFA_G32_CHAIN
    } else if (warp == G32_W + 1) {
        // =============================================================== second CONTROL warp: the chained MMAs
        // (f^T K_b)[i, 0..31] = sum_r z[i, r] W_b[r][.] with A = z = the first MMA's result, read from tensor memory.  A warp of its
        // own: 12 more MMA issues per step on the first control thread made that thread the bottleneck (1 083 us).
        if (nst > 0 && elect_one()) {
            const uint32_t id32 = idesc_tf32(128, H32);
            for (int n = 0; n < nst; ++n) {
                const int buf = n & 1, s = n % G32_NST;
                mbar_wait_or_trap(&bar_m1[buf], (n >> 1) & 1);
                fence_after_sync();
                const uint32_t dcol = tmem + buf * G32_TBUF;
                const uint32_t wb = sm_sa + G32_B_OFF + s * G32_STAGE + WOFF;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                    for (int sp = 0; sp < G32_SPS; ++sp)
                        umma_tf32_ts(dcol + G32_SPS * CROWS + sp * H32, dcol + sp * CROWS + 8 * kk,
                                     desc_sw128_kmajor(wb + sp * WBLK) + (uint64_t)(2 * kk), id32, kk > 0);
                }
                umma_commit(&bar_tm_full[buf]);
            }
        }
        __syncwarp();
#endif
    } else {
        // =============================================================== WORKER warps
        const int ic = pix_ok ? i : P - 1;
        const bool blue = pix_ok && i < Nb;
        const bool any_blue = pt * 128 + quad * 32 < Nb;  // warp-uniform
        const int iz = i < Nb ? i : (Nb > 0 ? Nb - 1 : 0);
        const float psi = __ldg(f.Psi + ic);
        const float om = blue ? __ldg(f.omega + i) : 0.0f;
        const float tau0 = __ldg(f.scal + 0), c0s = __ldg(f.scal + 1), beta = __ldg(f.scal + 2);
        const float one_m_c0 = 1.0f - c0s, nt0l2e = -tau0 * kLog2e, l2zn = f.llogzn * kLog2e;
        struct Cell { float x, e, z; unsigned m; };
        // cells are requested strictly in step order: running pointers (one 64-bit add each per step, no multiplies)
        const size_t o0 = (size_t)(st0 * G32_SPS + grp) * P + ic;
        const float* cx = f.x + o0;
        const float* ce = f.err + o0;
        const uint8_t* cm = f.mask + o0;
        const float* cz = f.zabs + (size_t)(st0 * G32_SPS + grp) * Nb + iz;
        const size_t cstride = (size_t)G32_SPS * P, zstride = (size_t)G32_SPS * Nb;
        int cleft = pix_ok ? (g.B - grp + G32_SPS - 1) / G32_SPS - st0 : 0;       // steps whose spectrum exists
#if QFA_G32_CPASYNC
        const uint8_t* const mask_end = f.mask + (size_t)g.B * P;
        const uint32_t cell_sa = sm_sa + G32_CELL_OFF + (uint32_t)tid * 4u;          // [slot][x | e | z | mask word][thread]
        constexpr uint32_t CARR = G32_W * 32 * 4, CSLOT = 4 * CARR;
        uint32_t wslot = 0, rslot = 0;
        uint32_t rsh = (uint32_t)((uintptr_t)cm & 3u) * 8u;                              // bit offset of this thread's mask byte in its word
        const uint32_t rsh_step = (uint32_t)(cstride & 3u) * 8u;
        auto cp4 = [](uint32_t dst, const void* src) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
        };
        auto issue_cell = [&]() {
            const uint32_t dst = cell_sa + wslot * CSLOT;
            wslot = wslot + 1 == G32_CL ? 0 : wslot + 1;
            if (cleft > 0) {
                cp4(dst, cx);
                cp4(dst + CARR, ce);
                if (any_blue) cp4(dst + 2 * CARR, cz);
                const uint8_t* wa = reinterpret_cast<const uint8_t*>((uintptr_t)cm & ~(uintptr_t)3);   // the aligned word holding the byte
                if (wa + 4 <= mask_end) cp4(dst + 3 * CARR, wa);
                else asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + 3 * CARR), "r"(ldg_stream_u8(cm) << (((uintptr_t)cm & 3u) * 8u)) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            --cleft; cx += cstride; ce += cstride; cm += cstride; cz += zstride;
        };
        int rleft = cleft;                                    // read-side copy of the "spectrum exists" counter
        auto read_cell = [&](Cell& c) {
            asm volatile("cp.async.wait_group %0;" ::"n"(G32_CL - 1) : "memory");
            const uint32_t src = cell_sa + rslot * CSLOT;
            rslot = rslot + 1 == G32_CL ? 0 : rslot + 1;
            if (rleft > 0) {
                unsigned w;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c.x) : "r"(src));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c.e) : "r"(src + CARR));
                if (any_blue) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c.z) : "r"(src + 2 * CARR)); else c.z = 0.f;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(src + 3 * CARR));
                c.m = (w >> rsh) & 0xffu;
            } else { c.m = 0u; c.x = 0.f; c.e = 1.f; c.z = 0.f; }
            --rleft; rsh = (rsh + rsh_step) & 31u;
        };
#pragma unroll 1
        for (int q = 0; q < G32_CL; ++q) issue_cell();        // empty groups beyond the last step keep the group count uniform
#else
        auto load_cell = [&](int, Cell& c) {
            if (cleft > 0) {
                c.m = ldg_stream_u8(cm);
                c.x = ldg_stream(cx);
                c.e = ldg_stream(ce);
                c.z = any_blue ? ldg_stream(cz) : 0.f;
            } else { c.m = 0u; c.x = 0.f; c.e = 1.f; c.z = 0.f; }
            --cleft; cx += cstride; ce += cstride; cm += cstride; cz += zstride;
        };
        // the cell of step n is requested G32_CELL_LEAD steps (>= 2 us) ahead: one step is less than the loaded HBM latency
        Cell cc[G32_CELL_LEAD];
#pragma unroll
        for (int q = 0; q < G32_CELL_LEAD; ++q)
            if (q < nst) load_cell(q, cc[q]);
#endif
        int wstage = 0;                                    // ring stage of the next step (n % G32_NST without the division)
#if QFA_G32_CPASYNC
        auto do_step = [&](int n) {
            Cell cb;
            read_cell(cb);
#else
        auto do_step = [&](int n, Cell& cb) {
#endif
            const int buf = n & 1, s = wstage;
            wstage = wstage + 1 == G32_NST ? 0 : wstage + 1;
            const int b = (st0 + n) * G32_SPS + grp;
            const bool valid = b < g.B;                       // warp-uniform
            long long* tr = (kTrace && g.trace && blockIdx.x == 0 && blockIdx.y == 0 && n < 256 && warp == 5 && lane == 0)
                                ? g.trace + n * 8 : nullptr;
            if (tr) tr[4] = clock64();
            mbar_wait_or_trap(&bar_tm_full[buf], (n >> 1) & 1);
            fence_after_sync();
            if (tr) tr[5] = clock64();
            // The TMEM buffer is handed back as soon as its 65 columns are in registers, BEFORE the math: the control warp then
            // runs a full step ahead and the workers never wait for an accumulator.  (The ring stage is still read below -- c,
            // image row 65 -- which is why the ring has 5 stages: the copy for step n+2 goes to the stage of step n-3.)
            float z[2][16], y[2][16], w8[8];
#if QFA_G32_CHAIN
            const uint32_t tz = ta + buf * G32_TBUF + grp * CROWS, tcol = ta + buf * G32_TBUF + G32_SPS * CROWS + grp * H32;
            tmem_ld16(tz, z[0]); tmem_ld16(tz + 16, z[1]); tmem_ld8(tz + 32, w8);
#else
            const uint32_t tcol = ta + buf * G32_TBUF + grp * G32_ROWS;
            tmem_ld16(tcol + 32, z[0]); tmem_ld16(tcol + 48, z[1]); tmem_ld8(tcol + 64, w8);
#endif
            tmem_wait_ld();
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int t = 0; t < 16; t += 2) {
                q0 = fmaf(z[0][t], z[0][t], q0); q1 = fmaf(z[0][t + 1], z[0][t + 1], q1);
                q2 = fmaf(z[1][t], z[1][t], q2); q3 = fmaf(z[1][t + 1], z[1][t + 1], q3);
            }
            tmem_ld16(tcol, y[0]); tmem_ld16(tcol + 16, y[1]);
            tmem_wait_ld();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tm_empty[buf]);
            const float q = (q0 + q1) + (q2 + q3), fa = w8[0];
            const bool mk = cb.m != 0u && valid;
            float A = 1.0f, zdep = 0.0f, powb = 0.0f, L2 = 0.0f;
            if (any_blue) {
                L2 = lg2f(1.0f + (mk ? cb.z : 0.0f));
                const float tau = fmaf(f.lt0, ex2f(f.lbe * (L2 - l2zn)), f.lC);       // utils.py:106 etc.
                const float Ab = ex2f(-kLog2e * tau);                                   // model.py:125
                powb = ex2f(beta * L2);                                                 // utils.py:72
                const float root = one_m_c0 - ex2f(nt0l2e * powb);                      // utils.py:91
                A = blue ? Ab : 1.0f;
                zdep = root * root;
            }
            const float A2 = A * A;
            const float oz = om * zdep;
            const float D = fmaf(A2, psi, fmaf(cb.e, cb.e, oz));                        // model.py:128-131
            const float w = mk ? rcpf(D) : 0.0f;
            const float r = mk ? cb.x : 0.0f;
            const float u = w * fmaf(-A, fa, r);                                        // (Sigma^-1 delta)_i
            const float s2 = w * A2;
            const float gd = 0.5f * (w - w * s2 * q - u * u);                           // model.py:136,138
            const float Au = A * u;
            if (valid) {                                       // an absent spectrum has no image: its TMEM columns are garbage
                gPsi = fmaf(A2, gd, gPsi);                                              // model.py:139
                cnt += mk ? 1.0f : 0.0f;
                dmu -= Au;
                s3sum = fmaf(s2, A, s3sum);
                if (any_blue) {
                    gOm = fmaf(gd, zdep, gOm);                                          // model.py:140
                    const float rootl = 1.0f - tau0 * powb - c0s;                       // model.py:141 (quirk Q3)
                    const float t = gd * oz * zdep * 2.0f * rootl;
                    sc0 = fmaf(-t, powb, sc0);                                          // model.py:142
                    sc1 -= t;                                                           // model.py:144
                    sc2 = fmaf(-t, tau0 * powb * (L2 * kLn2), sc2);                     // model.py:143
                }
                // gradF: - s2 (f^T K)_k - (A u) c_k ; c_b = image row 65 (shared memory, broadcast reads)
#if QFA_G32_CHAIN
                const uint32_t crow = sm_sa + G32_B_OFF + (uint32_t)s * G32_STAGE + (uint32_t)grp * CBLK;
                constexpr int CROW_IDX = H32 + 1;                         // compact block row 33 = c
#else
                const uint32_t crow = sm_sa + G32_B_OFF + (uint32_t)s * G32_STAGE + (uint32_t)grp * G32_IMG;
                constexpr int CROW_IDX = 2 * H32 + 1;                     // image row 65 = c
#endif
                const float ns2 = -s2, nAu = -Au;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float cx, cy, cz, cw;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cx), "=f"(cy), "=f"(cz), "=f"(cw)
                                 : "r"(crow + sw128_offset(CROW_IDX, 4 * c4)));
                    const int k = 4 * c4;
                    gF[k] = fmaf(ns2, y[k >> 4][k & 15], fmaf(nAu, cx, gF[k]));
                    gF[k + 1] = fmaf(ns2, y[k >> 4][(k & 15) + 1], fmaf(nAu, cy, gF[k + 1]));
                    gF[k + 2] = fmaf(ns2, y[k >> 4][(k & 15) + 2], fmaf(nAu, cz, gF[k + 2]));
                    gF[k + 3] = fmaf(ns2, y[k >> 4][(k & 15) + 3], fmaf(nAu, cw, gF[k + 3]));
                }
            }
            if (tr) tr[6] = clock64();
#if QFA_G32_CPASYNC
            issue_cell();                                     // the cell of step n + G32_CL
        };
        for (int n = 0; n < nst; n += G32_CELL_LEAD) {
#pragma unroll
            for (int q = 0; q < G32_CELL_LEAD; ++q)
                if (n + q < nst) do_step(n + q);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#else
            if (n + G32_CELL_LEAD < nst) load_cell(n + G32_CELL_LEAD, cb);
        };
        for (int n = 0; n < nst; n += G32_CELL_LEAD) {
#pragma unroll
            for (int q = 0; q < G32_CELL_LEAD; ++q)
                if (n + q < nst) do_step(n + q, cc[q]);
        }
#endif
    }

    // ---- end of the CTA: fold the three groups, finish gradF, write the per-split partials
    float* part = g.part + (size_t)blockIdx.y * part_len(P, Nb, Nh);
    float* red = reinterpret_cast<float*>(sm + G32_RED_OFF);     // [3 grp][128 pixels][40]
    fence_before_sync();
    __syncthreads();
    if (warp < G32_W) {
        float* r = red + ((size_t)grp * 128 + pi) * 40;
#pragma unroll
        for (int k = 0; k < H32; ++k) r[k] = gF[k];
        r[32] = gPsi; r[33] = cnt; r[34] = dmu; r[35] = s3sum; r[36] = gOm; r[37] = sc0; r[38] = sc1; r[39] = sc2;
    }
    __syncthreads();
    float sc3[3] = {0.f, 0.f, 0.f};
    if (warp < 4) {
        float v[40];
#pragma unroll
        for (int q = 0; q < 40; ++q)
            v[q] = red[((size_t)0 * 128 + pi) * 40 + q] + red[((size_t)1 * 128 + pi) * 40 + q] + red[((size_t)2 * 128 + pi) * 40 + q];
        if (pix_ok) {
            float* pF = part + (size_t)i * Nh;
            float* pPsi = part + (size_t)P * Nh + i;
            float* pOm = part + (size_t)P * Nh + P + i;
            float* pCnt = part + (size_t)P * Nh + P + Nb + i;
            float* pMu = part + (size_t)P * Nh + 2 * (size_t)P + Nb + i;
#pragma unroll
            for (int k = 0; k < H32; ++k)
                if (k < Nh) pF[k] = fmaf(__ldg(f.F + (size_t)i * Nh + k), v[35], v[k]);      // model.py:137 (quirk Q2)
            *pPsi = v[32]; *pCnt = v[33]; *pMu = v[34];
            if (i < Nb) *pOm = v[36];
        }
        sc3[0] = v[37]; sc3[1] = v[38]; sc3[2] = v[39];
    }
    if (pt * 128 < Nb) {
        float a0 = warp_sum(sc3[0]), a1 = warp_sum(sc3[1]), a2 = warp_sum(sc3[2]);
        if (lane == 0 && warp < 4) { sred2[warp] = a0; sred2[32 + warp] = a1; sred2[64 + warp] = a2; }
        __syncthreads();
        if (tid < 3) {
            float t = 0.f;
            for (int w = 0; w < 4; ++w) t += sred2[tid * 32 + w];
            g.spart[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 + tid] = t;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---------------------------------------------------------------------------------------
// k_out32: continuum and 1-sigma on the full grid for 8 < Nh <= 32 (model.py:180), cut from k_tc_grad32: for spectrum b and
// the CTA's 128-pixel tile   D_b[i, 32..63] = (L_b^-1 f_i)  ->  unc_i = |.|   ,   D_b[i, 64] = f_i^T hmean_b -> cont_i = mu_i + .
// A = the tile's rows of F, B = the image rows [L^-1 | a] that k_solve32<PRED> wrote; same ring / TMEM double buffering.
// lane = pixel: every store instruction writes 128 consecutive bytes of one spectrum's row.
// ---------------------------------------------------------------------------------------
struct TcOut32Args {
    Field<float> f;
    int B;
    int nsplit;
    const float* img;      // [B (+2 pad)][G32_IMG/4]
    float* cont;           // [B][P] or nullptr
    float* unc;            // [B][P] or nullptr
};

__global__ void __launch_bounds__(G32_THREADS, 1) k_out32(const TcOut32Args g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_b[G32_NST], bar_tm_full[2], bar_tm_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int quad = warp & 3, grp = warp >> 2;
    const Field<float>& f = g.f;
    const int P = f.P, Nh = f.Nh;
    const int pt = blockIdx.x;
    const int nsteps_all = (g.B + G32_SPS - 1) / G32_SPS;
    const int st0 = (int)((long long)blockIdx.y * nsteps_all / g.nsplit);
    const int st1 = (int)((long long)(blockIdx.y + 1) * nsteps_all / g.nsplit);
    const int nst = st1 - st0;
    if (tid == 0) {
        for (int s = 0; s < G32_NST; ++s) mbar_init(&bar_b[s], 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&bar_tm_full[s], 1); mbar_init(&bar_tm_empty[s], G32_W); }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    const int pi = quad * 32 + lane;
    const int i = pt * 128 + pi;
    const bool pix_ok = i < P;
    if (warp < 4) {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = c4 * 4 + q;
                v[q] = (pix_ok && k < Nh) ? tf32_rna(__ldg(f.F + (size_t)i * Nh + k)) : 0.0f;
            }
            sts_v4(smem_u32(sm) + G32_A_OFF + sw128_offset(pi, c4 * 4), v[0], v[1], v[2], v[3]);
        }
    }
    // the whole ring is zeroed once: image rows 0..31 and 65.. are never copied in the prediction flavour (the copies bring rows
    // 32..64 only), and an MMA must not read uninitialised shared memory (NaN patterns would not matter -- D columns are
    // independent -- but stay deterministic)
    for (int q = tid; q < G32_NST * G32_STAGE / 16; q += G32_THREADS) sts_v4(smem_u32(sm) + G32_B_OFF + q * 16, 0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t sm_sa = smem_u32(sm);
    const uint32_t ta = tmem + ((uint32_t)(quad * 32) << 16);
    constexpr int ROW0 = H32, NROWS = H32 + 1;          // image rows 32..64: L^-1 and a

    if (warp == G32_W) {
        if (nst > 0 && elect_one()) {
            constexpr int OROWS = 48;                                    // ring rows per spectrum: L^-1 (32), a (1), zero padding (15)
            const uint32_t idN = idesc_tf32(128, G32_SPS * OROWS);       // the three spectra of a step are ONE 144-row B operand
            auto issue_b = [&](int n) {
                const int s = n % G32_NST;
                const size_t b0 = (size_t)(st0 + n) * G32_SPS;
                mbar_expect_tx(&bar_b[s], G32_SPS * NROWS * 128);
#pragma unroll
                for (int sp = 0; sp < G32_SPS; ++sp)                     // image rows 32..64 -> rows 0..32 of the spectrum's compact block
                    bulk_g2s(sm + G32_B_OFF + s * G32_STAGE + sp * OROWS * 128, g.img + (b0 + sp) * (G32_IMG / 4) + ROW0 * 32,
                             NROWS * 128, &bar_b[s]);
            };
            for (int n = 0; n < 2 && n < nst; ++n) issue_b(n);
            const uint64_t dA = desc_sw128_kmajor(sm_sa + G32_A_OFF);
            for (int n = 0; n < nst; ++n) {
                const int buf = n & 1, s = n % G32_NST;
                if (n >= 2) mbar_wait_or_trap(&bar_tm_empty[buf], ((n >> 1) - 1) & 1);
                mbar_wait_or_trap(&bar_b[s], (n / G32_NST) & 1);
                fence_after_sync();
                // (the N = 240 MMA over whole 80-row images that k_tc_grad32 issues made this kernel tensor-bound: 617 us; one
                //  N = 48 MMA per spectrum made it issue-bound in the control thread: 749 us)
                const uint64_t dB = desc_sw128_kmajor(sm_sa + G32_B_OFF + s * G32_STAGE);
                const uint32_t dcol = tmem + buf * G32_TBUF;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_tf32(dcol, dA + (uint64_t)(2 * kk), dB + (uint64_t)(2 * kk), idN, kk > 0);
                umma_commit(&bar_tm_full[buf]);
                if (n + 2 < nst) issue_b(n + 2);
            }
        }
        __syncwarp();
    } else if (warp < G32_W) {
        const float mu = pix_ok ? __ldg(f.mu + i) : 0.f;
        for (int n = 0; n < nst; ++n) {
            const int buf = n & 1;
            const int b = (st0 + n) * G32_SPS + grp;
            mbar_wait_or_trap(&bar_tm_full[buf], (n >> 1) & 1);
            fence_after_sync();
            const uint32_t tcol = ta + buf * G32_TBUF + grp * 48;         // this spectrum's 48 columns: L^-1 f (32) | f.a (1) | 0
            float z[2][16], w8[8];
            tmem_ld16(tcol, z[0]); tmem_ld16(tcol + 16, z[1]); tmem_ld8(tcol + 32, w8);
            tmem_wait_ld();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tm_empty[buf]);
            float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
            for (int t = 0; t < 16; t += 2) {
                q0 = fmaf(z[0][t], z[0][t], q0); q1 = fmaf(z[0][t + 1], z[0][t + 1], q1);
                q2 = fmaf(z[1][t], z[1][t], q2); q3 = fmaf(z[1][t + 1], z[1][t + 1], q3);
            }
            if (b < g.B && pix_ok) {
                if (g.cont) st_stream(g.cont + (size_t)b * P + i, mu + w8[0]);                         // model.py:180
                if (g.unc) st_stream(g.unc + (size_t)b * P + i, sqrtaf((q0 + q1) + (q2 + q3)));
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace tcg32
}  // namespace qfa
