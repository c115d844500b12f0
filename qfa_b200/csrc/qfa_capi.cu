// C ABI of libqfa_b200.so (see include/qfa_b200.h) + the small parameter-side kernels.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/qfa_b200.h"
#include "../../include/qfa_b200_debug.h"
#include "qfa_aux.cuh"
#include "qfa_kernels.cuh"
#include "qfa_tc_selftest.cuh"
#include "qfa_tc_gram.cuh"
#include "qfa_tc_grad.cuh"
#include "qfa_tc_gram32.cuh"
#include "qfa_tc_gram32c.cuh"
#include "qfa_tc_gram_x3.cuh"
#include "qfa_peer.cuh"

using namespace qfa;

// ---------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
#define CK(call)                                           \
    do {                                                   \
        cudaError_t e_ = (call);                           \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

extern "C" int qfa_abi_version(void) { return QFA_ABI_VERSION; }

extern "C" const char* qfa_last_error_string(void) { return g_err; }

// Kernel launches issued by this library in this process (bench.py's gpu_launches is this counter, not a constant).
static unsigned long long g_launches = 0;
#define QFA_LAUNCHED() (++g_launches)
extern "C" unsigned long long qfa_launch_count(void) { return g_launches; }

// debug (qfa_b200_debug.h): clock64 trace buffers of k_tc_gram / k_tc_grad.  Only a -DQFA_ENABLE_TRACE build keeps these
// pointers; the production library keeps no pointer between calls and the setters report QFA_ERR_UNSUPPORTED.
#ifdef QFA_ENABLE_TRACE
static long long* g_trace = nullptr;
static long long* g_trace_grad = nullptr;
extern "C" int qfa_debug_set_trace(void* device_buffer) { g_trace = (long long*)device_buffer; return 0; }
extern "C" int qfa_debug_set_trace_grad(void* device_buffer) { g_trace_grad = (long long*)device_buffer; return 0; }
#else
static constexpr long long* g_trace = nullptr;
static constexpr long long* g_trace_grad = nullptr;
extern "C" int qfa_debug_set_trace(void*) { return fail(QFA_ERR_UNSUPPORTED, "built without -DQFA_ENABLE_TRACE"); }
extern "C" int qfa_debug_set_trace_grad(void*) { return fail(QFA_ERR_UNSUPPORTED, "built without -DQFA_ENABLE_TRACE"); }
#endif

// ---------------------------------------------------------------------------------------
// layouts
// ---------------------------------------------------------------------------------------
extern "C" size_t qfa_param_len(int Nb, int Nr, int Nh) {
    size_t P = (size_t)Nb + Nr;
    return P * Nh + P + Nb + 3;
}
extern "C" size_t qfa_acc_len(int Nb, int Nr, int Nh) {
    size_t P = (size_t)Nb + Nr;
    return P * Nh + P + Nb + 3 + P + 3 + 1 + 1 + P;
}
static inline int pad_h(int Nh) { return Nh <= 4 ? 4 : Nh <= 8 ? 8 : Nh <= 16 ? 16 : 32; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline size_t tsize(int precision) { return precision == QFA_PREC_FP64 ? 8 : 4; }

// Per-DEVICE state: the SM count and the "dynamic shared memory opt-in done" flags belong to the device that is current
// when an entry point runs (cudaFuncSetAttribute applies to the current device only), so they are keyed by its ordinal.
constexpr int kMaxDevices = 64;
static int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
static int num_sms() {
    static int sms[kMaxDevices] = {0};
    const int dev = current_device();
    if (sms[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms[dev] = n > 0 ? n : 148;
    }
    return sms[dev];
}
// one flag per (call site, device)
struct PerDeviceOnce {
    bool done[kMaxDevices] = {false};
};

// Spectra per sub-batch: sized so that one sub-batch of inputs stays L2-resident between the
// spectrum-major pass and the pixel-major pass (B200: 126 MB L2).
static int subbatch_size(int P, int Nb, int B) {
    static long mb = -1;
    if (mb < 0) {
        const char* s = getenv("QFA_SUBBATCH_MB");
        mb = s ? atol(s) : 48;
        if (mb <= 0) mb = 48;
    }
    size_t per = (size_t)9 * P + (size_t)4 * Nb;
    long sb = (long)((size_t)mb * 1024 * 1024 / per);
    if (sb < 128) sb = 128;
    if (sb > B) sb = B;
    return (int)sb;
}

struct TrainPlan {
    int HP, SB, ntiles, ntiles_blue, nsplit;
    size_t off_small, off_hasblue, off_nll, off_part, off_spart, total;
};
static TrainPlan plan_train(int Nb, int Nr, int Nh, int B, int precision) {
    TrainPlan p;
    const int P = Nb + Nr;
    const size_t ts = tsize(precision);
    p.HP = pad_h(Nh);
    p.SB = subbatch_size(P, Nb, B > 0 ? B : 1);
    p.ntiles = (P + 127) / 128;
    p.ntiles_blue = (Nb + 127) / 128;
    int want = (4 * num_sms() + p.ntiles - 1) / p.ntiles;
    if (want > p.SB) want = p.SB;
    if (want < 1) want = 1;
    p.nsplit = want;
    size_t small_len = 2 * (size_t)p.HP + 2 * (size_t)p.HP * p.HP;
    size_t o = 0;
    p.off_small = o;   o = align_up(o + (size_t)p.SB * small_len * ts, 256);
    p.off_hasblue = o; o = align_up(o + (size_t)(B > 0 ? B : 1) * 4, 256);
    p.off_nll = o;     o = align_up(o + (size_t)(B > 0 ? B : 1) * ts, 256);
    p.off_part = o;    o = align_up(o + (size_t)p.nsplit * part_len(P, Nb, Nh) * ts, 256);
    p.off_spart = o;   o = align_up(o + (size_t)p.nsplit * p.ntiles * 3 * ts, 256);
    p.total = o;
    return p;
}
// ---- tensor-core path (QFA_PREC_TF32, Nh <= 8): static operand images live in the workspace
struct TcPlan {
    int nkb, npt, ntiles;
    tcg::TileSched ts;
    size_t off_pb, off_qa, total;
};
static inline bool tc_eligible(int Nh, int precision) { return precision == QFA_PREC_TF32 && Nh <= tcg::HP; }
// The tensor-core kernels work on tiles of up to 120 spectra per CTA and have a latency floor of one pass over all K-blocks;
// below a path-specific batch size the float CUDA-core kernels (one CTA per spectrum) are faster.  Measured cross-overs on a
// B200 (profiles/r1_small_batch.txt): predict ~1 200 spectra, train step Nh <= 8 ~500, train step 8 < Nh <= 32 ~170.
// QFA_TC_MIN_BATCH (env) overrides all three; QFA_FLAG_FORCE_TENSOR forces the tensor-core kernels (tests, profiling).
enum { TC_PATH_PREDICT = 0, TC_PATH_TRAIN = 1, TC_PATH_TRAIN32 = 2, TC_PATH_PREDICT32 = 3 };
static int tc_min_batch(int path) {
    static int v = -2;
    if (v == -2) {
        const char* s = getenv("QFA_TC_MIN_BATCH");
        v = s ? atoi(s) : -1;
        if (s && v < 0) v = 0;
    }
    if (v >= 0) return v;
    return path == TC_PATH_PREDICT ? 1280 : (path == TC_PATH_TRAIN ? 800 : (path == TC_PATH_TRAIN32 ? 192 : 640));
}
static inline bool tc_use(int Nh, int precision, int B, int flags, int path) {
    return tc_eligible(Nh, precision) && (B >= tc_min_batch(path) || (flags & QFA_FLAG_FORCE_TENSOR));
}

// Tile heights: the tile count should be a whole number of waves of num_sms() persistent CTAs.  Heights are
// multiples of 8 (a worker warp of k_tc_gram owns 8 consecutive rows), at most tcg::TSH = 120 (15 worker warps);
// n_hi tiles are 8 rows taller than the others.
static tcg::TileSched tile_sched(int B, int quantum, int* ntiles) {
    tcg::TileSched ts;
    const long nsm = num_sms();
    const long H = tcg::TSH;
    if (B <= 0) { ts.r_hi = ts.r_lo = (int)H; ts.n_hi = 0; *ntiles = 0; return ts; }
    const long waves = ((long)B + nsm * H - 1) / (nsm * H);
    const long slots = nsm * waves;
    long r_lo = (B / slots) / quantum * quantum;
    if (r_lo < quantum) r_lo = quantum;
    if (r_lo > H) r_lo = H;
    long r_hi = r_lo + quantum > H ? H : r_lo + quantum;
    long rest = (long)B - slots * r_lo;
    long n_hi = (rest > 0 && r_hi > r_lo) ? (rest + (r_hi - r_lo) - 1) / (r_hi - r_lo) : 0;
    if (n_hi > slots) n_hi = slots;
    ts.r_hi = (int)r_hi; ts.r_lo = (int)r_lo; ts.n_hi = (int)n_hi;
    long covered = n_hi * r_hi, t = n_hi;
    if (covered < B) t += ((long)B - covered + r_lo - 1) / r_lo;
    else t = ((long)B + r_hi - 1) / r_hi;
    *ntiles = (int)t;
    return ts;
}
static TcPlan plan_tc(int Nb, int Nr, int B, bool want_qa, int quantum = 1) {
    TcPlan p;
    const int P = Nb + Nr;
    p.nkb = (P + tcg::KB - 1) / tcg::KB;
    p.npt = (P + tcg::PT - 1) / tcg::PT;
    p.ts = tile_sched(B, quantum < 8 ? 8 : quantum, &p.ntiles);
    size_t o = 0;
    p.off_pb = o; o = align_up(o + (size_t)p.nkb * tcg::PB_TILE, 1024);
    p.off_qa = o; if (want_qa) o = align_up(o + (size_t)p.npt * tcg::QA_TILE, 1024);
    p.total = o + 256;
    return p;
}
struct TcTrainPlan {
    TcPlan t;
    int nchunks, gct, nsplit, ntiles_blue;      // nsplit = max(ns_blue, ns_red): number of partial buffers
    int ns_blue, ns_red;
    size_t off_b2, off_kc, off_tsums, off_part, off_spart, total;
};
static TcTrainPlan plan_tc_train(int Nb, int Nr, int Nh, int B) {
    TcTrainPlan p;
    const int P = Nb + Nr;
    p.t = plan_tc(Nb, Nr, B, true, 8);
    {   // chunks of GC spectra per tile: enough for the tallest tile of this launch (a 56-row tile has 3, not 5)
        const int tall = p.t.ts.n_hi > 0 ? p.t.ts.r_hi : p.t.ts.r_lo;
        p.gct = (tall + tcg::GC - 1) / tcg::GC;
        if (p.gct < 1) p.gct = 1;
        if (p.gct > tcg::GCT) p.gct = tcg::GCT;
    }
    p.nchunks = p.t.ntiles * p.gct;
    p.ntiles_blue = (Nb + tcg::PT - 1) / tcg::PT;
    if (p.ntiles_blue > p.t.npt) p.ntiles_blue = p.t.npt;
    {
        // k_tc_grad runs ONE wave of CTAs; a pixel tile with blue pixels costs `ratio` times a red one per spectrum
        // (QFA_GRAD_BLUE_COST, default 1.5, measured).  Pick the CTAs per tile that minimise the slowest CTA.
        static double ratio = -1.0;
        if (ratio < 0) { const char* e = getenv("QFA_GRAD_BLUE_COST"); ratio = e ? atof(e) : 1.5; if (ratio < 1.0) ratio = 1.0; }
        const int ntb = p.ntiles_blue, ntr = p.t.npt - ntb, nsm = num_sms();
        int best_b = 1, best_r = 1; double best = 1e30;
        for (int r = 1; r <= nsm; ++r) {
            int b = ntb > 0 ? (nsm - ntr * r) / ntb : 1;
            if (ntr == 0) { b = nsm / (ntb > 0 ? ntb : 1); }
            if (b < 1) break;
            const double t = (ntb > 0 ? ratio / b : 0.0) > (ntr > 0 ? 1.0 / r : 0.0) ? ratio / b : 1.0 / r;
            if (t < best) { best = t; best_b = b; best_r = r; }
            if (ntr == 0) break;
        }
        if (best_b > p.nchunks) best_b = p.nchunks > 0 ? p.nchunks : 1;
        if (best_r > p.nchunks) best_r = p.nchunks > 0 ? p.nchunks : 1;
        p.ns_blue = best_b; p.ns_red = best_r;
        p.nsplit = best_b > best_r ? best_b : best_r;
    }
    size_t o = p.t.total;
    o = align_up(o, 1024);
    p.off_b2 = o;      o = align_up(o + (size_t)p.t.ntiles * tcg::B2_TILE, 1024);
    p.off_kc = o;      o = align_up(o + (size_t)p.t.ntiles * 4 * tcg::KC_TILE, 1024);
    p.off_tsums = o;   o = align_up(o + (size_t)(p.t.ntiles > 0 ? p.t.ntiles : 1) * 4 * 2 * 4, 256);
    p.off_part = o;    o = align_up(o + (size_t)p.nsplit * part_len(P, Nb, Nh) * 4, 256);
    p.off_spart = o;   o = align_up(o + (size_t)p.nsplit * p.t.npt * 3 * 4, 256);
    p.total = o;
    return p;
}

// ---- tensor-core train path for 8 < Nh <= 32 (QFA_PREC_TF32): k_tc_gram32 + k_solve32 + k_tc_grad32
struct Tc32Plan {
    int nkb, ntiles, npix_tiles, nsplit, ntiles_blue;
    tcg::TileSched ts;
    size_t off_pb, off_pb2, off_gram, off_small, off_hasblue, off_nll, off_part, off_spart, off_replay, total;
};
static inline bool tc32_eligible(int Nh, int precision) { return precision == QFA_PREC_TF32 && Nh > tcg::HP && Nh <= 32; }   // zero-padded to 32
static Tc32Plan plan_tc32(int Nb, int Nr, int Nh, int B, bool predict = false) {
    Tc32Plan p;
    const int P = Nb + Nr;
    p.nkb = (P + tcg::KB - 1) / tcg::KB;
    p.ts = tile_sched(B, 8, &p.ntiles);
    p.npix_tiles = (P + 127) / 128;
    p.ntiles_blue = (Nb + 127) / 128;
    int want = num_sms() / p.npix_tiles;         // one wave of k_tc_grad32 CTAs
    const int nsteps = (B + tcg32::G32_SPS - 1) / tcg32::G32_SPS;
    if (want > nsteps) want = nsteps;
    if (want < 1) want = 1;
    p.nsplit = want;
    const size_t Bn = B > 0 ? B : 1;
    size_t o = 0;
    p.off_pb = o;      o = align_up(o + (size_t)p.nkb * tcg32::PB32_KB_BYTES + 16, 1024);
    p.off_pb2 = o;     o = align_up(o + (predict ? (size_t)p.nkb * tcg32::PB32_KB_BYTES + 16 : 0), 1024);   // residual image (3xTF32)
    p.off_gram = o;    o = align_up(o + (predict ? 3 : 1) * Bn * tcg32::G32_STRIDE * 4, 256);
    p.off_small = o;   o = align_up(o + (Bn + 2) * tcg32::G32_IMG, 1024);      // per-spectrum images for k_tc_grad32 (+2: last step)
    p.off_hasblue = o; o = align_up(o + Bn * 4, 256);
    p.off_nll = o;     o = align_up(o + Bn * 4, 256);
    p.off_part = o;    o = align_up(o + (size_t)p.nsplit * part_len(P, Nb, Nh) * 4, 256);
    p.off_spart = o;   o = align_up(o + (size_t)p.nsplit * p.npix_tiles * 3 * 4, 256);
    // k_tc_gram32 replay buffer: one tile's s2 / s3 operand tiles per CTA (independent of B; 1 MB per CTA at 1000 pixels)
    const int gram_ctas = p.ntiles < num_sms() ? p.ntiles : num_sms();
    p.off_replay = o = align_up(o, 1024);
    o = align_up(o + (size_t)(gram_ctas > 0 ? gram_ctas : 1) * p.nkb * tcg32::RP_KB_FLOATS * 4, 1024);
    p.total = o;
    return p;
}

// ---- 3xTF32 path (QFA_FLAG_TF32X3, Nh <= 8): k_tc_gram_x3 (+ k_grad<float, 8> + k_reduce for the train step)
struct TcX3Plan {
    int nkb, npt, ntiles, nsplit, ntiles_blue;
    tcg::TileSched ts;
    size_t off_pb, off_qa, off_small, off_hasblue, off_nll, off_part, off_spart, total;
};
static TcX3Plan plan_tc_x3(int Nb, int Nr, int Nh, int B, bool train) {
    TcX3Plan p;
    const int P = Nb + Nr;
    p.nkb = (P + tcx::XKB - 1) / tcx::XKB;
    p.npt = (P + tcg::PT - 1) / tcg::PT;
    p.ts = tile_sched(B, 8, &p.ntiles);
    p.ntiles_blue = (Nb + 127) / 128;
    int want = (4 * num_sms() + p.npt - 1) / p.npt;
    if (want > B) want = B;
    if (want < 1) want = 1;
    p.nsplit = want;
    const size_t Bn = B > 0 ? B : 1;
    size_t o = 0;
    p.off_pb = o; o = align_up(o + (size_t)p.nkb * tcx::XPB_TILE, 1024);
    p.off_qa = o; o = align_up(o + (size_t)p.npt * tcx::XQA_TILE, 1024);
    p.off_small = p.off_hasblue = p.off_nll = p.off_part = p.off_spart = o;
    if (train) {
        p.off_small = o;   o = align_up(o + Bn * SmallLayout<8>::len * 4, 256);
        p.off_hasblue = o; o = align_up(o + Bn * 4, 256);
        p.off_nll = o;     o = align_up(o + Bn * 4, 256);
        p.off_part = o;    o = align_up(o + (size_t)p.nsplit * part_len(P, Nb, Nh) * 4, 256);
        p.off_spart = o;   o = align_up(o + (size_t)p.nsplit * p.npt * 3 * 4, 256);
    }
    p.total = o + 256;
    return p;
}

extern "C" size_t qfa_train_workspace_bytes(int Nb, int Nr, int Nh, int B, int precision) {
    if (Nb < 0 || Nr < 0 || Nb + Nr <= 0 || Nh < 1 || Nh > 32 || B < 0) return 0;
    size_t cc = plan_train(Nb, Nr, Nh, B, precision).total;
    if (tc_eligible(Nh, precision)) {
        size_t tc = plan_tc_train(Nb, Nr, Nh, B).total;
        size_t x3 = plan_tc_x3(Nb, Nr, Nh, B, true).total;
        if (x3 > tc) tc = x3;
        return tc > cc ? tc : cc;
    }
    if (tc32_eligible(Nh, precision)) {
        size_t tc = plan_tc32(Nb, Nr, Nh, B).total;
        return tc > cc ? tc : cc;
    }
    return cc;
}

extern "C" size_t qfa_predict_workspace_bytes(int Nb, int Nr, int Nh, int B, int precision) {
    if (Nb < 0 || Nr < 0 || Nb + Nr <= 0 || Nh < 1 || Nh > 32 || B < 0) return 0;
    if (tc_eligible(Nh, precision)) {
        size_t a = plan_tc(Nb, Nr, B, true).total, b = plan_tc_x3(Nb, Nr, Nh, B, false).total;
        return a > b ? a : b;
    }
    if (tc32_eligible(Nh, precision)) return plan_tc32(Nb, Nr, Nh, B, true).total;
    return 256;
}

template <typename T>
static Field<T> make_field(const QfaModel* m, const float* x, const float* err, const float* zabs,
                           const uint8_t* mask) {
    Field<T> f;
    const int P = m->Nb + m->Nr;
    f.x = x; f.err = err; f.zabs = zabs; f.mask = mask;
    f.F = m->params;
    f.Psi = m->params + (size_t)P * m->Nh;
    f.omega = f.Psi + P;
    f.scal = f.omega + m->Nb;
    f.mu = m->mu;
    f.Nb = m->Nb; f.P = P; f.Nh = m->Nh;
    LawConst lc = law_constants(m->tau_law);
    f.lt0 = (T)lc.t0; f.lbe = (T)lc.be; f.lC = (T)lc.C; f.llogzn = (T)log(lc.zn);
    return f;
}

static int check_model(const QfaModel* m, int precision) {
    if (!m || !m->params) return fail(QFA_ERR_NULL, "model / model->params is NULL");
    if (m->Nb < 0 || m->Nr < 0 || m->Nb + m->Nr <= 0) return fail(QFA_ERR_SHAPE, "bad grid Nb=%d Nr=%d", m->Nb, m->Nr);
    if (m->Nh < 1 || m->Nh > 32) return fail(QFA_ERR_NH, "Nh=%d unsupported (1..32)", m->Nh);
    if (m->tau_law < 0 || m->tau_law > 3) return fail(QFA_ERR_LAW, "unknown tau law %d", m->tau_law);
    if (precision != QFA_PREC_FP64 && precision != QFA_PREC_FP32 && precision != QFA_PREC_TF32)
        return fail(QFA_ERR_PRECISION, "unknown precision mode %d", precision);
    if (((uintptr_t)m->params & 15) != 0) return fail(QFA_ERR_ALIGN, "params must be 16-byte aligned");
    return 0;
}

// ---------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------
#ifndef QFA_SMEM_PER_SM_KB
#define QFA_SMEM_PER_SM_KB 227      // usable shared memory per SM (228 KB minus 1 KB reserved per CTA is accounted for below)
#endif
template <typename T, int HP, int MODE>
static cudaError_t launch_gram(const GramArgs<T>& a, cudaStream_t st) {
    using C = GramCfg<T, HP, MODE>;
    auto kern = k_gram_solve<T, HP, MODE>;
    static PerDeviceOnce attr_once;
    if (!attr_once.done[current_device()]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes);
        if (e != cudaSuccess) return e;
        attr_once.done[current_device()] = true;
    }
    int per_sm = (int)((size_t)QFA_SMEM_PER_SM_KB * 1024 / (C::smem_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int grid = num_sms() * per_sm;
    if (grid > a.B) grid = a.B;
    if (grid < 1) return cudaSuccess;
    kern<<<grid, C::NT, C::smem_bytes, st>>>(a); QFA_LAUNCHED();
    return cudaGetLastError();
}

template <typename T, int MODE>
static cudaError_t dispatch_gram(int HP, const GramArgs<T>& a, cudaStream_t st) {
    switch (HP) {
        case 4: return launch_gram<T, 4, MODE>(a, st);
        case 8: return launch_gram<T, 8, MODE>(a, st);
        case 16: return launch_gram<T, 16, MODE>(a, st);
        default: return launch_gram<T, 32, MODE>(a, st);
    }
}

template <typename T>
static cudaError_t dispatch_grad(int HP, const GradArgs<T>& a, int ntiles, cudaStream_t st) {
    dim3 grid(ntiles, a.nsplit);
    switch (HP) {
        case 4: k_grad<T, 4><<<grid, 128, 0, st>>>(a); QFA_LAUNCHED(); break;
        case 8: k_grad<T, 8><<<grid, 128, 0, st>>>(a); QFA_LAUNCHED(); break;
        case 16: k_grad<T, 16><<<grid, 128, 0, st>>>(a); QFA_LAUNCHED(); break;
        default: k_grad<T, 32><<<grid, 128, 0, st>>>(a); QFA_LAUNCHED(); break;
    }
    return cudaGetLastError();
}

template <typename T>
static int train_accumulate_t(const QfaModel* m, const float* delta, const float* error, const float* zabs,
                              const uint8_t* mask, int B, char* ws, const TrainPlan& pl, T* acc, T* nll_out,
                              int flags, cudaStream_t st) {
    const int P = m->Nb + m->Nr, Nb = m->Nb, Nh = m->Nh;
    if (flags & QFA_FLAG_ZERO_ACC) CK(cudaMemsetAsync(acc, 0, qfa_acc_len(m->Nb, m->Nr, Nh) * sizeof(T), st));
    if (B == 0) return 0;
    T* small = reinterpret_cast<T*>(ws + pl.off_small);
    float* hasblue = reinterpret_cast<float*>(ws + pl.off_hasblue);
    T* nll = nll_out ? nll_out : reinterpret_cast<T*>(ws + pl.off_nll);
    T* part = reinterpret_cast<T*>(ws + pl.off_part);
    T* spart = reinterpret_cast<T*>(ws + pl.off_spart);
    int nsplit_used = pl.nsplit;
    for (int s0 = 0, it = 0; s0 < B; s0 += pl.SB, ++it) {
        const int nb = (B - s0 < pl.SB) ? B - s0 : pl.SB;
        Field<T> f = make_field<T>(m, delta + (size_t)s0 * P, error + (size_t)s0 * P,
                                   zabs + (size_t)s0 * Nb, mask + (size_t)s0 * P);
        GramArgs<T> ga;
        ga.f = f; ga.B = nb; ga.small = small; ga.nll = nll + s0; ga.hasblue = hasblue + s0;
        ga.hmean = nullptr; ga.hcov = nullptr; ga.cont = nullptr; ga.unc = nullptr;
        CK((dispatch_gram<T, MODE_TRAIN>(pl.HP, ga, st)));
        GradArgs<T> gr;
        gr.f = f; gr.B = nb; gr.nsplit = nsplit_used; gr.small = small; gr.part = part; gr.spart = spart;
        gr.accumulate = it > 0;
        CK(dispatch_grad<T>(pl.HP, gr, pl.ntiles, st));
    }
    ReduceArgs<T> ra;
    ra.part = part; ra.spart = spart; ra.nll = nll; ra.hasblue = hasblue; ra.scal = make_field<T>(m, 0, 0, 0, 0).scal;
    ra.acc = acc; ra.P = P; ra.Nb = Nb; ra.Nh = Nh; ra.B = B; ra.stride = 1; ra.nsp = (double)B; ra.nsplit = nsplit_used;
    ra.ntiles_blue = pl.ntiles_blue; ra.ntiles = pl.ntiles; ra.nsplit_red = nsplit_used; ra.tile_px = 128;
    size_t n_el = part_len(P, Nb, Nh);
    int blocks = (int)((n_el + 255) / 256);
    k_reduce<T><<<blocks, 256, 0, st>>>(ra); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

static int train_accumulate_tc(const QfaModel* m, const float* delta, const float* error, const float* zabs,
                               const uint8_t* mask, int B, char* ws, const TcTrainPlan& pl, float* acc, float* nll_out,
                               int flags, cudaStream_t st) {
    using namespace tcg;
    const int P = m->Nb + m->Nr, Nb = m->Nb, Nh = m->Nh;
    if (flags & QFA_FLAG_ZERO_ACC) CK(cudaMemsetAsync(acc, 0, qfa_acc_len(m->Nb, m->Nr, Nh) * sizeof(float), st));
    if (B == 0) return 0;
    float* PB = reinterpret_cast<float*>(ws + pl.t.off_pb);
    float* QA = reinterpret_cast<float*>(ws + pl.t.off_qa);
    float* sm_b2 = reinterpret_cast<float*>(ws + pl.off_b2);
    float* sm_kc = reinterpret_cast<float*>(ws + pl.off_kc);
    float* tsums = reinterpret_cast<float*>(ws + pl.off_tsums);
    float* part = reinterpret_cast<float*>(ws + pl.off_part);
    float* spart = reinterpret_cast<float*>(ws + pl.off_spart);
    const size_t n_el = (size_t)pl.t.nkb * PB_ROWS * KB + (size_t)pl.t.npt * 2 * PT * KB;
    int blocks = (int)((n_el + 255) / 256);
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    k_tc_build_images<<<blocks, 256, 0, st>>>(m->params, P, Nh, PB, pl.t.nkb, QA, pl.t.npt); QFA_LAUNCHED();
    CK(cudaGetLastError());
    static PerDeviceOnce attr_once;
    if (!attr_once.done[current_device()]) {
        CK(cudaFuncSetAttribute(k_tc_gram<TC_TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<TC_TRAIN>::SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_tc_grad<GCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, GradSmem::BYTES));
        CK(cudaFuncSetAttribute(k_tc_grad<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GradSmem::BYTES));
        attr_once.done[current_device()] = true;
    }
    Field<float> f = make_field<float>(m, delta, error, zabs, mask);
    TcGramArgs a;
    a.f = f; a.B = B; a.ts = pl.t.ts; a.ntiles = pl.t.ntiles; a.nkb = pl.t.nkb; a.npt = pl.t.npt;
    a.PB = PB; a.QA = QA; a.nll = nll_out; a.hmean = nullptr; a.hcov = nullptr; a.cont = nullptr; a.unc = nullptr;
    a.sm_b2 = sm_b2; a.sm_kc = sm_kc; a.hasblue = nullptr; a.tile_sums = tsums; a.trace = g_trace;
    int grid = pl.t.ntiles < num_sms() ? pl.t.ntiles : num_sms();
    k_tc_gram<TC_TRAIN><<<grid, NTHREADS, Cfg<TC_TRAIN>::SMEM_BYTES, st>>>(a); QFA_LAUNCHED();
    CK(cudaGetLastError());
    TcGradArgs ga;
    ga.f = f; ga.B = B; ga.ts = pl.t.ts; ga.nchunks = pl.nchunks; ga.gct = pl.gct; ga.ns_blue = pl.ns_blue; ga.ns_red = pl.ns_red;
    ga.ntiles_blue = pl.ntiles_blue; ga.npt = pl.t.npt; ga.QA = QA; ga.sm_b2 = sm_b2; ga.sm_kc = sm_kc;
    ga.zero = reinterpret_cast<const uint8_t*>(PB) + sw128_offset_host(36, 0);
    ga.part = part; ga.spart = spart; ga.accumulate = 0; ga.trace = g_trace_grad;
    const int grad_ctas = pl.ntiles_blue * pl.ns_blue + (pl.t.npt - pl.ntiles_blue) * pl.ns_red;
    if (pl.gct == GCT) k_tc_grad<GCT><<<grad_ctas, GRAD_THREADS, GradSmem::BYTES, st>>>(ga);
    else k_tc_grad<0><<<grad_ctas, GRAD_THREADS, GradSmem::BYTES, st>>>(ga);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    ReduceArgs<float> ra;
    ra.part = part; ra.spart = spart; ra.nll = tsums; ra.hasblue = tsums + 1; ra.scal = f.scal;   // pre-folded per 32 rows
    ra.acc = acc; ra.P = P; ra.Nb = Nb; ra.Nh = Nh; ra.B = pl.t.ntiles * 4; ra.stride = 2; ra.nsp = (double)B;
    ra.nsplit = pl.ns_blue; ra.nsplit_red = pl.ns_red; ra.tile_px = PT; ra.ntiles_blue = pl.ntiles_blue; ra.ntiles = pl.t.npt;
    size_t n_pl = part_len(P, Nb, Nh);
    k_reduce<float><<<(int)((n_pl + 255) / 256), 256, 0, st>>>(ra); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

static int train_accumulate_tc32(const QfaModel* m, const float* delta, const float* error, const float* zabs,
                                 const uint8_t* mask, int B, char* ws, const Tc32Plan& pl, float* acc, float* nll_out,
                                 int flags, cudaStream_t st) {
    using namespace tcg32;
    const int P = m->Nb + m->Nr, Nb = m->Nb, Nh = m->Nh;
    if (flags & QFA_FLAG_ZERO_ACC) CK(cudaMemsetAsync(acc, 0, qfa_acc_len(m->Nb, m->Nr, Nh) * sizeof(float), st));
    if (B == 0) return 0;
    float* PB = reinterpret_cast<float*>(ws + pl.off_pb);
    float* gram = reinterpret_cast<float*>(ws + pl.off_gram);
    float* small = reinterpret_cast<float*>(ws + pl.off_small);
    float* hasblue = reinterpret_cast<float*>(ws + pl.off_hasblue);
    float* nll = nll_out ? nll_out : reinterpret_cast<float*>(ws + pl.off_nll);
    float* part = reinterpret_cast<float*>(ws + pl.off_part);
    float* spart = reinterpret_cast<float*>(ws + pl.off_spart);
    const size_t n_el = (size_t)pl.nkb * PB32_ROWS * tcg::KB;
    int blocks = (int)((n_el + 255) / 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    k_tc_build_images32<<<blocks, 256, 0, st>>>(m->params, P, Nh, PB, pl.nkb, 0); QFA_LAUNCHED();
    CK(cudaGetLastError());
    static PerDeviceOnce attr_once;
    if (!attr_once.done[current_device()]) {
        CK(cudaFuncSetAttribute(k_tc_gram32<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM32_BYTES));
        CK(cudaFuncSetAttribute(k_solve32<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SOLVE32_SMEM));
        CK(cudaFuncSetAttribute(k_solve32<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, SOLVE32_SMEM));
        CK(cudaFuncSetAttribute(k_tc_grad32, cudaFuncAttributeMaxDynamicSharedMemorySize, G32_SMEM));
        attr_once.done[current_device()] = true;
    }
    Field<float> f = make_field<float>(m, delta, error, zabs, mask);
    // Gram kernel: the three-pass form.  QFA_GRAM32_CLUSTER=1 selects the 4-CTA cluster form instead (one generation pass, columns
    // split over the cluster, operand tiles broadcast through distributed shared memory: qfa_tc_gram32c.cuh) -- same results,
    // MEASURED SLOWER (1 449 vs 711 us on 65 536 spectra, see the header there), kept as a tested experiment.
    int use_cluster = 0, nclusters = 0;
    {
        const char* e = getenv("QFA_GRAM32_CLUSTER");       // read per call: the parity test switches it inside one process
        if (e && atoi(e)) {
            static PerDeviceOnce once;
            static int maxc[kMaxDevices] = {0};
            if (!once.done[current_device()]) {
                CK(cudaFuncSetAttribute(k_tc_gram32c, cudaFuncAttributeMaxDynamicSharedMemorySize, CL_SMEM));
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(CL_NC * (num_sms() / CL_NC)); cfg.blockDim = dim3(tcg::NTHREADS); cfg.dynamicSmemBytes = CL_SMEM;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                int n = 0;
                if (cudaOccupancyMaxActiveClusters(&n, k_tc_gram32c, &cfg) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
                maxc[current_device()] = n;
                once.done[current_device()] = true;
            }
            nclusters = maxc[current_device()];
            if (nclusters > pl.ntiles) nclusters = pl.ntiles;
            use_cluster = nclusters > 0;
        }
    }
    if (use_cluster) {
        TcGram32cArgs c;
        c.f = f; c.B = B; c.ts = pl.ts; c.ntiles = pl.ntiles; c.nkb = pl.nkb; c.nclusters = nclusters; c.PB = PB; c.gram = gram;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CL_NC * nclusters); cfg.blockDim = dim3(tcg::NTHREADS); cfg.dynamicSmemBytes = CL_SMEM; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL_NC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, k_tc_gram32c, c)); QFA_LAUNCHED();
    } else {
    TcGram32Args a;
    a.f = f; a.B = B; a.ts = pl.ts; a.ntiles = pl.ntiles; a.nkb = pl.nkb; a.PB = PB; a.gram = gram; a.trace = g_trace;
    a.replay = reinterpret_cast<float*>(ws + pl.off_replay);
    int grid = pl.ntiles < num_sms() ? pl.ntiles : num_sms();
    k_tc_gram32<false><<<grid, tcg::NTHREADS, SMEM32_BYTES, st>>>(a); QFA_LAUNCHED();
    CK(cudaGetLastError());
    }
    int sblocks = (B + SOLVE32_WARPS - 1) / SOLVE32_WARPS;
    if (sblocks > QFA_SOLVE32_CTAS * num_sms()) sblocks = QFA_SOLVE32_CTAS * num_sms();
    if (flags & QFA_FLAG_SOLVE_FP64) k_solve32<double><<<sblocks, SOLVE32_WARPS * 32, SOLVE32_SMEM, st>>>(gram, B, small, nll, hasblue, nullptr, nullptr, H32, 1, use_cluster);
    else k_solve32<float><<<sblocks, SOLVE32_WARPS * 32, SOLVE32_SMEM, st>>>(gram, B, small, nll, hasblue, nullptr, nullptr, H32, 1, use_cluster);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    TcGrad32Args gr;
    gr.f = f; gr.B = B; gr.nsplit = pl.nsplit; gr.img = small; gr.part = part; gr.spart = spart; gr.trace = g_trace_grad;
    k_tc_grad32<<<dim3(pl.npix_tiles, pl.nsplit), QFA_G32_CHAIN ? G32_THREADS_CHAIN : G32_THREADS, G32_SMEM, st>>>(gr); QFA_LAUNCHED();
    CK(cudaGetLastError());
    ReduceArgs<float> ra;
    ra.part = part; ra.spart = spart; ra.nll = nll; ra.hasblue = hasblue; ra.scal = f.scal;
    ra.acc = acc; ra.P = P; ra.Nb = Nb; ra.Nh = Nh; ra.B = B; ra.stride = 1; ra.nsp = (double)B; ra.nsplit = pl.nsplit;
    ra.ntiles_blue = pl.ntiles_blue; ra.ntiles = pl.npix_tiles; ra.nsplit_red = pl.nsplit; ra.tile_px = 128;
    size_t n_pl = part_len(P, Nb, Nh);
    k_reduce<float><<<(int)((n_pl + 255) / 256), 256, 0, st>>>(ra); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

static int x3_build_and_gram(const QfaModel* m, const float* x, const float* error, const float* zabs, const uint8_t* mask,
                             int B, char* ws, const TcX3Plan& pl, bool train, bool want_o, tcx::TcGramX3Args& a, cudaStream_t st) {
    using namespace tcx;
    float* PB = reinterpret_cast<float*>(ws + pl.off_pb);
    float* QA = reinterpret_cast<float*>(ws + pl.off_qa);
    const int P = m->Nb + m->Nr;
    const size_t n_el = (size_t)pl.nkb * tcg::PB_ROWS * XKB + (want_o ? (size_t)pl.npt * XQA_BLK * tcg::PT * XKB : 0);
    int blocks = (int)((n_el + 255) / 256);
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    k_tc_build_images_x3<<<blocks, 256, 0, st>>>(m->params, P, m->Nh, PB, pl.nkb, want_o ? QA : nullptr, want_o ? pl.npt : 0);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    static PerDeviceOnce attr_once;
    if (!attr_once.done[current_device()]) {
        CK(cudaFuncSetAttribute(k_tc_gram_x3<tcg::TC_PREDICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, XCfg<tcg::TC_PREDICT>::SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_tc_gram_x3<tcg::TC_TRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, XCfg<tcg::TC_TRAIN>::SMEM_BYTES));
        attr_once.done[current_device()] = true;
    }
    a.f = make_field<float>(m, x, error, zabs, mask);
    a.B = B; a.ts = pl.ts; a.ntiles = pl.ntiles; a.nkb = pl.nkb; a.npt = pl.npt; a.PB = PB; a.QA = QA;
    int grid = pl.ntiles < num_sms() ? pl.ntiles : num_sms();
    if (train) k_tc_gram_x3<tcg::TC_TRAIN><<<grid, tcg::NTHREADS, XCfg<tcg::TC_TRAIN>::SMEM_BYTES, st>>>(a);
    else k_tc_gram_x3<tcg::TC_PREDICT><<<grid, tcg::NTHREADS, XCfg<tcg::TC_PREDICT>::SMEM_BYTES, st>>>(a);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// train step in 3xTF32 mode: Grams + solve on the tensor cores (k_tc_gram_x3<TRAIN>), gradient contraction by the float
// CUDA-core kernel k_grad<float, 8> from the same hand-off record the 'fp32' mode uses
static int train_accumulate_x3(const QfaModel* m, const float* delta, const float* error, const float* zabs,
                               const uint8_t* mask, int B, char* ws, const TcX3Plan& pl, float* acc, float* nll_out,
                               int flags, cudaStream_t st) {
    const int P = m->Nb + m->Nr, Nb = m->Nb, Nh = m->Nh;
    if (flags & QFA_FLAG_ZERO_ACC) CK(cudaMemsetAsync(acc, 0, qfa_acc_len(m->Nb, m->Nr, Nh) * sizeof(float), st));
    if (B == 0) return 0;
    float* small = reinterpret_cast<float*>(ws + pl.off_small);
    float* hasblue = reinterpret_cast<float*>(ws + pl.off_hasblue);
    float* nll = nll_out ? nll_out : reinterpret_cast<float*>(ws + pl.off_nll);
    float* part = reinterpret_cast<float*>(ws + pl.off_part);
    float* spart = reinterpret_cast<float*>(ws + pl.off_spart);
    tcx::TcGramX3Args a;
    a.nll = nll; a.hmean = nullptr; a.hcov = nullptr; a.cont = nullptr; a.unc = nullptr; a.small = small; a.hasblue = hasblue;
    if (int rc = x3_build_and_gram(m, delta, error, zabs, mask, B, ws, pl, true, false, a, st)) return rc;
    GradArgs<float> gr;
    gr.f = a.f; gr.B = B; gr.nsplit = pl.nsplit; gr.small = small; gr.part = part; gr.spart = spart; gr.accumulate = 0;
    CK(dispatch_grad<float>(8, gr, pl.npt, st));
    ReduceArgs<float> ra;
    ra.part = part; ra.spart = spart; ra.nll = nll; ra.hasblue = hasblue; ra.scal = a.f.scal;
    ra.acc = acc; ra.P = P; ra.Nb = Nb; ra.Nh = Nh; ra.B = B; ra.stride = 1; ra.nsp = (double)B; ra.nsplit = pl.nsplit;
    ra.ntiles_blue = pl.ntiles_blue; ra.ntiles = pl.npt; ra.nsplit_red = pl.nsplit; ra.tile_px = 128;
    k_reduce<float><<<(int)((part_len(P, Nb, Nh) + 255) / 256), 256, 0, st>>>(ra); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int qfa_train_accumulate(const QfaModel* model, const float* delta, const float* error,
                                    const float* zabs, const uint8_t* mask, int B, void* workspace,
                                    size_t workspace_bytes, void* acc, void* nll_per_spectrum, int precision,
                                    int flags, void* stream) {
    int rc = check_model(model, precision);
    if (rc) return rc;
    if (B < 0) return fail(QFA_ERR_SHAPE, "B=%d", B);
    if (!acc) return fail(QFA_ERR_NULL, "acc is NULL");
    if (B > 0 && (!delta || !error || !mask || (!zabs && model->Nb > 0)))
        return fail(QFA_ERR_NULL, "delta/error/zabs/mask is NULL");
    TrainPlan pl = plan_train(model->Nb, model->Nr, model->Nh, B, precision);
    if (B > 0 && (!workspace || workspace_bytes < pl.total))
        return fail(QFA_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", pl.total, workspace_bytes);
    if (((uintptr_t)workspace & 255) != 0) return fail(QFA_ERR_ALIGN, "workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if ((flags & QFA_FLAG_TF32X3) && precision == QFA_PREC_TF32) {
        if (model->Nh > tcg::HP) return fail(QFA_ERR_UNSUPPORTED, "QFA_FLAG_TF32X3 is implemented for Nh <= 8 (Nh = %d)", model->Nh);
        TcX3Plan xp = plan_tc_x3(model->Nb, model->Nr, model->Nh, B, true);
        if (B > 0 && workspace_bytes < xp.total)
            return fail(QFA_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", xp.total, workspace_bytes);
        return train_accumulate_x3(model, delta, error, zabs, mask, B, (char*)workspace, xp, (float*)acc,
                                   (float*)nll_per_spectrum, flags, st);
    }
    if (tc_use(model->Nh, precision, B, flags, TC_PATH_TRAIN)) {
        TcTrainPlan tp = plan_tc_train(model->Nb, model->Nr, model->Nh, B);
        if (B > 0 && workspace_bytes < tp.total)
            return fail(QFA_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", tp.total, workspace_bytes);
        return train_accumulate_tc(model, delta, error, zabs, mask, B, (char*)workspace, tp, (float*)acc,
                                   (float*)nll_per_spectrum, flags, st);
    }
    if (tc32_eligible(model->Nh, precision) && (B >= tc_min_batch(TC_PATH_TRAIN32) || (flags & QFA_FLAG_FORCE_TENSOR))) {
        Tc32Plan tp = plan_tc32(model->Nb, model->Nr, model->Nh, B);
        if (B > 0 && workspace_bytes < tp.total)
            return fail(QFA_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", tp.total, workspace_bytes);
        return train_accumulate_tc32(model, delta, error, zabs, mask, B, (char*)workspace, tp, (float*)acc,
                                     (float*)nll_per_spectrum, flags, st);
    }
    if (precision == QFA_PREC_FP64)
        return train_accumulate_t<double>(model, delta, error, zabs, mask, B, (char*)workspace, pl, (double*)acc,
                                          (double*)nll_per_spectrum, flags, st);
    // QFA_PREC_TF32 currently shares the float CUDA-core path
    return train_accumulate_t<float>(model, delta, error, zabs, mask, B, (char*)workspace, pl, (float*)acc,
                                     (float*)nll_per_spectrum, flags, st);
}

template <typename T>
static int predict_t(const QfaModel* m, const float* flux, const float* error, const float* zabs,
                     const uint8_t* mask, int B, T* nll, T* hmean, T* hcov, T* cont, T* unc, cudaStream_t st) {
    GramArgs<T> ga;
    ga.f = make_field<T>(m, flux, error, zabs, mask);
    ga.B = B; ga.small = nullptr; ga.nll = nll; ga.hasblue = nullptr;
    ga.hmean = hmean; ga.hcov = hcov; ga.cont = cont; ga.unc = unc;
    CK((dispatch_gram<T, MODE_PREDICT>(pad_h(m->Nh), ga, st)));
    return 0;
}

static int predict_tc(const QfaModel* m, const float* flux, const float* error, const float* zabs, const uint8_t* mask,
                      int B, char* ws, const TcPlan& pl, float* nll, float* hmean, float* hcov, float* cont, float* unc,
                      bool want_o, cudaStream_t st) {
    using namespace tcg;
    using C = Cfg<TC_PREDICT>;
    float* PB = reinterpret_cast<float*>(ws + pl.off_pb);
    float* QA = reinterpret_cast<float*>(ws + pl.off_qa);
    const int P = m->Nb + m->Nr;
    const size_t n_el = (size_t)pl.nkb * PB_ROWS * KB + (want_o ? (size_t)pl.npt * 2 * PT * KB : 0);
    int blocks = (int)((n_el + 255) / 256);
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    k_tc_build_images<<<blocks, 256, 0, st>>>(m->params, P, m->Nh, PB, pl.nkb, want_o ? QA : nullptr, want_o ? pl.npt : 0); QFA_LAUNCHED();
    CK(cudaGetLastError());
    static PerDeviceOnce attr_once;
    if (!attr_once.done[current_device()]) {
        CK(cudaFuncSetAttribute(k_tc_gram<TC_PREDICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr_once.done[current_device()] = true;
    }
    TcGramArgs a;
    a.f = make_field<float>(m, flux, error, zabs, mask);
    a.B = B; a.ts = pl.ts; a.ntiles = pl.ntiles; a.nkb = pl.nkb; a.npt = pl.npt;
    a.PB = PB; a.QA = QA; a.nll = nll; a.hmean = hmean; a.hcov = hcov; a.cont = cont; a.unc = unc;
    a.sm_b2 = nullptr; a.sm_kc = nullptr; a.hasblue = nullptr; a.tile_sums = nullptr; a.trace = g_trace;
    int grid = pl.ntiles < num_sms() ? pl.ntiles : num_sms();
    k_tc_gram<TC_PREDICT><<<grid, NTHREADS, C::SMEM_BYTES, st>>>(a); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// prediction for 8 < Nh <= 32 on the tensor cores: k_tc_gram32<PRED> (M, b) -> k_solve32<PRED> (NLL, hmean, hcov, [L^-1 | a]
// images) -> k_out32 (continuum, 1-sigma)                                                      (model.py:160-180)
static int predict_tc32(const QfaModel* m, const float* flux, const float* error, const float* zabs, const uint8_t* mask,
                        int B, char* ws, const Tc32Plan& pl, float* nll, float* hmean, float* hcov, float* cont, float* unc,
                        int flags, cudaStream_t st) {
    using namespace tcg32;
    const int P = m->Nb + m->Nr;
    float* PB = reinterpret_cast<float*>(ws + pl.off_pb);
    float* PBlo = reinterpret_cast<float*>(ws + pl.off_pb2);
    float* gram = reinterpret_cast<float*>(ws + pl.off_gram);
    float* img = reinterpret_cast<float*>(ws + pl.off_small);
    const size_t n_el = (size_t)pl.nkb * PB32_ROWS * tcg::KB;
    int blocks = (int)((n_el + 255) / 256);
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    k_tc_build_images32<<<blocks, 256, 0, st>>>(m->params, P, m->Nh, PB, pl.nkb, 0); QFA_LAUNCHED();
    k_tc_build_images32<<<blocks, 256, 0, st>>>(m->params, P, m->Nh, PBlo, pl.nkb, 1); QFA_LAUNCHED();
    CK(cudaGetLastError());
    static PerDeviceOnce attr_once;
    if (!attr_once.done[current_device()]) {
        CK((cudaFuncSetAttribute(k_tc_gram32<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM32_BYTES)));
        CK((cudaFuncSetAttribute(k_tc_gram32<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM32_BYTES)));
        CK((cudaFuncSetAttribute(k_solve32<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SOLVE32_SMEM)));
        CK((cudaFuncSetAttribute(k_solve32<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SOLVE32_SMEM)));
        CK(cudaFuncSetAttribute(k_out32, cudaFuncAttributeMaxDynamicSharedMemorySize, G32_SMEM));
        attr_once.done[current_device()] = true;
    }
    Field<float> f = make_field<float>(m, flux, error, zabs, mask);
    // The per-spectrum outputs of a prediction do not average over a batch, and M = I + F^T W F of an Nh = 32 model is
    // conditioned ~1e3: single-pass TF32 Grams put the continuum 1.6e-2 off (measured, scripts/p32_errors.py), above the 1e-3
    // bar.  So the Grams are 3xTF32: hi*hi + lo*hi + hi*lo as three launches of the same kernel (operand part / image part),
    // each into its own scratch row -- which also keeps the small cross terms out of the big accumulators (the tensor core's
    // fp32 accumulation truncates) -- summed by k_solve32.
    TcGram32Args a;
    a.f = f; a.B = B; a.ts = pl.ts; a.ntiles = pl.ntiles; a.nkb = pl.nkb; a.trace = nullptr;
    a.replay = reinterpret_cast<float*>(ws + pl.off_replay);
    int grid = pl.ntiles < num_sms() ? pl.ntiles : num_sms();
    const size_t gstride = (size_t)B * G32_STRIDE;
    a.PB = PB; a.gram = gram;
    k_tc_gram32<true, 0><<<grid, tcg::NTHREADS, SMEM32_BYTES, st>>>(a); QFA_LAUNCHED();
    a.PB = PB; a.gram = gram + gstride;
    k_tc_gram32<true, 1><<<grid, tcg::NTHREADS, SMEM32_BYTES, st>>>(a); QFA_LAUNCHED();
    a.PB = PBlo; a.gram = gram + 2 * gstride;
    k_tc_gram32<true, 0><<<grid, tcg::NTHREADS, SMEM32_BYTES, st>>>(a); QFA_LAUNCHED();
    CK(cudaGetLastError());
    int sblocks = (B + SOLVE32_WARPS - 1) / SOLVE32_WARPS;
    if (sblocks > QFA_SOLVE32_CTAS * num_sms()) sblocks = QFA_SOLVE32_CTAS * num_sms();
    if (flags & QFA_FLAG_SOLVE_FP64)
        k_solve32<double, true><<<sblocks, SOLVE32_WARPS * 32, SOLVE32_SMEM, st>>>(gram, B, img, nll, nullptr, hmean, hcov, m->Nh, 3);
    else
        k_solve32<float, true><<<sblocks, SOLVE32_WARPS * 32, SOLVE32_SMEM, st>>>(gram, B, img, nll, nullptr, hmean, hcov, m->Nh, 3);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    if (cont || unc) {
        TcOut32Args o;
        o.f = f; o.B = B; o.nsplit = pl.nsplit; o.img = img; o.cont = cont; o.unc = unc;
        k_out32<<<dim3(pl.npix_tiles, pl.nsplit), G32_THREADS, G32_SMEM, st>>>(o); QFA_LAUNCHED();
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" int qfa_predict(const QfaModel* model, const float* flux, const float* error, const float* zabs,
                           const uint8_t* mask, int B, void* workspace, size_t workspace_bytes, void* nll,
                           void* hmean, void* hcov, void* cont, void* unc, int precision, int flags,
                           void* stream) {
    int rc = check_model(model, precision);
    if (rc) return rc;
    if (B < 0) return fail(QFA_ERR_SHAPE, "B=%d", B);
    if (!model->mu) return fail(QFA_ERR_NULL, "model->mu is NULL (prediction needs the mean spectrum)");
    if (B == 0) return 0;
    if (!nll) return fail(QFA_ERR_NULL, "nll is NULL");
    if (!flux || !error || !mask || (!zabs && model->Nb > 0)) return fail(QFA_ERR_NULL, "flux/error/zabs/mask is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if ((flags & QFA_FLAG_TF32X3) && precision == QFA_PREC_TF32) {
        if (model->Nh > tcg::HP) return fail(QFA_ERR_UNSUPPORTED, "QFA_FLAG_TF32X3 is implemented for Nh <= 8 (Nh = %d)", model->Nh);
        TcX3Plan xp = plan_tc_x3(model->Nb, model->Nr, model->Nh, B, false);
        if (!workspace || workspace_bytes < xp.total)
            return fail(QFA_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", xp.total, workspace_bytes);
        if (((uintptr_t)workspace & 255) != 0) return fail(QFA_ERR_ALIGN, "workspace must be 256-byte aligned");
        tcx::TcGramX3Args a;
        a.nll = (float*)nll; a.hmean = (float*)hmean; a.hcov = (float*)hcov; a.cont = (float*)cont; a.unc = (float*)unc;
        a.small = nullptr; a.hasblue = nullptr;
        return x3_build_and_gram(model, flux, error, zabs, mask, B, (char*)workspace, xp, false, cont || unc, a, st);
    }
    if (tc_use(model->Nh, precision, B, flags, TC_PATH_PREDICT)) {
        const bool want_o = cont || unc;
        TcPlan pl = plan_tc(model->Nb, model->Nr, B, true);
        if (!workspace || workspace_bytes < pl.total)
            return fail(QFA_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", pl.total, workspace_bytes);
        if (((uintptr_t)workspace & 255) != 0) return fail(QFA_ERR_ALIGN, "workspace must be 256-byte aligned");
        return predict_tc(model, flux, error, zabs, mask, B, (char*)workspace, pl, (float*)nll, (float*)hmean,
                          (float*)hcov, (float*)cont, (float*)unc, want_o, st);
    }
    // measured cross-overs against the float CUDA-core kernels (scripts/p32_errors.py): ~600 spectra for 16 < Nh <= 32, ~4500 for
    // 8 < Nh <= 16 (whose CUDA-core kernel is the cheaper HP = 16 instance while the tensor path always pads to 32)
    const int p32_min = tc_min_batch(TC_PATH_PREDICT32) * (model->Nh > 16 ? 1 : 7);
    if (tc32_eligible(model->Nh, precision) && (B >= p32_min || (flags & QFA_FLAG_FORCE_TENSOR))) {
        Tc32Plan tp = plan_tc32(model->Nb, model->Nr, model->Nh, B, true);
        if (!workspace || workspace_bytes < tp.total)
            return fail(QFA_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", tp.total, workspace_bytes);
        if (((uintptr_t)workspace & 255) != 0) return fail(QFA_ERR_ALIGN, "workspace must be 256-byte aligned");
        return predict_tc32(model, flux, error, zabs, mask, B, (char*)workspace, tp, (float*)nll, (float*)hmean, (float*)hcov,
                            (float*)cont, (float*)unc, flags, st);
    }
    if (precision == QFA_PREC_FP64)
        return predict_t<double>(model, flux, error, zabs, mask, B, (double*)nll, (double*)hmean, (double*)hcov,
                                 (double*)cont, (double*)unc, st);
    return predict_t<float>(model, flux, error, zabs, mask, B, (float*)nll, (float*)hmean, (float*)hcov,
                            (float*)cont, (float*)unc, st);
}

// ---------------------------------------------------------------------------------------
// parameter-side kernels: finalize, Adam + clip, clip, smooth, batch preparation
// ---------------------------------------------------------------------------------------
struct PLayout {
    size_t PH, P, Nb, n;          // n = param_len
    size_t o_cnt, o_scnt, o_nll, o_nsp;
};
static PLayout playout(int Nb, int Nr, int Nh) {
    PLayout L;
    L.P = (size_t)Nb + Nr; L.Nb = Nb; L.PH = L.P * Nh; L.n = L.PH + L.P + L.Nb + 3;
    L.o_cnt = L.n; L.o_scnt = L.o_cnt + L.P; L.o_nll = L.o_scnt + 3; L.o_nsp = L.o_nll + 1;
    return L;
}

// gradient of packed element e from acc: sum / count (0/0 -> NaN, like model.py:104)
template <typename T>
__device__ __forceinline__ float grad_from_acc(const T* acc, const PLayout& L, size_t e, int Nh) {
    T cnt;
    if (e < L.PH) cnt = acc[L.o_cnt + e / Nh];
    else if (e < L.PH + L.P) cnt = acc[L.o_cnt + (e - L.PH)];
    else if (e < L.PH + L.P + L.Nb) cnt = acc[L.o_cnt + (e - L.PH - L.P)];
    else cnt = acc[L.o_scnt + (e - L.PH - L.P - L.Nb)];
    return (float)(acc[e] / cnt);
}

template <typename T>
__global__ void k_finalize(const T* acc, PLayout L, int Nh, float* grads, float* loss) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < L.n) grads[e] = grad_from_acc(acc, L, e, Nh);
    if (e == 0 && loss) loss[0] = (float)(acc[L.o_nll] / acc[L.o_nsp]);
}

__device__ __forceinline__ float clip_elem(float p, const PLayout& L, size_t e, float lo, float hi) {
    if (e < L.PH) return p;                                                   // F: unbounded
    if (e < L.PH + L.P + L.Nb) return fminf(fmaxf(p, lo), hi);                // Psi, omega (model.py:237-238)
    size_t s = e - (L.PH + L.P + L.Nb);
    if (s == 0) return fminf(fmaxf(p, 0.0f), 1.0f);                           // tau0 (model.py:239)
    if (s == 1) return fminf(fmaxf(p, -5.0f), 5.0f);                          // c0   (model.py:241)
    return fminf(fmaxf(p, 0.1f), 5.0f);                                       // beta (model.py:240)
}

// hyper (device, optional): {lr, bias1, bias2} -- overrides the by-value arguments, so that a captured CUDA graph of the step
// keeps working when the EPOCH counter (and with it the learning rate and the bias corrections, optimizer.py:50-52,98)
// changes.  Thread 0 also folds the step's mean NLL into loss_sum (model.py:213: total_loss += loss / Niter) and advances
// the data cursor of k_gather_prepare: a replayed step needs no host-side argument at all.
template <typename T>
__global__ void k_adam(float* p, float* m, float* v, const T* acc, const float* gin, PLayout L, int Nh,
                       float lr, float b1, float b2, float eps, float wd, float bias1, float bias2,
                       float lo, float hi, const float* hyper, double* loss_sum, double loss_scale,
                       long long* cursor, long long cursor_step) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) {
        if (loss_sum && acc) loss_sum[0] += (double)acc[L.o_nll] / (double)acc[L.o_nsp] * loss_scale;
        if (cursor) cursor[0] += cursor_step;
    }
    if (e >= L.n) return;
    if (hyper) { lr = hyper[0]; bias1 = hyper[1]; bias2 = hyper[2]; }
    float g = gin ? gin[e] : grad_from_acc(acc, L, e, Nh);
    float pe = p[e];
    g = g + wd * pe;                                  // optimizer.py:47
    float me = (1.0f - b1) * g + b1 * m[e];           // optimizer.py:48
    float ve = (1.0f - b2) * g * g + b2 * v[e];       // optimizer.py:49
    m[e] = me; v[e] = ve;
    float mh = me / bias1, vh = ve / bias2;           // optimizer.py:50-51
    pe = pe - lr * mh / (sqrtf(vh) + eps);            // optimizer.py:52
    p[e] = clip_elem(pe, L, e, lo, hi);               // model.py:316 -> 233-241
}

__global__ void k_clip(float* p, PLayout L, float lo, float hi) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < L.n) p[e] = clip_elem(p[e], L, e, lo, hi);
}

// box filter ignoring out-of-range taps; F columns: half-width 15, Psi/omega: half-width 7
__global__ void k_smooth(const float* in, float* out, PLayout L, int Nh) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= L.n) return;
    if (e < L.PH) {
        long i = (long)(e / Nh); int k = (int)(e % Nh);
        long lo = i - 15 < 0 ? 0 : i - 15, hi = i + 15 >= (long)L.P ? (long)L.P - 1 : i + 15;
        float s = 0.f;
        for (long j = lo; j <= hi; ++j) s += in[(size_t)j * Nh + k];
        out[e] = s / (float)(hi - lo + 1);
    } else if (e < L.PH + L.P + L.Nb) {
        size_t base = (e < L.PH + L.P) ? L.PH : L.PH + L.P;
        long n = (e < L.PH + L.P) ? (long)L.P : (long)L.Nb;
        long i = (long)(e - base);
        long lo = i - 7 < 0 ? 0 : i - 7, hi = i + 7 >= n ? n - 1 : i + 7;
        float s = 0.f;
        for (long j = lo; j <= hi; ++j) s += in[base + j];
        out[e] = s / (float)(hi - lo + 1);
    } else {
        out[e] = in[e];
    }
}

static int check_grid(int Nb, int Nr, int Nh) {
    if (Nb < 0 || Nr < 0 || Nb + Nr <= 0) return fail(QFA_ERR_SHAPE, "bad grid Nb=%d Nr=%d", Nb, Nr);
    if (Nh < 1 || Nh > 32) return fail(QFA_ERR_NH, "Nh=%d unsupported (1..32)", Nh);
    return 0;
}

extern "C" int qfa_grads_finalize(const void* acc, int Nb, int Nr, int Nh, int precision, float* grads,
                                  float* loss, void* stream) {
    int rc = check_grid(Nb, Nr, Nh);
    if (rc) return rc;
    if (!acc || !grads) return fail(QFA_ERR_NULL, "acc/grads is NULL");
    PLayout L = playout(Nb, Nr, Nh);
    int blocks = (int)((L.n + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == QFA_PREC_FP64) k_finalize<double><<<blocks, 256, 0, st>>>((const double*)acc, L, Nh, grads, loss);
    else k_finalize<float><<<blocks, 256, 0, st>>>((const float*)acc, L, Nh, grads, loss);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

static int adam_launch(float* params, float* m, float* v, const void* acc, const float* grads_in, int Nb, int Nr, int Nh,
                       int precision, float lr, float b1, float b2, float eps, float weight_decay, float bias1, float bias2,
                       float min_value, float max_value, const float* hyper, double* loss_sum, double loss_scale,
                       long long* cursor, long long cursor_step, void* stream) {
    int rc = check_grid(Nb, Nr, Nh);
    if (rc) return rc;
    if (!params || !m || !v || (!acc && !grads_in)) return fail(QFA_ERR_NULL, "params/m/v/acc is NULL");
    PLayout L = playout(Nb, Nr, Nh);
    int blocks = (int)((L.n + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == QFA_PREC_FP64 && !grads_in)
        k_adam<double><<<blocks, 256, 0, st>>>(params, m, v, (const double*)acc, grads_in, L, Nh, lr, b1, b2, eps,
                                               weight_decay, bias1, bias2, min_value, max_value, hyper, loss_sum, loss_scale,
                                               cursor, cursor_step);
    else
        k_adam<float><<<blocks, 256, 0, st>>>(params, m, v, (const float*)acc, grads_in, L, Nh, lr, b1, b2, eps,
                                              weight_decay, bias1, bias2, min_value, max_value, hyper, loss_sum, loss_scale,
                                              cursor, cursor_step);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int qfa_adam_clip_step(float* params, float* m, float* v, const void* acc, const float* grads_in,
                                  int Nb, int Nr, int Nh, int precision, float lr, float b1, float b2, float eps,
                                  float weight_decay, float bias1, float bias2, float min_value, float max_value,
                                  void* stream) {
    return adam_launch(params, m, v, acc, grads_in, Nb, Nr, Nh, precision, lr, b1, b2, eps, weight_decay, bias1, bias2,
                       min_value, max_value, nullptr, nullptr, 0.0, nullptr, 0, stream);
}

extern "C" int qfa_adam_clip_step_dev(float* params, float* m, float* v, const void* acc, int Nb, int Nr, int Nh,
                                      int precision, const float* hyper_dev, float b1, float b2, float eps,
                                      float weight_decay, float min_value, float max_value, double* loss_sum_dev,
                                      double loss_scale, long long* cursor_dev, long long cursor_step, void* stream) {
    if (!hyper_dev) return fail(QFA_ERR_NULL, "hyper_dev is NULL");
    return adam_launch(params, m, v, acc, nullptr, Nb, Nr, Nh, precision, 0.f, b1, b2, eps, weight_decay, 1.f, 1.f,
                       min_value, max_value, hyper_dev, loss_sum_dev, loss_scale, cursor_dev, cursor_step, stream);
}

// ---- one-shot all-reduce of `acc` over peer-mapped memory (qfa_peer.cuh): the exchange step of the data-parallel train step
extern "C" size_t qfa_peer_buffer_bytes(long long n, int precision, int world) {
    if (n <= 0 || world < 1 || world > peer::kMaxWorld) return 0;
    const size_t elem = precision == QFA_PREC_FP64 ? 8 : 4;
    return peer::kFlagBytes + 2 * peer::pub_bytes((size_t)n, elem);
}

extern "C" int qfa_peer_allreduce(void* acc, long long n, int precision, void* const* peer_base_dev, unsigned int* state_dev,
                                  int world, int rank, void* stream) {
    if (!acc || !peer_base_dev || !state_dev) return fail(QFA_ERR_NULL, "acc/peer_base_dev/state_dev is NULL");
    if (n <= 0) return fail(QFA_ERR_SHAPE, "bad accumulator length %lld", n);
    if (world < 1 || world > peer::kMaxWorld || rank < 0 || rank >= world)
        return fail(QFA_ERR_SHAPE, "bad world=%d rank=%d (1..%d ranks)", world, rank, peer::kMaxWorld);
    if (((uintptr_t)acc & 15u) || ((uintptr_t)peer_base_dev & 7u) || ((uintptr_t)state_dev & 7u))
        return fail(QFA_ERR_ALIGN, "acc must be 16-byte aligned (peer_base_dev, state_dev: 8)");
    const bool f64 = precision == QFA_PREC_FP64;
    const size_t nvec = (size_t)n / (f64 ? 2 : 4);
    int blocks = (int)((nvec + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > num_sms()) blocks = num_sms();              // CTAs spin on the peers' flags: all of them must be resident
    cudaStream_t st = (cudaStream_t)stream;
    char* const* pb = reinterpret_cast<char* const*>(peer_base_dev);
    static long long timeout_s = -1;
    if (timeout_s < 0) { const char* e = getenv("QFA_PEER_TIMEOUT_S"); timeout_s = e && atoll(e) > 0 ? atoll(e) : 600; }
    const unsigned long long tns = (unsigned long long)timeout_s * 1000000000ull;
    if (f64) peer::k_peer_allreduce<double><<<blocks, 256, 0, st>>>((double*)acc, (size_t)n, pb, state_dev, world, rank, tns);
    else peer::k_peer_allreduce<float><<<blocks, 256, 0, st>>>((float*)acc, (size_t)n, pb, state_dev, world, rank, tns);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int qfa_clip(float* params, int Nb, int Nr, int Nh, float min_value, float max_value, void* stream) {
    int rc = check_grid(Nb, Nr, Nh);
    if (rc) return rc;
    if (!params) return fail(QFA_ERR_NULL, "params is NULL");
    PLayout L = playout(Nb, Nr, Nh);
    k_clip<<<(int)((L.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, L, min_value, max_value); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int qfa_smooth(const float* params_in, float* params_out, int Nb, int Nr, int Nh, void* stream) {
    int rc = check_grid(Nb, Nr, Nh);
    if (rc) return rc;
    if (!params_in || !params_out) return fail(QFA_ERR_NULL, "params is NULL");
    if (params_in == params_out) return fail(QFA_ERR_UNSUPPORTED, "qfa_smooth cannot run in place");
    PLayout L = playout(Nb, Nr, Nh);
    k_smooth<<<(int)((L.n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params_in, params_out, L, Nh); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

static aux::Law aux_law(int tau_law) {
    LawConst lc = law_constants(tau_law);
    aux::Law lw; lw.t0 = (float)lc.t0; lw.be = (float)lc.be; lw.C = (float)lc.C; lw.zn = (float)lc.zn;
    return lw;
}

static int gather_prepare(const float* flux, const float* error, const uint8_t* mask, const float* zqso, const float* wav,
                          const float* mu, const long long* perm, const long long* cursor, int B, int Nb, int Nr, int tau_law,
                          float* zabs_out, float* delta_out, float* error_out, uint8_t* mask_out, void* stream) {
    if (Nb < 0 || Nr < 0 || Nb + Nr <= 0 || B < 0) return fail(QFA_ERR_SHAPE, "bad shape");
    if (tau_law < 0 || tau_law > 3) return fail(QFA_ERR_LAW, "unknown tau law %d", tau_law);
    if (!zqso || !wav || (delta_out && (!flux || !mu)) || (error_out && !error) || (mask_out && !mask))
        return fail(QFA_ERR_NULL, "NULL input");
    if (B == 0) return 0;
    aux::PrepArgs a;
    a.flux = flux; a.error = error; a.mask = mask; a.zq = zqso; a.wav = wav; a.mu = mu;
    a.perm = reinterpret_cast<const int64_t*>(perm); a.cursor = reinterpret_cast<const int64_t*>(cursor);
    a.B = B; a.Nb = Nb; a.P = Nb + Nr; a.max_series = aux::kNSeries; a.lw = aux_law(tau_law);
    a.zabs_out = zabs_out; a.delta_out = delta_out; a.error_out = error_out; a.mask_out = mask_out;
    // CTAs of 256 threads, each streaming whole rows.  With >= 4 rows per CTA the kernel builds the separable optical-depth
    // tables in shared memory (3 Nb floats) and runs one resident wave (6 CTAs per SM at 40 registers); smaller batches keep the
    // direct evaluation -- building the tables costs a CTA about as much as one row (batch 500: 18.5 vs 31 us).
    const size_t tab = 3 * (size_t)Nb * sizeof(float);
    const bool table = tab <= 48 * 1024 && (long long)B >= 4LL * 6 * num_sms();
    int grid = (table ? 6 : 8) * num_sms();
    if (grid > B) grid = B;
    if (table) aux::k_gather_prepare<true><<<grid, 256, tab, (cudaStream_t)stream>>>(a);
    else aux::k_gather_prepare<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int qfa_prepare_batch(const float* flux, const float* zqso, const float* wav, const float* mu, int B,
                                 int Nb, int Nr, int tau_law, float* zabs_out, float* delta_out, void* stream) {
    return gather_prepare(flux, nullptr, nullptr, zqso, wav, mu, nullptr, nullptr, B, Nb, Nr, tau_law, zabs_out, delta_out,
                          nullptr, nullptr, stream);
}

extern "C" int qfa_gather_prepare(const float* flux, const float* error, const uint8_t* mask, const float* zqso,
                                  const float* wav, const float* mu, const long long* perm, const long long* cursor_dev,
                                  int B, int Nb, int Nr, int tau_law, float* zabs_out, float* delta_out, float* error_out,
                                  uint8_t* mask_out, void* stream) {
    return gather_prepare(flux, error, mask, zqso, wav, mu, perm, cursor_dev, B, Nb, Nr, tau_law, zabs_out, delta_out,
                          error_out, mask_out, stream);
}

extern "C" int qfa_mean_spectrum_sums(const float* flux, const uint8_t* mask, const float* zqso, const float* wav, int N,
                                      int Nb, int Nr, int tau_law, double* sums, void* stream) {
    if (Nb < 0 || Nr < 0 || Nb + Nr <= 0 || N < 0) return fail(QFA_ERR_SHAPE, "bad shape");
    if (tau_law < 0 || tau_law > 3) return fail(QFA_ERR_LAW, "unknown tau law %d", tau_law);
    if (!flux || !mask || !zqso || !wav || !sums) return fail(QFA_ERR_NULL, "NULL input");
    const int P = Nb + Nr;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(sums, 0, 2 * (size_t)P * sizeof(double), st));
    if (N == 0) return 0;
    const int gx = (P + 127) / 128;
    int gy = (4 * num_sms() + gx - 1) / gx;
    if (gy > N) gy = N;
    aux::k_tau_weight_sums<<<dim3(gx, gy), 128, 0, st>>>(flux, mask, zqso, wav, N, Nb, P, aux::kNSeries, aux_law(tau_law), sums);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int qfa_ood_select(const float* nll, int B, float threshold, int k, int thr_cap, int* count_out, int* thr_idx,
                              int* top_idx, float* top_val, void* stream) {
    if (B < 0 || k < 0 || thr_cap < 0) return fail(QFA_ERR_SHAPE, "B=%d k=%d cap=%d", B, k, thr_cap);
    if (k > aux::kTopKMax) return fail(QFA_ERR_UNSUPPORTED, "k=%d: at most %d", k, aux::kTopKMax);
    if (!nll && B > 0) return fail(QFA_ERR_NULL, "nll is NULL");
    if (k > 0 && !top_idx) return fail(QFA_ERR_NULL, "top_idx is NULL");
    if (!count_out && k == 0) return 0;
    aux::k_ood_select<<<1, 1024, 0, (cudaStream_t)stream>>>(nll, B, threshold, k, thr_cap, count_out, thr_idx, top_idx, top_val);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

extern "C" int qfa_sample_posterior(const QfaModel* model, const float* hmean, const float* hcov, int B, int S,
                                    unsigned long long seed, float* z_out, float* h_out, float* cont_out, void* stream) {
    int rc = check_model(model, QFA_PREC_FP32);
    if (rc) return rc;
    if (B < 0 || S < 0) return fail(QFA_ERR_SHAPE, "B=%d S=%d", B, S);
    if (B == 0 || S == 0) return 0;
    if (!hmean || !hcov) return fail(QFA_ERR_NULL, "hmean/hcov is NULL");
    if (cont_out && !model->mu) return fail(QFA_ERR_NULL, "model->mu is NULL (continuum samples need the mean spectrum)");
    aux::k_sample_posterior<<<B, 128, 0, (cudaStream_t)stream>>>(hmean, hcov, model->params, model->mu, B, model->Nh,
                                                                 model->Nb + model->Nr, S, seed, z_out, h_out, cont_out);
    QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------
// tcgen05 self-test (see qfa_tc_selftest.cuh)
// ---------------------------------------------------------------------------------------
extern "C" int qfa_selftest_umma(const float* A, const float* Bimg_hi, const float* Bimg_lo, float* D, int split,
                                 int* err_flag, void* stream) {
    if (!A || !Bimg_hi || !Bimg_lo || !D || !err_flag) return fail(QFA_ERR_NULL, "NULL argument");
    const int smem = 49152 + 1024;
    CK(cudaFuncSetAttribute(k_selftest_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_selftest_umma<<<1, 128, smem, (cudaStream_t)stream>>>(A, Bimg_hi, Bimg_lo, D, split, err_flag); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// 2-D TMA tile load from a pitched array (see qfa_tc_selftest.cuh).  The tensor-map encoder lives in the driver: it is
// fetched at run time, the library does not link libcuda.
static int encode_tile_map(CUtensorMap* tm, const float* src, int rows, int npix, int pitch_px, int box_w, int box_h) {
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(QFA_ERR_SHAPE, "cuTensorMapEncodeTiled is not available");
    const cuuint64_t gdim[2] = {(cuuint64_t)npix, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch_px * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    // QFA_TMA_L2PROMO = 0 | 64 | 128 | 256 (design experiments: L2 promotion size of the tensor map)
    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (const char* e = getenv("QFA_TMA_L2PROMO")) {
        const int v = atoi(e);
        promo = v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
              : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    }
    const CUresult r = ((EncodeTiled)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)src, gdim, gstride, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(QFA_ERR_SHAPE, "cuTensorMapEncodeTiled failed: %d", (int)r);
    return 0;
}
extern "C" int qfa_selftest_tma2d(const float* src, int rows, int npix, int pitch_px, int x0, int y0, float* out,
                                  int* err_flag, void* stream) {
    if (!src || !out || !err_flag) return fail(QFA_ERR_NULL, "NULL argument");
    if (rows <= 0 || npix <= 0 || pitch_px < npix) return fail(QFA_ERR_SHAPE, "rows=%d npix=%d pitch=%d", rows, npix, pitch_px);
    if ((pitch_px * 4) % 16 != 0 || ((uintptr_t)src & 15) != 0)
        return fail(QFA_ERR_ALIGN, "a tensor map needs a 16-byte aligned base and a row pitch that is a multiple of 16 bytes");
    CUtensorMap tm;
    if (int rc = encode_tile_map(&tm, src, rows, npix, pitch_px, TMA_ST_COLS, TMA_ST_ROWS)) return rc;
    k_selftest_tma2d<<<1, 128, 0, (cudaStream_t)stream>>>(tm, out, x0, y0, err_flag); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}

// Streaming micro-benchmark of the TMA tile path (see k_bench_tma2d): reads src[rows][pitch_px] once, box_w = 32 | 64 | 128.
extern "C" int qfa_bench_tma2d(const float* src, int rows, int npix, int pitch_px, int box_w, float* sink, int* err_flag,
                               void* stream) {
    if (!src || !sink || !err_flag) return fail(QFA_ERR_NULL, "NULL argument");
    if (rows <= 0 || npix <= 0 || pitch_px < npix) return fail(QFA_ERR_SHAPE, "rows=%d npix=%d pitch=%d", rows, npix, pitch_px);
    if ((pitch_px * 4) % 16 != 0 || ((uintptr_t)src & 15) != 0) return fail(QFA_ERR_ALIGN, "pitch / base alignment");
    CUtensorMap tm;
    if (int rc = encode_tile_map(&tm, src, rows, npix, pitch_px, box_w, TMA_ST_ROWS)) return rc;
    const int smem = (box_w >= 128 ? 3 : (box_w == 64 ? 4 : QFA_TMA_NST32)) * TMA_ST_ROWS * box_w * 4 + 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (box_w == 32) {
        CK(cudaFuncSetAttribute(k_bench_tma2d<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_bench_tma2d<32><<<num_sms(), 256, smem, st>>>(tm, rows, npix, sink, err_flag); QFA_LAUNCHED();
    } else if (box_w == 64) {
        CK(cudaFuncSetAttribute(k_bench_tma2d<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_bench_tma2d<64><<<num_sms(), 256, smem, st>>>(tm, rows, npix, sink, err_flag); QFA_LAUNCHED();
    } else if (box_w == 128) {
        CK(cudaFuncSetAttribute(k_bench_tma2d<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_bench_tma2d<128><<<num_sms(), 256, smem, st>>>(tm, rows, npix, sink, err_flag); QFA_LAUNCHED();
    } else return fail(QFA_ERR_SHAPE, "box_w must be 32, 64 or 128");
    CK(cudaGetLastError());
    return 0;
}

// The same stream with the per-thread loader pattern of the production kernels (see k_bench_ldg); any pitch.
extern "C" int qfa_bench_ldg(const float* src, int rows, int npix, int pitch_px, float* sink, void* stream) {
    if (!src || !sink) return fail(QFA_ERR_NULL, "NULL argument");
    if (rows <= 0 || npix <= 0 || pitch_px < npix) return fail(QFA_ERR_SHAPE, "rows=%d npix=%d pitch=%d", rows, npix, pitch_px);
    k_bench_ldg<<<num_sms(), 512, 0, (cudaStream_t)stream>>>(src, rows, npix, pitch_px, sink); QFA_LAUNCHED();
    CK(cudaGetLastError());
    return 0;
}
