/*
 * qfa_b200.h -- C ABI of libqfa_b200.so: the B200-native replacement for the hot
 * path of ZechangSun/QFA (masked low-rank + diagonal Gaussian likelihood, its
 * gradient, and the posterior continuum prediction).
 *
 * The reference has NO FFI (it is pure Python/PyTorch);
 * the drop-in surface is the Python object API of QFA/model.py + QFA/optimizer.py.
 * qfa_b200/model.py and qfa_b200/optimizer.py mirror that API and call the entry
 * points below through ctypes with raw device pointers.  Each entry point names
 * the reference code it replaces (paths relative to the reference repository).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host; buffers are
 *    contiguous, row-major, 4-byte aligned (16-byte for the packed buffers); the
 *    library allocates nothing that outlives a call and keeps no pointers.
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it.
 *  - return value: 0 = ok; negative = argument/shape error found on the host
 *    before any launch (see QFA_ERR_*); positive = cudaError_t.  Text of the last
 *    error of the calling thread: qfa_last_error_string().  No exceptions cross.
 *  - precision: QFA_PREC_FP64 computes everything in double and its T-typed
 *    outputs are double; the other modes compute in float and T = float.
 *  - Nh <= 32 (internally zero-padded to 4/8/16/32, which is exact).
 */
#ifndef QFA_B200_H
#define QFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QFA_ABI_VERSION 2

/* precision modes */
#define QFA_PREC_FP64 0 /* double everywhere; parity mode (<=1e-5 vs fp64-promoted reference) */
#define QFA_PREC_FP32 1 /* float CUDA-core arithmetic (the reference itself is a float32 program) */
#define QFA_PREC_TF32 2 /* "mixed": float everywhere; above a per-path batch size (see QFA_FLAG_FORCE_TENSOR) the contractions (weighted Grams,
                           continuum/sigma, gradient) run on the tensor cores (tcgen05 kind::tf32, operands rounded to
                           TF32, fp32 accumulation in TMEM): Nh <= 8 train + predict, 8 < Nh <= 32 train */

/* mean optical depth laws, reference QFA/utils.py:95-141,149-171 (series = 1) */
#define QFA_TAU_BECKER 0
#define QFA_TAU_FG 1
#define QFA_TAU_KAMBLE 2
#define QFA_TAU_MOCK 3

/* flags */
#define QFA_FLAG_ZERO_ACC 1     /* qfa_train_accumulate: clear `acc` before accumulating */
#define QFA_FLAG_FORCE_TENSOR 2 /* QFA_PREC_TF32: use the tcgen05 kernels even for batches smaller than
                                   the path's cross-over (predict 1280 / 640 for Nh > 16, train 800, train with
                                   8 < Nh <= 32: 192 spectra; env QFA_TC_MIN_BATCH overrides) */
#define QFA_FLAG_SOLVE_FP64 4   /* QFA_PREC_TF32, 8 < Nh <= 32: per-spectrum Cholesky in double instead of float (7 % slower
                                   train step; no measurable difference on any parity case, kept for ill-conditioned models) */

#define QFA_FLAG_TF32X3 8       /* QFA_PREC_TF32 tensor-core kernels: 3xTF32 operand splitting (hi/lo parts of both operands,
                                   three MMAs per product, fp32 accumulation): operand rounding error ~2^-21 instead of
                                   2^-11, i.e. float-level results from the tensor cores at ~1/3 of the tensor rate */

/* error codes (negative) */
#define QFA_ERR_NULL -1
#define QFA_ERR_SHAPE -2
#define QFA_ERR_NH -3
#define QFA_ERR_PRECISION -4
#define QFA_ERR_WORKSPACE -5
#define QFA_ERR_ALIGN -6
#define QFA_ERR_LAW -7
#define QFA_ERR_UNSUPPORTED -8

/*
 * Model view.  `params` is the packed float32 parameter buffer
 *     [ F (Npix*Nh, row-major) | Psi (Npix) | omega (Nb) | tau0 | c0 | beta ]
 * i.e. the six tensors of reference QFA.parameters (model.py:297-306) laid out
 * back to back so that the optimiser update is one launch; Npix = Nb + Nr.
 * `mu` (Npix floats) is only read by qfa_predict (reference model.py:166,180).
 */
typedef struct QfaModel {
    int32_t Nb, Nr, Nh;
    int32_t tau_law;
    const float* params;
    const float* mu;
} QfaModel;

int qfa_abi_version(void);
const char* qfa_last_error_string(void);

/* number of elements of the packed parameter buffer: Npix*Nh + Npix + Nb + 3 (reference model.py:42) */
size_t qfa_param_len(int Nb, int Nr, int Nh);

/*
 * Accumulation buffer ("acc", element type T), the ONLY thing that has to be
 * all-reduced (sum) between GPUs in data-parallel training:
 *   [ sumF (Npix*Nh) | sumPsi (Npix) | sumOmega (Nb) | sum tau0,c0,beta (3) |
 *     pixel counts (Npix) | scalar counts tau0,c0,beta (3) | sum NLL (1) | #spectra (1) |
 *     sum dNLL/dmu (Npix)  (extra; not used by the reference-parity gradient) ]
 */
size_t qfa_acc_len(int Nb, int Nr, int Nh);
size_t qfa_train_workspace_bytes(int Nb, int Nr, int Nh, int B, int precision);
size_t qfa_predict_workspace_bytes(int Nb, int Nr, int Nh, int B, int precision);

/*
 * Replaces the per-spectrum loop of QFA.forward (model.py:98-103) and
 * QFA.loglikelihood_and_gradient_for_single_spectra (model.py:107-158, with
 * utils.py:12-54 MatrixInverse/MatrixLogDet and utils.py:57-92,95-171):
 * adds, for the B spectra given, the per-spectrum partials and the non-zero
 * counts into `acc`.  delta,error: (B,Npix) float; zabs: (B,Nb) float;
 * mask: (B,Npix) bytes (torch.bool), non-zero = pixel is used.
 * nll_per_spectrum: T[B] or NULL.
 */
int qfa_train_accumulate(const QfaModel* model, const float* delta, const float* error,
                         const float* zabs, const uint8_t* mask, int B,
                         void* workspace, size_t workspace_bytes,
                         void* acc, void* nll_per_spectrum,
                         int precision, int flags, void* stream);

/*
 * Replaces the tail of QFA.forward (model.py:100,104): loss = sum NLL / #spectra,
 * grad[e] = sum[e] / count[e] (0/0 -> NaN exactly like the reference).
 * grads: float[qfa_param_len] in the packed parameter layout; loss: float[1].
 */
int qfa_grads_finalize(const void* acc, int Nb, int Nr, int Nh, int precision,
                       float* grads, float* loss, void* stream);

/*
 * Replaces QFA.prediction_for_single_spectra (model.py:160-180) for a batch.
 * flux,error: (B,Npix) float; zabs (B,Nb); mask (B,Npix) bytes.
 * Outputs (element type T; any of hmean/hcov/cont/unc may be NULL = not wanted,
 * nll-only call = likelihood / out-of-distribution scoring):
 *   nll (B) NEGATIVE log-likelihood, hmean (B,Nh), hcov (B,Nh,Nh),
 *   cont (B,Npix) = mu + F hmean, unc (B,Npix) = sqrt(diag(F hcov F^T)).
 */
int qfa_predict(const QfaModel* model, const float* flux, const float* error,
                const float* zabs, const uint8_t* mask, int B,
                void* workspace, size_t workspace_bytes,
                void* nll, void* hmean, void* hcov, void* cont, void* unc,
                int precision, int flags, void* stream);

/*
 * Replaces Adam.update (optimizer.py:47-52) followed by the clipping setter
 * QFA.parameters (model.py:308-316, 233-241), fused on the packed buffers and
 * reading the gradient straight from `acc` (sum/count), so that
 * forward -> all-reduce(acc) -> update needs no intermediate tensor.
 * params, m, v: float[qfa_param_len], updated in place.
 * bias1 = 1 - b1^(i+1), bias2 = 1 - b2^(i+1), lr = scheduled lr (optimizer.py:98),
 * all evaluated on the host from the EPOCH counter i like the reference.
 * If grads_in != NULL it is used instead of acc (float[qfa_param_len]).
 */
int qfa_adam_clip_step(float* params, float* m, float* v,
                       const void* acc, const float* grads_in,
                       int Nb, int Nr, int Nh, int precision,
                       float lr, float b1, float b2, float eps, float weight_decay,
                       float bias1, float bias2,
                       float min_value, float max_value, void* stream);

/* Replaces QFA.clip (model.py:233-241) on the packed parameter buffer. */
int qfa_clip(float* params, int Nb, int Nr, int Nh, float min_value, float max_value, void* stream);

/*
 * Replaces QFA.smooth (model.py:243-252): box filters of width 15 (omega, Psi) and
 * 31 (each column of F) that ignore out-of-range taps (count_include_pad=False).
 * params_out must not alias params_in.
 */
int qfa_smooth(const float* params_in, float* params_out, int Nb, int Nr, int Nh, void* stream);

/*
 * Device-side data preparation (reference QFA/dataloader.py:102,135-136 with the multi-series
 * optical depth tau_total of utils.py:174-203):
 *   zabs[b,i]  = (1+zqso[b]) * wav[i] / 1215.67 - 1                  (i < Nb)
 *   delta[b,i] = flux[b,i] - mu[i] * exp(-tau_total(zqso[b], wav[i])) (A = 1 for i >= Nb)
 * wav: float[Npix] rest-frame grid; either output may be NULL.
 */
int qfa_prepare_batch(const float* flux, const float* zqso, const float* wav, const float* mu,
                      int B, int Nb, int Nr, int tau_law,
                      float* zabs_out, float* delta_out, void* stream);

/*
 * ---- rows "next" of SURVEY.md section 8f: the callers and data formats either side of the hot path ----
 */

/*
 * Same as qfa_adam_clip_step with the gradient read from `acc`, but the EPOCH-dependent scalars come from DEVICE
 * memory: hyper_dev = {lr_i, 1 - b1^(i+1), 1 - b2^(i+1)} (optimizer.py:50-52,98).  A CUDA graph that captured the
 * train step (gather/prepare -> accumulate -> all-reduce -> update) therefore stays valid across epochs.
 * Optional, all device pointers: loss_sum_dev[0] += (sum NLL / #spectra) * loss_scale  (model.py:213 with
 * loss_scale = 1 / Niter); cursor_dev[0] += cursor_step (the batch cursor of qfa_gather_prepare).
 */
int qfa_adam_clip_step_dev(float* params, float* m, float* v, const void* acc,
                           int Nb, int Nr, int Nh, int precision, const float* hyper_dev,
                           float b1, float b2, float eps, float weight_decay,
                           float min_value, float max_value,
                           double* loss_sum_dev, double loss_scale,
                           long long* cursor_dev, long long cursor_step, void* stream);

/*
 * qfa_prepare_batch for a SHUFFLED batch of a device-resident data set (reference QFA/dataloader.py:124-138 after
 * rewind(), dataloader.py:154-167): batch row b is data-set row perm[cursor + b] (perm == NULL: row cursor + b;
 * cursor_dev == NULL: 0; both are int64 DEVICE memory so that a captured graph needs no host argument).
 * Writes the four tensors QFA.forward takes: delta (B,Npix), error (B,Npix), zabs (B,Nb), mask (B,Npix bytes);
 * any output may be NULL.  delta uses the TOTAL Lyman-series optical depth (utils.py:174-203, all 30 lines of
 * QFA/Lyman_series.csv; on grids redward of Ly-beta this is Ly-alpha only).
 */
int qfa_gather_prepare(const float* flux, const float* error, const uint8_t* mask, const float* zqso,
                       const float* wav, const float* mu, const long long* perm, const long long* cursor_dev,
                       int B, int Nb, int Nr, int tau_law,
                       float* zabs_out, float* delta_out, float* error_out, uint8_t* mask_out, void* stream);

/*
 * Column sums of the mean spectrum (dataloader.py:110-111), N spectra resident on the device:
 *   sums[i]        = sum_b flux[b,i] * exp(+tau_total(b,i)) * mask[b,i]      (tau = 0 for i >= Nb)
 *   sums[Npix + i] = #{ b : flux[b,i] != -999 }
 * double[2*Npix]; all-reduce them across ranks, divide, then smooth on the host (utils.py:206-219).
 */
int qfa_mean_spectrum_sums(const float* flux, const uint8_t* mask, const float* zqso, const float* wav,
                           int N, int Nb, int Nr, int tau_law, double* sums, void* stream);

/*
 * Out-of-distribution scoring on per-spectrum NLLs (the consumer of qfa_predict's nll-only mode; BASELINE config 3):
 *   count_out[0] = #{ b : nll[b] > threshold } (NaN counts), the first min(count, thr_cap) hits go to thr_idx
 *   (unordered);  top_idx/top_val[0..k) = the k largest NLLs, descending, ties by ascending index (k <= 2048).
 * count_out == NULL skips the threshold pass; k == 0 skips the top-k.  All pointers are device memory.
 */
int qfa_ood_select(const float* nll, int B, float threshold, int k, int thr_cap,
                   int* count_out, int* thr_idx, int* top_idx, float* top_val, void* stream);

/*
 * Posterior samples of the latent vector and of the continuum (nb/predict.ipynb cell 11:
 * np.random.multivariate_normal(hmean, hcov); mu + F @ hsample):  h = hmean + chol(hcov) z, z ~ N(0, I) from
 * Philox4x32-10 keyed by (seed, spectrum, sample, component) -- reproducible for a given seed.
 * hmean (B,Nh), hcov (B,Nh,Nh) float as qfa_predict returns them.  Outputs (any may be NULL):
 * z_out, h_out (B,S,Nh); cont_out (B,S,Npix).
 */
int qfa_sample_posterior(const QfaModel* model, const float* hmean, const float* hcov, int B, int S,
                         unsigned long long seed, float* z_out, float* h_out, float* cont_out, void* stream);

/*
 * One-shot all-reduce(sum) of the accumulator over PEER-MAPPED device memory -- the exchange step of the data-parallel
 * train step (the reference sums over spectra at model.py:98-103; 8(f) row 4), in place of ncclAllReduce
 * for this 80-140 KB, latency-bound message.  Every rank calls it once per step, in the same order, with
 *   acc            n elements (float; double if precision == QFA_PREC_FP64), 16-byte aligned, summed IN PLACE
 *   peer_base_dev  DEVICE array of `world` pointers: entry q = the peer buffer of rank q as mapped in THIS process
 *                  (qfa_peer_buffer_bytes(n, precision, world) bytes each, zero-filled on every rank before the first call;
 *                  e.g. one torch.distributed._symmetric_memory allocation and its buffer_ptrs_dev)
 *   state_dev      two zero-initialised 32-bit words of LOCAL device memory (step counter, CTA ticket)
 * One kernel: publish acc in the own peer buffer, raise a flag in every peer's, wait for every peer's flag, sum the
 * world's buffers out of peer memory in rank order (the same bits on every rank).  No host argument changes between
 * steps, so the call can sit inside a captured CUDA graph.  A peer that has not arrived after QFA_PEER_TIMEOUT_S seconds
 * (environment, default 600) traps the kernel.
 */
size_t qfa_peer_buffer_bytes(long long n, int precision, int world);
int qfa_peer_allreduce(void* acc, long long n, int precision, void* const* peer_base_dev,
                       unsigned int* state_dev, int world, int rank, void* stream);

/* Number of kernel launches this library has issued in this process (what bench.py reports as gpu_launches). */
unsigned long long qfa_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QFA_B200_H */
