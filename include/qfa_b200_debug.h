/*
 * qfa_b200_debug.h -- test / design / profiling entry points of libqfa_b200.so.  NOT part of the reference-facing
 * surface (that is include/qfa_b200.h): hardware self-tests of the tcgen05 / TMA plumbing, streaming micro-benchmarks
 * used for design decisions, and the clock64 trace hooks (functional only in a -DQFA_ENABLE_TRACE build; the
 * production library keeps no pointer between calls and the setters return QFA_ERR_UNSUPPORTED).
 */
#ifndef QFA_B200_DEBUG_H
#define QFA_B200_DEBUG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Hardware self-test of the tcgen05 / TMEM / bulk-copy plumbing used by the QFA_PREC_TF32 kernels:
 * D[128][64] = [ A[128][32] * B[48][32]^T | A * B[32:48]^T ] with B given as the swizzled
 * shared-memory image (hi and lo parts).  split != 0 selects the 3xTF32 product.
 * err_flag (device int) is set to 1 if an mbarrier wait timed out.  Test-only entry point.
 */
int qfa_selftest_umma(const float* A, const float* Bimg_hi, const float* Bimg_lo, float* D, int split,
                      int* err_flag, void* stream);

/*
 * Self-test of a 2-D TMA tile load (cp.async.bulk.tensor.2d through a tensor map encoded at run time) from a pitched
 * row-major float array src[rows][pitch_px] with npix valid pixels per row: out[120][32] = the box starting at row y0,
 * pixel x0 (out-of-range elements = 0).  pitch_px * 4 must be a multiple of 16 (QFA_ERR_ALIGN otherwise: this is exactly
 * why the dense reference layout with odd Npix cannot use TMA).  Test-only entry point.
 */
int qfa_selftest_tma2d(const float* src, int rows, int npix, int pitch_px, int x0, int y0, float* out, int* err_flag,
                       void* stream);

/* Streaming micro-benchmark of the same TMA tile path: one persistent kernel reads src[rows][pitch_px] once through 4-stage
 * rings of 120 x box_w boxes (box_w = 32, 64 or 128).  The caller times it.  Test / design aid. */
int qfa_bench_tma2d(const float* src, int rows, int npix, int pitch_px, int box_w, float* sink, int* err_flag, void* stream);

/* ... and through the per-thread loader pattern of the production kernels (15 warps x 8 rows, 128-byte row segments, two
 * register buffers), any pitch.  Test / design aid. */
int qfa_bench_ldg(const float* src, int rows, int npix, int pitch_px, float* sink, void* stream);

/*
 * Debug/profiling aid: `device_buffer` (long long[nkb * 16 * 4], or NULL to switch off) receives clock64 stamps of
 * the first tile of CTA 0 of every following k_tc_gram launch: per K-block and warp {enter, stage free, operands
 * written, done}.  Not part of the reference-facing surface.
 */
int qfa_debug_set_trace(void* device_buffer);
/* same for k_tc_grad: CTA (0,0), long long[nchunks_of_that_cta * 16 * 8]: per chunk and warp 8 stamps */
int qfa_debug_set_trace_grad(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* QFA_B200_DEBUG_H */
