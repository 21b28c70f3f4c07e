/*
 * oracle/fft_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Single-precision real FFT used by the CPU restatement of zita-convolver
 * (oracle/zita_oracle.c).  It stands in for the two FFTW3f calls the
 * un-vendored library makes (fftwf_execute_dft_r2c / fftwf_execute_dft_c2r,
 * unnormalised, forward sign -1) -- FFTW is not installed in this image.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call this.
 */
#ifndef FOLVE_ORACLE_FFT_H
#define FOLVE_ORACLE_FFT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct offt_plan offt_plan;

/* Plan for real transforms of length nreal = 2^k, 8 <= nreal <= 2^20. */
offt_plan *offt_plan_create(int nreal);
void offt_plan_destroy(offt_plan *p);

/* r2c: in[nreal] real -> out[nreal/2+1] interleaved (re,im); unnormalised,
 * X[k] = sum_n x[n] exp(-2 pi i k n / nreal).  `in` is not modified. */
void offt_r2c(offt_plan *p, const float *in, float *out);

/* c2r: in[nreal/2+1] interleaved complex -> out[nreal] real; unnormalised
 * (c2r(r2c(x)) == nreal * x); imaginary parts of DC and Nyquist are ignored,
 * as FFTW does.  `in` is not modified. */
void offt_c2r(offt_plan *p, const float *in, float *out);

#ifdef __cplusplus
}
#endif
#endif
