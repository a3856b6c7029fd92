/* TEST INFRASTRUCTURE (oracle) — not part of the product path.
 *
 * Stand-in for FFTW3's header, which the reference's speedy.c includes in its
 * shipped (non-KISS_FFT) build (/root/reference/speedy.c:42, Makefile:13,78).
 * FFTW is absent from this environment.  Only the call surface speedy.c uses is
 * declared (speedy.c:148-150, 228-231, 274-277, 311-313, 462-470);
 * oracle/fft_oracle.c implements it with our own double-precision mixed-radix
 * FFT.  As in the real fftw3.h, fftw_complex is the C99 `double complex` type
 * when <complex.h> was included first (speedy.c:32 does so).
 */
#ifndef SPEEDY_B200_ORACLE_FFTW3_H_
#define SPEEDY_B200_ORACLE_FFTW3_H_

#include <complex.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double complex fftw_complex;
struct fftw_plan_s;
typedef struct fftw_plan_s* fftw_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)

void* fftw_malloc(size_t n);
void fftw_free(void* p);
fftw_plan fftw_plan_dft_1d(int n, fftw_complex* in, fftw_complex* out,
                           int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

#ifdef __cplusplus
}
#endif
#endif
