/* TEST INFRASTRUCTURE (oracle) — not part of the product path.
 *
 * Stand-in for the third-party kissfft header that the reference's speedy.c
 * includes under -DKISS_FFT (/root/reference/speedy.c:39-41).  kissfft is not
 * vendored by the reference (Makefile:9,73 clone it from GitHub) and is absent
 * from this environment, so this header declares only the call surface the
 * reference uses (speedy.c:144-146, 223-226, 269, 308-309, 434-449;
 * kiss_fft_test.cc:50-85) and oracle/fft_oracle.c implements it with our own
 * mixed-radix FFT.  A DFT is mathematically pinned, so any correct float FFT
 * behind this header reproduces the reference's known-answer values to ~1e-6.
 */
#ifndef SPEEDY_B200_ORACLE_KISS_FFT_H_
#define SPEEDY_B200_ORACLE_KISS_FFT_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  float r;
  float i;
} kiss_fft_cpx;

struct kiss_fft_state;
typedef struct kiss_fft_state* kiss_fft_cfg;

/* One malloc() block: the reference releases the plan with free()
 * (speedy.c:308). */
kiss_fft_cfg kiss_fft_alloc(int nfft, int inverse_fft, void* mem,
                            size_t* lenmem);
void kiss_fft(kiss_fft_cfg cfg, const kiss_fft_cpx* fin, kiss_fft_cpx* fout);
void kiss_fft_cleanup(void);
#define kiss_fft_free free

#ifdef __cplusplus
}
#endif
#endif
