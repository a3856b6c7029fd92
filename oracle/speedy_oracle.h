/* TEST INFRASTRUCTURE (oracle) — not part of the product path.
 * Interface of the CPU restatement in speedy_oracle.c; see that file. */
#ifndef SPEEDY_B200_ORACLE_H_
#define SPEEDY_B200_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_FEATURES 15 /* speedy.c:124 */

typedef struct {
  int rate;         /* samples per second */
  int channels;     /* interleaved channels in the input */
  int match_matlab; /* 1: Future=8/Past=12 (-DMATCH_MATLAB), 0: 12/8 */
  int fft_double;   /* 1: FFTW-style double FFT, 0: kissfft-style float */
  float speed;      /* global speed R_g (sonicSetSpeed) */
  float nonlinear;  /* sonicEnableNonlinearSpeedup factor */
  float feedback;   /* sonicSetDurationFeedbackStrength */
} oracle_cfg;

typedef struct {
  int window, fft, step, partial, future, past;
  int min_period, max_period, max_required, skip;
} oracle_geom;

/* Optional per-frame outputs; any pointer may be NULL.  `max_frames` is the row
 * capacity of every array. */
typedef struct {
  int max_frames;
  int n_analysis;     /* out: spectrogram frames (at_time 1..n) */
  int n_tension;      /* out: tension/speed values (r = 0..n-1) */
  float* spectrogram; /* [n_analysis][fft]   row k = window k = at_time k+1 */
  float* energy;      /* [n_analysis]        frame energy of window k */
  float* normalized;  /* [n_tension][fft/2]  normalised spectrum used at r */
  float* features;    /* [n_tension][15]     speedy.c:106-123 at r */
  float* tension;     /* [n_tension] */
  float* speed;       /* [n_tension]         value passed to sonicIntSetSpeed */
} oracle_taps;

void oracle_geometry(int rate, int match_matlab, oracle_geom* g);
int oracle_frames_analyzed(const oracle_geom* g, long total_frames);
int oracle_tensions_ready(const oracle_geom* g, int analysis_frames);
void oracle_hamming(int window, float* w);

/* Analysis only (K1-K3): fills the taps.  0 on success. */
int oracle_analyze(const oracle_cfg* cfg, const short* in, long n_frames,
                   oracle_taps* taps);

/* Resynthesis only (K4): feed `n_speeds` 10 ms buffers at the given speeds,
 * optionally flush.  Returns the number of output sample frames produced (may
 * exceed out_cap; only out_cap are stored). */
long oracle_resynthesize(const oracle_cfg* cfg, const short* in, long n_frames,
                         const float* speeds, int n_speeds, int flush,
                         short* out, long out_cap);

/* Whole stream: write everything, flush, read everything.  If speed_override
 * is non-NULL it replaces the computed per-frame speeds. */
long oracle_process(const oracle_cfg* cfg, const short* in, long n_frames,
                    const float* speed_override, short* out, long out_cap,
                    oracle_taps* taps);

/* Many independent streams over `n_threads` OS threads (the CPU baseline).
 * in: [n_streams][n_frames][channels], out: [n_streams][out_cap][channels]. */
int oracle_process_batch(const oracle_cfg* cfg, const short* in, long n_frames,
                         int n_streams, short* out, long out_cap,
                         long* out_counts, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
