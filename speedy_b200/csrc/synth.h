/* Deterministic, integer-only "speech-shaped" PCM generator.
 *
 * Every sample is a pure function of (stream id, sample rate, channel, sample
 * index): no floating point, no state, no libm, so the CUDA generator used by
 * bench.py and the C generator used by the CPU oracle/baseline produce
 * bit-identical streams and any stream can be regenerated anywhere.
 *
 * Time is cut into 200 ms segments (a ~5 Hz syllable rate).  A per-segment hash
 * picks the segment kind: 55 % voiced (8 harmonics with 1/h roll-off on an f0 of
 * 90-250 Hz, i.e. inside Sonic's 65-400 Hz pitch search range), 25 % unvoiced
 * (hash noise at about a quarter of the voiced amplitude), 20 % silence (+-1 LSB
 * dither).  A triangular envelope takes every segment to zero at its edges, so
 * the frame energy, the spectral difference and hence the tension all move the
 * way they do on speech.  Peak amplitude is about 12000.  Stereo channels are
 * (0.9 x, 1.1 x), in the spirit of /root/reference/sonic_test.cc:900-907.
 */
#ifndef SPEEDY_B200_SYNTH_H_
#define SPEEDY_B200_SYNTH_H_

#include <stdint.h>

#ifdef __CUDACC__
#define SYNTH_FN __host__ __device__ __forceinline__
#else
#define SYNTH_FN static inline
#endif

/* splitmix64 finaliser */
SYNTH_FN uint64_t synth_mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

/* Parabolic sine: 32-bit phase (2^32 = one cycle) -> [-32767, 32767]. */
SYNTH_FN int32_t synth_sin(uint32_t phase) {
  uint32_t half = (phase >> 16) & 0x7FFFu;
  int32_t par = (int32_t)((half * (32768u - half)) >> 13);
  if (par > 32767) par = 32767;
  return (phase & 0x80000000u) ? -par : par;
}

/* Mono sample `n` of stream `id` at `rate` Hz. */
SYNTH_FN int32_t synth_mono(uint64_t id, int32_t rate, int64_t n) {
  const int32_t seg_len = rate / 5;
  const int64_t seg = n / seg_len;
  const int32_t pos = (int32_t)(n - seg * seg_len);
  const uint64_t h =
      synth_mix64((0x5EEDC0DEULL ^ id) * 0x100000001B3ULL + (uint64_t)seg);
  const uint32_t kind = (uint32_t)(h & 0xFFu);
  /* triangular envelope in Q15 */
  int32_t tri = pos < seg_len / 2 ? pos : seg_len - 1 - pos;
  int32_t env = (int32_t)(((int64_t)tri * 65536) / seg_len);
  if (env > 32768) env = 32768;
  if (kind < 141u) {
    const uint32_t f0 = 90u + (uint32_t)((h >> 8) % 161u);
    const uint32_t step = (uint32_t)((((uint64_t)f0) << 32) / (uint32_t)rate);
    const int32_t amp = 2200 + (int32_t)((h >> 24) & 0x7FFu);
    const uint32_t base = step * (uint32_t)pos;
    int32_t acc = 0;
    for (uint32_t k = 1; k <= 8; k++) {
      acc += synth_sin(base * k + (uint32_t)(h >> 32) * k) / (int32_t)k;
    }
    return (((acc * amp) >> 15) * env) >> 15;
  }
  const uint64_t hs = synth_mix64(h ^ ((uint64_t)n * 0xD6E8FEB86659FD93ULL));
  if (kind < 205u) {
    const int32_t noise = (int32_t)(hs & 0xFFFFu) - 32768;
    return (((noise * 3000) >> 15) * env) >> 15;
  }
  return (hs & 1u) ? 1 : -1;
}

/* Interleaved sample: channel c of sample frame n. */
SYNTH_FN int16_t synth_sample(uint64_t id, int32_t rate, int32_t channels,
                              int32_t c, int64_t n) {
  int32_t x = synth_mono(id, rate, n);
  if (channels > 1) x = (c & 1) ? (x * 11) / 10 : (x * 9) / 10;
  if (x > 32767) x = 32767;
  if (x < -32768) x = -32768;
  return (int16_t)x;
}

#endif
