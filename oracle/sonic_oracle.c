/* TEST INFRASTRUCTURE (oracle) — not part of the product path.
 *
 * PARITY UNPINNED (value level).  This file restates the time-scale modifier of
 * upstream Sonic (github.com/waywardgeek/sonic, sonic.c), the third-party
 * library the reference calls for pitch-synchronous overlap-add
 * (/root/reference/soniclib.c:94,98,106,144,145,174,182,354,369,398,464,521,
 * 526,547,551).  The reference neither vendors nor pins it (Makefile:7,17-18,74
 * clone HEAD into ../sonic) and it cannot be fetched here, so the algorithm is
 * restated from its published description:
 *
 *   - the "classic" PICOLA bookkeeping with `remainingInputToCopy`, i.e. the
 *     Sonic the reference's tests were written against: sonic_test.cc:1019-1039
 *     documents 5 of 10 rapidly-varying-speed cases as failing, which is the
 *     behaviour of this bookkeeping (later upstream revisions replaced it with
 *     a play-time error accumulator to fix exactly those cases);
 *   - speed only: pitch, rate and chord-pitch stages are the identity on the
 *     reference's path (soniclib.c never calls them except sonicIntSetRate,
 *     which is stored and otherwise ignored here); volume scaling is kept.
 *
 * No golden output samples, pitch periods or output counts for this stage exist
 * anywhere in the reference, so the restatement is pinned only
 * through the reference's own property tests, re-expressed in
 * tests/test_oracle_sonic.py: output length (sonic_classic_test.cc:201-203,
 * 263-265, 514, 533, 574), mono == stereo sample-exact
 * (sonic_classic_test.cc:659-665, 703-708), silent channel stays 0
 * (sonic_test.cc:859-860), Teager-energy bounds (sonic_test.cc:528-530).
 *
 * Compiled with -DSONIC_INTERNAL the symbols are exported as sonicInt* (see
 * shim/sonic.h), which is what /root/reference/soniclib.c links against.
 */
#include "shim/sonic.h"

#include <stdlib.h>
#include <string.h>

struct sonicStreamStruct {
  short* inputBuffer;      /* interleaved FIFO of samples not yet consumed */
  short* outputBuffer;     /* interleaved FIFO of samples not yet read */
  short* downSampleBuffer; /* mono, decimated copy used by the AMDF */
  void* userData;
  float speed;
  float volume;
  float pitch;
  float rate;
  int quality;
  int numChannels;
  int inputBufferSize;
  int outputBufferSize;
  int numInputSamples;
  int numOutputSamples;
  int minPeriod;
  int maxPeriod;
  int maxRequired;
  int remainingInputToCopy;
  int sampleRate;
  int prevPeriod;
  int prevMinDiff;
};

/* ---- buffers ---------------------------------------------------------- */

static int growInput(sonicStream s, int numSamples) {
  if (s->numInputSamples + numSamples > s->inputBufferSize) {
    s->inputBufferSize += (s->inputBufferSize >> 1) + numSamples;
    short* p = (short*)realloc(
        s->inputBuffer, (size_t)s->inputBufferSize * sizeof(short) * s->numChannels);
    if (!p) return 0;
    s->inputBuffer = p;
  }
  return 1;
}

static int growOutput(sonicStream s, int numSamples) {
  if (s->numOutputSamples + numSamples > s->outputBufferSize) {
    s->outputBufferSize += (s->outputBufferSize >> 1) + numSamples;
    short* p = (short*)realloc(
        s->outputBuffer,
        (size_t)s->outputBufferSize * sizeof(short) * s->numChannels);
    if (!p) return 0;
    s->outputBuffer = p;
  }
  return 1;
}

static int appendOutput(sonicStream s, const short* samples, int numSamples) {
  if (!growOutput(s, numSamples)) return 0;
  memcpy(s->outputBuffer + (size_t)s->numOutputSamples * s->numChannels, samples,
         (size_t)numSamples * sizeof(short) * s->numChannels);
  s->numOutputSamples += numSamples;
  return 1;
}

static void dropInput(sonicStream s, int position) {
  int remaining = s->numInputSamples - position;
  if (remaining > 0) {
    memmove(s->inputBuffer, s->inputBuffer + (size_t)position * s->numChannels,
            (size_t)remaining * sizeof(short) * s->numChannels);
  }
  s->numInputSamples = remaining;
}

/* ---- create / destroy / accessors ------------------------------------- */

static int amdfSkip(sonicStream s) {
  /* The AMDF is evaluated at about SONIC_AMDF_FREQ unless quality is set. */
  if (s->sampleRate > SONIC_AMDF_FREQ && s->quality == 0) {
    return s->sampleRate / SONIC_AMDF_FREQ;
  }
  return 1;
}

sonicStream sonicCreateStream(int sampleRate, int numChannels) {
  sonicStream s = (sonicStream)calloc(1, sizeof(struct sonicStreamStruct));
  if (!s) return NULL;
  s->sampleRate = sampleRate;
  s->numChannels = numChannels;
  s->minPeriod = sampleRate / SONIC_MAX_PITCH;
  s->maxPeriod = sampleRate / SONIC_MIN_PITCH;
  s->maxRequired = 2 * s->maxPeriod;
  s->inputBufferSize = s->maxRequired + (s->maxRequired >> 2);
  s->outputBufferSize = s->inputBufferSize;
  s->inputBuffer =
      (short*)calloc((size_t)s->inputBufferSize, sizeof(short) * numChannels);
  s->outputBuffer =
      (short*)calloc((size_t)s->outputBufferSize, sizeof(short) * numChannels);
  /* Sized for skip == 1 so that sonicSetQuality can change the decimation. */
  s->downSampleBuffer = (short*)calloc((size_t)s->maxRequired, sizeof(short));
  if (!s->inputBuffer || !s->outputBuffer || !s->downSampleBuffer) {
    sonicDestroyStream(s);
    return NULL;
  }
  s->speed = 1.0f;
  s->pitch = 1.0f;
  s->volume = 1.0f;
  s->rate = 1.0f;
  s->quality = 0;
  s->prevPeriod = 0;
  return s;
}

void sonicDestroyStream(sonicStream s) {
  if (!s) return;
  free(s->inputBuffer);
  free(s->outputBuffer);
  free(s->downSampleBuffer);
  free(s);
}

void sonicSetUserData(sonicStream s, void* userData) { s->userData = userData; }
void* sonicGetUserData(sonicStream s) { return s->userData; }
float sonicGetSpeed(sonicStream s) { return s->speed; }
void sonicSetSpeed(sonicStream s, float speed) { s->speed = speed; }
float sonicGetPitch(sonicStream s) { return s->pitch; }
void sonicSetPitch(sonicStream s, float pitch) { s->pitch = pitch; }
float sonicGetRate(sonicStream s) { return s->rate; }
void sonicSetRate(sonicStream s, float rate) { s->rate = rate; }
float sonicGetVolume(sonicStream s) { return s->volume; }
void sonicSetVolume(sonicStream s, float volume) { s->volume = volume; }
int sonicGetQuality(sonicStream s) { return s->quality; }
void sonicSetQuality(sonicStream s, int quality) { s->quality = quality; }
int sonicGetSampleRate(sonicStream s) { return s->sampleRate; }
int sonicGetNumChannels(sonicStream s) { return s->numChannels; }
int sonicSamplesAvailable(sonicStream s) { return s->numOutputSamples; }

/* ---- AMDF pitch-period search ----------------------------------------- */

/* Average `skip` multi-channel samples into one mono value (int sum, C integer
 * division), producing maxRequired/skip values. */
static void decimate(sonicStream s, const short* samples, int skip) {
  int count = s->maxRequired / skip;
  int perValue = s->numChannels * skip;
  for (int i = 0; i < count; i++) {
    int value = 0;
    for (int j = 0; j < perValue; j++) value += *samples++;
    value /= perValue;
    s->downSampleBuffer[i] = (short)value;
  }
}

/* Average magnitude difference over lags minPeriod..maxPeriod.  The best lag
 * minimises diff/period, the worst maximises it; comparisons are done by
 * cross-multiplication in 64 bits, scanning lags upwards with strict
 * inequalities (so the smallest lag wins ties).  Returns the best lag and the
 * per-sample difference at the best and worst lags. */
static int searchRange(const short* samples, int minPeriod, int maxPeriod,
                       int* retMinDiff, int* retMaxDiff) {
  int bestPeriod = 0, worstPeriod = 255;
  unsigned long long minDiff = 1, maxDiff = 0;
  for (int period = minPeriod; period <= maxPeriod; period++) {
    unsigned long long diff = 0;
    const short* a = samples;
    const short* b = samples + period;
    for (int i = 0; i < period; i++) {
      int x = *a++, y = *b++;
      diff += (unsigned long long)(x >= y ? x - y : y - x);
    }
    if (bestPeriod == 0 || diff * (unsigned)bestPeriod < minDiff * (unsigned)period) {
      minDiff = diff;
      bestPeriod = period;
    }
    if (diff * (unsigned)worstPeriod > maxDiff * (unsigned)period) {
      maxDiff = diff;
      worstPeriod = period;
    }
  }
  *retMinDiff = (int)(minDiff / (unsigned)bestPeriod);
  *retMaxDiff = (int)(maxDiff / (unsigned)worstPeriod);
  return bestPeriod;
}

/* At abrupt voiced/unvoiced transitions the previous period is the better
 * guess.  preferNew is always 1 on the speed-change path. */
static int previousPeriodBetter(sonicStream s, int minDiff, int maxDiff,
                                int preferNew) {
  if (minDiff == 0 || s->prevPeriod == 0) return 0;
  if (preferNew) {
    if (maxDiff > minDiff * 3) return 0;            /* good match this time */
    if (minDiff * 2 <= s->prevMinDiff * 3) return 0; /* not much worse */
  } else {
    if (minDiff <= s->prevMinDiff) return 0;
  }
  return 1;
}

static int findPitchPeriod(sonicStream s, const short* samples, int preferNew) {
  int minPeriod = s->minPeriod, maxPeriod = s->maxPeriod;
  int minDiff, maxDiff, period, result;
  int skip = amdfSkip(s);

  if (s->numChannels == 1 && skip == 1) {
    period = searchRange(samples, minPeriod, maxPeriod, &minDiff, &maxDiff);
  } else {
    decimate(s, samples, skip);
    period = searchRange(s->downSampleBuffer, minPeriod / skip, maxPeriod / skip,
                         &minDiff, &maxDiff);
    if (skip != 1) {
      /* Refine around the coarse estimate at the full rate. */
      period *= skip;
      minPeriod = period - (skip << 2);
      maxPeriod = period + (skip << 2);
      if (minPeriod < s->minPeriod) minPeriod = s->minPeriod;
      if (maxPeriod > s->maxPeriod) maxPeriod = s->maxPeriod;
      if (s->numChannels == 1) {
        period = searchRange(samples, minPeriod, maxPeriod, &minDiff, &maxDiff);
      } else {
        decimate(s, samples, 1);
        period = searchRange(s->downSampleBuffer, minPeriod, maxPeriod, &minDiff,
                             &maxDiff);
      }
    }
  }
  result = previousPeriodBetter(s, minDiff, maxDiff, preferNew) ? s->prevPeriod
                                                               : period;
  s->prevMinDiff = minDiff;
  s->prevPeriod = period;
  return result;
}

/* ---- overlap-add ------------------------------------------------------ */

/* Linear cross-fade per channel: out[t] = (down[t]*(n-t) + up[t]*t) / n with C
 * integer arithmetic (truncating division). */
static void overlapAdd(int n, int numChannels, short* out, const short* rampDown,
                       const short* rampUp) {
  for (int c = 0; c < numChannels; c++) {
    short* o = out + c;
    const short* d = rampDown + c;
    const short* u = rampUp + c;
    for (int t = 0; t < n; t++) {
      *o = (short)((*d * (n - t) + *u * t) / n);
      o += numChannels;
      d += numChannels;
      u += numChannels;
    }
  }
}

/* speed > 1: drop (part of) a pitch period.  Returns output samples made. */
static int skipPitchPeriod(sonicStream s, const short* samples, float speed,
                           int period) {
  long newSamples;
  if (speed >= 2.0f) {
    newSamples = (long)(period / (speed - 1.0f));
  } else {
    newSamples = period;
    s->remainingInputToCopy = (int)(period * (2.0f - speed) / (speed - 1.0f));
  }
  if (!growOutput(s, (int)newSamples)) return 0;
  overlapAdd((int)newSamples, s->numChannels,
             s->outputBuffer + (size_t)s->numOutputSamples * s->numChannels,
             samples, samples + (size_t)period * s->numChannels);
  s->numOutputSamples += (int)newSamples;
  return (int)newSamples;
}

/* speed < 1: repeat (part of) a pitch period.  Returns input samples to skip. */
static int insertPitchPeriod(sonicStream s, const short* samples, float speed,
                             int period) {
  long newSamples;
  if (speed < 0.5f) {
    newSamples = (long)(period * speed / (1.0f - speed));
  } else {
    newSamples = period;
    s->remainingInputToCopy =
        (int)(period * (2.0f * speed - 1.0f) / (1.0f - speed));
  }
  if (!growOutput(s, period + (int)newSamples)) return 0;
  short* out = s->outputBuffer + (size_t)s->numOutputSamples * s->numChannels;
  memcpy(out, samples, (size_t)period * sizeof(short) * s->numChannels);
  out += (size_t)period * s->numChannels;
  overlapAdd((int)newSamples, s->numChannels, out,
             samples + (size_t)period * s->numChannels, samples);
  s->numOutputSamples += period + (int)newSamples;
  return (int)newSamples;
}

/* Copy through the input PICOLA wants played unmodified, at most maxRequired
 * samples at a time.  Returns the number copied. */
static int copyThrough(sonicStream s, int position) {
  int n = s->remainingInputToCopy;
  if (n > s->maxRequired) n = s->maxRequired;
  if (!appendOutput(s, s->inputBuffer + (size_t)position * s->numChannels, n)) {
    return 0;
  }
  s->remainingInputToCopy -= n;
  return n;
}

/* Consume as many pitch periods as are buffered. */
static int changeSpeed(sonicStream s, float speed) {
  int numSamples = s->numInputSamples;
  int position = 0, newSamples;
  int maxRequired = s->maxRequired;

  if (numSamples < maxRequired) return 1;
  do {
    if (s->remainingInputToCopy > 0) {
      newSamples = copyThrough(s, position);
      position += newSamples;
    } else {
      const short* samples = s->inputBuffer + (size_t)position * s->numChannels;
      int period = findPitchPeriod(s, samples, 1);
      if (speed > 1.0) {
        newSamples = skipPitchPeriod(s, samples, speed, period);
        position += period + newSamples;
      } else {
        newSamples = insertPitchPeriod(s, samples, speed, period);
        position += newSamples;
      }
    }
    if (newSamples == 0) return 0; /* nothing produced: give up on this write */
  } while (position + maxRequired <= numSamples);
  dropInput(s, position);
  return 1;
}

static void scaleVolume(short* samples, int count, float volume) {
  int fixedPoint = (int)(volume * 4096.0f);
  while (count--) {
    int value = (*samples * fixedPoint) >> 12;
    if (value > 32767) value = 32767;
    if (value < -32767) value = -32767;
    *samples++ = (short)value;
  }
}

static int processInput(sonicStream s) {
  int firstNew = s->numOutputSamples;
  float speed = s->speed / s->pitch;

  if (speed > 1.00001 || speed < 0.99999) {
    changeSpeed(s, speed);
  } else {
    if (!appendOutput(s, s->inputBuffer, s->numInputSamples)) return 0;
    s->numInputSamples = 0;
  }
  if (s->volume != 1.0f) {
    scaleVolume(s->outputBuffer + (size_t)firstNew * s->numChannels,
                (s->numOutputSamples - firstNew) * s->numChannels, s->volume);
  }
  return 1;
}

/* ---- stream I/O -------------------------------------------------------- */

int sonicWriteShortToStream(sonicStream s, const short* samples, int numSamples) {
  if (numSamples > 0) {
    if (!growInput(s, numSamples)) return 0;
    memcpy(s->inputBuffer + (size_t)s->numInputSamples * s->numChannels, samples,
           (size_t)numSamples * sizeof(short) * s->numChannels);
    s->numInputSamples += numSamples;
  }
  return processInput(s);
}

int sonicWriteFloatToStream(sonicStream s, const float* samples, int numSamples) {
  if (numSamples > 0) {
    if (!growInput(s, numSamples)) return 0;
    short* dst = s->inputBuffer + (size_t)s->numInputSamples * s->numChannels;
    int count = numSamples * s->numChannels;
    while (count--) *dst++ = (short)((*samples++) * 32767.0f);
    s->numInputSamples += numSamples;
  }
  return processInput(s);
}

int sonicReadShortFromStream(sonicStream s, short* samples, int maxSamples) {
  int n = s->numOutputSamples, left = 0;
  if (n == 0) return 0;
  if (n > maxSamples) {
    left = n - maxSamples;
    n = maxSamples;
  }
  memcpy(samples, s->outputBuffer, (size_t)n * sizeof(short) * s->numChannels);
  if (left > 0) {
    memmove(s->outputBuffer, s->outputBuffer + (size_t)n * s->numChannels,
            (size_t)left * sizeof(short) * s->numChannels);
  }
  s->numOutputSamples = left;
  return n;
}

int sonicReadFloatFromStream(sonicStream s, float* samples, int maxSamples) {
  int n = s->numOutputSamples, left = 0;
  if (n == 0) return 0;
  if (n > maxSamples) {
    left = n - maxSamples;
    n = maxSamples;
  }
  const short* src = s->outputBuffer;
  int count = n * s->numChannels;
  while (count--) *samples++ = (*src++) / 32767.0f;
  if (left > 0) {
    memmove(s->outputBuffer, s->outputBuffer + (size_t)n * s->numChannels,
            (size_t)left * sizeof(short) * s->numChannels);
  }
  s->numOutputSamples = left;
  return n;
}

/* Force out whatever is buffered: pad with 2*maxRequired zeros, process, then
 * trim the output to the length the real samples should have produced. */
int sonicFlushStream(sonicStream s) {
  int maxRequired = s->maxRequired;
  int remaining = s->numInputSamples;
  float speed = s->speed / s->pitch;
  float rate = s->rate * s->pitch;
  int expected =
      s->numOutputSamples + (int)((remaining / speed + 0 /* pitch FIFO */) / rate + 0.5f);

  if (!growInput(s, 2 * maxRequired)) return 0;
  memset(s->inputBuffer + (size_t)remaining * s->numChannels, 0,
         (size_t)2 * maxRequired * sizeof(short) * s->numChannels);
  s->numInputSamples += 2 * maxRequired;
  if (!sonicWriteShortToStream(s, NULL, 0)) return 0;
  if (s->numOutputSamples > expected) s->numOutputSamples = expected;
  s->numInputSamples = 0;
  s->remainingInputToCopy = 0;
  return 1;
}
