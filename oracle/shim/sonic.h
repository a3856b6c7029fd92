/* TEST INFRASTRUCTURE (oracle) — not part of the product path.
 *
 * Stand-in for the header of upstream Sonic (github.com/waywardgeek/sonic),
 * which the reference includes through sonic2.h (/root/reference/sonic2.h:34-35)
 * but does not vendor (Makefile:7,17-18,74 expect an un-pinned clone in
 * ../sonic).  It is absent from this environment.  This header declares the
 * call surface the reference uses (soniclib.c:94,98,106,144,145,174,182,354,
 * 369,398,464,521,526,547,551; sonic_test.cc:370,735-750) and
 * oracle/sonic_oracle.c restates the published algorithm.
 *
 * As upstream does, defining SONIC_INTERNAL renames every public symbol to
 * sonicInt* with macros, so that sonic2.h can #undef the few it overrides
 * (sonic2.h:38-48) and define its own functions under the original names.
 */
#ifndef SPEEDY_B200_ORACLE_SONIC_H_
#define SPEEDY_B200_ORACLE_SONIC_H_

#ifdef __cplusplus
extern "C" {
#endif

#ifdef SONIC_INTERNAL
#define sonicCreateStream sonicIntCreateStream
#define sonicDestroyStream sonicIntDestroyStream
#define sonicSetUserData sonicIntSetUserData
#define sonicGetUserData sonicIntGetUserData
#define sonicWriteFloatToStream sonicIntWriteFloatToStream
#define sonicWriteShortToStream sonicIntWriteShortToStream
#define sonicReadFloatFromStream sonicIntReadFloatFromStream
#define sonicReadShortFromStream sonicIntReadShortFromStream
#define sonicFlushStream sonicIntFlushStream
#define sonicSamplesAvailable sonicIntSamplesAvailable
#define sonicGetSpeed sonicIntGetSpeed
#define sonicSetSpeed sonicIntSetSpeed
#define sonicGetPitch sonicIntGetPitch
#define sonicSetPitch sonicIntSetPitch
#define sonicGetRate sonicIntGetRate
#define sonicSetRate sonicIntSetRate
#define sonicGetVolume sonicIntGetVolume
#define sonicSetVolume sonicIntSetVolume
#define sonicGetQuality sonicIntGetQuality
#define sonicSetQuality sonicIntSetQuality
#define sonicGetSampleRate sonicIntGetSampleRate
#define sonicGetNumChannels sonicIntGetNumChannels
#endif /* SONIC_INTERNAL */

/* Pitch range searched by the AMDF, and the rate the search is decimated to. */
#define SONIC_MIN_PITCH 65
#define SONIC_MAX_PITCH 400
#define SONIC_AMDF_FREQ 4000

struct sonicStreamStruct;
typedef struct sonicStreamStruct* sonicStream;

sonicStream sonicCreateStream(int sampleRate, int numChannels);
void sonicDestroyStream(sonicStream stream);
void sonicSetUserData(sonicStream stream, void* userData);
void* sonicGetUserData(sonicStream stream);
int sonicWriteFloatToStream(sonicStream stream, const float* samples,
                            int numSamples);
int sonicWriteShortToStream(sonicStream stream, const short* samples,
                            int numSamples);
int sonicReadFloatFromStream(sonicStream stream, float* samples,
                             int maxSamples);
int sonicReadShortFromStream(sonicStream stream, short* samples,
                             int maxSamples);
int sonicFlushStream(sonicStream stream);
int sonicSamplesAvailable(sonicStream stream);
float sonicGetSpeed(sonicStream stream);
void sonicSetSpeed(sonicStream stream, float speed);
float sonicGetPitch(sonicStream stream);
void sonicSetPitch(sonicStream stream, float pitch);
float sonicGetRate(sonicStream stream);
void sonicSetRate(sonicStream stream, float rate);
float sonicGetVolume(sonicStream stream);
void sonicSetVolume(sonicStream stream, float volume);
int sonicGetQuality(sonicStream stream);
void sonicSetQuality(sonicStream stream, int quality);
int sonicGetSampleRate(sonicStream stream);
int sonicGetNumChannels(sonicStream stream);

#ifdef __cplusplus
}
#endif
#endif
