// speedy_wave — command-line driver for the nonlinear speed-up path: WAV in, WAV out.
//
// Same flags and behaviour as the reference's tool (/root/reference/speedy_wave.cc:154-471):
// --input / --output, --speed, --nonlinear f (any positive value enables Speedy with
// factor 1.0, speedy_wave.cc:177), --linear, --duration_feedback_strength,
// --match_nonlinear (first measure the speed a nonlinear run achieves, then speed up
// linearly by that, :424-427), --length seconds (two-pass calibration, :428-462) and the
// five debug dumps --tension_file / --speed_file / --features_file / --spectrogram_file /
// --normalized_spectrogram_file (one "%g" line per frame, :68-122).  It talks to the
// library only through the Sonic/Speedy C API of include/speedy_b200.h, exactly as the
// reference tool talks to libspeedy.  The WAV reader/writer is ours (the reference uses
// upstream Sonic's wave.c, which is not vendored): 16-bit PCM RIFF files.
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "speedy_b200.h"

namespace {

struct Wave {
  int rate = 0;
  int channels = 0;
  std::vector<int16_t> samples;  // interleaved
};

uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

bool read_wave(const std::string& path, Wave* w) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<unsigned char> buf;
  unsigned char tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
  fclose(f);
  if (buf.size() < 12 || memcmp(&buf[0], "RIFF", 4) || memcmp(&buf[8], "WAVE", 4)) return false;
  size_t pos = 12;
  int bits = 0, format = 0;
  while (pos + 8 <= buf.size()) {
    const uint32_t len = rd32(&buf[pos + 4]);
    const unsigned char* body = &buf[pos + 8];
    if (!memcmp(&buf[pos], "fmt ", 4) && len >= 16) {
      format = rd16(body);
      w->channels = rd16(body + 2);
      w->rate = (int)rd32(body + 4);
      bits = rd16(body + 14);
    } else if (!memcmp(&buf[pos], "data", 4)) {
      size_t avail = buf.size() - (pos + 8);
      size_t bytes = len < avail ? len : avail;
      if (format != 1 || bits != 16 || w->channels < 1) return false;
      w->samples.resize(bytes / 2);
      for (size_t i = 0; i < w->samples.size(); i++) w->samples[i] = (int16_t)rd16(body + 2 * i);
      return true;
    }
    pos += 8 + len + (len & 1);
  }
  return false;
}

bool write_wave(const std::string& path, int rate, int channels, const std::vector<int16_t>& s) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const uint32_t bytes = (uint32_t)(s.size() * 2);
  unsigned char h[44];
  auto w32 = [&](int o, uint32_t v) { h[o] = v & 255; h[o + 1] = (v >> 8) & 255; h[o + 2] = (v >> 16) & 255; h[o + 3] = (v >> 24) & 255; };
  auto w16 = [&](int o, uint16_t v) { h[o] = v & 255; h[o + 1] = (v >> 8) & 255; };
  memcpy(h, "RIFF", 4); w32(4, 36 + bytes); memcpy(h + 8, "WAVEfmt ", 8); w32(16, 16); w16(20, 1);
  w16(22, (uint16_t)channels); w32(24, (uint32_t)rate); w32(28, (uint32_t)(rate * channels * 2));
  w16(32, (uint16_t)(channels * 2)); w16(34, 16); memcpy(h + 36, "data", 4); w32(40, bytes);
  bool ok = fwrite(h, 1, 44, f) == 44;
  for (size_t i = 0; ok && i < s.size(); i++) {
    unsigned char b[2] = {(unsigned char)(s[i] & 255), (unsigned char)((s[i] >> 8) & 255)};
    ok = fwrite(b, 1, 2, f) == 2;
  }
  fclose(f);
  return ok;
}

FILE *tension_fp, *speed_fp, *features_fp, *spectrogram_fp, *normalized_fp;

void tension_saver(sonicStream, int, float t) { if (tension_fp) fprintf(tension_fp, "%g\n", t); }
void speed_saver(sonicStream, int, float s) { if (speed_fp) fprintf(speed_fp, "%g\n", s); }
void features_saver(sonicStream, int, float* f) {
  if (!features_fp) return;
  for (int i = 0; i < SPEEDY_FEATURE_COUNT; i++) fprintf(features_fp, "%g ", f[i]);
  fprintf(features_fp, "\n");
}
void dump_row(FILE* fp, sonicStream s, float* v) {
  if (!fp) return;
  const int n = sonicSpectrogramSize(s);
  for (int i = 0; i < n; i++) fprintf(fp, "%g ", v[i]);
  fprintf(fp, "\n");
}
void spectrogram_saver(sonicStream s, int, float* v) { dump_row(spectrogram_fp, s, v); }
void normalized_saver(sonicStream s, int, float* v) { dump_row(normalized_fp, s, v); }

// speedy_wave.cc:154-242.  Returns the achieved speed-up.
double compress_sound(const Wave& in, double speed, double nonlinear, double feedback, const std::string& out_path) {
  const int max_samples = 1000;
  sonicStream s = sonicCreateStream(in.rate, in.channels);
  if (!s) {
    fprintf(stderr, "speedy_wave: cannot create a stream (no CUDA device?)\n");
    exit(2);
  }
  sonicSetSpeed(s, (float)speed);
  sonicEnableNonlinearSpeedup(s, nonlinear > 0.0 ? 1.0f : 0.0f);
  sonicSetDurationFeedbackStrength(s, (float)feedback);
  if (nonlinear > 0.0 && !out_path.empty()) {
    sonicTensionCallback(s, tension_saver);
    sonicSpeedCallback(s, speed_saver);
    sonicFeaturesCallback(s, features_saver);
    sonicSpectrogramCallback(s, spectrogram_saver);
    sonicNormalizedSpectrogramCallback(s, normalized_saver);
  }
  std::vector<int16_t> out, buf((size_t)max_samples * in.channels);
  const long total = (long)(in.samples.size() / in.channels);
  long produced = 0;
  for (long t = 0; t < total; t += max_samples) {
    const int n = (int)(total - t < max_samples ? total - t : max_samples);
    if (sonicWriteShortToStream(s, &in.samples[(size_t)t * in.channels], n) <= 0) {
      fprintf(stderr, "speedy_wave: sonicWriteShortToStream failed\n");
      exit(2);
    }
    const int got = sonicReadShortFromStream(s, buf.data(), max_samples);
    produced += got;
    out.insert(out.end(), buf.begin(), buf.begin() + (size_t)got * in.channels);
  }
  sonicFlushStream(s);
  for (;;) {
    const int got = sonicReadShortFromStream(s, buf.data(), max_samples);
    if (got <= 0) break;
    produced += got;
    out.insert(out.end(), buf.begin(), buf.begin() + (size_t)got * in.channels);
  }
  sonicDestroyStream(s);
  if (!out_path.empty() && !write_wave(out_path, in.rate, in.channels, out)) {
    fprintf(stderr, "speedy_wave: cannot write %s\n", out_path.c_str());
    exit(1);
  }
  printf("Compress_sound read %ld frames, and output %ld frames with nonlinear=%g.\n", total, produced, nonlinear);
  return produced > 0 ? (double)total / produced : 0.0;
}

FILE* open_dump(const char* arg) {
  FILE* fp = fopen(arg, "w");
  if (!fp) {
    fprintf(stderr, "speedy_wave: cannot open %s\n", arg);
    exit(1);
  }
  return fp;
}

}  // namespace

int main(int argc, char** argv) {
  static const char* usage =
      "Usage: %s [--speed 3.0]\n"
      "\t[--nonlinear 1.0] [--linear] [--match_nonlinear] [--length seconds]\n"
      "\t[--duration_feedback_strength 0.0]\n"
      "\t[--tension_file f] [--speed_file f] [--features_file f]\n"
      "\t[--spectrogram_file f] [--normalized_spectrogram_file f]\n"
      "\t--input sound.wav --output fastsound.wav\n"
      "\t [set nonlinear to 0.0 to get a linear speedup.]\n";
  double speed = 3.0, feedback = 0.0, nonlinear = 1.0, desired_length = 0.0;  // speedy_wave.cc:32-37
  int match_nonlinear = 0;
  std::string input, output;
  if (argc <= 1) {
    fprintf(stderr, usage, argv[0]);
    return 255;
  }
  static struct option opts[] = {
      {"match_nonlinear", no_argument, nullptr, 'm'}, {"linear", no_argument, nullptr, 'l'},
      {"input", required_argument, nullptr, 'i'}, {"output", required_argument, nullptr, 'o'},
      {"speed", required_argument, nullptr, 's'}, {"nonlinear", required_argument, nullptr, 'n'},
      {"length", required_argument, nullptr, 'e'}, {"tension_file", required_argument, nullptr, 't'},
      {"speed_file", required_argument, nullptr, 'p'}, {"features_file", required_argument, nullptr, 'f'},
      {"spectrogram_file", required_argument, nullptr, 'S'},
      {"duration_feedback_strength", required_argument, nullptr, 'd'},
      {"normalized_spectrogram_file", required_argument, nullptr, 'N'}, {"help", no_argument, nullptr, 'h'},
      {nullptr, 0, nullptr, 0}};
  int c;
  while ((c = getopt_long(argc, argv, "mli:o:s:n:h", opts, nullptr)) != -1) {
    switch (c) {
      case 'm': match_nonlinear = 1; break;
      case 'l': nonlinear = 0.0; break;
      case 'i': input = optarg; break;
      case 'o': output = optarg; break;
      case 's': speed = strtod(optarg, nullptr); break;
      case 'n': nonlinear = strtod(optarg, nullptr); break;
      case 'e': desired_length = strtod(optarg, nullptr); break;
      case 'd': feedback = strtod(optarg, nullptr); break;
      case 't': tension_fp = open_dump(optarg); break;
      case 'p': speed_fp = open_dump(optarg); break;
      case 'f': features_fp = open_dump(optarg); break;
      case 'S': spectrogram_fp = open_dump(optarg); break;
      case 'N': normalized_fp = open_dump(optarg); break;
      case 'h': printf(usage, argv[0]); return 0;
      default: fprintf(stderr, usage, argv[0]); return 1;
    }
  }
  if (speed <= 0.0 || feedback < 0.0 || nonlinear < 0.0 || nonlinear > 2.0 || desired_length < 0.0) {
    fprintf(stderr, "speedy_wave: bad argument value\n");
    return 1;
  }
  if (output.empty()) {
    printf("%s: Must specify an output file name.\n", argv[0]);
    return 1;
  }
  if (input.empty()) {
    printf("%s: Must specify an input file name.\n", argv[0]);
    return 1;
  }
  Wave in;
  if (!read_wave(input, &in)) {
    fprintf(stderr, "Can't open %s for speedy input.\n", input.c_str());
    return 255;
  }
  printf("Read %d channel data at a sample rate of %d.\n", in.channels, in.rate);
  if (match_nonlinear) {
    // speedy_wave.cc:424-427
    speed = compress_sound(in, speed, 1.0, feedback, "");
  } else if (desired_length > 0) {
    // speedy_wave.cc:428-462
    const double input_length = (double)(in.samples.size() / in.channels) / (float)in.rate;
    const double desired_speed = input_length / desired_length;
    const double new_speed = compress_sound(in, desired_speed, 1.0, feedback, "");
    speed = desired_speed * (desired_speed / new_speed);
    printf("First scaling by %g gave a speed of %g.\n", desired_speed, new_speed);
  }
  printf("Reading sound from %s and speeding it up %s by %gX into %s.\n", input.c_str(),
         nonlinear > 0.0 ? "non-linearly" : "linearly", speed, output.c_str());
  compress_sound(in, speed, nonlinear, feedback, output);
  for (FILE* fp : {tension_fp, speed_fp, features_fp, spectrogram_fp, normalized_fp}) {
    if (fp) fclose(fp);
  }
  return 0;
}
