// libspeedy_eval.so — DTW, Teager-energy and path-slope statistics (include/speedy_eval.h).
// Host-only restatement of the reference's evaluation helpers; float arithmetic in the
// reference's order so its thresholds carry over unchanged.
#include <math.h>

#include <algorithm>
#include <vector>

#include "speedy_eval.h"

namespace {

float point_distance(const float* a, const float* b, int dim) {
  float sum2 = 0.0f;  // sonic_test.cc:199-209
  for (int i = 0; i < dim; i++) {
    const float d = a[i] - b[i];
    sum2 += d * d;
  }
  return (float)sqrt(sum2);
}

template <class T>
void teager_variance(const T* x, int count, float* mean, float* variance) {
  float m2 = 0.0f;
  *mean = 0.0f;
  for (int n = 1; n < count - 1; n++) {
    const float t = (float)(1.0 * x[n] * x[n] - 1.0 * x[n - 1] * x[n + 1]);
    const float delta = t - *mean;
    *mean += delta / n;
    m2 += delta * (t - *mean);
  }
  *variance = m2 / (count - 3);
}

}  // namespace

extern "C" {

float speedyEvalDtw(const float* seq1, int len1, const float* seq2, int len2, int dim, int* path1, int* path2,
                    int* path_len) {
  if (path_len) *path_len = 0;
  if (len1 <= 0 || len2 <= 0 || dim <= 0) return 0.0f;
  const size_t w = (size_t)len2;
  std::vector<float> cost((size_t)len1 * w);
  std::vector<signed char> dir((size_t)len1 * w, 0);  // -1: from (i-1,j)  0: diagonal  1: from (i,j-1)
  for (int i = 0; i < len1; i++)
    for (int j = 0; j < len2; j++) cost[i * w + j] = point_distance(seq1 + (size_t)i * dim, seq2 + (size_t)j * dim, dim);
  for (int j = 1; j < len2; j++) {
    cost[j] += cost[j - 1];
    dir[j] = 1;
  }
  for (int i = 1; i < len1; i++) {
    cost[i * w] += cost[(i - 1) * w];
    dir[i * w] = -1;
  }
  for (int i = 1; i < len1; i++) {
    for (int j = 1; j < len2; j++) {
      const float up = cost[(i - 1) * w + j], left = cost[i * w + j - 1], diag = cost[(i - 1) * w + j - 1];
      cost[i * w + j] += std::min(std::min(up, left), diag);
      // strict inequalities: ties follow the diagonal (dynamic_time_warping.cc:66-74)
      dir[i * w + j] = (up < diag && up < left) ? -1 : ((left < up && left < diag) ? 1 : 0);
    }
  }
  if (path1 && path2 && path_len) {
    int n = 0;
    for (int i = len1 - 1, j = len2 - 1; i >= 0 && j >= 0;) {
      path1[n] = i;
      path2[n] = j;
      n++;
      const int d = dir[i * w + j];
      if (d <= 0) i--;
      if (d >= 0) j--;
    }
    std::reverse(path1, path1 + n);
    std::reverse(path2, path2 + n);
    *path_len = n;
  }
  return cost.back();
}

void speedyEvalTeagerVarianceShort(const short* data, int count, float* mean, float* variance) {
  teager_variance(data, count, mean, variance);
}
void speedyEvalTeagerVarianceFloat(const float* data, int count, float* mean, float* variance) {
  teager_variance(data, count, mean, variance);
}

int speedyEvalTeagerShort(const short* x, int count, float* out) {
  int k = 0;
  for (int n = 1; n < count - 1; n++) out[k++] = (float)x[n] * x[n] - (float)x[n - 1] * x[n + 1];
  return k;
}

int speedyEvalTeagerOutlierCountShort(const short* x, int count, float thresh_fraction) {
  float mean, variance;
  teager_variance(x, count, &mean, &variance);
  const float threshold = mean * thresh_fraction;
  int errors = 0;
  for (int n = 1; n < count - 1; n++) {
    const float t = (float)x[n] * x[n] - (float)x[n - 1] * x[n + 1];
    if (fabs(t - mean) > threshold) errors++;
  }
  return errors;
}

float speedyEvalLinearSlopeInt(const int* x, const int* y, int n) {
  float sx = 0, sy = 0, sxy = 0, sx2 = 0;
  for (int i = 0; i < n; i++) {
    sx += x[i];
    sy += y[i];
    sxy += x[i] * y[i];
    sx2 += x[i] * x[i];
  }
  return (n * sxy - sx * sy) / (n * sx2 - sx * sx);
}

int speedyEvalLinearSlopeEverywhereInt(const int* x, const int* y, int n, int half, float* slopes) {
  int k = 0;
  for (int i = half; i < n - half; i++) slopes[k++] = speedyEvalLinearSlopeInt(x + i - half, y + i - half, 2 * half);
  return k;
}

float speedyEvalMean(const float* v, int n) {
  double acc = 0.0;
  for (int i = 0; i < n; i++) acc += v[i];
  return (float)acc / n;
}

float speedyEvalStandardDeviation(const float* v, int n) {
  double acc = 0.0;
  for (int i = 0; i < n; i++) acc += v[i];
  const float mean = (float)acc / n;
  double sq = 0.0;
  for (int i = 0; i < n; i++) {
    const float d = v[i] - mean;
    sq += d * d;
  }
  return (float)sqrt((float)sq / n);
}

}  // extern "C"
