/* speedy_eval.h — the reference's evaluation tools, host C ABI.
 *
 * The reference's statistical tests judge a time-compressed signal with three
 * tools: dynamic time warping between spectrograms
 * (/root/reference/dynamic_time_warping.{h,cc}), the Teager energy operator on
 * sinusoids (sonic_test.cc:133-197) and least-squares slopes of the warping path
 * (sonic_test.cc:85-112).  They are restated here (libspeedy_eval.so, host C++
 * only, no CUDA) so that sonic_test.cc's tests can be re-run against the GPU
 * path's output with the reference's own thresholds.  All arithmetic is float,
 * in the reference's operation order.
 */
#ifndef SPEEDY_EVAL_H_
#define SPEEDY_EVAL_H_

#ifdef __cplusplus
extern "C" {
#endif

/* DynamicTimeWarping::Compute + BestPathSequence (dynamic_time_warping.cc:31-131)
 * with the Euclidean point distance of sonic_test.cc:199-209 (|a-b| when dim is 1,
 * as in dynamic_time_warping_test.cc:25-30).  seq1 is [len1][dim], seq2 is
 * [len2][dim].  Returns the optimal cost.  If path1/path2 are non-NULL they receive
 * the warping path (capacity len1 + len2 each) and *path_len its length. */
float speedyEvalDtw(const float* seq1, int len1, const float* seq2, int len2, int dim,
                    int* path1, int* path2, int* path_len);

/* TeagerVariance (sonic_test.cc:141-160): online mean / variance of
 * x[n]^2 - x[n-1] x[n+1], n = 1 .. count-2. */
void speedyEvalTeagerVarianceShort(const short* data, int count, float* mean, float* variance);
void speedyEvalTeagerVarianceFloat(const float* data, int count, float* mean, float* variance);

/* TeagerComputation (sonic_test.cc:162-171): out[n-1] for n = 1 .. count-2; returns count-2. */
int speedyEvalTeagerShort(const short* data, int count, float* out);

/* TeagerOutlierCount (sonic_test.cc:173-197): samples whose Teager energy is further
 * than thresh_fraction * mean from the mean. */
int speedyEvalTeagerOutlierCountShort(const short* data, int count, float thresh_fraction);

/* LinearSlope (sonic_test.cc:85-98) on integer paths, and LinearSlopeEverywhere
 * (sonic_test.cc:100-112): slopes of the windows [i-half, i+half), i = half .. n-half-1;
 * returns how many were written (n - 2 half, or 0). */
float speedyEvalLinearSlopeInt(const int* x, const int* y, int n);
int speedyEvalLinearSlopeEverywhereInt(const int* x, const int* y, int n, int half_width, float* slopes);

/* VectorMean / VectorStandardDeviation (sonic_test.cc:114-131). */
float speedyEvalMean(const float* v, int n);
float speedyEvalStandardDeviation(const float* v, int n);

#ifdef __cplusplus
}
#endif
#endif /* SPEEDY_EVAL_H_ */
