/* Oracle (TEST INFRASTRUCTURE): plain-C twin of oracle/vote_subset.py, used for the full-size
 * CPU baseline (bench.py cpu_baseline / --impl reference) and cross-checked against the numpy
 * version in tests/test_oracle_vote.py.  Not part of the product; never linked into libcpros.so.
 *
 * vote:   /root/reference/code/models.py:149-166 (prefix mode over the 25-sample window, torch-CPU
 *         tie rule = smallest label; equality count against arange(41)).
 * subset: README.md:11,15 (no runnable reference): restricted argmax -> window vote -> count.
 *
 * build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/liboracle_vote.so oracle/vote_subset.c
 */
#include <stdint.h>
#include <string.h>

#define MAXT 64

/* pred: (B,W,T) int32.  votes: (B,n_votes) int32 (#correct among T for window sizes 1..n_votes,
 * clamped at W rows).  y_pred: (B,T) int64 = full-window mode. */
void oracle_vote(const int32_t *pred, int64_t B, int W, int T, int n_votes,
                 int32_t *votes, int64_t *y_pred) {
    for (int64_t b = 0; b < B; ++b) {
        int32_t correct_at[256];
        int32_t mode_now[MAXT];
        int32_t cnt[MAXT][MAXT];
        memset(cnt, 0, sizeof(cnt));
        for (int w = 0; w < W; ++w) {
            int c = 0;
            for (int i = 0; i < T; ++i) {
                int l = pred[(b * W + w) * T + i];
                cnt[i][l]++;
                int best = 0;
                for (int j = 1; j < T; ++j) if (cnt[i][j] > cnt[i][best]) best = j;
                mode_now[i] = best;
                c += (best == i);
            }
            correct_at[w] = c;
        }
        for (int v = 0; v < n_votes; ++v) {
            int w = (v + 1 < W ? v + 1 : W) - 1;
            votes[b * n_votes + v] = correct_at[w];
        }
        for (int i = 0; i < T; ++i) y_pred[b * T + i] = mode_now[i];
    }
}

/* logits: (B,W,T,T) float32; masks: (n_trials,T) uint8; correct/total: (n_trials,) int64. */
void oracle_subset(const float *logits, int64_t B, int W, int T, const uint8_t *masks,
                   int64_t n_trials, int64_t *correct, int64_t *total) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < n_trials; ++t) {
        int S[MAXT], ns = 0;
        for (int j = 0; j < T; ++j) if (masks[t * T + j]) S[ns++] = j;
        int64_t c = 0;
        for (int64_t b = 0; b < B; ++b) {
            for (int a = 0; a < ns; ++a) {
                int i = S[a];
                int cnt[MAXT];
                memset(cnt, 0, sizeof(cnt));
                for (int w = 0; w < W; ++w) {
                    const float *row = logits + (((b * W + w) * T) + i) * T;
                    int best = S[0];
                    float bv = row[best];
                    for (int q = 1; q < ns; ++q) {
                        float v = row[S[q]];
                        if (v > bv) { bv = v; best = S[q]; }
                    }
                    cnt[best]++;
                }
                int mode = 0;
                for (int j = 1; j < T; ++j) if (cnt[j] > cnt[mode]) mode = j;
                c += (mode == i);
            }
        }
        correct[t] = c;
        total[t] = B * ns;
    }
}
