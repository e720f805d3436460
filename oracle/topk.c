/* oracle/topk.c -- TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/bruteforce.py).
 *
 * The selection half of the CPU brute-force baseline: merge one chunk of fp32 similarity scores [nq x m] into the
 * running per-query top-k by (score descending, row index ascending) -- the ordering rule of the oracle
 * (oracle/bruteforce.py::_select_topk; operator semantics terraform/lab2-vector-search/main.tf:292).  numpy's
 * argpartition does this on one thread and dominates the sgemm on a many-core host; this does it on all of them.
 * Built by __graft_entry__.build() with `gcc -O3 -fopenmp -shared -fPIC` into oracle/_build/ (git-ignored).
 */
#include <math.h>
#include <stdint.h>

/* best_s / best_i: [nq x k], sorted, unused slots hold (-inf, -1).  scores: [nq x m] row-major, may contain -inf for
 * rows that must never be returned.  Row r of the chunk has global index row_offset + r. */
void oracle_topk_merge_f32(const float* scores, int64_t nq, int64_t m, int k, int64_t row_offset, float* best_s,
                           int64_t* best_i) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t q = 0; q < nq; ++q) {
    const float* s = scores + q * m;
    float* bs = best_s + q * k;
    int64_t* bi = best_i + q * k;
    float thr = bs[k - 1];
    int full = bi[k - 1] >= 0;
    for (int64_t r = 0; r < m; ++r) {
      const float v = s[r];
      /* chunks arrive in ascending row order, so an equal score never displaces an earlier (lower) row */
      if (full ? !(v > thr) : !(v > -INFINITY)) continue;
      int pos = k - 1;
      while (pos > 0 && (bi[pos - 1] < 0 || v > bs[pos - 1])) {
        bs[pos] = bs[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
      }
      bs[pos] = v;
      bi[pos] = row_offset + r;
      thr = bs[k - 1];
      full = bi[k - 1] >= 0;
    }
  }
}
