// Kernels either side of the scan: ingest (fp32 -> bf16 + row L2 norms), candidate merge + exact
// rescoring, and the cross-shard merge that follows the all-gather.  All are HBM/latency-bound
// integer/byte work on CUDA cores: coalesced 16-byte accesses, one warp per row / one block per query.
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

namespace sa {

__host__ __device__ __forceinline__ uint32_t aux_f32_bits(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, sizeof u);
  return u;
#endif
}
__host__ __device__ __forceinline__ float aux_bits_f32(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, sizeof f);
  return f;
#endif
}
__host__ __device__ __forceinline__ float bf16_bits_to_f32(uint32_t b) { return aux_bits_f32(b << 16); }
// Round-to-nearest-even fp32 -> bf16 bit pattern (NaN kept quiet); same rule as the oracle's numpy code.
__host__ __device__ __forceinline__ uint32_t f32_to_bf16_bits(float f) {
  uint32_t u = aux_f32_bits(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (u >> 16) | 0x40u;
  return (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inv_norm[r] = 1/sqrt(sum_j row[r][j]^2) over the stored bf16 values, 0 for an all-zero row.
// One warp per row, 16-byte loads (dim % 8 == 0).
__global__ void sa_rownorm_kernel(const uint16_t* __restrict__ rows, float* __restrict__ inv_norm, long long first,
                                  long long n, int dim) {
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const uint4* src = reinterpret_cast<const uint4*>(rows + (first + w) * dim);
  float ss = 0.f;
  for (int i = lane; i < dim / 8; i += 32) {
    const uint4 x = __ldg(src + i);
    const uint32_t u[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = bf16_bits_to_f32(u[k] & 0xffffu), b = bf16_bits_to_f32(u[k] >> 16);
      ss = fmaf(a, a, ss);
      ss = fmaf(b, b, ss);
    }
  }
  ss = warp_sum(ss);
  if (lane == 0) inv_norm[first + w] = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
}

// dst_bf16[r][:] = RNE(src_f32[r][:]); optionally also inv_norm[r] (corpus ingest).  One warp per row.
__global__ void sa_convert_rows_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                       float* __restrict__ inv_norm, long long n, int dim) {
  const long long w = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const float4* s = reinterpret_cast<const float4*>(src + w * dim);
  uint2* d = reinterpret_cast<uint2*>(dst + w * dim);
  float ss = 0.f;
  for (int i = lane; i < dim / 4; i += 32) {
    const float4 x = __ldg(s + i);
    const uint32_t b0 = f32_to_bf16_bits(x.x), b1 = f32_to_bf16_bits(x.y), b2 = f32_to_bf16_bits(x.z),
                   b3 = f32_to_bf16_bits(x.w);
    d[i] = make_uint2(b0 | (b1 << 16), b2 | (b3 << 16));
    const float r0 = bf16_bits_to_f32(b0), r1 = bf16_bits_to_f32(b1), r2 = bf16_bits_to_f32(b2),
                r3 = bf16_bits_to_f32(b3);
    ss = fmaf(r0, r0, ss);
    ss = fmaf(r1, r1, ss);
    ss = fmaf(r2, r2, ss);
    ss = fmaf(r3, r3, ss);
  }
  if (inv_norm != nullptr) {
    ss = warp_sum(ss);
    if (lane == 0) inv_norm[w] = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
  }
}

// Order-preserving map: (score desc, row asc)  <=>  key desc.
__host__ __device__ __forceinline__ unsigned long long make_key(float s, int row) {
  uint32_t u = aux_f32_bits(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<unsigned long long>(u) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(row));
}
__host__ __device__ __forceinline__ int key_row(unsigned long long k) { return static_cast<int>(~static_cast<uint32_t>(k)); }

struct MergeParams {
  const float* part_score;  // [grid CTAs][128][qpu][kKL] from the scan
  const int* part_idx;
  const uint16_t* corpus;   // [capacity][dim] bf16
  const uint16_t* queries;  // [nq][dim] bf16 (this launch's queries)
  int dim;
  int nq;
  int k;
  int cg;                   // CTAs per unit in the scan that produced the lists
  int nqb;
  int tl_count;
  int qpu;                  // query blocks per unit in that scan (lists per CTA row)
  int unit_map;             // same mapping switch as ScanParams::unit_map
  float* out_score;         // [nq][k] fp32 cosine
  int* out_idx;             // [nq][k] shard-local row, -1 when fewer than k rows qualify
  double* out_score64;      // [nq][k] or nullptr: the unrounded cosine, for cross-shard merging
};

constexpr int kMergeThreads = 128;
constexpr int kMaxCand = 148 * 32;  // TL * kKL upper bound

// One block per query: (1) select the kKL best of the TL per-CTA lists by (approx score desc, row asc);
// (2) re-score those candidates exactly -- bf16 x bf16 products are exact in fp32, the sums of products and
// of squares run in fp64 -- so the final order equals the fp64 brute-force order; (3) sort, emit k.
//
// The rescoring set is kSel = 2*kKL wide.  What is guaranteed: every tile lane keeps its kKL >= k+4 best rows whose
// approximate score reaches the shared threshold, and that threshold never exceeds the query's global kKL-th best
// approximate score; so the union always contains the global approximate top-kKL, and usually (the shared threshold
// sits near the global (kKL*TL)-th best) the top-2*kKL as well.  A wrong answer therefore needs more than kKL-k rows
// (6 at k = 10) within the scan's fp32 rounding error (~1e-6 relative) of the query's k-th best score.
template <int kKL>
__global__ void __launch_bounds__(kMergeThreads) sa_merge_rescore_kernel(const MergeParams p) {
  constexpr int kSel = 2 * kKL;
  __shared__ unsigned long long keys[kMaxCand];
  __shared__ unsigned long long sel[kSel];
  __shared__ unsigned long long wbest[kMergeThreads / 32];
  __shared__ double cs[kSel];
  __shared__ double qq_s;
  __shared__ int nsel_s;

  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rows_per_unit = 128 * p.cg;
  const int qb = q / rows_per_unit;
  const int r = q % rows_per_unit;
  const int cta_in_unit = r / 128, row = r % 128;
  const int ncand = p.tl_count * kKL;

  for (int i = tid; i < ncand; i += kMergeThreads) {
    const int tl = i / kKL, e = i % kKL;
    const int nslots = (p.nqb + p.qpu - 1) / p.qpu;
    const int slot = qb / p.qpu, pass = qb % p.qpu;
    const int unit = p.unit_map == 0 ? tl * nslots + slot : slot * p.tl_count + tl;
    const size_t cta = static_cast<size_t>(unit) * p.cg + cta_in_unit;
    const size_t o = ((cta * 128 + row) * p.qpu + pass) * kKL + e;
    keys[i] = make_key(p.part_score[o], p.part_idx[o]);
  }
  const int max_sel = min(kSel, ncand);
  if (tid == 0) nsel_s = max_sel;
  __syncthreads();

  // up to kSel rounds of block-wide arg-max.  Keys of real candidates are unique (rows are unique per query).
  for (int round = 0; round < max_sel; ++round) {
    unsigned long long best = 0;
    int pos = -1;
    for (int i = tid; i < ncand; i += kMergeThreads) {
      const unsigned long long kk = keys[i];
      if (kk > best) {
        best = kk;
        pos = i;
      }
    }
    unsigned long long wb = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, wb, o);
      wb = other > wb ? other : wb;
    }
    if (lane == 0) wbest[warp] = wb;
    __syncthreads();
    unsigned long long gb = wbest[0];
#pragma unroll
    for (int w = 1; w < kMergeThreads / 32; ++w) gb = wbest[w] > gb ? wbest[w] : gb;
    if (pos >= 0 && best == gb && key_row(gb) >= 0) keys[pos] = 0;  // owner retires it
    if (tid == 0) {
      sel[round] = gb;
      if (key_row(gb) < 0 && nsel_s == max_sel) nsel_s = round;  // only empty slots remain
    }
    __syncthreads();
    if (nsel_s != max_sel) break;
  }
  const int nsel = nsel_s;

  // exact rescoring: warp w takes candidates w, w+4, ...
  const uint4* qv = reinterpret_cast<const uint4*>(p.queries + static_cast<size_t>(q) * p.dim);
  const int nvec = p.dim / 8;
  if (warp == 0) {
    double qq = 0.0;
    for (int i = lane; i < nvec; i += 32) {
      const uint4 x = __ldg(qv + i);
      const uint32_t u[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = bf16_bits_to_f32(u[k] & 0xffffu), b = bf16_bits_to_f32(u[k] >> 16);
        qq += static_cast<double>(a * a);
        qq += static_cast<double>(b * b);
      }
    }
    qq = warp_sum(qq);
    if (lane == 0) qq_s = qq;
  }
  __syncthreads();
  const double qq = qq_s;
  for (int c = warp; c < nsel; c += kMergeThreads / 32) {
    const int crow = key_row(sel[c]);
    const uint4* cv = reinterpret_cast<const uint4*>(p.corpus + static_cast<size_t>(crow) * p.dim);
    double dot = 0.0, dd = 0.0;
    for (int i = lane; i < nvec; i += 32) {
      const uint4 x = __ldg(qv + i);
      const uint4 y = __ldg(cv + i);
      const uint32_t u[4] = {x.x, x.y, x.z, x.w};
      const uint32_t v[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a0 = bf16_bits_to_f32(u[k] & 0xffffu), a1 = bf16_bits_to_f32(u[k] >> 16);
        const float b0 = bf16_bits_to_f32(v[k] & 0xffffu), b1 = bf16_bits_to_f32(v[k] >> 16);
        dot += static_cast<double>(a0 * b0);
        dot += static_cast<double>(a1 * b1);
        dd += static_cast<double>(b0 * b0);
        dd += static_cast<double>(b1 * b1);
      }
    }
    dot = warp_sum(dot);
    dd = warp_sum(dd);
    if (lane == 0) {
      const double den = qq * dd;
      cs[c] = den > 0.0 ? dot / sqrt(den) : 0.0;
    }
  }
  __syncthreads();

  if (tid == 0) {
    // insertion sort of <= kSel entries by (cosine desc, row asc)
    int ord[kSel];
    for (int i = 0; i < nsel; ++i) {
      const double ci = cs[i];
      const int ri = key_row(sel[i]);
      int j = i;
      while (j > 0) {
        const double cj = cs[ord[j - 1]];
        const int rj = key_row(sel[ord[j - 1]]);
        if (ci > cj || (ci == cj && ri < rj)) {
          ord[j] = ord[j - 1];
          --j;
        } else {
          break;
        }
      }
      ord[j] = i;
    }
    for (int i = 0; i < p.k; ++i) {
      const size_t o = static_cast<size_t>(q) * p.k + i;
      if (i < nsel) {  // nsel >= min(k, eligible rows): kSel > k
        p.out_score[o] = static_cast<float>(cs[ord[i]]);
        p.out_idx[o] = key_row(sel[ord[i]]);
        if (p.out_score64) p.out_score64[o] = cs[ord[i]];
      } else {
        p.out_score[o] = -INFINITY;
        p.out_idx[o] = -1;
        if (p.out_score64) p.out_score64[o] = -INFINITY;
      }
    }
  }
}

// After the all-gather: per query, merge G shard lists of k (already sorted, global row ids) into the
// global top-k by (cosine desc, global row asc).  One thread per query; G*k <= a few hundred.
__global__ void sa_merge_shards_kernel(const double* __restrict__ score64, const long long* __restrict__ gidx,
                                       int n_shards, int nq, int k, float* __restrict__ out_score,
                                       long long* __restrict__ out_idx) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  int head[64];
  for (int g = 0; g < n_shards; ++g) head[g] = 0;
  for (int i = 0; i < k; ++i) {
    int bg = -1;
    double bs = 0.0;
    long long bi = 0;
    for (int g = 0; g < n_shards; ++g) {
      if (head[g] >= k) continue;
      const size_t o = (static_cast<size_t>(g) * nq + q) * k + head[g];
      const long long ri = gidx[o];
      if (ri < 0) {
        head[g] = k;
        continue;
      }
      const double s = score64[o];
      if (bg < 0 || s > bs || (s == bs && ri < bi)) {
        bg = g;
        bs = s;
        bi = ri;
      }
    }
    const size_t oo = static_cast<size_t>(q) * k + i;
    if (bg >= 0) {
      out_score[oo] = static_cast<float>(bs);
      out_idx[oo] = bi;
      ++head[bg];
    } else {
      out_score[oo] = -INFINITY;
      out_idx[oo] = -1;
    }
  }
}

}  // namespace sa
