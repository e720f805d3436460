// The hot path: VECTOR_SEARCH_AGG(<corpus>, DESCRIPTOR(embedding), <query vector>, k)
// (reference call sites: terraform/lab2-vector-search/main.tf:292, LAB3-Walkthrough.md:343-350,
//  LAB4-Walkthrough.md:302-309) as one persistent, warp-specialised sm_100a kernel:
//
//   TMA (SWIZZLE_128B tiles of the bf16 corpus and of the query block)  ->  smem ring
//   tcgen05.mma  Q[128 x D] . C[256 x D]^T, fp32 accumulators in TMEM (two 256-column buffers)
//   epilogue warps: tcgen05.ld -> scale by the row's 1/|c| -> per-thread (thread == query) sorted register list
//
// Nothing but the per-CTA candidate lists (kKL entries per query, plus one "dropped" bound per query) leaves the SM.
//
// Work decomposition.  A "unit" is one CTA (kCG == 1, 128-query blocks) or one CTA pair (kCG == 2, 256-query
// blocks, tcgen05 cta_group::2).  Unit u owns query block qb = u % nqb and tile lane tl = u / nqb and walks corpus
// tiles tl, tl + TL, tl + 2 TL, ... (256 rows each).  All units of one tile lane touch the same corpus tile at about
// the same time (drift control below keeps it so): it crosses HBM once and is served from L2 to the others.
//
// Exactness contract with the merge kernel (sa_aux.cuh).  A thread's list holds the kKL best rows of its tile lane by
// the scan's approximate score a = fp32_accumulate(q.c) * (1/|c|), and `drop` is an upper bound on the approximate
// score of every row of the lane that is NOT in the list (evicted, rejected by the own threshold, or rejected by the
// bound shared between lanes).  The merge kernel turns (lists, drops) into either a certificate that the exactly
// re-scored candidates contain the true top-k, or a work item for the exact fallback scan.
#pragma once
#include "sm100_ptx.cuh"
#include <cmath>
#include <cstring>

namespace sa {

constexpr int kBlockM = 128;  // queries per CTA  (TMEM lanes)
constexpr int kBlockN = 256;  // corpus rows per tile (TMEM columns per accumulator)
constexpr int kBlockK = 64;   // bf16 per K slice = 128 B = one swizzle atom
constexpr int kUmmaK = 16;
constexpr int kScanThreads = 256;  // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, w4..7 epilogue
constexpr int kTmemCols = 512;
constexpr int kChunk = 32;         // TMEM columns per tcgen05.ld
constexpr int kWin = 16;           // tile lanes whose second-best scores an epilogue thread combines into a bound
constexpr int kWinWarmTiles = 16;  // the window is read on every one of a lane's first tiles, later only after a slow one

// kMode of the scan kernel
constexpr int kModeProd = 0;   // production
constexpr int kModeDots = 1;   // test hook: also dump the raw accumulators of one tile
constexpr int kModeProf = 2;   // profiling: per-role wait / busy cycle counters (ScanParams::prof)

template <int kCG>
struct ScanCfg {
  static constexpr int kStages = (kCG == 1) ? 4 : 6;
  static constexpr int kBRows = kBlockN / kCG;  // corpus rows staged by each CTA
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = kBRows * kBlockK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kIcBytes = 4 * kBlockN * sizeof(float);  // one 256-float scale vector per epilogue warp
  static constexpr uint32_t kBarBytes = (2 * kStages + 4) * 8 + 16;
  // +1024: the dynamic smem base is aligned up to 1024 B by hand (SWIZZLE_128B requirement).
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kIcBytes + kBarBytes + 1024;
};

// Per-CTA profile record (kModeProf): SM cycles, summed over the kernel.
struct ScanProf {
  long long prod_wait_empty;   // TMA producer blocked on a free smem slot
  long long mma_wait_full;     // MMA issuer blocked on TMA data
  long long mma_wait_tempty;   // MMA issuer blocked on the epilogue (accumulator not drained)
  long long epi_wait_tfull;    // epilogue warp 0 blocked on the MMA (accumulator not complete)
  long long epi_busy;          // epilogue warp 0 working on an accumulator
  long long epi_slow_chunks;   // 32-column chunks of epilogue warp 0 that took the insertion path
  long long total;             // CTA lifetime
  long long tiles;             // tiles walked
};

struct ScanParams {
  const float* inv_norm;  // [capacity] 1/|row| over the bf16-rounded row, 0 for an all-zero row; 16-byte aligned
  long long n_rows;       // committed rows (epoch snapshot); rows >= n_rows are masked
  int nq;                 // queries covered by tmap_q
  int num_kb;             // D / 64
  int num_tiles;          // ceil(n_rows / 256)
  int nqb;                // query blocks of 128*kCG rows
  int tl_count;           // tile lanes (TL)
  float* part_score;      // [gridDim.x][128][kKL]
  int* part_idx;          // [gridDim.x][128][kKL]
  float* part_drop;       // [gridDim.x][128]  upper bound on the approximate score of the lane's rows not in the list
  int corpus_evict_first; // 1: corpus tiles are read by a single query block -> stream them through L2
  int tile_stride;        // 1: walk every tile; S > 1: the sampling pre-pass walks tiles 0, S, 2S, ... only
  int wait_hint_ns;       // suspend-time hint of the epilogue's mbarrier waits (0 = plain polling)
  int* lane_progress;     // [tl_count][nqb] tiles whose loads each unit has issued (zero at launch), or nullptr
  int unit_map;           // 0: unit = tl*nqb + qb (lane-mates adjacent), 1: unit = qb*TL + tl (lane-mates TL apart)
  int max_drift;          // lead (in tiles) over the slowest lane-mate that is not paced
  int pace_gain;          // SM cycles of delay per K-slice issue per tile of lead beyond max_drift (0 = free-running)
  int pace_max;           // cap of that delay
  unsigned* thr_shared;   // [nqb*128*kCG] per-query lower bound on the kKL-th best score, order-preserving keys
                          // (zero at launch), or nullptr: lanes then learn their thresholds alone
  unsigned* lane2;        // [tl_count][nqb*128*kCG] each lane's SECOND-best score per query (keys, zero at launch), or
                          // nullptr.  kKL/2 lanes with two rows >= x each are kKL rows >= x: a much tighter bound than
                          // any single lane's kKL-th best while the lists are young (window_bound below)
  long long* dbg_times;   // optional [gridDim.x][2]: globaltimer at CTA start / end (ns), for drift studies
  float* dbg_dots;        // kModeDots only: raw accumulators of (unit 0 .. nqb-1, tile dbg_tile) [nqb*128*kCG][256]
  int dbg_tile;
  ScanProf* prof;         // kModeProf only: [gridDim.x]
};

// Bit casts usable on both sides of the compiler: the device path is the intrinsic, the host path (used only by the
// CPU unit tests through sa_debug_*) is a memcpy.
__host__ __device__ __forceinline__ unsigned f32_bits(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  unsigned u;
  memcpy(&u, &f, sizeof u);
  return u;
#endif
}
__host__ __device__ __forceinline__ float bits_f32(unsigned u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, sizeof f);
  return f;
#endif
}

// Order-preserving float <-> unsigned key (larger float <=> larger key; key 0 is below every float).
__host__ __device__ __forceinline__ unsigned float_to_key(float f) {
  const unsigned u = f32_bits(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float key_to_float(unsigned k) {
  return bits_f32((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// The largest float strictly less than x (x finite): `s > float_below(x)` <=> `s >= x`.
__host__ __device__ __forceinline__ float float_below(float x) {
  const int b = static_cast<int>(f32_bits(x));
  if (x > 0.f) return bits_f32(static_cast<unsigned>(b - 1));
  if (x < 0.f) return bits_f32(static_cast<unsigned>(b + 1));
  return bits_f32(0x80000001u);  // below +-0: the smallest negative denormal
}

// max that ignores a NaN operand (device: FMNMX; host: the same rule spelled out for the CPU unit tests)
__host__ __device__ __forceinline__ float max_nn(float a, float b) {
#ifdef __CUDA_ARCH__
  return fmaxf(a, b);
#else
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return (f32_bits(a) & 0x80000000u) ? b : a;  // max(+0, -0) = +0, as FMNMX
  return a > b ? a : b;
#endif
}

// Sorted (descending score, ascending row on ties) insertion into a register-resident list.
// Precondition: s > sc[kKL-1].  Rows reach a thread in ascending order, so a strict compare keeps the
// lower row index ahead of an equal score.
template <int kKL>
__host__ __device__ __forceinline__ void list_insert(float (&sc)[kKL], int (&id)[kKL], float s, int row) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = kKL - 1; i > 0; --i) {
    const bool shift = s > sc[i - 1];
    const bool here = s > sc[i];
    const float ns = shift ? sc[i - 1] : (here ? s : sc[i]);
    const int ni = shift ? id[i - 1] : (here ? row : id[i]);
    sc[i] = ns;
    id[i] = ni;
  }
  if (s > sc[0]) {
    sc[0] = s;
    id[0] = row;
  }
}

// The window bound.  Every tile lane publishes, per query, the SECOND best score it holds.  If kKL/2 different lanes each
// hold two rows scoring >= x, kKL rows score >= x (lanes own disjoint rows), so no row scoring < x is in the query's
// global top-kKL: x = the (kKL/2)-th largest of the lanes' second bests is a valid shared bound.  After t tiles per
// lane it sits near the top 1.7/(256 t) of the scores, the best single lane's kKL-th best near 9/(256 t): about five
// times fewer values pass it, which is what the first tiles of a short scan spend their time on.  16 keys, bitonic
// network (80 compare-exchanges, branch-free; key 0 = nothing published sorts last, so an early window yields 0 = no bound).
__host__ __device__ __forceinline__ void sort16_desc(unsigned (&x)[kWin]) {
  static_assert(kWin == 16, "the network below is the 16-input bitonic sorter");
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int k = 2; k <= 16; k <<= 1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = k >> 1; j > 0; j >>= 1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int i = 0; i < 16; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned a = x[i], b = x[l];
          const unsigned hi = a > b ? a : b, lo = a > b ? b : a;
          const bool desc = (i & k) == 0;
          x[i] = desc ? hi : lo;
          x[l] = desc ? lo : hi;
        }
      }
    }
  }
}
template <int kKL>
__host__ __device__ __forceinline__ unsigned window_bound(unsigned (&x)[kWin]) {
  static_assert(kKL / 2 <= kWin, "the window must hold kKL/2 lanes");
  sort16_desc(x);
  return x[kKL / 2 - 1];
}

// One query's candidate list as an epilogue thread holds it: all indices are compile-time, so it lives in registers.
template <int kKL>
struct TopList {
  float sc[kKL];
  int id[kKL];
  float thr;        // current insertion threshold = max(own kKL-th best, thr_floor)
  float thr_floor;  // largest float strictly below the bound shared by the other tile lanes
  float published;  // last own kKL-th best written to the shared bound
  float drop;       // max approximate score of any row this thread saw and does not hold (NaN-free; -inf = none)
  unsigned nxt_key; // shared bound fetched at the end of the previous accumulator (0 = nothing published yet)
  unsigned* slot;   // this query's shared bound (or nullptr)
  __host__ __device__ __forceinline__ void init(unsigned* shared_slot) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < kKL; ++i) {
      sc[i] = -INFINITY;
      id[i] = -1;
    }
    thr = thr_floor = published = drop = -INFINITY;
    nxt_key = 0u;
    slot = shared_slot;
  }
  // a bound published by another tile lane becomes visible
  __host__ __device__ __forceinline__ void apply_shared(unsigned key) {
    if (key != 0u) {
      thr_floor = max_nn(thr_floor, float_below(key_to_float(key)));  // bounds only ever tighten
      thr = max_nn(thr, thr_floor);
    }
  }
};

// One 32-column chunk: scale, reduce to the chunk maximum with full instruction-level parallelism, and only when that
// beats the threshold (probability ~ 32 kKL / n after n rows) take the insertion path.  Returns true if it ran.
//
// Insertion path: extract-max rounds over the whole chunk -- each round takes the thread's largest remaining value (the
// lowest row among equals), inserts it, masks it and re-reduces -- until nothing beats the threshold.  A warp executes
// as many rounds as its busiest lane needs (usually one or two), however the qualifying values are spread over the 32
// rows; walking the rows group by group instead costs a round per group that ANY lane has a hit in, which during the
// warm-up of a short scan (1M rows: 26 tiles per lane) was most of the epilogue's time (A/B on B200,
// profiles/r02_ab_insertion_walk.json: epilogue 15.1k -> 9.6k cycles per tile at 1M x 1536 / batch 256, scan +7..15 %;
// 7.3k -> 4.6k at 6.25M x 768 / batch 128; neutral at batch 1024).
template <int kKL>
__host__ __device__ __forceinline__ bool chunk_process(TopList<kKL>& L, float (&v)[kChunk], const float (&w)[kChunk],
                                                       int row_base) {
  auto reduce = [&]() {
    float g[8];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < 8; ++i)
      g[i] = max_nn(max_nn(v[4 * i + 0], v[4 * i + 1]), max_nn(v[4 * i + 2], v[4 * i + 3]));
    return max_nn(max_nn(max_nn(g[0], g[1]), max_nn(g[2], g[3])), max_nn(max_nn(g[4], g[5]), max_nn(g[6], g[7])));
  };
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int i = 0; i < kChunk; ++i) v[i] *= w[i];
  float m = reduce();
  if (!(m > L.thr)) {
    L.drop = max_nn(L.drop, m);
    return false;
  }
  do {
    int pos = kChunk - 1;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = kChunk - 2; j >= 0; --j) pos = (v[j] == m) ? j : pos;  // lowest row among equals first
    float sj = m;
    if (m == 0.f) {  // keep the element's own sign of zero (max(+0, -0) is +0)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
      for (int j = 0; j < kChunk; ++j) sj = (j == pos) ? v[j] : sj;
    }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 0; j < kChunk; ++j) v[j] = (j == pos) ? -INFINITY : v[j];
    L.drop = max_nn(L.drop, L.sc[kKL - 1]);  // the evicted tail (-inf while the list fills)
    list_insert<kKL>(L.sc, L.id, sj, row_base + pos);
    L.thr = max_nn(L.sc[kKL - 1], L.thr_floor);
    m = reduce();
  } while (m > L.thr);
  L.drop = max_nn(L.drop, m);  // whatever is left of the chunk was rejected (NaN = masked rows: ignored)
  return true;
}

#ifdef __CUDACC__
// Epilogue of one accumulator: 256 columns of this thread's TMEM lane -> scaled scores -> list.
// `ic` is this warp's private 256-float scale vector in shared memory (broadcast reads).
template <int kKL, int kMode>
__device__ __forceinline__ int epilogue_accumulator(TopList<kKL>& L, uint32_t taddr, const float* ic, int row0,
                                                    float* dbg_row) {
  L.apply_shared(L.nxt_key);
  int slow = 0;
  float va[kChunk], vb[kChunk], w[kChunk];
  const float4* ic4 = reinterpret_cast<const float4*>(ic);
  auto load_w = [&](int c) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x = ic4[c * 8 + i];
      w[4 * i + 0] = x.x;
      w[4 * i + 1] = x.y;
      w[4 * i + 2] = x.z;
      w[4 * i + 3] = x.w;
    }
  };
  auto dump = [&](int c, const float (&v)[kChunk]) {
    if constexpr (kMode == kModeDots) {
      if (dbg_row != nullptr) {
#pragma unroll
        for (int j = 0; j < kChunk; ++j) dbg_row[c * kChunk + j] = v[j];
      }
    }
  };
  __syncwarp();  // tcgen05.ld / wait::ld are .sync.aligned
  tmem_ld_32x32(taddr, va);
#pragma unroll 1
  for (int c = 0; c < kBlockN / kChunk; c += 2) {
    // chunk c is in flight into va: fetch its scales, wait, start chunk c+1 into vb, then work on va
    load_w(c);
    tmem_ld_wait(va);
    tmem_ld_32x32(taddr + static_cast<uint32_t>((c + 1) * kChunk), vb);
    dump(c, va);
    slow += chunk_process<kKL>(L, va, w, row0 + c * kChunk) ? 1 : 0;
    __syncwarp();  // reconverge after the divergent insertion path
    load_w(c + 1);
    tmem_ld_wait(vb);
    if (c + 2 < kBlockN / kChunk) tmem_ld_32x32(taddr + static_cast<uint32_t>((c + 2) * kChunk), va);
    dump(c + 1, vb);
    slow += chunk_process<kKL>(L, vb, w, row0 + (c + 1) * kChunk) ? 1 : 0;
    __syncwarp();
  }
  if (L.slot != nullptr) {
    if (L.sc[kKL - 1] > L.published) {  // list full and its tail improved: tell the other lanes
      L.published = L.sc[kKL - 1];
      atomicMax(L.slot, float_to_key(L.published));
    }
    L.nxt_key = ld_relaxed_gpu_u32(L.slot);  // consumed at the start of the next accumulator: latency hidden
  }
  return slow;
}

template <int kCG, int kKL, int kMode>
__global__ void __launch_bounds__(kScanThreads, 1)
sa_scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
               const ScanParams p) {
  using Cfg = ScanCfg<kCG>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kRowsPerQb = kBlockM * kCG;
  constexpr bool kProf = (kMode == kModeProf);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  float* icbuf = reinterpret_cast<float*>(smem_gen + kStages * Cfg::kStageBytes);  // [4 warps][256]
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes + Cfg::kIcBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + kStages * Cfg::kStageBytes + Cfg::kIcBytes +
                                                    (2 * kStages + 4) * 8);
  auto a_smem = [&](int s) { return smem_base + s * Cfg::kStageBytes; };
  auto b_smem = [&](int s) { return smem_base + s * Cfg::kStageBytes + Cfg::kABytes; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (kCG == 2) ? cluster_ctarank() : 0u;
  const int unit = blockIdx.x / kCG;
  const int TL = p.tl_count;
  const int nqb = p.nqb;  // units per tile lane
  const int qb = p.unit_map == 0 ? unit % nqb : unit / TL;
  const int tl = p.unit_map == 0 ? unit / nqb : unit % TL;
  const int walk_tiles = (p.num_tiles + p.tile_stride - 1) / p.tile_stride;  // tiles this launch visits (all lanes together)

  // ------------------------------------------------------------------ one-time setup
  long long t_start = 0;
  if constexpr (kProf) t_start = clock64();
  if (p.dbg_times != nullptr && threadIdx.x == 0) p.dbg_times[2 * blockIdx.x] = globaltimer_ns();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_c);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);   // the (leader) producer's arrive.expect_tx; TMA bytes complete it
      mbar_init(empty_bar(s), 1);  // one tcgen05.commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);         // tcgen05.commit after an accumulator's last MMA
      mbar_init(tempty_bar(a), 4 * kCG);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<kCG>(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish<kCG>();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ------------------------------------------------------------------ roles
  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      const uint64_t c_hint = p.corpus_evict_first ? kEvictFirst : kEvictNormal;
      int stage = 0;
      uint32_t phase = 0;
      long long waited = 0;
      // Drift control between the units of a tile lane.  Units that share a corpus tile run identical work but at
      // slightly different speeds (measured: ~3 % spread), so over thousands of tiles they drift tens of tiles
      // apart; once the spread exceeds what L2 holds, every unit re-reads its tiles from HBM (measured 3.05x the
      // algorithmic bytes at B = 1024).  A hard barrier costs a pipeline drain per tile (measured +20 %), so the
      // leader producers *pace* themselves instead: each publishes how many tiles it has issued, reads its
      // lane-mates' counters once per tile, and a unit that leads the slowest mate by more than `max_drift` tiles
      // delays every K-slice issue by pace_gain cycles per extra tile of lead (capped).  The kernel's duration is
      // set by its slowest unit anyway, so slowing the fast ones is free; it is only a hint (no waiting on
      // anyone), hence no co-residency assumption and no deadlock.
      const bool lockstep = p.lane_progress != nullptr && p.pace_gain > 0 && nqb > 1 && rank == 0;
      int pace = 0;
      int tile_no = 0;
      const int q_row = qb * kRowsPerQb + static_cast<int>(rank) * kBlockM;
      for (int ti = tl; ti < walk_tiles; ti += TL, ++tile_no) {
        const int t = ti * p.tile_stride;
        if (lockstep) {
          const int* pr = p.lane_progress + tl * nqb;
          int slowest = tile_no;
          for (int j = 0; j < nqb; ++j) slowest = min(slowest, ld_relaxed_gpu(pr + j));
          pace = min(max(tile_no - slowest - p.max_drift, 0) * p.pace_gain, p.pace_max);
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if constexpr (kProf) {
            const long long c0 = clock64();
            mbar_wait(empty_bar(stage), phase ^ 1u);
            waited += clock64() - c0;
          } else {
            mbar_wait(empty_bar(stage), phase ^ 1u);
          }
          if (pace > 0) {
            const long long c0 = clock64();
            while (clock64() - c0 < pace) {
            }
          }
          if constexpr (kCG == 1) {
            mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
            tma_load_2d(a_smem(stage), &tmap_q, full_bar(stage), kb * kBlockK, q_row, kEvictLast);
            tma_load_2d(b_smem(stage), &tmap_c, full_bar(stage), kb * kBlockK, t * kBlockN, c_hint);
          } else {
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
            tma_load_2d_pair(a_smem(stage), &tmap_q, full_bar(stage), kb * kBlockK, q_row, kEvictLast);
            tma_load_2d_pair(b_smem(stage), &tmap_c, full_bar(stage), kb * kBlockK,
                             t * kBlockN + static_cast<int>(rank) * Cfg::kBRows, c_hint);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (lockstep) st_relaxed_gpu(p.lane_progress + tl * nqb + qb, tile_no + 1);
      }
      if constexpr (kProf) {
        p.prof[blockIdx.x].prod_wait_empty = waited;
        p.prof[blockIdx.x].tiles = tile_no;
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader CTA of a pair) =====
      constexpr uint32_t idesc = make_idesc_bf16_f32(kBlockM * kCG, kBlockN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long w_full = 0, w_tempty = 0;
      for (int ti = tl; ti < walk_tiles; ti += TL, ++it) {
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1u;
        if constexpr (kProf) {
          const long long c0 = clock64();
          mbar_wait(tempty_bar(a), aph ^ 1u);
          w_tempty += clock64() - c0;
        } else {
          mbar_wait(tempty_bar(a), aph ^ 1u);  // epilogue has drained this accumulator
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a * kBlockN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if constexpr (kProf) {
            const long long c0 = clock64();
            mbar_wait(full_bar(stage), phase);
            w_full += clock64() - c0;
          } else {
            mbar_wait(full_bar(stage), phase);
          }
          tc_fence_after();
          const uint64_t a_desc = make_kmajor_sw128_desc(a_smem(stage));
          const uint64_t b_desc = make_kmajor_sw128_desc(b_smem(stage));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // +32 B along K inside the 128-B swizzle atom = +2 in the (addr >> 4) field
            umma_bf16<kCG>(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit<kCG>(empty_bar(stage));  // frees the smem slot (in both CTAs) once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit<kCG>(tfull_bar(a));  // accumulator complete -> epilogue
      }
      if constexpr (kProf) {
        p.prof[blockIdx.x].mma_wait_full = w_full;
        p.prof[blockIdx.x].mma_wait_tempty = w_tempty;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread == query row; 4 warps cover the 128 TMEM lanes; the warps never synchronise with each
    // other (each stages its own copy of the tile's 256 scales), only with the MMA issuer through the mbarriers =====
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    const int et = ew * 32 + lane;
    float* ic = icbuf + ew * kBlockN;
    // Threshold sharing.  A thread's list only ever sees its own tile lane, so alone it needs ~kKL*ln(n) insertions
    // to warm up, and a warp pays for every lane's insertions.  But if ANY lane already holds kKL rows scoring >= x
    // for this query, no row scoring < x can be in the query's global top-kKL.  So each epilogue thread publishes its
    // kKL-th best (atomicMax on an order-preserving key) and reads the shared bound once per accumulator: every lane
    // gets the threshold of the whole machine's progress, and the warm-up tail disappears.  The shared bound admits
    // ties (>=), the thread's own bound stays strict (>), so tie-breaking by row is unchanged.
    const int query = qb * kRowsPerQb + static_cast<int>(rank) * kBlockM + et;
    TopList<kKL> L;
    L.init((p.thr_shared != nullptr && query < p.nq) ? p.thr_shared + query : nullptr);

    // Scales of tile t: lane l fetches rows [8l, 8l+8) (two 16-byte loads), masks rows past the committed prefix
    // and all-zero rows with NaN (NaN never compares greater than a threshold, so they cannot enter a list and
    // are ignored by every max), and the warp shares them through its private smem vector.
    float4 nx0 = make_float4(0.f, 0.f, 0.f, 0.f), nx1 = nx0;
    auto fetch_ic = [&](int t) {
      const long long r0 = static_cast<long long>(t) * kBlockN + 8 * lane;
      const float4* src = reinterpret_cast<const float4*>(p.inv_norm + r0);
      if (r0 + 8 <= p.n_rows) {
        nx0 = __ldg(src);
        nx1 = __ldg(src + 1);
      } else {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (r0 + j < p.n_rows) ? __ldg(p.inv_norm + r0 + j) : 0.f;
        nx0 = make_float4(x[0], x[1], x[2], x[3]);
        nx1 = make_float4(x[4], x[5], x[6], x[7]);
      }
    };
    if (tl < walk_tiles) fetch_ic(tl * p.tile_stride);

    // Window bound (see window_bound): this thread's slot in the lanes' second-best table, and the kWin lanes it reads.
    const bool win_on = p.lane2 != nullptr && TL >= kWin && query < p.nq;
    const size_t win_stride = static_cast<size_t>(p.nqb) * kRowsPerQb;
    unsigned* const win_q = win_on ? p.lane2 + query : nullptr;
    float pub2 = -INFINITY;   // last second-best published
    unsigned nx2[kWin];       // the window as read at the end of the previous accumulator
    bool have2 = false;

    long long w_tfull = 0, busy = 0, slow_chunks = 0;
    int it = 0;
    for (int ti = tl; ti < walk_tiles; ti += TL, ++it) {
      if (have2) {  // in the shadow of the wait for the MMA
        const unsigned kb = window_bound<kKL>(nx2);
        L.nxt_key = kb > L.nxt_key ? kb : L.nxt_key;
        have2 = false;
      }
      const int t = ti * p.tile_stride;
      const float qnan = __int_as_float(0x7fc00000);
      auto sc = [&](float x) { return x > 0.f ? x : qnan; };
      float4* dst = reinterpret_cast<float4*>(ic + 8 * lane);
      dst[0] = make_float4(sc(nx0.x), sc(nx0.y), sc(nx0.z), sc(nx0.w));
      dst[1] = make_float4(sc(nx1.x), sc(nx1.y), sc(nx1.z), sc(nx1.w));
      if (ti + TL < walk_tiles) fetch_ic((ti + TL) * p.tile_stride);  // prefetch the next tile's inverse norms
      __syncwarp();                                // ic[] visible to the whole warp
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1u;
      long long c0 = 0;
      if constexpr (kProf) c0 = clock64();
      mbar_wait(tfull_bar(a), aph, static_cast<uint32_t>(p.wait_hint_ns));
      tc_fence_after();
      long long c1 = 0;
      if constexpr (kProf) c1 = clock64();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(a * kBlockN);
      float* dbg_row = nullptr;
      if constexpr (kMode == kModeDots) {
        if (p.dbg_dots != nullptr && t == p.dbg_tile) dbg_row = p.dbg_dots + static_cast<size_t>(query) * kBlockN;
      }
      const int slow = epilogue_accumulator<kKL, kMode>(L, taddr, ic, t * kBlockN, dbg_row);
      tc_fence_before();
      __syncwarp();  // also orders this tile's ic[] reads before the next tile's writes
      if (lane == 0) {
        if constexpr (kCG == 1)
          mbar_arrive(tempty_bar(a));
        else
          mbar_arrive_cluster(tempty_bar(a), 0);  // the MMA issuer lives in the pair's leader CTA
      }
      if constexpr (kProf) {
        w_tfull += c1 - c0;
        busy += clock64() - c1;
        slow_chunks += slow;
      }
      // accumulator released: publish this lane's second best and, while the lists are young (or whenever a warp just
      // paid for insertions), fetch the window for the next accumulator's bound
      if (p.lane2 != nullptr && TL >= kWin) {
        const bool fetch = (it < kWinWarmTiles) || __any_sync(0xffffffffu, slow > 0);
        if (win_on) {
          if (L.sc[1] > pub2) {
            pub2 = L.sc[1];
            st_relaxed_gpu_u32(win_q + static_cast<size_t>(tl) * win_stride, float_to_key(pub2));
          }
          if (fetch && ti + TL < walk_tiles) {
#pragma unroll
            for (int i = 0; i < kWin; ++i) {
              int ln = tl + i;
              ln -= (ln >= TL) ? TL : 0;
              nx2[i] = ld_relaxed_gpu_u32(win_q + static_cast<size_t>(ln) * win_stride);
            }
            have2 = true;
          }
        }
      }
    }

    // The only global writes of the scan: this CTA's candidate list and drop bound for each of its queries.
    if (query < p.nq) {
      const size_t o = (static_cast<size_t>(blockIdx.x) * kBlockM + et) * kKL;
      float4* ps = reinterpret_cast<float4*>(p.part_score + o);
      int4* pi = reinterpret_cast<int4*>(p.part_idx + o);
#pragma unroll
      for (int i = 0; i < kKL / 4; ++i) {
        ps[i] = make_float4(L.sc[4 * i], L.sc[4 * i + 1], L.sc[4 * i + 2], L.sc[4 * i + 3]);
        pi[i] = make_int4(L.id[4 * i], L.id[4 * i + 1], L.id[4 * i + 2], L.id[4 * i + 3]);
      }
      p.part_drop[static_cast<size_t>(blockIdx.x) * kBlockM + et] = L.drop;
    }
    if constexpr (kProf) {
      if (ew == 0 && lane == 0) {
        p.prof[blockIdx.x].epi_wait_tfull = w_tfull;
        p.prof[blockIdx.x].epi_busy = busy;
        p.prof[blockIdx.x].epi_slow_chunks = slow_chunks;
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (p.dbg_times != nullptr && threadIdx.x == 0) p.dbg_times[2 * blockIdx.x + 1] = globaltimer_ns();
  if constexpr (kProf) {
    if (threadIdx.x == 0) p.prof[blockIdx.x].total = clock64() - t_start;
  }
  if constexpr (kCG == 2) cluster_sync_all();  // the peer may still be signalling our barriers / reading our smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kCG>(tmem_base, kTmemCols);
  }
}
#endif  // __CUDACC__

}  // namespace sa
