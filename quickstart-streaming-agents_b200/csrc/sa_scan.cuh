// The hot path: VECTOR_SEARCH_AGG(<corpus>, DESCRIPTOR(embedding), <query vector>, k)
// (reference call sites: terraform/lab2-vector-search/main.tf:292, LAB3-Walkthrough.md:343-350,
//  LAB4-Walkthrough.md:302-309) as one persistent, warp-specialised sm_100a kernel:
//
//   TMA (SWIZZLE_128B tiles of the bf16 corpus and of the query block)  ->  smem ring
//   tcgen05.mma  Q[128 x D] . C[256 x D]^T, fp32 accumulators in TMEM (two 256-column buffers)
//   epilogue warps: tcgen05.ld -> scale by the row's 1/|c| -> per-thread (thread == query) sorted register list
//
// Nothing but the per-CTA candidate lists (kKL entries per query) leaves the SM.
//
// Work decomposition.  A "unit" is one CTA (kCG == 1, 128-query blocks) or one CTA pair (kCG == 2, 256-query
// blocks, tcgen05 cta_group::2) and carries kQPU (1 or 2) query blocks.  Unit u owns slot = u % nslots (its query
// blocks) and tile lane tl = u / nslots and walks corpus tiles tl, tl + TL, tl + 2 TL, ... (256 rows each), making
// one pass per query block over each.  All units of one tile lane touch the same corpus tile at about the same
// time (drift control below keeps it so): it crosses HBM once and is served from L2 to the others.
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

template <int kCG>
struct ScanCfg {
  static constexpr int kStages = (kCG == 1) ? 4 : 6;
  static constexpr int kBRows = kBlockN / kCG;  // corpus rows staged by each CTA
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = kBRows * kBlockK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kIcBytes = 2 * kBlockN * sizeof(float);
  static constexpr uint32_t kBarBytes = (2 * kStages + 4) * 8 + 16;
  // +1024: the dynamic smem base is aligned up to 1024 B by hand (SWIZZLE_128B requirement).
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kIcBytes + kBarBytes + 1024;
};

struct ScanParams {
  const float* inv_norm;  // [capacity] 1/|row| over the bf16-rounded row, 0 for an all-zero row
  long long n_rows;       // committed rows (epoch snapshot); rows >= n_rows are masked
  int nq;                 // queries covered by tmap_q
  int num_kb;             // D / 64
  int num_tiles;          // ceil(n_rows / 256)
  int nqb;                // query blocks of 128*kCG rows
  int tl_count;           // tile lanes (TL)
  float* part_score;      // [gridDim.x][128][kQPU][kKL]
  int* part_idx;          // [gridDim.x][128][kQPU][kKL]
  int corpus_evict_first; // 1: corpus tiles are read by a single query block -> stream them through L2
  int* lane_progress;     // [tl_count][nslots] tiles whose loads each unit has issued (zeroed before launch), or nullptr
  int unit_map;           // 0: unit = tl*nslots + slot (lane-mates adjacent), 1: unit = slot*TL + tl (lane-mates TL apart)
  int max_drift;          // lead (in tiles) over the slowest lane-mate that is not paced
  int pace_gain;          // SM cycles of delay per K-slice issue per tile of lead beyond max_drift (0 = free-running)
  int pace_max;           // cap of that delay
  unsigned* thr_shared;   // [nqb*128*kCG] per-query lower bound on the kKL-th best score, order-preserving keys
                          // (zeroed before launch), or nullptr: lanes then learn their thresholds alone
  long long* dbg_times;   // optional [gridDim.x][2]: globaltimer at CTA start / end (ns), for drift studies
  float* dbg_dots;        // debug builds only: raw accumulators of (unit 0 .. nqb-1, tile dbg_tile) [nqb*128*kCG][256]
  int dbg_tile;
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

// One query's candidate list as an epilogue thread holds it: all indices are compile-time, so it lives in registers.
template <int kKL>
struct TopList {
  float sc[kKL];
  int id[kKL];
  float thr;        // current insertion threshold = max(own kKL-th best, thr_floor)
  float thr_floor;  // largest float strictly below the bound shared by the other tile lanes
  float published;  // last own kKL-th best written to the shared bound
  unsigned* slot;   // this query's shared bound (or nullptr)
  __device__ __forceinline__ void init(unsigned* shared_slot) {
#pragma unroll
    for (int i = 0; i < kKL; ++i) {
      sc[i] = -INFINITY;
      id[i] = -1;
    }
    thr = thr_floor = published = -INFINITY;
    slot = shared_slot;
  }
};

// Epilogue of one accumulator: 256 columns of this thread's TMEM lane -> scaled scores -> list.
template <int kKL, bool kDebug>
__device__ __forceinline__ void epilogue_accumulator(TopList<kKL>& L, uint32_t taddr, const float4* ic4, int row0,
                                                     float* dbg_row) {
  if (L.slot != nullptr) {  // refresh the shared bound once per accumulator
    const unsigned key = ld_relaxed_gpu_u32(L.slot);
    if (key != 0u) {
      L.thr_floor = float_below(key_to_float(key));
      L.thr = fmaxf(L.thr, L.thr_floor);
    }
  }
#pragma unroll 1
  for (int c = 0; c < kBlockN / 32; ++c) {
    float v[32];
    __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent insert path
    tmem_ld_32x32(taddr + static_cast<uint32_t>(c * 32), v);
    tmem_ld_wait(v);
    if constexpr (kDebug) {
      if (dbg_row != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dbg_row[c * 32 + j] = v[j];
      }
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 w = ic4[c * 8 + g];
      float s0 = v[4 * g + 0] * w.x;
      float s1 = v[4 * g + 1] * w.y;
      float s2 = v[4 * g + 2] * w.z;
      float s3 = v[4 * g + 3] * w.w;
      float m = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
      while (m > L.thr) {  // rare after warm-up: ~kKL/n per value
        const int j = (s0 == m) ? 0 : (s1 == m) ? 1 : (s2 == m) ? 2 : 3;  // lowest row among equals first
        list_insert<kKL>(L.sc, L.id, m, row0 + c * 32 + g * 4 + j);
        L.thr = fmaxf(L.sc[kKL - 1], L.thr_floor);
        s0 = (j == 0) ? -INFINITY : s0;
        s1 = (j == 1) ? -INFINITY : s1;
        s2 = (j == 2) ? -INFINITY : s2;
        s3 = (j == 3) ? -INFINITY : s3;
        m = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
      }
    }
  }
  if (L.slot != nullptr && L.sc[kKL - 1] > L.published) {  // list full and its tail improved: tell the other lanes
    L.published = L.sc[kKL - 1];
    atomicMax(L.slot, float_to_key(L.published));
  }
}

// kQPU = query blocks per unit.  With kQPU == 2 a unit makes two passes over every corpus tile, one per query block
// (the second pass re-reads the tile from L2, where the first pass just put it) and keeps two candidate lists.  It
// exists to fill the machine: 4 query blocks on 74 SM pairs are 4 x 18 lanes = 72 pairs with kQPU == 1, but
// 2 x 37 lanes = 74 pairs with kQPU == 2, and a tile is shared by 2 units instead of 4.
template <int kCG, int kKL, int kQPU, bool kDebug>
__global__ void __launch_bounds__(kScanThreads, 1)
sa_scan_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c,
               const ScanParams p) {
  using Cfg = ScanCfg<kCG>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kRowsPerQb = kBlockM * kCG;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  float* icbuf = reinterpret_cast<float*>(smem_gen + kStages * Cfg::kStageBytes);  // [2][256]
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
  const int nslots = (p.nqb + kQPU - 1) / kQPU;  // units per tile lane
  const int slot = p.unit_map == 0 ? unit % nslots : unit / TL;
  const int tl = p.unit_map == 0 ? unit / nslots : unit % TL;
  const int qb0 = slot * kQPU;                    // first query block of this unit
  const int npass = min(kQPU, p.nqb - qb0);       // passes over each tile (the last slot of an odd nqb has one)

  // ------------------------------------------------------------------ one-time setup
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
      // Drift control between the units of a tile lane.  Units that share a corpus tile run identical work but at
      // slightly different speeds (measured: ~3 % spread), so over thousands of tiles they drift tens of tiles
      // apart; once the spread exceeds what L2 holds, every unit re-reads its tiles from HBM (measured 3.05x the
      // algorithmic bytes at B = 1024).  A hard barrier costs a pipeline drain per tile (measured +20 %), so the
      // leader producers *pace* themselves instead: each publishes how many tiles it has issued, reads its
      // lane-mates' counters once per tile, and a unit that leads the slowest mate by more than `max_drift` tiles
      // delays every K-slice issue by pace_gain cycles per extra tile of lead (capped).  The kernel's duration is
      // set by its slowest unit anyway, so slowing the fast ones is free; it is only a hint (no waiting on
      // anyone), hence no co-residency assumption and no deadlock.
      const bool lockstep = p.lane_progress != nullptr && p.pace_gain > 0 && nslots > 1 && rank == 0;
      int pace = 0;
      int tile_no = 0;
      for (int t = tl; t < p.num_tiles; t += TL, ++tile_no) {
        if (lockstep) {
          const int* pr = p.lane_progress + tl * nslots;
          int slowest = tile_no;
          for (int j = 0; j < nslots; ++j) slowest = min(slowest, ld_relaxed_gpu(pr + j));
          pace = min(max(tile_no - slowest - p.max_drift, 0) * p.pace_gain, p.pace_max);
        }
        for (int s = 0; s < npass; ++s) {
          const int q_row = (qb0 + s) * kRowsPerQb + static_cast<int>(rank) * kBlockM;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
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
        }
        if (lockstep) st_relaxed_gpu(p.lane_progress + tl * nslots + slot, tile_no + 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader CTA of a pair) =====
      constexpr uint32_t idesc = make_idesc_bf16_f32(kBlockM * kCG, kBlockN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = tl; t < p.num_tiles; t += TL) {
        for (int s = 0; s < npass; ++s, ++it) {
          const int a = it & 1;
          const uint32_t aph = (it >> 1) & 1u;
          mbar_wait(tempty_bar(a), aph ^ 1u);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a * kBlockN);
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(full_bar(stage), phase);
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
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread == query row; 4 warps cover the 128 TMEM lanes =====
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    const int et = ew * 32 + lane;
    // Threshold sharing.  A thread's list only ever sees its own tile lane, so alone it needs ~kKL*ln(n) insertions
    // to warm up, and a warp pays for every lane's insertions.  But if ANY lane already holds kKL rows scoring >= x
    // for this query, no row scoring < x can be in the query's global top-kKL.  So each epilogue thread publishes its
    // kKL-th best (atomicMax on an order-preserving key) and reads the shared bound once per accumulator: every lane
    // gets the threshold of the whole machine's progress, and the warm-up tail disappears.  The shared bound admits
    // ties (>=), the thread's own bound stays strict (>), so tie-breaking by row is unchanged.
    auto query_of = [&](int s) { return (qb0 + s) * kRowsPerQb + static_cast<int>(rank) * kBlockM + et; };
    auto shared_slot = [&](int s) -> unsigned* {
      return (p.thr_shared != nullptr && s < npass && query_of(s) < p.nq) ? p.thr_shared + query_of(s) : nullptr;
    };
    TopList<kKL> L0, L1;
    L0.init(shared_slot(0));
    L1.init(kQPU > 1 ? shared_slot(1) : nullptr);

    auto load_ic = [&](int t, float& x0, float& x1) {
      const long long r0 = static_cast<long long>(t) * kBlockN + 2 * et;
      x0 = (r0 < p.n_rows) ? __ldg(p.inv_norm + r0) : 0.f;
      x1 = (r0 + 1 < p.n_rows) ? __ldg(p.inv_norm + r0 + 1) : 0.f;
    };
    float nxt0 = 0.f, nxt1 = 0.f;
    if (tl < p.num_tiles) load_ic(tl, nxt0, nxt1);

    int it = 0;
    for (int t = tl; t < p.num_tiles; t += TL) {
      // Rows past the committed prefix and all-zero rows get a NaN scale: NaN never compares greater than the
      // threshold, so they can not enter a list.
      const float qnan = __int_as_float(0x7fc00000);
      const float2 scale = make_float2(nxt0 > 0.f ? nxt0 : qnan, nxt1 > 0.f ? nxt1 : qnan);
      if (t + TL < p.num_tiles) load_ic(t + TL, nxt0, nxt1);  // prefetch the next tile's inverse norms
      const int row0 = t * kBlockN;
      for (int s = 0; s < npass; ++s, ++it) {
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1u;
        float* ic = icbuf + a * kBlockN;
        reinterpret_cast<float2*>(ic)[et] = scale;
        asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue-only named barrier: ic[] visible

        mbar_wait(tfull_bar(a), aph);
        tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(a * kBlockN);
        const float4* ic4 = reinterpret_cast<const float4*>(ic);
        float* dbg_row = nullptr;
        if constexpr (kDebug) {
          if (p.dbg_dots != nullptr && t == p.dbg_tile)
            dbg_row = p.dbg_dots + static_cast<size_t>(query_of(s)) * kBlockN;
        }
        if (kQPU == 1 || s == 0)
          epilogue_accumulator<kKL, kDebug>(L0, taddr, ic4, row0, dbg_row);
        else
          epilogue_accumulator<kKL, kDebug>(L1, taddr, ic4, row0, dbg_row);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCG == 1)
            mbar_arrive(tempty_bar(a));
          else
            mbar_arrive_cluster(tempty_bar(a), 0);  // the MMA issuer lives in the pair's leader CTA
        }
      }
    }

    // The only global write of the scan: this CTA's candidate list(s) for each of its queries.
    auto write_list = [&](const TopList<kKL>& L, int s) {
      if (s >= npass || query_of(s) >= p.nq) return;
      const size_t o = ((static_cast<size_t>(blockIdx.x) * kBlockM + et) * kQPU + s) * kKL;
      float4* ps = reinterpret_cast<float4*>(p.part_score + o);
      int4* pi = reinterpret_cast<int4*>(p.part_idx + o);
#pragma unroll
      for (int i = 0; i < kKL / 4; ++i) {
        ps[i] = make_float4(L.sc[4 * i], L.sc[4 * i + 1], L.sc[4 * i + 2], L.sc[4 * i + 3]);
        pi[i] = make_int4(L.id[4 * i], L.id[4 * i + 1], L.id[4 * i + 2], L.id[4 * i + 3]);
      }
    };
    write_list(L0, 0);
    if constexpr (kQPU > 1) write_list(L1, 1);
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (p.dbg_times != nullptr && threadIdx.x == 0) p.dbg_times[2 * blockIdx.x + 1] = globaltimer_ns();
  if constexpr (kCG == 2) cluster_sync_all();  // the peer may still be signalling our barriers / reading our smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kCG>(tmem_base, kTmemCols);
  }
}

}  // namespace sa
