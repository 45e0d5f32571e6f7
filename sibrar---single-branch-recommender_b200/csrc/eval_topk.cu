// sibrar_b200 -- full-catalog evaluation: tiled user x item score GEMM (tcgen05, TMEM accumulators, TMA operand
// staging) fused with seen-item masking and an exact streaming top-k, so the [U, I] score matrix never reaches HBM.
//
// Orientation: USERS are the UMMA M dimension (TMEM lanes), ITEMS the N dimension (TMEM columns).  A CTA keeps the
// bf16 embeddings of 128 * UT users resident IN TENSOR MEMORY (the A operand of tcgen05.mma in TS form, so the MMA
// reads only the item tile from shared memory -- the SS form is bound by the 128 B/clk shared-memory read port at
// M = N = 128) and streams 128-item tiles (B operand) through a TMA ring.  One fp32 accumulator [128 users x 128 items]
// per user tile (UT = 2; the MMA of one user tile overlaps the drain of the other) or two alternating ones (UT = 1).
// Epilogue thread = (user, 64-column slice): it copies its 64 scores TMEM -> registers, hands the accumulator back,
// and compares a 3-input max tree against the user's running threshold (one register) -- ~0.6 instructions per score
// on the common path.  Scores that reach the threshold and are not in the user's seen set (a sorted-CSR cursor turned
// into a 64-bit mask per tile) are appended to the thread's candidate list in global scratch; when a list may
// overflow its warp prunes it to the exact top-k (bisection over the packed 64-bit keys) and raises the threshold.
// A second small kernel merges the lists of a user; sbr_topk_merge merges item splits / ranks.
// Larger key == better: higher score first, ties -> LOWEST item position.
//
// Replaces eval/eval.py:216-220 (scores, mask -> -inf) + the top-k inside rmet.calculate (eval/eval.py:99-102), and
// provides the k-way merge for item-split / item-sharded evaluation and the per-user metrics (eval/metrics.py:4-105).
#include <stdlib.h>

#include "common.cuh"

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

constexpr int TI = 128;         // items per tile (UMMA N)
constexpr int UM = 128;         // users per UMMA (M = TMEM lanes)
constexpr int MAX_STAGES = 12;  // ring of item K-chunks (actual depth chosen from the free shared memory)
constexpr int STAGE_BYTES = TI * 128;
constexpr int W = TI;           // accumulator columns (items) one epilogue thread scans per tile: the whole tile
constexpr int CS = TI / W;      // candidate lists per user (1)
constexpr int NBLK = W / 32;    // 32-column blocks per thread and tile

struct TopkParams {
  int64_t U, I, users_padded, ldu;
  int D, k, cap, prune_at, sbuf_trigger;
  int tiles_per_split, num_item_tiles, stages;
  int32_t item_offset;
  int debug;  // SBR_TOPK_DEBUG (measurement only): 1 = skip the scan, 2 = also skip the TMEM load, 3 = 2 + no TMA,
              // 4 = 2 + no MMA, -1 = common path of the scan only
  const bf16* users;
  const int64_t* seen_indptr;
  const int32_t* seen_indices;
  unsigned long long* part_keys;  // [n_splits, U, k]
  uint2* cand;                    // [n_splits, users_padded, CS, cap]  (raw score bits, local item position)
  int32_t* cand_cnt;              // [n_splits, users_padded, CS]
};

__device__ __forceinline__ uint32_t orderable(uint32_t fbits) {
  return (fbits & 0x80000000u) ? ~fbits : (fbits | 0x80000000u);
}
__device__ __forceinline__ float orderable_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t pos) {
  return ((unsigned long long)orderable(__float_as_uint(score)) << 32) | (unsigned long long)(0xFFFFFFFFu - pos);
}
__device__ __forceinline__ float key_score(unsigned long long key) { return orderable_to_float((uint32_t)(key >> 32)); }
__device__ __forceinline__ int32_t key_pos(unsigned long long key) {
  return (int32_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// k-th largest of the (distinct, non-zero) keys held by the warp, R per lane.  Returns 0 if fewer than k keys.
template <int R>
__device__ __forceinline__ unsigned long long warp_kth_largest(const unsigned long long (&keys)[R], int k) {
  unsigned long long prefix = 0;
#pragma unroll 1
  for (int bit = 63; bit >= 0; --bit) {
    const unsigned long long trial = prefix | (1ull << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) c += (keys[i] >= trial) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) prefix = trial;
  }
  return prefix;
}

// Lower bound `t` of the k-th largest key such that EXACTLY k keys are >= t.  Needs more than k valid (non-zero,
// distinct) keys in the warp.  Bisection over the key bits, starting below the keys' common prefix and stopping as
// soon as the count is exactly k (typically ~10 rounds instead of 64).
template <int R>
__device__ __forceinline__ unsigned long long warp_select_k(const unsigned long long (&keys)[R], int k) {
  uint32_t or_hi = 0, or_lo = 0, and_hi = 0xFFFFFFFFu, and_lo = 0xFFFFFFFFu;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    if (keys[i] != 0ull) {
      or_hi |= (uint32_t)(keys[i] >> 32); or_lo |= (uint32_t)keys[i];
      and_hi &= (uint32_t)(keys[i] >> 32); and_lo &= (uint32_t)keys[i];
    }
  }
  or_hi = __reduce_or_sync(0xffffffffu, or_hi);
  or_lo = __reduce_or_sync(0xffffffffu, or_lo);
  and_hi = __reduce_and_sync(0xffffffffu, and_hi);
  and_lo = __reduce_and_sync(0xffffffffu, and_lo);
  const unsigned long long orv = ((unsigned long long)or_hi << 32) | or_lo;
  const unsigned long long andv = ((unsigned long long)and_hi << 32) | and_lo;
  const unsigned long long diff = orv ^ andv;  // != 0: at least two distinct keys
  const int top = 63 - __clzll((long long)diff);
  unsigned long long prefix = (top == 63) ? 0ull : (andv & ~((2ull << top) - 1ull));
#pragma unroll 1
  for (int bit = top; bit >= 0; --bit) {
    const unsigned long long trial = prefix | (1ull << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < R; ++i) c += (keys[i] >= trial) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) {
      prefix = trial;
      if (c == k) break;
    }
  }
  return prefix;
}

// one candidate list (count entries, all unseen) as packed keys, R per lane
template <int R>
__device__ __forceinline__ void load_list_keys(const uint2* __restrict__ list, int count, int lane,
                                               unsigned long long (&keys)[R]) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int idx = lane + 32 * i;
    unsigned long long key = 0ull;
    if (idx < count) {
      const uint2 e = __ldcg(list + idx);
      key = ((unsigned long long)orderable(e.x) << 32) | (unsigned long long)(0xFFFFFFFFu - e.y);
    }
    keys[i] = key;
  }
}

// Prunes one candidate list in place to its k best entries (all of them if there are at most k).
// Warp-cooperative; returns the new length and the lower bound of the k-th key (0 = no bound yet).
template <int R>
__device__ __forceinline__ void prune_list(uint2* list, int count, int k, int lane, int& new_count,
                                           unsigned long long& bound) {
  unsigned long long keys[R];
  load_list_keys<R>(list, count, lane, keys);
  bound = 0ull;
  if (count > k) bound = warp_select_k<R>(keys, k);
  else if (count == k) {  // everything survives; the smallest key is the bound
    unsigned long long mn = ~0ull;
#pragma unroll
    for (int i = 0; i < R; ++i) if (keys[i] != 0ull && keys[i] < mn) mn = keys[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, mn, o);
      mn = other < mn ? other : mn;
    }
    bound = mn;
  }
  __syncwarp();
  int offset = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const bool keep = keys[i] != 0ull && keys[i] >= bound;
    const unsigned bits = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const uint32_t hi = (uint32_t)(keys[i] >> 32);
      const uint32_t raw = (hi & 0x80000000u) ? (hi & 0x7FFFFFFFu) : ~hi;
      const int dst = offset + __popc(bits & ((1u << lane) - 1u));
      __stcg(list + dst, make_uint2(raw, 0xFFFFFFFFu - (uint32_t)(keys[i] & 0xFFFFFFFFull)));
    }
    offset += __popc(bits);
  }
  new_count = offset;
  __syncwarp();
}

constexpr int ROW_PITCH = 36;  // floats per staged accumulator row (16-byte aligned, not a multiple of 32 banks)

constexpr int POOL = 64;   // staged candidate rows per warp (compact pool shared by the blocks of a tile)
constexpr int SBUF_N = 64; // per-user score buffer in shared memory (threshold maintenance without global memory)

// max of one 32-column block of the accumulator (lane = user): 3-input max tree, ~0.5 instructions per score
__device__ __forceinline__ float block_max(const uint32_t (&r)[32]) {
  float m[11];
#pragma unroll
  for (int g = 0; g < 10; ++g)
    m[g] = fmax3(__uint_as_float(r[3 * g]), __uint_as_float(r[3 * g + 1]), __uint_as_float(r[3 * g + 2]));
  m[10] = fmaxf(__uint_as_float(r[30]), __uint_as_float(r[31]));
  const float a = fmax3(m[0], m[1], m[2]), b = fmax3(m[3], m[4], m[5]), c = fmax3(m[6], m[7], m[8]);
  return fmax3(fmax3(a, b, c), m[9], m[10]);
}

// lanes that hold a candidate copy their 32 scores into consecutive pool slots starting at `base`
__device__ __forceinline__ void stage_rows(const uint32_t (&r)[32], unsigned flagged, bool mine, float* pool, int base,
                                           int lane) {
  if (mine) {
    const int slot = base + __popc(flagged & ((1u << lane) - 1u));
    float4* dst = reinterpret_cast<float4*>(pool + slot * ROW_PITCH);
#pragma unroll
    for (int v = 0; v < 8; ++v)
      dst[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                           __uint_as_float(r[4 * v + 3]));
  }
}

// The staged rows of one block, one user at a time with lane = column: one compare against that user's threshold and
// seen mask, one ballot, one coalesced append to that user's candidate list (and, SBUF, of the scores alone to the
// user's shared-memory buffer that the threshold is refreshed from).
template <bool SBUF>
__device__ __forceinline__ void handle_events(unsigned flagged, int base, float thr, uint32_t pos0, uint32_t skip,
                                              uint2* warp_list, size_t lane_stride, const float* pool,
                                              float* warp_sbuf, int lane, int& cnt, int& scnt) {
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned all = flagged;
  while (flagged) {
    const int L = __ffs(flagged) - 1;
    flagged &= flagged - 1;
    const float v = pool[(base + __popc(all & ((1u << L) - 1u))) * ROW_PITCH + lane];
    const float thrL = __shfl_sync(0xffffffffu, thr, L);
    const uint32_t skipL = __shfl_sync(0xffffffffu, skip, L);
    const int cntL = __shfl_sync(0xffffffffu, cnt, L);
    const bool pass = v >= thrL && !((skipL >> lane) & 1u);
    const unsigned bits = __ballot_sync(0xffffffffu, pass);
    const int rank = __popc(bits & lt_mask), n = __popc(bits);
    if (pass) __stcg(warp_list + L * lane_stride + cntL + rank, make_uint2(__float_as_uint(v), pos0 + (uint32_t)lane));
    if (SBUF) {
      const int scL = __shfl_sync(0xffffffffu, scnt, L);
      if (pass && scL + rank < SBUF_N) warp_sbuf[L * SBUF_N + scL + rank] = v;  // dropping is safe: bound of a subset
      if (lane == L) scnt = min(SBUF_N, scL + n);
    }
    if (lane == L) cnt = cntL + n;
  }
}

// Threshold refresh from one user's score buffer (n <= 64 scores in shared memory, two per lane): bisection for a
// value t with at least k scores >= t (exactly k unless scores tie), survivors compacted to the front.
__device__ __forceinline__ void refresh_threshold(float* sbuf, int n, int k, int lane, int& new_n, uint32_t& bound) {
  const uint32_t x0 = lane < n ? orderable(__float_as_uint(sbuf[lane])) : 0u;
  const uint32_t x1 = lane + 32 < n ? orderable(__float_as_uint(sbuf[lane + 32])) : 0u;
  bound = 0u;
  new_n = n;
  if (n <= k) return;
  const uint32_t orv = __reduce_or_sync(0xffffffffu, x0 | x1);
  const uint32_t andv = __reduce_and_sync(0xffffffffu, (x0 ? x0 : 0xFFFFFFFFu) & (x1 ? x1 : 0xFFFFFFFFu));
  const uint32_t diff = orv ^ andv;
  uint32_t prefix = andv;  // all scores equal: everything survives
  if (diff != 0u) {
    const int top = 31 - __clz((int)diff);
    prefix = top == 31 ? 0u : (andv & ~((2u << top) - 1u));
#pragma unroll 1
    for (int bit = top; bit >= 0; --bit) {
      const uint32_t trial = prefix | (1u << bit);
      const int c = __reduce_add_sync(0xffffffffu, (x0 >= trial ? 1 : 0) + (x1 >= trial ? 1 : 0));
      if (c >= k) {
        prefix = trial;
        if (c == k) break;
      }
    }
  }
  bound = prefix;
  __syncwarp();
  const bool k0 = x0 != 0u && x0 >= prefix, k1 = x1 != 0u && x1 >= prefix;
  const unsigned b0 = __ballot_sync(0xffffffffu, k0), b1 = __ballot_sync(0xffffffffu, k1);
  const float v0 = orderable_to_float(x0), v1 = orderable_to_float(x1);
  __syncwarp();
  if (k0) sbuf[__popc(b0 & ((1u << lane) - 1u))] = v0;
  if (k1) sbuf[__popc(b0) + __popc(b1 & ((1u << lane) - 1u))] = v1;
  new_n = __popc(b0) + __popc(b1);
  __syncwarp();
}

// UT = user tiles of 128 per CTA, R = cap / 32.  Warps [0, 4*UT) = epilogue (thread = one user, all 128 columns of
// the tile), then TMA producer, then MMA issuer.
// "Job" j = one [128 users x 128 items] accumulator: UT == 2: tile j / 2, user tile j % 2;  UT == 1: tile j.
// Job j uses accumulator j % NACC (NACC = 3 when the user operand leaves room in the 512 TMEM columns).
// SBUF (k <= 40): thresholds are refreshed from per-user score buffers in shared memory; the global lists are then
// append-only until they are nearly full.
template <int UT, int R, int NACC, bool SBUF>
__global__ void __launch_bounds__((4 * UT + 2) * 32, 1)
topk_scores_kernel(const __grid_constant__ CUtensorMap tmI, TopkParams p) {
  constexpr int NEW = 4 * UT;  // epilogue warps: 4 TMEM lane quarters per user tile
  constexpr int NUSERS = UT * UM;
  constexpr int CAP = 32 * R;
  constexpr int A_COL0 = NACC * TI;  // first TMEM column of the user operand (accumulators live below it)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KC = (p.D + 63) / 64;
  const int KA = KC * 32;  // TMEM columns of one user tile (two bf16 per column, zero-padded to whole K chunks)
  const int STAGES = p.stages;
  uint8_t* sI = smem;  // STAGES x [128 items x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sI + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;  // [NACC]
  uint64_t* tempty_bar = tfull_bar + 3;          // [NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 3);
  uint32_t* s_thr = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(full_bar) + 256);  // [NUSERS] orderable bits
  float* s_rows = reinterpret_cast<float*>(s_thr + NUSERS);  // [NEW warps][POOL slots][ROW_PITCH] staged candidate rows
  float* s_sbuf = s_rows + NEW * POOL * ROW_PITCH;            // SBUF: [NUSERS][SBUF_N] candidate scores per user

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t u0 = (int64_t)blockIdx.x * NUSERS;
  const int split = blockIdx.y;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(p.num_item_tiles, tile_begin + p.tiles_per_split);
  const int n_tiles = tile_end - tile_begin;
  const int n_jobs = n_tiles * UT;

  if (warp == NEW && lane == 0) {
    tma_prefetch_desc(&tmI);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == NEW + 1) tmem_alloc(tmem_slot, 512);
  for (int j = threadIdx.x; j < NUSERS; j += blockDim.x) s_thr[j] = 0x007FFFFFu;  // orderable(-inf)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- user operand: global bf16 rows -> TMEM (lane = user, column j = elements 2j, 2j+1)
  if (warp < NEW) {
    const int q = warp & 3;
    const int ut = warp >> 2;
    {
      const int64_t u = u0 + ut * UM + q * 32 + lane;
      const uint4* row = reinterpret_cast<const uint4*>(p.users + (u < p.U ? u : 0) * p.ldu);
      for (int c32 = 0; c32 < KA / 32; ++c32) {
        uint32_t w[32];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const int e0 = c32 * 64 + v * 8;  // first bf16 element of this 16-byte vector
          uint4 x = make_uint4(0u, 0u, 0u, 0u);
          if (u < p.U && e0 < p.D) x = __ldg(row + (e0 >> 3));
          w[4 * v] = x.x; w[4 * v + 1] = x.y; w[4 * v + 2] = x.z; w[4 * v + 3] = x.w;
        }
        tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(A_COL0 + ut * KA + c32 * 32), w);
      }
      tmem_st_wait();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == NEW) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected issue)
    if (p.debug != 3) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int row0 = (tile_begin + t) * TI;
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
            tma_load_2d(sI + s * STAGE_BYTES, &tmI, &full_bar[s], kc * 64, row0);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == NEW + 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected issue)
    const uint32_t idesc = umma_idesc_bf16(UM, TI, 0, 0);
    const int last_ksteps = (p.D - (KC - 1) * 64 + 15) / 16;
    const uint32_t sI_addr = smem_u32(sI);
    int s = 0;  // ring position of the current tile's first chunk
    uint32_t ph = 0;
    for (int t = 0; t < n_tiles; ++t) {
      int ss = s;
      uint32_t pp = ph;
#pragma unroll
      for (int ut = 0; ut < UT; ++ut) {
        const int j = t * UT + ut;
        const int a = j % NACC;
        mbar_wait(&tempty_bar[a], (uint32_t)(((j / NACC) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * TI);
        ss = s;
        pp = ph;
        for (int kc = 0; kc < KC; ++kc) {
          if (ut == 0 && p.debug != 3) {  // the second user tile re-uses the chunk the first one waited for
            mbar_wait(&full_bar[ss], pp);
            tc_fence_after();
          }
          const uint32_t b_addr = sI_addr + (uint32_t)(ss * STAGE_BYTES);
          const uint32_t a_tmem = tmem_base + (uint32_t)(A_COL0 + ut * KA + kc * 32);
          const int ksteps = (kc == KC - 1) ? last_ksteps : 4;
          if (elect_one()) {
            if (p.debug != 4) {
              for (int k = 0; k < ksteps; ++k)
                umma_bf16_ts(d_tmem, a_tmem + (uint32_t)(k * 8), umma_smem_desc(b_addr + k * 32, 16, 1024), idesc,
                             (kc > 0 || k > 0) ? 1u : 0u);
            }
            if (ut == UT - 1) umma_commit(&empty_bar[ss]);
          }
          __syncwarp();
          if (++ss == STAGES) { ss = 0; pp ^= 1; }
        }
        if (elect_one()) umma_commit(&tfull_bar[a]);
        __syncwarp();
      }
      s = ss;
      ph = pp;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: lane = user, columns = items
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int ut = warp >> 2;
    const int ul0 = ut * UM + q * 32;  // first user (CTA-local) of this warp
    const int64_t u_warp = u0 + ul0;
    const int64_t u = u_warp + lane;
    uint2* warp_list = p.cand + ((size_t)split * p.users_padded + u_warp) * (size_t)CAP;
    constexpr size_t LANE_STRIDE = (size_t)CAP;
    float thr = (u < p.U) ? -INFINITY : INFINITY;  // users past the end never collect candidates
    int cnt = 0;
    // cursor into the user's sorted seen row
    int64_t cur = 0, cend = 0;
    int32_t next_seen = 0x7FFFFFFF, after_next = 0x7FFFFFFF;  // two entries ahead: the load latency stays hidden
    float* pool = s_rows + (size_t)warp * POOL * ROW_PITCH;
    float* warp_sbuf = s_sbuf + (size_t)ul0 * SBUF_N;
    int scnt = 0;
    if (p.seen_indptr != nullptr && u < p.U) {
      int64_t lo = __ldg(p.seen_indptr + u), hi = __ldg(p.seen_indptr + u + 1);
      cend = hi;
      const int32_t first_item = tile_begin * TI;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(p.seen_indices + mid) < first_item) lo = mid + 1;
        else hi = mid;
      }
      cur = lo;
      if (cur < cend) next_seen = __ldg(p.seen_indices + cur);
      if (cur + 1 < cend) after_next = __ldg(p.seen_indices + cur + 1);
    }
    const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int32_t n_items32 = (int32_t)p.I;
    const int hard_limit = SBUF ? CAP - W : p.prune_at;  // global-list length that forces an exact prune

    for (int j = (UT == 2) ? ut : 0; j < n_jobs; j += UT) {
      const int a = j % NACC;
      const uint32_t use = (uint32_t)(j / NACC);
      const int t = (UT == 2) ? (j >> 1) : j;
      const int32_t item0 = (tile_begin + t) * TI;
      const float thr_scan = p.debug == -1 ? INFINITY : thr;  // (-1: measurement of the common path alone)
      // seen items of this user inside the tile (the cursor only ever moves forward) + columns beyond the catalogue
      uint32_t skip[NBLK];
#pragma unroll
      for (int bq = 0; bq < NBLK; ++bq) skip[bq] = 0u;
      if (__any_sync(0xffffffffu, next_seen < item0 + W)) {
        while (next_seen < item0 + W) {
          const int rel = next_seen - item0;
          if (rel >= 0) {
#pragma unroll
            for (int bq = 0; bq < NBLK; ++bq)
              if ((rel >> 5) == bq) skip[bq] |= 1u << (rel & 31);
          }
          ++cur;
          next_seen = after_next;
          after_next = cur + 1 < cend ? __ldg(p.seen_indices + cur + 1) : 0x7FFFFFFF;
        }
      }
      if (item0 + W > n_items32) {
        const int nvalid = n_items32 - item0;
#pragma unroll
        for (int bq = 0; bq < NBLK; ++bq) {
          const int nv = nvalid - 32 * bq;
          skip[bq] |= nv <= 0 ? 0xFFFFFFFFu : (nv >= 32 ? 0u : (0xFFFFFFFFu << nv));
        }
      }
      mbar_wait(&tfull_bar[a], use & 1);
      tc_fence_after();
      unsigned flg[NBLK];
      int base[NBLK];
#pragma unroll
      for (int bq = 0; bq < NBLK; ++bq) flg[bq] = 0u, base[bq] = 0;
      int pool_used = 0;
      // candidate rows wait in the pool until the whole tile has been compared (then TMEM is released first); if the
      // pool cannot take a block's rows (first tiles of a pass: every score is a candidate) the pending blocks are
      // handled on the spot
      auto flush = [&]() {
        __syncwarp();
#pragma unroll
        for (int bq = 0; bq < NBLK; ++bq) {
          if (flg[bq]) {
            handle_events<SBUF>(flg[bq], base[bq], thr_scan, (uint32_t)(item0 + 32 * bq), skip[bq], warp_list,
                                LANE_STRIDE, pool, warp_sbuf, lane, cnt, scnt);
            flg[bq] = 0u;
          }
        }
        pool_used = 0;
        __syncwarp();
      };
#pragma unroll
      for (int h = 0; h < NBLK / 2; ++h) {
        uint32_t r0[32], r1[32];
        if (p.debug < 2) {
          tmem_ld32(lane_taddr + (uint32_t)(a * TI + 64 * h), r0);
          tmem_ld32(lane_taddr + (uint32_t)(a * TI + 64 * h + 32), r1);
          tmem_ld_wait();
        }
        if (p.debug >= 1) {
          if (p.debug < 2 && r0[0] == 0x7fc12345u && r1[31] == 0x7fc54321u) cnt = 1;  // keep the loads alive
        } else {
          {
            const bool mine = block_max(r0) >= thr_scan;
            const unsigned f = __ballot_sync(0xffffffffu, mine);
            if (f) {
              if (pool_used + __popc(f) > POOL) flush();
              stage_rows(r0, f, mine, pool, pool_used, lane);
              flg[2 * h] = f; base[2 * h] = pool_used; pool_used += __popc(f);
            }
          }
          {
            const bool mine = block_max(r1) >= thr_scan;
            const unsigned f = __ballot_sync(0xffffffffu, mine);
            if (f) {
              if (pool_used + __popc(f) > POOL) flush();
              stage_rows(r1, f, mine, pool, pool_used, lane);
              flg[2 * h + 1] = f; base[2 * h + 1] = pool_used; pool_used += __popc(f);
            }
          }
        }
      }
      // every score of the tile has been compared (candidate rows are staged): hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
      if (pool_used > 0) {
        flush();
        if (SBUF) {  // cheap threshold refresh from the shared-memory score buffers
          unsigned need = __ballot_sync(0xffffffffu, scnt >= p.sbuf_trigger);
          while (need) {
            const int L = __ffs(need) - 1;
            need &= need - 1;
            const int nL = __shfl_sync(0xffffffffu, scnt, L);
            int nn;
            uint32_t bound;
            refresh_threshold(warp_sbuf + L * SBUF_N, nL, p.k, lane, nn, bound);
            if (lane == L) {
              scnt = nn;
              if (bound != 0u) thr = fmaxf(thr, orderable_to_float(bound));
            }
          }
        }
        // exact prune of the global lists that could overflow during the next tile
        unsigned need = __ballot_sync(0xffffffffu, cnt > hard_limit);
        while (need) {
          const int L = __ffs(need) - 1;
          need &= need - 1;
          const int cntL = __shfl_sync(0xffffffffu, cnt, L);
          int nc;
          unsigned long long bound;
          prune_list<R>(warp_list + L * LANE_STRIDE, cntL, p.k, lane, nc, bound);
          if (lane == L) {
            cnt = nc;
            if (bound != 0ull) thr = fmaxf(thr, orderable_to_float((uint32_t)(bound >> 32)));
          }
        }
      }
    }
    if (u < p.U) p.cand_cnt[(size_t)split * p.users_padded + u] = cnt;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NEW + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// one warp per (split, user): the CS candidate lists of the user -> exact top-k -> part_keys (unsorted packed keys
// with GLOBAL positions, 0 = empty slot)
template <int R>
__global__ void topk_finalize_kernel(TopkParams p, int n_splits) {
  const int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)n_splits * p.U) return;
  const int lane = threadIdx.x & 31;
  const int split = (int)(w / p.U);
  const int64_t u = w - (int64_t)split * p.U;
  constexpr int CAP = 32 * R;
  unsigned long long keys[CS * R];
  int total = 0;
#pragma unroll
  for (int cs = 0; cs < CS; ++cs) {
    const size_t li = ((size_t)split * p.users_padded + u) * CS + cs;
    const int c = p.cand_cnt[li];
    total += c;
    unsigned long long part[R];
    load_list_keys<R>(p.cand + li * CAP, c, lane, part);
#pragma unroll
    for (int i = 0; i < R; ++i) keys[cs * R + i] = part[i];
  }
  unsigned long long bound = 0ull;
  if (total > p.k) bound = warp_select_k<CS * R>(keys, p.k);
  unsigned long long* dst = p.part_keys + ((size_t)split * p.U + u) * p.k;
  int offset = 0;
#pragma unroll
  for (int i = 0; i < CS * R; ++i) {
    const bool keep = keys[i] != 0ull && keys[i] >= bound;
    const unsigned bits = __ballot_sync(0xffffffffu, keep);
    if (keep) dst[offset + __popc(bits & ((1u << lane) - 1u))] = keys[i] - (unsigned long long)(uint32_t)p.item_offset;
    offset += __popc(bits);
  }
  for (int i = offset + lane; i < p.k; i += 32) dst[i] = 0ull;
}

inline int topk_ut(int D) { return D <= 256 ? 2 : 1; }
inline int topk_stages(int UT, int D) {
  const int KC = (D + 63) / 64;
  int s = UT == 2 ? 4 : 8;
  const char* e = getenv("SBR_TOPK_STAGES");  // measurement only
  if (e) s = atoi(e);
  const int lo = UT == 2 ? KC + 1 : 2;  // a chunk stays resident until both user tiles have consumed it
  return s < lo ? lo : (s > MAX_STAGES ? MAX_STAGES : s);
}
inline size_t topk_smem_bytes(int UT, int stages) {
  return (size_t)stages * STAGE_BYTES + 256 + (size_t)UT * UM * 4 + (size_t)(4 * UT) * POOL * ROW_PITCH * 4 +
         (size_t)UT * UM * SBUF_N * 4 + 1024 + 64;
}
inline int topk_cap(int k) {
  // k <= 40: thresholds come from shared-memory score buffers and the global list is append-only -> roomy list;
  // larger k: the list itself is pruned at 2k + 28 entries
  int cap = ((k <= 40 && getenv("SBR_TOPK_SBUF")) || k > 48) ? 512 : 256;
  const char* e = getenv("SBR_TOPK_CAP");  // measurement only
  if (e && atoi(e) >= cap) cap = atoi(e);
  return cap;
}

// ------------------------------------------------------------------------------------------------ merge
// one warp per user: L * k packed keys -> exact top-k, sorted descending, decoded to (score, position)
template <int R>
__global__ void topk_merge_kernel(const unsigned long long* __restrict__ keys_in, int L, int64_t U, int k,
                                  float* __restrict__ out_vals, int32_t* __restrict__ out_idx,
                                  unsigned long long* __restrict__ out_keys) {
  const int64_t u = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (u >= U) return;
  const int lane = threadIdx.x & 31;
  const int total = L * k;
  unsigned long long keys[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int pos = lane + 32 * i;
    unsigned long long key = 0ull;
    if (pos < total) {
      const int l = pos / k, r = pos - l * k;
      key = keys_in[((size_t)l * U + u) * k + r];
    }
    keys[i] = key;
  }
  const unsigned long long kth = warp_kth_largest<R>(keys, k);
  // rank of every survivor = number of keys strictly greater (keys are distinct)
  int rank[R];
#pragma unroll
  for (int i = 0; i < R; ++i) rank[i] = 0;
#pragma unroll 1
  for (int src = 0; src < 32; ++src) {
#pragma unroll
    for (int i2 = 0; i2 < R; ++i2) {
      const unsigned long long other = __shfl_sync(0xffffffffu, keys[i2], src);
      if (other < kth || other == 0ull) continue;  // warp-uniform
#pragma unroll
      for (int i = 0; i < R; ++i) rank[i] += (other > keys[i]) ? 1 : 0;
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const unsigned long long mine = keys[i];
    const bool alive = mine != 0ull && mine >= kth;
    if (alive && rank[i] < k) {
      if (out_vals) out_vals[u * k + rank[i]] = key_score(mine);
      if (out_idx) out_idx[u * k + rank[i]] = key_pos(mine);
      if (out_keys) out_keys[u * k + rank[i]] = mine;
    }
  }
  // slots past the number of survivors: (-inf, -1)
  int alive_cnt = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) alive_cnt += (keys[i] != 0ull && keys[i] >= kth) ? 1 : 0;
  alive_cnt = __reduce_add_sync(0xffffffffu, alive_cnt);
  for (int r = alive_cnt + lane; r < k; r += 32) {
    if (out_vals) out_vals[u * k + r] = -INFINITY;
    if (out_idx) out_idx[u * k + r] = -1;
    if (out_keys) out_keys[u * k + r] = 0ull;
  }
}

// ------------------------------------------------------------------------------------------------ metrics
// one thread per user; ks ascending.  out[m][ki][u], m: 0 ndcg, 1 precision, 2 recall, 3 f_score, 4 hitrate, 5 ap, 6 rr
// (ap = sum over hits of precision@rank / min(k, n_targets); rr = 1 / rank of the first hit)
__global__ void metrics_kernel(const int32_t* __restrict__ topk_idx, int64_t U, int k,
                               const int64_t* __restrict__ tgt_indptr, const int32_t* __restrict__ tgt_indices,
                               const int32_t* __restrict__ ks, int n_ks, float* __restrict__ out,
                               int32_t* __restrict__ item_hits, int64_t n_items) {
  const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (u >= U) return;
  const int64_t beg = tgt_indptr[u], end = tgt_indptr[u + 1];
  const float nt = (float)(end - beg);
  float hits = 0.f, dcg = 0.f, idcg = 0.f, ap_sum = 0.f, rr = 0.f;
  int ki = 0;
  for (int r = 0; r < k && ki < n_ks; ++r) {
    const int32_t it = topk_idx[u * k + r];
    const float disc = 1.f / log2f((float)(r + 2));
    if (it >= 0) {
      int64_t lo = beg, hi = end;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (tgt_indices[mid] < it) lo = mid + 1;
        else hi = mid;
      }
      if (lo < end && tgt_indices[lo] == it) {
        hits += 1.f;
        dcg += disc;
        ap_sum += hits / (float)(r + 1);
        if (rr == 0.f) rr = 1.f / (float)(r + 1);
      }
      if (item_hits != nullptr && it < n_items) {
        for (int kk = ki; kk < n_ks; ++kk) item_hits[(int64_t)kk * n_items + it] = 1;
      }
    }
    if ((float)r < nt) idcg += disc;
    while (ki < n_ks && ks[ki] == r + 1) {
      const float kf = (float)(r + 1);
      const float prec = hits / kf;
      const float rec = nt > 0.f ? hits / nt : 0.f;
      const float ndcg = idcg > 0.f ? fminf(dcg / idcg, 1.f) : 0.f;
      const float fs = (prec + rec) > 0.f ? 2.f * prec * rec / (prec + rec) : 0.f;
      const float denom = fminf(nt, kf);
      const size_t o = (size_t)ki * U + u;
      const size_t plane = (size_t)n_ks * U;
      out[0 * plane + o] = ndcg;
      out[1 * plane + o] = prec;
      out[2 * plane + o] = rec;
      out[3 * plane + o] = fs;
      out[4 * plane + o] = fminf(hits, 1.f);
      out[5 * plane + o] = denom > 0.f ? ap_sum / denom : 0.f;
      out[6 * plane + o] = rr;
      ++ki;
    }
  }
}

template <int UT, int R, int NACC, bool SBUF>
int launch_topk_n(const CUtensorMap& tmI, const TopkParams& p, int user_tiles, int n_splits, cudaStream_t st) {
  const size_t smem = topk_smem_bytes(UT, p.stages);
  static size_t configured = 0;
  if (smem > configured) {
    SBR_CHECK_CUDA(cudaFuncSetAttribute(topk_scores_kernel<UT, R, NACC, SBUF>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  topk_scores_kernel<UT, R, NACC, SBUF><<<dim3(user_tiles, n_splits), (4 * UT + 2) * 32, smem, st>>>(tmI, p);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

template <int UT, int R>
int launch_topk(const CUtensorMap& tmI, const TopkParams& p, int user_tiles, int n_splits, cudaStream_t st) {
  // a third accumulator fits next to the user operand when UT * ceil(D / 64) * 32 <= 128 TMEM columns
  const bool three = UT == 2 && p.D <= 128 && !getenv("SBR_TOPK_NACC2");
  int rc;
  const bool sbuf = p.sbuf_trigger > 0;
  if (UT == 2 && three)
    rc = sbuf ? launch_topk_n<UT, R, (UT == 2 ? 3 : 2), true>(tmI, p, user_tiles, n_splits, st)
              : launch_topk_n<UT, R, (UT == 2 ? 3 : 2), false>(tmI, p, user_tiles, n_splits, st);
  else
    rc = sbuf ? launch_topk_n<UT, R, 2, true>(tmI, p, user_tiles, n_splits, st)
              : launch_topk_n<UT, R, 2, false>(tmI, p, user_tiles, n_splits, st);
  if (rc) return rc;
  SBR_LAUNCH_CHECK();
  topk_finalize_kernel<R><<<cdiv((int64_t)n_splits * p.U, 4), 128, 0, st>>>(p, n_splits);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

template <int UT>
int launch_topk_r(const CUtensorMap& tmI, const TopkParams& p, int user_tiles, int n_splits, cudaStream_t st) {
  if (p.cap == 256) return launch_topk<UT, 8>(tmI, p, user_tiles, n_splits, st);
  return launch_topk<UT, 16>(tmI, p, user_tiles, n_splits, st);
}

}  // namespace

extern "C" int sbr_topk_workspace_bytes(int64_t U, int64_t I, int D, int k, int n_splits, int64_t* bytes_out) {
  SBR_REQUIRE(bytes_out && U > 0 && I > 0 && k > 0 && n_splits > 0, "sbr_topk_workspace_bytes: bad arguments");
  const int NU = topk_ut(D) * UM;
  const int64_t users_padded = (U + NU - 1) / NU * NU;
  const int64_t lists = users_padded * n_splits * CS;
  int cap = topk_cap(k);
  if (cap != 256) cap = 512;
  *bytes_out = lists * cap * 8 + lists * 4;
  return SBR_OK;
}

extern "C" int sbr_topk_scores_masked(const void* users, int64_t ldu, const void* items, int64_t ldi, int64_t U,
                                      int64_t I, int D, const int64_t* seen_indptr, const int32_t* seen_indices, int k,
                                      int n_splits, int32_t item_offset, uint64_t* part_keys, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
  SBR_REQUIRE(users && items && part_keys && workspace, "sbr_topk_scores_masked: null argument");
  SBR_REQUIRE(U > 0 && I > 0 && U < (1ll << 31) && I < (1ll << 31), "sbr_topk_scores_masked: bad U/I");
  SBR_REQUIRE(D >= 8 && D <= 512 && D % 8 == 0, "sbr_topk_scores_masked: D=%d must be a multiple of 8 in [8, 512]", D);
  SBR_REQUIRE(ldu % 8 == 0 && ldu >= D && (reinterpret_cast<uintptr_t>(users) & 15) == 0,
              "sbr_topk_scores_masked: user rows must be 16-byte aligned (ldu %% 8 == 0)");
  SBR_REQUIRE(k >= 1 && k <= 256, "sbr_topk_scores_masked: k=%d not in [1, 256]", k);
  SBR_REQUIRE((seen_indptr == nullptr) == (seen_indices == nullptr), "sbr_topk_scores_masked: half a CSR given");
  SBR_REQUIRE(item_offset >= 0 && (int64_t)item_offset + I < (1ll << 32), "sbr_topk_scores_masked: bad item_offset");
  const int num_item_tiles = (int)((I + TI - 1) / TI);
  SBR_REQUIRE(n_splits >= 1 && n_splits <= num_item_tiles, "sbr_topk_scores_masked: n_splits=%d not in [1, %d]",
              n_splits, num_item_tiles);
  int64_t need = 0;
  sbr_topk_workspace_bytes(U, I, D, k, n_splits, &need);
  SBR_REQUIRE(workspace_bytes >= need, "sbr_topk_scores_masked: workspace too small (%lld < %lld)",
              (long long)workspace_bytes, (long long)need);
  const int UT = topk_ut(D);
  const int NU = UT * UM;
  const int user_tiles = (int)((U + NU - 1) / NU);
  CUtensorMap tmI;
  int rc = sbr_make_tmap_bf16_2d(&tmI, items, (uint64_t)D, (uint64_t)I, (uint64_t)ldi, 64, TI);
  if (rc) return rc;
  TopkParams p;
  p.U = U; p.I = I; p.D = D; p.k = k; p.ldu = ldu;
  p.cap = topk_cap(k);
  if (p.cap != 256) p.cap = 512;
  // prune trigger: early enough that the thresholds stay tight (every candidate costs ~30 warp instructions), late
  // enough that a prune (~2 us of one warp) stays rare -- measured optimum ~2k + 28 entries (gpurun_out/topk_v10)
  p.prune_at = 2 * k + 28;
  if (p.prune_at < 48) p.prune_at = 48;
  if (p.prune_at > p.cap - W) p.prune_at = p.cap - W;
  p.sbuf_trigger = 0;
  // (measured: no faster than pruning the list itself at 2k + 28 entries -> off unless SBR_TOPK_SBUF is set)
  if (k <= 40 && getenv("SBR_TOPK_SBUF")) {
    p.sbuf_trigger = k + 16;  // refresh when the buffer holds k survivors + 16 new scores
    const char* e = getenv("SBR_TOPK_TRIG");  // measurement only
    if (e && atoi(e) > k && atoi(e) <= SBUF_N) p.sbuf_trigger = atoi(e);
  }
  {
    const char* e = getenv("SBR_TOPK_PRUNE_AT");  // measurement only
    if (e && atoi(e) >= k && atoi(e) <= p.cap - W) p.prune_at = atoi(e);
  }
  p.users = reinterpret_cast<const bf16*>(users);
  p.users_padded = (int64_t)user_tiles * NU;
  p.num_item_tiles = num_item_tiles;
  p.tiles_per_split = (num_item_tiles + n_splits - 1) / n_splits;
  SBR_REQUIRE((int64_t)(n_splits - 1) * p.tiles_per_split < num_item_tiles,
              "sbr_topk_scores_masked: n_splits=%d leaves an empty split for %d item tiles", n_splits, num_item_tiles);
  p.item_offset = item_offset;
  p.stages = topk_stages(UT, D);
  {
    const char* dbg = getenv("SBR_TOPK_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  p.seen_indptr = seen_indptr;
  p.seen_indices = seen_indices;
  p.part_keys = reinterpret_cast<unsigned long long*>(part_keys);
  p.cand = reinterpret_cast<uint2*>(workspace);
  p.cand_cnt = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(workspace) +
                                          p.users_padded * n_splits * CS * (int64_t)p.cap * 8);
  if (UT == 2) return launch_topk_r<2>(tmI, p, user_tiles, n_splits, S(stream));
  return launch_topk_r<1>(tmI, p, user_tiles, n_splits, S(stream));
}

extern "C" int sbr_topk_merge(const uint64_t* keys, int L, int64_t U, int k, float* out_vals, int32_t* out_idx,
                              uint64_t* out_keys, void* stream) {
  SBR_REQUIRE(keys && L >= 1 && U > 0 && k >= 1, "sbr_topk_merge: bad arguments");
  SBR_REQUIRE((int64_t)L * k <= 1024, "sbr_topk_merge: L*k=%lld exceeds 1024 (merge hierarchically)",
              (long long)L * k);
  const int total = L * k;
  const unsigned long long* kin = reinterpret_cast<const unsigned long long*>(keys);
  unsigned long long* kout = reinterpret_cast<unsigned long long*>(out_keys);
  const unsigned blocks = cdiv(U, 4);
  if (total <= 64) topk_merge_kernel<2><<<blocks, 128, 0, S(stream)>>>(kin, L, U, k, out_vals, out_idx, kout);
  else if (total <= 128) topk_merge_kernel<4><<<blocks, 128, 0, S(stream)>>>(kin, L, U, k, out_vals, out_idx, kout);
  else if (total <= 256) topk_merge_kernel<8><<<blocks, 128, 0, S(stream)>>>(kin, L, U, k, out_vals, out_idx, kout);
  else if (total <= 512) topk_merge_kernel<16><<<blocks, 128, 0, S(stream)>>>(kin, L, U, k, out_vals, out_idx, kout);
  else topk_merge_kernel<32><<<blocks, 128, 0, S(stream)>>>(kin, L, U, k, out_vals, out_idx, kout);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_metrics_at_k(const int32_t* topk_idx, int64_t U, int k, const int64_t* tgt_indptr,
                                const int32_t* tgt_indices, const int32_t* ks_dev, int n_ks, float* out,
                                int32_t* item_hits, int64_t n_items, void* stream) {
  SBR_REQUIRE(topk_idx && tgt_indptr && tgt_indices && ks_dev && out && U > 0 && k >= 1 && n_ks >= 1,
              "sbr_metrics_at_k: bad arguments");
  metrics_kernel<<<cdiv(U, 128), 128, 0, S(stream)>>>(topk_idx, U, k, tgt_indptr, tgt_indices, ks_dev, n_ks, out,
                                                      item_hits, n_items);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}
