// sibrar_b200 -- full-catalog evaluation: tiled user x item score GEMM (tcgen05, TMEM accumulators, TMA operand
// staging) fused with seen-item masking and an exact streaming top-k, so the [U, I] score matrix never reaches HBM.
//
// Orientation: ITEMS are the UMMA M dimension (TMEM lanes), USERS the N dimension (TMEM columns).  A CTA keeps the
// bf16 embeddings of NU users resident in shared memory (B operand) and streams 128-item tiles (A operand) through a
// TMA ring; accumulators are double-buffered in TMEM so the MMA of tile t+1 overlaps the epilogue of tile t.
// Epilogue warps (one TMEM lane quarter each): lane = item, register c = user.  A score is a candidate when it beats
// the user's running k-th best (threshold in smem); candidates are found with warp ballots per user column, filtered
// by the seen-item bitmap of the tile (built from the sorted seen CSR with one cursor per user), and appended as
// packed 64-bit keys  (order-preserving score bits << 32 | ~position)  to a per-user list in global scratch.  When a
// list may overflow it is pruned to its exact top-k by one warp (bitwise k-th-largest selection with warp reductions),
// which also raises the threshold.  Larger key == better: higher score first, ties -> LOWEST item position.
//
// Replaces eval/eval.py:216-220 (scores, mask -> -inf) + the top-k inside rmet.calculate (eval/eval.py:99-102), and
// provides the k-way merge for item-split / item-sharded evaluation and the per-user metrics (eval/metrics.py:4-105).
#include "common.cuh"

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

constexpr int TI = 128;  // items per tile (UMMA M)
constexpr int STAGES = 4;
constexpr int ITEM_STAGE_BYTES = TI * 128;
constexpr int MAX_R = 16;  // candidate keys per lane in a prune: cap <= 512

struct TopkParams {
  int64_t U, I;
  int D, k, cap;
  int tiles_per_split, num_item_tiles;
  int32_t item_offset;
  const int64_t* seen_indptr;
  const int32_t* seen_indices;
  unsigned long long* part_keys;  // [n_splits, U, k]
  unsigned long long* cand;       // [gridDim.y, gridDim.x * NU, cap]
};

__device__ __forceinline__ unsigned long long make_key(float score, uint32_t pos) {
  uint32_t u = __float_as_uint(score);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - pos);
}
__device__ __forceinline__ float key_score(unsigned long long key) {
  uint32_t u = (uint32_t)(key >> 32);
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  return __uint_as_float(u);
}
__device__ __forceinline__ int32_t key_pos(unsigned long long key) {
  return (int32_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
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

// prune the candidate list of one user to its top-k (in place); returns the new count and k-th key
__device__ __forceinline__ void prune_list(unsigned long long* list, int count, int k, int lane, int& new_count,
                                           unsigned long long& kth) {
  unsigned long long keys[MAX_R];
#pragma unroll
  for (int i = 0; i < MAX_R; ++i) {
    int pos = lane + 32 * i;
    keys[i] = pos < count ? __ldcg(list + pos) : 0ull;
  }
  kth = warp_kth_largest<MAX_R>(keys, k);  // 0 when count < k: everything survives
  __syncwarp();
  int offset = 0;
#pragma unroll
  for (int i = 0; i < MAX_R; ++i) {
    const bool keep = keys[i] != 0ull && keys[i] >= kth;
    const unsigned bits = __ballot_sync(0xffffffffu, keep);
    if (keep) __stcg(list + offset + __popc(bits & ((1u << lane) - 1u)), keys[i]);
    offset += __popc(bits);
  }
  new_count = offset;
  __syncwarp();
}

template <int NU>
__global__ void __launch_bounds__(192, 1)
topk_scores_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmI, TopkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KC = (p.D + 63) / 64;
  uint8_t* sU = smem;                                  // KC chunks of [NU rows x 128 B]
  uint8_t* sI = sU + (size_t)KC * NU * 128;            // STAGES x [128 rows x 128 B]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sI + STAGES * ITEM_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* ufull_bar = empty_bar + STAGES;
  uint64_t* tfull_bar = ufull_bar + 1;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);
  float* s_thr = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 128);  // [NU], 16-byte aligned
  int* s_cnt = reinterpret_cast<int*>(s_thr + NU);                 // [NU]
  uint32_t* s_bm = reinterpret_cast<uint32_t*>(s_cnt + NU);        // [NU][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t u0 = (int64_t)blockIdx.x * NU;
  const int split = blockIdx.y;
  const int tile_begin = split * p.tiles_per_split;
  const int tile_end = min(p.num_item_tiles, tile_begin + p.tiles_per_split);
  const int n_tiles = tile_end - tile_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmU);
    tma_prefetch_desc(&tmI);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(ufull_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * NU);
  for (int j = threadIdx.x; j < NU; j += blockDim.x) {
    s_thr[j] = (u0 + j < p.U) ? -INFINITY : INFINITY;  // users past the end never collect candidates
    s_cnt[j] = 0;
  }
  if (threadIdx.x == 0) *s_flag = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(ufull_bar, (uint32_t)(KC * NU * 128));
      for (int kc = 0; kc < KC; ++kc) tma_load_2d(sU + (size_t)kc * NU * 128, &tmU, ufull_bar, kc * 64, (int)u0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int row0 = (tile_begin + t) * TI;
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], ITEM_STAGE_BYTES);
          tma_load_2d(sI + s * ITEM_STAGE_BYTES, &tmI, &full_bar[s], kc * 64, row0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TI, NU, 0, 0);
      mbar_wait(ufull_bar, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int acc = t & 1;
        const uint32_t use = (uint32_t)(t >> 1);
        mbar_wait(&tempty_bar[acc], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NU);
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sI + s * ITEM_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sU + (size_t)kc * NU * 128);
          const int ksteps = min(4, (p.D - kc * 64 + 15) / 16);
          for (int k = 0; k < ksteps; ++k) {
            umma_bf16(d_tmem, umma_smem_desc(a_addr + k * 32, 16, 1024), umma_smem_desc(b_addr + k * 32, 16, 1024),
                      idesc, (kc > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 4 warps = 128 threads
    const int q = warp & 3;
    const int et = threadIdx.x - 64;  // 0..127
    constexpr int UPT = NU / 128;     // users whose seen-cursor this thread owns
    int64_t cur[UPT], cend[UPT];
    const int64_t first_item = (int64_t)tile_begin * TI;
#pragma unroll
    for (int w = 0; w < UPT; ++w) {
      const int64_t u = u0 + et + 128 * w;
      cur[w] = cend[w] = 0;
      if (p.seen_indptr != nullptr && u < p.U) {
        int64_t lo = p.seen_indptr[u], hi = p.seen_indptr[u + 1];
        cend[w] = hi;
        while (lo < hi) {  // lower_bound(first item of this CTA's range)
          int64_t mid = (lo + hi) >> 1;
          if ((int64_t)p.seen_indices[mid] < first_item) lo = mid + 1;
          else hi = mid;
        }
        cur[w] = lo;
      }
    }
    unsigned long long* my_cand = p.cand + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)NU * p.cap;

    for (int t = 0; t < n_tiles; ++t) {
      const int acc = t & 1;
      const uint32_t use = (uint32_t)(t >> 1);
      const int64_t item0 = (int64_t)(tile_begin + t) * TI;
      // 1. seen bitmap of this tile: bit (item - item0) of user j lives in s_bm[j][(item - item0) / 32]
#pragma unroll
      for (int w = 0; w < UPT; ++w) {
        const int j = et + 128 * w;
        uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        int64_t c = cur[w];
        while (c < cend[w]) {
          const int64_t it = (int64_t)__ldg(p.seen_indices + c) - item0;
          if (it >= TI) break;
          if (it >= 0) {
            const uint32_t bit = 1u << (it & 31);
            if (it < 32) b0 |= bit;
            else if (it < 64) b1 |= bit;
            else if (it < 96) b2 |= bit;
            else b3 |= bit;
          }
          ++c;
        }
        cur[w] = c;
        *reinterpret_cast<uint4*>(s_bm + 4 * j) = make_uint4(b0, b1, b2, b3);
      }
      named_bar_sync(1, 128);
      // 2. scan the accumulator
      mbar_wait(&tfull_bar[acc], use & 1);
      tc_fence_after();
      const int64_t my_item = item0 + q * 32 + lane;
      const unsigned valid_bits = __ballot_sync(0xffffffffu, my_item < p.I);
      const uint32_t my_pos = (uint32_t)(my_item + p.item_offset);
#pragma unroll 1
      for (int c0 = 0; c0 < NU; c0 += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NU + c0), r);
        float thr[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          float4 v = *reinterpret_cast<const float4*>(s_thr + c0 + c);
          thr[c] = v.x; thr[c + 1] = v.y; thr[c + 2] = v.z; thr[c + 3] = v.w;
        }
        tmem_ld_wait();
        bool any = false;
#pragma unroll
        for (int c = 0; c < 32; ++c) any |= (__uint_as_float(r[c]) > thr[c]);
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float sc = __uint_as_float(r[c]);
            unsigned bits = __ballot_sync(0xffffffffu, sc > thr[c]) & valid_bits;
            if (bits) {
              const int j = c0 + c;
              bits &= ~s_bm[4 * j + q];
              if (bits) {
                const int npass = __popc(bits);
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_cnt[j], npass);
                base = __shfl_sync(0xffffffffu, base, 0);
                if ((bits >> lane) & 1u) {
                  const int pos = base + __popc(bits & ((1u << lane) - 1u));
                  if (pos < p.cap) __stcg(my_cand + (size_t)j * p.cap + pos, make_key(sc, my_pos));
                }
                if (lane == 0) {
                  const int nc = base + npass;
                  if (nc > p.cap - TI || (thr[c] == -INFINITY && nc >= p.k)) *s_flag = 1;
                }
              }
            }
          }
        }
      }
      // 3. all reads of this accumulator and all appends of this tile are done
      tc_fence_before();
      named_bar_sync(1, 128);
      if (et == 0) mbar_arrive(&tempty_bar[acc]);
      const int flagged = *s_flag;
      named_bar_sync(1, 128);
      if (et == 0) *s_flag = 0;
      // 4. prune the lists that could overflow during the next tile (or that can now define a threshold)
      if (flagged) {
        for (int j = q; j < NU; j += 4) {
          const int c = s_cnt[j];
          if (c > p.cap - TI || (s_thr[j] == -INFINITY && c >= p.k)) {
            int nc;
            unsigned long long kth;
            prune_list(my_cand + (size_t)j * p.cap, min(c, p.cap), p.k, lane, nc, kth);
            if (lane == 0) {
              s_cnt[j] = nc;
              if (kth != 0ull) s_thr[j] = key_score(kth);
            }
          }
        }
      }
      // (the bitmap barrier of the next tile orders these smem updates before the next scan)
    }
    // ------------------------------------------------------------------ final: <= k survivors per user -> part_keys
    named_bar_sync(1, 128);
    for (int j = q; j < NU; j += 4) {
      const int64_t u = u0 + j;
      if (u >= p.U) continue;
      int c = min(s_cnt[j], p.cap);
      unsigned long long* list = my_cand + (size_t)j * p.cap;
      if (c > p.k) {
        unsigned long long kth;
        prune_list(list, c, p.k, lane, c, kth);
      }
      unsigned long long* dst = p.part_keys + ((size_t)split * p.U + u) * p.k;
      for (int i = lane; i < p.k; i += 32) dst[i] = i < c ? __ldcg(list + i) : 0ull;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * NU);
  }
}

inline size_t topk_smem_bytes(int NU, int D) {
  const int KC = (D + 63) / 64;
  return (size_t)KC * NU * 128 + STAGES * ITEM_STAGE_BYTES + 256 + (size_t)NU * (4 + 4 + 16) + 1024 + 64;
}
inline int topk_nu(int D) { return D <= 256 ? 256 : 128; }
inline int topk_cap(int k) {
  int cap = ((k + 384 + 31) / 32) * 32;
  return cap > 512 ? 512 : cap;
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
// one thread per user; ks ascending.  out[m][ki][u], m: 0 ndcg, 1 precision, 2 recall, 3 f_score, 4 hitrate
__global__ void metrics_kernel(const int32_t* __restrict__ topk_idx, int64_t U, int k,
                               const int64_t* __restrict__ tgt_indptr, const int32_t* __restrict__ tgt_indices,
                               const int32_t* __restrict__ ks, int n_ks, float* __restrict__ out,
                               int32_t* __restrict__ item_hits, int64_t n_items) {
  const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (u >= U) return;
  const int64_t beg = tgt_indptr[u], end = tgt_indptr[u + 1];
  const float nt = (float)(end - beg);
  float hits = 0.f, dcg = 0.f, idcg = 0.f;
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
      const size_t o = (size_t)ki * U + u;
      out[0 * (size_t)n_ks * U + o] = ndcg;
      out[1 * (size_t)n_ks * U + o] = prec;
      out[2 * (size_t)n_ks * U + o] = rec;
      out[3 * (size_t)n_ks * U + o] = fs;
      out[4 * (size_t)n_ks * U + o] = fminf(hits, 1.f);
      ++ki;
    }
  }
}

template <int NU>
int launch_topk(const CUtensorMap& tmU, const CUtensorMap& tmI, const TopkParams& p, int user_tiles, int n_splits,
                cudaStream_t st) {
  const size_t smem = topk_smem_bytes(NU, p.D);
  static size_t configured = 0;
  if (smem > configured) {
    SBR_CHECK_CUDA(cudaFuncSetAttribute(topk_scores_kernel<NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  topk_scores_kernel<NU><<<dim3(user_tiles, n_splits), 192, smem, st>>>(tmU, tmI, p);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

}  // namespace

extern "C" int sbr_topk_workspace_bytes(int64_t U, int64_t I, int D, int k, int n_splits, int64_t* bytes_out) {
  SBR_REQUIRE(bytes_out && U > 0 && I > 0 && k > 0 && n_splits > 0, "sbr_topk_workspace_bytes: bad arguments");
  const int NU = topk_nu(D);
  const int64_t user_tiles = (U + NU - 1) / NU;
  *bytes_out = user_tiles * NU * (int64_t)n_splits * topk_cap(k) * 8;
  return SBR_OK;
}

extern "C" int sbr_topk_scores_masked(const void* users, int64_t ldu, const void* items, int64_t ldi, int64_t U,
                                      int64_t I, int D, const int64_t* seen_indptr, const int32_t* seen_indices, int k,
                                      int n_splits, int32_t item_offset, uint64_t* part_keys, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
  SBR_REQUIRE(users && items && part_keys && workspace, "sbr_topk_scores_masked: null argument");
  SBR_REQUIRE(U > 0 && I > 0 && U < (1ll << 31) && I < (1ll << 31), "sbr_topk_scores_masked: bad U/I");
  SBR_REQUIRE(D >= 8 && D <= 512 && D % 8 == 0, "sbr_topk_scores_masked: D=%d must be a multiple of 8 in [8, 512]", D);
  SBR_REQUIRE(k >= 1 && k <= 256, "sbr_topk_scores_masked: k=%d not in [1, 256]", k);
  SBR_REQUIRE((seen_indptr == nullptr) == (seen_indices == nullptr), "sbr_topk_scores_masked: half a CSR given");
  const int num_item_tiles = (int)((I + TI - 1) / TI);
  SBR_REQUIRE(n_splits >= 1 && n_splits <= num_item_tiles, "sbr_topk_scores_masked: n_splits=%d not in [1, %d]",
              n_splits, num_item_tiles);
  int64_t need = 0;
  sbr_topk_workspace_bytes(U, I, D, k, n_splits, &need);
  SBR_REQUIRE(workspace_bytes >= need, "sbr_topk_scores_masked: workspace too small (%lld < %lld)",
              (long long)workspace_bytes, (long long)need);
  const int NU = topk_nu(D);
  const int user_tiles = (int)((U + NU - 1) / NU);
  CUtensorMap tmU, tmI;
  int rc = sbr_make_tmap_bf16_2d(&tmU, users, (uint64_t)D, (uint64_t)U, (uint64_t)ldu, 64, (uint32_t)NU);
  if (rc) return rc;
  rc = sbr_make_tmap_bf16_2d(&tmI, items, (uint64_t)D, (uint64_t)I, (uint64_t)ldi, 64, TI);
  if (rc) return rc;
  TopkParams p;
  p.U = U; p.I = I; p.D = D; p.k = k; p.cap = topk_cap(k);
  p.num_item_tiles = num_item_tiles;
  p.tiles_per_split = (num_item_tiles + n_splits - 1) / n_splits;
  SBR_REQUIRE((int64_t)(n_splits - 1) * p.tiles_per_split < num_item_tiles,
              "sbr_topk_scores_masked: n_splits=%d leaves an empty split for %d item tiles", n_splits, num_item_tiles);
  p.item_offset = item_offset;
  p.seen_indptr = seen_indptr;
  p.seen_indices = seen_indices;
  p.part_keys = reinterpret_cast<unsigned long long*>(part_keys);
  p.cand = reinterpret_cast<unsigned long long*>(workspace);
  if (NU == 256) return launch_topk<256>(tmU, tmI, p, user_tiles, n_splits, S(stream));
  return launch_topk<128>(tmU, tmI, p, user_tiles, n_splits, S(stream));
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
