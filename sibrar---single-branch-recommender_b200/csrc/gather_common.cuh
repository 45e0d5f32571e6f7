// sibrar_b200 -- device helpers shared by the row-gather kernels (gather.cu) and the fused gather + SB-MLP kernels
// (mlp_fused.cu): source-row loads in the "group" layout, the dropout hash, group reductions.
#pragma once
#include "common.cuh"

// Dropout keep decision of element (row r, column c): explicit mask, or 16 random bits of the Philox block of the
// 8-column group c / 8 (one Philox call serves 8 elements; keep iff bits >= p * 65536).
__device__ __forceinline__ uint32_t drop_threshold(float p_drop) { return (uint32_t)ceilf(p_drop * 65536.f); }
// 128 random bits of the dropout block (row r, 8-column group c8): four murmur3-finalised words of a counter that
// mixes (r, c8, seed, step) -- a counter-based generator like Philox4x32-10 at a quarter of its instruction count
// (the mask only has to be reproducible between forward and backward and statistically flat).
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint4 philox_group(int64_t r, int c8, uint64_t seed, uint64_t step) {
  const uint32_t k0 = fmix32((uint32_t)seed ^ ((uint32_t)step * 0x9E3779B1u)) ^ (uint32_t)(seed >> 32);
  const uint32_t x = ((uint32_t)r * 0x9E3779B1u) ^ ((uint32_t)(r >> 32) * 0x7FEB352Du) ^ ((uint32_t)c8 * 0x846CA68Bu) ^ k0;
  return make_uint4(fmix32(x), fmix32(x + 0x68E31DA4u), fmix32(x + 0xB5297A4Du), fmix32(x + 0x1B56C4E9u));
}
// per-launch part of the counter (hoisted out of the row loops by the callers that care)
__device__ __forceinline__ uint32_t philox_key0(uint64_t seed, uint64_t step) {
  return fmix32((uint32_t)seed ^ ((uint32_t)step * 0x9E3779B1u)) ^ (uint32_t)(seed >> 32);
}
__device__ __forceinline__ uint4 philox_group_k(int64_t r, int c8, uint32_t k0) {
  const uint32_t x = ((uint32_t)r * 0x9E3779B1u) ^ ((uint32_t)(r >> 32) * 0x7FEB352Du) ^ ((uint32_t)c8 * 0x846CA68Bu) ^ k0;
  return make_uint4(fmix32(x), fmix32(x + 0x68E31DA4u), fmix32(x + 0xB5297A4Du), fmix32(x + 0x1B56C4E9u));
}
// x[j] = keep(j) ? x[j] * sc : 0 for the 8 columns of one dropout block: the same decisions as philox_lane16(blk, j) >=
// thr, taken as two unsigned compares per word (low half: (w << 16) >= thr << 16, high half: w >= thr << 16) whose
// predicates feed the selects directly.  thr_hi = drop_threshold(p) << 16 (p < 1).
__device__ __forceinline__ void dropout8_hash(float (&x)[8], const uint4& blk, uint32_t thr_hi, float sc) {
  const uint32_t w[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    x[2 * t] = ((w[t] << 16) >= thr_hi) ? x[2 * t] * sc : 0.f;
    x[2 * t + 1] = (w[t] >= thr_hi) ? x[2 * t + 1] * sc : 0.f;
  }
}
__device__ __forceinline__ uint32_t philox_lane16(const uint4& blk, int j) {  // j in [0, 8)
  const uint32_t w = (j >> 1) == 0 ? blk.x : ((j >> 1) == 1 ? blk.y : ((j >> 1) == 2 ? blk.z : blk.w));
  return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
}
__device__ __forceinline__ float keep_scale(const uint8_t* keep_mask, int64_t r, int C, int c, float p_drop,
                                            uint64_t seed, uint64_t step, uint4& cache, int& cache_c8) {
  if (p_drop <= 0.f) return 1.f;
  const float sc = 1.f / (1.f - p_drop);
  if (keep_mask != nullptr) return keep_mask[r * C + c] ? sc : 0.f;
  const int c8 = c >> 3;
  if (c8 != cache_c8) {
    cache = philox_group(r, c8, seed, step);
    cache_c8 = c8;
  }
  return philox_lane16(cache, c & 7) >= drop_threshold(p_drop) ? sc : 0.f;
}
// 8-bit keep mask of columns c0 .. c0+7 (c0 % 8 == 0)
__device__ __forceinline__ uint32_t keep8(const uint8_t* keep_mask, int64_t r, int C, int c0, float p_drop,
                                          uint64_t seed, uint64_t step) {
  if (p_drop <= 0.f) return 0xFFu;
  uint32_t m = 0;
  if (keep_mask != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < C && keep_mask[r * C + c0 + j]) m |= 1u << j;
    return m;
  }
  const uint4 blk = philox_group(r, c0 >> 3, seed, step);
  const uint32_t thr = drop_threshold(p_drop);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (philox_lane16(blk, j) >= thr) m |= 1u << j;
  return m;
}

// ------------------------------------------------------------------------------------------------ group kernels
// A "group" of LPR lanes owns one row; lane li of the group owns the 8 contiguous elements c = 8*li + 8*LPR*i + j
// (j < 8, i < NV): 32-byte loads, all lanes busy whatever C is (C = 64 -> 8 lanes per row, 4 rows per warp).
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int LPR>
__device__ __forceinline__ float group_sum_masked(float v, unsigned mask) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

__device__ __forceinline__ void load8(const float* __restrict__ p, int c, int C, bool vec_ok, float (&x)[8]) {
  if (vec_ok && c + 8 <= C) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p + c + 4));
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = (c + j < C) ? __ldg(p + c + j) : 0.f;
  }
}

// x = source row of (modality src, feature row) in the group layout; inv_cnt = 1 / #tags for TAG sources
template <int LPR, int NV>
__device__ __forceinline__ void load_source_row_g(const sbr_modality_src_t& s, int64_t feat_row, int C, int li,
                                                  float (&x)[NV * 8], float& inv_cnt) {
#pragma unroll
  for (int i = 0; i < NV * 8; ++i) x[i] = 0.f;
  inv_cnt = 1.f;
  if (feat_row < 0) return;
  const bool vec_ok = (C & 3) == 0;
  if (s.kind == SBR_SRC_TAG) {
    int cnt = 0;
    for (int t = 0; t < s.max_tags; ++t) {
      const int32_t tag = __ldg(s.codes + feat_row * s.max_tags + t);
      if (tag == s.pad_id) continue;
      ++cnt;
      const float* w = s.table + (int64_t)tag * C;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[8];
        load8(w, 8 * li + 8 * LPR * i, C, vec_ok, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[i * 8 + j] += v[j];
      }
    }
    inv_cnt = 1.f / (float)max(cnt, 1);
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) x[i] *= inv_cnt;
  } else {
    const int64_t src_row = (s.kind == SBR_SRC_CATEGORICAL) ? (int64_t)__ldg(s.codes + feat_row) : feat_row;
    const float* w = s.table + src_row * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float v[8];
      load8(w, 8 * li + 8 * LPR * i, C, vec_ok, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[i * 8 + j] = v[j];
    }
  }
}

