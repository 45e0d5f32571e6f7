// sibrar_b200 -- row gather: embedding / embedding-bag / projected-table lookup + L2-normalise + dropout, and backward.
#include <stdarg.h>

#include "common.cuh"

namespace {
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// one-warp-per-row kernels keep NV values per lane in registers: supported widths 64 / 128 / 256 / 768
#define DISPATCH_NV(n_elems, per, ...)                                   \
  do {                                                                   \
    int _nv = (int)(((n_elems) + (per) - 1) / (per));                    \
    if (_nv <= 2) { constexpr int NVv = 2; __VA_ARGS__; }                \
    else if (_nv <= 4) { constexpr int NVv = 4; __VA_ARGS__; }           \
    else if (_nv <= 8) { constexpr int NVv = 8; __VA_ARGS__; }           \
    else { constexpr int NVv = 24; __VA_ARGS__; }                        \
  } while (0)

// ------------------------------------------------------------------------------------------------ row gather
// lane owns elements c = 4*lane + j + 128*i  (j < 4, i < NV4): 16 contiguous bytes per lane, one Philox call per 4.
struct RowCtx {
  int64_t feat_row;
  int kind;
};

template <int NV4>
__device__ __forceinline__ void load_source_row(const sbr_modality_src_t& s, int64_t feat_row, int C, int lane,
                                                float (&x)[NV4 * 4], float& inv_cnt) {
#pragma unroll
  for (int i = 0; i < NV4 * 4; ++i) x[i] = 0.f;
  inv_cnt = 1.f;
  if (feat_row < 0) return;
  if (s.kind == SBR_SRC_TAG) {
    int cnt = 0;
    for (int t = 0; t < s.max_tags; ++t) {
      int32_t tag = __ldg(s.codes + feat_row * s.max_tags + t);
      if (tag == s.pad_id) continue;
      ++cnt;
      const float* w = s.table + (int64_t)tag * C;
#pragma unroll
      for (int i = 0; i < NV4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = 4 * lane + j + 128 * i;
          if (c < C) x[i * 4 + j] += __ldg(w + c);
        }
    }
    inv_cnt = 1.f / (float)max(cnt, 1);
#pragma unroll
    for (int i = 0; i < NV4 * 4; ++i) x[i] *= inv_cnt;
  } else {
    int64_t src_row = (s.kind == SBR_SRC_CATEGORICAL) ? (int64_t)__ldg(s.codes + feat_row) : feat_row;
    const float* w = s.table + src_row * C;
#pragma unroll
    for (int i = 0; i < NV4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = 4 * lane + j + 128 * i;
        if (c < C) x[i * 4 + j] = __ldg(w + c);
      }
  }
}

__device__ __forceinline__ float keep_scale(const uint8_t* keep_mask, int64_t r, int C, int c, float p_drop,
                                            uint64_t seed, uint64_t step, uint4& cache, int& cache_c4) {
  if (p_drop <= 0.f) return 1.f;
  const float sc = 1.f / (1.f - p_drop);
  if (keep_mask != nullptr) return keep_mask[r * C + c] ? sc : 0.f;
  int c4 = c >> 2;
  if (c4 != cache_c4) {
    cache = philox4x32(make_uint4((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)c4, 0x64726f70u),
                       make_uint2((uint32_t)seed ^ (uint32_t)step, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32)));
    cache_c4 = c4;
  }
  uint32_t bits = (c & 3) == 0 ? cache.x : ((c & 3) == 1 ? cache.y : ((c & 3) == 2 ? cache.z : cache.w));
  return u32_to_unit(bits) >= p_drop ? sc : 0.f;
}

template <int NV4>
__global__ void row_gather_fwd_kernel(const sbr_modality_src_t* __restrict__ srcs, int n_mods,
                                      const int64_t* __restrict__ idx, const uint8_t* __restrict__ mods, int64_t N,
                                      int k, int C, int normalize, float p_drop, uint64_t seed,
                                      const int64_t* __restrict__ step_dev, const uint8_t* __restrict__ keep_mask,
                                      bf16* __restrict__ out, int64_t ld_out, float* __restrict__ out_f32,
                                      int64_t ld_f32, int32_t* err_flag) {
  int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= N) return;
  const int lane = threadIdx.x & 31;
  const int m = mods ? (int)mods[r] : 0;
  const sbr_modality_src_t s = srcs[min(m, n_mods - 1)];
  const int64_t e = idx[r / k];
  int64_t feat_row = s.remap ? (int64_t)__ldg(s.remap + e) : e;
  if (feat_row < 0 && lane == 0 && err_flag) atomicExch(err_flag, 1);
  float x[NV4 * 4], inv_cnt;
  load_source_row<NV4>(s, feat_row, C, lane, x, inv_cnt);
  if (normalize) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV4 * 4; ++i) ss += x[i] * x[i];
    ss = warp_sum(ss);
    float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int i = 0; i < NV4 * 4; ++i) x[i] *= inv;
  }
  const uint64_t step = step_dev ? (uint64_t)*step_dev : 0;
  uint4 cache;
  int cache_c4 = -1;
#pragma unroll
  for (int i = 0; i < NV4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = 4 * lane + j + 128 * i;
      if (c < C) {
        float v = x[i * 4 + j] * keep_scale(keep_mask, r, C, c, p_drop, seed, step, cache, cache_c4);
        if (out) out[r * ld_out + c] = __float2bfloat16(v);
        if (out_f32) out_f32[r * ld_f32 + c] = v;
      }
    }
}

template <int NV4>
__global__ void row_gather_bwd_kernel(const sbr_modality_src_t* __restrict__ srcs, int n_mods,
                                      const int64_t* __restrict__ idx, const uint8_t* __restrict__ mods, int64_t N,
                                      int k, int C, int normalize, float p_drop, uint64_t seed,
                                      const int64_t* __restrict__ step_dev, const uint8_t* __restrict__ keep_mask,
                                      const float* __restrict__ dx, int64_t ld_dx) {
  int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= N) return;
  const int lane = threadIdx.x & 31;
  const int m = mods ? (int)mods[r] : 0;
  const sbr_modality_src_t s = srcs[min(m, n_mods - 1)];
  if (s.grad == nullptr) return;
  const int64_t e = idx[r / k];
  int64_t feat_row = s.remap ? (int64_t)__ldg(s.remap + e) : e;
  if (feat_row < 0) return;
  const uint64_t step = step_dev ? (uint64_t)*step_dev : 0;
  float g[NV4 * 4];
  uint4 cache;
  int cache_c4 = -1;
#pragma unroll
  for (int i = 0; i < NV4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = 4 * lane + j + 128 * i;
      g[i * 4 + j] = (c < C) ? dx[r * ld_dx + c] * keep_scale(keep_mask, r, C, c, p_drop, seed, step, cache, cache_c4)
                             : 0.f;
    }
  float inv_cnt = 1.f;
  if (normalize || s.kind == SBR_SRC_TAG) {
    float x[NV4 * 4];
    load_source_row<NV4>(s, feat_row, C, lane, x, inv_cnt);
    if (normalize) {
      float ss = 0.f, dot = 0.f;
#pragma unroll
      for (int i = 0; i < NV4 * 4; ++i) ss += x[i] * x[i];
      ss = warp_sum(ss);
      float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int i = 0; i < NV4 * 4; ++i) {
        x[i] *= inv;  // y
        dot += x[i] * g[i];
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int i = 0; i < NV4 * 4; ++i) g[i] = (g[i] - x[i] * dot) * inv;
    }
  }
  if (s.kind == SBR_SRC_TAG) {
    for (int t = 0; t < s.max_tags; ++t) {
      int32_t tag = __ldg(s.codes + feat_row * s.max_tags + t);
      if (tag == s.pad_id) continue;
      float* w = s.grad + (int64_t)tag * C;
#pragma unroll
      for (int i = 0; i < NV4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = 4 * lane + j + 128 * i;
          if (c < C) atomicAdd(w + c, g[i * 4 + j] * inv_cnt);
        }
    }
  } else {
    int64_t dst_row = (s.kind == SBR_SRC_CATEGORICAL) ? (int64_t)__ldg(s.codes + feat_row) : feat_row;
    float* w = s.grad + dst_row * C;
#pragma unroll
    for (int i = 0; i < NV4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = 4 * lane + j + 128 * i;
        if (c < C) atomicAdd(w + c, g[i * 4 + j]);
      }
  }
}


// ------------------------------------------------------------------------------------------------ segmented backward
// Rows of a batch that hit the same (modality, table row | category) are reduced by ONE warp instead of contending
// with atomics: plan = counting sort of the flat rows by key, then a segment reduce.  L2-normalise backward is linear
// in the incoming gradient, so it is applied once per segment (chunk) to the summed gradient.
__device__ __forceinline__ int64_t row_key(const sbr_modality_src_t& s, int64_t feat_row) {
  if (feat_row < 0) return -1;
  if (s.kind == SBR_SRC_CATEGORICAL) return s.key_base + (int64_t)__ldg(s.codes + feat_row);
  return s.key_base + feat_row;
}

__global__ void plan_count_kernel(const sbr_modality_src_t* __restrict__ srcs, int n_mods,
                                  const int64_t* __restrict__ idx, const uint8_t* __restrict__ mods, int64_t N, int k,
                                  int32_t* __restrict__ counts, int32_t* __restrict__ row_keys) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  const sbr_modality_src_t& s = srcs[min(mods ? (int)mods[r] : 0, n_mods - 1)];
  const int64_t e = idx[r / k];
  const int64_t feat_row = s.remap ? (int64_t)__ldg(s.remap + e) : e;
  const int64_t key = row_key(s, feat_row);
  row_keys[r] = (int32_t)key;
  if (key >= 0) atomicAdd(counts + key, 1);
}

// exclusive scan by one block (n up to a few million keys: n / 1024 iterations)
__global__ void plan_scan_kernel(const int32_t* __restrict__ counts, int64_t n, int32_t* __restrict__ offsets) {
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    const int32_t v = i < n ? counts[i] : 0;
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int32_t w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_sums[lane] = w;  // inclusive
    }
    __syncthreads();
    const int32_t carry = carry_s;
    const int32_t warp_off = warp > 0 ? warp_sums[warp - 1] : 0;
    if (i < n) offsets[i] = carry + warp_off + x - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_off + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = carry_s;
}

__global__ void plan_fill_kernel(const int32_t* __restrict__ row_keys, int64_t N, const int32_t* __restrict__ offsets,
                                 int32_t* __restrict__ cursor, int32_t* __restrict__ perm,
                                 int32_t* __restrict__ sorted_keys) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= N) return;
  const int32_t key = row_keys[r];
  if (key < 0) return;
  const int32_t pos = offsets[key] + atomicAdd(cursor + key, 1);
  perm[pos] = (int32_t)r;
  sorted_keys[pos] = key;
}

// flush the summed gradient of one run of equal keys into the owning source's gradient buffer
template <int NV4>
__device__ __forceinline__ void flush_run(const sbr_modality_src_t* __restrict__ srcs, int n_mods, int32_t key, int C,
                                          int normalize, int lane, float (&g)[NV4 * 4]) {
  int m = 0;
  for (int t = 1; t < n_mods; ++t)
    if ((int64_t)key >= srcs[t].key_base) m = t;
  const sbr_modality_src_t s = srcs[m];
  if (s.grad == nullptr) return;
  const int64_t local = (int64_t)key - s.key_base;  // table row | category | entity row (TAG)
  float inv_cnt = 1.f;
  if (normalize || s.kind == SBR_SRC_TAG) {
    // every row of the run gathered the same source vector x; the L2-normalise backward is linear in the gradient
    float x[NV4 * 4];
    if (s.kind == SBR_SRC_CATEGORICAL) {
      const float* w = s.table + local * C;
#pragma unroll
      for (int i = 0; i < NV4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * lane + j + 128 * i;
          x[i * 4 + j] = c < C ? __ldg(w + c) : 0.f;
        }
    } else {
      load_source_row<NV4>(s, local, C, lane, x, inv_cnt);
    }
    if (normalize) {
      float ss = 0.f, dot = 0.f;
#pragma unroll
      for (int i = 0; i < NV4 * 4; ++i) ss += x[i] * x[i];
      ss = warp_sum(ss);
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int i = 0; i < NV4 * 4; ++i) {
        x[i] *= inv;
        dot += x[i] * g[i];
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int i = 0; i < NV4 * 4; ++i) g[i] = (g[i] - x[i] * dot) * inv;
    }
  }
  if (s.kind == SBR_SRC_TAG) {
    for (int t = 0; t < s.max_tags; ++t) {
      const int32_t tag = __ldg(s.codes + local * s.max_tags + t);
      if (tag == s.pad_id) continue;
      float* w = s.grad + (int64_t)tag * C;
#pragma unroll
      for (int i = 0; i < NV4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * lane + j + 128 * i;
          if (c < C) atomicAdd(w + c, g[i * 4 + j] * inv_cnt);
        }
    }
  } else {
    float* w = s.grad + local * C;
#pragma unroll
    for (int i = 0; i < NV4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 4 * lane + j + 128 * i;
        if (c < C) atomicAdd(w + c, g[i * 4 + j]);
      }
  }
}

// One warp per chunk of `rows_per_warp` consecutive SORTED rows: runs of equal keys are summed in registers and
// flushed once, so a (modality, source row) that occurs n times in the batch costs ~n / run-length atomics instead
// of n, and the work per warp does not depend on how skewed the keys are (a 2-category feature, a popular item).
template <int NV4>
__global__ void seg_reduce_kernel(const sbr_modality_src_t* __restrict__ srcs, int n_mods, int64_t n_keys,
                                  const int32_t* __restrict__ offsets, const int32_t* __restrict__ perm,
                                  const int32_t* __restrict__ sorted_keys, int C, int normalize, float p_drop,
                                  uint64_t seed, const int64_t* __restrict__ step_dev,
                                  const uint8_t* __restrict__ keep_mask, const float* __restrict__ dx, int64_t ld_dx,
                                  int rows_per_warp) {
  const int64_t chunk = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_sorted = offsets[n_keys];  // rows that have a feature row
  const int64_t beg = chunk * rows_per_warp;
  if (beg >= n_sorted) return;
  const int64_t end = min(n_sorted, beg + rows_per_warp);
  const int lane = threadIdx.x & 31;
  const uint64_t step = step_dev ? (uint64_t)*step_dev : 0;
  float g[NV4 * 4];
#pragma unroll
  for (int i = 0; i < NV4 * 4; ++i) g[i] = 0.f;
  int32_t cur_key = __ldg(sorted_keys + beg);
  for (int64_t p = beg; p < end; ++p) {
    const int32_t key = __ldg(sorted_keys + p);
    const int64_t r = __ldg(perm + p);
    if (key != cur_key) {  // warp-uniform
      flush_run<NV4>(srcs, n_mods, cur_key, C, normalize, lane, g);
      cur_key = key;
#pragma unroll
      for (int i = 0; i < NV4 * 4; ++i) g[i] = 0.f;
    }
    uint4 cache;
    int cache_c4 = -1;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c0 = 4 * lane + 128 * i;
      if (c0 + 3 < C && (ld_dx & 3) == 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(dx + r * ld_dx + c0));
        g[i * 4 + 0] += v.x * keep_scale(keep_mask, r, C, c0 + 0, p_drop, seed, step, cache, cache_c4);
        g[i * 4 + 1] += v.y * keep_scale(keep_mask, r, C, c0 + 1, p_drop, seed, step, cache, cache_c4);
        g[i * 4 + 2] += v.z * keep_scale(keep_mask, r, C, c0 + 2, p_drop, seed, step, cache, cache_c4);
        g[i * 4 + 3] += v.w * keep_scale(keep_mask, r, C, c0 + 3, p_drop, seed, step, cache, cache_c4);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + j;
          if (c < C)
            g[i * 4 + j] += dx[r * ld_dx + c] * keep_scale(keep_mask, r, C, c, p_drop, seed, step, cache, cache_c4);
        }
      }
    }
  }
  flush_run<NV4>(srcs, n_mods, cur_key, C, normalize, lane, g);
}

}  // namespace

extern "C" int sbr_row_gather_fwd(const sbr_modality_src_t* srcs_dev, int n_mods, const int64_t* idx,
                                  const uint8_t* mods, int64_t n_idx, int k, int C, int normalize, float p_drop,
                                  uint64_t seed, const int64_t* step_dev, const uint8_t* keep_mask, void* out_bf16,
                                  int64_t ld_out, float* out_f32, int64_t ld_f32, int32_t* err_flag, void* stream) {
  SBR_REQUIRE(srcs_dev && idx && (out_bf16 || out_f32) && n_idx > 0 && k >= 1, "sbr_row_gather_fwd: bad arguments");
  SBR_REQUIRE(C > 0 && C <= 1024, "sbr_row_gather_fwd: C=%d not in [1, 1024]", C);
  SBR_REQUIRE((!out_bf16 || ld_out >= C) && (!out_f32 || ld_f32 >= C), "sbr_row_gather_fwd: output pitch < C");
  SBR_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "sbr_row_gather_fwd: dropout p must be in [0, 1)");
  const int64_t N = n_idx * k;
  DISPATCH_NV(C, 128, {
    constexpr int NV4 = NVv > 8 ? 8 : NVv;
    row_gather_fwd_kernel<NV4><<<cdiv(N, 8), 256, 0, S(stream)>>>(srcs_dev, n_mods, idx, mods, N, k, C, normalize,
                                                                  p_drop, seed, step_dev, keep_mask,
                                                                  reinterpret_cast<bf16*>(out_bf16), ld_out, out_f32,
                                                                  ld_f32, err_flag);
  });
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_row_gather_bwd(const sbr_modality_src_t* srcs_dev, int n_mods, const int64_t* idx,
                                  const uint8_t* mods, int64_t n_idx, int k, int C, int normalize, float p_drop,
                                  uint64_t seed, const int64_t* step_dev, const uint8_t* keep_mask, const float* dx,
                                  int64_t ld_dx, void* stream) {
  SBR_REQUIRE(srcs_dev && idx && dx && n_idx > 0 && k >= 1, "sbr_row_gather_bwd: bad arguments");
  SBR_REQUIRE(C > 0 && C <= 1024 && ld_dx >= C, "sbr_row_gather_bwd: C=%d not in [1, 1024] or ld_dx < C", C);
  const int64_t N = n_idx * k;
  DISPATCH_NV(C, 128, {
    constexpr int NV4 = NVv > 8 ? 8 : NVv;
    row_gather_bwd_kernel<NV4><<<cdiv(N, 8), 256, 0, S(stream)>>>(srcs_dev, n_mods, idx, mods, N, k, C, normalize,
                                                                  p_drop, seed, step_dev, keep_mask, dx, ld_dx);
  });
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}


extern "C" int sbr_gather_plan(const sbr_modality_src_t* srcs_dev, int n_mods, const int64_t* idx, const uint8_t* mods,
                               int64_t n_idx, int k, int64_t n_keys, int32_t* counts, int32_t* offsets,
                               int32_t* cursor, int32_t* row_keys, int32_t* perm, int32_t* sorted_keys, void* stream) {
  SBR_REQUIRE(srcs_dev && idx && counts && offsets && cursor && row_keys && perm && sorted_keys,
              "sbr_gather_plan: null argument");
  SBR_REQUIRE(n_idx > 0 && k >= 1 && n_keys > 0 && n_keys < (1ll << 31), "sbr_gather_plan: bad sizes");
  const int64_t N = n_idx * k;
  SBR_REQUIRE(N < (1ll << 31), "sbr_gather_plan: too many rows");
  SBR_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_keys, S(stream)));
  SBR_CHECK_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * n_keys, S(stream)));
  plan_count_kernel<<<cdiv(N, 256), 256, 0, S(stream)>>>(srcs_dev, n_mods, idx, mods, N, k, counts, row_keys);
  plan_scan_kernel<<<1, 1024, 0, S(stream)>>>(counts, n_keys, offsets);
  plan_fill_kernel<<<cdiv(N, 256), 256, 0, S(stream)>>>(row_keys, N, offsets, cursor, perm, sorted_keys);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_row_gather_bwd_segmented(const sbr_modality_src_t* srcs_dev, int n_mods, int64_t n_keys,
                                            const int32_t* offsets, const int32_t* perm, const int32_t* sorted_keys,
                                            int64_t n_rows, int C, int normalize, float p_drop, uint64_t seed,
                                            const int64_t* step_dev, const uint8_t* keep_mask, const float* dx,
                                            int64_t ld_dx, int rows_per_warp, void* stream) {
  SBR_REQUIRE(srcs_dev && offsets && perm && sorted_keys && dx && n_keys > 0 && n_rows > 0,
              "sbr_row_gather_bwd_segmented: bad arguments");
  SBR_REQUIRE(C > 0 && C <= 1024 && ld_dx >= C, "sbr_row_gather_bwd_segmented: C=%d not in [1, 1024] or ld_dx < C", C);
  SBR_REQUIRE(rows_per_warp >= 1, "sbr_row_gather_bwd_segmented: bad chunking");
  const unsigned blocks = cdiv(cdiv(n_rows, rows_per_warp), 8);
  DISPATCH_NV(C, 128, {
    constexpr int NV4 = NVv > 8 ? 8 : NVv;
    seg_reduce_kernel<NV4><<<blocks, 256, 0, S(stream)>>>(srcs_dev, n_mods, n_keys, offsets, perm, sorted_keys, C,
                                                          normalize, p_drop, seed, step_dev, keep_mask, dx, ld_dx,
                                                          rows_per_warp);
  });
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}
