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

