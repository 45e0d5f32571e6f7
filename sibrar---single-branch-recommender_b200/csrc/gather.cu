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

#include "gather_common.cuh"

template <int LPR, int NV>
__global__ void __launch_bounds__(256)
row_gather_fwd_g_kernel(const sbr_modality_src_t* __restrict__ srcs, int n_mods, const int64_t* __restrict__ idx,
                        const uint8_t* __restrict__ mods, int64_t N, int k, int C, int normalize, float p_drop,
                        uint64_t seed, const int64_t* __restrict__ step_dev, const uint8_t* __restrict__ keep_mask,
                        bf16* __restrict__ out, int64_t ld_out, float* __restrict__ out_f32, int64_t ld_f32,
                        int32_t* err_flag, uint8_t* __restrict__ keep_bits_out) {
  SBR_PDL_ENTRY();
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;
  const int li = threadIdx.x % LPR;
  const bool valid = gid < N;
  const int64_t r = valid ? gid : N - 1;  // keep every lane alive for the group shuffles
  const int m = mods ? (int)mods[r] : 0;
  const sbr_modality_src_t s = srcs[min(m, n_mods - 1)];
  const int64_t e = idx[r / k];
  const int64_t feat_row = s.remap ? (int64_t)__ldg(s.remap + e) : e;
  if (feat_row < 0 && li == 0 && valid && err_flag) atomicExch(err_flag, 1);
  float x[NV * 8], inv_cnt;
  load_source_row_g<LPR, NV>(s, feat_row, C, li, x, inv_cnt);
  if (normalize) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) ss += x[i] * x[i];
    ss = group_sum<LPR>(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) x[i] *= inv;
  }
  if (!valid) return;
  const uint64_t step = step_dev ? (uint64_t)*step_dev : 0;
  const float sc = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c0 = 8 * li + 8 * LPR * i;
    if (c0 >= C) continue;
    const uint32_t km = keep8(keep_mask, r, C, c0, p_drop, seed, step);
    if (keep_bits_out != nullptr) keep_bits_out[r * ((C + 7) >> 3) + (c0 >> 3)] = (uint8_t)km;  // for the backward
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ((km >> j) & 1u) ? x[i * 8 + j] * sc : 0.f;
    if (out) {
      bf16* dst = out + r * ld_out + c0;
      if (c0 + 8 <= C && (ld_out & 7) == 0) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[2 * t], v[2 * t + 1]);
        *reinterpret_cast<uint4*>(dst) = u;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < C) dst[j] = __float2bfloat16(v[j]);
      }
    }
    if (out_f32) {
      float* dst = out_f32 + r * ld_f32 + c0;
      if (c0 + 8 <= C && (ld_f32 & 3) == 0) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < C) dst[j] = v[j];
      }
    }
  }
}

constexpr int SEG_MAX_MODS = 16;
constexpr int SEG_SMEM_FLOATS = 10240;  // 40 KB of block-private gradient rows for small Embedding / EmbeddingBag tables

struct SegShared {
  sbr_modality_src_t src[SEG_MAX_MODS];
  int smem_off[SEG_MAX_MODS];  // offset of the modality's private gradient table in `priv`, or -1
  int n_rep, rep_stride;       // replicas of the private tables (warp w adds into replica w % n_rep)
  float priv[SEG_SMEM_FLOATS];
};

// 8 consecutive floats added to shared memory (float atomics in shared memory are CAS loops: keep the volume low)
__device__ __forceinline__ void smem_add8(float* w, int c, int C, const float* g, float scale) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (c + j < C) atomicAdd(w + c + j, g[j] * scale);
}

// 8 consecutive floats added to global memory: two 16-byte vector reductions (REDG.E.ADD.F32x4) when aligned
__device__ __forceinline__ void red_add8(float* w, int c, int C, const float* g, float scale) {
  if (c + 8 <= C && ((reinterpret_cast<uintptr_t>(w + c) & 15) == 0)) {
    const size_t a = __cvta_generic_to_global(w + c);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(g[0] * scale), "f"(g[1] * scale),
                 "f"(g[2] * scale), "f"(g[3] * scale)
                 : "memory");
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a + 16), "f"(g[4] * scale), "f"(g[5] * scale),
                 "f"(g[6] * scale), "f"(g[7] * scale)
                 : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c + j < C) atomicAdd(w + c + j, g[j] * scale);
  }
}

// flush the summed gradient of one run of equal keys (group layout) into the owning source's gradient buffer
template <int LPR, int NV>
__device__ __forceinline__ void flush_run_g(SegShared& sh, int n_mods, int32_t key, int C, int normalize, int li,
                                            unsigned gmask, float (&g)[NV * 8]) {
  int m = 0;
  for (int t = 1; t < n_mods; ++t)
    if ((int64_t)key >= sh.src[t].key_base) m = t;
  const sbr_modality_src_t& s = sh.src[m];
  if (s.grad == nullptr) return;
  const int64_t local = (int64_t)key - s.key_base;  // table row | category | entity row (TAG)
  float inv_cnt = 1.f;
  if (normalize || s.kind == SBR_SRC_TAG) {
    // every row of the run gathered the same source vector x; the L2-normalise backward is linear in the gradient
    float x[NV * 8];
    if (s.kind == SBR_SRC_CATEGORICAL) {
      const float* w = s.table + local * C;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[8];
        load8(w, 8 * li + 8 * LPR * i, C, (C & 3) == 0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[i * 8 + j] = v[j];
      }
    } else if (normalize) {
      load_source_row_g<LPR, NV>(s, local, C, li, x, inv_cnt);
    } else {  // TAG without normalisation: only the tag count is needed
      int cnt = 0;
      for (int t = 0; t < s.max_tags; ++t) cnt += __ldg(s.codes + local * s.max_tags + t) != s.pad_id ? 1 : 0;
      inv_cnt = 1.f / (float)max(cnt, 1);
    }
    if (normalize) {
      float ss = 0.f, dot = 0.f;
#pragma unroll
      for (int i = 0; i < NV * 8; ++i) ss += x[i] * x[i];
      ss = group_sum_masked<LPR>(ss, gmask);
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
      for (int i = 0; i < NV * 8; ++i) {
        x[i] *= inv;
        dot += x[i] * g[i];
      }
      dot = group_sum_masked<LPR>(dot, gmask);
#pragma unroll
      for (int i = 0; i < NV * 8; ++i) g[i] = (g[i] - x[i] * dot) * inv;
    }
  }
  const int off = sh.smem_off[m];
  // shared-memory float atomics are compare-and-swap loops: the replicas keep the groups of different warps apart
  float* priv = off >= 0 ? sh.priv + ((threadIdx.x >> 5) % sh.n_rep) * sh.rep_stride + off : nullptr;
  if (s.kind == SBR_SRC_TAG) {
    for (int t = 0; t < s.max_tags; ++t) {
      const int32_t tag = __ldg(s.codes + local * s.max_tags + t);
      if (tag == s.pad_id) continue;
      if (priv != nullptr) {
#pragma unroll
        for (int i = 0; i < NV; ++i) smem_add8(priv + (int64_t)tag * C, 8 * li + 8 * LPR * i, C, g + i * 8, inv_cnt);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) red_add8(s.grad + (int64_t)tag * C, 8 * li + 8 * LPR * i, C, g + i * 8, inv_cnt);
      }
    }
  } else if (priv != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem_add8(priv + local * C, 8 * li + 8 * LPR * i, C, g + i * 8, 1.f);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) red_add8(s.grad + local * C, 8 * li + 8 * LPR * i, C, g + i * 8, 1.f);
  }
}

// Persistent blocks; one GROUP per chunk of `rows_per_group` consecutive SORTED rows (grid-stride over the chunks):
// runs of equal keys are summed in registers and flushed once, so a (modality, source row) that occurs n times in the
// batch costs ~n / run-length atomics instead of n, and the work per group does not depend on how skewed the keys
// are.  Small Embedding / EmbeddingBag tables (a 2-category feature, 18 genre tags) are accumulated in a block-private
// shared-memory copy first: thousands of same-address global atomics serialise at ~100 cycles each in one L2 slice.
template <int LPR, int NV, int MINB>
__global__ void __launch_bounds__(256, MINB)
seg_reduce_g_kernel(const sbr_modality_src_t* __restrict__ srcs, int n_mods, int64_t n_keys,
                    const int32_t* __restrict__ offsets, const int32_t* __restrict__ perm,
                    const int32_t* __restrict__ sorted_keys, int C, int normalize, float p_drop, uint64_t seed,
                    const int64_t* __restrict__ step_dev, const uint8_t* __restrict__ keep_mask,
                    const float* __restrict__ dx, int64_t ld_dx, const uint8_t* __restrict__ keep_bits, int rpg) {
  SBR_PDL_LAUNCH();  // the shared-memory set-up below overlaps the previous kernel (srcs is a set-up-time constant)
  __shared__ SegShared sh;
  if (threadIdx.x == 0) {
    int used = 0;
    for (int m = 0; m < n_mods; ++m) {
      sh.src[m] = srcs[m];
      const int64_t need = sh.src[m].n_table_rows * C;
      const bool small = sh.src[m].grad != nullptr && sh.src[m].kind != SBR_SRC_TABLE && need > 0 &&
                         used + need <= SEG_SMEM_FLOATS;
      sh.smem_off[m] = small ? used : -1;
      if (small) used += (int)need;
    }
    sh.rep_stride = used;
    sh.n_rep = used > 0 ? max(1, min((int)(blockDim.x >> 5), SEG_SMEM_FLOATS / used)) : 1;
  }
  for (int i = threadIdx.x; i < SEG_SMEM_FLOATS; i += blockDim.x) sh.priv[i] = 0.f;
  __syncthreads();
  SBR_PDL_WAIT();

  const int li = threadIdx.x % LPR;
  const int lane = threadIdx.x & 31;
  const unsigned gmask = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << (lane / LPR * LPR));
  const int64_t n_sorted = offsets[n_keys];  // rows that have a feature row
  const uint64_t step = step_dev ? (uint64_t)*step_dev : 0;
  const float sc = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  const bool vec_ok = (ld_dx & 3) == 0;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / LPR);
  const int RPG = rpg;  // sorted rows per chunk
  for (int64_t chunk = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;; chunk += groups_total) {
    const int64_t beg = chunk * RPG;
    if (beg >= n_sorted) break;
    const int64_t end = min(n_sorted, beg + RPG);
    float g[NV * 8];
#pragma unroll
    for (int i = 0; i < NV * 8; ++i) g[i] = 0.f;
    int32_t cur_key = __ldg(sorted_keys + beg);
    // software pipeline: indices two rows ahead, the dx row one row ahead of the row being summed
    int32_t key_c = cur_key, key_n = cur_key;
    int64_t r_c = __ldg(perm + beg), r_n = r_c;
    if (beg + 1 < end) {
      key_n = __ldg(sorted_keys + beg + 1);
      r_n = __ldg(perm + beg + 1);
    }
    float vn[NV][8];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = 8 * li + 8 * LPR * i;
      if (c0 < C) load8(dx + r_c * ld_dx, c0, C, vec_ok, vn[i]);
    }
#pragma unroll 1
    for (int64_t p = beg; p < end; ++p) {
      const int32_t key = key_c;
      const int64_t r = r_c;
      float v[NV][8];
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = vn[i][j];
      key_c = key_n;
      r_c = r_n;
      if (p + 1 < end) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c0 = 8 * li + 8 * LPR * i;
          if (c0 < C) load8(dx + r_c * ld_dx, c0, C, vec_ok, vn[i]);
        }
        if (p + 2 < end) {
          key_n = __ldg(sorted_keys + p + 2);
          r_n = __ldg(perm + p + 2);
        }
      }
      if (key != cur_key) {  // group-uniform
        flush_run_g<LPR, NV>(sh, n_mods, cur_key, C, normalize, li, gmask, g);
        cur_key = key;
#pragma unroll
        for (int i = 0; i < NV * 8; ++i) g[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c0 = 8 * li + 8 * LPR * i;
        if (c0 >= C) continue;
        // the forward's keep bits (1 byte per lane and row) when it stored them, else the mask is regenerated
        const uint32_t km = keep_bits != nullptr ? (uint32_t)__ldg(keep_bits + r * ((C + 7) >> 3) + (c0 >> 3))
                                                 : keep8(keep_mask, r, C, c0, p_drop, seed, step);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[i * 8 + j] += ((km >> j) & 1u) ? v[i][j] * sc : 0.f;
      }
    }
    flush_run_g<LPR, NV>(sh, n_mods, cur_key, C, normalize, li, gmask, g);
  }
  __syncthreads();
  // block-private small tables -> global gradient buffers
  for (int m = 0; m < n_mods; ++m) {
    const int off = sh.smem_off[m];
    if (off < 0) continue;
    const int n = (int)(sh.src[m].n_table_rows * C);
    float* dst = sh.src[m].grad;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float v = 0.f;
      for (int rep = 0; rep < sh.n_rep; ++rep) v += sh.priv[rep * sh.rep_stride + off + i];
      if (v != 0.f) atomicAdd(dst + i, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------ EmbeddingBag tables
// nn.EmbeddingBag(mode="mean", padding_idx=pad) of EVERY feature row as a dense table (reference FeatureEmbedding for
// tag features, sgd_alg.py:1279-1396): bag[r] = mean over the non-pad tags t of weight[codes[r, t]].  The model
// gathers from this table like from a projected one (one row per lookup, no tag loop in the big gathers) and
// scatters the table gradient back onto the tag embeddings with the kernel below.
template <int LPR, int NV>
__global__ void __launch_bounds__(256)
tag_bag_fwd_kernel(const int32_t* __restrict__ codes, int max_tags, int32_t pad_id, const float* __restrict__ weight,
                   int64_t n_rows, int C, float* __restrict__ out) {
  SBR_PDL_ENTRY();
  const int64_t gid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR;
  const int li = threadIdx.x % LPR;
  if (gid >= n_rows) return;
  sbr_modality_src_t s;
  s.kind = SBR_SRC_TAG;
  s.table = weight;
  s.codes = codes;
  s.max_tags = max_tags;
  s.pad_id = pad_id;
  float x[NV * 8], inv_cnt;
  load_source_row_g<LPR, NV>(s, gid, C, li, x, inv_cnt);  // the arithmetic of the direct TAG gather, bit for bit
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 8 * li + 8 * LPR * i + j;
      if (c < C) out[gid * C + c] = x[i * 8 + j];
    }
}

// grad_weight[tag] += bag_grad[r] / #tags(r) for every tag of every feature row; bag_grad is cleared as it is read.
// Block-private shared-memory copy of the (small) tag table first, one global atomic per touched element and block.
template <int LPR, int NV>
__global__ void __launch_bounds__(256)
tag_bag_bwd_kernel(const int32_t* __restrict__ codes, int max_tags, int32_t pad_id, float* __restrict__ bag_grad,
                   int64_t n_rows, int C, float* __restrict__ grad_weight, int64_t n_weight_rows) {
  SBR_PDL_LAUNCH();
  __shared__ float priv[SEG_SMEM_FLOATS];
  const bool use_smem = n_weight_rows * C <= SEG_SMEM_FLOATS;
  const int n_priv = use_smem ? (int)(n_weight_rows * C) : 0;
  for (int i = threadIdx.x; i < n_priv; i += blockDim.x) priv[i] = 0.f;
  __syncthreads();
  SBR_PDL_WAIT();
  const int li = threadIdx.x % LPR;
  const int64_t groups_total = (int64_t)gridDim.x * (blockDim.x / LPR);
  for (int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / LPR; r < n_rows; r += groups_total) {
    float g[NV * 8];
    bool any = false;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = 8 * li + 8 * LPR * i + j;
        g[i * 8 + j] = c < C ? bag_grad[r * C + c] : 0.f;
        any |= g[i * 8 + j] != 0.f;
        if (c < C) bag_grad[r * C + c] = 0.f;
      }
    if (!any) continue;  // (per lane: a lane whose 8 elements are zero has nothing to add)
    int cnt = 0;
    for (int t = 0; t < max_tags; ++t) cnt += __ldg(codes + r * max_tags + t) != pad_id ? 1 : 0;
    const float inv_cnt = 1.f / (float)max(cnt, 1);
    for (int t = 0; t < max_tags; ++t) {
      const int32_t tag = __ldg(codes + r * max_tags + t);
      if (tag == pad_id) continue;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (use_smem) smem_add8(priv + (int64_t)tag * C, 8 * li + 8 * LPR * i, C, g + i * 8, inv_cnt);
        else red_add8(grad_weight + (int64_t)tag * C, 8 * li + 8 * LPR * i, C, g + i * 8, inv_cnt);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_priv; i += blockDim.x) {
    const float v = priv[i];
    if (v != 0.f) atomicAdd(grad_weight + i, v);
  }
}

// lanes per row / 8-element vectors per lane for a row of C elements
#define DISPATCH_GROUP(C_, ...)                                                        \
  do {                                                                                 \
    if ((C_) <= 8) { constexpr int LPRv = 1, NVg = 1; __VA_ARGS__; }                   \
    else if ((C_) <= 16) { constexpr int LPRv = 2, NVg = 1; __VA_ARGS__; }             \
    else if ((C_) <= 32) { constexpr int LPRv = 4, NVg = 1; __VA_ARGS__; }             \
    else if ((C_) <= 64) { constexpr int LPRv = 8, NVg = 1; __VA_ARGS__; }             \
    else if ((C_) <= 128) { constexpr int LPRv = 16, NVg = 1; __VA_ARGS__; }           \
    else if ((C_) <= 256) { constexpr int LPRv = 32, NVg = 1; __VA_ARGS__; }           \
    else if ((C_) <= 512) { constexpr int LPRv = 32, NVg = 2; __VA_ARGS__; }           \
    else { constexpr int LPRv = 32, NVg = 4; __VA_ARGS__; }                            \
  } while (0)

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

// Plan = counting sort in three launches, no memsets: `counts` is zero on entry (allocation / the previous call's scan
// clears it), the scan also seeds `cursor` with the segment starts, and the per-chunk scan state lives behind
// offsets[n_keys] (cleared by the count kernel).  Same-key slots of a warp are combined before the atomic (a 2-row
// gender table otherwise takes one same-address atomic per batch row).
constexpr int SCAN_CHUNK = 4096;  // keys per block of the scan (1024 threads x 4)

__global__ void __launch_bounds__(256)
plan_count_kernel(const sbr_modality_src_t* __restrict__ srcs, int n_mods, const int64_t* __restrict__ idx,
                  const uint8_t* __restrict__ mods, int64_t N, int k, int32_t* __restrict__ counts,
                  int32_t* __restrict__ row_keys, int32_t* __restrict__ scan_state, int n_state) {
  SBR_PDL_ENTRY();
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < n_state) scan_state[r] = 0;
  int32_t key = -1;
  if (r < N) {
    const sbr_modality_src_t& s = srcs[min(mods ? (int)mods[r] : 0, n_mods - 1)];
    const int64_t e = idx[r / k];
    const int64_t feat_row = s.remap ? (int64_t)__ldg(s.remap + e) : e;
    key = (int32_t)row_key(s, feat_row);
    row_keys[r] = key;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counts + key, __popc(peers));
}

// exclusive scan of counts[0, n) -> offsets[0, n] and cursor[0, n); counts is cleared.  One block per chunk of 4096
// keys; a block publishes its chunk total (bit 31 = valid) and adds up the totals of the chunks in front of it
// (in-order block dispatch: those blocks are running or done).
__global__ void __launch_bounds__(1024)
plan_scan_kernel(int32_t* __restrict__ counts, int64_t n, int32_t* __restrict__ offsets, int32_t* __restrict__ cursor,
                 volatile int32_t* __restrict__ scan_state) {
  SBR_PDL_ENTRY();
  __shared__ int32_t warp_sums[32];
  __shared__ int32_t s_prefix;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + 4 * threadIdx.x;
  int32_t c[4] = {0, 0, 0, 0};
  if (base + 4 <= n && (reinterpret_cast<uintptr_t>(counts) & 15) == 0) {
    const int4 v = *reinterpret_cast<const int4*>(counts + base);
    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    *reinterpret_cast<int4*>(counts + base) = make_int4(0, 0, 0, 0);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (base + j < n) {
        c[j] = counts[base + j];
        counts[base + j] = 0;
      }
  }
  const int32_t local = c[0] + c[1] + c[2] + c[3];
  int32_t x = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int32_t w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_sums[lane] = w;  // inclusive
    if (lane == 31) {
      __threadfence();
      scan_state[blockIdx.x] = (int32_t)(0x80000000u | (uint32_t)w);  // this chunk's total
    }
    // totals of the chunks in front (lanes stride over them)
    int32_t pre = 0;
    for (int b = lane; b < (int)blockIdx.x; b += 32) {
      int32_t v;
      do { v = scan_state[b]; } while (v >= 0);
      pre += (int32_t)((uint32_t)v & 0x7fffffffu);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
    if (lane == 0) s_prefix = pre;
  }
  __syncthreads();
  int32_t run = s_prefix + (warp > 0 ? warp_sums[warp - 1] : 0) + x - local;  // exclusive prefix of this thread's 4 keys
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (base + j < n) {
      offsets[base + j] = run;
      cursor[base + j] = run;
    }
    run += c[j];
    if (base + j == n - 1) offsets[n] = run;
  }
}

__global__ void __launch_bounds__(256)
plan_fill_kernel(const int32_t* __restrict__ row_keys, int64_t N, int32_t* __restrict__ cursor,
                 int32_t* __restrict__ perm, int32_t* __restrict__ sorted_keys) {
  SBR_PDL_ENTRY();
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int32_t key = r < N ? row_keys[r] : -1;
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  const int leader = __ffs(peers) - 1;
  int32_t first = 0;
  if (key >= 0 && lane == leader) first = atomicAdd(cursor + key, __popc(peers));
  first = __shfl_sync(0xffffffffu, first, leader);
  if (key < 0) return;
  const int32_t pos = first + __popc(peers & ((1u << lane) - 1u));
  perm[pos] = (int32_t)r;
  sorted_keys[pos] = key;
}

}  // namespace

extern "C" int sbr_row_gather_fwd(const sbr_modality_src_t* srcs_dev, int n_mods, const int64_t* idx,
                                  const uint8_t* mods, int64_t n_idx, int k, int C, int normalize, float p_drop,
                                  uint64_t seed, const int64_t* step_dev, const uint8_t* keep_mask, void* out_bf16,
                                  int64_t ld_out, float* out_f32, int64_t ld_f32, int32_t* err_flag,
                                  uint8_t* keep_bits_out, void* stream) {
  SBR_REQUIRE(srcs_dev && idx && (out_bf16 || out_f32) && n_idx > 0 && k >= 1, "sbr_row_gather_fwd: bad arguments");
  SBR_REQUIRE(C > 0 && C <= 1024, "sbr_row_gather_fwd: C=%d not in [1, 1024]", C);
  SBR_REQUIRE((!out_bf16 || ld_out >= C) && (!out_f32 || ld_f32 >= C), "sbr_row_gather_fwd: output pitch < C");
  SBR_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "sbr_row_gather_fwd: dropout p must be in [0, 1)");
  const int64_t N = n_idx * k;
  DISPATCH_GROUP(C, {
    const int64_t threads = N * LPRv;
    SBR_CHECK_CUDA(sbr_launch(row_gather_fwd_g_kernel<LPRv, NVg>, dim3(cdiv(threads, 256)), dim3(256), 0, S(stream),
                              srcs_dev, n_mods, idx, mods, N, k, C, normalize, p_drop, seed, step_dev, keep_mask,
                              reinterpret_cast<bf16*>(out_bf16), ld_out, out_f32, ld_f32, err_flag,
                              p_drop > 0.f ? keep_bits_out : (uint8_t*)nullptr));
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
  const int n_state = (int)((n_keys + SCAN_CHUNK - 1) / SCAN_CHUNK);
  int32_t* scan_state = offsets + n_keys + 1;
  SBR_CHECK_CUDA(sbr_launch(plan_count_kernel, dim3(cdiv(N > n_state ? N : n_state, 256)), dim3(256), 0, S(stream),
                            srcs_dev, n_mods, idx, mods, N, k, counts, row_keys, scan_state, n_state));
  SBR_CHECK_CUDA(sbr_launch(plan_scan_kernel, dim3((unsigned)n_state), dim3(1024), 0, S(stream), counts, n_keys, offsets,
                            cursor, (volatile int32_t*)scan_state));
  SBR_CHECK_CUDA(sbr_launch(plan_fill_kernel, dim3(cdiv(N, 256)), dim3(256), 0, S(stream), (const int32_t*)row_keys, N,
                            cursor, perm, sorted_keys));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_row_gather_bwd_segmented(const sbr_modality_src_t* srcs_dev, int n_mods, int64_t n_keys,
                                            const int32_t* offsets, const int32_t* perm, const int32_t* sorted_keys,
                                            int64_t n_rows, int C, int normalize, float p_drop, uint64_t seed,
                                            const int64_t* step_dev, const uint8_t* keep_mask, const float* dx,
                                            int64_t ld_dx, int rows_per_warp, const uint8_t* keep_bits, void* stream) {
  SBR_REQUIRE(srcs_dev && offsets && perm && sorted_keys && dx && n_keys > 0 && n_rows > 0,
              "sbr_row_gather_bwd_segmented: bad arguments");
  SBR_REQUIRE(C > 0 && C <= 1024 && ld_dx >= C, "sbr_row_gather_bwd_segmented: C=%d not in [1, 1024] or ld_dx < C", C);
  SBR_REQUIRE(rows_per_warp >= 1, "sbr_row_gather_bwd_segmented: bad chunking");
  SBR_REQUIRE(n_mods <= SEG_MAX_MODS, "sbr_row_gather_bwd_segmented: at most %d modalities", SEG_MAX_MODS);
  DISPATCH_GROUP(C, {
    const int64_t threads = (int64_t)cdiv(n_rows, 8) * LPRv;  // >= 8 sorted rows per lane group
    int64_t blocks = cdiv(threads, 256);
    int bps = 3, rpg = 0, minb = 3;  // rpg 0: one chunk of equal length per group (27.6 vs 29.7 us with 8-row chunks)
    if (const char* e = getenv("SBR_SEG_BPS")) bps = atoi(e);
    if (const char* e = getenv("SBR_SEG_MINB")) minb = atoi(e);
    const int64_t cap = (int64_t)sbr_num_sms() * bps;  // persistent: every block resident
    if (blocks > cap) blocks = cap;
    if (const char* e = getenv("SBR_SEG_RPG")) rpg = atoi(e);
    if (rpg <= 0) rpg = (int)std::max<int64_t>(8, cdiv(n_rows, blocks * (256 / LPRv)));  // one chunk per group
    if (minb == 4) {
      SBR_CHECK_CUDA(sbr_launch(seg_reduce_g_kernel<LPRv, NVg, 4>, dim3((unsigned)blocks), dim3(256), 0, S(stream),
                                srcs_dev, n_mods, n_keys, offsets, perm, sorted_keys, C, normalize, p_drop, seed,
                                step_dev, keep_mask, dx, ld_dx,
                                p_drop > 0.f ? keep_bits : (const uint8_t*)nullptr, rpg));
    } else {
      SBR_CHECK_CUDA(sbr_launch(seg_reduce_g_kernel<LPRv, NVg, 3>, dim3((unsigned)blocks), dim3(256), 0, S(stream),
                                srcs_dev, n_mods, n_keys, offsets, perm, sorted_keys, C, normalize, p_drop, seed,
                                step_dev, keep_mask, dx, ld_dx,
                                p_drop > 0.f ? keep_bits : (const uint8_t*)nullptr, rpg));
    }
  });
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_tag_bag_fwd(const int32_t* codes, int max_tags, int32_t pad_id, const float* weight, int64_t n_rows,
                               int C, float* out, void* stream) {
  SBR_REQUIRE(codes && weight && out && n_rows > 0 && max_tags >= 1, "sbr_tag_bag_fwd: bad arguments");
  SBR_REQUIRE(C > 0 && C <= 1024, "sbr_tag_bag_fwd: C=%d not in [1, 1024]", C);
  DISPATCH_GROUP(C, {
    const int64_t threads = n_rows * LPRv;
    SBR_CHECK_CUDA(sbr_launch(tag_bag_fwd_kernel<LPRv, NVg>, dim3(cdiv(threads, 256)), dim3(256), 0, S(stream), codes,
                              max_tags, pad_id, weight, n_rows, C, out));
  });
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_tag_bag_bwd(const int32_t* codes, int max_tags, int32_t pad_id, float* bag_grad, int64_t n_rows,
                               int C, float* grad_weight, int64_t n_weight_rows, void* stream) {
  SBR_REQUIRE(codes && bag_grad && grad_weight && n_rows > 0 && max_tags >= 1 && n_weight_rows > 0,
              "sbr_tag_bag_bwd: bad arguments");
  SBR_REQUIRE(C > 0 && C <= 1024, "sbr_tag_bag_bwd: C=%d not in [1, 1024]", C);
  DISPATCH_GROUP(C, {
    int64_t blocks = cdiv(n_rows * LPRv, 256);
    if (blocks > sbr_num_sms()) blocks = sbr_num_sms();
    SBR_CHECK_CUDA(sbr_launch(tag_bag_bwd_kernel<LPRv, NVg>, dim3((unsigned)blocks), dim3(256), 0, S(stream), codes,
                              max_tags, pad_id, bag_grad, n_rows, C, grad_weight, n_weight_rows));
  });
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}
