// sibrar_b200 -- the "referenced rows" route of the modality projections.
//
// The reference projects the feature rows of exactly the batch's entities (algorithms/sgd_alg.py:1949-1974:
// FeatureEmbedding.forward(indices_with_modality) per sampled modality, data/Feature.py:140-162 fetches those rows).  The
// table route of this library projects ALL rows of a modality once per step, which is the cheaper formulation when a step
// touches most of the catalogue, and wasteful when it does not (paper batch on a 10^5-item catalogue, item-sharded
// evaluation).  This file provides the per-step row subset:
//   * sbr_mark_referenced: one pass over the (entity, modality) slots of the step; the first slot that touches a feature
//     row appends it to the modality's row list and assigns it a compact position (epoch-stamped: nothing is cleared
//     between steps).  The projected "table" of the step then has one row per REFERENCED feature row and is presented to
//     the gather kernels as an indirect source (codes = row -> compact position).
//   * sbr_gather_rows_bf16: X_ref = X[list] (dense bf16 feature rows), the A operand of the projection GEMMs.
//   * sbr_spmm_scatter_wgrad: wgrad of a sparse-input Linear over the referenced rows only -- every stored entry (r, j)
//     adds dz[pos(r), :] into the TRANSPOSED gradient row j with vector reductions.
//   * sbr_transpose_add_f32: g[out, in] += gT[in, out]^T (and clears gT).
#include <stdarg.h>

#include "common.cuh"

namespace {
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

__global__ void __launch_bounds__(256)
mark_referenced_kernel(const sbr_modality_src_t* __restrict__ srcs, int n_mods, const int64_t* __restrict__ idx,
                       const uint8_t* __restrict__ mods, int64_t N, int k, int64_t* epoch_dev,
                       const sbr_ref_table_t* __restrict__ tabs) {
  SBR_PDL_ENTRY();
  // epoch of this call = stored epoch + 1; the LAST block to finish stores it (every block has read it by then)
  const int32_t epoch = (int32_t)((epoch_dev[0] + 1) & 0x7fffffff);
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < N) {
    const int m = min(mods ? (int)mods[r] : 0, n_mods - 1);
    const sbr_ref_table_t t = tabs[m];
    if (t.stamp != nullptr) {  // (nullptr: this modality keeps the whole-table route)
      const int64_t e = idx[r / k];
      const int32_t* remap = srcs[m].remap;
      const int64_t feat = remap ? (int64_t)__ldg(remap + e) : e;
      if (feat >= 0 && atomicExch(t.stamp + feat, epoch) != epoch) {  // first slot of this call that touches the row
        const int32_t p = atomicAdd(t.count, 1);
        t.list[p] = (int32_t)feat;
        t.pos[feat] = p;
        if (t.seg_first != nullptr) {  // sparse rows: the units of work are the row's segments
          const int32_t s0 = (int32_t)t.seg_first[feat], s1 = (int32_t)t.seg_first[feat + 1];
          const int32_t q = atomicAdd(t.seg_count, s1 - s0);
          for (int32_t s = s0; s < s1; ++s) t.seg_list[q + (s - s0)] = s;
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long done = atomicAdd(reinterpret_cast<unsigned long long*>(epoch_dev + 1), 1ULL);
    if (done == (unsigned long long)gridDim.x - 1) {
      epoch_dev[1] = 0;
      epoch_dev[0] += 1;
    }
  }
}

// dst[slot, :] = src[list[slot], :] for slot < *count, zero rows beyond (16-byte chunks; ld in elements, multiples of 8)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4* __restrict__ src, int64_t ld_src16, const int32_t* __restrict__ list,
                   const int32_t* __restrict__ count, int64_t capacity, int64_t chunks, uint4* __restrict__ dst,
                   int64_t ld_dst16) {
  SBR_PDL_ENTRY();
  const int64_t n = (int64_t)*count;
  const int64_t total = capacity * chunks;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t slot = i / chunks, c = i - slot * chunks;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (slot < n) v = __ldg(src + (int64_t)__ldg(list + slot) * ld_src16 + c);
    dst[slot * ld_dst16 + c] = v;
  }
}

// one warp per listed unit (row or row segment): gT[j, :] += vals[p] * dz[pos[row], :] for the unit's entries (p, j)
template <int NV8>
__global__ void __launch_bounds__(256)
spmm_scatter_wgrad_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                          const float* __restrict__ vals, const int32_t* __restrict__ unit_list,
                          const int32_t* __restrict__ n_units_dev, const int32_t* __restrict__ seg_row,
                          const int32_t* __restrict__ pos, const bf16* __restrict__ dz, int64_t ld_dz, int C,
                          float* __restrict__ gT, int64_t ld_gT) {
  SBR_PDL_ENTRY();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t slot = (int64_t)blockIdx.x * 8 + warp;
  if (slot >= (int64_t)*n_units_dev) return;
  const int64_t unit = __ldg(unit_list + slot);
  const int64_t row = seg_row ? (int64_t)(__ldg(seg_row + unit) & 0x7fffffff) : unit;
  const bf16* d = dz + (int64_t)__ldg(pos + row) * ld_dz;
  const int C8 = (C + 7) >> 3;
  float g[NV8][8];
#pragma unroll
  for (int i = 0; i < NV8; ++i) {
    const int c8 = lane + 32 * i;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (c8 < C8) u = __ldg(reinterpret_cast<const uint4*>(d) + c8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __bfloat1622float2(h[t]);
      g[i][2 * t] = f.x;
      g[i][2 * t + 1] = f.y;
    }
  }
  const int64_t beg = indptr[unit], end = indptr[unit + 1];
  const bool vec = (C & 7) == 0 && (ld_gT & 3) == 0;
  for (int64_t p = beg; p < end; p += 32) {
    const int32_t my = (p + lane < end) ? __ldg(indices + p + lane) : 0;
    const float myv = (vals != nullptr && p + lane < end) ? __ldg(vals + p + lane) : 1.f;
    const int cnt = (int)min((int64_t)32, end - p);
    for (int t = 0; t < cnt; ++t) {
      const int32_t j = __shfl_sync(0xffffffffu, my, t);
      const float w = __shfl_sync(0xffffffffu, myv, t);
      float* dst = gT + (int64_t)j * ld_gT;
#pragma unroll
      for (int i = 0; i < NV8; ++i) {
        const int c8 = lane + 32 * i;
        if (c8 >= C8) continue;
        if (vec) {
          const size_t a = __cvta_generic_to_global(dst + 8 * c8);
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(w * g[i][0]), "f"(w * g[i][1]),
                       "f"(w * g[i][2]), "f"(w * g[i][3])
                       : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a + 16), "f"(w * g[i][4]),
                       "f"(w * g[i][5]), "f"(w * g[i][6]), "f"(w * g[i][7])
                       : "memory");
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (8 * c8 + q < C) atomicAdd(dst + 8 * c8 + q, w * g[i][q]);
        }
      }
    }
  }
}

// dst[c, r] += src[r, c]; src cleared.  64 x 64 tiles through shared memory, 16-byte accesses on both sides (the
// transposed gradient of a sparse-input Linear is [n_entities, C]: tens of millions of elements every step).
__global__ void __launch_bounds__(256)
transpose_add_kernel(float* __restrict__ src, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst, int64_t rows,
                     int64_t cols, int vec) {
  SBR_PDL_ENTRY();
  __shared__ float tile[64][65];
  const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16: a thread owns 4 consecutive elements of 4 lines
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t r = r0 + ty + 16 * j, c = c0 + 4 * tx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows) {
      float* p = src + r * ld_src + c;
      if (vec && c + 4 <= cols) {
        v = *reinterpret_cast<float4*>(p);
        *reinterpret_cast<float4*>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c + q < cols) { e[q] = p[q]; p[q] = 0.f; }
        v = make_float4(e[0], e[1], e[2], e[3]);
      }
    }
    tile[ty + 16 * j][4 * tx + 0] = v.x;
    tile[ty + 16 * j][4 * tx + 1] = v.y;
    tile[ty + 16 * j][4 * tx + 2] = v.z;
    tile[ty + 16 * j][4 * tx + 3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t c = c0 + ty + 16 * j, r = r0 + 4 * tx;  // output line c, 4 consecutive r
    if (c >= cols) continue;
    float* p = dst + c * ld_dst + r;
    const float e[4] = {tile[4 * tx + 0][ty + 16 * j], tile[4 * tx + 1][ty + 16 * j], tile[4 * tx + 2][ty + 16 * j],
                        tile[4 * tx + 3][ty + 16 * j]};
    if (vec && r + 4 <= rows) {
      float4 o = *reinterpret_cast<float4*>(p);
      o.x += e[0]; o.y += e[1]; o.z += e[2]; o.w += e[3];
      *reinterpret_cast<float4*>(p) = o;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (r + q < rows) p[q] += e[q];
    }
  }
}

}  // namespace

extern "C" int sbr_mark_referenced(const sbr_modality_src_t* srcs_dev, int n_mods, const int64_t* idx,
                                   const uint8_t* mods, int64_t n_idx, int k, int64_t* epoch_dev,
                                   const sbr_ref_table_t* tabs_dev, void* stream) {
  SBR_REQUIRE(srcs_dev && idx && epoch_dev && tabs_dev && n_idx > 0 && k >= 1 && n_mods >= 1,
              "sbr_mark_referenced: bad arguments");
  const int64_t N = n_idx * k;
  SBR_CHECK_CUDA(sbr_launch(mark_referenced_kernel, dim3(cdiv(N, 256)), dim3(256), (size_t)0, S(stream), srcs_dev, n_mods,
                            idx, mods, N, k, epoch_dev, tabs_dev));
  return SBR_OK;
}

extern "C" int sbr_gather_rows_bf16(const void* src, int64_t ld_src, const int32_t* list, const int32_t* count_dev,
                                    int64_t capacity, int64_t cols, void* dst, int64_t ld_dst, void* stream) {
  SBR_REQUIRE(src && list && count_dev && dst && capacity > 0 && cols > 0, "sbr_gather_rows_bf16: bad arguments");
  SBR_REQUIRE(ld_src % 8 == 0 && ld_dst % 8 == 0 && cols % 8 == 0 && cols <= ld_src && cols <= ld_dst &&
                  (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "sbr_gather_rows_bf16: rows must be 16-byte aligned and padded to 8 elements");
  const int64_t chunks = cols / 8;
  int64_t blocks = (capacity * chunks + 255) / 256;
  const int64_t cap = 32 * (int64_t)sbr_num_sms();
  if (blocks > cap) blocks = cap;
  SBR_CHECK_CUDA(sbr_launch(gather_rows_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, S(stream),
                            reinterpret_cast<const uint4*>(src), ld_src / 8, list, count_dev, capacity, chunks,
                            reinterpret_cast<uint4*>(dst), ld_dst / 8));
  return SBR_OK;
}

extern "C" int sbr_spmm_scatter_wgrad(const int64_t* indptr, const int32_t* indices, const float* vals,
                                      const int32_t* unit_list, const int32_t* n_units_dev, int64_t max_units,
                                      const int32_t* seg_row, const int32_t* pos, const void* dz_bf16, int64_t ld_dz,
                                      int64_t C, float* gT, int64_t ld_gT, void* stream) {
  SBR_REQUIRE(indptr && indices && unit_list && n_units_dev && pos && dz_bf16 && gT && max_units > 0,
              "sbr_spmm_scatter_wgrad: bad arguments");
  SBR_REQUIRE(C > 0 && C <= 1024 && ld_dz % 8 == 0 && (reinterpret_cast<uintptr_t>(dz_bf16) & 15) == 0,
              "sbr_spmm_scatter_wgrad: C=%lld, ld_dz=%lld", (long long)C, (long long)ld_dz);
  const bf16* d = reinterpret_cast<const bf16*>(dz_bf16);
  const int nv8 = (int)(((C + 7) / 8 + 31) / 32);
  const dim3 grid(cdiv(max_units, 8));
#define SBR_SCAT(N)                                                                                                   \
  SBR_CHECK_CUDA(sbr_launch(spmm_scatter_wgrad_kernel<N>, grid, dim3(256), (size_t)0, S(stream), indptr, indices, vals, \
                            unit_list, n_units_dev, seg_row, pos, d, ld_dz, (int)C, gT, ld_gT))
  if (nv8 <= 1) { SBR_SCAT(1); }
  else if (nv8 <= 2) { SBR_SCAT(2); }
  else { SBR_SCAT(4); }
#undef SBR_SCAT
  return SBR_OK;
}

extern "C" int sbr_transpose_add_f32(float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int64_t cols,
                                     void* stream) {
  SBR_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= rows,
              "sbr_transpose_add_f32: bad arguments");
  dim3 grid(cdiv(cols, 64), cdiv(rows, 64));
  const int vec = (ld_src % 4 == 0) && (ld_dst % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  SBR_CHECK_CUDA(sbr_launch(transpose_add_kernel, grid, dim3(256), (size_t)0, S(stream), src, ld_src, dst, ld_dst, rows,
                            cols, vec));
  return SBR_OK;
}
