// sibrar_b200 -- symmetric InfoNCE (CLIP-style) modality-alignment loss with hand-written backward.
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"

namespace {
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// one-warp-per-row kernels keep NV values per lane in registers: supported widths D <= 64 / 128 / 512
#define DISPATCH_NV(n_elems, per, ...)                                   \
  do {                                                                   \
    int _nv = (int)(((n_elems) + (per) - 1) / (per));                    \
    if (_nv <= 2) { constexpr int NVv = 2; __VA_ARGS__; }                \
    else if (_nv <= 4) { constexpr int NVv = 4; __VA_ARGS__; }           \
    else { constexpr int NVv = 16; __VA_ARGS__; }                        \
  } while (0)

// ------------------------------------------------------------------------------------------------ InfoNCE
// e: [G, n, 2, D].  L[i, j] = <e[g,i,0], e[g,j,1]> / T.
// pass 1: lse[0][g][i] = logsumexp_j L[i, j]  (rows),  lse[1][g][j] = logsumexp_i L[i, j]  (columns).
template <int NV>
__global__ void infonce_lse_kernel(const float* __restrict__ e, int64_t G, int64_t n, int D, float inv_t,
                                   float* __restrict__ lse) {
  int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= 2 * G * n) return;
  const int lane = threadIdx.x & 31;
  const int side = (int)(w / (G * n));
  const int64_t gi = w - (int64_t)side * G * n;
  const int64_t g = gi / n, i = gi - g * n;
  const float* base = e + g * n * 2 * D;
  float a[NV];
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    int d = lane + 32 * t;
    a[t] = d < D ? base[(i * 2 + side) * D + d] : 0.f;
  }
  float mx = -INFINITY, se = 0.f;
  for (int64_t j = 0; j < n; ++j) {
    const float* o = base + (j * 2 + (1 - side)) * D;
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      int d = lane + 32 * t;
      if (d < D) dot += a[t] * o[d];
    }
    dot = warp_sum(dot) * inv_t;
    float nm = fmaxf(mx, dot);
    se = se * __expf(mx - nm) + __expf(dot - nm);
    mx = nm;
  }
  if (lane == 0) lse[w] = mx + logf(se);
}

// pass 2: gradient rows.  For side 0, row i:  de0_i = sum_j w_ij e1_j / T;  side 1, row j: de1_j = sum_i w_ij e0_i / T
// with w_ij = (exp(L_ij - lse_r[i]) + exp(L_ij - lse_c[j]) - 2 delta_ij) / R,  R = G * n.
// loss = sum_i (lse_r[i] - L_ii) / R + sum_j (lse_c[j] - L_jj) / R   (added by the side-0 warps).
template <int NV>
__global__ void infonce_grad_kernel(const float* __restrict__ e, int64_t G, int64_t n, int D, float inv_t,
                                    float weight, const float* __restrict__ lse, double* __restrict__ loss_acc,
                                    float* __restrict__ de, int accumulate) {
  int64_t w = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= 2 * G * n) return;
  const int lane = threadIdx.x & 31;
  const int side = (int)(w / (G * n));
  const int64_t gi = w - (int64_t)side * G * n;
  const int64_t g = gi / n, i = gi - g * n;
  const float* base = e + g * n * 2 * D;
  const float* lse_mine = lse + (int64_t)side * G * n + g * n;        // lse over my index
  const float* lse_other = lse + (int64_t)(1 - side) * G * n + g * n;  // lse over the other index
  const float inv_r = 1.f / (float)(G * n);
  float a[NV], acc[NV];
#pragma unroll
  for (int t = 0; t < NV; ++t) {
    int d = lane + 32 * t;
    a[t] = d < D ? base[(i * 2 + side) * D + d] : 0.f;
    acc[t] = 0.f;
  }
  const float my_lse = lse_mine[i];
  float l_ii = 0.f;
  for (int64_t j = 0; j < n; ++j) {
    const float* o = base + (j * 2 + (1 - side)) * D;
    float ov[NV];
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      int d = lane + 32 * t;
      ov[t] = d < D ? o[d] : 0.f;
      dot += a[t] * ov[t];
    }
    dot = warp_sum(dot) * inv_t;
    if (j == i) l_ii = dot;
    float wgt = (__expf(dot - my_lse) + __expf(dot - lse_other[j]) - (j == i ? 2.f : 0.f)) * inv_r * inv_t * weight;
#pragma unroll
    for (int t = 0; t < NV; ++t) acc[t] += wgt * ov[t];
  }
  if (de) {
    float* dst = de + (g * n * 2 + i * 2 + side) * D;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
      int d = lane + 32 * t;
      if (d < D) dst[d] = accumulate ? dst[d] + acc[t] : acc[t];
    }
  }
  if (lane == 0 && loss_acc) atomicAdd(loss_acc, (double)((my_lse - l_ii) * inv_r * weight));
}


// ------------------------------------------------------------------------------------------------ small groups
// Item side of the train step: G = B groups of n = 1 + n_neg rows (n <= 32).  ONE warp owns one group end to end: both
// slots of the group are staged in shared memory once (coalesced 16-byte loads), the n x n logits are computed with
// lane = (i, j) pair, row / column log-sum-exps with lane = row, the n x n gradient weights go back to shared memory
// and both gradient blocks are produced with lane = 4 contiguous columns -- e is read once and de written once per
// step (the two-pass kernels above read every row 2 (n + 1) times and issue one double-precision atomic per row).
// Persistent blocks; the loss is reduced per block before its single atomic.
constexpr int IG_MAX_N = 32;
template <int DV>  // float4 chunks of a row per lane: D <= 128 * DV
__global__ void __launch_bounds__(256)
infonce_group_kernel(const float* __restrict__ e, int64_t G, int n, int D, float inv_t, float weight,
                     double* __restrict__ loss_acc, float* __restrict__ de, int accumulate, int warps_per_block) {
  SBR_PDL_ENTRY();
  extern __shared__ float4 ig_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldr = D + 4;                                   // padded row pitch (floats): rows land in different banks
  const int per_warp = 2 * n * ldr + n * n + 2 * n + 2;    // e0 | e1 | L / W | lse_r | lse_c
  float* base_s = reinterpret_cast<float*>(ig_smem) + (size_t)warp * ((per_warp + 3) & ~3);
  float* s_e0 = base_s;  // slot `side` of the group: s_e0 + side * n * ldr
  float* s_l = base_s + 2 * n * ldr;
  float* s_lse_r = s_l + n * n;
  float* s_lse_c = s_lse_r + n;
  const int d4n = D >> 2;
  const float inv_r = 1.f / (float)(G * n);
  double loss_local = 0.0;
  if (warp < warps_per_block) {
    for (int64_t g = (int64_t)blockIdx.x * warps_per_block + warp; g < G; g += (int64_t)gridDim.x * warps_per_block) {
      const float* src = e + g * n * 2 * D;
      // ---- stage: global row (i, side) -> s_e[side][i]
      for (int c = lane; c < 2 * n * d4n; c += 32) {
        const int row = c / d4n, d4 = c - row * d4n;  // row = i * 2 + side
        const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (size_t)row * D) + d4);
        *reinterpret_cast<float4*>(s_e0 + ((row & 1) * n + (row >> 1)) * ldr + 4 * d4) = v;
      }
      __syncwarp();
      // ---- logits: lane = pair (i, j)
      for (int pq = lane; pq < n * n; pq += 32) {
        const int i = pq / n, j = pq - i * n;
        const float4* a = reinterpret_cast<const float4*>(s_e0 + i * ldr);
        const float4* b = reinterpret_cast<const float4*>(s_e0 + (n + j) * ldr);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        for (int t = 0; t < d4n; ++t) {
          const float4 x = a[t], y = b[t];
          d0 += x.x * y.x; d1 += x.y * y.y; d2 += x.z * y.z; d3 += x.w * y.w;
        }
        s_l[pq] = ((d0 + d1) + (d2 + d3)) * inv_t;
      }
      __syncwarp();
      // ---- log-sum-exp of every row (lanes 0 .. n-1) and every column (the same lanes, second pass)
      if (lane < n) {
        float mx = -INFINITY, mc = -INFINITY;
        for (int j = 0; j < n; ++j) {
          mx = fmaxf(mx, s_l[lane * n + j]);
          mc = fmaxf(mc, s_l[j * n + lane]);
        }
        float se = 0.f, sc = 0.f;
        for (int j = 0; j < n; ++j) {
          se += __expf(s_l[lane * n + j] - mx);
          sc += __expf(s_l[j * n + lane] - mc);
        }
        const float lr = mx + logf(se), lc = mc + logf(sc);
        s_lse_r[lane] = lr;
        s_lse_c[lane] = lc;
        const float lii = s_l[lane * n + lane];
        loss_local += (double)(((lr - lii) + (lc - lii)) * inv_r * weight);
      }
      __syncwarp();
      // ---- gradient weights W_ij (in place)
      for (int pq = lane; pq < n * n; pq += 32) {
        const int i = pq / n, j = pq - i * n;
        const float l = s_l[pq];
        s_l[pq] = (__expf(l - s_lse_r[i]) + __expf(l - s_lse_c[j]) - (i == j ? 2.f : 0.f)) * inv_r * inv_t * weight;
      }
      __syncwarp();
      // ---- de0_i = sum_j W_ij e1_j,  de1_j = sum_i W_ij e0_i : lane = float4 column chunks
      if (de != nullptr) {
        float* dst = de + g * n * 2 * D;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const float* other = s_e0 + (1 - side) * n * ldr;
          for (int i = 0; i < n; ++i) {
            float4 acc[DV];
#pragma unroll
            for (int v = 0; v < DV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = 0; j < n; ++j) {
              const float w = side == 0 ? s_l[i * n + j] : s_l[j * n + i];
#pragma unroll
              for (int v = 0; v < DV; ++v) {
                const int d4 = lane + 32 * v;
                if (d4 < d4n) {
                  const float4 o = *reinterpret_cast<const float4*>(other + j * ldr + 4 * d4);
                  acc[v].x += w * o.x; acc[v].y += w * o.y; acc[v].z += w * o.z; acc[v].w += w * o.w;
                }
              }
            }
#pragma unroll
            for (int v = 0; v < DV; ++v) {
              const int d4 = lane + 32 * v;
              if (d4 < d4n) {
                float4* q = reinterpret_cast<float4*>(dst + (size_t)(i * 2 + side) * D) + d4;
                float4 r = acc[v];
                if (accumulate) {
                  const float4 old = *q;
                  r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
                }
                *q = r;
              }
            }
          }
        }
      }
      __syncwarp();
    }
  }
  // ---- loss: fixed-order sum over the block's lanes, one atomic per block
  __shared__ double s_loss[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss_local += __shfl_xor_sync(0xffffffffu, loss_local, o);
  if (lane == 0) s_loss[warp] = loss_local;
  __syncthreads();
  if (threadIdx.x == 0 && loss_acc != nullptr) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_loss[w];
    atomicAdd(loss_acc, t);
  }
}

// ------------------------------------------------------------------------------------------------ large groups
// User side of the train step: ONE group of n = B rows (in-batch negatives).  The n x n logits, the two gradient products
// and nothing else are GEMMs (sbr_gemm_bf16, tcgen05); the kernels below are the memory-bound passes between them:
//   split   e (fp32) -> bf16 triples  A3 = [hi0 | lo0 | hi0],  B3 = [hi1 | hi1 | lo1]  so that ONE GEMM with K = 3 D gives
//           hi0 hi1 + lo0 hi1 + hi0 lo1 = the fp32 logits to ~2^-16 (single bf16 operands would move a logit of a
//           128-d unit-variance pair by ~0.04 / T)
//   lse     row / column log-sum-exp of L (online max / sum), the loss terms, then
//   weights W_ij = exp(L_ij - lse_r[i]) + exp(L_ij - lse_c[j]) - 2 delta_ij  as bf16, the A operand of  dE0 = W E1  and
//           (MN-major)  dE1 = W^T E0  (scale 1 / (R T) * weight folded into the GEMMs' alpha).
__global__ void infonce_split_kernel(const float* __restrict__ e, int64_t n, int D, int D8, bf16* __restrict__ a3,
                                     bf16* __restrict__ b3) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n * D8) return;
  const int64_t i = t / D8;
  const int d = (int)(t - i * D8);
  const float x0 = d < D ? e[(i * 2 + 0) * D + d] : 0.f, x1 = d < D ? e[(i * 2 + 1) * D + d] : 0.f;
  const bf16 h0 = __float2bfloat16(x0), h1 = __float2bfloat16(x1);
  const bf16 l0 = __float2bfloat16(x0 - __bfloat162float(h0)), l1 = __float2bfloat16(x1 - __bfloat162float(h1));
  bf16* a = a3 + i * 3 * D8 + d;
  bf16* b = b3 + i * 3 * D8 + d;
  a[0] = h0; a[D8] = l0; a[2 * D8] = h0;
  b[0] = h1; b[D8] = h1; b[2 * D8] = l1;
}

__device__ __forceinline__ void online_add(float& m, float& s, float v) {
  if (v > m) {
    s = s * __expf(m - v) + 1.f;
    m = v;
  } else {
    s += __expf(v - m);
  }
}
__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  s = (m == -INFINITY ? 0.f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mm));
  m = mm;
}

// one warp per row of L [n, n]: lse_r[i]; loss += (lse_r[i] - L_ii) * scale
__global__ void __launch_bounds__(256)
infonce_row_lse_kernel(const float* __restrict__ L, int64_t n, float* __restrict__ lse_r, float scale,
                       double* __restrict__ loss_acc) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = blockIdx.x * (int64_t)(blockDim.x >> 5) + warp;
  double part = 0.0;
  if (i < n) {
    const float4* row = reinterpret_cast<const float4*>(L + i * n);
    float m = -INFINITY, s = 0.f;
    for (int64_t j = lane; j < n / 4; j += 32) {
      const float4 v = __ldcs(row + j);
      online_add(m, s, v.x); online_add(m, s, v.y); online_add(m, s, v.z); online_add(m, s, v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      online_merge(m, s, m2, s2);
    }
    const float lse = m + logf(s);
    if (lane == 0) {
      lse_r[i] = lse;
      part = (double)((lse - L[i * n + i]) * scale);
    }
  }
  __shared__ double s_part[8];
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0 && loss_acc != nullptr) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
    atomicAdd(loss_acc, t);
  }
}

// column pass, step 1: block (x = 32-column group, y = row chunk) -> partial (max, sum) of its columns over its rows
__global__ void __launch_bounds__(256)
infonce_col_partial_kernel(const float* __restrict__ L, int64_t n, int64_t rows_per_chunk, float2* __restrict__ partial) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(n, r0 + rows_per_chunk);
  float m = -INFINITY, s = 0.f;
  if (c < n)
    for (int64_t r = r0 + warp; r < r1; r += 8) online_add(m, s, __ldcs(L + r * n + c));
  __shared__ float sm[8][32], ss[8][32];
  sm[warp][lane] = m;
  ss[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < n) {
#pragma unroll
    for (int w = 1; w < 8; ++w) online_merge(m, s, sm[w][lane], ss[w][lane]);
    partial[(int64_t)blockIdx.y * n + c] = make_float2(m, s);
  }
}
// step 2: one thread per column merges the chunk partials -> lse_c[j]; loss += (lse_c[j] - L_jj) * scale
__global__ void __launch_bounds__(256)
infonce_col_finish_kernel(const float2* __restrict__ partial, int n_chunks, const float* __restrict__ L, int64_t n,
                          float* __restrict__ lse_c, float scale, double* __restrict__ loss_acc) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double part = 0.0;
  if (j < n) {
    float m = -INFINITY, s = 0.f;
    for (int q = 0; q < n_chunks; ++q) {
      const float2 p = partial[(int64_t)q * n + j];
      online_merge(m, s, p.x, p.y);
    }
    const float lse = m + logf(s);
    lse_c[j] = lse;
    part = (double)((lse - L[j * n + j]) * scale);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  __shared__ double s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0 && loss_acc != nullptr) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
    atomicAdd(loss_acc, t);
  }
}

// W = exp(L - lse_r[i]) + exp(L - lse_c[j]) - 2 delta_ij  (bf16), four columns per thread
__global__ void __launch_bounds__(256)
infonce_weights_kernel(const float* __restrict__ L, int64_t n, const float* __restrict__ lse_r,
                       const float* __restrict__ lse_c, bf16* __restrict__ W) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // float4 index
  const int64_t per_row = n / 4;
  if (q >= n * per_row) return;
  const int64_t i = q / per_row, j = (q - i * per_row) * 4;
  const float4 l = __ldcs(reinterpret_cast<const float4*>(L) + q);
  const float4 lc = *reinterpret_cast<const float4*>(lse_c + j);
  const float lr = lse_r[i];
  float w[4] = {__expf(l.x - lr) + __expf(l.x - lc.x), __expf(l.y - lr) + __expf(l.y - lc.y),
                __expf(l.z - lr) + __expf(l.z - lc.z), __expf(l.w - lr) + __expf(l.w - lc.w)};
  if (i >= j && i < j + 4) w[i - j] -= 2.f;
  uint2 o;
  *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(w[0], w[1]);
  *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(w[2], w[3]);
  *reinterpret_cast<uint2*>(W + i * n + j) = o;
}

}  // namespace

extern "C" int sbr_infonce_split(const float* e, int64_t n, int D, void* a3, void* b3, void* stream) {
  SBR_REQUIRE(e && a3 && b3 && n > 0 && D > 0, "sbr_infonce_split: bad arguments");
  const int D8 = (D + 7) & ~7;
  infonce_split_kernel<<<cdiv(n * D8, 256), 256, 0, S(stream)>>>(e, n, D, D8, reinterpret_cast<bf16*>(a3),
                                                                  reinterpret_cast<bf16*>(b3));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_infonce_lse(const float* L, int64_t n, float scale, float* lse_r, float* lse_c, float* partial_ws,
                               int n_chunks, double* loss_acc, void* stream) {
  SBR_REQUIRE(L && lse_r && lse_c && partial_ws && n > 0 && (n & 3) == 0 && n_chunks >= 1 && n_chunks <= 65535,
              "sbr_infonce_lse: bad arguments (n must be a multiple of 4)");
  SBR_REQUIRE((reinterpret_cast<uintptr_t>(L) & 15) == 0 && (reinterpret_cast<uintptr_t>(partial_ws) & 7) == 0,
              "sbr_infonce_lse: unaligned buffers");
  infonce_row_lse_kernel<<<cdiv(n, 8), 256, 0, S(stream)>>>(L, n, lse_r, scale, loss_acc);
  const int64_t rpc = (n + n_chunks - 1) / n_chunks;
  infonce_col_partial_kernel<<<dim3(cdiv(n, 32), (unsigned)n_chunks), 256, 0, S(stream)>>>(
      L, n, rpc, reinterpret_cast<float2*>(partial_ws));
  infonce_col_finish_kernel<<<cdiv(n, 256), 256, 0, S(stream)>>>(reinterpret_cast<const float2*>(partial_ws), n_chunks, L,
                                                                 n, lse_c, scale, loss_acc);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_infonce_weights(const float* L, int64_t n, const float* lse_r, const float* lse_c, void* W,
                                   void* stream) {
  SBR_REQUIRE(L && lse_r && lse_c && W && n > 0 && (n & 3) == 0, "sbr_infonce_weights: bad arguments");
  SBR_REQUIRE((reinterpret_cast<uintptr_t>(L) & 15) == 0 && (reinterpret_cast<uintptr_t>(lse_c) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(W) & 7) == 0,
              "sbr_infonce_weights: unaligned buffers");
  infonce_weights_kernel<<<cdiv(n * (n / 4), 256), 256, 0, S(stream)>>>(L, n, lse_r, lse_c, reinterpret_cast<bf16*>(W));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_infonce(const float* e, int64_t G, int64_t n, int D, float temperature, float weight,
                           double* loss_acc, float* de, int accumulate, float* lse_ws, void* stream) {
  SBR_REQUIRE(e && lse_ws && G > 0 && n > 0, "sbr_infonce: bad arguments");
  SBR_REQUIRE(D > 0 && D <= 512, "sbr_infonce: D=%d not in [1, 512]", D);
  SBR_REQUIRE(temperature > 0.f, "sbr_infonce: temperature must be positive");
  // small groups (the item side: n = 1 + n_neg): one warp per group, everything staged in shared memory once
  if (n <= IG_MAX_N && (D & 3) == 0 && (reinterpret_cast<uintptr_t>(e) & 15) == 0 &&
      (de == nullptr || (reinterpret_cast<uintptr_t>(de) & 15) == 0) && getenv("SBR_INFONCE_TWO_PASS") == nullptr) {
    const int per_warp = ((2 * (int)n * (D + 4) + (int)(n * n) + 2 * (int)n + 2) + 3) & ~3;
    const size_t per_warp_bytes = (size_t)per_warp * sizeof(float);
    int wpb = (int)std::min<size_t>(8, (100 * 1024) / per_warp_bytes);  // two blocks per SM when the groups are small
    if (wpb >= 1) {
      const size_t smem = per_warp_bytes * wpb;
      int64_t blocks = (G + wpb - 1) / wpb;
      const int64_t cap = (int64_t)sbr_num_sms() * 2;
      if (blocks > cap) blocks = cap;
      auto launch = [&](auto kern) -> int {
        SBR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(100 * 1024)));
        SBR_CHECK_CUDA(sbr_launch(kern, dim3((unsigned)blocks), dim3(256), smem, S(stream), e, G, (int)n, D,
                                  1.f / temperature, weight, loss_acc, de, accumulate, wpb));
        return SBR_OK;
      };
      int rc = D <= 128 ? launch(infonce_group_kernel<1>) : (D <= 256 ? launch(infonce_group_kernel<2>)
                                                                      : launch(infonce_group_kernel<4>));
      if (rc) return rc;
      SBR_LAUNCH_CHECK();
      return SBR_OK;
    }
  }
  const int64_t warps = 2 * G * n;
  DISPATCH_NV(D, 32, {
    infonce_lse_kernel<NVv><<<cdiv(warps, 8), 256, 0, S(stream)>>>(e, G, n, D, 1.f / temperature, lse_ws);
    infonce_grad_kernel<NVv><<<cdiv(warps, 8), 256, 0, S(stream)>>>(e, G, n, D, 1.f / temperature, weight, lse_ws,
                                                                    loss_acc, de, accumulate);
  });
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

