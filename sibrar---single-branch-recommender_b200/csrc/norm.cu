// sibrar_b200 -- BatchNorm1d (train statistics finalize / apply / backward) and the table-level activation
// gradient with column sums.  Thread block = 8 row-lanes x 32 consecutive columns; grid = row blocks x column chunks,
// so every warp reads 128 contiguous bytes per row and column partial sums stay in one register per thread.
// Replaces torch.nn.BatchNorm1d as used at modules/polylinear.py:58-61,68-69 and algorithms/sgd_alg.py:1834-1837,
// and the autograd of the projection's output activation (algorithms/sgd_alg.py:1356).
#include <stdlib.h>

#include "common.cuh"

namespace {
constexpr int BNV_UNR = 4;  // rows in flight per thread in the vector kernels
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// rows per block: 256 for long tensors, 32 for short ones (a [3 706 x 64] table would otherwise run on 30 blocks
// whose threads walk 32 rows one dependent load at a time)
inline int rows_per_block(int64_t rows) {
  static const char* e = getenv("SBR_NORM_RB");  // measurement only
  if (e) return atoi(e);
  return rows >= 65536 ? 256 : 32;
}

__device__ __forceinline__ float load_y(const float* y_f32, const bf16* y_bf16, int64_t off) {
  return y_f32 ? y_f32[off] : __bfloat162float(y_bf16[off]);
}

// reduce `v` over the 8 row-lanes of the block and atomically add to dst[c]
__device__ __forceinline__ void block_col_flush(float v, float* dst, int c, int C) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  red[ty][tx] = v;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][tx];
    atomicAdd(dst + c, t);
  }
  __syncthreads();
}

__global__ void actgrad_colsum_kernel(float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ y_f32,
                                      const bf16* __restrict__ y_bf16, int64_t ld_y, int act, int64_t rows, int C,
                                      bf16* __restrict__ out_bf16, int64_t ld_out, float* __restrict__ out_f32,
                                      int64_t ld_out_f32, float* __restrict__ colsum, int zero_dy, int RB) {
  SBR_PDL_ENTRY();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.x * RB;
  float part = 0.f;
  if (c < C) {
    for (int rr = ty; rr < RB; rr += 8) {
      int64_t r = r0 + rr;
      if (r >= rows) break;
      float v = dy[r * ld_dy + c];
      if (zero_dy) dy[r * ld_dy + c] = 0.f;
      if (act != SBR_ACT_NONE) v *= act_grad_from_out(act, load_y(y_f32, y_bf16, r * ld_y + c));
      part += v;
      if (out_bf16) out_bf16[r * ld_out + c] = __float2bfloat16(v);
      if (out_f32) out_f32[r * ld_out_f32 + c] = v;
    }
  }
  if (colsum) block_col_flush(part, colsum, c, C);
}

// block = 32 columns x 32 partial-row lanes; partial rows are added in a fixed order (deterministic statistics):
// lane ty sums rows ty, ty + 32, ... with four independent accumulators, then the 32 lane sums are added in order
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ stats, int n_partials, int64_t n_rows, int C, float eps, float momentum,
                   float* __restrict__ mean_invstd, float* running_mean, float* running_var,
                   int64_t* num_batches_tracked) {
  SBR_PDL_ENTRY();
  // block = 8 columns x 128 row lanes (C = 64: 8 blocks, ~9 partial rows per lane), then a fixed pairwise tree in shared
  // memory: the result depends on n_partials only, never on the arrival order of the producers
  __shared__ float red[2][128][9];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + tx;
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
  float a = 0.f, b = 0.f;
  if (c < C) {
    for (int r = ty; r < n_partials; r += 128) {
      a += stats[(size_t)r * 2 * C + c];
      b += stats[(size_t)r * 2 * C + C + c];
    }
  }
  red[0][ty][tx] = a;
  red[1][ty][tx] = b;
  __syncthreads();
#pragma unroll
  for (int s = 64; s > 0; s >>= 1) {
    if (ty < s) {
      red[0][ty][tx] += red[0][ty + s][tx];
      red[1][ty][tx] += red[1][ty + s][tx];
    }
    __syncthreads();
  }
  if (ty != 0 || c >= C) return;
  const double s1 = (double)red[0][0][tx], s2 = (double)red[1][0][tx];
  double n = (double)n_rows;
  double mean = s1 / n;
  double var = s2 / n - mean * mean;
  if (var < 0.) var = 0.;
  mean_invstd[c] = (float)mean;
  mean_invstd[C + c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    double unbiased = n > 1. ? var * n / (n - 1.) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_coeffs_kernel(const float* __restrict__ rm, const float* __restrict__ rv, int C, float eps,
                                      float* __restrict__ mean_invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean_invstd[c] = rm[c];
  mean_invstd[C + c] = rsqrtf(rv[c] + eps);
}

__global__ void bn_apply_kernel(const float* __restrict__ z, int64_t ld_z, const float* __restrict__ mean_invstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                                int64_t rows, int C, bf16* __restrict__ out_bf16, int64_t ld_bf16,
                                float* __restrict__ out_f32, int64_t ld_f32, int RB) {
  SBR_PDL_ENTRY();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + tx;
  if (c >= C) return;
  const int64_t r0 = (int64_t)blockIdx.x * RB;
  const float mean = mean_invstd[c], sc = mean_invstd[C + c] * gamma[c], sh = beta[c];
  for (int rr = ty; rr < RB; rr += 8) {
    int64_t r = r0 + rr;
    if (r >= rows) break;
    float v = act_fwd(act, (z[r * ld_z + c] - mean) * sc + sh);
    if (out_bf16) out_bf16[r * ld_bf16 + c] = __float2bfloat16(v);
    if (out_f32) out_f32[r * ld_f32 + c] = v;
  }
}

// sums[0:C] += dzb, sums[C:2C] += dzb * xhat   with dzb = dy * act'(y)
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ y_f32,
                                     const bf16* __restrict__ y_bf16, int64_t ld_y, int act,
                                     const float* __restrict__ z, int64_t ld_z, const float* __restrict__ mean_invstd,
                                     int64_t rows, int C, float* __restrict__ sums, int RB) {
  SBR_PDL_ENTRY();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.x * RB;
  float p0 = 0.f, p1 = 0.f;
  if (c < C) {
    const float mean = mean_invstd[c], invstd = mean_invstd[C + c];
    for (int rr = ty; rr < RB; rr += 8) {
      int64_t r = r0 + rr;
      if (r >= rows) break;
      float g = dy[r * ld_dy + c];
      if (act != SBR_ACT_NONE) g *= act_grad_from_out(act, load_y(y_f32, y_bf16, r * ld_y + c));
      p0 += g;
      p1 += g * (z[r * ld_z + c] - mean) * invstd;
    }
  }
  block_col_flush(p0, sums, c, C);
  block_col_flush(p1, sums + C, c, C);
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ y_f32,
                                    const bf16* __restrict__ y_bf16, int64_t ld_y, int act,
                                    const float* __restrict__ z, int64_t ld_z, const float* __restrict__ mean_invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ sums, int n_replicas,
                                    int64_t rows, int C, bf16* __restrict__ dz_bf16, int64_t ld_dz, float* __restrict__ dz_f32,
                                    int64_t ld_dz_f32, float* dgamma, float* dbeta, int RB) {
  SBR_PDL_ENTRY();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + tx;
  if (c >= C) return;
  const int64_t r0 = (int64_t)blockIdx.x * RB;
  const float inv_n = 1.f / (float)rows;
  float s0 = 0.f, s1 = 0.f;
  for (int r = 0; r < n_replicas; ++r) {
    s0 += sums[(size_t)r * 2 * C + c];
    s1 += sums[(size_t)r * 2 * C + C + c];
  }
  if (blockIdx.x == 0 && ty == 0) {
    if (dbeta) dbeta[c] += s0;
    if (dgamma) dgamma[c] += s1;
  }
  const float mean = mean_invstd[c], invstd = mean_invstd[C + c], gi = gamma[c] * invstd;
  for (int rr = ty; rr < RB; rr += 8) {
    int64_t r = r0 + rr;
    if (r >= rows) break;
    float g = dy[r * ld_dy + c];
    if (act != SBR_ACT_NONE) g *= act_grad_from_out(act, load_y(y_f32, y_bf16, r * ld_y + c));
    float xh = (z[r * ld_z + c] - mean) * invstd;
    float v = gi * (g - s0 * inv_n - xh * s1 * inv_n);
    if (dz_bf16) dz_bf16[r * ld_dz + c] = __float2bfloat16(v);
    if (dz_f32) dz_f32[r * ld_dz_f32 + c] = v;
  }
}

inline dim3 tile_grid(int64_t rows, int64_t cols) { return dim3(cdiv(rows, rows_per_block(rows)), cdiv(cols, 32)); }
}  // namespace

// Vector path of actgrad_colsum (C % 4 == 0, 256 % (C / 4) == 0, 16-byte aligned rows): 4 columns per thread, the loads
// of 4 rows in flight, one atomic per column and BLOCK (a [3 706 x 64] table: 58 blocks instead of 232 whose
// same-address atomics queue up in one L2 slice).
__global__ void __launch_bounds__(256)
actgrad_colsum_vec_kernel(float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ y_f32,
                          const bf16* __restrict__ y_bf16, int64_t ld_y, int act, int64_t rows, int C,
                          bf16* __restrict__ out_bf16, int64_t ld_out, float* __restrict__ out_f32,
                          int64_t ld_out_f32, float* __restrict__ colsum, int zero_dy) {
  SBR_PDL_ENTRY();
  constexpr int UNR = 4;
  const int c4n = C >> 2;
  const int c = (threadIdx.x % c4n) * 4;
  const int rpb = 256 / c4n;
  const int64_t stride = (int64_t)gridDim.x * rpb;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t r0 = (int64_t)blockIdx.x * rpb + threadIdx.x / c4n; r0 < rows; r0 += stride * UNR) {
    float4 g[UNR], yv[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= rows) continue;
      g[u] = *reinterpret_cast<const float4*>(dy + r * ld_dy + c);
      if (act != SBR_ACT_NONE) {
        if (y_f32) {
          yv[u] = *reinterpret_cast<const float4*>(y_f32 + r * ld_y + c);
        } else {
          const uint2 raw = *reinterpret_cast<const uint2*>(y_bf16 + r * ld_y + c);
          const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
          const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
          yv[u] = make_float4(a.x, a.y, b.x, b.y);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= rows) break;
      float v[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
      if (act != SBR_ACT_NONE) {
        v[0] *= act_grad_from_out(act, yv[u].x);
        v[1] *= act_grad_from_out(act, yv[u].y);
        v[2] *= act_grad_from_out(act, yv[u].z);
        v[3] *= act_grad_from_out(act, yv[u].w);
      }
      if (zero_dy) *reinterpret_cast<float4*>(dy + r * ld_dy + c) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) part[j] += v[j];
      if (out_bf16) {
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(out_bf16 + r * ld_out + c) = o;
      }
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * ld_out_f32 + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  if (colsum == nullptr) return;
  __shared__ float4 red[256];
  red[threadIdx.x] = make_float4(part[0], part[1], part[2], part[3]);
  __syncthreads();
  if (threadIdx.x < c4n) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < rpb; ++k) {
      const float4 a = red[k * c4n + threadIdx.x];
      t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
    }
    atomicAdd(colsum + c, t.x);
    atomicAdd(colsum + c + 1, t.y);
    atomicAdd(colsum + c + 2, t.z);
    atomicAdd(colsum + c + 3, t.w);
  }
}


// ---- vector paths of bn_apply / bn_bwd_reduce and of bn_bwd_apply with an activation behind the BatchNorm (the wide
// single-branch networks: BatchNorm -> ReLU every second layer, 360 448 x 512 rows per step): 4 columns per thread, the
// loads of 4 rows in flight (the scalar kernels -- one 4-byte load per thread and row -- ran at 2.1 - 2.5 TB/s there)
__device__ __forceinline__ float4 load_y4(const float* y_f32, const bf16* y_bf16, int64_t off) {
  if (y_f32) return __ldcs(reinterpret_cast<const float4*>(y_f32 + off));
  const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(y_bf16 + off));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store_bf16x4(bf16* dst, const float (&v)[4]) {
  uint2 o;
  *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v[0], v[1]);
  *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(dst) = o;
}

__global__ void __launch_bounds__(256)
bn_apply_vec_kernel(const float* __restrict__ z, int64_t ld_z, const float* __restrict__ mean_invstd,
                    const float* __restrict__ gamma, const float* __restrict__ beta, int act, int64_t rows, int C,
                    bf16* __restrict__ out_bf16, int64_t ld_bf16, float* __restrict__ out_f32, int64_t ld_f32) {
  SBR_PDL_ENTRY();
  const int c4n = C >> 2;
  const int c = (threadIdx.x % c4n) * 4;
  const int rpb = 256 / c4n;
  const float4 mean = *reinterpret_cast<const float4*>(mean_invstd + c);
  const float4 invstd = *reinterpret_cast<const float4*>(mean_invstd + C + c);
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
  const float4 bt = *reinterpret_cast<const float4*>(beta + c);
  const float m[4] = {mean.x, mean.y, mean.z, mean.w};
  const float sc[4] = {gm.x * invstd.x, gm.y * invstd.y, gm.z * invstd.z, gm.w * invstd.w};
  const float sh[4] = {bt.x, bt.y, bt.z, bt.w};
  const int64_t stride = (int64_t)gridDim.x * rpb;
  for (int64_t r0 = (int64_t)blockIdx.x * rpb + threadIdx.x / c4n; r0 < rows; r0 += stride * BNV_UNR) {
    float4 zz[BNV_UNR];
#pragma unroll
    for (int u = 0; u < BNV_UNR; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < rows) zz[u] = __ldcs(reinterpret_cast<const float4*>(z + r * ld_z + c));
    }
#pragma unroll
    for (int u = 0; u < BNV_UNR; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= rows) break;
      const float zv[4] = {zz[u].x, zz[u].y, zz[u].z, zz[u].w};
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = act_fwd(act, (zv[j] - m[j]) * sc[j] + sh[j]);
      if (out_bf16) store_bf16x4(out_bf16 + r * ld_bf16 + c, v);
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * ld_f32 + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_vec_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ y_f32,
                         const bf16* __restrict__ y_bf16, int64_t ld_y, int act, const float* __restrict__ z,
                         int64_t ld_z, const float* __restrict__ mean_invstd, int64_t rows, int C,
                         float* __restrict__ sums) {
  SBR_PDL_ENTRY();
  const int c4n = C >> 2;
  const int c = (threadIdx.x % c4n) * 4;
  const int rpb = 256 / c4n;
  const float4 mean = *reinterpret_cast<const float4*>(mean_invstd + c);
  const float4 invstd = *reinterpret_cast<const float4*>(mean_invstd + C + c);
  const float m[4] = {mean.x, mean.y, mean.z, mean.w}, is[4] = {invstd.x, invstd.y, invstd.z, invstd.w};
  float p0[4] = {0.f, 0.f, 0.f, 0.f}, p1[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t stride = (int64_t)gridDim.x * rpb;
  for (int64_t r0 = (int64_t)blockIdx.x * rpb + threadIdx.x / c4n; r0 < rows; r0 += stride * BNV_UNR) {
    float4 g[BNV_UNR], zz[BNV_UNR], yv[BNV_UNR];
#pragma unroll
    for (int u = 0; u < BNV_UNR; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < rows) {
        g[u] = __ldg(reinterpret_cast<const float4*>(dy + r * ld_dy + c));  // (read again by bn_bwd_apply)
        zz[u] = __ldg(reinterpret_cast<const float4*>(z + r * ld_z + c));
        if (act != SBR_ACT_NONE) yv[u] = load_y4(y_f32, y_bf16, r * ld_y + c);
      }
    }
#pragma unroll
    for (int u = 0; u < BNV_UNR; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= rows) break;
      float gv[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
      const float zv[4] = {zz[u].x, zz[u].y, zz[u].z, zz[u].w};
      if (act != SBR_ACT_NONE) {
        gv[0] *= act_grad_from_out(act, yv[u].x);
        gv[1] *= act_grad_from_out(act, yv[u].y);
        gv[2] *= act_grad_from_out(act, yv[u].z);
        gv[3] *= act_grad_from_out(act, yv[u].w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p0[j] += gv[j];
        p1[j] += gv[j] * (zv[j] - m[j]) * is[j];
      }
    }
  }
  __shared__ float4 red[2][256];
  red[0][threadIdx.x] = make_float4(p0[0], p0[1], p0[2], p0[3]);
  red[1][threadIdx.x] = make_float4(p1[0], p1[1], p1[2], p1[3]);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * c4n; i += 256) {
    const int which = i / c4n, t = i % c4n;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < rpb; ++k) {
      const float4 b = red[which][k * c4n + t];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float* dst = sums + (size_t)which * C + 4 * t;
    atomicAdd(dst, a.x);
    atomicAdd(dst + 1, a.y);
    atomicAdd(dst + 2, a.z);
    atomicAdd(dst + 3, a.w);
  }
}

extern "C" int sbr_actgrad_colsum(float* dy, int64_t ld_dy, const float* y_f32, const void* y_bf16, int64_t ld_y,
                                  int act, int64_t rows, int64_t cols, void* out_bf16, int64_t ld_out, float* out_f32,
                                  int64_t ld_out_f32, float* colsum, int zero_dy, void* stream) {
  SBR_REQUIRE(dy && rows > 0 && cols > 0, "sbr_actgrad_colsum: bad arguments");
  SBR_REQUIRE(act == SBR_ACT_NONE || y_f32 || y_bf16, "sbr_actgrad_colsum: activation gradient needs the output y");
  {
    const auto al = [](const void* p, int a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
    const int C = (int)cols;
    const bool vec = (C & 3) == 0 && C <= 1024 && 256 % (C >> 2) == 0 && (ld_dy & 3) == 0 && al(dy, 16) &&
                     (act == SBR_ACT_NONE || (y_f32 ? ((ld_y & 3) == 0 && al(y_f32, 16)) : ((ld_y & 3) == 0 && al(y_bf16, 8)))) &&
                     (!out_bf16 || ((ld_out & 3) == 0 && al(out_bf16, 8))) &&
                     (!out_f32 || ((ld_out_f32 & 3) == 0 && al(out_f32, 16))) && getenv("SBR_NORM_SCALAR") == nullptr;
    if (vec) {
      const int rpb = 256 / (C >> 2);
      int64_t blocks = cdiv(rows, (int64_t)rpb * 4);
      const int64_t cap = (int64_t)sbr_num_sms() * 8;
      if (blocks > cap) blocks = cap;
      SBR_CHECK_CUDA(sbr_launch(actgrad_colsum_vec_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), S(stream), dy,
                                ld_dy, y_f32, reinterpret_cast<const bf16*>(y_bf16), ld_y, act, rows, C,
                                reinterpret_cast<bf16*>(out_bf16), ld_out, out_f32, ld_out_f32, colsum, zero_dy));
      SBR_LAUNCH_CHECK();
      return SBR_OK;
    }
  }
  SBR_CHECK_CUDA(sbr_launch(actgrad_colsum_kernel, dim3(tile_grid(rows, cols)), dim3(256), (size_t)(0), S(stream), 
      dy, ld_dy, y_f32, reinterpret_cast<const bf16*>(y_bf16), ld_y, act, rows, (int)cols,
      reinterpret_cast<bf16*>(out_bf16), ld_out, out_f32, ld_out_f32, colsum, zero_dy, rows_per_block(rows)));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_bn_finalize(const float* stats, int n_partials, int64_t n_rows, int C, float eps, float momentum,
                               float* mean_invstd, float* running_mean, float* running_var,
                               int64_t* num_batches_tracked, void* stream) {
  SBR_REQUIRE(stats && mean_invstd && n_rows > 0 && C > 0 && n_partials >= 1, "sbr_bn_finalize: bad arguments");
  SBR_CHECK_CUDA(sbr_launch(bn_finalize_kernel, dim3(cdiv(C, 8)), dim3(1024), (size_t)0, S(stream), stats, n_partials,
                            n_rows, C, eps, momentum, mean_invstd, running_mean, running_var, num_batches_tracked));
  return SBR_OK;
}

extern "C" int sbr_bn_eval_coeffs(const float* running_mean, const float* running_var, int C, float eps,
                                  float* mean_invstd, void* stream) {
  SBR_REQUIRE(running_mean && running_var && mean_invstd && C > 0, "sbr_bn_eval_coeffs: bad arguments");
  bn_eval_coeffs_kernel<<<cdiv(C, 128), 128, 0, S(stream)>>>(running_mean, running_var, C, eps, mean_invstd);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_bn_apply(const float* z, int64_t ld_z, const float* mean_invstd, const float* gamma,
                            const float* beta, int act, int64_t rows, int C, void* out_bf16, int64_t ld_bf16,
                            float* out_f32, int64_t ld_f32, void* stream) {
  SBR_REQUIRE(z && mean_invstd && gamma && beta && rows > 0 && C > 0, "sbr_bn_apply: bad arguments");
  {
    const auto al = [](const void* p, int a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
    const bool vec = (C & 3) == 0 && C <= 1024 && 256 % (C >> 2) == 0 && (ld_z & 3) == 0 && al(z, 16) &&
                     al(mean_invstd, 16) && al(gamma, 16) && al(beta, 16) &&
                     (!out_bf16 || ((ld_bf16 & 3) == 0 && al(out_bf16, 8))) &&
                     (!out_f32 || ((ld_f32 & 3) == 0 && al(out_f32, 16))) && rows >= 1024 &&
                     getenv("SBR_NORM_SCALAR") == nullptr;
    if (vec) {
      const int rpb = 256 / (C >> 2);
      int64_t blocks = cdiv(rows, (int64_t)rpb * BNV_UNR);
      const int64_t cap = (int64_t)sbr_num_sms() * 8;
      if (blocks > cap) blocks = cap;
      SBR_CHECK_CUDA(sbr_launch(bn_apply_vec_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), S(stream), z, ld_z,
                                mean_invstd, gamma, beta, act, rows, C, reinterpret_cast<bf16*>(out_bf16), ld_bf16,
                                out_f32, ld_f32));
      SBR_LAUNCH_CHECK();
      return SBR_OK;
    }
  }
  SBR_CHECK_CUDA(sbr_launch(bn_apply_kernel, dim3(tile_grid(rows, C)), dim3(256), (size_t)(0), S(stream), z, ld_z, mean_invstd, gamma, beta, act, rows, C,
                                                             reinterpret_cast<bf16*>(out_bf16), ld_bf16, out_f32,
                                                             ld_f32, rows_per_block(rows)));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_bn_bwd_reduce(const float* dy, int64_t ld_dy, const float* y_f32, const void* y_bf16, int64_t ld_y,
                                 int act, const float* z, int64_t ld_z, const float* mean_invstd, int64_t rows, int C,
                                 float* sums, void* stream) {
  SBR_REQUIRE(dy && z && mean_invstd && sums && rows > 0 && C > 0, "sbr_bn_bwd_reduce: bad arguments");
  SBR_REQUIRE(act == SBR_ACT_NONE || y_f32 || y_bf16, "sbr_bn_bwd_reduce: activation gradient needs the output y");
  {
    const auto al = [](const void* p, int a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
    const bool y_ok = act == SBR_ACT_NONE || ((ld_y & 3) == 0 && (y_f32 ? al(y_f32, 16) : al(y_bf16, 8)));
    const bool vec = y_ok && (C & 3) == 0 && C <= 1024 && 256 % (C >> 2) == 0 && (ld_dy & 3) == 0 && (ld_z & 3) == 0 &&
                     al(dy, 16) && al(z, 16) && al(mean_invstd, 16) && rows >= 1024 &&
                     getenv("SBR_NORM_SCALAR") == nullptr;
    if (vec) {
      const int rpb = 256 / (C >> 2);
      int64_t blocks = cdiv(rows, (int64_t)rpb * BNV_UNR);
      const int64_t cap = (int64_t)sbr_num_sms() * 4;  // (one atomic per column and block)
      if (blocks > cap) blocks = cap;
      SBR_CHECK_CUDA(sbr_launch(bn_bwd_reduce_vec_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), S(stream), dy,
                                ld_dy, y_f32, reinterpret_cast<const bf16*>(y_bf16), ld_y, act, z, ld_z, mean_invstd, rows,
                                C, sums));
      SBR_LAUNCH_CHECK();
      return SBR_OK;
    }
  }
  SBR_CHECK_CUDA(sbr_launch(bn_bwd_reduce_kernel, dim3(tile_grid(rows, C)), dim3(256), (size_t)(0), S(stream), 
      dy, ld_dy, y_f32, reinterpret_cast<const bf16*>(y_bf16), ld_y, act, z, ld_z, mean_invstd, rows, C, sums,
      rows_per_block(rows)));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

// Vector path of the above for the shapes of the train step (no activation in front, C % 4 == 0, 16-byte aligned
// rows, 256 % (C / 4) == 0): a thread owns 4 columns (its BatchNorm coefficients live in registers), grid-strides over
// the rows and keeps the loads of UNR rows in flight -- the scalar kernel (one 4-byte load per thread and row, no
// unrolling) ran at 2.4 TB/s.
__global__ void __launch_bounds__(256)
bn_bwd_apply_vec_kernel(const float* __restrict__ dy, int64_t ld_dy, const float* __restrict__ y_f32,
                        const bf16* __restrict__ y_bf16, int64_t ld_y, int act, const float* __restrict__ z, int64_t ld_z,
                        const float* __restrict__ mean_invstd, const float* __restrict__ gamma,
                        const float* __restrict__ sums, int n_replicas, int64_t rows, int C,
                        bf16* __restrict__ dz_bf16, int64_t ld_dz, float* __restrict__ dz_f32, int64_t ld_dz_f32,
                        float* dgamma, float* dbeta) {
  SBR_PDL_ENTRY();
  const int c4n = C >> 2;                 // float4 columns per row
  const int c = (threadIdx.x % c4n) * 4;  // this thread's first column (256 % c4n == 0)
  const int rpb = 256 / c4n;              // rows per block pass
  const float inv_n = 1.f / (float)rows;
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < n_replicas; ++r) {
    const float4 a = *reinterpret_cast<const float4*>(sums + (size_t)r * 2 * C + c);
    const float4 b = *reinterpret_cast<const float4*>(sums + (size_t)r * 2 * C + C + c);
    s0[0] += a.x; s0[1] += a.y; s0[2] += a.z; s0[3] += a.w;
    s1[0] += b.x; s1[1] += b.y; s1[2] += b.z; s1[3] += b.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < c4n) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (dbeta) dbeta[c + j] += s0[j];
      if (dgamma) dgamma[c + j] += s1[j];
    }
  }
  const float4 mean = *reinterpret_cast<const float4*>(mean_invstd + c);
  const float4 invstd = *reinterpret_cast<const float4*>(mean_invstd + C + c);
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
  const float m[4] = {mean.x, mean.y, mean.z, mean.w}, is[4] = {invstd.x, invstd.y, invstd.z, invstd.w};
  const float gi[4] = {gm.x * invstd.x, gm.y * invstd.y, gm.z * invstd.z, gm.w * invstd.w};
  float k0[4], k1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    k0[j] = s0[j] * inv_n;
    k1[j] = s1[j] * inv_n;
  }
  const int64_t stride = (int64_t)gridDim.x * rpb;
  for (int64_t r0 = (int64_t)blockIdx.x * rpb + threadIdx.x / c4n; r0 < rows; r0 += stride * BNV_UNR) {
    float4 g[BNV_UNR], zz[BNV_UNR], yv[BNV_UNR];
#pragma unroll
    for (int u = 0; u < BNV_UNR; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < rows) {
        g[u] = __ldcs(reinterpret_cast<const float4*>(dy + r * ld_dy + c));  // read once: streaming
        zz[u] = __ldcs(reinterpret_cast<const float4*>(z + r * ld_z + c));
        if (act != SBR_ACT_NONE) yv[u] = load_y4(y_f32, y_bf16, r * ld_y + c);
      }
    }
#pragma unroll
    for (int u = 0; u < BNV_UNR; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= rows) break;
      float gv[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
      const float zv[4] = {zz[u].x, zz[u].y, zz[u].z, zz[u].w};
      if (act != SBR_ACT_NONE) {
        gv[0] *= act_grad_from_out(act, yv[u].x);
        gv[1] *= act_grad_from_out(act, yv[u].y);
        gv[2] *= act_grad_from_out(act, yv[u].z);
        gv[3] *= act_grad_from_out(act, yv[u].w);
      }
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float xh = (zv[j] - m[j]) * is[j];
        v[j] = gi[j] * (gv[j] - k0[j] - xh * k1[j]);
      }
      if (dz_bf16) {
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(dz_bf16 + r * ld_dz + c) = o;
      }
      if (dz_f32) *reinterpret_cast<float4*>(dz_f32 + r * ld_dz_f32 + c) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

extern "C" int sbr_bn_bwd_apply(const float* dy, int64_t ld_dy, const float* y_f32, const void* y_bf16, int64_t ld_y,
                                int act, const float* z, int64_t ld_z, const float* mean_invstd, const float* gamma,
                                const float* sums, int n_replicas, int64_t rows, int C, void* dz_bf16, int64_t ld_dz,
                                float* dz_f32, int64_t ld_dz_f32, float* dgamma, float* dbeta, void* stream) {
  SBR_REQUIRE(dy && z && mean_invstd && gamma && sums && rows > 0 && C > 0 && n_replicas >= 1,
              "sbr_bn_bwd_apply: bad arguments");
  SBR_REQUIRE(act == SBR_ACT_NONE || y_f32 || y_bf16, "sbr_bn_bwd_apply: activation gradient needs the output y");
  const auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool y_ok = act == SBR_ACT_NONE || ((ld_y & 3) == 0 && (y_f32 ? al16(y_f32) : (reinterpret_cast<uintptr_t>(y_bf16) & 7) == 0));
  const bool vec = y_ok && (C & 3) == 0 && C <= 1024 && 256 % (C >> 2) == 0 && (ld_dy & 3) == 0 &&
                   (ld_z & 3) == 0 && al16(dy) && al16(z) && al16(mean_invstd) && al16(gamma) && al16(sums) &&
                   (!dz_bf16 || ((ld_dz & 3) == 0 && (reinterpret_cast<uintptr_t>(dz_bf16) & 7) == 0)) &&
                   (!dz_f32 || ((ld_dz_f32 & 3) == 0 && al16(dz_f32))) && getenv("SBR_NORM_SCALAR") == nullptr;
  if (vec) {
    const int rpb = 256 / (C >> 2);
    int64_t blocks = cdiv(rows, (int64_t)rpb * BNV_UNR);
    const int64_t cap = (int64_t)sbr_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    SBR_CHECK_CUDA(sbr_launch(bn_bwd_apply_vec_kernel, dim3((unsigned)blocks), dim3(256), (size_t)(0), S(stream), dy,
                              ld_dy, y_f32, reinterpret_cast<const bf16*>(y_bf16), ld_y, act, z, ld_z, mean_invstd, gamma,
                              sums, n_replicas, rows, C,
                              reinterpret_cast<bf16*>(dz_bf16), ld_dz, dz_f32, ld_dz_f32, dgamma, dbeta));
    SBR_LAUNCH_CHECK();
    return SBR_OK;
  }
  SBR_CHECK_CUDA(sbr_launch(bn_bwd_apply_kernel, dim3(tile_grid(rows, C)), dim3(256), (size_t)(0), S(stream), 
      dy, ld_dy, y_f32, reinterpret_cast<const bf16*>(y_bf16), ld_y, act, z, ld_z, mean_invstd, gamma, sums, n_replicas,
      rows, C, reinterpret_cast<bf16*>(dz_bf16), ld_dz, dz_f32, ld_dz_f32, dgamma, dbeta, rows_per_block(rows)));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}
