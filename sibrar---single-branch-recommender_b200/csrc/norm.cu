// sibrar_b200 -- BatchNorm1d (train statistics finalize / apply / backward) and the table-level activation
// gradient with column sums.  Thread block = 8 row-lanes x 32 consecutive columns; grid = row blocks x column chunks,
// so every warp reads 128 contiguous bytes per row and column partial sums stay in one register per thread.
// Replaces torch.nn.BatchNorm1d as used at modules/polylinear.py:58-61,68-69 and algorithms/sgd_alg.py:1834-1837,
// and the autograd of the projection's output activation (algorithms/sgd_alg.py:1356).
#include <stdlib.h>

#include "common.cuh"

namespace {
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
  __shared__ float red[2][32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    int r = ty;
    for (; r + 96 < n_partials; r += 128) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] += stats[(size_t)(r + 32 * q) * 2 * C + c];
        b[q] += stats[(size_t)(r + 32 * q) * 2 * C + C + c];
      }
    }
    for (int q = 0; r < n_partials; r += 32, ++q) {
      a[q] += stats[(size_t)r * 2 * C + c];
      b[q] += stats[(size_t)r * 2 * C + C + c];
    }
  }
  red[0][ty][tx] = (a[0] + a[1]) + (a[2] + a[3]);
  red[1][ty][tx] = (b[0] + b[1]) + (b[2] + b[3]);
  __syncthreads();
  if (ty != 0 || c >= C) return;
  double s1 = 0., s2 = 0.;
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    s1 += (double)red[0][q][tx];
    s2 += (double)red[1][q][tx];
  }
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

extern "C" int sbr_actgrad_colsum(float* dy, int64_t ld_dy, const float* y_f32, const void* y_bf16, int64_t ld_y,
                                  int act, int64_t rows, int64_t cols, void* out_bf16, int64_t ld_out, float* out_f32,
                                  int64_t ld_out_f32, float* colsum, int zero_dy, void* stream) {
  SBR_REQUIRE(dy && rows > 0 && cols > 0, "sbr_actgrad_colsum: bad arguments");
  SBR_REQUIRE(act == SBR_ACT_NONE || y_f32 || y_bf16, "sbr_actgrad_colsum: activation gradient needs the output y");
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
  SBR_CHECK_CUDA(sbr_launch(bn_finalize_kernel, dim3(cdiv(C, 32)), dim3(1024), (size_t)0, S(stream), stats, n_partials,
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
  SBR_CHECK_CUDA(sbr_launch(bn_bwd_reduce_kernel, dim3(tile_grid(rows, C)), dim3(256), (size_t)(0), S(stream), 
      dy, ld_dy, y_f32, reinterpret_cast<const bf16*>(y_bf16), ld_y, act, z, ld_z, mean_invstd, rows, C, sums,
      rows_per_block(rows)));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_bn_bwd_apply(const float* dy, int64_t ld_dy, const float* y_f32, const void* y_bf16, int64_t ld_y,
                                int act, const float* z, int64_t ld_z, const float* mean_invstd, const float* gamma,
                                const float* sums, int n_replicas, int64_t rows, int C, void* dz_bf16, int64_t ld_dz,
                                float* dz_f32, int64_t ld_dz_f32, float* dgamma, float* dbeta, void* stream) {
  SBR_REQUIRE(dy && z && mean_invstd && gamma && sums && rows > 0 && C > 0 && n_replicas >= 1,
              "sbr_bn_bwd_apply: bad arguments");
  SBR_REQUIRE(act == SBR_ACT_NONE || y_f32 || y_bf16, "sbr_bn_bwd_apply: activation gradient needs the output y");
  SBR_CHECK_CUDA(sbr_launch(bn_bwd_apply_kernel, dim3(tile_grid(rows, C)), dim3(256), (size_t)(0), S(stream), 
      dy, ld_dy, y_f32, reinterpret_cast<const bf16*>(y_bf16), ld_y, act, z, ld_z, mean_invstd, gamma, sums, n_replicas,
      rows, C, reinterpret_cast<bf16*>(dz_bf16), ld_dz, dz_f32, ld_dz_f32, dgamma, dbeta, rows_per_block(rows)));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}
