// sibrar_b200 -- symmetric InfoNCE (CLIP-style) modality-alignment loss with hand-written backward.
#include <stdarg.h>

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

}  // namespace

extern "C" int sbr_infonce(const float* e, int64_t G, int64_t n, int D, float temperature, float weight,
                           double* loss_acc, float* de, int accumulate, float* lse_ws, void* stream) {
  SBR_REQUIRE(e && lse_ws && G > 0 && n > 0, "sbr_infonce: bad arguments");
  SBR_REQUIRE(D > 0 && D <= 512, "sbr_infonce: D=%d not in [1, 512]", D);
  SBR_REQUIRE(temperature > 0.f, "sbr_infonce: temperature must be positive");
  const int64_t warps = 2 * G * n;
  DISPATCH_NV(D, 32, {
    infonce_lse_kernel<NVv><<<cdiv(warps, 8), 256, 0, S(stream)>>>(e, G, n, D, 1.f / temperature, lse_ws);
    infonce_grad_kernel<NVv><<<cdiv(warps, 8), 256, 0, S(stream)>>>(e, G, n, D, 1.f / temperature, weight, lse_ws,
                                                                    loss_acc, de, accumulate);
  });
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

