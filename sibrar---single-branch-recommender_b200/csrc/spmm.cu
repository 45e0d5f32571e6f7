// sibrar_b200 -- CSR SpMM for the 'interactions' modality (forward projection and its wgrad through the transposed CSR).
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

// ------------------------------------------------------------------------------------------------ CSR SpMM
// out[r, c] = act(sum_{p in row r} dense[indices[p], c] + bias[c]);  one warp per row, lanes across columns.
template <int NV>
__global__ void spmm_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t rows,
                            const float* __restrict__ dense, int64_t ld_dense, int C, const float* __restrict__ bias,
                            int act, float* __restrict__ out, int64_t ld_out, int transpose_out) {
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  const int64_t beg = indptr[row], end = indptr[row + 1];
  for (int64_t p = beg; p < end; p += 32) {
    int32_t my = (p + lane < end) ? indices[p + lane] : -1;
    int cnt = (int)min((int64_t)32, end - p);
    for (int t = 0; t < cnt; ++t) {
      int32_t j = __shfl_sync(0xffffffffu, my, t);
      const float* d = dense + (int64_t)j * ld_dense;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        int c = lane + 32 * i;
        if (c < C) acc[i] += __ldg(d + c);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int c = lane + 32 * i;
    if (c < C) {
      float v = acc[i] + (bias ? bias[c] : 0.f);
      v = act_fwd(act, v);
      if (transpose_out) out[(int64_t)c * ld_out + row] = v;
      else out[row * ld_out + c] = v;
    }
  }
}

}  // namespace

extern "C" int sbr_spmm_csr(const int64_t* indptr, const int32_t* indices, int64_t rows, const float* dense,
                            int64_t ld_dense, int64_t C, const float* bias, int act, float* out, int64_t ld_out,
                            int transpose_out, void* stream) {
  SBR_REQUIRE(indptr && indices && dense && out && rows > 0, "sbr_spmm_csr: bad arguments");
  SBR_REQUIRE(C > 0 && C <= 768, "sbr_spmm_csr: C=%lld not in [1, 768]", (long long)C);
  DISPATCH_NV(C, 32, spmm_kernel<NVv><<<cdiv(rows, 8), 256, 0, S(stream)>>>(
                         indptr, indices, rows, dense, ld_dense, (int)C, bias, act, out, ld_out, transpose_out));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

