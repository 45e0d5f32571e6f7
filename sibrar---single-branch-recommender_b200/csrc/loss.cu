// sibrar_b200 -- fused modality aggregation + user-item scoring + rec loss (+ gradients), InfoNCE.
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

// ------------------------------------------------------------------------------------------------ score + loss
// one warp per interaction.  lane owns d = lane + 32*i.
template <int NV>
__device__ __forceinline__ void aggregate_slots(const float* __restrict__ e, int k, int D, int lane, int agg_max,
                                                float (&out)[NV], uint32_t (&argmax)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int d = lane + 32 * i;
    float a = 0.f;
    uint32_t am = 0;
    if (d < D) {
      if (agg_max) {
        a = e[d];
        for (int s = 1; s < k; ++s) {
          float v = e[(int64_t)s * D + d];
          if (v > a) { a = v; am = s; }
        }
      } else {
        for (int s = 0; s < k; ++s) a += e[(int64_t)s * D + d];
        a /= (float)k;
      }
    }
    out[i] = a;
    argmax[i] = am;
  }
}

template <int NV>
__global__ void score_loss_kernel(const float* __restrict__ eu, const float* __restrict__ ei, int64_t B, int n, int ku,
                                  int ki, int D, int agg_max_user, int agg_max_item, int loss_kind, float inv_cnt,
                                  float ssm_shift, float* __restrict__ logits, double* __restrict__ loss_acc,
                                  float* __restrict__ deu, float* __restrict__ dei, float* __restrict__ u_agg_out,
                                  float* __restrict__ i_agg_out, const float* __restrict__ dlogits_in) {
  extern __shared__ float sh[];  // per warp: n scores + n grads
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t b = blockIdx.x * (int64_t)(blockDim.x >> 5) + wib;
  if (b >= B) return;
  float* sc = sh + (size_t)wib * 2 * n;
  float* gr = sc + n;
  float u[NV];
  uint32_t uam[NV];
  aggregate_slots<NV>(eu + b * ku * D, ku, D, lane, agg_max_user, u, uam);
  if (u_agg_out) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < D) u_agg_out[b * D + lane + 32 * i] = u[i];
  }
  for (int j = 0; j < n; ++j) {
    float it[NV];
    uint32_t iam[NV];
    aggregate_slots<NV>(ei + ((b * n + j) * ki) * D, ki, D, lane, agg_max_item, it, iam);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) dot += u[i] * it[i];
    dot = warp_sum(dot);
    if (lane == 0) sc[j] = dot;
    if (i_agg_out) {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (lane + 32 * i < D) i_agg_out[(b * n + j) * D + lane + 32 * i] = it[i];
    }
  }
  __syncwarp();
  // ---- loss and d loss / d score (lanes stride over j)
  float lsum = 0.f;
  if (dlogits_in != nullptr) {  // gradients of an externally computed loss (autograd path): d loss / d score given
    for (int j = lane; j < n; j += 32) gr[j] = dlogits_in[b * n + j];
  } else if (loss_kind == SBR_LOSS_BPR) {
    float s0 = sc[0], g0 = 0.f;
    for (int j = 1 + lane; j < n; j += 32) {
      float d = s0 - sc[j];
      lsum += (d > 0.f ? log1pf(__expf(-d)) : -d + log1pf(__expf(d)));
      float sg = 1.f / (1.f + __expf(d));  // sigmoid(-d)
      gr[j] = sg * inv_cnt;
      g0 -= sg * inv_cnt;
    }
    g0 = warp_sum(g0);
    if (lane == 0) gr[0] = g0;
  } else if (loss_kind == SBR_LOSS_BCE) {
    for (int j = lane; j < n; j += 32) {
      float s = sc[j], y = (j == 0) ? 1.f : 0.f;
      lsum += (s > 0.f ? s + log1pf(__expf(-s)) : log1pf(__expf(s))) - y * s;
      gr[j] = (1.f / (1.f + __expf(-s)) - y) * inv_cnt;
    }
  } else {
    float mx = -INFINITY;
    for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sc[j] + (j > 0 ? ssm_shift : 0.f));
    mx = warp_max(mx);
    float se = 0.f;
    for (int j = lane; j < n; j += 32) se += __expf(sc[j] + (j > 0 ? ssm_shift : 0.f) - mx);
    se = warp_sum(se);
    float lse = mx + logf(se);
    for (int j = lane; j < n; j += 32) {
      float zj = sc[j] + (j > 0 ? ssm_shift : 0.f);
      gr[j] = (__expf(zj - lse) - (j == 0 ? 1.f : 0.f)) * inv_cnt;
    }
    if (lane == 0) lsum = lse - sc[0];
  }
  lsum = warp_sum(lsum);
  if (lane == 0) {
    if (loss_acc) atomicAdd(loss_acc, (double)lsum * (double)inv_cnt);
  }
  if (logits) {
    for (int j = lane; j < n; j += 32) logits[b * n + j] = sc[j];
  }
  __syncwarp();
  if (deu == nullptr || dei == nullptr) return;
  // ---- gradients
  float du[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) du[i] = 0.f;
  for (int j = 0; j < n; ++j) {
    const float g = gr[j];
    float it[NV];
    uint32_t iam[NV];
    const float* ej = ei + ((b * n + j) * ki) * D;
    aggregate_slots<NV>(ej, ki, D, lane, agg_max_item, it, iam);
    float* dj = dei + ((b * n + j) * ki) * D;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int d = lane + 32 * i;
      if (d < D) {
        du[i] += g * it[i];
        float gi = g * u[i];
        for (int s = 0; s < ki; ++s)
          dj[(int64_t)s * D + d] = agg_max_item ? ((uint32_t)s == iam[i] ? gi : 0.f) : gi / (float)ki;
      }
    }
  }
  float* dub = deu + b * ku * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int d = lane + 32 * i;
    if (d < D)
      for (int s = 0; s < ku; ++s)
        dub[(int64_t)s * D + d] = agg_max_user ? ((uint32_t)s == uam[i] ? du[i] : 0.f) : du[i] / (float)ku;
  }
}

// Fast path (D in {16, 32, 64, 128}; any number of modality slots per entity): LPR = D / 4 lanes hold one row as
// float4, so one load instruction covers 32 / LPR item rows; the k slots of a row are aggregated (mean / max) in
// registers and re-read (L1 / L2 hits) for the gradient; the block's loss goes out with one atomic.
__device__ __forceinline__ float4 agg_slots4(const float4* __restrict__ row, int k, int stride4, int agg_max) {
  float4 a = __ldg(row);
  for (int s = 1; s < k; ++s) {
    const float4 b = __ldg(row + (size_t)s * stride4);
    if (agg_max) {
      a.x = b.x > a.x ? b.x : a.x; a.y = b.y > a.y ? b.y : a.y; a.z = b.z > a.z ? b.z : a.z; a.w = b.w > a.w ? b.w : a.w;
    } else {
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
  }
  if (!agg_max && k > 1) {
    a.x = a.x / (float)k; a.y = a.y / (float)k; a.z = a.z / (float)k; a.w = a.w / (float)k;  // (like the scalar kernel)
  }
  return a;
}
// gradient g (w.r.t. the aggregated row) -> the k slots: mean: g / k everywhere; max: g at the FIRST slot holding the
// maximum of each element (strict > in the forward scan), 0 elsewhere
__device__ __forceinline__ void scatter_slots4(float4* __restrict__ drow, const float4* __restrict__ row, int k, int stride4,
                                               int agg_max, const float4 g) {
  if (k == 1) {
    *drow = g;
    return;
  }
  if (!agg_max) {
    const float4 q = make_float4(g.x / (float)k, g.y / (float)k, g.z / (float)k, g.w / (float)k);
    for (int s = 0; s < k; ++s) drow[(size_t)s * stride4] = q;
    return;
  }
  float4 best = __ldg(row);
  int ax = 0, ay = 0, az = 0, aw = 0;
  for (int s = 1; s < k; ++s) {
    const float4 b = __ldg(row + (size_t)s * stride4);
    if (b.x > best.x) { best.x = b.x; ax = s; }
    if (b.y > best.y) { best.y = b.y; ay = s; }
    if (b.z > best.z) { best.z = b.z; az = s; }
    if (b.w > best.w) { best.w = b.w; aw = s; }
  }
  for (int s = 0; s < k; ++s)
    drow[(size_t)s * stride4] = make_float4(s == ax ? g.x : 0.f, s == ay ? g.y : 0.f, s == az ? g.z : 0.f, s == aw ? g.w : 0.f);
}

template <int LPR>
__global__ void __launch_bounds__(256)
score_loss_fast_kernel(const float* __restrict__ eu, const float* __restrict__ ei, int64_t B, int n, int ku, int ki,
                       int agg_max_user, int agg_max_item, int loss_kind, float inv_cnt, float ssm_shift,
                       float* __restrict__ logits, double* __restrict__ loss_acc, float* __restrict__ deu,
                       float* __restrict__ dei, const float* __restrict__ dlogits_in) {
  SBR_PDL_ENTRY();
  constexpr int D = 4 * LPR;
  constexpr int RPP = 32 / LPR;  // item rows per pass
  extern __shared__ float sh[];  // per warp: n scores + n grads; then one loss slot per warp
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int64_t b = blockIdx.x * (int64_t)nw + wib;
  float* sc = sh + (size_t)wib * 2 * n;
  float* gr = sc + n;
  float* s_loss = sh + (size_t)nw * 2 * n;
  const int sub = lane / LPR, li = lane % LPR;
  float lsum = 0.f;
  if (b < B) {
    const float4* urow = reinterpret_cast<const float4*>(eu + b * ku * D) + li;
    const float4 u4 = agg_slots4(urow, ku, LPR, agg_max_user);
    const float4* items = reinterpret_cast<const float4*>(ei + b * n * ki * D);
    for (int j0 = 0; j0 < n; j0 += RPP) {
      const int j = j0 + sub;
      float dot = 0.f;
      if (j < n) {
        const float4 v = agg_slots4(items + (size_t)j * ki * LPR + li, ki, LPR, agg_max_item);
        dot = u4.x * v.x + u4.y * v.y + u4.z * v.z + u4.w * v.w;
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if (li == 0 && j < n) sc[j] = dot;
    }
    __syncwarp();
    if (dlogits_in != nullptr) {  // gradients of an externally computed loss (autograd path): d loss / d score given
      for (int j = lane; j < n; j += 32) gr[j] = dlogits_in[b * n + j];
    } else if (loss_kind == SBR_LOSS_BPR) {
      float s0 = sc[0], g0 = 0.f;
      for (int j = 1 + lane; j < n; j += 32) {
        float d = s0 - sc[j];
        lsum += (d > 0.f ? log1pf(__expf(-d)) : -d + log1pf(__expf(d)));
        float sg = 1.f / (1.f + __expf(d));  // sigmoid(-d)
        gr[j] = sg * inv_cnt;
        g0 -= sg * inv_cnt;
      }
      g0 = warp_sum(g0);
      if (lane == 0) gr[0] = g0;
    } else if (loss_kind == SBR_LOSS_BCE) {
      for (int j = lane; j < n; j += 32) {
        float s = sc[j], y = (j == 0) ? 1.f : 0.f;
        lsum += (s > 0.f ? s + log1pf(__expf(-s)) : log1pf(__expf(s))) - y * s;
        gr[j] = (1.f / (1.f + __expf(-s)) - y) * inv_cnt;
      }
    } else {
      float mx = -INFINITY;
      for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sc[j] + (j > 0 ? ssm_shift : 0.f));
      mx = warp_max(mx);
      float se = 0.f;
      for (int j = lane; j < n; j += 32) se += __expf(sc[j] + (j > 0 ? ssm_shift : 0.f) - mx);
      se = warp_sum(se);
      float lse = mx + logf(se);
      for (int j = lane; j < n; j += 32) {
        float zj = sc[j] + (j > 0 ? ssm_shift : 0.f);
        gr[j] = (__expf(zj - lse) - (j == 0 ? 1.f : 0.f)) * inv_cnt;
      }
      if (lane == 0) lsum = lse - sc[0];
    }
    lsum = warp_sum(lsum);
    if (logits) {
      for (int j = lane; j < n; j += 32) logits[b * n + j] = sc[j];
    }
    __syncwarp();
    if (deu != nullptr && dei != nullptr) {
      float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
      float4* dit = reinterpret_cast<float4*>(dei + b * n * ki * D);
      for (int j0 = 0; j0 < n; j0 += RPP) {
        const int j = j0 + sub;
        if (j < n) {
          const float g = gr[j];
          const float4* irow = items + (size_t)j * ki * LPR + li;
          const float4 v = agg_slots4(irow, ki, LPR, agg_max_item);
          du.x += g * v.x; du.y += g * v.y; du.z += g * v.z; du.w += g * v.w;
          scatter_slots4(dit + (size_t)j * ki * LPR + li, irow, ki, LPR, agg_max_item,
                         make_float4(g * u4.x, g * u4.y, g * u4.z, g * u4.w));
        }
      }
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) {
        du.x += __shfl_xor_sync(0xffffffffu, du.x, o);
        du.y += __shfl_xor_sync(0xffffffffu, du.y, o);
        du.z += __shfl_xor_sync(0xffffffffu, du.z, o);
        du.w += __shfl_xor_sync(0xffffffffu, du.w, o);
      }
      if (sub == 0) scatter_slots4(reinterpret_cast<float4*>(deu + b * ku * D) + li, urow, ku, LPR, agg_max_user, du);
    }
  }
  if (lane == 0) s_loss[wib] = lsum;
  __syncthreads();
  if (threadIdx.x == 0 && loss_acc) {
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += (double)s_loss[w];
    atomicAdd(loss_acc, t * (double)inv_cnt);
  }
}

// Fused variant for entities whose single-branch net ends in a BatchNorm (the default configuration): the kernel
// reads the PRE-BatchNorm values z and applies  e = gamma * (z - mean) * invstd + beta  on the fly (the normalised
// embeddings are never written), and it accumulates the two BatchNorm-backward column sums  sum(de), sum(de * xhat)
// of the gradients it produces -- bn_apply and bn_bwd_reduce disappear from the step.  Persistent blocks; the sums
// go out once per block into one of `n_replicas` copies (same-address atomics serialise in one L2 slice).
struct BnInline {
  const float* z;            // [rows, D] pre-BatchNorm values, or nullptr: `e` holds the embeddings themselves
  const float* mean_invstd;  // [2 * D]
  const float* gamma;
  const float* beta;
  float* sums;               // [n_replicas, 2 * D], zeroed by the caller
};

template <int LPR, int MAXP>
__global__ void __launch_bounds__(256)
score_loss_bn_kernel(const float* __restrict__ eu, BnInline bu, const float* __restrict__ ei, BnInline bi, int64_t B,
                     int n, int loss_kind, float inv_cnt, float ssm_shift, float* __restrict__ logits,
                     double* __restrict__ loss_acc, float* __restrict__ deu, float* __restrict__ dei, int n_replicas) {
  SBR_PDL_ENTRY();
  constexpr int D = 4 * LPR;
  constexpr int RPP = 32 / LPR;  // item rows per pass
  extern __shared__ float sh[];  // per warp: n scores + n grads; then nw loss slots; then nw x 4 x D column sums
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float* sc = sh + (size_t)wib * 2 * n;
  float* gr = sc + n;
  float* s_loss = sh + (size_t)nw * 2 * n;
  float* s_sums = s_loss + nw;
  const int sub = lane / LPR, li = lane % LPR;
  const float4* usrc = reinterpret_cast<const float4*>(bu.z ? bu.z : eu);
  const float4* isrc = reinterpret_cast<const float4*>(bi.z ? bi.z : ei);
  // per-lane BatchNorm coefficients of its 4 columns: e = z * a + c,  xhat = (z - mean) * invstd
  float4 ua = make_float4(1.f, 1.f, 1.f, 1.f), uc = make_float4(0.f, 0.f, 0.f, 0.f), um = uc, uis = ua;
  float4 ia = ua, ic = uc, im = uc, iis = ua;
  if (bu.z) {
    um = reinterpret_cast<const float4*>(bu.mean_invstd)[li];
    uis = reinterpret_cast<const float4*>(bu.mean_invstd + D)[li];
    const float4 g = reinterpret_cast<const float4*>(bu.gamma)[li], b = reinterpret_cast<const float4*>(bu.beta)[li];
    ua = make_float4(g.x * uis.x, g.y * uis.y, g.z * uis.z, g.w * uis.w);
    uc = make_float4(b.x - um.x * ua.x, b.y - um.y * ua.y, b.z - um.z * ua.z, b.w - um.w * ua.w);
  }
  if (bi.z) {
    im = reinterpret_cast<const float4*>(bi.mean_invstd)[li];
    iis = reinterpret_cast<const float4*>(bi.mean_invstd + D)[li];
    const float4 g = reinterpret_cast<const float4*>(bi.gamma)[li], b = reinterpret_cast<const float4*>(bi.beta)[li];
    ia = make_float4(g.x * iis.x, g.y * iis.y, g.z * iis.z, g.w * iis.w);
    ic = make_float4(b.x - im.x * ia.x, b.y - im.y * ia.y, b.z - im.z * ia.z, b.w - im.w * ia.w);
  }
  float4 su0 = make_float4(0.f, 0.f, 0.f, 0.f), su1 = su0, si0 = su0, si1 = su0;  // sum(de), sum(de * xhat)
  float lsum_total = 0.f;
  for (int64_t b = blockIdx.x * (int64_t)nw + wib; b < B; b += (int64_t)gridDim.x * nw) {
    const float4 zu = __ldg(usrc + b * LPR + li);
    const float4 u4 = make_float4(zu.x * ua.x + uc.x, zu.y * ua.y + uc.y, zu.z * ua.z + uc.z, zu.w * ua.w + uc.w);
    const float4* items = isrc + b * n * LPR;
    // MAXP > 0 (n <= MAXP * RPP): all item rows of the user are loaded at once and stay in registers for the
    // gradient phase -- one 512-byte load in flight per warp (the loop below) caps the kernel at ~2.4 TB/s
    float4 zr[MAXP > 0 ? MAXP : 1];
    if (MAXP > 0) {
#pragma unroll
      for (int p = 0; p < MAXP; ++p) {
        const int j = p * RPP + sub;
        zr[p] = (p * RPP < n && j < n) ? __ldg(items + (size_t)j * LPR + li) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int p = 0; p < MAXP; ++p) {
        if (p * RPP >= n) break;  // warp-uniform
        const int j = p * RPP + sub;
        const float4 z = zr[p];
        float dot = u4.x * (z.x * ia.x + ic.x) + u4.y * (z.y * ia.y + ic.y) + u4.z * (z.z * ia.z + ic.z) +
                    u4.w * (z.w * ia.w + ic.w);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (li == 0 && j < n) sc[j] = dot;
      }
    } else {
      for (int j0 = 0; j0 < n; j0 += RPP) {
        const int j = j0 + sub;
        float dot = 0.f;
        if (j < n) {
          const float4 z = __ldg(items + (size_t)j * LPR + li);
          dot = u4.x * (z.x * ia.x + ic.x) + u4.y * (z.y * ia.y + ic.y) + u4.z * (z.z * ia.z + ic.z) +
                u4.w * (z.w * ia.w + ic.w);
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (li == 0 && j < n) sc[j] = dot;
      }
    }
    __syncwarp();
    float lsum = 0.f;
    if (loss_kind == SBR_LOSS_BPR) {
      float s0 = sc[0], g0 = 0.f;
      for (int j = 1 + lane; j < n; j += 32) {
        float d = s0 - sc[j];
        lsum += (d > 0.f ? log1pf(__expf(-d)) : -d + log1pf(__expf(d)));
        float sg = 1.f / (1.f + __expf(d));  // sigmoid(-d)
        gr[j] = sg * inv_cnt;
        g0 -= sg * inv_cnt;
      }
      g0 = warp_sum(g0);
      if (lane == 0) gr[0] = g0;
    } else if (loss_kind == SBR_LOSS_BCE) {
      for (int j = lane; j < n; j += 32) {
        float s = sc[j], y = (j == 0) ? 1.f : 0.f;
        lsum += (s > 0.f ? s + log1pf(__expf(-s)) : log1pf(__expf(s))) - y * s;
        gr[j] = (1.f / (1.f + __expf(-s)) - y) * inv_cnt;
      }
    } else {
      float mx = -INFINITY;
      for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sc[j] + (j > 0 ? ssm_shift : 0.f));
      mx = warp_max(mx);
      float se = 0.f;
      for (int j = lane; j < n; j += 32) se += __expf(sc[j] + (j > 0 ? ssm_shift : 0.f) - mx);
      se = warp_sum(se);
      float lse = mx + logf(se);
      for (int j = lane; j < n; j += 32) {
        float zj = sc[j] + (j > 0 ? ssm_shift : 0.f);
        gr[j] = (__expf(zj - lse) - (j == 0 ? 1.f : 0.f)) * inv_cnt;
      }
      if (lane == 0) lsum = lse - sc[0];
    }
    lsum_total += lsum;  // (lane-partial; reduced once at the end)
    if (logits) {
      for (int j = lane; j < n; j += 32) logits[b * n + j] = sc[j];
    }
    __syncwarp();
    float4 du = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* dit = reinterpret_cast<float4*>(dei + b * n * D);
    auto grad_row = [&](int j, const float4 z) {
      const float g = gr[j];
      du.x += g * (z.x * ia.x + ic.x); du.y += g * (z.y * ia.y + ic.y);
      du.z += g * (z.z * ia.z + ic.z); du.w += g * (z.w * ia.w + ic.w);
      const float4 d = make_float4(g * u4.x, g * u4.y, g * u4.z, g * u4.w);
      dit[(size_t)j * LPR + li] = d;
      si0.x += d.x; si0.y += d.y; si0.z += d.z; si0.w += d.w;
      si1.x += d.x * (z.x - im.x) * iis.x; si1.y += d.y * (z.y - im.y) * iis.y;
      si1.z += d.z * (z.z - im.z) * iis.z; si1.w += d.w * (z.w - im.w) * iis.w;
    };
    if (MAXP > 0) {
#pragma unroll
      for (int p = 0; p < MAXP; ++p) {
        const int j = p * RPP + sub;
        if (p * RPP < n && j < n) grad_row(j, zr[p]);
      }
    } else {
      for (int j0 = 0; j0 < n; j0 += RPP) {
        const int j = j0 + sub;
        if (j < n) grad_row(j, __ldg(items + (size_t)j * LPR + li));
      }
    }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      du.x += __shfl_xor_sync(0xffffffffu, du.x, o);
      du.y += __shfl_xor_sync(0xffffffffu, du.y, o);
      du.z += __shfl_xor_sync(0xffffffffu, du.z, o);
      du.w += __shfl_xor_sync(0xffffffffu, du.w, o);
    }
    if (sub == 0) {
      reinterpret_cast<float4*>(deu + b * D)[li] = du;
      su0.x += du.x; su0.y += du.y; su0.z += du.z; su0.w += du.w;
      su1.x += du.x * (zu.x - um.x) * uis.x; su1.y += du.y * (zu.y - um.y) * uis.y;
      su1.z += du.z * (zu.z - um.z) * uis.z; su1.w += du.w * (zu.w - um.w) * uis.w;
    }
    __syncwarp();
  }
  // ---- block reductions: loss, BatchNorm-backward column sums
  lsum_total = warp_sum(lsum_total);
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) {  // item sums live in every sub-group
    si0.x += __shfl_xor_sync(0xffffffffu, si0.x, o); si0.y += __shfl_xor_sync(0xffffffffu, si0.y, o);
    si0.z += __shfl_xor_sync(0xffffffffu, si0.z, o); si0.w += __shfl_xor_sync(0xffffffffu, si0.w, o);
    si1.x += __shfl_xor_sync(0xffffffffu, si1.x, o); si1.y += __shfl_xor_sync(0xffffffffu, si1.y, o);
    si1.z += __shfl_xor_sync(0xffffffffu, si1.z, o); si1.w += __shfl_xor_sync(0xffffffffu, si1.w, o);
  }
  if (lane == 0) s_loss[wib] = lsum_total;
  if (sub == 0) {
    float4* row = reinterpret_cast<float4*>(s_sums + (size_t)wib * 4 * D);
    row[li] = su0; row[LPR + li] = su1; row[2 * LPR + li] = si0; row[3 * LPR + li] = si1;
  }
  __syncthreads();
  if (threadIdx.x == 0 && loss_acc) {
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += (double)s_loss[w];
    atomicAdd(loss_acc, t * (double)inv_cnt);
  }
  const int rep = blockIdx.x % n_replicas;
  for (int i = threadIdx.x; i < 4 * D; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += s_sums[(size_t)w * 4 * D + i];
    const int which = i / (2 * D), c = i % (2 * D);  // 0: user (sum, sum*xhat), 1: item
    float* dst = which == 0 ? bu.sums : bi.sums;
    if (dst != nullptr && t != 0.f) atomicAdd(dst + (size_t)rep * 2 * D + c, t);
  }
}

__global__ void aggregate_kernel(const float* __restrict__ e, int64_t rows, int k, int D, int agg_max,
                                 float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int64_t ld_bf16) {
  int64_t total = rows * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / D;
    int d = (int)(i - r * D);
    const float* p = e + r * k * D + d;
    float a = p[0];
    for (int s = 1; s < k; ++s) {
      float v = p[(int64_t)s * D];
      a = agg_max ? fmaxf(a, v) : a + v;
    }
    if (!agg_max) a /= (float)k;
    if (out_f32) out_f32[i] = a;
    if (out_bf16) out_bf16[r * ld_bf16 + d] = __float2bfloat16(a);
  }
}

}  // namespace

extern "C" int sbr_score_loss(const float* eu, const float* ei, int64_t B, int n, int ku, int ki, int D,
                              int agg_max_user, int agg_max_item, int loss_kind, int aggregator_sum, float ssm_shift,
                              float* logits, double* loss_acc, float* deu, float* dei, float* u_agg, float* i_agg,
                              void* stream) {
  SBR_REQUIRE(eu && ei && B > 0 && n >= 1 && ku >= 1 && ki >= 1, "sbr_score_loss: bad arguments");
  SBR_REQUIRE(D > 0 && D <= 512, "sbr_score_loss: D=%d not in [1, 512]", D);
  SBR_REQUIRE(n <= 1024, "sbr_score_loss: at most 1024 items per interaction (got %d)", n);
  SBR_REQUIRE(loss_kind >= 0 && loss_kind <= 2, "sbr_score_loss: unknown loss kind %d", loss_kind);
  SBR_REQUIRE(!(loss_kind == SBR_LOSS_BPR && n < 2), "sbr_score_loss: BPR needs at least one negative");
  double cnt = 1.0;
  if (!aggregator_sum)
    cnt = loss_kind == SBR_LOSS_BPR ? (double)B * (n - 1) : (loss_kind == SBR_LOSS_BCE ? (double)B * n : (double)B);
  const float inv = (float)(1.0 / cnt);
  if (u_agg == nullptr && i_agg == nullptr && (D == 16 || D == 32 || D == 64 || D == 128) &&
      (reinterpret_cast<uintptr_t>(eu) & 15) == 0 && (reinterpret_cast<uintptr_t>(ei) & 15) == 0 &&
      (deu == nullptr || (reinterpret_cast<uintptr_t>(deu) & 15) == 0) &&
      (dei == nullptr || (reinterpret_cast<uintptr_t>(dei) & 15) == 0)) {
    const int nw = 8;
    const size_t sm = ((size_t)nw * 2 * n + nw) * sizeof(float);
    const unsigned blocks = cdiv(B, nw);
#define SBR_FAST(LPR_)                                                                                              \
  SBR_CHECK_CUDA(sbr_launch(score_loss_fast_kernel<LPR_>, dim3(blocks), dim3(nw * 32), sm, S(stream), eu, ei, B, n,  \
                            ku, ki, agg_max_user, agg_max_item, loss_kind, inv, ssm_shift, logits, loss_acc, deu, dei, \
                            (const float*)nullptr))
    if (D == 16) SBR_FAST(4);
    else if (D == 32) SBR_FAST(8);
    else if (D == 64) SBR_FAST(16);
    else SBR_FAST(32);
#undef SBR_FAST
    SBR_LAUNCH_CHECK();
    return SBR_OK;
  }
  const int wpb = 4;
  size_t shmem = (size_t)wpb * 2 * n * sizeof(float);
  DISPATCH_NV(D, 32, score_loss_kernel<NVv><<<cdiv(B, wpb), wpb * 32, shmem, S(stream)>>>(
                         eu, ei, B, n, ku, ki, D, agg_max_user, agg_max_item, loss_kind, (float)(1.0 / cnt), ssm_shift,
                         logits, loss_acc, deu, dei, u_agg, i_agg, nullptr));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_score_bwd(const float* eu, const float* ei, int64_t B, int n, int ku, int ki, int D,
                             int agg_max_user, int agg_max_item, const float* dlogits, float* deu, float* dei,
                             void* stream) {
  SBR_REQUIRE(eu && ei && dlogits && deu && dei && B > 0 && n >= 1 && ku >= 1 && ki >= 1,
              "sbr_score_bwd: bad arguments");
  SBR_REQUIRE(D > 0 && D <= 512 && n <= 1024, "sbr_score_bwd: D=%d / n=%d out of range", D, n);
  if ((D == 16 || D == 32 || D == 64 || D == 128) && (reinterpret_cast<uintptr_t>(eu) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(ei) & 15) == 0 && (reinterpret_cast<uintptr_t>(deu) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dei) & 15) == 0) {
    const int nw = 8;
    const size_t sm = ((size_t)nw * 2 * n + nw) * sizeof(float);
    const unsigned blocks = cdiv(B, nw);
#define SBR_FASTB(LPR_)                                                                                             \
  SBR_CHECK_CUDA(sbr_launch(score_loss_fast_kernel<LPR_>, dim3(blocks), dim3(nw * 32), sm, S(stream), eu, ei, B, n,  \
                            ku, ki, agg_max_user, agg_max_item, (int)SBR_LOSS_BCE, 1.f, 0.f, (float*)nullptr,        \
                            (double*)nullptr, deu, dei, dlogits))
    if (D == 16) SBR_FASTB(4);
    else if (D == 32) SBR_FASTB(8);
    else if (D == 64) SBR_FASTB(16);
    else SBR_FASTB(32);
#undef SBR_FASTB
    SBR_LAUNCH_CHECK();
    return SBR_OK;
  }
  const int wpb = 4;
  size_t shmem = (size_t)wpb * 2 * n * sizeof(float);
  DISPATCH_NV(D, 32, score_loss_kernel<NVv><<<cdiv(B, wpb), wpb * 32, shmem, S(stream)>>>(
                         eu, ei, B, n, ku, ki, D, agg_max_user, agg_max_item, SBR_LOSS_BCE, 1.f, 0.f, nullptr, nullptr,
                         deu, dei, nullptr, nullptr, dlogits));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_score_loss_bn(const float* eu, const sbr_bn_inline_t* bn_u, const float* ei,
                                 const sbr_bn_inline_t* bn_i, int64_t B, int n, int D, int loss_kind,
                                 int aggregator_sum, float ssm_shift, float* logits, double* loss_acc, float* deu,
                                 float* dei, int n_replicas, void* stream) {
  SBR_REQUIRE((eu || (bn_u && bn_u->z)) && (ei || (bn_i && bn_i->z)) && deu && dei && B > 0 && n >= 1,
              "sbr_score_loss_bn: bad arguments");
  SBR_REQUIRE(D == 16 || D == 32 || D == 64 || D == 128, "sbr_score_loss_bn: D=%d not in {16, 32, 64, 128}", D);
  SBR_REQUIRE(n <= 1024 && n_replicas >= 1, "sbr_score_loss_bn: bad n / n_replicas");
  SBR_REQUIRE(loss_kind >= 0 && loss_kind <= 2, "sbr_score_loss_bn: unknown loss kind %d", loss_kind);
  SBR_REQUIRE(!(loss_kind == SBR_LOSS_BPR && n < 2), "sbr_score_loss_bn: BPR needs at least one negative");
  double cnt = 1.0;
  if (!aggregator_sum)
    cnt = loss_kind == SBR_LOSS_BPR ? (double)B * (n - 1) : (loss_kind == SBR_LOSS_BCE ? (double)B * n : (double)B);
  BnInline bu{nullptr, nullptr, nullptr, nullptr, nullptr}, bi = bu;
  if (bn_u && bn_u->z) bu = BnInline{bn_u->z, bn_u->mean_invstd, bn_u->gamma, bn_u->beta, bn_u->sums};
  if (bn_i && bn_i->z) bi = BnInline{bn_i->z, bn_i->mean_invstd, bn_i->gamma, bn_i->beta, bn_i->sums};
  const int nw = 8;
  const size_t sm = ((size_t)nw * 2 * n + nw + (size_t)nw * 4 * D) * sizeof(float);
  int64_t blocks = (B + nw - 1) / nw;
  const int64_t cap = (int64_t)sbr_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  const float inv = (float)(1.0 / cnt);
  const bool generic = getenv("SBR_SCORE_GENERIC") != nullptr;
#define SBR_FUSED_P(LPR_, MAXP_)                                                                                  \
  do {                                                                                                            \
    int occ = 0; /* persistent grid = exactly the resident blocks (no second wave) */                             \
    SBR_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, score_loss_bn_kernel<LPR_, MAXP_>, nw * 32, sm)); \
    int64_t grid = (B + nw - 1) / nw;                                                                             \
    if (grid > (int64_t)sbr_num_sms() * (occ > 0 ? occ : 1)) grid = (int64_t)sbr_num_sms() * (occ > 0 ? occ : 1); \
    SBR_CHECK_CUDA(sbr_launch(score_loss_bn_kernel<LPR_, MAXP_>, dim3((unsigned)grid), dim3(nw * 32), sm, S(stream), \
                              eu, bu, ei, bi, B, n, loss_kind, inv, ssm_shift, logits, loss_acc, deu, dei,        \
                              n_replicas));                                                                       \
  } while (0)
#define SBR_FUSED(LPR_)                                            \
  do {                                                             \
    const int passes = (n + 32 / LPR_ - 1) / (32 / LPR_);          \
    if (!generic && passes <= 6) SBR_FUSED_P(LPR_, 6);             \
    else if (!generic && passes <= 12) SBR_FUSED_P(LPR_, 12);      \
    else SBR_FUSED_P(LPR_, 0);                                     \
  } while (0)
  if (D == 16) SBR_FUSED(4);
  else if (D == 32) SBR_FUSED(8);
  else if (D == 64) SBR_FUSED(16);
  else SBR_FUSED(32);
#undef SBR_FUSED
#undef SBR_FUSED_P
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

// ------------------------------------------------------------------------------------------------ MF bias terms
// logits[b, j] += user_bias[u[b]] + item_bias[i[b, j]] + global_bias   (SGDMatrixFactorization, sgd_alg.py:187-195;
// any of the three may be absent) and its backward: d item_bias[i[b, j]] += dl[b, j], d user_bias[u[b]] += sum_j dl[b, j],
// d global_bias += sum dl.
namespace {
__global__ void logit_bias_fwd_kernel(float* __restrict__ logits, int64_t B, int n, const int64_t* __restrict__ u_idx,
                                      const int64_t* __restrict__ i_idx, const float* __restrict__ user_bias,
                                      const float* __restrict__ item_bias, const float* __restrict__ global_bias) {
  SBR_PDL_ENTRY();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= B * n) return;
  float v = logits[t];
  if (user_bias) v += user_bias[u_idx[t / n]];
  if (item_bias) v += item_bias[i_idx[t]];
  if (global_bias) v += global_bias[0];
  logits[t] = v;
}
__global__ void __launch_bounds__(256)
logit_bias_bwd_kernel(const float* __restrict__ dlogits, int64_t B, int n, const int64_t* __restrict__ u_idx,
                      const int64_t* __restrict__ i_idx, float* __restrict__ d_user_bias,
                      float* __restrict__ d_item_bias, float* __restrict__ d_global_bias) {
  SBR_PDL_ENTRY();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float v = 0.f;
  if (t < B * n) {
    v = dlogits[t];
    if (d_item_bias) atomicAdd(d_item_bias + i_idx[t], v);
    if (d_user_bias) atomicAdd(d_user_bias + u_idx[t / n], v);
  }
  if (d_global_bias) {
    __shared__ float red[8];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s_ = 0.f;
      for (int w = 0; w < 8; ++w) s_ += red[w];
      atomicAdd(d_global_bias, s_);
    }
  }
}
}  // namespace

extern "C" int sbr_logit_bias_fwd(float* logits, int64_t B, int n, const int64_t* u_idx, const int64_t* i_idx,
                                  const float* user_bias, const float* item_bias, const float* global_bias,
                                  void* stream) {
  SBR_REQUIRE(logits && B > 0 && n >= 1 && (!user_bias || u_idx) && (!item_bias || i_idx),
              "sbr_logit_bias_fwd: bad arguments");
  SBR_CHECK_CUDA(sbr_launch(logit_bias_fwd_kernel, dim3(cdiv(B * n, 256)), dim3(256), (size_t)0, S(stream), logits, B, n,
                            u_idx, i_idx, user_bias, item_bias, global_bias));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_logit_bias_bwd(const float* dlogits, int64_t B, int n, const int64_t* u_idx, const int64_t* i_idx,
                                  float* d_user_bias, float* d_item_bias, float* d_global_bias, void* stream) {
  SBR_REQUIRE(dlogits && B > 0 && n >= 1 && (!d_user_bias || u_idx) && (!d_item_bias || i_idx),
              "sbr_logit_bias_bwd: bad arguments");
  SBR_CHECK_CUDA(sbr_launch(logit_bias_bwd_kernel, dim3(cdiv(B * n, 256)), dim3(256), (size_t)0, S(stream), dlogits, B, n,
                            u_idx, i_idx, d_user_bias, d_item_bias, d_global_bias));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

// x[i] < lo ? lo : x[i]  (in place);  backward: dx[i] = 0 where the forward input was below lo
__global__ void clamp_min_fwd_kernel(float* __restrict__ x, int64_t n, float lo, uint8_t* __restrict__ clamped) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const bool c = v < lo;
  if (clamped) clamped[i] = c ? 1 : 0;
  if (c) x[i] = lo;
}
__global__ void clamp_min_bwd_kernel(float* __restrict__ dx, int64_t n, const uint8_t* __restrict__ clamped) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && clamped[i]) dx[i] = 0.f;
}

extern "C" int sbr_clamp_min_fwd(float* x, int64_t n, float lo, uint8_t* clamped, void* stream) {
  SBR_REQUIRE(x && n > 0, "sbr_clamp_min_fwd: bad arguments");
  clamp_min_fwd_kernel<<<cdiv(n, 256), 256, 0, S(stream)>>>(x, n, lo, clamped);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_clamp_min_bwd(float* dx, int64_t n, const uint8_t* clamped, void* stream) {
  SBR_REQUIRE(dx && clamped && n > 0, "sbr_clamp_min_bwd: bad arguments");
  clamp_min_bwd_kernel<<<cdiv(n, 256), 256, 0, S(stream)>>>(dx, n, clamped);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_aggregate(const float* e, int64_t rows, int k, int D, int agg_max, float* out_f32, void* out_bf16,
                             int64_t ld_bf16, void* stream) {
  SBR_REQUIRE(e && rows > 0 && k >= 1 && D > 0 && (out_f32 || out_bf16), "sbr_aggregate: bad arguments");
  unsigned blocks = (unsigned)min((int64_t)sbr_num_sms() * 16, (rows * D + 255) / 256);
  aggregate_kernel<<<blocks, 256, 0, S(stream)>>>(e, rows, k, D, agg_max, out_f32, reinterpret_cast<bf16*>(out_bf16),
                                                  ld_bf16);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}
