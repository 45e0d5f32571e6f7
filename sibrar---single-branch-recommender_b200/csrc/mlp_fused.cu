// sibrar_b200 -- the single-branch MLP of a 64-wide entity as ONE persistent kernel per direction.
//
// Forward (sbr_mlp2_fwd):  X0 = dropout(normalize(gather(idx, mods)))  ->  [Linear + act]  ->  Linear (+ act)  ->  z
//   replaces _get_modality_embeddings + _embed of the reference (algorithms/sgd_alg.py:1934-1978, 1865-1877) and the
//   PolyLinear forward (modules/polylinear.py:50-76) for chains of one or two Linear layers whose widths are <= 64:
//   producer warps gather the source rows straight into the SWIZZLE_128B A stage (no X0 round trip through HBM), the
//   weights stay resident in shared memory, the hidden activation goes TMEM -> registers -> bias / activation -> shared
//   memory as the next A operand, and only the final pre-BatchNorm z (fp32) + its column statistics leave the SM.
// Backward (sbr_mlp2_bwd): per 128-row tile  re-gather X0, recompute Y1, dz = BatchNorm-backward(dE, z) (or the
//   activation gradient), dY1 = (dz W1) * act'(Y1), dX0 = dY1 W0, and BOTH weight gradients + BOTH bias gradients as one
//   accumulating MMA chain  [dz ; dY1]^T (stacked on M) x [Y1 | X0] (stacked on N)  /  x ones  whose fp32 accumulators
//   stay in TMEM across all tiles of the CTA and are flushed once (vector reductions).  Nothing but dE, z (in) and dX0
//   (out) touches HBM; replaces bn_bwd_apply + 2 dgrad + 2 wgrad GEMMs + their bf16 activations.
//
// One shared-memory image serves two operand roles: a [128 rows x 64] bf16 tile stored as 128-byte lines (line r = row r,
// 16-byte chunks XOR-swizzled by r & 7, 8-line groups 1024 B apart) is BOTH the K-major operand (M = rows, K = columns)
// of a dgrad / forward MMA and the MN-major operand (MN = columns, K = rows) of the wgrad MMA; a [64 x 64] weight image
// is both the K-major B of the forward (N = out, K = in) and the MN-major B of the dgrad (N = in, K = out).
#include <stdarg.h>

#include "common.cuh"

namespace {
#include "gather_common.cuh"

constexpr int TILE_ROWS = 128;
constexpr int TILE_BYTES = TILE_ROWS * 128;  // 16 KiB
constexpr int W_BYTES = 64 * 128;            // 8 KiB
constexpr int LPR = 8;                       // lanes per gathered row (8 columns each)
constexpr int MAX_SRC = 16;
constexpr int N_PRODUCER_WARPS = 4, MMA_WARP = 4, FIRST_EPI_WARP = 5, N_THREADS = 9 * 32;

struct GatherArgs {
  const sbr_modality_src_t* srcs;
  int n_mods;
  const int64_t* idx;
  const uint8_t* mods;
  int64_t N;  // rows = n_idx * k
  int k, C, normalize;
  float p_drop;
  uint64_t seed;
  const int64_t* step_dev;
  const uint8_t* keep_mask;
  int32_t* err_flag;
};
struct LayerArgs {
  const float* bias;
  int in_f, out_f, act;
};
struct FwdParams {
  int debug;  // SBR_MLP2_DEBUG bit mask (profiling only): 2 = no gather loads, 8 = no stores
  GatherArgs g;
  LayerArgs l[2];
  int n_layers;
  float* z;
  int64_t ldz;
  float* colstats;  // [colstats_rows, 2 D] partial column sums / sums of squares (deterministic), or nullptr
  int D;
};
struct BwdParams {
  int debug;  // SBR_MLP2_DEBUG bit mask (profiling only): 1 = no gradient flush, 2 = no gather loads, 4 = no dy / z loads,
              // 8 = no dX0 stores
  GatherArgs g;
  LayerArgs l[2];
  int n_layers;
  const float* dy;
  int64_t lddy;
  const float* z;
  int64_t ldz;
  // BatchNorm behind the chain (nullptr: none): dz = gamma * invstd * (dy - sum0 / N - xhat * sum1 / N)
  const float* mean_invstd;
  const float* gamma;
  const float* sums;
  int n_replicas;
  float* dgamma;
  float* dbeta;
  float* gw[2];
  float* gb[2];
  float* dx;
  int64_t lddx;
  int D;
};

// Profiling only (SBR_MLP2_DEBUG bit 16): block 0 stamps %globaltimer at the phase boundaries of its roles into a device
// buffer read back by sbr_mlp2_trace_read (scripts/trace_mlp2.py prints the per-tile timeline).
__device__ unsigned long long g_trace[4096];
__device__ unsigned int g_trace_n;
__device__ __forceinline__ void trace_ev(int debug, int id) {
  if ((debug & 16) && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned i = atomicAdd(&g_trace_n, 1u);
    if (i < 4096) g_trace[i] = (t << 8) | (unsigned long long)(id & 0xff);
  }
}

// Barrier wait of the latency-bound role hand-offs of these kernels: `mbarrier.test_wait` polling (no suspension).
// `try_wait` may park the thread for a system-dependent time; the hand-off chain of one tile (producer -> MMA -> epilogue ->
// MMA -> ...) pays that latency 6-8 times per tile with nothing else to overlap it.
template <bool POLL>
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  if (!POLL) {
    mbar_wait(bar, parity);
    return;
  }
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {  // byte offset of 16-byte chunk `chunk` of line `row`
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[2 * t], v[2 * t + 1]);
  return u;
}

// butterfly transpose-reduce (see gemm_sm100.cu): on return lane l holds the sum over the warp's lanes of v[l]
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 16; off >= 1; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      float send = upper ? v[i] : v[i + n];
      float keep = upper ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// activation / activation gradient with the (warp-uniform) switch hoisted out of the element loop
template <int NE>
__device__ __forceinline__ void act_n(int act, float (&v)[NE]) {
  switch (act) {
    case SBR_ACT_RELU:
#pragma unroll
      for (int j = 0; j < NE; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    case SBR_ACT_TANH:
#pragma unroll
      for (int j = 0; j < NE; ++j) v[j] = tanhf(v[j]);
      break;
    case SBR_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < NE; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
      break;
    case SBR_ACT_SELU:
#pragma unroll
      for (int j = 0; j < NE; ++j)
        v[j] = 1.0507009873554805f * (v[j] > 0.f ? v[j] : 1.6732632423543772f * (__expf(v[j]) - 1.f));
      break;
    default: break;
  }
}
// v[j] *= act'(y[j]) expressed through the output y
template <int NE>
__device__ __forceinline__ void actgrad_n(int act, float (&v)[NE], const float (&y)[NE]) {
  switch (act) {
    case SBR_ACT_RELU:
#pragma unroll
      for (int j = 0; j < NE; ++j) v[j] = y[j] > 0.f ? v[j] : 0.f;
      break;
    case SBR_ACT_TANH:
#pragma unroll
      for (int j = 0; j < NE; ++j) v[j] *= 1.f - y[j] * y[j];
      break;
    case SBR_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < NE; ++j) v[j] *= y[j] * (1.f - y[j]);
      break;
    case SBR_ACT_SELU:
#pragma unroll
      for (int j = 0; j < NE; ++j)
        v[j] *= y[j] > 0.f ? 1.0507009873554805f : y[j] + 1.0507009873554805f * 1.6732632423543772f;
      break;
    default: break;
  }
}
__device__ __forceinline__ void act32(int act, float (&v)[32]) { act_n<32>(act, v); }

// bias of one 32-column chunk from the shared-memory copy (zero beyond the layer's width): broadcast reads
__device__ __forceinline__ void add_bias32(float (&v)[32], const uint32_t (&r)[32], const float* s_bias) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(s_bias + j);
    v[j] = __uint_as_float(r[j]) + b.x;
    v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
    v[j + 2] = __uint_as_float(r[j + 2]) + b.z;
    v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
  }
}

// Producer, phase A (thread = row of the tile): entity index -> modality source -> feature row -> address of the fp32
// source row (nullptr: no row).  128 independent dependent-load chains in flight per CTA.
__device__ __forceinline__ void resolve_rows(const GatherArgs& g, const sbr_modality_src_t* s_src, int64_t tile,
                                             const float** s_ptr, int tid) {
  const int64_t gr = tile * TILE_ROWS + tid;
  const float* ptr = nullptr;
  if (gr < g.N) {
    int m = g.mods ? (int)__ldg(g.mods + gr) : 0;
    m = min(m, g.n_mods - 1);
    const sbr_modality_src_t& s = s_src[m];
    const int64_t e = __ldg(g.idx + gr / g.k);
    const int64_t feat = s.remap ? (int64_t)__ldg(s.remap + e) : e;
    if (feat < 0) {
      if (g.err_flag) atomicExch(g.err_flag, 1);
    } else {
      const int64_t src_row = (s.kind == SBR_SRC_CATEGORICAL) ? (int64_t)__ldg(s.codes + feat) : feat;
      ptr = s.table + src_row * g.C;
    }
  }
  s_ptr[tid] = ptr;
}

// Producer, phase B: groups of 8 lanes load the rows (4 rows per group in flight), normalise, drop, convert and store
// them into `dst` (the SWIZZLE_128B image).  TAG sources are not handled here (entities present them as tables).
__device__ __forceinline__ void gather_tile(const GatherArgs& g, const float* const* s_ptr, int64_t tile, uint8_t* dst,
                                            int tid, uint64_t step, bool no_loads) {
  const int grp = tid >> 3, li = tid & 7;
  const float sc = g.p_drop > 0.f ? 1.f / (1.f - g.p_drop) : 1.f;
  const bool vec_ok = (g.C & 3) == 0;
#pragma unroll 1
  for (int pass = 0; pass < 8; pass += 4) {
    float x[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* ptr = s_ptr[(pass + u) * 16 + grp];
      if (ptr != nullptr && !no_loads) {
        load8(ptr, 8 * li, g.C, vec_ok, x[u]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[u][j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = (pass + u) * 16 + grp;
      const int64_t gr = tile * TILE_ROWS + row;
      if (g.normalize) {
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += x[u][j] * x[u][j];
        ss = group_sum<LPR>(ss);
        const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[u][j] *= inv;
      }
      const uint32_t km = gr < g.N ? keep8(g.keep_mask, gr, g.C, 8 * li, g.p_drop, g.seed, step) : 0u;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ((km >> j) & 1u) ? x[u][j] * sc : 0.f;
      *reinterpret_cast<uint4*>(dst + tile_off(row, li)) = pack8(v);
    }
  }
}

__device__ __forceinline__ void producer_sync() {  // the 128 producer threads only
  asm volatile("bar.sync 1, 128;" ::: "memory");
}

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}

// K-major operand (rows = M or N, 64 K-elements per line): 4 K steps of 16 elements = +32 B inside the swizzle span
__device__ __forceinline__ uint64_t desc_k(uint32_t addr, int k16) { return umma_smem_desc(addr + k16 * 32, 16, 1024); }
// MN-major operand (lines = K, 64 MN-elements per line, further 64-element blocks `lbo` bytes apart): a K step of 16
// lines = +2048 B
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, int k16, uint32_t lbo) {
  return umma_smem_desc(addr + k16 * 2048, lbo, 1024);
}

// ================================================================================================ forward
constexpr int FWD_SMEM = 2 * W_BYTES + 3 * TILE_BYTES + 3072 + 1024;

template <int L, bool POLL>
__global__ void __launch_bounds__(N_THREADS, 2)
mlp2_fwd_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1, FwdParams p) {
  SBR_PDL_LAUNCH();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW0 = smem;
  uint8_t* sW1 = smem + W_BYTES;
  uint8_t* sX = smem + 2 * W_BYTES;        // 2 stages
  uint8_t* sA1 = sX + 2 * TILE_BYTES;      // hidden activation (L == 2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA1 + TILE_BYTES);
  uint64_t* w_full = bars;          // weights landed
  uint64_t* x_full = bars + 1;      // [2]
  uint64_t* x_empty = bars + 3;     // [2]
  uint64_t* dh_full = bars + 5;     // hidden accumulator ready
  uint64_t* a1_full = bars + 6;     // hidden activation written
  uint64_t* df_full = bars + 7;     // final accumulator ready
  uint64_t* df_empty = bars + 8;    // final accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  sbr_modality_src_t* s_src = reinterpret_cast<sbr_modality_src_t*>(bars + 12);
  const float** s_ptr = reinterpret_cast<const float**>(s_src + MAX_SRC);  // [128] source row of every tile row
  float* s_bias = reinterpret_cast<float*>(s_ptr + TILE_ROWS);                 // [2][64], zero beyond the widths

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.g.N + TILE_ROWS - 1) / TILE_ROWS;
  if (warp == MMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmW0);
    if (L == 2) tma_prefetch_desc(&tmW1);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], N_PRODUCER_WARPS);
      mbar_init(&x_empty[s], 1);
    }
    mbar_init(dh_full, 1);
    mbar_init(a1_full, 4);
    mbar_init(df_full, 1);
    mbar_init(df_empty, 4);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_h = tmem_base, tmem_f = tmem_base + 64;
  SBR_PDL_WAIT();  // nothing above touches global memory
  if (threadIdx.x < p.g.n_mods) s_src[threadIdx.x] = p.g.srcs[threadIdx.x];
  if (threadIdx.x >= 128 && threadIdx.x < 256) {
    const int l = (threadIdx.x - 128) >> 6, c = threadIdx.x & 63;
    s_bias[l * 64 + c] = (l < L && p.l[l].bias != nullptr && c < p.l[l].out_f) ? p.l[l].bias[c] : 0.f;
  }
  __syncthreads();

  if (warp < N_PRODUCER_WARPS) {
    // ---------------------------------------------------------------- producers: gather + normalise + dropout -> X0
    const uint64_t step = p.g.step_dev ? (uint64_t)*p.g.step_dev : 0;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      resolve_rows(p.g, s_src, tile, s_ptr, threadIdx.x);
      producer_sync();
      wait_bar<POLL>(&x_empty[s], (uint32_t)(((it >> 1) & 1) ^ 1));
      gather_tile(p.g, s_ptr, tile, sX + s * TILE_BYTES, threadIdx.x, step, (p.debug & 2) != 0);
      fence_proxy_async_smem();
      producer_sync();  // (also: s_ptr may be overwritten by the next tile)
      if (lane == 0) mbar_arrive(&x_full[s]);
    }
  } else if (warp == MMA_WARP) {
    // ---------------------------------------------------------------- MMA issuer (converged warp, elected issue)
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, (L == 2 ? 2 : 1) * W_BYTES);
      tma_load_2d(sW0, &tmW0, w_full, 0, 0);
      if (L == 2) tma_load_2d(sW1, &tmW1, w_full, 0, 0);
    }
    __syncwarp();
    wait_bar<POLL>(w_full, 0);
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t aX = smem_u32(sX), aA1 = smem_u32(sA1), aW0 = smem_u32(sW0), aW1 = smem_u32(sW1);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      wait_bar<POLL>(&x_full[s], (uint32_t)((it >> 1) & 1));
      if (L == 1) wait_bar<POLL>(df_empty, (uint32_t)((it & 1) ^ 1));
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(L == 2 ? tmem_h : tmem_f, desc_k(aX + s * TILE_BYTES, k), desc_k(aW0, k), idesc, k > 0 ? 1u : 0u);
        umma_commit(&x_empty[s]);
        umma_commit(L == 2 ? dh_full : df_full);
      }
      __syncwarp();
      if (L == 2) {
        wait_bar<POLL>(a1_full, (uint32_t)(it & 1));
        wait_bar<POLL>(df_empty, (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_f, desc_k(aA1, k), desc_k(aW1, k), idesc, k > 0 ? 1u : 0u);
          umma_commit(df_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = row of the tile
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int act0 = p.l[0].act, act_last = p.l[L - 1].act;
    float cs_acc[2] = {0.f, 0.f}, cq_acc[2] = {0.f, 0.f};
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int64_t row = tile * TILE_ROWS + row_in_tile;
      const bool row_ok = row < p.g.N;
      if (L == 2) {
        wait_bar<POLL>(dh_full, (uint32_t)(it & 1));
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_h + lane_off + c0, r);
          tmem_ld_wait();
          float v[32];
          add_bias32(v, r, s_bias + c0);
          act32(act0, v);
          // (columns beyond the layer's width hold act(0): the next weight's K columns there are zero)
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float w8[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w8[t] = v[j + t];
            *reinterpret_cast<uint4*>(sA1 + tile_off(row_in_tile, (c0 + j) >> 3)) = pack8(w8);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a1_full);
      }
      wait_bar<POLL>(df_full, (uint32_t)(it & 1));
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        if (c0 >= p.D) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(tmem_f + lane_off + c0, r);
        tmem_ld_wait();
        float v[32];
        add_bias32(v, r, s_bias + (L - 1) * 64 + c0);
        act32(act_last, v);
        if (p.colstats != nullptr) {
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = row_ok ? v[j] : 0.f;
            s1[j] = x;
            s2[j] = x * x;
          }
          cs_acc[c0 >> 5] += warp_colsum32(s1, lane);
          cq_acc[c0 >> 5] += warp_colsum32(s2, lane);
        }
        if (row_ok && !(p.debug & 8)) {
          float* dst = p.z + row * p.ldz + c0;
          if (c0 + 32 <= p.D && (p.ldz & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < p.D) dst[j] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(df_empty);
    }
    if (p.colstats != nullptr) {
      // one row of partial sums per CTA: the four lane quarters are added in a fixed order (deterministic statistics;
      // sbr_bn_finalize adds the rows of all CTAs in a fixed order as well)
      float* s_part = reinterpret_cast<float*>(sA1);  // (the hidden-activation tile is no longer needed)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        s_part[q * 128 + 32 * i + lane] = cs_acc[i];
        s_part[q * 128 + 64 + 32 * i + lane] = cq_acc[i];
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");  // the four epilogue warps
      if (q == 0) {
        float* rowp = p.colstats + (size_t)blockIdx.x * 2 * p.D;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int col = 32 * i + lane;
          if (col < p.D) {
            rowp[col] = ((s_part[col] + s_part[128 + col]) + s_part[256 + col]) + s_part[384 + col];
            rowp[p.D + col] = ((s_part[64 + col] + s_part[128 + 64 + col]) + s_part[256 + 64 + col]) + s_part[384 + 64 + col];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ================================================================================================ backward
// shared memory: W0 | W1 | DZ | DY1 | Y1 | X0[0] | X0[1].  Stacked operands of the wgrad MMA: A = [DZ ; DY1] (second
// 64-element block one tile further), B = [Y1 | X0[s]] (second block one or two tiles further: the descriptor's leading
// byte offset selects the X0 stage), L == 1: B = X0[s] alone.
constexpr int BWD_SMEM = 2 * W_BYTES + 5 * TILE_BYTES + 3072 + 1024;

template <int L, bool POLL>
__global__ void __launch_bounds__(N_THREADS, 2)
mlp2_bwd_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1, BwdParams p) {
  SBR_PDL_LAUNCH();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW0 = smem;
  uint8_t* sW1 = smem + W_BYTES;
  uint8_t* sDZ = smem + 2 * W_BYTES;
  uint8_t* sDY1 = sDZ + TILE_BYTES;
  uint8_t* sY1 = sDY1 + TILE_BYTES;
  uint8_t* sX = sY1 + TILE_BYTES;  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * TILE_BYTES);
  uint64_t* w_full = bars;        // weights landed
  uint64_t* x_full = bars + 1;    // [2] X0 gathered
  uint64_t* x_empty = bars + 3;   // [2] every MMA that reads the stage has completed
  uint64_t* dz_full = bars + 5;   // dz written
  uint64_t* da_full = bars + 6;   // the 64-column accumulator holds a result (used 1 or 3 times per tile)
  uint64_t* y1_full = bars + 7;   // Y1 written (and the accumulator read)
  uint64_t* dy1_full = bars + 8;  // dY1 written (and the accumulator read)
  uint64_t* da_free = bars + 9;   // dX0 read out of the accumulator
  uint64_t* w_done = bars + 10;   // the wgrad MMAs of the tile have read DZ / DY1 / Y1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  sbr_modality_src_t* s_src = reinterpret_cast<sbr_modality_src_t*>(bars + 12);
  const float** s_ptr = reinterpret_cast<const float**>(s_src + MAX_SRC);
  float* s_bias = reinterpret_cast<float*>(s_ptr + TILE_ROWS);  // [2][64], zero beyond the widths

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.g.N + TILE_ROWS - 1) / TILE_ROWS;
  if (warp == MMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmW0);
    if (L == 2) tma_prefetch_desc(&tmW1);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
    }
    mbar_init(dz_full, 1);
    mbar_init(da_full, 1);
    mbar_init(y1_full, 4);
    mbar_init(dy1_full, 4);
    mbar_init(da_free, 4);
    mbar_init(w_done, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, 256);
  if (L == 1)  // a defined second block of the stacked A operand (its accumulator lanes are never read)
    for (int i = threadIdx.x; i < TILE_BYTES / 16; i += N_THREADS)
      reinterpret_cast<uint4*>(sDY1)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base, tmem_w = tmem_base + 64;
  SBR_PDL_WAIT();
  if (threadIdx.x < p.g.n_mods) s_src[threadIdx.x] = p.g.srcs[threadIdx.x];
  if (threadIdx.x >= 128 && threadIdx.x < 256) {
    const int l = (threadIdx.x - 128) >> 6, c = threadIdx.x & 63;
    s_bias[l * 64 + c] = (l < L && p.l[l].bias != nullptr && c < p.l[l].out_f) ? p.l[l].bias[c] : 0.f;
  }
  __syncthreads();
  const int my_tiles = blockIdx.x < num_tiles ? (int)((num_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;

  if (warp < N_PRODUCER_WARPS) {
    // ---------------------------------------------------------------- producers: X0 (re-gathered, one tile ahead) and dz
    const uint64_t step = p.g.step_dev ? (uint64_t)*p.g.step_dev : 0;
    const int grp = threadIdx.x >> 3, li = threadIdx.x & 7;
    const int c0 = 8 * li;
    const int act_last = p.l[L - 1].act;
    // per-column coefficients of this lane's 8 columns: dz = A dy + B (z - mean) + C0  (BatchNorm backward), A = 1 else
    float cA[8], cB[8], cM[8], c0v[8], bsum[8];
    const bool bn = p.mean_invstd != nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      cA[j] = 1.f; cB[j] = 0.f; cM[j] = 0.f; c0v[j] = 0.f; bsum[j] = 0.f;
      if (bn && c < p.D) {
        float s0 = 0.f, s1 = 0.f;
        for (int r = 0; r < p.n_replicas; ++r) {
          s0 += p.sums[(size_t)r * 2 * p.D + c];
          s1 += p.sums[(size_t)r * 2 * p.D + p.D + c];
        }
        const float inv_n = 1.f / (float)p.g.N;
        const float istd = p.mean_invstd[p.D + c];
        const float gi = p.gamma[c] * istd;
        cM[j] = p.mean_invstd[c];
        cA[j] = gi;
        cB[j] = -gi * istd * (s1 * inv_n);
        c0v[j] = -gi * (s0 * inv_n);
        if (blockIdx.x == 0 && grp == 0) {  // d gamma / d beta come with the sums
          if (p.dbeta) p.dbeta[c] += s0;
          if (p.dgamma) p.dgamma[c] += s1;
        }
      }
    }
    const bool vec_ok = (p.lddy & 3) == 0 && (p.ldz & 3) == 0 && (p.D & 7) == 0;
    const bool no_dz_loads = (p.debug & 4) != 0;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      if (warp == 0) trace_ev(p.debug, 1);   // tile start
      resolve_rows(p.g, s_src, tile, s_ptr, threadIdx.x);
      producer_sync();
      if (warp == 0) trace_ev(p.debug, 2);   // rows resolved
      wait_bar<POLL>(&x_empty[s], (uint32_t)(((it >> 1) & 1) ^ 1));
      if (warp == 0) trace_ev(p.debug, 3);   // stage free
      gather_tile(p.g, s_ptr, tile, sX + s * TILE_BYTES, threadIdx.x, step, (p.debug & 2) != 0);
      fence_proxy_async_smem();
      producer_sync();
      if (threadIdx.x == 0) mbar_arrive(&x_full[s]);
      if (warp == 0) trace_ev(p.debug, 4);   // X0 published
      // dz of the tile (16 rows per pass, 2 passes in flight): the global loads are issued before the wait for the
      // previous tile's wgrad MMAs (the last readers of DZ)
#pragma unroll 1
      for (int pass = 0; pass < 8; pass += 2) {
        float gy[2][8], zz[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int row = (pass + u) * 16 + grp;
          const int64_t gr = tile * TILE_ROWS + row;
          if (gr < p.g.N && c0 < p.D && !no_dz_loads) {
            const float* dyp = p.dy + gr * p.lddy + c0;
            const float* zp = p.z + gr * p.ldz + c0;
            if (vec_ok) {
              const float4 a = __ldcs(reinterpret_cast<const float4*>(dyp)), b = __ldcs(reinterpret_cast<const float4*>(dyp + 4));
              const float4 c = __ldcs(reinterpret_cast<const float4*>(zp)), d = __ldcs(reinterpret_cast<const float4*>(zp + 4));
              gy[u][0] = a.x; gy[u][1] = a.y; gy[u][2] = a.z; gy[u][3] = a.w;
              gy[u][4] = b.x; gy[u][5] = b.y; gy[u][6] = b.z; gy[u][7] = b.w;
              zz[u][0] = c.x; zz[u][1] = c.y; zz[u][2] = c.z; zz[u][3] = c.w;
              zz[u][4] = d.x; zz[u][5] = d.y; zz[u][6] = d.z; zz[u][7] = d.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                gy[u][j] = c0 + j < p.D ? dyp[j] : 0.f;
                zz[u][j] = c0 + j < p.D ? zp[j] : 0.f;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) gy[u][j] = zz[u][j] = 0.f;
          }
        }
        if (pass == 0) {
          if (warp == 0) trace_ev(p.debug, 5);  // first dz loads issued
          wait_bar<POLL>(w_done, (uint32_t)((it & 1) ^ 1));
          if (warp == 0) trace_ev(p.debug, 6);  // DZ buffer free
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int row = (pass + u) * 16 + grp;
          const bool ok = tile * TILE_ROWS + row < p.g.N;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = cA[j] * gy[u][j] + cB[j] * (zz[u][j] - cM[j]) + c0v[j];  // (A = 1, B = C0 = 0 without BN)
          actgrad_n<8>(act_last, v, zz[u]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[j] = (ok && c0 + j < p.D) ? v[j] : 0.f;
            bsum[j] += v[j];  // bias gradient of the last layer: fp32 column sums BEFORE the bf16 rounding
          }
          *reinterpret_cast<uint4*>(sDZ + tile_off(row, li)) = pack8(v);
        }
      }
      fence_proxy_async_smem();
      producer_sync();
      if (threadIdx.x == 0) mbar_arrive(dz_full);
      if (warp == 0) trace_ev(p.debug, 7);   // dz published
    }
    if (p.gb[L - 1] != nullptr && my_tiles > 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = bsum[j];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane < 8 && c0 + j < p.D) atomicAdd(p.gb[L - 1] + c0 + j, v);
      }
    }
  } else if (warp == MMA_WARP) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, (L == 2 ? 2 : 1) * W_BYTES);
      tma_load_2d(sW0, &tmW0, w_full, 0, 0);
      if (L == 2) tma_load_2d(sW1, &tmW1, w_full, 0, 0);
    }
    __syncwarp();
    wait_bar<POLL>(w_full, 0);
    const uint32_t id_fwd = umma_idesc_bf16(128, 64, 0, 0);    // A K-major, B K-major   (forward)
    const uint32_t id_dg = umma_idesc_bf16(128, 64, 0, 1);     // A K-major, B MN-major  (dgrad: B = weight^T view)
    const uint32_t id_wg = umma_idesc_bf16(128, L == 2 ? 128 : 64, 1, 1);  // both MN-major (contraction over the rows)
    const uint32_t aX = smem_u32(sX), aY1 = smem_u32(sY1), aDZ = smem_u32(sDZ), aDY1 = smem_u32(sDY1),
                   aW0 = smem_u32(sW0), aW1 = smem_u32(sW1);
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int s = it & 1;
      const uint32_t aXs = aX + s * TILE_BYTES;
      trace_ev(p.debug, 16);
      wait_bar<POLL>(da_free, ph ^ 1);  // dX0 of the previous tile has been read out
      trace_ev(p.debug, 17);
      wait_bar<POLL>(&x_full[s], (uint32_t)((it >> 1) & 1));
      trace_ev(p.debug, 18);
      if (L == 2) {
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_a, desc_k(aXs, k), desc_k(aW0, k), id_fwd, k > 0 ? 1u : 0u);
          umma_commit(da_full);
        }
        __syncwarp();
        wait_bar<POLL>(y1_full, ph);   // Y1 in shared memory, accumulator free again
        trace_ev(p.debug, 19);
        wait_bar<POLL>(dz_full, ph);
        trace_ev(p.debug, 20);
        tc_fence_after();
        if (elect_one()) {        // dz W1  (contraction over out_1)
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_a, desc_k(aDZ, k), desc_mn(aW1, k, W_BYTES), id_dg, k > 0 ? 1u : 0u);
          umma_commit(da_full);
        }
        __syncwarp();
        wait_bar<POLL>(dy1_full, ph);
        trace_ev(p.debug, 21);
        tc_fence_after();
        if (elect_one()) {        // dX0 = dY1 W0
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_a, desc_k(aDY1, k), desc_mn(aW0, k, W_BYTES), id_dg, k > 0 ? 1u : 0u);
          umma_commit(da_full);
        }
        __syncwarp();
      } else {
        wait_bar<POLL>(dz_full, ph);
        tc_fence_after();
        if (elect_one()) {        // dX0 = dz W0
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_a, desc_k(aDZ, k), desc_mn(aW0, k, W_BYTES), id_dg, k > 0 ? 1u : 0u);
          umma_commit(da_full);
        }
        __syncwarp();
      }
      // weight gradients: [dz ; dY1]^T (M = 128) x [Y1 | X0] (N = 128) or x X0 (N = 64), K = the 128 rows of the tile,
      // accumulated in TMEM over every tile of this CTA
      if (elect_one()) {
        const uint32_t b_addr = (L == 2) ? aY1 : aXs;
        const uint32_t b_lbo = (L == 2) ? (uint32_t)((1 + s) * TILE_BYTES) : (uint32_t)TILE_BYTES;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16(tmem_w, desc_mn(aDZ, k, TILE_BYTES), desc_mn(b_addr, k, b_lbo), id_wg, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&x_empty[s]);
        umma_commit(w_done);
      }
      __syncwarp();
      trace_ev(p.debug, 22);
    }
  } else {
    // ---------------------------------------------------------------- epilogue: thread = row of the tile
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t da_ph = 0;  // parity of the next completion of da_full
    const int act0 = p.l[0].act;
    float cb_acc[2] = {0.f, 0.f};  // bias gradient of layer 0 (L == 2): lane l owns column 32 i + l
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int64_t row = tile * TILE_ROWS + row_in_tile;
      const bool row_ok = row < p.g.N;
      if (L == 2) {
        if (warp == 5) trace_ev(p.debug, 32);
        wait_bar<POLL>(da_full, da_ph);
        if (warp == 5) trace_ev(p.debug, 33);
        da_ph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_a + lane_off + c0, r);
          tmem_ld_wait();
          float v[32];
          add_bias32(v, r, s_bias + c0);
          act32(act0, v);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float w8[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w8[t] = v[j + t];
            const uint4 u = pack8(w8);
            *reinterpret_cast<uint4*>(sY1 + tile_off(row_in_tile, (c0 + j) >> 3)) = u;
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(y1_full);
        if (warp == 5) trace_ev(p.debug, 34);
        // dY1 = (dz W1) * act'(Y1)
        wait_bar<POLL>(da_full, da_ph);
        if (warp == 5) trace_ev(p.debug, 35);
        da_ph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_a + lane_off + c0, r);
          tmem_ld_wait();
          // (rows beyond N have dz = 0, columns beyond the width meet zero weight columns: no masks needed)
          float v[32], yv[32];
#pragma unroll
          for (int j = 0; j < 32; j += 8) {  // Y1 of this row back from its shared-memory tile (bf16)
            const uint4 u = *reinterpret_cast<const uint4*>(sY1 + tile_off(row_in_tile, (c0 + j) >> 3));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 y = __bfloat1622float2(h[t]);
              yv[j + 2 * t] = y.x;
              yv[j + 2 * t + 1] = y.y;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          actgrad_n<32>(act0, v, yv);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float w8[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w8[t] = v[j + t];
            *reinterpret_cast<uint4*>(sDY1 + tile_off(row_in_tile, (c0 + j) >> 3)) = pack8(w8);
          }
          if (p.gb[0] != nullptr && c0 < p.l[0].out_f) cb_acc[c0 >> 5] += warp_colsum32(v, lane);  // fp32, pre-rounding
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(dy1_full);
        if (warp == 5) trace_ev(p.debug, 36);
      }
      // dX0 -> global (fp32, consumed by the sorted-run gather backward)
      wait_bar<POLL>(da_full, da_ph);
      if (warp == 5) trace_ev(p.debug, 37);
      da_ph ^= 1;
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        if (c0 >= p.g.C) break;
        uint32_t r[32];
        tmem_ld32(tmem_a + lane_off + c0, r);
        tmem_ld_wait();
        if (row_ok && !(p.debug & 8)) {
          float* dst = p.dx + row * p.lddx + c0;
          if (c0 + 32 <= p.g.C && (p.lddx & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                 __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < p.g.C) dst[j] = __uint_as_float(r[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(da_free);
      if (warp == 5) trace_ev(p.debug, 38);
    }
    // ---- flush the gradient accumulators
    if (my_tiles > 0) {
      if (L == 2 && p.gb[0] != nullptr) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int col = 32 * i + lane;
          if (col < p.l[0].out_f) atomicAdd(p.gb[0] + col, cb_acc[i]);
        }
      }
      wait_bar<POLL>(w_done, (uint32_t)((my_tiles - 1) & 1));
      tc_fence_after();
      // lane = output feature of the stacked A operand.  L == 2: lanes 0-63 = dz (layer 1, x Y1 = columns 0-63), lanes
      // 64-127 = dY1 (layer 0, x X0 = columns 64-127).  L == 1: lanes 0-63 = dz (layer 0, x X0 = columns 0-63).
      const int half = q >> 1;
      const int layer = (L == 2) ? 1 - half : 0;
      const bool used = ((L == 2) || half == 0) && !(p.debug & 1);
      const int o = row_in_tile & 63;
      if (used) {
        const LayerArgs la = p.l[layer];
        float* gw = p.gw[layer];
        const int col_base = (L == 2 && half == 1) ? 64 : 0;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          if (c0 >= la.in_f) break;
          uint32_t r[32];
          tmem_ld32(tmem_w + lane_off + col_base + c0, r);
          tmem_ld_wait();
          if (o < la.out_f && gw != nullptr) {
            float* dst = gw + (int64_t)o * la.in_f + c0;
            if (c0 + 32 <= la.in_f && (la.in_f & 3) == 0 && (reinterpret_cast<uintptr_t>(gw) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const size_t a = __cvta_generic_to_global(dst + j);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(__uint_as_float(r[j])),
                             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                             "f"(__uint_as_float(r[j + 3]))
                             : "memory");
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c0 + j < la.in_f) atomicAdd(dst + j, __uint_as_float(r[j]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int mlp2_debug() {
  const char* e = getenv("SBR_MLP2_DEBUG");
  return e ? atoi(e) : 0;
}

inline bool mlp2_poll() {
  const char* e = getenv("SBR_MLP2_POLL");
  return e ? atoi(e) != 0 : false;  // (measured: test_wait polling is 4 % slower than try_wait here)
}

inline int64_t mlp2_grid(int64_t N) {
  const int64_t tiles = (N + TILE_ROWS - 1) / TILE_ROWS;
  const int64_t g = 2 * (int64_t)sbr_num_sms();
  return tiles < g ? (tiles < 1 ? 1 : tiles) : g;
}

int check_desc(const sbr_mlp2_desc_t* d, const char* who) {
  SBR_REQUIRE(d && d->srcs && d->idx && d->n_idx > 0 && d->k >= 1, "%s: bad gather arguments", who);
  SBR_REQUIRE(d->n_mods >= 1 && d->n_mods <= MAX_SRC, "%s: n_mods=%d not in [1, %d]", who, d->n_mods, MAX_SRC);
  SBR_REQUIRE(d->n_layers == 1 || d->n_layers == 2, "%s: n_layers=%d (1 or 2 supported)", who, d->n_layers);
  SBR_REQUIRE(d->C >= 1 && d->C <= 64, "%s: C=%d not in [1, 64]", who, d->C);
  int in = d->C;
  for (int l = 0; l < d->n_layers; ++l) {
    const sbr_mlp2_layer_t& y = d->layers[l];
    SBR_REQUIRE(y.w_bf16 && y.in_f == in && y.out_f >= 1 && y.out_f <= 64 && y.ldw >= y.in_f && y.ldw % 8 == 0,
                "%s: layer %d: in=%d (expected %d) out=%d ldw=%lld", who, l, y.in_f, in, y.out_f, (long long)y.ldw);
    in = y.out_f;
  }
  return SBR_OK;
}

void fill_gather(GatherArgs& g, const sbr_mlp2_desc_t* d) {
  g.srcs = d->srcs; g.n_mods = d->n_mods; g.idx = d->idx; g.mods = d->mods; g.N = d->n_idx * d->k; g.k = d->k;
  g.C = d->C; g.normalize = d->normalize; g.p_drop = d->p_drop; g.seed = d->seed; g.step_dev = d->step_dev;
  g.keep_mask = d->keep_mask; g.err_flag = d->err_flag;
}

int make_weight_maps(const sbr_mlp2_desc_t* d, CUtensorMap* tm) {
  for (int l = 0; l < 2; ++l) {
    const sbr_mlp2_layer_t& y = d->layers[l < d->n_layers ? l : 0];
    int rc = sbr_make_tmap_bf16_2d(&tm[l], y.w_bf16, (uint64_t)y.in_f, (uint64_t)y.out_f, (uint64_t)y.ldw, 64, 64);
    if (rc) return rc;
  }
  return SBR_OK;
}

}  // namespace

extern "C" int sbr_mlp2_trace_read(unsigned long long* host_out, int max_events) {
  unsigned int n = 0;
  SBR_CHECK_CUDA(cudaDeviceSynchronize());
  SBR_CHECK_CUDA(cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n)));
  if (n > 4096u) n = 4096u;
  if ((int)n > max_events) n = (unsigned)max_events;
  if (n > 0) SBR_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_trace, sizeof(unsigned long long) * n));
  const unsigned int zero = 0;
  SBR_CHECK_CUDA(cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero)));
  return (int)n;  // (number of events, not a status)
}

extern "C" int sbr_mlp2_colstats_rows(int64_t n_rows) { return (int)mlp2_grid(n_rows); }

extern "C" int sbr_mlp2_fwd(const sbr_mlp2_desc_t* d, int64_t n_rows, int C, float* z, int64_t ldz, float* colstats,
                            int colstats_rows, void* stream) {
  int rc = check_desc(d, "sbr_mlp2_fwd");
  if (rc) return rc;
  SBR_REQUIRE(z && n_rows == d->n_idx * d->k && C == d->C, "sbr_mlp2_fwd: bad output arguments");
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.debug = mlp2_debug();
  fill_gather(p.g, d);
  for (int l = 0; l < d->n_layers; ++l) p.l[l] = LayerArgs{d->layers[l].bias, d->layers[l].in_f, d->layers[l].out_f, d->layers[l].act};
  p.n_layers = d->n_layers;
  p.D = d->layers[d->n_layers - 1].out_f;
  p.z = z; p.ldz = ldz; p.colstats = colstats;
  SBR_REQUIRE(ldz >= p.D, "sbr_mlp2_fwd: ldz < D");
  const int64_t grid = mlp2_grid(n_rows);
  SBR_REQUIRE(colstats == nullptr || colstats_rows >= grid, "sbr_mlp2_fwd: colstats_rows=%d < %lld", colstats_rows,
              (long long)grid);
  CUtensorMap tm[2];
  rc = make_weight_maps(d, tm);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    SBR_CHECK_CUDA(cudaFuncSetAttribute(mlp2_fwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    SBR_CHECK_CUDA(cudaFuncSetAttribute(mlp2_fwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    SBR_CHECK_CUDA(cudaFuncSetAttribute(mlp2_fwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    SBR_CHECK_CUDA(cudaFuncSetAttribute(mlp2_fwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    configured = true;
  }
  if (colstats != nullptr && colstats_rows > grid)  // rows no CTA writes must not hold garbage
    SBR_CHECK_CUDA(cudaMemsetAsync(colstats + (size_t)grid * 2 * p.D, 0,
                                   (size_t)(colstats_rows - grid) * 2 * p.D * sizeof(float), S(stream)));
  const bool poll = mlp2_poll();
  auto kern = d->n_layers == 1 ? (poll ? mlp2_fwd_kernel<1, true> : mlp2_fwd_kernel<1, false>)
                               : (poll ? mlp2_fwd_kernel<2, true> : mlp2_fwd_kernel<2, false>);
  SBR_CHECK_CUDA(sbr_launch(kern, dim3((unsigned)grid), dim3(N_THREADS), (size_t)FWD_SMEM, S(stream), tm[0], tm[1], p));
  return SBR_OK;
}

extern "C" int sbr_mlp2_bwd(const sbr_mlp2_desc_t* d, int64_t n_rows, int C, const float* dy, int64_t lddy,
                            const float* z, int64_t ldz, const sbr_mlp2_bn_t* bn, float* const* grad_w,
                            float* const* grad_b, float* dx, int64_t lddx, void* stream) {
  int rc = check_desc(d, "sbr_mlp2_bwd");
  if (rc) return rc;
  SBR_REQUIRE(dy && z && dx && grad_w && grad_b && n_rows == d->n_idx * d->k && C == d->C,
              "sbr_mlp2_bwd: bad arguments");
  BwdParams p;
  memset(&p, 0, sizeof(p));
  p.debug = mlp2_debug();
  fill_gather(p.g, d);
  for (int l = 0; l < d->n_layers; ++l) {
    p.l[l] = LayerArgs{d->layers[l].bias, d->layers[l].in_f, d->layers[l].out_f, d->layers[l].act};
    p.gw[l] = grad_w[l];
    p.gb[l] = grad_b[l];
  }
  p.n_layers = d->n_layers;
  p.D = d->layers[d->n_layers - 1].out_f;
  p.dy = dy; p.lddy = lddy; p.z = z; p.ldz = ldz; p.dx = dx; p.lddx = lddx;
  if (bn != nullptr) {
    SBR_REQUIRE(bn->mean_invstd && bn->gamma && bn->sums && bn->n_replicas >= 1, "sbr_mlp2_bwd: bad BatchNorm arguments");
    p.mean_invstd = bn->mean_invstd; p.gamma = bn->gamma; p.sums = bn->sums; p.n_replicas = bn->n_replicas;
    p.dgamma = bn->dgamma; p.dbeta = bn->dbeta;
  }
  CUtensorMap tm[2];
  rc = make_weight_maps(d, tm);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    SBR_CHECK_CUDA(cudaFuncSetAttribute(mlp2_bwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    SBR_CHECK_CUDA(cudaFuncSetAttribute(mlp2_bwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    SBR_CHECK_CUDA(cudaFuncSetAttribute(mlp2_bwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    SBR_CHECK_CUDA(cudaFuncSetAttribute(mlp2_bwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    configured = true;
  }
  const int64_t grid = mlp2_grid(n_rows);
  const bool poll = mlp2_poll();
  auto kern = d->n_layers == 1 ? (poll ? mlp2_bwd_kernel<1, true> : mlp2_bwd_kernel<1, false>)
                               : (poll ? mlp2_bwd_kernel<2, true> : mlp2_bwd_kernel<2, false>);
  SBR_CHECK_CUDA(sbr_launch(kern, dim3((unsigned)grid), dim3(N_THREADS), (size_t)BWD_SMEM, S(stream), tm[0], tm[1], p));
  return SBR_OK;
}
