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
  // BatchNorm statistics finalised by the CTA that finishes last (counter == nullptr: not done here)
  unsigned int* bn_counter;
  float bn_eps, bn_momentum;
  float* bn_mean_invstd;
  float* bn_running_mean;
  float* bn_running_var;
  int64_t* bn_nbt;
};
struct BwdParams {
  int debug;  // SBR_MLP2_DEBUG bit mask (profiling only): 1 = no gradient flush, 2 = no gather loads, 4 = no dy / z loads,
              // 8 = no dX0 stores, 16 = wait-cycle profile of block 0, 32 = no L2 prefetch of dy / z
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

// Profiling only (SBR_MLP2_DEBUG bit 16 selects the PROF instantiation): warp 0 (producers), the MMA warp and the first
// epilogue warp of block 0 accumulate the SM cycles they spend inside every barrier wait; slot 0 of a role is the length of
// its tile loop.  Read back by sbr_mlp2_trace_read (scripts/trace_mlp2.py): loop - sum(waits) = the role's own work.
constexpr int PROF_SLOTS = 16;  // 1-7: waits, 8-15: sections of the role's own work
__device__ unsigned long long g_prof[4 * PROF_SLOTS];
// SBR_MLP2_DEBUG bit 64: every CTA of a backward launch with more than 100 000 rows stamps %globaltimer at its start and
// end (sbr_mlp2_cta_times_read): start skew / tail of the persistent grid inside the real step
__device__ unsigned long long g_cta_times[2 * 1024];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <bool PROF>
__device__ __forceinline__ void lap(uint32_t (&acc)[PROF_SLOTS], uint32_t& tmark, int work_slot) {
  if (PROF) {
    const uint32_t now = (uint32_t)clock();
    acc[work_slot] += now - tmark;
    tmark = now;
  }
}
// barrier wait; PROF: the time since the last lap goes to `work_slot`, the wait itself to `wait_slot`
template <bool PROF>
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity, uint32_t (&acc)[PROF_SLOTS], uint32_t& tmark,
                                         int wait_slot, int work_slot) {
  if (PROF) {
    lap<PROF>(acc, tmark, work_slot);
    mbar_wait(bar, parity);
    lap<PROF>(acc, tmark, wait_slot);
  } else {
    mbar_wait(bar, parity);
  }
}
template <bool PROF>
__device__ __forceinline__ void prof_flush(int role, uint32_t t_loop, const uint32_t (&acc)[PROF_SLOTS]) {
  if (PROF && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
    g_prof[role * PROF_SLOTS] = (uint32_t)clock() - t_loop;
#pragma unroll
    for (int i = 1; i < PROF_SLOTS; ++i) g_prof[role * PROF_SLOTS + i] = acc[i];
  }
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

// 16 values per lane: on return lanes 2m and 2m+1 hold the sum over the warp's lanes of v[m]
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 16, n = 8; off >= 2; off >>= 1, n >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      float send = upper ? v[i] : v[i + n];
      float keep = upper ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
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

// rows of a tile owned by gather / dz warp w (4 lane groups of 8 lanes, 8 passes): 16 i + 4 w + j  (i < 8, j < 4)
__device__ __forceinline__ int warp_row(int w, int lane) { return 16 * (lane >> 2) + 4 * w + (lane & 3); }

// X0 producer loop of one gather warp (w = 0..3 of the 4 gather warps): X0 = dropout(normalise(gather)) of every tile of
// this CTA -> the SWIZZLE_128B stage (it & 1).  A warp resolves the 32 rows it gathers itself (no CTA-level barrier), and
// the dependent index loads of tile t+1 (entity index -> feature row -> category) are issued between the gather passes of
// tile t.  x_full[s] expects one arrival per gather warp; x_empty[s] completes when the stage may be overwritten.
template <bool PROF>
__device__ __forceinline__ void gather_producer(const GatherArgs& g, const sbr_modality_src_t* s_src, const float** s_ptr,
                                                uint8_t* sX, uint64_t* x_full, uint64_t* x_empty, int w, int lane,
                                                int64_t num_tiles, bool no_loads, uint32_t (&wacc)[PROF_SLOTS],
                                                uint32_t& tmark) {
  const uint64_t step = g.step_dev ? (uint64_t)*g.step_dev : 0;
  const int grp = 4 * w + (lane >> 3), li = lane & 7;
  const int my_row = warp_row(w, lane);
  const float sc = g.p_drop > 0.f ? 1.f / (1.f - g.p_drop) : 1.f;
  const bool vec_ok = (g.C & 3) == 0;
  const bool hash_drop = g.p_drop > 0.f && g.keep_mask == nullptr;
  const uint32_t key0 = philox_key0(g.seed, step);
  const uint32_t thr_hi = min(drop_threshold(g.p_drop), 0xFFFFu) << 16;
  const uint32_t base_off = tile_off(grp, li);  // this thread's chunk of row `grp`; row 16 i + grp is 2048 i further
  // resolver state of the NEXT tile's row `my_row`
  int64_t r_e = 0, r_feat = -1, r_src = 0;
  int r_m = 0;
  bool r_ok = false;
  auto stage1 = [&](int64_t tile) {
    const int64_t gr = tile * TILE_ROWS + my_row;
    r_ok = tile < num_tiles && gr < g.N;
    if (r_ok) {
      r_m = g.mods ? min((int)__ldg(g.mods + gr), g.n_mods - 1) : 0;
      r_e = __ldg(g.idx + (g.k == 1 ? gr : gr / g.k));
    }
  };
  auto stage2 = [&]() {
    if (r_ok) {
      const sbr_modality_src_t& s = s_src[r_m];
      r_feat = s.remap ? (int64_t)__ldg(s.remap + r_e) : r_e;
    }
  };
  auto stage3 = [&]() {
    if (r_ok && r_feat >= 0)
      r_src = (s_src[r_m].kind == SBR_SRC_CATEGORICAL) ? (int64_t)__ldg(s_src[r_m].codes + r_feat) : r_feat;
  };
  auto publish = [&]() {  // -> s_ptr[my_row]
    const float* ptr = nullptr;
    if (r_ok) {
      if (r_feat < 0) {
        if (g.err_flag) atomicExch(g.err_flag, 1);
      } else {
        ptr = s_src[r_m].table + r_src * g.C;
      }
    }
    s_ptr[my_row] = ptr;
  };
  stage1(blockIdx.x);
  stage2();
  stage3();
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    uint8_t* dst = sX + s * TILE_BYTES;
    publish();
    __syncwarp();
    stage1(tile + gridDim.x);
    wait_bar<PROF>(&x_empty[s], (uint32_t)(((it >> 1) & 1) ^ 1), wacc, tmark, 1, 8);
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
      if (pass == 0) stage2(); else stage3();
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
        // (rows beyond N were loaded as zeros)
        if (hash_drop) {
          dropout8_hash(x[u], philox_group_k(gr, li, key0), thr_hi, sc);
        } else if (g.p_drop > 0.f) {
          const uint32_t km = gr < g.N ? keep8(g.keep_mask, gr, g.C, 8 * li, g.p_drop, g.seed, step) : 0u;
#pragma unroll
          for (int j = 0; j < 8; ++j) x[u][j] = ((km >> j) & 1u) ? x[u][j] * sc : 0.f;
        }
        *reinterpret_cast<uint4*>(dst + base_off + (pass + u) * 2048) = pack8(x[u]);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();  // (also: every lane has read s_ptr before the next publish)
    if (lane == 0) mbar_arrive(&x_full[s]);
    lap<PROF>(wacc, tmark, 9);
  }
}

template <bool GENERIC, int NE>
__device__ __forceinline__ void act_t(int act, float (&v)[NE]) {
  if (GENERIC) {
    act_n<NE>(act, v);
  } else if (act == SBR_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < NE; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}
template <bool GENERIC, int NE>
__device__ __forceinline__ void actgrad_t(int act, float (&v)[NE], const float (&y)[NE]) {
  if (GENERIC) {
    actgrad_n<NE>(act, v, y);
  } else if (act == SBR_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < NE; ++j) v[j] = y[j] > 0.f ? v[j] : 0.f;
  }
}

// ================================================================================================ forward
// Warp roles (13 warps, two CTAs per SM): 0-7 epilogue (warp & 3 = TMEM lane quarter, warp >> 2 = 32-column half of the
// row), 8-11 X0 gather, 12 MMA issuer.
constexpr int FWD_THREADS = 13 * 32;
constexpr int FW_P0 = 8, FW_MMA = 12;
constexpr int FWD_SMEM = 2 * W_BYTES + 3 * TILE_BYTES + 3072 + 1024;

template <int L, bool PROF, bool GENERIC>
__global__ void __launch_bounds__(512, 2)  // 13 warps are allocated as 16: 64 registers per thread for two CTAs per SM
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
  uint64_t* x_full = bars + 1;      // [2] (4 gather warps)
  uint64_t* x_empty = bars + 3;     // [2]
  uint64_t* dh_full = bars + 5;     // hidden accumulator ready
  uint64_t* a1_full = bars + 6;     // hidden activation written (8 epilogue warps)
  uint64_t* df_full = bars + 7;     // final accumulator ready
  uint64_t* df_empty = bars + 8;    // final accumulator drained (8 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  sbr_modality_src_t* s_src = reinterpret_cast<sbr_modality_src_t*>(bars + 12);
  const float** s_ptr = reinterpret_cast<const float**>(s_src + MAX_SRC);  // [128] source row of every tile row
  float* s_bias = reinterpret_cast<float*>(s_ptr + TILE_ROWS);                 // [2][64], zero beyond the widths

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.g.N + TILE_ROWS - 1) / TILE_ROWS;
  if (warp == FW_MMA && lane == 0) {
    tma_prefetch_desc(&tmW0);
    if (L == 2) tma_prefetch_desc(&tmW1);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 4);
      mbar_init(&x_empty[s], 1);
    }
    mbar_init(dh_full, 1);
    mbar_init(a1_full, 8);
    mbar_init(df_full, 1);
    mbar_init(df_empty, 8);
    fence_barrier_init();
  }
  if (warp == FW_MMA) tmem_alloc(tmem_slot, 128);
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
  uint32_t wacc[PROF_SLOTS] = {};  // (PROF only; dead otherwise)
  const uint32_t t_loop = (uint32_t)clock();
  uint32_t tmark = t_loop;

  if (warp >= FW_P0 && warp < FW_MMA) {
    // ---------------------------------------------------------------- producers: gather + normalise + dropout -> X0
    gather_producer<PROF>(p.g, s_src, s_ptr, sX, x_full, x_empty, warp - FW_P0, lane, num_tiles, (p.debug & 2) != 0, wacc,
                          tmark);
  } else if (warp == FW_MMA) {
    // ---------------------------------------------------------------- MMA issuer (converged warp, elected issue)
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, (L == 2 ? 2 : 1) * W_BYTES);
      tma_load_2d(sW0, &tmW0, w_full, 0, 0);
      if (L == 2) tma_load_2d(sW1, &tmW1, w_full, 0, 0);
    }
    __syncwarp();
    wait_bar<PROF>(w_full, 0, wacc, tmark, 7, 15);
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t aX = smem_u32(sX), aA1 = smem_u32(sA1), aW0 = smem_u32(sW0), aW1 = smem_u32(sW1);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      wait_bar<PROF>(&x_full[s], (uint32_t)((it >> 1) & 1), wacc, tmark, 1, 8);
      if (L == 1) wait_bar<PROF>(df_empty, (uint32_t)((it & 1) ^ 1), wacc, tmark, 2, 8);
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
        wait_bar<PROF>(a1_full, (uint32_t)(it & 1), wacc, tmark, 3, 8);
        wait_bar<PROF>(df_empty, (uint32_t)((it & 1) ^ 1), wacc, tmark, 2, 8);
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
    // ---------------------------------------------------------------- epilogue: thread = row, 32 of its 64 columns
    const int q = warp & 3, half = warp >> 2;
    const int cb = 32 * half;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int act0 = p.l[0].act, act_last = p.l[L - 1].act;
    const uint32_t e_row = (uint32_t)((row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128), e_r7 = (uint32_t)(row_in_tile & 7);
    float cs_acc[2] = {0.f, 0.f}, cq_acc[2] = {0.f, 0.f};  // lanes 2m, 2m+1: column cb + 16 i + m
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int64_t row = tile * TILE_ROWS + row_in_tile;
      const bool row_ok = row < p.g.N;
      if (L == 2) {
        wait_bar<PROF>(dh_full, (uint32_t)(it & 1), wacc, tmark, 1, 8);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_h + lane_off + cb + c0, r);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(s_bias + cb + c0 + j);
            v[j] = __uint_as_float(r[j]) + b.x;
            v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + b.z;
            v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
          }
          act_t<GENERIC, 16>(act0, v);
          // (columns beyond the layer's width hold act(0): the next weight's K columns there are zero)
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            float w8[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w8[t] = v[j + t];
            *reinterpret_cast<uint4*>(sA1 + e_row + ((((uint32_t)(cb + c0 + j) >> 3) ^ e_r7) << 4)) = pack8(w8);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a1_full);
      }
      wait_bar<PROF>(df_full, (uint32_t)(it & 1), wacc, tmark, 2, 9);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        if (cb + c0 >= p.D) break;  // warp-uniform
        uint32_t r[16];
        tmem_ld16(tmem_f + lane_off + cb + c0, r);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(s_bias + (L - 1) * 64 + cb + c0 + j);
          v[j] = __uint_as_float(r[j]) + b.x;
          v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
          v[j + 2] = __uint_as_float(r[j + 2]) + b.z;
          v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
        }
        act_t<GENERIC, 16>(act_last, v);
        if (row_ok && !(p.debug & 8)) {
          float* dst = p.z + row * p.ldz + cb + c0;
          if (cb + c0 + 16 <= p.D && (p.ldz & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (cb + c0 + j < p.D) dst[j] = v[j];
          }
        }
        if (p.colstats != nullptr) {
          float s1[16], s2[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = row_ok ? v[j] : 0.f;
            s1[j] = x;
            s2[j] = x * x;
          }
          cs_acc[c0 >> 4] += warp_colsum16(s1, lane);
          cq_acc[c0 >> 4] += warp_colsum16(s2, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(df_empty);
      lap<PROF>(wacc, tmark, 10);
    }
    if (p.colstats != nullptr) {
      // one row of partial sums per CTA: the four lane quarters are added in a fixed order (deterministic statistics;
      // sbr_bn_finalize adds the rows of all CTAs in a fixed order as well)
      float* s_part = reinterpret_cast<float*>(sA1);  // (the hidden-activation tile is no longer needed)
      if (!(lane & 1)) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          s_part[q * 128 + cb + 16 * i + (lane >> 1)] = cs_acc[i];
          s_part[q * 128 + 64 + cb + 16 * i + (lane >> 1)] = cq_acc[i];
        }
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // the eight epilogue warps
      if (q == 0) {
        float* rowp = p.colstats + (size_t)blockIdx.x * 2 * p.D;
        const int col = cb + lane;
        if (col < p.D) {
          rowp[col] = ((s_part[col] + s_part[128 + col]) + s_part[256 + col]) + s_part[384 + col];
          rowp[p.D + col] = ((s_part[64 + col] + s_part[128 + 64 + col]) + s_part[256 + 64 + col]) + s_part[384 + 64 + col];
        }
      }
    }
  }

  if (PROF && (lane == 0) && (warp == 0 || warp == FW_P0 || warp == FW_MMA))
    prof_flush<PROF>(warp == FW_P0 ? 0 : (warp == FW_MMA ? 1 : 2), t_loop, wacc);
  tc_fence_before();
  if (p.bn_counter != nullptr) __threadfence();  // this CTA's row of partial sums is visible before its ticket is taken
  __syncthreads();
  if (warp == FW_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
  if (p.bn_counter == nullptr) return;
  // ---- BatchNorm statistics: the CTA that takes the last ticket adds the rows of every CTA in a fixed order (13 row lanes
  // of 32 float4 column groups, then the lanes in order): the result does not depend on which CTA does it
  __shared__ unsigned int s_ticket;
  if (threadIdx.x == 0) s_ticket = atomicAdd(p.bn_counter, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  float* s_red = reinterpret_cast<float*>(sX);  // [13][128]
  const int W2 = 2 * p.D;                       // floats per row (D <= 64)
  const int g4 = lane * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g4 < W2 && (W2 & 3) == 0) {
#pragma unroll 4
    for (int r = warp; r < (int)gridDim.x; r += 13) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(p.colstats + (size_t)r * W2 + g4));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  } else if (g4 < W2) {
    for (int r = warp; r < (int)gridDim.x; r += 13) {
      const float* row = p.colstats + (size_t)r * W2 + g4;
      acc.x += __ldcg(row);
      if (g4 + 1 < W2) acc.y += __ldcg(row + 1);
      if (g4 + 2 < W2) acc.z += __ldcg(row + 2);
      if (g4 + 3 < W2) acc.w += __ldcg(row + 3);
    }
  }
  *reinterpret_cast<float4*>(s_red + warp * 128 + g4) = acc;
  __syncthreads();
  if (threadIdx.x < p.D) {
    const int c = threadIdx.x;
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int wv = 0; wv < 13; ++wv) {
      a += s_red[wv * 128 + c];
      b += s_red[wv * 128 + p.D + c];
    }
    const double n = (double)p.g.N;
    const double mean = (double)a / n;
    double var = (double)b / n - mean * mean;
    if (var < 0.) var = 0.;
    p.bn_mean_invstd[c] = (float)mean;
    p.bn_mean_invstd[p.D + c] = (float)(1.0 / sqrt(var + (double)p.bn_eps));
    if (p.bn_running_mean) p.bn_running_mean[c] = (1.f - p.bn_momentum) * p.bn_running_mean[c] + p.bn_momentum * (float)mean;
    if (p.bn_running_var) {
      const double unbiased = n > 1. ? var * n / (n - 1.) : var;
      p.bn_running_var[c] = (1.f - p.bn_momentum) * p.bn_running_var[c] + p.bn_momentum * (float)unbiased;
    }
  }
  if (threadIdx.x == 0) {
    if (p.bn_nbt) *p.bn_nbt += 1;
    *p.bn_counter = 0u;
  }
}

// ================================================================================================ backward
// Warp roles (13 warps, two CTAs per SM): 0-3 epilogue (warp = TMEM lane quarter, thread = row of the tile), 4-7 X0
// gather, 8-11 dz producers (+ the fp32 bias-gradient sums), 12 MMA issuer.  Every role runs at one warp per scheduler, so
// the tile time is the longest role, not their sum: the gather of tile t+1, the dz of tile t+1 (its buffer is released
// by the commit behind  dz W1  +  the layer-1 weight gradient, both issued as soon as dz and Y1 exist) and X0(t+1) W0^T
// (own accumulator) all run while the epilogue warps walk Y1 -> dY1 -> dX0 of tile t.
// shared memory: W0 | W1 | DZ | DY1 | Y1 | X0[0] | X0[1].  The weight-gradient MMAs take A = [DZ ; DY1] stacked on M (second
// 64-row block one tile further) against B = Y1 (-> accumulator columns 0-63, lanes 0-63 = layer 1) and B = X0[s]
// (-> columns 64-127, lanes 64-127 = layer 0); the other two quadrants hold cross terms nobody reads.
constexpr int BWD_THREADS = 13 * 32;
constexpr int BW_PG0 = 4, BW_PZ0 = 8, BW_MMA = 12;
constexpr int BWD_MISC = 4096;
constexpr int BWD_SMEM = 2 * W_BYTES + 5 * TILE_BYTES + BWD_MISC + 1024;

template <int L, bool PROF, bool GENERIC>
__global__ void __launch_bounds__(512, 2)  // 13 warps are allocated as 16: 64 registers per thread for two CTAs per SM
mlp2_bwd_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmW1, BwdParams p) {
  SBR_PDL_LAUNCH();
  const bool stamp = (p.debug & 64) && p.g.N > 100000 && threadIdx.x == 0 && blockIdx.x < 1024;
  if (stamp) g_cta_times[2 * blockIdx.x] = globaltimer_ns();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* sW0 = smem;
  uint8_t* sW1 = smem + W_BYTES;
  uint8_t* sDZ = smem + 2 * W_BYTES;
  uint8_t* sDY1 = sDZ + TILE_BYTES;
  uint8_t* sY1 = sDY1 + TILE_BYTES;
  uint8_t* sX = sY1 + TILE_BYTES;  // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * TILE_BYTES);
  uint64_t* w_full = bars;         // weights landed
  uint64_t* x_full = bars + 1;     // [2] X0 gathered (4 gather warps)
  uint64_t* x_empty = bars + 3;    // [2] the last MMA reading the stage has completed
  uint64_t* y_full = bars + 5;     // X0 W0^T in its accumulator
  uint64_t* y1_full = bars + 6;    // Y1 written (4 epilogue warps)
  uint64_t* dz_full = bars + 7;    // dz written (4 dz warps)
  uint64_t* dz_free = bars + 8;    // dz W1 and the layer-1 weight gradient have completed
  uint64_t* a_full = bars + 9;     // the shared accumulator holds dz W1 / dY1 W0 (two completions per tile, L == 2)
  uint64_t* dy1_full = bars + 10;  // dY1 written (4 epilogue warps)
  uint64_t* dy1_free = bars + 11;  // the dz warps have added dY1 into the layer-0 bias gradient (4 warps)
  uint64_t* da_free = bars + 12;   // L == 1: dX0 read out of the accumulator (4 epilogue warps)
  uint64_t* w_done = bars + 13;    // every MMA of this CTA has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  sbr_modality_src_t* s_src = reinterpret_cast<sbr_modality_src_t*>(bars + 16);
  const float** s_ptr = reinterpret_cast<const float**>(s_src + MAX_SRC);  // [128] source row of every tile row
  float* s_bias = reinterpret_cast<float*>(s_ptr + TILE_ROWS);                 // [2][64], zero beyond the widths
  float* s_coef = s_bias + 128;                                                // [4][64] BatchNorm-backward coefficients
  static_assert(16 * 8 + MAX_SRC * sizeof(sbr_modality_src_t) + TILE_ROWS * 8 + 128 * 4 + 256 * 4 <= BWD_MISC, "misc");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t num_tiles = (p.g.N + TILE_ROWS - 1) / TILE_ROWS;
  if (warp == BW_MMA && lane == 0) {
    tma_prefetch_desc(&tmW0);
    if (L == 2) tma_prefetch_desc(&tmW1);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&x_full[s], 4);
      mbar_init(&x_empty[s], 1);
    }
    mbar_init(y_full, 1);
    mbar_init(y1_full, 4);
    mbar_init(dz_full, 4);
    mbar_init(dz_free, 1);
    mbar_init(a_full, 1);
    mbar_init(dy1_full, 4);
    mbar_init(dy1_free, 4);
    mbar_init(da_free, 4);
    mbar_init(w_done, 1);
    fence_barrier_init();
  }
  if (warp == BW_MMA) tmem_alloc(tmem_slot, 256);
  if (L == 1)  // a defined second block of the stacked A operand (its accumulator lanes are never read)
    for (int i = threadIdx.x; i < TILE_BYTES / 16; i += BWD_THREADS)
      reinterpret_cast<uint4*>(sDY1)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base, tmem_y = tmem_base + 64, tmem_w = tmem_base + 128;
  SBR_PDL_WAIT();
  if (threadIdx.x < p.g.n_mods) s_src[threadIdx.x] = p.g.srcs[threadIdx.x];
  if (threadIdx.x >= 128 && threadIdx.x < 256) {
    const int l = (threadIdx.x - 128) >> 6, c = threadIdx.x & 63;
    s_bias[l * 64 + c] = (l < L && p.l[l].bias != nullptr && c < p.l[l].out_f) ? p.l[l].bias[c] : 0.f;
  }
  if (threadIdx.x >= 256 && threadIdx.x < 320) {
    // per-column coefficients  dz = A dy + B (z - mean) + C0  (BatchNorm backward; A = 1, B = C0 = 0 without one)
    const int c = threadIdx.x - 256;
    float cA = 1.f, cB = 0.f, cM = 0.f, c0v = 0.f;
    if (p.mean_invstd != nullptr && c < p.D) {
      float s0 = 0.f, s1 = 0.f;
      for (int r = 0; r < p.n_replicas; ++r) {
        s0 += p.sums[(size_t)r * 2 * p.D + c];
        s1 += p.sums[(size_t)r * 2 * p.D + p.D + c];
      }
      const float inv_n = 1.f / (float)p.g.N;
      const float istd = p.mean_invstd[p.D + c];
      const float gi = p.gamma[c] * istd;
      cM = p.mean_invstd[c];
      cA = gi;
      cB = -gi * istd * (s1 * inv_n);
      c0v = -gi * (s0 * inv_n);
      if (blockIdx.x == 0) {  // d gamma / d beta come with the sums
        if (p.dbeta) p.dbeta[c] += s0;
        if (p.dgamma) p.dgamma[c] += s1;
      }
    }
    s_coef[c] = cA; s_coef[64 + c] = cB; s_coef[128 + c] = cM; s_coef[192 + c] = c0v;
  }
  __syncthreads();
  const int my_tiles = blockIdx.x < num_tiles ? (int)((num_tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  uint32_t wacc[PROF_SLOTS] = {};  // (PROF only; dead otherwise)
  const uint32_t t_loop = (uint32_t)clock();
  uint32_t tmark = t_loop;

  if (warp >= BW_PG0 && warp < BW_PZ0) {
    // ---------------------------------------------------------------- X0 gather (one tile ahead of the MMAs)
    gather_producer<PROF>(p.g, s_src, s_ptr, sX, x_full, x_empty, warp - BW_PG0, lane, num_tiles, (p.debug & 2) != 0, wacc,
                          tmark);
  } else if (warp >= BW_PZ0 && warp < BW_MMA) {
    // ---------------------------------------------------------------- dz producers + bias-gradient sums
    const int w = warp - BW_PZ0;
    const int grp = 4 * w + (lane >> 3), li = lane & 7;
    const int c0 = 8 * li;
    const int act_last = p.l[L - 1].act;
    float bsum[8];  // fp32 column sums of dz (the last layer's bias gradient)
#pragma unroll
    for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
    const bool vec_ok = (p.lddy & 3) == 0 && (p.ldz & 3) == 0 && (p.D & 7) == 0;
    const bool no_dz_loads = (p.debug & 4) != 0;
    const bool want_gb = p.gb[L - 1] != nullptr;
    const uint32_t base_off = tile_off(grp, li);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      // 16 rows per pass, 2 passes in flight; the first loads are issued before the wait for the buffer.  Full tiles of
      // full-width rows (all but the last tile) take the path without per-element predicates.
      const int64_t r0 = tile * TILE_ROWS + grp;
      const bool full = vec_ok && p.D == 64 && tile * TILE_ROWS + TILE_ROWS <= p.g.N && !no_dz_loads;
      if (!(p.debug & 32)) {
        // dy / z of this CTA's NEXT tile -> L2 (thread = row, one prefetch per 128-byte line): the loads of the next
        // iteration see L2 latency, and a whole tile of DRAM requests is in flight without registers or shared memory
        const int64_t nr = (tile + gridDim.x) * TILE_ROWS + (threadIdx.x - BW_PZ0 * 32);
        if (nr < p.g.N && !no_dz_loads) {
          const char* a = reinterpret_cast<const char*>(p.dy + nr * p.lddy);
          const char* b = reinterpret_cast<const char*>(p.z + nr * p.ldz);
          for (int o = 0; o < p.D * 4; o += 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
          }
        }
      }
      const float* dyp = p.dy + r0 * p.lddy + c0;
      const float* zp = p.z + r0 * p.ldz + c0;
#pragma unroll 1
      for (int pass = 0; pass < 8; pass += 2) {
        float gy[2][8], zz[2][8];
        if (full) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(dyp + (int64_t)u * 16 * p.lddy));
            const float4 b = __ldcs(reinterpret_cast<const float4*>(dyp + (int64_t)u * 16 * p.lddy + 4));
            const float4 c = __ldcs(reinterpret_cast<const float4*>(zp + (int64_t)u * 16 * p.ldz));
            const float4 d = __ldcs(reinterpret_cast<const float4*>(zp + (int64_t)u * 16 * p.ldz + 4));
            gy[u][0] = a.x; gy[u][1] = a.y; gy[u][2] = a.z; gy[u][3] = a.w;
            gy[u][4] = b.x; gy[u][5] = b.y; gy[u][6] = b.z; gy[u][7] = b.w;
            zz[u][0] = c.x; zz[u][1] = c.y; zz[u][2] = c.z; zz[u][3] = c.w;
            zz[u][4] = d.x; zz[u][5] = d.y; zz[u][6] = d.z; zz[u][7] = d.w;
          }
        } else {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int64_t gr = r0 + (pass + u) * 16;
            const float* dq = dyp + (int64_t)u * 16 * p.lddy;
            const float* zq = zp + (int64_t)u * 16 * p.ldz;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const bool in = gr < p.g.N && c0 + j < p.D && !no_dz_loads;
              gy[u][j] = in ? dq[j] : 0.f;
              zz[u][j] = in ? zq[j] : 0.f;
            }
          }
        }
        dyp += (int64_t)32 * p.lddy;
        zp += (int64_t)32 * p.ldz;
        if (pass == 0) wait_bar<PROF>(dz_free, (uint32_t)((it & 1) ^ 1), wacc, tmark, 1, 8);
        // dz in place of dy, four columns at a time (the coefficients are broadcast reads of shared memory)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 a = *reinterpret_cast<const float4*>(s_coef + c0 + 4 * h);
          const float4 b = *reinterpret_cast<const float4*>(s_coef + 64 + c0 + 4 * h);
          const float4 m = *reinterpret_cast<const float4*>(s_coef + 128 + c0 + 4 * h);
          const float4 c = *reinterpret_cast<const float4*>(s_coef + 192 + c0 + 4 * h);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            gy[u][4 * h] = a.x * gy[u][4 * h] + b.x * (zz[u][4 * h] - m.x) + c.x;
            gy[u][4 * h + 1] = a.y * gy[u][4 * h + 1] + b.y * (zz[u][4 * h + 1] - m.y) + c.y;
            gy[u][4 * h + 2] = a.z * gy[u][4 * h + 2] + b.z * (zz[u][4 * h + 2] - m.z) + c.z;
            gy[u][4 * h + 3] = a.w * gy[u][4 * h + 3] + b.w * (zz[u][4 * h + 3] - m.w) + c.w;
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float (&v)[8] = gy[u];
          actgrad_t<GENERIC, 8>(act_last, v, zz[u]);
          if (!full) {  // rows beyond N / columns beyond D contribute nothing (the BatchNorm constant is not zero there)
            const bool ok = r0 + (pass + u) * 16 < p.g.N;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (ok && c0 + j < p.D) ? v[j] : 0.f;
          }
          if (want_gb) {  // bias gradient of the last layer: fp32 column sums BEFORE the bf16 rounding
#pragma unroll
            for (int j = 0; j < 8; ++j) bsum[j] += v[j];
          }
          *reinterpret_cast<uint4*>(sDZ + base_off + (pass + u) * 2048) = pack8(v);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(dz_full);
      lap<PROF>(wacc, tmark, 9);
    }
    if (my_tiles > 0 && want_gb) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = bsum[j];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane < 8 && c0 + j < p.D) atomicAdd(p.gb[L - 1] + c0 + j, v);
      }
    }
  } else if (warp == BW_MMA) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, (L == 2 ? 2 : 1) * W_BYTES);
      tma_load_2d(sW0, &tmW0, w_full, 0, 0);
      if (L == 2) tma_load_2d(sW1, &tmW1, w_full, 0, 0);
    }
    __syncwarp();
    wait_bar<PROF>(w_full, 0, wacc, tmark, 7, 15);
    const uint32_t id_fwd = umma_idesc_bf16(128, 64, 0, 0);   // A K-major, B K-major   (forward)
    const uint32_t id_dg = umma_idesc_bf16(128, 64, 0, 1);    // A K-major, B MN-major  (dgrad: B = weight^T view)
    const uint32_t id_wg = umma_idesc_bf16(128, 64, 1, 1);    // both MN-major (contraction over the rows of the tile)
    const uint32_t aX = smem_u32(sX), aY1 = smem_u32(sY1), aDZ = smem_u32(sDZ), aDY1 = smem_u32(sDY1),
                   aW0 = smem_u32(sW0), aW1 = smem_u32(sW1);
    if (L == 2 && my_tiles > 0) {  // X0(0) W0^T
      wait_bar<PROF>(&x_full[0], 0u, wacc, tmark, 2, 8);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_y, desc_k(aX, k), desc_k(aW0, k), id_fwd, k > 0 ? 1u : 0u);
        umma_commit(y_full);
      }
      __syncwarp();
    }
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t ph = (uint32_t)(it & 1);
      const int s = it & 1;
      const uint32_t aXs = aX + s * TILE_BYTES;
      if (L == 2) {
        wait_bar<PROF>(y1_full, ph, wacc, tmark, 3, 8);  // Y1 in shared memory; dX0 of the previous tile read out
        wait_bar<PROF>(dz_full, ph, wacc, tmark, 4, 8);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dz W1  (contraction over out_1)
            umma_bf16(tmem_a, desc_k(aDZ, k), desc_mn(aW1, k, W_BYTES), id_dg, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // [dz ; .]^T Y1 -> columns 0-63
            umma_bf16(tmem_w, desc_mn(aDZ, k, TILE_BYTES), desc_mn(aY1, k, TILE_BYTES), id_wg, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(a_full);
          umma_commit(dz_free);
        }
        __syncwarp();
        wait_bar<PROF>(dy1_full, ph, wacc, tmark, 5, 8);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dX0 = dY1 W0
            umma_bf16(tmem_a, desc_k(aDY1, k), desc_mn(aW0, k, W_BYTES), id_dg, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // [. ; dY1]^T X0 -> columns 64-127
            umma_bf16(tmem_w + 64, desc_mn(aDZ, k, TILE_BYTES), desc_mn(aXs, k, TILE_BYTES), id_wg, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(a_full);
          umma_commit(&x_empty[s]);
        }
        __syncwarp();
        if (it + 1 < my_tiles) {  // X0(t+1) W0^T into its own accumulator (read out by the epilogue before y1_full(t))
          wait_bar<PROF>(&x_full[s ^ 1], (uint32_t)(((it + 1) >> 1) & 1), wacc, tmark, 2, 8);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_y, desc_k(aX + (s ^ 1) * TILE_BYTES, k), desc_k(aW0, k), id_fwd, k > 0 ? 1u : 0u);
            umma_commit(y_full);
          }
          __syncwarp();
        }
      } else {
        wait_bar<PROF>(da_free, ph ^ 1, wacc, tmark, 1, 8);  // dX0 of the previous tile has been read out
        wait_bar<PROF>(&x_full[s], (uint32_t)((it >> 1) & 1), wacc, tmark, 2, 8);
        wait_bar<PROF>(dz_full, ph, wacc, tmark, 4, 8);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dX0 = dz W0
            umma_bf16(tmem_a, desc_k(aDZ, k), desc_mn(aW0, k, W_BYTES), id_dg, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // [dz ; 0]^T X0 -> columns 0-63
            umma_bf16(tmem_w, desc_mn(aDZ, k, TILE_BYTES), desc_mn(aXs, k, TILE_BYTES), id_wg, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(a_full);
          umma_commit(&x_empty[s]);
          umma_commit(dz_free);
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(w_done);
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue: thread = row of the tile
    const int q = warp;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t a_ph = 0;  // parity of the next completion of a_full
    const int act0 = p.l[0].act;
    const bool want_gb0 = L == 2 && p.gb[0] != nullptr;
    const uint32_t e_row = (uint32_t)((row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128), e_r7 = (uint32_t)(row_in_tile & 7);
    auto e_off = [&](int chunk) { return e_row + (((uint32_t)chunk ^ e_r7) << 4); };  // = tile_off(row_in_tile, chunk)
    float cb_acc[4] = {0.f, 0.f, 0.f, 0.f};  // bias gradient of layer 0: lanes 2m, 2m+1 hold column 16 i + m
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int64_t row = tile * TILE_ROWS + row_in_tile;
      const bool row_ok = row < p.g.N;
      if (L == 2) {
        // Y1 = act(X0 W0^T + b0) -> shared memory (bf16)
        wait_bar<PROF>(y_full, (uint32_t)(it & 1), wacc, tmark, 1, 8);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_y + lane_off + c0, r);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(s_bias + c0 + j);
            v[j] = __uint_as_float(r[j]) + b.x;
            v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + b.z;
            v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
          }
          act_t<GENERIC, 16>(act0, v);
          // (columns beyond the layer's width hold act(0): the next weight's K columns there are zero)
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            float w8[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w8[t] = v[j + t];
            *reinterpret_cast<uint4*>(sY1 + e_off((c0 + j) >> 3)) = pack8(w8);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(y1_full);
        // dY1 = (dz W1) * act'(Y1)
        wait_bar<PROF>(a_full, a_ph, wacc, tmark, 2, 9);
        a_ph ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_a + lane_off + c0, r);
          // (rows beyond N have dz = 0, columns beyond the width meet zero weight columns: no masks needed)
          float v[16], yv[16];
#pragma unroll
          for (int j = 0; j < 16; j += 8) {  // Y1 of this row back from its shared-memory tile (bf16)
            const uint4 u = *reinterpret_cast<const uint4*>(sY1 + e_off((c0 + j) >> 3));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 y = __bfloat1622float2(h[t]);
              yv[j + 2 * t] = y.x;
              yv[j + 2 * t + 1] = y.y;
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          actgrad_t<GENERIC, 16>(act0, v, yv);
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            float w8[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w8[t] = v[j + t];
            *reinterpret_cast<uint4*>(sDY1 + e_off((c0 + j) >> 3)) = pack8(w8);
          }
          if (want_gb0) cb_acc[c0 >> 4] += warp_colsum16(v, lane);  // fp32, before the bf16 rounding
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(dy1_full);
      }
      // dX0 -> global (fp32, consumed by the sorted-run gather backward)
      wait_bar<PROF>(a_full, a_ph, wacc, tmark, 3, 10);
      a_ph ^= 1;
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        if (c0 >= p.g.C) break;
        uint32_t r[16];
        tmem_ld16(tmem_a + lane_off + c0, r);
        tmem_ld_wait();
        if (row_ok && !(p.debug & 8)) {
          float* dst = p.dx + row * p.lddx + c0;
          if (c0 + 16 <= p.g.C && (p.lddx & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                 __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < p.g.C) dst[j] = __uint_as_float(r[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (L == 1 && lane == 0) mbar_arrive(da_free);
      lap<PROF>(wacc, tmark, 11);
    }
    // ---- flush the weight-gradient accumulators
    if (my_tiles > 0) {
      if (want_gb0 && !(lane & 1)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int col = 16 * i + (lane >> 1);
          if (col < p.l[0].out_f) atomicAdd(p.gb[0] + col, cb_acc[i]);
        }
      }
      wait_bar<PROF>(w_done, 0u, wacc, tmark, 5, 12);
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
        for (int c0 = 0; c0 < 64; c0 += 16) {
          if (c0 >= la.in_f) break;
          uint32_t r[16];
          tmem_ld16(tmem_w + lane_off + col_base + c0, r);
          tmem_ld_wait();
          if (o < la.out_f && gw != nullptr) {
            float* dst = gw + (int64_t)o * la.in_f + c0;
            if (c0 + 16 <= la.in_f && (la.in_f & 3) == 0 && (reinterpret_cast<uintptr_t>(gw) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const size_t a = __cvta_generic_to_global(dst + j);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(__uint_as_float(r[j])),
                             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                             "f"(__uint_as_float(r[j + 3]))
                             : "memory");
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < la.in_f) atomicAdd(dst + j, __uint_as_float(r[j]));
            }
          }
        }
      }
    }
  }

  if (PROF && (lane == 0) && (warp == 0 || warp == BW_PG0 || warp == BW_PZ0 || warp == BW_MMA))
    prof_flush<PROF>(warp == BW_PG0 ? 0 : (warp == BW_MMA ? 1 : (warp == 0 ? 2 : 3)), t_loop, wacc);
  tc_fence_before();
  __syncthreads();
  if (warp == BW_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
  if (stamp) g_cta_times[2 * blockIdx.x + 1] = globaltimer_ns();
}

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline int mlp2_debug() {
  const char* e = getenv("SBR_MLP2_DEBUG");
  return e ? atoi(e) : 0;
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
  SBR_CHECK_CUDA(cudaDeviceSynchronize());
  const int n = max_events < 4 * PROF_SLOTS ? max_events : 4 * PROF_SLOTS;
  if (n > 0) SBR_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_prof, sizeof(unsigned long long) * n));
  return n;  // (number of words, not a status)
}

extern "C" int sbr_mlp2_cta_times_read(unsigned long long* host_out, int n_ctas) {
  SBR_CHECK_CUDA(cudaDeviceSynchronize());
  const int n = n_ctas < 1024 ? n_ctas : 1024;
  if (n > 0) SBR_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_cta_times, sizeof(unsigned long long) * 2 * n));
  return n;
}

extern "C" int sbr_mlp2_colstats_rows(int64_t n_rows) { return (int)mlp2_grid(n_rows); }

extern "C" int sbr_mlp2_fwd(const sbr_mlp2_desc_t* d, int64_t n_rows, int C, float* z, int64_t ldz, float* colstats,
                            int colstats_rows, void* stream) {
  return sbr_mlp2_fwd_bn(d, n_rows, C, z, ldz, colstats, colstats_rows, nullptr, stream);
}

extern "C" int sbr_mlp2_fwd_bn(const sbr_mlp2_desc_t* d, int64_t n_rows, int C, float* z, int64_t ldz, float* colstats,
                               int colstats_rows, const sbr_mlp2_bn_tail_t* tail, void* stream) {
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
  if (tail != nullptr) {
    SBR_REQUIRE(colstats && tail->counter && tail->mean_invstd, "sbr_mlp2_fwd_bn: colstats, counter and mean_invstd needed");
    p.bn_counter = tail->counter; p.bn_eps = tail->eps; p.bn_momentum = tail->momentum;
    p.bn_mean_invstd = tail->mean_invstd; p.bn_running_mean = tail->running_mean; p.bn_running_var = tail->running_var;
    p.bn_nbt = tail->num_batches_tracked;
  }
  const int64_t grid = mlp2_grid(n_rows);
  SBR_REQUIRE(colstats == nullptr || colstats_rows >= grid, "sbr_mlp2_fwd: colstats_rows=%d < %lld", colstats_rows,
              (long long)grid);
  CUtensorMap tm[2];
  rc = make_weight_maps(d, tm);
  if (rc) return rc;
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, FwdParams);
  static const Kern kerns[8] = {mlp2_fwd_kernel<1, false, false>, mlp2_fwd_kernel<1, false, true>,
                                mlp2_fwd_kernel<1, true, false>,  mlp2_fwd_kernel<1, true, true>,
                                mlp2_fwd_kernel<2, false, false>, mlp2_fwd_kernel<2, false, true>,
                                mlp2_fwd_kernel<2, true, false>,  mlp2_fwd_kernel<2, true, true>};
  static bool configured = false;
  if (!configured) {
    for (int i = 0; i < 8; ++i)
      SBR_CHECK_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    configured = true;
  }
  if (colstats != nullptr && colstats_rows > grid)  // rows no CTA writes must not hold garbage
    SBR_CHECK_CUDA(cudaMemsetAsync(colstats + (size_t)grid * 2 * p.D, 0,
                                   (size_t)(colstats_rows - grid) * 2 * p.D * sizeof(float), S(stream)));
  const bool prof = (p.debug & 16) != 0;
  bool generic = false;  // activations other than none / ReLU take the instantiation with the full switch
  for (int l = 0; l < d->n_layers; ++l) generic |= d->layers[l].act != SBR_ACT_NONE && d->layers[l].act != SBR_ACT_RELU;
  const Kern kern = kerns[(d->n_layers == 2 ? 4 : 0) + (prof ? 2 : 0) + (generic ? 1 : 0)];
  SBR_CHECK_CUDA(sbr_launch(kern, dim3((unsigned)grid), dim3(FWD_THREADS), (size_t)FWD_SMEM, S(stream), tm[0], tm[1], p));
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
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, BwdParams);
  static const Kern kerns[8] = {mlp2_bwd_kernel<1, false, false>, mlp2_bwd_kernel<1, false, true>,
                                mlp2_bwd_kernel<1, true, false>,  mlp2_bwd_kernel<1, true, true>,
                                mlp2_bwd_kernel<2, false, false>, mlp2_bwd_kernel<2, false, true>,
                                mlp2_bwd_kernel<2, true, false>,  mlp2_bwd_kernel<2, true, true>};
  static bool configured = false;
  if (!configured) {
    for (int i = 0; i < 8; ++i)
      SBR_CHECK_CUDA(cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    configured = true;
  }
  const int64_t grid = mlp2_grid(n_rows);
  const bool prof = (p.debug & 16) != 0;
  bool generic = false;  // activations other than none / ReLU take the instantiation with the full switch
  for (int l = 0; l < d->n_layers; ++l) generic |= d->layers[l].act != SBR_ACT_NONE && d->layers[l].act != SBR_ACT_RELU;
  const Kern kern = kerns[(d->n_layers == 2 ? 4 : 0) + (prof ? 2 : 0) + (generic ? 1 : 0)];
  SBR_CHECK_CUDA(sbr_launch(kern, dim3((unsigned)grid), dim3(BWD_THREADS), (size_t)BWD_SMEM, S(stream), tm[0], tm[1], p));
  return SBR_OK;
}
