// sibrar_b200 -- bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// D[M,N] = alpha * A * B^T, A/B each K-major or MN-major (see include/sibrar_b200.h).  One 128 x BN output tile per
// CTA, optional split-K over blockIdx.z.  Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer
// (single elected thread), warps 2..9 (2..5 next to the bit converters) = epilogue (TMEM -> registers -> fused bias / BN statistics / activation /
// activation-gradient / store | transposed atomic accumulate).
//
// Replaces nn.Linear forward in the reference's PolyLinear (modules/polylinear.py:50-76) and the dgrad/wgrad
// GEMMs autograd runs for it (train/trainer.py:221).
#include <stdarg.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 B = one swizzle-128B span
constexpr int A_STAGE_BYTES = BM * 128;  // 16 KiB

struct GemmParams {
  int64_t M, N, K;
  const uint32_t* a_bits;  // A_BITS kernels: A is a 0/1 matrix, bit k of row m = a_bits[m * ld_words + k / 32] >> (k % 32)
  int64_t ld_words;
  int a_mn, b_mn;
  int kb_per_split, num_kb;
  sbr_gemm_epilogue_t ep;
};

constexpr int BITS_GROUP = 4;  // K blocks per TMA load of the bit-packed A operand (32 bytes per row)

template <int BN>
struct Cfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int LUT_BYTES = 4096;  // (comparison build -DSBR_GB_LUT=1: byte -> eight bf16 table in shared memory)
  // bit-packed A operand: ring of TMA-loaded groups of BITS_GROUP K blocks ([BM rows] x [8 bytes per K block])
  static constexpr int BITS_SLOTS = 2;
  static constexpr int BITS_SLOT_BYTES = BM * 8 * BITS_GROUP;
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 + LUT_BYTES + 1024;  // + barriers + align slack
  // (BN = 64: 111 872 bytes, so that two CTAs -- the user- and the item-side projection -- still share an SM)
  static constexpr int SMEM_BYTES_BITS = SMEM_BYTES + BITS_SLOTS * BITS_SLOT_BYTES;
};

// butterfly transpose-reduce: on return lane l holds sum over the warp's 32 lanes of v[l]
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

// activation with the (warp-uniform) switch hoisted out of the element loop
__device__ __forceinline__ void apply_act32(int act, float (&v)[32]) {
  switch (act) {
    case SBR_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    case SBR_ACT_TANH:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
      break;
    case SBR_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + __expf(-v[j]));
      break;
    case SBR_ACT_SELU:
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = 1.0507009873554805f * (v[j] > 0.f ? v[j] : 1.6732632423543772f * (__expf(v[j]) - 1.f));
      break;
    default: break;
  }
}
// v[j] *= act'(y[j]) expressed through the saved output y
__device__ __forceinline__ void apply_actgrad32(int act, float (&v)[32], const float (&y)[32]) {
  switch (act) {
    case SBR_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = y[j] > 0.f ? v[j] : 0.f;
      break;
    case SBR_ACT_TANH:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= 1.f - y[j] * y[j];
      break;
    case SBR_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= y[j] * (1.f - y[j]);
      break;
    case SBR_ACT_SELU:
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] *= y[j] > 0.f ? 1.0507009873554805f : y[j] + 1.0507009873554805f * 1.6732632423543772f;
      break;
    default: break;
  }
}

// A_BITS: the A operand is a bit-packed 0/1 matrix (multi-hot 'interactions' rows, data/Feature.py:147-150).  Four extra
// warps expand it into the bf16 SWIZZLE_128B K-major stage in shared memory (thread = tile row, 64 bits -> eight
// 16-byte chunks per K block, expanded in registers), so HBM sees 1 bit per element instead of 16 and the tensor cores
// see ordinary bf16 operands; the TMA producer loads B and, as 32-byte row segments, the bit words.
// epilogue warps: 4 (one per TMEM lane quarter) next to the bit converters, else 8 -- two warps per lane quarter, each
// draining every other 32-column chunk: skinny layers (K = N = 64) are bound by the epilogue's instruction stream
// timing diagnostics of the bit-packed GEMM (scripts/gemm_bits_variants.sh): each switch removes one part of the kernel
// (results are then wrong); the product build leaves all of them on
#ifndef SBR_GB_LOAD
#define SBR_GB_LOAD 1
#endif
#ifndef SBR_GB_STS
#define SBR_GB_STS 1
#endif
#ifndef SBR_GB_LUT
#define SBR_GB_LUT 0
#endif
#ifndef SBR_GB_FENCE
#define SBR_GB_FENCE 1
#endif
#ifndef SBR_GB_TMA
#define SBR_GB_TMA 1
#endif
#ifndef SBR_GB_EPI
#define SBR_GB_EPI 1
#endif

// per-CTA phase trace of the bit-packed GEMM (-DSBR_GB_TRACE, scripts/gemm_bits_trace.py): clock64 relative to the
// kernel entry at a handful of points, %globaltimer at entry and exit
#ifdef SBR_GB_TRACE
__device__ unsigned long long g_gb_trace[1024 * 16];
#define GB_T(slot) do { if (A_BITS && lane == 0 && cta_lin < 1024) g_gb_trace[cta_lin * 16 + (slot)] = (unsigned long long)(clock64() - t_entry); } while (0)
#define GB_T1(slot, cond) do { if (cond) GB_T(slot); } while (0)
#else
#define GB_T(slot) do { } while (0)
#define GB_T1(slot, cond) do { } while (0)
#endif

template <bool A_BITS>
struct Warps {
  static constexpr int EPI = A_BITS ? 4 : 8;
  static constexpr int THREADS = (2 + EPI + (A_BITS ? 4 : 0)) * 32;
};

template <int BN, bool A_BITS>
__global__ void __launch_bounds__(Warps<A_BITS>::THREADS)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  SBR_PDL_LAUNCH();  // (the wait follows the barrier / TMEM set-up: the prologue overlaps the previous kernel's tail)
#ifdef SBR_GB_TRACE
  const long long t_entry = clock64();
  const int cta_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (A_BITS && threadIdx.x == 0 && cta_lin < 1024) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_gb_trace[cta_lin * 16 + 0] = gt;
  }
#endif
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::STAGES * A_STAGE_BYTES;
  // A_BITS: [BITS_SLOTS][BM][8 * BITS_GROUP] bit words behind the operand ring (1024-byte aligned TMA destination)
  uint8_t* s_bits = sB + C::STAGES * C::B_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_bits + (A_BITS ? C::BITS_SLOTS * C::BITS_SLOT_BYTES : 0));
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint4* s_lut = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(full_bar) + 256);
  uint64_t* bits_full = reinterpret_cast<uint64_t*>(tmem_slot + 2);    // [BITS_SLOTS] group landed (TMA)
  uint64_t* bits_empty = bits_full + C::BITS_SLOTS;                    // [BITS_SLOTS] group read by the 4 converter warps

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (A_BITS && SBR_GB_LUT && threadIdx.x < 256) {
    const uint32_t t = threadIdx.x;
    const auto pair = [t](int b) { return ((t >> b) & 1u ? 0x3F80u : 0u) | ((t >> (b + 1)) & 1u ? 0x3F800000u : 0u); };
    s_lut[t] = make_uint4(pair(0), pair(2), pair(4), pair(6));
  }
  // persistent over the M tiles: CTA x handles tiles x, x + gridDim.x, ...; the two TMEM accumulators let the MMA of
  // tile t+1 run while the epilogue warps drain tile t (skinny layers are bound by the epilogue's HBM traffic)
  const int num_m_tiles = (int)((p.M + BM - 1) / BM);
  const int n0 = blockIdx.y * BN;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], A_BITS ? 5 : 1);  // TMA (+ the four converter warps)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], Warps<A_BITS>::EPI);
    }
    if (A_BITS) {
      for (int b = 0; b < C::BITS_SLOTS; ++b) {
        mbar_init(&bits_full[b], 1);
        mbar_init(&bits_empty[b], 4);
      }
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN < 32 ? 32 : 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  GB_T1(1, warp == 0);
  SBR_PDL_WAIT();  // nothing above reads or writes global memory
  GB_T1(2, warp == 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (converged warp, elected issue)
    int s = 0;
    uint32_t ph = 0;
    // A_BITS: the bit words of tile rows x BITS_GROUP K blocks arrive by TMA, one group ahead of the group whose B
    // tiles are being issued (sequence number bq over all (tile, group) pairs of this CTA; slot = bq % BITS_SLOTS).
    // The slot's previous group was read into registers by the converters before they filled the A stages this warp
    // has just seen released, so the wait below does not block.
    // (a TMA box has to start on a 16-byte boundary of the row: groups are counted from the even K block kb_al)
    const int kb_al = kb_begin & ~1;
    const int n_groups = (kb_end - kb_al + BITS_GROUP - 1) / BITS_GROUP;
    int bq = 0, b_tile = blockIdx.x, b_g = 0;
    auto issue_bits = [&]() {
      if (!SBR_GB_LOAD || b_tile >= num_m_tiles) return;
      const int slot = bq % C::BITS_SLOTS;
      mbar_wait(&bits_empty[slot], (uint32_t)(((bq / C::BITS_SLOTS) & 1) ^ 1));
      mbar_arrive_expect_tx(&bits_full[slot], C::BITS_SLOT_BYTES);
      tma_load_2d(s_bits + slot * C::BITS_SLOT_BYTES, &tmA, &bits_full[slot], (kb_al + b_g * BITS_GROUP) * 8,
                  b_tile * BM);
      ++bq;
      if (++b_g == n_groups) { b_g = 0; b_tile += gridDim.x; }
    };
    if (A_BITS && elect_one()) issue_bits();
    __syncwarp();
    for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x) {
      const int m0 = tile * BM;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one()) {
          if (A_BITS && (kb == kb_begin || (kb - kb_al) % BITS_GROUP == 0)) issue_bits();
          mbar_arrive_expect_tx(&full_bar[s], (A_BITS ? 0 : A_STAGE_BYTES) + (SBR_GB_TMA ? C::B_STAGE_BYTES : 0));
          uint8_t* a_dst = sA + s * A_STAGE_BYTES;
          uint8_t* b_dst = sB + s * C::B_STAGE_BYTES;
          if (A_BITS) {
            // A is written by the converter warps
          } else if (!p.a_mn) {
            tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[s], m0 + j * 64, kb * BK);
          }
          if (!SBR_GB_TMA) {
          } else if (!p.b_mn) {
            tma_load_2d(b_dst, &tmB, &full_bar[s], kb * BK, n0);
            GB_T1(11, kb == kb_begin);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[s], n0 + j * 64, kb * BK);
          }
        }
        __syncwarp();
        if (++s == C::STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected issue:
    // under `lane == 0` ptxas wraps every tcgen05.mma in a register->uniform-register broadcast loop)
    const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tempty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[s], ph);
        GB_T1(3, kb == kb_begin);
        GB_T1(4, kb == kb_begin + 4);
        GB_T1(13, kb == kb_begin + 8);
        GB_T1(14, kb == kb_begin + 12);
        tc_fence_after();
        const uint32_t a_addr = sA_addr + (uint32_t)(s * A_STAGE_BYTES);
        const uint32_t b_addr = sB_addr + (uint32_t)(s * C::B_STAGE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: +32 B per 16-element K step inside the 128 B swizzle span; MN-major: +2 K-groups of 1024 B
            const uint64_t adesc = p.a_mn ? umma_smem_desc(a_addr + k * 2048, 8192, 1024)
                                          : umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = p.b_mn ? umma_smem_desc(b_addr + k * 2048, 8192, 1024)
                                          : umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
        }
        __syncwarp();
        if (++s == C::STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull_bar[acc]);
      GB_T(5);
      __syncwarp();
    }
  } else if (A_BITS && warp >= 6) {
    // ------------------------------------------------------------------ bit -> bf16 converters (thread = tile row)
    const int row_in_tile = threadIdx.x - 192;
    const uint32_t sw = (uint32_t)(row_in_tile & 7);
    const uint32_t row_off = (uint32_t)((row_in_tile >> 3) * 1024 + (row_in_tile & 7) * 128);
    int s = 0;
    uint32_t ph = 0;
    int bq = 0;  // (tile, group) sequence number, as in the producer
    const uint4* my_bits = reinterpret_cast<const uint4*>(s_bits + row_in_tile * (8 * BITS_GROUP));
    for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x) {
      for (int kb0 = kb_begin & ~1; kb0 < kb_end; kb0 += BITS_GROUP, ++bq) {
        // this row's bit words of BITS_GROUP K blocks (rows past M and K blocks past the row pitch arrive as zeros)
        uint2 curw[BITS_GROUP];
        if (SBR_GB_LOAD) {
          const int slot = bq % C::BITS_SLOTS;
          mbar_wait(&bits_full[slot], (uint32_t)((bq / C::BITS_SLOTS) & 1));
          GB_T1(9, bq == 0 && warp == 6);
          const uint4* src = my_bits + slot * (C::BITS_SLOT_BYTES / 16);
#pragma unroll
          for (int q = 0; q < BITS_GROUP; q += 2) {
            const uint4 v = src[q / 2];
            curw[q] = make_uint2(v.x, v.y);
            curw[q + 1] = make_uint2(v.z, v.w);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bits_empty[slot]);
        } else {
#pragma unroll
          for (int q = 0; q < BITS_GROUP; ++q) curw[q] = make_uint2(0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < BITS_GROUP; ++q) {
          if (kb0 + q >= kb_begin && kb0 + q < kb_end) {  // 64 bits = one K block of this row
            const uint2 w = curw[q];
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* dst = sA + s * A_STAGE_BYTES + row_off;
#pragma unroll
            for (uint32_t c = 0; c < (SBR_GB_STS ? 8u : 0u); ++c) {  // chunk c = K elements 8c .. 8c+7 -> 16 bytes at the swizzled position
              const uint32_t byte = __byte_perm(c < 4 ? w.x : w.y, 0u, 0x4440u | (c & 3u));
              if (SBR_GB_LUT) {
                *reinterpret_cast<uint4*>(dst + ((c ^ sw) << 4)) = s_lut[byte];
              } else {
                // two bits -> two bf16 (0.0 | 1.0) in one word, in registers: bit 0 -> bit 7, bit 1 -> bit 23 (the two
                // shifted copies of the byte do not overlap, so the product has no carries), then x 0x7F turns each
                // single bit into the seven mantissa / exponent bits of 0x3F80.  The shared-memory pipe (table reads
                // with bank conflicts + stage stores + the tensor core's operand reads) bounded this kernel at
                // ~1000 cycles per K block (scripts/gemm_bits_trace.py).
                // (one multiply spreads all eight bits of the byte: bit j -> bits j + 7 and j + 22)
                const uint32_t y = byte * 0x00400080u;
                const auto two = [y](int i) { return ((y >> (2 * i)) & 0x00800080u) * 0x7Fu; };
                *reinterpret_cast<uint4*>(dst + ((c ^ sw) << 4)) = make_uint4(two(0), two(1), two(2), two(3));
              }
            }
            if (SBR_GB_FENCE) fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
            GB_T1(10, kb0 + q == kb_begin && warp == 6);
            if (++s == C::STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps, one TMEM lane quarter each)
    const int q = warp & 3;
    constexpr int NH = Warps<A_BITS>::EPI / 4;  // warps per lane quarter
    const int half = (warp - 2) >> 2;           // this warp drains the chunks c0 = 32 * (half + NH * i)
    const int row_in_tile = q * 32 + lane;
    const sbr_gemm_epilogue_t& ep = p.ep;
    // split-K without atomics: every K partition writes its own fp32 slice (reduced by sbr_splitk_reduce)
    float* out32 = ep.out_f32 != nullptr ? ep.out_f32 + (int64_t)blockIdx.z * ep.split_stride : nullptr;
    float cs_acc[BN / 32], cq_acc[BN / 32];  // lane l: partial column statistics of column 32 * i + l
#pragma unroll
    for (int i = 0; i < BN / 32; ++i) cs_acc[i] = cq_acc[i] = 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++it) {
    const int acc = it & 1;
    const int64_t row = (int64_t)tile * BM + row_in_tile;
    const bool row_ok = SBR_GB_EPI && row < p.M;
    mbar_wait(&tfull_bar[acc], (uint32_t)((it >> 1) & 1));
    GB_T1(6, warp == 2);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 32 * half; c0 < BN; c0 += 32 * NH) {
      const int64_t col0 = (int64_t)n0 + c0;
      if (col0 >= p.N) break;  // warp-uniform
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
      tmem_ld_wait();
      const bool full_cols = col0 + 32 <= p.N;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ep.alpha;
      if (ep.bias != nullptr) {
        if (full_cols && ((reinterpret_cast<uintptr_t>(ep.bias + col0) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {  // same address in every lane: one broadcast transaction each
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) v[j] += __ldg(ep.bias + col0 + j);
        }
      }
      if (ep.act != SBR_ACT_NONE) apply_act32(ep.act, v);
      if (ep.actgrad_y != nullptr) {
        float yv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) yv[j] = 1.f;  // rows / columns outside the matrix: derivative value irrelevant
        if (row_ok) {
          const bf16* y = reinterpret_cast<const bf16*>(ep.actgrad_y) + row * ep.ld_actgrad + col0;
          if (full_cols && (ep.ld_actgrad % 8 == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 u = __ldg(reinterpret_cast<const uint4*>(y + j));
              const bf16* yb = reinterpret_cast<const bf16*>(&u);
#pragma unroll
              for (int t = 0; t < 8; ++t) yv[j + t] = __bfloat162float(yb[t]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) yv[j] = __bfloat162float(y[j]);
          }
        }
        apply_actgrad32(ep.actgrad_act, v, yv);
      }
      if (ep.colstats != nullptr) {
        float s1[32], s2[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = row_ok ? v[j] : 0.f;
          s1[j] = x;
          s2[j] = x * x;
        }
        // accumulated over all tiles of this CTA; flushed with one atomic per column after the tile loop
        cs_acc[c0 / 32] += warp_colsum32(s1, lane);
        if (!ep.colstats_sum_only) cq_acc[c0 / 32] += warp_colsum32(s2, lane);
      }
      if (ep.transpose_out) {
        // out_f32 is [N, M]: for a fixed column the 32 lanes hit 32 consecutive floats
        if (out32 != nullptr && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (col0 + j < p.N) {
              float* dst = out32 + (col0 + j) * ep.ld_f32 + row;
              if (ep.atomic_out) atomicAdd(dst, v[j]);
              else *dst = v[j];
            }
          }
        }
        continue;
      }
      if (out32 != nullptr && row_ok) {
        float* dst = out32 + row * ep.ld_f32 + col0;
        if (ep.atomic_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) atomicAdd(dst + j, v[j]);
        } else if (full_cols && (ep.ld_f32 % 4 == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) dst[j] = v[j];
        }
      }
      if (ep.out_bf16 != nullptr && row_ok) {
        bf16* dst = reinterpret_cast<bf16*>(ep.out_bf16) + row * ep.ld_bf16 + col0;
        if (full_cols && (ep.ld_bf16 % 8 == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
            *reinterpret_cast<uint4*>(dst + j) = u;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) dst[j] = __float2bfloat16(v[j]);
        }
      }
    }
    GB_T1(7, warp == 2);
    // this warp is done with the accumulator
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }  // tile loop
    if (ep.colstats != nullptr) {
#pragma unroll
      for (int i = 0; i < BN / 32; ++i) {
        const int64_t col = (int64_t)n0 + 32 * i + lane;
        if (i % NH == half && col < p.N) {
          if (ep.colstats_rows > 0) {
            // deterministic statistics: this warp's private row of partial sums, added up in a fixed order by
            // sbr_bn_finalize (fp32 atomics would make the batch statistics -- and through bf16 roundings the whole
            // forward -- depend on the arrival order)
            float* rowp = ep.colstats + (size_t)(blockIdx.x * 4 + q) * 2 * p.N;
            rowp[col] = cs_acc[i];
            if (!ep.colstats_sum_only) rowp[p.N + col] = cq_acc[i];
          } else {
            atomicAdd(ep.colstats + col, cs_acc[i]);
            if (!ep.colstats_sum_only) atomicAdd(ep.colstats + p.N + col, cq_acc[i]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  GB_T1(8, warp == 0);
#ifdef SBR_GB_TRACE
  if (A_BITS && threadIdx.x == 0 && cta_lin < 1024) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_gb_trace[cta_lin * 16 + 12] = gt;
  }
#endif
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN < 32 ? 32 : 2 * BN);
  }
}

inline int64_t gemm_grid_x(int64_t M, int64_t N, int splits) {
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  const int64_t m_tiles = (M + BM - 1) / BM, n_tiles = (N + BN - 1) / BN;
  const int64_t ctas_per_sm = BN == 256 ? 1 : 2;  // 2 x BN TMEM columns and the operand ring per CTA
  int64_t gx = (sbr_num_sms() * ctas_per_sm) / (n_tiles * splits);
  return gx < 1 ? 1 : (gx > m_tiles ? m_tiles : gx);
}

template <int BN, bool A_BITS>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int splits, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    SBR_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, A_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        A_BITS ? Cfg<BN>::SMEM_BYTES_BITS : Cfg<BN>::SMEM_BYTES));
    configured = true;
  }
  const int64_t n_tiles = (p.N + BN - 1) / BN;
  const int64_t gx = gemm_grid_x(p.M, p.N, splits);
  SBR_REQUIRE(p.ep.colstats == nullptr || p.ep.colstats_rows == 0 || p.ep.colstats_rows >= 4 * gx,
              "sbr_gemm: colstats_rows=%d < %lld partial rows", p.ep.colstats_rows, (long long)(4 * gx));
  dim3 grid((unsigned)gx, (unsigned)n_tiles, (unsigned)splits);
  SBR_CHECK_CUDA(sbr_launch(gemm_bf16_kernel<BN, A_BITS>, grid, dim3(Warps<A_BITS>::THREADS),
                            (size_t)(A_BITS ? Cfg<BN>::SMEM_BYTES_BITS : Cfg<BN>::SMEM_BYTES), st,
                            tmA, tmB, p));
  return SBR_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int sbr_make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                          uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode_fn();
  SBR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  SBR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand base must be 16-byte aligned");
  SBR_REQUIRE((ld_elems * 2) % 16 == 0, "TMA operand row pitch must be a multiple of 8 bf16 elements (got %llu)",
              (unsigned long long)ld_elems);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SBR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu ld=%llu)",
              (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld_elems);
  return SBR_OK;
}

// bit-packed A operand as a byte matrix [M, ld_bytes]: box = BM rows x (8 bytes per K block x BITS_GROUP), no swizzle,
// rows / bytes outside the matrix arrive as zeros
static int sbr_make_tmap_bits_2d(CUtensorMap* out, const void* base, uint64_t ld_bytes, uint64_t rows) {
  PFN_encodeTiled enc = get_encode_fn();
  SBR_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
  cuuint64_t dims[2] = {ld_bytes, rows};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {8 * BITS_GROUP, BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SBR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (bit matrix) failed with CUresult %d (rows=%llu ld_bytes=%llu)",
              (int)r, (unsigned long long)rows, (unsigned long long)ld_bytes);
  return SBR_OK;
}

extern "C" int sbr_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                             int64_t M, int64_t N, int64_t K, const sbr_gemm_epilogue_t* ep, void* stream) {
  SBR_REQUIRE(A && B && ep, "sbr_gemm_bf16: null operand");
  SBR_REQUIRE(M > 0 && N > 0 && K > 0, "sbr_gemm_bf16: empty problem M=%lld N=%lld K=%lld", (long long)M,
              (long long)N, (long long)K);
  SBR_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "sbr_gemm_bf16: dimension too large");
  SBR_REQUIRE(ep->out_bf16 || ep->out_f32 || ep->colstats, "sbr_gemm_bf16: no output requested");
  SBR_REQUIRE(!(ep->transpose_out && ep->out_bf16), "sbr_gemm_bf16: transposed output is fp32 only");
  int splits = ep->split_k < 1 ? 1 : ep->split_k;
  const int num_kb = (int)((K + BK - 1) / BK);
  if (splits > num_kb) splits = num_kb;
  SBR_REQUIRE(splits == 1 || ((ep->atomic_out || ep->split_stride > 0) && !ep->out_bf16 && !ep->colstats &&
                              !ep->bias && ep->act == SBR_ACT_NONE && !ep->actgrad_y),
              "sbr_gemm_bf16: split_k > 1 needs a pure fp32 epilogue (atomic, or one slice per partition)");
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn_major) rc = sbr_make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  else rc = sbr_make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  if (rc) return rc;
  if (!b_mn_major) rc = sbr_make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)BN);
  else rc = sbr_make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return rc;

  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = a_mn_major ? 1 : 0;
  p.b_mn = b_mn_major ? 1 : 0;
  p.num_kb = num_kb;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  p.ep = *ep;
  if (p.ep.alpha == 0.f) p.ep.alpha = 1.f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  p.a_bits = nullptr;
  p.ld_words = 0;
  switch (BN) {
    case 64: return launch_gemm<64, false>(tmA, tmB, p, splits, st);
    case 128: return launch_gemm<128, false>(tmA, tmB, p, splits, st);
    default: return launch_gemm<256, false>(tmA, tmB, p, splits, st);
  }
}

#ifdef SBR_GB_TRACE
extern "C" int sbr_debug_gemm_bits_trace(unsigned long long* out_host) {
  return (int)cudaMemcpyFromSymbol(out_host, g_gb_trace, sizeof(g_gb_trace));
}
#endif

extern "C" int sbr_gemm_colstats_rows(int64_t M, int64_t N) { return (int)(4 * gemm_grid_x(M, N, 1)); }

extern "C" int sbr_gemm_bits_bf16(const uint32_t* A_bits, int64_t ld_words, const void* B, int64_t ldb, int b_mn_major,
                                  int64_t M, int64_t N, int64_t K, const sbr_gemm_epilogue_t* ep, void* stream) {
  SBR_REQUIRE(A_bits && B && ep, "sbr_gemm_bits_bf16: null operand");
  SBR_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31),
              "sbr_gemm_bits_bf16: bad problem M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
  const int num_kb = (int)((K + BK - 1) / BK);
  // (the bit words reach shared memory by TMA: 16-byte aligned rows)
  SBR_REQUIRE(ld_words % 4 == 0 && ld_words >= 2 * (int64_t)num_kb &&
                  (reinterpret_cast<uintptr_t>(A_bits) & 15) == 0,
              "sbr_gemm_bits_bf16: rows must be 16-byte aligned and padded to whole 64-bit K blocks (ld_words=%lld)",
              (long long)ld_words);
  SBR_REQUIRE(ep->out_bf16 || ep->out_f32 || ep->colstats, "sbr_gemm_bits_bf16: no output requested");
  SBR_REQUIRE(!(ep->transpose_out && ep->out_bf16), "sbr_gemm_bits_bf16: transposed output is fp32 only");
  int splits = ep->split_k < 1 ? 1 : ep->split_k;
  if (splits > num_kb) splits = num_kb;
  SBR_REQUIRE(splits == 1 || ((ep->atomic_out || ep->split_stride > 0) && !ep->out_bf16 && !ep->colstats &&
                              !ep->bias && ep->act == SBR_ACT_NONE && !ep->actgrad_y),
              "sbr_gemm_bits_bf16: split_k > 1 needs a pure fp32 epilogue (atomic, or one slice per partition)");
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  CUtensorMap tmBits, tmB;
  int rc = sbr_make_tmap_bits_2d(&tmBits, A_bits, (uint64_t)ld_words * 4, (uint64_t)M);
  if (rc) return rc;
  if (!b_mn_major) rc = sbr_make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, (uint32_t)BN);
  else rc = sbr_make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.a_bits = A_bits;
  p.ld_words = ld_words;
  p.a_mn = 0;
  p.b_mn = b_mn_major ? 1 : 0;
  p.num_kb = num_kb;
  p.kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.ep = *ep;
  if (p.ep.alpha == 0.f) p.ep.alpha = 1.f;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (BN) {
    case 64: return launch_gemm<64, true>(tmBits, tmB, p, splits, st);
    case 128: return launch_gemm<128, true>(tmBits, tmB, p, splits, st);
    default: return launch_gemm<256, true>(tmBits, tmB, p, splits, st);
  }
}
