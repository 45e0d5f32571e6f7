// sibrar_b200 -- shared device/host helpers for the sm_100a kernels (PTX wrappers: mbarrier, TMA, tcgen05/TMEM).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sibrar_b200.h"

// ------------------------------------------------------------------------------------------------ errors
void sbr_set_error(const char* fmt, ...);

#define SBR_CHECK_CUDA(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      sbr_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));         \
      return SBR_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define SBR_REQUIRE(cond, ...)                                                                     \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      sbr_set_error(__VA_ARGS__);                                                                  \
      return SBR_ERR_ARG;                                                                          \
    }                                                                                              \
  } while (0)

#define SBR_LAUNCH_CHECK() SBR_CHECK_CUDA(cudaGetLastError())

static inline int sbr_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------ launches
// Programmatic dependent launch (default on, SBR_PDL=0 disables): every kernel of the train step starts with SBR_PDL_ENTRY() -- it lets the NEXT kernel
// of the stream be scheduled as soon as all CTAs of this one are running (its prologue / block dispatch overlaps our
// tail) and then waits until the PREVIOUS kernel has completed and flushed its memory.  Nothing before the wait may
// touch global memory.  Without the launch attribute (plain <<<>>> launches) both instructions are no-ops.
// Timeline build only (-DSBR_STAMPS, `make stamps` -> libsibrar_b200_stamps.so, scripts/step_timeline.py): thread 0 of
// block 0 of every kernel records %globaltimer right behind its griddepcontrol.wait -- the moment its inputs are ready
// inside the replayed graph -- tagged with (hash of the source file, line of the macro).  The production library
// contains none of this.
#ifdef SBR_STAMPS
static __device__ unsigned long long* g_sbr_stamps;  // [0] = count, then (time ns, tag) pairs; one copy per translation unit
__host__ __device__ constexpr unsigned sbr_fid(const char* s) {
  unsigned h = 2166136261u;
  for (; *s; ++s) h = (h ^ (unsigned)*s) * 16777619u;
  return h & 0xffffu;
}
__device__ __forceinline__ void sbr_stamp(unsigned fid, int line) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 && threadIdx.y == 0) {
    unsigned long long* b = g_sbr_stamps;
    if (b != nullptr) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      const unsigned long long i = atomicAdd(b, 1ull);
      if (i < 4000) {
        b[1 + 2 * i] = t;
        b[2 + 2 * i] = ((unsigned long long)(gridDim.x * gridDim.y * gridDim.z) << 32) | ((unsigned long long)fid << 16) |
                       (unsigned long long)(line & 0xffff);
      }
    }
  }
}
void sbr_register_tu(int (*setter)(unsigned long long*));  // util.cu
static int sbr_tu_set_stamps(unsigned long long* p) {
  return (int)cudaMemcpyToSymbol(g_sbr_stamps, &p, sizeof(p));
}
namespace {
struct SbrTuReg {
  SbrTuReg() { sbr_register_tu(sbr_tu_set_stamps); }
};
static SbrTuReg sbr_tu_reg;
}  // namespace
#define SBR_STAMP() sbr_stamp(sbr_fid(__FILE__), __LINE__)
#else
#define SBR_STAMP() ((void)0)
#endif

#define SBR_PDL_ENTRY()                                                   \
  do {                                                                    \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");       \
    asm volatile("griddepcontrol.wait;" ::: "memory");                    \
    SBR_STAMP();                                                          \
  } while (0)

// the two halves, for kernels with a real prologue (barrier / TMEM set-up, shared-memory clears): the prologue runs
// while the previous kernel drains; only what follows SBR_PDL_WAIT() may touch memory the previous kernel wrote
#define SBR_PDL_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define SBR_PDL_WAIT()                                                    \
  do {                                                                    \
    asm volatile("griddepcontrol.wait;" ::: "memory");                    \
    SBR_STAMP();                                                          \
  } while (0)

static inline int sbr_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SBR_PDL");
    // on by default (SBR_PDL=0 disables): 3.6 % faster on the multi-branch step graph (0.339 -> 0.327 ms); it was 3 %
    // slower on the earlier single-stream graph
    v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return v;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t sbr_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  memset(&attr, 0, sizeof(attr));
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = sbr_pdl_enabled();
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------------ small device utils
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
// One lane of a CONVERGED warp.  Single-thread tcgen05 / TMA issue must sit under this (not under `lane == 0`):
// the warp then stays convergent, operands live in uniform registers and ptxas emits no broadcast loops.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// activations (ids shared with the host: SBR_ACT_*); derivative expressed through the OUTPUT y
__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case SBR_ACT_RELU: return fmaxf(x, 0.f);
    case SBR_ACT_TANH: return tanhf(x);
    case SBR_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    case SBR_ACT_SELU: return 1.0507009873554805f * (x > 0.f ? x : 1.6732632423543772f * (__expf(x) - 1.f));
    default: return x;
  }
}
__device__ __forceinline__ float act_grad_from_out(int act, float y) {
  switch (act) {
    case SBR_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case SBR_ACT_TANH: return 1.f - y * y;
    case SBR_ACT_SIGMOID: return y * (1.f - y);
    case SBR_ACT_SELU: return y > 0.f ? 1.0507009873554805f : y + 1.0507009873554805f * 1.6732632423543772f;
    default: return 1.f;
  }
}

// Philox4x32-10 (counter-based; one call -> 4 x 32 random bits)
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float u32_to_unit(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ------------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ------------------------------------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates {c0 = innermost (contiguous) element index, c1 = row index}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ------------------------------------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::f16 : bf16/f16 inputs, fp32 accumulate); one thread issues
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// lane = TMEM lane (row of D), 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// lane = TMEM lane, 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: lane = TMEM lane, 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: A is [128 lanes x K] bf16, two elements per 32-bit column (K-major only)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): SWIZZLE_128B, version 1.
//   K-major : rows of 128 B (64 bf16 of K); 8-row groups SBO = 1024 B apart; LBO unused.
//   MN-major: K-rows of 128 B (64 bf16 of MN); 8-K-row groups SBO = 1024 B apart; next 64 MN elements LBO bytes apart.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;  // layout_type = SWIZZLE_128B
  return d;
}
// Instruction descriptor (InstrDescriptor): bf16 x bf16 -> fp32, dense
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_mn_major & 1) << 15) | ((uint32_t)(b_mn_major & 1) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ host: tensor maps
// 2-D bf16 tensor map, 128-byte swizzle.  dims/strides in elements; inner dimension is contiguous.
int sbr_make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                          uint32_t box_inner, uint32_t box_outer);
