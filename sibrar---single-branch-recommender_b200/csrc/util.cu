// sibrar_b200 -- casts, CSR densify, modality sampling, step counter, multi-tensor Adam(W), negative sampling.
// Reference code replaced by each kernel is named in include/sibrar_b200.h.
#include <stdarg.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------ error plumbing
static thread_local char g_err[1024] = "";
void sbr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* sbr_last_error(void) { return g_err; }
extern "C" int sbr_version(void) { return 100; }


namespace {
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------ casts
__global__ void cast_kernel(const float* __restrict__ src, int64_t ld_src, bf16* __restrict__ dst, int64_t ld_dst,
                            int64_t rows, int64_t cols) {
  int64_t total = rows * ld_dst;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / ld_dst, c = i - r * ld_dst;
    dst[i] = __float2bfloat16(c < cols ? src[r * ld_src + c] : 0.f);
  }
}

__device__ __forceinline__ void store_as(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_as(bf16* p, float v) { *p = __float2bfloat16(v); }

template <typename OutT>
__global__ void transpose_cast_kernel(const float* __restrict__ src, int64_t ld_src, OutT* __restrict__ dst,
                                      int64_t ld_dst, int64_t rows, int64_t cols) {
  __shared__ float tile[32][33];
  int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) store_as(dst + c * ld_dst + r, tile[threadIdx.x][j]);
  }
}

__global__ void csr_to_dense_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                    const float* __restrict__ vals, int64_t rows, int64_t cols, bf16* __restrict__ dst,
                                    int64_t ld_dst) {
  int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  int lane = threadIdx.x & 31;
  bf16* d = dst + row * ld_dst;
  const bf16 zero = __float2bfloat16(0.f), one = __float2bfloat16(1.f);
  for (int64_t c = lane; c < ld_dst; c += 32) d[c] = zero;
  __syncwarp();
  for (int64_t p = indptr[row] + lane; p < indptr[row + 1]; p += 32) {
    int32_t c = indices[p];
    if (c >= 0 && c < cols) d[c] = vals ? __float2bfloat16(vals[p]) : one;
  }
}

// ------------------------------------------------------------------------------------------------ modality sampling
__global__ void sample_modalities_kernel(uint8_t* __restrict__ mods, int64_t n_rows, int k, int n_mods, int central,
                                         uint64_t seed, const int64_t* __restrict__ step_dev) {
  SBR_PDL_ENTRY();
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  uint64_t step = (uint64_t)*step_dev;
  uint4 rnd = philox4x32(make_uint4((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)step, 0x6d6f6473u),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32)));
  if (k == 1) {
    mods[r] = (uint8_t)min(n_mods - 1, (int)(u32_to_unit(rnd.x) * n_mods));
  } else if (central >= 0) {
    int o = min(n_mods - 2, (int)(u32_to_unit(rnd.x) * (n_mods - 1)));
    if (o >= central) ++o;
    mods[r * 2] = (uint8_t)central;
    mods[r * 2 + 1] = (uint8_t)o;
  } else {
    int a = min(n_mods - 1, (int)(u32_to_unit(rnd.x) * n_mods));
    int b = min(n_mods - 2, (int)(u32_to_unit(rnd.y) * (n_mods - 1)));
    if (b >= a) ++b;
    mods[r * 2] = (uint8_t)a;
    mods[r * 2 + 1] = (uint8_t)b;
  }
}

__global__ void tick_kernel(int64_t* c) {
  SBR_PDL_ENTRY();
  *c += 1;
}

// start of a train step in one launch: bump the step counters and clear the step's accumulator arena
__global__ void step_begin_kernel(int64_t* c0, int64_t* c1, uint4* zero, int64_t n16) {
  SBR_PDL_ENTRY();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (c0) *c0 += 1;
    if (c1) *c1 += 1;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
    zero[i] = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------------------------------------ Adam / AdamW
constexpr int ADAM_CHUNK = 1024;  // (ops.ADAM_CHUNK builds the chunk table)
__device__ __forceinline__ void adam_update(float& p, float& g, float& m, float& v, float grad_scale, float lr, float wd,
                                            int decoupled, float beta1, float beta2, float step_size, float bc2_sqrt,
                                            float eps) {
  g *= grad_scale;
  if (decoupled == 2) {
    // torch.optim.Adagrad (train/trainer.py:62-66; lr_decay = 0, initial accumulator 0): coupled weight decay,
    // v = running sum of squared gradients, no first moment
    g += wd * p;
    v += g * g;
    p -= lr * g / (sqrtf(v) + eps);
    return;
  }
  if (decoupled) p *= (1.f - lr * wd);
  else g += wd * p;
  m = beta1 * m + (1.f - beta1) * g;
  v = beta2 * v + (1.f - beta2) * g * g;
  p -= step_size * m / (sqrtf(v) / bc2_sqrt + eps);
}

// one chunk [off, off + ADAM_CHUNK) of tensor t; g_src = where the gradient is READ (t.grad, or the all-reduced copy of
// the data-parallel step); t.grad itself is cleared either way (zero_grad)
__device__ __forceinline__ void adam_chunk(const sbr_adam_tensor_t& t, int64_t off, const float* __restrict__ g_src,
                                           float lr, float beta1, float beta2, float eps, float wd, int decoupled,
                                           float grad_scale, float step_size, float bc2_sqrt) {
  const int64_t end = min(off + (int64_t)ADAM_CHUNK, t.numel);
  const auto al = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  const bool vec = (t.numel & 3) == 0 && (off & 3) == 0 && al(t.param, 16) && al(t.grad, 16) && al(g_src, 16) &&
                   al(t.exp_avg, 16) && al(t.exp_avg_sq, 16) &&
                   (t.shadow_bf16 == nullptr || ((t.cols & 3) == 0 && (t.shadow_ld & 3) == 0 && al(t.shadow_bf16, 8)));
  if (vec) {
    for (int64_t i = off + 4 * (int64_t)threadIdx.x; i < end; i += 4 * (int64_t)blockDim.x) {
      float4 p4 = *reinterpret_cast<const float4*>(t.param + i), g4 = *reinterpret_cast<const float4*>(g_src + i);
      float4 m4 = *reinterpret_cast<const float4*>(t.exp_avg + i), v4 = *reinterpret_cast<const float4*>(t.exp_avg_sq + i);
      adam_update(p4.x, g4.x, m4.x, v4.x, grad_scale, lr, wd, decoupled, beta1, beta2, step_size, bc2_sqrt, eps);
      adam_update(p4.y, g4.y, m4.y, v4.y, grad_scale, lr, wd, decoupled, beta1, beta2, step_size, bc2_sqrt, eps);
      adam_update(p4.z, g4.z, m4.z, v4.z, grad_scale, lr, wd, decoupled, beta1, beta2, step_size, bc2_sqrt, eps);
      adam_update(p4.w, g4.w, m4.w, v4.w, grad_scale, lr, wd, decoupled, beta1, beta2, step_size, bc2_sqrt, eps);
      *reinterpret_cast<float4*>(t.param + i) = p4;
      *reinterpret_cast<float4*>(t.exp_avg + i) = m4;
      *reinterpret_cast<float4*>(t.exp_avg_sq + i) = v4;
      *reinterpret_cast<float4*>(t.grad + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t.shadow_bf16) {
        const int64_t r = i / t.cols, c = i - r * t.cols;  // 4 consecutive elements stay inside one row (cols % 4 == 0)
        uint2 o;
        *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(p4.x, p4.y);
        *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(p4.z, p4.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(t.shadow_bf16) + r * t.shadow_ld + c) = o;
      }
    }
    return;
  }
  for (int64_t i = off + threadIdx.x; i < end; i += blockDim.x) {
    float p = t.param[i], g = g_src[i], m = t.exp_avg[i], v = t.exp_avg_sq[i];
    adam_update(p, g, m, v, grad_scale, lr, wd, decoupled, beta1, beta2, step_size, bc2_sqrt, eps);
    t.param[i] = p;
    t.exp_avg[i] = m;
    t.exp_avg_sq[i] = v;
    t.grad[i] = 0.f;
    if (t.shadow_bf16) {
      int64_t r = i / t.cols, c = i - r * t.cols;
      reinterpret_cast<bf16*>(t.shadow_bf16)[r * t.shadow_ld + c] = __float2bfloat16(p);
    }
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(const sbr_adam_tensor_t* __restrict__ tensors, const int32_t* __restrict__ chunk_to_tensor,
            const int64_t* __restrict__ chunk_offset, float lr, float beta1, float beta2, float eps, float wd,
            int decoupled, const int64_t* __restrict__ step_dev, float grad_scale) {
  SBR_PDL_ENTRY();
  // bias corrections: double-precision pow once per block (every thread doing it cost more than the update itself)
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {
    const double step = (double)*step_dev;
    s_bc[0] = (float)(1.0 - pow((double)beta1, step));
    s_bc[1] = (float)sqrt(1.0 - pow((double)beta2, step));
  }
  const sbr_adam_tensor_t t = tensors[chunk_to_tensor[blockIdx.x]];
  const int64_t off = chunk_offset[blockIdx.x];
  __syncthreads();
  adam_chunk(t, off, t.grad, lr, beta1, beta2, eps, wd, decoupled, grad_scale, lr / s_bc[0], s_bc[1]);
}

// ------------------------------------------------------------------------------------------------ data-parallel Adam
// Gradient all-reduce + optimizer in ONE kernel over NVSwitch multicast memory (replaces the NCCL all-reduce nodes in
// front of the optimizer of the data-parallel step).  Every rank runs the same persistent grid:
//   barrier A  (all ranks' gradients are complete)
//   phase 1    rank r sums slice r of the flat gradient buffer over all ranks INSIDE the switch
//              (multimem.ld_reduce on the multicast address) and broadcasts the sums into every rank's `sum` buffer
//              (multimem.st) -- each NVLink carries the buffer once in each direction
//   barrier B  (every slice has landed everywhere)
//   phase 2    multi-tensor Adam / AdamW / Adagrad on the local replica, reading the summed gradients, clearing the
//              local accumulators, refreshing the bf16 shadows.
// Cross-GPU barriers are epoch flags in symmetric memory (rank r writes epoch e into word [slot][r] of every peer and
// waits until its own words [slot][*] reach e; flags only grow, so CUDA-graph replays need no reset); block 0 talks to
// the peers, the other blocks of the grid follow a local release word.
__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu64(int64_t* p, int64_t v) {
  asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int64_t ld_acquire_gpu64(const int64_t* p) {
  int64_t v;
  asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// a peer that never arrives (crashed rank) must not hang the GPU: after ~10 s of polling the kernel traps
struct spin_guard {
  long long t0 = -1;
  unsigned n = 0;
  __device__ __forceinline__ void poll() {
#ifdef __CUDA_ARCH__
    if ((++n & 0x3ffu) == 0) {
      const long long now = clock64();
      if (t0 < 0) t0 = now;
      if (now - t0 > 20000000000ll) asm volatile("trap;");
    }
#endif
  }
};

// block 0: handshake with every peer on `slot`, then release the local grid; other blocks: wait for the release
__device__ __forceinline__ void mc_barrier(const sbr_mc_comm_t& c, int slot, int32_t epoch) {
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < c.world) {
      __threadfence_system();
      st_release_sys(c.peer_flags_dev[threadIdx.x] + slot * c.world + c.rank, epoch);
      spin_guard sg;
      while (ld_acquire_sys(c.flags_local + slot * c.world + threadIdx.x) - epoch < 0) sg.poll();
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu64(c.state + 1 + slot, (int64_t)epoch);
  } else {
    if (threadIdx.x == 0) {
      spin_guard sg;
      while (ld_acquire_gpu64(c.state + 1 + slot) - (int64_t)epoch < 0) sg.poll();
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
adam_mc_kernel(const sbr_adam_tensor_t* __restrict__ tensors, int64_t total_chunks,
               const int32_t* __restrict__ chunk_to_tensor, const int64_t* __restrict__ chunk_offset, float lr,
               float beta1, float beta2, float eps, float wd, int decoupled, const int64_t* __restrict__ step_dev,
               float grad_scale, int apply_adam, sbr_mc_comm_t c) {
  SBR_PDL_ENTRY();
  __shared__ float s_bc[2];
  const int32_t epoch = (int32_t)(ld_acquire_gpu64(c.state) + 1);  // (read by every block before the grid barrier below)
  if (threadIdx.x == 0) {
    const double step = (double)*step_dev;
    s_bc[0] = (float)(1.0 - pow((double)beta1, step));
    s_bc[1] = (float)sqrt(1.0 - pow((double)beta2, step));
  }
  mc_barrier(c, 0, epoch);
  // ---- phase 1: my slice, reduced in the switch, broadcast to every rank
  const int64_t slice = c.total / c.world;  // floats, a multiple of 4
  const int64_t lo = slice * c.rank;
  for (int64_t i = 4 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x); i < slice; i += 4 * (int64_t)gridDim.x * blockDim.x) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(c.mc_grads + lo + i)
                 : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c.sum_mc + lo + i), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
  // ---- grid barrier (all of this rank's stores issued), then barrier B over the ranks
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(reinterpret_cast<unsigned long long*>(c.state + 3), 1ull);
    if (blockIdx.x == 0) {
      const int64_t want = (int64_t)epoch * gridDim.x;  // (the grid size is fixed for the lifetime of `state`)
      spin_guard sg;
      while (ld_acquire_gpu64(c.state + 3) - want < 0) sg.poll();
      st_release_gpu64(c.state, (int64_t)epoch);  // every block has read the previous epoch by now
    }
  }
  __syncthreads();
  mc_barrier(c, 1, epoch);
  if (!apply_adam) return;
  // ---- phase 2: the optimizer on the summed gradients (grad_scale carries 1 / world)
  const float bc2_sqrt = s_bc[1], step_size = lr / s_bc[0];
  for (int64_t ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    const sbr_adam_tensor_t t = tensors[chunk_to_tensor[ch]];
    const float* g_src = c.sum_local + (t.grad - c.flat_grads);
    adam_chunk(t, chunk_offset[ch], g_src, lr, beta1, beta2, eps, wd, decoupled, grad_scale, step_size, bc2_sqrt);
  }
}

// ------------------------------------------------------------------------------------------------ negative sampling
__global__ void sample_batch_kernel(const int32_t* __restrict__ coo_user, const int32_t* __restrict__ coo_item,
                                    int64_t nnz, const int64_t* __restrict__ indptr,
                                    const int32_t* __restrict__ indices, const int32_t* __restrict__ items_in_split,
                                    int64_t n_items_in_split, int64_t B, int n_neg, uint64_t seed,
                                    const int64_t* __restrict__ step_dev, int64_t* __restrict__ out_u,
                                    int64_t* __restrict__ out_i, const int64_t* __restrict__ order, int64_t offset,
                                    int strategy, const int32_t* __restrict__ item_pos) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= B * (n_neg + 1)) return;
  const int64_t b = t / (n_neg + 1);
  const int j = (int)(t - b * (n_neg + 1));
  const uint64_t step = (uint64_t)*step_dev;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x6e656773u);
  int64_t e;
  if (order != nullptr) {
    e = order[offset + b];  // epoch order: slot b is the (offset + b)-th interaction of the shuffled epoch
  } else {
    // the positive (and thereby the user) of slot b: same draw for every j of the slot
    uint4 r0 = philox4x32(make_uint4((uint32_t)b, (uint32_t)(b >> 32), (uint32_t)step, (uint32_t)(step >> 32)), key);
    uint64_t r64 = ((uint64_t)r0.x << 32) | r0.y;
    e = (int64_t)__umul64hi(r64, (uint64_t)nnz);
  }
  const int32_t u = coo_user[e];
  if (j == 0) {
    out_u[b] = u;
    out_i[b * (n_neg + 1)] = coo_item[e];
    return;
  }
  const int64_t beg = indptr[u], end = indptr[u + 1];
  if (strategy == 1) {
    // 'uniform' (data/sampling.py:7-32): n_neg DISTINCT items of the split that are not train positives of the user --
    // raw = n_neg distinct draws from [0, M), M = n_choices - n_pos (np.random.choice(..., replace=False)), shifted past
    // the positives: neg = raw + #{j : pos_j - j <= raw} with pos_j the sorted positions of the positives among the
    // choices (searchsorted(pos - arange, raw, 'right')).  Thread j == 1 draws the whole slot (Floyd's subset sampling:
    // a uniformly random n_neg-subset; the order inside the row does not enter any loss).
    if (j != 1) return;
    const int64_t M = n_items_in_split - (end - beg);
    int64_t chosen[64];
    const int nn = n_neg < 64 ? n_neg : 64;
    for (int i = 0; i < nn; ++i) {
      const int64_t hi = M - nn + i;  // draw from [0, hi]
      uint4 r = philox4x32(make_uint4((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)step, 0x40000000u + i), key);
      int64_t v = (int64_t)__umul64hi(((uint64_t)r.x << 32) | r.y, (uint64_t)(hi + 1));
      bool dup = false;
      for (int q = 0; q < i; ++q) dup |= chosen[q] == v;
      chosen[i] = dup ? hi : v;
    }
    for (int i = 0; i < nn; ++i) {
      const int64_t raw = chosen[i];
      int64_t lo = 0, hi = end - beg;  // first positive with (position - rank) > raw
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int64_t pj = (item_pos ? (int64_t)item_pos[indices[beg + mid]] : (int64_t)indices[beg + mid]) - mid;
        if (pj <= raw) lo = mid + 1;
        else hi = mid;
      }
      out_i[b * (n_neg + 1) + 1 + i] = items_in_split[raw + lo];
    }
    return;
  }
  int32_t cand = 0;
  int64_t pos = 0;
  bool found = false;
  const auto is_positive = [&](int32_t c) {  // binary search in the sorted positives of u
    int64_t lo = beg, hi = end;
    while (lo < hi) {
      int64_t mid = (lo + hi) >> 1;
      if (indices[mid] < c) lo = mid + 1;
      else hi = mid;
    }
    return lo < end && indices[lo] == c;
  };
  for (int attempt = 0; attempt < 64 && !found; ++attempt) {
    uint4 r = philox4x32(make_uint4((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)step, 0x80000000u + attempt), key);
    uint64_t q = ((uint64_t)r.x << 32) | r.y;
    pos = (int64_t)__umul64hi(q, (uint64_t)n_items_in_split);
    cand = items_in_split[pos];
    found = !is_positive(cand);
  }
  // 64 rejected draws (a user who interacted with almost every item of the split): the reference keeps re-drawing
  // (data/dataloader.py:180-191); here the walk continues from the last draw to the next item that is NOT a positive,
  // so a positive is never emitted as a negative (the host checked that such an item exists)
  for (int64_t probe = 1; !found && probe < n_items_in_split; ++probe) {
    int64_t q = pos + probe;
    if (q >= n_items_in_split) q -= n_items_in_split;
    cand = items_in_split[q];
    found = !is_positive(cand);
  }
  out_i[t] = cand;
}

// out[r, c] (+)= act(sum_s part[s][r, c] + bias[c]).  Block = 32 consecutive elements x 8 partition lanes: every
// load instruction of a warp reads 128 contiguous bytes of one slice, lane group ty sums slices ty, ty + 8, ...
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, int n_splits, int64_t split_stride, int64_t ld_part, int64_t rows,
                     int64_t cols, const float* __restrict__ bias, int act, float* __restrict__ out_f32,
                     int64_t ld_f32, int accumulate, bf16* __restrict__ out_bf16, int64_t ld_bf16) {
  SBR_PDL_ENTRY();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t total = rows * cols;
  for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
    const int64_t i = base + tx;
    const int64_t r = i < total ? i / cols : 0, c = i < total ? i - r * cols : 0;
    float a0 = 0.f, a1 = 0.f;
    if (i < total) {
      const float* p = part + r * ld_part + c;
      int s = ty;
      for (; s + 8 < n_splits; s += 16) {
        a0 += p[(int64_t)s * split_stride];
        a1 += p[(int64_t)(s + 8) * split_stride];
      }
      if (s < n_splits) a0 += p[(int64_t)s * split_stride];
    }
    red[ty][tx] = a0 + a1;
    __syncthreads();
    if (ty == 0 && i < total) {
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) v += red[q][tx];
      if (bias) v += bias[c];
      v = act_fwd(act, v);
      if (out_f32) {
        float* d = out_f32 + r * ld_f32 + c;
        *d = accumulate ? *d + v : v;
      }
      if (out_bf16) out_bf16[r * ld_bf16 + c] = __float2bfloat16(v);
    }
    __syncthreads();
  }
}

// few partitions (the forward split-K of an 'interactions' table: 6-10 slices): a thread owns 4 consecutive elements
// and walks the slices with independent 16-byte loads -- no shared memory, no barriers
__global__ void __launch_bounds__(256)
splitk_reduce_vec_kernel(const float* __restrict__ part, int n_splits, int64_t split_stride, int64_t ld_part,
                         int64_t rows, int64_t cols, const float* __restrict__ bias, int act,
                         float* __restrict__ out_f32, int64_t ld_f32, int accumulate, bf16* __restrict__ out_bf16,
                         int64_t ld_bf16) {
  SBR_PDL_ENTRY();
  const int64_t c4n = cols >> 2, total4 = rows * c4n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4n, c = (i - r * c4n) * 4;
    const float* p = part + r * ld_part + c;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    int s = 0;
    for (; s + 1 < n_splits; s += 2) {
      const float4 x = __ldcs(reinterpret_cast<const float4*>(p + (int64_t)s * split_stride));
      const float4 y = __ldcs(reinterpret_cast<const float4*>(p + (int64_t)(s + 1) * split_stride));
      a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
      b.x += y.x; b.y += y.y; b.z += y.z; b.w += y.w;
    }
    if (s < n_splits) {
      const float4 x = __ldcs(reinterpret_cast<const float4*>(p + (int64_t)s * split_stride));
      a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
    }
    float v[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
    if (bias) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias + c);
      v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = act_fwd(act, v[j]);
    if (out_f32) {
      float4* d = reinterpret_cast<float4*>(out_f32 + r * ld_f32 + c);
      if (accumulate) {
        const float4 o = *d;
        v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
      }
      *d = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (out_bf16) {
      uint2 o;
      *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(v[0], v[1]);
      *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(v[2], v[3]);
      *reinterpret_cast<uint2*>(out_bf16 + r * ld_bf16 + c) = o;
    }
  }
}

}  // namespace

extern "C" int sbr_splitk_reduce(const float* partials, int n_splits, int64_t split_stride, int64_t ld_part,
                                 int64_t rows, int64_t cols, const float* bias, int act, float* out_f32,
                                 int64_t ld_f32, int accumulate, void* out_bf16, int64_t ld_bf16, void* stream) {
  SBR_REQUIRE(partials && n_splits >= 1 && rows > 0 && cols > 0 && (out_f32 || out_bf16),
              "sbr_splitk_reduce: bad arguments");
  const int64_t total = rows * cols;
  const auto al = [](const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  if (n_splits <= 32 && (cols & 3) == 0 && (ld_part & 3) == 0 && (split_stride & 3) == 0 && al(partials, 16) &&
      (!bias || al(bias, 16)) && (!out_f32 || ((ld_f32 & 3) == 0 && al(out_f32, 16))) &&
      (!out_bf16 || ((ld_bf16 & 3) == 0 && al(out_bf16, 8)))) {
    const unsigned blocks = (unsigned)min((int64_t)sbr_num_sms() * 8, (total / 4 + 255) / 256);
    SBR_CHECK_CUDA(sbr_launch(splitk_reduce_vec_kernel, dim3(blocks), dim3(256), (size_t)(0), S(stream), partials,
                              n_splits, split_stride, ld_part, rows, cols, bias, act, out_f32, ld_f32, accumulate,
                              reinterpret_cast<bf16*>(out_bf16), ld_bf16));
    SBR_LAUNCH_CHECK();
    return SBR_OK;
  }
  const unsigned blocks = (unsigned)min((int64_t)sbr_num_sms() * 16, (total + 31) / 32);
  SBR_CHECK_CUDA(sbr_launch(splitk_reduce_kernel, dim3(blocks), dim3(256), (size_t)(0), S(stream), partials, n_splits, split_stride, ld_part, rows, cols, bias, act,
                                                      out_f32, ld_f32, accumulate,
                                                      reinterpret_cast<bf16*>(out_bf16), ld_bf16));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_cast_f32_to_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows,
                                    int64_t cols, void* stream) {
  SBR_REQUIRE(src && dst && rows >= 0 && cols >= 0 && ld_dst >= cols, "sbr_cast_f32_to_bf16: bad arguments");
  if (rows * ld_dst == 0) return SBR_OK;
  unsigned blocks = (unsigned)min((int64_t)sbr_num_sms() * 16, (rows * ld_dst + 255) / 256);
  cast_kernel<<<blocks, 256, 0, S(stream)>>>(src, ld_src, reinterpret_cast<bf16*>(dst), ld_dst, rows, cols);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_transpose_f32_to_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows,
                                         int64_t cols, void* stream) {
  SBR_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_dst >= rows, "sbr_transpose_f32_to_bf16: bad arguments");
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32));
  transpose_cast_kernel<bf16><<<grid, dim3(32, 8), 0, S(stream)>>>(src, ld_src, reinterpret_cast<bf16*>(dst), ld_dst,
                                                                   rows, cols);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_transpose_f32(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows,
                                 int64_t cols, void* stream) {
  SBR_REQUIRE(src && dst && rows > 0 && cols > 0 && ld_dst >= rows, "sbr_transpose_f32: bad arguments");
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32));
  transpose_cast_kernel<float><<<grid, dim3(32, 8), 0, S(stream)>>>(src, ld_src, dst, ld_dst, rows, cols);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_csr_to_dense_bf16(const int64_t* indptr, const int32_t* indices, const float* vals, int64_t rows,
                                     int64_t cols, void* dst, int64_t ld_dst, void* stream) {
  SBR_REQUIRE(indptr && dst && rows > 0 && ld_dst >= cols, "sbr_csr_to_dense_bf16: bad arguments");
  csr_to_dense_kernel<<<cdiv(rows, 8), 256, 0, S(stream)>>>(indptr, indices, vals, rows, cols,
                                                            reinterpret_cast<bf16*>(dst), ld_dst);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_sample_modalities(uint8_t* mods, int64_t n_rows, int k, int n_mods, int central, uint64_t seed,
                                     const int64_t* step_dev, void* stream) {
  SBR_REQUIRE(mods && step_dev && n_rows > 0, "sbr_sample_modalities: bad arguments");
  SBR_REQUIRE(k == 1 || k == 2, "sbr_sample_modalities: k must be 1 or 2 (got %d)", k);
  SBR_REQUIRE(n_mods >= k && n_mods <= 255, "sbr_sample_modalities: need k <= n_mods <= 255 (k=%d n_mods=%d)", k,
              n_mods);
  SBR_REQUIRE(central < n_mods, "sbr_sample_modalities: central modality out of range");
  SBR_CHECK_CUDA(sbr_launch(sample_modalities_kernel, dim3(cdiv(n_rows, 256)), dim3(256), (size_t)(0), S(stream), mods, n_rows, k, n_mods, central, seed, step_dev));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_step_begin(int64_t* counter0, int64_t* counter1, void* zero, int64_t zero_bytes, void* stream) {
  SBR_REQUIRE(zero_bytes >= 0 && (zero_bytes & 15) == 0 && (zero_bytes == 0 || zero != nullptr) &&
                  (reinterpret_cast<uintptr_t>(zero) & 15) == 0,
              "sbr_step_begin: the cleared range must be 16-byte aligned and sized");
  const int64_t n16 = zero_bytes / 16;
  int64_t blocks = (n16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > sbr_num_sms()) blocks = sbr_num_sms();
  SBR_CHECK_CUDA(sbr_launch(step_begin_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, S(stream), counter0, counter1,
                            reinterpret_cast<uint4*>(zero), n16));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_tick(int64_t* counter_dev, void* stream) {
  SBR_REQUIRE(counter_dev, "sbr_tick: null counter");
  SBR_CHECK_CUDA(sbr_launch(tick_kernel, dim3(1), dim3(1), (size_t)(0), S(stream), counter_dev));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_adam_step(const sbr_adam_tensor_t* tensors_dev, int n_tensors, int64_t total_chunks,
                             const int32_t* chunk_to_tensor_dev, const int64_t* chunk_offset_dev, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int decoupled, const int64_t* step_dev,
                             float grad_scale, void* stream) {
  SBR_REQUIRE(tensors_dev && chunk_to_tensor_dev && chunk_offset_dev && step_dev && n_tensors > 0 && total_chunks > 0,
              "sbr_adam_step: bad arguments");
  SBR_CHECK_CUDA(sbr_launch(adam_kernel, dim3((unsigned)total_chunks), dim3(256), (size_t)(0), S(stream), tensors_dev, chunk_to_tensor_dev, chunk_offset_dev, lr,
                                                             beta1, beta2, eps, weight_decay, decoupled, step_dev,
                                                             grad_scale));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_adam_step_mc(const sbr_adam_tensor_t* tensors_dev, int n_tensors, int64_t total_chunks,
                                const int32_t* chunk_to_tensor_dev, const int64_t* chunk_offset_dev, float lr,
                                float beta1, float beta2, float eps, float weight_decay, int decoupled,
                                const int64_t* step_dev, float grad_scale, int apply_adam, const sbr_mc_comm_t* comm,
                                int grid_blocks, void* stream) {
  SBR_REQUIRE(tensors_dev && chunk_to_tensor_dev && chunk_offset_dev && step_dev && n_tensors > 0 && total_chunks > 0,
              "sbr_adam_step_mc: bad arguments");
  SBR_REQUIRE(comm && comm->flat_grads && comm->mc_grads && comm->sum_local && comm->sum_mc && comm->peer_flags_dev &&
                  comm->flags_local && comm->state && comm->world >= 1 && comm->rank >= 0 && comm->rank < comm->world,
              "sbr_adam_step_mc: bad communicator");
  SBR_REQUIRE(comm->world <= 256 && comm->total > 0 && comm->total % (4 * (int64_t)comm->world) == 0,
              "sbr_adam_step_mc: total=%lld must be a multiple of 4 * world", (long long)comm->total);
  // persistent grid, every block resident (the blocks wait for each other): at most 2 per SM
  SBR_REQUIRE(grid_blocks >= 1 && grid_blocks <= 2 * sbr_num_sms(), "sbr_adam_step_mc: grid_blocks=%d not in [1, %d]",
              grid_blocks, 2 * sbr_num_sms());
  SBR_CHECK_CUDA(sbr_launch(adam_mc_kernel, dim3((unsigned)grid_blocks), dim3(256), (size_t)0, S(stream), tensors_dev,
                            total_chunks, chunk_to_tensor_dev, chunk_offset_dev, lr, beta1, beta2, eps, weight_decay,
                            decoupled, step_dev, grad_scale, apply_adam, *comm));
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_sample_batch(const int32_t* coo_user, const int32_t* coo_item, int64_t nnz,
                                const int64_t* train_indptr, const int32_t* train_indices,
                                const int32_t* items_in_split, int64_t n_items_in_split, int64_t B, int n_neg,
                                uint64_t seed, const int64_t* step_dev, int64_t* out_u, int64_t* out_i, void* stream) {
  SBR_REQUIRE(coo_user && coo_item && train_indptr && train_indices && items_in_split && out_u && out_i && step_dev,
              "sbr_sample_batch: null argument");
  SBR_REQUIRE(nnz > 0 && n_items_in_split > 0 && B > 0 && n_neg >= 0, "sbr_sample_batch: bad sizes");
  sample_batch_kernel<<<cdiv(B * (n_neg + 1), 256), 256, 0, S(stream)>>>(coo_user, coo_item, nnz, train_indptr,
                                                                        train_indices, items_in_split, n_items_in_split,
                                                                        B, n_neg, seed, step_dev, out_u, out_i,
                                                                        (const int64_t*)nullptr, 0, 0,
                                                                        (const int32_t*)nullptr);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_sample_epoch_batch(const int32_t* coo_user, const int32_t* coo_item, int64_t nnz,
                                      const int64_t* order, int64_t offset, const int64_t* train_indptr,
                                      const int32_t* train_indices, const int32_t* items_in_split,
                                      int64_t n_items_in_split, int64_t B, int n_neg, uint64_t seed,
                                      const int64_t* step_dev, int64_t* out_u, int64_t* out_i, void* stream) {
  SBR_REQUIRE(coo_user && coo_item && order && train_indptr && train_indices && items_in_split && out_u && out_i &&
                  step_dev,
              "sbr_sample_epoch_batch: null argument");
  SBR_REQUIRE(nnz > 0 && n_items_in_split > 0 && B > 0 && n_neg >= 0 && offset >= 0 && offset + B <= nnz,
              "sbr_sample_epoch_batch: bad sizes (offset=%lld B=%lld nnz=%lld)", (long long)offset, (long long)B,
              (long long)nnz);
  sample_batch_kernel<<<cdiv(B * (n_neg + 1), 256), 256, 0, S(stream)>>>(coo_user, coo_item, nnz, train_indptr,
                                                                        train_indices, items_in_split, n_items_in_split,
                                                                        B, n_neg, seed, step_dev, out_u, out_i, order,
                                                                        offset, 0, (const int32_t*)nullptr);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}

extern "C" int sbr_sample_negatives(const int32_t* coo_user, const int32_t* coo_item, int64_t nnz, const int64_t* order,
                                    int64_t offset, const int64_t* train_indptr, const int32_t* train_indices,
                                    const int32_t* items_in_split, int64_t n_items_in_split, const int32_t* item_pos,
                                    int64_t B, int n_neg, int strategy, uint64_t seed, const int64_t* step_dev,
                                    int64_t* out_u, int64_t* out_i, void* stream) {
  SBR_REQUIRE(coo_user && coo_item && train_indptr && train_indices && items_in_split && out_u && out_i && step_dev,
              "sbr_sample_negatives: null argument");
  SBR_REQUIRE(nnz > 0 && n_items_in_split > 0 && B > 0 && n_neg >= 0 && offset >= 0 &&
                  (order == nullptr || offset + B <= nnz),
              "sbr_sample_negatives: bad sizes");
  SBR_REQUIRE(strategy == SBR_NEG_UNIFORM_RECBOLE || strategy == SBR_NEG_UNIFORM,
              "sbr_sample_negatives: Sampling strategy %d not yet supported.", strategy);
  SBR_REQUIRE(strategy != SBR_NEG_UNIFORM || n_neg <= 64, "sbr_sample_negatives: 'uniform' draws at most 64 negatives");
  sample_batch_kernel<<<cdiv(B * (n_neg + 1), 256), 256, 0, S(stream)>>>(coo_user, coo_item, nnz, train_indptr,
                                                                        train_indices, items_in_split, n_items_in_split,
                                                                        B, n_neg, seed, step_dev, out_u, out_i, order,
                                                                        offset, strategy, item_pos);
  SBR_LAUNCH_CHECK();
  return SBR_OK;
}


// ------------------------------------------------------------------------------------------------ timeline build
#ifdef SBR_STAMPS
static int (*g_tu_setters[32])(unsigned long long*);
static int g_n_tu = 0;
void sbr_register_tu(int (*setter)(unsigned long long*)) {
  if (g_n_tu < 32) g_tu_setters[g_n_tu++] = setter;
}
// buf: device uint64 [1 + 2 * 4000] (word 0 = number of stamps, zeroed by the caller), or NULL to switch stamping off
extern "C" int sbr_debug_stamps(unsigned long long* buf) {
  for (int i = 0; i < g_n_tu; ++i)
    if (g_tu_setters[i](buf) != 0) return 1;
  return 0;
}
#endif
