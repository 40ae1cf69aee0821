// One-shot tensor-parallel all-reduce over NVLink peer memory, fused with the residual add + RMSNorm
// that follows it.
//
// Replaces, for decode-sized messages, the reference's capturable NCCL wrapper
//   PyNCCLCommunicator.all_reduce  (python/minisgl/kernel/csrc/src/pynccl.cu:93-134: D2D copy into
//   the symmetric window, ncclAllReduce, D2D copy back), called once per layer after o_proj
//   (python/minisgl/layers/linear.py:102-106) and after the MLP down projection (linear.py:122-126),
// and optionally the flashinfer.fused_add_rmsnorm that consumes its result
//   (python/minisgl/layers/norm.py:32-38 via models/qwen3.py:38-41).
//
// Design ("push" one-shot, one kernel, one synchronisation phase):
//   every rank owns a region  data[2 parities][world senders][slot] | flags[world][kMaxCtas] | ctr
//   that all peers have mapped (CUDA IPC).  CTA b of rank s
//     1. PUSHES the rows it owns (r = b, b + grid, ...) from x -- any device memory, no staging copy --
//        into slot[parity][s] of EVERY rank with 16-byte stores over NVLink (posted writes, no round trip),
//     2. fences (system scope) and releases flag[s][b] = epoch on every rank,
//     3. acquires flag[0..world)[b] >= epoch in its own region (only its counterpart CTAs on the peers:
//        no grid-wide barrier), then
//     4. sums the world slots of its rows in rank order in fp32 (every rank gets bit-identical
//        results), rounds once to the 16-bit type (= the tensor the reference's all-reduce returns) and
//        either stores it, or applies  residual <- round(y + residual); out <- rmsnorm(y + residual) * w
//        with the arithmetic of rmsnorm_row_kernel<.., kFusedAdd> (elementwise.cu), bit for bit.
//   The epoch is a device-resident counter (advanced by the last CTA of each launch), so the kernel is
//   CUDA-graph capturable and replays correctly; parities alternate per launch, which is enough because
//   a rank can only enter launch e+1 after every peer has pushed launch e, i.e. has finished launch e-1.
//   Waits are bounded (globaltimer): a dead peer traps the kernel instead of hanging the GPU.
//
// Bytes per launch and rank: (world) x rows x dim x 2 pushed over NVLink (incl. the local copy),
// world x rows x dim x 2 read back locally; latency-bound for the <= 512 KB messages of decode.
#include "b200attn.h"
#include "common.cuh"

#include <cstring>
#include <type_traits>

namespace b200 {
namespace ar {

constexpr int kMaxWorld = 8;
constexpr int kMaxCtas = 512;  // one CTA per row up to 512 rows: the reduce phase is one row deep

struct Comm {
  int rank, world;
  size_t slot_bytes;   // capacity of one sender slot (max message bytes, 256-aligned)
  size_t flags_off, ctr_off, region_bytes;
  uint8_t* base[kMaxWorld];  // base[rank] = local region, others = IPC mappings
  void* opened[kMaxWorld];   // mappings this process opened (closed in destroy)
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

template <typename T>
struct Params {
  const T* x;
  int64_t x_rs;
  T* out;
  int64_t out_rs;
  T* residual;
  int64_t res_rs;
  const T* w;
  int rows, dim;
  float eps;
  int rank, world;
  size_t slot_bytes, flags_off, ctr_off;
  uint8_t* base[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ Vec8 ld_cg(const void* p) {
  Vec8 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3])
               : "l"(p));
  return r;
}

// block-wide sum with the reduction order of elementwise.cu::block_sum (bit-identical norms)
__device__ __forceinline__ float block_sum(float v, float* red /* [33] */) {
  const int lane = threadIdx.x % kWarp, wid = threadIdx.x / kWarp;
  const int nw = (blockDim.x + kWarp - 1) / kWarp;
  v = warp_sum(v);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

template <typename T, int kIter, bool kNorm>
__global__ void __launch_bounds__(1024) allreduce_push_kernel(const Params<T> p) {
  __shared__ float red[33];
  __shared__ uint32_t s_epoch;
  const int b = blockIdx.x, grid = gridDim.x, tid = threadIdx.x;
  const int chunks = p.dim / 8;
  uint8_t* mine = p.base[p.rank];
  uint32_t* ctr = reinterpret_cast<uint32_t*>(mine + p.ctr_off);
  // programmatic dependent launch: x is the predecessor's output (and the epoch counter was advanced by the
  // previous all-reduce, complete by transitivity); the successor may begin its own prologue right away
  pdl_wait();
  pdl_launch_dependents();
  if (tid == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(ctr) + 1u;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const size_t parity_off = (size_t)(epoch & 1u) * p.world * p.slot_bytes;

  // ---- 1. push my rows into slot[parity][rank] of every rank (peers first, own copy last)
  for (int r = b; r < p.rows; r += grid) {
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      const int c = tid + it * blockDim.x;
      if (c < chunks) {
        const Vec8 v = *reinterpret_cast<const Vec8*>(p.x + (int64_t)r * p.x_rs + c * 8);
        const size_t off = parity_off + (size_t)p.rank * p.slot_bytes + ((size_t)r * p.dim + c * 8) * sizeof(T);
        for (int i = 1; i <= p.world; ++i) {
          const int t = (p.rank + i) % p.world;
          *reinterpret_cast<Vec8*>(p.base[t] + off) = v;
        }
      }
    }
  }
  __syncthreads();
  // ---- 2. release: my pushes are visible system-wide before the flag is
  if (tid < p.world) {
    __threadfence_system();
    uint32_t* f = reinterpret_cast<uint32_t*>(p.base[tid] + p.flags_off) + p.rank * kMaxCtas + b;
    st_release_sys(f, epoch);
    // ---- 3. acquire the flag of sender `tid` for CTA b in my own region
    const uint32_t* g = reinterpret_cast<const uint32_t*>(mine + p.flags_off) + tid * kMaxCtas + b;
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(g) - epoch) < 0) {
      if (globaltimer_ns() - t0 > 8000000000ull) __trap();  // 8 s: a peer died
    }
  }
  __syncthreads();

  // ---- 4. reduce in rank order, epilogue
  for (int r = b; r < p.rows; r += grid) {
    float f[kIter][8];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      const int c = tid + it * blockDim.x;
      if (c < chunks) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        const size_t off = parity_off + ((size_t)r * p.dim + c * 8) * sizeof(T);
        for (int s = 0; s < p.world; ++s) {
          const Vec8 v = ld_cg(mine + off + (size_t)s * p.slot_bytes);
          float g[8];
          unpack8<T>(v, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += g[i];
        }
        // y = the all-reduced tensor, rounded once to the 16-bit type like the reference's result
        const Vec8 y = pack8<T>(acc);
        if constexpr (!kNorm) {
          *reinterpret_cast<Vec8*>(p.out + (int64_t)r * p.out_rs + c * 8) = y;
        } else {
          unpack8<T>(y, f[it]);
          const Vec8 rv = *reinterpret_cast<const Vec8*>(p.residual + (int64_t)r * p.res_rs + c * 8);
          float g[8];
          unpack8<T>(rv, g);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[it][i] += g[i];
          *reinterpret_cast<Vec8*>(p.residual + (int64_t)r * p.res_rs + c * 8) = pack8<T>(f[it]);
#pragma unroll
          for (int i = 0; i < 8; ++i) ss += f[it][i] * f[it][i];
        }
      }
    }
    if constexpr (kNorm) {
      const float tot = block_sum(ss, red);
      const float rcp = rsqrtf(tot / (float)p.dim + p.eps);
#pragma unroll
      for (int it = 0; it < kIter; ++it) {
        const int c = tid + it * blockDim.x;
        if (c < chunks) {
          const Vec8 wv = *reinterpret_cast<const Vec8*>(p.w + c * 8);
          float wf[8];
          unpack8<T>(wv, wf);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[it][i] = f[it][i] * rcp * wf[i];
          *reinterpret_cast<Vec8*>(p.out + (int64_t)r * p.out_rs + c * 8) = pack8<T>(f[it]);
        }
      }
      __syncthreads();  // red[] is reused by the next row
    }
  }

  // ---- the last CTA of the launch advances the epoch (every CTA has read it: all have arrived)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(ctr + 1, 1u) == (uint32_t)grid - 1u) {
      ctr[1] = 0u;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(ctr) = epoch;
    }
  }
}

template <typename T>
static int launch(const Comm* c, const void* x, int64_t x_rs, void* out, int64_t out_rs, void* residual,
                  int64_t res_rs, const void* w, int64_t rows, int dim, float eps, cudaStream_t st) {
  Params<T> p;
  p.x = (const T*)x;
  p.x_rs = x_rs;
  p.out = (T*)out;
  p.out_rs = out_rs;
  p.residual = (T*)residual;
  p.res_rs = res_rs;
  p.w = (const T*)w;
  p.rows = (int)rows;
  p.dim = dim;
  p.eps = eps;
  p.rank = c->rank;
  p.world = c->world;
  p.slot_bytes = c->slot_bytes;
  p.flags_off = c->flags_off;
  p.ctr_off = c->ctr_off;
  for (int i = 0; i < kMaxWorld; ++i) p.base[i] = c->base[i];
  // thread <-> chunk mapping of launch_fused_add_rmsnorm (elementwise.cu): identical reductions
  const int chunks = dim / 8;
  int threads = (int)ceil_div(chunks, 32) * 32;
  if (threads > 1024) threads = 1024;
  if (chunks > 512) threads = (int)ceil_div(ceil_div(chunks, 2), 32) * 32;
  const int iters = ceil_div(chunks, threads);
  B200_CHECK_ARG(iters <= 2, "allreduce: dim %d too large (max 16384)", dim);
  const int grid = rows < kMaxCtas ? (int)rows : kMaxCtas;
  const bool norm = residual != nullptr;
#define L(IT_, N_) B200_CHECK_CUDA(launch_pdl(allreduce_push_kernel<T, IT_, N_>, dim3(grid), dim3(threads), 0, st, p))
  if (iters == 1) {
    if (norm) L(1, true);
    else L(1, false);
  } else {
    if (norm) L(2, true);
    else L(2, false);
  }
#undef L
  B200_POST_LAUNCH();
  return 0;
}

}  // namespace ar
}  // namespace b200

using namespace b200;

extern "C" size_t b200_ar_region_bytes(int world, size_t max_bytes) {
  const size_t slot = ar::align_up(max_bytes, 256);
  return 2 * (size_t)world * slot + ar::align_up((size_t)world * ar::kMaxCtas * 4, 256) + 256;
}

extern "C" int b200_ar_alloc(size_t bytes, void** ptr) {
  B200_CHECK_ARG(ptr != nullptr && bytes > 0, "ar_alloc: bad arguments");
  B200_CHECK_CUDA(cudaMalloc(ptr, bytes));
  B200_CHECK_CUDA(cudaMemset(*ptr, 0, bytes));
  B200_CHECK_CUDA(cudaDeviceSynchronize());
  return 0;
}

extern "C" int b200_ar_ipc_handle(void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  B200_CHECK_CUDA(cudaIpcGetMemHandle(&h, ptr));
  std::memcpy(handle64, &h, 64);
  return 0;
}

extern "C" int b200_ar_ipc_open(const void* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  B200_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int b200_ar_create(int rank, int world, void* const* bases, const int* opened_mask,
                              size_t max_bytes, void** comm) {
  B200_CHECK_ARG(world >= 1 && world <= ar::kMaxWorld && rank >= 0 && rank < world,
                 "ar_create: bad rank/world %d/%d (world <= %d)", rank, world, ar::kMaxWorld);
  B200_CHECK_ARG(bases != nullptr && comm != nullptr && max_bytes > 0, "ar_create: bad arguments");
  auto* c = new ar::Comm();
  c->rank = rank;
  c->world = world;
  c->slot_bytes = ar::align_up(max_bytes, 256);
  c->flags_off = 2 * (size_t)world * c->slot_bytes;
  c->ctr_off = c->flags_off + ar::align_up((size_t)world * ar::kMaxCtas * 4, 256);
  c->region_bytes = b200_ar_region_bytes(world, max_bytes);
  for (int i = 0; i < ar::kMaxWorld; ++i) {
    c->base[i] = i < world ? static_cast<uint8_t*>(bases[i]) : nullptr;
    c->opened[i] = (i < world && opened_mask != nullptr && opened_mask[i]) ? bases[i] : nullptr;
  }
  *comm = c;
  return 0;
}

extern "C" int b200_ar_destroy(void* comm, int free_local) {
  if (comm == nullptr) return 0;
  auto* c = static_cast<ar::Comm*>(comm);
  for (int i = 0; i < c->world; ++i)
    if (c->opened[i] != nullptr) cudaIpcCloseMemHandle(c->opened[i]);
  if (free_local && c->base[c->rank] != nullptr) cudaFree(c->base[c->rank]);
  delete c;
  return 0;
}

extern "C" size_t b200_ar_max_bytes(const void* comm) {
  return comm == nullptr ? 0 : static_cast<const ar::Comm*>(comm)->slot_bytes;
}

extern "C" int b200_ar_allreduce(void* comm, const void* x, int64_t x_row_stride, void* out,
                                 int64_t out_row_stride, void* residual, int64_t res_row_stride,
                                 const void* weight, int64_t rows, int dim, float eps, int dtype,
                                 void* stream) {
  B200_CHECK_ARG(comm != nullptr, "allreduce: communicator is NULL");
  auto* c = static_cast<ar::Comm*>(comm);
  if (rows == 0) return 0;
  B200_CHECK_ARG(rows > 0 && dim > 0 && dim % 8 == 0, "allreduce: bad rows/dim %lld/%d", (long long)rows, dim);
  B200_CHECK_ARG((size_t)rows * dim * 2 <= c->slot_bytes, "allreduce: message of %lld x %d exceeds the %zu-byte slot",
                 (long long)rows, dim, c->slot_bytes);
  B200_CHECK_ARG((residual == nullptr) == (weight == nullptr), "allreduce: residual and weight go together");
  B200_CHECK_ARG(x_row_stride % 8 == 0 && out_row_stride % 8 == 0 && res_row_stride % 8 == 0 &&
                     ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)residual % 16) == 0 &&
                     ((uintptr_t)weight % 16) == 0,
                 "allreduce: pointers / row strides must be 16-byte aligned");
  auto st = (cudaStream_t)stream;
  if (dtype == B200_DTYPE_BF16)
    return ar::launch<__nv_bfloat16>(c, x, x_row_stride, out, out_row_stride, residual, res_row_stride, weight, rows, dim, eps, st);
  if (dtype == B200_DTYPE_FP16)
    return ar::launch<__half>(c, x, x_row_stride, out, out_row_stride, residual, res_row_stride, weight, rows, dim, eps, st);
  set_error("allreduce: bad dtype %d", dtype);
  return 1;
}
