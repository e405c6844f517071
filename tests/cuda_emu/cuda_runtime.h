// CPU stand-in for the CUDA constructs the emulation-ready sources use. TEST INFRASTRUCTURE ONLY
// (tests/test_cuda_emulation.py, tests/test_emulated_library.py); never part of the product, which
// has no CPU path.
//   kernels   a CTA runs as `blockDim.x` cooperative fibers, __syncthreads is a barrier among them,
//             a warp is 32 consecutive threads with their own barrier (__shfl_xor_sync,
//             __syncwarp), dynamic shared memory is a per-block heap buffer (NaN-filled), static
//             __shared__ a function static; blocks run one after another, launches are synchronous
//   runtime   the few cudaXxx calls of the host side: "device" memory is host memory (0xFF-filled,
//             so that an unset double reads as NaN and an unset index as -1), streams and events do
//             nothing but keep time
#pragma once
#include <fcntl.h>
#include <math.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <type_traits>
#include <memory>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static
#define __constant__ static

struct uint3
{
  unsigned x, y, z;
};
struct dim3
{
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1)
    : x(a)
    , y(b)
    , z(c)
  {}
};
struct uint2
{
  unsigned x, y;
};
struct int2
{
  int x, y;
};
struct alignas(16) double2
{
  double x, y;
};
struct alignas(8) float2
{
  float x, y;
};
inline uint2   make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline int2    make_int2(int x, int y) { return {x, y}; }
inline double2 make_double2(double x, double y) { return {x, y}; }
inline float2  make_float2(float x, float y) { return {x, y}; }

namespace gf_emu
{
  // A CTA runs as `blockDim.x` FIBERS on the calling thread (cooperative user-level contexts with
  // a ~20-instruction switch): deterministic, and a barrier costs a few switches instead of a
  // futex round trip per thread (OS threads made a 300-iteration CG on 130 nodes take 90 s).
  struct Bar
  {
    unsigned expected = 0, arrived = 0, gen = 0;
  };
  struct Warp
  {
    Bar bar;
    int n = 0;
    alignas(8) unsigned char slot[2][32][8];
  };
  struct Fiber
  {
    void *   sp   = nullptr;
    bool     done = false;
    uint3    thread_idx{0, 0, 0};
    Warp *   warp   = nullptr;
    unsigned parity = 0; // double-buffered shuffle slots: one barrier phase per shuffle
  };
  struct BlockState
  {
    uint3          block_idx{0, 0, 0};
    dim3           block_dim, grid_dim;
    Bar            bar;
    unsigned char *smem = nullptr;
  };
  inline Fiber *                      cur = nullptr;
  inline BlockState                   blk;
  inline void *                       sched_sp = nullptr;
  inline const std::function<void()> *cur_kernel = nullptr;
  inline long                         n_launches = 0;
  inline void *                       dynamic_smem() { return blk.smem; }

  // save the callee-saved registers and the stack pointer of the running context in *save_sp,
  // continue the context whose stack pointer is load_sp (System V x86-64)
  __attribute__((naked, noinline, used)) static void switch_context(void ** /*save_sp*/,
                                                                     void * /*load_sp*/)
  {
    asm volatile("pushq %rbp\n\t"
                 "pushq %rbx\n\t"
                 "pushq %r12\n\t"
                 "pushq %r13\n\t"
                 "pushq %r14\n\t"
                 "pushq %r15\n\t"
                 "movq %rsp, (%rdi)\n\t"
                 "movq %rsi, %rsp\n\t"
                 "popq %r15\n\t"
                 "popq %r14\n\t"
                 "popq %r13\n\t"
                 "popq %r12\n\t"
                 "popq %rbx\n\t"
                 "popq %rbp\n\t"
                 "ret\n\t");
  }
  inline void yield() { switch_context(&cur->sp, sched_sp); }
  inline void bar_wait(Bar &b)
  {
    const unsigned g = b.gen;
    if (++b.arrived >= b.expected)
      {
        b.arrived = 0;
        ++b.gen;
        return;
      }
    while (b.gen == g)
      yield();
  }
  inline void bar_drop(Bar &b) // a finished thread no longer takes part in barriers
  {
    --b.expected;
    if (b.expected > 0 && b.arrived >= b.expected)
      {
        b.arrived = 0;
        ++b.gen;
      }
  }
  static void fiber_entry()
  {
    (*cur_kernel)();
    bar_drop(cur->warp->bar);
    bar_drop(blk.bar);
    cur->done = true;
    for (;;)
      yield(); // never resumed
  }
  constexpr size_t FIBER_STACK = size_t(512) << 10;
  inline unsigned char *stack_of(unsigned t)
  {
    static std::vector<unsigned char *> stacks;
    while (stacks.size() <= t)
      stacks.push_back(static_cast<unsigned char *>(std::aligned_alloc(64, FIBER_STACK)));
    return stacks[t];
  }

  inline int schedule_mode()
  {
    const char *e = std::getenv("GF_EMU_SCHEDULE");
    if (e == nullptr)
      return 0;
    return std::strncmp(e, "reverse", 7) == 0 ? 1 : (std::strncmp(e, "random", 6) == 0 ? 2 : 0);
  }
  inline unsigned long long schedule_seed()
  {
    const char *e = std::getenv("GF_EMU_SCHEDULE");
    const char *c = e ? std::strchr(e, ':') : nullptr;
    return c ? std::strtoull(c + 1, nullptr, 10) + 0x9E3779B97F4A7C15ull : 0x9E3779B97F4A7C15ull;
  }

  // kernel<<<grid, block, smem_bytes>>>(args...) -> launch(grid, block, smem_bytes, [&] { kernel(args...); })
  inline void launch(dim3 grid3, unsigned block, size_t smem_bytes,
                     const std::function<void()> &kernel)
  {
    const unsigned grid = grid3.x * grid3.y * grid3.z; // blocks run one after another, x fastest
    ++n_launches;
    const unsigned      n_warps = (block + 31) / 32;
    std::vector<Fiber>  fibers(block);
    std::vector<Warp>   warps(n_warps);
    std::vector<double> smem(smem_bytes / sizeof(double) + 2);
    cur_kernel = &kernel;
    for (unsigned b = 0; b < grid; ++b)
      {
        std::fill(smem.begin(), smem.end(), nan("")); // uninitialised shared memory is not zero
        blk.block_idx = {b % grid3.x, (b / grid3.x) % grid3.y, b / (grid3.x * grid3.y)};
        blk.block_dim = dim3(block);
        blk.grid_dim  = grid3;
        blk.bar       = Bar{block, 0, 0};
        blk.smem      = reinterpret_cast<unsigned char *>(smem.data());
        for (unsigned w = 0; w < n_warps; ++w)
          {
            warps[w].n   = int(std::min(32u, block - 32 * w));
            warps[w].bar = Bar{unsigned(warps[w].n), 0, 0};
          }
        for (unsigned t = 0; t < block; ++t)
          {
            Fiber &f     = fibers[t];
            f.done       = false;
            f.thread_idx = {t, 0, 0};
            f.warp       = &warps[t / 32];
            f.parity     = 0;
            // initial frame: six callee-saved registers, then the entry point as return address;
            // at fiber_entry the stack pointer is 8 mod 16 as after a call
            uintptr_t top = (reinterpret_cast<uintptr_t>(stack_of(t)) + FIBER_STACK) & ~uintptr_t(15);
            void **   sp  = reinterpret_cast<void **>(top);
            *--sp         = nullptr;
            *--sp         = reinterpret_cast<void *>(&fiber_entry);
            for (int k = 0; k < 6; ++k)
              *--sp = nullptr;
            f.sp = sp;
          }
        // Scheduling order of the fibers between two yields. In-order execution would HIDE a
        // missing barrier of the kind "lower thread writes, higher thread reads", so the order is
        // selectable: GF_EMU_SCHEDULE = forward (default) | reverse | random:<seed> (a new
        // permutation every sweep). tests/test_emulated_library.py runs a subset under all three.
        static const int          mode = schedule_mode();
        static unsigned long long rng  = schedule_seed();
        std::vector<unsigned>     order(block);
        for (unsigned t = 0; t < block; ++t)
          order[t] = mode == 1 ? block - 1 - t : t;
        unsigned alive = block;
        while (alive > 0)
          {
            if (mode == 2)
              for (unsigned t = block; t > 1; --t)
                {
                  rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                  std::swap(order[t - 1], order[unsigned((rng >> 33) % t)]);
                }
            for (unsigned k = 0; k < block; ++k)
              {
                Fiber &f = fibers[order[k]];
                if (f.done)
                  continue;
                cur = &f;
                switch_context(&sched_sp, f.sp);
                if (f.done)
                  --alive;
              }
          }
      }
    cur = nullptr;
  }
  // <<<grid, block, smem, stream>>> as the sources write it (tests/cuda_emu/make_emu_library.py
  // rewrites the launch statements into this call)
  template <class G, class B, class S, class St>
  inline void launch4(G grid, B block, S smem_bytes, St /*stream*/,
                      const std::function<void()> &kernel)
  {
    if constexpr (std::is_same_v<G, dim3>)
      launch(grid, unsigned(block), size_t(smem_bytes), kernel);
    else
      launch(dim3(unsigned(grid)), unsigned(block), size_t(smem_bytes), kernel);
  }
} // namespace gf_emu

#define threadIdx (gf_emu::cur->thread_idx)
#define blockIdx (gf_emu::blk.block_idx)
#define blockDim (gf_emu::blk.block_dim)
#define gridDim (gf_emu::blk.grid_dim)

inline void __syncthreads() { gf_emu::bar_wait(gf_emu::blk.bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { gf_emu::bar_wait(gf_emu::cur->warp->bar); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask)
{
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  gf_emu::Fiber &f    = *gf_emu::cur;
  gf_emu::Warp & w    = *f.warp;
  const int      lane = int(f.thread_idx.x & 31);
  const unsigned par  = f.parity;
  f.parity ^= 1u;
  std::memcpy(w.slot[par][lane], &v, sizeof(T));
  gf_emu::bar_wait(w.bar);
  T         r   = v;
  const int src = lane ^ lane_mask;
  if (src < w.n)
    std::memcpy(&r, w.slot[par][src], sizeof(T));
  return r;
}
inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v)
{
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
template <class T>
inline T __ldg(const T *p)
{
  return *p;
}
template <class T>
inline T __ldcs(const T *p)
{
  return *p;
}
template <class T>
inline T __ldcg(const T *p)
{
  return *p;
}

// ---- runtime API --------------------------------------------------------------------------------
enum cudaError_t
{
  cudaSuccess      = 0,
  cudaErrorUnknown = 999
};
enum cudaMemcpyKind
{
  cudaMemcpyHostToHost,
  cudaMemcpyHostToDevice,
  cudaMemcpyDeviceToHost,
  cudaMemcpyDeviceToDevice,
  cudaMemcpyDefault
};
enum cudaFuncAttribute
{
  cudaFuncAttributeMaxDynamicSharedMemorySize = 8
};
typedef struct gf_emu_stream *cudaStream_t;
typedef double *              cudaEvent_t; // seconds at the record
constexpr unsigned            cudaStreamNonBlocking = 1;
struct cudaDeviceProp
{
  char   name[64]            = "gf_emu (CPU threads)";
  int    multiProcessorCount = 4;
  int    major = 10, minor = 0;
  size_t totalGlobalMem          = size_t(8) << 30;
  size_t sharedMemPerBlockOptin = 227 * 1024;
};
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n)
{
  *n = 8; // every emulated rank (process) may name its own device
  return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d)
{
  *d = 0;
  return cudaSuccess;
}
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
  *p = cudaDeviceProp();
  return cudaSuccess;
}
inline cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b)
{
  *free_b  = size_t(4) << 30;
  *total_b = size_t(8) << 30;
  return cudaSuccess;
}
// Blocks of >= 1 MiB live in named shared-memory files, so that a peer rank - another PROCESS - can
// map them (cudaIpcGetMemHandle / cudaIpcOpenMemHandle: the peer windows of comm.cu)
struct cudaIpcMemHandle_t
{
  char reserved[64];
};
constexpr unsigned cudaIpcMemLazyEnablePeerAccess = 1;
constexpr unsigned cudaHostAllocMapped            = 2;
namespace gf_emu
{
  struct Shared
  {
    std::string name;
    size_t      bytes;
    bool        owner;
  };
  inline std::map<void *, Shared> &shared_blocks()
  {
    static std::map<void *, Shared> m;
    return m;
  }
  inline void *map_shared(const std::string &name, size_t n, bool create)
  {
    const int fd = shm_open(name.c_str(), create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
    if (fd < 0)
      return nullptr;
    if (create && ftruncate(fd, off_t(n)) != 0)
      {
        close(fd);
        return nullptr;
      }
    if (!create)
      {
        struct stat st;
        fstat(fd, &st);
        n = size_t(st.st_size);
      }
    void *p = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED)
      return nullptr;
    shared_blocks()[p] = Shared{name, n, create};
    return p;
  }
} // namespace gf_emu
inline cudaError_t cudaMalloc(void **p, size_t n)
{
  if (n >= (size_t(1) << 20))
    {
      static int  counter = 0;
      const std::string name =
        "/gf_emu_" + std::to_string(long(getpid())) + "_" + std::to_string(counter++);
      *p = gf_emu::map_shared(name, n, true); // zero pages; large blocks are memset by their owners
      return *p ? cudaSuccess : cudaErrorUnknown;
    }
  *p = std::malloc(n ? n : 1);
  if (*p == nullptr)
    return cudaErrorUnknown;
  std::memset(*p, 0xFF, n);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void *p)
{
  auto it = gf_emu::shared_blocks().find(p);
  if (it != gf_emu::shared_blocks().end())
    {
      munmap(p, it->second.bytes);
      if (it->second.owner)
        shm_unlink(it->second.name.c_str());
      gf_emu::shared_blocks().erase(it);
      return cudaSuccess;
    }
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p)
{
  auto it = gf_emu::shared_blocks().find(p);
  if (it == gf_emu::shared_blocks().end() || it->second.name.size() >= sizeof(h->reserved))
    return cudaErrorUnknown;
  std::memset(h->reserved, 0, sizeof(h->reserved));
  std::memcpy(h->reserved, it->second.name.c_str(), it->second.name.size());
  return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned)
{
  h.reserved[sizeof(h.reserved) - 1] = 0;
  *p = gf_emu::map_shared(h.reserved, 0, false);
  return *p ? cudaSuccess : cudaErrorUnknown;
}
inline cudaError_t cudaIpcCloseMemHandle(void *p) { return cudaFree(p); }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned)
{
  *p = std::calloc(n ? n : 1, 1);
  return *p ? cudaSuccess : cudaErrorUnknown;
}
inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned)
{
  *d = h;
  return cudaSuccess;
}
// cuStreamWaitValue64 (ranks sharing a device): not in the emulation - every rank is its own device
enum cudaDriverEntryPointQueryResult
{
  cudaDriverEntryPointSuccess = 0,
  cudaDriverEntryPointSymbolNotFound = 1
};
constexpr unsigned long long cudaEnableDefault = 0;
inline cudaError_t cudaGetDriverEntryPoint(const char *, void **fn, unsigned long long,
                                           cudaDriverEntryPointQueryResult *qr)
{
  *fn = nullptr;
  *qr = cudaDriverEntryPointSymbolNotFound;
  return cudaSuccess;
}
inline cudaError_t cudaMallocHost(void **p, size_t n)
{
  *p = std::malloc(n ? n : 1);
  return *p ? cudaSuccess : cudaErrorUnknown;
}
template <class T>
inline cudaError_t cudaMallocHost(T **p, size_t n)
{
  return cudaMallocHost(reinterpret_cast<void **>(p), n);
}
inline cudaError_t cudaFreeHost(void *p)
{
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind)
{
  if (n)
    std::memmove(dst, src, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k,
                                   cudaStream_t = nullptr)
{
  return cudaMemcpy(dst, src, n, k);
}
inline cudaError_t cudaMemset(void *p, int v, size_t n)
{
  if (n)
    std::memset(p, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr)
{
  return cudaMemset(p, v, n);
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned)
{
  *s = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreate(cudaStream_t *s)
{
  *s = nullptr;
  return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e)
{
  *e = new double(0.0);
  return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e)
{
  delete e;
  return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr)
{
  *e = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
  return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
  *ms = float((*b - *a) * 1e3);
  return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F *, cudaFuncAttribute, int)
{
  return cudaSuccess;
}
template <class F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F *, int, size_t)
{
  *n = 1;
  return cudaSuccess;
}
