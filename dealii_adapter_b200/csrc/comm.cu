// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.
// The reference is single-process (adapter.h:152-154: this_mpi_process = 0, n_mpi_processes = 1),
// so this layer has no reference counterpart; it carries only the two exchanges the partitioned
// path needs: the ghost-DoF halo before a SpMV / assembly read and the scalar all-reduce of the
// CG dot products and Newton norms. NCCL is resolved at run time (dlopen) so that single-GPU use
// has no NCCL dependency.
//
// Two transports carry those exchanges:
//   peer windows (default)  every rank owns one device "window" (mailboxes + flags), exported with
//       cudaIpcGetMemHandle and mapped by all peers at gf_comm_create (the 64-byte handles travel
//       through one ncclAllGather). A halo exchange is then two of OUR kernels: `halo_push_kernel`
//       gathers the boundary values and stores them straight into the neighbour's mailbox over
//       NVLink, followed by a system-scope release of the neighbour's flag; `halo_wait_kernel`
//       acquires the local flag and unpacks the mailbox into the ghost entries. The scalar
//       all-reduce is ONE single-CTA kernel: store my partial sums into every peer's slot, release
//       the flags, acquire all local flags, add the slots in rank order (=> bitwise identical on
//       all ranks, independent of arrival order). No NCCL call, no proxy thread, ~2 launch
//       latencies instead of pack + grouped ncclSend/ncclRecv + unpack.
//       Mailboxes and flags are double-buffered by the parity of a communicator-wide epoch: rank A
//       can only push epoch e+2 after its wait of e+1, i.e. after B pushed e+1, which B's stream
//       orders after B's unpack of e — so the parity-(e mod 2) mailbox is free again.
//   NCCL (fallback, GF_COMM_P2P=0 or if IPC mapping fails on any rank): pack + grouped
//       ncclSend/ncclRecv + unpack, ncclAllReduce.
#include <dlfcn.h>
#ifdef GF_CUDA_EMULATION
#include <sched.h>
#include <time.h>
#endif

#include <cstdlib>
#include <cstring>

#include "reduce.cuh"

namespace gf
{
  namespace
  {
    struct NcclUniqueId
    {
      char internal[128];
    };
    typedef void *ncclComm_t;
    typedef int (*fn_GetUniqueId)(NcclUniqueId *);
    typedef int (*fn_CommInitRank)(ncclComm_t *, int, NcclUniqueId, int);
    typedef int (*fn_CommDestroy)(ncclComm_t);
    typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    typedef int (*fn_AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    typedef int (*fn_SendRecv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    typedef int (*fn_Group)(void);
    typedef const char *(*fn_ErrStr)(int);
    constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MIN = 3, NCCL_CHAR = 0, NCCL_INT32 = 2;

    struct NcclApi
    {
      void *          lib = nullptr;
      fn_GetUniqueId  GetUniqueId = nullptr;
      fn_CommInitRank CommInitRank = nullptr;
      fn_CommDestroy  CommDestroy = nullptr;
      fn_AllReduce    AllReduce = nullptr;
      fn_AllGather    AllGather = nullptr;
      fn_SendRecv     Send = nullptr, Recv = nullptr;
      fn_Group        GroupStart = nullptr, GroupEnd = nullptr;
      fn_ErrStr       GetErrorString = nullptr;
    };

    NcclApi &nccl()
    {
      static NcclApi api;
      if (api.lib)
        return api;
      const char *env = getenv("GF_NCCL_LIB");
      const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
      for (const char *n : names)
        {
          if (!n)
            continue;
          api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
          if (api.lib)
            break;
        }
      GF_REQUIRE(api.lib != nullptr, GF_ERR_NCCL,
                 "cannot load libnccl.so.2 (set GF_NCCL_LIB to its path)");
      api.GetUniqueId    = (fn_GetUniqueId)dlsym(api.lib, "ncclGetUniqueId");
      api.CommInitRank   = (fn_CommInitRank)dlsym(api.lib, "ncclCommInitRank");
      api.CommDestroy    = (fn_CommDestroy)dlsym(api.lib, "ncclCommDestroy");
      api.AllReduce      = (fn_AllReduce)dlsym(api.lib, "ncclAllReduce");
      api.AllGather      = (fn_AllGather)dlsym(api.lib, "ncclAllGather");
      api.Send           = (fn_SendRecv)dlsym(api.lib, "ncclSend");
      api.Recv           = (fn_SendRecv)dlsym(api.lib, "ncclRecv");
      api.GroupStart     = (fn_Group)dlsym(api.lib, "ncclGroupStart");
      api.GroupEnd       = (fn_Group)dlsym(api.lib, "ncclGroupEnd");
      api.GetErrorString = (fn_ErrStr)dlsym(api.lib, "ncclGetErrorString");
      GF_REQUIRE(api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce &&
                   api.Send && api.Recv && api.GroupStart && api.GroupEnd,
                 GF_ERR_NCCL, "libnccl is missing required symbols");
      return api;
    }

    void nccl_check(int rc, const char *what)
    {
      if (rc != 0)
        {
          NcclApi &   api = nccl();
          const char *msg = api.GetErrorString ? api.GetErrorString(rc) : "?";
          throw Error{GF_ERR_NCCL, std::string(what) + " failed: " + msg};
        }
    }

    __global__ void pack_kernel(int64_t n, const int32_t *__restrict__ idx,
                                const double *__restrict__ v, double *__restrict__ buf)
    {
      const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (k < n)
        buf[k] = v[idx[k]];
    }
    __global__ void unpack_kernel(int64_t n, const int32_t *__restrict__ idx,
                                  const double *__restrict__ buf, double *__restrict__ v)
    {
      const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (k < n)
        v[idx[k]] = buf[k];
    }
    // v[idx[k]] += buf[k] for ONE neighbour's list (its targets are distinct); the neighbours
    // are processed by consecutive launches, i.e. in a fixed order
    __global__ void unpack_add_kernel(int64_t n, const int32_t *__restrict__ idx,
                                      const double *__restrict__ buf, double *__restrict__ v)
    {
      const int64_t k = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
      if (k < n)
        v[idx[k]] += buf[k];
    }
  } // namespace

  // ---------------------------------------------------------------------------------------------
  // peer-window transport
  // ---------------------------------------------------------------------------------------------
  namespace
  {
    constexpr int PUSH_BLOCKS = 8, PUSH_THREADS = 512;

#ifndef GF_CUDA_EMULATION
    __device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
    {
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    }
    __device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
    {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
      return v;
    }
    __device__ __forceinline__ unsigned long long global_timer_ns()
    {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      return t;
    }
#else // tests/cuda_emu: ranks are processes, the windows shared-memory files
    inline void st_release_sys(unsigned long long *p, unsigned long long v)
    {
      __atomic_store_n(p, v, __ATOMIC_RELEASE);
    }
    inline unsigned long long ld_acquire_sys(const unsigned long long *p)
    {
      return __atomic_load_n(p, __ATOMIC_ACQUIRE);
    }
    inline unsigned long long global_timer_ns()
    {
      timespec ts;
      clock_gettime(CLOCK_MONOTONIC, &ts);
      return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
    }
    inline void __nanosleep(unsigned) { sched_yield(); }
#endif
    // spin until *flag >= epoch; a peer that never arrives sets *err instead of hanging the GPU
    __device__ __forceinline__ void wait_flag(const unsigned long long *flag,
                                              unsigned long long epoch, int *err,
                                              unsigned long long timeout_ns)
    {
      if (ld_acquire_sys(flag) >= epoch)
        return;
      // sticky: once one wait has timed out every later wait returns at once, so the error
      // reaches the host's next comm_check instead of costing a timeout per exchange
      if (*reinterpret_cast<volatile int *>(err) != 0)
        return;
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys(flag) < epoch)
        {
          __nanosleep(64);
          if (global_timer_ns() - t0 > timeout_ns)
            {
              *reinterpret_cast<volatile int *>(err) = 1;
              __threadfence_system();
              break;
            }
        }
    }

    struct HaloArgs
    {
      double *            mbox[P2P_MAX_RANKS];  // push: neighbour's mailbox for me; wait: my mailbox
      unsigned long long *flag[P2P_MAX_RANKS];  // push: neighbour's flag for me;    wait: my flag
      unsigned long long  epoch[P2P_MAX_RANKS]; // pair epoch of this exchange
      long long           off[P2P_MAX_RANKS + 1]; // offsets into the index list
    };

    // blockIdx.y = neighbour. Gather v[idx] straight into the neighbour's mailbox (NVLink
    // stores); the last CTA of a neighbour releases that neighbour's flag.
    __global__ void __launch_bounds__(PUSH_THREADS)
      halo_push_kernel(const HaloArgs a, const int32_t *__restrict__ idx,
                       const double *__restrict__ v, unsigned *counters)
    {
      const int       k  = blockIdx.y;
      const long long s0 = a.off[k], n = a.off[k + 1] - s0;
      double *        dst = a.mbox[k];
      for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
           i += (long long)gridDim.x * blockDim.x)
        dst[i] = v[idx[s0 + i]];
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0)
        {
          const unsigned old = atomicAdd(&counters[k], 1u);
          if (old == gridDim.x - 1)
            {
              counters[k] = 0; // next launch on this stream starts from zero
              __threadfence_system();
              st_release_sys(a.flag[k], a.epoch[k]);
            }
        }
    }

    // blockIdx.y + k0 = neighbour. Acquire my flag for that neighbour, then unpack its mailbox
    // into v[idx] (ADD: += for the reverse halo, one neighbour per launch => fixed order).
    template <bool ADD>
    __global__ void __launch_bounds__(PUSH_THREADS)
      halo_wait_kernel(const HaloArgs a, const int k0, const int32_t *__restrict__ idx,
                       double *__restrict__ v, int *err, const unsigned long long timeout_ns)
    {
      const int k = k0 + blockIdx.y;
      if (threadIdx.x == 0)
        wait_flag(a.flag[k], a.epoch[k], err, timeout_ns);
      __syncthreads();
      const long long r0 = a.off[k], n = a.off[k + 1] - r0;
      const double *  src = a.mbox[k];
      for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
           i += (long long)gridDim.x * blockDim.x)
        {
          const double x = __ldcg(src + i); // L2: the line was written by the peer over NVLink
          if (ADD)
            v[idx[r0 + i]] += x;
          else
            v[idx[r0 + i]] = x;
        }
    }

    struct ArArgs
    {
      unsigned char *win[P2P_MAX_RANKS];
      int            rank, n_ranks;
    };

    // Every exchange is TWO launches: a push kernel (remote stores + release of the peers' flags,
    // never conditional) and a wait kernel that acquires the local flags FIRST and then consumes.
    // No kernel ever waits for a flag after its own push: with ranks that share a device (tests)
    // the flag wait can then be a stream memory operation in front of the wait kernel, so a
    // waiting rank does not occupy the GPU (gf_comm_s::stream_waits).

    // all-reduce(sum) of `count` <= P2P_AR_MAX doubles: store my values into every rank's slot
    // (mine included) and release the flags ...
    __global__ void __launch_bounds__(64)
      p2p_ar_push_kernel(const ArArgs a, const double *vals, const int count,
                         const unsigned long long epoch)
    {
      const int t = threadIdx.x, par = int(epoch & 1ull);
      if (t < a.n_ranks * count)
        {
          const int r = t / count, j = t - r * count;
          double *  slot = reinterpret_cast<double *>(a.win[r] + P2P_AR_SLOT) +
                         (par * P2P_MAX_RANKS + a.rank) * P2P_AR_MAX + j;
          *slot = vals[j];
        }
      __threadfence_system();
      __syncthreads();
      if (t < a.n_ranks)
        st_release_sys(reinterpret_cast<unsigned long long *>(a.win[t] + P2P_AR_FLAG) +
                         par * P2P_MAX_RANKS + a.rank,
                       epoch);
    }
    // ... then acquire the flags of all ranks in my window and add the slots in rank order
    __global__ void __launch_bounds__(64)
      p2p_ar_sum_kernel(const ArArgs a, double *vals, const int count,
                        const unsigned long long epoch, int *err,
                        const unsigned long long timeout_ns)
    {
      const int t = threadIdx.x, par = int(epoch & 1ull);
      if (t < a.n_ranks)
        wait_flag(reinterpret_cast<const unsigned long long *>(a.win[a.rank] + P2P_AR_FLAG) +
                    par * P2P_MAX_RANKS + t,
                  epoch, err, timeout_ns);
      __syncthreads();
      if (t < count)
        {
          const double *slots = reinterpret_cast<const double *>(a.win[a.rank] + P2P_AR_SLOT) +
                                par * P2P_MAX_RANKS * P2P_AR_MAX;
          double s = 0.0;
          for (int r = 0; r < a.n_ranks; ++r)
            s += __ldcg(slots + r * P2P_AR_MAX + t);
          vals[t] = s;
        }
    }

    // Second stage of a partition-independent reduction across ranks (reduce.cuh): store my chunk
    // partials into every rank's gather window at their GLOBAL positions and release the flags ...
    __global__ void __launch_bounds__(TREE_THREADS)
      p2p_gather_push_kernel(const ArArgs a, const double *__restrict__ partials, const int stride,
                             const int n_local, const long long base, const int n_sums,
                             const unsigned long long epoch)
    {
      const int t = threadIdx.x, par = int(epoch & 1ull);
      for (int idx = t; idx < n_sums * n_local; idx += TREE_THREADS)
        {
          const int    k = idx / n_local, j = idx - k * n_local;
          const double v = partials[size_t(k) * stride + j];
          for (int r = 0; r < a.n_ranks; ++r)
            reinterpret_cast<double *>(a.win[r] + P2P_GATHER)[(size_t(par) * 3 + k) * P2P_GATHER_MAX +
                                                              size_t(base) + j] = v;
        }
      __threadfence_system();
      __syncthreads();
      if (t < a.n_ranks)
        st_release_sys(reinterpret_cast<unsigned long long *>(a.win[t] + P2P_GATHER_FLAG) +
                         par * P2P_MAX_RANKS + a.rank,
                       epoch);
    }
    // ... then acquire the flags of all ranks and run the fixed tree over the whole global chunk
    // list (every rank computes the same bits) and, for the CG, the scalar step. The handshake is
    // unconditional (a finished CG only skips the arithmetic): the parity double-buffering needs
    // every rank to have consumed epoch e before anybody pushes e + 2.
    __global__ void __launch_bounds__(TREE_THREADS)
      p2p_gather_tree_kernel(const ArArgs a, const int n_global, const int n_sums,
                             double *__restrict__ sums, CGScalars *s, const int phase,
                             const bool check_status, const unsigned long long epoch, int *err,
                             const unsigned long long timeout_ns)
    {
      __shared__ double sm[32];
      const int t = threadIdx.x, par = int(epoch & 1ull);
      if (t < a.n_ranks)
        wait_flag(reinterpret_cast<const unsigned long long *>(a.win[a.rank] + P2P_GATHER_FLAG) +
                    par * P2P_MAX_RANKS + t,
                  epoch, err, timeout_ns);
      __syncthreads();
      if (check_status && s->status != 0)
        return;
      const double *mine = reinterpret_cast<const double *>(a.win[a.rank] + P2P_GATHER) +
                           size_t(par) * 3 * P2P_GATHER_MAX;
      for (int k = 0; k < n_sums; ++k)
        {
          const double w = tree_sum_1024(mine + size_t(k) * P2P_GATHER_MAX, n_global, sm);
          if (t == 0)
            sums[k] = w;
        }
      if (phase >= 0 && t == 0)
        cg_scalar_step(s, sums, phase);
    }

    // NCCL transport: the padded all-gather result [rank][k][pad] -> the same fixed tree
    __global__ void __launch_bounds__(TREE_THREADS)
      gathered_tree_kernel(const double *__restrict__ gathered, double *__restrict__ scratch,
                           const int pad, const int n_ranks, const int *__restrict__ rank_off,
                           const int n_global, const int n_sums, double *__restrict__ sums,
                           CGScalars *s, const int phase, const bool check_status)
    {
      if (check_status && s->status != 0)
        return;
      __shared__ double sm[32];
      for (int k = 0; k < n_sums; ++k)
        {
          for (int r = 0; r < n_ranks; ++r)
            for (int j = threadIdx.x; j < rank_off[r + 1] - rank_off[r]; j += TREE_THREADS)
              scratch[rank_off[r] + j] = gathered[(size_t(r) * 3 + k) * pad + j];
          __threadfence_block();
          __syncthreads();
          const double w = tree_sum_1024(scratch, n_global, sm);
          if (threadIdx.x == 0)
            sums[k] = w;
          __syncthreads();
        }
      if (phase >= 0 && threadIdx.x == 0)
        cg_scalar_step(s, sums, phase);
    }

    // sum of a vector over all ranks through the halo mailboxes (count <= P2P_HALO_CAP): push my
    // vector into every peer's mailbox, then add the mailboxes in rank order (mine in its place)
    __global__ void __launch_bounds__(PUSH_THREADS)
      vec_push_kernel(const HaloArgs a, const long long n, const double *__restrict__ v,
                      unsigned *counters)
    {
      const int k   = blockIdx.y;
      double *  dst = a.mbox[k];
      for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
           i += (long long)gridDim.x * blockDim.x)
        dst[i] = v[i];
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0)
        {
          const unsigned old = atomicAdd(&counters[k], 1u);
          if (old == gridDim.x - 1)
            {
              counters[k] = 0;
              __threadfence_system();
              st_release_sys(a.flag[k], a.epoch[k]);
            }
        }
    }
    __global__ void __launch_bounds__(PUSH_THREADS)
      vec_sum_kernel(const HaloArgs a, const int n_peers, const int my_slot, const long long n,
                     double *__restrict__ v, int *err, const unsigned long long timeout_ns)
    {
      if (threadIdx.x < n_peers)
        wait_flag(a.flag[threadIdx.x], a.epoch[threadIdx.x], err, timeout_ns);
      __syncthreads();
      for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
           i += (long long)gridDim.x * blockDim.x)
        {
          double s = 0.0;
          for (int k = 0; k <= n_peers; ++k) // rank order: peers below me, me, peers above me
            {
              if (k == my_slot)
                s += v[i];
              if (k < n_peers)
                s += __ldcg(a.mbox[k] + i);
            }
          v[i] = s;
        }
    }

    // Ranks that share a device (gf_comm_s::stream_waits): the flag is awaited by a stream memory
    // operation in front of the wait kernel, so the stream blocks without a kernel spinning on the
    // SMs and the GPU can run the peer's context; the wait kernel then finds the flag set.
    typedef int (*fn_cuStreamWaitValue64)(cudaStream_t, unsigned long long, unsigned long long,
                                          unsigned);
    void stream_wait_flag(gf_comm cm, cudaStream_t s, unsigned char *flags_base, size_t index,
                          unsigned long long epoch)
    {
      if (!cm->stream_waits)
        return;
      static fn_cuStreamWaitValue64 fn = nullptr;
      if (!fn)
        {
          void *                           p = nullptr;
          cudaDriverEntryPointQueryResult qr;
          GF_CUDA_CHECK(cudaGetDriverEntryPoint("cuStreamWaitValue64", &p, cudaEnableDefault, &qr));
          GF_REQUIRE(p != nullptr && qr == cudaDriverEntryPointSuccess, GF_ERR_CUDA,
                     "cuStreamWaitValue64 is not available");
          fn = reinterpret_cast<fn_cuStreamWaitValue64>(p);
        }
      const unsigned long long addr =
        reinterpret_cast<unsigned long long>(flags_base) + index * sizeof(unsigned long long);
      const int rc = fn(s, addr, epoch, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
      GF_REQUIRE(rc == 0, GF_ERR_CUDA, "cuStreamWaitValue64 failed: " + std::to_string(rc));
    }

    // operations of one communicator must be ordered: if another handle used it on a different
    // stream, drain that stream first (multigrid levels share the finest level's stream)
    void comm_use_stream(gf_context &c)
    {
      gf_comm cm = c.comm;
      if (cm->last_stream != c.stream)
        {
          if (cm->last_stream)
            GF_CUDA_CHECK(cudaStreamSynchronize(cm->last_stream));
          cm->last_stream = c.stream;
        }
    }

    // direction: forward = owners -> ghosts (send lists out, recv lists in); reverse = ghosts ->
    // owners (recv lists out, send lists in, added)
    void p2p_halo(gf_context &c, double *v, bool reverse)
    {
      gf_comm   cm = c.comm;
      const int nn = int(c.nbr_rank.size());
      comm_use_stream(c);
      const std::vector<int64_t> &out_ptr = reverse ? c.recv_ptr : c.send_ptr;
      const std::vector<int64_t> &in_ptr  = reverse ? c.send_ptr : c.recv_ptr;
      const int32_t *out_idx = reverse ? c.recv_idx.p : c.send_idx.p;
      const int32_t *in_idx  = reverse ? c.send_idx.p : c.recv_idx.p;
      HaloArgs push{}, wait{};
      for (int k = 0; k < nn; ++k)
        {
          const int                r   = c.nbr_rank[k];
          const unsigned long long e   = ++cm->halo_epoch[r];
          const size_t             par = size_t(e & 1ull);
          // my data lands in the neighbour's window at [par][sender = me]
          push.mbox[k] = reinterpret_cast<double *>(cm->win[r] + P2P_MAILBOX) +
                         (par * P2P_MAX_RANKS + size_t(cm->rank)) * P2P_HALO_CAP;
          push.flag[k] = reinterpret_cast<unsigned long long *>(cm->win[r] + P2P_HALO_FLAG) +
                         par * P2P_MAX_RANKS + size_t(cm->rank);
          wait.mbox[k] = reinterpret_cast<double *>(cm->win[cm->rank] + P2P_MAILBOX) +
                         (par * P2P_MAX_RANKS + size_t(r)) * P2P_HALO_CAP;
          wait.flag[k] = reinterpret_cast<unsigned long long *>(cm->win[cm->rank] + P2P_HALO_FLAG) +
                         par * P2P_MAX_RANKS + size_t(r);
          push.epoch[k] = wait.epoch[k] = e;
          push.off[k]   = out_ptr[k];
          wait.off[k]   = in_ptr[k];
        }
      push.off[nn] = out_ptr[nn];
      wait.off[nn] = in_ptr[nn];
      halo_push_kernel<<<dim3(PUSH_BLOCKS, nn), PUSH_THREADS, 0, c.stream>>>(push, out_idx, v,
                                                                             cm->blk_counter);
      for (int k = 0; k < nn; ++k)
        stream_wait_flag(cm, c.stream, reinterpret_cast<unsigned char *>(wait.flag[k]), 0,
                         wait.epoch[k]);
      if (!reverse)
        halo_wait_kernel<false><<<dim3(PUSH_BLOCKS, nn), PUSH_THREADS, 0, c.stream>>>(
          wait, 0, in_idx, v, cm->d_err, cm->timeout_ns);
      else
        for (int k = 0; k < nn; ++k)
          halo_wait_kernel<true><<<dim3(PUSH_BLOCKS, 1), PUSH_THREADS, 0, c.stream>>>(
            wait, k, in_idx, v, cm->d_err, cm->timeout_ns);
      GF_CUDA_CHECK(cudaGetLastError());
      ++cm->n_halo;
    }
  } // namespace

  void comm_check(gf_context &c)
  {
    if (c.comm && c.comm->h_err && *reinterpret_cast<volatile int *>(c.comm->h_err) != 0)
      throw Error{GF_ERR_NCCL, "peer-window wait timed out: a neighbouring rank did not arrive "
                               "(ranks must issue the same sequence of exchanges)"};
  }

  void comm_forget_stream(gf_comm cm, cudaStream_t s)
  {
    if (cm && cm->last_stream == s)
      cm->last_stream = nullptr;
  }

  // collective over the communicator (gf_create is): all ranks use the peer mailboxes for this
  // handle's halo only if every rank's lists fit
  void comm_setup_context(gf_context &c)
  {
    c.halo_p2p = false;
    if (!c.comm || !c.comm->p2p)
      return;
    int fits = 1;
    for (size_t k = 0; k < c.nbr_rank.size(); ++k)
      {
        if (size_t(c.send_ptr[k + 1] - c.send_ptr[k]) > P2P_HALO_CAP ||
            size_t(c.recv_ptr[k + 1] - c.recv_ptr[k]) > P2P_HALO_CAP)
          fits = 0;
        if (c.nbr_rank[k] < 0 || c.nbr_rank[k] >= c.comm->n_ranks || c.nbr_rank[k] == c.comm->rank)
          fits = 0;
      }
    if (int(c.nbr_rank.size()) > P2P_MAX_RANKS)
      fits = 0;
    if (c.comm->nccl_comm == nullptr)
      {
        // host-bootstrapped communicator (gf_comm_ipc_*): no NCCL fallback transport exists, so
        // the number of ranks whose lists do NOT fit is agreed through the peer-window all-reduce
        double         bad = fits ? 0.0 : 1.0;
        DevBuf<double> d;
        d.upload(&bad, 1, c.stream);
        allreduce_sum(c, d.p, 1);
        d.download(&bad, c.stream);
        GF_REQUIRE(bad < 0.5, GF_ERR_UNSUPPORTED,
                   "halo lists exceed the peer mailboxes and the communicator has no NCCL transport");
        c.halo_p2p = true;
        return;
      }
    DevBuf<int> d;
    d.upload(&fits, 1, c.stream);
    comm_use_stream(c);
    nccl_check(nccl().AllReduce(d.p, d.p, 1, NCCL_INT32, NCCL_MIN, c.comm->nccl_comm, c.stream),
               "ncclAllReduce");
    d.download(&fits, c.stream);
    c.halo_p2p = fits != 0;
  }

  // second stage of a reduction across ranks (reduce.cuh)
  void comm_reduce_sums(gf_context &c, int n_sums, int cg_phase, bool check_status)
  {
    gf_comm cm = c.comm;
    comm_use_stream(c);
    ++cm->n_allreduce;
    if (cm->p2p)
      {
        ArArgs a{};
        for (int r = 0; r < cm->n_ranks; ++r)
          a.win[r] = cm->win[r];
        a.rank    = cm->rank;
        a.n_ranks = cm->n_ranks;
        const unsigned long long e = ++cm->gather_epoch;
        p2p_gather_push_kernel<<<1, TREE_THREADS, 0, c.stream>>>(
          a, c.partials.p, c.red_stride, c.n_red_chunks, (long long)c.red_chunk_base, n_sums, e);
        for (int r = 0; r < cm->n_ranks; ++r)
          stream_wait_flag(cm, c.stream, cm->win[cm->rank] + P2P_GATHER_FLAG,
                           (e & 1ull) * P2P_MAX_RANKS + r, e);
        p2p_gather_tree_kernel<<<1, TREE_THREADS, 0, c.stream>>>(
          a, int(c.n_red_chunks_global), n_sums, red_sums(c), c.cg_scalars.p, cg_phase,
          check_status, e, cm->d_err, cm->timeout_ns);
        GF_CUDA_CHECK(cudaGetLastError());
        return;
      }
    // NCCL transport: all-gather of the (padded) partials, then the same tree
    NcclApi & api = nccl();
    const int P   = cm->n_ranks;
    int       pad = 1;
    for (int r = 0; r < P; ++r)
      pad = std::max(pad, c.red_rank_chunks[r]);
    if (!c.red_gather.p)
      {
        c.red_gather.alloc_zero(size_t(P + 1) * 3 * pad + size_t(c.n_red_chunks_global) + 8, c.stream);
        std::vector<int> off(P + 1, 0);
        for (int r = 0; r < P; ++r)
          off[r + 1] = off[r] + c.red_rank_chunks[r];
        c.red_rank_off.upload(off.data(), off.size(), c.stream);
      }
    double *send    = c.red_gather.p + size_t(P) * 3 * pad;
    double *scratch = send + size_t(3) * pad;
    for (int k = 0; k < n_sums; ++k)
      GF_CUDA_CHECK(cudaMemcpyAsync(send + size_t(k) * pad, c.partials.p + size_t(k) * c.red_stride,
                                    size_t(c.n_red_chunks) * sizeof(double),
                                    cudaMemcpyDeviceToDevice, c.stream));
    nccl_check(api.AllGather(send, c.red_gather.p, size_t(3) * pad, NCCL_FLOAT64, cm->nccl_comm,
                             c.stream),
               "ncclAllGather");
    gathered_tree_kernel<<<1, TREE_THREADS, 0, c.stream>>>(
      c.red_gather.p, scratch, pad, P, c.red_rank_off.p, int(c.n_red_chunks_global), n_sums,
      red_sums(c), c.cg_scalars.p, cg_phase, check_status);
    GF_CUDA_CHECK(cudaGetLastError());
  }

  // exchange ghost values of v with the slab neighbours
  void halo_exchange(gf_context &c, double *v)
  {
    if (!c.comm || c.nbr_rank.empty())
      return;
    ProfScope ps(c, Profile::HALO, 2);
    if (c.halo_p2p)
      {
        p2p_halo(c, v, false);
        return;
      }
    comm_use_stream(c);
    NcclApi &     api = nccl();
    const int64_t ns = c.send_ptr.back(), nr = c.recv_ptr.back();
    if (ns)
      pack_kernel<<<unsigned((ns + 255) / 256), 256, 0, c.stream>>>(ns, c.send_idx.p, v,
                                                                   c.send_buf.p);
    nccl_check(api.GroupStart(), "ncclGroupStart");
    for (size_t k = 0; k < c.nbr_rank.size(); ++k)
      {
        const int64_t s0 = c.send_ptr[k], s1 = c.send_ptr[k + 1];
        const int64_t r0 = c.recv_ptr[k], r1 = c.recv_ptr[k + 1];
        if (s1 > s0)
          nccl_check(api.Send(c.send_buf.p + s0, size_t(s1 - s0), NCCL_FLOAT64, c.nbr_rank[k],
                              c.comm->nccl_comm, c.stream),
                     "ncclSend");
        if (r1 > r0)
          nccl_check(api.Recv(c.recv_buf.p + r0, size_t(r1 - r0), NCCL_FLOAT64, c.nbr_rank[k],
                              c.comm->nccl_comm, c.stream),
                     "ncclRecv");
      }
    nccl_check(api.GroupEnd(), "ncclGroupEnd");
    if (nr)
      unpack_kernel<<<unsigned((nr + 255) / 256), 256, 0, c.stream>>>(nr, c.recv_idx.p,
                                                                     c.recv_buf.p, v);
    GF_CUDA_CHECK(cudaGetLastError());
    ++c.comm->n_halo;
  }

  // reverse halo: every rank sends the values it accumulated on its GHOST dofs to their owners,
  // which add them to their own value in neighbour order (fixed => reproducible). Used by the
  // multigrid restriction, whose partial sums are taken over owned fine nodes only.
  void halo_reduce_add(gf_context &c, double *v)
  {
    if (!c.comm || c.nbr_rank.empty())
      return;
    ProfScope ps(c, Profile::HALO, 2);
    if (c.halo_p2p)
      {
        p2p_halo(c, v, true);
        return;
      }
    comm_use_stream(c);
    NcclApi &     api = nccl();
    const int64_t nr = c.recv_ptr.back();
    if (nr)
      pack_kernel<<<unsigned((nr + 255) / 256), 256, 0, c.stream>>>(nr, c.recv_idx.p, v,
                                                                   c.recv_buf.p);
    nccl_check(api.GroupStart(), "ncclGroupStart");
    for (size_t k = 0; k < c.nbr_rank.size(); ++k)
      {
        const int64_t s0 = c.send_ptr[k], s1 = c.send_ptr[k + 1];
        const int64_t r0 = c.recv_ptr[k], r1 = c.recv_ptr[k + 1];
        if (r1 > r0)
          nccl_check(api.Send(c.recv_buf.p + r0, size_t(r1 - r0), NCCL_FLOAT64, c.nbr_rank[k],
                              c.comm->nccl_comm, c.stream),
                     "ncclSend");
        if (s1 > s0)
          nccl_check(api.Recv(c.send_buf.p + s0, size_t(s1 - s0), NCCL_FLOAT64, c.nbr_rank[k],
                              c.comm->nccl_comm, c.stream),
                     "ncclRecv");
      }
    nccl_check(api.GroupEnd(), "ncclGroupEnd");
    for (size_t k = 0; k < c.nbr_rank.size(); ++k)
      {
        const int64_t s0 = c.send_ptr[k], n = c.send_ptr[k + 1] - s0;
        if (n)
          unpack_add_kernel<<<unsigned((n + 255) / 256), 256, 0, c.stream>>>(
            n, c.send_idx.p + s0, c.send_buf.p + s0, v);
      }
    GF_CUDA_CHECK(cudaGetLastError());
    ++c.comm->n_halo;
  }

  void allreduce_sum(gf_context &c, double *dev_values, int count)
  {
    if (!c.comm)
      return;
    comm_use_stream(c);
    gf_comm cm = c.comm;
    ++cm->n_allreduce;
    if (cm->p2p && count <= P2P_AR_MAX)
      {
        ArArgs a{};
        for (int r = 0; r < cm->n_ranks; ++r)
          a.win[r] = cm->win[r];
        a.rank    = cm->rank;
        a.n_ranks = cm->n_ranks;
        const unsigned long long e = ++cm->ar_epoch;
        p2p_ar_push_kernel<<<1, 64, 0, c.stream>>>(a, dev_values, count, e);
        for (int r = 0; r < cm->n_ranks; ++r)
          stream_wait_flag(cm, c.stream, cm->win[cm->rank] + P2P_AR_FLAG, (e & 1ull) * P2P_MAX_RANKS + r,
                           e);
        p2p_ar_sum_kernel<<<1, 64, 0, c.stream>>>(a, dev_values, count, e, cm->d_err,
                                                  cm->timeout_ns);
        GF_CUDA_CHECK(cudaGetLastError());
        return;
      }
    GF_REQUIRE(cm->nccl_comm != nullptr, GF_ERR_UNSUPPORTED,
               "all-reduce of more than 8 values without an NCCL transport");
    NcclApi &api = nccl();
    nccl_check(api.AllReduce(dev_values, dev_values, size_t(count), NCCL_FLOAT64, NCCL_SUM,
                             cm->nccl_comm, c.stream),
               "ncclAllReduce");
  }

  // sum of a long vector over all ranks (multigrid: restriction onto a replicated coarse level).
  // ncclAllReduce delivers the same bits to every rank, which the redundant coarse solves rely on.
  void allreduce_sum_vector(gf_context &c, double *dev_values, int64_t count)
  {
    if (!c.comm || count <= 0)
      return;
    ProfScope ps(c, Profile::HALO, 2);
    comm_use_stream(c);
    gf_comm cm = c.comm;
    ++cm->n_allreduce;
    if (cm->p2p && size_t(count) <= P2P_HALO_CAP)
      {
        // every pair (me, r) exchanges: the pair epochs advance exactly as in a halo exchange
        // with all ranks as neighbours, so the parity double-buffering argument carries over
        HaloArgs push{}, wait{};
        int      np = 0, my_slot = 0;
        for (int r = 0; r < cm->n_ranks; ++r)
          {
            if (r == cm->rank)
              {
                my_slot = np;
                continue;
              }
            const unsigned long long e   = ++cm->halo_epoch[r];
            const size_t             par = size_t(e & 1ull);
            push.mbox[np] = reinterpret_cast<double *>(cm->win[r] + P2P_MAILBOX) +
                            (par * P2P_MAX_RANKS + size_t(cm->rank)) * P2P_HALO_CAP;
            push.flag[np] = reinterpret_cast<unsigned long long *>(cm->win[r] + P2P_HALO_FLAG) +
                            par * P2P_MAX_RANKS + size_t(cm->rank);
            wait.mbox[np] = reinterpret_cast<double *>(cm->win[cm->rank] + P2P_MAILBOX) +
                            (par * P2P_MAX_RANKS + size_t(r)) * P2P_HALO_CAP;
            wait.flag[np] = reinterpret_cast<unsigned long long *>(cm->win[cm->rank] + P2P_HALO_FLAG) +
                            par * P2P_MAX_RANKS + size_t(r);
            push.epoch[np] = wait.epoch[np] = e;
            ++np;
          }
        const int blocks = int(std::min<int64_t>(32, (count + PUSH_THREADS - 1) / PUSH_THREADS));
        vec_push_kernel<<<dim3(blocks, np), PUSH_THREADS, 0, c.stream>>>(push, count, dev_values,
                                                                         cm->blk_counter);
        for (int k = 0; k < np; ++k)
          stream_wait_flag(cm, c.stream, reinterpret_cast<unsigned char *>(wait.flag[k]), 0,
                           wait.epoch[k]);
        vec_sum_kernel<<<blocks, PUSH_THREADS, 0, c.stream>>>(wait, np, my_slot, count, dev_values,
                                                              cm->d_err, cm->timeout_ns);
        GF_CUDA_CHECK(cudaGetLastError());
        return;
      }
    GF_REQUIRE(cm->nccl_comm != nullptr, GF_ERR_UNSUPPORTED,
               "vector all-reduce exceeds the peer mailboxes and the communicator has no NCCL");
    nccl_check(nccl().AllReduce(dev_values, dev_values, size_t(count), NCCL_FLOAT64, NCCL_SUM,
                                cm->nccl_comm, c.stream),
               "ncclAllReduce");
  }

  namespace
  {
    // map every rank's window into this process (cudaIpc over NVLink peer access); collective
    void p2p_setup(gf_comm cm)
    {
      NcclApi &   api = nccl();
      const char *env = getenv("GF_COMM_P2P");
      int         ok  = (env && atoi(env) == 0) ? 0 : 1;
      if (cm->n_ranks < 2 || cm->n_ranks > P2P_MAX_RANKS || !api.AllGather)
        ok = 0;
      if (const char *t = getenv("GF_P2P_TIMEOUT_S"))
        cm->timeout_ns = (unsigned long long)(atof(t) * 1e9);
      cudaStream_t s = nullptr;
      GF_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
      unsigned char *mine = nullptr;
      cudaIpcMemHandle_t my_handle{};
      if (ok)
        {
          if (cudaMalloc((void **)&mine, P2P_WINDOW_BYTES) != cudaSuccess ||
              cudaMemset(mine, 0, P2P_WINDOW_BYTES) != cudaSuccess ||
              cudaIpcGetMemHandle(&my_handle, mine) != cudaSuccess)
            {
              cudaGetLastError();
              ok = 0;
            }
        }
      // everyone learns everyone's handle (64 B each) through one all-gather
      DevBuf<unsigned char> d_handles;
      d_handles.alloc(size_t(cm->n_ranks) * sizeof(cudaIpcMemHandle_t));
      GF_CUDA_CHECK(cudaMemcpy(d_handles.p + size_t(cm->rank) * sizeof(cudaIpcMemHandle_t),
                               &my_handle, sizeof(cudaIpcMemHandle_t), cudaMemcpyHostToDevice));
      if (api.AllGather)
        nccl_check(api.AllGather(d_handles.p + size_t(cm->rank) * sizeof(cudaIpcMemHandle_t),
                                 d_handles.p, sizeof(cudaIpcMemHandle_t), NCCL_CHAR, cm->nccl_comm,
                                 s),
                   "ncclAllGather");
      GF_CUDA_CHECK(cudaStreamSynchronize(s));
      std::vector<cudaIpcMemHandle_t> handles(cm->n_ranks);
      GF_CUDA_CHECK(cudaMemcpy(handles.data(), d_handles.p,
                               size_t(cm->n_ranks) * sizeof(cudaIpcMemHandle_t),
                               cudaMemcpyDeviceToHost));
      // agree that every rank allocated before anybody maps
      DevBuf<int> d_ok;
      auto        agree = [&](int &flag) {
        d_ok.upload(&flag, 1, s);
        nccl_check(api.AllReduce(d_ok.p, d_ok.p, 1, NCCL_INT32, NCCL_MIN, cm->nccl_comm, s),
                   "ncclAllReduce");
        d_ok.download(&flag, s);
      };
      agree(ok);
      if (ok)
        {
          cm->win[cm->rank] = mine;
          for (int r = 0; r < cm->n_ranks && ok; ++r)
            if (r != cm->rank)
              {
                void *ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, handles[r], cudaIpcMemLazyEnablePeerAccess) !=
                    cudaSuccess)
                  {
                    cudaGetLastError();
                    ok = 0;
                  }
                cm->win[r] = static_cast<unsigned char *>(ptr);
              }
          if (ok && (cudaMalloc((void **)&cm->blk_counter, P2P_MAX_RANKS * sizeof(unsigned)) !=
                       cudaSuccess ||
                     cudaMemset(cm->blk_counter, 0, P2P_MAX_RANKS * sizeof(unsigned)) !=
                       cudaSuccess ||
                     cudaHostAlloc((void **)&cm->h_err, sizeof(int), cudaHostAllocMapped) !=
                       cudaSuccess))
            {
              cudaGetLastError();
              ok = 0;
            }
          if (ok)
            {
              *cm->h_err = 0;
              if (cudaHostGetDevicePointer((void **)&cm->d_err, cm->h_err, 0) != cudaSuccess)
                {
                  cudaGetLastError();
                  ok = 0;
                }
            }
          agree(ok); // a rank that could not map its peers takes everybody to the NCCL transport
        }
      if (!ok)
        {
          for (int r = 0; r < cm->n_ranks; ++r)
            if (r != cm->rank && cm->win[r])
              cudaIpcCloseMemHandle(cm->win[r]);
          for (int r = 0; r < P2P_MAX_RANKS; ++r)
            cm->win[r] = nullptr;
          if (mine)
            cudaFree(mine);
          cudaGetLastError();
        }
      cm->p2p = ok != 0;
      cudaStreamDestroy(s);
      if (getenv("GF_COMM_VERBOSE"))
        fprintf(stderr, "graft_fem: rank %d/%d transport = %s\n", cm->rank, cm->n_ranks,
                cm->p2p ? "peer windows (cudaIpc over NVLink)" : "NCCL");
    }
  } // namespace
  void comm_p2p_setup(gf_comm cm) { p2p_setup(cm); }
  void comm_p2p_teardown(gf_comm cm)
  {
    if (!cm->p2p)
      return;
    cudaDeviceSynchronize();
    for (int r = 0; r < cm->n_ranks; ++r)
      if (r != cm->rank && cm->win[r])
        cudaIpcCloseMemHandle(cm->win[r]);
    if (cm->win[cm->rank])
      cudaFree(cm->win[cm->rank]);
    if (cm->blk_counter)
      cudaFree(cm->blk_counter);
    if (cm->h_err)
      cudaFreeHost(cm->h_err);
    cm->p2p = false;
  }
} // namespace gf

extern "C"
{
  int gf_comm_unique_id(uint8_t id[128])
  {
    try
      {
        gf::NcclUniqueId uid;
        gf::nccl_check(gf::nccl().GetUniqueId(&uid), "ncclGetUniqueId");
        memcpy(id, uid.internal, 128);
        return GF_OK;
      }
    catch (gf::Error &e)
      {
        fprintf(stderr, "graft_fem: %s\n", e.msg.c_str());
        return e.code;
      }
  }
  int gf_comm_create(const uint8_t id[128], int rank, int n_ranks, int device, gf_comm *out)
  {
    try
      {
        GF_CUDA_CHECK(cudaSetDevice(device));
        gf::NcclUniqueId uid;
        memcpy(uid.internal, id, 128);
        gf_comm_s *cm = new gf_comm_s;
        cm->rank      = rank;
        cm->n_ranks   = n_ranks;
        cm->device    = device;
        gf::nccl_check(gf::nccl().CommInitRank(&cm->nccl_comm, n_ranks, uid, rank),
                       "ncclCommInitRank");
        gf::comm_p2p_setup(cm);
        *out = cm;
        return GF_OK;
      }
    catch (gf::Error &e)
      {
        fprintf(stderr, "graft_fem: %s\n", e.msg.c_str());
        return e.code;
      }
  }
  // NCCL-free bootstrap of the peer-window transport: the host moves the 64-byte cudaIpc handles
  // through whatever channel it has (MPI, torch.distributed/gloo, a file). Also the only way to
  // run several ranks on ONE device (tests: NCCL refuses two ranks per GPU).
  int gf_comm_ipc_begin(int rank, int n_ranks, int device, gf_comm *out, uint8_t handle[64])
  {
    gf_comm_s *cm = nullptr;
    try
      {
        GF_REQUIRE(out && handle && n_ranks >= 2 && n_ranks <= gf::P2P_MAX_RANKS && rank >= 0 &&
                     rank < n_ranks,
                   GF_ERR_INVALID_ARG, "gf_comm_ipc_begin: 2..8 ranks");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
        GF_CUDA_CHECK(cudaSetDevice(device));
        cm          = new gf_comm_s;
        cm->rank    = rank;
        cm->n_ranks = n_ranks;
        cm->device  = device;
        if (const char *t = getenv("GF_P2P_TIMEOUT_S"))
          cm->timeout_ns = (unsigned long long)(atof(t) * 1e9);
        unsigned char *mine = nullptr;
        GF_CUDA_CHECK(cudaMalloc((void **)&mine, gf::P2P_WINDOW_BYTES));
        cm->win[rank] = mine;
        GF_CUDA_CHECK(cudaMemset(mine, 0, gf::P2P_WINDOW_BYTES));
        GF_CUDA_CHECK(cudaDeviceSynchronize());
        cudaIpcMemHandle_t h;
        GF_CUDA_CHECK(cudaIpcGetMemHandle(&h, mine));
        memcpy(handle, &h, 64);
        *out = cm;
        return GF_OK;
      }
    catch (gf::Error &e)
      {
        fprintf(stderr, "graft_fem: %s\n", e.msg.c_str());
        if (cm)
          {
            if (cm->win[rank])
              cudaFree(cm->win[rank]);
            delete cm;
          }
        return e.code;
      }
  }
  int gf_comm_ipc_finish(gf_comm cm, const uint8_t *all_handles, int ranks_share_device)
  {
    try
      {
        GF_REQUIRE(cm && all_handles && !cm->p2p, GF_ERR_INVALID_ARG, "gf_comm_ipc_finish");
        GF_CUDA_CHECK(cudaSetDevice(cm->device));
        for (int r = 0; r < cm->n_ranks; ++r)
          if (r != cm->rank)
            {
              cudaIpcMemHandle_t h;
              memcpy(&h, all_handles + size_t(r) * 64, 64);
              void *ptr = nullptr;
              GF_CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
              cm->win[r] = static_cast<unsigned char *>(ptr);
            }
        GF_CUDA_CHECK(cudaMalloc((void **)&cm->blk_counter, gf::P2P_MAX_RANKS * sizeof(unsigned)));
        GF_CUDA_CHECK(cudaMemset(cm->blk_counter, 0, gf::P2P_MAX_RANKS * sizeof(unsigned)));
        GF_CUDA_CHECK(cudaHostAlloc((void **)&cm->h_err, sizeof(int), cudaHostAllocMapped));
        *cm->h_err = 0;
        GF_CUDA_CHECK(cudaHostGetDevicePointer((void **)&cm->d_err, cm->h_err, 0));
        GF_CUDA_CHECK(cudaDeviceSynchronize());
        cm->p2p          = true;
        cm->stream_waits = ranks_share_device != 0 || getenv("GF_COMM_STREAM_WAIT") != nullptr;
        return GF_OK;
      }
    catch (gf::Error &e)
      {
        fprintf(stderr, "graft_fem: %s\n", e.msg.c_str());
        return e.code;
      }
  }
  int gf_comm_transport(gf_comm cm, int64_t *n_halo, int64_t *n_allreduce)
  {
    if (!cm)
      return -1;
    if (n_halo)
      *n_halo = cm->n_halo;
    if (n_allreduce)
      *n_allreduce = cm->n_allreduce;
    return cm->p2p ? GF_TRANSPORT_PEER_WINDOWS : GF_TRANSPORT_NCCL;
  }
  void gf_comm_destroy(gf_comm cm)
  {
    if (!cm)
      return;
    gf::comm_p2p_teardown(cm);
    if (cm->nccl_comm)
      gf::nccl().CommDestroy(cm->nccl_comm);
    delete cm;
  }
}
