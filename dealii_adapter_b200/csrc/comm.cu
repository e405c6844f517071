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

#include <cstdlib>
#include <cstring>

#include "gf_context.h"

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
    // spin until *flag >= epoch; a peer that never arrives sets *err instead of hanging the GPU
    __device__ __forceinline__ void wait_flag(const unsigned long long *flag,
                                              unsigned long long epoch, int *err,
                                              unsigned long long timeout_ns)
    {
      if (ld_acquire_sys(flag) >= epoch)
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

    // all-reduce(sum) of `count` <= P2P_AR_MAX doubles in ONE single-CTA kernel: scatter my values
    // into every rank's slot (mine included), release the flags, acquire the flags of all ranks
    // in my window, add the slots in rank order.
    __global__ void __launch_bounds__(64)
      p2p_allreduce_kernel(const ArArgs a, double *vals, const int count,
                           const unsigned long long epoch, int *err,
                           const unsigned long long timeout_ns)
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
        {
          st_release_sys(reinterpret_cast<unsigned long long *>(a.win[t] + P2P_AR_FLAG) +
                           par * P2P_MAX_RANKS + a.rank,
                         epoch);
          wait_flag(reinterpret_cast<const unsigned long long *>(a.win[a.rank] + P2P_AR_FLAG) +
                      par * P2P_MAX_RANKS + t,
                    epoch, err, timeout_ns);
        }
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
    DevBuf<int> d;
    d.upload(&fits, 1, c.stream);
    comm_use_stream(c);
    nccl_check(nccl().AllReduce(d.p, d.p, 1, NCCL_INT32, NCCL_MIN, c.comm->nccl_comm, c.stream),
               "ncclAllReduce");
    d.download(&fits, c.stream);
    c.halo_p2p = fits != 0;
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
        p2p_allreduce_kernel<<<1, 64, 0, c.stream>>>(a, dev_values, count, ++cm->ar_epoch,
                                                     cm->d_err, cm->timeout_ns);
        GF_CUDA_CHECK(cudaGetLastError());
        return;
      }
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
    ProfScope ps(c, Profile::HALO, 1);
    comm_use_stream(c);
    ++c.comm->n_allreduce;
    nccl_check(nccl().AllReduce(dev_values, dev_values, size_t(count), NCCL_FLOAT64, NCCL_SUM,
                                c.comm->nccl_comm, c.stream),
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
