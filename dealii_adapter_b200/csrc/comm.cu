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

  // exchange ghost values of v with the slab neighbours (grouped ncclSend/ncclRecv)
  void halo_exchange(gf_context &c, double *v)
  {
    if (!c.comm || c.nbr_rank.empty())
      return;
    ProfScope     ps(c, Profile::HALO, 2);
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
  }

  // reverse halo: every rank sends the values it accumulated on its GHOST dofs to their owners,
  // which add them to their own value in neighbour order (fixed => reproducible). Used by the
  // multigrid restriction, whose partial sums are taken over owned fine nodes only.
  void halo_reduce_add(gf_context &c, double *v)
  {
    if (!c.comm || c.nbr_rank.empty())
      return;
    ProfScope     ps(c, Profile::HALO, 2);
    NcclApi &     api = nccl();
    const int64_t ns = c.send_ptr.back(), nr = c.recv_ptr.back();
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
    (void)ns;
    GF_CUDA_CHECK(cudaGetLastError());
  }

  void allreduce_sum(gf_context &c, double *dev_values, int count)
  {
    if (!c.comm)
      return;
    NcclApi &api = nccl();
    nccl_check(api.AllReduce(dev_values, dev_values, size_t(count), NCCL_FLOAT64, NCCL_SUM,
                             c.comm->nccl_comm, c.stream),
               "ncclAllReduce");
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
        *out = cm;
        return GF_OK;
      }
    catch (gf::Error &e)
      {
        fprintf(stderr, "graft_fem: %s\n", e.msg.c_str());
        return e.code;
      }
  }
  void gf_comm_destroy(gf_comm cm)
  {
    if (!cm)
      return;
    if (cm->nccl_comm)
      gf::nccl().CommDestroy(cm->nccl_comm);
    delete cm;
  }
}
