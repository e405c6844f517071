// TEST INFRASTRUCTURE ONLY. The CPU emulation build of the library (make_emu_library.py) covers
// the serial handle; what needs the hardware is stood in for here: the single-launch coarsest-level
// solver (a grid-wide spin barrier: blocks run one after another in the emulation) reports "not
// applicable", so the multigrid takes its multi-launch fallback, and everything that needs a
// communicator (NVLink peer windows, NCCL) answers GF_ERR_UNSUPPORTED.
#include "gf_context.h"
#include "reduce.cuh"

namespace gf
{
  namespace
  {
    [[noreturn]] void no(const char *what)
    {
      throw Error{GF_ERR_UNSUPPORTED, std::string(what) + " does not exist in the CPU emulation"};
    }
  } // namespace
  bool   coarse_solve_single_launch(gf_context &, const double *, const double *, double *, int, double)
  {
    return false;
  }
  void halo_reduce_add(gf_context &, double *) { no("a communicator"); }
  void halo_exchange(gf_context &, double *) { no("a communicator"); }
  void allreduce_sum(gf_context &, double *, int) { no("a communicator"); }
  void allreduce_sum_vector(gf_context &, double *, int64_t) { no("a communicator"); }
  void comm_setup_context(gf_context &c)
  {
    if (c.comm != nullptr)
      no("a communicator");
  }
  void comm_check(gf_context &) {}
  void comm_reduce_sums(gf_context &, int, int, bool) { no("a communicator"); }
  void comm_forget_stream(gf_comm, cudaStream_t) {}
} // namespace gf

extern "C"
{
  int  gf_comm_unique_id(uint8_t *) { return GF_ERR_UNSUPPORTED; }
  int  gf_comm_create(const uint8_t *, int, int, int, gf_comm *) { return GF_ERR_UNSUPPORTED; }
  int  gf_comm_ipc_begin(int, int, int, gf_comm *, uint8_t *) { return GF_ERR_UNSUPPORTED; }
  int  gf_comm_ipc_finish(gf_comm, const uint8_t *, int) { return GF_ERR_UNSUPPORTED; }
  int  gf_comm_transport(gf_comm, int64_t *, int64_t *) { return GF_ERR_UNSUPPORTED; }
  void gf_comm_destroy(gf_comm) {}
}
