// TEST INFRASTRUCTURE ONLY. The CPU emulation build of the library (make_emu_library.py) compiles
// the library's own sources; the one part that cannot run there is stood in for here: the
// single-launch coarsest-level solver (a grid-wide spin barrier among CTAs: blocks run one after
// another in the emulation) reports "not applicable", so the multigrid takes its multi-launch
// fallback. Communicators: the peer windows of comm.cu work across PROCESSES (shared-memory files
// behind cudaIpc*, see cuda_runtime.h); the NCCL transport is never reached (IPC bootstrap only).
#include "gf_context.h"

namespace gf
{
  bool coarse_solve_single_launch(gf_context &, const double *, const double *, double *, int, double)
  {
    return false;
  }
} // namespace gf
