// Reference-cell tables on the device: fe_tables_host.h computes them (FE_Q(p) in deal.II's
// hierarchical order at QGauss points, face tables in QProjector order), this file uploads them.
#include "gf_context.h"

namespace gf
{
  void build_tables(gf_context &c)
  {
    FETables &t = c.tables;
    build_host_tables(t, c.dim, c.p, c.model == GF_MODEL_NEO_HOOKEAN);
    t.N.upload(t.hN.data(), t.hN.size(), c.stream);
    t.dN.upload(t.hdN.data(), t.hdN.size(), c.stream);
    t.w.upload(t.hw.data(), t.hw.size(), c.stream);
    t.Nf.upload(t.hNf.data(), t.hNf.size(), c.stream);
    t.wf.upload(t.hwf.data(), t.hwf.size(), c.stream);
    t.Mref.upload(t.hMref.data(), t.hMref.size(), c.stream);
    t.dphi.upload(t.hdphi.data(), t.hdphi.size(), c.stream);
    t.dphif.upload(t.hdphif.data(), t.hdphif.size(), c.stream);
    t.dphip.upload(t.hdphip.data(), t.hdphip.size(), c.stream);
  }
} // namespace gf
