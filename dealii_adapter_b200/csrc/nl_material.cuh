// Compressible neo-Hookean material (compressible_neo_hook_material.h:37-49,62-138) in Voigt
// storage, shared by the assembly kernel (assemble_nl.cu) and the matrix-free operator (matfree.cu).
#pragma once
#include "graft_fem.h"
#include "kernel_utils.cuh"

namespace gf
{
  struct NLParams
  {
    double kappa, mu, rho, alpha_1;
    double body_force[3];
  };

  // compressible_neo_hook_material.h:37-49 in Voigt storage (SymmetricTensor order)
  template <int DIM>
  __device__ __forceinline__ void neo_hooke(const double kappa, const double mu, const double J,
                                            const double (&bbar)[DIM * (DIM + 1) / 2],
                                            double (&tau)[DIM * (DIM + 1) / 2],
                                            double (&D)[DIM * (DIM + 1) / 2][DIM * (DIM + 1) / 2])
  {
    constexpr int VO    = DIM * (DIM + 1) / 2;
    const double  dPsi  = (kappa / 2.0) * (J - 1.0 / J);         // :74-78
    const double  d2Psi = (kappa / 2.0) * (1.0 + 1.0 / (J * J)); // :100-104
    // tau_bar = 2 c_1 b_bar = mu b_bar ; tau_iso = dev_P : tau_bar  (:87-98)
    double tr = 0;
#pragma unroll
    for (int i = 0; i < DIM; ++i)
      tr += mu * bbar[i];
    double tau_iso[VO];
#pragma unroll
    for (int k = 0; k < VO; ++k)
      tau_iso[k] = mu * bbar[k] - (k < DIM ? tr / DIM : 0.0);
    const double tv = dPsi * J; // tau_vol = dPsi J I  (:80-85)
#pragma unroll
    for (int k = 0; k < VO; ++k)
      tau[k] = tau_iso[k] + (k < DIM ? tv : 0.0);
    // Jc_vol = J[(dPsi + J d2Psi) IxI - 2 dPsi S]  (:106-114)
    // Jc_iso = (2/dim) tr(tau_bar) dev_P - (2/dim)(tau_iso x I + I x tau_iso)  (:116-133)
    const double c_IxI = J * (dPsi + J * d2Psi) - (2.0 / DIM) * tr / DIM;
    const double c_S   = -J * (2.0 * dPsi) + (2.0 / DIM) * tr;
#pragma unroll
    for (int k = 0; k < VO; ++k)
#pragma unroll
      for (int l = 0; l < VO; ++l)
        {
          double v = 0;
          if (k < DIM && l < DIM)
            v += c_IxI;
          if (k == l)
            v += (k < DIM ? 1.0 : 0.5) * c_S;
          if (l < DIM)
            v -= (2.0 / DIM) * tau_iso[k];
          if (k < DIM)
            v -= (2.0 / DIM) * tau_iso[l];
          D[k][l] = v;
        }
  }

  inline NLParams make_nl_params(const gf_desc &d)
  {
    NLParams prm;
    prm.kappa   = (2.0 * d.mu * (1.0 + d.nu)) / (3.0 * (1.0 - 2.0 * d.nu)); // material.h:20
    prm.mu      = d.mu;
    prm.rho     = d.rho;
    prm.alpha_1 = 1. / (d.beta * d.delta_t * d.delta_t); // nonlinear_elasticity.h:242
    for (int k = 0; k < 3; ++k)
      prm.body_force[k] = d.body_force[k];
    return prm;
  }
} // namespace gf
