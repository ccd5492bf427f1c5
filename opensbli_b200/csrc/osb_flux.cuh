// osb_flux.cuh -- layout of the values a flux sweep stages per grid point.
//
// Per staged point the kernels keep in shared memory (constituent relations evaluated once per point while staging, not once
// per stencil use):  rho, y = 1/sqrt(rho), m_k, E, p, a, u_DIR.
//   * y serves the Roe average without further square roots or divisions: sqrt(rho) = rho y, u = m y^2, and the
//     Roe weights sqrt(rho_R)/(rho_R (sqrt(rho_L)+sqrt(rho_R))) = y_R / (sqrt(rho_L) + sqrt(rho_R))  (averaging.py:62-114);
//   * u_DIR = m_DIR y^2 is the velocity along the sweep: the local wave speeds u_d, u_d +- a and the factor of the
//     characteristic flux  CF_j = u_d CS_j + t_j  (see osb_flux3.cuh).
#pragma once
#include "osb_math.cuh"

namespace osb {

// index of the staged values
template <int ND> struct SV {
  static constexpr int RHO = 0, Y = 1, M0 = 2, E = 2 + ND, P = 3 + ND, A = 4 + ND, UD = 5 + ND, N = 6 + ND;
};

}  // namespace osb
