// 1-D operator tables of the Qk Gauss-Lobatto DG basis (host side, computed once per context).
//
// Everything the CUDA kernels need about the discretisation is one-dimensional because the
// mesh is an axis-parallel structured grid and the basis is a tensor product:
//   nodes   : GL nodes on [0,1], ascending        (reference: qkgllocalbasis.hh:222-234,
//                                                   gausslobattomatrices.hh:29-35)
//   M^{ab}  : int_0^1 l^a_i l^b_j                  exact mass, rectangular for a != b
//   S^a     : int_0^1 l^a_i' l^a_j'                exact stiffness
//   t_s,g_s : l_i(s), l_i'(s) at the end points s = 0,1 (face traces)
// and the derived tables used by the kernels' "M^-1-premultiplied" formulation:
//   MinvS = M^-1 S, mt_s = M^-1 t_s, mg_s = M^-1 g_s, P^{ab} = (M^{aa})^-1 M^{ab}.
// Integrals use a 20-point Gauss-Legendre rule evaluated in long double: exact for the
// polynomial degrees involved (<= 2*13), which is what the reference's rules of order 2p
// achieve on affine cubes (ipdgoperator.hh:133,251-253).
#pragma once
#include <vector>

namespace hpdg {

constexpr int kMaxP = 13;          // dynamicqknode.hh:84 : orders 0..13
constexpr int kMaxN = kMaxP + 1;   // table stride

// Device-side mirror: plain arrays with stride kMaxN so a kernel indexes tab[p].M[i*kMaxN+j].
struct DegTable {
  double nodes[kMaxN];
  double M[kMaxN * kMaxN];
  double MinvS[kMaxN * kMaxN];
  double S[kMaxN * kMaxN];
  double Minv[kMaxN * kMaxN];
  double t[2][kMaxN];    // l_i(s)
  double g[2][kMaxN];    // l_i'(s) (unit interval)
  double mt[2][kMaxN];   // M^-1 t_s
  double mg[2][kMaxN];   // M^-1 g_s
};

struct HostTables {
  std::vector<DegTable> deg;            // [kMaxP+1]
  // P[(a*(kMaxP+1)+b)*kMaxN*kMaxN + i*kMaxN + j] = ((M^{aa})^-1 M^{ab})_{ij}, i<=a, j<=b
  std::vector<double> P;
  // T[(c*(kMaxP+1)+f)*kMaxN*kMaxN + i*kMaxN + j] = l^c_j(x^f_i): p-transfer (prolongation) factor
  // (reference: dynamicordertransfer.hh:48-73)
  std::vector<double> T;
  // rectangular mass M^{ab}
  std::vector<double> Mab;
  // Non-conforming (hanging-node) faces, one level of refinement (reference: sfipdg.hh:472-491 evaluates both sides' 1-D bases at
  // the sub-face's mapped quadrature points).  tau = tangential reference coordinate of element e (degree a) on its side, the
  // sub-face covers I subset [0,1]; the neighbour o (degree b) sees it as tau_o(tau).
  //   Pnc_eo[((k*(kMaxP+1)+a)*(kMaxP+1)+b)*kMaxN*kMaxN + i*kMaxN + j] = ((M^{aa})^-1 int_I l^a_i(tau) l^b_j(tau_o(tau)) dtau)_{ij}
  //     k = 0 / 1: e is the COARSE side, I = [0,1/2] / [1/2,1], tau_o = 2 tau - {0,1}
  //     k = 2 / 3: e is the FINE side on the low / high half of o's side, I = [0,1], tau_o = (tau + {0,1}) / 2
  //   Pnc_ee[(k*(kMaxP+1)+a)*kMaxN*kMaxN + i*kMaxN + j] = ((M^{aa})^-1 int_I l^a_i l^a_j dtau)_{ij}, k = 0 / 1 (coarse side halves)
  std::vector<double> Pnc_eo, Pnc_ee;
};

const HostTables& host_tables();

// Generalised symmetric eigenproblem D v = lambda M v (n <= kMaxN), V^T M V = I.
// V is returned row-major n x n with eigenvectors in columns (stride n).
void gen_eig(int n, const double* D, const double* M, double* V, double* lambda);

}  // namespace hpdg
