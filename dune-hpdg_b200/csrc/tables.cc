// Host-side construction of the 1-D tables (see tables.hpp).  Product code: independent of oracle/.
#include "tables.hpp"

#include <cmath>
#include <cstring>
#include <mutex>

namespace hpdg {
namespace {
typedef long double ld;
constexpr int kQ = 20;  // Gauss points: exact to degree 39

// Node and quadrature generation.  Deliberately NOT the Newton-on-Legendre iteration the CPU test oracle uses:
// nodes and weights come from the eigen-decomposition of the Jacobi matrices of the orthogonal polynomials (Golub-Welsch), the
// Lagrange basis is evaluated in barycentric form.  A mistake in either generator then shows up as a product/oracle mismatch.

// cyclic Jacobi rotations on a symmetric n x n matrix (row-major, stride n): A -> diagonal, Q's columns -> eigenvectors
void jacobi_eig(int n, ld* A, ld* Q) {
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Q[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 200; sweep++) {
    ld offd = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) offd += A[i * n + j] * A[i * n + j];
    if (offd < 1e-70L) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      const ld apq = A[p * n + q];
      if (fabsl(apq) < 1e-300L) continue;
      const ld theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
      const ld t = (theta >= 0 ? 1 : -1) / (fabsl(theta) + sqrtl(theta * theta + 1));
      const ld c = 1 / sqrtl(t * t + 1), sn = t * c;
      for (int k = 0; k < n; k++) { ld akp = A[k * n + p], akq = A[k * n + q]; A[k * n + p] = c * akp - sn * akq; A[k * n + q] = sn * akp + c * akq; }
      for (int k = 0; k < n; k++) { ld apk = A[p * n + k], aqk = A[q * n + k]; A[p * n + k] = c * apk - sn * aqk; A[q * n + k] = sn * apk + c * aqk; }
      for (int k = 0; k < n; k++) { ld qkp = Q[k * n + p], qkq = Q[k * n + q]; Q[k * n + p] = c * qkp - sn * qkq; Q[k * n + q] = sn * qkp + c * qkq; }
    }
  }
}

// eigenvalues (ascending) of the symmetric tridiagonal matrix with zero diagonal and off-diagonal entries off[0..n-2];
// z0[i] = first component of the i-th normalised eigenvector
void tridiag_eig(int n, const ld* off, ld* ev, ld* z0) {
  static thread_local ld A[kQ * kQ], Q[kQ * kQ];
  for (int i = 0; i < n * n; i++) A[i] = 0;
  for (int i = 0; i + 1 < n; i++) A[i * n + i + 1] = A[(i + 1) * n + i] = off[i];
  jacobi_eig(n, A, Q);
  int idx[kQ];
  for (int i = 0; i < n; i++) idx[i] = i;
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (A[idx[j] * n + idx[j]] < A[idx[i] * n + idx[i]]) std::swap(idx[i], idx[j]);
  for (int i = 0; i < n; i++) { ev[i] = A[idx[i] * n + idx[i]]; z0[i] = Q[0 * n + idx[i]]; }
}

// Gauss-Lobatto nodes on [0,1] (qkgllocalbasis.hh:222-234): the end points and the zeros of P'_p, i.e. of the Jacobi polynomial
// P^(1,1)_{p-1}, whose Jacobi matrix has zero diagonal and off-diagonal entries sqrt(k (k+2) / ((2k+1)(2k+3))), k = 1..p-2
void gl_nodes(int p, ld* x) {
  if (p == 0) { x[0] = 0.5L; return; }
  x[0] = 0; x[p] = 1;
  if (p >= 2) {
    ld off[kMaxN], ev[kMaxN], z0[kMaxN];
    for (int k = 1; k <= p - 2; k++) off[k - 1] = sqrtl((ld)k * (k + 2) / ((ld)(2 * k + 1) * (2 * k + 3)));
    tridiag_eig(p - 1, off, ev, z0);
    for (int i = 1; i < p; i++) x[i] = (ev[i - 1] + 1) / 2;
  }
  for (int i = 0; i <= p / 2; i++) { ld a = (x[i] + (1 - x[p - i])) / 2; x[i] = a; x[p - i] = 1 - a; }  // exact mirror symmetry
}

// m-point Gauss-Legendre rule on [0,1] (Golub-Welsch): off-diagonal entries k / sqrt(4 k^2 - 1), weights = first components squared
void gauss(int m, ld* x, ld* w) {
  ld off[kQ], ev[kQ], z0[kQ];
  for (int k = 1; k < m; k++) off[k - 1] = (ld)k / sqrtl(4 * (ld)k * k - 1);
  tridiag_eig(m, off, ev, z0);
  for (int i = 0; i < m; i++) { x[i] = (ev[i] + 1) / 2; w[i] = z0[i] * z0[i]; }
}

// Lagrange basis on the nodes nd[0..p] in barycentric form: l_i(x) = (b_i / (x - x_i)) / sum_j b_j / (x - x_j)
void bary_weights(int p, const ld* nd, ld* b) {
  for (int i = 0; i <= p; i++) { ld d = 1; for (int j = 0; j <= p; j++) if (j != i) d *= (nd[i] - nd[j]); b[i] = 1 / d; }
}
ld lag(int p, const ld* nd, int i, ld x) {
  ld b[kMaxN];
  bary_weights(p, nd, b);
  for (int j = 0; j <= p; j++) if (x == nd[j]) return i == j ? 1 : 0;
  ld den = 0;
  for (int j = 0; j <= p; j++) den += b[j] / (x - nd[j]);
  return b[i] / (x - nd[i]) / den;
}
// l_i'(x): at a node x_k != x_i it is (b_i / b_k) / (x_k - x_i); otherwise l_i(x) * sum_{j != i} 1 / (x - x_j)
ld dlag(int p, const ld* nd, int i, ld x) {
  ld b[kMaxN];
  bary_weights(p, nd, b);
  for (int k = 0; k <= p; k++) if (x == nd[k] && k != i) return b[i] / b[k] / (nd[k] - nd[i]);
  ld sum = 0;
  for (int j = 0; j <= p; j++) if (j != i) sum += 1 / (x - nd[j]);
  return lag(p, nd, i, x) * sum;
}

// in-place inverse by Gauss-Jordan with partial pivoting, n x n, stride n
void invert(int n, ld* A) {
  ld B[kMaxN * kMaxN];
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) B[i * n + j] = (i == j);
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++) if (fabsl(A[r * n + c]) > fabsl(A[piv * n + c])) piv = r;
    if (piv != c) for (int j = 0; j < n; j++) { std::swap(A[c * n + j], A[piv * n + j]); std::swap(B[c * n + j], B[piv * n + j]); }
    ld d = 1 / A[c * n + c];
    for (int j = 0; j < n; j++) { A[c * n + j] *= d; B[c * n + j] *= d; }
    for (int r = 0; r < n; r++) if (r != c) {
      ld f = A[r * n + c];
      if (f != 0) for (int j = 0; j < n; j++) { A[r * n + j] -= f * A[c * n + j]; B[r * n + j] -= f * B[c * n + j]; }
    }
  }
  memcpy(A, B, sizeof(ld) * n * n);
}

HostTables build() {
  HostTables H;
  const int ND = kMaxP + 1;
  H.deg.resize(ND);
  H.P.assign((size_t)ND * ND * kMaxN * kMaxN, 0.0);
  H.T.assign((size_t)ND * ND * kMaxN * kMaxN, 0.0);
  H.Mab.assign((size_t)ND * ND * kMaxN * kMaxN, 0.0);
  ld qx[kQ], qw[kQ];
  gauss(kQ, qx, qw);
  static ld nodes[ND][kMaxN];
  static ld val[ND][kQ][kMaxN], der[ND][kQ][kMaxN];
  static ld Minv[ND][kMaxN * kMaxN];
  for (int p = 0; p < ND; p++) {
    gl_nodes(p, nodes[p]);
    for (int q = 0; q < kQ; q++) for (int i = 0; i <= p; i++) {
      val[p][q][i] = lag(p, nodes[p], i, qx[q]);
      der[p][q][i] = dlag(p, nodes[p], i, qx[q]);
    }
  }
  for (int p = 0; p < ND; p++) {
    int n = p + 1;
    DegTable& D = H.deg[p];
    memset(&D, 0, sizeof(D));
    ld M[kMaxN * kMaxN], S[kMaxN * kMaxN];
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      ld m = 0, s = 0;
      for (int q = 0; q < kQ; q++) { m += qw[q] * val[p][q][i] * val[p][q][j]; s += qw[q] * der[p][q][i] * der[p][q][j]; }
      M[i * n + j] = m; S[i * n + j] = s;
    }
    ld Mi[kMaxN * kMaxN];
    memcpy(Mi, M, sizeof(ld) * n * n);
    invert(n, Mi);
    memcpy(Minv[p], Mi, sizeof(ld) * n * n);
    ld t[2][kMaxN], g[2][kMaxN];
    for (int s = 0; s < 2; s++) for (int i = 0; i < n; i++) { t[s][i] = lag(p, nodes[p], i, (ld)s); g[s][i] = dlag(p, nodes[p], i, (ld)s); }
    for (int i = 0; i < n; i++) {
      D.nodes[i] = (double)nodes[p][i];
      for (int j = 0; j < n; j++) {
        D.M[i * kMaxN + j] = (double)M[i * n + j];
        D.S[i * kMaxN + j] = (double)S[i * n + j];
        D.Minv[i * kMaxN + j] = (double)Mi[i * n + j];
        ld ms = 0;
        for (int k = 0; k < n; k++) ms += Mi[i * n + k] * S[k * n + j];
        D.MinvS[i * kMaxN + j] = (double)ms;
      }
      for (int s = 0; s < 2; s++) {
        D.t[s][i] = (double)t[s][i]; D.g[s][i] = (double)g[s][i];
        ld a = 0, b = 0;
        for (int k = 0; k < n; k++) { a += Mi[i * n + k] * t[s][k]; b += Mi[i * n + k] * g[s][k]; }
        D.mt[s][i] = (double)a; D.mg[s][i] = (double)b;
      }
    }
  }
  for (int a = 0; a < ND; a++) for (int b = 0; b < ND; b++) {
    int na = a + 1, nb = b + 1;
    ld Mab[kMaxN * kMaxN];
    for (int i = 0; i < na; i++) for (int j = 0; j < nb; j++) {
      ld m = 0;
      for (int q = 0; q < kQ; q++) m += qw[q] * val[a][q][i] * val[b][q][j];
      Mab[i * nb + j] = m;
    }
    double* P = &H.P[((size_t)a * ND + b) * kMaxN * kMaxN];
    double* T = &H.T[((size_t)a * ND + b) * kMaxN * kMaxN];
    double* Mo = &H.Mab[((size_t)a * ND + b) * kMaxN * kMaxN];
    for (int i = 0; i < na; i++) for (int j = 0; j < nb; j++) {
      ld s = 0;
      for (int k = 0; k < na; k++) s += Minv[a][i * na + k] * Mab[k * nb + j];
      P[i * kMaxN + j] = (a == b) ? (double)(i == j) : (double)s;
      Mo[i * kMaxN + j] = (double)Mab[i * nb + j];
    }
    // T[c=a][f=b]: rows = fine nodes (b), cols = coarse functions (a)
    for (int i = 0; i < nb; i++) for (int j = 0; j < na; j++)
      T[i * kMaxN + j] = (a == b) ? (double)(i == j) : (double)lag(a, nodes[a], j, nodes[b][i]);
  }
  // non-conforming faces: sub-interval coupling of the tangential 1-D bases, exact 20-point Gauss quadrature on the sub-face
  H.Pnc_eo.assign((size_t)4 * ND * ND * kMaxN * kMaxN, 0.0);
  H.Pnc_ee.assign((size_t)2 * ND * kMaxN * kMaxN, 0.0);
  for (int k = 0; k < 4; k++) for (int a = 0; a < ND; a++) for (int b = 0; b < ND; b++) {
    const int na = a + 1, nb = b + 1;
    ld Meo[kMaxN * kMaxN];
    for (int i = 0; i < na; i++) for (int j = 0; j < nb; j++) {
      ld m = 0;
      for (int q = 0; q < kQ; q++) {
        const ld t = qx[q];   // sub-face coordinate in [0,1]
        ld te, to, w;         // the two sides' tangential coordinates, d tau_e / dt
        if (k < 2) { te = (t + k) / 2; to = t; w = (ld)0.5; } else { te = t; to = (t + (k - 2)) / 2; w = 1; }
        m += qw[q] * w * lag(a, nodes[a], i, te) * lag(b, nodes[b], j, to);
      }
      Meo[i * nb + j] = m;
    }
    double* P = &H.Pnc_eo[(((size_t)k * ND + a) * ND + b) * kMaxN * kMaxN];
    for (int i = 0; i < na; i++) for (int j = 0; j < nb; j++) {
      ld s = 0;
      for (int l = 0; l < na; l++) s += Minv[a][i * na + l] * Meo[l * nb + j];
      P[i * kMaxN + j] = (double)s;
    }
  }
  for (int k = 0; k < 2; k++) for (int a = 0; a < ND; a++) {
    const int na = a + 1;
    ld Mee[kMaxN * kMaxN];
    for (int i = 0; i < na; i++) for (int j = 0; j < na; j++) {
      ld m = 0;
      for (int q = 0; q < kQ; q++) { const ld te = (qx[q] + k) / 2; m += qw[q] * (ld)0.5 * lag(a, nodes[a], i, te) * lag(a, nodes[a], j, te); }
      Mee[i * na + j] = m;
    }
    double* P = &H.Pnc_ee[((size_t)k * ND + a) * kMaxN * kMaxN];
    for (int i = 0; i < na; i++) for (int j = 0; j < na; j++) {
      ld s = 0;
      for (int l = 0; l < na; l++) s += Minv[a][i * na + l] * Mee[l * na + j];
      P[i * kMaxN + j] = (double)s;
    }
  }
  return H;
}
}  // namespace

const HostTables& host_tables() {
  static HostTables H = build();
  return H;
}

void gen_eig(int n, const double* Din, const double* Min, double* V, double* lambda) {
  // M = L L^T ; C = L^-1 D L^-T ; cyclic Jacobi on C ; V = L^-T Q
  ld L[kMaxN * kMaxN] = {0}, C[kMaxN * kMaxN], Q[kMaxN * kMaxN];
  for (int j = 0; j < n; j++) {
    ld d = Min[j * n + j];
    for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
    d = sqrtl(d); L[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      ld s = Min[i * n + j];
      for (int k = 0; k < j; k++) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / d;
    }
  }
  // X = L^-1 D  (forward substitution on columns), then C = X L^-T
  ld X[kMaxN * kMaxN];
  for (int c = 0; c < n; c++)
    for (int i = 0; i < n; i++) {
      ld s = Din[i * n + c];
      for (int k = 0; k < i; k++) s -= L[i * n + k] * X[k * n + c];
      X[i * n + c] = s / L[i * n + i];
    }
  for (int r = 0; r < n; r++)
    for (int i = 0; i < n; i++) {
      ld s = X[r * n + i];
      for (int k = 0; k < i; k++) s -= L[i * n + k] * C[r * n + k];
      C[r * n + i] = s / L[i * n + i];
    }
  for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) { ld a = (C[i * n + j] + C[j * n + i]) / 2; C[i * n + j] = C[j * n + i] = a; }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Q[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 100; sweep++) {
    ld offd = 0;
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) offd += C[i * n + j] * C[i * n + j];
    if (offd < 1e-60L) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      ld apq = C[p * n + q];
      if (fabsl(apq) < 1e-300L) continue;
      ld theta = (C[q * n + q] - C[p * n + p]) / (2 * apq);
      ld t = (theta >= 0 ? 1 : -1) / (fabsl(theta) + sqrtl(theta * theta + 1));
      ld c = 1 / sqrtl(t * t + 1), s = t * c;
      for (int k = 0; k < n; k++) {
        ld akp = C[k * n + p], akq = C[k * n + q];
        C[k * n + p] = c * akp - s * akq; C[k * n + q] = s * akp + c * akq;
      }
      for (int k = 0; k < n; k++) {
        ld apk = C[p * n + k], aqk = C[q * n + k];
        C[p * n + k] = c * apk - s * aqk; C[q * n + k] = s * apk + c * aqk;
      }
      for (int k = 0; k < n; k++) {
        ld qkp = Q[k * n + p], qkq = Q[k * n + q];
        Q[k * n + p] = c * qkp - s * qkq; Q[k * n + q] = s * qkp + c * qkq;
      }
    }
  }
  // V = L^-T Q : back substitution
  for (int c = 0; c < n; c++) {
    ld v[kMaxN];
    for (int i = n - 1; i >= 0; i--) {
      ld s = Q[i * n + c];
      for (int k = i + 1; k < n; k++) s -= L[k * n + i] * v[k];
      v[i] = s / L[i * n + i];
    }
    for (int i = 0; i < n; i++) V[i * n + c] = (double)v[i];
    lambda[c] = (double)C[c * n + c];
  }
}

}  // namespace hpdg
