// C++ host-side test written like the reference's own drivers, against include/hpdg_b200.hh.
//   (1) matrix-free/test/testdg.cc:92-135     : SIPG apply with factor 0.5 vs an independent formulation (the CPU oracle)
//   (2) matrix-free/test/testdgblockjacobi.cc : damped (0.75) block-Jacobi iteration drives the energy norm of the residual < 1e-2
//   (3) test/test_solversetup.cc:25-50        : p-multigrid builds and runs (plus: residual decreases)
//   (4) test/test_dynamicblockgs.cc:24-44     : block-GS on the assembled matrix; (5) the same sweeps matrix-free;
//   (6) matrix-free/test/testoperator.cc:80-98 : operator tuple, on a hanging-node mesh
// Links libhpdg_b200.so (product) and libhpdg_oracle.so (checker; tests may).
#include <cmath>
#include <cstdio>
#include <memory>
#include <vector>

#include "hpdg_b200.hh"
#include "../../oracle/hpdg_oracle.h"

static int fails = 0;
#define CHECK(cond, msg) do { if (!(cond)) { std::printf("FAIL: %s\n", msg); fails++; } else std::printf("ok:   %s\n", msg); } while (0)

int main() {
  using V = hpdg::BlockVector;
  try {
    {  // (1)
      int n[2] = {16, 16}; double L[2] = {1, 1};
      auto ctx = std::make_shared<hpdg::Context>(2, n, L, 2, 2.0, true);
      std::vector<int> deg(256, 2);
      omesh* m = orc_mesh_create(2, n, L, deg.data(), nullptr, 2.0, 1);
      V x = ctx->makeVector(), Ax = ctx->makeVector(), ref = ctx->makeVector();
      orc_interpolate_normsq(m, x.data());
      hpdg::Operator op(ctx);
      op.setFactor(0.5);
      op.apply(x, Ax);
      Ax *= 1 / op.factor();
      orc_apply_mf(m, x.data(), ref.data(), 1.0, 1);
      Ax -= ref;
      CHECK(std::sqrt(Ax * Ax) < 1e-12 * std::sqrt(ref * ref), "Operator::apply with factor matches the quadrature-loop formulation");
      orc_mesh_destroy(m);
    }
    {  // (2)
      int n[2] = {10, 10}; double L[2] = {1, 1};
      auto ctx = std::make_shared<hpdg::Context>(2, n, L, 1, 2.0, true);
      auto A = hpdg::operatorFrom<V>(ctx);
      auto jac = std::make_shared<hpdg::BlockJacobiStep<V>>(ctx, HPDG_FINEST, HPDG_JACOBI_DENSE, 0.75);
      auto smoother = hpdg::smootherFrom<V>(jac);
      V x = ctx->makeVector(), b = ctx->makeVector(), r = ctx->makeVector(), c = ctx->makeVector(), Ac = ctx->makeVector();
      for (std::size_t i = 0; i < b.dimension(); i++) { double v = 1.0 / std::sqrt(100.0); b.data()[i] = v + (i % 5) * v / 10; }
      r = b;
      for (int it = 0; it < 100; it++) {
        A(Ac, c); r -= Ac;
        smoother(c, r);
        x += c;
      }
      A(Ac, r);
      CHECK(r * Ac < 1e-2, "damped block-Jacobi iteration reduces the residual energy below 1e-2");
      // the LinearIterationStep face of the same smoother
      V x2 = ctx->makeVector();
      jac->setProblem(x2, b);
      jac->preprocess();
      for (int it = 0; it < 100; it++) jac->iterate();
      x2 -= x;
      CHECK(std::sqrt(x2 * x2) < 1e-8 * std::sqrt(x * x) + 1e-12, "setProblem/preprocess/iterate agrees with the Smoother form");
    }
    {  // (3)
      int n[3] = {4, 4, 4}; double L[3] = {1, 1, 1};
      auto ctx = std::make_shared<hpdg::Context>(3, n, L, 4, 2.0, true);
      CHECK(ctx->buildPHierarchy() == 3, "p-hierarchy 4 -> 2 -> 1 has three levels");
      auto A = hpdg::operatorFrom<V>(ctx);
      hpdg::Multigrid mg(ctx);
      V x = ctx->makeVector(), b = ctx->makeVector(), t = ctx->makeVector();
      b = 1.0;
      double r0 = std::sqrt(b * b);
      for (int it = 0; it < 15; it++) { V rhs = b; mg.apply(x, rhs); }
      A(t, x); t -= b;
      CHECK(std::sqrt(t * t) < 0.5 * r0, "15 multigrid iterations reduce the residual");
      // transfer hooks
      auto R = hpdg::restrictFrom<V>(ctx, 2); auto Pm = hpdg::prolongFrom<V>(ctx, 2);
      V xc = ctx->makeVector(1), xf = ctx->makeVector(2), yc = ctx->makeVector(1), yf = ctx->makeVector(2);
      for (std::size_t i = 0; i < xc.dimension(); i++) xc.data()[i] = std::sin(0.1 * i);
      for (std::size_t i = 0; i < yf.dimension(); i++) yf.data()[i] = std::cos(0.07 * i);
      Pm(xf, xc); R(yc, yf);
      CHECK(std::fabs(xf * yf - xc * yc) < 1e-10 * std::fabs(xf * yf), "restrict is the transpose of prolong");
    }
    {  // (4) test/test_dynamicblockgs.cc:24-44: 2x2, k=2, penalty 1.5 k^dim (sigma 1.5), b = 1, x0 = 1, 100 sweeps
      int n[2] = {2, 2}; double L[2] = {1, 1};
      auto ctx = std::make_shared<hpdg::Context>(2, n, L, 2, 1.5, true);
      hpdg::AssembledMatrix A(ctx);
      V b = ctx->makeVector(), x = ctx->makeVector(), Ax = ctx->makeVector();
      b = 1.0; x = 1.0;
      hpdg::DynamicBlockGS<V> gs;
      gs.setProblem(A, x, b);
      for (int it = 0; it < 100; it++) gs.iterate();
      A.mv(x, Ax); Ax -= b;
      CHECK(std::sqrt(Ax * Ax) < 1e-13, "100 DynamicBlockGS sweeps solve the small system (|b - Ax| < 1e-13)");
      CHECK(A.blockCol.size() == 12 && A.entries.size() == 12u * 81u, "DynamicBCRSMatrix pattern: 4 rows x (self + 2 neighbours), 9x9 blocks");
    }
    {  // (5) the same sweeps without a matrix: MatrixFreeBlockGS reproduces DynamicBlockGS on the assembled matrix sweep by sweep
      int n[2] = {5, 4}; double L[2] = {1, 1.5};
      std::vector<int> deg(20);
      for (int e = 0; e < 20; e++) deg[e] = 1 + (e * 7) % 4;
      auto ctx = std::make_shared<hpdg::Context>(2, n, L, deg, 2.0, true);
      hpdg::AssembledMatrix A(ctx);
      V b = ctx->makeVector(), x1 = ctx->makeVector(), x2 = ctx->makeVector();
      for (std::size_t i = 0; i < b.dimension(); i++) b.data()[i] = std::sin(0.37 * (double)i);
      x1 = 1.0; x2 = 1.0;
      hpdg::DynamicBlockGS<V> gs;
      gs.setProblem(A, x1, b);
      hpdg::MatrixFreeBlockGS<V> mf(ctx);
      mf.setProblem(x2, b);
      for (int it = 0; it < 5; it++) { gs.iterate(); mf.iterate(); }
      x2 -= x1;
      CHECK(std::sqrt(x2 * x2) < 1e-12 * std::sqrt(x1 * x1), "MatrixFreeBlockGS == DynamicBlockGS on the assembled matrix (hp mesh, 5 sweeps)");
    }
    {  // (6) operator tuple (matrix-free/test/testoperator.cc:80-98: factors 1 and 2 accumulate to 3 A x) and a hanging-node mesh
      int n[2] = {3, 3}; double L[2] = {1, 1};
      std::vector<unsigned char> refine(9, 0);
      refine[4] = 1;                                   // the centre cell is split once
      std::vector<int> deg(9 + 3, 2);
      auto ctx = std::make_shared<hpdg::Context>(hpdg::Context::Refined2D{}, n, L, refine, deg, 2.0, false);
      V x = ctx->makeVector(), y1 = ctx->makeVector(), y3 = ctx->makeVector(), one = ctx->makeVector();
      for (std::size_t i = 0; i < x.dimension(); i++) x.data()[i] = std::cos(0.11 * (double)i);
      hpdg::Operator op(ctx);
      op.apply(x, y1);
      auto tuple = hpdg::Operator::fromLocalOperators(ctx, {hpdg::IPDGOperator(1.0), hpdg::IPDGOperator(2.0)});
      tuple.apply(x, y3);
      y3 *= 1.0 / 3.0; y3 -= y1;
      CHECK(std::sqrt(y3 * y3) < 1e-13 * std::sqrt(y1 * y1), "two local operators with factors 1 and 2 accumulate to 3 A x (hanging-node mesh)");
      one = 1.0;
      op.apply(one, y1);
      CHECK(std::sqrt(y1 * y1) < 1e-11, "natural boundary: constants are in the kernel of the operator on the hanging-node mesh");
      CHECK(ctx->blockOffsets().size() == 13, "leaf elements: 8 unrefined cells + 4 children");
    }
    {  // error behaviour: exceptions, not aborts
      bool threw = false;
      try { int n[1] = {4}; double L[1] = {1}; hpdg::Context bad(1, n, L, 1); } catch (const hpdg::Exception&) { threw = true; }
      CHECK(threw, "invalid arguments raise hpdg::Exception");
    }
  } catch (const std::exception& e) {
    std::printf("FAIL: exception %s\n", e.what());
    fails++;
  }
  std::printf(fails ? "CPP_SHIM FAIL\n" : "CPP_SHIM PASS\n");
  return fails != 0;
}
