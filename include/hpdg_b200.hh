// hpdg_b200.hh -- header-only C++17 host shim over the C ABI (hpdg_b200.h).
//
// Mirrors the reference-side interfaces of the hot path so that code written against dune-hpdg keeps its shape:
//   * vectors: anything exposing `double* data()` and `size_t dimension()` in the DynamicBlockVector layout
//     (dune/hpdg/common/dynamicbvector.hh:282,328-330,366-379) -- Dune::HPDG::DynamicBlockVector itself qualifies;
//     hpdg::BlockVector below is a minimal stand-in with the same members for builds without DUNE.
//   * hpdg::Operator::apply(x, Ax)            <- Dune::Fufem::MatrixFree::Operator::apply (matrix-free/operator.hh:41-56)
//     with setFactor/factor                    <- LocalOperator (matrix-free/localoperators/localoperator.hh:41-49)
//   * hpdg::BlockJacobiStep: setProblem / preprocess / iterate
//                                              <- Dune::Solvers::LinearIterationStep as used by DynamicBlockGS
//                                                 (iterationsteps/dynamicblockgs.hh:87-127)
//   * hpdg::operatorFrom / smootherFrom / restrictFrom / prolongFrom return
//     std::function<void(Vector&, const Vector&)> <- Operator<V>, Smoother<V>, TransferOperator<V>
//                                                 (iterationsteps/mg/multigrid.hh:13-23,83-155)
//   * hpdg::Multigrid::apply(x, b)             <- Dune::HPDG::Multigrid<Vector>::apply (mg/multigrid_impl.hh:16-61)
// Errors: every non-zero return of the C ABI is rethrown as hpdg::Exception (the reference throws Dune::Exception via
// DUNE_THROW, e.g. dynamicblockgs.hh:117).  No arithmetic lives here.
#pragma once
#include <cstddef>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "hpdg_b200.h"

namespace hpdg {

struct Exception : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// Minimal DynamicBlockVector stand-in: one contiguous array + block offsets.
class BlockVector {
 public:
  BlockVector() = default;
  explicit BlockVector(std::vector<long> offsets) : off_(std::move(offsets)), v_(off_.empty() ? 0 : off_.back(), 0.0) {}
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  std::size_t dimension() const { return v_.size(); }
  std::size_t size() const { return off_.empty() ? 0 : off_.size() - 1; }          // number of blocks
  std::size_t blockRows(std::size_t i) const { return off_[i + 1] - off_[i]; }       // dynamicbvector.hh:134-143
  double* block(std::size_t i) { return v_.data() + off_[i]; }
  BlockVector& operator=(double s) { for (auto& e : v_) e = s; return *this; }       // dynamicbvector.hh:185-192
  BlockVector& operator+=(const BlockVector& o) { for (std::size_t i = 0; i < v_.size(); i++) v_[i] += o.v_[i]; return *this; }
  BlockVector& operator-=(const BlockVector& o) { for (std::size_t i = 0; i < v_.size(); i++) v_[i] -= o.v_[i]; return *this; }
  BlockVector& operator*=(double s) { for (auto& e : v_) e *= s; return *this; }
  double operator*(const BlockVector& o) const { double s = 0; for (std::size_t i = 0; i < v_.size(); i++) s += v_[i] * o.v_[i]; return s; }
  const std::vector<long>& offsets() const { return off_; }

 private:
  std::vector<long> off_;
  std::vector<double> v_;
};

class Context {
 public:
  // uniform degree
  Context(int dim, const int* n, const double* L, int degree, double sigma = 2.0, bool dirichlet = false, int device = 0) {
    check(hpdg_create(&h_, dim, n, L, &degree, 1, sigma, dirichlet ? 1 : 0, device), nullptr);
  }
  // per-element degree map (DynamicDGQkGLBlockBasis(gridView, DegreeMap), dynamicdgqkglbasis.hh:54-69)
  Context(int dim, const int* n, const double* L, const std::vector<int>& degree, double sigma = 2.0, bool dirichlet = false,
          int device = 0) {
    check(hpdg_create(&h_, dim, n, L, degree.data(), (long)degree.size(), sigma, dirichlet ? 1 : 0, device), nullptr);
  }
  // non-conforming 2-D mesh: base grid n with the flagged cells split once into 2 x 2 children (hanging nodes; the grid the
  // reference's non-conforming branch sees, sfipdg.hh:213-222); one degree per leaf element
  struct Refined2D {};
  Context(Refined2D, const int* n, const double* L, const std::vector<unsigned char>& refine, const std::vector<int>& degree,
          double sigma = 2.0, bool dirichlet = false, int device = 0) {
    check(hpdg_create_refined_2d(&h_, n, L, refine.data(), degree.data(), (long)degree.size(), sigma, dirichlet ? 1 : 0, device), nullptr);
  }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  ~Context() { hpdg_destroy(h_); }

  hpdg_ctx* handle() const { return h_; }
  void check(int rc) const { check(rc, h_); }
  int numLevels() const { return hpdg_num_levels(h_); }
  long dimension(int level = HPDG_FINEST) const { return hpdg_dimension(h_, level); }
  std::vector<long> blockOffsets(int level = HPDG_FINEST) const {
    std::vector<long> off(hpdg_num_elements(h_) + 1);
    check(hpdg_block_offsets(h_, level, off.data()));
    return off;
  }
  BlockVector makeVector(int level = HPDG_FINEST) const { return BlockVector(blockOffsets(level)); }
  int buildPHierarchy() { check(hpdg_build_p_hierarchy(h_)); return numLevels(); }   // solversetup.hh:71-108

 private:
  static void check(int rc, const hpdg_ctx* h) {
    if (rc) throw Exception(hpdg_last_error(h));
  }
  hpdg_ctx* h_ = nullptr;
};

// The local operator of the tuple Operator iterates over: carries its own factor (localoperator.hh:41-49: factor() / setFactor())
class IPDGOperator {
 public:
  explicit IPDGOperator(double factor = 1.0) : factor_(factor) {}
  double factor() const { return factor_; }
  void setFactor(double f) { factor_ = f; }

 private:
  double factor_;
};

// Operator::apply(x, Ax) over a tuple of local operators (matrix-free/operator.hh:41-56): Ax is zeroed once, every local operator
// adds factor_k * A x (matrix-free/test/testoperator.cc:80-98: factors 1 and 2 give 3 A x).  The single-operator constructor is the
// common case: Ax = factor * A x.
class Operator {
 public:
  explicit Operator(std::shared_ptr<Context> c, int level = HPDG_FINEST, double factor = 1.0)
      : c_(std::move(c)), level_(level), ops_{IPDGOperator(factor)} {}
  static Operator fromLocalOperators(std::shared_ptr<Context> c, std::vector<IPDGOperator> ops, int level = HPDG_FINEST) {
    Operator o(std::move(c), level);
    if (!ops.empty()) o.ops_ = std::move(ops);
    return o;
  }
  double factor() const { return ops_[0].factor(); }
  void setFactor(double f) { ops_[0].setFactor(f); }
  std::vector<IPDGOperator>& localOperators() { return ops_; }
  template <class V>
  void apply(const V& x, V& Ax) const {
    c_->check(hpdg_op_apply(c_->handle(), level_, x.data(), Ax.data(), ops_[0].factor()));
    for (std::size_t k = 1; k < ops_.size(); k++)
      c_->check(hpdg_op_apply_accum(c_->handle(), level_, x.data(), Ax.data(), ops_[k].factor()));
  }

 private:
  std::shared_ptr<Context> c_;
  int level_;
  std::vector<IPDGOperator> ops_;
};

// LinearIterationStep-shaped damped block-Jacobi step: x += damping * D^-1 (rhs - A x)
template <class V>
class BlockJacobiStep {
 public:
  explicit BlockJacobiStep(std::shared_ptr<Context> c, int level = HPDG_FINEST, int form = HPDG_JACOBI_DENSE, double damping = 1.0)
      : c_(std::move(c)), level_(level), form_(form), damping_(damping) {}
  void setProblem(V& x, const V& rhs) { x_ = &x; rhs_ = &rhs; }
  void preprocess() { c_->check(hpdg_jacobi_setup(c_->handle(), level_, form_)); ready_ = true; }
  void iterate() {
    if (!ready_) preprocess();
    V r = *rhs_, t = *rhs_;
    c_->check(hpdg_op_apply(c_->handle(), level_, x_->data(), t.data(), 1.0));
    r -= t;
    apply(t, r);
    *x_ += t;
  }
  // Smoother<V>(c, r): correction from a zero start
  void apply(V& c, const V& r) {
    if (!ready_) preprocess();
    c_->check(hpdg_jacobi_apply(c_->handle(), level_, form_, r.data(), c.data(), damping_));
  }
  V* x_ = nullptr;         // public members like the dune-solvers step (multigrid.hh:100-103 pokes rhs_)
  const V* rhs_ = nullptr;

 private:
  std::shared_ptr<Context> c_;
  int level_, form_;
  double damping_;
  bool ready_ = false;
};

// DynamicBCRSMatrix-layout matrix assembled on the device (buildingblocks/matrices.hh:29-89, common/dynamicbcrs.hh:178-199).
// blockRowPtr/blockCol/blockOff/entries are exactly the arrays a DynamicBCRSMatrix holds; mv() is BCRSMatrix::mv.
class AssembledMatrix {
 public:
  explicit AssembledMatrix(std::shared_ptr<Context> c, int level = HPDG_FINEST) : c_(std::move(c)), level_(level) {
    long nb = 0, ne = 0;
    c_->check(hpdg_bcrs_sizes(c_->handle(), level_, &nb, &ne));
    blockRowPtr.resize(hpdg_num_elements(c_->handle()) + 1); blockCol.resize(nb); blockOff.resize(nb + 1); entries.resize(ne);
    c_->check(hpdg_assemble_bcrs(c_->handle(), level_, blockRowPtr.data(), blockCol.data(), blockOff.data(), entries.data()));
  }
  template <class V> void mv(const V& x, V& y) const { c_->check(hpdg_bcrs_mv(c_->handle(), level_, x.data(), y.data())); }
  std::vector<long> blockRowPtr, blockOff;
  std::vector<int> blockCol;
  std::vector<double> entries;
  std::shared_ptr<Context> context() const { return c_; }
  int level() const { return level_; }

 private:
  std::shared_ptr<Context> c_;
  int level_;
};

// Dune::HPDG::DynamicBlockGS<Matrix, Vector> (iterationsteps/dynamicblockgs.hh:87-127): setProblem(mat, x, rhs); iterate()
template <class V>
class DynamicBlockGS {
 public:
  void setProblem(const AssembledMatrix& m, V& x, const V& rhs) { mat_ = &m; x_ = &x; rhs_ = &rhs; }
  void preprocess() {}
  void iterate() { mat_->context()->check(hpdg_blockgs_iterate(mat_->context()->handle(), mat_->level(), rhs_->data(), x_->data())); }
  const AssembledMatrix* mat_ = nullptr;
  V* x_ = nullptr;
  const V* rhs_ = nullptr;
};

// The same iteration step without an assembled matrix (hpdg_blockgs_mf_iterate): the context takes the matrix's place in
// setProblem; the sweeps are those of DynamicBlockGS::iterate (dynamicblockgs.hh:94-126) at any mesh size.
template <class V>
class MatrixFreeBlockGS {
 public:
  explicit MatrixFreeBlockGS(std::shared_ptr<Context> c, int level = HPDG_FINEST) : c_(std::move(c)), level_(level) {}
  void setProblem(V& x, const V& rhs) { x_ = &x; rhs_ = &rhs; }
  void preprocess() {}
  void iterate() { c_->check(hpdg_blockgs_mf_iterate(c_->handle(), level_, rhs_->data(), x_->data())); }
  V* x_ = nullptr;
  const V* rhs_ = nullptr;
 private:
  std::shared_ptr<Context> c_;
  int level_;
};

// Dune::HPDG::L1Smoother<Matrix, Vector> (iterationsteps/l1smoother.hh:20-145): L1Smoother(ghosts); setProblem(mat, x, rhs);
// preprocess(); iterate()
template <class V>
class L1Smoother {
 public:
  explicit L1Smoother(const std::vector<std::size_t>& ghosts) : ghosts_(ghosts.begin(), ghosts.end()) {}
  void setProblem(const AssembledMatrix& m, V& x, const V& rhs) { mat_ = &m; x_ = &x; rhs_ = &rhs; }
  void preprocess() { mat_->context()->check(hpdg_l1_setup(mat_->context()->handle(), mat_->level(), ghosts_.data(), (long)ghosts_.size())); }
  void iterate() { mat_->context()->check(hpdg_l1_iterate(mat_->context()->handle(), mat_->level(), rhs_->data(), x_->data())); }
  const AssembledMatrix* mat_ = nullptr;
  V* x_ = nullptr;
  const V* rhs_ = nullptr;
 private:
  std::vector<long> ghosts_;
};

template <class V> using Fn = std::function<void(V&, const V&)>;

template <class V>
Fn<V> operatorFrom(std::shared_ptr<Context> c, int level = HPDG_FINEST) {   // operatorFromMatrix, multigrid.hh:137-155
  return [c, level](V& y, const V& x) { c->check(hpdg_op_apply(c->handle(), level, x.data(), y.data(), 1.0)); };
}
template <class V>
Fn<V> smootherFrom(std::shared_ptr<BlockJacobiStep<V>> step) {              // smootherFromIterationStep, multigrid.hh:83-93
  return [step](V& c, const V& r) { step->apply(c, r); };
}
template <class V>
Fn<V> restrictFrom(std::shared_ptr<Context> c, int fine_level) {            // restrictFromMultigridTransfer, multigrid.hh:108-116
  return [c, fine_level](V& coarse, const V& fine) { c->check(hpdg_restrict(c->handle(), fine_level, fine.data(), coarse.data())); };
}
template <class V>
Fn<V> prolongFrom(std::shared_ptr<Context> c, int fine_level) {             // prolongFromMultigridTransfer, multigrid.hh:118-126
  return [c, fine_level](V& fine, const V& coarse) { c->check(hpdg_prolong(c->handle(), fine_level, coarse.data(), fine.data())); };
}

// Multigrid<Vector>::apply(x, b): x += correction, b := residual
class Multigrid {
 public:
  explicit Multigrid(std::shared_ptr<Context> c, int form = HPDG_JACOBI_FD, double damping = 0.75, int pre = 5, int post = 5,
                     int coarse_its = 5)
      : c_(std::move(c)), form_(form), damping_(damping), pre_(pre), post_(post), coarse_(coarse_its) {}
  template <class V>
  void apply(V& x, V& b) const { c_->check(hpdg_vcycle(c_->handle(), form_, damping_, pre_, post_, coarse_, x.data(), b.data())); }

 private:
  std::shared_ptr<Context> c_;
  int form_;
  double damping_;
  int pre_, post_, coarse_;
};

}  // namespace hpdg
