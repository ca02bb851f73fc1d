// TEST INFRASTRUCTURE — a small, eagerly evaluated stand-in for the part of the Eigen3 API that the reference's vendored
// g2o (Thirdparty/g2o) and src/Optimizer.cc / src/Converter.cc use.  Eigen is not in this image and there is no network,
// so this header is what lets those reference sources compile IN PLACE, unmodified (oracle/Makefile, target
// _ref/libref_optimizer.so).  It is written from the public Eigen 3 API documentation, not from Eigen's sources:
//   * every expression is evaluated at once into a plain Matrix (no expression templates, no vectorisation);
//   * views (Map, Block) are writable windows on existing storage;
//   * products / sums run in index order, so rounding differs from a SIMD Eigen build by a few ulp (the parity bar for
//     LocalBA / PoseOptimization is 1e-5 relative on the deltas);
//   * Quaternion <-> rotation matrix, quaternion product and vector rotation follow Eigen's documented formulas;
//   * SimplicialLDLT is a dense LDL^T (no fill-reducing ordering: the solution is the same up to rounding), LDLT / LLT /
//     PartialPivLU are dense textbook factorisations.
// Nothing under self_commit_orb-slam2_b200/ includes this file.
#ifndef B2S_REFSHIM_EIGEN_H
#define B2S_REFSHIM_EIGEN_H

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <new>
#include <type_traits>
#include <vector>

#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 2
#define EIGEN_MINOR_VERSION 10
#define EIGEN_VERSION_AT_LEAST(x, y, z) \
  (EIGEN_WORLD_VERSION > x || (EIGEN_WORLD_VERSION >= x && (EIGEN_MAJOR_VERSION > y || (EIGEN_MAJOR_VERSION >= y && EIGEN_MINOR_VERSION >= z))))
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW_IF(x)
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)
#define EIGEN_STRONG_INLINE inline
#define EIGEN_DEVICE_FUNC

namespace Eigen {

typedef std::ptrdiff_t DenseIndex;
typedef DenseIndex Index;
const int Dynamic = -1;
const int Infinity = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { Unaligned = 0, Aligned = 1 };
enum { AlignedBit = 0x80 };
enum { Lower = 1, Upper = 2, UnitDiag = 4, ZeroDiag = 8, UnitLower = 5, UnitUpper = 6, StrictlyLower = 9, StrictlyUpper = 10 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
enum { ComputeEigenvectors = 0x80, EigenvaluesOnly = 0x40 };
enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x10 | Affine, Projective = 0x20 };

inline void initParallel() {}
inline void setNbThreads(int) {}

template <class T>
class aligned_allocator : public std::allocator<T> {
 public:
  template <class U>
  struct rebind {
    typedef aligned_allocator<U> other;
  };
  aligned_allocator() {}
  aligned_allocator(const aligned_allocator&) {}
  template <class U>
  aligned_allocator(const aligned_allocator<U>&) {}
};

template <class T>
struct traits;
template <class S, int R, int C, int O = 0, int MR = R, int MC = C>
class Matrix;
template <class M, int Opt = Unaligned, class StrideT = void>
class Map;
template <class X, int BR = Dynamic, int BC = Dynamic>
class Block;
template <class Derived>
class MatrixBase;
template <class X>
class DiagonalView;
template <class X>
class ArrayWrapper;

template <class S, int R, int C, int O, int MR, int MC>
struct traits<Matrix<S, R, C, O, MR, MC> > {
  typedef S Scalar;
  enum { Rows = R, Cols = C };
};
template <class M, int Opt, class St>
struct traits<Map<M, Opt, St> > : traits<typename std::remove_const<M>::type> {};
template <class X, int BR, int BC>
struct traits<Block<X, BR, BC> > {
  typedef typename traits<X>::Scalar Scalar;
  enum { Rows = BR, Cols = BC };
};

template <class X>
struct traits<DiagonalView<X> > {
  typedef typename traits<X>::Scalar Scalar;
  enum { Rows = Dynamic, Cols = 1 };
};
template <class X>
struct traits<ArrayWrapper<X> > : traits<X> {};

namespace internal {
template <int A, int B>
struct pick_dim {
  enum { value = (A != Dynamic) ? A : B };
};
template <class T>
struct plain_of {
  typedef Matrix<typename traits<T>::Scalar, traits<T>::Rows, traits<T>::Cols> type;
};
}  // namespace internal

// ------------------------------------------------------------------------------------------------------------------
// dense factorisations (results of llt() / ldlt() / lu())
template <class MatT>
class LLT;
template <class MatT>
class LDLT;
template <class MatT>
class PartialPivLU;

template <class Derived>
class CommaInitializer {
 public:
  typedef typename traits<Derived>::Scalar Scalar;
  CommaInitializer(Derived& m, Scalar first) : m_(m), k_(0) { put(first); }
  CommaInitializer& operator,(Scalar v) {
    put(v);
    return *this;
  }
  template <class O>
  CommaInitializer& operator,(const MatrixBase<O>& o) {  // only whole-row pieces of vectors are needed here
    for (int i = 0; i < o.size(); ++i) put(o[i]);
    return *this;
  }
  Derived& finished() { return m_; }

 private:
  void put(Scalar v) {
    int c = (int)m_.cols();
    m_.coeffRef(k_ / c, k_ % c) = v;
    ++k_;
  }
  Derived& m_;
  int k_;
};

template <class Derived>
class MatrixBase {
 public:
  typedef typename traits<Derived>::Scalar Scalar;
  typedef Scalar RealScalar;
  typedef DenseIndex Index;
  enum {
    RowsAtCompileTime = traits<Derived>::Rows,
    ColsAtCompileTime = traits<Derived>::Cols,
    SizeAtCompileTime = (traits<Derived>::Rows == Dynamic || traits<Derived>::Cols == Dynamic) ? Dynamic : traits<Derived>::Rows * traits<Derived>::Cols,
    IsVectorAtCompileTime = (traits<Derived>::Rows == 1 || traits<Derived>::Cols == 1),
    Flags = AlignedBit
  };
  typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> PlainObject;
  typedef Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> TransposedPlain;

  Derived& derived() { return *static_cast<Derived*>(this); }
  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  Index rows() const { return derived().rows(); }
  Index cols() const { return derived().cols(); }
  Index size() const { return rows() * cols(); }

  Scalar& operator()(Index i, Index j) { return derived().coeffRef(i, j); }
  const Scalar& operator()(Index i, Index j) const { return derived().coeffRef(i, j); }
  Scalar& operator()(Index i) { return lin(i); }
  const Scalar& operator()(Index i) const { return lin(i); }
  Scalar& operator[](Index i) { return lin(i); }
  const Scalar& operator[](Index i) const { return lin(i); }
  Scalar coeff(Index i, Index j) const { return derived().coeffRef(i, j); }
  Scalar coeff(Index i) const { return lin(i); }
  Scalar& x() { return lin(0); }
  Scalar& y() { return lin(1); }
  Scalar& z() { return lin(2); }
  Scalar& w() { return lin(3); }
  const Scalar& x() const { return lin(0); }
  const Scalar& y() const { return lin(1); }
  const Scalar& z() const { return lin(2); }
  const Scalar& w() const { return lin(3); }

  Derived& noalias() { return derived(); }
  const Derived& eval() const { return derived(); }
  PlainObject matrix() const { return PlainObject(derived()); }
  ArrayWrapper<Derived> array() { return ArrayWrapper<Derived>(derived()); }
  const ArrayWrapper<Derived> array() const { return ArrayWrapper<Derived>(const_cast<Derived&>(derived())); }

  // ---- fills
  Derived& setZero() { return fill(Scalar(0)); }
  Derived& setOnes() { return fill(Scalar(1)); }
  Derived& setConstant(Scalar v) { return fill(v); }
  Derived& fill(Scalar v) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = v;
    return derived();
  }
  Derived& setIdentity() {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = (i == j) ? Scalar(1) : Scalar(0);
    return derived();
  }
  static PlainObject Zero() {
    PlainObject m;
    m.setZero();
    return m;
  }
  static PlainObject Zero(Index r, Index c) {
    PlainObject m(r, c);
    m.setZero();
    return m;
  }
  static PlainObject Zero(Index n) {
    PlainObject m(n);
    m.setZero();
    return m;
  }
  static PlainObject Ones() {
    PlainObject m;
    m.setOnes();
    return m;
  }
  static PlainObject Constant(Scalar v) {
    PlainObject m;
    m.fill(v);
    return m;
  }
  static PlainObject Identity() {
    PlainObject m;
    m.setIdentity();
    return m;
  }
  static PlainObject Identity(Index r, Index c) {
    PlainObject m(r, c);
    m.setIdentity();
    return m;
  }

  // ---- views
  Block<Derived> block(Index r, Index c, Index nr, Index nc) { return Block<Derived>(derived(), r, c, nr, nc); }
  const Block<Derived> block(Index r, Index c, Index nr, Index nc) const {
    return Block<Derived>(const_cast<Derived&>(derived()), r, c, nr, nc);
  }
  template <int NR, int NC>
  Block<Derived, NR, NC> block(Index r, Index c) {
    return Block<Derived, NR, NC>(derived(), r, c, NR, NC);
  }
  template <int NR, int NC>
  const Block<Derived, NR, NC> block(Index r, Index c) const {
    return Block<Derived, NR, NC>(const_cast<Derived&>(derived()), r, c, NR, NC);
  }
  Block<Derived, RowsAtCompileTime, 1> col(Index j) { return Block<Derived, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1); }
  const Block<Derived, RowsAtCompileTime, 1> col(Index j) const {
    return Block<Derived, RowsAtCompileTime, 1>(const_cast<Derived&>(derived()), 0, j, rows(), 1);
  }
  Block<Derived, 1, ColsAtCompileTime> row(Index i) { return Block<Derived, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols()); }
  const Block<Derived, 1, ColsAtCompileTime> row(Index i) const {
    return Block<Derived, 1, ColsAtCompileTime>(const_cast<Derived&>(derived()), i, 0, 1, cols());
  }
  Block<Derived> segment(Index start, Index n) {
    return cols() == 1 ? Block<Derived>(derived(), start, 0, n, 1) : Block<Derived>(derived(), 0, start, 1, n);
  }
  const Block<Derived> segment(Index start, Index n) const { return const_cast<MatrixBase*>(this)->segment(start, n); }
  Block<Derived> head(Index n) { return segment(0, n); }
  const Block<Derived> head(Index n) const { return segment(0, n); }
  Block<Derived> tail(Index n) { return segment(size() - n, n); }
  const Block<Derived> tail(Index n) const { return segment(size() - n, n); }
  template <int N>
  Block<Derived, N, 1> head() {
    return Block<Derived, N, 1>(derived(), 0, 0, N, 1);
  }
  template <int N>
  const Block<Derived, N, 1> head() const {
    return Block<Derived, N, 1>(const_cast<Derived&>(derived()), 0, 0, N, 1);
  }
  template <int N>
  Block<Derived, N, 1> tail() {
    return Block<Derived, N, 1>(derived(), rows() - N, 0, N, 1);
  }
  template <int N>
  const Block<Derived, N, 1> tail() const {
    return Block<Derived, N, 1>(const_cast<Derived&>(derived()), rows() - N, 0, N, 1);
  }
  template <int N>
  Block<Derived, N, 1> segment(Index start) {
    return Block<Derived, N, 1>(derived(), start, 0, N, 1);
  }
  template <int N>
  const Block<Derived, N, 1> segment(Index start) const {
    return Block<Derived, N, 1>(const_cast<Derived&>(derived()), start, 0, N, 1);
  }
  Block<Derived> topLeftCorner(Index nr, Index nc) { return block(0, 0, nr, nc); }
  const Block<Derived> topLeftCorner(Index nr, Index nc) const { return block(0, 0, nr, nc); }
  template <int NR, int NC>
  Block<Derived, NR, NC> topLeftCorner() {
    return block<NR, NC>(0, 0);
  }
  template <int NR, int NC>
  const Block<Derived, NR, NC> topLeftCorner() const {
    return block<NR, NC>(0, 0);
  }
  template <int NR, int NC>
  Block<Derived, NR, NC> topRightCorner() {
    return block<NR, NC>(0, cols() - NC);
  }
  template <int NR, int NC>
  const Block<Derived, NR, NC> topRightCorner() const {
    return block<NR, NC>(0, cols() - NC);
  }
  DiagonalView<Derived> diagonal() { return DiagonalView<Derived>(derived()); }
  const DiagonalView<Derived> diagonal() const { return DiagonalView<Derived>(const_cast<Derived&>(derived())); }

  // ---- element-wise results
  TransposedPlain transpose() const {
    TransposedPlain t(cols(), rows());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) t.coeffRef(j, i) = coeff(i, j);
    return t;
  }
  TransposedPlain adjoint() const { return transpose(); }
  PlainObject operator-() const {
    PlainObject r(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = -coeff(i, j);
    return r;
  }
  PlainObject cwiseAbs() const {
    PlainObject r(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = std::abs(coeff(i, j));
    return r;
  }
  PlainObject cwiseSqrt() const {
    PlainObject r(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = std::sqrt(coeff(i, j));
    return r;
  }
  template <class O>
  PlainObject cwiseProduct(const MatrixBase<O>& o) const {
    PlainObject r(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = coeff(i, j) * o.coeff(i, j);
    return r;
  }
  template <class T>
  Matrix<T, RowsAtCompileTime, ColsAtCompileTime> cast() const {
    Matrix<T, RowsAtCompileTime, ColsAtCompileTime> r(rows(), cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) r.coeffRef(i, j) = (T)coeff(i, j);
    return r;
  }

  // ---- reductions
  Scalar sum() const {
    Scalar s = 0;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) s += coeff(i, j);
    return s;
  }
  Scalar trace() const {
    Scalar s = 0;
    for (Index i = 0; i < std::min(rows(), cols()); ++i) s += coeff(i, i);
    return s;
  }
  Scalar squaredNorm() const {
    Scalar s = 0;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) s += coeff(i, j) * coeff(i, j);
    return s;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  Scalar maxCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) m = std::max(m, coeff(i, j));
    return m;
  }
  Scalar minCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) m = std::min(m, coeff(i, j));
    return m;
  }
  template <int P>
  Scalar lpNorm() const {
    static_assert(P == Infinity || P == 1, "only the L1 / Linf norms are provided");
    Scalar m = 0;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) m = (P == 1) ? m + std::abs(coeff(i, j)) : std::max(m, std::abs(coeff(i, j)));
    return m;
  }
  void normalize() {
    Scalar n = norm();
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) /= n;
  }
  PlainObject normalized() const {
    PlainObject r(derived());
    r.normalize();
    return r;
  }
  template <class O>
  Scalar dot(const MatrixBase<O>& o) const {
    Scalar s = 0;
    for (Index i = 0; i < size(); ++i) s += lin(i) * o[i];
    return s;
  }
  template <class O>
  Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    Matrix<Scalar, 3, 1> r;
    r[0] = lin(1) * o[2] - lin(2) * o[1];
    r[1] = lin(2) * o[0] - lin(0) * o[2];
    r[2] = lin(0) * o[1] - lin(1) * o[0];
    return r;
  }
  bool allFinite() const {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i)
        if (!std::isfinite(coeff(i, j))) return false;
    return true;
  }
  bool hasNaN() const {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i)
        if (std::isnan(coeff(i, j))) return true;
    return false;
  }

  // ---- compound assignment
  template <class O>
  Derived& operator+=(const MatrixBase<O>& o) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) += o.coeff(i, j);
    return derived();
  }
  template <class O>
  Derived& operator-=(const MatrixBase<O>& o) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) -= o.coeff(i, j);
    return derived();
  }
  Derived& operator*=(Scalar s) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) *= s;
    return derived();
  }
  Derived& operator/=(Scalar s) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) /= s;
    return derived();
  }
  template <class O>
  Derived& operator*=(const MatrixBase<O>& o) {
    derived() = derived() * o;
    return derived();
  }
  template <class O>
  bool operator==(const MatrixBase<O>& o) const {
    if (rows() != o.rows() || cols() != o.cols()) return false;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i)
        if (coeff(i, j) != o.coeff(i, j)) return false;
    return true;
  }
  template <class O>
  bool operator!=(const MatrixBase<O>& o) const {
    return !(*this == o);
  }
  template <class O>
  bool isApprox(const MatrixBase<O>& o, Scalar prec = 1e-12) const {
    Scalar d = 0;
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) d += (coeff(i, j) - o.coeff(i, j)) * (coeff(i, j) - o.coeff(i, j));
    return d <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
  }

  CommaInitializer<Derived> operator<<(Scalar first) { return CommaInitializer<Derived>(derived(), first); }

  // ---- small dense algebra
  Scalar determinant() const;
  PlainObject inverse() const;
  LLT<PlainObject> llt() const;
  LDLT<PlainObject> ldlt() const;
  PartialPivLU<PlainObject> lu() const;
  PartialPivLU<PlainObject> partialPivLu() const;

 protected:
  template <class O>
  Derived& assign(const MatrixBase<O>& o) {
    derived().resizeLike(o.rows(), o.cols());
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) derived().coeffRef(i, j) = o.coeff(i, j);
    return derived();
  }

 private:
  Scalar& lin(Index i) const {
    Derived& d = const_cast<Derived&>(derived());
    if (ColsAtCompileTime == 1) return d.coeffRef(i, 0);
    if (RowsAtCompileTime == 1) return d.coeffRef(0, i);
    Index r = d.rows();
    if (d.cols() == 1) return d.coeffRef(i, 0);
    if (r == 1) return d.coeffRef(0, i);
    return d.coeffRef(i % r, i / r);
  }
};

// ------------------------------------------------------------------------------------------------------------------
namespace internal {
template <class S, int R, int C, bool Fixed = (R != Dynamic && C != Dynamic)>
struct DenseStorage;
template <class S, int R, int C>
struct DenseStorage<S, R, C, true> {
  S d[R * C > 0 ? R * C : 1];
  DenseStorage() {}
  DenseStorage(Index, Index) {}
  Index rows() const { return R; }
  Index cols() const { return C; }
  void resize(Index r, Index c) {
    assert(r == R && c == C);
    (void)r;
    (void)c;
  }
  S* data() { return d; }
  const S* data() const { return d; }
};
template <class S, int R, int C>
struct DenseStorage<S, R, C, false> {
  std::vector<S> v;
  Index r_, c_;
  DenseStorage() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C) {}
  DenseStorage(Index r, Index c) : v((size_t)(r * c)), r_(r), c_(c) {}
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  void resize(Index r, Index c) {
    if (r * c != r_ * c_) v.assign((size_t)(r * c), S());
    r_ = r;
    c_ = c;
  }
  S* data() { return v.data(); }
  const S* data() const { return v.data(); }
};
}  // namespace internal

template <class S, int R, int C, int O, int MR, int MC>
class Matrix : public MatrixBase<Matrix<S, R, C, O, MR, MC> > {
 public:
  typedef MatrixBase<Matrix> Base;
  typedef S Scalar;
  typedef Map<Matrix, Unaligned> MapType;
  typedef Map<const Matrix, Unaligned> ConstMapType;
  typedef Map<Matrix, Aligned> AlignedMapType;
  typedef Map<const Matrix, Aligned> ConstAlignedMapType;
  enum { Options = O };

  Matrix() {}
  explicit Matrix(Index n) : st_(R == 1 ? 1 : n, R == 1 ? n : (C == Dynamic ? 1 : C)) {
    if (R != Dynamic && C != Dynamic && R * C == 1) st_.data()[0] = (S)n;  // Matrix<double,1,1>(v)
  }
  // (rows, cols) for a resizable matrix, (x, y) for a fixed 2-vector — one template so that integer arguments are not ambiguous
  template <class T0, class T1>
  Matrix(const T0& a, const T1& b) : st_((R != Dynamic && C != Dynamic) ? R : (Index)a, (R != Dynamic && C != Dynamic) ? C : (Index)b) {
    if (R != Dynamic && C != Dynamic && R * C == 2) {
      st_.data()[0] = (S)a;
      st_.data()[1] = (S)b;
    }
  }
  Matrix(const S& x, const S& y, const S& z) {
    static_assert(R * C == 3, "3-coefficient constructor on a non-3-vector");
    st_.data()[0] = x;
    st_.data()[1] = y;
    st_.data()[2] = z;
  }
  Matrix(const S& x, const S& y, const S& z, const S& w) {
    static_assert(R * C == 4, "4-coefficient constructor on a non-4-vector");
    st_.data()[0] = x;
    st_.data()[1] = y;
    st_.data()[2] = z;
    st_.data()[3] = w;
  }
  explicit Matrix(const S* p) { std::memcpy(st_.data(), p, sizeof(S) * R * C); }
  Matrix(const Matrix& o) : st_(o.st_) {}
  template <class Od>
  Matrix(const MatrixBase<Od>& o) {
    this->assign(o);
  }
  Matrix& operator=(const Matrix& o) {
    st_ = o.st_;
    return *this;
  }
  template <class Od>
  Matrix& operator=(const MatrixBase<Od>& o) {
    return this->assign(o);
  }

  Index rows() const { return st_.rows(); }
  Index cols() const { return st_.cols(); }
  Index innerStride() const { return 1; }
  Index outerStride() const { return rows(); }
  S& coeffRef(Index i, Index j) const { return const_cast<S*>(st_.data())[i + j * st_.rows()]; }
  S* data() { return st_.data(); }
  const S* data() const { return st_.data(); }
  void resize(Index r, Index c) { st_.resize(r, c); }
  void resize(Index n) {
    if (R == 1)
      st_.resize(1, n);
    else
      st_.resize(n, C == Dynamic ? 1 : C);
  }
  void resizeLike(Index r, Index c) {
    if (r != rows() || c != cols()) st_.resize(r, c);
  }
  void conservativeResize(Index r, Index c) {
    Matrix old(*this);
    st_.resize(r, c);
    for (Index j = 0; j < c; ++j)
      for (Index i = 0; i < r; ++i) coeffRef(i, j) = (i < old.rows() && j < old.cols()) ? old.coeffRef(i, j) : S(0);
  }
  void conservativeResize(Index n) {
    if (R == 1)
      conservativeResize(1, n);
    else
      conservativeResize(n, 1);
  }
  Matrix& setZero() { return Base::setZero(); }
  Matrix& setZero(Index n) {
    resize(n);
    return Base::setZero();
  }
  Matrix& setZero(Index r, Index c) {
    resize(r, c);
    return Base::setZero();
  }
  void swap(Matrix& o) { std::swap(st_, o.st_); }

 private:
  internal::DenseStorage<S, R, C> st_;
};

// Map: a column-major window on external memory (placement-new re-seats it, as g2o does for its Hessian blocks)
template <class M, int Opt, class StrideT>
class Map : public MatrixBase<Map<M, Opt, StrideT> > {
 public:
  typedef typename std::remove_const<M>::type Plain;
  typedef typename traits<Plain>::Scalar Scalar;
  typedef typename std::conditional<std::is_const<M>::value, const Scalar*, Scalar*>::type Ptr;
  enum { R = traits<Plain>::Rows, C = traits<Plain>::Cols };
  Map(Ptr p) : p_(const_cast<Scalar*>(p)), r_(R), c_(C) {}
  Map(Ptr p, Index n) : p_(const_cast<Scalar*>(p)), r_(R == 1 ? 1 : n), c_(R == 1 ? n : 1) {}
  Map(Ptr p, Index r, Index c) : p_(const_cast<Scalar*>(p)), r_(r), c_(c) {}
  Map(const Map& o) : p_(o.p_), r_(o.r_), c_(o.c_) {}
  Map& operator=(const Map& o) { return this->assign(o); }
  template <class Od>
  Map& operator=(const MatrixBase<Od>& o) {
    return this->assign(o);
  }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Scalar& coeffRef(Index i, Index j) const { return p_[i + j * r_]; }
  Scalar* data() { return p_; }
  const Scalar* data() const { return p_; }
  void resizeLike(Index r, Index c) {
    assert(r == r_ && c == c_);
    (void)r;
    (void)c;
  }

 private:
  Scalar* p_;
  Index r_, c_;
};

template <class X, int BR, int BC>
class Block : public MatrixBase<Block<X, BR, BC> > {
 public:
  typedef typename traits<X>::Scalar Scalar;
  Block(X& x, Index r0, Index c0, Index nr, Index nc) : x_(&x), r0_(r0), c0_(c0), nr_(nr), nc_(nc) {}
  Block(const Block& o) : x_(o.x_), r0_(o.r0_), c0_(o.c0_), nr_(o.nr_), nc_(o.nc_) {}
  Block& operator=(const Block& o) { return this->assign(typename internal::plain_of<Block>::type(o)); }
  template <class Od>
  Block& operator=(const MatrixBase<Od>& o) {
    return this->assign(o);
  }
  Index rows() const { return nr_; }
  Index cols() const { return nc_; }
  Scalar& coeffRef(Index i, Index j) const { return x_->coeffRef(r0_ + i, c0_ + j); }
  void resizeLike(Index r, Index c) {
    assert(r == nr_ && c == nc_);
    (void)r;
    (void)c;
  }

 private:
  X* x_;
  Index r0_, c0_, nr_, nc_;
};

// writable view of the main diagonal (block_solver.hpp adds lambda through diagonal().array() += lambda)
template <class X>
class DiagonalView : public MatrixBase<DiagonalView<X> > {
 public:
  typedef typename traits<X>::Scalar Scalar;
  explicit DiagonalView(X& x) : x_(&x) {}
  DiagonalView(const DiagonalView& o) : x_(o.x_) {}
  DiagonalView& operator=(const DiagonalView& o) { return this->assign(typename internal::plain_of<DiagonalView>::type(o)); }
  template <class Od>
  DiagonalView& operator=(const MatrixBase<Od>& o) {
    return this->assign(o);
  }
  Index rows() const { return std::min(x_->rows(), x_->cols()); }
  Index cols() const { return 1; }
  Scalar& coeffRef(Index i, Index) const { return x_->coeffRef(i, i); }
  void resizeLike(Index r, Index c) {
    assert(r == rows() && c == 1);
    (void)r;
    (void)c;
  }

 private:
  X* x_;
};
// coefficient-wise ("array") window on an expression: scalar += / -= and coefficient-wise products
template <class X>
class ArrayWrapper : public MatrixBase<ArrayWrapper<X> > {
 public:
  typedef typename traits<X>::Scalar Scalar;
  explicit ArrayWrapper(X& x) : x_(&x) {}
  ArrayWrapper(const ArrayWrapper& o) : x_(o.x_) {}
  Index rows() const { return x_->rows(); }
  Index cols() const { return x_->cols(); }
  Scalar& coeffRef(Index i, Index j) const { return x_->coeffRef(i, j); }
  void resizeLike(Index, Index) {}
  using MatrixBase<ArrayWrapper<X> >::operator+=;
  using MatrixBase<ArrayWrapper<X> >::operator-=;
  ArrayWrapper& operator+=(Scalar s) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) coeffRef(i, j) += s;
    return *this;
  }
  ArrayWrapper& operator-=(Scalar s) {
    for (Index j = 0; j < cols(); ++j)
      for (Index i = 0; i < rows(); ++i) coeffRef(i, j) -= s;
    return *this;
  }
  template <class Od>
  ArrayWrapper& operator=(const MatrixBase<Od>& o) {
    return this->assign(o);
  }
  ArrayWrapper& operator=(const ArrayWrapper& o) { return this->assign(typename internal::plain_of<ArrayWrapper>::type(o)); }

 private:
  X* x_;
};

// ------------------------------------------------------------------------------------------------------------------
// binary operators (eager)
template <class A, class B>
Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value,
       internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value>
operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value,
         internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value>
      r(a.rows(), a.cols());
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value,
       internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value>
operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename traits<A>::Scalar, internal::pick_dim<traits<A>::Rows, traits<B>::Rows>::value,
         internal::pick_dim<traits<A>::Cols, traits<B>::Cols>::value>
      r(a.rows(), a.cols());
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  typedef typename traits<A>::Scalar S;
  Matrix<S, traits<A>::Rows, traits<B>::Cols> r(a.rows(), b.cols());
  assert(a.cols() == b.rows());
  const Index n = a.rows(), m = b.cols(), k = a.cols();
  for (Index j = 0; j < m; ++j)
    for (Index i = 0; i < n; ++i) {
      S s = 0;
      for (Index l = 0; l < k; ++l) s += a.coeff(i, l) * b.coeff(l, j);
      r.coeffRef(i, j) = s;
    }
  return r;
}
template <class A>
typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A>& a, typename traits<A>::Scalar s) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) * s;
  return r;
}
template <class A>
typename MatrixBase<A>::PlainObject operator*(typename traits<A>::Scalar s, const MatrixBase<A>& a) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = s * a.coeff(i, j);
  return r;
}
template <class A>
typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A>& a, typename traits<A>::Scalar s) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) r.coeffRef(i, j) = a.coeff(i, j) / s;
  return r;
}
template <class A>
std::ostream& operator<<(std::ostream& os, const MatrixBase<A>& a) {
  for (Index i = 0; i < a.rows(); ++i) {
    for (Index j = 0; j < a.cols(); ++j) os << (j ? " " : "") << a.coeff(i, j);
    if (i + 1 < a.rows()) os << "\n";
  }
  return os;
}

#define B2S_EIGEN_TYPEDEFS(T, s)                   \
  typedef Matrix<T, 2, 2> Matrix2##s;              \
  typedef Matrix<T, 3, 3> Matrix3##s;              \
  typedef Matrix<T, 4, 4> Matrix4##s;              \
  typedef Matrix<T, Dynamic, Dynamic> MatrixX##s;  \
  typedef Matrix<T, 2, 1> Vector2##s;              \
  typedef Matrix<T, 3, 1> Vector3##s;              \
  typedef Matrix<T, 4, 1> Vector4##s;              \
  typedef Matrix<T, Dynamic, 1> VectorX##s;        \
  typedef Matrix<T, 1, 2> RowVector2##s;           \
  typedef Matrix<T, 1, 3> RowVector3##s;           \
  typedef Matrix<T, 1, 4> RowVector4##s;           \
  typedef Matrix<T, 1, Dynamic> RowVectorX##s;
B2S_EIGEN_TYPEDEFS(double, d)
B2S_EIGEN_TYPEDEFS(float, f)
B2S_EIGEN_TYPEDEFS(int, i)
#undef B2S_EIGEN_TYPEDEFS

// ------------------------------------------------------------------------------------------------------------------
// dense factorisations
template <class MatT>
class LLT {
 public:
  typedef typename traits<MatT>::Scalar S;
  LLT() : ok_(false) {}
  template <class A>
  explicit LLT(const MatrixBase<A>& a) {
    compute(a);
  }
  template <class A>
  LLT& compute(const MatrixBase<A>& a) {
    const Index n = a.rows();
    L_ = MatT(a);
    ok_ = true;
    for (Index j = 0; j < n; ++j) {
      S d = L_(j, j);
      for (Index k = 0; k < j; ++k) d -= L_(j, k) * L_(j, k);
      if (!(d > S(0))) {
        ok_ = false;
        return *this;
      }
      d = std::sqrt(d);
      L_(j, j) = d;
      for (Index i = j + 1; i < n; ++i) {
        S s = L_(i, j);
        for (Index k = 0; k < j; ++k) s -= L_(i, k) * L_(j, k);
        L_(i, j) = s / d;
      }
    }
    for (Index j = 0; j < n; ++j)
      for (Index i = 0; i < j; ++i) L_(i, j) = 0;
    return *this;
  }
  template <class B>
  typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
    typename MatrixBase<B>::PlainObject x(b);
    const Index n = L_.rows();
    for (Index c = 0; c < x.cols(); ++c) {
      for (Index i = 0; i < n; ++i) {
        S s = x(i, c);
        for (Index k = 0; k < i; ++k) s -= L_(i, k) * x(k, c);
        x(i, c) = s / L_(i, i);
      }
      for (Index i = n - 1; i >= 0; --i) {
        S s = x(i, c);
        for (Index k = i + 1; k < n; ++k) s -= L_(k, i) * x(k, c);
        x(i, c) = s / L_(i, i);
      }
    }
    return x;
  }
  const MatT& matrixL() const { return L_; }
  MatT matrixU() const { return L_.transpose(); }
  ComputationInfo info() const { return ok_ ? Success : NumericalIssue; }

 private:
  MatT L_;
  bool ok_;
};

// LDL^T without pivoting (Eigen pivots; for the SPD systems g2o hands over the solution agrees to rounding)
template <class MatT>
class LDLT {
 public:
  typedef typename traits<MatT>::Scalar S;
  LDLT() : ok_(false), positive_(false) {}
  template <class A>
  explicit LDLT(const MatrixBase<A>& a) {
    compute(a);
  }
  template <class A>
  LDLT& compute(const MatrixBase<A>& a) {
    const Index n = a.rows();
    L_ = MatT(a);
    ok_ = true;
    positive_ = true;
    for (Index j = 0; j < n; ++j) {
      S d = L_(j, j);
      for (Index k = 0; k < j; ++k) d -= L_(j, k) * L_(j, k) * L_(k, k);
      L_(j, j) = d;
      if (!(d > S(0))) positive_ = false;
      if (d == S(0) || !std::isfinite(d)) {
        ok_ = false;
        return *this;
      }
      for (Index i = j + 1; i < n; ++i) {
        S s = L_(i, j);
        for (Index k = 0; k < j; ++k) s -= L_(i, k) * L_(j, k) * L_(k, k);
        L_(i, j) = s / d;
      }
    }
    return *this;
  }
  template <class B>
  typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
    typename MatrixBase<B>::PlainObject x(b);
    const Index n = L_.rows();
    for (Index c = 0; c < x.cols(); ++c) {
      for (Index i = 0; i < n; ++i) {
        S s = x(i, c);
        for (Index k = 0; k < i; ++k) s -= L_(i, k) * x(k, c);
        x(i, c) = s;
      }
      for (Index i = 0; i < n; ++i) x(i, c) /= L_(i, i);
      for (Index i = n - 1; i >= 0; --i) {
        S s = x(i, c);
        for (Index k = i + 1; k < n; ++k) s -= L_(k, i) * x(k, c);
        x(i, c) = s;
      }
    }
    return x;
  }
  bool isPositive() const { return ok_ && positive_; }
  bool isNegative() const { return ok_ && !positive_; }
  ComputationInfo info() const { return ok_ ? Success : NumericalIssue; }
  Matrix<S, Dynamic, 1> vectorD() const { return Matrix<S, Dynamic, 1>(L_.diagonal()); }

 private:
  MatT L_;
  bool ok_, positive_;
};

template <class MatT>
class PartialPivLU {
 public:
  typedef typename traits<MatT>::Scalar S;
  PartialPivLU() : sign_(1) {}
  template <class A>
  explicit PartialPivLU(const MatrixBase<A>& a) {
    compute(a);
  }
  template <class A>
  PartialPivLU& compute(const MatrixBase<A>& a) {
    const Index n = a.rows();
    lu_ = MatT(a);
    perm_.resize((size_t)n);
    sign_ = 1;
    for (Index i = 0; i < n; ++i) perm_[(size_t)i] = i;
    for (Index k = 0; k < n; ++k) {
      Index p = k;
      S best = std::abs(lu_(k, k));
      for (Index i = k + 1; i < n; ++i)
        if (std::abs(lu_(i, k)) > best) {
          best = std::abs(lu_(i, k));
          p = i;
        }
      if (p != k) {
        for (Index j = 0; j < n; ++j) std::swap(lu_(k, j), lu_(p, j));
        std::swap(perm_[(size_t)k], perm_[(size_t)p]);
        sign_ = -sign_;
      }
      if (lu_(k, k) == S(0)) continue;
      for (Index i = k + 1; i < n; ++i) {
        lu_(i, k) /= lu_(k, k);
        for (Index j = k + 1; j < n; ++j) lu_(i, j) -= lu_(i, k) * lu_(k, j);
      }
    }
    return *this;
  }
  template <class B>
  typename MatrixBase<B>::PlainObject solve(const MatrixBase<B>& b) const {
    const Index n = lu_.rows();
    typename MatrixBase<B>::PlainObject x(b.rows(), b.cols());
    for (Index c = 0; c < b.cols(); ++c) {
      for (Index i = 0; i < n; ++i) {
        S s = b.coeff(perm_[(size_t)i], c);
        for (Index k = 0; k < i; ++k) s -= lu_(i, k) * x(k, c);
        x(i, c) = s;
      }
      for (Index i = n - 1; i >= 0; --i) {
        S s = x(i, c);
        for (Index k = i + 1; k < n; ++k) s -= lu_(i, k) * x(k, c);
        x(i, c) = s / lu_(i, i);
      }
    }
    return x;
  }
  S determinant() const {
    S d = (S)sign_;
    for (Index i = 0; i < lu_.rows(); ++i) d *= lu_(i, i);
    return d;
  }
  MatT inverse() const {
    MatT I(lu_.rows(), lu_.cols());
    I.setIdentity();
    return solve(I);
  }

 private:
  MatT lu_;
  std::vector<Index> perm_;
  int sign_;
};
template <class MatT>
class FullPivLU : public PartialPivLU<MatT> {
 public:
  FullPivLU() {}
  template <class A>
  explicit FullPivLU(const MatrixBase<A>& a) : PartialPivLU<MatT>(a) {}
};

template <class D>
typename MatrixBase<D>::Scalar MatrixBase<D>::determinant() const {
  const Index n = rows();
  if (n == 1) return coeff(0, 0);
  if (n == 2) return coeff(0, 0) * coeff(1, 1) - coeff(1, 0) * coeff(0, 1);
  if (n == 3)
    return coeff(0, 0) * (coeff(1, 1) * coeff(2, 2) - coeff(1, 2) * coeff(2, 1)) -
           coeff(0, 1) * (coeff(1, 0) * coeff(2, 2) - coeff(1, 2) * coeff(2, 0)) +
           coeff(0, 2) * (coeff(1, 0) * coeff(2, 1) - coeff(1, 1) * coeff(2, 0));
  return PartialPivLU<PlainObject>(derived()).determinant();
}
template <class D>
typename MatrixBase<D>::PlainObject MatrixBase<D>::inverse() const {
  const Index n = rows();
  PlainObject r(n, n);
  if (n == 1) {
    r(0, 0) = Scalar(1) / coeff(0, 0);
    return r;
  }
  if (n == 2) {
    const Scalar invdet = Scalar(1) / determinant();
    r(0, 0) = coeff(1, 1) * invdet;
    r(1, 0) = -coeff(1, 0) * invdet;
    r(0, 1) = -coeff(0, 1) * invdet;
    r(1, 1) = coeff(0, 0) * invdet;
    return r;
  }
  if (n == 3) {  // cofactors of the first column give the determinant, the rest follow (fixed-size 3x3 inverse)
    auto cof = [&](int i1, int j1, int i2, int j2) { return coeff(i1, j1) * coeff(i2, j2) - coeff(i1, j2) * coeff(i2, j1); };
    const Scalar c00 = cof(1, 1, 2, 2), c10 = cof(2, 1, 0, 2), c20 = cof(0, 1, 1, 2);
    // note cof(2,1,0,2) = m21*m02 - m22*m01
    const Scalar det = coeff(0, 0) * c00 + coeff(1, 0) * c10 + coeff(2, 0) * c20;
    const Scalar invdet = Scalar(1) / det;
    r(0, 0) = c00 * invdet;
    r(0, 1) = c10 * invdet;
    r(0, 2) = c20 * invdet;
    r(1, 0) = cof(1, 2, 2, 0) * invdet;
    r(1, 1) = cof(0, 0, 2, 2) * invdet;
    r(1, 2) = cof(1, 0, 0, 2) * invdet;
    r(2, 0) = cof(1, 0, 2, 1) * invdet;
    r(2, 1) = cof(2, 0, 0, 1) * invdet;
    r(2, 2) = cof(0, 0, 1, 1) * invdet;
    return r;
  }
  return PartialPivLU<PlainObject>(derived()).inverse();
}
template <class D>
LLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::llt() const {
  return LLT<PlainObject>(derived());
}
template <class D>
LDLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::ldlt() const {
  return LDLT<PlainObject>(derived());
}
template <class D>
PartialPivLU<typename MatrixBase<D>::PlainObject> MatrixBase<D>::lu() const {
  return PartialPivLU<PlainObject>(derived());
}
template <class D>
PartialPivLU<typename MatrixBase<D>::PlainObject> MatrixBase<D>::partialPivLu() const {
  return PartialPivLU<PlainObject>(derived());
}

// symmetric eigenvalues by cyclic Jacobi (only g2o's information-matrix diagnostics use it)
template <class MatT>
class SelfAdjointEigenSolver {
 public:
  typedef typename traits<MatT>::Scalar S;
  SelfAdjointEigenSolver() {}
  template <class A>
  explicit SelfAdjointEigenSolver(const MatrixBase<A>& a, int = ComputeEigenvectors) {
    compute(a);
  }
  template <class A>
  SelfAdjointEigenSolver& compute(const MatrixBase<A>& a, int = ComputeEigenvectors) {
    const Index n = a.rows();
    Matrix<S, Dynamic, Dynamic> m(a);
    for (int sweep = 0; sweep < 64; ++sweep) {
      S off = 0;
      for (Index p = 0; p < n; ++p)
        for (Index q = p + 1; q < n; ++q) off += m(p, q) * m(p, q);
      if (off < std::numeric_limits<S>::min()) break;
      for (Index p = 0; p < n; ++p)
        for (Index q = p + 1; q < n; ++q) {
          if (m(p, q) == S(0)) continue;
          S theta = (m(q, q) - m(p, p)) / (2 * m(p, q));
          S t = (theta >= 0 ? S(1) : S(-1)) / (std::abs(theta) + std::sqrt(theta * theta + 1));
          S c = 1 / std::sqrt(t * t + 1), s = t * c;
          for (Index k = 0; k < n; ++k) {
            S akp = m(k, p), akq = m(k, q);
            m(k, p) = c * akp - s * akq;
            m(k, q) = s * akp + c * akq;
          }
          for (Index k = 0; k < n; ++k) {
            S apk = m(p, k), aqk = m(q, k);
            m(p, k) = c * apk - s * aqk;
            m(q, k) = s * apk + c * aqk;
          }
        }
    }
    ev_.resize(n);
    for (Index i = 0; i < n; ++i) ev_[i] = m(i, i);
    std::sort(ev_.data(), ev_.data() + n);
    return *this;
  }
  const Matrix<S, Dynamic, 1>& eigenvalues() const { return ev_; }
  ComputationInfo info() const { return Success; }

 private:
  Matrix<S, Dynamic, 1> ev_;
};

// ------------------------------------------------------------------------------------------------------------------
// Geometry
template <class S, int Opt = 0>
class Quaternion {
 public:
  typedef S Scalar;
  typedef Matrix<S, 4, 1> Coefficients;
  typedef Matrix<S, 3, 1> Vector3;
  typedef Matrix<S, 3, 3> Matrix3;
  Quaternion() {}
  Quaternion(const S& w, const S& x, const S& y, const S& z) { c_ << x, y, z, w; }
  Quaternion(const Quaternion& o) : c_(o.c_) {}
  explicit Quaternion(const S* d) { c_ << d[0], d[1], d[2], d[3]; }
  template <class D>
  explicit Quaternion(const MatrixBase<D>& m) {
    *this = m;
  }
  Quaternion& operator=(const Quaternion& o) {
    c_ = o.c_;
    return *this;
  }
  // rotation matrix (3x3) or coefficient vector (4x1: x, y, z, w)
  template <class D>
  Quaternion& operator=(const MatrixBase<D>& m) {
    if (m.rows() == 4 && m.cols() == 1) {
      for (int i = 0; i < 4; ++i) c_[i] = m[i];
      return *this;
    }
    S t = m.coeff(0, 0) + m.coeff(1, 1) + m.coeff(2, 2);
    if (t > S(0)) {
      t = std::sqrt(t + S(1.0));
      w() = S(0.5) * t;
      t = S(0.5) / t;
      x() = (m.coeff(2, 1) - m.coeff(1, 2)) * t;
      y() = (m.coeff(0, 2) - m.coeff(2, 0)) * t;
      z() = (m.coeff(1, 0) - m.coeff(0, 1)) * t;
    } else {
      int i = 0;
      if (m.coeff(1, 1) > m.coeff(0, 0)) i = 1;
      if (m.coeff(2, 2) > m.coeff(i, i)) i = 2;
      int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m.coeff(i, i) - m.coeff(j, j) - m.coeff(k, k) + S(1.0));
      c_[i] = S(0.5) * t;
      t = S(0.5) / t;
      w() = (m.coeff(k, j) - m.coeff(j, k)) * t;
      c_[j] = (m.coeff(j, i) + m.coeff(i, j)) * t;
      c_[k] = (m.coeff(k, i) + m.coeff(i, k)) * t;
    }
    return *this;
  }
  static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
  Quaternion& setIdentity() {
    c_ << 0, 0, 0, 1;
    return *this;
  }
  S& x() { return c_[0]; }
  S& y() { return c_[1]; }
  S& z() { return c_[2]; }
  S& w() { return c_[3]; }
  const S& x() const { return c_[0]; }
  const S& y() const { return c_[1]; }
  const S& z() const { return c_[2]; }
  const S& w() const { return c_[3]; }
  Coefficients& coeffs() { return c_; }
  const Coefficients& coeffs() const { return c_; }
  Vector3 vec() const { return Vector3(c_[0], c_[1], c_[2]); }
  S squaredNorm() const { return c_.squaredNorm(); }
  S norm() const { return c_.norm(); }
  void normalize() { c_.normalize(); }
  Quaternion normalized() const {
    Quaternion q(*this);
    q.normalize();
    return q;
  }
  Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
  Quaternion inverse() const {
    S n2 = squaredNorm();
    if (n2 > S(0)) return Quaternion(w() / n2, -x() / n2, -y() / n2, -z() / n2);
    return Quaternion(0, 0, 0, 0);
  }
  S dot(const Quaternion& o) const { return c_.dot(o.c_); }
  Quaternion operator*(const Quaternion& b) const {
    const Quaternion& a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Quaternion& operator*=(const Quaternion& b) {
    *this = *this * b;
    return *this;
  }
  // v' = v + w*t + q_xyz x t, t = 2 * (q_xyz x v)
  template <class D>
  Vector3 operator*(const MatrixBase<D>& v) const {
    Vector3 u = vec();
    Vector3 uv = u.cross(v);
    uv += uv;
    Vector3 uuv = u.cross(uv);
    Vector3 r;
    for (int i = 0; i < 3; ++i) r[i] = v[i] + w() * uv[i] + uuv[i];
    return r;
  }
  Vector3 _transformVector(const Vector3& v) const { return *this * v; }
  Matrix3 toRotationMatrix() const {
    Matrix3 res;
    const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
    const S twx = tx * w(), twy = ty * w(), twz = tz * w();
    const S txx = tx * x(), txy = ty * x(), txz = tz * x();
    const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res(0, 0) = S(1) - (tyy + tzz);
    res(0, 1) = txy - twz;
    res(0, 2) = txz + twy;
    res(1, 0) = txy + twz;
    res(1, 1) = S(1) - (txx + tzz);
    res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy;
    res(2, 1) = tyz + twx;
    res(2, 2) = S(1) - (txx + tyy);
    return res;
  }
  Matrix3 matrix() const { return toRotationMatrix(); }
  template <class T>
  Quaternion<T> cast() const {
    return Quaternion<T>((T)w(), (T)x(), (T)y(), (T)z());
  }

 private:
  Coefficients c_;
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class S>
class AngleAxis {
 public:
  typedef Matrix<S, 3, 1> Vector3;
  AngleAxis() : a_(0), ax_(1, 0, 0) {}
  template <class D>
  AngleAxis(S a, const MatrixBase<D>& ax) : a_(a), ax_(ax) {}
  S angle() const { return a_; }
  const Vector3& axis() const { return ax_; }
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> res;
    S s = std::sin(a_), c = std::cos(a_);
    Vector3 sin_axis = s * ax_;
    Vector3 cos1_axis = (S(1) - c) * ax_;
    S tmp;
    tmp = cos1_axis.x() * ax_.y();
    res(0, 1) = tmp - sin_axis.z();
    res(1, 0) = tmp + sin_axis.z();
    tmp = cos1_axis.x() * ax_.z();
    res(0, 2) = tmp + sin_axis.y();
    res(2, 0) = tmp - sin_axis.y();
    tmp = cos1_axis.y() * ax_.z();
    res(1, 2) = tmp - sin_axis.x();
    res(2, 1) = tmp + sin_axis.x();
    res(0, 0) = cos1_axis.x() * ax_.x() + c;
    res(1, 1) = cos1_axis.y() * ax_.y() + c;
    res(2, 2) = cos1_axis.z() * ax_.z() + c;
    return res;
  }

 private:
  S a_;
  Vector3 ax_;
};
typedef AngleAxis<double> AngleAxisd;

template <class S, int Dim, int Mode, int Opt = 0>
class Transform {
 public:
  typedef Matrix<S, Dim + 1, Dim + 1> MatrixType;
  typedef Matrix<S, Dim, Dim> LinearMatrixType;
  typedef Matrix<S, Dim, 1> VectorType;
  Transform() { m_.setIdentity(); }
  Transform(const Transform& o) : m_(o.m_) {}
  template <int O2>
  Transform(const Quaternion<S, O2>& q) {
    m_.setIdentity();
    m_.template block<Dim, Dim>(0, 0) = q.toRotationMatrix();
  }
  template <class D>
  explicit Transform(const MatrixBase<D>& m) {
    m_.setIdentity();
    if (m.rows() == Dim)
      m_.template block<Dim, Dim>(0, 0) = m;
    else
      m_ = m;
  }
  static Transform Identity() { return Transform(); }
  void setIdentity() { m_.setIdentity(); }
  MatrixType& matrix() { return m_; }
  const MatrixType& matrix() const { return m_; }
  Block<MatrixType, Dim, Dim> linear() { return m_.template block<Dim, Dim>(0, 0); }
  const Block<MatrixType, Dim, Dim> linear() const { return m_.template block<Dim, Dim>(0, 0); }
  Block<MatrixType, Dim, Dim> rotation() { return linear(); }
  const Block<MatrixType, Dim, Dim> rotation() const { return linear(); }
  Block<MatrixType, Dim, 1> translation() { return m_.template block<Dim, 1>(0, Dim); }
  const Block<MatrixType, Dim, 1> translation() const { return m_.template block<Dim, 1>(0, Dim); }
  S& operator()(Index i, Index j) { return m_(i, j); }
  const S& operator()(Index i, Index j) const { return m_(i, j); }
  Transform operator*(const Transform& o) const {
    Transform r;
    r.m_ = m_ * o.m_;
    return r;
  }
  template <class D>
  VectorType operator*(const MatrixBase<D>& v) const {
    VectorType r = LinearMatrixType(linear()) * v;
    for (int i = 0; i < Dim; ++i) r[i] += m_(i, Dim);
    return r;
  }
  Transform inverse() const {
    Transform r;
    if (Mode == Isometry) {
      LinearMatrixType Rt = LinearMatrixType(linear()).transpose();
      r.linear() = Rt;
      r.translation() = -(Rt * VectorType(translation()));
    } else {
      r.m_ = m_.inverse();
    }
    return r;
  }

 private:
  MatrixType m_;
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<double, 2, Isometry> Isometry2d;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 2, Affine> Affine2d;

// ------------------------------------------------------------------------------------------------------------------
// the little of Eigen/Sparse + Eigen/SparseCholesky that g2o's LinearSolverEigen touches
template <class S, class I = int>
class Triplet {
 public:
  Triplet() : r_(0), c_(0), v_(0) {}
  Triplet(const I& r, const I& c, const S& v = S(0)) : r_(r), c_(c), v_(v) {}
  const I& row() const { return r_; }
  const I& col() const { return c_; }
  const S& value() const { return v_; }

 private:
  I r_, c_;
  S v_;
};

template <class IndexT>
struct PermIndices {
  std::vector<IndexT> v;
  IndexT& operator()(Index i) { return v[(size_t)i]; }
  const IndexT& operator()(Index i) const { return v[(size_t)i]; }
  IndexT& operator[](Index i) { return v[(size_t)i]; }
  const IndexT& operator[](Index i) const { return v[(size_t)i]; }
  Index size() const { return (Index)v.size(); }
};
template <int SizeAtCompileTime, int MaxSize = SizeAtCompileTime, class IndexT = int>
class PermutationMatrix {
 public:
  PermutationMatrix() {}
  explicit PermutationMatrix(Index n) { resize(n); }
  void resize(Index n) { idx_.v.assign((size_t)n, 0); }
  Index size() const { return idx_.size(); }
  Index rows() const { return size(); }
  Index cols() const { return size(); }
  PermIndices<IndexT>& indices() { return idx_; }
  const PermIndices<IndexT>& indices() const { return idx_; }
  void setIdentity(Index n) {
    resize(n);
    for (Index i = 0; i < n; ++i) idx_.v[(size_t)i] = (IndexT)i;
  }
  PermutationMatrix inverse() const {
    PermutationMatrix r(size());
    for (Index i = 0; i < size(); ++i) r.idx_.v[(size_t)idx_.v[(size_t)i]] = (IndexT)i;
    return r;
  }

 private:
  PermIndices<IndexT> idx_;
};

template <class S, int Opt = ColMajor, class I = int>
class SparseMatrix;
template <class SM, unsigned UpLo>
struct SparseSelfAdjointView;
template <class SM, unsigned UpLo>
struct SparseSymmetricPermutationProduct {
  const SM* m;
  const PermutationMatrix<Dynamic, Dynamic>* p;
};
template <class SM, unsigned UpLo>
struct SparseSelfAdjointView {
  SM* m;
  SparseSymmetricPermutationProduct<SM, UpLo> twistedBy(const PermutationMatrix<Dynamic, Dynamic>& p) const {
    return SparseSymmetricPermutationProduct<SM, UpLo>{m, &p};
  }
  // dst(upper) = P * src(sym) * P^-1 with the convention new index = p.indices()(old index)
  template <unsigned SrcUpLo>
  SparseSelfAdjointView& operator=(const SparseSymmetricPermutationProduct<SM, SrcUpLo>& prod);
};

// compressed column storage; only what LinearSolverEigen needs (triplet build, raw value pointer, symmetric views)
template <class S, int Opt, class I>
class SparseMatrix {
 public:
  typedef S Scalar;
  typedef I Index;
  SparseMatrix() : r_(0), c_(0) { outer_.assign(1, 0); }
  SparseMatrix(Eigen::Index r, Eigen::Index c) { resize(r, c); }
  void resize(Eigen::Index r, Eigen::Index c) {
    r_ = r;
    c_ = c;
    outer_.assign((size_t)c + 1, 0);
    inner_.clear();
    val_.clear();
  }
  Eigen::Index rows() const { return r_; }
  Eigen::Index cols() const { return c_; }
  Eigen::Index nonZeros() const { return (Eigen::Index)val_.size(); }
  S* valuePtr() { return val_.data(); }
  const S* valuePtr() const { return val_.data(); }
  I* innerIndexPtr() { return inner_.data(); }
  const I* innerIndexPtr() const { return inner_.data(); }
  I* outerIndexPtr() { return outer_.data(); }
  const I* outerIndexPtr() const { return outer_.data(); }
  // column-major, rows ascending inside a column, duplicates summed (documented behaviour of setFromTriplets)
  template <class It>
  void setFromTriplets(It begin, It end) {
    std::vector<std::pair<std::pair<I, I>, S> > t;
    for (It it = begin; it != end; ++it) t.push_back(std::make_pair(std::make_pair((I)it->col(), (I)it->row()), it->value()));
    std::stable_sort(t.begin(), t.end(),
                     [](const std::pair<std::pair<I, I>, S>& a, const std::pair<std::pair<I, I>, S>& b) { return a.first < b.first; });
    outer_.assign((size_t)c_ + 1, 0);
    inner_.clear();
    val_.clear();
    for (size_t k = 0; k < t.size(); ++k) {
      if (k > 0 && t[k].first == t[k - 1].first) {
        val_.back() += t[k].second;
        continue;
      }
      inner_.push_back(t[k].first.second);
      val_.push_back(t[k].second);
      outer_[(size_t)t[k].first.first + 1]++;
    }
    for (size_t j = 0; j < (size_t)c_; ++j) outer_[j + 1] += outer_[j];
  }
  template <unsigned UpLo>
  SparseSelfAdjointView<SparseMatrix, UpLo> selfadjointView() {
    return SparseSelfAdjointView<SparseMatrix, UpLo>{this};
  }
  template <unsigned UpLo>
  SparseSelfAdjointView<SparseMatrix, UpLo> selfadjointView() const {
    return SparseSelfAdjointView<SparseMatrix, UpLo>{const_cast<SparseMatrix*>(this)};
  }
  // C = A.selfadjointView<Upper>()  (full symmetric copy)
  template <unsigned UpLo>
  SparseMatrix& operator=(const SparseSelfAdjointView<SparseMatrix, UpLo>& v) {
    const SparseMatrix& a = *v.m;
    std::vector<Triplet<S, I> > t;
    for (Eigen::Index j = 0; j < a.c_; ++j)
      for (I k = a.outer_[(size_t)j]; k < a.outer_[(size_t)j + 1]; ++k) {
        I i = a.inner_[(size_t)k];
        bool keep = (UpLo == Upper) ? (i <= (I)j) : (i >= (I)j);
        if (!keep) continue;
        t.push_back(Triplet<S, I>(i, (I)j, a.val_[(size_t)k]));
        if (i != (I)j) t.push_back(Triplet<S, I>((I)j, i, a.val_[(size_t)k]));
      }
    resize(a.r_, a.c_);
    setFromTriplets(t.begin(), t.end());
    return *this;
  }
  const SparseMatrix& nestedExpression() const { return *this; }

 private:
  Eigen::Index r_, c_;
  std::vector<I> outer_, inner_;
  std::vector<S> val_;
  template <class, unsigned>
  friend struct SparseSelfAdjointView;
};

template <class SM, unsigned UpLo>
template <unsigned SrcUpLo>
SparseSelfAdjointView<SM, UpLo>& SparseSelfAdjointView<SM, UpLo>::operator=(const SparseSymmetricPermutationProduct<SM, SrcUpLo>& prod) {
  typedef typename SM::Scalar S;
  typedef typename SM::Index I;
  const SM& a = *prod.m;
  std::vector<Triplet<S, I> > t;
  for (Eigen::Index j = 0; j < a.cols(); ++j)
    for (I k = a.outerIndexPtr()[j]; k < a.outerIndexPtr()[j + 1]; ++k) {
      I i = a.innerIndexPtr()[k];
      bool keep = (SrcUpLo == Upper) ? (i <= (I)j) : (i >= (I)j);
      if (!keep) continue;
      I pi = (I)prod.p->indices()(i), pj = (I)prod.p->indices()(j);
      if (UpLo == Upper ? (pi > pj) : (pi < pj)) std::swap(pi, pj);
      t.push_back(Triplet<S, I>(pi, pj, a.valuePtr()[k]));
    }
  m->resize(a.rows(), a.cols());
  m->setFromTriplets(t.begin(), t.end());
  return *this;
}

namespace internal {
// no fill-reducing ordering in the stand-in: the identity permutation (a dense factorisation does not need one)
template <class SM, class Perm>
void minimum_degree_ordering(SM& C, Perm& perm) {
  perm.setIdentity(C.cols());
}
}  // namespace internal

// hook for tests: called with (n, dense column-major symmetric matrix) at every numeric factorisation
typedef void (*b2s_factorize_hook_t)(int n, const double* dense);
inline b2s_factorize_hook_t& b2s_factorize_hook() {
  static b2s_factorize_hook_t h = nullptr;
  return h;
}

template <class SM, int UpLoT = Lower>
class SimplicialLDLT {
 public:
  typedef SM MatrixType;
  typedef SM CholMatrixType;
  typedef typename SM::Scalar Scalar;
  typedef Matrix<Scalar, Dynamic, 1> VectorType;
  enum { UpLo = UpLoT };
  SimplicialLDLT() : info_(Success), n_(0), analyzed_(false) {}
  ComputationInfo info() const { return info_; }
  void analyzePattern(const SM& a) {
    m_P.setIdentity(a.cols());
    m_Pinv.setIdentity(a.cols());
    analyzed_ = true;
  }
  void analyzePattern_preordered(const SM&, bool) { analyzed_ = true; }
  void compute(const SM& a) {
    analyzePattern(a);
    factorize(a);
  }
  void factorize(const SM& a) {
    n_ = a.cols();
    const Index n = n_;
    if (m_P.size() != n) {
      m_P.setIdentity(n);
      m_Pinv.setIdentity(n);
    }
    L_.assign((size_t)(n * n), Scalar(0));
    // dense symmetric copy of the stored triangle, in permuted order (new index = P(old index))
    for (Index j = 0; j < n; ++j)
      for (int k = a.outerIndexPtr()[j]; k < a.outerIndexPtr()[j + 1]; ++k) {
        Index i = a.innerIndexPtr()[k];
        bool keep = (UpLoT == Upper) ? (i <= j) : (i >= j);
        if (!keep) continue;
        Index pi = m_P.indices()(i), pj = m_P.indices()(j);
        L_[(size_t)(pi + pj * n)] = a.valuePtr()[k];
        L_[(size_t)(pj + pi * n)] = a.valuePtr()[k];
      }
    if (b2s_factorize_hook()) b2s_factorize_hook()((int)n, L_.data());
    info_ = Success;
    // column-oriented LDL^T on the lower triangle; row i only reaches back to its first non-zero (envelope), which is
    // what keeps a banded covisibility system cheap without a symbolic phase
    first_.assign((size_t)n, 0);
    for (Index i = 0; i < n; ++i) {
      Index f = 0;
      while (f < i && L_[(size_t)(i + f * n)] == Scalar(0)) ++f;
      first_[(size_t)i] = f;
    }
    for (Index j = 0; j < n; ++j) {
      Scalar d = L_[(size_t)(j + j * n)];
      for (Index k = first_[(size_t)j]; k < j; ++k) d -= L_[(size_t)(j + k * n)] * L_[(size_t)(j + k * n)] * L_[(size_t)(k + k * n)];
      L_[(size_t)(j + j * n)] = d;
      if (d == Scalar(0)) {
        info_ = NumericalIssue;
        return;
      }
      for (Index i = j + 1; i < n; ++i) {
        if (first_[(size_t)i] > j) continue;
        Scalar s = L_[(size_t)(i + j * n)];
        for (Index k = std::max(first_[(size_t)i], first_[(size_t)j]); k < j; ++k)
          s -= L_[(size_t)(i + k * n)] * L_[(size_t)(j + k * n)] * L_[(size_t)(k + k * n)];
        L_[(size_t)(i + j * n)] = s / d;
      }
    }
  }
  template <class B>
  VectorType solve(const MatrixBase<B>& b) const {
    const Index n = n_;
    VectorType y(n);
    for (Index i = 0; i < n; ++i) y[m_P.indices()(i)] = b[i];
    for (Index i = 0; i < n; ++i) {
      Scalar s = y[i];
      for (Index k = first_[(size_t)i]; k < i; ++k) s -= L_[(size_t)(i + k * n)] * y[k];
      y[i] = s;
    }
    for (Index i = 0; i < n; ++i) y[i] /= L_[(size_t)(i + i * n)];
    for (Index i = n - 1; i >= 0; --i) {
      Scalar s = y[i];
      for (Index k = i + 1; k < n; ++k)
        if (first_[(size_t)k] <= i) s -= L_[(size_t)(k + i * n)] * y[k];
      y[i] = s;
    }
    VectorType x(n);
    for (Index i = 0; i < n; ++i) x[i] = y[m_P.indices()(i)];
    return x;
  }
  struct LView {
    Index nnz;
    const LView& nestedExpression() const { return *this; }
    Index nonZeros() const { return nnz; }
  };
  LView matrixL() const {
    Index nnz = 0;
    for (Index i = 0; i < n_; ++i) nnz += i - first_[(size_t)i];
    return LView{nnz};
  }
  const PermutationMatrix<Dynamic, Dynamic>& permutationP() const { return m_P; }
  const PermutationMatrix<Dynamic, Dynamic>& permutationPinv() const { return m_Pinv; }

 protected:
  ComputationInfo info_;
  Index n_;
  bool analyzed_;
  std::vector<Scalar> L_;
  std::vector<Index> first_;
  // (names as in Eigen: g2o's CholeskyDecomposition subclass assigns them)
  PermutationMatrix<Dynamic, Dynamic> m_P, m_Pinv;
};
template <class SM, int UpLoT = Lower>
class SimplicialLLT : public SimplicialLDLT<SM, UpLoT> {};

}  // namespace Eigen

namespace std {
// Eigen/StdVector makes std::vector<T, Eigen::aligned_allocator<T>> usable; nothing to do for the stand-in
}

#endif  // B2S_REFSHIM_EIGEN_H
