// jalla_b200.hpp — header-only C++ mirror of Jalla's matrix interface over the C ABI (jalla_b200.h).
// Names and error behaviour follow /root/reference/Numeric/Jalla/Matrix.hs (matrixAlloc :785-787,
// matrixMult :446-461, unsafeMatrixMult :467-490, solveLinearSystem :513-521, invert :537-540):
// Haskell `error` -> std::runtime_error, `Maybe` -> std::optional.
#pragma once
#include <complex>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "jalla_b200.h"

namespace jalla {
enum Transpose { NoTrans = 111, Trans = 112 };
enum Order { RowMajor = 101, ColumnMajor = 102 };

template <typename T> struct Ops;
#define JALLA_OPS(T, CT, t)                                                                                          \
  template <> struct Ops<T> {                                                                                        \
    static void gemm(int o, int ta, int tb, int m, int n, int k, T al, const T *a, int lda, const T *b, int ldb, T be, T *c, int ldc) { \
      GEMM_CALL(t, CT) }                                                                                             \
    static int gesv(int o, int n, int nrhs, T *a, int lda, int *p, T *b, int ldb) { return LAPACKE_##t##gesv(o, n, nrhs, (CT *)a, lda, p, (CT *)b, ldb); } \
    static int getrf(int o, int m, int n, T *a, int lda, int *p) { return LAPACKE_##t##getrf(o, m, n, (CT *)a, lda, p); } \
    static int getri(int o, int n, T *a, int lda, const int *p) { return LAPACKE_##t##getri(o, n, (CT *)a, lda, p); } \
    static int lacpy(int tr, int m, int n, const T *a, int lda, T *b, int ldb) { return jb_##t##lacpy(tr, m, n, (const CT *)a, lda, (CT *)b, ldb); } \
  };
#define GEMM_CALL(t, CT) cblas_##t##gemm(o, ta, tb, m, n, k, al, a, lda, b, ldb, be, c, ldc);
JALLA_OPS(float, float, s)
JALLA_OPS(double, double, d)
#undef GEMM_CALL
#define GEMM_CALL(t, CT) cblas_##t##gemm(o, ta, tb, m, n, k, &al, a, lda, b, ldb, &be, c, ldc);   // complex scalars by pointer (BlasOps.hs:156,180)
JALLA_OPS(std::complex<float>, void, c)
JALLA_OPS(std::complex<double>, void, z)
#undef GEMM_CALL
#undef JALLA_OPS

// GPU-resident counterpart of `data Matrix e` (Matrix.hs:329-332): ColumnMajor, lda = rows, uninitialised.
template <typename T> class Matrix {
 public:
  Matrix(int rows, int cols) : r_(rows), c_(cols), lda_(rows > 1 ? rows : 1) {
    void *p = jb_malloc(sizeof(T) * (size_t)lda_ * (size_t)(cols > 1 ? cols : 1), JB_MEM_DEVICE);
    if (!p) throw std::runtime_error(std::string("matrixAlloc: ") + jb_last_error_string());
    buf_.reset(static_cast<T *>(p), [](T *q) { jb_free(q); });
  }
  static Matrix fromHost(int rows, int cols, const T *colmajor) {
    Matrix m(rows, cols);
    if (rows * cols && jb_memcpy_h2d(m.ptr(), colmajor, sizeof(T) * (size_t)rows * cols)) throw std::runtime_error(jb_last_error_string());
    return m;
  }
  std::vector<T> toHost() const {
    std::vector<T> h((size_t)r_ * c_);
    if (!h.empty() && jb_memcpy_d2h(h.data(), ptr(), sizeof(T) * h.size())) throw std::runtime_error(jb_last_error_string());
    return h;
  }
  int rowCount() const { return r_; }
  int colCount() const { return c_; }
  int lda() const { return lda_; }
  T *ptr() const { return buf_.get(); }
 private:
  int r_, c_, lda_;
  std::shared_ptr<T> buf_;
};

template <typename T>
void unsafeMatrixMult(T alpha, Transpose ta, const Matrix<T> &a, Transpose tb, const Matrix<T> &b, T beta, Matrix<T> &c) {
  int m = ta == NoTrans ? a.rowCount() : a.colCount(), k = ta == NoTrans ? a.colCount() : a.rowCount();
  int n = tb == NoTrans ? b.colCount() : b.rowCount();
  Ops<T>::gemm(ColumnMajor, ta, tb, m, n, k, alpha, a.ptr(), a.lda(), b.ptr(), b.lda(), beta, c.ptr(), c.lda());
  if (int e = jb_last_error()) throw std::runtime_error(std::string("gemm: ") + jb_last_error_string() + " (" + std::to_string(e) + ")");
}
template <typename T> Matrix<T> matrixMult(T alpha, Transpose ta, const Matrix<T> &a, Transpose tb, const Matrix<T> &b) {
  Matrix<T> c(ta == NoTrans ? a.rowCount() : a.colCount(), tb == NoTrans ? b.colCount() : b.rowCount());
  unsafeMatrixMult(alpha, ta, a, tb, b, T(0), c);
  return c;
}
template <typename T> Matrix<T> mm(const Matrix<T> &a, const Matrix<T> &b) {   // a ## b
  if (a.colCount() != b.rowCount()) throw std::runtime_error("(##): shapes do not match");
  return matrixMult(T(1), NoTrans, a, NoTrans, b);
}
template <typename T> Matrix<T> matrixCopy(const Matrix<T> &a, Transpose t = NoTrans) {
  Matrix<T> r(t == NoTrans ? a.rowCount() : a.colCount(), t == NoTrans ? a.colCount() : a.rowCount());
  if (Ops<T>::lacpy(t, a.rowCount(), a.colCount(), a.ptr(), a.lda(), r.ptr(), r.lda())) throw std::runtime_error(jb_last_error_string());
  return r;
}
template <typename T> void unsafeSolveLinearSystem(Matrix<T> &a, Matrix<T> &b) {
  if (!(a.rowCount() == a.colCount() && a.rowCount() == b.rowCount()))
    throw std::runtime_error("unsafeSolveLinearSystem: The shapes of the arguments do not match.");
  std::vector<int> ipiv((size_t)(a.colCount() > 1 ? a.colCount() : 1));
  int ret = Ops<T>::gesv(ColumnMajor, a.colCount(), b.colCount(), a.ptr(), a.lda(), ipiv.data(), b.ptr(), b.lda());
  if (ret != 0) throw std::runtime_error("unsafeSolveLinearSystem: ret /= 0");
}
template <typename T> Matrix<T> solveLinearSystem(const Matrix<T> &a, const Matrix<T> &b) {
  Matrix<T> x = matrixCopy(b), a2 = matrixCopy(a);
  unsafeSolveLinearSystem(a2, x);
  return x;
}
template <typename T> std::optional<Matrix<T>> invert(const Matrix<T> &a) {
  if (a.rowCount() != a.colCount()) return std::nullopt;
  Matrix<T> w = matrixCopy(a);
  std::vector<int> ipiv((size_t)(a.rowCount() > 1 ? a.rowCount() : 1));
  if (Ops<T>::getrf(ColumnMajor, w.rowCount(), w.colCount(), w.ptr(), w.lda(), ipiv.data()) != 0) return std::nullopt;
  if (Ops<T>::getri(ColumnMajor, w.colCount(), w.ptr(), w.lda(), ipiv.data()) != 0) return std::nullopt;
  return w;
}
}  // namespace jalla
