// Exercises include/jalla_b200.hpp — the C++ mirror of Jalla's matrix interface — against naive host loops.
// Built and run by tests/test_cpp_mirror.py on the GPU box.  Exit code 0 = all checks passed.
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>
#include "jalla_b200.hpp"

template <typename T> static double mag(T v) { return std::abs(v); }

template <typename T>
static std::vector<T> rnd(int n, std::mt19937 &g) {
  std::uniform_real_distribution<double> d(-1, 1);
  std::vector<T> v((size_t)n);
  for (auto &x : v) {
    if constexpr (std::is_same<T, std::complex<double>>::value || std::is_same<T, std::complex<float>>::value)
      x = T((typename T::value_type)d(g), (typename T::value_type)d(g));
    else x = (T)d(g);
  }
  return v;
}

template <typename T>
static int check_type(const char *name, double tol) {
  using namespace jalla;
  std::mt19937 g(7);
  const int m = 70, k = 45, n = 33;
  auto ha = rnd<T>(m * k, g), hb = rnd<T>(k * n, g);
  auto A = Matrix<T>::fromHost(m, k, ha.data()), B = Matrix<T>::fromHost(k, n, hb.data());
  auto C = mm(A, B).toHost();                                   // a ## b
  double err = 0, ref = 0;
  for (int j = 0; j < n; j++)
    for (int i = 0; i < m; i++) {
      T s = 0;
      for (int p = 0; p < k; p++) s += ha[i + p * m] * hb[p + j * k];
      err = std::max(err, mag(T(C[i + j * m] - s))); ref = std::max(ref, mag(s));
    }
  if (!(err <= tol * k * ref)) { std::printf("%s: gemm error %g\n", name, err); return 1; }
  auto At = matrixMult(T(1), Trans, A, NoTrans, A).toHost();      // (a,Trans) ##! (a,NoTrans)
  for (int j = 0; j < k; j++)
    for (int i = 0; i < k; i++) {
      T s = 0;
      for (int p = 0; p < m; p++) s += ha[p + i * m] * ha[p + j * m];
      if (!(mag(T(At[i + j * k] - s)) <= tol * m * (1 + mag(s)))) { std::printf("%s: A^T A error\n", name); return 1; }
    }
  bool threw = false;
  try { mm(A, A); } catch (const std::runtime_error &) { threw = true; }   // shape mismatch -> error
  if (!threw) { std::printf("%s: shape check missing\n", name); return 1; }
  // solveLinearSystem / invert on a well conditioned square system
  const int q = 48;
  auto hs = rnd<T>(q * q, g), hr = rnd<T>(q * 2, g);
  for (int i = 0; i < q; i++) hs[i + i * q] += T(q);
  auto S = Matrix<T>::fromHost(q, q, hs.data()), R = Matrix<T>::fromHost(q, 2, hr.data());
  auto X = solveLinearSystem(S, R).toHost();
  if (S.toHost() != hs) { std::printf("%s: solveLinearSystem clobbered its input\n", name); return 1; }
  for (int j = 0; j < 2; j++)
    for (int i = 0; i < q; i++) {
      T s = 0;
      for (int p = 0; p < q; p++) s += hs[i + p * q] * X[p + j * q];
      if (!(mag(T(s - hr[i + j * q])) <= tol * q * q)) { std::printf("%s: solve residual %g\n", name, mag(T(s - hr[i + j * q]))); return 1; }
    }
  auto inv = invert(S);
  if (!inv) { std::printf("%s: invert returned nothing\n", name); return 1; }
  auto I = mm(S, *inv).toHost();
  for (int j = 0; j < q; j++)
    for (int i = 0; i < q; i++)
      if (!(mag(T(I[i + j * q] - T(i == j ? 1 : 0))) <= tol * q * q)) { std::printf("%s: A*inv(A) != I\n", name); return 1; }
  if (invert(A)) { std::printf("%s: invert of a non-square matrix must be empty\n", name); return 1; }
  std::vector<T> ones((size_t)q * q, T(1));
  if (invert(Matrix<T>::fromHost(q, q, ones.data()))) { std::printf("%s: invert of a singular matrix must be empty\n", name); return 1; }
  return 0;
}

int main() {
  if (jb_init(-1) != 0) { std::printf("no GPU: %s\n", jb_last_error_string()); return 77; }
  int bad = 0;
  bad += check_type<float>("float", 2e-6);
  bad += check_type<double>("double", 4e-15);
  bad += check_type<std::complex<float>>("complex<float>", 4e-6);
  bad += check_type<std::complex<double>>("complex<double>", 8e-15);
  std::printf(bad ? "FAILED\n" : "cpp mirror ok\n");
  return bad;
}
