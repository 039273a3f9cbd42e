/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the arithmetic Jalla's hot path reaches through
 * its FFI: cblas_?gemm and LAPACKE_?{gesv,getrf,getrs,getri,potrf,potrs,posv}
 * for s/d/c/z.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libjalla_b200.so) never links or calls it.
 *
 * Where the algorithm lives: NOT in /root/reference.  Jalla links external
 * `lapacke`, `cblas`, `blas` (jalla.cabal:75, unpinned; README asks for LAPACK
 * >= 3.4, vendored headers are LAPACKE 3.4.0, lapacke/include/lapacke.h:31).
 * This file restates the published reference-BLAS/LAPACK algorithms (?gemm,
 * ?getf2, ?getrs, ?trti2+?getri, ?potf2, ?potrs) and the LAPACKE wrapper
 * conventions.  Pinning: the reference holds no golden vectors for this path
 * (tests/Test.hs:77-112 are properties, not vectors).  The restatement is
 * pinned in tests/test_oracle_pin.py against the same-ABI library Jalla would
 * link here — OpenBLAS 0.3.31.dev (scipy-bundled, LP64, `scipy_` prefix) with
 * OpenBLAS 0.3.15 (opencv-bundled, unprefixed symbols) as a second opinion —
 * and against the restated QuickCheck properties; vectors generated from
 * OpenBLAS are committed under tests/golden/ with their generating script.
 *
 * Build: gcc -O2 -std=c99 -fPIC -shared -ffp-contract=off -o oracle/libjalla_oracle.so oracle/jalla_oracle.c -lm
 * (-ffp-contract=off: every fma in the contract is explicit.)
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <stddef.h>

typedef float _Complex cfloat;
typedef double _Complex cdouble;

static inline cfloat mk_c(float r, float i) { cfloat z; ((float *)&z)[0] = r; ((float *)&z)[1] = i; return z; }
static inline cdouble mk_z(double r, double i) { cdouble z; ((double *)&z)[0] = r; ((double *)&z)[1] = i; return z; }
static inline cfloat cmul_c(cfloat a, cfloat b) {
  float ar = crealf(a), ai = cimagf(a), br = crealf(b), bi = cimagf(b);
  return mk_c(fmaf(-ai, bi, ar * br), fmaf(ai, br, ar * bi));
}
static inline cdouble cmul_z(cdouble a, cdouble b) {
  double ar = creal(a), ai = cimag(a), br = creal(b), bi = cimag(b);
  return mk_z(fma(-ai, bi, ar * br), fma(ai, br, ar * bi));
}
static inline cfloat recip_c(cfloat a) { float r = crealf(a), i = cimagf(a), d = fmaf(i, i, r * r); return mk_c(r / d, -i / d); }
static inline cdouble recip_z(cdouble a) { double r = creal(a), i = cimag(a), d = fma(i, i, r * r); return mk_z(r / d, -i / d); }

/* ---------------- float ---------------- */
#define T float
#define R float
#define ACC double
#define PFX(x) oracle_s##x
#define IS_CPLX 0
#define CONJ(x) (x)
#define ABS1(x) fabsf(x)
#define REALPART(x) (x)
#define IMAGPART(x) (0.0f)
#define MAKE_T(r, i) ((float)(r))
#define ISNAN_T(x) isnan(x)
#define SQRT_R(x) sqrtf(x)
#define FMA_R(a, b, c) fmaf(a, b, c)
#define CMUL(a, b) ((a) * (b))
#define RECIP(a) (1.0f / (a))
#define SCALAR_PARAM float
#define SCALAR_GET(x) (x)
#include "oracle_impl.inc"
#undef T
#undef R
#undef ACC
#undef PFX
#undef IS_CPLX
#undef CONJ
#undef ABS1
#undef REALPART
#undef IMAGPART
#undef MAKE_T
#undef ISNAN_T
#undef SQRT_R
#undef FMA_R
#undef CMUL
#undef RECIP
#undef SCALAR_PARAM
#undef SCALAR_GET

/* ---------------- double ---------------- */
#define T double
#define R double
#define ACC long double
#define PFX(x) oracle_d##x
#define IS_CPLX 0
#define CONJ(x) (x)
#define ABS1(x) fabs(x)
#define REALPART(x) (x)
#define IMAGPART(x) (0.0)
#define MAKE_T(r, i) ((double)(r))
#define ISNAN_T(x) isnan(x)
#define SQRT_R(x) sqrt(x)
#define FMA_R(a, b, c) fma(a, b, c)
#define CMUL(a, b) ((a) * (b))
#define RECIP(a) (1.0 / (a))
#define SCALAR_PARAM double
#define SCALAR_GET(x) (x)
#include "oracle_impl.inc"
#undef T
#undef R
#undef ACC
#undef PFX
#undef IS_CPLX
#undef CONJ
#undef ABS1
#undef REALPART
#undef IMAGPART
#undef MAKE_T
#undef ISNAN_T
#undef SQRT_R
#undef FMA_R
#undef CMUL
#undef RECIP
#undef SCALAR_PARAM
#undef SCALAR_GET

/* ---------------- complex float ---------------- */
#define T cfloat
#define R float
#define ACC cdouble
#define PFX(x) oracle_c##x
#define IS_CPLX 1
#define CONJ(x) conjf(x)
#define ABS1(x) (fabsf(crealf(x)) + fabsf(cimagf(x)))
#define REALPART(x) crealf(x)
#define IMAGPART(x) cimagf(x)
#define MAKE_T(r, i) mk_c((r), (i))
#define ISNAN_T(x) (isnan(crealf(x)) || isnan(cimagf(x)))
#define SQRT_R(x) sqrtf(x)
#define FMA_R(a, b, c) fmaf(a, b, c)
#define CMUL(a, b) cmul_c((a), (b))
#define RECIP(a) recip_c(a)
#define SCALAR_PARAM const void *
#define SCALAR_GET(x) (*(const cfloat *)(x))
#include "oracle_impl.inc"
#undef T
#undef R
#undef ACC
#undef PFX
#undef IS_CPLX
#undef CONJ
#undef ABS1
#undef REALPART
#undef IMAGPART
#undef MAKE_T
#undef ISNAN_T
#undef SQRT_R
#undef FMA_R
#undef CMUL
#undef RECIP
#undef SCALAR_PARAM
#undef SCALAR_GET

/* ---------------- complex double ---------------- */
typedef long double _Complex cldouble;
#define T cdouble
#define R double
#define ACC cldouble
#define PFX(x) oracle_z##x
#define IS_CPLX 1
#define CONJ(x) conj(x)
#define ABS1(x) (fabs(creal(x)) + fabs(cimag(x)))
#define REALPART(x) creal(x)
#define IMAGPART(x) cimag(x)
#define MAKE_T(r, i) mk_z((r), (i))
#define ISNAN_T(x) (isnan(creal(x)) || isnan(cimag(x)))
#define SQRT_R(x) sqrt(x)
#define FMA_R(a, b, c) fma(a, b, c)
#define CMUL(a, b) cmul_z((a), (b))
#define RECIP(a) recip_z(a)
#define SCALAR_PARAM const void *
#define SCALAR_GET(x) (*(const cdouble *)(x))
#include "oracle_impl.inc"
