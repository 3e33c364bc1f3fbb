// ref_abcd.cuh -- the reference's 2x2 ray-transfer algebra on the device, with the
// reference's rounding sequence (src/pathtracer/pathtracer.cpp:511-537 build each
// matrix entry in FLOAT; CGL/src/matrix3x3.cpp:99-114 multiplies in DOUBLE, two
// rounded products and one rounded sum per entry).  Every operation uses an explicit
// round-to-nearest intrinsic so nvcc can never contract it into an FMA: the x86-64
// reference build has no FMA, and parity with it is bit-for-bit.
#pragma once
#include "lfb_internal.h"

namespace lfb {

struct m2 { double a, b, c, d; };  // [[a,b],[c,d]]

__device__ __forceinline__ m2 make2(float a, float b, float c, float d) {
  m2 m; m.a = a; m.b = b; m.c = c; m.d = d; return m;
}
__device__ __forceinline__ m2 mT(float d) { return make2(1.f, d, 0.f, 1.f); }            // :527-529
__device__ __forceinline__ m2 mR(float c, float n1, float n2) {                             // :531-533
  return make2(1.f, 0.f, __fdiv_rn(__fmul_rn(c, __fsub_rn(n1, n2)), n2), __fdiv_rn(n1, n2));
}
__device__ __forceinline__ m2 mL(float c) { return make2(1.f, 0.f, __fmul_rn(2.f, c), 1.f); }  // :535-537

// C = A*B, entry (r,c) = B(0,c)*A(r,0) + B(1,c)*A(r,1)
__device__ __forceinline__ m2 mmul(const m2& A, const m2& B) {
  m2 C;
  C.a = __dadd_rn(__dmul_rn(B.a, A.a), __dmul_rn(B.c, A.b));
  C.c = __dadd_rn(__dmul_rn(B.a, A.c), __dmul_rn(B.c, A.d));
  C.b = __dadd_rn(__dmul_rn(B.b, A.a), __dmul_rn(B.d, A.b));
  C.d = __dadd_rn(__dmul_rn(B.b, A.c), __dmul_rn(B.d, A.d));
  return C;
}
// invert2x2 :519-525: float temporaries and determinant, double scale
__device__ __forceinline__ m2 minv(const m2& m) {
  float a = (float)m.a, b = (float)m.b, c = (float)m.c, d = (float)m.d;
  float det = __fsub_rn(__fmul_rn(a, d), __fmul_rn(b, c));
  double s = __ddiv_rn(1.0, (double)det);
  m2 r;
  r.a = __dmul_rn((double)d, s); r.b = __dmul_rn((double)(-b), s);
  r.c = __dmul_rn((double)(-c), s); r.d = __dmul_rn((double)a, s);
  return r;
}
__device__ __forceinline__ double mrow0(const m2& M, double r, double th) {
  return __dadd_rn(__dmul_rn(r, M.a), __dmul_rn(th, M.b));
}
__device__ __forceinline__ double mrow1(const m2& M, double r, double th) {
  return __dadd_rn(__dmul_rn(r, M.c), __dmul_rn(th, M.d));
}

__device__ __forceinline__ float n_before(const DevLens& L, int lambda, int k) {
  return k == 0 ? 1.00f : L.ior[lambda][k - 1];
}
__device__ __forceinline__ m2 surf_R(const DevLens& L, int lambda, int k) {  // create_Rs_for_color :559-569
  return mR(L.c[k], n_before(L, lambda, k), L.ior[lambda][k]);
}
__device__ __forceinline__ m2 step_TR(const DevLens& L, int lambda, int k, const m2& M) {
  return mmul(mmul(mT(L.d[k]), surf_R(L, lambda, k)), M);
}
__device__ __forceinline__ m2 step_back(const DevLens& L, int lambda, int k, const m2& M, int physical) {
  m2 back = physical ? mR(-L.c[k], L.ior[lambda][k], n_before(L, lambda, k)) : minv(surf_R(L, lambda, k));
  return mmul(mmul(back, mT(L.d[k])), M);
}
__device__ __forceinline__ m2 step_second_reflection(const DevLens& L, int i, const m2& M) {
  return mmul(mmul(mmul(mT(L.d[i]), minv(mL(L.c[i]))), mT(L.d[i])), M);
}

}  // namespace lfb
