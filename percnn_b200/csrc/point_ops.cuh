// Per-cell arithmetic of every cell variant, forward and adjoint, templated on the scalar type.
// Kernels gather the cross neighbourhood (radius 2 along each axis) and call into these.
#pragma once
#include "common.cuh"

namespace percnn {

// Centre value + cross neighbours of one field.  n[a][k]: axis a (0 = slowest), k = offsets -2,-1,+1,+2.
template <typename T, int NDIM>
struct Cross {
  T c;
  T n[NDIM][4];
};

// Sum_k taps * neighbours, with the prep-block tap layout (P_LAP_C0, P_LAP_AX).  W_laplace taps are the
// dense table of GS2D:20-24 / GS3D:22-39 restricted to its cross (all other entries are zero).
template <typename T, int NDIM>
__device__ __forceinline__ T lap_apply(const Cross<T, NDIM>& q, const T* __restrict__ P) {
  T acc = P[P_LAP_C0] * q.c;
#pragma unroll
  for (int a = 0; a < NDIM; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc = fma_t(P[P_LAP_AX + a * 4 + k], q.n[a][k], acc);
  return acc;
}

// Transposed stencil (taps mirrored): (Lap^T g)(x) = sum_o w[o] g(x - o).
template <typename T, int NDIM>
__device__ __forceinline__ T lap_apply_T(const Cross<T, NDIM>& q, const T* __restrict__ P) {
  T acc = P[P_LAP_C0] * q.c;
#pragma unroll
  for (int a = 0; a < NDIM; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc = fma_t(P[P_LAP_AX + a * 4 + (3 - k)], q.n[a][k], acc);
  return acc;
}

// Bivariate cubic in Horner form: 9 FMAs.  c = c00 c10 c01 c20 c11 c02 c30 c21 c12 c03.
template <typename T>
__device__ __forceinline__ T cubic_eval(const T* __restrict__ c, T u, T v) {
  T a0 = fma_t(u, fma_t(u, fma_t(u, c[6], c[3]), c[1]), c[0]);
  T a1 = fma_t(u, fma_t(u, c[7], c[4]), c[2]);
  T a2 = fma_t(u, c[8], c[5]);
  return fma_t(v, fma_t(v, fma_t(v, c[9], a2), a1), a0);
}

// Quadratic over monomials 1 u v u^2 uv v^2: 5 FMAs.
template <typename T>
__device__ __forceinline__ T quad_eval(const T* __restrict__ d, T u, T v) {
  T a0 = fma_t(u, fma_t(u, d[3], d[1]), d[0]);
  T a1 = fma_t(u, d[4], d[2]);
  return fma_t(v, fma_t(v, d[5], a1), a0);
}

// ---- Pi-block cell, k = 1 -------------------------------------------------------------------
// GS2D:115-118: res = (mu_up*sigmoid(CA)) * Lap(u) + Wh4(Wh1*Wh2*Wh3)(h);  u+ = u + res*dt.
template <typename T>
__device__ __forceinline__ void pi_k1_fwd_poly(T u, T v, T Lu, T Lv, const T* __restrict__ P, T& ou, T& ov) {
  T ru = fma_t(P[P_ALPHA + 0], Lu, cubic_eval(P + P_POLY, u, v));
  T rv = fma_t(P[P_ALPHA + 1], Lv, cubic_eval(P + P_POLY + 10, u, v));
  ou = fma_t(ru, P[P_DT], u);
  ov = fma_t(rv, P[P_DT], v);
}

// Channel-by-channel evaluation in the reference's own operation order ((P1*P2)*P3, then the 1x1 sum).
template <typename T>
__device__ __forceinline__ T pi_k1_branch_R(const T* __restrict__ B, int hc, T u, T v) {
  const T* W1 = B;
  const T* b1 = W1 + 2 * hc;
  const T* W2 = b1 + hc;
  const T* b2 = W2 + 2 * hc;
  const T* W3 = b2 + hc;
  const T* b3 = W3 + 2 * hc;
  const T* W4 = b3 + hc;
  T acc = W4[hc];  // b4
  for (int c = 0; c < hc; ++c) {
    T p1 = fma_t(W1[2 * c + 1], v, fma_t(W1[2 * c], u, b1[c]));
    T p2 = fma_t(W2[2 * c + 1], v, fma_t(W2[2 * c], u, b2[c]));
    T p3 = fma_t(W3[2 * c + 1], v, fma_t(W3[2 * c], u, b3[c]));
    acc = fma_t(W4[c], (p1 * p2) * p3, acc);
  }
  return acc;
}

template <typename T>
__device__ __forceinline__ void pi_k1_fwd_branch(T u, T v, T Lu, T Lv, const T* __restrict__ P, int hc, T& ou, T& ov) {
  const int per_field = 10 * hc + 1;
  T ru = fma_t(P[P_ALPHA + 0], Lu, pi_k1_branch_R(P + P_BRANCH, hc, u, v));
  T rv = fma_t(P[P_ALPHA + 1], Lv, pi_k1_branch_R(P + P_BRANCH + per_field, hc, u, v));
  ou = fma_t(ru, P[P_DT], u);
  ov = fma_t(rv, P[P_DT], v);
}

// Adjoint of the k=1 step, with the Pi-block folded to its cubic:
//   g_u = Gu + dt*(alpha_u * Lap^T Gu + Gu dRu/du + Gv dRv/du)      (same for v)
//   red[0..1]  += q * dt*Lap^T(Gq)            -> dL/dalpha_q   (sum_x Gq Lap(q) = sum_x Lap^T(Gq) q)
//   red[2+10f+m] += dt*G_f * monomial_m(u,v)  -> dL/dc^f_m
template <typename T>
__device__ __forceinline__ void pi_k1_bwd_poly(T u, T v, T Gu, T Gv, T LTu, T LTv, const T* __restrict__ P,
                                               T& gu, T& gv, T* __restrict__ red) {
  const T dt = P[P_DT];
  const T Gdu = dt * Gu, Gdv = dt * Gv;
  const T* D = P + P_DPOLY;
  T su = fma_t(Gdu, quad_eval(D + 0, u, v), Gdv * quad_eval(D + 12, u, v));
  T sv = fma_t(Gdu, quad_eval(D + 6, u, v), Gdv * quad_eval(D + 18, u, v));
  T lu = dt * LTu, lv = dt * LTv;
  gu = Gu + fma_t(P[P_ALPHA + 0], lu, su);
  gv = Gv + fma_t(P[P_ALPHA + 1], lv, sv);
  red[0] = fma_t(u, lu, red[0]);
  red[1] = fma_t(v, lv, red[1]);
  const T uu = u * u, uv = u * v, vv = v * v;
  const T m[10] = {T(1), u, v, uu, uv, vv, uu * u, uu * v, u * vv, vv * v};
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    red[2 + i] = fma_t(Gdu, m[i], red[2 + i]);
    red[12 + i] = fma_t(Gdv, m[i], red[12 + i]);
  }
}

// ---- first-derivative helper for the Stage-3 cells ------------------------------------------
// taps t[0..3] at offsets -2,-1,+1,+2 (centre tap of dx_2d_op is 0, BUR3:20-30)
template <typename T>
__device__ __forceinline__ T der_apply(const T* __restrict__ t, const T* __restrict__ n) {
  return fma_t(t[3], n[3], fma_t(t[2], n[2], fma_t(t[1], n[1], t[0] * n[0])));
}

template <typename T>
__device__ __forceinline__ T der_apply_T(const T* __restrict__ t, const T* __restrict__ n) {
  return fma_t(t[0], n[3], fma_t(t[1], n[2], fma_t(t[2], n[1], t[3] * n[0])));
}

// ---- Burgers physics cell (BUR3:154-157, 209-221) -------------------------------------------
// f_u = nu_u Lap(u) + C1_u u Dx(u) + C2_u v Dy(u);  f_v = nu_v Lap(v) + C1_v u Dx(v) + C2_v v Dy(v)
// Dx acts along tensor dim 2 (axis 0 here), Dy along dim 3 (axis 1).
template <typename T>
__device__ __forceinline__ void burgers_fwd(const Cross<T, 2>& U, const Cross<T, 2>& V, const T* __restrict__ P,
                                            T& ou, T& ov) {
  const T* C = P + P_PHYS;
  const T* tx = C + 4;
  const T* ty = C + 8;
  T fu = fma_t(C[1] * V.c, der_apply(ty, U.n[1]), fma_t(C[0] * U.c, der_apply(tx, U.n[0]), P[P_ALPHA + 0] * lap_apply<T, 2>(U, P)));
  T fv = fma_t(C[3] * V.c, der_apply(ty, V.n[1]), fma_t(C[2] * U.c, der_apply(tx, V.n[0]), P[P_ALPHA + 1] * lap_apply<T, 2>(V, P)));
  ou = fma_t(P[P_DT], fu, U.c);
  ov = fma_t(P[P_DT], fv, V.c);
}

// The right-hand side alone (f_rhs, BUR3:154-157): the RK4 cell evaluates it four times per step.
template <typename T>
__device__ __forceinline__ void burgers_rhs(const Cross<T, 2>& U, const Cross<T, 2>& V, const T* __restrict__ P, T& fu, T& fv) {
  const T* C = P + P_PHYS;
  const T* tx = C + 4;
  const T* ty = C + 8;
  fu = fma_t(C[1] * V.c, der_apply(ty, U.n[1]), fma_t(C[0] * U.c, der_apply(tx, U.n[0]), P[P_ALPHA + 0] * lap_apply<T, 2>(U, P)));
  fv = fma_t(C[3] * V.c, der_apply(ty, V.n[1]), fma_t(C[2] * U.c, der_apply(tx, V.n[0]), P[P_ALPHA + 1] * lap_apply<T, 2>(V, P)));
}

// Adjoint ("Adjoint for a4").  With D^T the mirrored-tap stencil (= -D for the antisymmetric
// reference taps):
//   g_u = Gu + dt[nu_u L^T Gu + C1_u Dx(u) Gu + Dx^T(C1_u u Gu) + Dy^T(C2_u v Gu) + C1_v Dx(v) Gv]
//   g_v = Gv + dt[nu_v L^T Gv + Dx^T(C1_v u Gv) + C2_v Dy(v) Gv + Dy^T(C2_v v Gv) + C2_u Dy(u) Gu]
//   dL/dnu_q = dt sum Gq L(q) = dt sum L^T(Gq) q ; dL/dC1_q = dt sum Gq u Dx(q) ; dL/dC2_q = dt sum Gq v Dy(q)
template <typename T>
__device__ __forceinline__ void burgers_bwd(const Cross<T, 2>& U, const Cross<T, 2>& V, const Cross<T, 2>& GU,
                                            const Cross<T, 2>& GV, const T* __restrict__ P, T& gu, T& gv,
                                            T* __restrict__ red) {
  const T* C = P + P_PHYS;
  const T* tx = C + 4;
  const T* ty = C + 8;
  const T dt = P[P_DT];
  const T dxu = der_apply(tx, U.n[0]), dyu = der_apply(ty, U.n[1]);
  const T dxv = der_apply(tx, V.n[0]), dyv = der_apply(ty, V.n[1]);
  // products at the neighbour positions, for the transposed (negated) first-derivative stencils
  T uGu_x[4], vGu_y[4], uGv_x[4], vGv_y[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uGu_x[k] = U.n[0][k] * GU.n[0][k];
    uGv_x[k] = U.n[0][k] * GV.n[0][k];
    vGu_y[k] = V.n[1][k] * GU.n[1][k];
    vGv_y[k] = V.n[1][k] * GV.n[1][k];
  }
  const T LTu = lap_apply_T<T, 2>(GU, P), LTv = lap_apply_T<T, 2>(GV, P);
  T au = P[P_ALPHA + 0] * LTu;
  au = fma_t(C[0] * dxu, GU.c, au);
  au = fma_t(C[0], der_apply_T(tx, uGu_x), au);
  au = fma_t(C[1], der_apply_T(ty, vGu_y), au);
  au = fma_t(C[2] * dxv, GV.c, au);
  T av = P[P_ALPHA + 1] * LTv;
  av = fma_t(C[2], der_apply_T(tx, uGv_x), av);
  av = fma_t(C[3] * dyv, GV.c, av);
  av = fma_t(C[3], der_apply_T(ty, vGv_y), av);
  av = fma_t(C[1] * dyu, GU.c, av);
  gu = fma_t(dt, au, GU.c);
  gv = fma_t(dt, av, GV.c);
  const T Gdu = dt * GU.c, Gdv = dt * GV.c;
  red[0] = fma_t(dt * LTu, U.c, red[0]);             // nu_u
  red[1] = fma_t(dt * LTv, V.c, red[1]);             // nu_v
  red[2] = fma_t(Gdu * U.c, dxu, red[2]);            // C1_u
  red[3] = fma_t(Gdu * V.c, dyu, red[3]);            // C2_u
  red[4] = fma_t(Gdv * U.c, dxv, red[4]);            // C1_v
  red[5] = fma_t(Gdv * V.c, dyv, red[5]);            // C2_v
}

// ---- lambda-omega physics cell (LO3:148-151, 203-215) -----------------------------------------
// f_u = nu_u Lap(u) + C1_u u + C2_u u^3 + C3_u u^2 v + C4_u u v^2 + C5_u v^3
// f_v = nu_v Lap(v) + C1_v v + C2_v u^3 + C3_v u^2 v + C4_v u v^2 + C5_v v^3 (+ C6_v u)
template <typename T>
__device__ __forceinline__ void lo_fwd(T u, T v, T Lu, T Lv, const T* __restrict__ P, T& ou, T& ov) {
  const T* Cu = P + P_PHYS;
  const T* Cv = Cu + 5;
  const T uu = u * u, vv = v * v;
  const T u3 = uu * u, u2v = uu * v, uv2 = u * vv, v3 = vv * v;
  T fu = P[P_ALPHA + 0] * Lu;
  fu = fma_t(Cu[0], u, fu);
  fu = fma_t(Cu[1], u3, fu);
  fu = fma_t(Cu[2], u2v, fu);
  fu = fma_t(Cu[3], uv2, fu);
  fu = fma_t(Cu[4], v3, fu);
  T fv = P[P_ALPHA + 1] * Lv;
  fv = fma_t(Cv[0], v, fv);
  fv = fma_t(Cv[1], u3, fv);
  fv = fma_t(Cv[2], u2v, fv);
  fv = fma_t(Cv[3], uv2, fv);
  fv = fma_t(Cv[4], v3, fv);
  fv = fma_t(Cv[5], u, fv);  // C6_v (0 when the cell has no such term)
  ou = fma_t(P[P_DT], fu, u);
  ov = fma_t(P[P_DT], fv, v);
}

// f_rhs of the lambda-omega cell (LO3:148-151).
template <typename T>
__device__ __forceinline__ void lo_rhs(T u, T v, T Lu, T Lv, const T* __restrict__ P, T& fu, T& fv) {
  const T* Cu = P + P_PHYS;
  const T* Cv = Cu + 5;
  const T uu = u * u, vv = v * v;
  const T u3 = uu * u, u2v = uu * v, uv2 = u * vv, v3 = vv * v;
  fu = P[P_ALPHA + 0] * Lu;
  fu = fma_t(Cu[0], u, fu);
  fu = fma_t(Cu[1], u3, fu);
  fu = fma_t(Cu[2], u2v, fu);
  fu = fma_t(Cu[3], uv2, fu);
  fu = fma_t(Cu[4], v3, fu);
  fv = P[P_ALPHA + 1] * Lv;
  fv = fma_t(Cv[0], v, fv);
  fv = fma_t(Cv[1], u3, fv);
  fv = fma_t(Cv[2], u2v, fv);
  fv = fma_t(Cv[3], uv2, fv);
  fv = fma_t(Cv[4], v3, fv);
  fv = fma_t(Cv[5], u, fv);  // C6_v (0 when the cell has no such term)
}

template <typename T>
__device__ __forceinline__ void lo_bwd(T u, T v, T Gu, T Gv, T LTu, T LTv, const T* __restrict__ P, T& gu, T& gv,
                                       T* __restrict__ red) {
  const T* Cu = P + P_PHYS;
  const T* Cv = Cu + 5;
  const T dt = P[P_DT];
  const T uu = u * u, uv = u * v, vv = v * v;
  const T dfu_du = Cu[0] + T(3) * Cu[1] * uu + T(2) * Cu[2] * uv + Cu[3] * vv;
  const T dfu_dv = Cu[2] * uu + T(2) * Cu[3] * uv + T(3) * Cu[4] * vv;
  const T dfv_du = T(3) * Cv[1] * uu + T(2) * Cv[2] * uv + Cv[3] * vv + Cv[5];
  const T dfv_dv = Cv[0] + Cv[2] * uu + T(2) * Cv[3] * uv + T(3) * Cv[4] * vv;
  gu = fma_t(dt, fma_t(P[P_ALPHA + 0], LTu, fma_t(dfu_du, Gu, dfv_du * Gv)), Gu);
  gv = fma_t(dt, fma_t(P[P_ALPHA + 1], LTv, fma_t(dfu_dv, Gu, dfv_dv * Gv)), Gv);
  const T Gdu = dt * Gu, Gdv = dt * Gv;
  const T u3 = uu * u, u2v = uu * v, uv2 = u * vv, v3 = vv * v;
  red[0] = fma_t(dt * LTu, u, red[0]);
  red[1] = fma_t(dt * LTv, v, red[1]);
  red[2] = fma_t(Gdu, u, red[2]);
  red[3] = fma_t(Gdu, u3, red[3]);
  red[4] = fma_t(Gdu, u2v, red[4]);
  red[5] = fma_t(Gdu, uv2, red[5]);
  red[6] = fma_t(Gdu, v3, red[6]);
  red[7] = fma_t(Gdv, v, red[7]);
  red[8] = fma_t(Gdv, u3, red[8]);
  red[9] = fma_t(Gdv, u2v, red[9]);
  red[10] = fma_t(Gdv, uv2, red[10]);
  red[11] = fma_t(Gdv, v3, red[11]);
  red[12] = fma_t(Gdv, u, red[12]);
}

}  // namespace percnn
