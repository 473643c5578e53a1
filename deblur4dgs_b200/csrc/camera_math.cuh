// camera_math.cuh -- row a7 of SURVEY.md section 8: the camera sub-exposure pose interpolation
// (flow3d/models/move_model.py:138-147, 168-176; flow3d/models/utils/spline_utils.py:371-408,
// 26-54, 204-215), written once over the scalar type S (float forward, Dual<12> for the
// vector-Jacobian product w.r.t. the two 6-vectors the MoveModel heads emit).
//
//   start6, end6 : se(3) tangents in pypose order [rho(3), phi(3)]
//   pp.se3(x).Exp()            -> (t = J_l(phi) rho, q = exp(phi))            [pypose se3_Exp]
//   linear_interpolation(u)    -> t_u = (1-u) t0 + u t1 ; q_u = q0 (x) Exp(u Log(q0^-1 (x) q1))
//   .Log()                     -> [rho' = J_l^-1(phi') t_u, phi' = Log(q_u)]  [pypose SE3_Log]
//   se3_to_SE3([rho', phi'])   -> the reference decodes this vector in BAD-NeRF order [w, u]:
//                                 R = I + A wx + B wx^2, V = I + B wx + C wx^2 with w = rho',
//                                 u = phi' and 11-term Taylor series A, B, C (spline_utils.py:26-54).
//                                 That order mix is reference behaviour and is reproduced literally.
// pypose (0.6.8) is an absent third-party dependency: its Exp/Log formulas are restated from the
// published implementation -- "parity unpinned" for those; the se3_to_SE3 / Taylor part IS pinned
// against the reference's own spline_utils.py (tests/golden/camera_se3.npz).
#pragma once
#include "deform_math.cuh"

namespace d4 {

D4_HD float d_sin(float a) { return sinf(a); }
D4_HD float d_cos(float a) { return cosf(a); }
D4_HD float d_atan(float a) { return atanf(a); }
template <int ND>
D4_HD Dual<ND> d_sin(const Dual<ND> &a) {
    Dual<ND> r;
    r.v = sinf(a.v);
    float c = cosf(a.v);
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = c * a.d[i];
    return r;
}
template <int ND>
D4_HD Dual<ND> d_cos(const Dual<ND> &a) {
    Dual<ND> r;
    r.v = cosf(a.v);
    float s = -sinf(a.v);
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = s * a.d[i];
    return r;
}
template <int ND>
D4_HD Dual<ND> d_atan(const Dual<ND> &a) {
    Dual<ND> r;
    r.v = atanf(a.v);
    float k = 1.0f / (1.0f + a.v * a.v);
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = k * a.d[i];
    return r;
}

template <typename S>
D4_HD void d_cross(const S *a, const S *b, S *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// Hamilton product, xyzw
template <typename S>
D4_HD void quat_mul_xyzw(const S *p, const S *q, S *o) {
    o[0] = p[3] * q[0] + q[3] * p[0] + (p[1] * q[2] - p[2] * q[1]);
    o[1] = p[3] * q[1] + q[3] * p[1] + (p[2] * q[0] - p[0] * q[2]);
    o[2] = p[3] * q[2] + q[3] * p[2] + (p[0] * q[1] - p[1] * q[0]);
    o[3] = p[3] * q[3] - (p[0] * q[0] + p[1] * q[1] + p[2] * q[2]);
}

constexpr float kLieEps = 1.1920928955078125e-07f;  // torch.finfo(float32).eps

template <typename S>
D4_HD void so3_exp(const S *phi, S *q) {
    S th2 = d_dot3(phi, phi);
    S imag, real;
    S c = d_const(th2, 1.0f);
    if (d_val(th2) > kLieEps * kLieEps) {
        S th = d_sqrt(th2);
        S half = th * d_const(th2, 0.5f);
        imag = d_sin(half) / th;
        real = d_cos(half);
    } else {
        S th4 = th2 * th2;
        imag = d_const(th2, 0.5f) - d_const(th2, 1.0f / 48.0f) * th2 + d_const(th2, 1.0f / 3840.0f) * th4;
        real = c - d_const(th2, 1.0f / 8.0f) * th2 + d_const(th2, 1.0f / 384.0f) * th4;
    }
    q[0] = phi[0] * imag; q[1] = phi[1] * imag; q[2] = phi[2] * imag; q[3] = real;
}

template <typename S>
D4_HD void so3_log(const S *q, S *r) {
    S vn2 = d_dot3(q, q);
    S f;
    const S &w = q[3];
    if (d_val(vn2) > kLieEps * kLieEps) {
        S vn = d_sqrt(vn2);
        if (fabsf(d_val(w)) > kLieEps) {
            f = d_const(vn2, 2.0f) * d_atan(vn / w) / vn;
        } else {
            f = d_const(vn2, d_val(w) >= 0.f ? 3.14159265358979323846f : -3.14159265358979323846f) / vn;
        }
    } else {
        f = d_const(vn2, 2.0f) / w - d_const(vn2, 2.0f / 3.0f) * vn2 / (w * w * w);
    }
    r[0] = f * q[0]; r[1] = f * q[1]; r[2] = f * q[2];
}

// t = J_l(phi) rho
template <typename S>
D4_HD void so3_jl_apply(const S *phi, const S *rho, S *t) {
    S th2 = d_dot3(phi, phi);
    S c1, c2;
    if (d_val(th2) > kLieEps * kLieEps) {
        S th = d_sqrt(th2);
        c1 = (d_const(th2, 1.0f) - d_cos(th)) / th2;
        c2 = (th - d_sin(th)) / (th2 * th);
    } else {
        c1 = d_const(th2, 0.5f) - d_const(th2, 1.0f / 24.0f) * th2;
        c2 = d_const(th2, 1.0f / 6.0f) - d_const(th2, 1.0f / 120.0f) * th2;
    }
    S k1[3], k2[3];
    d_cross(phi, rho, k1);
    d_cross(phi, k1, k2);
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = rho[i] + c1 * k1[i] + c2 * k2[i];
}

// rho = J_l^-1(phi) t
template <typename S>
D4_HD void so3_jl_inv_apply(const S *phi, const S *t, S *rho) {
    S th2 = d_dot3(phi, phi);
    S c2;
    if (d_val(th2) > kLieEps * kLieEps) {
        S th = d_sqrt(th2);
        S half = th * d_const(th2, 0.5f);
        c2 = (d_const(th2, 1.0f) - th * d_cos(half) / (d_const(th2, 2.0f) * d_sin(half))) / th2;
    } else {
        c2 = d_const(th2, 1.0f / 12.0f);
    }
    S k1[3], k2[3];
    d_cross(phi, t, k1);
    d_cross(phi, k1, k2);
#pragma unroll
    for (int i = 0; i < 3; ++i) rho[i] = t[i] - d_const(th2, 0.5f) * k1[i] + c2 * k2[i];
}

// spline_utils.py:26-54 Taylor series in x^2 (x = |w|), 11 terms each
template <typename S>
D4_HD void taylor_abc(const S &x2, S *A, S *B, S *C) {
    S a = d_const(x2, 0.f), b = d_const(x2, 0.f), c = d_const(x2, 0.f);
    S pw = d_const(x2, 1.f);  // x^(2i)
    float da = 1.0f, db = 1.0f, dc = 1.0f;
#pragma unroll
    for (int i = 0; i <= 10; ++i) {
        if (i > 0) da *= (float)((2 * i) * (2 * i + 1));
        db *= (float)((2 * i + 1) * (2 * i + 2));
        dc *= (float)((2 * i + 2) * (2 * i + 3));
        const float sgn = (i & 1) ? -1.0f : 1.0f;
        a = a + d_const(x2, sgn / da) * pw;
        b = b + d_const(x2, sgn / db) * pw;
        c = c + d_const(x2, sgn / dc) * pw;
        pw = pw * x2;
    }
    *A = a; *B = b; *C = c;
}

// se3_to_SE3 (spline_utils.py:204-215): wu = [w(3), u(3)] -> Rt[12] row-major 3x4
template <typename S>
D4_HD void se3_to_SE3_mat(const S *w, const S *u, S *Rt) {
    S x2 = d_dot3(w, w);
    S A, B, C;
    taylor_abc(x2, &A, &B, &C);
    S zero = d_const(x2, 0.f);
    // wx and wx^2
    S wx[9] = {zero, zero - w[2], w[1], w[2], zero, zero - w[0], zero - w[1], w[0], zero};
    S wx2[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) wx2[3 * i + j] = wx[3 * i] * wx[j] + wx[3 * i + 1] * wx[3 + j] + wx[3 * i + 2] * wx[6 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        S vu = zero;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            S idn = d_const(x2, i == j ? 1.f : 0.f);
            Rt[4 * i + j] = idn + A * wx[3 * i + j] + B * wx2[3 * i + j];
            vu = vu + (idn + B * wx[3 * i + j] + C * wx2[3 * i + j]) * u[j];
        }
        Rt[4 * i + 3] = vu;
    }
}

// one interpolated camera delta: start6/end6 in pypose [rho, phi] order, u in [0,1] -> Rt[12]
template <typename S>
D4_HD void camera_interp_one(const S *start6, const S *end6, float u, S *Rt) {
    S t0[3], q0[4], t1[3], q1[4];
    so3_jl_apply(start6 + 3, start6, t0);
    so3_exp(start6 + 3, q0);
    so3_jl_apply(end6 + 3, end6, t1);
    so3_exp(end6 + 3, q1);
    S uu = d_const(t0[0], u), um = d_const(t0[0], 1.0f - u);
    S tu[3] = {um * t0[0] + uu * t1[0], um * t0[1] + uu * t1[1], um * t0[2] + uu * t1[2]};
    S zero = d_const(t0[0], 0.f);
    S q0c[4] = {zero - q0[0], zero - q0[1], zero - q0[2], q0[3]};
    S qrel[4], r[3], qt[4], qu[4];
    quat_mul_xyzw(q0c, q1, qrel);
    so3_log(qrel, r);
    S ur[3] = {uu * r[0], uu * r[1], uu * r[2]};
    so3_exp(ur, qt);
    quat_mul_xyzw(q0, qt, qu);
    S phi[3], rho[3];
    so3_log(qu, phi);
    so3_jl_inv_apply(phi, tu, rho);
    se3_to_SE3_mat(rho, phi, Rt);  // reference quirk: [rho', phi'] decoded as [w, u]
}

}  // namespace d4
