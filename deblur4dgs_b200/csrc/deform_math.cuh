// deform_math.cuh -- per-Gaussian SE(3) deformation arithmetic (SURVEY.md rows
// a3-a5), written once as a template over the scalar type:
//   S = float     -> forward kernels
//   S = Dual<16>  -> backward kernels: forward-mode dual numbers carrying the 16
//                    partials (blended translation 3, blended rot6d 6, canonical
//                    mean 3, raw quaternion 4).  Seeding L = <v_out, out> gives
//                    the exact vector-Jacobian product of the SAME code the
//                    forward runs, including the branchy rotation-matrix ->
//                    quaternion conversion, with no hand-derived adjoint to keep
//                    in sync.
// Reference arithmetic: F.normalize (eps 1e-12) flow3d/params.py:39;
// cont_6d_to_rmat flow3d/transforms.py:41-53; R*mu+t and
// quat(R) (x) q flow3d/scene_model.py:89-102 (roma rotmat_to_unitquat /
// quat_product, XYZW).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define D4_HD __host__ __device__ __forceinline__
#else
#define D4_HD static inline
#endif

namespace d4 {

template <int ND>
struct Dual {
    float v;
    float d[ND];
};

// ---- scalar overloads (float) ---------------------------------------------------
D4_HD float d_const(float, float c) { return c; }
D4_HD float d_val(float a) { return a; }
D4_HD float d_sqrt(float a) { return sqrtf(a); }
D4_HD float d_max_const(float a, float c) { return a >= c ? a : c; }

// ---- dual overloads ---------------------------------------------------------------
template <int ND>
D4_HD Dual<ND> d_const(const Dual<ND> &, float c) {
    Dual<ND> r;
    r.v = c;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = 0.f;
    return r;
}
template <int ND>
D4_HD float d_val(const Dual<ND> &a) { return a.v; }
template <int ND>
D4_HD Dual<ND> operator+(const Dual<ND> &a, const Dual<ND> &b) {
    Dual<ND> r;
    r.v = a.v + b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
template <int ND>
D4_HD Dual<ND> operator-(const Dual<ND> &a, const Dual<ND> &b) {
    Dual<ND> r;
    r.v = a.v - b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
template <int ND>
D4_HD Dual<ND> operator*(const Dual<ND> &a, const Dual<ND> &b) {
    Dual<ND> r;
    r.v = a.v * b.v;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
template <int ND>
D4_HD Dual<ND> operator/(const Dual<ND> &a, const Dual<ND> &b) {
    Dual<ND> r;
    float inv = 1.0f / b.v;
    r.v = a.v * inv;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
template <int ND>
D4_HD Dual<ND> d_sqrt(const Dual<ND> &a) {
    Dual<ND> r;
    r.v = sqrtf(a.v);
    float k = a.v > 0.f ? 0.5f / r.v : 0.f;
#pragma unroll
    for (int i = 0; i < ND; ++i) r.d[i] = k * a.d[i];
    return r;
}
template <int ND>
D4_HD Dual<ND> d_max_const(const Dual<ND> &a, float c) { return a.v >= c ? a : d_const(a, c); }

// float arithmetic written through the same spelling as the dual code
template <typename S>
D4_HD S d_dot3(const S *a, const S *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// cont_6d_to_rmat (transforms.py:41-53): Gram-Schmidt of the 6-D rotation, columns x, y, z of R
template <typename S>
D4_HD void rot6d_to_cols(const S *r6, S *x, S *y, S *z) {
    const float eps = 1e-12f;
    S an = d_max_const(d_sqrt(d_dot3(r6, r6)), eps);
    x[0] = r6[0] / an; x[1] = r6[1] / an; x[2] = r6[2] / an;
    S bx = r6[3] * x[0] + r6[4] * x[1] + r6[5] * x[2];
    S yp[3] = {r6[3] - bx * x[0], r6[4] - bx * x[1], r6[5] - bx * x[2]};
    S yn = d_max_const(d_sqrt(d_dot3(yp, yp)), eps);
    y[0] = yp[0] / yn; y[1] = yp[1] / yn; y[2] = yp[2] / yn;
    z[0] = x[1] * y[2] - x[2] * y[1];
    z[1] = x[2] * y[0] - x[0] * y[2];
    z[2] = x[0] * y[1] - x[1] * y[0];
}

// (tl, r6, mu, q_raw[wxyz]) -> (mu' = R mu + tl, q' = normalize(wxyz(quat(R) (x) xyzw(normalize(q_raw)))))
template <typename S>
D4_HD void deform_point(const S *tl, const S *r6, const S *mu, const S *qraw, S *om, S *oq) {
    const float eps = 1e-12f;
    // q_hat = F.normalize(q_raw)
    S qn = d_sqrt(qraw[0] * qraw[0] + qraw[1] * qraw[1] + qraw[2] * qraw[2] + qraw[3] * qraw[3]);
    S qd = d_max_const(qn, eps);
    S qh[4] = {qraw[0] / qd, qraw[1] / qd, qraw[2] / qd, qraw[3] / qd};
    // Gram-Schmidt (cont_6d_to_rmat): columns x, y, z
    S x[3], y[3], z[3];
    rot6d_to_cols(r6, x, y, z);
    // R[i][0] = x[i], R[i][1] = y[i], R[i][2] = z[i]
#pragma unroll
    for (int i = 0; i < 3; ++i) om[i] = x[i] * mu[0] + y[i] * mu[1] + z[i] * mu[2] + tl[i];
    // roma.rotmat_to_unitquat (XYZW), SciPy-style: argmax over (R00, R11, R22, trace)
    S tr = x[0] + y[1] + z[2];
    float d0 = d_val(x[0]), d1 = d_val(y[1]), d2 = d_val(z[2]), d3 = d_val(tr);
    int choice = 0;
    float best = d0;
    if (d1 > best) { best = d1; choice = 1; }
    if (d2 > best) { best = d2; choice = 2; }
    if (d3 > best) { best = d3; choice = 3; }
    S one = d_const(tr, 1.0f), two = d_const(tr, 2.0f);
    S p[4];  // xyzw
    if (choice == 3) {
        p[0] = y[2] - z[1];  // R21 - R12
        p[1] = z[0] - x[2];  // R02 - R20
        p[2] = x[1] - y[0];  // R10 - R01
        p[3] = one + tr;
    } else if (choice == 0) {  // i=0, j=1, k=2
        p[0] = one - tr + two * x[0];
        p[1] = x[1] + y[0];  // R10 + R01
        p[2] = x[2] + z[0];  // R20 + R02
        p[3] = y[2] - z[1];  // R21 - R12
    } else if (choice == 1) {  // i=1, j=2, k=0
        p[1] = one - tr + two * y[1];
        p[2] = y[2] + z[1];  // R21 + R12
        p[0] = y[0] + x[1];  // R01 + R10
        p[3] = z[0] - x[2];  // R02 - R20
    } else {  // i=2, j=0, k=1
        p[2] = one - tr + two * z[2];
        p[0] = z[0] + x[2];  // R02 + R20
        p[1] = z[1] + y[2];  // R12 + R21
        p[3] = x[1] - y[0];  // R10 - R01
    }
    S pn = d_sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
    p[0] = p[0] / pn; p[1] = p[1] / pn; p[2] = p[2] / pn; p[3] = p[3] / pn;
    // roma.quat_product(p, q) with q = xyzw(q_hat) = (qh1, qh2, qh3, qh0)
    const S &qx = qh[1], &qy = qh[2], &qz = qh[3], &qw = qh[0];
    S vx = p[3] * qx + qw * p[0] + (p[1] * qz - p[2] * qy);
    S vy = p[3] * qy + qw * p[1] + (p[2] * qx - p[0] * qz);
    S vz = p[3] * qz + qw * p[2] + (p[0] * qy - p[1] * qx);
    S w = p[3] * qw - (p[0] * qx + p[1] * qy + p[2] * qz);
    // xyzw -> wxyz, F.normalize
    S on = d_max_const(d_sqrt(w * w + vx * vx + vy * vy + vz * vz), eps);
    oq[0] = w / on; oq[1] = vx / on; oq[2] = vy / on; oq[3] = vz / on;
}

// Reverse-mode vector-Jacobian product of deform_point, derived by hand (cotangents v_om[3], v_oq[4] of the
// outputs -> grad[16] = cotangents of (tl 3, r6 6, mu 3, q_raw 4)): one forward recomputation plus ~250 flops,
// against ~2 000 for the 16-partial forward-mode evaluation above.  tests/test_host_math.py checks it against the
// Dual<16> evaluation of the same code (all four rotation-matrix -> quaternion branches) on the CPU.
D4_HD void deform_point_vjp(const float *tl, const float *r6, const float *mu, const float *qraw, const float *v_om,
                            const float *v_oq, float *om_out, float *grad) {
    const float eps = 1e-12f;
    (void)tl;
    // ---- forward recomputation (same order of operations as deform_point<float>)
    const float qn = sqrtf(qraw[0] * qraw[0] + qraw[1] * qraw[1] + qraw[2] * qraw[2] + qraw[3] * qraw[3]);
    const float qd = qn >= eps ? qn : eps;
    const float qh[4] = {qraw[0] / qd, qraw[1] / qd, qraw[2] / qd, qraw[3] / qd};
    const float *a = r6, *b = r6 + 3;
    const float an_raw = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    const float an = an_raw >= eps ? an_raw : eps;
    const float x[3] = {a[0] / an, a[1] / an, a[2] / an};
    const float bx = b[0] * x[0] + b[1] * x[1] + b[2] * x[2];
    const float yp[3] = {b[0] - bx * x[0], b[1] - bx * x[1], b[2] - bx * x[2]};
    const float yn_raw = sqrtf(yp[0] * yp[0] + yp[1] * yp[1] + yp[2] * yp[2]);
    const float yn = yn_raw >= eps ? yn_raw : eps;
    const float y[3] = {yp[0] / yn, yp[1] / yn, yp[2] / yn};
    const float z[3] = {x[1] * y[2] - x[2] * y[1], x[2] * y[0] - x[0] * y[2], x[0] * y[1] - x[1] * y[0]};
    if (om_out) {
#pragma unroll
        for (int i = 0; i < 3; ++i) om_out[i] = x[i] * mu[0] + y[i] * mu[1] + z[i] * mu[2] + tl[i];
    }
    const float tr = x[0] + y[1] + z[2];
    int choice = 0;
    float best = x[0];
    if (y[1] > best) { best = y[1]; choice = 1; }
    if (z[2] > best) { best = z[2]; choice = 2; }
    if (tr > best) { best = tr; choice = 3; }
    float p[4];
    if (choice == 3) {
        p[0] = y[2] - z[1]; p[1] = z[0] - x[2]; p[2] = x[1] - y[0]; p[3] = 1.0f + tr;
    } else if (choice == 0) {
        p[0] = 1.0f - tr + 2.0f * x[0]; p[1] = x[1] + y[0]; p[2] = x[2] + z[0]; p[3] = y[2] - z[1];
    } else if (choice == 1) {
        p[1] = 1.0f - tr + 2.0f * y[1]; p[2] = y[2] + z[1]; p[0] = y[0] + x[1]; p[3] = z[0] - x[2];
    } else {
        p[2] = 1.0f - tr + 2.0f * z[2]; p[0] = z[0] + x[2]; p[1] = z[1] + y[2]; p[3] = x[1] - y[0];
    }
    const float pn = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
    const float ph[4] = {p[0] / pn, p[1] / pn, p[2] / pn, p[3] / pn};
    const float qx = qh[1], qy = qh[2], qz = qh[3], qw = qh[0];
    const float ux = ph[3] * qx + qw * ph[0] + (ph[1] * qz - ph[2] * qy);
    const float uy = ph[3] * qy + qw * ph[1] + (ph[2] * qx - ph[0] * qz);
    const float uz = ph[3] * qz + qw * ph[2] + (ph[0] * qy - ph[1] * qx);
    const float uw = ph[3] * qw - (ph[0] * qx + ph[1] * qy + ph[2] * qz);
    const float on_raw = sqrtf(uw * uw + ux * ux + uy * uy + uz * uz);
    // ---- backward
    // (1) oq = (uw, ux, uy, uz) / max(|u|, eps)
    float gw, gx, gy, gz;
    if (on_raw >= eps) {
        const float inv = 1.0f / on_raw;
        const float ow = uw * inv, ox = ux * inv, oy = uy * inv, oz = uz * inv;
        const float dt = v_oq[0] * ow + v_oq[1] * ox + v_oq[2] * oy + v_oq[3] * oz;
        gw = (v_oq[0] - dt * ow) * inv; gx = (v_oq[1] - dt * ox) * inv;
        gy = (v_oq[2] - dt * oy) * inv; gz = (v_oq[3] - dt * oz) * inv;
    } else {
        gw = v_oq[0] / eps; gx = v_oq[1] / eps; gy = v_oq[2] / eps; gz = v_oq[3] / eps;
    }
    // (2) quaternion product u = ph (x) q (xyzw), bilinear
    float v_ph[4];
    v_ph[0] = gx * qw - gy * qz + gz * qy - gw * qx;
    v_ph[1] = gx * qz + gy * qw - gz * qx - gw * qy;
    v_ph[2] = -gx * qy + gy * qx + gz * qw - gw * qz;
    v_ph[3] = gx * qx + gy * qy + gz * qz + gw * qw;
    const float v_qx = gx * ph[3] + gy * ph[2] - gz * ph[1] - gw * ph[0];
    const float v_qy = -gx * ph[2] + gy * ph[3] + gz * ph[0] - gw * ph[1];
    const float v_qz = gx * ph[1] - gy * ph[0] + gz * ph[3] - gw * ph[2];
    const float v_qw = gx * ph[0] + gy * ph[1] + gz * ph[2] + gw * ph[3];
    // (3) q = xyzw(q_hat), q_hat = q_raw / max(|q_raw|, eps)
    {
        const float v_qh[4] = {v_qw, v_qx, v_qy, v_qz};
        if (qn >= eps) {
            const float dt = v_qh[0] * qh[0] + v_qh[1] * qh[1] + v_qh[2] * qh[2] + v_qh[3] * qh[3];
            const float inv = 1.0f / qn;
#pragma unroll
            for (int j = 0; j < 4; ++j) grad[12 + j] = (v_qh[j] - dt * qh[j]) * inv;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) grad[12 + j] = v_qh[j] / eps;
        }
    }
    // (4) ph = p / |p|
    float v_p[4];
    {
        const float dt = v_ph[0] * ph[0] + v_ph[1] * ph[1] + v_ph[2] * ph[2] + v_ph[3] * ph[3];
        const float inv = 1.0f / pn;
#pragma unroll
        for (int j = 0; j < 4; ++j) v_p[j] = (v_ph[j] - dt * ph[j]) * inv;
    }
    // (5) p is linear in the entries of R = [x y z]
    float v_x[3] = {0.f, 0.f, 0.f}, v_y[3] = {0.f, 0.f, 0.f}, v_z[3] = {0.f, 0.f, 0.f};
    if (choice == 3) {
        v_y[2] += v_p[0]; v_z[1] -= v_p[0]; v_z[0] += v_p[1]; v_x[2] -= v_p[1]; v_x[1] += v_p[2]; v_y[0] -= v_p[2];
        v_x[0] += v_p[3]; v_y[1] += v_p[3]; v_z[2] += v_p[3];
    } else if (choice == 0) {
        v_x[0] += v_p[0]; v_y[1] -= v_p[0]; v_z[2] -= v_p[0]; v_x[1] += v_p[1]; v_y[0] += v_p[1];
        v_x[2] += v_p[2]; v_z[0] += v_p[2]; v_y[2] += v_p[3]; v_z[1] -= v_p[3];
    } else if (choice == 1) {
        v_x[0] -= v_p[1]; v_y[1] += v_p[1]; v_z[2] -= v_p[1]; v_y[2] += v_p[2]; v_z[1] += v_p[2];
        v_y[0] += v_p[0]; v_x[1] += v_p[0]; v_z[0] += v_p[3]; v_x[2] -= v_p[3];
    } else {
        v_x[0] -= v_p[2]; v_y[1] -= v_p[2]; v_z[2] += v_p[2]; v_z[0] += v_p[0]; v_x[2] += v_p[0];
        v_z[1] += v_p[1]; v_y[2] += v_p[1]; v_x[1] += v_p[3]; v_y[0] -= v_p[3];
    }
    // (6) om = x mu0 + y mu1 + z mu2 + tl
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        grad[i] = v_om[i];
        v_x[i] += v_om[i] * mu[0];
        v_y[i] += v_om[i] * mu[1];
        v_z[i] += v_om[i] * mu[2];
    }
    grad[9] = x[0] * v_om[0] + x[1] * v_om[1] + x[2] * v_om[2];
    grad[10] = y[0] * v_om[0] + y[1] * v_om[1] + y[2] * v_om[2];
    grad[11] = z[0] * v_om[0] + z[1] * v_om[1] + z[2] * v_om[2];
    // (7) z = x cross y:  v_x += y cross v_z,  v_y += v_z cross x
    v_x[0] += y[1] * v_z[2] - y[2] * v_z[1];
    v_x[1] += y[2] * v_z[0] - y[0] * v_z[2];
    v_x[2] += y[0] * v_z[1] - y[1] * v_z[0];
    v_y[0] += v_z[1] * x[2] - v_z[2] * x[1];
    v_y[1] += v_z[2] * x[0] - v_z[0] * x[2];
    v_y[2] += v_z[0] * x[1] - v_z[1] * x[0];
    // (8) y = yp / max(|yp|, eps)
    float v_yp[3];
    if (yn_raw >= eps) {
        const float dt = v_y[0] * y[0] + v_y[1] * y[1] + v_y[2] * y[2];
        const float inv = 1.0f / yn_raw;
#pragma unroll
        for (int i = 0; i < 3; ++i) v_yp[i] = (v_y[i] - dt * y[i]) * inv;
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) v_yp[i] = v_y[i] / eps;
    }
    // (9) yp = b - (b.x) x
    {
        const float dt = v_yp[0] * x[0] + v_yp[1] * x[1] + v_yp[2] * x[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            grad[6 + i] = v_yp[i] - dt * x[i];
            v_x[i] += -bx * v_yp[i] - dt * b[i];
        }
    }
    // (10) x = a / max(|a|, eps)
    if (an_raw >= eps) {
        const float dt = v_x[0] * x[0] + v_x[1] * x[1] + v_x[2] * x[2];
        const float inv = 1.0f / an_raw;
#pragma unroll
        for (int i = 0; i < 3; ++i) grad[3 + i] = (v_x[i] - dt * x[i]) * inv;
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) grad[3 + i] = v_x[i] / eps;
    }
}

}  // namespace d4
