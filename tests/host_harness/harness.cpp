// tests/host_harness/harness.cpp -- TEST ONLY.
// Compiles the product's __host__ __device__ arithmetic headers
// (csrc/project_math.cuh, csrc/deform_math.cuh) for the CPU with g++ so that
// the exact per-Gaussian math the CUDA kernels run can be checked against the
// oracle in the GPU-less container (-m "not gpu").  Nothing here is shipped or
// reachable from the product path.
#include <stdint.h>
#include <string.h>

#include "../../deblur4dgs_b200/csrc/project_math.cuh"
#include "../../deblur4dgs_b200/csrc/deform_math.cuh"
#include "../../deblur4dgs_b200/csrc/camera_math.cuh"

using namespace d4;

extern "C" {

void hh_project_fwd(const float *means, const float *quats, const float *scales, const float *V, const float *K,
                    int G, int width, int height, float eps2d, float near_plane, float far_plane, float radius_clip,
                    int tile_size, int tile_w, int tile_h, int32_t *radii, float *means2d, float *depths,
                    float *conics, int32_t *tiles_per_gauss) {
    for (int g = 0; g < G; ++g) {
        ProjOut o = project_one(means + 3 * g, quats + 4 * g, scales + 3 * g, V, K, width, height, eps2d,
                                near_plane, far_plane, radius_clip);
        radii[g] = o.radius;
        means2d[2 * g] = o.m2x; means2d[2 * g + 1] = o.m2y;
        depths[g] = o.depth;
        conics[3 * g] = o.ca; conics[3 * g + 1] = o.cb; conics[3 * g + 2] = o.cc;
        int n = 0;
        if (o.radius > 0) {
            int x0, y0, x1, y1;
            tile_rect(o.m2x, o.m2y, o.radius, tile_size, tile_w, tile_h, &x0, &y0, &x1, &y1);
            n = (y1 - y0) * (x1 - x0);
        }
        tiles_per_gauss[g] = n;
    }
}

void hh_project_bwd(const float *means, const float *quats, const float *scales, const float *V, const float *K,
                    int G, int width, int height, const int32_t *radii, const float *conics,
                    const float *v_means2d, const float *v_depths, const float *v_conics, float *v_means,
                    float *v_quats, float *v_scales, double *v_viewmat /*[16]*/) {
    for (int g = 0; g < G; ++g) {
        if (radii[g] <= 0) continue;
        ProjGrad pg;
        project_one_bwd(means + 3 * g, quats + 4 * g, scales + 3 * g, V, K, width, height, conics[3 * g],
                        conics[3 * g + 1], conics[3 * g + 2], v_means2d[2 * g], v_means2d[2 * g + 1], v_depths[g],
                        v_conics[3 * g], v_conics[3 * g + 1], v_conics[3 * g + 2], &pg);
        for (int j = 0; j < 3; ++j) v_means[3 * g + j] = pg.v_mean[j];
        for (int j = 0; j < 4; ++j) v_quats[4 * g + j] = pg.v_quat[j];
        for (int j = 0; j < 3; ++j) v_scales[3 * g + j] = pg.v_scale[j];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) v_viewmat[4 * i + j] += pg.v_R[3 * i + j];
            v_viewmat[4 * i + 3] += pg.v_t[i];
        }
    }
}

// blended inputs bl[9] = (tl 3, r6 6), mu[3], q_raw[4] -> out mu'[3], q'[4]
void hh_deform_point(const float *bl, const float *mu, const float *q, int n, float *om, float *oq) {
    for (int i = 0; i < n; ++i)
        deform_point<float>(bl + 9 * i, bl + 9 * i + 3, mu + 3 * i, q + 4 * i, om + 3 * i, oq + 4 * i);
}

// vector-Jacobian product through the dual-number path: grad[16] per point
void hh_deform_point_vjp(const float *bl, const float *mu, const float *q, const float *vm, const float *vq, int n,
                         float *grad) {
    typedef Dual<16> DU;
    for (int i = 0; i < n; ++i) {
        DU in[16];
        for (int a = 0; a < 16; ++a)
            for (int b = 0; b < 16; ++b) in[a].d[b] = (a == b) ? 1.f : 0.f;
        for (int j = 0; j < 9; ++j) in[j].v = bl[9 * i + j];
        for (int j = 0; j < 3; ++j) in[9 + j].v = mu[3 * i + j];
        for (int j = 0; j < 4; ++j) in[12 + j].v = q[4 * i + j];
        DU om[3], oq[4];
        deform_point<DU>(in, in + 3, in + 9, in + 12, om, oq);
        for (int a = 0; a < 16; ++a) {
            float acc = 0.f;
            for (int j = 0; j < 3; ++j) acc += vm[3 * i + j] * om[j].d[a];
            for (int j = 0; j < 4; ++j) acc += vq[4 * i + j] * oq[j].d[a];
            grad[16 * i + a] = acc;
        }
    }
}

// the hand-derived reverse-mode VJP of the same function (what the backward kernels run)
void hh_deform_point_vjp_rev(const float *bl, const float *mu, const float *q, const float *vm, const float *vq, int n,
                             float *grad, float *om) {
    for (int i = 0; i < n; ++i)
        deform_point_vjp(bl + 9 * i, bl + 9 * i + 3, mu + 3 * i, q + 4 * i, vm + 3 * i, vq + 4 * i, om + 3 * i,
                         grad + 16 * i);
}

// camera interpolation (row a7): forward and dual-number VJP for sub-exposure parameter u
void hh_camera_interp(const float *start6, const float *end6, const float *us, int n, float *Rt) {
    for (int i = 0; i < n; ++i) camera_interp_one<float>(start6, end6, us[i], Rt + 12 * i);
}
void hh_camera_interp_vjp(const float *start6, const float *end6, const float *us, int n, const float *v,
                          float *g12) {
    typedef Dual<12> DU;
    for (int j = 0; j < 12; ++j) g12[j] = 0.f;
    for (int i = 0; i < n; ++i) {
        DU s[6], e[6], Rt[12];
        for (int a = 0; a < 6; ++a) {
            s[a].v = start6[a]; e[a].v = end6[a];
            for (int j = 0; j < 12; ++j) { s[a].d[j] = (j == a) ? 1.f : 0.f; e[a].d[j] = (j == 6 + a) ? 1.f : 0.f; }
        }
        camera_interp_one<DU>(s, e, us[i], Rt);
        for (int o = 0; o < 12; ++o)
            for (int j = 0; j < 12; ++j) g12[j] += v[12 * i + o] * Rt[o].d[j];
    }
}
void hh_se3_to_SE3(const float *wu, int n, float *Rt) {
    for (int i = 0; i < n; ++i) se3_to_SE3_mat<float>(wu + 6 * i, wu + 6 * i + 3, Rt + 12 * i);
}
}
