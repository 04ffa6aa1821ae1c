// Scalar 3-D math shared by every kernel of the hot path (and by tests/host_harness.cpp, which compiles the
// same functions for the host so their logic can be checked without a GPU — the harness is test code only;
// the product never runs these on the CPU).
//
// Reference semantics: articulate/math/{general,angular,spatial}.py, net/smplify/temporal_smplify.py:25-59.
// Matrices are row-major float[9] (or [12] for 3x4 rigid transforms [R | t]).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RC_HD __host__ __device__ __forceinline__
#else
#define RC_HD inline
#endif

// Non-contracted float ops: where the reference evaluates a*b + c*d as separate ATen ops the device code
// must not fuse them into FMAs, or results drift by an ulp and threshold comparisons can flip.
#if defined(__CUDA_ARCH__)
#define RC_MUL(a, b) __fmul_rn((a), (b))
#define RC_ADD(a, b) __fadd_rn((a), (b))
#define RC_SUB(a, b) __fsub_rn((a), (b))
#define RC_DIV(a, b) __fdiv_rn((a), (b))
#else
#define RC_MUL(a, b) ((a) * (b))
#define RC_ADD(a, b) ((a) + (b))
#define RC_SUB(a, b) ((a) - (b))
#define RC_DIV(a, b) ((a) / (b))
#endif

// ---- 3x3 helpers ------------------------------------------------------------------------------------------
RC_HD void rc_mat3_mul(const float* a, const float* b, float* c) {          // c = a @ b
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
}
RC_HD void rc_mat3_tmul(const float* a, const float* b, float* c) {         // c = a^T @ b
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = a[0 * 3 + i] * b[0 * 3 + j] + a[1 * 3 + i] * b[1 * 3 + j] + a[2 * 3 + i] * b[2 * 3 + j];
}
RC_HD void rc_mat3_vec(const float* a, const float* v, float* o) {          // o = a @ v
    for (int i = 0; i < 3; ++i) o[i] = a[i * 3 + 0] * v[0] + a[i * 3 + 1] * v[1] + a[i * 3 + 2] * v[2];
}
RC_HD void rc_vec_mat3(const float* v, const float* a, float* o) {          // o = v @ a  (row vector)
    for (int j = 0; j < 3; ++j) o[j] = v[0] * a[0 * 3 + j] + v[1] * a[1 * 3 + j] + v[2] * a[2 * 3 + j];
}
RC_HD float rc_norm3(const float* v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// rigid 3x4 product: c = a * b with implicit bottom row [0 0 0 1]   (spatial.py:224-249, one tree edge)
RC_HD void rc_rigid_mul(const float* a, const float* b, float* c) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            c[i * 4 + j] = a[i * 4 + 0] * b[0 * 4 + j] + a[i * 4 + 1] * b[1 * 4 + j] + a[i * 4 + 2] * b[2 * 4 + j];
        c[i * 4 + 3] = a[i * 4 + 0] * b[0 * 4 + 3] + a[i * 4 + 1] * b[1 * 4 + 3] + a[i * 4 + 2] * b[2 * 4 + 3] + a[i * 4 + 3];
    }
}

// ---- rotation representations ------------------------------------------------------------------------------
// angular.py:249-264.  The three unit vectors are the COLUMNS of R; any NaN entry becomes 0.
RC_HD void rc_r6d_to_mat(const float* x, float* R) {
    float a0 = x[0], a1 = x[1], a2 = x[2], b0 = x[3], b1 = x[4], b2 = x[5];
    float na = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
    float c00 = RC_DIV(a0, na), c01 = RC_DIV(a1, na), c02 = RC_DIV(a2, na);
    float d = c00 * b0 + c01 * b1 + c02 * b2;
    float u0 = RC_SUB(b0, RC_MUL(d, c00)), u1 = RC_SUB(b1, RC_MUL(d, c01)), u2 = RC_SUB(b2, RC_MUL(d, c02));
    float nu = sqrtf(u0 * u0 + u1 * u1 + u2 * u2);
    float c10 = RC_DIV(u0, nu), c11 = RC_DIV(u1, nu), c12 = RC_DIV(u2, nu);
    float c20 = RC_SUB(RC_MUL(c01, c12), RC_MUL(c02, c11));
    float c21 = RC_SUB(RC_MUL(c02, c10), RC_MUL(c00, c12));
    float c22 = RC_SUB(RC_MUL(c00, c11), RC_MUL(c01, c10));
    float r[9] = {c00, c10, c20, c01, c11, c21, c02, c12, c22};
    for (int i = 0; i < 9; ++i) R[i] = (r[i] != r[i]) ? 0.f : r[i];
}
// angular.py:267-274
RC_HD void rc_mat_to_r6d(const float* R, float* x) {
    x[0] = R[0]; x[1] = R[3]; x[2] = R[6]; x[3] = R[1]; x[4] = R[4]; x[5] = R[7];
}
// angular.py:221-233: c I + (1-c) a a^T + s [a]x, axis NaN/Inf -> 0
RC_HD void rc_aa_to_mat(const float* v, float* R) {
    float ang = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float a[3];
    for (int i = 0; i < 3; ++i) {
        float q = RC_DIV(v[i], ang);
        a[i] = (q != q || isinf(q)) ? 0.f : q;
    }
    float c = cosf(ang), s = sinf(ang), t = 1.f - c;
    float K[9] = {0.f, -a[2], a[1], a[2], 0.f, -a[0], -a[1], a[0], 0.f};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            R[i * 3 + j] = RC_ADD(RC_ADD(i == j ? c : 0.f, RC_MUL(t, RC_MUL(a[i], a[j]))), RC_MUL(s, K[i * 3 + j]));
}
// temporal_smplify.py:25-59: angle = |v + 1e-8|, R = I + sin K + (1 - cos) K K
RC_HD void rc_batch_rodrigues(const float* v, float* R) {
    float e0 = v[0] + 1e-8f, e1 = v[1] + 1e-8f, e2 = v[2] + 1e-8f;
    float ang = sqrtf(e0 * e0 + e1 * e1 + e2 * e2);
    float d0 = RC_DIV(v[0], ang), d1 = RC_DIV(v[1], ang), d2 = RC_DIV(v[2], ang);
    float s = sinf(ang), t = 1.f - cosf(ang);
    float K[9] = {0.f, -d2, d1, d2, 0.f, -d0, -d1, d0, 0.f};
    float KK[9];
    rc_mat3_mul(K, K, KK);
    for (int i = 0; i < 9; ++i) R[i] = RC_ADD(RC_ADD((i % 4 == 0) ? 1.f : 0.f, RC_MUL(s, K[i])), RC_MUL(t, KK[i]));
}
// angular.py:236-246 == cv2.Rodrigues(matrix): SO(3) projection (the orthogonal polar factor U V^T, obtained
// here by Newton's iteration X <- (X + X^-T)/2 instead of an SVD), then the log map with OpenCV's branches
// (s < 1e-5: zero vector, or the theta ~ pi diagonal construction).  float64 inside, like OpenCV.
RC_HD void rc_mat_to_aa(const float* Rin, float* out) {
    double X[9];
    for (int i = 0; i < 9; ++i) X[i] = (double)Rin[i];
    for (int it = 0; it < 40; ++it) {
        double c0 = X[4] * X[8] - X[5] * X[7], c1 = X[5] * X[6] - X[3] * X[8], c2 = X[3] * X[7] - X[4] * X[6];
        double det = X[0] * c0 + X[1] * c1 + X[2] * c2;
        if (fabs(det) < 1e-300) { X[0] = X[4] = X[8] = 1.0; X[1] = X[2] = X[3] = X[5] = X[6] = X[7] = 0.0; break; }
        double inv = 1.0 / det;
        double C[9] = {c0, c1, c2,
                       X[2] * X[7] - X[1] * X[8], X[0] * X[8] - X[2] * X[6], X[1] * X[6] - X[0] * X[7],
                       X[1] * X[5] - X[2] * X[4], X[2] * X[3] - X[0] * X[5], X[0] * X[4] - X[1] * X[3]};
        // cofactor matrix / det == X^-T;  Frobenius-norm scaling accelerates far-from-orthogonal inputs
        double nx = 0, nc = 0;
        for (int i = 0; i < 9; ++i) { nx += X[i] * X[i]; nc += C[i] * C[i] * inv * inv; }
        double g = sqrt(sqrt(nc / nx));
        double delta = 0;
        for (int i = 0; i < 9; ++i) {
            double y = 0.5 * (g * X[i] + C[i] * inv / g);
            delta += (y - X[i]) * (y - X[i]);
            X[i] = y;
        }
        if (delta < 1e-30) break;
    }
    double rx = X[7] - X[5], ry = X[2] - X[6], rz = X[3] - X[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (X[0] + X[4] + X[8] - 1.0) * 0.5;
    c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { rx = ry = rz = 0.0; }
        else {
            double t = (X[0] + 1.0) * 0.5; rx = sqrt(t > 0 ? t : 0.0);
            t = (X[4] + 1.0) * 0.5; ry = sqrt(t > 0 ? t : 0.0) * (X[1] < 0 ? -1.0 : 1.0);
            t = (X[8] + 1.0) * 0.5; rz = sqrt(t > 0 ? t : 0.0) * (X[2] < 0 ? -1.0 : 1.0);
            if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && ((X[5] > 0) != (ry * rz > 0))) rz = -rz;
            theta /= sqrt(rx * rx + ry * ry + rz * rz);
            rx *= theta; ry *= theta; rz *= theta;
        }
    } else {
        double vth = theta / (2.0 * s);
        rx *= vth; ry *= vth; rz *= vth;
    }
    out[0] = (float)rx; out[1] = (float)ry; out[2] = (float)rz;
}
// angular.py:306-318 (wxyz; normalised first)
RC_HD void rc_quat_to_mat(const float* qin, float* R) {
    float n = sqrtf(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
    float a = RC_DIV(qin[0], n), b = RC_DIV(qin[1], n), c = RC_DIV(qin[2], n), d = RC_DIV(qin[3], n);
    R[0] = -2 * c * c - 2 * d * d + 1; R[1] = 2 * b * c - 2 * a * d;      R[2] = 2 * a * c + 2 * b * d;
    R[3] = 2 * b * c + 2 * a * d;      R[4] = -2 * b * b - 2 * d * d + 1; R[5] = 2 * c * d - 2 * a * b;
    R[6] = 2 * b * d - 2 * a * c;      R[7] = 2 * a * b + 2 * c * d;      R[8] = -2 * b * b - 2 * c * c + 1;
}
// angular.py:277-290
RC_HD void rc_quat_to_aa(const float* qin, float* o) {
    float n = sqrtf(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
    float w = RC_DIV(qin[0], n);
    w = w > 1.f ? 1.f : (w < -1.f ? -1.f : w);
    float th = acosf(w), sh = sinf(th);
    for (int i = 0; i < 3; ++i) {
        float v = RC_MUL(RC_MUL(RC_DIV(RC_DIV(qin[i + 1], n), sh), 2.f), th);
        o[i] = (v != v) ? 0.f : v;
    }
}
// angular.py:293-303
RC_HD void rc_aa_to_quat(const float* v, float* q) {
    float ang = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float h = ang / 2.f, s = sinf(h);
    q[0] = cosf(h);
    for (int i = 0; i < 3; ++i) {
        float a = RC_DIV(v[i], ang);
        a = (a != a) ? 0.f : a;
        q[i + 1] = RC_MUL(s, a);
    }
}
// angular.py:79-93
RC_HD void rc_quat_mul(const float* p, const float* q, float* o) {
    float w1 = p[0], x1 = p[1], y1 = p[2], z1 = p[3], w2 = q[0], x2 = q[1], y2 = q[2], z2 = q[3];
    o[0] = RC_SUB(RC_MUL(w1, w2), x1 * x2 + y1 * y2 + z1 * z2);
    o[1] = RC_ADD(RC_ADD(RC_SUB(RC_MUL(y1, z2), RC_MUL(z1, y2)), RC_MUL(w1, x2)), RC_MUL(w2, x1));
    o[2] = RC_ADD(RC_ADD(RC_SUB(RC_MUL(z1, x2), RC_MUL(x1, z2)), RC_MUL(w1, y2)), RC_MUL(w2, y1));
    o[3] = RC_ADD(RC_ADD(RC_SUB(RC_MUL(x1, y2), RC_MUL(y1, x2)), RC_MUL(w1, z2)), RC_MUL(w2, z1));
}
