/* TEST INFRASTRUCTURE -- plain-C oracle for the voxel lift.  NOT product code.
 *
 * Restates reference mmdet3d/models/detectors/nerfdet.py:396-403 (projection and
 * nearest pixel index), :414-416 (gather) and :171-181 (mean / all-view variance /
 * count) with explicit IEEE-754 single-precision operations, independent of any
 * BLAS: the K=4 dot product is the FMA chain of SURVEY.md Appendix A3, the
 * division is a correctly-rounded fp32 divide and the rounding is half-to-even.
 * tests/test_oracle_c.py checks it bit-for-bit against the reference-generated
 * fixtures (tests/golden/lift_*.npz) and against torch.bmm on this host.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off, see oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* q_r = P[r][3] + fma(P[r][2], Z, fma(P[r][1], Y, P[r][0]*X))  (nerfdet.py:398) */
static inline float chain4(const float *p, float x, float y, float z) {
    float t = p[0] * x;
    t = fmaf(p[1], y, t);
    t = fmaf(p[2], z, t);
    return fmaf(p[3], 1.0f, t);
}

/* pix[v*N+n] = y*W+x or -1; q (optional) gets the 3 homogeneous coordinates. */
void nd_oracle_project(const float *points /*[3][N]*/, const float *proj /*[nv][3][4]*/,
                       int nv, int64_t n_vox, int height, int width,
                       int32_t *pix /*[nv][N]*/, float *q /*[nv][3][N] or NULL*/) {
    for (int v = 0; v < nv; ++v) {
        const float *p = proj + (size_t)v * 12;
        for (int64_t n = 0; n < n_vox; ++n) {
            float X = points[n], Y = points[n_vox + n], Z = points[2 * n_vox + n];
            float q0 = chain4(p, X, Y, Z), q1 = chain4(p + 4, X, Y, Z), q2 = chain4(p + 8, X, Y, Z);
            if (q) {
                q[((size_t)v * 3 + 0) * n_vox + n] = q0;
                q[((size_t)v * 3 + 1) * n_vox + n] = q1;
                q[((size_t)v * 3 + 2) * n_vox + n] = q2;
            }
            float xf = nearbyintf(q0 / q2), yf = nearbyintf(q1 / q2);   /* nerfdet.py:400-401 */
            int ok = (xf >= 0.0f) && (yf >= 0.0f) && (xf < (float)width) && (yf < (float)height)
                     && (q2 > 0.0f);                                      /* nerfdet.py:403 */
            pix[(size_t)v * n_vox + n] = ok ? (int32_t)yf * width + (int32_t)xf : -1;
        }
    }
}

/* features: element strides (sv, sc, sy, sx).  mean/cov [C][N], count [N]. */
void nd_oracle_lift(const float *feat, int64_t sv, int64_t sc, int64_t sy, int64_t sx,
                    int nv, int channels, int height, int width,
                    const float *points, const float *proj, int64_t n_vox,
                    float *mean, float *cov, int64_t *count) {
    int32_t *pix = (int32_t *)malloc(sizeof(int32_t) * (size_t)nv * n_vox);
    float *g = (float *)malloc(sizeof(float) * (size_t)nv);
    nd_oracle_project(points, proj, nv, n_vox, height, width, pix, NULL);
    for (int64_t n = 0; n < n_vox; ++n) {
        int64_t c = 0;
        for (int v = 0; v < nv; ++v) c += pix[(size_t)v * n_vox + n] >= 0;
        count[n] = c;
        float denom = (float)c + 1e-8f;                                   /* nerfdet.py:175 */
        for (int ch = 0; ch < channels; ++ch) {
            float s = 0.0f;
            for (int v = 0; v < nv; ++v) {
                int32_t p = pix[(size_t)v * n_vox + n];
                g[v] = p < 0 ? 0.0f : feat[v * sv + ch * sc + (p / width) * sy + (p % width) * sx];
                s += g[v];
            }
            float m = c ? s / denom : 0.0f;                               /* nerfdet.py:175-176 */
            float acc = 0.0f;
            for (int v = 0; v < nv; ++v) { float d = g[v] - m; acc += d * d; }   /* :179, all views */
            float cv = c ? acc / denom : 1e6f;                            /* nerfdet.py:180 */
            mean[(size_t)ch * n_vox + n] = m;
            cov[(size_t)ch * n_vox + n] = expf(-cv);                      /* nerfdet.py:181 */
        }
    }
    free(pix);
    free(g);
}
