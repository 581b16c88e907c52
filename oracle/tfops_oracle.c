/*
 * oracle/tfops_oracle.c -- TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * Scalar CPU restatement of the two point-set ops of kujason/monopsr, used as the
 * parity checker for the sm_100a kernels in monopsr_b200/csrc.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 *
 * Two flavours exist for every op because the reference's CPU and GPU code paths
 * do NOT compute the same thing:
 *
 *   *_gpuorder : mirrors the arithmetic of the reference CUDA kernels
 *                (src/tf_ops/nn_distance/tf_nndistance_g.cu:5-157,
 *                 src/tf_ops/approxmatch/tf_approxmatch_g.cu:1-295):
 *                fp32 throughout, d = fma(dz,dz,fma(dx,dx,dy*dy)) (the contraction
 *                nvcc emits for x*x+y*y+z*z), 10 annealing levels,
 *                match stored (b,m,n) with element [l,k] at l*n+k.
 *                This is the PRIMARY oracle: it is what the product must reproduce
 *                (nearest-neighbour indices bit-exact; EMD within 1e-3 relative,
 *                because the GPU uses ex2.approx/rsqrt.approx and this file uses
 *                libm expf/sqrtf).
 *   *_cpuorder : mirrors the reference CPU kernels
 *                (src/tf_ops/nn_distance/tf_nndistance.cpp:21-43,126-163,
 *                 src/tf_ops/approxmatch/tf_approxmatch.cpp:23-140):
 *                d = (dx*dx+dy*dy)+dz*dz in fp32 widened to double for the compare,
 *                11 levels, double accumulation, match filled as [k*m+l].
 *                Pinned against the verbatim reference functions compiled into
 *                oracle/_ref (see oracle/build_ref.sh) and against the 9 KATs of
 *                tf_nndistance_test.py / tf_approxmatch_test.py.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile). The
 * -ffp-contract=off is load-bearing: every fma below is explicit.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ nn_distance */

/* fp32 squared distance in the GPU kernels' rounding order
 * (tf_nndistance_g.cu:25-28 after nvcc contraction). */
static inline float d2_gpu(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = bx - ax, dy = by - ay, dz = bz - az;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* One direction of NmDistanceKernel (tf_nndistance_g.cu:5-127).  The kernel scans
 * 512-point chunks with strict '<' inside a chunk and strict '>' across chunks, so
 * the globally lowest index among equal minima wins -- a plain ascending scan with
 * strict '<' reproduces it exactly. */
API void nn_search_gpuorder(int b, int n, int m, const float *xyz1, const float *xyz2,
                            float *dist, int *idx) {
    for (int i = 0; i < b; i++) {
        const float *p = xyz1 + (size_t)i * n * 3, *q = xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < n; j++) {
            float best = 0.f;
            int besti = 0;
            for (int k = 0; k < m; k++) {
                float d = d2_gpu(p[j * 3], p[j * 3 + 1], p[j * 3 + 2], q[k * 3], q[k * 3 + 1],
                                 q[k * 3 + 2]);
                if (k == 0 || d < best) {
                    best = d;
                    besti = k;
                }
            }
            dist[(size_t)i * n + j] = best;
            idx[(size_t)i * n + j] = besti;
        }
    }
}

/* CPU nnsearch (tf_nndistance.cpp:21-43): three separately rounded fp32 products,
 * summed in fp32, then widened to double for the comparison. */
API void nn_search_cpuorder(int b, int n, int m, const float *xyz1, const float *xyz2,
                            float *dist, int *idx) {
    for (int i = 0; i < b; i++) {
        const float *p = xyz1 + (size_t)i * n * 3, *q = xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < n; j++) {
            double best = 0;
            int besti = 0;
            for (int k = 0; k < m; k++) {
                float dx = q[k * 3] - p[j * 3], dy = q[k * 3 + 1] - p[j * 3 + 1],
                      dz = q[k * 3 + 2] - p[j * 3 + 2];
                float s = dx * dx + dy * dy;
                s = s + dz * dz;
                double d = s;
                if (k == 0 || d < best) {
                    best = d;
                    besti = k;
                }
            }
            dist[(size_t)i * n + j] = (float)best;
            idx[(size_t)i * n + j] = besti;
        }
    }
}

API void nn_distance_gpuorder(int b, int n, const float *xyz1, int m, const float *xyz2,
                              float *dist1, int *idx1, float *dist2, int *idx2) {
    nn_search_gpuorder(b, n, m, xyz1, xyz2, dist1, idx1);
    nn_search_gpuorder(b, m, n, xyz2, xyz1, dist2, idx2);
}

API void nn_distance_cpuorder(int b, int n, const float *xyz1, int m, const float *xyz2,
                              float *dist1, int *idx1, float *dist2, int *idx2) {
    nn_search_cpuorder(b, n, m, xyz1, xyz2, dist1, idx1);
    nn_search_cpuorder(b, m, n, xyz2, xyz1, dist2, idx2);
}

/* NnDistanceGrad (CPU loops tf_nndistance.cpp:126-163; GPU NmDistanceGradKernel
 * tf_nndistance_g.cu:132-157 computes the same terms with atomics).  Accumulated in
 * double here so the oracle is order-independent; the reference sums in fp32 in a
 * thread-schedule-dependent order, so parity is tolerance-based, not bit-exact. */
API void nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2,
                          const float *gd1, const int *idx1, const float *gd2, const int *idx2,
                          float *gx1, float *gx2) {
    size_t s1 = (size_t)n * 3, s2 = (size_t)m * 3;
    double *a1 = calloc(s1, sizeof(double)), *a2 = calloc(s2, sizeof(double));
    for (int i = 0; i < b; i++) {
        const float *p = xyz1 + i * s1, *q = xyz2 + i * s2;
        memset(a1, 0, s1 * sizeof(double));
        memset(a2, 0, s2 * sizeof(double));
        for (int j = 0; j < n; j++) {
            int j2 = idx1[(size_t)i * n + j];
            float g = gd1[(size_t)i * n + j] * 2;
            for (int c = 0; c < 3; c++) {
                float t = g * (p[j * 3 + c] - q[j2 * 3 + c]);
                a1[j * 3 + c] += t;
                a2[j2 * 3 + c] -= t;
            }
        }
        for (int j = 0; j < m; j++) {
            int j2 = idx2[(size_t)i * m + j];
            float g = gd2[(size_t)i * m + j] * 2;
            for (int c = 0; c < 3; c++) {
                float t = g * (q[j * 3 + c] - p[j2 * 3 + c]);
                a2[j * 3 + c] += t;
                a1[j2 * 3 + c] -= t;
            }
        }
        for (size_t t = 0; t < s1; t++) gx1[i * s1 + t] = (float)a1[t];
        for (size_t t = 0; t < s2; t++) gx2[i * s2 + t] = (float)a2[t];
    }
    free(a1);
    free(a2);
}

/* ------------------------------------------------------------------ approxmatch */

/* GPU-semantics approxmatch (tf_approxmatch_g.cu:1-179): fp32 state, 10 levels
 * j=7..-2 with level=-4^j (0 at j=-2), three sweeps per level, match[l*n+k].
 * expf here is libm's correctly-rounded-ish expf; the reference uses __expf. */
API void approxmatch_gpuorder(int b, int n, int m, const float *xyz1, const float *xyz2,
                              float *match) {
    float multiL, multiR;
    if (n >= m) {
        multiL = 1;
        multiR = (float)(n / m); /* integer division, tf_approxmatch_g.cu:4-10 */
    } else {
        multiL = (float)(m / n);
        multiR = 1;
    }
    float *remainL = malloc(sizeof(float) * n), *remainR = malloc(sizeof(float) * m);
    float *ratioL = malloc(sizeof(float) * n), *ratioR = malloc(sizeof(float) * m);
    for (int i = 0; i < b; i++) {
        const float *p = xyz1 + (size_t)i * n * 3, *q = xyz2 + (size_t)i * m * 3;
        float *mt = match + (size_t)i * n * m;
        memset(mt, 0, sizeof(float) * (size_t)n * m);
        for (int k = 0; k < n; k++) remainL[k] = multiL;
        for (int l = 0; l < m; l++) remainR[l] = multiR;
        for (int j = 7; j >= -2; j--) {
            float level = -powf(4.0f, (float)j);
            if (j == -2) level = 0;
            /* sweep 1 (:26-59) */
            for (int k = 0; k < n; k++) {
                float suml = 1e-9f;
                for (int l = 0; l < m; l++) {
                    float d = level * d2_gpu(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], q[l * 3],
                                             q[l * 3 + 1], q[l * 3 + 2]);
                    suml += expf(d) * remainR[l];
                }
                ratioL[k] = remainL[k] / suml;
            }
            /* sweep 2 (:75-108) */
            for (int l = 0; l < m; l++) {
                float sumr = 0;
                for (int k = 0; k < n; k++) {
                    float d = level * d2_gpu(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], q[l * 3],
                                             q[l * 3 + 1], q[l * 3 + 2]);
                    sumr += expf(d) * ratioL[k];
                }
                sumr *= remainR[l];
                float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
                ratioR[l] = consumption * remainR[l];
                remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
            }
            /* sweep 3 (:127-160) */
            for (int k = 0; k < n; k++) {
                float suml = 0;
                for (int l = 0; l < m; l++) {
                    float d = level * d2_gpu(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], q[l * 3],
                                             q[l * 3 + 1], q[l * 3 + 2]);
                    float w = expf(d) * ratioL[k] * ratioR[l];
                    mt[(size_t)l * n + k] += w;
                    suml += w;
                }
                remainL[k] = fmaxf(0.0f, remainL[k] - suml);
            }
        }
    }
    free(remainL);
    free(remainR);
    free(ratioL);
    free(ratioR);
}

/* matchcost kernel (tf_approxmatch_g.cu:183-225): fp32 sqrtf per pair, match read
 * as [l*n+k].  Accumulated in double so the oracle has no order dependence. */
API void matchcost_gpuorder(int b, int n, int m, const float *xyz1, const float *xyz2,
                            const float *match, float *cost) {
    for (int i = 0; i < b; i++) {
        const float *p = xyz1 + (size_t)i * n * 3, *q = xyz2 + (size_t)i * m * 3;
        const float *mt = match + (size_t)i * n * m;
        double s = 0;
        for (int l = 0; l < m; l++)
            for (int k = 0; k < n; k++) {
                float d = sqrtf(d2_gpu(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], q[l * 3],
                                       q[l * 3 + 1], q[l * 3 + 2]));
                s += (double)(d * mt[(size_t)l * n + k]);
            }
        cost[i] = (float)s;
    }
}

/* matchcostgrad1/2 (tf_approxmatch_g.cu:229-291): unit vector via
 * rsqrt(max(d2,1e-20)); double accumulation. */
API void matchcostgrad_gpuorder(int b, int n, int m, const float *xyz1, const float *xyz2,
                                const float *match, float *grad1, float *grad2) {
    for (int i = 0; i < b; i++) {
        const float *p = xyz1 + (size_t)i * n * 3, *q = xyz2 + (size_t)i * m * 3;
        const float *mt = match + (size_t)i * n * m;
        float *g1 = grad1 + (size_t)i * n * 3, *g2 = grad2 + (size_t)i * m * 3;
        for (int k = 0; k < n; k++) {
            double a[3] = {0, 0, 0};
            for (int l = 0; l < m; l++) {
                float dx = p[k * 3] - q[l * 3], dy = p[k * 3 + 1] - q[l * 3 + 1],
                      dz = p[k * 3 + 2] - q[l * 3 + 2];
                float d2 = fmaxf(dx * dx + dy * dy + dz * dz, 1e-20f);
                double w = (double)mt[(size_t)l * n + k] / sqrt((double)d2);
                a[0] += dx * w;
                a[1] += dy * w;
                a[2] += dz * w;
            }
            for (int c = 0; c < 3; c++) g1[k * 3 + c] = (float)a[c];
        }
        for (int l = 0; l < m; l++) {
            double a[3] = {0, 0, 0};
            for (int k = 0; k < n; k++) {
                float dx = q[l * 3] - p[k * 3], dy = q[l * 3 + 1] - p[k * 3 + 1],
                      dz = q[l * 3 + 2] - p[k * 3 + 2];
                float d2 = fmaxf(dx * dx + dy * dy + dz * dz, 1e-20f);
                double w = (double)mt[(size_t)l * n + k] / sqrt((double)d2);
                a[0] += dx * w;
                a[1] += dy * w;
                a[2] += dz * w;
            }
            for (int c = 0; c < 3; c++) g2[l * 3 + c] = (float)a[c];
        }
    }
}

/* CPU-semantics approxmatch (tf_approxmatch.cpp:23-84): 11 levels j=8..-2, double
 * state, match filled as [k*m+l] (n-major, quirk Q3).  Restated sweep by sweep. */
API void approxmatch_cpuorder(int b, int n, int m, const float *xyz1, const float *xyz2,
                              float *match) {
    int big = n > m ? n : m;
    double *satl = malloc(sizeof(double) * n), *satr = malloc(sizeof(double) * m);
    double *w = malloc(sizeof(double) * (size_t)n * m);
    double *colsum = malloc(sizeof(double) * m), *colsum2 = malloc(sizeof(double) * m);
    for (int i = 0; i < b; i++) {
        const float *p = xyz1 + (size_t)i * n * 3, *q = xyz2 + (size_t)i * m * 3;
        float *mt = match + (size_t)i * n * m;
        for (int k = 0; k < n; k++) satl[k] = (double)(big / n);
        for (int l = 0; l < m; l++) satr[l] = (double)(big / m);
        for (size_t t = 0; t < (size_t)n * m; t++) mt[t] = 0;
        for (int j = 8; j >= -2; j--) {
            double level = -powf(4.0f, (float)j);
            if (j == -2) level = 0;
            for (int k = 0; k < n; k++) {
                double x1 = p[k * 3], y1 = p[k * 3 + 1], z1 = p[k * 3 + 2];
                for (int l = 0; l < m; l++) {
                    double x2 = q[l * 3], y2 = q[l * 3 + 1], z2 = q[l * 3 + 2];
                    double arg = level * ((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) +
                                          (z1 - z2) * (z1 - z2));
                    w[(size_t)k * m + l] = expf((float)arg) * satr[l]; /* expf: cpp:44 */
                }
            }
            for (int l = 0; l < m; l++) colsum[l] = 1e-9;
            for (int k = 0; k < n; k++) {
                double s = 1e-9;
                for (int l = 0; l < m; l++) s += w[(size_t)k * m + l];
                for (int l = 0; l < m; l++) {
                    w[(size_t)k * m + l] = w[(size_t)k * m + l] / s * satl[k];
                    colsum[l] += w[(size_t)k * m + l];
                }
            }
            for (int l = 0; l < m; l++) {
                double r = satr[l] / colsum[l];
                colsum[l] = r < 1.0 ? r : 1.0;
                colsum2[l] = 0;
            }
            for (int k = 0; k < n; k++) {
                double s = 0;
                for (int l = 0; l < m; l++) {
                    w[(size_t)k * m + l] *= colsum[l];
                    s += w[(size_t)k * m + l];
                    colsum2[l] += w[(size_t)k * m + l];
                }
                satl[k] = satl[k] - s > 0.0 ? satl[k] - s : 0.0;
            }
            for (size_t t = 0; t < (size_t)n * m; t++) mt[t] += (float)w[t];
            for (int l = 0; l < m; l++)
                satr[l] = satr[l] - colsum2[l] > 0.0 ? satr[l] - colsum2[l] : 0.0;
        }
    }
    free(satl);
    free(satr);
    free(w);
    free(colsum);
    free(colsum2);
}

/* matchcost_cpu (tf_approxmatch.cpp:85-105): match read as [j*m+k] (n-major). */
API void matchcost_cpuorder(int b, int n, int m, const float *xyz1, const float *xyz2,
                            const float *match, float *cost) {
    for (int i = 0; i < b; i++) {
        const float *p = xyz1 + (size_t)i * n * 3, *q = xyz2 + (size_t)i * m * 3;
        const float *mt = match + (size_t)i * n * m;
        double s = 0;
        for (int k = 0; k < n; k++)
            for (int l = 0; l < m; l++) {
                float dx = q[l * 3] - p[k * 3], dy = q[l * 3 + 1] - p[k * 3 + 1],
                      dz = q[l * 3 + 2] - p[k * 3 + 2];
                float d = sqrtf(dx * dx + dy * dy + dz * dz) * mt[(size_t)k * m + l];
                s += d;
            }
        cost[i] = (float)s;
    }
}

/* matchcostgrad_cpu (tf_approxmatch.cpp:106-140) WITH grad1 fully zero-initialised
 * (the reference only zeroes the x component, quirk Q4) and the norm clamped at
 * 1e-20 as in cpp:119.  n-major match. */
API void matchcostgrad_cpuorder(int b, int n, int m, const float *xyz1, const float *xyz2,
                                const float *match, float *grad1, float *grad2) {
    for (int i = 0; i < b; i++) {
        const float *p = xyz1 + (size_t)i * n * 3, *q = xyz2 + (size_t)i * m * 3;
        const float *mt = match + (size_t)i * n * m;
        float *g1 = grad1 + (size_t)i * n * 3, *g2 = grad2 + (size_t)i * m * 3;
        for (int t = 0; t < n * 3; t++) g1[t] = 0;
        for (int l = 0; l < m; l++) {
            float s[3] = {0, 0, 0};
            for (int k = 0; k < n; k++) {
                float dx = q[l * 3] - p[k * 3], dy = q[l * 3 + 1] - p[k * 3 + 1],
                      dz = q[l * 3 + 2] - p[k * 3 + 2];
                float d = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-20f);
                float wv = mt[(size_t)k * m + l];
                float c[3] = {wv * (dx / d), wv * (dy / d), wv * (dz / d)};
                for (int a = 0; a < 3; a++) {
                    g1[k * 3 + a] -= c[a];
                    s[a] += c[a];
                }
            }
            for (int a = 0; a < 3; a++) g2[l * 3 + a] = s[a];
        }
    }
}
