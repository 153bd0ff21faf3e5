/*
 * tvl1_oracle.c -- CPU restatement of cv::DualTVL1OpticalFlow (OpenCV 3.4.1,
 * modules/video/src/tvl1flow.cpp) plus the wrapper semantics of fibsem-optflow's
 * flow stage.  TEST INFRASTRUCTURE ONLY -- see tvl1_oracle.h for the rules and for
 * the "parity unpinned" statement.
 *
 * Every function cites the SURVEY.md appendix section it follows (the OpenCV
 * source is not vendored by the reference) and, for wrapper code, the reference
 * file:line.  All per-pixel arithmetic is IEEE fp32 with no contraction: build
 * with -ffp-contract=off and without -ffast-math (oracle/Makefile does).
 *
 * Structure mirrors OpenCV's CPU class: seven separate row-parallel passes per
 * inner iteration (estimateV, 2x divergence, estimateU, 2x forwardGradient,
 * estimateDualVariables), which is also what makes it the timed CPU baseline.
 */
#include "tvl1_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_LEVELS 64

void orc_default_params(orc_params* p)
{
    /* SURVEY A.1, CPU class defaults */
    p->tau = 0.25; p->lambda = 0.15; p->theta = 0.3; p->epsilon = 0.01;
    p->scale_step = 0.8; p->gamma = 0.0;
    p->nscales = 5; p->warps = 5; p->inner_iterations = 30; p->outer_iterations = 10;
    p->median_filtering = 5; p->error_sum_mode = 0; p->nthreads = 0;
}

/* cvRound: round half to even */
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_floor_f(float v)
{
    int i = (int)v;
    return i - (v < (float)i);
}

int orc_scaled_size(int n, double f)
{
    return cv_round_d((double)n * f);
}

/* ---------------------------------------------------------------- A.2 resize */

void orc_resize_linear(const float* src, int sw, int sh, float* dst, int dw, int dh,
                       double inv_scale)
{
    if (inv_scale == 0.5) {
        /* resize(src, Size(), 0.5, 0.5, INTER_LINEAR) is switched to INTER_AREA by cv::resize (scale exactly
         * 2: the "area fast" path).  The mean of a 2x2 block is ((a + b) + (c + d)) * 0.25f in the SIMD part of
         * a row (four destination pixels at a time, ResizeAreaFastVec_SIMD_32f) and (((a + b) + c) + d) * 0.25f
         * for the up to three whole blocks left over; a block that hangs over the source (destination size
         * rounded up) is the mean of the pixels that exist, summed in row-major order and divided by their
         * count.  Pinned against cv2 on 300 random sizes (tests/test_oracle_primitives.py). */
        const int vec = (sw / 2) / 4 * 4;   /* destination columns the 4-wide loop covers */
#pragma omp parallel for schedule(static)
        for (int dy = 0; dy < dh; dy++)
            for (int dx = 0; dx < dw; dx++) {
                const int sx = 2 * dx, sy = 2 * dy;
                float d;
                if (sx + 1 < sw && sy + 1 < sh) {
                    const float* S0 = src + (size_t)sy * sw + sx;
                    const float* S1 = S0 + sw;
                    d = dx < vec ? ((S0[0] + S0[1]) + (S1[0] + S1[1])) * 0.25f
                                 : (S0[0] + S0[1] + S1[0] + S1[1]) * 0.25f;
                } else if (sx >= sw || sy >= sh) {
                    d = 0.f;
                } else {
                    float sum = 0.f;
                    int count = 0;
                    for (int r = 0; r < 2 && sy + r < sh; r++)
                        for (int c = 0; c < 2 && sx + c < sw; c++) { sum += src[(size_t)(sy + r) * sw + sx + c]; count++; }
                    d = sum / (float)count;
                }
                dst[(size_t)dy * dw + dx] = d;
            }
        return;
    }
    double inv_x, inv_y;
    if (inv_scale > 0) { inv_x = inv_scale; inv_y = inv_scale; }
    else { inv_x = (double)dw / sw; inv_y = (double)dh / sh; }
    const double scale_x = 1. / inv_x, scale_y = 1. / inv_y;

    int* xofs = (int*)malloc(sizeof(int) * (size_t)dw);
    float* alpha = (float*)malloc(sizeof(float) * 2 * (size_t)dw);
    int xmax = dw;
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor_f(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx + 1 >= sw) {
            if (dx < xmax) xmax = dx;
            if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        }
        xofs[dx] = sx;
        alpha[2 * dx] = 1.f - fx;
        alpha[2 * dx + 1] = fx;
    }

#pragma omp parallel for schedule(static)
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor_f(fy);
        fy -= sy;
        const float b0 = 1.f - fy, b1 = fy;
        int r0 = sy, r1 = sy + 1;
        if (r0 < 0) r0 = 0; if (r0 > sh - 1) r0 = sh - 1;
        if (r1 < 0) r1 = 0; if (r1 > sh - 1) r1 = sh - 1;
        const float* S0 = src + (size_t)r0 * sw;
        const float* S1 = src + (size_t)r1 * sw;
        float* D = dst + (size_t)dy * dw;
        int dx = 0;
        for (; dx < xmax; dx++) {
            const int sx = xofs[dx];
            const float a0 = alpha[2 * dx], a1 = alpha[2 * dx + 1];
            const float h0 = S0[sx] * a0 + S0[sx + 1] * a1;
            const float h1 = S1[sx] * a0 + S1[sx + 1] * a1;
            D[dx] = h0 * b0 + h1 * b1;
        }
        for (; dx < dw; dx++) {
            const int sx = xofs[dx];
            const float h0 = S0[sx] * 1.f;
            const float h1 = S1[sx] * 1.f;
            D[dx] = h0 * b0 + h1 * b1;
        }
    }
    free(xofs);
    free(alpha);
}

void orc_convert_u8(const unsigned char* src, long pitch, int w, int h, float* dst)
{
    /* A.2: I.convertTo(CV_32F, 1.0) for 8-bit input */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            dst[(size_t)y * w + x] = (float)src[(size_t)y * pitch + x];
}

/* ------------------------------------------------------ A.3 centred gradient */

void orc_centered_gradient(const float* src, int w, int h, float* dx, float* dy)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const int ym = y > 0 ? y - 1 : 0, yp = y < h - 1 ? y + 1 : h - 1;
        const float* c = src + (size_t)y * w;
        const float* a = src + (size_t)ym * w;
        const float* b = src + (size_t)yp * w;
        float* ox = dx + (size_t)y * w;
        float* oy = dy + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            const int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
            ox[x] = 0.5f * (c[xp] - c[xm]);
            oy[x] = 0.5f * (b[x] - a[x]);
        }
    }
}

/* ------------------------------------------------------------ A.4 cubic remap */

static float g_cubic_tab[32][4];
static int g_cubic_tab_ready = 0;

static void cubic_coeffs(float x, float* c)
{
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

static void cubic_tab_init(void)
{
    if (g_cubic_tab_ready) return;
    const float scale = 1.f / 32;
    for (int i = 0; i < 32; i++) cubic_coeffs(i * scale, g_cubic_tab[i]);
    g_cubic_tab_ready = 1;
}

static inline short sat_short(int v)
{
    return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
}

/* one output value of remap(INTER_CUBIC, BORDER_CONSTANT 0) for an fp32 plane */
static inline float remap_cubic_px(const float* src, int w, int h, float mx, float my)
{
    const int qx = cv_round_f(mx * 32), qy = cv_round_f(my * 32);
    const int sx = sat_short(qx >> 5) - 1, sy = sat_short(qy >> 5) - 1;
    const float* cx = g_cubic_tab[qx & 31];
    const float* cy = g_cubic_tab[qy & 31];
    const unsigned width1 = w - 3 > 0 ? (unsigned)(w - 3) : 0u;
    const unsigned height1 = h - 3 > 0 ? (unsigned)(h - 3) : 0u;

    if ((unsigned)sx < width1 && (unsigned)sy < height1) {
        /* interior: per-row grouped sums */
        const float* S = src + (size_t)sy * w + sx;
        float sum = S[0] * (cy[0] * cx[0]) + S[1] * (cy[0] * cx[1]) + S[2] * (cy[0] * cx[2]) +
                    S[3] * (cy[0] * cx[3]);
        S += w;
        sum += S[0] * (cy[1] * cx[0]) + S[1] * (cy[1] * cx[1]) + S[2] * (cy[1] * cx[2]) +
               S[3] * (cy[1] * cx[3]);
        S += w;
        sum += S[0] * (cy[2] * cx[0]) + S[1] * (cy[2] * cx[1]) + S[2] * (cy[2] * cx[2]) +
               S[3] * (cy[2] * cx[3]);
        S += w;
        sum += S[0] * (cy[3] * cx[0]) + S[1] * (cy[3] * cx[1]) + S[2] * (cy[3] * cx[2]) +
               S[3] * (cy[3] * cx[3]);
        return sum;
    }
    if (sx >= w || sx + 4 <= 0 || sy >= h || sy + 4 <= 0) return 0.f;
    /* border: one tap at a time, taps outside the image skipped (cval = 0) */
    float sum = 0.f;
    for (int i = 0; i < 4; i++) {
        const int yi = sy + i;
        if (yi < 0 || yi >= h) continue;
        const float* S = src + (size_t)yi * w;
        for (int j = 0; j < 4; j++) {
            const int xj = sx + j;
            if (xj >= 0 && xj < w) sum += (S[xj] - 0.f) * (cy[i] * cx[j]);
        }
    }
    return sum;
}

void orc_remap_cubic(const float* src, int w, int h, const float* mapx, const float* mapy,
                     float* dst)
{
    cubic_tab_init();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t i = (size_t)y * w + x;
            dst[i] = remap_cubic_px(src, w, h, mapx[i], mapy[i]);
        }
}

void orc_warp(const float* I0, const float* I1, const float* I1x, const float* I1y,
              const float* u1, const float* u2, int w, int h,
              float* I1w, float* I1wx, float* I1wy, float* grad, float* rho_c)
{
    cubic_tab_init();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t i = (size_t)y * w + x;
            /* buildFlowMap */
            const float mx = (float)x + u1[i], my = (float)y + u2[i];
            const float iw = remap_cubic_px(I1, w, h, mx, my);
            const float iwx = remap_cubic_px(I1x, w, h, mx, my);
            const float iwy = remap_cubic_px(I1y, w, h, mx, my);
            /* calcGradRho */
            const float Ix2 = iwx * iwx;
            const float Iy2 = iwy * iwy;
            if (I1w) I1w[i] = iw;
            I1wx[i] = iwx;
            I1wy[i] = iwy;
            grad[i] = Ix2 + Iy2;
            rho_c[i] = (iw - iwx * u1[i] - iwy * u2[i] - I0[i]);
        }
}

/* --------------------------------------------------------- A.5 inner iteration */

typedef struct {
    float *v1, *v2, *div1, *div2, *u1x, *u1y, *u2x, *u2y;
    float *v3, *div3, *u3x, *u3y;   /* gamma != 0 only (allocated on first use) */
    double* rowsum;
    size_t cap_px;
    int cap_rows;
} iter_ws;

static void ws_free(iter_ws* s)
{
    free(s->v1); free(s->v2); free(s->div1); free(s->div2);
    free(s->u1x); free(s->u1y); free(s->u2x); free(s->u2y); free(s->rowsum);
    free(s->v3); free(s->div3); free(s->u3x); free(s->u3y);
    memset(s, 0, sizeof(*s));
}

static int ws_reserve(iter_ws* s, int w, int h)
{
    const size_t n = (size_t)w * h;
    if (n <= s->cap_px && h <= s->cap_rows) return 0;
    ws_free(s);
    s->v1 = (float*)malloc(n * 4); s->v2 = (float*)malloc(n * 4);
    s->div1 = (float*)malloc(n * 4); s->div2 = (float*)malloc(n * 4);
    s->u1x = (float*)malloc(n * 4); s->u1y = (float*)malloc(n * 4);
    s->u2x = (float*)malloc(n * 4); s->u2y = (float*)malloc(n * 4);
    s->rowsum = (double*)malloc(sizeof(double) * (size_t)h);
    if (!s->v1 || !s->v2 || !s->div1 || !s->div2 || !s->u1x || !s->u1y || !s->u2x ||
        !s->u2y || !s->rowsum) {
        ws_free(s);
        return -1;
    }
    s->cap_px = n;
    s->cap_rows = h;
    return 0;
}

static void estimate_v(const float* I1wx, const float* I1wy, const float* u1, const float* u2,
                       const float* grad, const float* rho_c, float* v1, float* v2,
                       int w, int h, float l_t)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const size_t o = (size_t)y * w;
        for (int x = 0; x < w; x++) {
            const size_t i = o + x;
            const float rho = rho_c[i] + (I1wx[i] * u1[i] + I1wy[i] * u2[i]);
            float d1 = 0.0f, d2 = 0.0f;
            if (rho < -l_t * grad[i]) {
                d1 = l_t * I1wx[i];
                d2 = l_t * I1wy[i];
            } else if (rho > l_t * grad[i]) {
                d1 = -l_t * I1wx[i];
                d2 = -l_t * I1wy[i];
            } else if (grad[i] > FLT_EPSILON) {
                const float fi = -rho / grad[i];
                d1 = fi * I1wx[i];
                d2 = fi * I1wy[i];
            }
            v1[i] = u1[i] + d1;
            v2[i] = u2[i] + d2;
        }
    }
}

static void divergence(const float* a, const float* b, float* div, int w, int h)
{
#pragma omp parallel for schedule(static)
    for (int y = 1; y < h; y++) {
        const float* ar = a + (size_t)y * w;
        const float* bc = b + (size_t)y * w;
        const float* bp = b + (size_t)(y - 1) * w;
        float* d = div + (size_t)y * w;
        for (int x = 1; x < w; x++) {
            const float v1x = ar[x] - ar[x - 1];
            const float v2y = bc[x] - bp[x];
            d[x] = v1x + v2y;
        }
    }
    for (int x = 1; x < w; x++) div[x] = a[x] - a[x - 1] + b[x];
    for (int y = 1; y < h; y++)
        div[(size_t)y * w] = a[(size_t)y * w] + b[(size_t)y * w] - b[(size_t)(y - 1) * w];
    div[0] = a[0] + b[0];
}

static double estimate_u(const float* v1, const float* v2, const float* div1, const float* div2,
                         float* u1, float* u2, int w, int h, float theta, int mode,
                         double* rowsum)
{
    if (mode == 1) {
        /* OpenCV literal: one fp32 scalar, row-major, serial */
        float error = 0.0f;
        for (int y = 0; y < h; y++) {
            const size_t o = (size_t)y * w;
            for (int x = 0; x < w; x++) {
                const size_t i = o + x;
                const float u1k = u1[i], u2k = u2[i];
                u1[i] = v1[i] + theta * div1[i];
                u2[i] = v2[i] + theta * div2[i];
                error += (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k);
            }
        }
        return (double)error;
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const size_t o = (size_t)y * w;
        double acc = 0.0;
        for (int x = 0; x < w; x++) {
            const size_t i = o + x;
            const float u1k = u1[i], u2k = u2[i];
            u1[i] = v1[i] + theta * div1[i];
            u2[i] = v2[i] + theta * div2[i];
            const float term = (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k);
            acc += (double)term;
        }
        rowsum[y] = acc;
    }
    double error = 0.0;
    for (int y = 0; y < h; y++) error += rowsum[y];
    return error;
}

static void forward_gradient(const float* src, float* dx, float* dy, int w, int h)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const float* c = src + (size_t)y * w;
        const float* n = src + (size_t)(y < h - 1 ? y + 1 : y) * w;
        float* ox = dx + (size_t)y * w;
        float* oy = dy + (size_t)y * w;
        for (int x = 0; x < w - 1; x++) ox[x] = c[x + 1] - c[x];
        ox[w - 1] = 0.f;
        if (y < h - 1)
            for (int x = 0; x < w; x++) oy[x] = n[x] - c[x];
        else
            for (int x = 0; x < w; x++) oy[x] = 0.f;
    }
}

/* canonical hypot: the products are exact in double, one rounding in
 * the sum, one in sqrt, one in the cast -- this is also glibc's hypotf */
static inline float hypot_f(float a, float b)
{
    return (float)sqrt((double)a * (double)a + (double)b * (double)b);
}

static void estimate_dual(const float* u1x, const float* u1y, const float* u2x, const float* u2y,
                          float* p11, float* p12, float* p21, float* p22, int w, int h,
                          float taut)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const size_t o = (size_t)y * w;
        for (int x = 0; x < w; x++) {
            const size_t i = o + x;
            const float g1 = hypot_f(u1x[i], u1y[i]);
            const float g2 = hypot_f(u2x[i], u2y[i]);
            const float ng1 = 1.0f + taut * g1;
            const float ng2 = 1.0f + taut * g2;
            p11[i] = (p11[i] + taut * u1x[i]) / ng1;
            p12[i] = (p12[i] + taut * u1y[i]) / ng1;
            p21[i] = (p21[i] + taut * u2x[i]) / ng2;
            p22[i] = (p22[i] + taut * u2y[i]) / ng2;
        }
    }
}

/* ---- gamma != 0: the third channel u3 / p31, p32 (illumination term) of OpenCV 3.4.1's
 * tvl1flow.cpp (EstimateVBody, EstimateUBody, EstimateDualVariablesBody with use_gamma).
 * As recalled: rho gains + gamma*u3 (added last), d3 = +-l_t*gamma or fi*gamma,
 * v3 = u3 + d3, u3' = v3 + theta*div(p31, p32), the error term gains (u3' - u3)^2 (added last),
 * p31/p32 are updated like the other dual variables; grad and rho_c do not involve gamma. */
static void estimate_v_g(const float* I1wx, const float* I1wy, const float* u1, const float* u2,
                         const float* u3, const float* grad, const float* rho_c, float* v1,
                         float* v2, float* v3, int w, int h, float l_t, float gamma)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const size_t o = (size_t)y * w;
        for (int x = 0; x < w; x++) {
            const size_t i = o + x;
            const float rho = rho_c[i] + (I1wx[i] * u1[i] + I1wy[i] * u2[i]) + gamma * u3[i];
            float d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
            if (rho < -l_t * grad[i]) {
                d1 = l_t * I1wx[i];
                d2 = l_t * I1wy[i];
                d3 = l_t * gamma;
            } else if (rho > l_t * grad[i]) {
                d1 = -l_t * I1wx[i];
                d2 = -l_t * I1wy[i];
                d3 = -l_t * gamma;
            } else if (grad[i] > FLT_EPSILON) {
                const float fi = -rho / grad[i];
                d1 = fi * I1wx[i];
                d2 = fi * I1wy[i];
                d3 = fi * gamma;
            }
            v1[i] = u1[i] + d1;
            v2[i] = u2[i] + d2;
            v3[i] = u3[i] + d3;
        }
    }
}

static double estimate_u_g(const float* v1, const float* v2, const float* v3, const float* div1,
                           const float* div2, const float* div3, float* u1, float* u2, float* u3,
                           int w, int h, float theta, int mode, double* rowsum)
{
    if (mode == 1) {
        float error = 0.0f;
        for (int y = 0; y < h; y++) {
            const size_t o = (size_t)y * w;
            for (int x = 0; x < w; x++) {
                const size_t i = o + x;
                const float u1k = u1[i], u2k = u2[i], u3k = u3[i];
                u1[i] = v1[i] + theta * div1[i];
                u2[i] = v2[i] + theta * div2[i];
                u3[i] = v3[i] + theta * div3[i];
                error += (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k) +
                         (u3[i] - u3k) * (u3[i] - u3k);
            }
        }
        return (double)error;
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const size_t o = (size_t)y * w;
        double acc = 0.0;
        for (int x = 0; x < w; x++) {
            const size_t i = o + x;
            const float u1k = u1[i], u2k = u2[i], u3k = u3[i];
            u1[i] = v1[i] + theta * div1[i];
            u2[i] = v2[i] + theta * div2[i];
            u3[i] = v3[i] + theta * div3[i];
            const float term = (u1[i] - u1k) * (u1[i] - u1k) + (u2[i] - u2k) * (u2[i] - u2k) +
                               (u3[i] - u3k) * (u3[i] - u3k);
            acc += (double)term;
        }
        rowsum[y] = acc;
    }
    double error = 0.0;
    for (int y = 0; y < h; y++) error += rowsum[y];
    return error;
}

/* one channel of estimateDualVariables (the gamma = 0 form above does two at once) */
static void estimate_dual1(const float* ux, const float* uy, float* p1, float* p2, int w, int h,
                           float taut)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const size_t o = (size_t)y * w;
        for (int x = 0; x < w; x++) {
            const size_t i = o + x;
            const float g = hypot_f(ux[i], uy[i]);
            const float ng = 1.0f + taut * g;
            p1[i] = (p1[i] + taut * ux[i]) / ng;
            p2[i] = (p2[i] + taut * uy[i]) / ng;
        }
    }
}

static double iterate_ws_g(iter_ws* s, const float* I1wx, const float* I1wy, const float* grad,
                           const float* rho_c, float* u1, float* u2, float* u3, float* p11,
                           float* p12, float* p21, float* p22, float* p31, float* p32, int w, int h,
                           float l_t, float theta, float taut, float gamma, int mode)
{
    const size_t n = (size_t)w * h;
    if (!s->v3) {
        s->v3 = (float*)malloc(s->cap_px * 4); s->div3 = (float*)malloc(s->cap_px * 4);
        s->u3x = (float*)malloc(s->cap_px * 4); s->u3y = (float*)malloc(s->cap_px * 4);
    }
    (void)n;
    estimate_v_g(I1wx, I1wy, u1, u2, u3, grad, rho_c, s->v1, s->v2, s->v3, w, h, l_t, gamma);
    divergence(p11, p12, s->div1, w, h);
    divergence(p21, p22, s->div2, w, h);
    divergence(p31, p32, s->div3, w, h);
    const double err = estimate_u_g(s->v1, s->v2, s->v3, s->div1, s->div2, s->div3, u1, u2, u3, w, h,
                                    theta, mode, s->rowsum);
    forward_gradient(u1, s->u1x, s->u1y, w, h);
    forward_gradient(u2, s->u2x, s->u2y, w, h);
    forward_gradient(u3, s->u3x, s->u3y, w, h);
    estimate_dual(s->u1x, s->u1y, s->u2x, s->u2y, p11, p12, p21, p22, w, h, taut);
    estimate_dual1(s->u3x, s->u3y, p31, p32, w, h, taut);
    return err;
}

double orc_iterate_gamma(const float* I1wx, const float* I1wy, const float* grad, const float* rho_c,
                         float* u1, float* u2, float* u3, float* p11, float* p12, float* p21,
                         float* p22, float* p31, float* p32, int w, int h, float l_t, float theta,
                         float taut, float gamma, int error_sum_mode)
{
    iter_ws s;
    memset(&s, 0, sizeof(s));
    if (ws_reserve(&s, w, h)) return -1.0;
    const double e = iterate_ws_g(&s, I1wx, I1wy, grad, rho_c, u1, u2, u3, p11, p12, p21, p22, p31,
                                  p32, w, h, l_t, theta, taut, gamma, error_sum_mode);
    ws_free(&s);
    return e;
}

static double iterate_ws(iter_ws* s, const float* I1wx, const float* I1wy, const float* grad,
                         const float* rho_c, float* u1, float* u2, float* p11, float* p12,
                         float* p21, float* p22, int w, int h, float l_t, float theta,
                         float taut, int mode)
{
    estimate_v(I1wx, I1wy, u1, u2, grad, rho_c, s->v1, s->v2, w, h, l_t);
    divergence(p11, p12, s->div1, w, h);
    divergence(p21, p22, s->div2, w, h);
    const double err = estimate_u(s->v1, s->v2, s->div1, s->div2, u1, u2, w, h, theta, mode,
                                  s->rowsum);
    forward_gradient(u1, s->u1x, s->u1y, w, h);
    forward_gradient(u2, s->u2x, s->u2y, w, h);
    estimate_dual(s->u1x, s->u1y, s->u2x, s->u2y, p11, p12, p21, p22, w, h, taut);
    return err;
}

double orc_iterate(const float* I1wx, const float* I1wy, const float* grad, const float* rho_c,
                   float* u1, float* u2, float* p11, float* p12, float* p21, float* p22,
                   int w, int h, float l_t, float theta, float taut, int error_sum_mode)
{
    iter_ws s;
    memset(&s, 0, sizeof(s));
    if (ws_reserve(&s, w, h)) return -1.0;
    const double e = iterate_ws(&s, I1wx, I1wy, grad, rho_c, u1, u2, p11, p12, p21, p22, w, h,
                                l_t, theta, taut, error_sum_mode);
    ws_free(&s);
    return e;
}

/* ------------------------------------------------------------ A.7 5x5 median */

#define CSWAP(i, j) { const float lo = v[i] < v[j] ? v[i] : v[j]; \
                      const float hi = v[i] < v[j] ? v[j] : v[i]; v[i] = lo; v[j] = hi; }

/* 99-exchange median-of-25 selection network (exhaustively verified by
 * tests/test_oracle_primitives.py through orc_median25_selftest) */
static inline float median25(float* v)
{
    CSWAP(0, 1) CSWAP(3, 4) CSWAP(2, 4) CSWAP(2, 3) CSWAP(6, 7) CSWAP(5, 7) CSWAP(5, 6)
    CSWAP(9, 10) CSWAP(8, 10) CSWAP(8, 9) CSWAP(12, 13) CSWAP(11, 13) CSWAP(11, 12)
    CSWAP(15, 16) CSWAP(14, 16) CSWAP(14, 15) CSWAP(18, 19) CSWAP(17, 19) CSWAP(17, 18)
    CSWAP(21, 22) CSWAP(20, 22) CSWAP(20, 21) CSWAP(23, 24) CSWAP(2, 5) CSWAP(3, 6)
    CSWAP(0, 6) CSWAP(0, 3) CSWAP(4, 7) CSWAP(1, 7) CSWAP(1, 4) CSWAP(11, 14) CSWAP(8, 14)
    CSWAP(8, 11) CSWAP(12, 15) CSWAP(9, 15) CSWAP(9, 12) CSWAP(13, 16) CSWAP(10, 16)
    CSWAP(10, 13) CSWAP(20, 23) CSWAP(17, 23) CSWAP(17, 20) CSWAP(21, 24) CSWAP(18, 24)
    CSWAP(18, 21) CSWAP(19, 22) CSWAP(8, 17) CSWAP(9, 18) CSWAP(0, 18) CSWAP(0, 9)
    CSWAP(10, 19) CSWAP(1, 19) CSWAP(1, 10) CSWAP(11, 20) CSWAP(2, 20) CSWAP(2, 11)
    CSWAP(12, 21) CSWAP(3, 21) CSWAP(3, 12) CSWAP(13, 22) CSWAP(4, 22) CSWAP(4, 13)
    CSWAP(14, 23) CSWAP(5, 23) CSWAP(5, 14) CSWAP(15, 24) CSWAP(6, 24) CSWAP(6, 15)
    CSWAP(7, 16) CSWAP(7, 19) CSWAP(13, 21) CSWAP(15, 23) CSWAP(7, 13) CSWAP(7, 15)
    CSWAP(1, 9) CSWAP(3, 11) CSWAP(5, 17) CSWAP(11, 17) CSWAP(9, 17) CSWAP(4, 10)
    CSWAP(6, 12) CSWAP(7, 14) CSWAP(4, 6) CSWAP(4, 7) CSWAP(12, 14) CSWAP(10, 14)
    CSWAP(6, 7) CSWAP(10, 12) CSWAP(6, 10) CSWAP(6, 17) CSWAP(12, 17) CSWAP(7, 17)
    CSWAP(7, 10) CSWAP(12, 18) CSWAP(7, 12) CSWAP(10, 18) CSWAP(12, 20) CSWAP(10, 20)
    CSWAP(10, 12)
    return v[12];
}

/* exhaustive 0/1-principle check of the network: returns number of failing inputs */
long orc_median25_selftest(void)
{
    long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long m = 0; m < (1L << 25); m++) {
        float v[25];
        int ones = 0;
        for (int i = 0; i < 25; i++) {
            v[i] = (float)((m >> i) & 1);
            ones += (int)((m >> i) & 1);
        }
        const float med = median25(v);
        const float want = ones >= 13 ? 1.f : 0.f;
        if (med != want) bad++;
    }
    return bad;
}

void orc_median5(const float* src_in, int w, int h, float* dst)
{
    const float* src = src_in;
    float* copy = NULL;
    if (src_in == dst) {
        copy = (float*)malloc(sizeof(float) * (size_t)w * h);
        memcpy(copy, src_in, sizeof(float) * (size_t)w * h);
        src = copy;
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const float* rows[5];
        for (int k = 0; k < 5; k++) {
            int yy = y + k - 2;
            if (yy < 0) yy = 0;
            if (yy > h - 1) yy = h - 1;
            rows[k] = src + (size_t)yy * w;
        }
        for (int x = 0; x < w; x++) {
            float v[25];
            for (int k = 0; k < 5; k++)
                for (int j = 0; j < 5; j++) {
                    int xx = x + j - 2;
                    if (xx < 0) xx = 0;
                    if (xx > w - 1) xx = w - 1;
                    v[k * 5 + j] = rows[k][xx];
                }
            dst[(size_t)y * w + x] = median25(v);
        }
    }
    free(copy);
}

/* medianBlur(src, 3) on fp32: exact median of the 3x3 window, replicate border (the only other
 * aperture cv::medianBlur accepts for CV_32F).  Selection by a plain insertion sort. */
void orc_median3(const float* src_in, int w, int h, float* dst)
{
    const float* src = src_in;
    float* copy = NULL;
    if (src_in == dst) {
        copy = (float*)malloc(sizeof(float) * (size_t)w * h);
        memcpy(copy, src_in, sizeof(float) * (size_t)w * h);
        src = copy;
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float v[9];
            int n = 0;
            for (int k = -1; k <= 1; k++)
                for (int j = -1; j <= 1; j++) {
                    int yy = y + k, xx = x + j;
                    if (yy < 0) yy = 0;
                    if (yy > h - 1) yy = h - 1;
                    if (xx < 0) xx = 0;
                    if (xx > w - 1) xx = w - 1;
                    const float t = src[(size_t)yy * w + xx];
                    int i = n++;
                    while (i > 0 && v[i - 1] > t) { v[i] = v[i - 1]; i--; }
                    v[i] = t;
                }
            dst[(size_t)y * w + x] = v[4];
        }
    free(copy);
}

/* ---------------------------------------------------------- A.2/A.6/A.8 solve */

int orc_pyramid_sizes(int w, int h, int nscales, double scale_step, int* ws, int* hs)
{
    if (nscales > ORC_MAX_LEVELS) nscales = ORC_MAX_LEVELS;
    ws[0] = w; hs[0] = h;
    int used = nscales;
    for (int s = 1; s < nscales; s++) {
        ws[s] = orc_scaled_size(ws[s - 1], scale_step);
        hs[s] = orc_scaled_size(hs[s - 1], scale_step);
        if (ws[s] < 16 || hs[s] < 16) { used = s; break; }
    }
    return used;
}

int orc_tvl1_calc(const orc_params* p, const unsigned char* I0, long pitch0,
                  const unsigned char* I1, long pitch1, int w, int h,
                  float* u_out, float* v_out, int* iters_out)
{
    if (!p || p->nscales <= 0 || p->nscales > ORC_MAX_LEVELS || w <= 0 || h <= 0) return -1;
    if (p->median_filtering != 1 && p->median_filtering != 3 && p->median_filtering != 5) return -3;
#ifdef _OPENMP
    if (p->nthreads > 0) omp_set_num_threads(p->nthreads);
#endif
    cubic_tab_init();

    int ws[ORC_MAX_LEVELS + 1], hs[ORC_MAX_LEVELS + 1];
    float *I0s[ORC_MAX_LEVELS], *I1s[ORC_MAX_LEVELS], *u1s[ORC_MAX_LEVELS], *u2s[ORC_MAX_LEVELS];
    memset(I0s, 0, sizeof(I0s)); memset(I1s, 0, sizeof(I1s));
    memset(u1s, 0, sizeof(u1s)); memset(u2s, 0, sizeof(u2s));
    if (iters_out)
        for (int i = 0; i < p->nscales * p->warps; i++) iters_out[i] = -1;

    const size_t n0 = (size_t)w * h;
    int nscales = p->nscales;
    ws[0] = w; hs[0] = h;
    I0s[0] = (float*)malloc(n0 * 4); I1s[0] = (float*)malloc(n0 * 4);
    orc_convert_u8(I0, pitch0, w, h, I0s[0]);
    orc_convert_u8(I1, pitch1, w, h, I1s[0]);
    /* create the scales; the level that falls below 16 px is built, then dropped */
    for (int s = 1; s < nscales; s++) {
        ws[s] = orc_scaled_size(ws[s - 1], p->scale_step);
        hs[s] = orc_scaled_size(hs[s - 1], p->scale_step);
        if (ws[s] < 1 || hs[s] < 1) { nscales = s; break; }
        const size_t n = (size_t)ws[s] * hs[s];
        I0s[s] = (float*)malloc(n * 4); I1s[s] = (float*)malloc(n * 4);
        orc_resize_linear(I0s[s - 1], ws[s - 1], hs[s - 1], I0s[s], ws[s], hs[s], p->scale_step);
        orc_resize_linear(I1s[s - 1], ws[s - 1], hs[s - 1], I1s[s], ws[s], hs[s], p->scale_step);
        if (ws[s] < 16 || hs[s] < 16) { nscales = s; break; }
    }
    for (int s = 0; s < nscales; s++) {
        const size_t n = (size_t)ws[s] * hs[s];
        u1s[s] = (float*)malloc(n * 4); u2s[s] = (float*)malloc(n * 4);
    }
    memset(u1s[nscales - 1], 0, (size_t)ws[nscales - 1] * hs[nscales - 1] * 4);
    memset(u2s[nscales - 1], 0, (size_t)ws[nscales - 1] * hs[nscales - 1] * 4);

    float* I1x = (float*)malloc(n0 * 4); float* I1y = (float*)malloc(n0 * 4);
    float* I1wx = (float*)malloc(n0 * 4); float* I1wy = (float*)malloc(n0 * 4);
    float* grad = (float*)malloc(n0 * 4); float* rho_c = (float*)malloc(n0 * 4);
    float* p11 = (float*)malloc(n0 * 4); float* p12 = (float*)malloc(n0 * 4);
    float* p21 = (float*)malloc(n0 * 4); float* p22 = (float*)malloc(n0 * 4);
    float* med = (float*)malloc(n0 * 4);
    /* gamma != 0: u3 per level (zero at the coarsest, resized -- not scaled -- between levels), p31, p32 */
    const int use_gamma = p->gamma != 0.0;
    const float gamma = (float)p->gamma;
    float* u3s[ORC_MAX_LEVELS];
    memset(u3s, 0, sizeof(u3s));
    float *p31 = NULL, *p32 = NULL;
    if (use_gamma) {
        for (int s = 0; s < nscales; s++) u3s[s] = (float*)calloc((size_t)ws[s] * hs[s], 4);
        p31 = (float*)malloc(n0 * 4); p32 = (float*)malloc(n0 * 4);
    }
    iter_ws wsb;
    memset(&wsb, 0, sizeof(wsb));
    ws_reserve(&wsb, w, h);

    const float l_t = (float)(p->lambda * p->theta);
    const float taut = (float)(p->tau / p->theta);
    const float theta = (float)p->theta;

    for (int s = nscales - 1; s >= 0; --s) {
        const int lw = ws[s], lh = hs[s];
        const size_t n = (size_t)lw * lh;
        float *u1 = u1s[s], *u2 = u2s[s];
        /* procOneScale */
        const float scaledEpsilon = (float)(p->epsilon * p->epsilon * (double)(lw * lh));
        orc_centered_gradient(I1s[s], lw, lh, I1x, I1y);
        memset(p11, 0, n * 4); memset(p12, 0, n * 4);
        memset(p21, 0, n * 4); memset(p22, 0, n * 4);
        if (use_gamma) { memset(p31, 0, n * 4); memset(p32, 0, n * 4); }
        for (int warpings = 0; warpings < p->warps; ++warpings) {
            orc_warp(I0s[s], I1s[s], I1x, I1y, u1, u2, lw, lh, NULL, I1wx, I1wy, grad, rho_c);
            float error = FLT_MAX;
            int count = 0;
            for (int n_outer = 0; error > scaledEpsilon && n_outer < p->outer_iterations;
                 ++n_outer) {
                if (p->median_filtering == 5) {
                    orc_median5(u1, lw, lh, med); memcpy(u1, med, n * 4);
                    orc_median5(u2, lw, lh, med); memcpy(u2, med, n * 4);
                } else if (p->median_filtering == 3) {
                    orc_median3(u1, lw, lh, med); memcpy(u1, med, n * 4);
                    orc_median3(u2, lw, lh, med); memcpy(u2, med, n * 4);
                }
                for (int n_inner = 0; error > scaledEpsilon && n_inner < p->inner_iterations;
                     ++n_inner) {
                    const double e = use_gamma
                        ? iterate_ws_g(&wsb, I1wx, I1wy, grad, rho_c, u1, u2, u3s[s], p11, p12, p21, p22,
                                       p31, p32, lw, lh, l_t, theta, taut, gamma, p->error_sum_mode)
                        : iterate_ws(&wsb, I1wx, I1wy, grad, rho_c, u1, u2, p11, p12,
                                     p21, p22, lw, lh, l_t, theta, taut,
                                     p->error_sum_mode);
                    /* the reference holds the error in a float; the canonical fp64 sum is
                     * rounded to fp32 once here so that the comparison is float > float */
                    error = (float)e;
                    ++count;
                }
            }
            if (iters_out) iters_out[s * p->warps + warpings] = count;
        }
        if (s == 0) break;
        /* zoom the flow to the next finer level and scale it (u1, u2 only) */
        const float up = (float)(1 / p->scale_step);
        orc_resize_linear(u1, lw, lh, u1s[s - 1], ws[s - 1], hs[s - 1], 0.0);
        orc_resize_linear(u2, lw, lh, u2s[s - 1], ws[s - 1], hs[s - 1], 0.0);
        if (use_gamma) orc_resize_linear(u3s[s], lw, lh, u3s[s - 1], ws[s - 1], hs[s - 1], 0.0);
        const size_t nn = (size_t)ws[s - 1] * hs[s - 1];
        float *a = u1s[s - 1], *b = u2s[s - 1];
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)nn; i++) { a[i] = a[i] * up; b[i] = b[i] * up; }
    }
    memcpy(u_out, u1s[0], n0 * 4);
    memcpy(v_out, u2s[0], n0 * 4);

    for (int s = 0; s < ORC_MAX_LEVELS; s++) { free(I0s[s]); free(I1s[s]); free(u1s[s]); free(u2s[s]); }
    free(I1x); free(I1y); free(I1wx); free(I1wy); free(grad); free(rho_c);
    free(p11); free(p12); free(p21); free(p22); free(med);
    for (int s = 0; s < ORC_MAX_LEVELS; s++) free(u3s[s]);
    free(p31); free(p32);
    ws_free(&wsb);
    return nscales;
}

/* ------------------------------------------------------- wrapper: 8-bit prescale */

/* cv::resize(frame, frame, cv::Size(), scale, scale) on the decoded 8-bit frame (reference
 * src/optflow.cpp:111,124; default INTER_LINEAR).  OpenCV's 8-bit bilinear path works in fixed
 * point: 11-bit coefficients cvRound((1-f)*2048), cvRound(f*2048), horizontal pass in int, vertical
 * pass ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16), then (+2)>>2.  The column fraction is clamped at
 * the image border, the row fraction is not (only the row index is).  A decimation by exactly 2
 * takes the 2x2 area path instead: (a+b+c+d+2)>>2, and where the source ends early the mean of the
 * pixels that exist, cvRound((float)sum/count).  Pinned bit-exactly against cv2.resize
 * (tests/golden/prescale.npz). */
void orc_prescale_u8(const unsigned char* src, long spitch, int w, int h, double scale,
                     unsigned char* dst, long dpitch, int dw, int dh)
{
    const double inv = 1.0 / scale;
    if (inv == 2.0) {
        const int dw1 = w / 2;
        for (int dy = 0; dy < dh; dy++) {
            const int sy0 = dy * 2;
            const int wfast = sy0 + 2 <= h ? dw1 : 0;
            for (int dx = 0; dx < dw; dx++) {
                const int sx0 = dx * 2;
                unsigned char o = 0;
                if (dx < wfast) {
                    const unsigned char* a = src + (size_t)sy0 * spitch + sx0;
                    o = (unsigned char)((a[0] + a[1] + a[spitch] + a[spitch + 1] + 2) >> 2);
                } else if (sx0 < w && sy0 < h) {
                    int sum = 0, count = 0;
                    for (int yy = 0; yy < 2 && sy0 + yy < h; yy++)
                        for (int xx = 0; xx < 2 && sx0 + xx < w; xx++) {
                            sum += src[(size_t)(sy0 + yy) * spitch + sx0 + xx];
                            count++;
                        }
                    o = (unsigned char)cv_round_f((float)sum / (float)count);
                }
                dst[(size_t)dy * dpitch + dx] = o;
            }
        }
        return;
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)(((double)dy + 0.5) * inv - 0.5);
        const int sy = (int)floorf(fy);
        fy -= (float)sy;
        const int b0 = cv_round_f((1.f - fy) * 2048.f), b1 = cv_round_f(fy * 2048.f);
        const int r0 = sy < 0 ? 0 : (sy > h - 1 ? h - 1 : sy);
        const int r1 = sy + 1 < 0 ? 0 : (sy + 1 > h - 1 ? h - 1 : sy + 1);
        const unsigned char* S0 = src + (size_t)r0 * spitch;
        const unsigned char* S1 = src + (size_t)r1 * spitch;
        for (int dx = 0; dx < dw; dx++) {
            float fx = (float)(((double)dx + 0.5) * inv - 0.5);
            int sx = (int)floorf(fx);
            fx -= (float)sx;
            if (sx < 0) { fx = 0.f; sx = 0; }
            if (sx >= w - 1) { fx = 0.f; sx = w - 1; }
            const int a0 = cv_round_f((1.f - fx) * 2048.f), a1 = cv_round_f(fx * 2048.f);
            const int sx1 = sx + 1 > w - 1 ? w - 1 : sx + 1;
            const int h0 = S0[sx] * a0 + S0[sx1] * a1;
            const int h1 = S1[sx] * a0 + S1[sx1] * a1;
            int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)dy * dpitch + dx] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

/* ------------------------------------------------------- wrapper: mask, sample */

void orc_mask_flow(const unsigned char* f1, long pitch1, int w, int h, float* u, float* v)
{
    /* threshold(frame1, mask, 1, 1, THRESH_BINARY_INV) -> mask = frame1 <= 1;
     * flow.setTo(0, mask)   (reference src/optflow.cpp:471-473) */
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (f1[(size_t)y * pitch1 + x] <= 1) {
                u[(size_t)y * w + x] = 0.f;
                v[(size_t)y * w + x] = 0.f;
            }
}

static inline void point_pq(const float* u, const float* v, int w, int x, int y,
                            int roi0x, int roi0y, int roi1x, int roi1y, float inv_scale,
                            double* px, double* py, double* qx, double* qy)
{
    /* reference src/optflow.cpp:552-556 (features == false branch): int + int -> int,
     * int * float -> float; int + int + float -> float; all widened to double by jsoncpp */
    const size_t i = (size_t)y * w + x;
    *px = (double)((float)(x + roi0x) * inv_scale);
    *py = (double)((float)(y + roi0y) * inv_scale);
    *qx = (double)(((float)(x + roi1x) + u[i]) * inv_scale);
    *qy = (double)(((float)(y + roi1y) + v[i]) * inv_scale);
}

void orc_points_at(const float* u, const float* v, int w, int n, const int* positions,
                   int roi0x, int roi0y, int roi1x, int roi1y, float scale,
                   double* px, double* py, double* qx, double* qy)
{
    const float inv_scale = (float)(1. / scale);   /* float inv_scale = 1./scale (:528) */
    for (int k = 0; k < n; k++)
        point_pq(u, v, w, positions[2 * k], positions[2 * k + 1], roi0x, roi0y, roi1x, roi1y,
                 inv_scale, &px[k], &py[k], &qx[k], &qy[k]);
}

int orc_random_points(const unsigned char* f0, long pitch0, const unsigned char* f1, long pitch1,
                      const float* u, const float* v, int w, int h,
                      int roi0x, int roi0y, int roi1x, int roi1y, float scale,
                      int npoints, long seed,
                      double* px, double* py, double* qx, double* qy, double* wgt,
                      int* positions)
{
    const float inv_scale = (float)(1. / scale);
    /* mask = threshold(f0,1,1,BINARY) | threshold(f1,1,1,BINARY); findNonZero (:488-493,:531) */
    size_t count = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (f0[(size_t)y * pitch0 + x] > 1 || f1[(size_t)y * pitch1 + x] > 1) count++;
    if (count == 0) {
        px[0] = py[0] = qx[0] = qy[0] = -1.0;
        wgt[0] = 0.0;
        return 1;
    }
    int* loc = (int*)malloc(sizeof(int) * count);   /* linear index y*w + x */
    size_t k = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            if (f0[(size_t)y * pitch0 + x] > 1 || f1[(size_t)y * pitch1 + x] > 1)
                loc[k++] = y * w + x;
    if (seed >= 0) srand((unsigned)seed);
    /* libstdc++ std::random_shuffle(first, last): for i in 1..n-1 swap(v[i], v[rand() % (i+1)]) */
    for (size_t i = 1; i < count; i++) {
        const size_t j = (size_t)rand() % (i + 1);
        const int t = loc[i]; loc[i] = loc[j]; loc[j] = t;
    }
    int n = 0;
    for (; n < npoints && (size_t)n < count; n++) {
        const int x = loc[n] % w, y = loc[n] / w;
        if (positions) { positions[2 * n] = x; positions[2 * n + 1] = y; }
        point_pq(u, v, w, x, y, roi0x, roi0y, roi1x, roi1y, inv_scale,
                 &px[n], &py[n], &qx[n], &qy[n]);
        wgt[n] = 1.0;
    }
    free(loc);
    return n;
}
