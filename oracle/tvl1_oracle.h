/*
 * tvl1_oracle.h -- CPU restatement of the TV-L1 flow stage of fibsem-optflow.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library, and only as the checker or as
 * the timed CPU baseline.  The product (fibsem_optflow_b200/csrc) never links,
 * includes or calls it.
 *
 * PARITY UNPINNED (composition): the reference's flow stage is one call into
 * OpenCV 3.4.1's DualTVL1 (reference src/optflow.cpp:516-520), a third-party
 * dependency pinned only by a download URL (reference
 * singularity/optflow.def:22-23) that is neither vendored nor installed here,
 * and the reference ships no tests or golden vectors.  This file restates
 * the published algorithm of cv::DualTVL1OpticalFlow
 * (modules/video/src/tvl1flow.cpp, OpenCV 3.4.1) from SURVEY.md Appendix A.
 * What IS pinned: the three non-trivial primitives (bilinear resize, cubic
 * remap, 5x5 median) are checked against the installed cv2 4.13 and against
 * committed golden vectors (tests/golden/), and the sampling path against the
 * glibc rand()/std::random_shuffle known answers of SURVEY.md C6.
 */
#ifndef TVL1_ORACLE_H
#define TVL1_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_params {
    double tau;              /* 0.25 */
    double lambda;           /* OpenCV 0.15; reference wrapper 0.05 (src/optflow.cpp:504) */
    double theta;            /* 0.3 */
    double epsilon;          /* 0.01 */
    double scale_step;       /* 0.8 */
    double gamma;            /* 0; != 0: the third channel u3 / p31, p32 (as recalled) */
    int nscales;             /* OpenCV 5; reference wrapper 10 (src/optflow.cpp:506) */
    int warps;               /* 5 */
    int inner_iterations;    /* 30 */
    int outer_iterations;    /* 10 */
    int median_filtering;    /* 5 (1 = off; 3 and 5 are the apertures cv::medianBlur has for fp32) */
    int error_sum_mode;      /* 0 = fp32 terms summed in fp64 (canonical)
                                1 = one serial fp32 scalar, row-major (OpenCV literal) */
    int nthreads;            /* OpenMP threads, 0 = default */
} orc_params;

void orc_default_params(orc_params* p);

/* A.2: dst size of resize(src, Size(), f, f): saturate_cast<int>(n*f), round-half-even */
int orc_scaled_size(int n, double f);

/* A.2: bilinear resize.  inv_scale > 0: resize(src, Size(), inv_scale, inv_scale)
 * (dw, dh must equal orc_scaled_size); inv_scale <= 0: resize(src, Size(dw, dh)). */
void orc_resize_linear(const float* src, int sw, int sh, float* dst, int dw, int dh,
                       double inv_scale);

void orc_convert_u8(const unsigned char* src, long pitch, int w, int h, float* dst);

/* A.3 */
void orc_centered_gradient(const float* src, int w, int h, float* dx, float* dy);

/* A.4: one plane of remap(src, map1, map2, INTER_CUBIC, BORDER_CONSTANT 0) */
void orc_remap_cubic(const float* src, int w, int h, const float* mapx, const float* mapy,
                     float* dst);

/* A.4: buildFlowMap + 3 remaps + calcGradRho.  I1w may be NULL. */
void orc_warp(const float* I0, const float* I1, const float* I1x, const float* I1y,
              const float* u1, const float* u2, int w, int h,
              float* I1w, float* I1wx, float* I1wy, float* grad, float* rho_c);

/* A.5: one inner iteration, in place on u1,u2,p11..p22.  Returns the error sum
 * (as double; in mode 1 it is the fp32 scalar widened). */
double orc_iterate(const float* I1wx, const float* I1wy, const float* grad, const float* rho_c,
                   float* u1, float* u2, float* p11, float* p12, float* p21, float* p22,
                   int w, int h, float l_t, float theta, float taut, int error_sum_mode);

/* A.5 with gamma != 0: the same iteration with the third channel (u3, p31, p32). */
double orc_iterate_gamma(const float* I1wx, const float* I1wy, const float* grad, const float* rho_c,
                         float* u1, float* u2, float* u3, float* p11, float* p12, float* p21,
                         float* p22, float* p31, float* p32, int w, int h, float l_t, float theta,
                         float taut, float gamma, int error_sum_mode);

/* A.7: exact 5x5 median, replicate border; dst may equal src. */
void orc_median5(const float* src, int w, int h, float* dst);

/* medianBlur(src, 3): exact 3x3 median, replicate border; dst may equal src. */
void orc_median3(const float* src, int w, int h, float* dst);

/* number of pyramid levels actually used (A.2 stop rule) and their sizes */
int orc_pyramid_sizes(int w, int h, int nscales, double scale_step, int* ws, int* hs);

/* Whole solve (A.2-A.8).  I0/I1: 8-bit rows with byte pitch.  u, v: w*h floats.
 * iters_out (may be NULL): nscales*warps ints, [s*warps + w] = inner iterations
 * executed at level s (0 = finest), warp w; levels not used are -1.
 * Returns levels used, or < 0 on error. */
int orc_tvl1_calc(const orc_params* p, const unsigned char* I0, long pitch0,
                  const unsigned char* I1, long pitch1, int w, int h,
                  float* u, float* v, int* iters_out);

/* reference src/optflow.cpp:111,124: cv::resize(frame, frame, Size(), scale, scale) on the 8-bit frame.
 * dw, dh = orc_scaled_size(w, scale), orc_scaled_size(h, scale). */
void orc_prescale_u8(const unsigned char* src, long spitch, int w, int h, double scale,
                     unsigned char* dst, long dpitch, int dw, int dh);

/* reference src/optflow.cpp:471-473: flow = 0 where frame1 <= 1 */
void orc_mask_flow(const unsigned char* f1, long pitch1, int w, int h, float* u, float* v);

/* reference src/optflow.cpp:488-493 + 522-572 (random_points, features == false).
 * mask = (f0 > 1) | (f1 > 1); findNonZero row-major; std::random_shuffle driven by
 * glibc rand(): seed < 0 -> no srand (the reference's debug mode), else srand(seed).
 * Writes up to npoints entries: px,py,qx,qy (fp32 arithmetic, stored as the doubles
 * jsoncpp would hold) and w.  If the mask is empty writes the dummy (-1,-1,-1,-1,0).
 * positions (may be NULL) receives the sampled (x,y) pairs.  Returns entries written. */
int orc_random_points(const unsigned char* f0, long pitch0, const unsigned char* f1, long pitch1,
                      const float* u, const float* v, int w, int h,
                      int roi0x, int roi0y, int roi1x, int roi1y, float scale,
                      int npoints, long seed,
                      double* px, double* py, double* qx, double* qy, double* wgt,
                      int* positions);

/* fp32 p/q arithmetic only, for caller-given positions (T5 in SURVEY.md). */
void orc_points_at(const float* u, const float* v, int w, int n, const int* positions,
                   int roi0x, int roi0y, int roi1x, int roi1y, float scale,
                   double* px, double* py, double* qx, double* qy);

#ifdef __cplusplus
}
#endif
#endif
