/*
 * vj_oracle.h -- CPU restatement ("REF-SI") of the reference's Viola-Jones hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under clfacedetection_b200/ may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY STATUS: PINNED to the reference's own code.  oracle/build_ref.py compiles the reference's
 * Haar functions themselves -- tempcv.cpp:40-1516 (AgroupRectangles, hidden-cascade builder,
 * cvSetImagesForHaarClassifierCascade, cvRunHaarClassifierCascadeSum, both invokers,
 * cvHaarDetectObjectsForROC) and 1702-2089 (icvReadHaarClassifier) -- from where they lie under
 * /root/reference into oracle/_ref/libtempcv_ref.so, against a test-only stand-in for the OpenCV
 * 2.4 names they touch (oracle/ref_shim/).  tests/test_oracle_vs_reference.py holds this file equal
 * to that library: per-window return codes, stage sums, raw / grouped / reject-level rect lists
 * of the whole drivers (image-pyramid and scale-cascade), hidden-cascade weights and geometry, the
 * XML reader, on all 19 cascade files; tests/golden/reference_tempcv.npz keeps the library's
 * outputs for machines without it.  What the reference leaves to the OpenCV 2.4.2 dylibs (resize,
 * integral, tilted integral, colour conversion) is pinned bit-for-bit against cv2 4.13 in
 * tests/test_oracle_pins.py and tests/golden/opencv_pins.npz.
 *
 * All file:line citations are relative to /root/reference/CLFaceDetection/.
 */
#ifndef VJ_ORACLE_H
#define VJ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vjo_cascade vjo_cascade;

/* Flat description of a CvHaarClassifierCascade (tempcv.hpp:70-112) as produced by
 * oracle/cascade_xml.py (restating icvReadHaarClassifier, tempcv.cpp:1750-2089).
 *   st_*   : one entry per stage
 *   tr_nnodes : nodes per tree, trees concatenated in stage order
 *   nd_*   : one entry per node, nodes concatenated in tree order
 *   nd_rect: [N][3][4] = x,y,w,h   nd_weight: [N][3]
 *   nd_left/right: >0 = node index in the same tree, <=0 = leaf alpha[-idx]
 *   alpha  : (nnodes+1) floats per tree, concatenated
 * Returns NULL (and sets vjo_last_error) on a malformed cascade. */
vjo_cascade *vjo_cascade_create(int win_w, int win_h, int n_stages,
                                const int *st_ntrees, const float *st_thr,
                                const int *st_parent, const int *st_next,
                                const int *tr_nnodes,
                                const int *nd_tilted, const int *nd_rect,
                                const float *nd_weight, const float *nd_thr,
                                const int *nd_left, const int *nd_right,
                                const float *alpha);
void vjo_cascade_free(vjo_cascade *c);
const char *vjo_last_error(void);

/* bit0 is_tree, bit1 isStumpBased, bit2 has_tilted_features (tempcv.cpp:371,431,465) */
int vjo_cascade_flags(const vjo_cascade *c);
int vjo_cascade_counts(const vjo_cascade *c, int *n_stages, int *n_trees, int *n_nodes);
/* hidden-cascade values at scale 1 (tempcv.cpp:419,453-458,733,752-760) */
void vjo_cascade_hid(const vjo_cascade *c, float *node_weights /*[N][3]*/,
                     int *node_nrects /*[N]*/, float *stage_thr /*[S]*/,
                     int *stage_two_rects /*[S]*/, int *stage_child /*[S]*/);

/* cvResize(INTER_LINEAR) for 8-bit single channel (OpenCV imgproc, external; call
 * site tempcv.cpp:1301).  Returns 0 on success. */
int vjo_resize_linear(const uint8_t *src, int sw, int sh, int sstride,
                      uint8_t *dst, int dw, int dh, int dstride);

/* cvIntegral (external; call site tempcv.cpp:1302).  sum/tilted are int32
 * [(h+1)][(w+1)], sqsum is double (exact integers).  tilted may be NULL. */
void vjo_integral(const uint8_t *img, int w, int h, int stride,
                  int32_t *sum, double *sqsum, int32_t *tilted);

/* Level loop of the CV_HAAR_SCALE_IMAGE path (tempcv.cpp:1230-1234,1268-1288) and the
 * window grid of the invoker (tempcv.cpp:1013-1021,1079-1080).
 * Fills up to max_levels entries; returns the number of levels.  A zero max_w/max_h
 * means "image size" (tempcv.cpp:1230-1234). */
typedef struct vjo_level {
    double factor;
    int img_w, img_h;   /* sz   */
    int win_w, win_h;   /* winSize (output rect size) */
    int ystep;
    int nx, ny;         /* windows per row / rows of windows */
} vjo_level;
int vjo_plan_levels(int W, int H, int w0, int h0, double scale_factor,
                    int min_w, int min_h, int max_w, int max_h,
                    vjo_level *levels, int max_levels);

typedef struct vjo_stats {
    int64_t windows;           /* windows evaluated */
    int64_t weak_evals;        /* weak classifiers (trees) evaluated */
    int64_t node_evals;        /* tree nodes evaluated */
    int64_t accepted;          /* windows accepted */
    int64_t near_stage_thr;    /* windows with |stage_sum-thr| <= 1e-5*|thr| at some stage */
    int64_t stage_reach[64];   /* windows that evaluated stage i (i<64) */
    int64_t near_stage_events; /* (window, stage) pairs with |stage_sum-thr| <= 1e-5*|thr| */
} vjo_stats;

/* Whole REF-SI detection of one 8-bit gray frame.
 *   rects     : out, [cap][4] = x,y,w,h in raster order per level (NULL to skip)
 *   codes     : out, int16 per window, levels concatenated (level l at offset
 *               sum_{k<l} nx_k*ny_k); linear cascades: number of stages passed
 *               (count = accepted); stage-tree cascades: 2*last_stage + accepted.
 *   near      : out, uint8 per window, 1 if some evaluated stage sum was within
 *               1e-5 relative of its threshold.  codes / near may be NULL.
 * Returns the number of accepted windows (may exceed cap; only cap are written),
 * or -1 on error. */
int64_t vjo_detect(const vjo_cascade *c, const uint8_t *img, int W, int H, int stride,
                   double scale_factor, int min_w, int min_h, int max_w, int max_h,
                   int32_t *rects, int64_t cap, int16_t *codes, uint8_t *near,
                   vjo_stats *stats, int n_threads);

/* cvHaarDetectObjectsForROC with outputRejectLevels = true (tempcv.cpp:1084-1094), REF-SI path:
 * accepted windows (level = number of stages) and windows rejected by one of the last three
 * stages (level = that stage's index), level_weights = the stage sum of the last evaluated
 * stage.  reject_levels / level_weights hold `cap` entries like rects. */
int64_t vjo_detect_roc(const vjo_cascade *c, const uint8_t *img, int W, int H, int stride,
                       double scale_factor, int min_w, int min_h, int max_w, int max_h,
                       int32_t *rects, int32_t *reject_levels, double *level_weights, int64_t cap,
                       int16_t *codes, uint8_t *near, vjo_stats *stats, int n_threads);

/* Evaluate every grid window of ONE level whose image is given directly (no resize):
 * used by unit tests of the evaluator. */
int64_t vjo_eval_level(const vjo_cascade *c, const uint8_t *img, int w, int h, int stride,
                       int ystep, int16_t *codes, uint8_t *near, vjo_stats *stats,
                       int n_threads);

/* REF-SC: the scale-cascade path (tempcv.cpp:1330-1456, flags = 0; what main.cpp:145 runs):
 * one integral image, features scaled per factor (tempcv.cpp:549-768), grid and skip rule of
 * HaarDetectObjects_ScaleCascade_Invoker (tempcv.cpp:1132-1175).  vjo_plan_sc fills one
 * vjo_level per evaluated scale (img_w/h = frame size, ystep = 0: the step is max(2, factor),
 * nx/ny = endX/endY).  codes: as vjo_detect, plus VJO_CODE_SKIPPED for grid positions the skip
 * rule never evaluates and VJO_CODE_OUTSIDE for windows the bounds check rejects (result -1). */
#define VJO_CODE_SKIPPED (-32768)
#define VJO_CODE_OUTSIDE (-32767)
int vjo_plan_sc(int W, int H, int w0, int h0, double scale_factor, int min_w, int min_h,
                vjo_level *levels, int max_levels);
int64_t vjo_detect_sc(const vjo_cascade *c, const uint8_t *img, int W, int H, int stride,
                      double scale_factor, int min_w, int min_h,
                      int32_t *rects, int64_t cap, int16_t *codes, vjo_stats *stats, int n_threads);

/* AgroupRectangles(rectList, weights, groupThreshold, eps) (tempcv.cpp:130-243).
 * rects in/out [n][4]; weights out [n]; returns the new count. */
int vjo_group_rectangles(int32_t *rects, int n, int group_threshold, double eps,
                         int32_t *weights);
/* the ROC variant (tempcv.cpp:255-258): reject_levels / level_weights in (per rect) and out (per
 * kept class: the highest level in the class and the largest stage sum seen at that level) */
int vjo_group_rectangles_roc(int32_t *rects, int n, int group_threshold, double eps, int32_t *reject_levels,
                             double *level_weights);

#ifdef __cplusplus
}
#endif
#endif
