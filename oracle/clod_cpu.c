/*
 * clod_cpu.c -- "CLOD-CPU": CPU restatement of the reference's OWN detector, clodDetectObjects with
 * use_cl = FALSE (clod.cpp:1339-1500), the path BASELINE.json calls "the reference's CPU path"
 * (what main.cpp:79,90 run with CLOD_PER_STAGE_ITERATIONS | CLOD_PRECOMPUTE_FEATURES).
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see vj_oracle.h): bench.py times it beside the GPU number.
 * It is NOT the parity oracle of the CUDA path: this detector has other semantics than REF-SI
 * (tempcv.cpp) -- float arithmetic, features scaled by a float factor on one integral image, a
 * window step of max(2, scale), public (un-biased) stage thresholds, stumps only (first feature of
 * every classifier, clod.cpp:458,649), no tilted features, a skip of the NEXT LIST ENTRY after a
 * stage-0 reject (clod.cpp:729-731) -- see SURVEY.md Appendix A.
 *
 * PARITY STATUS: PINNED.  oracle/build_ref.py compiles clod.cpp:11-38, 371-527, 580-787, 1339-1500
 * from /root/reference into oracle/_ref/libclod_ref.so; tests/test_clod_cpu.py holds this file's raw
 * match lists equal to that library's (scale factor 1.1, the value clod.cpp:1349 hard-codes; here it
 * is a parameter), and tests/golden/reference_clod.npz keeps its outputs for machines without it.
 * Not restated: filterResult (clod.cpp:282-357) -- it accumulates into malloc'ed, never initialised
 * rectangles (clod.cpp:290,297-300), so its output is undefined; main.cpp calls with min_neighbors = 0.
 *
 * All file:line citations are relative to /root/reference/CLFaceDetection/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "vj_oracle.h"

#define CLOD_PRECOMPUTE_FEATURES (2 << 0)  /* clod.h:17 */
#define CLOD_PER_STAGE_ITERATIONS (2 << 2) /* clod.h:19 */

typedef struct {
    int win_w, win_h, n_stages;
    const int *st_ntrees;
    const float *st_thr;  /* public thresholds: no 0.0001 bias (that is tempcv.cpp's hidden cascade) */
    const int *nd_rect;   /* [N][3][4] */
    const float *nd_weight; /* [N][3] */
    const float *nd_thr;
    const float *alpha;   /* stumps: 2 per classifier */
} clod_cascade;

typedef struct { /* CLODOptimizedRect, clod.cpp:25-31 (offsets instead of pointers) */
    int64_t lt, rt, lb, rb;
    float weight;
} opt_rect;

typedef struct { /* CLODSubwindowData, clod.cpp:33-38 */
    uint32_t x, y, offset;
    float variance;
} subwindow;

typedef struct {
    float step;
    int sw, sh;       /* scaled window */
    int ex, ey, ew, eh; /* equ_rect */
    uint32_t area;
    int end_x, end_y;
} scale_setup;

/* setupScale, clod.cpp:371-415 */
static int setup_scale(float cs, int W, int H, int w0, int h0, int min_w, int min_h, int max_w, int max_h, scale_setup *s)
{
    s->step = (float)(2.0 > (double)cs ? 2.0 : (double)cs);
    s->sw = (int)(uint32_t)round((double)((float)w0 * cs));
    s->sh = (int)(uint32_t)round((double)((float)h0 * cs));
    if (s->sw < min_w || s->sh < min_h) return -1;
    if (max_w != 0 && s->sw > max_w) return -1;
    if (max_h != 0 && s->sh > max_h) return -1;
    if (s->sw > W || s->sh > H) return -1;
    s->ex = s->ey = (int)(uint32_t)round((double)cs);
    s->ew = (int)(uint32_t)round((double)((float)(w0 - 2) * cs));
    s->eh = (int)(uint32_t)round((double)((float)(h0 - 2) * cs));
    s->area = (uint32_t)(s->ew * s->eh);
    s->end_x = (int)lrint((double)((float)(W - s->sw) / s->step));
    s->end_y = (int)lrint((double)((float)(H - s->sh) / s->step));
    return 0;
}

/* computeVariance, clod.cpp:418-446: float mean, squares through (unsigned long) casts of the doubles */
static float compute_variance(const int32_t *sum, const double *sq, int stride, const scale_setup *s, uint32_t px, uint32_t py)
{
    const size_t x = (size_t)px + (size_t)s->ex, y = (size_t)py + (size_t)s->ey, w = (size_t)s->ew, h = (size_t)s->eh;
#define AT(m, xx, yy) ((m)[(size_t)stride * (yy) + (xx)])
    const int m4 = AT(sum, x, y) - AT(sum, x + w, y) - AT(sum, x, y + h) + AT(sum, x + w, y + h);
    const float mean = (float)m4 / (float)s->area;
    const unsigned long q4 = (unsigned long)AT(sq, x, y) - (unsigned long)AT(sq, x + w, y) - (unsigned long)AT(sq, x, y + h) +
                             (unsigned long)AT(sq, x + w, y + h);
#undef AT
    float variance = (float)q4;
    variance = (variance / (float)s->area) - (mean * mean);
    if (variance >= 0) variance = (float)sqrt((double)variance);
    else variance = 1;
    return variance;
}

/* precomputeFeatures, clod.cpp:448-493 (and the same arithmetic inside runClassifier, 580-634): a packed
 * list of 2 or 3 rects per classifier; returns the number of rects written */
static size_t precompute_features(const clod_cascade *c, int stride, float cs, uint32_t area, opt_rect *opt, int *n_rects_of)
{
    size_t k = 0, node = 0;
    for (int s = 0; s < c->n_stages; s++) {
        for (int t = 0; t < c->st_ntrees[s]; t++, node++) {
            float first_rect_area = 0.f, sum_rect_area = 0.f;
            const size_t first = k;
            for (int i = 0; i < 3; i++) {
                const float fw = c->nd_weight[node * 3 + i];
                if (fw != 0) {
                    const int *r = c->nd_rect + (node * 3 + i) * 4;
                    const uint32_t rx = (uint32_t)round((double)((float)r[0] * cs)), ry = (uint32_t)round((double)((float)r[1] * cs));
                    const uint32_t rw = (uint32_t)round((double)((float)r[2] * cs)), rh = (uint32_t)round((double)((float)r[3] * cs));
                    const float rect_weight = fw / (float)area;
                    opt[k].weight = rect_weight;
                    opt[k].lt = (int64_t)stride * ry + rx;
                    opt[k].rt = (int64_t)stride * ry + rx + rw;
                    opt[k].lb = (int64_t)stride * (ry + rh) + rx;
                    opt[k].rb = (int64_t)stride * (ry + rh) + rx + rw;
                    if (i > 0) sum_rect_area += rect_weight * (float)rw * (float)rh;
                    else first_rect_area = (float)(rw * rh);
                    k++;
                }
            }
            opt[first].weight = (-sum_rect_area / first_rect_area);
            if (n_rects_of) n_rects_of[node] = (int)(k - first);
        }
    }
    return k;
}

/* runClassifierWithPrecomputedFeatures, clod.cpp:636-679: unsigned corner arithmetic, float products and sums */
static inline float rect_term(const uint32_t *isum, const opt_rect *o, size_t off)
{
    const uint32_t v = isum[o->lt + off] - isum[o->rt + off] - isum[o->lb + off] + isum[o->rb + off];
    return (float)v * o->weight;
}

static inline void run_classifier(const clod_cascade *c, size_t node, const uint32_t *isum, const opt_rect *opt, size_t *k, size_t off,
                                  float variance, float *stage_sum)
{
    const float norm_threshold = c->nd_thr[node] * variance;
    float rect_sum = rect_term(isum, &opt[*k], off);
    (*k)++;
    rect_sum += rect_term(isum, &opt[*k], off);
    (*k)++;
    if (c->nd_weight[node * 3 + 2] != 0) {
        rect_sum += rect_term(isum, &opt[*k], off);
        (*k)++;
    }
    *stage_sum += c->alpha[node * 2 + (rect_sum >= norm_threshold)];
}

/* clodDetectObjects(use_cl = FALSE) on one 8-bit gray frame (clod.cpp:1339-1500; setupImage -> clifGrayscaleIntegral's
 * CPU branch, clif.cpp:326-331: cvIntegral into CV_32SC1 / CV_64FC1).  Returns the number of raw matches
 * (min_neighbors = 0), or -1 for a cascade this detector cannot run (trees, more than 220 classifiers in a stage). */
int64_t clodcpu_detect(int win_w, int win_h, int n_stages, const int *st_ntrees, const float *st_thr, const int *tr_nnodes,
                       const int *nd_rect, const float *nd_weight, const float *nd_thr, const float *alpha,
                       const uint8_t *img, int W, int H, int stride_bytes, float scale_factor,
                       int min_w, int min_h, int max_w, int max_h, unsigned flags,
                       int32_t *rects, int64_t cap, int64_t *windows, int64_t *classifier_evals)
{
    clod_cascade c = {win_w, win_h, n_stages, st_ntrees, st_thr, nd_rect, nd_weight, nd_thr, alpha};
    size_t n_nodes = 0;
    for (int s = 0; s < n_stages; s++) {
        if (st_ntrees[s] > 220) return -1;   /* MAX_STAGE_CLASSIFIER_COUNT, clod.cpp:13,1377 */
        for (int t = 0; t < st_ntrees[s]; t++)
            if (tr_nnodes[n_nodes + t] != 1) return -1;
        n_nodes += (size_t)st_ntrees[s];
    }
    const int stride = W + 1;   /* integral_image->width */
    int32_t *sum = (int32_t *)malloc((size_t)stride * (H + 1) * sizeof(int32_t));
    double *sq = (double *)malloc((size_t)stride * (H + 1) * sizeof(double));
    opt_rect *opt = (opt_rect *)malloc(n_nodes * 3 * sizeof(opt_rect));
    if (!sum || !sq || !opt) { free(sum); free(sq); free(opt); return -1; }
    vjo_integral(img, W, H, stride_bytes, sum, sq, NULL);
    const uint32_t *isum = (const uint32_t *)sum;

    /* clod.cpp:1366-1373 */
    unsigned scale_count = 0;
    for (float cs = 1; cs * (float)win_w < (float)(W - 10) && cs * (float)win_h < (float)(H - 10); cs *= scale_factor) scale_count++;

    int64_t n_match = 0, n_windows = 0, n_evals = 0;
    float cs = 1;
    for (unsigned si = 0; si < scale_count; si++, cs *= scale_factor) {
        scale_setup S;
        if (setup_scale(cs, W, H, win_w, win_h, min_w, min_h, max_w, max_h, &S) != 0) continue;
        precompute_features(&c, stride, cs, S.area, opt, NULL);   /* (the non-precomputed variant does the same arithmetic per call) */
        if (!(flags & CLOD_PER_STAGE_ITERATIONS)) {
            /* clod.cpp:1410-1432 + runCascade, 736-787: window at a time, x step 2 after a stage-0 exit */
            for (int yi = 0; yi < S.end_y; yi++) {
                int x_incr = 1;
                for (int xi = 0; xi < S.end_x; xi += x_incr) {
                    const uint32_t px = (uint32_t)round((double)((float)xi * S.step)), py = (uint32_t)round((double)((float)yi * S.step));
                    const float variance = compute_variance(sum, sq, stride, &S, px, py);
                    const size_t off = (size_t)stride * py + px;
                    int exit_stage = 1;
                    size_t k = 0, node = 0;
                    n_windows++;
                    for (int s = 0; s < n_stages; s++) {
                        float stage_sum = 0;
                        for (int t = 0; t < st_ntrees[s]; t++, node++) run_classifier(&c, node, isum, opt, &k, off, variance, &stage_sum);
                        n_evals += st_ntrees[s];
                        if (stage_sum < st_thr[s]) { exit_stage = -s; break; }
                    }
                    if (exit_stage > 0) {
                        if (rects && n_match < cap) {
                            int32_t *r = rects + 4 * n_match;
                            r[0] = (int32_t)px; r[1] = (int32_t)py; r[2] = S.sw; r[3] = S.sh;
                        }
                        n_match++;
                    }
                    x_incr = exit_stage != 0 ? 1 : 2;
                }
            }
        } else {
            /* clod.cpp:1434-1482: precomputeWindows (495-527), then one pass over the survivor list per stage (runSubwindow, 681-734) */
            const size_t n0 = (size_t)(S.end_y > 0 ? S.end_y : 0) * (size_t)(S.end_x > 0 ? S.end_x : 0);
            subwindow *in = (subwindow *)malloc((n0 + 1) * sizeof(subwindow)), *out = NULL;
            size_t n_in = 0, n_out = 0;
            for (int yi = 0; yi < S.end_y; yi++)
                for (int xi = 0; xi < S.end_x; xi++) {
                    const uint32_t px = (uint32_t)lrint((double)((float)xi * S.step)), py = (uint32_t)lrint((double)((float)yi * S.step));
                    in[n_in].x = px; in[n_in].y = py;
                    in[n_in].variance = compute_variance(sum, sq, stride, &S, px, py);
                    in[n_in].offset = (uint32_t)stride * py + px;
                    n_in++;
                }
            size_t k0 = 0, node0 = 0;
            for (int s = 0; s < n_stages; s++) {
                out = (subwindow *)malloc((n_in + 1) * sizeof(subwindow));
                n_out = 0;
                size_t k_end = k0;
                size_t incr = 1;
                for (size_t i = 0; i < n_in; i += incr) {
                    const subwindow w = in[i];
                    float stage_sum = 0;
                    size_t k = k0, node = node0;
                    for (int t = 0; t < st_ntrees[s]; t++, node++) run_classifier(&c, node, isum, opt, &k, w.offset, w.variance, &stage_sum);
                    k_end = k;
                    n_evals += st_ntrees[s];
                    if (s == 0) n_windows++;
                    incr = 1;
                    if (stage_sum >= st_thr[s]) out[n_out++] = w;
                    else if (s == 0) incr = 2;   /* skips the next LIST entry, also across rows (clod.cpp:729-731) */
                }
                free(in);
                in = out; n_in = n_out;
                k0 = k_end; node0 += (size_t)st_ntrees[s];
                if (n_out == 0) break;
            }
            for (size_t i = 0; i < n_in; i++) {
                if (rects && n_match < cap) {
                    int32_t *r = rects + 4 * n_match;
                    r[0] = (int32_t)in[i].x; r[1] = (int32_t)in[i].y; r[2] = S.sw; r[3] = S.sh;
                }
                n_match++;
            }
            free(in);
        }
    }
    free(sum); free(sq); free(opt);
    if (windows) *windows = n_windows;
    if (classifier_evals) *classifier_evals = n_evals;
    return n_match;
}

/* n_frames frames frame_stride bytes apart, one frame per thread at a time (the reference itself is single-threaded:
 * "Parallelize this", clod.cpp:700): match counts per frame; totals of windows / classifier evaluations */
int64_t clodcpu_detect_batch(int win_w, int win_h, int n_stages, const int *st_ntrees, const float *st_thr, const int *tr_nnodes,
                             const int *nd_rect, const float *nd_weight, const float *nd_thr, const float *alpha,
                             const uint8_t *frames, int n_frames, int64_t frame_stride, int W, int H, int stride_bytes,
                             float scale_factor, int min_w, int min_h, int max_w, int max_h, unsigned flags,
                             int64_t *match_counts, int64_t *windows, int64_t *classifier_evals, int n_threads)
{
    int64_t total = 0, tw = 0, te = 0;
    int bad = 0;
    if (n_threads <= 0) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads) reduction(+ : total, tw, te) reduction(| : bad)
    for (int f = 0; f < n_frames; f++) {
        int64_t w = 0, e = 0;
        const int64_t m = clodcpu_detect(win_w, win_h, n_stages, st_ntrees, st_thr, tr_nnodes, nd_rect, nd_weight, nd_thr, alpha,
                                         frames + (size_t)f * (size_t)frame_stride, W, H, stride_bytes, scale_factor, min_w, min_h,
                                         max_w, max_h, flags, NULL, 0, &w, &e);
        if (m < 0) bad |= 1;
        if (match_counts) match_counts[f] = m;
        total += m > 0 ? m : 0; tw += w; te += e;
    }
    if (windows) *windows = tw;
    if (classifier_evals) *classifier_evals = te;
    return bad ? -1 : total;
}
