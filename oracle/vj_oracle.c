/*
 * vj_oracle.c -- CPU restatement of the reference's Viola-Jones hot path ("REF-SI":
 * the CV_HAAR_SCALE_IMAGE path of tempcv.cpp).  See vj_oracle.h for status and scope.
 *
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline); never on the product path.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: the reference was built for x86-64 without FMA, so every
 * double/float multiply and add below rounds separately, exactly as written there.
 * C expression types are kept as in the reference (e.g. `int * float` is a FLOAT
 * product in the 3-rect/tree paths, tempcv.cpp:782-786,907-910, but a DOUBLE product in
 * the two_rects path, tempcv.cpp:880-885).
 *
 * Citations are file:line under /root/reference/CLFaceDetection/.
 */
#include "vj_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define VJO_FEATURE_MAX 3 /* CV_HAAR_FEATURE_MAX, tempcv.hpp:67 */

static __thread char g_err[512];
const char *vjo_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); } while (0)

/* cvRound: round-half-to-even under the default rounding mode. */
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }

/* ------------------------------------------------------------------------------------
 * Cascade: public structure (tempcv.hpp:70-112) + hidden cascade (tempcv.cpp:69-128)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int x, y, w, h;
} rect_t;

typedef struct {
    /* public */
    int tilted;
    rect_t r[VJO_FEATURE_MAX];
    float xml_weight[VJO_FEATURE_MAX];
    float threshold;
    int left, right;
    /* hidden (scale 1) */
    int nrects;                    /* 2 or 3: rect[2] kept? (tempcv.cpp:453-458) */
    float weight[VJO_FEATURE_MAX]; /* tempcv.cpp:752,760 */
    int off[VJO_FEATURE_MAX][4];   /* p0..p3 as (dy,dx) pairs packed later per level */
    int dy[VJO_FEATURE_MAX][4], dx[VJO_FEATURE_MAX][4];
} node_t;

typedef struct {
    int count;     /* nodes */
    node_t *node;  /* into cascade->nodes */
    float *alpha;  /* count+1 */
} tree_t;

typedef struct {
    int count;       /* trees */
    float threshold; /* xml - 0.0001f (tempcv.cpp:419) */
    float xml_threshold;
    tree_t *tree;
    int two_rects;
    int parent, next, child;
} stage_t;

struct vjo_cascade {
    int win_w, win_h;
    int count;
    int is_tree, is_stump_based, has_tilted;
    double inv_window_area;
    stage_t *stage;
    tree_t *trees;
    node_t *nodes;
    float *alphas;
    int n_trees, n_nodes;
};

void vjo_cascade_free(vjo_cascade *c)
{
    if (!c) return;
    free(c->stage); free(c->trees); free(c->nodes); free(c->alphas); free(c);
}

/* The reference halves the weights of tilted features (tempcv.cpp:733).  OpenCV 4.x's evaluator
 * does not; the soft pin against cv2.CascadeClassifier (tests/test_oracle_pins.py) sets
 * VJO_TEST_TILTED_CORRECTION=1 to compare the tilted GEOMETRY with that implementation.
 * Never set outside that test. */
static double tilted_correction(void)
{
    const char *e = getenv("VJO_TEST_TILTED_CORRECTION");
    return e ? atof(e) : 0.5;
}

vjo_cascade *vjo_cascade_create(int win_w, int win_h, int n_stages,
                                const int *st_ntrees, const float *st_thr,
                                const int *st_parent, const int *st_next,
                                const int *tr_nnodes,
                                const int *nd_tilted, const int *nd_rect,
                                const float *nd_weight, const float *nd_thr,
                                const int *nd_left, const int *nd_right,
                                const float *alpha)
{
    const float stage_threshold_bias = 0.0001f; /* tempcv.cpp:262 */
    if (n_stages <= 0) { FAIL("Number of stages should be positive"); return NULL; }
    if (win_w <= 0 || win_h <= 0) { FAIL("Invalid size node"); return NULL; }
    vjo_cascade *c = (vjo_cascade *)calloc(1, sizeof *c);
    c->win_w = win_w; c->win_h = win_h; c->count = n_stages;
    int T = 0, N = 0;
    for (int i = 0; i < n_stages; i++) {
        if (st_ntrees[i] <= 0) { FAIL("header of the stage classifier #%d is invalid", i); free(c); return NULL; }
        T += st_ntrees[i];
    }
    for (int t = 0; t < T; t++) {
        if (tr_nnodes[t] <= 0) { FAIL("Tree node is not a valid sequence (tree %d)", t); free(c); return NULL; }
        N += tr_nnodes[t];
    }
    c->n_trees = T; c->n_nodes = N;
    c->stage = (stage_t *)calloc(n_stages, sizeof(stage_t));
    c->trees = (tree_t *)calloc(T, sizeof(tree_t));
    c->nodes = (node_t *)calloc(N, sizeof(node_t));
    c->alphas = (float *)malloc(sizeof(float) * (N + T));
    memcpy(c->alphas, alpha, sizeof(float) * (N + T));

    /* --- icvReadHaarClassifier tail: parent/next/child (tempcv.cpp:2056-2083) and
     *     icvCreateHidHaarClassifierCascade (tempcv.cpp:308-467) --- */
    c->is_stump_based = 1; c->is_tree = 0; c->has_tilted = 0;
    for (int i = 0; i < n_stages; i++) c->stage[i].child = -1;
    int ti = 0, ni = 0, ai = 0;
    for (int i = 0; i < n_stages; i++) {
        stage_t *st = &c->stage[i];
        st->count = st_ntrees[i];
        st->xml_threshold = st_thr[i];
        st->threshold = st_thr[i] - stage_threshold_bias; /* float - float, tempcv.cpp:419 */
        st->tree = &c->trees[ti];
        st->two_rects = 1;
        st->parent = st_parent[i];
        st->next = st_next[i];
        if (st->parent < -1 || st->parent >= n_stages || st->next < -1 || st->next >= n_stages) {
            FAIL("parent/next must be a stage index or -1 (stage %d)", i); vjo_cascade_free(c); return NULL;
        }
        if (st->parent != -1 && c->stage[st->parent].child == -1) /* tempcv.cpp:2080-2083 */
            c->stage[st->parent].child = i;
        c->is_tree |= st->next != -1; /* tempcv.cpp:431 */
        for (int j = 0; j < st->count; j++, ti++) {
            tree_t *tr = &c->trees[ti];
            tr->count = tr_nnodes[ti];
            tr->node = &c->nodes[ni];
            tr->alpha = &c->alphas[ai];
            ai += tr->count + 1;
            for (int l = 0; l < tr->count; l++, ni++) {
                node_t *nd = &c->nodes[ni];
                nd->tilted = nd_tilted[ni] != 0;
                nd->threshold = nd_thr[ni];
                nd->left = nd_left[ni];
                nd->right = nd_right[ni];
                if (nd->left >= tr->count || nd->right >= tr->count ||
                    -nd->left > tr->count || -nd->right > tr->count) {
                    FAIL("Tree structure is broken (stage %d, tree %d, node %d)", i, j, l);
                    vjo_cascade_free(c); return NULL;
                }
                for (int k = 0; k < VJO_FEATURE_MAX; k++) {
                    nd->r[k].x = nd_rect[(ni * 3 + k) * 4 + 0];
                    nd->r[k].y = nd_rect[(ni * 3 + k) * 4 + 1];
                    nd->r[k].w = nd_rect[(ni * 3 + k) * 4 + 2];
                    nd->r[k].h = nd_rect[(ni * 3 + k) * 4 + 3];
                    nd->xml_weight[k] = nd_weight[ni * 3 + k];
                    if (nd->r[k].w) { /* bounds check, tempcv.cpp:367-386 */
                        rect_t r = nd->r[k];
                        c->has_tilted |= nd->tilted;
                        if (r.w < 0 || r.h < 0 || r.y < 0 || r.x + r.w > win_w ||
                            (!nd->tilted && (r.x < 0 || r.y + r.h > win_h)) ||
                            (nd->tilted && (r.x - r.h < 0 || r.y + r.w + r.h > win_h))) {
                            FAIL("rectangle #%d of the classifier #%d of the stage classifier #%d is not "
                                 "inside the reference (original) cascade window", k, j, i);
                            vjo_cascade_free(c); return NULL;
                        }
                    }
                }
                /* tempcv.cpp:453-458 */
                if (fabs(nd->xml_weight[2]) < DBL_EPSILON || nd->r[2].w == 0 || nd->r[2].h == 0)
                    nd->nrects = 2;
                else {
                    nd->nrects = 3;
                    st->two_rects = 0;
                }
            }
            c->is_stump_based &= tr->count == 1; /* tempcv.cpp:465 */
        }
    }

    /* --- cvSetImagesForHaarClassifierCascade at scale = 1 (tempcv.cpp:614-618,636-760).
     *     In REF-SI the scale is always 1 (tempcv.cpp:1321), so this is level-independent. */
    {
        const double scale = 1.;
        int eq_w = cv_round((win_w - 2) * scale), eq_h = cv_round((win_h - 2) * scale);
        double weight_scale = 1. / (eq_w * eq_h);
        c->inv_window_area = weight_scale;
        for (int n = 0; n < N; n++) {
            node_t *nd = &c->nodes[n];
            double sum0 = 0, area0 = 0;
            /* "align blocks" (tempcv.cpp:660-676): only matters when kx/ky <= 0, where the
             * reference would divide by zero (tempcv.cpp:681,688); reject such cascades. */
            int base_w = -1, base_h = -1;
            for (int k = 0; k < nd->nrects; k++) {
                unsigned a;
                a = (unsigned)(nd->r[k].w - 1); if (a < (unsigned)base_w) base_w = (int)a;
                a = (unsigned)(nd->r[k].x - nd->r[0].x - 1); if (a < (unsigned)base_w) base_w = (int)a;
                a = (unsigned)(nd->r[k].h - 1); if (a < (unsigned)base_h) base_h = (int)a;
                a = (unsigned)(nd->r[k].y - nd->r[0].y - 1); if (a < (unsigned)base_h) base_h = (int)a;
            }
            base_w += 1; base_h += 1;
            if (base_w <= 0 || base_h <= 0 || nd->r[0].w / base_w <= 0 || nd->r[0].h / base_h <= 0) {
                FAIL("node %d: kx/ky <= 0 (reference divides by zero, tempcv.cpp:678-690)", n);
                vjo_cascade_free(c); return NULL;
            }
            for (int k = 0; k < nd->nrects; k++) {
                rect_t tr;
                tr.x = cv_round(nd->r[k].x * scale); tr.w = cv_round(nd->r[k].w * scale);
                tr.y = cv_round(nd->r[k].y * scale); tr.h = cv_round(nd->r[k].h * scale);
                double correction_ratio = weight_scale * (!nd->tilted ? 1 : tilted_correction()); /* :733 */
                if (!nd->tilted) { /* :738-741 */
                    nd->dy[k][0] = tr.y;        nd->dx[k][0] = tr.x;
                    nd->dy[k][1] = tr.y;        nd->dx[k][1] = tr.x + tr.w;
                    nd->dy[k][2] = tr.y + tr.h; nd->dx[k][2] = tr.x;
                    nd->dy[k][3] = tr.y + tr.h; nd->dx[k][3] = tr.x + tr.w;
                } else { /* :745-749 */
                    nd->dy[k][2] = tr.y + tr.w;        nd->dx[k][2] = tr.x + tr.w;
                    nd->dy[k][3] = tr.y + tr.w + tr.h; nd->dx[k][3] = tr.x + tr.w - tr.h;
                    nd->dy[k][0] = tr.y;               nd->dx[k][0] = tr.x;
                    nd->dy[k][1] = tr.y + tr.h;        nd->dx[k][1] = tr.x - tr.h;
                }
                nd->weight[k] = (float)(nd->xml_weight[k] * correction_ratio); /* :752 */
                if (k == 0)
                    area0 = tr.w * tr.h;
                else
                    sum0 += nd->weight[k] * tr.w * tr.h; /* float*int*int -> float, :757 */
            }
            nd->weight[0] = (float)(-sum0 / area0); /* :760 */
        }
    }
    return c;
}

int vjo_cascade_flags(const vjo_cascade *c)
{
    return (c->is_tree ? 1 : 0) | (c->is_stump_based ? 2 : 0) | (c->has_tilted ? 4 : 0);
}

int vjo_cascade_counts(const vjo_cascade *c, int *n_stages, int *n_trees, int *n_nodes)
{
    if (n_stages) *n_stages = c->count;
    if (n_trees) *n_trees = c->n_trees;
    if (n_nodes) *n_nodes = c->n_nodes;
    return 0;
}

void vjo_cascade_hid(const vjo_cascade *c, float *node_weights, int *node_nrects,
                     float *stage_thr, int *stage_two_rects, int *stage_child)
{
    for (int n = 0; n < c->n_nodes; n++) {
        if (node_weights)
            for (int k = 0; k < 3; k++)
                node_weights[n * 3 + k] = k < c->nodes[n].nrects ? c->nodes[n].weight[k] : 0.f;
        if (node_nrects) node_nrects[n] = c->nodes[n].nrects;
    }
    for (int i = 0; i < c->count; i++) {
        if (stage_thr) stage_thr[i] = c->stage[i].threshold;
        if (stage_two_rects) stage_two_rects[i] = c->stage[i].two_rects;
        if (stage_child) stage_child[i] = c->stage[i].child;
    }
}

/* ------------------------------------------------------------------------------------
 * cvResize(INTER_LINEAR), 8UC1 (OpenCV imgproc; pinned against
 * cv2.resize in tests/test_oracle_pins.py)
 * ---------------------------------------------------------------------------------- */
#define RESIZE_COEF_BITS 11
#define RESIZE_COEF_SCALE (1 << RESIZE_COEF_BITS)

static inline short sat_short_round(float v)
{
    int iv = (int)lrintf(v);
    return (short)(iv < -32768 ? -32768 : iv > 32767 ? 32767 : iv);
}

int vjo_resize_linear(const uint8_t *src, int sw, int sh, int sstride,
                      uint8_t *dst, int dw, int dh, int dstride)
{
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) { FAIL("resize: empty size"); return -1; }
    double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    int *xofs = (int *)malloc(sizeof(int) * dw);
    short *ialpha = (short *)malloc(sizeof(short) * 2 * dw);
    int *yofs = (int *)malloc(sizeof(int) * dh);
    short *ibeta = (short *)malloc(sizeof(short) * 2 * dh);
    int *rows = (int *)malloc(sizeof(int) * 2 * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[dx * 2] = sat_short_round((1.f - fx) * RESIZE_COEF_SCALE);
        ialpha[dx * 2 + 1] = sat_short_round(fx * RESIZE_COEF_SCALE);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[dy * 2] = sat_short_round((1.f - fy) * RESIZE_COEF_SCALE);
        ibeta[dy * 2 + 1] = sat_short_round(fy * RESIZE_COEF_SCALE);
    }
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = yofs[dy], sy1 = yofs[dy] + 1;
        sy0 = sy0 < 0 ? 0 : sy0 >= sh ? sh - 1 : sy0;
        sy1 = sy1 < 0 ? 0 : sy1 >= sh ? sh - 1 : sy1;
        const uint8_t *S0 = src + (size_t)sy0 * sstride, *S1 = src + (size_t)sy1 * sstride;
        int *R0 = rows, *R1 = rows + dw;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            int a0 = ialpha[dx * 2], a1 = ialpha[dx * 2 + 1];
            R0[dx] = S0[sx] * a0 + S0[sx1] * a1;
            R1[dx] = S1[sx] * a0 + S1[sx1] * a1;
        }
        int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
        uint8_t *D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (R0[dx] >> 4)) >> 16) + ((b1 * (R1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xofs); free(ialpha); free(yofs); free(ibeta); free(rows);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * cvIntegral (OpenCV imgproc; pinned against cv2.integral3)
 * ---------------------------------------------------------------------------------- */
void vjo_integral(const uint8_t *img, int w, int h, int stride,
                  int32_t *sum, double *sqsum, int32_t *tilted)
{
    const int W1 = w + 1;
    for (int x = 0; x <= w; x++) { sum[x] = 0; sqsum[x] = 0; if (tilted) tilted[x] = 0; }
    for (int y = 0; y < h; y++) {
        const uint8_t *p = img + (size_t)y * stride;
        int32_t *s = sum + (size_t)(y + 1) * W1; const int32_t *sp = s - W1;
        double *q = sqsum + (size_t)(y + 1) * W1; const double *qp = q - W1;
        int32_t rs = 0; double rq = 0;
        s[0] = 0; q[0] = 0;
        for (int x = 0; x < w; x++) {
            rs += p[x]; rq += (double)(p[x] * p[x]);
            s[x + 1] = sp[x + 1] + rs;
            q[x + 1] = qp[x + 1] + rq;
        }
    }
    if (tilted) {
        /* tilted[Y][X] = sum_{y<Y, |x-(X-1)| <= (Y-1)-y} img[y][x]  (SURVEY A.3).
         * Column recurrence: T(Y,X) - T(Y-1,X) = A(Y-1,X-1) + B(Y-1,X-1) - I(Y-1,X-1) with
         * A/B the up-left / up-right diagonal prefix sums; X = 0 (apex at x = -1) only
         * picks up B(Y-2,0). */
        int32_t *A0 = (int32_t *)calloc((size_t)w + 2, sizeof(int32_t));
        int32_t *B0 = (int32_t *)calloc((size_t)w + 2, sizeof(int32_t));
        int32_t *A1 = (int32_t *)calloc((size_t)w + 2, sizeof(int32_t));
        int32_t *B1 = (int32_t *)calloc((size_t)w + 2, sizeof(int32_t));
        /* arrays are indexed x+1 so that x=-1 and x=w are zero pads */
        for (int Y = 1; Y <= h; Y++) {
            const uint8_t *p = img + (size_t)(Y - 1) * stride;
            int32_t *t = tilted + (size_t)Y * W1; const int32_t *tp = t - W1;
            /* previous row's B at x=0 is B0[1] */
            t[0] = tp[0] + B0[1];
            for (int x = 0; x < w; x++) {
                A1[x + 1] = p[x] + A0[x];     /* A(y,x) = I + A(y-1,x-1) */
                B1[x + 1] = p[x] + B0[x + 2]; /* B(y,x) = I + B(y-1,x+1) */
                t[x + 1] = tp[x + 1] + A1[x + 1] + B1[x + 1] - p[x];
            }
            int32_t *sw;
            sw = A0; A0 = A1; A1 = sw;
            sw = B0; B0 = B1; B1 = sw;
        }
        free(A0); free(B0); free(A1); free(B1);
    }
}

/* ------------------------------------------------------------------------------------
 * Level loop (tempcv.cpp:1230-1234,1268-1288) + window grid (tempcv.cpp:1013-1021)
 * ---------------------------------------------------------------------------------- */
int vjo_plan_levels(int W, int H, int w0, int h0, double scale_factor,
                    int min_w, int min_h, int max_w, int max_h,
                    vjo_level *levels, int max_levels)
{
    if (scale_factor <= 1) { FAIL("scale factor must be > 1"); return -1; }
    if (max_h == 0 || max_w == 0) { max_h = H; max_w = W; } /* :1230-1234 */
    int n = 0;
    for (double factor = 1;; factor *= scale_factor) {
        int win_w = cv_round(w0 * factor), win_h = cv_round(h0 * factor);
        int sz_w = cv_round(W / factor), sz_h = cv_round(H / factor);
        int sz1_w = sz_w - w0 + 1, sz1_h = sz_h - h0 + 1;
        if (sz1_w <= 0 || sz1_h <= 0) break;
        if (win_w > max_w || win_h > max_h) break;
        if (win_w < min_w || win_h < min_h) continue;
        if (n >= max_levels) { FAIL("too many pyramid levels"); return -1; }
        vjo_level *L = &levels[n++];
        L->factor = factor; L->img_w = sz_w; L->img_h = sz_h; L->win_w = win_w; L->win_h = win_h;
        L->ystep = factor > 2 ? 1 : 2; /* :1021 */
        /* invoker: y in [0, min(stripSize, sum.rows-1-h0)) = [0, sz_h-h0); x in [0, sz_w-w0) */
        int xe = sz_w - w0, ye = sz_h - h0;
        L->nx = xe > 0 ? (xe + L->ystep - 1) / L->ystep : 0;
        L->ny = ye > 0 ? (ye + L->ystep - 1) / L->ystep : 0;
        if (sz_w + 1 <= 1 + w0) { L->nx = 0; } /* :1017 */
        if (L->nx == 0 || L->ny == 0) n--; /* no window fits the grid (:1017-1018): nothing to do */
    }
    return n;
}

/* ------------------------------------------------------------------------------------
 * Window evaluation: cvRunHaarClassifierCascadeSum (tempcv.cpp:795-972) and
 * icvEvalHidHaarClassifier (tempcv.cpp:771-792)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    const int32_t *sum, *tilted;
    const double *sqsum;
    int step; /* elements per row, same for all three */
} level_img;

#define CALC_SUM(base, nd, k, o) \
    ((base)[(o) + (nd)->off[k][0]] - (base)[(o) + (nd)->off[k][1]] - \
     (base)[(o) + (nd)->off[k][2]] + (base)[(o) + (nd)->off[k][3]])

typedef struct {
    /* per-level resolved offsets: node offsets live in a per-thread-shared copy */
    node_t *nodes; /* copy of cascade nodes with off[][] resolved for this level step */
} level_nodes;

static inline double eval_tree(const tree_t *tr, const node_t *nodes_base, const node_t *lvl_nodes,
                               const level_img *im, double vnf, int p_offset, int64_t *node_evals)
{
    /* icvEvalHidHaarClassifier, tempcv.cpp:771-792 */
    int idx = 0;
    const node_t *first = lvl_nodes + (tr->node - nodes_base);
    do {
        const node_t *node = first + idx;
        const int32_t *base = node->tilted ? im->tilted : im->sum;
        double t = node->threshold * vnf;
        double sum = CALC_SUM(base, node, 0, p_offset) * node->weight[0]; /* int*float */
        sum += CALC_SUM(base, node, 1, p_offset) * node->weight[1];
        if (node->nrects == 3)
            sum += CALC_SUM(base, node, 2, p_offset) * node->weight[2];
        idx = sum < t ? node->left : node->right;
        (*node_evals)++;
    } while (idx > 0);
    return tr->alpha[-idx];
}

/* returns the parity code (see vj_oracle.h) */
static int run_window(const vjo_cascade *c, const node_t *lvl_nodes, const level_img *im,
                      int x, int y, int eq_off[4], int *near_flag, vjo_stats *st, double *last_sum)
{
    double last_stage_sum = 0.0; /* the stage_sum cvRunHaarClassifierCascadeSum hands back (tempcv.cpp:797) */
    int p_offset = y * im->step + x;
    int near = 0;
    /* variance normalisation, tempcv.cpp:822-832 */
    double mean = im->sum[p_offset + eq_off[0]] - im->sum[p_offset + eq_off[1]] -
                  im->sum[p_offset + eq_off[2]] + im->sum[p_offset + eq_off[3]];
    mean *= c->inv_window_area;
    double vnf = im->sqsum[p_offset + eq_off[0]] - im->sqsum[p_offset + eq_off[1]] -
                 im->sqsum[p_offset + eq_off[2]] + im->sqsum[p_offset + eq_off[3]];
    vnf = vnf * c->inv_window_area - mean * mean;
    if (vnf >= 0.)
        vnf = sqrt(vnf);
    else
        vnf = 1.;

    int64_t weak = 0, nodes = 0;
    int code;
#define NEAR_CHECK(S, T) do { double _t = (double)(T); \
        if (fabs((S) - _t) <= 1e-5 * fabs(_t)) { near = 1; st->near_stage_events++; } } while (0)

    if (c->is_tree) { /* tempcv.cpp:834-861 */
        int ptr = 0, last = 0, accepted = 0;
        for (;;) {
            const stage_t *s = &c->stage[ptr];
            double stage_sum = 0.0;
            last = ptr;
            if (ptr < 64) st->stage_reach[ptr]++;
            for (int j = 0; j < s->count; j++)
                stage_sum += eval_tree(&s->tree[j], c->nodes, lvl_nodes, im, vnf, p_offset, &nodes);
            weak += s->count;
            last_stage_sum = stage_sum;
            NEAR_CHECK(stage_sum, s->threshold);
            if (stage_sum >= s->threshold) {
                ptr = s->child;
                if (ptr < 0) { accepted = 1; break; }
            } else {
                while (ptr >= 0 && c->stage[ptr].next < 0) ptr = c->stage[ptr].parent;
                if (ptr < 0) { accepted = 0; break; }
                ptr = c->stage[ptr].next;
            }
        }
        code = 2 * last + accepted;
        if (accepted) st->accepted++;
    } else if (c->is_stump_based) { /* tempcv.cpp:862-949 */
        int i;
        for (i = 0; i < c->count; i++) {
            const stage_t *s = &c->stage[i];
            double stage_sum = 0.0;
            if (i < 64) st->stage_reach[i]++;
            if (s->two_rects) { /* :872-898 */
                for (int j = 0; j < s->count; j++) {
                    const node_t *node = lvl_nodes + (s->tree[j].node - c->nodes);
                    const int32_t *base = node->tilted ? im->tilted : im->sum;
                    double t = node->threshold * vnf;
                    double rect0 = CALC_SUM(base, node, 0, p_offset);
                    rect0 *= node->weight[0];
                    double rect1 = CALC_SUM(base, node, 1, p_offset);
                    rect1 *= node->weight[1];
                    double sum = rect1 + rect0;
                    stage_sum += s->tree[j].alpha[sum >= t];
                }
            } else { /* :899-930 */
                for (int j = 0; j < s->count; j++) {
                    const node_t *node = lvl_nodes + (s->tree[j].node - c->nodes);
                    const int32_t *base = node->tilted ? im->tilted : im->sum;
                    double t = node->threshold * vnf;
                    double sum = CALC_SUM(base, node, 0, p_offset) * node->weight[0];
                    sum += CALC_SUM(base, node, 1, p_offset) * node->weight[1];
                    if (node->nrects == 3)
                        sum += CALC_SUM(base, node, 2, p_offset) * node->weight[2];
                    stage_sum += s->tree[j].alpha[sum >= t];
                }
            }
            weak += s->count; nodes += s->count;
            last_stage_sum = stage_sum;
            NEAR_CHECK(stage_sum, s->threshold);
            if (stage_sum < s->threshold) break; /* return -i, :946 */
        }
        code = i; /* stages passed; == count => accepted (return 1, :971) */
        if (i == c->count) st->accepted++;
    } else { /* tempcv.cpp:950-966 */
        int i;
        for (i = 0; i < c->count; i++) {
            const stage_t *s = &c->stage[i];
            double stage_sum = 0.0;
            if (i < 64) st->stage_reach[i]++;
            for (int j = 0; j < s->count; j++)
                stage_sum += eval_tree(&s->tree[j], c->nodes, lvl_nodes, im, vnf, p_offset, &nodes);
            weak += s->count;
            last_stage_sum = stage_sum;
            NEAR_CHECK(stage_sum, s->threshold);
            if (stage_sum < s->threshold) break;
        }
        code = i;
        if (i == c->count) st->accepted++;
    }
    if (last_sum) *last_sum = last_stage_sum;
    st->windows++;
    st->weak_evals += weak;
    st->node_evals += nodes;
    st->near_stage_thr += near;
    *near_flag = near;
    return code;
}

static void stats_add(vjo_stats *a, const vjo_stats *b)
{
    a->windows += b->windows; a->weak_evals += b->weak_evals; a->node_evals += b->node_evals;
    a->accepted += b->accepted; a->near_stage_thr += b->near_stage_thr; a->near_stage_events += b->near_stage_events;
    for (int i = 0; i < 64; i++) a->stage_reach[i] += b->stage_reach[i];
}

/* evaluate the ystep grid of one level given its integral images */
static void eval_grid(const vjo_cascade *c, const level_img *im, int lw, int lh, int ystep,
                      int nx, int ny, int16_t *codes, uint8_t *near, vjo_stats *stats, int n_threads, double *last_sums)
{
    (void)lw; (void)lh;
    /* resolve corner offsets for this level's row step (the reference stores pointers,
     * tempcv.cpp:620-630,738-749) */
    node_t *lvl = (node_t *)malloc(sizeof(node_t) * c->n_nodes);
    memcpy(lvl, c->nodes, sizeof(node_t) * c->n_nodes);
    for (int n = 0; n < c->n_nodes; n++)
        for (int k = 0; k < lvl[n].nrects; k++)
            for (int q = 0; q < 4; q++)
                lvl[n].off[k][q] = lvl[n].dy[k][q] * im->step + lvl[n].dx[k][q];
    int ex = 1, ey = 1, ew = c->win_w - 2, eh = c->win_h - 2; /* equRect at scale 1, :614-616 */
    int eq_off[4] = { ey * im->step + ex, ey * im->step + ex + ew,
                      (ey + eh) * im->step + ex, (ey + eh) * im->step + ex + ew };
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
#pragma omp parallel num_threads(n_threads)
    {
        vjo_stats local; memset(&local, 0, sizeof local);
#pragma omp for schedule(dynamic, 4)
        for (int iy = 0; iy < ny; iy++) {
            for (int ix = 0; ix < nx; ix++) {
                int nf = 0;
                int code = run_window(c, lvl, im, ix * ystep, iy * ystep, eq_off, &nf, &local,
                                      last_sums ? last_sums + (size_t)iy * nx + ix : NULL);
                if (codes) codes[(size_t)iy * nx + ix] = (int16_t)code;
                if (near) near[(size_t)iy * nx + ix] = (uint8_t)nf;
                else (void)nf;
            }
        }
#pragma omp critical
        stats_add(stats, &local);
    }
    free(lvl);
}

static inline int code_accepts(const vjo_cascade *c, int code)
{
    return c->is_tree ? (code & 1) : (code == c->count);
}

int64_t vjo_eval_level(const vjo_cascade *c, const uint8_t *img, int w, int h, int stride,
                       int ystep, int16_t *codes, uint8_t *near, vjo_stats *stats, int n_threads)
{
    vjo_stats local; memset(&local, 0, sizeof local);
    int xe = w - c->win_w, ye = h - c->win_h;
    int nx = xe > 0 ? (xe + ystep - 1) / ystep : 0, ny = ye > 0 ? (ye + ystep - 1) / ystep : 0;
    if (nx == 0 || ny == 0) { if (stats) *stats = local; return 0; }
    size_t n1 = (size_t)(w + 1) * (h + 1);
    int32_t *sum = (int32_t *)malloc(n1 * sizeof(int32_t));
    double *sq = (double *)malloc(n1 * sizeof(double));
    int32_t *tl = c->has_tilted ? (int32_t *)malloc(n1 * sizeof(int32_t)) : NULL;
    vjo_integral(img, w, h, stride, sum, sq, tl);
    level_img im = { sum, tl, sq, w + 1 };
    eval_grid(c, &im, w, h, ystep, nx, ny, codes, near, &local, n_threads, NULL);
    free(sum); free(sq); free(tl);
    if (stats) *stats = local;
    return local.accepted;
}

int64_t vjo_detect(const vjo_cascade *c, const uint8_t *img, int W, int H, int stride,
                   double scale_factor, int min_w, int min_h, int max_w, int max_h,
                   int32_t *rects, int64_t cap, int16_t *codes, uint8_t *near,
                   vjo_stats *stats, int n_threads)
{
    return vjo_detect_roc(c, img, W, H, stride, scale_factor, min_w, min_h, max_w, max_h, rects, NULL, NULL, cap,
                          codes, near, stats, n_threads);
}

/* cvHaarDetectObjectsForROC with outputRejectLevels (tempcv.cpp:1084-1094): besides the accepted
 * windows (level = stage count) also the windows rejected by one of the last three stages, each
 * with the stage sum of the last stage it evaluated.  reject_levels == NULL: plain detection. */
int64_t vjo_detect_roc(const vjo_cascade *c, const uint8_t *img, int W, int H, int stride,
                       double scale_factor, int min_w, int min_h, int max_w, int max_h,
                       int32_t *rects, int32_t *reject_levels, double *level_weights, int64_t cap,
                       int16_t *codes, uint8_t *near, vjo_stats *stats, int n_threads)
{
    vjo_level lv[256];
    int nl = vjo_plan_levels(W, H, c->win_w, c->win_h, scale_factor, min_w, min_h, max_w, max_h, lv, 256);
    if (nl < 0) return -1;
    vjo_stats total; memset(&total, 0, sizeof total);
    size_t n1 = (size_t)(W + 1) * (H + 1);
    uint8_t *small = (uint8_t *)malloc((size_t)W * H);
    int32_t *sum = (int32_t *)malloc(n1 * sizeof(int32_t));
    double *sq = (double *)malloc(n1 * sizeof(double));
    int32_t *tl = c->has_tilted ? (int32_t *)malloc(n1 * sizeof(int32_t)) : NULL;
    int64_t n_out = 0; size_t woff = 0;
    for (int l = 0; l < nl; l++) {
        const vjo_level *L = &lv[l];
        size_t nwin = (size_t)L->nx * L->ny;
        if (nwin == 0) continue;
        /* tempcv.cpp:1301-1302: every level is resized from the ORIGINAL image */
        vjo_resize_linear(img, W, H, stride, small, L->img_w, L->img_h, L->img_w);
        vjo_integral(small, L->img_w, L->img_h, L->img_w, sum, sq, tl);
        level_img im = { sum, tl, sq, L->img_w + 1 };
        int16_t *lc = codes ? codes + woff : (int16_t *)malloc(nwin * sizeof(int16_t));
        double *ls = reject_levels ? (double *)malloc(nwin * sizeof(double)) : NULL;
        eval_grid(c, &im, L->img_w, L->img_h, L->ystep, L->nx, L->ny, lc, near ? near + woff : NULL,
                  &total, n_threads, ls);
        for (int iy = 0; iy < L->ny; iy++)
            for (int ix = 0; ix < L->nx; ix++) {
                const int code = lc[(size_t)iy * L->nx + ix];
                int take = code_accepts(c, code), level = c->count;
                if (reject_levels) { /* tempcv.cpp:1084-1094 */
                    /* result: 1 accepted; linear cascade -i for a rejection by stage i; stage tree 0 */
                    int result = take ? 1 : (c->is_tree ? 0 : -code);
                    if (result == 1) result = -1 * c->count;
                    take = c->count + result < 4;
                    level = -result;
                }
                if (take) {
                    if (rects && n_out < cap) { /* tempcv.cpp:1099-1100 */
                        rects[n_out * 4 + 0] = cv_round(ix * L->ystep * L->factor);
                        rects[n_out * 4 + 1] = cv_round(iy * L->ystep * L->factor);
                        rects[n_out * 4 + 2] = L->win_w;
                        rects[n_out * 4 + 3] = L->win_h;
                        if (reject_levels) { reject_levels[n_out] = level; level_weights[n_out] = ls[(size_t)iy * L->nx + ix]; }
                    }
                    n_out++;
                }
            }
        if (!codes) free(lc);
        free(ls);
        woff += nwin;
    }
    free(small); free(sum); free(sq); free(tl);
    if (stats) *stats = total;
    return n_out;
}

/* ------------------------------------------------------------------------------------
 * REF-SC: the scale-cascade path of cvHaarDetectObjectsForROC (tempcv.cpp:1330-1456 with
 * flags = 0: no Canny pruning, no biggest-object search), i.e. what main.cpp:145 runs:
 * ONE integral image of the full frame, the FEATURES scaled per factor by
 * cvSetImagesForHaarClassifierCascade (tempcv.cpp:549-768), windows on the grid of
 * HaarDetectObjects_ScaleCascade_Invoker (tempcv.cpp:1132-1175) including its skip rule:
 * after a window whose cvRunHaarClassifierCascade result is 0 the next x position is
 * skipped (result 0 = rejected by stage 0 of a linear cascade, tempcv.cpp:946 "-i", or ANY
 * rejection of a stage-tree cascade, tempcv.cpp:857).
 * ---------------------------------------------------------------------------------- */
int vjo_plan_sc(int W, int H, int w0, int h0, double scale_factor, int min_w, int min_h,
                vjo_level *levels, int max_levels)
{
    if (!(scale_factor > 1)) { FAIL("scale factor must be > 1"); return -1; }
    int n_factors = 0, n = 0;
    double factor;
    for (factor = 1; factor * w0 < W - 10 && factor * h0 < H - 10; factor *= scale_factor) n_factors++; /* :1344-1350 */
    factor = 1;
    for (; n_factors-- > 0; factor *= scale_factor) { /* :1361 */
        const double ystep = factor > 2. ? factor : 2.; /* :1365 */
        const int win_w = cv_round(w0 * factor), win_h = cv_round(h0 * factor);
        const int endX = cv_round((W - win_w) / ystep), endY = cv_round((H - win_h) / ystep); /* :1371-1372 */
        if (win_w < min_w || win_h < min_h) continue; /* :1374-1379 */
        if (n >= max_levels) { FAIL("too many scales"); return -1; }
        vjo_level *L = &levels[n++];
        L->factor = factor; L->img_w = W; L->img_h = H; L->win_w = win_w; L->win_h = win_h;
        L->ystep = 0; /* the step is max(2, factor), a double */
        L->nx = endX > 0 ? endX : 0; L->ny = endY > 0 ? endY : 0;
    }
    return n;
}

/* cvSetImagesForHaarClassifierCascade(scale) on a copy of the nodes (tempcv.cpp:614-618,636-760) */
static void sc_set_scale(const vjo_cascade *c, double scale, int step, node_t *lvl, double *inv_area, int eq_off[4])
{
    const int ex = cv_round(scale), ey = ex; /* :614 */
    const int ew = cv_round((c->win_w - 2) * scale), eh = cv_round((c->win_h - 2) * scale);
    const double weight_scale = 1. / (ew * eh);
    *inv_area = weight_scale;
    eq_off[0] = ey * step + ex; eq_off[1] = ey * step + ex + ew;
    eq_off[2] = (ey + eh) * step + ex; eq_off[3] = (ey + eh) * step + ex + ew;
    memcpy(lvl, c->nodes, sizeof(node_t) * c->n_nodes);
    for (int n = 0; n < c->n_nodes; n++) {
        node_t *nd = &lvl[n];
        double sum0 = 0, area0 = 0;
        for (int k = 0; k < nd->nrects; k++) { /* kx, ky >= 1 (checked at creation): the flagx/flagy branch is dead */
            rect_t tr;
            tr.x = cv_round(nd->r[k].x * scale); tr.w = cv_round(nd->r[k].w * scale);
            tr.y = cv_round(nd->r[k].y * scale); tr.h = cv_round(nd->r[k].h * scale);
            const double correction_ratio = weight_scale * (!nd->tilted ? 1 : 0.5);
            int dy[4], dx[4];
            if (!nd->tilted) {
                dy[0] = tr.y;        dx[0] = tr.x;
                dy[1] = tr.y;        dx[1] = tr.x + tr.w;
                dy[2] = tr.y + tr.h; dx[2] = tr.x;
                dy[3] = tr.y + tr.h; dx[3] = tr.x + tr.w;
            } else {
                dy[2] = tr.y + tr.w;        dx[2] = tr.x + tr.w;
                dy[3] = tr.y + tr.w + tr.h; dx[3] = tr.x + tr.w - tr.h;
                dy[0] = tr.y;               dx[0] = tr.x;
                dy[1] = tr.y + tr.h;        dx[1] = tr.x - tr.h;
            }
            for (int q = 0; q < 4; q++) nd->off[k][q] = dy[q] * step + dx[q];
            nd->weight[k] = (float)(nd->xml_weight[k] * correction_ratio);
            if (k == 0) area0 = tr.w * tr.h;
            else sum0 += nd->weight[k] * tr.w * tr.h;
        }
        nd->weight[0] = (float)(-sum0 / area0);
    }
}

int64_t vjo_detect_sc(const vjo_cascade *c, const uint8_t *img, int W, int H, int stride,
                      double scale_factor, int min_w, int min_h,
                      int32_t *rects, int64_t cap, int16_t *codes, vjo_stats *stats, int n_threads)
{
    vjo_level lv[256];
    const int nl = vjo_plan_sc(W, H, c->win_w, c->win_h, scale_factor, min_w, min_h, lv, 256);
    if (nl < 0) return -1;
    vjo_stats total; memset(&total, 0, sizeof total);
    const size_t n1 = (size_t)(W + 1) * (H + 1);
    int32_t *sum = (int32_t *)malloc(n1 * sizeof(int32_t));
    double *sq = (double *)malloc(n1 * sizeof(double));
    int32_t *tl = c->has_tilted ? (int32_t *)malloc(n1 * sizeof(int32_t)) : NULL;
    node_t *lvl = (node_t *)malloc(sizeof(node_t) * c->n_nodes);
    vjo_integral(img, W, H, stride, sum, sq, tl); /* :1335 */
    const level_img im = { sum, tl, sq, W + 1 };
    int64_t n_out = 0; size_t woff = 0;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    for (int l = 0; l < nl; l++) {
        const vjo_level *L = &lv[l];
        const size_t nwin = (size_t)L->nx * L->ny;
        if (nwin == 0) continue;
        const double ystep = L->factor > 2. ? L->factor : 2.;
        vjo_cascade cs = *c; /* shallow copy: only inv_window_area differs per scale */
        int eq_off[4];
        sc_set_scale(c, L->factor, im.step, lvl, &cs.inv_window_area, eq_off);
        int16_t *lc = codes ? codes + woff : (int16_t *)malloc(nwin * sizeof(int16_t));
#pragma omp parallel num_threads(n_threads)
        {
            vjo_stats local; memset(&local, 0, sizeof local);
#pragma omp for schedule(dynamic, 4)
            for (int iy = 0; iy < L->ny; iy++) { /* :1139-1165 */
                const int y = cv_round(iy * ystep);
                int ixstep = 1;
                for (int ix = 0; ix < L->nx; ix++) lc[(size_t)iy * L->nx + ix] = VJO_CODE_SKIPPED;
                for (int ix = 0; ix < L->nx; ix += ixstep) {
                    const int x = cv_round(ix * ystep);
                    int code, result;
                    if (x < 0 || y < 0 || x + L->win_w >= W + 1 || y + L->win_h >= H + 1) { /* :817-820: returns -1 */
                        code = VJO_CODE_OUTSIDE; result = -1;
                    } else {
                        int nf = 0;
                        code = run_window(&cs, lvl, &im, x, y, eq_off, &nf, &local, NULL);
                        if (c->is_tree) result = code & 1;               /* 0 on any rejection, :857 */
                        else result = code == c->count ? 1 : -code;      /* -i, :946,966 */
                    }
                    lc[(size_t)iy * L->nx + ix] = (int16_t)code;
                    ixstep = result != 0 ? 1 : 2; /* :1161 */
                }
            }
#pragma omp critical
            stats_add(&total, &local);
        }
        for (int iy = 0; iy < L->ny; iy++)
            for (int ix = 0; ix < L->nx; ix++) {
                const int code = lc[(size_t)iy * L->nx + ix];
                if (code >= 0 && code_accepts(c, code)) {
                    if (rects && n_out < cap) { /* :1159-1160 */
                        rects[n_out * 4 + 0] = cv_round(ix * ystep);
                        rects[n_out * 4 + 1] = cv_round(iy * ystep);
                        rects[n_out * 4 + 2] = L->win_w;
                        rects[n_out * 4 + 3] = L->win_h;
                    }
                    n_out++;
                }
            }
        if (!codes) free(lc);
        woff += nwin;
    }
    free(sum); free(sq); free(tl); free(lvl);
    if (stats) *stats = total;
    return n_out;
}

/* ------------------------------------------------------------------------------------
 * AgroupRectangles (tempcv.cpp:130-243); cv::partition is external OpenCV (call site
 * tempcv.cpp:160): connected components of the similarity graph, classes numbered in
 * order of their first member.
 * ---------------------------------------------------------------------------------- */
static int similar_rects(const int32_t *a, const int32_t *b, double eps)
{ /* ASimilarRects, tempcv.cpp:134-141 */
    int mw = a[2] < b[2] ? a[2] : b[2], mh = a[3] < b[3] ? a[3] : b[3];
    double delta = eps * (mw + mh) * 0.5;
    return abs(a[0] - b[0]) <= delta && abs(a[1] - b[1]) <= delta &&
           abs(a[0] + a[2] - b[0] - b[2]) <= delta && abs(a[1] + a[3] - b[1] - b[3]) <= delta;
}

static int uf_find(int *parent, int i)
{
    while (parent[i] != i) { parent[i] = parent[parent[i]]; i = parent[i]; }
    return i;
}

/* AgroupRectangles (tempcv.cpp:145-243).  level_weights != NULL: the ROC variant (:255-258) -- `weights`
 * then carries the reject levels in and the winning level per class out */
static int group_impl(int32_t *rects, int n, int group_threshold, double eps, int32_t *weights, double *level_weights)
{
    if (group_threshold <= 0 || n == 0) { /* tempcv.cpp:147-157 */
        if (weights) for (int i = 0; i < n; i++) weights[i] = 1;
        return n;
    }
    int *parent = (int *)malloc(sizeof(int) * n), *labels = (int *)malloc(sizeof(int) * n);
    int *cls_of_root = (int *)malloc(sizeof(int) * n);
    for (int i = 0; i < n; i++) { parent[i] = i; cls_of_root[i] = -1; }
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            if (similar_rects(rects + 4 * i, rects + 4 * j, eps)) {
                int a = uf_find(parent, i), b = uf_find(parent, j);
                if (a != b) parent[b] = a;
            }
    int nclasses = 0;
    for (int i = 0; i < n; i++) {
        int r = uf_find(parent, i);
        if (cls_of_root[r] < 0) cls_of_root[r] = nclasses++;
        labels[i] = cls_of_root[r];
    }
    int32_t *rr = (int32_t *)calloc((size_t)nclasses * 4, sizeof(int32_t));
    int *rw = (int *)calloc(nclasses, sizeof(int));
    for (int i = 0; i < n; i++) { /* :167-175 */
        int cl = labels[i];
        for (int k = 0; k < 4; k++) rr[cl * 4 + k] += rects[i * 4 + k];
        rw[cl]++;
    }
    int *rej = (int *)calloc(nclasses, sizeof(int));
    double *rejw = (double *)malloc(sizeof(double) * (nclasses ? nclasses : 1));
    for (int i = 0; i < nclasses; i++) rejw[i] = DBL_MIN; /* :165 */
    if (level_weights && weights) /* :176-189 */
        for (int i = 0; i < n; i++) {
            int cl = labels[i];
            if (weights[i] > rej[cl]) { rej[cl] = weights[i]; rejw[cl] = level_weights[i]; }
            else if (weights[i] == rej[cl] && level_weights[i] > rejw[cl]) rejw[cl] = level_weights[i];
        }
    for (int i = 0; i < nclasses; i++) { /* :191-199 */
        float s = 1.f / rw[i];
        for (int k = 0; k < 4; k++) {
            float v = rr[i * 4 + k] * s;
            rr[i * 4 + k] = v > (float)INT32_MAX ? INT32_MAX : (int)v;
        }
    }
    int out = 0;
    for (int i = 0; i < nclasses; i++) { /* :207-242 */
        const int32_t *r1 = rr + 4 * i;
        int n1 = level_weights ? rej[i] : rw[i], j; /* :210 */
        if (n1 <= group_threshold) continue;
        for (j = 0; j < nclasses; j++) {
            int n2 = rw[j];
            if (j == i || n2 <= group_threshold) continue;
            const int32_t *r2 = rr + 4 * j;
            int dx = (int)(r2[2] * eps), dy = (int)(r2[3] * eps);
            if (r1[0] >= r2[0] - dx && r1[1] >= r2[1] - dy &&
                r1[0] + r1[2] <= r2[0] + r2[2] + dx && r1[1] + r1[3] <= r2[1] + r2[3] + dy &&
                (n2 > (3 > n1 ? 3 : n1) || n1 < 3))
                break;
        }
        if (j == nclasses) {
            memcpy(rects + 4 * out, r1, 4 * sizeof(int32_t));
            if (weights) weights[out] = n1;
            if (level_weights) level_weights[out] = rejw[i];
            out++;
        }
    }
    free(parent); free(labels); free(cls_of_root); free(rr); free(rw); free(rej); free(rejw);
    return out;
}

int vjo_group_rectangles(int32_t *rects, int n, int group_threshold, double eps, int32_t *weights)
{
    return group_impl(rects, n, group_threshold, eps, weights, NULL);
}

int vjo_group_rectangles_roc(int32_t *rects, int n, int group_threshold, double eps, int32_t *reject_levels,
                             double *level_weights)
{
    return group_impl(rects, n, group_threshold, eps, reject_levels, level_weights);
}
