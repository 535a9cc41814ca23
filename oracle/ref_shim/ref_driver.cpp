/*
 * ref_driver.cpp -- C ABI around the REFERENCE'S OWN Haar code (oracle/_ref/libtempcv_ref.so).
 *
 * TEST INFRASTRUCTURE ONLY.  The two .inc files included below are produced at build time by
 * oracle/build_ref.py: verbatim line ranges of /root/reference/CLFaceDetection/tempcv.hpp
 * (60-155) and tempcv.cpp (40-1516, 1702-2089).  They are never committed; oracle/_ref/ is
 * git-ignored.  Everything in THIS file is glue: OpenCV-2.4 externals the reference links
 * from dylibs (resize / integral / colour conversion / XML persistence) and flat-array
 * accessors so that tests can compare the reference's results with oracle/vj_oracle.c.
 */
#include "cvmini.hpp"

extern "C" {
#include "tempcv_hpp_extract.inc"
}
#include "tempcv_cpp_extract.inc"

extern "C" {
#include "../vj_oracle.h"
}

#include <omp.h>

/* ====================== OpenCV externals, restated (cv2-pinned in vj_oracle.c) ============ */

void cvResize(const CvArr *_src, CvArr *_dst, int interpolation)
{
    const CvMat *src = (const CvMat *)_src;
    CvMat *dst = (CvMat *)_dst;
    if (interpolation != CV_INTER_LINEAR || CV_MAT_TYPE(src->type) != CV_8UC1 ||
        CV_MAT_TYPE(dst->type) != CV_8UC1)
        CV_Error(CV_StsUnsupportedFormat, "cvmini cvResize: 8UC1 INTER_LINEAR only");
    if (vjo_resize_linear(src->data.ptr, src->cols, src->rows, src->step, dst->data.ptr, dst->cols,
                          dst->rows, dst->step))
        CV_Error(CV_StsError, "vjo_resize_linear failed");
}

void cvIntegral(const CvArr *_img, CvArr *_sum, CvArr *_sqsum, CvArr *_tilted)
{
    const CvMat *img = (const CvMat *)_img;
    CvMat *sum = (CvMat *)_sum, *sq = (CvMat *)_sqsum, *tl = (CvMat *)_tilted;
    int w = img->cols, h = img->rows;
    if (CV_MAT_TYPE(img->type) != CV_8UC1 || sum->cols != w + 1 || sum->rows != h + 1)
        CV_Error(CV_StsUnmatchedSizes, "cvmini cvIntegral: bad arguments");
    /* vjo_integral writes tight (w+1)-pitch arrays; the drivers' headers are tight too
     * (tempcv.cpp:1292-1298 builds them with cvMat(), 1238-1244 with cvCreateMat) */
    if (sum->step != (w + 1) * 4 || (sq && sq->step != (w + 1) * 8) || (tl && tl->step != (w + 1) * 4))
        CV_Error(CV_StsUnmatchedSizes, "cvmini cvIntegral: padded rows are not supported");
    std::vector<double> sqtmp;
    double *sqp = sq ? sq->data.db : 0;
    if (!sqp) { sqtmp.resize((size_t)(w + 1) * (h + 1)); sqp = sqtmp.data(); }
    vjo_integral(img->data.ptr, w, h, img->step, sum->data.i, sqp, tl ? tl->data.i : 0);
}

void cvCvtColor(const CvArr *_src, CvArr *_dst, int code)
{
    const CvMat *src = (const CvMat *)_src;
    CvMat *dst = (CvMat *)_dst;
    if (code != CV_BGR2GRAY || CV_MAT_TYPE(src->type) != CV_8UC3 || CV_MAT_TYPE(dst->type) != CV_8UC1)
        CV_Error(CV_StsUnsupportedFormat, "cvmini cvCvtColor: BGR2GRAY 8UC3->8UC1 only");
    for (int y = 0; y < src->rows; y++) {
        const uchar *s = src->data.ptr + (size_t)y * src->step;
        uchar *d = dst->data.ptr + (size_t)y * dst->step;
        for (int x = 0; x < src->cols; x++, s += 3)
            d[x] = (uchar)((s[0] * 1868 + s[1] * 9617 + s[2] * 4899 + 8192) >> 14);
    }
}

void cvCanny(const CvArr *, CvArr *, double, double, int)
{
    CV_Error(CV_StsUnsupportedFormat, "cvmini: Canny pruning is out of scope (SURVEY section 2)");
}

/* ====================== a small tolerant XML reader for cvLoad =========================== */

namespace {

thread_local std::string g_err;

struct XmlReader {
    const char *p, *end;
    std::vector<CvSeq *> seqs;
    std::vector<CvMiniMap *> maps;

    void fail(const std::string &m) { throw CvMiniError(CV_StsError, "XML: " + m); }
    void skip_ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
    /* skips whitespace, comments, processing instructions */
    void skip_misc()
    {
        for (;;) {
            skip_ws();
            if (end - p >= 4 && !memcmp(p, "<!--", 4)) {
                const char *q = p + 4;
                while (q + 3 <= end && memcmp(q, "-->", 3)) q++;
                if (q + 3 > end) fail("unterminated comment");
                p = q + 3;
            } else if (end - p >= 2 && !memcmp(p, "<?", 2)) {
                while (p + 2 <= end && memcmp(p, "?>", 2)) p++;
                p += 2;
            } else
                return;
        }
    }
    /* at '<name attr="v"...>' : returns name, fills type_id attribute if present */
    std::string open_tag(std::string *type_id, bool *self_closed)
    {
        if (p >= end || *p != '<') fail("expected '<'");
        p++;
        const char *s = p;
        while (p < end && *p != '>' && *p != ' ' && *p != '/' && *p != '\n' && *p != '\t') p++;
        std::string name(s, p);
        *self_closed = false;
        while (p < end && *p != '>') {
            if (*p == '/') { *self_closed = true; p++; continue; }
            if (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r') { p++; continue; }
            const char *a = p;
            while (p < end && *p != '=' && *p != '>') p++;
            std::string an(a, p);
            if (p < end && *p == '=') {
                p++;
                char q = *p++;
                const char *v = p;
                while (p < end && *p != q) p++;
                if (type_id && an == "type_id") *type_id = std::string(v, p);
                p++;
            }
        }
        if (p >= end) fail("unterminated tag");
        p++;
        return name;
    }
    void close_tag(const std::string &name)
    {
        if (end - p < (ptrdiff_t)name.size() + 3 || p[0] != '<' || p[1] != '/' ||
            memcmp(p + 2, name.data(), name.size()))
            fail("expected </" + name + ">");
        p += 2 + name.size();
        skip_ws();
        if (p >= end || *p != '>') fail("bad closing tag " + name);
        p++;
    }
    /* OpenCV's scalar rule (icvXMLParseValue): a token that starts like a number is a real iff
     * a '.' or 'e' follows its leading digits, else an int */
    CvFileNode scalar(const std::string &tok)
    {
        CvFileNode n;
        memset(&n, 0, sizeof(n));
        const char *s = tok.c_str();
        char c = s[0], d = s[0] ? s[1] : 0;
        if (isdigit((unsigned char)c) || ((c == '-' || c == '+') && (isdigit((unsigned char)d) || d == '.')) ||
            (c == '.' && isalnum((unsigned char)d))) {
            const char *e = s + (c == '-' || c == '+');
            while (isdigit((unsigned char)*e)) e++;
            if (*e == '.' || *e == 'e') {
                n.tag = CV_NODE_REAL;
                n.data.f = strtod(s, 0);
            } else {
                n.tag = CV_NODE_INT;
                n.data.i = (int)strtol(s, 0, 0);
            }
        } else
            n.tag = CV_NODE_STR;
        return n;
    }
    /* content of an element whose open tag has been consumed, up to (not including) its close tag */
    CvFileNode content()
    {
        CvFileNode node;
        memset(&node, 0, sizeof(node));
        skip_misc();
        if (p < end && *p == '<' && p + 1 < end && p[1] != '/') {
            /* child elements: '_' items -> sequence, named items -> map */
            CvSeq *seq = 0;
            CvMiniMap *map = 0;
            while (p < end && *p == '<' && p[1] != '/') {
                bool sc;
                std::string name = open_tag(0, &sc);
                CvFileNode child;
                memset(&child, 0, sizeof(child));
                if (!sc) {
                    child = content();
                    close_tag(name);
                }
                if (name == "_") {
                    if (!seq) { seq = cvCreateSeq(0, sizeof(CvSeq), sizeof(CvFileNode), 0); seqs.push_back(seq); }
                    cvSeqPush(seq, &child);
                } else {
                    if (!map) { map = new CvMiniMap; maps.push_back(map); }
                    map->keys.push_back(name);
                    map->vals.push_back(child);
                }
                skip_misc();
            }
            if (seq && map) fail("element mixes list items and named items");
            if (seq) { node.tag = CV_NODE_SEQ; node.data.seq = seq; }
            else { node.tag = CV_NODE_MAP; node.data.map = map; }
            return node;
        }
        /* text: whitespace-separated scalars */
        std::vector<std::string> toks;
        while (p < end && *p != '<') {
            skip_ws();
            const char *s = p;
            while (p < end && *p != '<' && *p != ' ' && *p != '\n' && *p != '\r' && *p != '\t') p++;
            if (p > s) toks.push_back(std::string(s, p));
        }
        if (toks.size() == 1) return scalar(toks[0]);
        CvSeq *seq = cvCreateSeq(0, sizeof(CvSeq), sizeof(CvFileNode), 0);
        seqs.push_back(seq);
        for (size_t i = 0; i < toks.size(); i++) {
            CvFileNode c = scalar(toks[i]);
            cvSeqPush(seq, &c);
        }
        node.tag = CV_NODE_SEQ;
        node.data.seq = seq;
        return node;
    }
    ~XmlReader()
    {
        for (size_t i = 0; i < seqs.size(); i++) cvmini_free_seq(seqs[i]);
        for (size_t i = 0; i < maps.size(); i++) delete maps[i];
    }
};

CvHaarClassifierCascade *load_xml(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) throw CvMiniError(CV_StsError, std::string("cannot open ") + path);
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    fclose(f);
    XmlReader r;
    r.p = text.data();
    r.end = r.p + text.size();
    r.skip_misc();
    bool sc;
    std::string top = r.open_tag(0, &sc);
    if (top != "opencv_storage") r.fail("no <opencv_storage>");
    r.skip_misc();
    while (r.p < r.end && *r.p == '<' && r.p[1] != '/') {
        std::string type_id;
        std::string name = r.open_tag(&type_id, &sc);
        CvFileNode node;
        memset(&node, 0, sizeof(node));
        if (!sc) { node = r.content(); r.close_tag(name); }
        if (type_id == CV_TYPE_NAME_HAAR) {
            CvFileStorage fs;
            /* the reference's own reader, tempcv.cpp:1750-2089 */
            return (CvHaarClassifierCascade *)icvReadHaarClassifier(&fs, &node);
        }
        r.skip_misc();
    }
    r.fail("no opencv-haar-classifier object");
    return 0;
}

template <typename F> int guarded(F f)
{
    try {
        return f();
    } catch (const CvMiniError &e) {
        g_err = e.what();
        return -1;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

struct Counts { int stages, trees, nodes; };
Counts counts_of(const CvHaarClassifierCascade *c)
{
    Counts k = {c->count, 0, 0};
    for (int i = 0; i < c->count; i++) {
        k.trees += c->stage_classifier[i].count;
        for (int j = 0; j < c->stage_classifier[i].count; j++) k.nodes += c->stage_classifier[i].classifier[j].count;
    }
    return k;
}

/* integrals of a gray image in the layout the reference's drivers use (tight pitch) */
struct Integrals {
    CvMat *sum, *sq, *tilted;
    Integrals(const uint8_t *img, int w, int h, int stride, bool want_tilted)
    {
        CvMat im = cvMat(h, w, CV_8UC1, (void *)img);
        im.step = stride;
        sum = cvCreateMat(h + 1, w + 1, CV_32SC1);
        sq = cvCreateMat(h + 1, w + 1, CV_64FC1);
        tilted = want_tilted ? cvCreateMat(h + 1, w + 1, CV_32SC1) : 0;
        cvIntegral(&im, sum, sq, tilted);
    }
    ~Integrals() { cvReleaseMat(&sum); cvReleaseMat(&sq); cvReleaseMat(&tilted); }
};

}  // namespace

/* ====================== C ABI ============================================================ */
extern "C" {

const char *tcv_last_error(void) { return g_err.c_str(); }

void *tcv_cascade_load_xml(const char *path)
{
    CvHaarClassifierCascade *c = 0;
    if (guarded([&] { c = load_xml(path); return 0; })) return 0;
    return c;
}

/* same flat layout as vjo_cascade_create (oracle/vj_oracle.h) */
void *tcv_cascade_from_arrays(int win_w, int win_h, int n_stages, const int *st_ntrees, const float *st_thr,
                              const int *st_parent, const int *st_next, const int *tr_nnodes,
                              const int *nd_tilted, const int *nd_rect, const float *nd_weight,
                              const float *nd_thr, const int *nd_left, const int *nd_right, const float *alpha)
{
    CvHaarClassifierCascade *c = 0;
    if (guarded([&] {
            c = icvCreateHaarClassifierCascade(n_stages);   /* tempcv.cpp:263-283 */
            c->orig_window_size = cvSize(win_w, win_h);
            int t = 0, nd = 0, al = 0;
            for (int i = 0; i < n_stages; i++) {
                CvHaarStageClassifier *st = c->stage_classifier + i;
                st->count = st_ntrees[i];
                st->threshold = st_thr[i];
                st->parent = st_parent[i];
                st->next = st_next[i];
                st->child = -1;
                st->classifier = (CvHaarClassifier *)cvAlloc(sizeof(CvHaarClassifier) * st->count);
                for (int j = 0; j < st->count; j++, t++) {
                    CvHaarClassifier *cl = st->classifier + j;
                    int cnt = tr_nnodes[t];
                    cl->count = cnt;
                    /* one block, as icvReadHaarClassifier lays it out (tempcv.cpp:1830-1839) */
                    cl->haar_feature = (CvHaarFeature *)cvAlloc(
                        cnt * (sizeof(CvHaarFeature) + sizeof(float) + 2 * sizeof(int)) + (cnt + 1) * sizeof(float));
                    cl->threshold = (float *)(cl->haar_feature + cnt);
                    cl->left = (int *)(cl->threshold + cnt);
                    cl->right = (int *)(cl->left + cnt);
                    cl->alpha = (float *)(cl->right + cnt);
                    for (int l = 0; l < cnt; l++, nd++) {
                        cl->haar_feature[l].tilted = nd_tilted[nd];
                        for (int k = 0; k < 3; k++) {
                            const int *r = nd_rect + ((size_t)nd * 3 + k) * 4;
                            cl->haar_feature[l].rect[k].r = cvRect(r[0], r[1], r[2], r[3]);
                            cl->haar_feature[l].rect[k].weight = nd_weight[(size_t)nd * 3 + k];
                        }
                        cl->threshold[l] = nd_thr[nd];
                        cl->left[l] = nd_left[nd];
                        cl->right[l] = nd_right[nd];
                    }
                    memcpy(cl->alpha, alpha + al, (cnt + 1) * sizeof(float));
                    al += cnt + 1;
                }
                /* child links as the reader derives them (tempcv.cpp:2076-2083) */
                if (st->parent != -1 && c->stage_classifier[st->parent].child == -1)
                    c->stage_classifier[st->parent].child = i;
            }
            return 0;
        }))
        return 0;
    return c;
}

void tcv_cascade_free(void *h)
{
    CvHaarClassifierCascade *c = (CvHaarClassifierCascade *)h;
    if (c) cvReleaseHaarClassifierCascade(&c);   /* tempcv.cpp:1702-1719 */
}

int tcv_cascade_counts(const void *h, int *win_w, int *win_h, int *n_stages, int *n_trees, int *n_nodes)
{
    const CvHaarClassifierCascade *c = (const CvHaarClassifierCascade *)h;
    Counts k = counts_of(c);
    *win_w = c->orig_window_size.width;
    *win_h = c->orig_window_size.height;
    *n_stages = k.stages;
    *n_trees = k.trees;
    *n_nodes = k.nodes;
    return 0;
}

/* what icvReadHaarClassifier produced, as flat arrays (layout of vjo_cascade_create) */
int tcv_cascade_dump(const void *h, int *st_ntrees, float *st_thr, int *st_parent, int *st_next, int *st_child,
                     int *tr_nnodes, int *nd_tilted, int *nd_rect, float *nd_weight, float *nd_thr,
                     int *nd_left, int *nd_right, float *alpha)
{
    const CvHaarClassifierCascade *c = (const CvHaarClassifierCascade *)h;
    int t = 0, nd = 0, al = 0;
    for (int i = 0; i < c->count; i++) {
        const CvHaarStageClassifier *st = c->stage_classifier + i;
        st_ntrees[i] = st->count;
        st_thr[i] = st->threshold;
        st_parent[i] = st->parent;
        st_next[i] = st->next;
        st_child[i] = st->child;
        for (int j = 0; j < st->count; j++, t++) {
            const CvHaarClassifier *cl = st->classifier + j;
            tr_nnodes[t] = cl->count;
            for (int l = 0; l < cl->count; l++, nd++) {
                nd_tilted[nd] = cl->haar_feature[l].tilted;
                for (int k = 0; k < 3; k++) {
                    CvRect r = cl->haar_feature[l].rect[k].r;
                    int *o = nd_rect + ((size_t)nd * 3 + k) * 4;
                    o[0] = r.x; o[1] = r.y; o[2] = r.width; o[3] = r.height;
                    nd_weight[(size_t)nd * 3 + k] = cl->haar_feature[l].rect[k].weight;
                }
                nd_thr[nd] = cl->threshold[l];
                nd_left[nd] = cl->left[l];
                nd_right[nd] = cl->right[l];
            }
            memcpy(alpha + al, cl->alpha, (cl->count + 1) * sizeof(float));
            al += cl->count + 1;
        }
    }
    return 0;
}

/* hidden cascade after cvSetImagesForHaarClassifierCascade(scale) on a (W+1)x(H+1) dummy integral:
 * weights, rect counts, biased stage thresholds, two_rects, flags (bit0 is_tree, bit1 isStumpBased,
 * bit2 has_tilted_features), corner offsets relative to the integral origin (in elements):
 * node_corners [N][3][4] = p0..p3 */
int tcv_cascade_hid(void *h, int W, int H, double scale, float *node_weights, int *node_nrects,
                    int64_t *node_corners, float *stage_thr, int *stage_two_rects, int *flags,
                    double *inv_window_area, int64_t *eq_corners /*[4]*/)
{
    CvHaarClassifierCascade *c = (CvHaarClassifierCascade *)h;
    return guarded([&] {
        CvMat *sum = cvCreateMat(H + 1, W + 1, CV_32SC1), *sq = cvCreateMat(H + 1, W + 1, CV_64FC1),
              *tl = cvCreateMat(H + 1, W + 1, CV_32SC1);
        cvSetImagesForHaarClassifierCascade(c, sum, sq, tl, scale);   /* tempcv.cpp:549-768 */
        CvHidHaarClassifierCascade *hc = c->hid_cascade;
        *flags = (hc->is_tree ? 1 : 0) | (hc->isStumpBased ? 2 : 0) | (hc->has_tilted_features ? 4 : 0);
        *inv_window_area = hc->inv_window_area;
        eq_corners[0] = hc->p0 - sum->data.i; eq_corners[1] = hc->p1 - sum->data.i;
        eq_corners[2] = hc->p2 - sum->data.i; eq_corners[3] = hc->p3 - sum->data.i;
        int nd = 0;
        for (int i = 0; i < hc->count; i++) {
            CvHidHaarStageClassifier *st = hc->stage_classifier + i;
            stage_thr[i] = st->threshold;
            stage_two_rects[i] = st->two_rects;
            for (int j = 0; j < st->count; j++)
                for (int l = 0; l < st->classifier[j].count; l++, nd++) {
                    CvHidHaarTreeNode *n = st->classifier[j].node + l;
                    int tilted = c->stage_classifier[i].classifier[j].haar_feature[l].tilted;
                    const int *base = tilted ? tl->data.i : sum->data.i;
                    int nr = 0;
                    for (int k = 0; k < 3; k++) {
                        int64_t *o = node_corners + ((size_t)nd * 3 + k) * 4;
                        if (n->feature.rect[k].p0) {
                            nr = k + 1;
                            node_weights[(size_t)nd * 3 + k] = n->feature.rect[k].weight;
                            o[0] = n->feature.rect[k].p0 - base; o[1] = n->feature.rect[k].p1 - base;
                            o[2] = n->feature.rect[k].p2 - base; o[3] = n->feature.rect[k].p3 - base;
                        } else {
                            node_weights[(size_t)nd * 3 + k] = 0.f;
                            o[0] = o[1] = o[2] = o[3] = 0;
                        }
                    }
                    node_nrects[nd] = nr;
                }
        }
        cvReleaseMat(&sum); cvReleaseMat(&sq); cvReleaseMat(&tl);
        return 0;
    });
}

/* One pyramid level given directly (no resize): cvIntegral -> cvSetImages(scale 1) ->
 * cvRunHaarClassifierCascadeSum at every (x, y) of the ScaleImage invoker's grid
 * (tempcv.cpp:1013-1021, 1079-1083).  results: the raw return value per window
 * (1 accept, -i reject at stage i, 0 reject at stage 0 / any stage-tree reject);
 * stage_sums: the stage_sum handed back.  Returns the number of windows or -1. */
int64_t tcv_eval_level(void *h, const uint8_t *img, int w, int h_, int stride, int ystep, int32_t *results,
                       double *stage_sums, int n_threads)
{
    CvHaarClassifierCascade *c = (CvHaarClassifierCascade *)h;
    int64_t total = -1;
    guarded([&] {
        if (!c->hid_cascade) icvCreateHidHaarClassifierCascade(c);
        Integrals I(img, w, h_, stride, c->hid_cascade->has_tilted_features != 0);
        cvSetImagesForHaarClassifierCascade(c, I.sum, I.sq, I.tilted, 1.);
        int w0 = c->orig_window_size.width, h0 = c->orig_window_size.height;
        int ex = w - w0, ey = h_ - h0;   /* ssz.width, y2 with one strip */
        if (ex <= 0 || ey <= 0) { total = 0; return 0; }
        int nx = (ex + ystep - 1) / ystep, ny = (ey + ystep - 1) / ystep;
        #pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
        for (int iy = 0; iy < ny; iy++)
            for (int ix = 0; ix < nx; ix++) {
                double ss = 0;
                int r = cvRunHaarClassifierCascadeSum(c, cvPoint(ix * ystep, iy * ystep), ss, 0);
                results[(size_t)iy * nx + ix] = r;
                if (stage_sums) stage_sums[(size_t)iy * nx + ix] = ss;
            }
        total = (int64_t)nx * ny;
        return 0;
    });
    return total;
}

/* Scale-cascade formulation: cvIntegral of the frame once, cvSetImages(scale = factor), then
 * cvRunHaarClassifierCascade at (cvRound(ix*step), cvRound(iy*step)) for EVERY ix < nx, iy < ny
 * (no skip rule: the caller applies tempcv.cpp:1161 to compare with the invoker).  */
int64_t tcv_eval_scaled(void *h, const uint8_t *img, int W, int H, int stride, double factor, double step,
                        int nx, int ny, int32_t *results, int n_threads)
{
    CvHaarClassifierCascade *c = (CvHaarClassifierCascade *)h;
    int64_t total = -1;
    guarded([&] {
        if (!c->hid_cascade) icvCreateHidHaarClassifierCascade(c);
        Integrals I(img, W, H, stride, c->hid_cascade->has_tilted_features != 0);
        cvSetImagesForHaarClassifierCascade(c, I.sum, I.sq, I.tilted, factor);
        #pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : omp_get_max_threads())
        for (int iy = 0; iy < ny; iy++)
            for (int ix = 0; ix < nx; ix++)
                results[(size_t)iy * nx + ix] =
                    cvRunHaarClassifierCascade(c, cvPoint(cvRound(ix * step), cvRound(iy * step)), 0);
        total = (int64_t)nx * ny;
        return 0;
    });
    return total;
}

/* The reference's whole driver, cvHaarDetectObjectsForROC (tempcv.cpp:1188-1503), unmodified.
 * flags: CV_HAAR_SCALE_IMAGE = 2 selects REF-SI, 0 selects REF-SC.  channels 1 or 3 (BGR).
 * Outputs hold `cap` entries; returns the number of result rects (may exceed cap) or -1. */
int64_t tcv_detect(void *h, const uint8_t *img, int W, int H, int stride, int channels, double scale_factor,
                   int min_neighbors, int flags, int min_w, int min_h, int max_w, int max_h,
                   int output_reject_levels, int32_t *rects, int32_t *neighbors, int32_t *reject_levels,
                   double *level_weights, int64_t cap)
{
    CvHaarClassifierCascade *c = (CvHaarClassifierCascade *)h;
    int64_t total = -1;
    guarded([&] {
        CvMat im = cvMat(H, W, channels == 3 ? CV_8UC3 : CV_8UC1, (void *)img);
        im.step = stride;
        CvMemStorage storage;
        std::vector<int> levels;
        std::vector<double> weights;
        CvSeq *seq = cvHaarDetectObjectsForROC(&im, c, &storage, levels, weights, scale_factor, min_neighbors,
                                               flags, cvSize(min_w, min_h), cvSize(max_w, max_h),
                                               output_reject_levels != 0);
        total = seq->total;
        for (int i = 0; i < seq->total && i < cap; i++) {
            CvAvgComp *a = (CvAvgComp *)cvGetSeqElem(seq, i);
            rects[4 * i + 0] = a->rect.x; rects[4 * i + 1] = a->rect.y;
            rects[4 * i + 2] = a->rect.width; rects[4 * i + 3] = a->rect.height;
            if (neighbors) neighbors[i] = a->neighbors;
            if (output_reject_levels && reject_levels && i < (int)levels.size()) reject_levels[i] = levels[i];
            if (output_reject_levels && level_weights && i < (int)weights.size()) level_weights[i] = weights[i];
        }
        cvmini_free_seq(seq);
        return 0;
    });
    return total;
}

/* AgroupRectangles (tempcv.cpp:145-258) */
int tcv_group_rectangles(int32_t *rects, int n, int group_threshold, double eps, int32_t *weights)
{
    int out = -1;
    guarded([&] {
        std::vector<cv::Rect> v(n);
        for (int i = 0; i < n; i++) v[i] = cv::Rect(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]);
        std::vector<int> w;
        AgroupRectangles(v, w, group_threshold, eps);
        out = (int)v.size();
        for (int i = 0; i < out; i++) {
            rects[4 * i] = v[i].x; rects[4 * i + 1] = v[i].y; rects[4 * i + 2] = v[i].width; rects[4 * i + 3] = v[i].height;
            if (weights) weights[i] = w[i];
        }
        return 0;
    });
    return out;
}

int tcv_group_rectangles_roc(int32_t *rects, int n, int group_threshold, double eps, int32_t *reject_levels,
                             double *level_weights)
{
    int out = -1;
    guarded([&] {
        std::vector<cv::Rect> v(n);
        for (int i = 0; i < n; i++) v[i] = cv::Rect(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]);
        std::vector<int> lv(reject_levels, reject_levels + n);
        std::vector<double> lw(level_weights, level_weights + n);
        AgroupRectangles(v, lv, lw, group_threshold, eps);
        out = (int)v.size();
        for (int i = 0; i < out; i++) {
            rects[4 * i] = v[i].x; rects[4 * i + 1] = v[i].y; rects[4 * i + 2] = v[i].width; rects[4 * i + 3] = v[i].height;
            reject_levels[i] = lv[i];
            level_weights[i] = lw[i];
        }
        return 0;
    });
    return out;
}

}  // extern "C"
