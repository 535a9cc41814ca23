// cv_shim.cpp -- implementation of include/shim/cv_shim.h: the handful of OpenCV 2.4 C-API
// entry points the reference's main.cpp / clif / clod use, on top of the clfd C ABI.
// Pixel work on the hot path (resize, BGR->gray, integral, Haar detection) runs on the GPU;
// the rest is demo plumbing.  See SURVEY.md 8-b "shim surface".
#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "clfd_b200.h"
#include "cv_shim.h"

namespace {

[[noreturn]] void die(const char* what) {
    // error convention of the reference: CLUtil's clCheckOrExit prints and terminates
    fprintf(stderr, "clfd_b200: %s: %s\n", what, clfd_last_error());
    abort();
}
#define CHECK(call) do { if ((call) < 0) die(#call); } while (0)

std::mutex g_mu;
std::map<int, clfd_context*> g_ctx;
std::map<const CvHaarClassifierCascade*, clfd_cascade*> g_cascades;
// detector plans of cvHaarDetectObjects, keyed by cascade ID (never reused; an address can be)
typedef std::tuple<uint64_t, int, int, double, int, int, int, int, int> DetCacheKey;
std::map<DetCacheKey, clfd_detector*> g_det_cache;
uint64_t g_release_generation = 0;   // bumped by every cvReleaseHaarClassifierCascade

bool file_exists(const std::string& p) { FILE* f = fopen(p.c_str(), "rb"); if (f) fclose(f); return f != nullptr; }
std::string base_name(const std::string& p) { size_t i = p.find_last_of('/'); return i == std::string::npos ? p : p.substr(i + 1); }

std::string resolve_data_path(const char* filename) {
    std::string p(filename);
    if (file_exists(p)) return p;
    const std::string b = base_name(p);
    if (const char* d = getenv("CLFD_DATA_DIR")) { std::string q = std::string(d) + "/" + b; if (file_exists(q)) return q; }
    for (const char* d : {"data/haarcascades", "../data/haarcascades", "../../data/haarcascades"}) {
        std::string q = std::string(d) + "/" + b; if (file_exists(q)) return q;
    }
    return p;
}

inline bool is_image(const CvArr* a) { return ((const IplImage*)a)->nSize == (int)sizeof(IplImage); }

struct View { unsigned char* data; int w, h, step, channels; };
View view_of(const CvArr* a) {
    if (is_image(a)) { const IplImage* i = (const IplImage*)a; return {(unsigned char*)i->imageData, i->width, i->height, i->widthStep, i->nChannels}; }
    const CvMat* m = (const CvMat*)a;
    const int cn = ((m->type >> CV_CN_SHIFT) & 63) + 1;
    return {m->data.ptr, m->cols, m->rows, m->step, cn};
}

}  // namespace

clfd_context* cvShimContext(int device_index) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_ctx.find(device_index);
    if (it != g_ctx.end()) return it->second;
    clfd_context* ctx = nullptr;
    CHECK(clfd_context_create(device_index, &ctx));
    g_ctx[device_index] = ctx;
    return ctx;
}

clfd_cascade* cvShimCascadeHandle(const CvHaarClassifierCascade* c) {
    if (!c || (c->flags & 0xffff0000) != CV_HAAR_MAGIC_VAL) { fprintf(stderr, "clfd_b200: Invalid classifier cascade\n"); abort(); }
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cascades.find(c);
    if (it != g_cascades.end()) return it->second;
    // flatten CvHaarClassifierCascade (tempcv.hpp:70-112) into the arrays of clfd_cascade_from_arrays
    std::vector<int> st_ntrees, st_parent, st_next, tr_nnodes, nd_tilted, nd_rect, nd_left, nd_right;
    std::vector<float> st_thr, nd_weight, nd_thr, alpha;
    for (int i = 0; i < c->count; i++) {
        const CvHaarStageClassifier& s = c->stage_classifier[i];
        st_ntrees.push_back(s.count); st_thr.push_back(s.threshold);
        st_parent.push_back(s.parent); st_next.push_back(s.next);
        for (int j = 0; j < s.count; j++) {
            const CvHaarClassifier& t = s.classifier[j];
            tr_nnodes.push_back(t.count);
            for (int l = 0; l < t.count; l++) {
                const CvHaarFeature& f = t.haar_feature[l];
                nd_tilted.push_back(f.tilted);
                for (int k = 0; k < 3; k++) {
                    nd_rect.push_back(f.rect[k].r.x); nd_rect.push_back(f.rect[k].r.y);
                    nd_rect.push_back(f.rect[k].r.width); nd_rect.push_back(f.rect[k].r.height);
                    nd_weight.push_back(f.rect[k].weight);
                }
                nd_thr.push_back(t.threshold[l]);
                nd_left.push_back(t.left[l]); nd_right.push_back(t.right[l]);
            }
            for (int l = 0; l <= t.count; l++) alpha.push_back(t.alpha[l]);
        }
    }
    clfd_cascade* h = nullptr;
    CHECK(clfd_cascade_from_arrays(c->orig_window_size.width, c->orig_window_size.height, c->count, st_ntrees.data(),
                                   st_thr.data(), st_parent.data(), st_next.data(), tr_nnodes.data(), nd_tilted.data(),
                                   nd_rect.data(), nd_weight.data(), nd_thr.data(), nd_left.data(), nd_right.data(),
                                   alpha.data(), &h));
    g_cascades[c] = h;
    return h;
}

unsigned long long cvShimReleaseGeneration() { std::lock_guard<std::mutex> lock(g_mu); return g_release_generation; }

bool cvShimCascadeIdAlive(unsigned long long id) {
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto& kv : g_cascades)
        if (clfd_cascade_id(kv.second) == id) return true;
    return false;
}

// ---- cascade I/O -----------------------------------------------------------------------
void* cvLoad(const char* filename, CvMemStorage*, const char*, const char**) {
    const std::string path = resolve_data_path(filename);
    clfd_cascade* h = nullptr;
    if (clfd_cascade_load_xml(path.c_str(), &h) < 0) { fprintf(stderr, "cvLoad(%s): %s\n", filename, clfd_last_error()); return nullptr; }
    clfd_cascade_info info;
    CHECK(clfd_cascade_get_info(h, &info));
    const int S = info.n_stages, T = info.n_trees, N = info.n_nodes;
    std::vector<int> st_ntrees(S), st_parent(S), st_next(S), st_child(S), tr_nnodes(T), nd_tilted(N), nd_rect((size_t)N * 12), nd_left(N), nd_right(N);
    std::vector<float> st_thr(S), nd_weight((size_t)N * 3), nd_thr(N), alpha(N + T);
    CHECK(clfd_cascade_get_arrays(h, st_ntrees.data(), st_thr.data(), st_parent.data(), st_next.data(), st_child.data(),
                                  tr_nnodes.data(), nd_tilted.data(), nd_rect.data(), nd_weight.data(), nd_thr.data(),
                                  nd_left.data(), nd_right.data(), alpha.data()));
    // same allocation scheme as icvReadHaarClassifier (tempcv.cpp:264-282, 1805-1841)
    CvHaarClassifierCascade* c = (CvHaarClassifierCascade*)calloc(1, sizeof(*c) + S * sizeof(CvHaarStageClassifier));
    c->flags = CV_HAAR_MAGIC_VAL; c->count = S;
    c->orig_window_size = cvSize(info.win_w, info.win_h);
    c->stage_classifier = (CvHaarStageClassifier*)(c + 1);
    int t = 0, n = 0, a = 0;
    for (int i = 0; i < S; i++) {
        CvHaarStageClassifier& s = c->stage_classifier[i];
        s.count = st_ntrees[i]; s.threshold = st_thr[i];
        s.parent = st_parent[i]; s.next = st_next[i]; s.child = st_child[i];
        s.classifier = (CvHaarClassifier*)calloc(s.count, sizeof(CvHaarClassifier));
        for (int j = 0; j < s.count; j++, t++) {
            CvHaarClassifier& cl = s.classifier[j];
            const int cnt = tr_nnodes[t];
            cl.count = cnt;
            char* blk = (char*)calloc(1, cnt * (sizeof(CvHaarFeature) + sizeof(float) + 2 * sizeof(int)) + (cnt + 1) * sizeof(float));
            cl.haar_feature = (CvHaarFeature*)blk;
            cl.threshold = (float*)(cl.haar_feature + cnt);
            cl.left = (int*)(cl.threshold + cnt);
            cl.right = cl.left + cnt;
            cl.alpha = (float*)(cl.right + cnt);
            for (int l = 0; l < cnt; l++, n++) {
                cl.haar_feature[l].tilted = nd_tilted[n];
                for (int k = 0; k < 3; k++) {
                    const int* r = &nd_rect[((size_t)n * 3 + k) * 4];
                    cl.haar_feature[l].rect[k].r = cvRect(r[0], r[1], r[2], r[3]);
                    cl.haar_feature[l].rect[k].weight = nd_weight[(size_t)n * 3 + k];
                }
                cl.threshold[l] = nd_thr[n]; cl.left[l] = nd_left[n]; cl.right[l] = nd_right[n];
            }
            for (int l = 0; l <= cnt; l++) cl.alpha[l] = alpha[a++];
        }
    }
    std::lock_guard<std::mutex> lock(g_mu);
    g_cascades[c] = h;
    return c;
}

void cvReleaseHaarClassifierCascade(CvHaarClassifierCascade** pc) {
    if (!pc || !*pc) return;
    CvHaarClassifierCascade* c = *pc;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        auto it = g_cascades.find(c);
        if (it != g_cascades.end()) {
            // plans cached for this cascade go with it (they would never be hit again: ids are unique)
            const uint64_t id = clfd_cascade_id(it->second);
            for (auto d = g_det_cache.begin(); d != g_det_cache.end();) {
                if (std::get<0>(d->first) == id) { clfd_detector_destroy(d->second); d = g_det_cache.erase(d); }
                else ++d;
            }
            clfd_cascade_destroy(it->second);
            g_cascades.erase(it);
            g_release_generation++;
        }
    }
    for (int i = 0; i < c->count; i++) {
        for (int j = 0; j < c->stage_classifier[i].count; j++) free(c->stage_classifier[i].classifier[j].haar_feature);
        free(c->stage_classifier[i].classifier);
    }
    free(c);
    *pc = nullptr;
}

// ---- images / matrices -------------------------------------------------------------------
IplImage* cvCreateImageHeader(CvSize size, int depth, int channels) {
    if (depth != IPL_DEPTH_8U) { fprintf(stderr, "clfd_b200 shim: only 8-bit images\n"); abort(); }
    IplImage* im = (IplImage*)calloc(1, sizeof(IplImage));
    im->nSize = sizeof(IplImage); im->nChannels = channels; im->depth = depth;
    im->width = size.width; im->height = size.height;
    im->widthStep = (size.width * channels + 3) & ~3;   // OpenCV aligns rows to 4 bytes
    im->imageSize = im->widthStep * size.height;
    return im;
}
IplImage* cvCreateImage(CvSize size, int depth, int channels) {
    IplImage* im = cvCreateImageHeader(size, depth, channels);
    im->imageData = (char*)calloc(1, (size_t)im->imageSize + 16);
    im->owns_data = 1;
    return im;
}
void cvReleaseImageHeader(IplImage** im) { if (im && *im) { free(*im); *im = nullptr; } }
void cvReleaseImage(IplImage** im) { if (im && *im) { if ((*im)->owns_data) free((*im)->imageData); free(*im); *im = nullptr; } }

IplImage* cvLoadImage(const char* filename, int) {
    FILE* f = filename ? fopen(filename, "rb") : nullptr;
    if (f) {   // binary PGM (P5) / PPM (P6), maxval 255
        char magic[3] = {0};
        int w = 0, h = 0, maxv = 0;
        if (fscanf(f, "%2s %d %d %d", magic, &w, &h, &maxv) == 4 && (magic[1] == '5' || magic[1] == '6') && maxv == 255) {
            fgetc(f);
            const int cn = magic[1] == '6' ? 3 : 1;
            IplImage* im = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, 3);
            std::vector<unsigned char> row((size_t)w * cn);
            for (int y = 0; y < h; y++) {
                if (fread(row.data(), 1, row.size(), f) != row.size()) break;
                unsigned char* d = (unsigned char*)im->imageData + (size_t)y * im->widthStep;
                for (int x = 0; x < w; x++) {
                    if (cn == 3) { d[3 * x] = row[3 * x + 2]; d[3 * x + 1] = row[3 * x + 1]; d[3 * x + 2] = row[3 * x]; }
                    else d[3 * x] = d[3 * x + 1] = d[3 * x + 2] = row[x];
                }
            }
            fclose(f);
            return im;
        }
        fclose(f);
    }
    // deterministic synthetic frame (the reference's jobs.jpeg is not in its tree, main.cpp:48)
    IplImage* im = cvCreateImage(cvSize(640, 480), IPL_DEPTH_8U, 3);
    uint32_t seed = 0xC0FFEEu;
    for (int y = 0; y < 480; y++)
        for (int x = 0; x < 640; x++) {
            seed = seed * 1664525u + 1013904223u;
            const int base = 128 + (int)(60 * sin(x * 0.021) * cos(y * 0.017)) + (int)((seed >> 24) & 31) - 16;
            unsigned char* d = (unsigned char*)im->imageData + (size_t)y * im->widthStep + 3 * x;
            d[0] = (unsigned char)MIN(255, MAX(0, base - 8)); d[1] = (unsigned char)MIN(255, MAX(0, base)); d[2] = (unsigned char)MIN(255, MAX(0, base + 8));
        }
    return im;
}

void cvCopy(const CvArr* src, CvArr* dst, const CvArr*) {
    const View s = view_of(src), d = view_of(dst);
    if (s.w != d.w || s.h != d.h || s.channels != d.channels) { fprintf(stderr, "cvCopy: size mismatch\n"); abort(); }
    for (int y = 0; y < s.h; y++) memcpy(d.data + (size_t)y * d.step, s.data + (size_t)y * s.step, (size_t)s.w * s.channels);
}

static int elem_size(int type) {
    const int depth = type & 7, cn = ((type >> CV_CN_SHIFT) & 63) + 1;
    const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 0};
    return sz[depth] * cn;
}
CvMat* cvCreateMatHeader(int rows, int cols, int type) {
    CvMat* m = (CvMat*)calloc(1, sizeof(CvMat));
    m->type = CV_MAT_TYPE(type); m->rows = rows; m->cols = cols; m->step = cols * elem_size(type);
    return m;
}
CvMat* cvCreateMat(int rows, int cols, int type) {
    CvMat* m = cvCreateMatHeader(rows, cols, type);
    m->data.ptr = (unsigned char*)calloc(1, (size_t)m->step * rows + 16);
    m->owns_data = 1;
    return m;
}
void cvReleaseMat(CvMat** m) { if (m && *m) { if ((*m)->owns_data) free((*m)->data.ptr); free(*m); *m = nullptr; } }

// ---- pixel work on the GPU ----------------------------------------------------------------
void cvResize(const CvArr* src, CvArr* dst, int interpolation) {
    if (interpolation != CV_INTER_LINEAR) { fprintf(stderr, "cvResize shim: only CV_INTER_LINEAR\n"); abort(); }
    const View s = view_of(src), d = view_of(dst);
    if (s.channels != d.channels) { fprintf(stderr, "cvResize: channel mismatch\n"); abort(); }
    clfd_context* ctx = cvShimContext(0);
    if (s.channels == 1) { CHECK(clfd_resize(ctx, s.data, s.w, s.h, s.step, 0, d.data, d.w, d.h, d.step, 0)); return; }
    // multi-channel: the bilinear kernel is per plane
    std::vector<unsigned char> ps((size_t)s.w * s.h), pd((size_t)d.w * d.h);
    for (int c = 0; c < s.channels; c++) {
        for (int y = 0; y < s.h; y++) for (int x = 0; x < s.w; x++) ps[(size_t)y * s.w + x] = s.data[(size_t)y * s.step + x * s.channels + c];
        CHECK(clfd_resize(ctx, ps.data(), s.w, s.h, s.w, 0, pd.data(), d.w, d.h, d.w, 0));
        for (int y = 0; y < d.h; y++) for (int x = 0; x < d.w; x++) d.data[(size_t)y * d.step + x * d.channels + c] = pd[(size_t)y * d.w + x];
    }
}

void cvCvtColor(const CvArr* src, CvArr* dst, int code) {
    if (code != CV_BGR2GRAY) { fprintf(stderr, "cvCvtColor shim: only CV_BGR2GRAY\n"); abort(); }
    const View s = view_of(src), d = view_of(dst);
    CHECK(clfd_bgr_to_gray(cvShimContext(0), s.data, s.w, s.h, s.step, s.channels, 0, d.data, d.step, 0));
}

void cvIntegral(const CvArr* image, CvArr* sum, CvArr* sqsum, CvArr* tilted) {
    const View s = view_of(image);
    if (s.channels != 1) { fprintf(stderr, "cvIntegral shim: single channel only\n"); abort(); }
    const size_t n1 = (size_t)(s.w + 1) * (s.h + 1);
    std::vector<int32_t> hs(n1), ht(tilted ? n1 : 0);
    std::vector<uint64_t> hq(sqsum ? n1 : 0);
    CHECK(clfd_integral(cvShimContext(0), s.data, s.w, s.h, s.step, 0, hs.data(), sqsum ? hq.data() : nullptr,
                        tilted ? ht.data() : nullptr, 0));
    auto store = [&](CvArr* out, int kind) {
        CvMat* m = (CvMat*)out;
        for (int y = 0; y <= s.h; y++)
            for (int x = 0; x <= s.w; x++) {
                const size_t i = (size_t)y * (s.w + 1) + x;
                unsigned char* p = m->data.ptr + (size_t)y * m->step;
                if (kind == 0) ((int*)p)[x] = hs[i];
                else if (kind == 1) ((double*)p)[x] = (double)hq[i];   // CV_64F: exact below 2^53
                else ((int*)p)[x] = ht[i];
            }
    };
    store(sum, 0);
    if (sqsum) store(sqsum, 1);
    if (tilted) store(tilted, 2);
}

// flags & CV_HAAR_SCALE_IMAGE: image pyramid (tempcv.cpp:1257-1329); otherwise -- as in main.cpp:145,
// which passes 0 -- the scale-cascade path (tempcv.cpp:1330-1456).  Canny pruning / biggest-object
// / rough-search flags are not implemented (not on the hot path) and abort loudly.
CvSeq* cvHaarDetectObjects(const CvArr* image, CvHaarClassifierCascade* cascade, CvMemStorage*, double scale_factor,
                           int min_neighbors, int flags, CvSize min_size, CvSize max_size) {
    if (flags & ~CV_HAAR_SCALE_IMAGE) { fprintf(stderr, "cvHaarDetectObjects: unsupported flags 0x%x\n", flags); abort(); }
    const View s = view_of(image);
    clfd_context* ctx = cvShimContext(0);
    const clfd_cascade* cas = cvShimCascadeHandle(cascade);
    // detector plans are cached per (cascade, shape, parameters): OpenCV keeps its hidden cascade too
    const int mode = (flags & CV_HAAR_SCALE_IMAGE) ? CLFD_MODE_SCALE_IMAGE : CLFD_MODE_SCALE_CASCADE;
    // one call at a time: the plan cache and the detector it hands out are shared state (the
    // reference's API is single-threaded, SURVEY 8-b "Threading")
    static std::mutex call_mu;
    std::lock_guard<std::mutex> call_lock(call_mu);
    std::unique_lock<std::mutex> lock(g_mu);
    clfd_detector*& det = g_det_cache[DetCacheKey(clfd_cascade_id(cas), s.w, s.h, scale_factor, min_size.width, min_size.height,
                                                  max_size.width, max_size.height, mode)];
    if (!det) {
        clfd_detector_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.width = s.w; cfg.height = s.h; cfg.max_batch = 1; cfg.scale_factor = scale_factor;
        cfg.min_w = min_size.width; cfg.min_h = min_size.height; cfg.max_w = max_size.width; cfg.max_h = max_size.height;
        cfg.mode = mode;
        CHECK(clfd_detector_create(ctx, &cas, 1, &cfg, &det));
    }
    clfd_detector* const d = det;
    lock.unlock();
    static std::vector<clfd_rect> rects(1 << 20);   // guarded by call_mu
    int64_t n = 0;
    CHECK(clfd_detect_image(d, s.data, s.channels, s.step, rects.data(), (int64_t)rects.size(), &n));
    // the reference's output order: scale by scale, raster (tempcv.cpp:1079-1102, 1146-1160)
    std::sort(rects.begin(), rects.begin() + n, [](const clfd_rect& a, const clfd_rect& b) {
        return std::tie(a.w, a.h, a.y, a.x) < std::tie(b.w, b.h, b.y, b.x);
    });
    std::vector<int32_t> r4((size_t)n * 4), w(n > 0 ? n : 1, 0);
    for (int64_t i = 0; i < n; i++) { r4[4 * i] = rects[i].x; r4[4 * i + 1] = rects[i].y; r4[4 * i + 2] = rects[i].w; r4[4 * i + 3] = rects[i].h; }
    int m = (int)n;
    if (min_neighbors != 0 && n > 0) CHECK(clfd_group_rectangles(r4.data(), &m, MAX(min_neighbors, 1), 0.2, w.data()));   // tempcv.cpp:1462-1472
    CvSeq* seq = (CvSeq*)calloc(1, sizeof(CvSeq));
    seq->total = m; seq->elem_size = sizeof(CvAvgComp); seq->capacity = m;
    seq->data = (char*)calloc(m > 0 ? m : 1, sizeof(CvAvgComp));
    for (int i = 0; i < m; i++) {
        CvAvgComp* c = (CvAvgComp*)seq->data + i;
        c->rect = cvRect(r4[4 * i], r4[4 * i + 1], r4[4 * i + 2], r4[4 * i + 3]);
        c->neighbors = min_neighbors != 0 ? w[i] : 0;
    }
    return seq;   // owned by the caller's "storage" in OpenCV; here released by cvClearMemStorage-free leak-by-design demo
}

CvMemStorage* cvCreateMemStorage(int) { return (CvMemStorage*)calloc(1, sizeof(CvMemStorage)); }
void cvClearMemStorage(CvMemStorage*) {}
void cvReleaseMemStorage(CvMemStorage** s) { if (s && *s) { free(*s); *s = nullptr; } }
char* cvGetSeqElem(const CvSeq* seq, int index) { return (seq && index >= 0 && index < seq->total) ? seq->data + (size_t)index * seq->elem_size : nullptr; }

// ---- demo plumbing -------------------------------------------------------------------------
void cvRectangle(CvArr* img, CvPoint p1, CvPoint p2, CvScalar color, int thickness, int, int) {
    const View v = view_of(img);
    const int t = MAX(1, thickness);
    auto put = [&](int x, int y) {
        if (x < 0 || y < 0 || x >= v.w || y >= v.h) return;
        for (int c = 0; c < v.channels && c < 4; c++) v.data[(size_t)y * v.step + x * v.channels + c] = (unsigned char)color.val[c];
    };
    for (int k = 0; k < t; k++) {
        for (int x = p1.x; x <= p2.x; x++) { put(x, p1.y + k); put(x, p2.y - k); }
        for (int y = p1.y; y <= p2.y; y++) { put(p1.x + k, y); put(p2.x - k, y); }
    }
}
int cvNamedWindow(const char*, int) { return 0; }
void cvShowImage(const char*, const CvArr*) {}
void cvDestroyWindow(const char*) {}
int cvWaitKey(int) { return -1; }
CvCapture* cvCaptureFromCAM(int) { return nullptr; }
IplImage* cvQueryFrame(CvCapture*) { return nullptr; }
void cvReleaseCapture(CvCapture** c) { if (c) *c = nullptr; }
void cvInitFont(CvFont* f, int face, double hs, double vs, double, int, int) { if (f) { f->font_face = face; f->hscale = (float)hs; f->vscale = (float)vs; } }
void cvPutText(CvArr*, const char*, CvPoint, const CvFont*, CvScalar) {}
