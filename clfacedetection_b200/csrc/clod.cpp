// clod.cpp -- the reference's clod host API (include/clod.h) on top of the clfd C ABI.
// Mirrors the lifecycle and the detection entry point of clod.cpp:72-180,1339-1500 of the
// reference; the per-scale host loops, the per-stage kernel launches with a blocking round
// trip each (clod.cpp:1212-1321) and the CPU variants are replaced by ONE enqueue of the
// whole pyramid + cascade pipeline (clfd_detect).  Grouping stays on the host
// (clfd_group_rectangles = AgroupRectangles semantics), as the north star asks.
#include <algorithm>
#include <cstdio>
#include <map>
#include <tuple>
#include <vector>

#include "clfd_b200.h"
#include "clod.h"

namespace {

// keyed by the cascade's ID, not its address: a released cascade's address can be handed out again
struct DetKey {
    uint64_t cascade; int w, h, min_w, min_h, max_w, max_h; double sf; int mode;
    bool operator<(const DetKey& o) const {
        return std::tie(cascade, w, h, min_w, min_h, max_w, max_h, sf, mode) <
               std::tie(o.cascade, o.w, o.h, o.min_w, o.min_h, o.max_w, o.max_h, o.sf, o.mode);
    }
};

struct ClodState {
    clfd_context* ctx = nullptr;
    double scale_factor = 1.1;                       // clod.cpp:1349
    int mode = CLFD_MODE_SCALE_IMAGE;
    std::map<DetKey, clfd_detector*> detectors;      // plans are cached per (cascade, shape, limits)
    unsigned long long release_generation = 0;       // cvShimReleaseGeneration() at the last sweep
    std::vector<clfd_rect> rects;
    std::vector<unsigned char> gray;
};

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "clod: %s: %s\n", what, clfd_last_error());
    abort();
}
#define CHECK(call) do { if ((call) < 0) die(#call); } while (0)

ClodState* state(const CLODEnvironmentData* d) {
    if (!d || !d->environment.impl) { fprintf(stderr, "clod: environment not initialised\n"); abort(); }
    return (ClodState*)d->environment.impl;
}

}  // namespace

CLODEnvironmentData* clodInitEnvironment(const cl_uint device_index) {
    CLODEnvironmentData* data = (CLODEnvironmentData*)calloc(1, sizeof(CLODEnvironmentData));   // clod.cpp:75
    data->clif = clifInitEnvironment(device_index);   // the reference always passed 0 here (clod.cpp:76)
    ClodState* s = new ClodState();
    s->ctx = cvShimContext((int)device_index);
    data->environment.impl = s;
    data->environment.context = s->ctx;
    return data;
}

void clodReleaseEnvironment(CLODFEnvironmentData* data) {
    if (!data) return;
    clodReleaseBuffers(data);
    if (data->environment.impl) { delete (ClodState*)data->environment.impl; data->environment.impl = nullptr; }
    if (data->clif) { clifReleaseEnvironment(data->clif); free(data->clif); data->clif = nullptr; }   // clod.cpp:177-178
}

void clodInitBuffers(CLODEnvironmentData* data, const CvSize* integral_image_size) {
    // Device buffers depend on the cascade (window size), which is only known at detect time;
    // they are planned lazily there and cached.  Kept for source compatibility (main.cpp:55).
    (void)state(data);
    (void)integral_image_size;
}

void clodReleaseBuffers(CLODEnvironmentData* data) {
    if (!data || !data->environment.impl) return;
    ClodState* s = (ClodState*)data->environment.impl;
    for (auto& kv : s->detectors) clfd_detector_destroy(kv.second);
    s->detectors.clear();
    if (data->clif) clifReleaseBuffers(data->clif);   // clod.cpp:168
}

void clodSetScaleFactor(CLODEnvironmentData* data, double scale_factor) {
    if (!(scale_factor > 1)) { fprintf(stderr, "clod: scale factor must be > 1\n"); abort(); }
    state(data)->scale_factor = scale_factor;
}

void clodSetDetectionMode(CLODEnvironmentData* data, int scale_cascade) {
    state(data)->mode = scale_cascade ? CLFD_MODE_SCALE_CASCADE : CLFD_MODE_SCALE_IMAGE;
}

CLODDetectObjectsResult clodDetectObjects(const IplImage* image, const CvHaarClassifierCascade* cascade,
                                          const CLODEnvironmentData* data, const CvSize min_window_size,
                                          const CvSize max_window_size, const cl_uint min_neighbors,
                                          const clod_flags, const cl_bool) {
    ClodState* s = state(data);
    const clfd_cascade* cas = cvShimCascadeHandle(cascade);
    const int W = image->width, H = image->height;

    // plans of cascades released since the last call go now (cvReleaseHaarClassifierCascade)
    if (const unsigned long long gen = cvShimReleaseGeneration(); gen != s->release_generation) {
        s->release_generation = gen;
        for (auto it = s->detectors.begin(); it != s->detectors.end();) {
            if (!cvShimCascadeIdAlive(it->first.cascade)) { clfd_detector_destroy(it->second); it = s->detectors.erase(it); }
            else ++it;
        }
    }
    DetKey key{clfd_cascade_id(cas), W, H, min_window_size.width, min_window_size.height, max_window_size.width, max_window_size.height,
               s->scale_factor, s->mode};
    clfd_detector*& det = s->detectors[key];
    if (!det) {
        clfd_detector_config cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.width = W; cfg.height = H; cfg.max_batch = 1; cfg.scale_factor = s->scale_factor;
        cfg.mode = s->mode;
        cfg.min_w = key.min_w; cfg.min_h = key.min_h; cfg.max_w = key.max_w; cfg.max_h = key.max_h;   // 0 = unlimited (clod.cpp:394-397)
        CHECK(clfd_detector_create(s->ctx, &cas, 1, &cfg, &det));
    }
    if (s->rects.empty()) s->rects.resize(1 << 20);
    int64_t n = 0;
    // the reference's setupImage always converted from BGR (clod.cpp:360-369); here 1, 3 or 4 channels,
    // converted on the device on the way in
    CHECK(clfd_detect_image(det, (const uint8_t*)image->imageData, image->nChannels, image->widthStep, s->rects.data(),
                            (int64_t)s->rects.size(), &n));

    // the device appends accepted windows in no particular order; the reference emits them scale by
    // scale in raster order (tempcv.cpp:1079-1102), and AgroupRectangles numbers its classes by first
    // appearance -- so sort into that order to make the (grouped) result deterministic and the reference's
    std::sort(s->rects.begin(), s->rects.begin() + n, [](const clfd_rect& a, const clfd_rect& b) {
        return std::tie(a.w, a.h, a.y, a.x) < std::tie(b.w, b.h, b.y, b.x);
    });
    std::vector<int32_t> r4((size_t)n * 4), weights(n > 0 ? n : 1, 0);
    for (int64_t i = 0; i < n; i++) {
        r4[4 * i] = s->rects[i].x; r4[4 * i + 1] = s->rects[i].y; r4[4 * i + 2] = s->rects[i].w; r4[4 * i + 3] = s->rects[i].h;
    }
    int m = (int)n;
    if (min_neighbors != 0 && n > 0)   // clod.cpp:1325-1326 -> filterResult; semantics of tempcv.cpp:1462-1472
        CHECK(clfd_group_rectangles(r4.data(), &m, (int)MAX(min_neighbors, 1u), 0.2, weights.data()));

    CLODDetectObjectsResult result;
    result.match_count = (cl_uint)m;
    result.matches = (CLODWeightedRect*)malloc(sizeof(CLODWeightedRect) * (m > 0 ? m : 1));   // freed by the caller (main.cpp:183)
    for (int i = 0; i < m; i++) {
        result.matches[i].rect = cvRect(r4[4 * i], r4[4 * i + 1], r4[4 * i + 2], r4[4 * i + 3]);
        result.matches[i].weight = min_neighbors != 0 ? (cl_float)weights[i] : 0.f;   // clod.cpp:782
    }
    return result;
}
