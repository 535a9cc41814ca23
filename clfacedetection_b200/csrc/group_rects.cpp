// group_rects.cpp -- host-side rectangle grouping (the north star keeps it on the host).
// Semantics of AgroupRectangles(rectList, weights, groupThreshold, eps)
// (tempcv.cpp:145-243) with ASimilarRects (tempcv.cpp:130-143); cv::partition (external
// OpenCV, call site tempcv.cpp:160) = connected components of the similarity relation,
// classes numbered by their first member.  Replaces the reference's own filterResult
// (clod.cpp:282-357), which is defective.
#include <cfloat>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <numeric>
#include <thread>
#include <vector>

#include "clfd_internal.h"

namespace {

struct R4 { int x, y, w, h; };

inline bool similar(const R4 &a, const R4 &b, double eps) {
    const double delta = eps * (std::min(a.w, b.w) + std::min(a.h, b.h)) * 0.5;
    return std::abs(a.x - b.x) <= delta && std::abs(a.y - b.y) <= delta &&
           std::abs(a.x + a.w - b.x - b.w) <= delta && std::abs(a.y + a.h - b.y - b.h) <= delta;
}

struct DisjointSets {
    std::vector<int> parent;
    explicit DisjointSets(int n) : parent(n) { std::iota(parent.begin(), parent.end(), 0); }
    int find(int i) {
        int r = i;
        while (parent[r] != r) r = parent[r];
        while (parent[i] != r) { int nx = parent[i]; parent[i] = r; i = nx; }
        return r;
    }
    void unite(int a, int b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
};

inline int trunc_sat(float v) { return v > (float)INT_MAX ? INT_MAX : (int)v; }

}  // namespace

// weights: out = neighbour counts; with level_weights (the ROC variant, tempcv.cpp:255-258) weights carries
// the reject levels in and the winning level of each kept class out
static int group_impl(int32_t *rects_xywh, int *n_io, int group_threshold, double eps, int32_t *weights, double *level_weights) {
    if (!n_io || *n_io < 0) { clfd::set_error("bad argument"); return CLFD_ERR_INVALID; }
    const int n = *n_io;
    if (n == 0) return 0;   // an empty list needs no storage (std::vector::data() of one is NULL); tempcv.cpp:147
    if (!rects_xywh) { clfd::set_error("bad argument"); return CLFD_ERR_INVALID; }
    if (group_threshold <= 0 || n == 0) {  // tempcv.cpp:147-157
        if (weights) for (int i = 0; i < n; i++) weights[i] = 1;
        return 0;
    }
    const R4 *in = reinterpret_cast<const R4 *>(rects_xywh);
    DisjointSets ds(n);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            if (similar(in[i], in[j], eps)) ds.unite(i, j);
    // with min-index roots, a class's root IS its first member: label order = root order
    std::vector<int> label(n, -1);
    int nclasses = 0;
    for (int i = 0; i < n; i++)
        if (ds.find(i) == i) label[i] = nclasses++;
    std::vector<long long> acc((size_t)nclasses * 4, 0);
    std::vector<int> count(nclasses, 0);
    for (int i = 0; i < n; i++) {  // :167-175 (int accumulation in the reference)
        const int c = label[ds.find(i)];
        acc[c * 4 + 0] += in[i].x; acc[c * 4 + 1] += in[i].y; acc[c * 4 + 2] += in[i].w; acc[c * 4 + 3] += in[i].h;
        count[c]++;
    }
    std::vector<int> rej_level(nclasses, 0);
    std::vector<double> rej_weight(nclasses, DBL_MIN);   // :165 (DBL_MIN, not -DBL_MAX, as written there)
    if (level_weights && weights) {   // :176-189: highest level of the class, largest stage sum at that level
        for (int i = 0; i < n; i++) {
            const int c = label[ds.find(i)];
            if (weights[i] > rej_level[c]) { rej_level[c] = weights[i]; rej_weight[c] = level_weights[i]; }
            else if (weights[i] == rej_level[c] && level_weights[i] > rej_weight[c]) rej_weight[c] = level_weights[i];
        }
    }
    std::vector<R4> mean(nclasses);
    for (int c = 0; c < nclasses; c++) {  // :191-199: float reciprocal, truncation
        const float s = 1.f / count[c];
        mean[c].x = trunc_sat((int)acc[c * 4 + 0] * s); mean[c].y = trunc_sat((int)acc[c * 4 + 1] * s);
        mean[c].w = trunc_sat((int)acc[c * 4 + 2] * s); mean[c].h = trunc_sat((int)acc[c * 4 + 3] * s);
    }
    int out = 0;
    std::vector<R4> kept;
    std::vector<int> kept_w;
    std::vector<double> kept_lw;
    for (int i = 0; i < nclasses; i++) {  // :207-242
        const R4 r1 = mean[i];
        const int n1 = level_weights ? rej_level[i] : count[i];   // :210
        if (n1 <= group_threshold) continue;
        bool nested = false;
        for (int j = 0; j < nclasses && !nested; j++) {
            const int n2 = count[j];
            if (j == i || n2 <= group_threshold) continue;
            const R4 r2 = mean[j];
            // the double product is truncated as is (tempcv.cpp:221-222)
            const double vx = r2.w * eps, vy = r2.h * eps;
            const int dx = vx > INT_MAX ? INT_MAX : (int)vx, dy = vy > INT_MAX ? INT_MAX : (int)vy;
            nested = r1.x >= r2.x - dx && r1.y >= r2.y - dy && r1.x + r1.w <= r2.x + r2.w + dx &&
                     r1.y + r1.h <= r2.y + r2.h + dy && (n2 > std::max(3, n1) || n1 < 3);
        }
        if (!nested) { kept.push_back(r1); kept_w.push_back(n1); kept_lw.push_back(rej_weight[i]); out++; }
    }
    memcpy(rects_xywh, kept.data(), (size_t)out * sizeof(R4));
    if (weights) memcpy(weights, kept_w.data(), (size_t)out * sizeof(int));
    if (level_weights) memcpy(level_weights, kept_lw.data(), (size_t)out * sizeof(double));
    *n_io = out;
    return 0;
}

extern "C" int clfd_group_rectangles(int32_t *rects_xywh, int *n_io, int group_threshold, double eps,
                                     int32_t *weights) {
    return group_impl(rects_xywh, n_io, group_threshold, eps, weights, nullptr);
}

extern "C" int clfd_group_rectangles_roc(int32_t *rects_xywh, int *n_io, int group_threshold, double eps,
                                         int32_t *reject_levels, double *level_weights) {
    if (!reject_levels || !level_weights) { clfd::set_error("bad argument"); return CLFD_ERR_INVALID; }
    return group_impl(rects_xywh, n_io, group_threshold, eps, reject_levels, level_weights);
}

// Batch form: the raw rects of a whole batch, as clfd_detect / _collect return
// them, grouped per (frame, cascade) on `n_threads` host threads.  With clfd_detect_submit /
// _collect the caller runs this for batch i while the GPU evaluates batch i+1, so the O(N^2)
// grouping of tempcv.cpp:1462-1472 leaves the critical path.  Output: grouped rects sorted by
// (frame, cascade), the neighbour count of each in `weights`.
extern "C" int clfd_group_batch(const clfd_rect *rects, int64_t n, int group_threshold, double eps, int n_threads,
                                clfd_rect *out, int32_t *weights, int64_t cap, int64_t *n_out) {
    if ((!rects && n > 0) || n < 0 || !out || !n_out) { clfd::set_error("bad argument"); return CLFD_ERR_INVALID; }
    *n_out = 0;
    if (n == 0) return 0;
    // bucket by (frame, cascade), keeping the device order inside a bucket out of the result:
    // sort each bucket by (w, y, x) so the grouping is deterministic whatever the atomics did
    std::vector<int64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        const clfd_rect &p = rects[a], &q = rects[b];
        if (p.frame != q.frame) return p.frame < q.frame;
        if (p.cascade != q.cascade) return p.cascade < q.cascade;
        if (p.w != q.w) return p.w < q.w;
        if (p.y != q.y) return p.y < q.y;
        return p.x < q.x;
    });
    struct Bucket { int64_t first, count; int frame, cascade; std::vector<int32_t> r, w; int m; };
    std::vector<Bucket> buckets;
    for (int64_t i = 0; i < n;) {
        int64_t j = i;
        const clfd_rect &p = rects[order[i]];
        while (j < n && rects[order[j]].frame == p.frame && rects[order[j]].cascade == p.cascade) j++;
        buckets.push_back(Bucket{i, j - i, p.frame, p.cascade, {}, {}, 0});
        i = j;
    }
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
        for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= buckets.size()) return;
            Bucket &B = buckets[b];
            if (B.count > INT_MAX) { failed = 1; continue; }
            B.r.resize((size_t)B.count * 4);
            B.w.assign((size_t)B.count, 1);
            for (int64_t k = 0; k < B.count; k++) {
                const clfd_rect &p = rects[order[B.first + k]];
                B.r[k * 4 + 0] = p.x; B.r[k * 4 + 1] = p.y; B.r[k * 4 + 2] = p.w; B.r[k * 4 + 3] = p.h;
            }
            B.m = (int)B.count;
            if (clfd_group_rectangles(B.r.data(), &B.m, group_threshold, eps, B.w.data())) failed = 1;
        }
    };
    const int nt = std::max(1, std::min<int>(n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency(), (int)buckets.size()));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(work);
    work();
    for (auto &t : pool) t.join();
    if (failed) { clfd::set_error("rectangle grouping failed"); return CLFD_ERR_INVALID; }
    int64_t total = 0;
    for (const Bucket &B : buckets) total += B.m;
    *n_out = total;
    if (total > cap) { clfd::set_error("grouped rect buffer too small: need %lld", (long long)total); return CLFD_ERR_CAPACITY; }
    int64_t o = 0;
    for (const Bucket &B : buckets)
        for (int k = 0; k < B.m; k++, o++) {
            out[o].x = B.r[k * 4 + 0]; out[o].y = B.r[k * 4 + 1]; out[o].w = B.r[k * 4 + 2]; out[o].h = B.r[k * 4 + 3];
            out[o].frame = B.frame; out[o].cascade = B.cascade;
            if (weights) weights[o] = B.w[k];
        }
    return 0;
}
