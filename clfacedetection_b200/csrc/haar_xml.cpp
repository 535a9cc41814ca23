// haar_xml.cpp -- loader for OpenCV Haar cascade files: the old format (type_id
// "opencv-haar-classifier"), i.e. the product-side replacement of cvLoad() as called at
// main.cpp:36 = icvReadHaarClassifier (tempcv.cpp:1750-2089), and the new "opencv-cascade-
// classifier" format with HAAR features (read_new_cascade below).
//
// The reference reads these through OpenCV's CvFileStorage; here a small tolerant XML
// tokenizer builds an element tree and the cascade is read from it with the same field
// names, checks and error wording (stage/tree/node indices in every message).  Tolerant
// means: everything before <opencv_storage> is skipped (seven mcs_* files open with a
// "<!-----" comment that strict XML parsers reject) and comments may contain "--".
#include <cctype>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "clfd_internal.h"

namespace clfd {

namespace {

struct Elem {
    std::string name, type_id, text;
    std::vector<std::unique_ptr<Elem>> kids;
    const Elem *child(const char *n) const {
        for (auto &k : kids)
            if (k->name == n) return k.get();
        return nullptr;
    }
};

struct Parser {
    const char *p, *end;
    std::string err;

    bool fail(const char *msg) { if (err.empty()) err = msg; return false; }

    void skip_misc() {  // whitespace, comments, processing instructions
        for (;;) {
            while (p < end && isspace((unsigned char)*p)) p++;
            if (end - p >= 4 && !memcmp(p, "<!--", 4)) {
                const char *q = p + 4;
                while (q + 3 <= end && memcmp(q, "-->", 3)) q++;
                p = (q + 3 <= end) ? q + 3 : end;
            } else if (end - p >= 2 && !memcmp(p, "<?", 2)) {
                const char *q = p + 2;
                while (q + 2 <= end && memcmp(q, "?>", 2)) q++;
                p = (q + 2 <= end) ? q + 2 : end;
            } else
                return;
        }
    }

    // parses one element starting at '<'
    bool element(Elem &e, int depth) {
        if (depth > 64) return fail("XML nesting too deep");
        if (p >= end || *p != '<') return fail("expected '<'");
        p++;
        const char *s = p;
        while (p < end && !isspace((unsigned char)*p) && *p != '>' && *p != '/') p++;
        e.name.assign(s, p);
        if (e.name.empty()) return fail("empty tag name");
        // attributes
        bool self_closed = false;
        for (;;) {
            while (p < end && isspace((unsigned char)*p)) p++;
            if (p >= end) return fail("unterminated tag");
            if (*p == '>') { p++; break; }
            if (*p == '/') {
                if (p + 1 < end && p[1] == '>') { p += 2; self_closed = true; break; }
                return fail("malformed tag");
            }
            const char *an = p;
            while (p < end && *p != '=' && !isspace((unsigned char)*p) && *p != '>') p++;
            std::string aname(an, p);
            while (p < end && isspace((unsigned char)*p)) p++;
            if (p >= end || *p != '=') return fail("attribute without value");
            p++;
            while (p < end && isspace((unsigned char)*p)) p++;
            if (p >= end || (*p != '"' && *p != '\'')) return fail("attribute value must be quoted");
            char q = *p++;
            const char *av = p;
            while (p < end && *p != q) p++;
            if (p >= end) return fail("unterminated attribute value");
            if (aname == "type_id") e.type_id.assign(av, p);
            p++;
        }
        if (self_closed) return true;
        // content
        for (;;) {
            const char *t = p;
            while (p < end && *p != '<') p++;
            e.text.append(t, p);
            if (p >= end) return fail("unterminated element");
            if (end - p >= 4 && !memcmp(p, "<!--", 4)) { skip_misc(); continue; }
            if (p + 1 < end && p[1] == '/') {
                p += 2;
                const char *cn = p;
                while (p < end && *p != '>') p++;
                if (p >= end) return fail("unterminated closing tag");
                std::string cname(cn, p);
                while (!cname.empty() && isspace((unsigned char)cname.back())) cname.pop_back();
                p++;
                if (cname != e.name) return fail("mismatched closing tag");
                return true;
            }
            e.kids.emplace_back(new Elem());
            if (!element(*e.kids.back(), depth + 1)) return false;
        }
    }
};

// CvFileStorage typing: a token is an integer iff it is [+-]digits; otherwise real.
bool tok_is_int(const std::string &t) {
    size_t i = 0;
    if (i < t.size() && (t[i] == '+' || t[i] == '-')) i++;
    if (i >= t.size()) return false;
    for (; i < t.size(); i++)
        if (!isdigit((unsigned char)t[i])) return false;
    return true;
}

std::vector<std::string> split_ws(const std::string &s) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && isspace((unsigned char)s[i])) i++;
        size_t j = i;
        while (j < s.size() && !isspace((unsigned char)s[j])) j++;
        if (j > i) out.emplace_back(s, i, j - i);
        i = j;
    }
    return out;
}

bool scalar_int(const Elem *e, int &v) {
    if (!e) return false;
    auto t = split_ws(e->text);
    if (t.size() != 1 || !tok_is_int(t[0])) return false;
    v = (int)strtol(t[0].c_str(), nullptr, 10);
    return true;
}

// reals are read as double and cast to float (tempcv.cpp:1932,1958,1995,2033,2054)
bool tok_real(const std::string &t, float &v) {
    if (t.empty()) return false;
    char *endp = nullptr;
    errno = 0;
    double d = strtod(t.c_str(), &endp);
    if (endp == t.c_str() || *endp != '\0') return false;
    v = (float)d;
    return true;
}

bool scalar_real(const Elem *e, float &v) {
    if (!e) return false;
    auto t = split_ws(e->text);
    // CV_NODE_IS_REAL: OpenCV types "1." / "-1.5e-003" as real and bare digits as int
    if (t.size() != 1 || tok_is_int(t[0])) return false;
    return tok_real(t[0], v);
}

#define FMT_FAIL(...) do { set_error(__VA_ARGS__); return CLFD_ERR_FORMAT; } while (0)

int read_cascade(const Elem &node, HostCascade &c) {
    const Elem *stages = node.child("stages");
    if (!stages || stages->kids.empty()) FMT_FAIL("Invalid stages node");
    const int n = (int)stages->kids.size();

    const Elem *size = node.child("size");
    auto sz = size ? split_ws(size->text) : std::vector<std::string>();
    if (sz.size() != 2) FMT_FAIL("size node is not a valid sequence.");
    if (!tok_is_int(sz[0]) || atoi(sz[0].c_str()) <= 0)
        FMT_FAIL("Invalid size node: width must be positive integer");
    if (!tok_is_int(sz[1]) || atoi(sz[1].c_str()) <= 0)
        FMT_FAIL("Invalid size node: height must be positive integer");
    c.win_w = atoi(sz[0].c_str());
    c.win_h = atoi(sz[1].c_str());
    c.st_child.assign(n, -1);

    for (int i = 0; i < n; i++) {
        const Elem &stage = *stages->kids[i];
        if (stage.name != "_") FMT_FAIL("Invalid stage %d", i);
        const Elem *trees = stage.child("trees");
        if (!trees || trees->kids.empty())
            FMT_FAIL("Trees node is not a valid sequence. (stage %d)", i);
        c.st_ntrees.push_back((int)trees->kids.size());
        for (int j = 0; j < (int)trees->kids.size(); j++) {
            const Elem &tree = *trees->kids[j];
            const int count = (int)tree.kids.size();
            if (count <= 0)
                FMT_FAIL("Tree node is not a valid sequence. (stage %d, tree %d)", i, j);
            c.tr_nnodes.push_back(count);
            const size_t alpha0 = c.alpha.size();
            int last_idx = 0;
            for (int k = 0; k < count; k++) {
                const Elem &nd = *tree.kids[k];
                HostNode hn;
                memset(&hn, 0, sizeof hn);
                const Elem *feature = nd.child("feature");
                if (!feature)
                    FMT_FAIL("Feature node is not a valid map. (stage %d, tree %d, node %d)", i, j, k);
                const Elem *rects = feature->child("rects");
                if (!rects || rects->kids.size() < 1 || rects->kids.size() > 3)
                    FMT_FAIL("Rects node is not a valid sequence. (stage %d, tree %d, node %d)", i, j, k);
                for (int l = 0; l < (int)rects->kids.size(); l++) {
                    auto t = split_ws(rects->kids[l]->text);
                    if (t.size() != 5)
                        FMT_FAIL("Rect %d is not a valid sequence. (stage %d, tree %d, node %d)", l, i, j, k);
                    int v[4];
                    for (int q = 0; q < 4; q++) {
                        if (!tok_is_int(t[q]))
                            FMT_FAIL("rect coordinates must be integer. (stage %d, tree %d, node %d, rect %d)", i, j, k, l);
                        v[q] = atoi(t[q].c_str());
                    }
                    if (v[0] < 0)
                        FMT_FAIL("x coordinate must be non-negative integer. (stage %d, tree %d, node %d, rect %d)", i, j, k, l);
                    if (v[1] < 0)
                        FMT_FAIL("y coordinate must be non-negative integer. (stage %d, tree %d, node %d, rect %d)", i, j, k, l);
                    if (v[2] <= 0 || v[0] + v[2] > c.win_w)
                        FMT_FAIL("width must be positive integer and (x + width) must not exceed window width. "
                                 "(stage %d, tree %d, node %d, rect %d)", i, j, k, l);
                    if (v[3] <= 0 || v[1] + v[3] > c.win_h)
                        FMT_FAIL("height must be positive integer and (y + height) must not exceed window height. "
                                 "(stage %d, tree %d, node %d, rect %d)", i, j, k, l);
                    float w;
                    if (tok_is_int(t[4]) || !tok_real(t[4], w))
                        FMT_FAIL("weight must be real number. (stage %d, tree %d, node %d, rect %d)", i, j, k, l);
                    for (int q = 0; q < 4; q++) hn.rect[l][q] = v[q];
                    hn.weight[l] = w;
                }
                int tilted;
                if (!scalar_int(feature->child("tilted"), tilted))
                    FMT_FAIL("tilted must be 0 or 1. (stage %d, tree %d, node %d)", i, j, k);
                hn.tilted = tilted != 0;
                if (!scalar_real(nd.child("threshold"), hn.threshold))
                    FMT_FAIL("threshold must be real number. (stage %d, tree %d, node %d)", i, j, k);
                for (int side = 0; side < 2; side++) {
                    const char *sname = side ? "right" : "left";
                    int &dst = side ? hn.right : hn.left;
                    const Elem *cn = nd.child(side ? "right_node" : "left_node");
                    if (cn) {
                        int idx;
                        if (!scalar_int(cn, idx) || idx <= k || idx >= count)
                            FMT_FAIL("%s node must be valid node number. (stage %d, tree %d, node %d)", sname, i, j, k);
                        dst = idx;
                    } else {
                        const Elem *cv = nd.child(side ? "right_val" : "left_val");
                        if (!cv)
                            FMT_FAIL("%s node or %s value must be specified. (stage %d, tree %d, node %d)", sname, sname, i, j, k);
                        float val;
                        if (!scalar_real(cv, val))
                            FMT_FAIL("%s value must be real number. (stage %d, tree %d, node %d)", sname, i, j, k);
                        if (last_idx >= count + 1)
                            FMT_FAIL("Tree structure is broken: too many values. (stage %d, tree %d, node %d)", i, j, k);
                        dst = -last_idx;
                        c.alpha.push_back(val);
                        last_idx++;
                    }
                }
                c.nodes.push_back(hn);
            }
            if (last_idx != count + 1)
                FMT_FAIL("Tree structure is broken: too few values. (stage %d, tree %d)", i, j);
            (void)alpha0;
        }
        float thr;
        if (!scalar_real(stage.child("stage_threshold"), thr))
            FMT_FAIL("stage threshold must be real number. (stage %d)", i);
        c.st_thr.push_back(thr);
        int parent, next;
        if (!scalar_int(stage.child("parent"), parent) || parent < -1 || parent >= n)
            FMT_FAIL("parent must be integer number. (stage %d)", i);
        if (!scalar_int(stage.child("next"), next) || next < -1 || next >= n)
            FMT_FAIL("next must be integer number. (stage %d)", i);
        c.st_parent.push_back(parent);
        c.st_next.push_back(next);
        if (parent != -1 && c.st_child[parent] == -1) c.st_child[parent] = i;  // tempcv.cpp:2080-2083
    }
    return 0;
}

// New-format files (type_id "opencv-cascade-classifier", the on-disk format next to the one the
// reference loads: SURVEY 8-f row 4; declared in tempcv.hpp:370-491, no reader in the reference
// tree).  Only stageType BOOST + featureType HAAR; stages are linear (no stage tree).  The
// content maps one to one onto the old structure:
//   internalNodes = (left, right, featureIdx, threshold) per node, children > 0 are node indices,
//   <= 0 are -leafIndex into leafValues -- the convention HostNode::left/right already use;
//   features[featureIdx] = up to 3 rects "x y w h weight" (+ <tilted>).
// Stage thresholds are stored as in the old files (the -0.0001f bias of tempcv.cpp:419 is applied
// by build_hidden), so an old-format file and its converted copy give identical detections.
int read_new_cascade(const Elem &node, HostCascade &c) {
    const Elem *st = node.child("stageType"), *ft = node.child("featureType");
    auto word = [](const Elem *e) { auto t = e ? split_ws(e->text) : std::vector<std::string>(); return t.size() == 1 ? t[0] : std::string(); };
    if (word(st) != "BOOST") FMT_FAIL("stageType must be BOOST (found '%s')", word(st).c_str());
    if (word(ft) != "HAAR")
        FMT_FAIL("featureType '%s' is not supported: this path evaluates Haar features only (LBP / HOG cascades need another evaluator)",
                 word(ft).c_str());
    if (!scalar_int(node.child("width"), c.win_w) || c.win_w <= 0) FMT_FAIL("Invalid width node: width must be positive integer");
    if (!scalar_int(node.child("height"), c.win_h) || c.win_h <= 0) FMT_FAIL("Invalid height node: height must be positive integer");
    const Elem *features = node.child("features");
    if (!features || features->kids.empty()) FMT_FAIL("Invalid features node");
    struct Feat { int rect[3][4]; float weight[3]; int n, tilted; };
    std::vector<Feat> feats(features->kids.size());
    for (size_t fi = 0; fi < feats.size(); fi++) {
        const Elem &fe = *features->kids[fi];
        Feat &F = feats[fi];
        memset(&F, 0, sizeof F);
        const Elem *rects = fe.child("rects");
        if (!rects || rects->kids.size() < 1 || rects->kids.size() > 3)
            FMT_FAIL("Rects node is not a valid sequence. (feature %zu)", fi);
        F.n = (int)rects->kids.size();
        for (int l = 0; l < F.n; l++) {
            auto t = split_ws(rects->kids[l]->text);
            if (t.size() != 5) FMT_FAIL("Rect %d is not a valid sequence. (feature %zu)", l, fi);
            for (int q = 0; q < 4; q++) {
                if (!tok_is_int(t[q])) FMT_FAIL("rect coordinates must be integer. (feature %zu, rect %d)", fi, l);
                F.rect[l][q] = atoi(t[q].c_str());
            }
            if (F.rect[l][0] < 0 || F.rect[l][1] < 0 || F.rect[l][2] <= 0 || F.rect[l][3] <= 0)
                FMT_FAIL("rect must have non-negative origin and positive size. (feature %zu, rect %d)", fi, l);
            if (!tok_real(t[4], F.weight[l])) FMT_FAIL("weight must be real number. (feature %zu, rect %d)", fi, l);
        }
        const Elem *tl = fe.child("tilted");
        if (tl && !scalar_int(tl, F.tilted)) FMT_FAIL("tilted must be 0 or 1. (feature %zu)", fi);
    }
    const Elem *stages = node.child("stages");
    if (!stages || stages->kids.empty()) FMT_FAIL("Invalid stages node");
    const int n = (int)stages->kids.size();
    for (int i = 0; i < n; i++) {
        const Elem &stage = *stages->kids[i];
        const Elem *weak = stage.child("weakClassifiers");
        if (!weak || weak->kids.empty()) FMT_FAIL("weakClassifiers node is not a valid sequence. (stage %d)", i);
        c.st_ntrees.push_back((int)weak->kids.size());
        for (int j = 0; j < (int)weak->kids.size(); j++) {
            const Elem &tree = *weak->kids[j];
            auto in = tree.child("internalNodes") ? split_ws(tree.child("internalNodes")->text) : std::vector<std::string>();
            auto lv = tree.child("leafValues") ? split_ws(tree.child("leafValues")->text) : std::vector<std::string>();
            if (in.empty() || in.size() % 4 != 0)
                FMT_FAIL("internalNodes must hold (left, right, feature, threshold) quadruples. (stage %d, tree %d)", i, j);
            const int count = (int)in.size() / 4;
            if ((int)lv.size() != count + 1)
                FMT_FAIL("Tree structure is broken: %zu leaf values for %d nodes. (stage %d, tree %d)", lv.size(), count, i, j);
            c.tr_nnodes.push_back(count);
            for (int k = 0; k < count; k++) {
                HostNode hn;
                memset(&hn, 0, sizeof hn);
                if (!tok_is_int(in[4 * k]) || !tok_is_int(in[4 * k + 1]) || !tok_is_int(in[4 * k + 2]))
                    FMT_FAIL("node links and feature index must be integer. (stage %d, tree %d, node %d)", i, j, k);
                hn.left = atoi(in[4 * k].c_str());
                hn.right = atoi(in[4 * k + 1].c_str());
                const int fi = atoi(in[4 * k + 2].c_str());
                for (int side = 0; side < 2; side++) {
                    const int v = side ? hn.right : hn.left;
                    if (v > 0 ? (v <= k || v >= count) : -v > count)
                        FMT_FAIL("%s node must be valid node number. (stage %d, tree %d, node %d)", side ? "right" : "left", i, j, k);
                }
                if (fi < 0 || fi >= (int)feats.size())
                    FMT_FAIL("feature index %d outside 0..%zu. (stage %d, tree %d, node %d)", fi, feats.size() - 1, i, j, k);
                if (!tok_real(in[4 * k + 3], hn.threshold))
                    FMT_FAIL("threshold must be real number. (stage %d, tree %d, node %d)", i, j, k);
                const Feat &F = feats[fi];
                hn.tilted = F.tilted != 0;
                for (int l = 0; l < F.n; l++) {
                    if (F.rect[l][0] + F.rect[l][2] > c.win_w || F.rect[l][1] + F.rect[l][3] > c.win_h)
                        FMT_FAIL("rect %d of feature %d does not fit the %dx%d window. (stage %d, tree %d, node %d)", l, fi,
                                 c.win_w, c.win_h, i, j, k);
                    for (int q = 0; q < 4; q++) hn.rect[l][q] = F.rect[l][q];
                    hn.weight[l] = F.weight[l];
                }
                c.nodes.push_back(hn);
            }
            for (auto &t : lv) {
                float v;
                if (!tok_real(t, v)) FMT_FAIL("leaf value must be real number. (stage %d, tree %d)", i, j);
                c.alpha.push_back(v);
            }
        }
        float thr;
        const Elem *te = stage.child("stageThreshold");
        auto tt = te ? split_ws(te->text) : std::vector<std::string>();
        if (tt.size() != 1 || !tok_real(tt[0], thr)) FMT_FAIL("stage threshold must be real number. (stage %d)", i);
        c.st_thr.push_back(thr);
        c.st_parent.push_back(i - 1);   // linear cascade
        c.st_next.push_back(-1);
    }
    c.st_child.assign(n, -1);
    for (int i = 0; i + 1 < n; i++) c.st_child[i] = i + 1;
    return 0;
}

}  // namespace

int load_cascade_xml(const char *path, HostCascade &out) {
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("cannot open cascade file '%s': %s", path, strerror(errno)); return CLFD_ERR_IO; }
    std::string buf;
    char tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, got);
    fclose(f);
    size_t start = buf.find("<opencv_storage>");
    if (start == std::string::npos) FMT_FAIL("'%s': no <opencv_storage> element", path);
    Parser ps{buf.data() + start, buf.data() + buf.size(), {}};
    Elem root;
    if (!ps.element(root, 0)) FMT_FAIL("'%s': XML error: %s", path, ps.err.c_str());
    const Elem *node = nullptr, *node_new = nullptr;
    for (auto &k : root.kids) {
        if (k->type_id == "opencv-haar-classifier" && !node) node = k.get();
        if (k->type_id == "opencv-cascade-classifier" && !node_new) node_new = k.get();
    }
    if (!node && !node_new)
        FMT_FAIL("'%s': no node of type opencv-haar-classifier or opencv-cascade-classifier", path);
    out = HostCascade();
    out.name = (node ? node : node_new)->name;
    int rc = node ? read_cascade(*node, out) : read_new_cascade(*node_new, out);
    if (rc) return rc;
    return build_hidden(out);
}

}  // namespace clfd
