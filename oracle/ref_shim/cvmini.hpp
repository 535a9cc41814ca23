/*
 * cvmini.hpp -- the few OpenCV 2.4 names the reference's tempcv.cpp touches on the Haar path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/vj_oracle.h).  This is NOT OpenCV and not reference
 * code: it is the smallest set of types, macros and helpers that lets the reference's own
 * functions
 *     tempcv.cpp:40-1516   (AgroupRectangles, icvCreateHidHaarClassifierCascade,
 *                           cvSetImagesForHaarClassifierCascade, icvEvalHidHaarClassifier,
 *                           cvRunHaarClassifierCascadeSum, both invokers,
 *                           cvHaarDetectObjectsForROC, cvHaarDetectObjects)
 *     tempcv.cpp:1702-2089 (cvReleaseHaarClassifierCascade, icvReadHaarClassifier)
 *     tempcv.hpp:60-155    (CvHaar* structs and flags)
 * compile from where they lie under /root/reference (oracle/build_ref.py extracts those line
 * ranges verbatim into oracle/_ref/ at build time; nothing of them is committed).
 *
 * What the reference leaves to the OpenCV 2.4.2 dylibs is supplied here:
 *   cvResize(INTER_LINEAR), cvIntegral  -> the cv2-pinned restatements in oracle/vj_oracle.c
 *   cvCvtColor(BGR2GRAY)                -> OpenCV's 14-bit fixed point (1868, 9617, 4899)
 *   cv::partition                       -> union-find, classes numbered by first appearance
 *   CvFileStorage / CvFileNode          -> a small tolerant XML reader (oracle/ref_shim/ref_driver.cpp)
 *   cvRound                             -> lrint (round half to even)
 */
#ifndef CLFD_ORACLE_CVMINI_HPP
#define CLFD_ORACLE_CVMINI_HPP

#include <assert.h>
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

using std::vector;

typedef unsigned char uchar;

/* ---- status codes / error ------------------------------------------------------------- */
enum {
    CV_StsError = -2, CV_StsNullPtr = -27, CV_StsBadArg = -5, CV_StsOutOfRange = -211,
    CV_BadCOI = -24, CV_StsUnmatchedSizes = -209, CV_StsUnsupportedFormat = -210
};
struct CvMiniError : std::runtime_error {
    int code;
    CvMiniError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
#define CV_Error(code, msg) throw CvMiniError((code), std::string(msg))

#define CV_IMPL extern "C"
#define CVAPI(rettype) rettype
#define CV_DEFAULT(v) = v
#define CV_INLINE static inline
#define CV_OUT
#define CV_IN_OUT

#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif
#define CV_IMIN(a, b) ((a) ^ (((a) ^ (b)) & (((a) < (b)) - 1)))

static inline int cvRound(double v) { return (int)lrint(v); }

/* ---- basic C types --------------------------------------------------------------------- */
typedef void CvArr;
typedef struct CvRect { int x, y, width, height; } CvRect;
typedef struct CvSize { int width, height; } CvSize;
typedef struct CvPoint { int x, y; } CvPoint;
static inline CvRect cvRect(int x, int y, int w, int h) { CvRect r = {x, y, w, h}; return r; }
static inline CvSize cvSize(int w, int h) { CvSize s = {w, h}; return s; }
static inline CvPoint cvPoint(int x, int y) { CvPoint p = {x, y}; return p; }

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_MAT_TYPE_MASK 0xFFF
#define CV_MAT_TYPE(flags) ((flags) & CV_MAT_TYPE_MASK)
#define CV_MAT_DEPTH(flags) ((flags) & 7)
#define CV_MAT_CN(flags) ((((flags) & (511 << CV_CN_SHIFT)) >> CV_CN_SHIFT) + 1)
#define CV_MAGIC_MASK 0xFFFF0000

typedef struct CvMat {
    int type;
    int step;
    int *refcount;
    int hdr_refcount;
    union { uchar *ptr; short *s; int *i; float *fl; double *db; } data;
    union { int rows; int height; };
    union { int cols; int width; };
} CvMat;

static inline int cvmini_elem_size(int type)
{
    static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 0};
    return sz[CV_MAT_DEPTH(type)] * CV_MAT_CN(type);
}
static inline CvMat cvMat(int rows, int cols, int type, void *data)
{
    CvMat m;
    memset(&m, 0, sizeof(m));
    m.type = CV_MAT_TYPE(type);
    m.rows = rows;
    m.cols = cols;
    m.step = cols * cvmini_elem_size(type);
    m.data.ptr = (uchar *)data;
    return m;
}
static inline CvMat *cvCreateMat(int rows, int cols, int type)
{
    CvMat *m = (CvMat *)malloc(sizeof(CvMat));
    *m = cvMat(rows, cols, type, 0);
    m->data.ptr = (uchar *)calloc((size_t)rows * m->step + 64, 1);
    m->hdr_refcount = 1;
    return m;
}
static inline void cvReleaseMat(CvMat **m)
{
    if (m && *m) {
        if ((*m)->hdr_refcount) free((*m)->data.ptr);
        free(*m);
        *m = 0;
    }
}
static inline CvMat *cvGetMat(const CvArr *arr, CvMat *, int *coi = 0, int = 0)
{
    if (coi) *coi = 0;
    return (CvMat *)arr;   /* every caller here passes a CvMat */
}
#define CV_ARE_SIZES_EQ(a, b) ((a)->rows == (b)->rows && (a)->cols == (b)->cols)
#define CV_MAT_ELEM_PTR_FAST(mat, row, col, pix_size) \
    ((mat).data.ptr + (size_t)(mat).step * (row) + (pix_size) * (col))
static inline void cvZero(CvArr *arr)
{
    CvMat *m = (CvMat *)arr;
    memset(m->data.ptr, 0, (size_t)m->rows * m->step);
}

static inline void *cvAlloc(size_t n) { return malloc(n ? n : 1); }
static inline void cvFree_(void *p) { free(p); }
#define cvFree(ptr) (cvFree_(*(ptr)), *(ptr) = 0)
static inline void *cvAlignPtr(const void *ptr, int align = 32)
{
    return (void *)(((size_t)ptr + align - 1) & ~(size_t)(align - 1));
}

/* ---- sequences (CvSeq of fixed-size elements, contiguous) ------------------------------- */
typedef struct CvMemStorage { int unused; } CvMemStorage;
typedef struct CvSeq {
    int total;
    int elem_size;
    int cap;
    char *data;
} CvSeq;
typedef struct CvSeqReader { char *ptr; } CvSeqReader;
static inline CvSeq *cvCreateSeq(int, size_t, size_t elem_size, CvMemStorage *)
{
    CvSeq *s = (CvSeq *)calloc(1, sizeof(CvSeq));
    s->elem_size = (int)elem_size;
    return s;
}
static inline char *cvSeqPush(CvSeq *s, const void *elem)
{
    if (s->total == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 16;
        s->data = (char *)realloc(s->data, (size_t)s->cap * s->elem_size);
    }
    char *dst = s->data + (size_t)s->total++ * s->elem_size;
    if (elem) memcpy(dst, elem, s->elem_size);
    return dst;
}
static inline char *cvGetSeqElem(const CvSeq *s, int idx) { return s->data + (size_t)idx * s->elem_size; }
static inline void cvStartReadSeq(const CvSeq *s, CvSeqReader *r, int = 0) { r->ptr = s->data; }
#define CV_NEXT_SEQ_ELEM(elem_size, reader) ((reader).ptr += (elem_size))
#define CV_SEQ_ELEM(seq, elem_type, index) ((elem_type *)cvGetSeqElem((seq), (index)))
static inline void cvmini_free_seq(CvSeq *s) { if (s) { free(s->data); free(s); } }

/* ---- file nodes (what cvLoad's XML parser hands to icvReadHaarClassifier) --------------- */
#define CV_NODE_NONE 0
#define CV_NODE_INT 1
#define CV_NODE_REAL 2
#define CV_NODE_STR 3
#define CV_NODE_SEQ 5
#define CV_NODE_MAP 6
#define CV_NODE_TYPE_MASK 7
#define CV_NODE_TYPE(flags) ((flags) & CV_NODE_TYPE_MASK)
#define CV_NODE_IS_INT(flags) (CV_NODE_TYPE(flags) == CV_NODE_INT)
#define CV_NODE_IS_REAL(flags) (CV_NODE_TYPE(flags) == CV_NODE_REAL)
#define CV_NODE_IS_SEQ(flags) (CV_NODE_TYPE(flags) == CV_NODE_SEQ)
#define CV_NODE_IS_MAP(flags) (CV_NODE_TYPE(flags) == CV_NODE_MAP)
struct CvMiniMap;
typedef struct CvFileNode {
    int tag;
    union { double f; int i; CvSeq *seq; CvMiniMap *map; } data;
} CvFileNode;
struct CvMiniMap {
    std::vector<std::string> keys;
    std::vector<CvFileNode> vals;
};
typedef struct CvFileStorage { CvFileNode root; std::string root_name, type_id; } CvFileStorage;
typedef struct CvAttrList { const char **attr; struct CvAttrList *next; } CvAttrList;
static inline CvFileNode *cvGetFileNodeByName(const CvFileStorage *, const CvFileNode *map, const char *name)
{
    if (!map || !CV_NODE_IS_MAP(map->tag)) return 0;
    CvMiniMap *m = map->data.map;
    for (size_t i = 0; i < m->keys.size(); i++)
        if (m->keys[i] == name) return &m->vals[i];
    return 0;
}

/* ---- imgproc entry points the drivers call (bodies in ref_driver.cpp) ------------------- */
#define CV_INTER_LINEAR 1
#define CV_BGR2GRAY 6
void cvResize(const CvArr *src, CvArr *dst, int interpolation);
void cvIntegral(const CvArr *image, CvArr *sum, CvArr *sqsum = 0, CvArr *tilted_sum = 0);
void cvCvtColor(const CvArr *src, CvArr *dst, int code);
void cvCanny(const CvArr *image, CvArr *edges, double t1, double t2, int aperture);

/* ---- the C++ names ------------------------------------------------------------------------ */
namespace cv {
using std::max;
using std::min;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    Size(const CvSize &s) : width(s.width), height(s.height) {}
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int _w, int _h) : x(_x), y(_y), width(_w), height(_h) {}
    Rect(const CvRect &r) : x(r.x), y(r.y), width(r.width), height(r.height) {}
    int area() const { return width * height; }
    operator CvRect() const { return cvRect(x, y, width, height); }
};
struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
};
struct BlockedRange {
    int b, e;
    BlockedRange(int _b, int _e) : b(_b), e(_e) {}
    int begin() const { return b; }
    int end() const { return e; }
};
template <typename Body> static inline void parallel_for(const BlockedRange &r, const Body &body) { body(r); }
typedef std::vector<Rect> ConcurrentRectVector;

struct Mat {
    int rows, cols;
    Mat() : rows(0), cols(0) {}
    Mat(const CvMat *m) : rows(m->rows), cols(m->cols) {}
};

static inline void cvmini_release(CvMat *p) { cvReleaseMat(&p); }
static inline void cvmini_release(CvMemStorage *p) { free(p); }
template <typename T> struct Ptr {
    T *obj;
    Ptr() : obj(0) {}
    Ptr(T *p) : obj(p) {}
    ~Ptr() { if (obj) cvmini_release(obj); }
    Ptr &operator=(T *p) { if (obj && obj != p) cvmini_release(obj); obj = p; return *this; }
    T *operator->() { return obj; }
    const T *operator->() const { return obj; }
    operator T *() { return obj; }
    operator const T *() const { return obj; }
private:
    Ptr(const Ptr &);
    Ptr &operator=(const Ptr &);
};

/* cv::partition (OpenCV core operations.hpp, external): equivalence classes of the transitive
 * closure of `predicate` over all ordered pairs; classes numbered in order of first appearance */
template <typename T, class EqPredicate>
int partition(const std::vector<T> &vec, std::vector<int> &labels, EqPredicate predicate)
{
    int n = (int)vec.size();
    std::vector<int> parent(n), rank(n, 0);
    for (int i = 0; i < n; i++) parent[i] = i;
    struct F {
        static int find(std::vector<int> &p, int i)
        {
            int r = i;
            while (p[r] != r) r = p[r];
            while (p[i] != r) { int nx = p[i]; p[i] = r; i = nx; }
            return r;
        }
    };
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (i == j || !predicate(vec[i], vec[j])) continue;
            int ri = F::find(parent, i), rj = F::find(parent, j);
            if (ri == rj) continue;
            if (rank[ri] < rank[rj]) std::swap(ri, rj);
            parent[rj] = ri;
            if (rank[ri] == rank[rj]) rank[ri]++;
        }
    labels.assign(n, 0);
    std::vector<int> cls(n, -1);
    int nclasses = 0;
    for (int i = 0; i < n; i++) {
        int r = F::find(parent, i);
        if (cls[r] < 0) cls[r] = nclasses++;
        labels[i] = cls[r];
    }
    return nclasses;
}
}  // namespace cv

#endif
