/*
 * cv_shim.h -- the small part of OpenCV 2.4's C API and of the author's CLUtil library that
 * the reference's clif.h / clod.h / main.cpp touch, so that they compile and link without
 * OpenCV, CLUtil or an OpenCL runtime (none of which exist in this image).
 *
 * This is NOT a port of OpenCV: types carry only the members the reference dereferences
 * (main.cpp, clif.cpp, clod.cpp), and the functions are implemented in cv_shim.cpp on top of
 * the clfd C ABI (GPU) where they do pixel work on the hot path, or as trivial host code /
 * no-ops where they are demo plumbing (windows, drawing).  Type layouts follow
 * opencv2/core/types_c.h and tempcv.hpp:70-118 closely enough for source compatibility.
 */
#ifndef CLFD_B200_CV_SHIM_H
#define CLFD_B200_CV_SHIM_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- CLUtil / OpenCL scalar types (CLEnvironment.h, CLDevice.h, cl_platform.h) ---- */
typedef uint32_t cl_uint;
typedef int32_t cl_int;
typedef float cl_float;
typedef double cl_double;
typedef uint64_t cl_ulong;
typedef uint8_t cl_uchar;
typedef cl_uint cl_bool;
typedef void* cl_mem;          /* opaque: device pointer owned by the environment */
#ifndef CL_TRUE
#define CL_TRUE 1
#define CL_FALSE 0
#endif

typedef struct CLDeviceEnvironment {   /* CLUtil's {context, queue, kernels[]} */
    void* context;
    void* queue;
    void** kernels;
    cl_uint kernels_count;
    void* impl;                        /* implementation state of this library */
} CLDeviceEnvironment;

/* ---- OpenCV core types ---- */
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif

#define IPL_DEPTH_8U 8
#define CV_8U 0
#define CV_32S 4
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_MAT_TYPE(t) ((t) & 0xfff)
#define CV_BGR2GRAY 6
#define CV_INTER_LINEAR 1
#define CV_HAAR_SCALE_IMAGE 2
#define CV_LOAD_IMAGE_COLOR 1
#define CV_FONT_HERSHEY_PLAIN 1
#define CV_RGB(r, g, b) cvScalar((b), (g), (r), 0)

typedef struct CvSize { int width, height; } CvSize;
typedef struct CvPoint { int x, y; } CvPoint;
typedef struct CvRect { int x, y, width, height; } CvRect;
typedef struct CvScalar { double val[4]; } CvScalar;
typedef struct CvFont { int font_face; float hscale, vscale; } CvFont;

typedef struct IplImage {
    int nSize;
    int nChannels;
    int depth;
    int width, height;
    int widthStep;
    int imageSize;
    char* imageData;
    int owns_data;
} IplImage;

typedef struct CvMat {
    int type;
    int step;          /* bytes per row */
    int* refcount;
    union { unsigned char* ptr; short* s; int* i; float* fl; double* db; } data;
    union { int rows; int height; };
    union { int cols; int width; };
    int owns_data;
} CvMat;

typedef struct CvMemStorage { int unused; } CvMemStorage;
typedef struct CvSeq { int total; int elem_size; char* data; int capacity; } CvSeq;
typedef struct CvCapture { int unused; } CvCapture;
typedef void CvArr;

/* Haar cascade structures (tempcv.hpp:70-118) */
#define CV_HAAR_FEATURE_MAX 3
#define CV_HAAR_MAGIC_VAL 0x42500000
typedef struct CvHaarFeature {
    int tilted;
    struct { CvRect r; float weight; } rect[CV_HAAR_FEATURE_MAX];
} CvHaarFeature;
typedef struct CvHaarClassifier {
    int count;
    CvHaarFeature* haar_feature;
    float* threshold;
    int* left;
    int* right;
    float* alpha;
} CvHaarClassifier;
typedef struct CvHaarStageClassifier {
    int count;
    float threshold;
    CvHaarClassifier* classifier;
    int next, child, parent;
} CvHaarStageClassifier;
typedef struct CvHidHaarClassifierCascade CvHidHaarClassifierCascade;
typedef struct CvHaarClassifierCascade {
    int flags;
    int count;
    CvSize orig_window_size;
    CvSize real_window_size;
    double scale;
    CvHaarStageClassifier* stage_classifier;
    CvHidHaarClassifierCascade* hid_cascade;   /* here: cached clfd_cascade handle */
} CvHaarClassifierCascade;
typedef struct CvAvgComp { CvRect rect; int neighbors; } CvAvgComp;

/* ---- inline helpers ---- */
extern "C++" {
static inline CvSize cvSize(int w, int h) { CvSize s; s.width = w; s.height = h; return s; }
static inline CvPoint cvPoint(int x, int y) { CvPoint p; p.x = x; p.y = y; return p; }
static inline CvRect cvRect(int x, int y, int w, int h) { CvRect r; r.x = x; r.y = y; r.width = w; r.height = h; return r; }
static inline CvScalar cvScalar(double a, double b, double c, double d) { CvScalar s; s.val[0] = a; s.val[1] = b; s.val[2] = c; s.val[3] = d; return s; }
static inline int cvRound(double v) { return (int)lrint(v); }
}

/* ---- functions (cv_shim.cpp); C++ linkage even when this header is pulled in from inside an
 * extern "C" block, as clif.h does for the CLUtil headers (clif.h:4-7) ---- */
extern "C++" {
/* cascade I/O: cvLoad dispatches on the file's type_id; only opencv-haar-classifier is known.
 * A path that does not exist is retried as $CLFD_DATA_DIR/<basename>, then as
 * <repo>/data/haarcascades/<basename>, so the reference's absolute macOS paths keep working. */
void* cvLoad(const char* filename, CvMemStorage* storage = 0, const char* name = 0, const char** real_name = 0);
void cvReleaseHaarClassifierCascade(CvHaarClassifierCascade** cascade);

IplImage* cvCreateImage(CvSize size, int depth, int channels);
IplImage* cvCreateImageHeader(CvSize size, int depth, int channels);
void cvReleaseImage(IplImage** image);
void cvReleaseImageHeader(IplImage** image);
/* reads binary PGM/PPM; any other / missing file yields a deterministic synthetic 640x480 BGR frame */
IplImage* cvLoadImage(const char* filename, int iscolor = CV_LOAD_IMAGE_COLOR);
void cvCopy(const CvArr* src, CvArr* dst, const CvArr* mask = 0);
#define cvCopyImage(src, dst) cvCopy((src), (dst), 0)

CvMat* cvCreateMat(int rows, int cols, int type);
CvMat* cvCreateMatHeader(int rows, int cols, int type);
void cvReleaseMat(CvMat** mat);

/* pixel work: executed on the GPU through the clfd C ABI */
void cvResize(const CvArr* src, CvArr* dst, int interpolation = CV_INTER_LINEAR);
void cvCvtColor(const CvArr* src, CvArr* dst, int code);
void cvIntegral(const CvArr* image, CvArr* sum, CvArr* sqsum = 0, CvArr* tilted_sum = 0);
CvSeq* cvHaarDetectObjects(const CvArr* image, CvHaarClassifierCascade* cascade, CvMemStorage* storage,
                           double scale_factor = 1.1, int min_neighbors = 3, int flags = 0,
                           CvSize min_size = cvSize(0, 0), CvSize max_size = cvSize(0, 0));

CvMemStorage* cvCreateMemStorage(int block_size = 0);
void cvClearMemStorage(CvMemStorage* storage);
void cvReleaseMemStorage(CvMemStorage** storage);
char* cvGetSeqElem(const CvSeq* seq, int index);

/* demo plumbing: drawing is real (so results can be inspected), windows / camera are no-ops */
void cvRectangle(CvArr* img, CvPoint pt1, CvPoint pt2, CvScalar color, int thickness = 1, int line_type = 8, int shift = 0);
int cvNamedWindow(const char* name, int flags = 1);
void cvShowImage(const char* name, const CvArr* image);
void cvDestroyWindow(const char* name);
int cvWaitKey(int delay = 0);
CvCapture* cvCaptureFromCAM(int index);
IplImage* cvQueryFrame(CvCapture* capture);
void cvReleaseCapture(CvCapture** capture);
void cvInitFont(CvFont* font, int font_face, double hscale, double vscale, double shear = 0, int thickness = 1, int line_type = 8);
void cvPutText(CvArr* img, const char* text, CvPoint org, const CvFont* font, CvScalar color);

/* access for the clif / clod layers */
struct clfd_context;
struct clfd_cascade;
clfd_context* cvShimContext(int device_index);                     /* lazily created, one per device */
clfd_cascade* cvShimCascadeHandle(const CvHaarClassifierCascade* cascade);   /* packed once, cached */
unsigned long long cvShimReleaseGeneration();            /* changes whenever a cascade is released */
bool cvShimCascadeIdAlive(unsigned long long cascade_id); /* is a cascade with this clfd_cascade_id still loaded */
}  /* extern "C++" */

#endif
