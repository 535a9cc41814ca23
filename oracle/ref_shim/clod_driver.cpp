/*
 * clod_driver.cpp -- C ABI around the REFERENCE'S OWN CPU detector, clodDetectObjects(use_cl = FALSE)
 * (part of oracle/_ref/libtempcv_ref.so).
 *
 * TEST INFRASTRUCTURE ONLY.  The .inc files included below are produced at build time by
 * oracle/build_ref.py: verbatim line ranges of /root/reference/CLFaceDetection/clod.h (17-21, 39-47)
 * and clod.cpp (11-38 macros + list structs, 182-357 filterResult, 371-527 setupScale ..
 * precomputeWindows, 580-787 runClassifier .. runCascade, 1339-1500 clodDetectObjects).  They are never
 * committed.  Everything in THIS file is glue: the OpenCL typedefs of <CL/cl.h>, the IplImage fields
 * the driver reads, the two branches that need an OpenCL runtime (they abort), and setupImage
 * (clod.cpp:360-369 -> clifGrayscaleIntegral's CPU branch, clif.cpp:326-331: cvIntegral into a
 * CV_32SC1 and a CV_64FC1 matrix) on top of the cv2-pinned integral of oracle/vj_oracle.c.
 */
#include "cvmini.hpp"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

extern "C" {
#include "tempcv_hpp_extract.inc"
#include "../vj_oracle.h"
}

/* <CL/cl.h> */
typedef uint32_t cl_uint;
typedef int32_t cl_int;
typedef float cl_float;
typedef cl_uint cl_bool;
#define CL_SUCCESS 0
#define CL_FALSE 0
#define CL_TRUE 1

typedef struct IplImage { int width, height, widthStep; char *imageData; } IplImage;
typedef struct CLIFEnvironmentData { int unused; } CLIFEnvironmentData;
typedef struct CLODFEnvironmentData { CLIFEnvironmentData *clif; } CLODEnvironmentData;

#include "clod_h_extract.inc"

static CLODDetectObjectsResult clodDetectObjectsOpenCL(const IplImage *, const CvHaarClassifierCascade *, const CLODEnvironmentData *,
                                                       const CvSize, const CvSize, const cl_uint)
{
    fprintf(stderr, "clod_driver: the OpenCL branch needs an OpenCL runtime\n");
    abort();
}
static CLODDetectObjectsResult clodDetectObjectsBlock(const IplImage *, const CvHaarClassifierCascade *, CLIFEnvironmentData *, const CvSize,
                                                      const CvSize, const cl_uint, const cl_uint)
{
    fprintf(stderr, "clod_driver: CLOD_BLOCK_IMPLEMENTATION is out of scope\n");
    abort();
}
static void setupImage(const IplImage *src, CvMat **sum, CvMat **square_sum, cl_bool)
{
    *sum = cvCreateMat(src->height + 1, src->width + 1, CV_32SC1);
    *square_sum = cvCreateMat(src->height + 1, src->width + 1, CV_64FC1);
    vjo_integral((const uint8_t *)src->imageData, src->width, src->height, src->widthStep, (*sum)->data.i, (*square_sum)->data.db, 0);
}

#define printf(...) ((void)0) /* clod.cpp:1497 prints a newline per call */
#include "clod_cpp_extract.inc"
#undef printf

extern "C" int64_t tcv_clod_detect(void *h, const uint8_t *img, int W, int H, int stride, int min_w, int min_h, int max_w, int max_h,
                                   unsigned flags, int32_t *rects, int64_t cap)
{
    const CvHaarClassifierCascade *c = (const CvHaarClassifierCascade *)h;
    IplImage im = {W, H, stride, (char *)img};
    CLIFEnvironmentData clif = {0};
    CLODEnvironmentData env = {&clif};
    CLODDetectObjectsResult r = clodDetectObjects(&im, c, &env, cvSize(min_w, min_h), cvSize(max_w, max_h), 0, flags, CL_FALSE);
    for (int64_t i = 0; i < (int64_t)r.match_count && i < cap; i++) {
        rects[4 * i + 0] = r.matches[i].rect.x; rects[4 * i + 1] = r.matches[i].rect.y;
        rects[4 * i + 2] = r.matches[i].rect.width; rects[4 * i + 3] = r.matches[i].rect.height;
    }
    free(r.matches);
    return (int64_t)r.match_count;
}
