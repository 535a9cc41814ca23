/*
 * clod.h -- "CL object detection" host API, B200-native implementation.
 *
 * Same public names, types and signatures as the reference's clod.h
 * (CLFaceDetection/clod.h:17-81); main.cpp of the reference compiles against this header
 * unchanged.  clodDetectObjects runs the whole hot path on one B200: pyramid downscale,
 * integral / squared-integral (/ tilted) images and cascade evaluation with the semantics
 * of OpenCV's CV_HAAR_SCALE_IMAGE detector (tempcv.cpp:1257-1329), then -- on the host, as
 * in the reference -- optional rectangle grouping.  `flags` and `use_opencl` are accepted
 * and ignored, exactly like the reference's OpenCL branch ignores `flags`
 * (clod.cpp:1355-1356); there is no CPU path to fall back to.
 *
 * Additive API (not in the reference): clodSetScaleFactor (the reference hard-codes 1.1,
 * clod.cpp:831,1184,1349, which stays the default) and clodSetDetectionMode.
 */
#ifndef CLFD_B200_CLOD_H
#define CLFD_B200_CLOD_H

#include <opencv2/imgproc/imgproc.hpp>
#include <opencv2/highgui/highgui.hpp>
#include <opencv/cvaux.hpp>
#include <sys/time.h>
#include <limits.h>
#include <params.h>
#include "clif.h"

/* flag bits of the reference (clod.h:17-19); kept for source compatibility */
#define CLOD_PRECOMPUTE_FEATURES  (2 << 0)
#define CLOD_BLOCK_IMPLEMENTATION (2 << 1)
#define CLOD_PER_STAGE_ITERATIONS (2 << 2)

typedef cl_uint clod_flags;

/* wall-clock stopwatch in milliseconds used by main.cpp:57-96 (clod.h:23-36) */
typedef struct ElapseTime {
    double s;
    double e;
    struct timeval time;
    void start() { gettimeofday(&time, NULL); s = now_ms(); }
    double get() { gettimeofday(&time, NULL); e = now_ms(); return e - s; }
private:
    double now_ms() const { return (double)time.tv_sec * 1000.0 + (double)time.tv_usec / 1000.0; }
} ElapseTime;

/* one raw (or grouped) detection; weight = neighbour count after grouping, else 0 */
typedef struct CLODWeightedRect {
    CvRect rect;
    cl_float weight;
} CLODWeightedRect;

/* `matches` is malloc'd by clodDetectObjects and freed by the caller (main.cpp:183) */
typedef struct CLODDetectObjectsResult {
    CLODWeightedRect* matches;
    cl_uint match_count;
} CLODDetectObjectsResult;

typedef struct CLODDetectsObjectsData {
    cl_mem buffers[5];
    size_t global_size[1];
    size_t local_size[1];
} CLODDetectObjectsData;

typedef struct CLODFEnvironmentData {
    CLIFEnvironmentData* clif;            /* owned; released by clodReleaseEnvironment */
    CLDeviceEnvironment environment;      /* environment.impl -> detector cache */
    CLODDetectObjectsData detect_objects_data;
} CLODEnvironmentData;

/* lifecycle (clod.cpp:72-180); the caller frees the struct itself (main.cpp:128) */
CLODEnvironmentData* clodInitEnvironment(const cl_uint device_index);
void clodReleaseEnvironment(CLODFEnvironmentData* data);
void clodInitBuffers(CLODEnvironmentData* data, const CvSize* integral_image_size);
void clodReleaseBuffers(CLODEnvironmentData* data);

/* clod.cpp:1339-1500.  image: 8-bit, 1 or 3 (BGR) channels.  max_window_size 0 = unlimited,
 * min_neighbors 0 = raw rects. */
CLODDetectObjectsResult clodDetectObjects(const IplImage* image, const CvHaarClassifierCascade* cascade,
                                          const CLODEnvironmentData* data, const CvSize min_window_size,
                                          const CvSize max_window_size, const cl_uint min_neighbors,
                                          const clod_flags flags, const cl_bool use_opencl);

/* additive: pyramid scale factor (> 1) used by subsequent clodDetectObjects calls */
void clodSetScaleFactor(CLODEnvironmentData* data, double scale_factor);
/* additive: 0 (default) = image pyramid (CV_HAAR_SCALE_IMAGE, tempcv.cpp:1257-1329);
 * 1 = scaled features on one integral image, the semantics of the cvHaarDetectObjects call the
 * reference's main.cpp:145 compares against (tempcv.cpp:1330-1456) and the formulation of the
 * reference's own clodDetectObjects (clod.cpp:1176-1336). */
void clodSetDetectionMode(CLODEnvironmentData* data, int scale_cascade);

#endif
