/*
 * clif.h -- "CL integral/filter" host API, B200-native implementation.
 *
 * Same public names, argument meaning and ownership rules as the reference's clif.h
 * (CLFaceDetection/clif.h:12-73) so that code written against it -- including the reference's
 * own main.cpp -- compiles unchanged.  Behind it there is no OpenCL: every call goes through
 * the C ABI of include/clfd_b200.h to hand-written sm_100a kernels.  There is no CPU fallback:
 * the `use_opencl` flag is accepted for source compatibility and ignored (the reference's CPU
 * branches were cvCvtColor / cvIntegral from OpenCV, clif.cpp:247-251,280-285,326-335).
 *
 * Deliberate deviations from the reference, all of them bug fixes:
 *   - clifIntegral uploads the pixels, not the IplImage struct (clif.cpp:290);
 *   - the squared integral is really computed (clif.cl:118 wrote the plain sum into it);
 *   - row 0 / column 0 of both outputs are zero, widthStep is honoured, steps are in bytes.
 */
#ifndef CLFD_B200_CLIF_H
#define CLFD_B200_CLIF_H

extern "C" {
#include "CLEnvironment.h"
#include "CLDevice.h"
}
#include <stdio.h>
#include <opencv2/imgproc/imgproc.hpp>
#include <opencv/cvaux.hpp>

/* Per-kernel bookkeeping blocks.  The reference kept cl_mem handles and NDRange sizes in
 * them (clif.h:12-25); the fields are kept so that user code touching them still compiles.
 * `buffers` now hold device pointers owned by the environment. */
typedef struct CLIFBgrToGrayData {
    cl_mem buffers[2];            /* [0] BGR input, [1] gray output */
    void* ptr;                    /* host view of the last gray result */
    size_t global_size[2];
    size_t local_size[2];
} CLIFBgrToGayData;               /* (sic) the reference's typedef name, clif.h:17 */

typedef struct CLIFIntegralImageData {
    cl_mem buffers[5];            /* [0] input, [1]/[2] sum / squared sum */
    void* ptr;                    /* host copy of the int32 sum, (h+1) x (w+1) */
    void* square_ptr;             /* host copy of the uint64 squared sum */
    size_t global_size[2];
    size_t local_size[2];
} CLIFIntegralImageData;

typedef struct CLIFEnvironmentData {
    CLDeviceEnvironment environment;      /* environment.impl -> clfd_context and host buffers */
    CLIFBgrToGayData bgr_to_gray_data;
    CLIFIntegralImageData integral_image_data;
} CLIFEnvironmentData;

/* Headers over environment-owned memory (clif.cpp:308-314): `image` is CV_32SC1, and
 * `square_image` is typed CV_64FC1 but its payload is uint64 exactly as in the reference's
 * device branch (clif.cpp:305,313-314; main.cpp:69 reads it through an unsigned long*). */
typedef struct CLIFIntegralResult {
    CvMat* image;
    CvMat* square_image;
} CLIFIntegralResult;

typedef struct CLIFGrayscaleResult {
    IplImage* image;
} CLIFGrayscaleResult;

/* Environment lifecycle (clif.cpp:80-118, 226-238).  `device_index` selects the GPU. The
 * struct is malloc'd; clifReleaseEnvironment releases what it owns, the caller frees it
 * (clodReleaseEnvironment frees the one it created, clod.cpp:177-178). */
CLIFEnvironmentData* clifInitEnvironment(const cl_uint device_index);
void clifReleaseEnvironment(CLIFEnvironmentData* data);

/* Buffer lifecycle for one image shape (clif.cpp:120-224). */
void clifInitBuffers(CLIFEnvironmentData* data, const cl_uint image_width, const cl_uint image_height,
                     const cl_uint image_stride, const cl_uint image_channels);
void clifReleaseBuffers(CLIFEnvironmentData* data);

/* Computations (clif.cpp:241-374).  `source` is 8-bit, 1 or 3 channels. */
CLIFGrayscaleResult clifGrayscale(const IplImage* source, CLIFEnvironmentData* data, const cl_bool use_opencl);
CLIFIntegralResult clifIntegral(const IplImage* source, CLIFEnvironmentData* data, const cl_bool use_opencl);
CLIFIntegralResult clifGrayscaleIntegral(const IplImage* source, CLIFEnvironmentData* data, const cl_bool use_opencl);

#endif
