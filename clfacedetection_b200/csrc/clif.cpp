// clif.cpp -- the reference's clif host API (include/clif.h) on top of the clfd C ABI.
// Mirrors clif.cpp:80-374 of the reference: environment + buffer lifecycle, clifGrayscale,
// clifIntegral, clifGrayscaleIntegral.  All pixel work happens in sm_100a kernels; results
// are CvMat / IplImage headers over environment-owned host memory, as in the reference's
// device branch (clif.cpp:302-314).  Errors are fatal (the reference's clCheckOrExit).
#include <cstdio>
#include <vector>

#include "clfd_b200.h"
#include "clif.h"

namespace {

struct ClifState {
    clfd_context* ctx = nullptr;
    int device = 0;
    cl_uint width = 0, height = 0, stride = 0, channels = 0;
    std::vector<int32_t> sum;
    std::vector<uint64_t> sqsum;
    std::vector<unsigned char> gray;
    CvMat* sum_hdr = nullptr;
    CvMat* sq_hdr = nullptr;
    IplImage* gray_hdr = nullptr;
};

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "clif: %s: %s\n", what, clfd_last_error());
    abort();
}
#define CHECK(call) do { if ((call) < 0) die(#call); } while (0)

ClifState* state(CLIFEnvironmentData* d) {
    if (!d || !d->environment.impl) { fprintf(stderr, "clif: environment not initialised\n"); abort(); }
    return (ClifState*)d->environment.impl;
}

void ensure_headers(ClifState* s, int w, int h) {
    const size_t n1 = (size_t)(w + 1) * (h + 1);
    if (s->sum.size() != n1) { s->sum.assign(n1, 0); s->sqsum.assign(n1, 0); }
    if (!s->sum_hdr) s->sum_hdr = cvCreateMatHeader(h + 1, w + 1, CV_32SC1);
    if (!s->sq_hdr) s->sq_hdr = cvCreateMatHeader(h + 1, w + 1, CV_64FC1);
    s->sum_hdr->rows = s->sq_hdr->rows = h + 1;
    s->sum_hdr->cols = s->sq_hdr->cols = w + 1;
    s->sum_hdr->step = (w + 1) * 4;          // bytes (the reference set elements, clif.cpp:370,372)
    s->sq_hdr->step = (w + 1) * 8;
    s->sum_hdr->data.i = s->sum.data();
    s->sq_hdr->data.db = (double*)s->sqsum.data();   // uint64 payload, as clif.cpp:305,313-314
}

// gray plane of `source` on the host: as is for 1 channel, BGR->gray kernel for 3/4
const unsigned char* gray_plane(ClifState* s, const IplImage* src, int* step) {
    if (src->nChannels == 1) { *step = src->widthStep; return (const unsigned char*)src->imageData; }
    s->gray.resize((size_t)src->width * src->height);
    CHECK(clfd_bgr_to_gray(s->ctx, (const uint8_t*)src->imageData, src->width, src->height, src->widthStep,
                           src->nChannels, 0, s->gray.data(), src->width, 0));
    *step = src->width;
    return s->gray.data();
}

}  // namespace

CLIFEnvironmentData* clifInitEnvironment(const cl_uint device_index) {
    CLIFEnvironmentData* data = (CLIFEnvironmentData*)calloc(1, sizeof(CLIFEnvironmentData));   // clif.cpp:83
    ClifState* s = new ClifState();
    s->device = (int)device_index;
    s->ctx = cvShimContext((int)device_index);
    data->environment.impl = s;
    data->environment.context = s->ctx;
    return data;
}

void clifReleaseEnvironment(CLIFEnvironmentData* data) {
    if (!data || !data->environment.impl) return;
    clifReleaseBuffers(data);
    delete (ClifState*)data->environment.impl;   // the context is shared per device and stays alive
    data->environment.impl = nullptr;
}

void clifInitBuffers(CLIFEnvironmentData* data, const cl_uint image_width, const cl_uint image_height,
                     const cl_uint image_stride, const cl_uint image_channels) {
    ClifState* s = state(data);
    s->width = image_width; s->height = image_height; s->stride = image_stride; s->channels = image_channels;
    ensure_headers(s, (int)image_width, (int)image_height);
    data->integral_image_data.ptr = s->sum.data();
    data->integral_image_data.square_ptr = s->sqsum.data();
    data->integral_image_data.global_size[0] = image_height;   // informational, clif.cpp:182-185
    data->integral_image_data.global_size[1] = image_width;
    data->bgr_to_gray_data.global_size[0] = image_width;
    data->bgr_to_gray_data.global_size[1] = image_height;
}

void clifReleaseBuffers(CLIFEnvironmentData* data) {
    if (!data || !data->environment.impl) return;
    ClifState* s = (ClifState*)data->environment.impl;
    if (s->sum_hdr) cvReleaseMat(&s->sum_hdr);
    if (s->sq_hdr) cvReleaseMat(&s->sq_hdr);
    if (s->gray_hdr) cvReleaseImageHeader(&s->gray_hdr);
    s->sum.clear(); s->sqsum.clear(); s->gray.clear();
    data->integral_image_data.ptr = data->integral_image_data.square_ptr = nullptr;
}

CLIFGrayscaleResult clifGrayscale(const IplImage* source, CLIFEnvironmentData* data, const cl_bool) {
    ClifState* s = state(data);
    CLIFGrayscaleResult r;
    int step = 0;
    const unsigned char* g = gray_plane(s, source, &step);
    if (source->nChannels == 1) {   // already gray: copy so that the result is environment-owned
        s->gray.resize((size_t)source->width * source->height);
        for (int y = 0; y < source->height; y++) memcpy(&s->gray[(size_t)y * source->width], g + (size_t)y * step, source->width);
    }
    if (!s->gray_hdr) s->gray_hdr = cvCreateImageHeader(cvSize(source->width, source->height), IPL_DEPTH_8U, 1);
    s->gray_hdr->width = source->width; s->gray_hdr->height = source->height;
    s->gray_hdr->widthStep = source->width; s->gray_hdr->imageData = (char*)s->gray.data();
    data->bgr_to_gray_data.ptr = s->gray.data();
    r.image = s->gray_hdr;
    return r;
}

static CLIFIntegralResult integral_of(const IplImage* source, CLIFEnvironmentData* data) {
    ClifState* s = state(data);
    ensure_headers(s, source->width, source->height);
    // one upload of the frame as it is; colour conversion and integral images on the device (no gray round trip)
    CHECK(clfd_integral_image(s->ctx, (const uint8_t*)source->imageData, source->width, source->height, source->widthStep,
                              source->nChannels, s->sum.data(), s->sqsum.data(), nullptr, nullptr, 0));
    data->integral_image_data.ptr = s->sum.data();
    data->integral_image_data.square_ptr = s->sqsum.data();
    CLIFIntegralResult r;
    r.image = s->sum_hdr;
    r.square_image = s->sq_hdr;
    return r;
}

// The reference's clifIntegral expects a gray image but main.cpp:68 hands it the BGR frame
// (and its upload was broken, clif.cpp:290); here 3-channel input is converted first, which
// makes clifIntegral == clifGrayscaleIntegral for colour input.
CLIFIntegralResult clifIntegral(const IplImage* source, CLIFEnvironmentData* data, const cl_bool) { return integral_of(source, data); }
CLIFIntegralResult clifGrayscaleIntegral(const IplImage* source, CLIFEnvironmentData* data, const cl_bool) { return integral_of(source, data); }
