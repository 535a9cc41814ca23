// clod_demo.cpp -- minimal client of the reference-facing C++ API (include/clif.h, clod.h):
// what a user of the reference's clod library writes, minus the highgui windows.
//   clod_demo <cascade.xml> <image.pgm|-> <scale_factor> <min_neighbors> [min_w min_h [scale_cascade]]
// Prints one "x y w h weight" line per detection and a summary of clifIntegral.
#include <cstdio>
#include <cstdlib>

#include "clod.h"

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s cascade.xml image.pgm scale min_neighbors [min_w min_h]\n", argv[0]); return 2; }
    CvHaarClassifierCascade* cascade = (CvHaarClassifierCascade*)cvLoad(argv[1], 0, 0, 0);
    if (!cascade) return 1;
    IplImage* frame = cvLoadImage(argv[2]);            // BGR, like main.cpp:48
    CvSize min_size = cvSize(argc > 5 ? atoi(argv[5]) : 0, argc > 6 ? atoi(argv[6]) : 0), max_size = cvSize(0, 0);

    CLODEnvironmentData* data = clodInitEnvironment(0);
    clifInitBuffers(data->clif, frame->width, frame->height, frame->widthStep, 3);
    CvSize isz = cvSize(frame->width, frame->height);
    clodInitBuffers(data, &isz);
    clodSetScaleFactor(data, atof(argv[3]));
    if (argc > 7) clodSetDetectionMode(data, atoi(argv[7]));   // 1: scaled features on one integral image (main.cpp:145's semantics)

    CLIFIntegralResult r = clifIntegral(frame, data->clif, CL_TRUE);
    const int W1 = frame->width + 1;
    printf("integral %d %llu\n", r.image->data.i[(size_t)frame->height * W1 + frame->width],
           (unsigned long long)((cl_ulong*)r.square_image->data.db)[(size_t)frame->height * W1 + frame->width]);

    CLODDetectObjectsResult res = clodDetectObjects(frame, cascade, data, min_size, max_size, (cl_uint)atoi(argv[4]),
                                                    CLOD_PRECOMPUTE_FEATURES | CLOD_PER_STAGE_ITERATIONS, CL_TRUE);
    printf("matches %u\n", res.match_count);
    for (cl_uint i = 0; i < res.match_count; i++)
        printf("%d %d %d %d %g\n", res.matches[i].rect.x, res.matches[i].rect.y, res.matches[i].rect.width,
               res.matches[i].rect.height, res.matches[i].weight);
    free(res.matches);
    clodReleaseBuffers(data);
    clodReleaseEnvironment(data);
    free(data);
    cvReleaseImage(&frame);
    cvReleaseHaarClassifierCascade(&cascade);
    return 0;
}
