// clod_lifecycle.cpp -- a client that loads, uses, releases and reloads cascades the way a
// long-running user of the reference's API does, through BOTH detection entry points
// (clodDetectObjects and the shim's cvHaarDetectObjects):
//   clod_lifecycle <image.pgm> <scale_factor> <min_neighbors> <cascade_a.xml> <cascade_b.xml> [...]
// For each cascade in turn: cvLoad, detect on the same frame (same plan-cache key but for the
// cascade), print the rects, cvReleaseHaarClassifierCascade.  A plan cached under the address
// of a released cascade must never serve the next one (its address is often reused).
#include <cstdio>
#include <cstdlib>

#include "clod.h"

static void print_rect(const CvRect& r, double w) { printf("%d %d %d %d %g\n", r.x, r.y, r.width, r.height, w); }

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s image.pgm scale min_neighbors cascade.xml...\n", argv[0]); return 2; }
    IplImage* frame = cvLoadImage(argv[1]);
    const double sf = atof(argv[2]);
    const int mn = atoi(argv[3]);
    CLODEnvironmentData* data = clodInitEnvironment(0);
    clodSetScaleFactor(data, sf);
    CvMemStorage* storage = cvCreateMemStorage(0);
    for (int round = 0; round < 2; round++)
        for (int a = 4; a < argc; a++) {
            CvHaarClassifierCascade* cascade = (CvHaarClassifierCascade*)cvLoad(argv[a], 0, 0, 0);
            if (!cascade) return 1;
            CLODDetectObjectsResult res = clodDetectObjects(frame, cascade, data, cvSize(0, 0), cvSize(0, 0), (cl_uint)mn, 0, CL_TRUE);
            printf("clod %d %s %u\n", round, argv[a], res.match_count);
            for (cl_uint i = 0; i < res.match_count; i++) print_rect(res.matches[i].rect, res.matches[i].weight);
            free(res.matches);
            CvSeq* seq = cvHaarDetectObjects(frame, cascade, storage, sf, mn, CV_HAAR_SCALE_IMAGE, cvSize(0, 0), cvSize(0, 0));
            printf("cv %d %s %d\n", round, argv[a], seq->total);
            for (int i = 0; i < seq->total; i++) {
                CvAvgComp* c = (CvAvgComp*)cvGetSeqElem(seq, i);
                print_rect(c->rect, c->neighbors);
            }
            cvReleaseHaarClassifierCascade(&cascade);
        }
    clodReleaseBuffers(data);
    clodReleaseEnvironment(data);
    free(data);
    cvReleaseImage(&frame);
    return 0;
}
