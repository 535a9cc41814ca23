/* shim: stands in for params.h of the reference build (OpenCV 2.4 / CLUtil are absent); see cv_shim.h */
#include "cv_shim.h"
