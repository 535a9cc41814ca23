/* shim: stands in for opencv2/highgui/highgui.hpp of the reference build (OpenCV 2.4 / CLUtil are absent); see cv_shim.h */
#include "../../cv_shim.h"
