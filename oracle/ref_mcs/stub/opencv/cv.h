// forwards to the stand-in (see ../opencv2/opencv.hpp): test infrastructure
#include "../opencv2/opencv.hpp"
