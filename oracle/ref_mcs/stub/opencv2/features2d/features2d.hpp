// forwards to the stand-in (see ../opencv.hpp): test infrastructure
#include "../opencv.hpp"
