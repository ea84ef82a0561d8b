/* stand-in for <opencv/cv.h>: everything lives in cvshim.hpp (oracle/_ref build only) */
#include "cvshim.hpp"
