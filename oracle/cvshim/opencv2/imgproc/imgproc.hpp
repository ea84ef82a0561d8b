/* stand-in for <opencv2/imgproc/imgproc.hpp>: everything lives in cvshim.hpp (oracle/_ref build only) */
#include "cvshim.hpp"
