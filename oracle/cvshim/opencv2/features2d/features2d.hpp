/* stand-in for <opencv2/features2d/features2d.hpp>: everything lives in cvshim.hpp (oracle/_ref build only) */
#include "cvshim.hpp"
