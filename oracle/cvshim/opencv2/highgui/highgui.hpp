/* stand-in for <opencv2/highgui/highgui.hpp>: everything lives in cvshim.hpp (oracle/_ref build only) */
#include "cvshim.hpp"
