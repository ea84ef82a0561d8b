/* stand-in for <opencv2/core/core.hpp>: everything lives in cvshim.hpp (oracle/_ref build only) */
#include "cvshim.hpp"
