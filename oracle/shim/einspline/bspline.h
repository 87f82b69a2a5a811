// TEST INFRASTRUCTURE ONLY -- see einspline_shim.h
#include "einspline_shim.h"
