// TEST INFRASTRUCTURE: matcher.cpp:3 includes <opencv2/features2d.hpp> (OpenCV 3 layout); forwards to the compat shim.
#pragma once
#include "opencv2/features2d/features2d.hpp"
