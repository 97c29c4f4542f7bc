// oracle/compat/opencv2/highgui/highgui.hpp -- TEST INFRASTRUCTURE (CPU oracle shim); nothing from highgui is used.
#pragma once
#include "../core/core.hpp"
