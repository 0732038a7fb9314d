// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <opencv2/opencv.hpp>; see opencv2/core.hpp.
#pragma once
#include "opencv2/core.hpp"
