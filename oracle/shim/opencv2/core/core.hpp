// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <opencv2/core/core.hpp>.
#pragma once
#include "opencv2/core.hpp"
