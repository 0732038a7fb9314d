// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <ros/ros.h>.
#pragma once
#include "ros/time.h"
