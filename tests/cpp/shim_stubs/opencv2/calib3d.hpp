// stand-in for <opencv2/calib3d.hpp> (see README.md)
#pragma once
#include "opencv2/core/mat.hpp"
