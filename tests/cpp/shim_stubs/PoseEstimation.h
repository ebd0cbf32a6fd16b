// stand-in for the reference's PoseEstimation.h: estimatePoseFromImage (PoseEstimation.h:18), declaration only
#pragma once
#include <opencv2/core/mat.hpp>
cv::Mat estimatePoseFromImage(cv::Mat cameraMatrix, cv::Mat distCoeffs, cv::Mat image, bool visualize);
