// stand-in for the reference's ColorReconstruction.h: its two prototypes (ColorReconstruction.h:131,142); the voxel_pass macros
// of the real header (:10-74) are only used by ColorReconstruction.cpp, which the shim replaces
#pragma once
#include <vector>
#include "Model.h"
#include <opencv2/core/mat.hpp>
void reconstructClosestColor(cv::Mat& cameraMatrix, cv::Mat& distCoeffs, Model& model, std::vector<cv::Mat>& images, std::vector<cv::Mat>& masks);
void reconstructAvgColor(cv::Mat& cameraMatrix, cv::Mat& distCoeffs, Model& model, std::vector<cv::Mat>& images, std::vector<cv::Mat>& masks);
