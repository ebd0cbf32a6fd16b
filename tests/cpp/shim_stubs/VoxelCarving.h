// stand-in for the reference's VoxelCarving.h: its two prototypes, default argument included (VoxelCarving.h:19,31)
#pragma once
#include <vector>
#include "Model.h"
#include <opencv2/core/mat.hpp>
void carve(cv::Mat& cameraMatrix, cv::Mat& distCoeffs, Model& model, std::vector<cv::Mat>& images, std::vector<cv::Mat>& masks, bool intermediateMeshes = false);
void fastCarve(cv::Mat& cameraMatrix, cv::Mat& distCoeffs, Model& model, std::vector<cv::Mat>& images, std::vector<cv::Mat>& masks);
