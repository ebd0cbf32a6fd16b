// A translation unit that uses the shim the way main.cpp:260-303 uses the reference (type check only, never linked).
#include <string>
#include <vector>
// the reference's headers in the order main.cpp:4-11 includes them (stand-ins with the same prototypes and default
// arguments), THEN the shim - exactly what INTEGRATION.md step 3 prescribes
#include "Calibration.h"
#include "PoseEstimation.h"
#include "Segmentation.h"
#include "VoxelCarving.h"
#include "ColorReconstruction.h"
#include "MarchingCubes.h"
#include "Postprocessing3d.h"
#include "Benchmark.h"

#include "voxcarve_shim.hpp"

#ifndef VOXCARVE_SHIM_ENABLED
#error "the shim did not switch itself on although the reference's headers are on the include path"
#endif

int pipeline(cv::Mat& cameraMatrix, cv::Mat& distCoeffs, std::vector<cv::Mat>& images, std::vector<cv::Mat>& masks) {
    Model model(100, 100, 100, 0.0028f);
    carve(cameraMatrix, distCoeffs, model, images, masks);
    carve(cameraMatrix, distCoeffs, model, images, masks, true);
    fastCarve(cameraMatrix, distCoeffs, model, images, masks);
    reconstructClosestColor(cameraMatrix, distCoeffs, model, images, masks);
    reconstructAvgColor(cameraMatrix, distCoeffs, model, images, masks);
    model.handleUnseen();
#ifdef VOXCARVE_SHIM_REPLACE_POSTPROCESSING
    applyClosure(&model, 3);
    return marchingCubes(&model, 1.0f, Vector3f(0, 0, 0), 0.5f, "out/mesh.off") ? 0 : 1;
#else
    return 0;
#endif
}
