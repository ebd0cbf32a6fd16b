// stand-in for the reference's Segmentation.h (colour / k-means masks; nothing of it is used by the shim), see Calibration.h
#pragma once
