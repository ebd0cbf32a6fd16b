// stand-in for the reference's Calibration.h (interactive ChArUco calibration; nothing of it is used by the shim): present so
// that the test's caller can include the reference's headers in main.cpp:4-11 order
#pragma once
