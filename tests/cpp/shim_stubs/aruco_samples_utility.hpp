// stand-in for the reference's aruco_samples_utility.hpp (nothing of it is used by the shim itself)
#pragma once
