// stand-in for the reference's Postprocessing3d.h (Postprocessing3d.h:10)
#pragma once
#include "Model.h"
int applyClosure(Model* model, int kernelSize);
