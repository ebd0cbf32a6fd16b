// stand-in for the reference's MarchingCubes.h: the prototype with its default arguments (MarchingCubes.h:596)
#pragma once
#include <string>
#include "Model.h"
bool marchingCubes(Model* model, float scale = 1, Vector3f translation = Vector3f(0, 0, 0), float threshold = 0.5f,
                   std::string outFileName = "out/mesh.off");
