"""B200-native voxel-carving engine behind the reference's VoxelCarving.h / Model.h interface.

The compute path is libvoxcarve.so (hand-written CUDA for sm_100a behind a C ABI, include/voxcarve.h);
this package is the host-side mirror of the reference interface for that path.
"""
from .engine import VoxelEngine, VoxCarveError  # noqa: F401
from .model import Model, MODEL_COLOR, UNSEEN_COLOR  # noqa: F401
from .api import carve, fastCarve, reconstructAvgColor, reconstructClosestColor, marchingCubesClassify, ViewSet  # noqa: F401
