"""Host mirror of the reference `Model` (Model.h:93-163, Model.cpp:9-47) for SMALL grids.

Dense RGBA float voxels exactly like `std::vector<Vector4f> voxels` (alpha = occupancy) plus the
`seen` bits, indexed with Model::flatten = x + X*(y + Y*z) (Model.h:104-106).  The engine never uses
this class on its hot path: it is what the shim fills from the device bit volumes so that the
rest of the reference pipeline (handleUnseen, applyClosure, marchingCubes) can keep running, and
what the parity tests read.
"""
import numpy as np

MODEL_COLOR = (50.0, 168.0, 141.0, 1.0)   # Model.h:90
UNSEEN_COLOR = (204.0, 0.0, 0.0, 1.0)     # Model.h:91


class Model:
    def __init__(self, x, y, z, size):
        if x < 1 or y < 1 or z < 1:
            raise ValueError("You need to define a valid number of voxels for the model. (--x/--y/--z)")  # main.cpp:235-238
        if not size > 0:
            raise ValueError("You need to define a strictly positive voxel size. (--size)")  # main.cpp:241-245
        self.size_x, self.size_y, self.size_z = int(x), int(y), int(z)
        self.voxel_size = np.float32(size)
        self.voxels = np.tile(np.array(MODEL_COLOR, np.float32), (x * y * z, 1))  # Model.cpp:10-13
        self.seen = np.zeros(x * y * z, dtype=bool)

    def getX(self): return self.size_x
    def getY(self): return self.size_y
    def getZ(self): return self.size_z
    def getSize(self): return self.voxel_size

    def flatten(self, x, y, z):
        return x + self.size_x * (y + self.size_y * z)

    def get(self, x, y, z):
        if x < 0 or x >= self.size_x or y < 0 or y >= self.size_y or z < 0 or z >= self.size_z:
            return np.zeros(4, np.float32)  # Model.h:119-122
        return self.voxels[self.flatten(x, y, z)]

    def set(self, x, y, z, v):
        self.voxels[self.flatten(x, y, z)] = v

    def isInner(self, x, y, z):  # Model.h:126-132
        return all(self.get(*p)[3] != 0 for p in ((x - 1, y, z), (x + 1, y, z), (x, y - 1, z), (x, y + 1, z), (x, y, z - 1), (x, y, z + 1)))

    def toWord(self, x, y, z):  # Model.h:134-136 (sic)
        s = self.voxel_size
        return np.array([np.float32(y) * s, np.float32(x) * s, np.float32(-1 * z) * s, 1.0], np.float32)

    def see(self, x, y, z): self.seen[self.flatten(x, y, z)] = True
    def visit(self, v): self.see(v[0], v[1], v[2])
    def visited(self, v): return bool(self.seen[self.flatten(v[0], v[1], v[2])])

    def handleUnseen(self):  # Model.cpp:36-47: unseen voxels stay solid, painted UNSEEN_COLOR
        self.voxels[~self.seen] = np.array(UNSEEN_COLOR, np.float32)

    # ---- bridge to the engine's bit volumes -------------------------------------------
    def occupied_grid(self):
        """bool[Z][Y][X] of alpha != 0"""
        return (self.voxels[:, 3] != 0).reshape(self.size_z, self.size_y, self.size_x)

    def apply_carve(self, occ_words, seen_words):
        """What the C++ shim does after vc_carve: set(x,y,z,(0,0,0,0)) for cleared bits, see() for seen bits."""
        from .synth import unpack_bits
        occ = unpack_bits(occ_words, self.size_x).reshape(-1)
        seen = unpack_bits(seen_words, self.size_x).reshape(-1)
        self.voxels[~occ] = 0.0   # VoxelCarving.cpp:52
        self.seen |= seen         # VoxelCarving.cpp:54

    def apply_colors(self, idx, rgbn):
        """model.set(x,y,z,(r,g,b,1)) for every surface voxel with >= 1 observation (ColorReconstruction.cpp:41,66)."""
        m = rgbn[:, 3] > 0
        self.voxels[idx[m].astype(np.int64), :3] = rgbn[m, :3].astype(np.float32)
        self.voxels[idx[m].astype(np.int64), 3] = 1.0
