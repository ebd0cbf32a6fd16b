"""SimpleMesh::WriteMesh (MarchingCubes.h:59-87): the text .off the reference writes after marchingCubes().

vertex = index-space coordinate * scaleFactor + translation in f32 (scaleFactor = scale * model->getSize(),
MarchingCubes.cpp:23); numbers are printed the way std::ostream prints a float by default (6 significant
digits, %g); faces are "3 i j k r g b" with three unshared vertices per triangle (MarchingCubes.h:561-575).
"""
import numpy as np


def write_off(path, verts, rgb, scale=1.0, translation=(0.0, 0.0, 0.0)):
    v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    c = np.ascontiguousarray(rgb).reshape(-1, 3)
    sf = np.float32(scale)
    t = np.asarray(translation, np.float32)
    w = (v * sf).astype(np.float32) + t  # f32 multiply, then f32 add
    with open(path, "w") as f:
        f.write("OFF\n%d %d 0\n" % (len(v), len(c)))
        f.write("".join("%g %g %g\n" % (float(a), float(b), float(d)) for a, b, d in w))
        f.write("".join("3 %d %d %d %d %d %d\n" % (3 * i, 3 * i + 1, 3 * i + 2, r, g, b) for i, (r, g, b) in enumerate(c)))
    return True
