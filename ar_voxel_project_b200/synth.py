"""Synthetic carve workloads: ring cameras, analytic silhouettes, seed-determined.

Grid N^3 with voxel size s = 0.28f/N (the physical extent of the reference's default grid,
main.cpp:26-29), so world x,y in [0,0.28), z in (-0.28,0] (Model::toWord, Model.h:134-136).
Cameras v = 0..V-1 on a ring of radius 0.60 around the grid centre, azimuth 2*pi*v/V, elevation
alternating 20/50 degrees, looking at the centre; K = [[f,0,W/2],[0,f,H/2],[0,0,1]], f = 0.78*W
(the datasets' 496.5/640), zero distortion.  M = [R|t] world->camera in f32, P = K32*M evaluated the
way cv::gemm does for 3x3.3x4 CV_32F (plain f32, left to right, no FMA; VoxelCarving.cpp:19).
Silhouette = exact per-pixel ray/quadric test of an ellipsoid plus three seeded spheres.
"""
import numpy as np

f32 = np.float32
CENTRE = np.array([0.14, 0.14, -0.14])


def gemm_k_m_f32(K32, M):
    """K32 (3x3) . M (3x4) exactly as OpenCV's small f32 gemm: (a0*b0 + a1*b1) + a2*b2, no FMA."""
    K32 = K32.astype(f32)
    M = M.astype(f32)
    t = (K32[:, 0:1] * M[0:1, :]).astype(f32) + (K32[:, 1:2] * M[1:2, :]).astype(f32)
    return (t.astype(f32) + (K32[:, 2:3] * M[2:3, :]).astype(f32)).astype(f32)


def cameras(V, W, H):
    f = 0.78 * W
    K = np.array([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=np.float64)
    K32 = K.astype(f32)
    Ms, Ps = [], []
    for v in range(V):
        az = 2 * np.pi * v / V
        el = np.deg2rad(20.0 if v % 2 == 0 else 50.0)
        pos = CENTRE + 0.60 * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), -np.sin(el)])
        zc = CENTRE - pos
        zc /= np.linalg.norm(zc)
        up = np.array([0.0, 0.0, -1.0])
        xc = np.cross(zc, up)
        xc /= np.linalg.norm(xc)
        yc = np.cross(zc, xc)
        R = np.stack([xc, yc, zc])
        t = -R @ pos
        M = np.concatenate([R, t[:, None]], axis=1).astype(f32)
        Ms.append(M)
        Ps.append(gemm_k_m_f32(K32, M))
    return K32, np.stack(Ms), np.stack(Ps)


def scene(seed):
    """Quadrics (centre, semi-axes): one ellipsoid + 3 spheres in the grid's middle half."""
    rng = np.random.default_rng(seed)
    q = [(CENTRE.copy(), np.array([0.08, 0.06, 0.11]))]
    for _ in range(3):
        r = rng.uniform(0.02, 0.04)
        c = np.array([rng.uniform(0.07, 0.21), rng.uniform(0.07, 0.21), -rng.uniform(0.07, 0.21)])
        q.append((c, np.array([r, r, r])))
    return q


def pack_bits(bg):
    """bool[..., H, W] (True = background) -> uint32[..., H, ceil(W/32)], bit x&31 of word x>>5."""
    *lead, H, W = bg.shape
    Ww = (W + 31) // 32
    pad = np.zeros((*lead, H, Ww * 32), dtype=np.uint8)
    pad[..., :W] = bg
    b = np.packbits(pad, axis=-1, bitorder="little")
    return np.ascontiguousarray(b).view("<u4").reshape(*lead, H, Ww)


def unpack_bits(words, n):
    b = np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")
    return b.reshape(*words.shape[:-1], -1)[..., :n].astype(bool)


def silhouettes(Ms, K32, W, H, quadrics):
    """-> uint32[V][H][ceil(W/32)], bit = 1 where NO quadric is hit (background => carve)."""
    V = len(Ms)
    Kd = K32.astype(np.float64)
    f, cx, cy = Kd[0, 0], Kd[0, 2], Kd[1, 2]
    out = np.empty((V, H, (W + 31) // 32), np.uint32)
    for v in range(V):
        R = Ms[v][:, :3].astype(np.float64)
        t = Ms[v][:, 3].astype(np.float64)
        pos = -R.T @ t
        fg = np.zeros((H, W), dtype=bool)
        for c, ax in quadrics:
            cc = R @ c + t  # centre in camera coordinates
            rb = float(ax.max())
            if cc[2] <= rb:  # not used by the stock scenes; render full frame if it ever happens
                x0, x1, y0, y1 = 0, W, 0, H
            else:
                rad = f * rb / (cc[2] - rb) * 1.05 + 2
                u, w_ = f * cc[0] / cc[2] + cx, f * cc[1] / cc[2] + cy
                x0, x1 = int(max(0, np.floor(u - rad))), int(min(W, np.ceil(u + rad) + 1))
                y0, y1 = int(max(0, np.floor(w_ - rad))), int(min(H, np.ceil(w_ + rad) + 1))
            if x0 >= x1 or y0 >= y1:
                continue
            px, py = np.meshgrid(np.arange(x0, x1, dtype=np.float64), np.arange(y0, y1, dtype=np.float64))
            dc = np.stack([(px - cx) / f, (py - cy) / f, np.ones_like(px)], axis=-1)  # ray dir, camera frame
            d = dc @ R  # = R^T dc, world frame
            o = (pos - c) / ax
            d = d / ax
            a = (d * d).sum(-1)
            b = (d * o).sum(-1)
            cq = (o * o).sum() - 1.0
            fg[y0:y1, x0:x1] |= (b * b - a * cq) >= 0.0
        out[v] = pack_bits(~fg)
    return out


def images(V, W, H, seed=0):
    """uint8[V][H][W][3] BGR, a cheap integer hash of (v, x, y, channel)."""
    v = np.arange(V, dtype=np.uint32)[:, None, None, None]
    y = np.arange(H, dtype=np.uint32)[None, :, None, None]
    x = np.arange(W, dtype=np.uint32)[None, None, :, None]
    c = np.arange(3, dtype=np.uint32)[None, None, None, :]
    h = (v * np.uint32(2654435761) + y * np.uint32(40503) + x * np.uint32(2246822519) + c * np.uint32(3266489917) + np.uint32(seed))
    h ^= h >> np.uint32(15)
    h *= np.uint32(2246822519)
    h ^= h >> np.uint32(13)
    return (h & np.uint32(255)).astype(np.uint8)


class Workload:
    """One synthetic config: grid N^3 (or X,Y,Z), V views of W x H."""

    def __init__(self, N, V, W, H, seed=0, dims=None):
        self.X, self.Y, self.Z = dims if dims else (N, N, N)
        self.N = N
        self.s = f32(0.28) / f32(N)
        self.V, self.W, self.H, self.seed = V, W, H, seed
        self.K32, self.M, self.P = cameras(V, W, H)
        self.mask_bits = silhouettes(self.M, self.K32, W, H, scene(seed))

    def mask_bgr(self):
        """the same silhouettes as 8UC3 images: background (0,0,0), object (255,255,255)."""
        fg = ~unpack_bits(self.mask_bits, self.W)
        return np.repeat((fg.astype(np.uint8) * 255)[..., None], 3, axis=-1)

    def images_bgr(self):
        return images(self.V, self.W, self.H, self.seed)

    @property
    def name(self):
        return f"synthetic {self.X}x{self.Y}x{self.Z} x {self.V} views {self.W}x{self.H} seed {self.seed}"


CONFIGS = {  # BASELINE.json configs[2..4]
    "C3": dict(N=512, V=36, W=640, H=480),
    "C4": dict(N=1024, V=72, W=1920, H=1080),
    "C5": dict(N=2048, V=72, W=3840, H=2160),
}


def noisy_masks(w, seed=0, p_noise=0.01, fringe=2):
    """Hostile version of a workload's silhouettes, as 8UC3 masks (the form the reference holds them in, VoxelCarving.cpp:36):
    a ragged `fringe`-pixel band of JPEG-like greys around every silhouette edge (30 % of the band exactly black) plus
    `p_noise` salt-and-pepper (half black holes inside the object, half white specks in the background; SURVEY §8c-8 measured
    0.5-1.3 % non-binary pixels on the real masks).  Only (0,0,0) carves (VoxelCarving.cpp:49-50).
    -> (uint8[V][H][W][3], the bit masks they pack to)"""
    from scipy import ndimage
    rng = np.random.default_rng(seed + 12345)
    fg = ~unpack_bits(w.mask_bits, w.W)
    out = np.empty((w.V, w.H, w.W, 3), np.uint8)
    for v in range(w.V):
        f = fg[v]
        g = f.astype(np.uint8) * 255
        band = ndimage.binary_dilation(f, iterations=fringe) & ~ndimage.binary_erosion(f, iterations=fringe)
        r = rng.integers(0, 256, size=f.shape, dtype=np.uint8)
        r[rng.random(f.shape) < 0.3] = 0
        g[band] = r[band]
        n = rng.random(f.shape)
        g[n < p_noise / 2] = 0
        g[(n >= p_noise / 2) & (n < p_noise)] = 255
        out[v] = g[..., None]
    return out, pack_bits((out == 0).all(-1))
