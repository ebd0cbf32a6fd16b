/*
 * voxcarve_oracle.c — CPU restatement of the reference voxel hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this. The product
 * (libvoxcarve.so) never links, imports or calls anything in oracle/.
 *
 * Parity pin: the projection arithmetic below is checked bit-for-bit against
 * cv2.gemm 4.13.0 (tests/golden/gemm_kat.npz, 52 800 vectors incl. adversarial
 * near-tie rows) and the whole carve/colour loop against a literal per-voxel
 * cv2.gemm run of the reference algorithm (tests/golden/{box,human}_literal.npz);
 * generator: tools/make_goldens.py. The reference C++ itself cannot be built here
 * (needs OpenCV C++ + Eigen, neither installed) so there is no oracle/_ref.
 * Not pinned by any reference fixture (recalled from OpenCV matx.hpp): the f64
 * accumulation inside cv::norm(Vec4f) used for the colour depth (vo_depth).
 *
 * Every function cites the reference lines it restates (paths under /root/reference/src).
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -pthread)
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

#define VO_API __attribute__((visibility("default")))

/* VoxelCarving.cpp:19 first product `intr * pose` (3x3 . 3x4, CV_32F): OpenCV's small
 * matrix path = plain f32, left to right, no FMA [pinned: gemm_kat.npz KM]. */
VO_API void vo_gemm3x3_3x4(const float* K, const float* M, float* P) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) {
            float a = K[i * 3 + 0] * M[0 * 4 + j]; /* -ffp-contract=off: no FMA */
            float b = K[i * 3 + 1] * M[1 * 4 + j];
            float c = K[i * 3 + 2] * M[2 * 4 + j];
            float t = a + b;
            P[i * 4 + j] = t + c;
        }
}

/* VoxelCarving.cpp:19 second product `(intr*pose) * world` (3x4 . 4x1, CV_32F):
 * products and a SEQUENTIAL accumulation ((s0+s1)+s2)+s3 in f64, one rounding to f32
 * [pinned: gemm_kat.npz proj; SURVEY §8c-5 states a different association, which the
 * near-tie vectors refute]. */
VO_API void vo_project(const float* P, const float* w, float* proj) {
    for (int i = 0; i < 3; i++) {
        double s = (double)P[i * 4 + 0] * (double)w[0];
        s = s + (double)P[i * 4 + 1] * (double)w[1];
        s = s + (double)P[i * 4 + 2] * (double)w[2];
        s = s + (double)P[i * 4 + 3] * (double)w[3];
        proj[i] = (float)s;
    }
}

/* Model.h:134-136 toWord: (y*size, x*size, -1*z*size, 1) — int*float products in f32. */
VO_API void vo_world(int x, int y, int z, float s, float* w) {
    w[0] = (float)y * s;
    w[1] = (float)x * s;
    w[2] = (float)(-1 * z) * s;
    w[3] = 1.0f;
}

/* VoxelCarving.cpp:44 `(int)std::round(float)`: half away from zero; NaN, inf and
 * values outside int range convert to INT_MIN on x86 (cvttss2si), i.e. out of bounds. */
VO_API int32_t vo_round_to_int(float v) {
    float r = roundf(v);
    if (!(r >= -2147483648.0f && r < 2147483648.0f)) return INT32_MIN;
    return (int32_t)r;
}

/* VoxelCarving.cpp:18-21 + :44: pixel of a voxel in one view. Returns 1 if inside the image (:45). */
static inline int vo_pixel(const float* P, int x, int y, int z, float s, int W, int H, int* px, int* py) {
    float w[4], p[3];
    vo_world(x, y, z, s, w);
    vo_project(P, w, p);
    float u = p[0] / p[2], v = p[1] / p[2]; /* :20, IEEE f32 divides, no depth-sign test */
    *px = vo_round_to_int(u);
    *py = vo_round_to_int(v);
    return *px >= 0 && *px < W && *py >= 0 && *py < H; /* cv::Point::inside(Rect(0,0,W,H)) */
}

VO_API int vo_pixel_of(const float* P, int x, int y, int z, float s, int W, int H, int* px, int* py, float* uv) {
    float w[4], p[3];
    vo_world(x, y, z, s, w);
    vo_project(P, w, p);
    uv[0] = p[0] / p[2];
    uv[1] = p[1] / p[2];
    return vo_pixel(P, x, y, z, s, W, H, px, py);
}

static inline int mask_is_bg(const uint32_t* bits, const uint8_t* bgr, int v, int W, int H, int px, int py) {
    if (bits) {
        int Ww = (W + 31) / 32;
        return (bits[((size_t)v * H + py) * Ww + (px >> 5)] >> (px & 31)) & 1u;
    }
    const uint8_t* p = bgr + (((size_t)v * H + py) * W + px) * 3;
    return p[0] == 0 && p[1] == 0 && p[2] == 0; /* VoxelCarving.cpp:50 */
}

/* One z-range of VoxelCarving.cpp:60-72 (views outer) / :39-55 (x, y, z loops in the order
 * of Model.h:10-35), on byte-per-voxel scratch indexed like Model::flatten (Model.h:104-106). */
static void carve_range(int X, int Y, int zlo, int zhi, int z0, float s, int V, int W, int H, const float* P,
                        const uint32_t* bits, const uint8_t* bgr, uint8_t* occ8, uint8_t* seen8) {
    for (int v = 0; v < V; v++) {
        const float* Pv = P + (size_t)v * 12;
        for (int x = 0; x < X; x++)
            for (int y = 0; y < Y; y++)
                for (int z = zlo; z < zhi; z++) {
                    int px, py;
                    if (!vo_pixel(Pv, x, y, z, s, W, H, &px, &py)) continue; /* :45-48 */
                    size_t f = (size_t)x + (size_t)X * ((size_t)y + (size_t)Y * (size_t)(z - z0));
                    if (mask_is_bg(bits, bgr, v, W, H, px, py)) occ8[f] = 0; /* :50-53 */
                    seen8[f] = 1;                                            /* :54 */
                }
    }
}

static void pack_bits(const uint8_t* b8, int X, int Y, int nz, uint32_t* words) {
    int Wx = (X + 31) / 32;
    for (size_t row = 0; row < (size_t)Y * nz; row++)
        for (int j = 0; j < Wx; j++) {
            uint32_t wd = 0;
            for (int b = 0; b < 32 && j * 32 + b < X; b++) wd |= (uint32_t)(b8[row * X + j * 32 + b] != 0) << b;
            words[row * Wx + j] = wd;
        }
}

typedef struct {
    int X, Y, z0, z1, V, W, H;
    float s;
    const float* P;
    const uint32_t* bits;
    const uint8_t* bgr;
    uint8_t *occ8, *seen8;
    atomic_int next_z;
} carve_job;

static void* carve_worker(void* arg) {
    carve_job* j = (carve_job*)arg;
    for (;;) { /* z-planes are independent (flatten is z-major), so threads take one plane at a time */
        int z = atomic_fetch_add(&j->next_z, 1);
        if (z >= j->z1) break;
        carve_range(j->X, j->Y, z, z + 1, j->z0, j->s, j->V, j->W, j->H, j->P, j->bits, j->bgr, j->occ8, j->seen8);
    }
    return NULL;
}

VO_API int vo_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (int)n;
}

/* carve() VoxelCarving.cpp:60-72 on the z-slab [z0,z1) of an X*Y*Z grid. Exactly one of
 * mask_bits / mask_bgr is non-NULL. Outputs bit volumes word[((z-z0)*Y + y)*ceil(X/32) + (x>>5)],
 * bit x&31; occupied starts all-ones (Model.cpp:9-14, alpha = 1), seen all-zero.
 * nthreads = 1 is the reference as written (single thread); nthreads > 1 splits z-planes over
 * pthreads (nthreads < 1: all online cores). */
VO_API int vo_carve(int X, int Y, int Z, float s, int z0, int z1, int V, int W, int H, const float* P,
                    const uint32_t* mask_bits, const uint8_t* mask_bgr, uint32_t* occ, uint32_t* seen,
                    int nthreads) {
    if (X < 1 || Y < 1 || Z < 1 || z0 < 0 || z1 > Z || z0 >= z1 || (!mask_bits == !mask_bgr)) return -1;
    int nz = z1 - z0;
    size_t n = (size_t)X * Y * nz;
    uint8_t* occ8 = (uint8_t*)malloc(n);
    uint8_t* seen8 = (uint8_t*)calloc(n, 1);
    if (!occ8 || !seen8) return -2;
    memset(occ8, 1, n);
    if (nthreads < 1) nthreads = vo_max_threads();
    if (nthreads > nz) nthreads = nz;
    carve_job job = {X, Y, z0, z1, V, W, H, s, P, mask_bits, mask_bgr, occ8, seen8, 0};
    atomic_store(&job.next_z, z0);
    if (nthreads == 1) {
        carve_range(X, Y, z0, z1, z0, s, V, W, H, P, mask_bits, mask_bgr, occ8, seen8);
    } else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
        for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, carve_worker, &job);
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
        free(th);
    }
    pack_bits(occ8, X, Y, nz, occ);
    pack_bits(seen8, X, Y, nz, seen);
    free(occ8);
    free(seen8);
    return 0;
}

static inline int getbit(const uint32_t* w, int X, int Y, int Z, int x, int y, int z) {
    if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) return 0; /* Model::get OOB -> alpha 0 (Model.h:119-124) */
    int Wx = (X + 31) / 32;
    return (w[((size_t)z * Y + y) * Wx + (x >> 5)] >> (x & 31)) & 1u;
}
static inline void putbit(uint32_t* w, int X, int Y, int x, int y, int z, int v) {
    int Wx = (X + 31) / 32;
    uint32_t* p = &w[((size_t)z * Y + y) * Wx + (x >> 5)];
    if (v) *p |= 1u << (x & 31); else *p &= ~(1u << (x & 31));
}

/* fastCarve() VoxelCarving.cpp:74-167: BFS from (0,0,0); `visited` is the seen bit (Model.h:154-160). */
VO_API int vo_fast_carve(int X, int Y, int Z, float s, int V, int W, int H, const float* P,
                         const uint32_t* mask_bits, const uint8_t* mask_bgr, uint32_t* occ, uint32_t* seen) {
    int Wx = (X + 31) / 32;
    size_t nw = (size_t)Wx * Y * Z;
    memset(seen, 0, nw * 4);
    for (size_t i = 0; i < nw; i++) /* Model ctor: every voxel occupied; row padding bits stay 0 */
        occ[i] = ((i % Wx) == (size_t)Wx - 1 && X % 32) ? (1u << (X % 32)) - 1 : 0xffffffffu;
    size_t cap = 1 << 16, head = 0, tail = 0;
    int32_t* q = (int32_t*)malloc(cap * 3 * sizeof(int32_t));
    if (!q) return -2;
#define PUSH(a, b, c) do { if (tail == cap) { cap *= 2; q = (int32_t*)realloc(q, cap * 3 * sizeof(int32_t)); if (!q) return -2; } \
        q[tail * 3] = (a); q[tail * 3 + 1] = (b); q[tail * 3 + 2] = (c); tail++; } while (0)
    PUSH(0, 0, 0); /* :101 */
    while (head < tail) {
        int x = q[head * 3], y = q[head * 3 + 1], z = q[head * 3 + 2];
        head++;
        if (getbit(seen, X, Y, Z, x, y, z)) continue; /* :107 */
        putbit(seen, X, Y, x, y, z, 1);               /* :110 */
        int carved = 0;
        for (int v = 0; v < V; v++) { /* :113-130 */
            int px, py;
            if (!vo_pixel(P + (size_t)v * 12, x, y, z, s, W, H, &px, &py)) continue;
            if (mask_is_bg(mask_bits, mask_bgr, v, W, H, px, py)) { putbit(occ, X, Y, x, y, z, 0); carved = 1; break; }
        }
        if (carved) { /* :132-163 */
            if (x > 0 && !getbit(seen, X, Y, Z, x - 1, y, z)) PUSH(x - 1, y, z);
            if (x < X - 1 && !getbit(seen, X, Y, Z, x + 1, y, z)) PUSH(x + 1, y, z);
            if (y > 0 && !getbit(seen, X, Y, Z, x, y - 1, z)) PUSH(x, y - 1, z);
            if (y < Y - 1 && !getbit(seen, X, Y, Z, x, y + 1, z)) PUSH(x, y + 1, z);
            if (z > 0 && !getbit(seen, X, Y, Z, x, y, z - 1)) PUSH(x, y, z - 1);
            if (z < Z - 1 && !getbit(seen, X, Y, Z, x, y, z + 1)) PUSH(x, y, z + 1);
        }
    }
#undef PUSH
    free(q);
    return 0;
}

/* ColorReconstruction.h:59 depth = cv::norm(cameras[i] - word_coord): Vec4f difference in f32,
 * squares accumulated in f64, sqrt, narrowed to the `float depth` of Model::addColor (Model.h:142). */
VO_API float vo_depth(const float* cam4, const float* w4) {
    double acc = 0;
    for (int k = 0; k < 4; k++) {
        float d = cam4[k] - w4[k];
        acc += (double)d * (double)d;
    }
    return (float)sqrt(acc);
}

/* voxel_pass_loops (ColorReconstruction.h:34-74) + body reconstructClosestColor (.cpp:33-41, mode 1)
 * or reconstructAvgColor (.cpp:59-66, mode 2), on a full-grid occupancy bit volume.
 * Emits one record per SURFACE voxel (alpha != 0 && !isInner, .h:46) in ascending flatten order
 * z,y,x: idx = x + X*(y + Y*z), rgbn = (r, g, b, min(n_observations, 255)); voxels with n = 0 keep
 * MODEL_COLOR (50,168,141) because the body `continue`s (.cpp:29-31). Returns the record count
 * (counts beyond cap are still counted, not written). */
VO_API uint64_t vo_color(int X, int Y, int Z, float s, int V, int W, int H, const float* P, const float* M,
                         const uint8_t* images_bgr, const uint32_t* occ, int mode, uint64_t* idx_out,
                         uint8_t* rgbn_out, uint64_t cap) {
    uint64_t n = 0;
    for (int z = 0; z < Z; z++)
        for (int y = 0; y < Y; y++)
            for (int x = 0; x < X; x++) {
                if (!getbit(occ, X, Y, Z, x, y, z)) continue;
                int inner = getbit(occ, X, Y, Z, x - 1, y, z) && getbit(occ, X, Y, Z, x + 1, y, z) &&
                            getbit(occ, X, Y, Z, x, y - 1, z) && getbit(occ, X, Y, Z, x, y + 1, z) &&
                            getbit(occ, X, Y, Z, x, y, z - 1) && getbit(occ, X, Y, Z, x, y, z + 1); /* Model.h:126-132 */
                if (inner) continue;
                float w[4];
                vo_world(x, y, z, s, w);
                int nobs = 0;
                float sum[3] = {0, 0, 0}, best[3] = {50, 168, 141}, bestd = 0;
                for (int v = 0; v < V; v++) {
                    int px, py;
                    if (!vo_pixel(P + (size_t)v * 12, x, y, z, s, W, H, &px, &py)) continue; /* .h:54-57 */
                    const uint8_t* pix = images_bgr + (((size_t)v * H + py) * W + px) * 3;   /* .h:58 */
                    float rgb[3] = {(float)pix[2], (float)pix[1], (float)pix[0]};            /* .h:59 BGR->RGB */
                    float cam[4] = {M[v * 12 + 3], M[v * 12 + 7], M[v * 12 + 11], 1.0f};     /* .h:21 */
                    float d = vo_depth(cam, w);
                    if (nobs == 0 || d < bestd) { bestd = d; memcpy(best, rgb, sizeof best); } /* .cpp:34-40 */
                    for (int c = 0; c < 3; c++) sum[c] = sum[c] + rgb[c];                     /* .cpp:60-64 */
                    nobs++;
                }
                uint8_t out[4] = {50, 168, 141, (uint8_t)(nobs > 255 ? 255 : nobs)}; /* MODEL_COLOR, Model.h:90 */
                if (nobs > 0) {
                    for (int c = 0; c < 3; c++) {
                        float val = mode == 2 ? roundf(sum[c] / (float)nobs) : best[c]; /* .cpp:65-66 */
                        out[c] = (uint8_t)val;
                    }
                }
                if (n < cap) {
                    idx_out[n] = (uint64_t)x + (uint64_t)X * ((uint64_t)y + (uint64_t)Y * (uint64_t)z);
                    memcpy(rgbn_out + n * 4, out, 4);
                }
                n++;
            }
    return n;
}

/* number of triangles of each of the 256 cube configurations = entries of triTable[idx] / 3
 * (MarchingCubes.h:147-404); filled by the caller from the shared table file so that oracle and
 * product do not share code — see oracle/oracle.py. */
VO_API void vo_mc_classify(int X, int Y, int Z, const uint32_t* occ, const uint8_t* tri_count256,
                           uint64_t* hist256, uint64_t* n_active, uint64_t* n_tris) {
    memset(hist256, 0, 256 * sizeof(uint64_t));
    *n_active = 0;
    *n_tris = 0;
    for (int x = -1; x < X; x++) /* MarchingCubes.cpp:12-14 */
        for (int y = -1; y < Y; y++)
            for (int z = -1; z < Z; z++) {
                /* corner order MarchingCubes.h:537-552; bit i set iff corner EMPTY (w < 0.5, :479-484) */
                int c[8] = {getbit(occ, X, Y, Z, x + 1, y, z),     getbit(occ, X, Y, Z, x, y, z),
                            getbit(occ, X, Y, Z, x, y + 1, z),     getbit(occ, X, Y, Z, x + 1, y + 1, z),
                            getbit(occ, X, Y, Z, x + 1, y, z + 1), getbit(occ, X, Y, Z, x, y, z + 1),
                            getbit(occ, X, Y, Z, x, y + 1, z + 1), getbit(occ, X, Y, Z, x + 1, y + 1, z + 1)};
                int idx = 0;
                for (int i = 0; i < 8; i++) if (!c[i]) idx |= 1 << i;
                hist256[idx]++;
                if (idx != 0 && idx != 255) (*n_active)++; /* edgeTable[idx] != 0 (:486) */
                *n_tris += tri_count256[idx];
            }
}

/* ------------------------------------------------------------------------------------------------
 * "Next" rows of the hot path: applyClosure and the full marchingCubes on the dense
 * RGBA Model (std::vector<Vector4f> voxels, index = Model::flatten = x + X*(y + Y*z), Model.h:104-106).
 * ---------------------------------------------------------------------------------------------- */
static inline const float* vget(const float* rgba, int X, int Y, int Z, int x, int y, int z) {
    static const float zero[4] = {0, 0, 0, 0}; /* Model::get out of range (Model.h:119-122) */
    if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) return zero;
    return rgba + 4 * ((size_t)x + (size_t)X * ((size_t)y + (size_t)Y * (size_t)z));
}

/* applyClosure(model, kernelSize) Postprocessing3d.cpp:4-100.  "Dilution" (:20-58): a voxel with w > 0 is kept,
 * any other becomes the mean (sum / count, f32, neighbours visited x, y, z = outer..inner) of its in-grid
 * neighbours with w > 0, or (0,0,0,0).  The erosion (:60-96) tests `w < thresh` with thresh = 0 and can never
 * fire, so the model becomes the dilated copy.  Returns -1 for an even kernel size (:8-11). */
VO_API int vo_closure(int X, int Y, int Z, int kernelSize, const float* in, float* out) {
    if (kernelSize % 2 != 1) return -1;
    const int size = (kernelSize - 1) / 2;
    for (int x = 0; x < X; x++)
        for (int y = 0; y < Y; y++)
            for (int z = 0; z < Z; z++) {
                const float* o = vget(in, X, Y, Z, x, y, z);
                float* dst = out + 4 * ((size_t)x + (size_t)X * ((size_t)y + (size_t)Y * (size_t)z));
                if (o[3] > 0.0f) { memcpy(dst, o, 16); continue; }
                int count = 0;
                float sum[4] = {0, 0, 0, 0};
                for (int i = -size; i <= size; i++) {
                    int xn = x + i;
                    if (xn < 0 || xn >= X) continue;
                    for (int j = -size; j <= size; j++) {
                        int yn = y + j;
                        if (yn < 0 || yn >= Y) continue;
                        for (int k = -size; k <= size; k++) {
                            int zn = z + k;
                            if (zn < 0 || zn >= Z) continue;
                            const float* v = vget(in, X, Y, Z, xn, yn, zn);
                            if (v[3] > 0.0f) { count++; for (int c = 0; c < 4; c++) sum[c] = sum[c] + v[c]; }
                        }
                    }
                }
                if (count > 0) for (int c = 0; c < 4; c++) sum[c] = sum[c] / (float)count;
                memcpy(dst, sum, 16);
            }
    return 0;
}

static int is_default_color(const float* c) { /* MODEL_COLOR / UNSEEN_COLOR heads (Model.h:90-91, MarchingCubes.h:453,457) */
    return (c[0] == 50.0f && c[1] == 168.0f && c[2] == 141.0f) || (c[0] == 204.0f && c[1] == 0.0f && c[2] == 0.0f);
}

/* VertexInterp MarchingCubes.h:428-468 */
static void vertex_interp(float thr, const float* p0, const float* v0, const float* p1, const float* v1, float* coord, float* color) {
    if (v0[3] == 0.0f && v1[3] != 0.0f) { memcpy(color, v1, 12); memcpy(coord, p1, 12); return; }
    if (v0[3] != 0.0f && v1[3] == 0.0f) { memcpy(color, v0, 12); memcpy(coord, p0, 12); return; }
    float f = (v0[3] == v1[3]) ? 0.5f : (thr - v0[3]) / (v1[3] - v0[3]);
    for (int c = 0; c < 3; c++) coord[c] = (1 - f) * p0[c] + f * p1[c];
    if (is_default_color(v0)) memcpy(color, v1, 12);
    else if (is_default_color(v1)) memcpy(color, v0, 12);
    else for (int c = 0; c < 3; c++) color[c] = (1 - f) * v0[c] + f * v1[c];
}

/* marchingCubes() MarchingCubes.cpp:8-31 -> ProcessVoxel MarchingCubes.h:532-578 -> Polygonise :478-511.
 * tri_table = triTable (256 x 16, -1 terminated, MarchingCubes.h:147-404); edgeTable follows from the topology.
 * Output per triangle, in the reference's emission order (x outer, y, z inner): 9 floats (3 unshared vertices, :561-568)
 * and 3 colours MeanColorFloats (:414-416, :570-573, incl. the col[2] = vertex i+1 quirk :506). Returns #triangles. */
VO_API uint64_t vo_marching_cubes(int X, int Y, int Z, const float* rgba, float thr, const int8_t* tri_table, float* verts,
                                  uint32_t* rgb, uint64_t cap) {
    static const int second[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
    uint64_t n = 0;
    for (int x = -1; x < X; x++)
        for (int y = -1; y < Y; y++)
            for (int z = -1; z < Z; z++) {
                const int cx[8] = {x + 1, x, x, x + 1, x + 1, x, x, x + 1}, cy[8] = {y, y, y + 1, y + 1, y, y, y + 1, y + 1},
                          cz[8] = {z, z, z, z, z + 1, z + 1, z + 1, z + 1};
                const float* val[8];
                float p[8][3];
                int idx = 0;
                for (int i = 0; i < 8; i++) {
                    val[i] = vget(rgba, X, Y, Z, cx[i], cy[i], cz[i]);
                    p[i][0] = (float)cx[i]; p[i][1] = (float)cy[i]; p[i][2] = (float)cz[i];
                    if (val[i][3] < thr) idx |= 1 << i;
                }
                int edges = 0;
                for (int e = 0; e < 12; e++) if (((idx >> (e % 8)) & 1) != ((idx >> second[e]) & 1)) edges |= 1 << e;
                if (edges == 0) continue;
                float vc[12][3], vcol[12][3];
                for (int e = 0; e < 12; e++)
                    if (edges & (1 << e)) vertex_interp(thr, p[e % 8], val[e % 8], p[second[e]], val[second[e]], vc[e], vcol[e]);
                const int8_t* t = tri_table + idx * 16;
                for (int i = 0; t[i] != -1; i += 3) {
                    if (n < cap) {
                        for (int k = 0; k < 3; k++) memcpy(verts + n * 9 + k * 3, vc[t[i + k]], 12);
                        const float *c0 = vcol[t[i]], *c1 = vcol[t[i + 1]], *c2 = vcol[t[i + 1]];
                        for (int c = 0; c < 3; c++) rgb[n * 3 + c] = (uint32_t)roundf(((c0[c] + c1[c]) + c2[c]) / 3);
                    }
                    n++;
                }
            }
    return n;
}

/* SimpleMesh::WriteMesh MarchingCubes.h:59-87: text .off, ostream default float formatting (== printf %g). */
VO_API int vo_write_off(const char* path, uint64_t ntris, const float* verts, const uint32_t* rgb, float scale, float tx, float ty, float tz) {
    FILE* f = fopen(path, "w");
    if (!f) return -1;
    fprintf(f, "OFF\n%llu %llu 0\n", (unsigned long long)(ntris * 3), (unsigned long long)ntris);
    for (uint64_t i = 0; i < ntris * 3; i++) {
        float a = verts[i * 3] * scale, b = verts[i * 3 + 1] * scale, c = verts[i * 3 + 2] * scale;
        a = a + tx; b = b + ty; c = c + tz;
        fprintf(f, "%g %g %g\n", a, b, c);
    }
    for (uint64_t i = 0; i < ntris; i++)
        fprintf(f, "3 %llu %llu %llu %u %u %u\n", (unsigned long long)(3 * i), (unsigned long long)(3 * i + 1), (unsigned long long)(3 * i + 2),
                rgb[i * 3], rgb[i * 3 + 1], rgb[i * 3 + 2]);
    fclose(f);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * cv::undistort as the reference calls it (VoxelCarving.cpp:36,86,89; ColorReconstruction.h:23,26): 8UC3 image,
 * camera matrix K (3x3 f64), distortion (k1 k2 p1 p2 [k3 [k4 k5 k6]]).  The algorithm lives in OpenCV (un-vendored,
 * un-pinned by the reference; pinned here to cv2 4.13.0 through tests/golden/undistort_kat.npz):
 *   stripes of min(max(1, 4096/cols), rows) rows; per stripe the new camera matrix is K with cy - y0 and the map is
 *   initUndistortRectifyMap(K, dist, I, K') in f64:  (x, y) = K'^-1 (j, i, 1);  radial/tangential model;  (u, v) = K (xd, yd, 1);
 *   fixed point: iu = round_half_even(u * 32), sx = iu >> 5, fx = iu & 31 (same for v);
 *   remap INTER_LINEAR, BORDER_CONSTANT(0): out = (sum_taps src * w + 2^14) >> 15 with w = (32-fx|fx)(32-fy|fy) * 32.
 * ---------------------------------------------------------------------------------------------- */
static void inv3x3_lu(const double* A, double* inv) { /* Gaussian elimination with partial pivoting on [A | I], as cv::invert(DECOMP_LU) */
    double a[3][3], b[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { a[i][j] = A[i * 3 + j]; b[i][j] = i == j; }
    for (int i = 0; i < 3; i++) {
        int k = i;
        for (int j = i + 1; j < 3; j++) if (fabs(a[j][i]) > fabs(a[k][i])) k = j;
        if (k != i) for (int j = 0; j < 3; j++) { double t = a[i][j]; a[i][j] = a[k][j]; a[k][j] = t; t = b[i][j]; b[i][j] = b[k][j]; b[k][j] = t; }
        double d = -1 / a[i][i];
        for (int j = i + 1; j < 3; j++) {
            double alpha = a[j][i] * d;
            for (int c = i + 1; c < 3; c++) a[j][c] += alpha * a[i][c];
            for (int c = 0; c < 3; c++) b[j][c] += alpha * b[i][c];
        }
    }
    for (int i = 2; i >= 0; i--)
        for (int j = 0; j < 3; j++) {
            double s = b[i][j];
            for (int k = i + 1; k < 3; k++) s -= a[i][k] * b[k][j];
            b[i][j] = s / a[i][i];
        }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) inv[i * 3 + j] = b[i][j];
}

VO_API int vo_undistort(int W, int H, const uint8_t* src, const double* K, const double* dist, int n_dist, uint8_t* dst) {
    if (n_dist != 4 && n_dist != 5 && n_dist != 8) return -1;
    double k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n_dist; i++) k[i] = dist[i];
    const double k1 = k[0], k2 = k[1], p1 = k[2], p2 = k[3], k3 = k[4], k4 = k[5], k5 = k[6], k6 = k[7];
    const double fx = K[0], fy = K[4], u0 = K[2], v0 = K[5];
    int stripe0 = (1 << 12) / (W > 1 ? W : 1);
    if (stripe0 < 1) stripe0 = 1;
    if (stripe0 > H) stripe0 = H;
    for (int y0 = 0; y0 < H; y0 += stripe0) {
        int n = stripe0 < H - y0 ? stripe0 : H - y0;
        double Ar[9], ir[9];
        memcpy(Ar, K, sizeof Ar);
        Ar[5] = v0 - y0;
        inv3x3_lu(Ar, ir);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < W; j++) {
                double _x = (i * ir[1] + ir[2]) + j * ir[0], _y = (i * ir[4] + ir[5]) + j * ir[3], _w = (i * ir[7] + ir[8]) + j * ir[6];
                double w = 1. / _w, x = _x * w, y = _y * w;
                double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
                double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
                double xd = (x * kr + p1 * _2xy) + p2 * (r2 + 2 * x2);
                double yd = (y * kr + p1 * (r2 + 2 * y2)) + p2 * _2xy;
                double u = fx * xd + u0, v = fy * yd + v0;
                double ru = nearbyint(u * 32), rv = nearbyint(v * 32); /* cvRound: round half to even, saturating */
                int iu = ru >= 2147483647.0 ? INT32_MAX : (ru <= -2147483648.0 || ru != ru ? INT32_MIN : (int)ru);
                int iv = rv >= 2147483647.0 ? INT32_MAX : (rv <= -2147483648.0 || rv != rv ? INT32_MIN : (int)rv);
                int sx = iu >> 5, sy = iv >> 5, ax = iu & 31, ay = iv & 31;
                int wt[4] = {(32 - ax) * (32 - ay) * 32, ax * (32 - ay) * 32, (32 - ax) * ay * 32, ax * ay * 32};
                uint8_t* o = dst + ((size_t)(y0 + i) * W + j) * 3;
                for (int c = 0; c < 3; c++) {
                    int acc = 0;
                    for (int t = 0; t < 4; t++) {
                        int xx = sx + (t & 1), yy = sy + (t >> 1);
                        int val = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? src[((size_t)yy * W + xx) * 3 + c] : 0;
                        acc += val * wt[t];
                    }
                    acc = (acc + (1 << 14)) >> 15;
                    o[c] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
                }
            }
    }
    return 0;
}
