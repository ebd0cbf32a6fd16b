// Drives include/voxcarve_host.hpp the way main.cpp:248-303 drives the reference: a Model on the stack,
// carve -> reconstructAvgColor -> handleUnseen -> cube-index pass.  MiniModel is a stand-in with the
// interface of the reference Model (Model.h:93-163) so this builds without OpenCV/Eigen.
// usage: host_roundtrip <case.bin> <out.bin> [<mesh.off> [<intermediate dir>]]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "voxcarve_host.hpp"

struct Vec4 {
    float v[4];
    Vec4() : v{0, 0, 0, 0} {}
    Vec4(float a, float b, float c, float d) : v{a, b, c, d} {}
    float operator()(int i) const { return v[i]; }
};

class MiniModel {
   public:
    MiniModel(int x, int y, int z, float size) : X(x), Y(y), Z(z), s(size), voxels((size_t)x * y * z, Vec4(50, 168, 141, 1)), seen_((size_t)x * y * z, 0) {}
    int getX() { return X; }
    int getY() { return Y; }
    int getZ() { return Z; }
    float getSize() { return s; }
    Vec4 get(int x, int y, int z) {
        if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) return Vec4(0, 0, 0, 0);
        return voxels[flatten(x, y, z)];
    }
    void set(int x, int y, int z, const Vec4& v) { voxels[flatten(x, y, z)] = v; }
    void see(int x, int y, int z) { seen_[flatten(x, y, z)] = 1; }
    void handleUnseen() {
        for (size_t i = 0; i < voxels.size(); i++)
            if (!seen_[i]) voxels[i] = Vec4(204, 0, 0, 1);
    }
    std::vector<Vec4> voxels;
    std::vector<char> seen_;

   private:
    size_t flatten(int x, int y, int z) { return x + (size_t)X * (y + (size_t)Y * z); }
    int X, Y, Z;
    float s;
};

template <class T>
static void rd(FILE* f, std::vector<T>& v, size_t n) {
    v.resize(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[6];
    float s;
    if (fread(hdr, 4, 6, f) != 6 || fread(&s, 4, 1, f) != 1) return 2;
    const int X = hdr[0], Y = hdr[1], Z = hdr[2];
    vc::ViewCache views;
    views.V = hdr[3]; views.W = hdr[4]; views.H = hdr[5];
    rd(f, views.P, (size_t)views.V * 12);
    rd(f, views.M, (size_t)views.V * 12);
    rd(f, views.mask_bits, (size_t)views.V * views.H * ((views.W + 31) / 32));
    rd(f, views.images_bgr, (size_t)views.V * views.H * views.W * 3);
    fclose(f);
    try {
        MiniModel model(X, Y, Z, s);
        std::vector<vc::McSummary> perView;
        vc::carve(views, model, /*intermediateMeshes=*/true, &perView, argc > 4 ? argv[4] : "out/intermediate");
        vc::reconstructAvgColor(views, model);
        model.handleUnseen();
        const vc::McSummary mc = vc::marchingCubesClassify(model);
        FILE* o = fopen(argv[2], "wb");
        fwrite(model.voxels.data(), sizeof(Vec4), model.voxels.size(), o);
        fwrite(model.seen_.data(), 1, model.seen_.size(), o);
        fwrite(mc.hist, 8, 256, o);
        fwrite(&mc.triangles, 8, 1, o);
        const uint64_t nv = perView.size();
        fwrite(&nv, 8, 1, o);
        for (const auto& p : perView) fwrite(&p.triangles, 8, 1, o);
        fclose(o);
        if (argc > 3) {  // main.cpp:297-303: applyClosure(&model, 3); marchingCubes(&model, scale, translation, 0.5f, outFile)
            if (vc::applyClosure(&model, 2) != -1) return 4;
            if (vc::applyClosure(&model, 3) != 0) return 4;
            const float t[3] = {0.5f, -0.25f, 2.0f};
            if (!vc::marchingCubes(&model, 1.5f, t, 0.5f, argv[3])) return 4;
        }

    } catch (const vc::Error& e) {
        fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    return 0;
}
