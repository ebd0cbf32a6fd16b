// Drives several engines from C++ the way a multi-GPU host built on the reference would (one engine per device, one thread per
// engine), through include/voxcarve_host.hpp and the C ABI only - no Python, no torch:
//   n_dev >= 2: NCCL inside libvoxcarve.so (vc_comm_init / vc_exchange_halos / vc_comm_allreduce_u64 / vc_gather)
//   n_dev == 1: the same slabs as several engines on device 0, halos swapped with vc_exchange_halos_peer (NCCL refuses two
//               ranks on one GPU)
// and in both cases vc::carveOnDevices.  Everything is compared with a single engine on the whole grid; the Python test
// compares that single-engine result with the oracle.
// usage: multi_gpu <case.bin> <n_dev> <n_slabs> <out.bin>      exit code 0 = every check passed
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "voxcarve_host.hpp"

struct Vec4 {
    float v[4];
    Vec4() : v{0, 0, 0, 0} {}
    Vec4(float a, float b, float c, float d) : v{a, b, c, d} {}
    float operator()(int i) const { return v[i]; }
};
class MiniModel {  // the interface of the reference Model (Model.h:93-163)
   public:
    MiniModel(int x, int y, int z, float size) : voxels((size_t)x * y * z, Vec4(50, 168, 141, 1)), seen_((size_t)x * y * z, 0), X(x), Y(y), Z(z), s(size) {}
    int getX() { return X; }
    int getY() { return Y; }
    int getZ() { return Z; }
    float getSize() { return s; }
    Vec4 get(int x, int y, int z) { return voxels[x + (size_t)X * (y + (size_t)Y * z)]; }
    void set(int x, int y, int z, const Vec4& v) { voxels[x + (size_t)X * (y + (size_t)Y * z)] = v; }
    void see(int x, int y, int z) { seen_[x + (size_t)X * (y + (size_t)Y * z)] = 1; }
    std::vector<Vec4> voxels;
    std::vector<char> seen_;

   private:
    int X, Y, Z;
    float s;
};

template <class T>
static void rd(FILE* f, std::vector<T>& v, size_t n) {
    v.resize(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}
static int g_fail = 0;
static void expect(bool ok, const char* what) {
    if (!ok) { fprintf(stderr, "CHECK FAILED: %s\n", what); g_fail++; }
}

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    const int n_dev = atoi(argv[2]), n_slabs = atoi(argv[3]);
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
        // single engine, whole grid
        std::vector<uint32_t> occ, seen;
        vc::McSummary mc{};
        std::vector<uint64_t> cidx;
        std::vector<uint8_t> crgbn;
        std::vector<int32_t> bounds;
        {
            vc::Engine e(X, Y, Z, s);
            e.setViews(views, true);
            bounds = e.planSlabs(n_slabs);
            e.carve();
            occ = e.occupied();
            seen = e.seen();
            mc = e.mcClassify();
            e.color(VC_COLOR_AVG);
            e.colors(cidx, crgbn);
        }
        const size_t plane = (size_t)Y * ((X + 31) / 32);
        std::vector<uint64_t> hist_sum(256, 0);
        std::vector<std::vector<uint64_t>> idx((size_t)n_slabs);
        std::vector<std::vector<uint8_t>> rgbn((size_t)n_slabs);
        std::vector<std::vector<uint32_t>> full_occ((size_t)n_slabs), full_seen((size_t)n_slabs);
        std::vector<std::string> errors((size_t)n_slabs);
        if (n_dev >= 2) {  // one thread and one NCCL rank per engine, engine r on device r % n_dev ... but NCCL needs distinct GPUs
            if (n_slabs > n_dev) { fprintf(stderr, "n_slabs must be <= n_dev with NCCL\n"); return 2; }
            unsigned char id[128];
            if (vc_comm_unique_id(id) != VC_OK) { fprintf(stderr, "%s\n", vc_last_error(nullptr)); return 3; }
            std::vector<std::vector<uint64_t>> reduced((size_t)n_slabs, std::vector<uint64_t>(256));
            std::vector<std::thread> th;
            for (int r = 0; r < n_slabs; r++)
                th.emplace_back([&, r] {
                    try {
                        vc::Engine e(X, Y, Z, s, bounds[r], bounds[r + 1], r);
                        e.allocFullVolumes();
                        e.setViews(views, true);
                        e.commInit(r, n_slabs, id);
                        e.carve();
                        e.exchangeHalos();
                        vc::McSummary m = e.mcClassify();
                        memcpy(reduced[r].data(), m.hist, sizeof m.hist);
                        e.allreduce(reduced[r].data(), 256);
                        e.color(VC_COLOR_AVG);
                        e.colors(idx[r], rgbn[r]);
                        e.gather(bounds, 3);
                        full_occ[r] = e.downloadFull(0);
                        full_seen[r] = e.downloadFull(1);
                    } catch (const std::exception& ex) {
                        errors[r] = ex.what();
                    }
                });
            for (auto& t : th) t.join();
            for (int r = 0; r < n_slabs; r++) {
                if (!errors[r].empty()) { fprintf(stderr, "rank %d: %s\n", r, errors[r].c_str()); return 3; }
                expect(memcmp(reduced[r].data(), mc.hist, sizeof mc.hist) == 0, "all-reduced cube-index histogram == single engine");
                expect(full_occ[r] == occ, "gathered occupied == single engine (every rank)");
                expect(full_seen[r] == seen, "gathered seen == single engine (every rank)");
            }
        } else {  // several engines on device 0
            std::vector<std::unique_ptr<vc::Engine>> es;
            std::vector<vc_engine*> hs;
            for (int r = 0; r < n_slabs; r++) {
                es.emplace_back(new vc::Engine(X, Y, Z, s, bounds[r], bounds[r + 1], 0));
                es.back()->setViews(views, true);
                es.back()->carve();
                hs.push_back(es.back()->handle());
            }
            if (vc_exchange_halos_peer(hs.data(), n_slabs) != VC_OK) { fprintf(stderr, "%s\n", vc_last_error(hs[0])); return 3; }
            for (int r = 0; r < n_slabs; r++) {
                vc::McSummary m = es[r]->mcClassify();
                for (int i = 0; i < 256; i++) hist_sum[i] += m.hist[i];
                es[r]->color(VC_COLOR_AVG);
                es[r]->colors(idx[r], rgbn[r]);
                std::vector<uint32_t> o = es[r]->occupied(), sn = es[r]->seen();
                expect(memcmp(o.data(), occ.data() + plane * bounds[r], o.size() * 4) == 0, "slab occupied == single engine");
                expect(memcmp(sn.data(), seen.data() + plane * bounds[r], sn.size() * 4) == 0, "slab seen == single engine");
            }
            expect(memcmp(hist_sum.data(), mc.hist, sizeof mc.hist) == 0, "summed cube-index histograms == single engine");
        }
        std::vector<uint64_t> all_idx;
        std::vector<uint8_t> all_rgbn;
        for (int r = 0; r < n_slabs; r++) {
            all_idx.insert(all_idx.end(), idx[r].begin(), idx[r].end());
            all_rgbn.insert(all_rgbn.end(), rgbn[r].begin(), rgbn[r].end());
        }
        expect(all_idx == cidx && all_rgbn == crgbn, "concatenated colour records == single engine");
        // the host-Model flow on several devices
        MiniModel a(X, Y, Z, s), b(X, Y, Z, s);
        vc::carve(views, a);
        std::vector<int> devs;
        for (int r = 0; r < n_slabs; r++) devs.push_back(n_dev >= 2 ? r % n_dev : 0);
        vc::carveOnDevices(views, b, devs);
        expect(memcmp(a.voxels.data(), b.voxels.data(), a.voxels.size() * sizeof(Vec4)) == 0 && a.seen_ == b.seen_, "carveOnDevices Model == carve Model");
        FILE* o = fopen(argv[4], "wb");
        fwrite(occ.data(), 4, occ.size(), o);
        fwrite(seen.data(), 4, seen.size(), o);
        fwrite(mc.hist, 8, 256, o);
        fclose(o);
    } catch (const vc::Error& e) {
        fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    return g_fail ? 1 : 0;
}
