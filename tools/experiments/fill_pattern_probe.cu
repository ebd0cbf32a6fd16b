// Probe: store-only bandwidth of two volumes written (A) in linear chunk order, (B) like vc_fill4_kernel: one block per
// (4 KB chunk of a plane, layer of 8 planes), 16 stores per thread at plane stride.  C5 geometry: 2048 rows x 64 words, nz planes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fill_pattern_probe tools/experiments/fill_pattern_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define RT(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(r_)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) linear_kernel(uint4* occ, uint4* seen, size_t n_quads, int streaming) {
    const uint4 o = make_uint4(0, 0, 0, 0), s = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_quads; i += (size_t)gridDim.x * 256) {
        if (streaming) { __stcs(occ + i, o); __stcs(seen + i, s); } else { occ[i] = o; seen[i] = s; }
    }
}
// grid (chunks per plane, layers or fewer): block (c, l) writes chunk c of the 8 planes of layers l, l + gridDim.y, ...
__global__ void __launch_bounds__(256) layered_kernel(uint4* occ, uint4* seen, unsigned quads_per_plane, unsigned nz, int streaming) {
    const uint4 o = make_uint4(0, 0, 0, 0), s = make_uint4(~0u, ~0u, ~0u, ~0u);
    const unsigned t = blockIdx.x * 256 + threadIdx.x;
    if (t >= quads_per_plane) return;
    for (unsigned l = blockIdx.y; l * 8 < nz; l += gridDim.y) {
        size_t i = (size_t)l * 8 * quads_per_plane + t;
        for (unsigned z = 0; z < 8 && l * 8 + z < nz; z++, i += quads_per_plane) {
            if (streaming) { __stcs(occ + i, o); __stcs(seen + i, s); } else { occ[i] = o; seen[i] = s; }
        }
    }
}
// the fused variant: `nb` blocks, block b takes chunk b % fx of layers b / fx, b / fx + fy, ...
__global__ void __launch_bounds__(256) fused_like_kernel(uint4* occ, uint4* seen, unsigned quads_per_plane, unsigned nz, int streaming) {
    const uint4 o = make_uint4(0, 0, 0, 0), s = make_uint4(~0u, ~0u, ~0u, ~0u);
    const unsigned per_plane = (quads_per_plane + 255) / 256;
    const unsigned fx = min(per_plane, gridDim.x), fy = gridDim.x / fx;
    if (blockIdx.x >= fx * fy) return;
    for (unsigned c = blockIdx.x % fx; c < per_plane; c += fx) {
        const unsigned t = c * 256 + threadIdx.x;
        if (t >= quads_per_plane) continue;
        for (unsigned l = blockIdx.x / fx; l * 8 < nz; l += fy) {
            size_t i = (size_t)l * 8 * quads_per_plane + t;
            for (unsigned z = 0; z < 8 && l * 8 + z < nz; z++, i += quads_per_plane) {
                if (streaming) { __stcs(occ + i, o); __stcs(seen + i, s); } else { occ[i] = o; seen[i] = s; }
            }
        }
    }
}
template <class F> static float best_of(F launch) {
    cudaEvent_t a, b; RT(cudaEventCreate(&a)); RT(cudaEventCreate(&b));
    float best = 1e9f;
    for (int r = 0; r < 6; r++) {
        RT(cudaEventRecord(a)); launch(); RT(cudaEventRecord(b)); RT(cudaEventSynchronize(b));
        float ms; RT(cudaEventElapsedTime(&ms, a, b)); if (r && ms < best) best = ms;
    }
    RT(cudaGetLastError());
    return best;
}
int main() {
    for (int cfg = 0; cfg < 2; cfg++) {
        const unsigned Y = cfg ? 2048 : 1024, Wx = cfg ? 64 : 32, nz = cfg ? 2048 : 1024;
        const unsigned qpp = Y * Wx / 4;
        const size_t n_quads = (size_t)qpp * nz, bytes = n_quads * 16;
        uint4 *occ, *seen; RT(cudaMalloc(&occ, bytes)); RT(cudaMalloc(&seen, bytes));
        printf("%s: 2 x %.0f MiB\n", cfg ? "C5" : "C4", bytes / 1048576.0);
        for (int st = 0; st < 2; st++) {
            for (int blocks : {148, 592, 2368}) {
                float ms = best_of([&] { linear_kernel<<<blocks, 256>>>(occ, seen, n_quads, st); });
                printf("  linear   %s grid %5d: %.4f ms  %.2f TB/s\n", st ? "stcs " : "plain", blocks, ms, 2.0 * bytes / ms * 1e-9);
            }
            float ms = best_of([&] { layered_kernel<<<dim3((qpp + 255) / 256, nz / 8), 256>>>(occ, seen, qpp, nz, st); });
            printf("  layered  %s grid (%u,%u): %.4f ms  %.2f TB/s\n", st ? "stcs " : "plain", (qpp + 255) / 256, nz / 8, ms, 2.0 * bytes / ms * 1e-9);
            for (int blocks : {148, 296, 592}) {
                ms = best_of([&] { fused_like_kernel<<<blocks, 256>>>(occ, seen, qpp, nz, st); });
                printf("  fused-like %s grid %5d: %.4f ms  %.2f TB/s\n", st ? "stcs " : "plain", blocks, ms, 2.0 * bytes / ms * 1e-9);
            }
        }
        RT(cudaFree(occ)); RT(cudaFree(seen));
    }
    return 0;
}
