// Probe: does this GPU grant compressible (CU_MEM_ALLOCATION_COMP_GENERIC) memory, and what does a uniform fill cost in it?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o compress_probe tools/experiments/compress_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); printf("%s -> %s\n", #x, s_); exit(1); } } while (0)
#define RT(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(r_)); exit(1); } } while (0)

__global__ void fill_kernel(uint4* p, size_t n, uint32_t v, int mode) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 w;
        if (mode == 0) w = make_uint4(v, v, v, v);
        else { uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; w = make_uint4(h, h * 3u + 1u, h ^ 0x9e3779b9u, h * 7u); }
        __stcs(p + i, w);
    }
}
__global__ void sum_kernel(const uint4* p, size_t n, unsigned long long* out) {
    unsigned long long s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 w = __ldcs(p + i);
        s += w.x + w.y + w.z + w.w;
    }
    if (s == 0x1234567ull) *out = s;
}

static float time_fill(uint4* p, size_t n, uint32_t v, int mode, int blocks) {
    cudaEvent_t a, b;
    RT(cudaEventCreate(&a)); RT(cudaEventCreate(&b));
    float best = 1e9f;
    for (int r = 0; r < 6; r++) {
        RT(cudaEventRecord(a));
        fill_kernel<<<blocks, 256>>>(p, n, v, mode);
        RT(cudaEventRecord(b));
        RT(cudaEventSynchronize(b));
        float ms; RT(cudaEventElapsedTime(&ms, a, b));
        if (r && ms < best) best = ms;
    }
    return best;
}
static float time_sum(const uint4* p, size_t n, unsigned long long* out, int blocks) {
    cudaEvent_t a, b;
    RT(cudaEventCreate(&a)); RT(cudaEventCreate(&b));
    float best = 1e9f;
    for (int r = 0; r < 6; r++) {
        RT(cudaEventRecord(a));
        sum_kernel<<<blocks, 256>>>(p, n, out);
        RT(cudaEventRecord(b));
        RT(cudaEventSynchronize(b));
        float ms; RT(cudaEventElapsedTime(&ms, a, b));
        if (r && ms < best) best = ms;
    }
    return best;
}

int main() {
    RT(cudaSetDevice(0));
    RT(cudaFree(0));
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    int comp = 0; CK(cuDeviceGetAttribute(&comp, CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev));
    printf("GENERIC_COMPRESSION_SUPPORTED = %d\n", comp);
    const size_t bytes = (size_t)1 << 30;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    prop.allocFlags.compressionType = CU_MEM_ALLOCATION_COMP_GENERIC;
    size_t gran = 0; CK(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    printf("granularity %zu\n", gran);
    const size_t sz = (bytes + gran - 1) / gran * gran;
    CUmemGenericAllocationHandle h; CK(cuMemCreate(&h, sz, &prop, 0));
    CUmemAllocationProp got = {}; CK(cuMemGetAllocationPropertiesFromHandle(&got, h));
    printf("granted compressionType = %d\n", (int)got.allocFlags.compressionType);
    CUdeviceptr va; CK(cuMemAddressReserve(&va, sz, 0, 0, 0));
    CK(cuMemMap(va, sz, 0, h, 0));
    CUmemAccessDesc acc = {}; acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemSetAccess(va, sz, &acc, 1));
    uint4* pc = (uint4*)va;
    uint4* pn; RT(cudaMalloc(&pn, bytes));
    unsigned long long* out; RT(cudaMalloc(&out, 8));
    const size_t n = bytes / 16;
    for (int blocks : {148, 148 * 4, 148 * 16}) {
        printf("blocks %d:\n", blocks);
        printf("  plain        fill ones %.4f ms  zeros %.4f ms  random %.4f ms", time_fill(pn, n, 0xffffffffu, 0, blocks), time_fill(pn, n, 0u, 0, blocks), time_fill(pn, n, 0u, 1, blocks));
        printf("  read(random) %.4f ms", time_sum(pn, n, out, blocks));
        time_fill(pn, n, 0xffffffffu, 0, blocks);
        printf("  read(ones) %.4f ms\n", time_sum(pn, n, out, blocks));
        printf("  compressible fill ones %.4f ms  zeros %.4f ms  random %.4f ms", time_fill(pc, n, 0xffffffffu, 0, blocks), time_fill(pc, n, 0u, 0, blocks), time_fill(pc, n, 0u, 1, blocks));
        printf("  read(random) %.4f ms", time_sum(pc, n, out, blocks));
        time_fill(pc, n, 0xffffffffu, 0, blocks);
        printf("  read(ones) %.4f ms\n", time_sum(pc, n, out, blocks));
    }
    RT(cudaMemset(pn, 0xff, bytes));
    cudaEvent_t a, b; RT(cudaEventCreate(&a)); RT(cudaEventCreate(&b));
    for (int which = 0; which < 2; which++) {
        float best = 1e9f;
        for (int r = 0; r < 5; r++) {
            RT(cudaEventRecord(a)); RT(cudaMemsetAsync(which ? (void*)pc : (void*)pn, 0xff, bytes)); RT(cudaEventRecord(b)); RT(cudaEventSynchronize(b));
            float ms; RT(cudaEventElapsedTime(&ms, a, b)); if (r && ms < best) best = ms;
        }
        printf("cudaMemset 1 GiB %s: %.4f ms\n", which ? "compressible" : "plain", best);
    }
    // a D2H copy out of compressible memory still works?
    uint32_t hostw[4] = {0, 0, 0, 0};
    RT(cudaMemcpy(hostw, pc, 16, cudaMemcpyDeviceToHost));
    printf("first word of compressible buffer: %08x\n", hostw[0]);
    return 0;
}
