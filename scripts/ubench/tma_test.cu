// Minimal bulk-tensor (TMA) load probe: which box shapes / ranks load on this GPU.  nvcc -arch sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int RANK, int BX, int BY>
__global__ void k(const __grid_constant__ CUtensorMap tmapParam, const CUtensorMap *tmapGlobal, float *out, int cx, int cy, int cz) {
    const CUtensorMap &tmap = tmapGlobal ? *tmapGlobal : tmapParam;
    __shared__ __align__(128) float tile[BY][BX];
    __shared__ __align__(8) unsigned long long mbar;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&mbar), dst = (unsigned)__cvta_generic_to_shared(&tile[0][0]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)sizeof(tile)) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(cx), "r"(cy), "r"(cz), "r"(bar) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(cx), "r"(cy), "r"(bar) : "memory");
    }
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
    if (threadIdx.x == 0) { out[0] = tile[2][2]; out[1] = tile[0][0]; out[2] = tile[BY - 1][BX - 1]; }
}
static int gMode = 0;
template <int RANK, int BX, int BY>
void run(EncodeTiledFn fn, float *d, int N, int H, int W, float *out, const char *name) {
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {BX, BY, 1}, es[3] = {1, 1, 1};
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, RANK, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", name, (int)r); return; }
    CUtensorMap *gm = nullptr;
    if (gMode & 1) { cudaMalloc(&gm, sizeof(m)); cudaMemcpy(gm, &m, sizeof(m), cudaMemcpyHostToDevice); }
    const int c0 = gMode >> 4, c1 = (gMode & 2) ? 4 : -2;   // mode = x0 * 16 + flags (x0 may be negative)
    k<RANK, BX, BY><<<1, 128>>>(m, gm, out, c0, c1, 1);
    cudaError_t e = cudaDeviceSynchronize();
    float h[3] = {0, 0, 0};
    if (e == cudaSuccess) cudaMemcpy(h, out, 12, cudaMemcpyDeviceToHost);
    printf("%s: %s  tile[2][2]=%g (expect %g) tile[0][0]=%g last=%g\n", name, cudaGetErrorString(e), h[0], (float)(H * W), h[1], h[2]);
}
int main(int argc, char **argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    gMode = argc > 2 ? atoi(argv[2]) : 0;
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int N = 2, H = 384, W = 1280;
    float *d, *out; cudaMalloc(&d, (size_t)N * H * W * 4); cudaMalloc(&out, 64);
    float *h = (float *)malloc((size_t)N * H * W * 4);
    for (size_t i = 0; i < (size_t)N * H * W; i++) h[i] = (float)i;
    cudaMemcpy(d, h, (size_t)N * H * W * 4, cudaMemcpyHostToDevice);
    printf("mode %d (1 = descriptor in global memory, 2 = non-negative coords): ", gMode);
    if (which == 0) run<3, 128, 32>(fn, d, N, H, W, out, "3d 128x32");
    if (which == 1) run<3, 132, 35>(fn, d, N, H, W, out, "3d 132x35");
    if (which == 2) run<2, 128, 32>(fn, d, N, H, W, out, "2d 128x32");
    if (which == 3) run<2, 64, 16>(fn, d, N, H, W, out, "2d 64x16");
    if (which == 4) run<3, 32, 8>(fn, d, N, H, W, out, "3d 32x8");
    return 0;
}
