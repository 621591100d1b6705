// Version / error strings / workspace header helper of the C ABI (include/mte.h).
#include "common.cuh"

extern "C" int mte_version(void) { return MTE_VERSION; }

extern "C" const char *mte_error_string(int code) {
    switch (code) {
        case MTE_OK: return "ok";
        case MTE_ERR_NULL: return "required pointer is NULL";
        case MTE_ERR_SHAPE: return "non-positive or inconsistent shape";
        case MTE_ERR_WORKSPACE: return "workspace too small";
        case MTE_ERR_ARG: return "attribute out of range";
        case MTE_ERR_ALIGN: return "pointer not aligned to its element type";
        case MTE_ERR_NOT_NESTED: return "threshold list is not nested (strictest first)";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

extern "C" int mte_workspace_init(void *workspace, size_t bytes, mte_stream_t stream) {
    if (!workspace) return MTE_ERR_NULL;
    if (bytes < MTE_WS_HEADER_BYTES) return MTE_ERR_WORKSPACE;
    cudaError_t e = cudaMemsetAsync(workspace, 0, MTE_WS_HEADER_BYTES, reinterpret_cast<cudaStream_t>(stream));
    return e == cudaSuccess ? MTE_OK : (int)e;
}
