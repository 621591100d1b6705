// Entry points not implemented yet (replaced as their kernels land).
#include "common.cuh"
extern "C" size_t mte_chamfer_workspace_bytes(int, int, int) { return 0; }
extern "C" int mte_chamfer_counts(const uint8_t *, const uint8_t *, int, int, int, double, double *, void *, size_t, mte_stream_t) { return MTE_ERR_ARG; }
