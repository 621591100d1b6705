// Entry points not implemented yet (replaced as their kernels land).
#include "common.cuh"
extern "C" size_t mte_dee_workspace_bytes(int, int, int) { return 0; }
extern "C" int mte_dee_postprocess(const void *, int, int, int, int, int, int, double, double, uint8_t *, void *, int, void *, size_t, mte_stream_t) { return MTE_ERR_ARG; }
extern "C" size_t mte_pr_workspace_bytes(int, int, int, int, double) { return 0; }
extern "C" int mte_pr_counts(const void *, int, const uint8_t *, int, int, int, const int32_t *, const double *, int, double, int, int64_t *, void *, size_t, mte_stream_t) { return MTE_ERR_ARG; }
extern "C" size_t mte_match_workspace_bytes(int, int, int, double) { return 0; }
extern "C" int mte_correspond_pixels(const uint8_t *, const uint8_t *, int, int, int, double, uint8_t *, uint8_t *, int64_t *, void *, size_t, mte_stream_t) { return MTE_ERR_ARG; }
extern "C" size_t mte_thin_workspace_bytes(int, int, int) { return 0; }
extern "C" int mte_binary_thin(const uint8_t *, uint8_t *, int, int, int, int, void *, size_t, mte_stream_t) { return MTE_ERR_ARG; }
extern "C" size_t mte_chamfer_workspace_bytes(int, int, int) { return 0; }
extern "C" int mte_chamfer_counts(const uint8_t *, const uint8_t *, int, int, int, double, double *, void *, size_t, mte_stream_t) { return MTE_ERR_ARG; }
