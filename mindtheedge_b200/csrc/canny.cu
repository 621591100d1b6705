// Depth -> edge extraction for evaluation (sm_100a), bit-exact with the reference path
//   edge_from_depth            edge.py:73-93  (twin packnet_code/packnet_sfm/utils/edge.py:64-89)
//   cv2.Canny(u8, low, high)   third-party call at edge.py:88 (aperture 3, L1 gradient), SURVEY.md A.2
//
// Three kernels, all integer work after the quantisation:
//  1. canny_nms_kernel   depth -> u8 (clamp, *255/max, trunc) -> Sobel3 (BORDER_REPLICATE) -> L1 magnitude ->
//                        fixed-point NMS, staged through a shared-memory halo tile; emits per pixel the FIRST
//                        threshold index at which it is a candidate (cl) and at which it is strong (sl).
//  2. canny_hyst_kernel  multi-threshold hysteresis as one monotone relaxation
//                            E(p) = max(cl(p), min(E(p), min_{n in N8(p)} E(n))),  E initialised to sl,
//                        so E(p) <= t  <=>  p is an edge of cv2.Canny at pair t, for ALL nested pairs at once
//                        (one u8 plane instead of T).  Tile-local fix-point in shared memory, tiles re-queued
//                        through dirty flags, grid-wide iterations inside ONE cooperative launch (no host sync).
//  3. canny_expand_kernel  optional: the T {0,255} planes the reference API returns.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mte {
namespace canny {

constexpr int TG22 = 13573;  // round(tan(22.5 deg) * 2^15)
constexpr int TW = 128, TH = 32, kThreads = 256;
constexpr int kNever = 255;

struct Thresholds {
    int n;
    int low[MTE_MAX_THRESHOLDS];
    int high[MTE_MAX_THRESHOLDS];
};

template <typename T>
__device__ __forceinline__ unsigned char quantise(T d, T lo, T hi, T factor) {
    // edge.py:81-87 (the comparisons leave NaN alone; the uint8 cast of NaN is 0 on x86 and here)
    if (d < lo) d = lo;
    if (d > hi) d = hi;
    const T v = d * factor;
    return (unsigned char)(int)v;
}
template <>
__device__ __forceinline__ unsigned char quantise<unsigned char>(unsigned char d, unsigned char, unsigned char,
                                                                 unsigned char) {
    return d;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) canny_nms_kernel(const T *__restrict__ depth, int N, int H, int W, T lo,
                                                             T hi, T factor, const __grid_constant__ Thresholds thr,
                                                             unsigned char *__restrict__ cl,
                                                             unsigned char *__restrict__ E) {
    __shared__ unsigned char q[TH + 4][TW + 4];
    __shared__ unsigned short mag[TH + 2][TW + 4];
    const int tilesX = ceil_div(W, TW), tilesY = ceil_div(H, TH);
    const int tile = blockIdx.x % (tilesX * tilesY), img = blockIdx.x / (tilesX * tilesY);
    const int x0 = (tile % tilesX) * TW, y0 = (tile / tilesX) * TH;
    const T *src = depth + (size_t)img * H * W;

    // quantised image, halo 2, coordinates clamped (BORDER_REPLICATE)
    for (int i = threadIdx.x; i < (TH + 4) * (TW + 4); i += kThreads) {
        const int r = i / (TW + 4), c = i - r * (TW + 4);
        const int y = min(max(y0 + r - 2, 0), H - 1), x = min(max(x0 + c - 2, 0), W - 1);
        q[r][c] = quantise<T>(src[(size_t)y * W + x], lo, hi, factor);
    }
    __syncthreads();
    // L1 magnitude, halo 1; outside the image the magnitude is 0
    for (int i = threadIdx.x; i < (TH + 2) * (TW + 2); i += kThreads) {
        const int r = i / (TW + 2), c = i - r * (TW + 2);
        const int y = y0 + r - 1, x = x0 + c - 1;
        int m = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const int a = q[r][c], b = q[r][c + 1], cc = q[r][c + 2];
            const int d = q[r + 1][c], f = q[r + 1][c + 2];
            const int g = q[r + 2][c], h = q[r + 2][c + 1], k = q[r + 2][c + 2];
            const int dx = (cc + 2 * f + k) - (a + 2 * d + g);
            const int dy = (g + 2 * h + k) - (a + 2 * b + cc);
            m = abs(dx) + abs(dy);
        }
        mag[r][c] = (unsigned short)m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
        const int r = i / TW, c = i - r * TW;
        const int y = y0 + r, x = x0 + c;
        if (y >= H || x >= W) continue;
        const int a = q[r + 1][c + 1], b = q[r + 1][c + 2], cc = q[r + 1][c + 3];
        const int d = q[r + 2][c + 1], f = q[r + 2][c + 3];
        const int g = q[r + 3][c + 1], h = q[r + 3][c + 2], k = q[r + 3][c + 3];
        const int dx = (cc + 2 * f + k) - (a + 2 * d + g);
        const int dy = (g + 2 * h + k) - (a + 2 * b + cc);
        const int m = mag[r + 1][c + 1];
        const int ax = abs(dx), ay = abs(dy) << 15;
        const int tg22x = ax * TG22;
        const int tg67x = tg22x + (ax << 16);
        bool keep;
        if (ay < tg22x) {
            keep = m > mag[r + 1][c] && m >= mag[r + 1][c + 2];
        } else if (ay > tg67x) {
            keep = m > mag[r][c + 1] && m >= mag[r + 2][c + 1];
        } else {
            const int s = ((dx ^ dy) < 0) ? -1 : 1;
            keep = m > mag[r][c + 1 - s] && m > mag[r + 2][c + 1 + s];
        }
        const int v = keep ? m : 0;
        int first_c = kNever, first_s = kNever;
        for (int t = thr.n - 1; t >= 0; t--) {  // nested pairs: the last hit going down is the first index
            if (v > thr.low[t]) first_c = t;
            if (v > thr.high[t]) first_s = t;
        }
        const size_t o = (size_t)img * H * W + (size_t)y * W + x;
        cl[o] = (unsigned char)first_c;
        E[o] = (unsigned char)first_s;
    }
}

struct HystP {
    const unsigned char *cl;
    unsigned char *E;
    int N, H, W, tilesX, tilesY, nTiles;
    unsigned char *active;   // [2][nTiles]
    unsigned int *counters;  // [2] number of tiles queued for iteration parity
};

__global__ void __launch_bounds__(kThreads) canny_hyst_kernel(const HystP P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned char sE[TH + 2][TW + 2];
    __shared__ unsigned char sC[TH][TW];
    __shared__ int sChanged, sBorder, sGo;
    const int perImg = P.tilesX * P.tilesY;
    for (int it = 0;; it++) {
        unsigned char *cur = P.active + (size_t)(it & 1) * P.nTiles;
        unsigned char *nxt = P.active + (size_t)((it + 1) & 1) * P.nTiles;
        for (int tile = blockIdx.x; tile < P.nTiles; tile += gridDim.x) {
            if (threadIdx.x == 0) {
                sGo = (it == 0) || cur[tile];
                cur[tile] = 0;
            }
            __syncthreads();
            if (!sGo) {
                __syncthreads();
                continue;
            }
            const int img = tile / perImg, tl = tile - img * perImg;
            const int tx = tl % P.tilesX, ty = tl / P.tilesX;
            const int x0 = tx * TW, y0 = ty * TH;
            const size_t base = (size_t)img * P.H * P.W;
            for (int i = threadIdx.x; i < (TH + 2) * (TW + 2); i += kThreads) {
                const int r = i / (TW + 2), c = i - r * (TW + 2);
                const int y = y0 + r - 1, x = x0 + c - 1;
                const bool in = y >= 0 && y < P.H && x >= 0 && x < P.W;
                sE[r][c] = in ? __ldcg(P.E + base + (size_t)y * P.W + x) : (unsigned char)kNever;
            }
            for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
                const int r = i / TW, c = i - r * TW;
                const int y = y0 + r, x = x0 + c;
                sC[r][c] = (y < P.H && x < P.W) ? P.cl[base + (size_t)y * P.W + x] : (unsigned char)kNever;
            }
            if (threadIdx.x == 0) sBorder = 0;
            __syncthreads();
            // tile-local fix-point (in place: any value read is a valid upper bound, the map is monotone)
            bool anyChange = false;
            for (;;) {
                if (threadIdx.x == 0) sChanged = 0;
                __syncthreads();
                bool ch = false;
                for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
                    const int r = i / TW, c = i - r * TW;
                    const int cur_e = sE[r + 1][c + 1], lim = sC[r][c];
                    if (cur_e <= lim) continue;  // already at its candidate level (or never a candidate)
                    int m = min(min(sE[r][c], sE[r][c + 1]), sE[r][c + 2]);
                    m = min(m, min(sE[r + 1][c], sE[r + 1][c + 2]));
                    m = min(m, min(min(sE[r + 2][c], sE[r + 2][c + 1]), sE[r + 2][c + 2]));
                    const int ne = max(lim, min(cur_e, m));
                    if (ne < cur_e) {
                        sE[r + 1][c + 1] = (unsigned char)ne;
                        ch = true;
                        if (r == 0 || c == 0 || r == TH - 1 || c == TW - 1) sBorder = 1;
                    }
                }
                if (ch) sChanged = 1;
                __syncthreads();
                if (!sChanged) break;
                anyChange = true;
                __syncthreads();
            }
            if (anyChange) {
                for (int i = threadIdx.x; i < TH * TW; i += kThreads) {
                    const int r = i / TW, c = i - r * TW;
                    const int y = y0 + r, x = x0 + c;
                    if (y < P.H && x < P.W) __stcg(P.E + base + (size_t)y * P.W + x, sE[r + 1][c + 1]);
                }
                if (sBorder && threadIdx.x < 9 && threadIdx.x != 4) {
                    const int nx = tx + (int)(threadIdx.x % 3) - 1, ny = ty + (int)(threadIdx.x / 3) - 1;
                    if (nx >= 0 && nx < P.tilesX && ny >= 0 && ny < P.tilesY) {
                        const int nt = img * perImg + ny * P.tilesX + nx;
                        if (!nxt[nt]) {
                            nxt[nt] = 1;
                            atomicAdd(P.counters + ((it + 1) & 1), 1u);
                        }
                    }
                }
            }
            __syncthreads();
        }
        __threadfence();
        grid.sync();
        const unsigned pending = *(volatile unsigned int *)(P.counters + ((it + 1) & 1));
        grid.sync();
        if (blockIdx.x == 0 && threadIdx.x == 0) P.counters[(it + 1) & 1] = 0;  // consumed; clean for reuse
        if (pending == 0) break;
    }
}

__global__ void canny_expand_kernel(const unsigned char *__restrict__ E, unsigned char *__restrict__ edges, size_t n,
                                    int T) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int e = E[i];
        for (int t = 0; t < T; t++) edges[(size_t)t * n + i] = (e <= t) ? 255 : 0;
    }
}

struct Layout {
    size_t offCl, offE, offActive, total;
    int nTiles, tilesX, tilesY;
};

static Layout layout(int N, int H, int W) {
    Layout L;
    L.tilesX = ceil_div(W, TW);
    L.tilesY = ceil_div(H, TH);
    L.nTiles = L.tilesX * L.tilesY * N;
    size_t off = MTE_WS_HEADER_BYTES;
    const size_t plane = align_up((size_t)N * H * W, 256);
    L.offCl = off; off += plane;
    L.offE = off; off += plane;
    L.offActive = off; off += align_up((size_t)2 * L.nTiles, 256);
    L.total = off;
    return L;
}

// Shared with the DEE post-process (dee.cu): relax E towards cl over 8-connected paths.
int run_level_hysteresis(const unsigned char *cl, unsigned char *E, int N, int H, int W, unsigned char *active,
                         unsigned int *counters, cudaStream_t st) {
    HystP P;
    P.cl = cl; P.E = E; P.N = N; P.H = H; P.W = W;
    P.tilesX = ceil_div(W, TW); P.tilesY = ceil_div(H, TH); P.nTiles = P.tilesX * P.tilesY * N;
    P.active = active;
    P.counters = counters;
    int dev = 0, sms = kNumSMs, perSm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, canny_hyst_kernel, kThreads, 0);
    int grid = sms * (perSm < 1 ? 1 : perSm);
    if (grid > P.nTiles) grid = P.nTiles;
    cudaError_t e = cudaMemsetAsync(P.active, 0, (size_t)2 * P.nTiles, st);
    if (e != cudaSuccess) return (int)e;
    void *args[] = {&P};
    e = cudaLaunchCooperativeKernel((const void *)canny_hyst_kernel, dim3(grid), dim3(kThreads), args, 0, st);
    return e == cudaSuccess ? MTE_OK : (int)e;
}
size_t hysteresis_active_bytes(int N, int H, int W) {
    return align_up((size_t)2 * ceil_div(W, TW) * ceil_div(H, TH) * N, 256);
}

static int run_pairs(const void *depth, int dtype, int N, int H, int W, double min_depth, double max_depth,
                     const Thresholds &thr, unsigned char *edges, unsigned char *levels, char *ws, const Layout &L,
                     cudaStream_t st) {
    unsigned char *cl = reinterpret_cast<unsigned char *>(ws + L.offCl);
    unsigned char *E = levels ? levels : reinterpret_cast<unsigned char *>(ws + L.offE);
    const int grid1 = L.nTiles;
    const double factor = 255.0 / max_depth;
    if (dtype == MTE_F32)
        canny_nms_kernel<float><<<grid1, kThreads, 0, st>>>(static_cast<const float *>(depth), N, H, W, (float)min_depth,
                                                           (float)max_depth, (float)factor, thr, cl, E);
    else if (dtype == MTE_F64)
        canny_nms_kernel<double><<<grid1, kThreads, 0, st>>>(static_cast<const double *>(depth), N, H, W, min_depth,
                                                            max_depth, factor, thr, cl, E);
    else
        canny_nms_kernel<unsigned char><<<grid1, kThreads, 0, st>>>(static_cast<const unsigned char *>(depth), N, H, W,
                                                                   0, 0, 0, thr, cl, E);
    MTE_RETURN_IF_CUDA_ERROR();

    int rc = run_level_hysteresis(cl, E, N, H, W, reinterpret_cast<unsigned char *>(ws + L.offActive),
                                  reinterpret_cast<WsHeader *>(ws)->flag, st);
    if (rc) return rc;
    int sms = kNumSMs;
    if (edges) {
        const size_t n = (size_t)N * H * W;
        const int g = (int)((n + 255) / 256 < (size_t)sms * 16 ? (n + 255) / 256 : (size_t)sms * 16);
        canny_expand_kernel<<<g, 256, 0, st>>>(E, edges, n, thr.n);
        MTE_RETURN_IF_CUDA_ERROR();
    }
    return MTE_OK;
}

}  // namespace canny
}  // namespace mte

using namespace mte;
using namespace mte::canny;

extern "C" size_t mte_canny_workspace_bytes(int N, int H, int W, int T) {
    if (N < 1 || H < 1 || W < 1 || T < 1) return 0;
    return layout(N, H, W).total;
}

extern "C" int mte_canny_from_depth(const void *depth, int dtype, int N, int H, int W, double min_depth,
                                    double max_depth, const int32_t *lows, const int32_t *highs, int T,
                                    uint8_t *edges, uint8_t *levels, void *workspace, size_t ws_bytes,
                                    mte_stream_t stream) {
    if (!depth || !lows || !highs || !workspace) return MTE_ERR_NULL;
    if (!edges && !levels) return MTE_ERR_NULL;
    if (N < 1 || H < 1 || W < 1) return MTE_ERR_SHAPE;
    if (T < 1 || T > MTE_MAX_THRESHOLDS) return MTE_ERR_ARG;
    if (dtype != MTE_F32 && dtype != MTE_F64 && dtype != MTE_U8) return MTE_ERR_ARG;
    if (dtype != MTE_U8 && !(max_depth != 0.0)) return MTE_ERR_ARG;
    const Layout L = layout(N, H, W);
    if (ws_bytes < L.total) return MTE_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char *ws = static_cast<char *>(workspace);

    Thresholds thr;
    thr.n = T;
    bool nested = true;
    for (int t = 0; t < T; t++) {
        int lo = lows[t], hi = highs[t];
        if (lo > hi) { const int s = lo; lo = hi; hi = s; }  // cv2.Canny swaps
        thr.low[t] = lo;
        thr.high[t] = hi;
        if (t > 0 && (thr.low[t] > thr.low[t - 1] || thr.high[t] > thr.high[t - 1])) nested = false;
    }
    if (nested) return run_pairs(depth, dtype, N, H, W, min_depth, max_depth, thr, edges, levels, ws, L, st);
    // arbitrary pair lists: one independent pass per pair (a level plane only exists for nested lists)
    if (levels) return MTE_ERR_NOT_NESTED;
    const size_t n = (size_t)N * H * W;
    for (int t = 0; t < T; t++) {
        Thresholds one;
        one.n = 1; one.low[0] = thr.low[t]; one.high[0] = thr.high[t];
        int rc = run_pairs(depth, dtype, N, H, W, min_depth, max_depth, one, edges + (size_t)t * n, nullptr, ws, L, st);
        if (rc) return rc;
    }
    return MTE_OK;
}
