// Training edge loss, forward and backward, for sm_100a.
//
// Replaces (reference paths relative to the reference root):
//   GradLayer.forward          packnet_code/packnet_sfm/losses/grad_loss.py:65-95
//   GradLoss.forward           packnet_code/packnet_sfm/losses/grad_loss.py:122-159
//   GradLoss.comp_cross_entropy packnet_code/packnet_sfm/losses/grad_loss.py:161-219
//   the per-scale loop         packnet_code/packnet_sfm/models/SemiSupEdgeModel.py:164-198
//   inv2depth (optional fuse)  packnet_code/packnet_sfm/utils/depth.py:104-121
// and the autograd backward of that chain.
//
// Design (HBM-bound stencil + reduction, no tensor cores):
//  * one launch covers every image of up to 4 pyramid scales;
//  * a warp owns a 128-px wide strip (32 lanes x float4, 128-bit coalesced loads),
//    horizontal neighbours come from warp shuffles, vertical ones from registers;
//  * forward: 3x3 directional responses -> pick by quantised normal -> sigmoid ->
//    soft-label BCE terms, per-thread fp32 partials -> warp shuffle -> one fp64
//    partial row per CTA -> the last CTA (atomic ticket) folds all partials in a
//    fixed order, computes the per-image class balance alpha and writes the loss:
//    single launch, no host sync, bit-reproducible run to run;
//  * backward: recomputes the per-pixel coefficient s(n) from depth/edge/normal
//    (16 B/px of traffic, nothing stashed by the forward except alpha) on a strip
//    that overlaps its neighbours by one lane each side, then gathers the 3x3
//    adjoint from registers + shuffles.
#include <math.h>

#include "common.cuh"

namespace mte {
namespace loss {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kAcc = 8;
enum { A_WP = 0, A_WN, A_SPU, A_SNU, A_SPM, A_SNM, A_SUMM, A_FLAGS };
enum { F_HAS0 = 1u, F_HAS1 = 2u, F_OTHER = 4u };
enum { MODE_NONE = 0, MODE_MAG = 1, MODE_DIR = 2 };

constexpr float kEps = 0.001f;        // grad_loss.py:167,180
constexpr float kLn2 = 0.6931471805599453f;

struct ScaleP {
    const float *x, *e, *n, *m;
    float *g, *dx;
    int B, H, W;
    int strips, rowBlocks, items, ctasPerImage;
    int ctaBase, imgBase;
    float scaleWeight;
};

struct LossP {
    ScaleP s[MTE_MAX_SCALES];
    int nScales, totalCtas, totalImages;
    int isSigmoid, predInv;
    float T, weight, p2n;
    double *partials;      // [totalCtas][kAcc]
    double *segsums;       // [totalImages][kAcc]
    unsigned *ticket;
    float *lossOut;        // [1+nScales]
    float *ctx;            // [totalImages] alpha, then per scale {coef, maskBinary}
    const float *gradLoss; // bwd: [1+nScales]
};

// fp32-rounded k*pi/8, exactly the constants torch compares against (grad_loss.py:80-93)
#define MTE_B1 ((float)(1 * M_PI / 8))
#define MTE_B3 ((float)(3 * M_PI / 8))
#define MTE_B5 ((float)(5 * M_PI / 8))
#define MTE_B7 ((float)(7 * M_PI / 8))

// 0:h 1:rl 2:v 3:lr  (NaN -> 0 -> h, as the reference's untouched default)
__device__ __forceinline__ int dir_index(float t) {
    int idx = (t >= -MTE_B7) + (t >= -MTE_B5) + (t >= -MTE_B3) + (t >= -MTE_B1) + (t >= MTE_B1) + (t >= MTE_B3) +
              (t >= MTE_B5) + (t >= MTE_B7);
    return idx & 3;
}

template <int VEC>
struct Row {
    float c[VEC];
    float l, r;
    __device__ __forceinline__ float at(int v) const { return v < 0 ? l : (v >= VEC ? r : c[v]); }
};

__device__ __forceinline__ float inv_to_depth(float v) { return 1.0f / fmaxf(v, 1e-6f); }

// Load one row of the strip: centre values by a 128-bit load, the two outer
// neighbours by shuffle (lanes 0/31 fetch theirs from memory).  Out-of-image
// reads give 0 (the zero padding of F.conv2d(padding=1)).
template <int VEC>
__device__ __forceinline__ void load_row(Row<VEC> &R, const float *img, int row, int H, int W, int col0, int lane,
                                         bool predInv) {
    const bool rowOk = (row >= 0) && (row < H);
    const float *p = img + (size_t)(rowOk ? row : 0) * W;
    if (VEC == 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rowOk && col0 >= 0 && col0 < W) v = ld_cached4(p + col0);
        R.c[0] = v.x; R.c[1 % VEC] = v.y; R.c[2 % VEC] = v.z; R.c[3 % VEC] = v.w;
    } else {
        R.c[0] = (rowOk && col0 >= 0 && col0 < W) ? __ldg(p + col0) : 0.f;
    }
    if (predInv) {
        if (rowOk && col0 >= 0 && col0 < W) {
#pragma unroll
            for (int v = 0; v < VEC; v++) R.c[v] = inv_to_depth(R.c[v]);
        }
    }
    float l = __shfl_up_sync(MTE_FULL_MASK, R.c[VEC - 1], 1);
    float r = __shfl_down_sync(MTE_FULL_MASK, R.c[0], 1);
    if (lane == 0) {
        const int c = col0 - 1;
        l = (rowOk && c >= 0 && c < W) ? __ldg(p + c) : 0.f;
        if (predInv && rowOk && c >= 0 && c < W) l = inv_to_depth(l);
    }
    if (lane == 31) {
        const int c = col0 + VEC;
        r = (rowOk && c >= 0 && c < W) ? __ldg(p + c) : 0.f;
        if (predInv && rowOk && c >= 0 && c < W) r = inv_to_depth(r);
    }
    R.l = l;
    R.r = r;
}

template <int VEC>
__device__ __forceinline__ void load_plane_row(float (&out)[VEC], const float *img, int row, int H, int W, int col0,
                                               float fill) {
    const bool ok = (row >= 0) && (row < H) && (col0 >= 0) && (col0 < W);
    if (VEC == 4) {
        float4 v = make_float4(fill, fill, fill, fill);
        if (ok) v = ld_stream4(img + (size_t)row * W + col0);
        out[0] = v.x; out[1 % VEC] = v.y; out[2 % VEC] = v.z; out[3 % VEC] = v.w;
    } else {
        out[0] = ok ? __ldcs(img + (size_t)row * W + col0) : fill;
    }
}

// The four zero-padded 3x3 cross-correlations at column v of the middle row
// (grad_loss.py:20-31 written through their separable parts).
template <int VEC>
__device__ __forceinline__ void responses(const Row<VEC> &up, const Row<VEC> &mid, const Row<VEC> &dn, int v,
                                          float &cv, float &ch, float &clr, float &crl) {
    const float tl = up.at(v - 1), tc = up.at(v), tr = up.at(v + 1);
    const float ml = mid.at(v - 1), mr = mid.at(v + 1);
    const float bl = dn.at(v - 1), bc = dn.at(v), br = dn.at(v + 1);
    const float P = (bl + bc + br) - (tl + tc + tr);
    const float Dm = mr - ml;
    const float R = (tr - tl) + Dm + (br - bl);
    cv = P + (bc - tc);
    ch = R + Dm;
    clr = P + R;
    crl = R - P;
}

// Forward: the loss is a 2M-term average, MUFU-accuracy sigmoid is far inside 1e-5.
__device__ __forceinline__ float sigmoidf_fast(float z) { return __frcp_rn(1.0f + __expf(-z)); }
// Backward: p*(1-p)/(1-p+eps) amplifies the last bits of p where the sigmoid saturates, so p is computed
// the way eager PyTorch does (accurate expf, IEEE divide) to reproduce the reference gradient, not just the math.
__device__ __forceinline__ float sigmoidf_ref(float z) { return 1.0f / (1.0f + expf(-z)); }

// ---------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------
__device__ __noinline__ void finalize_loss(const LossP &P, bool hasMask) {
    // Called by every thread of the LAST CTA.  Fixed summation order.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int si = 0; si < P.nScales; si++) {
        const ScaleP &S = P.s[si];
        for (int b = warp; b < S.B; b += kWarps) {
            double a[kAcc];
            unsigned fl = 0;
#pragma unroll
            for (int k = 0; k < kAcc; k++) a[k] = 0.0;
            const double *base = P.partials + (size_t)(S.ctaBase + b * S.ctasPerImage) * kAcc;
            for (int c = lane; c < S.ctasPerImage; c += 32) {
                const double *q = base + (size_t)c * kAcc;
#pragma unroll
                for (int k = 0; k < kAcc - 1; k++) a[k] += __ldcg(q + k);
                fl |= (unsigned)__ldcg(q + A_FLAGS);
            }
#pragma unroll
            for (int k = 0; k < kAcc - 1; k++) a[k] = warp_sum(a[k]);
            fl = warp_or(fl);
            if (lane == 0) {
                double *o = P.segsums + (size_t)(S.imgBase + b) * kAcc;
#pragma unroll
                for (int k = 0; k < kAcc - 1; k++) o[k] = a[k];
                o[A_FLAGS] = (double)fl;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int si = 0; si < P.nScales; si++) {
            const ScaleP &S = P.s[si];
            const double npix = (double)S.H * (double)S.W;
            unsigned fl = 0;
            double wnAll = 0.0, sumM = 0.0;
            for (int b = 0; b < S.B; b++) {
                const double *o = P.segsums + (size_t)(S.imgBase + b) * kAcc;
                fl |= (unsigned)o[A_FLAGS];
                wnAll += hasMask ? o[A_WN] : (npix - o[A_WP]);
                sumM += o[A_SUMM];
            }
            const bool binary = hasMask && fl == (F_HAS0 | F_HAS1);
            const double valid = binary ? sumM : npix * (double)S.B;
            double acc = 0.0;
            for (int b = 0; b < S.B; b++) {
                const double *o = P.segsums + (size_t)(S.imgBase + b) * kAcc;
                const double wp = o[A_WP];
                const double wn = hasMask ? o[A_WN] : (npix - wp);
                const float alpha = (wnAll == 0.0) ? 1.0f : (float)(wn / (wp + wn));
                const double sp = (double)kLn2 * (binary ? o[A_SPM] : o[A_SPU]);
                const double sn = (double)kLn2 * (binary ? o[A_SNM] : o[A_SNU]);
                acc += -(double)P.p2n * (double)alpha * sp - (1.0 - (double)alpha) * sn;
                P.ctx[S.imgBase + b] = alpha;
            }
            const double lossS = (double)P.weight * (acc / valid);
            P.lossOut[1 + si] = (float)lossS;
            P.ctx[P.totalImages + 2 * si] = (float)((double)P.weight / valid);
            P.ctx[P.totalImages + 2 * si + 1] = binary ? 1.0f : 0.0f;
            total += (double)S.scaleWeight * lossS;
        }
        P.lossOut[0] = (float)total;
        *P.ticket = 0u;  // leave the workspace header clean for the next launch
    }
}

template <int VEC, int MODE, bool MASK, int RH>
__global__ void __launch_bounds__(kThreads) edge_loss_fwd_kernel(const __grid_constant__ LossP P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int si = 0;
#pragma unroll
    for (int k = 1; k < MTE_MAX_SCALES; k++)
        if (k < P.nScales && (int)blockIdx.x >= P.s[k].ctaBase) si = k;
    const ScaleP &S = P.s[si];
    const int local = blockIdx.x - S.ctaBase;
    const int img = local / S.ctasPerImage;
    const int item = (local - img * S.ctasPerImage) * kWarps + warp;

    float acc[kAcc - 1];
#pragma unroll
    for (int k = 0; k < kAcc - 1; k++) acc[k] = 0.f;
    unsigned flags = 0;

    if (item < S.items) {
        const int H = S.H, W = S.W;
        const int strip = item / S.rowBlocks;  // vertically adjacent row blocks share a CTA (L1 halo reuse)
        const int rb = item - strip * S.rowBlocks;
        const int row0 = rb * RH;
        const int col0 = (strip * 32 + lane) * VEC;
        const size_t plane = (size_t)img * H * W;
        const float *x = S.x + plane;
        const bool colOk = col0 < W;

        Row<VEC> rows[(MODE == MODE_NONE) ? RH : RH + 2];
        if (MODE == MODE_NONE) {
#pragma unroll
            for (int r = 0; r < RH; r++) {
                load_plane_row<VEC>(rows[r].c, x, row0 + r, H, W, col0, 0.f);
                if (P.predInv && colOk && row0 + r < H) {
#pragma unroll
                    for (int v = 0; v < VEC; v++) rows[r].c[v] = inv_to_depth(rows[r].c[v]);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < RH + 2; r++) load_row<VEC>(rows[r], x, row0 - 1 + r, H, W, col0, lane, P.predInv != 0);
        }
        float e[RH][VEC], th[RH][VEC], m[RH][VEC];
#pragma unroll
        for (int r = 0; r < RH; r++) {
            load_plane_row<VEC>(e[r], S.e + plane, row0 + r, H, W, col0, 0.f);
            if (MODE == MODE_DIR) load_plane_row<VEC>(th[r], S.n + plane, row0 + r, H, W, col0, 0.f);
            if (MASK) load_plane_row<VEC>(m[r], S.m + plane, row0 + r, H, W, col0, 0.f);
        }
#pragma unroll
        for (int r = 0; r < RH; r++) {
            const int row = row0 + r;
            const bool ok = colOk && row < H;
            float g[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                if (MODE == MODE_NONE) {
                    g[v] = rows[r].c[v];
                } else {
                    float cv, ch, clr, crl;
                    responses<VEC>(rows[r], rows[r + 1], rows[r + 2], v, cv, ch, clr, crl);
                    if (MODE == MODE_MAG) {
                        g[v] = sqrtf(cv * cv + ch * ch + 1e-6f);
                    } else {
                        const int k = dir_index(th[r][v]);
                        const float c = (k == 0) ? ch : (k == 1) ? crl : (k == 2) ? cv : clr;
                        g[v] = fabsf(c);
                    }
                }
                const float p = P.isSigmoid ? sigmoidf_fast(g[v] - P.T) : g[v];
                const float ee = e[r][v];
                const float lp = __log2f(p + kEps);
                const float ln = __log2f((1.0f - p) + kEps);
                if (ok) {
                    const float ne = 1.0f - ee;
                    acc[A_SPU] += ee * lp;
                    acc[A_SNU] += ne * ln;
                    if (MASK) {
                        const float mm = m[r][v];
                        acc[A_WP] += ee * mm;
                        acc[A_WN] += ne * mm;
                        acc[A_SUMM] += mm;
                        if (mm != 0.f) {
                            acc[A_SPM] += ee * lp;
                            acc[A_SNM] += ne * ln;
                        }
                        flags |= (mm == 0.f) ? F_HAS0 : ((mm == 1.f) ? F_HAS1 : F_OTHER);
                    } else {
                        acc[A_WP] += ee;
                    }
                }
            }
            if (S.g != nullptr && ok) {
                float *gp = S.g + plane + (size_t)row * W + col0;
                if (VEC == 4) st_stream4(gp, make_float4(g[0], g[1 % VEC], g[2 % VEC], g[3 % VEC]));
                else __stcs(gp, g[0]);
            }
        }
    }

    // warp -> CTA -> one fp64 partial row per CTA
    __shared__ float sAcc[kWarps][kAcc];
#pragma unroll
    for (int k = 0; k < kAcc - 1; k++)
        if (MASK || k == A_WP || k == A_SPU || k == A_SNU) acc[k] = warp_sum(acc[k]);
    if (MASK) flags = warp_or(flags);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kAcc - 1; k++) sAcc[warp][k] = acc[k];
        sAcc[warp][A_FLAGS] = __uint_as_float(flags);
    }
    __syncthreads();
    __shared__ bool sLast;
    if (threadIdx.x < kAcc) {
        const int k = threadIdx.x;
        double v;
        if (k == A_FLAGS) {
            unsigned f = 0;
            for (int w = 0; w < kWarps; w++) f |= __float_as_uint(sAcc[w][k]);
            v = (double)f;
        } else {
            v = 0.0;
            for (int w = 0; w < kWarps; w++) v += (double)sAcc[w][k];
        }
        __stcg(P.partials + (size_t)blockIdx.x * kAcc + k, v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(P.ticket, 1u);
        sLast = (t == (unsigned)P.totalCtas - 1u);
    }
    __syncthreads();
    if (sLast) {
        __threadfence();
        finalize_loss(P, MASK);
    }
}

// ---------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------
// Per-pixel adjoint coefficients: contribution of response pixel n to its 3x3
// neighbourhood is  U/V on the diagonals, A2 above/below, C2 left/right
// (derived from K_d[a][b] = alpha*a*(1+beta*(1-|b|)) + gamma*b*(1+beta*(1-|a|))).
struct Coef {
    float U, V, A2, C2;
};

template <int VEC>
struct CoefRow {
    Coef c[VEC];
    Coef l, r;
    __device__ __forceinline__ const Coef &at(int v) const { return v < 0 ? l : (v >= VEC ? r : c[v]); }
};

struct BwdImg {
    float cp, cn;     // -G*coef*lambda*alpha , G*coef*(1-alpha)
    bool maskBinary;
};

template <int MODE, bool MASK>
__device__ __forceinline__ float dloss_dg(float g, float ee, float mm, const BwdImg &I, bool isSigmoid, float T) {
    const float p = isSigmoid ? sigmoidf_ref(g - T) : g;
    float d = I.cp * ee * __frcp_rn(p + kEps) + I.cn * (1.0f - ee) * __frcp_rn((1.0f - p) + kEps);
    if (MASK) {
        if (I.maskBinary && mm == 0.f) d = 0.f;
    }
    return isSigmoid ? d * p * (1.0f - p) : d;
}

template <int VEC, int MODE, bool MASK, int RH>
__global__ void __launch_bounds__(kThreads) edge_loss_bwd_kernel(const __grid_constant__ LossP P) {
    static_assert(MODE != MODE_NONE, "pointwise backward has its own kernel");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int si = 0;
#pragma unroll
    for (int k = 1; k < MTE_MAX_SCALES; k++)
        if (k < P.nScales && (int)blockIdx.x >= P.s[k].ctaBase) si = k;
    const ScaleP &S = P.s[si];
    const int local = blockIdx.x - S.ctaBase;
    const int img = local / S.ctasPerImage;
    const int item = (local - img * S.ctasPerImage) * kWarps + warp;
    if (item >= S.items) return;

    const int H = S.H, W = S.W;
    const int strip = item / S.rowBlocks;
    const int rb = item - strip * S.rowBlocks;
    const int row0 = rb * RH;
    // 30 writing lanes per strip; lanes 0 and 31 only supply the halo coefficients
    const int col0 = (strip * 30 + lane - 1) * VEC;
    const size_t plane = (size_t)img * H * W;
    const float *x = S.x + plane;
    const bool colOk = col0 >= 0 && col0 < W;

    BwdImg I;
    {
        const float G = __ldg(P.gradLoss) * S.scaleWeight + __ldg(P.gradLoss + 1 + si);
        const float coef = __ldg(P.ctx + P.totalImages + 2 * si) * G;
        const float alpha = __ldg(P.ctx + S.imgBase + img);
        I.cp = -coef * P.p2n * alpha;
        I.cn = coef * (1.0f - alpha);
        I.maskBinary = MASK && (__ldg(P.ctx + P.totalImages + 2 * si + 1) != 0.f);
    }

    Row<VEC> xr[3];      // rolling depth rows: xr[j % 3] holds image row (row0 - 2 + j)
    CoefRow<VEC> cr[3];  // rolling coefficient rows: cr[j % 3] holds image row (row0 - 1 + j)
    load_row<VEC>(xr[0], x, row0 - 2, H, W, col0, lane, P.predInv != 0);
    load_row<VEC>(xr[1], x, row0 - 1, H, W, col0, lane, P.predInv != 0);

#pragma unroll
    for (int j = 0; j < RH + 2; j++) {
        const int row = row0 - 1 + j;  // row of the coefficient computed in this step
        load_row<VEC>(xr[(j + 2) % 3], x, row + 1, H, W, col0, lane, P.predInv != 0);
        const Row<VEC> &up = xr[j % 3], &mid = xr[(j + 1) % 3], &dn = xr[(j + 2) % 3];
        float e[VEC], th[VEC], m[VEC];
        load_plane_row<VEC>(e, S.e + plane, row, H, W, col0, 0.f);
        if (MODE == MODE_DIR) load_plane_row<VEC>(th, S.n + plane, row, H, W, col0, 0.f);
        if (MASK) load_plane_row<VEC>(m, S.m + plane, row, H, W, col0, 1.f);
        const bool rowOk = row >= 0 && row < H;
        CoefRow<VEC> &C = cr[j % 3];
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            float cv, ch, clr, crl;
            responses<VEC>(up, mid, dn, v, cv, ch, clr, crl);
            Coef k;
            if (MODE == MODE_MAG) {
                const float g = sqrtf(cv * cv + ch * ch + 1e-6f);
                const float d = dloss_dg<MODE, MASK>(g, e[v], MASK ? m[v] : 1.f, I, P.isSigmoid != 0, P.T);
                const float rg = d / g;
                const float sv = rg * cv, sh = rg * ch;
                k.U = sv + sh; k.V = sv - sh; k.A2 = 2.f * sv; k.C2 = 2.f * sh;
            } else {
                const int di = dir_index(th[v]);
                const float c = (di == 0) ? ch : (di == 1) ? crl : (di == 2) ? cv : clr;
                const float d = dloss_dg<MODE, MASK>(fabsf(c), e[v], MASK ? m[v] : 1.f, I, P.isSigmoid != 0, P.T);
                const float s = (c > 0.f) ? d : ((c < 0.f) ? -d : 0.f);
                // h: U=s V=-s A2=0 C2=2s | rl: U=0 V=-2s A2=-s C2=s | v: U=s V=s A2=2s C2=0 | lr: U=2s V=0 A2=s C2=s
                k.U = (di == 1) ? 0.f : ((di == 3) ? 2.f * s : s);
                k.V = (di == 0) ? -s : ((di == 1) ? -2.f * s : ((di == 2) ? s : 0.f));
                k.A2 = (di == 0) ? 0.f : ((di == 1) ? -s : ((di == 2) ? 2.f * s : s));
                k.C2 = (di == 0) ? 2.f * s : ((di == 2) ? 0.f : s);
            }
            if (!(rowOk && colOk)) { k.U = 0.f; k.V = 0.f; k.A2 = 0.f; k.C2 = 0.f; }
            C.c[v] = k;
        }
        // neighbours across the lane boundary
        C.l.U = __shfl_up_sync(MTE_FULL_MASK, C.c[VEC - 1].U, 1);
        C.l.V = __shfl_up_sync(MTE_FULL_MASK, C.c[VEC - 1].V, 1);
        C.l.C2 = __shfl_up_sync(MTE_FULL_MASK, C.c[VEC - 1].C2, 1);
        C.r.U = __shfl_down_sync(MTE_FULL_MASK, C.c[0].U, 1);
        C.r.V = __shfl_down_sync(MTE_FULL_MASK, C.c[0].V, 1);
        C.r.C2 = __shfl_down_sync(MTE_FULL_MASK, C.c[0].C2, 1);
        C.l.A2 = 0.f; C.r.A2 = 0.f;

        if (j >= 2) {
            const int orow = row - 1;  // == row0 + j - 2
            const CoefRow<VEC> &cu = cr[(j - 2) % 3], &cm = cr[(j - 1) % 3], &cd = cr[j % 3];
            float out[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                float d = (cu.at(v - 1).U + cu.at(v).A2 + cu.at(v + 1).V) + (cm.at(v - 1).C2 - cm.at(v + 1).C2) -
                          (cd.at(v - 1).V + cd.at(v).A2 + cd.at(v + 1).U);
                if (P.predInv) {
                    // d depth / d inv = -depth^2 where inv >= 1e-6 (clamp passes the gradient), else 0
                    // (inv <= 1e-6 maps to depth == 1e6: treated as clamped)
                    const float dep = xr[j % 3].c[v];  // image row orow == row0-2+j
                    d = (dep < 1e6f) ? -d * dep * dep : 0.f;
                }
                out[v] = d;
            }
            if (lane >= 1 && lane <= 30 && colOk && orow < H) {
                float *o = S.dx + plane + (size_t)orow * W + col0;
                if (VEC == 4) st_stream4(o, make_float4(out[0], out[1 % VEC], out[2 % VEC], out[3 % VEC]));
                else __stcs(o, out[0]);
            }
        }
    }
}

// is_grad == 0 (DEE-training mode, EdgeEstimationLIDARModel.py:139-144): the loss acts on
// the map itself, the backward is pointwise.
template <bool MASK>
__global__ void __launch_bounds__(kThreads) edge_loss_bwd_pointwise_kernel(const __grid_constant__ LossP P) {
    for (int si = 0; si < P.nScales; si++) {
        const ScaleP &S = P.s[si];
        const size_t npix = (size_t)S.H * S.W;
        const size_t n = npix * S.B;
        const float G = __ldg(P.gradLoss) * S.scaleWeight + __ldg(P.gradLoss + 1 + si);
        const float coef = __ldg(P.ctx + P.totalImages + 2 * si) * G;
        const bool binary = MASK && (__ldg(P.ctx + P.totalImages + 2 * si + 1) != 0.f);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const int img = (int)(i / npix);
            const float alpha = __ldg(P.ctx + S.imgBase + img);
            BwdImg I{-coef * P.p2n * alpha, coef * (1.0f - alpha), binary};
            float xv = S.x[i];
            const float dep = P.predInv ? inv_to_depth(xv) : xv;
            float d = dloss_dg<MODE_NONE, MASK>(dep, S.e[i], MASK ? S.m[i] : 1.f, I, P.isSigmoid != 0, P.T);
            if (P.predInv) d = (dep < 1e6f) ? -d * dep * dep : 0.f;
            S.dx[i] = d;
        }
    }
}

// ---------------------------------------------------------------------------
// Bilinear resize (F.interpolate(mode='bilinear', align_corners=False),
// grad_loss.py:127) -- only used when the prediction and target sizes differ.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void src_index(int o, float scale, int n, int &i0, int &i1, float &f) {
    float s = scale * ((float)o + 0.5f) - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = min((int)s, n - 1);
    i1 = min(i0 + 1, n - 1);
    f = s - (float)i0;
}

__global__ void resize_fwd_kernel(const float *__restrict__ in, float *__restrict__ out, int B, int h, int w, int H,
                                  int W) {
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    const size_t n = (size_t)B * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % W), Y = (int)((i / W) % H), b = (int)(i / ((size_t)W * H));
        int y0, y1, x0, x1;
        float fy, fx;
        src_index(Y, sy, h, y0, y1, fy);
        src_index(X, sx, w, x0, x1, fx);
        const float *p = in + (size_t)b * h * w;
        out[i] = (1.f - fy) * ((1.f - fx) * p[(size_t)y0 * w + x0] + fx * p[(size_t)y0 * w + x1]) +
                 fy * ((1.f - fx) * p[(size_t)y1 * w + x0] + fx * p[(size_t)y1 * w + x1]);
    }
}

// Deterministic adjoint: every source pixel gathers from the destination pixels that read it.
__global__ void resize_bwd_kernel(const float *__restrict__ gout, float *__restrict__ gin, int B, int h, int w, int H,
                                  int W) {
    const float sy = (float)h / (float)H, sx = (float)w / (float)W;
    const size_t n = (size_t)B * h * w;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)((i / w) % h), b = (int)(i / ((size_t)w * h));
        // destination rows/cols whose source interval can touch (y, x)
        int Y0 = (int)floorf(((float)y - 1.0f + 0.5f) / sy - 0.5f) - 1, Y1 = (int)ceilf(((float)y + 1.0f + 0.5f) / sy - 0.5f) + 1;
        int X0 = (int)floorf(((float)x - 1.0f + 0.5f) / sx - 0.5f) - 1, X1 = (int)ceilf(((float)x + 1.0f + 0.5f) / sx - 0.5f) + 1;
        if (y == 0) Y0 = 0;
        if (x == 0) X0 = 0;
        if (y == h - 1) Y1 = H - 1;
        if (x == w - 1) X1 = W - 1;
        Y0 = max(Y0, 0); X0 = max(X0, 0); Y1 = min(Y1, H - 1); X1 = min(X1, W - 1);
        const float *g = gout + (size_t)b * H * W;
        float acc = 0.f;
        for (int Y = Y0; Y <= Y1; Y++) {
            int a0, a1; float fy;
            src_index(Y, sy, h, a0, a1, fy);
            float wy = 0.f;
            if (a0 == y) wy += 1.f - fy;
            if (a1 == y) wy += fy;
            if (wy == 0.f) continue;
            for (int X = X0; X <= X1; X++) {
                int b0, b1; float fx;
                src_index(X, sx, w, b0, b1, fx);
                float wx = 0.f;
                if (b0 == x) wx += 1.f - fx;
                if (b1 == x) wx += fx;
                if (wx != 0.f) acc += wy * wx * g[(size_t)Y * W + X];
            }
        }
        gin[i] = acc;
    }
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
constexpr int kFwdRH = 4;
constexpr int kBwdRH = 8;

struct Plan {
    LossP P;
    bool vec, hasNormal, hasMask, isGrad;
    size_t offPartials, offSegsums, offResize[MTE_MAX_SCALES], offResizeDx[MTE_MAX_SCALES], total;
    bool resized[MTE_MAX_SCALES];
};

static int validate(const mte_loss_scale_t *sc, int n) {
    if (!sc) return MTE_ERR_NULL;
    if (n < 1 || n > MTE_MAX_SCALES) return MTE_ERR_ARG;
    for (int i = 0; i < n; i++) {
        if (!sc[i].pred || !sc[i].edge) return MTE_ERR_NULL;
        if (sc[i].B < 1 || sc[i].h < 1 || sc[i].w < 1 || sc[i].H < 1 || sc[i].W < 1) return MTE_ERR_SHAPE;
        if ((sc[i].normal != nullptr) != (sc[0].normal != nullptr)) return MTE_ERR_ARG;
        if ((sc[i].mask != nullptr) != (sc[0].mask != nullptr)) return MTE_ERR_ARG;
        const void *ptrs[] = {sc[i].pred, sc[i].edge, sc[i].normal, sc[i].mask, sc[i].grad_map, sc[i].grad_pred};
        for (const void *p : ptrs)
            if (reinterpret_cast<uintptr_t>(p) & 3u) return MTE_ERR_ALIGN;
    }
    return MTE_OK;
}

// Lay out CTAs / workspace.  bwd == true uses the overlapped 30-lane strips.
static int make_plan(Plan &pl, const mte_loss_scale_t *sc, int n, const mte_loss_attrs_t *at, bool bwd) {
    int rc = validate(sc, n);
    if (rc) return rc;
    LossP &P = pl.P;
    memset(&P, 0, sizeof(P));
    pl.hasNormal = sc[0].normal != nullptr;
    pl.hasMask = sc[0].mask != nullptr;
    pl.isGrad = at ? at->is_grad != 0 : true;
    pl.vec = true;
    for (int i = 0; i < n; i++) {
        if (sc[i].W % 4) pl.vec = false;
        // the kernels read the (possibly resized) prediction at target resolution
        const void *ptrs[] = {sc[i].edge, sc[i].normal, sc[i].mask, sc[i].grad_map};
        for (const void *p : ptrs)
            if (!aligned16(p)) pl.vec = false;
        const bool rs = sc[i].h != sc[i].H || sc[i].w != sc[i].W;
        if (!rs && (!aligned16(sc[i].pred) || (bwd && !aligned16(sc[i].grad_pred)))) pl.vec = false;
    }
    const int VEC = pl.vec ? 4 : 1;
    const int RH = bwd ? kBwdRH : kFwdRH;
    const int lanesOut = (bwd && pl.isGrad) ? 30 : 32;
    int cta = 0, img = 0;
    for (int i = 0; i < n; i++) {
        ScaleP &S = P.s[i];
        S.B = sc[i].B; S.H = sc[i].H; S.W = sc[i].W;
        S.e = sc[i].edge; S.n = sc[i].normal; S.m = sc[i].mask; S.g = sc[i].grad_map; S.dx = sc[i].grad_pred;
        S.x = sc[i].pred;
        S.strips = ceil_div(S.W, lanesOut * VEC);
        S.rowBlocks = ceil_div(S.H, RH);
        S.items = S.strips * S.rowBlocks;
        S.ctasPerImage = ceil_div(S.items, kWarps);
        S.ctaBase = cta; S.imgBase = img;
        S.scaleWeight = sc[i].scale_weight;
        cta += S.ctasPerImage * S.B;
        img += S.B;
        pl.resized[i] = sc[i].h != sc[i].H || sc[i].w != sc[i].W;
    }
    P.nScales = n; P.totalCtas = cta; P.totalImages = img;
    if (at) {
        P.isSigmoid = at->is_sigmoid; P.predInv = at->pred_is_inverse;
        P.T = at->sigmoid_thresh; P.weight = at->weight; P.p2n = at->pos_to_neg;
    }
    size_t off = MTE_WS_HEADER_BYTES;
    // forward partial layout is what sizes the workspace (bwd uses none of it)
    int fwdCtas = 0;
    for (int i = 0; i < n; i++) {
        const int items = ceil_div(sc[i].W, 32) * ceil_div(sc[i].H, kFwdRH);  // upper bound (VEC=1)
        fwdCtas += ceil_div(items, kWarps) * sc[i].B;
    }
    pl.offPartials = off; off += align_up((size_t)fwdCtas * kAcc * sizeof(double), 256);
    pl.offSegsums = off; off += align_up((size_t)img * kAcc * sizeof(double), 256);
    for (int i = 0; i < n; i++) {
        const size_t planeBytes = align_up((size_t)sc[i].B * sc[i].H * sc[i].W * sizeof(float), 256);
        pl.offResize[i] = off;
        if (pl.resized[i]) off += planeBytes;
        pl.offResizeDx[i] = off;
        if (pl.resized[i]) off += planeBytes;
    }
    pl.total = off;
    return MTE_OK;
}

template <int VEC, int MODE, bool MASK>
static void launch_fwd(const LossP &P, cudaStream_t st) {
    edge_loss_fwd_kernel<VEC, MODE, MASK, kFwdRH><<<P.totalCtas, kThreads, 0, st>>>(P);
}
template <int VEC, int MODE, bool MASK>
static void launch_bwd(const LossP &P, cudaStream_t st) {
    edge_loss_bwd_kernel<VEC, MODE, MASK, kBwdRH><<<P.totalCtas, kThreads, 0, st>>>(P);
}

template <int VEC, bool MASK>
static void dispatch_fwd(int mode, const LossP &P, cudaStream_t st) {
    if (mode == MODE_NONE) launch_fwd<VEC, MODE_NONE, MASK>(P, st);
    else if (mode == MODE_MAG) launch_fwd<VEC, MODE_MAG, MASK>(P, st);
    else launch_fwd<VEC, MODE_DIR, MASK>(P, st);
}
template <int VEC, bool MASK>
static void dispatch_bwd(int mode, const LossP &P, cudaStream_t st) {
    if (mode == MODE_MAG) launch_bwd<VEC, MODE_MAG, MASK>(P, st);
    else launch_bwd<VEC, MODE_DIR, MASK>(P, st);
}

}  // namespace loss
}  // namespace mte

using namespace mte;
using namespace mte::loss;

extern "C" size_t mte_edge_loss_workspace_bytes(const mte_loss_scale_t *sc, int n) {
    Plan pl;
    if (make_plan(pl, sc, n, nullptr, false)) return 0;
    return pl.total;
}

extern "C" size_t mte_edge_loss_ctx_bytes(const mte_loss_scale_t *sc, int n) {
    if (validate(sc, n)) return 0;
    size_t imgs = 0;
    for (int i = 0; i < n; i++) imgs += sc[i].B;
    return align_up((imgs + 2 * MTE_MAX_SCALES) * sizeof(float), 16);
}

extern "C" int mte_edge_loss_fwd(const mte_loss_scale_t *sc, int n, const mte_loss_attrs_t *at, float *loss_out,
                                 void *ctx, void *ws, size_t ws_bytes, mte_stream_t stream) {
    if (!at || !loss_out || !ctx || !ws) return MTE_ERR_NULL;
    Plan pl;
    int rc = make_plan(pl, sc, n, at, false);
    if (rc) return rc;
    if (ws_bytes < pl.total) return MTE_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char *w = static_cast<char *>(ws);
    LossP &P = pl.P;
    P.partials = reinterpret_cast<double *>(w + pl.offPartials);
    P.segsums = reinterpret_cast<double *>(w + pl.offSegsums);
    P.ticket = reinterpret_cast<WsHeader *>(w)->ticket;
    P.lossOut = loss_out;
    P.ctx = static_cast<float *>(ctx);
    for (int i = 0; i < n; i++) {
        if (pl.resized[i]) {
            float *tmp = reinterpret_cast<float *>(w + pl.offResize[i]);
            const size_t npx = (size_t)sc[i].B * sc[i].H * sc[i].W;
            const int grid = (int)((npx + 255) / 256 < 148 * 16 ? (npx + 255) / 256 : 148 * 16);
            resize_fwd_kernel<<<grid, 256, 0, st>>>(sc[i].pred, tmp, sc[i].B, sc[i].h, sc[i].w, sc[i].H, sc[i].W);
            P.s[i].x = tmp;
        }
    }
    const int mode = !at->is_grad ? MODE_NONE : (pl.hasNormal ? MODE_DIR : MODE_MAG);
    if (pl.vec) {
        if (pl.hasMask) dispatch_fwd<4, true>(mode, P, st); else dispatch_fwd<4, false>(mode, P, st);
    } else {
        if (pl.hasMask) dispatch_fwd<1, true>(mode, P, st); else dispatch_fwd<1, false>(mode, P, st);
    }
    MTE_RETURN_IF_CUDA_ERROR();
    return MTE_OK;
}

extern "C" int mte_edge_loss_bwd(const mte_loss_scale_t *sc, int n, const mte_loss_attrs_t *at,
                                 const float *grad_loss, const void *ctx, void *ws, size_t ws_bytes,
                                 mte_stream_t stream) {
    if (!at || !grad_loss || !ctx || !ws) return MTE_ERR_NULL;
    Plan pl;
    int rc = make_plan(pl, sc, n, at, true);
    if (rc) return rc;
    for (int i = 0; i < n; i++)
        if (!sc[i].grad_pred) return MTE_ERR_NULL;
    if (ws_bytes < pl.total) return MTE_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char *w = static_cast<char *>(ws);
    LossP &P = pl.P;
    P.ctx = const_cast<float *>(static_cast<const float *>(ctx));
    P.gradLoss = grad_loss;
    for (int i = 0; i < n; i++) {
        if (pl.resized[i]) {
            // recompute the resized prediction (the workspace is not carried from the forward)
            float *tmp = reinterpret_cast<float *>(w + pl.offResize[i]);
            const size_t npx = (size_t)sc[i].B * sc[i].H * sc[i].W;
            const int grid = (int)((npx + 255) / 256 < 148 * 16 ? (npx + 255) / 256 : 148 * 16);
            resize_fwd_kernel<<<grid, 256, 0, st>>>(sc[i].pred, tmp, sc[i].B, sc[i].h, sc[i].w, sc[i].H, sc[i].W);
            P.s[i].x = tmp;
        }
    }
    // resized scales: the gradient is produced at target resolution, then pulled back by the resize adjoint
    for (int i = 0; i < n; i++)
        if (pl.resized[i]) P.s[i].dx = reinterpret_cast<float *>(w + pl.offResizeDx[i]);
    if (!at->is_grad) {
        if (pl.hasMask) edge_loss_bwd_pointwise_kernel<true><<<kNumSMs * 8, kThreads, 0, st>>>(P);
        else edge_loss_bwd_pointwise_kernel<false><<<kNumSMs * 8, kThreads, 0, st>>>(P);
    } else {
        const int mode = pl.hasNormal ? MODE_DIR : MODE_MAG;
        if (pl.vec) {
            if (pl.hasMask) dispatch_bwd<4, true>(mode, P, st); else dispatch_bwd<4, false>(mode, P, st);
        } else {
            if (pl.hasMask) dispatch_bwd<1, true>(mode, P, st); else dispatch_bwd<1, false>(mode, P, st);
        }
    }
    MTE_RETURN_IF_CUDA_ERROR();
    for (int i = 0; i < n; i++) {
        if (pl.resized[i]) {
            const size_t npx = (size_t)sc[i].B * sc[i].h * sc[i].w;
            const int grid = (int)((npx + 255) / 256 < 148 * 16 ? (npx + 255) / 256 : 148 * 16);
            resize_bwd_kernel<<<grid, 256, 0, st>>>(P.s[i].dx, sc[i].grad_pred, sc[i].B, sc[i].h, sc[i].w, sc[i].H,
                                                   sc[i].W);
        }
    }
    MTE_RETURN_IF_CUDA_ERROR();
    return MTE_OK;
}
