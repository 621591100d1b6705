// Alternative edge-loss types of GradLoss.forward (packnet_code/packnet_sfm/losses/grad_loss.py:143-156):
//   'attention_loss'      attention_loss2(p, e, mask, False)   losses/attention_loss.py:21-49  (one global alpha)
//   'spatially_adaptive'  attention_loss2(p, e, mask, True)    (alpha from a 15x15 box filter of the target, :29-34)
//   'dice'                1000 * ((sum p^2 + sum e^2 + 1e-4) / (2 sum p e + 1e-4)) / numel, ADDED to the base loss
// They are not on the shipped path (configs/default_config.py:130 = 'cross_entropy'), so they are written as plain
// pointwise / gather kernels around the fused kernels' side outputs instead of new streaming variants:
//   forward   the fused cross-entropy forward produces the grad map g (= |directional response|, or the map itself
//             when !is_grad) and the 1 B/px stash; alt_fwd_kernel turns g into p and accumulates the type's sums;
//             alt_finalize_kernel folds them into the loss and the context the backward needs
//   backward  alt_dldg_kernel writes dL/dg, alt_scatter_kernel gathers it through the 3x3 adjoint of the picked
//             direction (stash byte) -- or passes it through when !is_grad -- and ADDS to / overwrites grad_pred.
#include <string.h>

#include "common.cuh"

namespace mte {
namespace alt {

constexpr int kThreads = 256;
enum { ACC_NPOS = 0, ACC_NNEG, ACC_SA, ACC_SB, ACC_PP, ACC_EE, ACC_PE, ACC_SW, ACC_N };
enum { CTX_ALPHA = 0, CTX_DICE_N, CTX_DICE_D, CTX_N };

struct AltP {
    const float *g, *e, *m, *alphaMap;  // [B,H,W] planes (m, alphaMap optional)
    size_t n;                           // B*H*W
    int types, isSigmoid;
    float T, weight;
    double *acc;                        // [ACC_N]
    float *ctx;                         // [CTX_N]
    const float *ceLoss;                // weighted cross-entropy loss of the fused kernel, or nullptr
    float *lossOut;                     // [2]
    const float *gradLoss;              // [2]: d total / d lossOut[0], d / d lossOut[1] (both apply)
    float *dldg;                        // [B,H,W]
};

__device__ __forceinline__ float prob_of(float g, const AltP &P) { return P.isSigmoid ? 1.0f / (1.0f + expf(-(g - P.T))) : g; }

// weight (detached) and BCE of attention_loss2 for one pixel; alpha is the positive-class share of the weight
__device__ __forceinline__ void attention_terms(float p, float t, float &wa, float &wb, float &bce) {
    const float pc = fminf(fmaxf(p, 1e-14f), 1.0f - 1e-14f);              // attention_loss.py:38
    wa = t * powf(4.0f, sqrtf(1.0f - pc));                                // * alpha
    wb = (1.0f - t) * powf(4.0f, sqrtf(pc));                              // * (1 - alpha)
    bce = -(t * fmaxf(logf(p), -100.0f) + (1.0f - t) * fmaxf(logf(1.0f - p), -100.0f));  // F.binary_cross_entropy
}

// 15x15 zero-padded box mean of the target -> alpha map (attention_loss.py:29-34)
__global__ void __launch_bounds__(kThreads) box15_kernel(const float *__restrict__ e, float *__restrict__ alpha, int B, int H,
                                                         int W) {
    const size_t n = (size_t)B * H * W;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const float *p = e + (i - (size_t)y * W - x);
        float s = 0.f;
        for (int dy = -7; dy <= 7; dy++) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            for (int dx = -7; dx <= 7; dx++) {
                const int xx = x + dx;
                if (xx >= 0 && xx < W) s += p[(size_t)yy * W + xx];
            }
        }
        float neg = 1.0f - s / 225.0f;
        if (neg >= 1.0f - 1e-14f) neg = 0.5f;
        alpha[i] = neg;
    }
}

__global__ void __launch_bounds__(kThreads) alt_fwd_kernel(const AltP P) {
    double a[ACC_N];
#pragma unroll
    for (int k = 0; k < ACC_N; k++) a[k] = 0.0;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < P.n; i += (size_t)gridDim.x * kThreads) {
        const float p = prob_of(P.g[i], P), t = P.e[i];
        const float mm = P.m ? P.m[i] : 1.0f;
        if (P.types & (MTE_LOSS_ATTENTION | MTE_LOSS_SPATIAL)) {
            float wa, wb, bce;
            attention_terms(p, t, wa, wb, bce);
            a[ACC_NPOS] += (t == 1.0f) ? 1.0 : 0.0;
            a[ACC_NNEG] += (t == 0.0f) ? 1.0 : 0.0;
            a[ACC_SA] += (double)(wa * mm * bce);
            a[ACC_SB] += (double)(wb * mm * bce);
            if (P.alphaMap) {
                const float al = P.alphaMap[i];
                a[ACC_SW] += (double)((wa * al + wb * (1.0f - al)) * mm * bce);
            }
        }
        if (P.types & MTE_LOSS_DICE) {
            a[ACC_PP] += (double)(p * p);
            a[ACC_EE] += (double)(t * t);
            a[ACC_PE] += (double)(p * t);
        }
    }
    __shared__ double sh[kThreads / 32][ACC_N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < ACC_N; k++) {
        const double v = warp_sum(a[k]);
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < ACC_N) {
        double v = 0.0;
        for (int w = 0; w < kThreads / 32; w++) v += sh[w][threadIdx.x];
        atomicAdd(P.acc + threadIdx.x, v);
    }
}

__global__ void alt_finalize_kernel(const AltP P) {
    const double numel = (double)P.n;
    double loss = 0.0;
    float alpha = 0.f;
    if (P.types & MTE_LOSS_CE) loss = P.weight != 0.f ? (double)*P.ceLoss / (double)P.weight : 0.0;  // the fused kernel already applied the weight
    if (P.types & MTE_LOSS_ATTENTION) {
        alpha = (float)P.acc[ACC_NNEG] / ((float)P.acc[ACC_NPOS] + (float)P.acc[ACC_NNEG]);   // attention_loss.py:25-27
        loss = ((double)alpha * P.acc[ACC_SA] + (1.0 - (double)alpha) * P.acc[ACC_SB]) / numel;
    }
    if (P.types & MTE_LOSS_SPATIAL) loss = P.acc[ACC_SW] / numel;                              // later type wins, as in :146-147
    const double dN = P.acc[ACC_PP] + P.acc[ACC_EE] + 0.0001, dD = 2.0 * P.acc[ACC_PE] + 0.0001;
    if (P.types & MTE_LOSS_DICE) loss += 1000.0 * (dN / dD) / numel;                           // grad_loss.py:154-156
    P.ctx[CTX_ALPHA] = alpha;
    P.ctx[CTX_DICE_N] = (float)dN;
    P.ctx[CTX_DICE_D] = (float)dD;
    P.lossOut[0] = (float)((double)P.weight * loss);
    P.lossOut[1] = P.lossOut[0];
}

// dL/dg of the non-cross-entropy terms (the weight of attention_loss2 is detached, attention_loss.py:42)
__global__ void __launch_bounds__(kThreads) alt_dldg_kernel(const AltP P) {
    const float G = (P.gradLoss[0] + P.gradLoss[1]) * P.weight / (float)P.n;
    const float alpha = P.ctx[CTX_ALPHA], dN = P.ctx[CTX_DICE_N], dD = P.ctx[CTX_DICE_D];
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < P.n; i += (size_t)gridDim.x * kThreads) {
        const float p = prob_of(P.g[i], P), t = P.e[i];
        float dp = 0.f;
        if (P.types & (MTE_LOSS_ATTENTION | MTE_LOSS_SPATIAL)) {
            float wa, wb, bce;
            attention_terms(p, t, wa, wb, bce);
            const float al = (P.types & MTE_LOSS_SPATIAL) ? P.alphaMap[i] : alpha;
            const float w = (wa * al + wb * (1.0f - al)) * (P.m ? P.m[i] : 1.0f);
            dp += w * (p - t) / fmaxf((1.0f - p) * p, 1e-12f);    // binary_cross_entropy backward
        }
        if (P.types & MTE_LOSS_DICE) dp += 1000.0f * (2.0f * p / dD - dN * 2.0f * t / (dD * dD));
        P.dldg[i] = G * dp * (P.isSigmoid ? p * (1.0f - p) : 1.0f);
    }
}

// 3x3 cross-correlation kernels of grad_loss.py:20-31, indexed by the stash direction 0:h 1:rl 2:v 3:lr
__constant__ float kK[4][3][3] = {
    {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}},
    {{0, 1, 2}, {-1, 0, 1}, {-2, -1, 0}},
    {{-1, -2, -1}, {0, 0, 0}, {1, 2, 1}},
    {{-2, -1, 0}, {-1, 0, 1}, {0, 1, 2}},
};

// d loss / d pred(m) = sum over the 3x3 neighbours n of K_dir(n)[m - n] * dldg(n) * sign(c(n))   
__global__ void __launch_bounds__(kThreads) alt_scatter_kernel(const float *__restrict__ dldg,
                                                               const float *__restrict__ gmap,
                                                               const unsigned char *__restrict__ stash,
                                                               const float *__restrict__ pred, float *__restrict__ dx, int B,
                                                               int H, int W, int isGrad, int predInv, int accumulate) {
    const size_t n = (size_t)B * H * W;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        float d;
        if (!isGrad) {
            d = dldg[i];
        } else {
            const int x = (int)(i % W), y = (int)((i / W) % H);
            const size_t plane = i - (size_t)y * W - x;
            d = 0.f;
#pragma unroll
            for (int a = -1; a <= 1; a++)
#pragma unroll
                for (int b = -1; b <= 1; b++) {
                    const int yy = y - a, xx = x - b;  // response pixel n with n + (a, b) = m
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    const size_t j = plane + (size_t)yy * W + xx;
                    const unsigned code = stash[j];
                    const float sg = (gmap[j] == 0.f) ? 0.f : ((code & 8u) ? -1.f : 1.f);  // sign(0) = 0
                    d += kK[code & 3u][a + 1][b + 1] * sg * dldg[j];
                }
        }
        if (predInv) {  // chain through depth = 1 / clamp(inv, 1e-6)
            const float inv = pred[i];
            const float dep = 1.0f / fmaxf(inv, 1e-6f);
            d = (inv >= 1e-6f) ? -d * dep * dep : 0.f;
        }
        dx[i] = accumulate ? dx[i] + d : d;
    }
}

// The same adjoint without normals (GradLayer's magnitude path, grad_loss.py:70-73: g = sqrt(c_v^2 + c_h^2 + 1e-6)):
// d g / d x = (c_v K_v + c_h K_h) / g, so every response pixel n contributes K_v[m-n] * dldg * c_v / g +
// K_h[m-n] * dldg * c_h / g; c_v, c_h are recomputed from the prediction (zero padding).  Off the shipped path: a
// plain gather (9 neighbours x two 3x3 stencils per pixel).
__global__ void __launch_bounds__(kThreads) alt_scatter_mag_kernel(const float *__restrict__ dldg,
                                                                   const float *__restrict__ gmap,
                                                                   const float *__restrict__ pred, float *__restrict__ dx,
                                                                   int B, int H, int W, int accumulate) {
    const size_t n = (size_t)B * H * W;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const float *img = pred + (i - (size_t)y * W - x);
        const size_t plane = i - (size_t)y * W - x;
        // fp64 accumulation: the stencil sums are then exact, so the only rounding left is the reference's own
        double d = 0.0;
        for (int a = -1; a <= 1; a++)
            for (int b = -1; b <= 1; b++) {
                const int yy = y - a, xx = x - b;  // response pixel n with n + (a, b) = m
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                double cv = 0.0, ch = 0.0;
                for (int u = -1; u <= 1; u++)
                    for (int v = -1; v <= 1; v++) {
                        const int py = yy + u, px = xx + v;
                        if (py < 0 || py >= H || px < 0 || px >= W) continue;
                        const double val = (double)img[(size_t)py * W + px];
                        cv += (double)kK[2][u + 1][v + 1] * val;
                        ch += (double)kK[0][u + 1][v + 1] * val;
                    }
                const size_t j = plane + (size_t)yy * W + xx;
                const double r = (double)dldg[j] / (double)gmap[j];   // g >= 1e-3 by construction
                d += ((double)kK[2][a + 1][b + 1] * (double)(float)cv + (double)kK[0][a + 1][b + 1] * (double)(float)ch) * r;
            }
        dx[i] = accumulate ? dx[i] + (float)d : (float)d;
    }
}

static int grid_for(size_t n) {
    const size_t b = (n + kThreads - 1) / kThreads;
    return (int)(b < (size_t)num_sms() * 8 ? (b ? b : 1) : (size_t)num_sms() * 8);
}

}  // namespace alt
}  // namespace mte

using namespace mte;
using namespace mte::alt;

static bool types_ok(int types) {
    if (types & ~(MTE_LOSS_CE | MTE_LOSS_ATTENTION | MTE_LOSS_SPATIAL | MTE_LOSS_DICE)) return false;
    return (types & (MTE_LOSS_CE | MTE_LOSS_ATTENTION | MTE_LOSS_SPATIAL)) != 0;  // 'dice' alone has no base loss (NameError in the reference)
}

extern "C" size_t mte_edge_loss_alt_workspace_bytes(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return 0;
    const size_t plane = align_up((size_t)B * H * W * sizeof(float), 256);
    return MTE_WS_HEADER_BYTES + 256 + 2 * plane;  // accumulators, alpha map, dL/dg
}

extern "C" int mte_edge_loss_alt_fwd(const float *grad_map, const float *edge, const float *mask, int B, int H, int W,
                                     int loss_types, int is_sigmoid, float sigmoid_thresh, float weight,
                                     const float *ce_loss, float *loss_out, float *ctx, void *workspace, size_t ws_bytes,
                                     mte_stream_t stream) {
    if (!grad_map || !edge || !loss_out || !ctx || !workspace) return MTE_ERR_NULL;
    if (B < 1 || H < 1 || W < 1) return MTE_ERR_SHAPE;
    if (!types_ok(loss_types) || ((loss_types & MTE_LOSS_CE) && !ce_loss)) return MTE_ERR_ARG;
    if (ws_bytes < mte_edge_loss_alt_workspace_bytes(B, H, W)) return MTE_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char *w = static_cast<char *>(workspace);
    AltP P;
    memset(&P, 0, sizeof(P));
    P.g = grad_map; P.e = edge; P.m = mask;
    P.n = (size_t)B * H * W;
    P.types = loss_types; P.isSigmoid = is_sigmoid; P.T = sigmoid_thresh; P.weight = weight;
    P.acc = reinterpret_cast<double *>(w + MTE_WS_HEADER_BYTES);
    P.ctx = ctx; P.ceLoss = ce_loss; P.lossOut = loss_out;
    cudaError_t e = cudaMemsetAsync(P.acc, 0, ACC_N * sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
    if (loss_types & MTE_LOSS_SPATIAL) {
        float *am = reinterpret_cast<float *>(w + MTE_WS_HEADER_BYTES + 256);
        box15_kernel<<<grid_for(P.n), kThreads, 0, st>>>(edge, am, B, H, W);
        P.alphaMap = am;
    }
    alt_fwd_kernel<<<grid_for(P.n), kThreads, 0, st>>>(P);
    alt_finalize_kernel<<<1, 1, 0, st>>>(P);
    MTE_RETURN_IF_CUDA_ERROR();
    return MTE_OK;
}

extern "C" int mte_edge_loss_alt_bwd(const float *grad_map, const float *edge, const float *mask, const uint8_t *stash,
                                     const float *pred, int B, int H, int W, int loss_types, int is_grad, int is_sigmoid,
                                     int pred_is_inverse, float sigmoid_thresh, float weight, const float *grad_loss,
                                     const float *ctx, float *grad_pred, int accumulate, void *workspace, size_t ws_bytes,
                                     mte_stream_t stream) {
    if (!grad_map || !edge || !grad_loss || !ctx || !grad_pred || !workspace) return MTE_ERR_NULL;
    // is_grad without a stash = the magnitude path (no normals): the adjoint is recomputed from the prediction
    if ((is_grad && !stash && !pred) || (pred_is_inverse && !pred)) return MTE_ERR_NULL;
    if (is_grad && !stash && pred_is_inverse) return MTE_ERR_ARG;
    if (B < 1 || H < 1 || W < 1) return MTE_ERR_SHAPE;
    if (!types_ok(loss_types)) return MTE_ERR_ARG;
    if (ws_bytes < mte_edge_loss_alt_workspace_bytes(B, H, W)) return MTE_ERR_WORKSPACE;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char *w = static_cast<char *>(workspace);
    const size_t plane = align_up((size_t)B * H * W * sizeof(float), 256);
    AltP P;
    memset(&P, 0, sizeof(P));
    P.g = grad_map; P.e = edge; P.m = mask;
    P.n = (size_t)B * H * W;
    P.types = loss_types; P.isSigmoid = is_sigmoid; P.T = sigmoid_thresh; P.weight = weight;
    P.ctx = const_cast<float *>(ctx); P.gradLoss = grad_loss;
    P.dldg = reinterpret_cast<float *>(w + MTE_WS_HEADER_BYTES + 256 + plane);
    if (loss_types & MTE_LOSS_SPATIAL) {
        float *am = reinterpret_cast<float *>(w + MTE_WS_HEADER_BYTES + 256);
        box15_kernel<<<grid_for(P.n), kThreads, 0, st>>>(edge, am, B, H, W);
        P.alphaMap = am;
    }
    alt_dldg_kernel<<<grid_for(P.n), kThreads, 0, st>>>(P);
    if (is_grad && !stash)
        alt_scatter_mag_kernel<<<grid_for(P.n), kThreads, 0, st>>>(P.dldg, grad_map, pred, grad_pred, B, H, W, accumulate);
    else
        alt_scatter_kernel<<<grid_for(P.n), kThreads, 0, st>>>(P.dldg, grad_map, stash, pred, grad_pred, B, H, W, is_grad,
                                                              pred_is_inverse, accumulate);
    MTE_RETURN_IF_CUDA_ERROR();
    return MTE_OK;
}
