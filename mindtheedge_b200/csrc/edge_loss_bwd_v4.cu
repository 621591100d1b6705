// Instantiations of the edge-loss bwd kernels for VEC=4 (one TU per direction/VEC: parallel builds).
#include "edge_loss_kernels.cuh"
namespace mte { namespace loss {
template <int MODE> static void go_bwd_v4(const LossP &P, bool mask, bool inv, bool sig, cudaStream_t st) {
    MTE_LOSS_DISPATCH_BOOL(mask, MASK, MTE_LOSS_DISPATCH_BOOL(inv, INV, MTE_LOSS_DISPATCH_BOOL(sig, SIG,
        edge_loss_bwd_kernel<4, MODE, MASK, INV, SIG><<<P.totalCtas, kThreads, 0, st>>>(P);)))
}
void launch_bwd_v4(const LossP &P, int mode, bool mask, bool inv, bool sig, cudaStream_t st) {
    if (mode == MODE_MAG) go_bwd_v4<MODE_MAG>(P, mask, inv, sig, st);
    else go_bwd_v4<MODE_DIR>(P, mask, inv, sig, st);
}
}}  // namespace mte::loss
