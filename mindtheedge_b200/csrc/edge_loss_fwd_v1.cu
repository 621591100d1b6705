// Instantiations of the edge-loss fwd kernels for VEC=1 (one TU per direction/VEC: parallel builds).
#include "edge_loss_kernels.cuh"
namespace mte { namespace loss {
template <int MODE> static void go_fwd_v1(const LossP &P, bool mask, bool inv, bool sig, cudaStream_t st) {
    MTE_LOSS_DISPATCH_BOOL(mask, MASK, MTE_LOSS_DISPATCH_BOOL(inv, INV, MTE_LOSS_DISPATCH_BOOL(sig, SIG,
        launch_fwd_one<1, MODE, MASK, INV, SIG>(P, st);)))
}
void launch_fwd_v1(const LossP &P, int mode, bool mask, bool inv, bool sig, cudaStream_t st) {
    if (mode == MODE_NONE) go_fwd_v1<MODE_NONE>(P, mask, inv, sig, st); else
    if (mode == MODE_MAG) go_fwd_v1<MODE_MAG>(P, mask, inv, sig, st);
    else go_fwd_v1<MODE_DIR>(P, mask, inv, sig, st);
}
}}  // namespace mte::loss
