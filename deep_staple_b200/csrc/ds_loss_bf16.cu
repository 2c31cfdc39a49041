// Instantiations of the fused loss kernel for __nv_bfloat16 logits (see ds_loss.cuh).
#include "ds_loss_launch.cuh"

namespace ds {
template int launch_loss_typed<__nv_bfloat16>(const LossParams&, int, bool, bool, bool, cudaStream_t);
}  // namespace ds
