// Instantiations of the fused loss kernel for float logits (see ds_loss.cuh).
#include "ds_loss_launch.cuh"

namespace ds {
template int launch_loss_typed<float>(const LossParams&, int, bool, bool, bool, cudaStream_t);
}  // namespace ds
