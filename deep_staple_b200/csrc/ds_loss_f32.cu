// Instantiations of the fused loss kernel for float logits (see ds_loss.cuh).
#include "ds_loss_launch.cuh"

namespace ds {
template int launch_loss_typed<float>(const LossParams&, int, bool, bool, bool, cudaStream_t);
}  // namespace ds

#ifdef DS_TIMING
// tuning builds only: copy the phase stamps of the last float-logits launch to the host
extern "C" int ds_debug_read(void* dst, size_t bytes) {
  return (int)cudaMemcpyFromSymbol(dst, ds::g_dbg, bytes < sizeof(ds::g_dbg) ? bytes : sizeof(ds::g_dbg));
}
#endif
