// Host-side dispatch of loss_kernel over (label dtype, class count, vector width, outputs).
#pragma once

#include "ds_loss.cuh"
#include "ds_loss_tma.cuh"

namespace ds {

// Launch with programmatic stream serialisation allowed (programmatic dependent launch): if the kernel right before
// this one in the stream is one of ours that has executed griddepcontrol.launch_dependents (every CTA, after its
// streaming loop), this grid's CTAs may become resident while that kernel's tail is still running; they do their
// prologue and stream their static share, and stop at griddepcontrol.wait -- placed before their first access to the
// workspace -- until the previous grid has completed and flushed.  After any other kernel (cuDNN, the optimiser
// step, ...) the attribute changes nothing: a kernel that never triggers releases its dependents only on completion.
// DS_LOSS_PDL=0 launches the plain way.
template <typename K>
void launch_kernel(K kernel, dim3 grid, int smem, cudaStream_t stream, const LossParams& p) {
  static const bool pdl = env_int("DS_LOSS_PDL", 1) != 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = (size_t)smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, p);  // errors surface through check_launch (cudaPeekAtLastError)
}

template <typename LT, typename TT, int C, int VEC>
int launch_cfg(const LossParams& p, bool grad, bool stats, cudaStream_t stream) {
  const dim3 grid((unsigned)p.grid);
  const bool dice = stats && p.dice_target != nullptr;
  if constexpr (C >= 2 && C <= 4 && VEC > 1) {
    if (p.use_tma && stats && !dice) {
      constexpr int smem = tma_shape((int)sizeof(LT), C).stages * tma_stage_bytes<LT, TT, C, VEC>();
      auto launch = [&](auto kernel) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail((int)e, "loss_kernel_tma: %s", cudaGetErrorString(e));
        launch_kernel(kernel, grid, smem, stream, p);
        count_launch();
        return check_launch("loss_kernel_tma");
      };
      return grad ? launch(loss_kernel_tma<LT, TT, C, VEC, true>) : launch(loss_kernel_tma<LT, TT, C, VEC, false>);
    }
  }
  if (grad && stats) {
    if (dice) launch_kernel(loss_kernel<LT, TT, C, VEC, true, true, true>, grid, 0, stream, p);
    else launch_kernel(loss_kernel<LT, TT, C, VEC, true, true, false>, grid, 0, stream, p);
  } else if (stats) {
    if (dice) launch_kernel(loss_kernel<LT, TT, C, VEC, false, true, true>, grid, 0, stream, p);
    else launch_kernel(loss_kernel<LT, TT, C, VEC, false, true, false>, grid, 0, stream, p);
  } else {
    loss_kernel<LT, TT, C, VEC, true, false, false><<<grid, kThreads, 0, stream>>>(p);  // no tail, no workspace
  }
  count_launch();
  return check_launch("loss_kernel");
}

template <typename LT, typename TT>
int launch_tt(const LossParams& p, bool grad, bool stats, bool vec, cudaStream_t stream) {
  if (!vec) return launch_cfg<LT, TT, 0, 1>(p, grad, stats, stream);
  constexpr int W = sizeof(LT) == 4 ? 4 : 8;
  switch (p.C) {
    case 2: return launch_cfg<LT, TT, 2, W>(p, grad, stats, stream);
    case 3: return launch_cfg<LT, TT, 3, (sizeof(LT) == 2 ? DS_HALF_C3_VEC : W)>(p, grad, stats, stream);
    case 4: return launch_cfg<LT, TT, 4, W>(p, grad, stats, stream);
    default: return launch_cfg<LT, TT, 0, 4>(p, grad, stats, stream);
  }
}

template <typename LT>
int launch_loss_typed(const LossParams& p, int target_dtype, bool grad, bool stats, bool vec, cudaStream_t stream) {
  return target_dtype == DS_I64 ? launch_tt<LT, int64_t>(p, grad, stats, vec, stream)
                                : launch_tt<LT, uint8_t>(p, grad, stats, vec, stream);
}

}  // namespace ds
