// Instantiations of the fused K@W FP32-pipe kernel (kmatw_kernel.cuh) for epilogue mode PM_SQEXP.
#include "kmatw_kernel.cuh"

namespace jrk {
JRK_KMATW_DEFINE_DISPATCH(PM_SQEXP)
}  // namespace jrk
