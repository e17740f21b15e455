// Instantiates the tcgen05 3xTF32 Gram kernels (gram_tf32_kernel.cuh) for the PM_LAPLACE epilogue.
#include "gram_tf32_kernel.cuh"
namespace jrk {
namespace tf32 {
JRK_TF32_DEFINE_RUN_PM(PM_LAPLACE)
}  // namespace tf32
}  // namespace jrk
