"""jaxrk_b200 -- B200-native Gram / fused K@W path behind JaxRK's kern / reduce / rkhs API.

Sub-packages mirror the reference's names: `kern` (GenGaussKernel, ScaledPairwiseDistance,
SimpleScaler), `reduce` (Reduce family), `rkhs` (FiniteVec, inner, FiniteOp, Cov*/Cmo/Cdo),
`utilities` (eucldist, dist).  All kernel evaluations run in libjaxrk_b200.so
(hand-written sm_100a CUDA behind the C ABI of include/jrk.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import kern, reduce, rkhs, utilities  # noqa: E402,F401
