from .base import Kernel, DensityKernel
from .util import Scaler, NoScaler, SimpleScaler, ScaledPairwiseDistance
from .rbf import GenGaussKernel, PeriodicKernel, ThreshSpikeKernel
