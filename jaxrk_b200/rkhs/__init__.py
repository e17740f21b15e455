from .base import Vec, LinOp
from .vector import FiniteVec, CombVec, inner
from .operator import FiniteOp, CrossCovOp, CovOp, CovOp_from_Samples, Cov_regul, Cov_inv, Cov_solve, Cmo, Cdo
