from .base import Reduce, NoReduce, Prefactors, TileView, Sum, Mean, BalancedRed, Center
from .lincomb import LinearReduce, SparseReduce, BlockReduce, KernelMapReduce
