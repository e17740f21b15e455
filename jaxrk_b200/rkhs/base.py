"""ABCs of RKHS objects (jaxrk/rkhs/base.py:7-28)."""
from abc import ABC, abstractmethod
from collections.abc import Sized


class Vec(Sized, ABC):
    @abstractmethod
    def reduce_gram(self, gram, axis=0):
        pass

    @abstractmethod
    def inner(self, Y=None, full=True):
        pass


class LinOp(Vec, ABC):
    @abstractmethod
    def __matmul__(self, right_inp):
        pass
