"""``pyspark.ml.linalg``: DenseVector / SparseVector / Vectors (numpy-backed)."""
from __future__ import annotations

from typing import Iterable, Sequence

import numpy as np


class Vector:
    def toArray(self) -> np.ndarray:  # pragma: no cover
        raise NotImplementedError

    def __array__(self, dtype=None, copy=None):
        arr = self.toArray()
        return arr.astype(dtype) if dtype is not None else arr

    def __len__(self):
        return self.size

    def __iter__(self):
        return iter(self.toArray())

    def __eq__(self, other):
        return isinstance(other, Vector) and self.size == other.size and np.array_equal(self.toArray(), other.toArray())

    def __hash__(self):
        return hash(self.toArray().tobytes())


class DenseVector(Vector):
    def __init__(self, values: Iterable[float]):
        self.array = np.asarray(list(values) if not isinstance(values, np.ndarray) else values, dtype=np.float64).reshape(-1)

    values = property(lambda self: self.array)
    size = property(lambda self: int(self.array.shape[0]))

    def toArray(self) -> np.ndarray:
        return self.array

    def __getitem__(self, i):
        return self.array[i]

    def dot(self, other) -> float:
        return float(np.dot(self.array, np.asarray(other)))

    def norm(self, p) -> float:
        return float(np.linalg.norm(self.array, p))

    def __repr__(self):
        return "DenseVector([%s])" % ", ".join("%g" % v for v in self.array)

    def __reduce__(self):
        return (DenseVector, (self.array,))


class SparseVector(Vector):
    def __init__(self, size: int, *args):
        self._size = int(size)
        if len(args) == 1:
            pairs = sorted(args[0].items()) if isinstance(args[0], dict) else sorted(args[0])
            self.indices = np.asarray([p[0] for p in pairs], dtype=np.int32)
            self.values = np.asarray([p[1] for p in pairs], dtype=np.float64)
        else:
            self.indices = np.asarray(args[0], dtype=np.int32)
            self.values = np.asarray(args[1], dtype=np.float64)
        if len(self.indices) != len(self.values):
            raise ValueError("index and value arrays not the same length")

    size = property(lambda self: self._size)

    def toArray(self) -> np.ndarray:
        out = np.zeros(self._size, dtype=np.float64)
        out[self.indices] = self.values
        return out

    def __getitem__(self, i):
        return self.toArray()[i]

    def __repr__(self):
        return "SparseVector(%d, {%s})" % (self._size, ", ".join("%d: %g" % kv for kv in zip(self.indices, self.values)))

    def __reduce__(self):
        return (SparseVector, (self._size, self.indices, self.values))


class Vectors:
    @staticmethod
    def dense(*elements) -> DenseVector:
        if len(elements) == 1 and not isinstance(elements[0], (float, int)):
            elements = elements[0]
        return DenseVector(elements)

    @staticmethod
    def sparse(size: int, *args) -> SparseVector:
        return SparseVector(size, *args)

    @staticmethod
    def zeros(size: int) -> DenseVector:
        return DenseVector(np.zeros(size))


VectorUDT = Vector
