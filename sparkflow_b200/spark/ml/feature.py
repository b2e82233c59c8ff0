"""The ``pyspark.ml.feature`` stages the reference's examples use, plus ``StopWordsRemover`` – the
JVM class the reference abuses as a *carrier* for pickled python stages (pipeline_util.py:16-31)."""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np

from ..context import keyword_only
from ..sql import DataFrame, Row
from .base import DefaultParamsPersistence, Transformer, java_class
from .linalg import DenseVector, SparseVector, Vector, Vectors
from .param import HasInputCol, HasInputCols, HasOutputCol, Param, Params, TypeConverters


def _append_column(df: DataFrame, name: str, fn) -> DataFrame:
    cols = df.columns + ([name] if name not in df.columns else [])
    idx = {c: i for i, c in enumerate(df.columns)}
    out = []
    for part in df._parts:
        rows = []
        for r in part:
            vals = list(r)
            v = fn(r)
            if name in idx:
                vals[idx[name]] = v
            else:
                vals.append(v)
            rows.append(Row._make(cols, vals))
        out.append(rows)
    return DataFrame(out, cols, df.ctx)


@java_class("org.apache.spark.ml.feature.VectorAssembler")
class VectorAssembler(Transformer, HasInputCols, HasOutputCol, DefaultParamsPersistence):
    handleInvalid = Param(Params._dummy(), "handleInvalid", "how to handle invalid data", typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, inputCols=None, outputCol=None, handleInvalid="error"):
        super().__init__()
        self._setDefault(handleInvalid="error")
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, inputCols=None, outputCol=None, handleInvalid="error"):
        return self._set(**self._input_kwargs)

    def _transform(self, dataset: DataFrame) -> DataFrame:
        cols, out = self.getInputCols(), self.getOutputCol()
        idx = [dataset.columns.index(c) for c in cols]

        def build(r):
            parts = []
            for i in idx:
                v = r[i]
                parts.append(np.asarray(v.toArray() if isinstance(v, Vector) else v, dtype=np.float64).reshape(-1))
            return DenseVector(np.concatenate(parts) if parts else np.zeros(0))

        return _append_column(dataset, out, build)


@java_class("org.apache.spark.ml.feature.OneHotEncoder")
class OneHotEncoder(Transformer, HasInputCol, HasOutputCol, DefaultParamsPersistence):
    dropLast = Param(Params._dummy(), "dropLast", "whether to drop the last category", typeConverter=TypeConverters.toBoolean)

    @keyword_only
    def __init__(self, dropLast=True, inputCol=None, outputCol=None):
        super().__init__()
        self._setDefault(dropLast=True)
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, dropLast=True, inputCol=None, outputCol=None):
        return self._set(**self._input_kwargs)

    def _transform(self, dataset: DataFrame) -> DataFrame:
        inp, out = self.getInputCol(), self.getOutputCol()
        i = dataset.columns.index(inp)
        n = int(max((int(r[i]) for p in dataset._parts for r in p), default=-1)) + 1
        size = n - 1 if self.getOrDefault(self.dropLast) else n

        def build(r):
            k = int(r[i])
            return SparseVector(size, [k], [1.0]) if k < size else SparseVector(size, [], [])

        return _append_column(dataset, out, build)


@java_class("org.apache.spark.ml.feature.Normalizer")
class Normalizer(Transformer, HasInputCol, HasOutputCol, DefaultParamsPersistence):
    p = Param(Params._dummy(), "p", "the p norm value.", typeConverter=TypeConverters.toFloat)

    @keyword_only
    def __init__(self, p=2.0, inputCol=None, outputCol=None):
        super().__init__()
        self._setDefault(p=2.0)
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, p=2.0, inputCol=None, outputCol=None):
        return self._set(**self._input_kwargs)

    def _transform(self, dataset: DataFrame) -> DataFrame:
        inp, out, p = self.getInputCol(), self.getOutputCol(), float(self.getOrDefault(self.p))
        i = dataset.columns.index(inp)

        def build(r):
            v = np.asarray(r[i].toArray() if isinstance(r[i], Vector) else r[i], dtype=np.float64)
            nrm = np.linalg.norm(v, p)
            return DenseVector(v / nrm if nrm > 0 else v)

        return _append_column(dataset, out, build)


@java_class("org.apache.spark.ml.feature.StopWordsRemover")
class StopWordsRemover(Transformer, HasInputCol, HasOutputCol, DefaultParamsPersistence):
    stopWords = Param(Params._dummy(), "stopWords", "The words to be filtered out", typeConverter=TypeConverters.toListString)
    caseSensitive = Param(Params._dummy(), "caseSensitive", "whether to do a case sensitive comparison over the stop words",
                          typeConverter=TypeConverters.toBoolean)

    @keyword_only
    def __init__(self, inputCol=None, outputCol=None, stopWords=None, caseSensitive=False):
        super().__init__()
        self._setDefault(stopWords=[], caseSensitive=False)
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, inputCol=None, outputCol=None, stopWords=None, caseSensitive=False):
        return self._set(**self._input_kwargs)

    def setStopWords(self, value: List[str]):
        return self._set(stopWords=list(value))

    def getStopWords(self) -> List[str]:
        return self.getOrDefault(self.stopWords)

    def _transform(self, dataset: DataFrame) -> DataFrame:
        inp, out = self.getInputCol(), self.getOutputCol()
        i = dataset.columns.index(inp)
        cs = self.getOrDefault(self.caseSensitive)
        stop = set(self.getStopWords() if cs else [w.lower() for w in self.getStopWords()])
        return _append_column(dataset, out, lambda r: [w for w in r[i] if (w if cs else w.lower()) not in stop])
