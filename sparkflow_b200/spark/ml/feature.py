"""The ``pyspark.ml.feature`` stages the reference's examples use, plus ``StopWordsRemover`` – the
JVM class the reference abuses as a *carrier* for pickled python stages (pipeline_util.py:16-31)."""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np

from ..context import keyword_only
from ..sql import DataFrame, Row
from .base import DefaultParamsPersistence, Estimator, Model, Transformer, java_class
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


# -------------------------------------------------------------------------------------------------
# Common pre-processing stages seen next to SparkAsyncDL in user pipelines.  Fitted state (labels, statistics) is kept in
# Params so the metadata-only persistence of this shim round-trips it (real Spark writes a parquet `data/` directory).
# -------------------------------------------------------------------------------------------------
def _as_array(v) -> np.ndarray:
    return np.asarray(v.toArray() if isinstance(v, Vector) else v, dtype=np.float64).reshape(-1)


@java_class("org.apache.spark.ml.feature.Binarizer")
class Binarizer(Transformer, HasInputCol, HasOutputCol, DefaultParamsPersistence):
    threshold = Param(Params._dummy(), "threshold", "features greater than the threshold become 1.0, the others 0.0",
                      typeConverter=TypeConverters.toFloat)

    @keyword_only
    def __init__(self, threshold=0.0, inputCol=None, outputCol=None):
        super().__init__()
        self._setDefault(threshold=0.0)
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, threshold=0.0, inputCol=None, outputCol=None):
        return self._set(**self._input_kwargs)

    def _transform(self, dataset: DataFrame) -> DataFrame:
        i, thr = dataset.columns.index(self.getInputCol()), float(self.getOrDefault(self.threshold))

        def build(r):
            v = r[i]
            if isinstance(v, Vector) or isinstance(v, (list, tuple, np.ndarray)):
                return DenseVector((_as_array(v) > thr).astype(np.float64))
            return 1.0 if float(v) > thr else 0.0

        return _append_column(dataset, self.getOutputCol(), build)


@java_class("org.apache.spark.ml.feature.StringIndexerModel")
class StringIndexerModel(Model, HasInputCol, HasOutputCol, DefaultParamsPersistence):
    labels = Param(Params._dummy(), "labels", "ordered list of labels, index = position", typeConverter=TypeConverters.toListString)
    handleInvalid = Param(Params._dummy(), "handleInvalid", "error | skip | keep", typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, inputCol=None, outputCol=None, labels=None, handleInvalid="error"):
        super().__init__()
        self._setDefault(labels=[], handleInvalid="error")
        self._set(**self._input_kwargs)

    def _transform(self, dataset: DataFrame) -> DataFrame:
        i = dataset.columns.index(self.getInputCol())
        labels = list(self.getOrDefault(self.labels))
        index = {l: float(k) for k, l in enumerate(labels)}
        invalid = self.getOrDefault(self.handleInvalid)
        if invalid == "skip":
            dataset = DataFrame([[r for r in p if str(r[i]) in index] for p in dataset._parts], dataset.columns, dataset.ctx)

        def build(r):
            key = str(r[i])
            if key in index:
                return index[key]
            if invalid == "keep":
                return float(len(labels))
            raise ValueError(f"Unseen label: {key}")

        return _append_column(dataset, self.getOutputCol(), build)


@java_class("org.apache.spark.ml.feature.StringIndexer")
class StringIndexer(Estimator, HasInputCol, HasOutputCol, DefaultParamsPersistence):
    """Most frequent label -> 0.0 (ties broken alphabetically, Spark's ``frequencyDesc``)."""
    handleInvalid = Param(Params._dummy(), "handleInvalid", "error | skip | keep", typeConverter=TypeConverters.toString)
    stringOrderType = Param(Params._dummy(), "stringOrderType", "frequencyDesc | frequencyAsc | alphabetDesc | alphabetAsc",
                            typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, inputCol=None, outputCol=None, handleInvalid="error", stringOrderType="frequencyDesc"):
        super().__init__()
        self._setDefault(handleInvalid="error", stringOrderType="frequencyDesc")
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, inputCol=None, outputCol=None, handleInvalid="error", stringOrderType="frequencyDesc"):
        return self._set(**self._input_kwargs)

    def _fit(self, dataset: DataFrame) -> StringIndexerModel:
        i = dataset.columns.index(self.getInputCol())
        counts: dict = {}
        for p in dataset._parts:
            for r in p:
                counts[str(r[i])] = counts.get(str(r[i]), 0) + 1
        order = self.getOrDefault(self.stringOrderType)
        if order == "frequencyDesc":
            labels = sorted(counts, key=lambda l: (-counts[l], l))
        elif order == "frequencyAsc":
            labels = sorted(counts, key=lambda l: (counts[l], l))
        elif order == "alphabetDesc":
            labels = sorted(counts, reverse=True)
        else:
            labels = sorted(counts)
        m = StringIndexerModel(inputCol=self.getInputCol(), outputCol=self.getOutputCol(), labels=labels,
                               handleInvalid=self.getOrDefault(self.handleInvalid))
        return m


class _ScalerModelBase(Model, HasInputCol, HasOutputCol, DefaultParamsPersistence):
    def _stats(self, name):
        return np.asarray(self.getOrDefault(getattr(self, name)), dtype=np.float64)


@java_class("org.apache.spark.ml.feature.StandardScalerModel")
class StandardScalerModel(_ScalerModelBase):
    mean = Param(Params._dummy(), "mean", "per-feature mean", typeConverter=TypeConverters.toListFloat)
    std = Param(Params._dummy(), "std", "per-feature (unbiased) standard deviation", typeConverter=TypeConverters.toListFloat)
    withMean = Param(Params._dummy(), "withMean", "center with the mean", typeConverter=TypeConverters.toBoolean)
    withStd = Param(Params._dummy(), "withStd", "scale to unit standard deviation", typeConverter=TypeConverters.toBoolean)

    @keyword_only
    def __init__(self, inputCol=None, outputCol=None, mean=None, std=None, withMean=False, withStd=True):
        super().__init__()
        self._setDefault(mean=[], std=[], withMean=False, withStd=True)
        self._set(**self._input_kwargs)

    def _transform(self, dataset: DataFrame) -> DataFrame:
        i = dataset.columns.index(self.getInputCol())
        mean, std = self._stats("mean"), self._stats("std")
        wm, ws = self.getOrDefault(self.withMean), self.getOrDefault(self.withStd)
        inv = np.where(std > 0, 1.0 / np.where(std > 0, std, 1.0), 0.0)

        def build(r):
            v = _as_array(r[i])
            if wm:
                v = v - mean
            if ws:
                v = v * inv
            return DenseVector(v)

        return _append_column(dataset, self.getOutputCol(), build)


@java_class("org.apache.spark.ml.feature.StandardScaler")
class StandardScaler(Estimator, HasInputCol, HasOutputCol, DefaultParamsPersistence):
    withMean = Param(Params._dummy(), "withMean", "center with the mean", typeConverter=TypeConverters.toBoolean)
    withStd = Param(Params._dummy(), "withStd", "scale to unit standard deviation", typeConverter=TypeConverters.toBoolean)

    @keyword_only
    def __init__(self, withMean=False, withStd=True, inputCol=None, outputCol=None):
        super().__init__()
        self._setDefault(withMean=False, withStd=True)
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, withMean=False, withStd=True, inputCol=None, outputCol=None):
        return self._set(**self._input_kwargs)

    def _fit(self, dataset: DataFrame) -> StandardScalerModel:
        i = dataset.columns.index(self.getInputCol())
        X = np.stack([_as_array(r[i]) for p in dataset._parts for r in p])
        std = X.std(axis=0, ddof=1) if X.shape[0] > 1 else np.zeros(X.shape[1])
        return StandardScalerModel(inputCol=self.getInputCol(), outputCol=self.getOutputCol(), mean=[float(v) for v in X.mean(axis=0)],
                                   std=[float(v) for v in std], withMean=self.getOrDefault(self.withMean), withStd=self.getOrDefault(self.withStd))


@java_class("org.apache.spark.ml.feature.MinMaxScalerModel")
class MinMaxScalerModel(_ScalerModelBase):
    originalMin = Param(Params._dummy(), "originalMin", "per-feature minimum seen in fit", typeConverter=TypeConverters.toListFloat)
    originalMax = Param(Params._dummy(), "originalMax", "per-feature maximum seen in fit", typeConverter=TypeConverters.toListFloat)
    min = Param(Params._dummy(), "min", "lower bound after transformation", typeConverter=TypeConverters.toFloat)  # noqa: A003
    max = Param(Params._dummy(), "max", "upper bound after transformation", typeConverter=TypeConverters.toFloat)  # noqa: A003

    @keyword_only
    def __init__(self, inputCol=None, outputCol=None, originalMin=None, originalMax=None, min=0.0, max=1.0):  # noqa: A002
        super().__init__()
        self._setDefault(originalMin=[], originalMax=[], min=0.0, max=1.0)
        self._set(**self._input_kwargs)

    def _transform(self, dataset: DataFrame) -> DataFrame:
        i = dataset.columns.index(self.getInputCol())
        lo, hi = self._stats("originalMin"), self._stats("originalMax")
        a, b = float(self.getOrDefault(self.min)), float(self.getOrDefault(self.max))
        span = hi - lo

        def build(r):
            v = _as_array(r[i])
            unit = np.where(span > 0, (v - lo) / np.where(span > 0, span, 1.0), 0.5)      # constant features map to the middle (Spark)
            return DenseVector(unit * (b - a) + a)

        return _append_column(dataset, self.getOutputCol(), build)


@java_class("org.apache.spark.ml.feature.MinMaxScaler")
class MinMaxScaler(Estimator, HasInputCol, HasOutputCol, DefaultParamsPersistence):
    min = Param(Params._dummy(), "min", "lower bound after transformation", typeConverter=TypeConverters.toFloat)  # noqa: A003
    max = Param(Params._dummy(), "max", "upper bound after transformation", typeConverter=TypeConverters.toFloat)  # noqa: A003

    @keyword_only
    def __init__(self, min=0.0, max=1.0, inputCol=None, outputCol=None):  # noqa: A002
        super().__init__()
        self._setDefault(min=0.0, max=1.0)
        self.setParams(**self._input_kwargs)

    @keyword_only
    def setParams(self, min=0.0, max=1.0, inputCol=None, outputCol=None):  # noqa: A002
        return self._set(**self._input_kwargs)

    def _fit(self, dataset: DataFrame) -> MinMaxScalerModel:
        i = dataset.columns.index(self.getInputCol())
        X = np.stack([_as_array(r[i]) for p in dataset._parts for r in p])
        return MinMaxScalerModel(inputCol=self.getInputCol(), outputCol=self.getOutputCol(), originalMin=[float(v) for v in X.min(axis=0)],
                                 originalMax=[float(v) for v in X.max(axis=0)], min=float(self.getOrDefault(self.min)),
                                 max=float(self.getOrDefault(self.max)))
