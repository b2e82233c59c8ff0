from __future__ import annotations

import numpy as np

from ..context import keyword_only
from .base import Evaluator
from .linalg import Vector
from .param import HasLabelCol, HasPredictionCol, Param, Params, TypeConverters


class MulticlassClassificationEvaluator(Evaluator, HasLabelCol, HasPredictionCol):
    metricName = Param(Params._dummy(), "metricName", "metric name in evaluation (f1|weightedPrecision|weightedRecall|accuracy)",
                       typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, predictionCol="prediction", labelCol="label", metricName="f1"):
        super().__init__()
        self._setDefault(predictionCol="prediction", labelCol="label", metricName="f1")
        self._set(**self._input_kwargs)

    def _evaluate(self, dataset) -> float:
        lc, pc = self.getLabelCol(), self.getPredictionCol()
        rows = dataset.collect()
        y = np.asarray([float(r[lc]) for r in rows])
        p = np.asarray([float(r[pc]) for r in rows])
        metric = self.getOrDefault(self.metricName)
        if metric == "accuracy":
            return float((y == p).mean()) if len(y) else 0.0
        classes = np.unique(y)
        w = np.asarray([(y == c).mean() for c in classes])
        prec = np.asarray([((p == c) & (y == c)).sum() / max((p == c).sum(), 1) for c in classes])
        rec = np.asarray([((p == c) & (y == c)).sum() / max((y == c).sum(), 1) for c in classes])
        if metric == "weightedPrecision":
            return float((w * prec).sum())
        if metric == "weightedRecall":
            return float((w * rec).sum())
        f1 = np.where(prec + rec > 0, 2 * prec * rec / np.maximum(prec + rec, 1e-30), 0.0)
        return float((w * f1).sum())


class RegressionEvaluator(Evaluator, HasLabelCol, HasPredictionCol):
    metricName = Param(Params._dummy(), "metricName", "metric name in evaluation (rmse|mse|r2|mae)", typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, predictionCol="prediction", labelCol="label", metricName="rmse"):
        super().__init__()
        self._setDefault(predictionCol="prediction", labelCol="label", metricName="rmse")
        self._set(**self._input_kwargs)

    def _evaluate(self, dataset) -> float:
        lc, pc = self.getLabelCol(), self.getPredictionCol()
        rows = dataset.collect()
        y = np.asarray([float(r[lc]) for r in rows])
        p = np.asarray([float(r[pc]) for r in rows])
        metric = self.getOrDefault(self.metricName)
        err = y - p
        if metric == "mse":
            return float(np.mean(err ** 2))
        if metric == "mae":
            return float(np.mean(np.abs(err)))
        if metric == "r2":
            ss_tot = float(np.sum((y - y.mean()) ** 2))
            return float(1.0 - np.sum(err ** 2) / ss_tot) if ss_tot > 0 else 0.0
        return float(np.sqrt(np.mean(err ** 2)))

    def isLargerBetter(self) -> bool:
        return self.getOrDefault(self.metricName) == "r2"


class BinaryClassificationEvaluator(Evaluator, HasLabelCol):
    """``rawPredictionCol`` may hold a score, a probability, or a 2-element vector (score of class 1 = last element)."""
    rawPredictionCol = Param(Params._dummy(), "rawPredictionCol", "raw prediction (a.k.a. confidence) column name.",
                             typeConverter=TypeConverters.toString)
    metricName = Param(Params._dummy(), "metricName", "metric name in evaluation (areaUnderROC|areaUnderPR)", typeConverter=TypeConverters.toString)

    @keyword_only
    def __init__(self, rawPredictionCol="rawPrediction", labelCol="label", metricName="areaUnderROC"):
        super().__init__()
        self._setDefault(rawPredictionCol="rawPrediction", labelCol="label", metricName="areaUnderROC")
        self._set(**self._input_kwargs)

    def _evaluate(self, dataset) -> float:
        lc, rc = self.getLabelCol(), self.getOrDefault(self.rawPredictionCol)
        rows = dataset.collect()
        y = np.asarray([float(r[lc]) for r in rows])

        def score(v):
            if isinstance(v, Vector):
                return float(v.toArray()[-1])
            if isinstance(v, (list, tuple, np.ndarray)):
                return float(np.asarray(v, dtype=np.float64).reshape(-1)[-1])
            return float(v)

        s = np.asarray([score(r[rc]) for r in rows])
        order = np.argsort(-s, kind="mergesort")
        y, s = y[order], s[order]
        pos, neg = float((y > 0.5).sum()), float((y <= 0.5).sum())
        if pos == 0 or neg == 0:
            return 0.0
        # thresholds at distinct scores
        last = np.r_[np.nonzero(np.diff(s))[0], len(s) - 1]
        tp = np.cumsum(y > 0.5)[last].astype(np.float64)
        fp = np.cumsum(y <= 0.5)[last].astype(np.float64)
        if self.getOrDefault(self.metricName) == "areaUnderPR":
            prec = tp / np.maximum(tp + fp, 1.0)
            rec = tp / pos
            prec, rec = np.r_[prec[0], prec], np.r_[0.0, rec]
            return float(np.sum(np.diff(rec) * (prec[1:] + prec[:-1]) / 2.0))
        tpr, fpr = np.r_[0.0, tp / pos], np.r_[0.0, fp / neg]
        return float(np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) / 2.0))
