from __future__ import annotations

import numpy as np

from ..context import keyword_only
from .base import Evaluator
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
