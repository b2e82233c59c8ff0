"""Drop-in aliases so lifeomic/sparkflow user code runs unchanged.

``install()`` registers, for every package that is NOT importable in this environment:

* ``pyspark`` (+ ``pyspark.sql``, ``pyspark.sql.functions``, ``pyspark.ml``, ``pyspark.ml.feature``,
  ``pyspark.ml.linalg``, ``pyspark.ml.param``, ``pyspark.ml.param.shared``, ``pyspark.ml.base``,
  ``pyspark.ml.util``, ``pyspark.ml.pipeline``, ``pyspark.ml.evaluation``, ``pyspark.context``)
  -> ``sparkflow_b200.spark``
* ``tensorflow``  -> ``sparkflow_b200.graph.tfcompat``
* ``sparkflow`` (+ the reference's seven module names) -> the same-named ``sparkflow_b200`` modules

Real installations always win: nothing is overridden if the genuine package imports.
"""
from __future__ import annotations

import importlib
import sys
import types


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        importlib.import_module(name)
        return False
    except Exception:
        return True


def _alias(name: str, module: types.ModuleType) -> None:
    sys.modules[name] = module


def _ns(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install(force: bool = False) -> dict:
    done = {}
    if force or _missing("pyspark"):
        from . import spark
        from .spark import context, sql
        from .spark.ml import base, evaluation, feature, linalg, param

        pyspark = _ns("pyspark", __sparkflow_shim__=True, SparkContext=context.SparkContext, SparkConf=context.SparkConf, keyword_only=context.keyword_only,
                      __version__=spark.__version__, __path__=[])
        sql_mod = _ns("pyspark.sql", SparkSession=sql.SparkSession, DataFrame=sql.DataFrame, Row=sql.Row, __path__=[])
        fn_mod = _ns("pyspark.sql.functions", rand=sql.rand, col=sql.col)
        ml_mod = _ns("pyspark.ml", Pipeline=base.Pipeline, PipelineModel=base.PipelineModel, Estimator=base.Estimator,
                     Model=base.Model, Transformer=base.Transformer, __path__=[])
        util = _ns("pyspark.ml.util", Identifiable=param.Identifiable, MLReadable=base.MLReadable, MLWritable=base.MLWritable,
                   MLReader=base.MLReader, MLWriter=base.MLWriter, JavaMLReader=base.MLReader, JavaMLWriter=base.MLWriter)
        shared = _ns("pyspark.ml.param.shared", HasInputCol=param.HasInputCol, HasOutputCol=param.HasOutputCol,
                     HasLabelCol=param.HasLabelCol, HasPredictionCol=param.HasPredictionCol, HasFeaturesCol=param.HasFeaturesCol,
                     HasInputCols=param.HasInputCols)
        param_mod = _ns("pyspark.ml.param", Param=param.Param, Params=param.Params, TypeConverters=param.TypeConverters, __path__=[])
        pipeline_mod = _ns("pyspark.ml.pipeline", Pipeline=base.Pipeline, PipelineModel=base.PipelineModel)
        for name, mod in {"pyspark": pyspark, "pyspark.context": context, "pyspark.sql": sql_mod, "pyspark.sql.functions": fn_mod,
                          "pyspark.ml": ml_mod, "pyspark.ml.base": base, "pyspark.ml.util": util, "pyspark.ml.param": param_mod,
                          "pyspark.ml.param.shared": shared, "pyspark.ml.feature": feature, "pyspark.ml.linalg": linalg,
                          "pyspark.ml.pipeline": pipeline_mod, "pyspark.ml.evaluation": evaluation}.items():
            _alias(name, mod)
        pyspark.sql, pyspark.ml, pyspark.context = sql_mod, ml_mod, context
        sql_mod.functions = fn_mod
        ml_mod.feature, ml_mod.linalg, ml_mod.param, ml_mod.util = feature, linalg, param_mod, util
        ml_mod.pipeline, ml_mod.evaluation, ml_mod.base = pipeline_mod, evaluation, base
        param_mod.shared = shared
        done["pyspark"] = "sparkflow_b200.spark"
    if force or _missing("tensorflow"):
        from .graph import tfcompat

        tfcompat._sparkflow_shim = True
        _alias("tensorflow", tfcompat)
        done["tensorflow"] = "sparkflow_b200.graph.tfcompat"
    if force or _missing("sparkflow"):
        import sparkflow_b200 as pkg
        from . import HogwildSparkModel, RWLock, graph_utils, ml_util, pipeline_util, tensorflow_async, tensorflow_model_loader

        _alias("sparkflow", pkg)
        for m in (HogwildSparkModel, RWLock, graph_utils, ml_util, pipeline_util, tensorflow_async, tensorflow_model_loader):
            _alias("sparkflow." + m.__name__.rsplit(".", 1)[-1], m)
        done["sparkflow"] = "sparkflow_b200"
    return done


class json_format:
    """``google.protobuf.json_format`` look-alike for tfcompat MetaGraphDefs (tests/dl_runner.py:200)."""

    @staticmethod
    def MessageToJson(message, **_unused) -> str:
        if hasattr(message, "to_json"):
            return message.to_json()
        from google.protobuf import json_format as real  # pragma: no cover

        return real.MessageToJson(message)

    @staticmethod
    def Parse(text, message=None, **_unused):
        import json

        from .graph.tfcompat import MetaGraphDef

        mg = MetaGraphDef(json.loads(text))
        if message is not None and hasattr(message, "data"):
            message.data = mg.data
            return message
        return mg
