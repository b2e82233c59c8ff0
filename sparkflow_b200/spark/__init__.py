"""A dependency-free stand-in for the slice of PySpark that lifeomic/sparkflow builds on.

The reference is a ``pyspark.ml`` Estimator/Model (tensorflow_async.py:51-121) fed by DataFrames whose
partitions become workers.  PySpark (and a JVM) are not installable here, so this package implements
the used surface with the same names and behaviour: ``SparkSession`` / ``DataFrame`` / ``RDD`` / ``Row``
(partitioned, in-process), the ``Param`` system with ``keyword_only``, ``Pipeline`` / ``PipelineModel``,
the feature stages used by the examples, and Spark's on-disk ML persistence layout.
``sparkflow_b200.compat.install()`` aliases it to ``pyspark`` when the real one is missing.
"""
from .context import SparkConf, SparkContext, keyword_only  # noqa: F401
from .sql import DataFrame, RDD, Row, SparkSession  # noqa: F401

__version__ = "2.4.3-sparkflow_b200"
