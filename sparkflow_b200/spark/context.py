from __future__ import annotations

import functools
import re
from typing import Any, Dict, Optional


def keyword_only(func):
    """Same contract as ``pyspark.keyword_only``: force kwargs and stash them in ``_input_kwargs``."""

    @functools.wraps(func)
    def wrapper(self, *args, **kwargs):
        if len(args) > 0:
            raise TypeError("Method %s forces keyword arguments." % func.__name__)
        self._input_kwargs = kwargs
        return func(self, **kwargs)

    return wrapper


class SparkConf:
    def __init__(self, values: Optional[Dict[str, str]] = None):
        self._v: Dict[str, str] = dict(values or {})

    def set(self, key: str, value: Any) -> "SparkConf":
        self._v[key] = str(value)
        return self

    def get(self, key: str, default: Optional[str] = None) -> Optional[str]:
        return self._v.get(key, default)

    def getAll(self):
        return list(self._v.items())


class SparkContext:
    _active_spark_context: Optional["SparkContext"] = None

    def __init__(self, master: str = "local[*]", appName: str = "sparkflow_b200", conf: Optional[SparkConf] = None):
        self.master, self.appName = master, appName
        self._conf = conf or SparkConf()
        self._conf.set("spark.driver.host", self._conf.get("spark.driver.host", "127.0.0.1"))
        self._conf.set("spark.master", master)
        m = re.match(r"local\[(\d+|\*)\]", master or "")
        if m and m.group(1) != "*":
            self.defaultParallelism = int(m.group(1))
        else:
            import os

            self.defaultParallelism = max(1, min(os.cpu_count() or 2, 8)) if m else 2
        SparkContext._active_spark_context = self

    def getConf(self) -> SparkConf:
        return self._conf

    def parallelize(self, data, numSlices: Optional[int] = None):
        from .sql import RDD

        data = list(data)
        n = max(1, numSlices or self.defaultParallelism)
        return RDD(_split(data, n), self)

    def stop(self):
        if SparkContext._active_spark_context is self:
            SparkContext._active_spark_context = None

    @classmethod
    def getOrCreate(cls, conf: Optional[SparkConf] = None) -> "SparkContext":
        return cls._active_spark_context or cls(conf=conf)


def _split(data, n: int):
    """Contiguous, near-equal slices (Spark's ParallelCollectionRDD slicing)."""
    total = len(data)
    return [data[(i * total) // n:((i + 1) * total) // n] for i in range(n)]
