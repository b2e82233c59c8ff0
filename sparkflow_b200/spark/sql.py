"""In-process, partitioned ``DataFrame`` / ``RDD`` / ``Row`` / ``SparkSession``."""
from __future__ import annotations

import random as _random
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Union

import numpy as np

from .context import SparkConf, SparkContext, _split


class Row(tuple):
    """``pyspark.sql.Row``: a tuple with named fields (``row.f``, ``row['f']``, ``asDict()``)."""

    def __new__(cls, *args, **kwargs):
        if args and kwargs:
            raise ValueError("Can not use both args and kwargs to create Row")
        if kwargs:
            names = sorted(kwargs.keys())          # Spark 2.4 sorts keyword fields
            row = tuple.__new__(cls, [kwargs[n] for n in names])
            row.__fields__ = names
            return row
        return tuple.__new__(cls, args)

    @classmethod
    def _make(cls, names: Sequence[str], values: Sequence[Any]) -> "Row":
        row = tuple.__new__(cls, values)
        row.__fields__ = list(names)
        return row

    def asDict(self, recursive: bool = False) -> Dict[str, Any]:
        if not hasattr(self, "__fields__"):
            raise TypeError("Cannot convert a Row class into dict")
        return dict(zip(self.__fields__, self))

    def __getitem__(self, item):
        if isinstance(item, (int, slice)):
            return super().__getitem__(item)
        try:
            return super().__getitem__(self.__fields__.index(item))
        except (AttributeError, ValueError):
            raise ValueError(item)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        try:
            return self[self.__fields__.index(item)]
        except (AttributeError, ValueError):
            raise AttributeError(item)

    def __contains__(self, item):
        return item in getattr(self, "__fields__", ()) or super().__contains__(item)

    def __repr__(self):
        if hasattr(self, "__fields__"):
            return "Row(%s)" % ", ".join("%s=%r" % kv for kv in zip(self.__fields__, tuple(self)))
        return "<Row(%s)>" % ", ".join(repr(v) for v in self)

    def __reduce__(self):
        if hasattr(self, "__fields__"):
            return (_rebuild_row, (self.__fields__, tuple(self)))
        return tuple.__reduce__(self)


def _rebuild_row(names, values):
    return Row._make(names, values)


class RDD:
    def __init__(self, partitions: List[List[Any]], ctx: Optional[SparkContext] = None):
        self._parts = partitions
        self.ctx = ctx or SparkContext.getOrCreate()

    # -- transformations (eager; data sets here are host-resident python objects) -------------------
    def map(self, f: Callable[[Any], Any]) -> "RDD":
        return RDD([[f(x) for x in p] for p in self._parts], self.ctx)

    def flatMap(self, f) -> "RDD":
        return RDD([[y for x in p for y in f(x)] for p in self._parts], self.ctx)

    def filter(self, f) -> "RDD":
        return RDD([[x for x in p if f(x)] for p in self._parts], self.ctx)

    def mapPartitions(self, f: Callable[[Iterator[Any]], Iterable[Any]], preservesPartitioning: bool = False) -> "RDD":
        return RDD([list(f(iter(p))) for p in self._parts], self.ctx)

    def mapPartitionsWithIndex(self, f, preservesPartitioning: bool = False) -> "RDD":
        return RDD([list(f(i, iter(p))) for i, p in enumerate(self._parts)], self.ctx)

    def glom(self) -> "RDD":
        return RDD([[list(p)] for p in self._parts], self.ctx)

    def coalesce(self, numPartitions: int, shuffle: bool = False) -> "RDD":
        n = max(1, int(numPartitions))
        if n >= len(self._parts):
            return self
        groups = _split(list(range(len(self._parts))), n)       # merge neighbouring partitions, no shuffle
        return RDD([[x for i in g for x in self._parts[i]] for g in groups], self.ctx)

    def repartition(self, numPartitions: int) -> "RDD":
        """Full shuffle: rows are redistributed at random over ``numPartitions`` partitions."""
        rows = [x for p in self._parts for x in p]
        _random.shuffle(rows)
        return RDD(_split(rows, max(1, int(numPartitions))), self.ctx)

    # -- actions --------------------------------------------------------------------------------------
    def getNumPartitions(self) -> int:
        return len(self._parts)

    def collect(self) -> List[Any]:
        return [x for p in self._parts for x in p]

    def take(self, n: int) -> List[Any]:
        out: List[Any] = []
        for p in self._parts:
            for x in p:
                if len(out) >= n:
                    return out
                out.append(x)
        return out

    def first(self):
        got = self.take(1)
        if not got:
            raise ValueError("RDD is empty")
        return got[0]

    def count(self) -> int:
        return sum(len(p) for p in self._parts)

    def foreach(self, f) -> None:
        for p in self._parts:
            for x in p:
                f(x)

    def foreachPartition(self, f: Callable[[Iterator[Any]], None]) -> None:
        """Partitions run concurrently, one task per partition (Spark ``local[N]``)."""
        workers = max(1, min(len(self._parts), self.ctx.defaultParallelism))
        if workers == 1 or len(self._parts) == 1:
            for p in self._parts:
                f(iter(p))
            return
        with ThreadPoolExecutor(max_workers=workers) as ex:
            for fut in [ex.submit(f, iter(p)) for p in self._parts]:
                fut.result()

    def toDF(self, schema: Optional[Sequence[str]] = None) -> "DataFrame":
        return DataFrame(self._parts, schema, self.ctx)

    def partitions(self) -> List[List[Any]]:
        return self._parts


class Column:
    def __init__(self, kind: str, name: Optional[str] = None, seed: Optional[int] = None):
        self.kind, self.name, self.seed = kind, name, seed


class DataFrame:
    def __init__(self, partitions: List[List[Any]], schema: Optional[Sequence[str]] = None, ctx: Optional[SparkContext] = None):
        self.ctx = ctx or SparkContext.getOrCreate()
        cols = list(schema) if schema is not None else None
        parts: List[List[Row]] = []
        for p in partitions:
            rows = []
            for r in p:
                if isinstance(r, Row) and hasattr(r, "__fields__"):
                    if cols is None:
                        cols = list(r.__fields__)
                    rows.append(r if list(r.__fields__) == cols else Row._make(cols, [r[c] for c in cols]))
                elif isinstance(r, dict):
                    if cols is None:
                        cols = sorted(r.keys())
                    rows.append(Row._make(cols, [r[c] for c in cols]))
                else:
                    vals = list(r) if isinstance(r, (tuple, list)) else [r]
                    if cols is None:
                        cols = ["_%d" % (i + 1) for i in range(len(vals))]
                    rows.append(Row._make(cols, vals))
            parts.append(rows)
        self._parts = parts
        self.columns: List[str] = cols or []

    # -- basic API ------------------------------------------------------------------------------------
    @property
    def rdd(self) -> RDD:
        return RDD(self._parts, self.ctx)

    @property
    def schema(self):
        return self.columns

    def count(self) -> int:
        return sum(len(p) for p in self._parts)

    def collect(self) -> List[Row]:
        return [r for p in self._parts for r in p]

    def take(self, n: int) -> List[Row]:
        return self.rdd.take(n)

    def head(self, n: Optional[int] = None):
        return self.take(1)[0] if n is None else self.take(n)

    def first(self) -> Row:
        return self.take(1)[0]

    def limit(self, n: int) -> "DataFrame":
        return DataFrame([self.take(n)], self.columns, self.ctx)

    def show(self, n: int = 20, truncate: bool = True) -> None:
        print(" | ".join(self.columns))
        for r in self.take(n):
            print(" | ".join((str(v)[:20] if truncate else str(v)) for v in r))

    def select(self, *cols) -> "DataFrame":
        names = [c for arg in cols for c in (arg if isinstance(arg, (list, tuple)) else [arg])]
        names = [c.name if isinstance(c, Column) else c for c in names]
        idx = [self.columns.index(c) for c in names]
        return DataFrame([[Row._make(names, [r[i] for i in idx]) for r in p] for p in self._parts], names, self.ctx)

    def drop(self, *cols) -> "DataFrame":
        return self.select([c for c in self.columns if c not in cols])

    def withColumn(self, name: str, fn_or_values) -> "DataFrame":
        cols = self.columns + ([name] if name not in self.columns else [])
        out = []
        for p in self._parts:
            rows = []
            for r in p:
                d = r.asDict()
                d[name] = fn_or_values(r) if callable(fn_or_values) else fn_or_values
                rows.append(Row._make(cols, [d[c] for c in cols]))
            out.append(rows)
        return DataFrame(out, cols, self.ctx)

    def withColumnRenamed(self, old: str, new: str) -> "DataFrame":
        cols = [new if c == old else c for c in self.columns]
        return DataFrame([[Row._make(cols, list(r)) for r in p] for p in self._parts], cols, self.ctx)

    def orderBy(self, *cols, ascending: bool = True) -> "DataFrame":
        rows = self.collect()
        if len(cols) == 1 and isinstance(cols[0], Column) and cols[0].kind == "rand":
            rng = _random.Random(cols[0].seed)
            rng.shuffle(rows)
        else:
            names = [c.name if isinstance(c, Column) else c for c in cols]
            rows.sort(key=lambda r: tuple(r[n] for n in names), reverse=not ascending)
        return DataFrame(_split(rows, len(self._parts)), self.columns, self.ctx)

    sort = orderBy

    def filter(self, f: Callable[[Row], bool]) -> "DataFrame":
        return DataFrame([[r for r in p if f(r)] for p in self._parts], self.columns, self.ctx)

    where = filter

    def coalesce(self, n: int) -> "DataFrame":
        return DataFrame(self.rdd.coalesce(n).partitions(), self.columns, self.ctx)

    def repartition(self, n: int) -> "DataFrame":
        return DataFrame(self.rdd.repartition(n).partitions(), self.columns, self.ctx)

    def randomSplit(self, weights: Sequence[float], seed: Optional[int] = None) -> List["DataFrame"]:
        rng = _random.Random(seed)
        total = float(sum(weights))
        bounds = np.cumsum([w / total for w in weights])
        buckets: List[List[Row]] = [[] for _ in weights]
        for r in self.collect():
            buckets[int(np.searchsorted(bounds, rng.random(), side="right").clip(0, len(weights) - 1))].append(r)
        return [DataFrame(_split(b, len(self._parts)), self.columns, self.ctx) for b in buckets]

    def cache(self) -> "DataFrame":
        return self

    persist = cache

    def toPandas(self):
        import pandas as pd

        return pd.DataFrame([list(r) for r in self.collect()], columns=self.columns)

    def __repr__(self):
        return "DataFrame[%s]" % ", ".join(self.columns)


class DataFrameReader:
    def __init__(self, session: "SparkSession"):
        self._session = session
        self._options: Dict[str, str] = {}

    def option(self, key: str, value: Any) -> "DataFrameReader":
        self._options[key.lower()] = str(value)
        return self

    def options(self, **kw) -> "DataFrameReader":
        for k, v in kw.items():
            self.option(k, v)
        return self

    def csv(self, path: str, header: Optional[bool] = None, inferSchema: Optional[bool] = None) -> DataFrame:
        """Numeric CSV reader (native parser).  Columns are named ``_c0.._cN`` like Spark; with
        ``inferSchema`` integral columns come back as ``int`` otherwise ``float``; without it, ``str``."""
        from ..io.csvio import read_numeric_csv

        has_header = (str(header).lower() == "true") if header is not None else self._options.get("header", "false") == "true"
        infer = (str(inferSchema).lower() == "true") if inferSchema is not None else self._options.get("inferschema", "false") == "true"
        arr, names = read_numeric_csv(path, has_header)
        cols = names or ["_c%d" % i for i in range(arr.shape[1])]
        if infer:
            integral = np.all(arr == np.round(arr), axis=0)
            rows = [Row._make(cols, [int(v) if integral[j] else float(v) for j, v in enumerate(r)]) for r in arr]
        else:
            rows = [Row._make(cols, [("%g" % v) for v in r]) for r in arr]
        return DataFrame(_split(rows, self._session.sparkContext.defaultParallelism), cols, self._session.sparkContext)


class SparkSession:
    _instantiated: Optional["SparkSession"] = None

    class Builder:
        def __init__(self):
            self._master = "local[*]"
            self._app = "sparkflow_b200"
            self._conf: Dict[str, str] = {}

        def master(self, m: str):
            self._master = m
            return self

        def appName(self, a: str):
            self._app = a
            return self

        def config(self, key: Optional[str] = None, value: Any = None, conf: Optional[SparkConf] = None):
            if conf is not None:
                self._conf.update(dict(conf.getAll()))
            elif key is not None:
                self._conf[key] = str(value)
            return self

        def enableHiveSupport(self):
            return self

        def getOrCreate(self) -> "SparkSession":
            if SparkSession._instantiated is None:
                SparkSession._instantiated = SparkSession(SparkContext(self._master, self._app, SparkConf(self._conf)))
            return SparkSession._instantiated

    class _BuilderAccessor:
        def __get__(self, obj, owner):
            return SparkSession.Builder()

    builder = _BuilderAccessor()

    def __init__(self, sc: Optional[SparkContext] = None):
        self.sparkContext = sc or SparkContext.getOrCreate()
        self.read = DataFrameReader(self)

    def createDataFrame(self, data, schema: Optional[Sequence[str]] = None, numPartitions: Optional[int] = None) -> DataFrame:
        if isinstance(data, RDD):
            return DataFrame(data.partitions(), schema, self.sparkContext)
        try:
            import pandas as pd

            if isinstance(data, pd.DataFrame):
                schema = schema or list(data.columns)
                data = [tuple(r) for r in data.itertuples(index=False)]
        except ImportError:  # pragma: no cover
            pass
        rows = list(data)
        n = numPartitions or self.sparkContext.defaultParallelism
        return DataFrame(_split(rows, max(1, n)), schema, self.sparkContext)

    def stop(self):
        self.sparkContext.stop()
        SparkSession._instantiated = None

    @property
    def conf(self):
        return self.sparkContext.getConf()


# pyspark.sql.functions ---------------------------------------------------------------------------
def rand(seed: Optional[int] = None) -> Column:
    return Column("rand", seed=seed)


def col(name: str) -> Column:
    return Column("col", name=name)
