"""``pyspark.ml`` base classes, ``Pipeline`` / ``PipelineModel`` and Spark's ML persistence layout.

On-disk layout (Spark 2.4 ``DefaultParamsWriter`` / ``Pipeline.SharedReadWrite``):

    <path>/metadata/part-00000   one JSON line {"class","timestamp","sparkVersion","uid","paramMap","defaultParamMap"}
    <path>/metadata/_SUCCESS
    <path>/stages/<idx>_<uid>/   (pipelines only; idx zero-padded to len(str(numStages)))
"""
from __future__ import annotations

import importlib
import json
import os
import shutil
import time
from typing import Any, Dict, List, Optional

from .param import Identifiable, Param, Params

SPARK_VERSION = "2.4.3"


# ---------------------------------------------------------------------------------------------
# persistence helpers
# ---------------------------------------------------------------------------------------------
def _write_metadata(path: str, java_class: str, uid: str, param_map: Dict[str, Any], default_map: Optional[Dict[str, Any]] = None) -> None:
    meta_dir = os.path.join(path, "metadata")
    os.makedirs(meta_dir, exist_ok=True)
    meta = {"class": java_class, "timestamp": int(time.time() * 1000), "sparkVersion": SPARK_VERSION, "uid": uid,
            "paramMap": param_map, "defaultParamMap": default_map or {}}
    with open(os.path.join(meta_dir, "part-00000"), "w") as fh:
        fh.write(json.dumps(meta, separators=(",", ":")) + "\n")
    open(os.path.join(meta_dir, "_SUCCESS"), "w").close()


def _read_metadata(path: str) -> Dict[str, Any]:
    meta_dir = os.path.join(path, "metadata")
    parts = sorted(f for f in os.listdir(meta_dir) if f.startswith("part-"))
    if not parts:
        raise IOError(f"no ML metadata under {meta_dir}")
    with open(os.path.join(meta_dir, parts[0])) as fh:
        return json.loads(fh.readline())


class MLWriter:
    def __init__(self, instance):
        self.instance = instance
        self.shouldOverwrite = False

    def overwrite(self) -> "MLWriter":
        self.shouldOverwrite = True
        return self

    def option(self, key, value):
        return self

    def session(self, sparkSession):
        return self

    def context(self, sqlContext):
        return self

    def save(self, path: str) -> None:
        if os.path.exists(path):
            if not self.shouldOverwrite:
                raise IOError(f"Path {path} already exists. To overwrite it, please use write.overwrite().save(path) for Scala and use "
                              "write().overwrite().save(path) for Java and Python.")
            shutil.rmtree(path)
        self.saveImpl(path)

    def saveImpl(self, path: str) -> None:
        self.instance._save_impl(path)


class MLReader:
    def __init__(self, clazz):
        self.clazz = clazz

    def session(self, sparkSession):
        return self

    def load(self, path: str):
        return self.clazz._load_impl(path)


class MLWritable:
    def write(self) -> MLWriter:
        return MLWriter(self)

    def save(self, path: str) -> None:
        self.write().save(path)


class MLReadable:
    @classmethod
    def read(cls) -> MLReader:
        return MLReader(cls)

    @classmethod
    def load(cls, path: str):
        return cls.read().load(path)


# registry: JVM class name -> python class (what py4j + JavaParams._from_java do in real Spark)
_JAVA_CLASSES: Dict[str, type] = {}


def java_class(name: str):
    def deco(cls):
        cls._java_class_name = name
        _JAVA_CLASSES[name] = cls
        return cls
    return deco


def load_stage(path: str):
    meta = _read_metadata(path)
    cls = _JAVA_CLASSES.get(meta["class"])
    if cls is None:
        mod, _, name = meta["class"].rpartition(".")
        try:
            cls = getattr(importlib.import_module(mod), name)
        except Exception:
            raise ValueError(f"cannot load ML stage of class {meta['class']}")
    return cls._load_impl(path)


class DefaultParamsPersistence(MLWritable, MLReadable):
    """Param-only stages: everything lives in metadata (Spark ``DefaultParamsWritable``)."""

    _java_class_name = ""

    def _param_json(self, m: Dict[Param, Any]) -> Dict[str, Any]:
        out = {}
        for p, v in m.items():
            try:
                json.dumps(v)
                out[p.name] = v
            except TypeError:
                out[p.name] = str(v)
        return out

    def _save_impl(self, path: str) -> None:
        _write_metadata(path, self._java_class_name or f"{type(self).__module__}.{type(self).__name__}", self.uid,
                        self._param_json(self._paramMap), self._param_json(self._defaultParamMap))

    @classmethod
    def _load_impl(cls, path: str):
        meta = _read_metadata(path)
        inst = cls()
        inst._resetUid(meta["uid"])
        for k, v in meta.get("defaultParamMap", {}).items():
            if inst.hasParam(k):
                inst._setDefault(**{k: v})
        for k, v in meta.get("paramMap", {}).items():
            if inst.hasParam(k):
                inst._set(**{k: v})
        return inst


# ---------------------------------------------------------------------------------------------
# pipeline API
# ---------------------------------------------------------------------------------------------
class Transformer(Params):
    def transform(self, dataset, params: Optional[Dict[Param, Any]] = None):
        if params:
            return self.copy(params)._transform(dataset)
        return self._transform(dataset)

    def _transform(self, dataset):  # pragma: no cover
        raise NotImplementedError


class Estimator(Params):
    def fit(self, dataset, params: Optional[Dict[Param, Any]] = None):
        if isinstance(params, (list, tuple)):
            return [self.fit(dataset, p) for p in params]
        if params:
            return self.copy(params)._fit(dataset)
        return self._fit(dataset)

    def _fit(self, dataset):  # pragma: no cover
        raise NotImplementedError


class Model(Transformer):
    pass


class Evaluator(Params):
    def evaluate(self, dataset, params: Optional[Dict[Param, Any]] = None) -> float:
        if params:
            return self.copy(params)._evaluate(dataset)
        return self._evaluate(dataset)

    def isLargerBetter(self) -> bool:
        return True


def _stage_dir(stages_dir: str, idx: int, n: int, uid: str) -> str:
    return os.path.join(stages_dir, "%0*d_%s" % (len(str(n)), idx, uid))


def _save_stages(path: str, java_cls: str, uid: str, stages: List[Any]) -> None:
    # stage uids come from the persisted object (carrier-wrapped python stages keep their own uid)
    _write_metadata(path, java_cls, uid, {"stageUids": [s.uid for s in stages]})
    stages_dir = os.path.join(path, "stages")
    os.makedirs(stages_dir, exist_ok=True)
    for i, st in enumerate(stages):
        st.write().save(_stage_dir(stages_dir, i, len(stages), st.uid))


def _load_stages(path: str):
    meta = _read_metadata(path)
    uids = meta["paramMap"]["stageUids"]
    stages_dir = os.path.join(path, "stages")
    return meta, [load_stage(_stage_dir(stages_dir, i, len(uids), u)) for i, u in enumerate(uids)]


@java_class("org.apache.spark.ml.Pipeline")
class Pipeline(Estimator, MLWritable, MLReadable):
    stages = Param(Params._dummy(), "stages", "a list of pipeline stages")

    def __init__(self, stages: Optional[List[Any]] = None):
        super().__init__()
        if stages is not None:
            self.setStages(stages)

    def setStages(self, value: List[Any]) -> "Pipeline":
        return self._set(stages=list(value))

    def getStages(self) -> List[Any]:
        return self.getOrDefault(self.stages)

    def _fit(self, dataset) -> "PipelineModel":
        stages = self.getStages()
        for st in stages:
            if not isinstance(st, (Estimator, Transformer)):
                raise TypeError("Cannot recognize a pipeline stage of type %s." % type(st))
        last_est = max([i for i, s in enumerate(stages) if isinstance(s, Estimator)], default=-1)
        fitted: List[Transformer] = []
        for i, st in enumerate(stages):
            if i <= last_est:
                if isinstance(st, Transformer):
                    fitted.append(st)
                    dataset = st.transform(dataset)
                else:
                    model = st.fit(dataset)
                    fitted.append(model)
                    if i < last_est:
                        dataset = model.transform(dataset)
            else:
                fitted.append(st)
        pm = PipelineModel(fitted)
        pm._resetUid(self.uid)
        return pm

    def copy(self, extra=None):
        that = Params.copy(self, extra)
        return that.setStages([s.copy(extra) for s in self.getStages()])

    def _save_impl(self, path: str) -> None:
        _save_stages(path, "org.apache.spark.ml.Pipeline", self.uid, self.getStages())

    @classmethod
    def _load_impl(cls, path: str) -> "Pipeline":
        meta, stages = _load_stages(path)
        p = cls(stages)
        p._resetUid(meta["uid"])
        return p


@java_class("org.apache.spark.ml.PipelineModel")
class PipelineModel(Model, MLWritable, MLReadable):
    def __init__(self, stages: List[Transformer]):
        super().__init__()
        self.stages = list(stages)

    def _transform(self, dataset):
        for t in self.stages:
            dataset = t.transform(dataset)
        return dataset

    def copy(self, extra=None):
        return PipelineModel([s.copy(extra) for s in self.stages])

    def _save_impl(self, path: str) -> None:
        _save_stages(path, "org.apache.spark.ml.PipelineModel", self.uid, self.stages)

    @classmethod
    def _load_impl(cls, path: str) -> "PipelineModel":
        meta, stages = _load_stages(path)
        pm = cls(stages)
        pm._resetUid(meta["uid"])
        return pm
