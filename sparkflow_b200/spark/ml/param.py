"""``pyspark.ml.param``: Param / Params / TypeConverters with Spark's semantics (defaults vs set
values, ``getOrDefault``, ``_set`` type conversion, ``copy``, ``extractParamMap``)."""
from __future__ import annotations

import copy as _copy
import uuid
from typing import Any, Dict, List, Optional

import numpy as np


class TypeConverters:
    @staticmethod
    def identity(v):
        return v

    @staticmethod
    def toString(v):
        if isinstance(v, str):
            return v
        raise TypeError("Could not convert %r to string type" % (v,))

    @staticmethod
    def toInt(v):
        if isinstance(v, bool):
            raise TypeError("Could not convert %r to int" % (v,))
        if isinstance(v, (int, np.integer)) or (isinstance(v, (float, np.floating)) and float(v).is_integer()):
            return int(v)
        raise TypeError("Could not convert %r to int" % (v,))

    @staticmethod
    def toFloat(v):
        if isinstance(v, bool):
            raise TypeError("Could not convert %r to float" % (v,))
        if isinstance(v, (int, float, np.integer, np.floating)):
            return float(v)
        raise TypeError("Could not convert %r to float" % (v,))

    @staticmethod
    def toBoolean(v):
        if isinstance(v, (bool, np.bool_)):
            return bool(v)
        raise TypeError("Boolean Param requires value of type bool. Found %s." % type(v))

    @staticmethod
    def toList(v):
        if isinstance(v, (list, tuple, np.ndarray)):
            return list(v)
        raise TypeError("Could not convert %r to list" % (v,))

    @staticmethod
    def toListString(v):
        return [TypeConverters.toString(x) for x in TypeConverters.toList(v)]

    @staticmethod
    def toListFloat(v):
        return [TypeConverters.toFloat(x) for x in TypeConverters.toList(v)]

    @staticmethod
    def toListInt(v):
        return [TypeConverters.toInt(x) for x in TypeConverters.toList(v)]


class Param:
    def __init__(self, parent, name: str, doc: str, typeConverter=None):
        self.parent = parent.uid if isinstance(parent, Identifiable) else parent
        self.name, self.doc = str(name), str(doc)
        self.typeConverter = typeConverter or TypeConverters.identity

    def _copy_new_parent(self, parent):
        if self.parent == "undefined":
            p = _copy.copy(self)
            p.parent = parent.uid
            return p
        raise ValueError("Cannot copy from non-dummy parent %s." % self.parent)

    def __str__(self):
        return "%s__%s" % (self.parent, self.name)

    __repr__ = lambda self: "Param(parent=%r, name=%r, doc=%r)" % (self.parent, self.name, self.doc)

    def __hash__(self):
        return hash(str(self))

    def __eq__(self, other):
        return isinstance(other, Param) and self.parent == other.parent and self.name == other.name


class Identifiable:
    def __init__(self):
        self.uid = self._randomUID()

    @classmethod
    def _randomUID(cls) -> str:
        return "%s_%s" % (cls.__name__, uuid.uuid4().hex[-12:])

    def __repr__(self):
        return self.uid


class _Dummy(Identifiable):
    def __init__(self):
        self.uid = "undefined"


class Params(Identifiable):
    def __init__(self):
        super().__init__()
        self._paramMap: Dict[Param, Any] = {}
        self._defaultParamMap: Dict[Param, Any] = {}
        self._params: Optional[List[Param]] = None
        self._copy_params()

    @staticmethod
    def _dummy():
        return _Dummy()

    def _copy_params(self):
        cls = type(self)
        for name in dir(cls):
            attr = getattr(cls, name, None)
            if isinstance(attr, Param):
                setattr(self, name, attr._copy_new_parent(self))

    @property
    def params(self) -> List[Param]:
        if self._params is None:
            self._params = sorted([getattr(self, n) for n in dir(self) if n != "params" and not n.startswith("__")
                                   and isinstance(getattr(type(self), n, None), Param)], key=lambda p: p.name)
        return self._params

    def hasParam(self, name: str) -> bool:
        return isinstance(getattr(self, name, None), Param)

    def getParam(self, name: str) -> Param:
        p = getattr(self, name, None)
        if isinstance(p, Param):
            return p
        raise ValueError("Cannot find param with name %s." % name)

    def _resolveParam(self, param) -> Param:
        if isinstance(param, Param):
            if param.parent != self.uid:
                raise ValueError("Param %r does not belong to %r." % (param, self))
            return param
        return self.getParam(param)

    def isSet(self, param) -> bool:
        return self._resolveParam(param) in self._paramMap

    def hasDefault(self, param) -> bool:
        return self._resolveParam(param) in self._defaultParamMap

    def isDefined(self, param) -> bool:
        return self.isSet(param) or self.hasDefault(param)

    def getOrDefault(self, param):
        p = self._resolveParam(param)
        if p in self._paramMap:
            return self._paramMap[p]
        if p in self._defaultParamMap:
            return self._defaultParamMap[p]
        raise KeyError("Failed to find a default value for %s" % p.name)

    def getDefault(self, param):
        return self._defaultParamMap[self._resolveParam(param)]

    def extractParamMap(self, extra: Optional[Dict[Param, Any]] = None) -> Dict[Param, Any]:
        m = dict(self._defaultParamMap)
        m.update(self._paramMap)
        m.update(extra or {})
        return m

    def explainParams(self) -> str:
        return "\n".join("%s: %s" % (p.name, p.doc) for p in self.params)

    def set(self, param, value):
        p = self._resolveParam(param)
        self._paramMap[p] = p.typeConverter(value) if value is not None else None
        return self

    def _set(self, **kwargs):
        for name, value in kwargs.items():
            p = self.getParam(name)
            if value is not None:
                try:
                    value = p.typeConverter(value)
                except TypeError as e:
                    raise TypeError('Invalid param value given for param "%s". %s' % (p.name, e))
            self._paramMap[p] = value
        return self

    def _setDefault(self, **kwargs):
        for name, value in kwargs.items():
            p = self.getParam(name)
            if value is not None and not isinstance(value, Params):
                try:
                    value = p.typeConverter(value)
                except TypeError as e:
                    raise TypeError('Invalid default param value given for param "%s". %s' % (p.name, e))
            self._defaultParamMap[p] = value
        return self

    def clear(self, param):
        self._paramMap.pop(self._resolveParam(param), None)

    def copy(self, extra: Optional[Dict[Param, Any]] = None):
        that = _copy.copy(self)
        that._paramMap = dict(self._paramMap)
        that._defaultParamMap = dict(self._defaultParamMap)
        for p, v in (extra or {}).items():
            that._paramMap[that.getParam(p.name)] = v
        return that

    def _resetUid(self, newUid: str):
        self.uid = str(newUid)
        pm, dm = {}, {}
        for p in list(self.params):
            q = _copy.copy(p)
            q.parent = self.uid
            if p in self._paramMap:
                pm[q] = self._paramMap[p]
            if p in self._defaultParamMap:
                dm[q] = self._defaultParamMap[p]
            setattr(self, p.name, q)
        self._paramMap, self._defaultParamMap, self._params = pm, dm, None
        return self


# pyspark.ml.param.shared ---------------------------------------------------------------------------
def _shared(cls_name: str, pname: str, doc: str, conv, getter: str, default=None):
    def getter_fn(self):
        return self.getOrDefault(getattr(self, pname))

    def setter_fn(self, value):
        return self._set(**{pname: value})

    def init(self):
        Params.__init__(self)

    ns = {pname: Param(Params._dummy(), pname, doc, typeConverter=conv), getter: getter_fn,
          "set" + getter[3:]: setter_fn, "__init__": init}
    return type(cls_name, (Params,), ns)


HasInputCol = _shared("HasInputCol", "inputCol", "input column name.", TypeConverters.toString, "getInputCol")
HasOutputCol = _shared("HasOutputCol", "outputCol", "output column name.", TypeConverters.toString, "getOutputCol")
HasLabelCol = _shared("HasLabelCol", "labelCol", "label column name.", TypeConverters.toString, "getLabelCol")
HasPredictionCol = _shared("HasPredictionCol", "predictionCol", "prediction column name.", TypeConverters.toString, "getPredictionCol")
HasFeaturesCol = _shared("HasFeaturesCol", "featuresCol", "features column name.", TypeConverters.toString, "getFeaturesCol")
HasInputCols = _shared("HasInputCols", "inputCols", "input column names.", TypeConverters.toListString, "getInputCols")
