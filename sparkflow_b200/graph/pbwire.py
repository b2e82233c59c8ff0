"""Schema-driven protobuf wire codec for the handful of TensorFlow messages the framework touches.

TensorFlow is not installable here, yet the artefact formats are part of the public contract
(SURVEY.md section 2.5): MetaGraphDef (binary ``.meta`` and proto3-JSON), ``VariableDef`` bytes in
``collection_def``, and the TF-V2 checkpoint protos.  This module decodes binary protobuf into the
same dict shape ``google.protobuf.json_format.MessageToJson`` produces (lowerCamelCase keys, int64
as strings, bytes as base64, enums by name) and encodes such dicts back to binary.

Field numbers follow tensorflow/core/{framework,protobuf}/*.proto (TF 1.x).
"""
from __future__ import annotations

import base64
import struct
from typing import Any, Dict, List, Tuple

# ---------------------------------------------------------------------------
# enums
# ---------------------------------------------------------------------------
DATA_TYPE = {
    0: "DT_INVALID", 1: "DT_FLOAT", 2: "DT_DOUBLE", 3: "DT_INT32", 4: "DT_UINT8", 5: "DT_INT16", 6: "DT_INT8",
    7: "DT_STRING", 8: "DT_COMPLEX64", 9: "DT_INT64", 10: "DT_BOOL", 14: "DT_BFLOAT16", 19: "DT_HALF",
    20: "DT_RESOURCE", 101: "DT_FLOAT_REF", 103: "DT_INT32_REF", 109: "DT_INT64_REF",
}
ENUMS: Dict[str, Dict[int, str]] = {
    "DataType": DATA_TYPE,
    "SaverVersion": {0: "LEGACY", 1: "V1", 2: "V2"},
    "Endianness": {0: "LITTLE", 1: "BIG"},
}
_ENUM_REV = {k: {v: i for i, v in d.items()} for k, d in ENUMS.items()}

# ---------------------------------------------------------------------------
# schemas: message -> {field_number: (json_name, type, repeated)}
# scalar types: int32 int64 uint64 bool float double string bytes fixed32 enum:<Name> ; msg:<Name> ;
# map:<keytype>,<valuetype>
# ---------------------------------------------------------------------------
S: Dict[str, Dict[int, Tuple[str, str, bool]]] = {
    "MetaGraphDef": {
        1: ("metaInfoDef", "msg:MetaInfoDef", False),
        2: ("graphDef", "msg:GraphDef", False),
        3: ("saverDef", "msg:SaverDef", False),
        4: ("collectionDef", "map:string,msg:CollectionDef", False),
        5: ("signatureDef", "map:string,msg:Opaque", False),
        6: ("assetFileDef", "msg:Opaque", True),
    },
    "MetaInfoDef": {
        1: ("metaGraphVersion", "string", False),
        2: ("strippedOpList", "msg:OpList", False),
        3: ("anyInfo", "msg:Opaque", False),
        4: ("tags", "string", True),
        5: ("tensorflowVersion", "string", False),
        6: ("tensorflowGitVersion", "string", False),
        7: ("strippedDefaultAttrs", "bool", False),
    },
    "OpList": {1: ("op", "msg:OpDef", True)},
    "OpDef": {
        1: ("name", "string", False),
        2: ("inputArg", "msg:ArgDef", True),
        3: ("outputArg", "msg:ArgDef", True),
        4: ("attr", "msg:OpAttrDef", True),
        5: ("summary", "string", False),
        6: ("description", "string", False),
        8: ("deprecation", "msg:Opaque", False),
        16: ("isAggregate", "bool", False),
        17: ("isStateful", "bool", False),
        18: ("isCommutative", "bool", False),
        19: ("allowsUninitializedInput", "bool", False),
    },
    "ArgDef": {
        1: ("name", "string", False),
        2: ("description", "string", False),
        3: ("type", "enum:DataType", False),
        4: ("typeAttr", "string", False),
        5: ("numberAttr", "string", False),
        6: ("typeListAttr", "string", False),
        16: ("isRef", "bool", False),
    },
    "OpAttrDef": {
        1: ("name", "string", False),
        2: ("type", "string", False),
        3: ("defaultValue", "msg:AttrValue", False),
        4: ("description", "string", False),
        5: ("hasMinimum", "bool", False),
        6: ("minimum", "int64", False),
        7: ("allowedValues", "msg:AttrValue", False),
    },
    "GraphDef": {
        1: ("node", "msg:NodeDef", True),
        2: ("library", "msg:Opaque", False),
        3: ("version", "int32", False),
        4: ("versions", "msg:VersionDef", False),
    },
    "VersionDef": {1: ("producer", "int32", False), 2: ("minConsumer", "int32", False), 3: ("badConsumers", "int32", True)},
    "NodeDef": {
        1: ("name", "string", False),
        2: ("op", "string", False),
        3: ("input", "string", True),
        4: ("device", "string", False),
        5: ("attr", "map:string,msg:AttrValue", False),
    },
    "AttrValue": {
        1: ("list", "msg:AttrList", False),
        2: ("s", "bytes", False),
        3: ("i", "int64", False),
        4: ("f", "float", False),
        5: ("b", "bool", False),
        6: ("type", "enum:DataType", False),
        7: ("shape", "msg:TensorShapeProto", False),
        8: ("tensor", "msg:TensorProto", False),
        9: ("placeholder", "string", False),
        10: ("func", "msg:Opaque", False),
    },
    "AttrList": {
        2: ("s", "bytes", True),
        3: ("i", "int64", True),
        4: ("f", "float", True),
        5: ("b", "bool", True),
        6: ("type", "enum:DataType", True),
        7: ("shape", "msg:TensorShapeProto", True),
        8: ("tensor", "msg:TensorProto", True),
        9: ("func", "msg:Opaque", True),
    },
    "TensorShapeProto": {2: ("dim", "msg:Dim", True), 3: ("unknownRank", "bool", False)},
    "Dim": {1: ("size", "int64", False), 2: ("name", "string", False)},
    "TensorProto": {
        1: ("dtype", "enum:DataType", False),
        2: ("tensorShape", "msg:TensorShapeProto", False),
        3: ("versionNumber", "int32", False),
        4: ("tensorContent", "bytes", False),
        5: ("floatVal", "float", True),
        6: ("doubleVal", "double", True),
        7: ("intVal", "int32", True),
        8: ("stringVal", "bytes", True),
        10: ("int64Val", "int64", True),
        11: ("boolVal", "bool", True),
        13: ("halfVal", "int32", True),
    },
    "SaverDef": {
        1: ("filenameTensorName", "string", False),
        2: ("saveTensorName", "string", False),
        3: ("restoreOpName", "string", False),
        4: ("maxToKeep", "int32", False),
        5: ("sharded", "bool", False),
        6: ("keepCheckpointEveryNHours", "float", False),
        7: ("version", "enum:SaverVersion", False),
    },
    "CollectionDef": {
        1: ("nodeList", "msg:NodeList", False),
        2: ("bytesList", "msg:BytesList", False),
        3: ("int64List", "msg:Int64List", False),
        4: ("floatList", "msg:FloatList", False),
        5: ("anyList", "msg:Opaque", False),
    },
    "NodeList": {1: ("value", "string", True)},
    "BytesList": {1: ("value", "bytes", True)},
    "Int64List": {1: ("value", "int64", True)},
    "FloatList": {1: ("value", "float", True)},
    "VariableDef": {
        1: ("variableName", "string", False),
        2: ("initializerName", "string", False),
        3: ("snapshotName", "string", False),
        4: ("saveSliceInfoDef", "msg:Opaque", False),
        5: ("isResource", "bool", False),
        6: ("initialValueName", "string", False),
        7: ("trainable", "bool", False),
    },
    # tensor_bundle.proto
    "BundleHeaderProto": {
        1: ("numShards", "int32", False),
        2: ("endianness", "enum:Endianness", False),
        3: ("version", "msg:VersionDef", False),
    },
    "BundleEntryProto": {
        1: ("dtype", "enum:DataType", False),
        2: ("shape", "msg:TensorShapeProto", False),
        3: ("shardId", "int32", False),
        4: ("offset", "int64", False),
        5: ("size", "int64", False),
        6: ("crc32c", "fixed32", False),
        7: ("slices", "msg:Opaque", True),
    },
    "CheckpointState": {1: ("modelCheckpointPath", "string", False), 2: ("allModelCheckpointPaths", "string", True)},
    "Opaque": {},
}

_PACKABLE = {"int32", "int64", "uint64", "bool", "float", "double", "fixed32"}


# ---------------------------------------------------------------------------
# decoding
# ---------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _scalar_from_varint(t: str, v: int) -> Any:
    if t == "bool":
        return bool(v)
    if t in ("int64",):
        return str(_signed64(v))
    if t == "uint64":
        return str(v)
    if t == "int32":
        v = _signed64(v)
        return int(v)
    if t.startswith("enum:"):
        return ENUMS[t[5:]].get(v, v)
    raise ValueError(t)


def _decode_scalar(t: str, wt: int, buf: bytes, pos: int) -> Tuple[Any, int]:
    if wt == 0:
        v, pos = _varint(buf, pos)
        return _scalar_from_varint(t, v), pos
    if wt == 5:
        raw = buf[pos:pos + 4]
        pos += 4
        if t == "float":
            return struct.unpack("<f", raw)[0], pos
        return struct.unpack("<I", raw)[0], pos
    if wt == 1:
        raw = buf[pos:pos + 8]
        pos += 8
        if t == "double":
            return struct.unpack("<d", raw)[0], pos
        return str(struct.unpack("<Q", raw)[0]), pos
    raise ValueError(f"wire type {wt} for scalar {t}")


def _skip(wt: int, buf: bytes, pos: int) -> int:
    if wt == 0:
        _, pos = _varint(buf, pos)
        return pos
    if wt == 1:
        return pos + 8
    if wt == 2:
        n, pos = _varint(buf, pos)
        return pos + n
    if wt == 5:
        return pos + 4
    raise ValueError(f"unsupported wire type {wt}")


def decode(msg: str, buf: bytes) -> Dict[str, Any]:
    """Binary protobuf -> proto3-JSON-shaped dict."""
    schema = S[msg]
    out: Dict[str, Any] = {}
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if fn not in schema:
            pos = _skip(wt, buf, pos)
            continue
        name, t, rep = schema[fn]
        if t.startswith("map:"):
            n, pos = _varint(buf, pos)
            k, v = _decode_map_entry(t, buf[pos:pos + n])
            pos += n
            out.setdefault(name, {})[k] = v
            continue
        if t.startswith("msg:"):
            n, pos = _varint(buf, pos)
            val: Any = decode(t[4:], buf[pos:pos + n])
            pos += n
        elif t == "string":
            n, pos = _varint(buf, pos)
            val = buf[pos:pos + n].decode("utf-8", "replace")
            pos += n
        elif t == "bytes":
            n, pos = _varint(buf, pos)
            val = base64.b64encode(buf[pos:pos + n]).decode("ascii")
            pos += n
        elif wt == 2 and (t in _PACKABLE or t.startswith("enum:")):
            n, pos = _varint(buf, pos)
            sub_end = pos + n
            vals: List[Any] = []
            sub_wt = 5 if t in ("float", "fixed32") else 1 if t == "double" else 0
            while pos < sub_end:
                v, pos = _decode_scalar(t, sub_wt, buf, pos)
                vals.append(v)
            out.setdefault(name, []).extend(vals)
            continue
        else:
            val, pos = _decode_scalar(t, wt, buf, pos)
        if rep:
            out.setdefault(name, []).append(val)
        else:
            out[name] = val
    return out


def _decode_map_entry(t: str, buf: bytes) -> Tuple[Any, Any]:
    kt, vt = t[4:].split(",", 1)
    key: Any = "" if kt == "string" else 0
    val: Any = {} if vt.startswith("msg:") else None
    pos = 0
    while pos < len(buf):
        k, pos = _varint(buf, pos)
        fn, wt = k >> 3, k & 7
        if fn == 1:
            n, pos = _varint(buf, pos)
            key = buf[pos:pos + n].decode("utf-8")
            pos += n
        elif fn == 2 and vt.startswith("msg:"):
            n, pos = _varint(buf, pos)
            val = decode(vt[4:], buf[pos:pos + n])
            pos += n
        else:
            pos = _skip(wt, buf, pos)
    return key, val


# ---------------------------------------------------------------------------
# encoding
# ---------------------------------------------------------------------------
def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_key(fn: int, wt: int) -> bytes:
    return _enc_varint((fn << 3) | wt)


def _enc_scalar(fn: int, t: str, v: Any) -> bytes:
    if t == "string":
        raw = v.encode("utf-8")
        return _enc_key(fn, 2) + _enc_varint(len(raw)) + raw
    if t == "bytes":
        raw = base64.b64decode(v) if isinstance(v, str) else bytes(v)
        return _enc_key(fn, 2) + _enc_varint(len(raw)) + raw
    if t == "float":
        return _enc_key(fn, 5) + struct.pack("<f", float(v))
    if t == "double":
        return _enc_key(fn, 1) + struct.pack("<d", float(v))
    if t == "fixed32":
        return _enc_key(fn, 5) + struct.pack("<I", int(v))
    if t == "bool":
        return _enc_key(fn, 0) + _enc_varint(1 if v else 0)
    if t.startswith("enum:"):
        iv = _ENUM_REV[t[5:]].get(v, v) if isinstance(v, str) else int(v)
        return _enc_key(fn, 0) + _enc_varint(int(iv))
    return _enc_key(fn, 0) + _enc_varint(int(v))


# messages whose scalar fields are members of a oneof: explicit presence, zeros ARE serialised
_ONEOF_MESSAGES = {"AttrValue"}


def _is_default(t: str, v: Any) -> bool:
    if t in ("string", "bytes"):
        return v == "" or v == b""
    if t == "bool":
        return not v
    if t.startswith("enum:"):
        return v == 0 or v == ENUMS[t[5:]].get(0)
    try:
        return float(v) == 0.0
    except (TypeError, ValueError):
        return False


def encode(msg: str, obj: Dict[str, Any]) -> bytes:
    """proto3-JSON-shaped dict -> binary protobuf (fields emitted in field-number order)."""
    schema = S[msg]
    by_name = {name: (fn, t, rep) for fn, (name, t, rep) in schema.items()}
    out = bytearray()
    for name, (fn, t, rep) in sorted(by_name.items(), key=lambda kv: kv[1][0]):
        if name not in obj or obj[name] is None:
            continue
        v = obj[name]
        if t.startswith("map:"):
            kt, vt = t[4:].split(",", 1)
            for k in v:
                entry = _enc_scalar(1, kt, k)
                sub = encode(vt[4:], v[k])
                entry += _enc_key(2, 2) + _enc_varint(len(sub)) + sub
                out += _enc_key(fn, 2) + _enc_varint(len(entry)) + entry
            continue
        if not rep and msg not in _ONEOF_MESSAGES and not t.startswith("msg:") and _is_default(t, v):
            continue                      # proto3: default-valued scalars are not serialised
        vals = v if rep else [v]
        for item in vals:
            if t.startswith("msg:"):
                sub = encode(t[4:], item)
                out += _enc_key(fn, 2) + _enc_varint(len(sub)) + sub
            else:
                out += _enc_scalar(fn, t, item)
    return bytes(out)
