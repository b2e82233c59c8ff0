"""PyTorch interpreter for GraphIR – the semantics oracle, the CPU/gloo engine and the generic path.

The reference evaluates graphs with ``tf.Session.run`` (HogwildSparkModel.py:47-53, ml_util.py:64-73).
Here a GraphIR is interpreted lazily from the requested fetches backwards (so gradient / optimizer /
saver sub-graphs of real TF-1.x MetaGraphs are never touched), with autograd providing
``tf.gradients``.  Unlike the reference, ONE backward pass yields every variable's gradient
(the reference re-runs forward+backward once per variable, HogwildSparkModel.py:66-67).

Conventions: floating tensors are ``torch.Tensor`` on the program's device; integer "meta" values
(shapes, axes, indices) are ``numpy`` arrays on the host.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from .ir import GraphIR, Node, split_ref

_TORCH_DT = {
    "DT_FLOAT": torch.float32, "DT_DOUBLE": torch.float64, "DT_HALF": torch.float16, "DT_INT32": torch.int32,
    "DT_INT64": torch.int64, "DT_BOOL": torch.bool, "DT_UINT8": torch.uint8, "DT_BFLOAT16": torch.bfloat16,
}
_FLOAT_DT = {"DT_FLOAT", "DT_DOUBLE", "DT_HALF", "DT_BFLOAT16"}


class UnsupportedOp(NotImplementedError):
    pass


def _is_np(x) -> bool:
    return isinstance(x, (np.ndarray, np.generic, int, float, bool))


def _to_np(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class _Ctx:
    def __init__(self, program: "GraphProgram", weights, feeds, training: bool, generator):
        self.p, self.weights, self.feeds, self.training, self.gen = program, weights, feeds, training, generator
        self.device = program.device

    def t(self, x, dtype=None) -> torch.Tensor:
        if isinstance(x, torch.Tensor):
            return x if dtype is None or x.dtype == dtype else x.to(dtype)
        arr = np.asarray(x)
        if dtype is None:
            dtype = torch.float32 if arr.dtype.kind == "f" else torch.int64 if arr.dtype.kind in "iu" else torch.bool
        return torch.as_tensor(arr, device=self.device).to(dtype)


OpFn = Callable[[Node, List[Any], _Ctx], Tuple[Any, ...]]
OPS: Dict[str, OpFn] = {}


def op(*names: str):
    def deco(fn: OpFn) -> OpFn:
        for n in names:
            OPS[n] = fn
        return fn
    return deco


# ---------------------------------------------------------------------------
# sources
# ---------------------------------------------------------------------------
@op("Const")
def _const(n, ins, c):
    v = n.attrs.get("value")
    if v is None:
        raise ValueError(f"Const node {n.name} has no value")
    if v.dtype.kind == "f":
        return (torch.as_tensor(v, device=c.device),)
    return (v,)


@op("Placeholder", "PlaceholderV2")
def _placeholder(n, ins, c):
    raise KeyError(f"You must feed a value for placeholder tensor '{n.name}'")


@op("PlaceholderWithDefault")
def _placeholder_default(n, ins, c):
    return (ins[0],)


@op("VariableV2", "Variable", "VarHandleOp")
def _variable(n, ins, c):
    if n.name in c.weights:
        return (c.weights[n.name],)
    # not one of the trained variables (moving statistics, step counters, frozen layers): the engine never runs
    # update ops on it - exactly like the reference, which only feeds gradients of trainable_variables - so it keeps
    # the value its initializer produces
    return (c.p.frozen_variable(n.name),)


@op("Identity", "ReadVariableOp", "StopGradient", "PreventGradient", "Snapshot", "CheckNumerics")
def _identity(n, ins, c):
    x = ins[0]
    if n.op in ("StopGradient", "PreventGradient") and isinstance(x, torch.Tensor):
        x = x.detach()
    return (x,)


@op("NoOp", "Assert")
def _noop(n, ins, c):
    return (None,)


# ---------------------------------------------------------------------------
# shape / structure
# ---------------------------------------------------------------------------
@op("Shape")
def _shape(n, ins, c):
    return (np.asarray(tuple(ins[0].shape), dtype=np.int64),)


@op("Size")
def _size(n, ins, c):
    return (np.asarray(int(np.prod(tuple(ins[0].shape))), dtype=np.int64),)


@op("Rank")
def _rank(n, ins, c):
    return (np.asarray(len(ins[0].shape), dtype=np.int64),)


@op("Reshape")
def _reshape(n, ins, c):
    shape = [int(s) for s in _to_np(ins[1]).reshape(-1)]
    x = ins[0]
    return (x.reshape(shape) if isinstance(x, torch.Tensor) else np.reshape(x, shape),)


@op("Squeeze")
def _squeeze(n, ins, c):
    dims = n.attrs.get("squeeze_dims") or []
    x = ins[0]
    if not dims:
        return (x.squeeze(),)
    for d in sorted((d % x.dim() for d in dims), reverse=True):
        x = x.squeeze(d)
    return (x,)


@op("ExpandDims")
def _expand(n, ins, c):
    d = int(_to_np(ins[1]))
    x = ins[0]
    return (x.unsqueeze(d if d >= 0 else d + x.dim() + 1) if isinstance(x, torch.Tensor) else np.expand_dims(x, d),)


@op("Transpose")
def _transpose(n, ins, c):
    return (ins[0].permute([int(p) for p in _to_np(ins[1])]),)


@op("Pack")
def _pack(n, ins, c):
    ax = n.attrs.get("axis", 0)
    if all(_is_np(i) for i in ins):
        return (np.stack([np.asarray(i) for i in ins], axis=ax),)
    return (torch.stack([c.t(i) for i in ins], dim=ax),)


@op("ConcatV2")
def _concat(n, ins, c):
    ax = int(_to_np(ins[-1]))
    vals = ins[:-1]
    if all(_is_np(v) for v in vals):
        return (np.concatenate([np.atleast_1d(v) for v in vals], axis=ax),)
    return (torch.cat([c.t(v) for v in vals], dim=ax),)


@op("StridedSlice")
def _strided_slice(n, ins, c):
    x = ins[0]
    begin, end, strides = (_to_np(v).reshape(-1).astype(np.int64) for v in ins[1:4])
    bm, em, sm = n.attrs.get("begin_mask", 0), n.attrs.get("end_mask", 0), n.attrs.get("shrink_axis_mask", 0)
    if n.attrs.get("ellipsis_mask", 0) or n.attrs.get("new_axis_mask", 0):
        raise UnsupportedOp("StridedSlice with ellipsis/new_axis masks")
    idx: List[Any] = []
    for i in range(len(begin)):
        if sm & (1 << i):
            idx.append(int(begin[i]))
        else:
            b = None if bm & (1 << i) else int(begin[i])
            e = None if em & (1 << i) else int(end[i])
            idx.append(slice(b, e, int(strides[i])))
    return (x[tuple(idx)],)


@op("Slice")
def _slice(n, ins, c):
    x = ins[0]
    begin, size = _to_np(ins[1]).reshape(-1), _to_np(ins[2]).reshape(-1)
    idx = tuple(slice(int(b), None if s < 0 else int(b + s)) for b, s in zip(begin, size))
    return (x[idx],)


@op("Fill")
def _fill(n, ins, c):
    dims = [int(d) for d in _to_np(ins[0]).reshape(-1)]
    v = ins[1]
    if _is_np(v) and np.asarray(v).dtype.kind != "f":
        return (np.full(dims, np.asarray(v).item()),)
    return (torch.full(dims, float(_to_np(v)), device=c.device),)


@op("ZerosLike")
def _zeros_like(n, ins, c):
    return (torch.zeros_like(ins[0]) if isinstance(ins[0], torch.Tensor) else np.zeros_like(ins[0]),)


@op("OnesLike")
def _ones_like(n, ins, c):
    return (torch.ones_like(ins[0]) if isinstance(ins[0], torch.Tensor) else np.ones_like(ins[0]),)


@op("Tile")
def _tile(n, ins, c):
    return (ins[0].repeat([int(m) for m in _to_np(ins[1]).reshape(-1)]),)


@op("Cast", "ToFloat")
def _cast(n, ins, c):
    dst = n.attrs.get("DstT", "DT_FLOAT")
    x = ins[0]
    if dst in _FLOAT_DT:
        return (c.t(x, _TORCH_DT[dst]),)
    if isinstance(x, torch.Tensor):
        return (x.to(_TORCH_DT[dst]),)
    return (np.asarray(x).astype({"DT_INT32": np.int64, "DT_INT64": np.int64, "DT_BOOL": np.bool_}.get(dst, np.int64)),)


@op("Range")
def _range(n, ins, c):
    return (np.arange(int(_to_np(ins[0])), int(_to_np(ins[1])), int(_to_np(ins[2]))),)


# ---------------------------------------------------------------------------
# elementwise math
# ---------------------------------------------------------------------------
def _binary(fn_t, fn_np):
    def run(n, ins, c):
        a, b = ins
        if _is_np(a) and _is_np(b) and np.asarray(a).dtype.kind != "f" and np.asarray(b).dtype.kind != "f":
            return (fn_np(np.asarray(a), np.asarray(b)),)
        ta = a if isinstance(a, torch.Tensor) else None
        tb = b if isinstance(b, torch.Tensor) else None
        ref = ta if ta is not None else tb
        dt = ref.dtype if ref is not None and ref.dtype.is_floating_point else torch.float32
        return (fn_t(c.t(a, dt) if ta is None else ta, c.t(b, dt) if tb is None else tb),)
    return run


OPS["Add"] = OPS["AddV2"] = _binary(torch.add, np.add)
OPS["Sub"] = _binary(torch.sub, np.subtract)
OPS["Mul"] = _binary(torch.mul, np.multiply)
OPS["RealDiv"] = OPS["Div"] = _binary(torch.div, lambda a, b: a / b)
OPS["FloorDiv"] = _binary(lambda a, b: torch.floor(a / b), np.floor_divide)
OPS["Maximum"] = _binary(torch.maximum, np.maximum)
OPS["Minimum"] = _binary(torch.minimum, np.minimum)
OPS["Pow"] = _binary(torch.pow, np.power)
OPS["SquaredDifference"] = _binary(lambda a, b: (a - b) ** 2, lambda a, b: (a - b) ** 2)
OPS["Greater"] = _binary(torch.gt, np.greater)
OPS["GreaterEqual"] = _binary(torch.ge, np.greater_equal)
OPS["Less"] = _binary(torch.lt, np.less)
OPS["LessEqual"] = _binary(torch.le, np.less_equal)
OPS["Equal"] = _binary(torch.eq, np.equal)
OPS["NotEqual"] = _binary(torch.ne, np.not_equal)
OPS["DivNoNan"] = _binary(lambda a, b: torch.where(b == 0, torch.zeros_like(a), a / b), lambda a, b: np.where(b == 0, 0, a / np.where(b == 0, 1, b)))


@op("AddN")
def _addn(n, ins, c):
    out = ins[0]
    for x in ins[1:]:
        out = out + x
    return (out,)


def _unary(fn_t, fn_np=None):
    def run(n, ins, c):
        x = ins[0]
        if fn_np is not None and _is_np(x) and np.asarray(x).dtype.kind != "f":
            return (fn_np(np.asarray(x)),)
        return (fn_t(c.t(x)),)
    return run


OPS["Neg"] = _unary(torch.neg, np.negative)
OPS["Abs"] = _unary(torch.abs, np.abs)
OPS["Square"] = _unary(torch.square, np.square)
OPS["Sqrt"] = _unary(torch.sqrt)
OPS["Rsqrt"] = _unary(torch.rsqrt)
OPS["Exp"] = _unary(torch.exp)
OPS["Log"] = _unary(torch.log)
OPS["Log1p"] = _unary(torch.log1p)
OPS["Floor"] = _unary(torch.floor)
OPS["Ceil"] = _unary(torch.ceil)
OPS["Sign"] = _unary(torch.sign, np.sign)
OPS["Reciprocal"] = OPS["Inv"] = _unary(torch.reciprocal)
OPS["Relu"] = _unary(torch.relu)
OPS["Relu6"] = _unary(lambda x: torch.clamp(x, 0.0, 6.0))
OPS["Sigmoid"] = _unary(torch.sigmoid)
OPS["Tanh"] = _unary(torch.tanh)
OPS["Softplus"] = _unary(F.softplus)
OPS["Softsign"] = _unary(F.softsign)
OPS["Elu"] = _unary(F.elu)
OPS["Selu"] = _unary(F.selu)
OPS["Softmax"] = _unary(lambda x: torch.softmax(x, dim=-1))
OPS["LogSoftmax"] = _unary(lambda x: torch.log_softmax(x, dim=-1))
OPS["LogicalNot"] = _unary(torch.logical_not, np.logical_not)


@op("LeakyRelu")
def _leaky(n, ins, c):
    return (F.leaky_relu(ins[0], n.attrs.get("alpha", 0.2)),)


@op("Select", "SelectV2")
def _select(n, ins, c):
    cond, a, b = ins
    cond = c.t(cond, torch.bool)
    ta, tb = c.t(a), c.t(b)
    if n.op == "Select" and cond.dim() == 1 and ta.dim() > 1:
        cond = cond.reshape([-1] + [1] * (ta.dim() - 1))
    return (torch.where(cond, ta, tb),)


@op("LogicalAnd")
def _land(n, ins, c):
    return (torch.logical_and(c.t(ins[0], torch.bool), c.t(ins[1], torch.bool)),)


# ---------------------------------------------------------------------------
# reductions
# ---------------------------------------------------------------------------
def _reduction(fn):
    def run(n, ins, c):
        x = ins[0]
        axes = [int(a) for a in _to_np(ins[1]).reshape(-1)]
        keep = bool(n.attrs.get("keep_dims", n.attrs.get("keepdims", False)))
        if _is_np(x):
            x = c.t(x)
        if x.dim() == 0 or not axes:
            return (x,)
        return (fn(x, dim=axes, keepdim=keep),)
    return run


OPS["Sum"] = _reduction(torch.sum)
OPS["Mean"] = _reduction(torch.mean)
OPS["Max"] = _reduction(torch.amax)
OPS["Min"] = _reduction(torch.amin)


@op("Prod")
def _prod(n, ins, c):
    x = ins[0]
    axes = [int(a) for a in _to_np(ins[1]).reshape(-1)]
    if _is_np(x):
        return (np.prod(np.asarray(x), axis=tuple(axes) if axes else None, keepdims=bool(n.attrs.get("keep_dims", False))),)
    for a in sorted(axes, reverse=True):
        x = x.prod(dim=a, keepdim=bool(n.attrs.get("keep_dims", False)))
    return (x,)


@op("ArgMax")
def _argmax(n, ins, c):
    return (torch.argmax(ins[0], dim=int(_to_np(ins[1]))),)


@op("ArgMin")
def _argmin(n, ins, c):
    return (torch.argmin(ins[0], dim=int(_to_np(ins[1]))),)


# ---------------------------------------------------------------------------
# neural-network ops
# ---------------------------------------------------------------------------
@op("MatMul")
def _matmul(n, ins, c):
    a, b = ins
    if n.attrs.get("transpose_a"):
        a = a.t()
    if n.attrs.get("transpose_b"):
        b = b.t()
    return (a @ b,)


@op("BiasAdd")
def _bias_add(n, ins, c):
    x, b = ins
    if n.attrs.get("data_format", "NHWC") == "NCHW" and x.dim() == 4:
        return (x + b.reshape(1, -1, 1, 1),)
    return (x + b,)


def _same_pad(size: int, k: int, s: int) -> Tuple[int, int]:
    out = (size + s - 1) // s
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


@op("Conv2D")
def _conv2d(n, ins, c):
    x, w = ins                                     # NHWC, HWIO
    if n.attrs.get("data_format", "NHWC") != "NHWC":
        raise UnsupportedOp("Conv2D data_format NCHW")
    st = n.attrs.get("strides", [1, 1, 1, 1])
    dil = n.attrs.get("dilations", [1, 1, 1, 1])
    xt = x.permute(0, 3, 1, 2)
    wt = w.permute(3, 2, 0, 1)
    if n.attrs.get("padding", "VALID") == "SAME":
        ph = _same_pad(xt.shape[2], (w.shape[0] - 1) * dil[1] + 1, st[1])
        pw = _same_pad(xt.shape[3], (w.shape[1] - 1) * dil[2] + 1, st[2])
        xt = F.pad(xt, (pw[0], pw[1], ph[0], ph[1]))
    y = F.conv2d(xt, wt, stride=(st[1], st[2]), dilation=(dil[1], dil[2]))
    return (y.permute(0, 2, 3, 1),)


def _pool(kind):
    def run(n, ins, c):
        x = ins[0]
        ks, st = n.attrs.get("ksize", [1, 2, 2, 1]), n.attrs.get("strides", [1, 2, 2, 1])
        xt = x.permute(0, 3, 1, 2)
        if n.attrs.get("padding", "VALID") == "SAME":
            ph, pw = _same_pad(xt.shape[2], ks[1], st[1]), _same_pad(xt.shape[3], ks[2], st[2])
            xt = F.pad(xt, (pw[0], pw[1], ph[0], ph[1]), value=float("-inf") if kind == "max" else 0.0)
        if kind == "max":
            y = F.max_pool2d(xt, (ks[1], ks[2]), (st[1], st[2]))
        else:
            y = F.avg_pool2d(xt, (ks[1], ks[2]), (st[1], st[2]))
            if n.attrs.get("padding", "VALID") == "SAME":
                # TF excludes the padded cells from the divisor: rescale by window / (number of real cells in the window)
                ones = F.pad(torch.ones_like(x.permute(0, 3, 1, 2)[:1, :1]), (pw[0], pw[1], ph[0], ph[1]))
                frac = F.avg_pool2d(ones, (ks[1], ks[2]), (st[1], st[2]))
                y = y / frac
        return (y.permute(0, 2, 3, 1),)
    return run


OPS["MaxPool"] = _pool("max")
OPS["AvgPool"] = _pool("avg")


@op("SoftmaxCrossEntropyWithLogits")
def _softmax_xent(n, ins, c):
    logits, labels = ins
    logp = torch.log_softmax(logits, dim=-1)
    loss = -(labels * logp).sum(dim=-1)
    return (loss, torch.softmax(logits, dim=-1) - labels)


@op("SparseSoftmaxCrossEntropyWithLogits")
def _sparse_softmax_xent(n, ins, c):
    logits, labels = ins
    lab = c.t(labels, torch.int64)
    loss = F.cross_entropy(logits, lab, reduction="none")
    return (loss, torch.softmax(logits, dim=-1) - F.one_hot(lab, logits.shape[-1]).to(logits.dtype))


@op("L2Loss")
def _l2loss(n, ins, c):
    return (0.5 * (ins[0] ** 2).sum(),)


@op("OneHot")
def _one_hot(n, ins, c):
    idx, depth, on, off = ins
    oh = F.one_hot(c.t(idx, torch.int64), int(_to_np(depth))).to(torch.float32)
    return (oh * float(_to_np(on)) + (1.0 - oh) * float(_to_np(off)),)


# ---------------------------------------------------------------------------
# ops beyond what the reference's own examples need: real TF-1.x graphs (embeddings, tf.cond around dropout / batch-norm,
# fused batch-norm, transposed / depthwise convolutions, ...) train on the interpreter engine through these
# ---------------------------------------------------------------------------
class _Dead:
    """Value on the untaken side of a tf.cond (`Switch`): every op that consumes it produces it again; `Merge` forwards
    whichever input is alive."""

    def __repr__(self):
        return "<dead tensor>"


DEAD = _Dead()


class _DeadOutputs:
    def __getitem__(self, i):
        return DEAD


@op("Switch", "RefSwitch")
def _switch(n, ins, c):
    data, pred = ins
    take_true = bool(_to_np(pred).reshape(-1)[0])
    return (DEAD, data) if take_true else (data, DEAD)


@op("Merge", "RefMerge")
def _merge(n, ins, c):
    for i, v in enumerate(ins):
        if v is not DEAD:
            return (v, np.asarray(i, dtype=np.int64))
    return (DEAD, DEAD)


@op("GatherV2", "Gather", "ResourceGather")
def _gather(n, ins, c):
    params, idx = ins[0], ins[1]
    axis = int(_to_np(ins[2])) if len(ins) > 2 else 0
    if _is_np(params) and _is_np(idx):
        return (np.take(np.asarray(params), np.asarray(idx), axis=axis),)
    p = c.t(params)
    i = c.t(idx, torch.int64)
    out = torch.index_select(p, axis % p.dim(), i.reshape(-1))
    shape = list(p.shape[:axis % p.dim()]) + list(i.shape) + list(p.shape[axis % p.dim() + 1:])
    return (out.reshape(shape),)


@op("Split")
def _split(n, ins, c):
    axis, x = int(_to_np(ins[0])), ins[1]
    return tuple(torch.chunk(c.t(x), int(n.attrs.get("num_split", 1)), dim=axis))


@op("SplitV")
def _splitv(n, ins, c):
    x, sizes, axis = c.t(ins[0]), [int(v) for v in _to_np(ins[1]).reshape(-1)], int(_to_np(ins[2]))
    if -1 in sizes:
        sizes[sizes.index(-1)] = x.shape[axis] - (sum(sizes) + 1)
    return tuple(torch.split(x, sizes, dim=axis))


@op("Unpack")
def _unpack(n, ins, c):
    return tuple(torch.unbind(c.t(ins[0]), dim=int(n.attrs.get("axis", 0))))


@op("Pad", "PadV2")
def _pad(n, ins, c):
    x = c.t(ins[0])
    pads = _to_np(ins[1]).reshape(-1, 2).astype(np.int64)
    flat: List[int] = []
    for lo, hi in pads[::-1]:                      # torch pads the LAST dimension first
        flat += [int(lo), int(hi)]
    value = float(_to_np(ins[2])) if n.op == "PadV2" and len(ins) > 2 else 0.0
    return (F.pad(x, flat, value=value),)


@op("MirrorPad")
def _mirror_pad(n, ins, c):
    x = c.t(ins[0])
    pads = _to_np(ins[1]).reshape(-1, 2).astype(np.int64)
    mode = n.attrs.get("mode", "REFLECT")
    mode = (mode.decode() if isinstance(mode, bytes) else str(mode)).upper()
    for d, (lo, hi) in enumerate(pads):
        if lo or hi:                               # one gather per padded dimension; numpy supplies the index pattern
            idx = np.pad(np.arange(x.shape[d]), (int(lo), int(hi)), mode="reflect" if mode == "REFLECT" else "symmetric")
            x = torch.index_select(x, d, torch.as_tensor(idx, device=x.device))
    return (x,)


@op("BatchMatMul", "BatchMatMulV2")
def _batch_matmul(n, ins, c):
    a, b = c.t(ins[0]), c.t(ins[1])
    if n.attrs.get("adj_x"):
        a = a.transpose(-1, -2)
    if n.attrs.get("adj_y"):
        b = b.transpose(-1, -2)
    return (a @ b,)


@op("FusedBatchNorm", "FusedBatchNormV2", "FusedBatchNormV3")
def _fused_batch_norm(n, ins, c):
    x, scale, offset, mean, var = (c.t(v) for v in ins[:5])
    if n.attrs.get("data_format", "NHWC") != "NHWC":
        raise UnsupportedOp("FusedBatchNorm data_format NCHW")
    eps = float(n.attrs.get("epsilon", 1e-3))
    if n.attrs.get("is_training", True):
        dims = list(range(x.dim() - 1))
        bm = x.mean(dim=dims)
        bv = x.var(dim=dims, unbiased=False)
        cnt = x.numel() // x.shape[-1]
        y = (x - bm) * torch.rsqrt(bv + eps) * scale + offset
        return (y, bm, bv * (cnt / max(cnt - 1, 1)), bm, bv, bv)
    y = (x - mean) * torch.rsqrt(var + eps) * scale + offset
    return (y, mean, var, mean, var, var)


OPS["Erf"] = _unary(torch.erf)
OPS["Sin"] = _unary(torch.sin)
OPS["Cos"] = _unary(torch.cos)
OPS["Round"] = OPS["Rint"] = _unary(torch.round, np.round)
OPS["IsNan"] = _unary(torch.isnan)
OPS["FloorMod"] = _binary(torch.remainder, np.mod)
OPS["LogicalOr"] = lambda n, ins, c: (torch.logical_or(c.t(ins[0], torch.bool), c.t(ins[1], torch.bool)),)


@op("ClipByValue")
def _clip(n, ins, c):
    x, lo, hi = (c.t(v) for v in ins)
    return (torch.minimum(torch.maximum(x, lo), hi),)


@op("Cumsum")
def _cumsum(n, ins, c):
    x, axis = c.t(ins[0]), int(_to_np(ins[1]))
    if n.attrs.get("reverse"):
        x = x.flip(axis)
    y = torch.cumsum(x, dim=axis)
    if n.attrs.get("exclusive"):
        y = y - x
    return (y.flip(axis) if n.attrs.get("reverse") else y,)


def _bool_reduction(fn):
    def run(n, ins, c):
        x = c.t(ins[0], torch.bool)
        axes = [int(a) for a in _to_np(ins[1]).reshape(-1)]
        keep = bool(n.attrs.get("keep_dims", False))
        for a in sorted((a % max(x.dim(), 1) for a in axes), reverse=True) if x.dim() else []:
            x = fn(x, dim=a, keepdim=keep)
        return (x,)
    return run


OPS["Any"] = _bool_reduction(torch.any)
OPS["All"] = _bool_reduction(torch.all)


@op("TopKV2", "TopK")
def _topk(n, ins, c):
    k = int(_to_np(ins[1])) if len(ins) > 1 else int(n.attrs.get("k", 1))
    v, i = torch.topk(c.t(ins[0]), k, dim=-1, sorted=True)
    return (v, i)


@op("ReverseV2")
def _reverse(n, ins, c):
    return (c.t(ins[0]).flip([int(a) for a in _to_np(ins[1]).reshape(-1)]),)


@op("Conv2DBackpropInput")
def _conv2d_transpose(n, ins, c):
    """Forward use = tf.nn.conv2d_transpose: inputs (output_shape, filter HWOI-as-HWIO-of-the-forward-conv, value)."""
    out_shape, w, y = [int(v) for v in _to_np(ins[0]).reshape(-1)], c.t(ins[1]), c.t(ins[2])
    if n.attrs.get("data_format", "NHWC") != "NHWC":
        raise UnsupportedOp("Conv2DBackpropInput data_format NCHW")
    st = n.attrs.get("strides", [1, 1, 1, 1])
    kh, kw = w.shape[0], w.shape[1]
    pad_h = pad_w = (0, 0)
    if n.attrs.get("padding", "VALID") == "SAME":
        pad_h, pad_w = _same_pad(out_shape[1], kh, st[1]), _same_pad(out_shape[2], kw, st[2])
    yt = y.permute(0, 3, 1, 2)
    wt = w.permute(3, 2, 0, 1)                     # [C_of_value, C_out, kh, kw] is what conv_transpose2d expects
    full = F.conv_transpose2d(yt, wt, stride=(st[1], st[2]))
    # crop the padding of the forward convolution / extend to the requested output size
    need_h, need_w = out_shape[1] + pad_h[0] + pad_h[1], out_shape[2] + pad_w[0] + pad_w[1]
    if full.shape[2] < need_h or full.shape[3] < need_w:
        full = F.pad(full, (0, max(need_w - full.shape[3], 0), 0, max(need_h - full.shape[2], 0)))
    full = full[:, :, pad_h[0]:pad_h[0] + out_shape[1], pad_w[0]:pad_w[0] + out_shape[2]]
    return (full.permute(0, 2, 3, 1),)


@op("DepthwiseConv2dNative")
def _depthwise(n, ins, c):
    x, w = c.t(ins[0]), c.t(ins[1])                 # NHWC, [kh, kw, C, multiplier]
    if n.attrs.get("data_format", "NHWC") != "NHWC":
        raise UnsupportedOp("DepthwiseConv2dNative data_format NCHW")
    st = n.attrs.get("strides", [1, 1, 1, 1])
    dil = n.attrs.get("dilations", [1, 1, 1, 1])
    kh, kw, ch, mult = w.shape
    xt = x.permute(0, 3, 1, 2)
    wt = w.permute(2, 3, 0, 1).reshape(ch * mult, 1, kh, kw)
    if n.attrs.get("padding", "VALID") == "SAME":
        ph = _same_pad(xt.shape[2], (kh - 1) * dil[1] + 1, st[1])
        pw = _same_pad(xt.shape[3], (kw - 1) * dil[2] + 1, st[2])
        xt = F.pad(xt, (pw[0], pw[1], ph[0], ph[1]))
    y = F.conv2d(xt, wt, stride=(st[1], st[2]), dilation=(dil[1], dil[2]), groups=ch)
    return (y.permute(0, 2, 3, 1),)


@op("ResizeNearestNeighbor")
def _resize_nearest(n, ins, c):
    x = c.t(ins[0])
    oh, ow = (int(v) for v in _to_np(ins[1]).reshape(-1))
    if n.attrs.get("align_corners") or n.attrs.get("half_pixel_centers"):
        raise UnsupportedOp("ResizeNearestNeighbor with align_corners / half_pixel_centers")
    ih, iw = x.shape[1], x.shape[2]
    ys = torch.clamp((torch.arange(oh, device=x.device) * (ih / oh)).floor().long(), max=ih - 1)
    xs = torch.clamp((torch.arange(ow, device=x.device) * (iw / ow)).floor().long(), max=iw - 1)
    return (x[:, ys][:, :, xs],)


# ---------------------------------------------------------------------------
# random ops (initializers, dropout)
# ---------------------------------------------------------------------------
def _rand_shape(ins):
    return [int(d) for d in _to_np(ins[0]).reshape(-1)]


@op("RandomUniform")
def _random_uniform(n, ins, c):
    return (torch.rand(_rand_shape(ins), device=c.device, generator=c.gen),)


@op("RandomStandardNormal")
def _random_normal(n, ins, c):
    return (torch.randn(_rand_shape(ins), device=c.device, generator=c.gen),)


@op("TruncatedNormal")
def _truncated_normal(n, ins, c):
    out = torch.empty(_rand_shape(ins), device=c.device)
    torch.nn.init.trunc_normal_(out, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=c.gen)
    return (out,)


# ---------------------------------------------------------------------------
# program
# ---------------------------------------------------------------------------
class GraphProgram:
    """Executable view of a GraphIR."""

    def __init__(self, ir: GraphIR, device: Union[str, torch.device] = "cpu"):
        self.ir = ir
        self.device = torch.device(device)
        self._gen: Optional[torch.Generator] = None

    # -- variables ---------------------------------------------------------------------------
    @property
    def var_names(self) -> List[str]:
        return [v.name for v in self.ir.trainable]

    def seed(self, seed: Optional[int]) -> None:
        if seed is None:
            self._gen = None
        else:
            self._gen = torch.Generator(device=self.device)
            self._gen.manual_seed(int(seed))

    def frozen_variable(self, name: str) -> torch.Tensor:
        """Value of a variable that is not in ``trainable_variables``: its initializer, evaluated once and cached."""
        cache = self.__dict__.setdefault("_frozen", {})
        if name not in cache:
            assign = self.ir.nodes.get(f"{name}/Assign")
            if assign is None or len(assign.inputs) < 2:
                raise KeyError(f"no value bound for variable '{name}' and it has no initializer")
            src, idx = assign.inputs[1]
            with torch.no_grad():
                val = self.run([f"{src}:{idx}"], {}, {})[0]
            cache[name] = (val if isinstance(val, torch.Tensor) else torch.as_tensor(np.asarray(val), device=self.device)).detach()
        return cache[name]

    def init_weights(self, seed: Optional[int] = None) -> List[np.ndarray]:
        """Evaluate every trainable variable's initializer sub-graph (``global_variables_initializer``)."""
        if seed is not None:
            self.seed(seed)
        out = []
        with torch.no_grad():
            for v in self.ir.trainable:
                if v.initial_value and self.ir.has_tensor(v.initial_value):
                    val = self.run([v.initial_value], {}, {})[0]
                    out.append(_to_np(val).astype(np.float32).reshape(v.shape))
                else:
                    out.append(np.zeros(v.shape, dtype=np.float32))
        return out

    def bind(self, weights: Sequence[Any], requires_grad: bool = False) -> Dict[str, torch.Tensor]:
        names = self.var_names
        if len(weights) != len(names):
            raise ValueError(f"expected {len(names)} weight arrays, got {len(weights)}")
        bound = {}
        for name, w, info in zip(names, weights, self.ir.trainable):
            t = w if isinstance(w, torch.Tensor) else torch.as_tensor(np.asarray(w, dtype=np.float32))
            t = t.to(self.device, torch.float32).reshape(info.shape)
            if requires_grad:
                t = t.detach().clone().requires_grad_(True)
            bound[name] = t
        return bound

    # -- evaluation ----------------------------------------------------------------------------
    def run(self, fetches: Sequence[str], feeds: Dict[str, Any], weights: Dict[str, torch.Tensor], training: bool = False) -> List[Any]:
        ctx = _Ctx(self, weights, feeds, training, self._gen)
        fed: Dict[Tuple[str, int], Any] = {}
        for k, v in feeds.items():
            key = split_ref(k)
            if isinstance(v, torch.Tensor):
                fed[key] = v.to(self.device)
            else:
                arr = np.asarray(v)
                fed[key] = torch.as_tensor(arr.astype(np.float32) if arr.dtype.kind in "fiub" else arr, device=self.device)
        cache: Dict[str, Tuple[Any, ...]] = {}
        targets = [split_ref(f) for f in fetches]
        for name, _ in targets:
            if name not in self.ir.nodes:
                raise KeyError(f"The name '{name}' refers to a Tensor which does not exist in the graph.")
        stack: List[Tuple[str, bool]] = [(name, False) for name, idx in targets if (name, idx) not in fed]
        while stack:
            name, expanded = stack.pop()
            if name in cache:
                continue
            node = self.ir.nodes[name]
            if not expanded:
                stack.append((name, True))
                for src, idx in node.inputs:
                    if (src, idx) not in fed and src not in cache:
                        if src not in self.ir.nodes:
                            raise KeyError(f"node '{name}' refers to missing input '{src}'")
                        stack.append((src, False))
                continue
            ins = [fed[(s, i)] if (s, i) in fed else cache[s][i] for s, i in node.inputs]
            if node.op not in ("Merge", "RefMerge") and any(v is DEAD for v in ins):
                cache[name] = _DeadOutputs()          # untaken branch of a tf.cond
                continue
            fn = OPS.get(node.op)
            if fn is None:
                raise UnsupportedOp(f"graph op '{node.op}' (node '{node.name}') is not supported by sparkflow_b200; "
                                    f"supported ops: {sorted(OPS)}")
            cache[name] = fn(node, ins, ctx)
        return [fed[(n, i)] if (n, i) in fed else cache[n][i] for n, i in targets]

    def forward(self, output: str, feeds: Dict[str, Any], weights: Sequence[Any]) -> torch.Tensor:
        with torch.no_grad():
            out = self.run([output], feeds, self.bind(weights))[0]
        return out if isinstance(out, torch.Tensor) else torch.as_tensor(np.asarray(out))

    def loss_and_grads(self, feeds: Dict[str, Any], weights: Sequence[Any], loss_name: Optional[str] = None) -> Tuple[float, List[torch.Tensor]]:
        """loss value and d loss / d variable for every trainable variable, from ONE backward pass."""
        loss_name = loss_name or (self.ir.losses[0] if self.ir.losses else None)
        if loss_name is None:
            raise ValueError("the graph has no entry in the 'losses' collection (use tf.losses.* or tf.losses.add_loss)")
        bound = self.bind(weights, requires_grad=True)
        with torch.enable_grad():
            loss = self.run([loss_name], feeds, bound, training=True)[0]
            if loss.dim() > 0:
                loss = loss.sum()
            params = [bound[n] for n in self.var_names]
            grads = torch.autograd.grad(loss, params, allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, params)]
        return float(loss.detach()), [g.detach() for g in grads]

    def loss(self, feeds: Dict[str, Any], weights: Sequence[Any], loss_name: Optional[str] = None) -> float:
        loss_name = loss_name or self.ir.losses[0]
        with torch.no_grad():
            v = self.run([loss_name], feeds, self.bind(weights))[0]
        return float(v.sum())
