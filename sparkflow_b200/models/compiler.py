"""GraphIR -> LayerPlan: recognise the layer structure the B200 engine has hand-written kernels for.

The reference runs whatever graph the user serialised through a TF session.  The B200 engine
instead executes a *compiled* plan: the forward chain between ``tfInput`` and the loss is pattern-
matched into dense / conv / pool / reshape layers with fused bias+activation, the loss into
softmax-cross-entropy or mean-squared-error, ``tf.nn.dropout`` / ``tf.layers.dropout`` after a dense layer into a
fused Philox mask in that layer's epilogue, and ``tfOutput`` into "activation of layer k" (+ optional ArgMax /
Softmax).  Graphs outside this family (custom losses, exotic ops) are reported as unsupported so the caller can use the
generic interpreter engine instead (loudly: a RuntimeWarning names the reason).

Covers every model the reference ships: simple_dnn, cnn_example,
autoencoder_example, the test MLPs / auto-encoder and the checkpoint fixture.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from ..graph.ir import GraphIR, Node, split_ref

ACT_IDS = {None: 0, "Relu": 1, "Sigmoid": 2, "Tanh": 3}


class UnsupportedGraph(ValueError):
    """The graph is valid but outside the family the compiled B200 plan covers."""


@dataclass
class Layer:
    kind: str                                   # dense | conv | pool | reshape
    kernel: Optional[str] = None                # variable node names
    bias: Optional[str] = None
    act: Optional[str] = None                   # Relu | Sigmoid | Tanh | None
    in_shape: Tuple[int, ...] = ()              # per-sample shape (no batch dim)
    out_shape: Tuple[int, ...] = ()
    ksize: Tuple[int, int] = (0, 0)
    tensors: Dict[str, str] = field(default_factory=dict)   # stage -> tensor name (linear / bias / act / dropout)
    dropout_keep: float = 0.0                   # dense: keep probability of a tf.nn.dropout applied to the layer's output (0 = none)

    @property
    def in_features(self) -> int:
        n = 1
        for d in self.in_shape:
            n *= d
        return n

    @property
    def out_features(self) -> int:
        n = 1
        for d in self.out_shape:
            n *= d
        return n


@dataclass
class OutputSpec:
    layer: int                                  # index into LayerPlan.layers whose output is fetched (-1 = the input)
    stage: str                                  # linear | bias | act
    post: Optional[str] = None                  # ArgMax | Softmax | None


@dataclass
class LayerPlan:
    layers: List[Layer]
    loss: Optional[str]                         # softmax_xent | mse | None (inference only)
    input_name: str
    label_name: Optional[str]                   # placeholder fed with labels; None if target is the input
    target_is_input: bool
    input_dim: int
    label_dim: int
    output: Optional[OutputSpec]
    var_order: List[str]

    def trainable_layers(self) -> List[Layer]:
        return [l for l in self.layers if l.kind in ("dense", "conv")]

    def is_mlp(self) -> bool:
        return all(l.kind in ("dense", "reshape") for l in self.layers)


_PASS = ("Identity", "StopGradient", "PlaceholderWithDefault")


def _var_of(ir: GraphIR, ref: Tuple[str, int]) -> Optional[str]:
    node = ir.nodes[ref[0]]
    hops = 0
    while node.op in ("Identity", "ReadVariableOp") and node.inputs and hops < 4:
        node = ir.nodes[node.inputs[0][0]]
        hops += 1
    return node.name if node.op in ("VariableV2", "Variable", "VarHandleOp") else None


def _const_scalar(ir: GraphIR, ref: Tuple[str, int]) -> Optional[float]:
    """Value of a scalar that is a Const, possibly behind Identity / PlaceholderWithDefault / `1 - rate` arithmetic."""
    node = ir.nodes[ref[0]]
    hops = 0
    while node.op in ("Identity", "PlaceholderWithDefault", "Cast") and node.inputs and hops < 6:
        node = ir.nodes[node.inputs[0][0]]
        hops += 1
    if node.op == "Const":
        try:
            return float(node.attrs["value"].reshape(-1)[0])
        except Exception:
            return None
    if node.op == "Sub" and len(node.inputs) == 2:
        a, b = _const_scalar(ir, node.inputs[0]), _const_scalar(ir, node.inputs[1])
        return None if a is None or b is None else a - b
    return None


def _match_dropout(ir: GraphIR, node: Node) -> Optional[Tuple[Tuple[str, int], float]]:
    """``tf.nn.dropout`` sub-graph: Mul(RealDiv(x, keep), Floor(Add(keep, RandomUniform(Shape(x))))) -> (x, keep)."""
    if node.op != "Mul" or len(node.inputs) != 2:
        return None
    for a, b in ((node.inputs[0], node.inputs[1]), (node.inputs[1], node.inputs[0])):
        div, flo = ir.nodes[a[0]], ir.nodes[b[0]]
        if div.op not in ("RealDiv", "Div") or flo.op != "Floor":
            continue
        add = ir.nodes[flo.inputs[0][0]]
        if add.op not in ("Add", "AddV2") or not any(ir.nodes[i[0]].op == "RandomUniform" for i in add.inputs):
            continue
        keep = _const_scalar(ir, div.inputs[1])
        if keep is None:
            raise UnsupportedGraph(f"dropout '{node.name}': keep_prob is a placeholder without a default - it is never fed during "
                                   "training (reference: HogwildSparkModel.py never feeds tfDropout), use tf.placeholder_with_default")
        return div.inputs[0], float(keep)
    return None


def _find_loss_core(ir: GraphIR, loss_ref: str) -> Tuple[str, Node]:
    """Walk back from the scalar loss through reductions / scalings to the per-element loss node."""
    seen = set()
    frontier = [split_ref(loss_ref)[0]]
    while frontier:
        name = frontier.pop(0)
        if name in seen:
            continue
        seen.add(name)
        node = ir.nodes[name]
        if node.op == "SoftmaxCrossEntropyWithLogits":
            return "softmax_xent", node
        if node.op == "SquaredDifference":
            return "mse", node
        if node.op == "Square" and node.inputs and ir.nodes[node.inputs[0][0]].op == "Sub":
            return "mse", ir.nodes[node.inputs[0][0]]
        if node.op in ("Sum", "Mean", "RealDiv", "Div", "DivNoNan", "Mul", "Identity", "Select", "Reshape", "Cast"):
            # follow data inputs that are not plain constants / shape plumbing
            for src, _ in node.inputs[: 1 if node.op in ("Sum", "Mean", "Reshape") else len(node.inputs)]:
                if ir.nodes[src].op not in ("Const", "Size", "Shape", "Fill", "Greater", "Equal", "ZerosLike", "OnesLike"):
                    frontier.append(src)
        else:
            continue
    raise UnsupportedGraph(f"loss '{loss_ref}' is neither softmax-cross-entropy nor mean-squared-error")


def _check_mean_reduction(ir: GraphIR, loss_ref: str, kind: str, core: Node) -> None:
    """The kernels implement the *mean* over elements (tf.losses default).  Verify numerically cheap:
    the path from core to loss must contain a Sum followed by a division, or a Mean."""
    name = split_ref(loss_ref)[0]
    ops_on_path = []
    node = ir.nodes[name]
    hops = 0
    while node.name != core.name and hops < 16:
        ops_on_path.append(node.op)
        nxt = None
        for src, _ in node.inputs:
            if ir.nodes[src].op not in ("Const", "Size", "Shape", "Cast") or src == core.name:
                nxt = ir.nodes[src]
                break
        if nxt is None:
            break
        node = nxt
        hops += 1
    if "Mean" in ops_on_path:
        if kind == "mse" and "Mul" in ops_on_path:
            raise UnsupportedGraph("scaled MSE variants are not covered by the compiled plan")
        return
    if "Sum" in ops_on_path and any(o in ops_on_path for o in ("RealDiv", "Div", "DivNoNan")):
        return
    raise UnsupportedGraph("only mean-reduced losses are covered by the compiled plan")


def compile_graph(ir: GraphIR, tf_input: str, tf_label: Optional[str], tf_output: Optional[str] = None,
                  need_loss: bool = True) -> LayerPlan:
    input_node = split_ref(tf_input)[0]
    if input_node not in ir.nodes:
        raise KeyError(f"tfInput '{tf_input}' is not in the graph")
    in_shape = ir.nodes[input_node].attrs.get("shape")
    if not in_shape or len(in_shape) != 2 or in_shape[1] is None or in_shape[1] < 0:
        raise UnsupportedGraph("the compiled plan needs a rank-2 input placeholder [batch, features]")
    input_dim = int(in_shape[1])

    loss_kind: Optional[str] = None
    pred_ref: Optional[Tuple[str, int]] = None
    target_ref: Optional[Tuple[str, int]] = None
    if need_loss:
        if not ir.losses:
            raise ValueError("the graph has no 'losses' collection entry; training needs tf.losses.* (or add_loss)")
        loss_kind, core = _find_loss_core(ir, ir.losses[0])
        _check_mean_reduction(ir, ir.losses[0], loss_kind, core)
        a, b = core.inputs[0], core.inputs[1]
        if loss_kind == "softmax_xent":
            pred_ref, target_ref = a, b               # (features=logits, labels)
        else:
            # mean_squared_error is symmetric: the side that is a placeholder is the target
            def is_ph(ref):
                n = ir.nodes[ref[0]]
                while n.op in _PASS and n.inputs:
                    n = ir.nodes[n.inputs[0][0]]
                return n.op in ("Placeholder", "PlaceholderV2")
            pred_ref, target_ref = (b, a) if is_ph(a) and not is_ph(b) else (a, b)
    else:
        if not tf_output:
            raise ValueError("inference compilation needs tfOutput")
        pred_ref = split_ref(tf_output)

    # ---- resolve the prediction-side post op (ArgMax / Softmax) for inference fetches ------------
    def strip_post(ref: Tuple[str, int]) -> Tuple[Tuple[str, int], Optional[str]]:
        n = ir.nodes[ref[0]]
        if n.op == "ArgMax":
            return n.inputs[0], "ArgMax"
        if n.op == "Softmax":
            return n.inputs[0], "Softmax"
        return ref, None

    chain_end = pred_ref
    post_of_end = None
    if not need_loss:
        chain_end, post_of_end = strip_post(pred_ref)

    # ---- walk the forward chain backwards -------------------------------------------------------
    layers_rev: List[Layer] = []
    cur = chain_end
    pending_act: Optional[Tuple[str, str]] = None      # (act, tensor name)
    pending_bias: Optional[Tuple[str, str]] = None     # (var, tensor name)
    pending_drop: Optional[Tuple[float, str]] = None   # (keep_prob, tensor name) of a dropout above the next layer found
    guard = 0
    while True:
        guard += 1
        if guard > 10000:
            raise UnsupportedGraph("forward chain too long or cyclic")
        node = ir.nodes[cur[0]]
        tname = f"{node.name}:{cur[1]}"
        if node.name == input_node:
            if pending_drop:
                raise UnsupportedGraph("dropout applied directly to the input is not covered by the compiled plan")
            break
        if node.op in _PASS:
            cur = node.inputs[0]
            continue
        drop = _match_dropout(ir, node)
        if drop is not None:
            if pending_act or pending_bias or pending_drop:
                raise UnsupportedGraph(f"dropout '{node.name}' is not directly on top of a layer output")
            if drop[1] < 1.0:
                if drop[1] <= 0.0:
                    raise UnsupportedGraph("dropout with keep_prob <= 0")
                pending_drop = (drop[1], tname)
            cur = drop[0]
            continue
        if node.op in ("Relu", "Sigmoid", "Tanh"):
            if pending_act or pending_bias:
                raise UnsupportedGraph(f"activation '{node.name}' is not directly on top of a linear layer")
            pending_act = (node.op, tname)
            cur = node.inputs[0]
            continue
        if node.op == "BiasAdd" or (node.op in ("Add", "AddV2") and _var_of(ir, node.inputs[1]) is not None):
            if pending_bias:
                raise UnsupportedGraph("two bias additions in a row")
            var = _var_of(ir, node.inputs[1])
            if var is None:
                raise UnsupportedGraph(f"BiasAdd '{node.name}' does not add a variable")
            pending_bias = (var, tname)
            cur = node.inputs[0]
            continue
        if node.op == "MatMul":
            if node.attrs.get("transpose_a") or node.attrs.get("transpose_b"):
                raise UnsupportedGraph("transposed MatMul is not covered by the compiled plan")
            kvar = _var_of(ir, node.inputs[1])
            if kvar is None:
                raise UnsupportedGraph(f"MatMul '{node.name}' right operand is not a variable")
            kin, kout = ir.nodes[kvar].attrs["shape"]
            tensors = {"linear": tname}
            if pending_bias:
                tensors["bias"] = pending_bias[1]
            if pending_act:
                tensors["act"] = pending_act[1]
            if pending_drop:
                tensors["dropout"] = pending_drop[1]
            layers_rev.append(Layer("dense", kernel=kvar, bias=pending_bias[0] if pending_bias else None,
                                    act=pending_act[0] if pending_act else None, in_shape=(kin,), out_shape=(kout,), tensors=tensors,
                                    dropout_keep=pending_drop[0] if pending_drop else 0.0))
            pending_act = pending_bias = pending_drop = None
            cur = node.inputs[0]
            continue
        if pending_drop:
            raise UnsupportedGraph(f"dropout on top of '{node.op}' ({node.name}): the compiled plan fuses dropout into dense layers only")
        if node.op == "Conv2D":
            if node.attrs.get("padding", "VALID") != "VALID" or list(node.attrs.get("strides", [1, 1, 1, 1])) != [1, 1, 1, 1] \
                    or node.attrs.get("data_format", "NHWC") != "NHWC" or list(node.attrs.get("dilations", [1, 1, 1, 1])) != [1, 1, 1, 1]:
                raise UnsupportedGraph("only NHWC / VALID / stride-1 Conv2D is covered by the compiled plan")
            kvar = _var_of(ir, node.inputs[1])
            if kvar is None:
                raise UnsupportedGraph(f"Conv2D '{node.name}' filter is not a variable")
            kh, kw, cin, cout = ir.nodes[kvar].attrs["shape"]
            tensors = {"linear": tname}
            if pending_bias:
                tensors["bias"] = pending_bias[1]
            if pending_act:
                tensors["act"] = pending_act[1]
            layers_rev.append(Layer("conv", kernel=kvar, bias=pending_bias[0] if pending_bias else None,
                                    act=pending_act[0] if pending_act else None, in_shape=(-1, -1, cin), out_shape=(-1, -1, cout),
                                    ksize=(kh, kw), tensors=tensors))
            pending_act = pending_bias = None
            cur = node.inputs[0]
            continue
        if pending_act or pending_bias:
            raise UnsupportedGraph(f"bias/activation stacked on unsupported op '{node.op}' ({node.name})")
        if node.op == "MaxPool":
            if list(node.attrs.get("ksize")) != [1, 2, 2, 1] or list(node.attrs.get("strides")) != [1, 2, 2, 1] \
                    or node.attrs.get("padding", "VALID") != "VALID":
                raise UnsupportedGraph("only 2x2 / stride-2 / VALID MaxPool is covered by the compiled plan")
            layers_rev.append(Layer("pool", ksize=(2, 2), tensors={"act": tname}))
            cur = node.inputs[0]
            continue
        if node.op == "Reshape":
            layers_rev.append(Layer("reshape", tensors={"act": tname, "shape_src": f"{node.inputs[1][0]}:{node.inputs[1][1]}"}))
            cur = node.inputs[0]
            continue
        raise UnsupportedGraph(f"op '{node.op}' ({node.name}) is not covered by the compiled plan")

    layers = layers_rev[::-1]

    # ---- shape propagation ---------------------------------------------------------------------
    shape: Tuple[int, ...] = (input_dim,)
    for l in layers:
        if l.kind == "reshape":
            src = ir.node_of(l.tensors["shape_src"])
            tgt: Optional[List[int]] = None
            if src.op == "Const":
                tgt = [int(v) for v in src.attrs["value"].reshape(-1)][1:]
            elif src.op == "Pack":                       # tf.layers.flatten: [batch, -1]
                tgt = [-1]
            if tgt is None:
                raise UnsupportedGraph("dynamic Reshape target")
            total = 1
            for d in shape:
                total *= d
            if tgt.count(-1) == 1:
                known = 1
                for d in tgt:
                    if d != -1:
                        known *= d
                tgt[tgt.index(-1)] = total // known
            l.in_shape, l.out_shape = shape, tuple(tgt)
        elif l.kind == "dense":
            if len(shape) != 1 or shape[0] != l.in_shape[0]:
                raise UnsupportedGraph(f"dense layer '{l.kernel}' expects {l.in_shape} features but receives {shape}")
        elif l.kind == "conv":
            if len(shape) != 3 or shape[2] != l.in_shape[2]:
                raise UnsupportedGraph(f"conv layer '{l.kernel}' expects NHWC input, got per-sample shape {shape}")
            l.in_shape = shape
            l.out_shape = (shape[0] - l.ksize[0] + 1, shape[1] - l.ksize[1] + 1, l.out_shape[2])
        elif l.kind == "pool":
            if len(shape) != 3:
                raise UnsupportedGraph("MaxPool on non-image tensor")
            l.in_shape, l.out_shape = shape, (shape[0] // 2, shape[1] // 2, shape[2])
        shape = l.out_shape

    if not any(l.kind in ("dense", "conv") for l in layers):
        raise UnsupportedGraph("no trainable layer between tfInput and the loss")

    # ---- target ----------------------------------------------------------------------------------
    label_name = None
    target_is_input = False
    label_dim = 0
    if need_loss:
        tnode = ir.nodes[target_ref[0]]
        while tnode.op in _PASS and tnode.inputs:
            tnode = ir.nodes[tnode.inputs[0][0]]
        if tnode.name == input_node:
            target_is_input = True
            label_dim = input_dim
        elif tnode.op in ("Placeholder", "PlaceholderV2"):
            label_name = f"{tnode.name}:0"
            tshape = tnode.attrs.get("shape")
            label_dim = int(tshape[1]) if tshape and len(tshape) == 2 and tshape[1] and tshape[1] > 0 else layers[-1].out_features
            if tf_label is not None and split_ref(tf_label)[0] != tnode.name:
                raise UnsupportedGraph(f"the loss target is placeholder '{tnode.name}' but tfLabel is '{tf_label}'")
        else:
            raise UnsupportedGraph("loss target must be a placeholder or the input itself")
        if label_dim != layers[-1].out_features:
            raise UnsupportedGraph("label width does not match the network output width")
        if loss_kind == "softmax_xent" and layers[-1].act is not None:
            raise UnsupportedGraph("softmax cross-entropy expects raw logits")
        if layers[-1].dropout_keep:
            raise UnsupportedGraph("dropout on the network output is not covered by the compiled plan")

    # ---- output fetch ------------------------------------------------------------------------------
    out_spec: Optional[OutputSpec] = None
    if tf_output:
        ref, post = strip_post(split_ref(tf_output))
        want = f"{ref[0]}:{ref[1]}"
        if split_ref(want)[0] == input_node:
            out_spec = OutputSpec(-1, "act", post)
        else:
            for i, l in enumerate(layers):
                for stage in ("dropout", "act", "bias", "linear"):
                    if l.tensors.get(stage) == want:
                        # the fetched stage must be the layer's final stage (we do not keep pre-activations)
                        final = "dropout" if l.dropout_keep else ("act" if l.act or l.kind in ("pool", "reshape") else ("bias" if l.bias else "linear"))
                        if stage != final:
                            raise UnsupportedGraph(f"tfOutput '{tf_output}' fetches a pre-activation tensor")
                        out_spec = OutputSpec(i, stage, post)
            if out_spec is None and need_loss:
                out_spec = None            # tfOutput may be unrelated to training (reference cnn_example); resolved at predict time
            elif out_spec is None:
                raise UnsupportedGraph(f"tfOutput '{tf_output}' is not a tensor of the forward chain")

    return LayerPlan(layers=layers, loss=loss_kind, input_name=f"{input_node}:0", label_name=label_name,
                     target_is_input=target_is_input, input_dim=input_dim, label_dim=label_dim, output=out_spec,
                     var_order=[v.name for v in ir.trainable])
