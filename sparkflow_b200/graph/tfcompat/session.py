"""``tf.Session`` on top of the PyTorch graph interpreter (enough for TF-1.x style eval/inspection)."""
from __future__ import annotations

import threading
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from . import core


class _Local(threading.local):
    def __init__(self):
        self.stack: List["Session"] = []


_SESSIONS = _Local()


def get_default_session() -> "Session":
    if not _SESSIONS.stack:
        raise ValueError("Cannot evaluate tensor using `eval()`: No default session is registered.")
    return _SESSIONS.stack[-1]


class Session:
    def __init__(self, target: str = "", graph: Optional[core.Graph] = None, config=None):
        self.graph = graph or core.get_default_graph()
        self._values: Dict[str, torch.Tensor] = {}
        self._program = None
        self._n_nodes = -1

    # context manager / default-session handling
    def __enter__(self):
        _SESSIONS.stack.append(self)
        self._graph_ctx = self.graph.as_default()
        self._graph_ctx.__enter__()
        return self

    def __exit__(self, *exc):
        self._graph_ctx.__exit__(*exc)
        _SESSIONS.stack.pop()
        return False

    def as_default(self):
        return self

    def close(self):
        self._values.clear()

    def _prog(self):
        from ..executor import GraphProgram
        from ..ir import GraphIR

        if self._program is None or self._n_nodes != len(self.graph.nodes):
            self._program = GraphProgram(GraphIR.from_metagraph(core.export_meta_graph(graph=self.graph)))
            self._n_nodes = len(self.graph.nodes)
        return self._program

    def _initialize(self, var_node: str):
        prog = self._prog()
        info = next((v for v in prog.ir.variables + prog.ir.trainable if v.name == var_node), None)
        if info is None or info.initial_value is None:
            raise ValueError(f"cannot initialise variable {var_node}")
        with torch.no_grad():
            self._values[var_node] = prog.run([info.initial_value], {}, self._values)[0].to(torch.float32)

    def set_variable(self, name: str, value) -> None:
        self._values[name.split(":")[0]] = torch.as_tensor(np.asarray(value, dtype=np.float32))

    def run(self, fetches, feed_dict: Optional[Dict[Any, Any]] = None, **_unused):
        single = not isinstance(fetches, (list, tuple))
        items = [fetches] if single else list(fetches)
        feeds = {}
        for k, v in (feed_dict or {}).items():
            feeds[k.name if hasattr(k, "name") else str(k)] = v
        out: List[Any] = []
        prog = self._prog()
        for it in items:
            if isinstance(it, core.Variable):
                it = it.value()
            if isinstance(it, str):
                it = self.graph.as_graph_element(it)
            if isinstance(it, core.Operation):
                if it.type == "NoOp":                     # global_variables_initializer
                    for ref in it.node_def.get("input", []):
                        tgt = ref.lstrip("^")
                        if tgt.endswith("/Assign"):
                            self._initialize(tgt[: -len("/Assign")])
                elif it.type == "Assign":
                    self._initialize(it.node_def["input"][0])
                out.append(None)
                continue
            with torch.no_grad():
                val = prog.run([it.name], feeds, self._values)[0]
            out.append(val.detach().cpu().numpy() if isinstance(val, torch.Tensor) else np.asarray(val))
        return out[0] if single else out


InteractiveSession = Session
