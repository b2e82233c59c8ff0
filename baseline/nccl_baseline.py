"""The comparison arm: the reference's training semantics as a stock PyTorch + NCCL (+cuBLAS / cuDNN) build.

Per BASELINE.md the reference itself cannot run here (TF-1.x / pyspark / flask / JVM are absent), so the
comparison target is "parameter server replaced by NCCL broadcast / reduce".  Every step each rank does

  pull  = ``dist.broadcast(flat_params, src=0)``                        (NCCL)
  grad  = autograd on a stock ``nn.Module`` under bf16 autocast          (cuBLAS / cuDNN tensor-core kernels)
  push  = ``dist.reduce(flat_grad, dst=0)`` + ONE fused Adam step on rank 0's master copy

None of sparkflow_b200's kernels or engine is on this path.  It is built to be a *strong* baseline:

* parameters and gradients are views into ONE flat fp32 buffer each (no per-step ``cat`` / ``vector_to_parameters``),
* ``torch.optim.Adam(fused=True, capturable=True)``: one multi-tensor kernel, no host sync,
* the WHOLE step (both collectives included) is captured into a CUDA graph and replayed (``use_graph=True``),
* optionally the forward/backward is ``torch.compile``d before capture (``compile=True``).

Workloads: the MLPs (simple_dnn, autoencoders, wide_dnn) and the reference's CNN (cuDNN, channels-last).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F


def build_mlp(dims: Sequence[int], acts: Sequence[Optional[str]]) -> nn.Module:
    layers: List[nn.Module] = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if acts[i] == "relu":
            layers.append(nn.ReLU())
        elif acts[i] == "sigmoid":
            layers.append(nn.Sigmoid())
        elif acts[i] == "tanh":
            layers.append(nn.Tanh())
    return nn.Sequential(*layers)


class RefCNN(nn.Module):
    """examples/cnn_example.py:10-22 of the reference: conv5x5x32-relu-pool2-conv3x3x64-relu-pool2-flatten-dense10."""

    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(1, 32, 5)
        self.c2 = nn.Conv2d(32, 64, 3)
        self.fc = nn.Linear(5 * 5 * 64, 10)

    def forward(self, x):
        x = x.view(-1, 1, 28, 28).contiguous(memory_format=torch.channels_last)
        x = F.max_pool2d(F.relu(self.c1(x)), 2)
        x = F.max_pool2d(F.relu(self.c2(x)), 2)
        return self.fc(x.permute(0, 2, 3, 1).reshape(x.shape[0], -1))


MODEL_SPECS = {
    "simple_dnn": dict(dims=[784, 256, 256, 10], acts=["relu", "relu", None], loss="softmax_xent"),
    "autoencoder": dict(dims=[784, 256, 128, 256, 784], acts=["relu", "sigmoid", "relu", "sigmoid"], loss="mse"),
    "autoencoder_small": dict(dims=[784, 32, 784], acts=["sigmoid", "sigmoid"], loss="mse"),
    "wide_dnn": dict(dims=[4096, 4096, 4096, 4096, 4096, 1000], acts=["relu"] * 4 + [None], loss="softmax_xent"),
    "cnn": dict(cnn=True, loss="softmax_xent"),
}


class NcclBaselineWorker:
    """One rank of the NCCL build.  ``step(x, y)`` runs one full pull / fwd / bwd / push(+Adam) iteration on the device
    tensors ``x`` / ``y`` (static buffers when graphed: refill them, then call ``step``)."""

    def __init__(self, model: str = "simple_dnn", batch: int = 300, lr: float = 1e-3, device="cuda", world: int = 1, rank: int = 0,
                 use_graph: bool = True, compile: bool = False):
        spec = MODEL_SPECS[model]
        self.device, self.world, self.rank = torch.device(device), world, rank
        self.loss_kind = spec["loss"]
        net = RefCNN() if spec.get("cnn") else build_mlp(spec["dims"], spec["acts"])
        net = net.to(self.device)
        if spec.get("cnn"):
            net = net.to(memory_format=torch.channels_last)
        self.model = net
        self.in_dim = 784 if spec.get("cnn") else spec["dims"][0]
        self.out_dim = 10 if spec.get("cnn") else spec["dims"][-1]
        self.target_is_input = self.loss_kind == "mse"
        # flat parameter / gradient buffers; every parameter and its .grad are views into them
        params = list(net.parameters())
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, device=self.device)
        self.flat_grad = torch.zeros(n, device=self.device)
        off = 0
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
            off += k
        self.params = params
        self.master = nn.Parameter(self.flat)          # the optimizer sees ONE tensor: a single fused multi-tensor launch
        self.master.grad = self.flat_grad
        self.opt = torch.optim.Adam([self.master], lr=lr, fused=True, capturable=True) if rank == 0 else None
        self.x = torch.zeros(batch, self.in_dim, device=self.device)
        self.y = None if self.target_is_input else torch.zeros(batch, self.out_dim, device=self.device)
        self.loss = torch.zeros((), device=self.device)
        self._fwd = torch.compile(self._fwd_loss) if compile else self._fwd_loss
        self.compiled = bool(compile)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graphed = False
        if use_graph:
            self._capture()

    # ------------------------------------------------------------------------------------------
    def _fwd_loss(self, x, y):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = self.model(x)
        out = out.float()
        if self.loss_kind == "softmax_xent":
            return -(y * torch.log_softmax(out, dim=1)).sum(dim=1).mean()
        return ((out - (x if y is None else y)) ** 2).mean()

    def _step_impl(self) -> None:
        if self.world > 1:
            dist.broadcast(self.flat, src=0)                      # pull
        self.flat_grad.zero_()
        loss = self._fwd(self.x, self.y)
        loss.backward()                                           # accumulates into the flat gradient views
        self.loss.copy_(loss.detach())
        if self.world > 1:
            dist.reduce(self.flat_grad, dst=0)                    # push
        if self.rank == 0:
            self.opt.step()                                       # one fused Adam step on the master copy

    def _capture(self) -> None:
        try:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(3):                                # warm-up: cuBLAS handles, NCCL channels, Adam state
                    self._step_impl()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            if self.world > 1:
                dist.barrier()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step_impl()
            self.graph, self.graphed = g, True
        except Exception as exc:  # pragma: no cover - depends on the NCCL / driver combination of the box
            self.graph, self.graphed = None, False
            self.capture_error = f"{type(exc).__name__}: {exc}"
            torch.cuda.synchronize(self.device)

    def step(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x, non_blocking=True)
        if y is not None and self.y is not None and y.data_ptr() != self.y.data_ptr():
            self.y.copy_(y, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_impl()
        return self.loss
