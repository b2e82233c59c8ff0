"""The baseline to beat: the reference's training semantics as a stock PyTorch + NCCL(+cuBLAS) build.

Per BASELINE.md the reference itself cannot run here (TF-1.x / pyspark / flask / JVM are absent), so the
comparison target is "param-server replaced by NCCL broadcast/reduce": every step each rank
  pull  = ``dist.broadcast(flat_params, src=0)``
  grad  = ``torch`` autograd on an ``nn.Sequential`` MLP (bf16 autocast -> cuBLAS tensor-core GEMMs)
  push  = ``dist.reduce(flat_grad, dst=0)``; rank 0 applies ``torch.optim.Adam`` (fused) to the master copy.
None of sparkflow_b200's kernels or engine is on this path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn


def build_mlp(dims, acts):
    layers = []
    for i in range(len(dims) - 1):
        layers.append(nn.Linear(dims[i], dims[i + 1]))
        if acts[i] == "relu":
            layers.append(nn.ReLU())
        elif acts[i] == "sigmoid":
            layers.append(nn.Sigmoid())
    return nn.Sequential(*layers)


class NcclBaselineWorker:
    def __init__(self, dims, acts, loss="softmax_xent", lr=1e-3, device="cuda", world=1, rank=0, use_graph=False):
        self.device, self.world, self.rank = torch.device(device), world, rank
        self.model = build_mlp(dims, acts).to(self.device)
        self.params = [p for p in self.model.parameters()]
        self.flat = torch.nn.utils.parameters_to_vector(self.params).detach().clone()
        self.master = self.flat.clone().requires_grad_(False)
        self.master_param = nn.Parameter(self.master)
        self.opt = torch.optim.Adam([self.master_param], lr=lr, fused=True) if rank == 0 else None
        self.loss_kind = loss
        self.flat_grad = torch.zeros_like(self.flat)

    def step(self, x, y):
        # pull
        if self.world > 1:
            dist.broadcast(self.master_param.data, src=0)
        torch.nn.utils.vector_to_parameters(self.master_param.data, self.params)
        # forward / backward (bf16 tensor cores through cuBLAS)
        for p in self.params:
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = self.model(x)
        out = out.float()
        if self.loss_kind == "softmax_xent":
            loss = -(y * torch.log_softmax(out, dim=1)).sum(dim=1).mean()
        else:
            loss = ((out - y) ** 2).mean()
        loss.backward()
        # push
        torch.cat([p.grad.reshape(-1) for p in self.params], out=self.flat_grad)
        if self.world > 1:
            dist.reduce(self.flat_grad, dst=0)
        if self.rank == 0:
            self.master_param.grad = self.flat_grad
            self.opt.step()
        return loss
