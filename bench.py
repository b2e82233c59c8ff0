#!/usr/bin/env python
"""Flagship benchmark: MNIST-DNN (simple_dnn 784-256-256-10, Adam, minibatch 300 per worker) samples/s.

    python bench.py --gpus 1 --steps 200 --warmup 20
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

One rank per GPU.  A *step* is one full parameter-server iteration of every worker: pull the master
parameters, forward + backward on a 300-row minibatch, push the gradient (ONE Adam step on the master
per push).  Weak scaling: per-GPU work is fixed, ``value`` is the whole-job samples/s.

Two numbers per run:
  value      device-timed: K CUDA-graph replays of the compiled step, each bracketed by its own pair of
             CUDA events with a 256 MiB L2 flush in between (outside the events); max over ranks.
  e2e.value  the same steps driven through the public engine API (TrainingSession / B200Engine.train):
             every step copies its minibatch host->device from a pinned partition that is larger than L2
             and reads the step's loss back device->host; one event pair around all K steps; max over ranks.

``--impl reference`` would run the unmodified lifeomic/sparkflow from baseline/_ref; it needs TensorFlow
1.x + pyspark + flask + a JVM, none of which exist in this image, so it reports ``unavailable``.
``--impl nccl`` runs the stock PyTorch + NCCL build of the same semantics (baseline/nccl_baseline.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MNIST-DNN samples/sec (device-timed, max over ranks)"
BATCH = 300
DIMS = [784, 256, 256, 10]
# --model: the other workloads the reference ships (BASELINE.json configs 3-5); (input dim, label dim or 0, description)
MODELS = {
    "simple_dnn": (784, 10, "simple_dnn 784-256-256-10"),
    "cnn": (784, 10, "cnn_example conv5x5x32-pool-conv3x3x64-pool-dense256-10"),
    "autoencoder": (784, 0, "autoencoder_example 784-256-128-256-784"),
    "autoencoder_small": (784, 0, "autoencoder 784-32-784"),
    "wide_dnn": (4096, 1000, "wide_dnn 4096-4096x4-1000"),
}


def reference_arm(args) -> int:
    why = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import importlib

        for dep in ("tensorflow", "pyspark", "flask"):
            try:
                importlib.import_module(dep)
            except Exception as exc:
                why = f"reference needs {dep} ({type(exc).__name__}); TF-1.x/pyspark/flask/JVM are not in this image or /opt/wheelhouse"
                break
        if why is None:
            why = "reference dependencies import but the TF-1.x graph API (tf.placeholder/tf.layers/tf.Session) is required"
    except Exception as exc:  # pragma: no cover
        why = f"{type(exc).__name__}: {exc}"
    print(json.dumps({"impl": "reference", "unavailable": why}))
    return 0


def _max_over_ranks(ctx, value: float) -> float:
    from sparkflow_b200.parallel import dist as D

    return max(D.all_gather_object(ctx, float(value)))


def _synthetic_partition(rows: int, seed: int, in_dim: int = DIMS[0], n_classes: int = DIMS[-1]):
    import numpy as np

    rng = np.random.default_rng(seed)
    x = rng.random((rows, in_dim), dtype=np.float32)
    y = np.eye(n_classes, dtype=np.float32)[rng.integers(0, n_classes, rows)] if n_classes else None
    return x, y


def run_ours(args, ctx) -> dict:
    import torch

    from sparkflow_b200.models import zoo
    from sparkflow_b200.ops.optimizers import OptimizerSpec
    from sparkflow_b200.parallel import dist as D
    from sparkflow_b200.parallel.session import TrainingSession
    from sparkflow_b200.utils.clocks import ClockSampler

    dev = torch.device("cuda", ctx.local_rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    lock = args.mode == "lock"
    spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.001, beta1=0.9, beta2=0.999))
    in_dim, n_classes, model_desc = MODELS[args.model]
    sess = TrainingSession(zoo.build(args.model), "x:0", "y:0" if n_classes else None, spec, acquire_lock=lock, iters=1, mini_batch=BATCH,
                           mini_stochastic_iters=1, shuffle=True, engine="b200", seed=1234, pull_mode=args.pull_mode,
                           push_mode=args.push_mode, devices=[dev.index]).open()
    eng = sess.make_engine(dev)
    rows = max(args.partition_rows, BATCH * 2)
    if in_dim * rows * 4 > (1 << 30):
        rows = max(BATCH * 2, (1 << 30) // (in_dim * 4))
    x, y = _synthetic_partition(rows, seed=100 + ctx.rank, in_dim=in_dim, n_classes=n_classes)
    eng.load_partition(x, y)
    K, W = args.steps, args.warmup
    n_batches = rows // BATCH

    def e2e_steps(n, start):
        starts = [((start + k) % n_batches) * BATCH for k in range(n)]
        if args.python_loop:
            for r in starts:
                eng.train(slice(r, r + BATCH), pull=True)
        else:
            eng.train_contiguous(starts, BATCH, pull=True)       # native StepDriver: H2D + graph replay + loss D2H per step

    # ---------------- e2e: public engine API, H2D of every minibatch + D2H of every loss ----------------
    e2e_steps(W, 0)
    eng.finish()
    h2d0, d2h0 = eng.h2d_bytes, eng.d2h_bytes
    sess.quiesce()                       # barrier + device-wide torch.cuda.synchronize() (applier paused around it)
    sampler = ClockSampler(dev.index or 0).start() if ctx.rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    host0 = list(eng._driver.host_ns()) if getattr(eng, "_driver", None) is not None else None
    e0.record(eng.w.stream)
    e2e_steps(K, W)
    e1.record(eng.w.stream)
    eng.finish()                         # stream sync + wait until the master has applied this rank's last push
    t_wall = time.perf_counter() - t_wall0
    host_us = None
    if host0 is not None:
        names = ("wait_slot", "h2d_enqueue", "event_handoff", "graph_launch", "record")
        host_us = {n: (b - a) / K / 1e3 for n, a, b in zip(names, host0, eng._driver.host_ns())}
    sess.quiesce()
    e2e_ms = _max_over_ranks(ctx, e0.elapsed_time(e1))
    wall_ms = _max_over_ranks(ctx, t_wall * 1e3)
    last_loss = eng.last_loss()
    h2d_per_step = (eng.h2d_bytes - h2d0) // K
    d2h_per_step = (eng.d2h_bytes - d2h0) // K

    # ---------------- device-timed: graph replays, L2 flushed between steps ----------------
    w = eng.w
    plan, bufs = w.build_plan(BATCH, 0, with_pull=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = w.stream
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with torch.cuda.stream(st):
        for _ in range(max(W, 3)):
            w.run_plan(plan)
        st.synchronize()
        sess.quiesce()
        for a, b in evs:
            flush.zero_()
            a.record(st)
            w.run_plan(plan)
            b.record(st)
        st.synchronize()
    sess.quiesce()
    dev_ms = _max_over_ranks(ctx, sum(a.elapsed_time(b) for a, b in evs))
    # back-to-back replays without the flush (what the training loop actually sees)
    with torch.cuda.stream(st):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier(ctx)
        a.record(st)
        for _ in range(K):
            w.run_plan(plan)
        b.record(st)
        st.synchronize()
    warm_ms = _max_over_ranks(ctx, a.elapsed_time(b))
    clocks = sampler.stop() if sampler else {}
    counters = sess.counters()
    launches = len(plan)
    res = {
        "metric": METRIC if args.model == "simple_dnn" else f"{args.model} samples/sec (device-timed, max over ranks)",
        "value": ctx.world * BATCH * K / (dev_ms / 1e3), "unit": "samples/s", "n_gpus": ctx.world, "steps": K,
        "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "impl": "sparkflow_b200",
        "config": {"model": model_desc, "global_batch": ctx.world * BATCH, "seq_len": 1,
                   "parallelism": f"async-ps dp{ctx.world} ({'rw-lock' if lock else 'hogwild'}, master on gpu0)",
                   "optimizer": "adam(1e-3), one step per push on the master", "pull_mode": w.pull_mode, "push_mode": sess.push_mode,
                   "l2": "device-timed value: 256 MiB flush between steps (outside the event pairs); e2e: pinned partition "
                         f"{rows * in_dim * 4 >> 20} MiB > L2, a fresh minibatch H2D every step",
                   "kernels_per_step": plan.names(), "cuda_graph": bool(w.use_graphs)},
        "e2e": {"value": ctx.world * BATCH * K / (e2e_ms / 1e3), "unit": "samples/s", "ms_per_step": e2e_ms / K,
                "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": int(d2h_per_step), "wall_ms_per_step": wall_ms / K,
                "host_us_per_step": host_us},
        "gpu_launches": int(launches * K * ctx.world),
        "warm_cache_ms_per_step": warm_ms / K,
        "clocks": clocks, "final_loss": last_loss, "master_counters": counters,
    }
    sess.close()
    return res


def run_nccl(args, ctx) -> dict:
    import torch
    import torch.distributed as dist

    from baseline.nccl_baseline import NcclBaselineWorker
    from sparkflow_b200.parallel import dist as D
    from sparkflow_b200.utils.clocks import ClockSampler

    dev = torch.device("cuda", ctx.local_rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    torch.manual_seed(1234)
    wk = NcclBaselineWorker(DIMS, ["relu", "relu", None], device=dev, world=ctx.world, rank=ctx.rank)
    rows = max(args.partition_rows, BATCH * 2)
    x, y = _synthetic_partition(rows, seed=100 + ctx.rank)
    xp, yp = torch.from_numpy(x).pin_memory(), torch.from_numpy(y).pin_memory()
    n_batches = rows // BATCH
    K, W = args.steps, args.warmup
    xs = [torch.empty(BATCH, DIMS[0], device=dev) for _ in range(2)]
    ys = [torch.empty(BATCH, DIMS[-1], device=dev) for _ in range(2)]
    loss_host = torch.zeros(1).pin_memory()

    def step(k):
        r = (k % n_batches) * BATCH
        s = k & 1
        xs[s].copy_(xp[r:r + BATCH], non_blocking=True)
        ys[s].copy_(yp[r:r + BATCH], non_blocking=True)
        loss = wk.step(xs[s], ys[s])
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

    for k in range(W):
        step(k)
    torch.cuda.synchronize(dev)
    if ctx.world > 1:
        dist.barrier()
    sampler = ClockSampler(dev.index or 0).start() if ctx.rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        step(W + k)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = _max_over_ranks(ctx, e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else {}
    return {"metric": METRIC, "value": ctx.world * BATCH * K / (ms / 1e3), "unit": "samples/s", "n_gpus": ctx.world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "impl": "nccl_baseline",
            "config": {"model": "simple_dnn 784-256-256-10", "global_batch": ctx.world * BATCH, "seq_len": 1,
                       "parallelism": f"nccl broadcast/reduce dp{ctx.world}", "l2": f"pinned partition {rows * DIMS[0] * 4 >> 20} MiB > L2"},
            "e2e": {"value": ctx.world * BATCH * K / (ms / 1e3), "unit": "samples/s", "h2d_bytes_per_step": BATCH * (DIMS[0] + DIMS[-1]) * 4,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": 0, "clocks": clocks, "final_loss": float(loss_host[0])}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "nccl"])
    ap.add_argument("--model", default="simple_dnn", choices=sorted(MODELS), help="workload (the headline metric is simple_dnn)")
    ap.add_argument("--mode", default="lock", choices=["lock", "hogwild"], help="acquire_lock=True (BASELINE config 2) or Hogwild")
    ap.add_argument("--pull-mode", default=None, choices=[None, "copy", "direct"])
    ap.add_argument("--push-mode", default=None, choices=[None, "direct", "served"],
                    help="direct: worker applies the optimizer on master memory over NVLink; served: mailbox + applier kernel "
                         "on the master GPU (default: served when more than one GPU shares the master)")
    ap.add_argument("--partition-rows", type=int, default=50_100, help="rows of the pinned per-rank partition (157 MiB > L2)")
    ap.add_argument("--with-nccl-baseline", action="store_true", help="also time the NCCL baseline in the same launch")
    ap.add_argument("--python-loop", action="store_true", help="drive the e2e steps from Python instead of the native StepDriver")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    import torch

    if not torch.cuda.is_available():
        print(json.dumps({"metric": METRIC, "error": "no CUDA device: bench.py measures the B200 engine"}))
        return 1
    from sparkflow_b200.parallel import dist as D

    ctx = D.get_context()
    if ctx.world != args.gpus:
        if ctx.rank == 0:
            print(json.dumps({"metric": METRIC, "error": f"--gpus {args.gpus} but WORLD_SIZE is {ctx.world}; launch with torchrun"}))
        return 1
    res = run_ours(args, ctx) if args.impl == "ours" else run_nccl(args, ctx)
    if args.impl == "ours" and args.with_nccl_baseline:
        base = run_nccl(args, ctx)
        res["nccl_baseline_same_run"] = {"value": base["value"], "ms_per_step": base["ms_per_step"]}
        res["e2e_vs_nccl_baseline"] = res["e2e"]["value"] / base["value"]
    if ctx.rank == 0:
        print(json.dumps(res))
    D.barrier(ctx)
    return 0


if __name__ == "__main__":
    sys.exit(main())
