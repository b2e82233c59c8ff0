"""Multi-GPU (one process per GPU, CUDA-IPC master, NVLink push/pull) end-to-end checks."""
import json
import os
import subprocess
import sys
import textwrap

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys
    import numpy as np, torch
    sys.path.insert(0, {root!r})
    from sparkflow_b200.models import zoo
    from sparkflow_b200.graph.executor import GraphProgram
    from sparkflow_b200.graph.ir import GraphIR
    from sparkflow_b200.ops.optimizers import OptimizerSpec
    from sparkflow_b200.parallel import dist as D
    from sparkflow_b200.parallel.session import TrainingSession
    lock = sys.argv[1].startswith("lock")
    push_mode = "direct" if sys.argv[1].endswith("direct") else ("sharded" if sys.argv[1].endswith("sharded") else "served")
    ctx = D.get_context()
    rng = np.random.default_rng(11)
    centers = rng.normal(0, 1, (10, 784)).astype(np.float32)
    parts = []
    for p in range(ctx.world):
        lab = rng.integers(0, 10, 1500)
        parts.append((centers[lab] + 0.3 * rng.normal(0, 1, (1500, 784)).astype(np.float32), np.eye(10, dtype=np.float32)[lab], lab))
    graph = zoo.build("simple_dnn")
    sess = TrainingSession(graph, "x:0", "y:0", OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.002)), acquire_lock=lock,
                           iters=6, mini_batch=300, shuffle=False, engine="b200", seed=3, push_mode=push_mode)
    sess.train_partitions([(x, y) for x, y, _ in parts])
    import time
    t0 = time.time()
    while sess.counters()["pushes"] < ctx.world * 30 and time.time() - t0 < 20:
        time.sleep(0.02)            # served mode: the applier may still be draining the last mailboxes
    D.barrier(ctx)
    w = sess.weights()
    c = sess.counters()
    prog = GraphProgram(GraphIR.from_metagraph(graph))
    X = np.concatenate([p[0] for p in parts]); L = np.concatenate([p[2] for p in parts])
    acc = float((prog.forward("out:0", {{"x:0": X}}, w).numpy() == L).mean())
    with open(os.path.join(sys.argv[2], "rank%d.json" % ctx.rank), "w") as fh:
        json.dump({{"rank": ctx.rank, "acc": acc, "counters": c, "w0": float(np.abs(w[0]).sum()), "master": sess.master_desc()}}, fh)
    sess.close()
""")


def _run(n, mode, tmp_path, port):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), mode, str(tmp_path)]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                          env=dict(os.environ, GLOO_SOCKET_IFNAME="lo"))
    assert proc.returncode == 0, proc.stdout[-4000:]
    return [json.load(open(tmp_path / f)) for f in sorted(os.listdir(tmp_path)) if f.startswith("rank")]


MODES = ["hogwild", "lock", "hogwild-direct", "lock-direct", "hogwild-sharded", "lock-sharded"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("mode", MODES)
def test_two_gpu_async_parameter_server(mode, tmp_path):
    n = min(torch.cuda.device_count(), 8)
    res = _run(n, mode, tmp_path, 29541 + MODES.index(mode))
    assert len(res) == n
    total = n * 6 * 5                       # ranks x iters x batches
    for r in res:
        assert r["counters"]["lock"] == 0
        assert r["counters"]["pushes"] == total, r["counters"]
        assert r["acc"] > 0.9, r
    assert len({round(r["w0"], 3) for r in res}) == 1      # every rank reads the same master
