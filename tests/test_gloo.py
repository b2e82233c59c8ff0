"""BASELINE.json config 1: simple_dnn Hogwild, world_size=2, CPU / gloo (plumbing, no GPU)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import json, os, sys
    import numpy as np
    sys.path.insert(0, {root!r})
    os.environ["SPARKFLOW_ENGINE"] = "torch"
    from sparkflow_b200 import compat; compat.install()
    from sparkflow_b200.models import zoo
    from sparkflow_b200.graph.executor import GraphProgram
    from sparkflow_b200.graph.ir import GraphIR
    from sparkflow_b200.ops.optimizers import OptimizerSpec
    from sparkflow_b200.spark import SparkSession
    from sparkflow_b200.HogwildSparkModel import HogwildSparkModel
    from sparkflow_b200.parallel import dist as D
    lock = sys.argv[1] == "lock"
    n_part = int(sys.argv[2])
    out_dir = sys.argv[3]
    ctx = D.get_context()
    assert ctx.world == 2
    rng = np.random.default_rng(7)                       # identical data on every rank
    centers = rng.normal(0, 1, (10, 784)).astype(np.float32)
    lab = rng.integers(0, 10, 400)
    X = centers[lab] + 0.2 * rng.normal(0, 1, (400, 784)).astype(np.float32)
    Y = np.eye(10, dtype=np.float32)[lab]
    spark = SparkSession.builder.master("local[2]").getOrCreate()
    rdd = spark.sparkContext.parallelize([(X[i], Y[i]) for i in range(400)], n_part)
    graph = zoo.build("simple_dnn")
    model = HogwildSparkModel(tensorflowGraph=graph, iters=3, tfInput="x:0", tfLabel="y:0", acquire_lock=lock,
                              optimizer=OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.002)), mini_batch=50, seed=5)
    l0 = None
    weights = model.train(rdd)
    prog = GraphProgram(GraphIR.from_metagraph(graph))
    loss = prog.loss({{"x:0": X, "y:0": Y}}, weights)
    acc = float((prog.forward("out:0", {{"x:0": X}}, weights).numpy() == lab).mean())
    # one file per rank: interleaved rank stdout is not parseable
    with open(os.path.join(out_dir, "rank%d.json" % ctx.rank), "w") as fh:
        json.dump({{"rank": ctx.rank, "loss": loss, "acc": acc, "n": len(weights), "w0": float(np.abs(weights[0]).sum())}}, fh)
""")


def _run(mode: str, tmp_path, n_part: int = 2, port: int = 29533):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, GLOO_SOCKET_IFNAME="lo", CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script), mode, str(n_part), str(tmp_path)]
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:]
    import json

    return [json.load(open(tmp_path / f"rank{r}.json")) for r in range(2)]


def test_hogwild_world_size_2_gloo(tmp_path):
    res = _run("hogwild", tmp_path)
    assert sorted(r["rank"] for r in res) == [0, 1]
    assert res[0]["w0"] == res[1]["w0"] and res[0]["n"] == 6          # every rank returns the same master weights
    assert res[0]["loss"] < 1.0 and res[0]["acc"] > 0.8                 # 24 pushes of Adam on separable blobs


def test_locked_world_size_2_gloo(tmp_path):
    res = _run("lock", tmp_path, port=29534)
    assert res[0]["w0"] == res[1]["w0"] and res[0]["acc"] > 0.8


def test_more_partitions_than_ranks_gloo(tmp_path):
    """6 partitions on 2 ranks: every rank trains 3 partitions on threads that share ONE transport (used to hang)."""
    res = _run("hogwild", tmp_path, n_part=6, port=29535)
    assert res[0]["w0"] == res[1]["w0"] and res[0]["acc"] > 0.8


def test_fewer_partitions_than_ranks_gloo(tmp_path):
    """1 partition on 2 ranks: the idle rank still closes its transport exactly once (no 10 s join timeout)."""
    res = _run("lock", tmp_path, n_part=1, port=29536)
    assert res[0]["w0"] == res[1]["w0"]
