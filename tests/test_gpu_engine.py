"""End-to-end numerics of the compiled B200 step (pull -> fwd -> loss -> bwd -> push) against the
PyTorch-autograd oracle running the reference semantics on the CPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from sparkflow_b200.graph.executor import GraphProgram
from sparkflow_b200.graph.ir import GraphIR
from sparkflow_b200.models import zoo
from sparkflow_b200.models.compiler import compile_graph
from sparkflow_b200.ops.layout import ParamLayout
from sparkflow_b200.ops.optimizers import OptimizerSpec
from sparkflow_b200.parallel.device_engine import DeviceWorker, MasterState, plan_publish_needs
from sparkflow_b200.parallel.param_server import LocalTransport, ParameterServer
from sparkflow_b200.parallel.worker import B200Engine, TorchEngine

CASES = {
    "cnn": ("x:0", "y:0", 784, 10, "onehot"),
    "simple_dnn": ("x:0", "y:0", 784, 10, "onehot"),
    "autoencoder": ("x:0", None, 784, 0, None),
    "test_mlp": ("x:0", "y:0", 10, 1, "binary"),
    "test_autoencoder": ("x:0", None, 10, 0, None),
}


def _data(n, d, c, kind, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d), dtype=np.float32)
    if kind == "onehot":
        Y = np.eye(c, dtype=np.float32)[rng.integers(0, c, n)]
    elif kind == "binary":
        Y = rng.integers(0, 2, (n, 1)).astype(np.float32)
    else:
        Y = None
    return X, Y


def _setup(name, spec, lock=False, pull_mode="copy", served=False):
    tf_in, tf_lab, d, c, kind = CASES[name]
    ir = GraphIR.from_metagraph(zoo.build(name))
    lp = compile_graph(ir, tf_in, tf_lab)
    need_w, need_wt = plan_publish_needs(lp)
    lay = ParamLayout.build(ir.param_shapes(), need_w, need_wt)
    dev = torch.device("cuda:0")
    master = MasterState(lay, spec, dev, n_mailboxes=1 if served else 0)
    w0 = GraphProgram(ir).init_weights(seed=1)
    master.load_weights(w0)
    if served:
        master.start_applier(lock, scope_sys=False, grid=32)
    worker = DeviceWorker(ir, tf_in, tf_lab, spec, master, acquire_lock=lock, pull_mode=pull_mode, shared=False)
    return ir, master, worker, w0, (tf_in, tf_lab, d, c, kind)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("lock", [False, True])
def test_compiled_step_tracks_oracle(name, lock):
    spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.001))
    ir, master, worker, w0, (tf_in, tf_lab, d, c, kind) = _setup(name, spec, lock)
    X, Y = _data(256, d, c, kind)
    eng = B200Engine(worker)
    eng.load_partition(X, Y)
    ps = ParameterServer(w0, spec, acquire_lock=lock)
    ref = TorchEngine(ir, tf_in, tf_lab, LocalTransport(ps))
    ref.load_partition(X, Y)
    rows = [slice(0, 64), slice(64, 128), np.arange(100, 164), slice(192, 256), slice(0, 64)]
    for r in rows:
        eng.train(r, pull=True)
        ref.train(r, pull=True)
    eng.finish()
    got, exp = master.weights(), ps.weights()
    for a, b, v in zip(got, exp, ir.trainable):
        # adam steps are +-lr per element; bf16 forward/backward may flip the sign of tiny gradients
        assert np.abs(a - b).max() < 5 * 0.001 * len(rows), v.name
        assert np.mean(np.abs(a - b)) < 0.35 * 0.001 * len(rows), v.name
    cnt = master.counters()
    assert cnt["pushes"] == len(rows) and cnt["lock"] == 0
    # loss of the last step agrees with the oracle's loss on the same batch before its update
    loss_gpu = worker.partition_loss(eng.X, eng.Y)
    loss_ref = GraphProgram(ir).loss(ref._feed(slice(0, 256)), got)
    assert abs(loss_gpu - loss_ref) < 2e-2 * max(1.0, abs(loss_ref))
    master.close()


def test_training_reduces_loss_and_predicts():
    spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.002))
    ir, master, worker, w0, (tf_in, tf_lab, d, c, kind) = _setup("simple_dnn", spec)
    rng = np.random.default_rng(3)
    centers = rng.normal(0, 1, (10, 784)).astype(np.float32)
    lab = rng.integers(0, 10, 3000)
    X = centers[lab] + 0.3 * rng.normal(0, 1, (3000, 784)).astype(np.float32)
    Y = np.eye(10, dtype=np.float32)[lab]
    eng = B200Engine(worker)
    eng.load_partition(X, Y)
    l0 = eng.partition_loss()
    for it in range(30):
        for r in range(0, 3000, 300):
            eng.train(slice(r, r + 300), pull=True)
    eng.finish()
    l1 = eng.partition_loss()
    assert l1 < 0.2 * l0, (l0, l1)
    # prediction through the forward-only plan (+ArgMax kernel)
    pred = worker.predict(X[:500], upto=2, post="ArgMax")
    assert (pred.astype(np.int64) == lab[:500]).mean() > 0.95
    master.close()


@pytest.mark.parametrize("pull_mode", ["copy", "direct"])
def test_graph_replay_equals_eager(pull_mode):
    spec = OptimizerSpec.from_tf_kwargs("gradient_descent", dict(learning_rate=0.05))
    outs = []
    for use_graphs in (False, True):
        ir, master, worker, w0, (tf_in, tf_lab, d, c, kind) = _setup("test_mlp", spec, pull_mode=pull_mode)
        worker.use_graphs = use_graphs
        X, Y = _data(128, d, c, kind, seed=5)
        eng = B200Engine(worker)
        eng.load_partition(X, Y)
        for _ in range(6):
            eng.train(slice(0, 128), pull=True)
        eng.finish()
        outs.append(master.weights())
        master.close()
    for a, b in zip(*outs):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("lock", [False, True])
def test_served_push_mailbox_applier_tracks_oracle(lock):
    """push = post to the mailbox + persistent applier kernel on the master GPU; same numbers as direct mode."""
    spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.001))
    ir, master, worker, w0, (tf_in, tf_lab, d, c, kind) = _setup("simple_dnn", spec, lock, served=True)
    assert worker.served and master.applier.alive()
    X, Y = _data(256, d, c, kind)
    eng = B200Engine(worker)
    eng.load_partition(X, Y)
    ps = ParameterServer(w0, spec, acquire_lock=lock)
    ref = TorchEngine(ir, tf_in, tf_lab, LocalTransport(ps))
    ref.load_partition(X, Y)
    rows = [slice(0, 64), slice(64, 128), slice(128, 192), slice(192, 256)] * 3
    for r in rows:
        eng.train(r, pull=True)          # each pull waits until the applier has consumed this worker's previous post
        ref.train(r, pull=True)
    eng.finish()
    import time
    t0 = time.time()
    while master.counters()["pushes"] < len(rows) and time.time() - t0 < 10:
        time.sleep(0.01)
    cnt = master.counters()
    assert cnt["pushes"] == len(rows) and cnt["lock"] == 0, cnt
    for a, b, v in zip(master.weights(), ps.weights(), ir.trainable):
        assert np.abs(a - b).max() < 5 * 0.001 * len(rows), v.name
        assert np.mean(np.abs(a - b)) < 0.35 * 0.001 * len(rows), v.name
    assert master.applier.alive()
    master.close()
    assert master.applier is None
