"""Sharded master (parallel/sharded.py): every piece exercised on ONE GPU (two shards + two workers on the same device,
system-scope protocol, peer stores instead of NVLS) so the driver's single-GPU box runs it; the multi-GPU / multicast
form of the same protocol is in test_gpu_multi.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from sparkflow_b200.graph.executor import GraphProgram
from sparkflow_b200.graph.ir import GraphIR
from sparkflow_b200.models import zoo
from sparkflow_b200.ops import native
from sparkflow_b200.ops.optimizers import OptimizerSpec
from sparkflow_b200.parallel.param_server import LocalTransport, ParameterServer
from sparkflow_b200.parallel.session import TrainingSession
from sparkflow_b200.parallel.worker import TorchEngine


def test_wgrad_epilogue_routes_tiles_to_the_owning_mailbox():
    """GEMM epilogue with a route: element (r, c) of the fp32 output must land in mailbox[owner(tile(r, c))] only."""
    C = native.cuda_ext()
    dev = torch.device("cuda:0")
    M, N, K = 200, 150, 64
    torch.manual_seed(0)
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, K, device=dev).to(torch.bfloat16)
    tiles_c = (N + 63) // 64
    n_tiles = ((M + 31) // 32) * tiles_c
    tile0, off = 5, 1024                                   # pretend the variable starts at push tile 5, flat offset 1024
    bounds = [0, tile0 + 4, tile0 + 11, tile0 + n_tiles]     # 3 shards
    mbs = [torch.zeros(off + M * N + 64, device=dev) for _ in range(3)]
    route = torch.frombuffer(bytearray(C.pack_route(3, bounds, [m.data_ptr() for m in mbs])), dtype=torch.uint8).to(dev)
    g = C.Gemm(dict(a=a.data_ptr(), lda=K, b=b.data_ptr(), ldb=K, M=M, N=N, K=K, route=route.data_ptr(), route_tile0=tile0,
                    route_tiles_c=tiles_c, route_off=off, ld_f32=N))
    g.launch(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    rows, cols = torch.meshgrid(torch.arange(M), torch.arange(N), indexing="ij")
    tile = tile0 + (rows // 32) * tiles_c + cols // 64
    owner = (tile >= bounds[1]).int() + (tile >= bounds[2]).int()
    for r in range(3):
        got = mbs[r][off:off + M * N].view(M, N).cpu()
        mine = owner == r
        assert mine.any()
        assert torch.allclose(got[mine], ref.cpu()[mine], rtol=2e-2, atol=2e-2)
        assert (got[~mine] == 0).all()                      # nothing leaks into a mailbox that does not own the tile
        assert (mbs[r][:off] == 0).all()


def _blobs(n, seed):
    rng = np.random.default_rng(seed)
    centers = np.random.default_rng(11).normal(0, 1, (10, 784)).astype(np.float32)
    lab = rng.integers(0, 10, n)
    return centers[lab] + 0.3 * rng.normal(0, 1, (n, 784)).astype(np.float32), np.eye(10, dtype=np.float32)[lab], lab


@pytest.mark.parametrize("lock", [False, True])
def test_one_worker_two_shards_tracks_the_host_parameter_server(lock):
    """Deterministic: ONE worker pushing through a master sharded in two (both shards on GPU 0) must follow the host
    parameter server step for step - mailbox routing, per-shard apply, publish, acknowledgement, snapshot pull."""
    spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.001))
    graph = zoo.build("simple_dnn")
    ir = GraphIR.from_metagraph(graph)
    w0 = GraphProgram(ir).init_weights(seed=1)
    X, Y, _ = _blobs(512, 0)
    sess = TrainingSession(graph, "x:0", "y:0", spec, acquire_lock=lock, engine="b200", initial_weights=w0, push_mode="sharded",
                           devices=[0, 0]).open()
    assert sess.master.n == 2 and not sess.master.multicast
    eng = sess.make_engine(torch.device("cuda:0"), lane=0)
    eng.load_partition(X, Y)
    ps = ParameterServer(w0, spec, acquire_lock=lock)
    ref = TorchEngine(ir, "x:0", "y:0", LocalTransport(ps))
    ref.load_partition(X, Y)
    rows = [slice(0, 128), slice(128, 256), np.arange(100, 228), slice(256, 384), slice(0, 128), slice(384, 512)]
    for r in rows:
        eng.train(r, pull=True)
        ref.train(r, pull=True)
    eng.finish()
    got, exp = sess.weights(), ps.weights()
    for a, b, v in zip(got, exp, ir.trainable):
        assert np.abs(a - b).max() < 5 * 0.001 * len(rows), v.name
        assert np.mean(np.abs(a - b)) < 0.45 * 0.001 * len(rows), v.name
    c = sess.counters()
    assert c["pushes"] == len(rows) and c["shards"] == 2
    names = eng.w.last_plan_names()
    # lock mode: seqlock snapshot kernel; Hogwild: no pull kernel at all (the wait is fused into the first GEMM)
    assert (("sync_pull@1" in names or "sync_pull" in names) if lock else not any(n.startswith("sync_pull") or n.startswith("pull") for n in names)), names
    assert "post_flags" in names and "push" not in names and "post" not in names
    loss_gpu = eng.partition_loss()
    loss_ref = GraphProgram(ir).loss(ref._feed(slice(0, 512)), got)
    assert abs(loss_gpu - loss_ref) < 2e-2 * max(1.0, abs(loss_ref))
    sess.close()


def _train_with_watchdog(sess, parts, stall_s=6.0):
    """train_partitions on a thread; if no push is applied for `stall_s` the protocol words of every shard / worker are
    dumped into the failure message (the bounded device waits would only trap - and kill the context - after 20 s)."""
    import json
    import threading
    import time

    sess.open()
    done, err = threading.Event(), []

    def run():
        try:
            sess.train_partitions(parts)
        except BaseException as exc:  # noqa: BLE001
            err.append(exc)
        done.set()

    threading.Thread(target=run, daemon=True).start()
    last, t_last = -1, time.time()
    while not done.wait(0.25):
        st = sess.master.debug_state()
        cur = sum(st[f"shard{r}.ctrl"][3] for r in range(sess.master.n))
        if cur != last:
            last, t_last = cur, time.time()
        elif time.time() - t_last > stall_s:
            pytest.fail("sharded run stalled: " + json.dumps(st))
    if err:
        raise err[0]


@pytest.mark.parametrize("lock", [False, True])
@pytest.mark.parametrize("model", ["simple_dnn", "cnn"])
def test_two_workers_two_shards_on_one_gpu_train(lock, model):
    """Two worker threads + two shard appliers sharing GPU 0: asynchronous pushes from both workers are all applied
    (one optimizer step per push per shard) and the model learns."""
    spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.002))
    graph = zoo.build(model)
    parts = [_blobs(1500, 20 + p) for p in range(2)]
    iters = 6 if model == "simple_dnn" else 3
    sess = TrainingSession(graph, "x:0", "y:0", spec, acquire_lock=lock, iters=iters, mini_batch=300, shuffle=False, engine="b200", seed=3,
                           push_mode="sharded", devices=[0, 0])
    _train_with_watchdog(sess, [(x, y) for x, y, _ in parts])
    w = sess.weights()
    c = sess.counters()
    assert c["pushes"] == 2 * iters * 5, c
    prog = GraphProgram(GraphIR.from_metagraph(graph))
    X = np.concatenate([p[0] for p in parts])
    L = np.concatenate([p[2] for p in parts])
    acc = float((prog.forward("out:0", {"x:0": X}, w).numpy() == L).mean())
    assert acc > (0.9 if model == "simple_dnn" else 0.5), acc
    sess.close()
