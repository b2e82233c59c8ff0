#!/usr/bin/env python
"""After a few sharded pushes compare every replica's published bf16 W / W^T (+ fp32 tail) with the fp32 master state."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sparkflow_b200.graph.executor import GraphProgram
from sparkflow_b200.graph.ir import GraphIR
from sparkflow_b200.models import zoo
from sparkflow_b200.ops.optimizers import OptimizerSpec
from sparkflow_b200.parallel.session import TrainingSession

lock = sys.argv[1] == "lock"
graph = zoo.build("simple_dnn")
ir = GraphIR.from_metagraph(graph)
w0 = GraphProgram(ir).init_weights(seed=1)
rng = np.random.default_rng(0)
X = rng.random((512, 784), dtype=np.float32); Y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 512)]
sess = TrainingSession(graph, "x:0", "y:0", OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.01)), acquire_lock=lock, engine="b200",
                       initial_weights=w0, push_mode="sharded", devices=[0, 0]).open()
eng = sess.make_engine(torch.device("cuda:0"), lane=0)
eng.load_partition(X, Y)
for k in range(5):
    eng.train(slice(0, 128), pull=True)
eng.finish()
prog = GraphProgram(ir)
for rep in range(3):
    lg = eng.partition_loss()
    lr_ = prog.loss({"x:0": X, "y:0": Y}, sess.weights())
    print("partition_loss gpu", lg, "oracle", lr_)
m, lay = sess.master, sess.layout
flat = lay.flatten(sess.weights())
ref = torch.from_numpy(lay.publish_reference(flat)).to(torch.bfloat16).float().numpy()
ref0 = torch.from_numpy(lay.publish_reference(lay.flatten(w0))).to(torch.bfloat16).float().numpy()
print("counters", sess.counters(), "stamps", m.debug_state()["replica0.stamps(begin,end)"])
for rep in range(0):
    for slot, nm in ((0, "shadow"), (1, "shadow1")):
        got = m.view(nm, rep).float().cpu().numpy()
        for s in lay.segments:
            for kind, off, n in (("W", s.w_off, s.rows * s.w_ld), ("WT", s.wt_off, s.cols * s.wt_ld)):
                if off < 0: continue
                d = np.abs(got[off:off + n] - ref[off:off + n]); d0 = np.abs(got[off:off + n] - ref0[off:off + n])
                print(f"replica {rep} slot {slot} {s.name:16s} {kind:2s} max|pub-state|={d.max():.4f} frac_bad={(d > 0.02).mean():.3f}   (vs initial: max={d0.max():.4f})")
sess.close()
