"""Runs the flagship step eagerly (no CUDA graph) so ncu can attribute time per kernel."""
import sys

import torch

sys.path.insert(0, ".")
from sparkflow_b200.models import zoo
from sparkflow_b200.ops.optimizers import OptimizerSpec
from sparkflow_b200.parallel.session import TrainingSession

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.001))
sess = TrainingSession(zoo.build("simple_dnn"), "x:0", "y:0", spec, acquire_lock="--lock" in sys.argv, engine="b200", seed=0, devices=[0]).open()
eng = sess.make_engine(torch.device("cuda", 0))
w = eng.w
w.use_graphs = False
plan, bufs = w.build_plan(300, 0)
bufs.x_stage.uniform_()
bufs.y_stage.zero_()
bufs.y_stage[:, 3] = 1
for _ in range(steps):
    plan.run(w.stream.cuda_stream)
w.stream.synchronize()
print("kernels per step:", plan.names())
sess.close()
