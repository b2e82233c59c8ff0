"""Runs the flagship step eagerly (no CUDA graph) so ncu can attribute time per kernel."""
import sys

import torch

sys.path.insert(0, ".")
from sparkflow_b200.models import zoo
from sparkflow_b200.ops.optimizers import OptimizerSpec
from sparkflow_b200.parallel.session import TrainingSession

steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 6
model = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--model=")), "simple_dnn")
spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.001))
sess = TrainingSession(zoo.build(model), "x:0", "y:0", spec, acquire_lock="--lock" in sys.argv, engine="b200", seed=0, devices=[0]).open()
eng = sess.make_engine(torch.device("cuda", 0))
w = eng.w
w.use_graphs = False
plan, bufs = w.build_plan(300, 0)
bufs.x_stage.uniform_()
bufs.y_stage.zero_()
bufs.y_stage[:, 3] = 1
# warm steps outside the capture range, ONE step inside it (run ncu with --profile-from-start off)
for _ in range(steps - 1):
    plan.run(w.stream.cuda_stream)
w.stream.synchronize()
torch.cuda.cudart().cudaProfilerStart()
plan.run(w.stream.cuda_stream)
w.stream.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("kernels per step:", plan.names())
sess.close()
