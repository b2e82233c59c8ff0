"""Device timeline of the END-TO-END loop (StepDriver: H2D + graph replay + loss hand-off per step), from the
in-kernel %globaltimer tracer: per step, first-kernel start, last-kernel end and the idle gap to the next step."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from sparkflow_b200.models import zoo
from sparkflow_b200.ops.optimizers import OptimizerSpec
from sparkflow_b200.parallel.session import TrainingSession
from sparkflow_b200.utils.trace import DeviceTrace

lock = "--lock" in sys.argv
B = 300
spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.001))
sess = TrainingSession(zoo.build("simple_dnn"), "x:0", "y:0", spec, acquire_lock=lock, engine="b200", seed=0, devices=[0]).open()
eng = sess.make_engine(torch.device("cuda", 0))
rng = np.random.default_rng(0)
rows = 50100
X = rng.random((rows, 784), dtype=np.float32)
Y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, rows)]
eng.load_partition(X, Y)
nb = rows // B
eng.train_contiguous([(k % nb) * B for k in range(40)], B, pull=True)
eng.finish()
N = 40
with DeviceTrace(1 << 15) as tr:
    eng.train_contiguous([((40 + k) % nb) * B for k in range(N)], B, pull=True)
    eng.finish()
rec = [r for r in tr.records() if int(r["kid"]) < 100]
rec.sort(key=lambda r: int(r["t0"]))
# a step starts with the pull / cast kernels and ends with push: split on push records (kid 6)
steps, cur = [], None
for r in rec:
    t0, t2, kid = int(r["t0"]), int(r["t2"]), int(r["kid"])
    if cur is None:
        cur = dict(start=t0, end=t2, push_seen=False)
    elif cur["push_seen"] and kid != 6:
        steps.append(cur)
        cur = dict(start=t0, end=t2, push_seen=False)
    cur["end"] = max(cur["end"], t2)
    if kid == 6:
        cur["push_seen"] = True
if cur:
    steps.append(cur)
dur = np.array([(s["end"] - s["start"]) / 1e3 for s in steps])
gap = np.array([(steps[i + 1]["start"] - steps[i]["end"]) / 1e3 for i in range(len(steps) - 1)])
period = np.array([(steps[i + 1]["start"] - steps[i]["start"]) / 1e3 for i in range(len(steps) - 1)])
out = dict(steps=len(steps), busy_us_median=float(np.median(dur)), gap_us_median=float(np.median(gap)), period_us_median=float(np.median(period)),
           busy_us=[round(float(v), 2) for v in dur[:16]], gap_us=[round(float(v), 2) for v in gap[:16]],
           host_us_per_step={n: v / eng._driver.steps() / 1e3 for n, v in
                             zip(("wait_slot", "h2d_enqueue", "event_handoff", "graph_launch", "record"), eng._driver.host_ns())})
print(json.dumps(out))
json.dump(out, open("gpurun_out/trace_e2e%s.json" % ("_lock" if lock else ""), "w"))
sess.close()
