"""Probe of the served-push protocol: start the applier, post one gradient, watch the flag / sync words."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from sparkflow_b200.ops import native
from sparkflow_b200.ops.layout import ParamLayout
from sparkflow_b200.ops.optimizers import OptimizerSpec
from sparkflow_b200.parallel.device_engine import MasterState, _view

C = native.cuda_ext()
dev = torch.device("cuda", 0)
lay = ParamLayout.build([("a/kernel", (784, 256)), ("a/bias", (256,)), ("b/kernel", (256, 10)), ("b/bias", (10,))])
spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.01))
m = MasterState(lay, spec, dev, n_mailboxes=2)
m.load_weights([np.ones(s.shape, np.float32) for s in lay.segments])
print("layout", m.ml)
flags = _view(m.base + m.ml.flags, 2 * C.MB_WORDS * 4, torch.int32, dev)
sync = _view(m.base + m.ml.applier_sync, 32, torch.int32, dev)
m.start_applier(acquire_lock="--lock" in sys.argv, scope_sys=False, grid=32)
time.sleep(0.2)
print("alive", m.applier.alive(), "flags", flags[[0, 16, 32, 48]].tolist(), "sync", sync[:8].tolist())
print("launches so far", m.applier.launches())
st = torch.cuda.Stream()
grads = torch.full((lay.total,), 0.5, device=dev)
lsync = torch.zeros(8, dtype=torch.int32, device=dev)
loss = torch.zeros(2, device=dev)
for k in range(3):
    args = dict(grad=native.ptr(grads), mailbox=m.mailbox_ptr(1), flags=m.flags_ptr(1), loss_acc=native.ptr(loss), loss_out=native.ptr(loss) + 4,
                n=lay.total)
    with torch.cuda.stream(st):
        grads.fill_(0.5)
        C.post(args, native.ptr(lsync), 0, st.cuda_stream)
    t0 = time.time()
    while time.time() - t0 < 1.5:
        f = flags[[0, 16, 32, 48]].tolist()
        if f[3] == k + 1:
            break
        time.sleep(0.01)
    print(f"post {k}: flags(posted0, applied0, posted1, applied1)={flags[[0, 16, 32, 48]].tolist()} sync={sync[:4].tolist()} "
          f"lsync={lsync.tolist()} counters={m.counters()} alive={m.applier.alive()} err={native.describe_device_error()}")
print("p[0..3]", m.p[:3].tolist(), "slots", m.slots[0][:2].tolist(), m.slots[1][:2].tolist())
m.close()
print("closed ok")
