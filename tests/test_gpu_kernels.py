"""Numerics of every sm_100a kernel against a plain PyTorch fp32 reference of the same op."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from sparkflow_b200.ops import native
from sparkflow_b200.ops.layout import ParamLayout, round_up
from sparkflow_b200.ops.optimizers import OPT_IDS, OptimizerSpec, apply_update, init_slots

ACT = {None: 0, "relu": 1, "sigmoid": 2, "tanh": 3}


def _act(x, a):
    return {None: lambda v: v, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh}[a](x)


def _dact(a_out, a):
    if a == "relu":
        return (a_out > 0).float()
    if a == "sigmoid":
        return a_out * (1 - a_out)
    if a == "tanh":
        return 1 - a_out * a_out
    return torch.ones_like(a_out)


@pytest.fixture(scope="module")
def C():
    return native.cuda_ext()


def _bf16_pad(x, ld):
    out = torch.zeros(x.shape[0], ld, dtype=torch.bfloat16, device="cuda")
    out[:, : x.shape[1]] = x.to(torch.bfloat16)
    return out


def _run_gemm(C, M, N, K, bn=0, split_k=1, bias=False, act=None, aux_act=None, colsum=False, want_bf16=True,
              want_t=True, accumulate=False, seed=0, pair=-1, pair_ctas=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(M, K, device="cuda", generator=g) * 0.5
    b = torch.randn(N, K, device="cuda", generator=g) * 0.5
    lda, ldb = round_up(K, 8), round_up(K, 8)
    a16, b16 = _bf16_pad(a, lda), _bf16_pad(b, ldb)
    ref = a16[:, :K].float() @ b16[:, :K].float().t()
    ldn = round_up(N, 8)
    ldt = round_up(M, 8)
    out_f32 = torch.full((M, N), 7.0 if accumulate else float("nan"), device="cuda")
    out_bf16 = torch.zeros(M, ldn, dtype=torch.bfloat16, device="cuda") if want_bf16 else None
    outT = torch.zeros(N, ldt, dtype=torch.bfloat16, device="cuda") if want_t else None
    bias_t = torch.randn(N, device="cuda", generator=g) if bias else None
    aux = None
    if aux_act is not None:
        aux = _bf16_pad(_act(torch.randn(M, N, device="cuda", generator=g), aux_act), ldn)
    cs = torch.zeros(N, device="cuda") if colsum else None
    d = dict(a=native.ptr(a16), b=native.ptr(b16), M=M, N=N, K=K, lda=lda, ldb=ldb, bn=bn, split_k=split_k,
             out_f32=native.ptr(out_f32), ld_f32=N, out_bf16=native.ptr(out_bf16), ld_bf16=ldn,
             outT_bf16=native.ptr(outT), ld_t=ldt, bias=native.ptr(bias_t), act=ACT[act], aux=native.ptr(aux),
             ld_aux=ldn, aux_act=ACT[aux_act], colsum=native.ptr(cs), alpha=1.0, accumulate=int(accumulate), pair=pair,
             pair_ctas=pair_ctas)
    gm = C.Gemm(d)
    gm.launch(native.current_stream())
    torch.cuda.synchronize()
    assert C.read_error_code() == 0
    exp = ref
    if bias:
        exp = exp + bias_t
    exp = _act(exp, act)
    if aux is not None:
        exp = exp * _dact(aux[:, :N].float(), aux_act)
    tol = dict(rtol=2e-2, atol=2e-2 * max(1.0, K ** 0.5 * 0.25))
    got = out_f32 - 7.0 if accumulate else out_f32
    torch.testing.assert_close(got, exp, **tol)
    if want_bf16:
        torch.testing.assert_close(out_bf16[:, :N].float(), exp, rtol=3e-2, atol=tol["atol"] + 0.05)
        assert torch.all(out_bf16[:, N:] == 0)
    if want_t:
        torch.testing.assert_close(outT[:, :M].float().t(), exp, rtol=3e-2, atol=tol["atol"] + 0.05)
    if colsum:
        torch.testing.assert_close(cs, exp.sum(0), rtol=2e-2, atol=tol["atol"] * 4)
    return gm


@pytest.mark.parametrize("M,N,K", [(128, 32, 64), (128, 64, 128), (300, 256, 784), (300, 10, 256),
                                     (784, 256, 300), (256, 10, 300), (1000, 520, 1000)])
def test_gemm_plain(C, M, N, K):
    _run_gemm(C, M, N, K)


@pytest.mark.parametrize("bn", [32, 64, 128, 256])
def test_gemm_tile_widths(C, bn):
    _run_gemm(C, 384, 512, 320, bn=bn)


def test_gemm_bias_relu(C):
    _run_gemm(C, 300, 256, 784, bias=True, act="relu")


def test_gemm_bias_sigmoid_tanh(C):
    _run_gemm(C, 257, 130, 200, bias=True, act="sigmoid")
    _run_gemm(C, 257, 130, 200, bias=True, act="tanh")


def test_gemm_dgrad_epilogue(C):
    _run_gemm(C, 300, 256, 10, aux_act="relu", colsum=True)
    _run_gemm(C, 300, 256, 256, aux_act="sigmoid", colsum=True)


def test_gemm_split_k_accumulate(C):
    gm = _run_gemm(C, 32, 64, 4096, split_k=8, want_bf16=False, want_t=False, accumulate=True)
    assert gm.split_k == 8


def test_gemm_large(C):
    _run_gemm(C, 2048, 2048, 2048, want_t=False)


@pytest.mark.parametrize("M,N,K,bn,split_k", [(784, 256, 300, 0, 1), (256, 256, 300, 64, 1), (256, 10, 300, 0, 1), (1600, 10, 300, 0, 1),
                                              (288, 64, 30000, 0, 8), (25, 32, 4000, 64, 1), (4096, 4096, 300, 128, 1), (130, 70, 17, 0, 1)])
def test_gemm_mn_major_operands(C, M, N, K, bn, split_k):
    """D = a^T . b with a [K, M] and b [K, N] row-major (both operands MN-major for the tensor core): the weight-gradient
    form dW = activations^T . dz consumed straight from the row-major activation / dz buffers, no transposed copies."""
    g = torch.Generator(device="cuda").manual_seed(7)
    lda, ldb = round_up(M, 8), round_up(N, 8)
    a = torch.zeros(K, lda, dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(K, ldb, dtype=torch.bfloat16, device="cuda")
    a[:, :M] = (torch.randn(K, M, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    b[:, :N] = (torch.randn(K, N, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    ref = a[:, :M].float().t() @ b[:, :N].float()
    out = torch.zeros(M, N, device="cuda") if split_k > 1 else torch.full((M, N), float("nan"), device="cuda")
    cs = torch.zeros(N, device="cuda")
    gm = C.Gemm(dict(a=native.ptr(a), lda=lda, b=native.ptr(b), ldb=ldb, M=M, N=N, K=K, mn_major=1, bn=bn, split_k=split_k,
                     out_f32=native.ptr(out), ld_f32=N, colsum=native.ptr(cs) if split_k == 1 else 0, accumulate=1 if split_k > 1 else 0))
    assert gm.mn_major() == 1 and gm.bn in (64, 128)
    gm.launch(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    scale = max(1.0, ref.abs().max().item())
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-2 * scale * (K / 300) ** 0.5 + 1e-3
    if split_k == 1:
        assert torch.allclose(cs, ref.sum(0), rtol=2e-2, atol=2e-2 * scale)


@pytest.mark.parametrize("M,N,K,ctas", [(512, 512, 256, 0), (1000, 520, 1000, 0), (256, 256, 64, 0), (2048, 2048, 512, 8),
                                          (1536, 1280, 320, 6), (4096, 4096, 1024, 0)])
def test_gemm_pair_kernel(C, M, N, K, ctas):
    """persistent 2-CTA (cta_group::2) kernel: tails, more tiles than CTA pairs (double-buffered TMEM phases)"""
    gm = _run_gemm(C, M, N, K, pair=1, pair_ctas=ctas, want_t=(M * N <= 1 << 22))
    assert gm.pair == 1


def test_gemm_pair_kernel_epilogues(C):
    gm = _run_gemm(C, 768, 512, 512, pair=1, bias=True, act="relu")
    assert gm.pair == 1
    _run_gemm(C, 768, 512, 512, pair=1, aux_act="sigmoid", colsum=True)
    _run_gemm(C, 512, 768, 256, pair=1, want_bf16=False, want_t=False, accumulate=True)


def test_cast_transpose(C):
    x = torch.randn(300, 784, device="cuda")
    ld, ldt = round_up(784, 8), round_up(300, 8)
    out = torch.full((300, ld), 9.0, dtype=torch.bfloat16, device="cuda")
    outT = torch.full((784, ldt), 9.0, dtype=torch.bfloat16, device="cuda")
    C.cast_transpose(native.ptr(x), 784, 0, native.ptr(out), ld, native.ptr(outT), ldt, 300, 784,
                     native.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(out[:, :784], x.to(torch.bfloat16))
    assert torch.equal(outT[:, :300], x.to(torch.bfloat16).t())
    assert torch.all(outT[:, 300:] == 0)
    idx = torch.randperm(300, device="cuda", dtype=torch.int32)[:128].contiguous()
    out2 = torch.zeros(128, ld, dtype=torch.bfloat16, device="cuda")
    C.cast_transpose(native.ptr(x), 784, native.ptr(idx), native.ptr(out2), ld, 0, 0, 128, 784,
                     native.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(out2[:, :784], x[idx.long()].to(torch.bfloat16))


@pytest.mark.parametrize("B,Cc", [(300, 10), (64, 1000), (7, 3)])
def test_softmax_xent(C, B, Cc):
    z = torch.randn(B, Cc, device="cuda") * 3
    y = torch.nn.functional.one_hot(torch.randint(0, Cc, (B,), device="cuda"), Cc).float()
    ld, ldt = round_up(Cc, 8), round_up(B, 8)
    loss = torch.zeros(1, device="cuda")
    dz = torch.full((B, ld), 5.0, dtype=torch.bfloat16, device="cuda")
    dzT = torch.zeros(Cc, ldt, dtype=torch.bfloat16, device="cuda")
    db = torch.zeros(Cc, device="cuda")
    C.softmax_xent(native.ptr(z), Cc, native.ptr(y), Cc, native.ptr(loss), native.ptr(dz), ld, native.ptr(dzT), ldt,
                   native.ptr(db), B, Cc, native.current_stream())
    torch.cuda.synchronize()
    zr = z.clone().requires_grad_(True)
    ref = -(y * torch.log_softmax(zr, 1)).sum(1).mean()
    ref.backward()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dz[:, :Cc].float(), zr.grad, rtol=2e-2, atol=1e-4)
    torch.testing.assert_close(dzT[:, :B].float().t(), zr.grad, rtol=2e-2, atol=1e-4)
    assert torch.all(dz[:, Cc:] == 0)
    torch.testing.assert_close(db, zr.grad.sum(0), rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("act", [None, "sigmoid"])
def test_mse(C, act):
    B, Cc = 256, 784
    pre = torch.randn(B, Cc, device="cuda", requires_grad=True)
    out = _act(pre, act)
    tgt = torch.rand(B, Cc, device="cuda")
    ref = ((out - tgt) ** 2).mean()
    ref.backward()
    ld, ldt = round_up(Cc, 8), round_up(B, 8)
    loss = torch.zeros(1, device="cuda")
    dz = torch.zeros(B, ld, dtype=torch.bfloat16, device="cuda")
    dzT = torch.zeros(Cc, ldt, dtype=torch.bfloat16, device="cuda")
    db = torch.zeros(Cc, device="cuda")
    o = out.detach().contiguous()
    C.mse_loss(native.ptr(o), Cc, native.ptr(tgt), Cc, ACT[act], native.ptr(loss), native.ptr(dz), ld,
               native.ptr(dzT), ldt, native.ptr(db), B, Cc, native.current_stream())
    torch.cuda.synchronize()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(dz[:, :Cc].float(), pre.grad, rtol=2e-2, atol=1e-7)
    torch.testing.assert_close(dzT[:, :B].float().t(), pre.grad, rtol=2e-2, atol=1e-7)
    torch.testing.assert_close(db, pre.grad.sum(0), rtol=1e-3, atol=1e-7)


def test_argmax(C):
    z = torch.randn(333, 10, device="cuda")
    out = torch.zeros(333, device="cuda")
    C.argmax_rows(native.ptr(z), 10, native.ptr(out), 333, 10, native.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(out.long(), z.argmax(1))


SHAPES = [("dense/kernel", (784, 256)), ("dense/bias", (256,)), ("dense_1/kernel", (256, 256)),
          ("dense_1/bias", (256,)), ("dense_2/kernel", (256, 10)), ("dense_2/bias", (10,))]


def _push_setup(C, spec, lock_mode=0):
    lay = ParamLayout.build(SHAPES)
    dev = "cuda"
    state = torch.zeros(lay.total, 4, device=dev)
    mask = torch.from_numpy(lay.valid_mask()).to(dev)
    state[mask, 0] = torch.randn(int(mask.sum()), device=dev) * 0.1
    for i in range(spec.num_slots):
        state[:, 1 + i] = spec.slot_init(i)
    p = state[:, 0]
    slots = [state[:, 1 + i] for i in range(spec.num_slots)] + [None] * (3 - spec.num_slots)
    ctrl = torch.zeros(C.CTRL_WORDS, dtype=torch.int32, device=dev)
    shadow = torch.zeros(lay.shadow_total, dtype=torch.bfloat16, device=dev)
    grad = torch.zeros(lay.total, device=dev)
    segs = torch.frombuffer(bytearray(C.pack_segs(lay.seg_rows())), dtype=torch.uint8).to(dev)
    tmap = torch.from_numpy(lay.tile_map()).to(dev)
    lsync = torch.zeros(8, dtype=torch.int32, device=dev)
    loss_acc = torch.zeros(1, device=dev)
    loss_out = torch.zeros(1, device=dev)
    args = dict(state=native.ptr(state), ctrl=native.ptr(ctrl), shadow_dst=[native.ptr(shadow)], grad=native.ptr(grad),
                loss_acc=native.ptr(loss_acc), loss_out=native.ptr(loss_out), segs=native.ptr(segs),
                tile_map=native.ptr(tmap), num_tiles=int(tmap.shape[0]), optimizer=spec.opt_id,
                lock_mode=lock_mode, grad_scale=1.0, hyper=spec.native_hyper())
    keep = dict(lay=lay, state=state, p=p, slots=slots, ctrl=ctrl, shadow=shadow, grad=grad, segs=segs, tmap=tmap, lsync=lsync,
                loss_acc=loss_acc, loss_out=loss_out, mask=mask)
    return args, keep


@pytest.mark.parametrize("name,kw", [
    ("gradient_descent", dict(learning_rate=0.1)),
    ("momentum", dict(learning_rate=0.1, momentum=0.9)),
    ("momentum", dict(learning_rate=0.1, momentum=0.9, use_nesterov=True)),
    ("adam", dict(learning_rate=0.01, beta1=0.85, beta2=0.98, epsilon=1e-8)),
    ("rmsprop", dict(learning_rate=0.05, decay=0.95, momentum=0.1, epsilon=1e-10)),
    ("rmsprop", dict(learning_rate=0.05, decay=0.95, momentum=0.1, epsilon=1e-6, centered=True)),
    ("adagrad", dict(learning_rate=0.1, initial_accumulator_value=0.1)),
    ("adadelta", dict(learning_rate=1.0, rho=0.95, epsilon=1e-6)),
    ("adagrad_da", dict(learning_rate=0.1, l1_regularization_strength=0.001, l2_regularization_strength=0.01)),
    ("ftrl", dict(learning_rate=0.1, l1_regularization_strength=0.001, l2_regularization_strength=0.01)),
    ("proximal_adagrad", dict(learning_rate=0.1, l1_regularization_strength=0.001, l2_regularization_strength=0.01)),
    ("proximal_gradient_descent", dict(learning_rate=0.1, l1_regularization_strength=0.001,
                                       l2_regularization_strength=0.01)),
])
@pytest.mark.parametrize("lock_mode", [0, 1])
def test_push_matches_reference_optimizer(C, name, kw, lock_mode):
    spec = OptimizerSpec.from_tf_kwargs(name, kw)
    args, k = _push_setup(C, spec, lock_mode)
    p_ref = k["p"].clone().contiguous()
    s_ref = [s.clone().contiguous() for s in k["slots"] if s is not None]
    for step in range(1, 4):
        g = torch.zeros_like(k["grad"])
        g[k["mask"]] = torch.randn(int(k["mask"].sum()), device="cuda") * 0.05
        k["grad"].copy_(g)
        k["loss_acc"].fill_(float(step))
        C.push(args, native.ptr(k["lsync"]), 0, native.current_stream())
        torch.cuda.synchronize()
        assert C.read_error_code() == 0
        apply_update(spec, p_ref, g, s_ref, step)
        torch.testing.assert_close(k["p"][k["mask"]], p_ref[k["mask"]], rtol=2e-4, atol=2e-6)
        for a, b in zip([s for s in k["slots"] if s is not None], s_ref):
            torch.testing.assert_close(a[k["mask"]], b[k["mask"]], rtol=2e-4, atol=1e-6)
        assert torch.all(k["grad"] == 0), "push must consume (zero) the gradient buffer"
        assert float(k["loss_out"][0]) == float(step) and float(k["loss_acc"][0]) == 0.0
        ctrl = k["ctrl"].cpu().numpy()
        assert ctrl[0] == 0 and ctrl[1] == step and ctrl[2] == step and ctrl[3] == step
    exp = torch.from_numpy(k["lay"].publish_reference(k["p"].contiguous().cpu().numpy())).to(torch.bfloat16)
    assert torch.equal(k["shadow"].cpu(), exp)


def test_push_drop_fault_injection(C):
    spec = OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.01))
    args, k = _push_setup(C, spec)
    args["drop"] = 1
    before = k["p"].clone().contiguous()
    k["grad"][k["mask"]] = 1.0
    C.push(args, native.ptr(k["lsync"]), 0, native.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(before, k["p"].contiguous()) and torch.all(k["grad"] == 0)
    assert int(k["ctrl"][5]) == 1 and int(k["ctrl"][3]) == 0


@pytest.mark.parametrize("lock_mode", [0, 1])
def test_pull_copies_publish_buffer(C, lock_mode):
    n16, n32 = 64 * 1000, 4 * 300
    src = torch.randn(n16, device="cuda").to(torch.bfloat16)
    dst = torch.zeros_like(src)
    sstate = torch.randn(n32, 4, device="cuda")
    s32 = sstate[:, 0].contiguous()
    d32 = torch.zeros_like(s32)
    ctrl = torch.zeros(C.CTRL_WORDS, dtype=torch.int32, device="cuda")
    ctrl[1] = 41
    seen = torch.zeros(1, dtype=torch.int32, device="cuda")
    lsync = torch.zeros(8, dtype=torch.int32, device="cuda")
    d = dict(src=native.ptr(src), dst=native.ptr(dst), src_state=native.ptr(sstate), dst_f32=native.ptr(d32),
             n_bf16=n16, n_f32=n32, ctrl=native.ptr(ctrl), seen_version=native.ptr(seen), lock_mode=lock_mode)
    for _ in range(2):
        C.pull(d, native.ptr(lsync), 0, native.current_stream())
    torch.cuda.synchronize()
    assert torch.equal(src, dst) and torch.equal(s32, d32) and int(seen[0]) == 41 and int(ctrl[0]) == 0


def test_device_rwlock_word(C):
    ctrl = torch.zeros(C.CTRL_WORDS, dtype=torch.int32, device="cuda")
    st = native.current_stream()
    C.lock_op(native.ptr(ctrl), 0, st); C.lock_op(native.ptr(ctrl), 0, st)
    torch.cuda.synchronize()
    assert int(ctrl[0]) == 2
    C.lock_op(native.ptr(ctrl), 1, st); C.lock_op(native.ptr(ctrl), 1, st)
    C.lock_op(native.ptr(ctrl), 2, st)
    torch.cuda.synchronize()
    assert int(ctrl[0]) == 1 << 16
    C.lock_op(native.ptr(ctrl), 3, st)
    torch.cuda.synchronize()
    assert int(ctrl[0]) == 0


def test_plan_capture_replay(C):
    x = torch.randn(64, 32, device="cuda")
    out = torch.zeros(64, 32, dtype=torch.bfloat16, device="cuda")
    plan = C.Plan()
    plan.add_cast_transpose(native.ptr(x), 32, 0, native.ptr(out), 32, 0, 0, 64, 32)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        plan.run(st.cuda_stream)
        st.synchronize()
        plan.capture(st.cuda_stream)
        out.zero_()
        plan.replay(st.cuda_stream)
    st.synchronize()
    assert torch.equal(out, x.to(torch.bfloat16))


def test_gemm_fused_softmax_xent_epilogue(C):
    """Last dense layer + softmax-CE + gradient in one kernel (dz, dz^T, dbias, loss)."""
    M, N, K = 300, 10, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    y = torch.nn.functional.one_hot(torch.randint(0, N, (M,), device="cuda"), N).float()
    ldn, ldt = round_up(N, 8), round_up(M, 8)
    dz = torch.full((M, ldn), 3.0, dtype=torch.bfloat16, device="cuda")
    dzT = torch.zeros(N, ldt, dtype=torch.bfloat16, device="cuda")
    db = torch.zeros(N, device="cuda")
    loss = torch.zeros(1, device="cuda")
    gm = C.Gemm(dict(a=native.ptr(a), b=native.ptr(w), M=M, N=N, K=K, lda=K, ldb=K, bias=native.ptr(bias), loss_mode=1,
                     target=native.ptr(y), ld_target=N, loss=native.ptr(loss), out_bf16=native.ptr(dz), ld_bf16=ldn,
                     outT_bf16=native.ptr(dzT), ld_t=ldt, colsum=native.ptr(db)))
    gm.launch(native.current_stream())
    torch.cuda.synchronize()
    z = (a.float() @ w.float().t() + bias).requires_grad_(True)
    ref = -(y * torch.log_softmax(z, 1)).sum(1).mean()
    ref.backward()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(dz[:, :N].float(), z.grad, rtol=2e-2, atol=2e-5)
    torch.testing.assert_close(dzT[:, :M].float().t(), z.grad, rtol=2e-2, atol=2e-5)
    assert torch.all(dz[:, N:] == 0)
    torch.testing.assert_close(db, z.grad.sum(0), rtol=1e-2, atol=1e-5)


@pytest.mark.parametrize("act", [None, "sigmoid"])
def test_gemm_fused_mse_epilogue(C, act):
    M, N, K = 256, 784, 256
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.2).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.2).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    tgt = torch.rand(M, N, device="cuda", generator=g)
    ldn, ldt = round_up(N, 8), round_up(M, 8)
    dz = torch.zeros(M, ldn, dtype=torch.bfloat16, device="cuda")
    dzT = torch.zeros(N, ldt, dtype=torch.bfloat16, device="cuda")
    db = torch.zeros(N, device="cuda")
    loss = torch.zeros(1, device="cuda")
    gm = C.Gemm(dict(a=native.ptr(a), b=native.ptr(w), M=M, N=N, K=K, lda=K, ldb=K, bias=native.ptr(bias), act=ACT[act],
                     loss_mode=2, target=native.ptr(tgt), ld_target=N, loss=native.ptr(loss), out_bf16=native.ptr(dz),
                     ld_bf16=ldn, outT_bf16=native.ptr(dzT), ld_t=ldt, colsum=native.ptr(db)))
    gm.launch(native.current_stream())
    torch.cuda.synchronize()
    z = (a.float() @ w.float().t() + bias).requires_grad_(True)
    ref = ((_act(z, act) - tgt) ** 2).mean()
    ref.backward()
    torch.testing.assert_close(loss[0], ref.detach(), rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(dz[:, :N].float(), z.grad, rtol=2e-2, atol=1e-7)
    torch.testing.assert_close(dzT[:, :M].float().t(), z.grad, rtol=2e-2, atol=1e-7)
    torch.testing.assert_close(db, z.grad.sum(0), rtol=1e-2, atol=1e-6)


def test_plan_branches_fork_join(C):
    x = torch.randn(64, 32, device="cuda")
    a = torch.zeros(64, 32, dtype=torch.bfloat16, device="cuda")
    b = torch.zeros(64, 32, dtype=torch.bfloat16, device="cuda")
    plan = C.Plan()
    plan.fork(1)
    plan.branch(1)
    plan.add_cast_transpose(native.ptr(x), 32, 0, native.ptr(a), 32, 0, 0, 64, 32)
    plan.branch(0)
    plan.add_cast_transpose(native.ptr(x), 32, 0, native.ptr(b), 32, 0, 0, 64, 32)
    plan.join(1)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        plan.run(st.cuda_stream)
        st.synchronize()
        plan.capture(st.cuda_stream)
        a.zero_(); b.zero_()
        plan.replay(st.cuda_stream)
    st.synchronize()
    assert torch.equal(a, x.to(torch.bfloat16)) and torch.equal(b, x.to(torch.bfloat16))
    assert plan.graph_nodes() >= 2 and plan.names() == ["cast_transpose@1", "cast_transpose"]


def test_im2col_and_col2im(C):
    n, h, w, c, kh, kw = 5, 12, 12, 32, 3, 3
    x = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
    oh, ow = h - kh + 1, w - kw + 1
    M, K = n * oh * ow, kh * kw * c
    ldK, ldM = round_up(K, 8), round_up(M, 8)
    out = torch.zeros(M, ldK, dtype=torch.bfloat16, device="cuda")
    outT = torch.zeros(K, ldM, dtype=torch.bfloat16, device="cuda")
    st = native.current_stream()
    C.im2col(native.ptr(x), n, h, w, c, kh, kw, native.ptr(out), ldK, native.ptr(outT), ldM, st)
    torch.cuda.synchronize()
    ref = torch.nn.functional.unfold(x.float().permute(0, 3, 1, 2), (kh, kw))            # [n, c*kh*kw, L] in (c, kh, kw) order
    ref = ref.reshape(n, c, kh * kw, oh * ow).permute(0, 3, 2, 1).reshape(M, K)       # -> (kh, kw, c) order
    assert torch.equal(out[:, :K].float(), ref) and torch.equal(outT[:, :M].float().t(), ref)
    # col2im == adjoint of im2col
    dcols = torch.randn(M, ldK, device="cuda").to(torch.bfloat16)
    din = torch.zeros(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    C.col2im(native.ptr(dcols), ldK, n, h, w, c, kh, kw, native.ptr(din), st)
    torch.cuda.synchronize()
    d = dcols[:, :K].float().reshape(n, oh * ow, kh * kw, c).permute(0, 3, 2, 1).reshape(n, c * kh * kw, oh * ow)
    ref_in = torch.nn.functional.fold(d, (h, w), (kh, kw)).permute(0, 2, 3, 1)
    torch.testing.assert_close(din.float(), ref_in, rtol=2e-2, atol=5e-2)


def test_maxpool_fwd_bwd(C):
    n, h, w, c = 4, 10, 10, 64
    a = torch.relu(torch.randn(n, h, w, c, device="cuda")).to(torch.bfloat16)
    oh, ow = h // 2, w // 2
    out = torch.zeros(n, oh, ow, c, dtype=torch.bfloat16, device="cuda")
    am = torch.zeros(n, oh, ow, c, dtype=torch.uint8, device="cuda")
    ldB = round_up(n, 8)
    outT = torch.zeros(oh * ow * c, ldB, dtype=torch.bfloat16, device="cuda")
    st = native.current_stream()
    C.maxpool_fwd(native.ptr(a), n, h, w, c, native.ptr(out), native.ptr(am), native.ptr(outT), ldB, st)
    torch.cuda.synchronize()
    af = a.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = torch.nn.functional.max_pool2d(torch.relu(af), 2)
    assert torch.equal(out.float(), ref.detach().permute(0, 2, 3, 1))
    assert torch.equal(outT[:, :n].float().t(), ref.detach().permute(0, 2, 3, 1).reshape(n, -1))
    dout = torch.randn(n, oh, ow, c, device="cuda").to(torch.bfloat16)
    ref.backward(dout.float().permute(0, 3, 1, 2))
    M = n * h * w
    ldM = round_up(M, 8)
    dz = torch.zeros(M, c, dtype=torch.bfloat16, device="cuda")
    dzT = torch.zeros(c, ldM, dtype=torch.bfloat16, device="cuda")
    db = torch.zeros(c, device="cuda")
    C.maxpool_bwd(native.ptr(dout), native.ptr(am), n, h, w, c, native.ptr(a), 1, native.ptr(dz), c, native.ptr(dzT), ldM, native.ptr(db), st)
    torch.cuda.synchronize()
    g = af.grad.permute(0, 2, 3, 1).reshape(M, c)
    # ties inside a window may pick a different (equal-valued) winner than torch: compare where unambiguous
    torch.testing.assert_close(dz.float().sum(), g.sum(), rtol=1e-2, atol=1e-1)
    mism = (dz.float() != g).float().mean().item()
    assert mism < 0.02, mism
    assert torch.equal(dzT[:, :M].float().t(), dz.float())
    torch.testing.assert_close(db, dz.float().sum(0), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("rows,cols,ycols,B,start,ldpad", [(1000, 784, 10, 300, 37, 0), (500, 50, 3, 64, 436, 0), (256, 36, 1, 33, 0, 0),
                                                           (400, 30, 2, 50, 11, 6)])
def test_fetch_kernel_zero_copy_minibatch(C, rows, cols, ycols, B, start, ldpad):
    """in-graph minibatch fetch: SM loads stream a minibatch (features + labels) from pinned host memory into staging"""
    g = torch.Generator().manual_seed(1)
    Xfull = torch.randn(rows, cols + ldpad, generator=g).pin_memory()
    Yfull = torch.randn(rows, ycols + ldpad, generator=g).pin_memory()
    x_out = torch.full((B, cols), 7.0, device="cuda")
    y_out = torch.zeros(B, ycols, device="cuda")
    sched = torch.zeros(8, dtype=torch.int64).pin_memory()
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    sync = torch.zeros(1, dtype=torch.int32, device="cuda")
    desc = torch.tensor([Xfull.data_ptr(), Yfull.data_ptr(), cols + ldpad, ycols + ldpad], dtype=torch.int64).cuda()
    args = dict(desc=native.ptr(desc), sched=sched.data_ptr(), ring_mask=7, counter=native.ptr(counter), sync=native.ptr(sync),
                x_out=native.ptr(x_out), y_out=native.ptr(y_out), rows=B, cols=cols, y_cols=ycols)
    sched[0] = 5            # first fetch (sequence 0) reads entry 0 ...
    sched[1] = start        # ... the second one entry 1
    for _ in range(2):
        C.fetch(args, 5, native.current_stream())
    torch.cuda.synchronize()
    assert int(counter.item()) == 2 and int(sync.item()) == 0
    assert torch.equal(x_out.cpu(), Xfull[start:start + B, :cols])
    assert torch.equal(y_out.cpu(), Yfull[start:start + B, :ycols])


@pytest.mark.parametrize("n,h,w,cin,kh,kw,cout,act", [(5, 28, 28, 1, 5, 5, 32, "relu"), (3, 12, 12, 2, 3, 3, 16, "tanh"), (2, 9, 9, 3, 2, 2, 64, None)])
def test_conv_first_fused_fwd_and_wgrad(C, n, h, w, cin, kh, kw, cout, act):
    """First conv + bias + activation + 2x2 max-pool as one direct kernel, and its fused weight / bias gradient, against
    torch (conv2d -> act -> max_pool2d and autograd through them)."""
    import torch.nn.functional as F

    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).to(torch.bfloat16)
    K = kh * kw * cin
    wmat = (torch.randn(K, cout, device="cuda", generator=g) * 0.3).to(torch.bfloat16)      # [K = (dy, dx, ci), cout] (HWIO flattened)
    ld_w = round_up(K, 8)
    wT = torch.zeros(cout, ld_w, dtype=torch.bfloat16, device="cuda")
    wT[:, :K] = wmat.t()
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1
    oh, ow = h - kh + 1, w - kw + 1
    ph, pw = oh // 2, ow // 2
    pooled = torch.zeros(n, ph, pw, cout, dtype=torch.bfloat16, device="cuda")
    argmax = torch.zeros(n, ph, pw, cout, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    C.conv_first_fwd(x.data_ptr(), n, h, w, cin, kh, kw, cout, wT.data_ptr(), ld_w, bias.data_ptr(), ACT[act], pooled.data_ptr(), argmax.data_ptr(), st)
    torch.cuda.synchronize()
    xt = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wt = wmat.float().reshape(kh, kw, cin, cout).permute(3, 2, 0, 1).clone().requires_grad_(True)
    bt = bias.clone().requires_grad_(True)
    conv = _act(F.conv2d(xt, wt, bt), act)
    ref = F.max_pool2d(conv, 2)[:, :, :ph, :pw]
    assert torch.allclose(pooled.float().permute(0, 3, 1, 2), ref, rtol=2e-2, atol=2e-2)
    # gradient of sum(g_pool * pooled) through the SAME arg-max positions the kernel recorded
    gp = torch.randn(n, ph, pw, cout, device="cuda", generator=g).to(torch.bfloat16)
    dW = torch.zeros(K, cout, device="cuda")
    db = torch.zeros(cout, device="cuda")
    C.conv_first_wgrad(x.data_ptr(), n, h, w, cin, kh, kw, cout, gp.data_ptr(), pooled.data_ptr(), argmax.data_ptr(), ACT[act], dW.data_ptr(),
                       db.data_ptr(), st)
    torch.cuda.synchronize()
    am = argmax.long().permute(0, 3, 1, 2)
    ys = 2 * torch.arange(ph, device="cuda").view(1, 1, ph, 1) + am // 2
    xs = 2 * torch.arange(pw, device="cuda").view(1, 1, 1, pw) + am % 2
    picked = conv[torch.arange(n, device="cuda").view(n, 1, 1, 1), torch.arange(cout, device="cuda").view(1, cout, 1, 1), ys, xs]
    # the kernel takes act'(.) from the bf16-rounded pooled value: use the same mask convention for ReLU ties
    (picked * gp.float().permute(0, 3, 1, 2)).sum().backward()
    ref_dW = wt.grad.permute(2, 3, 1, 0).reshape(K, cout)
    scale = max(1.0, ref_dW.abs().max().item())
    assert (dW - ref_dW).abs().max().item() < 3e-2 * scale
    assert torch.allclose(db, bt.grad, rtol=3e-2, atol=3e-2 * max(1.0, bt.grad.abs().max().item()))
