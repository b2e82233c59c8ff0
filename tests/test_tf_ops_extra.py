"""CPU tests of the wider TF-1.x op surface of the builder + interpreter engine (embeddings, sparse labels, splits / pads,
transposed / depthwise convolutions, fused batch-norm, tf.cond): every op against a numpy / torch closed form, and one
model that needs them trained through the public session API."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from sparkflow_b200.graph import tfcompat as tf
from sparkflow_b200.graph.executor import GraphProgram
from sparkflow_b200.graph.ir import GraphIR
from sparkflow_b200.graph.tfcompat.core import attr_b, attr_f, attr_s, attr_type
from sparkflow_b200.graph_utils import build_graph
from sparkflow_b200.ops.optimizers import OptimizerSpec


def _run(fetch_fn, feeds):
    g = tf.Graph()
    with g.as_default():
        phs = {k: tf.placeholder(tf.float32, shape=[None] + list(v.shape[1:]), name=k) for k, v in feeds.items()}
        out = fetch_fn(**phs)
        with tf.Session(graph=g) as sess:
            return sess.run(out, feed_dict={k + ":0": v for k, v in feeds.items()})


def test_gather_one_hot_split_pad_ops():
    rng = np.random.default_rng(0)
    table = rng.standard_normal((11, 4)).astype(np.float32)
    ids = rng.integers(0, 11, (5, 3)).astype(np.float32)
    got = _run(lambda t, i: tf.nn.embedding_lookup(t, tf.cast(i, tf.int32)), dict(t=table, i=ids))
    np.testing.assert_allclose(got, table[ids.astype(int)])
    got = _run(lambda t, i: tf.gather(t, tf.cast(i, tf.int32), axis=1), dict(t=table, i=np.asarray([[3.0], [0.0]], np.float32)))
    np.testing.assert_allclose(got, np.take(table, [[3], [0]], axis=1))
    np.testing.assert_allclose(_run(lambda i: tf.one_hot(tf.cast(i, tf.int32), 11), dict(i=ids)), np.eye(11, dtype=np.float32)[ids.astype(int)])
    a = rng.standard_normal((6, 8)).astype(np.float32)
    parts = _run(lambda a: tf.split(a, 4, axis=1), dict(a=a))
    assert len(parts) == 4 and all(np.array_equal(p, q) for p, q in zip(parts, np.split(a, 4, axis=1)))
    parts = _run(lambda a: tf.split(a, [1, 5, 2], axis=1), dict(a=a))
    assert [p.shape[1] for p in parts] == [1, 5, 2] and np.array_equal(np.concatenate(parts, 1), a)
    rows = _run(lambda a: tf.unstack(a, num=6, axis=0), dict(a=a))
    assert len(rows) == 6 and np.array_equal(np.stack(rows), a)
    np.testing.assert_allclose(_run(lambda a: tf.pad(a, [[0, 0], [2, 1]]), dict(a=a)), np.pad(a, [(0, 0), (2, 1)]))
    np.testing.assert_allclose(_run(lambda a: tf.pad(a, [[1, 0], [0, 3]], constant_values=7.5), dict(a=a)), np.pad(a, [(1, 0), (0, 3)], constant_values=7.5))
    np.testing.assert_allclose(_run(lambda a: tf.pad(a, [[0, 0], [2, 3]], mode="REFLECT"), dict(a=a)), np.pad(a, [(0, 0), (2, 3)], mode="reflect"))
    np.testing.assert_allclose(_run(lambda a: tf.pad(a, [[1, 1], [2, 2]], mode="SYMMETRIC"), dict(a=a)), np.pad(a, [(1, 1), (2, 2)], mode="symmetric"))


def test_elementwise_and_reduction_additions():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((5, 7)).astype(np.float32) * 3
    b = (rng.random((5, 7)).astype(np.float32) + 0.5) * 2
    np.testing.assert_allclose(_run(lambda a: tf.erf(a) + tf.sin(a) * tf.cos(a), dict(a=a)), torch.erf(torch.tensor(a)).numpy() + np.sin(a) * np.cos(a),
                               rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_run(lambda a: tf.round(a), dict(a=a)), np.round(a))
    np.testing.assert_allclose(_run(lambda a, b: tf.floormod(a, b), dict(a=a, b=b)), np.mod(a, b), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(_run(lambda a: tf.cumsum(a, axis=1), dict(a=a)), np.cumsum(a, 1), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(_run(lambda a: tf.cumsum(a, axis=0, exclusive=True, reverse=True), dict(a=a)),
                               np.flip(np.cumsum(np.flip(a, 0), 0) - np.flip(a, 0), 0), rtol=1e-5, atol=1e-5)
    assert np.array_equal(_run(lambda a: tf.reduce_any(tf.logical_or(tf.greater(a, 5.0), tf.less(a, -5.0)), axis=1), dict(a=a)), (np.abs(a) > 5).any(1))
    assert np.array_equal(_run(lambda a: tf.reduce_all(tf.less(a, 6.0), axis=0, keepdims=True), dict(a=a)), (a < 6).all(0, keepdims=True))
    vals, idx = _run(lambda a: tf.nn.top_k(a, k=3), dict(a=a))
    order = np.argsort(-a, axis=1)[:, :3]
    np.testing.assert_allclose(vals, np.take_along_axis(a, order, 1))
    assert np.array_equal(idx, order)


def test_sparse_softmax_cross_entropy_matches_dense_form():
    rng = np.random.default_rng(2)
    logits = rng.standard_normal((9, 6)).astype(np.float32)
    lab = rng.integers(0, 6, (9,)).astype(np.float32)
    got = _run(lambda z, y: tf.nn.sparse_softmax_cross_entropy_with_logits(labels=tf.cast(y, tf.int32), logits=z), dict(z=logits, y=lab))
    logp = logits - np.log(np.exp(logits).sum(1, keepdims=True))
    np.testing.assert_allclose(got, -logp[np.arange(9), lab.astype(int)], rtol=1e-5)
    mean = _run(lambda z, y: tf.losses.sparse_softmax_cross_entropy(tf.cast(y, tf.int32), z), dict(z=logits, y=lab))
    np.testing.assert_allclose(mean, got.mean(), rtol=1e-5)


def _conv_fwd(x, w, stride, padding):
    g = tf.Graph()
    with g.as_default():
        xt = tf.constant(x)
        wt = tf.constant(w)
        out = tf.nn.conv2d(xt, wt, [1, stride, stride, 1], padding)
        with tf.Session(graph=g) as sess:
            return sess.run(out)


def _conv_t(y, w, out_shape, stride, padding):
    g = tf.Graph()
    with g.as_default():
        out = tf.nn.conv2d_transpose(tf.constant(y), tf.constant(w), out_shape, [1, stride, stride, 1], padding)
        with tf.Session(graph=g) as sess:
            return sess.run(out)


@pytest.mark.parametrize("padding,stride,k,h", [("VALID", 2, 3, 9), ("SAME", 2, 3, 8), ("SAME", 1, 5, 6), ("VALID", 1, 2, 5), ("SAME", 2, 4, 7)])
def test_conv2d_transpose_adjoint_property(padding, stride, k, h):
    rng = np.random.default_rng(4)
    cin, cout = 3, 4
    x = rng.standard_normal((2, h, h, cin)).astype(np.float32)
    w = rng.standard_normal((k, k, cin, cout)).astype(np.float32)           # forward filter HWIO == transpose filter [kh, kw, out, in]
    y_like = _conv_fwd(x, w, stride, padding)
    y = rng.standard_normal(y_like.shape).astype(np.float32)
    xt = _conv_t(y, w, list(x.shape), stride, padding)
    assert xt.shape == x.shape
    np.testing.assert_allclose((y_like * y).sum(), (x * xt).sum(), rtol=2e-4)


def test_depthwise_conv_matches_grouped_conv():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 7, 7, 3)).astype(np.float32)
    w = rng.standard_normal((3, 3, 3, 2)).astype(np.float32)
    g = tf.Graph()
    with g.as_default():
        out = tf.nn.depthwise_conv2d(tf.constant(x), tf.constant(w), [1, 1, 1, 1], "SAME")
        with tf.Session(graph=g) as sess:
            got = sess.run(out)
    ref = F.conv2d(torch.tensor(x).permute(0, 3, 1, 2), torch.tensor(w).permute(2, 3, 0, 1).reshape(6, 1, 3, 3), padding=1, groups=3)
    np.testing.assert_allclose(got, ref.permute(0, 2, 3, 1).numpy(), rtol=1e-4, atol=1e-5)


def test_fused_batch_norm_and_batch_matmul_nodes():
    """Ops that only REAL TF graphs carry (the builder lowers batch-norm to primitive ops): nodes added by hand."""
    rng = np.random.default_rng(6)
    x = (rng.standard_normal((4, 5, 5, 3)) * 2 + 1).astype(np.float32)
    scale, offset = rng.random(3).astype(np.float32) + 0.5, rng.standard_normal(3).astype(np.float32)
    mean, var = rng.standard_normal(3).astype(np.float32), rng.random(3).astype(np.float32) + 0.5
    for training in (True, False):
        g = tf.Graph()
        with g.as_default():
            ins = [tf.constant(v) for v in (x, scale, offset, mean, var)]
            op = g.add_node("FusedBatchNormV3", "bn", ins, {"T": attr_type(tf.float32), "U": attr_type(tf.float32), "epsilon": attr_f(1e-3),
                                                           "is_training": attr_b(training), "data_format": attr_s("NHWC")},
                            [tf.float32] * 6, [x.shape] + [(3,)] * 5)
            with tf.Session(graph=g) as sess:
                y, bm, bv = sess.run([op.outputs[0], op.outputs[1], op.outputs[2]])
        if training:
            m, v = x.mean((0, 1, 2)), x.var((0, 1, 2))
            np.testing.assert_allclose(bm, m, rtol=1e-5)
            np.testing.assert_allclose(bv, v * (100 / 99), rtol=1e-4)
        else:
            m, v = mean, var
        np.testing.assert_allclose(y, (x - m) / np.sqrt(v + 1e-3) * scale + offset, rtol=1e-4, atol=1e-5)
    a, b = rng.standard_normal((3, 4, 5)).astype(np.float32), rng.standard_normal((3, 6, 5)).astype(np.float32)
    g = tf.Graph()
    with g.as_default():
        op = g.add_node("BatchMatMulV2", "bmm", [tf.constant(a), tf.constant(b)], {"T": attr_type(tf.float32), "adj_x": attr_b(False), "adj_y": attr_b(True)},
                        [tf.float32], [(3, 4, 6)])
        with tf.Session(graph=g) as sess:
            np.testing.assert_allclose(sess.run(op.outputs[0]), a @ b.transpose(0, 2, 1), rtol=1e-5)


def test_cond_switch_merge_takes_the_live_branch():
    rng = np.random.default_rng(7)
    a = rng.standard_normal((4, 3)).astype(np.float32)
    g = tf.Graph()
    with g.as_default():
        x = tf.placeholder(tf.float32, [None, 3], name="x")
        training = tf.placeholder_with_default(False, shape=(), name="training")
        y = tf.cond(training, lambda: tf.nn.relu(x) * 2.0, lambda: x - 1.0)
        z = tf.identity(y + 10.0, name="z")                                    # ops after the Merge see the live value
        with tf.Session(graph=g) as sess:
            np.testing.assert_allclose(sess.run(z, {"x:0": a}), a - 1 + 10)
            np.testing.assert_allclose(sess.run(z, {"x:0": a, "training:0": True}), np.maximum(a, 0) * 2 + 10)
    assert tf.cond(True, lambda: 1, lambda: 2) == 1                            # python predicate: decided while building


def test_embedding_classifier_trains_through_the_session_api():
    """tokens -> embedding_lookup -> mean -> dense -> sparse softmax CE: needs GatherV2 (with a gradient into the table), Cast,
    sparse labels; trained with the interpreter engine through TrainingSession like any graph outside the compiled family."""
    from sparkflow_b200.parallel.session import TrainingSession

    def model():
        x = tf.placeholder(tf.float32, shape=[None, 6], name="x")              # token ids arrive as floats (DataFrame vectors)
        y = tf.placeholder(tf.float32, shape=[None, 1], name="y")
        table = tf.get_variable("embedding", [20, 8], initializer=tf.random_normal_initializer(stddev=0.5))
        emb = tf.nn.embedding_lookup(table, tf.cast(x, tf.int32))
        h = tf.reduce_mean(emb, axis=1)
        logits = tf.layers.dense(h, 2, name="cls")
        tf.argmax(logits, 1, name="out")
        return tf.losses.sparse_softmax_cross_entropy(tf.cast(tf.reshape(y, [-1]), tf.int32), logits)

    graph = build_graph(model)
    ir = GraphIR.from_metagraph(graph)
    assert [v.name for v in ir.trainable] == ["embedding", "cls/kernel", "cls/bias"]
    rng = np.random.default_rng(8)
    X = rng.integers(0, 20, (400, 6)).astype(np.float32)
    Y = ((X < 10).sum(1) >= 3).astype(np.float32).reshape(-1, 1)               # class = "at least half the tokens are small ids"
    prog = GraphProgram(ir)
    w0 = prog.init_weights(seed=0)
    l0, grads = prog.loss_and_grads({"x:0": X, "y:0": Y}, w0)
    assert float(grads[0].abs().sum()) > 0                                      # the table receives a gradient through the gather
    sess = TrainingSession(graph, "x:0", "y:0", OptimizerSpec.from_tf_kwargs("adam", dict(learning_rate=0.05)), iters=40, mini_batch=100,
                           engine="torch", seed=0, initial_weights=[np.asarray(w) for w in w0]).open()
    sess.train_partitions([(X, Y)])
    w1 = sess.weights()
    sess.close()
    l1 = prog.loss({"x:0": X, "y:0": Y}, w1)
    acc = float((prog.forward("out:0", {"x:0": X}, w1).numpy() == Y.reshape(-1)).mean())
    assert l1 < 0.5 * l0 and acc > 0.85, (l0, l1, acc)


def test_tensor_indexing_and_comparison_sugar():
    rng = np.random.default_rng(9)
    a = rng.standard_normal((5, 6, 4)).astype(np.float32)
    np.testing.assert_allclose(_run(lambda a: a[:, 1:4], dict(a=a)), a[:, 1:4])
    np.testing.assert_allclose(_run(lambda a: a[:, :, 0], dict(a=a)), a[:, :, 0])
    np.testing.assert_allclose(_run(lambda a: a[:, -1], dict(a=a)), a[:, -1])
    np.testing.assert_allclose(_run(lambda a: a[1:, ::2, 1:-1], dict(a=a)), a[1:, ::2, 1:-1])
    np.testing.assert_allclose(_run(lambda a: a[0], dict(a=a)), a[0])
    np.testing.assert_allclose(_run(lambda a: tf.where(a > 0.0, a ** 2.0, tf.zeros_like(a)), dict(a=a)), np.where(a > 0, a ** 2, 0), rtol=1e-5)
    assert np.array_equal(_run(lambda a: tf.logical_and(a >= -1.0, a <= 1.0), dict(a=a)), (a >= -1) & (a <= 1))
