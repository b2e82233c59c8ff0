"""Model zoo: every workload the reference ships plus the BASELINE.json
stress configs, as plain graph-builder functions usable with ``build_graph``.

Each function mirrors a model definition of the reference:
  * :func:`simple_dnn`        – examples/simple_dnn.py:13-21 (784-256-256-10, softmax-CE, ``out`` = argmax)
  * :func:`cnn`               – examples/cnn_example.py:10-22 (conv5x5x32-pool-conv3x3x64-pool-dense10)
  * :func:`autoencoder`       – examples/autoencoder_example.py:9-16 (784-256-128-256-784, MSE vs input)
  * :func:`test_mlp` / :func:`test_autoencoder` – tests/dl_runner.py:45-73
  * :func:`fixture_mlp`       – the graph stored in tests/test_model/to_load.meta (2-10-10-1)
  * :func:`autoencoder_small` / :func:`wide_dnn` – BASELINE.json configs 4 and 5
"""
from __future__ import annotations

from typing import Callable, Dict

from ..graph import tfcompat as tf


def simple_dnn():
    x = tf.placeholder(tf.float32, shape=[None, 784], name="x")
    y = tf.placeholder(tf.float32, shape=[None, 10], name="y")
    h1 = tf.layers.dense(x, 256, activation=tf.nn.relu, kernel_initializer=tf.glorot_uniform_initializer())
    h2 = tf.layers.dense(h1, 256, activation=tf.nn.relu, kernel_initializer=tf.glorot_uniform_initializer())
    logits = tf.layers.dense(h2, 10, kernel_initializer=tf.glorot_uniform_initializer())
    tf.argmax(logits, 1, name="out")
    return tf.losses.softmax_cross_entropy(y, logits)


def cnn():
    x = tf.placeholder(tf.float32, shape=[None, 784], name="x")
    y = tf.placeholder(tf.float32, shape=[None, 10], name="y")
    img = tf.reshape(x, shape=[-1, 28, 28, 1])
    c1 = tf.layers.max_pooling2d(tf.layers.conv2d(img, 32, 5, activation=tf.nn.relu), 2, 2)
    c2 = tf.layers.max_pooling2d(tf.layers.conv2d(c1, 64, 3, activation=tf.nn.relu), 2, 2)
    logits = tf.layers.dense(tf.layers.flatten(c2), 10)
    tf.argmax(logits, 1, name="out")
    return tf.losses.softmax_cross_entropy(y, logits)


def autoencoder():
    x = tf.placeholder("float", shape=[None, 784], name="x")
    e1 = tf.layers.dense(x, 256, activation=tf.nn.relu)
    code = tf.layers.dense(e1, 128, activation=tf.nn.sigmoid, name="out")
    d1 = tf.layers.dense(code, 256, activation=tf.nn.relu)
    rec = tf.layers.dense(d1, 784, activation=tf.nn.sigmoid)
    return tf.losses.mean_squared_error(rec, x)


def autoencoder_small():
    x = tf.placeholder(tf.float32, shape=[None, 784], name="x")
    code = tf.layers.dense(x, 32, activation=tf.nn.sigmoid, name="out")
    rec = tf.layers.dense(code, 784, activation=tf.nn.sigmoid)
    return tf.losses.mean_squared_error(rec, x)


def wide_dnn(width: int = 4096, depth: int = 4, classes: int = 1000):
    x = tf.placeholder(tf.float32, shape=[None, width], name="x")
    y = tf.placeholder(tf.float32, shape=[None, classes], name="y")
    h = x
    for _ in range(depth):
        h = tf.layers.dense(h, width, activation=tf.nn.relu)
    logits = tf.layers.dense(h, classes)
    tf.argmax(logits, 1, name="out")
    return tf.losses.softmax_cross_entropy(y, logits)


def test_mlp(in_dim: int = 10):
    x = tf.placeholder(tf.float32, shape=[None, in_dim], name="x")
    h1 = tf.layers.dense(x, 12, activation=tf.nn.relu)
    h2 = tf.layers.dense(h1, 7, activation=tf.nn.relu)
    out = tf.layers.dense(h2, 1, name="outer", activation=tf.nn.sigmoid)
    y = tf.placeholder(tf.float32, shape=[None, 1], name="y")
    return tf.losses.mean_squared_error(y, out)


def test_autoencoder():
    x = tf.placeholder(tf.float32, shape=[None, 10], name="x")
    enc = tf.layers.dense(x, 5, activation=tf.nn.relu)
    code = tf.layers.dense(enc, 2, activation=tf.nn.sigmoid, name="out")
    dec = tf.layers.dense(code, 5, activation=tf.nn.relu)
    rec = tf.layers.dense(dec, 10)
    return tf.losses.mean_squared_error(x, rec)


def fixture_mlp():
    x = tf.placeholder(tf.float32, shape=[None, 2], name="x")
    y = tf.placeholder(tf.float32, shape=[None, 1], name="y")
    h1 = tf.layers.dense(x, 10, activation=tf.nn.tanh)
    h2 = tf.layers.dense(h1, 10, activation=tf.nn.tanh)
    out = tf.layers.dense(h2, 1, activation=tf.nn.sigmoid, name="out")
    return tf.losses.mean_squared_error(y, out)


MODELS: Dict[str, Callable] = {
    "simple_dnn": simple_dnn,
    "cnn": cnn,
    "autoencoder": autoencoder,
    "autoencoder_small": autoencoder_small,
    "wide_dnn": wide_dnn,
    "test_mlp": test_mlp,
    "test_autoencoder": test_autoencoder,
    "fixture_mlp": fixture_mlp,
}


def build(name: str, **kwargs) -> str:
    """MetaGraphDef JSON of a zoo model (same as ``build_graph(MODELS[name])``)."""
    g = tf.Graph()
    with g.as_default():
        MODELS[name](**kwargs)
        return tf.train.export_meta_graph().to_json(indent=None)
