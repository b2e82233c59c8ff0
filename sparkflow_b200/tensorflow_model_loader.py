"""Load a pre-trained TensorFlow checkpoint (``<path>.meta`` + V2 bundle) as a ``SparkAsyncDLModel``.

API parity with /root/reference/sparkflow/tensorflow_model_loader.py:8-45, implemented without
TensorFlow: the binary MetaGraphDef is decoded by ``graph.pbwire`` and the ``.index`` / ``.data``
bundle by the native TF-bundle reader (``io.bundle``).
"""
from __future__ import annotations

import json
import os

from .graph import pbwire
from .graph.ir import GraphIR
from .io.bundle import read_bundle, read_checkpoint_state
from .spark.ml.base import PipelineModel
from .tensorflow_async import SparkAsyncDLModel


def load_tensorflow_model(path, inputCol, tfInput, tfOutput, predictionCol="predicted", tfDropout=None, toKeepDropout=False):
    with open(path + ".meta", "rb") as fh:
        meta = pbwire.decode("MetaGraphDef", fh.read())
    ckpt_dir = os.path.dirname(path) or "."
    prefix = read_checkpoint_state(ckpt_dir) or path      # ``tf.train.latest_checkpoint`` of the directory
    tensors = read_bundle(prefix)
    ir = GraphIR.from_metagraph(meta)
    weights = []
    for v in ir.trainable:
        if v.name not in tensors:
            raise KeyError(f"variable '{v.name}' is missing from checkpoint {prefix}")
        weights.append(tensors[v.name].reshape(v.shape).tolist())
    return SparkAsyncDLModel(inputCol=inputCol, modelJson=json.dumps(meta), modelWeights=json.dumps(weights), tfInput=tfInput,
                             tfOutput=tfOutput, predictionCol=predictionCol, tfDropout=tfDropout, toKeepDropout=toKeepDropout)


def attach_tensorflow_model_to_pipeline(path, pipelineModel, inputCol, tfInput, tfOutput, predictionCol="predicted",
                                        tfDropout=None, toKeepDropout=False):
    spark_model = load_tensorflow_model(path, inputCol, tfInput, tfOutput, predictionCol, tfDropout, toKeepDropout)
    return PipelineModel(stages=[pipelineModel, spark_model])
