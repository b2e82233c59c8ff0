"""Unsupervised auto-encoder 784-256-128-256-784 on L1-normalised MNIST rows (tfLabel=None: the input is the target);
the 128-wide sigmoid bottleneck (`out/Sigmoid:0`) is what transform() writes into the prediction column.
Same workload as the reference's autoencoder example.

    python examples/autoencoder_example.py [--rows N] [--iters K]
"""
from _common import Stopwatch, mnist_frame, parse_args, pixel_columns

import tensorflow as tf
from pyspark.ml.feature import Normalizer, VectorAssembler
from sparkflow.graph_utils import build_graph
from sparkflow.tensorflow_async import SparkAsyncDL

BATCH = 256
# (units, activation, layer name); the layer called "out" is the code the fitted model emits
ENCODER = ((256, "relu", None), (128, "sigmoid", "out"))
DECODER = ((256, "relu", None), (784, "sigmoid", None))


def autoencoder():
    pixels = tf.placeholder("float", shape=[None, 784], name="x")
    h = pixels
    for units, act, name in ENCODER + DECODER:
        h = tf.layers.dense(h, units, activation=getattr(tf.nn, act), name=name)
    return tf.losses.mean_squared_error(h, pixels)


def main():
    args = parse_args(__doc__.splitlines()[0], iters=10)
    spark, frame = mnist_frame(args)
    assembled = VectorAssembler(inputCols=pixel_columns(frame), outputCol="raw").transform(frame).select(["raw"])
    rows = Normalizer(inputCol="raw", outputCol="features", p=1.0).transform(assembled).select(["features"])
    estimator = SparkAsyncDL(inputCol="features", predictionCol="predicted", tensorflowGraph=build_graph(autoencoder), tfInput="x:0",
                             tfLabel=None, tfOutput="out/Sigmoid:0", tfOptimizer="adam", tfLearningRate=0.001, iters=args.iters,
                             partitions=args.partitions, miniBatchSize=BATCH, verbose=0 if args.quiet else 1)
    with Stopwatch("autoencoder", rows.count(), args.iters, BATCH):
        model = estimator.fit(rows)
    code = model.transform(rows).take(1)[0]["predicted"]
    print("bottleneck code of the first row:", code)
    spark.stop()


if __name__ == "__main__":
    main()
