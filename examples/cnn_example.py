"""MNIST CNN (conv5x5x32 - pool - conv3x3x64 - pool - dense10), counterpart of examples/cnn_example.py.
On a B200 the whole step runs on the compiled sm_100a plan (im2col + tcgen05 GEMMs + fused pool backward)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkflow_b200 import compat

compat.install()

import tensorflow as tf
from pyspark.ml.feature import OneHotEncoder, VectorAssembler
from pyspark.ml.pipeline import Pipeline
from pyspark.sql import SparkSession
from pyspark.sql.functions import rand
from sparkflow.graph_utils import build_graph
from sparkflow.tensorflow_async import SparkAsyncDL

from _data import mnist_csv


def cnn_model():
    x = tf.placeholder(tf.float32, shape=[None, 784], name="x")
    y = tf.placeholder(tf.float32, shape=[None, 10], name="y")
    x = tf.reshape(x, shape=[-1, 28, 28, 1])
    conv1 = tf.layers.max_pooling2d(tf.layers.conv2d(x, 32, 5, activation=tf.nn.relu), 2, 2)
    conv2 = tf.layers.max_pooling2d(tf.layers.conv2d(conv1, 64, 3, activation=tf.nn.relu), 2, 2)
    out = tf.layers.dense(tf.layers.flatten(conv2), 10)
    tf.argmax(out, 1, name="out")
    return tf.losses.softmax_cross_entropy(y, out)


if __name__ == "__main__":
    rows = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else None
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 50
    spark = SparkSession.builder.appName("examples").master("local[4]").config("spark.driver.memory", "4g").getOrCreate()
    df = spark.read.option("inferSchema", "true").csv(mnist_csv()).orderBy(rand(seed=1))
    if rows:
        df = df.limit(rows).repartition(4)
    mg = build_graph(cnn_model)
    va = VectorAssembler(inputCols=df.columns[1:785], outputCol="features")
    encoded = OneHotEncoder(inputCol="_c0", outputCol="labels", dropLast=False)
    spark_model = SparkAsyncDL(inputCol="features", tensorflowGraph=mg, tfInput="x:0", tfLabel="y:0", tfOutput="out:0",
                               tfOptimizer="adam", miniBatchSize=300, miniStochasticIters=-1, shufflePerIter=True, iters=iters,
                               partitions=4, tfLearningRate=.0001, predictionCol="predicted", labelCol="labels", verbose=1)
    p = Pipeline(stages=[va, encoded, spark_model]).fit(df)
    p.write().overwrite().save("/tmp/cnn")
    print("saved /tmp/cnn;", p.transform(df).take(1)[0]["predicted"])
