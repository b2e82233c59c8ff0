"""MNIST MLP through the Spark-ML pipeline API (counterpart of the reference's examples/simple_dnn.py).

Run:  python examples/simple_dnn.py [--rows N]            (single process: one worker per GPU / thread)
      torchrun --nproc-per-node 8 examples/simple_dnn.py   (one rank per B200)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkflow_b200 import compat

compat.install()

import tensorflow as tf
from pyspark.ml.evaluation import MulticlassClassificationEvaluator
from pyspark.ml.feature import OneHotEncoder, VectorAssembler
from pyspark.ml.pipeline import Pipeline, PipelineModel
from pyspark.sql import SparkSession
from pyspark.sql.functions import rand
from sparkflow.graph_utils import build_adam_config, build_graph
from sparkflow.pipeline_util import PysparkPipelineWrapper
from sparkflow.tensorflow_async import SparkAsyncDL

from _data import mnist_csv


def small_model():
    x = tf.placeholder(tf.float32, shape=[None, 784], name="x")
    y = tf.placeholder(tf.float32, shape=[None, 10], name="y")
    layer1 = tf.layers.dense(x, 256, activation=tf.nn.relu, kernel_initializer=tf.glorot_uniform_initializer())
    layer2 = tf.layers.dense(layer1, 256, activation=tf.nn.relu, kernel_initializer=tf.glorot_uniform_initializer())
    out = tf.layers.dense(layer2, 10, kernel_initializer=tf.glorot_uniform_initializer())
    tf.argmax(out, 1, name="out")
    return tf.losses.softmax_cross_entropy(y, out)


if __name__ == "__main__":
    rows = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else None
    spark = SparkSession.builder.appName("examples").master("local[4]").config("spark.driver.memory", "2g").getOrCreate()
    df = spark.read.option("inferSchema", "true").csv(mnist_csv()).orderBy(rand(seed=1))
    if rows:
        df = df.limit(rows).repartition(4)

    mg = build_graph(small_model)
    adam_config = build_adam_config(learning_rate=0.001, beta1=0.9, beta2=0.999)
    vector_assembler = VectorAssembler(inputCols=df.columns[1:785], outputCol="features")
    encoder = OneHotEncoder(inputCol="_c0", outputCol="labels", dropLast=False)
    spark_model = SparkAsyncDL(inputCol="features", tensorflowGraph=mg, tfInput="x:0", tfLabel="y:0", tfOutput="out:0",
                               tfOptimizer="adam", miniBatchSize=300, miniStochasticIters=1, shufflePerIter=True, iters=50,
                               predictionCol="predicted", labelCol="labels", partitions=4, verbose=1, optimizerOptions=adam_config)

    p = Pipeline(stages=[vector_assembler, encoder, spark_model]).fit(df)
    p.write().overwrite().save("/tmp/simple_dnn")
    loaded_pipeline = PysparkPipelineWrapper.unwrap(PipelineModel.load("/tmp/simple_dnn"))
    predictions = loaded_pipeline.transform(df)
    evaluator = MulticlassClassificationEvaluator(labelCol="_c0", predictionCol="predicted", metricName="accuracy")
    print("Test Error = %g" % (1.0 - evaluator.evaluate(predictions)))
