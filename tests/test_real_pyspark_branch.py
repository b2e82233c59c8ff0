"""The code paths that bind the API classes to a GENUINE pyspark (`spark/backend.py: REAL_PYSPARK`) never run in this image
(no pyspark, no JVM).  This test puts a minimal look-alike `pyspark` package (no `__sparkflow_shim__` marker; its classes are
distinct subclasses, plus `JavaMLWriter` / `JavaMLReader` / an RDD that only offers `glom().collect()`) on `sys.path` in a
subprocess and checks that every import of the real branch resolves, that the Estimator / Model derive from ITS classes, that
persistence goes through `JavaMLWriter` / `JavaMLReader` and the JVM carrier builder, and that partitions of a genuine-style
RDD are collected.  It cannot prove interoperability with Spark itself - only that the branch is wired and importable."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FAKE = {
    "pyspark/__init__.py": """
        from sparkflow_b200.spark import context as _c
        class SparkContext(_c.SparkContext):
            pass
        keyword_only = _c.keyword_only
    """,
    "pyspark/ml/__init__.py": """
        from sparkflow_b200.spark.ml import base as _b
        class Estimator(_b.Estimator):
            pass
        class Model(_b.Model):
            pass
        class Pipeline(_b.Pipeline):
            pass
        class PipelineModel(_b.PipelineModel):
            pass
    """,
    "pyspark/ml/feature.py": """
        from sparkflow_b200.spark.ml import feature as _f
        class StopWordsRemover(_f.StopWordsRemover):
            pass
    """,
    "pyspark/ml/linalg.py": "from sparkflow_b200.spark.ml.linalg import Vectors\n",
    "pyspark/ml/param/__init__.py": "from sparkflow_b200.spark.ml.param import Param, Params, TypeConverters\n",
    "pyspark/ml/param/shared.py": "from sparkflow_b200.spark.ml.param import HasInputCol, HasLabelCol, HasPredictionCol\n",
    "pyspark/ml/util.py": """
        from sparkflow_b200.spark.ml.base import MLReadable, MLReader, MLWritable, MLWriter
        from sparkflow_b200.spark.ml.param import Identifiable
        class JavaMLWriter(object):
            def __init__(self, instance):
                self.instance = instance
        class JavaMLReader(object):
            def __init__(self, clazz):
                self.clazz = clazz
    """,
    "pyspark/sql/__init__.py": "from sparkflow_b200.spark.sql import Row\n",
}

SCRIPT = """
    import sys
    import numpy as np
    import sparkflow_b200.spark.backend as B
    assert B.REAL_PYSPARK, "backend did not select the pyspark found on sys.path"
    import pyspark, pyspark.ml, pyspark.ml.feature, pyspark.ml.util
    from sparkflow_b200.tensorflow_async import SparkAsyncDL, SparkAsyncDLModel
    from sparkflow_b200.pipeline_util import PysparkObjId, PysparkReaderWriter
    from sparkflow_b200.HogwildSparkModel import collect_partitions
    assert issubclass(SparkAsyncDL, pyspark.ml.Estimator) and issubclass(SparkAsyncDLModel, pyspark.ml.Model)
    assert PysparkObjId._getCarrierClass() is pyspark.ml.feature.StopWordsRemover
    est = SparkAsyncDL(inputCol='features', tensorflowGraph='{}', tfInput='x:0', tfLabel='y:0', tfOutput='out:0', iters=3)
    assert est.getIters() == 3 and est.getAqcuireLock() is False
    w = est.write()
    assert isinstance(w, pyspark.ml.util.JavaMLWriter) and w.instance is est
    r = SparkAsyncDLModel.read()
    assert isinstance(r, pyspark.ml.util.JavaMLReader) and r.clazz is pyspark.ml.feature.StopWordsRemover
    try:
        est._to_java()
    except (ImportError, AttributeError) as exc:          # no pyspark.ml.wrapper / no JVM gateway in the look-alike
        assert 'wrapper' in str(exc) or '_gateway' in str(exc) or '_active_spark_context' in str(exc), exc
    else:
        raise AssertionError('_to_java should need a JVM')
    class GenuineStyleRDD:                                 # no .partitions(): only the public pyspark API
        def glom(self):
            return self
        def collect(self):
            return [[(np.zeros(2), 0.0)], [(np.ones(2), 1.0), (np.ones(2), 1.0)]]
    parts = collect_partitions(GenuineStyleRDD())
    assert [len(p) for p in parts] == [1, 2]
    print('real-pyspark branch wired')
"""


def test_api_classes_bind_to_a_genuine_looking_pyspark(tmp_path):
    for rel, body in FAKE.items():
        path = tmp_path / rel
        path.parent.mkdir(parents=True, exist_ok=True)
        path.write_text(textwrap.dedent(body))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(tmp_path), ROOT]))
    env.pop("SPARKFLOW_FORCE_SHIM", None)
    out = subprocess.run([sys.executable, "-c", textwrap.dedent(SCRIPT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0 and "real-pyspark branch wired" in out.stdout, out.stdout[-3000:]
    # and the switch back: SPARKFLOW_FORCE_SHIM=1 ignores an installed pyspark
    env["SPARKFLOW_FORCE_SHIM"] = "1"
    out = subprocess.run([sys.executable, "-c", "import sparkflow_b200.spark.backend as B; assert not B.REAL_PYSPARK; print('shim forced')"],
                         env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0 and "shim forced" in out.stdout, out.stdout[-2000:]
