"""Shared plumbing of the example scripts: command line, Spark session, the MNIST frame, a throughput report."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from sparkflow_b200 import compat  # noqa: E402

compat.install()          # `sparkflow`, and stand-ins for `pyspark` / `tensorflow` when those are not installed

from pyspark.sql import SparkSession  # noqa: E402
from pyspark.sql.functions import rand  # noqa: E402

from _data import mnist_csv  # noqa: E402

PIXELS = 784


def parse_args(description, iters, partitions=4):
    ap = argparse.ArgumentParser(description=description)
    ap.add_argument("--rows", type=int, default=None, help="train on the first N shuffled rows only")
    ap.add_argument("--iters", type=int, default=iters, help="iterations per partition")
    ap.add_argument("--partitions", type=int, default=partitions)
    ap.add_argument("--quiet", action="store_true", help="no per-iteration loss lines")
    ap.add_argument("--out", default=None, help="where to save the fitted pipeline")
    return ap.parse_args()


def mnist_frame(args, driver_memory="2g"):
    """(spark, DataFrame with columns _c0 = label, _c1.._c784 = pixels), shuffled and optionally truncated."""
    spark = (SparkSession.builder.appName("sparkflow_b200-examples").master(f"local[{args.partitions}]")
             .config("spark.driver.memory", driver_memory).getOrCreate())
    frame = spark.read.option("inferSchema", "true").csv(mnist_csv()).orderBy(rand(seed=1))
    if args.rows:
        frame = frame.limit(args.rows).repartition(args.partitions)
    return spark, frame


def pixel_columns(frame):
    return frame.columns[1:1 + PIXELS]


class Stopwatch:
    def __init__(self, what, rows, iters, batch):
        self.what, self.samples = what, iters * batch
        self.rows = rows

    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        dt = time.perf_counter() - self.t0
        if exc[0] is None:
            print(f"[{self.what}] {self.rows} rows, fit took {dt:.2f} s")
