"""Reference import path ``sparkflow.RWLock`` -> :class:`sparkflow_b200.parallel.rwlock.RWLock`."""
from .parallel.rwlock import RWLock  # noqa: F401
