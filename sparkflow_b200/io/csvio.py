"""Numeric CSV reader (native parser with a numpy fallback)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def read_numeric_csv(path: str, has_header: bool = False) -> Tuple[np.ndarray, Optional[List[str]]]:
    names: Optional[List[str]] = None
    if has_header:
        with open(path) as fh:
            names = [c.strip() for c in fh.readline().rstrip("\n").split(",")]
    try:
        from ..ops.native import host_ext

        arr = host_ext().read_csv(path, 1 if has_header else 0)
        return np.asarray(arr, dtype=np.float64), names
    except Exception:
        return np.atleast_2d(np.loadtxt(path, delimiter=",", skiprows=1 if has_header else 0, dtype=np.float64)), names
