"""Loader for the in-tree native extensions.

``_C``    – sm_100a kernels + C++ step-plan runtime (needs a B200 at run time, imports anywhere)
``_host`` – CPU-only native helpers (TF bundle codec, crc32c, carrier codec, CSV reader)

On a GPU box the CUDA path is THE path: if ``_C.so`` is missing we fail loudly instead of silently
falling back to eager PyTorch.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path
from types import ModuleType

_PKG_DIR = Path(__file__).resolve().parents[1]
_cache: dict[str, ModuleType] = {}


class NativeExtensionMissing(RuntimeError):
    pass


def _load(name: str) -> ModuleType:
    if name in _cache:
        return _cache[name]
    so = _PKG_DIR / f"{name}.so"
    if not so.exists():
        if os.environ.get("SPARKFLOW_NO_AUTOBUILD") != "1":
            try:
                sys.path.insert(0, str(_PKG_DIR.parent))
                from tools.build_ext import build  # type: ignore

                build(only="cuda" if name == "_C" else "host")
            except Exception as exc:  # pragma: no cover - build environment specific
                raise NativeExtensionMissing(
                    f"native extension {so} is missing and could not be built: {exc}. "
                    "Run `python tools/build_ext.py`."
                ) from exc
            finally:
                sys.path.pop(0)
        if not so.exists():
            raise NativeExtensionMissing(f"native extension {so} is missing; run `python tools/build_ext.py`")
    full = f"sparkflow_b200.{name}"
    spec = importlib.util.spec_from_file_location(full, so)
    assert spec and spec.loader
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[full] = mod
    _cache[name] = mod
    return mod


def cuda_ext() -> ModuleType:
    """The sm_100a kernel/runtime module (``sparkflow_b200._C``)."""
    return _load("_C")


def host_ext() -> ModuleType:
    """The CPU-only native helper module (``sparkflow_b200._host``)."""
    return _load("_host")


def ptr(t) -> int:
    """Device address of a torch tensor (0 for ``None``)."""
    return 0 if t is None else int(t.data_ptr())


def current_stream() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)
