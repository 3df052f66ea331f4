"""Read the reference's golden pickles without CasADi / do_mpc.

TEST INFRASTRUCTURE ONLY (see oracle/README.md). The reference stores its regression
results with ``do_mpc.data.save_results`` (reference do_mpc/data.py:376-434): a dict
{'mpc': MPCData, 'simulator': Data, 'estimator': Data}.  The Data objects pickle CasADi
Functions/structs that cannot be reconstructed here, so every class that is not a plain
numpy / builtin type is replaced by a stub that just keeps its state.
"""
import io
import pickle

import numpy as np

_SAFE_PREFIXES = ("numpy", "builtins", "collections", "copyreg", "_codecs", "pandas")


class _Stub:
    def __init__(self, *a, **k):
        self._args = a
        self._kwargs = k

    def __setstate__(self, state):
        self._state = state

    def __reduce_ex__(self, protocol):  # pragma: no cover - never re-pickled
        raise TypeError("stub objects are read-only")


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in [p.split(".")[0] for p in _SAFE_PREFIXES]:
            return super().find_class(module, name)
        return type(name, (_Stub,), {"__module__": module})


def load_results(path):
    """Return {'mpc': {field: ndarray}, 'simulator': {...}, 'estimator': {...}}."""
    with open(path, "rb") as f:
        raw = _StubUnpickler(f).load()
    out = {}
    for key, obj in raw.items():
        state = getattr(obj, "_state", None)
        if state is None:
            state = getattr(obj, "__dict__", {})
        out[key] = state
    return out


def numeric_fields(state):
    return {k: v for k, v in state.items() if isinstance(v, np.ndarray)}
