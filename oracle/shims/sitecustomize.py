"""Loaded automatically by subprocesses that run the unmodified reference scripts with oracle/shims on PYTHONPATH.

The reference was written for pandas 1.x / numpy 1.23; this image has pandas 3 / numpy 2.  Restored here, without touching the
reference's files (the same five behaviours tests/golden/make_golden_isoforms.py restores in-process):
  * DataFrame.append (removed in pandas 2.0);
  * assigning a string into a float column upcasts to object instead of raising;
  * Series.values is writable;
  * text columns are object dtype, not the new str dtype;
  * np.array of ragged per-contig lists (bin/pipeline.py:385) gives an object array.
Only active when MGCOD_REFERENCE_SHIMS=1 is set, so an accidental PYTHONPATH leak changes nothing.
"""
import os

if os.environ.get("MGCOD_REFERENCE_SHIMS") == "1":
    import numpy as np
    import pandas as pd

    def _append(self, other, ignore_index=False, verify_integrity=False, sort=False):
        def row(o):
            return o.to_frame().T if isinstance(o, pd.Series) else o
        objs = [self] + ([row(o) for o in other] if isinstance(other, (list, tuple)) else [row(other)])
        return pd.concat(objs, ignore_index=ignore_index, sort=sort)

    pd.DataFrame.append = _append
    pd.options.future.infer_string = False
    from pandas.core.internals import blocks as _blocks
    _orig = _blocks.Block.coerce_to_target_dtype
    _blocks.Block.coerce_to_target_dtype = lambda self, other, raise_on_upcast=False: _orig(self, other, raise_on_upcast=False)
    _values = pd.Series.values

    def _writable_values(self):
        a = _values.fget(self)
        if isinstance(a, np.ndarray) and not a.flags.writeable:
            try:
                a.flags.writeable = True
            except ValueError:
                a = a.copy()
        return a

    pd.Series.values = property(_writable_values)
    _np_array = np.array

    def _array(x, *a, **kw):
        try:
            return _np_array(x, *a, **kw)
        except ValueError:
            o = np.empty(len(x), dtype=object)
            for i, v in enumerate(x):
                o[i] = v
            return o

    np.array = _array
