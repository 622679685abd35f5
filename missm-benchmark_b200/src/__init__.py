"""Drop-in `src` package: provides `src.model` (the B200 path).  Any other `src` directory found on
sys.path (the reference's own, with `src.dataset` / `src.utils`) is appended to this package's search
path, so `train_ddp.py` / `test.py` keep importing their loaders unchanged."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(sys.path):
    _cand = os.path.join(_p or ".", "src")
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and _cand not in __path__:
        __path__.append(_cand)
