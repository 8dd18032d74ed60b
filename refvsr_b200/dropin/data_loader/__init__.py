"""Drop-in for the reference's `data_loader` package (frame IO / window assembly).

`data_loader/datasets.py:222-316` (`Test_datasets.__getitem__`) re-opens and re-decodes ALL frames of a window for
every output frame: 4 streams x T frames (LR, Ref-W, Ref-T and the 1080p ground truth), i.e. 4T decodes where 4 are
new.  With `<repo>/refvsr_b200/dropin` in front of the reference checkout on sys.path this package takes the place of
`data_loader`; every sub-module except `utils` still comes from the reference (its directory is appended to
`__path__`), and `utils.read_frame` keeps the decoded frames of the last few windows (see utils.py).  The tensors
the datasets return are bit-identical (tests/test_loader_dropin.py)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(sys.path) + [os.getcwd()]:
    _cand = os.path.join(_p or os.getcwd(), 'data_loader')      # '' on sys.path = the current directory
    if os.path.isdir(_cand) and os.path.abspath(_cand) != _here and os.path.isfile(os.path.join(_cand, 'datasets.py')):
        __path__.append(_cand)          # noqa: F821  (package attribute): the reference's modules resolve from here
        REFERENCE_DIR = _cand
        break
else:
    raise ImportError('refvsr_b200.dropin.data_loader: the reference checkout (data_loader/datasets.py) is not on sys.path')
