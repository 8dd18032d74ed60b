"""`models.archs`: RefVSR.py here shadows the reference's; every other arch module (`SPyNet`, `RefVSR_IR`, `edvr_net`,
the `RefVSR_` helper package) resolves to the reference checkout through `__path__` (see ../__init__.py)."""
import os

from .. import REFERENCE_DIR as _ref_models

if _ref_models is not None and os.path.isdir(os.path.join(_ref_models, 'archs')):
    __path__.append(os.path.join(_ref_models, 'archs'))          # noqa: F821
