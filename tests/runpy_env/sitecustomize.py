"""TEST INFRASTRUCTURE - loaded by the subprocesses tests/test_dropin_runpy.py starts (this directory is first on their
PYTHONPATH).  It stands in for packages the reference imports but this image lacks, and never ships with the product.

  * torchvision.models.vgg19(pretrained=True) would download weights (attention.py:28): build it with weights=None - the
    VGG parameters are part of the RefVSR checkpoint anyway (attention.py:44-45).
  * REFVSR_TEST_ORACLE_OPS=1 (CPU runs only): the product has no CPU path, so the drop-in's operator set is replaced by the
    oracle's test double and forced to fp32 - this checks the drop-in wiring / schedule through run.py, not the kernels.
"""
import os


def _patch():
    try:
        import torchvision
    except Exception:                                   # noqa: BLE001
        return
    _vgg19 = torchvision.models.vgg19
    torchvision.models.vgg19 = lambda pretrained=False, **kw: _vgg19(weights=None)
    if os.environ.get('REFVSR_TEST_ORACLE_OPS') == '1':
        import refvsr_b200.lib as lib
        import refvsr_b200.network as network
        from oracle.oracle_ops import OracleOps
        lib.CudaOps = OracleOps
        _init = network.Network.__init__

        def init(self, config, ops=None):
            config.b200_precision = 'fp32'
            _init(self, config, ops)
        network.Network.__init__ = init
    prec = os.environ.get('REFVSR_TEST_PRECISION')
    if prec:
        import refvsr_b200.network as network
        _init2 = network.Network.__init__

        def init2(self, config, ops=None):
            config.b200_precision = prec
            _init2(self, config, ops)
        network.Network.__init__ = init2


if os.environ.get('REFVSR_TEST_ENV') == '1':
    _patch()
