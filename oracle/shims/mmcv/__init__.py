"""Minimal stand-in for `mmcv` so that the UNMODIFIED reference (`/root/reference`) imports on CPU.

TEST INFRASTRUCTURE ONLY (oracle side).  mmcv is an un-vendored, unpinned dependency of the
reference (install/install_cudnn113.sh:9-16).  On the RefVSR hot path it contributes no arithmetic
of its own: `ConvModule` (models/archs/SPyNet.py:152-191) is conv2d(+bias) followed by ReLU, and
`load_checkpoint` is only file IO.  Everything else here exists so that
`mmedit/models/common/__init__.py:2-20` can be imported.
"""
__version__ = "0.0.0-oracle-shim"
