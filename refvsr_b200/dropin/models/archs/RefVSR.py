from refvsr_b200.network import Network  # noqa: F401  (replaces models/archs/RefVSR.py:14-325)
