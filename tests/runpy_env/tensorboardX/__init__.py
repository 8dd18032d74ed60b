"""test stand-in for tensorboardX (run.py:36-37 constructs a SummaryWriter on rank 0; nothing is asserted on it)"""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None
