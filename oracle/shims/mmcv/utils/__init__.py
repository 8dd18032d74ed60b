import logging


def get_logger(name, log_file=None, log_level=logging.INFO, **kw):
    return logging.getLogger(name)


def collect_env():
    return {}


def get_git_hash(*a, **k):
    return 'unknown'


def print_log(*a, **k):
    pass


class Registry:
    def __init__(self, name, *a, **k):
        self.name = name

    def register_module(self, *a, **k):
        def deco(cls):
            return cls
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return deco


def build_from_cfg(*a, **k):
    raise NotImplementedError
