"""The C-ABI library loads and exports every symbol include/refvsr_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'refvsr_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(rv_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from refvsr_b200.lib import LIB_PATH, SIGNATURES
    lib = ctypes.CDLL(LIB_PATH)
    names = header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
    assert sorted(SIGNATURES) == names, 'python prototypes and header are out of sync'
    lib.rv_version.restype = ctypes.c_int
    assert lib.rv_version() == 100


def test_descriptor_struct_matches_header():
    """field order of rv_conv_desc in the ctypes mirror == the header's struct"""
    from refvsr_b200.lib import rv_conv_desc
    src = open(os.path.join(ROOT, 'include', 'refvsr_b200.h')).read()
    body = src[src.index('typedef struct rv_conv_desc {'):src.index('} rv_conv_desc;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = []
    for decl in body.split('{', 1)[1].split(';'):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if not decl.startswith('const') else decl.split(None, 2)[2]
        for nm in names.split(','):
            fields.append(nm.replace('*', '').strip())
    assert fields == [f[0] for f in rv_conv_desc._fields_]


def test_invalid_arguments_return_error_codes_without_a_gpu():
    from refvsr_b200.lib import load_library
    lib = load_library()
    rc = lib.rv_conv2d(None, None)
    assert rc == -1 and b'null descriptor' in lib.rv_last_error()
    rc = lib.rv_conf_max(None, None, None, 0, None)
    assert rc == -1
