"""The C-ABI library builds, loads and exports every symbol include/vt_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from tests.conftest import ROOT


def test_library_exports_every_declared_symbol():
    from videotransformer_pytorch_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    hdr = open(os.path.join(ROOT, 'include', 'vt_b200.h')).read()
    declared = sorted(set(re.findall(r'^int\s+(vt_\w+)\s*\(', hdr, flags=re.M)))
    assert len(declared) >= 15
    dll = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(dll, name), f'{name} declared in vt_b200.h but not exported'
    assert sorted(_lib.EXPORTS) == declared
    assert dll.vt_version() == 1


def test_sass_is_blackwell_native():
    import shutil
    import subprocess
    from videotransformer_pytorch_b200 import build
    if not shutil.which('cuobjdump'):
        return
    sass = subprocess.run(['cuobjdump', '-sass', build.build()], capture_output=True, text=True).stdout
    assert 'UTCHMMA' in sass        # tcgen05.mma
    assert 'UTMALDG' in sass        # TMA loads
    assert 'LDTM' in sass           # tcgen05.ld
    assert 'HMMA.16816' not in sass  # no legacy mma.sync path


def test_missing_library_fails_loudly(monkeypatch):
    import pytest
    from videotransformer_pytorch_b200 import _lib
    monkeypatch.setattr(_lib, '_dll', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libvt_b200.so')
    with pytest.raises(RuntimeError, match='no CPU / library fallback'):
        _lib.load_library()
