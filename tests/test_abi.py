"""The C-ABI library loads and exports every symbol include/plas.h declares; the ctypes mirror covers
them all; descriptor structs have the sizes the C side expects (no compute: runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "plas.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plas_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from phones_las_b200 import _lib
    return ctypes.CDLL(_lib.LIB_PATH)


def test_header_declares_entry_points():
    names = _header_functions()
    for must in ("plas_frontend_fwd", "plas_gemm_bf16", "plas_gemm_bf16_f32out", "plas_gemm_f32",
                 "plas_bilstm_rec_fwd", "plas_decoder_fwd", "plas_mask_time", "plas_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in _header_functions():
        assert hasattr(lib, name), f"libplas.so does not export {name}"


def test_ctypes_mirror_covers_header():
    from phones_las_b200 import _lib
    assert sorted(_lib.EXPORTS) == _header_functions()


def test_descriptor_layouts_match_c():
    """sizeof/offsetof of the ctypes structs vs the C compiler's view of include/plas.h."""
    import subprocess
    import tempfile
    from phones_las_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "plas.h"
int main(void) {
  printf("%zu %zu %zu\n", sizeof(plas_frontend_desc), sizeof(plas_rec_desc), sizeof(plas_dec_desc));
  printf("%zu %zu %zu %zu\n", offsetof(plas_frontend_desc, window), offsetof(plas_rec_desc, whh_tc),
         offsetof(plas_dec_desc, keys), offsetof(plas_dec_desc, pv_ld));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes[0] == ctypes.sizeof(_lib.FrontendDesc)
    assert sizes[1] == ctypes.sizeof(_lib.RecDesc)
    assert sizes[2] == ctypes.sizeof(_lib.DecDesc)
    assert sizes[3] == _lib.FrontendDesc.window.offset
    assert sizes[4] == _lib.RecDesc.whh_tc.offset
    assert sizes[5] == _lib.DecDesc.keys.offset
    assert sizes[6] == _lib.DecDesc.pv_ld.offset


def test_missing_library_fails_loudly(monkeypatch):
    from phones_las_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libplas.so")
    with pytest.raises(_lib.PlasError):
        _lib.lib()


def test_no_cuda_fails_loudly():
    import torch
    from phones_las_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PlasError):
        _lib.require_cuda()


def test_training_descriptor_layouts_match_c():
    """Every field offset (and the size) of the training-path / fp32-inference descriptors: ctypes mirror vs gcc on plas.h."""
    import subprocess
    import tempfile
    from phones_las_b200 import _lib
    structs = {"plas_gemm_ex_desc": _lib.GemmExDesc, "plas_rec_train_desc": _lib.RecTrainDesc,
               "plas_dec_train_desc": _lib.DecTrainDesc, "plas_dec_infer_desc": _lib.DecInferDesc}
    lines = []
    for cname, cls in structs.items():
        lines.append(f'printf("%zu\\n", sizeof({cname}));')
        for fname, *_ in cls._fields_:
            lines.append(f'printf("%zu\\n", offsetof({cname}, {fname}));')
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "plas.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0;\n}\n"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = [int(x) for x in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    i = 0
    for cname, cls in structs.items():
        assert out[i] == ctypes.sizeof(cls), cname
        i += 1
        for fname, *_ in cls._fields_:
            assert out[i] == getattr(cls, fname).offset, f"{cname}.{fname}"
            i += 1
