"""CPU checks of the drop-in boundary: the C-ABI library builds/loads and exports exactly
the symbols include/videoblip_b200.h declares; the ctypes struct layouts match the header;
the product path fails loudly without CUDA (no fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "videoblip_b200.h").read_text()


@pytest.fixture(scope="module")
def lib():
    from eilev_b200 import _lib
    if not _lib.LIB_PATH.exists():
        _lib.build()
    return _lib.lib()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", body)))


def test_header_symbols_are_exported_and_bound(lib):
    from eilev_b200 import _lib
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)
    assert lib.vb_abi_version() == int(re.search(r"#define VB_ABI_VERSION (\d+)", HEADER).group(1))


def _c_struct_fields(name):
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HEADER, flags=re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.replace("*", " ").split()
        # "int64_t m, n, k" -> m n k ; "const void* a" -> a ; "vb_attn_args fwd" -> fwd
        tail = decl.split(",")
        first = tail[0].replace("*", " ").split()[-1]
        fields.append(first)
        for t in tail[1:]:
            fields.append(t.replace("*", " ").split()[-1])
        del names
    return fields


@pytest.mark.parametrize("cname,pyname", [("vb_gemm_args", "GemmArgs"), ("vb_attn_args", "AttnArgs"),
                                           ("vb_attn_bwd_args", "AttnBwdArgs")])
def test_ctypes_structs_follow_the_header(cname, pyname):
    from eilev_b200 import _lib
    py = [f[0] for f in getattr(_lib, pyname)._fields_]
    assert py == _c_struct_fields(cname)


def test_struct_sizes():
    from eilev_b200 import _lib
    assert C.sizeof(_lib.GemmArgs) == 5 * 8 + 7 * 8 + 2 * 4 + 2 * 8 + 4 * 4 + 8 + 16
    assert C.sizeof(_lib.AttnArgs) == 6 * 8 + 13 * 8 + 8 + 8 + 16 + 16  # + rel_bias, rel_bias_stride
    assert C.sizeof(_lib.AttnBwdArgs) == C.sizeof(_lib.AttnArgs) + 4 * 8 + 6 * 8 + 2 * 8 + 8


def test_error_reporting_without_gpu(lib):
    from eilev_b200 import _lib
    args = _lib.GemmArgs()
    assert lib.vb_gemm(C.byref(args), None) != 0
    assert b"vb_gemm" in lib.vb_last_error()
    assert lib.vb_adamw(None, None, None, None, 0, 0.0, 0.0, 0.0, 0.0, 0.0, 0, None, None) != 0
    assert b"step" in lib.vb_last_error()


def test_product_path_refuses_cpu_tensors():
    from eilev_b200 import _lib, ops
    with pytest.raises(_lib.VbError):
        ops.gemm(torch.zeros(2, 8, dtype=torch.bfloat16), torch.zeros(4, 8, dtype=torch.bfloat16))
    with pytest.raises(_lib.VbError):
        ops.layernorm(torch.zeros(2, 8, dtype=torch.bfloat16), torch.ones(8), torch.zeros(8), 1e-5)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under eilev_b200/ may reference it."""
    for path in (ROOT / "eilev_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_patch_gather_u8_validates_arguments_before_touching_the_device(lib):
    """Argument marshalling of the uint8 frame entry point (double rescale, host float arrays):
    a non-positive std / null pointers / too many channels are refused with a message and no launch."""
    three = (C.c_float * 3)
    buf = (C.c_uint8 * 64)()
    out = (C.c_uint16 * 64)()
    ok_mean, ok_std = three(0.5, 0.5, 0.5), three(0.25, 0.25, 0.25)
    args = (1, 3, 1, 4, 4, 2, 12, 1 / 255)
    assert lib.vb_patch_gather_u8(C.addressof(buf), C.addressof(out), *args, ok_mean, three(0.25, 0.0, 0.25), None) != 0
    assert b"std must be positive" in lib.vb_last_error()
    assert lib.vb_patch_gather_u8(None, C.addressof(out), *args, ok_mean, ok_std, None) != 0
    assert b"bad arguments" in lib.vb_last_error()
    assert lib.vb_patch_gather_u8(C.addressof(buf), C.addressof(out), 1, 5, 1, 4, 4, 2, 20, 1 / 255, ok_mean, ok_std, None) != 0
    assert lib.vb_patch_gather_u8(C.addressof(buf), C.addressof(out), 1, 3, 1, 4, 4, 2, 11, 1 / 255, ok_mean, ok_std, None) != 0  # kpad < C*P*P
