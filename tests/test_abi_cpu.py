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
    # + ln_stats, ln_colsum, (ln_eps, reserved3), stats_out, stats_zero (ABI 5)
    assert C.sizeof(_lib.GemmArgs) == 5 * 8 + 7 * 8 + 2 * 4 + 2 * 8 + 4 * 4 + 8 + 16 + 5 * 8
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


def test_resize_coefficient_tables_equal_the_pillow_restatement(lib):
    """vb_resize_bicubic_coeffs is host arithmetic (no device work): its window bounds and 22-bit
    fixed-point weights must equal the oracle's restatement of Pillow's precompute_coeffs +
    normalize_coeffs_8bpc — itself pinned bit-exactly to PIL.Image.resize — for every size pair."""
    import numpy as np
    from eilev_b200 import ops
    from oracle import pil_resize_ref as P
    pairs = [(448, 224), (224, 224), (100, 224), (160, 224), (300, 224), (500, 224), (1080, 224), (1920, 224),
             (37, 56), (53, 56), (17, 64), (19, 48), (225, 224), (223, 224), (7, 3), (3, 7), (1, 5), (5, 1)]
    for n_in, n_out in pairs:
        ksize, bounds, kk = ops.resize_coeffs(n_in, n_out)
        rk, rb, rkk = P.precompute_coeffs(n_in, n_out)
        assert ksize == rk, (n_in, n_out)
        assert np.array_equal(bounds.numpy(), rb), (n_in, n_out)
        assert np.array_equal(kk.numpy(), rkk), (n_in, n_out)
        assert int((bounds[:, 0] + bounds[:, 1]).max()) <= n_in and int(bounds[:, 0].min()) >= 0
    with pytest.raises(ValueError):
        ops.resize_coeffs(0, 4)
    assert lib.vb_resize_bicubic_coeffs(8, 4, None, None, 0) != 0
    assert lib.vb_resize_u8_pass(None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, None) != 0


def _emulate_resize_pass(src, out, bounds, kk, planes, lines, out_len, ksize, ips, ils, ies, ops_, ols, oes,
                         lines_fastest):
    """resize_u8_pass_kernel (csrc/elementwise.cu) thread for thread, in numpy on flat byte arrays."""
    import numpy as np
    idx = np.arange(planes * lines * out_len, dtype=np.int64)
    per_plane = lines * out_len
    p, r = idx // per_plane, idx % per_plane
    l = np.where(lines_fastest, r % lines, r // out_len)
    o = np.where(lines_fastest, r // lines, r % out_len)
    acc = np.full(idx.shape, 1 << 21, dtype=np.int32)
    for x in range(ksize):
        live = x < bounds[o, 1]
        pos = p * ips + l * ils + (bounds[o, 0].astype(np.int64) + x) * ies
        tap = src[np.where(live, pos, 0)].astype(np.int32)
        acc = acc + np.where(live, tap * kk[o, x], 0).astype(np.int32)
    out[p * ops_ + l * ols + o * oes] = np.clip(acc >> 22, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("shape,size", [((2, 3, 40, 72), (56, 56)), ((1, 2, 90, 60), (32, 48)), ((3, 56, 80), (56, 56)),
                                        ((2, 70, 56), (56, 56)), ((1, 448, 448), (224, 224)), ((2, 56, 56), (56, 56))])
def test_resize_plan_emulated_on_cpu_equals_pillow(lib, shape, size):
    """The pass plan ops.resize_bicubic_u8 hands to vb_resize_u8_pass (offsets, strides, shifted
    vertical windows, warp index order), executed by a numpy emulation of the kernel's per-thread
    formula with the C library's own coefficient tables, reproduces PIL.Image.resize bit for bit."""
    import numpy as np
    from PIL import Image
    from eilev_b200 import ops
    rs = np.random.RandomState(sum(shape))
    frames = rs.randint(0, 256, shape).astype(np.uint8)
    in_h, in_w = shape[-2:]
    out_h, out_w = size
    planes = int(np.prod(shape[:-2]))
    cur = frames.reshape(-1)
    for ps in ops.resize_plan(in_h, in_w, out_h, out_w):
        ksize, bounds, kk = ops.resize_coeffs(*ps["axis"])
        bounds = bounds.numpy().copy()
        bounds[:, 0] -= ps["shift"]
        out = np.zeros(planes * ps["out_shape"][0] * ps["out_shape"][1], dtype=np.uint8)
        _emulate_resize_pass(cur[ps["in_offset"]:], out, bounds, kk.numpy(), planes, ps["lines"], ps["out_len"], ksize,
                             *ps["in_strides"], *ps["out_strides"], ps["lines_fastest"])
        cur = out
    got = cur.reshape(*shape[:-2], out_h, out_w)
    want = np.stack([np.asarray(Image.fromarray(pl).resize((out_w, out_h), resample=Image.BICUBIC))
                     for pl in frames.reshape(-1, in_h, in_w)]).reshape(got.shape)
    assert np.array_equal(got, want)
