"""The C-ABI library builds, loads and exports every symbol include/mcan_b200.h declares; and the
product path fails loudly (no CPU fallback) without a CUDA device."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mcan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcan_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("mcan_build", os.path.join(ROOT, "mcan-vqa_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build(verbose=False)
    from mcan_vqa_b200 import capi
    return capi


def test_header_and_binding_declare_the_same_symbols(lib):
    assert _header_symbols() == sorted(lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    handle = lib.load()
    for name in _header_symbols():
        assert hasattr(handle, name), name
    assert handle.mcan_version() == 1


def test_struct_layouts_match_the_header(lib, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "mcan_b200.h"\n#include <stdio.h>\n#include <stddef.h>\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(mcan_gemm_args), sizeof(mcan_attn_args),'
                   ' sizeof(mcan_attn_bwd_args), offsetof(mcan_gemm_args, bias), offsetof(mcan_gemm_args, stream),'
                   ' offsetof(mcan_attn_bwd_args, dout)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(lib.GemmArgs), ctypes.sizeof(lib.AttnArgs), ctypes.sizeof(lib.AttnBwdArgs),
            lib.GemmArgs.bias.offset, lib.GemmArgs.stream.offset, lib.AttnBwdArgs.dout.offset]
    assert got == want


def test_errors_are_reported_not_thrown(lib):
    handle = lib.load()
    args = lib.GemmArgs()          # all zeros: invalid
    rc = handle.mcan_gemm(ctypes.byref(args))
    assert rc < 0
    assert b"mcan_gemm" in handle.mcan_last_error()
    assert handle.mcan_gemm(None) < 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from mcan_vqa_b200 import ops
    x = torch.randn(4, 64)
    with pytest.raises(lib.McanError):
        ops.layernorm_fwd(x, torch.ones(64), torch.zeros(64), 1e-6, y_f32=torch.empty_like(x))
    with pytest.raises(lib.McanError):
        ops.gemm(x.to(torch.bfloat16), x.to(torch.bfloat16), out_f32=torch.empty(4, 4))


def test_struct_packer_matches_field_by_field_fill(lib):
    """ops.gemm fills mcan_gemm_args with one struct.pack_into; the bytes must equal a ctypes field-by-field fill."""
    import random
    rnd = random.Random(7)
    pk = lib.StructPacker(lib.GemmArgs)
    assert pk.packer.size == ctypes.sizeof(lib.GemmArgs)
    for _ in range(20):
        args = lib.GemmArgs()
        vals = []
        for name, ctype in lib.GemmArgs._fields_:
            if hasattr(ctype, "_length_"):
                arr = getattr(args, name)
                for i in range(ctype._length_):
                    v = rnd.randrange(1, 1 << 47)
                    arr[i] = v
                    vals.append(v)
            elif ctype is ctypes.c_float:
                v = rnd.choice([0.0, 0.1, 1.0, 0.25])
                setattr(args, name, v)
                vals.append(v)
            elif ctype is ctypes.c_void_p:
                v = rnd.choice([0, rnd.randrange(1, 1 << 47)])
                setattr(args, name, v or None)
                vals.append(v)
            elif ctype is ctypes.c_uint32:
                v = rnd.randrange(0, 1 << 32)
                setattr(args, name, v)
                vals.append(v)
            else:
                v = rnd.randrange(0, 1 << 20)
                setattr(args, name, v)
                vals.append(v)
        assert len(vals) == pk.count
        ptr = pk.pack(*vals)
        assert ctypes.string_at(ptr, pk.packer.size) == ctypes.string_at(ctypes.byref(args), ctypes.sizeof(args))
