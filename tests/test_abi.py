"""The C-ABI boundary on a box without a GPU: the library loads, exports every symbol include/hedgehog_mc.h
declares, the ctypes mirror has the C compiler's struct layout, and the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

import hedgehog_jl_b200 as hh
from hedgehog_jl_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hedgehog_mc.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(hh_[a-z0-9_]+)\s*\(", src))
    names -= {"hh_allreduce_fn"}
    return sorted(names)


def test_header_declares_what_the_mirror_binds():
    assert set(_declared_symbols()) == set(abi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = hh.load_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.hh_version() == abi.HH_VERSION


def test_struct_layout_matches_the_c_compiler(tmp_path):
    structs = {"hh_model": abi.hh_model, "hh_bk_config": abi.hh_bk_config, "hh_sim": abi.hh_sim,
               "hh_payoff": abi.hh_payoff, "hh_result": abi.hh_result, "hh_tangent": abi.hh_tangent,
               "hh_lsm_result": abi.hh_lsm_result, "hh_comm": abi.hh_comm, "hh_path_payoff": abi.hh_path_payoff}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for name, st in structs.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f, _ in st._fields_:
            lines.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, st in structs.items():
        assert int(out[name]) == C.sizeof(st), name
        for f, _ in st._fields_:
            assert int(out[f"{name}.{f}"]) == getattr(st, f).offset, (name, f)


def test_default_bk_config_is_the_reference_defaults():
    lib = hh.load_library()
    c = abi.hh_bk_config()
    lib.hh_default_bk_config(C.byref(c))
    # sample_from_cf.jl:27 (n=5), :50 (h=1e-2), :75 (cf_tol=1e-3), :110-112 (atol=1e-4, 10, 100)
    assert (c.n_std, c.h_fd, c.cf_tol, c.atol, c.maxiter_newton, c.maxiter_bisection) == (5, 1e-2, 1e-3, 1e-4, 10, 100)


def test_no_cpu_fallback_without_a_gpu():
    """Without a usable sm_100 device hh_create must fail and the Python host must raise — never compute on the CPU."""
    code = ("import hedgehog_jl_b200 as hh, sys\n"
            "try:\n    hh.CudaEngine(0)\nexcept hh.HedgehogB200Error as e:\n    print('RAISED', e); sys.exit(0)\n"
            "print('CREATED')\n")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "RAISED" in r.stdout and "no usable CUDA device" in r.stdout


def test_missing_library_raises(tmp_path):
    with pytest.raises(hh.HedgehogB200Error):
        abi.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hedgehog.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "hh_oracle" not in text.replace("oracle/hh_oracle.c", ""), f
