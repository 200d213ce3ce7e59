"""The C-ABI library loads and exports every symbol include/ssba.h declares; no compute calls
here (no GPU in the CPU suite).  Also: the product has no CPU fallback and never touches oracle/."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ssba.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ssba_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(ssba_lib):
    from ssvio_b200 import ba
    names = declared_symbols()
    assert set(names) == set(ba.ABI_SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", ba.LIB_PATH], capture_output=True, text=True,
                         check=True).stdout
    exported = set(re.findall(r"\bT (ssba_[a-z0-9_]+)", out))
    for n in names:
        assert n in exported, f"{n} declared in ssba.h but not exported by libssba.so"
        assert getattr(ssba_lib, n) is not None


def test_header_is_plain_c():
    """include/ssba.h must compile as C (no C++ / torch types at the boundary)."""
    src = '#include "ssba.h"\nint main(void){ssba_options o; ssba_default_options(&o); return (int)sizeof(ssba_report) == 0;}\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-x", "c", "-"], input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_struct_layouts_match_header(ssba_lib):
    """ctypes mirrors vs sizeof() from the real header."""
    from ssvio_b200 import ba
    src = ('#include <stdio.h>\n#include "ssba.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(ssba_options),'
           'sizeof(ssba_iter_record), sizeof(ssba_report), sizeof(ssba_profile), sizeof(ssba_problem_info), sizeof(ssba_status));return 0;}\n')
    exe = "/tmp/ssba_sizeof"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe], input=src, text=True, check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes[:5] == [C.sizeof(ba.Options), C.sizeof(ba.IterRecord), C.sizeof(ba.Report), C.sizeof(ba.Profile),
                         C.sizeof(ba.ProblemInfo)]
    from oracle import bindings
    assert C.sizeof(bindings.Report) == sizes[2]


def test_default_options_mirror_levenberg_constants(ssba_lib):
    """g2o/core/optimization_algorithm_levenberg.cpp:44-56"""
    from ssvio_b200 import ba
    o = ba.Options()
    ssba_lib.ssba_default_options(C.byref(o))
    assert o.tau == 1e-5
    assert o.good_step_lower_scale == pytest.approx(1 / 3) and o.good_step_upper_scale == pytest.approx(2 / 3)
    assert o.max_trials_after_failure == 10 and o.user_lambda_init == 0.0
    assert o.world_size == 1 and o.jacobian_mode == 0


def _cuda_device_present():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20)
        return out.returncode == 0 and "GPU" in out.stdout
    except Exception:
        return False


def test_no_cpu_fallback_without_device(ssba_lib):
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    if _cuda_device_present():
        pytest.skip("a CUDA device is present")
    from ssvio_b200 import ba
    with pytest.raises(ba.SsbaError) as ei:
        ba.BundleAdjuster()
    assert ei.value.status == 3  # SSBA_ERR_NO_DEVICE
    assert "no CPU path" in str(ei.value)


def test_product_never_references_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "ssvio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in (r"from\s+oracle", r"import\s+oracle", r"oracle/", r"ssba_oracle", r"ssba_ref", r"_ref/"):
                    assert not re.search(pat, text), f"{f} references the oracle ({pat})"
    out = subprocess.run(["ldd", os.path.join(pkg, "lib", "libssba.so")], capture_output=True, text=True).stdout
    assert "ssba_oracle" not in out and "ssba_ref" not in out
