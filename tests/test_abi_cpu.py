"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/beep_b200.h declares; without a GPU it fails loudly."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_cuda, load_golden


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "beep_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(beep_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from beep_b200 import engine
    L = ctypes.CDLL(engine.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/beep_b200.h but not exported"
    # and the binding knows every one of them
    assert set(names) == set(engine._SIGNATURES)


def test_struct_layout_matches_def():
    """ctypes mirror of BeepModelDesc has the size the C compiler gives it."""
    import subprocess, tempfile
    from beep_b200.model import BeepModelDescStruct
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write('#include <stdio.h>\n#include "beep_b200.h"\nint main(){printf("%zu\\n",sizeof(BeepModelDesc));return 0;}\n')
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        size = int(subprocess.check_output([exe]).decode())
    assert size == ctypes.sizeof(BeepModelDescStruct)


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU behaviour")
def test_no_device_fails_loudly():
    from beep_b200 import engine
    m, _ = load_golden("EX_sim")
    with pytest.raises(engine.BeepError) as e:
        engine.Engine(m)
    assert e.value.code == -4          # BEEP_ERR_NO_DEVICE: there is no CPU fallback


def test_bootstrap_matches_oracle():
    """PMCMC::bootstrap (pmcmc.cc:211-268): host arithmetic of the C-ABI vs the oracle."""
    from beep_b200 import engine
    import oracle_lib
    rng = np.random.default_rng(5)
    for n in (1, 7, 256):
        L = rng.normal(-500, 3, size=n)
        for sec in (0, 3):
            inc_e, back_e = engine.bootstrap(L, 0.7, 99, sec)
            inc_o, back_o = oracle_lib.bootstrap(L, 0.7, 99, sec)
            assert abs(inc_e - inc_o) <= 1e-12 * abs(inc_o)
            assert np.array_equal(back_e, back_o)
            # stay-in-place rule: every selected ancestor keeps its own slot
            for i in set(back_e.tolist()):
                assert back_e[i] == i
