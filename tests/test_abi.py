"""The C-ABI library loads on a CPU-only box and exports every symbol include/smenv.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from safemotionsrisk_b200 import abi, cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "smenv.h")).read()
    return sorted(set(re.findall(r"\b(smenv_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_match_abi_list():
    assert declared_symbols() == sorted(abi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = cabi.load()
    for sym in declared_symbols():
        assert hasattr(lib, sym), sym


def test_struct_layout_matches_library():
    lib = cabi.load()
    assert lib.smenv_sizeof_scene() == C.sizeof(abi.SmScene)
    assert lib.smenv_sizeof_shape() == C.sizeof(abi.SmShape)
    assert lib.smenv_abi_version() == 1


def test_oracle_shares_the_scene_layout():
    from oracle import oracle
    assert oracle.lib().smo_sizeof_scene() == C.sizeof(abi.SmScene)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from safemotionsrisk_b200.vec_env import SafeMotionsVecEnv
    with pytest.raises(cabi.SmEnvError):
        SafeMotionsVecEnv(num_envs=1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "safemotionsrisk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "smenv_oracle" not in text, f
