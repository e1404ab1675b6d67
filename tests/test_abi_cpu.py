"""CPU: the C-ABI library loads and exports every symbol include/pcgrl_b200.h declares; host-only queries
work; launch entry points fail loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import pytest

from control_pcgrl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pcgrl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcgrl_[a-z_]+)\s*\(", src)))


def test_library_built_and_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pcgrl_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "ctypes table and header disagree"


def test_struct_layout_matches_header():
    # sizes computed by gcc for the header's structs (see tests: compiled on the fly)
    import subprocess
    import tempfile
    code = ('#include "pcgrl_b200.h"\n#include <stdio.h>\n#include <stddef.h>\n'
            'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(pcgrl_config), sizeof(pcgrl_state), '
            'sizeof(pcgrl_obs_args), offsetof(pcgrl_config, init_probs), offsetof(pcgrl_config, weights), '
            'offsetof(pcgrl_obs_args, out));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(code)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    want = [ctypes.sizeof(_lib.Config), ctypes.sizeof(_lib.State), ctypes.sizeof(_lib.ObsArgs),
            _lib.Config.init_probs.offset, _lib.Config.weights.offset, _lib.ObsArgs.out.offset]
    assert [int(v) for v in out] == want


def _cfg(problem="binary", rep="narrow", dims=(16, 16), n_tiles=2, n_stats=2):
    c = _lib.Config()
    c.abi_version = _lib.PCGRL_ABI_VERSION
    c.problem = _lib.PROB_IDS[problem]
    c.representation = _lib.REP_IDS[rep]
    c.action_kind = _lib.ACT_INT32
    c.ndim = len(dims)
    for i, d in enumerate(dims):
        c.dims[i] = d
    c.n_tiles, c.n_stats = n_tiles, n_stats
    c.max_iterations, c.max_changes = 769, -1
    return c


def test_host_queries_and_argument_errors():
    lib = _lib.load()
    assert lib.pcgrl_abi_version() == _lib.PCGRL_ABI_VERSION
    c = _cfg()
    assert lib.pcgrl_config_check(c) == 0 and c.row_stride == 256
    assert lib.pcgrl_step_bytes(c) == 2 * 256 + 4 + 8 * 2 + 5          # 537 B
    assert lib.pcgrl_scratch_bytes(c, 1024) >= 0
    z = _cfg("zelda", "turtle", (7, 11), 8, 7)
    assert lib.pcgrl_config_check(z) == 0 and z.row_stride == 80
    assert lib.pcgrl_step_bytes(z) == 2 * 77 + 4 + 56 + 5              # 219 B
    # ABI 5: narrow action elements and packed result records
    assert lib.pcgrl_record_stride(c) == 0
    c8 = _cfg()
    c8.action_elem_bytes, c8.record_stat_bytes = 1, 1
    assert lib.pcgrl_config_check(c8) == 0
    assert lib.pcgrl_record_stride(c8) == 8                            # f32 reward | 2 x u8 stats | done | changed
    assert lib.pcgrl_step_bytes(c8) == 2 * 256 + 1 + 8 * 2 + 5
    z.record_stat_bytes = 2
    assert lib.pcgrl_record_stride(z) == 20                            # 4 + 7 * 2 + 2
    bad = _cfg("zelda", "turtle", (7, 11), 8, 7)
    bad.action_kind, bad.representation, bad.act_h, bad.act_w = _lib.ACT_WIDE_FLAT, _lib.REP_IDS["wide"], 7, 11
    bad.action_elem_bytes = 1                                          # 7 * 11 * 8 = 616 actions do not fit a byte
    assert lib.pcgrl_config_check(bad) == _lib_err("PCGRL_E_ARG")
    bad.action_elem_bytes = 2
    assert lib.pcgrl_config_check(bad) == 0
    bad = _cfg()
    bad.record_stat_bytes = 3
    assert lib.pcgrl_config_check(bad) == _lib_err("PCGRL_E_ARG")
    bad = _cfg()
    bad.n_stats = 5
    assert lib.pcgrl_config_check(bad) == _lib_err("PCGRL_E_ARG")
    assert b"n_stats" in lib.pcgrl_last_error()
    bad = _cfg()
    bad.action_kind = _lib.ACT_CA_TILES
    assert lib.pcgrl_config_check(bad) < 0
    st = _lib.State()
    assert lib.pcgrl_step(c, st, None, None) < 0                       # NULL pointers are rejected, not dereferenced
    with pytest.raises(_lib.PcgrlError):
        _lib.check(-1, "probe")


def _lib_err(name):
    return {"PCGRL_E_ARG": -1, "PCGRL_E_CUDA": -2, "PCGRL_E_UNSUPPORTED": -3}[name]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import control_pcgrl_b200 as P
    with pytest.raises(_lib.PcgrlError):
        P.BatchedPcgrlEnv(P.make_config(), 4)
    with pytest.raises(_lib.PcgrlError):
        P.make("binary-narrow-v0")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "control_pcgrl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} mentions the oracle"
                assert "/root/reference" not in txt or f.endswith((".py",)) and "relative to" in txt or "/root/reference/" in txt
