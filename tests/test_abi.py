"""The C-ABI library loads without a GPU, exports every symbol include/bnmpc.h declares, and refuses to run on the CPU."""
import ctypes as C
import os
import re

import pytest
import torch

import drone_attitude_control_b200 as pkg
from drone_attitude_control_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'bnmpc.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(bnmpc_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f'{n} declared in include/bnmpc.h but not exported by libbnmpc.so'
    assert L.bnmpc_version() == 210


def test_struct_sizes_and_defaults_match_reference_constants():
    cfg = _lib.default_config('force')
    assert cfg.horizon == 30 and cfg.erk_stages == 4 and cfg.sqp_max_iter == 100 and cfg.qp_max_iter == 50
    assert cfg.dt == 1 / 50 and list(cfg.W)[:6] == [100, 100, 1, 1, 0.1, 0.1] and list(cfg.W_e)[:4] == [100, 100, 1, 1]
    assert cfg.lbu[0] == pytest.approx(-0.06429474, abs=1e-12) and cfg.ubu[1] == pytest.approx(0.41791581, abs=1e-12)
    assert (cfg.sim_erk_stages, cfg.sim_substeps, cfg.sim_dt) == (4, 1, 1 / 50)
    cj = _lib.default_config('jerk')
    assert cj.erk_stages == 1 and list(cj.W)[:8] == [100, 100, 1, 1, 0, 0, 0.1, 0.1]
    assert list(cj.lbx)[:6] == [-1.2, -1.2, -1, -1, -5, -5 + 9.81] and cj.ubu[0] == 5
    assert (cj.sim_erk_stages, cj.sim_substeps, cj.sim_dt) == (1, 10, 1 / 500)
    bad = _lib.Config()
    assert _lib.lib().bnmpc_config_default(17, C.byref(bad)) == -1
    assert b'unknown model' in _lib.lib().bnmpc_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback():
    with pytest.raises(pkg.BnmpcError):
        pkg.BatchedAcadosOcpSolver('force')
    cfg = _lib.default_config('force')
    h = C.c_void_p()
    assert _lib.lib().bnmpc_create(C.byref(cfg), 4, 0, C.byref(h)) == -4          # BNMPC_E_CUDA
    assert b'no CUDA device' in _lib.lib().bnmpc_last_error()
