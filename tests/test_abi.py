"""CPU: the C-ABI library loads and exports every symbol include/tampc_mppi.h declares; the ctypes mirror of the
structs has the sizes the C compiler gives; without a device the library fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT
from tampc_b200 import _lib, build

HEADER = os.path.join(ROOT, 'include', 'tampc_mppi.h')


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_lib.lib_path()):
        build.build()
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    return re.findall(r'TAMPC_API\s+[\w\s\*]+?\b(tampc_(?:mppi|gp)_\w+)\s*\(', src)


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(set(names)) == sorted(set(_lib.EXPORTS))
    assert lib.tampc_mppi_abi_version() == _lib.ABI_VERSION


def test_ctypes_structs_match_the_c_layout(tmp_path):
    prog = tmp_path / 'sizes.c'
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "tampc_mppi.h"\nint main(void){'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(tampc_mlp_desc), sizeof(tampc_model_desc),'
                    'sizeof(tampc_cost_desc), sizeof(tampc_mppi_config), sizeof(tampc_partial),'
                    'offsetof(tampc_cost_desc, set_goals), offsetof(tampc_model_desc, pendulum), offsetof(tampc_mppi_config, seed),'
                    'sizeof(tampc_gp_desc), offsetof(tampc_gp_desc, lengthscale), offsetof(tampc_gp_desc, rollout_var_cost));return 0;}')
    exe = tmp_path / 'sizes'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(prog), '-o', str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.MlpDesc), C.sizeof(_lib.ModelDesc), C.sizeof(_lib.CostDesc), C.sizeof(_lib.MppiConfig),
            C.sizeof(_lib.Partial), _lib.CostDesc.set_goals.offset, _lib.ModelDesc.pendulum.offset,
            _lib.MppiConfig.seed.offset, C.sizeof(_lib.GpDesc), _lib.GpDesc.lengthscale.offset,
            _lib.GpDesc.rollout_var_cost.offset]
    assert got == want


def test_create_without_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is visible; the no-device path is exercised on the CPU box")
    cfg = _lib.MppiConfig()
    cfg.abi_version, cfg.nx, cfg.nu, cfg.num_samples, cfg.num_local, cfg.max_horizon = 1, 5, 3, 8, 8, 8
    cfg.precision, cfg.lambda_ = 1, 1.0
    h = C.c_void_p()
    rc = lib.tampc_mppi_create(C.byref(cfg), C.byref(h))
    assert rc == _lib.ERR_NO_DEVICE
    assert b'no CPU fallback' in lib.tampc_mppi_last_error(None)


def test_create_rejects_bad_arguments(lib):
    cfg = _lib.MppiConfig()
    cfg.abi_version = 99
    h = C.c_void_p()
    assert lib.tampc_mppi_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID
    cfg.abi_version, cfg.nx, cfg.nu = 1, 99, 3
    assert lib.tampc_mppi_create(C.byref(cfg), C.byref(h)) == _lib.ERR_INVALID


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure; nothing under tampc_b200/ may import it."""
    for root, _, files in os.walk(os.path.join(ROOT, 'tampc_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f
                assert 'mppi_oracle' not in src, f


def test_descriptor_packing_of_golden_specs():
    from conftest import GOLDEN_CASES, load_golden
    for case in GOLDEN_CASES:
        z, spec = load_golden(case)
        md, params = _lib.model_desc(spec.dynamics)
        cd = _lib.cost_desc(spec.cost)
        cfg = _lib.mppi_config(spec.sampler, spec.dynamics.nx, 'mixed')
        assert md.nx == spec.dynamics.nx and cd.nu == spec.cost.nu and cfg.num_samples == spec.sampler.K
        if spec.dynamics.kind == 1:
            assert params.size > 3000 and md.net.dims[0] == 5 and md.decoder.dims[3] == 25
