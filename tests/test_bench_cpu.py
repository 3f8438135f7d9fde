"""bench.py on a machine without a GPU: the reference arm (the oracle's CPU port, the one other place that may execute
oracle/) prints the contract's JSON line; the product arm refuses to run — there is no CPU path to fall back to."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '3')
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d['impl'] == 'reference'
    assert d['metric'] == 'mppi_sample_steps_per_sec' and d['unit'] == 'sample-steps/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['warmup'] == 3
    assert d['config']['workload'].startswith('push_rex_extract_K8192_T40')
    assert d['value'] > 0 and abs(d['value'] - 8192 * 40 / (d['ms_per_step'] * 1e-3)) < 1e-6 * d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['gpu_launches'] == 0


def test_reference_arm_runs_on_rank_zero_only():
    r = _run('--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '3', env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is visible: the product arm would run")
    r = _run('--steps', '1', '--warmup', '3')
    assert r.returncode != 0
    assert 'no CPU path' in (r.stderr + r.stdout)
