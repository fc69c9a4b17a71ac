"""bench.py prints ONE JSON line with the keys the driver's contract names."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
             'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline'}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True,
                         text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run('--impl', 'reference', '--workload', 'tiny', '--steps', '1', '--warmup', '0')
    assert BASE_KEYS <= set(d) and d['impl'] == 'reference'
    assert d['metric'] == 'exposures_per_s' and d['unit'] == 'exposures/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['vs_baseline'] is None and d['scaling'] == 'weak'
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['sample'] and cb['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'exposures/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


@pytest.mark.gpu
def test_native_arm_line():
    d = _run('--workload', 'tiny', '--steps', '6', '--warmup', '3')
    assert BASE_KEYS | {'clocks', 'gpu_launches', 'roofline', 'roofline_hbm'} <= set(d)
    assert d['n_gpus'] == 1 and d['steps'] == 6 and d['warmup'] >= 3 and d['value'] > 0
    assert d['gpu_launches'] >= 6 * 5 and d['dtype'] == 'f64' and d['data'] == 'synthetic'
    e = d['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert abs(e['value'] - d['value']) / d['value'] > 1e-6            # measured separately
    for r in (d['roofline'], d['roofline_hbm']):
        assert {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} <= set(r)
    assert d['roofline_hbm']['bound'] == 'hbm' and d['roofline_hbm']['unit'] == 'GB/s'
    c = d['clocks']
    assert {'sm_mhz', 'sm_max_mhz', 'reasons'} <= set(c)
    assert {'value', 'unit', 'cores', 'kind', 'sample'} <= set(d['cpu_baseline'])


def test_every_rank_gets_physical_inputs():
    """Ranks key their exposures with rank * 100000 + i.  The per-exposure inputs derived from
    that index must stay physical for every rank (a visit-trend factor that went negative for
    ranks > 0 once made every multi-GPU run throw zero electrons and fail its own bookkeeping)."""
    sys.path.insert(0, ROOT)
    import bench
    wk = bench.WORKLOADS['c4']
    for rank in (0, 1, 7):
        for i in (0, 7, 63, 999):
            kw = bench.frame_kwargs(wk, rank * 100000 + i)
            assert 0.98 < kw['scale_factor'] <= 1.0, (rank, i, kw['scale_factor'])
            assert abs(kw['x_ref'] - wk['x_ref']) < 0.5 and abs(kw['y_ref'] - wk['y_ref']) < 0.5
