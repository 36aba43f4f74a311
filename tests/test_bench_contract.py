"""The bench JSON lines committed under profiles/ carry every key the benchmark contract names (metric line,
roofline, cpu_baseline, e2e, gpu_launches, clocks; the reference arm's own keys).  CPU only: reads the recorded
lines of the last GPU run, so a change to bench.py that drops a key shows up as soon as the line is refreshed."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, 'profiles', name)
    if not os.path.exists(path):
        pytest.skip(f'{name} not recorded yet')
    rows = [l for l in open(path) if l.startswith('{')]
    assert len(rows) == 1, 'stdout of bench.py must be ONE JSON line'
    return json.loads(rows[0])


BASE = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
        'vs_baseline', 'dtype', 'data', 'config']


@pytest.mark.parametrize('name', ['r01_bench_n1.json', 'r01_bench_n2.json', 'r01_bench_n4.json', 'r01_bench_n8.json',
                                  'r02_bench_n1.json', 'r02_bench_n2.json', 'r02_bench_n4.json', 'r02_bench_n8.json'])
def test_own_arm_line(name):
    d = _line(name)
    for k in BASE + ['roofline', 'e2e', 'gpu_launches', 'clocks']:
        assert k in d, k
    assert d['metric'].startswith('elements assembled/sec') and d['unit'] == 'elements/s' and d['dtype'] == 'f64'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['data'] == 'synthetic'
    assert 'workload' in d['config'] and 'model' not in d['config']
    assert d['warmup'] >= 3 and d['gpu_launches'] == d['steps'] > 0
    r = d['roofline']
    for k in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert k in r, k
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12
    e = d['e2e']
    assert e['unit'] == d['unit'] and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert 0 < e['value'] < d['value']                      # host buffers + PCIe inside the timed region
    c = d['clocks']
    assert c['sm_mhz'] > 0 and c['sm_max_mhz'] >= c['sm_mhz'] and isinstance(c['reasons'], list)
    assert not set(c['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    if d['n_gpus'] == 1:
        b = d['cpu_baseline']
        for k in ('value', 'unit', 'cores', 'kind', 'sample'):
            assert k in b, k
        assert b['kind'] in ('port', 'reference') and b['cores'] >= 1


@pytest.mark.parametrize('name', ['r01_bench_reference.json', 'r02_bench_reference.json'])
def test_reference_arm_line(name):
    d = _line(name)
    for k in BASE + ['impl', 'cpu_baseline', 'e2e']:
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'elements/s'
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['cpu_baseline']['value'] == d['value']
