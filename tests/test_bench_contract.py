"""bench.py keeps the driver's contract: ONE JSON line on stdout with the agreed keys, for our arm
(GPU) and for the reference arm (CPU port of the path, `--impl reference`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
        'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline']


def run(args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, 'exactly one JSON line on stdout'
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run(['--impl', 'reference', '--steps', '1', '--warmup', '0'])
    for k in BASE + ['impl']:
        assert k in d, k
    assert d['impl'] == 'reference' and d['n_gpus'] == 1 and d['value'] > 0
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] and d['e2e']['value'] == d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


@pytest.mark.gpu
def test_our_arm_line():
    d = run(['--steps', '6', '--warmup', '3', '--clip-frames', '192', '--cpu-sample', '1'])
    for k in BASE + ['gpu_launches', 'roofline', 'clocks', 'kernels']:
        assert k in d, k
    assert d['metric'].startswith('frames/sec') and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 6 and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['dtype'] == 'int8' and d['data'] == 'synthetic' and 'workload' in d['config']
    assert d['value'] > 1000 and d['gpu_launches'] > 100
    e = d['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert e['value'] < d['value'], 'end to end includes the host<->device copies'
    r = d['roofline']
    assert r['bound'] in ('hbm', 'tensor') and r['unit'] in ('GB/s', 'TFLOP/s')
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and 0 < r['frac'] < 1
    c = d['cpu_baseline']
    assert c['value'] > 0 and c['cores'] >= 1 and c['kind'] == 'port' and c['sample']
    assert 'sm_mhz' in d['clocks'] and 'reasons' in d['clocks']
    # a step is one whole clip; the NCCL gather is timed apart; the end-to-end roofline fraction and the
    # kernel with the largest time share (K7 is allowed, bound "latency"; the roofs by algorithmic intensity) are reported
    assert d['config']['step'].startswith('192 frames') and d['gather_ms'] >= 0 and d['ms_per_batch'] > 0
    assert 0 < r['e2e_frac'] < 1 and r['dominant_by_time']['bound'] in ('hbm', 'tensor', 'latency')
    assert {'K7_tracker', 'K1_preprocess'} <= set(d['kernels']) and d['kernels']['K7_tracker']['bound'] == 'latency'
    assert any(k == 'mbconv_fused' for k in d['kernels']), 'the backbone runs as fused MBConv blocks'


@pytest.mark.gpu
def test_configs4_workload_line():
    """34 fixture-length clips through shard.track_videos (one rank here): every gathered table equals the
    1-rank table; the line carries the LPT bound the sharding is measured against."""
    d = run(['--workload', 'configs4', '--clip-frames', '256'])
    assert d['config']['frames'] == 55001 and len(d['config']['rank_loads_frames']) == 1
    assert d['parity']['videos_byte_identical_to_1_rank_run'] == d['parity']['videos'] == 34
    assert d['parity']['rows'] > 0 and d['value'] > 1000 and d['scaling'] == 'strong'
