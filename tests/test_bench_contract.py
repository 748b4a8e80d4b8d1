"""CPU-only: the reference arm of bench.py (the oracle port timed on the host cores) prints exactly one JSON line
on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--ref-scale', '9',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'link structural features/sec' and d['unit'] == 'links/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['steps'] == 1 and d['warmup'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'links/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'].startswith('rmat24') and d['vs_baseline'] is None
    assert d['native_so_loaded'] == [], 'the reference arm must not map any library of this repository'


def test_reference_arm_non_zero_ranks_stay_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                          '--ref-scale', '9', '--steps', '1', '--warmup', '0'], capture_output=True, text=True,
                         timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
