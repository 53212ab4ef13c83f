"""bench.py's reference arm (the restated CPU oracle on the host cores) prints exactly one JSON line with the keys the
driver reads; ranks other than 0 print nothing. No GPU involved."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2',
                           '--warmup', '1', '--ref-envs-per-core', '8'], capture_output=True, text=True, env=e,
                          timeout=300)


def test_reference_arm_prints_one_json_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'env-steps/sec' and d['unit'] == 'env-steps/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 2 and d['warmup'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] == (os.cpu_count() or 1)
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['gpu_launches'] == 0 and 'workload' in d['config']


def test_reference_arm_other_ranks_stay_silent():
    r = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''
