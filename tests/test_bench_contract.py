"""bench.py contract, CPU side: the reference arm prints ONE JSON line with the keys the driver reads, ranks other than 0
print nothing, and the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None):
    e = dict(os.environ, T2B200_BENCH_CPU_BUDGET='0.5')
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_line():
    r = run(['--impl', 'reference', '--steps', '1', '--warmup', '0'])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['metric'] == 'ldpc_codewords_per_s' and d['unit'] == 'codewords/s'
    assert d['value'] > 0 and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    assert 'workload' in d['config']
    # same-config arms: both print the config of bench.workload_config (the driver compares them), and the reference arm is
    # the whole stock chain over C32 frames, not the LDPC stage alone
    sys.path.insert(0, ROOT)
    import bench
    assert d['config'] == bench.workload_config(1)
    assert set(d['cpu_baseline']['ms_per_frame_per_core']) == {'fft', 'equalize', 'ti_demap', 'ldpc_bch'}
    assert d['cpu_baseline']['e2e_1core']['bbframes'] == 192


def test_reference_arm_other_ranks_are_silent():
    r = run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--gpus', '2'], env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = run(['--steps', '1'])
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
