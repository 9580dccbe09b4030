"""The CPU arm of bench.py (`--impl reference`: the oracle timed on the host)
prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '0', '--sample-grid', '40', '--trend-grid', '56', '--grid', '1024'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
              'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config',
              'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['dtype'] == 'f64' and d['vs_baseline'] is None
    kind = 'reference' if os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'numpad', 'adarray.py')) else 'port'
    assert d['cpu_baseline']['kind'] == kind and d['cpu_baseline']['cores'] == 1
    assert d['loaded_engine_modules'] == []      # the CPU arm never imports the engine
    assert len(d['cpu_baseline']['trend']) == 2 and d['cpu_baseline']['trend_extrapolation']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    assert 'workload' in d['config'] and d['value'] > 0


def test_reference_arm_under_torchrun_prints_once():
    """launched like the driver does for N > 1: rank 0 alone runs and prints, the other ranks
    exit 0 without work; the line describes the same (strong-scaling) job"""
    import socket
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                          '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                          '--master-port', str(port), os.path.join(ROOT, 'bench.py'),
                          '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0',
                          '--sample-grid', '32', '--trend-grid', '0', '--grid', '1024'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['scaling'] == 'strong'
    assert 'row-sharded' in d['config']['parallelism']
