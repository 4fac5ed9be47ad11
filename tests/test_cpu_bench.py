"""bench.py's reference arm (the CPU implementation of the path, which is what the driver's ratio is computed against)
runs without a GPU: one JSON line with the contract's keys.  The GPU arm is exercised on the B200 box only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2',
                          '--warmup', '1', '--cpu-seconds', '1', '--blocks', '1', '--playouts', '16'],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    line = json.loads(out.stdout.decode().strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'mcts_simulations_per_sec'
    assert line['unit'] == 'simulations/s' and line['value'] > 0 and line['higher_is_better'] is True
    assert line['n_gpus'] == 1 and line['steps'] == 2 and line['warmup'] == 1
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == line['value'] and 'playouts' in cb['sample']
    assert line['e2e'] == {'value': line['value'], 'unit': 'simulations/s', 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                          '--steps', '2', '--warmup', '1'], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.decode().strip() == ''
