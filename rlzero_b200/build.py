"""Build the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m rlzero_b200.build [--force] [--verbose]

The shared object lands next to this file (``librlzero_b200.so``): it is git-ignored but
travels with the repository snapshot to the GPU box.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'librlzero_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-diag-suppress', '68']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _deps():
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh'))
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'rlzero_b200.h'))
    deps.append(os.path.abspath(__file__))
    return deps


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared object (object files cached in build/)."""
    if not force and not is_stale():
        return LIB
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    headers = [d for d in _deps() if not d.endswith('.cu')]
    hdr_time = max(os.path.getmtime(h) for h in headers)
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hdr_time):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (src, out))
        if verbose:
            print(out)
    cmd = [nvcc, '-shared', '--cudart', 'shared', '-o', LIB + '.tmp'] + objs
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if out.returncode != 0:
        raise RuntimeError('link failed:\n' + out.stdout.decode())
    os.replace(LIB + '.tmp', LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
