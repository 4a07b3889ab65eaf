"""Build recipe for libb200phy.so (hand-written sm_100a CUDA behind a C ABI).

``python -m pyphysim_b200._build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a
GPU; the resulting .so lives in-tree (git-ignored) so it travels to the GPU box with the snapshot.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(PKG, 'build')
LIB = os.path.join(PKG, 'libb200phy.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--expt-relaxed-constexpr', '-Xcompiler', '-fPIC', '-Xcompiler', '-O2',
              '-Xptxas', '-v']


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build libb200phy.so')
    return exe


def have_nvcc():
    return bool(shutil.which('nvcc')) or os.path.exists('/usr/local/cuda/bin/nvcc')


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(PKG), 'include')):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), 'rb') as fh:
                h.update(f.encode())
                h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, extra_flags=(), lib_path=None, obj_dir=None):
    """Compile every .cu under csrc/ for sm_100a and link libb200phy.so.  Returns the path."""
    LIB = lib_path or globals()['LIB']
    OBJ = obj_dir or globals()['OBJ']
    os.makedirs(OBJ, exist_ok=True)
    stamp = LIB + '.digest'          # next to the library: it travels with it (the object directory does not)
    dig = _digest() + ' '.join(extra_flags)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj[:-2] + '.log', 'w') as fh:
            fh.write(' '.join(cmd) + '\n' + log)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (src, log))
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    with open(stamp, 'w') as fh:
        fh.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
