"""Build libttk.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python -m upliftingtabletennis_b200.build [--force]

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels to the GPU
box with the repository snapshot.  cudart is linked statically, the driver API
(cuTensorMapEncodeTiled) is resolved at run time, so the library loads on a GPU-less host.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libttk.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden'] + os.environ.get('TTK_NVCC_FLAGS', '').split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.h') or f.endswith('.cuh')]
    hs.append(os.path.join(HERE, '..', 'include', 'ttk.h'))
    return max(os.path.getmtime(h) for h in hs)


def compile_one(src, force):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    path = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), headers_mtime()):
        return obj, False
    cmd = [NVCC] + FLAGS + ['-c', path, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(lambda s: compile_one(s, force), sources()))
    objs = [o for o, _ in results]
    if force or any(c for _, c in results) or not os.path.exists(LIB):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
        if verbose:
            print('built', LIB)
    elif verbose:
        print('up to date', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)
