"""In-tree build of libtxl_b200.so: nvcc -> sm_100a objects -> one shared library with the C ABI of include/txl_b200.h.

Used by `__graft_entry__.build()`.  Objects are cached under `build/` (git-ignored) and rebuilt when a source or header
is newer; the .so lands next to this file so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(ROOT, 'build', 'txl_b200')
LIB = os.path.join(HERE, 'libtxl_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--use_fast_math',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '-diag-suppress', '550']
# --use_fast_math would change expf/logf/division in the fp32 parity kernels, so it is NOT applied to simt_*.cu
PRECISE = {'simt_ops.cu', 'simt_gemm.cu', 'simt_relattn.cu', 'sample.cu'}


def _nvcc():
    return shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(ROOT, 'include', '*.h'))
    sources = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        if force or _newer([src] + headers, obj):
            flags = [f for f in NVCC_FLAGS if not (f == '--use_fast_math' and os.path.basename(src) in PRECISE)]
            jobs.append((src, obj, [_nvcc()] + flags + ['-c', src, '-o', obj]))

    def run(job):
        src, obj, cmd = job
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{p.stdout}\n{p.stderr}')
        with open(obj + '.log', 'w') as f:
            f.write(p.stdout + p.stderr)
        if verbose:
            print(p.stderr, file=sys.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + '.o') for s in sources]
    if force or jobs or _newer(objs, LIB):
        cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f'link failed:\n{p.stdout}\n{p.stderr}')
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose=True))
