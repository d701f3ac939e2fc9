"""Builds libdockgpu.so (sm_100a only) in-tree with nvcc; also used by __graft_entry__.build()."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libdockgpu.so')
SOURCES = ['capi.cu', 'msm_g1.cu', 'msm_g2.cu', 'batch_g1.cu', 'batch_g2.cu', 'pairing.cu', 'ntt.cu', 'serialize.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-std=c++17', '-O3', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(HERE, '..', 'include', 'dockgpu.h'))
    objdir = os.path.join(HERE, '..', 'build', 'obj')
    os.makedirs(objdir, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace('.cu', '.o'))
        if force or _stale(obj, [src] + headers):
            cmd = ['nvcc'] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose or res.returncode:
                    sys.stderr.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
                if res.returncode:
                    raise RuntimeError('nvcc failed for ' + cmd[-3])
    objs = [os.path.join(objdir, s.replace('.cu', '.o')) for s in srcs]
    if force or jobs or _stale(OUT, objs):
        cmd = ['nvcc', '-shared', '-o', OUT] + objs + ['-lcudart']
        subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
