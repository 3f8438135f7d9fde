"""Build libtampc_mppi.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m tampc_b200.build [--force] [--verbose]

One object per translation unit, compiled in parallel; objects are reused when newer than every source/header.
The shared library is what travels to the GPU box (built .so files are git-ignored, not gpurun-ignored)."""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')
BUILD = os.path.join(CSRC, 'build')
LIB = os.path.join(CSRC, 'libtampc_mppi.so')

SOURCES = ['mppi_api.cu', 'gp_fit.cu', 'variant_tc.cu', 'variant_generic.cu', 'variant_rex_push.cu', 'variant_rex_peg.cu', 'variant_mlp.cu', 'variant_mlp64.cu', 'variant_mlp128.cu', 'variant_gp.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; the CUDA extension cannot be built')
    return nvcc


def _deps_mtime():
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    paths += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return max(os.path.getmtime(p) for p in paths)


def _compile(src, force, verbose):
    obj = os.path.join(BUILD, src.replace('.cu', '.o'))
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), _deps_mtime()):
        return obj, 'cached'
    cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-I', INCLUDE, '-c', srcp, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed for {}:\n{}\n{}'.format(src, r.stdout, r.stderr))
    return obj, (r.stderr if verbose else 'built')


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for (o, log) in results:
            print(os.path.basename(o), log)
    need_link = force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if need_link:
        cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n{}\n{}'.format(r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    lib = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(lib)
