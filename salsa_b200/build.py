"""Builds libsalsa_b200.so in-tree with nvcc for sm_100a (no torch headers involved)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libsalsa_b200.so')
SOURCES = ['salsa_abi.cu', 'crnn_abi.cu', 'crnn_model.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-I' + os.path.join(ROOT, 'include'), '-I' + CSRC]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', f) for f in os.listdir(os.path.join(ROOT, 'include'))]
    if not force and os.path.isfile(OUT) and os.path.getmtime(OUT) >= _newest(deps):
        return OUT
    objs, procs = [], []
    for src in srcs:                      # one nvcc per translation unit, side by side
        obj = os.path.join(HERE, '_build', os.path.basename(src) + '.o')
        os.makedirs(os.path.dirname(obj), exist_ok=True)
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd)))
        objs.append(obj)
    for cmd, proc in procs:
        if proc.wait() != 0:
            raise subprocess.CalledProcessError(proc.returncode, cmd)
    # the link step gets the arch flags too: without them nvcc adds an (empty) default-architecture device-link stub
    subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', OUT] + objs + ['-lcudart'], check=True)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
