"""Build the CUDA library in-tree (opensbli_b200/libosbli_b200.so) for sm_100a with nvcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'osb_driver.cu')
DEPS = [SRC, os.path.join(HERE, 'csrc', 'osb_kernels.cuh'), os.path.join(HERE, 'csrc', 'osb_math.cuh'), os.path.join(HERE, 'csrc', 'osb_flux.cuh'),
        os.path.join(os.path.dirname(HERE), 'include', 'osbli_b200.h')]
LIB = os.path.join(HERE, 'libosbli_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-ldl']


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + [SRC, '-o', LIB]
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
