"""Build the CUDA library in-tree (opensbli_b200/libosbli_b200.so) for sm_100a with nvcc.

Translation units: csrc/osb_driver.cu (context, C ABI, every kernel but the flux sweeps) and csrc/osb_flux_tu.cu compiled once
per (ndim, reconstruction) pair -- 13 objects built in parallel, rebuilt only when one of their sources changed."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# experiments: OSB_BUILD_TAG=<tag> OSB_BUILD_FLAGS="-DOSB_F3_XBLOCKS=4 ..." builds libosbli_b200_<tag>.so beside the product library;
# OSB_B200_LIB=<path> makes the runtime load it
TAG = os.environ.get('OSB_BUILD_TAG', '')
OBJ = os.path.join(HERE, 'build' + ('_' + TAG if TAG else ''))
LIB = os.path.join(HERE, 'libosbli_b200%s.so' % ('_' + TAG if TAG else ''))
HEADERS = [os.path.join(CSRC, f) for f in ('osb_kernels.cuh', 'osb_math.cuh', 'osb_flux.cuh', 'osb_types.cuh', 'osb_flux_api.h')] + \
          [os.path.join(os.path.dirname(HERE), 'include', 'osbli_b200.h')]
FLUX_HEADERS = [os.path.join(CSRC, f) for f in ('osb_math.cuh', 'osb_flux.cuh', 'osb_types.cuh', 'osb_flux3.cuh', 'osb_flux_api.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC'] + os.environ.get('OSB_BUILD_FLAGS', '').split()


def units():
    """(object file, source, extra flags, dependencies)"""
    u = [(os.path.join(OBJ, 'osb_driver.o'), os.path.join(CSRC, 'osb_driver.cu'), [], HEADERS)]
    for nd in (1, 2, 3):
        for recon in range(4):
            u.append((os.path.join(OBJ, 'osb_flux_%d_%d.o' % (nd, recon)), os.path.join(CSRC, 'osb_flux_tu.cu'),
                      ['-DOSB_FLUX_ND=%d' % nd, '-DOSB_FLUX_RECON=%d' % recon], FLUX_HEADERS))
    return u


def _stale(target, deps):
    return not os.path.exists(target) or any(not os.path.exists(d) or os.path.getmtime(target) < os.path.getmtime(d) for d in deps)


def up_to_date():
    return not _stale(LIB, [o for o, _, _, _ in units()]) and all(not _stale(o, [src] + deps) for o, src, _, deps in units())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    nvcc = os.environ.get('NVCC', 'nvcc')
    os.makedirs(OBJ, exist_ok=True)
    todo = [(o, src, fl) for o, src, fl, deps in units() if force or _stale(o, [src] + deps)]

    def compile_one(job):
        o, src, fl = job
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + fl + ['-c', src, '-o', o]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return o, r.returncode, r.stdout
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        for o, rc, out in ex.map(compile_one, todo):
            if verbose or rc:
                print(out)
            if rc:
                raise subprocess.CalledProcessError(rc, 'nvcc ... -o ' + o)
    subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC'] +
                          [o for o, _, _, _ in units()] + ['-ldl', '-o', LIB])
    return LIB


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
