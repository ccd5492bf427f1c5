"""Host-side execution of the run-time compiled kernels (test infrastructure): the CUDA C that opensbli_b200.run prints for
user / boundary kernels is compiled as C++ with g++ (one thread per point becomes a loop nest) and run on numpy arrays.  The CPU
suite checks the PRINTED kernels of the generic paths (filters, statistics, boundary classes without a hand-written kernel)
against the reference's golden states this way, without a GPU; the GPU tests then only confirm the same sources under NVRTC."""
import ctypes
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_PRE = r'''
#include <cmath>
#include <algorithm>
#define __global__
#define __device__
struct D3 { int x, y, z; };
static D3 blockIdx, threadIdx, blockDim;
'''


class HostKernels(object):
    """kernels: resolved user kernels of a plan (dicts with source, entry, fields, range, when, writes)."""

    def __init__(self, kernels, shape, halo=5, workdir=None):
        self.kernels = kernels
        self.shape = tuple(shape)
        self.nd = len(self.shape)
        self.h = halo
        self.fields = {}
        src = [_PRE]
        for n, k in enumerate(kernels):
            body = k['source'].replace('extern "C" ', '')
            src.append('namespace hk%d {\n%s\n}' % (n, body))
            src.append('extern "C" void hostrun_%d(long long off, int n0, int n1, int n2, int lo0, int lo1, int lo2, long long s1, long long s2, double **p, int nf, long long iter) {\n'
                       '  hk%d::UserFields f; for (int i = 0; i < nf; i++) f.p[i] = p[i]; f.iter = iter;\n  blockDim = {1, 1, 1}; threadIdx = {0, 0, 0};\n'
                       '  for (int z = 0; z < n2; z++) for (int y = 0; y < n1; y++) for (int x = 0; x < n0; x++) { blockIdx = {x, y, z};\n'
                       '    hk%d::%s(off, n0, n1, n2, lo0, lo1, lo2, s1, s2, f); }\n}' % (n, n, n, k['entry']))
        text = '\n'.join(src)
        d = workdir or os.path.join(tempfile.gettempdir(), 'osb_hostsim')
        os.makedirs(d, exist_ok=True)
        tag = hashlib.sha1(text.encode()).hexdigest()[:16]
        so = os.path.join(d, 'hk_%s.so' % tag)
        if not os.path.exists(so):
            cpp = os.path.join(d, 'hk_%s.cpp' % tag)
            open(cpp, 'w').write(text)
            subprocess.check_call(['g++', '-O1', '-ffp-contract=off', '-w', '-shared', '-fPIC', '-o', so, cpp])
        self.lib = ctypes.CDLL(so)

    def field(self, name):
        if name not in self.fields:
            self.fields[name] = np.zeros(self.shape)          # datasets start zeroed (OPS semantics)
        return self.fields[name]

    def run(self, when, iteration=0):
        h, nd = self.h, self.nd
        pd = list(reversed(self.shape))                         # padded extents, x first
        s1 = pd[0] if nd > 1 else 0
        s2 = pd[0] * pd[1] if nd > 2 else 0
        off = h * (1 + (s1 if nd > 1 else 0) + (s2 if nd > 2 else 0))
        for n, k in enumerate(self.kernels):
            if k['when'] != when and not (when.startswith('stage_') and k['when'] == 'stage'):      # 'stage': every RK stage
                continue
            arrs = [self.field(f) for f in k['fields']]
            assert all(a.flags['C_CONTIGUOUS'] and a.dtype == np.float64 for a in arrs)
            ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
            r = list(k['range']) + [0, 1] * (3 - nd)
            getattr(self.lib, 'hostrun_%d' % n)(ctypes.c_longlong(off), r[1] - r[0], r[3] - r[2], r[5] - r[4], r[0], r[2], r[4],
                                                ctypes.c_longlong(s1), ctypes.c_longlong(s2), ptrs, len(arrs), ctypes.c_longlong(iteration))
