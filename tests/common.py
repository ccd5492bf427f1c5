"""Shared helpers of the parity tests."""
import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')

# Per-step tolerance on conserved fields, relative to the field's L-inf norm (momentum components share the
# norm of the momentum vector).  north_star: 1e-12.  WENO-Z is the exception: with eps = 1e-14 the
# reference's own result moves by ~5e-11 per step between two builds of its generated C (-ffp-contract off
# vs fast, see DESIGN.md "reference noise floor"), so parity is asserted at 1e-9 for it.
TOL = {'default': 1e-12, 'weno_Z': 1e-9}


def tol_for(plan, nsteps=1):
    base = TOL['weno_Z'] if (plan['conv'] == 'weno' and plan.get('weno_formulation') == 'Z') else TOL['default']
    return base * max(1.0, nsteps / 10.0)


def fixtures():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz')))


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    plan = json.loads(str(z['plan']))
    states = {int(k[1:]): z[k] for k in z.files if k.startswith('q') and k[1:].isdigit()}
    # general-path fixtures carry metric fields, a Dirichlet table and the padded initial state
    fields = {k[6:]: z[k] for k in z.files if k.startswith('field_')}
    if fields:
        plan['fields'] = fields
    for k in z.files:
        if k.startswith('bc_table_'):
            d, s = int(k.split('_')[2]), int(k.split('_')[3])
            plan['bc'][d][s]['table'] = z[k]
    if 'q0_padded' in z.files:
        plan['q0_padded'] = z['q0_padded']
    stats = {k[5:]: z[k] for k in z.files if k.startswith('stat_')}
    if stats:
        plan['stats_golden'] = stats
    return plan, states


def initial_padded(plan, states):
    """padded initial state: the stored one if the halos matter (general path), else zero-padded interior."""
    if 'q0_padded' in plan:
        return [np.array(a, dtype=np.float64, order='C', copy=True) for a in plan['q0_padded']]
    return pad(plan, states[0])


def pad(plan, q_inner, halo=5):
    """interior array(s) (nv, ...) -> list of zero-padded arrays in the reference layout."""
    nd = plan['ndim']
    return [np.pad(a, [(halo, halo)] * nd) for a in q_inner]


def inner(plan, q, halo=5):
    s = (slice(halo, -halo),) * plan['ndim']
    return np.stack([a[s] for a in q])


def field_errors(plan, q, ref):
    """max |q-ref| / norm per conserved field; q, ref: (nv, ...) interior arrays."""
    nd = plan['ndim']
    norms = [np.abs(ref[0]).max()] + [max(np.abs(ref[1:1 + nd]).max(), 1e-300)] * nd + [np.abs(ref[nd + 1]).max()]
    return [float(np.abs(q[m] - ref[m]).max() / norms[m]) for m in range(nd + 2)]
