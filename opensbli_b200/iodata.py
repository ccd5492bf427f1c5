"""Dataset files of the runner: output, checkpoints, restart input -- in the reference's HDF5 layout.

The reference writes its fields with `ops_fetch_block_hdf5_file` / `ops_fetch_dat_hdf5_file` (core/io_hdf5.py:99-127) and
reads initial or restart data with `ops_decl_dat_hdf5` (code_generation/opsc.py:702-705).  Layout (what the apps' plot.py
scripts and utilities/helperfunctions.py:80-127 `output_hdf5` rely on):

    /<block name>                       group, attrs: dims, ops_type = "ops_block", index
    /<block name>/<dataset>_B0          padded array (numpy order k, j, i), attrs: d_m = [-5, ..], d_p = [5, ..], size = interior
                                        points per direction (x first), base, dim, type = "double", block, block_index, ops_type

When `h5py` can be imported the files are real HDF5 (`opensbli_output.h5`, `opensbli_output_%06d.h5` for the in-loop dumps of
`iohdf5(save_every=N)`).  This image has no h5py: the same group / dataset names and attributes then go into an `.npz` archive
(`<block>/<dataset>` keys, attributes under `<block>/<dataset>@<attr>`), which `read_datasets` reads back the same way.
"""
import os

import numpy as np

BLOCK = 'opensbliblock00'


def have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except Exception:
        return False


def dataset_attrs(nd, np_, halo, block=BLOCK):
    """attributes OPS attaches to a dataset (helperfunctions.py:80-101)"""
    return {'d_m': np.array([-halo] * nd, dtype=np.int32), 'd_p': np.array([halo] * nd, dtype=np.int32),
            'size': np.array([int(n) for n in np_[:nd]], dtype=np.int32), 'base': np.zeros(nd, dtype=np.int32),
            'dim': np.array([1], dtype=np.int32), 'block_index': np.array([0], dtype=np.int32),
            'ops_type': 'ops_dat', 'type': 'double', 'block': block}


def write_datasets(path_base, arrays, np_, halo=5, block=BLOCK, force_npz=False):
    """arrays: {dataset name without the _B0 suffix: padded ndarray}.  Returns the path written (.h5 or .npz)."""
    nd = len(np_)
    if have_h5py() and not force_npz:
        import h5py
        path = path_base + '.h5'
        with h5py.File(path, 'w') as hf:
            g = hf.create_group(block)
            g.attrs.create('dims', [nd], dtype='int32')
            g.attrs.create('ops_type', u'ops_block', dtype='S9')
            g.attrs.create('index', [0], dtype='int32')
            for name, a in arrays.items():
                d = g.create_dataset(name + '_B0', data=np.ascontiguousarray(a, dtype=np.float64))
                for k, v in dataset_attrs(nd, np_, halo, block).items():
                    if isinstance(v, str):
                        d.attrs.create(k, v, dtype='S%d' % max(len(v) + 1, 10))
                    else:
                        d.attrs.create(k, v, dtype='int32')
        return path
    path = path_base + '.npz'
    out = {block + '@dims': np.array([nd], dtype=np.int32), block + '@ops_type': np.array('ops_block'), block + '@index': np.array([0], dtype=np.int32)}
    for name, a in arrays.items():
        key = '%s/%s_B0' % (block, name)
        out[key] = np.ascontiguousarray(a, dtype=np.float64)
        for k, v in dataset_attrs(nd, np_, halo, block).items():
            out['%s@%s' % (key, k)] = np.array(v)
    tmp = path + '.tmp.npz'
    np.savez(tmp, **out)
    os.replace(tmp, path)                      # a checkpoint is either complete or absent
    return path


def read_datasets(path, block=BLOCK):
    """-> ({dataset name without _B0: padded ndarray}, {dataset name: attrs}) from a file written by write_datasets, by the
    reference (HDF5) or by `output_hdf5`."""
    if not os.path.exists(path):
        for ext in ('.h5', '.npz'):
            if os.path.exists(path + ext):
                path = path + ext
                break
            base = os.path.splitext(path)[0]
            if os.path.exists(base + ext):
                path = base + ext
                break
        else:
            raise IOError('dataset file %s not found (.h5 / .npz)' % path)
    data, attrs = {}, {}
    if path.endswith('.npz'):
        z = np.load(path)
        pre = block + '/'
        for k in z.files:
            if k.startswith(pre) and '@' not in k:
                name = k[len(pre):]
                name = name[:-3] if name.endswith('_B0') else name
                data[name] = z[k]
                attrs[name] = {a.split('@', 1)[1]: z[a] for a in z.files if a.startswith(k + '@')}
        return data, attrs
    if not have_h5py():
        raise IOError('%s is an HDF5 file and h5py is not available in this environment' % path)
    import h5py
    with h5py.File(path, 'r') as hf:
        g = hf[block]
        for k in g.keys():
            name = k[:-3] if k.endswith('_B0') else k
            data[name] = np.array(g[k])
            attrs[name] = {a: np.array(v) for a, v in g[k].attrs.items()}
    return data, attrs


def strip_halos(array, attrs):
    """interior of a dataset, as the apps' plot.py scripts cut it (read_start = |d_m|, read_end = shape - |d_m|)"""
    d_m = [abs(int(v)) for v in np.atleast_1d(attrs['d_m'])]
    return array[tuple(slice(h, array.shape[n] - h) for n, h in enumerate(d_m))]
