"""CPU tests of the dataset files (output / checkpoint / restart) in the reference's HDF5 layout (core/io_hdf5.py:99-127,
utilities/helperfunctions.py:80-127): same group, dataset names and attributes whether h5py is present (HDF5) or not (npz)."""
import numpy as np
import pytest

from opensbli_b200 import iodata


def _arrays(nd, np_, seed=0):
    rng = np.random.default_rng(seed)
    shape = tuple(n + 10 for n in reversed(np_))
    names = ['rho'] + ['rhou%d' % d for d in range(nd)] + ['rhoE']
    return {n: rng.standard_normal(shape) for n in names}


@pytest.mark.parametrize('np_', [[12], [9, 7], [6, 8, 7]])
def test_round_trip_npz(tmp_path, np_):
    nd = len(np_)
    arrays = _arrays(nd, np_)
    path = iodata.write_datasets(str(tmp_path / 'opensbli_output'), arrays, np_, force_npz=True)
    assert path.endswith('.npz')
    data, attrs = iodata.read_datasets(path)
    assert sorted(data) == sorted(arrays)
    for n, a in arrays.items():
        assert np.array_equal(data[n], a)                       # bit-exact: a checkpoint restarts the run exactly
        assert list(attrs[n]['d_m']) == [-5] * nd and list(attrs[n]['d_p']) == [5] * nd
        assert list(attrs[n]['size']) == np_
        inner = iodata.strip_halos(data[n], attrs[n])           # what the apps' plot.py scripts cut out
        assert inner.shape == tuple(reversed(np_))
        assert np.array_equal(inner, a[(slice(5, -5),) * nd])
    # lookup by base name, as the runner's --restart option does
    data2, _ = iodata.read_datasets(str(tmp_path / 'opensbli_output'))
    assert np.array_equal(data2['rho'], arrays['rho'])


def test_hdf5_layout_when_h5py_is_available(tmp_path):
    if not iodata.have_h5py():
        pytest.skip('h5py not installed in this image: the npz stand-in carries the same names and attributes')
    import h5py
    np_ = [9, 7]
    arrays = _arrays(2, np_)
    path = iodata.write_datasets(str(tmp_path / 'opensbli_output'), arrays, np_)
    with h5py.File(path, 'r') as f:                             # read exactly as apps/*/plot.py do
        g = f['opensbliblock00']
        d_m = g['rho_B0'].attrs['d_m']
        size = g['rho_B0'].shape
        rs, re_ = [abs(d) for d in d_m], [s - abs(d) for d, s in zip(d_m, size)]
        assert np.array_equal(g['rho_B0'][rs[0]:re_[0], rs[1]:re_[1]], arrays['rho'][5:-5, 5:-5])


def test_missing_file_is_an_error(tmp_path):
    with pytest.raises(IOError):
        iodata.read_datasets(str(tmp_path / 'nothing_here'))
