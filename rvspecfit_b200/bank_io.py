"""On-disk template products of the reference -> device-resident banks
(SURVEY.md section 8 row f2).

Reads what the reference's preparation scripts write under
config['template_lib'] (paths under /root/reference/py/rvspecfit/):
    interp_<setup>.h5 + interpdat_<setup>.npy   (make_nd.py:14-15,161-177; keys
        consumed as in spec_inter.py:330-370)
    ccf_<setup>.h5 + ccfdat_<setup>.npz + ccfmod_<setup>.npy
        (make_ccf.py:27-36,483-493; fitter_ccf.py:39-56)
The HDF5 dictionaries use the reference serializer's typed-attribute scheme
(serializer.py:112-157).  h5py is needed only here and only at load time; banks
can always be registered from memory instead (spec_inter.register_bank,
fitter_ccf.register_ccf_bank).
"""
import os
import pickle

import numpy as np

H5_VERSION = 1      # serializer.CURRENT_VERSION the products are written with


def _h5py():
    try:
        import h5py
    except ImportError as exc:
        raise RuntimeError('reading the reference\'s interp_*.h5 / ccf_*.h5 products needs '
                           'h5py, which is not installed; register the banks from memory '
                           'instead (spec_inter.register_bank)') from exc
    return h5py


def _decode(group, h5py):
    out = {}
    for key, item in group.items():
        if isinstance(item, h5py.Group):
            sub = _decode(item, h5py)
            kind = item.attrs.get('type')
            if kind in ('flattened_tuple', 'flattened_list'):
                seq = [sub['__item_%d' % i] for i in range(len(sub))]
                sub = tuple(seq) if kind == 'flattened_tuple' else seq
            out[key] = sub
            continue
        kind = item.attrs['type']
        if kind in ('list', 'tuple', 'ndarray'):
            val = item[:]
            if item.dtype.kind == 'O':
                val = val.astype(str)
            out[key] = {'list': list, 'tuple': tuple, 'ndarray': lambda v: v}[kind](val)
        elif kind == 'str':
            out[key] = item[()].decode('utf-8')
        elif kind in ('scalar', 'empty_array'):
            out[key] = item[()]
        elif kind == 'pickle':
            out[key] = pickle.loads(item[()])
        elif kind == 'None':
            out[key] = None
        else:
            raise ValueError(f'unsupported entry type {kind!r} for key {key!r}')
    return out


def load_h5_dict(filename):
    """Dictionary stored by the reference's serializer.save_dict_to_hdf5."""
    if not os.path.exists(filename):
        raise RuntimeError(f'Filename {filename} does not exist')
    h5py = _h5py()
    with h5py.File(filename, 'r') as fh:
        version = fh.attrs.get('version', None)
        if version != H5_VERSION:
            raise ValueError(f'Incompatible version: {version}')
        return _decode(fh['/'], h5py)


def bank_from_dict(name, fd, dats):
    """TemplateBank from the loaded interp_<setup>.h5 dictionary and the
    interpdat rows (spec_inter.py:330-370)."""
    from . import spec_inter
    kind = fd.get('interpolation_type')
    if kind is None:
        kind = 'triangulation' if 'triang' in fd else ('regulargrid' if 'regular' in fd else None)
    if kind not in ('triangulation', 'regulargrid'):
        raise RuntimeError('Unrecognized interpolation file')
    if fd.get('mapper_class_name', 'LogParamMapper') != 'LogParamMapper':
        raise RuntimeError(f"parameter mapper {fd['mapper_class_name']} is not supported")
    args = fd.get('mapper_args') or ([0],)
    log_ids = tuple(int(_) for _ in np.atleast_1d(args[0]))
    common = dict(log_ids=log_ids, log_step=bool(fd['log_step']),
                  log_spec=bool(fd.get('log_spec', True)))
    parnames = [str(_) for _ in fd['parnames']]
    if kind == 'regulargrid':
        return spec_inter.TemplateBank(name, fd['lam'], dats, parnames, kind=kind,
                                       uvecs=fd['uvecs'], idgrid=fd['idgrid'], vecs=fd['vec'],
                                       **common)
    return spec_inter.TemplateBank(name, fd['lam'], dats, parnames, kind=kind,
                                   triang=fd['triang'], extraflags=fd['extraflags'], **common)


def load_bank(setup, config):
    prefix = config['template_lib'] + '/'
    fd = load_h5_dict(prefix + 'interp_%s.h5' % setup)
    dats = np.load(prefix + 'interpdat_%s.npy' % setup, mmap_mode='r')
    return bank_from_dict(setup, fd, dats)


def load_ccf_bank(setup, config):
    from . import fitter_ccf
    cont = config.get('ccf_continuum_normalize')
    pref = '' if (cont is None or cont) else 'nocont_'
    prefix = config['template_lib'] + '/'
    info = load_h5_dict(prefix + 'ccf_' + pref + '%s.h5' % setup)
    dat = np.load(prefix + 'ccfdat_' + pref + '%s.npz' % setup, mmap_mode='r')
    models = np.load(prefix + 'ccfmod_' + pref + '%s.npy' % setup, mmap_mode='r')
    return fitter_ccf.CcfBank(setup, dat['fft'], dat['fft2'], models, info['params'],
                              info['vsinis'], info['parnames'], info['ccfconf'])
