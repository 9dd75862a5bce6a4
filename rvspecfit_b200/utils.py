"""Configuration handling with the reference's semantics (reference
py/rvspecfit/utils.py:9-110): same keys, same defaults, missing keys filled
from the defaults, result frozen so that it can key caches."""
import logging
import os

import yaml


class frozendict(dict):
    """Hashable, immutable dict (role of reference frozendict.py:24)."""

    def _blocked(self, *a, **k):
        raise TypeError('frozendict is immutable')

    __setitem__ = __delitem__ = clear = pop = popitem = setdefault = update = _blocked

    def __hash__(self):
        return hash(tuple(sorted(self.items(), key=lambda kv: str(kv[0]))))


def get_default_config():
    """Defaults of reference utils.py:9-28."""
    return {'min_vel': -1000, 'max_vel': 1000, 'vel_step0': 5, 'max_vsini': 500,
            'min_vsini': 1e-2, 'min_vel_step': 0.2, 'second_minimizer': True,
            'template_lib': 'templ_data/'}


def freezeDict(d):
    if isinstance(d, dict):
        return frozendict({k: freezeDict(v) for k, v in d.items()})
    if isinstance(d, list):
        return tuple(d)
    return d


def read_config(fname=None, override_options=None):
    """Read config.yaml (reference utils.py:31-82)."""
    given = fname is not None
    if fname is None:
        fname = 'config.yaml'
    if os.path.exists(fname):
        with open(fname, 'r') as fp:
            D = yaml.safe_load(fp)
        if D is None:
            D = {}
            logging.warning('Configuration file is empty. Using default settings')
    else:
        if given:
            raise RuntimeError(f"Configuration file '{fname}' not found.")
        logging.warning(f"Configuration file '{fname}' not found. Using default settings")
        D = {}
    for k, v in get_default_config().items():
        if k not in D:
            D[k] = v
    D['config_file_path'] = os.path.abspath(fname)
    if override_options is not None:
        for k, v in override_options.items():
            if k in D and v != D[k]:
                logging.warning(f'Provided option {k} overrides the value in '
                                'the configuration file')
            D[k] = v
    return freezeDict(D)
