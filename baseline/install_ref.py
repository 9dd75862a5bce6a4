"""Install the UNMODIFIED reference package into baseline/_ref/ (git-ignored; it travels to
the GPU box with the working tree) so that `bench.py --impl reference` and the
`cpu_baseline` leg time the reference's own code, not a port.

    python baseline/install_ref.py

Recipe (recorded in DESIGN.md section 4):
  1. /root/reference is read-only and its build writes into the source tree, so the tree is
     copied to a scratch directory;
  2. the copy gets py/rvspecfit/_version.py -- the file setuptools_scm would generate
     (pyproject.toml:50-51; setuptools_scm is not in this image), imported by
     rvspecfit/__init__.py:1;
  3. `python -m pip install --no-index --no-build-isolation --no-deps --find-links
     /opt/wheelhouse --target baseline/_ref <copy>` -- `--no-deps` because the wheelhouse has
     no numpy/scipy wheels to "resolve" (they are installed); the build compiles the cffi
     spline `_spliner` exactly as the reference's setup.py does.
Nothing of the reference enters the repository's history.  Absent third-party modules
(h5py, astropy, numdifftools, matplotlib) are stubbed at import time by baseline/ref_loader.py.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')


def install(force=False):
    marker = os.path.join(DST, 'rvspecfit', '_version.py')
    if os.path.exists(marker) and not force:
        return DST
    if not os.path.isdir(SRC):
        raise RuntimeError(f'{SRC} is not mounted: the reference can only be installed in the '
                           'build container')
    work = tempfile.mkdtemp(prefix='rvs_ref_src_')
    try:
        copy = os.path.join(work, 'reference')
        shutil.copytree(SRC, copy, ignore=shutil.ignore_patterns('.git'))
        with open(os.path.join(copy, 'py', 'rvspecfit', '_version.py'), 'w') as fp:
            fp.write("version = '0+local'\n__version__ = version\n")
        if os.path.exists(DST):
            shutil.rmtree(DST)
        subprocess.check_call([sys.executable, '-m', 'pip', 'install', '--no-index',
                               '--no-build-isolation', '--no-deps', '--find-links',
                               '/opt/wheelhouse', '--target', DST, copy],
                              stdout=subprocess.DEVNULL)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    assert os.path.exists(marker), 'reference install incomplete'
    return DST


if __name__ == '__main__':
    print(install(force='--force' in sys.argv))
