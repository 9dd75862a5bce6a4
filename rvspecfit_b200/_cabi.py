"""ctypes binding of librvs_b200.so (include/rvs_b200.h).

The product has no CPU path: if the shared library is missing, or no CUDA
device is present when a compute entry point is called, this module raises.
PyTorch is used only to own device memory and streams; every compute call goes
through the C ABI with raw device pointers.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RVS_B200_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get('RVS_B200_LIB') or os.path.join(_HERE, 'librvs_b200.so')

c_dp = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double

ST_OK, ST_TEMPLATE_BAD, ST_NOT_PD, ST_RANGE, ST_TAPS, ST_LIMIT = 0, 1, 2, 4, 8, 16
MAX_NPOLY = 16
MAX_FUSED_TAPS = 128
MAX_ARMS = 4


class Knots(ctypes.Structure):
    """struct rvs_knots"""
    _fields_ = [('d_lam_t', c_dp), ('d_h', c_dp), ('d_hinv', c_dp), ('d_cp', c_dp),
                ('d_winv', c_dp), ('npix_t', ctypes.c_int32), ('log_step', ctypes.c_int32),
                ('x0', c_dbl), ('xlast', c_dbl), ('q0', c_dbl), ('qstep_inv', c_dbl),
                ('lnstep', c_dbl), ('ratio', c_dbl), ('ratio_dev', c_dbl)]


class Obs(ctypes.Structure):
    """struct rvs_obs"""
    _fields_ = [('d_lam', c_dp), ('d_loglam', c_dp), ('d_P', c_dp), ('d_goff', c_dp),
                ('d_dn', c_dp), ('d_einv', c_dp), ('d_sumlog2', c_dp), ('d_off', c_dp),
                ('npoly', ctypes.c_int32), ('npp', ctypes.c_int32), ('nobj', ctypes.c_int32),
                ('shared_grid', ctypes.c_int32), ('d_resol', c_dp), ('d_resol_offs', c_dp),
                ('nresol', ctypes.c_int32), ('resol_halfwidth', ctypes.c_int32)]


class GridMap(ctypes.Structure):
    """struct rvs_gridmap"""
    _fields_ = [('d_uvec', c_dp), ('d_idgrid', c_dp), ('ndim', ctypes.c_int32),
                ('len', ctypes.c_int32 * 5), ('uoff', ctypes.c_int32 * 5),
                ('nnode', ctypes.c_int32), ('d_vnorm', c_dp), ('ptp', c_dbl * 5)]


class GridBox(ctypes.Structure):
    """struct rvs_gridbox"""
    _fields_ = [('tmap', ctypes.c_ubyte * 128), ('len', ctypes.c_int32 * 4),
                ('cols', ctypes.c_int32), ('rows', ctypes.c_int32)]


class CcfArm(ctypes.Structure):
    """struct rvs_ccf_arm"""
    _fields_ = [('d_fft', c_dp), ('d_fft2', c_dp), ('d_lo', c_dp), ('d_hi', c_dp),
                ('d_dxn', c_dp), ('d_dx', c_dp), ('npoints', ctypes.c_int32),
                ('ntempl', ctypes.c_int32), ('continuum', ctypes.c_int32),
                ('nvel', ctypes.c_int32)]


class FusedArm(ctypes.Structure):
    """struct rvs_fused_arm"""
    _fields_ = [('d_grid', c_dp), ('grid_f64', ctypes.c_int32), ('log_spec', ctypes.c_int32),
                ('ld', c_i64), ('knots', ctypes.POINTER(Knots)), ('obs', ctypes.POINTER(Obs)),
                ('d_oix', c_dp), ('d_tn', c_dp), ('tn_stride', c_i64), ('d_work', c_dp),
                ('d_chisq', c_dp), ('d_status', c_dp), ('box', ctypes.POINTER(GridBox))]


class FitLayout(ctypes.Structure):
    """struct rvs_fit_layout"""
    _fields_ = [('nfit', ctypes.c_int32), ('nspec', ctypes.c_int32), ('fit_vsini', ctypes.c_int32),
                ('has_vsini', ctypes.c_int32), ('fixmask', ctypes.c_int32),
                ('logmask', ctypes.c_int32), ('priormask', ctypes.c_int32),
                ('narm', ctypes.c_int32), ('nobj', ctypes.c_int32), ('pad_', ctypes.c_int32),
                ('min_vel', c_dbl), ('max_vel', c_dbl), ('max_vsini', c_dbl),
                ('h_p0', c_dp), ('h_q0', c_dp), ('h_vsini0', c_dp), ('h_prior_mu', c_dp),
                ('h_prior_sig', c_dp), ('h_oix', c_dp), ('h_badchi', c_dp), ('h_cover', c_dp)]


class Drive(ctypes.Structure):
    """struct rvs_drive"""
    _fields_ = [('stream', c_dp), ('h_in', c_dp), ('h_oix', c_dp), ('h_chi', c_dp),
                ('h_flags', c_dp), ('f_prior', c_dp), ('f_pen', c_dp), ('f_wall', c_dp),
                ('f_out', c_dp), ('f_redo', c_dp), ('cap', c_i64),
                ('shared_locate', ctypes.c_int32), ('ngraph', ctypes.c_int32),
                ('g_kp', c_dp), ('g_vmax', c_dp), ('g_exec', c_dp), ('g_nk', c_dp),
                ('objmap', c_dp), ('nprob', c_i64), ('speculate_below', ctypes.c_int32),
                ('state', ctypes.c_int32), ('stop_stopped', c_i64), ('fused_vmax', c_dbl),
                ('K', c_i64), ('Kp', c_i64), ('vmax', c_dbl),
                ('rounds', c_i64), ('items', c_i64), ('graph_launches', c_i64),
                ('graph_kernels', c_i64), ('h2d_bytes', c_i64), ('d2h_bytes', c_i64),
                ('epoch_event', c_dp), ('t_rec', c_dp), ('t_cap', c_i64), ('t_n', c_i64),
                ('timed', ctypes.c_int32), ('pad_', ctypes.c_int32)]


DRIVE_DONE, DRIVE_PEEL, DRIVE_LAUNCH, DRIVE_REDO, DRIVE_PYEVAL = range(5)
DRIVE_IDLE, DRIVE_PACKED, DRIVE_LAUNCHED, DRIVE_COLLECTED = range(4)

# name -> (restype, argtypes); every symbol include/rvs_b200.h declares
SIGNATURES = {
    'rvs_last_error': (ctypes.c_char_p, []),
    'rvs_version': (c_int, []),
    'rvs_launch_count': (c_i64, []),
    'rvs_profile_enable': (None, [c_int]),
    'rvs_profile_active': (c_int, []),
    'rvs_profile_read': (c_int, [c_dp, c_dp, c_int]),
    'rvs_profile_timeline': (c_int, [c_dp, c_int]),
    'rvs_spline_construct': (None, [c_dp, c_dp, c_int, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'rvs_spline_eval': (c_int, [c_dp, c_int, c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_int,
                                c_dp]),
    'rvs_knot_tables': (None, [c_dp, c_int, c_dp, c_dp, c_dp, c_dp]),
    'rvs_knot_info': (c_int, [c_dp, c_int, c_int, ctypes.POINTER(Knots)]),
    'rvs_locate_grid': (c_int, [ctypes.POINTER(GridMap), c_dp, c_i64, c_int, c_dp, c_dp, c_dp,
                                c_dp, c_dp]),
    'rvs_template_build': (c_int, [c_dp, c_int, c_i64, ctypes.POINTER(Knots), c_dp, c_dp, c_int,
                                   c_dp, c_int, c_int, c_dp, c_i64, c_dp, c_dp]),
    'rvs_obs_prepare': (c_int, [c_dp, c_dp, c_dp, c_int, c_dbl, c_dp, c_dp, c_dp, c_dp]),
    'rvs_basis_build': (c_int, [c_dp, c_dp, c_int, c_i64, c_int, c_int, c_int, c_dp, c_dp,
                                c_dp]),
    'rvs_chisq_scan': (c_int, [c_dp, c_i64, c_dp, ctypes.POINTER(Knots), ctypes.POINTER(Obs),
                               c_dp, c_dp, c_int, c_int, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                               c_int, c_dp]),
    'rvs_chisq_scan_ragged': (c_int, [c_dp, c_i64, c_dp, ctypes.POINTER(Knots),
                                      ctypes.POINTER(Obs), c_dp, c_dp, c_int, c_dp, c_int, c_dp,
                                      c_dp, c_dp, c_dp, c_dp, c_dp, c_int, c_dp]),
    'rvs_gridbox_init': (c_int, [ctypes.POINTER(GridBox), c_dp, c_i64, c_int, c_dp]),
    'rvs_fused_chunks': (c_int, [c_int, c_int]),
    'rvs_fused_workspace': (c_i64, [c_int, c_int, c_int]),
    'rvs_chisq_fused': (c_int, [c_dp, c_int, c_i64, ctypes.POINTER(Knots), c_dp, c_dp, c_int,
                                c_dp, c_dbl, c_int, ctypes.POINTER(Obs), c_dp, c_dp, c_int,
                                c_dp, c_i64, c_dp, c_dp, c_dp, ctypes.POINTER(GridBox), c_dp]),
    'rvs_chisq_fused_multi': (c_int, [ctypes.POINTER(FusedArm), c_int, c_dp, c_dp, c_int, c_dp, c_dbl,
                                      c_dp, c_int, c_dp]),
    'rvs_scan_stats': (c_int, [c_dp, c_dp, c_int, c_int, c_int, c_int, c_dp, c_dp, c_dp]),
    'rvs_scan_stats_ragged': (c_int, [c_dp, c_dp, c_int, c_int, c_int, c_dp, c_int, c_dp, c_dp,
                                      c_dp]),
    'rvs_ccf_workspace': (c_i64, [ctypes.POINTER(CcfArm), c_int]),
    'rvs_ccf_accumulate': (c_int, [ctypes.POINTER(CcfArm), c_dp, c_dp, c_int, c_dp, c_dp, c_dp,
                                   c_dp, c_i64, c_dp]),
    'rvs_ccf_prep_smem': (c_i64, [c_int, c_int]),
    'rvs_ccf_preprocess': (c_int, [c_dp, c_dp, c_dp, c_dp, c_int, c_int, c_dp, c_int, c_dp, c_dp,
                                   c_dp, c_int, c_int, c_dbl, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'rvs_nm_create': (ctypes.c_void_p, [c_int, c_int, c_dp, c_dbl, c_dbl, c_i64]),
    'rvs_nm_destroy': (None, [ctypes.c_void_p]),
    'rvs_nm_request': (c_i64, [ctypes.c_void_p, c_int, c_dp, c_dp, c_i64]),
    'rvs_nm_feed': (c_int, [ctypes.c_void_p, c_dp, c_i64]),
    'rvs_nm_live': (c_i64, [ctypes.c_void_p, c_dp]),
    'rvs_nm_result': (c_int, [ctypes.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'rvs_fit_pack': (c_int, [ctypes.POINTER(FitLayout), c_i64, c_i64, c_dp, c_dp, c_dp, c_dp, c_dp,
                             c_dp, c_dp, c_dp, c_dp]),
    'rvs_fit_collect': (c_i64, [ctypes.POINTER(FitLayout), c_i64, c_i64, c_dp, c_dp, c_dp, c_dp,
                                c_int, c_int, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'rvs_stream_create': (ctypes.c_void_p, [c_int]),
    'rvs_stream_destroy': (None, [ctypes.c_void_p]),
    'rvs_fit_round_items': (c_i64, [c_i64]),
    'rvs_drive_create': (ctypes.c_void_p, [c_i64, c_int]),
    'rvs_drive_destroy': (None, [ctypes.c_void_p]),
    'rvs_drive_request': (c_int, [ctypes.c_void_p, c_dp, c_dp, c_dp]),
    'rvs_nm_drive': (c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(FitLayout),
                             ctypes.POINTER(Drive)]),
    'rvs_bfgs_drive': (c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(FitLayout),
                               ctypes.POINTER(Drive)]),
    'rvs_bfgs_create': (ctypes.c_void_p, [c_int, c_int, c_dp, c_dp, c_dbl, c_i64]),
    'rvs_bfgs_destroy': (None, [ctypes.c_void_p]),
    'rvs_bfgs_request': (c_i64, [ctypes.c_void_p, c_dp, c_dp, c_i64]),
    'rvs_bfgs_feed': (c_int, [ctypes.c_void_p, c_dp, c_i64]),
    'rvs_bfgs_live': (c_i64, [ctypes.c_void_p, c_dp]),
    'rvs_bfgs_result': (c_int, [ctypes.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp]),
    'rvs_ccf_best': (c_int, [c_dp, c_dp, c_dp, c_int, c_int, c_int, c_dp, c_dp, c_dp]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get('RVS_LIB', LIB_PATH)   # tuning builds (csrc/Makefile OUT=...)
        if not os.path.exists(path):
            raise RuntimeError(
                f'{path} not found: build it with `python __graft_entry__.py build` '
                '(make -C rvspecfit_b200/csrc).  There is no CPU fallback.')
        L = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


class RvsError(RuntimeError):
    pass


def check(rc, what=''):
    if rc != 0:
        msg = lib().rvs_last_error().decode()
        raise RvsError(f'{what} failed with code {rc}: {msg}')


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('rvspecfit_b200 needs a CUDA device (sm_100a); there is no CPU path')
    return torch
