"""DESI spectra -> SpecData: the data conditioning the reference's DESI driver applies
between the files and the likelihood (reference desi/desi_fit.py:682-888, SURVEY.md
section 8 row f4): masking, bridging of masked runs, the error assigned to masked
pixels, the clamp on implausibly small errors, the dichroic mask, and -- in
resolution-matrix mode -- the deconvolved, edge-normalised banded matrix of every
spectrum.  Host code, as in the reference (it runs once per spectrum, before the hot
path); the SpecData it returns feed spec_fit / vel_fit / fitter_ccf of this package,
which apply the matrices on the device."""
import logging

import numpy as np
import scipy.linalg
import scipy.sparse

from . import spec_fit

LARGE_ERROR = 1000      # error of masked pixels, in units of the median flux
MINERR_FRAC = 0.3       # errors below this fraction of the median error are clamped
EDGE_PIXELS = 5         # masked at both ends in resolution-matrix mode


def band_to_rows(mat):
    """DESI stores the band of the resolution matrix by columns; this gives it by rows
    (desi_fit.py:682-685)."""
    w = mat.shape[0]
    return np.array([np.roll(mat[k], k - w // 2) for k in range(w)])[::-1]


def band_to_columns(rows):
    """Inverse of band_to_rows (desi_fit.py:688-691)."""
    w = rows.shape[0]
    flipped = rows[::-1]
    return np.array([np.roll(flipped[k], w // 2 - k) for k in range(w)])


def deconvolve_resolution_matrix(mat0, sigma0_angstrom=0.5, pix_size_angstrom=0.8):
    """Remove a Gaussian of sigma0 from the line-spread function the band describes: the
    templates already carry that much resolution (desi_fit.py:694-720)."""
    width, npix = mat0.shape
    sig = sigma0_angstrom / pix_size_angstrom
    xs = np.arange(width)
    gau = np.array([1. / np.sqrt(2 * np.pi) / sig * np.exp(-0.5 * ((xs - i) / sig)**2)
                    for i in range(width)])
    w2 = width // 2
    rows = band_to_rows(mat0)
    for i in range(w2):             # band entries that fall off the matrix
        rows[:w2 - i - 1, i] = 0
        rows[w2 + 1 + i:, npix - 1 - i] = 0
    return band_to_columns(scipy.linalg.solve(gau, rows))


def construct_resolution_sparse_matrix(mat, pix_size_angstrom=None, sigma0_angstrom=None):
    """Band data of one spectrum [width, npix] -> scipy.sparse.dia_matrix, deconvolved and
    with the truncated rows at both ends renormalised (desi_fit.py:723-748)."""
    width, npix = mat.shape
    w2 = width // 2
    mat = deconvolve_resolution_matrix(mat.copy(), pix_size_angstrom=pix_size_angstrom,
                                       sigma0_angstrom=sigma0_angstrom)
    rows = band_to_rows(mat)
    mult = np.median(rows.sum(axis=0))
    if mult == 0:
        mult = 1
    for i in range(w2):
        n1 = rows[w2 - i:, i].sum()
        rows[:, i] = rows[:, i] / (n1 + (n1 == 0)) * mult
        j = npix - 1 - i
        n2 = rows[:w2 + 1 + i, j].sum()
        rows[:, j] = rows[:, j] / (n2 + (n2 == 0)) * mult
    return scipy.sparse.dia_matrix((band_to_columns(rows), np.arange(w2, -w2 - 1, -1)),
                                   (npix, npix))


def interpolate_bad_regions(spec, mask):
    """Masked runs bridged linearly in pixel index, flat at the ends
    (desi_fit.py:751-778)."""
    bad = np.flatnonzero(mask)
    if len(bad) == 0 or len(bad) == len(spec):
        return spec
    out = spec * 1
    starts = np.flatnonzero(np.diff(bad, prepend=-10) > 1)
    ends = np.append(starts[1:] - 1, len(bad) - 1)
    for lh, rh in zip(bad[starts], bad[ends]):
        if lh == 0:
            out[:rh + 1] = spec[rh + 1]
        elif rh == len(spec) - 1:
            out[lh:] = spec[lh - 1]
        else:
            out[lh:rh + 1] = np.interp(np.arange(lh, rh + 1), [lh - 1, rh + 1],
                                       [spec[lh - 1], spec[rh + 1]])
    return out


def get_specdata(waves, fluxes, ivars, masks, resolutions, seqid, setups,
                 use_resolution_matrix=False, mask_dicroic=True, lsf_sigma0_angstrom=None):
    """The SpecData tuple of one fibre (None if nothing is usable): reference
    desi/desi_fit.py:781-888, same arguments and results."""
    sds = []
    for s in setups:
        spec = fluxes[s][seqid] * 1
        ivar = ivars[s][seqid] * 1
        badmask = masks[s][seqid] > 0
        med = np.nanmedian(spec)
        if badmask.all():
            continue
        if med == 0:
            med = np.nanmedian(spec[(spec > 0) & (~badmask)])
            if not np.isfinite(med):
                med = np.nanmedian(np.abs(spec))
        if not np.isfinite(med) or med == 0:
            continue
        baddat = ~np.isfinite(spec + ivar)
        dichroic = ((waves[s] > 4300) & (waves[s] < 4450)) if mask_dicroic else \
            np.zeros(len(waves[s]), dtype=bool)
        baderr = ivar <= 0
        edge = np.zeros(len(spec), dtype=bool)
        resol = None
        if use_resolution_matrix:
            resol = spec_fit.ResolMatrix(construct_resolution_sparse_matrix(
                resolutions[s][seqid], pix_size_angstrom=waves[s][1] - waves[s][0],
                sigma0_angstrom=lsf_sigma0_angstrom[s]))
            edge[:EDGE_PIXELS] = True       # the matrix is unreliable at the ends
            edge[-EDGE_PIXELS:] = True
        badall = baddat | badmask | baderr | dichroic | edge
        ivar[badall] = 1. / med**2 / LARGE_ERROR**2
        spec[:] = interpolate_bad_regions(spec, baddat | badmask | baderr)
        espec = 1. / ivar**.5
        if badall.all():
            logging.warning('The whole spectrum was masked...')
        else:
            thresh = np.median(espec[~badall]) * MINERR_FRAC
            small = (espec < thresh) & (~badall)
            if small.sum() / (~badall).sum() > .01:
                logging.warning('More than 1% of spectra had the uncertainty clamped')
            espec[small] = thresh
        sds.append(spec_fit.SpecData('desi_%s' % s, waves[s], spec, espec, resolution=resol,
                                     badmask=badall))
    if not sds:
        logging.warning(f'No good data found for fiber {seqid}')
        return None
    return tuple(sds)
