"""Parity of the CUDA cross-correlation path (rvs_ccf_accumulate / rvs_ccf_best
through fitter_ccf.fit) with the fixtures the reference's fitter_ccf.fit
produced (tests/golden/ccf.npz) and with the CPU oracle.  Needs a B200.

Tolerances: the chi-square curve of the best template to 1e-9 relative (the
reference's two inverse transforms are merged into one by linearity and cuFFT's
butterfly order differs from pocketfft's; both are fp64), best velocity to
1e-5 km/s (BASELINE.json asks 0.01 km/s)."""
import numpy as np
import pytest

import oracle
from helpers import close, config, setup, unpack_objects
from rvspecfit_b200 import fitter_ccf, make_ccf, spec_fit

pytestmark = pytest.mark.gpu

CASES = (('rvs', ('gaiarvs',), 600), ('two', ('desi_b', 'desi_r'), 1000))


def _banks(g, tag, shapes, splinestep=1000):
    """Register the reference-built model bank on the device and return the
    oracle's copy of it."""
    obanks = {}
    for k, s in enumerate(shapes):
        st = setup(s, 'tiny', 31 + k)
        c = g[f'{tag}_{s}_conf']
        conf = make_ccf.get_ccf_config(c[0], c[1], int(c[2]), splinestep=splinestep)
        models = g[f'{tag}_{s}_models']
        fft, fft2 = np.fft.rfft(models, axis=1), np.fft.rfft(models**2, axis=1)
        fitter_ccf.register_ccf_bank(s, fft, fft2, models, g[f'{tag}_{s}_params'],
                                     list(g[f'{tag}_{s}_vsinis']), st['parnames'], conf)
        obanks[s] = dict(fft=fft, fft2=fft2, models=models, params=g[f'{tag}_{s}_params'],
                         vsinis=list(g[f'{tag}_{s}_vsinis']), parnames=st['parnames'],
                         ccfconf=oracle.ccf_config(c[0], c[1], int(c[2]), splinestep=splinestep))
    return obanks


def _sd(obj, cls=spec_fit.SpecData):
    return [cls(nm, lam, sp, es, bad) for nm, lam, sp, es, bad in obj['arms']]


def test_ccf_fit_matches_reference(golden):
    g = golden('ccf')
    for tag, shapes, maxvel in CASES:
        cfg = config(max_vel=maxvel, vel_step0=2.5)
        _banks(g, tag, shapes)
        st = setup(shapes[0], 'tiny', 31)
        for i, o in enumerate(unpack_objects(g, tag + '_')):
            sd = _sd(o)
            res = fitter_ccf.fit(sd, cfg)
            for a in range(len(sd)):
                close(res['proc_spec'][sd[a].name], g[f'{tag}_{i}_{a}_proc_spec'], rtol=1e-7,
                      atol=1e-9, what='proc_spec')
            close(res['vel_grid'], g[f'{tag}_{i}_vel_grid'], rtol=0, atol=0)
            close(res['best_ccf'], g[f'{tag}_{i}_best_ccf'], rtol=1e-7, what='best_ccf')
            assert abs(res['best_vel'] - g[f'{tag}_{i}_best_vel']) < 1e-5
            assert res['best_vsini'] == g[f'{tag}_{i}_best_vsini']
            close([res['best_par'][k] for k in st['parnames']], g[f'{tag}_{i}_best_par'],
                  rtol=1e-12)


def test_ccf_batch_matches_oracle_on_reference_preprocessing(golden):
    """The device part alone: the reference's own proc_spec / proc_ivar go in, so
    the host-side continuum fit is out of the comparison; every output of the
    hot loop is held to the oracle, for all objects in one batch."""
    g = golden('ccf')
    for tag, shapes, maxvel in CASES:
        cfg = config(max_vel=maxvel, vel_step0=2.5)
        obanks = _banks(g, tag, shapes)
        objs = unpack_objects(g, tag + '_')
        pre = [{nm: (g[f'{tag}_{i}_{a}_proc_spec'], g[f'{tag}_{i}_{a}_proc_ivar'])
                for a, (nm, *_) in enumerate(o['arms'])} for i, o in enumerate(objs)]
        # ragged batch: object 0, object 1, and object 0 with its first arm only
        sds = [_sd(o) for o in objs] + [_sd(objs[0])[:1]]
        pre.append(pre[0])
        got = fitter_ccf.fit_batch(sds, cfg, preprocessed=pre)
        for i, (sd, res) in enumerate(zip(sds, got)):
            osd = [oracle.SpecData(s.name, s.lam, s.spec, s.espec, s.badmask) for s in sd]
            want = oracle.ccf_fit(osd, cfg, obanks, preprocessed=pre[i])
            assert res['best_id'] == want['best_id']
            close(res['best_ccf'], want['best_ccf'], rtol=1e-9, what='best_ccf')
            assert abs(res['best_vel'] - want['best_vel']) < 1e-6
            assert res['best_vsini'] == want['best_vsini']
            for s in sd:
                close(res['best_model'][s.name], want['best_model'][s.name], rtol=0, atol=0)
        if len(objs[0]['arms']) == 1:       # golden value of the reference itself
            close(got[0]['best_ccf'], g[f'{tag}_0_best_ccf'], rtol=1e-9)


def test_device_preprocessing_matches_reference(golden):
    """Row f3 on the device (rvs_ccf_preprocess: masks, gap bridging, medians, soft-L1
    continuum fit, resampling) against the reference's make_ccf.preprocess_data outputs.
    Everything but the continuum fit is the same arithmetic as the reference; the fit is a
    damped Gauss-Newton run to a tighter stop than scipy's least_squares (ftol = xtol =
    gtol = 1e-8), so the two continua agree to what the reference's own stop leaves:
    measured <= 1.2e-6 of the peak on these fixtures, tolerance 1e-5.  The first guess
    it feeds must not move: best velocity to 0.01 km/s (BASELINE.json), same template."""
    g = golden('ccf')
    for tag, shapes, maxvel in CASES:
        cfg = config(max_vel=maxvel, vel_step0=2.5)
        _banks(g, tag, shapes)
        objs = unpack_objects(g, tag + '_')
        for i, o in enumerate(objs):
            for a, (nm, lam, sp, es, bad) in enumerate(o['arms']):
                c = g[f'{tag}_{nm}_conf']
                conf = make_ccf.get_ccf_config(c[0], c[1], int(c[2]))
                prep = make_ccf.DevicePrep(lam, conf)
                assert prep.ok, prep.why
                d_ps, d_pi, d_cont, d_info = prep(sp[None], es[None], bad[None], want_cont=True)
                ps, pi = d_ps.cpu().numpy()[0], d_pi.cpu().numpy()[0]
                info = d_info.cpu().numpy()[0]
                assert 1 <= info[0] < 60 and info[1] == 0, info
                want_s, want_i = g[f'{tag}_{i}_{a}_proc_spec'], g[f'{tag}_{i}_{a}_proc_ivar']
                assert np.abs(ps - want_s).max() <= 1e-5 * np.abs(want_s).max()
                assert np.abs(pi - want_i).max() <= 1e-5 * np.abs(want_i).max()
                # exactly the same pixels carry weight
                assert np.array_equal(pi == 0, want_i == 0)
                # the host route of this package (the reference's own steps) agrees too
                hs, hi = make_ccf.preprocess_data(lam, sp, es, ccfconf=conf, badmask=bad)
                assert np.abs(ps - hs).max() <= 1e-5 * np.abs(hs).max()
        sds = [_sd(o) for o in objs] + [_sd(objs[0])[:1]]
        dev = fitter_ccf.fit_batch(sds, cfg, preprocess='device')
        host = fitter_ccf.fit_batch(sds, cfg, preprocess='host')
        for i, (d, h) in enumerate(zip(dev, host)):
            assert d['best_id'] == h['best_id'] and d['best_vsini'] == h['best_vsini']
            assert abs(d['best_vel'] - h['best_vel']) < 0.01
            close(d['best_ccf'], h['best_ccf'], rtol=1e-4, what='best_ccf (device preprocessing)')
            if i < len(objs):
                assert abs(d['best_vel'] - g[f'{tag}_{i}_best_vel']) < 0.01


def test_device_preprocessing_edge_cases():
    """Masked runs at both ends and in the middle, a non-positive stretch (median filter
    mask), an error spike, and a spectrum with every pixel masked: the device route
    against the host route of this package (a mirror of the reference)."""
    rs = np.random.RandomState(4)
    lam = np.arange(8460., 8700., 0.1)
    n = len(lam)
    conf = make_ccf.get_ccf_config(np.log(8410.), np.log(8750.), 4096)
    prep = make_ccf.DevicePrep(lam, conf)
    assert prep.ok
    base = 50 * (1 + 0.2 * np.sin(lam / 30.)) * (1 - 0.5 * np.exp(-0.5 * ((lam - 8542) / 1.5)**2))
    specs, errs, bads = [], [], []
    for k in range(5):
        es = 1.0 + 0.3 * rs.uniform(size=n)
        sp = base + es * rs.normal(size=n)
        bad = np.zeros(n, dtype=bool)
        if k == 0:
            bad[:40] = True
            bad[-25:] = True
            bad[700:760] = True
        if k == 1:
            sp[1000:1030] = -5.0            # median filter mask
            es[300] = 500.0                 # error spike
        if k == 2:
            bad[rs.choice(n, 400, replace=False)] = True
        if k == 3:
            bad[:] = True
        specs.append(sp), errs.append(es), bads.append(bad)
    d_ps, d_pi, d_cont, d_info = prep(np.array(specs), np.array(errs), np.array(bads),
                                      want_cont=True)
    ps, pi, info = d_ps.cpu().numpy(), d_pi.cpu().numpy(), d_info.cpu().numpy()
    assert info[3, 1] == 1 and not info[[0, 1, 2, 4], 1].any()
    for k in range(5):
        hs, hi = make_ccf.preprocess_data(lam, specs[k], errs[k], ccfconf=conf, badmask=bads[k])
        if k == 3:          # nothing to fit: no weight anywhere, as on the host
            assert not pi[k].any() and not hi.any()
            continue
        assert np.abs(ps[k] - hs).max() <= 1e-5 * np.abs(hs).max(), k
        assert np.abs(pi[k] - hi).max() <= 1e-5 * np.abs(hi).max(), k
        assert np.array_equal(pi[k] == 0, hi == 0), k


def test_ccf_ratio_mode_and_small_workspace(golden, monkeypatch):
    """ccf_continuum_normalize = False (chi2 = -ccf0^2/ccf1, two inverse
    transforms) and a workspace so small that objects go through one at a time."""
    g = golden('ccf')
    tag, shapes, maxvel = CASES[1]
    cfg = config(max_vel=maxvel, vel_step0=2.5)
    obanks = _banks(g, tag, shapes, splinestep=None)
    objs = unpack_objects(g, tag + '_')
    pre = [{nm: (g[f'{tag}_{i}_{a}_proc_spec'], g[f'{tag}_{i}_{a}_proc_ivar'] + 1e-3)
            for a, (nm, *_) in enumerate(o['arms'])} for i, o in enumerate(objs)]
    sds = [_sd(o) for o in objs]
    monkeypatch.setattr(fitter_ccf, 'WORKSPACE_BYTES', 1)
    fitter_ccf._ws.clear()
    got = fitter_ccf.fit_batch(sds, cfg, preprocessed=pre)
    for i, (sd, res) in enumerate(zip(sds, got)):
        osd = [oracle.SpecData(s.name, s.lam, s.spec, s.espec, s.badmask) for s in sd]
        want = oracle.ccf_fit(osd, cfg, obanks, preprocessed=pre[i])
        assert res['best_id'] == want['best_id']
        close(res['best_ccf'], want['best_ccf'], rtol=1e-8, what='best_ccf (ratio mode)')
        assert abs(res['best_vel'] - want['best_vel']) < 1e-5
    fitter_ccf._ws.clear()


def test_ccf_failure_raises(golden):
    """All-NaN input -> the reference's RuntimeError (fitter_ccf.py:224-226)."""
    g = golden('ccf')
    tag, shapes, maxvel = CASES[0]
    cfg = config(max_vel=maxvel, vel_step0=2.5)
    _banks(g, tag, shapes)
    o = unpack_objects(g, tag + '_')[0]
    sd = _sd(o)
    n = int(g[f'{tag}_{shapes[0]}_conf'][2])
    pre = [{shapes[0]: (np.full(n, np.nan), np.ones(n))}]
    with pytest.raises(RuntimeError, match='Cross-correlation step failed'):
        fitter_ccf.fit_batch([sd], cfg, preprocessed=pre)
