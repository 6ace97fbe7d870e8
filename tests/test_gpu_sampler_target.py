"""GPU: the device-resident PT sampler (ttvb_pt_*) on a target with a KNOWN answer.

A single planet transits at t_n = t_c + n P exactly, so with the period and the mean anomaly free
(everything else constant) the transit times are linear in the two hypercube coordinates over the
width of the posterior: chi2 is a quadratic form and the tempered posteriors are the Gaussians
N(theta_hat, Sigma / beta_t), with theta_hat and Sigma from weighted least squares. A sampler whose
stretch moves, acceptance rule, walker pairing or temperature swaps were biased would miss them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _one_planet_spec(oracle):
    # free: the period and slot 4, which in the reference's parameterisation (utils.py:403-438) is
    # argument + mean anomaly, i.e. the orbital phase of a circular orbit; slot 5 (argument - mean
    # anomaly) and everything else are constants
    P, x4, c5 = 10.0, 60.0, 10.0       # first transit 0.83 d after t0
    const = {"0": 5.0, "2": 0.0, "3": 90.0, "5": c5, "6": 0.0}
    spec = {"system_name": "one", "mstar": 1.0, "rstarAU": 0.00465, "t0": 0.0, "ftime": 205.0, "dt": 0.5,
            "npla": 1, "ndim": 2, "planets_IDs": {"Planet-b": 0},
            "bounds": [[9.99, 10.01], [50.0, 70.0]], "bi": [9.99, 50.0], "bf": [10.01, 70.0],
            "hypercube": [[0.0, 1.0], [0.0, 1.0]], "constant_params": const,
            "observed": {}, "second_term_sum": 0.0}
    flat = [5.0, P, 0.0, 90.0, (x4 + c5) / 2.0, (x4 - c5) / 2.0, 0.0]
    st, pl, ep, tm, rs, vs = oracle.ephemeris(flat, spec["mstar"], 0.0, spec["dt"], spec["ftime"])
    assert st == 0 and len(tm) >= 20
    ep, tm = ep[:19], tm[:19]                            # keep clear of the end of the integration
    sigma = 5e-4
    rng = np.random.default_rng(77)
    spec["observed"] = {"Planet-b": {"planet_index": 0, "epochs": [int(e) for e in ep],
                                     "times": [float(t) for t in tm + sigma * rng.standard_normal(len(tm))],
                                     "sigma": [sigma] * len(tm)}}
    truth_cube = np.array([(P - 9.99) / 0.02, (x4 - 50.0) / 20.0])
    return spec, truth_cube, sigma


def _least_squares(spec, truth_cube, sigma, oracle):
    """theta_hat and Sigma of chi2(theta) in cube coordinates: Gauss-Newton on the oracle's transit
    times (linear in theta to ~1e-9 d over the posterior, so one step from the truth is exact)."""
    osys = oracle.OracleSystem(spec)
    obs = np.array(spec["observed"]["Planet-b"]["times"])

    def times(c):
        r = oracle.ephemeris(osys.cube_to_physical(c), spec["mstar"], 0.0, spec["dt"], spec["ftime"])
        return r[3][:len(obs)]
    t0 = times(truth_cube)
    h = np.array([1e-3, 1e-3])
    A = np.stack([(times(truth_cube + h * np.eye(2)[k]) - times(truth_cube - h * np.eye(2)[k])) / (2 * h[k])
                  for k in range(2)], axis=1)
    # our chi2 = sum ((obs - t)/sigma)^2 and logL = -chi2/2 (+const) -> Sigma = sigma^2 (A^T A)^-1
    N = A.T @ A
    theta = truth_cube + np.linalg.solve(N, A.T @ (obs - t0))
    return theta, sigma ** 2 * np.linalg.inv(N)


def test_device_sampler_recovers_a_gaussian_posterior_at_every_temperature(oracle):
    import nauyaca_b200 as nb
    from nauyaca_b200.ptsampler import DeviceSampler
    spec, truth, sigma = _one_planet_spec(oracle)
    theta, Sigma = _least_squares(spec, truth, sigma, oracle)
    sd = np.sqrt(np.diag(Sigma))
    assert np.all(sd < 5e-3) and np.all(np.abs(theta - truth) < 5 * sd)       # well inside the unit hypercube
    eng = nb.BatchEngine(spec)
    betas = np.array([1.0, 0.5, 0.2])
    nw, iters = 32, 3000
    smp = DeviceSampler(eng, nwalkers=nw, betas=betas, a=2.0, seed=20240607)
    rng = np.random.default_rng(3)
    L = np.linalg.cholesky(Sigma)
    p0 = theta[None, None, :] + 2.0 * (rng.standard_normal((3, nw, 2)) @ L.T)
    chain = np.empty((iters, 3, nw, 2))
    for i, (p, logpost, logl) in enumerate(smp.sample(p0, iterations=iters, adapt=False, storechain=False)):
        chain[i] = p
    assert np.array_equal(smp.betas, betas)                                   # adapt=False: the ladder stays
    acc = smp.acceptance_fraction.mean(axis=1)
    assert np.all(acc > 0.3) and np.all(acc < 0.9)
    assert np.all(smp.tswap_acceptance_fraction > 0.2)
    # integrated autocorrelation time of the ensemble mean -> effective sample size
    for t, beta in enumerate(betas):
        x = chain[iters // 3:, t]                                             # (n, nw, 2), burn-in dropped
        flat = x.reshape(-1, 2)
        mean, cov = flat.mean(axis=0), np.cov(flat.T)
        want = Sigma / beta
        m = x.mean(axis=1)                                                    # ensemble mean per iteration
        tau = max(1.0, max(_iat(m[:, k]) for k in range(2)))
        neff = flat.shape[0] / tau
        se_mean = np.sqrt(np.diag(want) / neff)
        assert np.all(np.abs(mean - theta) < 5 * se_mean + 0.02 * np.sqrt(np.diag(want))), (t, mean, theta, se_mean)
        ratio = np.diag(cov) / np.diag(want)
        tol = 5 * np.sqrt(2.0 / neff) + 0.03
        assert np.all(np.abs(ratio - 1.0) < tol), (t, ratio, tol, neff)
        corr = cov[0, 1] / np.sqrt(cov[0, 0] * cov[1, 1]); wcorr = want[0, 1] / np.sqrt(want[0, 0] * want[1, 1])
        assert abs(corr - wcorr) < 0.05, (t, corr, wcorr)
    smp.close(); eng.close()


def _iat(x, window=200):
    x = np.asarray(x) - np.mean(x)
    n = len(x)
    f = np.fft.rfft(x, 2 * n)
    acf = np.fft.irfft(f * np.conj(f))[:n]
    acf /= acf[0]
    return 1.0 + 2.0 * np.sum(acf[1:window])
