"""Host-side Rippe contact-law fit (SciPy), restating /root/reference/optim_rippe_curve_update.py.

Only the INPUT of the fit (the distance histogram, cuda_lib_gl.py:1236-1270) moves to the device
(graal_dist_histogram); the 4-parameter least-squares fit and the fsolve for the cis/trans
cross-over distance stay on the host, as in the reference.
"""
import numpy as np
from scipy.optimize import leastsq, fsolve

D_CONST = 3   # module-level ``d`` of optim_rippe_curve_update.py:9


def rippe_curve(x, kuhn, lm, slope, d, A):
    with np.errstate(all="ignore"):
        n = lm * x / kuhn
        return A * (0.53 * (kuhn ** -3.) * np.power(n, slope) * np.exp((d - 2) / (np.power(n, 2) + d)))


def peval(x, param):
    """optim_rippe_curve_update.py:22-28: param = [kuhn, lm, slope, A(, ...)], d = module constant.
    NOTE (reference quirk Q10): callers that pass [kuhn, lm, slope, d, fact] get ``d`` as amplitude."""
    return rippe_curve(x, param[0], param[1], param[2], D_CONST, param[3])


def _log_residuals(p, y, x):
    kuhn, lm, slope, A = p
    with np.errstate(all="ignore"):
        model = np.log(A) + np.log(0.53) - 3 * np.log(kuhn) + slope * (np.log(lm * x) - np.log(kuhn)) + \
            (D_CONST - 2) / (np.power(lm * x / kuhn, 2) + D_CONST)
    return y - model


def estimate_param_rippe(y_meas, x_bins):
    """optim_rippe_curve_update.py:73-115 -> ([kuhn, lm, slope, d, A], fitted curve)."""
    kuhn, lm, slope = 1, 9.6, -1.5
    A = np.sum(y_meas)
    sol = leastsq(_log_residuals, [kuhn, lm, slope, A], args=(np.log(y_meas), x_bins))[0]
    y_estim = peval(x_bins, sol)
    out = [sol[0], sol[1], sol[2], D_CONST, sol[3]]
    if np.any(np.isnan(np.array(out))) or slope >= 0:     # the reference tests the initial constant
        out = [kuhn, lm, slope, D_CONST, A]
    return out, y_estim


def estimate_max_dist_intra(p, val_inter):
    """optim_rippe_curve_update.py:117-135: distance (kb) where the cis law meets the trans level."""
    kuhn, lm, slope, d, A = p
    f = lambda x: val_inter - rippe_curve(x, kuhn, lm, slope, d, A)
    return fsolve(f, 500)[0]
