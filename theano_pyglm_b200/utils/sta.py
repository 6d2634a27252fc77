"""Spike-triggered average (interface of pyglm/utils/sta.py) and the basis projection that turns it into initial
stimulus weights (pyglm/utils/basis.py:416-436).  Host NumPy: it runs once, before MAP."""
import numpy as np


def sta(stim, data, L, Ns=None):
    """A[i, l, d] = sum_t S[t, Ns[i]] * istim[t - l, d] / sum_t S[t, Ns[i]]  (sta.py:6-84): `istim` is the stimulus
    interpolated at the spike bins and divided by dt_stim/dt (sta.py:34-38), zero before the recording starts."""
    S = np.asarray(data['S'], dtype=np.float64)
    nt, N = S.shape
    Ns = list(range(N)) if Ns is None else ([Ns] if isinstance(Ns, (int, np.integer)) else list(Ns))
    D = stim.shape[1]
    dt, dt_stim = data['dt'], data['dt_stim']
    t = dt * np.arange(nt)
    t_stim = dt_stim * np.arange(stim.shape[0])
    istim = np.stack([np.interp(t, t_stim, stim[:, d]) for d in range(D)], axis=1) / (dt_stim / dt)
    istim = np.vstack([np.zeros((L, D)), istim])
    Sn = S[:, Ns]                                                      # (nt, n)
    A = np.empty((len(Ns), L, D))
    for l in range(L):                                                 # lag l looks l bins back
        A[:, l, :] = Sn.T @ istim[L - l:L - l + nt, :]
    return A / np.sum(Sn, axis=0)[:, None, None]


def project_onto_basis(f, basis, lam=0):
    """Least-squares coefficients of f (R,) or (R, k) on basis (R, B), optional ridge lam (basis.py:416-436)."""
    R, B = basis.shape
    f = np.asarray(f, dtype=np.float64)
    assert f.shape[0] == R, "Function is not the same length as the basis!"
    if f.ndim == 1:
        f = f.reshape(R, 1)
    return np.linalg.inv(basis.T @ basis + lam * np.eye(B)) @ basis.T @ f
