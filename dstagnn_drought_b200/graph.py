"""Host-side graph preprocessing, run once at model-build time (NumPy/SciPy).

Same results as the reference's lib/utils.py:149-203; kept on the host because it is
O(N^2) work done once.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch


def scaled_Laplacian(W):
    """2 L / lambda_max - I, L = diag(rowsum(W)) - W  (lib/utils.py:149-177).

    lambda_max is the eigenvalue with the largest real part (ARPACK, ``which='LR'``),
    as in the reference; W may be a numpy array or a torch tensor and the result keeps
    that kind (tensor results are float32 on W's device).

    The reference hands ARPACK the float32 adjacency and lets it draw a random start vector, so
    its own lambda_max -- and with it T_1, T_2 -- moves by ~3e-6 (relative) from run to run and from
    machine to machine.  Here the eigenvalue is computed in float64 from a fixed start vector: one
    value, inside the reference's own spread, identical on every rank and every box."""
    from scipy.sparse.linalg import eigs

    is_tensor = torch.is_tensor(W)
    w = W.detach().cpu().numpy() if is_tensor else np.asarray(W)
    if w.ndim != 2 or w.shape[0] != w.shape[1]:
        raise AssertionError("adjacency matrix must be square")
    n = w.shape[0]
    lap = np.diag(w.sum(axis=1)) - w
    lam_max = eigs(lap.astype(np.float64), k=1, which="LR",
                   v0=np.random.RandomState(0).uniform(-1.0, 1.0, n))[0].real
    lam_max = lap.dtype.type(lam_max) if np.issubdtype(lap.dtype, np.floating) else lam_max
    lt = 2.0 * lap / lam_max - np.eye(n)
    if is_tensor:
        return torch.from_numpy(lt).float().to(W.device)
    return lt


def cheb_polynomial(L_tilde, K: int) -> List[np.ndarray]:
    """[T_0, ..., T_{K-1}] with the reference's ELEMENT-WISE recurrence
    T_k = 2 L~ (.) T_{k-1} - T_{k-2}  (lib/utils.py:198-203; Hadamard product, sic)."""
    lt = L_tilde.detach().cpu().numpy() if torch.is_tensor(L_tilde) else np.asarray(L_tilde)
    polys = [np.eye(lt.shape[0]), np.array(lt, copy=True)]
    while len(polys) < K:
        polys.append(2.0 * lt * polys[-1] - polys[-2])
    return polys
