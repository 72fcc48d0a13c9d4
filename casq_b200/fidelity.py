"""Gate fidelities from batched propagators (north_star: "final states, populations and gate fidelities").

The reference's only fidelity is the Hellinger fidelity of sampled counts inside PulseOptimizer
(src/casq/optimizers/pulse_optimizer.py:501-503); there is no process / gate fidelity in casq.  These are the standard
leakage-aware definitions, evaluated on the device for a whole batch of propagators [B, d, d] as returned by
``B200PulseBackend.solve_propagators_batch`` (rotating frame of the model, lab basis):

    M      = P U_target^+ U P          restricted to the computational subspace (n = 2^k levels of the d)
    F_pro  = |Tr M|^2 / n^2
    F_avg  = (Tr(M M^+) + |Tr M|^2) / (n (n + 1))        (Pedersen, Moller, Molmer 2007; reduces to the usual
                                                          (n F_pro + 1) / (n + 1) for a unitary M)
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from .common.exceptions import CasqError


def computational_indices(dims: Sequence[int], levels: int = 2) -> np.ndarray:
    """Indices of the basis states whose every subsystem level is < ``levels`` (subsystem 0 = least-significant digit),
    ordered so that the restricted operator is in the usual qubit ordering."""
    dims = list(dims)
    idx = []
    for k in range(int(np.prod([min(levels, dd) for dd in dims]))):
        rem, full, stride = k, 0, 1
        for dd in dims:
            lv = rem % min(levels, dd)
            rem //= min(levels, dd)
            full += lv * stride
            stride *= dd
        idx.append(full)
    return np.asarray(idx, dtype=np.int64)


def gate_fidelities(U: torch.Tensor, target, dims: Optional[Sequence[int]] = None, levels: int = 2, virtual_z: bool = False):
    """(F_avg [B], F_pro [B], leakage [B]) of propagators U [B, d, d] against ``target`` [n, n] on the computational subspace.

    ``virtual_z`` (single qubit only): maximise over a frame update Z(phi) after the gate -- the phase a software virtual-Z
    absorbs -- which for a 2 x 2 M replaces |m00 + m11| by |m00| + |m11|.
    leakage = 1 - Tr(M M^+) / n: the population that leaves the computational subspace, averaged over its basis states.
    """
    if U.dim() != 3 or U.shape[1] != U.shape[2]:
        raise CasqError("gate_fidelities: U must be [B, d, d]")
    d = U.shape[1]
    dims = [d] if dims is None else list(dims)
    if int(np.prod(dims)) != d:
        raise CasqError(f"gate_fidelities: dims {dims} do not multiply to {d}")
    idx = torch.as_tensor(computational_indices(dims, levels), device=U.device)
    n = idx.numel()
    tgt = torch.as_tensor(np.asarray(target, dtype=complex), dtype=U.dtype, device=U.device)
    if tgt.shape != (n, n):
        raise CasqError(f"gate_fidelities: target must be [{n},{n}] for dims {dims}, got {tuple(tgt.shape)}")
    sub = U.index_select(1, idx).index_select(2, idx)
    M = tgt.conj().T.unsqueeze(0) @ sub
    diag = torch.diagonal(M, dim1=-2, dim2=-1)
    if virtual_z:
        if n != 2:
            raise CasqError("virtual_z is implemented for a single qubit")
        tr_abs = diag.abs().sum(-1)
    else:
        tr_abs = diag.sum(-1).abs()
    mm = (M.abs() ** 2).sum(dim=(-2, -1))
    f_pro = tr_abs**2 / n**2
    f_avg = (mm + tr_abs**2) / (n * (n + 1))
    return f_avg, f_pro, 1.0 - mm / n
