"""C-ABI library: loads, exports every symbol include/casq_b200.h declares, host-side model preparation
matches the oracle, and bad arguments come back as status codes + messages (no GPU needed)."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import oracle as O
from casq_b200 import _native
from casq_b200.common.exceptions import CasqError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "casq_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(casq_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 10
    cdll = ctypes.CDLL(str(_native.lib_path()))
    missing = [n for n in declared if not hasattr(cdll, n)]
    assert not missing, missing
    assert set(_native._SIGNATURES) == set(declared)
    assert native_lib.casq_version() >= 100


def test_library_is_sm100a_only(native_lib):
    import shutil
    import subprocess

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", str(_native.lib_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("subsystems,frame", [([0], "full"), ([0, 1], "full"), ([0, 1], "diag"), ([1, 2], "none")])
def test_model_preparation_matches_oracle(native_lib, subsystems, frame):
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, subsystems)
    rf = {"full": st, "diag": np.diag(st).real, "none": None}[frame]
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, O.MANILA_DT, rotating_frame=rf)
    m = _native.NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], O.MANILA_DT, rotating_frame=rf)
    d = st.shape[0]
    assert m.dim == d and native_lib.casq_model_dim(m.handle) == d
    scale = np.abs(st).max()
    U = m.frame_basis
    assert np.abs(U.conj().T @ U - np.eye(d)).max() < 1e-13
    if rf is not None:
        hf = np.diag(rf) if np.ndim(rf) == 1 else rf
        assert np.abs(U.conj().T @ hf @ U - np.diag(m.frame_evals)).max() / scale < 1e-14
        assert np.abs(np.sort(m.frame_evals) - np.sort(om.frame_evals)).max() / scale < 1e-14
    de, V = om.dressed()
    assert np.abs(m.dressed_evals - de).max() / scale < 1e-14
    assert np.abs(m.dressed_states - V).max() < 1e-9
    m.close()


def test_model_create_rejects_bad_arguments(native_lib):
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    with pytest.raises(CasqError, match="dt must be > 0"):
        _native.NativeModel(st, ops, [1e9, 1e9], 0.0)
    with pytest.raises(CasqError, match="carrier"):
        _native.NativeModel(st, ops, [1e9], 1e-10)
    with pytest.raises(CasqError, match="rotating_frame"):
        _native.NativeModel(st, ops, [1e9, 1e9], 1e-10, rotating_frame=np.zeros((2, 2)))
    h = ctypes.c_void_p()
    rc = native_lib.casq_model_create(ctypes.byref(h), 0, 0, None, None, None, 1e-10, 0, None, math.nan, 0)
    assert rc == -1 and b"dim" in native_lib.casq_last_error()
    assert native_lib.casq_model_invalidate_cache(None) == -1
    assert native_lib.casq_envelope_sample(9, 16, 4, None, 0, 0.0, None, None) == -1 and b"kind" in native_lib.casq_last_error()


def test_rwa_masks_match_oracle(native_lib):
    """RWA changes results only through the masks; check them through a silent probe: a model whose cutoff
    removes everything must equal (in dressed/frame read-backs) the unmasked one, and creation must succeed for
    NaN (= no RWA) and finite cutoffs."""
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    nu = [O.MANILA_CARRIERS[c] for c in ch]
    a = _native.NativeModel(st, ops, nu, O.MANILA_DT, rotating_frame=st, rwa_cutoff_freq=None)
    b = _native.NativeModel(st, ops, nu, O.MANILA_DT, rotating_frame=st, rwa_cutoff_freq=4.9e9)
    assert np.array_equal(a.frame_evals, b.frame_evals)
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, O.MANILA_DT, rotating_frame=st, rwa_cutoff_freq=4.9e9)
    # oracle side: the 0<->1 and 1<->2 co-rotating elements survive, counter-rotating ones do not
    assert np.count_nonzero(om.pos[0]) == 2 and np.count_nonzero(om.neg[0]) == 2
    assert np.count_nonzero(om.pos[0] * om.neg[0]) == 0


def test_dissipators_accepted(native_lib):
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    m = _native.NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], O.MANILA_DT, rotating_frame=np.diag(st).real)
    a = np.diag(np.sqrt([1.0, 2.0]), k=1)
    m.set_dissipators(np.array([a * 100.0, a.T @ a * 50.0]))
    with pytest.raises(Exception):
        m.set_dissipators(np.zeros((1, 2, 2)))


def test_solver_entry_points_validate_arguments_before_touching_the_gpu(native_lib):
    """Every solver entry point rejects a NULL handle / inconsistent arguments with a status and a message (no CUDA call
    is needed for that, so this runs on the CPU box); B = 0 is a successful no-op."""
    L = native_lib
    assert L.casq_solve_sv_rk4(None, 0, None, 0, None, None, 0, 0.0, 1.0, 0.1, 2, None, None, None) == -1
    assert b"NULL model" in L.casq_last_error()
    assert L.casq_solve_sv_dopri5(None, 0, None, 0, None, None, 0, 0.0, 1.0, 1e-6, 1e-6, 0.0, 0, 2, None, None, None, None, None) == -1
    assert b"NULL model" in L.casq_last_error()
    assert L.casq_propagators_expm(None, 0, None, 0, None, 0.0, 1.0, 0.1, None, None) == -1
    assert b"NULL model" in L.casq_last_error()
    assert L.casq_solve_lindblad_rk4(None, 0, None, 0, None, None, 0, 0.0, 1.0, 0.1, 2, None, None, None, None) == -1
    assert b"NULL model" in L.casq_last_error()
    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    m = _native.NativeModel(st, ops, [O.MANILA_CARRIERS[c] for c in ch], O.MANILA_DT, rotating_frame=st)
    h = m.handle
    # reversed time span and negative batch are caught before any launch; an empty batch succeeds
    assert L.casq_solve_sv_rk4(h, 0, None, 4, None, None, 0, 1.0, 0.0, 0.1, 2, None, None, None) < 0
    assert L.casq_solve_sv_dopri5(h, 0, None, -1, None, None, 0, 0.0, 1.0, 1e-6, 1e-6, 0.0, 0, 2, None, None, None, None, None) < 0
    assert b"negative" in L.casq_last_error()
    assert L.casq_propagators_expm(h, 0, None, 0, None, 0.0, 1e-9, 1e-10, None, None) == 0
    m.close()
