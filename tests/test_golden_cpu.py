"""The oracle reproduces the committed golden vectors (tests/golden/make_golden.py wrote them)."""
import os

import numpy as np

import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_envelopes_golden():
    g = np.load(os.path.join(G, "envelopes.npz"))
    for kind, keys in [("constant", ()), ("gaussian", ("sigma",)), ("drag", ("sigma", "beta")),
                       ("gaussian_square", ("sigma", "width")), ("gaussian_square_drag", ("sigma", "width", "beta"))]:
        got = O.envelope_samples(kind, int(g["duration"]), amp=g["p_amp"], angle=g["p_angle"], **{k: g[f"p_{k}"] for k in keys})
        assert np.array_equal(got, g[kind])


def test_rk4_golden(manila_q0):
    om, st, ops, ch = manila_q0
    g = np.load(os.path.join(G, "rk4_drag_d3.npz"))
    sel = slice(0, 6)
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", int(g["duration"]))],
                                [dict(amp=g["amp"][sel], angle=0.0, sigma=float(g["sigma"]), beta=g["beta"][sel])])
    ts, y = O.rk4_solve(om, samples, np.array([1, 0, 0]), (0.0, int(g["t_final"]) * O.MANILA_DT), O.MANILA_DT / 40, g["t_eval"])
    assert np.abs(y - g["raw"][sel]).max() < 1e-14
    assert np.abs(O.postprocess_statevectors(om, ts, y) - g["states"][sel]).max() < 1e-13
