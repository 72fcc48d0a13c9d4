"""Parallel oracle solves for the GPU tests: chunks run in plain subprocesses (`python tests/_oracle_pool.py in out`) that
import only numpy/scipy + oracle -- the parent holds a CUDA context, so neither fork nor a re-imported pytest main is
involved.  TEST INFRASTRUCTURE."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg5_chunk(sigma, beta, rtol, atol, truth):
    sys.path.insert(0, ROOT)
    import oracle as O

    st, ops, ch, _ = O.parse_hamiltonian_dict(O.MANILA, [0])
    om = O.OracleModel(st, ops, ch, O.MANILA_CARRIERS, O.MANILA_DT, rotating_frame=st)
    samples = O.channel_samples(ch, [O.PulseSpec("drag", "d0", 256)], [dict(amp=0.12, angle=0.0, sigma=sigma, beta=beta)])
    span = (0.0, 256 * O.MANILA_DT)
    y0 = np.array([1, 0, 0])
    if truth:
        return O.dop853_solve(om, samples, y0, span)[1][:, -1], np.zeros((len(sigma), 2), dtype=np.int64)
    _, y, steps = O.dopri5_jax_solve(om, samples, y0, span, rtol=rtol, atol=atol, hmax=O.MANILA_DT)
    return y[:, -1], steps


def cfg5_oracle(sigma, beta, tol=(1e-8, 1e-6), truth=False, workers=None):
    """Final states [B,3] (and step counts [B,2]) of the cfg 5 problem for arrays sigma, beta; chunks run in parallel."""
    workers = workers or min(16, os.cpu_count() or 1)
    chunks = [c for c in np.array_split(np.arange(len(sigma)), max(1, min(len(sigma), workers))) if len(c)]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        for i, c in enumerate(chunks):
            fin, fout = os.path.join(tmp, f"in{i}.npz"), os.path.join(tmp, f"out{i}.npz")
            np.savez(fin, sigma=sigma[c], beta=beta[c], rtol=tol[0], atol=tol[1], truth=truth)
            procs.append((subprocess.Popen([sys.executable, os.path.abspath(__file__), fin, fout], env=env), fout))
        ys, steps = [], []
        for pr, fout in procs:
            if pr.wait() != 0:
                raise RuntimeError("oracle worker failed")
            r = np.load(fout)
            ys.append(r["y"])
            steps.append(r["steps"])
    return np.concatenate(ys), (None if truth else np.concatenate(steps))


if __name__ == "__main__":
    a = np.load(sys.argv[1])
    y, st = _cfg5_chunk(a["sigma"], a["beta"], float(a["rtol"]), float(a["atol"]), bool(a["truth"]))
    np.savez(sys.argv[2], y=y, steps=st)
