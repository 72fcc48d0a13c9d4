#!/usr/bin/env python
"""Time BASELINE.json configs 2-5 on one GPU (CUDA events, warm-up, best/mean of N) and write a JSON summary.
cfg 1 is the CPU reference sanity case; its GPU parity lives in tests/test_gpu_parity_sv.py.
Usage: python scripts/bench_configs.py [--out FILE] [--quick]"""
import argparse, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from casq_b200 import engine
from scripts import _problems as P
from scripts._problems import t1t2  # noqa: F401  (re-exported for the other scripts)

DT = P.DT


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--out", default=None); ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    res = {"gpu": torch.cuda.get_device_name(0), "fp64_fma_probe_tflops": engine.fp64_fma_probe()}
    rng = np.random.default_rng(20260101)
    # ---- cfg 2
    st, ops, ch = P.manila([0])
    nm = P.native(st, ops, ch)
    bb, aa = np.meshgrid(np.linspace(-5, 5, 250), np.linspace(0.02, 0.5, 400), indexing="ij")
    p = engine.pack_params([dict(amp=aa.ravel(), angle=0.0, sigma=64.0, beta=bb.ravel())], device="cuda")
    y0 = np.array([1, 0, 0])
    best, mean = timed(lambda: engine.solve_sv_rk4(nm, [("drag", 0, 0, 256)], p, y0, (0, 257 * DT), DT / 40), reps=5, warm=2)
    res["cfg2_rk4_d3"] = {"B": 100000, "steps": 10280, "ms_best": best, "traj_per_s": 1e5 / best * 1e3}
    # ---- cfg 5: 1e4 DRAG candidates, Dopri5 atol 1e-6 rtol 1e-8 hmax dt
    B5 = 10000
    p5 = engine.pack_params([dict(amp=0.12, angle=0.0, sigma=rng.uniform(16, 128, B5), beta=rng.uniform(-4, 4, B5))], device="cuda")
    out5 = {}
    def run5():
        out5["r"] = engine.solve_sv_dopri5(nm, [("drag", 0, 0, 256)], p5, y0, (0, 256 * DT), rtol=1e-8, atol=1e-6, hmax=DT, return_stats=True)
    best, mean = timed(run5, reps=3, warm=1)
    y5, ns5, st5 = out5["r"]
    states, probs, pops = engine.postprocess_sv(nm, np.array([0, 256 * DT]), y5)
    res["cfg5_dopri5_d3"] = {"B": B5, "ms_best": best, "traj_per_s": B5 / best * 1e3, "attempted_steps_total": int(ns5[:, 0].sum()),
                             "accepted_steps_total": int(ns5[:, 1].sum()), "failed": int((st5 != 0).sum()),
                             "best_infidelity_1_minus_P1fold": float((1 - pops[:, -1, 1]).min()),
                             "best_infidelity_true_1_minus_psi1sq": float((1 - probs[:, -1, 1]).min()),
                             "flop_per_attempted_step": 1152,
                             "algorithmic_TFLOPs": int(ns5[:, 0].sum()) * 1152 / (best * 1e-3) / 1e12}
    # ---- cfg 3: 2 transmons, GaussianSquare on u0, 64x64 sweep, expm-midpoint, RWA m=1 and no-RWA m=20
    st2, ops2, ch2 = P.chain(2)
    ww, aa3 = np.meshgrid(np.linspace(0, 768, 64), np.linspace(0.05, 0.8, 64), indexing="ij")
    p3 = engine.pack_params([dict(amp=aa3.ravel(), angle=0.0, sigma=64.0, width=ww.ravel())], device="cuda")
    for tag, rwa, sub in (("rwa_m1", 0.75 * P.CARRIERS["d0"], 1), ("norwa_m20", None, 20)):
        if args.quick and sub == 20:
            continue
        nm2 = P.native(st2, ops2, ch2, rwa=rwa)
        keep = {}
        def run3():
            keep["U"] = engine.propagators_expm(nm2, [("gaussian_square", ch2.index("u0"), 0, 1024)], p3, (0, 1024 * DT), DT / sub)
        best, mean = timed(run3, reps=2, warm=1)
        U = keep["U"]
        unit = float((U.conj().transpose(1, 2) @ U - torch.eye(9, device="cuda", dtype=U.dtype)).abs().max())
        res[f"cfg3_expm_d9_{tag}"] = {"B": 4096, "steps": 1024 * sub, "ms_best": best, "traj_per_s": 4096 / best * 1e3,
                                      "max_unitarity_defect": unit, "rwa_cutoff_hz": rwa,
                                      # F_expm(9, s=2) ~ 6.0e4 flop per step (Pade-13 count; the kernel executes orders 3-5)
                                      "algorithmic_TFLOPs": 4096 * 1024 * sub * 6.0e4 / (best * 1e-3) / 1e12}
    # ---- cfg 4: 5 transmons d=243, Lindblad T1/T2, RWA, max_dt = dt/4, S = 1028, B = 128 per GPU
    st5m, ops5, ch5 = P.manila(None)
    nm5 = P.native(st5m, ops5, ch5, frame="diag", rwa=0.75 * P.CARRIERS["d0"])
    nm5.set_dissipators(t1t2(5))
    B4 = 16 if args.quick else 128
    p4 = engine.pack_params([dict(amp=np.linspace(0.05, 0.5, B4), angle=0.0, sigma=64.0, beta=1.0)], device="cuda")
    rho0 = np.zeros(243, dtype=complex); rho0[0] = 1
    keep = {}
    def run4():
        keep["r"] = engine.solve_lindblad_rk4(nm5, [("drag", ch5.index("d0"), 0, 256)], p4, rho0, (0, 257 * DT), DT / 4, want_rho=False)
    best, mean = timed(run4, reps=1, warm=1)
    pops = keep["r"][1][:, -1]
    alg_bytes = 8 * 243 * 243 * 16 * 1028 * B4
    res["cfg4_lindblad_d243"] = {"B": B4, "steps": 1028, "ms_best": best, "traj_per_s": B4 / best * 1e3,
                                 "algorithmic_GBps": alg_bytes / (best * 1e-3) / 1e9, "hbm_peak_GBps_measured": 6554.6,
                                 "trace_min": float(pops.sum(-1).min()), "trace_max": float(pops.sum(-1).max()),
                                 "P_ground_first_last": [float(pops[0, 0]), float(pops[-1, 0])]}
    # ---- beyond d = 4: lane-group RK4 and warp-per-trajectory Dopri5 on 2 and 3 coupled transmons (not BASELINE configs)
    if not args.quick:
        Bg = 4096
        for nq in (2, 3):
            stg, opsg, chg = P.chain(nq)
            dg = stg.shape[0]
            nmg = P.native(stg, opsg, chg)
            pg = engine.pack_params([dict(amp=rng.uniform(0.05, 0.5, Bg), angle=0.0, sigma=64.0, beta=rng.uniform(-2, 2, Bg))], device="cuda")
            yg = np.zeros(dg, dtype=complex); yg[0] = 1
            sl = [("drag", chg.index("d0"), 0, 256)]
            best, _ = timed(lambda: engine.solve_sv_rk4(nmg, sl, pg, yg, (0, 257 * DT), DT / 40), reps=2, warm=1)
            res[f"extra_rk4_general_d{dg}"] = {"B": Bg, "steps": 10280, "ms_best": best, "traj_per_s": Bg / best * 1e3}
            outg = {}
            def rung():
                outg["r"] = engine.solve_sv_dopri5(nmg, sl, pg, yg, (0, 256 * DT), rtol=1e-8, atol=1e-6, hmax=DT, return_stats=True)
            best, _ = timed(rung, reps=2, warm=1)
            res[f"extra_dopri5_general_d{dg}"] = {"B": Bg, "ms_best": best, "traj_per_s": Bg / best * 1e3,
                                                  "attempted_steps_total": int(outg["r"][1][:, 0].sum()), "failed": int((outg["r"][2] != 0).sum())}
        # three-transmon unitaries (block-per-trajectory propagator kernel) and the full five-transmon state-vector solve
        st3, ops3, ch3 = P.chain(3)
        nm3 = P.native(st3, ops3, ch3, rwa=0.75 * P.CARRIERS["d0"])
        B27 = 1024
        p27 = engine.pack_params([dict(amp=rng.uniform(0.05, 0.8, B27), angle=0.0, sigma=64.0, width=rng.uniform(0, 768, B27))], device="cuda")
        best, _ = timed(lambda: engine.propagators_expm(nm3, [("gaussian_square", ch3.index("u0"), 0, 1024)], p27, (0, 1024 * DT), DT), reps=2, warm=1)
        res["extra_expm_block_d27"] = {"B": B27, "steps": 1024, "ms_best": best, "traj_per_s": B27 / best * 1e3}
        nm5s = P.native(st5m, ops5, ch5, frame="diag", rwa=0.75 * P.CARRIERS["d0"])
        B243 = 4096
        p243 = engine.pack_params([dict(amp=np.linspace(0.05, 0.5, B243), angle=0.0, sigma=64.0, beta=1.0)], device="cuda")
        y243 = np.zeros(243, dtype=complex); y243[0] = 1
        best, _ = timed(lambda: engine.solve_sv_rk4(nm5s, [("drag", ch5.index("d0"), 0, 256)], p243, y243, (0, 257 * DT), DT / 4), reps=2, warm=1)
        res["extra_rk4_general_d243"] = {"B": B243, "steps": 1028, "ms_best": best, "traj_per_s": B243 / best * 1e3}
    txt = json.dumps(res, indent=1)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
