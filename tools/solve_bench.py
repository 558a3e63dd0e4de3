"""Wall-clock cost of one keyframe's window solve on the benched shape (at128, 10 references):
fflins_window_solve (one launch, the LM loop on the device) against the host-driven sequence
(17 x fflins_window_eval + 1 x fflins_window_chi2 round trips, with and without a host LM between them).
Usage: python tools/solve_bench.py [reps]  -> one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ff-lins_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def main():
    from fflins import api, synth
    from oracle import pyoracle   # only to stage the window exactly as the parity tests do
    import lm_ref
    from util import build_window, window_params
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    seq = synth.Sequence("at128")
    c = api.Context(api.default_config(point_filter_num=seq.filter_num, window_size=10))
    keys = build_window(c, pyoracle, seq, n_keys=11, frames_per_key=5)
    rng = np.random.default_rng(7)
    st = c.associate()
    pose_refs, pose_cur, ext, td = window_params(seq, keys, rng)
    refs = [k["id"] for k in keys[:-1]]
    n = len(refs)
    x = np.concatenate([pose_refs.reshape(-1), pose_cur, ext, [td]])
    D = 6 * n + 13
    J0 = rng.normal(size=(D + 5, D)) * 20.0
    prior = (J0.T @ J0, np.zeros(D), x.copy(), 0.0)
    zero = [np.zeros(len(c.features(r)["outlier"]), np.uint8) for r in refs]

    def clear():
        for r, z in zip(refs, zero):
            c.set_outlier(r, z)
    out = {"n_refs": n, "queries": int(st.n_queries), "D": D, "reps": reps}
    for name, kw in (("device_full_prior", dict(flags=1)), ("device_registration_only", dict(flags=1 | 2 | 8 | 16))):
        opt = c.solve_options(function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0, **kw)
        pr = prior if name == "device_full_prior" else None
        ts, evals = [], 0
        for i in range(reps + 5):
            clear()
            t0 = time.perf_counter()
            _, _, _, _, rep = c.window_solve(refs, pose_refs, pose_cur, ext, td, options=opt, prior=pr)
            t1 = time.perf_counter()
            if i >= 5:
                ts.append(t1 - t0)
                evals = rep.evaluations[0] + rep.evaluations[1]
        out[name] = {"us_per_solve_median": round(1e6 * float(np.median(ts)), 2), "us_min": round(1e6 * min(ts), 2),
                     "evaluations": int(evals), "iterations": [int(rep.iterations[0]), int(rep.iterations[1])],
                     "solver_cta_phase_cycles": dict(zip(("wait_eval", "assemble", "damped_copy", "back_substitution", "model", "plus", "elimination"),
                                                         [int(v) for v in rep.phase_cycles[:7]]))}
    # host-driven: the same number of evaluations as round trips, no solver between them (what bench.py's step does)
    prep, res = c.prepare_window(refs, pose_refs, pose_cur, ext)
    ts = []
    for i in range(reps + 5):
        clear()
        t0 = time.perf_counter()
        for _ in range(6):
            c.window_eval_prepared(prep, td, 1, res)
        c.window_chi2_prepared(prep, td)
        for _ in range(11):
            c.window_eval_prepared(prep, td, 1, res)
        t1 = time.perf_counter()
        if i >= 5:
            ts.append(t1 - t0)
    out["host_round_trips_only"] = {"us_per_keyframe_median": round(1e6 * float(np.median(ts)), 2), "us_min": round(1e6 * min(ts), 2),
                                    "evaluations": 17}
    # host-driven with a host LM (numpy; the interpreter's share is large, reported for completeness)
    clear()
    t0 = time.perf_counter()

    def ev(xx):
        o = c.window_eval(refs, xx[:7 * n].reshape(n, 7), xx[7 * n:7 * n + 7], xx[7 * n + 7:7 * n + 14], float(xx[7 * n + 14]), 1)
        return o["H"], o["b"], o["cost"]

    def chi2(xx, thr):
        return c.window_chi2(refs, xx[:7 * n].reshape(n, 7), xx[7 * n:7 * n + 7], xx[7 * n + 7:7 * n + 14], float(xx[7 * n + 14]), thr)
    lm_ref.window_solve(ev, chi2, n, x, 1, prior=prior, ftol=0.0, gtol=0.0, ptol=0.0)
    out["host_numpy_lm_us"] = round(1e6 * (time.perf_counter() - t0), 1)
    clear()
    print(json.dumps(out))
    c.close()


if __name__ == "__main__":
    main()
