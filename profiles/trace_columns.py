"""Per-column timeline of one working-set round (REVS_DEBUG_TRACE): where a QP column spends its time.

    REVS_DEBUG_TRACE=4,0 REVS_DEBUG_TRACE_FILE=gpurun_out/trace.bin python profiles/trace_columns.py [workload]

Record layout (utility_qp_warp.cu:trace_out / utility_qp.cu): start ns, end ns, smid, (class<<40 | m<<20 | its),
4-5 phase cycle counts, passes / pdas, exit kind / evals, stored rows / -, cycles."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import revs_admm_b200 as R  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic-refshape-125k-homes-per-gpu-x96"
    trees, hm, cost, sizes, T = bench.make_rank_problem(wl, 0)
    with R.Solver(sizes, T) as s:
        s.set_feeder_trees(trees)
        s.set_homes(**hm)
        s.set_tariff(cost)
        s.solve_admm(**bench.ADMM)
        st = s.stats()
    path = os.environ.get("REVS_DEBUG_TRACE_FILE", "revs_trace.bin")
    tr = np.fromfile(path, dtype=np.int64).reshape(-1, 12)
    ran = tr[:, 1] > 0
    t = tr[ran]
    dur = (t[:, 1] - t[:, 0]) * 1e-3
    span = (t[:, 1].max() - t[:, 0].min()) * 1e-3
    n_of_col = np.repeat(np.asarray(sizes), T)[ran]
    warp = (t[:, 3] >> 40) == 0
    print(f"columns traced {ran.sum()} of {len(tr)}; round span {span:.0f} us; total_ms {st['total_ms']:.2f}")
    w = t[warp]
    m = (w[:, 3] >> 20) & 0xFFFFF
    its = w[:, 3] & 0xFFFFF
    kind = w[:, 9] & 15
    print(f"warp-kernel columns {len(w)}: exit kinds finished/deferred/handed = {(kind == 0).sum()}/{(kind == 1).sum()}/{(kind == 2).sum()}")
    print(" m  columns  mean_us  p90_us  its  passes | cycles: load admit solve verify total | n mean")
    for mm in sorted(set(m.tolist())):
        sel = m == mm
        d = dur[warp][sel]
        print(f"{mm:2d} {sel.sum():8d} {d.mean():8.1f} {np.percentile(d, 90):7.1f} {its[sel].mean():5.1f} {w[sel, 8].mean():6.2f} | "
              f"{w[sel, 4].mean():7.0f} {w[sel, 5].mean():7.0f} {w[sel, 6].mean():7.0f} {w[sel, 7].mean():7.0f} {w[sel, 11].mean():8.0f} | {n_of_col[warp][sel].mean():.0f}")
    handed = kind == 2
    if handed.any():
        reasons = (w[handed, 9] >> 4)
        print("hand-over reasons (1 full set, 2 passes, 3 pdas, 4 line search, 5 other):", np.bincount(reasons.astype(int)).tolist())
    c = t[~warp]
    if len(c):
        cls = c[:, 3] >> 40
        mc = (c[:, 3] >> 20) & 0xFFFFF
        dc = dur[~warp]
        for cl in sorted(set(cls.tolist())):
            sel = cls == cl
            print(f"CTA class {cl}: {sel.sum()} columns, mean {dc[sel].mean():.1f} us, p90 {np.percentile(dc[sel], 90):.1f}, m mean {mc[sel].mean():.1f}, "
                  f"phase Mcycles grad {c[sel, 4].sum() * 1e-6:.1f} hess {c[sel, 5].sum() * 1e-6:.1f} pdas {c[sel, 6].sum() * 1e-6:.1f} search {c[sel, 7].sum() * 1e-6:.1f} eval {c[sel, 8].sum() * 1e-6:.1f}")
    sm = t[:, 2]
    busy = np.zeros(int(sm.max()) + 1)
    np.add.at(busy, sm, dur)
    print(f"per-SM busy column-time: mean {busy.mean():.0f} us, max {busy.max():.0f} us (round span {span:.0f} us)")


if __name__ == "__main__":
    main()
