#!/usr/bin/env python
"""Build the committed golden fixtures from the read-only reference checkout.

Run ONCE in the build container (needs /root/reference); the outputs below are
committed so that nothing at test/bench time reads /root/reference.

  tests/golden/input/121144-dist-net.gpickle   input data, byte copy
  tests/golden/input/121144-com.txt            input data, byte copy
  tests/golden/input/DVP-tariff.txt            input data, byte copy
  tests/golden/input/121-home-load.csv         RECONSTRUCTED (the reference lists it
        in .MISSING_LARGE_BLOBS): hourly load = "Residence Usage Profile" minus
        "EV Charger Usage Profile" of the reference's own result files, un-rolled
        by shift_time=6 and converted kW -> W, i.e. the inverse of
        extract.py:26-46 (GetHomeLoad).
  tests/golden/ref_out_121144_com2.npz         the reference's result files
        out/121144-com2/{individual,centralized,distributed}/adopt90-rating4800-seed1234.txt,
        individual/adopt90-rating3600-seed1234.txt and individual/adopt70-rating4800-seed1234.txt
        parsed into arrays (home ids, P_res, P_ev, SOC, per-iteration diff).
"""
import os
import shutil
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_result(path):
    """Parse the text layout written by extract.py:134-174 (combine_result)."""
    with open(path) as f:
        lines = f.read().split("\n")
    secs, cur, i = {}, None, 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("####"):
            cur = lines[i + 1]
            secs[cur] = {}
            i += 3
            continue
        if ln.strip() and cur is not None:
            h, v = ln.split(":\t")
            secs[cur][int(h)] = np.array([float(x) for x in v.split(" ")])
        i += 1
    return secs


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    inp = os.path.join(HERE, "input")
    os.makedirs(inp, exist_ok=True)
    for name in ("121144-dist-net.gpickle", "121144-com.txt", "DVP-tariff.txt"):
        shutil.copyfile(os.path.join(REF, "input", name), os.path.join(inp, name))

    out = {}
    base = os.path.join(REF, "out", "121144-com2")
    for mode in ("individual", "centralized", "distributed"):
        s = parse_result(os.path.join(base, mode, "adopt90-rating4800-seed1234.txt"))
        res = np.array(list(s["Residence Usage Profile"]), dtype=np.int64)
        ev = np.array(list(s["EV Charger Usage Profile"]), dtype=np.int64)
        out[f"{mode}_res_ids"] = res
        out[f"{mode}_ev_ids"] = ev
        out[f"{mode}_P_res"] = np.array([s["Residence Usage Profile"][h] for h in res])
        out[f"{mode}_P_ev"] = np.array([s["EV Charger Usage Profile"][h] for h in ev])
        out[f"{mode}_SOC"] = np.array([s["EV Charger State of Charge Profile"][h] for h in ev])
        if "EV Convergence over Iterations" in s:
            out[f"{mode}_diff"] = np.array([s["EV Convergence over Iterations"][h] for h in ev])
    # an extra individual run shipped by the reference (different rating)
    s = parse_result(os.path.join(base, "individual", "adopt90-rating3600-seed1234.txt"))
    ev = np.array(list(s["EV Charger Usage Profile"]), dtype=np.int64)
    out["individual3600_ev_ids"] = ev
    out["individual3600_P_ev"] = np.array([s["EV Charger Usage Profile"][h] for h in ev])
    out["individual3600_SOC"] = np.array([s["EV Charger State of Charge Profile"][h] for h in ev])
    # ... and one with 70 % adoption: pins the seeded EV-home draw at a second adoption level
    s = parse_result(os.path.join(base, "individual", "adopt70-rating4800-seed1234.txt"))
    ev = np.array(list(s["EV Charger Usage Profile"]), dtype=np.int64)
    out["individual70_ev_ids"] = ev
    out["individual70_P_ev"] = np.array([s["EV Charger Usage Profile"][h] for h in ev])
    out["individual70_SOC"] = np.array([s["EV Charger State of Charge Profile"][h] for h in ev])
    np.savez_compressed(os.path.join(HERE, "ref_out_121144_com2.npz"), **out)

    # reconstruct the home-load CSV (hid,hour1..hour24 in W, un-shifted)
    res = out["distributed_res_ids"]
    P = out["distributed_P_res"].copy()
    evpos = {int(h): i for i, h in enumerate(out["distributed_ev_ids"])}
    for i, h in enumerate(res):
        if int(h) in evpos:
            P[i] -= out["distributed_P_ev"][evpos[int(h)]]
    raw = np.roll(P, 6, axis=1) * 1e3  # undo np.roll(.., -shift) and the 1e-3
    with open(os.path.join(inp, "121-home-load.csv"), "w") as f:
        f.write("hid," + ",".join(f"hour{i+1}" for i in range(24)) + "\n")
        for h, row in zip(res, raw):
            f.write(str(int(h)) + "," + ",".join(repr(float(x)) for x in row) + "\n")
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
