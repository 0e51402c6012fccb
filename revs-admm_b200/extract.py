"""Input/output formats of the reference, unchanged (reference: extract.py).

Same function names, arguments and return values as the reference module so that a
pipeline written against it keeps working:

  GetTariff(path, region, shift)            <region>-tariff.txt            extract.py:16-24
  GetHomeLoad(path, region_list, shift)     <region>-home-load.csv         extract.py:26-46
  GetDistNet(path, code)                    <code>-dist-net.gpickle        extract.py:48-80
  GetCommunity(filename, com_index)         <network>-com.txt              extract.py:82-89
  get_homes_ev_param(...)                   per-home EV dictionaries       extract.py:91-132
  combine_result(...)                       result text file               extract.py:134-174

GetDistNet differs in mechanism only: networkx >= 3 no longer ships read_gpickle, and the
pickles embed shapely geometries that are irrelevant to the optimisation, so the file is
read with a pickle.Unpickler that substitutes a placeholder for classes of packages that
are not installed.
"""
import os
import pickle

import numpy as np


def GetTariff(path, region, shift):
    fname = f"{path}/{region}-tariff.txt"
    if not os.path.exists(fname):
        raise ValueError(f"{fname} doesn't exist!")
    with open(fname) as f:
        values = [float(tok) for tok in f.readline().split()]
    return np.roll(values, -shift).tolist()


def GetHomeLoad(path, region_list, shift):
    """{hid: [24 hourly loads in kW, day rolled to start `shift` hours later]}."""
    import pandas as pd
    if not isinstance(region_list, (list, tuple)):
        region_list = [region_list]
    homes = {}
    hour_cols = [f"hour{i + 1}" for i in range(24)]
    for reg in region_list:
        fname = f"{path}/{reg}-home-load.csv"
        if not os.path.exists(fname):
            raise ValueError(f"{fname} doesn't exist!")
        df = pd.read_csv(fname)
        kw = 1e-3 * df[hour_cols].to_numpy(dtype=float)
        kw = np.roll(kw, -shift, axis=1)
        for hid, row in zip(df["hid"].to_numpy(), kw):
            homes[int(hid)] = row.tolist()
    return homes


class _Placeholder:
    """Stands in for objects of packages that are not installed (shapely geometries)."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        pass


class _TolerantUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return _Placeholder


def _read_gpickle(fname):
    if not os.path.exists(fname):
        raise ValueError(f"{fname} doesn't exist!")
    with open(fname, "rb") as f:
        return _TolerantUnpickler(f).load()


def GetDistNet(path, code):
    """networkx graph of the synthetic distribution network(s) `code` (node attr label in
    {'H','T','R','S'}, edge attrs r, x, type, label ...)."""
    if isinstance(code, list):
        import networkx as nx
        graph = nx.Graph()
        for c in code:
            graph = nx.compose(graph, _read_gpickle(f"{path}/{c}-dist-net.gpickle"))
        return graph
    return _read_gpickle(f"{path}/{code}-dist-net.gpickle")


def GetCommunity(filename, com_index):
    if not os.path.exists(f"{filename}"):
        raise ValueError(f"{filename} doesn't exist!")
    with open(filename) as f:
        rows = f.readlines()
    return [int(tok) for tok in rows[int(com_index) - 1].split()]


def _per_home(value, ev_homes):
    return value if isinstance(value, dict) else {h: value for h in ev_homes}


def get_homes_ev_param(homes, dist, ev_homes, rating, capacity, initial, start, end):
    """{h: {"LOAD": [...], "EV": {} | {rating, capacity, initial, start, end}}} for every
    residence of the network; scalars are broadcast to all adopters."""
    rating, capacity, initial = (_per_home(v, ev_homes) for v in (rating, capacity, initial))
    start, end = _per_home(start, ev_homes), _per_home(end, ev_homes)
    adopters = set(int(h) for h in ev_homes)
    out = {}
    for h in (n for n in dist if dist.nodes[n]["label"] == "H"):
        entry = {"LOAD": list(homes[h]), "EV": {}}
        if h in adopters:
            entry["EV"] = {"rating": rating[h], "capacity": float(capacity[h]),
                           "initial": initial[h], "start": start[h], "end": end[h]}
        out[h] = entry
    return out


_BAR = "#############################################"


def _section(title, table, keys):
    body = "\n".join(f"{h}:\t" + " ".join(str(float(x)) for x in table[h]) for h in keys)
    return f"\n{_BAR}\n{title}\n{_BAR}\n{body}"


def combine_result(P_res, P_ev, SOC, ev_homes, diff=None):
    """Text layout of the reference's result files (four '#'-delimited sections)."""
    data = _section("Residence Usage Profile", P_res, list(P_res))
    data += _section("EV Charger Usage Profile", P_ev, ev_homes)
    data += _section("EV Charger State of Charge Profile", SOC, ev_homes)
    if diff:
        conv = {h: [diff[k + 1][h] for k in range(len(diff))] for h in ev_homes}
        data += _section("EV Convergence over Iterations", conv, ev_homes)
    return data


def read_result(path):
    """Inverse of combine_result: {section title: {home id: np.ndarray}}."""
    with open(path) as f:
        rows = f.read().split("\n")
    out, cur, i = {}, None, 0
    while i < len(rows):
        if rows[i].startswith("####"):
            cur = rows[i + 1]
            out[cur] = {}
            i += 3
            continue
        if rows[i].strip() and cur is not None:
            hid, vals = rows[i].split(":\t")
            out[cur][int(hid)] = np.array([float(x) for x in vals.split(" ")])
        i += 1
    return out
