"""REVS pipeline object with the reference's interface (reference: revs_fixture.py).

    fx = REVS(**config["run_parameters"]["input_filepath"])
    tariff, homes, dist, saved = fx.read_inputs(**input_parameters)
    fx.get_distributed_optimal(tariff, homes, dist, save=True, **optimizer_parameters)

Method names, keyword names and defaults follow revs_fixture.py:60-330, including its
quirks that matter for parity: the constructor reads the community from the key
``comunityID`` (sic, revs_fixture.py:64) and the distributed run reads ``vlow``/``vhigh``
(revs_fixture.py:258-259) while the YAML provides ``vmin``/``vmax``, so the distributed
optimiser always runs with vlow=0.95, vhigh=1.05.  Optimisation runs on the GPU through
lpsolver.py of this package.
"""
import os

import numpy as np

from .extract import (GetCommunity, GetDistNet, GetHomeLoad, GetTariff, combine_result,
                      get_homes_ev_param)
from .lpsolver import (compute_flows, compute_voltage, solve_ADMM, solve_central,
                       solve_residences)


class REVS:
    def __init__(self, **kwargs):
        self.netID = kwargs.get("networkID", 121144)
        self.regID = kwargs.get("regionID", 121)
        self.com = kwargs.get("comunityID", 2)
        self.tariffID = kwargs.get("tariffID", "DVP")
        self.optim = kwargs.get("optimizer_mode", "individual")
        self.data_path = kwargs.get("data_path")
        out_path = kwargs.get("out_path")
        grb_path = kwargs.get("grb_path")
        self.out_dir = f"{out_path}/{self.netID}-com{self.com}/{self.optim}"
        self.grb_dir = f"{grb_path}/{self.netID}-com{self.com}/{self.optim}"   # kept for API parity; unused
        self.fig_dir = kwargs.get("fig_path")
        self.last_stats = None

    # ------------------------------------------------------------ inputs
    def read_tariff(self, tariffID=None, shift=6):
        return GetTariff(self.data_path, tariffID or "DVP", shift)

    def read_homes(self, regionID=None, shift=6):
        return GetHomeLoad(self.data_path, regionID or self.regID, shift=shift)

    def read_network(self, networkID=None):
        return GetDistNet(self.data_path, networkID or self.netID)

    def read_community(self, networkID=None, com_index=2):
        return GetCommunity(f"{self.data_path}/{networkID or self.netID}-com.txt", com_index)

    def read_inputs(self, regionID=None, networkID=None, tariffID=None, ev_homes=None, **kwargs):
        adoption = kwargs.get("adoption", 90)
        rating = kwargs.get("rating", 4800)
        capacity = kwargs.get("capacity", 20)
        initial = kwargs.get("initial_soc", 0.2)
        start = kwargs.get("start_time", 11)
        end = kwargs.get("end_time", 23)
        shift = kwargs.get("shift_time", 6)
        seed = kwargs.get("seed", 1234)

        tariff = self.read_tariff(tariffID=tariffID, shift=shift)
        all_homes = self.read_homes(regionID=regionID, shift=shift)
        dist = self.read_network(networkID=networkID)
        com = self.read_community(networkID=networkID, com_index=self.com)
        if ev_homes is None or len(ev_homes) == 0:
            np.random.seed(int(seed))                       # same draw as revs_fixture.py:174-177
            ev_homes = np.random.choice(com, int(adoption * 1e-2 * len(com)), replace=False)
        homes = get_homes_ev_param(all_homes, dist, ev_homes, rating * 1e-3, capacity,
                                   initial, start, end)
        return tariff, homes, dist, dict(ev_homes=ev_homes, community=com)

    # ------------------------------------------------------------ optimisation
    def _save(self, Pres, Pev, soc, kwargs, diff=None):
        os.makedirs(self.out_dir, exist_ok=True)
        name = f"adopt{kwargs.get('adoption', 90)}-rating{kwargs.get('rating', 4800)}-seed{kwargs.get('seed', None)}.txt"
        with open(f"{self.out_dir}/{name}", "w") as f:
            f.write(combine_result(Pres, Pev, soc, kwargs.get("ev_homes", None), diff))

    def get_individual_optimal(self, tariff, homes, save=False, **kwargs):
        Pev, soc, Pres = solve_residences(tariff, homes)
        if save:
            self._save(Pres, Pev, soc, kwargs)
        return Pres, Pev, soc

    def get_centralized_optimal(self, tariff, homes, dist, save=False, **kwargs):
        Pev, soc, Pres = solve_central(tariff, homes, dist, self.grb_dir, kwargs.get("v0", 1.03),
                                       kwargs.get("vmin", 0.90), kwargs.get("vmax", 1.05))
        if save:
            self._save(Pres, Pev, soc, kwargs)
        return Pres, Pev, soc

    def get_distributed_optimal(self, tariff, homes, dist, save=False, **kwargs):
        diff, Pres, Pev, soc, self.last_stats = solve_ADMM(
            homes, dist, tariff, self.grb_dir,
            kappa=kwargs.get("kappa", 5.0), iter_max=kwargs.get("max_iterations", 15),
            vset=kwargs.get("v0", 1.03), vlow=kwargs.get("vlow", 0.95), vhigh=kwargs.get("vhigh", 1.05),
            return_stats=True)
        if save:
            self._save(Pres, Pev, soc, kwargs, diff)
        return Pres, Pev, soc

    # ------------------------------------------------------------ reliability check
    def reliability(self, demand, dist, vset=1.0):
        """Numbers behind plot_result (revs_fixture.py:282-330): per-line loading and
        per-node voltage of a schedule, computed on the GPU."""
        return compute_flows(dist, demand), compute_voltage(dist, demand, vset=vset)

    def plot_result(self, demand, dist, **kwargs):
        raise NotImplementedError(
            "plotting (matplotlib/seaborn/geopandas) is outside this package; use "
            "REVS.reliability() for the flows and voltages the reference's box plots show")
