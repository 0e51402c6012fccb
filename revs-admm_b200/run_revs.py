"""Command-line runner with the reference's calling convention (test-optimizer.py):

    python revs-admm_b200/run_revs.py revs-admm_b200/revs_config.yaml
"""
import os
import sys

import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from revs_admm_b200.revs_fixture import REVS  # noqa: E402


def main(config_file):
    with open(config_file) as f:
        config = yaml.safe_load(f)
    run = config["run_parameters"]
    file_params, inp, opt = run["input_filepath"], run["input_parameters"], run["optimizer_parameters"]
    fx = REVS(**file_params)
    tariff, homes, dist, saved = fx.read_inputs(**inp)
    opt.update(saved)
    opt.update(inp)
    if fx.optim == "individual":
        fx.get_individual_optimal(tariff, homes, save=True, **opt)
    elif fx.optim == "centralized":
        fx.get_centralized_optimal(tariff, homes, dist, save=True, **opt)
    elif fx.optim == "distributed":
        fx.get_distributed_optimal(tariff, homes, dist, save=True, **opt)
        print(fx.last_stats)
    print("results in", fx.out_dir)


if __name__ == "__main__":
    main(sys.argv[1])
