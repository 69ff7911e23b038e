"""Host-side utilities under the reference's names: the training-configuration file
(evstore_utils.py:31-53) and the per-request latency CDF the inference driver writes
(``calculate_and_write_cdf``, dlrm_s_pytorch_C1_C2_C3.py:291-319) -- the file the paper's CDF plots
(experiments.md:694-802) are drawn from."""
from __future__ import annotations

import os

from .storage_manager import read_training_config, store_training_config  # noqa: F401  (evstore_utils.py:31-53)

TRAINING_CONFIG_FILE = "training_config.txt"       # evstore_utils.py:10


def calculate_and_write_cdf(cdf_output_dir, cache_algo, arr_time_start, n_points: int = 1000):
    """arr_time_start: the start time (seconds) of every request plus one final entry for the completion of the
    last one.  Like the reference: latency i = start[i+1] - start[i] for i < len - 2 (its loop drops the last
    interval, :297), sorted, thinned to every int(n / n_points)-th point, written as ``y,latency_ms`` with
    y = rank / points to ``<cdf_output_dir>/<cache_algo>-cdf.csv``.  Returns the path.
    Departures: fewer than n_points latencies are written unthinned (the reference divides by a zero step), and
    the plot script (``./script/plot_cdf.py``) is not invoked."""
    os.makedirs(cdf_output_dir, exist_ok=True)
    lat = sorted(float(arr_time_start[i + 1]) - float(arr_time_start[i]) for i in range(0, len(arr_time_start) - 2))
    step = max(1, int(len(lat) / n_points))
    lat = lat[0::step]
    n = len(lat)
    output = os.path.join(cdf_output_dir, str(cache_algo) + "-cdf.csv")
    with open(output, "w") as f:
        f.write("y,latency_ms\n")
        for i, v in enumerate(lat):
            f.write(f"{(i + 1) / n!r},{v * 1000!r}\n")
    print("CDF Latency data points is written to: " + output)
    return output
