"""BASELINE.json configs[1]: MouseOrganogenesis (Stereo-seq MOSTA, E9.5-E16.5) *shaped* run of the analyze-side
hot path on one B200: per-timepoint domain k-means (K7) on the latents, then one unbalanced OT coupling per
adjacent pair of timepoints with wot's conventions (exact median, 3 growth iterations of the 6-stage duality-gap
solver to 1e-8, transition table over the domains) — `spadot_b200.analyze.ot_analysis` end to end.

The dataset itself is not shipped with the reference (only examples/MouseOrganogenesis_output/*.csv gene lists),
so the latents are synthetic (bench.synth mixture, d = z_dim = 20, utils/config.yaml) at MOSTA's published bin-50
spot counts per stage.  The reference's path for this config is wot on dense N x M fp64 matrices: at the last pair
(113k x 122k) K, _K, R, C alone are 4 x 110 GB.  One JSON line per pair + a summary line.

    python tools/mosta_bench.py [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from spadot_b200 import analyze, kmeans  # noqa: E402

STAGES = ["E9.5", "E10.5", "E11.5", "E12.5", "E13.5", "E14.5", "E15.5", "E16.5"]
SPOTS = [5913, 18408, 30124, 51365, 77369, 102519, 113350, 121767]      # MOSTA bin-50 spots per embryo section


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="multiply every spot count (CPU-sized smoke: 0.01)")
    ap.add_argument("--d", type=int, default=20)
    ap.add_argument("--k", type=int, default=10)
    a = ap.parse_args()
    # under torchrun: the rows of every source timepoint are partitioned over the ranks (k-means is replicated, it is 1 % of the time)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist = None
    if world > 1:
        import torch.distributed as td
        from spadot_b200 import sinkhorn
        td.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
        dist = sinkhorn.Dist()
    say = print if rank == 0 else (lambda *args, **kw: None)
    sizes = [max(64, int(s * a.scale)) for s in SPOTS]
    lat = []
    for t, n in enumerate(sizes):
        x, _ = bench.synth(n, 8, a.d, seed=1993)            # same mixture for every stage, fresh draws, drifting
        rng = np.random.default_rng(t)
        lat.append(x[rng.permutation(n)] + 0.05 * t)
    t_all = time.perf_counter()
    labels, t_km = [None] * len(lat), [0.0] * len(lat)
    for t, x in enumerate(lat):
        if t % world != rank:                                # the timepoints' k-means runs are independent: dealt over the ranks
            continue
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        km = kmeans.KMeans(n_clusters=a.k, n_init=10, random_state=1993).fit(x)
        torch.cuda.synchronize()
        t_km[t] = time.perf_counter() - t0
        labels[t] = np.asarray(km.labels_)
    if dist is not None:
        for t in range(len(lat)):
            buf = torch.from_numpy(labels[t].astype(np.int64)).cuda() if labels[t] is not None else torch.empty(sizes[t], dtype=torch.int64, device="cuda")
            td.broadcast(buf, src=t % world)
            labels[t] = buf.cpu().numpy()
            tk = torch.tensor([t_km[t]], dtype=torch.float64, device="cuda")
            td.broadcast(tk, src=t % world)
            t_km[t] = float(tk.item())
    rows = []
    if dist is not None:                                     # first-use costs of the communicators stay out of the pair times
        analyze.transport_between(lat[0][rank::world][:256], lat[1], dist=dist)
    for t in range(len(lat) - 1):
        r0, r1 = (sizes[t] * rank) // world, (sizes[t] * (rank + 1)) // world
        torch.cuda.synchronize()
        if dist is not None:
            td.barrier()
        t0 = time.perf_counter()
        cp, growth = analyze.transport_between(lat[t][r0:r1], lat[t + 1], dist=dist)
        torch.cuda.synchronize()
        if dist is not None:
            td.barrier()
        t_solve = time.perf_counter() - t0
        t0 = time.perf_counter()
        table = cp.transition_table(labels[t][r0:r1], labels[t + 1], a.k, a.k).cpu().numpy()
        torch.cuda.synchronize()
        t_table = time.perf_counter() - t0
        prob = analyze.transition_probabilities(table)
        mass = float(growth[-1].sum())
        if dist is not None:
            mass = float(dist.sum_(torch.tensor([mass], dtype=torch.float64, device="cuda")).item())
        row = dict(pair=f"{STAGES[t]}->{STAGES[t + 1]}", n=sizes[t], m=sizes[t + 1], n_gpus=world, seconds=t_solve, table_s=t_table,
                   iters_last_growth=cp.info["total_iters"], gap=cp.info["gap"], tc=bool(cp.ops.use_tc),
                   plan_mass=mass, table_mass=float(table.sum()),
                   argmax=[int(v) for v in prob.argmax(axis=1)],
                   dense_fp64_gb_reference=4 * 8.0 * sizes[t] * sizes[t + 1] / 1e9)
        rows.append(row)
        say(json.dumps(row), flush=True)
    total = time.perf_counter() - t_all
    say(json.dumps(dict(n_gpus=world, workload="MouseOrganogenesis-shaped analyze (8 stages, MOSTA bin-50 spot counts x %.3g, d=%d, k=%d)"
                                   % (a.scale, a.d, a.k), kmeans_s=t_km, ot_pairs_s=[r["seconds"] for r in rows],
                          total_s=total, peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)))
    if dist is not None:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
