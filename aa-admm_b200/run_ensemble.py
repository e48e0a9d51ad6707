"""Ensemble runner (BASELINE configs[4] = cfg 5): N independent beam scenes with the material
sweep of `ensemble.scene_material`, scene s on GPU s mod G, no data-path collective; one gather of the
per-scene result records at the end.

    python aa-admm_b200/run_ensemble.py                      # 64 scenes 88x22x22, one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        aa-admm_b200/run_ensemble.py                         # 8 scenes per GPU

Prints one JSON line on rank 0: scenes/s (setup included), aggregate iterations/s of the ADMM loops
(device time, max over ranks of the per-rank sums) and the per-scene table.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aa_admm_b200 as A  # noqa: E402
from aa_admm_b200 import ensemble as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=64)
    ap.add_argument("--dims", type=int, nargs=3, default=[88, 22, 22])
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--anderson-m", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    device = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=device)
        dist.barrier()
    A.set_device(local)
    dt = 1.0 / 30.0
    t_all = time.perf_counter()
    recs = []
    for s in E.scenes_of_rank(args.scenes, rank, world):
        youngs, poisson = E.scene_material(s)
        t0 = time.perf_counter()
        solver, scene = A.make_beam_solver(*args.dims, iters=args.iters, anderson_m=args.anderson_m, youngs=youngs, poisson=poisson)
        pidx = scene.arrays()[3]
        iters = rejects = 0
        loop_ms = 0.0
        last = None
        for _ in range(args.frames):
            solver.set_pins(pidx, scene.stretch(dt))
            last = solver.step()  # rows: primal residual, combined residual, is_reject
            info = solver.info()
            iters += last.shape[0]
            rejects += int(last[:, 2].sum())
            loop_ms += info["loop_ms"]
        recs.append(E.make_record(s, iters, rejects, last[-1, 0], last[-1, 1], loop_ms,
                                  1e3 * (time.perf_counter() - t0), rank))
        del solver, scene
    table = E.gather_records(np.array(recs), dist, device)
    wall = time.perf_counter() - t_all
    if dist is not None:
        import torch
        w = torch.tensor([wall], dtype=torch.float64, device=device)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        wall = float(w.item())
    if rank == 0:
        per_rank_loop = [table[table[:, 7] == r][:, 5].sum() for r in range(world)]
        out = {"workload": "cfg5: %d scenes, beam %dx%dx%d, material sweep, %d frame(s) x %d iterations, m=%d" %
               (args.scenes, *args.dims, args.frames, args.iters, args.anderson_m),
               "n_gpus": world, "scenes_per_s": args.scenes / wall, "wall_s": wall,
               "aggregate_iterations_per_s": float(table[:, 1].sum() / (max(per_rank_loop) * 1e-3)),
               "record_fields": list(E.RECORD_FIELDS), "records": table.tolist()}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
