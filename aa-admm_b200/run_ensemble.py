"""Ensemble runner (BASELINE configs[4] = cfg 5): N independent beam scenes with the material sweep of
`ensemble.scene_material`, scene s on GPU s mod G, no data-path collective; one gather of the per-scene result
records at the end. Per GPU `--slots` scenes are resident (ensemble.SceneSlot: the mesh, its analysis and all device
buffers persist, every member only redoes the numeric setup) and are pipelined by one host thread each.

    python aa-admm_b200/run_ensemble.py                      # 64 scenes 88x22x22, one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        aa-admm_b200/run_ensemble.py                         # 8 scenes per GPU

Prints one JSON line on rank 0: scenes/s of the first pass (one-time analysis included) and of the following passes
(per-scene numeric setup included), aggregate iterations/s and the per-scene table. `bench.py --config cfg5` measures
the same workload under the bench contract.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=64)
    ap.add_argument("--dims", type=int, nargs=3, default=[88, 22, 22])
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--anderson-m", type=int, default=5)
    ap.add_argument("--slots", type=int, default=2)
    ap.add_argument("--passes", type=int, default=2, help="passes over the ensemble (the first one includes the analysis)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // (world * args.slots)))
    import aa_admm_b200 as A
    from aa_admm_b200 import ensemble as E
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
    mine = E.scenes_of_rank(args.scenes, rank, world)
    slots = [E.SceneSlot(A, args.dims, iters=args.iters, anderson_m=args.anderson_m, device=local) for _ in range(args.slots)]
    walls = []
    for p in range(args.passes):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        recs, setups = E.run_sweep(A, args.dims, mine, slots, args.frames, rank)
        wall = time.perf_counter() - t0
        if dist is not None:
            import torch
            w = torch.tensor([wall], dtype=torch.float64, device=device)
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
            wall = float(w.item())
        walls.append(wall)
    table = E.gather_records(recs, dist, device)
    if rank == 0:
        out = {"workload": "cfg5: %d scenes, beam %dx%dx%d, material sweep, %d frame(s) x %d iterations, m=%d, %d resident scenes per GPU" %
               (args.scenes, *args.dims, args.frames, args.iters, args.anderson_m, args.slots),
               "n_gpus": world, "pass_wall_s": walls, "scenes_per_s_first_pass": args.scenes / walls[0],
               "scenes_per_s": args.scenes / walls[-1],
               "aggregate_iterations_per_s": float(table[:, 1].sum() / walls[-1]),
               "setup_ms_last_pass_rank0": [[int(s), round(ms, 2), bool(inc)] for s, ms, inc in setups],
               "record_fields": list(E.RECORD_FIELDS), "records": table.tolist()}
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
