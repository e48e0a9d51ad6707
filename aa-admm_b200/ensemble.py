"""Ensemble sharding (BASELINE configs[4]): independent scenes, static scene -> rank map,
no per-iteration collective; one gather of fixed-size result records at the end."""
import numpy as np

RECORD_FIELDS = ("scene_id", "iters", "rejects", "final_prim", "final_comb", "loop_ms", "wall_ms", "rank")


def scene_material(s):
    """Material sweep of cfg 5: E_s = 10^(6 + 2*floor(s/8)/7) Pa, nu_s = 0.30 + 0.02*(s mod 8)."""
    return 10.0 ** (6.0 + 2.0 * (s // 8) / 7.0), 0.30 + 0.02 * (s % 8)


def scenes_of_rank(n_scenes, rank, world):
    """Scene s runs on GPU s mod G."""
    return [s for s in range(n_scenes) if s % world == rank]


def make_record(scene_id, iters, rejects, final_prim, final_comb, loop_ms, wall_ms, rank):
    return np.array([scene_id, iters, rejects, final_prim, final_comb, loop_ms, wall_ms, rank], dtype=np.float64)


def gather_records(records, dist=None, device=None):
    """records: (k, 8) array of this rank. Returns the (n_scenes, 8) table sorted by scene id on
    every rank. `dist` is torch.distributed (NCCL on GPUs, gloo on CPU) or None for one process."""
    records = np.asarray(records, dtype=np.float64).reshape(-1, len(RECORD_FIELDS))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return records[np.argsort(records[:, 0])]
    import torch
    world = dist.get_world_size()
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([records.shape[0]], dtype=torch.int64, device=device))
    kmax = int(max(int(c.item()) for c in counts))
    pad = np.full((kmax, len(RECORD_FIELDS)), -1.0)
    pad[:records.shape[0]] = records
    bufs = [torch.zeros(kmax, len(RECORD_FIELDS), dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(pad).to(device) if device is not None else torch.from_numpy(pad))
    table = np.concatenate([b.cpu().numpy()[: int(c.item())] for b, c in zip(bufs, counts)], axis=0)
    return table[np.argsort(table[:, 0])]
