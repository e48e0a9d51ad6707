"""Ensemble sharding (BASELINE configs[4]): independent scenes, static scene -> rank map,
no per-iteration collective; one gather of fixed-size result records at the end."""
import numpy as np

RECORD_FIELDS = ("scene_id", "iters", "rejects", "final_prim", "final_comb", "loop_ms", "wall_ms", "rank")


def scene_material(s):
    """Material sweep of cfg 5: E_s = 10^(6 + 2*floor(s/8)/7) Pa, nu_s = 0.30 + 0.02*(s mod 8)."""
    return 10.0 ** (6.0 + 2.0 * (s // 8) / 7.0), 0.30 + 0.02 * (s % 8)


def scenes_of_rank(n_scenes, rank, world):
    """Static scene -> GPU map: scene s runs on GPU (s + s // 8) mod G. The sweep is an 8 x 8 grid (stiffness group
    s // 8, Poisson ratio s mod 8) and both axes change the iterations a scene needs (36 ... 100): the diagonal map
    gives every GPU each stiffness and each Poisson ratio equally often, where the plain s mod G of round 1 gave GPU r
    the r-th Poisson ratio only (20 % more iterations on the last GPU than on the first)."""
    return [s for s in range(n_scenes) if (s + s // 8) % world == rank]


def scenes_by_cost(table, rank, world):
    """Scene -> GPU map of the NEXT pass of a sweep from the gathered records of the previous one (gather_records: every
    rank holds the same table). The scenes of a sweep share one mesh, so any resident slot can run any member; their cost
    (iterations to the tolerance, 36 ... 100 at cfg 5) repeats from pass to pass. Longest-processing-time-first: the scenes
    in descending order of their measured device-loop time, each to the GPU with the least load so far (ties: lowest
    rank). Deterministic: every rank computes the same partition. A rank's list stays in descending order, so its slot
    threads finish the pass on the short members."""
    table = np.asarray(table, dtype=np.float64).reshape(-1, len(RECORD_FIELDS))
    order = sorted(range(table.shape[0]), key=lambda i: (-table[i, 5], table[i, 0]))
    load = [0.0] * world
    parts = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += table[i, 5]
        parts[r].append(int(table[i, 0]))
    return parts[rank]


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


class SceneSlot:
    """One resident scene of a sweep: a Solver whose mesh, pins, analysis (ordering, pattern of L, fronts, schedules)
    and device buffers persist; every member of the sweep only changes the material (Solver::set_material +
    incremental initialize(): values of the system matrix, numeric LDL^T on the device, element moduli)."""

    def __init__(self, A, dims, dt=1.0 / 30.0, iters=100, anderson_m=5, penalty=1.0, device=0):
        self.A, self.dims, self.dt, self.iters, self.m, self.penalty = A, tuple(dims), dt, iters, anderson_m, penalty
        self.device = device
        A.set_device(device)
        scene = A.BeamScene().add(*dims, 0.0)
        self.verts, self.tets, self.masses, self.pidx, self.ppts, self.pside = scene.arrays()
        self.rest = self.verts.astype(np.float64).reshape(-1)
        self.solver = None
        self.setup_ms = None

    def _pins_of_frame(self, p):
        # stretch_beams (beams.cpp:74-87): the left pins move by -dt, the right pins by +dt along x, per frame
        p[self.pside == 0, 0] -= 1.0 * self.dt
        p[self.pside == 1, 0] += 1.0 * self.dt
        return p

    def run_member(self, scene_id, frames=1, rank=0):
        """One member: material of `scene_id`, `frames` frames from the rest state. Returns the result record."""
        import time
        A = self.A
        t0 = time.perf_counter()
        youngs, poisson = scene_material(scene_id)
        if self.solver is None:  # first member on this slot: full setup
            self.solver = A.Solver()
            self.solver.add_tetmesh(self.verts, self.tets, self.masses, youngs, poisson, 0)
        else:
            self.solver.set_material(youngs, poisson)
            self.solver.set_x(self.rest)
        s = self.solver
        # the sample's call sequence (beams.cpp: stretch_beams before initialize and at the start of every frame)
        p = self._pins_of_frame(self.ppts.copy())
        s.set_pins(self.pidx, p)
        s.initialize(self.dt, self.iters, -9.8, self.m, True, self.penalty, A.ORDER_HARD_ZXU)
        t1 = time.perf_counter()
        iters = rejects = 0
        loop_ms = 0.0
        last = None
        for _ in range(frames):
            s.set_pins(self.pidx, self._pins_of_frame(p))
            last = s.step()  # rows: primal residual, combined residual, is_reject
            iters += last.shape[0]
            rejects += int(last[:, 2].sum())
            loop_ms += s.info()["loop_ms"]
        self.setup_ms = 1e3 * (t1 - t0)
        self.incremental = s.was_incremental()
        return make_record(scene_id, iters, rejects, last[-1, 0], last[-1, 1], loop_ms, 1e3 * (time.perf_counter() - t0), rank)


def run_sweep(A, dims, scene_ids, slots, frames=1, rank=0):
    """The scenes of this rank on `slots` (a list of SceneSlot of ONE GPU), one host thread per slot: while one member
    iterates on the GPU the next one's numeric setup runs (host: matrix values; device: numeric factorisation on the
    slot's own stream), and the loops of the resident scenes overlap on the device. Returns (records, per-slot setup ms)."""
    import queue
    import threading
    q = queue.Queue()
    for sid in scene_ids:
        q.put(sid)
    recs, errors, setups = [], [], []
    lock = threading.Lock()

    def work(slot, device):
        try:
            A.set_device(device)  # the CUDA device is per host thread
            while True:
                try:
                    sid = q.get_nowait()
                except queue.Empty:
                    return
                r = slot.run_member(sid, frames, rank)
                with lock:
                    recs.append(r)
                    setups.append((sid, slot.setup_ms, slot.incremental))
        except Exception as e:  # surfaced by the caller
            with lock:
                errors.append(e)

    threads = [threading.Thread(target=work, args=(sl, sl.device if hasattr(sl, "device") else 0)) for sl in slots]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return np.array(recs).reshape(-1, len(RECORD_FIELDS)), setups
