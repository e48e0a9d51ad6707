"""Regenerates tests/golden/hard_cloth_*.npz from the UNMODIFIED reference compiled into oracle/_ref
(run in the build container, where /root/reference exists):  python tests/golden/make_golden_cloth.py
Scenes: tests/scenes.py run_cloth (TriEnergyTerm cloth under the hard_zxu ordering, alone and together with a
tet beam in the same solver). Flags of the reference build are pinned in oracle/Makefile."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import aa_admm_b200 as A  # noqa: E402  (host-side scene builder only)
from oracle import refbind as R  # noqa: E402
from scenes import run_cloth, run_plinko  # noqa: E402

CASES = {
    "hard_cloth_8_m5": dict(n=8, m=5, accel=True, limits=(-100.0, 100.0), beam=None),
    "hard_cloth_8_noacc_limits": dict(n=8, m=0, accel=False, limits=(0.95, 1.05), beam=None),
    "hard_cloth_6_beam_6x2x2_m5": dict(n=6, m=5, accel=True, limits=(0.9, 1.1), beam=(6, 2, 2)),
    # windyflag.cpp's material, strain limits and wind on the synthetic cloth; pins stay in place
    "hard_windyflag_10_m5": dict(n=10, m=5, accel=True, limits=(0.95, 1.05), beam=None, youngs=50.0, poisson=0.1,
                                 wind=(25.0, 0.0, 5.0), pin_speed=0.0, frames=3, iters=100),
}


def case(n, m, accel, limits, beam, frames=2, iters=60, youngs=1e5, poisson=0.3, wind=None, pin_speed=0.3):
    hist, xs = run_cloth(lambda: R.RefSolver("hard"), frames=frames, n=n, m=m, accel=accel, iters=iters, limits=limits,
                         with_beam=(A, beam) if beam else None, youngs=youngs, poisson=poisson, wind=wind,
                         pin_speed=pin_speed)
    prim, comb, rej, rows = [], [], [], []
    for h in hist:
        p, c, r = np.zeros(iters), np.zeros(iters), np.zeros(iters)
        p[:len(h)], c[:len(h)], r[:len(h)] = h[:, 1], h[:, 2], h[:, 3]
        prim.append(p), comb.append(c), rej.append(r), rows.append(len(h))
    return dict(prim=np.array(prim), comb=np.array(comb), rej=np.array(rej), rows=np.array(rows), x=np.array(xs),
                n=n, m=m, accel=int(accel), limits=np.array(limits), beam=np.array(beam if beam else (0, 0, 0)), iters=iters,
                frames=frames, youngs=youngs, poisson=poisson, wind=np.array(wind if wind else (0.0, 0.0, 0.0)),
                pin_speed=pin_speed)


def plinko_case(m, accel, dims=(8, 2, 2), frames=3, iters=40):
    """Free beam on analytic obstacles with a Collision term on every vertex (tests/scenes.py run_plinko)."""
    hist, xs = run_plinko(lambda: R.RefSolver("hard"), A, frames=frames, dims=dims, m=m, accel=accel, iters=iters)
    comb, rej, rows = [], [], []
    for h in hist:
        c, r = np.zeros(iters), np.zeros(iters)
        c[:len(h)], r[:len(h)] = h[:, 2], h[:, 3]
        comb.append(c), rej.append(r), rows.append(len(h))
    return dict(comb=np.array(comb), rej=np.array(rej), rows=np.array(rows), x=np.array(xs), dims=np.array(dims), m=m,
                accel=int(accel), frames=frames, iters=iters)


if __name__ == "__main__":
    assert R.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    np.savez_compressed(os.path.join(HERE, "hard_plinko_8x2x2_m5.npz"), **plinko_case(5, True))
    np.savez_compressed(os.path.join(HERE, "hard_plinko_8x2x2_noacc.npz"), **plinko_case(0, False))
    for name, kw in CASES.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **case(**kw))
        print("wrote", name)
