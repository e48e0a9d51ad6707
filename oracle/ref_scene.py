"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): numpy restatement of the beam scene of the reference's
sample admm_anderson_hard_zxu/samples/Asia2019/beams.cpp:66-160 (mcl::factory::make_tet_blocks + centre / scale /
offset, weighted_masses at 1522 kg/m^3, find_pins, stretch_beams), so that the CPU arms of bench.py
(`--impl reference`, `cpu_baseline`) and the golden generators can build the scene WITHOUT loading the product's
libraries. tests/test_oracle_cpu.py pins it bitwise against the reference's own generator (oracle/_ref) and against
the product's host scene builder."""
import numpy as np

_CORNER = np.array([[1, 1, 1], [0, 1, 1], [0, 1, 0], [1, 1, 0], [1, 0, 1], [0, 0, 1], [0, 0, 0], [1, 0, 0]])
_SPLIT = np.array([[0, 5, 7, 4], [5, 7, 2, 0], [5, 0, 2, 1], [7, 2, 0, 3], [5, 2, 7, 6]])


def make_beam(cx, cy, cz, y_shift=0.0, density=1522.0):
    """verts float32 (nv,3), tets int32 (nt,4), masses float32 (nv,). Vertices are numbered by first appearance while
    the cubes are visited x-major (ShapeFactory.hpp:452-488 followed by the lowest-index merge of refine())."""
    cx, cy, cz = max(1, cx), max(1, cy), max(1, cz)
    ny, nz = cy + 1, cz + 1
    X, Y, Z = np.meshgrid(np.arange(cx), np.arange(cy), np.arange(cz), indexing="ij")
    cube = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)                       # cubes in visiting order
    g = cube[:, None, :] + _CORNER[None, :, :]                                 # (cubes, 8, 3) grid coordinates
    lin = ((g[..., 0] * ny + g[..., 1]) * nz + g[..., 2]).ravel()                # visit sequence of grid vertices
    uniq, first = np.unique(lin, return_index=True)
    order = np.argsort(first, kind="stable")                                   # vertex id = rank of first appearance
    vid_of_lin = np.empty(uniq.size, np.int64)
    vid_of_lin[order] = np.arange(uniq.size)
    ids = vid_of_lin[np.searchsorted(uniq, lin)].reshape(-1, 8)
    tets = ids[:, _SPLIT].reshape(-1, 4).astype(np.int32)
    glin = uniq[order]
    gverts = np.stack([glin // (ny * nz), (glin // nz) % ny, glin % nz], 1)
    # float32 affine map of beams.cpp:83-90: s = 1/size_y, v' = s*v + s*(-centre), then the y offset
    f32 = np.float32
    mx = np.array([cx, cy, cz], f32)
    s = f32(1.0) / (mx[1] - f32(0.0))
    cen = (f32(0.0) + mx) / f32(2.0)
    tr = (s * (-cen)).astype(f32)
    verts = (s * gverts.astype(f32)).astype(f32) + tr
    if y_shift != 0.0:
        verts[:, 1] = verts[:, 1] + f32(y_shift)
    verts = verts.astype(f32)
    # weighted_masses: float32 volumes (Eigen's fixed-size 3x3 determinant order), sequential float adds in tet order
    p0 = verts[tets[:, 0]]
    e = [verts[tets[:, c + 1]] - p0 for c in range(3)]          # e[c][:, r] = column c, row r

    def h(a, b, c):  # m[0][a] * (m[1][b] * m[2][c] - m[1][c] * m[2][b]), m[r][c] = e[c][:, r]
        return (e[a][:, 0] * (e[b][:, 1] * e[c][:, 2] - e[c][:, 1] * e[b][:, 2])).astype(f32)

    det = (h(0, 1, 2) - h(1, 0, 2)).astype(f32) + h(2, 0, 1)
    vol = np.abs(det / f32(6.0)).astype(f32)
    quarter = ((f32(density) * vol).astype(f32) / f32(4.0)).astype(f32)
    masses = np.zeros(len(verts), f32)
    np.add.at(masses, tets.ravel(), np.repeat(quarter, 4))
    return verts, tets, masses


class RefBeamScene:
    """Same surface as aa_admm_b200.BeamScene (add / arrays / stretch), numpy only."""

    def __init__(self):
        self.verts, self.tets, self.masses = [], [], []
        self.pidx, self.ppts, self.pside = [], [], []
        self.nv = 0

    def add(self, cx, cy, cz, y_shift=0.0, density=1522.0):
        v, t, m = make_beam(cx, cy, cz, y_shift, density)
        lo, hi = v[:, 0].min(), v[:, 0].max()
        min_x, max_x = np.float32(lo + np.float32(1e-2)), np.float32(hi - np.float32(1e-2))
        for j in np.nonzero((v[:, 0] < min_x) | (v[:, 0] > max_x))[0]:   # find_pins: vertex order, left test first
            if v[j, 0] < min_x:
                self.pidx.append(j + self.nv), self.ppts.append(v[j].astype(np.float64)), self.pside.append(0)
            if v[j, 0] > max_x:
                self.pidx.append(j + self.nv), self.ppts.append(v[j].astype(np.float64)), self.pside.append(1)
        self.verts.append(v), self.tets.append(t + self.nv), self.masses.append(m)
        self.nv += len(v)
        return self

    def arrays(self):
        return (np.concatenate(self.verts), np.concatenate(self.tets).astype(np.int32), np.concatenate(self.masses),
                np.array(self.pidx, np.int32), np.array(self.ppts, np.float64).reshape(-1, 3), np.array(self.pside, np.int32))

    def stretch(self, dt):
        """stretch_beams (beams.cpp:74-87): the left pins move by -dt, the right pins by +dt along x; returns the targets."""
        p = np.array(self.ppts, np.float64).reshape(-1, 3)
        side = np.array(self.pside)
        p[side == 0, 0] -= 1.0 * dt
        p[side == 1, 0] += 1.0 * dt
        self.ppts = list(p)
        return p.copy()
