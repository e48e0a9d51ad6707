#include "beam_scene.hpp"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace aaadmm {

namespace {
// float 3x3 determinant in the operation order of Eigen 3.3.4's fixed-size path
// (Eigen/src/LU/Determinant.h: bruteforce_det3_helper), which weighted_masses relies on.
inline float det3_helper(const float m[3][3], int a, int b, int c) {
    return m[0][a] * (m[1][b] * m[2][c] - m[1][c] * m[2][b]);
}
inline float det3f(const float m[3][3]) {
    return det3_helper(m, 0, 1, 2) - det3_helper(m, 1, 0, 2) + det3_helper(m, 2, 0, 1);
}
}  // namespace

BeamMesh make_beam(int cx, int cy, int cz, float y_shift, float density) {
    cx = std::max(1, cx);
    cy = std::max(1, cy);
    cz = std::max(1, cz);
    BeamMesh mesh;
    const int nx = cx + 1, ny = cy + 1, nz = cz + 1;
    std::vector<int> grid_id((size_t)nx * ny * nz, -1);
    auto gid = [&](int x, int y, int z) -> int & { return grid_id[((size_t)x * ny + y) * nz + z]; };
    mesh.tets.reserve((size_t)cx * cy * cz * 20);
    std::vector<int> gverts;  // integer grid coordinates, 3 per vertex
    // Corner order a..h and the 5-tet split of ShapeFactory.hpp:452-488; vertices are
    // numbered by first appearance, which is what refine()'s lowest-index merge produces.
    static const int corner[8][3] = {{1, 1, 1}, {0, 1, 1}, {0, 1, 0}, {1, 1, 0},
                                     {1, 0, 1}, {0, 0, 1}, {0, 0, 0}, {1, 0, 0}};
    static const int split[5][4] = {{0, 5, 7, 4}, {5, 7, 2, 0}, {5, 0, 2, 1}, {7, 2, 0, 3}, {5, 2, 7, 6}};
    for (int x = 0; x < cx; ++x)
        for (int y = 0; y < cy; ++y)
            for (int z = 0; z < cz; ++z) {
                int id[8];
                for (int k = 0; k < 8; ++k) {
                    int gx = x + corner[k][0], gy = y + corner[k][1], gz = z + corner[k][2];
                    int &slot = gid(gx, gy, gz);
                    if (slot < 0) {
                        slot = (int)(gverts.size() / 3);
                        gverts.push_back(gx);
                        gverts.push_back(gy);
                        gverts.push_back(gz);
                    }
                    id[k] = slot;
                }
                for (int t = 0; t < 5; ++t)
                    for (int k = 0; k < 4; ++k) mesh.tets.push_back(id[split[t][k]]);
            }
    const int nv = (int)(gverts.size() / 3);
    // bounds of the integer grid: min 0, max (cx,cy,cz); beams.cpp:83-90 in float32:
    //   center = (min+max)/2, s = 1/size_y, v' = s*v + s*(-center)   (Eigen affine product)
    const float mx[3] = {(float)cx, (float)cy, (float)cz};
    float cen[3], tr[3];
    const float s = 1.f / (mx[1] - 0.f);
    for (int k = 0; k < 3; ++k) {
        cen[k] = (0.f + mx[k]) / 2.f;
        tr[k] = s * (-cen[k]);
    }
    mesh.verts.resize((size_t)nv * 3);
    for (int i = 0; i < nv; ++i)
        for (int k = 0; k < 3; ++k) {
            float v = s * (float)gverts[3 * i + k] + tr[k];
            if (k == 1 && y_shift != 0.f) v = v + y_shift;
            mesh.verts[3 * i + k] = v;
        }
    // weighted_masses: float32 volume, tet order, 4 sequential float adds per tet.
    mesh.masses.assign(nv, 0.f);
    const int nt = mesh.n_tets();
    for (int t = 0; t < nt; ++t) {
        const int *tet = &mesh.tets[4 * (size_t)t];
        float e[3][3];
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) e[r][c] = mesh.verts[3 * tet[c + 1] + r] - mesh.verts[3 * tet[0] + r];
        float v = std::fabs(det3f(e) / 6.f);
        float tet_mass = density * v;
        for (int k = 0; k < 4; ++k) mesh.masses[tet[k]] += tet_mass / 4.f;
    }
    return mesh;
}

void append_mesh(BeamMesh &a, const BeamMesh &b) {
    const int off = a.n_verts();
    a.verts.insert(a.verts.end(), b.verts.begin(), b.verts.end());
    a.masses.insert(a.masses.end(), b.masses.begin(), b.masses.end());
    for (int t : b.tets) a.tets.push_back(t + off);
}

void find_pins(const BeamMesh &m, int vertex_offset, BeamPins &pins) {
    const int nv = m.n_verts();
    float lo = m.verts[0], hi = m.verts[0];
    for (int i = 1; i < nv; ++i) {
        lo = std::min(lo, m.verts[3 * i]);
        hi = std::max(hi, m.verts[3 * i]);
    }
    const float min_x = lo + 1e-2f, max_x = hi - 1e-2f;
    for (int j = 0; j < nv; ++j) {
        const float *v = &m.verts[3 * (size_t)j];
        if (v[0] < min_x) {
            pins.idx.push_back(j + vertex_offset);
            for (int k = 0; k < 3; ++k) pins.points.push_back((double)v[k]);
            pins.side.push_back(0);
        }
        if (v[0] > max_x) {
            pins.idx.push_back(j + vertex_offset);
            for (int k = 0; k < 3; ++k) pins.points.push_back((double)v[k]);
            pins.side.push_back(1);
        }
    }
}

void stretch_pins(BeamPins &pins, double dt) {
    // Eigen::Vector3d(1.f,0,0)*dt ; all_points -= / += move   (beams.cpp:74-87)
    const double move[3] = {1.0 * dt, 0.0 * dt, 0.0 * dt};
    const size_t n = pins.idx.size();
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            if (pins.side[i] == 0)
                pins.points[3 * i + k] -= move[k];
            else
                pins.points[3 * i + k] += move[k];
        }
}

}  // namespace aaadmm
