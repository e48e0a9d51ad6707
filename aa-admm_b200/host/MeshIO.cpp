#include "MeshIO.hpp"

#include <algorithm>

#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <numeric>
#include <sstream>
#include <stdexcept>

namespace mcl {

namespace {
// Matrix3f::determinant() of the reference (Eigen/src/LU/Determinant.h: bruteforce_det3_helper), m[row][col].
inline float det3_helper(const float m[3][3], int a, int b, int c) { return m[0][a] * (m[1][b] * m[2][c] - m[1][c] * m[2][b]); }
inline float det3f(const float m[3][3]) { return det3_helper(m, 0, 1, 2) - det3_helper(m, 1, 0, 2) + det3_helper(m, 2, 0, 1); }
inline Vec3f sub(const Vec3f &a, const Vec3f &b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
inline Vec3f cross(const Vec3f &a, const Vec3f &b) {
    return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
// Eigen's unrolled reduction of a 3-vector splits it 1 + 2: x0 + (x1 + x2) (Core/Redux.h redux_novec_unroller)
inline float dot(const Vec3f &a, const Vec3f &b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
}  // namespace

// TetMesh.hpp:297-315
void TetMesh::weighted_masses(std::vector<float> &m, float density_kgm3) {
    m.resize(vertices.size(), 0.f);
    const int n_tets = (int)tets.size();
    for (int t = 0; t < n_tets; ++t) {
        const Vec4i tet = tets[t];
        float e[3][3];
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) e[r][c] = vertices[tet[c + 1]][r] - vertices[tet[0]][r];
        const float v = std::abs(det3f(e) / 6.f);
        const float tet_mass = density_kgm3 * v;
        for (int k = 0; k < 4; ++k) m[tet[k]] += tet_mass / 4.f;
    }
}

// TriangleMesh.hpp:281-296
void TriangleMesh::weighted_masses(std::vector<float> &m, float density_kgm2) {
    m.resize(vertices.size(), 0.f);
    const int n_faces = (int)faces.size();
    for (int f = 0; f < n_faces; ++f) {
        const Vec3i face = faces[f];
        const Vec3f n = cross(sub(vertices[face[1]], vertices[face[0]]), sub(vertices[face[2]], vertices[face[0]]));
        const float area = 0.5f * std::sqrt(dot(n, n));
        const float tri_mass = density_kgm2 * area;
        for (int k = 0; k < 3; ++k) m[face[k]] += tri_mass / 3.f;
    }
}

namespace meshio {

// MeshIO.hpp:55-130
bool load_obj(TriangleMesh *mesh, std::string file) {
    mesh->clear();
    std::ifstream infile(file.c_str());
    if (!infile.is_open()) {
        std::cerr << "\n**mcl::meshio::load_obj Error: Could not open file " << file << std::endl;
        return false;
    }
    std::string line;
    while (std::getline(infile, line)) {
        std::stringstream ss(line);
        std::string tok;
        ss >> tok;
        for (char &c : tok) c = (char)::tolower(c);
        if (tok == "v") {
            float x = 0, y = 0, z = 0;
            ss >> x >> y >> z;
            mesh->vertices.push_back({x, y, z});
        } else if (tok == "f") {
            Vec3i face = {-1, -1, -1};
            for (int i = 0; i < 3; ++i) {  // the first index of each `a/b/c` token, 1-based in the file
                std::string f_str, s2;
                ss >> f_str;
                std::stringstream ss2(f_str);
                bool have = false;
                while (std::getline(ss2, s2, '/')) {
                    if (s2.empty()) continue;
                    const int val = std::stoi(s2) - 1;
                    if (!have) face[i] = val, have = true;
                }
            }
            if (face[0] >= 0 && face[1] >= 0 && face[2] >= 0) mesh->faces.push_back(face);
        }
    }
    return true;
}

// MeshIO.hpp:133-183 (no normals / texture coordinates are kept here: plain `f a b c` records)
bool save_obj(const TriangleMesh *mesh, std::string filename) {
    const size_t fsize = filename.size();
    if (fsize < 4 || filename.substr(fsize - 4, 4) != ".obj") {
        printf("\n**TriangleMesh::save Error: Filetype must be .obj\n");
        return false;
    }
    std::ofstream fs(filename.c_str());
    fs << "# Generated with mclscene by Matt Overby (www.mattoverby.net)";
    for (const Vec3f &v : mesh->vertices) fs << "\nv " << v[0] << ' ' << v[1] << ' ' << v[2];
    for (const Vec3i &f : mesh->faces) fs << "\nf " << f[0] + 1 << ' ' << f[1] + 1 << ' ' << f[2] + 1;
    fs << "\n";
    return true;
}

// MeshIO.hpp:186-296
bool load_elenode(TetMesh *mesh, std::string file) {
    mesh->clear();
    {
        const std::string name = file + ".ele";
        std::ifstream fs(name.c_str());
        if (!fs) {
            std::cerr << "\n**TetMesh Error: Could not load " << name << std::endl;
            return false;
        }
        std::string header;
        std::getline(fs, header);
        std::stringstream hs(header);
        int n_tets = 0;
        hs >> n_tets;
        mesh->tets.resize(std::max(n_tets, 0));
        std::vector<int> seen(mesh->tets.size(), 0);
        bool starts_with_one = false;
        for (int i = 0; i < n_tets; ++i) {
            std::string line;
            std::getline(fs, line);
            std::stringstream ls(line);
            size_t idx = 0;
            int ids[4] = {0, 0, 0, 0};
            ls >> idx >> ids[0] >> ids[1] >> ids[2] >> ids[3];
            if (i == 0 && idx == 1) starts_with_one = true;
            if (starts_with_one) {
                idx -= 1;
                for (int j = 0; j < 4; ++j) ids[j] -= 1;
            }
            if (idx >= mesh->tets.size()) {  // the reference tests `>` and then writes out of bounds for `==`
                std::cerr << "\n**TetMesh Error: Your indices are bad for file " << name << std::endl;
                return false;
            }
            mesh->tets[idx] = {ids[0], ids[1], ids[2], ids[3]};
            seen[idx] = 1;
        }
        for (int s : seen)
            if (!s) {
                std::cerr << "\n**TetMesh Error: Your indices are bad for file " << name << std::endl;
                return false;
            }
    }
    {
        const std::string name = file + ".node";
        std::ifstream fs(name.c_str());
        if (!fs) {
            std::cerr << "\n**TetMesh Error: Could not load " << name << std::endl;
            return false;
        }
        std::string header;
        std::getline(fs, header);
        std::stringstream hs(header);
        int n_nodes = 0;
        hs >> n_nodes;
        mesh->vertices.resize(std::max(n_nodes, 0));
        std::vector<int> seen(mesh->vertices.size(), 0);
        bool starts_with_one = false;
        for (int i = 0; i < n_nodes; ++i) {
            std::string line;
            std::getline(fs, line);
            std::stringstream ls(line);
            double x = 0, y = 0, z = 0;
            size_t idx = 0;
            ls >> idx >> x >> y >> z;
            if (i == 0 && idx == 1) starts_with_one = true;
            if (starts_with_one) idx -= 1;
            if (idx >= mesh->vertices.size()) {
                std::cerr << "\n**TetMesh Error: Your indices are bad for file " << name << std::endl;
                return false;
            }
            mesh->vertices[idx] = {(float)x, (float)y, (float)z};  // parsed as double, stored as float32
            seen[idx] = 1;
        }
        for (int s : seen)
            if (!s) {
                std::cerr << "\n**TetMesh Error: Your indices are bad for file " << name << std::endl;
                return false;
            }
    }
    // inverted tets are re-ordered (float32 signed volume)
    for (Vec4i &t : mesh->tets) {
        for (int k = 0; k < 4; ++k)
            if (t[k] < 0 || (size_t)t[k] >= mesh->vertices.size()) {
                std::cerr << "\n**TetMesh Error: Your indices are bad for file " << file << ".ele" << std::endl;
                return false;
            }
        const Vec3f a = mesh->vertices[t[0]];
        const float V = dot(sub(mesh->vertices[t[1]], a), cross(sub(mesh->vertices[t[2]], a), sub(mesh->vertices[t[3]], a))) / 6.f;
        if (V < 0) std::swap(t[1], t[2]);
    }
    if (mesh->vertices.empty() || mesh->tets.empty()) throw std::runtime_error("\n**TetMesh Error: Problem loading files");
    return true;
}

// MeshIO.hpp:299-345: "<n> 4 0" / "<n> 3 0 0" headers, 0-based rows
bool save_elenode(const TetMesh *mesh, std::string file) {
    {
        std::ofstream fs((file + ".ele").c_str());
        if (!fs) return false;
        fs << mesh->tets.size() << " 4 0\n";
        for (size_t i = 0; i < mesh->tets.size(); ++i) {
            const Vec4i &t = mesh->tets[i];
            fs << "\t" << i << ' ' << t[0] << ' ' << t[1] << ' ' << t[2] << ' ' << t[3] << "\n";
        }
    }
    {
        std::ofstream fs((file + ".node").c_str());
        if (!fs) return false;
        fs << mesh->vertices.size() << " 3 0 0\n";
        fs.precision(9);  // enough to restore a float32 exactly
        for (size_t i = 0; i < mesh->vertices.size(); ++i) {
            const Vec3f &v = mesh->vertices[i];
            fs << "\t" << i << ' ' << v[0] << ' ' << v[1] << ' ' << v[2] << "\n";
        }
    }
    return true;
}

}  // namespace meshio
}  // namespace mcl

namespace binding {

// AddMeshes.hpp:97-177
}  // namespace binding

namespace mcl {
void TetMesh::surface_inds(std::vector<int> &surf_inds) {
    // a face is on the surface when it belongs to one tet only
    std::vector<std::array<int, 3>> faces;
    faces.reserve(tets.size() * 4);
    static const int fc[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
    for (const Vec4i &t : tets)
        for (int f = 0; f < 4; ++f) {
            std::array<int, 3> k = {t[fc[f][0]], t[fc[f][1]], t[fc[f][2]]};
            std::sort(k.begin(), k.end());
            faces.push_back(k);
        }
    std::sort(faces.begin(), faces.end());
    std::vector<char> on(vertices.size(), 0);
    for (size_t i = 0; i < faces.size();) {
        size_t j = i + 1;
        while (j < faces.size() && faces[j] == faces[i]) ++j;
        if (j - i == 1)
            for (int v : faces[i]) on[v] = 1;
        i = j;
    }
    for (size_t v = 0; v < on.size(); ++v)
        if (on[v]) surf_inds.push_back((int)v);
}
}  // namespace mcl

namespace binding {
void add_tetmesh(admm::Solver *solver, std::shared_ptr<mcl::TetMesh> &mesh, const admm::Lame &lame, bool verbose) {
    const int num_tet_verts = (int)mesh->vertices.size();
    const int prev_tet_verts = (int)solver->m_x.size() / 3;
    const int num_tets = (int)mesh->tets.size();
    std::vector<float> masses;
    mesh->weighted_masses(masses, 1522.f);
    for (float m : masses)
        if (m <= 0.f) throw std::runtime_error("TetMesh Error: Zero mass");
    solver->m_x.resize((size_t)(prev_tet_verts + num_tet_verts) * 3);
    solver->m_v.resize((size_t)(prev_tet_verts + num_tet_verts) * 3, 0.0);
    solver->m_masses.resize((size_t)(prev_tet_verts + num_tet_verts) * 3);
    for (int i = 0; i < num_tet_verts; ++i)
        for (int j = 0; j < 3; ++j) {
            solver->m_x[(size_t)(i + prev_tet_verts) * 3 + j] = (double)mesh->vertices[i][j];
            solver->m_masses[(size_t)(i + prev_tet_verts) * 3 + j] = (double)masses[i];
        }
    {  // AddMeshes.hpp:130-136
        std::vector<int> surf_inds;
        mesh->surface_inds(surf_inds);
        for (int i : surf_inds) solver->surface_inds.emplace_back(i + prev_tet_verts);
    }
    const float *verts = num_tet_verts ? &mesh->vertices[0][0] : nullptr;
    const int *tets = num_tets ? &mesh->tets[0][0] : nullptr;
    if ((mesh->flags & LINEAR) || mesh->flags == 0)
        admm::create_tets_from_mesh<float, admm::TetEnergyTerm>(solver->energyterms, verts, tets, num_tets, lame, prev_tet_verts);
    else if (mesh->flags & NEOHOOKEAN)
        admm::create_tets_from_mesh<float, admm::NeoHookeanTet>(solver->energyterms, verts, tets, num_tets, lame, prev_tet_verts);
    else if (mesh->flags & STVK)
        admm::create_tets_from_mesh<float, admm::StVKTet>(solver->energyterms, verts, tets, num_tets, lame, prev_tet_verts);
    if (verbose)
        std::cout << "Added mesh: "
                  << "\n\tmass: " << std::accumulate(masses.begin(), masses.end(), 0.f) << "kg"
                  << "\n\tvertices: " << num_tet_verts << "\n\ttets: " << num_tets
                  << "\n\t(total) verts: " << solver->m_x.size() / 3 << std::endl;
}

// AddMeshes.hpp:180-237
void add_trimesh(admm::Solver *solver, std::shared_ptr<mcl::TriangleMesh> &mesh, const admm::Lame &lame, bool verbose) {
    const int num_tri_verts = (int)mesh->vertices.size();
    const int prev_tri_verts = (int)solver->m_x.size() / 3;
    const int num_tris = (int)mesh->faces.size();
    std::vector<float> masses;
    mesh->weighted_masses(masses, 1.0f);
    for (float m : masses)
        if (m <= 0.f) throw std::runtime_error("TriMesh Error: Zero mass");
    solver->m_x.resize((size_t)(prev_tri_verts + num_tri_verts) * 3);
    solver->m_v.resize((size_t)(prev_tri_verts + num_tri_verts) * 3, 0.0);
    solver->m_masses.resize((size_t)(prev_tri_verts + num_tri_verts) * 3);
    for (int i = 0; i < num_tri_verts; ++i)
        for (int j = 0; j < 3; ++j) {
            solver->m_x[(size_t)(i + prev_tri_verts) * 3 + j] = (double)mesh->vertices[i][j];
            solver->m_masses[(size_t)(i + prev_tri_verts) * 3 + j] = (double)masses[i];
        }
    if ((mesh->flags & LINEAR) || mesh->flags == 0)
        admm::create_tris_from_mesh<float, admm::TriEnergyTerm>(solver->energyterms, num_tri_verts ? &mesh->vertices[0][0] : nullptr,
                                                                num_tris ? &mesh->faces[0][0] : nullptr, num_tris, lame,
                                                                prev_tri_verts);
    else
        throw std::runtime_error("**binding::add_trimesh Error: Unknown triangle mesh material type");
    if (verbose)
        std::cout << "Added mesh: "
                  << "\n\tmass: " << std::accumulate(masses.begin(), masses.end(), 0.f) << "kg"
                  << "\n\tvertices: " << num_tri_verts << "\n\ttris: " << num_tris
                  << "\n\t(total) verts: " << solver->m_x.size() / 3 << std::endl;
}

}  // namespace binding
