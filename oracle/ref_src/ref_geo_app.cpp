// TEST INFRASTRUCTURE ONLY (oracle). The reference's own Geometry applications, unmodified, compiled
// with their main() renamed so that the golden-fixture generator can run them in-process:
//   -DREF_APP_PLANARITY  -> Geometry/PlanarityOpt.cpp   (oracle/_ref/libref_planarity.so)
//   -DREF_APP_WIREMESH   -> Geometry/WireMeshOpt.cpp    (oracle/_ref/libref_wiremesh.so)
#define main ref_app_main_impl
#ifdef REF_APP_PLANARITY
#include "PlanarityOpt.cpp"
#else
#include "WireMeshOpt.cpp"
#endif
#undef main
extern "C" int ref_app_main(int argc, char **argv) { return ref_app_main_impl(argc, argv); }
#ifdef REF_APP_WIREMESH
// the subdivided + smoothed input mesh of WireMeshOpt (MeshTypes.h:214-342), for fixture generation
extern "C" int ref_app_subdivided(const char *path, double *verts, int *quads, int *n_verts, int *n_quads, double *edge_len) {
    PolyMesh mesh;
    if (!OpenMesh::IO::read_mesh(mesh, path)) return -1;
    *edge_len = average_edge_length(mesh) * 0.5;
    PolyMesh sub = subdivide_and_smooth_mesh(mesh);
    *n_verts = (int)sub.n_vertices();
    *n_quads = (int)sub.n_faces();
    if (!verts) return 0;
    Matrix3X p;
    get_vertex_points(sub, p);
    memcpy(verts, p.data(), sizeof(double) * 3 * sub.n_vertices());
    int f = 0;
    for (PolyMesh::ConstFaceIter f_it = sub.faces_begin(); f_it != sub.faces_end(); ++f_it, ++f) {
        int k = 0;
        for (PolyMesh::ConstFaceVertexIter fv = sub.cfv_iter(*f_it); fv.is_valid(); ++fv) quads[4 * f + (k++)] = fv->idx();
    }
    return 0;
}
// edge list in OpenMesh edge order (halfedge 0 from/to), as WireMeshOpt.cpp:277-284 iterates it
extern "C" int ref_app_edges(const char *path, int *edges, int *n_edges) {
    PolyMesh mesh;
    if (!OpenMesh::IO::read_mesh(mesh, path)) return -1;
    PolyMesh sub = subdivide_and_smooth_mesh(mesh);
    *n_edges = (int)sub.n_edges();
    if (!edges) return 0;
    int e = 0;
    for (PolyMesh::ConstEdgeIter ce = sub.edges_begin(); ce != sub.edges_end(); ++ce, ++e) {
        PolyMesh::HalfedgeHandle heh = sub.halfedge_handle(*ce, 0);
        edges[2 * e] = sub.from_vertex_handle(heh).idx();
        edges[2 * e + 1] = sub.to_vertex_handle(heh).idx();
    }
    return 0;
}
#endif
