// C ABI of libaaadmm_b200.so (include/aaadmm.h): handle management and the stream-ordered
// launch sequences of the ADMM loops.  All numerics live in the kernels
// (tet_kernels.cuh, aa_kernels.cuh, ldlt_apply.cu); nothing here computes on the host.
#include "../../include/aaadmm.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "aa_kernels.cuh"
#include "common.cuh"
#include "ldlt_apply.cuh"
#include "extra_terms.cuh"
#include "tet_kernels.cuh"

namespace aaadmm {

static thread_local std::string g_err;
void set_last_error(const std::string &msg) { g_err = msg; }

int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaDeviceProp p;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&p, dev) == cudaSuccess)
            sms = p.multiProcessorCount;
        else
            sms = 148;
    }
    return sms;
}
int stream_grid(int ctas_per_sm) { return std::min(RED_MAX_BLOCKS, sm_count() * ctas_per_sm); }

#define AA_DISPATCH(m, CALL)                 \
    switch (m) {                             \
        case 1: { constexpr int MM = 1; CALL; } break;  \
        case 2: { constexpr int MM = 2; CALL; } break;  \
        case 3: { constexpr int MM = 3; CALL; } break;  \
        case 4: { constexpr int MM = 4; CALL; } break;  \
        case 5: { constexpr int MM = 5; CALL; } break;  \
        case 6: { constexpr int MM = 6; CALL; } break;  \
        case 7: { constexpr int MM = 7; CALL; } break;  \
        case 8: { constexpr int MM = 8; CALL; } break;  \
        default: { constexpr int MM = AA_MAX_M; CALL; } break; \
    }

int launch_aa_pass1(int m, int grid, cudaStream_t s, const double *g_u, const double *g_x, double *gx_dst,
                    double *ucur, double *dF, double *dG, int64_t Ne, int64_t Nt, SolveState *st, double *partials,
                    double *g_copy) {
    AA_DISPATCH(m, (k_aa_pass1<MM><<<grid, AA_BLOCK, 0, s>>>(g_u, g_x, gx_dst, ucur, dF, dG, Ne, Nt, st, partials, g_copy)));
    AAADMM_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_aa_pass2(int m, int grid, cudaStream_t s, const double *g_u, const double *g_x, double *ucur,
                    double *dF, double *dG, int64_t Ne, int64_t Nt, const SolveState *st) {
    AA_DISPATCH(m, (k_aa_pass2<MM><<<grid, AA_BLOCK, 0, s>>>(g_u, g_x, ucur, dF, dG, Ne, Nt, st)));
    AAADMM_CUDA_OK(cudaGetLastError());
    return 0;
}

// One warp: the QR through cod_qr_warp (what the Anderson passes use), the rest on lane 0.
__global__ void k_cod(int m, const double *M, const double *rhs, double *x, int *rank) {
    __shared__ double A[AA_MAX_M * AA_MAX_M], b[AA_MAX_M], xx[AA_MAX_M], hc[AA_MAX_M];
    __shared__ int trans[AA_MAX_M];
    __shared__ CodWarpWork w;
    for (int i = threadIdx.x; i < m * m; i += 32) A[i] = M[i];
    if (threadIdx.x < m) b[threadIdx.x] = rhs[threadIdx.x];
    __syncwarp();
    int nonzero_pivots;
    double maxpivot;
    cod_qr_warp(A, m, hc, trans, nonzero_pivots, maxpivot, w);
    if (threadIdx.x == 0) *rank = cod_finish(A, m, hc, trans, nonzero_pivots, maxpivot, b, xx);
    __syncwarp();
    if (threadIdx.x < m) x[threadIdx.x] = xx[threadIdx.x];
}

// AoS (9 consecutive doubles per tet, reference layout) <-> SoA planes
__global__ void k_soa_to_aos9(const double *__restrict__ soa, double *__restrict__ aos, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    for (int k = 0; k < 9; ++k) aos[9 * (size_t)t + k] = soa[(size_t)k * T + t];
}

__global__ void k_init_state(SolveState *st, int accel, int m, double eps, int max_iters) {
    st->loop_it = 0;
    st->max_iters = max_iters;
    st->skip_redo = 1;
    st->aa_skip = 0;
    st->prim2 = 0.0;
    st->hyper_prim2 = 0.0;
    st->tri_prim2 = 0.0;
    st->tri_comb = 0.0;
    st->pt_prim2 = 0.0;
    st->pt_comb = 0.0;
    st->prev_prim = 1e+20;
    st->comb = 0.0;
    st->eps = eps;
    st->reject = 0;
    st->done = 0;
    st->iter = 0;
    st->n_rejects = 0;
    st->accel = accel;
    st->aa_iter = 0;
    st->aa_col = 0;
    st->aa_m = m;
    st->aa_mk = 0;
    st->ticket = 0u;
}
// The iteration loop begins here (after the warm start): time base of the per-iteration time stamps.
__global__ void k_mark_loop_start(SolveState *st) { st->t0 = global_timer_ns(); }
// Tail of one loop turn inside the graph WHILE node: advance the device-side iteration counter and
// decide whether the body runs again (hard/src/Solver.cpp:130 `for` bound and :188 `break`).
__global__ void k_loop_cond(cudaGraphConditionalHandle h, SolveState *st) {
    const int it = st->loop_it + 1;
    st->loop_it = it;
    cudaGraphSetConditional(h, (!st->done && it < st->max_iters) ? 1u : 0u);
}
__global__ void k_aa_set_counters(SolveState *st, int iter, int col) {
    st->aa_iter = iter;
    st->aa_col = col;
}

}  // namespace aaadmm

using namespace aaadmm;

#define API_TRY_BEGIN try {
#define API_TRY_END                                     \
    }                                                   \
    catch (const std::exception &e) {                   \
        set_last_error(std::string("exception: ") + e.what()); \
        return -2;                                      \
    }

struct aaadmm_aa {
    int m = 0;
    int64_t Nt = 0, Ne = 0;
    double *ucur = nullptr, *dF = nullptr, *dG = nullptr, *partials = nullptr, *gtmp = nullptr;
    SolveState *st = nullptr;
    cudaStream_t stream = nullptr;
    int grid = 0;
};

struct aaadmm_ldlt {
    LdltDev *f = nullptr;
    cudaStream_t stream = nullptr;
    double *b_dev = nullptr, *x_dev = nullptr;
};

struct aaadmm_tetscene {
    int T = 0, V = 0, NF = 0, NP = 0;
    double rho_dt2 = 0;
    aaadmm_ldlt *factor = nullptr;
    cudaStream_t stream = nullptr;
    // constants
    int4 *idx = nullptr;
    double *binv = nullptr, *w = nullptr, *kvol = nullptr, *mass = nullptr;
    int *material = nullptr, *hyper_ids = nullptr;
    double *mu = nullptr, *lambda = nullptr, *volume = nullptr;
    int n_hyper = 0;
    // triangle terms (hard_zxu only): u / z planes follow the tets' inside Ubuf / Gbuf / z, contributions behind 12 T
    int NT = 0;
    int4 *tri_idx = nullptr;
    double *tri_rp = nullptr, *tri_w = nullptr, *tri_lmin = nullptr, *tri_lmax = nullptr;
    // collision terms: u / z planes behind the triangles', contributions behind theirs
    int NC = 0, n_objs = 0;
    int *pt_vert = nullptr, *obj_type = nullptr;
    double *pt_w = nullptr, *obj_prm = nullptr;
    int64_t *inc_ptr = nullptr;
    int *inc = nullptr;
    // state
    int64_t Ne = 0, Nt = 0, Nbuf = 0;  // Nbuf = 9T + 3V (pinned tail rides along)
    double *Ubuf = nullptr, *Gbuf = nullptr, *xs = nullptr, *z = nullptr, *contrib = nullptr;
    double *bconst = nullptr, *xbar = nullptr, *xpin = nullptr;
    double *dF = nullptr, *dG = nullptr;
    double *xz_a = nullptr, *xz_b = nullptr;  // xzu ordering: default_u and comb_z / last_z
    int hist_cap = 0, m_cap = 0;
    double *hist_prim = nullptr, *hist_comb = nullptr;
    int *hist_rej = nullptr;
    double *partials = nullptr;
    SolveState *st = nullptr;
    double *pin_h = nullptr, *xbar_h = nullptr, *xout_h = nullptr;  // pinned host staging
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int launches = 0;
    bool has_inputs = false;
    // the whole iteration loop as ONE graph launch: a conditional WHILE node whose body is one loop turn
    cudaGraph_t loop_graph = nullptr;
    cudaGraphExec_t loop_exec = nullptr;
    int loop_key = -1;  // accel * 64 + m the graph was built for
    int body_launches = 0;
    // where the last step left the tets' z and u (depends on ordering and acceleration): aaadmm_tetscene_read_zu
    const double *last_z = nullptr, *last_u = nullptr;
    double *zu_scratch = nullptr;
    SolveState *st_h = nullptr;  // pinned host copy of the control block
    int last_rows = 0, last_max_iters = 0;  // of the last step (aaadmm_tetscene_iteration_times)
};

extern "C" {

const char *aaadmm_last_error(void) { return g_err.c_str(); }

int aaadmm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int aaadmm_set_device(int device) {
    AAADMM_CUDA_OK(cudaSetDevice(device));
    return 0;
}
int aaadmm_device_sms(void) { return sm_count(); }

// ------------------------------------------------------------------------------------------
// Anderson acceleration object
// ------------------------------------------------------------------------------------------
int aaadmm_aa_create(aaadmm_aa **out, int m, int64_t total_dim, int64_t effective_dim) {
    API_TRY_BEGIN
    if (m <= 0 || m > AA_MAX_M || total_dim <= 0 || effective_dim <= 0 || effective_dim > total_dim) {
        set_last_error("aa_create: need 0 < m <= 16 and 0 < effective_dim <= total_dim");
        return -1;
    }
    if (aaadmm_device_count() <= 0) {
        set_last_error("aa_create: no CUDA device (this library has no CPU path)");
        return -1;
    }
    aaadmm_aa *a = new aaadmm_aa();
    a->m = m;
    a->Nt = total_dim;
    a->Ne = effective_dim;
    a->grid = stream_grid(4);
    AAADMM_CUDA_OK(cudaStreamCreate(&a->stream));
    AAADMM_CUDA_OK(cudaMalloc((void **)&a->ucur, sizeof(double) * total_dim));
    AAADMM_CUDA_OK(cudaMalloc((void **)&a->gtmp, sizeof(double) * total_dim));
    AAADMM_CUDA_OK(cudaMalloc((void **)&a->dF, sizeof(double) * effective_dim * m));
    AAADMM_CUDA_OK(cudaMalloc((void **)&a->dG, sizeof(double) * total_dim * m));
    AAADMM_CUDA_OK(cudaMalloc((void **)&a->partials, sizeof(double) * RED_MAX_Q * RED_MAX_BLOCKS));
    AAADMM_CUDA_OK(cudaMalloc((void **)&a->st, sizeof(SolveState)));
    AAADMM_CUDA_OK(cudaMemsetAsync(a->st, 0, sizeof(SolveState), a->stream));
    AAADMM_CUDA_OK(cudaMemsetAsync(a->dF, 0, sizeof(double) * effective_dim * m, a->stream));
    AAADMM_CUDA_OK(cudaMemsetAsync(a->dG, 0, sizeof(double) * total_dim * m, a->stream));
    k_init_state<<<1, 1, 0, a->stream>>>(a->st, 1, m, 0.0, 0);
    k_aa_set_counters<<<1, 1, 0, a->stream>>>(a->st, -1, -1);
    AAADMM_CUDA_OK(cudaStreamSynchronize(a->stream));
    *out = a;
    return 0;
    API_TRY_END
}

int aaadmm_aa_destroy(aaadmm_aa *a) {
    if (!a) return 0;
    cudaFree(a->ucur);
    cudaFree(a->gtmp);
    cudaFree(a->dF);
    cudaFree(a->dG);
    cudaFree(a->partials);
    cudaFree(a->st);
    if (a->stream) cudaStreamDestroy(a->stream);
    delete a;
    return 0;
}

static int aa_set_u(aaadmm_aa *a, const double *u, int64_t n, bool host, bool reset_counters) {
    if (n != a->Nt) {
        set_last_error("aa: vector size != total_dim");
        return -1;
    }
    AAADMM_CUDA_OK(cudaMemcpyAsync(a->ucur, u, sizeof(double) * n, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                   a->stream));
    if (reset_counters) k_aa_set_counters<<<1, 1, 0, a->stream>>>(a->st, 0, 0);
    AAADMM_CUDA_OK(cudaStreamSynchronize(a->stream));
    return 0;
}
int aaadmm_aa_init(aaadmm_aa *a, const double *u, int64_t n) { return aa_set_u(a, u, n, true, true); }
int aaadmm_aa_reset(aaadmm_aa *a, const double *u, int64_t n) { return aa_set_u(a, u, n, true, true); }
int aaadmm_aa_replace(aaadmm_aa *a, const double *u, int64_t n) { return aa_set_u(a, u, n, true, false); }
int aaadmm_aa_init_dev(aaadmm_aa *a, const double *u, int64_t n) { return aa_set_u(a, u, n, false, true); }
int aaadmm_aa_reset_dev(aaadmm_aa *a, const double *u, int64_t n) { return aa_set_u(a, u, n, false, true); }
int aaadmm_aa_replace_dev(aaadmm_aa *a, const double *u, int64_t n) { return aa_set_u(a, u, n, false, false); }

static int aa_compute_impl(aaadmm_aa *a, const double *d_g, double *accel, int64_t n, bool host_out) {
    if (n != a->Nt) {
        set_last_error("aa_compute: vector size != total_dim");
        return -1;
    }
    if (launch_aa_pass1(a->m, a->grid, a->stream, d_g, d_g + a->Ne, nullptr, a->ucur, a->dF, a->dG, a->Ne, a->Nt, a->st,
                        a->partials))
        return -1;
    if (launch_aa_pass2(a->m, a->grid, a->stream, d_g, d_g + a->Ne, a->ucur, a->dF, a->dG, a->Ne, a->Nt, a->st))
        return -1;
    AAADMM_CUDA_OK(cudaMemcpyAsync(accel, a->ucur, sizeof(double) * n,
                                   host_out ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, a->stream));
    AAADMM_CUDA_OK(cudaStreamSynchronize(a->stream));
    return 0;
}
int aaadmm_aa_compute(aaadmm_aa *a, const double *g, double *accel_u, int64_t n) {
    if (n != a->Nt) {
        set_last_error("aa_compute: vector size != total_dim");
        return -1;
    }
    AAADMM_CUDA_OK(cudaMemcpyAsync(a->gtmp, g, sizeof(double) * n, cudaMemcpyHostToDevice, a->stream));
    return aa_compute_impl(a, a->gtmp, accel_u, n, true);
}
int aaadmm_aa_compute_dev(aaadmm_aa *a, const double *d_g, double *d_accel_u, int64_t n) {
    return aa_compute_impl(a, d_g, d_accel_u, n, false);
}
int aaadmm_aa_state(aaadmm_aa *a, int *iter, int *col) {
    SolveState h;
    AAADMM_CUDA_OK(cudaMemcpy(&h, a->st, sizeof(SolveState), cudaMemcpyDeviceToHost));
    if (iter) *iter = h.aa_iter;
    if (col) *col = h.aa_col;
    return 0;
}

// ------------------------------------------------------------------------------------------
// LDL^T
// ------------------------------------------------------------------------------------------
int aaadmm_ldlt_create(aaadmm_ldlt **out, int n, const int64_t *Lp, const int *Li, const double *Lx,
                       const double *D, const int *perm, int nrhs) {
    API_TRY_BEGIN
    if (aaadmm_device_count() <= 0) {
        set_last_error("ldlt_create: no CUDA device (this library has no CPU path)");
        return -1;
    }
    if (n <= 0 || !Lp || !D || !perm) {
        set_last_error("ldlt_create: bad arguments");
        return -1;
    }
    aaadmm_ldlt *h = new aaadmm_ldlt();
    if (ldlt_dev_create(&h->f, n, Lp, Li, Lx, D, perm, nrhs)) {
        delete h;
        return -1;
    }
    AAADMM_CUDA_OK(cudaStreamCreate(&h->stream));
    AAADMM_CUDA_OK(cudaMalloc((void **)&h->b_dev, sizeof(double) * n * nrhs));
    AAADMM_CUDA_OK(cudaMalloc((void **)&h->x_dev, sizeof(double) * n * nrhs));
    *out = h;
    return 0;
    API_TRY_END
}
int aaadmm_ldlt_create_from_matrix(aaadmm_ldlt **out, int n, const int64_t *Ap, const int *Ai, const double *Ax,
                                   const int64_t *Lp, const int *Li, const int *perm, int nrhs) {
    API_TRY_BEGIN
    if (aaadmm_device_count() <= 0) {
        set_last_error("ldlt_create_from_matrix: no CUDA device (this library has no CPU path)");
        return -1;
    }
    if (n <= 0 || !Ap || !Ai || !Ax || !Lp || !perm || (Lp[n] > 0 && !Li)) {
        set_last_error("ldlt_create_from_matrix: bad arguments");
        return -1;
    }
    aaadmm_ldlt *h = new aaadmm_ldlt();
    if (ldlt_dev_create_from_matrix(&h->f, n, Ap, Ai, Ax, Lp, Li, perm, nrhs)) {
        delete h;
        return -1;
    }
    *out = h;  // from here on aaadmm_ldlt_destroy frees everything
    AAADMM_CUDA_OK(cudaStreamCreate(&h->stream));
    AAADMM_CUDA_OK(cudaMalloc((void **)&h->b_dev, sizeof(double) * n * nrhs));
    AAADMM_CUDA_OK(cudaMalloc((void **)&h->x_dev, sizeof(double) * n * nrhs));
    return 0;
    API_TRY_END
}
int aaadmm_ldlt_refactor(aaadmm_ldlt *h, const double *Ax) {
    API_TRY_BEGIN
    if (!h || !Ax) {
        set_last_error("ldlt_refactor: bad arguments");
        return -1;
    }
    return ldlt_dev_refactor(h->f, Ax, h->stream);
    API_TRY_END
}
int aaadmm_ldlt_destroy(aaadmm_ldlt *h) {
    if (!h) return 0;
    ldlt_dev_destroy(h->f);
    cudaFree(h->b_dev);
    cudaFree(h->x_dev);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}
int aaadmm_ldlt_solve_dev(aaadmm_ldlt *h, const double *d_b, double *d_x) {
    if (ldlt_dev_apply(h->f, d_b, d_x, h->stream, nullptr)) return -1;
    AAADMM_CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}
int aaadmm_ldlt_solve(aaadmm_ldlt *h, const double *b, double *x) {
    const size_t bytes = sizeof(double) * (size_t)h->f->n * h->f->nrhs;
    AAADMM_CUDA_OK(cudaMemcpyAsync(h->b_dev, b, bytes, cudaMemcpyHostToDevice, h->stream));
    if (ldlt_dev_apply(h->f, h->b_dev, h->x_dev, h->stream, nullptr)) return -1;
    AAADMM_CUDA_OK(cudaMemcpyAsync(x, h->x_dev, bytes, cudaMemcpyDeviceToHost, h->stream));
    AAADMM_CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}
int aaadmm_ldlt_dump_trace(aaadmm_ldlt *h, const char *path) { return ldlt_dev_dump_trace(h->f, path); }
int aaadmm_ldlt_stats(aaadmm_ldlt *h, double *s) {
    const LdltStats &t = h->f->stats;
    s[0] = t.n;
    s[1] = t.n_blocks;
    s[2] = t.n_levels;
    s[3] = t.max_block;
    s[4] = (double)t.nnz_L;
    s[5] = (double)t.nnz_offdiag;
    s[6] = (double)t.nnz_diag_dense;
    s[7] = t.bytes_per_solve;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Tet scene
// ------------------------------------------------------------------------------------------
int aaadmm_tetscene_destroy(aaadmm_tetscene *s) {
    if (!s) return 0;
    cudaFree(s->idx);
    cudaFree(s->binv);
    cudaFree(s->w);
    cudaFree(s->kvol);
    cudaFree(s->mass);
    cudaFree(s->material);
    cudaFree(s->hyper_ids);
    cudaFree(s->mu);
    cudaFree(s->lambda);
    cudaFree(s->volume);
    cudaFree(s->tri_idx);
    cudaFree(s->tri_rp);
    cudaFree(s->tri_w);
    cudaFree(s->tri_lmin);
    cudaFree(s->tri_lmax);
    cudaFree(s->pt_vert);
    cudaFree(s->obj_type);
    cudaFree(s->pt_w);
    cudaFree(s->obj_prm);
    cudaFree(s->inc_ptr);
    cudaFree(s->inc);
    cudaFree(s->Ubuf);
    cudaFree(s->Gbuf);
    cudaFree(s->xs);
    cudaFree(s->z);
    cudaFree(s->contrib);
    cudaFree(s->bconst);
    cudaFree(s->xbar);
    cudaFree(s->xpin);
    cudaFree(s->dF);
    cudaFree(s->dG);
    cudaFree(s->xz_a);
    cudaFree(s->xz_b);
    cudaFree(s->hist_prim);
    cudaFree(s->hist_comb);
    cudaFree(s->hist_rej);
    cudaFree(s->partials);
    cudaFree(s->st);
    cudaFree(s->zu_scratch);
    if (s->st_h) cudaFreeHost(s->st_h);
    if (s->pin_h) cudaFreeHost(s->pin_h);
    if (s->xbar_h) cudaFreeHost(s->xbar_h);
    if (s->xout_h) cudaFreeHost(s->xout_h);
    for (auto &e : s->ev)
        if (e) cudaEventDestroy(e);
    if (s->loop_exec) cudaGraphExecDestroy(s->loop_exec);
    if (s->loop_graph) cudaGraphDestroy(s->loop_graph);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return 0;
}

int aaadmm_tetscene_create(aaadmm_tetscene **out, const aaadmm_tetscene_desc *d, aaadmm_ldlt *factor) {
    API_TRY_BEGIN
    if (aaadmm_device_count() <= 0) {
        set_last_error("tetscene_create: no CUDA device (this library has no CPU path)");
        return -1;
    }
    if (!d || !factor || d->n_tets < 0 || d->n_tris < 0 || d->n_tets + d->n_tris <= 0 || d->n_free <= 0 ||
        d->n_free > d->n_verts || d->n_collisions < 0 || d->n_obstacles < 0) {
        set_last_error("tetscene_create: bad arguments");
        return -1;
    }
    if (d->n_tris > 0 && (!d->tri || !d->tri_rest_pose || !d->tri_weight)) {
        set_last_error("tetscene_create: triangle terms need tri, tri_rest_pose and tri_weight");
        return -1;
    }
    // Ahat (x) I3 as a scalar n_free factor with 3 right-hand sides (this repo's setup), or a factor of the full
    // 3 n_free system in the reference's degree-of-freedom order 3*vertex + axis with one right-hand side (what
    // Eigen::SimplicialLDLT gives the reference, LinearSolver.hpp:79-84)
    if (!((factor->f->n == d->n_free && factor->f->nrhs == 3) || (factor->f->n == 3 * d->n_free && factor->f->nrhs == 1))) {
        set_last_error("tetscene_create: factor must be n_free x n_free with nrhs = 3, or 3 n_free x 3 n_free with nrhs = 1");
        return -1;
    }
    if (d->n_collisions > 0) {
        if (!d->collision_vert || !d->collision_weight || (d->n_obstacles > 0 && (!d->obstacle_type || !d->obstacle_prm))) {
            set_last_error("tetscene_create: collision terms need collision_vert, collision_weight and the obstacle arrays");
            return -1;
        }
        for (int i = 0; i < d->n_collisions; ++i)
            if (d->collision_vert[i] < 0 || d->collision_vert[i] >= d->n_free) {
                set_last_error("tetscene_create: collision terms must sit on free vertices");
                return -1;
            }
        for (int j = 0; j < d->n_obstacles; ++j)
            if (d->obstacle_type[j] < AAADMM_PASSIVE_FLOOR || d->obstacle_type[j] > AAADMM_PASSIVE_CYLINDER) {
                set_last_error("tetscene_create: unknown obstacle type");
                return -1;
            }
    }
    std::vector<int> hyper_ids;
    if (d->material) {
        for (int t = 0; t < d->n_tets; ++t) {
            if (d->material[t] < 0 || d->material[t] > 2) {
                set_last_error("tetscene_create: material must be 0 (linear), 1 (neo-hookean) or 2 (stvk)");
                return -1;
            }
            if (d->material[t] != 0) hyper_ids.push_back(t);
        }
        if (!hyper_ids.empty() && (!d->mu || !d->lambda || !d->volume)) {
            set_last_error("tetscene_create: hyper-elastic tets need mu, lambda and volume");
            return -1;
        }
    }
    aaadmm_tetscene *s = new aaadmm_tetscene();
    auto build = [&]() -> int {
    const int T = d->n_tets, V = d->n_verts, NF = d->n_free;
    s->T = T;
    s->V = V;
    s->NF = NF;
    s->NP = V - NF;
    s->rho_dt2 = d->rho_dt2;
    s->factor = factor;
    const int NT = d->n_tris;
    s->NT = NT;
    const int NC = d->n_collisions;
    s->NC = NC;
    s->n_objs = d->n_obstacles;
    s->Ne = 9 * (int64_t)T + 6 * (int64_t)NT + 3 * (int64_t)NC;
    s->Nt = s->Ne + 3 * (int64_t)NF;
    s->Nbuf = s->Ne + 3 * (int64_t)V;
    AAADMM_CUDA_OK(cudaStreamCreate(&s->stream));
    for (auto &e : s->ev) AAADMM_CUDA_OK(cudaEventCreate(&e));
    // constants: AoS -> SoA on the host once
    std::vector<double> binv_soa((size_t)9 * T);
    for (int t = 0; t < T; ++t)
        for (int k = 0; k < 9; ++k) binv_soa[(size_t)k * T + t] = d->binv[9 * (size_t)t + k];
    if (NT > 0) {
        std::vector<int> idx4((size_t)4 * NT, 0);
        std::vector<double> rp_soa((size_t)4 * NT), lmin(NT, -100.0), lmax(NT, 100.0);
        for (int t = 0; t < NT; ++t) {
            for (int c = 0; c < 3; ++c) idx4[4 * (size_t)t + c] = d->tri[3 * (size_t)t + c];
            for (int k = 0; k < 4; ++k) rp_soa[(size_t)k * NT + t] = d->tri_rest_pose[4 * (size_t)t + k];
            if (d->tri_limit_min) lmin[t] = d->tri_limit_min[t];
            if (d->tri_limit_max) lmax[t] = d->tri_limit_max[t];
        }
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->tri_idx, sizeof(int4) * NT));
        AAADMM_CUDA_OK(cudaMemcpy(s->tri_idx, idx4.data(), sizeof(int4) * NT, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->tri_rp, sizeof(double) * 4 * NT));
        AAADMM_CUDA_OK(cudaMemcpy(s->tri_rp, rp_soa.data(), sizeof(double) * 4 * NT, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->tri_w, sizeof(double) * NT));
        AAADMM_CUDA_OK(cudaMemcpy(s->tri_w, d->tri_weight, sizeof(double) * NT, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->tri_lmin, sizeof(double) * NT));
        AAADMM_CUDA_OK(cudaMemcpy(s->tri_lmin, lmin.data(), sizeof(double) * NT, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->tri_lmax, sizeof(double) * NT));
        AAADMM_CUDA_OK(cudaMemcpy(s->tri_lmax, lmax.data(), sizeof(double) * NT, cudaMemcpyHostToDevice));
    }
    if (NC > 0) {
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->pt_vert, sizeof(int) * NC));
        AAADMM_CUDA_OK(cudaMemcpy(s->pt_vert, d->collision_vert, sizeof(int) * NC, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->pt_w, sizeof(double) * NC));
        AAADMM_CUDA_OK(cudaMemcpy(s->pt_w, d->collision_weight, sizeof(double) * NC, cudaMemcpyHostToDevice));
        const int no = std::max(1, d->n_obstacles);
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->obj_type, sizeof(int) * no));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->obj_prm, sizeof(double) * 7 * no));
        if (d->n_obstacles > 0) {
            AAADMM_CUDA_OK(cudaMemcpy(s->obj_type, d->obstacle_type, sizeof(int) * d->n_obstacles, cudaMemcpyHostToDevice));
            AAADMM_CUDA_OK(cudaMemcpy(s->obj_prm, d->obstacle_prm, sizeof(double) * 7 * d->n_obstacles, cudaMemcpyHostToDevice));
        }
    }
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->idx, sizeof(int4) * std::max(T, 1)));
    AAADMM_CUDA_OK(cudaMemcpy(s->idx, d->tet, sizeof(int) * 4 * T, cudaMemcpyHostToDevice));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->binv, sizeof(double) * 9 * std::max(T, 1)));
    AAADMM_CUDA_OK(cudaMemcpy(s->binv, binv_soa.data(), sizeof(double) * 9 * T, cudaMemcpyHostToDevice));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->w, sizeof(double) * std::max(T, 1)));
    AAADMM_CUDA_OK(cudaMemcpy(s->w, d->weight, sizeof(double) * T, cudaMemcpyHostToDevice));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->kvol, sizeof(double) * std::max(T, 1)));
    AAADMM_CUDA_OK(cudaMemcpy(s->kvol, d->kvol, sizeof(double) * T, cudaMemcpyHostToDevice));
    s->n_hyper = (int)hyper_ids.size();
    if (s->n_hyper > 0) {
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->material, sizeof(int) * T));
        AAADMM_CUDA_OK(cudaMemcpy(s->material, d->material, sizeof(int) * T, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->hyper_ids, sizeof(int) * s->n_hyper));
        AAADMM_CUDA_OK(cudaMemcpy(s->hyper_ids, hyper_ids.data(), sizeof(int) * s->n_hyper, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->mu, sizeof(double) * T));
        AAADMM_CUDA_OK(cudaMemcpy(s->mu, d->mu, sizeof(double) * T, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->lambda, sizeof(double) * T));
        AAADMM_CUDA_OK(cudaMemcpy(s->lambda, d->lambda, sizeof(double) * T, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->volume, sizeof(double) * T));
        AAADMM_CUDA_OK(cudaMemcpy(s->volume, d->volume, sizeof(double) * T, cudaMemcpyHostToDevice));
    }
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->mass, sizeof(double) * NF));
    AAADMM_CUDA_OK(cudaMemcpy(s->mass, d->mass_free, sizeof(double) * NF, cudaMemcpyHostToDevice));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->inc_ptr, sizeof(int64_t) * (NF + 1)));
    AAADMM_CUDA_OK(cudaMemcpy(s->inc_ptr, d->inc_ptr, sizeof(int64_t) * (NF + 1), cudaMemcpyHostToDevice));
    const int64_t ninc = d->inc_ptr[NF];
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->inc, sizeof(int) * std::max<int64_t>(ninc, 1)));
    AAADMM_CUDA_OK(cudaMemcpy(s->inc, d->inc, sizeof(int) * ninc, cudaMemcpyHostToDevice));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->Ubuf, sizeof(double) * s->Nbuf));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->Gbuf, sizeof(double) * s->Nbuf));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->xs, sizeof(double) * 3 * V));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->z, sizeof(double) * s->Ne));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->contrib, sizeof(double) * (12 * (size_t)T + 9 * (size_t)NT + 3 * (size_t)NC)));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->bconst, sizeof(double) * 3 * NF));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->xbar, sizeof(double) * 3 * NF));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->xpin, sizeof(double) * 3 * std::max(1, s->NP)));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->partials, sizeof(double) * RED_MAX_Q * RED_MAX_BLOCKS));
    AAADMM_CUDA_OK(cudaMalloc((void **)&s->st, sizeof(SolveState)));
    AAADMM_CUDA_OK(cudaMemset(s->st, 0, sizeof(SolveState)));
    AAADMM_CUDA_OK(cudaMallocHost((void **)&s->pin_h, sizeof(double) * 3 * std::max(1, s->NP)));
    AAADMM_CUDA_OK(cudaMallocHost((void **)&s->xbar_h, sizeof(double) * 3 * NF));
    AAADMM_CUDA_OK(cudaMallocHost((void **)&s->xout_h, sizeof(double) * 3 * NF));
    AAADMM_CUDA_OK(cudaMallocHost((void **)&s->st_h, sizeof(SolveState)));
    return 0;
    };
    if (build() != 0) {  // nothing of a partially built scene stays behind
        aaadmm_tetscene_destroy(s);
        return -1;
    }
    *out = s;
    return 0;
    API_TRY_END
}

int aaadmm_tetscene_update_material(aaadmm_tetscene *s, const double *weight, const double *kvol, const double *mu,
                                    const double *lambda, const double *tri_weight, const double *tri_limit_min,
                                    const double *tri_limit_max, double rho_dt2) {
    API_TRY_BEGIN
    if (!s || (s->T > 0 && (!weight || !kvol)) || (s->NT > 0 && !tri_weight) || (s->n_hyper > 0 && (!mu || !lambda))) {
        set_last_error("tetscene_update_material: null array");
        return -1;
    }
    cudaStream_t st = s->stream;
    if (s->T > 0) {
        AAADMM_CUDA_OK(cudaMemcpyAsync(s->w, weight, sizeof(double) * s->T, cudaMemcpyHostToDevice, st));
        AAADMM_CUDA_OK(cudaMemcpyAsync(s->kvol, kvol, sizeof(double) * s->T, cudaMemcpyHostToDevice, st));
        if (s->n_hyper > 0) {
            AAADMM_CUDA_OK(cudaMemcpyAsync(s->mu, mu, sizeof(double) * s->T, cudaMemcpyHostToDevice, st));
            AAADMM_CUDA_OK(cudaMemcpyAsync(s->lambda, lambda, sizeof(double) * s->T, cudaMemcpyHostToDevice, st));
        }
    }
    if (s->NT > 0) {
        AAADMM_CUDA_OK(cudaMemcpyAsync(s->tri_w, tri_weight, sizeof(double) * s->NT, cudaMemcpyHostToDevice, st));
        if (tri_limit_min) AAADMM_CUDA_OK(cudaMemcpyAsync(s->tri_lmin, tri_limit_min, sizeof(double) * s->NT, cudaMemcpyHostToDevice, st));
        if (tri_limit_max) AAADMM_CUDA_OK(cudaMemcpyAsync(s->tri_lmax, tri_limit_max, sizeof(double) * s->NT, cudaMemcpyHostToDevice, st));
    }
    if (rho_dt2 != s->rho_dt2) {
        s->rho_dt2 = rho_dt2;
        s->loop_key = -1;  // the captured loop body holds rho dt^2 by value: capture again at the next step
    }
    s->has_inputs = false;
    AAADMM_CUDA_OK(cudaStreamSynchronize(st));  // the host arrays may go away
    return 0;
    API_TRY_END
}

static int scene_reserve(aaadmm_tetscene *s, int iters, int m) {
    if (iters > s->hist_cap) {
        cudaFree(s->hist_prim);
        cudaFree(s->hist_comb);
        cudaFree(s->hist_rej);
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->hist_prim, sizeof(double) * iters));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->hist_comb, sizeof(double) * 2 * iters));  // residuals | time stamps
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->hist_rej, sizeof(int) * iters));
        s->hist_cap = iters;
        s->loop_key = -1;
    }
    if (m > s->m_cap) {
        cudaFree(s->dF);
        cudaFree(s->dG);
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->dF, sizeof(double) * s->Ne * m));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->dG, sizeof(double) * s->Nt * m));
        s->m_cap = m;
        s->loop_key = -1;
    }
    return 0;
}

namespace {
struct PhaseProf {
    bool on = false;
    std::vector<cudaEvent_t> ev;  // start/stop pairs
    std::vector<int> phase;
    cudaStream_t s;
    void begin(int p) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        ev.push_back(e);
        phase.push_back(p);
    }
    void end() {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        ev.push_back(e);
    }
};
}  // namespace

// Graph WHILE-loop plumbing shared by the two orderings. begin: returns 1 if a cached graph for
// `key` was launched, 0 if capture of one loop turn has started, -1 on error.
static int loop_graph_begin(aaadmm_tetscene *s, int key, cudaGraphConditionalHandle *cond_handle) {
    cudaStream_t st = s->stream;
    if (s->loop_key == key && s->loop_exec) {
        AAADMM_CUDA_OK(cudaGraphLaunch(s->loop_exec, st));
        return 1;
    }
    if (s->loop_exec) cudaGraphExecDestroy(s->loop_exec), s->loop_exec = nullptr;
    if (s->loop_graph) cudaGraphDestroy(s->loop_graph), s->loop_graph = nullptr;
    AAADMM_CUDA_OK(cudaGraphCreate(&s->loop_graph, 0));
    AAADMM_CUDA_OK(cudaGraphConditionalHandleCreate(cond_handle, s->loop_graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams np = {};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = *cond_handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    AAADMM_CUDA_OK(cudaGraphAddNode(&node, s->loop_graph, nullptr, 0, &np));
    AAADMM_CUDA_OK(cudaStreamBeginCaptureToGraph(st, np.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                                 cudaStreamCaptureModeThreadLocal));
    s->loop_key = key;
    return 0;
}
// A launch failed while one loop turn was being captured: leave capture mode and forget the half-built graph, so that
// the stream is usable and the next step() starts a fresh capture. Returns -1 (the caller's error code).
static int loop_graph_abort(aaadmm_tetscene *s) {
    cudaGraph_t dummy = nullptr;
    cudaStreamEndCapture(s->stream, &dummy);  // the body graph is owned by the conditional node
    cudaGetLastError();
    if (s->loop_exec) cudaGraphExecDestroy(s->loop_exec), s->loop_exec = nullptr;
    if (s->loop_graph) cudaGraphDestroy(s->loop_graph), s->loop_graph = nullptr;
    s->loop_key = -1;
    return -1;
}
static int loop_graph_end(aaadmm_tetscene *s, cudaGraphConditionalHandle cond_handle, int &L, int L_before) {
    cudaStream_t st = s->stream;
    k_loop_cond<<<1, 1, 0, st>>>(cond_handle, s->st);
    s->body_launches = L - L_before + 1;
    L = L_before;
    cudaError_t e = cudaStreamEndCapture(st, nullptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (s->loop_graph) cudaGraphDestroy(s->loop_graph), s->loop_graph = nullptr;
        s->loop_key = -1;
        set_last_error(std::string("loop graph capture failed: ") + cudaGetErrorString(e));
        return -1;
    }
    e = cudaGraphInstantiate(&s->loop_exec, s->loop_graph, 0);
    if (e != cudaSuccess) {
        s->loop_key = -1;
        set_last_error(std::string("loop graph instantiate failed: ") + cudaGetErrorString(e));
        return -1;
    }
    AAADMM_CUDA_OK(cudaGraphLaunch(s->loop_exec, st));
    return 0;
}

// hard_zxu ordering: hard/src/Solver.cpp:74-214 from "Initialize ADMM vars" to the end of the loop.
static int run_hard(aaadmm_tetscene *s, const aaadmm_step_opts *o, PhaseProf *prof, int iters, bool init_frame,
                    bool use_graph = false) {
    cudaStream_t st = s->stream;
    LdltDev *f = s->factor->f;
    const int T = s->T, NF = s->NF;
    const bool accel = o->accel && o->anderson_m > 0;
    const int m = accel ? o->anderson_m : 1;
    TetArrays A{T, NF, s->V, s->idx, s->binv, s->w, s->kvol, s->rho_dt2, s->material, s->mu, s->lambda, s->volume, s->hyper_ids, s->n_hyper};
    const int gt = std::max(1, std::min((T + TET_BLOCK - 1) / TET_BLOCK, stream_grid(8)));
    const int gv = (NF + 127) / 128;
    const int gs = stream_grid(4);
    double *Uu = s->Ubuf, *Ux = s->Ubuf + s->Ne;
    double *Gu = s->Gbuf, *Gx = s->Gbuf + s->Ne;
    int &L = s->launches;
    PhaseProf none;
    if (!prof) prof = &none;
    // triangle terms: their u / z planes and contribution slots follow the tets'
    const int NT = s->NT;
    const size_t o9 = 9 * (size_t)T;
    TriArrays R{NT, NF, s->tri_idx, s->tri_rp, s->tri_w, s->tri_lmin, s->tri_lmax, s->rho_dt2};
    double *tri_contrib = s->contrib + 12 * (size_t)T;
    // collision terms behind the triangles
    const int NC = s->NC;
    const size_t o15 = o9 + 6 * (size_t)NT;
    PointArrays Pt{NC, s->pt_vert, s->pt_w, s->n_objs, s->obj_type, s->obj_prm, s->rho_dt2};
    double *pt_contrib = tri_contrib + 9 * (size_t)NT;
    auto update_z = [&](int mode) {
        if (NC > 0) {
            launch_pt_update_z_hard(mode, st, Pt, Ux, Uu + o15, s->z + o15, pt_contrib, s->st, s->partials);
            ++L;
        }
        if (NT > 0) {
            launch_tri_update_z_hard(mode, st, R, Ux, Uu + o9, s->z + o9, tri_contrib, s->st, s->partials);
            ++L;
        }
        launch_update_z_hard(mode, gt, st, A, Ux, Uu, s->z, s->contrib, s->st, s->partials);
    };
    auto update_u = [&](int mode, double *u_out) {
        if (NC > 0) {
            launch_pt_update_u_hard(mode, st, Pt, s->xs, Ux, s->z + o15, Uu + o15, u_out + o15, s->st, s->partials);
            ++L;
        }
        if (NT > 0) {
            launch_tri_update_u_hard(mode, st, R, s->xs, Ux, s->z + o9, Uu + o9, u_out + o9, s->st, s->partials);
            ++L;
        }
        launch_update_u_hard(mode, gt, st, A, s->xs, Ux, s->z, Uu, u_out, s->st, s->partials, s->hist_prim,
                             s->hist_comb, s->hist_rej);
    };

    if (init_frame) {
        k_init_state<<<1, 1, 0, st>>>(s->st, accel ? 1 : 0, m, o->eps, iters);
        // pinned tails of the three position arrays; x = x_bar; u = 0
        if (s->NP > 0) {
            const size_t pb = sizeof(double) * 3 * s->NP;
            AAADMM_CUDA_OK(cudaMemcpyAsync(Ux + 3 * (size_t)NF, s->xpin, pb, cudaMemcpyDeviceToDevice, st));
            AAADMM_CUDA_OK(cudaMemcpyAsync(Gx + 3 * (size_t)NF, s->xpin, pb, cudaMemcpyDeviceToDevice, st));
            AAADMM_CUDA_OK(cudaMemcpyAsync(s->xs + 3 * (size_t)NF, s->xpin, pb, cudaMemcpyDeviceToDevice, st));
        }
        AAADMM_CUDA_OK(cudaMemcpyAsync(Ux, s->xbar, sizeof(double) * 3 * NF, cudaMemcpyDeviceToDevice, st));
        AAADMM_CUDA_OK(cudaMemsetAsync(Uu, 0, sizeof(double) * s->Ne, st));
        launch_bconst(st, A, s->inc_ptr, s->inc, Ux, s->mass, s->xbar, s->bconst);
        if (NT > 0) {
            launch_tri_bconst(st, R, 4 * T, s->inc_ptr, s->inc, Ux, s->bconst);
            ++L;
        }
        // warm start (hard/src/Solver.cpp:99-114)
        update_z(MODE_WARM);
        launch_rhs_gather(st, NF, s->inc_ptr, s->inc, s->contrib, s->bconst, f->iperm, f->W, s->st, 0, f->nrhs == 1);
        if (ldlt_dev_apply_permuted(f, s->xs, st, &s->st->done)) return -1;
        update_u(MODE_WARM, Gu);
        AAADMM_CUDA_OK(cudaMemcpyAsync(Gx, s->xs, sizeof(double) * 3 * NF, cudaMemcpyDeviceToDevice, st));
        // default_(u,x) = curr_(u,x); accelerator->init(curr_u, curr_x)
        AAADMM_CUDA_OK(cudaMemcpyAsync(s->Ubuf, s->Gbuf, sizeof(double) * s->Nt, cudaMemcpyDeviceToDevice, st));
        L += 6 + f->n_launches;
        k_mark_loop_start<<<1, 1, 0, st>>>(s->st);
        ++L;
    }
    // ---- the loop: one graph launch (WHILE node, device-side break), or plain launches when profiling ----
    cudaGraphConditionalHandle cond_handle = 0;
    bool capturing = false;
    if (use_graph && iters > 0) {
        const int rc = loop_graph_begin(s, (accel ? 64 : 0) + m, &cond_handle);
        if (rc < 0) return -1;
        if (rc == 1) return 0;  // cached graph launched
        capturing = true;
        iters = 1;
    }
    const int L_before = L;
    for (int it = 0; it < iters; ++it) {
        prof->begin(0);
        update_z(MODE_ITER);
        prof->end();
        ++L;
        if (accel) {
            prof->begin(6);
            launch_restore_if_reject(gs, st, s->Ubuf, s->Gbuf, s->Nt, s->st);
            update_z(MODE_REDO);
            prof->end();
            L += 2;
        }
        prof->begin(1);
        launch_rhs_gather(st, NF, s->inc_ptr, s->inc, s->contrib, s->bconst, f->iperm, f->W, s->st, 0, f->nrhs == 1);
        prof->end();
        prof->begin(2);
        if (ldlt_dev_apply_permuted(f, s->xs, st, &s->st->done)) return capturing ? loop_graph_abort(s) : -1;
        prof->end();
        L += 1 + f->n_launches;
        prof->begin(3);
        update_u(MODE_ITER, accel ? Gu : Uu);
        prof->end();
        ++L;
        if (accel) {
            prof->begin(4);
            if (launch_aa_pass1(m, gs, st, Gu, s->xs, Gx, s->Ubuf, s->dF, s->dG, s->Ne, s->Nt, s->st, s->partials)) return capturing ? loop_graph_abort(s) : -1;
            prof->end();
            prof->begin(5);
            if (launch_aa_pass2(m, gs, st, Gu, Gx, s->Ubuf, s->dF, s->dG, s->Ne, s->Nt, s->st)) return capturing ? loop_graph_abort(s) : -1;
            prof->end();
            L += 2;
        } else {
            prof->begin(4);
            launch_copy_if_not_done(st, Ux, s->xs, 3 * (int64_t)NF, s->st);
            prof->end();
            ++L;
        }
    }
    if (capturing) return loop_graph_end(s, cond_handle, L, L_before);
    AAADMM_CUDA_OK(cudaGetLastError());
    return 0;
}

// xzu ordering: admm_anderson_xzu/src/Solver.cpp:78-257. Anderson variable = z (Ne = Nt = 9T).
//   Zcur = Ubuf[0:9T] (accelerator's current_u_ = curr_z), Zdef = Gbuf[0:9T] (default_z),
//   curr_x = xs, default_x = Gbuf x-part, comb_x = Ubuf x-part, curr_u = s->z, default_u = xz_a,
//   comb_z / last_z = xz_b.
static int run_xzu(aaadmm_tetscene *s, const aaadmm_step_opts *o, int iters, bool use_graph) {
    cudaStream_t st = s->stream;
    LdltDev *f = s->factor->f;
    const int T = s->T, NF = s->NF;
    const bool accel = o->accel && o->anderson_m > 0;
    const int m = o->anderson_m > 0 ? o->anderson_m : 1;
    TetArrays A{T, NF, s->V, s->idx, s->binv, s->w, s->kvol, s->rho_dt2, s->material, s->mu, s->lambda, s->volume, s->hyper_ids, s->n_hyper};
    const int gt = std::min((T + TET_BLOCK - 1) / TET_BLOCK, stream_grid(8));
    const int gs = stream_grid(4);
    const int64_t NZ = s->Ne, NX = 3 * (int64_t)NF;
    double *Zcur = s->Ubuf, *Zdef = s->Gbuf, *cx = s->xs, *dx = s->Gbuf + s->Ne, *combx = s->Ubuf + s->Ne;
    double *u = s->z, *du = s->xz_a, *combz = s->xz_b;
    int &L = s->launches;
    auto solve = [&](const double *zz, double *xout, int when) -> int {
        launch_contrib(gt, st, A, zz, u, s->contrib, s->st, when);
        launch_rhs_gather(st, NF, s->inc_ptr, s->inc, s->contrib, s->bconst, f->iperm, f->W, s->st, when, f->nrhs == 1);
        if (ldlt_dev_apply_permuted(f, xout, st, when ? &s->st->skip_redo : &s->st->done)) return -1;
        L += 2 + f->n_launches;
        return 0;
    };
    // ---- frame init + warm start (Solver.cpp:78-117) ----
    k_init_state<<<1, 1, 0, st>>>(s->st, accel ? 1 : 0, m, o->eps, iters);
    if (s->NP > 0) {
        const size_t pb = sizeof(double) * 3 * s->NP;
        AAADMM_CUDA_OK(cudaMemcpyAsync(cx + NX, s->xpin, pb, cudaMemcpyDeviceToDevice, st));
        AAADMM_CUDA_OK(cudaMemcpyAsync(dx + NX, s->xpin, pb, cudaMemcpyDeviceToDevice, st));
        AAADMM_CUDA_OK(cudaMemcpyAsync(combx + NX, s->xpin, pb, cudaMemcpyDeviceToDevice, st));
    }
    AAADMM_CUDA_OK(cudaMemcpyAsync(cx, s->xbar, sizeof(double) * NX, cudaMemcpyDeviceToDevice, st));
    AAADMM_CUDA_OK(cudaMemsetAsync(u, 0, sizeof(double) * NZ, st));
    launch_bconst(st, A, s->inc_ptr, s->inc, cx, s->mass, s->xbar, s->bconst);
    launch_z_from_x(gt, st, A, cx, Zcur);
    if (solve(Zcur, cx, 0)) return -1;
    launch_update_z_plain(gt, st, A, cx, u, Zcur, s->st);
    // default_(z,x,u) = curr; accelerator.init(m, z_size, curr_z)
    AAADMM_CUDA_OK(cudaMemcpyAsync(Zdef, Zcur, sizeof(double) * NZ, cudaMemcpyDeviceToDevice, st));
    AAADMM_CUDA_OK(cudaMemcpyAsync(dx, cx, sizeof(double) * NX, cudaMemcpyDeviceToDevice, st));
    AAADMM_CUDA_OK(cudaMemsetAsync(du, 0, sizeof(double) * NZ, st));
    k_mark_loop_start<<<1, 1, 0, st>>>(s->st);
    L += 6;

    cudaGraphConditionalHandle cond_handle = 0;
    bool capturing = false;
    if (use_graph && iters > 0) {
        const int rc = loop_graph_begin(s, 1024 + (accel ? 64 : 0) + m, &cond_handle);
        if (rc < 0) return -1;
        if (rc == 1) return 0;
        capturing = true;
        iters = 1;
    }
    const int L_before = L;
    for (int it = 0; it < iters; ++it) {
        if (accel) {
            launch_grad_u_xzu(gt, st, A, Zcur, u, s->st);                                  // :125-133
        } else {
            launch_update_u_plain(gt, st, A, cx, Zcur, u, s->st, 0);                       // :135-141
        }
        if (solve(Zcur, cx, 0)) return capturing ? loop_graph_abort(s) : -1;                                                 // :147-149
        launch_prim_xzu(MODE_ITER, gt, st, A, cx, Zcur, s->st, s->partials);               // :153-159
        L += 2;
        if (accel) {
            launch_restore_xzu(gs, st, u, du, Zcur, Zdef, NZ, cx, dx, NX, s->st);          // :162-166
            launch_update_u_plain(gt, st, A, cx, Zcur, u, s->st, 1);                       // :168-172
            if (solve(Zcur, cx, 1)) return capturing ? loop_graph_abort(s) : -1;                                             // :174-176
            launch_prim_xzu(MODE_REDO, gt, st, A, cx, Zcur, s->st, s->partials);           // :178-180
            // default_x = curr_x; default_u = curr_u; default_z = update_z(curr_x, curr_u)   :192-201
            launch_copy2_if_not_done(gs, st, dx, cx, NX, du, u, NZ, s->st);
            launch_update_z_plain(gt, st, A, cx, u, Zdef, s->st);
            if (launch_aa_pass1(m, gs, st, Zdef, nullptr, nullptr, Zcur, s->dF, s->dG, NZ, NZ, s->st, s->partials)) return capturing ? loop_graph_abort(s) : -1;
            if (launch_aa_pass2(m, gs, st, Zdef, nullptr, Zcur, s->dF, s->dG, NZ, NZ, s->st)) return capturing ? loop_graph_abort(s) : -1;
            // combined residual "for drawing figures": extra solve + local step on copies     :217-233
            if (solve(Zdef, combx, 0)) return capturing ? loop_graph_abort(s) : -1;
            launch_update_z_plain(gt, st, A, combx, u, combz, s->st);
            launch_comb_xzu(gt, st, A, combx, combz, Zdef, s->st, s->partials, s->hist_prim, s->hist_comb, s->hist_rej);
            L += 10;
        } else {
            // last_z = curr_z; curr_z = update_z(curr_x, curr_u)                               :205-213, :234-238
            launch_copy2_if_not_done(gs, st, combz, Zcur, NZ, nullptr, nullptr, 0, s->st);
            launch_update_z_plain(gt, st, A, cx, u, Zcur, s->st);
            launch_comb_xzu(gt, st, A, cx, Zcur, combz, s->st, s->partials, s->hist_prim, s->hist_comb, s->hist_rej);
            L += 3;
        }
    }
    if (capturing) return loop_graph_end(s, cond_handle, L, L_before);
    AAADMM_CUDA_OK(cudaGetLastError());
    return 0;
}

static int step_common(aaadmm_tetscene *s, const aaadmm_step_opts *o, bool host_io, const double *x_bar,
                       const double *x_pin, double *x_out, double *hp, double *hc, int *hr, aaadmm_step_result *res) {
    if (!o || o->admm_iters < 0 || (o->accel && (o->anderson_m <= 0 || o->anderson_m > AA_MAX_M))) {
        set_last_error("tetscene_step: bad options (Anderson_m must be in 1..16 when accel is on)");
        return -1;
    }
    if (o->ordering != AAADMM_ORDER_HARD_ZXU && o->ordering != AAADMM_ORDER_XZU) {
        set_last_error("tetscene_step: unknown ordering");
        return -1;
    }
    const bool xzu = o->ordering == AAADMM_ORDER_XZU;
    if (xzu && (s->NT > 0 || s->NC > 0)) {
        set_last_error("tetscene_step: triangle and collision terms run under the hard_zxu ordering only");
        return -1;
    }
    if (xzu && !s->xz_a) {
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->xz_a, sizeof(double) * s->Ne));
        AAADMM_CUDA_OK(cudaMalloc((void **)&s->xz_b, sizeof(double) * s->Ne));
    }
    if (!host_io && !s->has_inputs) {
        set_last_error("tetscene_step_resident: call aaadmm_tetscene_step once first");
        return -1;
    }
    const bool accel = o->accel && o->anderson_m > 0;
    if (scene_reserve(s, std::max(1, o->admm_iters), (accel || xzu) ? std::max(1, o->anderson_m) : 1)) return -1;
    cudaStream_t st = s->stream;
    const int NF = s->NF;
    s->launches = 0;
    AAADMM_CUDA_OK(cudaEventRecord(s->ev[0], st));
    if (host_io) {
        memcpy(s->xbar_h, x_bar, sizeof(double) * 3 * NF);
        if (s->NP > 0) memcpy(s->pin_h, x_pin, sizeof(double) * 3 * s->NP);
        AAADMM_CUDA_OK(cudaMemcpyAsync(s->xbar, s->xbar_h, sizeof(double) * 3 * NF, cudaMemcpyHostToDevice, st));
        if (s->NP > 0)
            AAADMM_CUDA_OK(cudaMemcpyAsync(s->xpin, s->pin_h, sizeof(double) * 3 * s->NP, cudaMemcpyHostToDevice, st));
        s->has_inputs = true;
    }
    AAADMM_CUDA_OK(cudaEventRecord(s->ev[1], st));
    // AAADMM_NO_GRAPH=1: plain stream launches instead of the graph WHILE node (for profilers that
    // want every kernel as its own launch); same kernels, same order.
    static const bool no_graph = getenv("AAADMM_NO_GRAPH") != nullptr;
    if (xzu ? run_xzu(s, o, o->admm_iters, !no_graph) : run_hard(s, o, nullptr, o->admm_iters, true, !no_graph)) return -1;
    AAADMM_CUDA_OK(cudaEventRecord(s->ev[2], st));
    // which buffers hold the tets' u and z now (hard: default_u = Gbuf when accelerated, Ubuf otherwise; xzu: curr_u
    // lives in s->z and curr_z in Ubuf)
    s->last_u = xzu ? s->z : (accel ? s->Gbuf : s->Ubuf);
    s->last_z = xzu ? s->Ubuf : s->z;
    if (host_io) {
        // hard: default_x when ANDERSON, curr_x otherwise (hard/src/Solver.cpp:216-223)
        // xzu: curr_x (xzu/src/Solver.cpp:255)
        const double *xfinal = xzu ? s->xs : (accel ? (s->Gbuf + s->Ne) : s->xs);
        AAADMM_CUDA_OK(cudaMemcpyAsync(s->xout_h, xfinal, sizeof(double) * 3 * NF, cudaMemcpyDeviceToHost, st));
    }
    AAADMM_CUDA_OK(cudaMemcpyAsync(s->st_h, s->st, sizeof(SolveState), cudaMemcpyDeviceToHost, st));
    AAADMM_CUDA_OK(cudaEventRecord(s->ev[3], st));
    AAADMM_CUDA_OK(cudaStreamSynchronize(st));
    const SolveState hs = *s->st_h;
    s->last_rows = hs.iter;
    s->last_max_iters = hs.max_iters;
    if (host_io) {
        memcpy(x_out, s->xout_h, sizeof(double) * 3 * NF);
        const int rows = hs.iter;
        if (hp && rows) AAADMM_CUDA_OK(cudaMemcpy(hp, s->hist_prim, sizeof(double) * rows, cudaMemcpyDeviceToHost));
        if (hc && rows) AAADMM_CUDA_OK(cudaMemcpy(hc, s->hist_comb, sizeof(double) * rows, cudaMemcpyDeviceToHost));
        if (hr && rows) AAADMM_CUDA_OK(cudaMemcpy(hr, s->hist_rej, sizeof(int) * rows, cudaMemcpyDeviceToHost));
    }
    if (res) {
        res->iters_logged = hs.iter;
        res->rejects = hs.n_rejects;
        res->broke_early = hs.done;
        cudaEventElapsedTime(&res->loop_ms, s->ev[1], s->ev[2]);
        cudaEventElapsedTime(&res->step_ms, s->ev[0], s->ev[3]);
        res->kernel_launches = s->launches + hs.loop_it * s->body_launches;
    }
    return 0;
}

int aaadmm_tetscene_step(aaadmm_tetscene *s, const aaadmm_step_opts *o, const double *x_bar, const double *x_pin,
                         double *x_out, double *hp, double *hc, int *hr, aaadmm_step_result *res) {
    API_TRY_BEGIN
    if (!x_bar || !x_out || (s->NP > 0 && !x_pin)) {
        set_last_error("tetscene_step: null buffer");
        return -1;
    }
    return step_common(s, o, true, x_bar, x_pin, x_out, hp, hc, hr, res);
    API_TRY_END
}
int aaadmm_tetscene_step_resident(aaadmm_tetscene *s, const aaadmm_step_opts *o, aaadmm_step_result *res) {
    API_TRY_BEGIN
    return step_common(s, o, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, res);
    API_TRY_END
}

int aaadmm_tetscene_iteration_times(aaadmm_tetscene *s, double *ms, int n) {
    API_TRY_BEGIN
    if (!s || !ms || n < 0 || n > s->last_rows) {
        set_last_error("tetscene_iteration_times: bad arguments (n must not exceed the rows of the last step)");
        return -1;
    }
    if (n > 0) AAADMM_CUDA_OK(cudaMemcpy(ms, s->hist_comb + s->last_max_iters, sizeof(double) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) ms[i] *= 1e-6;
    return 0;
    API_TRY_END
}

int aaadmm_tetscene_read_zu(aaadmm_tetscene *s, double *z, double *u) {
    API_TRY_BEGIN
    if (!s) {
        set_last_error("tetscene_read_zu: null scene");
        return -1;
    }
    if (s->T <= 0) return 0;  // tet terms only (reference layout of the tets' z / u)
    if (!s->last_z || !s->last_u) {
        set_last_error("tetscene_read_zu: call aaadmm_tetscene_step first");
        return -1;
    }
    if (!s->zu_scratch) AAADMM_CUDA_OK(cudaMalloc((void **)&s->zu_scratch, sizeof(double) * 9 * s->T));
    const int g = (s->T + 255) / 256;
    if (z) {
        k_soa_to_aos9<<<g, 256, 0, s->stream>>>(s->last_z, s->zu_scratch, s->T);
        AAADMM_CUDA_OK(cudaMemcpyAsync(z, s->zu_scratch, sizeof(double) * 9 * s->T, cudaMemcpyDeviceToHost, s->stream));
        AAADMM_CUDA_OK(cudaStreamSynchronize(s->stream));
    }
    if (u) {
        k_soa_to_aos9<<<g, 256, 0, s->stream>>>(s->last_u, s->zu_scratch, s->T);
        AAADMM_CUDA_OK(cudaMemcpyAsync(u, s->zu_scratch, sizeof(double) * 9 * s->T, cudaMemcpyDeviceToHost, s->stream));
        AAADMM_CUDA_OK(cudaStreamSynchronize(s->stream));
    }
    return 0;
    API_TRY_END
}

int aaadmm_tetscene_profile(aaadmm_tetscene *s, const aaadmm_step_opts *o, int iters, float *ms) {
    API_TRY_BEGIN
    if (!s->has_inputs) {
        set_last_error("tetscene_profile: call aaadmm_tetscene_step once first");
        return -1;
    }
    const bool accel = o->accel && o->anderson_m > 0;
    if (scene_reserve(s, std::max(1, iters), accel ? o->anderson_m : 1)) return -1;
    PhaseProf prof;
    prof.on = true;
    prof.s = s->stream;
    aaadmm_step_opts oo = *o;
    oo.eps = -1.0;  // never break: every phase runs in every iteration
    if (run_hard(s, &oo, nullptr, 0, true)) return -1;
    if (run_hard(s, &oo, &prof, iters, false)) return -1;
    AAADMM_CUDA_OK(cudaStreamSynchronize(s->stream));
    double sum[AAADMM_NPROF] = {0};
    for (size_t k = 0; k < prof.phase.size(); ++k) {
        float t = 0;
        cudaEventElapsedTime(&t, prof.ev[2 * k], prof.ev[2 * k + 1]);
        sum[prof.phase[k]] += t;
        sum[AAADMM_NPROF - 1] += t;
    }
    for (auto e : prof.ev) cudaEventDestroy(e);
    for (int p = 0; p < AAADMM_NPROF; ++p) ms[p] = (float)(sum[p] / std::max(1, iters));
    return 0;
    API_TRY_END
}

int aaadmm_tetscene_algo_bytes(aaadmm_tetscene *s, int m, double *b) {
    const double T = s->T, V = s->V, NF = s->NF, Ne = (double)s->Ne, Nt = (double)s->Nt, R = s->NT;
    const double ninc = 4.0 * T + 3.0 * R;  // upper bound: incidences of free vertices
    b[0] = T * (16 + 72 + 8 + 72 + 72 + 96) + 24 * V;           // update_z: idx,B^-1,w,u in; z,contrib out; positions
    b[0] += R * (16 + 32 + 8 + 16 + 48 + 48 + 72);              // triangles: idx,rest pose,w,limits,u in; z,contrib out
    b[1] = T * 96 + R * 72 + 4 * ninc + NF * (8 + 24 + 24 + 4); // rhs gather: contrib, inc, ptr, bconst, out, iperm
    b[2] = s->factor->f->stats.bytes_per_solve;                 // ldlt apply
    b[3] = T * (16 + 72 + 8 + 72 + 72 + 72) + 2 * 24 * V;       // update_u + residuals
    b[3] += R * (16 + 32 + 8 + 48 + 48 + 48);
    const double C = s->NC;  // collision terms: vertex id, weight, position, u in; z, contribution out (+ u out)
    b[0] += C * (4 + 8 + 24 + 24 + 24 + 24);
    b[1] += C * (24 + 4);
    b[3] += C * (4 + 8 + 48 + 24 + 24 + 24);
    b[4] = 8.0 * ((m + 2) * Ne + 3 * Nt);                       // aa pass 1
    b[5] = 8.0 * ((m + 3) * Nt + 2 * Ne);                       // aa pass 2
    b[6] = 0;
    b[7] = b[0] + b[1] + b[2] + b[3] + b[4] + b[5];
    return 0;
}

// ------------------------------------------------------------------------------------------
// Batched element kernels (unit parity)
// ------------------------------------------------------------------------------------------
// ---- rows I / J: triangle, collision and spring-pin prox batches on host arrays -------------------------
namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { cudaFree(p); }
    int up(const void *src, size_t bytes) {
        AAADMM_CUDA_OK(cudaMalloc(&p, std::max<size_t>(bytes, 8)));
        if (bytes) AAADMM_CUDA_OK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
        return 0;
    }
};
}  // namespace
int aaadmm_tri_prox(int variant, double *z, int64_t n, double limit_min, double limit_max) {
    if (aaadmm_device_count() <= 0) {
        set_last_error("no CUDA device (this library has no CPU path)");
        return -1;
    }
    if (!z || n < 0 || (variant != AAADMM_ORDER_HARD_ZXU && variant != AAADMM_ORDER_XZU)) {
        set_last_error("tri_prox: bad arguments");
        return -1;
    }
    DevBuf d;
    if (d.up(z, sizeof(double) * 6 * n)) return -1;
    launch_tri_prox(variant == AAADMM_ORDER_XZU ? 0 : 1, (double *)d.p, n, limit_min, limit_max, nullptr);
    AAADMM_CUDA_OK(cudaGetLastError());
    AAADMM_CUDA_OK(cudaMemcpy(z, d.p, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost));
    return 0;
}
int aaadmm_collision_prox(int n_objs, const int *types, const double *params7, double *z, int64_t n) {
    if (aaadmm_device_count() <= 0) {
        set_last_error("no CUDA device (this library has no CPU path)");
        return -1;
    }
    if (!z || n < 0 || n_objs < 0 || (n_objs > 0 && (!types || !params7))) {
        set_last_error("collision_prox: bad arguments");
        return -1;
    }
    for (int j = 0; j < n_objs; ++j)
        if (types[j] < AAADMM_PASSIVE_FLOOR || types[j] > AAADMM_PASSIVE_CYLINDER) {
            set_last_error("collision_prox: unknown passive object type (triangle-mesh obstacles have no device batch)");
            return -1;
        }
    DevBuf d, t, p;
    if (d.up(z, sizeof(double) * 3 * n) || t.up(types, sizeof(int) * n_objs) || p.up(params7, sizeof(double) * 7 * n_objs)) return -1;
    launch_collision_prox(n_objs, (const int *)t.p, (const double *)p.p, (double *)d.p, n, nullptr);
    AAADMM_CUDA_OK(cudaGetLastError());
    AAADMM_CUDA_OK(cudaMemcpy(z, d.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    return 0;
}
int aaadmm_spring_prox(double *z, const double *pins, const int *active, int64_t n) {
    if (aaadmm_device_count() <= 0) {
        set_last_error("no CUDA device (this library has no CPU path)");
        return -1;
    }
    if (!z || !pins || !active || n < 0) {
        set_last_error("spring_prox: bad arguments");
        return -1;
    }
    DevBuf d, p, a;
    if (d.up(z, sizeof(double) * 3 * n) || p.up(pins, sizeof(double) * 3 * n) || a.up(active, sizeof(int) * n)) return -1;
    launch_spring_prox((double *)d.p, (const double *)p.p, (const int *)a.p, n, nullptr);
    AAADMM_CUDA_OK(cudaGetLastError());
    AAADMM_CUDA_OK(cudaMemcpy(z, d.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    return 0;
}
int aaadmm_tet_prox_linear(double *z, int64_t n) {
    if (aaadmm_device_count() <= 0) {
        set_last_error("no CUDA device (this library has no CPU path)");
        return -1;
    }
    double *d = nullptr;
    AAADMM_CUDA_OK(cudaMalloc((void **)&d, sizeof(double) * 9 * n));
    AAADMM_CUDA_OK(cudaMemcpy(d, z, sizeof(double) * 9 * n, cudaMemcpyHostToDevice));
    launch_prox_batch(d, n);
    AAADMM_CUDA_OK(cudaGetLastError());
    AAADMM_CUDA_OK(cudaMemcpy(z, d, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}
int aaadmm_tet_f_minus_uvt(const double *z, double *out, int64_t n) {
    if (aaadmm_device_count() <= 0) {
        set_last_error("no CUDA device (this library has no CPU path)");
        return -1;
    }
    double *d = nullptr, *o = nullptr;
    AAADMM_CUDA_OK(cudaMalloc((void **)&d, sizeof(double) * 9 * n));
    AAADMM_CUDA_OK(cudaMalloc((void **)&o, sizeof(double) * 9 * n));
    AAADMM_CUDA_OK(cudaMemcpy(d, z, sizeof(double) * 9 * n, cudaMemcpyHostToDevice));
    launch_fmuvt_batch(d, o, n);
    AAADMM_CUDA_OK(cudaGetLastError());
    AAADMM_CUDA_OK(cudaMemcpy(out, o, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    cudaFree(o);
    return 0;
}
int aaadmm_tet_prox_hyper(int material, double mu, double lambda, double vol, double *z, double *grad, int64_t n) {
    if (aaadmm_device_count() <= 0) {
        set_last_error("no CUDA device (this library has no CPU path)");
        return -1;
    }
    double *d = nullptr, *g = nullptr;
    AAADMM_CUDA_OK(cudaMalloc((void **)&d, sizeof(double) * 9 * n));
    AAADMM_CUDA_OK(cudaMalloc((void **)&g, sizeof(double) * 9 * n));
    AAADMM_CUDA_OK(cudaMemcpy(d, z, sizeof(double) * 9 * n, cudaMemcpyHostToDevice));
    launch_prox_hyper_batch(material, mu, lambda, vol, d, g, n);
    AAADMM_CUDA_OK(cudaGetLastError());
    AAADMM_CUDA_OK(cudaMemcpy(z, d, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost));
    if (grad) AAADMM_CUDA_OK(cudaMemcpy(grad, g, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost));
    cudaFree(d);
    cudaFree(g);
    return 0;
}
int aaadmm_cod_solve(int m, const double *M, const double *rhs, double *x, int *rank) {
    if (aaadmm_device_count() <= 0) {
        set_last_error("no CUDA device (this library has no CPU path)");
        return -1;
    }
    if (m <= 0 || m > AA_MAX_M) {
        set_last_error("cod_solve: m out of range");
        return -1;
    }
    double *dM = nullptr, *dr = nullptr, *dx = nullptr;
    int *dk = nullptr;
    AAADMM_CUDA_OK(cudaMalloc((void **)&dM, sizeof(double) * m * m));
    AAADMM_CUDA_OK(cudaMalloc((void **)&dr, sizeof(double) * m));
    AAADMM_CUDA_OK(cudaMalloc((void **)&dx, sizeof(double) * m));
    AAADMM_CUDA_OK(cudaMalloc((void **)&dk, sizeof(int)));
    AAADMM_CUDA_OK(cudaMemcpy(dM, M, sizeof(double) * m * m, cudaMemcpyHostToDevice));
    AAADMM_CUDA_OK(cudaMemcpy(dr, rhs, sizeof(double) * m, cudaMemcpyHostToDevice));
    k_cod<<<1, 32>>>(m, dM, dr, dx, dk);
    AAADMM_CUDA_OK(cudaGetLastError());
    AAADMM_CUDA_OK(cudaMemcpy(x, dx, sizeof(double) * m, cudaMemcpyDeviceToHost));
    if (rank) AAADMM_CUDA_OK(cudaMemcpy(rank, dk, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(dM);
    cudaFree(dr);
    cudaFree(dx);
    cudaFree(dk);
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// Geometry ADMM
// ------------------------------------------------------------------------------------------
#include "geo_kernels.cuh"

#include <cfloat>
#include <numeric>

namespace {

// Median-split BVH over triangle centroids (built once on the host; the mesh is static).
struct BvhBuild {
    std::vector<BvhNode> nodes;
    std::vector<int> order;
    const double *tri;
    int build(int lo, int hi) {
        const int id = (int)nodes.size();
        nodes.emplace_back();
        BvhNode nd;
        for (int r = 0; r < 3; ++r) nd.lo[r] = DBL_MAX, nd.hi[r] = -DBL_MAX;
        double clo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, chi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int k = lo; k < hi; ++k) {
            const double *t = tri + 9 * (size_t)order[k];
            for (int r = 0; r < 3; ++r) {
                double c = 0;
                for (int v = 0; v < 3; ++v) {
                    nd.lo[r] = std::min(nd.lo[r], t[3 * v + r]);
                    nd.hi[r] = std::max(nd.hi[r], t[3 * v + r]);
                    c += t[3 * v + r];
                }
                clo[r] = std::min(clo[r], c);
                chi[r] = std::max(chi[r], c);
            }
        }
        if (hi - lo <= 4) {
            nd.left = -(lo + 1);
            nd.right = hi - lo;
            nodes[id] = nd;
            return id;
        }
        int ax = 0;
        if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
        if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
        const int mid = (lo + hi) / 2;
        const double *T = tri;
        std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, [T, ax](int a, int b) {
            const double ca = T[9 * (size_t)a + ax] + T[9 * (size_t)a + 3 + ax] + T[9 * (size_t)a + 6 + ax];
            const double cb = T[9 * (size_t)b + ax] + T[9 * (size_t)b + 3 + ax] + T[9 * (size_t)b + 6 + ax];
            return ca < cb || (ca == cb && a < b);
        });
        const int l = build(lo, mid), r = build(mid, hi);
        nd.left = l;
        nd.right = r;
        nodes[id] = nd;
        return id;
    }
};

struct RefMeshDev {
    BvhNode *nodes = nullptr;
    int *order = nullptr;
    double *tri = nullptr;
    int n_tris = 0;
    int upload(const double *verts, int nv, const int *tris, int nt) {
        (void)nv;
        std::vector<double> T((size_t)9 * std::max(nt, 1));
        for (int t = 0; t < nt; ++t)
            for (int v = 0; v < 3; ++v)
                for (int r = 0; r < 3; ++r) T[9 * (size_t)t + 3 * v + r] = verts[3 * (size_t)tris[3 * t + v] + r];
        BvhBuild B;
        B.tri = T.data();
        B.order.resize(nt);
        std::iota(B.order.begin(), B.order.end(), 0);
        if (nt > 0) B.build(0, nt);
        n_tris = nt;
        AAADMM_CUDA_OK(cudaMalloc((void **)&nodes, sizeof(BvhNode) * std::max<size_t>(B.nodes.size(), 1)));
        AAADMM_CUDA_OK(cudaMemcpy(nodes, B.nodes.data(), sizeof(BvhNode) * B.nodes.size(), cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&order, sizeof(int) * std::max(nt, 1)));
        AAADMM_CUDA_OK(cudaMemcpy(order, B.order.data(), sizeof(int) * nt, cudaMemcpyHostToDevice));
        AAADMM_CUDA_OK(cudaMalloc((void **)&tri, sizeof(double) * T.size()));
        AAADMM_CUDA_OK(cudaMemcpy(tri, T.data(), sizeof(double) * T.size(), cudaMemcpyHostToDevice));
        return 0;
    }
    void release() {
        cudaFree(nodes);
        cudaFree(order);
        cudaFree(tri);
    }
};

template <typename T>
int up(T **dst, const T *src, size_t n) {
    AAADMM_CUDA_OK(cudaMalloc((void **)dst, sizeof(T) * std::max<size_t>(n, 1)));
    if (n) AAADMM_CUDA_OK(cudaMemcpy(*dst, src, sizeof(T) * n, cudaMemcpyHostToDevice));
    return 0;
}

__global__ void k_geo_init_state(SolveState *st, int m, int max_iters) {
    st->prim2 = 0.0;
    st->hyper_prim2 = 0.0;
    st->prev_prim = DBL_MAX;  // prev_residual
    st->comb = 0.0;
    st->eps = 0.0;
    st->reject = 0;  // reset
    st->done = 0;
    st->iter = 0;
    st->n_rejects = 0;
    st->accel = m > 0;
    st->aa_iter = 0;
    st->aa_col = 0;
    st->aa_m = m > 0 ? m : 1;
    st->aa_mk = 0;
    st->ticket = 0u;
    st->t0 = global_timer_ns();  // per-iteration time stamps of the log count from here
    st->loop_it = 0;
    st->max_iters = max_iters;
    st->skip_redo = 1;
    st->aa_skip = 0;
}
// ALMGeometrySolver.h:258-263: the loop ends when iter_count >= max_iter (accepted iterations)
__global__ void k_geo_loop_cond(cudaGraphConditionalHandle h, SolveState *st) {
    st->loop_it += 1;
    cudaGraphSetConditional(h, (st->iter < st->max_iters && st->loop_it < 4 * st->max_iters + 8) ? 1u : 0u);
}

}  // namespace

struct aaadmm_geo {
    int P = 0, n_hard = 0, zc = 0, n_soft = 0, variant = 0;
    int zc_all = 0;  // z / u columns: zc (ALM) or zc + n_soft (GS)
    aaadmm_ldlt *factor = nullptr;
    cudaStream_t stream = nullptr;
    int *type = nullptr, *idx_ptr = nullptr, *idx = nullptr, *col0 = nullptr, *dt_col = nullptr, *soft_point = nullptr,
        *soft_of_point = nullptr, *last_tri = nullptr, *soft_order = nullptr;
    std::vector<int> soft_point_h;  // host copy: the first solve sorts the soft points along a Morton curve
    bool soft_order_built = false;
    int64_t *dt_ptr = nullptr;
    double *param = nullptr, *dt_val = nullptr, *rhs_fixed = nullptr;
    double soft_weight = 0, rho = 1.0;
    RefMeshDev mesh;
    int64_t N = 0;  // 3 zc + 3 P: the Anderson variable (u | x)
    double *Ubuf = nullptr, *Nbuf = nullptr, *Dbuf = nullptr, *z = nullptr, *zmu = nullptr, *prev_dx = nullptr, *cp = nullptr;
    double *dF = nullptr, *dG = nullptr, *hist = nullptr, *partials = nullptr;
    int m_cap = 0, hist_cap = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    SolveState *st_h = nullptr;  // pinned host copy of the control block
    int last_iters = 0, last_max_iter = 0;  // of the last solve (aaadmm_geo_reset_flags)
    SolveState *st = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int graph_key = -1;
    int body_launches = 0;
};

extern "C" {

int aaadmm_geo_destroy(aaadmm_geo *g) {
    if (!g) return 0;
    for (auto &e : g->ev)
        if (e) cudaEventDestroy(e);
    if (g->st_h) cudaFreeHost(g->st_h);
    cudaFree(g->type);
    cudaFree(g->idx_ptr);
    cudaFree(g->idx);
    cudaFree(g->col0);
    cudaFree(g->dt_col);
    cudaFree(g->soft_point);
    cudaFree(g->soft_of_point);
    cudaFree(g->last_tri);
    cudaFree(g->soft_order);
    cudaFree(g->dt_ptr);
    cudaFree(g->param);
    cudaFree(g->dt_val);
    cudaFree(g->rhs_fixed);
    g->mesh.release();
    cudaFree(g->Ubuf);
    cudaFree(g->Nbuf);
    cudaFree(g->Dbuf);
    cudaFree(g->z);
    cudaFree(g->zmu);
    cudaFree(g->prev_dx);
    cudaFree(g->cp);
    cudaFree(g->dF);
    cudaFree(g->dG);
    cudaFree(g->hist);
    cudaFree(g->partials);
    cudaFree(g->st);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
    return 0;
}

int aaadmm_geo_create(aaadmm_geo **out, const aaadmm_geo_desc *d, aaadmm_ldlt *factor) {
    API_TRY_BEGIN
    if (aaadmm_device_count() <= 0) {
        set_last_error("geo_create: no CUDA device (this library has no CPU path)");
        return -1;
    }
    if (!d || !factor || d->n_points <= 0 || factor->f->n != d->n_points || factor->f->nrhs != 3) {
        set_last_error("geo_create: bad arguments (factor must be n_points x n_points with nrhs = 3)");
        return -1;
    }
    std::vector<int> col0(std::max(d->n_hard, 1));
    int zc = 0;
    for (int c = 0; c < d->n_hard; ++c) {
        const int k = d->idx_ptr[c + 1] - d->idx_ptr[c];
        if (d->type[c] < 0 || d->type[c] > 2 || k < 2 || k > GEO_MAX_K || (d->type[c] == GEO_EDGE && k != 2) ||
            (d->type[c] == GEO_ANGLE && k != 3)) {
            set_last_error("geo_create: unsupported constraint (type must be plane/edge/angle, plane size <= 16)");
            return -1;
        }
        col0[c] = zc;
        zc += d->type[c] == GEO_PLANE ? k : k - 1;
    }
    if (zc != d->n_zcols) {
        set_last_error("geo_create: n_zcols does not match the constraint list");
        return -1;
    }
    aaadmm_geo *g = new aaadmm_geo();
    g->P = d->n_points;
    g->n_hard = d->n_hard;
    g->zc = zc;
    g->n_soft = d->n_soft;
    g->soft_weight = d->soft_weight;
    g->factor = factor;
    g->variant = d->variant == AAADMM_GEO_GS ? AAADMM_GEO_GS : AAADMM_GEO_ALM;
    g->rho = d->rho;
    g->zc_all = zc + (g->variant == AAADMM_GEO_GS ? d->n_soft : 0);
    g->N = 3 * (int64_t)g->zc_all + 3 * (int64_t)g->P;
    AAADMM_CUDA_OK(cudaStreamCreate(&g->stream));
    for (auto &e : g->ev) AAADMM_CUDA_OK(cudaEventCreate(&e));
    AAADMM_CUDA_OK(cudaMallocHost((void **)&g->st_h, sizeof(SolveState)));
    int rc = 0;
    rc |= up(&g->type, d->type, d->n_hard);
    rc |= up(&g->idx_ptr, d->idx_ptr, d->n_hard + 1);
    rc |= up(&g->idx, d->idx, d->n_hard ? d->idx_ptr[d->n_hard] : 0);
    rc |= up(&g->col0, col0.data(), d->n_hard);
    rc |= up(&g->param, d->param, (size_t)4 * d->n_hard);
    rc |= up(&g->dt_ptr, d->dt_ptr, (size_t)g->P + 1);
    rc |= up(&g->dt_col, d->dt_col, (size_t)d->dt_ptr[g->P]);
    rc |= up(&g->dt_val, d->dt_val, (size_t)d->dt_ptr[g->P]);
    rc |= up(&g->rhs_fixed, d->rhs_fixed, (size_t)3 * g->P);
    rc |= up(&g->soft_point, d->soft_point, d->n_soft);
    if (d->n_soft > 0) {
        g->soft_point_h.assign(d->soft_point, d->soft_point + d->n_soft);
        std::vector<int> ident(d->n_soft);
        for (int i = 0; i < d->n_soft; ++i) ident[i] = i;
        rc |= up(&g->soft_order, ident.data(), d->n_soft);
    }
    std::vector<int> sop(g->P, -1), lt(std::max(d->n_soft, 1), -1);
    for (int i = 0; i < d->n_soft; ++i) sop[d->soft_point[i]] = i;
    rc |= up(&g->soft_of_point, sop.data(), g->P);
    rc |= up(&g->last_tri, lt.data(), d->n_soft);
    if (d->n_soft > 0) rc |= g->mesh.upload(d->ref_verts, d->n_ref_verts, d->ref_tris, d->n_ref_tris);
    if (rc) {
        aaadmm_geo_destroy(g);
        return -1;
    }
    AAADMM_CUDA_OK(cudaMalloc((void **)&g->Ubuf, sizeof(double) * g->N));
    AAADMM_CUDA_OK(cudaMalloc((void **)&g->Nbuf, sizeof(double) * g->N));
    AAADMM_CUDA_OK(cudaMalloc((void **)&g->Dbuf, sizeof(double) * g->N));
    AAADMM_CUDA_OK(cudaMalloc((void **)&g->z, sizeof(double) * 3 * std::max(g->zc_all, 1)));
    AAADMM_CUDA_OK(cudaMalloc((void **)&g->zmu, sizeof(double) * 3 * std::max(g->zc_all, 1)));
    AAADMM_CUDA_OK(cudaMemset(g->zmu, 0, sizeof(double) * 3 * std::max(g->zc_all, 1)));
    AAADMM_CUDA_OK(cudaMalloc((void **)&g->prev_dx, sizeof(double) * 3 * std::max(g->zc_all, 1)));
    AAADMM_CUDA_OK(cudaMalloc((void **)&g->cp, sizeof(double) * 3 * std::max(d->n_soft, 1)));
    AAADMM_CUDA_OK(cudaMalloc((void **)&g->partials, sizeof(double) * RED_MAX_Q * RED_MAX_BLOCKS));
    AAADMM_CUDA_OK(cudaMalloc((void **)&g->st, sizeof(SolveState)));
    AAADMM_CUDA_OK(cudaMemset(g->st, 0, sizeof(SolveState)));
    *out = g;
    return 0;
    API_TRY_END
}

static void geo_views(aaadmm_geo *g, GeoConstraints &C, GeoSoft &S) {
    C.n = g->n_hard;
    C.type = g->type;
    C.idx_ptr = g->idx_ptr;
    C.idx = g->idx;
    C.col0 = g->col0;
    C.param = g->param;
    S.n = g->n_soft;
    S.point = g->soft_point;
    S.weight = g->soft_weight;
    S.nodes = g->mesh.nodes;
    S.tri_order = g->mesh.order;
    S.tri = g->mesh.tri;
    S.last_tri = g->last_tri;
    S.order = g->soft_order;
}

// one turn of the while loop of ALMGeometrySolver.h:197-268
static int geo_enqueue_turn(aaadmm_geo *g, int m, int &L) {
    cudaStream_t st = g->stream;
    LdltDev *f = g->factor->f;
    GeoConstraints C;
    GeoSoft S;
    geo_views(g, C, S);
    const int64_t NU = 3 * (int64_t)g->zc;
    double *cu = g->Ubuf, *cx = g->Ubuf + NU, *nu = g->Nbuf, *nx = g->Nbuf + NU;
    const bool accel = m > 0;
    launch_geo_local(st, C, cx, cu, g->prev_dx, g->z, g->zmu, g->st);
    launch_geo_soft(st, S, cx, g->cp, g->st);
    launch_geo_rhs(st, g->P, g->dt_ptr, g->dt_col, g->dt_val, g->zmu, nullptr, g->rhs_fixed, g->n_soft ? g->soft_of_point : nullptr,
                   g->soft_weight, g->cp, f->iperm, f->W, g->st);
    if (ldlt_dev_apply_permuted(f, nx, st, &g->st->done)) return -1;
    launch_geo_u_resid(st, C, nx, cu, g->z, g->prev_dx, nu, g->st, g->partials, g->hist, accel ? 1 : 0);
    L += 4 + f->n_launches;
    if (accel) {
        const int gs = stream_grid(4);
        if (launch_aa_pass1(m, gs, st, g->Nbuf, nullptr, nullptr, g->Ubuf, g->dF, g->dG, g->N, g->N, g->st, g->partials, g->Dbuf))
            return -1;
        if (launch_aa_pass2(m, gs, st, g->Nbuf, nullptr, g->Ubuf, g->dF, g->dG, g->N, g->N, g->st)) return -1;
        L += 2;
    }
    launch_geo_select(st, g->Ubuf, g->Dbuf, g->Nbuf, g->N, g->st, accel ? 1 : 0);
    L += 1;
    return 0;
}

// GeometrySolver<3>: the warm-up turn of ADMM_init_variables (Geometry/GeometrySolver.h:383-407)
static int gs_enqueue_x_u(aaadmm_geo *g, const GeoConstraints &C, const GeoSoft &S, int &L) {
    cudaStream_t st = g->stream;
    LdltDev *f = g->factor->f;
    const int64_t NU = 3 * (int64_t)g->zc_all;
    double *cu = g->Ubuf, *du = g->Dbuf, *dx = g->Dbuf + NU;
    // x_update: default_x = solve(rhs_fixed + rho D^T (z - current_u))   (:436-443)
    launch_geo_rhs(st, g->P, g->dt_ptr, g->dt_col, g->dt_val, g->z, cu, g->rhs_fixed, nullptr, 0.0, nullptr, f->iperm,
                   f->W, g->st);
    if (ldlt_dev_apply_permuted(f, dx, st, &g->st->done)) return -1;
    launch_gs_dx(st, C, S, g->zc, dx, g->prev_dx, g->st, 0);
    launch_gs_u(st, cu, g->prev_dx, g->z, du, NU, g->st);
    L += 3 + f->n_launches;
    return 0;
}
static int gs_enqueue_warmup(aaadmm_geo *g, int &L) {
    cudaStream_t st = g->stream;
    GeoConstraints C;
    GeoSoft S;
    geo_views(g, C, S);
    const int64_t NU = 3 * (int64_t)g->zc_all;
    launch_gs_dx(st, C, S, g->zc, g->Ubuf + NU, g->prev_dx, g->st, 0);
    launch_gs_z(st, 0, C, S, g->zc, g->rho, g->prev_dx, g->Ubuf, g->z, nullptr, g->st, g->partials, g->hist);
    if (gs_enqueue_x_u(g, C, S, L)) return -1;
    launch_gs_take_default(st, g->Ubuf, g->Dbuf, g->N, g->st, 0);
    L += 3;
    return 0;
}
// one turn of the while loop of GeometrySolver.h:181-254; Dx of current_x is in prev_dx on entry
static int gs_enqueue_turn(aaadmm_geo *g, int m, int &L) {
    cudaStream_t st = g->stream;
    GeoConstraints C;
    GeoSoft S;
    geo_views(g, C, S);
    const int64_t NU = 3 * (int64_t)g->zc_all;
    double *cu = g->Ubuf, *cx = g->Ubuf + NU;
    launch_gs_z(st, 1, C, S, g->zc, g->rho, g->prev_dx, cu, g->z, nullptr, g->st, g->partials, g->hist);
    L += 1;
    if (m > 0) {  // need_reset: back to the un-accelerated iterate, z again
        launch_gs_take_default(st, g->Ubuf, g->Dbuf, g->N, g->st, 1);
        launch_gs_dx(st, C, S, g->zc, cx, g->prev_dx, g->st, 1);
        launch_gs_z(st, 2, C, S, g->zc, g->rho, g->prev_dx, cu, g->z, nullptr, g->st, g->partials, g->hist);
        L += 3;
    }
    if (gs_enqueue_x_u(g, C, S, L)) return -1;
    if (m > 0) {
        const int gs = stream_grid(4);
        if (launch_aa_pass1(m, gs, st, g->Dbuf, g->Dbuf + NU, nullptr, g->Ubuf, g->dF, g->dG, NU, g->N, g->st, g->partials))
            return -1;
        if (launch_aa_pass2(m, gs, st, g->Dbuf, g->Dbuf + NU, g->Ubuf, g->dF, g->dG, NU, g->N, g->st)) return -1;
        L += 2;
    } else {
        launch_gs_take_default(st, g->Ubuf, g->Dbuf, g->N, g->st, 0);
        L += 1;
    }
    launch_gs_dx(st, C, S, g->zc, cx, g->prev_dx, g->st, 0);
    L += 1;
    return 0;
}
static int geo_enqueue_any(aaadmm_geo *g, int m, int &L) {
    return g->variant == AAADMM_GEO_GS ? gs_enqueue_turn(g, m, L) : geo_enqueue_turn(g, m, L);
}

int aaadmm_geo_solve(aaadmm_geo *g, const double *init_x, int max_iter, int anderson_m, double *x_out, double *hist,
                     aaadmm_step_result *res) {
    API_TRY_BEGIN
    if (!g || !init_x || !x_out || max_iter < 0 || anderson_m > AA_MAX_M) {
        set_last_error("geo_solve: bad arguments");
        return -1;
    }
    const int m = anderson_m > 0 ? anderson_m : 0;
    cudaStream_t st = g->stream;
    if (max_iter > g->hist_cap) {
        cudaFree(g->hist);
        AAADMM_CUDA_OK(cudaMalloc((void **)&g->hist, sizeof(double) * 3 * std::max(1, max_iter)));  // residuals | reset flags | ms
        g->hist_cap = max_iter;
        g->graph_key = -1;
    }
    if (m > g->m_cap) {
        cudaFree(g->dF);
        cudaFree(g->dG);
        AAADMM_CUDA_OK(cudaMalloc((void **)&g->dF, sizeof(double) * g->N * m));
        AAADMM_CUDA_OK(cudaMalloc((void **)&g->dG, sizeof(double) * g->N * m));
        g->m_cap = m;
        g->graph_key = -1;
    }
    const int64_t NU = 3 * (int64_t)g->zc_all;
    cudaEvent_t e0 = g->ev[0], e1 = g->ev[1];
    // init_variables (ALMGeometrySolver.h:404-409) + aa->init(current_u, current_x)
    AAADMM_CUDA_OK(cudaMemsetAsync(g->Ubuf, 0, sizeof(double) * g->N, st));
    AAADMM_CUDA_OK(cudaMemsetAsync(g->Dbuf, 0, sizeof(double) * g->N, st));
    AAADMM_CUDA_OK(cudaMemcpyAsync(g->Ubuf + NU, init_x, sizeof(double) * 3 * g->P, cudaMemcpyHostToDevice, st));
    AAADMM_CUDA_OK(cudaMemcpyAsync(g->Dbuf + NU, init_x, sizeof(double) * 3 * g->P, cudaMemcpyHostToDevice, st));
    if (g->n_soft) AAADMM_CUDA_OK(cudaMemsetAsync(g->last_tri, 0xff, sizeof(int) * g->n_soft, st));
    if (g->n_soft >= 1024 && !g->soft_order_built) {
        // processing order of the closest-point queries: Morton order of the first positions (the points move little
        // during a solve); results do not depend on it
        const int n = g->n_soft;
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < 3; ++r) {
                const double c = init_x[3 * (size_t)g->soft_point_h[i] + r];
                lo[r] = std::min(lo[r], c);
                hi[r] = std::max(hi[r], c);
            }
        std::vector<std::pair<uint64_t, int>> key(n);
        for (int i = 0; i < n; ++i) {
            uint64_t code = 0;
            uint32_t q[3];
            for (int r = 0; r < 3; ++r) {
                const double c = init_x[3 * (size_t)g->soft_point_h[i] + r], w = hi[r] - lo[r];
                q[r] = w > 0 ? (uint32_t)std::min(2097151.0, (c - lo[r]) / w * 2097152.0) : 0u;
            }
            for (int bit = 20; bit >= 0; --bit)
                for (int r = 0; r < 3; ++r) code = (code << 1) | ((q[r] >> bit) & 1u);
            key[i] = {code, i};
        }
        std::sort(key.begin(), key.end());
        std::vector<int> ord(n);
        for (int i = 0; i < n; ++i) ord[i] = key[i].second;
        AAADMM_CUDA_OK(cudaMemcpyAsync(g->soft_order, ord.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
        AAADMM_CUDA_OK(cudaStreamSynchronize(st));  // `ord` is a local
        g->soft_order_built = true;
    }
    k_geo_init_state<<<1, 1, 0, st>>>(g->st, m, max_iter);
    int launches = 0;
    static const bool no_graph = getenv("AAADMM_NO_GRAPH") != nullptr;
    SolveState &hs = *g->st_h;
    if (max_iter > 0 && !no_graph) {
        // the loop graph is built (once per window size) before the timed region starts
        const int key = m;
        if (g->graph_key != key || !g->exec) {
            if (g->exec) cudaGraphExecDestroy(g->exec), g->exec = nullptr;
            if (g->graph) cudaGraphDestroy(g->graph), g->graph = nullptr;
            cudaGraphConditionalHandle h;
            AAADMM_CUDA_OK(cudaGraphCreate(&g->graph, 0));
            AAADMM_CUDA_OK(cudaGraphConditionalHandleCreate(&h, g->graph, 1, cudaGraphCondAssignDefault));
            cudaGraphNodeParams np = {};
            np.type = cudaGraphNodeTypeConditional;
            np.conditional.handle = h;
            np.conditional.type = cudaGraphCondTypeWhile;
            np.conditional.size = 1;
            cudaGraphNode_t node;
            AAADMM_CUDA_OK(cudaGraphAddNode(&node, g->graph, nullptr, 0, &np));
            AAADMM_CUDA_OK(cudaStreamBeginCaptureToGraph(st, np.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                                         cudaStreamCaptureModeThreadLocal));
            int L = 0;
            const int rc = geo_enqueue_any(g, m, L);
            k_geo_loop_cond<<<1, 1, 0, st>>>(h, g->st);
            cudaError_t e = cudaStreamEndCapture(st, nullptr);
            if (rc || e != cudaSuccess) {
                cudaGetLastError();
                cudaGraphDestroy(g->graph), g->graph = nullptr;
                g->graph_key = -1;
                set_last_error(std::string("geo loop graph capture failed: ") + cudaGetErrorString(e));
                return -1;
            }
            AAADMM_CUDA_OK(cudaGraphInstantiate(&g->exec, g->graph, 0));
            g->graph_key = key;
            g->body_launches = L + 1;
        }
    }
    AAADMM_CUDA_OK(cudaEventRecord(e0, st));
    if (g->variant == AAADMM_GEO_GS && gs_enqueue_warmup(g, launches)) return -1;
    if (max_iter > 0 && !no_graph) {
        AAADMM_CUDA_OK(cudaGraphLaunch(g->exec, st));
    } else if (max_iter > 0) {
        for (int turn = 0; turn < 4 * max_iter + 8; ++turn) {
            if (geo_enqueue_any(g, m, launches)) return -1;
            AAADMM_CUDA_OK(cudaMemcpyAsync(&hs, g->st, sizeof(SolveState), cudaMemcpyDeviceToHost, st));
            AAADMM_CUDA_OK(cudaStreamSynchronize(st));
            if (hs.iter >= max_iter) break;
        }
    }
    AAADMM_CUDA_OK(cudaEventRecord(e1, st));
    // the solution is default_x (ALMGeometrySolver.h:285) resp. current_x (GeometrySolver.h:265)
    AAADMM_CUDA_OK(cudaMemcpyAsync(x_out, (g->variant == AAADMM_GEO_GS ? g->Ubuf : g->Dbuf) + NU, sizeof(double) * 3 * g->P,
                                   cudaMemcpyDeviceToHost, st));
    AAADMM_CUDA_OK(cudaMemcpyAsync(&hs, g->st, sizeof(SolveState), cudaMemcpyDeviceToHost, st));
    AAADMM_CUDA_OK(cudaStreamSynchronize(st));
    if (hist && hs.iter > 0) AAADMM_CUDA_OK(cudaMemcpy(hist, g->hist, sizeof(double) * hs.iter, cudaMemcpyDeviceToHost));
    g->last_iters = hs.iter;
    g->last_max_iter = max_iter;
    if (res) {
        res->iters_logged = hs.iter;
        res->rejects = hs.n_rejects;
        res->broke_early = 0;
        cudaEventElapsedTime(&res->loop_ms, e0, e1);
        res->step_ms = res->loop_ms;
        res->kernel_launches = launches + hs.loop_it * g->body_launches;
    }
    return 0;
    API_TRY_END
}

int aaadmm_geo_reset_flags(aaadmm_geo *g, int *flags, int n) {
    API_TRY_BEGIN
    if (!g || !flags || n < 0 || n > g->last_iters) {
        set_last_error("geo_reset_flags: bad arguments (n must not exceed the iterations of the last solve)");
        return -1;
    }
    std::vector<double> f((size_t)std::max(n, 1));
    if (n > 0) AAADMM_CUDA_OK(cudaMemcpy(f.data(), g->hist + g->last_max_iter, sizeof(double) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) flags[i] = f[i] != 0.0;
    return 0;
    API_TRY_END
}

int aaadmm_geo_iteration_times(aaadmm_geo *g, double *ms, int n) {
    API_TRY_BEGIN
    if (!g || !ms || n < 0 || n > g->last_iters) {
        set_last_error("geo_iteration_times: bad arguments (n must not exceed the iterations of the last solve)");
        return -1;
    }
    if (n > 0) AAADMM_CUDA_OK(cudaMemcpy(ms, g->hist + 2 * (size_t)g->last_max_iter, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
    API_TRY_END
}

int aaadmm_geo_project(int type, int n, int k, const double *cols, const double *param4, double *out) {
    API_TRY_BEGIN
    if (aaadmm_device_count() <= 0) {
        set_last_error("no CUDA device (this library has no CPU path)");
        return -1;
    }
    const int kc = type == GEO_PLANE ? k : k - 1;
    std::vector<int> types(n, type), ptr(n + 1), idx((size_t)n * k, 0), col0(n);
    std::vector<double> prm((size_t)4 * n);
    for (int c = 0; c <= n; ++c) ptr[c] = c * k;
    for (int c = 0; c < n; ++c) {
        col0[c] = c * kc;
        for (int q = 0; q < 4; ++q) prm[4 * (size_t)c + q] = param4[q];
    }
    GeoConstraints C;
    int *dt = nullptr, *dp = nullptr, *di = nullptr, *dc = nullptr;
    double *dprm = nullptr, *dv = nullptr, *dz = nullptr;
    if (up(&dt, types.data(), n) || up(&dp, ptr.data(), n + 1) || up(&di, idx.data(), idx.size()) || up(&dc, col0.data(), n) ||
        up(&dprm, prm.data(), prm.size()) || up(&dv, cols, (size_t)3 * n * kc))
        return -1;
    AAADMM_CUDA_OK(cudaMalloc((void **)&dz, sizeof(double) * 3 * n * kc));
    C.n = n;
    C.type = dt;
    C.idx_ptr = dp;
    C.idx = di;
    C.col0 = dc;
    C.param = dprm;
    launch_geo_project_only(C, dv, dz);
    AAADMM_CUDA_OK(cudaGetLastError());
    AAADMM_CUDA_OK(cudaMemcpy(out, dz, sizeof(double) * 3 * n * kc, cudaMemcpyDeviceToHost));
    cudaFree(dt);
    cudaFree(dp);
    cudaFree(di);
    cudaFree(dc);
    cudaFree(dprm);
    cudaFree(dv);
    cudaFree(dz);
    return 0;
    API_TRY_END
}

int aaadmm_geo_closest_points(const double *verts, int nv, const int *tris, int nt, const double *q, int nq,
                              double *closest, int *tri) {
    API_TRY_BEGIN
    if (aaadmm_device_count() <= 0) {
        set_last_error("no CUDA device (this library has no CPU path)");
        return -1;
    }
    RefMeshDev M;
    if (M.upload(verts, nv, tris, nt)) return -1;
    double *dq = nullptr, *dc = nullptr;
    int *dt = nullptr;
    if (up(&dq, q, (size_t)3 * nq)) return -1;
    AAADMM_CUDA_OK(cudaMalloc((void **)&dc, sizeof(double) * 3 * std::max(nq, 1)));
    AAADMM_CUDA_OK(cudaMalloc((void **)&dt, sizeof(int) * std::max(nq, 1)));
    GeoSoft S;
    S.n = nq;
    S.point = nullptr;
    S.order = nullptr;
    S.weight = 1.0;
    S.nodes = M.nodes;
    S.tri_order = M.order;
    S.tri = M.tri;
    S.last_tri = nullptr;
    launch_geo_closest_only(S, dq, dc, dt);
    AAADMM_CUDA_OK(cudaGetLastError());
    AAADMM_CUDA_OK(cudaMemcpy(closest, dc, sizeof(double) * 3 * nq, cudaMemcpyDeviceToHost));
    if (tri) AAADMM_CUDA_OK(cudaMemcpy(tri, dt, sizeof(int) * nq, cudaMemcpyDeviceToHost));
    cudaFree(dq);
    cudaFree(dc);
    cudaFree(dt);
    M.release();
    return 0;
    API_TRY_END
}

}  // extern "C"
