// Multifrontal triangular solves for a sparse LDL^T factor resident in HBM.
//
// The factor's columns are grouped into FRONTS (supernodes: elimination-tree chains j -> j+1; the k
// rows below the chain's last column contain the rows of all its columns). Per front the setup
// builds ONE dense column-major matrix
//        M = [ Linv ]   ns x ns   inverse of the unit-lower diagonal block
//            [  Q   ]   k  x ns   Q = L_below * Linv
// so that both sweeps are plain dense products with no dependency inside a front:
//   forward   [ y ; u_own ] = M w          w = rhs - (updates of the children that land in the front's columns)
//             u = u_own + (updates of the children that land in the k rows below)      -> handed to the parent
//   backward  x = M^T [ D^-1 y ; -x(rows below) ]
// One kernel launch per sweep: the fronts are cut into tasks (row tiles forward, column chunks backward), the tasks
// are handed out in topological order through a ticket and synchronise through per-front arrival counters; every
// task streams its part of M through a cp.async.bulk + mbarrier ring (see k_fwd_front / k_bwd_front). No indices are
// streamed (8 bytes per factor entry and sweep), the summation order is fixed (no atomics on data): results are
// bit-reproducible. Right-hand sides are interleaved (n x NR, NR = 3 for A = Ahat (x) I3): every
// factor entry is read once per sweep and used NR times.
#include "ldlt_apply.cuh"
#include "ldlt_factor.cuh"
#include "pipe.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace aaadmm {

namespace {

constexpr int CTA = 256;
constexpr int FCH = 1024;  // columns of w staged in shared memory at a time (forward)
constexpr int BCH = 1024;  // rows of v staged in shared memory at a time (backward)

// Child updates that add into one front row: up to 4 slots inline (-1 = none), the rest (rare) in a CSR.
__device__ __forceinline__ SweepTask load_task(const SweepTask *t) {
    union {
        SweepTask s;
        int4 v[4];
    } u;
    const int4 *p = reinterpret_cast<const int4 *>(t);
#pragma unroll
    for (int i = 0; i < 4; ++i) u.v[i] = __ldg(p + i);
    return u.s;
}

struct Gather {
    const int4 *ell;
    const int64_t *ptr;  // null when no row has more than 4 contributions
    const int *idx;
};

template <int NR>
__device__ __forceinline__ void gather_overflow(const Gather &G, int64_t grow, const double *U, double sign, double (&a)[NR]) {
    const int64_t g1 = G.ptr[grow + 1];
    for (int64_t g = G.ptr[grow]; g < g1; ++g) {
        const size_t s = (size_t)G.idx[g] * NR;
#pragma unroll
        for (int q = 0; q < NR; ++q) a[q] += sign * __ldcg(U + s + q);
    }
}

template <int NR>
__device__ __forceinline__ void gather_add(const Gather &G, int64_t grow, const double *U, double sign, double (&a)[NR]) {
    const int4 e = __ldg(G.ell + grow);
    const int s4[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (s4[i] >= 0) {
            const size_t s = (size_t)s4[i] * NR;
#pragma unroll
            for (int q = 0; q < NR; ++q) a[q] += sign * __ldcg(U + s + q);
        }
    }
    if (G.ptr) gather_overflow<NR>(G, grow, U, sign, a);
}

// ---- bulk-copy pipeline primitives (cp.async.bulk + mbarrier, sm_90+) ----------------------------
// A 9th warp streams the CTA's part of the factor into a ring of NSTG stages of STG doubles in shared
// memory (one elected lane issues cp.async.bulk; the bytes signal the stage's `full` barrier); the 8
// consumer warps wait on `full`, use the stage and release it through `empty`. Loads in flight hold no
// registers and cost one instruction per copy (up to 16 KB).
constexpr int NCONS = CTA;          // consumer threads
constexpr int NTHR = CTA + 32;      // + producer warp
constexpr int STG = 2048;           // doubles per stage (16 KB)
constexpr int NSTG = 3;
constexpr int WCH = 256;            // wide fronts: columns / rows of the assembled vector per buffer
constexpr int NWB = 4;              // ... and the number of such buffers (NWB - 1 chunks in flight)
constexpr int ACOLS = 256;          // entries one assemble task handles

// Programmatic dependent launch: the next level's kernel may start (and fill its ring with factor data,
// which does not depend on the previous level) while this one drains; it must not touch anything a
// previous kernel wrote before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
// Dataflow between the tasks of one sweep kernel: a task starts its dependent reads when the counter of the
// fronts it needs has reached `need`, and bumps its own front's counter when its results are written.
// Tasks are handed out in topological order through an atomic ticket, so everything a task waits for is
// already running or finished: the wait cannot deadlock.
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void task_wait(const int *cnt, const SweepTask &F) {
    if (F.need > 0) {
        if (threadIdx.x == 0) {
            while (ld_acquire(cnt + F.wait_idx) < F.need) __nanosleep(40);
        }
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
    }
}
// called by all consumer threads after their global writes
__device__ __forceinline__ void task_signal(int *cnt, const SweepTask &F) {
    if (F.signal_idx < 0) return;  // nobody waits for this task (root forward, leaf fronts backward)
    // the CTA's writes are ordered before the barrier, the barrier before thread 0's release at device scope (the
    // semaphore idiom: one release instead of a device-wide fence in every thread); the waiting side reads the counter
    // with ld.acquire.gpu
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;\n" ::"l"(cnt + F.signal_idx) : "memory");
}

__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

struct PipeBars {
    uint64_t full[NSTG], empty[NSTG];
    uint64_t wfull[NWB];  // the buffers of the assembled vector of a wide front
};
__device__ __forceinline__ void pipe_init(PipeBars &B) {
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NSTG; ++i) {
            mbar_init(&B.full[i], 1);
            mbar_init(&B.empty[i], NCONS / 32);
        }
#pragma unroll
        for (int i = 0; i < NWB; ++i) mbar_init(&B.wfull[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
}

// skip flag set: drain the copies already issued (the CTA's shared memory must not be recycled under them)
__device__ __forceinline__ void pipe_drain(PipeBars &B, int issued) {
    for (int i = 0; i < issued; ++i) mbar_wait(&B.full[i], 0);
}

// rows a forward tile of 2^shape rows stores per column when `left` rows of the front remain at its first row: a tile
// that reaches past the front's last row (the only or last tile of a front, e.g. 176 rows in a 256-row tile) is stored
// - in HBM and in the stages - with its real (even) row count as column stride, so no padding rows are streamed
__host__ __device__ __forceinline__ int fwd_rows_stored(int shape, int left) {
    const int rt = 1 << shape, r = (left + 1) & ~1;
    return r < rt ? r : rt;
}

// ---- forward: one CTA = RT rows of one front, the columns interleaved over CS = 512 / RT slices ----
// RT = 256 .. 16 per task (narrow fronts want many rows per CTA, wide ones many slices). A consumer thread
// owns two adjacent rows. The tile is stored contiguously (tile-major copy Mf of the front matrices: column
// j of the tile at offset j * RT), so one stage = STG / RT columns = ONE bulk copy.
template <int NR, int LRT>
__device__ __forceinline__ void fwd_tile(const SweepTask &F, const Gather &G,
                                         const double *W, const double *__restrict__ dinv,
                                         double *__restrict__ Yd, double *U, int ws_cap, double *sm, PipeBars &B) {
    constexpr int RT = 1 << LRT, CS = (2 * CTA) >> LRT, NCS = STG / RT, PER = NCS / CS;
    static_assert(PER >= 1, "stage narrower than the slices");
    double *ring = sm;                         // [NSTG][STG]; reused by the final reduction
    double *ws = sm + (size_t)NSTG * STG;      // [ws_cap][NR]
    const int r0 = F.start, m = F.ns + F.k;
    const int ncols = min(F.ns, r0 + RT);      // columns stored for this tile (rows of the diagonal block
                                               // have nothing to the right of the tile's last row)
    const int nstages = (ncols + NCS - 1) / NCS;
    const int RTS = fwd_rows_stored(LRT, m - r0);  // column stride of the stored tile (== RT unless the tile ends the front)
    const int lane = threadIdx.x & 31;
    const int lp = threadIdx.x & (RT / 2 - 1), cs = threadIdx.x >> (LRT - 1);
    double acc[2][NR], pass[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) acc[0][q] = acc[1][q] = pass[q] = 0.0;
    // thread t < RT finishes row r0 + t; the children's updates of a row below the front are passed on to
    // the parent together with the front's own
    const int frow = r0 + (int)threadIdx.x;
    const bool fin = (int)threadIdx.x < RT && frow < m;
    if (fin && frow >= F.ns) gather_add<NR>(G, F.g_off + frow, U, 1.0, pass);
    int it = 0;
    // Wide fronts (F.cw != 0): w was assembled in place in W by the front's assemble tasks; it arrives in
    // chunks of WCH columns by bulk copy, NWB - 1 chunks in flight while one is used.
    const bool wide = F.cw != 0;
    const int chunk = wide ? WCH : ws_cap;
    constexpr int HALF = WCH * NR + 2;
    auto issue_w = [&](int c) {
        const int jc = c * WCH;
        const size_t idx = (size_t)(F.first + jc) * NR;
        const int sh = (int)(idx & 1);  // bulk copies want 16-byte aligned sources
        const unsigned bytes = (unsigned)(((min(ncols - jc, WCH) * NR + sh + 1) & ~1) * 8);
        mbar_expect_tx(&B.wfull[c % NWB], bytes);
        bulk_g2s(ws + (c % NWB) * HALF, W + idx - sh, bytes, &B.wfull[c % NWB]);
    };
    if (wide && threadIdx.x == 0) {
        for (int c = 0; c < NWB && c * WCH < ncols; ++c) issue_w(c);
    }
    for (int jc = 0, ci = 0; jc < ncols; jc += chunk, ++ci) {
        const int jn = min(ncols - jc, chunk);
        const double *wbase = ws;
        if (wide) {
            mbar_wait(&B.wfull[ci % NWB], (ci / NWB) & 1);
            wbase = ws + (ci % NWB) * HALF + (((size_t)(F.first + jc) * NR) & 1);
        } else {
        if (jc > 0) cons_sync();
        // w_j = rhs_j - sum of the child updates that land on column j (4 columns per thread in flight)
        for (int j0 = threadIdx.x; j0 < jn; j0 += 4 * NCONS) {
            double a[4][NR];
            int4 e[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = j0 + i * NCONS;
                if (j < jn) {
                    e[i] = __ldg(G.ell + F.g_off + jc + j);
#pragma unroll
                    for (int q = 0; q < NR; ++q) a[i][q] = W[(size_t)(F.first + jc + j) * NR + q];
                } else {
                    e[i] = make_int4(-1, -1, -1, -1);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int s4[4] = {e[i].x, e[i].y, e[i].z, e[i].w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (s4[c] >= 0) {
                        const size_t s = (size_t)s4[c] * NR;
#pragma unroll
                        for (int q = 0; q < NR; ++q) a[i][q] -= __ldcg(U + s + q);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = j0 + i * NCONS;
                if (j < jn) {
                    if (G.ptr) gather_overflow<NR>(G, F.g_off + jc + j, U, -1.0, a[i]);
#pragma unroll
                    for (int q = 0; q < NR; ++q) ws[j * NR + q] = a[i][q];
                }
            }
        }
        cons_sync();
        }
        // the stages whose columns fall into this chunk (chunks are multiples of NCS)
        const int je = jc + jn;
        for (; it * NCS < je; ++it) {
            const int slot = it % NSTG;
            mbar_wait(&B.full[slot], (it / NSTG) & 1);
            const double *st = ring + (size_t)slot * STG + 2 * lp;
            const double *wp = wbase + (size_t)(it * NCS - jc) * NR;
            const int colsin = ncols - it * NCS;
            if (colsin >= NCS) {
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int c = cs + i * CS;
                    const double2 v = *reinterpret_cast<const double2 *>(st + c * RTS);
#pragma unroll
                    for (int q = 0; q < NR; ++q) {
                        const double w = wp[c * NR + q];
                        acc[0][q] += v.x * w;
                        acc[1][q] += v.y * w;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int c = cs + i * CS;
                    if (c < colsin) {
                        const double2 v = *reinterpret_cast<const double2 *>(st + c * RTS);
#pragma unroll
                        for (int q = 0; q < NR; ++q) {
                            const double w = wp[c * NR + q];
                            acc[0][q] += v.x * w;
                            acc[1][q] += v.y * w;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&B.empty[slot]);
        }
        if (wide) {
            cons_sync();  // everyone is done with this half
            if (threadIdx.x == 0 && (ci + NWB) * WCH < ncols) issue_w(ci + NWB);
        }
    }
    cons_sync();  // every consumer is done with the ring
    double *part = ring;  // [CS][RT][NR], CS * RT = 512
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int q = 0; q < NR; ++q) part[(cs * RT + 2 * lp + e) * NR + q] = acc[e][q];
    cons_sync();
    if (fin) {
        double sum[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q) sum[q] = 0.0;
#pragma unroll
        for (int c = 0; c < CS; ++c) {
#pragma unroll
            for (int q = 0; q < NR; ++q) sum[q] += part[(c * RT + (int)threadIdx.x) * NR + q];
        }
        if (frow < F.ns) {
            const double di = dinv[F.first + frow];
#pragma unroll
            for (int q = 0; q < NR; ++q) Yd[(size_t)(F.first + frow) * NR + q] = sum[q] * di;
        } else {
            const size_t o = (size_t)(F.u_off + frow - F.ns) * NR;
#pragma unroll
            for (int q = 0; q < NR; ++q) U[o + q] = sum[q] + pass[q];
        }
    }
}

template <int NR>
__global__ void __launch_bounds__(NTHR, 3)
k_fwd_front(const SweepTask *__restrict__ tasks, int *ctl, const double *__restrict__ Mf,
            Gather G, double *W, const double *__restrict__ dinv, double *__restrict__ Yd, double *U,
            const int *skip, int ws_cap, unsigned long long *trace) {
    extern __shared__ __align__(128) double sm[];
    __shared__ PipeBars B;
    __shared__ int s_task;
    const unsigned long long t0 = trace ? gtime() : 0;
    if (threadIdx.x == 0) s_task = atomicAdd(ctl, 1);
    pipe_init(B);
    const SweepTask F = load_task(tasks + s_task);
    pdl_launch_dependents();
    // producer: the tile is one contiguous run of Mf; the first NSTG stages go out before the previous
    // level is known to be finished
    const int RT = 1 << F.shape, NCS = STG / RT;
    const int ncols = min(F.ns, F.start + RT);
    const int nstages = F.shape == 0 ? 0 : (ncols + NCS - 1) / NCS;  // shape 0: assemble task, no factor data
    const double *tile = Mf + F.m_off;
    const int RTS = fwd_rows_stored(F.shape, F.ns + F.k - F.start);
    auto produce = [&](int it) {
        const int slot = it % NSTG;
        mbar_wait(&B.empty[slot], ((it / NSTG) & 1) ^ 1);
        const unsigned bytes = (unsigned)(min(NCS, ncols - it * NCS) * RTS * 8);
        mbar_expect_tx(&B.full[slot], bytes);
        bulk_g2s(sm + (size_t)slot * STG, tile + (size_t)it * NCS * RTS, bytes, &B.full[slot]);
    };
    const int first = min(nstages, NSTG);
    if (threadIdx.x == NCONS)
        for (int it = 0; it < first; ++it) produce(it);
    pdl_wait();
    if (skip && *skip) {
        if (threadIdx.x == NCONS) pipe_drain(B, first);
        return;
    }
    if (threadIdx.x >= NCONS) {
        if (threadIdx.x == NCONS)
            for (int it = first; it < nstages; ++it) produce(it);
        return;
    }
    const unsigned long long t1 = trace ? gtime() : 0;
    if (F.shape == 0) {
        // assemble task of a wide front: w_j = rhs_j - (child updates landing on column j), in place in W. It sits on
        // the dependency chain of the sweep: what does not depend on the children (the slot list, the right-hand side)
        // is fetched before the wait.
        const int j = F.start + (int)threadIdx.x;
        const bool act = (int)threadIdx.x < ACOLS && j < F.ns;
        double a[NR];
        int4 e = make_int4(-1, -1, -1, -1);
        if (act) {
            e = __ldg(G.ell + F.g_off + j);
#pragma unroll
            for (int q = 0; q < NR; ++q) a[q] = __ldcg(W + (size_t)(F.first + j) * NR + q);
        }
        task_wait(ctl + 2, F);  // the children's updates are written
        if (act) {
            const int s4[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (s4[i] >= 0) {
                    const size_t so = (size_t)s4[i] * NR;
#pragma unroll
                    for (int q = 0; q < NR; ++q) a[q] += -1.0 * __ldcg(U + so + q);
                }
            }
            if (G.ptr) gather_overflow<NR>(G, F.g_off + j, U, -1.0, a);
#pragma unroll
            for (int q = 0; q < NR; ++q) W[(size_t)(F.first + j) * NR + q] = a[q];
        }
        task_signal(ctl + 2, F);
        if (trace && threadIdx.x == 0) {
            unsigned long long *r = trace + 4 * (size_t)s_task;
            r[0] = t0;
            r[1] = t1;
            r[2] = t1;
            r[3] = gtime();
        }
        return;
    }
    task_wait(ctl + 2, F);  // the children's updates are written
    const unsigned long long t2 = trace ? gtime() : 0;
    switch (F.shape) {
    case 8: fwd_tile<NR, 8>(F, G, W, dinv, Yd, U, ws_cap, sm, B); break;
    case 7: fwd_tile<NR, 7>(F, G, W, dinv, Yd, U, ws_cap, sm, B); break;
    case 6: fwd_tile<NR, 6>(F, G, W, dinv, Yd, U, ws_cap, sm, B); break;
    case 5: fwd_tile<NR, 5>(F, G, W, dinv, Yd, U, ws_cap, sm, B); break;
    case 4: fwd_tile<NR, 4>(F, G, W, dinv, Yd, U, ws_cap, sm, B); break;
    default: fwd_tile<NR, 3>(F, G, W, dinv, Yd, U, ws_cap, sm, B); break;
    }
    task_signal(ctl + 2, F);
    if (trace && threadIdx.x == 0) {
        unsigned long long *r = trace + 4 * (size_t)s_task;
        r[0] = t0;
        r[1] = t1;
        r[2] = t2;
        r[3] = gtime();
    }
}

// ---- backward: one CTA = `shape` columns of one front, 8 * CW at a time (one warp per CW columns, a lane
// owns two adjacent rows). A stage = RB rows of those 8 * CW columns, stored contiguously in Mb (the
// backward copy of the front matrices, stage blocks in the order they are used): ONE bulk copy. ----
template <int NR, int CW>
__device__ __forceinline__ void bwd_task(const SweepTask &F, const int *__restrict__ rows, const double *Yd, double *X,
                                         const double *Va, const int *__restrict__ perm, double *__restrict__ x_out,
                                         int v_cap, double *sm, PipeBars &B) {
    constexpr int NC = 8 * CW, RB = STG / NC;  // columns per pass, rows per stage
    double *ring = sm;                     // [NSTG][NC][RB]
    double *vs = sm + (size_t)NSTG * STG;  // v[v_cap][NR]
    const int c0 = F.start, m = F.ns + F.k;
    const int cend = min(F.ns, c0 + F.shape);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool single = (F.ld - c0) <= v_cap;  // v fits at once: staged once, several column passes allowed
    // Wide fronts (F.cw & 8): v was assembled once per front in Va by the front's assemble tasks and arrives
    // in chunks of WCH rows by bulk copy, NWB - 1 chunks in flight while one is used.
    const bool wide = (F.cw & 8) != 0;
    const int chunk = wide ? WCH : v_cap;
    constexpr int HALF = WCH * NR + 2;
    auto issue_v = [&](int c) {
        const int rc = c0 + c * WCH;
        const unsigned bytes = (unsigned)(((min(F.ld - rc, WCH) * NR + 1) & ~1) * 8);
        mbar_expect_tx(&B.wfull[c % NWB], bytes);
        bulk_g2s(vs + (c % NWB) * HALF, Va + (size_t)(F.u_off + rc) * NR, bytes, &B.wfull[c % NWB]);
    };
    if (wide && threadIdx.x == 0) {
        for (int c = 0; c < NWB && c0 + c * WCH < F.ld; ++c) issue_v(c);
    }
    int it = 0;
    for (int jb0 = c0; jb0 < cend; jb0 += NC) {
        const int jb = jb0 + warp * CW;
        const int rs = max(c0, jb0 & ~15);
        double acc[CW][NR];
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int q = 0; q < NR; ++q) acc[c][q] = 0.0;
        for (int rc = c0, ci = 0; rc < F.ld; rc += chunk, ++ci) {
            const int re = min(F.ld, rc + chunk);
            const double *vbase = vs;
            if (wide) {
                mbar_wait(&B.wfull[ci % NWB], (ci / NWB) & 1);
                vbase = vs + (ci % NWB) * HALF;
            } else if (!(single && jb0 > c0)) {
                if (rc > c0) cons_sync();
                // v = [ D^-1 y of the front's columns ; -x of the rows below ; 0 ] (4 rows per thread in flight)
                for (int r0 = rc + threadIdx.x; r0 < re; r0 += 4 * NCONS) {
                    size_t src[4];
                    double sg[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = min(r0 + i * NCONS, re - 1);
                        if (r < F.ns) {
                            src[i] = (size_t)(F.first + r) * NR;
                            sg[i] = 1.0;
                        } else if (r < m) {
                            src[i] = (size_t)__ldg(rows + F.g_off + r - F.ns) * NR;
                            sg[i] = -1.0;
                        } else {
                            src[i] = 0;
                            sg[i] = 0.0;
                        }
                    }
                    double vv[4][NR];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double *p = sg[i] > 0.0 ? Yd : X;
#pragma unroll
                        for (int q = 0; q < NR; ++q) vv[i][q] = __ldcg(p + src[i] + q);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = r0 + i * NCONS;
                        if (r < re) {
#pragma unroll
                            for (int q = 0; q < NR; ++q) vs[(r - rc) * NR + q] = sg[i] != 0.0 ? sg[i] * vv[i][q] : 0.0;  // padding rows: an exact zero, whatever X[0] holds
                        }
                    }
                }
                cons_sync();
            }
            // stages of this pass inside the chunk (chunk boundaries are stage boundaries)
            for (int rr = max(rs, rc); rr < re; rr += RB, ++it) {
                const int slot = it % NSTG;
                mbar_wait(&B.full[slot], (it / NSTG) & 1);
                const int nrow = min(RB, F.ld - rr);
                const double *st = ring + (size_t)slot * STG + (warp * CW) * nrow;  // block = [ncol][nrow]
                const double *vp = vbase + (size_t)(rr - rc) * NR;
#pragma unroll
                for (int t = 0; t < RB / 64; ++t) {
                    const int r = 64 * t + 2 * lane;
                    if (r < nrow) {
                        // rows r and r + 1 of v: 2 * NR consecutive doubles, 16-byte aligned (r is even) - read as
                        // 16-byte pieces: 8-byte loads at this 48-byte lane stride run into 4-way bank conflicts
                        double v0[NR], v1[NR];
                        {
                            double2 tq[NR];
                            const double2 *vp2 = reinterpret_cast<const double2 *>(vp + r * NR);
#pragma unroll
                            for (int q = 0; q < NR; ++q) tq[q] = vp2[q];
                            const double *tf = reinterpret_cast<const double *>(tq);
#pragma unroll
                            for (int q = 0; q < NR; ++q) {
                                v0[q] = tf[q];
                                v1[q] = tf[NR + q];
                            }
                        }
#pragma unroll
                        for (int c = 0; c < CW; ++c) {
                            const double2 a = *reinterpret_cast<const double2 *>(st + c * nrow + r);
#pragma unroll
                            for (int q = 0; q < NR; ++q) {
                                acc[c][q] += a.x * v0[q];
                                acc[c][q] += a.y * v1[q];
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&B.empty[slot]);
            }
            if (wide) {
                cons_sync();  // everyone is done with this half
                if (threadIdx.x == 0 && c0 + (ci + NWB) * WCH < F.ld) issue_v(ci + NWB);
            }
        }
#pragma unroll
        for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int q = 0; q < NR; ++q)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[c][q] += __shfl_xor_sync(0xffffffffu, acc[c][q], o);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                const int j = jb + c;
                if (j < cend) {
                    const int o = perm[F.first + j];
#pragma unroll
                    for (int q = 0; q < NR; ++q) {
                        X[(size_t)(F.first + j) * NR + q] = acc[c][q];
                        x_out[(size_t)o * NR + q] = acc[c][q];
                    }
                }
            }
        }
    }
}

template <int NR>
__global__ void __launch_bounds__(NTHR, 3)
k_bwd_front(const SweepTask *__restrict__ tasks, int *ctl, const double *__restrict__ Mb,
            const int *__restrict__ rows, const double *Yd, double *X, double *Va, const int *__restrict__ perm,
            double *__restrict__ x_out, const int *skip, int v_cap, unsigned long long *trace) {
    extern __shared__ __align__(128) double sm[];
    __shared__ PipeBars B;
    __shared__ int s_task;
    const unsigned long long t0 = trace ? gtime() : 0;
    if (threadIdx.x == 0) s_task = atomicAdd(ctl + 1, 1);
    pipe_init(B);
    const SweepTask F = load_task(tasks + s_task);
    pdl_launch_dependents();
    const int NC = max(8 * (F.cw & 7), 8), RB = STG / NC;  // cw == 0: assemble task (shape 0: no columns)
    const int c0 = F.start, cend = min(F.ns, c0 + F.shape);
    // producer state: the task's stage blocks lie one after the other in Mb, in the order they are used
    const double *psrc = Mb + F.m_off;
    int pit = 0, pjb0 = c0, prr = c0;  // next stage: pass pjb0, first row prr
    auto produce = [&](int limit) {
        while (pjb0 < cend && pit < limit) {
            const int slot = pit % NSTG;
            const unsigned bytes = (unsigned)(min(NC, cend - pjb0) * min(RB, F.ld - prr) * 8);
            mbar_wait(&B.empty[slot], ((pit / NSTG) & 1) ^ 1);
            mbar_expect_tx(&B.full[slot], bytes);
            bulk_g2s(sm + (size_t)slot * STG, psrc, bytes, &B.full[slot]);
            psrc += bytes / 8;
            ++pit;
            prr += RB;
            if (prr >= F.ld) {
                pjb0 += NC;
                prr = max(c0, pjb0 & ~15);
            }
        }
    };
    if (threadIdx.x == NCONS) produce(NSTG);
    pdl_wait();
    if (skip && *skip) {
        if (threadIdx.x == NCONS) pipe_drain(B, pit);
        return;
    }
    if (threadIdx.x >= NCONS) {
        if (threadIdx.x == NCONS) produce(0x7fffffff);
        return;
    }
    const unsigned long long t1 = trace ? gtime() : 0;
    if ((F.cw & 7) == 0) {
        // assemble task of a wide front: v = [ D^-1 y ; -x(rows below) ; 0 ] for rows start .. start + ACOLS. On the
        // dependency chain of the sweep: y (the forward sweep is complete) and the row ids are fetched before the wait.
        const int r = F.start + (int)threadIdx.x;
        const bool act = (int)threadIdx.x < ACOLS && r < F.ld;
        double a[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q) a[q] = 0.0;
        size_t xi = 0;
        const bool below = act && r >= F.ns && r < F.ns + F.k;
        if (act && r < F.ns) {
#pragma unroll
            for (int q = 0; q < NR; ++q) a[q] = __ldcg(Yd + (size_t)(F.first + r) * NR + q);
        } else if (below) {
            xi = (size_t)__ldg(rows + F.g_off + r - F.ns) * NR;
        }
        task_wait(ctl + 2, F);  // the parent's (hence every ancestor's) x is written
        if (below) {
#pragma unroll
            for (int q = 0; q < NR; ++q) a[q] = -__ldcg(X + xi + q);
        }
        if (act) {
#pragma unroll
            for (int q = 0; q < NR; ++q) Va[(size_t)(F.u_off + r) * NR + q] = a[q];
        }
        task_signal(ctl + 2, F);
        if (trace && threadIdx.x == 0) {
            unsigned long long *tr = trace + 4 * (size_t)s_task;
            tr[0] = t0;
            tr[1] = t1;
            tr[2] = t1;
            tr[3] = gtime();
        }
        return;
    }
    task_wait(ctl + 2, F);  // the parent's (hence every ancestor's) x is written
    const unsigned long long t2 = trace ? gtime() : 0;
    switch (F.cw & 7) {
    case 4: bwd_task<NR, 4>(F, rows, Yd, X, Va, perm, x_out, v_cap, sm, B); break;
    case 2: bwd_task<NR, 2>(F, rows, Yd, X, Va, perm, x_out, v_cap, sm, B); break;
    default: bwd_task<NR, 1>(F, rows, Yd, X, Va, perm, x_out, v_cap, sm, B); break;
    }
    task_signal(ctl + 2, F);
    if (trace && threadIdx.x == 0) {
        unsigned long long *r = trace + 4 * (size_t)s_task;
        r[0] = t0;
        r[1] = t1;
        r[2] = t2;
        r[3] = gtime();
    }
}

// Tile-major copy of the front matrices for the forward sweep: one CTA per forward task.
__global__ void __launch_bounds__(256)
k_make_tiles(const SweepTask *__restrict__ tasks, const int64_t *__restrict__ src_off, const double *__restrict__ M,
             double *__restrict__ Mf) {
    const SweepTask F = tasks[blockIdx.x];
    if (F.shape == 0) return;  // assemble task: no factor data
    const int RT = 1 << F.shape, r0 = F.start;
    const int ncols = min(F.ns, r0 + RT);
    const int RTS = fwd_rows_stored(F.shape, F.ns + F.k - r0);
    const double *src = M + src_off[blockIdx.x];
    double *dst = Mf + F.m_off;
    for (int e = threadIdx.x; e < ncols * RTS; e += 256) {
        const int j = e / RTS, r = r0 + (e - j * RTS);
        dst[e] = r < F.ld ? src[(size_t)j * F.ld + r] : 0.0;
    }
}

// Stage-major copy of the front matrices for the backward sweep: one CTA per backward task, walking the
// passes and stages exactly like the producer of k_bwd_front.
__global__ void __launch_bounds__(256)
k_make_btiles(const SweepTask *__restrict__ tasks, const int64_t *__restrict__ src_off, const double *__restrict__ M,
              double *__restrict__ Mb) {
    const SweepTask F = tasks[blockIdx.x];
    if ((F.cw & 7) == 0) return;  // assemble task: no factor data
    const int NC = 8 * (F.cw & 7), RB = STG / NC;
    const int c0 = F.start, cend = min(F.ns, c0 + F.shape);
    const double *src = M + src_off[blockIdx.x];
    double *dst = Mb + F.m_off;
    for (int jb0 = c0; jb0 < cend; jb0 += NC) {
        const int rs = max(c0, jb0 & ~15);
        const int ncol = min(NC, cend - jb0);
        for (int rr = rs; rr < F.ld; rr += RB) {
            const int nrow = min(RB, F.ld - rr);
            for (int e = threadIdx.x; e < ncol * nrow; e += 256) {
                const int c = e / nrow, r = e - c * nrow;
                dst[e] = src[(size_t)(jb0 + c) * F.ld + rr + r];
            }
            dst += ncol * nrow;
        }
    }
}

template <int NR>
__global__ void k_permute_in(const double *__restrict__ b, const int *__restrict__ perm, int n, double *__restrict__ W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = perm[i];
#pragma unroll
    for (int r = 0; r < NR; ++r) W[(size_t)i * NR + r] = b[(size_t)o * NR + r];
}

// ---- setup kernels ---------------------------------------------------------------------
// X = T^-1 for the unit-lower diagonal block of every front (T in A, X into M, both column-major with
// the front's ld). One thread per row i of X, from X T = I: X(i,j) = -sum_{k=j+1..i} X(i,k) T(k,j), j < i.
// M must be zero on entry.
__global__ void __launch_bounds__(128)
k_invert_fronts(const int2 *__restrict__ tasks, const FrontDesc *__restrict__ fronts, const double *__restrict__ A,
                double *__restrict__ M) {
    const int2 task = tasks[blockIdx.x];
    const FrontDesc F = fronts[task.x];
    const int i0 = task.y;
    const int i = i0 + threadIdx.x;
    const bool active = i < F.ns;
    const double *T = A + F.m_off;
    double *Xm = M + F.m_off;
    if (active) Xm[(size_t)i * F.ld + i] = 1.0;
    const int itop = min(i0 + 127, F.ns - 1);
    for (int j = itop - 1; j >= 0; --j) {
        if (!active || j >= i) continue;
        double s = 0.0;
        const double *Tj = T + (size_t)j * F.ld;
        for (int k = j + 1; k <= i; ++k) s += Xm[(size_t)k * F.ld + i] * Tj[k];
        Xm[(size_t)j * F.ld + i] = -s;
    }
}

// ---- blocked inverse of the unit-lower diagonal blocks (replaces the row-serial k_invert_fronts for everything but the
// 64 x 64 blocks on the diagonal): X = T^-1 by block distance r = i - j,
//     X(i,i) = T(i,i)^-1,      X(i,j) = -X(i,i) * sum_{k=j}^{i-1} T(i,k) X(k,j)   (i > j),
// every block at distance r depends on blocks of smaller distance in its own block column only: one launch per
// distance over all fronts, each CTA one 64 x 64 block (two chained products, 4 x 4 outputs per thread). The critical
// path of a 2,888-column front is 45 launches instead of 4 M dependent iterations of one thread.
constexpr int IB = 64;
__global__ void __launch_bounds__(IB)
k_invert_diag_blocks(const int2 *__restrict__ tasks, const FrontDesc *__restrict__ fronts, const double *__restrict__ A,
                     double *__restrict__ M) {
    __shared__ double Ts[IB][IB + 1];
    const int2 task = tasks[blockIdx.x];
    const FrontDesc F = fronts[task.x];
    const int d0 = task.y * IB, nb = min(IB, F.ns - d0);
    const double *T = A + F.m_off + (size_t)d0 * F.ld + d0;
    double *X = M + F.m_off + (size_t)d0 * F.ld + d0;
    const int i = threadIdx.x;
    for (int c = 0; c < nb; ++c) Ts[i][c] = (i < nb && i > c) ? T[(size_t)c * F.ld + i] : 0.0;
    __syncthreads();
    if (i >= nb) return;
    // row i of X from X T = I: X(i,j) = -sum_{k=j+1..i} X(i,k) T(k,j); the row lives in this thread's registers
    double x[IB];
#pragma unroll
    for (int j = 0; j < IB; ++j) x[j] = j == i ? 1.0 : 0.0;
#pragma unroll
    for (int j = IB - 2; j >= 0; --j) {
        if (j < i) {
            double s = 0.0;
#pragma unroll
            for (int k = j + 1; k < IB; ++k)
                if (k <= i) s += x[k] * Ts[k][j];
            x[j] = -s;
        }
    }
#pragma unroll
    for (int j = 0; j < IB; ++j)
        if (j <= i) X[(size_t)j * F.ld + i] = x[j];
}

__global__ void __launch_bounds__(256)
k_invert_offdiag(const int2 *__restrict__ tasks, int r, const FrontDesc *__restrict__ fronts, const double *__restrict__ A,
                 double *__restrict__ M) {
    extern __shared__ __align__(16) double ism[];
    double (*As)[IB + 1] = reinterpret_cast<double (*)[IB + 1]>(ism);                   // sum_k T(i,k) X(k,j)
    double (*Xd)[IB + 1] = reinterpret_cast<double (*)[IB + 1]>(ism + IB * (IB + 1));   // X(i,i); staging before
    double (*Ps)[IB + 1] = Xd;                                                          // [16][65] rows of T(i, .)
    double (*Ls)[IB + 1] = reinterpret_cast<double (*)[IB + 1]>(ism + IB * (IB + 1) + 16 * (IB + 1));  // [16][65] X(., j)
    const int2 task = tasks[blockIdx.x];
    const FrontDesc F = fronts[task.x];
    const int jb = task.y, ib = jb + r;
    const int i0 = ib * IB, j0 = jb * IB;
    const double *T = A + F.m_off;
    double *X = M + F.m_off;
    const int tr = threadIdx.x & 15, tj = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    for (int t0 = j0; t0 < i0; t0 += 16) {
        for (int e = threadIdx.x; e < 16 * IB; e += 256) {
            const int tt = e >> 6, q = e & 63;
            Ps[tt][q] = (i0 + q < F.ns) ? T[(size_t)(t0 + tt) * F.ld + i0 + q] : 0.0;   // T(i0 + q, t0 + tt)
        }
        for (int e = threadIdx.x; e < 16 * IB; e += 256) {
            const int c = e >> 4, tt = e & 15;
            Ls[tt][c] = X[(size_t)(j0 + c) * F.ld + t0 + tt];                          // X(t0 + tt, j0 + c)
        }
        __syncthreads();
#pragma unroll
        for (int tt = 0; tt < 16; ++tt) {
            double p[4], l[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) p[a] = Ps[tt][tr + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; ++b) l[b] = Ls[tt][tj + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += p[a] * l[b];
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) As[tr + 16 * a][tj + 16 * b] = acc[a][b];
    for (int e = threadIdx.x; e < IB * IB; e += 256) {
        const int c = e >> 6, q = e & 63;   // X(i0 + q, i0 + c)
        Xd[q][c] = (i0 + q < F.ns && i0 + c < F.ns && q >= c) ? X[(size_t)(i0 + c) * F.ld + i0 + q] : 0.0;
    }
    __syncthreads();
    // X(i,j) = -X(i,i) * As
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
#pragma unroll 8
    for (int t = 0; t < IB; ++t) {
        double p[4], l[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) p[a] = Xd[tr + 16 * a][t];
#pragma unroll
        for (int b = 0; b < 4; ++b) l[b] = As[t][tj + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += p[a] * l[b];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int q = i0 + tr + 16 * a, c = j0 + tj + 16 * b;
            if (q < F.ns) X[(size_t)c * F.ld + q] = -acc[a][b];
        }
}

// Q = P * Linv: P = rows ns.. of A, Linv = rows 0..ns of M, Q -> rows ns.. of M. 64 x 64 tile per CTA,
// 4 x 4 outputs per thread, 16-deep shared-memory stages.
__global__ void __launch_bounds__(256)
k_front_q(const int4 *__restrict__ tasks, const FrontDesc *__restrict__ fronts, const double *__restrict__ A,
          double *__restrict__ M) {
    __shared__ double Ps[16][64 + 1], Ls[16][64 + 1];
    const int4 task = tasks[blockIdx.x];
    const FrontDesc F = fronts[task.x];
    const int rt = task.y, jt = task.z;  // first row of the tile (within the k rows) / first column
    const double *P = A + F.m_off + F.ns;
    const double *Li = M + F.m_off;
    double *Q = M + F.m_off + F.ns;
    const int tr = threadIdx.x & 15, tj = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    // Linv(t, j) = 0 for t < j: start at the tile's first column
    for (int t0 = jt & ~15; t0 < F.ns; t0 += 16) {
        // Ps[tt][r] = P(rt + r, t0 + tt); Ls[tt][j] = Linv(t0 + tt, jt + j)
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int tt = e >> 6, r = e & 63;
            Ps[tt][r] = (t0 + tt < F.ns && rt + r < F.k) ? P[(size_t)(t0 + tt) * F.ld + rt + r] : 0.0;
        }
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {
            const int j = e >> 4, tt = e & 15;
            Ls[tt][j] = (t0 + tt < F.ns && jt + j < F.ns) ? Li[(size_t)(jt + j) * F.ld + t0 + tt] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int tt = 0; tt < 16; ++tt) {
            double p[4], l[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) p[a] = Ps[tt][tr + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; ++b) l[b] = Ls[tt][tj + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] += p[a] * l[b];
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int r = rt + tr + 16 * a, j = jt + tj + 16 * b;
            if (r < F.k && j < F.ns) Q[(size_t)j * F.ld + r] = acc[a][b];
        }
}

template <typename T>
int upload(T **dst, const std::vector<T> &src) {
    size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    AAADMM_CUDA_OK(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) AAADMM_CUDA_OK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

}  // namespace

void ldlt_dev_destroy(LdltDev *f) {
    if (!f) return;
    cudaFree(f->perm);
    cudaFree(f->iperm);
    cudaFree(f->dinv);
    cudaFree(f->fronts);
    cudaFree(f->M);
    cudaFree(f->Mf);
    cudaFree(f->Mb);
    cudaFree(f->rows);
    cudaFree(f->gell);
    cudaFree(f->gptr);
    cudaFree(f->gidx);
    cudaFree(f->tasks);
    cudaFree(f->W);
    cudaFree(f->Yd);
    cudaFree(f->X);
    cudaFree(f->U);
    cudaFree(f->ctl);
    cudaFree(f->Va);
    cudaFree(f->trace);
    cudaFree(f->inv_tasks);
    cudaFree(f->invd_tasks);
    cudaFree(f->invo_tasks);
    cudaFree(f->q_tasks);
    cudaFree(f->tile_src);
    cudaFree(f->A);
    cudaFree(f->D);
    factor_plan_destroy(f->plan);
    if (f->setup_stream) cudaStreamDestroy(f->setup_stream);
    delete f;
}

namespace {
// [Linv ; Q] of every front from its [T ; P] (f->A) and the two sweep-ordered copies Mf / Mb; on `s`.
int finish_numeric(LdltDev *f, cudaStream_t s) {
    AAADMM_CUDA_OK(cudaMemsetAsync(f->M, 0, (size_t)std::max<int64_t>(f->m_tot, 1) * sizeof(double), s));
    static const bool serial_inverse = getenv("AAADMM_INV_SERIAL") != nullptr;  // the row-serial kernel (experiments)
    if (serial_inverse) {
        if (f->n_inv_tasks > 0) k_invert_fronts<<<f->n_inv_tasks, 128, 0, s>>>(f->inv_tasks, f->fronts, f->A, f->M);
    } else {
        if (f->n_invd_tasks > 0) k_invert_diag_blocks<<<f->n_invd_tasks, IB, 0, s>>>(f->invd_tasks, f->fronts, f->A, f->M);
        const size_t ism = sizeof(double) * 2 * IB * (IB + 1);
        for (size_t r = 1; r < f->invo_off.size(); ++r) {
            const int cnt = f->invo_off[r] - f->invo_off[r - 1];
            if (cnt > 0) k_invert_offdiag<<<cnt, 256, ism, s>>>(f->invo_tasks + f->invo_off[r - 1], (int)r, f->fronts, f->A, f->M);
        }
    }
    if (f->n_q_tasks > 0) k_front_q<<<f->n_q_tasks, 256, 0, s>>>(f->q_tasks, f->fronts, f->A, f->M);
    if (f->n_ftasks > 0) k_make_tiles<<<f->n_ftasks, 256, 0, s>>>(f->tasks, f->tile_src, f->M, f->Mf);
    if (f->n_btasks > 0) k_make_btiles<<<f->n_btasks, 256, 0, s>>>(f->tasks + f->n_ftasks, f->tile_src + f->n_ftasks, f->M, f->Mb);
    AAADMM_CUDA_OK(cudaGetLastError());
    return 0;
}
}  // namespace

// Structure from the pattern of L; values either from (Lx, D) (a factor computed elsewhere) or - when the pattern of
// the matrix (Ap, Ai) is given instead - left to the device-side factorisation (f->plan).
static int build(LdltDev **out, int n, const int64_t *Lp, const int *Li, const double *Lx, const double *D, const int *perm,
                 int nrhs, const int64_t *Ap, const int *Ai) {
    if (nrhs != 1 && nrhs != 3) {
        set_last_error("ldlt: nrhs must be 1 or 3");
        return -1;
    }
    LdltDev *f = new LdltDev();
    f->n = n;
    f->nrhs = nrhs;
    const int64_t nnz = Lp[n];

    // ---- fronts. (1) Small subtrees of the elimination tree whose columns are consecutive (the leaf domains
    // of a nested-dissection ordering) become ONE front each: the rows below the subtree's root contain the
    // rows of all its columns that leave the subtree, the rest is stored as explicit zeros. Without this a
    // surface mesh (SURVEY cfg 3) decomposes into tens of thousands of fronts of 3 columns. (2) Elsewhere
    // column j joins the front of j-1 when j is the parent of j-1 and either the patterns nest exactly
    // (fundamental supernode) or the front is still small (relaxed chain).
    // tuning knobs (environment overrides are for experiments only)
    // Factors with little fill per column (surface meshes, 2-D separators: cfg 3 has nnz(L) / n = 39, a 2 M-triangle
    // cloth 58; the volume meshes of cfg 4 / cfg 5 513 / 279) are deep trees of small fronts; every level costs about 5 us
    // per sweep whatever it holds, so they merge twice as much (measured: cfg 3 0.876 -> 0.836 ms per loop turn, cloth
    // apply 0.899 -> 0.868 ms; on the volume meshes the larger caps only add zeros: cfg 4 0.481 -> 0.488 ms).
    const int merge_cap = nnz < (int64_t)128 * std::max(n, 1) ? 128 : 64;
    const int kSmall = env_int("AAADMM_KSMALL", merge_cap), kCap = env_int("AAADMM_KCAP", 6144);
    const int kSubtree = env_int("AAADMM_KSUBTREE", merge_cap);
    std::vector<int> sub_root(std::max(n, 1), -1);  // root of the maximal small contiguous subtree holding j
    if (kSubtree > 1) {
        std::vector<int> par(n, -1), size(n, 1), lo(n);
        for (int j = 0; j < n; ++j) {
            lo[j] = j;
            if (Lp[j + 1] > Lp[j]) par[j] = Li[Lp[j]];
        }
        for (int j = 0; j < n; ++j)  // children precede their parents
            if (par[j] >= 0) {
                size[par[j]] += size[j];
                lo[par[j]] = std::min(lo[par[j]], lo[j]);
            }
        auto small = [&](int j) { return size[j] <= kSubtree && j - lo[j] + 1 == size[j]; };
        for (int j = n - 1; j >= 0; --j) {
            if (sub_root[j] >= 0 || !small(j) || size[j] < 2) continue;
            if (par[j] >= 0 && small(par[j])) continue;  // the parent's subtree will take it
            for (int c = lo[j]; c <= j; ++c) sub_root[c] = j;
        }
    }
    // (3) A column whose pattern is ALMOST that of the chain below it (it has a second child, so its pattern is a union,
    // typically one or two rows more) also joins when the explicit zeros this costs stay below relax_pct percent of the
    // front: otherwise it becomes a front of one column between two big fronts - two more hops of the dependency chain
    // in every sweep for a few KB of data.
    const int relax_pct = env_int("AAADMM_RELAX_PCT", 5);
    std::vector<int> blk_of(std::max(n, 1)), blk_first;
    int64_t front_zeros = 0;  // explicit zeros the relaxed joins have put into the current front
    for (int j = 0; j < n; ++j) {
        bool join = false;
        if (j > 0) {
            if (sub_root[j] >= 0) {
                join = sub_root[j - 1] == sub_root[j];  // same subtree; a subtree always starts a front
            } else {
                const int64_t c0 = Lp[j] - Lp[j - 1], c1 = Lp[j + 1] - Lp[j];
                const bool chain = c0 > 0 && Li[Lp[j - 1]] == j;
                const int cur_size = j - blk_first.back();
                if (chain && cur_size < kCap) {
                    const int64_t extra = std::max<int64_t>(c1 + 1 - c0, 0) * cur_size;  // zeros this join adds
                    const int64_t entries = (int64_t)(cur_size + 1) * (c1 + (cur_size + 2) / 2);
                    if (c0 == c1 + 1 || cur_size < kSmall) {
                        join = true;
                    } else if (sub_root[j - 1] < 0 && (front_zeros + extra) * 100 <= (int64_t)relax_pct * entries) {
                        join = true;
                        front_zeros += extra;
                    }
                }
            }
        }
        if (!join) {
            blk_first.push_back(j);
            front_zeros = 0;
        }
        blk_of[j] = (int)blk_first.size() - 1;
    }
    const int nb = (int)blk_first.size();
    blk_first.push_back(n);
    f->n_blocks = nb;

    std::vector<FrontDesc> fr(std::max(nb, 1));
    std::vector<int> parent(nb, -1), level(nb, 0);
    int64_t m_tot = 0, r_tot = 0, g_rows = 0;
    int max_block = 0;
    int64_t dense_entries = 0;
    for (int b = 0; b < nb; ++b) {
        FrontDesc &F = fr[b];
        const int jl = blk_first[b + 1] - 1;
        F.first = blk_first[b];
        F.ns = blk_first[b + 1] - blk_first[b];
        F.k = (int)(Lp[jl + 1] - Lp[jl]);
        F.ld = (F.ns + F.k + 3) & ~3;
        F.m_off = m_tot;
        F.r_off = r_tot;
        F.u_off = r_tot;
        F.g_off = g_rows;
        F.pad0 = F.pad1 = 0;
        m_tot += ((int64_t)F.ld * F.ns + 31) & ~(int64_t)31;
        r_tot += F.k;
        g_rows += F.ns + F.k;
        max_block = std::max(max_block, F.ns);
        dense_entries += (int64_t)F.k * F.ns + (int64_t)F.ns * (F.ns + 1) / 2;
        if (F.k > 0) parent[b] = blk_of[Li[Lp[jl]]];
    }
    for (int b = 0; b < nb; ++b)
        if (parent[b] >= 0) level[parent[b]] = std::max(level[parent[b]], level[b] + 1);
    int nlev = 0;
    for (int b = 0; b < nb; ++b) nlev = std::max(nlev, level[b] + 1);
    f->n_levels = nlev;
    f->n_launches = 2;

    // ---- rows below each front, front matrices A = [T ; P] (column-major), child update slots ----
    std::vector<int> rows((size_t)std::max<int64_t>(r_tot, 1));
    for (int b = 0; b < nb; ++b) {
        const int jl = blk_first[b + 1] - 1;
        std::copy(Li + Lp[jl], Li + Lp[jl + 1], rows.begin() + fr[b].r_off);
    }
    std::vector<double> A;
    int64_t noff = 0;
    if (Lx) {
    try {
        A.assign((size_t)std::max<int64_t>(m_tot, 1), 0.0);
    } catch (const std::bad_alloc &) {
        set_last_error("ldlt: out of host memory for the front matrices");
        delete f;
        return -1;
    }
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : bad, noff)
    for (int b = 0; b < nb; ++b) {
        const FrontDesc &F = fr[b];
        double *Ab = A.data() + F.m_off;
        const int *R = rows.data() + F.r_off;
        for (int j = F.first; j < F.first + F.ns; ++j) {
            double *col = Ab + (size_t)(j - F.first) * F.ld;
            col[j - F.first] = 1.0;
            int pos = 0;
            for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) {
                const int i = Li[p];
                if (i < F.first + F.ns) {
                    col[i - F.first] = Lx[p];
                    continue;
                }
                while (pos < F.k && R[pos] < i) ++pos;
                if (pos >= F.k || R[pos] != i) {
                    ++bad;
                    break;
                }
                col[F.ns + pos] = Lx[p];
                ++noff;
            }
        }
    }
    if (bad) {
        set_last_error("ldlt: column patterns of the factor do not nest along the elimination tree");
        delete f;
        return -1;
    }
    } else {
        for (int b = 0; b < nb; ++b) {  // pattern-only statistics
            const FrontDesc &F = fr[b];
            for (int j = F.first; j < F.first + F.ns; ++j)
                for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) noff += Li[p] >= F.first + F.ns;
        }
    }
    // gather lists: for every front row (columns first, then the rows below) the slots of the children's
    // update vectors that add into it, children in ascending order
    std::vector<int64_t> gptr((size_t)g_rows + 1, 0);
    auto target_row = [&](const FrontDesc &P, int i, int &pos) {  // row of global index i inside front P
        if (i < P.first + P.ns) return i - P.first;
        const int *R = rows.data() + P.r_off;
        while (pos < P.k && R[pos] < i) ++pos;
        return P.ns + pos;
    };
    for (int b = 0; b < nb; ++b) {
        if (parent[b] < 0) continue;
        const FrontDesc &C = fr[b], &P = fr[parent[b]];
        int pos = 0;
        for (int q = 0; q < C.k; ++q) gptr[P.g_off + target_row(P, rows[C.r_off + q], pos) + 1]++;
    }
    // the first 4 contributions of a row sit inline (ELL), the rest in the CSR
    bool overflow = false;
    for (int64_t i = 0; i < g_rows; ++i) {
        const int64_t c = gptr[i + 1];
        if (c > 4) overflow = true;
        gptr[i + 1] = gptr[i] + std::max<int64_t>(c - 4, 0);
    }
    std::vector<int> gidx((size_t)std::max<int64_t>(gptr[g_rows], 1));
    std::vector<int4> gell((size_t)std::max<int64_t>(g_rows, 1), make_int4(-1, -1, -1, -1));
    {
        std::vector<int64_t> fill(gptr.begin(), gptr.end() - 1);
        for (int b = 0; b < nb; ++b) {
            if (parent[b] < 0) continue;
            const FrontDesc &C = fr[b], &P = fr[parent[b]];
            int pos = 0;
            for (int q = 0; q < C.k; ++q) {
                const int64_t grow = P.g_off + target_row(P, rows[C.r_off + q], pos);
                const int slot = (int)(C.u_off + q);
                int4 &e = gell[grow];
                if (e.x < 0) e.x = slot;
                else if (e.y < 0) e.y = slot;
                else if (e.z < 0) e.z = slot;
                else if (e.w < 0) e.w = slot;
                else gidx[fill[grow]++] = slot;
            }
        }
    }

    // ---- schedule: tile shapes and task lists ----
    // Forward: a CTA takes RT = 2^lrt rows of a front and splits the columns over 512 / RT slices: narrow
    // fronts get tall tiles (long CTAs amortise their fixed latency), wide ones many slices. Backward: a
    // CTA takes `ncols` columns, each warp CW of them at a time. Tree levels with few fronts are cut finer
    // until they have at least min_ctas tasks. One launch per sweep: the tasks are listed in topological
    // order (forward bottom-up, backward top-down) and synchronise through per-front arrival counters.
    const bool wide_on = env_int("AAADMM_WIDE", 1) != 0;
    const int min_ctas = env_int("AAADMM_MIN_CTAS", 148);  // one CTA per SM; each keeps 48 KB of loads in flight
    const int tile_entries = env_int("AAADMM_TILE_ENTRIES", 65536);
    const int min_lrt = env_int("AAADMM_MIN_LRT", 3);  // smallest forward tile: 8 rows
    const int min_ctas_f = env_int("AAADMM_MIN_CTAS_F", min_ctas);
    const int min_ctas_wide_f = env_int("AAADMM_MIN_CTAS_WIDE_F", min_ctas_f), min_ctas_wide_b = env_int("AAADMM_MIN_CTAS_WIDE_B", min_ctas);
    std::vector<std::vector<int>> by_level(nlev);
    for (int b = 0; b < nb; ++b) by_level[level[b]].push_back(b);
    std::vector<SweepTask> tasks;
    std::vector<int> bcw(std::max(nlev, 1), 4);
    auto make_task = [&](int b, bool fwd, int start, int shape) {
        const FrontDesc &F = fr[b];
        SweepTask t;
        t.m_off = F.m_off;
        t.g_off = fwd ? F.g_off : F.r_off;
        t.u_off = F.u_off;
        t.first = F.first;
        t.ns = F.ns;
        t.k = F.k;
        t.ld = F.ld;
        t.start = start;
        t.shape = shape;
        t.wait_idx = 0;
        t.need = 0;
        t.signal_idx = -1;
        t.cw = 4;
        return t;
    };
    auto ceil_log2 = [](int v) {
        int l = 0;
        while ((1 << l) < v) ++l;
        return l;
    };
    int max_ns_all = 1, max_ld_all = 4;
    for (int b = 0; b < nb; ++b) {
        max_ns_all = std::max(max_ns_all, fr[b].ns);
        max_ld_all = std::max(max_ld_all, fr[b].ld);
    }
    // chunk capacities (the environment overrides, multiples of 128 / 256 not above the defaults, are for tests)
    const int ws_cap = std::min((max_ns_all + 127) & ~127, std::min(FCH, env_int("AAADMM_FCH", FCH)));
    const int v_cap = std::min((max_ld_all + 255) & ~255, std::min(BCH, env_int("AAADMM_BCH", BCH)));
    f->ws_cap = ws_cap;
    f->v_cap = v_cap;
    std::vector<int> lrt(std::max(nb, 1), 8), bcols(std::max(nb, 1), 32), bcw_f(std::max(nb, 1), 4);
    const int task_slots = env_int("AAADMM_TASK_SLOTS", 200);                       // 0: no per-front cap
    const int task_slots_f = env_int("AAADMM_TASK_SLOTS_F", task_slots), task_slots_b = env_int("AAADMM_TASK_SLOTS_B", task_slots);
    const int64_t task_min_entries = (int64_t)env_int("AAADMM_TASK_MIN_KENTRIES", 8) * 1024;
    for (int l = 0; l < nlev; ++l) {
        std::vector<int> &v = by_level[l];
        std::sort(v.begin(), v.end(), [&](int a, int b2) {
            const int64_t wa = (int64_t)fr[a].ns * (fr[a].ns + fr[a].k), wb = (int64_t)fr[b2].ns * (fr[b2].ns + fr[b2].k);
            return wa != wb ? wa > wb : a < b2;
        });
        for (int b : v) {
            const int ns = fr[b].ns;
            lrt[b] = ns <= 128 ? 8 : (ns <= 256 ? 7 : (ns <= 512 ? 6 : (ns <= 1024 ? 5 : 4)));
        }
        auto count_f = [&]() {
            int64_t c = 0;
            for (int b : v) c += (fr[b].ns + fr[b].k + (1 << lrt[b]) - 1) >> lrt[b];
            return c;
        };
        // levels of wide fronts (their vector arrives by bulk copy: splitting costs no redundant gathers) are cut
        // finer than the others
        bool wide_f = false, wide_b = false;
        for (int b : v) {
            wide_f = wide_f || (wide_on && fr[b].ns > ws_cap);
            wide_b = wide_b || (wide_on && fr[b].ld > v_cap);
        }
        const int target_f = wide_f ? min_ctas_wide_f : min_ctas_f, target_b = wide_b ? min_ctas_wide_b : min_ctas;
        for (int pass = 0; pass < 6 && count_f() < target_f; ++pass)
            for (int b : v)
                if (lrt[b] > min_lrt && fr[b].ns >= 4 * ((2 * CTA) >> (lrt[b] - 1))) lrt[b]--;
        int cap = 128;
        auto set_cols = [&](int cw) {
            int64_t c = 0;
            for (int b : v) {
                const int m = fr[b].ns + fr[b].k, g = 8 * cw;
                int nc = g;
                if (fr[b].ld <= v_cap) nc = std::max(g, std::min(cap, tile_entries / std::max(m, 1)) / g * g);
                bcols[b] = nc;
                c += (fr[b].ns + nc - 1) / nc;
            }
            return c;
        };
        while (set_cols(bcw[l]) < target_b) {
            if (cap > 8 * bcw[l])
                cap /= 2;
            else if (bcw[l] > 1)
                bcw[l] /= 2;
            else
                break;
        }
        // Per-front cap on the size of a task. A task streams its part of the factor at the rate one CTA's ring sustains
        // (about 28 GB/s), and near the root a level is one to eight big fronts whose tasks all start together: the level
        // then lasts as long as its biggest task (the per-task trace shows 15-25 us for 45-90 k entries per task there).
        // A level's entries are therefore spread over about `slots` tasks (what the device holds at once), never
        // below e_min entries per task (the fixed cost of a task is a few us).
        {
            int64_t level_entries = 0;
            for (int b : v) level_entries += (int64_t)fr[b].ns * (fr[b].ns / 2 + fr[b].k);
            const int64_t e_star = std::max<int64_t>(task_min_entries, level_entries / std::max(task_slots_f, 1));
            const int64_t e_star_b = std::max<int64_t>(task_min_entries, level_entries / std::max(task_slots_b, 1));
            for (int b : v) {
                const int ns = fr[b].ns, m = fr[b].ns + fr[b].k;
                if (task_slots > 0) {
                    while (lrt[b] > min_lrt && ((int64_t)ns << lrt[b]) > e_star && ns >= 4 * ((2 * CTA) >> (lrt[b] - 1))) lrt[b]--;
                }
                int cw = bcw[l];
                if (task_slots > 0) {
                    const int64_t nc_des = std::max<int64_t>(8, e_star_b / std::max(m, 1));
                    cw = std::min(cw, nc_des >= 32 ? 4 : (nc_des >= 16 ? 2 : 1));
                    const int g = 8 * cw;
                    if (fr[b].ld <= v_cap)
                        bcols[b] = std::max(g, (int)std::min<int64_t>(bcols[b], nc_des) / g * g);
                    else
                        bcols[b] = g;
                }
                bcw_f[b] = cw;
            }
        }
    }
    // counters: [2 + b] forward arrivals at front b (tiles of its children), [2 + nb + b] backward arrivals
    // of front b (its own column chunks)
    std::vector<int> ntiles_f(std::max(nb, 1), 0), ntasks_b(std::max(nb, 1), 0), need_f(std::max(nb, 1), 0);
    // Emission order = ticket order. By tree level, or (default) by the estimated time a front becomes ready
    // (children's ready time + a latency/bandwidth estimate of their duration): tasks then take their tickets
    // roughly in the order they can start, and fewer CTA slots are held by tasks that only wait.
    // Measured (profiles/r02_ldlt_schedule_sweeps.log): with wide fronts near the root (volume meshes: their assemble steps
    // and slot-limited streams are what the duration estimate models worst) plain level order is 2.5 % faster, on the
    // deep trees of small fronts of a surface mesh the ready-time order is 1.5 % faster.
    const bool by_ready = env_int("AAADMM_ORDER_BY_LEVEL", (wide_on && max_ns_all > ws_cap) ? 1 : 0) == 0;
    const double est_lat = 0.1 * env_int("AAADMM_EST_LAT_TENTHS_US", 60), est_rate = env_int("AAADMM_EST_RATE_GBS", 25);
    auto front_dur_us = [&](int b, bool fwd) {
        const double m = fr[b].ns + fr[b].k;
        const double tile_bytes = fwd ? 8.0 * (double)(1 << lrt[b]) * fr[b].ns : 8.0 * m * std::min(fr[b].ns, bcols[b]);
        return est_lat + tile_bytes / (est_rate * 1e3) + ((fwd ? fr[b].ns > ws_cap : fr[b].ld > v_cap) ? est_lat : 0.0);
    };
    std::vector<double> ready_f(std::max(nb, 1), 0.0), ready_b(std::max(nb, 1), 0.0);
    for (int b = 0; b < nb; ++b)  // children have smaller indices than their parents
        if (parent[b] >= 0) ready_f[parent[b]] = std::max(ready_f[parent[b]], ready_f[b] + front_dur_us(b, true));
    for (int b = nb - 1; b >= 0; --b)
        if (parent[b] >= 0) ready_b[b] = ready_b[parent[b]] + front_dur_us(parent[b], false);
    std::vector<int> order_f, order_b;
    for (int l = 0; l < nlev; ++l) order_f.insert(order_f.end(), by_level[l].begin(), by_level[l].end());
    for (int l = nlev - 1; l >= 0; --l) order_b.insert(order_b.end(), by_level[l].begin(), by_level[l].end());
    if (by_ready) {
        std::stable_sort(order_f.begin(), order_f.end(), [&](int a, int b2) { return ready_f[a] < ready_f[b2]; });
        std::stable_sort(order_b.begin(), order_b.end(), [&](int a, int b2) { return ready_b[a] < ready_b[b2]; });
    }
    {
        for (int b : order_f) {
            const int m = fr[b].ns + fr[b].k;
            // wide fronts: assemble tasks write w once; the tiles wait for them and receive w by bulk copy
            const bool wide = wide_on && fr[b].ns > ws_cap;
            int n_asm = 0;
            if (wide)
                for (int j0 = 0; j0 < fr[b].ns; j0 += ACOLS, ++n_asm) {
                    SweepTask t = make_task(b, true, j0, 0);
                    t.wait_idx = b;
                    t.need = -1;
                    t.signal_idx = 2 * nb + b;
                    t.cw = 0;
                    tasks.push_back(t);
                }
            std::vector<SweepTask> tiles;
            for (int r0 = 0; r0 < m;) {
                const int shape = std::min(lrt[b], std::max(4, ceil_log2(m - r0)));  // tail tiles stay >= 16 rows
                SweepTask t = make_task(b, true, r0, shape);
                t.wait_idx = wide ? 2 * nb + b : b;
                t.need = wide ? n_asm : -1;  // -1: filled in below from the children's tile counts
                t.cw = wide ? 1 : 0;
                t.signal_idx = parent[b];
                tiles.push_back(t);
                ntiles_f[b]++;
                r0 += 1 << shape;
            }
            // longest first: a tile of the diagonal block only has the columns up to its last row
            std::stable_sort(tiles.begin(), tiles.end(), [](const SweepTask &x, const SweepTask &y) {
                const int64_t wx = (int64_t)std::min(x.ns, x.start + (1 << x.shape)) << x.shape;
                const int64_t wy = (int64_t)std::min(y.ns, y.start + (1 << y.shape)) << y.shape;
                return wx > wy;
            });
            tasks.insert(tasks.end(), tiles.begin(), tiles.end());
        }
    }
    f->n_ftasks = (int)tasks.size();
    for (int b = 0; b < nb; ++b)
        if (parent[b] >= 0) need_f[parent[b]] += ntiles_f[b];
    for (int i = 0; i < f->n_ftasks; ++i)
        if (tasks[i].need < 0) tasks[i].need = need_f[tasks[i].wait_idx];
    // tile-major copy for the forward sweep: the tile of a task is contiguous (column j at j * RT)
    std::vector<int64_t> tile_src(tasks.size());
    int64_t mf_tot = 0;
    for (size_t i = 0; i < tasks.size(); ++i) {
        SweepTask &t = tasks[i];
        tile_src[i] = t.m_off;
        t.m_off = mf_tot;
        const int64_t RT = (int64_t)1 << t.shape;
        mf_tot += ((int64_t)fwd_rows_stored(t.shape, t.ns + t.k - t.start) * std::min(t.ns, t.start + (int)RT) + 15) & ~(int64_t)15;
    }
    // stage-major copy for the backward sweep
    int64_t mb_tot = 0;
    int64_t va_tot = 0;
    std::vector<int> n_asm_b(std::max(nb, 1), -1);
    {
        for (int b : order_b) {
            const int l = level[b];
            // wide fronts: assemble tasks write v once into Va; the column tasks receive it by bulk copy
            const bool wide = wide_on && fr[b].ld > v_cap;
            int64_t va_off = 0;
            if (wide) {
                va_off = va_tot;
                va_tot += (fr[b].ld + 1) & ~1;
                n_asm_b[b] = 0;
                for (int r0 = 0; r0 < fr[b].ld; r0 += ACOLS, ++n_asm_b[b]) {
                    SweepTask t = make_task(b, false, r0, 0);
                    t.cw = 0;
                    t.u_off = va_off;
                    t.wait_idx = nb + std::max(parent[b], 0);
                    t.need = -1;  // the parent's column chunks, filled in below
                    t.signal_idx = 3 * nb + b;
                    tile_src.push_back(0);
                    tasks.push_back(t);
                }
            }
            for (int c0 = 0; c0 < fr[b].ns; c0 += bcols[b]) {
                SweepTask t = make_task(b, false, c0, bcols[b]);
                t.cw = bcw_f[b] | (wide ? 8 : 0);
                t.u_off = va_off;
                t.wait_idx = wide ? 3 * nb + b : nb + std::max(parent[b], 0);
                t.need = wide ? n_asm_b[b] : -1;
                t.signal_idx = nb + b;
                tile_src.push_back(t.m_off);
                t.m_off = mb_tot;
                const int NC = 8 * bcw_f[b], cend = std::min(t.ns, c0 + bcols[b]);
                for (int jb0 = c0; jb0 < cend; jb0 += NC)
                    mb_tot += (int64_t)std::min(NC, cend - jb0) * (t.ld - std::max(c0, jb0 & ~15));
                mb_tot = (mb_tot + 15) & ~(int64_t)15;
                tasks.push_back(t);
                ntasks_b[b]++;
            }
        }
    }
    f->n_btasks = (int)tasks.size() - f->n_ftasks;
    for (size_t i = f->n_ftasks; i < tasks.size(); ++i) {
        if (tasks[i].need >= 0) continue;
        const int b = tasks[i].signal_idx >= 3 * nb ? tasks[i].signal_idx - 3 * nb : tasks[i].signal_idx - nb;
        tasks[i].need = parent[b] >= 0 ? ntasks_b[parent[b]] : 0;
    }
    {
        // leaf fronts have no children: nothing waits for their backward tasks (1,168 of 6,000 tasks at cfg 4, the tail of
        // the sweep): no fence, no barrier, no atomic at their end
        std::vector<char> has_child(std::max(nb, 1), 0);
        for (int b = 0; b < nb; ++b)
            if (parent[b] >= 0) has_child[parent[b]] = 1;
        for (size_t i = f->n_ftasks; i < tasks.size(); ++i) {
            const int si = tasks[i].signal_idx;
            if (si >= nb && si < 2 * nb && !has_child[si - nb]) tasks[i].signal_idx = -1;
        }
    }
    f->n_ctl = 2 + 4 * nb;  // + one 'assembled' counter per front and sweep (wide fronts)
    if (getenv("AAADMM_LDLT_TRACE")) {  // developer aid: per-task time stamps (start, ring primed, dependencies met, done)
        cudaMalloc((void **)&f->trace, sizeof(unsigned long long) * 4 * tasks.size());
        cudaMemset(f->trace, 0, sizeof(unsigned long long) * 4 * tasks.size());
        f->host_tasks = tasks;
        f->host_level.assign(tasks.size(), 0);
        for (size_t i = 0; i < tasks.size(); ++i) f->host_level[i] = level[blk_of[tasks[i].first]];
    }

    std::vector<int> permv(perm, perm + n), iperm(std::max(n, 1));
    for (int k = 0; k < n; ++k) iperm[perm[k]] = k;
    std::vector<double> dinv(std::max(n, 1), 1.0);
    if (D)
        for (int k = 0; k < n; ++k) dinv[k] = 1.0 / D[k];

    // ---- upload ----
    int rc = 0;
    rc |= upload(&f->perm, permv);
    rc |= upload(&f->iperm, iperm);
    rc |= upload(&f->dinv, dinv);
    rc |= upload(&f->fronts, fr);
    rc |= upload(&f->rows, rows);
    rc |= upload(&f->gell, gell);
    if (overflow) {
        rc |= upload(&f->gptr, gptr);
        rc |= upload(&f->gidx, gidx);
    }
    rc |= upload(&f->tasks, tasks);
    if (Lx) rc |= upload(&f->A, A);
    std::vector<double>().swap(A);
    f->m_tot = m_tot;
    f->mf_tot = mf_tot;
    f->mb_tot = mb_tot;
    const size_t mat_bytes = (size_t)std::max<int64_t>(m_tot, 1) * sizeof(double);
    const size_t vec_bytes = (std::max<size_t>((size_t)n * nrhs, 1) + 2) * sizeof(double);  // + slack for aligned bulk copies
    const size_t u_bytes = std::max<size_t>((size_t)r_tot * nrhs, 1) * sizeof(double);
    if (rc || (!Lx && cudaMalloc((void **)&f->A, mat_bytes) != cudaSuccess) ||
        (!Lx && cudaMalloc((void **)&f->D, sizeof(double) * std::max(n, 1)) != cudaSuccess) ||
        cudaMalloc((void **)&f->M, mat_bytes) != cudaSuccess || cudaMalloc((void **)&f->W, vec_bytes) != cudaSuccess ||
        cudaMalloc((void **)&f->Yd, vec_bytes) != cudaSuccess || cudaMalloc((void **)&f->X, vec_bytes) != cudaSuccess ||
        cudaMalloc((void **)&f->U, u_bytes) != cudaSuccess || cudaMalloc((void **)&f->ctl, sizeof(int) * f->n_ctl) != cudaSuccess ||
        cudaMalloc((void **)&f->Va, (std::max<size_t>((size_t)va_tot * nrhs, 1) + 2) * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&f->Mf, (size_t)std::max<int64_t>(mf_tot, 1) * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&f->Mb, (size_t)std::max<int64_t>(mb_tot, 1) * sizeof(double)) != cudaSuccess) {
        set_last_error("ldlt: cudaMalloc failed");
        ldlt_dev_destroy(f);
        return -1;
    }
    // nothing the sweeps read is ever uninitialised (padding entries are multiplied by zeros of the factor: they must
    // be finite)
    cudaMemset(f->W, 0, vec_bytes);
    cudaMemset(f->Yd, 0, vec_bytes);
    cudaMemset(f->X, 0, vec_bytes);
    cudaMemset(f->U, 0, u_bytes);
    cudaMemset(f->Va, 0, (std::max<size_t>((size_t)va_tot * nrhs, 1) + 2) * sizeof(double));
    // ---- task lists of the numeric setup ([Linv ; Q], sweep-ordered copies) ----
    {
        std::vector<int2> inv_tasks;
        std::vector<int4> q_tasks;
        for (int b = 0; b < nb; ++b) {
            for (int i0 = 0; i0 < fr[b].ns; i0 += 128) inv_tasks.push_back(make_int2(b, i0));
            for (int rt = 0; rt < fr[b].k; rt += 64)
                for (int jt = 0; jt < fr[b].ns; jt += 64) q_tasks.push_back(make_int4(b, rt, jt, 0));
        }
        f->n_inv_tasks = (int)inv_tasks.size();
        f->n_q_tasks = (int)q_tasks.size();
        // blocked inverse: diagonal blocks, then the off-diagonal blocks by block distance
        std::vector<int2> invd, invo;
        int max_nbk = 0;
        for (int b = 0; b < nb; ++b) {
            const int nbk = (fr[b].ns + IB - 1) / IB;
            max_nbk = std::max(max_nbk, nbk);
            for (int d = 0; d < nbk; ++d) invd.push_back(make_int2(b, d));
        }
        f->invo_off.assign(1, 0);
        for (int r = 1; r < max_nbk; ++r) {
            for (int b = 0; b < nb; ++b) {
                const int nbk = (fr[b].ns + IB - 1) / IB;
                for (int j = 0; j + r < nbk; ++j) invo.push_back(make_int2(b, j));
            }
            f->invo_off.push_back((int)invo.size());
        }
        f->n_invd_tasks = (int)invd.size();
        cudaFuncSetAttribute(k_invert_offdiag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 2 * IB * (IB + 1)));
        if (upload(&f->invd_tasks, invd) || upload(&f->invo_tasks, invo)) {
            ldlt_dev_destroy(f);
            return -1;
        }
        if (upload(&f->inv_tasks, inv_tasks) || upload(&f->q_tasks, q_tasks) || upload(&f->tile_src, tile_src)) {
            ldlt_dev_destroy(f);
            return -1;
        }
    }
    if (Lx) {
        // values given: [Linv ; Q] and the sweep-ordered copies now; the intermediate matrices are not kept
        cudaError_t e = finish_numeric(f, 0) == 0 ? cudaDeviceSynchronize() : cudaErrorUnknown;
        cudaFree(f->M);
        f->M = nullptr;
        cudaFree(f->A);
        f->A = nullptr;
        if (e != cudaSuccess) {
            set_last_error(std::string("ldlt: front setup failed: ") + cudaGetErrorString(e));
            ldlt_dev_destroy(f);
            return -1;
        }
    } else {
        std::vector<int> parent_v(parent), level_v(level);
        if (cudaStreamCreate(&f->setup_stream) != cudaSuccess ||
            factor_plan_build(&f->plan, n, fr, nb, rows, parent_v, level_v, blk_of, perm, Ap, Ai, m_tot)) {
            ldlt_dev_destroy(f);
            return -1;
        }
    }
    const int max_smem = NSTG * STG * (int)sizeof(double) + (BCH + 256) * 3 * (int)sizeof(double);
    cudaFuncSetAttribute(k_fwd_front<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(k_fwd_front<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(k_bwd_front<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(k_bwd_front<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    if (getenv("AAADMM_LDLT_VERBOSE")) {
        int64_t small = 0;
        for (int b = 0; b < nb; ++b) small += fr[b].ns + fr[b].k <= 64;
        fprintf(stderr, "ldlt: n %d nnz %lld fronts %d (<=64 rows: %lld) levels %d max front %d | fwd tasks %d bwd tasks %d | Mf %.1f MB Mb %.1f MB | "
                "entries: L %.2f M, dense fronts (relaxed zeros incl.) %.2f M, forward tiles %.2f M, backward stages %.2f M\n",
                n, (long long)nnz, nb, (long long)small, nlev, max_block, f->n_ftasks, f->n_btasks, mf_tot * 8e-6, mb_tot * 8e-6,
                nnz * 1e-6, dense_entries * 1e-6, mf_tot * 1e-6, mb_tot * 1e-6);
    }
    f->stats.n = n;
    f->stats.n_blocks = nb;
    f->stats.n_levels = nlev;
    f->stats.max_block = max_block;
    f->stats.nnz_L = nnz;
    f->stats.nnz_offdiag = noff;
    f->stats.nnz_diag_dense = nnz - noff;
    f->stats.dense_entries = dense_entries;
    // per apply: every factor value once per sweep (8 bytes, no indices); rhs in, y out and in, x out twice,
    // the update vectors out and in
    f->stats.bytes_per_solve = 2.0 * 8.0 * (double)nnz + 8.0 * nrhs * (5.0 * (double)n + 3.0 * (double)r_tot);
    *out = f;
    return 0;
}

int ldlt_dev_create(LdltDev **out, int n, const int64_t *Lp, const int *Li, const double *Lx,
                    const double *D, const int *perm, int nrhs) {
    if (!Lx && Lp[n] > 0) {
        set_last_error("ldlt: null values");
        return -1;
    }
    static const double one = 1.0;
    return build(out, n, Lp, Li, Lx ? Lx : &one, D, perm, nrhs, nullptr, nullptr);
}

int ldlt_dev_refactor(LdltDev *f, const double *Ax, cudaStream_t stream) {
    if (!f->plan) {
        set_last_error("ldlt refactor: this factor was not created from a matrix");
        return -1;
    }
    static const bool trace = getenv("AAADMM_SETUP_TRACE") != nullptr;  // setup telemetry: device times of the numeric phases
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (trace)
        for (auto &e : ev) cudaEventCreate(&e);
    if (trace) cudaEventRecord(ev[0], stream);
    AAADMM_CUDA_OK(cudaMemcpyAsync(factor_plan_values(f->plan), Ax, sizeof(double) * (size_t)factor_plan_nnz(f->plan),
                                   cudaMemcpyHostToDevice, stream));
    if (trace) cudaEventRecord(ev[1], stream);
    if (factor_plan_run(f->plan, f->A, f->D, f->dinv, stream)) return -1;
    if (trace) cudaEventRecord(ev[2], stream);
    if (finish_numeric(f, stream)) return -1;
    if (trace) cudaEventRecord(ev[3], stream);
    const int rc = factor_plan_check(f->plan, stream);
    if (trace) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        cudaEventElapsedTime(&c, ev[2], ev[3]);
        fprintf(stderr, "[setup]   device numeric: matrix H2D %.2f ms, multifrontal LDL^T %.2f ms (%d launches), [Linv ; Q] + sweep copies %.2f ms\n",
                a, b, factor_plan_launches(f->plan), c);
        for (auto &e : ev) cudaEventDestroy(e);
    }
    return rc;
}

int ldlt_dev_create_from_matrix(LdltDev **out, int n, const int64_t *Ap, const int *Ai, const double *Ax, const int64_t *Lp,
                                const int *Li, const int *perm, int nrhs) {
    LdltDev *f = nullptr;
    if (build(&f, n, Lp, Li, nullptr, nullptr, perm, nrhs, Ap, Ai)) return -1;
    if (ldlt_dev_refactor(f, Ax, f->setup_stream)) {
        ldlt_dev_destroy(f);
        return -1;
    }
    *out = f;
    return 0;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait above).
template <typename... KArgs, typename... Args>
static cudaError_t launch_sweep(void (*kern)(KArgs...), int grid, size_t smem, cudaStream_t s, Args... args) {
    static const bool no_pdl = getenv("AAADMM_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NTHR);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int NR>
static int apply_impl(LdltDev *f, double *x_out, cudaStream_t s, const int *skip) {
    if (f->n_ftasks == 0) return 0;
    AAADMM_CUDA_OK(cudaMemsetAsync(f->ctl, 0, sizeof(int) * f->n_ctl, s));
    Gather G;
    G.ell = f->gell;
    G.ptr = f->gptr;
    G.idx = f->gidx;
    const size_t ring = (size_t)NSTG * STG * sizeof(double);
    AAADMM_CUDA_OK(launch_sweep(k_fwd_front<NR>, f->n_ftasks, ring + (std::max<size_t>((size_t)f->ws_cap * NR, NWB * (WCH * NR + 2)) + 8) * sizeof(double), s, f->tasks, f->ctl,
                                f->Mf, G, f->W, f->dinv, f->Yd, f->U, skip, f->ws_cap, f->trace));
    AAADMM_CUDA_OK(launch_sweep(k_bwd_front<NR>, f->n_btasks, ring + (std::max<size_t>((size_t)f->v_cap * NR, NWB * (WCH * NR + 2)) + 8) * sizeof(double), s,
                                f->tasks + f->n_ftasks, f->ctl, f->Mb, f->rows, f->Yd, f->X, f->Va, f->perm, x_out, skip, f->v_cap,
                                f->trace ? f->trace + 4 * (size_t)f->n_ftasks : nullptr));
    AAADMM_CUDA_OK(cudaGetLastError());
    return 0;
}

int ldlt_dev_dump_trace(LdltDev *f, const char *path) {
    if (!f->trace) {
        set_last_error("ldlt trace: set AAADMM_LDLT_TRACE before creating the factor");
        return -1;
    }
    const size_t nt = f->host_tasks.size();
    std::vector<unsigned long long> t(4 * nt);
    AAADMM_CUDA_OK(cudaMemcpy(t.data(), f->trace, sizeof(unsigned long long) * 4 * nt, cudaMemcpyDeviceToHost));
    FILE *fp = fopen(path, "w");
    if (!fp) {
        set_last_error("ldlt trace: cannot open the output file");
        return -1;
    }
    fprintf(fp, "task,sweep,level,first,ns,k,start,shape,cw,need,wait_idx,signal_idx,t_start,t_primed,t_deps,t_done\n");
    for (size_t i = 0; i < nt; ++i) {
        const SweepTask &k = f->host_tasks[i];
        fprintf(fp, "%zu,%s,%d,%d,%d,%d,%d,%d,%d,%d,%d,%d,%llu,%llu,%llu,%llu\n", i, i < (size_t)f->n_ftasks ? "fwd" : "bwd", f->host_level[i],
                k.first, k.ns, k.k, k.start, k.shape, k.cw, k.need, k.wait_idx, k.signal_idx, t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
    }
    fclose(fp);
    return 0;
}

int ldlt_dev_apply_permuted(LdltDev *f, double *x_out, cudaStream_t stream, const int *skip) {
    return f->nrhs == 3 ? apply_impl<3>(f, x_out, stream, skip) : apply_impl<1>(f, x_out, stream, skip);
}

int ldlt_dev_apply(LdltDev *f, const double *b, double *x_out, cudaStream_t stream, const int *skip) {
    const int g = (f->n + 255) / 256;
    if (f->n > 0) {
        if (f->nrhs == 3)
            k_permute_in<3><<<g, 256, 0, stream>>>(b, f->perm, f->n, f->W);
        else
            k_permute_in<1><<<g, 256, 0, stream>>>(b, f->perm, f->n, f->W);
    }
    return ldlt_dev_apply_permuted(f, x_out, stream, skip);
}

}  // namespace aaadmm
