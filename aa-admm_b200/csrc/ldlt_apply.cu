// Level-scheduled block triangular solves for a sparse LDL^T factor resident in HBM.
//
// Layout (all built once at setup from the CSC factor the host hands over):
//   * columns are grouped into BLOCKS = maximal elimination-tree chains j -> j+1 whose column
//     patterns nest (fundamental supernodes; short non-nesting chains are merged too);
//   * the unit-lower DIAGONAL block of every group is inverted once (FP64, on the GPU) and kept
//     dense, row-major (forward) and transposed (backward): inside a group the solve is a dense
//     GEMV with no sequential dependency;
//   * the OFF-BLOCK entries are kept twice: CSR (forward sweep, one warp gathers one row: no
//     atomics, fixed summation order) and CSC (backward sweep, one warp per column);
//   * groups are scheduled by their level in the group dependency DAG: all groups of one level
//     run in one launch.  Per level: off-block kernel, then diagonal-block kernel.
// Right-hand sides are interleaved (n x NR, NR = 3 for the xyz-Kronecker system A = Ahat (x) I3):
// every factor entry is read once per apply and used NR times.
#include "ldlt_apply.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace aaadmm {

namespace {

constexpr int WARPS_PER_CTA = 8;

template <int NR>
__device__ __forceinline__ void warp_sum(double (&a)[NR]) {
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
    }
}

// Work distribution shared by the four sweep kernels: the first n_long CTAs take one LONG row
// each (all 8 warps stride over it, fixed-shape combine through shared memory); the remaining
// CTAs take 8 SHORT rows each, one warp per row. Either way one row is summed by one fixed set
// of lanes in a fixed order: the result is deterministic and needs no atomics.
struct RowLists {
    const int *long_rows;   // one CTA per row
    int n_long;
    const int *short_rows;  // one warp per row
    int n_short;
    const int *tiny_rows;   // 8 lanes per row (rows of <= TINY_ROW entries: the leaf levels)
    int n_tiny;
};
constexpr int TINY_LANES = 8;
constexpr int TINY_PER_CTA = WARPS_PER_CTA * 32 / TINY_LANES;

__host__ __device__ __forceinline__ int short_ctas(const RowLists &L) { return (L.n_short + WARPS_PER_CTA - 1) / WARPS_PER_CTA; }

// cls: 0 long, 1 short, 2 tiny. Inactive lanes of a tiny CTA keep running (their warp still shuffles).
template <int NR>
__device__ __forceinline__ bool pick_row(const RowLists &L, int &row, int &tid, int &nthr, int &cls) {
    const int b = (int)blockIdx.x;
    if (b < L.n_long) {
        cls = 0;
        row = L.long_rows[b];
        tid = threadIdx.x;
        nthr = WARPS_PER_CTA * 32;
        return true;
    }
    const int ns = short_ctas(L);
    if (b < L.n_long + ns) {
        cls = 1;
        const int wid = (b - L.n_long) * WARPS_PER_CTA + (threadIdx.x >> 5);
        tid = threadIdx.x & 31;
        nthr = 32;
        if (wid >= L.n_short) return false;
        row = L.short_rows[wid];
        return true;
    }
    cls = 2;
    const int gid = (b - L.n_long - ns) * TINY_PER_CTA + (threadIdx.x / TINY_LANES);
    tid = threadIdx.x & (TINY_LANES - 1);
    nthr = TINY_LANES;
    if (gid >= L.n_tiny) return false;
    row = L.tiny_rows[gid];
    return true;
}

// Reduces acc over the lanes that worked on the row; returns true in the one thread holding it.
template <int NR>
__device__ __forceinline__ bool row_reduce(int cls, bool active, double (&acc)[NR]) {
    if (cls == 2) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
#pragma unroll
            for (int o = TINY_LANES / 2; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        }
        return active && (threadIdx.x & (TINY_LANES - 1)) == 0;
    }
    warp_sum<NR>(acc);
    if (cls == 1) return active && (threadIdx.x & 31) == 0;
    __shared__ double s_part[WARPS_PER_CTA][NR];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) s_part[warp][r] = acc[r];
    }
    __syncthreads();
    if (threadIdx.x != 0) return false;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        double x = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS_PER_CTA; ++w) x += s_part[w][r];
        acc[r] = x;
    }
    return true;
}

// W[row] -= sum_{off-block cols c} L(row,c) * Y[c]
template <int NR>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_fwd_off(RowLists L, const int64_t *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val,
          const double *__restrict__ Y, double *__restrict__ W, const int *skip) {
    if (skip && *skip) return;
    int row = 0, tid, nthr, cls;
    const bool active = pick_row<NR>(L, row, tid, nthr, cls);
    double acc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = 0.0;
    if (active) {
        const int64_t p0 = ptr[row], p1 = ptr[row + 1];
#pragma unroll 4
        for (int64_t p = p0 + tid; p < p1; p += nthr) {
            const int c = __ldg(col + p);
            const double v = __ldg(val + p);
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[r] += v * Y[(size_t)c * NR + r];
        }
    }
    if (row_reduce<NR>(cls, active, acc)) {
#pragma unroll
        for (int r = 0; r < NR; ++r) W[(size_t)row * NR + r] -= acc[r];
    }
}

// Y[row] = sum_{j<=i} Linv(i,j) W[first+j]     (dense inverse of the unit-lower diagonal block)
template <int NR>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_fwd_diag(RowLists L, const int *__restrict__ blk_of, const int *__restrict__ blk_first,
           const int64_t *__restrict__ linv_off, const double *__restrict__ Linv, const double *__restrict__ W,
           double *__restrict__ Y, const int *skip) {
    if (skip && *skip) return;
    int row = 0, tid, nthr, cls;
    const bool active = pick_row<NR>(L, row, tid, nthr, cls);
    double acc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = 0.0;
    if (active) {
        const int b = blk_of[row];
        const int first = blk_first[b];
        const int ns = blk_first[b + 1] - first;
        const int i = row - first;
        const double *Lrow = Linv + linv_off[b] + (size_t)i * ns;
#pragma unroll 4
        for (int j = tid; j <= i; j += nthr) {
            const double v = __ldg(Lrow + j);
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[r] += v * W[(size_t)(first + j) * NR + r];
        }
    }
    if (row_reduce<NR>(cls, active, acc)) {
#pragma unroll
        for (int r = 0; r < NR; ++r) Y[(size_t)row * NR + r] = acc[r];
    }
}

// W[j] = Y[j]/D[j] - sum_{off-block rows i} L(i,j) X[i]
template <int NR>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_bwd_off(RowLists L, const int64_t *__restrict__ ptr, const int *__restrict__ rowidx, const double *__restrict__ val,
          const double *__restrict__ dinv, const double *__restrict__ Y, const double *__restrict__ X,
          double *__restrict__ W, const int *skip) {
    if (skip && *skip) return;
    int j = 0, tid, nthr, cls;
    const bool active = pick_row<NR>(L, j, tid, nthr, cls);
    double acc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = 0.0;
    if (active) {
        const int64_t p0 = ptr[j], p1 = ptr[j + 1];
#pragma unroll 4
        for (int64_t p = p0 + tid; p < p1; p += nthr) {
            const int i = __ldg(rowidx + p);
            const double v = __ldg(val + p);
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[r] += v * X[(size_t)i * NR + r];
        }
    }
    if (row_reduce<NR>(cls, active, acc)) {
        const double di = dinv[j];
#pragma unroll
        for (int r = 0; r < NR; ++r) W[(size_t)j * NR + r] = Y[(size_t)j * NR + r] * di - acc[r];
    }
}

// X[row] = sum_{j>=i} LinvT(i,j) W[first+j]; also scatter to the caller's ordering.
template <int NR>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_bwd_diag(RowLists L, const int *__restrict__ blk_of, const int *__restrict__ blk_first,
           const int64_t *__restrict__ linv_off, const double *__restrict__ LinvT, const double *__restrict__ W,
           double *__restrict__ X, const int *__restrict__ perm, double *__restrict__ x_out, const int *skip) {
    if (skip && *skip) return;
    int row = 0, tid, nthr, cls;
    const bool active = pick_row<NR>(L, row, tid, nthr, cls);
    double acc[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[r] = 0.0;
    if (active) {
        const int b = blk_of[row];
        const int first = blk_first[b];
        const int ns = blk_first[b + 1] - first;
        const int i = row - first;
        const double *Lrow = LinvT + linv_off[b] + (size_t)i * ns;
#pragma unroll 4
        for (int j = i + tid; j < ns; j += nthr) {
            const double v = __ldg(Lrow + j);
#pragma unroll
            for (int r = 0; r < NR; ++r) acc[r] += v * W[(size_t)(first + j) * NR + r];
        }
    }
    if (row_reduce<NR>(cls, active, acc)) {
        const int o = perm[row];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            X[(size_t)row * NR + r] = acc[r];
            x_out[(size_t)o * NR + r] = acc[r];
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
// X = T^-1 for every dense unit-lower diagonal block; one thread per column of X.
__global__ void __launch_bounds__(128)
k_invert_unit_lower(const int2 *__restrict__ tasks, const int *__restrict__ blk_first,
                    const int64_t *__restrict__ linv_off, const double *__restrict__ T, double *__restrict__ X) {
    const int2 task = tasks[blockIdx.x];
    const int b = task.x, col0 = task.y;
    const int first = blk_first[b];
    const int ns = blk_first[b + 1] - first;
    const int j = col0 + threadIdx.x;
    const bool active = j < ns;
    const double *Tb = T + linv_off[b];
    double *Xb = X + linv_off[b];
    if (active) {
        for (int i = 0; i < j; ++i) Xb[(size_t)i * ns + j] = 0.0;
        Xb[(size_t)j * ns + j] = 1.0;
    }
    for (int i = col0 + 1; i < ns; ++i) {
        double s = 0.0;
        if (active && i > j) {
            const double *Trow = Tb + (size_t)i * ns;
            for (int k = col0; k < i; ++k) s += Trow[k] * Xb[(size_t)k * ns + j];
            Xb[(size_t)i * ns + j] = -s;
        }
    }
}

__global__ void __launch_bounds__(128)
k_transpose_blocks(const int2 *__restrict__ tasks, const int *__restrict__ blk_first,
                   const int64_t *__restrict__ linv_off, const double *__restrict__ X, double *__restrict__ XT) {
    const int2 task = tasks[blockIdx.x];
    const int b = task.x, col0 = task.y;
    const int first = blk_first[b];
    const int ns = blk_first[b + 1] - first;
    const int j = col0 + threadIdx.x;
    if (j >= ns) return;
    const double *Xb = X + linv_off[b];
    double *Tb = XT + linv_off[b];
    for (int i = 0; i < ns; ++i) Tb[(size_t)i * ns + j] = Xb[(size_t)j * ns + i];
}

template <typename T>
int upload(T **dst, const std::vector<T> &src) {
    size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    AAADMM_CUDA_OK(cudaMalloc((void **)dst, bytes));
    if (!src.empty()) AAADMM_CUDA_OK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace

void ldlt_dev_destroy(LdltDev *f) {
    if (!f) return;
    cudaFree(f->perm);
    cudaFree(f->iperm);
    cudaFree(f->blk_of);
    cudaFree(f->blk_first);
    cudaFree(f->linv_off);
    cudaFree(f->Linv);
    cudaFree(f->LinvT);
    cudaFree(f->dinv);
    cudaFree(f->fr_ptr);
    cudaFree(f->fr_col);
    cudaFree(f->fr_val);
    cudaFree(f->bc_ptr);
    cudaFree(f->bc_row);
    cudaFree(f->bc_val);
    cudaFree(f->lev_rows);
    cudaFree(f->W);
    cudaFree(f->Y);
    cudaFree(f->X);
    delete f;
}

int ldlt_dev_create(LdltDev **out, int n, const int64_t *Lp, const int *Li, const double *Lx,
                    const double *D, const int *perm, int nrhs) {
    if (nrhs != 1 && nrhs != 3) {
        set_last_error("ldlt: nrhs must be 1 or 3");
        return -1;
    }
    LdltDev *f = new LdltDev();
    f->n = n;
    f->nrhs = nrhs;
    const int64_t nnz = Lp[n];

    // ---- block partition: elimination-tree chains with nested patterns ----
    // tuning knobs (environment overrides are for experiments only)
    auto env_int = [](const char *name, int dflt) { const char *v = getenv(name); return v ? atoi(v) : dflt; };
    const int kSmall = env_int("AAADMM_KSMALL", 96), kCap = 6144;
    std::vector<int> blk_of(n), blk_first;
    for (int j = 0; j < n; ++j) {
        bool join = false;
        if (j > 0) {
            const int64_t c0 = Lp[j] - Lp[j - 1], c1 = Lp[j + 1] - Lp[j];
            const bool chain = c0 > 0 && Li[Lp[j - 1]] == j;
            const int cur_size = j - blk_first.back();
            if (chain && cur_size < kCap && (c0 == c1 + 1 || cur_size < kSmall)) join = true;
        }
        if (!join) blk_first.push_back(j);
        blk_of[j] = (int)blk_first.size() - 1;
    }
    const int nb = (int)blk_first.size();
    blk_first.push_back(n);
    f->n_blocks = nb;

    // ---- levels of the block DAG ----
    std::vector<int> level(nb, 0);
    for (int b = 0; b < nb; ++b)
        for (int j = blk_first[b]; j < blk_first[b + 1]; ++j)
            for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) {
                const int bi = blk_of[Li[p]];
                if (bi != b && level[bi] < level[b] + 1) level[bi] = level[b] + 1;
            }
    int nlev = 0;
    for (int b = 0; b < nb; ++b) nlev = std::max(nlev, level[b] + 1);
    f->n_levels = nlev;

    // ---- split entries: dense diagonal blocks / off-block CSC + CSR ----
    std::vector<int64_t> linv_off(nb + 1, 0);
    int max_block = 0;
    for (int b = 0; b < nb; ++b) {
        const int64_t ns = blk_first[b + 1] - blk_first[b];
        linv_off[b + 1] = linv_off[b] + ns * ns;
        max_block = std::max<int>(max_block, (int)ns);
    }
    std::vector<double> Tdense((size_t)linv_off[nb], 0.0);
    std::vector<int64_t> bc_ptr(n + 1, 0), fr_ptr(n + 1, 0);
    for (int j = 0; j < n; ++j) {
        const int b = blk_of[j];
        const int first = blk_first[b];
        const int64_t ns = blk_first[b + 1] - first;
        double *Tb = Tdense.data() + linv_off[b];
        Tb[(size_t)(j - first) * ns + (j - first)] = 1.0;
        int64_t cnt = 0;
        for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) {
            const int i = Li[p];
            if (blk_of[i] == b)
                Tb[(size_t)(i - first) * ns + (j - first)] = Lx[p];
            else {
                ++cnt;
                fr_ptr[i + 1]++;
            }
        }
        bc_ptr[j + 1] = bc_ptr[j] + cnt;
    }
    for (int i = 0; i < n; ++i) fr_ptr[i + 1] += fr_ptr[i];
    const int64_t noff = bc_ptr[n];
    std::vector<int> bc_row((size_t)noff), fr_col((size_t)noff);
    std::vector<double> bc_val((size_t)noff), fr_val((size_t)noff);
    {
        std::vector<int64_t> pos(fr_ptr.begin(), fr_ptr.end() - 1);
        for (int j = 0; j < n; ++j) {
            const int b = blk_of[j];
            int64_t q = bc_ptr[j];
            for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) {
                const int i = Li[p];
                if (blk_of[i] == b) continue;
                bc_row[q] = i;
                bc_val[q] = Lx[p];
                ++q;
                const int64_t r = pos[i]++;
                fr_col[r] = j;  // columns arrive in ascending order: CSR rows are sorted
                fr_val[r] = Lx[p];
            }
        }
    }

    // ---- level lists: per level and per sweep kernel, LONG rows (one CTA each) and SHORT rows ----
    // kinds: 0 fwd_off (rows with off-block entries), 1 fwd_diag, 2 bwd_off, 3 bwd_diag
    const int64_t kLong = env_int("AAADMM_KLONG", 1024), kTiny = env_int("AAADMM_KTINY", 64);
    auto cls_of = [&](int64_t len) { return len > kLong ? 0 : (len > kTiny ? 1 : 2); };
    std::vector<std::vector<int>> lists((size_t)nlev * 12);
    for (int j = 0; j < n; ++j) {
        const int b = blk_of[j];
        const int l = level[b];
        const int i = j - blk_first[b], ns = blk_first[b + 1] - blk_first[b];
        const int64_t fo = fr_ptr[j + 1] - fr_ptr[j], bo = bc_ptr[j + 1] - bc_ptr[j];
        if (fo > 0) lists[(size_t)l * 12 + 0 + cls_of(fo)].push_back(j);
        lists[(size_t)l * 12 + 3 + cls_of(i + 1)].push_back(j);
        lists[(size_t)l * 12 + 6 + cls_of(bo)].push_back(j);
        lists[(size_t)l * 12 + 9 + cls_of(ns - i)].push_back(j);
    }
    std::vector<int> lev_rows;
    f->list_ptr.assign((size_t)nlev * 12 + 1, 0);
    for (size_t k = 0; k < lists.size(); ++k) {
        lev_rows.insert(lev_rows.end(), lists[k].begin(), lists[k].end());
        f->list_ptr[k + 1] = (int)lev_rows.size();
    }

    std::vector<int> permv(perm, perm + n), iperm(n);
    for (int k = 0; k < n; ++k) iperm[perm[k]] = k;
    std::vector<double> dinv(n);
    for (int k = 0; k < n; ++k) dinv[k] = 1.0 / D[k];

    // ---- upload ----
    int rc = 0;
    rc |= upload(&f->perm, permv);
    rc |= upload(&f->iperm, iperm);
    rc |= upload(&f->blk_of, blk_of);
    rc |= upload(&f->blk_first, blk_first);
    rc |= upload(&f->linv_off, linv_off);
    rc |= upload(&f->dinv, dinv);
    rc |= upload(&f->fr_ptr, fr_ptr);
    rc |= upload(&f->fr_col, fr_col);
    rc |= upload(&f->fr_val, fr_val);
    rc |= upload(&f->bc_ptr, bc_ptr);
    rc |= upload(&f->bc_row, bc_row);
    rc |= upload(&f->bc_val, bc_val);
    rc |= upload(&f->lev_rows, lev_rows);
    rc |= upload(&f->LinvT, Tdense);  // holds T until the inversion below has run
    if (rc) {
        ldlt_dev_destroy(f);
        return -1;
    }
    const size_t dense_bytes = std::max<size_t>(Tdense.size(), 1) * sizeof(double);
    const size_t vec_bytes = std::max<size_t>((size_t)n * nrhs, 1) * sizeof(double);
    if (cudaMalloc((void **)&f->Linv, dense_bytes) != cudaSuccess || cudaMalloc((void **)&f->W, vec_bytes) != cudaSuccess ||
        cudaMalloc((void **)&f->Y, vec_bytes) != cudaSuccess || cudaMalloc((void **)&f->X, vec_bytes) != cudaSuccess) {
        set_last_error("ldlt: cudaMalloc failed");
        ldlt_dev_destroy(f);
        return -1;
    }
    // ---- invert the diagonal blocks on the device ----
    {
        std::vector<int2> tasks;
        for (int b = 0; b < nb; ++b) {
            const int ns = blk_first[b + 1] - blk_first[b];
            for (int c0 = 0; c0 < ns; c0 += 128) tasks.push_back(make_int2(b, c0));
        }
        int2 *d_tasks = nullptr;
        if (upload(&d_tasks, tasks)) {
            ldlt_dev_destroy(f);
            return -1;
        }
        const int nt = (int)tasks.size();
        if (nt > 0) {
            k_invert_unit_lower<<<nt, 128>>>(d_tasks, f->blk_first, f->linv_off, f->LinvT, f->Linv);
            k_transpose_blocks<<<nt, 128>>>(d_tasks, f->blk_first, f->linv_off, f->Linv, f->LinvT);
        }
        cudaError_t e = cudaDeviceSynchronize();
        cudaFree(d_tasks);
        if (e != cudaSuccess) {
            set_last_error(std::string("ldlt: diagonal-block inversion failed: ") + cudaGetErrorString(e));
            ldlt_dev_destroy(f);
            return -1;
        }
    }
    f->stats.n = n;
    f->stats.n_blocks = nb;
    f->stats.n_levels = nlev;
    f->stats.max_block = max_block;
    f->stats.nnz_L = nnz;
    f->stats.nnz_offdiag = noff;
    f->stats.nnz_diag_dense = linv_off[nb];
    // per apply: off-block values+indices once per sweep, the dense triangles once per sweep,
    // the three n x nrhs vectors a few times
    f->stats.bytes_per_solve = 2.0 * (12.0 * (double)noff + 8.0 * 0.5 * (double)linv_off[nb]) +
                               8.0 * (double)n * nrhs * 8.0;
    *out = f;
    return 0;
}

template <int NR>
static int apply_impl(LdltDev *f, double *x_out, cudaStream_t s, const int *skip) {
    const int nlev = f->n_levels;
    auto lists = [&](int l, int kind) {
        RowLists L;
        const int *p = &f->list_ptr[(size_t)l * 12 + 3 * kind];
        L.long_rows = f->lev_rows + p[0];
        L.n_long = p[1] - p[0];
        L.short_rows = f->lev_rows + p[1];
        L.n_short = p[2] - p[1];
        L.tiny_rows = f->lev_rows + p[2];
        L.n_tiny = p[3] - p[2];
        return L;
    };
    auto grid = [](const RowLists &L) { return L.n_long + short_ctas(L) + (L.n_tiny + TINY_PER_CTA - 1) / TINY_PER_CTA; };
    const int T = WARPS_PER_CTA * 32;
    for (int l = 0; l < nlev; ++l) {
        RowLists a = lists(l, 0), b = lists(l, 1);
        if (grid(a) > 0) k_fwd_off<NR><<<grid(a), T, 0, s>>>(a, f->fr_ptr, f->fr_col, f->fr_val, f->Y, f->W, skip);
        if (grid(b) > 0)
            k_fwd_diag<NR><<<grid(b), T, 0, s>>>(b, f->blk_of, f->blk_first, f->linv_off, f->Linv, f->W, f->Y, skip);
    }
    for (int l = nlev - 1; l >= 0; --l) {
        RowLists a = lists(l, 2), b = lists(l, 3);
        if (grid(a) > 0)
            k_bwd_off<NR><<<grid(a), T, 0, s>>>(a, f->bc_ptr, f->bc_row, f->bc_val, f->dinv, f->Y, f->X, f->W, skip);
        if (grid(b) > 0)
            k_bwd_diag<NR><<<grid(b), T, 0, s>>>(b, f->blk_of, f->blk_first, f->linv_off, f->LinvT, f->W, f->X,
                                                 f->perm, x_out, skip);
    }
    AAADMM_CUDA_OK(cudaGetLastError());
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
