// Batched greedy structure search: nem(search = "greedy") of the reference, R/mnems.r:111-142, 203-307, 362-379,
// for many (start, component, restart) searches in lock step.
//
// Per greedy iteration three kernels run over ALL active searches:
//   gen_kernel     c(get.ins.fast, get.rev.tc, get.del.tc)(better)  R/mnems_low.r:1025-1084 as bit-matrix algebra
//                  (one CTA per search; a candidate is 32 column masks; the one-step closure of src/mm.cpp:50-65
//                  in column form is  col'_j = col_j | (x[v,j] ? col_u : 0))
//   score_kernel   scoreAdj(D, model, P, oldadj, trans.close = FALSE)  R/mnems.r:205-217, 439-454: thread = E-gene row,
//                  only the columns that differ from the current graph are re-summed (ascending S-gene order, exactly
//                  the order of the Eigen product they replace), row maximum against 0, fixed-shape sum over rows
//   finish_kernel  which.max (first), accept rule R/mnems.r:279-289, graph update
// Graphs are 32-bit row/column masks (S <= 32), so closure, reduction and neighbour generation are shuffles,
// ballots and popcounts.
#include <climits>

#include "common.cuh"

namespace mn {

constexpr int kCT = 32;        // candidates per score CTA
constexpr int kRows = 128;     // E-gene rows per score CTA (= threads)

struct SearchBatch {
    int nmax = 0, E = 0, S = 0, etiles = 0, maxc = 0;
    DevBuf<const double *> Rptr;
    DevBuf<uint32_t> Prow, Pcol, cand_cols, cand_J, adj_rows, closed_rows, closed_cols;
    DevBuf<int32_t> ncand, active, niter, n_active, theta;
    DevBuf<double> partial, oldscore, maxtrace;
    DevBuf<unsigned long long> nscored;
    int32_t *h_n_active = nullptr;   // pinned
};

// ---- candidate generation --------------------------------------------------------------------------------
// mode 0: the single "candidate" closure(better) used for the initial score (R/mnems.r:135-141)
// mode 1: all single-edge neighbours of `better`
__global__ void __launch_bounds__(256) gen_kernel(int S, int maxc, int mode, const uint32_t *__restrict__ Prow,
                                                  const int32_t *__restrict__ active, uint32_t *__restrict__ Pcol,
                                                  uint32_t *__restrict__ cand_cols, uint32_t *__restrict__ cand_J,
                                                  int32_t *__restrict__ ncand, int8_t *__restrict__ cand_kind)
{
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (!active[q]) {
        if (threadIdx.x == 0) ncand[q] = 0;
        return;
    }
    const uint32_t smask = S == 32 ? 0xffffffffu : ((1u << S) - 1u);
    const uint32_t row = lane < S ? (Prow[q * 32 + lane] | (1u << lane)) & smask : 0u;
    const uint32_t col = warp_transpose(row, lane, S);
    uint32_t *cc = cand_cols + (size_t)q * maxc * 32;
    uint32_t *cj = cand_J + (size_t)q * maxc;
    if (warp == 0) Pcol[q * 32 + lane] = col;
    if (mode == 0) {
        if (warp == 0) {
            const uint32_t clo = warp_closure(row, S);
            const uint32_t ccol = warp_transpose(clo, lane, S);
            cc[lane] = ccol;
            if (lane == 0) { cj[0] = smask; ncand[q] = 1; }
        }
        return;
    }
    const uint32_t T = warp_reduce_graph(row, lane, S);          // rows of transitive.reduction(better)
    const uint32_t Tcol = warp_transpose(T, lane, S);
    // lane v owns column v: rows u that yield a candidate, per kind
    uint32_t km[3];
    km[0] = lane < S ? (~col & smask) : 0u;                       // which(Phi == 0)            get.ins.fast :1049
    km[1] = lane < S ? ((Tcol ^ T) & smask) : 0u;                 // Phitr + t(Phitr) == 1      get.rev.tc  :1027
    km[2] = lane < S ? (Tcol & smask) : 0u;                       // which(Phi2 == 1)           get.del.tc  :1073
    int off[3], base = 0;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
        const int cnt = __popc(km[kd]);
        int incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        off[kd] = base + incl - cnt;
        base += __shfl_sync(kFull, incl, 31);
    }
    if (threadIdx.x == 0) ncand[q] = base;
    for (int item = warp; item < 3 * S; item += nwarp) {
        const int kd = item / S, v = item - kd * S;
        uint32_t m = __shfl_sync(kFull, kd == 0 ? km[0] : (kd == 1 ? km[1] : km[2]), v);
        int idx = __shfl_sync(kFull, kd == 0 ? off[0] : (kd == 1 ? off[1] : off[2]), v);
        while (m) {
            const int u = __ffs(m) - 1;
            m &= m - 1;
            uint32_t nc;
            if (kd == 0) {
                // Phinew[u,v] = 1; transClose_Ins(u,v): x[i,j] |= x[i,u] && x[v,j]
                const uint32_t b = col | (lane == v ? (1u << u) : 0u);
                const uint32_t cu = __shfl_sync(kFull, col, u);
                nc = ((b >> v) & 1u) ? (b | cu) : b;
            } else if (kd == 1) {
                // flip (u,v) and (v,u) in better (row u, column v), then mytc(Phinew, uv): Ins if the edge is there, else Del
                const int r = u, c = v;
                const uint32_t f = col ^ (lane == c ? (1u << r) : 0u) ^ (lane == r ? (1u << c) : 0u);
                const uint32_t fc = __shfl_sync(kFull, f, c);
                int uu, vv;
                if ((fc >> r) & 1u) { uu = r; vv = c; } else { uu = c; vv = r; }
                const uint32_t fv = __shfl_sync(kFull, f, vv);
                const uint32_t fu = __shfl_sync(kFull, f, uu);
                if ((fv >> uu) & 1u) {
                    nc = ((f >> vv) & 1u) ? (f | fu) : f;
                } else {
                    // transClose_Del: x[uu,vv] = exists k: x[uu,k] && x[k,vv]   (src/mm.cpp:32-48)
                    const unsigned any = __ballot_sync(kFull, lane < S && ((f >> uu) & 1u) && ((fv >> lane) & 1u));
                    nc = f | ((lane == vv && any) ? (1u << uu) : 0u);
                }
            } else {
                nc = col & ~(lane == v ? (1u << u) : 0u);
            }
            nc &= smask;
            if (lane >= S) nc = 0u;
            const unsigned J = __ballot_sync(kFull, nc != col);
            cc[(size_t)idx * 32 + lane] = nc;
            if (lane == 0) {
                cj[idx] = J;
                if (cand_kind) cand_kind[(size_t)q * maxc + idx] = (int8_t)kd;
            }
            ++idx;
        }
    }
}

// ---- candidate scoring -----------------------------------------------------------------------------------
template <int SP>
__device__ __forceinline__ double col_sum(const double (&r)[SP], uint32_t m)
{   // (R %*% adj)[e, j] with adj[, j] = m: ascending S-gene order, products with 1.0 are exact
    double w = 0.0;
#pragma unroll
    for (int s = 0; s < SP; ++s)
        if ((m >> s) & 1u) { MN_KEEP_BRANCH(); w += r[s]; }
    return w;
}

template <int SP>
__global__ void __launch_bounds__(kRows) score_kernel(int E, int S, int maxc, int etiles,
                                                      const double *const *__restrict__ Rptr,
                                                      const int32_t *__restrict__ ncand,
                                                      const uint32_t *__restrict__ Pcol,
                                                      const uint32_t *__restrict__ cand_cols,
                                                      const uint32_t *__restrict__ cand_J,
                                                      double *__restrict__ partial)
{
    __shared__ uint32_t s_pcol[32];
    __shared__ uint32_t s_ccol[kCT][32];
    __shared__ uint32_t s_J[kCT];
    __shared__ double s_W[SP][kRows];
    __shared__ double s_red[kCT][kRows / 32];
    const int q = blockIdx.z;
    const int nc = ncand[q];
    const int c0 = blockIdx.x * kCT;
    if (c0 >= nc) return;
    const int nloc = min(kCT, nc - c0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 32) s_pcol[tid] = Pcol[q * 32 + tid];
    for (int i = tid; i < nloc * 32; i += kRows) s_ccol[i >> 5][i & 31] = cand_cols[((size_t)q * maxc + c0) * 32 + i];
    if (tid < nloc) s_J[tid] = cand_J[(size_t)q * maxc + c0 + tid];
    const int e = blockIdx.y * kRows + tid;
    const double *R = Rptr[q];
    double r[SP];
#pragma unroll
    for (int s = 0; s < SP; ++s) r[s] = (s < S && e < E) ? R[(size_t)s * E + e] : 0.0;
    __syncthreads();
    // subweights of the current graph (`P` of R/mnems.r:442) and the two largest entries of the row
    double m1 = -INFINITY, m2 = -INFINITY;
    int a1 = 0, a2 = 0;
#pragma unroll
    for (int j = 0; j < SP; ++j) {
        if (j < S) {
            const double w = col_sum<SP>(r, s_pcol[j]);
            s_W[j][tid] = w;
            if (w > m1) { m2 = m1; a2 = a1; m1 = w; a1 = j; }
            else if (w > m2) { m2 = w; a2 = j; }
        }
    }
    for (int ci = 0; ci < nloc; ++ci) {
        uint32_t J = s_J[ci];
        double best;                       // maximum over the columns the candidate shares with the current graph
        if (!((J >> a1) & 1u)) best = m1;
        else if (!((J >> a2) & 1u) && a2 != a1) best = m2;
        else {
            best = -INFINITY;
            for (int j = 0; j < S; ++j)
                if (!((J >> j) & 1u)) best = fmax(best, s_W[j][tid]);
        }
        while (J) {                        // changed columns: score[, changeidx] <- llrScore(...)  R/mnems.r:443-445
            const int j = __ffs(J) - 1;
            J &= J - 1;
            best = fmax(best, col_sum<SP>(r, s_ccol[ci][j]));
        }
        double val = fmax(best, 0.0);      // rowMaxs(cbind(0, score))  R/mnems.r:452-454
        for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(kFull, val, o);
        if (lane == 0) s_red[ci][warp] = val;
    }
    __syncthreads();
    if (tid < nloc) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kRows / 32; ++w) t += s_red[tid][w];
        partial[((size_t)q * etiles + blockIdx.y) * maxc + c0 + tid] = t;
    }
}

// ---- argmax, accept rule, graph update -------------------------------------------------------------------
__global__ void __launch_bounds__(128) finish_kernel(int S, int maxc, int etiles, int mode,
                                                     const int32_t *__restrict__ ncand,
                                                     const double *__restrict__ partial,
                                                     const uint32_t *__restrict__ cand_cols,
                                                     const uint32_t *__restrict__ Pcol, uint32_t *__restrict__ Prow,
                                                     int32_t *__restrict__ active, double *__restrict__ oldscore,
                                                     double *__restrict__ maxtrace, int32_t *__restrict__ niter,
                                                     unsigned long long *__restrict__ nscored,
                                                     int32_t *__restrict__ n_active)
{
    __shared__ double s_best[128];
    __shared__ int s_idx[128];
    const int q = blockIdx.x, tid = threadIdx.x;
    if (!active[q]) return;
    const int nc = ncand[q];
    double best = -INFINITY;
    int bidx = INT_MAX;
    for (int c = tid; c < nc; c += 128) {
        double sc = 0.0;
        for (int et = 0; et < etiles; ++et) sc += partial[((size_t)q * etiles + et) * maxc + c];
        if (isnan(sc)) sc = 0.0;                                   // scores[is.na(scores)] <- 0   R/mnems.r:235
        if (sc > best) { best = sc; bidx = c; }                    // ascending c: first maximum kept
    }
    s_best[tid] = best;
    s_idx[tid] = bidx;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (tid < o) {
            const double ob = s_best[tid + o];
            const int oi = s_idx[tid + o];
            if (ob > s_best[tid] || (ob == s_best[tid] && oi < s_idx[tid])) { s_best[tid] = ob; s_idx[tid] = oi; }
        }
        __syncthreads();
    }
    if (tid >= 32) return;
    best = s_best[0];
    bidx = s_idx[0];
    if (nc == 0 || bidx == INT_MAX) {
        if (tid == 0) active[q] = 0;
        return;
    }
    if (mode == 0) {
        if (tid == 0) { oldscore[q] = best; maxtrace[q] = best; }
        return;
    }
    const uint32_t ccol = cand_cols[((size_t)q * maxc + bidx) * 32 + tid];
    int np = __popc(Pcol[q * 32 + tid]), ncn = __popc(ccol);
    for (int o = 16; o > 0; o >>= 1) {
        np += __shfl_down_sync(kFull, np, o);
        ncn += __shfl_down_sync(kFull, ncn, o);
    }
    np = __shfl_sync(kFull, np, 0);
    ncn = __shfl_sync(kFull, ncn, 0);
    const double old = oldscore[q];
    // R/mnems.r:279-283: better score, or equal score and fewer edges
    const bool accept = best > old || (best == old && np > ncn);
    const uint32_t newrow = warp_transpose(ccol, tid, S);
    if (tid == 0) nscored[q] += (unsigned long long)nc;
    if (accept) {
        Prow[q * 32 + tid] = newrow;
        if (tid == 0) {
            oldscore[q] = best;
            if (best > maxtrace[q]) maxtrace[q] = best;
            niter[q] += 1;
            atomicAdd(n_active, 1);
        }
    } else if (tid == 0) {
        active[q] = 0;
    }
}

__global__ void search_init_kernel(int n, int S, const uint32_t *__restrict__ start_rows, uint32_t *__restrict__ Prow,
                                   int32_t *active, int32_t *niter, unsigned long long *nscored)
{
    const int q = blockIdx.x, lane = threadIdx.x;
    const uint32_t smask = S == 32 ? 0xffffffffu : ((1u << S) - 1u);
    uint32_t r = start_rows ? start_rows[q * 32 + lane] : 0u;
    r = lane < S ? ((r | (1u << lane)) & smask) : 0u;             // diag(better) <- 1, not closed  R/mnems.r:133
    Prow[q * 32 + lane] = r;
    if (lane == 0) { active[q] = 1; niter[q] = 0; nscored[q] = 0ull; }
}

// final graph -> closure (for theta, R/mnems.r:362-369) and transitive reduction (:374)
__global__ void search_final_kernel(int S, const uint32_t *__restrict__ Prow, uint32_t *__restrict__ adj_rows,
                                    uint32_t *__restrict__ closed_rows, uint32_t *__restrict__ closed_cols)
{
    const int q = blockIdx.x, lane = threadIdx.x;
    const uint32_t row = Prow[q * 32 + lane];
    const uint32_t clo = warp_closure(row | (lane < S ? (1u << lane) : 0u), S);
    closed_rows[q * 32 + lane] = clo;
    closed_cols[q * 32 + lane] = warp_transpose(clo, lane, S);
    adj_rows[q * 32 + lane] = warp_reduce_graph(row, lane, S);
}

// theta and (optionally) subweights for one graph per search
template <int SP>
__global__ void __launch_bounds__(kRows) theta_kernel(int E, int S, const double *const *__restrict__ Rptr,
                                                      const uint32_t *__restrict__ cols, int null_last,
                                                      int32_t *__restrict__ theta, double *__restrict__ W_out)
{
    __shared__ uint32_t s_col[32];
    const int q = blockIdx.y, tid = threadIdx.x;
    if (tid < 32) s_col[tid] = cols[q * 32 + tid];
    __syncthreads();
    const int e = blockIdx.x * kRows + tid;
    if (e >= E) return;
    const double *R = Rptr[q];
    double r[SP];
#pragma unroll
    for (int s = 0; s < SP; ++s) r[s] = s < S ? R[(size_t)s * E + e] : 0.0;
    double best = null_last ? -INFINITY : 0.0;
    int arg = 0;
#pragma unroll
    for (int j = 0; j < SP; ++j) {
        if (j < S) {
            const double w = col_sum<SP>(r, s_col[j]);
            if (W_out) W_out[((size_t)q * S + j) * E + e] = w;
            if (w > best) { best = w; arg = j + 1; }               // first strict maximum (src/mm.cpp:74-79)
        }
    }
    if (null_last && 0.0 > best) arg = 0;                          // appended "0" column, R/mnems_low.r:617
    theta[(size_t)q * E + e] = arg;
}

__global__ void closure_kernel(int S, const uint32_t *__restrict__ rows_in, uint32_t *__restrict__ clo_rows,
                               uint32_t *__restrict__ clo_cols)
{
    const int q = blockIdx.x, lane = threadIdx.x;
    const uint32_t smask = S == 32 ? 0xffffffffu : ((1u << S) - 1u);
    uint32_t row = lane < S ? ((rows_in[q * 32 + lane] | (1u << lane)) & smask) : 0u;    // mytc: diag(x) <- 1
    row = warp_closure(row, S);
    if (clo_rows) clo_rows[q * 32 + lane] = row;
    if (clo_cols) clo_cols[q * 32 + lane] = warp_transpose(row, lane, S);
}

__global__ void rows_from_u8_kernel(const uint8_t *__restrict__ g, int S, uint32_t *__restrict__ rows)
{
    rows[blockIdx.x * 32 + threadIdx.x] = load_row_mask(g + (size_t)blockIdx.x * S * S, threadIdx.x, S);
}
__global__ void rows_to_u8_kernel(const uint32_t *__restrict__ rows, int S, uint8_t *__restrict__ g)
{
    store_row_mask(g + (size_t)blockIdx.x * S * S, rows[blockIdx.x * 32 + threadIdx.x], threadIdx.x, S);
}

int launch_closure(const uint32_t *rows_in, int n, int S, uint32_t *clo_rows, uint32_t *clo_cols, cudaStream_t st)
{
    if (n <= 0) return 0;
    MN_LAUNCH(closure_kernel, n, 32, 0, st, S, rows_in, clo_rows, clo_cols);
    return 0;
}
int launch_rows_from_u8(const uint8_t *g, int n, int S, uint32_t *rows, cudaStream_t st)
{
    if (n <= 0) return 0;
    MN_LAUNCH(rows_from_u8_kernel, n, 32, 0, st, g, S, rows);
    return 0;
}
int launch_rows_to_u8(const uint32_t *rows, int n, int S, uint8_t *g, cudaStream_t st)
{
    if (n <= 0) return 0;
    MN_LAUNCH(rows_to_u8_kernel, n, 32, 0, st, rows, S, g);
    return 0;
}

int launch_theta(const double *const *Rptr_dev, int n, int E, int S, const uint32_t *cols, int null_last, int32_t *theta,
                 double *W_out, cudaStream_t st)
{
    if (n <= 0) return 0;
    dim3 grid((E + kRows - 1) / kRows, n);
    if (S <= 8) MN_LAUNCH(theta_kernel<8>, grid, kRows, 0, st, E, S, Rptr_dev, cols, null_last, theta, W_out);
    else if (S <= 16) MN_LAUNCH(theta_kernel<16>, grid, kRows, 0, st, E, S, Rptr_dev, cols, null_last, theta, W_out);
    else if (S <= 24) MN_LAUNCH(theta_kernel<24>, grid, kRows, 0, st, E, S, Rptr_dev, cols, null_last, theta, W_out);
    else MN_LAUNCH(theta_kernel<32>, grid, kRows, 0, st, E, S, Rptr_dev, cols, null_last, theta, W_out);
    return 0;
}

static int launch_score(SearchBatch *sb, int n, cudaStream_t st)
{
    dim3 grid((sb->maxc + kCT - 1) / kCT, sb->etiles, n);
    const int S = sb->S;
#define MN_SCORE(SPV)                                                                                              \
    MN_LAUNCH(score_kernel<SPV>, grid, kRows, 0, st, sb->E, S, sb->maxc, sb->etiles, sb->Rptr.p, sb->ncand.p,      \
              sb->Pcol.p, sb->cand_cols.p, sb->cand_J.p, sb->partial.p)
    if (S <= 8) MN_SCORE(8);
    else if (S <= 16) MN_SCORE(16);
    else if (S <= 24) MN_SCORE(24);
    else MN_SCORE(32);
#undef MN_SCORE
    return 0;
}

int search_create(int nmax, int E, int S, SearchBatch **out)
{
    MN_ARG(nmax >= 1 && E >= 1 && S >= 1 && S <= MNEMCU_MAX_S, "bad search workspace size");
    SearchBatch *sb = new SearchBatch();
    sb->nmax = nmax; sb->E = E; sb->S = S;
    sb->etiles = (E + kRows - 1) / kRows;
    sb->maxc = std::max(1, 3 * S * (S - 1));
    auto body = [&]() -> int {
        MN_TRY(sb->Rptr.alloc(nmax));
        MN_TRY(sb->Prow.alloc((size_t)nmax * 32, true)); MN_TRY(sb->Pcol.alloc((size_t)nmax * 32, true));
        MN_TRY(sb->cand_cols.alloc((size_t)nmax * sb->maxc * 32)); MN_TRY(sb->cand_J.alloc((size_t)nmax * sb->maxc));
        MN_TRY(sb->adj_rows.alloc((size_t)nmax * 32)); MN_TRY(sb->closed_rows.alloc((size_t)nmax * 32));
        MN_TRY(sb->closed_cols.alloc((size_t)nmax * 32));
        MN_TRY(sb->ncand.alloc(nmax, true)); MN_TRY(sb->active.alloc(nmax, true)); MN_TRY(sb->niter.alloc(nmax, true));
        MN_TRY(sb->n_active.alloc(1, true)); MN_TRY(sb->theta.alloc((size_t)nmax * E));
        MN_TRY(sb->partial.alloc((size_t)nmax * sb->etiles * sb->maxc));
        MN_TRY(sb->oldscore.alloc(nmax)); MN_TRY(sb->maxtrace.alloc(nmax)); MN_TRY(sb->nscored.alloc(nmax, true));
        MN_CU(cudaMallocHost((void **)&sb->h_n_active, sizeof(int32_t)));
        return 0;
    };
    int rc = body();
    if (rc) { search_destroy(sb); return rc; }
    *out = sb;
    return 0;
}

void search_destroy(SearchBatch *sb)
{
    if (!sb) return;
    if (sb->h_n_active) cudaFreeHost(sb->h_n_active);
    delete sb;
}

int search_run(SearchBatch *sb, int n, const double *const *R_ptrs_host, const uint32_t *start_rows_dev, cudaStream_t st)
{
    MN_ARG(n >= 1 && n <= sb->nmax, "too many searches for the workspace");
    const int S = sb->S;
    MN_CU(cudaMemcpyAsync(sb->Rptr.p, R_ptrs_host, sizeof(double *) * n, cudaMemcpyHostToDevice, st));
    MN_LAUNCH(search_init_kernel, n, 32, 0, st, n, S, start_rows_dev, sb->Prow.p, sb->active.p, sb->niter.p, sb->nscored.p);
    // initial score of closure(better)
    MN_LAUNCH(gen_kernel, n, 256, 0, st, S, sb->maxc, 0, sb->Prow.p, sb->active.p, sb->Pcol.p, sb->cand_cols.p,
              sb->cand_J.p, sb->ncand.p, (int8_t *)nullptr);
    MN_TRY(launch_score(sb, n, st));
    MN_LAUNCH(finish_kernel, n, 128, 0, st, S, sb->maxc, sb->etiles, 0, sb->ncand.p, sb->partial.p, sb->cand_cols.p,
              sb->Pcol.p, sb->Prow.p, sb->active.p, sb->oldscore.p, sb->maxtrace.p, sb->niter.p, sb->nscored.p,
              sb->n_active.p);
    const int hard_cap = 4 * S * S + 16;   // a strict hill-climb cannot take more moves than this in practice
    for (int it = 0; it < hard_cap; ++it) {
        MN_CU(cudaMemsetAsync(sb->n_active.p, 0, sizeof(int32_t), st));
        MN_LAUNCH(gen_kernel, n, 256, 0, st, S, sb->maxc, 1, sb->Prow.p, sb->active.p, sb->Pcol.p, sb->cand_cols.p,
                  sb->cand_J.p, sb->ncand.p, (int8_t *)nullptr);
        MN_TRY(launch_score(sb, n, st));
        MN_LAUNCH(finish_kernel, n, 128, 0, st, S, sb->maxc, sb->etiles, 1, sb->ncand.p, sb->partial.p, sb->cand_cols.p,
                  sb->Pcol.p, sb->Prow.p, sb->active.p, sb->oldscore.p, sb->maxtrace.p, sb->niter.p, sb->nscored.p,
                  sb->n_active.p);
        MN_CU(cudaMemcpyAsync(sb->h_n_active, sb->n_active.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        MN_CU(cudaStreamSynchronize(st));
        if (*sb->h_n_active == 0) break;
    }
    MN_LAUNCH(search_final_kernel, n, 32, 0, st, S, sb->Prow.p, sb->adj_rows.p, sb->closed_rows.p, sb->closed_cols.p);
    MN_TRY(launch_theta(sb->Rptr.p, n, sb->E, S, sb->closed_cols.p, 0, sb->theta.p, nullptr, st));
    return 0;
}

const uint32_t *search_adj_rows(const SearchBatch *sb) { return sb->adj_rows.p; }
const uint32_t *search_closed_rows(const SearchBatch *sb) { return sb->closed_rows.p; }
const int32_t *search_theta(const SearchBatch *sb) { return sb->theta.p; }
const double *search_score(const SearchBatch *sb) { return sb->oldscore.p; }
const double *search_maxtrace(const SearchBatch *sb) { return sb->maxtrace.p; }
const int32_t *search_niter(const SearchBatch *sb) { return sb->niter.p; }
const unsigned long long *search_nscored(const SearchBatch *sb) { return sb->nscored.p; }

}  // namespace mn

using namespace mn;

// ---- ABI ---------------------------------------------------------------------------------------------------
extern "C" int mnemcu_nem_greedy(const double *R, int E, int S, const uint8_t *start, int n, uint8_t *adj_out,
                                 int32_t *theta_out, double *score, double *max_trace, int32_t *n_iter,
                                 int64_t *n_scored)
{
    MN_ARG(R && adj_out && theta_out, "NULL pointer");
    MN_ARG(E >= 1 && S >= 1 && S <= MNEMCU_MAX_S && n >= 1, "bad size");
    SearchBatch *sb = nullptr;
    MN_TRY(search_create(n, E, S, &sb));
    auto body = [&]() -> int {
        DevBuf<double> dR;
        DevBuf<uint8_t> dg;
        DevBuf<uint32_t> drows;
        MN_TRY(dR.alloc((size_t)n * E * S));
        MN_CU(cudaMemcpy(dR.p, R, sizeof(double) * (size_t)n * E * S, cudaMemcpyHostToDevice));
        MN_TRY(dg.alloc((size_t)n * S * S));
        if (start) {
            MN_TRY(drows.alloc((size_t)n * 32));
            MN_CU(cudaMemcpy(dg.p, start, (size_t)n * S * S, cudaMemcpyHostToDevice));
            MN_TRY(launch_rows_from_u8(dg.p, n, S, drows.p, 0));
        }
        std::vector<const double *> ptrs(n);
        for (int q = 0; q < n; ++q) ptrs[q] = dR.p + (size_t)q * E * S;
        MN_TRY(search_run(sb, n, ptrs.data(), start ? drows.p : nullptr, 0));
        MN_TRY(launch_rows_to_u8(search_adj_rows(sb), n, S, dg.p, 0));
        MN_CU(cudaMemcpy(adj_out, dg.p, (size_t)n * S * S, cudaMemcpyDeviceToHost));
        MN_CU(cudaMemcpy(theta_out, search_theta(sb), sizeof(int32_t) * (size_t)n * E, cudaMemcpyDeviceToHost));
        if (score) MN_CU(cudaMemcpy(score, search_score(sb), sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (max_trace) MN_CU(cudaMemcpy(max_trace, search_maxtrace(sb), sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (n_iter) MN_CU(cudaMemcpy(n_iter, search_niter(sb), sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
        if (n_scored) {
            static_assert(sizeof(unsigned long long) == sizeof(int64_t), "");
            MN_CU(cudaMemcpy(n_scored, search_nscored(sb), sizeof(int64_t) * n, cudaMemcpyDeviceToHost));
        }
        return 0;
    };
    int rc = body();
    search_destroy(sb);
    return rc;
}

// scoreAdj for n candidate graphs against one R: uses the same scoring kernel with every column marked changed
__global__ void score_api_prep_kernel(int S, int maxc, int n, const uint32_t *__restrict__ cols_in,
                                      uint32_t *__restrict__ cand_cols, uint32_t *__restrict__ cand_J,
                                      uint32_t *__restrict__ Pcol, int32_t *__restrict__ ncand)
{
    const int c = blockIdx.x, lane = threadIdx.x;
    const uint32_t smask = S == 32 ? 0xffffffffu : ((1u << S) - 1u);
    cand_cols[(size_t)c * 32 + lane] = cols_in[c * 32 + lane];
    if (lane == 0) cand_J[c] = smask;
    if (c == 0) { Pcol[lane] = 0u; if (lane == 0) ncand[0] = n; }
}

__global__ void score_api_sum_kernel(int n, int maxc, int etiles, const double *__restrict__ partial, double *score)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double sc = 0.0;
    for (int et = 0; et < etiles; ++et) sc += partial[(size_t)et * maxc + c];
    score[c] = sc;
}

__global__ void cols_from_rows_kernel(int S, int trans_close, const uint32_t *__restrict__ rows, uint32_t *__restrict__ cols)
{
    const int q = blockIdx.x, lane = threadIdx.x;
    const uint32_t smask = S == 32 ? 0xffffffffu : ((1u << S) - 1u);
    uint32_t row = lane < S ? (rows[q * 32 + lane] & smask) : 0u;
    if (trans_close) row = warp_closure(lane < S ? (row | (1u << lane)) : 0u, S);     // mytc, R/mnems.r:428-430
    cols[q * 32 + lane] = warp_transpose(row, lane, S);
}

extern "C" int mnemcu_score_adj(const double *R, int E, int S, const uint8_t *adj, int n, int trans_close,
                                double *score, int32_t *theta, double *W)
{
    MN_ARG(R && adj && score, "NULL pointer");
    MN_ARG(E >= 1 && S >= 1 && S <= MNEMCU_MAX_S && n >= 1, "bad size");
    SearchBatch sb;
    sb.nmax = 1; sb.E = E; sb.S = S; sb.etiles = (E + kRows - 1) / kRows; sb.maxc = n;
    DevBuf<double> dR, dscore, dW;
    DevBuf<uint8_t> dg;
    DevBuf<uint32_t> drows, dcols;
    DevBuf<int32_t> dtheta;
    DevBuf<const double *> dptr;
    MN_TRY(dR.alloc((size_t)E * S));
    MN_CU(cudaMemcpy(dR.p, R, sizeof(double) * (size_t)E * S, cudaMemcpyHostToDevice));
    MN_TRY(dg.alloc((size_t)n * S * S));
    MN_CU(cudaMemcpy(dg.p, adj, (size_t)n * S * S, cudaMemcpyHostToDevice));
    MN_TRY(drows.alloc((size_t)n * 32)); MN_TRY(dcols.alloc((size_t)n * 32));
    MN_TRY(launch_rows_from_u8(dg.p, n, S, drows.p, 0));
    MN_LAUNCH(cols_from_rows_kernel, n, 32, 0, 0, S, trans_close, drows.p, dcols.p);
    MN_TRY(sb.Rptr.alloc(1)); MN_TRY(sb.Pcol.alloc(32)); MN_TRY(sb.cand_cols.alloc((size_t)n * 32));
    MN_TRY(sb.cand_J.alloc(n)); MN_TRY(sb.ncand.alloc(1)); MN_TRY(sb.partial.alloc((size_t)sb.etiles * n));
    const double *rp = dR.p;
    MN_CU(cudaMemcpy(sb.Rptr.p, &rp, sizeof(double *), cudaMemcpyHostToDevice));
    MN_LAUNCH(score_api_prep_kernel, n, 32, 0, 0, S, n, n, dcols.p, sb.cand_cols.p, sb.cand_J.p, sb.Pcol.p, sb.ncand.p);
    MN_TRY(launch_score(&sb, 1, 0));
    MN_TRY(dscore.alloc(n));
    MN_LAUNCH(score_api_sum_kernel, (n + 127) / 128, 128, 0, 0, n, n, sb.etiles, sb.partial.p, dscore.p);
    MN_CU(cudaMemcpy(score, dscore.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (theta || W) {
        std::vector<const double *> ptrs(n, dR.p);
        MN_TRY(dptr.alloc(n));
        MN_CU(cudaMemcpy(dptr.p, ptrs.data(), sizeof(double *) * n, cudaMemcpyHostToDevice));
        MN_TRY(dtheta.alloc((size_t)n * E));
        if (W) MN_TRY(dW.alloc((size_t)n * E * S));
        MN_TRY(launch_theta(dptr.p, n, E, S, dcols.p, 0, dtheta.p, W ? dW.p : nullptr, 0));
        if (theta) MN_CU(cudaMemcpy(theta, dtheta.p, sizeof(int32_t) * (size_t)n * E, cudaMemcpyDeviceToHost));
        if (W) MN_CU(cudaMemcpy(W, dW.p, sizeof(double) * (size_t)n * E * S, cudaMemcpyDeviceToHost));
    }
    MN_CU(cudaDeviceSynchronize());
    return 0;
}

// c(get.ins.fast, get.rev.tc, get.del.tc)(P) for one graph, as uint8 matrices (test / inspection entry point)
__global__ void cand_to_u8_kernel(int S, const uint32_t *__restrict__ cand_cols, uint8_t *__restrict__ models)
{
    const int c = blockIdx.x, lane = threadIdx.x;
    const uint32_t row = warp_transpose(cand_cols[(size_t)c * 32 + lane], lane, S);
    store_row_mask(models + (size_t)c * S * S, row, lane, S);
}

extern "C" int mnemcu_neighbours(const uint8_t *P, int S, uint8_t *models, int8_t *kind, int32_t *n_out)
{
    MN_ARG(P && models && n_out, "NULL pointer");
    MN_ARG(S >= 1 && S <= MNEMCU_MAX_S, "S must be in 1..32");
    const int maxc = std::max(1, 3 * S * (S - 1));
    DevBuf<uint8_t> dg, dmodels;
    DevBuf<uint32_t> drow, dPcol, dcc, dJ;
    DevBuf<int32_t> dn, dact;
    DevBuf<int8_t> dkind;
    MN_TRY(dg.alloc((size_t)S * S)); MN_TRY(drow.alloc(32)); MN_TRY(dPcol.alloc(32));
    MN_TRY(dcc.alloc((size_t)maxc * 32)); MN_TRY(dJ.alloc(maxc)); MN_TRY(dn.alloc(1, true)); MN_TRY(dact.alloc(1));
    MN_TRY(dkind.alloc(maxc)); MN_TRY(dmodels.alloc((size_t)maxc * S * S));
    const int32_t one = 1;
    MN_CU(cudaMemcpy(dg.p, P, (size_t)S * S, cudaMemcpyHostToDevice));
    MN_CU(cudaMemcpy(dact.p, &one, sizeof(one), cudaMemcpyHostToDevice));
    MN_TRY(launch_rows_from_u8(dg.p, 1, S, drow.p, 0));
    MN_LAUNCH(gen_kernel, 1, 256, 0, 0, S, maxc, 1, drow.p, dact.p, dPcol.p, dcc.p, dJ.p, dn.p, dkind.p);
    int32_t n = 0;
    MN_CU(cudaMemcpy(&n, dn.p, sizeof(n), cudaMemcpyDeviceToHost));
    if (n > 0) {
        MN_LAUNCH(cand_to_u8_kernel, n, 32, 0, 0, S, dcc.p, dmodels.p);
        MN_CU(cudaMemcpy(models, dmodels.p, (size_t)n * S * S, cudaMemcpyDeviceToHost));
        if (kind) MN_CU(cudaMemcpy(kind, dkind.p, (size_t)n, cudaMemcpyDeviceToHost));
    }
    *n_out = n;
    return 0;
}
