// Responsibilities, mixture weights and likelihood on the k x C matrix of per-cell log ratios.
//
//   affinity_kernel  replaces getAffinity,  R/mnems.r:1390-1441
//   mix_stats_kernel replaces getLL (R/mnems_low.r:540-552) fused with the mixture-weight update
//                    mw <- normalise(rowSums(getAffinity(L, mw)))  (R/mnems_low.r:591-595, 653-657; R/mnems.r:1936-1940, 2113-2121)
//
// Reductions are fixed-shape (per-thread strided sums, warp tree, block order, then block partials in
// ascending order) so results are bit-reproducible from run to run.  Element (i, c) of L lives at
// L[i*si + c*sc]: si = Cp, sc = 1 for the sorted internal layout; si = 1, sc = k for R's column-major k x C.
#include "common.cuh"

namespace mn {

struct MixGeom {
    int ncell;              // cells (columns) to visit
    long long si, sc;       // strides of L (component, cell)
    const int32_t *valid;   // optional [ncell]: < 0 marks padding that must not enter any sum
    long long gsi, gsc;     // strides of the gamma output of affinity_kernel (0, 0: same as L)
};

__device__ __forceinline__ double pow_b(double b, double lb, double y)
{   // logtype^y (R_pow): exp2 for the default base 2
    return b == 2.0 ? exp2(y) : pow(b, y);
}

// One cell: y_i = mw_i * b^(L_i - shift); returns column sum; xs[] receives y_i.
// KT > 0: k known at compile time (loops unroll, the small per-cell arrays stay in registers); KT = 0: runtime k.
template <bool kComplete, int KT = 0>
__device__ __forceinline__ double cell_terms(const double *L, long long si, int k_rt, const double *mw, bool shift_on,
                                             double shrink_k, double b, double lb, double *xs, double *xraw)
{
    const int k = KT > 0 ? KT : k_rt;
    double xmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < k; ++i) {
        const double x = L[i * si];
        xraw[i] = x;
        if (x > xmax) xmax = x;
    }
    double cs = 0.0;
#pragma unroll
    for (int i = 0; i < k; ++i) {
        double v = xraw[i];
        if (kComplete && shift_on) v = v - (xmax - shrink_k);      // R/mnems.r:1427
        const double y = pow_b(b, lb, v) * mw[i];                   // :1431-1432
        xs[i] = y;
        cs += y;                                                    // colSums :1433
    }
    return cs;
}

template <bool kComplete>
__global__ void __launch_bounds__(256) affinity_kernel(MixSlots s, MixGeom g, int k, int hard, double b, double lb,
                                                       double shrink_k)
{
    const int slot = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncell) return;
    const long long osi = g.gsi ? g.gsi : g.si, osc = g.gsc ? g.gsc : g.sc;
    double *out = s.gam[slot] + c * osc;
    if (g.valid && g.valid[c] < 0) {
        for (int i = 0; i < k; ++i) out[i * osi] = 0.0;
        return;
    }
    double mw[MNEMCU_MAX_K], xs[MNEMCU_MAX_K], xraw[MNEMCU_MAX_K];
    for (int i = 0; i < k; ++i) mw[i] = s.mw[slot] ? s.mw[slot][i] : 1.0 / k;   // :1392
    const double *L = s.L[slot] + c * g.sc;
    if (hard) {                                                     // affinity == 1, :1403-1417
        double mx = -INFINITY;
        for (int i = 0; i < k; ++i) {
            xs[i] = pow_b(b, lb, L[i * g.si]) * mw[i];
            if (xs[i] > mx) mx = xs[i];
        }
        for (int i = 0; i < k; ++i) {
            double y = xs[i] < mx ? 0.0 : xs[i];
            if (y != 0.0) y = 1.0;
            if (isnan(y)) y = 0.0;
            out[i * osi] = y;
        }
        return;
    }
    if (k == 1) { out[0] = 1.0; return; }                           // :1435-1437
    const bool shift_on = kComplete && s.anypos[slot] && *s.anypos[slot] != 0;   // :1422 any(y > 0)
    const double cs = cell_terms<kComplete>(L, g.si, k, mw, shift_on, shrink_k, b, lb, xs, xraw);
    for (int i = 0; i < k; ++i) {
        double v = xs[i] / cs;
        if (isnan(v)) v = 0.0;                                      // :1439
        out[i * osi] = v;
    }
}

// The EM driver's gamma is cell-major [Cp][gw] with the mixtures of a tick side by side in a row (the B operand the
// M-step statistics kernel stages with one bulk copy per pipeline stage).  One thread per cell would write k doubles
// per mixture at a stride of a whole row, so a CTA takes kAffCells cells x all listed mixtures, collects the rows in shared
// memory and writes them out row by row.  Same arithmetic per cell as affinity_kernel (soft affinity only).
constexpr int kAffCells = 128;
template <bool kComplete, int KT>
__global__ void __launch_bounds__(2 * kAffCells) affinity_rows_kernel(MixSlots s, MixGeom g, int k_rt, double b, double lb,
                                                                      double shrink_k)
{
    extern __shared__ double aff_tile[];                           // [kAffCells][wp]
    constexpr int KA = KT > 0 ? KT : MNEMCU_MAX_K;
    const int k = KT > 0 ? KT : k_rt;
    const int w = s.n * k, wp = w | 1;                             // odd row pitch: conflict-free 8-byte columns
    const int cl = threadIdx.x & (kAffCells - 1);
    const int c0 = blockIdx.x * kAffCells, c = c0 + cl;
    const bool live = c < g.ncell && !(g.valid && g.valid[c] < 0);
    for (int j = threadIdx.x / kAffCells; j < s.n; j += 2) {       // the two halves of the CTA alternate over the mixtures
        double *t = aff_tile + cl * wp + j * k;
        if (!live || k == 1) {                                      // padding: 0; k == 1: R/mnems.r:1435-1437
#pragma unroll
            for (int i = 0; i < k; ++i) t[i] = live ? 1.0 : 0.0;
            continue;
        }
        double mw[KA], xs[KA], xraw[KA];
#pragma unroll
        for (int i = 0; i < k; ++i) mw[i] = s.mw[j] ? s.mw[j][i] : 1.0 / k;
        const bool shift_on = kComplete && s.anypos[j] && *s.anypos[j] != 0;
        const double cs = cell_terms<kComplete, KT>(s.L[j] + c * g.sc, g.si, k, mw, shift_on, shrink_k, b, lb, xs, xraw);
#pragma unroll
        for (int i = 0; i < k; ++i) {
            double v = xs[i] / cs;
            if (isnan(v)) v = 0.0;
            t[i] = v;
        }
    }
    __syncthreads();
    const int nc = min(kAffCells, g.ncell - c0);
    for (int idx = threadIdx.x; idx < nc * w; idx += 2 * kAffCells) {
        const int cc = idx / w, q = idx - cc * w, j = q / k, i = q - j * k;
        s.gam[j][(long long)(c0 + cc) * g.gsc + i] = aff_tile[cc * wp + q];
    }
}

__device__ __forceinline__ double block_sum_256(double v, double *sm)
{   // fixed tree: warp shuffle, then the 8 warp results in order
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    return t;   // valid in thread 0
}

// partial[(slot*nb + blockIdx.x)*(k+1) + i] : i < k row sums of gamma, i == k log-likelihood
// hard != 0: getAffinity(affinity = 1), R/mnems.r:1393-1417 -- per cell the components whose mw_i b^L_i equals the column
// maximum count 1, the others 0 (only without `complete`: the reference's hard + complete branch is an error)
template <bool kComplete, int KT>
__global__ void __launch_bounds__(256) mix_stats_kernel(MixSlots s, MixGeom g, int k_rt, int hard, double b, double lb,
                                                        double shrink_k, double *__restrict__ partial)
{
    __shared__ double sm[8];
    const int slot = blockIdx.y;
    constexpr int KA = KT > 0 ? KT : MNEMCU_MAX_K;
    const int k = KT > 0 ? KT : k_rt;
    double mw[KA], lmw[KA], xs[KA], xraw[KA], rs[KA];
#pragma unroll
    for (int i = 0; i < k; ++i) {
        mw[i] = s.mw[slot] ? s.mw[slot][i] : 1.0 / k;
        lmw[i] = log(mw[i]) / lb;                                   // R/mnems_low.r:545
        rs[i] = 0.0;
    }
    const bool shift_on = kComplete && s.anypos[slot] && *s.anypos[slot] != 0;
    double ll = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < g.ncell; c += gridDim.x * blockDim.x) {
        if (g.valid && g.valid[c] < 0) continue;
        const double *L = s.L[slot] + c * g.sc;
        if (kComplete) {
            // getLL(complete = TRUE): sum_i Z_ic * (L_ic + log_b mw_i), Z = getAffinity(complete = TRUE)
            double cs;
            if (k == 1) { xraw[0] = L[0]; xs[0] = 1.0; cs = 1.0; }
            else cs = cell_terms<true, KT>(L, g.si, k, mw, shift_on, shrink_k, b, lb, xs, xraw);
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < k; ++i) {
                double z = xs[i] / cs;
                if (isnan(z)) z = 0.0;
                rs[i] += z;
                t += z * (xraw[i] + lmw[i]);
            }
            ll += t;
        } else {
            // getLL: log(sum_i mw_i b^L_ic)/log(b), un-stabilised (R/mnems_low.r:547-549)
            const double cs = cell_terms<false, KT>(L, g.si, k, mw, false, 0.0, b, lb, xs, xraw);
            ll += log(cs) / lb;
            if (hard) {
                double mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < k; ++i) if (xs[i] > mx) mx = xs[i];
#pragma unroll
                for (int i = 0; i < k; ++i) {
                    double z = xs[i] < mx ? 0.0 : xs[i];
                    if (z != 0.0) z = 1.0;
                    if (isnan(z)) z = 0.0;
                    rs[i] += z;
                }
            } else {
#pragma unroll
                for (int i = 0; i < k; ++i) {
                    double z = k == 1 ? 1.0 : xs[i] / cs;
                    if (isnan(z)) z = 0.0;
                    rs[i] += z;
                }
            }
        }
    }
    double *out = partial + ((size_t)slot * gridDim.x + blockIdx.x) * (k + 1);
#pragma unroll
    for (int i = 0; i < k; ++i) {
        const double t = block_sum_256(rs[i], sm);
        if (threadIdx.x == 0) out[i] = t;
    }
    const double t = block_sum_256(ll, sm);
    if (threadIdx.x == 0) out[k] = t;
}

// one CTA per slot: block partials in ascending order (staged through shared memory a chunk at a time, so the
// sequential sums read from shared memory instead of chasing nb dependent global loads), then mw <- rs / sum(rs),
// NA guard, ll
constexpr int kFinThreads = 256;
constexpr int kFinStage = 3072;          // doubles staged per chunk
__global__ void __launch_bounds__(kFinThreads) mix_finalize_kernel(MixSlots s, int k, int nb, const double *__restrict__ partial)
{
    __shared__ double sp[kFinStage];
    const int slot = blockIdx.x;
    const int tid = threadIdx.x, lane = tid;
    const int w = k + 1;                                       // k row sums + the log-likelihood per block
    const int rows_per = kFinStage / w;
    const double *p = partial + (size_t)slot * nb * w;
    double v = 0.0;                                            // thread i <= k: running sum of column i
    for (int b0 = 0; b0 < nb; b0 += rows_per) {
        const int nr = min(rows_per, nb - b0);
        for (int i = tid; i < nr * w; i += kFinThreads) sp[i] = p[(size_t)b0 * w + i];
        __syncthreads();
        if (tid < w)
            for (int r = 0; r < nr; ++r) v += sp[r * w + tid];
        __syncthreads();
    }
    if (tid < w) sp[tid] = v;                                  // k may be 32: column k (the ll) then lives in warp 1
    __syncthreads();
    if (tid >= 32) return;
    v = lane < k ? sp[lane] : 0.0;
    const double ll = sp[k];
    double tot = 0.0;
    for (int i = 0; i < k; ++i) tot += __shfl_sync(kFull, v, i);
    double m = v / tot;
    const unsigned bad = __ballot_sync(kFull, lane < k && isnan(m));
    if (bad && s.na_guard[slot]) m = 1.0 / k;
    if (lane < k && s.mw_out[slot]) s.mw_out[slot][lane] = m;
    if (lane == 0 && s.ll_out[slot]) *s.ll_out[slot] = ll;
}

__global__ void anypos_kernel(const double *__restrict__ L, size_t n, int32_t *flag)
{
    bool pos = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        pos |= L[i] > 0.0;
    if (__any_sync(kFull, pos) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

static void consts(double logtype, int k, double *lb, double *shrink_k)
{
    *lb = log(logtype);
    *shrink_k = (log(ldexp(1.0, 1023)) / log(logtype)) / k;          // R/mnems.r:1425-1427
}

int launch_affinity_geom(const MixSlots &s, const MixGeom &g, int k, int hard, int complete, double logtype,
                         cudaStream_t st)
{
    double lb, sk;
    consts(logtype, k, &lb, &sk);
    if (!hard && g.gsi == 1 && g.gsc > 1 && s.n * k <= 32) {           // the EM driver's row layout (k * batch <= 32)
        const int nblk = (g.ncell + kAffCells - 1) / kAffCells;
        const size_t smem = sizeof(double) * kAffCells * ((s.n * k) | 1);
#define MN_AFF(KT)                                                                                                        \
    do {                                                                                                                  \
        if (complete) MN_LAUNCH((affinity_rows_kernel<true, KT>), nblk, 2 * kAffCells, smem, st, s, g, k, logtype, lb, sk); \
        else MN_LAUNCH((affinity_rows_kernel<false, KT>), nblk, 2 * kAffCells, smem, st, s, g, k, logtype, lb, sk);         \
    } while (0)
        switch (k) {
            case 2: MN_AFF(2); break;
            case 3: MN_AFF(3); break;
            case 4: MN_AFF(4); break;
            case 5: MN_AFF(5); break;
            case 6: MN_AFF(6); break;
            default: MN_AFF(0); break;
        }
#undef MN_AFF
        return 0;
    }
    dim3 grid((g.ncell + 255) / 256, s.n);
    if (complete) MN_LAUNCH(affinity_kernel<true>, grid, 256, 0, st, s, g, k, hard, logtype, lb, sk);
    else MN_LAUNCH(affinity_kernel<false>, grid, 256, 0, st, s, g, k, hard, logtype, lb, sk);
    return 0;
}

int launch_mix_stats_geom(const MixSlots &s, const MixGeom &g, int k, int hard, int complete, double logtype,
                          double *scratch, cudaStream_t st)
{
    MN_ARG(!(hard && complete), "affinity = 1 with complete = TRUE is an error in the reference too");
    double lb, sk;
    consts(logtype, k, &lb, &sk);
    dim3 grid(kMixBlocks, s.n);
#define MN_MIX(KT)                                                                                                 \
    do {                                                                                                           \
        if (complete) MN_LAUNCH((mix_stats_kernel<true, KT>), grid, 256, 0, st, s, g, k, 0, logtype, lb, sk, scratch); \
        else MN_LAUNCH((mix_stats_kernel<false, KT>), grid, 256, 0, st, s, g, k, hard, logtype, lb, sk, scratch);    \
    } while (0)
    switch (k) {       // the common mixture sizes get a compile-time k (registers instead of local-memory arrays)
        case 1: MN_MIX(1); break;
        case 2: MN_MIX(2); break;
        case 3: MN_MIX(3); break;
        case 4: MN_MIX(4); break;
        case 5: MN_MIX(5); break;
        case 6: MN_MIX(6); break;
        case 7: MN_MIX(7); break;
        case 8: MN_MIX(8); break;
        default: MN_MIX(0); break;
    }
#undef MN_MIX
    MN_LAUNCH(mix_finalize_kernel, s.n, kFinThreads, 0, st, s, k, kMixBlocks, scratch);
    return 0;
}

int launch_affinity(const mnemcu_ctx *cx, const MixSlots &s, int k, int gw, int hard, int complete, double logtype,
                    cudaStream_t st)
{
    MN_ARG(!(hard && complete), "affinity = 1 with complete = TRUE is an error in the reference too");
    // L is [k][Cp]; gamma goes to the cell-major [Cp][gw] layout the M-step statistics kernel stages (s.gam[j] points at
    // the mixture's first column)
    MixGeom g{cx->Cp, (long long)cx->Cp, 1, cx->orig, 1, (long long)gw};
    return launch_affinity_geom(s, g, k, hard, complete, logtype, st);
}

int launch_mix_stats(const mnemcu_ctx *cx, const MixSlots &s, int k, int hard, int complete, double logtype, double *scratch,
                     cudaStream_t st)
{
    MixGeom g{cx->Cp, (long long)cx->Cp, 1, cx->orig, 0, 0};
    return launch_mix_stats_geom(s, g, k, hard, complete, logtype, scratch, st);
}

int launch_anypos(const mnemcu_ctx *cx, const double *L, int k, int32_t *flag, cudaStream_t st)
{
    MN_LAUNCH(anypos_kernel, 296, 256, 0, st, L, (size_t)k * cx->Cp, flag);
    return 0;
}

}  // namespace mn

using namespace mn;

// ---- standalone ABI on R's k x C column-major layout ------------------------------------------------------
static int mix_api(const double *L, int k, int C, const double *mw, int affinity, int complete, double logtype,
                   double *gam_out, double *ll_out)
{
    MN_ARG(L != nullptr, "L is NULL");
    MN_ARG(k >= 1 && k <= MNEMCU_MAX_K && C >= 1, "k must be in 1..32 and C positive");
    MN_ARG(logtype > 0 && logtype != 1.0, "logtype must be a valid log base");
    MN_ARG(!(affinity == 1 && complete), "affinity = 1 with complete = TRUE is an error in the reference too");
    const size_t n = (size_t)k * C;
    DevBuf<double> dL, dG, dmw, dscratch, dll;
    DevBuf<int32_t> dflag;
    MN_TRY(dL.alloc(n));
    MN_TRY(dflag.alloc(1, true));
    MN_CU(cudaMemcpy(dL.p, L, sizeof(double) * n, cudaMemcpyHostToDevice));
    if (mw) {
        MN_TRY(dmw.alloc(k));
        MN_CU(cudaMemcpy(dmw.p, mw, sizeof(double) * k, cudaMemcpyHostToDevice));
    }
    MixSlots s{};
    s.n = 1;
    s.L[0] = dL.p;
    s.mw[0] = mw ? dmw.p : nullptr;
    s.anypos[0] = dflag.p;
    MixGeom g{C, 1, (long long)k, nullptr, 0, 0};
    if (complete) MN_LAUNCH(anypos_kernel, 296, 256, 0, 0, dL.p, n, dflag.p);
    if (gam_out) {
        MN_TRY(dG.alloc(n));
        s.gam[0] = dG.p;
        MN_TRY(launch_affinity_geom(s, g, k, affinity, complete, logtype, 0));
        MN_CU(cudaMemcpy(gam_out, dG.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    }
    if (ll_out) {
        MN_TRY(dscratch.alloc((size_t)kMixBlocks * (k + 1)));
        MN_TRY(dll.alloc(1));
        s.ll_out[0] = dll.p;
        MN_TRY(launch_mix_stats_geom(s, g, k, 0, complete, logtype, dscratch.p, 0));
        MN_CU(cudaMemcpy(ll_out, dll.p, sizeof(double), cudaMemcpyDeviceToHost));
    }
    MN_CU(cudaDeviceSynchronize());
    return 0;
}

extern "C" int mnemcu_get_affinity(const double *L, int k, int C, const double *mw, int affinity, int complete,
                                   double logtype, double *out)
{
    MN_ARG(out != nullptr, "out is NULL");
    MN_ARG(affinity == 0 || affinity == 1, "affinity must be 0 or 1");
    return mix_api(L, k, C, mw, affinity, complete, logtype, out, nullptr);
}

extern "C" int mnemcu_get_ll(const double *L, int k, int C, const double *mw, int complete, double logtype, double *ll)
{
    MN_ARG(ll != nullptr, "ll is NULL");
    return mix_api(L, k, C, mw, 0, complete, logtype, nullptr, ll);
}
