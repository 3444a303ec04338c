// The two HBM-streaming kernels of the EM loop and their small helpers.
//
//   estep_kernel   replaces the body of getProbs, R/mnems_low.r:606-638:  L[i,c] = colSums(data * t(adj2))
//   mstats_kernel  replaces doMean on the Rho path, R/mnems_low.r:893-906 reached from nem() R/mnems.r:101-107
//
// Both read the expression matrix exactly once per launch and serve up to 32 (start, component) pairs, so
// the algorithmic traffic is 8*E*C bytes in plus the small outputs.  FP64 throughout.
#include <algorithm>

#include "common.cuh"

namespace mn {

// ---------------------------------------------------------------------------------------------------------
// E-step.  Layout Dblk[blk][e][64]: for one E-gene the 64 cells of a block are contiguous, so a warp reads
// 512 contiguous bytes per E-gene (one double2 per lane) and every lane owns two cells.  All 64 cells of a
// block were perturbed in the same S-gene (segments are padded), so the effect mask F[e] = phi[s, theta(e)]
// is warp-uniform: 32 consecutive mask words are fetched with one coalesced load and broadcast by shuffle.
// Bit m of a mask word belongs to (start, component) pair m; pair m writes to its own output row out.L[m]
// ([Cp] doubles) because the EM slots rotate their k x C buffers independently.  No cross-lane reduction is needed: each lane
// accumulates its own cells in ascending E-gene order, exactly the order of colSums() in the reference.
//
// grid = nblk, block = 32 * groups * esplit threads: warp (g, w) serves pairs [g*MT, (g+1)*MT) over the E-genes
// [w*eper, (w+1)*eper).  The `groups` warps that share a cell block sit in the same CTA, so only the first of them
// pulls a line from HBM; the others hit L1/L2.
//
// The masked add is written as a real branch (MN_KEEP_BRANCH) so that ptxas emits predicated DADDs fed by R2P
// (1 issue slot per pair and cell); left alone it if-converts to DADD + 2 FSEL, three issue slots per add.
// ---------------------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(MT > 11 ? 256 : 384) estep_kernel(const double *__restrict__ Dblk, int E,
                                                    const uint32_t *__restrict__ maskT,
                                                    const int32_t *__restrict__ blk_seg, int M, int eper,
                                                    int groups, PairPtrs out)
{
    extern __shared__ double2 s_part[];   // [groups][esplit-1][MT][32] partial sums of warps 1..
    const int blk = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int grp = (threadIdx.x >> 5) % groups;
    const int warp = (threadIdx.x >> 5) / groups;          // E-gene range index
    const int nwarp = (blockDim.x >> 5) / groups;
    const int m0 = grp * MT;
    const int ebeg = warp * eper;
    const int eend = min(E, ebeg + eper);
    const double2 *base = reinterpret_cast<const double2 *>(Dblk + (size_t)blk * E * kBlkCells) + lane;
    const uint32_t *mk = maskT + (size_t)blk_seg[blk] * E;

    // scalar accumulators (not double2): ptxas then adds in place under the predicate instead of via a temporary
    double ax[MT], ay[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) { ax[m] = 0.0; ay[m] = 0.0; }

    // E-genes are taken 32 at a time and EVERY warp runs the same number of groups (eper / 32): with a uniform trip
    // count the compiler knows the __shfl_sync below is convergent and emits one copy of the loop.  A range may run
    // past eend / E; those reads stay inside the allocation (256 E-genes of slack) and their mask words are zero.
    const int nit = eper >> 5;
    for (int it = 0; it < nit; ++it) {
        const int e0 = ebeg + (it << 5);
        const uint32_t wv = (e0 + lane < eend) ? (__ldg(mk + e0 + lane) >> m0) : 0u;
#pragma unroll
        for (int jb = 0; jb < 32; jb += 8) {
            double2 d[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = ld_stream2(base + (size_t)(e0 + jb + j) * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t w = __shfl_sync(kFull, wv, jb + j);
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    if ((w >> m) & 1u) { MN_KEEP_BRANCH(); ax[m] += d[j].x; ay[m] += d[j].y; }
            }
        }
    }
    if (nwarp > 1) {   // combine the E-gene ranges in ascending order (fixed shape => deterministic)
        if (warp > 0) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
                s_part[((grp * (nwarp - 1) + warp - 1) * MT + m) * 32 + lane] = make_double2(ax[m], ay[m]);
        }
        __syncthreads();
        if (warp == 0) {
            for (int w = 1; w < nwarp; ++w)
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const double2 p = s_part[((grp * (nwarp - 1) + w - 1) * MT + m) * 32 + lane];
                    ax[m] += p.x;
                    ay[m] += p.y;
                }
        }
    }
    if (warp == 0) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (m0 + m < M) {   // warp-uniform
                reinterpret_cast<double2 *>(out.L[m0 + m] + (size_t)blk * kBlkCells)[lane] = make_double2(ax[m], ay[m]);
                const bool pos = ax[m] > 0.0 || ay[m] > 0.0;
                const unsigned any = __ballot_sync(kFull, pos);
                if (any && lane == 0) atomicOr(out.flag[m0 + m], 1);   // any(L > 0), R/mnems.r:1422
            }
        }
    }
}

template <int MT>
static int launch_estep_t(const mnemcu_ctx *cx, const uint32_t *maskT, int M, const PairPtrs &out, int esplit,
                          cudaStream_t st)
{
    const int groups = (M + MT - 1) / MT;
    const int eper = ((cx->E + esplit - 1) / esplit + 31) / 32 * 32;
    size_t smem = esplit > 1 ? (size_t)groups * (esplit - 1) * MT * 32 * sizeof(double2) : 0;
    // the attribute belongs to the current device's context (a process may hold contexts on several devices): set it
    // whenever it is needed -- the call is cheap next to a pass over D
    if (smem > 48 * 1024)                    // small C: up to 7 partial-sum slabs of MT x 32 double2 (70 KB at MT = 20)
        MN_CU(cudaFuncSetAttribute(estep_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 20 * 32 * 16));
    MN_LAUNCH(estep_kernel<MT>, cx->nblk, 32 * groups * esplit, smem, st, cx->Dblk, cx->E, maskT, cx->blk_seg, M, eper,
              groups, out);
    return 0;
}

int launch_estep(const mnemcu_ctx *cx, const uint32_t *maskT, int M, const PairPtrs &out, cudaStream_t st)
{
    MN_ARG(M >= 1 && M <= 32, "E-step serves 1..32 (start, component) pairs per launch");
    // Pairs per warp: up to 20 (40 FP64 accumulators per lane, 125 registers); above that several warps of the same CTA
    // split the pairs and all but the first re-read the block from L1/L2, not from HBM.
    // Measured at 30 pairs, cfg5 (16 GB per pass): one warp with 60 accumulators 4.58 ms; two warps of 15 pairs 4.25 ms;
    // THREE warps of 10 pairs 3.94 ms (85 registers, 24 warps per SM; 4 x 8: 4.45 ms).  At that point the kernel is at
    // 82 % of the FP64 add rate (30 x E x C predicated DADDs per pass), which binds before HBM does.
    // 25 pairs: two warps of 13 pairs 3.76 ms, three warps of 9 pairs 4.72 ms (the 9-pair instantiation schedules badly).
    int groups = M <= 20 ? 1 : (M <= 26 ? 2 : 3);
    int MT = (M + groups - 1) / groups;
    if (groups == 3) MT = std::max(MT, 10);
    if (const char *e = getenv("MNEMCU_ESTEP_MT")) {                                    // experiments
        MT = std::max(1, std::min(20, atoi(e)));
        groups = (M + MT - 1) / MT;
        MT = (M + groups - 1) / groups;
    }
    MN_ARG(groups <= 3, "E-step: at most three warps share a cell block");       // 3 groups x 4 E-gene ranges = 384 threads
    // Few blocks (small C): split the E-gene range over up to 4 warps so that every SM still has loads in flight.  The split
    // changes the ORDER of a cell's sum (ranges are summed separately and then added in ascending order), so it must depend
    // on the shape of the data only -- never on M -- or a start's fit would depend on how many starts share its passes.
    // (Found with full-precision synthetic data: with inputs that are multiples of 2^-16 every order gives the same sum.)
    // The heuristics below therefore assume the nominal 20 pairs in one group whatever M is.
    int esplit = 1;
    const long warps = (long)cx->nblk;
    while (esplit < 4 && warps * esplit < 16L * cx->sm_count && cx->E / (esplit * 2) >= 64) esplit *= 2;
    // Register-limited residency (about 45 + 4 * 20 registers per thread): a grid of between one and ~2.5 waves leaves a
    // long tail, so split once more (cfg3 shape, 20 pairs: 1.32 waves at 2 warps per block, 0.189 -> 0.164 ms at 4).
    {
        const long resident = (long)cx->sm_count * std::min(64, 65536 / ((45 + 4 * 20) * 32));
        const double waves = (double)(warps * esplit) / (double)resident;
        if (waves > 1.0 && waves < 2.5 && esplit < 4 && cx->E / (esplit * 2) >= 64) esplit *= 2;
    }
    switch (MT) {
#define MN_CASE(v) case v: return launch_estep_t<v>(cx, maskT, M, out, esplit, st);
        MN_CASE(1) MN_CASE(2) MN_CASE(3) MN_CASE(4) MN_CASE(5) MN_CASE(6) MN_CASE(7) MN_CASE(8) MN_CASE(9) MN_CASE(10)
        MN_CASE(11) MN_CASE(12) MN_CASE(13) MN_CASE(14) MN_CASE(15) MN_CASE(16) MN_CASE(17) MN_CASE(18) MN_CASE(19)
#undef MN_CASE
        default: return launch_estep_t<20>(cx, maskT, M, out, esplit, st);
    }
}

// maskT[g][e]: bit m = adj1_m[cells of group g, theta_m(e)]  (adj1[, subtopo] of R/mnems_low.r:621-624, theta 0 -> null
// column), where adj1 = min(1, t(Rho) %*% closure(phi_m)) (:613-616): the OR of the closure rows of the group's S-genes.
// For single knock-downs the group is one S-gene and this is phi_m closure [s, theta_m(e)].
__global__ void build_masks_kernel(const uint32_t *__restrict__ clo, const int32_t *__restrict__ theta, int M, int E,
                                   int S, const uint32_t *__restrict__ gmask, uint32_t *__restrict__ maskT)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (e >= E) return;
    const uint32_t gm = gmask[g];
    uint32_t w = 0;
    for (int m = 0; m < M; ++m) {
        const int t = theta[(size_t)m * E + e];
        if (t > 0 && t <= S) {
            uint32_t hit = 0;
            for (uint32_t r = gm; r; r &= r - 1) hit |= clo[m * 32 + (__ffs(r) - 1)];
            if ((hit >> (t - 1)) & 1u) w |= 1u << m;
        }
    }
    maskT[(size_t)g * E + e] = w;
}

int launch_build_masks(const mnemcu_ctx *cx, const uint32_t *clo, const int32_t *theta, int M, uint32_t *maskT,
                       cudaStream_t st)
{
    dim3 grid((cx->E + 127) / 128, cx->G);
    MN_LAUNCH(build_masks_kernel, grid, 128, 0, st, clo, theta, M, cx->E, cx->S, cx->gmask, maskT);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// M-step sufficient statistics.  Layout Dcol[cp][Epad]: a cell's E-genes are contiguous, so a CTA streams its
// slice of D with coalesced 16-byte copies; the cells of one chunk of one segment are contracted against the
// responsibilities gamma[m][cp] on the FP64 tensor cores (see mstats_kernel).  Chunk partials are combined in
// ascending chunk order by mstats_reduce_kernel: no atomics, bit-reproducible.
// ---------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}

constexpr int kStCells = 8;    // cells per pipeline stage (two k-steps of the FP64 MMA)
constexpr int kStages = 3;     // stages in flight: 2 ahead of the one being consumed
constexpr int kMsWarpE = 64;   // E-genes per warp: 8 n-tiles of 8
constexpr int kMsThreads = 128;

// ---- TMA (1-D bulk copy) + mbarrier helpers: one thread arms the stage's barrier with the byte count and issues the
// copies; the copy engine writes shared memory and completes the transaction count; everybody waits on the phase bit.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// D += A * B on the FP64 tensor cores, m8n8k4: A 8x4 row-major (a = A[lane/4][lane%4]), B 4x8 column-major
// (b = B[lane%4][lane/4]), C/D 8x8 (c0, c1 = C[lane/4][2*(lane%4) + {0, 1}]).
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}


// grid = (groups of pairs, E-gene tiles, chunks), block = tpb threads (a multiple of 32, <= 128); warp w owns the 64
// E-genes [tile0 + 64 w, +64).  The contraction over the cells of the chunk is a GEMM
//     Rt[pair][e] += gamma[pair][cell] * D[cell][e]          (M = pairs in 8-row tiles, N = E-genes, K = cells)
// issued as m8n8k4 FP64 MMAs: one instruction does 256 FMAs, so the FP64 pipe runs at its peak rate with an eighth
// of the issue slots the scalar DFMA form needs (ncu on the DFMA form: FP64 pipe 53 %, issue-bound).
// Staging: a 3-stage ring filled by TMA 1-D bulk copies (cp.async.bulk, one elected thread): per stage 8 cells, i.e. 8
// rows of the CTA's D slice (2 KB each, contiguous in Dcol) and ONE block of responsibilities (gamma is kept cell-major
// [Cp][gw], so the 8 cells x gw weights of a stage are contiguous); completion is signalled on the stage's mbarrier, so
// no thread spends issue slots on address arithmetic or per-thread copies.  Row strides (D rows in shared memory, gw)
// are 8 mod 16 doubles so that the 4 x 8 fragment loads take the minimum of two shared-memory wavefronts.
template <int MT8>
__global__ void __launch_bounds__(kMsThreads, 3) mstats_kernel(const double *__restrict__ Dcol, int Epad,
                                                               const double *__restrict__ gam, int gw, int M, int MT,
                                                               const int32_t *__restrict__ chunk_cp0,
                                                               const int32_t *__restrict__ chunk_n,
                                                               double *__restrict__ partial)
{
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ __align__(8) uint64_t s_full[kStages];
    const int tpb = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ESTR = 2 * tpb + 8;                                           // == 8 (mod 16): tpb is a multiple of 32
    double *sD = reinterpret_cast<double *>(s_dyn);                         // [kStages][kStCells][ESTR]
    double *sG = sD + (size_t)kStages * kStCells * ESTR;                    // [kStages][kStCells][gw]
    const int chunk = blockIdx.z;
    const int m0 = blockIdx.x * MT;                     // groups vary fastest: the CTAs that re-read a tile run together
    const int etile0 = 2 * blockIdx.y * tpb;
    const int cp0 = chunk_cp0[chunk];
    const int n = chunk_n[chunk];                       // multiple of 64
    const uint32_t row_bytes = (uint32_t)(min(2 * tpb, Epad - etile0) * (int)sizeof(double));   // multiple of 16 (Epad even)
    const uint32_t gam_bytes = (uint32_t)(kStCells * gw) * (uint32_t)sizeof(double);
    const uint32_t stage_bytes = kStCells * row_bytes + (gam ? gam_bytes : 0u);
    double acc[8][MT8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int mt = 0; mt < MT8; ++mt) { acc[nt][mt][0] = 0.0; acc[nt][mt][1] = 0.0; }
    // no weights (unweighted doMean): constant rows, written once (generic proxy) and never touched by the copies
    if (!gam)
        for (int i = tid; i < kStages * kStCells * gw; i += tpb) sG[i] = (i % gw) < M ? 1.0 : 0.0;
    if (tid == 0) {
        for (int s2 = 0; s2 < kStages; ++s2) mbar_init(&s_full[s2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const double *Dsrc = Dcol + (size_t)cp0 * Epad + etile0;

    auto issue = [&](int st, int c0) {                  // called by thread 0 only
        if (c0 >= n) return;
        mbar_expect_tx(&s_full[st], stage_bytes);
        for (int j = 0; j < kStCells; ++j)
            tma_load_1d(&sD[(size_t)(st * kStCells + j) * ESTR], Dsrc + (size_t)(c0 + j) * Epad, row_bytes, &s_full[st]);
        if (gam) tma_load_1d(&sG[st * kStCells * gw], gam + (size_t)(cp0 + c0) * gw, gam_bytes, &s_full[st]);
    };

    if (tid == 0) { issue(0, 0); issue(1, kStCells); }
    int st = 0;
    uint32_t parity = 0;                                // phase of the barrier of stage `st`: flips each time the ring wraps
    const int kq = lane & 3, gq = lane >> 2;            // k index / row-or-column index of this lane's fragment element
    for (int c0 = 0; c0 < n; c0 += kStCells) {
        // stage (st + 2) % 3 was consumed in the previous iteration (everyone passed the barrier at its end)
        if (tid == 0) issue(st == 0 ? 2 : st - 1, c0 + 2 * kStCells);
        mbar_wait(&s_full[st], parity);
#pragma unroll
        for (int ks = 0; ks < kStCells / 4; ++ks) {
            const double *gd = sG + (st * kStCells + ks * 4 + kq) * gw + m0 + gq;
            const double *dd = sD + (size_t)(st * kStCells + ks * 4 + kq) * ESTR + warp * kMsWarpE + gq;
            double a[MT8];
#pragma unroll
            for (int mt = 0; mt < MT8; ++mt) a[mt] = gd[mt * 8];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const double b = dd[nt * 8];
#pragma unroll
                for (int mt = 0; mt < MT8; ++mt) dmma884(acc[nt][mt][0], acc[nt][mt][1], a[mt], b);
            }
        }
        __syncthreads();                                // the stage may be refilled
        if (st == 2) { st = 0; parity ^= 1u; } else ++st;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int eo = etile0 + warp * kMsWarpE + nt * 8 + 2 * kq;
#pragma unroll
        for (int mt = 0; mt < MT8; ++mt) {
            const int m = mt * 8 + gq;
            if (eo < Epad && m < MT && m0 + m < M)
                *reinterpret_cast<double2 *>(partial + ((size_t)chunk * M + m0 + m) * Epad + eo) =
                    make_double2(acc[nt][mt][0], acc[nt][mt][1]);
        }
    }
}

// R[m][s][e] (E x S column-major per pair) = sum over the groups g that contain S-gene s (ascending) of
// w_g * (sum over the chunks of segment g, ascending), w_g = 1 / |g|: (D diag gamma) %*% t(Rho') with the columns of
// Rho' normalised to sum 1, R/mnems_low.r:896-904.  Single knock-down data: one group per S-gene, w = 1, and the
// result is the plain ascending chunk sum (0 + x * 1 is exact).
__global__ void mstats_reduce_kernel(const double *__restrict__ partial, int M, int E, int Epad, int S,
                                     const int32_t *__restrict__ seg_chunk0, const int32_t *__restrict__ sg_ptr,
                                     const int32_t *__restrict__ sg_idx, const double *__restrict__ gweight,
                                     double *__restrict__ R)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y, m = blockIdx.z;
    if (e >= E) return;
    double tot = 0.0;
    for (int i = sg_ptr[s]; i < sg_ptr[s + 1]; ++i) {
        const int g = sg_idx[i];
        double acc = 0.0;
        for (int ch = seg_chunk0[g]; ch < seg_chunk0[g + 1]; ++ch) acc += partial[((size_t)ch * M + m) * Epad + e];
        tot += acc * gweight[g];
    }
    R[((size_t)m * S + s) * E + e] = tot;
}

// Rg[m][g][e] = sum over the chunks of segment g (un-normalised): the per-group collapsed data the theta-free E-step of
// the Rho path needs, because min(1, t(Rho) %*% adj1) is an OR over a cell's S-genes, not a sum.
__global__ void mstats_groups_kernel(const double *__restrict__ partial, int M, int E, int Epad, int G,
                                     const int32_t *__restrict__ seg_chunk0, double *__restrict__ Rg)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y, m = blockIdx.z;
    if (e >= E) return;
    double acc = 0.0;
    for (int ch = seg_chunk0[g]; ch < seg_chunk0[g + 1]; ++ch) acc += partial[((size_t)ch * M + m) * Epad + e];
    Rg[((size_t)m * G + g) * E + e] = acc;
}

// theta-free E-step on the Rho path, R/mnems_low.r:617-619: theta[m][e] = first argmax_j over
// ((D diag gamma_m) %*% cbind(adj1, 0))[e, j], adj1[c, j] = OR_{s in group(c)} closure_m[s, j]; the cells of a group
// share their row of adj1, so W[e, j] = sum_{g : gmask_g & col_j != 0} Rg[m][g][e] (ascending g).  Null column LAST.
__global__ void __launch_bounds__(128) theta_groups_kernel(const double *__restrict__ Rg, int E, int S, int G,
                                                           const uint32_t *__restrict__ gmask,
                                                           const uint32_t *__restrict__ clo_cols,
                                                           int32_t *__restrict__ theta)
{
    __shared__ uint32_t s_col[32];
    const int m = blockIdx.y, tid = threadIdx.x;
    if (tid < 32) s_col[tid] = clo_cols[m * 32 + tid];
    __syncthreads();
    const int e = blockIdx.x * blockDim.x + tid;
    if (e >= E) return;
    const double *r = Rg + (size_t)m * G * E + e;
    double best = -INFINITY;
    int arg = 0;
    for (int j = 0; j < S; ++j) {
        const uint32_t col = s_col[j];
        double w = 0.0;
        for (int g = 0; g < G; ++g)
            if (gmask[g] & col) w += r[(size_t)g * E];
        if (w > best) { best = w; arg = j + 1; }
    }
    if (0.0 > best) arg = 0;
    theta[(size_t)m * E + e] = arg;
}

int launch_mstats_groups(const mnemcu_ctx *cx, const double *partial, int M, double *Rg, cudaStream_t st)
{
    dim3 grid((cx->E + 127) / 128, cx->G, M);
    MN_LAUNCH(mstats_groups_kernel, grid, 128, 0, st, partial, M, cx->E, cx->Epad, cx->G, cx->seg_chunk0, Rg);
    return 0;
}

int launch_theta_groups(const mnemcu_ctx *cx, const double *Rg, int M, const uint32_t *clo_cols, int32_t *theta,
                        cudaStream_t st)
{
    dim3 grid((cx->E + 127) / 128, M);
    MN_LAUNCH(theta_groups_kernel, grid, 128, 0, st, Rg, cx->E, cx->S, cx->G, cx->gmask, clo_cols, theta);
    return 0;
}

size_t mstats_partial_elems(const mnemcu_ctx *cx, int M) { return (size_t)cx->nchunk * M * cx->Epad; }

template <int MT8>
static int launch_mstats_t(const mnemcu_ctx *cx, const double *gam, int M, int MT, double *partial, cudaStream_t st)
{
    const int pairs = cx->Epad / 2;                    // E-gene pairs: one per staging thread
    int tpb = kMsThreads;
    while (tpb > 32 && tpb / 2 >= pairs) tpb /= 2;
    dim3 grid((M + MT - 1) / MT, (pairs + tpb - 1) / tpb, cx->nchunk);
    const int gw = gamma_stride(M);
    const size_t smem = (size_t)kStages * kStCells * ((size_t)(2 * tpb + 8) + gw) * sizeof(double);
    // per device / context, see launch_estep_t
    MN_CU(cudaFuncSetAttribute(mstats_kernel<MT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    MN_LAUNCH(mstats_kernel<MT8>, grid, tpb, smem, st, cx->Dcol, cx->Epad, gam, gw, M, MT, cx->chunk_cp0, cx->chunk_n,
              partial);
    return 0;
}

int launch_mstats(const mnemcu_ctx *cx, const double *gam, int M, double *partial, double *R, cudaStream_t st)
{
    MN_ARG(M >= 1 && M <= 32, "M-step statistics serve 1..32 pairs per launch");
    // pairs per CTA: up to 24 (three 8-pair MMA tiles; 48 FP64 accumulators per lane)
    const int groups = (M + 23) / 24;
    const int MT = (M + groups - 1) / groups;
    int rc;
    switch ((MT + 7) / 8) {
        case 1: rc = launch_mstats_t<1>(cx, gam, M, MT, partial, st); break;
        case 2: rc = launch_mstats_t<2>(cx, gam, M, MT, partial, st); break;
        default: rc = launch_mstats_t<3>(cx, gam, M, MT, partial, st); break;
    }
    if (rc) return rc;
    dim3 grid((cx->E + 127) / 128, cx->S, M);
    MN_LAUNCH(mstats_reduce_kernel, grid, 128, 0, st, partial, M, cx->E, cx->Epad, cx->S, cx->seg_chunk0, cx->sg_ptr,
              cx->sg_idx, cx->gweight, R);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// k x C column-major in the caller's cell order  <->  [k][Cp] sorted by perturbation
// ---------------------------------------------------------------------------------------------------------
__global__ void L_to_internal_kernel(const double *__restrict__ Lh, int k, int Cp, const int32_t *__restrict__ orig,
                                     double *__restrict__ Li)
{
    const int cp = blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= Cp) return;
    const int c = orig[cp];
    for (int i = 0; i < k; ++i) Li[(size_t)i * Cp + cp] = c >= 0 ? Lh[(size_t)c * k + i] : 0.0;
}
__global__ void L_from_internal_kernel(const double *__restrict__ Li, int k, int C, int Cp,
                                       const int32_t *__restrict__ pos, double *__restrict__ Lh)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int cp = pos[c];
    for (int i = 0; i < k; ++i) Lh[(size_t)c * k + i] = Li[(size_t)i * Cp + cp];
}

int launch_L_to_internal(const mnemcu_ctx *cx, const double *Lh, int k, double *Li, cudaStream_t st)
{
    MN_LAUNCH(L_to_internal_kernel, (cx->Cp + 255) / 256, 256, 0, st, Lh, k, cx->Cp, cx->orig, Li);
    return 0;
}
int launch_L_from_internal(const mnemcu_ctx *cx, const double *Li, int k, double *Lh, cudaStream_t st)
{
    MN_LAUNCH(L_from_internal_kernel, (cx->C + 255) / 256, 256, 0, st, Li, k, cx->C, cx->Cp, cx->pos, Lh);
    return 0;
}

}  // namespace mn
