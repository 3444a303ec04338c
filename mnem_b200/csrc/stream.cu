// The two HBM-streaming kernels of the EM loop and their small helpers.
//
//   estep_kernel   replaces the body of getProbs, R/mnems_low.r:606-638:  L[i,c] = colSums(data * t(adj2))
//   mstats_kernel  replaces doMean on the Rho path, R/mnems_low.r:893-906 reached from nem() R/mnems.r:101-107
//
// Both read the expression matrix exactly once per launch and serve up to 32 (start, component) pairs, so
// the algorithmic traffic is 8*E*C bytes in plus the small outputs.  FP64 throughout.
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
// grid = (nblk, groups of MT pairs), block = 32 * esplit threads; warp w covers E-genes [w*eper, (w+1)*eper).
// ---------------------------------------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(256) estep_kernel(const double *__restrict__ Dblk, int E,
                                                    const uint32_t *__restrict__ maskT,
                                                    const int32_t *__restrict__ blk_seg, int M, int eper,
                                                    PairPtrs out)
{
    extern __shared__ double2 s_part[];   // [esplit-1][MT][32] partial sums of warps 1..
    const int blk = blockIdx.x;
    const int m0 = blockIdx.y * MT;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarp = blockDim.x >> 5;
    const int seg = blk_seg[blk];
    const int ebeg = warp * eper;
    const int eend = min(E, ebeg + eper);
    const double2 *base = reinterpret_cast<const double2 *>(Dblk + (size_t)blk * E * kBlkCells) + lane;
    const uint32_t *mk = maskT + (size_t)seg * E;

    double2 acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = make_double2(0.0, 0.0);

    for (int e0 = ebeg; e0 < eend; e0 += 32) {
        uint32_t wv = (e0 + lane < eend) ? (__ldg(mk + e0 + lane) >> m0) : 0u;
        const int n = min(32, eend - e0);
        if (n == 32) {
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
                        if ((w >> m) & 1u) { acc[m].x += d[j].x; acc[m].y += d[j].y; }
                }
            }
        } else {
            for (int j = 0; j < n; ++j) {
                const double2 d = ld_stream2(base + (size_t)(e0 + j) * 32);
                const uint32_t w = __shfl_sync(kFull, wv, j);
#pragma unroll
                for (int m = 0; m < MT; ++m)
                    if ((w >> m) & 1u) { acc[m].x += d.x; acc[m].y += d.y; }
            }
        }
    }
    if (nwarp > 1) {   // combine the E-gene ranges in ascending order (fixed shape => deterministic)
        if (warp > 0) {
#pragma unroll
            for (int m = 0; m < MT; ++m) s_part[((warp - 1) * MT + m) * 32 + lane] = acc[m];
        }
        __syncthreads();
        if (warp == 0) {
            for (int w = 1; w < nwarp; ++w)
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const double2 p = s_part[((w - 1) * MT + m) * 32 + lane];
                    acc[m].x += p.x;
                    acc[m].y += p.y;
                }
        }
    }
    if (warp == 0) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            if (m0 + m < M) {   // warp-uniform
                reinterpret_cast<double2 *>(out.L[m0 + m] + (size_t)blk * kBlkCells)[lane] = acc[m];
                const bool pos = acc[m].x > 0.0 || acc[m].y > 0.0;
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
    dim3 grid(cx->nblk, groups);
    size_t smem = esplit > 1 ? (size_t)(esplit - 1) * MT * 32 * sizeof(double2) : 0;
    MN_LAUNCH(estep_kernel<MT>, grid, 32 * esplit, smem, st, cx->Dblk, cx->E, maskT, cx->blk_seg, M, eper, out);
    return 0;
}

int launch_estep(const mnemcu_ctx *cx, const uint32_t *maskT, int M, const PairPtrs &out, cudaStream_t st)
{
    MN_ARG(M >= 1 && M <= 32, "E-step serves 1..32 (start, component) pairs per launch");
    // Pairs per warp: at most 10 (20 FP64 accumulators per lane); split evenly above that.  The second group
    // re-reads the block from L1/L2, not from HBM.
    const int groups = (M + 9) / 10;
    const int MT = (M + groups - 1) / groups;
    // Few blocks (small C): split the E-gene range over up to 8 warps so that every SM still has loads in flight.
    int esplit = 1;
    const long warps = (long)cx->nblk * groups;
    while (esplit < 8 && warps * esplit < 16L * cx->sm_count && cx->E / (esplit * 2) >= 64) esplit *= 2;
    switch (MT) {
        case 1: return launch_estep_t<1>(cx, maskT, M, out, esplit, st);
        case 2: return launch_estep_t<2>(cx, maskT, M, out, esplit, st);
        case 3: return launch_estep_t<3>(cx, maskT, M, out, esplit, st);
        case 4: return launch_estep_t<4>(cx, maskT, M, out, esplit, st);
        case 5: return launch_estep_t<5>(cx, maskT, M, out, esplit, st);
        case 6: return launch_estep_t<6>(cx, maskT, M, out, esplit, st);
        case 7: return launch_estep_t<7>(cx, maskT, M, out, esplit, st);
        case 8: return launch_estep_t<8>(cx, maskT, M, out, esplit, st);
        case 9: return launch_estep_t<9>(cx, maskT, M, out, esplit, st);
        default: return launch_estep_t<10>(cx, maskT, M, out, esplit, st);
    }
}

// maskT[s][e]: bit m = phi_m closure [s, theta_m(e)]  (adj1[, subtopo] of R/mnems_low.r:621-624, theta 0 -> null column)
__global__ void build_masks_kernel(const uint32_t *__restrict__ clo, const int32_t *__restrict__ theta, int M, int E,
                                   int S, uint32_t *__restrict__ maskT)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (e >= E) return;
    uint32_t w = 0;
    for (int m = 0; m < M; ++m) {
        const int t = theta[(size_t)m * E + e];
        if (t > 0 && t <= S && ((clo[m * 32 + s] >> (t - 1)) & 1u)) w |= 1u << m;
    }
    maskT[(size_t)s * E + e] = w;
}

int launch_build_masks(const mnemcu_ctx *cx, const uint32_t *clo, const int32_t *theta, int M, uint32_t *maskT,
                       cudaStream_t st)
{
    dim3 grid((cx->E + 127) / 128, cx->S);
    MN_LAUNCH(build_masks_kernel, grid, 128, 0, st, clo, theta, M, cx->E, cx->S, maskT);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// M-step sufficient statistics.  Layout Dcol[cp][Epad]: a cell's E-genes are contiguous, thread t owns the
// E-gene pair (2t, 2t+1) and walks the cells of one chunk of one segment in ascending order, so loads are
// coalesced 16-byte reads and the responsibility gamma[m][cp] is warp-uniform (broadcast from shared memory).
// Chunk partials are combined in ascending chunk order by mstats_reduce_kernel: no atomics, bit-reproducible.
//
// grid = (E-gene tiles, chunks, groups of MT pairs), block = tpb threads.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSub = 64;   // cells whose weights are staged in shared memory at a time

template <int MT>
__global__ void __launch_bounds__(256) mstats_kernel(const double *__restrict__ Dcol, int Epad,
                                                     const double *__restrict__ gam, int M, int Cp,
                                                     const int32_t *__restrict__ chunk_cp0,
                                                     const int32_t *__restrict__ chunk_n, double *__restrict__ partial)
{
    __shared__ double s_g[kSub][MT];
    const int chunk = blockIdx.y;
    const int m0 = blockIdx.z * MT;
    const int e = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const bool live = e < Epad;
    const int cp0 = chunk_cp0[chunk];
    const int n = chunk_n[chunk];
    double2 acc[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[m] = make_double2(0.0, 0.0);
    for (int c0 = 0; c0 < n; c0 += kSub) {
        const int nc = min(kSub, n - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < kSub * MT; i += blockDim.x) {
            const int m = i / kSub, c = i - m * kSub;     // consecutive threads read consecutive cells
            double g = 0.0;
            if (c < nc && m0 + m < M) g = gam ? gam[(size_t)(m0 + m) * Cp + cp0 + c0 + c] : 1.0;
            s_g[c][m] = g;
        }
        __syncthreads();
        if (live) {
            const double2 *p = reinterpret_cast<const double2 *>(Dcol + (size_t)(cp0 + c0) * Epad + e);
            const size_t stride = (size_t)Epad / 2;
            int c = 0;
            for (; c + 4 <= nc; c += 4) {
                double2 d[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) d[j] = ld_stream2(p + (size_t)(c + j) * stride);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const double g = s_g[c + j][m];
                        acc[m].x = fma(d[j].x, g, acc[m].x);
                        acc[m].y = fma(d[j].y, g, acc[m].y);
                    }
            }
            for (; c < nc; ++c) {
                const double2 d = ld_stream2(p + (size_t)c * stride);
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    const double g = s_g[c][m];
                    acc[m].x = fma(d.x, g, acc[m].x);
                    acc[m].y = fma(d.y, g, acc[m].y);
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int m = 0; m < MT; ++m)
            if (m0 + m < M)
                *reinterpret_cast<double2 *>(partial + ((size_t)chunk * M + m0 + m) * Epad + e) = acc[m];
    }
}

// R[m][s][e] (E x S column-major per pair) = sum over the chunks of segment s, ascending
__global__ void mstats_reduce_kernel(const double *__restrict__ partial, int M, int E, int Epad, int S,
                                     const int32_t *__restrict__ seg_chunk0, double *__restrict__ R)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y, m = blockIdx.z;
    if (e >= E) return;
    double acc = 0.0;
    for (int ch = seg_chunk0[s]; ch < seg_chunk0[s + 1]; ++ch) acc += partial[((size_t)ch * M + m) * Epad + e];
    R[((size_t)m * S + s) * E + e] = acc;
}

size_t mstats_partial_elems(const mnemcu_ctx *cx, int M) { return (size_t)cx->nchunk * M * cx->Epad; }

template <int MT>
static int launch_mstats_t(const mnemcu_ctx *cx, const double *gam, int M, double *partial, cudaStream_t st)
{
    const int pairs = cx->Epad / 2;
    int tpb = 256;
    while (tpb > 32 && tpb / 2 >= pairs) tpb /= 2;
    dim3 grid((pairs + tpb - 1) / tpb, cx->nchunk, (M + MT - 1) / MT);
    MN_LAUNCH(mstats_kernel<MT>, grid, tpb, 0, st, cx->Dcol, cx->Epad, gam, M, cx->Cp, cx->chunk_cp0, cx->chunk_n,
              partial);
    return 0;
}

int launch_mstats(const mnemcu_ctx *cx, const double *gam, int M, double *partial, double *R, cudaStream_t st)
{
    MN_ARG(M >= 1 && M <= 32, "M-step statistics serve 1..32 pairs per launch");
    const int groups = (M + 9) / 10;
    const int MT = (M + groups - 1) / groups;
    int rc;
    switch (MT) {
        case 1: rc = launch_mstats_t<1>(cx, gam, M, partial, st); break;
        case 2: rc = launch_mstats_t<2>(cx, gam, M, partial, st); break;
        case 3: rc = launch_mstats_t<3>(cx, gam, M, partial, st); break;
        case 4: rc = launch_mstats_t<4>(cx, gam, M, partial, st); break;
        case 5: rc = launch_mstats_t<5>(cx, gam, M, partial, st); break;
        case 6: rc = launch_mstats_t<6>(cx, gam, M, partial, st); break;
        case 7: rc = launch_mstats_t<7>(cx, gam, M, partial, st); break;
        case 8: rc = launch_mstats_t<8>(cx, gam, M, partial, st); break;
        case 9: rc = launch_mstats_t<9>(cx, gam, M, partial, st); break;
        default: rc = launch_mstats_t<10>(cx, gam, M, partial, st); break;
    }
    if (rc) return rc;
    dim3 grid((cx->E + 127) / 128, cx->S, M);
    MN_LAUNCH(mstats_reduce_kernel, grid, 128, 0, st, partial, M, cx->E, cx->Epad, cx->S, cx->seg_chunk0, R);
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
