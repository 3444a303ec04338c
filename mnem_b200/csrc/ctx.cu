// Device-resident data context: uploads (or synthesises) the E x C expression matrix once and keeps it in the
// two HBM layouts the streaming kernels want (DESIGN.md "Data layout in HBM").
#include <algorithm>
#include <thread>

#include "common.cuh"

namespace mn {

// ---- synthetic data: counter-based, identical on host and device ------------------------------------------
// noise(seed, idx) = (sum of twelve 16-bit uniforms - 6*65535) / 65536  [Irwin-Hall(12): mean 0, variance ~1]
//                  + (a 53-bit uniform - 0.5) / 65536                   [fills the mantissa below 2^-16]
// Every operation is a single correctly rounded IEEE operation on exactly representable operands (no libm), so host
// and device agree bit for bit; the second term makes the values full-precision doubles, so that sums over them round
// differently in different orders (with multiples of 2^-16 alone every E-step sum would be exact in ANY order and the
// parity of summation orders could not be observed).
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double syn_noise(uint64_t seed, uint64_t idx)
{
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uint64_t w = mix64(seed + 0x9E3779B97F4A7C15ULL * (3 * idx + j + 1));
        sum += (uint32_t)(w & 0xffff) + (uint32_t)((w >> 16) & 0xffff) + (uint32_t)((w >> 32) & 0xffff) +
               (uint32_t)(w >> 48);
    }
    const uint64_t w4 = mix64((seed ^ 0xA0761D6478BD642FULL) + 0xE7037ED1A0B428DBULL * (idx + 1));
    const double coarse = ((double)sum - 393210.0) * (1.0 / 65536.0);                       // exact
    const double fine = ((double)(w4 >> 11) * (1.0 / 9007199254740992.0) - 0.5) * (1.0 / 65536.0);   // exact
    return coarse + fine;                                                                   // one rounding
}
// D[e,c] = (bin - 0.5)/0.5 + noise (README.md:38-39); bin = phi_true[comp(c)][pert(c), theta_true[comp(c)][e]]
// (simData R/mnems.r:3084-3138), uninformative rows (theta 0) are fair coin flips (:3147-3151).
__host__ __device__ __forceinline__ double syn_value(int e, int64_t c, int E, int pert1, int comp, int S,
                                                     const uint8_t *phi_true, const int32_t *theta_true,
                                                     uint64_t seed)
{
    const uint64_t idx = (uint64_t)e + (uint64_t)E * (uint64_t)c;
    const int t = theta_true[(size_t)comp * E + e];
    int bin;
    if (t > 0) bin = phi_true[(size_t)comp * S * S + (pert1 - 1) + (size_t)(t - 1) * S] != 0;
    else bin = (int)(mix64(seed ^ 0xD1B54A32D192ED03ULL ^ (idx * 0x2545F4914F6CDD1DULL)) & 1);
    return (bin ? 1.0 : -1.0) + syn_noise(seed, idx);
}

__global__ void syn_fill_kernel(double *__restrict__ Dcol, int E, int Epad, int Cp, const int32_t *__restrict__ orig,
                                const int32_t *__restrict__ pert, const uint8_t *__restrict__ comp, int S,
                                const uint8_t *__restrict__ phi_true, const int32_t *__restrict__ theta_true,
                                uint64_t seed)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int cp = blockIdx.y;
    if (e >= Epad) return;
    const int c = orig[cp];
    double v = 0.0;
    if (c >= 0 && e < E) v = syn_value(e, c, E, pert[c], comp[c], S, phi_true, theta_true, seed);
    Dcol[(size_t)cp * Epad + e] = v;
}

// staging[c_local][E] (a slab of original columns) -> Dcol[pos[c]][Epad]
__global__ void scatter_cols_kernel(const double *__restrict__ slab, int E, int Epad, int c0, int nc,
                                    const int32_t *__restrict__ pos, double *__restrict__ Dcol)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int cl = blockIdx.y;
    if (e >= E || cl >= nc) return;
    Dcol[(size_t)pos[c0 + cl] * Epad + e] = slab[(size_t)cl * E + e];
}

// Dcol[cp][Epad] -> Dblk[blk][e][64]   (32 E-genes x 64 cells tile through shared memory)
__global__ void transpose_kernel(const double *__restrict__ Dcol, int E, int Epad, double *__restrict__ Dblk)
{
    __shared__ double tile[kBlkCells][33];
    const int blk = blockIdx.x;
    const int e0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 256 threads: 8 rows of 32
    for (int c = ty; c < kBlkCells; c += 8) {
        const int e = e0 + tx;
        tile[c][tx] = e < E ? Dcol[((size_t)blk * kBlkCells + c) * Epad + e] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * kBlkCells; i += 256) {
        const int el = i / kBlkCells, c = i - el * kBlkCells;
        if (e0 + el < E) Dblk[((size_t)blk * E + e0 + el) * kBlkCells + c] = tile[c][el];
    }
}

__global__ void gather_cols_kernel(const double *__restrict__ Dcol, int E, int Epad, int c0, int nc,
                                   const int32_t *__restrict__ pos, double *__restrict__ slab)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int cl = blockIdx.y;
    if (e >= E || cl >= nc) return;
    slab[(size_t)cl * E + e] = Dcol[(size_t)pos[c0 + cl] * Epad + e];
}

// cellmask[c]: set of knocked-down S-genes of cell c (bit s); exactly one bit for ordinary data
static int ctx_layout(mnemcu_ctx *cx, int E, int C, const std::vector<uint32_t> &cellmask, int S, int device)
{
    MN_ARG(E >= 1 && C >= 1, "E and C must be positive");
    MN_ARG(S >= 1 && S <= MNEMCU_MAX_S, "S must be in 1..32");
    int ndev = 0;
    MN_CU(cudaGetDeviceCount(&ndev));
    MN_ARG(device >= 0 && device < ndev, "no such CUDA device");
    MN_CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    MN_CU(cudaGetDeviceProperties(&prop, device));
    cx->device = device;
    cx->sm_count = prop.multiProcessorCount;
    cx->E = E;
    cx->Epad = E + (E & 1);
    cx->C = C;
    cx->S = S;
    // groups: the S single knock-downs first, then the other distinct sets in ascending mask order
    cx->h_gmask.clear();
    for (int s = 0; s < S; ++s) cx->h_gmask.push_back(1u << s);
    {
        std::vector<uint32_t> extra;
        for (int c = 0; c < C; ++c)
            if (__builtin_popcount(cellmask[c]) != 1) extra.push_back(cellmask[c]);
        std::sort(extra.begin(), extra.end());
        extra.erase(std::unique(extra.begin(), extra.end()), extra.end());
        cx->h_gmask.insert(cx->h_gmask.end(), extra.begin(), extra.end());
        cx->multi = !extra.empty();
    }
    const int G = cx->G = (int)cx->h_gmask.size();
    auto group_of = [&](uint32_t m) -> int {
        if (__builtin_popcount(m) == 1) return __builtin_ctz(m);
        return (int)(std::lower_bound(cx->h_gmask.begin() + S, cx->h_gmask.end(), m) - cx->h_gmask.begin());
    };
    std::vector<int32_t> cell_group(C);
    std::vector<int64_t> cnt(G, 0);
    for (int c = 0; c < C; ++c) {
        cell_group[c] = group_of(cellmask[c]);
        cnt[cell_group[c]]++;
    }
    cx->seg_start.assign(G + 1, 0);
    int64_t cp = 0;
    for (int g = 0; g < G; ++g) {
        cx->seg_start[g] = (int32_t)cp;
        cp += (cnt[g] + kBlkCells - 1) / kBlkCells * kBlkCells;
    }
    MN_ARG(cp < (int64_t)1 << 31, "too many cells");
    cx->seg_start[G] = (int32_t)cp;
    cx->Cp = (int)cp;
    cx->nblk = cx->Cp / kBlkCells;
    cx->h_pos.assign(C, 0);
    cx->h_orig.assign(std::max(cx->Cp, 1), -1);
    std::vector<int32_t> fill(cx->seg_start.begin(), cx->seg_start.end() - 1);
    for (int c = 0; c < C; ++c) {   // stable: original order inside a segment
        const int p = fill[cell_group[c]]++;
        cx->h_pos[c] = p;
        cx->h_orig[p] = c;
    }
    cx->h_blk_seg.assign(std::max(cx->nblk, 1), 0);
    for (int g = 0; g < G; ++g)
        for (int b = cx->seg_start[g] / kBlkCells; b < cx->seg_start[g + 1] / kBlkCells; ++b) cx->h_blk_seg[b] = g;
    {   // device tables of the groups
        std::vector<double> gw(G);
        for (int g = 0; g < G; ++g) gw[g] = 1.0 / std::max(1, __builtin_popcount(cx->h_gmask[g]));
        std::vector<int32_t> ptr(S + 1, 0), idx;
        for (int s = 0; s < S; ++s) {
            ptr[s] = (int32_t)idx.size();
            for (int g = 0; g < G; ++g)
                if ((cx->h_gmask[g] >> s) & 1u) idx.push_back(g);
        }
        ptr[S] = (int32_t)idx.size();
        MN_CU(cudaMalloc((void **)&cx->gmask, sizeof(uint32_t) * G));
        MN_CU(cudaMalloc((void **)&cx->gweight, sizeof(double) * G));
        MN_CU(cudaMalloc((void **)&cx->sg_ptr, sizeof(int32_t) * (S + 1)));
        MN_CU(cudaMalloc((void **)&cx->sg_idx, sizeof(int32_t) * std::max<size_t>(1, idx.size())));
        MN_CU(cudaMemcpy(cx->gmask, cx->h_gmask.data(), sizeof(uint32_t) * G, cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(cx->gweight, gw.data(), sizeof(double) * G, cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(cx->sg_ptr, ptr.data(), sizeof(int32_t) * (S + 1), cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(cx->sg_idx, idx.data(), sizeof(int32_t) * idx.size(), cudaMemcpyHostToDevice));
    }
    // chunks for the M-step statistics kernel: aim at >= 4 CTAs per SM, chunk a multiple of 64 cells
    {
        const int pairs = cx->Epad / 2;
        const int etiles = (pairs + 255) / 256;
        int64_t want = 4LL * cx->sm_count / std::max(1, etiles);
        int64_t cc = cx->Cp / std::max<int64_t>(1, want);
        cc = std::max<int64_t>(kBlkCells, std::min<int64_t>(4096, cc / kBlkCells * kBlkCells));
        cx->chunk_cells = (int)cc;
        std::vector<int32_t> cp0, cn, cs;
        cx->h_seg_chunk0.assign(G + 1, 0);
        for (int s = 0; s < G; ++s) {
            cx->h_seg_chunk0[s] = (int32_t)cp0.size();
            for (int p = cx->seg_start[s]; p < cx->seg_start[s + 1]; p += cx->chunk_cells) {
                cp0.push_back(p);
                cn.push_back(std::min(cx->chunk_cells, cx->seg_start[s + 1] - p));
                cs.push_back(s);
            }
        }
        cx->h_seg_chunk0[G] = (int32_t)cp0.size();
        cx->nchunk = (int)cp0.size();
        MN_CU(cudaMalloc((void **)&cx->chunk_cp0, sizeof(int32_t) * std::max(1, cx->nchunk)));
        MN_CU(cudaMalloc((void **)&cx->chunk_n, sizeof(int32_t) * std::max(1, cx->nchunk)));
        MN_CU(cudaMalloc((void **)&cx->chunk_seg, sizeof(int32_t) * std::max(1, cx->nchunk)));
        MN_CU(cudaMalloc((void **)&cx->seg_chunk0, sizeof(int32_t) * (G + 1)));
        MN_CU(cudaMemcpy(cx->chunk_cp0, cp0.data(), sizeof(int32_t) * cx->nchunk, cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(cx->chunk_n, cn.data(), sizeof(int32_t) * cx->nchunk, cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(cx->chunk_seg, cs.data(), sizeof(int32_t) * cx->nchunk, cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(cx->seg_chunk0, cx->h_seg_chunk0.data(), sizeof(int32_t) * (G + 1), cudaMemcpyHostToDevice));
    }
    MN_CU(cudaStreamCreateWithFlags(&cx->stream, cudaStreamNonBlocking));
    MN_CU(cudaMalloc((void **)&cx->pos, sizeof(int32_t) * C));
    MN_CU(cudaMalloc((void **)&cx->orig, sizeof(int32_t) * cx->h_orig.size()));
    MN_CU(cudaMalloc((void **)&cx->blk_seg, sizeof(int32_t) * cx->h_blk_seg.size()));
    MN_CU(cudaMemcpy(cx->pos, cx->h_pos.data(), sizeof(int32_t) * C, cudaMemcpyHostToDevice));
    MN_CU(cudaMemcpy(cx->orig, cx->h_orig.data(), sizeof(int32_t) * cx->h_orig.size(), cudaMemcpyHostToDevice));
    MN_CU(cudaMemcpy(cx->blk_seg, cx->h_blk_seg.data(), sizeof(int32_t) * cx->h_blk_seg.size(), cudaMemcpyHostToDevice));
    const size_t ncol = (size_t)cx->Cp * cx->Epad;
    const size_t nblkE = ((size_t)cx->nblk * E + 256) * kBlkCells;   // 256 E-genes of slack: the E-step reads whole groups of 32 per warp
    cudaError_t e1 = cudaMalloc((void **)&cx->Dcol, sizeof(double) * std::max<size_t>(1, ncol));
    if (e1 != cudaSuccess) return set_err(MNEMCU_ERR_NOMEM, "cudaMalloc Dcol (%zu bytes): %s", ncol * 8, cudaGetErrorString(e1));
    e1 = cudaMalloc((void **)&cx->Dblk, sizeof(double) * std::max<size_t>(1, nblkE));
    if (e1 == cudaSuccess) e1 = cudaMemset(cx->Dblk + (size_t)cx->nblk * E * kBlkCells, 0, sizeof(double) * 256 * kBlkCells);
    if (e1 == cudaSuccess) e1 = cudaDeviceSynchronize();      // legacy-stream memset vs the non-blocking work stream
    if (e1 != cudaSuccess) return set_err(MNEMCU_ERR_NOMEM, "cudaMalloc Dblk (%zu bytes): %s", nblkE * 8, cudaGetErrorString(e1));
    return 0;
}

static int ctx_finish(mnemcu_ctx *cx)
{
    dim3 grid(cx->nblk, (cx->E + 31) / 32);
    MN_LAUNCH(transpose_kernel, grid, 256, 0, cx->stream, cx->Dcol, cx->E, cx->Epad, cx->Dblk);
    MN_CU(cudaStreamSynchronize(cx->stream));
    return 0;
}

static void ctx_free(mnemcu_ctx *cx)
{
    if (!cx) return;
    cudaSetDevice(cx->device);
    cudaFree(cx->Dcol); cudaFree(cx->Dblk); cudaFree(cx->orig); cudaFree(cx->pos); cudaFree(cx->blk_seg);
    cudaFree(cx->chunk_cp0); cudaFree(cx->chunk_n); cudaFree(cx->chunk_seg); cudaFree(cx->seg_chunk0);
    cudaFree(cx->gmask); cudaFree(cx->gweight); cudaFree(cx->sg_ptr); cudaFree(cx->sg_idx);
    if (cx->stream) cudaStreamDestroy(cx->stream);
    delete cx;
}

}  // namespace mn

using namespace mn;

static int pert_to_masks(const int32_t *pert, int C, int S, std::vector<uint32_t> &out)
{
    MN_ARG(pert != nullptr, "pert is NULL");
    MN_ARG(S >= 1 && S <= MNEMCU_MAX_S && C >= 1, "bad S / C");
    out.resize(C);
    for (int c = 0; c < C; ++c) {
        if (pert[c] < 1 || pert[c] > S) return set_err(MNEMCU_ERR_ARG, "pert[%d] = %d outside 1..%d", c, pert[c], S);
        out[c] = 1u << (pert[c] - 1);
    }
    return 0;
}

// D on the device already (column-major E x C, caller's cell order): one scatter pass, no staging
static int ctx_from_device(const double *D_dev, int E, int C, const std::vector<uint32_t> &cellmask, int S, int device,
                           mnemcu_ctx **out)
{
    mnemcu_ctx *cx = new mnemcu_ctx();
    int rc = ctx_layout(cx, E, C, cellmask, S, device);
    if (rc) { ctx_free(cx); return rc; }
    auto body = [&]() -> int {
        MN_CU(cudaMemsetAsync(cx->Dcol, 0, sizeof(double) * (size_t)cx->Cp * cx->Epad, cx->stream));
        for (int c0 = 0; c0 < C; c0 += 32768) {
            const int nc = std::min(32768, C - c0);
            dim3 grid((E + 127) / 128, nc);
            MN_LAUNCH(scatter_cols_kernel, grid, 128, 0, cx->stream, D_dev + (size_t)c0 * E, E, cx->Epad, c0, nc, cx->pos,
                      cx->Dcol);
        }
        return ctx_finish(cx);
    };
    rc = body();
    if (rc) { ctx_free(cx); return rc; }
    *out = cx;
    return 0;
}

static int ctx_upload(const double *D, int E, int C, const std::vector<uint32_t> &cellmask, int S, int device,
                      mnemcu_ctx **out)
{
    mnemcu_ctx *cx = new mnemcu_ctx();
    int rc = ctx_layout(cx, E, C, cellmask, S, device);
    if (rc) { ctx_free(cx); return rc; }
    // zero the padding, then stream D up in slabs of original columns through two staging buffers
    cudaError_t ce = cudaMemsetAsync(cx->Dcol, 0, sizeof(double) * (size_t)cx->Cp * cx->Epad, cx->stream);
    if (ce != cudaSuccess) { ctx_free(cx); return set_err(MNEMCU_ERR_CUDA, "memset: %s", cudaGetErrorString(ce)); }
    const size_t slab_bytes = 256u << 20;
    int slab_cells = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((size_t)C, 32768), slab_bytes / (sizeof(double) * (size_t)E)));
    DevBuf<double> slab[2];
    cudaEvent_t ev[2] = {nullptr, nullptr};
    rc = slab[0].alloc((size_t)slab_cells * E);
    if (!rc) rc = slab[1].alloc((size_t)slab_cells * E);
    if (rc) { ctx_free(cx); return rc; }
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
    auto body = [&]() -> int {
        int which = 0;
        for (int c0 = 0; c0 < C; c0 += slab_cells, which ^= 1) {
            const int nc = std::min(slab_cells, C - c0);
            MN_CU(cudaEventSynchronize(ev[which]));   // the scatter that last read this slab is done
            MN_CU(cudaMemcpyAsync(slab[which].p, D + (size_t)c0 * E, sizeof(double) * (size_t)nc * E,
                                  cudaMemcpyHostToDevice, cx->stream));
            dim3 grid((E + 127) / 128, nc);
            MN_LAUNCH(scatter_cols_kernel, grid, 128, 0, cx->stream, slab[which].p, E, cx->Epad, c0, nc, cx->pos,
                      cx->Dcol);
            MN_CU(cudaEventRecord(ev[which], cx->stream));
        }
        return ctx_finish(cx);
    };
    rc = body();
    for (int i = 0; i < 2; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
    if (rc) { ctx_free(cx); return rc; }
    *out = cx;
    return 0;
}

extern "C" int mnemcu_ctx_create(const double *D, int E, int C, const int32_t *pert, int S, int device,
                                 mnemcu_ctx **out)
{
    MN_ARG(out != nullptr && D != nullptr, "NULL pointer");
    *out = nullptr;
    std::vector<uint32_t> cm;
    MN_TRY(pert_to_masks(pert, C, S, cm));
    return ctx_upload(D, E, C, cm, S, device, out);
}

extern "C" int mnemcu_ctx_create_device(const double *D_dev, int E, int C, const int32_t *pert, int S, int device,
                                        mnemcu_ctx **out)
{
    MN_ARG(out != nullptr && D_dev != nullptr, "NULL pointer");
    *out = nullptr;
    cudaPointerAttributes at;
    MN_CU(cudaPointerGetAttributes(&at, D_dev));
    MN_ARG(at.type == cudaMemoryTypeDevice && at.device == device, "D_dev must be device memory of `device`");
    std::vector<uint32_t> cm;
    MN_TRY(pert_to_masks(pert, C, S, cm));
    return ctx_from_device(D_dev, E, C, cm, S, device, out);
}

extern "C" int mnemcu_ctx_create_rho(const double *D, int E, int C, const double *Rho, int S, int device,
                                     mnemcu_ctx **out)
{
    MN_ARG(out != nullptr && D != nullptr && Rho != nullptr, "NULL pointer");
    MN_ARG(S >= 1 && S <= MNEMCU_MAX_S && C >= 1, "bad S / C");
    *out = nullptr;
    std::vector<uint32_t> cm(C, 0u);
    for (int c = 0; c < C; ++c)
        for (int s = 0; s < S; ++s)
            if (Rho[(size_t)c * S + s] != 0.0) cm[c] |= 1u << s;
    return ctx_upload(D, E, C, cm, S, device, out);
}

extern "C" int mnemcu_ctx_create_synthetic(int E, int C, const int32_t *pert, int S, const uint8_t *comp, int n_comp,
                                           const uint8_t *phi_true, const int32_t *theta_true, uint64_t seed,
                                           int device, mnemcu_ctx **out)
{
    MN_ARG(out && comp && phi_true && theta_true, "NULL pointer");
    MN_ARG(n_comp >= 1, "n_comp must be positive");
    *out = nullptr;
    for (int c = 0; c < C; ++c)
        if (comp[c] >= n_comp) return set_err(MNEMCU_ERR_ARG, "comp[%d] = %d outside 0..%d", c, comp[c], n_comp - 1);
    std::vector<uint32_t> cm;
    MN_TRY(pert_to_masks(pert, C, S, cm));
    mnemcu_ctx *cx = new mnemcu_ctx();
    int rc = ctx_layout(cx, E, C, cm, S, device);
    if (rc) { ctx_free(cx); return rc; }
    auto body = [&]() -> int {
        DevBuf<int32_t> d_pert, d_theta;
        DevBuf<uint8_t> d_comp, d_phi;
        MN_TRY(d_pert.alloc(C)); MN_TRY(d_comp.alloc(C));
        MN_TRY(d_phi.alloc((size_t)n_comp * S * S)); MN_TRY(d_theta.alloc((size_t)n_comp * E));
        MN_CU(cudaMemcpy(d_pert.p, pert, sizeof(int32_t) * C, cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(d_comp.p, comp, (size_t)C, cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(d_phi.p, phi_true, (size_t)n_comp * S * S, cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(d_theta.p, theta_true, sizeof(int32_t) * (size_t)n_comp * E, cudaMemcpyHostToDevice));
        for (int cp0 = 0; cp0 < cx->Cp; cp0 += 32768) {
            const int n = std::min(32768, cx->Cp - cp0);
            dim3 grid((cx->Epad + 127) / 128, n);
            MN_LAUNCH(syn_fill_kernel, grid, 128, 0, cx->stream, cx->Dcol + (size_t)cp0 * cx->Epad, E, cx->Epad, cx->Cp,
                      cx->orig + cp0, d_pert.p, d_comp.p, S, d_phi.p, d_theta.p, seed);
        }
        MN_TRY(ctx_finish(cx));
        return 0;
    };
    rc = body();
    if (rc) { ctx_free(cx); return rc; }
    *out = cx;
    return 0;
}

extern "C" int mnemcu_synthetic_fill_host(double *D, int E, int64_t c0, int64_t nc, const int32_t *pert, int S,
                                          const uint8_t *comp, int n_comp, const uint8_t *phi_true,
                                          const int32_t *theta_true, uint64_t seed, int nthreads)
{
    MN_ARG(D && pert && comp && phi_true && theta_true, "NULL pointer");
    MN_ARG(E >= 1 && nc >= 0 && c0 >= 0 && S >= 1 && S <= MNEMCU_MAX_S && n_comp >= 1, "bad size");
    if (nthreads < 1) nthreads = 1;
    auto work = [&](int t) {
        for (int64_t c = c0 + t; c < c0 + nc; c += nthreads) {
            double *col = D + (size_t)(c - c0) * E;
            for (int e = 0; e < E; ++e) col[e] = syn_value(e, c, E, pert[c], comp[c], S, phi_true, theta_true, seed);
        }
    };
    if (nthreads == 1) { work(0); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto &x : th) x.join();
    return 0;
}

extern "C" int mnemcu_ctx_destroy(mnemcu_ctx *ctx)
{
    ctx_free(ctx);
    return 0;
}

extern "C" int mnemcu_ctx_dims(const mnemcu_ctx *ctx, int *E, int *C, int *S)
{
    MN_ARG(ctx != nullptr, "ctx is NULL");
    if (E) *E = ctx->E;
    if (C) *C = ctx->C;
    if (S) *S = ctx->S;
    return 0;
}

extern "C" int mnemcu_ctx_read_data(mnemcu_ctx *cx, int64_t c0, int64_t nc, double *D_out)
{
    MN_ARG(cx && D_out, "NULL pointer");
    MN_ARG(c0 >= 0 && nc >= 0 && c0 + nc <= cx->C, "cell range outside the matrix");
    MN_CU(cudaSetDevice(cx->device));
    const int64_t step = std::max<int64_t>(1, std::min<int64_t>(32768, (64 << 20) / (8 * (int64_t)cx->E)));
    DevBuf<double> slab;
    MN_TRY(slab.alloc((size_t)std::min<int64_t>(step, std::max<int64_t>(nc, 1)) * cx->E));
    for (int64_t c = c0; c < c0 + nc; c += step) {
        const int n = (int)std::min<int64_t>(step, c0 + nc - c);
        dim3 grid((cx->E + 127) / 128, n);
        MN_LAUNCH(gather_cols_kernel, grid, 128, 0, cx->stream, cx->Dcol, cx->E, cx->Epad, (int)c, n, cx->pos, slab.p);
        MN_CU(cudaMemcpyAsync(D_out + (size_t)(c - c0) * cx->E, slab.p, sizeof(double) * (size_t)n * cx->E,
                              cudaMemcpyDeviceToHost, cx->stream));
        MN_CU(cudaStreamSynchronize(cx->stream));
    }
    return 0;
}
