// Internal declarations shared by the translation units of libmnemcu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/mnemcu.h"

namespace mn {

// ---- errors / bookkeeping -------------------------------------------------------------------------------
int set_err(int code, const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define MN_CU(call)                                                                                      \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return mn::set_err(MNEMCU_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
                               cudaGetErrorString(e_));                                                  \
    } while (0)

#define MN_ARG(cond, msg)                                                                                \
    do {                                                                                                 \
        if (!(cond)) return mn::set_err(MNEMCU_ERR_ARG, "%s:%d: bad argument: %s", __FILE__, __LINE__, msg); \
    } while (0)

// every kernel launch goes through this so that gpu_launches is a count, not a guess
#define MN_LAUNCH(kernel, grid, block, smem, stream, ...)                                                \
    do {                                                                                                 \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                      \
        mn::g_launches.fetch_add(1, std::memory_order_relaxed);                                          \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess)                                                                           \
            return mn::set_err(MNEMCU_ERR_CUDA, "%s:%d: launch %s -> %s", __FILE__, __LINE__, #kernel,   \
                               cudaGetErrorString(e_));                                                  \
    } while (0)

#define MN_TRY(expr)                                                                                     \
    do {                                                                                                 \
        int rc_ = (expr);                                                                                \
        if (rc_ != 0) return rc_;                                                                        \
    } while (0)

// RAII device buffer (freed on scope exit; never thrown across the ABI)
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count, bool zero = false)
    {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            // only a genuine out-of-memory is NOMEM; no device / no driver is a CUDA failure like everywhere else
            return set_err(e == cudaErrorMemoryAllocation ? MNEMCU_ERR_NOMEM : MNEMCU_ERR_CUDA, "cudaMalloc(%zu bytes): %s",
                           count * sizeof(T), cudaGetErrorString(e));
        }
        n = count;
        if (zero) {
            e = cudaMemset(p, 0, count * sizeof(T));
            if (e != cudaSuccess) return set_err(MNEMCU_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(e));
        }
        return 0;
    }
};

// An empty volatile asm inside a conditional block keeps the compiler from if-converting `if (bit) acc += x` into
// an unconditional FP64 add plus two 32-bit selects; ptxas then emits a predicated DADD (one issue slot, and no
// FP64-pipe work when the bit is clear).
#define MN_KEEP_BRANCH() asm volatile("")

constexpr int kBlkCells = 64;      // cells per streamed block (one warp, two cells per lane)
constexpr unsigned kFull = 0xffffffffu;

// ---- device-resident data (see DESIGN.md "Data layout in HBM") ----------------------------------------
}  // namespace mn

struct mnemcu_ctx {
    int device = 0;
    int E = 0, Epad = 0, C = 0, Cp = 0, S = 0, nblk = 0;
    // Cells are grouped by their SET of knocked-down S-genes.  Groups 0..S-1 are the single knock-downs (present even
    // when empty), further groups are the distinct multi-knock-down sets of the Rho path (R/mnems_low.r:308-324,
    // 613-616) in ascending mask order, plus the empty set if control cells exist.  G == S and multi == false for
    // ordinary single-knock-down data, and everything below then reduces to "one segment per S-gene".
    int G = 0;
    bool multi = false;
    std::vector<uint32_t> h_gmask;    // [G] bit s = S-gene s knocked down in the cells of the group
    uint32_t *gmask = nullptr;        // device copy
    double *gweight = nullptr;        // [G] 1 / max(1, popcount): the normalised Rho column of doMean, :897-902
    int32_t *sg_ptr = nullptr, *sg_idx = nullptr;   // CSR: groups that contain S-gene s, ascending (device)
    // cells sorted by group, every segment padded to a multiple of 64 cells (padding = zeros)
    double *Dcol = nullptr;        // [Cp][Epad]   a cell's E-genes contiguous       (M-step statistics kernel)
    double *Dblk = nullptr;        // [nblk][E][64] 64 cells contiguous per E-gene   (E-step kernel)
    int32_t *orig = nullptr;       // [Cp] original column of a sorted cell, -1 for padding
    int32_t *pos = nullptr;        // [C]  sorted position of an original column
    int32_t *blk_seg = nullptr;    // [nblk] group of each 64-cell block
    std::vector<int32_t> h_pos, h_orig, h_blk_seg;
    std::vector<int32_t> seg_start;   // [G+1] sorted/padded start of each segment
    // M-step statistics chunking: chunk j covers sorted cells [chunk_cp0[j], +chunk_n[j]) of one segment
    int chunk_cells = 0, nchunk = 0;
    int32_t *chunk_cp0 = nullptr, *chunk_n = nullptr, *chunk_seg = nullptr;
    int32_t *seg_chunk0 = nullptr;   // [G+1] first chunk of each segment
    std::vector<int32_t> h_seg_chunk0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
};

namespace mn {

// ---- bit-matrix helpers: one warp per graph, lane = row (or column), S <= 32 ------------------------------
__device__ __forceinline__ uint32_t warp_closure(uint32_t row, int S)
{   // Warshall, src/mm.cpp:15-30: for k: every row that reaches k absorbs row k
    for (int k = 0; k < S; ++k) {
        uint32_t rk = __shfl_sync(kFull, row, k);
        if ((row >> k) & 1u) row |= rk;
    }
    return row;
}

__device__ __forceinline__ uint32_t warp_reduce_graph(uint32_t row, int lane, int S)
{   // transitive.reduction R/mnems.r:517-532: closure (diag 1), zero diag, for y: rows x with g[x,y] drop row y
    row = warp_closure(row | (lane < S ? (1u << lane) : 0u), S);
    row &= ~(1u << lane);
    for (int y = 0; y < S; ++y) {
        uint32_t ry = __shfl_sync(kFull, row, y);
        if ((row >> y) & 1u) row &= ~ry;
    }
    return row;
}

__device__ __forceinline__ uint32_t warp_transpose(uint32_t row, int lane, int S)
{   // lane j receives column j: bit i = row_i bit j
    uint32_t col = 0;
    for (int j = 0; j < S; ++j) {
        uint32_t b = __ballot_sync(kFull, (row >> j) & 1u);
        if (lane == j) col = b;
    }
    return col;
}

// uint8 column-major S x S  <->  row masks (bit j of row i = adj[i + j*S])
__device__ __forceinline__ uint32_t load_row_mask(const uint8_t *g, int lane, int S)
{
    uint32_t r = 0;
    if (lane < S)
        for (int j = 0; j < S; ++j) r |= (g[lane + j * S] != 0 ? 1u : 0u) << j;
    return r;
}
__device__ __forceinline__ void store_row_mask(uint8_t *g, uint32_t row, int lane, int S)
{
    if (lane < S)
        for (int j = 0; j < S; ++j) g[lane + j * S] = (uint8_t)((row >> j) & 1u);
}

__device__ __forceinline__ double ld_stream(const double *p)
{
    return __ldcs(p);
}
__device__ __forceinline__ double2 ld_stream2(const double2 *p)
{
    return __ldcs(p);
}

// ---- stream.cu: the two HBM-streaming kernels and their helpers ------------------------------------------
struct PairPtrs {          // by-value kernel argument, indexed by (start, component) pair m
    double *L[32];         // [Cp] output row of pair m
    int32_t *flag[32];     // device flag that receives any(L > 0) for the mixture pair m belongs to
};
// E-step: out.L[m][cp] = sum_e Dblk[cp][e] * bit_m(maskT[seg(cp)][e]),  m < M <= 32.
int launch_estep(const mnemcu_ctx *cx, const uint32_t *maskT, int M, const PairPtrs &out, cudaStream_t st);
// maskT[s][e] bit m = closure_m[s][theta_m[e] - 1] (theta 0 -> 0).  clo: [M][32] row masks, theta: [M][E].
int launch_build_masks(const mnemcu_ctx *cx, const uint32_t *clo, const int32_t *theta, int M, uint32_t *maskT,
                       cudaStream_t st);
// M-step statistics: R[m][s][e] = sum_{cp in segment s} Dcol[cp][e] * gam[cp][m]; gam == nullptr: weights 1.
// gam is cell-major [Cp][gamma_stride(M)], columns M.. zero (one TMA bulk copy per pipeline stage).
inline int gamma_stride(int M) { return M <= 8 ? 8 : (M <= 24 ? 24 : 40); }    // == 8 (mod 16): conflict-free fragments
int launch_mstats(const mnemcu_ctx *cx, const double *gam, int M, double *partial, double *R, cudaStream_t st);
size_t mstats_partial_elems(const mnemcu_ctx *cx, int M);
// Rho path only (ctx->multi): per-group sums Rg[m][g][e] from the chunk partials of the last launch_mstats, and the
// theta-free argmax over them (null column last)
int launch_mstats_groups(const mnemcu_ctx *cx, const double *partial, int M, double *Rg, cudaStream_t st);
int launch_theta_groups(const mnemcu_ctx *cx, const double *Rg, int M, const uint32_t *clo_cols, int32_t *theta,
                        cudaStream_t st);
// host layout k x C (column-major, original cell order)  <->  internal [k][Cp] (sorted, padded with 0)
int launch_L_to_internal(const mnemcu_ctx *cx, const double *L_host_dev, int k, double *L_int, cudaStream_t st);
int launch_L_from_internal(const mnemcu_ctx *cx, const double *L_int, int k, double *L_host_dev, cudaStream_t st);

// ---- mixture.cu: responsibilities, mixture weights, likelihood on [k][Cp] ---------------------------------
constexpr int kMaxSlots = 32;
struct MixSlots {             // by-value kernel argument: one entry per mixture in the launch
    const double *L[kMaxSlots];     // [k][Cp]
    double *gam[kMaxSlots];         // output of the affinity kernel: first column of the mixture in [Cp][gw]
    const double *mw[kMaxSlots];    // [k] device
    const int32_t *anypos[kMaxSlots];   // device flag: any(L > 0)
    double *ll_out[kMaxSlots];      // device scalar receiving ll
    double *mw_out[kMaxSlots];      // [k] device receiving normalise(rowSums gamma)
    int na_guard[kMaxSlots];        // NA -> 1/k (R/mnems_low.r:595, R/mnems.r:1940)
    int n;
};
constexpr int kMixBlocks = 296;    // blocks per mixture in the statistics kernel (2 per SM)
// gamma = getAffinity(L, mw) for every slot (pad cells get 0)
int launch_affinity(const mnemcu_ctx *cx, const MixSlots &s, int k, int gw, int hard, int complete, double logtype,
                    cudaStream_t st);
// ll = getLL(L, mw), mw_out = normalise(rowSums(getAffinity(L, mw))) for every slot; scratch: n*kMixBlocks*(k+1)
int launch_mix_stats(const mnemcu_ctx *cx, const MixSlots &s, int k, int hard, int complete, double logtype, double *scratch,
                     cudaStream_t st);
// any(L > 0) over the valid cells of [k][Cp] into *flag (flag must be zeroed by the caller)
int launch_anypos(const mnemcu_ctx *cx, const double *L, int k, int32_t *flag, cudaStream_t st);

// ---- search.cu: batched greedy structure search -----------------------------------------------------------
struct SearchBatch;   // opaque workspace
int search_create(int nmax, int E, int S, SearchBatch **out);
void search_destroy(SearchBatch *sb);
// n searches; R_dev[q] points at an E x S column-major device matrix; start rows given as [n][32] row masks
// (diag is forced to 1).  Results stay on the device in the workspace until fetched.
int search_run(SearchBatch *sb, int n, const double *const *R_dev_ptrs_host, const uint32_t *start_rows_dev,
               cudaStream_t st);
// device views of the results of the last search_run
const uint32_t *search_adj_rows(const SearchBatch *sb);     // [n][32] transitively reduced, diag 0
const uint32_t *search_closed_rows(const SearchBatch *sb);  // [n][32] closure of the final graph (diag 1)
const int32_t *search_theta(const SearchBatch *sb);         // [n][E]
const double *search_score(const SearchBatch *sb);          // [n] nem$score
const double *search_maxtrace(const SearchBatch *sb);       // [n] max(nem$scores)
const int32_t *search_niter(const SearchBatch *sb);         // [n]
const unsigned long long *search_nscored(const SearchBatch *sb);   // [n]
// fraction of their resident time the CTAs of the device-resident greedy loop spent on tasks, over all runs of sb
double search_sm_busy(const SearchBatch *sb);
// decision tape (audit runs): rows of the graph after every accepted move, [q][search_hard_cap][32]
int search_enable_tape(SearchBatch *sb);
const uint32_t *search_tape(const SearchBatch *sb);
int search_hard_cap(const SearchBatch *sb);
// theta[q][e] = first argmax_j (R_q %*% cols_q)[e, j] with the null column FIRST (null_last = 0, R/mnems.r:447-450)
// or LAST (null_last = 1, R/mnems_low.r:617-619); W_out (n*E*S) optional.  cols: [n][32] column masks.
int launch_theta(const double *const *R_dev_ptrs_dev, int n, int E, int S, const uint32_t *cols, int null_last,
                 int32_t *theta, double *W_out, cudaStream_t st);
// rows [n][32] -> closure rows / closure columns
int launch_closure(const uint32_t *rows_in, int n, int S, uint32_t *clo_rows, uint32_t *clo_cols, cudaStream_t st);
// uint8 graphs <-> row masks
int launch_rows_from_u8(const uint8_t *g, int n, int S, uint32_t *rows, cudaStream_t st);
int launch_rows_to_u8(const uint32_t *rows, int n, int S, uint8_t *g, cudaStream_t st);

}  // namespace mn
