// Batched greedy structure search: nem(search = "greedy") of the reference, R/mnems.r:111-142, 203-307, 362-379,
// for many (start, component, restart) searches at once.
//
// The searches of a call are dealt to (by default two) CHAINS that iterate independently on their own streams; one greedy
// iteration of a chain is three launches over the chain's searches:
//   finish_gen_kernel   [finish of iteration i-1] which.max (first), accept rule R/mnems.r:279-289, graph update;
//                       [generation of iteration i] c(get.ins.fast, get.rev.tc, get.del.tc)(better),
//                       R/mnems_low.r:1025-1084, as bit-matrix algebra (one CTA of 1024 threads per search; a candidate is
//                       32 column masks; the one-step closure of src/mm.cpp:50-65 in column form is
//                       col'_j = col_j | (x[v,j] ? col_u : 0)), plus the per-iteration dictionary of distinct column masks
//   score_dict_kernel   scoreAdj(D, model, P, oldadj, trans.close = FALSE), R/mnems.r:205-217, 439-454, for every candidate:
//                       each distinct mask is summed once per E-gene row (ascending S-gene order, the order of the Eigen
//                       product it replaces), candidates take row maxima over table look-ups, rows are added in a fixed
//                       shape (exact ties between candidates stay exact)
//   reduce_tiles_kernel adds the 32-row tiles of every candidate in ascending order
// score_kernel is the generic scorer (any graph, every changed column re-summed; marginal = TRUE variant) behind
// mnemcu_score_adj and the marginal search; gen_kernel is the unfused generator used for the initial score and by
// mnemcu_neighbours.  Graphs are 32-bit row/column masks (S <= 32), so closure, reduction and neighbour generation
// are shuffles, ballots and popcounts.
#include <climits>

#include "common.cuh"

namespace mn {

constexpr int kCT = 32;        // candidates per score CTA (generic scorer)
constexpr int kRows = 128;     // E-gene rows per score CTA (= threads, generic scorer)

constexpr int kDictHash = 2048;   // hash slots of the per-search mask dictionary (power of two)
constexpr int kDictRows = 32;     // E-gene rows per tile of the dictionary scoring kernel
constexpr int kDictThreads = 512;
constexpr int kChains = 4;         // most independent iteration chains of the host-driven loop (see search_run_host)
constexpr int kUP = 8;             // insertion sources per work item of the dictionary scorer
constexpr int kMaxChunk = 8;       // most row tiles per task of the device-resident loop
constexpr unsigned long long kStallNs = 5000000000ull;   // watchdog of the device-resident loop: 5 s without a task

// Device view of a search workspace: one by-value kernel argument instead of twenty pointers.
struct SearchDev {
    int E = 0, S = 0, maxc = 0, dcap = 0, ntiles = 0, nsm = 0, hard_cap = 0;
    const double *const *Rptr = nullptr;
    uint32_t *Prow = nullptr, *Pcol = nullptr, *cand_cols = nullptr, *cand_J = nullptr, *dict_masks = nullptr, *rowm = nullptr;
    uint16_t *cand_id = nullptr, *base_id = nullptr, *pair_id = nullptr;
    uint8_t *chain_v = nullptr, *chain_start = nullptr;      // [q][32] targets chain after chain (bottom-up), [q][33] offsets
    int32_t *ncand = nullptr, *active = nullptr, *niter = nullptr, *nd = nullptr, *nins = nullptr, *overflow = nullptr,
            *nchain = nullptr;
    int8_t *cand_kind = nullptr;
    double *partial = nullptr, *scores = nullptr, *oldscore = nullptr, *maxtrace = nullptr;
    unsigned long long *nscored = nullptr;
    // the device-resident loop: a ring of tasks (search, first tile, tiles) and its counters
    unsigned long long *ring = nullptr;
    unsigned ring_mask = 0;
    unsigned *qctl = nullptr;         // [0] head (next ticket), [1] reserved, [2] committed, [3] done flag
    int32_t *live = nullptr;          // searches still climbing
    int32_t *tiles_done = nullptr;    // [q] tiles of the current iteration already scored
    int32_t *stage = nullptr;         // [q] 0: the pending scores are the initial score of closure(start); 1: neighbours
    unsigned long long *busy_ns = nullptr;   // [0] ns the CTAs spent on tasks, [1] ns they were resident (utilisation of the loop),
                                             // [2] ns in finish steps, [3] ns in generation + publishing, [4] iterations closed
    uint32_t *tape = nullptr;         // decision tape (audit runs only): [q][hard_cap][32] rows of the graph after every accepted move
};

struct SearchBatch {
    int nmax = 0, E = 0, S = 0, etiles = 0, maxc = 0;
    int dcap = 0;                     // dictionary capacity (entries)
    int marginal = 0;                 // scoreAdj(marginal = TRUE): log-sum-exp over the columns instead of the row maximum
    int generic = 0;                  // fallback after a dictionary overflow: generic scorer, no dictionary (search_run)
    int32_t last_flags = 0;           // overflow flags of the last loop (1: dictionary full, 2: move limit)
    double logtype = 2.0;
    size_t dict_smem = 0;
    int sm_count = 148;
    DevBuf<const double *> Rptr;
    DevBuf<uint32_t> Prow, Pcol, cand_cols, cand_J, adj_rows, closed_rows, closed_cols, dict_masks, rowm;
    DevBuf<uint16_t> cand_id, base_id, pair_id;
    DevBuf<uint8_t> chain_v, chain_start;
    DevBuf<int32_t> ncand, active, niter, n_active, theta, nd, nins, overflow, nchain, live, tiles_done, stage;
    DevBuf<double> partial, scores, oldscore, maxtrace;
    DevBuf<unsigned long long> nscored, ring, busy_ns;
    DevBuf<uint32_t> tape;            // see SearchDev::tape; allocated by search_enable_tape
    unsigned long long *h_busy = nullptr;     // pinned copy of busy_ns
    double busy_acc = 0.0, resident_acc = 0.0;   // accumulated over the search_run calls of this workspace
    double fin_acc = 0.0, gen_acc = 0.0, closed_acc = 0.0, calls_acc = 0.0, ph_acc[16] = {};
    DevBuf<unsigned> qctl;
    unsigned ring_cap = 0;
    // Host-driven loop only.  Chains: the searches are dealt round-robin (by pair) to kChains groups that iterate
    // independently on their own streams, so that one chain's single-CTA-per-search phases (argmax + candidate
    // generation), launch gaps and scoring tails are filled by the scoring kernels of the others.
    int32_t *h_n_active = nullptr;   // pinned: [2*h + (it & 1)] moves accepted by chain h in iteration it, [2*kChains] overflow
    DevBuf<int32_t> qlist;            // [nmax] search ids, chain by chain
    cudaStream_t cst[kChains] = {};   // streams of chains 1.. (chain 0 runs on the caller's stream)
    cudaEvent_t ev[kChains][2] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kChains] = {};

    SearchDev dev() const
    {
        SearchDev d;
        d.E = E; d.S = S; d.maxc = maxc; d.dcap = dcap; d.ntiles = (E + kDictRows - 1) / kDictRows; d.nsm = sm_count;
        d.hard_cap = 4 * S * S + 16;      // a strict hill-climb cannot take more moves than this in practice
        d.Rptr = Rptr.p; d.Prow = Prow.p; d.Pcol = Pcol.p; d.cand_cols = cand_cols.p; d.cand_J = cand_J.p;
        d.dict_masks = (marginal || generic) ? nullptr : dict_masks.p;      // no dictionary for the generic scorer
        d.rowm = rowm.p; d.cand_id = cand_id.p; d.base_id = base_id.p; d.pair_id = pair_id.p;
        d.chain_v = chain_v.p; d.chain_start = chain_start.p; d.ncand = ncand.p; d.active = active.p; d.niter = niter.p;
        d.nd = nd.p; d.nins = nins.p; d.overflow = overflow.p; d.nchain = nchain.p; d.partial = partial.p;
        d.scores = scores.p; d.oldscore = oldscore.p; d.maxtrace = maxtrace.p; d.nscored = nscored.p; d.ring = ring.p;
        d.ring_mask = ring_cap ? ring_cap - 1 : 0; d.qctl = qctl.p; d.live = live.p; d.tiles_done = tiles_done.p;
        d.stage = stage.p; d.busy_ns = busy_ns.p; d.tape = tape.p;
        return d;
    }
};

// Loads of per-search state that ANOTHER CTA may have written earlier in the same launch (the device-resident loop) must
// not be served from this SM's L1: CG = true reads through L2 (ld.global.cg).  The host-driven kernels use plain loads.
template <bool CG, typename T>
__device__ __forceinline__ T ldv(const T *p)
{
    if constexpr (CG) return __ldcg(p);
    else return *p;
}

// ---- candidate generation + mask dictionary ----------------------------------------------------------------
// mode 0: the single "candidate" closure(better) used for the initial score (R/mnems.r:135-141)
// mode 1: all single-edge neighbours of `better`
//
// Dictionary.  Every candidate differs from the current graph in a few columns, and the same new column mask
// recurs in many candidates (an insertion u -> v turns column j into col_j | col_u for every j below v, whatever v
// is).  While the candidates are generated, each changed column mask is entered into a shared-memory hash set;
// afterwards the occupied slots get dense ids and every candidate keeps, per changed column, the id of its mask.
// The scoring kernel then sums each distinct mask once per E-gene row instead of once per candidate.  Ids follow
// hash order; they only index a table, so no result depends on them.
// For insertions the "changed" set written to cand_J is the whole row of v (the columns the one-step closure may
// touch): candidates that share v then share the set, and the scoring kernel reuses their unchanged-column maximum.
//
// Chains (mode 1, dictionary only).  Inserting u -> v needs, for every u, the maximum over the columns j in D_v (row v
// of the graph) of the table entry of col_j | col_u.  Whenever D_c is a subset of D_v (always for a descendant c of v in
// a closed graph) that maximum is max(the maximum over D_c, the look-ups over D_v \ D_c), so the scorer walks chains
// c_1, c_2, ... of targets with nested rows bottom-up and pays |D_top| look-up rounds per chain instead of sum |D_c|.
// The greedy cover built here (largest uncovered row first, then repeatedly its largest uncovered subset) tests the
// nesting directly, so it needs no closure assumption (intermediate graphs need not be closed, R/mnems.r:128-134).
struct GenSmem {
    uint32_t *keys;     // [kDictHash]
    uint16_t *ids;      // [kDictHash]
    uint16_t *pair;     // [32 * 32]
    int *warp;          // [NT / 32]
    // optional staging (device-resident loop: the CTA has the whole scoring buffer to itself while it generates): hash
    // slots and changed-column sets of the reversal / deletion candidates, so that turning slots into dense ids needs no
    // read-back of what this CTA has just written to global memory (two dependent L2 round trips per entry otherwise)
    uint16_t *st_ci = nullptr;   // [rest][32]
    uint32_t *st_J = nullptr;    // [rest]
};
constexpr size_t kGenSmemBytes = kDictHash * 4 + kDictHash * 2 + 1024 * 2 + 64 * 4;

__device__ __forceinline__ GenSmem gen_smem_at(unsigned char *base)
{
    GenSmem g;
    g.keys = reinterpret_cast<uint32_t *>(base);
    g.ids = reinterpret_cast<uint16_t *>(base + kDictHash * 4);
    g.pair = reinterpret_cast<uint16_t *>(base + kDictHash * 6);
    g.warp = reinterpret_cast<int *>(base + kDictHash * 6 + 2048);
    return g;
}

// diagnostic phase clock of the closing path (device-resident loop): thread 0 adds the time since the last mark to ph[i];
// ph[15] holds the last time stamp.  ph == nullptr: nothing is measured.
__device__ __forceinline__ void ph_mark(unsigned long long *ph, int i)
{
    if (ph && threadIdx.x == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        ph[i] += now - ph[15];
        ph[15] = now;
    }
}

template <int NT, bool CG>
__device__ __forceinline__ void gen_body(const SearchDev &d, const GenSmem sm, int q, int mode, unsigned long long *ph = nullptr)
{
    uint32_t *keys = sm.keys;
    uint16_t *ids = sm.ids, *s_pair = sm.pair;
    int *s_warp = sm.warp;
    constexpr int kPer = kDictHash / NT;          // hash slots per thread in the dense-id scan
    static_assert(kPer * NT == kDictHash, "thread count must divide the hash size");
    const int S = d.S, maxc = d.maxc, dcap = d.dcap;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
    const bool dict = d.dict_masks != nullptr;
    if (!ldv<CG>(d.active + q)) {
        if (tid == 0) { d.ncand[q] = 0; if (dict) { d.nd[q] = 0; d.nins[q] = 0; d.nchain[q] = 0; } }
        return;
    }
    if (dict)
        for (int i = tid; i < kDictHash; i += NT) keys[i] = 0u;
    __syncthreads();
    auto insert = [&](uint32_t m) -> int {          // returns the hash slot of m (a table that is full reports overflow)
        int h = (int)((m * 2654435761u) >> 21);
        for (int probe = 0; probe < kDictHash; ++probe) {
            const uint32_t prev = atomicCAS(&keys[h], 0u, m);
            if (prev == 0u || prev == m) return h;
            h = (h + 1) & (kDictHash - 1);
        }
        atomicOr(d.overflow, 1);
        return 0;
    };
    const uint32_t smask = S == 32 ? 0xffffffffu : ((1u << S) - 1u);
    const uint32_t row = lane < S ? (ldv<CG>(d.Prow + q * 32 + lane) | (1u << lane)) & smask : 0u;
    const uint32_t col = warp_transpose(row, lane, S);
    uint32_t *cc = d.cand_cols + (size_t)q * maxc * 32;
    uint32_t *cj = d.cand_J + (size_t)q * maxc;
    uint16_t *ci = dict ? d.cand_id + (size_t)q * maxc * 32 : nullptr;
    int total = 0;
    uint16_t base_slot = 0;                                  // warp 0, lane j: hash slot of the current column j
    if (warp == 0) {
        d.Pcol[q * 32 + lane] = col;
        if (dict && lane < S) base_slot = (uint16_t)insert(col);
    }
    const bool staged = dict && sm.st_ci != nullptr && mode != 0;
    int nins_all = 0;                                        // number of insertions (mode 1), known to every thread
    ph_mark(ph, 4);                                          // hash reset, current columns
    if (mode == 0) {
        if (warp == 0) {
            const uint32_t clo = warp_closure(row, S);
            const uint32_t ccol = warp_transpose(clo, lane, S);
            cc[lane] = ccol;
            if (dict && lane < S) ci[lane] = (uint16_t)insert(ccol);
            if (lane == 0) { cj[0] = smask; d.ncand[q] = 1; if (dict) { d.nins[q] = 0; d.nchain[q] = 0; } }
        }
        total = 1;
    } else {
        // insertion u -> v turns column j (for every j with x[v,j] = 1) into col_j | col_u: enter all pairs
        if (dict)
            for (int u = warp; u < S; u += nwarp) {
                const uint32_t cu = __shfl_sync(kFull, col, u);
                if (lane < S) s_pair[u * 32 + lane] = (uint16_t)insert(col | cu);
            }
        ph_mark(ph, 5);                                              // pair unions entered
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
        total = base;
        nins_all = __shfl_sync(kFull, off[1], 0);                            // lane 0: off[1] = number of insertions
        if (tid == 0) { d.ncand[q] = base; if (dict) d.nins[q] = off[1]; }
        ph_mark(ph, 11);                                             // (warp 0) reduction, kinds, offsets
        if (dict && warp == nwarp - 1) {
            // chains of nested insertion targets for the scorer (see above); lane v: row v = D_v, km[0] = sources into v
            d.rowm[q * 32 + lane] = row;
            const int cost = __popc(row);
            unsigned unclaimed = __ballot_sync(kFull, km[0] != 0u);       // targets that have candidates at all
            int nch = 0, npos = 0;
            uint8_t *cv = d.chain_v + q * 32, *cs = d.chain_start + q * 33;
            if (lane == 0) cs[0] = 0;
            while (unclaimed) {
                unsigned long long list = 0ull;                           // the chain top-down, 5 bits per node
                int len = 0;
                uint32_t cur = 0xffffffffu;                               // row of the node above (none yet: everything nests)
                while (len < 12) {
                    // largest uncovered target whose row nests inside `cur`; ties -> lowest index
                    const bool ok = ((unclaimed >> lane) & 1u) && (row & ~cur) == 0u;
                    int key = ok ? (cost << 5) | (31 - lane) : -1;
                    for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(kFull, key, o));
                    if (key < 0) break;
                    const int v = 31 - (key & 31);
                    list |= (unsigned long long)v << (5 * len);
                    ++len;
                    unclaimed &= ~(1u << v);
                    cur = __shfl_sync(kFull, row, v);
                }
                if (lane < len) cv[npos + lane] = (uint8_t)((list >> (5 * (len - 1 - lane))) & 31ull);    // bottom-up
                npos += len;
                ++nch;
                if (lane == 0) cs[nch] = (uint8_t)npos;
            }
            if (lane == 0) d.nchain[q] = nch;
        }
        // Device-resident loop (staged): the insertion candidates are scored from the pair table and the winner's columns
        // are rebuilt from (u, v) by the finish step, so nothing is written per insertion -- only reversals and deletions
        // are materialised (most of the ~S^2 candidates of an iteration are insertions; this loop was 14 us of the 36 us a
        // CTA spends between two iterations of a search).
        // (the last warp builds the chains meanwhile: in the staged loop it takes no candidates, the others share them)
        const int nwork = (staged && nwarp > 1) ? nwarp - 1 : nwarp;
        for (int item = warp + (staged ? S : 0); item < 3 * S && warp < nwork; item += nwork) {
            const int kd = item / S, v = item - kd * S;
            uint32_t m = __shfl_sync(kFull, kd == 0 ? km[0] : (kd == 1 ? km[1] : km[2]), v);
            int idx = __shfl_sync(kFull, kd == 0 ? off[0] : (kd == 1 ? off[1] : off[2]), v);
            const uint32_t rowv = __shfl_sync(kFull, row, v);
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
                unsigned J = __ballot_sync(kFull, nc != col);
                if (kd == 0) J = rowv;                     // superset shared by every insertion into v (diag included)
                cc[(size_t)idx * 32 + lane] = nc;
                // insertions are scored from the pair table; only reversals / deletions need per-candidate ids
                if (dict && kd != 0 && ((J >> lane) & 1u)) {
                    const uint16_t slot = (uint16_t)insert(nc);
                    if (staged) sm.st_ci[(size_t)(idx - nins_all) * 32 + lane] = slot;
                    else ci[(size_t)idx * 32 + lane] = slot;
                }
                if (lane == 0) {
                    if (staged && kd != 0) sm.st_J[idx - nins_all] = J;
                    cj[idx] = J;
                    if (d.cand_kind) d.cand_kind[(size_t)q * maxc + idx] = (int8_t)kd;
                }
                ++idx;
            }
        }
    }
    ph_mark(ph, 12);                                         // (warp 0) its share of the candidates
    if (!dict) return;
    __syncthreads();
    ph_mark(ph, 6);                                          // waiting for the slowest warp (chains + candidates)
    // dense ids: exclusive scan of the occupied slots (kPer consecutive slots per thread)
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) cnt += keys[tid * kPer + i] != 0u;
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int base = incl - cnt;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const uint32_t k = keys[tid * kPer + i];
        if (k != 0u) {
            ids[tid * kPer + i] = (uint16_t)min(base, dcap - 1);   // overflow: stay inside the table (the host reports it)
            if (base < dcap) d.dict_masks[(size_t)q * dcap + base] = k;
            ++base;
        }
    }
    if (tid == NT - 1) {
        d.nd[q] = base;
        if (base > dcap) atomicOr(d.overflow, 1);
    }
    __syncthreads();
    ph_mark(ph, 7);                                          // dense ids
    // hash slots -> dense ids
    if (tid < S) d.base_id[q * 32 + tid] = ids[base_slot];         // warp 0 holds the slots of the current columns
    // insertions carry no per-candidate ids: convert from the first reversal on (mode 0: the single candidate)
    const int nins_q = mode == 0 ? 0 : nins_all;
    if (staged) {
        for (int i = tid; i < (total - nins_q) * 32; i += NT) {
            const int c = i >> 5, j = i & 31;
            if ((sm.st_J[c] >> j) & 1u) ci[(size_t)nins_q * 32 + i] = ids[sm.st_ci[i]];
        }
    } else {
        // (cand_id written by this CTA above; __syncthreads makes it visible)
        for (int i = nins_q * 32 + tid; i < total * 32; i += NT) {
            const int c = i >> 5, j = i & 31;
            if ((ldv<CG>(cj + c) >> j) & 1u) ci[i] = ids[ldv<CG>(ci + i)];
        }
    }
    if (mode != 0)
        for (int i = tid; i < 32 * 32; i += NT)
            d.pair_id[(size_t)q * 1024 + i] = ((i >> 5) < S && (i & 31) < S) ? ids[s_pair[i]] : (uint16_t)0;
    ph_mark(ph, 8);                                          // ids of the candidates and of the pair table
}

__global__ void __launch_bounds__(256) gen_kernel(SearchDev d, int mode)
{
    __shared__ __align__(16) unsigned char s_gen[kGenSmemBytes];
    gen_body<256, false>(d, gen_smem_at(s_gen), blockIdx.x, mode);
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

// MARG: scoreAdj(marginal = TRUE), R/mnems.r:455-457: a row contributes log_b(sum_j b^cbind(0, score)[e, j]) instead of
// its maximum.  The powers of the unchanged columns are cached per row; the row sum runs over the null column and then
// the S-gene columns in ascending order.
__device__ __forceinline__ double pow_logtype(double b, double y) { return b == 2.0 ? exp2(y) : pow(b, y); }

template <int SP, bool MARG>
__global__ void __launch_bounds__(kRows) score_kernel(int E, int S, int maxc, int etiles,
                                                      const double *const *__restrict__ Rptr,
                                                      const int32_t *__restrict__ ncand,
                                                      const uint32_t *__restrict__ Pcol,
                                                      const uint32_t *__restrict__ cand_cols,
                                                      const uint32_t *__restrict__ cand_J,
                                                      double *__restrict__ partial, double logtype)
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
            s_W[j][tid] = MARG ? pow_logtype(logtype, w) : w;
            if (w > m1) { m2 = m1; a2 = a1; m1 = w; a1 = j; }
            else if (w > m2) { m2 = w; a2 = j; }
        }
    }
    const double inv_lb = MARG ? 1.0 / log(logtype) : 0.0;
    for (int ci = 0; ci < nloc; ++ci) {
        uint32_t J = s_J[ci];
        if (MARG) {
            double rs = pow_logtype(logtype, 0.0);
            for (int j = 0; j < S; ++j)
                rs += ((J >> j) & 1u) ? pow_logtype(logtype, col_sum<SP>(r, s_ccol[ci][j])) : s_W[j][tid];
            double val = e < E ? log(rs) * inv_lb : 0.0;
            for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(kFull, val, o);
            if (lane == 0) s_red[ci][warp] = val;
            continue;
        }
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


// ---- dictionary scoring ----------------------------------------------------------------------------------
// scoreAdj for every candidate of one search over a set of 32-row tiles.  block = 16 warps, lane = E-gene row of the tile.
// Per tile:
// (1) table T[entry][row] = sum of the R columns in the entry's mask, ascending S-gene order (the order of the Eigen
//     product in src/mm.cpp:10); each warp takes every 16th entry, four at a time.
// (2) floors: for every insertion target v the maximum A_v over the columns OUTSIDE D_v (row v of the graph) of the
//     current subweights, floored at 0, one target per warp; one more warp keeps the four largest current subweights of
//     each row for the reversal / deletion candidates (whose changed sets are arbitrary).
// (3) work items, handed out by a shared counter so that the 16 warps finish together:
//     insertions: one item = one chain of nested targets (see gen_body) x 8 sources u.  The item walks the chain
//     bottom-up; for every NEW column j of D_v the 8 look-ups T[col_j | col_u] update the running maxima B[u] held in
//     registers; at every target the values max(A_v, B[u]) of the 32 rows are added by a butterfly that halves the
//     number of candidates per lane in its first three stages (lanes 16..31 hold the sources in swapped halves, so the
//     first stage needs no selects).  The item waits for its chain's floors only when it reaches the first target.
//     reversals / deletions: two candidates per item, worked on together; changed columns by id, unchanged-column
//     maximum from the four largest.
// Per-candidate sums of every tile go to partial[q][tile][c]; tiles are added in ascending order later.  Every row
// reduction pairs the rows as (l, l^16), (l, l^8), ...: the same function of the 32 row values, so equal rows give
// bit-equal sums whatever the candidate kind (exact score ties decide graphs, R/mnems.r:236, 279-283).
struct ScoreSmem {
    double *T;          // [dcap][32]
    uint32_t *dmask;    // [dcap]
    uint32_t *pid;      // [32][32]: pid[j*32 + u] = byte offset in T of the row of col_j | col_u
    double *A0;         // [32][32]: A0[v*32 + row] = max(0, max over the columns outside D_v)
    double *tv;         // [4][32] the four largest current subweights of each row ...
    int *ti;            // [4][32] ... and their columns (32 = none)
};
__host__ __device__ inline size_t score_smem_bytes(int dcap)
{
    return (size_t)dcap * kDictRows * 8 + (((size_t)dcap * 4 + 15) & ~(size_t)15) + 4096 + 32 * 32 * 8 + 4 * 32 * 8 + 4 * 32 * 4;
}
__device__ __forceinline__ ScoreSmem score_smem_at(unsigned char *base, int dcap)
{
    ScoreSmem s;
    s.T = reinterpret_cast<double *>(base);
    s.dmask = reinterpret_cast<uint32_t *>(s.T + (size_t)dcap * kDictRows);
    s.pid = s.dmask + ((dcap + 3) & ~3);
    s.A0 = reinterpret_cast<double *>(s.pid + 1024);
    s.tv = s.A0 + 32 * 32;
    s.ti = reinterpret_cast<int *>(s.tv + 4 * 32);
    return s;
}

template <int SP, bool CG>
__device__ __forceinline__ void score_tiles(const SearchDev &d, unsigned char *s_raw, int q, int tile_begin, int tile_end,
                                            int tile_step)
{
    __shared__ int s_vlist[32];          // chain nodes, chain after chain, each chain bottom-up
    __shared__ int s_cstart[33];         // chain h = s_vlist[s_cstart[h] .. s_cstart[h + 1])
    __shared__ uint32_t s_rowm[32];      // D_v
    __shared__ int s_ctr[2];
    __shared__ int s_aflag[33];          // [h] floors of chain h, [32] top four: == gen once written for the current tile
    const int E = d.E, S = d.S, maxc = d.maxc, dcap = d.dcap;
    const int nc = ldv<CG>(d.ncand + q);
    if (nc == 0) return;
    const int ni = ldv<CG>(d.nins + q);
    const int ndq = min(ldv<CG>(d.nd + q), dcap);
    const ScoreSmem sm = score_smem_at(s_raw, dcap);
    double *T = sm.T;
    uint32_t *dmask = sm.dmask, *pid = sm.pid;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kDictThreads / 32;
    constexpr int kRestChunk = 2;
    constexpr int kParts = (SP + kUP - 1) / kUP;    // work items per chain
    static_assert(kUP == 8, "the row reduction below halves 8 -> 4 -> 2 -> 1 over the lane bits 16, 8, 4");
    for (int i = tid; i < ndq; i += kDictThreads) dmask[i] = ldv<CG>(d.dict_masks + (size_t)q * dcap + i);
    const int nv = ni > 0 ? ldv<CG>(d.nchain + q) : 0;
    if (ni > 0) {
        for (int i = tid; i < 1024; i += kDictThreads) pid[i] = (uint32_t)ldv<CG>(d.pair_id + (size_t)q * 1024 + i) * (kDictRows * 8u);
        if (tid < 32) { s_vlist[tid] = ldv<CG>(d.chain_v + q * 32 + tid); s_rowm[tid] = ldv<CG>(d.rowm + q * 32 + tid); }
        if (tid < 33) s_cstart[tid] = ldv<CG>(d.chain_start + q * 33 + tid);
    }
    if (tid == 0) { s_ctr[0] = 0; s_ctr[1] = 0; }
    if (tid < 33) s_aflag[tid] = 0;
    int gen = 0;
    const uint32_t bid = lane < S ? ldv<CG>(d.base_id + q * 32 + lane) : 0u;
    const uint32_t smask = S == 32 ? 0xffffffffu : ((1u << S) - 1u);
    const uint32_t pcol = lane < S ? ldv<CG>(d.Pcol + q * 32 + lane) : 0u;     // lane = column
    const uint32_t zmask = lane < S ? (~pcol & smask) : 0u;            // rows u with P[u, lane] == 0: insertions into lane
    int offv;                                                           // first candidate index of column `lane`
    {
        const int cnt = __popc(zmask);
        int incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        offv = incl - cnt;
    }
    const double *R = d.Rptr[q];
    const uint32_t *cJ = d.cand_J + (size_t)q * maxc;
    const uint16_t *cI = d.cand_id + (size_t)q * maxc * 32;
    const int nrest = nc - ni;
    const int nchunks = (nrest + kRestChunk - 1) / kRestChunk;
    __syncthreads();
    int parity = 0;
    const int ntiles = d.ntiles;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        const int row = tile * kDictRows + lane;
        double *ptile = d.partial + ((size_t)q * ntiles + tile) * maxc;   // this tile's per-candidate sums
        {
            double r[SP];
#pragma unroll
            for (int s = 0; s < SP; ++s) r[s] = (s < S && row < E) ? R[(size_t)s * E + row] : 0.0;
            // (1) the table: four independent sums per warp in flight (the adds of one entry form a dependent chain)
            for (int ent = warp; ent < ndq; ent += 4 * NW) {
                const uint32_t m0 = dmask[ent];
                const bool h1 = ent + NW < ndq, h2 = ent + 2 * NW < ndq, h3 = ent + 3 * NW < ndq;
                const uint32_t m1 = h1 ? dmask[ent + NW] : 0u, m2 = h2 ? dmask[ent + 2 * NW] : 0u,
                               m3 = h3 ? dmask[ent + 3 * NW] : 0u;
                double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
#pragma unroll
                for (int s = 0; s < SP; ++s) {
                    if ((m0 >> s) & 1u) { MN_KEEP_BRANCH(); w0 += r[s]; }
                    if ((m1 >> s) & 1u) { MN_KEEP_BRANCH(); w1 += r[s]; }
                    if ((m2 >> s) & 1u) { MN_KEEP_BRANCH(); w2 += r[s]; }
                    if ((m3 >> s) & 1u) { MN_KEEP_BRANCH(); w3 += r[s]; }
                }
                T[ent * kDictRows + lane] = w0;
                if (h1) T[(ent + NW) * kDictRows + lane] = w1;
                if (h2) T[(ent + 2 * NW) * kDictRows + lane] = w2;
                if (h3) T[(ent + 3 * NW) * kDictRows + lane] = w3;
            }
        }
        if (tid == 0) s_ctr[parity ^ 1] = 0;                  // next tile's counter (idle until after the barrier below)
        ++gen;
        __syncthreads();
        // (2) + (3) work items: [0, nv) floors of chain h; [nv] top four of the current row (if there are reversal /
        // deletion candidates); then the insertion items (chain x 8 sources), then the other candidates in pairs.  The
        // floor items come first in the queue, so whoever needs a floor finds its producer already at work and waits on
        // a shared-memory flag instead of a block-wide barrier.
        // A warp draws the ticket of its NEXT item before it starts on the current one, so the shared-memory atomic and its
        // broadcast are off the critical path.  No deadlock: the items that others wait for (floors, top four) have the lowest
        // tickets and never wait themselves, so a reserved one is always behind an item that finishes.
        const int nA = nv + (nrest > 0 ? 1 : 0);
        const int nitems = nA + nv * kParts + nchunks;
        int ticket = 0;
        if (lane == 0) ticket = atomicAdd(&s_ctr[parity], 1);
        for (int item = __shfl_sync(kFull, ticket, 0); item < nitems; item = __shfl_sync(kFull, ticket, 0)) {
            if (lane == 0) ticket = atomicAdd(&s_ctr[parity], 1);
            if (item < nv) {
                // floors A_v = max(0, max over the columns outside D_v of the current subweights) for the targets of chain
                // `item`, top-down: the complements grow, so the running maximum carries over
                const int h = item;
                const int cbeg = s_cstart[h];
                uint32_t outside = 0u;                                 // columns already covered by the running maximum
                double sc = -INFINITY;
                for (int ci = s_cstart[h + 1] - 1; ci >= cbeg; --ci) {
                    const int v = s_vlist[ci];
                    uint32_t m = ~s_rowm[v] & smask & ~outside;
                    outside |= m;
                    while (m) {                                        // four look-ups in flight (a maximum: any order)
                        int j[4];
                        j[0] = __ffs(m) - 1;
                        m &= m - 1;
#pragma unroll
                        for (int t = 1; t < 4; ++t) { j[t] = m ? __ffs(m) - 1 : j[0]; m &= m - 1; }
                        double w[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) w[t] = T[__shfl_sync(kFull, bid, j[t]) * kDictRows + lane];
#pragma unroll
                        for (int t = 0; t < 4; ++t) sc = w[t] > sc ? w[t] : sc;
                    }
                    sm.A0[v * 32 + lane] = sc > 0.0 ? sc : 0.0;
                }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) *reinterpret_cast<volatile int *>(&s_aflag[h]) = gen;
                continue;
            }
            if (item < nA) {
                double t0 = -INFINITY, t1 = -INFINITY, t2 = -INFINITY, t3 = -INFINITY;
                int i0 = 32, i1 = 32, i2 = 32, i3 = 32;              // 32 = no column
                for (int j = 0; j < S; ++j) {
                    const int id = __shfl_sync(kFull, bid, j);
                    const double w = T[id * kDictRows + lane];
                    if (w > t0) { t3 = t2; i3 = i2; t2 = t1; i2 = i1; t1 = t0; i1 = i0; t0 = w; i0 = j; }
                    else if (w > t1) { t3 = t2; i3 = i2; t2 = t1; i2 = i1; t1 = w; i1 = j; }
                    else if (w > t2) { t3 = t2; i3 = i2; t2 = w; i2 = j; }
                    else if (w > t3) { t3 = w; i3 = j; }
                }
                sm.tv[lane] = t0; sm.tv[32 + lane] = t1; sm.tv[64 + lane] = t2; sm.tv[96 + lane] = t3;
                sm.ti[lane] = i0; sm.ti[32 + lane] = i1; sm.ti[64 + lane] = i2; sm.ti[96 + lane] = i3;
                __syncwarp();
                __threadfence_block();
                if (lane == 0) *reinterpret_cast<volatile int *>(&s_aflag[32]) = gen;
                continue;
            }
            item -= nA;
            if (item < nv * kParts) {
                // insertions into the targets of chain `item / kParts`, bottom-up, for the sources u in [u0, u0 + kUP)
                const int h = item / kParts, u0 = (item - h * kParts) * kUP;
                bool floors_ready = false;                            // waited for at the first target, after its look-ups
                // B[i] = max over the columns seen so far of T[col_j | col_u], u = u0 + (i ^ hsw): lanes 16..31 keep the upper
                // half of the sources in the lower slots, so the first stage of the row butterfly below sends slots
                // kUP/2.. and keeps slots 0..kUP/2-1 in every lane (no selects)
                double B[kUP];
#pragma unroll
                for (int i = 0; i < kUP; ++i) B[i] = -INFINITY;
                const int hsw = (lane & 16) ? kUP / 2 : 0;
                uint32_t seen = 0u;
                const int cend = s_cstart[h + 1];
                const unsigned char *Tl = reinterpret_cast<const unsigned char *>(T + lane);
                for (int ci = s_cstart[h]; ci < cend; ++ci) {
                    const int v = s_vlist[ci];
                    const uint32_t zfull = __shfl_sync(kFull, zmask, v);
                    const uint32_t zm = zfull >> u0;
                    const uint32_t Dv = s_rowm[v];                                   // columns j with P[v, j] = 1 (diag included)
                    uint32_t m = Dv & ~seen;
                    seen = Dv;
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        // ids of col_j | col_u: uniform 16-byte loads; the look-ups are unconditional (also for u that are
                        // no candidates) so that they pipeline instead of serialising behind uniform branches
                        uint32_t ofs[kUP];
#pragma unroll
                        for (int w4 = 0; w4 < kUP / 4; ++w4) {
                            const uint4 x = *reinterpret_cast<const uint4 *>(pid + j * 32 + u0 + ((w4 * 4) ^ hsw));
                            ofs[w4 * 4 + 0] = x.x; ofs[w4 * 4 + 1] = x.y; ofs[w4 * 4 + 2] = x.z; ofs[w4 * 4 + 3] = x.w;
                        }
#pragma unroll
                        for (int i = 0; i < kUP; ++i) {
                            const double t = *reinterpret_cast<const double *>(Tl + ofs[i]);
                            B[i] = t > B[i] ? t : B[i];
                        }
                    }
                    if (zm & ((1u << kUP) - 1u)) {                    // warp-uniform: any candidate among these sources?
                        if (!floors_ready) {                              // the chain's floors are in shared memory once its flag is up
                            while (*reinterpret_cast<volatile int *>(&s_aflag[h]) != gen) { }
                            __threadfence_block();
                            floors_ready = true;
                        }
                        const double A0 = sm.A0[v * 32 + lane];
                        // value of candidate u in this row: max(0, A_v, B[u]).  Rows are added pairwise over the lane bits
                        // 16, 8, 4, 2, 1 (the shape of every row reduction here); the first three stages halve the number
                        // of candidates a lane carries, after which lane l holds candidate 4 b16 + 2 b8 + b4.  B itself stays
                        // intact for the next target of the chain.
                        double V[kUP];
#pragma unroll
                        for (int i = 0; i < kUP; ++i) V[i] = B[i] > A0 ? B[i] : A0;   // non-candidates: summed, never stored
#pragma unroll
                        for (int i = 0; i < kUP / 2; ++i) V[i] += __shfl_xor_sync(kFull, V[i + kUP / 2], 16);
#pragma unroll
                        for (int dd = 8, n = kUP / 4; n >= 1; dd >>= 1, n >>= 1) {
                            const bool hi = (lane & dd) != 0;
#pragma unroll
                            for (int i = 0; i < n; ++i) {
                                const double send = hi ? V[i] : V[i + n];
                                const double recv = __shfl_xor_sync(kFull, send, dd);
                                V[i] = (hi ? V[i + n] : V[i]) + recv;
                            }
                        }
                        V[0] += __shfl_xor_sync(kFull, V[0], 2);
                        V[0] += __shfl_xor_sync(kFull, V[0], 1);
                        const int ui = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                        const int off = __shfl_sync(kFull, offv, v);
                        if ((lane & 3) == 0 && ((zm >> ui) & 1u))
                            ptile[off + __popc(zfull & ((1u << (u0 + ui)) - 1u))] = V[0];
                    }
                }
                continue;
            }
            // reversals and deletions (and the single closure "candidate" of the initial score)
            while (*reinterpret_cast<volatile int *>(&s_aflag[32]) != gen) { }
            __threadfence_block();
            const double t0 = sm.tv[lane], t1 = sm.tv[32 + lane], t2 = sm.tv[64 + lane], t3 = sm.tv[96 + lane];
            const int i0 = sm.ti[lane], i1 = sm.ti[32 + lane], i2 = sm.ti[64 + lane], i3 = sm.ti[96 + lane];
            // the maximum over the columns a candidate leaves unchanged is the first of the four largest whose column is
            // outside the changed set (full scan only if all four are inside)
            auto unchanged_max = [&](uint32_t Jset) -> double {
                const unsigned long long Jx = Jset;
                const bool in0 = (Jx >> i0) & 1ull, in1 = (Jx >> i1) & 1ull, in2 = (Jx >> i2) & 1ull, in3 = (Jx >> i3) & 1ull;
                double A = !in0 ? t0 : (!in1 ? t1 : (!in2 ? t2 : t3));
                const bool need_scan = in0 && in1 && in2 && in3;
                if (__any_sync(kFull, need_scan)) {
                    double sc = -INFINITY;
                    for (int j = 0; j < S; ++j) {
                        const int id = __shfl_sync(kFull, bid, j);
                        if (!((Jset >> j) & 1u)) { const double t = T[id * kDictRows + lane]; sc = t > sc ? t : sc; }
                    }
                    if (need_scan) A = sc;
                }
                return A;
            };
            // two candidates per item, worked on together: their look-ups and row sums are independent dependent chains,
            // so the warp has twice the work in flight (this part of a tile is latency-bound)
            static_assert(kRestChunk == 2, "the reversal / deletion items take candidates in pairs");
            const int ca = ni + (item - nv * kParts) * kRestChunk;
            const bool two = ca + 1 < nc;
            const int cb = two ? ca + 1 : ca;
            uint32_t Ja = ldv<CG>(cJ + ca), Jb = ldv<CG>(cJ + cb);
            const uint32_t ida = ldv<CG>(cI + (size_t)ca * 32 + lane), idb = ldv<CG>(cI + (size_t)cb * 32 + lane);
            double besta = unchanged_max(Ja);
            double bestb = Jb == Ja ? besta : unchanged_max(Jb);
            if (!two) Jb = 0u;
            while (Ja | Jb) {                               // up to four look-ups in flight; ids of lanes outside J are never used
                if (Ja) {
                    const int j0 = __ffs(Ja) - 1;
                    Ja &= Ja - 1;
                    const int j1 = Ja ? __ffs(Ja) - 1 : j0;
                    Ja &= Ja - 1;
                    const double t0a = T[__shfl_sync(kFull, ida, j0) * kDictRows + lane];
                    const double t1a = T[__shfl_sync(kFull, ida, j1) * kDictRows + lane];
                    besta = t0a > besta ? t0a : besta;      // plain compare: NaN never wins, like which.max
                    besta = t1a > besta ? t1a : besta;
                }
                if (Jb) {
                    const int j0 = __ffs(Jb) - 1;
                    Jb &= Jb - 1;
                    const int j1 = Jb ? __ffs(Jb) - 1 : j0;
                    Jb &= Jb - 1;
                    const double t0b = T[__shfl_sync(kFull, idb, j0) * kDictRows + lane];
                    const double t1b = T[__shfl_sync(kFull, idb, j1) * kDictRows + lane];
                    bestb = t0b > bestb ? t0b : bestb;
                    bestb = t1b > bestb ? t1b : bestb;
                }
            }
            double va = besta > 0.0 ? besta : 0.0, vb = bestb > 0.0 ? bestb : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                va += __shfl_xor_sync(kFull, va, o);
                vb += __shfl_xor_sync(kFull, vb, o);
            }
            if (lane == 0) {
                ptile[ca] = va;
                if (two) ptile[cb] = vb;
            }
        }
        parity ^= 1;
        __syncthreads();
    }
}

// host-driven loop: grid = (G row groups, searches of the chain)
template <int SP>
__global__ void __launch_bounds__(kDictThreads, 1) score_dict_kernel(SearchDev d, const int32_t *__restrict__ qlist)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int q = qlist ? qlist[blockIdx.y] : blockIdx.y;
    score_tiles<SP, false>(d, s_raw, q, blockIdx.x, d.ntiles, gridDim.x);
}

// ---- argmax, accept rule, graph update -------------------------------------------------------------------
// The fixed shape in which the row tiles of a candidate are added: Gp (a power of two <= 8, so that the lanes of a
// candidate sit in one warp) groups of TG consecutive tiles; ascending inside a group, then ascending over the groups.
__host__ __device__ __forceinline__ void tile_groups(int ntiles, int *Gp, int *TG)
{
    int g = 1;
    while (g < 8 && g * 16 < ntiles) g <<= 1;
    *Gp = g;
    *TG = (ntiles + g - 1) / g;
}

struct FinSmem {
    double *best;     // [NT]
    int *idx;         // [NT]
    int *flag;        // [1] 1: the search goes on (a move was accepted, or the initial score was taken)
};
template <int NT>
__device__ __forceinline__ FinSmem fin_smem_at(unsigned char *base)
{
    FinSmem f;
    f.best = reinterpret_cast<double *>(base);
    f.idx = reinterpret_cast<int *>(base + NT * 8);
    f.flag = f.idx + NT;
    return f;
}
__host__ __device__ constexpr size_t fin_smem_bytes(int nt) { return (size_t)nt * 12 + 16; }

// scores of candidate c: sum over `etiles` tile rows of src[(q*etiles + et)*maxc + c], ascending (etiles = 1: src = scores)
template <int NT, bool CG>
__device__ __forceinline__ void finish_body(const SearchDev &d, const FinSmem fs, int q, int mode, int etiles,
                                            const double *__restrict__ src, int32_t *n_active, unsigned long long *ph = nullptr)
{
    double *s_best = fs.best;
    int *s_idx = fs.idx;
    const int S = d.S, maxc = d.maxc;
    const int tid = threadIdx.x;
    if (tid == 0) fs.flag[0] = 0;
    if (!ldv<CG>(d.active + q)) { __syncthreads(); return; }
    const int nc = ldv<CG>(d.ncand + q);
    double best = -INFINITY;
    int bidx = INT_MAX;
    // warp 0 fetches the scalars of the accept step now: their L2 round trips hide behind the tile sums
    uint32_t pf_pcol = 0u;
    double pf_old = 0.0, pf_maxtrace = 0.0;
    unsigned long long pf_nscored = 0ull;
    int pf_niter = 0, pf_nins = 0;
    if (tid < 32 && mode != 0) {
        pf_pcol = ldv<CG>(d.Pcol + q * 32 + tid);
        pf_old = ldv<CG>(d.oldscore + q);
        pf_maxtrace = ldv<CG>(d.maxtrace + q);
        pf_nscored = ldv<CG>(d.nscored + q);
        pf_niter = ldv<CG>(d.niter + q);
        if constexpr (CG) pf_nins = ldv<CG>(d.nins + q);
    }
    // Score of a candidate = sum of its row tiles in a FIXED shape, whatever CTA produced them: the tiles form Gp groups
    // (tile_groups), each group is added in ascending order, then the group sums in ascending order (reduce_tiles_kernel
    // computes the same function).  A thread owns two neighbouring candidates and fetches a whole group (16 tiles) of both
    // with 16-byte loads before it adds: this step is on the critical path of every greedy iteration and is bound by how
    // fast one SM can pull the 63 x ~700 partial sums out of L2.
    {
        int Gp, TG;
        tile_groups(etiles, &Gp, &TG);
        const double *base = src + (size_t)q * etiles * maxc;      // rows of maxc doubles (maxc is even: 16-byte aligned pairs)
        for (int c = 2 * tid; c < nc; c += 2 * NT) {
            double s0 = 0.0, s1 = 0.0;
            for (int g = 0; g < Gp; ++g) {
                const int t0 = g * TG, t1 = min(etiles, t0 + TG);
                double p0 = 0.0, p1 = 0.0;
                for (int e0 = t0; e0 < t1; e0 += 16) {
                    double2 v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        v[i] = e0 + i < t1 ? ldv<CG>(reinterpret_cast<const double2 *>(base + (size_t)(e0 + i) * maxc + c))
                                           : make_double2(0.0, 0.0);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (e0 + i < t1) { p0 += v[i].x; p1 += v[i].y; }
                }
                s0 = g == 0 ? p0 : s0 + p0;
                s1 = g == 0 ? p1 : s1 + p1;
            }
            if (isnan(s0)) s0 = 0.0;                               // scores[is.na(scores)] <- 0   R/mnems.r:235
            if (isnan(s1)) s1 = 0.0;
            if (s0 > best) { best = s0; bidx = c; }                // ascending c: first maximum kept
            if (c + 1 < nc && s1 > best) { best = s1; bidx = c + 1; }
        }
    }
    ph_mark(ph, 1);                                          // tile sums
    // block argmax, ties -> the lower candidate index (which.max): warp shuffles, then the warp results
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(kFull, best, o);
        const int oi = __shfl_down_sync(kFull, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_idx[tid >> 5] = bidx; }
    __syncthreads();
    if (tid < 32) {
        best = tid < NT / 32 ? s_best[tid] : -INFINITY;
        bidx = tid < NT / 32 ? s_idx[tid] : INT_MAX;
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_down_sync(kFull, best, o);
            const int oi = __shfl_down_sync(kFull, bidx, o);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        if (tid == 0) { s_best[0] = best; s_idx[0] = bidx; }
        __syncwarp();
    }
    ph_mark(ph, 2);                                          // argmax
    if (tid < 32) {
        best = s_best[0];
        bidx = s_idx[0];
        if (nc == 0 || bidx == INT_MAX) {
            if (tid == 0) d.active[q] = 0;
        } else if (mode == 0) {
            if (tid == 0) { d.oldscore[q] = best; d.maxtrace[q] = best; fs.flag[0] = 1; }
        } else {
            const uint32_t pcol = pf_pcol;
            uint32_t ccol;
            bool rebuilt = false;
            if constexpr (CG) {
                // insertion candidates are not materialised in the device-resident loop: candidate bidx < nins is the
                // r-th missing edge u -> v of column v (ascending u), columns as in gen_body (transClose_Ins, src/mm.cpp:50-65)
                const int nins = pf_nins;
                if (bidx < nins) {
                    const uint32_t smask = S == 32 ? 0xffffffffu : ((1u << S) - 1u);
                    const uint32_t zmask = tid < S ? (~pcol & smask) : 0u;
                    const int cnt = __popc(zmask);
                    int incl = cnt;
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(kFull, incl, o);
                        if (tid >= o) incl += t;
                    }
                    const int offv = incl - cnt;
                    const unsigned hit = __ballot_sync(kFull, bidx >= offv && bidx < offv + cnt);
                    const int v = __ffs(hit) - 1;
                    const int r = bidx - __shfl_sync(kFull, offv, v);
                    const int u = (int)__fns(__shfl_sync(kFull, zmask, v), 0, r + 1);
                    const uint32_t cu = __shfl_sync(kFull, pcol, u);
                    const uint32_t b = pcol | (tid == v ? (1u << u) : 0u);
                    ccol = (((b >> v) & 1u) ? (b | cu) : b) & smask;
                    if (tid >= S) ccol = 0u;
                    rebuilt = true;
                }
            }
            if (!rebuilt) ccol = ldv<CG>(d.cand_cols + ((size_t)q * maxc + bidx) * 32 + tid);
            int np = __popc(pcol), ncn = __popc(ccol);
            for (int o = 16; o > 0; o >>= 1) {
                np += __shfl_down_sync(kFull, np, o);
                ncn += __shfl_down_sync(kFull, ncn, o);
            }
            np = __shfl_sync(kFull, np, 0);
            ncn = __shfl_sync(kFull, ncn, 0);
            const double old = pf_old;
            // R/mnems.r:279-283: better score, or equal score and fewer edges
            const bool accept = best > old || (best == old && np > ncn);
            const uint32_t newrow = warp_transpose(ccol, tid, S);
            if (tid == 0) d.nscored[q] = pf_nscored + (unsigned long long)nc;
            if (accept) {
                d.Prow[q * 32 + tid] = newrow;
                if (d.tape) d.tape[((size_t)q * d.hard_cap + pf_niter) * 32 + tid] = newrow;
                if (tid == 0) {
                    d.oldscore[q] = best;
                    if (best > pf_maxtrace) d.maxtrace[q] = best;
                    const int it = pf_niter + 1;
                    d.niter[q] = it;
                    if (n_active) atomicAdd(n_active, 1);
                    if (it >= d.hard_cap) { d.active[q] = 0; atomicOr(d.overflow, 2); }    // not converged: reported by the host
                    else fs.flag[0] = 1;
                }
            } else if (tid == 0) {
                d.active[q] = 0;
            }
        }
    }
    __syncthreads();       // the accepted graph (Prow), the active flag and fs.flag are visible to the whole CTA
    ph_mark(ph, 3);                                          // accept rule, graph update
}

// finish of greedy iteration i fused with the candidate generation of iteration i + 1 (same CTA per search): one
// launch less per iteration and no host round trip between the accept decision and the next neighbourhood.
constexpr int kFgThreads = 1024;   // 32 warps: the 3*S candidate groups of a search spread over all of them
__global__ void __launch_bounds__(kFgThreads) finish_gen_kernel(SearchDev d, int fmode, int32_t *n_active,
                                                                const int32_t *__restrict__ qlist)
{
    __shared__ __align__(16) unsigned char s_gen[kGenSmemBytes];
    __shared__ __align__(16) unsigned char s_fin[fin_smem_bytes(kFgThreads)];
    const int q = qlist[blockIdx.x];                 // the searches of one chain (see search_run_host)
    finish_body<kFgThreads, false>(d, fin_smem_at<kFgThreads>(s_fin), q, fmode, 1, d.scores, n_active);
    gen_body<kFgThreads, false>(d, gen_smem_at(s_gen), q, 1);
}

// ---- the device-resident greedy loop -----------------------------------------------------------------------
// One persistent CTA per SM.  A task is (search, first tile, number of tiles) of the current iteration of that search;
// CTAs draw tickets from a ring.  The CTA that scores the LAST tile of a search's iteration finishes it on the spot:
// adds the tiles (ascending), takes which.max, applies the accept rule, generates the next neighbourhood with its
// dictionary and chains, and publishes the tasks of the next iteration.  Searches therefore advance independently of
// each other (no launch, no grid-wide barrier, no host round trip per iteration), a long search never holds back the
// short ones, and the SMs stay busy as long as any search has tiles left.  Results do not depend on the schedule: every
// number is produced by the same arithmetic in the same order as in the host-driven loop (tested bit for bit).
__device__ __forceinline__ unsigned long long task_encode(int q, int tile0, int nt)
{
    return ((unsigned long long)q << 40) | ((unsigned long long)tile0 << 8) | (unsigned long long)nt;
}

// (re)publishes the tiles of search q as tasks; called by one thread
__device__ __forceinline__ void publish_search(const SearchDev &d, int q)
{
    const int live = max(1, *reinterpret_cast<volatile int32_t *>(d.live));
    int chunk = (int)(((long)d.ntiles * live + 2L * d.nsm - 1) / (2L * d.nsm));
    chunk = max(1, min(kMaxChunk, chunk));
    const int nch = (d.ntiles + chunk - 1) / chunk;
    d.tiles_done[q] = 0;
    const unsigned r = atomicAdd(&d.qctl[1], (unsigned)nch);
    for (int i = 0; i < nch; ++i)
        d.ring[(r + i) & d.ring_mask] = task_encode(q, i * chunk, min(chunk, d.ntiles - i * chunk));
    __threadfence();
    unsigned spins = 0;
    while (atomicCAS(&d.qctl[2], r, r + (unsigned)nch) != r) {                      // commit in reservation order
        __nanosleep(40);
        if (*reinterpret_cast<volatile unsigned *>(&d.qctl[3]) || ++spins > (1u << 27)) break;   // aborted run: do not wait
    }
}

__global__ void search_seed_kernel(SearchDev d, int n)
{
    // one thread: every search starts with the score of closure(start) pending (stage 0)
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    d.qctl[0] = 0u; d.qctl[1] = 0u; d.qctl[2] = 0u; d.qctl[3] = 0u;
    *d.live = n;
    for (int q = 0; q < n; ++q) d.stage[q] = 0;
    __threadfence();
    for (int q = 0; q < n; ++q) publish_search(d, q);
}

template <int SP>
__global__ void __launch_bounds__(kDictThreads, 1) search_persistent_kernel(SearchDev d)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ unsigned long long s_task;
    __shared__ int s_last;
    __shared__ unsigned long long s_ph[16];
    const int tid = threadIdx.x;
    if (tid < 16) s_ph[tid] = 0ull;
    __syncthreads();
    unsigned long long t_begin = 0ull, t_task = 0ull, busy = 0ull, t_fin = 0ull, t_gen = 0ull, n_closed = 0ull;
    if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    bool first = true;
    while (true) {
        if (tid == 0) {
            if (!first) { unsigned long long now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now)); busy += now - t_task; }
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_task));          // start of this wait (watchdog reference)
            first = false;
            const unsigned idx = atomicAdd(&d.qctl[0], 1u);
            unsigned long long task = ~0ull;
            unsigned spins = 0;
            while (true) {
                if (*reinterpret_cast<volatile unsigned *>(&d.qctl[3])) break;      // every search has ended (or the run was aborted)
                const unsigned committed = *reinterpret_cast<volatile unsigned *>(&d.qctl[2]);
                if ((int)(committed - idx) > 0) {
                    __threadfence();
                    task = __ldcg(d.ring + (idx & d.ring_mask));
                    break;
                }
                __nanosleep(100);
                if ((++spins & 4095u) == 0u) {
                    // watchdog: the longest legitimate wait is one iteration of one search (well under a millisecond); a
                    // CTA that has seen no task for kStallNs aborts the run instead of spinning on the GPU for ever
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (now - t_task > kStallNs && t_task != 0ull) {
                        atomicOr(d.overflow, 4);
                        *reinterpret_cast<volatile unsigned *>(&d.qctl[3]) = 1u;
                    }
                }
            }
            s_task = task;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_task));
        }
        __syncthreads();
        const unsigned long long task = s_task;
        if (task == ~0ull) break;
        const int q = (int)(task >> 40), tile0 = (int)((task >> 8) & 0xffffffffu), nt = (int)(task & 0xffu);
        score_tiles<SP, true>(d, s_raw, q, tile0, tile0 + nt, 1);
        __threadfence();                                   // this CTA's partial sums are out before the count moves
        __syncthreads();
        if (tid == 0) {
            const int prev = atomicAdd(&d.tiles_done[q], nt);
            s_last = prev + nt == d.ntiles;
            if (s_last) __threadfence();
        }
        __syncthreads();
        if (!s_last) continue;
        // ---- this CTA closes the iteration of search q ----
        unsigned long long tp0 = 0ull, tp1 = 0ull, tp2 = 0ull;
        if (tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tp0)); s_ph[15] = tp0; }
        const int mode = __ldcg(d.stage + q);
        finish_body<kDictThreads, true>(d, fin_smem_at<kDictThreads>(s_raw + kGenSmemBytes), q, mode, d.ntiles, d.partial,
                                        nullptr, s_ph);
        const int go = fin_smem_at<kDictThreads>(s_raw + kGenSmemBytes).flag[0];
        __syncthreads();
        if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tp1));
        if (go) {
            GenSmem gs = gen_smem_at(s_raw);
            {   // staging behind the generation and finish areas; capacity checked against the workspace in search_create
                unsigned char *stg = s_raw + ((kGenSmemBytes + fin_smem_bytes(kDictThreads) + 15) & ~(size_t)15);
                gs.st_J = reinterpret_cast<uint32_t *>(stg);
                gs.st_ci = reinterpret_cast<uint16_t *>(stg + (((size_t)d.maxc * 4 + 15) & ~(size_t)15));
            }
            gen_body<kDictThreads, true>(d, gs, q, 1, s_ph);
            __threadfence();
            __syncthreads();
            ph_mark(s_ph, 9);                                // fence
            if (tid == 0) {
                d.stage[q] = 1;
                if (*reinterpret_cast<volatile int32_t *>(d.overflow) & 1) {      // dictionary overflow: stop everything
                    *reinterpret_cast<volatile unsigned *>(&d.qctl[3]) = 1u;
                } else {
                    __threadfence();
                    publish_search(d, q);
                }
            }
        } else if (tid == 0) {
            __threadfence();
            if (atomicSub(d.live, 1) == 1) { __threadfence(); *reinterpret_cast<volatile unsigned *>(&d.qctl[3]) = 1u; }
        }
        ph_mark(s_ph, 10);                                   // publish / retire
        if (tid == 0) {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tp2));
            t_fin += tp1 - tp0; t_gen += tp2 - tp1; ++n_closed;
        }
        __syncthreads();
    }
    if (tid == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        atomicAdd(&d.busy_ns[0], busy);
        atomicAdd(&d.busy_ns[1], now - t_begin);
        atomicAdd(&d.busy_ns[2], t_fin);
        atomicAdd(&d.busy_ns[3], t_gen);
        atomicAdd(&d.busy_ns[4], n_closed);
        for (int i = 1; i <= 12; ++i) atomicAdd(&d.busy_ns[8 + i], s_ph[i]);
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
    do {                                                                                                           \
        if (sb->marginal)                                                                                          \
            MN_LAUNCH((score_kernel<SPV, true>), grid, kRows, 0, st, sb->E, S, sb->maxc, sb->etiles, sb->Rptr.p,   \
                      sb->ncand.p, sb->Pcol.p, sb->cand_cols.p, sb->cand_J.p, sb->partial.p, sb->logtype);         \
        else                                                                                                       \
            MN_LAUNCH((score_kernel<SPV, false>), grid, kRows, 0, st, sb->E, S, sb->maxc, sb->etiles, sb->Rptr.p,  \
                      sb->ncand.p, sb->Pcol.p, sb->cand_cols.p, sb->cand_J.p, sb->partial.p, sb->logtype);         \
    } while (0)
    if (S <= 8) MN_SCORE(8);
    else if (S <= 16) MN_SCORE(16);
    else if (S <= 24) MN_SCORE(24);
    else MN_SCORE(32);
#undef MN_SCORE
    return 0;
}

// scores[q][c] = sum over the row tiles in the fixed shape of tile_groups -- fully parallel over candidates, so the
// single-CTA finish step of the host-driven loop reads one value per candidate instead of one per (candidate, tile)
__global__ void __launch_bounds__(256) reduce_tiles_kernel(int maxc, int ntiles, const int32_t *__restrict__ ncand,
                                                           const double *__restrict__ partial, double *__restrict__ scores,
                                                           const int32_t *__restrict__ qlist)
{
    const int q = qlist ? qlist[blockIdx.y] : blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncand[q]) return;
    int Gp, TG;
    tile_groups(ntiles, &Gp, &TG);                                  // the same shape as finish_body
    double sc = 0.0;
    for (int g = 0; g < Gp; ++g) {
        double p = 0.0;
#pragma unroll 8
        for (int et = g * TG; et < min(ntiles, (g + 1) * TG); ++et) p += partial[((size_t)q * ntiles + et) * maxc + c];
        sc = g == 0 ? p : sc + p;
    }
    scores[(size_t)q * maxc + c] = sc;
}

static int launch_gen(SearchBatch *sb, int n, int mode, cudaStream_t st)
{
    MN_LAUNCH(gen_kernel, n, 256, 0, st, sb->dev(), mode);
    return 0;
}

// the templates are instantiated for these numbers of S-gene registers
#define MN_FOR_SP(S, M)                                                                                            \
    do {                                                                                                           \
        if ((S) <= 8) M(8);                                                                                        \
        else if ((S) <= 16) M(16);                                                                                 \
        else if ((S) <= 24) M(24);                                                                                 \
        else if ((S) <= 28) M(28);                                                                                 \
        else if ((S) <= 30) M(30);                                                                                 \
        else M(32);                                                                                                \
    } while (0)

// scores of the `n` searches listed in qlist (device; nullptr = searches 0..n-1) with G row groups each
static int launch_score_dict(SearchBatch *sb, int n, int G, const int32_t *qlist, cudaStream_t st)
{
    const int S = sb->S;
    if (sb->marginal || sb->generic) {   // the dictionary trick needs a row MAXIMUM; the marginal score re-sums the changed columns per candidate
        MN_TRY(launch_score(sb, n, st));
        dim3 rg((sb->maxc + 255) / 256, n);
        MN_LAUNCH(reduce_tiles_kernel, rg, 256, 0, st, sb->maxc, sb->etiles, sb->ncand.p, sb->partial.p, sb->scores.p,
                  (const int32_t *)nullptr);
        return 0;
    }
    dim3 grid(G, n);
    const SearchDev d = sb->dev();
#define MN_SCORE(SPV) MN_LAUNCH(score_dict_kernel<SPV>, grid, kDictThreads, sb->dict_smem, st, d, qlist)
    MN_FOR_SP(S, MN_SCORE);
#undef MN_SCORE
    dim3 rgrid((sb->maxc + 255) / 256, n);
    MN_LAUNCH(reduce_tiles_kernel, rgrid, 256, 0, st, sb->maxc, d.ntiles, sb->ncand.p, sb->partial.p, sb->scores.p, qlist);
    return 0;
}

constexpr int kGcap = 64;          // most row groups (CTAs) per search

// One CTA fits per SM (the table fills shared memory), so the launch runs in ceil(G*n/SMs) rounds of ceil(tiles/G)
// tiles each: pick the G that minimises the product for the current number of active searches.
static int pick_groups(int n_active, int ntiles, int sms)
{
    int best = 1;
    long best_cost = -1;
    for (int G = 1; G <= std::min(kGcap, ntiles); ++G) {
        const long cost = (long)((G * (long)n_active + sms - 1) / sms) * ((ntiles + G - 1) / G);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = G; }
    }
    return best;
}

int search_create(int nmax, int E, int S, SearchBatch **out)
{
    MN_ARG(nmax >= 1 && E >= 1 && S >= 1 && S <= MNEMCU_MAX_S, "bad search workspace size");
    SearchBatch *sb = new SearchBatch();
    sb->nmax = nmax; sb->E = E; sb->S = S;
    sb->etiles = (E + kRows - 1) / kRows;
    sb->maxc = std::max(2, 3 * S * (S - 1));          // even: finish_body reads the partial sums as double2
    const int ntiles = (E + kDictRows - 1) / kDictRows;
    {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            sb->sm_count = sms;
        cudaGetLastError();
    }
    {   // dictionary capacity: pairwise unions + deletion masks + slack, limited by the 227 KB of shared memory
        const int need = S * (S + 1) / 2 + S * S / 4 + 2 * S + 8;
        int fit = 1;
        while (score_smem_bytes(fit + 1) + 1536 <= 232448) ++fit;   // 1.5 KB of static shared memory in the kernels
        sb->dcap = std::max(1, std::min(need + need / 2, fit));     // 50 % slack where shared memory allows (S <= 26)
        if (const char *e = getenv("MNEMCU_DICT_CAP")) sb->dcap = std::max(1, std::min(sb->dcap, atoi(e)));   // tests: force the fallback
        // generation phase of the device-resident loop: hash set + finish scratch + the staged ids of the reversal /
        // deletion candidates (at most S(S-1) of each kind), see GenSmem
        const size_t stage_need = ((kGenSmemBytes + fin_smem_bytes(kDictThreads) + 15) & ~(size_t)15) +
                                  (((size_t)sb->maxc * 4 + 15) & ~(size_t)15) + (size_t)(2 * S * (S - 1) + S) * 64;
        sb->dict_smem = std::max(score_smem_bytes(sb->dcap), stage_need);
        // per device: set on every workspace creation (cheap), the attribute belongs to the current context
#define MN_ATTR(SPV)                                                                                               \
        cudaFuncSetAttribute(score_dict_kernel<SPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb->dict_smem); \
        cudaFuncSetAttribute(search_persistent_kernel<SPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb->dict_smem)
        MN_ATTR(8); MN_ATTR(16); MN_ATTR(24); MN_ATTR(28); MN_ATTR(30); MN_ATTR(32);
#undef MN_ATTR
    }
    sb->ring_cap = 1;
    while (sb->ring_cap < (unsigned)nmax * (unsigned)ntiles + 1u) sb->ring_cap <<= 1;
    auto body = [&]() -> int {
        MN_TRY(sb->Rptr.alloc(nmax));
        MN_TRY(sb->Prow.alloc((size_t)nmax * 32, true)); MN_TRY(sb->Pcol.alloc((size_t)nmax * 32, true));
        MN_TRY(sb->cand_cols.alloc((size_t)nmax * sb->maxc * 32)); MN_TRY(sb->cand_J.alloc((size_t)nmax * sb->maxc));
        MN_TRY(sb->adj_rows.alloc((size_t)nmax * 32)); MN_TRY(sb->closed_rows.alloc((size_t)nmax * 32));
        MN_TRY(sb->closed_cols.alloc((size_t)nmax * 32));
        MN_TRY(sb->ncand.alloc(nmax, true)); MN_TRY(sb->active.alloc(nmax, true)); MN_TRY(sb->niter.alloc(nmax, true));
        MN_TRY(sb->n_active.alloc(2 * kChains, true)); MN_TRY(sb->theta.alloc((size_t)nmax * E));
        MN_TRY(sb->qlist.alloc(nmax, true));
        MN_TRY(sb->partial.alloc((size_t)nmax * ntiles * sb->maxc));
        MN_TRY(sb->scores.alloc((size_t)nmax * sb->maxc));
        MN_TRY(sb->dict_masks.alloc((size_t)nmax * sb->dcap)); MN_TRY(sb->nd.alloc(nmax, true));
        MN_TRY(sb->cand_id.alloc((size_t)nmax * sb->maxc * 32, true)); MN_TRY(sb->base_id.alloc((size_t)nmax * 32, true));
        MN_TRY(sb->overflow.alloc(1, true)); MN_TRY(sb->pair_id.alloc((size_t)nmax * 1024, true));
        MN_TRY(sb->nins.alloc(nmax, true));
        MN_TRY(sb->rowm.alloc((size_t)nmax * 32, true)); MN_TRY(sb->chain_v.alloc((size_t)nmax * 32, true));
        MN_TRY(sb->chain_start.alloc((size_t)nmax * 33, true)); MN_TRY(sb->nchain.alloc(nmax, true));
        MN_TRY(sb->ring.alloc(sb->ring_cap)); MN_TRY(sb->qctl.alloc(4, true)); MN_TRY(sb->live.alloc(1, true));
        MN_TRY(sb->tiles_done.alloc(nmax, true)); MN_TRY(sb->stage.alloc(nmax, true)); MN_TRY(sb->busy_ns.alloc(24, true));
        MN_CU(cudaMallocHost((void **)&sb->h_busy, 24 * sizeof(unsigned long long)));
        MN_TRY(sb->oldscore.alloc(nmax)); MN_TRY(sb->maxtrace.alloc(nmax)); MN_TRY(sb->nscored.alloc(nmax, true));
        MN_CU(cudaMallocHost((void **)&sb->h_n_active, (2 * kChains + 1) * sizeof(int32_t)));
        for (int i = 0; i <= 2 * kChains; ++i) sb->h_n_active[i] = 0;
        for (int h = 0; h < kChains; ++h) {
            MN_CU(cudaEventCreateWithFlags(&sb->ev[h][0], cudaEventDisableTiming));
            MN_CU(cudaEventCreateWithFlags(&sb->ev[h][1], cudaEventDisableTiming));
            MN_CU(cudaEventCreateWithFlags(&sb->ev_join[h], cudaEventDisableTiming));
            if (h > 0) MN_CU(cudaStreamCreateWithFlags(&sb->cst[h], cudaStreamNonBlocking));
        }
        MN_CU(cudaEventCreateWithFlags(&sb->ev_fork, cudaEventDisableTiming));
        // the zero fills above ran on the legacy default stream; the searches run on non-blocking streams
        MN_CU(cudaDeviceSynchronize());
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
    if (getenv("MNEMCU_SEARCH_STATS") && sb->calls_acc > 0)        // phase times of the device-resident loop (diagnostic)
        fprintf(stderr, "search loop: %.0f launches, %.0f iterations closed, CTA time busy %.1f ms of %.1f ms resident (sum over "
                        "CTAs); per closed iteration: finish %.2f us, generate + publish %.2f us\n", sb->calls_acc, sb->closed_acc,
                sb->busy_acc * 1e-6, sb->resident_acc * 1e-6, sb->fin_acc / std::max(1.0, sb->closed_acc) * 1e-3,
                sb->gen_acc / std::max(1.0, sb->closed_acc) * 1e-3);
    if (getenv("MNEMCU_SEARCH_STATS") && sb->calls_acc > 0) {
        static const char *nm[13] = {"", "tile sums", "argmax", "accept + update", "hash reset", "pair unions",
                                     "wait for the slowest warp", "dense ids", "candidate / pair ids", "fence", "publish",
                                     "warp 0: reduction, kinds, offsets", "warp 0: its candidates"};
        for (int i = 1; i <= 12; ++i)
            fprintf(stderr, "  %-34s %6.2f us\n", nm[i], sb->ph_acc[i] / std::max(1.0, sb->closed_acc) * 1e-3);
    }
    if (sb->h_n_active) cudaFreeHost(sb->h_n_active);
    if (sb->h_busy) cudaFreeHost(sb->h_busy);
    for (int h = 0; h < kChains; ++h) {
        for (int i = 0; i < 2; ++i) if (sb->ev[h][i]) cudaEventDestroy(sb->ev[h][i]);
        if (sb->ev_join[h]) cudaEventDestroy(sb->ev_join[h]);
        if (sb->cst[h]) cudaStreamDestroy(sb->cst[h]);
    }
    if (sb->ev_fork) cudaEventDestroy(sb->ev_fork);
    delete sb;
}

static int search_check_flags(SearchBatch *sb, int32_t flags)
{
    sb->last_flags = flags;
    if (flags & 1) return set_err(MNEMCU_ERR_NOMEM, "mask dictionary overflow (more than %d distinct columns)", sb->dcap);
    if (flags & 2)
        return set_err(MNEMCU_ERR_ARG, "a greedy search did not converge within %d moves", 4 * sb->S * sb->S + 16);
    if (flags & 4) return set_err(MNEMCU_ERR_CUDA, "the device-resident search loop stalled (watchdog)");
    return 0;
}

// The greedy loops of all n searches inside ONE launch (search_persistent_kernel); the host only seeds the queue.
static int search_loop_device(SearchBatch *sb, int n, cudaStream_t st)
{
    const SearchDev d = sb->dev();
    MN_CU(cudaMemsetAsync(sb->busy_ns.p, 0, 24 * sizeof(unsigned long long), st));
    MN_LAUNCH(search_seed_kernel, 1, 32, 0, st, d, n);
    const int grid = std::max(1, std::min(sb->sm_count, n * d.ntiles));
#define MN_PERS(SPV) MN_LAUNCH(search_persistent_kernel<SPV>, grid, kDictThreads, sb->dict_smem, st, d)
    MN_FOR_SP(sb->S, MN_PERS);
#undef MN_PERS
    int32_t *h_over = sb->h_n_active + 2 * kChains;
    MN_CU(cudaMemcpyAsync(h_over, sb->overflow.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MN_CU(cudaMemcpyAsync(sb->h_busy, sb->busy_ns.p, 24 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    MN_CU(cudaStreamSynchronize(st));
    sb->busy_acc += (double)sb->h_busy[0];
    sb->resident_acc += (double)sb->h_busy[1];
    sb->fin_acc += (double)sb->h_busy[2]; sb->gen_acc += (double)sb->h_busy[3]; sb->closed_acc += (double)sb->h_busy[4];
    sb->calls_acc += 1.0;
    for (int i = 1; i <= 12; ++i) sb->ph_acc[i] += (double)sb->h_busy[8 + i];
    return search_check_flags(sb, *h_over);
}

// The same loops driven from the host, one greedy iteration of a chain = three launches.  Kept for
// scoreAdj(marginal = TRUE) (whose scorer is not dictionary-based) and as the reference schedule the device-resident
// loop is tested against bit for bit (MNEMCU_SEARCH=host).
static int search_loop_host(SearchBatch *sb, int n, cudaStream_t st)
{
    const int ntiles = (sb->E + kDictRows - 1) / kDictRows;
    const int sms = sb->sm_count;
    const SearchDev d = sb->dev();
    MN_TRY(launch_score_dict(sb, n, pick_groups(n, ntiles, sms), nullptr, st));
    // Chains.  Searches come in pairs (2p: empty start, 2p+1: previous graph, R/mnems.r:2036-2073) with very different
    // lengths, so pairs -- not single searches -- are dealt round-robin.
    int want = 2;                                      // default number of chains
    if (const char *e = getenv("MNEMCU_CHAINS")) want = std::max(1, std::min(kChains, atoi(e)));
    const int nch = (sb->marginal || sb->generic || n < 2 * want) ? 1 : want;
    struct Chain { int n = 0, off = 0, pending = -1, G = 1; bool done = false; cudaStream_t st = nullptr; };
    Chain ch[kChains];
    {
        std::vector<int32_t> ql;
        for (int h = 0; h < nch; ++h) {
            ch[h].off = (int)ql.size();
            for (int q = 0; q < n; ++q)
                if ((q >> 1) % nch == h) ql.push_back(q);
            ch[h].n = (int)ql.size() - ch[h].off;
            ch[h].st = h == 0 ? st : sb->cst[h];
            ch[h].G = pick_groups(ch[h].n, ntiles, sms);
            ch[h].done = ch[h].n == 0;
        }
        // pageable source: the call returns once the bytes are staged, so the stack vector may go out of scope
        MN_CU(cudaMemcpyAsync(sb->qlist.p, ql.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    }
    for (int h = 1; h < nch; ++h) {                   // fork: the other chains start after the list and the initial scores
        MN_CU(cudaEventRecord(sb->ev_fork, st));
        MN_CU(cudaStreamWaitEvent(ch[h].st, sb->ev_fork, 0));
    }
    // iteration it of a chain: [finish(it-1) + gen(it)] -> score(it).  The host learns the number of searches that
    // accepted a move one iteration late (the copy of iteration it is awaited while iteration it+1 is already queued), so
    // the GPU never idles on a host round trip; the one surplus iteration at the end finds every search of the chain
    // inactive and does nothing.
    int32_t *h_over = sb->h_n_active + 2 * kChains;
    auto all_done = [&]() { for (int h = 0; h < nch; ++h) if (!ch[h].done) return false; return true; };
    for (int it = 0; it < d.hard_cap + 2 && !all_done(); ++it) {
        const int slot = it & 1;
        for (int h = 0; h < nch; ++h) {
            Chain &c = ch[h];
            if (c.done) continue;
            int32_t *na = sb->n_active.p + 2 * h + slot;
            const int32_t *ql = sb->qlist.p + c.off;
            MN_CU(cudaMemsetAsync(na, 0, sizeof(int32_t), c.st));
            // *na counts the searches of the chain that accept a move in this fused finish
            MN_LAUNCH(finish_gen_kernel, c.n, kFgThreads, 0, c.st, d, it == 0 ? 0 : 1, na, ql);
            MN_TRY(launch_score_dict(sb, c.n, c.G, ql, c.st));
            MN_CU(cudaMemcpyAsync(sb->h_n_active + 2 * h + slot, na, sizeof(int32_t), cudaMemcpyDeviceToHost, c.st));
            MN_CU(cudaMemcpyAsync(h_over, sb->overflow.p, sizeof(int32_t), cudaMemcpyDeviceToHost, c.st));
            MN_CU(cudaEventRecord(sb->ev[h][slot], c.st));
            if (c.pending >= 0) {
                const int ps = c.pending & 1;
                MN_CU(cudaEventSynchronize(sb->ev[h][ps]));
                if (*h_over & 1) return search_check_flags(sb, *h_over);
                // iteration 0's finish is the initial score (mode 0): it accepts nothing by design
                const int acc = sb->h_n_active[2 * h + ps];
                if (c.pending > 0 && acc == 0) c.done = true;
                else if (c.pending > 0) c.G = pick_groups(acc, ntiles, sms);
            }
            c.pending = it;
        }
    }
    for (int h = 1; h < nch; ++h) {                   // join
        MN_CU(cudaEventRecord(sb->ev_join[h], ch[h].st));
        MN_CU(cudaStreamWaitEvent(st, sb->ev_join[h], 0));
    }
    MN_CU(cudaStreamSynchronize(st));
    return search_check_flags(sb, *h_over);
}

int search_run(SearchBatch *sb, int n, const double *const *R_ptrs_host, const uint32_t *start_rows_dev, cudaStream_t st)
{
    MN_ARG(n >= 1 && n <= sb->nmax, "too many searches for the workspace");
    const int S = sb->S;
    MN_CU(cudaMemcpyAsync(sb->Rptr.p, R_ptrs_host, sizeof(double *) * n, cudaMemcpyHostToDevice, st));
    MN_CU(cudaMemsetAsync(sb->overflow.p, 0, sizeof(int32_t), st));
    MN_LAUNCH(search_init_kernel, n, 32, 0, st, n, S, start_rows_dev, sb->Prow.p, sb->active.p, sb->niter.p, sb->nscored.p);
    // initial score of closure(better), R/mnems.r:135-141
    MN_TRY(launch_gen(sb, n, 0, st));
    bool host_loop = sb->marginal != 0;
    if (const char *e = getenv("MNEMCU_SEARCH")) host_loop = host_loop || strcmp(e, "host") == 0;
    int rc = host_loop ? search_loop_host(sb, n, st) : search_loop_device(sb, n, st);
    if (rc && (sb->last_flags & 1) && !sb->marginal) {
        // More distinct column masks in one neighbourhood than the shared-memory table holds (capacity is a heuristic for
        // S >= 27): run these searches again from their start graphs with the generic scorer, which re-sums the changed
        // columns of every candidate and needs no dictionary.  Slower, never wrong; scores may differ from the dictionary
        // path in the last bits (different row-tile shape).
        sb->generic = 1;
        auto again = [&]() -> int {
            MN_CU(cudaMemsetAsync(sb->overflow.p, 0, sizeof(int32_t), st));
            MN_LAUNCH(search_init_kernel, n, 32, 0, st, n, S, start_rows_dev, sb->Prow.p, sb->active.p, sb->niter.p, sb->nscored.p);
            MN_TRY(launch_gen(sb, n, 0, st));
            return search_loop_host(sb, n, st);
        };
        rc = again();
        sb->generic = 0;
    }
    if (rc) return rc;
    MN_LAUNCH(search_final_kernel, n, 32, 0, st, S, sb->Prow.p, sb->adj_rows.p, sb->closed_rows.p, sb->closed_cols.p);
    MN_TRY(launch_theta(sb->Rptr.p, n, sb->E, S, sb->closed_cols.p, 0, sb->theta.p, nullptr, st));
    return 0;
}

int search_enable_tape(SearchBatch *sb)
{
    if (sb->tape.p) return 0;
    MN_TRY(sb->tape.alloc((size_t)sb->nmax * (4 * sb->S * sb->S + 16) * 32));
    return 0;
}
const uint32_t *search_tape(const SearchBatch *sb) { return sb->tape.p; }
int search_hard_cap(const SearchBatch *sb) { return 4 * sb->S * sb->S + 16; }
const uint32_t *search_adj_rows(const SearchBatch *sb) { return sb->adj_rows.p; }
const uint32_t *search_closed_rows(const SearchBatch *sb) { return sb->closed_rows.p; }
const int32_t *search_theta(const SearchBatch *sb) { return sb->theta.p; }
const double *search_score(const SearchBatch *sb) { return sb->oldscore.p; }
const double *search_maxtrace(const SearchBatch *sb) { return sb->maxtrace.p; }
const int32_t *search_niter(const SearchBatch *sb) { return sb->niter.p; }
const unsigned long long *search_nscored(const SearchBatch *sb) { return sb->nscored.p; }
double search_sm_busy(const SearchBatch *sb) { return sb->resident_acc > 0 ? sb->busy_acc / sb->resident_acc : 0.0; }

}  // namespace mn

using namespace mn;

// ---- ABI ---------------------------------------------------------------------------------------------------
static int nem_greedy_impl(const double *R, int E, int S, const uint8_t *start, int n, int marginal, double logtype,
                           uint8_t *adj_out, int32_t *theta_out, double *score, double *max_trace, int32_t *n_iter,
                           int64_t *n_scored)
{
    MN_ARG(R && adj_out && theta_out, "NULL pointer");
    MN_ARG(E >= 1 && S >= 1 && S <= MNEMCU_MAX_S && n >= 1, "bad size");
    MN_ARG(!marginal || (logtype > 0 && logtype != 1.0), "logtype must be a valid log base");
    SearchBatch *sb = nullptr;
    MN_TRY(search_create(n, E, S, &sb));
    sb->marginal = marginal ? 1 : 0;
    sb->logtype = logtype;
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

extern "C" int mnemcu_nem_greedy(const double *R, int E, int S, const uint8_t *start, int n, uint8_t *adj_out,
                                 int32_t *theta_out, double *score, double *max_trace, int32_t *n_iter,
                                 int64_t *n_scored)
{
    return nem_greedy_impl(R, E, S, start, n, 0, 2.0, adj_out, theta_out, score, max_trace, n_iter, n_scored);
}

extern "C" int mnemcu_nem_greedy_marginal(const double *R, int E, int S, const uint8_t *start, int n, double logtype,
                                          uint8_t *adj_out, int32_t *theta_out, double *score, double *max_trace,
                                          int32_t *n_iter, int64_t *n_scored)
{
    return nem_greedy_impl(R, E, S, start, n, 1, logtype, adj_out, theta_out, score, max_trace, n_iter, n_scored);
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

static int score_adj_impl(const double *R, int E, int S, const uint8_t *adj, int n, int trans_close, int marginal,
                          double logtype, double *score, int32_t *theta, double *W)
{
    MN_ARG(R && adj && score, "NULL pointer");
    MN_ARG(E >= 1 && S >= 1 && S <= MNEMCU_MAX_S && n >= 1, "bad size");
    MN_ARG(!marginal || (logtype > 0 && logtype != 1.0), "logtype must be a valid log base");
    SearchBatch sb;
    sb.nmax = 1; sb.E = E; sb.S = S; sb.etiles = (E + kRows - 1) / kRows; sb.maxc = n;
    sb.marginal = marginal ? 1 : 0;
    sb.logtype = logtype;
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

extern "C" int mnemcu_score_adj(const double *R, int E, int S, const uint8_t *adj, int n, int trans_close,
                                double *score, int32_t *theta, double *W)
{
    return score_adj_impl(R, E, S, adj, n, trans_close, 0, 2.0, score, theta, W);
}

extern "C" int mnemcu_score_adj_marginal(const double *R, int E, int S, const uint8_t *adj, int n, int trans_close,
                                         double logtype, double *score, int32_t *theta, double *W)
{
    return score_adj_impl(R, E, S, adj, n, trans_close, 1, logtype, score, theta, W);
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
    SearchDev d;                              // candidates only: no dictionary, no chains
    d.S = S; d.maxc = maxc; d.Prow = drow.p; d.active = dact.p; d.Pcol = dPcol.p; d.cand_cols = dcc.p; d.cand_J = dJ.p;
    d.ncand = dn.p; d.cand_kind = dkind.p;
    MN_LAUNCH(gen_kernel, 1, 256, 0, 0, d, 1);
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
