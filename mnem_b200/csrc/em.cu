// The EM driver: mnem()::do_inits (R/mnems.r:1884-2165) for many starts in lock step, plus the standalone
// getProbs / doMean entry points built from the same pieces.
//
// Scheduling.  B "slots" each hold one EM start.  All slots advance together in ticks; one tick is
//   gamma = getAffinity(L, mw)          (affinity_kernel, all slots)
//   R     = doMean(D, gamma_i)          (ONE pass over D for all k*B (start, component) pairs)
//   structure step                      INIT slots: theta = argmax with the graph fixed (R/mnems_low.r:618-619)
//                                       EM slots:   2 greedy searches per component (R/mnems.r:2036-2073)
//   L'    = E-step                      (ONE pass over D for all pairs)
//   mixture weights / likelihood        (k x C reductions, 1 per INIT slot, 12 per EM slot)
// followed by the host-side accept / stop logic of the reference.  A slot that finishes is refilled with the
// next start, so the passes over D stay shared until the queue drains.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace mn {

// weights given as n row blocks of C (caller's cell order) -> [Cp][gw] sorted, padding cells and columns n.. 0
__global__ void weights_to_internal_kernel(const double *__restrict__ w, int n, int gw, int C, int Cp,
                                           const int32_t *__restrict__ orig, double *__restrict__ out)
{
    const int cp = blockIdx.x * blockDim.x + threadIdx.x;
    if (cp >= Cp) return;
    const int c = orig[cp];
    for (int j = 0; j < gw; ++j) out[(size_t)cp * gw + j] = (c >= 0 && j < n) ? w[(size_t)j * C + c] : 0.0;
}

// After the two greedy searches of every (slot, component) pair: keep the better one (R/mnems.r:2069-2073, tie ->
// the empty start), count edge / theta changes against the previous M-step (:2080-2083), install the new graph.
// grid = pairs in `pairs`, block = 128.
__global__ void __launch_bounds__(128) pick_kernel(int E, int S, const int32_t *__restrict__ pairs,
                                                   const int32_t *__restrict__ have_theta,
                                                   const double *__restrict__ maxtrace,
                                                   const uint32_t *__restrict__ adj_rows,
                                                   const uint32_t *__restrict__ closed_rows,
                                                   const int32_t *__restrict__ theta_s,
                                                   const int32_t *__restrict__ niter,
                                                   const unsigned long long *__restrict__ nscored,
                                                   uint32_t *__restrict__ res1_rows, uint32_t *__restrict__ clo_rows,
                                                   int32_t *__restrict__ theta, int32_t *__restrict__ edgechange,
                                                   int32_t *__restrict__ thetachange,
                                                   unsigned long long *__restrict__ counters)
{
    __shared__ int s_cnt[4];
    const int p = blockIdx.x, tid = threadIdx.x;
    const int m = pairs[p];
    const int q0 = 2 * p, q1 = 2 * p + 1;
    const int pick = maxtrace[q1] > maxtrace[q0] ? q1 : q0;       // which.max: first on ties
    int tc = 0;
    const int ht = have_theta[m];
    for (int e = tid; e < E; e += 128) {
        const int t = theta_s[(size_t)pick * E + e];
        if (ht && t != theta[(size_t)m * E + e]) ++tc;
        theta[(size_t)m * E + e] = t;
    }
    for (int o = 16; o > 0; o >>= 1) tc += __shfl_down_sync(kFull, tc, o);
    if ((tid & 31) == 0) s_cnt[tid >> 5] = tc;
    __syncthreads();
    if (tid < 32) {
        const uint32_t nr = adj_rows[pick * 32 + tid];
        const uint32_t old = res1_rows[m * 32 + tid];
        int ec = tid < S ? __popc(nr ^ old) : 0;                    // sum(abs(res$adj - res1$adj))
        for (int o = 16; o > 0; o >>= 1) ec += __shfl_down_sync(kFull, ec, o);
        res1_rows[m * 32 + tid] = nr;
        clo_rows[m * 32 + tid] = closed_rows[pick * 32 + tid];
        if (tid == 0) {
            edgechange[m] = ec;
            thetachange[m] = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
            atomicAdd(&counters[0], nscored[q0] + nscored[q1]);
            atomicAdd(&counters[1], (unsigned long long)(niter[q0] + niter[q1]));
        }
    }
}

__global__ void build_starts_kernel(const int32_t *__restrict__ pairs, const uint32_t *__restrict__ res1_rows,
                                    uint32_t *__restrict__ start_rows)
{
    const int p = blockIdx.x, lane = threadIdx.x;
    start_rows[(2 * p) * 32 + lane] = 0u;                                      // start0 * 0   R/mnems.r:2042
    start_rows[(2 * p + 1) * 32 + lane] = res1_rows[pairs[p] * 32 + lane];     // start0       R/mnems.r:2040
}

struct Slot {
    int start = -1;
    int phase = 0;                // 0 INIT (theta-free E-step iterations), 1 EM loop
    int inner = 0;
    int cur = 0, nw = 1, best = 0;
    double llold_inner = -INFINITY;
    // do_inits state
    double ll = -INFINITY, llold = -INFINITY, bestll = -INFINITY, probs_ll = -INFINITY;
    double maxll = -INFINITY;     // max(lls) over ALL iterations (limits$ll, R/mnems.r:2157), whatever the trace capacity
    bool time0 = true, stop = false, have_theta = false, pending_outer = false, pending_outer_guard = false;
    int count = 0;
    double mw[32], bestmw[32], mwold[32];
    std::vector<uint32_t> best_rows;     // [k*32]
    bool best_is_res1 = false;
};

struct Engine {
    mnemcu_ctx *cx = nullptr;
    int k = 0, B = 0, M = 0, T = 10, complete = 0, affinity = 0;    // affinity: mnem(affinity =), 1 = hard assignments
    // Slots 0..Be-1 are the ones in use (Be = B until the queue of starts is drained; then finished slots are squeezed
    // out, see compact()): the passes over D serve Me() = k * Be pairs, so the tail of a fit -- and every rank of an
    // N-GPU job spends a good part of its time there -- streams 5 pairs at HBM speed instead of 30 at the FP64 limit.
    int Be = 0;
    int Me() const { return k * Be; }
    int nullcomp = 0;             // mnem(nullcomp = TRUE): component k-1 of every slot is the null component
    bool fixed_theta = false;     // mnem(theta =): theta of a slot never moves (d_theta_fixed holds it)
    DevBuf<int32_t> d_theta_fixed;
    double logtype = 2.0;
    cudaStream_t st = nullptr;
    std::vector<Slot> slots;
    DevBuf<double> Lbuf, gam, R, Rg, partial, mixscratch, d_mw, d_mwtmp, d_ll, Lstage;
    DevBuf<int32_t> d_flag, d_theta, d_theta_tmp, d_theta_best, d_edge, d_thch, d_pairs, d_have_theta;
    DevBuf<uint32_t> d_res1, d_clo, d_clo_cols, d_start, maskT, d_best_rows;
    DevBuf<unsigned long long> d_counters;
    DevBuf<const double *> d_Rptr;
    DevBuf<uint8_t> d_u8;
    SearchBatch *sb = nullptr;
    // pinned readback
    double *h_ll = nullptr, *h_mw = nullptr;
    int32_t *h_chg = nullptr;
    unsigned long long *h_counters = nullptr;
    size_t kCp = 0;
    // phase timers: event pairs recorded on the stream and read back at the next point where the host waits for the
    // device anyway (no extra synchronisation on the critical path)
    static constexpr int kTimers = 16;
    cudaEvent_t tev[kTimers][2] = {};
    double *tacc[kTimers] = {};
    int ntimer = 0;
    mnemcu_em_stats stats{};
    // Decision tape (audit runs, mnemcu_em_opts::tape): every decision of the fit that compares floating-point sums --
    // accepted greedy moves, restart picks, theta attachments, EM-level accept / best updates -- as int32 records
    // [code, start, index, ...] (codes in include/mnemcu.h), so that the CPU oracle can replay the fit decision by decision.
    bool tape_on = false;
    std::vector<int32_t> tape;
    std::vector<int32_t> h_i32;
    std::vector<uint32_t> h_u32;
    std::vector<double> h_f64;
    void rec(std::initializer_list<int32_t> w) { tape.insert(tape.end(), w.begin(), w.end()); }

    ~Engine()
    {
        if (sb) search_destroy(sb);
        if (h_ll) cudaFreeHost(h_ll);
        if (h_mw) cudaFreeHost(h_mw);
        if (h_chg) cudaFreeHost(h_chg);
        if (h_counters) cudaFreeHost(h_counters);
        for (int i = 0; i < kTimers; ++i)
            for (int j = 0; j < 2; ++j)
                if (tev[i][j]) cudaEventDestroy(tev[i][j]);
    }

    double *buf(int slot, int which) { return Lbuf.p + ((size_t)slot * 3 + which) * kCp; }
    int32_t *flag(int slot, int which) { return d_flag.p + slot * 4 + which; }
    int free_buf(const Slot &s) const
    {
        for (int i = 0; i < 3; ++i)
            if (i != s.cur && i != s.best) return i;
        return 0;
    }

    int init(mnemcu_ctx *c, int k_, int B_, int complete_, double logtype_, bool with_search)
    {
        cx = c; k = k_; B = B_; M = k * B; Be = B; complete = complete_; logtype = logtype_;
        MN_ARG(k >= 1 && B >= 1 && M <= 32, "k * batch must be in 1..32");
        MN_ARG(logtype > 0 && logtype != 1.0, "logtype must be a valid log base");
        MN_CU(cudaSetDevice(cx->device));
        st = cx->stream;
        T = cx->C <= 100 ? 100 : 10;                                           // R/mnems_low.r:584-588
        kCp = (size_t)k * cx->Cp;
        slots.assign(B, Slot());
        MN_TRY(Lbuf.alloc((size_t)B * 3 * kCp, true));
        MN_TRY(gam.alloc((size_t)gamma_stride(M) * cx->Cp, true));     // [Cp][gw], columns M.. stay 0
        MN_TRY(R.alloc((size_t)M * cx->S * cx->E, true));
        MN_TRY(partial.alloc(mstats_partial_elems(cx, M)));
        MN_TRY(mixscratch.alloc((size_t)B * kMixBlocks * (k + 1)));
        MN_TRY(d_mw.alloc((size_t)B * 32, true)); MN_TRY(d_mwtmp.alloc((size_t)B * 32, true));
        MN_TRY(d_ll.alloc((size_t)B * 128, true));
        MN_TRY(d_flag.alloc((size_t)B * 4, true));
        MN_TRY(d_theta.alloc((size_t)M * cx->E, true)); MN_TRY(d_theta_tmp.alloc((size_t)M * cx->E, true));
        MN_TRY(d_theta_best.alloc((size_t)M * cx->E, true));
        MN_TRY(d_edge.alloc(M, true)); MN_TRY(d_thch.alloc(M, true)); MN_TRY(d_pairs.alloc(M, true));
        MN_TRY(d_have_theta.alloc(M, true));
        MN_TRY(d_res1.alloc((size_t)M * 32, true)); MN_TRY(d_clo.alloc((size_t)M * 32, true));
        MN_TRY(d_clo_cols.alloc((size_t)M * 32, true)); MN_TRY(d_start.alloc((size_t)2 * M * 32, true));
        MN_TRY(d_best_rows.alloc((size_t)M * 32, true));
        MN_TRY(maskT.alloc((size_t)cx->G * cx->E, true));
        if (cx->multi) MN_TRY(Rg.alloc((size_t)M * cx->G * cx->E, true));
        MN_TRY(d_counters.alloc(2, true));
        MN_TRY(d_Rptr.alloc(M));
        MN_TRY(d_u8.alloc((size_t)M * cx->S * cx->S));
        MN_TRY(Lstage.alloc((size_t)k * cx->C));
        std::vector<const double *> rp(M);
        for (int m = 0; m < M; ++m) rp[m] = R.p + (size_t)m * cx->S * cx->E;
        MN_CU(cudaMemcpy(d_Rptr.p, rp.data(), sizeof(double *) * M, cudaMemcpyHostToDevice));
        if (with_search) MN_TRY(search_create(2 * M, cx->E, cx->S, &sb));
        MN_CU(cudaMallocHost((void **)&h_ll, sizeof(double) * B * 128));
        MN_CU(cudaMallocHost((void **)&h_mw, sizeof(double) * B * 32));
        MN_CU(cudaMallocHost((void **)&h_chg, sizeof(int32_t) * 2 * M));
        MN_CU(cudaMallocHost((void **)&h_counters, sizeof(unsigned long long) * 2));
        for (int i = 0; i < kTimers; ++i) { MN_CU(cudaEventCreate(&tev[i][0])); MN_CU(cudaEventCreate(&tev[i][1])); }
        for (int b = 0; b < B; ++b) slots[b].best_rows.assign((size_t)k * 32, 0u);
        // the zero fills of DevBuf::alloc ran on the legacy default stream; the work stream does not synchronise with it
        MN_CU(cudaDeviceSynchronize());
        return 0;
    }

    // ---- timing helpers ---------------------------------------------------------------------------------
    int tic()
    {
        if (ntimer == kTimers) MN_TRY(flush_timers());
        MN_CU(cudaEventRecord(tev[ntimer][0], st));
        return 0;
    }
    int toc(double *acc)
    {
        MN_CU(cudaEventRecord(tev[ntimer][1], st));
        tacc[ntimer++] = acc;
        return 0;
    }
    int flush_timers()          // waits for the last recorded event: call where the host synchronises anyway
    {
        for (int i = 0; i < ntimer; ++i) {
            MN_CU(cudaEventSynchronize(tev[i][1]));
            float ms = 0.f;
            MN_CU(cudaEventElapsedTime(&ms, tev[i][0], tev[i][1]));
            *tacc[i] += ms;
        }
        ntimer = 0;
        return 0;
    }

    // ---- primitives over sets of slots ------------------------------------------------------------------
    // graphs (uint8, k per slot) -> row masks as given + closure rows/cols
    int load_graphs(int slot, const uint8_t *phi_host)
    {
        const int S = cx->S;
        uint8_t *d = d_u8.p + (size_t)slot * k * S * S;
        MN_CU(cudaMemcpyAsync(d, phi_host, (size_t)k * S * S, cudaMemcpyHostToDevice, st));
        MN_TRY(launch_rows_from_u8(d, k, S, d_res1.p + (size_t)slot * k * 32, st));
        MN_TRY(launch_closure(d_res1.p + (size_t)slot * k * 32, k, S, d_clo.p + (size_t)slot * k * 32,
                              d_clo_cols.p + (size_t)slot * k * 32, st));
        return 0;
    }

    int set_mw(int slot, const double *mw_host)
    {
        for (int i = 0; i < k; ++i) slots[slot].mw[i] = mw_host ? mw_host[i] : 1.0 / k;
        MN_CU(cudaMemcpyAsync(d_mw.p + slot * 32, slots[slot].mw, sizeof(double) * k, cudaMemcpyHostToDevice, st));
        return 0;
    }

    // mixture statistics for a list of slots.  which[j]: buffer index; from_tmp/to_tmp: read/write the inner-loop mw
    // copy instead of the outer mw; ll goes to d_ll[slot*128 + ll_idx].
    int op_stats(const std::vector<int> &list, const std::vector<int> &which, bool from_tmp, bool to_tmp, int ll_idx,
                 bool guard)
    {
        if (list.empty()) return 0;
        MixSlots ms{};
        ms.n = (int)list.size();
        for (int j = 0; j < ms.n; ++j) {
            const int b = list[j];
            ms.L[j] = buf(b, which[j]);
            ms.mw[j] = (from_tmp ? d_mwtmp.p : d_mw.p) + b * 32;
            ms.anypos[j] = flag(b, which[j]);
            ms.ll_out[j] = d_ll.p + b * 128 + ll_idx;
            ms.mw_out[j] = (to_tmp ? d_mwtmp.p : d_mw.p) + b * 32;
            ms.na_guard[j] = guard ? 1 : 0;
        }
        return launch_mix_stats(cx, ms, k, affinity, complete, logtype, mixscratch.p, st);
    }

    int op_affinity(const std::vector<int> &list)
    {
        if (list.empty()) return 0;
        MixSlots ms{};
        ms.n = (int)list.size();
        for (int j = 0; j < ms.n; ++j) {
            const int b = list[j];
            ms.L[j] = buf(b, slots[b].cur);
            ms.gam[j] = gam.p + (size_t)b * k;
            ms.mw[j] = d_mw.p + b * 32;
            ms.anypos[j] = flag(b, slots[b].cur);
        }
        return launch_affinity(cx, ms, k, gamma_stride(Me()), affinity, complete, logtype, st);
    }

    int op_mstats(bool unweighted = false)
    {
        MN_TRY(tic());
        MN_TRY(launch_mstats(cx, unweighted ? nullptr : gam.p, Me(), partial.p, R.p, st));
        MN_TRY(toc(&stats.ms_mstats));
        stats.passes_mstats++;
        return 0;
    }

    // theta = argmax with the closed graph fixed, null column last (R/mnems_low.r:617-619) for the listed slots
    int op_theta_init(const std::vector<int> &list)
    {
        if (list.empty()) return 0;
        if (cx->multi) {
            MN_TRY(launch_mstats_groups(cx, partial.p, Me(), Rg.p, st));
            MN_TRY(launch_theta_groups(cx, Rg.p, Me(), d_clo_cols.p, d_theta_tmp.p, st));
        } else {
            MN_TRY(launch_theta(d_Rptr.p, Me(), cx->E, cx->S, d_clo_cols.p, 1, d_theta_tmp.p, nullptr, st));
        }
        for (int b : list)
            MN_CU(cudaMemcpyAsync(d_theta.p + (size_t)b * k * cx->E, d_theta_tmp.p + (size_t)b * k * cx->E,
                                  sizeof(int32_t) * (size_t)k * cx->E, cudaMemcpyDeviceToDevice, st));
        if (nullcomp)                                                    // tmp <- 0 for the null component, R/mnems_low.r:607-608
            for (int b : list)
                MN_CU(cudaMemsetAsync(d_theta.p + ((size_t)b * k + k - 1) * cx->E, 0, sizeof(int32_t) * cx->E, st));
        if (tape_on) MN_TRY(tape_theta(list, 1));
        return 0;
    }

    // tape: theta of the listed slots (code 1: inside the theta-free getProbs, index = inner iteration; code 4: after an
    // M-step, index = EM iteration)
    int tape_theta(const std::vector<int> &list, int code)
    {
        const size_t kE = (size_t)k * cx->E;
        h_i32.resize(kE);
        for (int b : list) {
            MN_CU(cudaMemcpyAsync(h_i32.data(), d_theta.p + (size_t)b * kE, sizeof(int32_t) * kE, cudaMemcpyDeviceToHost, st));
            MN_CU(cudaStreamSynchronize(st));
            rec({code, slots[b].start, code == 1 ? slots[b].inner : slots[b].count});
            tape.insert(tape.end(), h_i32.begin(), h_i32.end());
        }
        return 0;
    }

    // tape: the accepted moves of every search of the tick (code 2) and the restart pick of every pair (code 3)
    int tape_searches(const std::vector<int32_t> &pairs)
    {
        const int np = (int)pairs.size(), S = cx->S, cap = search_hard_cap(sb);
        std::vector<int32_t> nit(2 * np);
        std::vector<double> mtr(2 * np);
        MN_CU(cudaMemcpyAsync(nit.data(), search_niter(sb), sizeof(int32_t) * 2 * np, cudaMemcpyDeviceToHost, st));
        MN_CU(cudaMemcpyAsync(mtr.data(), search_maxtrace(sb), sizeof(double) * 2 * np, cudaMemcpyDeviceToHost, st));
        MN_CU(cudaStreamSynchronize(st));
        for (int q = 0; q < 2 * np; ++q) {
            const int m = pairs[q >> 1], b = m / k, comp = m - b * k;
            h_u32.resize((size_t)std::max(1, nit[q]) * 32);
            if (nit[q] > 0) {
                MN_CU(cudaMemcpyAsync(h_u32.data(), search_tape(sb) + (size_t)q * cap * 32, sizeof(uint32_t) * 32 * nit[q],
                                      cudaMemcpyDeviceToHost, st));
                MN_CU(cudaStreamSynchronize(st));
            }
            rec({2, slots[b].start, slots[b].count, comp, q & 1, nit[q]});
            for (int i = 0; i < nit[q]; ++i)
                for (int r = 0; r < S; ++r) tape.push_back((int32_t)h_u32[(size_t)i * 32 + r]);
            if (q & 1) rec({3, slots[b].start, slots[b].count, comp, mtr[q] > mtr[q - 1] ? 1 : 0});   // = pick_kernel
        }
        return 0;
    }

    // E-step for all slots into their `nw` buffers
    int op_estep()
    {
        MN_TRY(launch_build_masks(cx, d_clo.p, d_theta.p, Me(), maskT.p, st));
        PairPtrs pp{};
        for (int b = 0; b < Be; ++b) {
            MN_CU(cudaMemsetAsync(flag(b, slots[b].nw), 0, sizeof(int32_t), st));
            for (int i = 0; i < k; ++i) {
                pp.L[b * k + i] = buf(b, slots[b].nw) + (size_t)i * cx->Cp;
                pp.flag[b * k + i] = flag(b, slots[b].nw);
            }
        }
        MN_TRY(tic());
        MN_TRY(launch_estep(cx, maskT.p, Me(), pp, st));
        MN_TRY(toc(&stats.ms_estep));
        stats.passes_estep++;
        stats.estep_bytes += 8.0 * cx->E * cx->C + 8.0 * (double)Me() * cx->C;
        return 0;
    }

    // everything slot `from` holds on the device and on the host moves to the (free) slot `to`
    int move_slot(int from, int to)
    {
        const size_t kE = (size_t)k * cx->E;
        auto cp = [&](void *dst, const void *src, size_t bytes) { return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st); };
        for (int i = 0; i < 3; ++i) MN_CU(cp(buf(to, i), buf(from, i), sizeof(double) * kCp));
        MN_CU(cp(d_flag.p + to * 4, d_flag.p + from * 4, sizeof(int32_t) * 4));
        MN_CU(cp(d_theta.p + to * kE, d_theta.p + from * kE, sizeof(int32_t) * kE));
        MN_CU(cp(d_theta_best.p + to * kE, d_theta_best.p + from * kE, sizeof(int32_t) * kE));
        if (fixed_theta) MN_CU(cp(d_theta_fixed.p + to * kE, d_theta_fixed.p + from * kE, sizeof(int32_t) * kE));
        for (DevBuf<uint32_t> *g : {&d_res1, &d_clo, &d_clo_cols, &d_best_rows})
            MN_CU(cp(g->p + (size_t)to * k * 32, g->p + (size_t)from * k * 32, sizeof(uint32_t) * k * 32));
        MN_CU(cp(d_mw.p + to * 32, d_mw.p + from * 32, sizeof(double) * 32));
        MN_CU(cp(d_mwtmp.p + to * 32, d_mwtmp.p + from * 32, sizeof(double) * 32));
        MN_CU(cp(d_ll.p + to * 128, d_ll.p + from * 128, sizeof(double) * 128));
        std::swap(slots[to], slots[from]);
        slots[from].start = -1;
        return 0;
    }

    // Once no start is left to refill a finished slot: move the active slots to the front and shrink the passes to them.
    // A start's fit does not depend on which slot it sits in nor on how many pairs share its passes (tested bit for bit).
    int compact()
    {
        for (;;) {
            int hi = -1, lo = -1;
            for (int b = 0; b < Be; ++b) {
                if (slots[b].start >= 0) hi = b;
                else if (lo < 0) lo = b;
            }
            if (hi < 0 || lo < 0 || lo > hi) break;
            MN_TRY(move_slot(hi, lo));
        }
        int nb = 0;
        while (nb < Be && slots[nb].start >= 0) ++nb;
        nb = std::max(1, nb);
        if (nb != Be) {
            // gamma is [Cp][gamma_stride(pairs)] with zero padding columns: a new stride needs a clean buffer
            if (gamma_stride(k * nb) != gamma_stride(Me())) MN_CU(cudaMemsetAsync(gam.p, 0, sizeof(double) * gam.n, st));
            Be = nb;
        }
        return 0;
    }

    // greedy searches for the listed (EM) slots, then winner selection and change counts
    int op_search(const std::vector<int> &list)
    {
        if (list.empty()) return 0;
        std::vector<int32_t> pairs, ht(M, 0);
        for (int b : list)
            for (int i = 0; i < k - (nullcomp ? 1 : 0); ++i) pairs.push_back(b * k + i);      // no M-step for a null component
        for (int b = 0; b < B; ++b)
            for (int i = 0; i < k; ++i) ht[b * k + i] = slots[b].have_theta ? 1 : 0;
        const int np = (int)pairs.size();
        MN_CU(cudaMemcpyAsync(d_pairs.p, pairs.data(), sizeof(int32_t) * np, cudaMemcpyHostToDevice, st));
        MN_CU(cudaMemcpyAsync(d_have_theta.p, ht.data(), sizeof(int32_t) * M, cudaMemcpyHostToDevice, st));
        MN_LAUNCH(build_starts_kernel, np, 32, 0, st, d_pairs.p, d_res1.p, d_start.p);
        std::vector<const double *> rp(2 * np);
        for (int p = 0; p < np; ++p) rp[2 * p] = rp[2 * p + 1] = R.p + (size_t)pairs[p] * cx->S * cx->E;
        // (pairs / ht are stack vectors: copies from pageable memory are staged before the call returns)
        MN_TRY(tic());
        MN_TRY(search_run(sb, 2 * np, rp.data(), d_start.p, st));
        MN_LAUNCH(pick_kernel, np, 128, 0, st, cx->E, cx->S, d_pairs.p, d_have_theta.p, search_maxtrace(sb),
                  search_adj_rows(sb), search_closed_rows(sb), search_theta(sb), search_niter(sb), search_nscored(sb),
                  d_res1.p, d_clo.p, d_theta.p, d_edge.p, d_thch.p, d_counters.p);
        if (fixed_theta)                                                  // nem(subtopo = thetaX) returns thetaX: theta never moves
            for (int b : list) {
                const size_t kE = (size_t)k * cx->E;
                MN_CU(cudaMemcpyAsync(d_theta.p + b * kE, d_theta_fixed.p + b * kE, sizeof(int32_t) * kE, cudaMemcpyDeviceToDevice, st));
                MN_CU(cudaMemsetAsync(d_thch.p + (size_t)b * k, 0, sizeof(int32_t) * k, st));
            }
        MN_TRY(toc(&stats.ms_search));
        if (tape_on) { MN_TRY(tape_searches(pairs)); MN_TRY(tape_theta(list, 4)); }
        return 0;
    }
};

}  // namespace mn

using namespace mn;

// ---------------------------------------------------------------------------------------------------------
// doMean
// ---------------------------------------------------------------------------------------------------------
extern "C" int mnemcu_do_mean(mnemcu_ctx *cx, const double *weights, int n, double *R_out)
{
    MN_ARG(cx && R_out, "NULL pointer");
    MN_ARG(n >= 1, "n must be positive");
    MN_ARG(weights || n == 1, "unweighted doMean takes n = 1");
    MN_CU(cudaSetDevice(cx->device));
    cudaStream_t st = cx->stream;
    const size_t ES = (size_t)cx->E * cx->S;
    DevBuf<double> dw, dgam, dpart, dR;
    const int G = std::min(n, 20);
    MN_TRY(dgam.alloc((size_t)gamma_stride(G) * cx->Cp));
    MN_TRY(dpart.alloc(mstats_partial_elems(cx, G)));
    MN_TRY(dR.alloc((size_t)G * ES));
    if (weights) MN_TRY(dw.alloc((size_t)G * cx->C));
    for (int j0 = 0; j0 < n; j0 += G) {
        const int g = std::min(G, n - j0);
        if (weights) {
            MN_CU(cudaMemcpyAsync(dw.p, weights + (size_t)j0 * cx->C, sizeof(double) * (size_t)g * cx->C,
                                  cudaMemcpyHostToDevice, st));
            MN_LAUNCH(weights_to_internal_kernel, (cx->Cp + 255) / 256, 256, 0, st, dw.p, g, gamma_stride(g), cx->C, cx->Cp,
                      cx->orig, dgam.p);
        }
        MN_TRY(launch_mstats(cx, weights ? dgam.p : nullptr, g, dpart.p, dR.p, st));
        MN_CU(cudaMemcpyAsync(R_out + (size_t)j0 * ES, dR.p, sizeof(double) * (size_t)g * ES, cudaMemcpyDeviceToHost, st));
        MN_CU(cudaStreamSynchronize(st));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// getProbs
// ---------------------------------------------------------------------------------------------------------
extern "C" int mnemcu_get_probs(mnemcu_ctx *cx, int n, int k, const uint8_t *phi, const int32_t *theta,
                                const double *L_in, const double *mw_in, int complete, double logtype, double *L_out,
                                double *mw_out, double *ll_out, int32_t *theta_used)
{
    MN_ARG(cx && phi && L_in && L_out && mw_out && ll_out, "NULL pointer");
    MN_ARG(n >= 1 && k >= 1 && k <= MNEMCU_MAX_K, "bad n / k");
    const int E = cx->E, C = cx->C, S = cx->S;
    if (theta)
        for (size_t i = 0; i < (size_t)n * k * E; ++i) MN_ARG(theta[i] >= 0 && theta[i] <= S, "theta outside 0..S");
    const int B = std::max(1, std::min(n, 20 / k > 0 ? 20 / k : 1));
    Engine en;
    MN_TRY(en.init(cx, k, B, complete, logtype, false));
    cudaStream_t st = en.st;
    const size_t kC = (size_t)k * C;
    for (int g0 = 0; g0 < n; g0 += B) {
        const int nb = std::min(B, n - g0);
        std::vector<int> list;
        for (int b = 0; b < nb; ++b) list.push_back(b);
        for (int b = 0; b < B; ++b) {
            Slot &s = en.slots[b];
            s = Slot();
            s.best_rows.assign((size_t)k * 32, 0u);
            s.cur = 0; s.best = 0; s.nw = 1;
            const int src = g0 + std::min(b, nb - 1);          // unused slots replay the last mixture (results dropped)
            MN_TRY(en.load_graphs(b, phi + (size_t)src * k * S * S));
            MN_CU(cudaMemcpyAsync(en.Lstage.p, L_in + (size_t)src * kC, sizeof(double) * kC, cudaMemcpyHostToDevice, st));
            MN_TRY(launch_L_to_internal(cx, en.Lstage.p, k, en.buf(b, 0), st));
            MN_CU(cudaMemsetAsync(en.flag(b, 0), 0, sizeof(int32_t), st));
            MN_TRY(launch_anypos(cx, en.buf(b, 0), k, en.flag(b, 0), st));
            MN_TRY(en.set_mw(b, mw_in ? mw_in + (size_t)src * k : nullptr));
            if (theta)
                MN_CU(cudaMemcpyAsync(en.d_theta.p + (size_t)b * k * E, theta + (size_t)src * k * E,
                                      sizeof(int32_t) * (size_t)k * E, cudaMemcpyHostToDevice, st));
            MN_CU(cudaStreamSynchronize(st));                   // Lstage is reused
        }
        std::vector<int> all;
        for (int b = 0; b < B; ++b) all.push_back(b);
        std::vector<int> wcur(B), wnew(B);
        auto refresh = [&]() { for (int b = 0; b < B; ++b) { wcur[b] = en.slots[b].cur; wnew[b] = en.slots[b].nw; } };
        refresh();
        // mw <- normalise(rowSums(getAffinity(probsold, mw))), NA -> 1/k        R/mnems_low.r:591-595
        MN_TRY(en.op_stats(all, wcur, false, false, 127, true));
        std::vector<double> llold(B, -INFINITY);
        if (theta) {
            MN_TRY(en.op_estep());
            for (int t = 0; t < en.T; ++t) MN_TRY(en.op_stats(all, wnew, false, false, t, false));
            MN_CU(cudaMemcpyAsync(en.h_ll, en.d_ll.p, sizeof(double) * B * 128, cudaMemcpyDeviceToHost, st));
            MN_CU(cudaStreamSynchronize(st));
            for (int b = 0; b < B; ++b) {
                Slot &s = en.slots[b];
                // bestprobs <- probs0 at the first inner iteration whose ll beats -Inf (:646-648)
                bool improved = false;
                double lo = -INFINITY;
                for (int t = 0; t < en.T; ++t) { if (en.h_ll[b * 128 + t] - lo > 0) improved = true; lo = en.h_ll[b * 128 + t]; }
                if (improved) s.best = s.nw;
                llold[b] = en.h_ll[b * 128 + en.T - 1];
            }
        } else {
            for (int t = 0; t < en.T; ++t) {
                MN_TRY(en.op_affinity(all));
                MN_TRY(en.op_mstats());
                MN_TRY(en.op_theta_init(all));
                MN_TRY(en.op_estep());
                refresh();
                MN_TRY(en.op_stats(all, wnew, false, false, 0, false));
                MN_CU(cudaMemcpyAsync(en.h_ll, en.d_ll.p, sizeof(double) * B * 128, cudaMemcpyDeviceToHost, st));
                MN_CU(cudaStreamSynchronize(st));
                for (int b = 0; b < B; ++b) {
                    Slot &s = en.slots[b];
                    const double ll0 = en.h_ll[b * 128];
                    if (ll0 - llold[b] > 0) s.best = s.nw;
                    s.cur = s.nw;
                    s.nw = en.free_buf(s);
                    llold[b] = ll0;
                }
                refresh();
            }
        }
        MN_CU(cudaMemcpyAsync(en.h_mw, en.d_mw.p, sizeof(double) * B * 32, cudaMemcpyDeviceToHost, st));
        MN_CU(cudaStreamSynchronize(st));
        for (int b = 0; b < nb; ++b) {
            Slot &s = en.slots[b];
            MN_TRY(launch_L_from_internal(cx, en.buf(b, s.best), k, en.Lstage.p, st));
            MN_CU(cudaMemcpyAsync(L_out + (size_t)(g0 + b) * kC, en.Lstage.p, sizeof(double) * kC, cudaMemcpyDeviceToHost, st));
            if (theta_used)
                MN_CU(cudaMemcpyAsync(theta_used + (size_t)(g0 + b) * k * E, en.d_theta.p + (size_t)b * k * E,
                                      sizeof(int32_t) * (size_t)k * E, cudaMemcpyDeviceToHost, st));
            MN_CU(cudaStreamSynchronize(st));
            for (int i = 0; i < k; ++i) mw_out[(size_t)(g0 + b) * k + i] = en.h_mw[b * 32 + i];
            ll_out[g0 + b] = llold[b];
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// the EM loop
// ---------------------------------------------------------------------------------------------------------
extern "C" int mnemcu_em_run(mnemcu_ctx *cx, int starts, int k, const uint8_t *init_phi, const mnemcu_em_opts *opts,
                             uint8_t *adj_out, int32_t *theta_out, double *probs_out, double *mw_out, double *ll_out,
                             int32_t *n_iter_out, double *lls, double *phievo, double *thetaevo, double *mwevo,
                             int32_t *best_out, mnemcu_em_stats *stats_out)
{
    MN_ARG(cx && init_phi && opts && adj_out && theta_out && mw_out && ll_out && n_iter_out && best_out, "NULL pointer");
    MN_ARG(starts >= 1 && k >= 2 && k <= MNEMCU_MAX_K, "starts must be positive and k in 2..32 (k = 1 is nem() + getProbs)");
    MN_ARG(opts->max_trace >= 1 || !lls, "max_trace must be positive");
    const int E = cx->E, C = cx->C, S = cx->S;
    // EM starts per pass over D.  Default: as many as give k * B <= 30 pairs.  Up to 20 pairs the two streaming kernels are
    // HBM-bound (E-step 0.89 of the measured copy bandwidth at 20), beyond that they are bound by the FP64 rate, but a pass
    // still gets cheaper per start (cfg5: (2.81 + 3.43) / 4 = 1.56 ms per start and tick at B = 4, (3.94 + 4.47) / 6 =
    // 1.40 ms at B = 6) and the device-resident search loop has more searches to fill the SMs with.
    int B = opts->batch > 0 ? opts->batch : std::max(1, 30 / k);
    B = std::min(B, std::min(starts, 32 / k));
    MN_ARG(opts->affinity == 0 || opts->affinity == 1, "affinity must be 0 or 1");
    MN_ARG(!(opts->affinity && opts->complete), "affinity = 1 with complete = TRUE is an error in the reference too");
    Engine en;
    MN_TRY(en.init(cx, k, B, opts->complete, opts->logtype, true));
    en.affinity = opts->affinity;
    en.nullcomp = opts->nullcomp ? 1 : 0;
    if (opts->theta) {
        for (size_t i = 0; i < (size_t)starts * k * E; ++i) MN_ARG(opts->theta[i] >= 0 && opts->theta[i] <= S, "theta outside 0..S");
        en.fixed_theta = true;
        MN_TRY(en.d_theta_fixed.alloc((size_t)en.M * E));
    }
    MN_ARG(opts->tape_cap >= 0 && (opts->tape || opts->tape_cap == 0), "tape_cap without a tape buffer");
    if (opts->tape) { en.tape_on = true; MN_TRY(search_enable_tape(en.sb)); }
    cudaStream_t st = en.st;
    const int T = en.T;
    const size_t kC = (size_t)k * C, kE = (size_t)k * E, kSS = (size_t)k * S * S;
    const int mt = opts->max_trace;
    int next_seq = 0, running = 0;
    // The next EM start: 0, 1, 2, ... or, when several callers (one per GPU) drain one queue, whatever the caller's
    // ticket function hands out.  Results are always stored at the start's own index.
    for (int s = 0; s < starts; ++s) { n_iter_out[s] = -1; ll_out[s] = -INFINITY; }
    auto take_start = [&]() -> int {
        if (opts->next_start) {
            const int s = opts->next_start(opts->next_start_user);
            return (s >= 0 && s < starts) ? s : -1;
        }
        return next_seq < starts ? next_seq++ : -1;
    };
    MN_ARG(opts->probs_mode == 0 || opts->probs_mode == 1, "probs_mode must be 0 or 1");
    if (opts->mw)
        for (int i = 0; i < k; ++i) MN_ARG(opts->mw[i] >= 0.0 && std::isfinite(opts->mw[i]), "mw must be finite and non-negative");
    const bool winner_probs = probs_out && opts->probs_mode == 1;
    DevBuf<double> winner_L;                  // probs of the best start so far, internal layout
    if (winner_probs) MN_TRY(winner_L.alloc(en.kCp));
    cudaEvent_t evA, evB;
    MN_CU(cudaEventCreate(&evA)); MN_CU(cudaEventCreate(&evB));
    MN_CU(cudaEventRecord(evA, st));
    double winner_ll = -INFINITY;
    int winner = -1;

    auto finish_slot = [&](int b) -> int {
        Slot &s = en.slots[b];
        const int sidx = s.start;
        // final best update, R/mnems.r:2144-2147
        if (s.ll - s.bestll > 0) {
            if (en.tape_on) en.rec({7, sidx, s.count});
            s.bestll = s.ll;
            MN_CU(cudaMemcpyAsync(en.d_best_rows.p + (size_t)b * k * 32, en.d_res1.p + (size_t)b * k * 32,
                                  sizeof(uint32_t) * k * 32, cudaMemcpyDeviceToDevice, st));
            MN_CU(cudaMemcpyAsync(en.d_theta_best.p + (size_t)b * kE, en.d_theta.p + (size_t)b * kE, sizeof(int32_t) * kE,
                                  cudaMemcpyDeviceToDevice, st));
            s.best = s.cur;
            for (int i = 0; i < k; ++i) s.bestmw[i] = s.mw[i];
        }
        MN_TRY(launch_rows_to_u8(en.d_best_rows.p + (size_t)b * k * 32, k, S, en.d_u8.p + (size_t)b * kSS, st));
        MN_CU(cudaMemcpyAsync(adj_out + (size_t)sidx * kSS, en.d_u8.p + (size_t)b * kSS, kSS, cudaMemcpyDeviceToHost, st));
        MN_CU(cudaMemcpyAsync(theta_out + (size_t)sidx * kE, en.d_theta_best.p + (size_t)b * kE, sizeof(int32_t) * kE,
                              cudaMemcpyDeviceToHost, st));
        if (probs_out && !winner_probs) {
            MN_TRY(launch_L_from_internal(cx, en.buf(b, s.best), k, en.Lstage.p, st));
            MN_CU(cudaMemcpyAsync(probs_out + (size_t)sidx * kC, en.Lstage.p, sizeof(double) * kC, cudaMemcpyDeviceToHost, st));
        }
        MN_CU(cudaStreamSynchronize(st));
        for (int i = 0; i < k; ++i) mw_out[(size_t)sidx * k + i] = s.bestmw[i];
        const double mx = s.maxll;                                          // limits$ll <- max(lls)  :2157
        ll_out[sidx] = mx;
        n_iter_out[sidx] = s.count;
        if (en.tape_on) en.rec({8, sidx, s.count});
        en.stats.em_iters += s.count;
        // `<=` in an upward scan over the starts (:2222-2230): the maximal start with the LARGEST index wins, whatever
        // the order in which the slots finish
        if (winner < 0 || mx > winner_ll || (mx == winner_ll && sidx > winner)) {
            winner_ll = mx; winner = sidx;
            if (winner_probs) {
                MN_CU(cudaMemcpyAsync(winner_L.p, en.buf(b, s.best), sizeof(double) * en.kCp, cudaMemcpyDeviceToDevice, st));
                MN_CU(cudaStreamSynchronize(st));          // the slot's buffers are reused by the next start
            }
        }
        s.start = -1;
        --running;
        return 0;
    };

    auto begin_slot = [&](int b, int sidx) -> int {
        Slot &s = en.slots[b];
        std::vector<uint32_t> keep;
        keep.swap(s.best_rows);
        s = Slot();
        s.best_rows.swap(keep);
        s.start = sidx;
        s.phase = 0; s.inner = 0; s.cur = 0; s.best = 0; s.nw = 1;
        MN_TRY(en.load_graphs(b, init_phi + (size_t)sidx * kSS));
        if (en.fixed_theta) {                                                     // res1[[i]]$subtopo <- theta[[s]][[i]]  :1890-1892
            MN_CU(cudaMemcpyAsync(en.d_theta_fixed.p + (size_t)b * kE, opts->theta + (size_t)sidx * kE, sizeof(int32_t) * kE,
                                  cudaMemcpyHostToDevice, st));
            MN_CU(cudaMemcpyAsync(en.d_theta.p + (size_t)b * kE, en.d_theta_fixed.p + (size_t)b * kE, sizeof(int32_t) * kE,
                                  cudaMemcpyDeviceToDevice, st));
            MN_CU(cudaStreamSynchronize(st));                                     // opts->theta is pageable caller memory
        }
        if (en.nullcomp) {                                                        // res1[[k]]$adj <- adj * 0, subtopo * 0  :1894-1898
            MN_CU(cudaMemsetAsync(en.d_res1.p + ((size_t)b * k + k - 1) * 32, 0, sizeof(uint32_t) * 32, st));
            MN_CU(cudaMemsetAsync(en.d_theta.p + ((size_t)b * k + k - 1) * E, 0, sizeof(int32_t) * E, st));
            if (en.fixed_theta)
                MN_CU(cudaMemsetAsync(en.d_theta_fixed.p + ((size_t)b * k + k - 1) * E, 0, sizeof(int32_t) * E, st));
        }
        MN_CU(cudaMemsetAsync(en.buf(b, 0), 0, sizeof(double) * en.kCp, st));      // probs <- matrix(0, k, C)  :1905
        MN_CU(cudaMemsetAsync(en.flag(b, 0), 0, sizeof(int32_t), st));
        MN_TRY(en.set_mw(b, opts->mw));                                           // mw <- rep(1, k)/k unless given  :1771-1773
        // getProbs preamble, R/mnems_low.r:591-595
        MN_TRY(en.op_stats({b}, {0}, false, false, 127, true));
        // best graphs default to the initial ones (bestres is only assigned inside the loop)
        MN_CU(cudaMemcpyAsync(en.d_best_rows.p + (size_t)b * k * 32, en.d_res1.p + (size_t)b * k * 32,
                              sizeof(uint32_t) * k * 32, cudaMemcpyDeviceToDevice, st));
        MN_CU(cudaMemsetAsync(en.d_theta_best.p + (size_t)b * kE, 0, sizeof(int32_t) * kE, st));
        ++running;
        return 0;
    };

    auto body = [&]() -> int {
        for (int b = 0; b < B; ++b) {
            const int s0 = take_start();
            if (s0 < 0) break;
            MN_TRY(begin_slot(b, s0));
        }
        long guard_ticks = 0;
        bool drained = false;
        while (running > 0) {
            ++guard_ticks;
            std::vector<int> act, ini, em;
            for (int b = 0; b < en.Be; ++b)
                if (en.slots[b].start >= 0) {
                    act.push_back(b);
                    (en.slots[b].phase == 0 ? ini : em).push_back(b);
                }
            // ---- top of the EM loop body for EM slots: best tracking, R/mnems.r:1959-1966 ----
            for (int b : em) {
                Slot &s = en.slots[b];
                if (!s.time0) {
                    if (s.ll - s.bestll > 0) {
                        if (en.tape_on) en.rec({7, s.start, s.count});
                        s.bestll = s.ll;
                        MN_CU(cudaMemcpyAsync(en.d_best_rows.p + (size_t)b * k * 32, en.d_res1.p + (size_t)b * k * 32,
                                              sizeof(uint32_t) * k * 32, cudaMemcpyDeviceToDevice, st));
                        MN_CU(cudaMemcpyAsync(en.d_theta_best.p + (size_t)b * kE, en.d_theta.p + (size_t)b * kE,
                                              sizeof(int32_t) * kE, cudaMemcpyDeviceToDevice, st));
                        s.best = s.cur;
                        for (int i = 0; i < k; ++i) s.bestmw[i] = s.mw[i];
                    }
                    s.llold = s.ll;
                }
                s.nw = en.free_buf(s);
            }
            for (int b : ini) en.slots[b].nw = en.free_buf(en.slots[b]);
            // ---- gamma, M-step statistics ----
            MN_TRY(en.op_affinity(act));
            MN_TRY(en.op_mstats());
            // ---- structure ----
            if (!en.fixed_theta) MN_TRY(en.op_theta_init(ini));                   // fixed theta: the theta-given path of getProbs
            MN_TRY(en.op_search(em));
            // ---- E-step ----
            MN_TRY(en.op_estep());
            // ---- mixture weights and likelihood ----
            MN_TRY(en.tic());
            {
                std::vector<int> wcur, wnew;
                for (int b : em) { wcur.push_back(en.slots[b].cur); wnew.push_back(en.slots[b].nw); }
                MN_TRY(en.op_stats(em, wcur, false, true, 127, true));            // getProbs preamble on probsold, outer mw
                for (int t = 0; t < T; ++t) MN_TRY(en.op_stats(em, wnew, true, true, t, false));
                std::vector<int> inew;
                for (int b : ini) inew.push_back(en.slots[b].nw);
                // theta-free getProbs: one inner iteration per tick.  Fixed theta: the contraction is the same in all T inner
                // iterations (R/mnems_low.r:621-631), so the whole call is this one tick: T evaluations of (ll, mw) on it
                for (int t = 0; t < (en.fixed_theta ? T : 1); ++t) MN_TRY(en.op_stats(ini, inew, false, false, t, false));
            }
            MN_CU(cudaMemcpyAsync(en.h_ll, en.d_ll.p, sizeof(double) * B * 128, cudaMemcpyDeviceToHost, st));
            MN_CU(cudaMemcpyAsync(en.h_chg, en.d_edge.p, sizeof(int32_t) * en.M, cudaMemcpyDeviceToHost, st));
            MN_CU(cudaMemcpyAsync(en.h_chg + en.M, en.d_thch.p, sizeof(int32_t) * en.M, cudaMemcpyDeviceToHost, st));
            MN_TRY(en.toc(&en.stats.ms_mixture));
            MN_CU(cudaStreamSynchronize(st));                   // h_ll / h_chg are read below
            // ---- host logic, INIT slots: R/mnems_low.r:643-659 ----
            std::vector<int> outer, outer_guard;
            for (int b : ini) {
                Slot &s = en.slots[b];
                for (int t = 0; t < (en.fixed_theta ? T - 1 : 0); ++t) {          // fixed theta: inner iterations 0 .. T-2 of this tick
                    const double llt = en.h_ll[b * 128 + t];
                    if (en.tape_on) en.rec({6, s.start, s.inner, llt - s.llold_inner > 0 ? 1 : 0});
                    if (llt - s.llold_inner > 0) s.best = s.nw;
                    s.llold_inner = llt;
                    ++s.inner;
                }
                const double ll0 = en.h_ll[b * 128 + (en.fixed_theta ? T - 1 : 0)];
                if (en.tape_on) en.rec({6, s.start, s.inner, ll0 - s.llold_inner > 0 ? 1 : 0});
                if (ll0 - s.llold_inner > 0) s.best = s.nw;
                s.cur = s.nw;
                s.llold_inner = ll0;
                if (++s.inner >= T) {
                    // getProbs returns (bestprobs, mw, llold); do_inits :1912-1915, then the mw update of :1936-1940
                    s.cur = s.best;
                    s.probs_ll = s.llold_inner;
                    s.phase = 1;
                    s.time0 = true;
                    s.have_theta = en.fixed_theta;                                    // res1 carries subtopo from the start
                    outer_guard.push_back(b);
                }
            }
            // ---- host logic, EM slots: R/mnems.r:2104-2121 ----
            for (int b : em) {
                Slot &s = en.slots[b];
                bool improved = false;
                double lo = -INFINITY;
                for (int t = 0; t < T; ++t) { if (en.h_ll[b * 128 + t] - lo > 0) improved = true; lo = en.h_ll[b * 128 + t]; }
                const double newll = en.h_ll[b * 128 + T - 1];
                if (en.tape_on) en.rec({5, s.start, s.count, newll > s.probs_ll ? 1 : 0, improved ? 1 : 0});
                if (newll > s.probs_ll) {                                          // :2104-2106
                    if (improved) s.cur = s.nw;                                   // else getProbs returned its input
                    s.probs_ll = newll;
                }
                s.ll = s.probs_ll;                                                // :2111
                for (int i = 0; i < k; ++i) s.mwold[i] = s.mw[i];
                outer.push_back(b);
            }
            // outer mixture-weight update on the accepted probs, :2113-2121 (and :1936-1940 with the NA guard)
            {
                std::vector<int> w;
                for (int b : outer) w.push_back(en.slots[b].cur);
                MN_TRY(en.op_stats(outer, w, false, false, 126, false));
                w.clear();
                for (int b : outer_guard) w.push_back(en.slots[b].cur);
                MN_TRY(en.op_stats(outer_guard, w, false, false, 126, true));
            }
            MN_CU(cudaMemcpyAsync(en.h_mw, en.d_mw.p, sizeof(double) * B * 32, cudaMemcpyDeviceToHost, st));
            MN_CU(cudaStreamSynchronize(st));
            MN_TRY(en.flush_timers());
            for (int b : act)
                for (int i = 0; i < k; ++i) en.slots[b].mw[i] = en.h_mw[b * 32 + i];
            for (int b : outer_guard)
                for (int i = 0; i < k; ++i) en.slots[b].bestmw[i] = en.slots[b].mw[i];   // bestmw <- mw  :1949
            for (int b : em) {
                Slot &s = en.slots[b];
                double ec = 0, tc = 0;
                for (int i = 0; i < k; ++i) { ec += en.h_chg[b * k + i]; tc += en.h_chg[en.M + b * k + i]; }
                long double dm = 0;
                for (int i = 0; i < k; ++i) dm += fabs(s.mw[i] - s.mwold[i]);
                const int sidx = s.start;
                if (s.ll > s.maxll) s.maxll = s.ll;
                if (s.count < mt) {
                    if (lls) lls[(size_t)sidx * mt + s.count] = s.ll;
                    if (phievo) phievo[(size_t)sidx * mt + s.count] = ec;
                    if (thetaevo) thetaevo[(size_t)sidx * mt + s.count] = tc;
                    if (mwevo) mwevo[(size_t)sidx * mt + s.count] = (double)dm;
                }
                s.count++;
                if (ec == 0 && tc == 0 && !s.time0) s.stop = true;                 // :2138-2141
                s.time0 = false;
                s.have_theta = true;
                // loop condition :1955-1958 (converged = -Inf, increase = TRUE)
                const bool again = ((s.ll - s.llold > -INFINITY) && fabs(s.ll - s.llold) > 0) && !s.stop;
                const bool guard = opts->max_iter_guard > 0 && s.count >= opts->max_iter_guard;
                if (!again || guard) {
                    MN_TRY(finish_slot(b));
                    const int s1 = take_start();
                    if (s1 >= 0) MN_TRY(begin_slot(b, s1));
                    else drained = true;
                }
            }
            if (drained && running > 0) MN_TRY(en.compact());
            if (guard_ticks > 100000) return set_err(MNEMCU_ERR_ARG, "EM did not terminate");
        }
        return 0;
    };
    int rc = body();
    if (!rc && winner_probs && winner >= 0) {
        rc = launch_L_from_internal(cx, winner_L.p, k, en.Lstage.p, st);
        if (!rc && cudaMemcpyAsync(probs_out, en.Lstage.p, sizeof(double) * kC, cudaMemcpyDeviceToHost, st) != cudaSuccess)
            rc = set_err(MNEMCU_ERR_CUDA, "copy of the winner's probs failed");
    }
    if (!rc) rc = en.flush_timers();
    if (!rc) {
        cudaEventRecord(evB, st);
        cudaEventSynchronize(evB);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evA, evB);
        en.stats.ms_total = ms;
        cudaMemcpy(en.h_counters, en.d_counters.p, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost);
        en.stats.cand_scored = (int64_t)en.h_counters[0];
        en.stats.greedy_iters = (int64_t)en.h_counters[1];
        *best_out = winner;
        en.stats.search_sm_busy = en.sb ? search_sm_busy(en.sb) : 0.0;
        en.stats.tape_words = (int64_t)en.tape.size();
        if (en.tape_on) {
            if ((int64_t)en.tape.size() > opts->tape_cap) {
                if (stats_out) *stats_out = en.stats;        // tape_words = the capacity this fit needs
                rc = set_err(MNEMCU_ERR_ARG, "decision tape: %lld words needed, capacity %lld", (long long)en.tape.size(),
                             (long long)opts->tape_cap);
            } else std::copy(en.tape.begin(), en.tape.end(), opts->tape);
        }
        if (stats_out) *stats_out = en.stats;
    }
    cudaEventDestroy(evA); cudaEventDestroy(evB);
    return rc;
}

// ---------------------------------------------------------------------------------------------------------
// timing probes for bench.py
// ---------------------------------------------------------------------------------------------------------
extern "C" int mnemcu_bench_estep(mnemcu_ctx *cx, int batch, int k, const uint8_t *phi, const int32_t *theta, int n_rep,
                                  double *ms_out, double *bytes_per_launch)
{
    MN_ARG(cx && phi && theta && ms_out, "NULL pointer");
    MN_ARG(batch >= 1 && k >= 1 && batch * k <= 32 && n_rep >= 1, "bad batch / k / n_rep");
    Engine en;
    MN_TRY(en.init(cx, k, batch, 0, 2.0, false));
    for (int b = 0; b < batch; ++b) {
        MN_TRY(en.load_graphs(b, phi + (size_t)b * k * cx->S * cx->S));
        MN_CU(cudaMemcpyAsync(en.d_theta.p + (size_t)b * k * cx->E, theta + (size_t)b * k * cx->E,
                              sizeof(int32_t) * (size_t)k * cx->E, cudaMemcpyHostToDevice, en.st));
    }
    MN_CU(cudaStreamSynchronize(en.st));
    for (int r = 0; r < n_rep; ++r) {
        double ms = 0.0;
        en.stats.ms_estep = 0.0;
        MN_TRY(en.op_estep());
        MN_TRY(en.flush_timers());
        ms = en.stats.ms_estep;
        ms_out[r] = ms;
    }
    if (bytes_per_launch) *bytes_per_launch = 8.0 * cx->E * cx->C + 8.0 * (double)batch * k * cx->C;
    return 0;
}

extern "C" int mnemcu_bench_mstats(mnemcu_ctx *cx, int batch, int k, int n_rep, double *ms_out, double *bytes_per_launch)
{
    MN_ARG(cx && ms_out, "NULL pointer");
    MN_ARG(batch >= 1 && k >= 1 && batch * k <= 32 && n_rep >= 1, "bad batch / k / n_rep");
    Engine en;
    MN_TRY(en.init(cx, k, batch, 0, 2.0, false));
    for (int r = 0; r < n_rep; ++r) {
        en.stats.ms_mstats = 0.0;
        MN_TRY(en.op_mstats());
        MN_TRY(en.flush_timers());
        ms_out[r] = en.stats.ms_mstats;
    }
    if (bytes_per_launch)
        *bytes_per_launch = 8.0 * cx->E * cx->C + 8.0 * (double)batch * k * cx->C + 8.0 * cx->E * cx->S * (double)batch * k;
    return 0;
}
