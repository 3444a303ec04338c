// A compiled C++ consumer of the C ABI (include/mnemcu.h), passing column-major buffers exactly as the Rcpp shim would.
// Built with plain g++ against libmnemcu.so (no CUDA headers needed): proves the boundary is a C ABI, and lets the tests
// compare struct layouts and results with the ctypes binding.
//
//   cabi_check layout                 print sizeof / offsetof of the ABI structs (compared with ctypes in test_abi.py)
//   cabi_check nogpu                  without a device every compute entry must fail with MNEMCU_ERR_CUDA and a message
//   cabi_check fit <file>             read a problem written by the test (header of int32: E C S k starts complete, then
//                                     D (E*C doubles, column-major), pert (C int32), init graphs (starts*k*S*S bytes)),
//                                     run mnemcu_em_run and print "best ll n_iter..." plus a digest of the winner's graphs
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/mnemcu.h"

static int layout()
{
    std::printf("em_opts %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(mnemcu_em_opts), offsetof(mnemcu_em_opts, complete),
                offsetof(mnemcu_em_opts, logtype), offsetof(mnemcu_em_opts, batch), offsetof(mnemcu_em_opts, max_trace),
                offsetof(mnemcu_em_opts, max_iter_guard), offsetof(mnemcu_em_opts, next_start),
                offsetof(mnemcu_em_opts, next_start_user), offsetof(mnemcu_em_opts, affinity), offsetof(mnemcu_em_opts, tape),
                offsetof(mnemcu_em_opts, tape_cap), offsetof(mnemcu_em_opts, probs_mode), offsetof(mnemcu_em_opts, mw), offsetof(mnemcu_em_opts, nullcomp), offsetof(mnemcu_em_opts, theta));
    std::printf("em_stats %zu %zu %zu %zu %zu\n", sizeof(mnemcu_em_stats), offsetof(mnemcu_em_stats, em_iters),
                offsetof(mnemcu_em_stats, ms_total), offsetof(mnemcu_em_stats, estep_bytes), offsetof(mnemcu_em_stats, tape_words));
    std::printf("version %d\n", mnemcu_version());
    return 0;
}

static int nogpu()
{
    int n = -1;
    int rc = mnemcu_device_count(&n);
    if (rc == 0 && n > 0) { std::printf("has_gpu %d\n", n); return 0; }
    const int E = 3, C = 8, S = 2;
    std::vector<double> D((size_t)E * C, 0.5);
    std::vector<int32_t> pert(C);
    for (int c = 0; c < C; ++c) pert[c] = 1 + c % S;
    mnemcu_ctx *cx = nullptr;
    rc = mnemcu_ctx_create(D.data(), E, C, pert.data(), S, 0, &cx);
    const char *msg = mnemcu_last_error();
    std::printf("ctx_create rc=%d msg=%s\n", rc, msg ? msg : "(null)");
    if (rc != MNEMCU_ERR_CUDA || cx != nullptr || !msg || !msg[0]) return 1;
    double ll = 0;
    rc = mnemcu_get_ll(D.data(), 1, E * C, nullptr, 0, 2.0, &ll);
    std::printf("get_ll rc=%d\n", rc);
    if (rc != MNEMCU_ERR_CUDA) return 1;
    rc = mnemcu_ctx_create(D.data(), E, C, nullptr, S, 0, &cx);          // argument errors are reported before any CUDA call
    std::printf("null_pert rc=%d\n", rc);
    return rc == MNEMCU_ERR_ARG ? 0 : 1;
}

static int fit(const char *path)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) { std::perror("open"); return 2; }
    int32_t h[6];
    if (std::fread(h, sizeof(int32_t), 6, f) != 6) return 2;
    const int E = h[0], C = h[1], S = h[2], k = h[3], starts = h[4], complete = h[5];
    std::vector<double> D((size_t)E * C);
    std::vector<int32_t> pert(C);
    std::vector<uint8_t> init((size_t)starts * k * S * S);
    if (std::fread(D.data(), sizeof(double), D.size(), f) != D.size()) return 2;
    if (std::fread(pert.data(), sizeof(int32_t), pert.size(), f) != pert.size()) return 2;
    if (std::fread(init.data(), 1, init.size(), f) != init.size()) return 2;
    std::fclose(f);
    mnemcu_ctx *cx = nullptr;
    if (mnemcu_ctx_create(D.data(), E, C, pert.data(), S, 0, &cx)) { std::printf("error %s\n", mnemcu_last_error()); return 1; }
    mnemcu_em_opts o;
    std::memset(&o, 0, sizeof(o));
    o.complete = complete; o.logtype = 2.0; o.batch = 0; o.max_trace = 64; o.max_iter_guard = 0;
    std::vector<uint8_t> adj((size_t)starts * k * S * S);
    std::vector<int32_t> theta((size_t)starts * k * E), nit(starts);
    std::vector<double> mw((size_t)starts * k), ll(starts), lls((size_t)starts * 64), pe((size_t)starts * 64),
        te((size_t)starts * 64), me((size_t)starts * 64);
    int32_t best = -1;
    mnemcu_em_stats st;
    int rc = mnemcu_em_run(cx, starts, k, init.data(), &o, adj.data(), theta.data(), nullptr, mw.data(), ll.data(), nit.data(),
                           lls.data(), pe.data(), te.data(), me.data(), &best, &st);
    if (rc) { std::printf("error %s\n", mnemcu_last_error()); mnemcu_ctx_destroy(cx); return 1; }
    uint64_t dig = 1469598103934665603ull;                             // FNV-1a over the winner's graphs and theta
    for (size_t i = 0; i < (size_t)k * S * S; ++i) { dig ^= adj[(size_t)best * k * S * S + i]; dig *= 1099511628211ull; }
    for (size_t i = 0; i < (size_t)k * E; ++i) { dig ^= (uint64_t)(uint32_t)theta[(size_t)best * k * E + i]; dig *= 1099511628211ull; }
    std::printf("best %d ll %.17g iters %lld digest %llu\n", best, ll[best], (long long)st.em_iters, (unsigned long long)dig);
    mnemcu_ctx_destroy(cx);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !std::strcmp(argv[1], "layout")) return layout();
    if (argc >= 2 && !std::strcmp(argv[1], "nogpu")) return nogpu();
    if (argc >= 3 && !std::strcmp(argv[1], "fit")) return fit(argv[2]);
    std::fprintf(stderr, "usage: cabi_check layout | nogpu | fit <file>\n");
    return 2;
}
