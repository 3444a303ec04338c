// Compiles and RUNS rshim/mnemcu_shim.cpp (the Rcpp glue of INTEGRATION.md) against the functional Rcpp stand-in in
// rcpp_stub/Rcpp.h: a driver in the role of the R code in rshim/mnem_cuda.R.  Input: the same binary file cabi_check
// reads (header, D, pert, init graphs); output: the line cabi_check prints for the same fit, so that the test can
// compare the two consumers of the C ABI and the Python host mirror.
//   shim_check syntax          -> "ok" (the shim compiled and linked: no device needed)
//   shim_check fit <file>      -> best / ll / iters / digest through mnemcu_context + mnemcu_em of the shim
#include "../../rshim/mnemcu_shim.cpp"

#include <cstring>

static int fit(const char *path)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) return 2;
    int32_t h[6];
    if (std::fread(h, sizeof(int32_t), 6, f) != 6) return 2;
    const int E = h[0], C = h[1], S = h[2], k = h[3], starts = h[4], complete = h[5];
    NumericMatrix D(E, C);
    IntegerVector pert(C);
    std::vector<uint8_t> init((size_t)starts * k * S * S);
    if (std::fread(D.begin(), sizeof(double), (size_t)E * C, f) != (size_t)E * C) return 2;
    if (std::fread(pert.begin(), sizeof(int32_t), C, f) != (size_t)C) return 2;
    if (std::fread(init.data(), 1, init.size(), f) != init.size()) return 2;
    std::fclose(f);
    try {
        SEXP ctx = mnemcu_context(D, pert, S, 0);
        List phi(starts);                                   // mnem()'s phi argument: list over starts of lists of matrices
        for (int s = 0; s < starts; ++s) {
            List comps(k);
            for (int i = 0; i < k; ++i) {
                NumericMatrix A(S, S);
                for (int q = 0; q < S * S; ++q) A[q] = init[((size_t)s * k + i) * S * S + q];
                comps[i] = A;
            }
            phi[s] = comps;
        }
        List out = mnemcu_em(ctx, phi, complete != 0, 2.0, 0, 64, 0, R_NilValue, false, R_NilValue);
        const int best = IntegerVector(out["best"])[0] - 1;
        List limits(out["limits"]);
        List lim(limits[best]);
        List res(lim["res"]);
        long iters = 0;
        for (int s = 0; s < starts; ++s) iters += NumericVector(List(limits[s])["lls"]).size();
        uint64_t dig = 1469598103934665603ull;              // FNV-1a over the winner's graphs and theta (as cabi_check)
        for (int i = 0; i < k; ++i) {
            NumericMatrix A(List(res[i])["adj"]);
            for (int q = 0; q < S * S; ++q) { dig ^= (uint64_t)(A[q] != 0); dig *= 1099511628211ull; }
        }
        for (int i = 0; i < k; ++i) {
            IntegerVector t(List(res[i])["subtopo"]);
            for (int e = 0; e < E; ++e) { dig ^= (uint64_t)(uint32_t)t[e]; dig *= 1099511628211ull; }
        }
        std::printf("best %d ll %.17g iters %ld digest %llu\n", best, NumericVector(lim["ll"])[0], iters, (unsigned long long)dig);
        // the legacy symbols through the shim: closure of a chain, and a scoreAdj round trip
        NumericMatrix g(3, 3);
        g[0 + 1 * 3] = 1; g[1 + 2 * 3] = 1;                 // 1 -> 2 -> 3
        NumericMatrix c(transClose_W(g));
        std::printf("closure_1_3 %d\n", (int)(c[0 + 2 * 3] != 0));
        rstub_finalize(ctx);                                // R's gc: the finalizer releases the device context
        std::printf("ctx_released %d\n", (int)(R_ExternalPtrAddr(ctx) == nullptr));
    } catch (const std::exception &ex) {
        std::printf("error %s\n", ex.what());
        return 1;
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !std::strcmp(argv[1], "syntax")) { std::printf("ok\n"); return 0; }
    if (argc >= 3 && !std::strcmp(argv[1], "fit")) return fit(argv[2]);
    std::fprintf(stderr, "usage: shim_check syntax | fit <file>\n");
    return 2;
}
