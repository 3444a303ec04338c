/* simData-style synthetic input generator for the CPU arm -- TEST / BENCH INFRASTRUCTURE.
 *
 * A plain-C twin of the device generator of the product library (syn_value in mnem_b200/csrc/ctx.cu; NumPy twin:
 * mnem_b200/simdata.py::host_data), so that bench.py --impl reference and the oracle-side tools can build the same
 * matrix WITHOUT loading libmnemcu.so.  The generator restates simData of the reference (R/mnems.r:2989-3160,
 * README.md:36-40) with a counter-based noise:
 *   D[e, c] = (bin - 0.5)/0.5 + noise(seed, e + E*c)
 *   bin     = closure(Nem_i)[pert(c), theta_i(e)] for the component i of cell c; uninformative rows: a fair coin
 *   noise   = (sum of twelve 16-bit uniforms - 6*65535)/65536 + (53-bit uniform - 0.5)/65536
 * Bit-identical to the other two twins (tests/test_simdata.py). */
#include <stddef.h>
#include <stdint.h>

static uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static double syn_noise(uint64_t seed, uint64_t idx)
{
    uint32_t sum = 0;
    for (int j = 0; j < 3; ++j) {
        uint64_t w = mix64(seed + 0x9E3779B97F4A7C15ULL * (3 * idx + j + 1));
        sum += (uint32_t)(w & 0xffff) + (uint32_t)((w >> 16) & 0xffff) + (uint32_t)((w >> 32) & 0xffff) + (uint32_t)(w >> 48);
    }
    const uint64_t w4 = mix64((seed ^ 0xA0761D6478BD642FULL) + 0xE7037ED1A0B428DBULL * (idx + 1));
    const double coarse = ((double)sum - 393210.0) * (1.0 / 65536.0);
    const double fine = ((double)(w4 >> 11) * (1.0 / 9007199254740992.0) - 0.5) * (1.0 / 65536.0);
    return coarse + fine;
}

/* cells [c0, c0 + nc) of D, column-major E x nc; pert 1-based S-gene of every cell, comp its component */
void orc_simgen_fill(double *D, int E, int64_t c0, int64_t nc, const int32_t *pert, int S, const uint8_t *comp,
                     const uint8_t *phi_true, const int32_t *theta_true, uint64_t seed, int nthreads)
{
    int64_t c;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (c = c0; c < c0 + nc; ++c) {
        double *col = D + (size_t)(c - c0) * E;
        const int cm = comp[c], p1 = pert[c];
        for (int e = 0; e < E; ++e) {
            const uint64_t idx = (uint64_t)e + (uint64_t)E * (uint64_t)c;
            const int t = theta_true[(size_t)cm * E + e];
            int bin;
            if (t > 0) bin = phi_true[(size_t)cm * S * S + (p1 - 1) + (size_t)(t - 1) * S] != 0;
            else bin = (int)(mix64(seed ^ 0xD1B54A32D192ED03ULL ^ (idx * 0x2545F4914F6CDD1DULL)) & 1);
            col[e] = (bin ? 1.0 : -1.0) + syn_noise(seed, idx);
        }
    }
}
