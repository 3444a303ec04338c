// Library bookkeeping and the small native routines of the reference's src/mm.cpp as batched CUDA kernels.
#include "common.cuh"

namespace mn {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// one warp per graph; op 0 transClose_W, 1 transClose_Ins, 2 transClose_Del, 3 transitive.reduction
__global__ void graph_op_kernel(int S, int op, const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                const int32_t *__restrict__ u, const int32_t *__restrict__ v)
{
    const int g = blockIdx.x, lane = threadIdx.x;
    uint32_t row = load_row_mask(in + (size_t)g * S * S, lane, S);
    if (op == 0) {
        row = warp_closure(row, S);                                  // src/mm.cpp:15-30 (no diagonal forcing)
    } else if (op == 1) {
        const int uu = u[g] - 1, vv = v[g] - 1;                      // src/mm.cpp:50-65: x[i,j] |= x[i,u] && x[v,j]
        const uint32_t rv = __shfl_sync(kFull, row, vv);
        if ((row >> uu) & 1u) row |= rv;
    } else if (op == 2) {
        const int uu = u[g] - 1, vv = v[g] - 1;                      // src/mm.cpp:32-48: x[u,v] |= any_k x[u,k] && x[k,v]
        const uint32_t ru = __shfl_sync(kFull, row, uu);
        const unsigned any = __ballot_sync(kFull, lane < S && ((ru >> lane) & 1u) && ((row >> vv) & 1u));
        if (lane == uu && any) row |= 1u << vv;
    } else {
        row = warp_reduce_graph(row, lane, S);                       // R/mnems.r:511-533
    }
    store_row_mask(out + (size_t)g * S * S, row, lane, S);
}

static int graph_op(int op, const uint8_t *x, int S, int n, const int32_t *u, const int32_t *v, uint8_t *out)
{
    MN_ARG(x && out, "NULL pointer");
    MN_ARG(S >= 1 && S <= MNEMCU_MAX_S && n >= 1, "S must be in 1..32 and n positive");
    if (op == 1 || op == 2) {
        MN_ARG(u && v, "u / v are NULL");
        for (int g = 0; g < n; ++g) MN_ARG(u[g] >= 1 && u[g] <= S && v[g] >= 1 && v[g] <= S, "u, v must be in 1..S");
    }
    DevBuf<uint8_t> din, dout;
    DevBuf<int32_t> du, dv;
    MN_TRY(din.alloc((size_t)n * S * S));
    MN_TRY(dout.alloc((size_t)n * S * S));
    MN_CU(cudaMemcpy(din.p, x, (size_t)n * S * S, cudaMemcpyHostToDevice));
    if (u) {
        MN_TRY(du.alloc(n)); MN_TRY(dv.alloc(n));
        MN_CU(cudaMemcpy(du.p, u, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
        MN_CU(cudaMemcpy(dv.p, v, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    }
    MN_LAUNCH(graph_op_kernel, n, 32, 0, 0, S, op, din.p, dout.p, du.p, dv.p);
    MN_CU(cudaMemcpy(out, dout.p, (size_t)n * S * S, cudaMemcpyDeviceToHost));
    return 0;
}

__global__ void max_col_row_kernel(const double *__restrict__ x, int nr, int nc, int32_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr) return;
    double best = -INFINITY;
    int arg = 0;
    for (int j = 0; j < nc; ++j) {
        const double v = x[i + (size_t)j * nr];
        if (v > best) { best = v; arg = j + 1; }
    }
    out[i] = arg;
}

// C = A * B with one accumulator per output and ascending-p mul/add (no FMA): the order Eigen's GEBP uses for
// the depths that occur on this path (src/mm.cpp:8-13).
__global__ void matmul_kernel(const double *__restrict__ A, const double *__restrict__ B, int m, int p, int n,
                              double *__restrict__ Cm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= m) return;
    double s = 0.0;
    for (int l = 0; l < p; ++l) s = __dadd_rn(s, __dmul_rn(A[i + (size_t)l * m], B[l + (size_t)j * p]));
    Cm[i + (size_t)j * m] = s;
}

}  // namespace mn

using namespace mn;

extern "C" int mnemcu_version(void) { return 200; }   // 2.0: em_opts gained tape, probs_mode, mw
extern "C" const char *mnemcu_last_error(void) { return g_err; }
extern "C" int64_t mnemcu_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" int mnemcu_device_count(int *n)
{
    MN_ARG(n != nullptr, "n is NULL");
    *n = 0;
    MN_CU(cudaGetDeviceCount(n));
    return 0;
}

extern "C" int mnemcu_trans_close_w(uint8_t *x, int S, int n) { return graph_op(0, x, S, n, nullptr, nullptr, x); }
extern "C" int mnemcu_trans_close_ins(uint8_t *x, int S, int n, const int32_t *u, const int32_t *v)
{
    return graph_op(1, x, S, n, u, v, x);
}
extern "C" int mnemcu_trans_close_del(uint8_t *x, int S, int n, const int32_t *u, const int32_t *v)
{
    return graph_op(2, x, S, n, u, v, x);
}
extern "C" int mnemcu_trans_reduce(const uint8_t *x, int S, int n, uint8_t *out)
{
    return graph_op(3, x, S, n, nullptr, nullptr, out);
}

extern "C" int mnemcu_max_col_row(const double *x, int nr, int nc, int32_t *out)
{
    MN_ARG(x && out, "NULL pointer");
    MN_ARG(nr >= 1 && nc >= 1, "empty matrix");
    DevBuf<double> dx;
    DevBuf<int32_t> dout;
    MN_TRY(dx.alloc((size_t)nr * nc));
    MN_TRY(dout.alloc(nr));
    MN_CU(cudaMemcpy(dx.p, x, sizeof(double) * (size_t)nr * nc, cudaMemcpyHostToDevice));
    MN_LAUNCH(max_col_row_kernel, (nr + 127) / 128, 128, 0, 0, dx.p, nr, nc, dout.p);
    MN_CU(cudaMemcpy(out, dout.p, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int mnemcu_matmul(const double *A, const double *B, int m, int p, int n, double *Cm)
{
    MN_ARG(A && B && Cm, "NULL pointer");
    MN_ARG(m >= 1 && p >= 1 && n >= 1 && n <= 65535, "bad matrix size");
    DevBuf<double> dA, dB, dC;
    MN_TRY(dA.alloc((size_t)m * p)); MN_TRY(dB.alloc((size_t)p * n)); MN_TRY(dC.alloc((size_t)m * n));
    MN_CU(cudaMemcpy(dA.p, A, sizeof(double) * (size_t)m * p, cudaMemcpyHostToDevice));
    MN_CU(cudaMemcpy(dB.p, B, sizeof(double) * (size_t)p * n, cudaMemcpyHostToDevice));
    dim3 grid((m + 127) / 128, n);
    MN_LAUNCH(matmul_kernel, grid, 128, 0, 0, dA.p, dB.p, m, p, n, dC.p);
    MN_CU(cudaMemcpy(Cm, dC.p, sizeof(double) * (size_t)m * n, cudaMemcpyDeviceToHost));
    return 0;
}
