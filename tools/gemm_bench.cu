// stand-alone timing of the DMMA update kernel with experiment switches (development aid)
#define FGP_GEMM_DEBUG 1
#include "../fungp_b200/csrc/gemm_dmma.cu"      // the TMA kernel (gemm_tma.cu) is linked as its own translation unit
#include <cstdlib>
void fgp_set_err(fgp_ctx*, const std::string&) {}
void fgp_set_err_cuda(fgp_ctx*, const char*, cudaError_t, const char*, int) {}
std::atomic<long long> g_fgp_launches{0};
int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 48;   // trailing tiles
    const int n = (T + 4) * 128;
    const long ld = n + 16;
    double* M;
    cudaMalloc(&M, sizeof(double) * ld * n);
    cudaMemset(M, 0, sizeof(double) * ld * n);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // operand path: cp.async kernel with its experiment switches, then the TMA kernel (dbg column = -1)
    for (int K : {128, 256, 512, 1024, 2048}) for (int dbg : {0, 32, 1, 3, 16, -1}) {
        GemmArgs g{};
        g.tma = dbg < 0;
        g.A = M; g.B = M; g.C = M; g.lda = g.ldb = g.ldc = ld; g.K = K; g.accumulate = 1;
        g.tri0 = 4; g.triT = T; g.rowLo = n; g.rowHoleHi = n; g.nrows = n; g.colLo = n; g.colHoleHi = n; g.ncols = n;
        g.batch = 1; g.dbg = dbg < 0 ? 0 : dbg;
        launch_gemm(g, 0);
        float best = 1e9;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0); launch_gemm(g, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double tiles = T * (T + 1) / 2.0;
        printf("T=%d K=%d dbg=%2d: %.3f ms  %.2f TF  (%s)\n", T, K, dbg, best, tiles * 2.0 * 128 * 128 * K / (best * 1e-3) / 1e12,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
