// stand-alone timing of the diagonal-block kernel with per-phase clock64 stamps (development aid)
#define FGP_POTF2_PROF 1
#include "../fungp_b200/csrc/potf2.cu"
#include <cstdlib>
#include <vector>
void fgp_set_err(fgp_ctx*, const std::string&) {}
void fgp_set_err_cuda(fgp_ctx*, const char*, cudaError_t, const char*, int) {}
std::atomic<long long> g_fgp_launches{0};
int main(int argc, char** argv) {
    const int batch = argc > 1 ? atoi(argv[1]) : 1;
    const int n = 128; const long ld = 144;
    std::vector<double> h((size_t)ld * n * batch, 0.0);
    for (int b = 0; b < batch; ++b)
        for (int j = 0; j < n; ++j)
            for (int i = j; i < n; ++i) h[(size_t)b * ld * n + i + j * ld] = (i == j) ? 2.0 + 0.01 * i : 1.0 / (1.0 + (i - j) * (i - j));
    double *M, *W; int* info;
    cudaMalloc(&M, sizeof(double) * h.size()); cudaMalloc(&W, sizeof(double) * 128 * 128 * batch); cudaMalloc(&info, 4 * batch);
    cudaMemset(info, 0, 4 * batch);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaMemcpy(M, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        launch_potf2(M, ld, ld * n, batch, 0, 128, W, 128 * 128, info, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    long long p[32];
    cudaMemcpyFromSymbol(p, g_potf2_prof, sizeof(p));
    printf("batch %d: %.2f us  (%s)\n", batch, best * 1e3, cudaGetErrorString(cudaGetLastError()));
    printf("load %lld\n", p[1] - p[0]);
    for (int s = 0; s < 4; ++s)
        printf("s=%d: (c of s-1 | a) %lld  (b) %lld  cycles  [%lld]\n", s, p[2 + 3 * s] - p[1 + 3 * s], p[3 + 3 * s] - p[2 + 3 * s],
               (s < 3 ? p[4 + 3 * s] : p[13]) - p[3 + 3 * s]);
    printf("store W %lld  store L %lld  total %lld cycles\n", p[14] - p[13], p[15] - p[14], p[15] - p[0]);      // W goes first (potf2_body.cuh)
    std::vector<double> o(h.size());
    cudaMemcpy(o.data(), M, sizeof(double) * h.size(), cudaMemcpyDeviceToHost);
    // residual check of the first matrix
    double err = 0;
    for (int j = 0; j < n; ++j) for (int i = j; i < n; ++i) {
        double s = 0; for (int k = 0; k <= j; ++k) s += o[i + k * ld] * o[j + k * ld];
        err = fmax(err, fabs(s - h[i + j * ld]));
    }
    printf("max |LL' - A| = %.3e\n", err);
    return 0;
}
