// tools/syev_probe.cu — latency of the symmetric eigen-solvers cuSOLVER offers for the small projected problem
// (m <= 3k): Xsyevd (divide & conquer, what csrc/smalldense.cu uses), Dsyevj (Jacobi), XsyevBatched (batch = 1).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/syev_probe.cu -lcusolver -lcublas -o tools/syev_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#define CK(x) do { auto e = (x); if (e != 0) { printf("error %d at %s:%d\n", (int)e, __FILE__, __LINE__); exit(1); } } while (0)
int main() {
  cusolverDnHandle_t h; CK(cusolverDnCreate(&h));
  cudaStream_t st; CK(cudaStreamCreate(&st)); CK(cusolverDnSetStream(h, st));
  cusolverDnParams_t par; CK(cusolverDnCreateParams(&par));
  syevjInfo_t jinfo; CK(cusolverDnCreateSyevjInfo(&jinfo));
  CK(cusolverDnXsyevjSetTolerance(jinfo, 1e-14)); CK(cusolverDnXsyevjSetMaxSweeps(jinfo, 30));
  const int sizes[] = {20, 40, 60, 96, 128, 192, 256, 300, 384, 450, 600, 900};
  for (int n : sizes) {
    std::vector<double> A((size_t)n * n);
    srand(1);
    for (int j = 0; j < n; j++) for (int i = 0; i <= j; i++) { double v = rand() / (double)RAND_MAX - 0.5; if (i == j) v += 2.0 + i; A[i + (size_t)j * n] = A[j + (size_t)i * n] = v; }
    double *dA, *dA0, *dW; int* dinfo;
    CK(cudaMalloc(&dA, sizeof(double) * n * n)); CK(cudaMalloc(&dA0, sizeof(double) * n * n)); CK(cudaMalloc(&dW, sizeof(double) * n)); CK(cudaMalloc(&dinfo, 4));
    CK(cudaMemcpy(dA0, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    size_t wd = 0, wh = 0;
    CK(cusolverDnXsyevd_bufferSize(h, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, &wd, &wh));
    void* ws; CK(cudaMalloc(&ws, wd + 16)); std::vector<char> hws(wh + 16);
    int lw = 0; CK(cusolverDnDsyevj_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, dA, n, dW, &lw, jinfo));
    double* wj; CK(cudaMalloc(&wj, sizeof(double) * (lw + 2)));
    size_t bd = 0, bh = 0;
    cusolverStatus_t bst = cusolverDnXsyevBatched_bufferSize(h, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, &bd, &bh, 1);
    void* bws = nullptr; std::vector<char> bhws(bh + 16);
    if (bst == 0) CK(cudaMalloc(&bws, bd + 16));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float t[3] = {0, 0, 0}; double w0[3] = {0, 0, 0};
    for (int alg = 0; alg < 3; alg++) {
      if (alg == 2 && bst != 0) { t[2] = -1; continue; }
      float best = 1e30f;
      for (int rep = 0; rep < 6; rep++) {
        CK(cudaMemcpyAsync(dA, dA0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        cudaEventRecord(e0, st);
        if (alg == 0) CK(cusolverDnXsyevd(h, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, ws, wd, hws.data(), wh, dinfo));
        else if (alg == 1) CK(cusolverDnDsyevj(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, dA, n, dW, wj, lw, dinfo, jinfo));
        else CK(cusolverDnXsyevBatched(h, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, bws, bd, bhws.data(), bh, dinfo, 1));
        cudaEventRecord(e1, st); CK(cudaStreamSynchronize(st));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
      }
      t[alg] = best;
      CK(cudaMemcpy(&w0[alg], dW, 8, cudaMemcpyDeviceToHost));
    }
    printf("{\"n\": %d, \"syevd_ms\": %.3f, \"syevj_ms\": %.3f, \"syev_batched_ms\": %.3f, \"w0\": [%.15g, %.15g, %.15g]}\n", n, t[0], t[1], t[2], w0[0], w0[1], w0[2]);
    cudaFree(dA); cudaFree(dA0); cudaFree(dW); cudaFree(dinfo); cudaFree(ws); cudaFree(wj); if (bws) cudaFree(bws);
  }
  return 0;
}
