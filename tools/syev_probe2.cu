// tools/syev_probe2.cu — what a faster small projected eigenproblem could be built from (m = 600, 900): cuSOLVER Xsyevd in
// double and in float, Dsyevdx over the whole spectrum, sytrd alone, and a 900^3 DGEMM (the unit of a refinement step).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/syev_probe2.cu -lcusolver -lcublas -o tools/syev_probe2
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#define CK(x) do { auto e = (x); if (e != 0) { printf("error %d at %s:%d\n", (int)e, __FILE__, __LINE__); exit(1); } } while (0)
int main() {
  cusolverDnHandle_t h; CK(cusolverDnCreate(&h));
  cublasHandle_t cb; CK(cublasCreate(&cb));
  cudaStream_t st; CK(cudaStreamCreate(&st)); CK(cusolverDnSetStream(h, st)); CK(cublasSetStream(cb, st));
  cusolverDnParams_t par; CK(cusolverDnCreateParams(&par));
  for (int n : {300, 600, 900}) {
    std::vector<double> A((size_t)n * n); std::vector<float> Af((size_t)n * n);
    srand(1);
    for (int j = 0; j < n; j++) for (int i = 0; i <= j; i++) { double v = rand() / (double)RAND_MAX - 0.5; if (i == j) v += 2.0 + i; A[i + (size_t)j * n] = A[j + (size_t)i * n] = v; }
    for (size_t i = 0; i < A.size(); i++) Af[i] = (float)A[i];
    double *dA, *dA0, *dW, *dB, *dC; float *fA, *fA0, *fW; int* dinfo;
    CK(cudaMalloc(&dA, 8 * n * n)); CK(cudaMalloc(&dA0, 8 * n * n)); CK(cudaMalloc(&dB, 8 * n * n)); CK(cudaMalloc(&dC, 8 * n * n)); CK(cudaMalloc(&dW, 8 * n));
    CK(cudaMalloc(&fA, 4 * n * n)); CK(cudaMalloc(&fA0, 4 * n * n)); CK(cudaMalloc(&fW, 4 * n)); CK(cudaMalloc(&dinfo, 4));
    CK(cudaMemcpy(dA0, A.data(), 8 * n * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(fA0, Af.data(), 4 * n * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, A.data(), 8 * n * n, cudaMemcpyHostToDevice));
    size_t wd = 0, wh = 0, wdf = 0, whf = 0;
    CK(cusolverDnXsyevd_bufferSize(h, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, &wd, &wh));
    CK(cusolverDnXsyevd_bufferSize(h, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_32F, fA, n, CUDA_R_32F, fW, CUDA_R_32F, &wdf, &whf));
    void *ws, *wsf; CK(cudaMalloc(&ws, wd + 16)); CK(cudaMalloc(&wsf, wdf + 16)); std::vector<char> hws(wh + 16), hwsf(whf + 16);
    int lwx = 0, meig = 0;
    CK(cusolverDnDsyevdx_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_ALL, CUBLAS_FILL_MODE_UPPER, n, dA, n, 0, 0, 1, n, &meig, dW, &lwx));
    double* wx; CK(cudaMalloc(&wx, 8 * (size_t)(lwx + 2)));
    int lwt = 0; double *dD, *dE, *dTau, *wt;
    CK(cudaMalloc(&dD, 8 * n)); CK(cudaMalloc(&dE, 8 * n)); CK(cudaMalloc(&dTau, 8 * n));
    CK(cusolverDnDsytrd_bufferSize(h, CUBLAS_FILL_MODE_UPPER, n, dA, n, dD, dE, dTau, &lwt)); CK(cudaMalloc(&wt, 8 * (size_t)(lwt + 2)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float t[5];
    for (int alg = 0; alg < 5; alg++) {
      float best = 1e30f;
      for (int rep = 0; rep < 6; rep++) {
        CK(cudaMemcpyAsync(dA, dA0, 8 * n * n, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(fA, fA0, 4 * n * n, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        cudaEventRecord(e0, st);
        if (alg == 0) CK(cusolverDnXsyevd(h, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_64F, dA, n, CUDA_R_64F, dW, CUDA_R_64F, ws, wd, hws.data(), wh, dinfo));
        else if (alg == 1) CK(cusolverDnXsyevd(h, par, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, n, CUDA_R_32F, fA, n, CUDA_R_32F, fW, CUDA_R_32F, wsf, wdf, hwsf.data(), whf, dinfo));
        else if (alg == 2) CK(cusolverDnDsyevdx(h, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_ALL, CUBLAS_FILL_MODE_UPPER, n, dA, n, 0, 0, 1, n, &meig, dW, wx, lwx, dinfo));
        else if (alg == 3) CK(cusolverDnDsytrd(h, CUBLAS_FILL_MODE_UPPER, n, dA, n, dD, dE, dTau, wt, lwt, dinfo));
        else { const double one = 1, zero = 0; CK(cublasDgemm(cb, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, dA, n, dB, n, &zero, dC, n)); }
        cudaEventRecord(e1, st); CK(cudaStreamSynchronize(st));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
      }
      t[alg] = best;
    }
    printf("{\"n\": %d, \"dsyevd_ms\": %.3f, \"ssyevd_ms\": %.3f, \"dsyevdx_all_ms\": %.3f, \"dsytrd_ms\": %.3f, \"dgemm_tn_ms\": %.3f}\n", n, t[0], t[1], t[2], t[3], t[4]);
  }
  return 0;
}
