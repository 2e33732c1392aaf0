// Probe: FP64 peaks on B200 — DMMA.8x8x4 issue rate, DFMA issue rate, cuBLAS DGEMM 8192^3,
// cuBLAS tall-skinny DGEMM-TN / DSYRK at the LOBPCG Gram shapes. Output: JSON lines.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak_probe fp64_peak_probe.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int ILP>
__global__ void __launch_bounds__(256) dmma_rate(double* out, int iters){
  double c[ILP][2]; double a=1.0+threadIdx.x*1e-9, b=1.0-threadIdx.x*1e-9;
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=i;c[i][1]=-i;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]),"+d"(c[i][1]) : "d"(a),"d"(b));
  }
  double s=0; 
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int ILP>
__global__ void __launch_bounds__(256) dfma_rate(double* out, int iters){
  double c[ILP]; double a=1.0+threadIdx.x*1e-9, b=1e-9*threadIdx.x;
  #pragma unroll
  for(int i=0;i<ILP;i++) c[i]=i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) c[i]=fma(c[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
static float time_ms(cudaEvent_t a, cudaEvent_t b){float ms; cudaEventElapsedTime(&ms,a,b); return ms;}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount; printf("{\"gpu\":\"%s\",\"sms\":%d}\n",p.name,sms);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*256));
  for(int bps: {1,2,4,8}){
    int iters=20000;
    dmma_rate<8><<<sms*bps,256>>>(out,100); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); dmma_rate<8><<<sms*bps,256>>>(out,iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms=time_ms(e0,e1); double fl=(double)sms*bps*8/*warps*/*iters*8*512.0;
    printf("{\"probe\":\"dmma_8x8x4\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n",bps,ms,fl/ms/1e9);
    dfma_rate<8><<<sms*bps,256>>>(out,100); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); dfma_rate<8><<<sms*bps,256>>>(out,iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    ms=time_ms(e0,e1); fl=(double)sms*bps*256*(double)iters*8*2.0;
    printf("{\"probe\":\"dfma\",\"ctas_per_sm\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n",bps,ms,fl/ms/1e9);
  }
  cublasHandle_t h; cublasCreate(&h);
  // square DGEMM
  {
    int N=8192; double *A,*B,*C; CK(cudaMalloc(&A,8ull*N*N)); CK(cudaMalloc(&B,8ull*N*N)); CK(cudaMalloc(&C,8ull*N*N));
    CK(cudaMemset(A,0,8ull*N*N)); CK(cudaMemset(B,0,8ull*N*N));
    double al=1,be=0; 
    for(int i=0;i<2;i++) cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,N,N,N,&al,A,N,B,N,&be,C,N);
    CK(cudaDeviceSynchronize()); float best=1e9;
    for(int r=0;r<5;r++){cudaEventRecord(e0); cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,N,N,N,&al,A,N,B,N,&be,C,N); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); float ms=time_ms(e0,e1); if(ms<best)best=ms;}
    printf("{\"probe\":\"cublas_dgemm_8192\",\"ms\":%.3f,\"tflops\":%.2f}\n",best,2.0*N*N*N/best/1e9);
    // sustained: 3 seconds
    int reps=(int)(3000.0/best)+1; cudaEventRecord(e0); for(int r=0;r<reps;r++) cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,N,N,N,&al,A,N,B,N,&be,C,N); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms=time_ms(e0,e1)/reps; printf("{\"probe\":\"cublas_dgemm_8192_sustained\",\"ms\":%.3f,\"tflops\":%.2f,\"reps\":%d}\n",ms,2.0*N*N*N/ms/1e9,reps);
    cudaFree(A);cudaFree(B);cudaFree(C);
  }
  // tall skinny
  struct Sh{long n; int m;}; Sh shapes[]={{2097152,384},{4096000,600},{4096000,900},{1000000,60},{1000000,192}};
  for(auto s: shapes){
    double *S,*AS,*G; size_t bytes=8ull*s.n*s.m; if(bytes*2>100ull<<30) continue;
    CK(cudaMalloc(&S,bytes)); CK(cudaMalloc(&AS,bytes)); CK(cudaMalloc(&G,8ull*s.m*s.m));
    CK(cudaMemset(S,0,bytes)); CK(cudaMemset(AS,0,bytes));
    double al=1,be=0;
    for(int i=0;i<2;i++) cublasDgemm(h,CUBLAS_OP_T,CUBLAS_OP_N,s.m,s.m,(int)s.n,&al,S,(int)s.n,AS,(int)s.n,&be,G,s.m);
    CK(cudaDeviceSynchronize()); float best=1e9;
    for(int r=0;r<3;r++){cudaEventRecord(e0); cublasDgemm(h,CUBLAS_OP_T,CUBLAS_OP_N,s.m,s.m,(int)s.n,&al,S,(int)s.n,AS,(int)s.n,&be,G,s.m); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); float ms=time_ms(e0,e1); if(ms<best)best=ms;}
    printf("{\"probe\":\"cublas_dgemm_tn\",\"n\":%ld,\"m\":%d,\"ms\":%.3f,\"tflops\":%.2f,\"gbs\":%.1f}\n",s.n,s.m,best,2.0*s.n*s.m*s.m/best/1e9, 2.0*bytes/best/1e6);
    for(int i=0;i<2;i++) cublasDsyrk(h,CUBLAS_FILL_MODE_UPPER,CUBLAS_OP_T,s.m,(int)s.n,&al,S,(int)s.n,&be,G,s.m);
    CK(cudaDeviceSynchronize()); best=1e9;
    for(int r=0;r<3;r++){cudaEventRecord(e0); cublasDsyrk(h,CUBLAS_FILL_MODE_UPPER,CUBLAS_OP_T,s.m,(int)s.n,&al,S,(int)s.n,&be,G,s.m); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); float ms=time_ms(e0,e1); if(ms<best)best=ms;}
    printf("{\"probe\":\"cublas_dsyrk\",\"n\":%ld,\"m\":%d,\"ms\":%.3f,\"tflops\":%.2f,\"gbs\":%.1f}\n",s.n,s.m,best,1.0*s.n*s.m*(s.m+1)/best/1e9, 1.0*bytes/best/1e6);
    // NN projection: n x m times m x (2m/3)
    { int nb=2*s.m/3; double *C,*O; CK(cudaMalloc(&C,8ull*s.m*nb)); CK(cudaMemset(C,0,8ull*s.m*nb)); O=AS;
      for(int i=0;i<2;i++) cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,(int)s.n,nb,s.m,&al,S,(int)s.n,C,s.m,&be,O,(int)s.n);
      CK(cudaDeviceSynchronize()); best=1e9;
      for(int r=0;r<3;r++){cudaEventRecord(e0); cublasDgemm(h,CUBLAS_OP_N,CUBLAS_OP_N,(int)s.n,nb,s.m,&al,S,(int)s.n,C,s.m,&be,O,(int)s.n); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); float ms=time_ms(e0,e1); if(ms<best)best=ms;}
      printf("{\"probe\":\"cublas_dgemm_nn_proj\",\"n\":%ld,\"m\":%d,\"nb\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n",s.n,s.m,nb,best,2.0*s.n*s.m*nb/best/1e9);
      cudaFree(C);}
    cudaFree(S);cudaFree(AS);cudaFree(G);
  }
  return 0;
}
