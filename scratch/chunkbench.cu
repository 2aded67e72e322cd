// isolated fast_chunk throughput: how close does the instruction stream itself get to the FP32 pipe peak?
#include <cstdio>
#include "../aero-cli_b200/csrc/ddc_kernels.cuh"
using namespace aeroddc;
template<int NF> __global__ void __launch_bounds__(128, 4) kb(float* out, int iters, float one, int nco_len){
  __shared__ float4 tile[kTile];
  for(int i=threadIdx.x;i<kTile;i+=blockDim.x){ float c=0.001f*i, d=0.5f-0.002f*i; tile[i]=make_float4(c,d,-d,c);} 
  __syncthreads();
  Ones k1; k1.one=bcast2(one);
  Rot rot; float cr=0.99f+1e-6f*threadIdx.x, sr=0.14f; rot.a=pack2(cr,sr); rot.b=pack2(-sr,cr);
  HbState hb[4];
  for(int s=0;s<4;s++){ for(int k=0;k<5;k++) hb[s].e[k]=pzero(); for(int k=0;k<3;k++) hb[s].o[k]=pzero(); }
  float oa=1.f, ob=0.f; int idx=0; long long n_abs=1;
  P2 acc=pzero();
  for(int it=0; it<iters; it++){
    #pragma unroll 1
    for(int c=0;c<kTile;c+=kChunk){
      P2 o[kChunk>>NF];
      fast_chunk<NF,false>(k1,oa,ob,rot,hb,tile+c,o,idx,nco_len,n_abs,0.f,0.f);
      acc=add2(k1,acc,o[0]);
    }
  }
  float a,b; unpack2(acc,a,b); out[blockIdx.x*blockDim.x+threadIdx.x]=a+b+idx;
}
int main(){
  float* out; cudaMalloc(&out, 148*8*128*4);
  for(int ctas: {1,2,3,4}){
    int iters=200;
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kb<4><<<148*ctas,128>>>(out,10,1.0f,1<<30); cudaDeviceSynchronize();
    cudaEventRecord(e0); kb<4><<<148*ctas,128>>>(out,iters,1.0f,1<<30); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double vs=(double)148*ctas*128*iters*kTile; // VFO-samples
    printf("ctas/SM=%d  %.3f ms  %.1f Gsps  pipe-bound(36.75 lane-ops/sample @37.1T)=%.1f%%\n",ctas,ms,vs/ms/1e6, 100*vs/ms/1e6/(37100/36.75));
  }
  return 0;
}
