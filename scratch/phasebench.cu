#include <cstdio>
#include "../aero-cli_b200/csrc/ddc_kernels.cuh"
using namespace aeroddc;
// MODE 0: NCO chain only. MODE 1: mix + HB cascade only (osc values constant). MODE 2: NCO with packed add/sub instead of scalar FADD
template<int MODE> __global__ void __launch_bounds__(128, 4) kb(float* out, int iters, float one, float mone){
  __shared__ float2 tile[kTile];
  for(int i=threadIdx.x;i<kTile;i+=blockDim.x){ float c=0.001f*i, d=0.5f-0.002f*i; tile[i]=make_float2(c,d);} 
  __syncthreads();
  Ones k1; k1.one=bcast2(one);
  Rot rot; float cr=0.99f+1e-6f*threadIdx.x, sr=0.14f; rot.a=pack2(cr,sr); rot.b=pack2(-sr,cr);
  HbState hb[4];
  for(int s=0;s<4;s++){ for(int k=0;k<5;k++) hb[s].e[k]=pzero(); for(int k=0;k<3;k++) hb[s].o[k]=pzero(); }
  float oa=1.f, ob=0.f; P2 acc=pzero(); P2 pmone=bcast2(mone);
  for(int it=0; it<iters; it++){
    #pragma unroll 1
    for(int c=0;c<kTile;c+=kChunk){
      if(MODE==0){
        #pragma unroll
        for(int i=0;i<kChunk;i++) nco_step(k1,oa,ob,rot);
      } else if(MODE==2){
        #pragma unroll
        for(int i=0;i<kChunk;i++){
          const P2 n = add2(k1, mul2s(rot.a, oa), mul2s(rot.b, ob));
          P2 sq = mul2(n,n); float r2,i2; unpack2(sq,r2,i2);
          P2 s = add2(k1, sq, pack2(i2,r2));            // (s,s)
          P2 nm = fma2(s, pmone, bcast2(1.95f));        // 1.95 - s in both lanes
          unpack2(mul2(n,nm), oa, ob);
        }
      } else {
        P2 x0[kChunk];
        #pragma unroll
        for(int i=0;i<kChunk;i++) x0[i]=mix(k1,oa,ob,tile[c+i]);
        P2 x1[8],x2[4],x3[2];
        #pragma unroll
        for(int j=0;j<8;j++) x1[j]=hb_pair(k1,hb[0],x0[2*j],x0[2*j+1]);
        #pragma unroll
        for(int j=0;j<4;j++) x2[j]=hb_pair(k1,hb[1],x1[2*j],x1[2*j+1]);
        #pragma unroll
        for(int j=0;j<2;j++) x3[j]=hb_pair(k1,hb[2],x2[2*j],x2[2*j+1]);
        acc=add2(k1,acc,hb_pair(k1,hb[3],x3[0],x3[1]));
      }
    }
  }
  float a,b; unpack2(acc,a,b); out[blockIdx.x*blockDim.x+threadIdx.x]=a+b+oa+ob;
}
template<int MODE> void run(const char* nm, double pipe_cycles_per_chunk){
  float* out; cudaMalloc(&out, 148*4*128*4);
  for(int ctas: {1,2,4}){
    int iters=200; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kb<MODE><<<148*ctas,128>>>(out,10,1.0f,-1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0); kb<MODE><<<148*ctas,128>>>(out,iters,1.0f,-1.0f); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double warp_chunks=(double)148*ctas*4*iters*(kTile/kChunk);
    double need = warp_chunks*pipe_cycles_per_chunk/(148*4);   // cycles per SMSP
    double have = ms*1e-3*1.92e9;
    printf("%-28s ctas/SM=%d %.3f ms pipe util (at 1.92GHz) %.1f%%\n",nm,ctas,ms,100*need/have);
  }
}
int main(){
  run<0>("NCO chain (5 packed+2 scalar)", 16*(5*2+2*1));
  run<2>("NCO chain all packed (7)", 16*7*2);
  run<1>("mix+HB cascade", (48+150)*2);
  return 0;
}
