#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a,float b){ u64 r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v,float&a,float&b){ asm("mov.b64 {%0,%1},%2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a,u64 b){ u64 d; asm volatile("mul.rn.f32x2 %0,%1,%2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ u64 fma2(u64 a,u64 b,u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
// MODE: 0 FMUL2 pair*pair(reg) ; 1 FMUL2 pair*scalar.F32(reg) ; 2 FMUL2 swizzled(-hi,lo)*scalar ; 3 FFMA2 R,UR,R ; 4 FFMA2 3 distinct reg pairs
// 5: 7 FMUL2 + 1 FADD ; 6: 7 FMUL2 + 2 FADD ; 7: 8 FMUL2 + 1 scalar FMUL; 8: FMUL2 pair*imm
template<int MODE> __global__ void __launch_bounds__(128,4) k(float* out,int iters,float m,float c,float one){
  u64 p[8]; float s[8];
  #pragma unroll
  for(int i=0;i<8;i++){ float a=threadIdx.x*0.001f+i; p[i]=pk(a,a+0.5f); s[i]=1.0f+1e-7f*(threadIdx.x+i);} 
  u64 pm=pk(m+threadIdx.x*1e-9f,m), pc=pk(c,c+threadIdx.x*1e-9f), pone=pk(one,one);
  float fs=m+threadIdx.x*1e-9f;
  __shared__ float4 sm[64]; if(threadIdx.x<64) sm[threadIdx.x]=make_float4(m,c,m,c); __syncthreads();
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int u=0;u<8;u++){
      #pragma unroll
      for(int i=0;i<8;i++){
        if(MODE==0) p[i]=mul2(p[i],pm);
        if(MODE==1) p[i]=mul2(p[i],pk(s[i&3],s[i&3]));
        if(MODE==2){ float a,b; upk(pm,a,b); p[i]=mul2(pk(-b,a),pk(s[i&3],s[i&3])); pm=p[(i+3)&7]; }
        if(MODE==3) p[i]=fma2(p[i],pone,pc);
        if(MODE==4) p[i]=fma2(p[i],pm,pc);
        if(MODE==5){ p[i]=mul2(p[i],pm); if(i==7) s[u]=__fadd_rn(s[u],fs); }
        if(MODE==6){ p[i]=mul2(p[i],pm); if(i>=6) s[u]=__fadd_rn(s[u],fs); }
        if(MODE==7){ p[i]=mul2(p[i],pm); if(i==7) s[u]=__fmul_rn(s[u],fs); }
        if(MODE==8) p[i]=mul2(p[i],pk(1.0000001f,1.0000001f));
        if(MODE==9){ p[i]=mul2(p[i],pm); if(i==7){ float t; asm volatile("mov.b32 %0,%1;":"=f"(t):"f"(s[u])); s[(u+1)&7]=t; } }
        if(MODE==10){ p[i]=mul2(p[i],pm); if(i>=6){ float t; asm volatile("mov.b32 %0,%1;":"=f"(t):"f"(s[(u+i)&7])); s[(u+i+1)&7]=t; } }
        if(MODE==11){ p[i]=mul2(p[i],pm); if(i==7){ float4 v=sm[(it+u)&63]; s[u]=v.x; s[(u+1)&7]=v.y; } }
        if(MODE==12){ p[i]=mul2(p[i],pm); if(i>=4){ float t; asm volatile("mov.b32 %0,%1;":"=f"(t):"f"(s[(u+i)&7])); s[(u+i+1)&7]=t; } }
      }
    }
  }
  float r=0; for(int i=0;i<8;i++){ float a,b; upk(p[i],a,b); r+=a+b+s[i]; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
}
template<int MODE> void run(const char* nm,double packed_per_iter,double scalar_per_iter){
  float* out; cudaMalloc(&out,148*4*128*4); int iters=10000;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148*4,128>>>(out,100,1.0000001f,1e-9f,1.0f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148*4,128>>>(out,iters,1.0000001f,1e-9f,1.0f); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double thr=(double)148*4*128*iters;
  double lane = thr*(packed_per_iter*2+scalar_per_iter);
  printf("%-34s %.3f ms  lane-ops %.2f T/s (%.1f%% of 37.1)\n",nm,ms,lane/ms/1e9,100*lane/ms/1e9/37.1);
  cudaFree(out);
}
int main(){
  run<0>("FMUL2 pair*pair",64,0); run<1>("FMUL2 pair*R.F32",64,0); run<2>("FMUL2 swz(-hi,lo)*R.F32",64,0);
  run<3>("FFMA2 R,UR.F32,R",64,0); run<4>("FFMA2 3 reg pairs",64,0);
  run<5>("7 FMUL2:1 FADD (of 8 packed)",64,8); run<6>("8 FMUL2: 2 FADD",64,16); run<7>("8 FMUL2: 1 FMUL",64,8); run<8>("FMUL2 pair*imm",64,0);
  run<9>("8 FMUL2: 1 MOV",64,0); run<10>("8 FMUL2: 2 MOV",64,0); run<11>("8 FMUL2: 1 LDS.128 bcast",64,0); run<12>("8 FMUL2: 4 MOV",64,0);
  return 0;
}
