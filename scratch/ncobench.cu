// which ingredient of the NCO step keeps a 4-warp/SMSP dependent chain from saturating the FP32 pipe?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a,float b){ u64 r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v,float&a,float&b){ asm("mov.b64 {%0,%1},%2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a,u64 b){ u64 d; asm("mul.rn.f32x2 %0,%1,%2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ u64 fma2(u64 a,u64 b,u64 c){ u64 d; asm("fma.rn.f32x2 %0,%1,%2,%3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d;}
template<int V> __global__ void __launch_bounds__(128,4) k(float* out,int iters,float one,float rc,float rs,u64 opaque){
  float a=1.0f,b=0.0f; u64 ONE=pk(one,one);
  float c=rc+threadIdx.x*1e-7f, s=rs;
  u64 rotA=pk(c,s), rotB=pk(-s,c);
  if(V==1) rotB ^= opaque;   // opaque (=0): keeps rotB an independent register pair
  u64 ONEr = ONE; if(V==2||V==5) ONEr ^= (opaque+threadIdx.x*0);  // still uniform...
  u64 A=pk(a,a), B=pk(b,b);
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int u=0;u<32;u++){
      if(V==0||V==1||V==2){
        u64 n=fma2(mul2(rotA,pk(a,a)), (V==2?ONEr:ONE), mul2(rotB,pk(b,b)));
        float r2,i2; upk(mul2(n,n),r2,i2);
        float nm=__fsub_rn(1.95f,__fadd_rn(r2,i2));
        upk(mul2(n,pk(nm,nm)),a,b);
      }
      if(V==3){ // all packed, no scalar ops, state as scalars still
        u64 n=fma2(mul2(rotA,pk(a,a)), ONE, mul2(rotB,pk(b,b)));
        u64 sq=mul2(n,n); float r2,i2; upk(sq,r2,i2);
        u64 ss=fma2(sq,ONE,pk(i2,r2));
        u64 nm=fma2(ss,pk(-one,-one),pk(1.95f,1.95f));
        upk(mul2(n,nm),a,b);
      }
      if(V==4){ // pure pair ops, no broadcasts / swizzles: same dependency shape (2 mul, fma, mul, fma, fma, mul)
        u64 n=fma2(mul2(rotA,A), ONE, mul2(rotB,B));
        u64 sq=mul2(n,n);
        u64 ss=fma2(sq,ONE,sq);
        u64 nm=fma2(ss,rotB,rotA);
        A=mul2(n,nm); B=A;
      }
      if(V==5){ // V4 but B chain independent (two parallel chains)
        u64 n=fma2(mul2(rotA,A), ONE, rotB);
        u64 sq=mul2(n,n);
        u64 ss=fma2(sq,ONE,sq);
        u64 nm=fma2(ss,rotB,rotA);
        A=mul2(n,nm);
      }
    }
  }
  float x,y; upk(A,x,y); out[blockIdx.x*blockDim.x+threadIdx.x]=a+b+x+y;
}
template<int V> void run(const char* nm,int ninstr){
  float* out; cudaMalloc(&out,148*4*128*4); int iters=3000;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<148*4,128>>>(out,10,1.0f,0.99f,0.14f,0ull); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<V><<<148*4,128>>>(out,iters,1.0f,0.99f,0.14f,0ull); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double cyc = ms*1e-3*1.92e9;               // per SMSP
  double need = 4.0*iters*32*ninstr*2;       // 4 warps per SMSP, 2 pipe cycles per FP instr (scalar too)
  printf("%-46s %.3f ms  pipe util %.1f%%\n",nm,ms,100*need/cyc);
  cudaFree(out);
}
int main(){
  run<0>("V0 NCO as in kernel (5 packed + 2 scalar)",7);
  run<1>("V1 rotB as its own register pair",7);
  run<2>("V2 ONE from a register pair?",7);
  run<3>("V3 all packed (7), scalar state",7);
  run<4>("V4 pure pair ops, same shape",7);
  run<5>("V5 pure pair, single chain 6 deep (6 instr)",6);
  return 0;
}
