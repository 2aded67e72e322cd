// FP32 pipe microbenchmark for sm_100a: scalar vs packed f32x2 issue rates.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint64_t pk(float a,float b){ uint64_t r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ float lo(uint64_t v){ float a,b; asm("mov.b64 {%0,%1},%2;":"=f"(a),"=f"(b):"l"(v)); return a+b;}

template<int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s, long long* cyc){
  float a[8]; uint64_t p[8];
  #pragma unroll
  for(int i=0;i<8;i++){ a[i]=threadIdx.x*0.001f+i; p[i]=pk(a[i],a[i]+0.5f);}
  float m=s, c=s*0.5f; uint64_t pm=pk(m,m), pc=pk(c,c);
  __shared__ float4 sm[64];
  if(threadIdx.x<64) sm[threadIdx.x]=make_float4(s,s,s,s);
  __syncthreads();
  int xi=threadIdx.x;
  long long t0=clock64();
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int u=0;u<8;u++){
      #pragma unroll
      for(int i=0;i<8;i++){
        if(MODE==0) a[i]=__fmaf_rn(a[i],m,c);
        if(MODE==1) a[i]=__fmul_rn(a[i],m);
        if(MODE==2) a[i]=__fadd_rn(a[i],c);
        if(MODE==3){ if(i&1) a[i]=__fmul_rn(a[i],m); else a[i]=__fadd_rn(a[i],c);} 
        if(MODE==4) asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p[i]):"l"(pm),"l"(pc));
        if(MODE==5) asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pm));
        if(MODE==6) asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pc));
        if(MODE==7){ if(i&1) asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pm)); else asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pc)); }
        if(MODE==8){ // packed + 1 int ALU op per 2 packed
          asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pm));
          if(i&1) xi = (xi ^ it) + u;
        }
        if(MODE==9){ // scalar + int op per 2 scalar
          a[i]=__fmul_rn(a[i],m);
          if(i&1) xi = (xi ^ it) + u;
        }
        if(MODE==10){ // packed + 1 int op per packed
          asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pm));
          xi = (xi ^ it) + u + i;
        }
        if(MODE==11){ // packed fed by a broadcast LDS.128 every 4 packed
          if((i&3)==0){ float4 v=sm[(it+u+i)&63]; pm=pk(v.x,v.y); pc=pk(v.z,v.w);} 
          if(i&1) asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pm)); else asm volatile("add.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pc));
        }
        if(MODE==12){ // scalar fed by LDS.64 every 4 scalar
          if((i&3)==0){ float2 v=*(float2*)&sm[(it+u+i)&63]; m=v.x; c=v.y;} 
          if(i&1) a[i]=__fmul_rn(a[i],m); else a[i]=__fadd_rn(a[i],c);
        }
        if(MODE==13){ // packed + 1 MOV-ish (fp32 register shuffle among pairs) per 2 packed
          asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pm));
        }
      }
      if(MODE==13){ uint64_t t=p[0]; 
        #pragma unroll
        for(int i=0;i<7;i++) p[i]=p[i+1]; p[7]=t; }
    }
  }
  long long t1=clock64();
  float r=0; for(int i=0;i<8;i++) r+=a[i]+lo(p[i]);
  out[blockIdx.x*blockDim.x+threadIdx.x]=r+xi;
  if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}

template<int MODE> int run(const char* name, int flop_per, int warps_per_sm){
  int nsm=148; int threads=256; int blocks=nsm*(warps_per_sm*32/threads);
  float* out; long long* cyc; CK(cudaMalloc(&out,blocks*threads*4)); CK(cudaMalloc(&cyc,blocks*8));
  int iters=20000;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks,threads>>>(out,200,1.0000001f,cyc); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); k<MODE><<<blocks,threads>>>(out,iters,1.0000001f,cyc); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  long long h[1]; cudaMemcpy(h,cyc,8,cudaMemcpyDeviceToHost);
  double ninstr=(double)iters*64; // FP instrs per thread
  double warp_instr_per_sm = ninstr*warps_per_sm;
  double ipc = warp_instr_per_sm/(double)h[0];
  double total_lane_ops = ninstr*blocks*threads*flop_per; // flop_per = lanes-worth per instr (1 scalar, 2 packed)
  printf("%-28s warps/SM=%2d  ms=%8.3f cyc=%lld  FPinstr/clk/SM=%.3f  laneops/clk/SM=%.1f  Tlaneops/s=%.2f  eff_clk=%.0f MHz\n",
         name,warps_per_sm,ms,h[0],ipc,ipc*32*flop_per,total_lane_ops/ms/1e9,(double)h[0]/ms/1e3);
  cudaFree(out); cudaFree(cyc); return 0;
}
int main(){
  for(int w: {16,32,64}){
    run<0>("FFMA scalar",1,w); run<1>("FMUL scalar",1,w); run<2>("FADD scalar",1,w); run<3>("FMUL/FADD alt scalar",1,w);
    run<4>("FFMA2 packed",2,w); run<5>("FMUL2 packed",2,w); run<6>("FADD2 packed",2,w); run<7>("FMUL2/FADD2 alt",2,w);
    run<8>("FMUL2 + 0.5 int/instr",2,w); run<9>("FMUL + 0.5 int/instr",1,w); run<10>("FMUL2 + 1 int/instr",2,w);
    run<11>("packed alt + LDS128 bcast/4",2,w); run<12>("scalar alt + LDS64 bcast/4",1,w); run<13>("FMUL2 + 64b rot moves",2,w);
  }
  return 0;
}
