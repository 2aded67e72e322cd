// dependent-issue latency of packed FP32 ops, and throughput vs (warps/SMSP, ILP)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template<int ILP, int OP> __global__ void k(float* out, int iters, float m, float c, long long* cyc){
  unsigned long long p[ILP], pm, pc;
  for(int i=0;i<ILP;i++){ float a=threadIdx.x*0.001f+i; asm("mov.b64 %0,{%1,%2};":"=l"(p[i]):"f"(a),"f"(a+0.5f)); }
  asm("mov.b64 %0,{%1,%1};":"=l"(pm):"f"(m)); asm("mov.b64 %0,{%1,%1};":"=l"(pc):"f"(c));
  float s[ILP]; for(int i=0;i<ILP;i++) s[i]=threadIdx.x+i;
  long long t0=clock64();
  for(int it=0; it<iters; it++){
    #pragma unroll
    for(int u=0;u<16;u++){
      #pragma unroll
      for(int i=0;i<ILP;i++){
        if(OP==0) asm volatile("fma.rn.f32x2 %0,%0,%1,%2;":"+l"(p[i]):"l"(pm),"l"(pc));
        if(OP==1) asm volatile("mul.rn.f32x2 %0,%0,%1;":"+l"(p[i]):"l"(pm));
        if(OP==2) s[i]=__fmaf_rn(s[i],m,c);
      }
    }
  }
  long long t1=clock64();
  float r=0; for(int i=0;i<ILP;i++){ float a,b; asm("mov.b64 {%0,%1},%2;":"=f"(a),"=f"(b):"l"(p[i])); r+=a+b+s[i]; }
  out[blockIdx.x*blockDim.x+threadIdx.x]=r;
  if(threadIdx.x==0 && blockIdx.x==0) cyc[0]=t1-t0;
}
template<int ILP,int OP> void run(int warps, const char* nm){
  float* out; long long* cyc; cudaMalloc(&out, 148*warps*32*4); cudaMalloc(&cyc,8);
  int iters=4000;
  k<ILP,OP><<<148, warps*32>>>(out,100,1.0000001f,1e-9f,cyc); cudaDeviceSynchronize();
  k<ILP,OP><<<148, warps*32>>>(out,iters,1.0000001f,1e-9f,cyc); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h,cyc,8,cudaMemcpyDeviceToHost);
  double per_warp_instr = (double)iters*16*ILP;
  printf("%-6s warps/SM=%2d ILP=%d: cycles/instr(per warp)=%.2f  instr/clk/SMSP=%.3f\n", nm, warps, ILP, h/per_warp_instr, per_warp_instr*(warps/4.0)/h);
  cudaFree(out); cudaFree(cyc);
}
int main(){
  run<1,0>(4,"FFMA2"); run<2,0>(4,"FFMA2"); run<3,0>(4,"FFMA2"); run<4,0>(4,"FFMA2"); run<8,0>(4,"FFMA2");
  run<1,0>(8,"FFMA2"); run<2,0>(8,"FFMA2"); run<3,0>(8,"FFMA2"); run<4,0>(8,"FFMA2");
  run<1,0>(12,"FFMA2"); run<2,0>(12,"FFMA2"); run<1,0>(16,"FFMA2"); run<2,0>(16,"FFMA2");
  run<1,1>(4,"FMUL2"); run<2,1>(4,"FMUL2"); run<4,1>(4,"FMUL2");
  run<1,2>(4,"FFMA"); run<2,2>(4,"FFMA"); run<4,2>(4,"FFMA"); run<8,2>(4,"FFMA");
  return 0;
}
