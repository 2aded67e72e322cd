#include <cstdint>
__device__ __forceinline__ uint64_t pk(float a,float b){ uint64_t r; asm("mov.b64 %0,{%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b){ uint64_t d; asm("mul.rn.f32x2 %0,%1,%2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b){ uint64_t d; asm("add.rn.f32x2 %0,%1,%2;":"=l"(d):"l"(a),"l"(b)); return d;}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b){ uint64_t d; asm("sub.rn.f32x2 %0,%1,%2;":"=l"(d):"l"(a),"l"(b)); return d;}
__global__ void k(const float2* __restrict__ in, uint64_t* out, int n){
  extern __shared__ float2 sm[];
  for(int i=threadIdx.x;i<n;i+=blockDim.x) sm[i]=in[i];
  __syncthreads();
  uint64_t a=out[threadIdx.x], b=out[threadIdx.x+128];
  uint64_t accr=0, acci=0;
  #pragma unroll 4
  for(int i=0;i<n;i++){
    float2 s=sm[i];
    uint64_t cc=pk(s.x,s.x), dd=pk(s.y,s.y);
    uint64_t re=sub2(mul2(a,cc),mul2(b,dd));
    uint64_t im=add2(mul2(a,dd),mul2(b,cc));
    accr=add2(accr,re); acci=add2(acci,im);
  }
  out[threadIdx.x]=accr; out[threadIdx.x+128]=acci;
}
