// ubench.cu -- issue/pipe throughput probes for the B200 kernel design decisions (scratch, not product):
// scalar FFMA vs packed FFMA2/FADD2/FMUL2, MUFU, LDS.64/128, shuffle.  Prints warp-instr/clk/SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v){ float a,b; asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); return a+b; }
#define ITER 2048
template<int MODE> __global__ void k(float* out, float s, int iters){
  __shared__ float4 sm[1024];
  float a[8]; u64 p[8];
  for(int i=0;i<8;i++){ a[i]=threadIdx.x*0.001f+i; p[i]=pk(a[i],a[i]+1.f);} 
  u64 ps = pk(s,s*0.5f), pt = pk(0.25f, 0.125f);
  if (MODE>=6) { for (int i=threadIdx.x;i<1024;i+=blockDim.x) sm[i]=make_float4(i,i,i,i); __syncthreads(); }
  int idx = threadIdx.x;
  for(int it=0; it<iters; ++it){
#pragma unroll
    for(int r=0;r<4;r++){
#pragma unroll
      for(int i=0;i<8;i++){
        if (MODE==0) a[i] = fmaf(a[i], s, 0.25f);
        if (MODE==1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;":"+l"(p[i]):"l"(ps),"l"(pt));
        if (MODE==2) asm volatile("add.rn.f32x2 %0, %0, %1;":"+l"(p[i]):"l"(ps));
        if (MODE==3) asm volatile("mul.rn.f32x2 %0, %0, %1;":"+l"(p[i]):"l"(ps));
        if (MODE==4) a[i] = __fadd_rn(a[i], s);
        if (MODE==5) asm volatile("rsqrt.approx.ftz.f32 %0, %0;":"+f"(a[i]));
        if (MODE==6) { float vx,vy; unsigned ad=(unsigned)__cvta_generic_to_shared((float2*)sm + ((idx + i*32 + it) & 2047)); asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];":"=f"(vx),"=f"(vy):"r"(ad)); a[i]+=vx; }
        if (MODE==7) { float vx,vy,vz,vw; unsigned ad=(unsigned)__cvta_generic_to_shared(sm + ((idx + i*32 + it) & 1023)); asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];":"=f"(vx),"=f"(vy),"=f"(vz),"=f"(vw):"r"(ad)); a[i]+=vx; }
        if (MODE==8) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1);
        if (MODE==9) { a[i] = fmaf(a[i], s, 0.25f); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;":"+l"(p[i]):"l"(ps),"l"(pt)); }
        if (MODE==10) { a[i] = __fadd_rn(a[i], s); asm volatile("lop3.b32 %0, %0, 0x55, %1, 0x96;":"+r"(idx):"r"(it)); }
      }
    }
  }
  float acc=0; for(int i=0;i<8;i++) acc += a[i] + lo(p[i]);
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc + idx;
}
template<int MODE> void run(const char* name, double instPerIter){
  float* d; cudaMalloc(&d, 148*8*256*4);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148*8,256>>>(d,1.0001f,16);
  cudaEventRecord(e0); k<MODE><<<148*8,256>>>(d,1.0001f,ITER); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double warps = 148.0*8*8, insts = warps*ITER*32.0*instPerIter; // warp-instr of the probed kind
  double clks = ms*1e-3*clk*1e3;
  printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, insts/clks/148.0, clk/1000);
  cudaFree(d);
}
int main(){
  run<0>("FFMA scalar",1); run<1>("FFMA2 packed",1); run<2>("FADD2 packed",1); run<3>("FMUL2 packed",1);
  run<4>("FADD scalar",1); run<5>("MUFU.RSQ",1); run<6>("LDS.64",1); run<7>("LDS.128",1); run<8>("SHFL.UP",1);
  run<9>("FFMA+FFMA2 pair (pairs)",1); run<10>("FADD+LOP3 pair (pairs)",1);
  return 0;
}
