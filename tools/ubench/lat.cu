// Latency microbenchmarks for the sequential critical paths (Jacobi rotation chain, Householder).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsq64(double x){double y;asm volatile("rsqrt.approx.ftz.f64 %0, %1;":"=d"(y):"d"(x));return y;}
__device__ __forceinline__ double rcp64(double x){double y;asm volatile("rcp.approx.ftz.f64 %0, %1;":"=d"(y):"d"(x));return y;}
template<int OP> __global__ void k(double* out, long long* cyc, double seed, int n){
  __shared__ double sh[64];
  double x=seed+threadIdx.x*1e-9, y=1.0000001;
  sh[threadIdx.x&63]=x; __syncthreads();
  unsigned long long g0; asm volatile("mov.u64 %0, %globaltimer;":"=l"(g0));
  long long t0=clock64();
  for(int i=0;i<n;++i){
    if(OP==0) x=fma(x,y,1e-9);
    if(OP==1) x=x*y;
    if(OP==2) x=x+y;
    if(OP==3) x=rsq64(x)+1.0;      // MUFU.RSQ64H + DADD
    if(OP==4) x=double(float(x))*y; // F2F both ways + DMUL
    if(OP==5) { __syncthreads(); }
    if(OP==6) { x=sh[(threadIdx.x+int(x))&63]; } // LDS dependent (x stays small)
    if(OP==7) x=rcp64(x)+1.0;
    if(OP==8) { float f=__frcp_rn(float(x)); x=double(f)+1.0; }
    if(OP==9) { x=__shfl_xor_sync(0xffffffffu,x,1)+y; }
    if(OP==10){ sh[threadIdx.x&63]=x; __syncwarp(); x=sh[(threadIdx.x+1)&63]+y; __syncwarp(); }
  }
  long long t1=clock64();
  unsigned long long g1; asm volatile("mov.u64 %0, %globaltimer;":"=l"(g1));
  if(threadIdx.x==0) { cyc[0]=t1-t0; cyc[1]=(long long)(g1-g0); }
  out[threadIdx.x]=x;
}
int main(){
  double* out; long long* cyc; cudaMalloc(&out,8192); cudaMalloc(&cyc,64);
  const char* names[]={"DFMA dep","DMUL dep","DADD dep","RSQ64H+DADD","F2F x2 + DMUL","BAR.SYNC","LDS dep (+cvt)","RCP64H+DADD","F2F+MUFU.RCP+F2F+DADD","SHFL+DADD","STS+LDS+DADD"};
  int n=65536;
  for(int threads: {32,512}){
    printf("threads=%d\n",threads);
    for(int op=0;op<11;++op){
      long long h=0, hn[2]={0,0};
      for(int rep=0;rep<2;++rep){
      switch(op){
        case 0:k<0><<<1,threads>>>(out,cyc,1.0,n);break; case 1:k<1><<<1,threads>>>(out,cyc,1.0,n);break;
        case 2:k<2><<<1,threads>>>(out,cyc,1.0,n);break; case 3:k<3><<<1,threads>>>(out,cyc,1.0,n);break;
        case 4:k<4><<<1,threads>>>(out,cyc,1.0,n);break; case 5:k<5><<<1,threads>>>(out,cyc,1.0,n);break;
        case 6:k<6><<<1,threads>>>(out,cyc,0.0,n);break; case 7:k<7><<<1,threads>>>(out,cyc,1.0,n);break;
        case 8:k<8><<<1,threads>>>(out,cyc,1.0,n);break; case 9:k<9><<<1,threads>>>(out,cyc,1.0,n);break;
        case 10:k<10><<<1,threads>>>(out,cyc,1.0,n);break;
      }
      cudaDeviceSynchronize(); cudaMemcpy(hn,cyc,16,cudaMemcpyDeviceToHost); h=hn[0];}
      printf("  %-28s %.1f cycles/iter   (%.0f MHz)\n",names[op],double(h)/n, 1e3*double(hn[0])/double(hn[1]));
    }
  }
  return 0;
}
