// Issue rate of single SASS opcodes on sm_100a: 1024 threads (8 warps per scheduler), 8 independent chains per thread.
// Prints cycles per warp instruction and scheduler. nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
#define ITERS 2048
template <int OP>
__device__ __forceinline__ unsigned step(unsigned x, unsigned c) {
   unsigned d;
   if (OP == 0) asm volatile("lop3.b32 %0, %1, %2, %1, 0x96;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 1) asm volatile("shr.u32 %0, %1, 3;" : "=r"(d) : "r"(x));
   else if (OP == 2) asm volatile("shf.l.wrap.b32 %0, %1, %1, %2;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 3) asm volatile("shf.r.wrap.b32 %0, %1, %1, %2;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 4) asm volatile("mad.hi.u32 %0, %1, %2, %1;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 5) asm volatile("popc.b32 %0, %1;" : "=r"(d) : "r"(x));
   else if (OP == 6) asm volatile("mad.lo.u32 %0, %1, %2, %1;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 7) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 8) asm volatile("prmt.b32 %0, %1, %2, 0x3021;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 9) asm volatile("bfe.u32 %0, %1, 5, 11;" : "=r"(d) : "r"(x));
   else if (OP == 10) asm volatile("shl.b32 %0, %1, 3;" : "=r"(d) : "r"(x));
   else if (OP == 11) asm volatile("{.reg .pred p; setp.lt.s32 p, %1, %2; selp.u32 %0, %1, %2, p;}" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 12) { d = c + (x >> 31); asm volatile("" : "+r"(d)); }             // LEA.HI
   else if (OP == 13) { d = c + (x >> 3); asm volatile("" : "+r"(d)); }              // LEA.HI
   else if (OP == 14) { d = (x << 2) + c; asm volatile("" : "+r"(d)); }              // LEA / IMAD
   else if (OP == 15) asm volatile("shf.l.clamp.b32 %0, %1, %1, %2;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 16) asm volatile("shr.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 17) asm volatile("bfind.u32 %0, %1;" : "=r"(d) : "r"(x));
   else if (OP == 18) asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 19) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %1;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 20) asm volatile("dp4a.u32.u32 %0, %1, %2, %1;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 21) asm volatile("min.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 22) asm volatile("bmsk.clamp.b32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(c));
   else if (OP == 23) asm volatile("szext.wrap.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(c));
   else d = x;
   return d;
}
template <int OP>
__global__ void probe(unsigned* out, unsigned long long* cycles, unsigned c) {
   unsigned x[CHAINS];
   for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x * 7919u + i;
   __syncthreads();
   const unsigned long long begin = clock64();
   for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) x[i] = step<OP>(x[i], c);
   }
   const unsigned long long end = clock64();
   unsigned sum = 0;
   for (int i = 0; i < CHAINS; ++i) sum += x[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
   if (threadIdx.x == 0) cycles[blockIdx.x] = end - begin;
}
template <int OP>
void run(const char* name, unsigned* out, unsigned long long* cycles) {
   probe<OP><<<148, 1024>>>(out, cycles, 5);
   cudaDeviceSynchronize();
   probe<OP><<<148, 1024>>>(out, cycles, 5);
   cudaDeviceSynchronize();
   unsigned long long host[148];
   cudaMemcpy(host, cycles, sizeof(host), cudaMemcpyDeviceToHost);
   double mean = 0;
   for (int i = 0; i < 148; ++i) mean += host[i];
   mean /= 148;
   // per scheduler: 8 warps x ITERS x CHAINS instructions
   printf("%-28s %6.2f cycles per warp instruction and scheduler\n", name, mean / (8.0 * ITERS * CHAINS));
}
int main() {
   unsigned* out; unsigned long long* cycles;
   cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cycles, 148 * 8);
   run<0>("lop3", out, cycles); run<1>("shr imm (SHF.R.U32)", out, cycles); run<2>("shf.l.wrap reg", out, cycles);
   run<3>("shf.r.wrap reg", out, cycles); run<4>("mad.hi (IMAD.HI)", out, cycles); run<5>("popc", out, cycles);
   run<6>("mad.lo (IMAD)", out, cycles); run<7>("add", out, cycles); run<8>("prmt", out, cycles); run<9>("bfe", out, cycles);
   run<10>("shl imm", out, cycles); run<11>("setp+selp", out, cycles); run<12>("c + (x >> 31)", out, cycles);
   run<13>("c + (x >> 3)", out, cycles); run<14>("(x << 2) + c", out, cycles); run<15>("shf.l.clamp reg", out, cycles);
   run<16>("shr reg", out, cycles); run<17>("bfind", out, cycles); run<18>("mul.hi", out, cycles); run<19>("vabsdiff4", out, cycles);
   run<20>("dp4a", out, cycles); run<21>("min.u32", out, cycles); run<22>("bmsk", out, cycles); run<23>("szext", out, cycles);
   return 0;
}
