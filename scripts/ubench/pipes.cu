// pipes.cu — measurement tool (not product): issue throughput of the instruction classes the integrate / raycast kernels are
// made of, on the box's B200, in warp-instructions per cycle per SM. Built and run by scripts/ubench/run.sh under gpurun.
// Each kernel keeps 8 independent dependency chains per thread, 32 warps per SM (8 per scheduler), and brackets the loop with
// clock64() on every SM; the reported figure is (instructions of the class issued by the SM) / (cycles of that SM), median over SMs.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

enum Op { FADD, FMUL, FFMA, FADD2, FMUL2, FFMA2, IADD3, LOP3, IMAD, FMNMX, FSETP_SEL, F2I, I2F, MUFU_RCP, MIX_FADD_LOP, MIX_FADD2_LOP, MIX_FFMA_IADD,
          MIX_FADD_FADD2, LDS32, SHFL, POPC, FADD_DEP, NOPS };
static const char *names[] = {"FADD", "FMUL", "FFMA", "FADD2 (f32x2)", "FMUL2 (f32x2)", "FFMA2 (f32x2)", "IADD3", "LOP3", "IMAD", "FMNMX", "FSETP+SEL",
                              "F2I", "I2F", "MUFU.RCP", "mix FADD+LOP3 (1:1)", "mix FADD2+LOP3 (1:1)", "mix FFMA+IADD3 (1:1)", "mix FADD+FADD2 (1:1)",
                              "LDS.32", "SHFL", "POPC", "FADD dependent chain"};

template <int OP>
__global__ void __launch_bounds__(1024) k(float *out, long long *cycles, float seed) {
  __shared__ float sm[1024];
  sm[threadIdx.x] = seed + threadIdx.x;
  __syncthreads();
  float a[CHAINS], b[CHAINS];
  int ia[CHAINS];
  unsigned long long pa[CHAINS];
  for (int c = 0; c < CHAINS; ++c) {
    a[c] = seed + c + threadIdx.x * 0.001f; b[c] = 1.0f + seed * (c + 1); ia[c] = threadIdx.x + c;
    float2 t = make_float2(a[c], b[c]); pa[c] = *reinterpret_cast<unsigned long long *>(&t);
  }
  const float m = 1.0000001f, s = 0.5f;
  unsigned long long pm; { float2 t = make_float2(m, m); pm = *reinterpret_cast<unsigned long long *>(&t); }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(s));
      if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(m));
      if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m), "f"(b[c]));
      if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[c]) : "l"(pm));
      if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(pa[c]) : "l"(pm));
      if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(pa[c]) : "l"(pm));
      if (OP == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[c]) : "r"(ia[(c + 1) % CHAINS]));
      if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[c]) : "r"(ia[(c + 1) % CHAINS]), "r"(it));
      if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(ia[c]) : "r"(ia[(c + 1) % CHAINS]), "r"(it));
      if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b[c]));
      if (OP == FSETP_SEL) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %1, %0, p;}" : "+f"(a[c]) : "f"(b[c]));
      if (OP == F2I) asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(ia[c]) : "f"(a[c]));
      if (OP == I2F) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(a[c]) : "r"(ia[c]));
      if (OP == MUFU_RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[c]));
      if (OP == MIX_FADD_LOP) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(s)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[c]) : "r"(ia[(c + 1) % CHAINS]), "r"(it)); }
      if (OP == MIX_FADD2_LOP) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[c]) : "l"(pm)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[c]) : "r"(ia[(c + 1) % CHAINS]), "r"(it)); }
      if (OP == MIX_FFMA_IADD) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(m), "f"(b[c])); asm volatile("add.s32 %0, %0, %1;" : "+r"(ia[c]) : "r"(ia[(c + 1) % CHAINS])); }
      if (OP == MIX_FADD_FADD2) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(s)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[c]) : "l"(pm)); }
      if (OP == LDS32) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a[c]) : "r"((unsigned)__cvta_generic_to_shared(&sm[(threadIdx.x + c * 32) & 1023])) : "memory");
      if (OP == SHFL) asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(ia[c]));
      if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(ia[c]));
      if (OP == FADD_DEP) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[0]) : "f"(s));
    }
  }
  const long long t1 = clock64();
  float acc = 0; int iacc = 0;
  for (int c = 0; c < CHAINS; ++c) { float2 t = *reinterpret_cast<float2 *>(&pa[c]); acc += a[c] + t.x + t.y; iacc += ia[c]; }
  if (acc == 12345.678f || iacc == 0x7fffffff) out[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(int sms, float *out, long long *dcyc) {
  k<OP><<<sms, 1024>>>(out, dcyc, 0.25f);
  cudaDeviceSynchronize();
  k<OP><<<sms, 1024>>>(out, dcyc, 0.25f);
  cudaDeviceSynchronize();
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double cyc = (double)h[sms / 2];
  const int perIter = (OP >= MIX_FADD_LOP && OP <= MIX_FADD_FADD2) ? 2 : 1;
  const double inst = 32.0 * ITERS * CHAINS * perIter;   // warp-instructions of the class per SM (32 warps)
  printf("%-26s %7.3f warp-inst/clk/SM   (%.0f cycles)\n", names[OP], inst / cyc, cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  float *out; long long *dcyc;
  cudaMalloc(&out, 4); cudaMalloc(&dcyc, sizeof(long long) * sms);
  printf("%s, %d SMs; 32 warps/SM, %d independent chains per thread\n", p.name, sms, CHAINS);
  run<FADD>(sms, out, dcyc); run<FMUL>(sms, out, dcyc); run<FFMA>(sms, out, dcyc); run<FADD2>(sms, out, dcyc); run<FMUL2>(sms, out, dcyc);
  run<FFMA2>(sms, out, dcyc); run<IADD3>(sms, out, dcyc); run<LOP3>(sms, out, dcyc); run<IMAD>(sms, out, dcyc); run<FMNMX>(sms, out, dcyc);
  run<FSETP_SEL>(sms, out, dcyc); run<F2I>(sms, out, dcyc); run<I2F>(sms, out, dcyc); run<MUFU_RCP>(sms, out, dcyc);
  run<MIX_FADD_LOP>(sms, out, dcyc); run<MIX_FADD2_LOP>(sms, out, dcyc); run<MIX_FFMA_IADD>(sms, out, dcyc); run<MIX_FADD_FADD2>(sms, out, dcyc);
  run<LDS32>(sms, out, dcyc); run<SHFL>(sms, out, dcyc); run<POPC>(sms, out, dcyc); run<FADD_DEP>(sms, out, dcyc);
  return 0;
}
