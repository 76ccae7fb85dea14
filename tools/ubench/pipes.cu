// Instruction-throughput microbenchmark for the integer ops k_minimizers is made of (sm_100a).
// Each kernel runs NCH independent dependency chains per thread; 148 x 4 CTAs x 256 threads keep 8 warps per SM
// sub-partition busy.  Output: warp-instructions per cycle per sub-partition (1.0 = the issue limit).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NIT 2048
#define BODY(NAME, STMT)                                                                        \
    __global__ void __launch_bounds__(256) NAME(uint32_t *out, uint32_t k, long long *cyc)     \
    {                                                                                           \
        uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7; \
        uint32_t b0 = k, b1 = k + 1, b2 = k + 2, b3 = k + 3, b4 = k + 4, b5 = k + 5, b6 = k + 6, b7 = k + 7; \
        const long long t0 = clock64();                                                         \
        for (int i = 0; i < NIT; ++i) { STMT }                                                  \
        const long long t1 = clock64();                                                         \
        out[blockIdx.x * 256 + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ b0 ^ b1 ^ b2 ^ b3 ^ b4 ^ b5 ^ b6 ^ b7; \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                        \
    }
#define X8(OP) OP(a0, b0) OP(a1, b1) OP(a2, b2) OP(a3, b3) OP(a4, b4) OP(a5, b5) OP(a6, b6) OP(a7, b7)
#define LOP(a, b)  asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(b));
#define SHF(a, b)  asm volatile("shf.l.wrap.b32 %0, %0, %0, 1;" : "+r"(a));
#define SHL(a, b)  asm volatile("shl.b32 %0, %0, 1;" : "+r"(a));
#define MAD(a, b)  asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a) : "r"(b));
#define MADK(a, b) asm volatile("mad.lo.u32 %0, %1, 2, %0;" : "+r"(a) : "r"(b));
#define MHI(a, b)  asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
#define MHIK(a, b) asm volatile("mad.hi.u32 %0, %0, 2, %1;" : "+r"(a) : "r"(b));
#define WIDE(a, b) { unsigned long long w_; asm volatile("mul.wide.u32 %0, %1, 2;" : "=l"(w_) : "r"(a)); asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(w_)); }
#define PRM(a, b)  asm volatile("prmt.b32 %0, %0, %1, 0x7650;" : "+r"(a) : "r"(b));
#define ADD(a, b)  asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
#define ADDK(a, b) asm volatile("add.u32 %0, %0, 5;" : "+r"(a));
#define MNX(a, b)  asm volatile("min.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
#define SETP(a, b) asm volatile("{.reg .pred p; setp.le.u32 p, %0, %1; @p add.u32 %0, %0, 3;}" : "+r"(a) : "r"(b));
#define LOPMAD(a, b) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(b)); asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b) : "r"(k));
#define LOPMHI(a, b) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(b)); asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(b) : "r"(k));
#define LOP2MAD(a, b) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(k)); asm volatile("shf.l.wrap.b32 %0, %0, %0, 1;" : "+r"(a)); asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b) : "r"(k));
// rol1(x) ^ t two ways: SHF + LOP3 (ALU, ALU) vs IMAD.SHL + IMAD.HI + LOP3 (FMA, FMA, ALU)
#define ROLA(a, b) { uint32_t r_; asm volatile("shf.l.wrap.b32 %0, %1, %1, 1;" : "=r"(r_) : "r"(a)); asm volatile("xor.b32 %0, %1, %2;" : "=r"(a) : "r"(r_), "r"(b)); }
#define ROLF(a, b) { uint32_t l_, h_; asm volatile("mul.lo.u32 %0, %1, 2;" : "=r"(l_) : "r"(a)); asm volatile("mul.hi.u32 %0, %1, 2;" : "=r"(h_) : "r"(a)); asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a) : "r"(l_), "r"(h_), "r"(b)); }
#define ROLW(a, b) { unsigned long long w_; uint32_t l_, h_; asm volatile("mul.wide.u32 %0, %1, 2;" : "=l"(w_) : "r"(a)); asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(l_), "=r"(h_) : "l"(w_)); asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(a) : "r"(l_), "r"(h_), "r"(b)); }
BODY(k_lop, X8(LOP))
BODY(k_shf, X8(SHF))
BODY(k_shl, X8(SHL))
BODY(k_mad, X8(MAD))
BODY(k_madk, X8(MADK))
BODY(k_mhi, X8(MHI))
BODY(k_mhik, X8(MHIK))
BODY(k_wide, X8(WIDE))
BODY(k_prmt, X8(PRM))
BODY(k_add, X8(ADD))
BODY(k_addk, X8(ADDK))
BODY(k_mnx, X8(MNX))
BODY(k_setp, X8(SETP))
BODY(k_lopmad, X8(LOPMAD))
BODY(k_lopmhi, X8(LOPMHI))
BODY(k_lop2mad, X8(LOP2MAD))
BODY(k_rola, X8(ROLA))
BODY(k_rolf, X8(ROLF))
BODY(k_rolw, X8(ROLW))
typedef void (*K)(uint32_t *, uint32_t, long long *);
int main()
{
    const int grid = 148 * 4;
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, grid * 8);
    struct { const char *name; K k; int per_chain; } ks[] = {
        {"LOP3 (xor)", k_lop, 1}, {"SHF (rotate)", k_shf, 1}, {"SHL by 1", k_shl, 1}, {"IMAD (mad.lo r,r,r)", k_mad, 1},
        {"IMAD (mad.lo r,imm)", k_madk, 1}, {"IMAD.HI (mul.hi)", k_mhi, 1}, {"IMAD.HI (mad.hi imm)", k_mhik, 1},
        {"IMAD.WIDE (mul.wide)", k_wide, 1}, {"PRMT", k_prmt, 1}, {"IADD (add r,r)", k_add, 1}, {"IADD imm (VIADD?)", k_addk, 1},
        {"VIMNMX", k_mnx, 1}, {"ISETP + @p add", k_setp, 2}, {"LOP3 + IMAD", k_lopmad, 2}, {"LOP3 + IMAD.HI", k_lopmhi, 2},
        {"LOP3 + SHF + IMAD", k_lop2mad, 3}, {"rol1^t: SHF+LOP3", k_rola, 2}, {"rol1^t: IMAD.SHL+IMAD.HI+LOP3", k_rolf, 3},
        {"rol1^t: IMAD.WIDE+LOP3", k_rolw, 2}};
    static long long h[grid];
    for (auto &e : ks) {
        e.k<<<grid, 256>>>(out, 3u, cyc);
        e.k<<<grid, 256>>>(out, 3u, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < grid; ++i) s += (double)h[i];
        const double cycles = s / grid;
        // per SM sub-partition: 4 CTAs x 8 warps / 4 = 8 warps, each NIT * 8 chains * per_chain instructions
        const double winst = 8.0 * NIT * 8 * e.per_chain;
        printf("%-34s %8.0f cycles  %.3f warp-instr/cycle/SMSP  (%.2f cycles per chain step per warp-slot)\n", e.name, cycles, winst / cycles,
               cycles / (8.0 * NIT * 8));
    }
    return 0;
}
