// warp_ipc.cu — single-warp latency / issue-rate probe for B200 (one warp per SM, nothing else resident).
// For each op class and ILP (independent chains) it prints cycles per instruction of ONE warp, i.e. the
// dependent-issue latency (chains = 1) and the best single-warp issue rate (chains = 8).  The config-2
// workload (1000 games on 592 warp schedulers) runs ~1 warp per scheduler, so these numbers — not the
// all-warps throughput of int_peak.cu — bound its per-piece latency.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/warp_ipc warp_ipc.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { LOP3, IMAD, MIXED, POPC, FLO, DADD, DMUL, SHFL, REDUX, LDS, VOTE, NOPS };
static const char* NAME[NOPS] = {"LOP3", "IMAD", "LOP3+IMAD", "POPC", "FLO", "DADD", "DMUL", "SHFL", "REDUX", "LDS", "VOTE"};

template <int OP, int CH>
__global__ void k(long long* out, uint32_t seed, int iters)
{
    __shared__ uint32_t sm[256];
    sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 32] = threadIdx.x;
    __syncwarp();
    uint32_t x[CH]; double d[CH];
    uint32_t k1 = seed + threadIdx.x, k2 = seed * 77u;
#pragma unroll
    for (int c = 0; c < CH; ++c) { x[c] = k1 + c; d[c] = 1.0 + c; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 32; ++s) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                else if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                else if (OP == MIXED) { if ((s + c) & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                                        else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2)); }
                else if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
                else if (OP == FLO) asm volatile("bfind.u32 %0, %0;" : "+r"(x[c]));
                else if (OP == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(1.25));
                else if (OP == DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(1.0000001));
                else if (OP == SHFL) asm volatile("shfl.sync.idx.b32 %0, %0, 3, 0x1f, 0xffffffff;" : "+r"(x[c]));
                else if (OP == REDUX) asm volatile("redux.sync.max.u32 %0, %0, 0xffffffff;" : "+r"(x[c]));
                else if (OP == LDS) { uint32_t a = (uint32_t)__cvta_generic_to_shared(&sm[x[c] & 63u]); asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[c]) : "r"(a)); }
                else if (OP == VOTE) { uint32_t r; asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "=r"(r) : "r"(x[c])); x[c] = r | 1u; }
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= x[c] ^ (uint32_t)__double2loint(d[c]);
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
}

template <int OP, int CH>
double run(long long* d_out)
{
    const int iters = 200;
    k<OP, CH><<<148, 32>>>(d_out, 3u, 10);
    cudaDeviceSynchronize();
    k<OP, CH><<<148, 32>>>(d_out, 5u, iters);
    long long h[2];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    return (double)h[0] / ((double)iters * 32 * CH);
}

template <int OP>
void row(long long* d_out)
{
    printf("%-10s  ILP1 %6.2f   ILP2 %6.2f   ILP4 %6.2f   ILP8 %6.2f   cycles per warp-instruction\n", NAME[OP],
           run<OP, 1>(d_out), run<OP, 2>(d_out), run<OP, 4>(d_out), run<OP, 8>(d_out));
}

int main()
{
    long long* d_out; cudaMalloc(&d_out, 16);
    row<LOP3>(d_out); row<IMAD>(d_out); row<MIXED>(d_out); row<POPC>(d_out); row<FLO>(d_out); row<DADD>(d_out);
    row<DMUL>(d_out); row<SHFL>(d_out); row<REDUX>(d_out); row<LDS>(d_out); row<VOTE>(d_out);
    return 0;
}
