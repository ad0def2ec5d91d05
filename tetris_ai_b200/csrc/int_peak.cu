// int_peak.cu — integer-issue roofline micro-benchmark for B200 (sm_100a).
//
// The Tetris placement path is integer/bitwise work, so its roofline denominator is the SM's
// integer issue rate, which MEASURED_PEAKS.json does not contain.  This program
// measures, on the box it runs on, the sustained per-SM throughput (thread-ops / clk / SM) and the
// whole-chip rate (thread-ops / s) of every SASS op class the kernels lean on:
//   LOP3, IADD3, SHF, IMAD, PRMT, IMNMX, POPC, FLO, I2F.F64, DADD, DMUL, SHFL, REDUX, LDS
// plus a LOP3+IADD3+IMAD "mixed" stream (both integer pipes busy) that is used as INT_PEAK.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o bin/int_peak int_peak.cu
// Run:    bin/int_peak [out.json]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <string>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

constexpr int CHAINS = 8;      // independent dependency chains per thread (ILP)
constexpr int INNER  = 64;     // unrolled ops per chain per outer iteration

enum Op { OP_LOP3, OP_IADD3, OP_SHF, OP_IMAD, OP_PRMT, OP_IMNMX, OP_POPC, OP_FLO, OP_I2F64,
          OP_DADD, OP_DMUL, OP_SHFL, OP_REDUX, OP_LDS, OP_MIX, OP_COUNT };
static const char* OP_NAME[OP_COUNT] = { "LOP3", "IADD3", "SHF", "IMAD", "PRMT", "IMNMX", "POPC", "FLO",
  "I2F.F64", "DADD", "DMUL", "SHFL.BFLY", "REDUX.MAX", "LDS.32", "MIX(LOP3+IADD3+IMAD)" };
// ops counted per inner step per chain (MIX issues three)
static const int OP_WEIGHT[OP_COUNT] = {1,1,1,1,1,1,1,1,1,1,1,1,1,1,3};

template <int OP>
__global__ void __launch_bounds__(256) bench_kernel(uint32_t* out, long long* cycles, int outer, uint32_t seed)
{
    __shared__ uint32_t smem[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) smem[i] = (i * 37u + seed) & 1023u;
    __syncthreads();
    uint32_t x[CHAINS];
    double   d[CHAINS];
    uint32_t k1 = seed * 2654435761u + threadIdx.x, k2 = seed ^ 0x9e3779b9u;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = k1 + c * 977u; d[c] = 1.0 + 1e-9 * (double)(c + threadIdx.x); }
    long long t0 = clock64();
    for (int it = 0; it < outer; ++it) {
#pragma unroll
        for (int s = 0; s < INNER; ++s) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
                if (OP == OP_LOP3)       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                else if (OP == OP_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(k1));
                else if (OP == OP_SHF)   asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                else if (OP == OP_IMAD)  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                else if (OP == OP_PRMT)  asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                else if (OP == OP_IMNMX) asm volatile("max.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(k2 + s));
                else if (OP == OP_POPC)  asm volatile("popc.b32 %0, %0;" : "+r"(x[c]));
                else if (OP == OP_FLO)   asm volatile("bfind.u32 %0, %0;" : "+r"(x[c]));
                else if (OP == OP_I2F64) { double t; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t) : "r"(x[c])); x[c] = (uint32_t)__double2hiint(t) ^ (uint32_t)__double2loint(t); }
                else if (OP == OP_DADD)  asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(1.25));
                else if (OP == OP_DMUL)  asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(1.0000001));
                else if (OP == OP_SHFL)  asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(x[c]));
                else if (OP == OP_REDUX) asm volatile("redux.sync.max.u32 %0, %0, 0xffffffff;" : "+r"(x[c]));
                else if (OP == OP_LDS)   { uint32_t a = (uint32_t)__cvta_generic_to_shared(&smem[x[c] & 1023u]);
                                           asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[c]) : "r"(a)); }
                else if (OP == OP_MIX) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(k1));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(k1), "r"(k2));
                }
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= x[c] ^ (uint32_t)__double2loint(d[c]) ^ (uint32_t)__double2hiint(d[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double g_clock_hz = 1.965e9;
struct Result { double per_sm_clk; double chip_per_s; double ms; };

template <int OP>
Result run(int sms, int blocks_per_sm, int outer, uint32_t* d_out, long long* d_cyc)
{
    int grid = sms * blocks_per_sm, block = 256;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench_kernel<OP><<<grid, block>>>(d_out, d_cyc, outer / 8 + 1, 12345u);   // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f; std::vector<long long> cyc(grid); double best_cyc = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        bench_kernel<OP><<<grid, block>>>(d_out, d_cyc, outer, 777u + rep);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) {
            best = ms;
            CK(cudaMemcpy(cyc.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            double s = 0; for (auto c : cyc) s += (double)c; best_cyc = s / grid;
        }
    }
    double ops_per_thread = (double)outer * INNER * CHAINS * OP_WEIGHT[OP];
    double total = ops_per_thread * (double)grid * block;
    Result r;
    r.ms = best;
    r.chip_per_s = total / (best * 1e-3);
    // per-SM rate at the maximum SM clock (clock64 deltas per block under-count when blocks start staggered,
    // so the event time is the authority; nvidia-smi showed clocks.sm == clocks.max.sm right after the run)
    (void)best_cyc;
    r.per_sm_clk = r.chip_per_s / ((double)sms * g_clock_hz);
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    return r;
}

int main(int argc, char** argv)
{
    const char* out_path = argc > 1 ? argv[1] : nullptr;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    if (khz > 0) g_clock_hz = 1e3 * (double)khz;
    const int bps = 4;             // 4 x 256 threads = 32 warps / SM, 8 warps / SMSP
    const int outer = 400;
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * sms * bps * 256));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * bps));
    Result res[OP_COUNT];
    res[OP_LOP3]  = run<OP_LOP3 >(sms, bps, outer, d_out, d_cyc);
    res[OP_IADD3] = run<OP_IADD3>(sms, bps, outer, d_out, d_cyc);
    res[OP_SHF]   = run<OP_SHF  >(sms, bps, outer, d_out, d_cyc);
    res[OP_IMAD]  = run<OP_IMAD >(sms, bps, outer, d_out, d_cyc);
    res[OP_PRMT]  = run<OP_PRMT >(sms, bps, outer, d_out, d_cyc);
    res[OP_IMNMX] = run<OP_IMNMX>(sms, bps, outer, d_out, d_cyc);
    res[OP_POPC]  = run<OP_POPC >(sms, bps, outer / 4, d_out, d_cyc);
    res[OP_FLO]   = run<OP_FLO  >(sms, bps, outer / 4, d_out, d_cyc);
    res[OP_I2F64] = run<OP_I2F64>(sms, bps, outer / 4, d_out, d_cyc);
    res[OP_DADD]  = run<OP_DADD >(sms, bps, outer / 2, d_out, d_cyc);
    res[OP_DMUL]  = run<OP_DMUL >(sms, bps, outer / 2, d_out, d_cyc);
    res[OP_SHFL]  = run<OP_SHFL >(sms, bps, outer / 4, d_out, d_cyc);
    res[OP_REDUX] = run<OP_REDUX>(sms, bps, outer / 8, d_out, d_cyc);
    res[OP_LDS]   = run<OP_LDS  >(sms, bps, outer / 4, d_out, d_cyc);
    res[OP_MIX]   = run<OP_MIX  >(sms, bps, outer, d_out, d_cyc);

    std::string js = "{\n";
    char buf[512];
    snprintf(buf, sizeof buf, " \"gpu_name\": \"%s\", \"sms\": %d, \"attr_clock_khz\": %d,\n \"threads_per_sm\": %d, \"ilp_chains\": %d,\n \"ops\": {\n",
             prop.name, sms, khz, bps * 256, CHAINS);
    js += buf;
    for (int i = 0; i < OP_COUNT; ++i) {
        snprintf(buf, sizeof buf, "  \"%s\": {\"thread_ops_per_clk_per_sm\": %.2f, \"chip_thread_ops_per_s\": %.4e, \"ms\": %.3f}%s\n",
                 OP_NAME[i], res[i].per_sm_clk, res[i].chip_per_s, res[i].ms, i + 1 < OP_COUNT ? "," : "");
        js += buf;
        printf("%-22s %8.2f thread-ops/clk/SM   %.3e thread-ops/s chip   (%.3f ms)\n",
               OP_NAME[i], res[i].per_sm_clk, res[i].chip_per_s, res[i].ms);
    }
    // INT_PEAK: the best whole-chip integer rate observed (mixed stream uses both INT-capable pipes)
    double peak = res[OP_MIX].chip_per_s;
    for (int i : {OP_LOP3, OP_IADD3, OP_IMAD}) if (res[i].chip_per_s > peak) peak = res[i].chip_per_s;
    snprintf(buf, sizeof buf, " },\n \"int_peak_thread_ops_per_s\": %.4e,\n \"how\": \"8 independent chains/thread, 32 warps/SM, %d SMs, CUDA events, best of 3; per-SM rate = chip rate / (SMs x max SM clock)\"\n}\n", peak, sms);
    js += buf;
    printf("INT_PEAK = %.4e thread-ops/s\n", peak);
    if (out_path) { FILE* f = fopen(out_path, "w"); if (f) { fputs(js.c_str(), f); fclose(f); } }
    return 0;
}
