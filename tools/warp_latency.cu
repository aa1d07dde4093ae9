// tools/warp_latency.cu -- dependent-chain latency (cycles) of the warp primitives the QLFC coder leans on,
// measured with ONE resident warp (the situation of a coder warp: nothing else hides latency).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/warp_latency tools/warp_latency.cu && gpurun_out/warp_latency
#include <cstdio>
#include <cuda_runtime.h>

#define REPS 256

template <int WHAT> __global__ void k(unsigned *out, long long *cyc, unsigned seed, int distinct)
{
    __shared__ unsigned sm[1024];
    const unsigned lane = threadIdx.x;
    for (int i = lane; i < 1024; i += 32) sm[i] = (i * 7 + 1) & 1023;
    __syncwarp();
    unsigned v = seed + lane, acc = 0;
    unsigned key = distinct >= 32 ? lane : lane % distinct;
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < REPS; ++r) {
        if (WHAT == 0) { v = __shfl_sync(0xffffffffu, v, (lane + 1) & 31) + 1; }
        if (WHAT == 1) { v = __match_any_sync(0xffffffffu, key + (v & 0x40000000u)) ; acc += v; v = acc & 0x40000000u; }
        if (WHAT == 2) { v = __reduce_max_sync(0xffffffffu, v) + lane; }
        if (WHAT == 3) { v = sm[v & 1023]; }
        if (WHAT == 4) { v = __ballot_sync(0xffffffffu, v & 1) + lane; }
        if (WHAT == 5) { v = (v * 4093u + 77u) >> 12; }
        if (WHAT == 6) { sm[lane] = v; __syncwarp(); v = sm[(lane + 1) & 31] + 1; __syncwarp(); }
        if (WHAT == 7) { if ((v & 31) == (unsigned)(r & 31)) v = v * 3 + 1; __syncwarp(); }
        if (WHAT == 8) { v = __reduce_add_sync(0xffffffffu, v) + lane; }
    }
    // ---- issue rate of a LONE warp (round 2: does "one instruction per 4-5 cycles whatever the dependency structure" hold?) ----
    // N independent multiply-add chains in one instruction stream: cycles per INSTRUCTION = t / (REPS * N)
    if (WHAT >= 100) {
        unsigned a0 = v, a1 = v + 1, a2 = v + 2, a3 = v + 3, a4 = v + 4, a5 = v + 5, a6 = v + 6, a7 = v + 7;
        t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < REPS; ++r) {
            a0 = a0 * 4093u + 77u;
            if (WHAT >= 102) { a1 = a1 * 4091u + 79u; }
            if (WHAT >= 104) { a2 = a2 * 4089u + 81u; a3 = a3 * 4087u + 83u; }
            if (WHAT >= 108) { a4 = a4 * 4085u + 85u; a5 = a5 * 4083u + 87u; a6 = a6 * 4081u + 89u; a7 = a7 * 4079u + 91u; }
        }
        acc = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    }
    long long t1 = clock64();
    out[lane] = v + acc;
    if (lane == 0) *cyc = t1 - t0;
}

template <int WHAT> static void run(const char *name, int distinct = 32)
{
    unsigned *out; long long *cyc, h;
    cudaMalloc(&out, 128); cudaMalloc(&cyc, 8);
    k<WHAT><<<1, 32>>>(out, cyc, 12345u, distinct);
    k<WHAT><<<1, 32>>>(out, cyc, 12345u, distinct);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f cycles per dependent op\n", name, (double)h / REPS);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<5>("IMAD + SHF (counter move)");
    run<0>("__shfl_sync");
    run<4>("__ballot_sync");
    run<2>("__reduce_max_sync");
    run<8>("__reduce_add_sync");
    run<3>("LDS (pointer chase)");
    run<6>("STS + syncwarp + LDS + syncwarp");
    run<7>("divergent if + syncwarp");
    run<1>("__match_any_sync, 32 distinct values", 32);
    run<1>("__match_any_sync, 16 distinct values", 16);
    run<1>("__match_any_sync, 8 distinct values", 8);
    run<1>("__match_any_sync, 4 distinct values", 4);
    run<1>("__match_any_sync, 1 distinct value", 1);
    // cycles per LOOP ITERATION below; divide by the number of chains for cycles per instruction
    run<100>("lone warp, 1 IMAD chain  (per iteration)");
    run<102>("lone warp, 2 independent IMAD chains");
    run<104>("lone warp, 4 independent IMAD chains");
    run<108>("lone warp, 8 independent IMAD chains");
    return 0;
}
