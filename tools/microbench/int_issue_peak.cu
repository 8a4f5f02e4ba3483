/*
 * int_issue_peak.cu -- measured integer issue peaks of the GPU this runs on (SURVEY.md 8(d) asks for the number the
 * roofline of an integer-only path is quoted against; it is not in MEASURED_PEAKS.json).
 *
 * Every thread runs 8 independent dependency chains of one instruction kind (IMAD on the FMA pipe; IADD3, LOP3, SHF,
 * ISETP+SEL-free PRMT on the ALU pipe), or an alternating IMAD / IADD3 / LOP3 mix, 32 resident warps per SM
 * scheduler-quarter so latency is hidden and only the pipe width limits.  Output: one JSON object with lane-ops/s,
 * warp-instructions per cycle per SM (from clock64) and the SM clock observed during the runs.
 *
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_issue_peak int_issue_peak.cu && ./int_issue_peak
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

enum { OP_IMAD, OP_IADD3, OP_LOP3, OP_SHF, OP_PRMT, OP_MIX2, OP_MIX3, N_OPS };
static const char *kNames[N_OPS] = {"imad", "iadd3", "lop3", "shf", "prmt", "mix_imad_iadd3", "mix_imad_iadd3_lop3"};

template <int OP> __device__ __forceinline__ void step(unsigned &a, unsigned b, unsigned c) {
    if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == OP_IADD3) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a) : "r"(b), "r"(c));
    if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == OP_PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
}

template <int OP> __global__ void __launch_bounds__(1024) chain_kernel(unsigned *out, unsigned b, unsigned c, int iters, long long *cycles) {
    unsigned v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = threadIdx.x * 8 + k;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (OP == OP_MIX2) { if ((k & 1) == 0) step<OP_IMAD>(v[k], b, c); else step<OP_IADD3>(v[k], b, c); }
                else if (OP == OP_MIX3) { if (k % 4 < 2) step<OP_IMAD>(v[k], b, c); else if (k % 4 == 2) step<OP_IADD3>(v[k], b, c); else step<OP_LOP3>(v[k], b, c); }
                else step<OP>(v[k], b, c);
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc ^= v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> static void run(int sms, unsigned *d_out, long long *d_cyc, double *lane_ops_per_s, double *warp_inst_per_cyc_sm, double *mhz) {
    const int blocks = sms * 2, threads = 1024, iters = 4096; /* 64 warps per SM = 16 per scheduler */
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain_kernel<OP><<<blocks, threads>>>(d_out, 3u, 5u, 64, d_cyc); /* warm-up */
    cudaDeviceSynchronize();
    float best = 1e30f;
    long long cyc_best = 1;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        chain_kernel<OP><<<blocks, threads>>>(d_out, 3u, 5u, iters, d_cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        long long cyc[1024];
        cudaMemcpy(cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int k = 0; k < blocks; k++) if (cyc[k] > mx) mx = cyc[k];
        if (ms < best) { best = ms; cyc_best = mx; }
    }
    const double inst_per_thread = (double)iters * 16 * 8;
    const double lane_ops = inst_per_thread * blocks * threads;
    *lane_ops_per_s = lane_ops / (best * 1e-3);
    *warp_inst_per_cyc_sm = inst_per_thread * (2.0 * threads / 32) / (double)cyc_best; /* two blocks per SM */
    *mhz = (double)cyc_best / (best * 1e-3) / 1e6;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
    const int sms = prop.multiProcessorCount;
    unsigned *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, sizeof(unsigned) * sms * 2 * 1024);
    cudaMalloc(&d_cyc, sizeof(long long) * 1024);
    double ops[N_OPS], ipc[N_OPS], mhz[N_OPS];
    run<OP_IMAD>(sms, d_out, d_cyc, &ops[0], &ipc[0], &mhz[0]);
    run<OP_IADD3>(sms, d_out, d_cyc, &ops[1], &ipc[1], &mhz[1]);
    run<OP_LOP3>(sms, d_out, d_cyc, &ops[2], &ipc[2], &mhz[2]);
    run<OP_SHF>(sms, d_out, d_cyc, &ops[3], &ipc[3], &mhz[3]);
    run<OP_PRMT>(sms, d_out, d_cyc, &ops[4], &ipc[4], &mhz[4]);
    run<OP_MIX2>(sms, d_out, d_cyc, &ops[5], &ipc[5], &mhz[5]);
    run<OP_MIX3>(sms, d_out, d_cyc, &ops[6], &ipc[6], &mhz[6]);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"how\": \"8 independent chains per thread, 64 warps per SM, best of 5, CUDA events + clock64\", \"ops\": {", prop.name, sms);
    for (int k = 0; k < N_OPS; k++)
        printf("%s\"%s\": {\"lane_ops_per_s\": %.4e, \"warp_inst_per_cycle_per_sm\": %.3f, \"sm_mhz\": %.0f}", k ? ", " : "", kNames[k], ops[k], ipc[k], mhz[k]);
    double peak = 0;
    for (int k = 0; k < N_OPS; k++) if (ops[k] > peak) peak = ops[k];
    printf("}, \"peak_lane_ops_per_s\": %.4e}\n", peak);
    return 0;
}
