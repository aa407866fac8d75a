// Does a small kernel on a non-blocking stream run beside a kernel whose CTAs spin on a flag it is going to set?
// Variants: spinning kernel on the legacy stream / on a created stream; with / without a large dynamic shared-memory request;
// spinning CTAs on every SM / on a few.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(volatile int *flag, int nspin, unsigned long long timeout_ns, int *timed_out) {
    extern __shared__ char sm[];
    if (blockIdx.x < nspin && threadIdx.x == 0) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        while (*flag == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) { *timed_out = 1; break; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sm[0] = 1;
}
__global__ void setflag(int *flag) { if (threadIdx.x == 0 && blockIdx.x == 0) *flag = 1; }
int main() {
    int *flag, *to;
    cudaMalloc(&flag, 4); cudaMalloc(&to, 4);
    cudaStream_t a, b;
    cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    setflag<<<1, 32, 0, b>>>(flag); cudaDeviceSynchronize();
    for (int legacy = 0; legacy < 2; ++legacy)
        for (int smem = 0; smem < 2; ++smem)
            for (int nspin : {8, 148 * 4}) {
                cudaMemset(flag, 0, 4); cudaMemset(to, 0, 4); cudaDeviceSynchronize();
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaStream_t sa = legacy ? 0 : a;
                cudaEventRecord(e0, sa);
                spin<<<148 * 4, 128, smem ? 50 * 1024 : 64, sa>>>(flag, nspin, 2000000000ull, to);
                cudaEventRecord(e1, sa);
                setflag<<<32, 256, 0, b>>>(flag);
                cudaDeviceSynchronize();
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                int h = -1; cudaError_t ce = cudaMemcpy(&h, to, 4, cudaMemcpyDeviceToHost); if (ce != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(ce)); return 1; }
                printf("spin kernel on %s stream, %s smem, %d spinning CTAs: %.3f ms%s\n", legacy ? "legacy" : "created", smem ? "50 KB" : "no", nspin, ms, h ? "  TIMED OUT" : "");
            }
    // the pattern of the external publish: stream A: K1, record(ev), K2 (spins for the flag); stream B: wait(ev), setflag
    for (int legacy = 0; legacy < 2; ++legacy)
        for (int timing = 0; timing < 2; ++timing) {
            cudaMemset(flag, 0, 4); cudaMemset(to, 0, 4); cudaDeviceSynchronize();
            cudaEvent_t ev, e0, e1;
            cudaEventCreateWithFlags(&ev, timing ? cudaEventDefault : cudaEventDisableTiming);
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaStream_t sa = legacy ? 0 : a;
            int *dummy; cudaMalloc(&dummy, 4); cudaMemset(dummy, 0, 4);
            cudaEventRecord(e0, sa);
            setflag<<<1, 32, 0, sa>>>(dummy);                       // K1
            cudaEventRecord(ev, sa);
            cudaStreamWaitEvent(b, ev, 0);
            setflag<<<32, 256, 0, b>>>(flag);                       // the "publish"
            spin<<<148 * 4, 128, 50 * 1024, sa>>>(flag, 592, 2000000000ull, to);   // K2, enqueued after the publish like iteration j + 1
            cudaEventRecord(e1, sa);
            cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            int h = -1; cudaMemcpy(&h, to, 4, cudaMemcpyDeviceToHost);
            printf("pattern K1,record,[B: wait,set],K2 on %s stream, event %s timing: %.3f ms%s\n", legacy ? "legacy" : "created", timing ? "with" : "without", ms, h ? "  TIMED OUT" : "");
        }
    // the same pattern with a LONG K1 (2 ms): stream B's wait is evaluated before the event has completed, so B's channel blocks on
    // the semaphore; does it wake up while K2 is spinning on stream A?
    for (int legacy = 0; legacy < 2; ++legacy) {
        cudaMemset(flag, 0, 4); cudaMemset(to, 0, 4); cudaDeviceSynchronize();
        cudaEvent_t ev, e0, e1;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaStream_t sa = legacy ? 0 : a;
        int *never; cudaMalloc(&never, 4); cudaMemset(never, 0, 4);
        int *to2; cudaMalloc(&to2, 4); cudaMemset(to2, 0, 4);
        spin<<<148, 128, 64, sa>>>(never, 148, 2000000ull, to2);          // K1: every CTA spins for 2 ms
        cudaEventRecord(ev, sa);
        cudaStreamWaitEvent(b, ev, 0);
        setflag<<<32, 256, 0, b>>>(flag);                                  // the "publish"
        cudaEventRecord(e0, sa);
        spin<<<148 * 4, 128, 50 * 1024, sa>>>(flag, 592, 2000000000ull, to);   // K2 spins until the publish has run (or 2 s)
        cudaEventRecord(e1, sa);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        int h = -1; cudaMemcpy(&h, to, 4, cudaMemcpyDeviceToHost);
        printf("long K1 (2 ms), record, [B: wait, set], K2 on %s stream: K2 waited %.3f ms%s\n", legacy ? "legacy" : "created", ms, h ? "  TIMED OUT" : "");
    }
    return 0;
}
