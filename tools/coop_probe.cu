// Hardware / driver behaviour probe: are cooperative launches gang-scheduled when several streams compete for the SMs?
// A kernel with an in-kernel grid barrier (atomic counter + spin) is only safe if ALL of its CTAs are co-resident.  With
// plain launches two such kernels on two streams can each hold half of the SMs and wait for the other half forever.
// Every spin here has a clock64 timeout, so nothing can hang: a timeout is reported as a failed barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/coop_probe tools/coop_probe.cu && ./tools/coop_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void spin_kernel(long long cycles) {
    extern __shared__ char smem[];
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (threadIdx.x == 0) smem[0] = 1;
}

// result[0] += 1 per CTA that passed the barrier, result[1] += 1 per CTA that timed out
__global__ void barrier_kernel(unsigned* counter, unsigned* result, long long timeout_cycles, long long work_cycles) {
    extern __shared__ char smem[];
    const long long tw = clock64();
    while (clock64() - tw < work_cycles) { }
    if (threadIdx.x == 0) {
        smem[0] = 1;
        __threadfence();
        atomicAdd(counter, 1u);
        const long long t0 = clock64();
        bool ok = false;
        while (clock64() - t0 < timeout_cycles) {
            if (*(volatile unsigned*)counter >= gridDim.x) { ok = true; break; }
            __nanosleep(100);
        }
        atomicAdd(&result[ok ? 0 : 1], 1u);
    }
}

static cudaError_t launch_barrier(bool coop, int grid, int smem, cudaStream_t st, unsigned* counter, unsigned* result,
                                  long long timeout, long long work) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = coop ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, barrier_kernel, counter, result, timeout, work);
}

int main() {
    int dev = 0, sms = 0, coop = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    const int SMEM = 200 * 1024;                       // one CTA per SM, like the convolution kernels
    CK(cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    CK(cudaFuncSetAttribute(barrier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    printf("SMs %d, cooperative launch supported %d\n", sms, coop);
    cudaStream_t sa, sb;
    CK(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
    unsigned *counters, *result;
    const int NCNT = 4096;
    CK(cudaMalloc(&counters, NCNT * sizeof(unsigned)));
    CK(cudaMalloc(&result, 2 * sizeof(unsigned)));
    const long long TIMEOUT = 40000000LL;              // ~20 ms
    auto reset = [&]() { CK(cudaMemset(counters, 0, NCNT * sizeof(unsigned))); CK(cudaMemset(result, 0, 2 * sizeof(unsigned))); CK(cudaDeviceSynchronize()); };
    auto report = [&](const char* what, int expect) {
        CK(cudaDeviceSynchronize());
        unsigned r[2];
        CK(cudaMemcpy(r, result, sizeof(r), cudaMemcpyDeviceToHost));
        printf("%-78s passed %u timed-out %u (of %d)\n", what, r[0], r[1], expect);
    };

    // 1. cooperative barrier kernel behind a kernel that occupies every SM on another stream
    reset();
    spin_kernel<<<sms, 128, SMEM, sa>>>(4000000LL);
    CK(launch_barrier(true, sms, SMEM, sb, counters, result, TIMEOUT, 0));
    report("1. cooperative, full grid, other stream busy with a full-GPU kernel:", sms);

    // 2. two cooperative full-grid barrier kernels on two streams, 200 rounds
    for (int c = 1; c >= 0; --c) {
        reset();
        int k = 0;
        for (int it = 0; it < 200; ++it) {
            CK(launch_barrier(c, sms, SMEM, sa, counters + k++, result, TIMEOUT, 20000));
            CK(launch_barrier(c, sms, SMEM, sb, counters + k++, result, TIMEOUT, 20000));
        }
        report(c ? "2. cooperative, two streams x 200 full-grid barrier kernels:" : "3. PLAIN launches, same pattern (expected to dead-lock -> time-outs):", 400 * sms);
    }

    // 4. cooperative, smaller grids that can co-reside (2 x 60 CTAs), 200 rounds
    reset();
    { int k = 0;
      for (int it = 0; it < 200; ++it) {
        CK(launch_barrier(true, 60, SMEM, sa, counters + k++, result, TIMEOUT, 20000));
        CK(launch_barrier(true, 100, SMEM, sb, counters + k++, result, TIMEOUT, 20000));
      } }
    report("4. cooperative, two streams x 200 (60 + 100 CTAs):", 200 * 160);

    // 5. the same two-stream pattern captured into a CUDA graph (fork/join) and replayed
    reset();
    {
        cudaStream_t cap;
        CK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
        cudaEvent_t fork, join;
        CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
        cudaGraph_t graph; cudaGraphExec_t exec;
        CK(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
        CK(cudaEventRecord(fork, cap));
        CK(cudaStreamWaitEvent(sb, fork, 0));
        int k = 0;
        for (int it = 0; it < 20; ++it) {
            CK(launch_barrier(true, sms, SMEM, cap, counters + k++, result, TIMEOUT, 20000));
            CK(launch_barrier(true, sms, SMEM, sb, counters + k++, result, TIMEOUT, 20000));
        }
        CK(cudaEventRecord(join, sb));
        CK(cudaStreamWaitEvent(cap, join, 0));
        cudaError_t e = cudaStreamEndCapture(cap, &graph);
        if (e != cudaSuccess) { printf("5. capture of cooperative launches failed: %s\n", cudaGetErrorString(e)); return 0; }
        e = cudaGraphInstantiate(&exec, graph, 0);
        if (e != cudaSuccess) { printf("5. instantiate failed: %s\n", cudaGetErrorString(e)); return 0; }
        CK(cudaGraphLaunch(exec, cap));
        report("5. cooperative launches in a captured two-branch graph (1 replay, 40 kernels):", 40 * sms);
        // counters are not reset between replays: a second replay would see counters >= grid at once, so time it only
        cudaEvent_t t0, t1; CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
        CK(cudaEventRecord(t0, cap));
        for (int r = 0; r < 10; ++r) CK(cudaGraphLaunch(exec, cap));
        CK(cudaEventRecord(t1, cap));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
        printf("   10 replays: %.3f ms (%.2f us per kernel)\n", ms, ms * 1000.f / 400.f);
    }

    // 6. cost of the barrier itself: 200 cooperative full-grid kernels on one stream vs 200 kernels without barrier
    reset();
    {
        cudaEvent_t t0, t1; CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
        CK(cudaEventRecord(t0, sa));
        for (int it = 0; it < 200; ++it) CK(launch_barrier(true, sms, SMEM, sa, counters + it, result, TIMEOUT, 0));
        CK(cudaEventRecord(t1, sa));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, t0, t1));
        CK(cudaEventRecord(t0, sa));
        for (int it = 0; it < 200; ++it) spin_kernel<<<sms, 128, SMEM, sa>>>(0);
        CK(cudaEventRecord(t1, sa));
        CK(cudaDeviceSynchronize());
        float ms2; CK(cudaEventElapsedTime(&ms2, t0, t1));
        printf("6. per kernel: cooperative + grid barrier %.2f us, plain empty kernel %.2f us\n", ms * 5.f, ms2 * 5.f);
    }
    return 0;
}
