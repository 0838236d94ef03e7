// Micro-benchmark of the synchronisation primitives the group PDHG kernel is built from (developer tool).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_sync tools/ubench_sync.cu && tools/ubench_sync
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ unsigned ld_acq(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

template <int V>
__global__ void __launch_bounds__(1024, 1) k(int iters, int G, double *buf, unsigned *bar, long long *out, int halo) {
    const int tid = threadIdx.x, rank = blockIdx.x % G;
    unsigned epoch = 0;
    double acc = 0.0;
    __shared__ double sh[4096];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (V == 0) { __syncthreads(); }
        if (V == 1) { asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
        if (V == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
        if (V == 3 || V == 4 || V == 5 || V == 7) {
            buf[(size_t)blockIdx.x * 1024 + tid] = acc + it;   // exchange store
            __syncthreads();
            if (tid == 0) { if (V == 4) __threadfence(); else asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
            asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
            if (V == 5 || V == 7) {                            // halo fetch from the next block of the group
                const int peer = (blockIdx.x / G) * G + (rank + 1) % G;
                for (int h = 0; h < halo; ++h) sh[tid + 1024 * h] = __ldcg(buf + (size_t)peer * 1024 + ((tid * 7 + h * 131) & 1023));
                __syncthreads();
                acc += sh[(tid * 5) & 1023];
            }
            if (V == 7) {   // second barrier so the peer's next store cannot overtake our read (as in the real kernel there is work between)
                asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
            }
        }
        if (V == 6 || V == 8) {   // global-counter barrier (GRID mode)
            buf[(size_t)blockIdx.x * 1024 + tid] = acc + it;
            __syncthreads();
            if (tid == 0) {
                epoch += G;
                if (V == 6) __threadfence(); else asm volatile("fence.acq_rel.gpu;" ::: "memory");
                atomicAdd(bar + blockIdx.x / G, 1u);
                while ((int)(ld_acq(bar + blockIdx.x / G) - epoch) < 0) {}
            }
            __syncthreads();
            const int peer = (blockIdx.x / G) * G + (rank + 1) % G;
            acc += __ldcg(buf + (size_t)peer * 1024 + tid);
        }
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 123.456) buf[0] = acc;
}

template <int V>
void run(const char *name, int G, bool cluster, int iters, int halo = 2) {
    double *buf; unsigned *bar; long long *out;
    CK(cudaMalloc(&buf, sizeof(double) * 1024 * 256)); CK(cudaMalloc(&bar, 1024)); CK(cudaMalloc(&out, 8 * 256));
    CK(cudaMemset(bar, 0, 1024)); CK(cudaMemset(buf, 0, sizeof(double) * 1024 * 256));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(bar, 0, 1024));
        CK(cudaEventRecord(e0));
        if (cluster) {
            if (G > 8) CK(cudaFuncSetAttribute(k<V>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(G); cfg.blockDim = dim3(1024);
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, k<V>, iters, G, buf, bar, out, halo));
        } else {
            void *args[] = {&iters, &G, &buf, &bar, &out, &halo};
            CK(cudaLaunchCooperativeKernel((void *)k<V>, dim3(G), dim3(1024), args, 0, 0));
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h; CK(cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost));
    printf("%-52s G=%3d  %8.1f ns/iter  %8.0f cycles/iter\n", name, G, ms * 1e6 / iters, (double)h / iters);
    cudaFree(buf); cudaFree(bar); cudaFree(out);
}

int main() {
    const int N = 20000;
    run<0>("__syncthreads", 1, true, N);
    for (int G : {2, 4, 8, 16}) run<1>("cluster barrier relaxed", G, true, N);
    for (int G : {2, 8, 16}) run<2>("cluster barrier release/acquire", G, true, N);
    for (int G : {2, 8, 16}) run<3>("store+sync+t0 fence.acq_rel.gpu+cluster relaxed", G, true, N);
    for (int G : {2, 8, 16}) run<4>("store+sync+t0 __threadfence+cluster relaxed", G, true, N);
    for (int G : {2, 8, 16}) run<5>("  ... + halo fetch 2/thread", G, true, N);
    for (int G : {16}) run<5>("  ... + halo fetch 4/thread", G, true, N, 4);
    for (int G : {2, 8, 16}) run<7>("  ... + halo fetch 2/thread + 2nd barrier", G, true, N);
    for (int G : {2, 16, 32, 74, 148}) run<6>("grid counter barrier (__threadfence) + 1 load", G, false, N);
    for (int G : {2, 16, 32, 74, 148}) run<8>("grid counter barrier (fence.acq_rel) + 1 load", G, false, N);
    return 0;
}
