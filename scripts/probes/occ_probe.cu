// Which feature pins a kernel to one CTA per SM on sm_100a?  Prints cudaOccupancyMaxActiveBlocksPerMultiprocessor for
// small kernels that differ in one feature each.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o occ_probe occ_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_plain(float* p) { p[threadIdx.x] = 1.f; }

__global__ void k_tmem(float* p) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(slot));
    p[threadIdx.x] = 1.f;
}

__global__ void k_cluster(float* p) {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    p[threadIdx.x] = (float)r;
}

__global__ void k_mbar_tma(float* p) {
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar)));
    __syncthreads();
    p[threadIdx.x] = 1.f;
}

__global__ void k_pdl(float* p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    p[threadIdx.x] = 1.f;
}

template <typename K>
void report(const char* name, K k) {
    printf("%-10s", name);
    for (int kb : {0, 32, 64, 96, 112}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        int n = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, 288, (size_t)kb * 1024);
        printf("  %3dK:%2d", kb, n);
    }
    printf("\n");
}

int main() {
    report("plain", k_plain);
    report("tmem", k_tmem);
    report("cluster", k_cluster);
    report("mbarrier", k_mbar_tma);
    report("pdl", k_pdl);
    return 0;
}
