/* cpmask_probe.cu — a single byte-masked bulk store (cp.async.bulk ... .cp_mask, SASS UBLKCP.G.S.DST_G_BYTE_MASK), entirely
 * in bounds: 256 bytes of dynamic shared memory, one 16-byte block at offset 64, mask 0xFFC0 (bytes 6..15).  Prints the
 * destination; run it under `compute-sanitizer --tool memcheck` to see how the tool treats the instruction.
 * build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cpmask_probe tools/cpmask_probe.cu */
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(uint8_t *g, unsigned short m)
{
    extern __shared__ __align__(16) uint8_t sm[];
    sm[threadIdx.x] = (uint8_t)(threadIdx.x + 1);
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + 64);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;" ::"l"(g), "r"(s), "h"(m) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
int main()
{
    uint8_t *g, h[16];
    cudaMalloc(&g, 256);
    cudaMemset(g, 0xEE, 256);
    k<<<1, 256, 256>>>(g, 0xFFC0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, g, 16, cudaMemcpyDeviceToHost);
    printf("status %s; dst:", cudaGetErrorString(e));
    for (int i = 0; i < 16; i++) printf(" %02x", h[i]);
    printf("   (want ee x6 then 47 48 .. 50)\n");
    return 0;
}
