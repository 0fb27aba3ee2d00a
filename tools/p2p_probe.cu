/*
 * p2p_probe.cu — how fast can ONE kernel move bytes into a peer GPU's memory over NVLink 5 / NVSwitch?
 * (development tool behind the design of the fused compose + all-gather kernels)
 *
 * Single process, G GPUs with peer access.  Every GPU pushes `bytes` from its own HBM into a buffer on each of
 * the other GPUs at the same time (the all-gather traffic pattern), by one of:
 *   stg    : one thread = one 16-byte block, ld.global.nc -> st.global (to every peer), UNROLL blocks in flight
 *   bulk   : TMA — cp.async.bulk global->shared (own HBM), then cp.async.bulk shared->global to every peer,
 *            CHUNK-byte stages, two stages per CTA
 *   ce     : cudaMemcpyPeerAsync (copy engines), one stream per destination — the reference point
 * Prints GB/s received per GPU (max time over GPUs).
 *
 * build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/p2p_probe tools/p2p_probe.cu
 * run:   tools/p2p_probe [MB per shard, default 256]
 */
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int MAXG = 8;
struct Dst { uint8_t *p[MAXG]; int n; };

__global__ void __launch_bounds__(256) push_stg(const uint4 *__restrict__ src, Dst d, size_t nblk)
{
    constexpr int UNROLL = 4;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    for (size_t b0 = tid; b0 < nblk; b0 += nth * UNROLL) {
        uint4 v[UNROLL];
        #pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t b = b0 + u * nth;
            if (b < nblk)
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + b));
        }
        #pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const size_t b = b0 + u * nth;
            if (b < nblk)
                for (int q = 0; q < d.n; q++)
                    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(d.p[q] + b * 16), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w) : "memory");
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int CHUNK>
__global__ void __launch_bounds__(32) push_bulk(const uint8_t *__restrict__ src, Dst d, size_t bytes)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar[2];
    const size_t nchunk = (bytes + CHUNK - 1) / CHUNK;
    if (threadIdx.x != 0)
        return;
    for (int s = 0; s < 2; s++)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    uint32_t phase[2] = {0, 0};
    int it = 0;
    for (size_t c = blockIdx.x; c < nchunk; c += gridDim.x, it++) {
        const int s = it & 1;
        const size_t off = c * CHUNK;
        const uint32_t n = (uint32_t)((bytes - off) < CHUNK ? (bytes - off) : CHUNK);
        if (it >= 2)            /* the stores that read this stage two trips ago must have left shared memory */
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(n) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm + s * CHUNK)), "l"(src + off), "r"(n), "r"(smem_u32(&bar[s])) : "memory");
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(phase[s]) : "memory");
        } while (!done);
        phase[s] ^= 1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int q = 0; q < d.n; q++)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d.p[q] + off), "r"(smem_u32(sm + s * CHUNK)), "r"(n) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv)
{
    const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 256) << 20;
    int G = 0;
    CK(cudaGetDeviceCount(&G));
    if (G > MAXG) G = MAXG;
    if (G < 2) { printf("need >= 2 GPUs\n"); return 0; }
    std::vector<uint8_t *> src(G), dst(G);
    std::vector<cudaStream_t> st(G);
    std::vector<cudaEvent_t> e0(G), e1(G);
    std::vector<std::vector<cudaStream_t>> ce(G);
    for (int g = 0; g < G; g++) {
        CK(cudaSetDevice(g));
        for (int q = 0; q < G; q++)
            if (q != g) {
                int ok = 0;
                CK(cudaDeviceCanAccessPeer(&ok, g, q));
                if (!ok) { printf("no peer access %d->%d\n", g, q); return 0; }
                cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
                cudaGetLastError();
            }
        CK(cudaMalloc(&src[g], bytes));
        CK(cudaMalloc(&dst[g], bytes * G));                 /* slot q of dst[g] receives GPU q's shard */
        CK(cudaMemset(src[g], g + 1, bytes));
        CK(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking));
        CK(cudaEventCreate(&e0[g]));
        CK(cudaEventCreate(&e1[g]));
        ce[g].resize(G);
        for (int q = 0; q < G; q++)
            CK(cudaStreamCreateWithFlags(&ce[g][q], cudaStreamNonBlocking));
    }
    auto dsts = [&](int g, bool self) {
        Dst d{};
        for (int q = 0; q < G; q++)
            if (q != g || self)
                d.p[d.n++] = dst[q] + (size_t)g * bytes;
        return d;
    };
    auto run = [&](const char *name, int mode, int param, int reps) {
        float worst = 0;
        for (int rep = 0; rep < reps + 1; rep++) {           /* first repetition = warm-up */
            for (int g = 0; g < G; g++) {
                CK(cudaSetDevice(g));
                CK(cudaDeviceSynchronize());
            }
            for (int g = 0; g < G; g++) {
                CK(cudaSetDevice(g));
                Dst d = dsts(g, false);
                CK(cudaEventRecord(e0[g], st[g]));
                if (mode == 0) {
                    push_stg<<<148 * param, 256, 0, st[g]>>>((const uint4 *)src[g], d, bytes / 16);
                } else if (mode == 1) {
                    CK(cudaFuncSetAttribute(push_bulk<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16384));
                    push_bulk<16384><<<148 * param, 32, 2 * 16384, st[g]>>>(src[g], d, bytes);
                } else if (mode == 2) {
                    CK(cudaFuncSetAttribute(push_bulk<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096));
                    push_bulk<4096><<<148 * param, 32, 2 * 4096, st[g]>>>(src[g], d, bytes);
                } else {
                    for (int q = 0; q < d.n; q++)
                        CK(cudaMemcpyAsync(d.p[q], src[g], bytes, cudaMemcpyDeviceToDevice, st[g]));
                }
                CK(cudaGetLastError());
                CK(cudaEventRecord(e1[g], st[g]));
            }
            float mx = 0;
            for (int g = 0; g < G; g++) {
                CK(cudaSetDevice(g));
                CK(cudaEventSynchronize(e1[g]));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0[g], e1[g]));
                mx = ms > mx ? ms : mx;
            }
            if (rep)
                worst = mx > worst ? mx : worst;
        }
        const double recv = (double)bytes * (G - 1);
        printf("{\"probe\": \"%s\", \"param\": %d, \"gpus\": %d, \"shard_MB\": %zu, \"ms\": %.3f, \"recv_GBps_per_gpu\": %.1f}\n", name, param, G,
            bytes >> 20, worst, recv / (worst * 1e-3) / 1e9);
        fflush(stdout);
    };
    for (int p : {2, 4, 8, 16})
        run("stg_v4_unroll4", 0, p, 3);
    for (int p : {1, 2, 4, 6})
        run("tma_bulk_16K_x2", 1, p, 3);
    for (int p : {2, 4, 8, 16})
        run("tma_bulk_4K_x2", 2, p, 3);
    run("copy_engine", 3, 0, 3);
    /* verify one destination */
    CK(cudaSetDevice(0));
    std::vector<uint8_t> h(64);
    CK(cudaMemcpy(h.data(), dst[0] + bytes * 1 + 12345, 64, cudaMemcpyDeviceToHost));
    printf("{\"check\": \"dst[0] slot 1 byte = %d (want 2)\"}\n", h[0]);
    return 0;
}
