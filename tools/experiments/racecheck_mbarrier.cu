// Control experiment for profiles/r2_compute_sanitizer.txt: a minimal, textbook-correct producer/consumer hand-off between two warps
// through an mbarrier (the producer writes shared memory, __syncwarp, one lane arrives with release semantics; the consumer waits with
// acquire semantics, then reads).  If compute-sanitizer --tool racecheck flags THIS, its reports about the same pattern in
// brick_query_kernel (producer warps -> consumer warps across the `full` barrier) say nothing about the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o racecheck_mbarrier racecheck_mbarrier.cu && compute-sanitizer --tool racecheck ./racecheck_mbarrier
#include <cstdint>
#include <cstdio>

__global__ void handoff(int* out, int rounds)
{
    __shared__ __align__(8) unsigned long long bar[2];      // full, empty
    __shared__ int data[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t full = (uint32_t)__cvta_generic_to_shared(&bar[0]), empty = (uint32_t)__cvta_generic_to_shared(&bar[1]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto wait = [](uint32_t b, uint32_t parity) {
        asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b), "r"(parity) : "memory");
    };
    int acc = 0;
    for (int r = 0; r < rounds; r++) {
        if (warp == 1) {                                     // producer
            if (r >= 1) wait(empty, (uint32_t)(r - 1) & 1u);
            data[lane] = r * 32 + lane;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full) : "memory");
        } else {                                             // consumer
            wait(full, (uint32_t)r & 1u);
            acc += data[(lane + 1) & 31];
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty) : "memory");
        }
    }
    if (warp == 0) out[lane] = acc;
}

int main()
{
    int* d;
    cudaMalloc(&d, 32 * sizeof(int));
    handoff<<<1, 64>>>(d, 8);
    int h[32];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long expect = 0;
    for (int r = 0; r < 8; r++) expect += r * 32 + 1;
    printf("lane 0 sum %d (expected %lld), %s\n", h[0], expect, cudaGetLastError() == cudaSuccess ? "ok" : "error");
    return 0;
}
