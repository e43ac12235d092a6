#!/usr/bin/env python
"""Experiment: rate of SM-issued stores into mapped pinned host memory over PCIe, 4 bytes per lane (128 B per warp instruction) versus
16 bytes per lane (512 B per warp instruction), streaming (.cs) stores, grid of persistent CTAs.  Decides whether the host-output path
of the query kernel should widen its stores."""
import torch
from torch.utils.cpp_extension import load_inline

src = r'''
#include <torch/extension.h>
#include <cuda_runtime.h>
__global__ void w4(int* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(out + i), "r"((int)i) : "memory");
}
__global__ void w16(int4* __restrict__ out, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        int v = (int)i;
        asm volatile("st.global.cs.v4.s32 [%0], {%1,%1,%1,%1};" ::"l"(out + i), "r"(v) : "memory");
    }
}
// warp-contiguous chunks written at random chunk positions (like the lists of warp tasks landing wherever the cursor put them)
__global__ void w4_chunks(int* __restrict__ out, long long n_chunks, int chunk_words) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = warp; c < n_chunks; c += n_warps) {
        const long long cc = (c * 2654435761ll) % n_chunks;
        int* dst = out + cc * chunk_words;
        for (int w = lane; w < chunk_words; w += 32) asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(dst + w), "r"(w) : "memory");
    }
}
__global__ void w16_chunks(int* __restrict__ out, long long n_chunks, int chunk_words) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long c = warp; c < n_chunks; c += n_warps) {
        const long long cc = (c * 2654435761ll) % n_chunks;
        int4* dst = reinterpret_cast<int4*>(out + cc * chunk_words);
        for (int w = lane; w < chunk_words / 4; w += 32) asm volatile("st.global.cs.v4.s32 [%0], {%1,%1,%1,%1};" ::"l"(dst + w), "r"(w) : "memory");
    }
}
void run(torch::Tensor host, int which, int grid) {
    int* p = nullptr;
    cudaHostGetDevicePointer((void**)&p, host.data_ptr(), 0);
    const long long n = host.numel();
    if (which == 0) w4<<<grid, 512>>>(p, n);
    else if (which == 1) w16<<<grid, 512>>>((int4*)p, n / 4);
    else if (which == 2) w4_chunks<<<grid, 512>>>(p, n / 1024, 1024);
    else w16_chunks<<<grid, 512>>>(p, n / 1024, 1024);
}
'''
mod = load_inline(name="zc_width", cpp_sources="void run(torch::Tensor host, int which, int grid);", cuda_sources=src, functions=["run"],
                  extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"], verbose=False)
n = 320 * 1024 * 1024          # 1.28 GB of int32
host = torch.empty(n, dtype=torch.int32).pin_memory()
dev = torch.empty(n, dtype=torch.int32, device="cuda")
names = {0: "4 B/lane linear", 1: "16 B/lane linear", 2: "4 B/lane, 4 KB chunks at scattered positions", 3: "16 B/lane, 4 KB chunks at scattered positions"}
for grid in (148, 148 * 4):
    for which in (0, 1, 2, 3):
        mod.run(host, which, grid)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mod.run(host, which, grid)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"grid {grid:4d}  {names[which]:48s} {ms:7.2f} ms  {n * 4 / ms / 1e6:6.1f} GB/s", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
e0.record()
host.copy_(dev, non_blocking=True)
e1.record()
e1.synchronize()
print(f"cudaMemcpyAsync D2H (copy engine)                          {e0.elapsed_time(e1):7.2f} ms  {n * 4 / e0.elapsed_time(e1) / 1e6:6.1f} GB/s")
