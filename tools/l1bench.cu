// tools/l1bench.cu -- how many L1 data-pipe wavefronts a warp-wide global access costs, by pattern (run under ncu:
// l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum / l1tex__t_requests / l1tex__t_sectors).  One kernel per pattern, each warp
// issues ITER loads of the same shape from an L1-resident 64 KiB window (so DRAM/L2 are not involved after the first touch).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ITER 256
// PATTERN: 0 = 8 B/lane contiguous (256 B); 1 = 8 B/lane, 8 rows x 32 B (row stride 2 KiB); 2 = 4 rows x 64 B; 3 = 2 rows x 128 B;
//          4 = 16 B/lane contiguous (512 B); 5 = 16 B/lane, 16 rows x 32 B; 6 = 16 B/lane, lane stride 32 B; 7 = 16 B/lane, lane stride 64 B;
//          8 = 16 B/lane uniform address; 9 = 16 B/lane, 8 distinct entries (lane & 7) contiguous; 10 = 8 B store, 8 rows x 32 B; 11 = 8 B store contiguous
template <int PATTERN>
__global__ void k(u64 *buf, u64 *out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u64 acc = 0;
    size_t off;   // in 8-byte words
    if (PATTERN == 0 || PATTERN == 11) off = lane;
    if (PATTERN == 1 || PATTERN == 10) off = (size_t)(lane >> 2) * 256 + (lane & 3);
    if (PATTERN == 2) off = (size_t)(lane >> 3) * 256 + (lane & 7);
    if (PATTERN == 3) off = (size_t)(lane >> 4) * 256 + (lane & 15);
    if (PATTERN == 4) off = lane * 2;
    if (PATTERN == 5) off = (size_t)(lane >> 1) * 256 + (lane & 1) * 2;
    if (PATTERN == 6) off = lane * 4;
    if (PATTERN == 7) off = lane * 8;
    if (PATTERN == 8) off = 0;
    if (PATTERN == 9) off = (lane & 7) * 2;
    u64 *p = buf + (size_t)w * 8192 + off;   // 64 KiB per warp window
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
        u64 *q = p + (i & 7) * 32;
        if (PATTERN >= 10) { asm volatile("st.global.u64 [%0], %1;" :: "l"(q), "l"(acc + i) : "memory"); }
        else if (PATTERN >= 4) { u64 a, b; asm volatile("ld.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(q) : "memory"); acc += a ^ b; }
        else { u64 a; asm volatile("ld.global.u64 %0, [%1];" : "=l"(a) : "l"(q) : "memory"); acc += a; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int P> void run(u64 *buf, u64 *out) { k<P><<<148, 128>>>(buf, out); cudaDeviceSynchronize(); }
int main() {
    u64 *buf, *out;
    cudaMalloc(&buf, 4 * 8192 * 8 + (1 << 20)); cudaMalloc(&out, 148 * 128 * 8);
    cudaMemset(buf, 1, 4 * 8192 * 8 + (1 << 20));
    run<0>(buf, out); run<1>(buf, out); run<2>(buf, out); run<3>(buf, out); run<4>(buf, out); run<5>(buf, out);
    run<6>(buf, out); run<7>(buf, out); run<8>(buf, out); run<9>(buf, out); run<10>(buf, out); run<11>(buf, out);
    printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
