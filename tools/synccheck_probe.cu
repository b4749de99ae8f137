// Which named-barrier patterns does compute-sanitizer --tool synccheck accept?  Each variant is its own kernel; run as
//   nvcc -arch=sm_100a -lineinfo -o /tmp/scp tools/synccheck_probe.cu && for v in 0 1 2 3 4 5; do compute-sanitizer --tool synccheck /tmp/scp $v; done
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ void bsync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void barrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// 0: two warps, 64-thread block, bar.sync 1,64 from two different program locations
__global__ void v0(int *out) {
    __shared__ int s[32];
    if (threadIdx.x < 32) { s[threadIdx.x] = threadIdx.x; bsync(1, 64); }
    else { bsync(1, 64); out[threadIdx.x - 32] = s[threadIdx.x - 32]; }
}
// 1: as 0 in a 96-thread block whose third warp exits at once
__global__ void v1(int *out) {
    __shared__ int s[32];
    if (threadIdx.x >= 64) return;
    if (threadIdx.x < 32) { s[threadIdx.x] = threadIdx.x; bsync(1, 64); }
    else { bsync(1, 64); out[threadIdx.x - 32] = s[threadIdx.x - 32]; }
}
// 2: as 1, the third warp waits on barrier 2 which warp 1 arrives at later
__global__ void v2(int *out) {
    __shared__ int s[32];
    if (threadIdx.x >= 64) { bsync(2, 64); out[threadIdx.x] = s[threadIdx.x - 64]; return; }
    if (threadIdx.x < 32) { s[threadIdx.x] = threadIdx.x; bsync(1, 64); }
    else { bsync(1, 64); out[threadIdx.x - 32] = s[threadIdx.x - 32]; barrive(2, 64); }
}
// 3: as 0 with a lane-0 branch right before the barrier (no __syncwarp)
__global__ void v3(int *out) {
    __shared__ int s[33];
    if (threadIdx.x < 32) { if (threadIdx.x == 0) s[32] = out[63]; s[threadIdx.x] = threadIdx.x; bsync(1, 64); }
    else { bsync(1, 64); out[threadIdx.x - 32] = s[threadIdx.x - 32] + s[32]; }
}
// 4: same program location for both warps (the barrier sits after the role branch), 96-thread block, third warp at barrier 2
__global__ void v4(int *out) {
    __shared__ int s[32];
    if (threadIdx.x >= 64) { bsync(2, 64); out[threadIdx.x] = s[threadIdx.x - 64]; return; }
    if (threadIdx.x < 32) s[threadIdx.x] = threadIdx.x;
    bsync(1, 64);
    if (threadIdx.x >= 32) { out[threadIdx.x - 32] = s[threadIdx.x - 32]; barrive(2, 64); }
}
// 5: as 2 with the barriers in a loop (the K1 pattern: many phases on one id)
__global__ void v5(int *out) {
    __shared__ int s[32];
    if (threadIdx.x >= 64) { bsync(2, 64); out[threadIdx.x] = s[threadIdx.x - 64]; return; }
    if (threadIdx.x < 32) { for (int k = 0; k < 8; k++) { bsync(1, 64); s[threadIdx.x] = k; bsync(1, 64); } }
    else { int acc = 0; for (int k = 0; k < 8; k++) { bsync(1, 64); bsync(1, 64); acc += s[threadIdx.x - 32]; } out[threadIdx.x - 32] = acc; barrive(2, 64); }
}
int main(int argc, char **argv) {
    int v = argc > 1 ? atoi(argv[1]) : 0, *out;
    cudaMalloc(&out, 4096);
    cudaMemset(out, 0, 4096);
    switch (v) {
    case 0: v0<<<1, 64>>>(out); break;
    case 1: v1<<<1, 96>>>(out); break;
    case 2: v2<<<1, 96>>>(out); break;
    case 3: v3<<<1, 64>>>(out); break;
    case 4: v4<<<1, 96>>>(out); break;
    default: v5<<<1, 96>>>(out); break;
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: %s\n", v, cudaGetErrorString(e));
    return e != cudaSuccess;
}
