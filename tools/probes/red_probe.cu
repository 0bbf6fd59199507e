// tools/probes/red_probe.cu -- how fast can votes be accumulated with vector reductions?  (GPU box; measurement aid)
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/red_probe tools/probes/red_probe.cu && gpurun_out/red_probe
//
// V votes with uniformly random base cells in a G^3 grid, 8 corners each, workspace [G^3][8] floats (one 32-byte sector
// per voxel).  Variants differ in how the lanes of a warp instruction map onto sectors:
//   v0  one lane per vote: red.v4 + red.v2 per corner            (16 instructions per vote, 32 sectors per instruction)
//   v1  two lanes per vote, each a red.v4 on one half of a sector  (8 instructions, 16 sectors per instruction)
//   v2  four lanes per vote covering the corners z and z+1         (4 instructions, 8 x 64 contiguous bytes per instruction)
//   v3  like v0 but a single red.v4 per corner (4 of 6 channels): the cost of the second reduction of v0
// and a streaming pass: read the workspace, write 24 bytes per voxel (the write-out's traffic).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void red4(float *a, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void red2(float *a, float x, float y) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(a), "f"(x), "f"(y) : "memory");
}

__global__ void v0(const int *cell, int V, int G, float *work) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const long long c = cell[i];
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int d = 0; d < 2; d++) {
        float *s = work + 8 * (c + (long long)a * G * G + b * G + d);
        red4(s, 1.f, 2.f, 3.f, 4.f);
        red2(s + 4, 5.f, 6.f);
    }
}
__global__ void v3(const int *cell, int V, int G, float *work) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const long long c = cell[i];
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int d = 0; d < 2; d++)
        red4(work + 8 * (c + (long long)a * G * G + b * G + d), 1.f, 2.f, 3.f, 4.f);
}
__global__ void v1(const int *cell, int V, int G, float *work) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 1, h = t & 1;
    if (i >= V) return;
    const long long c = cell[i];
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) for (int d = 0; d < 2; d++)
        red4(work + 8 * (c + (long long)a * G * G + b * G + d) + 4 * h, 1.f, 2.f, 3.f, 4.f);
}
__global__ void v2(const int *cell, int V, int G, float *work) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 2, q = t & 3;
    if (i >= V) return;
    const long long c = cell[i];
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++)
        red4(work + 8 * (c + (long long)a * G * G + b * G) + 4 * q, 1.f, 2.f, 3.f, 4.f);   // q = 0,1: corner z; 2,3: corner z+1
}
__global__ void v4(const int *cell, int V, int G, float *work) {   // 8 lanes per vote: (b, d, half); 2 instructions (a = 0, 1)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 3, q = t & 7;
    if (i >= V) return;
    const long long c = cell[i];
    const int b = q >> 2, dq = q & 3;
    for (int a = 0; a < 2; a++)
        red4(work + 8 * (c + (long long)a * G * G + b * G) + 4 * dq, 1.f, 2.f, 3.f, 4.f);
}
__global__ void v5(const int *cell, int V, int G, float *work) {   // 16 lanes per vote: (a, b, d, half); 1 instruction
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 4, q = t & 15;
    if (i >= V) return;
    const long long c = cell[i];
    const int a = q >> 3, b = (q >> 2) & 1, dq = q & 3;
    red4(work + 8 * (c + (long long)a * G * G + b * G) + 4 * dq, 1.f, 2.f, 3.f, 4.f);
}
// like v5, but the vote data arrive by shuffles from the lane that "computed" the vote (the real kernel's structure)
__global__ void v6(const int *cell, int V, int G, float *work) {
    const int lane = threadIdx.x & 31;
    const int base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;
    const int mine = base + lane < V ? cell[base + lane] : -1;
    const int q = lane & 15, a = q >> 3, b = (q >> 2) & 1, dq = q & 3;
    const long long off = (long long)a * G * G + b * G;
    for (int j = 0; j < 32; j += 2) {
        const int c = __shfl_sync(0xffffffffu, mine, j + (lane >> 4));
        if (c >= 0) red4(work + 8 * (c + off) + 4 * dq, 1.f, 2.f, 3.f, 4.f);
    }
}
__global__ void stream_pass(const float4 *work, long long Gv, float4 *o1, float4 *o2) {
    // 4 voxels per thread: 8 x 16-byte loads, 6 x 16-byte stores
    const long long v = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (v >= Gv) return;
    float4 a[8];
    for (int j = 0; j < 8; j++) a[j] = __ldcs(work + 2 * v + j);
    float s = 0.f;
    for (int j = 0; j < 8; j++) s += a[j].x + a[j].w;
    __stcs(o1 + v / 4, make_float4(s, a[0].y, a[1].z, a[2].w));
    for (int j = 0; j < 5; j++) __stcs(o2 + (v / 4) * 5 + j, make_float4(a[j].x, a[j + 1].y, a[j + 2].z, s));
}
__global__ void write_only(long long Gv, float4 *o1, float4 *o2) {
    const long long v = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (v >= Gv) return;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    __stcs(o1 + v / 4, z);
    for (int j = 0; j < 5; j++) __stcs(o2 + (v / 4) * 5 + j, z);
}

int main(int argc, char **argv) {
    const int G = argc > 1 ? atoi(argv[1]) : 128, V = argc > 2 ? atoi(argv[2]) : 381529;
    const long long Gv = (long long)G * G * G;
    int *h = (int *)malloc(sizeof(int) * V), *d;
    srand(1);
    for (int i = 0; i < V; i++) {
        const int x = rand() % (G - 1), y = rand() % (G - 1), z = rand() % (G - 1);
        h[i] = (x * G + y) * G + z;
    }
    // half of the votes concentrated in 12 hot spots, like object centres
    for (int i = 0; i < V / 2; i++) {
        const int o = i % 12, cx = 20 + 7 * o, cy = 10 + 3 * o, cz = 100 - 6 * o;
        h[i] = ((cx + rand() % 5) * G + cy + rand() % 5) * G + cz + rand() % 5;
    }
    CK(cudaMalloc(&d, sizeof(int) * V));
    CK(cudaMemcpy(d, h, sizeof(int) * V, cudaMemcpyHostToDevice));
    float *work, *o1, *o2, *flush;
    CK(cudaMalloc(&work, Gv * 32));
    CK(cudaMalloc(&o1, Gv * 4));
    CK(cudaMalloc(&o2, Gv * 20));
    CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemset(work, 0, Gv * 32));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time_it = [&](const char *name, auto launch, bool cold) {
        float best = 1e9f, sum = 0.f;
        const int reps = 8;
        for (int r = 0; r < reps + 2; r++) {
            if (cold) CK(cudaMemsetAsync(flush, r, 256 << 20));
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 2) { best = ms < best ? ms : best; sum += ms; }
        }
        CK(cudaGetLastError());
        printf("%-34s %s  best %7.1f us  mean %7.1f us\n", name, cold ? "L2 cold" : "L2 warm", best * 1e3f, sum / reps * 1e3f);
    };
    printf("G=%d  votes=%d  workspace %.1f MB\n", G, V, Gv * 32 / 1e6);
    for (int cold = 0; cold < 2; cold++) {
        time_it("v0 lane/vote v4+v2 (16 instr)", [&] { v0<<<(V + 255) / 256, 256>>>(d, V, G, work); }, cold);
        time_it("v3 lane/vote v4 only (8 instr)", [&] { v3<<<(V + 255) / 256, 256>>>(d, V, G, work); }, cold);
        time_it("v1 2 lanes/vote v4 (8 instr)", [&] { v1<<<(2 * V + 255) / 256, 256>>>(d, V, G, work); }, cold);
        time_it("v2 4 lanes/vote v4 (4 instr)", [&] { v2<<<(4 * V + 255) / 256, 256>>>(d, V, G, work); }, cold);
        time_it("v4 8 lanes/vote v4 (2 instr)", [&] { v4<<<(8 * V + 255) / 256, 256>>>(d, V, G, work); }, cold);
        time_it("v5 16 lanes/vote v4 (1 instr)", [&] { v5<<<(16 * V + 255) / 256, 256>>>(d, V, G, work); }, cold);
        time_it("v6 16 lanes/vote, shuffled", [&] { v6<<<(V + 255) / 256, 256>>>(d, V, G, work); }, cold);
        time_it("stream: read 32 B + write 24 B / voxel", [&] { stream_pass<<<(unsigned)((Gv / 4 + 255) / 256), 256>>>((float4 *)work, Gv, (float4 *)o1, (float4 *)o2); }, cold);
        time_it("write only 24 B / voxel", [&] { write_only<<<(unsigned)((Gv / 4 + 255) / 256), 256>>>(Gv, (float4 *)o1, (float4 *)o2); }, cold);
    }
    return 0;
}
