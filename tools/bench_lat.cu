// Development micro-benchmark: latency of the 128-bit strong (gpu-scope) load the exchange protocol
// polls with, of a plain load, and store -> remote-visible latency between two SMs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct W { unsigned long long lo, hi; };
__device__ __forceinline__ W ldb(void const* p){ W w; asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}" : "=l"(w.lo), "=l"(w.hi) : "l"(p) : "memory"); return w; }
__device__ __forceinline__ void stb(void* p, W w){ asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}" :: "l"(p), "l"(w.lo), "l"(w.hi) : "memory"); }
// dependent chain of loads through a pointer table
__global__ void chase(uint4* tab, int n, long long* out, int strong)
{
    if (threadIdx.x != 0) return;
    unsigned idx = blockIdx.x * 977u % n;
    long long t0 = clock64();
    for (int i = 0; i < 64; ++i)
    {
        if (strong) { W w = ldb(tab + idx); idx = (unsigned)w.lo; }
        else { uint4 v = __ldcg(tab + idx); idx = v.x; }
    }
    long long t1 = clock64();
    out[blockIdx.x] = (t1 - t0) / 64 + (idx == 0xffffffffu);
}
// ping-pong between block 0 and block b: round trip of store -> poll on another SM
__global__ void pingpong(uint4* flags, long long* out, int partner, int iters)
{
    if (threadIdx.x != 0) return;
    if (blockIdx.x != 0 && (int)blockIdx.x != partner) return;
    bool const me0 = blockIdx.x == 0;
    long long t0 = clock64();
    for (unsigned i = 1; i <= (unsigned)iters; ++i)
    {
        if (me0)
        {
            stb(flags, W{i, i});
            while (ldb(flags + 8).hi != i) {}
        }
        else
        {
            while (ldb(flags).hi != i) {}
            stb(flags + 8, W{i, i});
        }
    }
    long long t1 = clock64();
    if (me0) out[0] = (t1 - t0) / iters;
}
int main()
{
    int n = 1 << 20;
    uint4* tab; long long* out; cudaMalloc(&tab, sizeof(uint4) * n); cudaMalloc(&out, 8 * 256);
    uint4* h = new uint4[n];
    for (int i = 0; i < n; ++i) { unsigned nx = (unsigned)((i * 1103515245ull + 12345ull) % n); h[i] = make_uint4(nx, 0, 0, 0); }
    cudaMemcpy(tab, h, sizeof(uint4) * n, cudaMemcpyHostToDevice);
    for (int strong = 0; strong < 2; ++strong)
        for (int rep = 0; rep < 2; ++rep)
        {
            chase<<<148, 32>>>(tab, n, out, strong);
            cudaDeviceSynchronize();
            long long hc[148]; cudaMemcpy(hc, out, sizeof hc, cudaMemcpyDeviceToHost);
            double m = 0; long long mx = 0, mn = 1 << 30; for (auto c : hc) { m += c; mx = c > mx ? c : mx; mn = c < mn ? c : mn; }
            printf("%s load, dependent chain, 16 MB table: mean %.0f min %lld max %lld cycles\n", strong ? "strong.gpu b128" : "ld.cg", m / 148, mn, mx);
        }
    cudaMemset(tab, 0, 4096);
    for (int partner : {1, 2, 37, 74, 100, 147})
    {
        pingpong<<<148, 32>>>(tab, out, partner, 2000);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
        printf("ping-pong block 0 <-> block %d: %lld cycles per round trip (2 store->visible hops)\n", partner, c);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
