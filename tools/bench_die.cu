// Development micro-benchmark: which die an SM sits on, which die a 2 KB chunk of global memory is homed on, and
// what a mailbox hop (strong 128-bit store -> strong 128-bit poll on another SM) costs for every combination of
// writer die, reader die and home die.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bench_die.bin tools/bench_die.cu
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
struct W { unsigned long long lo, hi; };
__device__ __forceinline__ W ldb(void const* p){ W w; asm volatile("{\n\t.reg .b128 q;\n\tld.relaxed.gpu.global.b128 q, [%2];\n\tmov.b128 {%0, %1}, q;\n\t}" : "=l"(w.lo), "=l"(w.hi) : "l"(p) : "memory"); return w; }
__device__ __forceinline__ void stb(void* p, W w){ asm volatile("{\n\t.reg .b128 q;\n\tmov.b128 q, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], q;\n\t}" :: "l"(p), "l"(w.lo), "l"(w.hi) : "memory"); }
__device__ __forceinline__ unsigned smid(){ unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }

constexpr int kChunk = 2048;            // bytes
constexpr int kLines = kChunk / 128;    // 128-byte lines per chunk

// every chunk holds a cyclic pointer chain over its 16 lines (byte offsets from the buffer start in .lo)
__global__ void fill(char* buf, unsigned const* chunk_off, int n_chunks)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    for (int l = 0; l < kLines; ++l)
    {
        unsigned long long next = (unsigned long long)chunk_off[c] + 128ull * ((l * 5 + 3) % kLines);
        *reinterpret_cast<W*>(buf + chunk_off[c] + 128 * l) = W{next, 0ull};
    }
}
// one CTA per SM (dynamic shared memory forces it): lat[sm][chunk] = cycles per dependent strong load, chain inside the chunk
__global__ void probe(char const* buf, unsigned const* chunk_off, int n_chunks, int reps, float* lat, unsigned* sm_of_block)
{
    if (threadIdx.x != 0) return;
    unsigned const sm = smid();
    sm_of_block[blockIdx.x] = sm;
    for (int cc = 0; cc < n_chunks; ++cc)
    {
        int const c = (cc + blockIdx.x * 7) % n_chunks; // SMs do not all hit one chunk at the same time
        unsigned long long at = chunk_off[c];
        for (int i = 0; i < kLines; ++i) at = ldb(buf + at).lo; // warm: the chunk into the L2
        long long best = 1ll << 60;
        for (int r = 0; r < reps; ++r)
        {
            long long t0 = clock64();
            for (int i = 0; i < 32; ++i) at = ldb(buf + at).lo;
            long long t1 = clock64();
            best = min(best, t1 - t0);
        }
        lat[sm * n_chunks + c] = float(best) / 32.f + (at == 1ull ? 1.f : 0.f);
    }
}
// ping-pong: the CTA on SM a stores into fa and polls fb, the CTA on SM b polls fa and stores into fb
__global__ void pingpong(char* buf, unsigned fa, unsigned fb, unsigned sm_a, unsigned sm_b, int iters, long long* out, unsigned base)
{
    if (threadIdx.x != 0) return;
    unsigned const sm = smid();
    if (sm != sm_a && sm != sm_b) return;
    bool const me_a = sm == sm_a;
    long long t0 = clock64();
    for (unsigned i = base + 1; i <= base + (unsigned)iters; ++i)
    {
        if (me_a)
        {
            stb(buf + fa, W{i, i});
            while (ldb(buf + fb).hi != i) {}
        }
        else
        {
            while (ldb(buf + fa).hi != i) {}
            stb(buf + fb, W{i, i});
        }
    }
    long long t1 = clock64();
    if (me_a) out[0] = (t1 - t0) / iters;
}

int main()
{
    int const n_sm = 148, n_probe = 48, n_more = 2048;
    size_t const bytes = 256ull << 20;
    char* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    std::vector<unsigned> off(n_probe + n_more);
    unsigned seed = 12345u;
    for (int c = 0; c < n_probe; ++c) { seed = seed * 1664525u + 1013904223u; off[c] = (seed >> 4) % unsigned(bytes / kChunk) * kChunk; }
    for (int c = 0; c < n_more; ++c) off[n_probe + c] = (64u << 20) + unsigned(c) * kChunk; // a contiguous 4 MB run
    unsigned* d_off; cudaMalloc(&d_off, off.size() * 4); cudaMemcpy(d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice);
    int const n_all = int(off.size());
    fill<<<(n_all + 127) / 128, 128>>>(buf, d_off, n_all);
    float* d_lat; cudaMalloc(&d_lat, sizeof(float) * 256 * n_all); cudaMemset(d_lat, 0, sizeof(float) * 256 * n_all);
    unsigned* d_smb; cudaMalloc(&d_smb, 4 * n_sm);
    size_t const smem = 120 * 1024; // one CTA per SM
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(pingpong, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<<<n_sm, 32, smem>>>(buf, d_off, n_probe, 4, d_lat, d_smb);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("probe failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<float> lat(256 * n_all); cudaMemcpy(lat.data(), d_lat, lat.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<unsigned> smb(n_sm); cudaMemcpy(smb.data(), d_smb, 4 * n_sm, cudaMemcpyDeviceToHost);
    std::vector<int> sms;
    for (int b = 0; b < n_sm; ++b) sms.push_back(int(smb[b]));
    std::sort(sms.begin(), sms.end());
    sms.erase(std::unique(sms.begin(), sms.end()), sms.end());
    printf("probe of %d chunks from %zu distinct SMs (max smid %d): %.2f ms\n", n_probe, sms.size(), sms.back(), ms);
    printf("block -> smid of the first 24 blocks:"); for (int b = 0; b < 24; ++b) printf(" %u", smb[b]); printf("\n");
    // classify: per chunk threshold half way between the fastest and the slowest SM
    std::vector<std::vector<int>> bit(sms.size(), std::vector<int>(n_probe));
    double gap = 0, lo_mean = 0, hi_mean = 0;
    for (int c = 0; c < n_probe; ++c)
    {
        float mn = 1e9f, mx = 0;
        for (int s : sms) { mn = std::min(mn, lat[s * n_probe + c]); mx = std::max(mx, lat[s * n_probe + c]); }
        float const th = 0.5f * (mn + mx);
        double lo = 0, hi = 0; int nlo = 0, nhi = 0;
        for (size_t i = 0; i < sms.size(); ++i)
        {
            float const v = lat[sms[i] * n_probe + c];
            bit[i][c] = v > th;
            if (v > th) { hi += v; ++nhi; } else { lo += v; ++nlo; }
        }
        lo_mean += lo / std::max(1, nlo); hi_mean += hi / std::max(1, nhi); gap += mx - mn;
        if (c < 6) printf("chunk %d: min %.1f max %.1f cycles, %d near / %d far SMs\n", c, mn, mx, nlo, nhi);
    }
    printf("mean near %.1f far %.1f (max - min %.1f) cycles per strong 128-bit load\n", lo_mean / n_probe, hi_mean / n_probe, gap / n_probe);
    // side of an SM: its pattern against the pattern of the first SM, majority over the chunks
    std::vector<int> side(sms.size());
    int n1 = 0; double agree = 0;
    for (size_t i = 0; i < sms.size(); ++i)
    {
        int same = 0;
        for (int c = 0; c < n_probe; ++c) same += bit[i][c] == bit[0][c];
        side[i] = same * 2 < n_probe;
        n1 += side[i];
        agree += double(std::max(same, n_probe - same)) / n_probe;
    }
    printf("SM sides: %zu / %d, mean agreement of an SM's pattern with its side %.3f\n", sms.size() - n1, n1, agree / sms.size());
    printf("side by smid:"); for (size_t i = 0; i < sms.size(); ++i) printf("%s%d", i % 37 == 0 ? "\n  " : "", side[i]); printf("\n");
    // chunk home: near for the SMs of side 0 -> home 0
    std::vector<int> home(n_probe);
    int h1 = 0;
    for (int c = 0; c < n_probe; ++c)
    {
        int far0 = 0, n0 = 0;
        for (size_t i = 0; i < sms.size(); ++i) if (side[i] == 0) { far0 += bit[i][c]; ++n0; }
        home[c] = far0 * 2 > n0;
        h1 += home[c];
    }
    printf("chunk homes: %d on side 0, %d on side 1\n", n_probe - h1, h1);
    // the contiguous run: homes of consecutive 2 KB chunks, probed from one SM of each side
    {
        probe<<<n_sm, 32, smem>>>(buf, d_off + n_probe, n_more, 2, d_lat, d_smb);
        cudaDeviceSynchronize();
        // d_lat is indexed [sm * n_chunks + c] with n_chunks = n_more here
        std::vector<float> l2(256 * n_more); cudaMemcpy(l2.data(), d_lat, l2.size() * 4, cudaMemcpyDeviceToHost);
        int s0 = -1, s1 = -1;
        for (size_t i = 0; i < sms.size(); ++i) { if (side[i] == 0 && s0 < 0) s0 = sms[i]; if (side[i] == 1 && s1 < 0) s1 = sms[i]; }
        int runs = 1, ones = 0, undecided = 0; int prev = -1;
        printf("homes of 96 consecutive 2 KB chunks: ");
        for (int c = 0; c < n_more; ++c)
        {
            float const a = l2[s0 * n_more + c], b = l2[s1 * n_more + c];
            int const h = a > b;
            if (fabsf(a - b) < 8.f) ++undecided;
            ones += h;
            if (c > 0 && h != prev) ++runs;
            prev = h;
            if (c < 96) printf("%d", h);
        }
        printf("\n%d chunks: %d homed on side 1, %d runs, %d undecided (difference below 8 cycles)\n", n_more, ones, runs, undecided);
    }
    // ping-pong for the combinations
    long long* d_out; cudaMalloc(&d_out, 8);
    auto pick_sm = [&](int sd, int k) { for (size_t i = 0; i < sms.size(); ++i) if (side[i] == sd && k-- == 0) return sms[i]; return -1; };
    auto pick_chunk = [&](int h, int k) { for (int c = 0; c < n_probe; ++c) if (home[c] == h && k-- == 0) return int(off[c]); return -1; };
    unsigned base = 0;
    auto run = [&](char const* name, int a, int b, int fa, int fb) {
        double sum = 0; int const iters = 2000;
        for (int rep = 0; rep < 3; ++rep)
        {
            pingpong<<<n_sm, 32, smem>>>(buf, unsigned(fa), unsigned(fb) + 1024u, unsigned(a), unsigned(b), iters, d_out, base);
            base += iters;
            cudaDeviceSynchronize();
            long long o; cudaMemcpy(&o, d_out, 8, cudaMemcpyDeviceToHost);
            if (rep) sum += double(o);
        }
        printf("%-78s %6.0f cycles per round trip (2 hops)\n", name, sum / 2);
    };
    for (int k = 0; k < 2; ++k)
    {
        int const a0 = pick_sm(0, 3 + 20 * k), b0 = pick_sm(0, 40 + 11 * k), b1 = pick_sm(1, 5 + 30 * k);
        int const c0 = pick_chunk(0, 2 * k), c0b = pick_chunk(0, 2 * k + 1), c1 = pick_chunk(1, 2 * k), c1b = pick_chunk(1, 2 * k + 1);
        printf("SMs %d, %d on side 0, %d on side 1\n", a0, b0, b1);
        run("same die, both mailboxes homed on that die", a0, b0, c0, c0b);
        run("same die, both mailboxes homed on the other die", a0, b0, c1, c1b);
        run("same die, one mailbox homed here and one there", a0, b0, c0, c1);
        run("different dies, every mailbox homed on its reader's die", a0, b1, c1, c0);
        run("different dies, every mailbox homed on its writer's die", a0, b1, c0, c1);
        run("different dies, both mailboxes homed on one die", a0, b1, c0, c0b);
    }
    return 0;
}
