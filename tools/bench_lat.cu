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

// The exchange pattern of the resident kernel: every thread of a few warps per SM pushes 8 records to
// scattered mailboxes (16-byte stores), then polls 8 records laid out [slot][thread] (coalesced 16-byte
// loads).  Cycles until the stores are issued, and until the loads that follow have answered.
// STORE: 0 st.relaxed.gpu.b128, 1 plain st.global.v4.u32, 2 st.global.cg.v4.u32
// LOAD : 0 ld.relaxed.gpu.b128, 1 ld.global.cg.v4.u32, 2 ld.volatile.global.v4.u32
template <int STORE, int LOAD, bool LOADS_FIRST>
__global__ void exchange(uint4* box, unsigned n_box, long long* out, unsigned* sink, int iters)
{
    unsigned const tid = threadIdx.x, nt = blockDim.x, gid = blockIdx.x * nt + tid;
    unsigned acc = 0;
    long long t_store = 0, t_load = 0;
    for (int it = 0; it < iters; ++it)
    {
        unsigned seed = (gid * 2654435761u) ^ (unsigned(it) * 40503u);
        uint4* dst[8];
        for (int e = 0; e < 8; ++e)
        {
            seed   = seed * 1664525u + 1013904223u;
            dst[e] = box + (seed >> 8) % n_box;                       // scattered
        }
        uint4 const* src = box + ((blockIdx.x * 8u) * nt + tid) % n_box; // [slot][thread]: + e * nt
        __syncthreads();
        long long t0 = clock64();
        uint4 got[8];
        auto loads = [&] {
#pragma unroll
            for (int e = 0; e < 8; ++e)
            {
                uint4 const* p = src + e * nt;
                if (LOAD == 0) { W w = ldb(p); got[e] = make_uint4((unsigned)w.lo, (unsigned)(w.lo >> 32), (unsigned)w.hi, (unsigned)(w.hi >> 32)); }
                if (LOAD == 1) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(got[e].x), "=r"(got[e].y), "=r"(got[e].z), "=r"(got[e].w) : "l"(p));
                if (LOAD == 2) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(got[e].x), "=r"(got[e].y), "=r"(got[e].z), "=r"(got[e].w) : "l"(p));
            }
        };
        auto stores = [&] {
#pragma unroll
            for (int e = 0; e < 8; ++e)
            {
                if (STORE == 0) stb(dst[e], W{gid, (unsigned long long)it});
                if (STORE == 1) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst[e]), "r"(gid), "r"(0u), "r"(unsigned(it)), "r"(0u));
                if (STORE == 2) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst[e]), "r"(gid), "r"(0u), "r"(unsigned(it)), "r"(0u));
            }
        };
        if (LOADS_FIRST) loads();
        stores();
        long long t1;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1));
        if (!LOADS_FIRST) loads();
#pragma unroll
        for (int e = 0; e < 8; ++e)
            acc += got[e].x ^ got[e].w;
        long long t2;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t2) : "r"(acc));
        t_store += t1 - t0;
        t_load += t2 - t1;
    }
    if (tid == 0)
    {
        out[2 * blockIdx.x]     = t_store / iters;
        out[2 * blockIdx.x + 1] = t_load / iters;
    }
    sink[gid] = acc;
}

template <int STORE, int LOAD, bool LOADS_FIRST>
static void run_exchange(char const* name, uint4* box, unsigned n_box, long long* out, unsigned* sink)
{
    for (int threads : {96, 192})
    {
        for (int rep = 0; rep < 2; ++rep)
        {
            exchange<STORE, LOAD, LOADS_FIRST><<<148, threads>>>(box, n_box, out, sink, 200);
            cudaDeviceSynchronize();
        }
        long long hc[296]; cudaMemcpy(hc, out, sizeof hc, cudaMemcpyDeviceToHost);
        double a = 0, b = 0; for (int i = 0; i < 148; ++i) { a += hc[2 * i]; b += hc[2 * i + 1]; }
        printf("exchange %-44s %3d threads/SM: %5.0f cycles to issue%s 8 scattered stores, %5.0f more until 8 coalesced loads answered\n",
               name, threads, a / 148, LOADS_FIRST ? " 8 loads +" : "", b / 148);
    }
}


// Round trip of a poll: LOADS strong 128-bit loads per thread, issued together, until all have answered.
// Three warps per SM poll; with `writers` three more warps of every SM keep storing (strong, scattered) into the
// mailboxes another SM polls, as the neighbours of a region do.  layout 0: mailboxes [slot][thread] (a warp's
// load covers four 128-byte lines, eight mailboxes of eight different writers per line ... as in the kernel);
// layout 1: one mailbox per 128-byte line.
template <int LOADS>
__global__ void poll_probe(uint4* box, long long* out, unsigned* sink, int iters, int writers, int layout)
{
    unsigned const tid = threadIdx.x, warp = tid / 32;
    unsigned const stride = layout ? 8u : 1u;                 // in 16-byte units
    unsigned const per_cta = 8u * 96u * stride;                // mailboxes of one CTA's pollers
    unsigned acc = 0;
    long long total = 0;
    if (warp < 3)
    {
        uint4 const* mine = box + blockIdx.x * per_cta + tid * stride;
        for (int it = 0; it < iters; ++it)
        {
            long long t0 = clock64();
            W got[LOADS];
#pragma unroll
            for (int e = 0; e < LOADS; ++e)
                got[e] = ldb(mine + e * 96u * stride);
#pragma unroll
            for (int e = 0; e < LOADS; ++e)
                acc += (unsigned)got[e].lo ^ (unsigned)(got[e].hi >> 32);
            long long t1;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "r"(acc));
            total += t1 - t0;
        }
        if (tid == 0)
            out[blockIdx.x] = total / iters;
    }
    else if (writers)
    {
        unsigned const victim = (blockIdx.x + 1 + warp) % gridDim.x;
        unsigned seed = blockIdx.x * 7919u + tid;
        for (int it = 0; it < iters * 2; ++it)
#pragma unroll
            for (int e = 0; e < 8; ++e)
            {
                seed = seed * 1664525u + 1013904223u;
                stb(box + victim * per_cta + ((seed >> 10) % (8u * 96u)) * stride, W{seed, (unsigned long long)it});
            }
    }
    sink[blockIdx.x * blockDim.x + tid] = acc;
}

template <int LOADS>
static void run_poll(uint4* box, long long* out, unsigned* sink)
{
    for (int layout = 0; layout < 2; ++layout)
        for (int writers = 0; writers < 2; ++writers)
        {
            for (int rep = 0; rep < 2; ++rep)
            {
                poll_probe<LOADS><<<148, 192>>>(box, out, sink, 300, writers, layout);
                cudaDeviceSynchronize();
            }
            long long hc[148]; cudaMemcpy(hc, out, sizeof hc, cudaMemcpyDeviceToHost);
            double m = 0; for (auto c : hc) m += double(c);
            printf("poll round trip, %d strong 128-bit loads per thread, %s, %s: %5.0f cycles\n", LOADS,
                   layout ? "one mailbox per 128-byte line" : "mailboxes [slot][thread]         ",
                   writers ? "neighbours writing" : "nobody writing    ", m / 148);
        }
}


// The same poll probe with other access flavours: SF 0 st.relaxed.gpu.b128, 1 st.global.v4.u32 (weak), 2 st.global.cg.v4.u32;
// LF 0 ld.relaxed.gpu.b128, 1 ld.global.cg.v4.u32, 2 ld.volatile.global.v4.u32.  `gap`: the writers pause that many
// nanoseconds between batches of 8 stores (0 = hammer), to see how the polls degrade with the write rate.
template <int LOADS, int SF, int LF>
__global__ void poll_flavours(uint4* box, long long* out, unsigned* sink, int iters, unsigned gap)
{
    unsigned const tid = threadIdx.x, warp = tid / 32;
    unsigned const per_cta = 8u * 96u;
    unsigned acc = 0;
    long long total = 0;
    if (warp < 3)
    {
        uint4 const* mine = box + blockIdx.x * per_cta + tid;
        for (int it = 0; it < iters; ++it)
        {
            long long t0 = clock64();
            uint4 got[LOADS];
#pragma unroll
            for (int e = 0; e < LOADS; ++e)
            {
                uint4 const* p = mine + e * 96u;
                if (LF == 0) { W w = ldb(p); got[e] = make_uint4((unsigned)w.lo, (unsigned)(w.lo >> 32), (unsigned)w.hi, (unsigned)(w.hi >> 32)); }
                if (LF == 1) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(got[e].x), "=r"(got[e].y), "=r"(got[e].z), "=r"(got[e].w) : "l"(p));
                if (LF == 2) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(got[e].x), "=r"(got[e].y), "=r"(got[e].z), "=r"(got[e].w) : "l"(p));
            }
#pragma unroll
            for (int e = 0; e < LOADS; ++e)
                acc += got[e].x ^ got[e].w;
            long long t1;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "r"(acc));
            total += t1 - t0;
        }
        if (tid == 0)
            out[blockIdx.x] = total / iters;
    }
    else
    {
        unsigned const victim = (blockIdx.x + 1 + warp) % gridDim.x;
        unsigned seed = blockIdx.x * 7919u + tid;
        int const rounds = gap ? iters / 2 : iters * 2;
        for (int it = 0; it < rounds; ++it)
        {
#pragma unroll
            for (int e = 0; e < 8; ++e)
            {
                seed = seed * 1664525u + 1013904223u;
                uint4* d = box + victim * per_cta + (seed >> 10) % per_cta;
                if (SF == 0) stb(d, W{seed, (unsigned long long)it});
                if (SF == 1) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(d), "r"(seed), "r"(0u), "r"(unsigned(it)), "r"(0u));
                if (SF == 2) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(d), "r"(seed), "r"(0u), "r"(unsigned(it)), "r"(0u));
            }
            if (gap)
                __nanosleep(gap);
        }
    }
    sink[blockIdx.x * blockDim.x + tid] = acc;
}

template <int LOADS, int SF, int LF>
static void run_flavours(char const* name, uint4* box, long long* out, unsigned* sink)
{
    for (unsigned gap : {0u, 2000u, 5000u})
    {
        for (int rep = 0; rep < 2; ++rep)
        {
            poll_flavours<LOADS, SF, LF><<<148, 192>>>(box, out, sink, 300, gap);
            cudaDeviceSynchronize();
        }
        long long hc[148]; cudaMemcpy(hc, out, sizeof hc, cudaMemcpyDeviceToHost);
        double m = 0; for (auto c : hc) m += double(c);
        printf("poll flavours, %d loads, %-40s writers pause %4u ns: %5.0f cycles\n", LOADS, name, gap, m / 148);
    }
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
    {
        unsigned const n_box = 1u << 20; // 16 MB of mailboxes, as config 3
        unsigned* sink; cudaMalloc(&sink, 4 * 148 * 256);
        long long* out2; cudaMalloc(&out2, 8 * 2 * 148);
        run_exchange<0, 0, false>("st.relaxed.gpu.b128 / ld.relaxed.gpu.b128", tab, n_box, out2, sink);
        run_exchange<1, 0, false>("st.global.v4 (weak) / ld.relaxed.gpu.b128", tab, n_box, out2, sink);
        run_exchange<2, 0, false>("st.global.cg.v4 / ld.relaxed.gpu.b128", tab, n_box, out2, sink);
        run_exchange<0, 1, false>("st.relaxed.gpu.b128 / ld.global.cg.v4", tab, n_box, out2, sink);
        run_exchange<1, 1, false>("st.global.v4 (weak) / ld.global.cg.v4", tab, n_box, out2, sink);
        run_exchange<0, 2, false>("st.relaxed.gpu.b128 / ld.volatile.v4", tab, n_box, out2, sink);
        run_exchange<0, 0, true>("loads first: st.relaxed.gpu / ld.relaxed.gpu", tab, n_box, out2, sink);
        run_exchange<1, 1, true>("loads first: st.global.v4 / ld.global.cg.v4", tab, n_box, out2, sink);
    }
    {
        uint4* big; cudaMalloc(&big, sizeof(uint4) * 148u * 8u * 96u * 8u); cudaMemset(big, 0, sizeof(uint4) * 148u * 8u * 96u * 8u);
        unsigned* sink; cudaMalloc(&sink, 4 * 148 * 256);
        long long* out3; cudaMalloc(&out3, 8 * 148);
        run_poll<1>(big, out3, sink);
        run_poll<2>(big, out3, sink);
        run_poll<4>(big, out3, sink);
        run_poll<8>(big, out3, sink);
    }
    {
        uint4* big; cudaMalloc(&big, sizeof(uint4) * 148u * 8u * 96u); cudaMemset(big, 0, sizeof(uint4) * 148u * 8u * 96u);
        unsigned* sink; cudaMalloc(&sink, 4 * 148 * 256);
        long long* out4; cudaMalloc(&out4, 8 * 148);
        run_flavours<4, 0, 0>("st.relaxed.gpu / ld.relaxed.gpu", big, out4, sink);
        run_flavours<4, 1, 0>("st.global.v4 (weak) / ld.relaxed.gpu", big, out4, sink);
        run_flavours<4, 2, 0>("st.global.cg.v4 / ld.relaxed.gpu", big, out4, sink);
        run_flavours<4, 0, 1>("st.relaxed.gpu / ld.global.cg.v4", big, out4, sink);
        run_flavours<4, 0, 2>("st.relaxed.gpu / ld.volatile.v4", big, out4, sink);
        run_flavours<4, 1, 1>("st.global.v4 (weak) / ld.global.cg.v4", big, out4, sink);
        run_flavours<8, 0, 0>("st.relaxed.gpu / ld.relaxed.gpu", big, out4, sink);
        run_flavours<8, 1, 1>("st.global.v4 (weak) / ld.global.cg.v4", big, out4, sink);
        run_flavours<8, 0, 2>("st.relaxed.gpu / ld.volatile.v4", big, out4, sink);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
