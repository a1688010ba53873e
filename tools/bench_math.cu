// Development micro-benchmark: cycles per Green projection for one warp alone on an SM (latency)
// and for many warps (throughput).  Positions and tet constants come from shared memory, as in the
// resident kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/bench_math tools/bench_math.cu
#include "xpbd_math_pair.cuh"
#include <cstdio>
#include <vector>
using namespace sbsb200;

template <int VARIANT>
__global__ void k(Real4<float> const* pos, Real4<float> const* rec, long long* cycles, float* sink, int iters, float strain)
{
    __shared__ Real4<float> sx[4 * 256];
    __shared__ Real4<float> sr[3 * 256];
    int const tid = threadIdx.x;
    for (int k = 0; k < 4; ++k)
    {
        Real4<float> p = pos[k];
        p.x *= strain; p.y *= (2.f - strain) ; p.x += 0.001f * tid; p.y += 0.002f * (tid % 7) * k;
        sx[k * blockDim.x + tid] = p;
    }
    for (int k = 0; k < 3; ++k)
        sr[k * blockDim.x + tid] = rec[k];
    __syncthreads();
    float lambda = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
        Real4<float> p1 = sx[tid], p2 = sx[blockDim.x + tid], p3 = sx[2 * blockDim.x + tid], p4 = sx[3 * blockDim.x + tid];
        Real4<float> r0 = sr[tid], r1 = sr[blockDim.x + tid], r2 = sr[2 * blockDim.x + tid];
        Vec3<float> z{};
        float l = lambda;
        green_project_at<float, false>(p1, p2, p3, p4, z, z, z, z, r0, r1, r2, 384615.4f, 576923.1f, 39.0625f, 0.f, 0.0016f, l);
        // keep the strain alive: write back a damped update so every iteration does real work
        if (VARIANT == 0)
        {
            sx[tid] = p1; sx[blockDim.x + tid] = p2; sx[2 * blockDim.x + tid] = p3; sx[3 * blockDim.x + tid] = p4;
            lambda = l;
        }
        else
        { // discard the update (constant strain level)
            lambda += (l - lambda) * 1e-30f;
            sx[tid].w = p1.w + (p1.x - sx[tid].x) * 1e-30f;
        }
    }
    long long t1 = clock64();
    if (tid == 0)
        cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + tid] = lambda + sx[tid].x;
}


// Two tets per thread on packed pairs (xpbd_math_pair.cuh): same data flow as k<>, lane 1 uses the
// positions of thread tid + 1.  Cycles are per PAIR iteration; main() halves them.
template <int VARIANT>
__global__ void k_pair(Real4<float> const* pos, Real4<float> const* rec, long long* cycles, float* sink, int iters, float strain)
{
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    __shared__ Real4<float> sx[2 * 4 * 256];
    __shared__ Real4<float> sr[3 * 256];
    int const tid = threadIdx.x, nt = blockDim.x;
    for (int l = 0; l < 2; ++l)
        for (int k = 0; k < 4; ++k)
        {
            Real4<float> p = pos[k];
            p.x *= strain; p.y *= (2.f - strain); p.x += 0.001f * (tid + l); p.y += 0.002f * ((tid + l) % 7) * k;
            sx[(l * 4 + k) * nt + tid] = p;
        }
    for (int k = 0; k < 3; ++k)
        sr[k * nt + tid] = rec[k];
    __syncthreads();
    P2 lambda(0.f);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
        Real4Pair p1 = pair_of(sx[tid], sx[4 * nt + tid]), p2 = pair_of(sx[nt + tid], sx[5 * nt + tid]),
                  p3 = pair_of(sx[2 * nt + tid], sx[6 * nt + tid]), p4 = pair_of(sx[3 * nt + tid], sx[7 * nt + tid]);
        Real4Pair const r0 = pair_of(sr[tid], sr[tid]), r1 = pair_of(sr[nt + tid], sr[nt + tid]),
                        r2 = pair_of(sr[2 * nt + tid], sr[2 * nt + tid]);
        P2 l = lambda;
        if (!green_project_pair(p1, p2, p3, p4, r0, r1, r2, P2(384615.4f), P2(576923.1f), P2(39.0625f), 0.0016f, l))
        { // a lane off the fast route: both through the scalar code
            Real4<float> a1 = lane0(p1), a2 = lane0(p2), a3 = lane0(p3), a4 = lane0(p4);
            Real4<float> b1 = lane1(p1), b2 = lane1(p2), b3 = lane1(p3), b4 = lane1(p4);
            Vec3<float> const z{};
            float la = l.v.x, lb = l.v.y;
            green_project_at<float, false>(a1, a2, a3, a4, z, z, z, z, lane0(r0), lane0(r1), lane0(r2), 384615.4f, 576923.1f, 39.0625f, 0.f, 0.0016f, la);
            green_project_at<float, false>(b1, b2, b3, b4, z, z, z, z, lane1(r0), lane1(r1), lane1(r2), 384615.4f, 576923.1f, 39.0625f, 0.f, 0.0016f, lb);
            p1 = pair_of(a1, b1); p2 = pair_of(a2, b2); p3 = pair_of(a3, b3); p4 = pair_of(a4, b4);
            l = P2(la, lb);
        }
        if (VARIANT == 0)
        {
            sx[tid] = lane0(p1); sx[nt + tid] = lane0(p2); sx[2 * nt + tid] = lane0(p3); sx[3 * nt + tid] = lane0(p4);
            sx[4 * nt + tid] = lane1(p1); sx[5 * nt + tid] = lane1(p2); sx[6 * nt + tid] = lane1(p3); sx[7 * nt + tid] = lane1(p4);
            lambda = l;
        }
        else
        {
            lambda = lambda + (l - lambda) * P2(1e-30f);
            sx[tid].w = p1.w.v.x + (p1.x.v.x - sx[tid].x) * 1e-30f + (p1.x.v.y - sx[4 * nt + tid].x) * 1e-30f;
        }
    }
    long long t1 = clock64();
    if (tid == 0)
        cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * nt + tid] = lambda.v.x + lambda.v.y + sx[tid].x;
#endif
}

// the pair routine against the scalar one on the same inputs (max relative deviation of positions, lambda)
__global__ void k_check(Real4<float> const* pos, Real4<float> const* rec, float* out, float strain)
{
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    int const tid = threadIdx.x;
    Real4<float> a[4], b[4];
    for (int k = 0; k < 4; ++k)
    {
        Real4<float> p = pos[k];
        p.x *= strain; p.y *= (2.f - strain);
        a[k] = p; a[k].x += 0.001f * tid; a[k].y += 0.002f * (tid % 7) * k;
        b[k] = p; b[k].x -= 0.003f * tid; b[k].z += 0.001f * (tid % 5) * k;
    }
    Real4<float> const r0 = rec[0], r1 = rec[1], r2 = rec[2];
    Real4Pair p1 = pair_of(a[0], b[0]), p2 = pair_of(a[1], b[1]), p3 = pair_of(a[2], b[2]), p4 = pair_of(a[3], b[3]);
    P2 l(0.25f, -0.5f);
    bool const fast = green_project_pair(p1, p2, p3, p4, pair_of(r0, r0), pair_of(r1, r1), pair_of(r2, r2), P2(384615.4f),
                                         P2(576923.1f), P2(39.0625f), 0.0016f, l);
    Vec3<float> z{};
    float la = 0.25f, lb = -0.5f;
    green_project_at<float, false>(a[0], a[1], a[2], a[3], z, z, z, z, r0, r1, r2, 384615.4f, 576923.1f, 39.0625f, 0.f, 0.0016f, la);
    green_project_at<float, false>(b[0], b[1], b[2], b[3], z, z, z, z, r0, r1, r2, 384615.4f, 576923.1f, 39.0625f, 0.f, 0.0016f, lb);
    float dev = 0.f;
    Real4Pair const* pp[4] = {&p1, &p2, &p3, &p4};
    for (int k = 0; k < 4; ++k)
    {
        Real4<float> const u = lane0(*pp[k]), v = lane1(*pp[k]);
        dev = fmaxf(dev, fmaxf(fmaxf(fabsf(u.x - a[k].x), fabsf(u.y - a[k].y)), fabsf(u.z - a[k].z)));
        dev = fmaxf(dev, fmaxf(fmaxf(fabsf(v.x - b[k].x), fabsf(v.y - b[k].y)), fabsf(v.z - b[k].z)));
    }
    out[2 * tid]     = fast ? dev : -1.f; // -1: off the fast route, nothing to compare
    out[2 * tid + 1] = fmaxf(fabsf(l.v.x - la) / fmaxf(1e-30f, fabsf(la)), fabsf(l.v.y - lb) / fmaxf(1e-30f, fabsf(lb)));
#endif
}

// Pipe probes: OP issued from 8 independent accumulators, `warps` warps per SM, cycles per warp-instruction
// per sub-partition.  0 FFMA, 1 FFMA2, 2 FMUL, 3 FADD, 4 FADD2, 5 FFMA with a dependent chain (latency),
// 6 FFMA2 dependent chain
template <int OP>
__global__ void k_pipe(long long* cycles, float* sink, int iters, float seed)
{
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    float a[8];
    unsigned long long q[8];
    float const m = seed, c = seed * 0.5f;
    unsigned long long mm, cc;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    for (int i = 0; i < 8; ++i)
    {
        a[i] = seed + i + threadIdx.x;
        asm volatile("mov.b64 %0, {%1, %1};" : "=l"(q[i]) : "f"(a[i]));
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                int const j = (OP == 5 || OP == 6) ? 0 : i;
                if (OP == 0 || OP == 5) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(m), "f"(c));
                if (OP == 1 || OP == 6) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[j]) : "l"(mm), "l"(cc));
                if (OP == 2) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(m));
                if (OP == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(c));
                if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[j]) : "l"(cc));
            }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i)
    {
        float lo, hi;
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(q[i]));
        s += a[i] + lo + hi;
    }
    if (threadIdx.x == 0)
        cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
#endif
}

template <int OP>
static void pipe_probe(char const* name, long long* dc, float* ds)
{
    int const iters = 2000;
    for (int warps : {1, 4, 8, 16})
    {
        for (int rep = 0; rep < 2; ++rep)
        {
            k_pipe<OP><<<148, 32 * warps>>>(dc, ds, iters, 1.0000001f);
            cudaDeviceSynchronize();
        }
        long long hc[148];
        cudaMemcpy(hc, dc, sizeof hc, cudaMemcpyDeviceToHost);
        double mean = 0; for (auto c : hc) mean += double(c); mean /= 148;
        double const per_warp = mean / (iters * 32.0);                    // cycles per instruction of one warp
        double const per_smsp = per_warp / ((warps + 3) / 4);             // warps share 4 sub-partitions
        printf("pipe %-22s warps/SM %2d: %.2f cycles per warp-instruction, %.2f per sub-partition slot\n", name, warps, per_warp, per_smsp);
    }
}

int main()
{
    // one lattice tet (p3,p1,p4,p0 of a unit cell): rest positions and DmInv
    double x0[4][3] = {{0, 1, 0}, {1, 0, 0}, {0, 0, 1}, {0, 0, 0}};
    double m[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[3 * r + c] = x0[c][r] - x0[3][r];
    double det = m[0]*(m[4]*m[8]-m[5]*m[7]) - m[1]*(m[3]*m[8]-m[5]*m[6]) + m[2]*(m[3]*m[7]-m[4]*m[6]);
    double id = 1.0 / det;
    double inv[9] = {(m[4]*m[8]-m[5]*m[7])*id,(m[2]*m[7]-m[1]*m[8])*id,(m[1]*m[5]-m[2]*m[4])*id,(m[5]*m[6]-m[3]*m[8])*id,(m[0]*m[8]-m[2]*m[6])*id,(m[2]*m[3]-m[0]*m[5])*id,(m[3]*m[7]-m[4]*m[6])*id,(m[1]*m[6]-m[0]*m[7])*id,(m[0]*m[4]-m[1]*m[3])*id};
    std::vector<Real4<float>> hp(4), hr(3);
    for (int k = 0; k < 4; ++k) hp[k] = {float(x0[k][0]), float(x0[k][1]), float(x0[k][2]), 1.f};
    hr[0] = {float(inv[0]), float(inv[1]), float(inv[2]), float(inv[3])};
    hr[1] = {float(inv[4]), float(inv[5]), float(inv[6]), float(inv[7])};
    hr[2] = {float(inv[8]), float(det / 6), 0.f, 0.f};
    Real4<float>*dp, *dr; long long* dc; float* ds;
    cudaMalloc(&dp, 64); cudaMalloc(&dr, 48); cudaMalloc(&dc, 8 * 1024); cudaMalloc(&ds, 4 * 1024 * 256);
    cudaMemcpy(dp, hp.data(), 64, cudaMemcpyHostToDevice); cudaMemcpy(dr, hr.data(), 48, cudaMemcpyHostToDevice);
    int const iters = 200;
    for (float strain : {1.10f, 1.01f, 1.0f})
        for (int threads : {32, 128, 256})
            for (int variant : {1, 0})
            {
                for (int rep = 0; rep < 2; ++rep)
                {
                    if (variant == 0) k<0><<<148, threads>>>(dp, dr, dc, ds, iters, strain);
                    else k<1><<<148, threads>>>(dp, dr, dc, ds, iters, strain);
                    cudaDeviceSynchronize();
                }
                long long hc[148];
                cudaMemcpy(hc, dc, sizeof hc, cudaMemcpyDeviceToHost);
                double mean = 0; for (auto c : hc) mean += double(c); mean /= 148;
                printf("strain %.2f threads %3d %s: %.0f cycles/projection (per thread), %.2f cycles/projection/SM throughput\n",
                       strain, threads, variant ? "const-strain" : "relaxing   ", mean / iters, mean / iters / threads);
            }
    // ---- two tets per thread on packed pairs
    for (float strain : {1.10f, 1.01f, 1.0f})
        for (int threads : {32, 64, 128})
            for (int variant : {1, 0})
            {
                for (int rep = 0; rep < 2; ++rep)
                {
                    if (variant == 0) k_pair<0><<<148, threads>>>(dp, dr, dc, ds, iters, strain);
                    else k_pair<1><<<148, threads>>>(dp, dr, dc, ds, iters, strain);
                    cudaDeviceSynchronize();
                }
                long long hc[148];
                cudaMemcpy(hc, dc, sizeof hc, cudaMemcpyDeviceToHost);
                double mean = 0; for (auto c : hc) mean += double(c); mean /= 148;
                printf("PAIR strain %.2f threads %3d (= %3d tets in flight) %s: %.0f cycles/projection (per thread, pair/2), %.2f cycles/projection/SM throughput\n",
                       strain, threads, 2 * threads, variant ? "const-strain" : "relaxing   ", mean / iters / 2, mean / iters / threads / 2);
            }
    {
        float* dout; cudaMalloc(&dout, 2 * 128 * sizeof(float));
        for (float strain : {1.10f, 1.01f, 0.5f})
        {
            k_check<<<1, 128>>>(dp, dr, dout, strain);
            float ho[256]; cudaMemcpy(ho, dout, sizeof ho, cudaMemcpyDeviceToHost);
            float dx = 0, dl = 0; int slow = 0;
            for (int i = 0; i < 128; ++i)
            {
                if (ho[2 * i] < 0.f) { ++slow; continue; }
                dx = fmaxf(dx, ho[2 * i]); dl = fmaxf(dl, ho[2 * i + 1]);
            }
            printf("pair vs scalar, strain %.2f: max |dx| %.3e, max rel |dlambda| %.3e (%d of 128 threads off the fast route)\n", strain, dx, dl, slow);
        }
    }
    pipe_probe<0>("FFMA (3 registers)", dc, ds);
    pipe_probe<1>("FFMA2", dc, ds);
    pipe_probe<2>("FMUL", dc, ds);
    pipe_probe<3>("FADD", dc, ds);
    pipe_probe<4>("FADD2", dc, ds);
    pipe_probe<5>("FFMA dependent chain", dc, ds);
    pipe_probe<6>("FFMA2 dependent chain", dc, ds);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
