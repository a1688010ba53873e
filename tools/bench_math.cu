// Development micro-benchmark: cycles per Green projection for one warp alone on an SM (latency)
// and for many warps (throughput).  Positions and tet constants come from shared memory, as in the
// resident kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/bench_math tools/bench_math.cu
#include "../soft-body-simulator_b200/csrc/xpbd_kernels.cuh"
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
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
