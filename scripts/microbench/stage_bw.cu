// Microbenchmark: how fast can a warp-strip kernel with the staging pattern of subcycle_strip_umevp pull plane data
// from HBM?  One warp walks R rows; per row it needs NP planes x 32 doubles (256 B per plane), staged one row ahead.
// Variants: 0 = cp.async.ca 8 B per lane (what the strip kernels do), 1 = cp.async.cg 16 B per lane (cooperative),
//           2 = cp.async.bulk 256 B per plane + mbarrier, 3 = plain LDG into registers (no staging).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/stage_bw scripts/microbench/stage_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int NP = 32; // planes per row
constexpr int WARPS = 4;

__device__ __forceinline__ void cpAsync8(void* s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(unsigned(__cvta_generic_to_shared(s))), "l"(g)); }
__device__ __forceinline__ void cpAsync16cg(void* s, const void* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(unsigned(__cvta_generic_to_shared(s))), "l"(g)); }
__device__ __forceinline__ void cpCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mbarInit(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(unsigned(__cvta_generic_to_shared(b))), "r"(n)); }
__device__ __forceinline__ void mbarExpect(uint64_t* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(unsigned(__cvta_generic_to_shared(b))), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbarWait(uint64_t* b, unsigned parity)
{
    asm volatile("{\n.reg .pred p;\nW%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D%=;\nbra W%=;\nD%=:\n}" ::"r"(unsigned(__cvta_generic_to_shared(b))), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulkCopy(void* s, const void* g, unsigned bytes, uint64_t* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(unsigned(__cvta_generic_to_shared(s))), "l"(g), "r"(bytes), "r"(unsigned(__cvta_generic_to_shared(b))) : "memory");
}

struct Stage { double buf[2][NP][32]; uint64_t bar[2]; uint64_t pad[2]; };

template <int MODE, bool STORE>
__global__ void __launch_bounds__(32 * WARPS) kern(const double* __restrict__ src, double* __restrict__ dst, size_t Npad, int nxs, int R, int nsx, int nsy, double* out)
{
    extern __shared__ __align__(128) unsigned char raw[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int w = blockIdx.x * WARPS + wib;
    if (w >= nsx * nsy) return;
    Stage& st = reinterpret_cast<Stage*>(raw)[wib];
    const int sx = w % nsx, sy = w / nsx;
    const int ey0 = sy * R, ey1 = ey0 + R;
    double acc = 0;
    if (MODE == 2) {
        if (lane == 0) { mbarInit(&st.bar[0], 1); mbarInit(&st.bar[1], 1); }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
    auto issue = [&](int row, int b) {
        if (row < ey1) {
            const size_t e0 = size_t(row) * nxs + 32 * sx;
            if (MODE == 0) {
#pragma unroll
                for (int p = 0; p < NP; ++p) cpAsync8(&st.buf[b][p][lane], src + size_t(p) * Npad + e0 + lane);
            } else if (MODE == 1) {
#pragma unroll
                for (int p = 0; p < NP; p += 2) { const int pp = p + (lane >> 4); cpAsync16cg(&st.buf[b][pp][2 * (lane & 15)], src + size_t(pp) * Npad + e0 + 2 * (lane & 15)); }
            } else if (MODE == 2) {
                if (lane == 0) mbarExpect(&st.bar[b], NP * 256);
                __syncwarp();
                for (int p = lane; p < NP; p += 32) bulkCopy(&st.buf[b][p][0], src + size_t(p) * Npad + e0, 256, &st.bar[b]);
            }
        }
        if (MODE < 2) cpCommit();
    };
    if (MODE < 3) issue(ey0, 0);
    unsigned phase[2] = { 0, 0 };
    for (int ey = ey0; ey < ey1; ++ey) {
        const int b = (ey - ey0) & 1;
        const size_t e = size_t(ey) * nxs + 32 * sx + lane;
        if (MODE < 3) issue(ey + 1, b ^ 1);
        if (MODE < 2) { cpWait<1>(); if (MODE == 1) __syncwarp(); }
        if (MODE == 2) { mbarWait(&st.bar[b], phase[b]); phase[b] ^= 1; }
        double v[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) v[p] = (MODE == 3) ? src[size_t(p) * Npad + e] : st.buf[b][p][lane];
#pragma unroll
        for (int p = 0; p < NP; ++p) acc += v[p];
        if (STORE) {
#pragma unroll
            for (int p = 0; p < 24; ++p) dst[size_t(p) * Npad + e] = v[p] * 0.999;
        }
        if (MODE == 1 || MODE == 2) __syncwarp();
    }
    if (MODE < 2) cpWait<0>();
    if (acc == 123.456) out[0] = acc;
}

template <int MODE, bool STORE> void run(const char* name, const double* src, double* dst, size_t Npad, int n, int R, double* out, int blocksPerSM)
{
    const int nsx = n / 32, nsy = n / R;
    const int nb = (nsx * nsy + WARPS - 1) / WARPS;
    size_t smem = sizeof(Stage) * WARPS;
    // pad dynamic smem so that exactly blocksPerSM blocks fit (227 KB usable)
    size_t want = (227 * 1024) / blocksPerSM - 1024;
    if (want > smem) smem = want;
    CK(cudaFuncSetAttribute(kern<MODE, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) kern<MODE, STORE><<<nb, 32 * WARPS, smem>>>(src, dst, Npad, n, R, nsx, nsy, out);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    const int it = 10;
    for (int i = 0; i < it; ++i) kern<MODE, STORE><<<nb, 32 * WARPS, smem>>>(src, dst, Npad, n, R, nsx, nsy, out);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= it;
    const double bytes = double(n) * n * 8 * (NP + (STORE ? 24 : 0));
    printf("%-34s blocks/SM %d warps/SM %2d  %.3f ms  %.0f GB/s\n", name, blocksPerSM, blocksPerSM * WARPS, ms, bytes / ms / 1e6);
}

int main()
{
    const int n = 2048, R = 16;
    const size_t Npad = size_t(n) * n + 1024 + 32;
    double *src, *dst, *out;
    CK(cudaMalloc(&src, NP * Npad * 8)); CK(cudaMalloc(&dst, 24 * Npad * 8)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(src, 0, NP * Npad * 8)); CK(cudaMemset(dst, 0, 24 * Npad * 8));
    for (int bps : { 1, 2, 3 }) {
        run<0, false>("cp.async.ca 8B  read-only", src, dst, Npad, n, R, out, bps);
        run<1, false>("cp.async.cg 16B read-only", src, dst, Npad, n, R, out, bps);
        run<2, false>("cp.async.bulk 256B read-only", src, dst, Npad, n, R, out, bps);
        run<3, false>("plain LDG read-only", src, dst, Npad, n, R, out, bps);
        run<0, true>("cp.async.ca 8B  read+write", src, dst, Npad, n, R, out, bps);
        run<1, true>("cp.async.cg 16B read+write", src, dst, Npad, n, R, out, bps);
        run<2, true>("cp.async.bulk 256B read+write", src, dst, Npad, n, R, out, bps);
        run<3, true>("plain LDG read+write", src, dst, Npad, n, R, out, bps);
    }
    return 0;
}
