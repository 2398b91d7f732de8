/*
 * nsdg_halo.cuh -- halo exchange between the 2-D boxes of a partitioned domain, over NVLink peer memory.
 *
 * The reference has no counterpart: its dynamics is not MPI-parallel; only the box
 * decomposition itself is the reference's (run/partition.cdl, core/src/ModelMetadata.cpp:42-62).  Each box
 * carries a one-element overlap ring towards every neighbour.  What travels:
 *   - every subcycle: the CG node lines of u, v that the box does not own (2 lines from the left/bottom
 *     neighbour, 3 from the right/top one: the right/top box owns the shared boundary line);
 *   - every transport RK stage: the ring column/row of the advected DG field.
 * Mechanism: each box exposes one "arena" (per side: two parity slots of payload + one flag) through a CUDA
 * IPC handle.  The sender's kernel packs its lines straight into the RECEIVER's arena with plain stores
 * through the peer mapping, then releases a system-scope flag carrying the exchange epoch; the receiver's
 * kernel acquires the flag and unpacks.  Two parity slots make the scheme race-free without
 * acknowledgements: slot (epoch % 2) is rewritten only after the receiver has unpacked epoch - 2, which it
 * must have done before it could send the epoch - 1 message the sender waited for.  x-direction first, then y
 * over the full local width, so corner values arrive by two hops.
 */
#pragma once
#include "nsdg_state.cuh"

namespace nsdg {

constexpr int kHaloSides = 4; // 0 bottom, 1 right, 2 top, 3 left (enum nsdg_side)

//! layout of a box's arena (all in one cudaMalloc so that one IPC handle covers it)
struct HaloArenaLayout {
    size_t slotDoubles; //!< payload capacity of one parity slot
    __host__ __device__ size_t slotOffset(int side, int parity) const { return (size_t(side) * 2 + parity) * slotDoubles; }
    __host__ __device__ size_t flagsOffsetBytes() const { return size_t(kHaloSides) * 2 * slotDoubles * sizeof(double); }
    __host__ __device__ size_t totalBytes() const { return flagsOffsetBytes() + 256; }
};

__device__ __forceinline__ void stReleaseSys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ldAcquireSys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long globalTimerNs()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
//! how long a box waits for a neighbour's message before it gives up and raises the handle's halo error (the host
//! reports it after the next stream synchronisation: Handle::checkHaloError).  Long enough for first-launch skew
//! between ranks (module load, graph instantiation), short enough not to look like a hung GPU.
constexpr unsigned long long kHaloTimeoutNs = 30ull * 1000ull * 1000ull * 1000ull;

//! one direction of one exchange: which lines of which fields go where
struct HaloLineDesc {
    int nFields; //!< number of arrays (node fields: 2; DG field: ncomp planes)
    int nLines; //!< lines sent to this neighbour
    int lineLen; //!< entries per line
    long firstLine[3]; //!< index of the first entry of each line in the array
    long stride; //!< distance between consecutive entries of a line
};

struct HaloPushArgs {
    double* fields[8];
    size_t fieldPitch; //!< DG planes: distance between planes when fields[1..] are not given (0 = use fields[])
    HaloLineDesc send[kHaloSides];
    double* peerSlot[kHaloSides]; //!< mapped pointer to the neighbour's slot for THIS exchange's parity (nullptr: no neighbour)
    unsigned* peerFlag[kHaloSides];
    unsigned epoch[kHaloSides];
    int sideMask; //!< which sides take part in this phase (x: left|right, y: bottom|top)
    unsigned* done; //!< [kHaloSides] block-completion counters in MY memory (zero between launches)
};

constexpr int kHaloBlocksPerSide = 24;

//! grid = (blocks per side, 4 sides).  Packs the lines into the neighbour's arena; the last block of a side
//! to finish publishes the epoch.
__global__ void __launch_bounds__(256) halo_push_kernel(const __grid_constant__ HaloPushArgs a)
{
    const int side = blockIdx.y;
    if (!(a.sideMask & (1 << side)) || a.peerSlot[side] == nullptr)
        return;
    const HaloLineDesc& d = a.send[side];
    const long perField = long(d.nLines) * d.lineLen;
    const long total = perField * d.nFields;
    double* dst = a.peerSlot[side];
    for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
        const int f = int(i / perField);
        const long r = i % perField;
        const int line = int(r / d.lineLen);
        const long k = r % d.lineLen;
        const double* src = a.fieldPitch ? a.fields[0] + size_t(f) * a.fieldPitch : a.fields[f];
        dst[i] = src[d.firstLine[line] + k * d.stride];
    }
    __threadfence_system(); // my peer stores are visible system-wide before I report completion
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.done + side, 1u);
        if (prev == gridDim.x - 1) { // every block of this side has fenced its stores
            a.done[side] = 0;
            __threadfence_system();
            stReleaseSys(a.peerFlag[side], a.epoch[side]);
        }
    }
}

struct HaloUnpackArgs {
    double* fields[8];
    size_t fieldPitch;
    HaloLineDesc recv[kHaloSides];
    const double* mySlot[kHaloSides]; //!< my own arena slot of this parity (nullptr: no neighbour)
    const unsigned* myFlag[kHaloSides];
    unsigned epoch[kHaloSides];
    int sideMask;
    int* errorFlag; //!< set to 1 if a neighbour never showed up (spin timeout)
};

//! grid = (blocks per side, 4 sides).  Waits until the neighbour's message of this epoch has landed, then scatters it.
__global__ void __launch_bounds__(256) halo_unpack_kernel(const __grid_constant__ HaloUnpackArgs a)
{
    const int side = blockIdx.y;
    if (!(a.sideMask & (1 << side)) || a.mySlot[side] == nullptr)
        return;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        ok = 1;
        const unsigned long long t0 = globalTimerNs();
        // flags only ever grow: >= tolerates a neighbour that is already one phase ahead
        while (int(ldAcquireSys(a.myFlag[side]) - a.epoch[side]) < 0) {
            if (globalTimerNs() - t0 > kHaloTimeoutNs) { // wall clock, independent of the SM clock: the neighbour died; do not hang the GPU
                ok = 0;
                *a.errorFlag = 1;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (!ok)
        return;
    const HaloLineDesc& d = a.recv[side];
    const long perField = long(d.nLines) * d.lineLen;
    const long total = perField * d.nFields;
    const double* src = a.mySlot[side];
    for (long i = long(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += long(gridDim.x) * blockDim.x) {
        const int f = int(i / perField);
        const long r = i % perField;
        const int line = int(r / d.lineLen);
        const long k = r % d.lineLen;
        double* dstf = a.fieldPitch ? a.fields[0] + size_t(f) * a.fieldPitch : a.fields[f];
        dstf[d.firstLine[line] + k * d.stride] = __ldcv(src + i);
    }
}

} // namespace nsdg
