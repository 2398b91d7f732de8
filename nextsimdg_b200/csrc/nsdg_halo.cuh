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
 * IPC handle.  ONE kernel per phase (halo_exchange_kernel): its push blocks pack the box's lines straight into the
 * RECEIVER's arena with plain stores through the peer mapping, then release a system-scope flag carrying the
 * exchange epoch; its unpack blocks acquire the box's own flag and scatter what the neighbour delivered (the lines
 * sent and the lines received are disjoint).  The epoch lives in DEVICE memory and is advanced by the kernel
 * itself, so the launch arguments never change and the whole subcycle loop of a partitioned box replays as one
 * CUDA graph.  Two parity slots make the scheme race-free without acknowledgements: slot (epoch % 2) is rewritten
 * only by the kernel of epoch + 2, which starts after this box has waited for the neighbour's epoch + 1 message,
 * which the neighbour sent from a kernel that started after its unpack of this epoch had finished.  x-direction
 * first, then y over the full local width, so corner values arrive by two hops.
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

constexpr int kHaloBlocksPerSide = 24;

//! device-resident exchange state of a box (zero-initialised)
struct HaloDevState {
    unsigned epoch[kHaloSides]; //!< exchanges completed per side
    unsigned pushed[kHaloSides]; //!< push blocks of the running kernel that have fenced their stores
    unsigned finished[kHaloSides]; //!< blocks (push and unpack) of the running kernel that are done
};

struct HaloExchangeArgs {
    double* fields[8];
    size_t fieldPitch; //!< DG planes: distance between planes when fields[1..] are not given (0 = use fields[])
    HaloLineDesc send[kHaloSides], recv[kHaloSides];
    double* peerArena[kHaloSides]; //!< mapped base of the neighbour's arena (nullptr: no neighbour)
    double* myArena;
    HaloArenaLayout layout;
    int sideMask; //!< which sides take part in this phase (x: left|right, y: bottom|top)
    HaloDevState* state;
    int* errorFlag; //!< set to 1 if a neighbour never showed up (wall-clock timeout)
};

//! grid = (2 * kHaloBlocksPerSide, 4 sides): blocks [0, B) push to the neighbour across the side, blocks [B, 2B) wait for
//! and unpack what that neighbour pushed.  All blocks are resident at once (192 blocks of 256 threads), so the waiting
//! unpack blocks cannot starve the push blocks.
__global__ void __launch_bounds__(256) halo_exchange_kernel(const __grid_constant__ HaloExchangeArgs a)
{
    const int side = blockIdx.y;
    if (!(a.sideMask & (1 << side)) || a.peerArena[side] == nullptr)
        return;
    const int opposite = (side + 2) & 3;
    const unsigned epoch = a.state->epoch[side] + 1u; // not advanced before every block of this kernel has read it (see below)
    const int parity = int(epoch & 1u);
    const bool push = blockIdx.x < kHaloBlocksPerSide;
    const int b = push ? blockIdx.x : blockIdx.x - kHaloBlocksPerSide;
    if (push) {
        const HaloLineDesc& d = a.send[side];
        const long perField = long(d.nLines) * d.lineLen;
        const long total = perField * d.nFields;
        // my message lands in the neighbour's slot for ITS side facing me
        double* dst = a.peerArena[side] + a.layout.slotOffset(opposite, parity);
        for (long i = long(b) * blockDim.x + threadIdx.x; i < total; i += long(kHaloBlocksPerSide) * blockDim.x) {
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
            const unsigned prev = atomicAdd(a.state->pushed + side, 1u);
            if (prev == kHaloBlocksPerSide - 1) { // every push block of this side has fenced its stores
                __threadfence_system();
                unsigned* peerFlag = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(a.peerArena[side]) + a.layout.flagsOffsetBytes()) + opposite;
                stReleaseSys(peerFlag, epoch);
            }
        }
    } else {
        __shared__ int ok;
        if (threadIdx.x == 0) {
            ok = 1;
            const unsigned* myFlag = reinterpret_cast<const unsigned*>(reinterpret_cast<const unsigned char*>(a.myArena) + a.layout.flagsOffsetBytes()) + side;
            const unsigned long long t0 = globalTimerNs();
            // flags only ever grow: >= tolerates a neighbour that is already one phase ahead
            while (int(ldAcquireSys(myFlag) - epoch) < 0) {
                if (globalTimerNs() - t0 > kHaloTimeoutNs) { // wall clock: the neighbour died; do not hang the GPU
                    ok = 0;
                    *a.errorFlag = 1;
                    break;
                }
                __nanosleep(100);
            }
        }
        __syncthreads();
        if (ok) {
            const HaloLineDesc& d = a.recv[side];
            const long perField = long(d.nLines) * d.lineLen;
            const long total = perField * d.nFields;
            const double* src = a.myArena + a.layout.slotOffset(side, parity);
            for (long i = long(b) * blockDim.x + threadIdx.x; i < total; i += long(kHaloBlocksPerSide) * blockDim.x) {
                const int f = int(i / perField);
                const long r = i % perField;
                const int line = int(r / d.lineLen);
                const long k = r % d.lineLen;
                double* dstf = a.fieldPitch ? a.fields[0] + size_t(f) * a.fieldPitch : a.fields[f];
                dstf[d.firstLine[line] + k * d.stride] = __ldcv(src + i);
            }
        }
    }
    // the last block of this side to finish advances the epoch and re-arms the counters for the next kernel.  Every
    // block read `epoch` at its very start and no block can be the last one before all others have passed that point.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(a.state->finished + side, 1u);
        if (prev == 2 * kHaloBlocksPerSide - 1) {
            a.state->pushed[side] = 0;
            a.state->finished[side] = 0;
            a.state->epoch[side] = epoch;
            __threadfence();
        }
    }
}

} // namespace nsdg
