// Peer-memory exchange over NVLink: the device half.
//
// Every rank (one process per GPU) owns a WINDOW in its own HBM, mapped into every other rank's address space through
// CUDA IPC.  A window holds one 64-bit arrival counter and two sets of `world` slots.  An exchange is: store my
// contribution into slot[rank] of EVERY rank's window (plain stores; remote ones travel over NVLink), make them visible
// system-wide, bump every rank's counter, then wait until my own counter shows that all ranks' blocks have arrived and
// combine the slots IN RANK ORDER (deterministic).  No NCCL call, no host round trip; consecutive exchanges alternate
// between the two slot sets, which is enough because nobody can finish exchange k+1 before everybody finished reading k.
#pragma once
#include <cstdint>
#include <cstdio>

namespace ndb {

constexpr int NDB_MAX_PEERS = 16;
constexpr int PEER_MAX_TILES = 1000;  // per-tile arrival counters for reductions that exchange tile by tile (reduce.cu)
constexpr size_t PEER_HEADER = 8192;  // bytes in front of the slots: counter 0 = whole-exchange arrivals, 1 + t = tile t

struct PeerArgs {
    char *win[NDB_MAX_PEERS];      // window base of every rank, as mapped in THIS process
    int rank, world;
    unsigned long long target;     // value my counter reaches when every block of every rank has arrived
    size_t set_off;                // byte offset of the slot set used by this exchange (from the window base)
    size_t slot_bytes;
};

__device__ __forceinline__ char *peer_slot(const PeerArgs &p, int window_of, int slot)
{
    return p.win[window_of] + p.set_off + (size_t)slot * p.slot_bytes;
}

// called by all threads of a block after its stores into the peers' slots
__device__ __forceinline__ void peer_arrive(const PeerArgs &p)
{
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < p.world)
        atomicAdd_system(reinterpret_cast<unsigned long long *>(p.win[threadIdx.x]), 1ull);
}

// one thread spins until counter `index` of MY window reaches `target`
__device__ __forceinline__ void peer_spin(const PeerArgs &p, int index, unsigned long long target)
{
    const unsigned long long *ctr = reinterpret_cast<const unsigned long long *>(p.win[p.rank]) + index;
    unsigned long long seen, t0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(ctr) : "memory");
        if (seen >= target) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > 20000000000ull) {  // 20 s: a peer never arrived; fail loudly instead of hanging the GPU
            printf("ndb peer exchange: rank %d timed out at %llu of %llu arrivals (counter %d)\n", p.rank, seen, target, index);
            __trap();
        }
        __nanosleep(200);
    }
}

// called by all threads of a block; returns once every rank's contribution is visible in MY window
__device__ __forceinline__ void peer_wait(const PeerArgs &p)
{
    if (threadIdx.x == 0) {
        const unsigned long long *ctr = reinterpret_cast<const unsigned long long *>(p.win[p.rank]);
        unsigned long long seen, t0, now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(ctr) : "memory");
            if (seen >= p.target) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > 20000000000ull) {  // 20 s: a peer never arrived; fail loudly instead of hanging the GPU
                printf("ndb peer exchange: rank %d timed out at %llu of %llu arrivals\n", p.rank, seen, p.target);
                __trap();
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
}

}  // namespace ndb
