// peer.cuh — all-reduce of a few doubles INSIDE a kernel through peer-mapped mailboxes (NVLink), for volumes sharded over
// the GPUs of one box: no NCCL call and no extra launch in the epoch.
#pragma once
#include "common.cuh"

namespace trb {

// The last CTA of every rank's epoch kernel pushes its partial sums into every rank's mailbox over NVLink (peer stores),
// waits for the others' and adds them in rank order, so all ranks continue with identical values.
struct PeerExchange {
    double *mailbox[8];            // mailbox[r]: rank r's buffer mapped into this process, [2 parities][8 ranks][48] + 8 doubles
    int rank, world;               // world <= 1: off
    unsigned long long seq;        // sequence number of this epoch's exchange (>= 1, identical on every rank)
};
constexpr int kMailSlot = 48;      // 41 moments, flag (u64) at [47]
constexpr int kMailPoison = 2 * 8 * kMailSlot;   // own mailbox, after the slots: set once a peer failed to show up

// All-reduce of `count` (<= 47) doubles over the ranks through peer memory, executed by one warp.
// Push model: remote stores are posted over NVLink, every rank polls its OWN memory.  Two parities of slots: a rank
// can only be one epoch ahead of its slowest peer (it needs that peer's previous contribution to finish an epoch).
// A peer that never shows up (bounded spin, ~seconds) poisons the moments with NaN instead of hanging the GPU.
__device__ __forceinline__ void peer_allreduce(double *row, int count, const PeerExchange &x, int lane)
{
    const unsigned long long seq = x.seq;
    const size_t par = (size_t)(seq & 1ull) * 8 * kMailSlot;
    for (int r = 0; r < x.world; ++r) {
        volatile double *dst = x.mailbox[r] + par + (size_t)x.rank * kMailSlot;
        for (int v = lane; v < count; v += 32) dst[v] = row[v];
    }
    __threadfence_system();
    __syncwarp();
    if (lane < x.world) {
        unsigned long long *flag = reinterpret_cast<unsigned long long *>(x.mailbox[lane] + par + (size_t)x.rank * kMailSlot + (kMailSlot - 1));
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
    }
    double *mine = x.mailbox[x.rank] + par;
    // once a peer has timed out every later exchange gives up at once (NaN sums) instead of spinning for seconds per
    // queued epoch
    volatile unsigned long long *poison = reinterpret_cast<volatile unsigned long long *>(x.mailbox[x.rank] + kMailPoison);
    bool ok = *poison == 0ull;
    if (ok && lane < x.world) {
        const unsigned long long *flag = reinterpret_cast<const unsigned long long *>(mine + (size_t)lane * kMailSlot + (kMailSlot - 1));
        unsigned long long got = 0;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(flag) : "memory");
        } while (got != seq && ++spins < (1ll << 23));
        ok = got == seq;
    }
    ok = __all_sync(kFull, ok);
    if (!ok && lane == 0) *poison = 1ull;
    for (int v = lane; v < count; v += 32) {
        double t = 0.0;
        for (int r = 0; r < x.world; ++r) t += reinterpret_cast<volatile const double *>(mine)[(size_t)r * kMailSlot + v];
        row[v] = ok ? t : __longlong_as_double(0x7ff8000000000000ll);
    }
    __syncwarp();
}

}  // namespace trb
