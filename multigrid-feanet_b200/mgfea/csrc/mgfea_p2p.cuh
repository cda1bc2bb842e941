// Halo exchange / gather / all-reduce of the row-slab V-cycle over NVLink peer memory (no NCCL on the data path).
//
// Every rank's slab arrays and a small "mailbox" live in memory that the other ranks of the node have mapped
// (cudaIpc).  One exchange step is ONE kernel per rank:
//
//   push    the copy jobs store this rank's boundary rows straight into the neighbours' ghost rows (or its coarse rows /
//           partial sums into every peer) with 16-byte peer stores over NVLink;
//   signal  every CTA, once its own stores are done (barrier + one system-scope fence), increments one flag in each
//           target's mailbox; all ranks launch a step with the same number of CTAs (mgfea_xchg.grid);
//   wait    one thread spins (ld.acquire.sys) until the flags in ITS mailbox, written by the ranks that push to it,
//           have advanced by that number of CTAs.  The expected count lives in device memory and is advanced by the
//           kernel itself, so the whole step is CUDA-graph replayable.
//
// The kernel that follows on the stream therefore sees complete ghost rows.  Overwriting a neighbour's ghost rows is safe
// without a second handshake because of the ping-pong buffers: the array pushed in step s+1 is never the one a kernel
// between steps s and s+1 reads (see FEANet/distributed.py), and a rank cannot be more than one step ahead of a
// neighbour.  Waits are bounded (a dead peer must not hang the GPU box): on timeout the error word is set and the host
// raises.
#pragma once
#include <stdint.h>

#include "../../../include/mgfea.h"
#include "mgfea_ptx.cuh"

namespace mgfea {

struct XchgParams {
    mgfea_xchg x;
    long long timeout_clocks;
};

__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_sys_add_u32(unsigned int *p, unsigned int v) {
    asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int XCHG_THREADS = 256;

__global__ void __launch_bounds__(XCHG_THREADS) p2p_exchange_kernel(const XchgParams p) {
    const mgfea_xchg &x = p.x;
    if (x.ctl != nullptr && ld_volatile_s32(&x.ctl->done) != 0) return;  // converged: every rank skips the step
    if (x.mode & MGFEA_XCHG_PUSH) {
        // ---- copy jobs: 16-byte chunks, grid-stride within each job (coalesced peer stores)
        const long long gtid = (long long)blockIdx.x * XCHG_THREADS + threadIdx.x;
        const long long gstride = (long long)gridDim.x * XCHG_THREADS;
        for (int j = 0; j < x.njobs; ++j) {
            const uint4 *src = reinterpret_cast<const uint4 *>(x.src[j]);
            uint4 *dst = reinterpret_cast<uint4 *>(x.dst[j]);
            const long long nchunk = (long long)(x.bytes[j] >> 4);
            long long i = gtid;
            for (; i + 3 * gstride < nchunk; i += 4 * gstride) {  // 4 independent loads in flight per thread
                const uint4 a = __ldcg(src + i), b = __ldcg(src + i + gstride), c = __ldcg(src + i + 2 * gstride),
                            d = __ldcg(src + i + 3 * gstride);
                dst[i] = a;
                dst[i + gstride] = b;
                dst[i + 2 * gstride] = c;
                dst[i + 3 * gstride] = d;
            }
            for (; i < nchunk; i += gstride) dst[i] = __ldcg(src + i);
        }
        // ---- every CTA signals for itself: the barrier makes the CTA's stores precede thread 0's system-scope fence,
        // the fence makes them visible before the flag increment (no grid-wide ticket, one fence per CTA)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();
            for (int s = 0; s < x.nsignal; ++s) red_relaxed_sys_add_u32(x.signal[s], 1u);
        }
    }
    if ((!(x.mode & MGFEA_XCHG_WAIT) && x.nwait2 <= 0) || blockIdx.x != 0 || threadIdx.x != 0) return;
    // ---- CTA 0 waits for the ranks that push to this one: each of their `x.grid` CTAs increments the flag once
    const unsigned int expect = (x.mode & MGFEA_XCHG_WAIT) ? *x.seq + (unsigned int)x.grid : 0u;
    const bool dead = (x.err != nullptr) && (*x.err != 0);  // after one timeout do not wait again (fail fast on the host)
    const long long t0 = clock64();
    for (int w = 0; w < ((x.mode & MGFEA_XCHG_WAIT) ? x.nwait : 0) && !dead; ++w) {
        while ((int)(ld_acquire_sys_u32(x.wait[w]) - expect) < 0) {
            if (clock64() - t0 > p.timeout_clocks) {
                if (x.err) *x.err = 1 + w;
                break;
            }
        }
    }
    if (x.mode & MGFEA_XCHG_WAIT) *x.seq = expect;
    // ---- second flag set: raised once per launch by the neighbours' fused-push kernels
    if (x.nwait2 > 0) {
        const unsigned int expect2 = *x.seq2 + 1u;
        for (int w = 0; w < x.nwait2 && !dead; ++w) {
            while ((int)(ld_acquire_sys_u32(x.wait2[w]) - expect2) < 0) {
                if (clock64() - t0 > p.timeout_clocks) {
                    if (x.err) *x.err = 17 + w;
                    break;
                }
            }
        }
        *x.seq2 = expect2;
    }
    // ---- optional reduction of the slots the peers filled (fixed rank order: identical result on every rank)
    if (x.nred > 0 && x.red_dst != nullptr) {
        double s = 0.0;
        for (int i = 0; i < x.nred; ++i) s += __ldcg(reinterpret_cast<const double *>(
                                               reinterpret_cast<const unsigned char *>(x.red_src) + (size_t)i * x.red_stride));
        *x.red_dst = s;
        if (x.ctl != nullptr) {  // Multigrid.Solve's loop condition on the all-rank total (same bits on every rank)
            mgfea_ctl *c = x.ctl;
            const int cyc = c->cycle;
            if (x.hist != nullptr && cyc < x.hist_cap) x.hist[cyc] = s;
            c->cycle = cyc + 1;
            bool done = false;
            if (c->eps2 >= 0.0 && cyc + 1 >= c->min_cycles && s <= c->eps2) done = true;
            if (cyc + 1 >= c->max_cycles) done = true;
            if (!(s == s) || s > 1.7e308) done = true;
            if (done) c->done = 1;
            __threadfence();
        }
    }
}

}  // namespace mgfea
