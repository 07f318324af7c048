// simt.h -- a small SIMT emulator for the CPU-only suite: runs CUDA kernels (compiled as plain C++ with
// SMB_SIMT_EMUL defined) one CTA at a time, every thread of the CTA a cooperative fiber (ucontext), so that
// __syncthreads(), the *_sync warp collectives, shared memory and atomics behave as on the device for the
// purposes of a parity test.  Test infrastructure only; no performance model, no memory model (one OS
// thread, so every access is sequentially consistent).
//
//   smb_emu::launch(grid, block, dynamic_smem_bytes, [&] { kernel(args...); });
//
// Semantics that matter for the kernels under test:
//   * __syncthreads(): releases when every thread of the CTA that has not exited is waiting at it;
//   * __ballot_sync / __shfl_sync / __reduce_add_sync with the full mask: complete when every lane of the warp
//     that has not exited has arrived (the kernels keep their collectives warp-uniform); a lane only
//     sees the values the other lanes passed to the SAME collective call;
//   * a CTA whose fibers are all blocked without a release condition is reported as a deadlock.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <cuda_runtime.h>        // first: its structs have members named gridDim / blockDim, the macros below come after

#include <algorithm>
#include <functional>
#include <vector>

namespace smb_emu {

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
    Dim3() {}
    Dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
    Dim3(const dim3& d) : x(d.x), y(d.y), z(d.z) {}
};

struct Fiber {
    ucontext_t ctx;
    Dim3 tid;
    unsigned linear = 0;
    bool done = false;
    unsigned long bar_gen = 0;          // block-barrier generation this fiber waits for (0: not waiting)
};
struct WarpColl {
    uint32_t arrived = 0;
    uint64_t val[32];
    uint64_t snap[32];
    uint32_t snap_mask = 0;
    unsigned long gen = 1;
};
struct State {
    Dim3 grid, block, bid;
    std::vector<Fiber> fibers;
    std::vector<WarpColl> warps;
    std::vector<char> stacks;
    std::vector<unsigned char> smem;
    ucontext_t main_ctx;
    Fiber* cur = nullptr;
    unsigned long bar_gen = 1;
    unsigned bar_arrived = 0, live = 0;
    const std::function<void()>* body = nullptr;
    unsigned long switches = 0, progress = 0;          // progress: barriers released, collectives completed, exits
};
inline State& st() { static State s; return s; }

inline unsigned char* dyn_smem() { return st().smem.data(); }
inline void yield() { State& S = st(); ++S.switches; swapcontext(&S.cur->ctx, &S.main_ctx); }

inline uint32_t live_lanes(unsigned warp) {
    State& S = st();
    uint32_t m = 0;
    const unsigned n = (unsigned)S.fibers.size();
    for (unsigned l = 0; l < 32; ++l) { const unsigned t = warp * 32 + l; if (t < n && !S.fibers[t].done) m |= 1u << l; }
    return m;
}

inline void syncthreads() {
    State& S = st();
    const unsigned long my = S.bar_gen;
    S.cur->bar_gen = my;
    ++S.bar_arrived;
    while (S.bar_gen == my) {
        if (S.bar_arrived >= S.live) { S.bar_arrived = 0; ++S.bar_gen; ++S.progress; break; }
        yield();
    }
    S.cur->bar_gen = 0;
}

// one warp collective: every live lane deposits a value; returns after all of them did
inline const WarpColl& collective(uint64_t v) {
    State& S = st();
    const unsigned warp = S.cur->linear / 32, lane = S.cur->linear & 31;
    WarpColl& c = S.warps[warp];
    const unsigned long my = c.gen;
    c.val[lane] = v;
    c.arrived |= 1u << lane;
    while (c.gen == my) {
        const uint32_t need = live_lanes(warp);
        if ((c.arrived & need) == need) {
            memcpy(c.snap, c.val, sizeof c.val);
            c.snap_mask = c.arrived;
            c.arrived = 0;
            ++c.gen;
            ++S.progress;
            break;
        }
        yield();
    }
    return c;
}

inline void fiber_entry() {
    State& S = st();
    (*S.body)();
    S.cur->done = true;
    --S.live;
    ++S.progress;
    // a thread that leaves while others wait at the barrier counts as arrived (lenient, like returning early
    // from a kernel whose remaining threads still synchronise among themselves)
    swapcontext(&S.cur->ctx, &S.main_ctx);
}

inline void launch(Dim3 grid, Dim3 block, size_t smem_bytes, const std::function<void()>& body) {
    State& S = st();
    const unsigned nthreads = block.x * block.y * block.z;
    const size_t stack_bytes = 64 * 1024;
    S.grid = grid; S.block = block; S.body = &body;
    S.stacks.resize((size_t)nthreads * stack_bytes);
    S.smem.assign(smem_bytes + 64, 0xAB);                 // shared memory starts undefined: poison it
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                S.bid = Dim3(bx, by, bz);
                S.fibers.assign(nthreads, Fiber());
                S.warps.assign((nthreads + 31) / 32, WarpColl());
                S.bar_gen = 1; S.bar_arrived = 0; S.live = nthreads;
                std::fill(S.smem.begin(), S.smem.end(), (unsigned char)0xAB);
                for (unsigned t = 0; t < nthreads; ++t) {
                    Fiber& f = S.fibers[t];
                    f.linear = t;
                    f.tid = Dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = S.stacks.data() + (size_t)t * stack_bytes;
                    f.ctx.uc_stack.ss_size = stack_bytes;
                    f.ctx.uc_link = &S.main_ctx;
                    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
                }
                unsigned long idle_rounds = 0;
                while (S.live > 0) {
                    const unsigned long before = S.progress;
                    for (unsigned t = 0; t < nthreads && S.live > 0; ++t) {
                        Fiber& f = S.fibers[t];
                        if (f.done) continue;
                        if (f.bar_gen && f.bar_gen == S.bar_gen && S.bar_arrived < S.live) continue;   // still blocked
                        S.cur = &f;
                        swapcontext(&S.main_ctx, &f.ctx);
                    }
                    // progress = somebody finished or a barrier / collective completed; detect a stuck CTA
                    if (S.live > 0 && S.bar_arrived >= S.live) { S.bar_arrived = 0; ++S.bar_gen; ++S.progress; }
                    if (S.progress == before) ++idle_rounds; else idle_rounds = 0;
                    if (idle_rounds > 4) { fprintf(stderr, "simt: deadlock in block (%u,%u,%u)\n", bx, by, bz); exit(97); }
                }
            }
}

}  // namespace smb_emu

// ---- the CUDA names the kernels use -------------------------------------------------------------------------
#define threadIdx (smb_emu::st().cur->tid)
#define blockIdx (smb_emu::st().bid)
#define blockDim (smb_emu::st().block)
#define gridDim (smb_emu::st().grid)
#define __syncthreads() smb_emu::syncthreads()
#undef __launch_bounds__
#define __launch_bounds__(...)

inline uint32_t lane_id() { return smb_emu::st().cur->linear & 31u; }
inline uint64_t ld_nc_u64(const uint64_t* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline uint32_t __ballot_sync(uint32_t, bool pred) {
    const smb_emu::WarpColl& c = smb_emu::collective(pred ? 1 : 0);
    uint32_t m = 0;
    for (int l = 0; l < 32; ++l) if (((c.snap_mask >> l) & 1u) && c.snap[l]) m |= 1u << l;
    return m;
}
template <class T> inline T __shfl_sync(uint32_t, T v, int src) {
    uint64_t raw = 0;
    static_assert(sizeof(T) <= 8, "shfl payload");
    memcpy(&raw, &v, sizeof(T));
    const smb_emu::WarpColl& c = smb_emu::collective(raw);
    T out;
    memcpy(&out, &c.snap[src & 31], sizeof(T));
    return out;
}
template <class T> inline T __shfl_xor_sync(uint32_t, T v, int lane_mask) {
    uint64_t raw = 0;
    static_assert(sizeof(T) <= 8, "shfl payload");
    memcpy(&raw, &v, sizeof(T));
    const unsigned me = smb_emu::st().cur->linear & 31u;
    const smb_emu::WarpColl& c = smb_emu::collective(raw);
    T out;
    memcpy(&out, &c.snap[(me ^ (unsigned)lane_mask) & 31], sizeof(T));
    return out;
}
inline bool __any_sync(uint32_t, bool pred) {
    const smb_emu::WarpColl& c = smb_emu::collective(pred ? 1 : 0);
    for (int l = 0; l < 32; ++l) if (((c.snap_mask >> l) & 1u) && c.snap[l]) return true;
    return false;
}
inline bool __all_sync(uint32_t, bool pred) {
    const smb_emu::WarpColl& c = smb_emu::collective(pred ? 1 : 0);
    for (int l = 0; l < 32; ++l) if (((c.snap_mask >> l) & 1u) && !c.snap[l]) return false;
    return true;
}
inline uint32_t __reduce_add_sync(uint32_t, uint32_t v) {
    const smb_emu::WarpColl& c = smb_emu::collective(v);
    uint32_t s = 0;
    for (int l = 0; l < 32; ++l) if ((c.snap_mask >> l) & 1u) s += (uint32_t)c.snap[l];
    return s;
}
inline void __syncwarp(uint32_t = 0xffffffffu) { smb_emu::collective(0); }
inline int __ffs(uint32_t x) { return x ? __builtin_ctz(x) + 1 : 0; }
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
using std::max;
using std::min;
// the typed overload cuda_runtime.h only offers to nvcc
template <class T> inline cudaError_t cudaFuncSetAttribute(T* f, cudaFuncAttribute a, int v) {
    return ::cudaFuncSetAttribute(reinterpret_cast<const void*>(f), a, v);
}
