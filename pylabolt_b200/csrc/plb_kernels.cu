// plb_kernels.cu -- sm_100a kernels of the fluidLB step.
//
// One time step of the reference (pylabolt/solvers/fluidLB.py:206-253) makes
// five full passes over array-of-structures fields.  Here it is ONE pass:
// every fluid node reads its nine populations (structure of arrays, fully
// coalesced), computes rho / u / force, collides in registers and PUSHES the
// post-collision populations into the second lattice.
// 144 B per node per step is all that touches HBM, plus one code byte.
//
//   k_bulk_vec2 / k_bulk_scalar : nodes whose eight neighbours are plain
//                                 fluid nodes (NODE_BULK) -- >99.9 % of a
//                                 large lattice; no flags beyond the code byte
//   k_links                     : the O(perimeter) NODE_LINK nodes (domain
//                                 edges, obstacle surfaces) from a list, with
//                                 per-direction link codes: halfway bounce
//                                 back, moving wall, anti-bounce-back
//                                 pressure, periodic wrap, uncovered ghost
//   k_zero_gradient             : thin pass for zero_gradient elements
//   k_face_unpack               : slab-face / periodic-x delivery of the
//                                 three populations that cross a face
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plb_collide.cuh"
#ifndef PLB_EMU_RUNTIME
#include <cuda.h>            // CUtensorMap and its enums (no libcuda linkage)
#endif

// Every kernel launch goes through PLB_LAUNCH.  MODE documents (and, in the
// host-side SIMT emulation used by tests/emu, selects) how the kernel's
// threads cooperate: SIMPLE = independent threads, COOP = warp shuffles /
// votes / __syncthreads.  The kernel name is parenthesised so that template
// arguments with commas survive the preprocessor.
#define PLB_UNPAREN(...) __VA_ARGS__
#ifdef PLB_EMU_RUNTIME
#define PLB_LAUNCH(MODE, KERNEL, GRID, BLOCK, STREAM, ...)                    \
    PLB_EMU_LAUNCH(MODE, (PLB_UNPAREN KERNEL), GRID, BLOCK, STREAM, __VA_ARGS__)
#define PLB_LAUNCH_SMEM(MODE, KERNEL, GRID, BLOCK, SMEM, STREAM, ...)         \
    PLB_EMU_LAUNCH(MODE, (PLB_UNPAREN KERNEL), GRID, BLOCK, STREAM, __VA_ARGS__)
#else
#define PLB_LAUNCH(MODE, KERNEL, GRID, BLOCK, STREAM, ...)                    \
    PLB_UNPAREN KERNEL<<<(GRID), (BLOCK), 0, (STREAM)>>>(__VA_ARGS__)
// SMEM bytes of dynamic shared memory (the launcher has raised the kernel's
// cudaFuncAttributeMaxDynamicSharedMemorySize if that is more than 48 KB)
#define PLB_LAUNCH_SMEM(MODE, KERNEL, GRID, BLOCK, SMEM, STREAM, ...)         \
    PLB_UNPAREN KERNEL<<<(GRID), (BLOCK), (SMEM), (STREAM)>>>(__VA_ARGS__)
#endif

namespace plb {

// ---------------------------------------------------------------------------
// memory access helpers
// ---------------------------------------------------------------------------
// PLB_LD_MODE / PLB_ST_MODE: 0 default, 1 streaming (.cs), 2 (loads only)
// read-only path without L1 allocation.  The populations are touched exactly
// once per step, so nothing is gained by keeping them in L1 / L2.
#ifndef PLB_LD_MODE
#define PLB_LD_MODE 0
#endif
#ifndef PLB_ST_MODE
#define PLB_ST_MODE 0
#endif
#ifndef PLB_BLOCK
#define PLB_BLOCK 128
#endif
// Issue the population loads before the node-code check (see k_bulk_vec2).
// Measured on B200 (profiles/r01_occupancy_sweep.txt): +1 % (BGK) to +3.7 %
// (BGK + Guo) for the reference-ordered BGK kernels, -0.3 .. -1.7 % for the
// two-stress-moment MRT, hence per collision model.
__host__ __device__ constexpr bool bulk_speculate(int coll)
{
#ifdef PLB_SPECULATE
    return PLB_SPECULATE != 0;
#else
    return coll == 0;
#endif
}

// Resident CTAs per SM requested from ptxas for the 128-bit bulk kernel.  The
// kernel is latency bound on HBM: what matters is bytes in flight per SM, i.e.
// occupancy, so registers are capped as low as each instantiation allows
// without spilling (measured on B200, profiles/r01_occupancy_sweep.txt:
// 6 CTAs x 128 threads = 80 registers reaches the measured copy bandwidth;
// the Guo second-order MRT kernel needs 96 registers, the nine-rate MRT in moment space 94 at 5 CTAs).
__host__ __device__ constexpr int bulk_min_blocks(int coll, int forcing)
{
#ifdef PLB_MINBLOCKS
    return PLB_MINBLOCKS;
#else
    return coll == 1 ? 5 : 6;
#endif
}

__device__ __forceinline__ double2 ld2(const double *p)
{
#if PLB_LD_MODE == 1
    return __ldcs(reinterpret_cast<const double2 *>(p));
#elif PLB_LD_MODE == 2
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return *reinterpret_cast<const double2 *>(p);
#endif
}
__device__ __forceinline__ void st2(double *p, double a, double b)
{
#if PLB_ST_MODE == 1
    __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
#else
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
#endif
}
__device__ __forceinline__ void st1(double *p, double a)
{
#if PLB_ST_MODE == 1
    __stcs(p, a);
#else
    *p = a;
#endif
}
// The pair (lo, hi) to p[0], p[1]: both halves (mode 3) as one 128-bit store,
// one half (mode 1: lo, mode 2: hi) as a 64-bit store, nothing (mode 0) --
// predicated, so that lanes with different modes do not diverge.
__device__ __forceinline__ void st_pair(double *p, double lo, double hi, unsigned mode)
{
#if defined(PLB_EMU_RUNTIME) || PLB_ST_MODE == 1
    if (mode == 3) st2(p, lo, hi);
    else if (mode == 1) st1(p, lo);
    else if (mode == 2) st1(p + 1, hi);
#else
    asm volatile(
        "{\n\t.reg .pred both, only_lo, only_hi;\n\t"
        "setp.eq.u32 both, %3, 3;\n\t"
        "setp.eq.u32 only_lo, %3, 1;\n\t"
        "setp.eq.u32 only_hi, %3, 2;\n\t"
        "@both st.global.v2.f64 [%0], {%1, %2};\n\t"
        "@only_lo st.global.f64 [%0], %1;\n\t"
        "@only_hi st.global.f64 [%0+8], %2;\n\t}"
        ::"l"(p), "d"(lo), "d"(hi), "r"(mode)
        : "memory");
#endif
}

// ---------------------------------------------------------------------------
// one bulk node, scalar accesses
// ---------------------------------------------------------------------------
template <int COLL, int FORCING, bool STORE>
__device__ __forceinline__ void bulk_node(const StepArgs &a, int64_t idx)
{
    const int64_t plane = a.p.L.plane, pitch = a.p.L.pitch;
    double f[Q], g[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) f[k] = a.fin[k * plane + idx];
    const Moments m = collide<COLL, FORCING>(a.p, f, g);
    if constexpr (STORE) {
        a.rho[idx] = m.rho;
        a.ux[idx] = m.ux;
        a.uy[idx] = m.uy;
    }
#pragma unroll
    for (int k = 0; k < Q; ++k)
        a.fout[k * plane + idx + d_cx[k] * pitch + d_cy[k]] = g[k];
}

template <int COLL, int FORCING, bool STORE>
__global__ void __launch_bounds__(256)
k_bulk_scalar(StepArgs a, int64_t x_begin, int32_t chunks_per_row)
{
    const int64_t row = blockIdx.x / chunks_per_row;
    const int32_t chunk = blockIdx.x - row * chunks_per_row;
    const int64_t y = int64_t(chunk) * 256 + threadIdx.x;
    if (y >= a.p.L.ny) return;
    const int64_t idx = a.p.L.at(x_begin + row, y);
    if (a.code[idx] != NODE_BULK) return;
    bulk_node<COLL, FORCING, STORE>(a, idx);
}

// ---------------------------------------------------------------------------
// slab-edge columns with peer-to-peer faces
// ---------------------------------------------------------------------------
// Where population q pushed to (tx, ty) lands: a ghost row that stands for the
// neighbouring slab is redirected into that rank's receive buffer (peer memory
// mapped over NVLink), everything else is the local lattice.
__device__ __forceinline__ double *push_dst(const StepArgs &a, int q, int64_t tx,
                                            int64_t ty)
{
    const Layout &L = a.p.L;
    // populations 1, 5, 8 cross the right face, 3, 6, 7 the left one
    const int slot = (q == 1 || q == 3) ? 0 : ((q == 5 || q == 6) ? 1 : 2);
    if (tx < 0 && a.face_lo)
        return a.face_lo + slot * a.face_stride + L.y0 + ty;
    if (tx >= L.nx && a.face_hi)
        return a.face_hi + slot * a.face_stride + L.y0 + ty;
    return a.fout + q * a.fout_plane + lat_off(a.fout_map, L.at(tx, ty));
}

template <int COLL, int FORCING, bool STORE>
__global__ void __launch_bounds__(256)
k_bulk_edge(StepArgs a, int64_t x_begin, int32_t chunks_per_row)
{
    const int64_t row = blockIdx.x / chunks_per_row;
    const int32_t chunk = blockIdx.x - row * chunks_per_row;
    const int64_t y = int64_t(chunk) * 256 + threadIdx.x;
    if (y >= a.p.L.ny) return;
    const int64_t x = x_begin + row;
    const int64_t idx = a.p.L.at(x, y);
    if (a.code[idx] != NODE_BULK) return;
    double f[Q], g[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) f[k] = a.fin[k * a.p.L.plane + idx];
    const Moments m = collide<COLL, FORCING>(a.p, f, g);
    if constexpr (STORE) {
        a.rho[idx] = m.rho;
        a.ux[idx] = m.ux;
        a.uy[idx] = m.uy;
    }
#pragma unroll
    for (int k = 0; k < Q; ++k)
        *push_dst(a, k, x + d_cx[k], y + d_cy[k]) = g[k];
}

// ---------------------------------------------------------------------------
// two nodes per thread, 128-bit loads and stores
// ---------------------------------------------------------------------------
// A warp owns 64 consecutive y of one row.  Loads are aligned double2.  The
// three populations with c_y = 0 are stored as aligned double2 into rows
// x-1, x, x+1.  For c_y = +-1 the destination is shifted by one double, so the
// aligned pair [y, y+1] of the destination row is assembled from this lane
// and its neighbour with one shuffle; only the two ends of the 64-node run
// fall back to 8-byte stores.  Any warp that contains a non-bulk node takes
// the scalar path (domain edges, obstacle surfaces).
template <int COLL, int FORCING, bool STORE>
__global__ void __launch_bounds__(PLB_BLOCK, bulk_min_blocks(COLL, FORCING))
k_bulk_vec2(StepArgs a, int64_t x_begin, int32_t chunks_per_row)
{
    const Layout &L = a.p.L;
    const int64_t row = blockIdx.x / chunks_per_row;
    const int32_t chunk = blockIdx.x - row * chunks_per_row;
    const int64_t y = int64_t(chunk) * (2 * PLB_BLOCK) + 2 * threadIdx.x;
    const int64_t idx = L.at(x_begin + row, y);
    const int lane = threadIdx.x & 31;

    uint16_t codes = 0x0303;
    if (y + 1 < L.ny)
        codes = *reinterpret_cast<const uint16_t *>(a.code + idx);
    else if (y < L.ny)
        codes = uint16_t(a.code[idx]) | 0x0300;

    const int64_t plane = L.plane, pitch = L.pitch;
    double fa[Q], fb[Q], ga[Q], gb[Q];
    // Speculative variant: the population loads do not wait for the code
    // byte.  Any pair inside the padded row is a valid address, so they are
    // issued together with the code load and one DRAM round trip leaves the
    // critical path.  (Threads past the row end hold no bulk node; their warp
    // takes the scalar path.)
    if constexpr (bulk_speculate(COLL)) {
        if (L.y0 + y + 1 < pitch) {
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                const double2 v = ld2(a.fin + k * plane + idx);
                fa[k] = v.x;
                fb[k] = v.y;
            }
        }
    }
    if (!__all_sync(0xffffffffu, codes == 0)) {
        if ((codes & 0xff) == NODE_BULK)
            bulk_node<COLL, FORCING, STORE>(a, idx);
        if ((codes >> 8) == NODE_BULK)
            bulk_node<COLL, FORCING, STORE>(a, idx + 1);
        return;
    }
    if constexpr (!bulk_speculate(COLL)) {
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            const double2 v = ld2(a.fin + k * plane + idx);
            fa[k] = v.x;
            fb[k] = v.y;
        }
    }
    const Moments ma = collide<COLL, FORCING>(a.p, fa, ga);
    const Moments mb = collide<COLL, FORCING>(a.p, fb, gb);
    if constexpr (STORE) {
        st2(a.rho + idx, ma.rho, mb.rho);
        st2(a.ux + idx, ma.ux, mb.ux);
        st2(a.uy + idx, ma.uy, mb.uy);
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        double *dst = a.fout + k * plane + idx + d_cx[k] * pitch;
        if (d_cy[k] == 0) {
            st2(dst, ga[k], gb[k]);
        } else if (d_cy[k] == 1) {
            // values move to y+1, y+2: pair [y, y+1] = (left lane's b, own a)
            const double up = __shfl_up_sync(0xffffffffu, gb[k], 1);
            if (lane != 0) st2(dst, up, ga[k]);
            else st1(dst + 1, ga[k]);
            if (lane == 31) st1(dst + 2, gb[k]);
        } else {
            // values move to y-1, y: pair [y, y+1] = (own b, right lane's a)
            const double dn = __shfl_down_sync(0xffffffffu, ga[k], 1);
            if (lane != 31) st2(dst, gb[k], dn);
            else st1(dst, gb[k]);
            if (lane == 0) st1(dst - 1, ga[k]);
        }
    }
}

// ---------------------------------------------------------------------------
// several time steps in one pass (temporal blocking on chip)
// ---------------------------------------------------------------------------
// The single-step kernels above move 144 B per node and step and run at the
// HBM roofline; the only way past it is to touch HBM less.  k_bulk_fused
// advances DEEP nodes by DEPTH = 2, 3 or 4 steps per pass: lattice A (time t)
// is read once, lattice B (time t + DEPTH) is written once, 144 / DEPTH bytes
// per node and step.  The intermediate lattices never exist in memory:
//
//   * a warp owns a strip of 64 consecutive y (two per lane, 128-bit loads) and
//     marches along x; a row is loaded and collided once ("level 0"), and its
//     post-collision populations are kept in registers, already shifted in y
//     (one shuffle per population with c_y != 0) to the lane that PULLS them;
//   * the time-(t+1) state of the row before it is then complete in registers:
//     k = 1, 5, 8 came from two rows back (two iterations old), k = 0, 2, 4
//     from one row back (one iteration old), k = 3, 6, 7 are fresh.  It is
//     collided again; with DEPTH = 3 (4) the same hand-over happens once
//     (twice) more, one row further back each time; the last collision is
//     pushed into B exactly like k_bulk_vec2 pushes;
//   * every hand-over costs the strip its two outer nodes (no y-neighbour
//     inside the warp), so a warp delivers 62 (60, 58) of its 64 nodes and
//     strips overlap by two (four, six); chunks of rows overlap by two (four,
//     six) rows in x.  The overlapped loads hit L2.  No block barrier anywhere.
//
// A node is deep enough for DEPTH steps if every node within Chebyshev
// distance DEPTH - 1 is a bulk node (`deep` plane: that distance, capped).
// Every other node (domain edges, obstacle surfaces, slab-edge columns, and the
// rings around them) is advanced by DEPTH ordinary list passes on the edge
// stream (step_fused in plb_api.cu).  The arithmetic per node and step is the
// same collide<>() as everywhere else, so the strict build stays bit-identical
// to the reference.
__host__ __device__ constexpr int fused_span(int depth) { return 64 - 2 * (depth - 1); }
#ifndef PLB_FUSED_BLOCK
#define PLB_FUSED_BLOCK 128
#endif
// Shipped configuration since round 2 (profiles/r02_fused_sweep_*.txt, 4096 x
// 16384 sweep lattice, MRT + Guo / BGK):
//   * the carried populations live in shared memory (PLB_FUSED_CARRY_SMEM), which
//     takes three steps per pass from 200 to ~120 registers;
//   * the row ring is ONE slot per warp filled by the TMA unit (PLB_FUSED_BULK,
//     PLB_FUSED_STAGES = 1) with one tensor copy per row (PLB_FUSED_TENSOR):
//     refilled as soon as it has been read, it still fetches one row ahead,
//     needs no destination registers and no LSU instruction in 31 of 32 lanes;
//   * CTAs of 128 threads; how many per SM is set by shared memory (37 / 55 /
//     74 KB at two / three / four steps per pass) and, per collision model, by
//     the register cap of fused_min_blocks().
// History: round 1's two-step kernel (registers, cp.async ring) 82.2 / 81.5
// GLUPS; three steps per pass with the carry in registers (200 registers, 8
// warps) 88.6 / 80.5; carry in shared memory + nine bulk copies per row 109.0 /
// 96.3; tensor copy + instruction diet 114.5 / 113.7; four steps per pass
// 132.3 / 110.5 (the default for the two-stress-moment MRT kernel).
#ifndef PLB_FUSED_MINBLOCKS
#define PLB_FUSED_MINBLOCKS 4
#endif
// Row prefetch ring.  Warps wait on their row loads (round-1 ncu: long
// scoreboard 6 of 10 cycles per issue), so rows are fetched AHEAD of their use
// into a per-warp shared-memory ring -- asynchronous copies need no
// destination registers.  PLB_FUSED_STAGES = slots of the ring; 0 = plain loads.
// Every lane reads only its own 16 bytes per population, so the ring needs no
// block barrier.
#ifndef PLB_FUSED_STAGES
#define PLB_FUSED_STAGES 1
#endif
// PLB_FUSED_BULK=1: the ring is filled by the TMA unit -- a warp's row is nine
// contiguous runs of 512 bytes, so one elected lane issues nine cp.async.bulk
// copies (SASS UBLKCP) that complete on a per-warp, per-slot mbarrier.  A slot
// is refilled right after it has been read into registers (behind a
// cross-proxy fence, see fill()), so PLB_FUSED_STAGES slots keep that many rows
// in flight and a single slot (18 KB per CTA) already fetches one row ahead.
// PLB_FUSED_BULK=0: per-lane cp.async.cg copies (needs PLB_FUSED_STAGES >= 2).
#ifndef PLB_FUSED_BULK
#define PLB_FUSED_BULK 1
#endif
#if PLB_FUSED_BULK && PLB_FUSED_STAGES < 1
#error "PLB_FUSED_BULK needs a ring (PLB_FUSED_STAGES >= 1)"
#endif
// PLB_FUSED_TENSOR=1 (with PLB_FUSED_BULK): the lattice is described to the TMA
// unit as a rank-3 tensor (column, row, population) and a warp's row -- a box
// of 64 columns x 1 row x 9 populations -- is fetched by ONE
// cp.async.bulk.tensor.3d (SASS UTMALDG) instead of nine linear bulk copies:
// the nine-copy issue sequence was 185 of the ~1250 instructions a warp issues
// per row (round-2 SASS, profiles/r02_sass_k_bulk_fused_mrt_guo2_depth3.txt),
// executed by one lane.  Columns beyond the padded row are zero-filled by the
// unit, which is what the lanes outside the row read before, too.  The box
// lands densely, so a warp's ring slot is 9 x 512 contiguous bytes.
// (Default below, once PLB_FUSED_CARRY_SMEM is known.)
#if !PLB_FUSED_BULK && PLB_FUSED_STAGES == 1
#error "a cp.async ring needs two slots (PLB_FUSED_STAGES >= 2), or none (0)"
#endif
// PLB_FUSED_CARRY_SMEM=1: the post-collision populations that wait one or two
// iterations for their row (FusedCarry: 18 doubles per lane and level) live in
// shared memory instead of registers -- every lane reads and writes only its
// own 16-byte slots, once per iteration, so no barrier is needed and the
// accesses are conflict free.  Frees ~36 registers per level.
#ifndef PLB_FUSED_CARRY_SMEM
#define PLB_FUSED_CARRY_SMEM 1
#endif
#ifndef PLB_FUSED_TENSOR
#define PLB_FUSED_TENSOR (PLB_FUSED_BULK && PLB_FUSED_CARRY_SMEM)
#endif
#if PLB_FUSED_TENSOR && !PLB_FUSED_BULK
#error "PLB_FUSED_TENSOR is a way of filling the PLB_FUSED_BULK ring"
#endif
#if PLB_FUSED_TENSOR && !PLB_FUSED_CARRY_SMEM
#error "PLB_FUSED_TENSOR addresses the ring in dynamic shared memory (PLB_FUSED_CARRY_SMEM=1)"
#endif
// Shared memory of one CTA of k_bulk_fused<.., DEPTH>: the ring, the carried
// populations of DEPTH - 1 levels, the ring's mbarriers.  The shipped build (ring
// of two rows: 36 KB) declares it statically; a variant that needs more than
// the 48 KB a kernel may declare statically takes all of it from dynamic
// shared memory (PLB_FUSED_DYN_SMEM).
constexpr int FUSED_RING_BYTES = PLB_FUSED_STAGES * Q * PLB_FUSED_BLOCK * 16;
constexpr int FUSED_CARRY_LEVEL_BYTES = PLB_FUSED_CARRY_SMEM ? Q * PLB_FUSED_BLOCK * 16 : 0;
constexpr int FUSED_BAR_BYTES = PLB_FUSED_BULK ? PLB_FUSED_STAGES * (PLB_FUSED_BLOCK / 32) * 8 : 0;
__host__ __device__ constexpr int fused_smem_bytes(int depth)
{
    return FUSED_RING_BYTES + (depth - 1) * FUSED_CARRY_LEVEL_BYTES + FUSED_BAR_BYTES;
}
#ifndef PLB_FUSED_DYN_SMEM
#define PLB_FUSED_DYN_SMEM \
    (PLB_FUSED_CARRY_SMEM || PLB_FUSED_STAGES * 9 * PLB_FUSED_BLOCK * 16 > 48 * 1024 - 256)
#endif

// Resident CTAs per SM asked of ptxas, per collision model (PLB_FUSED_MINBLOCKS
// for all, PLB_FUSED_MINBLOCKS_BGK for the reference-ordered BGK kernels, which
// need more registers than the two-stress-moment MRT).
#ifndef PLB_FUSED_BULK_FENCE
#define PLB_FUSED_BULK_FENCE 1
#endif
#ifndef PLB_FUSED_BULK_LATE
#define PLB_FUSED_BULK_LATE 0
#endif
#ifndef PLB_FUSED_MINBLOCKS_D3
#define PLB_FUSED_MINBLOCKS_D3 (512 / PLB_FUSED_BLOCK)
#endif
#ifndef PLB_FUSED_MINBLOCKS_D3_MRT
#define PLB_FUSED_MINBLOCKS_D3_MRT (384 / PLB_FUSED_BLOCK)
#endif
__host__ __device__ constexpr int fused_min_blocks(int coll, int depth)
{
    // three steps per pass: shared memory (55 KB per CTA) admits four CTAs of
    // 128 threads per SM, i.e. 128 registers per thread
    // Measured (profiles/r02_fused_sweep_v5_tensor.txt): the two-stress-moment
    // MRT kernel is faster with three CTAs of 142 registers (117.3 GLUPS) than
    // with four of 119 (114.5); the reference-ordered BGK kernels the other
    // way round (109.1 with three of 134, 113.7 with four of 128); the
    // nine-rate MRT spills beyond 128.
    if (depth >= 3) return coll == 2 ? PLB_FUSED_MINBLOCKS_D3_MRT : PLB_FUSED_MINBLOCKS_D3;
#ifdef PLB_FUSED_MINBLOCKS_BGK
    return coll == 0 ? PLB_FUSED_MINBLOCKS_BGK : PLB_FUSED_MINBLOCKS;
#else
    return PLB_FUSED_MINBLOCKS;
#endif
}

#ifdef PLB_EMU_RUNTIME
#define PLB_GRID_CONSTANT
#else
#define PLB_GRID_CONSTANT __grid_constant__
#endif

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
#ifdef PLB_EMU_RUNTIME
    *static_cast<double2 *>(smem) = *static_cast<const double2 *>(gmem);
#else
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gmem)
                 : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit()
{
#ifndef PLB_EMU_RUNTIME
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
#ifndef PLB_EMU_RUNTIME
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// ---- TMA bulk copies completing on an mbarrier (PLB_FUSED_BULK) ------------
// The emulation copies at issue time; there the warp-level barrier that
// stands in for the wait orders the elected lane's copy before the reads.
__device__ __forceinline__ void mbar_init(unsigned long long *bar)
{
#ifdef PLB_EMU_RUNTIME
    *bar = 0;
#else
    const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
#endif
}
__device__ __forceinline__ void mbar_init_fence()
{
#ifndef PLB_EMU_RUNTIME
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
#ifndef PLB_EMU_RUNTIME
    const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b),
                 "r"(bytes)
                 : "memory");
#endif
}
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, unsigned bytes,
                                         unsigned long long *bar)
{
#ifdef PLB_EMU_RUNTIME
    const char *src = static_cast<const char *>(gmem);
    char *dst = static_cast<char *>(smem);
    for (unsigned i = 0; i < bytes; ++i) dst[i] = src[i];
#else
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(d), "l"(gmem), "r"(bytes), "r"(b)
        : "memory");
#endif
}
// One box of (64 columns, 1 row, Q populations) of the lattice behind `map`,
// first column c0 / row c1, into 9 x 512 dense bytes of shared memory.
__device__ __forceinline__ void tensor_g2s(void *smem, const TensorMap *map, int c0,
                                           int c1, unsigned long long *bar)
{
#ifdef PLB_EMU_RUNTIME
    const double *base = reinterpret_cast<const double *>(map->opaque[0]);
    const int64_t pitch = int64_t(map->opaque[1]), rows = int64_t(map->opaque[2]);
    const int64_t plane = int64_t(map->opaque[3]);
    double *dst = static_cast<double *>(smem);
    for (int k = 0; k < Q; ++k)
        for (int c = 0; c < 64; ++c) {
            const int64_t col = int64_t(c0) + c;
            const bool inside = col >= 0 && col < pitch && c1 >= 0 && c1 < rows;
            dst[k * 64 + c] = inside ? base[k * plane + int64_t(c1) * pitch + col] : 0.0;
        }
    (void)bar;
#else
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(d), "l"(map), "r"(c0), "r"(c1), "r"(0), "r"(b)
        : "memory");
#endif
}
// Orders this thread's earlier generic-proxy accesses to shared memory (the
// ld.shared of a ring slot) before later async-proxy accesses (the bulk copy
// that refills the slot).
__device__ __forceinline__ void fence_proxy_async_smem()
{
#ifndef PLB_EMU_RUNTIME
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
// Waits for the phase of `bar` with the given parity.  A copy that never
// completes traps after ~1 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
#ifdef PLB_EMU_RUNTIME
    __syncwarp();
#else
    const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
    unsigned done = 0;
    for (unsigned spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(b), "r"(parity)
            : "memory");
        if (done) break;
        if (spins > (1u << 20)) __trap();
    }
#endif
}

// PLB_FUSED_PIN=1: the collision constants a warp uses 3 x 2 times per row live in
// registers for the whole kernel.  ptxas otherwise re-reads them from the
// constant bank inside the loop (round-2 SASS: 130 LDC / LDCU of ~1070
// instructions per row); the kernel's occupancy is set by its shared memory
// (four CTAs of 55 KB), which leaves 128 registers per thread, 32 more than
// the loop needs.  ptxas rematerialises anything it can trace back to the
// constant bank (it even folds a warp shuffle of a uniform value), so the bit
// pattern is XORed with a zero it cannot prove to be zero: bit 63 of the
// cycle counter, read once per thread.
#ifndef PLB_FUSED_PIN
#define PLB_FUSED_PIN 1
#endif
// Neighbouring chunks march in opposite x directions (see k_bulk_fused);
// PLB_FUSED_ALTERNATE in the environment overrides.
#ifndef PLB_FUSED_ALTERNATE_DEFAULT
#define PLB_FUSED_ALTERNATE_DEFAULT 1
#endif
__device__ __forceinline__ unsigned long long opaque_zero()
{
#ifdef PLB_EMU_RUNTIME
    return 0;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t >> 63;
#endif
}
__device__ __forceinline__ void pin(double &v, unsigned long long zero)
{
#ifdef PLB_EMU_RUNTIME
    (void)zero;
#else
    v = __longlong_as_double(__double_as_longlong(v) ^ (long long)zero);
#endif
}
template <int COLL, int FORCING>
__device__ __forceinline__ void pin_collision_constants(KParams &kp)
{
#if PLB_FUSED_PIN
    const unsigned long long z = opaque_zero();
    if constexpr (COLL == 2) {
        pin(kp.mrt.k2, z); pin(kp.mrt.k4, z); pin(kp.mrt.k7, z); pin(kp.mrt.k8, z);
        pin(kp.mrt.qa, z); pin(kp.mrt.qb, z);
        pin(kp.w[0], z); pin(kp.w[1], z); pin(kp.w[5], z);
        pin(kp.inv_cs_2, z); pin(kp.eps, z);
        if constexpr (FORCING != 0) { pin(kp.mrt.hgx, z); pin(kp.mrt.hgy, z); }
#if PLB_FUSED_PIN >= 2
        if constexpr (FORCING == 2) {
            pin(kp.gx, z); pin(kp.gy, z);
#pragma unroll
            for (int j = 0; j < 4; ++j) { pin(kp.mrt.k4cg[j], z); pin(kp.mrt.k2cg[j], z); }
        }
#endif
    } else if constexpr (COLL == 0) {
        pin(kp.omega, z); pin(kp.inv_cs_2, z); pin(kp.inv_cs_4, z); pin(kp.eps, z);
        pin(kp.w[0], z); pin(kp.w[1], z); pin(kp.w[5], z);
    }
#else
    (void)kp;
#endif
}

// Stage 1 of one row: collide the pair (y, y + 1) and hand every
// post-collision population to the lane that pulls it in the next step:
// sa[k] is what node y receives in slot k, sb[k] what node y + 1 receives.
// (For c_y = +1 lane 0's sa and for c_y = -1 lane 31's sb come from outside
// the warp and are meaningless: those two nodes are not delivered.)
template <int COLL, int FORCING>
__device__ __forceinline__ void fused_stage1(const KParams &kp, const double fa[Q],
                                             const double fb[Q], double sa[Q],
                                             double sb[Q])
{
    double ga[Q], gb[Q];
    collide<COLL, FORCING>(kp, fa, ga);
    collide<COLL, FORCING>(kp, fb, gb);
#pragma unroll
    for (int k = 0; k < Q; ++k) {
        if (d_cy[k] == 0) {
            sa[k] = ga[k];
            sb[k] = gb[k];
        } else if (d_cy[k] == 1) {          // pulled from y - 1
            sa[k] = __shfl_up_sync(0xffffffffu, gb[k], 1);
            sb[k] = ga[k];
        } else {                            // pulled from y + 1
            sa[k] = gb[k];
            sb[k] = __shfl_down_sync(0xffffffffu, ga[k], 1);
        }
    }
}

// Post-collision populations of the two previous rows of one level, already
// shifted to the lanes that pull them.
struct FusedCarry {
    double pa[3], pb[3];   // two rows back: k = 1, 5, 8
    double na[3], nb[3];   // one row back : k = 1, 5, 8
    double ca[3], cb[3];   // one row back : k = 0, 2, 4
};

// Hand-over between two levels with the carry in shared memory: stores what
// this row has just produced for its successors and completes the row one
// back, one step later.  DIR = +1: the warp marches towards larger x, so the
// populations with c_x = +1 (1, 5, 8) were produced two iterations ago by the
// row behind and those with c_x = -1 (3, 6, 7) are fresh from the row ahead;
// DIR = -1 (march towards smaller x): the other way round.  c: this level's
// nine slots (0-2 / 3-5: the behind-sourced populations of the even / odd
// rows, 6-8: k = 0, 2, 4 of the previous row).
template <int DIR>
__device__ __forceinline__ void fused_hand_over(double2 (*c)[PLB_FUSED_BLOCK], int two_back,
                                                const double sa[Q], const double sb[Q],
                                                double fa[Q], double fb[Q])
{
    constexpr int B0 = DIR > 0 ? 1 : 3, B1 = DIR > 0 ? 5 : 6, B2 = DIR > 0 ? 8 : 7;
    constexpr int A0 = DIR > 0 ? 3 : 1, A1 = DIR > 0 ? 6 : 5, A2 = DIR > 0 ? 7 : 8;
    const int t = threadIdx.x;
    const double2 p0 = c[two_back + 0][t];
    const double2 p1 = c[two_back + 1][t];
    const double2 p2 = c[two_back + 2][t];
    const double2 c0 = c[6][t];
    const double2 c2 = c[7][t];
    const double2 c4 = c[8][t];
    c[two_back + 0][t] = make_double2(sa[B0], sb[B0]);
    c[two_back + 1][t] = make_double2(sa[B1], sb[B1]);
    c[two_back + 2][t] = make_double2(sa[B2], sb[B2]);
    c[6][t] = make_double2(sa[0], sb[0]);
    c[7][t] = make_double2(sa[2], sb[2]);
    c[8][t] = make_double2(sa[4], sb[4]);
    fa[0] = c0.x; fa[2] = c2.x; fa[4] = c4.x;
    fb[0] = c0.y; fb[2] = c2.y; fb[4] = c4.y;
    fa[B0] = p0.x; fa[B1] = p1.x; fa[B2] = p2.x;
    fb[B0] = p0.y; fb[B1] = p1.y; fb[B2] = p2.y;
    fa[A0] = sa[A0]; fa[A1] = sa[A1]; fa[A2] = sa[A2];
    fb[A0] = sb[A0]; fb[A1] = sb[A1]; fb[A2] = sb[A2];
}

template <int COLL, int FORCING, int DEPTH>
__global__ void __launch_bounds__(PLB_FUSED_BLOCK, fused_min_blocks(COLL, DEPTH))
k_bulk_fused(StepArgs a, const uint8_t *__restrict__ deep, int64_t x_begin,
             int64_t x_end, int32_t strips, int32_t rows_per_chunk,
             unsigned *work_counter, const PLB_GRID_CONSTANT TensorMap tmap,
             int32_t alternate)
{
    constexpr int LEVELS = DEPTH - 1;              // hand-overs in registers
    const Layout &L = a.p.L;
    const int lane = threadIdx.x & 31;
    KParams kp = a.p;
    pin_collision_constants<COLL, FORCING>(kp);
#if PLB_FUSED_DYN_SMEM
#ifdef PLB_EMU_RUNTIME
    static __align__(128) unsigned char fused_smem[fused_smem_bytes(DEPTH)];
#else
    extern __shared__ __align__(128) unsigned char fused_smem[];
#endif
    typedef double2 RingSlot[Q][PLB_FUSED_BLOCK];
    RingSlot *ring = reinterpret_cast<RingSlot *>(fused_smem);
    // per level: slots 0-2 = k 1, 5, 8 of the even rows, 3-5 = of the odd
    // rows, 6-8 = k 0, 2, 4 of the previous row; (.x, .y) = nodes (y, y + 1).
    // Whatever an earlier work item left behind is never used: a row is
    // complete only after this item has written its two predecessors.
    RingSlot *carry_s = reinterpret_cast<RingSlot *>(fused_smem + FUSED_RING_BYTES);
    typedef unsigned long long RingBars[PLB_FUSED_BLOCK / 32];
    RingBars *ring_bar = reinterpret_cast<RingBars *>(
        fused_smem + FUSED_RING_BYTES + LEVELS * FUSED_CARRY_LEVEL_BYTES);
    (void)ring;
    (void)carry_s;
    (void)ring_bar;
#else
#if PLB_FUSED_STAGES >= 2 || PLB_FUSED_BULK
    __shared__ __align__(128) double2 ring[PLB_FUSED_STAGES][Q][PLB_FUSED_BLOCK];
#endif
#if PLB_FUSED_BULK
    __shared__ unsigned long long ring_bar[PLB_FUSED_STAGES][PLB_FUSED_BLOCK / 32];
#endif
#endif
#if PLB_FUSED_BULK
    // one mbarrier per warp and slot; `filled` counts the rows this warp has
    // sent through the ring since the kernel began (slot = filled % STAGES,
    // phase parity = filled / STAGES & 1), across work items
    const int wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < PLB_FUSED_STAGES; ++st) mbar_init(&ring_bar[st][wid]);
        mbar_init_fence();
    }
    __syncwarp();
    unsigned filled = 0;
    // this lane's 16 bytes of population k in ring slot `slot`
#if PLB_FUSED_TENSOR
    // a warp's slot is one dense box: [slot][warp][k][lane]
    auto ring_row = [&](int slot, int k) -> double2 * {
        return reinterpret_cast<double2 *>(fused_smem) +
               ((slot * (PLB_FUSED_BLOCK / 32) + wid) * Q + k) * 32 + lane;
    };
#else
    auto ring_row = [&](int slot, int k) -> double2 * {
        return &ring[slot][k][threadIdx.x];
    };
#endif
#endif
    // One work item = one chunk of rows of one strip.  Statically a warp takes
    // the item of its own number; with a work counter (PLB_FUSED_DYNAMIC=1,
    // persistent grid) warps draw items until none is left, so that no SM
    // idles while the last wave of a static assignment drains.
    int64_t warp = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (;;) {
        if (work_counter) {
            unsigned item = 0;
            if (lane == 0) item = atomicAdd(work_counter, 1u);
            warp = __shfl_sync(0xffffffffu, item, 0);
        }
        // item -> (chunk, strip): strips fastest, so that the warps of a CTA work
        // on y-adjacent strips of one chunk; or (alternate & 2) the warps of a
        // CTA take x-adjacent chunks of ONE strip, which -- marching in
        // alternating directions, started together, equally long -- meet at
        // every boundary between them at the same time
        int64_t chunk = warp / strips;
        int32_t strip = int32_t(warp - chunk * strips);
        if (alternate & 2) {
            constexpr int wpb = PLB_FUSED_BLOCK / 32;
            const int64_t group = warp / wpb;
            chunk = (group / strips) * wpb + (warp - group * wpb);
            strip = int32_t(group % strips);
        }
        const int64_t xs = x_begin + chunk * rows_per_chunk;
        if (xs >= x_end) {                             // whole warp
            // (with chunk groups an item past the last chunk may be followed by
            // valid items of the next strip: a warp that draws its items from
            // the queue goes on until the queue itself is exhausted)
            const int64_t n_chunks = (x_end - x_begin + rows_per_chunk - 1) / rows_per_chunk;
            constexpr int64_t wpb = PLB_FUSED_BLOCK / 32;
            if (!work_counter || !(alternate & 2) ||
                warp >= (n_chunks + wpb - 1) / wpb * wpb * strips)
                return;
            continue;
        }
        const int64_t xe = (xs + rows_per_chunk < x_end) ? xs + rows_per_chunk : x_end;
        const int64_t y = int64_t(strip) * fused_span(DEPTH) - 2 + 2 * lane;
        // the pair lies inside the padded row (y >= -2 is column >= 14)
        const bool in_row = L.y0 + y + 1 < L.pitch;
        const int64_t plane = L.plane, pitch = L.pitch;
        // After LEVELS hand-overs the strip has lost LEVELS nodes at either end:
        // this lane still delivers node y / node y + 1 if ...
        const bool lane_a = 2 * lane >= LEVELS && 2 * lane <= 63 - LEVELS;
        const bool lane_b = 2 * lane + 1 >= LEVELS && 2 * lane + 1 <= 63 - LEVELS;
        // which halves of the pair [y, y + 1] this lane stores, per c_y of the
        // population (bit 0: the low half, bit 1: the high half): c_y = 0 own
        // nodes, c_y = +1 (left neighbour's b, own a), c_y = -1 (own b, right
        // neighbour's a)
        const unsigned mode_0 = unsigned(lane_a) | unsigned(lane_b) << 1;
        const unsigned mode_p =
            unsigned(lane > 0 && 2 * lane - 1 <= 63 - LEVELS && 2 * lane - 1 >= LEVELS) |
            unsigned(lane_a) << 1;
        const unsigned mode_m =
            unsigned(lane_b) |
            unsigned(lane < 31 && 2 * lane + 2 >= LEVELS && 2 * lane + 2 <= 63 - LEVELS) << 1;

        // Rows xs - LEVELS .. xe - 1 + LEVELS go through level 0 in order (row
        // number i = 0 ..); row number i leaves level l (0-based) as the complete
        // state of row i - 1 one step later, meaningful from i = 2 (l + 1) on; the
        // row pushed into B in iteration i is x = xs + i - 2 LEVELS.
        // With `alternate`, odd chunks march the other way (row number i is
        // x = xe - 1 + LEVELS - i): a chunk and its neighbour then reach the rows
        // they share at the same time -- both at their start or both at their
        // end -- and the second read of those 2 LEVELS rows is an L2 hit instead
        // of a second trip to HBM (they were read ~100 us apart before, which
        // no line survives at 5 TB/s).
        const int n_rows = int(xe - xs) + 2 * LEVELS;
        const bool down = PLB_FUSED_CARRY_SMEM && (alternate & 1) && (chunk & 1);
        const int64_t x_first = down ? xe - 1 + LEVELS : xs - LEVELS;   // row number 0
        const int64_t row_step = down ? -pitch : pitch;
        const double *row0 = a.fin + L.at(x_first, y);          // pair of row i = 0
#if PLB_FUSED_BULK
        constexpr int AHEAD = PLB_FUSED_STAGES;
#if PLB_FUSED_TENSOR
        // the warp's box: 64 columns from the strip's first, row number j of the
        // item; a strip always begins inside the padded row
        const int box_col = int(L.y0 + y) - 2 * lane;
        const int box_row0 = int(x_first) + 1;
        const int box_step = down ? -1 : 1;
        constexpr unsigned run_bytes = 512;
        auto fill = [&](int j) {
            // cross-proxy fence + warp barrier: see the linear variant below
#if PLB_FUSED_BULK_FENCE
            fence_proxy_async_smem();
#endif
            __syncwarp();
            if (lane == 0 && j < n_rows) {
                const unsigned g = filled + unsigned(j);
                const int slot = int(g % PLB_FUSED_STAGES);
                unsigned long long *bar = &ring_bar[slot][wid];
                mbar_expect_tx(bar, Q * run_bytes);
                tensor_g2s(ring_row(slot, 0) - lane, &tmap, box_col, box_row0 + box_step * j,
                           bar);
            }
        };
#else
        // bytes of the warp's run that lie inside the padded row (the lanes
        // with in_row are a prefix of the warp), and the elected lane's view
        // of the run
        const unsigned run_bytes = 16u * unsigned(__popc(__ballot_sync(0xffffffffu, in_row)));
        const double *run0 = row0 - 2 * lane;
        // row number j of this item into its slot, which every lane has just
        // read into registers (__syncwarp: those reads are done)
        auto fill = [&](int j) {
            // The slot is rewritten by the async proxy (TMA) right after the
            // generic-proxy reads above: the cross-proxy fence orders every
            // lane's reads before the copy that lane 0 issues after the warp
            // barrier.  (Without it the round-2 hardware sweep produced
            // run-to-run different fields; the host emulation cannot see this.)
#if PLB_FUSED_BULK_FENCE
            fence_proxy_async_smem();
#endif
            __syncwarp();
            if (lane == 0 && run_bytes != 0 && j < n_rows) {
                const unsigned g = filled + unsigned(j);
                unsigned long long *bar = &ring_bar[g % PLB_FUSED_STAGES][wid];
                mbar_expect_tx(bar, Q * run_bytes);
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    bulk_g2s(ring_row(int(g % PLB_FUSED_STAGES), k) - lane,
                             run0 + k * plane + int64_t(j) * row_step, run_bytes, bar);
            }
        };
#endif
#pragma unroll
        for (int i = 0; i < AHEAD; ++i) fill(i);
#elif PLB_FUSED_STAGES >= 2
        constexpr int AHEAD = PLB_FUSED_STAGES - 1;
#pragma unroll
        for (int i = 0; i < AHEAD; ++i) {
            if (in_row && i < n_rows) {
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    cp_async16(&ring[i][k][threadIdx.x], row0 + k * plane + i * row_step);
            }
            cp_async_commit();
        }
#endif

#if !PLB_FUSED_CARRY_SMEM
        FusedCarry carry[LEVELS];
#pragma unroll
        for (int l = 0; l < LEVELS; ++l)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                carry[l].pa[j] = carry[l].pb[j] = carry[l].na[j] = carry[l].nb[j] =
                    carry[l].ca[j] = carry[l].cb[j] = 0.0;
#endif

        // deep flags of the row that is pushed in the NEXT iteration: fetched one
        // iteration ahead, so that the vote below never waits for them
        uint16_t dd_next = 0;
        // node index of this lane's pair in the row pushed in iteration i (LEVELS
        // rows behind the row loaded in that iteration), advanced by one row
        // per iteration
        int64_t idx = L.at(down ? x_first + LEVELS : x_first - LEVELS, y);
        for (int i = 0; i < n_rows; ++i, idx += row_step) {
            const uint16_t dd = dd_next;
            if (in_row && i + 1 >= 2 * LEVELS && i + 1 < n_rows)
                dd_next = *reinterpret_cast<const uint16_t *>(deep + idx + row_step);
            double fa[Q], fb[Q];
#if PLB_FUSED_BULK
            {
                const unsigned g = filled + unsigned(i);
                const int slot = int(g % PLB_FUSED_STAGES);
                if (run_bytes != 0)         // row i has landed
                    mbar_wait(&ring_bar[slot][wid], (g / PLB_FUSED_STAGES) & 1u);
#pragma unroll
                for (int k = 0; k < Q; ++k) {
#if PLB_FUSED_TENSOR
                    // columns outside the padded row were zero-filled by the unit
                    const double2 v = *ring_row(slot, k);
#else
                    double2 v = make_double2(0.0, 0.0);
                    if (in_row) v = *ring_row(slot, k);
#endif
                    fa[k] = v.x;
                    fb[k] = v.y;
                }
#if !PLB_FUSED_BULK_LATE
                fill(i + AHEAD);            // the same slot, AHEAD rows on
#endif
            }
#elif PLB_FUSED_STAGES >= 2
            {
                // refill the slot that was read in the previous iteration
                const int j = i + AHEAD;
                if (in_row && j < n_rows) {
#pragma unroll
                    for (int k = 0; k < Q; ++k)
                        cp_async16(&ring[j % PLB_FUSED_STAGES][k][threadIdx.x],
                                   row0 + k * plane + j * row_step);
                }
                cp_async_commit();
                cp_async_wait<AHEAD>();                 // row i has landed
                const int slot = i % PLB_FUSED_STAGES;
#pragma unroll
                for (int k = 0; k < Q; ++k) {
                    double2 v = make_double2(0.0, 0.0);
                    if (in_row) v = ring[slot][k][threadIdx.x];
                    fa[k] = v.x;
                    fb[k] = v.y;
                }
            }
#else
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                double2 v = make_double2(0.0, 0.0);
                if (in_row) v = ld2(row0 + k * plane + i * row_step);
                fa[k] = v.x;
                fb[k] = v.y;
            }
#endif
            bool complete = true;
#pragma unroll
            for (int l = 0; l < LEVELS; ++l) {
                if (!complete) break;
                double sa[Q], sb[Q];
                fused_stage1<COLL, FORCING>(kp, fa, fb, sa, sb);
#if PLB_FUSED_BULK && PLB_FUSED_BULK_LATE
                // refill only after the row has been collided: its values have
                // then provably left shared memory (data dependence)
                if (l == 0) fill(i + AHEAD);
#endif
#if PLB_FUSED_CARRY_SMEM
                {
                    const int two_back = (i & 1) * 3;       // written two iterations ago
                    if (down) fused_hand_over<-1>(carry_s[l], two_back, sa, sb, fa, fb);
                    else fused_hand_over<1>(carry_s[l], two_back, sa, sb, fa, fb);
                }
#else
                FusedCarry &c = carry[l];
                // one step later: the populations of the row one back
                fa[0] = c.ca[0]; fa[1] = c.pa[0]; fa[2] = c.ca[1]; fa[3] = sa[3]; fa[4] = c.ca[2];
                fa[5] = c.pa[1]; fa[6] = sa[6]; fa[7] = sa[7]; fa[8] = c.pa[2];
                fb[0] = c.cb[0]; fb[1] = c.pb[0]; fb[2] = c.cb[1]; fb[3] = sb[3]; fb[4] = c.cb[2];
                fb[5] = c.pb[1]; fb[6] = sb[6]; fb[7] = sb[7]; fb[8] = c.pb[2];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    c.pa[j] = c.na[j];
                    c.pb[j] = c.nb[j];
                }
                c.na[0] = sa[1]; c.na[1] = sa[5]; c.na[2] = sa[8];
                c.nb[0] = sb[1]; c.nb[1] = sb[5]; c.nb[2] = sb[8];
                c.ca[0] = sa[0]; c.ca[1] = sa[2]; c.ca[2] = sa[4];
                c.cb[0] = sb[0]; c.cb[1] = sb[2]; c.cb[2] = sb[4];
#endif
                complete = i >= 2 * (l + 1);
            }
            if (!complete) continue;

            const bool da = (dd & 0xff) >= LEVELS && lane_a;
            const bool db = (dd >> 8) >= LEVELS && lane_b;
            if (!__any_sync(0xffffffffu, da || db)) continue;

            double ha[Q], hb[Q];
            collide<COLL, FORCING>(kp, fa, ha);
            collide<COLL, FORCING>(kp, fb, hb);

            if (__all_sync(0xffffffffu, (da || !lane_a) && (db || !lane_b))) {
                // every node the warp can deliver is deep: 128-bit stores, pairs
                // re-aligned by shuffle (a pair whose other half belongs to a lost
                // node shrinks to a 64-bit store) -- three predicated stores per
                // population, no divergent branch (st_pair)
                char *row = reinterpret_cast<char *>(a.fout + idx);
#pragma unroll
                for (int k = 0; k < Q; ++k) {
                    double *dst = reinterpret_cast<double *>(row + a.push_off[k]);
                    if (d_cy[k] == 0) {
                        st_pair(dst, ha[k], hb[k], mode_0);
                    } else if (d_cy[k] == 1) {
                        // values move to y + 1, y + 2: pair [y, y+1] = (left b, own a)
                        st_pair(dst, __shfl_up_sync(0xffffffffu, hb[k], 1), ha[k], mode_p);
                    } else {
                        // values move to y - 1, y: pair [y, y+1] = (own b, right a)
                        st_pair(dst, hb[k], __shfl_down_sync(0xffffffffu, ha[k], 1), mode_m);
                    }
                }
            } else {
                // strip touches a node that is not deep (domain edge, obstacle, ring)
#pragma unroll
                for (int k = 0; k < Q; ++k) {
                    double *dst = a.fout + k * plane + idx + d_cx[k] * pitch + d_cy[k];
                    if (da) st1(dst, ha[k]);
                    if (db) st1(dst + 1, hb[k]);
                }
            }
        }
        if (!work_counter) return;
#if PLB_FUSED_BULK
        if (run_bytes != 0) filled += unsigned(n_rows);
#endif
    }
}

// ---------------------------------------------------------------------------
// link nodes
// ---------------------------------------------------------------------------
// velocity[ind] as the reference's fixed_pressure kernel would read it
// (cpu/fluid_boundary_kernels.py:131-140): phase-4 velocity of this step for a
// fluid node (recomputed from the intact old lattice), the stored rigid-body
// velocity for a solid node, zero on the ghost ring.
__device__ __forceinline__ void node_velocity(const StepArgs &a, int64_t idx,
                                              double &ux, double &uy)
{
    const uint8_t c = a.code[idx];
    if (c == NODE_BULK || c == NODE_LINK) {
        double f[Q];
        const int64_t off = lat_off(a.fin_map, idx);
#pragma unroll
        for (int k = 0; k < Q; ++k) f[k] = a.fin[k * a.fin_plane + off];
        const Moments m = moments(a.p, f);
        ux = m.ux;
        uy = m.uy;
    } else {
        ux = a.ux[idx];
        uy = a.uy[idx];
    }
}

template <int COLL, int FORCING, bool STORE>
__global__ void __launch_bounds__(128)
k_links(StepArgs a, const LinkNode *__restrict__ nodes, int64_t n_nodes,
        const ElementDev *__restrict__ elements)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const LinkNode nd = nodes[i];
    const Layout &L = a.p.L;
    const int64_t pitch = L.pitch;
    const int64_t idx = L.at(nd.x, nd.y);
    // the node's own place in the input / output lattice (full or compact)
    const int64_t in_off = lat_off(a.fin_map, idx);
    const int64_t out_off = lat_off(a.fout_map, idx);

    double f[Q], g[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) f[k] = a.fin[k * a.fin_plane + in_off];
    const Moments m = collide<COLL, FORCING>(a.p, f, g);
    if constexpr (STORE) {
        a.rho[idx] = m.rho;
        a.ux[idx] = m.ux;
        a.uy[idx] = m.uy;
    }
    a.fout[out_off] = g[0];   // pop_new[ind, 0] = pop[ind, 0], streaming_kernels.py:34

#pragma unroll
    for (int q = 1; q < Q; ++q) {
        const int code = int((nd.links >> (8 * (q - 1))) & 0xff);
        const int cx = d_cx[q], cy = d_cy[q], qi = d_inv[q];
        bool own = true;      // this node writes its own slot f'[i, inv q]
        double back = 0.0;    // ... with this value
        if (code == LINK_PUSH) {
            *push_dst(a, q, nd.x + cx, nd.y + cy) = g[q];
            own = false;
        } else if (code == LINK_WRAP) {
            int64_t ty = nd.y + cy;
            if (ty < 0) ty = L.ny - 1;
            else if (ty >= L.ny) ty = 0;
            *push_dst(a, q, nd.x + cx, ty) = g[q];
            own = false;
        } else if (code == LINK_SOLID_BB) {
            // cpu/streaming_kernels.py:40-47 (k_inv = q, ind_nb = target)
            const int64_t t = idx + cx * pitch + cy;
            const double temp = 2 * a.p.w[q] * m.rho * a.p.inv_cs_2 *
                                (double(cx) * a.ux[t] + double(cy) * a.uy[t]);
            back = g[q] - temp;
        } else if (code == LINK_ZERO) {
            // pulled from a ghost node that nothing ever fills
            back = 0.0;
        } else if (code >= LINK_ELEMENT0) {
            const ElementDev e = elements[code - LINK_ELEMENT0];
            if (e.type == 0) {
                // bounce_back, cpu/fluid_boundary_kernels.py:21-25
                back = g[q];
            } else if (e.type == 1) {
                // fixed_velocity_density_based, :51-62
                const double temp = 2 * a.p.w[q] * m.rho * a.p.inv_cs_2 *
                                    (double(cx) * e.v0 + double(cy) * e.v1);
                back = g[q] - temp;
            } else {
                // fixed_pressure_density_based, :126-152
                double unx, uny;
                node_velocity(a, L.at(nd.x + e.normal_x, nd.y + e.normal_y),
                              unx, uny);
                const double ex = m.ux + 0.5 * (m.ux - unx);
                const double ey = m.uy + 0.5 * (m.uy - uny);
                const double u2 = ex * ex + ey * ey;
                const double cu = double(cx) * ex + double(cy) * ey;
                const double temp = 2 * a.p.w[q] * e.scalar *
                                    (1 + 0.5 * a.p.inv_cs_4 * cu * cu -
                                     0.5 * a.p.inv_cs_2 * u2);
                back = -g[q] + temp;
            }
        } else {
            own = false;      // LINK_ZG: written by k_zero_gradient afterwards
        }
        if (own) a.fout[qi * a.fout_plane + out_off] = back;
        // momentum exchanged across the link, for the wall / obstacle force:
        // pop[k] c_k - pop_new[k_inv] c_kinv = c_k (g_k + f'_kinv)
        // (cpu/force_torque_kernels.py:73-81, 128-137)
        if (a.exch) a.exch[i * 8 + (q - 1)] = own ? g[q] + back : 0.0;
    }
}

// zero_gradient (our definition, see oracle/plb_oracle.c bc_zero_gradient)
__global__ void k_zero_gradient(double *fout, int64_t plane,
                                const int32_t *__restrict__ map,
                                const ZgLink *__restrict__ links, int64_t n)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ZgLink l = links[i];
    fout[l.v * plane + lat_off(map, l.dst)] = fout[l.v * plane + lat_off(map, l.src)];
}

// Peer-to-peer hand-shake.  The edge-column and link kernels of step t have
// stored this rank's outgoing face populations into the neighbours' receive
// buffers; those kernels have completed (stream order), so their stores have
// been performed and publishing the step number releases them.
__global__ void k_face_signal(unsigned long long *flag_a,
                              unsigned long long *flag_b,
                              unsigned long long value)
{
    __threadfence_system();
    if (flag_a) *reinterpret_cast<volatile unsigned long long *>(flag_a) = value;
    if (flag_b) *reinterpret_cast<volatile unsigned long long *>(flag_b) = value;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(
    const unsigned long long *p)
{
#ifdef PLB_EMU_RUNTIME
    return *reinterpret_cast<const volatile unsigned long long *>(p);
#else
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p)
                 : "memory");
    return v;
#endif
}

// Delivery of the three populations that crossed a slab face (or the
// periodic-x seam) into the first / last interior column.  With peer-to-peer
// faces every CTA first waits until the neighbour has published step
// `wait_value` in this rank's mailbox; a neighbour that never arrives sets
// *status after `spin_budget` clock cycles instead of hanging the GPU.
__global__ void k_face_unpack(Layout L, double *fout, int64_t plane,
                              const int32_t *__restrict__ map, int64_t x_col,
                              int32_t k0, int32_t k1, int32_t k2,
                              const double *__restrict__ src, int64_t s0,
                              int64_t s1, int64_t s2,
                              const uint8_t *__restrict__ mask,
                              const unsigned long long *wait_flag,
                              unsigned long long wait_value,
                              unsigned long long *status, long long spin_budget)
{
    if (wait_flag) {
        // a time-out is sticky: later steps do not wait again, plb_sync reports
        if (threadIdx.x == 0 &&
            *reinterpret_cast<volatile unsigned long long *>(status) == 0) {
            const long long t0 = clock64();
            while (ld_acquire_sys(wait_flag) < wait_value) {
                if (clock64() - t0 > spin_budget) {
                    atomicExch(status, 1ull);
                    break;
                }
                __nanosleep(200);
            }
        }
        __syncthreads();
    }
    const int64_t y = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (y >= L.ny) return;
    const uint8_t mk = mask[y];
    if (!mk) return;
    const int64_t idx = lat_off(map, L.at(x_col, y));
    if (mk & 1) fout[k0 * plane + idx] = __ldcg(src + s0 + y);
    if (mk & 2) fout[k1 * plane + idx] = __ldcg(src + s1 + y);
    if (mk & 4) fout[k2 * plane + idx] = __ldcg(src + s2 + y);
}

// ---------------------------------------------------------------------------
// initialisation: f = feq(rho, u) on fluid nodes, 0 elsewhere
// cpu/equilibrium_kernels.py:38-78
// ---------------------------------------------------------------------------
__global__ void k_init_pop(KParams p, double *f, const uint8_t *__restrict__ code,
                           const double *__restrict__ rho,
                           const double *__restrict__ ux,
                           const double *__restrict__ uy)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= p.L.plane) return;
    const uint8_t c = code[i];
    double e[Q];
    if (c == NODE_BULK || c == NODE_LINK) {
        Moments m;
        m.rho = rho[i];
        m.ux = ux[i];
        m.uy = uy[i];
        m.fx = m.fy = 0.0;
        feq_all(p, m, m.ux * m.ux + m.uy * m.uy, e);
    } else {
#pragma unroll
        for (int k = 0; k < Q; ++k) e[k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) f[k * p.L.plane + i] = e[k];
}

// ---------------------------------------------------------------------------
// reference layout (padded, array of structures) <-> device planes
// ---------------------------------------------------------------------------
__global__ void k_unpack_rows(Layout L, const double *__restrict__ staging,
                              int ncomp, double *planes, int64_t plane_stride,
                              int64_t row0, int64_t nrows)
{
    const int64_t nyp = L.ny + 2;
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nrows * nyp) return;
    const int64_t r = i / nyp, yy = i - r * nyp;
    const int64_t d = (row0 + r) * L.pitch + L.y0 - 1 + yy;
    for (int c = 0; c < ncomp; ++c)
        planes[c * plane_stride + d] = staging[i * ncomp + c];
}

// zero_mode 1: entries on the ghost ring are reported as 0 (the reference
// never writes pop_fluid_new there; our ghost columns are face staging).
__global__ void k_pack_rows(Layout L, double *staging, int ncomp,
                            const double *__restrict__ planes,
                            int64_t plane_stride, int64_t row0, int64_t nrows,
                            int zero_mode)
{
    const int64_t nyp = L.ny + 2;
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nrows * nyp) return;
    const int64_t r = i / nyp, yy = i - r * nyp;
    const int64_t row = row0 + r;
    const int64_t d = row * L.pitch + L.y0 - 1 + yy;
    const bool ring = row == 0 || row == L.nx + 1 || yy == 0 || yy == nyp - 1;
    for (int c = 0; c < ncomp; ++c)
        staging[i * ncomp + c] =
            (zero_mode && ring) ? 0.0 : planes[c * plane_stride + d];
}

__global__ void k_pack_inner(Layout L, double *staging, int ncomp,
                             const double *__restrict__ planes,
                             int64_t plane_stride, int64_t x0, int64_t nrows)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nrows * L.ny) return;
    const int64_t r = i / L.ny, y = i - r * L.ny;
    const int64_t d = L.at(x0 + r, y);
    for (int c = 0; c < ncomp; ++c)
        staging[i * ncomp + c] = planes[c * plane_stride + d];
}

// ---------------------------------------------------------------------------
// residues, cpu/compute_residues_kernels.py:6-73 (deterministic two-stage sum)
// ---------------------------------------------------------------------------
constexpr int RES_THREADS = 256;

__global__ void __launch_bounds__(RES_THREADS)
k_residue_partial(Layout L, const uint8_t *__restrict__ code,
                  const double *__restrict__ rho, const double *__restrict__ ux,
                  const double *__restrict__ uy, double *rho_old, double *ux_old,
                  double *uy_old, double *partials)
{
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
         i < L.plane; i += stride) {
        const uint8_t c = code[i];
        if (c == NODE_BULK || c == NODE_LINK) {
            const double r = rho[i], ro = rho_old[i];
            const double x = ux[i], xo = ux_old[i];
            const double y = uy[i], yo = uy_old[i];
            acc[0] += (r - ro) * (r - ro);
            acc[1] += ro * ro;
            acc[2] += (x - xo) * (x - xo);
            acc[3] += xo * xo;
            acc[4] += (y - yo) * (y - yo);
            acc[5] += yo * yo;
            rho_old[i] = r;
            ux_old[i] = x;
            uy_old[i] = y;
        }
    }
    __shared__ double sh[6][RES_THREADS];
    for (int j = 0; j < 6; ++j) sh[j][threadIdx.x] = acc[j];
    __syncthreads();
    for (int s = RES_THREADS / 2; s > 0; s >>= 1) {
        if (int(threadIdx.x) < s)
            for (int j = 0; j < 6; ++j)
                sh[j][threadIdx.x] += sh[j][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 6) partials[blockIdx.x * 6 + threadIdx.x] = sh[threadIdx.x][0];
}

__global__ void k_residue_final(const double *__restrict__ partials, int n_blocks,
                                double *out6)
{
    if (threadIdx.x < 6) {
        double s = 0.0;
        for (int b = 0; b < n_blocks; ++b) s += partials[b * 6 + threadIdx.x];
        out6[threadIdx.x] = s;
    }
}

__global__ void k_fill(double *buf, int64_t n, double value)
{
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += stride)
        buf[i] = value;
}

__global__ void k_fill_inner(Layout L, double *plane, double value)
{
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < L.plane;
         i += stride) {
        const int64_t x = i / L.pitch - 1, y = i % L.pitch - L.y0;
        plane[i] = (x >= 0 && x < L.nx && y >= 0 && y < L.ny) ? value : 0.0;
    }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
template <template <int, int, bool> class Launcher, typename... Args>
static void dispatch(int collision, int forcing, bool store, Args &&...args)
{
#define PLB_CASE(C, F)                                                        \
    if (collision == C && forcing == F) {                                     \
        if (store) Launcher<C, F, true>::run(args...);                        \
        else Launcher<C, F, false>::run(args...);                             \
        return;                                                               \
    }
    PLB_CASE(0, 0) PLB_CASE(0, 1) PLB_CASE(0, 2)
    PLB_CASE(1, 0) PLB_CASE(1, 1) PLB_CASE(1, 2)
    PLB_CASE(2, 0) PLB_CASE(2, 1) PLB_CASE(2, 2)
#undef PLB_CASE
}

template <int C, int F, bool S>
struct BulkScalar {
    static void run(const StepArgs &a, int64_t x_begin, int64_t n_rows,
                    cudaStream_t st)
    {
        const int32_t chunks = int32_t((a.p.L.ny + 255) / 256);
        PLB_LAUNCH(SIMPLE, (k_bulk_scalar<C, F, S>), unsigned(n_rows * chunks), 256, st, a,
                   x_begin, chunks);
    }
};

template <int C, int F, bool S>
struct BulkVec2 {
    static void run(const StepArgs &a, int64_t x_begin, int64_t n_rows,
                    cudaStream_t st)
    {
        constexpr int span = 2 * PLB_BLOCK;
        const int32_t chunks = int32_t((a.p.L.ny + span - 1) / span);
        PLB_LAUNCH(COOP, (k_bulk_vec2<C, F, S>), unsigned(n_rows * chunks), PLB_BLOCK, st, a,
                   x_begin, chunks);
    }
};

template <int C, int F, bool S>
struct BulkEdge {
    static void run(const StepArgs &a, int64_t x_begin, int64_t n_rows,
                    cudaStream_t st)
    {
        const int32_t chunks = int32_t((a.p.L.ny + 255) / 256);
        PLB_LAUNCH(SIMPLE, (k_bulk_edge<C, F, S>), unsigned(n_rows * chunks), 256, st, a,
                   x_begin, chunks);
    }
};

template <int C, int F, bool S>
struct Links {
    static void run(const StepArgs &a, const LinkNode *nodes, int64_t n,
                    const ElementDev *el, cudaStream_t st)
    {
        PLB_LAUNCH(SIMPLE, (k_links<C, F, S>), unsigned((n + 127) / 128), 128, st, a, nodes, n, el);
    }
};

int launch_bulk(const StepArgs &a, int64_t x_begin, int64_t x_end, int variant,
                cudaStream_t stream)
{
    const int64_t n_rows = x_end - x_begin;
    if (n_rows <= 0) return 0;
    if (variant == 0)
        dispatch<BulkScalar>(a.collision, a.forcing, a.store != 0, a, x_begin,
                             n_rows, stream);
    else
        dispatch<BulkVec2>(a.collision, a.forcing, a.store != 0, a, x_begin,
                           n_rows, stream);
    return 1;
}

int fused_strips(const Layout &L, int depth)
{
    // strip j delivers y in [span j + depth - 3, span (j + 1) + depth - 3)
    const int span = fused_span(depth);
    return int((L.ny + 3 - depth + span - 1) / span);
}

template <int C, int F, int D>
static void run_fused(const StepArgs &a, const uint8_t *deep, int64_t x_begin,
                      int64_t x_end, int32_t rows_per_chunk, unsigned *work_counter,
                      const TensorMap &tmap, int32_t alternate, cudaStream_t st)
{
    const int32_t strips = fused_strips(a.p.L, D);
    const int64_t chunks = (x_end - x_begin + rows_per_chunk - 1) / rows_per_chunk;
    constexpr int wpb = PLB_FUSED_BLOCK / 32;
    // (alternate & 2: a CTA takes wpb x-adjacent chunks of one strip; items
    // beyond the last chunk return at once)
    int64_t warps = (alternate & 2) ? (chunks + wpb - 1) / wpb * wpb * strips : chunks * strips;
    constexpr int dyn_smem = PLB_FUSED_DYN_SMEM ? fused_smem_bytes(D) : 0;
#ifndef PLB_EMU_RUNTIME
    static bool attributes_set = false;
    if (!attributes_set) {
        // the prefetch ring wants shared memory, nothing here wants L1
        if (PLB_FUSED_STAGES >= 2 || PLB_FUSED_BULK || PLB_FUSED_CARRY_SMEM)
            cudaFuncSetAttribute(k_bulk_fused<C, F, D>,
                                 cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (dyn_smem > 0)
            cudaFuncSetAttribute(k_bulk_fused<C, F, D>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
        attributes_set = true;
    }
#endif
    if (work_counter) {
        // persistent grid: as many CTAs as are resident at once
        static int resident = 0;
        if (!resident) {
#ifdef PLB_EMU_RUNTIME
            resident = 8;
#else
            int per_sm = 0, dev = 0, sms = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bulk_fused<C, F, D>,
                                                          PLB_FUSED_BLOCK, dyn_smem);
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            resident = std::max(1, per_sm * sms);
#endif
        }
        warps = std::min<int64_t>(warps, int64_t(resident) * wpb);
        cudaMemsetAsync(work_counter, 0, sizeof(unsigned), st);
    }
    PLB_LAUNCH_SMEM(COOP, (k_bulk_fused<C, F, D>), unsigned((warps + wpb - 1) / wpb),
                    PLB_FUSED_BLOCK, dyn_smem, st, a, deep, x_begin, x_end, strips,
                    rows_per_chunk, work_counter, tmap, alternate);
}

const char *kernel_build_info()
{
    static char text[400];
    if (!text[0]) {
        const char *ring = PLB_FUSED_TENSOR       ? "tma-tensor"
                           : PLB_FUSED_BULK       ? "tma-bulk"
                           : PLB_FUSED_STAGES < 2 ? "none"
                                                  : "cp.async";
        snprintf(text, sizeof text,
                 "bulk: block=%d ld_mode=%d st_mode=%d; fused: block=%d "
                 "ctas_per_sm=%d/%d/%d/%d (depth2/bgk3/mrt3/mrt4) stages=%d ring=%s carry=%s "
                 "pin=%d smem=%s (%d / %d / %d bytes per cta at depth 2 / 3 / 4)%s",
                 PLB_BLOCK, PLB_LD_MODE, PLB_ST_MODE, PLB_FUSED_BLOCK,
                 fused_min_blocks(2, 2), fused_min_blocks(0, 3), fused_min_blocks(2, 3),
                 fused_min_blocks(2, 4),
                 PLB_FUSED_STAGES, ring, PLB_FUSED_CARRY_SMEM ? "shared" : "registers",
                 PLB_FUSED_PIN, PLB_FUSED_DYN_SMEM ? "dynamic" : "static", fused_smem_bytes(2),
                 fused_smem_bytes(3), fused_smem_bytes(4),
#ifdef PLB_EMU_RUNTIME
                 "; host emulation (test infrastructure)"
#else
                 ""
#endif
        );
    }
    return text;
}

bool fused_needs_tensor_map() { return PLB_FUSED_TENSOR != 0; }

int make_lattice_tensor_map(TensorMap *out, const double *lattice, const Layout &L,
                            char *why, size_t why_len)
{
    memset(out, 0, sizeof *out);
#ifdef PLB_EMU_RUNTIME
    // the emulated copy (tensor_g2s) reads the lattice through these four words
    out->opaque[0] = reinterpret_cast<unsigned long long>(lattice);
    out->opaque[1] = (unsigned long long)L.pitch;
    out->opaque[2] = (unsigned long long)(L.nx + 2);
    out->opaque[3] = (unsigned long long)L.plane;
    (void)why;
    (void)why_len;
    return 0;
#else
    static_assert(sizeof(TensorMap) == sizeof(CUtensorMap), "CUtensorMap is 128 bytes");
    typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                               const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                               const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    // the driver entry point through the runtime: libplb does not link libcuda
    static Encode encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult found;
        const cudaError_t rc = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn,
                                                       cudaEnableDefault, &found);
        if (rc != cudaSuccess || found != cudaDriverEntryPointSuccess || !fn) {
            snprintf(why, why_len, "cuTensorMapEncodeTiled is not available (%s)",
                     cudaGetErrorString(rc));
            return rc != cudaSuccess ? int(rc) : -1;
        }
        encode = reinterpret_cast<Encode>(fn);
    }
    const cuuint64_t dims[3] = {cuuint64_t(L.pitch), cuuint64_t(L.nx + 2), cuuint64_t(Q)};
    const cuuint64_t strides[2] = {cuuint64_t(L.pitch) * sizeof(double),
                                   cuuint64_t(L.plane) * sizeof(double)};
    const cuuint32_t box[3] = {64, 1, cuuint32_t(Q)};
    const cuuint32_t elem[3] = {1, 1, 1};
    int promo = 2;      // 128-byte L2 sectors promoted to 256 bytes: rows are 512 B runs
    if (const char *v = getenv("PLB_TMA_L2_PROMOTION")) promo = atoi(v);
    const CUresult rc = encode(
        reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
        const_cast<double *>(lattice), dims, strides, box, elem,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
        promo == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
        : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
        : promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                     : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        snprintf(why, why_len,
                 "cuTensorMapEncodeTiled failed with CUresult %d (pitch %lld, rows %lld, "
                 "plane %lld doubles)", int(rc), (long long)L.pitch, (long long)(L.nx + 2),
                 (long long)L.plane);
        return int(rc);
    }
    return 0;
#endif
}

int launch_bulk_fused(const StepArgs &a, const uint8_t *deep, int depth,
                      int64_t x_begin, int64_t x_end, int32_t rows_per_chunk,
                      unsigned *work_counter, const TensorMap *tmap_ptr,
                      cudaStream_t stream)
{
    if (x_end <= x_begin || depth < 2 || depth > 4) return 0;
    // PLB_FUSED_ALTERNATE=0: every chunk marches towards larger x (A/B switch)
    // bit 0: odd chunks march towards smaller x; bit 1: the warps of a CTA take
    // x-adjacent chunks of one strip instead of y-adjacent strips of one chunk
    const char *alt_env = getenv("PLB_FUSED_ALTERNATE");
    const int32_t alternate = alt_env ? atoi(alt_env) & 3 : PLB_FUSED_ALTERNATE_DEFAULT;
    static const TensorMap no_map = {};
    if (PLB_FUSED_TENSOR && !tmap_ptr) return 0;
    const TensorMap &tmap = tmap_ptr ? *tmap_ptr : no_map;
#define PLB_CASE(C, F)                                                        \
    if (a.collision == C && a.forcing == F) {                                 \
        if (depth == 2)                                                       \
            run_fused<C, F, 2>(a, deep, x_begin, x_end, rows_per_chunk,       \
                               work_counter, tmap, alternate, stream);        \
        else if (depth == 3)                                                  \
            run_fused<C, F, 3>(a, deep, x_begin, x_end, rows_per_chunk,       \
                               work_counter, tmap, alternate, stream);        \
        else                                                                  \
            run_fused<C, F, 4>(a, deep, x_begin, x_end, rows_per_chunk,       \
                               work_counter, tmap, alternate, stream);        \
        return 1;                                                             \
    }
    PLB_CASE(0, 0) PLB_CASE(0, 1) PLB_CASE(0, 2)
    PLB_CASE(1, 0) PLB_CASE(1, 1) PLB_CASE(1, 2)
    PLB_CASE(2, 0) PLB_CASE(2, 1) PLB_CASE(2, 2)
#undef PLB_CASE
    return 0;
}

int launch_bulk_edge(const StepArgs &a, int64_t x_begin, int64_t x_end,
                     cudaStream_t stream)
{
    const int64_t n_rows = x_end - x_begin;
    if (n_rows <= 0) return 0;
    dispatch<BulkEdge>(a.collision, a.forcing, a.store != 0, a, x_begin, n_rows,
                       stream);
    return 1;
}

int launch_face_signal(unsigned long long *flag_a, unsigned long long *flag_b,
                       unsigned long long value, cudaStream_t stream)
{
    PLB_LAUNCH(SIMPLE, (k_face_signal), 1, 1, stream, flag_a, flag_b, value);
    return 1;
}

int launch_links(const StepArgs &a, const LinkNode *nodes, int64_t n_nodes,
                 const ElementDev *elements, cudaStream_t stream)
{
    if (n_nodes <= 0) return 0;
    dispatch<Links>(a.collision, a.forcing, a.store != 0, a, nodes, n_nodes,
                    elements, stream);
    return 1;
}

int launch_zero_gradient(double *fout, int64_t plane, const int32_t *map,
                         const ZgLink *links, int64_t n_links, cudaStream_t stream)
{
    if (n_links <= 0) return 0;
    PLB_LAUNCH(SIMPLE, (k_zero_gradient), unsigned((n_links + 127) / 128), 128, stream, fout,
               plane, map, links, n_links);
    return 1;
}

int launch_face_unpack(const Layout &L, double *fout, int64_t plane,
                       const int32_t *map, int64_t x_col,
                       const int32_t dirs[3], const double *src,
                       int64_t src_stride0, int64_t src_stride1,
                       int64_t src_stride2, const uint8_t *mask,
                       cudaStream_t stream, const unsigned long long *wait_flag,
                       unsigned long long wait_value, unsigned long long *status,
                       long long spin_budget)
{
    PLB_LAUNCH(COOP, (k_face_unpack), unsigned((L.ny + 255) / 256), 256, stream, L, fout,
               plane, map, x_col, dirs[0], dirs[1], dirs[2], src, src_stride0, src_stride1, src_stride2,
               mask, wait_flag, wait_value, status, spin_budget);
    return 1;
}

int launch_init_pop(const KParams &p, double *f, const uint8_t *code,
                    const double *rho, const double *ux, const double *uy,
                    cudaStream_t stream)
{
    PLB_LAUNCH(SIMPLE, (k_init_pop), unsigned((p.L.plane + 255) / 256), 256, stream, p, f,
               code, rho, ux, uy);
    return 1;
}

int launch_unpack_rows(const Layout &L, const double *staging, int ncomp,
                       double *planes, int64_t plane_stride, int64_t row0,
                       int64_t nrows, cudaStream_t stream)
{
    const int64_t n = nrows * (L.ny + 2);
    PLB_LAUNCH(SIMPLE, (k_unpack_rows), unsigned((n + 255) / 256), 256, stream, L, staging,
               ncomp, planes, plane_stride, row0, nrows);
    return 1;
}

int launch_pack_rows(const Layout &L, double *staging, int ncomp,
                     const double *planes, int64_t plane_stride, int64_t row0,
                     int64_t nrows, const uint8_t *, int zero_mode,
                     cudaStream_t stream)
{
    const int64_t n = nrows * (L.ny + 2);
    PLB_LAUNCH(SIMPLE, (k_pack_rows), unsigned((n + 255) / 256), 256, stream, L, staging,
               ncomp, planes, plane_stride, row0, nrows, zero_mode);
    return 1;
}

int launch_pack_inner(const Layout &L, double *staging, int ncomp,
                      const double *planes, int64_t plane_stride, int64_t x0,
                      int64_t nrows, cudaStream_t stream)
{
    const int64_t n = nrows * L.ny;
    PLB_LAUNCH(SIMPLE, (k_pack_inner), unsigned((n + 255) / 256), 256, stream, L, staging,
               ncomp, planes, plane_stride, x0, nrows);
    return 1;
}

int launch_residue(const Layout &L, const uint8_t *code, const double *rho,
                   const double *ux, const double *uy, double *rho_old,
                   double *ux_old, double *uy_old, double *partials,
                   int n_blocks, double *out6, cudaStream_t stream)
{
    PLB_LAUNCH(COOP, (k_residue_partial), n_blocks, RES_THREADS, stream, L, code, rho, ux,
               uy, rho_old, ux_old, uy_old, partials);
    PLB_LAUNCH(SIMPLE, (k_residue_final), 1, 32, stream, partials, n_blocks, out6);
    return 2;
}

int launch_fill_inner(const Layout &L, double *plane, double value, cudaStream_t stream)
{
    PLB_LAUNCH(SIMPLE, (k_fill_inner), 148 * 8, 256, stream, L, plane, value);
    return 1;
}

int launch_fill(double *buf, int64_t n, double value, cudaStream_t stream)
{
    PLB_LAUNCH(SIMPLE, (k_fill), 148 * 8, 256, stream, buf, n, value);
    return 1;
}

}  // namespace plb
