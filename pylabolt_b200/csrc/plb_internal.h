// plb_internal.h -- structures shared by the host side (plb_api.cu) and the
// kernels (plb_kernels.cu) of libplb.  Not part of the ABI (include/plb.h is).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace plb {

constexpr int Q = 9;
// most steps k_bulk_fused advances per pass (instantiated for 2 .. this)
constexpr int MAX_FUSE_DEPTH = 4;

// D2Q9 direction tables, pylabolt/base/lattice.py:50-60.
#define PLB_CX_LIST {0, 1, 0, -1, 0, 1, -1, -1, 1}
#define PLB_CY_LIST {0, 0, 1, 0, -1, 1, 1, -1, -1}
#define PLB_INV_LIST {0, 3, 4, 1, 2, 7, 8, 5, 6}

// Device layout of one scalar plane (a population k, rho, ux, uy or the node
// codes): (nx + 2) rows of `pitch` elements, y contiguous.  Interior node
// (x, y), x in [0, nx), y in [0, ny), lives at (x + 1) * pitch + y0 + y, so
// y = 0 is 128-byte aligned (y0 = 16 doubles) and every row start is too
// (pitch is a multiple of 16).  The ghost ring of the reference layout maps
// to x = -1, nx and y = -1, ny.
struct Layout {
    int64_t nx, ny;
    int64_t pitch;   // elements per row
    int64_t plane;   // elements per plane = (nx + 2) * pitch
    int32_t y0;      // column of interior y = 0
    __host__ __device__ int64_t at(int64_t x, int64_t y) const {
        return (x + 1) * pitch + y0 + y;
    }
};

// Constants used verbatim by the kernels (values computed by the caller the
// way the reference computes them, see plb_config in include/plb.h).
// Derived constants of the two-stress-moment MRT path (collide_mrt_stress in
// plb_collide.cuh), computed once on the host from the values above.
struct MrtStress {
    double hgx, hgy;     // g / 2
    double k2, k4;       // 1 / (2 cs^2), 1 / (2 cs^4)
    double k7, k8;       // w_1 / cs^4, 4 w_5 / cs^4
    double qa, qb;       // (1 - s_7) / 4, (1 - s_8) / 4
    double k4cg[4];      // k4 * (c . g) for c = (1,0), (0,1), (1,1), (1,-1)
    double k2cg[4];      // k2 * (c . g)
};

// Nine-rate MRT in moment space (collide_mrt_moments in plb_collide.cuh): the
// moments of the second-order equilibrium and of the Guo source are
// polynomials in (rho, u, F) whose coefficients are lattice sums of the rows
// of M against the weights -- computed once on the host from the caller's own
// w / inv_cs_2 / inv_cs_4 (so that nothing assumes cs^2 = 1/3 exactly).
//   rows 0-2 (rho, e, eps):  m_eq = rho (A + Qx ux^2 + Qy uy^2)
//                            m_F  = -GA (u.F) + Gx ux Fx + Gy uy Fy
//   rows 3-6 (jx, qx, jy, qy): m_eq = rho B u_a,  m_F = B F_a   (a = x, x, y, y)
//   row 7 (pxx): m_eq = rho (Q7x ux^2 + Q7y uy^2),  m_F = 2 (Q7x ux Fx + Q7y uy Fy)
//   row 8 (pxy): m_eq = rho Q8 ux uy,               m_F = Q8 (ux Fy + uy Fx)
struct MrtMoments {
    double A[3], Qx[3], Qy[3];
    double GA[3], Gx[3], Gy[3];
    double B[4];
    double Q7x, Q7y, Q8;
    double sn[Q];        // s_r / |row r|^2
    double hn[Q];        // (1 - s_r / 2) / |row r|^2
};

struct KParams {
    Layout L;
    double omega, gx, gy, inv_cs_2, inv_cs_4, eps;
    double w[Q];
    double s[Q];   // MRT relaxation rates
    MrtStress mrt;
    MrtMoments mrtm;
};

// Node classes in the code plane.
enum NodeCode : uint8_t {
    NODE_BULK = 0,    // fluid, all 8 pushes go to plain neighbours
    NODE_LINK = 1,    // fluid, handled by the link-list kernel
    NODE_SOLID = 2,
    NODE_GHOST = 3    // ghost ring and alignment padding
};

// Per-direction treatment of a link node's outgoing population q = 1..8,
// one byte each, packed little-endian into a uint64 (byte q-1).
enum LinkCode : uint8_t {
    LINK_PUSH = 0,       // f'[i + c_q, q] = g_q
    LINK_SOLID_BB = 1,   // halfway bounce back off a solid node (moving wall)
    LINK_ZERO = 2,       // uncovered non-periodic ghost: f'[i, inv q] = 0
    LINK_WRAP = 3,       // push with the y index wrapped (y-periodic)
    LINK_ZG = 4,         // zero_gradient element: written by the zg pass
    LINK_ELEMENT0 = 8    // 8 + e: boundary element e (bounce_back,
                         // fixed_velocity, fixed_pressure)
};
constexpr int MAX_ELEMENTS = 247;

struct ElementDev {
    int32_t type;        // enum plb_bc_type
    int32_t normal_x, normal_y;
    int32_t pad;
    double v0, v1;       // vector_fluid
    double scalar;       // scalar_fluid
};

struct LinkNode {
    int32_t x, y;        // interior coordinates
    uint64_t links;      // 8 LinkCode bytes
};

// One zero_gradient copy: f'[v][dst] = f'[v][src].
struct ZgLink {
    int64_t dst, src;
    int32_t v;
    int32_t pad;
};

struct StepArgs {
    KParams p;
    const double *fin;
    double *fout;
    const uint8_t *code;
    double *rho, *ux, *uy;   // moment planes (ux/uy also hold solid velocities)
    double *exch;            // per link node x 8: momentum exchange, or null
    int32_t collision, forcing, store;
    // Peer-to-peer slab faces: where the ghost row x = -1 (populations 3, 6, 7)
    // and x = nx (populations 1, 5, 8) live -- three rows `face_stride` apart
    // in the NEIGHBOUR RANK's receive buffer, written directly over NVLink by
    // the edge-column and link kernels.  Null: this rank's own ghost rows.
    double *face_lo = nullptr, *face_hi = nullptr;
    int64_t face_stride = 0;
    // Compact lattices (the scratch lattices of the several-steps-per-pass
    // path hold time t + 1 (t + 2) on the O(perimeter) listed nodes only):
    // `*_map` translates a node index into the compact lattice, 16 nodes (one
    // 128-byte line) at a time, `*_plane` is the stride between populations.
    // Null map: an ordinary full lattice (plane = L.plane).  Only the
    // list-driven kernels (k_links, k_zero_gradient, k_face_unpack) and
    // push_dst translate; the bulk kernels always work on full lattices.
    const int32_t *fin_map = nullptr, *fout_map = nullptr;
    int64_t fin_plane = 0, fout_plane = 0;
    // Byte offset of the slot population k of a node is pushed into in a full
    // lattice, without the c_y part: (k * plane + c_x[k] * pitch) * 8 -- nine
    // loop-invariant 64-bit constants the fused kernel adds to its running row
    // pointer straight from the constant bank.
    int64_t push_off[Q] = {};
};

// node index -> offset inside one population plane of a (possibly compact) lattice
__host__ __device__ inline int64_t lat_off(const int32_t *map, int64_t idx)
{
    return map ? ((int64_t(map[idx >> 4]) << 4) | (idx & 15)) : idx;
}

// A lattice as the TMA unit sees it: a rank-3 tensor of doubles, (column,
// row, population) = (pitch, nx + 2, Q) with the strides of Layout.  128
// opaque bytes (a CUtensorMap), 64-byte aligned, handed to the kernel as a
// __grid_constant__ parameter.
struct alignas(64) TensorMap {
    unsigned long long opaque[16];
};
// Encodes the descriptor of `lattice` for boxes of (64 columns, 1 row, Q
// populations); 0 on success, else a cudaError_t / CUresult-style code and a
// message in `why`.
int make_lattice_tensor_map(TensorMap *out, const double *lattice, const Layout &L,
                            char *why, size_t why_len);
// true if k_bulk_fused of this build fills its ring with tensor copies
bool fused_needs_tensor_map();

// ---- launchers implemented in plb_kernels.cu ---------------------------
// All return the number of kernels launched (0 if nothing to do).
int launch_bulk(const StepArgs &a, int64_t x_begin, int64_t x_end, int variant,
                cudaStream_t stream);
// `depth` (2 .. MAX_FUSE_DEPTH) steps in one pass over the nodes of columns
// [x_begin, x_end) whose deep[] value is >= depth - 1 (fin = time t, fout =
// time t + depth); deep[] is one byte per node: the Chebyshev distance up to
// which all neighbours are bulk nodes, capped at the largest depth in use - 1.
// work_counter: null = one work item per warp; else a device word (zeroed by
// the launcher) from which the warps of a persistent grid draw their items.
// tmap: the TMA descriptor of lattice a.fin (make_lattice_tensor_map), through
// which a warp fetches a whole row -- nine populations x 64 nodes -- with ONE
// tensor copy; null only in builds whose ring is not tensor-filled.
int launch_bulk_fused(const StepArgs &a, const uint8_t *deep, int depth,
                      int64_t x_begin, int64_t x_end, int32_t rows_per_chunk,
                      unsigned *work_counter, const TensorMap *tmap,
                      cudaStream_t stream);
int fused_strips(const Layout &L, int depth);
// compile-time configuration of the kernels (plb_build_info)
const char *kernel_build_info();
// One slab-edge column with the face redirection of StepArgs::face_lo/hi.
int launch_bulk_edge(const StepArgs &a, int64_t x_begin, int64_t x_end,
                     cudaStream_t stream);
int launch_links(const StepArgs &a, const LinkNode *nodes, int64_t n_nodes,
                 const ElementDev *elements, cudaStream_t stream);
// Peer-to-peer face hand-shake: publishes `value` in the neighbours' mailboxes
// (either pointer may be null).
int launch_face_signal(unsigned long long *flag_a, unsigned long long *flag_b,
                       unsigned long long value, cudaStream_t stream);
int launch_zero_gradient(double *fout, int64_t plane, const int32_t *map,
                         const ZgLink *links, int64_t n_links, cudaStream_t stream);
// Copies the three face populations from `src` (three rows of ny doubles,
// src_stride apart) into column x_col of fout where mask bit j is set
// (fout: stride `plane` between populations, node index through `map`).
int launch_face_unpack(const Layout &L, double *fout, int64_t plane,
                       const int32_t *map, int64_t x_col,
                       const int32_t dirs[3], const double *src,
                       int64_t src_stride0, int64_t src_stride1,
                       int64_t src_stride2, const uint8_t *mask,
                       cudaStream_t stream,
                       const unsigned long long *wait_flag = nullptr,
                       unsigned long long wait_value = 0,
                       unsigned long long *status = nullptr,
                       long long spin_budget = 0);
int launch_init_pop(const KParams &p, double *f, const uint8_t *code,
                    const double *rho, const double *ux, const double *uy,
                    cudaStream_t stream);
// reference-layout (padded AoS) <-> device planes, rows [row0, row0 + nrows)
// of the padded array (row = x + 1).
int launch_unpack_rows(const Layout &L, const double *staging, int ncomp,
                       double *planes, int64_t plane_stride, int64_t row0,
                       int64_t nrows, cudaStream_t stream);
int launch_pack_rows(const Layout &L, double *staging, int ncomp,
                     const double *planes, int64_t plane_stride, int64_t row0,
                     int64_t nrows, const uint8_t *code, int zero_mode,
                     cudaStream_t stream);
int launch_pack_inner(const Layout &L, double *staging, int ncomp,
                      const double *planes, int64_t plane_stride, int64_t x0,
                      int64_t nrows, cudaStream_t stream);
int launch_residue(const Layout &L, const uint8_t *code, const double *rho,
                   const double *ux, const double *uy, double *rho_old,
                   double *ux_old, double *uy_old, double *partials,
                   int n_blocks, double *out6, cudaStream_t stream);
int launch_fill(double *buf, int64_t n, double value, cudaStream_t stream);
// plane[node] = value on interior nodes, 0 on the ghost ring and the padding
int launch_fill_inner(const Layout &L, double *plane, double value, cudaStream_t stream);

}  // namespace plb
