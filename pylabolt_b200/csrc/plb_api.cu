// plb_api.cu -- host side of libplb: the C ABI of include/plb.h.
//
// Owns the device lattices, classifies nodes into bulk / link / solid / ghost
// from the reference's own flags and boundary elements, builds the link lists
// and drives the per-step kernel sequence (optionally with a slab-face
// exchange over NCCL, loaded at run time with dlopen so that a single-GPU
// process needs no NCCL at all).
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <nccl.h>   // types and prototypes only; symbols come from dlopen

#include "../../include/plb.h"
#include "plb_internal.h"

using namespace plb;

namespace {

thread_local std::string g_error;

int fail(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                        \
    do {                                                                      \
        cudaError_t err__ = (expr);                                           \
        if (err__ != cudaSuccess)                                             \
            return fail(PLB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,         \
                        cudaGetErrorString(err__), __FILE__, __LINE__);       \
    } while (0)

#ifndef PLB_FUSE_DEFAULT
#define PLB_FUSE_DEFAULT 1
#endif

constexpr size_t XBUF_MAILBOX = 256;   // bytes reserved for the p2p mailbox words

constexpr int CX[Q] = PLB_CX_LIST;
constexpr int CY[Q] = PLB_CY_LIST;
constexpr int INV[Q] = PLB_INV_LIST;

// ---- NCCL through dlopen ---------------------------------------------------
struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.lib) return PLB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return fail(PLB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                      \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name)); \
    if (!g_nccl.field) return fail(PLB_ERR_NCCL, "NCCL symbol %s missing", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = lib;
    return PLB_OK;
}

#define NCCL_TRY(expr)                                                        \
    do {                                                                      \
        ncclResult_t r__ = (expr);                                            \
        if (r__ != ncclSuccess)                                               \
            return fail(PLB_ERR_NCCL, "%s failed: %s", #expr,                 \
                        g_nccl.GetErrorString(r__));                          \
    } while (0)

struct ElementHost {
    int32_t type;
    std::vector<int64_t> nodes;
    int64_t out[3], inv[3], normal[2];
    double vector[2], scalar;
};

}  // namespace

struct plb_solver {
    plb_config cfg;
    Layout L;
    KParams kp;
    int variant = 1;                 // 0 scalar, 1 vec2 (env PLB_KERNEL)
    int kernel_collision = 0;        // 0 BGK, 1 MRT (free rates), 2 MRT (reference rates)

    cudaStream_t stream = nullptr, edge_stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_edge = nullptr, ev_comm = nullptr;
    cudaEvent_t ev_filled[2] = {}, ev_drained[2] = {};   // staging halves
    cudaEvent_t events[8] = {};

    double *f[2] = {nullptr, nullptr};   // two lattices, 9 planes each
    TensorMap f_tmap[2];                 // their TMA descriptors (k_bulk_fused's row fetch)
    bool tmap_ok = true;
    int cur = 0;
    double *mom = nullptr;               // rho, ux, uy planes
    double *mom_old = nullptr;           // residue field_old (lazy)
    double *res_partials = nullptr, *res_out = nullptr;
    uint8_t *code = nullptr;
    double *staging = nullptr;
    size_t staging_bytes = 0;
    double *flush_buf = nullptr;
    double *exch_dev = nullptr;          // n_links x 8 momentum exchange (lazy)
    std::vector<int64_t> link_inds;      // padded flat index of each link node

    std::vector<uint8_t> solid_host;
    std::vector<ElementHost> elements;
    bool finalized = false;

    ElementDev *elements_dev = nullptr;
    LinkNode *links_dev = nullptr;
    int64_t n_links = 0;
    std::vector<std::pair<ZgLink *, int64_t>> zg_dev;
    uint8_t *mask_left = nullptr, *mask_right = nullptr;
    int64_t n_bulk = 0, n_solid = 0;
    int64_t n_bulk_edge[2] = {0, 0};     // bulk nodes in columns 0 and nx - 1

    ncclComm_t comm = nullptr;
    int rank = 0, n_ranks = 1, left_rank = -1, right_rank = -1;
    double *recv_left = nullptr, *recv_right = nullptr;   // 3 * ny each (NCCL faces)

    // Peer-to-peer faces (see setup_p2p): one exchange buffer per rank,
    //   [mailbox: 32 x u64][recv_left: 2 x 3 x pitch doubles][recv_right: same]
    // exported with CUDA IPC and mapped by both neighbours.
    bool p2p = false;
    unsigned char *xbuf = nullptr;
    unsigned char *peer_left = nullptr, *peer_right = nullptr;   // their xbuf
    long long spin_budget = 0;

    // Several steps per pass (step_fused): PLB_FUSE = 0 off, 1 when the
    // geometry qualifies (default), 2 whenever there is a deep node at all;
    // PLB_FUSE_DEPTH = steps per pass, 2 .. MAX_FUSE_DEPTH (default per
    // collision model, see plb_create).
    int fuse_mode = PLB_FUSE_DEFAULT;
    int fuse_depth = 3;
    bool fuse_depth_from_env = false;
    int fused_depth_ok = 0;              // largest depth the geometry qualifies for (0: none)
    double *f_mid[MAX_FUSE_DEPTH - 1] = {};   // compact scratch lattices (lazy)
    // node index -> compact scratch lattice, one entry per 16 nodes (lat_off);
    // entry 0 is a spare segment that absorbs anything unlisted
    int32_t *mid_map_dev = nullptr;
    std::vector<int32_t> mid_map_host;
    int64_t mid_plane = 0;               // doubles per population of a scratch lattice
    uint8_t *deep_dev = nullptr;         // distance to the nearest non-bulk node - 1, capped at fuse_depth - 1
    // list p (0-based) = nodes of list pass p + 1 of a depth-d group,
    // lists[d - 2][p]; the last one holds the fluid nodes that are not deep enough
    LinkNode *lists_dev[MAX_FUSE_DEPTH - 1][MAX_FUSE_DEPTH] = {};
    int64_t n_lists[MAX_FUSE_DEPTH - 1][MAX_FUSE_DEPTH] = {};
    int64_t n_deep[MAX_FUSE_DEPTH - 1] = {};   // nodes the depth-2 / -3 / -4 kernel advances
    // rows a warp marches over per work item, for two / three / four steps per
    // pass (PLB_FUSED_ROWS sets all)
    int32_t fused_rows[MAX_FUSE_DEPTH - 1] = {32, 64, 64};
    unsigned *work_counter = nullptr;    // PLB_FUSED_DYNAMIC=1: persistent grid, work queue
    int64_t pending = 0;                 // plain steps held back for grouping
    int64_t groups_done[MAX_FUSE_DEPTH - 1] = {};

    int64_t launches = 0;
    int64_t steps_done = 0;

    bool profile = false;
    std::vector<cudaEvent_t> prof_events;   // start/stop pairs
    size_t prof_used = 0;

    double *rho() const { return mom; }
    double *ux() const { return mom + L.plane; }
    double *uy() const { return mom + 2 * L.plane; }
};

namespace {

// Host <-> device transfers go through two staging halves so that the PCIe
// copy of one chunk (copy stream) overlaps the layout transpose of the other
// (main stream): reference layout (padded / inner, ncomp interleaved) on the
// host side, planes on the device side.
constexpr size_t CHUNK_BYTES = size_t(64) << 20;

int ensure_staging(plb_solver *s, size_t half_bytes)
{
    if (!s->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_filled[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->ev_drained[i], cudaEventDisableTiming));
        }
    }
    if (s->staging_bytes >= half_bytes) return PLB_OK;
    if (s->staging) CUDA_TRY(cudaFree(s->staging));
    s->staging = nullptr;
    s->staging_bytes = 0;
    CUDA_TRY(cudaMalloc(&s->staging, 2 * half_bytes));
    s->staging_bytes = half_bytes;
    return PLB_OK;
}

// host rows [0, n_rows) of row_elems doubles -> unpack(staging, row0, nrows)
template <typename Unpack>
int upload_rows(plb_solver *s, const double *host, int64_t row_elems, int64_t n_rows,
                Unpack unpack)
{
    const int64_t rows_per_chunk =
        std::max<int64_t>(1, int64_t(CHUNK_BYTES / (row_elems * sizeof(double))));
    const size_t half = size_t(rows_per_chunk) * row_elems * sizeof(double);
    if (int rc = ensure_staging(s, half)) return rc;
    int64_t chunk = 0;
    for (int64_t row0 = 0; row0 < n_rows; row0 += rows_per_chunk, ++chunk) {
        const int h = int(chunk & 1);
        double *stage = s->staging + h * (half / sizeof(double));
        const int64_t nrows = std::min(rows_per_chunk, n_rows - row0);
        if (chunk >= 2) CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->ev_drained[h], 0));
        CUDA_TRY(cudaMemcpyAsync(stage, host + row0 * row_elems,
                                 size_t(nrows) * row_elems * sizeof(double),
                                 cudaMemcpyHostToDevice, s->copy_stream));
        CUDA_TRY(cudaEventRecord(s->ev_filled[h], s->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_filled[h], 0));
        s->launches += unpack(stage, row0, nrows);
        CUDA_TRY(cudaEventRecord(s->ev_drained[h], s->stream));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return PLB_OK;
}

// pack(staging, row0, nrows) -> host rows [0, n_rows) of row_elems doubles
template <typename Pack>
int download_rows(plb_solver *s, double *host, int64_t row_elems, int64_t n_rows,
                  Pack pack)
{
    const int64_t rows_per_chunk =
        std::max<int64_t>(1, int64_t(CHUNK_BYTES / (row_elems * sizeof(double))));
    const size_t half = size_t(rows_per_chunk) * row_elems * sizeof(double);
    if (int rc = ensure_staging(s, half)) return rc;
    int64_t chunk = 0;
    for (int64_t row0 = 0; row0 < n_rows; row0 += rows_per_chunk, ++chunk) {
        const int h = int(chunk & 1);
        double *stage = s->staging + h * (half / sizeof(double));
        const int64_t nrows = std::min(rows_per_chunk, n_rows - row0);
        if (chunk >= 2) CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_drained[h], 0));
        s->launches += pack(stage, row0, nrows);
        CUDA_TRY(cudaEventRecord(s->ev_filled[h], s->stream));
        CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->ev_filled[h], 0));
        CUDA_TRY(cudaMemcpyAsync(host + row0 * row_elems, stage,
                                 size_t(nrows) * row_elems * sizeof(double),
                                 cudaMemcpyDeviceToHost, s->copy_stream));
        CUDA_TRY(cudaEventRecord(s->ev_drained[h], s->copy_stream));
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return PLB_OK;
}

int upload_padded(plb_solver *s, const double *host, int ncomp, double *planes,
                  int64_t plane_stride)
{
    const Layout &L = s->L;
    return upload_rows(s, host, (L.ny + 2) * ncomp, L.nx + 2,
                       [&](double *stage, int64_t row0, int64_t nrows) {
                           return launch_unpack_rows(L, stage, ncomp, planes, plane_stride,
                                                     row0, nrows, s->stream);
                       });
}

int download_padded(plb_solver *s, double *host, int ncomp, const double *planes,
                    int64_t plane_stride, int zero_mode)
{
    const Layout &L = s->L;
    return download_rows(s, host, (L.ny + 2) * ncomp, L.nx + 2,
                         [&](double *stage, int64_t row0, int64_t nrows) {
                             return launch_pack_rows(L, stage, ncomp, planes, plane_stride,
                                                     row0, nrows, s->code, zero_mode,
                                                     s->stream);
                         });
}

int download_inner(plb_solver *s, double *host, int ncomp, const double *planes,
                   int64_t plane_stride)
{
    const Layout &L = s->L;
    return download_rows(s, host, L.ny * ncomp, L.nx,
                         [&](double *stage, int64_t x0, int64_t nrows) {
                             return launch_pack_inner(L, stage, ncomp, planes, plane_stride,
                                                      x0, nrows, s->stream);
                         });
}

// Per-direction link codes of fluid node (x, y); see LinkCode.
struct Classifier {
    const plb_solver *s;
    int64_t nx, ny, nyp;
    const uint8_t *solid;
    const std::unordered_map<int64_t, int32_t> *element_of;   // ind*9+o -> e
    const std::unordered_set<int64_t> *zg_cover;

    bool is_solid(int64_t x, int64_t y) const
    {
        return solid[(x + 1) * nyp + (y + 1)] != 0;
    }
    uint64_t links(int64_t x, int64_t y) const
    {
        const int64_t ind = (x + 1) * nyp + (y + 1);
        const bool edge = x == 0 || y == 0 || x == nx - 1 || y == ny - 1;
        uint64_t out = 0;
        for (int q = 1; q < Q; ++q) {
            const int64_t tx = x + CX[q], ty = y + CY[q];
            uint8_t code = LINK_PUSH;
            bool done = false;
            if (edge) {
                const int64_t key = ind * Q + q;
                if (zg_cover->count(key)) {
                    code = LINK_ZG;
                    done = true;
                } else {
                    auto it = element_of->find(key);
                    if (it != element_of->end()) {
                        code = uint8_t(LINK_ELEMENT0 + it->second);
                        done = true;
                    }
                }
            }
            if (!done) {
                if (is_solid(tx, ty)) {
                    code = LINK_SOLID_BB;
                } else {
                    const bool x_out = tx < 0 || tx >= nx;
                    const bool y_out = ty < 0 || ty >= ny;
                    if (x_out || y_out) {
                        bool reachable = true;
                        if (x_out)
                            reachable = tx < 0 ? s->cfg.left_neighbor != 0
                                               : s->cfg.right_neighbor != 0;
                        if (y_out && !s->cfg.y_periodic) reachable = false;
                        code = !reachable ? LINK_ZERO
                                          : (y_out ? LINK_WRAP : LINK_PUSH);
                    }
                }
            }
            out |= uint64_t(code) << (8 * (q - 1));
        }
        return out;
    }
};

int run_zero_gradient(plb_solver *s, const StepArgs &a, cudaStream_t st)
{
    for (auto &z : s->zg_dev)
        s->launches += launch_zero_gradient(a.fout, a.fout_plane, a.fout_map, z.first,
                                            z.second, st);
    return PLB_OK;
}

// Launches the bulk kernel on columns [x0, x1).  `timed`: the launch is the
// dominant one of the step (main stream) and, while profiling is enabled, is
// bracketed by CUDA events on its stream.  `edge`: a slab-edge column whose
// face pushes go to the neighbour's memory (peer-to-peer faces).
int bulk_timed(plb_solver *s, const StepArgs &a, int64_t x0, int64_t x1,
               cudaStream_t stream, bool timed, bool edge = false)
{
    if (x1 <= x0) return PLB_OK;
    const bool prof = s->profile && timed;
    if (prof) {
        if (s->prof_used + 2 > s->prof_events.size()) {
            for (int i = 0; i < 2; ++i) {
                cudaEvent_t e;
                CUDA_TRY(cudaEventCreate(&e));
                s->prof_events.push_back(e);
            }
        }
        CUDA_TRY(cudaEventRecord(s->prof_events[s->prof_used], stream));
    }
    s->launches += edge ? launch_bulk_edge(a, x0, x1, stream)
                        : launch_bulk(a, x0, x1, s->variant, stream);
    if (prof) {
        CUDA_TRY(cudaEventRecord(s->prof_events[s->prof_used + 1], stream));
        s->prof_used += 2;
    }
    return PLB_OK;
}

StepArgs step_args(const plb_solver *s, const double *fin, double *fout)
{
    StepArgs a;
    a.p = s->kp;
    a.fin = fin;
    a.fout = fout;
    a.code = s->code;
    a.rho = s->rho();
    a.ux = s->ux();
    a.uy = s->uy();
    a.collision = s->kernel_collision;
    a.forcing = s->cfg.forcing;
    a.store = 0;
    a.exch = nullptr;
    a.fin_plane = a.fout_plane = s->L.plane;
    {
        static const int cx[Q] = PLB_CX_LIST;
        for (int k = 0; k < Q; ++k)
            a.push_off[k] = (int64_t(k) * s->L.plane + cx[k] * s->L.pitch) *
                            int64_t(sizeof(double));
    }
    for (int m = 0; m < MAX_FUSE_DEPTH - 1; ++m) {
        if (!s->f_mid[m]) continue;
        if (fin == s->f_mid[m]) {
            a.fin_map = s->mid_map_dev;
            a.fin_plane = s->mid_plane;
        }
        if (fout == s->f_mid[m]) {
            a.fout_map = s->mid_map_dev;
            a.fout_plane = s->mid_plane;
        }
    }
    return a;
}

// The O(perimeter) part of step number t, all on the edge stream: with slab
// faces the two edge columns first (they feed the faces; `edge_columns`), then
// the listed nodes, then the face exchange (peer-to-peer stores or NCCL
// between ranks, this rank's own ghost rows for a single-rank periodic seam)
// and its delivery into a.fout.  Returns in [x_lo, x_hi) the columns that are
// left for the bulk kernel.
int edge_chain(plb_solver *s, StepArgs a, unsigned long long t, const LinkNode *list,
               int64_t n_list, bool edge_columns, int64_t *x_lo, int64_t *x_hi)
{
    const Layout &L = s->L;
    cudaStream_t es = s->edge_stream;
    double *fout = a.fout;
    const int64_t oplane = a.fout_plane;
    const int32_t *omap = a.fout_map;
    // offset of population k of node `idx` in fout; a ghost row is contiguous
    // in a compact lattice too (its segments are allocated in index order)
    const int32_t *host_map = omap ? s->mid_map_host.data() : nullptr;
    auto at = [&](int k, int64_t idx) { return k * oplane + lat_off(host_map, idx); };
    const int32_t right_dirs[3] = {1, 5, 8}, left_dirs[3] = {3, 6, 7};
    const bool faces = s->cfg.left_neighbor || s->cfg.right_neighbor;
    // peer-to-peer faces: this step's parity selects the half of the
    // neighbours' receive buffers that the edge kernels store into
    const int64_t half = 3 * L.pitch;                 // doubles per parity
    const int64_t par_off = int64_t(t & 1) * half;
    if (s->p2p) {
        a.face_stride = L.pitch;
        if (s->left_rank >= 0)       // the left neighbour's recv_right
            a.face_lo = reinterpret_cast<double *>(s->peer_left + XBUF_MAILBOX) +
                        2 * half + par_off;
        if (s->right_rank >= 0)      // the right neighbour's recv_left
            a.face_hi = reinterpret_cast<double *>(s->peer_right + XBUF_MAILBOX) +
                        par_off;
    }
    *x_lo = 0;
    *x_hi = L.nx;
    if (edge_columns && faces && L.nx > 2) {
        if (int rc = bulk_timed(s, a, 0, 1, es, false, s->p2p)) return rc;
        if (int rc = bulk_timed(s, a, L.nx - 1, L.nx, es, false, s->p2p)) return rc;
        *x_lo = 1;
        *x_hi = L.nx - 1;
    } else if (edge_columns && faces) {
        if (int rc = bulk_timed(s, a, 0, L.nx, es, true, s->p2p)) return rc;
        *x_lo = *x_hi = 0;
    }
    s->launches += launch_links(a, list, n_list, s->elements_dev, es);
    if (!s->comm) {
        // single rank: the periodic image is this rank's own ghost column
        if (s->cfg.left_neighbor)
            s->launches += launch_face_unpack(
                L, fout, oplane, omap, 0, right_dirs, fout, at(1, L.at(L.nx, 0)),
                at(5, L.at(L.nx, 0)), at(8, L.at(L.nx, 0)), s->mask_left, es);
        if (s->cfg.right_neighbor)
            s->launches += launch_face_unpack(
                L, fout, oplane, omap, L.nx - 1, left_dirs, fout, at(3, L.at(-1, 0)),
                at(6, L.at(-1, 0)), at(7, L.at(-1, 0)), s->mask_right, es);
    } else if (s->p2p) {
        // the stores are done (stream order): publish step t to the neighbours
        // (mailbox word 0 = "data from your left", word 1 = "from your right")
        auto mailbox = [](unsigned char *base, int word) {
            return reinterpret_cast<unsigned long long *>(base) + word;
        };
        s->launches += launch_face_signal(
            s->right_rank >= 0 ? mailbox(s->peer_right, 0) : nullptr,
            s->left_rank >= 0 ? mailbox(s->peer_left, 1) : nullptr, t, es);
        const double *mine = reinterpret_cast<const double *>(s->xbuf + XBUF_MAILBOX);
        if (s->left_rank >= 0)
            s->launches += launch_face_unpack(
                L, fout, oplane, omap, 0, right_dirs, mine + par_off, L.y0, L.pitch + L.y0,
                2 * L.pitch + L.y0, s->mask_left, es, mailbox(s->xbuf, 0), t,
                mailbox(s->xbuf, 2), s->spin_budget);
        if (s->right_rank >= 0)
            s->launches += launch_face_unpack(
                L, fout, oplane, omap, L.nx - 1, left_dirs, mine + 2 * half + par_off, L.y0,
                L.pitch + L.y0, 2 * L.pitch + L.y0, s->mask_right, es,
                mailbox(s->xbuf, 1), t, mailbox(s->xbuf, 2), s->spin_budget);
    } else {
        NCCL_TRY(g_nccl.GroupStart());
        if (s->right_rank >= 0)
            for (int j = 0; j < 3; ++j)
                NCCL_TRY(g_nccl.Send(fout + at(right_dirs[j], L.at(L.nx, 0)),
                                     size_t(L.ny), ncclDouble, s->right_rank,
                                     s->comm, es));
        if (s->left_rank >= 0)
            for (int j = 0; j < 3; ++j)
                NCCL_TRY(g_nccl.Send(fout + at(left_dirs[j], L.at(-1, 0)),
                                     size_t(L.ny), ncclDouble, s->left_rank,
                                     s->comm, es));
        if (s->left_rank >= 0)
            for (int j = 0; j < 3; ++j)
                NCCL_TRY(g_nccl.Recv(s->recv_left + j * L.ny, size_t(L.ny),
                                     ncclDouble, s->left_rank, s->comm, es));
        if (s->right_rank >= 0)
            for (int j = 0; j < 3; ++j)
                NCCL_TRY(g_nccl.Recv(s->recv_right + j * L.ny, size_t(L.ny),
                                     ncclDouble, s->right_rank, s->comm, es));
        NCCL_TRY(g_nccl.GroupEnd());
        if (s->left_rank >= 0)
            s->launches += launch_face_unpack(L, fout, oplane, omap, 0, right_dirs,
                                              s->recv_left, 0, L.ny, 2 * L.ny,
                                              s->mask_left, es);
        if (s->right_rank >= 0)
            s->launches += launch_face_unpack(L, fout, oplane, omap, L.nx - 1, left_dirs,
                                              s->recv_right, 0, L.ny, 2 * L.ny,
                                              s->mask_right, es);
    }
    return PLB_OK;
}

int step_once(plb_solver *s, bool store, bool record)
{
    StepArgs a = step_args(s, s->f[s->cur], s->f[s->cur ^ 1]);
    a.store = store ? 1 : 0;
    if (record && s->n_links > 0) {
        if (!s->exch_dev)
            CUDA_TRY(cudaMalloc(&s->exch_dev, size_t(s->n_links) * 8 * sizeof(double)));
        a.exch = s->exch_dev;
    }

    // Two concurrent chains that write disjoint slots of lattice B.
    //   edge stream (high priority): the O(perimeter) work (edge_chain);
    //   main stream: the bulk kernel on all other columns.
    // They join before the zero_gradient pass, so the small kernels and the
    // exchange latency hide behind the bulk pass.
    cudaStream_t es = s->edge_stream;
    CUDA_TRY(cudaEventRecord(s->ev_edge, s->stream));
    CUDA_TRY(cudaStreamWaitEvent(es, s->ev_edge, 0));
    const unsigned long long t = (unsigned long long)(s->steps_done + 1);
    int64_t x_lo = 0, x_hi = 0;
    if (int rc = edge_chain(s, a, t, s->links_dev, s->n_links, true, &x_lo, &x_hi))
        return rc;
    CUDA_TRY(cudaEventRecord(s->ev_comm, es));
    if (int rc = bulk_timed(s, a, x_lo, x_hi, s->stream, true)) return rc;
    CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_comm, 0));
    run_zero_gradient(s, a, s->stream);
    s->cur ^= 1;
    s->steps_done += 1;
    return PLB_OK;
}

// DEPTH (2 or 3) steps as one pass (k_bulk_fused in plb_kernels.cu).  A = time t,
// B = time t + DEPTH, M1 (M2) = scratch lattices that hold time t + 1 (t + 2)
// on the nodes that are NOT deep enough:
//   edge stream: list pass 1: A -> M1 (-> ... -> B in the last pass), each with
//                its face exchange and, except the last, its zero_gradient pass;
//                list p holds the not-deep-enough fluid nodes plus the deep
//                nodes within distance DEPTH - p of them (they feed the next
//                pass);
//   main stream: all DEPTH steps of every deep node, A -> B;
//   join, zero_gradient on B.
// Slot (n, k) of B has one writer: the owner n - c_k (deep: the fused kernel,
// else the last list pass), n itself (bounce back / element / uncovered
// ghost), or the face delivery.  Every slot of a scratch lattice that the next
// list pass reads (all nine of each of its nodes) is written by the pass
// before: its source is on that pass's (larger) list, the node itself, or the
// face delivery.  The per-step protocol between ranks (stores, mailbox value
// t, delivery) is the one of step_once, so a neighbour may run the same steps
// unfused.
int step_fused(plb_solver *s, int depth)
{
    const Layout &L = s->L;
    for (int m = 0; m < depth - 1; ++m) {
        if (s->f_mid[m]) continue;
        // compact: only the 16-node segments the list passes touch (mid_map)
        const size_t bytes = size_t(Q) * s->mid_plane * sizeof(double);
        if (cudaMalloc(&s->f_mid[m], bytes) != cudaSuccess) {
            cudaGetLastError();
            s->f_mid[m] = nullptr;
            // no room for the scratch lattice: fall back to what fits
            s->fused_depth_ok = m == 0 ? 0 : m + 1;
            return PLB_ERR_NOMEM;
        }
        CUDA_TRY(cudaMemsetAsync(s->f_mid[m], 0, bytes, s->stream));
    }
    double *A = s->f[s->cur], *B = s->f[s->cur ^ 1];
    cudaStream_t es = s->edge_stream;
    CUDA_TRY(cudaEventRecord(s->ev_edge, s->stream));
    CUDA_TRY(cudaStreamWaitEvent(es, s->ev_edge, 0));
    const unsigned long long t1 = (unsigned long long)(s->steps_done + 1);
    int64_t x_lo = 0, x_hi = 0;
    for (int p = 0; p < depth; ++p) {
        const double *fin = p == 0 ? A : s->f_mid[p - 1];
        double *fout = p == depth - 1 ? B : s->f_mid[p];
        const StepArgs pa = step_args(s, fin, fout);
        if (int rc = edge_chain(s, pa, t1 + p, s->lists_dev[depth - 2][p],
                                s->n_lists[depth - 2][p], false, &x_lo, &x_hi))
            return rc;
        if (p < depth - 1) run_zero_gradient(s, pa, es);
    }
    CUDA_TRY(cudaEventRecord(s->ev_comm, es));

    const StepArgs a = step_args(s, A, B);
    const bool prof = s->profile;
    if (prof) {
        if (s->prof_used + 2 > s->prof_events.size())
            for (int i = 0; i < 2; ++i) {
                cudaEvent_t e;
                CUDA_TRY(cudaEventCreate(&e));
                s->prof_events.push_back(e);
            }
        CUDA_TRY(cudaEventRecord(s->prof_events[s->prof_used], s->stream));
    }
    // level 0 reads `depth - 1` rows either side of the rows it delivers and
    // the lattice has one ghost row: columns closer than depth - 2 to the slab
    // edge are left out (they cannot be deep enough anyway)
    // A persistent grid (PLB_FUSED_DYNAMIC) holds every CTA slot until the pass is
    // over, so the later list passes of this group -- and with them the steps
    // the neighbour ranks are waiting for -- would queue up behind it instead
    // of running beside it: between ranks the grid is always the plain one.
    unsigned *work_counter = s->comm ? nullptr : s->work_counter;
    s->launches += launch_bulk_fused(a, s->deep_dev, depth, depth - 2, L.nx - (depth - 2),
                                     s->fused_rows[depth - 2], work_counter,
                                     &s->f_tmap[s->cur], s->stream);
    if (prof) {
        CUDA_TRY(cudaEventRecord(s->prof_events[s->prof_used + 1], s->stream));
        s->prof_used += 2;
    }
    CUDA_TRY(cudaStreamWaitEvent(s->stream, s->ev_comm, 0));
    run_zero_gradient(s, a, s->stream);
    s->cur ^= 1;
    s->steps_done += depth;
    s->groups_done[depth - 2] += 1;
    return PLB_OK;
}

// Steps per pass by default, per collision kernel (PLB_FUSE_DEPTH overrides).
// Measured on the B200 (profiles/r02_fused_sweep_v7_depth4.txt): the
// two-stress-moment MRT kernel is bound by HBM at three steps per pass and
// gains 14 % from a fourth (132.3 against 115.9 GLUPS; fp64 pipe 70 % busy
// then); the reference-ordered BGK kernels are bound by the fp64 pipe at three
// already (110.5 against 108.5), where a fourth costs a scratch lattice, a
// list pass and a face exchange per group for 2 %.
#ifndef PLB_FUSE_DEPTH_MRT
#define PLB_FUSE_DEPTH_MRT 4
#endif
int default_fuse_depth(int kernel_collision)
{
    return kernel_collision == 2 ? PLB_FUSE_DEPTH_MRT : 3;
}

// Steps per pass the solver groups plain steps into right now (1: none).
int fused_depth(const plb_solver *s)
{
    if (s->fuse_mode == 0) return 1;
    const int d = std::min(s->fuse_depth, s->fused_depth_ok);
    return d >= 2 ? d : 1;
}

bool fused_active(const plb_solver *s) { return fused_depth(s) >= 2; }

// n steps; `flags` apply to the last one.  Steps that need neither moments nor
// the link record go fused_depth() at a time (a remainder of two as a pair).
int run_steps(plb_solver *s, int64_t n, int32_t flags)
{
    int64_t i = 0;
    const int64_t plain = flags ? n - 1 : n;
    for (;;) {
        int d = fused_depth(s);
        if (d > plain - i) d = d >= 2 ? int(plain - i) : 1;   // a remainder as a smaller group
        if (d < 2) break;
        const int rc = step_fused(s, d);
        if (rc == PLB_ERR_NOMEM) continue;      // fused_depth() has shrunk
        if (rc) return rc;
        i += d;
    }
    for (; i < n; ++i) {
        const bool last = i == n - 1;
        if (int rc = step_once(s, last && (flags & PLB_STORE_MOMENTS),
                               last && (flags & PLB_RECORD_LINKS)))
            return rc;
    }
    return PLB_OK;
}

// plb_step() is asynchronous; with the fused path a single plain step is held
// back until its partner arrives (or until any other entry point needs the
// state), so that a host loop that issues one step per call -- the
// reference's Solver.run -- still advances several steps per pass.
int flush_pending(plb_solver *s)
{
    const int64_t n = s->pending;
    s->pending = 0;
    return n > 0 ? run_steps(s, n, 0) : PLB_OK;
}

// Peer-to-peer slab faces.  Every rank exports one exchange buffer with CUDA
// IPC; the 64-byte handles travel to both neighbours over the (bootstrap) NCCL
// communicator, each neighbour maps the buffer, and from then on a step moves
// its face populations with plain stores over NVLink from the kernels that
// produce them -- no pack, no NCCL call, no proxy thread on the step path.
// The decision is collective: if any rank cannot map a neighbour, all ranks
// keep the NCCL send/recv faces.
struct Hello {
    cudaIpcMemHandle_t handle;
    int64_t ny, pitch;
};

int setup_p2p(plb_solver *s)
{
    const Layout &L = s->L;
    const size_t bytes = std::max<size_t>(
        XBUF_MAILBOX + size_t(12) * L.pitch * sizeof(double), size_t(2) << 20);
    CUDA_TRY(cudaMalloc(&s->xbuf, bytes));
    CUDA_TRY(cudaMemset(s->xbuf, 0, bytes));

    Hello mine;
    memset(&mine, 0, sizeof mine);
    int ok = cudaIpcGetMemHandle(&mine.handle, s->xbuf) == cudaSuccess;
    cudaGetLastError();
    mine.ny = L.ny;
    mine.pitch = L.pitch;

    // hello[0] = mine, hello[1] = left neighbour's, hello[2] = right neighbour's
    Hello *hello = nullptr;
    CUDA_TRY(cudaMalloc(&hello, 3 * sizeof(Hello)));
    CUDA_TRY(cudaMemset(hello, 0, 3 * sizeof(Hello)));
    CUDA_TRY(cudaMemcpy(hello, &mine, sizeof mine, cudaMemcpyHostToDevice));
    NCCL_TRY(g_nccl.GroupStart());
    if (s->left_rank >= 0) {
        NCCL_TRY(g_nccl.Send(hello, sizeof(Hello), ncclChar, s->left_rank, s->comm, s->stream));
        NCCL_TRY(g_nccl.Recv(hello + 1, sizeof(Hello), ncclChar, s->left_rank, s->comm, s->stream));
    }
    if (s->right_rank >= 0) {
        NCCL_TRY(g_nccl.Send(hello, sizeof(Hello), ncclChar, s->right_rank, s->comm, s->stream));
        NCCL_TRY(g_nccl.Recv(hello + 2, sizeof(Hello), ncclChar, s->right_rank, s->comm, s->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    Hello theirs[3];
    CUDA_TRY(cudaMemcpy(theirs, hello, sizeof theirs, cudaMemcpyDeviceToHost));

    auto map = [&](const Hello &h, unsigned char **out) {
        if (h.ny != L.ny || h.pitch != L.pitch) return false;
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h.handle, cudaIpcMemLazyEnablePeerAccess) !=
            cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        *out = static_cast<unsigned char *>(ptr);
        return true;
    };
    if (ok && s->left_rank >= 0) ok = map(theirs[1], &s->peer_left);
    if (ok && s->right_rank >= 0) {
        if (s->right_rank == s->left_rank) s->peer_right = s->peer_left;   // ring of two
        else ok = map(theirs[2], &s->peer_right);
    }

    // collective decision: min over ranks of "everything mapped"
    int *flag = reinterpret_cast<int *>(hello);
    CUDA_TRY(cudaMemcpy(flag, &ok, sizeof ok, cudaMemcpyHostToDevice));
    NCCL_TRY(g_nccl.AllReduce(flag, flag, 1, ncclInt, ncclMin, s->comm, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(&ok, flag, sizeof ok, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaFree(hello));
    s->p2p = ok != 0;

    double timeout_s = 120.0;
    if (const char *v = getenv("PLB_P2P_TIMEOUT_S")) timeout_s = atof(v);
    int khz = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, s->cfg.device));
    s->spin_budget = (long long)(timeout_s * 1e3 * double(khz));
    return PLB_OK;
}

// Peer-to-peer faces: has a slab-face wait of this rank timed out?  Called by
// every entry point that hands device results to the host, after it has
// synchronised the stream -- a neighbour rank that stalled makes later steps
// run on stale face buffers, so nothing computed since may leave the library
// looking like a result.
int check_faces(plb_solver *s)
{
    if (!s->p2p || !s->xbuf) return PLB_OK;
    unsigned long long status = 0;
    CUDA_TRY(cudaMemcpy(&status, s->xbuf + 2 * sizeof status, sizeof status,
                        cudaMemcpyDeviceToHost));
    if (status)
        return fail(PLB_ERR_NCCL, "slab-face exchange timed out: a neighbour "
                    "rank never published its step (PLB_P2P_TIMEOUT_S)");
    return PLB_OK;
}

void close_p2p(plb_solver *s)
{
    if (s->peer_left) cudaIpcCloseMemHandle(s->peer_left);
    if (s->peer_right && s->peer_right != s->peer_left)
        cudaIpcCloseMemHandle(s->peer_right);
    s->peer_left = s->peer_right = nullptr;
    cudaFree(s->xbuf);
    s->xbuf = nullptr;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char *plb_last_error(void) { return g_error.c_str(); }

int plb_create(const plb_config *c, plb_handle *out)
{
    if (!c || !out) return fail(PLB_ERR_INVALID, "null argument");
    if (c->abi_version != PLB_ABI_VERSION)
        return fail(PLB_ERR_INVALID, "abi_version %d != %d", c->abi_version,
                    PLB_ABI_VERSION);
    if (c->nx < 1 || c->ny < 1)
        return fail(PLB_ERR_INVALID, "nx, ny must be >= 1 (got %lld, %lld)",
                    (long long)c->nx, (long long)c->ny);
    if (c->nx > 2000000000LL || c->ny > 2000000000LL)
        return fail(PLB_ERR_INVALID, "nx, ny must fit int32");
    if (c->collision != PLB_BGK && c->collision != PLB_MRT)
        return fail(PLB_ERR_INVALID, "unknown collision model %d", c->collision);
    if (c->forcing < 0 || c->forcing > 2)
        return fail(PLB_ERR_INVALID, "unknown forcing model %d", c->forcing);

    int n_dev = 0;
    CUDA_TRY(cudaGetDeviceCount(&n_dev));
    if (c->device < 0 || c->device >= n_dev)
        return fail(PLB_ERR_CUDA, "CUDA device %d not available (%d visible)",
                    c->device, n_dev);
    CUDA_TRY(cudaSetDevice(c->device));

    plb_solver *s = new plb_solver();
    s->cfg = *c;
    Layout &L = s->L;
    L.nx = c->nx;
    L.ny = c->ny;
    L.y0 = 16;
    L.pitch = ((L.y0 + L.ny + 1) + 15) / 16 * 16;
    L.plane = (L.nx + 2) * L.pitch;
    KParams &p = s->kp;
    p.L = L;
    p.omega = c->omega;
    p.gx = c->gravity[0];
    p.gy = c->gravity[1];
    p.inv_cs_2 = c->inv_cs_2;
    p.inv_cs_4 = c->inv_cs_4;
    p.eps = c->float_min;
    for (int k = 0; k < Q; ++k) {
        p.w[k] = c->weights[k];
        p.s[k] = c->mrt_rates[k];
    }
    {
        MrtStress &t = p.mrt;
        t.hgx = 0.5 * p.gx;
        t.hgy = 0.5 * p.gy;
        t.k2 = 0.5 * p.inv_cs_2;
        t.k4 = 0.5 * p.inv_cs_4;
        t.k7 = p.w[1] * p.inv_cs_4;
        t.k8 = 4.0 * p.w[5] * p.inv_cs_4;
        t.qa = 0.25 * (1.0 - p.s[7]);
        t.qb = 0.25 * (1.0 - p.s[8]);
        const double cg[4] = {p.gx, p.gy, p.gx + p.gy, p.gx - p.gy};
        for (int j = 0; j < 4; ++j) {
            t.k4cg[j] = t.k4 * cg[j];
            t.k2cg[j] = t.k2 * cg[j];
        }
    }
    {
        // lattice sums for the nine-rate MRT in moment space (MrtMoments)
        static const int M[Q][Q] = {
            {1, 1, 1, 1, 1, 1, 1, 1, 1},      {-4, -1, -1, -1, -1, 2, 2, 2, 2},
            {4, -2, -2, -2, -2, 1, 1, 1, 1},  {0, 1, 0, -1, 0, 1, -1, -1, 1},
            {0, -2, 0, 2, 0, 1, -1, -1, 1},   {0, 0, 1, 0, -1, 1, 1, -1, -1},
            {0, 0, -2, 0, 2, 1, 1, -1, -1},   {0, 1, -1, 1, -1, 0, 0, 0, 0},
            {0, 0, 0, 0, 0, 1, -1, 1, -1}};
        double A[Q], Bx[Q], By[Q], Cxx[Q], Cyy[Q], Cxy[Q], norm2[Q];
        for (int r = 0; r < Q; ++r) {
            A[r] = Bx[r] = By[r] = Cxx[r] = Cyy[r] = Cxy[r] = norm2[r] = 0.0;
            for (int k = 0; k < Q; ++k) {
                const double mw = M[r][k] * p.w[k];
                A[r] += mw;
                Bx[r] += mw * CX[k];
                By[r] += mw * CY[k];
                Cxx[r] += mw * CX[k] * CX[k];
                Cyy[r] += mw * CY[k] * CY[k];
                Cxy[r] += mw * CX[k] * CY[k];
                norm2[r] += double(M[r][k] * M[r][k]);
            }
        }
        MrtMoments &t = p.mrtm;
        for (int r = 0; r < 3; ++r) {
            t.A[r] = A[r];
            t.Qx[r] = 0.5 * p.inv_cs_4 * Cxx[r] - 0.5 * p.inv_cs_2 * A[r];
            t.Qy[r] = 0.5 * p.inv_cs_4 * Cyy[r] - 0.5 * p.inv_cs_2 * A[r];
            t.GA[r] = p.inv_cs_2 * A[r];
            t.Gx[r] = p.inv_cs_4 * Cxx[r];
            t.Gy[r] = p.inv_cs_4 * Cyy[r];
        }
        t.B[0] = p.inv_cs_2 * Bx[3];
        t.B[1] = p.inv_cs_2 * Bx[4];
        t.B[2] = p.inv_cs_2 * By[5];
        t.B[3] = p.inv_cs_2 * By[6];
        t.Q7x = 0.5 * p.inv_cs_4 * Cxx[7];
        t.Q7y = 0.5 * p.inv_cs_4 * Cyy[7];
        t.Q8 = p.inv_cs_4 * Cxy[8];
        for (int r = 0; r < Q; ++r) {
            t.sn[r] = p.s[r] / norm2[r];
            t.hn[r] = (1.0 - 0.5 * p.s[r]) / norm2[r];
        }
    }
    if (const char *v = getenv("PLB_KERNEL"))
        s->variant = (strcmp(v, "scalar") == 0) ? 0 : 1;
    if (const char *v = getenv("PLB_FUSE")) s->fuse_mode = atoi(v);
    if (s->fuse_mode < 0 || s->fuse_mode > 2) s->fuse_mode = PLB_FUSE_DEFAULT;
    s->fuse_depth_from_env = false;
    if (const char *v = getenv("PLB_FUSE_DEPTH")) {
        s->fuse_depth = atoi(v);
        s->fuse_depth_from_env = true;
    }
    if (s->fuse_depth < 2 || s->fuse_depth > MAX_FUSE_DEPTH) {
        s->fuse_depth = 3;
        s->fuse_depth_from_env = false;
    }
    s->kernel_collision = c->collision;
    if (c->collision == PLB_MRT) {
        // S = (1,..,1,s7,s8) as in base/collision_operator.py:159-163 needs
        // only the two stress moments (mrt_reduced_all)
        bool reference_rates = true;
        for (int k = 0; k < 7; ++k) reference_rates &= c->mrt_rates[k] == 1.0;
        const char *g = getenv("PLB_MRT_GENERAL");
        if (reference_rates && !(g && g[0] == '1')) s->kernel_collision = 2;
    }
    if (!s->fuse_depth_from_env) s->fuse_depth = default_fuse_depth(s->kernel_collision);

    auto cleanup = [&](int rc) {
        plb_destroy(s);
        return rc;
    };
#define TRY_OR_CLEAN(expr)                                                    \
    do {                                                                      \
        cudaError_t err__ = (expr);                                           \
        if (err__ != cudaSuccess)                                             \
            return cleanup(fail(err__ == cudaErrorMemoryAllocation            \
                                    ? PLB_ERR_NOMEM : PLB_ERR_CUDA,           \
                                "%s failed: %s", #expr,                       \
                                cudaGetErrorString(err__)));                  \
    } while (0)
    TRY_OR_CLEAN(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        TRY_OR_CLEAN(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        TRY_OR_CLEAN(cudaStreamCreateWithPriority(&s->edge_stream, cudaStreamNonBlocking, hi));
        TRY_OR_CLEAN(cudaEventCreateWithFlags(&s->ev_edge, cudaEventDisableTiming));
        TRY_OR_CLEAN(cudaEventCreateWithFlags(&s->ev_comm, cudaEventDisableTiming));
    }
    for (auto &e : s->events) TRY_OR_CLEAN(cudaEventCreate(&e));
    const size_t plane_bytes = size_t(L.plane) * sizeof(double);
    for (int i = 0; i < 2; ++i) {
        TRY_OR_CLEAN(cudaMalloc(&s->f[i], Q * plane_bytes));
        TRY_OR_CLEAN(cudaMemsetAsync(s->f[i], 0, Q * plane_bytes, s->stream));
        if (fused_needs_tensor_map()) {
            char why[200] = "";
            if (make_lattice_tensor_map(&s->f_tmap[i], s->f[i], L, why, sizeof why)) {
                // the single-step kernels need no descriptor: carry on with them
                fprintf(stderr, "libplb: %s -- several steps per pass switched off\n", why);
                s->tmap_ok = false;
            }
        }
    }
    TRY_OR_CLEAN(cudaMalloc(&s->mom, 3 * plane_bytes));
    TRY_OR_CLEAN(cudaMemsetAsync(s->mom, 0, 3 * plane_bytes, s->stream));
    TRY_OR_CLEAN(cudaMalloc(&s->code, size_t(L.plane)));
    TRY_OR_CLEAN(cudaMemsetAsync(s->code, NODE_GHOST, size_t(L.plane), s->stream));
    TRY_OR_CLEAN(cudaStreamSynchronize(s->stream));
#undef TRY_OR_CLEAN
    s->solid_host.assign(size_t((L.nx + 2) * (L.ny + 2)), 0);
    *out = s;
    return PLB_OK;
}

void plb_destroy(plb_handle s)
{
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->edge_stream) cudaStreamSynchronize(s->edge_stream);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    close_p2p(s);
    if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
    for (int i = 0; i < 2; ++i) cudaFree(s->f[i]);
    for (int m = 0; m < MAX_FUSE_DEPTH - 1; ++m) cudaFree(s->f_mid[m]);
    cudaFree(s->mid_map_dev);
    cudaFree(s->deep_dev);
    cudaFree(s->work_counter);
    for (int d = 0; d < MAX_FUSE_DEPTH - 1; ++d)
        for (int p = 0; p < MAX_FUSE_DEPTH; ++p) cudaFree(s->lists_dev[d][p]);
    cudaFree(s->mom);
    cudaFree(s->mom_old);
    cudaFree(s->res_partials);
    cudaFree(s->res_out);
    cudaFree(s->code);
    cudaFree(s->staging);
    cudaFree(s->flush_buf);
    cudaFree(s->exch_dev);
    cudaFree(s->elements_dev);
    cudaFree(s->links_dev);
    for (auto &z : s->zg_dev) cudaFree(z.first);
    cudaFree(s->mask_left);
    cudaFree(s->mask_right);
    cudaFree(s->recv_left);
    cudaFree(s->recv_right);
    for (auto &e : s->events)
        if (e) cudaEventDestroy(e);
    for (auto &e : s->prof_events) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (s->ev_filled[i]) cudaEventDestroy(s->ev_filled[i]);
        if (s->ev_drained[i]) cudaEventDestroy(s->ev_drained[i]);
    }
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    if (s->ev_edge) cudaEventDestroy(s->ev_edge);
    if (s->ev_comm) cudaEventDestroy(s->ev_comm);
    if (s->edge_stream) cudaStreamDestroy(s->edge_stream);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int plb_add_boundary_element(plb_handle s, int32_t bc_type,
                             const int64_t *boundary_nodes, int64_t n_nodes,
                             const int64_t out_list[3], const int64_t inv_list[3],
                             const int64_t normal[2], const double vector_fluid[2],
                             double scalar_fluid)
{
    if (!s) return fail(PLB_ERR_INVALID, "null handle");
    if (s->finalized)
        return fail(PLB_ERR_STATE, "geometry already finalized");
    if (bc_type < PLB_BC_BOUNCE_BACK || bc_type > PLB_BC_ZERO_GRADIENT)
        return fail(PLB_ERR_INVALID, "unknown boundary type %d", bc_type);
    if (n_nodes < 0 || (n_nodes > 0 && !boundary_nodes))
        return fail(PLB_ERR_INVALID, "bad boundary node list");
    if ((int)s->elements.size() >= MAX_ELEMENTS)
        return fail(PLB_ERR_INVALID, "more than %d boundary elements", MAX_ELEMENTS);
    const int64_t size = (s->L.nx + 2) * (s->L.ny + 2);
    ElementHost e;
    e.type = bc_type;
    e.nodes.assign(boundary_nodes, boundary_nodes + n_nodes);
    for (int64_t n : e.nodes)
        if (n < 0 || n >= size)
            return fail(PLB_ERR_INVALID, "boundary node %lld out of range",
                        (long long)n);
    for (int j = 0; j < 3; ++j) {
        if (out_list[j] < 1 || out_list[j] >= Q || inv_list[j] != INV[out_list[j]])
            return fail(PLB_ERR_INVALID,
                        "out_list/inv_list are not opposite D2Q9 directions");
        e.out[j] = out_list[j];
        e.inv[j] = inv_list[j];
    }
    e.normal[0] = normal[0];
    e.normal[1] = normal[1];
    e.vector[0] = vector_fluid ? vector_fluid[0] : 0.0;
    e.vector[1] = vector_fluid ? vector_fluid[1] : 0.0;
    e.scalar = scalar_fluid;
    s->elements.push_back(std::move(e));
    return PLB_OK;
}

int plb_finalize_geometry(plb_handle s)
{
    if (!s) return fail(PLB_ERR_INVALID, "null handle");
    if (s->finalized) return fail(PLB_ERR_STATE, "geometry already finalized");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    const Layout &L = s->L;
    const int64_t nx = L.nx, ny = L.ny, nyp = ny + 2;
    // PLB_SETUP_TIMING=1: wall clock of the phases below on stderr
    const bool timing = getenv("PLB_SETUP_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[plb setup] %-28s %8.1f ms\n", what,
                std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };

    // links owned by boundary elements; later elements win, zero_gradient
    // elements are a final pass (see oracle_set_boundary)
    std::unordered_map<int64_t, int32_t> element_of;
    std::unordered_set<int64_t> zg_cover;
    for (size_t e = 0; e < s->elements.size(); ++e) {
        const ElementHost &el = s->elements[e];
        if (el.type == PLB_BC_PERIODIC) continue;
        for (int64_t ind : el.nodes)
            for (int j = 0; j < 3; ++j) {
                if (el.type == PLB_BC_ZERO_GRADIENT) zg_cover.insert(ind * Q + el.out[j]);
                else element_of[ind * Q + el.out[j]] = int32_t(e);
            }
    }
    Classifier cls{s, nx, ny, nyp, s->solid_host.data(), &element_of, &zg_cover};
    lap("element maps");

    // rows (reference x) that contain any solid node, ghost ring included
    std::vector<uint8_t> row_has_solid(size_t(nx + 2), 0);
    for (int64_t r = 0; r < nx + 2; ++r) {
        const uint8_t *row = s->solid_host.data() + r * nyp;
        row_has_solid[r] = memchr(row, 1, size_t(nyp)) != nullptr;
        if (!row_has_solid[r])
            for (int64_t j = 0; j < nyp; ++j)
                if (row[j]) { row_has_solid[r] = 1; break; }
    }

    std::vector<uint8_t> code(size_t(L.plane), NODE_GHOST);
    std::vector<LinkNode> link_nodes;
    int64_t n_bulk = 0, n_solid = 0;
    // row_class[r] (padded row r = x + 1): rows whose interior columns
    // 1 .. ny - 2 are all BULK get the class (bulk-ness of y = 0, of y = ny - 1)
    // in 0 .. 3, every other row -1 -- rows of one class have identical
    // BULK / non-BULK patterns, which is all the deep flags depend on
    std::vector<int8_t> row_class(size_t(nx + 2), -1);
    auto classify = [&](int64_t x, int64_t y, uint8_t *crow) {
        if (cls.is_solid(x, y)) {
            crow[y] = NODE_SOLID;
            ++n_solid;
            return;
        }
        const uint64_t lk = cls.links(x, y);
        if (lk == 0) {
            crow[y] = NODE_BULK;
            ++n_bulk;
            if (x == 0) ++s->n_bulk_edge[0];
            if (x == nx - 1 && nx > 1) ++s->n_bulk_edge[1];
        } else {
            crow[y] = NODE_LINK;
            link_nodes.push_back(LinkNode{int32_t(x), int32_t(y), lk});
        }
    };
    for (int64_t x = 0; x < nx; ++x) {
        uint8_t *crow = code.data() + L.at(x, 0);
        const bool near_solid =
            row_has_solid[x] || row_has_solid[x + 1] || row_has_solid[x + 2];
        const bool edge_row = x == 0 || x == nx - 1;
        if (!near_solid && !edge_row && ny >= 2) {
            // every interior column pushes to plain fluid neighbours
            classify(x, 0, crow);
            memset(crow + 1, NODE_BULK, size_t(ny - 2));
            n_bulk += ny - 2;
            classify(x, ny - 1, crow);
            row_class[size_t(x + 1)] = int8_t((crow[0] == NODE_BULK ? 1 : 0) |
                                              (crow[ny - 1] == NODE_BULK ? 2 : 0));
            continue;
        }
        for (int64_t y = 0; y < ny; ++y) classify(x, y, crow);
    }
    s->n_bulk = n_bulk;
    s->n_solid = n_solid;
    lap("node codes");

    // slab-face acceptance masks: slot (node, k) is fed from across the face
    // iff the node is fluid and its own link in direction inv(k) is a push
    std::vector<uint8_t> mask_l(size_t(ny), 0), mask_r(size_t(ny), 0);
    const int right_dirs[3] = {1, 5, 8}, left_dirs[3] = {3, 6, 7};
    for (int64_t y = 0; y < ny; ++y) {
        if (s->cfg.left_neighbor && !cls.is_solid(0, y)) {
            const uint64_t lk = cls.links(0, y);
            for (int j = 0; j < 3; ++j) {
                const int c = int((lk >> (8 * (INV[right_dirs[j]] - 1))) & 0xff);
                if (c == LINK_PUSH || c == LINK_WRAP) mask_l[y] |= uint8_t(1 << j);
            }
        }
        if (s->cfg.right_neighbor && !cls.is_solid(nx - 1, y)) {
            const uint64_t lk = cls.links(nx - 1, y);
            for (int j = 0; j < 3; ++j) {
                const int c = int((lk >> (8 * (INV[left_dirs[j]] - 1))) & 0xff);
                if (c == LINK_PUSH || c == LINK_WRAP) mask_r[y] |= uint8_t(1 << j);
            }
        }
    }

    // Several steps per pass: deep[n] = 1 if n and its eight neighbours are bulk
    // nodes, 2 if all 24 nodes within distance two are (capped there).  A
    // depth-d group advances the nodes with deep >= d - 1 in k_bulk_fused and
    // everything else in d list passes: the last list holds the fluid nodes
    // that are not deep enough, each earlier one adds the deep nodes that
    // touch the list after it (step_fused).  Whole-plane work is limited to
    // two separable 3 x 3 erosions and one scan for bytes != 2, all of which
    // vectorise; the lists are built from the O(perimeter) candidates.
    if (s->fuse_mode != 0) {
        const int64_t P = L.pitch, N = L.plane;
        std::vector<uint8_t> deep(size_t(N), 0);
        {
            // bad1 = a non-bulk code (BULK = 0) in the 3 x 3 neighbourhood,
            // bad2 = a bad1 in the 3 x 3 neighbourhood, deep = !bad1 + !bad2:
            // two separable 3 x 3 dilations, streamed row by row through
            // cache-sized row buffers (no whole-plane temporaries).  Five
            // consecutive rows of one row_class have one and the same deep
            // row, which is then copied instead of recomputed -- all but
            // O(1) rows of a lattice without obstacles.
            const int64_t R = nx + 2;
            const uint8_t *c = code.data();
            auto hor = [P](const uint8_t *in, uint8_t *out) {
                // out[i] = in[i-1] | in[i] | in[i+1]; the row ends are
                // alignment padding (GHOST, non-bulk) and stay "bad"
                out[0] = out[P - 1] = 1;
                for (int64_t i = 1; i + 1 < P; ++i) out[i] = in[i - 1] | in[i] | in[i + 1];
            };
            // ring buffers over padded rows: h1[r] (codes dilated along y),
            // b1[r] (bad1), h2[r] (bad1 dilated along y)
            std::vector<uint8_t> h1buf(size_t(3 * P)), b1buf(size_t(3 * P)), h2buf(size_t(3 * P));
            auto h1 = [&](int64_t r) { return h1buf.data() + (r % 3) * P; };
            auto b1 = [&](int64_t r) { return b1buf.data() + (r % 3) * P; };
            auto h2 = [&](int64_t r) { return h2buf.data() + (r % 3) * P; };
            // b1[r] needs h1[r-1 .. r+1], deep[r] needs b1[r] and h2[r-1 .. r+1];
            // the ghost rows 0 and R - 1 are bad by definition
            auto make_h1 = [&](int64_t r) { hor(c + r * P, h1(r)); };
            std::vector<uint8_t> tmpl[4];
            int64_t h1_next = 0;       // next padded row whose h1 is to be made
            int64_t b1_next = 0;       // next row whose b1 / h2 is to be made
            auto advance_b1 = [&](int64_t upto) {
                // make b1 / h2 for rows b1_next .. upto (inclusive)
                for (; b1_next <= upto; ++b1_next) {
                    const int64_t r = b1_next;
                    for (; h1_next <= r + 1 && h1_next < R; ++h1_next) make_h1(h1_next);
                    uint8_t *o = b1(r);
                    if (r == 0 || r == R - 1) {
                        memset(o, 1, size_t(P));
                    } else {
                        const uint8_t *u = h1(r - 1), *m = h1(r), *d = h1(r + 1);
                        for (int64_t i = 0; i < P; ++i) o[i] = (u[i] | m[i] | d[i]) != 0;
                    }
                    hor(o, h2(r));
                }
            };
            for (int64_t r = 1; r + 1 < R; ++r) {
                uint8_t *out = deep.data() + r * P;
                const int8_t k = row_class[size_t(r)];
                const bool uniform = k >= 0 && r >= 3 && r + 3 < R &&
                    row_class[size_t(r - 2)] == k && row_class[size_t(r - 1)] == k &&
                    row_class[size_t(r + 1)] == k && row_class[size_t(r + 2)] == k;
                if (uniform && !tmpl[k].empty()) {
                    memcpy(out, tmpl[k].data(), size_t(P));
                    continue;
                }
                // the streaming state may lag behind after copied rows
                if (b1_next < r - 1) {
                    b1_next = r - 1;
                    h1_next = r - 2;
                }
                advance_b1(r + 1);
                const uint8_t *u = h2(r - 1), *m = h2(r), *d = h2(r + 1), *bad = b1(r);
                for (int64_t i = 0; i < P; ++i)
                    out[i] = uint8_t(!bad[i]) + uint8_t((u[i] | m[i] | d[i]) == 0);
                if (uniform) tmpl[k].assign(out, out + P);
            }
        }
        // One more erosion per further step per pass: a node whose nine
        // neighbours all carry the cap so far is one ring deeper.  In place,
        // row by row: hq[r][i] = rows r's nodes i-1, i, i+1 all at the cap, made
        // from the still unmodified row before row r - 1 is raised.
        uint8_t cap = 2;
        for (; cap < uint8_t(s->fuse_depth - 1); ++cap) {
            const int64_t R = nx + 2;
            std::vector<uint8_t> hq(size_t(3 * P), 0);
            auto row = [&](int64_t r) { return hq.data() + (r % 3) * P; };
            auto make = [&](int64_t r) {
                const uint8_t *in = deep.data() + r * P;
                uint8_t *o = row(r);
                o[0] = o[P - 1] = 0;
                for (int64_t i = 1; i + 1 < P; ++i)
                    o[i] = uint8_t((in[i - 1] == cap) & (in[i] == cap) & (in[i + 1] == cap));
            };
            make(0);
            make(1);
            for (int64_t r = 1; r + 1 < R; ++r) {
                make(r + 1);
                const uint8_t *u = row(r - 1), *m = row(r), *d = row(r + 1);
                uint8_t *out = deep.data() + r * P;
                for (int64_t i = 0; i < P; ++i) out[i] = uint8_t(out[i] + (u[i] & m[i] & d[i]));
            }
        }
        lap("deep flags");
        // candidates: interior nodes that are not fully deep (a few rings)
        std::vector<int64_t> candidates;
        const uint64_t all_cap = 0x0101010101010101ull * cap;
        for (int64_t x = 0; x < nx; ++x) {
            const uint8_t *d = deep.data() + L.at(x, 0);
            int64_t y = 0;
            while (y < ny) {
                uint64_t w;
                if (y + 8 <= ny && (memcpy(&w, d + y, 8), w == all_cap)) {
                    y += 8;
                    continue;
                }
                if (d[y] != cap) candidates.push_back(L.at(x, y));
                ++y;
            }
        }
        lap("candidates");
        const int64_t n_fluid = n_bulk + int64_t(link_nodes.size());
        std::vector<uint8_t> on(size_t(N), 0);
        // 16-node segments of the plane that a list pass reads or writes in a
        // scratch lattice: the listed nodes and their eight neighbours
        std::vector<uint8_t> seg_used(size_t(N / 16), 0);
        const int64_t nb8[8] = {-1, 1, -P, -P - 1, -P + 1, P, P - 1, P + 1};
        for (int depth = 2; depth <= s->fuse_depth; ++depth) {
            const uint8_t need = uint8_t(depth - 1);
            // last list: fluid nodes that are not deep enough (sorted by index)
            std::vector<int64_t> cur;
            for (int64_t idx : candidates) {
                const uint8_t c = code[size_t(idx)];
                if ((c == NODE_BULK || c == NODE_LINK) && deep[size_t(idx)] < need) {
                    on[size_t(idx)] = 1;
                    cur.push_back(idx);
                }
            }
            const int64_t n_deep = n_fluid - int64_t(cur.size());
            std::vector<std::vector<int64_t>> lists;
            lists.resize(size_t(depth));
            lists[size_t(depth - 1)] = cur;
            for (int p = depth - 2; p >= 0; --p) {
                // the list before: add the deep nodes next to this one (bulk
                // nodes with bulk neighbours, so every neighbour index is valid)
                std::vector<int64_t> add;
                for (int64_t idx : cur)
                    for (int64_t off : nb8) {
                        const int64_t n = idx + off;
                        if (!on[size_t(n)] && deep[size_t(n)] >= need) {
                            on[size_t(n)] = 1;
                            add.push_back(n);
                        }
                    }
                std::sort(add.begin(), add.end());
                std::vector<int64_t> merged(cur.size() + add.size());
                std::merge(cur.begin(), cur.end(), add.begin(), add.end(), merged.begin());
                cur.swap(merged);
                lists[size_t(p)] = cur;
            }
            for (int64_t idx : lists[0]) on[size_t(idx)] = 0;   // clean for the next depth
            s->n_deep[depth - 2] = n_deep;
            const bool ok = n_deep > 0 && s->tmap_ok &&
                            (s->fuse_mode == 2 ||
                             (int64_t(lists[0].size()) * 8 <= n_fluid && n_fluid >= 4096));
            if (!ok) continue;
            s->fused_depth_ok = depth;
            for (int64_t idx : lists[0])
                for (int64_t dx = -P; dx <= P; dx += P) {
                    seg_used[size_t((idx + dx - 1) >> 4)] = 1;
                    seg_used[size_t((idx + dx + 1) >> 4)] = 1;
                }
            for (int p = 0; p < depth; ++p) {
                // index -> record; link nodes (all on every list) keep their codes
                std::vector<LinkNode> l;
                l.reserve(lists[size_t(p)].size());
                size_t next_link = 0;
                for (int64_t idx : lists[size_t(p)]) {
                    const int64_t x = idx / P - 1, y = idx % P - L.y0;
                    if (code[size_t(idx)] == NODE_LINK) {
                        while (next_link < link_nodes.size() &&
                               (link_nodes[next_link].x != x || link_nodes[next_link].y != y))
                            ++next_link;
                        if (next_link == link_nodes.size())
                            return fail(PLB_ERR_STATE, "internal: link node (%lld, %lld) "
                                        "missing from the link list", (long long)x, (long long)y);
                        l.push_back(link_nodes[next_link++]);
                    } else {
                        l.push_back(LinkNode{int32_t(x), int32_t(y), 0});
                    }
                }
                s->n_lists[depth - 2][p] = int64_t(l.size());
                CUDA_TRY(cudaMalloc(&s->lists_dev[depth - 2][p],
                                    std::max<size_t>(1, l.size()) * sizeof(LinkNode)));
                CUDA_TRY(cudaMemcpy(s->lists_dev[depth - 2][p], l.data(),
                                    l.size() * sizeof(LinkNode), cudaMemcpyHostToDevice));
            }
        }
        lap("lists");
        if (s->fused_depth_ok >= 2) {
            // compact scratch lattices: the two ghost rows (send staging of the
            // slab faces, contiguous) and every used segment, numbered in index
            // order from 1; segment 0 is a spare that absorbs the rest
            const int64_t segs_per_row = P / 16;
            for (int64_t i = 0; i < segs_per_row; ++i)
                seg_used[size_t(i)] = seg_used[size_t((nx + 1) * segs_per_row + i)] = 1;
            s->mid_map_host.assign(seg_used.size(), 0);
            int32_t n_seg = 1;
            for (size_t i = 0; i < seg_used.size(); ++i)
                if (seg_used[i]) s->mid_map_host[i] = n_seg++;
            s->mid_plane = int64_t(n_seg) * 16;
            CUDA_TRY(cudaMalloc(&s->mid_map_dev, s->mid_map_host.size() * sizeof(int32_t)));
            CUDA_TRY(cudaMemcpy(s->mid_map_dev, s->mid_map_host.data(),
                                s->mid_map_host.size() * sizeof(int32_t),
                                cudaMemcpyHostToDevice));
            lap("scratch map");
            if (const char *v = getenv("PLB_FUSED_DYNAMIC"))
                if (atoi(v) != 0) CUDA_TRY(cudaMalloc(&s->work_counter, 256));
            CUDA_TRY(cudaMalloc(&s->deep_dev, deep.size()));
            CUDA_TRY(cudaMemcpy(s->deep_dev, deep.data(), deep.size(), cudaMemcpyHostToDevice));
            if (const char *v = getenv("PLB_FUSED_ROWS")) {
                s->fused_rows[0] = s->fused_rows[1] = s->fused_rows[2] = std::max(1, atoi(v));
            } else {
                // Chunk height measured on B200 (profiles/): two steps per pass:
                // 32 rows beat 64 / 128 / 512 although 2 of 34 row loads are
                // then redundant (more warps in their prologue at any time =
                // more loads in flight); three steps per pass: 64 rows (4 of 68
                // redundant) beat 32 and tie with 128; four steps per pass: 64
                // rows (132.3 GLUPS) against 127.4 / 131.3 / 132.1 / 130.7 with
                // 32 / 48 / 96 / 128.  Small lattices: >= 4 waves of work items.
                const int64_t want = nx * fused_strips(L, 2) / (148 * 16 * 4);
                s->fused_rows[0] = int32_t(std::min<int64_t>(32, std::max<int64_t>(8, want)));
                s->fused_rows[1] = int32_t(std::min<int64_t>(64, std::max<int64_t>(8, want)));
                s->fused_rows[2] = int32_t(std::min<int64_t>(64, std::max<int64_t>(8, want)));
            }
        }
    }

    lap("deep upload");
    // device copies
    CUDA_TRY(cudaMemcpy(s->code, code.data(), code.size(), cudaMemcpyHostToDevice));
    s->n_links = int64_t(link_nodes.size());
    s->link_inds.resize(link_nodes.size());
    for (size_t i = 0; i < link_nodes.size(); ++i)
        s->link_inds[i] = (int64_t(link_nodes[i].x) + 1) * nyp + link_nodes[i].y + 1;
    if (s->n_links) {
        CUDA_TRY(cudaMalloc(&s->links_dev, link_nodes.size() * sizeof(LinkNode)));
        CUDA_TRY(cudaMemcpy(s->links_dev, link_nodes.data(),
                            link_nodes.size() * sizeof(LinkNode),
                            cudaMemcpyHostToDevice));
    }
    std::vector<ElementDev> edev(std::max<size_t>(1, s->elements.size()));
    for (size_t e = 0; e < s->elements.size(); ++e) {
        const ElementHost &el = s->elements[e];
        edev[e] = ElementDev{el.type, int32_t(el.normal[0]), int32_t(el.normal[1]),
                             0, el.vector[0], el.vector[1], el.scalar};
    }
    CUDA_TRY(cudaMalloc(&s->elements_dev, edev.size() * sizeof(ElementDev)));
    CUDA_TRY(cudaMemcpy(s->elements_dev, edev.data(), edev.size() * sizeof(ElementDev),
                        cudaMemcpyHostToDevice));
    for (const ElementHost &el : s->elements) {
        if (el.type != PLB_BC_ZERO_GRADIENT) continue;
        std::vector<ZgLink> zl;
        for (int64_t ind : el.nodes) {
            if (s->solid_host[size_t(ind)]) continue;
            const int64_t x = ind / nyp - 1, y = ind % nyp - 1;
            if (x < 0 || x >= nx || y < 0 || y >= ny) continue;
            for (int j = 0; j < 3; ++j)
                zl.push_back(ZgLink{L.at(x, y),
                                    L.at(x + el.normal[0], y + el.normal[1]),
                                    int32_t(el.inv[j]), 0});
        }
        ZgLink *dev = nullptr;
        if (!zl.empty()) {
            CUDA_TRY(cudaMalloc(&dev, zl.size() * sizeof(ZgLink)));
            CUDA_TRY(cudaMemcpy(dev, zl.data(), zl.size() * sizeof(ZgLink),
                                cudaMemcpyHostToDevice));
        }
        s->zg_dev.emplace_back(dev, int64_t(zl.size()));
    }
    CUDA_TRY(cudaMalloc(&s->mask_left, size_t(ny)));
    CUDA_TRY(cudaMalloc(&s->mask_right, size_t(ny)));
    CUDA_TRY(cudaMemcpy(s->mask_left, mask_l.data(), size_t(ny), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(s->mask_right, mask_r.data(), size_t(ny), cudaMemcpyHostToDevice));
    lap("device copies");
    s->finalized = true;
    return PLB_OK;
}

int plb_upload(plb_handle s, int32_t field, const void *host, size_t bytes)
{
    if (!s || !host) return fail(PLB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    const Layout &L = s->L;
    const size_t size = size_t((L.nx + 2) * (L.ny + 2));
    switch (field) {
    case PLB_SOLID:
        if (bytes != size) return fail(PLB_ERR_INVALID, "PLB_SOLID expects %zu bytes", size);
        if (s->finalized) return fail(PLB_ERR_STATE, "geometry already finalized");
        memcpy(s->solid_host.data(), host, size);
        return PLB_OK;
    case PLB_DENSITY:
        if (bytes != size * 8) return fail(PLB_ERR_INVALID, "PLB_DENSITY expects %zu bytes", size * 8);
        return upload_padded(s, static_cast<const double *>(host), 1, s->rho(), L.plane);
    case PLB_VELOCITY:
        if (bytes != size * 16) return fail(PLB_ERR_INVALID, "PLB_VELOCITY expects %zu bytes", size * 16);
        return upload_padded(s, static_cast<const double *>(host), 2, s->ux(), L.plane);
    case PLB_POP:
        if (bytes != size * 72) return fail(PLB_ERR_INVALID, "PLB_POP expects %zu bytes", size * 72);
        return upload_padded(s, static_cast<const double *>(host), Q, s->f[s->cur], L.plane);
    default:
        return fail(PLB_ERR_INVALID, "field %d cannot be uploaded", field);
    }
}

int plb_fill(plb_handle s, int32_t field, const double *value)
{
    if (!s || !value) return fail(PLB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    switch (field) {
    case PLB_DENSITY:
        s->launches += launch_fill_inner(s->L, s->rho(), value[0], s->stream);
        break;
    case PLB_VELOCITY:
        s->launches += launch_fill_inner(s->L, s->ux(), value[0], s->stream);
        s->launches += launch_fill_inner(s->L, s->uy(), value[1], s->stream);
        break;
    default:
        return fail(PLB_ERR_INVALID, "plb_fill: field %d is not PLB_DENSITY / "
                    "PLB_VELOCITY", field);
    }
    CUDA_TRY(cudaGetLastError());
    return PLB_OK;
}

int plb_download(plb_handle s, int32_t field, void *host, size_t bytes)
{
    if (!s || !host) return fail(PLB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    const Layout &L = s->L;
    const size_t size = size_t((L.nx + 2) * (L.ny + 2));
    const size_t inner = size_t(L.nx * L.ny);
    double *out = static_cast<double *>(host);
    auto checked = [&](int rc) { return rc ? rc : check_faces(s); };
    switch (field) {
    case PLB_SOLID:
        if (bytes != size) return fail(PLB_ERR_INVALID, "PLB_SOLID expects %zu bytes", size);
        memcpy(host, s->solid_host.data(), size);
        return PLB_OK;
    case PLB_DENSITY:
        if (bytes != size * 8) return fail(PLB_ERR_INVALID, "PLB_DENSITY expects %zu bytes", size * 8);
        return checked(download_padded(s, out, 1, s->rho(), L.plane, 0));
    case PLB_VELOCITY:
        if (bytes != size * 16) return fail(PLB_ERR_INVALID, "PLB_VELOCITY expects %zu bytes", size * 16);
        return checked(download_padded(s, out, 2, s->ux(), L.plane, 0));
    case PLB_POP:
        if (bytes != size * 72) return fail(PLB_ERR_INVALID, "PLB_POP expects %zu bytes", size * 72);
        return checked(download_padded(s, out, Q, s->f[s->cur], L.plane, 1));
    case PLB_DENSITY_INNER:
        if (bytes != inner * 8) return fail(PLB_ERR_INVALID, "PLB_DENSITY_INNER expects %zu bytes", inner * 8);
        return checked(download_inner(s, out, 1, s->rho(), L.plane));
    case PLB_VELOCITY_INNER:
        if (bytes != inner * 16) return fail(PLB_ERR_INVALID, "PLB_VELOCITY_INNER expects %zu bytes", inner * 16);
        return checked(download_inner(s, out, 2, s->ux(), L.plane));
    default:
        return fail(PLB_ERR_INVALID, "field %d cannot be downloaded", field);
    }
}

int plb_initialize_pop(plb_handle s)
{
    if (!s) return fail(PLB_ERR_INVALID, "null handle");
    if (!s->finalized) return fail(PLB_ERR_STATE, "plb_finalize_geometry not called");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    s->launches += launch_init_pop(s->kp, s->f[s->cur], s->code, s->rho(), s->ux(),
                                   s->uy(), s->stream);
    CUDA_TRY(cudaMemsetAsync(s->f[s->cur ^ 1], 0, size_t(Q) * s->L.plane * sizeof(double),
                             s->stream));
    CUDA_TRY(cudaGetLastError());
    return PLB_OK;
}

int plb_step(plb_handle s, int64_t n_steps, int32_t flags)
{
    if (!s) return fail(PLB_ERR_INVALID, "null handle");
    if (!s->finalized) return fail(PLB_ERR_STATE, "plb_finalize_geometry not called");
    if (n_steps < 0) return fail(PLB_ERR_INVALID, "n_steps < 0");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    flags &= PLB_STORE_MOMENTS | PLB_RECORD_LINKS;
    int64_t n = s->pending + n_steps;
    s->pending = 0;
    if (fused_active(s) && flags == 0) {
        // plain steps that do not fill a group wait for the rest of it
        // (flush_pending)
        s->pending = n % fused_depth(s);
        n -= s->pending;
    }
    if (int rc = run_steps(s, n, flags)) return rc;
    CUDA_TRY(cudaGetLastError());
    return PLB_OK;
}

int plb_fused_info(plb_handle s, int64_t out[10])
{
    if (!s || !out) return fail(PLB_ERR_INVALID, "null argument");
    const int d = fused_depth(s);
    out[0] = d >= 2 ? d : 0;
    out[1] = s->n_deep[0];
    out[2] = s->n_deep[1];
    out[3] = d >= 2 ? s->n_lists[d - 2][0] : 0;
    out[4] = s->groups_done[0];
    out[5] = s->fused_rows[d >= 2 ? d - 2 : 0];
    out[6] = fused_strips(s->L, d >= 2 ? d : 2);
    out[7] = s->groups_done[1];
    out[8] = s->n_deep[2];
    out[9] = s->groups_done[2];
    return PLB_OK;
}

int plb_memory_info(plb_handle s, int64_t out[6])
{
    if (!s || !out) return fail(PLB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    const int64_t plane_bytes = s->L.plane * int64_t(sizeof(double));
    out[0] = 2 * Q * plane_bytes;
    out[1] = (s->mom ? 3 : 0) * plane_bytes + (s->mom_old ? 3 : 0) * plane_bytes;
    out[2] = 0;
    for (int m = 0; m < MAX_FUSE_DEPTH - 1; ++m)
        if (s->f_mid[m]) out[2] += Q * s->mid_plane * int64_t(sizeof(double));
    if (s->mid_map_dev) out[2] += int64_t(s->mid_map_host.size() * sizeof(int32_t));
    int64_t lists = s->n_links * int64_t(sizeof(LinkNode));
    for (int d = 0; d < MAX_FUSE_DEPTH - 1; ++d)
        for (int p = 0; p < MAX_FUSE_DEPTH; ++p)
            lists += s->n_lists[d][p] * int64_t(sizeof(LinkNode));
    out[3] = s->L.plane * (s->deep_dev ? 2 : 1) + lists + 2 * int64_t(s->staging_bytes) +
             (s->exch_dev ? s->n_links * 8 * int64_t(sizeof(double)) : 0);
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    out[4] = int64_t(free_b);
    out[5] = int64_t(total_b);
    return PLB_OK;
}

int plb_link_nodes(plb_handle s, int64_t *padded_index, int64_t capacity,
                   int64_t *n_links)
{
    if (!s || !n_links) return fail(PLB_ERR_INVALID, "null argument");
    if (!s->finalized) return fail(PLB_ERR_STATE, "plb_finalize_geometry not called");
    *n_links = s->n_links;
    if (padded_index) {
        if (capacity < s->n_links)
            return fail(PLB_ERR_INVALID, "capacity %lld < %lld link nodes",
                        (long long)capacity, (long long)s->n_links);
        memcpy(padded_index, s->link_inds.data(), size_t(s->n_links) * sizeof(int64_t));
    }
    return PLB_OK;
}

int plb_download_link_exchange(plb_handle s, double *out, int64_t n_values)
{
    if (!s || !out) return fail(PLB_ERR_INVALID, "null argument");
    if (n_values != s->n_links * 8)
        return fail(PLB_ERR_INVALID, "expected %lld values", (long long)(s->n_links * 8));
    if (!s->exch_dev)
        return fail(PLB_ERR_STATE, "no step was run with PLB_RECORD_LINKS");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, s->exch_dev, size_t(n_values) * sizeof(double),
                             cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return check_faces(s);
}

int plb_sync(plb_handle s)
{
    if (!s) return fail(PLB_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    CUDA_TRY(cudaStreamSynchronize(s->edge_stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return check_faces(s);
}

int plb_residue_sums(plb_handle s, double out[6])
{
    if (!s || !out) return fail(PLB_ERR_INVALID, "null argument");
    if (!s->finalized) return fail(PLB_ERR_STATE, "plb_finalize_geometry not called");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    const int n_blocks = 148 * 4;
    const size_t plane_bytes = size_t(s->L.plane) * sizeof(double);
    if (!s->mom_old) {
        CUDA_TRY(cudaMalloc(&s->mom_old, 3 * plane_bytes));
        CUDA_TRY(cudaMemsetAsync(s->mom_old, 0, 3 * plane_bytes, s->stream));
        CUDA_TRY(cudaMalloc(&s->res_partials, size_t(n_blocks) * 6 * sizeof(double)));
        CUDA_TRY(cudaMalloc(&s->res_out, 6 * sizeof(double)));
    }
    s->launches += launch_residue(s->L, s->code, s->rho(), s->ux(), s->uy(), s->mom_old,
                                  s->mom_old + s->L.plane, s->mom_old + 2 * s->L.plane,
                                  s->res_partials, n_blocks, s->res_out, s->stream);
    CUDA_TRY(cudaMemcpyAsync(out, s->res_out, 6 * sizeof(double), cudaMemcpyDeviceToHost,
                             s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return check_faces(s);
}

int plb_comm_unique_id(void *id128)
{
    if (!id128) return fail(PLB_ERR_INVALID, "null argument");
    if (int rc = load_nccl()) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return PLB_OK;
}

int plb_comm_init(plb_handle s, const void *id128, int32_t rank, int32_t n_ranks,
                  int32_t left_rank, int32_t right_rank)
{
    if (!s || !id128) return fail(PLB_ERR_INVALID, "null argument");
    if (s->comm) return fail(PLB_ERR_STATE, "communicator already initialised");
    if (n_ranks < 2 || rank < 0 || rank >= n_ranks)
        return fail(PLB_ERR_INVALID, "bad rank %d of %d", rank, n_ranks);
    if (left_rank >= n_ranks || right_rank >= n_ranks || left_rank == rank ||
        right_rank == rank)
        return fail(PLB_ERR_INVALID, "bad neighbour ranks %d, %d", left_rank, right_rank);
    if ((left_rank >= 0) != (s->cfg.left_neighbor != 0) ||
        (right_rank >= 0) != (s->cfg.right_neighbor != 0))
        return fail(PLB_ERR_INVALID,
                    "neighbour ranks disagree with plb_config.left/right_neighbor");
    if (int rc = load_nccl()) return rc;
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NCCL_TRY(g_nccl.CommInitRank(&s->comm, n_ranks, id, rank));
    s->rank = rank;
    s->n_ranks = n_ranks;
    s->left_rank = left_rank;
    s->right_rank = right_rank;
    CUDA_TRY(cudaMalloc(&s->recv_left, size_t(3 * s->L.ny) * sizeof(double)));
    CUDA_TRY(cudaMalloc(&s->recv_right, size_t(3 * s->L.ny) * sizeof(double)));
    // face transport: peer-to-peer stores unless PLB_FACE=nccl (A/B runs) or a
    // neighbour cannot be mapped
    const char *mode = getenv("PLB_FACE");
    if (!(mode && strcmp(mode, "nccl") == 0))
        if (int rc = setup_p2p(s)) return rc;
    return PLB_OK;
}

int plb_event_record(plb_handle s, int32_t slot)
{
    if (!s || slot < 0 || slot >= 8) return fail(PLB_ERR_INVALID, "bad event slot");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    CUDA_TRY(cudaEventRecord(s->events[slot], s->stream));
    return PLB_OK;
}

int plb_event_elapsed_ms(plb_handle s, int32_t a, int32_t b, float *ms)
{
    if (!s || !ms || a < 0 || a >= 8 || b < 0 || b >= 8)
        return fail(PLB_ERR_INVALID, "bad event slot");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    CUDA_TRY(cudaEventSynchronize(s->events[b]));
    CUDA_TRY(cudaEventElapsedTime(ms, s->events[a], s->events[b]));
    return PLB_OK;
}

int plb_profile_enable(plb_handle s, int32_t enable)
{
    if (!s) return fail(PLB_ERR_INVALID, "null handle");
    s->profile = enable != 0;
    s->prof_used = 0;
    return PLB_OK;
}

int plb_profile_read(plb_handle s, double *bulk_ms, int64_t *n_launches)
{
    if (!s || !bulk_ms || !n_launches) return fail(PLB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    double total = 0.0;
    for (size_t i = 0; i + 1 < s->prof_used; i += 2) {
        float ms = 0.f;
        CUDA_TRY(cudaEventSynchronize(s->prof_events[i + 1]));
        CUDA_TRY(cudaEventElapsedTime(&ms, s->prof_events[i], s->prof_events[i + 1]));
        total += ms;
    }
    *bulk_ms = total;
    *n_launches = int64_t(s->prof_used / 2);
    s->prof_used = 0;
    return PLB_OK;
}

int plb_info(plb_handle s, int64_t out[8])
{
    if (!s || !out) return fail(PLB_ERR_INVALID, "null argument");
    out[0] = s->n_bulk;
    out[1] = s->n_links;
    out[2] = s->n_solid;
    out[3] = s->L.pitch;
    out[4] = s->L.plane;
    out[5] = s->variant;
    // bulk nodes of the dominant (profiled) launch, and the face transport:
    // 0 none, 1 this rank's own ghost rows (periodic seam), 2 NCCL, 3 p2p
    const bool faces = s->cfg.left_neighbor || s->cfg.right_neighbor;
    out[6] = (faces && s->L.nx > 2)
                 ? s->n_bulk - s->n_bulk_edge[0] - s->n_bulk_edge[1]
                 : s->n_bulk;
    out[7] = !faces ? 0 : (!s->comm ? 1 : (s->p2p ? 3 : 2));
    return PLB_OK;
}

int64_t plb_kernel_launches(plb_handle s, int32_t reset)
{
    if (!s) return 0;
    cudaSetDevice(s->cfg.device);
    flush_pending(s);
    const int64_t n = s->launches;
    if (reset) s->launches = 0;
    return n;
}

int plb_device_pci_bus_id(int32_t device, char *buf, int32_t len)
{
    if (!buf || len < 16) return fail(PLB_ERR_INVALID, "buffer too small");
    CUDA_TRY(cudaDeviceGetPCIBusId(buf, len, device));
    return PLB_OK;
}

int plb_host_alloc(void **ptr, size_t bytes)
{
    if (!ptr) return fail(PLB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return PLB_OK;
}

int plb_host_free(void *ptr)
{
    CUDA_TRY(cudaFreeHost(ptr));
    return PLB_OK;
}

int plb_flush_l2(plb_handle s)
{
    if (!s) return fail(PLB_ERR_INVALID, "null handle");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    if (int rc = flush_pending(s)) return rc;
    const int64_t n = (int64_t(256) << 20) / 8;   // 256 MiB > 126 MB L2
    if (!s->flush_buf) CUDA_TRY(cudaMalloc(&s->flush_buf, size_t(n) * 8));
    s->launches += launch_fill(s->flush_buf, n, 0.0, s->stream);
    CUDA_TRY(cudaGetLastError());
    return PLB_OK;
}

const char *plb_build_info(void) { return kernel_build_info(); }

int plb_copy_bandwidth(plb_handle s, double *gbs)
{
    if (!s || !gbs) return fail(PLB_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(s->cfg.device));
    const size_t bytes = size_t(1) << 30;
    char *a = nullptr, *b = nullptr;
    CUDA_TRY(cudaMalloc(&a, bytes));
    if (cudaMalloc(&b, bytes) != cudaSuccess) {
        cudaFree(a);
        return fail(PLB_ERR_NOMEM, "no room for the 2 GiB copy probe");
    }
    CUDA_TRY(cudaMemsetAsync(a, 1, bytes, s->stream));
    CUDA_TRY(cudaMemsetAsync(b, 2, bytes, s->stream));
    float best = 0.f;
    for (int i = 0; i < 6; ++i) {
        CUDA_TRY(cudaEventRecord(s->events[6], s->stream));
        CUDA_TRY(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, s->stream));
        CUDA_TRY(cudaEventRecord(s->events[7], s->stream));
        CUDA_TRY(cudaEventSynchronize(s->events[7]));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, s->events[6], s->events[7]));
        if (i > 0 && (best == 0.f || ms < best)) best = ms;   // first pass warms up
    }
    CUDA_TRY(cudaFree(a));
    CUDA_TRY(cudaFree(b));
    *gbs = 2.0 * double(bytes) / (double(best) * 1e-3) / 1e9;
    return PLB_OK;
}

}  // extern "C"
