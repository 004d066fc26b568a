/*
 * plb.h -- C ABI of libplb, the B200 (sm_100a) back end of PyLaBolt's fluidLB
 * time step (D2Q9, fp64).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy
 * types.  Every entry point replaces a piece of the reference's Python/numba
 * path (Malyadeep/pylabolt; citations are file:line in that repository):
 *
 *   reference seam: Solver.execute_single_time_step, bound in Solver.compile
 *   (pylabolt/solvers/fluidLB.py:273-280) and called once per step by
 *   Solver.run (:356).  A "b200" back end rebinds that slot to plb_step().
 *
 * Host arrays cross the boundary in the REFERENCE'S OWN LAYOUTS
 * (pylabolt/base/fields.py:50-92): one ghost ring, flat node index
 * ind = x * (ny + 2) + y with y contiguous, populations (size, 9), vectors
 * (size, 2), flags one byte per node.  The library owns all device memory
 * behind the opaque handle and keeps its own structure-of-arrays layout;
 * host pointers are borrowed for the duration of a call only.
 *
 * Error convention: every function returns PLB_OK (0) or a negative code;
 * plb_last_error() returns the message of the last failure on this thread.
 * A handle is driven by one host thread.  plb_step() is asynchronous; it is
 * ordered before any later plb_download / plb_sync on the same handle.
 */
#ifndef PLB_H
#define PLB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLB_ABI_VERSION 1

enum plb_status {
    PLB_OK = 0,
    PLB_ERR_INVALID = -1,   /* bad argument / configuration               */
    PLB_ERR_CUDA = -2,      /* CUDA runtime failure (message has details) */
    PLB_ERR_STATE = -3,     /* call made in the wrong state               */
    PLB_ERR_NCCL = -4,      /* NCCL failure or NCCL not loadable          */
    PLB_ERR_NOMEM = -5
};

/* collision_dict.fluid.model, pylabolt/solvers/fluidLB.py:35-40 */
enum plb_collision { PLB_BGK = 0, PLB_MRT = 1 };
/* collision_dict.fluid.forcing_model, pylabolt/solvers/fluidLB.py:28-34 */
enum plb_forcing { PLB_FORCING_NONE = 0, PLB_GUO_LINEAR = 1,
                   PLB_GUO_SECOND_ORDER = 2 };
/* boundary_dict.<name>.fluid.type, pylabolt/base/boundary.py:377-382;
 * zero_gradient is listed by README.rst:83 but has no upstream kernel. */
enum plb_bc_type { PLB_BC_BOUNCE_BACK = 0, PLB_BC_FIXED_VELOCITY = 1,
                   PLB_BC_FIXED_PRESSURE = 2, PLB_BC_PERIODIC = 3,
                   PLB_BC_ZERO_GRADIENT = 4 };

/* Host-side fields that can be uploaded / downloaded.  "padded" = the
 * reference layout with the ghost ring, (nx+2)*(ny+2) nodes; "inner" = ghost
 * ring stripped, nx*ny nodes, x-major (what
 * pylabolt/parallel/cpu/io_operator_kernels.py:5-52 produces for np.savez). */
enum plb_field {
    PLB_SOLID = 0,          /* uint8  padded  fields.solid (incl. ghost flags
                               filled by base/obstacle_operator.py:36-41)  */
    PLB_DENSITY = 1,        /* double padded  fields.density               */
    PLB_VELOCITY = 2,       /* double padded (size,2) fields.velocity      */
    PLB_POP = 3,            /* double padded (size,9) fields.pop_fluid_new */
    PLB_DENSITY_INNER = 4,  /* double inner                                */
    PLB_VELOCITY_INNER = 5  /* double inner (n,2)                          */
};

/* One rank's configuration.  Values are computed by the caller exactly as the
 * reference computes them and are used verbatim by the kernels. */
typedef struct plb_config {
    int32_t abi_version;       /* PLB_ABI_VERSION                                   */
    int32_t device;            /* CUDA device ordinal                               */
    int64_t nx, ny;            /* this rank's interior nodes: Domain.Nx_rank/Ny_rank,
                                  pylabolt/parallel/domain.py:54-76                 */
    int32_t x_periodic;        /* Boundary.x_periodic, base/boundary.py:553-556     */
    int32_t y_periodic;        /* Boundary.y_periodic                               */
    int32_t left_neighbor;     /* 1 if column x=-1 is fed by a neighbour rank or by
                                  the periodic image (MPIOperator.left_rank is not
                                  None, parallel/MPI_operator.py:131-136)           */
    int32_t right_neighbor;    /* same for column x=nx (MPI_operator.py:138-142)    */
    int32_t collision;         /* enum plb_collision                                */
    int32_t forcing;           /* enum plb_forcing                                  */
    double omega;              /* 1/tau, base/collision_operator.py:89-91           */
    double mrt_rates[9];       /* S, base/collision_operator.py:159-163             */
    double gravity[2];         /* ForceOperator.gravity, base/force_operator.py:59-79 */
    double inv_cs_2, inv_cs_4; /* Lattice, base/lattice.py:41-44                    */
    double float_min;          /* Control.float_min, base/control.py:49             */
    double weights[9];         /* Lattice.weights, base/lattice.py:54-57            */
} plb_config;

typedef struct plb_solver *plb_handle;

/* ---- life cycle -------------------------------------------------------- */

/* Allocates the device lattices (two-lattice SoA, fp64) for one rank.
 * Replaces State.set_backend / Fields.set_backend device mirroring
 * (pylabolt/base/fields.py:192-227). */
int plb_create(const plb_config *config, plb_handle *out);
void plb_destroy(plb_handle h);
const char *plb_last_error(void);

/* ---- geometry ---------------------------------------------------------- */

/* One boundary element, in boundary_dict order (later elements win shared
 * links, base/boundary_operator.py:153-157).  Mirrors BoundaryElement
 * (pylabolt/base/boundary.py:7-127): boundary_nodes are padded flat indices,
 * out_list / inv_list the three outgoing / incoming directions, normal the
 * inward surface normal as integers, vector / scalar the fluid value. */
int plb_add_boundary_element(plb_handle h, int32_t bc_type,
                             const int64_t *boundary_nodes, int64_t n_nodes,
                             const int64_t out_list[3],
                             const int64_t inv_list[3],
                             const int64_t normal[2],
                             const double vector_fluid[2],
                             double scalar_fluid);

/* Classifies every node from the uploaded PLB_SOLID flags and the boundary
 * elements (bulk / skip / link node), builds the link lists and the slab-face
 * masks and uploads them.  Must be called once after PLB_SOLID and all
 * elements are set and before plb_initialize_pop / plb_step. */
int plb_finalize_geometry(plb_handle h);

/* ---- data -------------------------------------------------------------- */

int plb_upload(plb_handle h, int32_t field, const void *host, size_t bytes);
int plb_download(plb_handle h, int32_t field, void *host, size_t bytes);

/* A uniform field without a host array: PLB_DENSITY (value[0]) or PLB_VELOCITY
 * (value[0], value[1]) on every interior node, 0 on the ghost ring -- what
 * set_field_scalar / set_field_vector (pylabolt/base/init_fields.py:325-375)
 * leave behind for a `type: "fixed"` entry of initial_fields_dict, produced on
 * the device instead of being copied over PCIe.  Solid nodes carry the body's
 * own density / velocity in the reference; a case with obstacles uploads. */
int plb_fill(plb_handle h, int32_t field, const double *value);

/* f = f_eq(rho, u) on fluid nodes, 0 elsewhere: CollisionOperator.initialize_pop,
 * pylabolt/parallel/cpu/equilibrium_kernels.py:38-78. */
int plb_initialize_pop(plb_handle h);

/* ---- the hot path ------------------------------------------------------ */

/* Advances n_steps reference time steps (Solver.single_time_step,
 * pylabolt/solvers/fluidLB.py:206-253, phases 2-8 fused).  flags apply to the
 * LAST of the n_steps:
 *   PLB_STORE_MOMENTS  the step also stores rho and u (phases 2-4 of that
 *                      step), which is what fields.density / fields.velocity
 *                      hold after the reference has executed the same steps;
 *   PLB_RECORD_LINKS   the step also records, per link node and direction,
 *                      the momentum it exchanged with a wall / solid / edge,
 *                      for plb_download_link_exchange. */
#define PLB_STORE_MOMENTS 1
#define PLB_RECORD_LINKS 2
int plb_step(plb_handle h, int64_t n_steps, int32_t flags);
int plb_sync(plb_handle h);
/* Several steps per pass.  Steps that need neither flag are advanced several
 * at a time where the geometry allows it: nodes whose whole neighbourhood is
 * plain fluid go through all steps of a group on chip (lattice read once,
 * written once: 144 / d B per node and step for d steps per pass), all other
 * nodes through d ordinary passes hidden behind that kernel.  The result is
 * the one of single steps (same per-node arithmetic).  Because plb_step() is
 * asynchronous, plain steps that do not fill a group may be held back until
 * the rest arrives; every other entry point first completes what was held
 * back.  Environment: PLB_FUSE=0 disables the path, PLB_FUSE=2 uses it on any
 * lattice that has such nodes (default: lattices where they dominate);
 * PLB_FUSE_DEPTH=2 / 3 / 4 sets the steps per pass (default: per collision
 * model, from measurement; 48 / 36 B per node and step at 3 / 4; a depth-d
 * pass advances the nodes whose neighbourhood of radius d - 1 is plain fluid).
 * out = {steps per pass in use (0: off), nodes advanced by the two-step kernel,
 *        by the three-step kernel, nodes of the first list pass,
 *        two-step passes executed, rows per warp chunk, y strips,
 *        three-step passes executed, nodes advanced by the four-step kernel,
 *        four-step passes executed} */
int plb_fused_info(plb_handle h, int64_t out[10]);

/* Device memory held by the handle, in bytes (the reference keeps its
 * `*_device` mirrors for the process lifetime, base/fields.py:192-227):
 * out = {the two lattices, moment planes (rho, u and the residue's old copy),
 *        compact scratch lattices of the several-steps-per-pass path incl.
 *        their index map, node codes / deep flags / lists / staging,
 *        free device memory, total device memory} */
int plb_memory_info(plb_handle h, int64_t out[6]);

/* ---- diagnostics ("next" rows) ----------------------------------------- */

/* Residue sums of utils/residues.py:171-222 on the stored moments:
 * out = {num_rho, den_rho, num_ux, den_ux, num_uy, den_uy}; the library keeps
 * field_old and updates it, like cpu/compute_residues_kernels.py:6-73. */
int plb_residue_sums(plb_handle h, double out[6]);

/* Momentum exchange for wall and obstacle forces
 * (pylabolt/parallel/cpu/force_torque_kernels.py:11-141).  Link nodes are the
 * fluid nodes on a domain edge or next to a solid node.  plb_link_nodes
 * returns their padded flat indices in list order (padded_index may be NULL to
 * query the count).  plb_download_link_exchange returns, for link node i and
 * direction k = 1..8, out[8*i + k-1] = pop[i,k] + pop_new[i,inv k] of the last
 * step run with PLB_RECORD_LINKS when the node itself wrote pop_new[i,inv k]
 * (bounce back, boundary element, uncovered edge), else 0; the force of the
 * link is c_k times that value.  Links owned by zero_gradient elements are
 * not recorded. */
int plb_link_nodes(plb_handle h, int64_t *padded_index, int64_t capacity,
                   int64_t *n_links);
int plb_download_link_exchange(plb_handle h, double *out, int64_t n_values);

/* ---- multi-GPU (x-slabs, one process per GPU) -------------------------- */

/* 128-byte NCCL unique id, created on one rank and broadcast by the caller
 * (torch.distributed is only the bootstrap). */
int plb_comm_unique_id(void *id128);
/* Joins the slab ring.  left_rank / right_rank are the ranks that own the
 * neighbouring slabs (-1 = none).  Replaces MPIOperator.find_neighbor_ranks
 * and halo_exchange, pylabolt/parallel/MPI_operator.py:116-259.
 * Collective over all ranks.  The three populations that cross a face are
 * stored by the kernels that produce them directly into the neighbour's
 * receive buffer (CUDA IPC mapping over NVLink, mailbox hand-shake); if a
 * neighbour cannot be mapped, or with PLB_FACE=nccl in the environment, they
 * travel by ncclSend / ncclRecv instead. */
int plb_comm_init(plb_handle h, const void *id128, int32_t rank,
                  int32_t n_ranks, int32_t left_rank, int32_t right_rank);

/* ---- measurement helpers ----------------------------------------------- */

/* CUDA events on the library's own stream (slot 0..7). */
int plb_event_record(plb_handle h, int32_t slot);
int plb_event_elapsed_ms(plb_handle h, int32_t start_slot, int32_t stop_slot,
                         float *ms);
/* Per-kernel timing of the dominant (bulk collide-stream) kernel: while
 * enabled, every bulk launch is bracketed by CUDA events on the launching
 * stream.  plb_profile_read returns the summed duration and the number of
 * launches since the last read (it synchronises the events).  Only the
 * dominant launch of a step is bracketed (all columns, or all but the two
 * slab-edge columns when the slab has faces); plb_info reports how many bulk
 * nodes it covers. */
int plb_profile_enable(plb_handle h, int32_t enable);
int plb_profile_read(plb_handle h, double *bulk_ms, int64_t *n_launches);
/* out = {n_bulk, n_link, n_solid, pitch, plane, kernel_variant,
 *        bulk nodes of the profiled (dominant) bulk launch,
 *        face transport: 0 none, 1 own ghost rows (single-rank periodic seam),
 *                        2 NCCL send/recv, 3 peer-to-peer stores} */
int plb_info(plb_handle h, int64_t out[8]);
/* Kernels launched by this handle since creation / since the last reset. */
int64_t plb_kernel_launches(plb_handle h, int32_t reset);
/* PCI bus id ("0000:1b:00.0") of a CUDA device ordinal, so that the host side
 * can bind the rank to the NUMA node the GPU hangs off before it allocates
 * its transfer buffers (no handle needed). */
int plb_device_pci_bus_id(int32_t device, char *buf, int32_t len);
/* Pinned host memory for upload / download buffers. */
int plb_host_alloc(void **ptr, size_t bytes);
int plb_host_free(void *ptr);
/* Writes a scratch buffer larger than L2 (bench hygiene). */
int plb_flush_l2(plb_handle h);
/* Device-to-device copy rate of this GPU right now, in GB/s (read + write
 * bytes of a 1 GiB cudaMemcpyAsync, best of 5): the practical HBM ceiling
 * next to which a roofline fraction is read. */
int plb_copy_bandwidth(plb_handle h, double *gbs);
/* Compile-time configuration of the kernels in this build of the library
 * ("fused: block=128 minblocks=3 ring=cp.async stages=2 carry=registers ..."):
 * which tuning variant a bench line or a sweep was measured with.  Static
 * string, never NULL. */
const char *plb_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* PLB_H */
