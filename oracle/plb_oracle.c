/*
 * plb_oracle.c -- CPU restatement of the reference's fluidLB time step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the timed CPU baseline.  pylabolt_b200 never imports it.
 *
 * What it restates: Malyadeep/pylabolt, pylabolt/solvers/fluidLB.py:206-253
 * (Solver.single_time_step) and the numba CPU kernels it dispatches, pass by
 * pass, on the reference's own array-of-structures layouts:
 *     pop[size][9], pop_new[size][9], velocity[size][2], force[size][2],
 *     density[size], solid[size] (bool), ghost[size] (bool),
 *     ind = x * Ny_pad + y   (y contiguous, one ghost ring).
 * Every function cites the reference file:line it follows.  Expression trees
 * are kept exactly as Python parses them, and the file is compiled with
 * -ffp-contract=off, so on BGK paths the results are BIT-IDENTICAL to the
 * numba kernels (numba's @njit emits no FMA).
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against
 * the .npz fixtures in tests/golden, which tests/golden/make_golden.py produced by running
 * the reference itself.  Two features have NO upstream kernel and are our
 * own definitions -- "parity unpinned upstream": MRT collision
 * (oracle_collide_mrt) and the zero_gradient boundary (oracle_bc_zero_gradient).
 *
 * OpenMP `parallel for` stands where the reference has numba `prange`, so the
 * five-pass structure, the memory traffic and the threading model match the
 * reference's CPU path; that is what makes this a fair CPU baseline "port".
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Q 9

/* base/lattice.py:50-60 */
static const int64_t CX[Q] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int64_t CY[Q] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const int64_t INV[Q] = {0, 3, 4, 1, 2, 7, 8, 5, 6};

typedef struct {
    int64_t nx_pad, ny_pad;          /* domain.shape, parallel/domain.py:78-80 */
    double inv_cs_2, inv_cs_4;       /* base/lattice.py:41-44                 */
    double float_min;                /* base/control.py:49                    */
    double weights[Q];               /* base/lattice.py:54-57                 */
    double omega;                    /* base/collision_operator.py:89-91      */
    double gravity[2];               /* base/force_operator.py:59-79          */
    double s[Q];                     /* MRT rates, collision_operator.py:159  */
    int32_t collision;               /* 0 = BGK, 1 = MRT                      */
    int32_t forcing;                 /* 0 None, 1 guo_linear, 2 guo_second_order */
    int32_t x_periodic, y_periodic;  /* base/boundary.py:553-556              */
} oracle_params;

typedef struct {
    uint8_t *solid, *ghost;
    double *density, *velocity, *force, *pop, *pop_new;
} oracle_fields;

/* boundary element, base/boundary.py:7-127 */
typedef struct {
    int32_t type;             /* 0 bounce_back 1 fixed_velocity 2 fixed_pressure
                                 3 periodic 4 zero_gradient (ours)            */
    int64_t n_nodes;
    const int64_t *nodes;     /* boundary_nodes                                */
    int64_t out_list[3], inv_list[3];
    int64_t normal[2];        /* surface_normals cast to int (shim 4)          */
    double vector[2];         /* vector_fluid                                  */
    double scalar;            /* scalar_fluid                                  */
} oracle_element;

void oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* second-order equilibrium, cpu/equilibrium_kernels.py:26-35 and the
 * identical inline copies in cpu/collision_kernels.py:36-44 */
static inline double feq(const oracle_params *p, int k, double rho, double ux,
                         double uy, double u2)
{
    double cu = (double)CX[k] * ux + (double)CY[k] * uy;
    return p->weights[k] * rho *
           (1 + p->inv_cs_2 * cu + 0.5 * p->inv_cs_4 * cu * cu -
            0.5 * p->inv_cs_2 * u2);
}

/* cpu/equilibrium_kernels.py:38-78 initialize_pop_density_based_second_order */
void oracle_initialize_pop(const oracle_params *p, oracle_fields *f)
{
    const int64_t size = p->nx_pad * p->ny_pad;
#pragma omp parallel for schedule(static)
    for (int64_t ind = 0; ind < size; ++ind) {
        if (!f->solid[ind] && !f->ghost[ind]) {
            double rho = f->density[ind];
            double ux = f->velocity[2 * ind], uy = f->velocity[2 * ind + 1];
            double u2 = ux * ux + uy * uy;
            for (int k = 0; k < Q; ++k) {
                double e = feq(p, k, rho, ux, uy, u2);
                f->pop[Q * ind + k] = e;
                f->pop_new[Q * ind + k] = e;
            }
        }
    }
}

/* cpu/compute_fields_kernels.py:5-28 density_compute_density_based */
void oracle_density(const oracle_params *p, oracle_fields *f)
{
    const int64_t size = p->nx_pad * p->ny_pad;
#pragma omp parallel for schedule(static)
    for (int64_t ind = 0; ind < size; ++ind) {
        if (!f->solid[ind] && !f->ghost[ind]) {
            double sum = 0.;
            const double *pl = f->pop_new + Q * ind;
            for (int k = 0; k < Q; ++k) sum += pl[k];
            f->density[ind] = sum;
        }
    }
}

/* cpu/force_field_kernels.py:5-25 compute_gravity_force */
void oracle_gravity_force(const oracle_params *p, oracle_fields *f)
{
    const int64_t size = p->nx_pad * p->ny_pad;
#pragma omp parallel for schedule(static)
    for (int64_t ind = 0; ind < size; ++ind) {
        if (!f->solid[ind] && !f->ghost[ind]) {
            double rho = f->density[ind];
            f->force[2 * ind] = rho * p->gravity[0];
            f->force[2 * ind + 1] = rho * p->gravity[1];
        }
    }
}

/* cpu/compute_fields_kernels.py:31-65 velocity_compute_density_based */
void oracle_velocity(const oracle_params *p, oracle_fields *f)
{
    const int64_t size = p->nx_pad * p->ny_pad;
#pragma omp parallel for schedule(static)
    for (int64_t ind = 0; ind < size; ++ind) {
        if (!f->solid[ind] && !f->ghost[ind]) {
            double sx = 0., sy = 0.;
            const double *pl = f->pop_new + Q * ind;
            for (int k = 0; k < Q; ++k) {
                sx += (double)CX[k] * pl[k];
                sy += (double)CY[k] * pl[k];
            }
            double inv = 1 / (f->density[ind] + p->float_min);
            f->velocity[2 * ind] = sx * inv + 0.5 * f->force[2 * ind] * inv;
            f->velocity[2 * ind + 1] =
                sy * inv + 0.5 * f->force[2 * ind + 1] * inv;
        }
    }
}

/* Guo source term without the (1 - omega/2) prefactor.
 * forcing 1: cpu/collision_kernels.py:90-92 ; forcing 2: :142-148 */
static inline double guo_term(const oracle_params *p, int k, double ux,
                              double uy, double fx, double fy)
{
    if (p->forcing == 1) {
        return p->weights[k] * ((double)CX[k] * fx + (double)CY[k] * fy) *
               p->inv_cs_2;
    } else {
        double cu = (double)CX[k] * ux + (double)CY[k] * uy;
        double const_x = ((double)CX[k] - ux) * p->inv_cs_2 +
                         cu * (double)CX[k] * p->inv_cs_4;
        double const_y = ((double)CY[k] - uy) * p->inv_cs_2 +
                         cu * (double)CY[k] * p->inv_cs_4;
        return p->weights[k] * (const_x * fx + const_y * fy);
    }
}

/* cpu/collision_kernels.py:5-45 (None), :48-97 (guo_linear),
 * :100-153 (guo_second_order): reads pop_new, writes pop */
void oracle_collide_bgk(const oracle_params *p, oracle_fields *f)
{
    const int64_t size = p->nx_pad * p->ny_pad;
    const double omega = p->omega;
#pragma omp parallel for schedule(static)
    for (int64_t ind = 0; ind < size; ++ind) {
        if (!f->solid[ind] && !f->ghost[ind]) {
            double rho = f->density[ind];
            double ux = f->velocity[2 * ind], uy = f->velocity[2 * ind + 1];
            double fx = f->force[2 * ind], fy = f->force[2 * ind + 1];
            double u2 = ux * ux + uy * uy;
            for (int k = 0; k < Q; ++k) {
                double e = feq(p, k, rho, ux, uy, u2);
                if (p->forcing == 0) {
                    f->pop[Q * ind + k] =
                        (1 - omega) * f->pop_new[Q * ind + k] + omega * e;
                } else {
                    double ft = guo_term(p, k, ux, uy, fx, fy);
                    f->pop[Q * ind + k] =
                        (1 - omega) * f->pop_new[Q * ind + k] + omega * e +
                        (1 - 0.5 * omega) * ft;
                }
            }
        }
    }
}

/* Lallemand-Luo moment matrix, base/collision_operator.py:147-157
 * (rows rho, e, eps, jx, qx, jy, qy, pxx, pxy; columns = direction k). */
static const double MRT_M[Q][Q] = {
    {1, 1, 1, 1, 1, 1, 1, 1, 1},
    {-4, -1, -1, -1, -1, 2, 2, 2, 2},
    {4, -2, -2, -2, -2, 1, 1, 1, 1},
    {0, 1, 0, -1, 0, 1, -1, -1, 1},
    {0, -2, 0, 2, 0, 1, -1, -1, 1},
    {0, 0, 1, 0, -1, 1, 1, -1, -1},
    {0, 0, -2, 0, 2, 1, 1, -1, -1},
    {0, 1, -1, 1, -1, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 1, -1, 1, -1}};
/* rows of M are orthogonal: inv(M) = M^T diag(1/|row|^2) */
static const double MRT_NORM2[Q] = {9, 36, 36, 6, 12, 6, 12, 4, 4};

/* MRT -- OUR DEFINITION, parity unpinned upstream (the reference has the
 * matrices, base/collision_operator.py:147-163, but no kernel, and its setup
 * code is broken at :93 and :164-165).  SURVEY.md App. A.2:
 *   g = f - Minv diag(S) M (f - feq) + Minv (I - diag(S)/2) M Phi
 * with the same second-order feq as BGK and Phi = the Guo bracket.  For
 * S = omega * 1 this is algebraically the BGK update above. */
void oracle_collide_mrt(const oracle_params *p, oracle_fields *f)
{
    const int64_t size = p->nx_pad * p->ny_pad;
#pragma omp parallel for schedule(static)
    for (int64_t ind = 0; ind < size; ++ind) {
        if (!f->solid[ind] && !f->ghost[ind]) {
            double rho = f->density[ind];
            double ux = f->velocity[2 * ind], uy = f->velocity[2 * ind + 1];
            double fx = f->force[2 * ind], fy = f->force[2 * ind + 1];
            double u2 = ux * ux + uy * uy;
            double fneq[Q], phi[Q], m[Q], mphi[Q];
            for (int k = 0; k < Q; ++k) {
                fneq[k] = f->pop_new[Q * ind + k] - feq(p, k, rho, ux, uy, u2);
                phi[k] = p->forcing ? guo_term(p, k, ux, uy, fx, fy) : 0.0;
            }
            for (int r = 0; r < Q; ++r) {
                double a = 0., b = 0.;
                for (int k = 0; k < Q; ++k) {
                    a += MRT_M[r][k] * fneq[k];
                    b += MRT_M[r][k] * phi[k];
                }
                m[r] = p->s[r] * a / MRT_NORM2[r];
                mphi[r] = (1 - 0.5 * p->s[r]) * b / MRT_NORM2[r];
            }
            for (int k = 0; k < Q; ++k) {
                double relax = 0., src = 0.;
                for (int r = 0; r < Q; ++r) {
                    relax += MRT_M[r][k] * m[r];
                    src += MRT_M[r][k] * mphi[r];
                }
                f->pop[Q * ind + k] = f->pop_new[Q * ind + k] - relax + src;
            }
        }
    }
}

/* Periodic ghost fill: x-phase on full columns, then y-phase on full rows so
 * that corners are right.  Wrap semantics of gpu/MPI_kernels.py:34-99 and of
 * the multi-rank CPU exchange parallel/MPI_operator.py:155-259 (the
 * single-rank self-Sendrecv mirror is an upstream defect). */
void oracle_wrap_ghosts(const oracle_params *p, void *field, int64_t item_bytes)
{
    const int64_t nxp = p->nx_pad, nyp = p->ny_pad;
    char *base = (char *)field;
    if (p->x_periodic) {
        memcpy(base, base + (nxp - 2) * nyp * item_bytes, nyp * item_bytes);
        memcpy(base + (nxp - 1) * nyp * item_bytes, base + nyp * item_bytes,
               nyp * item_bytes);
    }
    if (p->y_periodic) {
        for (int64_t x = 0; x < nxp; ++x) {
            char *row = base + x * nyp * item_bytes;
            memcpy(row, row + (nyp - 2) * item_bytes, item_bytes);
            memcpy(row + (nyp - 1) * item_bytes, row + item_bytes, item_bytes);
        }
    }
}

/* cpu/streaming_kernels.py:5-47 scalar_based_kernel: pull + halfway
 * bounce back off `solid` with the moving-wall term */
void oracle_stream(const oracle_params *p, oracle_fields *f)
{
    const int64_t size = p->nx_pad * p->ny_pad, nyp = p->ny_pad;
#pragma omp parallel for schedule(static)
    for (int64_t ind = 0; ind < size; ++ind) {
        if (!f->solid[ind] && !f->ghost[ind]) {
            int64_t x = ind / nyp, y = ind - x * nyp;
            double rho = f->density[ind];
            f->pop_new[Q * ind] = f->pop[Q * ind];
            for (int k = 1; k < Q; ++k) {
                int64_t nb = (x - CX[k]) * nyp + (y - CY[k]);
                if (!f->solid[nb]) {
                    f->pop_new[Q * ind + k] = f->pop[Q * nb + k];
                } else {
                    int64_t ki = INV[k];
                    double temp = 2 * p->weights[ki] * rho * p->inv_cs_2 *
                                  ((double)CX[ki] * f->velocity[2 * nb] +
                                   (double)CY[ki] * f->velocity[2 * nb + 1]);
                    f->pop_new[Q * ind + k] = f->pop[Q * ind + ki] - temp;
                }
            }
        }
    }
}

/* cpu/fluid_boundary_kernels.py:5-25 bounce_back */
static void bc_bounce_back(oracle_fields *f, const oracle_element *e)
{
#pragma omp parallel for schedule(static)
    for (int64_t it = 0; it < e->n_nodes; ++it) {
        int64_t ind = e->nodes[it];
        if (!f->solid[ind])
            for (int k = 0; k < 3; ++k)
                f->pop_new[Q * ind + e->inv_list[k]] =
                    f->pop[Q * ind + e->out_list[k]];
    }
}

/* cpu/fluid_boundary_kernels.py:28-62 fixed_velocity_density_based */
static void bc_fixed_velocity(const oracle_params *p, oracle_fields *f,
                              const oracle_element *e)
{
#pragma omp parallel for schedule(static)
    for (int64_t it = 0; it < e->n_nodes; ++it) {
        int64_t ind = e->nodes[it];
        if (!f->solid[ind]) {
            double rho = f->density[ind];
            for (int k = 0; k < 3; ++k) {
                int64_t o = e->out_list[k], v = e->inv_list[k];
                double temp = 2 * p->weights[o] * rho * p->inv_cs_2 *
                              ((double)CX[o] * e->vector[0] +
                               (double)CY[o] * e->vector[1]);
                f->pop_new[Q * ind + v] = f->pop[Q * ind + o] - temp;
            }
        }
    }
}

/* cpu/fluid_boundary_kernels.py:100-152 fixed_pressure_density_based
 * (anti bounce back; surface_normals as integers, shim 4) */
static void bc_fixed_pressure(const oracle_params *p, oracle_fields *f,
                              const oracle_element *e)
{
    const int64_t nyp = p->ny_pad;
#pragma omp parallel for schedule(static)
    for (int64_t it = 0; it < e->n_nodes; ++it) {
        int64_t ind = e->nodes[it];
        if (!f->solid[ind]) {
            int64_t i = ind / nyp, j = ind - i * nyp;
            double ux = f->velocity[2 * ind], uy = f->velocity[2 * ind + 1];
            int64_t nrm = (i + e->normal[0]) * nyp + (j + e->normal[1]);
            double ex = ux + 0.5 * (ux - f->velocity[2 * nrm]);
            double ey = uy + 0.5 * (uy - f->velocity[2 * nrm + 1]);
            double u2 = ex * ex + ey * ey;
            for (int k = 0; k < 3; ++k) {
                int64_t o = e->out_list[k], v = e->inv_list[k];
                double cu = (double)CX[o] * ex + (double)CY[o] * ey;
                double temp = 2 * p->weights[o] * e->scalar *
                              (1 + 0.5 * p->inv_cs_4 * cu * cu -
                               0.5 * p->inv_cs_2 * u2);
                f->pop_new[Q * ind + v] = -f->pop[Q * ind + o] + temp;
            }
        }
    }
}

/* zero_gradient -- OUR DEFINITION, parity unpinned upstream (listed in
 * README.rst:83 and docs/boundary_conditions.rst:26, absent from
 * base/boundary.py:377-382, no kernel anywhere).  The three unknown incoming
 * populations are copied from the neighbour along the inward normal after
 * streaming.  zero_gradient elements are applied in a final pass, in
 * boundary_dict order, after every other element type. */
static void bc_zero_gradient(const oracle_params *p, oracle_fields *f,
                             const oracle_element *e)
{
    const int64_t nyp = p->ny_pad;
#pragma omp parallel for schedule(static)
    for (int64_t it = 0; it < e->n_nodes; ++it) {
        int64_t ind = e->nodes[it];
        if (!f->solid[ind]) {
            int64_t i = ind / nyp, j = ind - i * nyp;
            int64_t nrm = (i + e->normal[0]) * nyp + (j + e->normal[1]);
            for (int k = 0; k < 3; ++k) {
                int64_t v = e->inv_list[k];
                f->pop_new[Q * ind + v] = f->pop_new[Q * nrm + v];
            }
        }
    }
}

/* base/boundary_operator.py:138-157 set_boundary_cpu: one kernel per element
 * in boundary_dict order; periodic elements have no kernel (:325-339). */
void oracle_set_boundary(const oracle_params *p, oracle_fields *f,
                         const oracle_element *elements, int32_t n_elements)
{
    for (int32_t n = 0; n < n_elements; ++n) {
        const oracle_element *e = elements + n;
        if (e->type == 0) bc_bounce_back(f, e);
        else if (e->type == 1) bc_fixed_velocity(p, f, e);
        else if (e->type == 2) bc_fixed_pressure(p, f, e);
    }
    for (int32_t n = 0; n < n_elements; ++n)
        if (elements[n].type == 4) bc_zero_gradient(p, f, elements + n);
}

/* solvers/fluidLB.py:206-253 single_time_step, phases 2-8 (phase 1 and 9 are
 * obstacle motion / force reduction: no-ops for static bodies). */
void oracle_step(const oracle_params *p, oracle_fields *f,
                 const oracle_element *elements, int32_t n_elements,
                 int64_t n_steps)
{
    for (int64_t s = 0; s < n_steps; ++s) {
        oracle_density(p, f);                       /* :211-215 */
        oracle_gravity_force(p, f);                 /* :216-219 */
        oracle_velocity(p, f);                      /* :220-224 */
        if (p->collision == 0) oracle_collide_bgk(p, f);   /* :225-229 */
        else oracle_collide_mrt(p, f);
        oracle_wrap_ghosts(p, f->pop, Q * sizeof(double)); /* :230-234 */
        oracle_stream(p, f);                        /* :235-239 */
        oracle_set_boundary(p, f, elements, n_elements);   /* :240-244 */
    }
}

/* ------------------------------------------------------------------------
 * "next" rows of SURVEY.md section 8(f)
 * ------------------------------------------------------------------------ */

/* cpu/compute_residues_kernels.py:6-35 (scalar) and :38-73 (vector):
 * numerator = sum (phi - phi_old)^2, denominator = sum phi_old^2 over fluid
 * nodes, per component; phi_old <- phi.  utils/residues.py:171-222 then takes
 * sqrt(num / (den + float_min)).  The numba prange reduction order is
 * unspecified, so sums are compared to 1e-12 relative, not bit for bit.
 * out = {num_rho, den_rho, num_ux, den_ux, num_uy, den_uy}. */
void oracle_residue_sums(const oracle_params *p, oracle_fields *f,
                         double *density_old, double *velocity_old,
                         double out[6])
{
    const int64_t size = p->nx_pad * p->ny_pad;
    double nr = 0, dr = 0, nx = 0, dx = 0, ny = 0, dy = 0;
#pragma omp parallel for schedule(static) reduction(+ : nr, dr, nx, dx, ny, dy)
    for (int64_t ind = 0; ind < size; ++ind) {
        if (!f->solid[ind] && !f->ghost[ind]) {
            double old = density_old[ind], cur = f->density[ind];
            double d = cur - old;
            nr += d * d;
            dr += old * old;
            density_old[ind] = cur;
            double ox = velocity_old[2 * ind], oy = velocity_old[2 * ind + 1];
            double cx = f->velocity[2 * ind], cy = f->velocity[2 * ind + 1];
            double ddx = cx - ox, ddy = cy - oy;
            nx += ddx * ddx;
            ny += ddy * ddy;
            dx += ox * ox;
            dy += oy * oy;
            velocity_old[2 * ind] = cx;
            velocity_old[2 * ind + 1] = cy;
        }
    }
    out[0] = nr; out[1] = dr; out[2] = nx; out[3] = dx; out[4] = ny; out[5] = dy;
}

/* cpu/force_torque_kernels.py:100-141 compute_boundary_force_single_phase:
 * momentum exchange over the three links of every (non-solid) node of one
 * boundary element, with pop = post-collision and pop_new = post-stream
 * populations of the same step.  Reduction order is unspecified upstream. */
void oracle_boundary_force(const oracle_fields *f, const oracle_element *e,
                           double out[2])
{
    double fx = 0., fy = 0.;
#pragma omp parallel for schedule(static) reduction(+ : fx, fy)
    for (int64_t it = 0; it < e->n_nodes; ++it) {
        int64_t ind = e->nodes[it];
        if (!f->solid[ind]) {
            for (int k = 0; k < 3; ++k) {
                int64_t o = e->out_list[k], v = e->inv_list[k];
                fx += f->pop[Q * ind + o] * (double)CX[o] -
                      f->pop_new[Q * ind + v] * (double)CX[v];
                fy += f->pop[Q * ind + o] * (double)CY[o] -
                      f->pop_new[Q * ind + v] * (double)CY[v];
            }
        }
    }
    out[0] = fx; out[1] = fy;
}

/* cpu/force_torque_kernels.py:11-94 compute_force_torque_single_phase:
 * force and torque (about ref_point, minimum image) on the obstacle
 * `current_solid_id`, summed over its fluid boundary nodes and their links
 * into solid nodes. */
void oracle_obstacle_force_torque(const oracle_params *p, const oracle_fields *f,
                                  const int64_t *solid_id,
                                  const uint8_t *fluid_boundary,
                                  const int64_t offset[2],
                                  const int64_t grid_global_shape[2],
                                  const double ref_point[2],
                                  int64_t current_solid_id, double out[3])
{
    const int64_t size = p->nx_pad * p->ny_pad, nyp = p->ny_pad;
    const double Nx = (double)grid_global_shape[0], Ny = (double)grid_global_shape[1];
    double fx = 0., fy = 0., tq = 0.;
#pragma omp parallel for schedule(static) reduction(+ : fx, fy, tq)
    for (int64_t ind = 0; ind < size; ++ind) {
        if (!f->ghost[ind] && fluid_boundary[ind] &&
            solid_id[ind] == current_solid_id) {
            int64_t x = ind / nyp, y = ind - x * nyp;
            double rx = (double)(x - 1 + offset[0]) - ref_point[0];
            double ry = (double)(y - 1 + offset[1]) - ref_point[1];
            double rx_min = rx, ry_min = ry;
            if (p->x_periodic) {
                if (__builtin_fabs(rx + Nx) < __builtin_fabs(rx_min)) rx_min = rx + Nx;
                if (__builtin_fabs(rx - Nx) < __builtin_fabs(rx_min)) rx_min = rx - Nx;
            }
            if (p->y_periodic) {
                if (__builtin_fabs(ry + Ny) < __builtin_fabs(ry_min)) ry_min = ry + Ny;
                if (__builtin_fabs(ry - Ny) < __builtin_fabs(ry_min)) ry_min = ry - Ny;
            }
            for (int k = 0; k < Q; ++k) {
                int64_t nb = (x + CX[k]) * nyp + (y + CY[k]);
                if (f->solid[nb]) {
                    int64_t ki = INV[k];
                    double vx = f->pop[Q * ind + k] * (double)CX[k] -
                                f->pop_new[Q * ind + ki] * (double)CX[ki];
                    double vy = f->pop[Q * ind + k] * (double)CY[k] -
                                f->pop_new[Q * ind + ki] * (double)CY[ki];
                    fx += vx;
                    fy += vy;
                    tq += rx_min * vy - ry_min * vx;
                }
            }
        }
    }
    out[0] = fx; out[1] = fy; out[2] = tq;
}
