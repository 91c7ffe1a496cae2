/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's ray-cast hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libspica_b200.so) never does and has no CPU fallback.
 *
 * Parity status: PINNED.  The restatement is checked (tests/test_oracle.py) against
 *   - the reference's own live known-answer tests (tests/test_geometry.cc:50-77,119-125,
 *     tests/test_ray.cc:6-23), and
 *   - outputs of the unmodified reference compiled here (oracle/_ref/raycast_ref): BVH topology
 *     dumps and per-ray (prim, t) for box/bunny/kitten.ply, committed under tests/golden/.
 *
 * Every function cites the reference file:line (relative to /root/reference/sources) it follows.
 * All arithmetic is double, compiled -O2 -ffp-contract=off (x86-64 baseline: no FMA), like the
 * reference build the parity contract is defined against.
 */
#ifndef SPICA_ORACLE_H_
#define SPICA_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One node of the reference's binary BVH (accelerators/bvh.h:28-52), index-linked. */
typedef struct so_node {
    double lo[3], hi[3];
    int32_t left, right; /* -1 for leaves                        */
    int32_t prim;        /* >= 0 for leaves, -1 for forks        */
    int32_t axis;
} so_node;

/* core/ray.cc:11-19,43-47 + core/vector3d_detail.h:134-137,196-200:
 * dir = d * (1.0 / sqrt(d.d)); invdir = (dir==0) ? 1e32 : 1/dir.  Returns 0 for a zero direction
 * (the reference aborts there, tests/test_ray.cc:20-22). */
int so_ray_init(const double o[3], const double d[3], double dir_out[3], double invdir_out[3]);

/* core/triangle.cc:98-117 (closest) == :159-178 (any): Moeller-Trumbore on world-space points.
 * dir must already be normalised (so_ray_init). Returns 1 on hit and writes t,u,v. */
int so_triangle_intersect(const double tri[9], const double org[3], const double dir[3],
                          double max_dist, double* t, double* u, double* v);

/* core/bounds3d_detail.h:64-79: slab test over [0, max_dist]. */
int so_bounds_intersect(const double lo[3], const double hi[3], const double org[3],
                        const double invdir[3], double max_dist, double* t_near, double* t_far);

/* accelerators/bvh.cc:139-237: the top-down builder, 1 primitive per leaf, 2N-1 nodes written in
 * creation (pre-)order; root is node 0. tris = n x 9 doubles. nodes must hold 2n-1 entries.
 * Returns the node count. */
int64_t so_bvh_build(const double* tris, int64_t n, so_node* nodes);

/* rays: n x 8 {o[3], d[3], tmin(ignored), tmax}, float32 (ray_f64 = 0) or float64 (ray_f64 = 1);
 * each goes through so_ray_init first, as the reference's Ray constructor does. */

/* accelerators/bvh.cc:331-360 (+ core/primitive.cc:49-62): unordered DFS, right child first,
 * accept t <= maxDist and shrink. prim = -1 on miss. u,v may be NULL. */
void so_trace_closest(const so_node* nodes, int32_t root, const double* tris, const void* rays,
                      int ray_f64, int64_t n, int32_t* prim, double* t, double* u, double* v,
                      int threads);

/* accelerators/bvh.cc:362-387. */
void so_trace_any(const so_node* nodes, int32_t root, const double* tris, const void* rays,
                  int ray_f64, int64_t n, uint8_t* occluded, int threads);

/* Ground truth a la tests/test_trimesh.cc:153-205: every triangle in index order, same accept
 * rule (t <= maxDist, shrink). */
void so_trace_bruteforce(const double* tris, int64_t n_tris, const void* rays, int ray_f64,
                         int64_t n, int32_t* prim, double* t, int threads);

/* node / leaf visits of an ORDERED (near child first), tmax-pruned traversal of
 * the same binary tree; defines the algorithmic bytes per ray. sums are over all n rays. */
void so_count_ordered_visits(const so_node* nodes, int32_t root, const double* tris,
                             const void* rays, int ray_f64, int64_t n, int64_t* sum_inner,
                             int64_t* sum_leaf, int threads);

#ifdef __cplusplus
}
#endif
#endif
