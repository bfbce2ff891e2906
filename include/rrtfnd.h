/*
 * rrtfnd.h — C ABI of the B200-native RRT / RRT* / RRT*-FN / RRT*-FND hot path.
 *
 * The reference (grzesiek2201/RRT-STAR-FND-Algorithm) is pure Python and has no FFI layer; its
 * boundary is the Python surface of src/algorithm.py, src/graph.py and src/map.py. Every entry point
 * below names the reference function(s) it replaces (file:line relative to the reference root). The
 * Python shim in rrt_star_fnd_algorithm_b200/ binds these with ctypes.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - no allocation or ownership transfer inside the device-pointer entry points;
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream);
 *   - return value: 0 = OK, > 0 = cudaError_t, < 0 = RRTFND_E_* below;
 *   - all coordinates / costs are IEEE fp64 evaluated one rounding per written operation (no FMA
 *     contraction) except the one fused multiply-add of `calc_distance` (DESIGN.md "Arithmetic contract").
 *
 * The same parameter / state structs are used by the CPU oracle (oracle/rrt_oracle.c), which is test
 * infrastructure only.
 */
#ifndef RRTFND_H
#define RRTFND_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRTFND_ABI_VERSION 3

/* planner kinds: src/algorithm.py:55 (RRT), :98 (RRT_star), :190 (RRT_star_FN) */
#define RRTFND_MODE_RRT 0
#define RRTFND_MODE_STAR 1
#define RRTFND_MODE_FN 2

/* random word source */
#define RRTFND_STREAM_WORDS 0  /* explicit uint32 buffer per planner (drop-in: MT19937 words of `random`) */
#define RRTFND_STREAM_PHILOX 1 /* counter-based Philox4x32-10 keyed by (seed, stream_id) */

/* flags */
#define RRTFND_FLAG_EVAL_CHOOSE_PARENT 1   /* evaluate choose_parent_kdtree's return value (src/algorithm.py:822-852) */
#define RRTFND_FLAG_COMMIT_CHOOSE_PARENT 2 /* EXTENSION: use that return value as the new node's edge (the reference drops it, :146/:224) */
#define RRTFND_FLAG_TRACE 4                /* write one rrtfnd_trace_t per successful iteration */
#define RRTFND_FLAG_NODE_GRID 8            /* answer nearest / near-set queries of large trees through batch.node_grid (RRT, RRT* only) */
#define RRTFND_FLAG_OBST_HAS_K 16          /* rrtfnd_plan_host only: obst_host rows are (ox, oy, r, K) instead of (ox, oy, r) */
#define RRTFND_FLAG_RECTANGLES 32          /* the obstacle table holds rectangle rows (extension below): rrtfnd_plan_batch then runs the
                                            * kernel instances that look for them; without the flag its batch kernels treat every row
                                            * as a circle (the dispatch costs 6 % on circle-only maps). The stateless operators and the
                                            * FND operators always look. rrtfnd_plan_host sets it itself. */

/* per-planner status (rrtfnd_planner_t.status) */
#define RRTFND_ST_RUNNING 0        /* iter < iter_num, can be resumed */
#define RRTFND_ST_DONE 1           /* iter == iter_num (or RRT reached the goal) */
#define RRTFND_ST_NEED_WORDS 2     /* word buffer exhausted at an attempt boundary; refill and call again */
#define RRTFND_ST_ATTEMPT_CAP 3    /* max_attempts reached (reference would still be looping, src/algorithm.py:74-75) */
#define RRTFND_ST_CAPACITY (-1)    /* node capacity exhausted */
#define RRTFND_ST_NEAR_CAP (-2)    /* reserved (v1: near-set larger than near_cap); never returned by ABI v2 */
#define RRTFND_ST_FINISH_CAP (-3)  /* more goal-reaching nodes than finish_cap */
#define RRTFND_ST_PATH_CAP (-4)    /* best path longer than path_cap */
#define RRTFND_ST_DUPLICATE (-5)   /* steered point coincides with an existing vertex (src/graph.py:54-56 corner) */
#define RRTFND_ST_NO_CHILDLESS (-6)/* forced_removal would spin forever (src/algorithm.py:815-816) */
#define RRTFND_ST_STREAM_EXHAUSTED (-7) /* word buffer ran out inside forced_removal's redraw loop */
#define RRTFND_ST_REGROW_NEED_WORDS (-8) /* rrtfnd_fnd_regrow: word buffer exhausted at an attempt boundary; the planner state is
                                          * consistent - refill and call again with the remaining iterations */

/* rrtfnd_fnd_regrow gives up with RRTFND_ST_ATTEMPT_CAP after this many attempts per allowed extension
 * (max_iter * RRTFND_REGROW_ATTEMPTS_PER_ITER in total): the reference's loop (src/algorithm.py:490-495) only counts
 * successful extensions and spins forever when no sample can be connected to the root tree (root boxed in or gone). */
#define RRTFND_REGROW_ATTEMPTS_PER_ITER 32

/* library errors (negative return values) */
#define RRTFND_E_BADARG (-100)
#define RRTFND_E_NOGPU (-101)
#define RRTFND_E_SMEM (-102)

typedef struct rrtfnd_params {
    int32_t mode;          /* RRTFND_MODE_* */
    int32_t iter_num;      /* successful iterations to reach (src/algorithm.py:71, :118, :214) */
    int32_t max_nodes;     /* RRT*-FN node cap (src/algorithm.py:233) */
    int32_t capacity;      /* node slots per planner */
    int32_t n_obst;        /* circles in the obstacle table */
    int32_t near_cap;      /* re-parented nodes of one rewire_kdtree call kept in shared memory (more spill to nflags; a performance knob) */
    int32_t finish_cap;    /* max goal-reaching nodes tracked (finish_nodes_of_path, :116) */
    int32_t path_cap;      /* max ids stored for best_path["path"] */
    int32_t stream_mode;   /* RRTFND_STREAM_* */
    int32_t flags;         /* RRTFND_FLAG_* */
    int32_t trace_cap;     /* trace records per planner */
    int32_t reserved;
    int64_t words_per_planner; /* RRTFND_STREAM_WORDS: words available per planner */
    int64_t max_attempts;      /* attempts allowed per planner over its lifetime (<=0: unlimited) */
    double step_length;    /* steer() max_length (src/algorithm.py:731) */
    double radius;         /* near-set / rewire radius (:837, :893) */
    double node_radius;    /* Map.node_radius: inflation of Map.is_occupied_c (src/map.py:45) */
    double bias;           /* Graph.random_node bias (src/graph.py:96) */
    double x_lo, x_hi, y_lo, y_hi; /* Graph.xy_range (src/graph.py:98) */
    double goal_node_radius; /* the planners' own `node_radius` argument: goal test dist < 2 * it (src/algorithm.py:757-758) */
} rrtfnd_params_t;

/* Per-planner scalars. One entry per planner, device resident, read back by the host. */
typedef struct rrtfnd_planner {
    double start_x, start_y;   /* Graph.start (root position) */
    double goal_x, goal_y;     /* Graph.goal */
    double best_cost;          /* best_path["cost"], +inf when none (src/algorithm.py:115) */
    uint64_t seed, stream_id;  /* Philox key / stream */
    uint64_t cursor;           /* next random word index */
    int64_t attempts;          /* new_node() calls (src/algorithm.py:36) */
    int64_t n_edge_checks;     /* unique segments submitted to through_obstacle (:581) */
    int64_t n_edge_checks_ref; /* reference-equivalent count: nearest rounds + 2 x near-set (:847 and :904) */
    int64_t n_circle_tests;    /* intersection_circle evaluations actually performed on the device */
    int64_t n_scan_nodes;      /* nodes visited by nearest / near-set scans (x 16 B = algorithmic bytes) */
    int64_t n_chain_hops;      /* parent-chain hops walked by get_cost / find_path (x 12 B) */
    int32_t n_slots;           /* slots in use (live + tombstones) */
    int32_t n_live;            /* live nodes (n_of_nodes, :209) */
    int32_t last_id;           /* Graph.last_id (src/graph.py:32) */
    int32_t iter;              /* successful iterations so far */
    int32_t root_slot;         /* slot of the tree root (Graph.id_vertex[Graph.start]) */
    int32_t n_finish;          /* len(finish_nodes_of_path) */
    int32_t best_len;          /* len(best_path["path"]) */
    int32_t status;            /* RRTFND_ST_* */
    int32_t n_rewired;         /* total rewired nodes */
    int32_t n_removed;         /* total forced removals */
    int32_t solution_found;    /* src/algorithm.py:114 */
    int32_t n_compactions;
    int32_t fn_count;          /* RRT_star_FN's n_of_nodes (src/algorithm.py:209): 1 + this call's insertions - removals;
                                  NOT the live count when the Graph already held nodes */
    int32_t reserved;
} rrtfnd_planner_t;

/* One record per successful iteration (RRTFND_FLAG_TRACE). */
typedef struct rrtfnd_trace {
    int32_t id_new;     /* id given by Graph.add_vertex (src/graph.py:57-58) */
    int32_t id_near;    /* nearest visible node (src/algorithm.py:44) */
    int32_t n_near;     /* near-set size without the new node itself (:837) */
    int32_t n_rewired;  /* nodes re-parented by rewire_kdtree (:905-910) */
    int32_t removed_id; /* forced_removal result or -1 (:234) */
    int32_t cp_parent;  /* parent id in choose_parent_kdtree's RETURNED edge (:852), -1 if not evaluated */
    double cp_cost;     /* cost in that returned edge */
} rrtfnd_trace_t;

/* Uniform grid over the circular obstacles, used to cull circle tests without changing a single boolean
 * (DESIGN.md "Obstacle grid": a circle is listed in every cell its box, inflated by r + margin, overlaps; the
 * margin bounds the rounding error of the reference's own quadratic for |slope| <= slope_max and
 * |coordinates| <= coord_max; other segments fall back to the full table). nx == 0 disables culling.
 * cell(v) = min(max((int)floor((v - v0) * inv_cell), 0), n - 1) on both the host builder and the device. */
typedef struct rrtfnd_obst_grid {
    double x0, y0, inv_cell;
    double slope_max;           /* segments with |Line.dir| above this (or vertical / NaN) test every circle */
    double coord_max;           /* bound on |coordinate| and radius the margin was derived for */
    double margin;              /* inflation used for seg_* lists (informational) */
    int32_t nx, ny;
    const int32_t* seg_start;   /* [nx*ny + 1] CSR offsets into seg_items (row-major cells, x fastest) */
    const int32_t* seg_items;   /* obstacle indices for through_obstacle (src/algorithm.py:581-591) */
    const int32_t* pt_start;    /* [nx*ny + 1] CSR offsets into pt_items */
    const int32_t* pt_items;    /* obstacle indices for Map.is_occupied_c (src/map.py:38-47), inflated by node_radius */
} rrtfnd_obst_grid_t;

/* Uniform grid over the NODES of each planner (RRTFND_FLAG_NODE_GRID): replaces the k-d tree the reference rebuilds
 * twice per attempt (src/algorithm.py:41-43, :145) for large trees. Every inserted node is pushed onto the singly
 * linked list of its cell; a query gathers the nodes of a window of cells that provably contains every node within
 * the distance it needs and then runs the SAME exact predicates as the full scans on that candidate list, so results
 * are identical to the scans (and to the reference) by construction. Used once a tree has min_nodes nodes.
 * cell(v) as in rrtfnd_obst_grid_t. No removals: not available with RRTFND_MODE_FN or after FND surgery. */
typedef struct rrtfnd_node_grid {
    double x0, y0, inv_cell, cell;
    int32_t nx, ny;
    int32_t min_nodes;          /* trees smaller than this are scanned */
    int32_t cand_cap;           /* candidate slots per planner (a window holding more falls back to the scan) */
    int32_t* head;              /* [P][nx*ny] newest node of each cell, -1 = empty (reset by rrtfnd_init_batch) */
    int32_t* next;              /* [P][capacity] next (older) node of the same cell, -1 = end */
    int32_t* cand;              /* [P][2*cand_cap] scratch */
} rrtfnd_node_grid_t;

/* Walk record of one node (rrtfnd_batch_t.walk). gp is only a hint: it is used when the parent's record confirms it. */
typedef struct rrtfnd_walk {
    double ecost;               /* = ecost[slot] */
    int32_t parent;             /* = parent[slot] */
    int32_t gp;                 /* guess of parent[parent[slot]], -1 = none */
} rrtfnd_walk_t;

/* Device-resident batch of P independent planners, structure-of-arrays, fixed capacity. */
typedef struct rrtfnd_batch {
    int32_t n_planners;
    int32_t reserved;
    rrtfnd_planner_t* planners; /* [P] */
    double* xy;                 /* [P][capacity][2] node position (x, y); NaN in a tombstone slot */
    double* ecost;              /* [P][capacity] edge cost to parent (Graph.cost, src/graph.py:29) */
    int32_t* parent;            /* [P][capacity] parent SLOT, -1 = none */
    int32_t* nchild;            /* [P][capacity] len(Graph.children[id]) */
    int32_t* id;                /* [P][capacity] reference node id, -1 = tombstone; ascending in slot order */
    uint8_t* nflags;            /* [P][capacity] scratch: bit 0 = some goal-reaching node descends from this node */
    int32_t* finish;            /* [P][finish_cap] slots of goal-reaching nodes, in discovery order */
    int32_t* best_path;         /* [P][path_cap] ids leaf -> root (best_path["path"]) */
    const double* obst;         /* [n_obst][4]: ox, oy, r, K = ((-(r*r)) + ox*ox) + oy*oy */
    const uint32_t* words;      /* [P][words_per_planner] or NULL */
    rrtfnd_trace_t* trace;      /* [P][trace_cap] or NULL */
    rrtfnd_obst_grid_t grid;    /* optional culling grid over obst */
    rrtfnd_node_grid_t node_grid; /* optional grid over the nodes (RRTFND_FLAG_NODE_GRID) */
    rrtfnd_walk_t* walk;        /* [P][capacity] scratch of rrtfnd_plan_batch, required with RRTFND_FLAG_NODE_GRID, else unused (may
                                 * be NULL): a 16-byte copy of (ecost, parent) per node plus a self-healing guess of the
                                 * grandparent slot, rebuilt from ecost / parent at the start of every call. The node-grid kernel's
                                 * Graph.get_cost walks (src/graph.py:41-46; hundreds of hops in a large tree) read ONE record per
                                 * hop and, while the guesses hold, fetch two hops per memory round trip. Never read by the host;
                                 * ecost / parent stay the truth. */
} rrtfnd_batch_t;

/* ---- library info ------------------------------------------------------------------------ */
int rrtfnd_abi_version(void);
const char* rrtfnd_status_string(int status);
/* kernel launches one rrtfnd_plan_batch call makes for a batch of n_planners (large batches are grown in slices whose
 * launches overlap, csrc/planner.cu launch_plan) */
int32_t rrtfnd_plan_launches(const rrtfnd_params_t* prm, int32_t n_planners);
/* dynamic shared memory (bytes) per planner the fused planner kernel needs for these parameters; <0 = does not fit */
int64_t rrtfnd_plan_smem_bytes(const rrtfnd_params_t* prm);

/* ---- obstacle grid (host side, no CUDA calls) ------------------------------------------------- */

/* Build the culling grid for obst_host [M][3] (ox, oy, r) over the box [lo_x, hi_x] x [lo_y, hi_y] that
 * bounds every coordinate the planners can produce (xy_range, starts, goals). cell <= 0 picks
 * sqrt(area / M). Two-call pattern: with seg_items_host == NULL only the sizes are returned in
 * n_items_out[0] (seg) and n_items_out[1] (pt) and grid_out->nx/ny are set; otherwise the CSR arrays are
 * filled (seg_start_host / pt_start_host hold nx*ny + 1 entries). Pointers inside grid_out are left NULL:
 * the caller uploads the four arrays and stores the device pointers. */
int rrtfnd_obst_grid_build_host(const double* obst_host, int32_t n_obst, double node_radius, double lo_x,
                                double hi_x, double lo_y, double hi_y, double cell, rrtfnd_obst_grid_t* grid_out,
                                int32_t* seg_start_host, int32_t* seg_items_host, int32_t* pt_start_host,
                                int32_t* pt_items_host, int64_t* n_items_out);
/* The same for rows of the device table's own form [M][4] (ox, oy, r, K), which may hold rectangle rows (below). */
int rrtfnd_obst_grid_build_host4(const double* obst4_host, int32_t n_obst, double node_radius, double lo_x,
                                 double hi_x, double lo_y, double hi_y, double cell, rrtfnd_obst_grid_t* grid_out,
                                 int32_t* seg_start_host, int32_t* seg_items_host, int32_t* pt_start_host,
                                 int32_t* pt_items_host, int64_t* n_items_out);

/* ---- EXTENSION: rectangular obstacles -----------------------------------------------------------------------------
 * The reference declares Map.obstacles_r (src/map.py:18, :54) but never fills or tests it, so this is a builder-defined
 * extension: PARITY UNPINNED BY THE REFERENCE (pinned on oracle/rrt_oracle.c's restatement of the same expressions).
 * Wherever an obstacle table [M][4] is taken, a row whose third column is NEGATIVE is an axis-aligned rectangle
 * (cx, cy, -hx, hy): centre, negated half width (hx > 0), half height. Semantics, by analogy with the circles (segments
 * against the raw shape, points against the shape inflated by node_radius; csrc/arith.cuh):
 *   segment: the closed segment p1-p2 intersects the closed rectangle (separating axes on the segment's midpoint and half
 *            vector; no division, one fp64 rounding per written operation);
 *   point:   its distance to the rectangle is < node_radius (strict), or it lies inside the closed rectangle.
 * Circles and rectangles may be mixed in one table; the culling grid lists both. */

/* ---- stateless batched operators (device pointers) ------------------------------------------ */

/* through_obstacle(Line(p1, p2), obstacles) for n_edges segments (src/algorithm.py:581-591,
 * :555-578, :594-635; src/line.py:5-16). One warp per edge, lanes stride over the obstacle tile in
 * shared memory. out_hit[i] = 1 if segment i collides. */
int rrtfnd_edges_vs_circles(const double* p1x, const double* p1y, const double* p2x, const double* p2y,
                            int64_t n_edges, const double* obst, int32_t n_obst, uint8_t* out_hit, void* stream);

/* Map.is_occupied_c(p) for n points (src/map.py:38-47): any circle with
 * sqrt(dx*dx+dy*dy) < node_radius + r. */
int rrtfnd_points_occupied(const double* px, const double* py, int64_t n_points, const double* obst,
                           int32_t n_obst, double node_radius, uint8_t* out_occ, void* stream);

/* The same two predicates evaluated the way the fused planner does: circles culled through `grid` (host struct
 * holding device pointers). Results are identical to the un-culled operators above by construction; the parity
 * tests compare them on adversarial inputs. */
int rrtfnd_edges_vs_circles_grid(const double* p1x, const double* p1y, const double* p2x, const double* p2y,
                                 int64_t n_edges, const double* obst, int32_t n_obst, const rrtfnd_obst_grid_t* grid,
                                 uint8_t* out_hit, void* stream);
int rrtfnd_points_occupied_grid(const double* px, const double* py, int64_t n_points, const double* obst,
                                int32_t n_obst, const rrtfnd_obst_grid_t* grid, double node_radius, uint8_t* out_occ,
                                void* stream);

/* nearest_node_kdtree (src/algorithm.py:666-699) for one query per planner: the slot of the nearest node
 * (plain d2, lowest id on ties) whose segment Line(q, node) is collision free; -1 if none is visible,
 * -2 if q coincides exactly with a vertex (:676-678). q is [P][2]. Tests every circle (no grid). */
int rrtfnd_nearest_visible(const rrtfnd_batch_t* batch, int32_t capacity, const double* q, int32_t n_obst,
                           int32_t* out_slot, int32_t* out_rounds, void* stream);

/* cKDTree.query_ball_point(q, r) (src/algorithm.py:837, :893): slots with d2 <= r*r in ascending slot
 * order, per planner. out_slots is [P][near_cap]; out_count[p] may exceed near_cap (then truncated). */
int rrtfnd_near_set(const rrtfnd_batch_t* batch, int32_t capacity, const double* q, double radius,
                    int32_t near_cap, int32_t* out_slots, int32_t* out_count, void* stream);

/* Graph.get_cost (src/graph.py:34-46) for n (planner, slot) pairs: leaf -> root sequential fp64 sum. */
int rrtfnd_chain_cost(const rrtfnd_batch_t* batch, int32_t capacity, const int32_t* planner_idx,
                      const int32_t* slot, int64_t n, double* out_cost, void* stream);

/* `x ** 2` (which = 0) or `x ** 0.5` (which = 1) for n values as the reference's interpreter evaluates them: libm pow() of
 * glibc 2.39 x86-64 (FMA build), which is not fl(x*x) / sqrt(x) in ~0.1 % of arguments (src/algorithm.py:602-603, :617,
 * :619-620; src/map.py:44; csrc/libm_pow.h). The planners use x*x / sqrt (no predicate changed in 4.2 G audited
 * evaluations); this operator is the device half of closing that gap by construction. */
int rrtfnd_pow_libm(const double* x, int64_t n, int32_t which, double* out, void* stream);

/* ---- fused planner driver ---------------------------------------------------------------------- */

/* Advance every planner of the batch until iter == prm->iter_num (or a status other than RUNNING).
 * Replaces the loops of RRT / RRT_star / RRT_star_FN (src/algorithm.py:71-94, :118-186, :214-265) with
 * new_node (:36-51), choose_parent_kdtree (:822-852), rewire_kdtree (:882-910), forced_removal (:800-819),
 * check_solution (:749-760), find_path (:778-797) fused into one persistent warp per planner. Resumable. */
int rrtfnd_plan_batch(const rrtfnd_params_t* prm, const rrtfnd_batch_t* batch, void* stream);

/* Initialise planner p's tree to the single root node `start` (Graph.__init__, src/graph.py:25-32) and
 * zero its counters. planners[p].start/goal/seed/stream_id must already be set. */
int rrtfnd_init_batch(const rrtfnd_params_t* prm, const rrtfnd_batch_t* batch, void* stream);

/* ---- RRT*-FND tree surgery (src/algorithm.py:294-434), one warp per planner ----------------------------------
 * Node ids at the interface, like the reference's arguments. `path` is the planner's stored best path
 * (batch.best_path, ids goal side first); obstacles are batch.obst / prm->n_obst / prm->node_radius as they are
 * at the time of the call (the culling grid is not used: the map has usually just changed). */

/* select_branch + remove_children (src/algorithm.py:294-341): re-root planner p at node current_id[p] of its path
 * and delete everything outside that node's subtree; start <- its position, path <- path[:index + 1], best cost
 * re-summed to the new root. out_status[p] (may be NULL): 0, -1 = not on the path / not alive (ValueError in the
 * reference), -2 = it is already the root (KeyError in the reference). */
int rrtfnd_fnd_select_branch(const rrtfnd_params_t* prm, const rrtfnd_batch_t* batch, const int32_t* current_id,
                             int32_t* out_status, void* stream);
/* valid_path (src/algorithm.py:344-367): delete the path nodes occupied under Map.is_occupied_c together with the
 * off-path subtrees hanging from them; the on-path child of each becomes a root. out_separate_root_id[p] = id of the
 * last such child (goal side), previous_root_id[p] if no path node is occupied, -1 with out_index_error[p] = 1 when
 * the goal-side leaf itself is occupied (the reference raises IndexError at :358 after removing its children). */
int rrtfnd_fnd_valid_path(const rrtfnd_params_t* prm, const rrtfnd_batch_t* batch, const int32_t* previous_root_id,
                          int32_t* out_separate_root_id, int32_t* out_index_error, void* stream);
/* reconnect + get_near_nodes + get_distance_dict + check_for_tree_associativity (src/algorithm.py:370-434, :938-951):
 * link the separate root to the first node in ascending id that belongs to the root tree, lies within `radius`
 * (get_distance_dict's formula) and sees it (through_obstacle(Line(node, separate root)) is False); edge cost =
 * calc_distance. On success the stored best path / cost become find_path(last_node = path[0], root).
 * out_linked_id[p] = id linked to, or -1. */
int rrtfnd_fnd_reconnect(const rrtfnd_params_t* prm, const rrtfnd_batch_t* batch, const int32_t* separate_root_id,
                         double radius, int32_t* out_linked_id, void* stream);

/* regrow + nearest_node + reconnect_ancestors (src/algorithm.py:482-543, :702-728): up to max_iter (500 in the reference)
 * plain extensions of the root tree from the planner's word stream (planner cursor), each followed by an attempt to hang
 * the separate tree - re-rooted at the first of its nodes, in ascending id, that the new node sees closer than
 * step_length - under the new node. Uses the culling grid of the current obstacle table. out_reconnected[p] = 1 / 0,
 * RRTFND_ST_ATTEMPT_CAP (3) after max_iter * RRTFND_REGROW_ATTEMPTS_PER_ITER attempts, RRTFND_E_BADARG for a planner
 * whose separate_root_id is not a live node (-1 = skip this planner), or a negative status; out_iters[p] (may be NULL) =
 * extensions made. On success the stored best path / cost become
 * find_path(path[0], root). */
int rrtfnd_fnd_regrow(const rrtfnd_params_t* prm, const rrtfnd_batch_t* batch, const int32_t* separate_root_id,
                      int32_t max_iter, int32_t* out_reconnected, int32_t* out_iters, void* stream);

/* EXTENSION (DetectCollision on path EDGES; the reference's valid_path tests the path nodes only, src/algorithm.py:356): after
 * rrtfnd_fnd_valid_path, cut every still-linked pair (child = path[i], parent = path[i + 1]) of the stored path whose segment
 * through_obstacle(Line(parent, child)) (:581-591) is blocked - the child becomes a separate root, nothing is deleted.
 * separate_root_id[p] = valid_path's result (negative = skip the planner); out_separate_root_id[p] = the root of the tree
 * holding the goal-side leaf if anything was cut, else the input (may alias it); out_n_cut[p] (may be NULL) = edges cut.
 * Parity unpinned by the reference; pinned on the oracle's composition of the reference's own through_obstacle. */
int rrtfnd_fnd_cut_blocked_edges(const rrtfnd_params_t* prm, const rrtfnd_batch_t* batch, const int32_t* separate_root_id,
                                 int32_t* out_separate_root_id, int32_t* out_n_cut, void* stream);

/* ---- the replanning tick (BASELINE.json config 5) as ONE call without device -> host traffic --------------------------------
 * The reference stops at the init_movement() stub (src/algorithm.py:283-291); its usage example test_select_branch (:437-479)
 * implies the loop: step to the next path node (select_branch), validate the path against the moved obstacles (valid_path),
 * mend what was split off (reconnect, else regrow). rrtfnd_fnd_tick runs that for every planner: per-planner state and the
 * tick's outcome live in device arrays (all int32 [P] unless noted), the bookkeeping between the operators is a handful of
 * one-thread-per-planner kernels, so a tick is a sequence of launches on `stream` and nothing is copied to the host. The
 * caller moves the obstacles first (batch->obst / batch->grid, e.g. rrtfnd_obst_grid_build_device). */
typedef struct rrtfnd_fnd_loop {
    int32_t max_rounds;     /* separate trees mended per tick (one per occupied stretch of the path), goal side first */
    int32_t reserved;
    int32_t* phase;         /* in/out: 0 MOVING, 1 ARRIVED (one path node left), 2 LOST */
    int32_t* current;       /* out: node the robot stepped to (path[-2]), -1 = did not move */
    int32_t* status;        /* out: select_branch status */
    int32_t* prev_root;     /* scratch: valid_path argument */
    int32_t* separate;      /* out: valid_path result (after the edge cut, if enabled) */
    int32_t* index_error;   /* out: valid_path's IndexError flag */
    int32_t* edge_cuts;     /* out: edges cut by the edge invalidation (may be NULL) */
    int32_t* todo;          /* scratch: separate root to mend in the current round */
    int32_t* regrow_arg;    /* scratch */
    int32_t* link_tmp;      /* scratch */
    int32_t* regrow_tmp;    /* scratch */
    int32_t* iters_tmp;     /* scratch */
    int32_t* linked;        /* out [P][max_rounds]: id each separate root was linked to by reconnect, -1 */
    int32_t* regrow;        /* out [P][max_rounds]: regrow result (1 reconnected, 0 not, 3 attempt cap, < 0 status), 0 = not run */
    int32_t* regrow_iters;  /* out [P][max_rounds]: extensions regrow made */
    int64_t* stats;         /* in/out [16], accumulated: moves, splits, reconnected, regrown, regrow extensions, then (last tick
                             * only meaningful if zeroed before it) lost, arrived; ticks; ... */
} rrtfnd_fnd_loop_t;
int rrtfnd_fnd_tick(const rrtfnd_params_t* prm, const rrtfnd_batch_t* batch, const rrtfnd_fnd_loop_t* loop,
                    double reconnect_radius, int32_t regrow_iters, int32_t edge_invalidation, void* stream);

/* The culling grid of rrtfnd_obst_grid_build_host4, built by a kernel from the DEVICE table obst [M][4] into the caller's
 * device arrays (grid->seg_start / pt_start hold cells_cap entries, seg_items / pt_items seg_cap / pt_cap). The grid
 * geometry depends on the box, M and `cell` only; coord_max must bound every |coordinate| and radius the table will hold
 * (it sets the exactness margin; larger is still exact). *overflow (device int, zeroed by the caller, never cleared here) is set
 * to 1 if a list did not fit. */
int rrtfnd_obst_grid_build_device(const double* obst, int32_t n_obst, double node_radius, double lo_x, double hi_x,
                                  double lo_y, double hi_y, double cell, double coord_max, rrtfnd_obst_grid_t* grid,
                                  int32_t seg_cap, int32_t pt_cap, int32_t cells_cap, int32_t* overflow, void* stream);

/* ---- host-buffer entry point (what the Python drop-in and bench `e2e` call) -------------------- */

/* Plan P independent problems given HOST buffers; allocates device memory, copies in, runs to completion,
 * copies results back. starts_goals_host is [P][4] (sx, sy, gx, gy); seeds_host is [P][2] (seed, stream_id);
 * obst_host is [M][3] (ox, oy, r): the per-circle constant K = -r ** 2 + ox ** 2 + oy ** 2 (src/algorithm.py:619) is
 * then evaluated inside with the C library's pow(), the call the reference's `**` makes on floats; with
 * RRTFND_FLAG_OBST_HAS_K obst_host is [M][4] (ox, oy, r, K) and the caller's K is used (integer-valued obstacles of
 * Map.generate_obstacles, src/map.py:59-60, are squared exactly by the interpreter). Outputs (all host, may be NULL to
 * skip): planners_host [P]; tree arrays xy [P][capacity][2], ecost, parent (slot), nchild, id [P][capacity];
 * best_path_host [P][path_cap]. The obstacle grid is built inside the call. The device memory is an arena the library keeps
 * between calls (grow-only; calls are serialised); rrtfnd_host_release() frees it. */
int rrtfnd_plan_host(const rrtfnd_params_t* prm, int32_t n_planners, const double* starts_goals_host,
                     const uint64_t* seeds_host, const double* obst_host, const uint32_t* words_host,
                     rrtfnd_planner_t* planners_host, double* xy_host, double* ecost_host,
                     int32_t* parent_host, int32_t* nchild_host, int32_t* id_host, int32_t* best_path_host,
                     int32_t device);

int rrtfnd_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* RRTFND_H */
