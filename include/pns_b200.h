/*
 * pns_b200.h -- C-ABI of the B200-native PedNStream Link-Transmission-Model timestep.
 *
 * The reference (WaimenMak/PedNStream) has no FFI layer: its hot path is the Python method
 * `Network.network_loading(t)` (src/LTM/network.py:266-287) and the per-object methods it calls.
 * Each entry point below replaces one slice of that call tree; the reference line ranges are
 * cited per function.  Conventions:
 *
 *   - plain C: POD structs of raw *device* pointers and sizes, no torch / C++ types;
 *   - the library owns no memory: every buffer is allocated by the caller (the Python host uses
 *     torch tensors) and must stay alive while a launch that uses it is in flight;
 *   - every launch is asynchronous on the `stream` argument (a cudaStream_t passed as void*);
 *   - return value 0 = launched, non-zero = error, message via pns_last_error() (thread local);
 *     physics-level faults that the reference raises as Python exceptions (negative flows,
 *     out-of-range history index) set bits in `pns_state.err[replica]` instead.
 *
 * Memory layout ("time-major structure of arrays", replica index fastest):
 *   hist64[f][t][c*R + r]   f < n_f64, t <= sim_steps, c < n_cols64      (double)
 *   hist32[f][t][l*R + r]   f < 6,                      l < n_links       (float)
 * Column c < n_links is physical link c (network.links order; the reverse direction of link l
 * is l^1); columns >= n_links are the virtual origin/destination links.
 */
#ifndef PNS_B200_H
#define PNS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNS_ABI_VERSION 8
#define PNS_MAX_DEGREE 8 /* link slots per node handled by the node kernel */

/* fp64 history fields (reference src/LTM/link.py:12-17, 56, 425) */
enum {
    PNS_F64_INFLOW = 0,
    PNS_F64_OUTFLOW = 1,
    PNS_F64_CUM_INFLOW = 2,
    PNS_F64_CUM_OUTFLOW = 3,
    PNS_F64_SENDING = 4,
    PNS_F64_RECEIVING = 5,
    PNS_F64_BACK_GATE = 6,
    PNS_F64_SEP_WIDTH = 7 /* present only when n_f64 == 8 */
};
/* fp32 history fields (reference src/LTM/link.py:82-97) */
enum {
    PNS_F32_NUM_PED = 0,
    PNS_F32_DENSITY = 1,
    PNS_F32_SPEED = 2,
    PNS_F32_TRAVEL_TIME = 3,
    PNS_F32_AVG_TRAVEL_TIME = 4,
    PNS_F32_LINK_FLOW = 5
};

/* where the in-step random draws come from (SURVEY.md "RNG ledger", sites R1..R4) */
enum {
    PNS_RNG_TABLE = 0,   /* outcomes supplied by the caller (replay / numpy-compatible stepping) */
    PNS_RNG_PHILOX = 1,  /* Philox4x32-10 keyed (seed, replica; t, link, site) on the device */
    PNS_RNG_REQUEST = 2  /* write the draw requests of this step, change no state (pass 1 of
                            numpy-compatible stepping) */
};

/* bits of pns_state.err[replica] */
enum {
    PNS_ERR_NEG_SENDING = 1,     /* link.py:346,366 ValueError */
    PNS_ERR_NEG_NODE_FLOW = 2,   /* node.py:194,219,238 Warning */
    PNS_ERR_HISTORY_INDEX = 4,   /* numpy IndexError in link.py:210-212 */
    PNS_ERR_ZERO_LAG = 8,        /* tau == 0: the reference result depends on node visiting order */
    PNS_ERR_LP_FAILED = 16       /* 'optimal' node model: the simplex ended unbounded or at its pivot limit (the
                                    reference keeps the previous step's q when linprog fails, node.py:265) */
};

/* One parameter class of links (reference src/LTM/link.py:52-100): networks have few distinct
 * parameter sets, so links carry a class index and the table stays L1-resident.  Every derived
 * constant is evaluated on the host with the reference's own expression order and stored in the
 * precision the reference uses it in (numpy demotes Python scalars to float32 next to a float32). */
typedef struct pns_link_class {
    double length, area, space; /* area = length*width (link.py:131), space = k_jam*area (link.py:386) */
    double kc, vf, kj, act, sigma;
    float kc32, kj32, kj_minus_kc32, gamma32, bi32, area32, length32;
    float yp_coef32;   /* (kc*vf)/(kj-kc)   functions.py:123 */
    float neg_vf32;    /* -vf               functions.py:118 */
    float vf32;        /* vf                functions.py:126 */
    float sm_gamma32;  /* vf*kc             functions.py:128 */
    float inv_kj32;    /* 1/kj              functions.py:128 */
    float max_tt32;    /* length/0.05       link.py:63 */
    float tt0;         /* travel_time[0]    link.py:83 */
    int32_t fftau, swtau; /* free-flow lag (link.py:86), shock-wave lag (link.py:380) */
    int32_t flags;        /* bit0 separator, bits1-2 fd type (0 yperman 1 greenshields 2 smulders) */
    int32_t pad_;
} pns_link_class;

/* Immutable network description.  All pointers are device pointers. */
typedef struct pns_net {
    int32_t abi_version;
    int32_t n_links, n_nodes, n_cols64, sim_steps, replicas, window, n_edges, n_od, n_demand_rows;
    int32_t n_routed, n_groups, n_opts, n_rows, n_terms, n_classes, max_degree, nd_stride; /* nd_stride: 4 or 8 */
    double unit_time;
    const pns_link_class *classes; /* [n_classes] */
    pns_link_class class0;         /* host copy of classes[0]: single-class networks read parameters from the
                                      kernel-parameter (constant) bank instead of loading the table */
    const int32_t *lk_class;       /* [n_links], or [n_links*replicas] with per_replica_scenario */
    const double *lk_width;        /* [n_links] corridor width `_width` (link.py:53): initial gate / lane widths */
    /* node table -- CSR over link slots (virtual link first, then neighbours by ascending id).
     * The outgoing link of a slot is the reverse of its incoming link: out column = in column ^ 1
     * (physical pairs are adjacent, virtual in/out links are allocated as adjacent columns). */
    const int32_t *nd_meta;     /* [n_nodes][4]: slot offset (host use), m | kind<<8 | tf_mode<<16, demand row
                                   (-1 none; the node's virtual in/out links are columns n_links + 2*row, +1),
                                   offset of the node's m(m-1) turning fractions.
                                   kind: 0 one-to-one (node.py:230), 1 regular/classic (node.py:272), 2 regular/optimal
                                   (node.py:249-271, the linear program; such nodes are also listed in lp_nodes);
                                   tf_mode: 0 uniform 1/(m-1) (network.py:269-271), 1 tf_static, 2 routed */
    const int32_t *nd_routed;   /* [n_nodes] index into the routed-node arrays, -1 = none */
    const int32_t *lk_slots;    /* [n_links][2] node-major exchange slots of each link: where its sending flow goes
                                   (end node * nd_stride + its slot there) and where its receiving flow goes
                                   (start node * nd_stride + its slot there) */
    const int32_t *nd_in_link;  /* [n_nodes*nd_stride] the inverse map: history column of the incoming link of each
                                   slot (virtual O/D link: n_links + 2*demand row), -1 = unused slot.  The node pass
                                   stores outflow[t] of that column and inflow[t] of column ^ 1 (node.py:146-162) */
    /* route plan (path_finder.py:510-546 structures, flattened by PathFinder.export_route_plan) */
    const int32_t *rt_routed_nodes, *rt_routed_edge0, *rt_routed_row0;
    const int32_t *rt_row_routed;     /* [n_rows] routed node (index into rt_routed_*) of each upstream-slot row */
    const int32_t *rt_row_grp_ptr, *rt_row_grp; /* [n_rows+1], [n_groups]: the (od, upstream) groups registered at each row */
    const int32_t *rt_term_od;        /* [n_terms] OD column of each accumulation term (= rt_row_od[rt_term_row_entry]) */
    const int32_t *rt_dyn_rows;       /* [n_dyn_rows] rows whose fractions can change between steps (a group with several
                                         options, or several registered ODs); the other rows evaluate to constants and
                                         are computed only on a step with pns_step_io.route_all_rows set */
    const int32_t *rt_grp_node, *rt_grp_up, *rt_grp_od, *rt_grp_has_virtual, *rt_opt_ptr;
    const int32_t *rt_opt_link, *rt_opt_slot;
    const double *rt_opt_dist;
    const int32_t *rt_row_ptr, *rt_row_od, *rt_term_ptr, *rt_term_opt, *rt_term_row_entry;
    double rt_temp, rt_alpha, rt_beta, rt_omega, rt_eps; /* path_finder.py:158-163 */
    /* optional launch order of the single-replica link kernel: CTA i processes the block of
     * lane_order_block consecutive links number lane_order[i] (a permutation of the blocks; NULL = index
     * order).  A schedule, not a result: blocks whose links sit next to demand origins -- where queues form and
     * the blockers draw of a jammed link is a long serial walk -- go first, so that their latency is covered by
     * the rest of the grid instead of extending its tail.  Ignored unless lane_order_block equals
     * pns_lane_block_size(). */
    const int32_t *lane_order;
    int32_t lane_order_block, n_lane_blocks;
    /* per-replica scenarios (domain randomisation, env_loader.py:160-424): when set, lk_class is
     * [n_links*replicas] (replica fastest) -- every replica has its own parameter class per link -- and
     * pns_step_io.od_w is [sim_steps+1][n_od*replicas] (replica fastest).  Topology, widths, OD nodes and the
     * route plan are shared by all replicas. */
    int32_t per_replica_scenario, n_dyn_rows;
    /* assign_flows_type 'optimal' (node.py:249-271): the nodes of kind 2, the largest slot count among them and the
     * weight of the turning-fraction penalty (Node.w = 0.01, node.py:14) */
    const int32_t *lp_nodes;   /* [n_lp_nodes] */
    int32_t n_lp_nodes, lp_max_m;
    double lp_w;
    /* optional schedule of the batched node kernel when nd_stride is 8: all nodes, those with at most 4 link slots
     * first (n_nodes_small of them; they share a CTA in twos).  NULL = one CTA per node in index order.  A schedule,
     * not a result. */
    const int32_t *nd_cols_order;
    int32_t n_nodes_small, pad3_;
} pns_net;

/* Mutable simulation state; all device pointers, caller-owned. */
typedef struct pns_state {
    double *hist64;      /* [n_f64][sim_steps+1][n_cols64*R] */
    float *hist32;       /* [6][sim_steps+1][n_links*R] */
    double *gate;        /* [n_links*R] back gate width of plain links / separator width of separators.
                            The other widths follow from the reference's setters (link.py:110-126, 462-478):
                            front_gate(l) = gate[l^1] for a plain link, = gate[l] for a separator;
                            a separator's back gate and lane width are both gate[l]. */
    int32_t *sep_np64;   /* [n_links*R] separator width was set from a numpy float64 (dtype ledger) */
    float *runsum;       /* [n_links*R] fp32 running sum of travel times (link.py:84,183-186) */
    double *tf_static;   /* [n_edges] host-owned turning fractions (uniform default / user supplied) */
    double *tf_routed;   /* [n_edges*R] fractions computed this step for routed nodes */
    double *probs;       /* [n_opts*R] scratch: P(down | up, od) of this step */
    double *nm_s, *nm_r;   /* [n_nodes*nd_stride*R] node-major sending / receiving flows of the step (link -> node) */
    int32_t *err;        /* [R] error bits */
    int32_t n_f64;       /* 7, or 8 when the network has separators */
} pns_state;

/* Per-call inputs of one or more consecutive steps. */
typedef struct pns_step_io {
    const double *demand;   /* [>= t][n_demand_rows*R]: row t-1 is node.demand[t-1] (node.py:176) */
    const double *od_w;     /* [sim_steps+1][n_od]: od_flows[(o,d)][t] (od_manager.py:52-54); [..][n_od*replicas] with
                               pns_net.per_replica_scenario */
    const int32_t *draw_b;  /* TABLE: [rows][3][n_links*R] outcomes of sites R1, R2, R3 */
    const double *draw_n;   /* TABLE: [rows][n_links*R] speed noise (site R4) */
    int64_t draw_row_stride;/* rows advance by one per step when != 0 (0: one row reused) */
    /* REQUEST outputs, [n_links*R] each */
    int32_t *req_kind;      /* 0 no R1 draw and flow fixed, 1 diffusion branch (flow in req_sval), 2 binomial */
    int32_t *req_n1;        /* trials of R1 */
    float *req_rf;          /* releasing factor: p = 0.7 + 0.15*rf**0.8 is evaluated by the host in numpy */
    double *req_sval;       /* sending flow after the release stage when kind != 2 */
    int32_t *req_n3;        /* trials of R3 (-1: separator, no draw) */
    double *req_exp;        /* REQUEST output [n_opts*R]: the arguments -temp*U of the route-choice exponentials of the
                               step (path_finder.py:585) -- numpy-compatible stepping evaluates them with the host's
                               numpy, whose exp differs from CUDA's in the last bit of a few percent of arguments */
    const double *draw_exp; /* TABLE input [rows][n_opts*R]: exp of those arguments (NULL: evaluated on the device) */
    uint64_t seed;          /* PHILOX */
    uint32_t replica_base;  /* PHILOX: global index of local replica 0 (replicas sharded over GPUs) */
    uint32_t route_all_rows;/* non-zero: the first step of the call evaluates every route row, not only rt_dyn_rows (the
                               first step after pns_state_init, a change of the route plan; step 1 always does) */
    double *lp_x;           /* optional output [n_edges*R]: the turn flows x of the step's linear programs before the
                               floor (nodes of kind 2; indexed like tf_routed); NULL = not kept */
} pns_step_io;

/* Control environment over R replicas (reference rl/pz_pednet_env.py, rl/builders.py, rl/discovery.py).
 * Actions and observations are replica-major: actions[r][n_act], obs[r][n_obs], reward[r]. */
enum { /* observation sources (rl/builders.py:119-177) */
    PNS_OBS_INFLOW = 0, PNS_OBS_OUTFLOW = 1, PNS_OBS_REV_INFLOW = 2, PNS_OBS_REV_OUTFLOW = 3,
    PNS_OBS_SHARED_DENSITY = 4,        /* link.get_density(t) */
    PNS_OBS_SHARED_DENSITY_OVER_KJ = 5,/* link.get_density(t) / k_jam */
    PNS_OBS_SPEED = 6, PNS_OBS_GATE = 7 /* link.back_gate_width */
};
typedef struct pns_env {
    int32_t n_act, n_obs, n_reward_links, pad_;
    const int32_t *act_link;   /* [n_act] controlled link (a separator agent lists its forward link) */
    const int32_t *act_sep;    /* [n_act] 1 = separator action (also sets the reverse lane to W - v) */
    const double *act_lo, *act_hi;     /* [n_act] clip bounds: gate [0, width]; separator [w_min, W - w_min] */
    const double *act_max_delta;       /* [n_act] per-env-step rate limit (rl/builders.py:281-311) */
    const double *act_total_width;     /* [n_act] corridor width W (separator actions) */
    const int32_t *obs_link;   /* [n_obs] */
    const int32_t *obs_src;    /* [n_obs] PNS_OBS_* */
    const float *obs_div;      /* [n_obs] fixed normalisation divisor (1 = none, rl/builders.py:179-238) */
    const int32_t *reward_link;/* [n_reward_links] controlled links of the first agent when it is a gate agent
                                  (the reference rewards only that agent, pz_pednet_env.py:548-581) */
    /* The same programs indexed by directed link, for the batched link kernel of pns_env_step, which applies the
     * actions inside the link pass and emits observations and the reward from the state update (no separate
     * environment launches).  Optional as a set: with lk_act == NULL pns_env_step runs the stand-alone kernels. */
    const int32_t *lk_act;     /* [n_links] index of the action that sets this link's width, -1 = none */
    const int32_t *lk_obs_ptr; /* [n_links+1] observation entries produced by each directed link ... */
    const int32_t *lk_obs_col; /* ... their column in obs[r][.] ... */
    const int32_t *lk_obs_src; /* ... PNS_OBS_* (REV_* entries are listed under the reverse link as INFLOW/OUTFLOW) ... */
    const float *lk_obs_div;   /* ... and divisor */
    const int32_t *lk_reward;  /* [n_links] 1 = the link or its reverse is a reward link */
    int32_t *reward_count;     /* [R] scratch, zero before the first step (the kernel leaves it zero) */
} pns_env;

int pns_abi_version(void);
const char *pns_last_error(void);

/* Initial state of every history field, width table and running sum
 * (reference Link.__init__/Separator.__init__, link.py:12-17, 32-100, 420-425).
 * `gate` must already hold the initial widths. */
int pns_state_init(const pns_net *net, const pns_state *st, void *stream);

/* Link demand/supply pass for step t (time index tau = t-1), one thread per link pair and replica:
 * Link.cal_sending_flow (link.py:216-370) incl. get_outflow (:199-214) and
 * Link/Separator.cal_receiving_flow_with_reverse (:372-416, :480-512).
 * Writes sending_flow[tau], receiving_flow[tau] and the node-major copies nm_s / nm_r the node pass reads
 * (or only the draw requests in REQUEST mode; with io->req_exp set and a routed network that call also runs
 * pns_route_fractions in REQUEST mode, so one call collects everything the host must supply for the step).  The single-replica kernel does not rewrite an nm_s slot whose
 * previous and new sending flow are both 0: a caller that repeats steps clears nm_s first (see pns_node_flows).
 * Inside pns_step this pass for step t+1 is fused with pns_link_update of step t (same thread,
 * state kept in registers). */
int pns_link_flows(const pns_net *net, const pns_state *st, const pns_step_io *io, int t, int rng_mode,
                   void *stream);

/* Route choice for step t, one thread per (upstream slot of a routed node, replica): the logit P(down | up, od) of
 * the (od, upstream) groups registered at the slot (PathFinder.update_node_turn_probs, path_finder.py:561-589; writes
 * st->probs), then the OD mixing and row check of that row of turning fractions (update_turning_fractions + check_fractions,
 * path_finder.py:591-715; writes st->tf_routed, which the node pass reads).  Runs for every routed node on every step,
 * like the reference (network.py:273-278).  rng_mode REQUEST: writes only the exponentials' arguments to io->req_exp;
 * TABLE with io->draw_exp: takes the exponentials from that table; otherwise they are evaluated on the device
 * (libdevice exp, or the operation-exact det_exp in PHILOX mode). */
int pns_route_fractions(const pns_net *net, const pns_state *st, const pns_step_io *io, int t, int rng_mode,
                        void *stream);

/* Node pass for step t, one thread per node and replica: Node.assign_flows / solve / update_links
 * (node.py:146-300) with the turning fractions pns_route_fractions left in st->tf_routed (routed nodes), the
 * host-owned st->tf_static, or uniform 1/(m-1).
 * Writes inflow[t] / outflow[t] of every link (physical and virtual) and the cumulative counts of the virtual
 * O/D links at row t.
 * Row contract: a node that receives no sending flow stores nothing, so rows inflow[t] and outflow[t] must be zero
 * when the step starts.  pns_state_init leaves every row zero and each step writes only its own row, so a run that
 * moves forward in time satisfies this by construction; a caller that repeats a step clears those two rows first
 * (the Python engine does so whenever it is asked to go back in time). */
int pns_node_flows(const pns_net *net, const pns_state *st, const pns_step_io *io, int t, void *stream);

/* Link state update for step t, one thread per link pair and replica: cumulative counts from the
 * inflow/outflow the node pass stored (link.py:19-25),
 * Link.update_link_density_flow + update_speeds (link.py:133-188, 430-452) and
 * BiDirectionalFd.__call__ (src/utils/functions.py:112-134). */
int pns_link_update(const pns_net *net, const pns_state *st, const pns_step_io *io, int t, int rng_mode,
                    void *stream);

/* Network.network_loading(t) for t = t0 .. t0+n_steps-1 (network.py:266-287): the four passes
 * above per step, back to back on `stream` (TABLE or PHILOX mode). */
int pns_step(const pns_net *net, const pns_state *st, const pns_step_io *io, int t0, int n_steps,
             int rng_mode, void *stream);

/* pns_step with CUDA-event timing of every launch on `stream` (bench/roofline instrumentation):
 * adds the elapsed milliseconds of each pass to ms[0..2] (link kernel, route_probs, node_flows;
 * ms[3] is unused) and the launch counts to launches[0..2].  Synchronises the stream before returning. */
int pns_step_profiled(const pns_net *net, const pns_state *st, const pns_step_io *io, int t0, int n_steps,
                      int rng_mode, void *stream, double *ms, int64_t *launches);

/* pns_step with the host traffic of a driving loop folded in (reference callers:
 * `for t in range(1, steps): net.network_loading(t)` with demand edited between steps, examples/long_corridor.py:65-66,
 * 126-134): before the node pass of every step its demand row is copied from pinned host memory
 * (`host_demand`, same layout as pns_step_io.demand) into the device table, and after every step the network-wide
 * pedestrian count of that step is reduced on the device to PNS_METRIC_SLOTS partial sums and copied to the host.
 * `dev_metric` (device scratch) and `host_metric` (pinned) hold [n_steps][PNS_METRIC_SLOTS * PNS_METRIC_STRIDE]
 * doubles; the count of step t0+k is the sum over j of host_metric[k][j * PNS_METRIC_STRIDE] (exact in any
 * order whenever the counts are integer-valued, i.e. unless a gate capacity makes a sending flow fractional).  The copies (one H2D and one D2H per step) run on an internal
 * second stream that meets `stream` once per group of 8 steps: demand rows are copied one group ahead of their step,
 * results up to one group after it.  The call returns with everything enqueued; after synchronising `stream` all
 * results are on the host. */
#define PNS_METRIC_SLOTS 64   /* partial sums per step (a power of two) */
#define PNS_METRIC_STRIDE 4   /* doubles between slots: one 32-byte sector each */
int pns_step_streamed(const pns_net *net, const pns_state *st, const pns_step_io *io, int t0, int n_steps,
                      int rng_mode, const double *host_demand, double *dev_metric, double *host_metric, void *stream);

/* ActionApplier.apply_all_actions (rl/builders.py:264-352): rate-limit, clip, write the gate table. */
int pns_env_apply_actions(const pns_net *net, const pns_state *st, const pns_env *env, const float *actions,
                          void *stream);

/* ObservationBuilder.build_observation for every agent (rl/builders.py:68-177) and
 * PedNetParallelEnv._compute_rewards (rl/pz_pednet_env.py:548-581) at simulation row t. */
int pns_env_observe(const pns_net *net, const pns_state *st, const pns_env *env, int t, float *obs,
                    float *reward, void *stream);

/* One environment step (rl/pz_pednet_env.py:194-254 with action_gap 1) in one call: apply the actions (skipped when
 * `actions` is NULL), Network.network_loading(t), build observations and the reward; `cum_reward` (optional, [R])
 * accumulates the rewards.  Same results as pns_env_apply_actions + pns_step + pns_env_observe.
 * Batched replicas with the per-link programs of pns_env run three launches: actions ride in the flow launch,
 * observations and reward in the state-update launch. */
int pns_env_step(const pns_net *net, const pns_state *st, const pns_step_io *io, const pns_env *env,
                 const float *actions, int t, int rng_mode, float *obs, float *reward, float *cum_reward,
                 void *stream);

/* n_steps environment steps t0 .. t0+n_steps-1 in one call (a rollout with given actions, e.g. random gate actions).
 * Device-resident form (host_* NULL): actions [n_steps][R*n_act], obs [n_steps][R*n_obs], reward [n_steps][R] on the
 * device.  Host form: host_actions / host_obs / host_reward are pinned host arrays of those shapes and actions / obs /
 * reward are device staging buffers for TWO steps ([2][...]); every step has its own host->device copy of its actions
 * and device->host copy of its observations and rewards, on two internal copy streams, double-buffered against the
 * step kernels.  Returns with everything enqueued; after synchronising `stream` the results are complete. */
int pns_env_rollout(const pns_net *net, const pns_state *st, const pns_step_io *io, const pns_env *env, int t0,
                    int n_steps, int rng_mode, const float *actions, float *obs, float *reward, float *cum_reward,
                    const float *host_actions, float *host_obs, float *host_reward, void *stream);

/* Origin demand of a batch of replicas, drawn on the device (reference od_manager.py:100-155): demand[t][row*R + r]
 * for t = 0..sim_steps.  pattern[row*R + r]: 0 gaussian_peaks = Poisson(base + peak*bump1[t] + peak*bump2[t]),
 * 1 constant = base, 2 sudden_demand = gaussian_peaks plus a burst of 10..19 steps at a random start with height
 * 20..49, -1 none.  Every value is a pure function of (seed, replica_base + r; t, row): independent of the sharding
 * of replicas over GPUs.  bump1/bump2: [sim_steps] (host-evaluated gaussian bumps); base/peak/pattern: [rows*R]. */
int pns_env_draw_demand(int sim_steps, int rows, int replicas, uint32_t replica_base, uint64_t seed,
                        const double *bump1, const double *bump2, const double *base, const double *peak,
                        const int32_t *pattern, double *demand, void *stream);

/* Domain randomisation on the device: per-replica scenarios of one episode without a host loop over replicas
 * (reference NetworkEnvGenerator.generate_random_link_params / generate_random_od_flows / generate_random_demand_params,
 * env_loader.py:183-258, 363-424; same distributions, counter-based draws keyed (seed; replica_base + r)):
 *   classes      [n_base_classes + R*n_change]: on entry the first n_base_classes hold the unperturbed parameter
 *                classes; the call appends the classes of every replica's n_change perturbed corridors
 *                (n_change = int(0.2 * corridors), at most 32);
 *   lk_class     [n_links*R] (replica fastest): base_class[l], or the replica's own class on a perturbed corridor;
 *   od_w         [sim_steps+1][n_od*R]: one weight ~ U[1, 10) per OD pair and replica, constant over the episode;
 *   dem_base / dem_peak / dem_pattern  [n_demand_rows*R]: the inputs of pns_env_draw_demand (rows with
 *                row_is_origin == 0 get pattern -1).
 * The caller then points pns_net.classes / lk_class at these arrays with per_replica_scenario = 1. */
int pns_env_randomize(const pns_net *net, pns_link_class *classes, int n_base_classes, const int32_t *base_class,
                      int n_change, int32_t *lk_class, double *od_w, int n_demand_rows, const int32_t *row_is_origin,
                      double *dem_base, double *dem_peak, int32_t *dem_pattern, uint64_t seed, uint32_t replica_base,
                      void *stream);

/* Episode KPIs of every replica from the history rows 0..t_last (reference rl/rl_utils.py, which computes them
 * from the JSON that handlers/output_handler.py saves): out[replica][PNS_KPI_COUNT].
 *   TOTAL_DEMAND     sum of the origin demand                                  (compute_network_throughput :827-835)
 *   TOTAL_OUTFLOW    sum of cumulative_outflow[t_last] over links ending at a destination   (:840-858, :1243-1262)
 *   TOTAL_INFLOW     sum of cumulative_inflow[t_last] over links starting at an origin      (:1136-1156, :1225-1241)
 *   PERSON_TIME      sum_t sum_l N * dt                                        (compute_average_travel_time_spent :1124-1133)
 *   PERSON_TIME_MOVING, TOTAL_DELAY   the same sum and N * max(0, 1 - T_ff/T) * dt over entries with T > 0
 *                                                                              (compute_total_network_delay :1010-1052)
 *   CONGESTION_TIME, AREA_TIME, STEPS, CONGESTED_STEPS                         (compute_network_congestion_metric :1462-1486)
 *   AVG_TRAVEL_TIME  mean over OD-path links of the time-mean travel time      (compute_network_travel_time :915-948)
 * The ratios the reference returns (throughput, served-trips rate, delay intensity, ...) are quotients of these.
 * lk_role[l]: bit0 link starts at an origin, bit1 ends at a destination, bit2 lies on an OD path (any_od_path = 0:
 * no paths known, every link counts).  scratch: [n_links*replicas*8] doubles. */
enum { PNS_KPI_TOTAL_DEMAND = 0, PNS_KPI_TOTAL_OUTFLOW, PNS_KPI_TOTAL_INFLOW, PNS_KPI_PERSON_TIME,
       PNS_KPI_PERSON_TIME_MOVING, PNS_KPI_TOTAL_DELAY, PNS_KPI_CONGESTION_TIME, PNS_KPI_AREA_TIME, PNS_KPI_STEPS,
       PNS_KPI_CONGESTED_STEPS, PNS_KPI_AVG_TRAVEL_TIME, PNS_KPI_COUNT };
int pns_kpi(const pns_net *net, const pns_state *st, const pns_step_io *io, int t_last, const int32_t *lk_role,
            int any_od_path, double *scratch, double *out, void *stream);

/* RegularNode.solve(type='optimal') (node.py:249-271) for a batch of `n` independent nodes of `m` link slots: the
 * linear program that the reference gives to scipy.optimize.linprog (constraint matrices node.py:73-104, :110-137;
 * objective -sum x + w sum |phi X - x|), one warp per program.  s, r: [n][m] sending / receiving flows (>= 0), phi:
 * [n][m(m-1)] turning fractions; x: [n][m(m-1)] turn flows at the optimum (before the reference's floor), objective:
 * [n], info: [n] pivot count, | 1<<28 when a non-basic reduced cost is zero (the optimum may not be unique), | 1<<29
 * unbounded, | 1<<30 pivot limit.  All pointers are device memory.  Inside a step the same solver runs for the nodes
 * of kind 2 (pns_node_flows / pns_step). */
int pns_lp_solve(int m, int n, const double *s, const double *r, const double *phi, double w, double *x,
                 double *objective, int32_t *info, void *stream);

/* Links per CTA of the single-replica link kernel (the granularity of pns_net.lane_order). */
int pns_lane_block_size(void);

/* Draw `n` samples with the on-device Philox samplers (test hook for oracle/philox.py):
 * kind 0: binomial(n_trials[i], p[i]) -> out_i; kind 1: the four normals of quad key i -> out_d[4i..4i+3];
 * kind 2: det_pow08(p[i]) -> out_d[i]. */
int pns_rng_selftest(int kind, int n, const int32_t *n_trials, const double *p, uint64_t seed, int t,
                     int site, int32_t *out_i, double *out_d, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PNS_B200_H */
