// B200 (sm_100a) kernels of the Link-Transmission-Model timestep + their C-ABI launchers.
//
// One simulation step t (reference Network.network_loading, src/LTM/network.py:266-287) is
//
//   link kernel    FLOWS : sending + receiving flow at tau = t-1, handed to the node pass in node-major order
//                  UPDATE: the link's inflow/outflow of step t picked up from the node pass, cumulative
//                          counts, pedestrians, density, speed, travel time at t
//       k_link_pair<R1,PHASE,MODE>      one thread per (link pair, replica)      any replica count
//       k_link_lane<PHASE,MODE,ONECLASS> one thread per directed link            single replica (GPU only)
//       k_link_rep<PHASE,MODE,ONECLASS,ENV> one thread per (directed link, replica)  batched replicas (GPU only)
//   k_route_fractions one thread per (routed node, replica) logit P(down | up, od) of the node's groups + OD mixing
//                                                          into its turning fractions (routed nets only)
//   k_node_flows   one thread per (node, replica)         node model on contiguous node-major records
//
// and inside a multi-step call the UPDATE of step t and the FLOWS of step t+1 run in the same
// thread (the link state stays in registers), so a step costs two launches, chained with
// programmatic dependent launch:
//
//   FLOWS(t0) | node(t0) | UPDATE(t0)+FLOWS(t0+1) | node(t0+1) | ... | UPDATE(t0+n-1)
//
// Nodes are mutually independent within a step (every read is of rows <= t-1 or of values the same
// pass produced for the same link pair), which is what makes the step
// data-parallel.  All state is time-major structure-of-arrays history in HBM (include/pns_b200.h);
// the replica index is the fastest-varying one, so for batched replicas a warp touches 32
// consecutive elements of every row; for a single replica adjacent lanes are the two directions of
// a corridor.  Per-link parameters come from a small class table (L1) or, for single-class
// networks, from the kernel-parameter constant bank.
//
// Numerics: the reference mixes float32 history with float64 counters under numpy-2 scalar
// promotion; every expression below states its precision explicitly and the file is compiled
// with -fmad=false so that no multiply-add is contracted (SURVEY.md "dtype ledger").
#ifdef PNS_HOST_EMULATION
#include "pns_emu.h"   // tests/emu: sequential host build of these kernels for CPU-side unit tests
#else
#include <cuda_runtime.h>
#include <mutex>
#define PNS_LAUNCH(kern, nblk, nthr, stream, ...) kern<<<(nblk), (nthr), 0, (stream)>>>(__VA_ARGS__)
// Programmatic dependent launch (sm_90+): the kernels of a step form a strict chain on one stream.
// Each kernel lets its successor start launching at once (TRIGGER) and waits for its predecessor's
// memory only after its own index arithmetic and static-table loads (WAIT), so launch latency and
// the ramp-up of one kernel overlap the tail of the previous one.
#define PNS_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define PNS_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
template <typename K>
static inline void pns_launch_chain(K kern, dim3 nblk, unsigned nthr, cudaStream_t stream, const void* ctx_arg,
                                    size_t smem = 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = nblk; cfg.blockDim = dim3(nthr); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void* args[1] = {const_cast<void*>(ctx_arg)};
    cudaLaunchKernelExC(&cfg, (const void*)kern, args);
}
#define PNS_LAUNCH_CHAIN(kern, nblk, nthr, stream, ctx) pns_launch_chain(kern, (nblk), (nthr), (stream), &(ctx))
#endif
#ifdef PNS_HOST_EMULATION
#define PNS_PDL_TRIGGER() do { } while (0)
#define PNS_PDL_WAIT() do { } while (0)
#define PNS_LAUNCH_CHAIN(kern, nblk, nthr, stream, ctx) PNS_LAUNCH(kern, nblk, nthr, stream, ctx)
#endif
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pns_b200.h"
#include "pns_rng.cuh"
#include "pns_lp.cuh"

namespace {

thread_local char g_err[512] = "";

int fail(const char* what, cudaError_t e = cudaSuccess) {
    if (e != cudaSuccess)
        snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    else
        snprintf(g_err, sizeof g_err, "%s", what);
    return 1;
}

constexpr int kBlock = 128;
#ifndef PNS_MIN_BLOCKS
#define PNS_MIN_BLOCKS 4   // resident CTAs per SM the register allocator must allow (tuning knob)
#endif
#ifndef PNS_LANE_BLOCK
#define PNS_LANE_BLOCK 64      // threads per CTA of k_link_lane (measured: 64 < 128 < 256 < 512 in step time)
#endif
#ifndef PNS_PF_AHEAD_CTAS
#define PNS_PF_AHEAD_CTAS (99456 / PNS_LANE_BLOCK)   // k_link_lane: each CTA pulls the rows of the CTA this far ahead
                                                     // into L2: three quarters of a resident wave of 148 SMs x 896
                                                     // threads (round 1, at 1024 threads per SM, 1/4 .. 1 wave: 43.0 42.3
                                                     // 41.4 41.0 41.6 us; now 1/2, 3/4, 1 wave: 39.1 38.4 38.7; 0 = off)
#endif
#ifndef PNS_SKIP_ZERO_HANDOVER
#define PNS_SKIP_ZERO_HANDOVER 1
#endif
#ifndef PNS_QUIET_FAST_PATH
#define PNS_QUIET_FAST_PATH 1  // k_link_lane: exact shortcut for the sending flow of empty links
#endif
#ifndef PNS_PF_TAPS
#define PNS_PF_TAPS 1          // k_link_lane: early fetch of the diffusion taps of occupied links
#endif
#ifndef PNS_LANE_MIN_BLOCKS
#define PNS_LANE_MIN_BLOCKS (896 / PNS_LANE_BLOCK)   // 14 CTAs of 64 threads: 72 registers, no spills.  16 (64 registers) spilled
                                                     // 56 bytes around the calls of the round-2 samplers and cost 8 % early in
                                                     // the run, 5 % late (41.8 / 50.7 -> 38.5 / 48.2 us per step); 12: 40.4 / 50.4
#endif
#ifndef PNS_NODE_BLOCK
#define PNS_NODE_BLOCK 128     // threads per CTA of k_node_flows
#endif
#ifndef PNS_NODE_COLS_MAX_REPLICAS
#define PNS_NODE_COLS_MAX_REPLICAS (1 << 30)   // k_node_cols up to this many replicas per GPU (measured faster at 1024 .. 8192)
#endif
#ifndef PNS_NODE_MIN_BLOCKS
#define PNS_NODE_MIN_BLOCKS 4  // 72 registers, no spills; 4..8 measured within 1%
#endif
constexpr int PH_UPDATE = 1, PH_FLOWS = 2;

struct Ctx {
    pns_net n;
    pns_state s;
    pns_step_io io;
    int t;        // UPDATE acts on row t; route/node kernels compute step t
    int t_flows;  // FLOWS computes the flows of step t_flows (time index t_flows-1)
    int phase;    // PH_* mask
    int mode;     // PNS_RNG_*
    const int32_t* draw_b;  // TABLE: R1..R3 outcomes for step t_flows
    const double* draw_n;   // TABLE: R4 noise for step t
    const double* draw_exp; // TABLE: route-choice exponentials of step t (null: evaluated on the device)
    size_t row64, row32;    // elements per history row
    size_t fld64, fld32;    // elements per history field
    // rows known at launch time, resolved on the host (saves 64-bit index arithmetic per access)
    const float *u_num_prev, *u_tt_old;                 // row t-1; row t-window (null while t < window)
    float *u_num, *u_dens, *u_speed, *u_tt, *u_flow, *u_avg;   // UPDATE outputs, row t
    double *u_bgw, *u_sepw;
    const float *f_num, *f_dens, *f_avg;                // FLOWS inputs, row tau = t_flows-1
    const double *f_cin, *f_cou, *f_sndp, *f_rcvp;      // rows tau, tau, tau-1 (wrapped), tau-1
    double *f_snd, *f_rcv;                              // FLOWS outputs, row tau
    const double *n_coutp, *n_cinp, *n_demand;          // cumulative counts of row t-1; demand row of step t
    double *n_outflow, *n_inflow, *n_cout, *n_cin;      // flows and cumulative counts of row t
    // single-class networks: every link has the same lags, so the lagged rows of FLOWS are launch constants too
    const double *c0_coulag;                            // cumulative_outflow[tau+1-swtau] (null while negative)
    const double *c0_pre0, *c0_pre1;                    // cumulative_inflow rows of the two likeliest arrival lags
    int c0_pre_i0, c0_pre_i1;                           // their indices (-1: not applicable)
    double* metric;                                     // streamed runs: PNS_METRIC_SLOTS partial sums of num_pedestrians[t]
    // Route choice riding in a link launch that computes FLOWS(route_t): route_blocks extra CTAs (lane kernel: at the
    // front of the grid; batched kernel: extra grid rows) run route_thread for step route_t next to the link CTAs.  It
    // is independent of the FLOWS part; when the launch also runs UPDATE(route_t-1) (route_recompute), the pedestrian
    // counts of row route_t-1 are being written by the link CTAs, and the route threads rebuild the few they need from
    // the UPDATE inputs (num[t-2], inflow[t-1], outflow[t-1]) with the same arithmetic.
    int route_blocks, route_t, route_recompute;
    int route_all;   // 1: every route row is evaluated; 0: only pns_net.rt_dyn_rows (the others hold constants)
};

// Control environment (reference rl/builders.py, rl/pz_pednet_env.py): the step context plus the action /
// observation / reward programs
struct EnvCtx {
    Ctx c;
    pns_env env;
    const float* actions;
    float* obs;
    float* reward;
    float* cum_reward;      // optional running sum of the rewards (null: not kept)
    // software prefetch of k_link_rep (a schedule): each CTA pulls the first-batch rows of the corridor `pf_pairs`
    // further on (same replicas; about one resident wave of CTAs later) into L2, `pf_off` elements ahead; 0 = off
    unsigned pf_pairs;
    size_t pf_off;
};

template <bool R1> struct Lanes;   // how a thread's two links sit in a history row
template <> struct Lanes<true> {   // single replica: adjacent columns -> one vector access
    static __device__ __forceinline__ void ld(const double* row, size_t e0, size_t, double* o) {
        const double2 v = *reinterpret_cast<const double2*>(row + e0); o[0] = v.x; o[1] = v.y;
    }
    static __device__ __forceinline__ void ld(const float* row, size_t e0, size_t, float* o) {
        const float2 v = *reinterpret_cast<const float2*>(row + e0); o[0] = v.x; o[1] = v.y;
    }
    static __device__ __forceinline__ void st(double* row, size_t e0, size_t, double a, double b) {
        double2 v; v.x = a; v.y = b; *reinterpret_cast<double2*>(row + e0) = v;
    }
    static __device__ __forceinline__ void st(float* row, size_t e0, size_t, float a, float b) {
        float2 v; v.x = a; v.y = b; *reinterpret_cast<float2*>(row + e0) = v;
    }
};
template <> struct Lanes<false> {  // batched replicas: the two links are R elements apart
    static __device__ __forceinline__ void ld(const double* row, size_t e0, size_t e1, double* o) { o[0] = row[e0]; o[1] = row[e1]; }
    static __device__ __forceinline__ void ld(const float* row, size_t e0, size_t e1, float* o) { o[0] = row[e0]; o[1] = row[e1]; }
    static __device__ __forceinline__ void st(double* row, size_t e0, size_t e1, double a, double b) { row[e0] = a; row[e1] = b; }
    static __device__ __forceinline__ void st(float* row, size_t e0, size_t e1, float a, float b) { row[e0] = a; row[e1] = b; }
};

// ---- addressing ---------------------------------------------------------------------------------
__device__ __forceinline__ double* H64(const Ctx& c, int f, int t) {
    return c.s.hist64 + (size_t)f * c.fld64 + (size_t)t * c.row64;
}
__device__ __forceinline__ float* H32(const Ctx& c, int f, int t) {
    return c.s.hist32 + (size_t)f * c.fld32 + (size_t)t * c.row32;
}
// numpy-style index of a length-(S+1) series: negative wraps once, anything else is an IndexError
__device__ __forceinline__ int wrap_index(const Ctx& c, int i, int replica) {
    const int n = c.n.sim_steps + 1;
    if (i < 0) i += n;
    if (i < 0 || i >= n) {
        atomicOr(c.s.err + replica, PNS_ERR_HISTORY_INDEX);
        return -1;
    }
    return i;
}

#ifndef PNS_HOST_EMULATION
// ---- L2 residency hints (sm_80+ cache policies) -------------------------------------------------
// One step of a large network streams ~190 MB through a 126 MB L2, so with plain LRU little of what
// one kernel writes survives until the next kernel (or the next step) reads it.  Accesses can carry
// an evict_last policy (hand-over data: exchange arrays, static incidence tables, gate array, the
// rows the next step reads back) or evict_first (rows written or read once).  PNS_L2_STAGE selects
// how much is marked: 0 nothing, 1 exchange arrays + static tables, 2 + gate and running sums,
// 3 + the rows the next step reads back (with evict_first on their final read).
// Measured on the 512x512 lattice (profiles/r1_h_l2_*.md): stage 1 shortens the link kernel by 4 us
// (the index load at the head of its only dependent load chain now hits L2); stages 2 and 3 lower
// DRAM traffic a little but not the time, and a persisting-L2 set-aside
// (cudaLimitPersistingL2CacheSize) cuts traffic by 15% while making the step 25% slower -- the
// step is latency-bound, not bandwidth-bound.  Stage 1 is the default.
#ifndef PNS_L2_STAGE
#define PNS_L2_STAGE 1
#endif
struct L2Pol { uint64_t keep, once; };
__device__ __forceinline__ L2Pol l2_policies() {
    L2Pol p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.once));
    return p;
}
__device__ __forceinline__ double ldh(const double* a, uint64_t pol) {
    double v; asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol) : "memory"); return v;
}
__device__ __forceinline__ float ldh(const float* a, uint64_t pol) {
    float v; asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol) : "memory"); return v;
}
__device__ __forceinline__ double2 ldh(const double2* a, uint64_t pol) {
    double2 v; asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol) : "memory"); return v;
}
__device__ __forceinline__ int4 ldh_nc(const int4* a, uint64_t pol) {
    int4 v; asm volatile("ld.global.nc.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol)); return v;
}
__device__ __forceinline__ int2 ldh_nc(const int2* a, uint64_t pol) {
    int2 v; asm volatile("ld.global.nc.L2::cache_hint.v2.s32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(a), "l"(pol)); return v;
}
__device__ __forceinline__ void sth(double* a, double v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" :: "l"(a), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void sth(float* a, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(a), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* a) { asm volatile("prefetch.global.L2 [%0];" :: "l"(a)); }
__device__ __forceinline__ void prefetch_l1(const void* a) { asm volatile("prefetch.global.L1 [%0];" :: "l"(a)); }
// stage-gated forms: S = the stage from which the hint applies
template <int S, typename T> __device__ __forceinline__ T ld_keep(const T* a, const L2Pol p) { if (PNS_L2_STAGE >= S) return ldh(a, p.keep); return *a; }
template <int S, typename T> __device__ __forceinline__ T ld_once(const T* a, const L2Pol p) { if (PNS_L2_STAGE >= S) return ldh(a, p.once); return *a; }
template <int S, typename T> __device__ __forceinline__ void st_keep(T* a, T v, const L2Pol p) { if (PNS_L2_STAGE >= S) sth(a, v, p.keep); else *a = v; }
template <int S, typename T> __device__ __forceinline__ void st_once(T* a, T v, const L2Pol p) { if (PNS_L2_STAGE >= S) sth(a, v, p.once); else *a = v; }
__device__ __forceinline__ int4 ld_meta(const int4* a, const L2Pol p) { if (PNS_L2_STAGE >= 1) return ldh_nc(a, p.keep); return __ldg(a); }
#else   // host-emulation test build: plain accesses
struct L2Pol { };
static inline L2Pol l2_policies() { return L2Pol(); }
static inline void prefetch_l2(const void*) { }
static inline void prefetch_l1(const void*) { }
template <int S, typename T> static inline T ld_keep(const T* a, const L2Pol) { return *a; }
template <int S, typename T> static inline void st_keep(T* a, T v, const L2Pol) { *a = v; }
static inline int4 ld_meta(const int4* a, const L2Pol) { return *a; }
#endif

typedef pns_link_class LinkP;
__device__ __forceinline__ bool is_sep(const LinkP& p) { return p.flags & 1; }
// parameter class of link l in replica rep (per-replica tables under domain randomisation)
__device__ __forceinline__ int class_of(const Ctx& c, int l, int rep) {
    return c.n.per_replica_scenario ? __ldg(c.n.lk_class + (size_t)l * c.n.replicas + rep) : __ldg(c.n.lk_class + l);
}

// Walkable area (link.py:128-131, 454-456).  For a separator it follows the lane width; `f64`
// tells whether the reference holds that width as a numpy float64 (assigned from np.clip), in
// which case density is a double-precision divide instead of a float32 one.
struct Area {
    double area, space;
    float area32;
    bool f64;
};
__device__ __forceinline__ Area link_area(const Ctx& c, const LinkP& p, size_t e, double gate, bool np64_now = false) {
    Area a;
    if (is_sep(p)) {
        a.area = p.length * gate;
        a.space = p.kj * a.area;
        a.area32 = (float)a.area;
        a.f64 = np64_now || c.s.sep_np64[e] != 0;
    } else {
        a.area = p.area; a.space = p.space; a.area32 = p.area32; a.f64 = false;
    }
    return a;
}
__device__ __forceinline__ float div_by_area(float x, const Area& a) {
    if (x == 0.0f) return 0.0f;   // exact (areas are positive); keeps empty links off the division slow path
    return a.f64 ? (float)((double)x / a.area) : x / a.area32;
}

// Python min/max on scalars: min(a, b) -> b if b < a else a ; max(a, b) -> b if b > a else a
__device__ __forceinline__ double pymin(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double pymax(double a, double b) { return b > a ? b : a; }
__device__ __forceinline__ float clip01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

struct SendOut {
    double flow;   // sending_flow[tau]
    double sval;   // REQUEST: flow after the release stage when no R1 draw is needed
    int kind;      // REQUEST: 0 nothing to draw, 1 diffusion branch, 2 binomial R1
    int n1;
    float rf;
};

// Link.get_outflow (link.py:199-214): 4-tap geometric smoothing of lagged inflow
__device__ __noinline__ double diffusion_outflow(const Ctx& c, const LinkP& p, size_t e, int tau, int lag,
                                                 float avg_tt, int replica) {
    const float F = 1.0f / (1.0f + p.gamma32 * avg_tt);
    const float u = 1.0f - F;
    const float c1 = F * u;
    const float c2 = F * (float)((double)u * (double)u);                 // powf(u, 2)
    const float c3 = F * (float)(((double)u * (double)u) * (double)u);   // powf(u, 3)
    double in[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = wrap_index(c, tau - lag - k, replica);
        in[k] = i >= 0 ? H64(c, PNS_F64_INFLOW, i)[e] : 0.0;
    }
    const double total = (((double)F * in[0] + (double)c1 * in[1]) + (double)c2 * in[2]) + (double)c3 * in[3];
    const double up = ceil(total);
    return 0.0 > up ? 0.0 : up;
}

struct LinkNow {      // state of one link at time index tau, either just computed or loaded
    float num, dens, avg_tt;
};

// Link.cal_sending_flow at time index tau (link.py:216-370)
template <int MODE>
__device__ __forceinline__ SendOut sending_flow(const Ctx& c, const LinkP& p, size_t e, int tau, const LinkNow& me,
                                                float num_rev, const Area& ar, double front_gate, double cum_out_tau,
                                                double snd_prev, int replica, const pns::DrawKey& key,
                                                int pre_idx0 = -1, double pre_val0 = 0.0, int pre_idx1 = -1,
                                                double pre_val1 = 0.0) {
    SendOut o;
    o.kind = 0; o.n1 = 0; o.rf = 0.0f; o.sval = 0.0; o.flow = 0.0;
    if (tau < p.fftau) return o;                                          // link.py:267-269
    const float dens = is_sep(p) ? me.dens : div_by_area(me.num + num_rev, ar);
    const int lag = __float2int_rn(me.avg_tt / (float)c.n.unit_time);     // round(), half to even (link.py:260)
    const int idx = max(0, tau + 1 - lag);
    const float cong = clip01((me.dens - p.kc32) / p.kj_minus_kc32);
    // cumulative_inflow[idx]: the caller may have fetched the rows of the two most likely lags early
    const double cin_idx = idx == pre_idx0 ? pre_val0 : idx == pre_idx1 ? pre_val1 : H64(c, PNS_F64_CUM_INFLOW, idx)[e];
    const double arrived_raw = cin_idx - cum_out_tau;
    const double arrived = arrived_raw > 0.0 ? arrived_raw : 0.0;
    const double boundary = (double)(cong * me.num) + (double)(1.0f - cong) * arrived;
    const double gate_cap = ((front_gate * p.kc) * p.vf) * c.n.unit_time;
    double flow = pymin(boundary, gate_cap);
    const double original = flow;
    // a zero lag makes the reference's result depend on its node visiting order; it only matters when the link sends
    if (lag == 0 && flow > 0.0) atomicOr(c.s.err + replica, PNS_ERR_ZERO_LAG);
    if (flow > 0.0) {
        const float rf = dens == 0.0f ? 0.0f : clip01(dens / p.kj32);
        o.rf = rf;
        bool draw = true;
        if (dens <= p.kc32) {
            const double spread = diffusion_outflow(c, p, e, tau, lag, me.avg_tt, replica);
            if (spread > 0.0) {
                flow = floor(pymin(0.8 * spread + 0.19999999999999996 * flow, flow));   // link.py:330
                draw = false;
                o.kind = 1;
            }
        }
        if (draw) {
            const int trials = (int)floor(flow);
            o.kind = 2;
            o.n1 = trials;
            if (MODE == PNS_RNG_TABLE) {
                flow = (double)c.draw_b[e];
            } else if (MODE == PNS_RNG_PHILOX) {
                const float p32 = 0.7f + 0.15f * pns::det_pow08(rf);
                flow = (double)pns::binomial_release(key, trials, (double)p32);
            } else {
                return o;   // REQUEST: the host draws R1 (and R2, which depends on it)
            }
        }
    }
    o.sval = flow;
    if (MODE == PNS_RNG_REQUEST) return o;
    if (p.act > 0.0 && flow > 1.0) {                                       // link.py:351-358
        const int trials = (int)floor(flow);
        int stay;
        if (MODE == PNS_RNG_TABLE) stay = c.draw_b[c.row32 + e];
        else stay = pns::binomial_philox(key, 2u, trials, p.act);
        flow = flow - (double)stay;
    }
    flow = flow > 0.0 ? flow : 0.0;
    flow = pymin(floor(0.8 * flow + 0.2 * snd_prev), original);           // link.py:363-364
    if (flow < 0.0) atomicOr(c.s.err + replica, PNS_ERR_NEG_SENDING);
    o.flow = flow;
    return o;
}

// Link/Separator.cal_receiving_flow at tau, before the reverse sending flow is subtracted
// (link.py:372-405, 480-507)
template <int MODE>
__device__ __forceinline__ double receiving_flow(const Ctx& c, const LinkP& p, size_t e, int tau, float num_rev,
                                                 const Area& ar, double back_gate, double cum_in_tau,
                                                 double cum_out_lag, double rcv_prev, const pns::DrawKey& key,
                                                 int* n3) {
    const int lag_i = tau + 1 - p.swtau;   // cum_out_lag = cumulative_outflow[lag_i] when lag_i >= 0
    double bound;
    if (is_sep(p)) {
        *n3 = -1;
        if (lag_i < 0) bound = ar.space;
        else bound = (cum_out_lag + ar.space) - cum_in_tau;
    } else {
        const int trials = (int)num_rev;        // numpy casts the float32 count to int64 (truncation)
        *n3 = trials;
        int blockers = 0;
        if (MODE == PNS_RNG_TABLE) blockers = c.draw_b[2 * c.row32 + e];
        else if (MODE == PNS_RNG_PHILOX) blockers = pns::binomial_blockers(key, trials);
        else return 0.0;
        if (lag_i < 0) {
            bound = ar.space - (double)blockers;
        } else {
            const double x = ((cum_out_lag + ar.space) - (double)blockers) - cum_in_tau;
            bound = x > 0.0 ? x : 0.0;
        }
    }
    if (MODE == PNS_RNG_REQUEST) return 0.0;
    const double gate_cap = ((back_gate * p.kc) * p.vf) * c.n.unit_time;
    double flow = pymin(bound, gate_cap);
    flow = pymax(flow, 0.0);
    if (rcv_prev >= 0.0) flow = pymin(floor(flow * 0.8 + rcv_prev * 0.2), flow);   // link.py:400-401
    return flow;
}

// BiDirectionalFd.__call__ (src/utils/functions.py:112-134) + travel time (link.py:176-177).
// Returns the float32 value stored in speed[t]; *tt is the float32 stored in travel_time[t].
__device__ __forceinline__ float speed_and_travel_time(const LinkP& p, float k_self, float k_opp, bool have_noise,
                                                       double z, float* tt) {
    const float k_eff = k_self + p.bi32 * k_opp;
    const int fd = (p.flags >> 1) & 3;
    // the speed is a Python double in the free-flow branch (and after `0 + noise`), a float32 otherwise
    bool is_f64 = false, is_zero = false;
    double v64 = 0.0;
    float v32 = 0.0f;
    const bool free_flow = k_eff <= p.kc32;
    if (fd == 2) {                                    // smulders
        if (free_flow) v32 = p.vf32 * (1.0f - k_eff / p.kj32);
        else {
            v32 = p.sm_gamma32 * (1.0f / k_eff - p.inv_kj32);
            if (!(v32 > 0.0f)) is_zero = true;
        }
    } else if (free_flow) {
        is_f64 = true;
        v64 = p.vf;
    } else if (fd == 0) {                             // yperman
        v32 = p.yp_coef32 * (p.kj32 / k_eff - 1.0f);
        if (!(v32 > 0.0f)) is_zero = true;
    } else {                                          // greenshields
        v32 = (p.neg_vf32 * (k_eff - p.kj32)) / p.kj_minus_kc32;
        if (!(v32 > 0.0f)) is_zero = true;
    }
    if (have_noise) {
        if (is_zero) { is_f64 = true; is_zero = false; v64 = 0.0 + z; }
        else if (is_f64) v64 = v64 + z;
        else v32 = v32 + (float)z;
    }
    if (is_zero) { *tt = p.max_tt32; return 0.0f; }
    if (is_f64) {
        if (!(v64 > 0.0)) { *tt = p.max_tt32; return 0.0f; }
        *tt = (float)(p.length / v64);
        return (float)v64;
    }
    if (!(v32 > 0.0f)) { *tt = p.max_tt32; return 0.0f; }
    *tt = p.length32 / v32;
    return v32;
}

// =================================================================================================
// UPDATE: Link.update_link_density_flow + update_speeds at row t (link.py:133-188; Separator :430-452)
// FLOWS : sending/receiving flows of step t_flows (time index t_flows-1)
// Every load that does not depend on a computed lag is issued before the first arithmetic so the
// whole batch is in flight at once; only cumulative_inflow[idx] and the diffusion taps are dependent.
template <bool R1, int PHASE, int MODE>
__device__ __forceinline__ void link_pair_body(const Ctx& c) {
    const int R = R1 ? 1 : c.n.replicas;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_pairs = (size_t)(c.n.n_links / 2);
    if (gid >= n_pairs * R) return;
    const int pair = R1 ? (int)gid : (int)(gid / R);
    const int rep = R1 ? 0 : (int)(gid % R);
    const int l0 = 2 * pair;
    const size_t e[2] = {(size_t)l0 * R + rep, (size_t)(l0 + 1) * R + rep};
    PNS_PDL_TRIGGER();
    PNS_PDL_WAIT();
    typedef Lanes<R1> V;
    constexpr bool upd = (PHASE & PH_UPDATE) != 0, flw = (PHASE & PH_FLOWS) != 0;
    const int tau = c.t_flows - 1;

    // networks with a single parameter class (lattices, the shipped 45_intersections) skip the index load
    const bool one_class = c.n.n_classes == 1;
    const LinkP* const pp[2] = {c.n.classes + (one_class ? 0 : class_of(c, l0, rep)),
                                c.n.classes + (one_class ? 0 : class_of(c, l0 + 1, rep))};
    double gate[2];
    V::ld(c.s.gate, e[0], e[1], gate);
    // ---- batch of independent loads --------------------------------------------------------
    double din[2] = {0, 0}, dout[2] = {0, 0}, cin_p[2] = {0, 0}, cou_p[2] = {0, 0};
    float np_[2] = {0, 0}, rs[2] = {0, 0}, tt_old[2] = {0, 0};
    const bool windowed = c.u_tt_old != nullptr;
    // where this pair's flows live in the node-major exchange arrays (see k_node_flows)
    const int4 slots = __ldg(reinterpret_cast<const int4*>(c.n.lk_slots) + pair);   // {s0, r0, s1, r1}
    if (upd) {
        // Node.update_links for this pair (node.py:146-162): the node pass wrote inflow[t] / outflow[t]
        // of every link; fetch ours and extend the cumulative counts
        V::ld(c.n_inflow, e[0], e[1], din);
        V::ld(c.n_outflow, e[0], e[1], dout);
        V::ld(c.n_cinp, e[0], e[1], cin_p);
        V::ld(c.n_coutp, e[0], e[1], cou_p);
        V::ld(c.u_num_prev, e[0], e[1], np_);
        V::ld(c.s.runsum, e[0], e[1], rs);
        if (windowed) V::ld(c.u_tt_old, e[0], e[1], tt_old);
    }
    double cin_tau[2] = {0, 0}, cou_tau[2] = {0, 0}, snd_prev[2] = {0, 0}, rcv_prev[2] = {0, 0}, cou_lag[2] = {0, 0};
    float ld_num[2] = {0, 0}, ld_dens[2] = {0, 0}, ld_avg[2] = {0, 0};
    if (flw) {
        if (!upd) {                     // fused launches carry the cumulative counts in registers
            V::ld(c.f_cin, e[0], e[1], cin_tau);
            V::ld(c.f_cou, e[0], e[1], cou_tau);
        }
        V::ld(c.f_sndp, e[0], e[1], snd_prev);
        V::ld(c.f_rcvp, e[0], e[1], rcv_prev);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int lag_i = tau + 1 - pp[a]->swtau;                      // static shock-wave lag (link.py:380)
            if (lag_i >= 0) cou_lag[a] = H64(c, PNS_F64_CUM_OUTFLOW, lag_i)[e[a]];
        }
        if (!upd) {
            V::ld(c.f_num, e[0], e[1], ld_num);
            V::ld(c.f_dens, e[0], e[1], ld_dens);
            V::ld(c.f_avg, e[0], e[1], ld_avg);
        }
    }
    Area ar[2];
    ar[0] = link_area(c, *pp[0], e[0], gate[0]);
    ar[1] = link_area(c, *pp[1], e[1], gate[1]);
    const uint32_t k0 = (uint32_t)c.io.seed, k1 = (uint32_t)(c.io.seed >> 32);

    LinkNow now[2];
    if (upd) {
        const int t = c.t;
#pragma unroll
        for (int a = 0; a < 2; ++a) {                                      // link.py:19-25
            cin_tau[a] = cin_p[a] + din[a];
            cou_tau[a] = cou_p[a] + dout[a];
        }
        V::st(c.n_cin, e[0], e[1], cin_tau[0], cin_tau[1]);
        V::st(c.n_cout, e[0], e[1], cou_tau[0], cou_tau[1]);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            now[a].num = (float)((double)np_[a] + (din[a] - dout[a]));    // link.py:134-135
            now[a].dens = div_by_area(now[a].num, ar[a]);                  // link.py:136
        }
        // speed noise (functions.py:132-133): one Philox block serves both directions of the pair
        double z[2] = {0.0, 0.0};
        const bool noisy[2] = {pp[0]->sigma > 0.0, pp[1]->sigma > 0.0};
        if (noisy[0] | noisy[1]) {
            if (MODE == PNS_RNG_TABLE) {
                V::ld(c.draw_n, e[0], e[1], z);
            } else {
                pns::DrawKey key;
                key.t = (uint32_t)t; key.link = (uint32_t)l0; key.replica = (uint32_t)rep + c.io.replica_base; key.k0 = k0; key.k1 = k1;
                double g0, g1;
                pns::normal_pair_philox(key, 4u, &g0, &g1);
                z[0] = pp[0]->sigma * g0;
                z[1] = pp[1]->sigma * g1;
            }
        }
        float v[2], tt[2], sum[2], lf[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const LinkP& p = *pp[a];
            v[a] = speed_and_travel_time(p, now[a].dens, is_sep(p) ? 0.0f : now[1 - a].dens, noisy[a], z[a], &tt[a]);
            lf[a] = v[a] * now[a].dens;                                    // functions.py:97-101
            sum[a] = rs[a] + tt[a];                                        // link.py:183-186
            if (windowed) {
                sum[a] = sum[a] - tt_old[a];
                now[a].avg_tt = sum[a] / (float)c.n.window;
            } else {
                now[a].avg_tt = p.tt0;                                     // rows < window keep travel_time[0]
            }
        }
        V::st(c.u_num, e[0], e[1], now[0].num, now[1].num);
        V::st(c.u_dens, e[0], e[1], now[0].dens, now[1].dens);
        V::st(c.u_speed, e[0], e[1], v[0], v[1]);
        V::st(c.u_tt, e[0], e[1], tt[0], tt[1]);
        V::st(c.u_flow, e[0], e[1], lf[0], lf[1]);
        if (windowed) V::st(c.u_avg, e[0], e[1], now[0].avg_tt, now[1].avg_tt);
        V::st(c.s.runsum, e[0], e[1], sum[0], sum[1]);
        V::st(c.u_bgw, e[0], e[1], gate[0], gate[1]);                      // link.py:188, 451-452
#pragma unroll
        for (int a = 0; a < 2; ++a)
            if (is_sep(*pp[a])) c.u_sepw[e[a]] = gate[a];
    }
    if (!flw) return;

    if (!upd) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            now[a].num = ld_num[a];
            now[a].dens = ld_dens[a];
            now[a].avg_tt = ld_avg[a];
        }
    }
    SendOut s[2];
    double r[2];
    int n3[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const LinkP& p = *pp[a];
        pns::DrawKey key;
        key.t = (uint32_t)c.t_flows; key.link = (uint32_t)(l0 + a); key.replica = (uint32_t)rep + c.io.replica_base; key.k0 = k0; key.k1 = k1;
        // front gate of a plain link is the back gate of its reverse (link.py:110-126); a separator's
        // gates both equal its lane width (link.py:462-478)
        const double front = is_sep(p) ? gate[a] : gate[1 - a];
        s[a] = sending_flow<MODE>(c, p, e[a], tau, now[a], now[1 - a].num, ar[a], front, cou_tau[a], snd_prev[a], rep,
                                  key);
        r[a] = receiving_flow<MODE>(c, p, e[a], tau, now[1 - a].num, ar[a], gate[a], cin_tau[a], cou_lag[a],
                                    rcv_prev[a], key, &n3[a]);
    }
    if (MODE == PNS_RNG_REQUEST) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            c.io.req_kind[e[a]] = s[a].kind;
            c.io.req_n1[e[a]] = s[a].n1;
            c.io.req_rf[e[a]] = s[a].rf;
            c.io.req_sval[e[a]] = s[a].sval;
            c.io.req_n3[e[a]] = n3[a];
        }
        return;
    }
    // cal_receiving_flow_with_reverse (link.py:407-416; separators ignore the reverse flow, :509-512)
    const double q0 = pymax(is_sep(*pp[0]) ? r[0] : r[0] - s[1].flow, 0.0);
    const double q1 = pymax(is_sep(*pp[1]) ? r[1] : r[1] - s[0].flow, 0.0);
    V::st(c.f_snd, e[0], e[1], s[0].flow, s[1].flow);
    V::st(c.f_rcv, e[0], e[1], q0, q1);
    // hand the flows to the node pass in node-major order (slot of each link at its end / start node)
    c.s.nm_s[(size_t)slots.x * R + rep] = s[0].flow; c.s.nm_r[(size_t)slots.y * R + rep] = q0;
    c.s.nm_s[(size_t)slots.z * R + rep] = s[1].flow; c.s.nm_r[(size_t)slots.w * R + rep] = q1;
}

template <bool R1, int PHASE, int MODE>
__global__ void __launch_bounds__(kBlock, PNS_MIN_BLOCKS) k_link_pair(const __grid_constant__ Ctx c) {
    link_pair_body<R1, PHASE, MODE>(c);
}

// =================================================================================================
// PathFinder.update_node_turn_probs (path_finder.py:561-589) for one (od, upstream) group g: the logit
// P(down | up, od) over the group's options, written to st->probs.  REQUEST mode writes the arguments of the
// exponentials instead (numpy-compatible stepping evaluates them with the host's numpy: its exp differs from
// CUDA's in the last bit of a few percent of arguments, and a last-bit change of a turning fraction can move
// floor(P*s) across an integer); TABLE mode takes the exponentials from draw_exp when it is given.
__device__ __forceinline__ void group_probs(const Ctx& c, int g, int rep, int t) {
    const int R = c.n.replicas;
    const int o0 = __ldg(c.n.rt_opt_ptr + g), o1 = __ldg(c.n.rt_opt_ptr + g + 1);
    const int n = o1 - o0;
    if (n == 1 && c.mode != PNS_RNG_REQUEST) {
        // a single option: exp(x) / exp(x) = 1 exactly (x = -temp*U is far from the over/underflow range)
        c.s.probs[(size_t)o0 * R + rep] = 1.0;
        return;
    }
    const bool wide = __ldg(c.n.rt_grp_has_virtual + g) != 0;   // np.array([... float32 ..., 0]) is float64
    const int tm1 = t - 1;
    const int tm2 = wrap_index(c, t - 2, rep);
    const float* num = H32(c, PNS_F32_NUM_PED, tm1);
    const float* dens_row = H32(c, PNS_F32_DENSITY, tm1);
    const double* rcv = H64(c, PNS_F64_RECEIVING, tm2);
    // pedestrians at t-1; rebuilt from the inputs of UPDATE(t-1) when that update runs in the same launch
    // (link.py:134-135: float32(num[t-2] + (inflow[t-1] - outflow[t-1])))
    const bool rebuild = c.route_recompute != 0;
    const float* num_prev = rebuild ? H32(c, PNS_F32_NUM_PED, tm1 - 1) : nullptr;
    const double* in_row = rebuild ? H64(c, PNS_F64_INFLOW, tm1) : nullptr;
    const double* out_row = rebuild ? H64(c, PNS_F64_OUTFLOW, tm1) : nullptr;
#define PNS_NUM_AT(x) (rebuild ? (float)((double)num_prev[x] + (in_row[x] - out_row[x])) : num[x])

    float dens[PNS_MAX_DEGREE];
    double cap[PNS_MAX_DEGREE];
    double sum_d = 0.0, sum_c = 0.0;
    for (int k = 0; k < n; ++k) {
        const int l = __ldg(c.n.rt_opt_link + o0 + k);
        sum_d = k == 0 ? __ldg(c.n.rt_opt_dist + o0) : sum_d + __ldg(c.n.rt_opt_dist + o0 + k);
        if (l >= 0) {
            const LinkP& p = c.n.classes[class_of(c, l, rep)];
            const size_t e = (size_t)l * R + rep;
            const double gate = c.s.gate[e];
            const Area ar = link_area(c, p, e, gate);
            if (is_sep(p)) dens[k] = rebuild ? div_by_area(PNS_NUM_AT(e), ar) : dens_row[e];
            else dens[k] = div_by_area(PNS_NUM_AT(e) + PNS_NUM_AT((size_t)(l ^ 1) * R + rep), ar);
            const double last = rcv[e];
            cap[k] = last >= 0.0 ? last : ((gate * p.vf) * p.kc) * c.n.unit_time;
        } else {
            dens[k] = 0.0f;
            cap[k] = 100.0;
        }
        sum_c = k == 0 ? cap[0] : sum_c + cap[k];
    }
    double ex[PNS_MAX_DEGREE];
    double sum_e = 0.0;
    for (int k = 0; k < n; ++k) {
        double crowd_term;
        if (wide) {
            const double d = (double)dens[k] - 2.0;
            crowd_term = c.n.rt_beta * ((d > 0.0 ? d : 0.0) / 8.0);
        } else {
            const float d = dens[k] - 2.0f;
            crowd_term = (double)((float)c.n.rt_beta * (fmaxf(d, 0.0f) / 8.0f));
        }
        const double util = (((c.n.rt_alpha * __ldg(c.n.rt_opt_dist + o0 + k)) / (sum_d + 1e-6) + crowd_term) -
                             (c.n.rt_omega * cap[k]) / (sum_c + 1e-6)) + c.n.rt_eps;
        const double arg = -c.n.rt_temp * util;
        if (c.mode == PNS_RNG_REQUEST) {
            c.io.req_exp[(size_t)(o0 + k) * R + rep] = arg;
            continue;
        }
        // the operation-exact restatement when the whole step is counter-based (oracle/philox.py reproduces it
        // bit for bit anywhere); the host's exponentials in numpy-compatible stepping; libdevice otherwise
        ex[k] = c.mode == PNS_RNG_PHILOX ? pns::det_exp(arg)
              : c.draw_exp ? c.draw_exp[(size_t)(o0 + k) * R + rep] : exp(arg);
        sum_e = k == 0 ? ex[0] : sum_e + ex[k];
    }
#undef PNS_NUM_AT
    if (c.mode == PNS_RNG_REQUEST) return;
    for (int k = 0; k < n; ++k) c.s.probs[(size_t)(o0 + k) * R + rep] = ex[k] / sum_e;
}

// =================================================================================================
// PathFinder.update_turning_fractions + check_fractions (path_finder.py:591-715) for one upstream slot i of a
// routed node (row = the node's first row + i): P(od | up) from this step's OD weights, the m-1 fractions of
// the row as sum over ODs of P(down | up, od) * P(od | up) in the reference's accumulation order, and the row
// check.  Writes element j of the row to out[j*R].
__device__ __forceinline__ void routed_row_fractions(const Ctx& c, int routed, int m, int i, int row, int t, int rep,
                                                     double* out) {
    const int R = c.n.replicas;
    const int edge0 = __ldg(c.n.rt_routed_edge0 + routed) + i * (m - 1);
    // od weights of this step; one set per replica under domain randomisation
    const size_t ws = c.n.per_replica_scenario ? (size_t)R : 1;
    const double* w = c.io.od_w + (size_t)t * c.n.n_od * ws + (c.n.per_replica_scenario ? rep : 0);
    const int ra = __ldg(c.n.rt_row_ptr + row), rb = __ldg(c.n.rt_row_ptr + row + 1);
    // a single registered OD: P(od | up) = w/w = 1, or the uniform 1/1 when w = 0 (weights are finite, >= 0)
    const bool single = rb - ra == 1;
    double total = 0.0;
    if (!single)
        for (int x = ra; x < rb; ++x) total = total + w[(size_t)__ldg(c.n.rt_row_od + x) * ws];
    const double uniform = rb > ra ? 1.0 / (double)(rb - ra) : 0.0;
    double row_sum = 0.0;
    for (int j = 0; j < m - 1; ++j) {
        const int ta = __ldg(c.n.rt_term_ptr + edge0 + j), tb = __ldg(c.n.rt_term_ptr + edge0 + j + 1);
        double acc = 0.0;
        for (int x = ta; x < tb; ++x) {
            const double od_p = single ? 1.0 : total > 0.0 ? w[(size_t)__ldg(c.n.rt_term_od + x) * ws] / total : uniform;
            acc = acc + c.s.probs[(size_t)__ldg(c.n.rt_term_opt + x) * R + rep] * od_p;
        }
        out[(size_t)j * R] = acc;
        row_sum = j == 0 ? acc : row_sum + acc;
    }
    if (fabs(row_sum - 1.0) > 1e-3) {                                      // check_fractions
        if (row_sum > 1e-6) {
            for (int j = 0; j < m - 1; ++j) out[(size_t)j * R] = out[(size_t)j * R] / row_sum;
        } else {
            for (int j = 0; j < m - 1; ++j) out[(size_t)j * R] = 1.0 / (double)(m - 1);
        }
    }
}

// Route choice, one thread per (upstream slot of a routed node, replica): the probabilities of the (od, upstream)
// groups registered at that slot, then the slot's row of turning fractions (tf_routed; read by the node pass of
// the same step).  Runs for every routed node on every step, as the reference does (network.py:273-278), so the
// fractions a caller reads back are those of the last step.
__device__ __forceinline__ int route_rows(const Ctx& c) { return c.route_all ? c.n.n_rows : c.n.n_dyn_rows; }
__device__ __forceinline__ void route_thread(const Ctx& c, int t, int row_index, int rep) {
    const int R = c.n.replicas;
    PNS_PDL_TRIGGER();
    const int row = c.route_all ? row_index : __ldg(c.n.rt_dyn_rows + row_index);
    const int routed = __ldg(c.n.rt_row_routed + row);
    const int ga = __ldg(c.n.rt_row_grp_ptr + row), gb = __ldg(c.n.rt_row_grp_ptr + row + 1);
    const int i = row - __ldg(c.n.rt_routed_row0 + routed);
    const int4 meta = __ldg(reinterpret_cast<const int4*>(c.n.nd_meta) + __ldg(c.n.rt_routed_nodes + routed));
    const int m = meta.y & 0xff;
    PNS_PDL_WAIT();
    for (int x = ga; x < gb; ++x) group_probs(c, __ldg(c.n.rt_row_grp + x), rep, t);
    if (c.mode == PNS_RNG_REQUEST) return;
    routed_row_fractions(c, routed, m, i, row, t, rep, c.s.tf_routed + (size_t)(meta.w + i * (m - 1)) * R + rep);
}
#ifndef PNS_HOST_EMULATION
// Single replica: one *warp* per row.  A row thread is a chain of a few thousand dependent fp64 instructions (the
// logit of each of its (od, upstream) groups, then the OD mixing): 26 us for nine_intersections, more than the
// link and node kernels together.  The groups are independent, so the lanes take one each; lane 0 then mixes.
__device__ __forceinline__ void route_warp(const Ctx& c, int t, int row_index, unsigned lane) {
    PNS_PDL_TRIGGER();
    const int row = c.route_all ? row_index : __ldg(c.n.rt_dyn_rows + row_index);
    const int routed = __ldg(c.n.rt_row_routed + row);
    const int ga = __ldg(c.n.rt_row_grp_ptr + row), gb = __ldg(c.n.rt_row_grp_ptr + row + 1);
    const int i = row - __ldg(c.n.rt_routed_row0 + routed);
    const int4 meta = __ldg(reinterpret_cast<const int4*>(c.n.nd_meta) + __ldg(c.n.rt_routed_nodes + routed));
    const int m = meta.y & 0xff;
    PNS_PDL_WAIT();
    for (int x = ga + (int)lane; x < gb; x += 32) group_probs(c, __ldg(c.n.rt_row_grp + x), 0, t);
    __syncwarp();                              // orders the lanes' stores of the probabilities before lane 0's loads
    if (lane == 0 && c.mode != PNS_RNG_REQUEST)
        routed_row_fractions(c, routed, m, i, row, t, 0, c.s.tf_routed + (size_t)(meta.w + i * (m - 1)));
}
#endif
__global__ void __launch_bounds__(kBlock) k_route_fractions(const __grid_constant__ Ctx c) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned R = (unsigned)c.n.replicas;
    if (gid >= (size_t)route_rows(c) * R) return;
    route_thread(c, c.t, (int)((unsigned)gid / R), (int)((unsigned)gid % R));
}

// floor(min(w, r * (w / D))) of node.py:296-298 with two exact shortcuts that avoid the IEEE division:
//  * w == 0  ->  0/D = 0, r*0 = 0, min(0, 0) = 0;
//  * r >= 2D ->  r*(w/D) >= 2w(1 - 2^-52) > w for w > 0, so the minimum is w itself.
// Otherwise the division is evaluated as the reference does (quotient first, then the product).
__device__ __forceinline__ double turn_flow(double w, double r, double D) {
    if (w == 0.0) return 0.0;
    if (w > 0.0 && D > 0.0 && r >= 2.0 * D && r < 1e300) return floor(w);
    return floor(pymin(w, r * (w / D)));
}

// One node.  The link kernels hand over sending/receiving flows in *node-major* order
// (nm_s / nm_r: slot k of node n at index n*stride + k, replica fastest), so this pass reads
// contiguous, coalesced records; it answers by storing every link's inflow[t] / outflow[t] straight
// into the history rows (link-major).  Both directions of the link<->node incidence are therefore
// crossed by *stores* (scattered 8-byte writes nobody waits for) and all loads are coalesced.
// M > 0: slot count known at compile time (loops unrolled, everything in registers);
// M == 0: generic path for rare high-degree nodes (arrays in local memory).
template <int M, bool R1, bool ROUTED>
__device__ __forceinline__ void node_body(const Ctx& c, int node, int rep, int m_dyn, int kind, int tf_mode,
                                          int dem_row, int tf_ptr, const L2Pol pol) {
    constexpr int CAP = M ? M : PNS_MAX_DEGREE;
    const int m = M ? M : m_dyn;
    const int R = R1 ? 1 : c.n.replicas;
    const size_t base = (size_t)node * c.n.nd_stride;
    double s[CAP], r[CAP];
    int in_link[CAP];                                                       // incoming link column of each slot
    // The m(m-1) routed fractions are used one by one deep inside the column loop; read there, each costs its own
    // memory round trip (measured: a third of the kernel's stall samples, 6 us of serial latency per CTA).  With a
    // compile-time slot count they are fetched here in one batch with the hand-over loads.
    constexpr bool PRELOAD = ROUTED && M > 0;
    double tfv[PRELOAD ? M * (M - 1) : 1];
    if (PRELOAD && tf_mode == 2) {
        const double* t0 = c.s.tf_routed + (size_t)tf_ptr * R + rep;
#pragma unroll
        for (int k = 0; k < M * (M - 1); ++k) tfv[k] = t0[(size_t)k * R];
    }
    if (M == 4) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(c.n.nd_in_link + base));
        in_link[0] = v.x; in_link[1 % CAP] = v.y; in_link[2 % CAP] = v.z; in_link[3 % CAP] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < m; ++i) in_link[i] = __ldg(c.n.nd_in_link + base + i);
    }
    if (R1 && M == 4) {
        const double2* ps = reinterpret_cast<const double2*>(c.s.nm_s + base);
        const double2* pr = reinterpret_cast<const double2*>(c.s.nm_r + base);
        const double2 a0 = ld_keep<1>(ps, pol), a1 = ld_keep<1>(ps + 1, pol);
        const double2 b0 = ld_keep<1>(pr, pol), b1 = ld_keep<1>(pr + 1, pol);
        s[0] = a0.x; s[1 % CAP] = a0.y; s[2 % CAP] = a1.x; s[3 % CAP] = a1.y;
        r[0] = b0.x; r[1 % CAP] = b0.y; r[2 % CAP] = b1.x; r[3 % CAP] = b1.y;
    } else {
#pragma unroll
        for (int i = 0; i < m; ++i) {
            s[i] = R1 ? ld_keep<1>(c.s.nm_s + base + i, pol) : c.s.nm_s[(base + i) * R + rep];
            r[i] = R1 ? ld_keep<1>(c.s.nm_r + base + i, pol) : c.s.nm_r[(base + i) * R + rep];
        }
    }
    double v_cout = 0.0, v_cin = 0.0;                                       // counters of the virtual links
    if (dem_row >= 0) {                                                     // slot 0 is the virtual O/D link pair
        s[0] = c.n_demand[(size_t)dem_row * R + rep];                       // node.py:176
        r[0] = 1e6;                                                         // node.py:186
        const size_t vin = (size_t)(c.n.n_links + 2 * dem_row) * R + rep;
        v_cout = c.n_coutp[vin]; v_cin = c.n_cinp[vin + R];                 // fetched with the rest of the batch
    }
    bool negative = false;
#pragma unroll
    for (int i = 0; i < m; ++i) negative |= (s[i] < 0.0) | (r[i] < 0.0);
    if (negative) atomicOr(c.s.err + rep, PNS_ERR_NEG_NODE_FLOW);

    // Nothing is sent into this node: every flow through it is zero, and the rows of step t already hold
    // zeros (pns_state_init; see the row contract at pns_node_flows in the header), so neither the solve
    // nor the eight scattered stores are needed.  Most nodes of a sparsely occupied network end here.
    bool any_flow = false;
#pragma unroll
    for (int i = 0; i < m; ++i) any_flow |= s[i] != 0.0;
    if (!any_flow) {
        if (dem_row >= 0) {               // the cumulative counts of the virtual links still carry forward
            const size_t vin = (size_t)(c.n.n_links + 2 * dem_row) * R + rep, vout = vin + R;
            c.n_cout[vin] = v_cout;
            c.n_cin[vout] = v_cin;
        }
        return;
    }
    double q_out[CAP], q_in[CAP];
    if (kind == 0) {
        // OneToOneNode.solve (node.py:230-242): exactly two slots
        const double a = fmin(s[0], r[CAP > 1 ? 1 : 0]), b = fmin(s[CAP > 1 ? 1 : 0], r[0]);
        q_out[0] = a; q_out[CAP > 1 ? 1 : 0] = b;
        q_in[0] = b;  q_in[CAP > 1 ? 1 : 0] = a;
    } else {
        // RegularNode.solve, 'classic' (node.py:272-300), one column (outgoing slot j) at a time:
        // W[i][j] = P[i][j]*s[i], D[j] = sum_i W[i][j] in slot order, f = floor(min(W, r[j]*(W/D[j]))).
        // tf_mode 0: P = phi = 1/(m-1) (network.py:269-271); otherwise P[i][j] = tf[i*(m-1) + (j<i ? j : j-1)].
        const double* tf = nullptr;
        size_t ts = 1;
        if (ROUTED && tf_mode == 2) {
            tf = c.s.tf_routed + (size_t)tf_ptr * R + rep;      // written by k_route_fractions for this step
            ts = (size_t)R;
        } else if (tf_mode == 1) {
            tf = c.s.tf_static + tf_ptr;
        }
        const double phi = 1.0 / (double)(m - 1);
        if (tf == nullptr) {
#pragma unroll
            for (int i = 0; i < m; ++i) s[i] = phi * s[i];        // from here on s[] holds W[i][*]
        }
#pragma unroll
        for (int i = 0; i < m; ++i) q_out[i] = 0.0;
#pragma unroll
        for (int j = 0; j < m; ++j) {
            double w[CAP];
            double D = 0.0;               // np.sum(axis=0): rows added in order (the diagonal adds an exact 0)
#pragma unroll
            for (int i = 0; i < m; ++i) {
                w[i] = 0.0;
                if (i == j) continue;
                if (PRELOAD && tf_mode == 2) w[i] = tfv[(i * (M - 1) + (j < i ? j : j - 1)) % (PRELOAD ? M * (M - 1) : 1)] * s[i];
                else w[i] = tf ? tf[(size_t)(i * (m - 1) + (j < i ? j : j - 1)) * ts] * s[i] : s[i];
                D = D + w[i];
            }
            D = D != 0.0 ? D : 1e-5;
            double in_j = 0.0;
#pragma unroll
            for (int i = 0; i < m; ++i) {
                if (i == j) continue;
                const double f = turn_flow(w[i], r[j], D);
                q_out[i] += f;
                in_j += f;
            }
            q_in[j] = fmax(0.0, in_j);
        }
#pragma unroll
        for (int i = 0; i < m; ++i) q_out[i] = fmax(0.0, q_out[i]);
    }
    // Node.update_links (node.py:146-162): slot i's incoming link gets outflow[t] = q_out[i], its reverse
    // (the slot's outgoing link) gets inflow[t] = q_in[i].  Scattered 8-byte stores straight into the
    // history rows: nothing waits on them, and the link pass reads its own column back coalesced.
#pragma unroll
    for (int i = 0; i < m; ++i) {
        const int col = in_link[i];
        c.n_outflow[(size_t)col * R + rep] = q_out[i];
        c.n_inflow[(size_t)(col ^ 1) * R + rep] = q_in[i];
    }
    if (dem_row >= 0) {
        // the virtual links have no link thread: keep their counters here (link.py:19-25)
        const size_t vin = (size_t)(c.n.n_links + 2 * dem_row) * R + rep, vout = vin + R;
        c.n_cout[vin] = v_cout + q_out[0];
        c.n_cin[vout] = v_cin + q_in[0];
    }
}

template <bool R1, bool ROUTED>
__device__ __noinline__ void node_body_generic(const Ctx& c, int node, int rep, int m, int kind, int tf_mode,
                                               int dem_row, int tf_ptr) {
    node_body<0, R1, ROUTED>(c, node, rep, m, kind, tf_mode, dem_row, tf_ptr, l2_policies());
}

// Node.assign_flows / solve (node.py:164-300) + turning fractions (path_finder.py:591-715)
// ROUTED: some node takes its fractions from the route-choice model (tf_routed)
template <bool R1, bool ROUTED>
__global__ void __launch_bounds__(PNS_NODE_BLOCK, PNS_NODE_MIN_BLOCKS) k_node_flows(const __grid_constant__ Ctx c) {
    const int R = R1 ? 1 : c.n.replicas;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)c.n.n_nodes * R) return;
    const int node = R1 ? (int)gid : (int)(gid / R);
    const int rep = R1 ? 0 : (int)(gid % R);
    PNS_PDL_TRIGGER();
    const L2Pol pol = l2_policies();
    const int4 meta = R1 ? ld_meta(reinterpret_cast<const int4*>(c.n.nd_meta) + node, pol)
                         : __ldg(reinterpret_cast<const int4*>(c.n.nd_meta) + node);   // {-, m|kind|mode, demand row, tf offset}
    const int m = meta.y & 0xff, kind = (meta.y >> 8) & 0xff, tf_mode = (meta.y >> 16) & 0xff;
    PNS_PDL_WAIT();
#ifndef PNS_NODE_PF_AHEAD
#define PNS_NODE_PF_AHEAD 113664   // nodes: three quarters of a resident wave of 148 SMs x 1024 threads
#endif
    if (R1 && PNS_NODE_PF_AHEAD > 0) {
        // fewer than two resident waves: the first wave pulls the records of the second into L2
        const size_t na = gid + (size_t)PNS_NODE_PF_AHEAD;
        if (na < (size_t)c.n.n_nodes) {
            prefetch_l2(reinterpret_cast<const int4*>(c.n.nd_meta) + na);
            prefetch_l2(c.s.nm_s + na * c.n.nd_stride);
            prefetch_l2(c.s.nm_r + na * c.n.nd_stride);
        }
    }
    if (kind == 2) return;       // 'optimal' node model: k_node_lp
    switch (m) {
        case 0: case 1: break;   // isolated node / dead end without any turn
        case 2: node_body<2, R1, ROUTED>(c, node, rep, 2, kind, tf_mode, meta.z, meta.w, pol); break;
        case 3: node_body<3, R1, ROUTED>(c, node, rep, 3, kind, tf_mode, meta.z, meta.w, pol); break;
        case 4: node_body<4, R1, ROUTED>(c, node, rep, 4, kind, tf_mode, meta.z, meta.w, pol); break;
        default: node_body_generic<R1, ROUTED>(c, node, rep, m, kind, tf_mode, meta.z, meta.w); break;
    }
}

// =================================================================================================
// RegularNode.solve(type='optimal') (node.py:249-271): nodes of kind 2 solve the linear program of pns_lp.cuh, one
// warp per node and replica (a sequential "thread" in the host build), then floor the turn flows and add them up
// per link like the classic model (`A_ub @ np.floor(res.x)`, node.py:266-268).  Loads, the zero-flow shortcut (no
// sending flow: X_i <= 0 forces x = 0), the stores into the history rows and the counters of the virtual links are
// those of node_body.
constexpr int kLpHeader = 24;   // doubles in front of the solver's scratch: s[8], r[8], uniform fraction
template <int LANES>
__device__ __forceinline__ void node_lp_body(const Ctx& c, int node, int rep, int lane, double* scratch) {
    const int R = c.n.replicas;
    const int4 meta = __ldg(reinterpret_cast<const int4*>(c.n.nd_meta) + node);
    const int m = meta.y & 0xff, tf_mode = (meta.y >> 16) & 0xff, dem_row = meta.z, tf_ptr = meta.w;
    const size_t base = (size_t)node * c.n.nd_stride;
    double *sv = scratch, *rv = scratch + 8;
    for (int i = lane; i < m; i += LANES) {
        sv[i] = c.s.nm_s[(base + i) * R + rep];
        rv[i] = c.s.nm_r[(base + i) * R + rep];
    }
    double v_cout = 0.0, v_cin = 0.0;
    const size_t vin = dem_row >= 0 ? (size_t)(c.n.n_links + 2 * dem_row) * R + rep : 0, vout = vin + R;
    if (dem_row >= 0 && lane == 0) {
        sv[0] = c.n_demand[(size_t)dem_row * R + rep];                       // node.py:176
        rv[0] = 1e6;                                                         // node.py:186
        v_cout = c.n_coutp[vin]; v_cin = c.n_cinp[vout];
    }
    if (lane == 0) scratch[16] = 1.0 / (double)(m - 1);                      // network.py:269-271
    PNS_LP_SYNC();
    bool negative = false, any_flow = false;
    for (int i = 0; i < m; ++i) { negative |= (sv[i] < 0.0) | (rv[i] < 0.0); any_flow |= sv[i] != 0.0; }
    if (negative && lane == 0) atomicOr(c.s.err + rep, PNS_ERR_NEG_NODE_FLOW);
    if (negative || !any_flow) {
        if (dem_row >= 0 && lane == 0) { c.n_cout[vin] = v_cout; c.n_cin[vout] = v_cin; }
        return;
    }
    const double* phi = scratch + 16;
    size_t phi_stride = 0;
    if (tf_mode == 2) { phi = c.s.tf_routed + (size_t)tf_ptr * R + rep; phi_stride = (size_t)R; }
    else if (tf_mode == 1) { phi = c.s.tf_static + tf_ptr; phi_stride = 1; }
    double* x = nullptr;
    double objective = 0.0;
    const int info = pns::lp_node_solve<LANES>(m, sv, rv, phi, phi_stride, c.n.lp_w, lane, scratch + kLpHeader, &x,
                                               &objective);
    if ((info & (pns::LP_UNBOUNDED | pns::LP_PIVOT_LIMIT)) && lane == 0) atomicOr(c.s.err + rep, PNS_ERR_LP_FAILED);
#ifdef PNS_HOST_EMULATION
    if ((info & (pns::LP_UNBOUNDED | pns::LP_PIVOT_LIMIT)) && getenv("PNS_LP_DEBUG")) {    // test build: show the program
        fprintf(stderr, "LP failed: info %x node %d rep %d t %d m %d\n", info, node, rep, c.t, m);
        for (int i = 0; i < m; ++i) fprintf(stderr, " s %.17g r %.17g\n", sv[i], rv[i]);
        for (int e = 0; e < m * (m - 1); ++e) fprintf(stderr, " phi %.17g\n", phi[(size_t)e * phi_stride]);
    }
#endif
    if (c.io.lp_x)
        for (int e = lane; e < m * (m - 1); e += LANES) c.io.lp_x[(size_t)(tf_ptr + e) * R + rep] = x[e];
    double q0_out = 0.0, q0_in = 0.0;
    for (int i = lane; i < m; i += LANES) {
        double q_out = 0.0, q_in = 0.0;
        for (int k = 0; k < m - 1; ++k) q_out += floor(x[i * (m - 1) + k]);              // row i of A_ub
        for (int k = 0; k < m; ++k)
            if (k != i) q_in += floor(x[k * (m - 1) + (i < k ? i : i - 1)]);             // row m + i of A_ub
        q_out = fmax(0.0, q_out); q_in = fmax(0.0, q_in);
        const int col = __ldg(c.n.nd_in_link + base + i);
        c.n_outflow[(size_t)col * R + rep] = q_out;                                      // node.py:146-162
        c.n_inflow[(size_t)(col ^ 1) * R + rep] = q_in;
        if (i == 0) { q0_out = q_out; q0_in = q_in; }
    }
    if (dem_row >= 0 && lane == 0) { c.n_cout[vin] = v_cout + q0_out; c.n_cin[vout] = v_cin + q0_in; }
}

__host__ __device__ inline size_t node_lp_warp_doubles(int max_m) {
    return (size_t)kLpHeader + pns::lp_scratch_bytes(max_m) / sizeof(double);
}

#ifdef PNS_HOST_EMULATION
__global__ void k_node_lp(const Ctx c) {
    static thread_local double* scratch = nullptr;
    if (!scratch) scratch = (double*)malloc(node_lp_warp_doubles(PNS_MAX_DEGREE) * sizeof(double));
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)c.n.n_lp_nodes * c.n.replicas) return;
    node_lp_body<1>(c, c.n.lp_nodes[gid / c.n.replicas], (int)(gid % c.n.replicas), 0, scratch);
}
// batch of stand-alone programs (pns_lp_solve)
struct LpBatch { int m, n; const double *s, *r, *phi; double w; double *x, *objective; int32_t* info; };
__global__ void k_lp_batch(const LpBatch b) {
    static thread_local double* scratch = nullptr;
    if (!scratch) scratch = (double*)malloc(node_lp_warp_doubles(PNS_MAX_DEGREE) * sizeof(double));
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (size_t)b.n) return;
    const int E = b.m * (b.m - 1);
    double* x = nullptr;
    b.info[k] = pns::lp_node_solve<1>(b.m, b.s + k * b.m, b.r + k * b.m, b.phi + k * E, 1, b.w, 0, scratch, &x,
                                      b.objective + k);
    for (int e = 0; e < E; ++e) b.x[k * E + e] = x[e];
}
#else
__global__ void __launch_bounds__(256) k_node_lp(const __grid_constant__ Ctx c) {
    extern __shared__ double lp_smem[];
    const unsigned warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
    const unsigned R = (unsigned)c.n.replicas;
    PNS_PDL_TRIGGER();
    PNS_PDL_WAIT();
    if (warp >= (size_t)c.n.n_lp_nodes * R) return;
    const int node = __ldg(c.n.lp_nodes + (unsigned)(warp / R));
    node_lp_body<32>(c, node, (int)(warp % R), (int)lane, lp_smem + warp_in_cta * node_lp_warp_doubles(c.n.lp_max_m));
}
struct LpBatch { int m, n; const double *s, *r, *phi; double w; double *x, *objective; int32_t* info; };
__global__ void __launch_bounds__(256) k_lp_batch(const LpBatch b) {
    extern __shared__ double lp_smem[];
    const unsigned warp_in_cta = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t k = (size_t)blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
    if (k >= (size_t)b.n) return;
    const int E = b.m * (b.m - 1);
    double* x = nullptr;
    double objective = 0.0;
    const int info = pns::lp_node_solve<32>(b.m, b.s + k * b.m, b.r + k * b.m, b.phi + k * E, 1, b.w, (int)lane,
                                            lp_smem + warp_in_cta * node_lp_warp_doubles(b.m), &x, &objective);
    for (int e = (int)lane; e < E; e += 32) b.x[k * E + e] = x[e];
    if (lane == 0) { b.objective[k] = objective; b.info[k] = info; }
}
#endif

#ifndef PNS_HOST_EMULATION
// =================================================================================================
// Batched replicas: the node model with one *warp per slot*.  A CTA is one node for 32 replicas; warp j owns slot j
// (outgoing link j, receiving flow r[j]): it forms column j of the classic model -- W[i][j] = P[i][j] s[i],
// D[j] = sum_i W[i][j] in slot order, f[i][j] = floor(min(W, r[j] (W/D))) -- and the inflow of its outgoing link,
// q_in[j] = sum_i f[i][j]; the flows meet in shared memory, and after one barrier warp j adds up row j, the
// outflow of its incoming link, q_out[j] = sum_j' f[j][j'] (same summation orders as node_body).  Against one
// thread per node this is a quarter of the dependent arithmetic per thread, a third of the registers and four
// times the warps: the thread-per-node kernel spent 25 us at 4096 replicas on 25 MB of traffic (9 us its loads,
// 7 us the solve, 8 us the routed fractions), its duration almost independent of the replica count.
// W = slots per node the launch provides warps for (pns_net.nd_stride: 4 or 8); warps beyond the node's slot count
// leave at once, the others meet at a named barrier sized to the node.
// Body for a node with M slots (compile-time: every loop is exact and every small array lives in registers).
#ifndef PNS_REP_L2
#define PNS_REP_L2 0   // L2 eviction hints in the batched kernels (see k_link_rep): measured no gain, off
#endif
#if PNS_REP_L2
#define NLD_ONCE(a) ldh((a), pol.once)
#define NST_KEEP(a, v) sth((a), (v), pol.keep)
#else
#define NLD_ONCE(a) (*(a))
#define NST_KEEP(a, v) (*(a) = (v))
#endif
template <bool ROUTED, int W, int M>
__device__ __forceinline__ void node_cols_body(const Ctx& c, double (*sh_f)[W][32], int node, int j, int lane, int rep,
                                               bool valid, int kind, int tf_mode, int dem_row, int tf_ptr,
                                               int barrier_id = 1) {
    const int R = c.n.replicas;
    const size_t base = (size_t)node * c.n.nd_stride;
    const int col = __ldg(c.n.nd_in_link + base + j);
    PNS_PDL_WAIT();
    // ---- one batch of loads: all sending flows, this slot's receiving flow, column j of the fractions ----
    const L2Pol pol = l2_policies();           // PNS_REP_L2: hand-over records read once, the answer kept
    (void)pol;
    double s[M];
#pragma unroll
    for (int i = 0; i < M; ++i) s[i] = NLD_ONCE(c.s.nm_s + (base + i) * R + rep);
    double r_j = NLD_ONCE(c.s.nm_r + (base + j) * R + rep);
    double r_other = (M == 2 && kind == 0) ? c.s.nm_r[(base + (1 - j)) * R + rep] : 0.0;
    double P[M];
    if (kind != 0) {
        const double phi = 1.0 / (double)(M - 1);
        const double* tf = nullptr;
        size_t ts = 1;
        if (ROUTED && tf_mode == 2) { tf = c.s.tf_routed + (size_t)tf_ptr * R + rep; ts = (size_t)R; }
        else if (tf_mode == 1) tf = c.s.tf_static + tf_ptr;
#pragma unroll
        for (int i = 0; i < M; ++i)
            P[i] = i != j ? (tf ? tf[(size_t)(i * (M - 1) + (j < i ? j : j - 1)) * ts] : phi) : 0.0;
    }
    double v_cout = 0.0, v_cin = 0.0;          // counters of the virtual links (kept by warp 0, their slot)
    size_t vin = 0;
    if (dem_row >= 0) {                        // slot 0 is the virtual O/D link pair
        s[0] = c.n_demand[(size_t)dem_row * R + rep];                       // node.py:176
        vin = (size_t)(c.n.n_links + 2 * dem_row) * R + rep;
        if (j == 0) r_j = 1e6;                                              // node.py:186
        if (M == 2 && kind == 0 && j == 1) r_other = 1e6;
        if (j == 0) { v_cout = c.n_coutp[vin]; v_cin = c.n_cinp[vin + R]; }
    }
    double s_j = 0.0, s_other = 0.0;
#pragma unroll
    for (int i = 0; i < M; ++i) { if (i == j) s_j = s[i]; if (i == 1 - j) s_other = s[i]; }
    if (valid && ((s_j < 0.0) | (r_j < 0.0))) atomicOr(c.s.err + rep, PNS_ERR_NEG_NODE_FLOW);
    bool any_flow = false;
#pragma unroll
    for (int i = 0; i < M; ++i) any_flow |= s[i] != 0.0;
    double q_in = 0.0, q_out = 0.0;
    if (M == 2 && kind == 0) {
        // OneToOneNode.solve (node.py:230-242): exactly two slots
        q_out = fmin(s_j, r_other);
        q_in = fmin(s_other, r_j);
    } else {
        // RegularNode.solve, 'classic' (node.py:272-300), column j
        double D = 0.0;                        // np.sum(axis=0): rows added in order (the diagonal adds an exact 0)
        double w[M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
            w[i] = 0.0;
            if (i == j) continue;
            w[i] = P[i] * s[i];
            D = D + w[i];
        }
        D = D != 0.0 ? D : 1e-5;
        double in_j = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) {
            if (i == j) continue;
            const double f = turn_flow(w[i], r_j, D);
            sh_f[j][i][lane] = f;
            in_j += f;
        }
        q_in = fmax(0.0, in_j);
        asm volatile("bar.sync %0, %1;" :: "r"(barrier_id), "r"(32 * M) : "memory");
#pragma unroll
        for (int jj = 0; jj < M; ++jj)
            if (jj != j) q_out += sh_f[jj][j][lane];
        q_out = fmax(0.0, q_out);
    }
    if (!valid) return;
    // Node.update_links (node.py:146-162); a node that receives no sending flow stores nothing (rows are zero by
    // contract, see pns_node_flows), only the counters of its virtual links carry forward
    if (any_flow) {
        NST_KEEP(c.n_outflow + (size_t)col * R + rep, q_out);
        NST_KEEP(c.n_inflow + (size_t)(col ^ 1) * R + rep, q_in);
    }
    if (dem_row >= 0 && j == 0) {
        // the virtual links have no link thread: slot 0's flows extend their counters (link.py:19-25)
        c.n_cout[vin] = v_cout + (any_flow ? q_out : 0.0);
        c.n_cin[vin + R] = v_cin + (any_flow ? q_in : 0.0);
    }
}

template <bool ROUTED, int W>
__global__ void __launch_bounds__(32 * W) k_node_cols(const __grid_constant__ Ctx c) {
    __shared__ double sh_f[W][W][32];          // f[j][i][lane]: flow from slot i into slot j's outgoing link
    const int R = c.n.replicas;
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
    const int rep_raw = (int)blockIdx.x * 32 + lane;
    const bool valid = rep_raw < R;
    const int rep = valid ? rep_raw : R - 1;
    PNS_PDL_TRIGGER();
    // W == 8 (some node has more than 4 slots): registers are granted per CTA, so a 256-thread CTA for a node with
    // 3 or 4 slots would hold half of them idle.  Nodes with at most 4 slots therefore share a CTA in twos (warps
    // 0-3 / 4-7, named barriers 1 / 2, the two halves of sh_f); pns_net.nd_cols_order lists those nodes first.
    int node, j, half = 0;
    if (W == 8 && c.n.nd_cols_order) {
        const int n_small = c.n.n_nodes_small, small_ctas = (n_small + 1) >> 1;
        if ((int)blockIdx.y < small_ctas) {
            half = warp >> 2;
            const int k = 2 * (int)blockIdx.y + half;
            if (k >= n_small) return;
            node = __ldg(c.n.nd_cols_order + k);
            j = warp & 3;
        } else {
            node = __ldg(c.n.nd_cols_order + n_small + ((int)blockIdx.y - small_ctas));
            j = warp;
        }
    } else {
        node = (int)blockIdx.y;
        j = warp;
    }
    const int4 meta = __ldg(reinterpret_cast<const int4*>(c.n.nd_meta) + node);   // {-, m|kind|mode, demand row, tf offset}
    const int m = meta.y & 0xff, kind = (meta.y >> 8) & 0xff, tf_mode = (meta.y >> 16) & 0xff;
    if (m < 2 || j >= m || kind == 2) return;  // dead end / isolated node; warps without a slot; k_node_lp's nodes
    // a node of <= 4 slots uses a 4 x 4 x 32 block of sh_f (the first or the second half of the array)
    double (*sh4)[4][32] = reinterpret_cast<double (*)[4][32]>(&sh_f[0][0][0] + (size_t)half * 4 * 4 * 32);
#define PNS_COLS4(MM) node_cols_body<ROUTED, 4, MM>(c, sh4, node, j, lane, rep, valid, kind, tf_mode, meta.z, meta.w, 1 + half)
#define PNS_COLS(MM) node_cols_body<ROUTED, W, MM>(c, sh_f, node, j, lane, rep, valid, kind, tf_mode, meta.z, meta.w)
    switch (m) {
        case 2: PNS_COLS4(2); break;
        case 3: PNS_COLS4(3); break;
        case 4: PNS_COLS4(4); break;
        default:
            if (W > 4) {
                switch (m) {
                    case 5: PNS_COLS((W > 4 ? 5 : 2)); break;
                    case 6: PNS_COLS((W > 4 ? 6 : 2)); break;
                    case 7: PNS_COLS((W > 4 ? 7 : 2)); break;
                    default: PNS_COLS((W > 4 ? 8 : 2)); break;
                }
            }
            break;
    }
#undef PNS_COLS
#undef PNS_COLS4
}
#endif

#ifndef PNS_HOST_EMULATION
// =================================================================================================
// Single-replica fast path: one thread per *directed link*; the two directions of a corridor sit
// in adjacent lanes and trade the three values they need from each other (pedestrians, density,
// sending flow) with warp shuffles.  Same arithmetic as link_pair_body, half the critical path
// per thread and no cross-direction state to keep in registers.  (The host-emulation test build
// runs the pair-per-thread kernel above instead; the GPU parity tests cover this one.)
#ifndef PNS_LANE_BULK
#define PNS_LANE_BULK 0   // k_link_lane: stage the first batch of rows in shared memory with cp.async.bulk (experiment)
#endif
#if PNS_LANE_BULK
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif
template <int PHASE, int MODE, bool ONECLASS>
__global__ void __launch_bounds__(PNS_LANE_BLOCK, PNS_LANE_MIN_BLOCKS) k_link_lane(const __grid_constant__ Ctx c) {
    constexpr bool upd = (PHASE & PH_UPDATE) != 0, flw = (PHASE & PH_FLOWS) != 0;
    constexpr unsigned FULL = 0xffffffffu;
    // launch order (pns_net.lane_order): CTA i works on link block order[i]; blocks likely to hold long
    // sampler walks come first so that they overlap the rest of the grid
    unsigned bx = blockIdx.x, n_link_ctas = gridDim.x;
    if (flw && c.route_blocks > 0) {                     // route choice rides along (see Ctx::route_blocks)
        if (bx < (unsigned)c.route_blocks) {             // one warp per row
            const unsigned row = bx * (unsigned)(PNS_LANE_BLOCK / 32) + (threadIdx.x >> 5);
            if (row < (unsigned)route_rows(c)) route_warp(c, c.route_t, (int)row, threadIdx.x & 31u);
            return;
        }
        bx -= (unsigned)c.route_blocks;
        n_link_ctas -= (unsigned)c.route_blocks;
    }
    const bool ordered = c.n.lane_order != nullptr && c.n.lane_order_block == PNS_LANE_BLOCK;
    const unsigned blk = ordered ? (unsigned)__ldg(c.n.lane_order + bx) : bx;
    const unsigned pf_slot = bx + (unsigned)PNS_PF_AHEAD_CTAS;       // the CTA this one prefetches for (below)
    const unsigned pf_blk = pf_slot < n_link_ctas ? (ordered ? (unsigned)__ldg(c.n.lane_order + pf_slot) : pf_slot) : 0xffffffu;
    const unsigned gid = blk * blockDim.x + threadIdx.x;
    const bool valid = gid < (unsigned)c.n.n_links;           // whole pairs: a lane and its partner agree
    const int l = valid ? (int)gid : 0;
    const size_t e = (size_t)l;
    const int tau = c.t_flows - 1;
    PNS_PDL_TRIGGER();
    // single-class networks: parameters are kernel-parameter constants, no table loads at all
    const LinkP& p = ONECLASS ? c.n.class0 : c.n.classes[class_of(c, l, 0)];
    const L2Pol pol = l2_policies();
    const int2 slots = PNS_L2_STAGE >= 1 ? ldh_nc(reinterpret_cast<const int2*>(c.n.lk_slots) + l, pol.keep)
                                         : __ldg(reinterpret_cast<const int2*>(c.n.lk_slots) + l);   // {sending slot, receiving slot}
    const int fftau = p.fftau, swtau = p.swtau;
    PNS_PDL_WAIT();             // everything above is static; below reads what the previous kernel wrote
#if PNS_LANE_BULK
    // Experiment: the first batch of rows of a full block of the fused launch of a
    // single-class network comes in through the bulk-copy engine -- one thread issues up to thirteen 512 / 256 byte
    // copies into shared memory, the block waits on one mbarrier -- instead of thirteen loads per thread.
    constexpr bool kBulk = upd && flw && ONECLASS;
    __shared__ alignas(128) double bk64[kBulk ? 10 : 1][PNS_LANE_BLOCK];
    __shared__ alignas(128) float bk32[kBulk ? 3 : 1][PNS_LANE_BLOCK];
    __shared__ alignas(8) uint64_t bk_bar;
    const bool bulk = kBulk && (size_t)(blk + 1) * PNS_LANE_BLOCK <= (size_t)c.n.n_links;
    if (bulk) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bk_bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const size_t e0 = (size_t)blk * PNS_LANE_BLOCK;
            const bool win = c.u_tt_old != nullptr;
            const uint32_t b64 = PNS_LANE_BLOCK * 8u, b32 = PNS_LANE_BLOCK * 4u;
            const uint32_t total = b64 * (7u + (c.c0_coulag ? 1u : 0u) + (c.c0_pre0 ? 1u : 0u) + (c.c0_pre1 ? 1u : 0u)) +
                                   b32 * (2u + (win ? 1u : 0u));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bk_bar)), "r"(total) : "memory");
            bulk_g2s(bk64[0], c.s.gate + e0, b64, &bk_bar);
            bulk_g2s(bk64[1], c.n_cinp + e0, b64, &bk_bar);
            bulk_g2s(bk64[2], c.n_coutp + e0, b64, &bk_bar);
            bulk_g2s(bk64[3], c.n_outflow + e0, b64, &bk_bar);
            bulk_g2s(bk64[4], c.n_inflow + e0, b64, &bk_bar);
            bulk_g2s(bk64[5], c.f_sndp + e0, b64, &bk_bar);
            bulk_g2s(bk64[6], c.f_rcvp + e0, b64, &bk_bar);
            if (c.c0_coulag) bulk_g2s(bk64[7], c.c0_coulag + e0, b64, &bk_bar);
            if (c.c0_pre0) bulk_g2s(bk64[8], c.c0_pre0 + e0, b64, &bk_bar);
            if (c.c0_pre1) bulk_g2s(bk64[9], c.c0_pre1 + e0, b64, &bk_bar);
            bulk_g2s(bk32[0], c.u_num_prev + e0, b32, &bk_bar);
            bulk_g2s(bk32[1], c.s.runsum + e0, b32, &bk_bar);
            if (win) bulk_g2s(bk32[2], c.u_tt_old + e0, b32, &bk_bar);
        }
    }
#else
    constexpr bool bulk = false;
#endif
    double gate_ld = 0.0;
    if (!bulk) gate_ld = ld_keep<2>(c.s.gate + e, pol);
    // ---- batch of independent loads --------------------------------------------------------
    double din = 0, dout = 0;
    float np_ = 0, rs = 0, tt_old = 0;
    const bool windowed = c.u_tt_old != nullptr;
    double cin_prev = 0, cou_prev = 0;
    if (upd && !bulk) {
        // Node.update_links for this link (node.py:146-162): the node pass wrote inflow[t] / outflow[t]
        cin_prev = ld_once<3>(c.n_cinp + e, pol); cou_prev = ld_once<3>(c.n_coutp + e, pol);
        np_ = ld_once<3>(c.u_num_prev + e, pol); rs = ld_keep<2>(c.s.runsum + e, pol);
        if (windowed) tt_old = ld_once<3>(c.u_tt_old + e, pol);
        dout = c.n_outflow[e]; din = c.n_inflow[e];
    }
    double cin_tau = 0, cou_tau = 0, snd_prev = 0, rcv_prev = 0, cou_lag = 0;
    LinkNow me;
    me.num = 0; me.dens = 0; me.avg_tt = 0;
    if (flw && !bulk) {
        if (!upd) { cin_tau = c.f_cin[e]; cou_tau = c.f_cou[e]; }
        snd_prev = ld_once<3>(c.f_sndp + e, pol); rcv_prev = ld_once<3>(c.f_rcvp + e, pol);
        if (ONECLASS) {
            if (c.c0_coulag) cou_lag = ld_once<3>(c.c0_coulag + e, pol);
        } else {
            const int lag_i = tau + 1 - swtau;
            if (lag_i >= 0) cou_lag = ld_once<3>(H64(c, PNS_F64_CUM_OUTFLOW, lag_i) + e, pol);
        }
        if (!upd) { me.num = c.f_num[e]; me.dens = c.f_dens[e]; me.avg_tt = c.f_avg[e]; }
    }
    // The arrival row cumulative_inflow[tau+1-lag] depends on the travel-time lag computed below; in
    // free flow the lag is the free-flow lag or one less (speed noise), so fetch both rows now and
    // fall back to a dependent load only when the link is congested.
    int pre_i0 = -1, pre_i1 = -1;
    double pre_v0 = 0.0, pre_v1 = 0.0;
    const double *pre_row0 = nullptr, *pre_row1 = nullptr;
    if (flw) {
        if (ONECLASS) {
            pre_i0 = c.c0_pre_i0; pre_i1 = c.c0_pre_i1; pre_row0 = c.c0_pre0; pre_row1 = c.c0_pre1;
        } else if (tau >= fftau) {
            pre_i0 = max(0, tau + 1 - fftau);
            pre_row0 = H64(c, PNS_F64_CUM_INFLOW, pre_i0);
            if (fftau > 1) {
                pre_i1 = max(0, tau + 2 - fftau);
                pre_row1 = H64(c, PNS_F64_CUM_INFLOW, pre_i1);
            }
        }
        if (!bulk) {
            if (pre_row0) pre_v0 = ld_once<3>(pre_row0 + e, pol);
            if (pre_row1) pre_v1 = ld_keep<3>(pre_row1 + e, pol);   // next step's pre_v0
        }
    }
    // ---- software prefetch for later CTAs ----------------------------------------------------
    // A thread spends most of its memory time waiting for the batch above to come back from DRAM.
    // CTAs are dispatched in index order, so while these loads are in flight the kernel asks L2 to
    // fetch the same columns for the links PNS_PF_AHEAD_CTAS CTAs further on (most of a resident wave,
    // a few microseconds ahead); their batch then hits in L2.
    if (PNS_PF_AHEAD_CTAS > 0 && threadIdx.x < 32u) {
        // the CTA's first warp covers the PNS_LANE_BLOCK links of the target CTA, 16 bytes apart
        const unsigned ga = pf_blk * (unsigned)PNS_LANE_BLOCK + (unsigned)(PNS_LANE_BLOCK / 32) * threadIdx.x;
        if (pf_slot < n_link_ctas && ga < (unsigned)c.n.n_links) {
            const size_t ea = ga;
            prefetch_l2(reinterpret_cast<const int2*>(c.n.lk_slots) + ea);
            prefetch_l2(c.s.gate + ea);
            if (upd) {
                prefetch_l2(c.n_cinp + ea); prefetch_l2(c.n_coutp + ea);
                prefetch_l2(c.u_num_prev + ea); prefetch_l2(c.s.runsum + ea);
                if (windowed) prefetch_l2(c.u_tt_old + ea);
                prefetch_l2(c.n_outflow + ea); prefetch_l2(c.n_inflow + ea);
            }
            if (flw) {
                if (!upd) { prefetch_l2(c.f_cin + ea); prefetch_l2(c.f_cou + ea); }
                prefetch_l2(c.f_sndp + ea); prefetch_l2(c.f_rcvp + ea);
                if (ONECLASS) {
                    if (c.c0_coulag) prefetch_l2(c.c0_coulag + ea);
                } else if (tau + 1 - swtau >= 0) {
                    prefetch_l2(H64(c, PNS_F64_CUM_OUTFLOW, tau + 1 - swtau) + ea);
                }
                if (pre_row0) prefetch_l2(pre_row0 + ea);
                if (pre_row1) prefetch_l2(pre_row1 + ea);
            }
        }
    }
#if PNS_LANE_BULK
    if (bulk) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bk_bar)) : "memory");
        const unsigned k = threadIdx.x;
        gate_ld = bk64[0][k];
        cin_prev = bk64[1][k]; cou_prev = bk64[2][k]; dout = bk64[3][k]; din = bk64[4][k];
        snd_prev = bk64[5][k]; rcv_prev = bk64[6][k];
        if (c.c0_coulag) cou_lag = bk64[7][k];
        if (c.c0_pre0) pre_v0 = bk64[8][k];
        if (c.c0_pre1) pre_v1 = bk64[9][k];
        np_ = bk32[0][k]; rs = bk32[1][k];
        if (windowed) tt_old = bk32[2][k];
    }
#endif
    const double gate = gate_ld;
    const double gate_rev = __shfl_xor_sync(FULL, gate, 1);
    const Area ar = link_area(c, p, e, gate);
    const uint32_t k0 = (uint32_t)c.io.seed, k1 = (uint32_t)(c.io.seed >> 32);

    if (upd) {
        cin_tau = cin_prev + din;                                           // link.py:19-25
        cou_tau = cou_prev + dout;
        if (valid) {
            st_keep<3>(c.n_cin + e, cin_tau, pol); st_keep<3>(c.n_cout + e, cou_tau, pol);     // read back by the next step
        }
        me.num = (float)((double)np_ + (din - dout));                      // link.py:134-135
        me.dens = div_by_area(me.num, ar);                                  // link.py:136
        if (c.metric) {
            // streamed runs report the network-wide pedestrian count of every step: one atomic per warp, spread
            // over slots.  Counts are whole numbers unless a gate capacity made a sending flow fractional, so the
            // warp sum is normally one integer reduction (REDUX); otherwise five double-precision shuffles.
            const float x = valid ? me.num : 0.0f;
            const int xi = (int)x;
            const bool whole = (float)xi == x && x < 6.0e7f;
            double v;
            if (__all_sync(FULL, whole)) {
                v = (double)__reduce_add_sync(FULL, xi);
            } else {
                v = (double)x;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
            }
            if ((threadIdx.x & 31u) == 0 && v != 0.0)
                atomicAdd(c.metric + (size_t)PNS_METRIC_STRIDE * ((gid >> 5) & (PNS_METRIC_SLOTS - 1)), v);
        }
        const float dens_rev = __shfl_xor_sync(FULL, me.dens, 1);
        const bool noisy = p.sigma > 0.0;
        double z = 0.0;
        if (noisy) {
            if (MODE == PNS_RNG_TABLE) z = c.draw_n[e];
            else {
                // every lane evaluates its quad's Philox block and keeps its own normal.  (Drawing once per
                // CTA in its first warp and handing the normals over through shared memory executes a
                // quarter of the sampler instructions but measured 1.5% slower: the CTA then lives as
                // long as its longest warp.)
                pns::DrawKey key;
                key.t = (uint32_t)c.t; key.link = (uint32_t)(l & ~1); key.replica = c.io.replica_base; key.k0 = k0; key.k1 = k1;
                double g0, g1;
                pns::normal_pair_philox(key, 4u, &g0, &g1);
                z = p.sigma * ((l & 1) ? g1 : g0);
            }
        }
        float tt;
        const float v = speed_and_travel_time(p, me.dens, is_sep(p) ? 0.0f : dens_rev, noisy, z, &tt);
        float sum = rs + tt;                                                // link.py:183-186
        if (windowed) {
            sum = sum - tt_old;
            me.avg_tt = sum / (float)c.n.window;
        } else {
            me.avg_tt = p.tt0;
        }
        if (valid) {
            st_keep<3>(c.u_num + e, me.num, pol); st_once<3>(c.u_dens + e, me.dens, pol);
            st_once<3>(c.u_speed + e, v, pol); st_once<3>(c.u_tt + e, tt, pol);
            st_once<3>(c.u_flow + e, v * me.dens, pol);                     // functions.py:97-101
            if (windowed) st_once<3>(c.u_avg + e, me.avg_tt, pol);
            st_keep<2>(c.s.runsum + e, sum, pol);
            st_once<3>(c.u_bgw + e, gate, pol);                             // link.py:188, 451-452
            if (is_sep(p)) c.u_sepw[e] = gate;
        }
    }
    if (!flw) return;
    // Links that hold pedestrians will probably smooth their outflow over four lagged inflow rows
    // (get_outflow, link.py:199-214): rows tau-lag-k of a lag that is only known deep inside the
    // sending-flow computation.  In free flow the lag is the free-flow lag or one less, so ask for the
    // five candidate rows now; by the time they are needed the dependent DRAM round trip is over.
    if (PNS_PF_TAPS && me.num > 0.0f && tau >= fftau) {
        const double* row = H64(c, PNS_F64_INFLOW, tau - fftau + 1) + e;
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (tau - fftau + 1 - k >= 0) prefetch_l1(row - (size_t)k * c.row64);
    }
    const float num_rev = __shfl_xor_sync(FULL, me.num, 1);
    pns::DrawKey key;
    key.t = (uint32_t)c.t_flows; key.link = (uint32_t)l; key.replica = c.io.replica_base; key.k0 = k0; key.k1 = k1;
    const double front = is_sep(p) ? gate : gate_rev;                       // link.py:110-126, 462-478
    SendOut s;
    double r = 0.0;
    int n3 = -1;
    if (valid) {
        // A link with nobody on it and nobody in transit (everything that entered has left) sends nothing:
        // arrived <= cin[tau] - cout[tau] = 0, so boundary = 0 and only the smoothing with the previous
        // sending flow remains (link.py:363-364).  Same result as the general path, a seventh of its
        // instructions; whole warps of such links take it on sparsely occupied networks.
        if (PNS_QUIET_FAST_PATH && me.num == 0.0f && cin_tau == cou_tau && front >= 0.0 && tau >= fftau) {
            s.kind = 0; s.n1 = 0; s.rf = 0.0f; s.sval = 0.0;
            const double f = pymin(floor(0.8 * 0.0 + 0.2 * snd_prev), 0.0);
            if (MODE != PNS_RNG_REQUEST && f < 0.0) atomicOr(c.s.err, PNS_ERR_NEG_SENDING);
            s.flow = MODE == PNS_RNG_REQUEST ? 0.0 : f;
        } else
        s = sending_flow<MODE>(c, p, e, tau, me, num_rev, ar, front, cou_tau, snd_prev, 0, key, pre_i0, pre_v0, pre_i1,
                               pre_v1);
        r = receiving_flow<MODE>(c, p, e, tau, num_rev, ar, gate, cin_tau, cou_lag, rcv_prev, key, &n3);
    } else {
        s.flow = 0; s.sval = 0; s.kind = 0; s.n1 = 0; s.rf = 0;
    }
    if (MODE == PNS_RNG_REQUEST) {
        if (valid) {
            c.io.req_kind[e] = s.kind; c.io.req_n1[e] = s.n1; c.io.req_rf[e] = s.rf;
            c.io.req_sval[e] = s.sval; c.io.req_n3[e] = n3;
        }
        return;
    }
    const double s_rev = __shfl_xor_sync(FULL, s.flow, 1);
    if (valid) {
        // cal_receiving_flow_with_reverse (link.py:407-416; separators ignore the reverse flow, :509-512)
        const double rcv = pymax(is_sep(p) ? r : r - s_rev, 0.0);
        st_keep<3>(c.f_snd + e, s.flow, pol);      // read back by the next step's smoothing (link.py:363, 400)
        st_keep<3>(c.f_rcv + e, rcv, pol);
        // node-major hand-over to the node pass; a slot that held 0 (the previous sending flow) and gets 0 again
        // needs no store (scattered 8-byte stores are the expensive kind)
        if (!PNS_SKIP_ZERO_HANDOVER || s.flow != 0.0 || snd_prev != 0.0) st_keep<1>(c.s.nm_s + slots.x, s.flow, pol);
        st_keep<1>(c.s.nm_r + slots.y, rcv, pol);
    }
}

#endif  // !PNS_HOST_EMULATION

// =================================================================================================
// Control environment (reference rl/builders.py, rl/pz_pednet_env.py)

// ActionApplier.clip_gater_action_value / clip_separator_action_value (rl/builders.py:281-352): the width an
// action asks for, rate-limited around the current width and clipped to the agent's bounds
__device__ __forceinline__ double clipped_action(const pns_env& env, int a, double v /* float(action[i]) */, double cur) {
    const double md = env.act_max_delta[a];
    if (fabs(v - cur) > md) {
        const double d = fmin(fmax(v - cur, -md), md);                   // np.clip
        v = cur + d;
    }
    return fmin(fmax(v, env.act_lo[a]), env.act_hi[a]);
}

// ActionApplier.clip_gater_action_value / clip_separator_action_value + setters
// (rl/builders.py:281-352; link.py:121-126, 462-478)
__global__ void __launch_bounds__(kBlock) k_env_actions(const __grid_constant__ EnvCtx x) {
    const Ctx& c = x.c;
    const int R = c.n.replicas;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)x.env.n_act * R) return;
    PNS_PDL_TRIGGER();
    const int a = (int)(gid / R), rep = (int)(gid % R);
    const int l = x.env.act_link[a];
    PNS_PDL_WAIT();                      // actions and gate table may come from the previous kernel
    const size_t e = (size_t)l * R + rep;
    const double v = clipped_action(x.env, a, (double)x.actions[(size_t)rep * x.env.n_act + a], c.s.gate[e]);
    c.s.gate[e] = v;
    if (x.env.act_sep[a]) {
        const size_t er = (size_t)(l ^ 1) * R + rep;
        c.s.gate[er] = x.env.act_total_width[a] - v;
        c.s.sep_np64[e] = 1;       // np.clip returns numpy float64: the lane area becomes a float64
        c.s.sep_np64[er] = 1;
    }
}

// Link.get_density(t) (link.py:190-197, 427): both directions' pedestrians over the shared area; a separator
// lane's own density.  CG: read around L1 (values another CTA of the same launch has just written).
template <bool CG>
__device__ __forceinline__ float shared_density(const Ctx& c, int l, int rep, int t) {
    const int R = c.n.replicas;
    const LinkP& p = c.n.classes[class_of(c, l, rep)];
    const size_t e = (size_t)l * R + rep;
    if (is_sep(p)) return CG ? __ldcg(H32(c, PNS_F32_DENSITY, t) + e) : H32(c, PNS_F32_DENSITY, t)[e];
    const float* num = H32(c, PNS_F32_NUM_PED, t);
    const size_t er = (size_t)(l ^ 1) * R + rep;
    const Area ar = link_area(c, p, e, 0.0);
    return div_by_area((CG ? __ldcg(num + e) : num[e]) + (CG ? __ldcg(num + er) : num[er]), ar);
}

// PedNetParallelEnv._compute_rewards (pz_pednet_env.py:548-581) of one replica at row t, float32 arithmetic as numpy
// evaluates it: the first agent's controlled links only (the reference returns inside its loop)
template <bool CG>
__device__ __forceinline__ float env_reward(const EnvCtx& x, int rep, int t) {
    const Ctx& c = x.c;
    const int R = c.n.replicas;
    const int n = x.env.n_reward_links;
    float total = 0.0f, rho[PNS_MAX_DEGREE];
    const float* tt = H32(c, PNS_F32_TRAVEL_TIME, t);
    for (int i = 0; i < n; ++i) {
        const int l = x.env.reward_link[i];
        const size_t e = (size_t)l * R + rep, er = (size_t)(l ^ 1) * R + rep;
        rho[i] = shared_density<CG>(c, l, rep, t);
        total = total - ((CG ? __ldcg(tt + e) : tt[e]) + (CG ? __ldcg(tt + er) : tt[er]));
        if (rho[i] > 4.0f) total = total - 10.0f * (rho[i] - c.n.classes[class_of(c, l, rep)].kc32);
    }
    if (n > 1) {
        float s = rho[0];
        for (int i = 1; i < n; ++i) s = s + rho[i];
        const float mean = s / (float)n;
        float d = fabsf(rho[0] - mean);
        for (int i = 1; i < n; ++i) d = d + fabsf(rho[i] - mean);
        total = total - 10.0f * (d / (float)n);
    }
    return total;
}

__global__ void __launch_bounds__(kBlock) k_env_observe(const __grid_constant__ EnvCtx x) {
    const Ctx& c = x.c;
    const int R = c.n.replicas;
    const int t = c.t;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_obs_threads = (size_t)x.env.n_obs * R;
    PNS_PDL_TRIGGER();
    PNS_PDL_WAIT();                      // reads the rows the link pass has just written
    if (gid < n_obs_threads) {
        const int k = (int)(gid / R), rep = (int)(gid % R);
        const int l = x.env.obs_link[k];
        const size_t e = (size_t)l * R + rep, er = (size_t)(l ^ 1) * R + rep;
        float v;
        switch (x.env.obs_src[k]) {
            case PNS_OBS_INFLOW: v = (float)H64(c, PNS_F64_INFLOW, t)[e]; break;
            case PNS_OBS_OUTFLOW: v = (float)H64(c, PNS_F64_OUTFLOW, t)[e]; break;
            case PNS_OBS_REV_INFLOW: v = (float)H64(c, PNS_F64_INFLOW, t)[er]; break;
            case PNS_OBS_REV_OUTFLOW: v = (float)H64(c, PNS_F64_OUTFLOW, t)[er]; break;
            case PNS_OBS_SHARED_DENSITY: v = shared_density<false>(c, l, rep, t); break;
            case PNS_OBS_SHARED_DENSITY_OVER_KJ: v = shared_density<false>(c, l, rep, t) / c.n.classes[class_of(c, l, rep)].kj32; break;
            case PNS_OBS_SPEED: v = H32(c, PNS_F32_SPEED, t)[e]; break;
            default: v = (float)c.s.gate[e]; break;
        }
        const float d = x.env.obs_div[k];
        if (d != 1.0f) v = v / d;
        x.obs[(size_t)rep * x.env.n_obs + k] = v;
        return;
    }
    // reward: one thread per replica (pz_pednet_env.py:548-581, float32 arithmetic as numpy does it)
    const size_t rid = gid - n_obs_threads;
    if (rid >= (size_t)R) return;
    const int rep = (int)rid;
    const float total = env_reward<false>(x, rep, t);
    x.reward[rep] = total;
    if (x.cum_reward) x.cum_reward[rep] = x.cum_reward[rep] + total;
}

#ifndef PNS_HOST_EMULATION
// =================================================================================================
// Batched replicas: one thread per (directed link, replica).  A CTA is one warp pair: its two warps are the two
// directions of one corridor for the same 32 replicas (every access of a warp is 32 consecutive elements of a
// history row) and trade pedestrians, density, speed noise and sending flow through shared memory behind the
// CTA barrier.  (One barrier per CTA: a dynamic named-barrier id makes ptxas reserve all 16 hardware barriers,
// which caps an SM at 4 resident CTAs.)  Same device functions as k_link_pair, half the state per thread
// (twice the resident warps), every load that does not depend on a computed lag issued up front, the arrival
// row fetched for the two likeliest lags and the diffusion taps requested as soon as the link is known to be
// occupied.  ENV: the control environment rides along -- the FLOWS phase applies the step's actions to the
// widths it reads (ActionApplier, rl/builders.py:264-352), the UPDATE phase emits the observation entries of
// its link and, as the last thread of a replica's reward links to finish, the reward (pz_pednet_env.py:548-581).
// (The host-emulation test build runs k_link_pair and the stand-alone environment kernels instead; the GPU
// parity tests compare the two paths.)
#ifndef PNS_REP_MIN_BLOCKS
#define PNS_REP_MIN_BLOCKS 12
#endif
constexpr int kRepBlock = 64;

template <int PHASE, int MODE, bool ONECLASS, bool ENV>
#ifndef PNS_REP_L2
#define PNS_REP_L2 0
#endif
#if PNS_REP_L2
#define RLD_KEEP(a) ldh((a), pol.keep)
#define RLD_ONCE(a) ldh((a), pol.once)
#define RST_KEEP(a, v) sth((a), (v), pol.keep)
#define RST_ONCE(a, v) sth((a), (v), pol.once)
#else
#define RLD_KEEP(a) (*(a))
#define RLD_ONCE(a) (*(a))
#define RST_KEEP(a, v) (*(a) = (v))
#define RST_ONCE(a, v) (*(a) = (v))
#endif
__global__ void __launch_bounds__(kRepBlock, PNS_REP_MIN_BLOCKS) k_link_rep(const __grid_constant__ EnvCtx x) {
    constexpr bool upd = (PHASE & PH_UPDATE) != 0, flw = (PHASE & PH_FLOWS) != 0;
    const Ctx& c = x.c;
    // L2 residency (PNS_REP_L2 = 1): a step moves about twice what L2 holds, so what the next launch reads back --
    // the hand-over to the node pass, the row the update leaves for the next flow launch -- would be stored
    // evict-last, and what is used once (lagged rows, the node pass's answer, series nobody reads inside a step)
    // evict-first.  Measured: no difference (8192 replicas 110.0 vs 108.8 us per step, 4096: 61.4 vs 62.8) -- the
    // kernels wait on dependent loads and arithmetic, not on DRAM bandwidth; left off.
    const L2Pol pol = l2_policies();
    (void)pol;
    __shared__ float sh_num[kRepBlock], sh_dens[kRepBlock];
    __shared__ double sh_noise[kRepBlock], sh_send[kRepBlock];
    const int R = c.n.replicas;
    const unsigned dir = threadIdx.x >> 5;
    const int rep_raw = (int)(blockIdx.x * 32u + (threadIdx.x & 31u));
    const unsigned n_pairs = (unsigned)c.n.n_links >> 1;
    if (flw && blockIdx.y >= n_pairs) {                     // route choice rides along: two rows per CTA
        // (one row per CTA with its groups split between the two warps was tried: 35.9 vs 34.3 us per step at 1024
        // replicas, 115 vs 108 us at 8192 -- with 32 replicas per warp the rows are not the longest threads)
        const unsigned row = 2u * (blockIdx.y - n_pairs) + dir;
        if (row < (unsigned)route_rows(c) && rep_raw < R) route_thread(c, c.route_t, (int)row, rep_raw);
        return;
    }
    const unsigned pair = blockIdx.y;
    const bool valid = rep_raw < R;
    const int rep = valid ? rep_raw : R - 1;          // idle lanes shadow the last replica and store nothing
    const int l = (int)(2u * pair + dir);
    const size_t e = (size_t)l * R + rep, er = (size_t)(l ^ 1) * R + rep;
    const unsigned mate = threadIdx.x ^ 32u;
    const int tau = c.t_flows - 1;
    PNS_PDL_TRIGGER();
    const LinkP& p = ONECLASS ? c.n.class0 : c.n.classes[class_of(c, l, rep)];
    const int2 slots = __ldg(reinterpret_cast<const int2*>(c.n.lk_slots) + l);   // {sending slot, receiving slot}
    const int fftau = p.fftau, swtau = p.swtau;
    int act_own = -1, act_mate = -1, obs0 = 0, obs1 = 0, rewarded = 0;
    if (ENV && flw && x.actions) {
        act_own = __ldg(x.env.lk_act + l);
        act_mate = __ldg(x.env.lk_act + (l ^ 1));
    }
    if (ENV && upd && x.obs) {
        obs0 = __ldg(x.env.lk_obs_ptr + l); obs1 = __ldg(x.env.lk_obs_ptr + l + 1);
        rewarded = x.env.n_reward_links > 0 ? __ldg(x.env.lk_reward + l) : 0;
    }
    PNS_PDL_WAIT();             // everything above is static; below reads what the previous kernel wrote
    double gate = c.s.gate[e];
    double gate_rev = flw ? c.s.gate[er] : 0.0;
    const bool windowed = c.u_tt_old != nullptr;
    if (x.pf_pairs && pair + x.pf_pairs < n_pairs) {
        // the loads below, for the CTA one resident wave later: its batch then finds the rows in L2.  One lane per
        // 128-byte line asks (16 doubles / 32 floats).
        const size_t a = e + x.pf_off;
        const unsigned ln = threadIdx.x & 31u;
        if ((ln & 15u) == 0) {
            prefetch_l2(c.s.gate + a);
            if (upd) {
                prefetch_l2(c.n_cinp + a); prefetch_l2(c.n_coutp + a);
                prefetch_l2(c.n_outflow + a); prefetch_l2(c.n_inflow + a);
            }
            if (flw) {
                if (!upd) { prefetch_l2(c.f_cin + a); prefetch_l2(c.f_cou + a); }
                prefetch_l2(c.f_sndp + a); prefetch_l2(c.f_rcvp + a);
                if (ONECLASS) {
                    if (c.c0_coulag) prefetch_l2(c.c0_coulag + a);
                    if (c.c0_pre0) prefetch_l2(c.c0_pre0 + a);
                    if (c.c0_pre1) prefetch_l2(c.c0_pre1 + a);
                    // (the lagged inflow rows of get_outflow were tried too: 112.6 vs 108.3 us per step at 8192
                    // replicas -- most links do not need them and the extra 40 B per link-step cost more)
                }
            }
        }
        if (ln == 0) {
            if (upd) {
                prefetch_l2(c.u_num_prev + a); prefetch_l2(c.s.runsum + a);
                if (windowed) prefetch_l2(c.u_tt_old + a);
            }
            if (flw && !upd) { prefetch_l2(c.f_num + a); prefetch_l2(c.f_dens + a); prefetch_l2(c.f_avg + a); }
        }
    }
    // ---- batch of independent loads --------------------------------------------------------
    double din = 0, dout = 0, cin_prev = 0, cou_prev = 0;
    float np_ = 0, rs = 0, tt_old = 0;
    if (upd) {
        cin_prev = RLD_ONCE(c.n_cinp + e); cou_prev = RLD_ONCE(c.n_coutp + e);
        np_ = RLD_ONCE(c.u_num_prev + e); rs = c.s.runsum[e];
        if (windowed) tt_old = RLD_ONCE(c.u_tt_old + e);
        dout = RLD_ONCE(c.n_outflow + e); din = RLD_ONCE(c.n_inflow + e);
    }
    double cin_tau = 0, cou_tau = 0, snd_prev = 0, rcv_prev = 0, cou_lag = 0;
    LinkNow me;
    me.num = 0; me.dens = 0; me.avg_tt = 0;
    float num_rev = 0.0f;
    int pre_i0 = -1, pre_i1 = -1;
    double pre_v0 = 0.0, pre_v1 = 0.0;
    if (flw) {
        if (!upd) {
            cin_tau = RLD_KEEP(c.f_cin + e); cou_tau = RLD_KEEP(c.f_cou + e);       // the update reads them again
            me.num = RLD_KEEP(c.f_num + e); me.dens = RLD_ONCE(c.f_dens + e); me.avg_tt = RLD_ONCE(c.f_avg + e);
            num_rev = RLD_KEEP(c.f_num + er);
        }
        snd_prev = RLD_ONCE(c.f_sndp + e); rcv_prev = RLD_ONCE(c.f_rcvp + e);
        if (ONECLASS) {
            if (c.c0_coulag) cou_lag = RLD_ONCE(c.c0_coulag + e);
        } else {
            const int lag_i = tau + 1 - swtau;
            if (lag_i >= 0) cou_lag = H64(c, PNS_F64_CUM_OUTFLOW, lag_i)[e];
        }
        // the arrival row cumulative_inflow[tau+1-lag]: in free flow the lag is the free-flow lag or one less
        // (speed noise), so both rows are fetched with the batch; a congested link falls back to a dependent load
        const double *pre_row0 = nullptr, *pre_row1 = nullptr;
        if (ONECLASS) {
            pre_i0 = c.c0_pre_i0; pre_i1 = c.c0_pre_i1; pre_row0 = c.c0_pre0; pre_row1 = c.c0_pre1;
        } else if (tau >= fftau) {
            pre_i0 = max(0, tau + 1 - fftau);
            pre_row0 = H64(c, PNS_F64_CUM_INFLOW, pre_i0);
            if (fftau > 1) {
                pre_i1 = max(0, tau + 2 - fftau);
                pre_row1 = H64(c, PNS_F64_CUM_INFLOW, pre_i1);
            }
        }
        if (pre_row0) pre_v0 = RLD_ONCE(pre_row0 + e);
        if (pre_row1) pre_v1 = RLD_ONCE(pre_row1 + e);
    }
    // ---- actions of this environment step (rl/builders.py:264-352; setters link.py:121-126, 462-478) --------
    // Each thread derives the new width of its own link and of its mate's from the old widths, so neither waits
    // for the other; its own store happens behind the sending-flow barrier, after the mate has read the old value.
    bool gate_changed = false, np64_now = false;
    if (ENV && flw && (act_own >= 0 || act_mate >= 0)) {
        const double old_own = gate, old_rev = gate_rev;
        const float* act = x.actions + (size_t)rep * x.env.n_act;
        double v_own = 0.0, v_mate = 0.0;
        if (act_own >= 0) v_own = clipped_action(x.env, act_own, (double)act[act_own], old_own);
        if (act_mate >= 0) v_mate = clipped_action(x.env, act_mate, (double)act[act_mate], old_rev);
        const bool sep_own = act_own >= 0 && __ldg(x.env.act_sep + act_own) != 0;
        const bool sep_mate = act_mate >= 0 && __ldg(x.env.act_sep + act_mate) != 0;
        if (act_own >= 0) { gate = v_own; gate_changed = true; np64_now = sep_own; }
        else if (sep_mate) { gate = x.env.act_total_width[act_mate] - v_mate; gate_changed = true; np64_now = true; }
        if (act_mate >= 0) gate_rev = v_mate;
        else if (sep_own) gate_rev = x.env.act_total_width[act_own] - v_own;
    }
    const Area ar = link_area(c, p, e, gate, np64_now);
    const uint32_t k0 = (uint32_t)c.io.seed, k1 = (uint32_t)(c.io.seed >> 32);
    const uint32_t rkey = (uint32_t)rep + c.io.replica_base;

    if (upd) {
        cin_tau = cin_prev + din;                                           // link.py:19-25
        cou_tau = cou_prev + dout;
        if (valid) { RST_KEEP(c.n_cin + e, cin_tau); RST_KEEP(c.n_cout + e, cou_tau); }
        me.num = (float)((double)np_ + (din - dout));                      // link.py:134-135
        me.dens = div_by_area(me.num, ar);                                  // link.py:136
        const bool noisy = p.sigma > 0.0;
        double z = 0.0;
        sh_num[threadIdx.x] = me.num; sh_dens[threadIdx.x] = me.dens;
        if (noisy) {
            if (MODE == PNS_RNG_TABLE) z = c.draw_n[e];
            else if (dir == 0) {
                // one Philox block and one Box-Muller pair per corridor: the forward warp draws, keeps the cosine
                // branch and hands the sine branch to its mate
                pns::DrawKey key;
                key.t = (uint32_t)c.t; key.link = (uint32_t)l; key.replica = rkey; key.k0 = k0; key.k1 = k1;
                double g0, g1;
                pns::normal_pair_philox(key, 4u, &g0, &g1);
                z = g0;
                sh_noise[threadIdx.x] = g1;
            }
        }
        __syncthreads();
        num_rev = sh_num[mate];
        const float dens_rev = sh_dens[mate];
        if (noisy) {
            if (MODE != PNS_RNG_TABLE && dir == 1) z = sh_noise[mate];
            if (MODE != PNS_RNG_TABLE) z = p.sigma * z;
        }
        float tt;
        const float v = speed_and_travel_time(p, me.dens, is_sep(p) ? 0.0f : dens_rev, noisy, z, &tt);
        float sum = rs + tt;                                                // link.py:183-186
        if (windowed) {
            sum = sum - tt_old;
            me.avg_tt = sum / (float)c.n.window;
        } else {
            me.avg_tt = p.tt0;
        }
        if (valid) {
            RST_KEEP(c.u_num + e, me.num); RST_KEEP(c.u_dens + e, me.dens);
            RST_ONCE(c.u_speed + e, v); RST_ONCE(c.u_tt + e, tt);
            RST_ONCE(c.u_flow + e, v * me.dens);                            // functions.py:97-101
            if (windowed) RST_KEEP(c.u_avg + e, me.avg_tt);
            c.s.runsum[e] = sum;
            RST_ONCE(c.u_bgw + e, gate);                                    // link.py:188, 451-452
            if (is_sep(p)) RST_ONCE(c.u_sepw + e, gate);
        }
        if (ENV && x.obs && valid) {
            // ObservationBuilder (rl/builders.py:68-177): the entries this directed link contributes
            for (int k = obs0; k < obs1; ++k) {
                float f;
                switch (__ldg(x.env.lk_obs_src + k)) {
                    case PNS_OBS_INFLOW: f = (float)din; break;
                    case PNS_OBS_OUTFLOW: f = (float)dout; break;
                    case PNS_OBS_SHARED_DENSITY: f = is_sep(p) ? me.dens : div_by_area(me.num + num_rev, ar); break;
                    case PNS_OBS_SHARED_DENSITY_OVER_KJ:
                        f = (is_sep(p) ? me.dens : div_by_area(me.num + num_rev, ar)) / p.kj32; break;
                    case PNS_OBS_SPEED: f = v; break;
                    default: f = (float)gate; break;
                }
                const float d = __ldg(x.env.lk_obs_div + k);
                if (d != 1.0f) f = f / d;
                x.obs[(size_t)rep * x.env.n_obs + __ldg(x.env.lk_obs_col + k)] = f;
            }
            // reward: needs the rows of all reward links of the replica; whichever of their threads finishes last
            // evaluates it, reading the others' values around L1
            if (rewarded) {
                __threadfence();
                const int seen = atomicAdd(x.env.reward_count + rep, 1);
                if (seen == 2 * x.env.n_reward_links - 1) {
                    x.env.reward_count[rep] = 0;
                    __threadfence();
                    const float total = env_reward<true>(x, rep, c.t);
                    x.reward[rep] = total;
                    if (x.cum_reward) x.cum_reward[rep] = x.cum_reward[rep] + total;
                }
            }
        }
    }
    if (!flw) return;
    // occupied links will probably smooth their outflow over four lagged inflow rows (get_outflow,
    // link.py:199-214) of a lag known only deep inside the sending-flow computation: request the five candidate
    // rows now
    if (me.num > 0.0f && tau >= fftau) {
        const double* row = H64(c, PNS_F64_INFLOW, tau - fftau + 1) + e;
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (tau - fftau + 1 - k >= 0) prefetch_l1(row - (size_t)k * c.row64);
    }
    pns::DrawKey key;
    key.t = (uint32_t)c.t_flows; key.link = (uint32_t)l; key.replica = rkey; key.k0 = k0; key.k1 = k1;
    const double front = is_sep(p) ? gate : gate_rev;                       // link.py:110-126, 462-478
    SendOut s;
    // nobody on the link and nobody in transit: boundary flow 0, only the smoothing with the previous sending
    // flow remains (link.py:363-364); exact, see k_link_lane
    if (me.num == 0.0f && cin_tau == cou_tau && front >= 0.0 && tau >= fftau) {
        s.kind = 0; s.n1 = 0; s.rf = 0.0f; s.sval = 0.0;
        s.flow = pymin(floor(0.8 * 0.0 + 0.2 * snd_prev), 0.0);
        if (s.flow < 0.0) atomicOr(c.s.err + rep, PNS_ERR_NEG_SENDING);
    } else {
        s = sending_flow<MODE>(c, p, e, tau, me, num_rev, ar, front, cou_tau, snd_prev, rep, key, pre_i0, pre_v0,
                               pre_i1, pre_v1);
    }
    int n3 = -1;
    const double r = receiving_flow<MODE>(c, p, e, tau, num_rev, ar, gate, cin_tau, cou_lag, rcv_prev, key, &n3);
    sh_send[threadIdx.x] = s.flow;
    __syncthreads();
    const double s_rev = sh_send[mate];
    if (valid) {
        // cal_receiving_flow_with_reverse (link.py:407-416; separators ignore the reverse flow, :509-512)
        const double rcv = pymax(is_sep(p) ? r : r - s_rev, 0.0);
        RST_ONCE(c.f_snd + e, s.flow);
        RST_ONCE(c.f_rcv + e, rcv);
        RST_KEEP(c.s.nm_s + (size_t)slots.x * R + rep, s.flow);
        RST_KEEP(c.s.nm_r + (size_t)slots.y * R + rep, rcv);
        if (ENV && gate_changed) {
            c.s.gate[e] = gate;
            if (np64_now) c.s.sep_np64[e] = 1;   // np.clip returns numpy float64: the lane area becomes a float64
        }
    }
}
#endif  // !PNS_HOST_EMULATION

// =================================================================================================
// Initial state (link.py:12-17, 56, 82-97, 425)
__global__ void k_state_init(const __grid_constant__ Ctx c) {
    const int R = c.n.replicas;
    const size_t n64 = c.row64, n32 = c.row32;
    const int T = c.n.sim_steps + 1;
    const size_t total = (size_t)T * n64;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / n64);
        const size_t e = i % n64;
        const bool physical = e < n32;
        const int l = physical ? (int)(e / R) : 0;
        const LinkP* p = physical ? c.n.classes + class_of(c, l, (int)(e % R)) : nullptr;
        for (int f = 0; f < c.s.n_f64; ++f) {
            double v = 0.0;
            if (f == PNS_F64_SENDING || f == PNS_F64_RECEIVING) v = -1.0;
            else if (f == PNS_F64_BACK_GATE && physical) v = c.n.lk_width[l];                  // link.py:56
            else if (f == PNS_F64_SEP_WIDTH && physical && is_sep(*p)) v = c.n.lk_width[l] / 2;  // link.py:425
            H64(c, f, t)[e] = v;
        }
        if (physical) {
            const float tt0 = p->tt0;
            H32(c, PNS_F32_NUM_PED, t)[e] = 0.0f;
            H32(c, PNS_F32_DENSITY, t)[e] = 0.0f;
            H32(c, PNS_F32_SPEED, t)[e] = 0.0f;
            H32(c, PNS_F32_LINK_FLOW, t)[e] = 0.0f;
            H32(c, PNS_F32_TRAVEL_TIME, t)[e] = t == 0 ? tt0 : 0.0f;
            H32(c, PNS_F32_AVG_TRAVEL_TIME, t)[e] = t < c.n.window ? tt0 : 0.0f;
            if (t == 0) c.s.runsum[e] = tt0;
        }
    }
}

// Network-wide pedestrian count of row t (sum over links and replicas), accumulated in double.
__global__ void __launch_bounds__(256) k_metric_pedestrians(const __grid_constant__ Ctx c, double* out) {
    const float* row = c.u_num;
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < c.row32; i += (size_t)gridDim.x * blockDim.x)
        acc += (double)row[i];
#ifndef PNS_HOST_EMULATION
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(out, acc);
#else
    *out += acc;
#endif
}

// =================================================================================================
// Origin demand of the batched environment, drawn on the device (reference od_manager.py:100-155:
// Poisson around base + peak * (bump(S/4) + bump(3S/4)); 'constant'; 'sudden_demand' adds a burst of
// random length, start and height).  One thread per (step, demand row, replica); every value is a pure
// function of (seed, global replica; step, row), so it does not depend on how replicas are sharded.
struct DemandCtx {
    int S, rows, R;
    uint32_t replica_base;
    uint64_t seed;
    const double *bump1, *bump2;     // [S] the two gaussian bumps, evaluated on the host like the reference does
    const double *base, *peak;       // [rows*R]
    const int32_t* pattern;          // [rows*R] 0 gaussian_peaks, 1 constant, 2 sudden_demand, -1 no demand
    double* out;                     // [S+1][rows*R]
};
__global__ void __launch_bounds__(kBlock) k_demand_draw(const DemandCtx d) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_row = (size_t)d.rows * d.R;
    if (gid >= (size_t)(d.S + 1) * per_row) return;
    const int t = (int)(gid / per_row);
    const size_t col = gid % per_row;
    const int row = (int)(col / d.R), rep = (int)(col % d.R);
    const int pat = d.pattern[col];
    double v = 0.0;
    if (pat == 1) {
        v = d.base[col];                                              // od_manager.py:106-109, all S+1 entries
    } else if (pat >= 0 && t < d.S) {
        pns::DrawKey key;
        key.t = (uint32_t)t; key.link = (uint32_t)row; key.replica = d.replica_base + (uint32_t)rep;
        key.k0 = (uint32_t)d.seed; key.k1 = (uint32_t)(d.seed >> 32);
        const double lam = (d.base[col] + d.peak[col] * d.bump1[t]) + d.peak[col] * d.bump2[t];
        const pns::Philox4 w = pns::philox4x32_10(key.t, key.link, 8u, key.replica, key.k0, key.k1);
        v = (double)pns::poisson_inversion(lam, pns::u53(w.v[0], w.v[1]));
        if (pat == 2) {                                               // od_manager.py:111-123
            const pns::Philox4 b = pns::philox4x32_10(0u, key.link, 9u, key.replica, key.k0, key.k1);
            const int period = 10 + (int)(b.v[0] % 10u);
            const int span = d.S - period > 1 ? d.S - period : 1;
            const int start = (int)(b.v[1] % (uint32_t)span);
            if (t >= start && t < start + period) v += (double)(20 + (int)(b.v[2] % 30u));
        }
    }
    d.out[gid] = v;
}

// =================================================================================================
// Domain randomisation on the device (reference NetworkEnvGenerator.generate_random_link_params /
// generate_random_od_flows / generate_random_demand_params, env_loader.py:183-258, 363-424): every replica and
// episode gets its own scenario without a host loop over replicas.  Same distributions as the reference's
// generators, drawn from Philox blocks keyed (seed; replica) instead of numpy's stream (restated in
// oracle/philox.py: device_scenario):
//   * n_change = int(0.2 * corridors) corridors chosen uniformly without replacement (partial Fisher-Yates);
//     each, with probability 1/2, gets a capacity factor f ~ U[0.6, 1.2): k_c' = max(0.5, k_c f),
//     k_j' = max(2 k_c', k_j f), and, with probability 1/2, a free-flow speed factor ~ U[0.6, 0.9);
//   * every OD pair a weight ~ U[1, 10), constant over the episode;
//   * every origin a demand pattern (gaussian_peaks / constant / sudden_demand, equally likely), a base rate
//     ~ U[2, 10) and a peak rate ~ U[10, 30), raised to base + 5 when it is closer than that.
// One thread per replica; the parameter classes of its perturbed corridors are built here with the arithmetic
// of plan.class_record.
struct ScenarioCtx {
    int n_links, R, n_change, n_base, n_od, S, rows;
    uint32_t replica_base;
    uint64_t seed;
    double unit_time;
    pns_link_class* classes;       // [n_base + R*n_change]; the first n_base are the unperturbed classes
    int32_t* lk_class;             // [n_links*R]
    const int32_t* base_class;     // [n_links]
    const double* lk_width;        // [n_links]
    double* od_w;                  // [S+1][n_od*R]
    double *dem_base, *dem_peak;   // [rows*R]
    int32_t* dem_pattern;          // [rows*R]; rows that are not origins keep -1
    const int32_t* row_is_origin;  // [rows]
};

__device__ __forceinline__ pns_link_class make_link_class(const pns_link_class& b, double width, double vf, double kc,
                                                          double kj, double unit_time) {
    pns_link_class r = b;                                       // length, gamma, act, bi, sigma, flags unchanged
    const double area = b.length * width;                       // link.py:131
    const double slow = b.length / 0.05;                        // link.py:63
    const double free = b.length / vf;
    const float tt0 = (float)(free < slow ? free : slow);       // link.py:83
    const double shock = (vf * kc) / (kj - kc);                 // link.py:58,61
    r.area = area; r.space = kj * area;
    r.kc = kc; r.vf = vf; r.kj = kj;
    r.kc32 = (float)kc; r.kj32 = (float)kj; r.kj_minus_kc32 = (float)(kj - kc);
    r.area32 = (float)area;
    r.yp_coef32 = (float)((kc * vf) / (kj - kc));
    r.neg_vf32 = (float)(-vf);
    r.vf32 = (float)vf;
    r.sm_gamma32 = (float)(vf * kc);
    r.inv_kj32 = (float)(1.0 / kj);
    r.max_tt32 = (float)slow;
    r.tt0 = tt0;
    r.fftau = __float2int_rn(tt0 / (float)unit_time);           // link.py:86
    r.swtau = (int)rint(b.length / (shock * unit_time));        // link.py:380
    return r;
}

__global__ void __launch_bounds__(kBlock) k_scenario_draw(const ScenarioCtx x) {
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= x.R) return;
    const uint32_t rkey = x.replica_base + (uint32_t)rep, k0 = (uint32_t)x.seed, k1 = (uint32_t)(x.seed >> 32);
    const int R = x.R, C = x.n_links / 2;
    for (int l = 0; l < x.n_links; ++l) x.lk_class[(size_t)l * R + rep] = x.base_class[l];
    // corridors to perturb: partial Fisher-Yates over 0..C-1, positions swapped so far kept in a short list
    int moved_pos[64], moved_val[64];
    int n_moved = 0;
    const int n_change = x.n_change < 32 ? x.n_change : 32;
    for (int k = 0; k < n_change; ++k) {
        const pns::Philox4 w = pns::philox4x32_10((uint32_t)k, 0u, 10u, rkey, k0, k1);
        int j = k + (int)(pns::u53(w.v[0], w.v[1]) * (double)(C - k));
        if (j >= C) j = C - 1;
        int vj = j, vk = k, ij = -1, ik = -1;                   // current contents of positions j and k
        for (int q = 0; q < n_moved; ++q) {
            if (moved_pos[q] == j) { vj = moved_val[q]; ij = q; }
            if (moved_pos[q] == k) { vk = moved_val[q]; ik = q; }
        }
        if (ij >= 0) moved_val[ij] = vk; else { moved_pos[n_moved] = j; moved_val[n_moved] = vk; ++n_moved; }
        if (ik >= 0) moved_val[ik] = vj; else if (j != k) { moved_pos[n_moved] = k; moved_val[n_moved] = vj; ++n_moved; }
        const int corridor = vj;                                 // chosen: links 2*corridor, 2*corridor + 1
        const pns::Philox4 a = pns::philox4x32_10((uint32_t)k, 0u, 11u, rkey, k0, k1);
        const pns::Philox4 b = pns::philox4x32_10((uint32_t)k, 0u, 12u, rkey, k0, k1);
        const pns_link_class base = x.classes[x.base_class[2 * corridor]];
        double kc = base.kc, kj = base.kj, vf = base.vf;
        if (a.v[0] & 1u) {
            const double f = 0.6 + 0.6 * pns::u53(a.v[2], a.v[3]);
            kc = fmax(0.5, base.kc * f);
            kj = fmax(kc * 2.0, base.kj * f);
        }
        if (a.v[1] & 1u) vf = base.vf * (0.6 + 0.3 * pns::u53(b.v[0], b.v[1]));
        const int slot = x.n_base + rep * x.n_change + k;
        x.classes[slot] = make_link_class(base, x.lk_width[2 * corridor], vf, kc, kj, x.unit_time);
        x.lk_class[(size_t)(2 * corridor) * R + rep] = slot;
        x.lk_class[(size_t)(2 * corridor + 1) * R + rep] = slot;
    }
    for (int od = 0; od < x.n_od; ++od) {
        const pns::Philox4 w = pns::philox4x32_10((uint32_t)od, 0u, 13u, rkey, k0, k1);
        const double weight = 1.0 + 9.0 * pns::u53(w.v[0], w.v[1]);
        for (int t = 0; t <= x.S; ++t) x.od_w[((size_t)t * x.n_od + od) * R + rep] = weight;
    }
    for (int row = 0; row < x.rows; ++row) {
        const size_t col = (size_t)row * R + rep;
        if (!x.row_is_origin[row]) { x.dem_pattern[col] = -1; x.dem_base[col] = 0.0; x.dem_peak[col] = 0.0; continue; }
        const pns::Philox4 w = pns::philox4x32_10((uint32_t)row, 0u, 14u, rkey, k0, k1);
        const pns::Philox4 v = pns::philox4x32_10((uint32_t)row, 0u, 15u, rkey, k0, k1);
        const double base = 2.0 + 8.0 * pns::u53(w.v[1], w.v[2]);
        double peak = 10.0 + 20.0 * pns::u53(v.v[0], v.v[1]);
        if (peak < base + 5.0) peak = base + 5.0;
        x.dem_pattern[col] = (int32_t)(w.v[0] % 3u);            // 0 gaussian_peaks, 1 constant, 2 sudden_demand
        x.dem_base[col] = base;
        x.dem_peak[col] = peak;
    }
}

// =================================================================================================
// Episode KPIs per replica (reference rl/rl_utils.py:770-1512, which reads them from the JSON that
// OutputHandler saves): reductions over the whole history, rows 0..t_last.
// Pass 1: one thread per (link, replica) walks its column through time (coalesced over replicas).
__global__ void __launch_bounds__(kBlock) k_kpi_links(const __grid_constant__ Ctx c, int t_last, double* part) {
    const int R = c.n.replicas;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)c.n.n_links * R) return;
    const int l = (int)(gid / R), rep = (int)(gid % R);
    const LinkP& p = c.n.classes[class_of(c, l, rep)];
    const double dt = c.n.unit_time;
    const double t_free = p.length / p.vf;                       // rl_utils.py:1023
    const double area_time = (p.length * c.n.lk_width[l]) * dt;   // rl_utils.py:1469-1477
    double spent = 0.0, ptime = 0.0, delay = 0.0, cong = 0.0, n_dens = 0.0, n_cong = 0.0, sum_tt = 0.0, n_tt = 0.0;
    for (int t = 0; t <= t_last; ++t) {
        const double num = (double)H32(c, PNS_F32_NUM_PED, t)[gid];
        const double tt = (double)H32(c, PNS_F32_TRAVEL_TIME, t)[gid];
        const double dens = (double)H32(c, PNS_F32_DENSITY, t)[gid];
        if (num >= 0.0) spent = spent + num * dt;                 // rl_utils.py:1131-1133
        if (tt > 0.0) {                                           // rl_utils.py:1041-1050
            const double frac = fmax(0.0, 1.0 - t_free / tt);
            delay = delay + (num * frac) * dt;
            ptime = ptime + num * dt;
        }
        if (dens >= 0.0) {                                        // rl_utils.py:1472-1484
            n_dens += 1.0;
            if (dens > p.kc) { n_cong += 1.0; cong = cong + (dens - p.kc) * area_time; }
        }
        if (tt >= 0.0) { sum_tt = sum_tt + tt; n_tt += 1.0; }     // rl_utils.py:935-941
    }
    double* o = part + gid * 8;
    o[0] = spent; o[1] = ptime; o[2] = delay; o[3] = cong; o[4] = n_dens; o[5] = n_cong; o[6] = sum_tt; o[7] = n_tt;
}

// Pass 2: one thread per replica adds the links up in network.links order.
// out[rep][PNS_KPI_COUNT]; lk_role bit0: starts at an origin, bit1: ends at a destination, bit2: on an OD path.
__global__ void __launch_bounds__(kBlock) k_kpi_reduce(const __grid_constant__ Ctx c, int t_last, const double* part,
                                                       const int32_t* lk_role, int any_od_path, double* out) {
    const int R = c.n.replicas;
    const int rep = blockIdx.x * blockDim.x + threadIdx.x;
    if (rep >= R) return;
    double k[PNS_KPI_COUNT];
    for (int i = 0; i < PNS_KPI_COUNT; ++i) k[i] = 0.0;
    const double* cin = H64(c, PNS_F64_CUM_INFLOW, t_last);
    const double* cou = H64(c, PNS_F64_CUM_OUTFLOW, t_last);
    double tt_links = 0.0, n_links_tt = 0.0;
    for (int l = 0; l < c.n.n_links; ++l) {
        const size_t e = (size_t)l * R + rep;
        const double* p = part + e * 8;
        const int role = lk_role[l];
        if (role & 1) k[PNS_KPI_TOTAL_INFLOW] += cin[e];
        if (role & 2) k[PNS_KPI_TOTAL_OUTFLOW] += cou[e];
        k[PNS_KPI_PERSON_TIME] += p[0];
        k[PNS_KPI_PERSON_TIME_MOVING] += p[1];
        k[PNS_KPI_TOTAL_DELAY] += p[2];
        k[PNS_KPI_CONGESTION_TIME] += p[3];
        k[PNS_KPI_AREA_TIME] += p[4] * ((c.n.classes[class_of(c, l, rep)].length * c.n.lk_width[l]) * c.n.unit_time);
        k[PNS_KPI_STEPS] += p[4];
        k[PNS_KPI_CONGESTED_STEPS] += p[5];
        if ((!any_od_path || (role & 4)) && p[7] > 0.0) { tt_links += p[6] / p[7]; n_links_tt += 1.0; }
    }
    k[PNS_KPI_AVG_TRAVEL_TIME] = n_links_tt > 0.0 ? tt_links / n_links_tt : 0.0;
    const int rows = c.n.n_demand_rows;
    for (int t = 0; t < c.n.sim_steps; ++t)
        for (int r0 = 0; r0 < rows; ++r0) k[PNS_KPI_TOTAL_DEMAND] += c.io.demand[((size_t)t * rows + r0) * R + rep];
    for (int i = 0; i < PNS_KPI_COUNT; ++i) out[(size_t)rep * PNS_KPI_COUNT + i] = k[i];
}

// sampler tables of pns_rng.cuh, once per device and process
__global__ void k_init_sampler_tables() {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= (PNS_MODE_TABLE_N > PNS_INV_TABLE ? PNS_MODE_TABLE_N : PNS_INV_TABLE)) pns::init_sampler_table_entry(i);
}
void ensure_sampler_tables(cudaStream_t s) {
#ifndef PNS_HOST_EMULATION
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || done[dev]) return;
    done[dev] = true;
#else
    static bool done = false;
    if (done) return;
    done = true;
#endif
    const int n = (PNS_MODE_TABLE_N > PNS_INV_TABLE ? PNS_MODE_TABLE_N : PNS_INV_TABLE) + 1;
    PNS_LAUNCH(k_init_sampler_tables, (n + 127) / 128, 128, s);
}

__global__ void k_rng_selftest(int kind, int n, const int32_t* n_trials, const double* p, uint64_t seed, int t,
                               int site, int32_t* out_i, double* out_d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pns::DrawKey key;
    key.t = (uint32_t)t; key.link = (uint32_t)i; key.replica = 0;
    key.k0 = (uint32_t)seed; key.k1 = (uint32_t)(seed >> 32);
    if (kind == 0) out_i[i] = pns::binomial_philox(key, (uint32_t)site, n_trials[i], p[i]);
    else if (kind == 3) {          // the shared block of a link's draws: site 1 -> R1 (any p), site 3 -> R3 (p = 0.9)
        out_i[i] = site == 3 ? pns::binomial_blockers(key, n_trials[i]) : pns::binomial_release(key, n_trials[i], p[i]);
    }
    else if (kind == 1) {
        float g[4];
        pns::normal_quad_philox(key, (uint32_t)site, g);
        for (int k = 0; k < 4; ++k) out_d[4 * i + k] = (double)g[k];
    } else out_d[i] = (double)pns::det_pow08((float)p[i]);
}

// ---- host side ----------------------------------------------------------------------------------
// row_update / row_flows: draw-table rows (relative to the first step of the call) of the UPDATE
// step and of the FLOWS step
Ctx make_ctx(const pns_net* net, const pns_state* st, const pns_step_io* io, int phase, int t, int t_flows,
             int mode, int row_update, int row_flows) {
    Ctx c;
    c.n = *net;
    c.s = *st;
    if (io) c.io = *io; else memset(&c.io, 0, sizeof c.io);
    c.t = t;
    c.t_flows = t_flows;
    c.phase = phase;
    c.mode = mode;
    c.row64 = (size_t)net->n_cols64 * net->replicas;
    c.row32 = (size_t)net->n_links * net->replicas;
    c.fld64 = c.row64 * (size_t)(net->sim_steps + 1);
    c.fld32 = c.row32 * (size_t)(net->sim_steps + 1);
    {   // rows known at launch time
        const int S1 = net->sim_steps + 1;
        auto h64 = [&](int f, int row) { return st->hist64 + (size_t)f * c.fld64 + (size_t)row * c.row64; };
        auto h32 = [&](int f, int row) { return st->hist32 + (size_t)f * c.fld32 + (size_t)row * c.row32; };
        const int tu = t >= 1 && t < S1 ? t : 1;                 // UPDATE / node row (clamped when unused)
        c.u_num_prev = h32(PNS_F32_NUM_PED, tu - 1);
        c.u_tt_old = tu >= net->window ? h32(PNS_F32_TRAVEL_TIME, tu - net->window) : nullptr;
        c.u_num = h32(PNS_F32_NUM_PED, tu); c.u_dens = h32(PNS_F32_DENSITY, tu); c.u_speed = h32(PNS_F32_SPEED, tu);
        c.u_tt = h32(PNS_F32_TRAVEL_TIME, tu); c.u_flow = h32(PNS_F32_LINK_FLOW, tu);
        c.u_avg = h32(PNS_F32_AVG_TRAVEL_TIME, tu);
        c.u_bgw = h64(PNS_F64_BACK_GATE, tu);
        c.u_sepw = st->n_f64 > PNS_F64_SEP_WIDTH ? h64(PNS_F64_SEP_WIDTH, tu) : nullptr;
        const int tf_ = t_flows >= 1 && t_flows <= S1 ? t_flows : 1;
        const int tau = tf_ - 1;
        const int prev = tau - 1 < 0 ? tau - 1 + S1 : tau - 1;   // numpy wrap of index -1 (link.py:364, 400)
        c.f_num = h32(PNS_F32_NUM_PED, tau); c.f_dens = h32(PNS_F32_DENSITY, tau);
        c.f_avg = h32(PNS_F32_AVG_TRAVEL_TIME, tau);
        c.f_cin = h64(PNS_F64_CUM_INFLOW, tau); c.f_cou = h64(PNS_F64_CUM_OUTFLOW, tau);
        c.f_sndp = h64(PNS_F64_SENDING, prev); c.f_rcvp = h64(PNS_F64_RECEIVING, prev);
        c.f_snd = h64(PNS_F64_SENDING, tau); c.f_rcv = h64(PNS_F64_RECEIVING, tau);
        c.n_coutp = h64(PNS_F64_CUM_OUTFLOW, tu - 1); c.n_cinp = h64(PNS_F64_CUM_INFLOW, tu - 1);
        c.n_outflow = h64(PNS_F64_OUTFLOW, tu); c.n_inflow = h64(PNS_F64_INFLOW, tu);
        c.n_cout = h64(PNS_F64_CUM_OUTFLOW, tu); c.n_cin = h64(PNS_F64_CUM_INFLOW, tu);
        c.n_demand = (io && io->demand) ? io->demand + (size_t)(tu - 1) * net->n_demand_rows * net->replicas : nullptr;
        // lagged rows of a single-class network (same expressions as k_link_lane evaluates per link otherwise)
        const int fftau = net->class0.fftau, swtau = net->class0.swtau;
        c.c0_coulag = tau + 1 - swtau >= 0 ? h64(PNS_F64_CUM_OUTFLOW, tau + 1 - swtau) : nullptr;
        c.c0_pre_i0 = c.c0_pre_i1 = -1;
        c.c0_pre0 = c.c0_pre1 = nullptr;
        if (tau >= fftau) {
            c.c0_pre_i0 = tau + 1 - fftau > 0 ? tau + 1 - fftau : 0;
            c.c0_pre0 = h64(PNS_F64_CUM_INFLOW, c.c0_pre_i0);
            if (fftau > 1) {
                c.c0_pre_i1 = tau + 2 - fftau > 0 ? tau + 2 - fftau : 0;
                c.c0_pre1 = h64(PNS_F64_CUM_INFLOW, c.c0_pre_i1);
            }
        }
    }
    c.metric = nullptr;
    c.route_blocks = 0; c.route_t = 0; c.route_recompute = 0;
    c.route_all = 1;
    const int64_t stride = io ? io->draw_row_stride : 0;
    c.draw_b = (io && io->draw_b) ? io->draw_b + (size_t)(stride * row_flows) * 3 * c.row32 : nullptr;
    c.draw_n = (io && io->draw_n) ? io->draw_n + (size_t)(stride * row_update) * c.row32 : nullptr;
    c.draw_exp = (io && io->draw_exp && mode == PNS_RNG_TABLE)
                     ? io->draw_exp + (size_t)(stride * row_flows) * net->n_opts * net->replicas : nullptr;
    return c;
}

int check_common(const pns_net* net, const pns_state* st, int t) {
    if (!net || !st) return fail("null net/state");
    if (net->abi_version != PNS_ABI_VERSION) return fail("pns_net.abi_version mismatch");
    if (t < 1 || t > net->sim_steps) return fail("time step out of range [1, sim_steps]");
    if (net->n_links % 2) return fail("links must come in forward/reverse pairs");
    return 0;
}

unsigned blocks_for(size_t n) { return (unsigned)((n + kBlock - 1) / kBlock); }

int launched(const char* what) {
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail(what, e);
}

int check_step_io(const pns_net* net, const pns_step_io* io, int rng_mode) {
    if (rng_mode != PNS_RNG_TABLE && rng_mode != PNS_RNG_PHILOX) return fail("TABLE or PHILOX mode only");
    if (rng_mode == PNS_RNG_TABLE && !(io && io->draw_b)) return fail("TABLE mode needs draw_b");
    if (net->n_demand_rows > 0 && !(io && io->demand)) return fail("demand table missing");
    if (net->n_routed > 0 && !(io && io->od_w)) return fail("od weight table missing");
    return 0;
}

template <bool R1, int PHASE>
void launch_pair_mode(size_t n, cudaStream_t s, const Ctx& c) {
    if (c.mode == PNS_RNG_PHILOX) PNS_LAUNCH_CHAIN((k_link_pair<R1, PHASE, PNS_RNG_PHILOX>), blocks_for(n), kBlock, s, c);
    else if (c.mode == PNS_RNG_TABLE) PNS_LAUNCH_CHAIN((k_link_pair<R1, PHASE, PNS_RNG_TABLE>), blocks_for(n), kBlock, s, c);
    else PNS_LAUNCH_CHAIN((k_link_pair<R1, PHASE, PNS_RNG_REQUEST>), blocks_for(n), kBlock, s, c);
}
template <bool R1>
void launch_pair_phase(size_t n, cudaStream_t s, const Ctx& c) {
    if (c.phase == (PH_UPDATE | PH_FLOWS)) launch_pair_mode<R1, PH_UPDATE | PH_FLOWS>(n, s, c);
    else if (c.phase == PH_UPDATE) launch_pair_mode<R1, PH_UPDATE>(n, s, c);
    else if (c.phase == PH_FLOWS) launch_pair_mode<R1, PH_FLOWS>(n, s, c);
}
#ifndef PNS_HOST_EMULATION
template <int PHASE>
void launch_lane_mode(size_t n_links, cudaStream_t s, const Ctx& c) {
    const unsigned nb = (unsigned)((n_links + PNS_LANE_BLOCK - 1) / PNS_LANE_BLOCK) + (unsigned)c.route_blocks;
    if (c.n.n_classes == 1) {
        if (c.mode == PNS_RNG_PHILOX) PNS_LAUNCH_CHAIN((k_link_lane<PHASE, PNS_RNG_PHILOX, true>), nb, PNS_LANE_BLOCK, s, c);
        else if (c.mode == PNS_RNG_TABLE) PNS_LAUNCH_CHAIN((k_link_lane<PHASE, PNS_RNG_TABLE, true>), nb, PNS_LANE_BLOCK, s, c);
        else PNS_LAUNCH_CHAIN((k_link_lane<PHASE, PNS_RNG_REQUEST, true>), nb, PNS_LANE_BLOCK, s, c);
    } else {
        if (c.mode == PNS_RNG_PHILOX) PNS_LAUNCH_CHAIN((k_link_lane<PHASE, PNS_RNG_PHILOX, false>), nb, PNS_LANE_BLOCK, s, c);
        else if (c.mode == PNS_RNG_TABLE) PNS_LAUNCH_CHAIN((k_link_lane<PHASE, PNS_RNG_TABLE, false>), nb, PNS_LANE_BLOCK, s, c);
        else PNS_LAUNCH_CHAIN((k_link_lane<PHASE, PNS_RNG_REQUEST, false>), nb, PNS_LANE_BLOCK, s, c);
    }
}
#endif
#ifndef PNS_HOST_EMULATION
// batched replicas on the GPU: one thread per (directed link, replica), optionally with the environment riding along
template <int PHASE, bool ENV>
void launch_rep_mode(const pns_net* net, cudaStream_t s, const EnvCtx& x0) {
    const dim3 nb((unsigned)(((size_t)net->replicas + 31) / 32), (unsigned)(net->n_links / 2) + (unsigned)x0.c.route_blocks);
    const bool one = net->n_classes == 1;
    EnvCtx x = x0;
    {   // prefetch distance: the corridors half a resident wave of CTAs covers (148 SMs x PNS_REP_MIN_BLOCKS / 2), plus
        // one (measured at 8192 replicas: off 113.0, quarter.. wave 108.3 / 109.5 / 112.6 us per environment step)
        static const int wave = getenv("PNS_REP_PF_WAVE") ? atoi(getenv("PNS_REP_PF_WAVE")) : 148 * PNS_REP_MIN_BLOCKS / 2;
        const unsigned ahead = wave > 0 ? (unsigned)((wave + nb.x - 1) / nb.x) + 1u : 0u;
        if (ahead && ahead < (unsigned)(net->n_links / 2)) {
            x.pf_pairs = ahead;
            x.pf_off = (size_t)ahead * 2u * (size_t)net->replicas;
        }
    }
    if (x.c.mode == PNS_RNG_PHILOX) {
        if (one) PNS_LAUNCH_CHAIN((k_link_rep<PHASE, PNS_RNG_PHILOX, true, ENV>), nb, kRepBlock, s, x);
        else PNS_LAUNCH_CHAIN((k_link_rep<PHASE, PNS_RNG_PHILOX, false, ENV>), nb, kRepBlock, s, x);
    } else {
        if (one) PNS_LAUNCH_CHAIN((k_link_rep<PHASE, PNS_RNG_TABLE, true, ENV>), nb, kRepBlock, s, x);
        else PNS_LAUNCH_CHAIN((k_link_rep<PHASE, PNS_RNG_TABLE, false, ENV>), nb, kRepBlock, s, x);
    }
}
bool rep_kernel_applies(const pns_net* net, int mode) {
    return net->replicas > 1 && mode != PNS_RNG_REQUEST && !getenv("PNS_PAIR_THREADS");
}
// the environment part of a launch: null members = nothing to do in that phase
struct EnvRide {
    const pns_env* env;
    const float* actions;     // FLOWS phase
    float *obs, *reward, *cum_reward;   // UPDATE phase
};
void launch_rep(const pns_net* net, cudaStream_t s, const Ctx& c, const EnvRide* ride, bool with_route) {
    EnvCtx x;
    memset(&x, 0, sizeof x);
    x.c = c;
    if (with_route && (c.phase & PH_FLOWS)) {
        x.c.route_blocks = ((c.route_all ? net->n_rows : net->n_dyn_rows) + 1) / 2;   // extra grid rows: two route rows per CTA
        x.c.route_t = c.t_flows;
        x.c.route_recompute = (c.phase & PH_UPDATE) ? 1 : 0;
    }
    const bool env_flows = ride && ride->actions && (c.phase & PH_FLOWS);
    const bool env_update = ride && ride->obs && (c.phase & PH_UPDATE);
    if (env_flows || env_update) {
        x.env = *ride->env;
        x.actions = env_flows ? ride->actions : nullptr;
        if (env_update) { x.obs = ride->obs; x.reward = ride->reward; x.cum_reward = ride->cum_reward; }
        if (c.phase == PH_FLOWS) launch_rep_mode<PH_FLOWS, true>(net, s, x);
        else if (c.phase == PH_UPDATE) launch_rep_mode<PH_UPDATE, true>(net, s, x);
        else launch_rep_mode<PH_UPDATE | PH_FLOWS, true>(net, s, x);
        return;
    }
    if (c.phase == (PH_UPDATE | PH_FLOWS)) launch_rep_mode<PH_UPDATE | PH_FLOWS, false>(net, s, x);
    else if (c.phase == PH_UPDATE) launch_rep_mode<PH_UPDATE, false>(net, s, x);
    else if (c.phase == PH_FLOWS) launch_rep_mode<PH_FLOWS, false>(net, s, x);
}
#else
struct EnvRide { const pns_env* env; const float* actions; float *obs, *reward, *cum_reward; };
#endif
void launch_pair(const pns_net* net, size_t n, cudaStream_t s, const Ctx& c, const EnvRide* ride = nullptr,
                 bool with_route = false) {
#ifndef PNS_HOST_EMULATION
    if (rep_kernel_applies(net, c.mode)) { launch_rep(net, s, c, ride, with_route); return; }
    if (net->replicas == 1 && !getenv("PNS_PAIR_THREADS")) {      // single replica: one thread per directed link
        Ctx cl = c;
        if (with_route && (c.phase & PH_FLOWS)) {
            constexpr int rows_per_cta = PNS_LANE_BLOCK / 32;               // route choice: one warp per row
            cl.route_blocks = ((c.route_all ? net->n_rows : net->n_dyn_rows) + rows_per_cta - 1) / rows_per_cta;
            cl.route_t = c.t_flows;
            cl.route_recompute = (c.phase & PH_UPDATE) ? 1 : 0;
        }
        if (c.phase == (PH_UPDATE | PH_FLOWS)) launch_lane_mode<PH_UPDATE | PH_FLOWS>(2 * n, s, cl);
        else if (c.phase == PH_UPDATE) launch_lane_mode<PH_UPDATE>(2 * n, s, cl);
        else if (c.phase == PH_FLOWS) launch_lane_mode<PH_FLOWS>(2 * n, s, cl);
        return;
    }
#endif
    if (net->replicas == 1) launch_pair_phase<true>(n, s, c);
    else launch_pair_phase<false>(n, s, c);
}
void launch_node(const pns_net* net, size_t n, cudaStream_t s, const Ctx& c) {
    const bool routed = net->n_routed > 0;
#ifndef PNS_HOST_EMULATION
    static const int cols_max = getenv("PNS_NODE_COLS_MAX") ? atoi(getenv("PNS_NODE_COLS_MAX")) : PNS_NODE_COLS_MAX_REPLICAS;
    if (net->replicas > 1 && net->replicas <= cols_max && !getenv("PNS_PAIR_THREADS")) {
        // batched replicas, small batches: one warp per node slot (latency); large batches keep one thread per node
        // (fewer instructions per node)
        dim3 grid((unsigned)(((size_t)net->replicas + 31) / 32), (unsigned)net->n_nodes);
        if (net->nd_stride > 4 && net->nd_cols_order)          // nodes of <= 4 slots share a CTA in twos
            grid.y = (unsigned)((net->n_nodes_small + 1) / 2 + (net->n_nodes - net->n_nodes_small));
        if (net->nd_stride <= 4) {
            if (routed) PNS_LAUNCH_CHAIN((k_node_cols<true, 4>), grid, 128, s, c);
            else PNS_LAUNCH_CHAIN((k_node_cols<false, 4>), grid, 128, s, c);
        } else {
            if (routed) PNS_LAUNCH_CHAIN((k_node_cols<true, 8>), grid, 256, s, c);
            else PNS_LAUNCH_CHAIN((k_node_cols<false, 8>), grid, 256, s, c);
        }
        return;
    }
#endif
    const unsigned nb = (unsigned)((n + PNS_NODE_BLOCK - 1) / PNS_NODE_BLOCK);
    if (net->replicas == 1) {
        if (routed) PNS_LAUNCH_CHAIN((k_node_flows<true, true>), nb, PNS_NODE_BLOCK, s, c);
        else PNS_LAUNCH_CHAIN((k_node_flows<true, false>), nb, PNS_NODE_BLOCK, s, c);
    } else {
        if (routed) PNS_LAUNCH_CHAIN((k_node_flows<false, true>), nb, PNS_NODE_BLOCK, s, c);
        else PNS_LAUNCH_CHAIN((k_node_flows<false, false>), nb, PNS_NODE_BLOCK, s, c);
    }
}

// warps per CTA and dynamic shared memory of the LP kernels for programs of up to max_m slots
struct LpLaunch { unsigned warps; size_t smem; };
LpLaunch lp_launch_shape(int max_m) {
    const size_t per_warp = node_lp_warp_doubles(max_m) * sizeof(double);
    size_t warps = (96 * 1024) / per_warp;
    warps = warps < 1 ? 1 : (warps > 8 ? 8 : warps);
    return {(unsigned)warps, warps * per_warp};
}
#ifndef PNS_HOST_EMULATION
template <typename K>
int lp_allow_smem(K kern, size_t smem) {
    if (smem <= 48 * 1024) return 0;
    if (smem > 227 * 1024) return fail("linear program too large for shared memory");
    const cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return e == cudaSuccess ? 0 : fail("cudaFuncSetAttribute(k_node_lp)", e);
}
#endif
// nodes of kind 2 ('optimal' node model), after the classic pass
int launch_node_lp(const pns_net* net, cudaStream_t s, const Ctx& c) {
    if (net->n_lp_nodes <= 0) return 0;
    const size_t n = (size_t)net->n_lp_nodes * net->replicas;
#ifdef PNS_HOST_EMULATION
    PNS_LAUNCH(k_node_lp, (unsigned)n, 1, s, c);
#else
    const LpLaunch z = lp_launch_shape(net->lp_max_m);
    if (lp_allow_smem(k_node_lp, z.smem)) return 1;
    pns_launch_chain(k_node_lp, dim3((unsigned)((n + z.warps - 1) / z.warps)), z.warps * 32, s, &c, z.smem);
#endif
    return 0;
}

struct StepSizes { size_t n_pair, n_grp, n_node; };
StepSizes sizes_of(const pns_net* net) {
    StepSizes z;
    z.n_pair = (size_t)(net->n_links / 2) * net->replicas;
    z.n_grp = (size_t)net->n_rows * net->replicas;       // route choice: one thread per routed upstream slot and replica
    z.n_node = (size_t)net->n_nodes * net->replicas;
    return z;
}

constexpr size_t kMetricRow = (size_t)PNS_METRIC_SLOTS * PNS_METRIC_STRIDE;   // doubles per step in the metric buffers
#ifndef PNS_HOST_EMULATION
// copy stream + event ring of pns_step_streamed, one per device (created on first use, kept for the life of the
// process; a re-recorded event does not disturb waits that were enqueued on its earlier record).  The table is
// guarded by a mutex: ctypes callers release the GIL, so two host threads may be inside the library at once.
struct SideStream {
    cudaStream_t stream = nullptr, stream2 = nullptr;    // host->device copies / device->host copies
    cudaEvent_t ring[256];
    int next = 0;
    bool ok = false;
    cudaEvent_t event() { next = (next + 1) & 255; return ring[next]; }
};
SideStream* side_stream_for_current_device() {
    static SideStream table[64];
    static std::mutex guard;
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(guard);
    SideStream& s = table[d];
    if (!s.ok) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (auto& e : s.ring)
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        s.ok = true;
    }
    return &s;
}
#endif
struct Streamed {           // per-step host traffic of pns_step_streamed
    const double* host_demand;   // pinned [rows][n_demand_rows*R]
    double* dev_metric;          // [n_steps][PNS_METRIC_SLOTS * PNS_METRIC_STRIDE]
    double* host_metric;         // pinned, same shape
};

int step_impl(const pns_net* net, const pns_state* st, const pns_step_io* io, int t0, int n_steps, int rng_mode,
              cudaStream_t s, double* ms, int64_t* launches, const Streamed* sx = nullptr,
              const EnvRide* ride = nullptr) {
    if (n_steps <= 0) return 0;
    if (check_common(net, st, t0) || check_common(net, st, t0 + n_steps - 1)) return 1;
    if (check_step_io(net, io, rng_mode)) return 1;
    ensure_sampler_tables(s);
    const StepSizes z = sizes_of(net);
#ifndef PNS_HOST_EMULATION
    constexpr int kSyncGroup = 8;
    cudaEvent_t demand_next = nullptr;
    int results_copied = 0;
    SideStream* side_p = sx ? side_stream_for_current_device() : nullptr;
    if (sx && !side_p) return fail("pns_step_streamed: copy stream / events could not be created", cudaGetLastError());
    SideStream dummy;
    SideStream& side = side_p ? *side_p : dummy;
    if (sx) {
        cudaMemsetAsync(sx->dev_metric, 0, (size_t)n_steps * kMetricRow * sizeof(double), s);
        cudaEvent_t e = side.event();                  // the side stream starts after everything queued so far
        cudaEventRecord(e, s);
        cudaStreamWaitEvent(side.stream, e, 0);
    }
    cudaEvent_t* ev = nullptr;
    const int per_step = 4;   // before pair | after pair | after route | after node
    if (ms) {
        ev = (cudaEvent_t*)malloc(sizeof(cudaEvent_t) * (per_step * (n_steps + 1)));
        for (int i = 0; i < per_step * (n_steps + 1); ++i) cudaEventCreate(&ev[i]);
    }
#define PNS_MARK(k, j) do { if (ev) cudaEventRecord(ev[per_step * (k) + (j)], s); } while (0)
#else
    (void)ms; (void)launches;
#define PNS_MARK(k, j) do { } while (0)
#endif
    // launch k (0..n_steps): pair kernel = UPDATE(t0+k-1) [k>0] + FLOWS(t0+k) [k<n_steps]; then route+node(t0+k)
    for (int k = 0; k <= n_steps; ++k) {
        const int phase = (k > 0 ? PH_UPDATE : 0) | (k < n_steps ? PH_FLOWS : 0);
        Ctx cp = make_ctx(net, st, io, phase, t0 + k - 1, t0 + k, rng_mode, k - 1, k);
        // constant route rows are evaluated on step 1 and whenever the caller asks (first step after an init)
        const int route_all = (t0 + k == 1 || (k == 0 && io && io->route_all_rows)) ? 1 : 0;
        cp.route_all = route_all;
        const size_t n_route = (size_t)(route_all ? net->n_rows : net->n_dyn_rows) * net->replicas;
#ifndef PNS_HOST_EMULATION
        // streamed runs: the single-replica link kernel accumulates the step's pedestrian count itself
        const bool lane_metric = sx && k > 0 && net->replicas == 1 && !getenv("PNS_PAIR_THREADS");
        if (lane_metric) cp.metric = sx->dev_metric + (size_t)(k - 1) * kMetricRow;
#endif
#ifndef PNS_HOST_EMULATION
        // Host traffic of streamed runs rides on a second stream: every step still has its own H2D copy (its
        // demand row) and its own D2H copy (its result), but the two streams meet only once per group of
        // kSyncGroup steps -- an event between two kernels of the chain costs the programmatic overlap of that
        // launch pair, a copy there additionally its own latency.  Demand rows are copied one group ahead.
        if (sx && k < n_steps && k % kSyncGroup == 0 && net->n_demand_rows) {
            const size_t row = (size_t)net->n_demand_rows * net->replicas;
            auto copy_group = [&](int first) {          // demand rows of steps t0+first .. t0+first+kSyncGroup-1
                for (int j = first; j < first + kSyncGroup && j < n_steps; ++j) {
                    const size_t off = (size_t)(t0 + j - 1) * row;
                    cudaMemcpyAsync(const_cast<double*>(io->demand) + off, sx->host_demand + off, row * sizeof(double),
                                    cudaMemcpyHostToDevice, side.stream);
                }
                cudaEvent_t e = side.event();
                cudaEventRecord(e, side.stream);
                return e;
            };
            if (k == 0) demand_next = copy_group(0);
            cudaStreamWaitEvent(s, demand_next, 0);     // this group's rows are on the device
            if (k + kSyncGroup < n_steps) demand_next = copy_group(k + kSyncGroup);
        }
#endif
        // GPU link kernels: the route choice of step t0+k is independent of the FLOWS part of the link pass and needs
        // only a few pedestrian counts of its UPDATE part, so it rides along as extra CTAs of the link launch (not at
        // step 1 of an environment, where it reads the widths this launch sets)
        bool route_rides = false;
#ifndef PNS_HOST_EMULATION
        route_rides = (phase & PH_FLOWS) && n_route && z.n_pair && !ev && !getenv("PNS_ROUTE_SEPARATE") &&
                      !(t0 + k == 1 && ride && ride->actions) &&
                      (rep_kernel_applies(net, rng_mode) || (net->replicas == 1 && !getenv("PNS_PAIR_THREADS")));
#endif
        PNS_MARK(k, 0);
        if (z.n_pair) launch_pair(net, z.n_pair, s, cp, ride, route_rides);
        PNS_MARK(k, 1);
#ifndef PNS_HOST_EMULATION
        if (sx && k > 0) {          // result of step t0+k-1: partial sums reduced on the device
            double* row = sx->dev_metric + (size_t)(k - 1) * kMetricRow;
            if (!lane_metric) k_metric_pedestrians<<<148 * 4, 256, 0, s>>>(cp, row);
            if (k % kSyncGroup == 0 || k == n_steps) {   // ... and copied to the host, one copy per step
                cudaEvent_t e = side.event();
                cudaEventRecord(e, s);
                cudaStreamWaitEvent(side.stream, e, 0);
                for (int j = results_copied; j < k; ++j)
                    cudaMemcpyAsync(sx->host_metric + (size_t)j * kMetricRow, sx->dev_metric + (size_t)j * kMetricRow,
                                    kMetricRow * sizeof(double), cudaMemcpyDeviceToHost, side.stream);
                results_copied = k;
            }
        }
#endif
        if (k == n_steps) break;
        Ctx cn = make_ctx(net, st, io, 0, t0 + k, t0 + k, rng_mode, k, k);
        cn.route_all = route_all;
        if (n_route && !route_rides) PNS_LAUNCH_CHAIN(k_route_fractions, blocks_for(n_route), kBlock, s, cn);
        PNS_MARK(k, 2);
        if (z.n_node) launch_node(net, z.n_node, s, cn);
        if (launch_node_lp(net, s, cn)) return 1;
        PNS_MARK(k, 3);
    }
#undef PNS_MARK
#ifndef PNS_HOST_EMULATION
    if (sx) {                                          // the caller synchronises `s` only
        cudaEvent_t e = side.event();
        cudaEventRecord(e, side.stream);
        cudaStreamWaitEvent(s, e, 0);
    }
    if (ev) {
        const cudaError_t err = cudaStreamSynchronize(s);
        for (int k = 0; k <= n_steps && err == cudaSuccess; ++k) {
            float dt = 0.f;
            cudaEventElapsedTime(&dt, ev[per_step * k], ev[per_step * k + 1]);
            if (z.n_pair) { ms[0] += dt; launches[0] += 1; }
            if (k == n_steps) break;
            cudaEventElapsedTime(&dt, ev[per_step * k + 1], ev[per_step * k + 2]);
            if (z.n_grp) { ms[1] += dt; launches[1] += 1; }
            cudaEventElapsedTime(&dt, ev[per_step * k + 2], ev[per_step * k + 3]);
            if (z.n_node) { ms[2] += dt; launches[2] += 1; }
        }
        for (int i = 0; i < per_step * (n_steps + 1); ++i) cudaEventDestroy(ev[i]);
        free(ev);
        if (err != cudaSuccess) return fail("pns_step_profiled", err);
    }
#endif
    return launched("pns_step");
}

}  // namespace

extern "C" {

int pns_abi_version(void) { return PNS_ABI_VERSION; }
int pns_lane_block_size(void) { return PNS_LANE_BLOCK; }
const char* pns_last_error(void) { return g_err; }

int pns_state_init(const pns_net* net, const pns_state* st, void* stream) {
    if (!net || !st) return fail("null net/state");
    if (net->abi_version != PNS_ABI_VERSION) return fail("pns_net.abi_version mismatch");
    if (cudaMemsetAsync(st->err, 0, sizeof(int32_t) * net->replicas, (cudaStream_t)stream) != cudaSuccess)
        return fail("memset err", cudaGetLastError());
    ensure_sampler_tables((cudaStream_t)stream);
    const Ctx c = make_ctx(net, st, nullptr, 0, 1, 1, PNS_RNG_TABLE, 0, 0);
    PNS_LAUNCH(k_state_init, 148 * 8, 256, (cudaStream_t)stream, c);
    return launched("k_state_init");
}

int pns_link_flows(const pns_net* net, const pns_state* st, const pns_step_io* io, int t, int rng_mode,
                   void* stream) {
    if (check_common(net, st, t)) return 1;
    if (rng_mode == PNS_RNG_TABLE && !(io && io->draw_b)) return fail("TABLE mode needs draw_b");
    if (rng_mode == PNS_RNG_REQUEST &&
        !(io && io->req_kind && io->req_n1 && io->req_rf && io->req_sval && io->req_n3))
        return fail("REQUEST mode needs the req_* buffers");
    const size_t n = sizes_of(net).n_pair;
    if (n == 0) return 0;
    ensure_sampler_tables((cudaStream_t)stream);
    const Ctx c = make_ctx(net, st, io, PH_FLOWS, t, t, rng_mode, 0, 0);
    launch_pair(net, n, (cudaStream_t)stream, c);
    if (rng_mode == PNS_RNG_REQUEST && io->req_exp && net->n_routed > 0) {
        const size_t nr = sizes_of(net).n_grp;
        PNS_LAUNCH(k_route_fractions, blocks_for(nr), kBlock, (cudaStream_t)stream, c);
    }
    return launched("k_link_pair[flows]");
}

int pns_route_fractions(const pns_net* net, const pns_state* st, const pns_step_io* io, int t, int rng_mode,
                        void* stream) {
    if (check_common(net, st, t)) return 1;
    const size_t n = sizes_of(net).n_grp;
    if (n == 0) return 0;
    if (rng_mode == PNS_RNG_REQUEST && !(io && io->req_exp)) return fail("REQUEST mode needs req_exp");
    if (rng_mode != PNS_RNG_REQUEST && !(io && io->od_w)) return fail("od weight table missing");
    const Ctx c = make_ctx(net, st, io, 0, t, t, rng_mode, 0, 0);
    PNS_LAUNCH(k_route_fractions, blocks_for(n), kBlock, (cudaStream_t)stream, c);
    return launched("k_route_fractions");
}

int pns_node_flows(const pns_net* net, const pns_state* st, const pns_step_io* io, int t, void* stream) {
    if (check_common(net, st, t)) return 1;
    if (net->n_demand_rows > 0 && !(io && io->demand)) return fail("demand table missing");
    if (net->n_routed > 0 && !(io && io->od_w)) return fail("od weight table missing");
    const size_t n = sizes_of(net).n_node;
    if (n == 0) return 0;
    const Ctx c = make_ctx(net, st, io, 0, t, t, PNS_RNG_TABLE, 0, 0);
    launch_node(net, n, (cudaStream_t)stream, c);
    if (launch_node_lp(net, (cudaStream_t)stream, c)) return 1;
    return launched("k_node_flows");
}

int pns_link_update(const pns_net* net, const pns_state* st, const pns_step_io* io, int t, int rng_mode,
                    void* stream) {
    if (check_common(net, st, t)) return 1;
    const size_t n = sizes_of(net).n_pair;
    if (n == 0) return 0;
    const Ctx c = make_ctx(net, st, io, PH_UPDATE, t, t, rng_mode, 0, 0);
    launch_pair(net, n, (cudaStream_t)stream, c);
    return launched("k_link_pair[update]");
}

int pns_step(const pns_net* net, const pns_state* st, const pns_step_io* io, int t0, int n_steps, int rng_mode,
             void* stream) {
    return step_impl(net, st, io, t0, n_steps, rng_mode, (cudaStream_t)stream, nullptr, nullptr);
}

int pns_step_profiled(const pns_net* net, const pns_state* st, const pns_step_io* io, int t0, int n_steps,
                      int rng_mode, void* stream, double* ms, int64_t* launches) {
    return step_impl(net, st, io, t0, n_steps, rng_mode, (cudaStream_t)stream, ms, launches);
}

int pns_env_apply_actions(const pns_net* net, const pns_state* st, const pns_env* env, const float* actions,
                          void* stream) {
    if (!net || !st || !env || !actions) return fail("pns_env_apply_actions: null argument");
    if (net->abi_version != PNS_ABI_VERSION) return fail("pns_net.abi_version mismatch");
    const size_t n = (size_t)env->n_act * net->replicas;
    if (n == 0) return 0;
    EnvCtx x;
    x.c = make_ctx(net, st, nullptr, 0, 1, 1, PNS_RNG_TABLE, 0, 0);
    x.env = *env; x.actions = actions; x.obs = nullptr; x.reward = nullptr; x.cum_reward = nullptr;
    PNS_LAUNCH_CHAIN(k_env_actions, blocks_for(n), kBlock, (cudaStream_t)stream, x);
    return launched("k_env_actions");
}

int pns_env_observe(const pns_net* net, const pns_state* st, const pns_env* env, int t, float* obs, float* reward,
                    void* stream) {
    if (!net || !st || !env || !obs || !reward) return fail("pns_env_observe: null argument");
    if (net->abi_version != PNS_ABI_VERSION) return fail("pns_net.abi_version mismatch");
    if (t < 0 || t > net->sim_steps) return fail("pns_env_observe: row out of range");
    if (env->n_reward_links > PNS_MAX_DEGREE) return fail("pns_env_observe: too many reward links");
    const size_t n = (size_t)env->n_obs * net->replicas + (size_t)net->replicas;
    EnvCtx x;
    x.c = make_ctx(net, st, nullptr, 0, t, t, PNS_RNG_TABLE, 0, 0);
    x.env = *env; x.actions = nullptr; x.obs = obs; x.reward = reward; x.cum_reward = nullptr;
    PNS_LAUNCH_CHAIN(k_env_observe, blocks_for(n), kBlock, (cudaStream_t)stream, x);
    return launched("k_env_observe");
}

namespace {
bool env_fused_path(const pns_net* net, const pns_env* env, int rng_mode) {
#ifndef PNS_HOST_EMULATION
    return env->lk_act && env->lk_obs_ptr && env->lk_reward && env->reward_count && rep_kernel_applies(net, rng_mode);
#else
    (void)net; (void)env; (void)rng_mode;
    return false;
#endif
}

// one environment step
int env_step_impl(const pns_net* net, const pns_state* st, const pns_step_io* io, const pns_env* env,
                  const float* actions, int t, int rng_mode, float* obs, float* reward, float* cum_reward,
                  cudaStream_t stream) {
    if (env_fused_path(net, env, rng_mode)) {
        // batched replicas: actions ride in the FLOWS launch, observations and reward in the UPDATE launch
        EnvRide ride;
        memset(&ride, 0, sizeof ride);
        ride.env = env; ride.actions = (actions && env->n_act > 0) ? actions : nullptr;
        ride.obs = obs; ride.reward = reward; ride.cum_reward = cum_reward;
        return step_impl(net, st, io, t, 1, rng_mode, stream, nullptr, nullptr, nullptr, &ride);
    }
    if (actions && env->n_act > 0 && pns_env_apply_actions(net, st, env, actions, stream)) return 1;
    if (step_impl(net, st, io, t, 1, rng_mode, stream, nullptr, nullptr)) return 1;
    const size_t n = (size_t)env->n_obs * net->replicas + (size_t)net->replicas;
    EnvCtx x;
    memset(&x, 0, sizeof x);
    x.c = make_ctx(net, st, nullptr, 0, t, t, PNS_RNG_TABLE, 0, 0);
    x.env = *env; x.obs = obs; x.reward = reward; x.cum_reward = cum_reward;
    PNS_LAUNCH_CHAIN(k_env_observe, blocks_for(n), kBlock, stream, x);
    return launched("k_env_observe");
}
}  // namespace

int pns_env_step(const pns_net* net, const pns_state* st, const pns_step_io* io, const pns_env* env,
                 const float* actions, int t, int rng_mode, float* obs, float* reward, float* cum_reward,
                 void* stream) {
    if (!net || !st || !env || !obs || !reward) return fail("pns_env_step: null argument");
    if (env->n_reward_links > PNS_MAX_DEGREE) return fail("pns_env_step: too many reward links");
    return env_step_impl(net, st, io, env, actions, t, rng_mode, obs, reward, cum_reward, (cudaStream_t)stream);
}

int pns_env_rollout(const pns_net* net, const pns_state* st, const pns_step_io* io, const pns_env* env, int t0,
                    int n_steps, int rng_mode, const float* actions, float* obs, float* reward, float* cum_reward,
                    const float* host_actions, float* host_obs, float* host_reward, void* stream) {
    if (!net || !st || !env || !obs || !reward) return fail("pns_env_rollout: null argument");
    if (env->n_reward_links > PNS_MAX_DEGREE) return fail("pns_env_rollout: too many reward links");
    if (n_steps <= 0) return 0;
    if (t0 < 1 || t0 + n_steps - 1 > net->sim_steps) return fail("pns_env_rollout: steps out of range");
    const bool host = host_actions || host_obs || host_reward;
    if (host && !(host_obs && host_reward && (host_actions || env->n_act == 0)))
        return fail("pns_env_rollout: host form needs all host buffers");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t na = (size_t)net->replicas * env->n_act, no = (size_t)net->replicas * env->n_obs, nr = (size_t)net->replicas;
#ifndef PNS_HOST_EMULATION
    SideStream* side = host ? side_stream_for_current_device() : nullptr;
    if (host && !side) return fail("pns_env_rollout: copy streams / events could not be created", cudaGetLastError());
    cudaEvent_t up_done[2] = {nullptr, nullptr}, step_done[2] = {nullptr, nullptr}, down_done[2] = {nullptr, nullptr};
    if (host) {
        cudaEvent_t e = side->event();                 // the copy streams start after everything queued so far
        cudaEventRecord(e, s);
        cudaStreamWaitEvent(side->stream, e, 0);
        cudaStreamWaitEvent(side->stream2, e, 0);
        if (na) cudaMemcpyAsync(const_cast<float*>(actions), host_actions, na * sizeof(float), cudaMemcpyHostToDevice, side->stream);
        up_done[0] = side->event();
        cudaEventRecord(up_done[0], side->stream);
    }
#else
    if (host) return fail("pns_env_rollout: host form needs the CUDA build");
#endif
    for (int k = 0; k < n_steps; ++k) {
        const int slot = host ? (k & 1) : k;
#ifndef PNS_HOST_EMULATION
        if (host) {
            cudaStreamWaitEvent(s, up_done[k & 1], 0);                       // this step's actions are on the device
            if (k >= 2) cudaStreamWaitEvent(s, down_done[k & 1], 0);         // results of step k-2 have left this slot
        }
#endif
        if (env_step_impl(net, st, io, env, env->n_act ? actions + (size_t)slot * na : nullptr, t0 + k, rng_mode,
                          obs + (size_t)slot * no, reward + (size_t)slot * nr, cum_reward, s))
            return 1;
#ifndef PNS_HOST_EMULATION
        if (host) {
            step_done[k & 1] = side->event();
            cudaEventRecord(step_done[k & 1], s);
            if (k + 1 < n_steps) {                     // next step's actions (its slot was read by step k-1)
                if (k >= 1) cudaStreamWaitEvent(side->stream, step_done[(k + 1) & 1], 0);
                if (na) cudaMemcpyAsync(const_cast<float*>(actions) + (size_t)((k + 1) & 1) * na, host_actions + (size_t)(k + 1) * na,
                                        na * sizeof(float), cudaMemcpyHostToDevice, side->stream);
                up_done[(k + 1) & 1] = side->event();
                cudaEventRecord(up_done[(k + 1) & 1], side->stream);
            }
            cudaStreamWaitEvent(side->stream2, step_done[k & 1], 0);          // this step's results
            cudaMemcpyAsync(host_obs + (size_t)k * no, obs + (size_t)(k & 1) * no, no * sizeof(float), cudaMemcpyDeviceToHost, side->stream2);
            cudaMemcpyAsync(host_reward + (size_t)k * nr, reward + (size_t)(k & 1) * nr, nr * sizeof(float), cudaMemcpyDeviceToHost, side->stream2);
            down_done[k & 1] = side->event();
            cudaEventRecord(down_done[k & 1], side->stream2);
        }
#endif
    }
#ifndef PNS_HOST_EMULATION
    if (host) {                                        // the caller synchronises `stream` only
        cudaEvent_t e1 = side->event(), e2 = side->event();
        cudaEventRecord(e1, side->stream);
        cudaEventRecord(e2, side->stream2);
        cudaStreamWaitEvent(s, e1, 0);
        cudaStreamWaitEvent(s, e2, 0);
    }
#endif
    return launched("pns_env_rollout");
}

int pns_step_streamed(const pns_net* net, const pns_state* st, const pns_step_io* io, int t0, int n_steps,
                      int rng_mode, const double* host_demand, double* dev_metric, double* host_metric, void* stream) {
#ifdef PNS_HOST_EMULATION
    (void)host_demand; (void)dev_metric; (void)host_metric;
    return step_impl(net, st, io, t0, n_steps, rng_mode, (cudaStream_t)stream, nullptr, nullptr);
#else
    if (!host_demand || !dev_metric || !host_metric) return fail("pns_step_streamed: null buffer");
    Streamed sx;
    sx.host_demand = host_demand; sx.dev_metric = dev_metric; sx.host_metric = host_metric;
    return step_impl(net, st, io, t0, n_steps, rng_mode, (cudaStream_t)stream, nullptr, nullptr, &sx);
#endif
}

int pns_env_draw_demand(int sim_steps, int rows, int replicas, uint32_t replica_base, uint64_t seed,
                        const double* bump1, const double* bump2, const double* base, const double* peak,
                        const int32_t* pattern, double* demand, void* stream) {
    if (sim_steps <= 0 || rows <= 0 || replicas <= 0) return 0;
    if (!bump1 || !bump2 || !base || !peak || !pattern || !demand) return fail("pns_env_draw_demand: null argument");
    DemandCtx d;
    d.S = sim_steps; d.rows = rows; d.R = replicas; d.replica_base = replica_base; d.seed = seed;
    d.bump1 = bump1; d.bump2 = bump2; d.base = base; d.peak = peak; d.pattern = pattern; d.out = demand;
    const size_t n = (size_t)(sim_steps + 1) * rows * replicas;
    PNS_LAUNCH(k_demand_draw, blocks_for(n), kBlock, (cudaStream_t)stream, d);
    return launched("k_demand_draw");
}

int pns_env_randomize(const pns_net* net, pns_link_class* classes, int n_base_classes, const int32_t* base_class,
                      int n_change, int32_t* lk_class, double* od_w, int n_demand_rows, const int32_t* row_is_origin,
                      double* dem_base, double* dem_peak, int32_t* dem_pattern, uint64_t seed, uint32_t replica_base,
                      void* stream) {
    if (!net || !classes || !base_class || !lk_class) return fail("pns_env_randomize: null argument");
    if (net->abi_version != PNS_ABI_VERSION) return fail("pns_net.abi_version mismatch");
    if (n_change < 0 || n_change > 32) return fail("pns_env_randomize: at most 32 perturbed corridors per replica");
    if (net->n_od > 0 && !od_w) return fail("pns_env_randomize: od weight table missing");
    if (n_demand_rows > 0 && !(row_is_origin && dem_base && dem_peak && dem_pattern))
        return fail("pns_env_randomize: demand parameter arrays missing");
    ScenarioCtx x;
    x.n_links = net->n_links; x.R = net->replicas; x.n_change = n_change; x.n_base = n_base_classes;
    x.n_od = net->n_od; x.S = net->sim_steps; x.rows = n_demand_rows;
    x.replica_base = replica_base; x.seed = seed; x.unit_time = net->unit_time;
    x.classes = classes; x.lk_class = lk_class; x.base_class = base_class; x.lk_width = net->lk_width;
    x.od_w = od_w; x.dem_base = dem_base; x.dem_peak = dem_peak; x.dem_pattern = dem_pattern;
    x.row_is_origin = row_is_origin;
    PNS_LAUNCH(k_scenario_draw, blocks_for((size_t)net->replicas), kBlock, (cudaStream_t)stream, x);
    return launched("k_scenario_draw");
}

int pns_kpi(const pns_net* net, const pns_state* st, const pns_step_io* io, int t_last, const int32_t* lk_role,
            int any_od_path, double* scratch, double* out, void* stream) {
    if (!net || !st || !lk_role || !scratch || !out) return fail("pns_kpi: null argument");
    if (net->abi_version != PNS_ABI_VERSION) return fail("pns_net.abi_version mismatch");
    if (t_last < 0 || t_last > net->sim_steps) return fail("pns_kpi: row out of range");
    if (net->n_demand_rows > 0 && !(io && io->demand)) return fail("pns_kpi: demand table missing");
    const Ctx c = make_ctx(net, st, io, 0, 1, 1, PNS_RNG_TABLE, 0, 0);
    const size_t n = (size_t)net->n_links * net->replicas;
    if (n) PNS_LAUNCH(k_kpi_links, blocks_for(n), kBlock, (cudaStream_t)stream, c, t_last, scratch);
    PNS_LAUNCH(k_kpi_reduce, blocks_for((size_t)net->replicas), kBlock, (cudaStream_t)stream, c, t_last,
               (const double*)scratch, lk_role, any_od_path, out);
    return launched("k_kpi");
}

int pns_lp_solve(int m, int n, const double* s, const double* r, const double* phi, double w, double* x,
                 double* objective, int32_t* info, void* stream) {
    if (m < 2 || m > PNS_MAX_DEGREE) return fail("pns_lp_solve: 2 <= m <= PNS_MAX_DEGREE");
    if (!s || !r || !phi || !x || !objective || !info) return fail("pns_lp_solve: null argument");
    if (n <= 0) return 0;
    const LpBatch b = {m, n, s, r, phi, w, x, objective, info};
#ifdef PNS_HOST_EMULATION
    PNS_LAUNCH(k_lp_batch, (unsigned)n, 1, (cudaStream_t)stream, b);
#else
    const LpLaunch z = lp_launch_shape(m);
    if (lp_allow_smem(k_lp_batch, z.smem)) return 1;
    k_lp_batch<<<(unsigned)((n + z.warps - 1) / z.warps), z.warps * 32, z.smem, (cudaStream_t)stream>>>(b);
#endif
    return launched("k_lp_batch");
}

int pns_rng_selftest(int kind, int n, const int32_t* n_trials, const double* p, uint64_t seed, int t, int site,
                     int32_t* out_i, double* out_d, void* stream) {
    if (n <= 0) return 0;
    ensure_sampler_tables((cudaStream_t)stream);
    PNS_LAUNCH(k_rng_selftest, (n + 127) / 128, 128, (cudaStream_t)stream, kind, n, n_trials, p, seed, t, site, out_i,
               out_d);
    return launched("k_rng_selftest");
}

}  // extern "C"
