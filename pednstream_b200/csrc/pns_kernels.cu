// B200 (sm_100a) kernels of the Link-Transmission-Model timestep + their C-ABI launchers.
//
// One simulation step t (reference Network.network_loading, src/LTM/network.py:266-287) is four
// passes; all state is time-major structure-of-arrays history in HBM (include/pns_b200.h):
//
//   k_link_flows   one thread per (link pair, replica)   sending + receiving flow at tau = t-1
//   k_route_probs  one thread per (route group, replica) logit P(down | up, od)
//   k_node_flows   one thread per (node, replica)        turning fractions, node model, cum. counts
//   k_link_update  one thread per (link pair, replica)   pedestrians, density, speed, travel time
//
// Nodes are mutually independent within a step (every read is of rows <= t-1 or of values the
// same pass produced for the same link pair), which is what makes the
// step data-parallel.  The replica index is the fastest-varying one, so for batched replicas a
// warp touches 32 consecutive elements of every row it reads or writes.
//
// Numerics: the reference mixes float32 history with float64 counters under numpy-2 scalar
// promotion; every expression below states its precision explicitly and the file is compiled
// with -fmad=false so that no multiply-add is contracted (SURVEY.md "dtype ledger").
#ifdef PNS_HOST_EMULATION
#include "pns_emu.h"   // tests/emu: sequential host build of these kernels for CPU-side unit tests
#else
#include <cuda_runtime.h>
#define PNS_LAUNCH(kern, nblk, nthr, stream, ...) kern<<<(nblk), (nthr), 0, (stream)>>>(__VA_ARGS__)
#endif
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pns_b200.h"
#include "pns_rng.cuh"

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

struct Ctx {
    pns_net n;
    pns_state s;
    pns_step_io io;
    int t;      // step being computed
    int mode;   // PNS_RNG_*
    const int32_t* draw_b;  // row of this step (TABLE)
    const double* draw_n;
};

// ---- addressing ---------------------------------------------------------------------------------
__device__ __forceinline__ size_t row64(const Ctx& c) { return (size_t)c.n.n_cols64 * c.n.replicas; }
__device__ __forceinline__ size_t row32(const Ctx& c) { return (size_t)c.n.n_links * c.n.replicas; }
__device__ __forceinline__ double* H64(const Ctx& c, int f, int t) {
    return c.s.hist64 + ((size_t)f * (c.n.sim_steps + 1) + t) * row64(c);
}
__device__ __forceinline__ float* H32(const Ctx& c, int f, int t) {
    return c.s.hist32 + ((size_t)f * (c.n.sim_steps + 1) + t) * row32(c);
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

struct LinkP {
    double length, width, vf, kc, kj, gamma, act, bi, sigma;
    int fftau, swtau, flags;
};
__device__ __forceinline__ LinkP load_link(const Ctx& c, int l) {
    LinkP p;
    p.length = __ldg(c.n.lk_length + l);
    p.width = __ldg(c.n.lk_width + l);
    p.vf = __ldg(c.n.lk_vf + l);
    p.kc = __ldg(c.n.lk_kc + l);
    p.kj = __ldg(c.n.lk_kj + l);
    p.gamma = __ldg(c.n.lk_gamma + l);
    p.act = __ldg(c.n.lk_act + l);
    p.bi = __ldg(c.n.lk_bi + l);
    p.sigma = __ldg(c.n.lk_sigma + l);
    p.fftau = __ldg(c.n.lk_fftau + l);
    p.swtau = __ldg(c.n.lk_swtau + l);
    p.flags = __ldg(c.n.lk_flags + l);
    return p;
}
__device__ __forceinline__ bool is_sep(const LinkP& p) { return p.flags & 1; }

// area of the walkable surface (link.py:128-131, 454-456); `f64` tells whether the reference
// holds it as a numpy float64 (a separator width assigned from np.clip) -- then density is a
// double-precision divide, otherwise the area is demoted to float32 first.
__device__ __forceinline__ double link_area(const Ctx& c, const LinkP& p, size_t e, bool* f64) {
    if (is_sep(p)) {
        *f64 = c.s.sep_np64[e] != 0;
        return p.length * c.s.widths[2 * row32(c) + e];
    }
    *f64 = false;
    return p.length * p.width;
}
__device__ __forceinline__ float div_by_area(float x, double area, bool f64) {
    return f64 ? (float)((double)x / area) : x / (float)area;
}

// Python min/max on scalars: min(a, b) -> b if b < a else a ; max(a, b) -> b if b > a else a
__device__ __forceinline__ double pymin(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double pymax(double a, double b) { return b > a ? b : a; }
__device__ __forceinline__ float clip01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

struct SendOut {
    double flow;   // sending_flow[tau]
    int kind;      // request kind
    int n1;
    float rf;
    double sval;
};

// Link.get_outflow (link.py:199-214)
__device__ double diffusion_outflow(const Ctx& c, const LinkP& p, size_t e, int tau_idx, int tau, float avg_tt,
                                    int replica) {
    const float F = 1.0f / (1.0f + (float)p.gamma * avg_tt);
    const float u = 1.0f - F;
    const float c1 = F * u;
    const float c2 = F * (float)((double)u * (double)u);                 // powf(u, 2)
    const float c3 = F * (float)(((double)u * (double)u) * (double)u);   // powf(u, 3)
    double in[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = wrap_index(c, tau_idx - tau - k, replica);
        in[k] = i >= 0 ? H64(c, PNS_F64_INFLOW, i)[e] : 0.0;
    }
    const double total = (((double)F * in[0] + (double)c1 * in[1]) + (double)c2 * in[2]) + (double)c3 * in[3];
    const double up = ceil(total);
    return 0.0 > up ? 0.0 : up;
}

// Link.cal_sending_flow at time index tau (link.py:216-370)
__device__ SendOut sending_flow(const Ctx& c, const LinkP& p, int l, size_t e, int tau, float num_self,
                                float num_rev, int replica, const pns::DrawKey& key) {
    SendOut o;
    o.kind = 0; o.n1 = 0; o.rf = 0.0f; o.sval = 0.0; o.flow = 0.0;
    const float dens_self = H32(c, PNS_F32_DENSITY, tau)[e];
    bool a64;
    const double area = link_area(c, p, e, &a64);
    const float dens = is_sep(p) ? dens_self : div_by_area(num_self + num_rev, area, a64);
    const float avg_tt = H32(c, PNS_F32_AVG_TRAVEL_TIME, tau)[e];
    const int lag = __float2int_rn(avg_tt / (float)c.n.unit_time);        // round(), half to even
    if (tau < p.fftau) return o;                                          // link.py:267-269
    if (lag == 0) atomicOr(c.s.err + replica, PNS_ERR_ZERO_LAG);
    const int idx = max(0, tau + 1 - lag);
    const float cong = clip01((dens_self - (float)p.kc) / (float)(p.kj - p.kc));
    const double arrived_raw = H64(c, PNS_F64_CUM_INFLOW, idx)[e] - H64(c, PNS_F64_CUM_OUTFLOW, tau)[e];
    const double arrived = arrived_raw > 0.0 ? arrived_raw : 0.0;
    const double boundary = (double)(cong * num_self) + (double)(1.0f - cong) * arrived;
    const double fgw = c.s.widths[e];
    const double gate_cap = ((fgw * p.kc) * p.vf) * c.n.unit_time;
    double flow = pymin(boundary, gate_cap);
    const double original = flow;
    if (flow > 0.0) {
        const float rf = clip01(dens / (float)p.kj);
        o.rf = rf;
        bool draw = true;
        if (dens <= (float)p.kc) {
            const double spread = diffusion_outflow(c, p, e, tau, lag, avg_tt, replica);
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
            if (c.mode == PNS_RNG_TABLE) {
                flow = (double)c.draw_b[e];
            } else if (c.mode == PNS_RNG_PHILOX) {
                const float p32 = 0.7f + 0.15f * pns::det_pow08(rf);
                flow = (double)pns::binomial_philox(key, 1u, trials, (double)p32);
            } else {
                return o;   // REQUEST: the host draws R1 (and R2, which depends on it)
            }
        }
    }
    o.sval = flow;
    if (c.mode == PNS_RNG_REQUEST) return o;
    if (p.act > 0.0 && flow > 1.0) {                                       // link.py:351-358
        const int trials = (int)floor(flow);
        int stay;
        if (c.mode == PNS_RNG_TABLE) stay = c.draw_b[row32(c) + e];
        else stay = pns::binomial_philox(key, 2u, trials, p.act);
        flow = flow - (double)stay;
    }
    flow = flow > 0.0 ? flow : 0.0;
    const int prev_i = wrap_index(c, tau - 1, replica);
    const double prev = H64(c, PNS_F64_SENDING, prev_i)[e];
    flow = pymin(floor(0.8 * flow + 0.2 * prev), original);               // link.py:363-364
    if (flow < 0.0) atomicOr(c.s.err + replica, PNS_ERR_NEG_SENDING);
    o.flow = flow;
    return o;
}

// Link/Separator.cal_receiving_flow at tau, before the reverse sending flow is subtracted
// (link.py:372-405, 480-507)
__device__ double receiving_flow(const Ctx& c, const LinkP& p, size_t e, int tau, float num_rev, int replica,
                                 const pns::DrawKey& key, int* n3) {
    bool a64;
    const double area = link_area(c, p, e, &a64);
    const double space = p.kj * area;
    const int lag_i = tau + 1 - p.swtau;
    double bound;
    if (is_sep(p)) {
        *n3 = -1;
        if (lag_i < 0) bound = space;
        else bound = (H64(c, PNS_F64_CUM_OUTFLOW, lag_i)[e] + space) - H64(c, PNS_F64_CUM_INFLOW, tau)[e];
    } else {
        const int trials = (int)num_rev;        // numpy casts the float32 count to int64 (truncation)
        *n3 = trials;
        int blockers = 0;
        if (c.mode == PNS_RNG_TABLE) blockers = c.draw_b[2 * row32(c) + e];
        else if (c.mode == PNS_RNG_PHILOX) blockers = pns::binomial_philox(key, 3u, trials, 0.9);
        else return 0.0;
        if (lag_i < 0) {
            bound = space - (double)blockers;
        } else {
            const double x = ((H64(c, PNS_F64_CUM_OUTFLOW, lag_i)[e] + space) - (double)blockers) -
                             H64(c, PNS_F64_CUM_INFLOW, tau)[e];
            bound = x > 0.0 ? x : 0.0;
        }
    }
    if (c.mode == PNS_RNG_REQUEST) return 0.0;
    const double bgw = c.s.widths[row32(c) + e];
    const double gate_cap = ((bgw * p.kc) * p.vf) * c.n.unit_time;
    double flow = pymin(bound, gate_cap);
    flow = pymax(flow, 0.0);
    const int prev_i = wrap_index(c, tau - 1, replica);
    const double prev = H64(c, PNS_F64_RECEIVING, prev_i)[e];
    if (prev >= 0.0) flow = pymin(floor(flow * 0.8 + prev * 0.2), flow);   // link.py:400-401
    return flow;
}

// =================================================================================================
__global__ void __launch_bounds__(kBlock) k_link_flows(const __grid_constant__ Ctx c) {
    const int R = c.n.replicas;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_pairs = (size_t)(c.n.n_links / 2);
    if (gid >= n_pairs * R) return;
    const int pair = (int)(gid / R);
    const int rep = (int)(gid % R);
    const int tau = c.t - 1;
    const int l0 = 2 * pair, l1 = l0 + 1;
    const size_t e0 = (size_t)l0 * R + rep, e1 = e0 + R;
    const LinkP p0 = load_link(c, l0), p1 = load_link(c, l1);
    const float* num = H32(c, PNS_F32_NUM_PED, tau);
    const float n0 = num[e0], n1 = num[e1];
    pns::DrawKey k0, k1;
    k0.t = (uint32_t)c.t; k0.link = (uint32_t)l0; k0.replica = (uint32_t)rep;
    k0.k0 = (uint32_t)c.io.seed; k0.k1 = (uint32_t)(c.io.seed >> 32);
    k1 = k0; k1.link = (uint32_t)l1;

    const SendOut s0 = sending_flow(c, p0, l0, e0, tau, n0, n1, rep, k0);
    const SendOut s1 = sending_flow(c, p1, l1, e1, tau, n1, n0, rep, k1);
    int n3_0, n3_1;
    const double r0 = receiving_flow(c, p0, e0, tau, n1, rep, k0, &n3_0);
    const double r1 = receiving_flow(c, p1, e1, tau, n0, rep, k1, &n3_1);

    if (c.mode == PNS_RNG_REQUEST) {
        c.io.req_kind[e0] = s0.kind; c.io.req_kind[e1] = s1.kind;
        c.io.req_n1[e0] = s0.n1;     c.io.req_n1[e1] = s1.n1;
        c.io.req_rf[e0] = s0.rf;     c.io.req_rf[e1] = s1.rf;
        c.io.req_sval[e0] = s0.sval; c.io.req_sval[e1] = s1.sval;
        c.io.req_n3[e0] = n3_0;      c.io.req_n3[e1] = n3_1;
        return;
    }
    double* snd = H64(c, PNS_F64_SENDING, tau);
    double* rcv = H64(c, PNS_F64_RECEIVING, tau);
    snd[e0] = s0.flow;
    snd[e1] = s1.flow;
    // cal_receiving_flow_with_reverse (link.py:407-416; separators ignore the reverse flow, :509-512)
    const double q0 = is_sep(p0) ? r0 : r0 - s1.flow;
    const double q1 = is_sep(p1) ? r1 : r1 - s0.flow;
    rcv[e0] = pymax(q0, 0.0);
    rcv[e1] = pymax(q1, 0.0);
}

// =================================================================================================
// PathFinder.update_node_turn_probs (path_finder.py:561-589)
__global__ void __launch_bounds__(kBlock) k_route_probs(const __grid_constant__ Ctx c) {
    const int R = c.n.replicas;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)c.n.n_groups * R) return;
    const int g = (int)(gid / R);
    const int rep = (int)(gid % R);
    const int o0 = c.n.rt_opt_ptr[g], o1 = c.n.rt_opt_ptr[g + 1];
    const int n = o1 - o0;
    const bool wide = c.n.rt_grp_has_virtual[g] != 0;   // np.array([... float32 ..., 0]) is float64
    const int tm1 = c.t - 1;
    const int tm2 = wrap_index(c, c.t - 2, rep);
    const float* num = H32(c, PNS_F32_NUM_PED, tm1);
    const float* dens_row = H32(c, PNS_F32_DENSITY, tm1);
    const double* rcv = H64(c, PNS_F64_RECEIVING, tm2);

    float dens[PNS_MAX_DEGREE];
    double cap[PNS_MAX_DEGREE];
    double sum_d = 0.0, sum_c = 0.0;
    for (int k = 0; k < n; ++k) {
        const int l = c.n.rt_opt_link[o0 + k];
        sum_d = k == 0 ? c.n.rt_opt_dist[o0] : sum_d + c.n.rt_opt_dist[o0 + k];
        if (l >= 0) {
            const LinkP p = load_link(c, l);
            const size_t e = (size_t)l * R + rep;
            if (is_sep(p)) {
                dens[k] = dens_row[e];
            } else {
                bool a64;
                const double area = link_area(c, p, e, &a64);
                dens[k] = div_by_area(num[e] + num[(size_t)(l ^ 1) * R + rep], area, a64);
            }
            const double last = rcv[e];
            cap[k] = last >= 0.0 ? last : ((c.s.widths[row32(c) + e] * p.vf) * p.kc) * c.n.unit_time;
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
        const double util = (((c.n.rt_alpha * c.n.rt_opt_dist[o0 + k]) / (sum_d + 1e-6) + crowd_term) -
                             (c.n.rt_omega * cap[k]) / (sum_c + 1e-6)) + c.n.rt_eps;
        ex[k] = exp(-c.n.rt_temp * util);
        sum_e = k == 0 ? ex[0] : sum_e + ex[k];
    }
    for (int k = 0; k < n; ++k) c.s.probs[(size_t)(o0 + k) * R + rep] = ex[k] / sum_e;
}

// =================================================================================================
// Node.assign_flows / solve / update_links (node.py:146-300) + turning fractions
// (path_finder.py:591-715)
__global__ void __launch_bounds__(kBlock) k_node_flows(const __grid_constant__ Ctx c) {
    const int R = c.n.replicas;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)c.n.n_nodes * R) return;
    const int node = (int)(gid / R);
    const int rep = (int)(gid % R);
    const int base = c.n.nd_ptr[node];
    const int m = c.n.nd_ptr[node + 1] - base;
    if (m == 0) return;
    const int t = c.t, tau = c.t - 1;
    const int L = c.n.n_links;

    double s[PNS_MAX_DEGREE], r[PNS_MAX_DEGREE];
    int icol[PNS_MAX_DEGREE], ocol[PNS_MAX_DEGREE];
    const double* snd = H64(c, PNS_F64_SENDING, tau);
    const double* rcv = H64(c, PNS_F64_RECEIVING, tau);
    bool negative = false;
    for (int i = 0; i < m; ++i) {
        icol[i] = c.n.nd_in_col[base + i];
        ocol[i] = c.n.nd_out_col[base + i];
        if (icol[i] >= L) {
            const int row = c.n.nd_dem_row[node];
            s[i] = c.io.demand[(size_t)tau * c.n.n_demand_rows * R + (size_t)row * R + rep];   // node.py:176
        } else {
            s[i] = snd[(size_t)icol[i] * R + rep];
        }
        r[i] = ocol[i] >= L ? 1e6 : rcv[(size_t)ocol[i] * R + rep];                            // node.py:186
        negative |= (s[i] < 0.0) | (r[i] < 0.0);
    }
    if (negative) atomicOr(c.s.err + rep, PNS_ERR_NEG_NODE_FLOW);

    double q_out[PNS_MAX_DEGREE], q_in[PNS_MAX_DEGREE];
    if (c.n.nd_kind[node] == 0) {
        // OneToOneNode.solve (node.py:230-242)
        const double a = fmin(s[0], r[1]), b = fmin(s[1], r[0]);
        q_out[0] = a; q_out[1] = b; q_in[0] = b; q_in[1] = a;
    } else {
        const int e0 = c.n.nd_tf_ptr[node];
        const double* tf;
        const int routed = c.n.nd_routed[node];
        if (routed >= 0) {
            // PathFinder.update_turning_fractions (path_finder.py:591-689)
            double* out = c.s.tf_routed + (size_t)e0 * R + rep;   // element k at out[k*R]
            const int row0 = c.n.rt_routed_row0[routed];
            const int edge0 = c.n.rt_routed_edge0[routed];
            const double* w = c.io.od_w + (size_t)t * c.n.n_od;
            int k = 0;
            for (int i = 0; i < m; ++i) {
                const int ra = c.n.rt_row_ptr[row0 + i], rb = c.n.rt_row_ptr[row0 + i + 1];
                double total = 0.0;
                for (int x = ra; x < rb; ++x) total = total + w[c.n.rt_row_od[x]];
                const double uniform = rb > ra ? 1.0 / (double)(rb - ra) : 0.0;
                double row_sum = 0.0;
                for (int j = 0; j < m - 1; ++j, ++k) {
                    const int ta = c.n.rt_term_ptr[edge0 + k], tb = c.n.rt_term_ptr[edge0 + k + 1];
                    double acc = 0.0;
                    for (int x = ta; x < tb; ++x) {
                        const double od_p = total > 0.0 ? w[c.n.rt_row_od[c.n.rt_term_row_entry[x]]] / total : uniform;
                        acc = acc + c.s.probs[(size_t)c.n.rt_term_opt[x] * R + rep] * od_p;
                    }
                    out[(size_t)k * R] = acc;
                    row_sum = j == 0 ? acc : row_sum + acc;
                }
                // PathFinder.check_fractions (path_finder.py:691-715)
                if (fabs(row_sum - 1.0) > 1e-3) {
                    double* rowp = out + (size_t)(k - (m - 1)) * R;
                    if (row_sum > 1e-6) {
                        for (int j = 0; j < m - 1; ++j) rowp[(size_t)j * R] = rowp[(size_t)j * R] / row_sum;
                    } else {
                        for (int j = 0; j < m - 1; ++j) rowp[(size_t)j * R] = 1.0 / (double)(m - 1);
                    }
                }
            }
            tf = out;
        } else {
            tf = c.s.tf_static + e0;
        }
        const size_t tstride = routed >= 0 ? (size_t)R : 1;
        // RegularNode.solve, 'classic' (node.py:272-300).  P[i][j] = tf[i*(m-1) + (j<i ? j : j-1)]
        double D[PNS_MAX_DEGREE];
        for (int j = 0; j < m; ++j) {
            double acc = 0.0;
            for (int i = 0; i < m; ++i) {
                const double pij = i == j ? 0.0 : tf[(size_t)(i * (m - 1) + (j < i ? j : j - 1)) * tstride];
                const double wij = pij * s[i];
                acc = i == 0 ? wij : acc + wij;      // np.sum(axis=0): rows added in order
            }
            D[j] = acc != 0.0 ? acc : 1e-5;
            q_in[j] = 0.0;
        }
        for (int i = 0; i < m; ++i) {
            double out_i = 0.0;
            for (int j = 0; j < m; ++j) {
                if (i == j) continue;
                const double pij = tf[(size_t)(i * (m - 1) + (j < i ? j : j - 1)) * tstride];
                const double wij = pij * s[i];
                const double supply = r[j] * (wij / D[j]);
                const double f = floor(pymin(wij, supply));
                out_i += f;
                q_in[j] += f;
            }
            q_out[i] = out_i;
        }
        for (int i = 0; i < m; ++i) {
            q_out[i] = fmax(0.0, q_out[i]);
            q_in[i] = fmax(0.0, q_in[i]);
        }
    }
    // Node.update_links (node.py:146-162, link.py:19-25)
    double* outflow = H64(c, PNS_F64_OUTFLOW, t);
    double* inflow = H64(c, PNS_F64_INFLOW, t);
    double* cout_t = H64(c, PNS_F64_CUM_OUTFLOW, t);
    double* cin_t = H64(c, PNS_F64_CUM_INFLOW, t);
    const double* cout_p = H64(c, PNS_F64_CUM_OUTFLOW, tau);
    const double* cin_p = H64(c, PNS_F64_CUM_INFLOW, tau);
    for (int i = 0; i < m; ++i) {
        const size_t ei = (size_t)icol[i] * R + rep, eo = (size_t)ocol[i] * R + rep;
        outflow[ei] = q_out[i];
        cout_t[ei] = cout_p[ei] + q_out[i];
        inflow[eo] = q_in[i];
        cin_t[eo] = cin_p[eo] + q_in[i];
    }
}

// =================================================================================================
// BiDirectionalFd.__call__ (src/utils/functions.py:112-134) + travel time (link.py:176-177).
// Returns the float32 value stored in speed[t]; *tt is the float32 stored in travel_time[t].
__device__ float speed_and_travel_time(const LinkP& p, float k_self, float k_opp, bool have_noise, double z,
                                       float* tt) {
    const float k_eff = k_self + (float)p.bi * k_opp;
    const int fd = (p.flags >> 1) & 3;
    // the speed is a Python double in the free-flow branch (and after `0 + noise`), a float32 otherwise
    bool is_f64 = false, is_zero = false;
    double v64 = 0.0;
    float v32 = 0.0f;
    const bool free_flow = k_eff <= (float)p.kc;
    if (fd == 2) {                                    // smulders
        if (free_flow) v32 = (float)p.vf * (1.0f - k_eff / (float)p.kj);
        else {
            v32 = (float)(p.vf * p.kc) * (1.0f / k_eff - (float)(1.0 / p.kj));
            if (!(v32 > 0.0f)) is_zero = true;
        }
    } else if (free_flow) {
        is_f64 = true;
        v64 = p.vf;
    } else if (fd == 0) {                             // yperman
        v32 = (float)((p.kc * p.vf) / (p.kj - p.kc)) * ((float)p.kj / k_eff - 1.0f);
        if (!(v32 > 0.0f)) is_zero = true;
    } else {                                          // greenshields
        v32 = ((float)(-p.vf) * (k_eff - (float)p.kj)) / (float)(p.kj - p.kc);
        if (!(v32 > 0.0f)) is_zero = true;
    }
    if (have_noise) {
        if (is_zero) { is_f64 = true; is_zero = false; v64 = 0.0 + z; }
        else if (is_f64) v64 = v64 + z;
        else v32 = v32 + (float)z;
    }
    if (is_zero) { *tt = (float)(p.length / 0.05); return 0.0f; }
    if (is_f64) {
        if (!(v64 > 0.0)) { *tt = (float)(p.length / 0.05); return 0.0f; }
        *tt = (float)(p.length / v64);
        return (float)v64;
    }
    if (!(v32 > 0.0f)) { *tt = (float)(p.length / 0.05); return 0.0f; }
    *tt = (float)p.length / v32;
    return v32;
}

// Link.update_link_density_flow + update_speeds (link.py:133-188; Separator :430-452)
__global__ void __launch_bounds__(kBlock) k_link_update(const __grid_constant__ Ctx c) {
    const int R = c.n.replicas;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_pairs = (size_t)(c.n.n_links / 2);
    if (gid >= n_pairs * R) return;
    const int pair = (int)(gid / R);
    const int rep = (int)(gid % R);
    const int t = c.t, tau = c.t - 1;
    const int l0 = 2 * pair;
    const size_t e[2] = {(size_t)l0 * R + rep, (size_t)(l0 + 1) * R + rep};
    const LinkP p[2] = {load_link(c, l0), load_link(c, l0 + 1)};
    const double* inflow = H64(c, PNS_F64_INFLOW, t);
    const double* outflow = H64(c, PNS_F64_OUTFLOW, t);
    const float* num_prev = H32(c, PNS_F32_NUM_PED, tau);
    float num[2], dens[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const double delta = inflow[e[a]] - outflow[e[a]];
        num[a] = (float)((double)num_prev[e[a]] + delta);
        bool a64;
        const double area = link_area(c, p[a], e[a], &a64);
        dens[a] = div_by_area(num[a], area, a64);
    }
    float* num_t = H32(c, PNS_F32_NUM_PED, t);
    float* dens_t = H32(c, PNS_F32_DENSITY, t);
    float* speed_t = H32(c, PNS_F32_SPEED, t);
    float* tt_t = H32(c, PNS_F32_TRAVEL_TIME, t);
    float* flow_t = H32(c, PNS_F32_LINK_FLOW, t);
    float* avg_t = H32(c, PNS_F32_AVG_TRAVEL_TIME, t);
    double* bgw_t = H64(c, PNS_F64_BACK_GATE, t);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const bool sep = is_sep(p[a]);
        const bool noisy = p[a].sigma > 0.0;
        double z = 0.0;
        if (noisy) {
            if (c.mode == PNS_RNG_TABLE) z = c.draw_n[e[a]];
            else {
                pns::DrawKey key;
                key.t = (uint32_t)t; key.link = (uint32_t)(l0 + a); key.replica = (uint32_t)rep;
                key.k0 = (uint32_t)c.io.seed; key.k1 = (uint32_t)(c.io.seed >> 32);
                z = p[a].sigma * pns::normal_philox(key, 4u);
            }
        }
        float tt;
        const float v = speed_and_travel_time(p[a], dens[a], sep ? 0.0f : dens[1 - a], noisy, z, &tt);
        num_t[e[a]] = num[a];
        dens_t[e[a]] = dens[a];
        speed_t[e[a]] = v;
        tt_t[e[a]] = tt;
        flow_t[e[a]] = v * dens[a];
        float rs = c.s.runsum[e[a]] + tt;                                   // link.py:183-186
        if (t >= c.n.window) {
            rs = rs - H32(c, PNS_F32_TRAVEL_TIME, t - c.n.window)[e[a]];
            avg_t[e[a]] = rs / (float)c.n.window;
        }
        c.s.runsum[e[a]] = rs;
        if (sep) {
            const double w = c.s.widths[2 * row32(c) + e[a]];
            bgw_t[e[a]] = w;
            H64(c, PNS_F64_SEP_WIDTH, t)[e[a]] = w;
        } else {
            bgw_t[e[a]] = c.s.widths[row32(c) + e[a]];
        }
    }
}

// =================================================================================================
// Initial state (link.py:12-17, 56, 82-97, 425)
__global__ void k_state_init(const __grid_constant__ Ctx c) {
    const int R = c.n.replicas;
    const size_t n64 = row64(c), n32 = row32(c);
    const int T = c.n.sim_steps + 1;
    const size_t total = (size_t)T * n64;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / n64);
        const size_t e = i % n64;
        const bool physical = e < n32;
        const int l = physical ? (int)(e / R) : 0;
        for (int f = 0; f < c.s.n_f64; ++f) {
            double v = 0.0;
            if (f == PNS_F64_SENDING || f == PNS_F64_RECEIVING) v = -1.0;
            else if (f == PNS_F64_BACK_GATE && physical) v = c.n.lk_width[l];
            else if (f == PNS_F64_SEP_WIDTH && physical && (c.n.lk_flags[l] & 1)) v = c.n.lk_width[l] / 2;   // link.py:425
            H64(c, f, t)[e] = v;
        }
        if (physical) {
            const float tt0 = c.n.lk_tt0[l];
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

__global__ void k_rng_selftest(int kind, int n, const int32_t* n_trials, const double* p, uint64_t seed, int t,
                               int site, int32_t* out_i, double* out_d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pns::DrawKey key;
    key.t = (uint32_t)t; key.link = (uint32_t)i; key.replica = 0;
    key.k0 = (uint32_t)seed; key.k1 = (uint32_t)(seed >> 32);
    if (kind == 0) out_i[i] = pns::binomial_philox(key, (uint32_t)site, n_trials[i], p[i]);
    else if (kind == 1) out_d[i] = pns::normal_philox(key, (uint32_t)site);
    else out_d[i] = (double)pns::det_pow08((float)p[i]);
}

// ---- host side ----------------------------------------------------------------------------------
Ctx make_ctx(const pns_net* net, const pns_state* st, const pns_step_io* io, int t, int mode, int step_no) {
    Ctx c;
    c.n = *net;
    c.s = *st;
    if (io) c.io = *io; else memset(&c.io, 0, sizeof c.io);
    c.t = t;
    c.mode = mode;
    const size_t n32 = (size_t)net->n_links * net->replicas;
    const int64_t row = io ? io->draw_row_stride * step_no : 0;
    c.draw_b = (io && io->draw_b) ? io->draw_b + (size_t)row * 3 * n32 : nullptr;
    c.draw_n = (io && io->draw_n) ? io->draw_n + (size_t)row * n32 : nullptr;
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

}  // namespace

extern "C" {

int pns_abi_version(void) { return PNS_ABI_VERSION; }
const char* pns_last_error(void) { return g_err; }

int pns_state_init(const pns_net* net, const pns_state* st, void* stream) {
    if (!net || !st) return fail("null net/state");
    if (cudaMemsetAsync(st->err, 0, sizeof(int32_t) * net->replicas, (cudaStream_t)stream) != cudaSuccess)
        return fail("memset err", cudaGetLastError());
    Ctx c = make_ctx(net, st, nullptr, 1, PNS_RNG_TABLE, 0);
    PNS_LAUNCH(k_state_init, 148 * 8, 256, (cudaStream_t)stream, c);
    return launched("k_state_init");
}

int pns_link_flows(const pns_net* net, const pns_state* st, const pns_step_io* io, int t, int rng_mode,
                   void* stream) {
    if (check_common(net, st, t)) return 1;
    if (rng_mode == PNS_RNG_TABLE && !io->draw_b) return fail("TABLE mode needs draw_b");
    if (rng_mode == PNS_RNG_REQUEST && !(io->req_kind && io->req_n1 && io->req_rf && io->req_sval && io->req_n3))
        return fail("REQUEST mode needs the req_* buffers");
    const size_t n = (size_t)(net->n_links / 2) * net->replicas;
    if (n == 0) return 0;
    const Ctx c = make_ctx(net, st, io, t, rng_mode, 0);
    PNS_LAUNCH(k_link_flows, blocks_for(n), kBlock, (cudaStream_t)stream, c);
    return launched("k_link_flows");
}

int pns_route_probs(const pns_net* net, const pns_state* st, const pns_step_io* io, int t, void* stream) {
    if (check_common(net, st, t)) return 1;
    const size_t n = (size_t)net->n_groups * net->replicas;
    if (n == 0) return 0;
    const Ctx c = make_ctx(net, st, io, t, PNS_RNG_TABLE, 0);
    PNS_LAUNCH(k_route_probs, blocks_for(n), kBlock, (cudaStream_t)stream, c);
    return launched("k_route_probs");
}

int pns_node_flows(const pns_net* net, const pns_state* st, const pns_step_io* io, int t, void* stream) {
    if (check_common(net, st, t)) return 1;
    if (net->n_demand_rows > 0 && !io->demand) return fail("demand table missing");
    if (net->n_routed > 0 && !io->od_w) return fail("od weight table missing");
    const size_t n = (size_t)net->n_nodes * net->replicas;
    if (n == 0) return 0;
    const Ctx c = make_ctx(net, st, io, t, PNS_RNG_TABLE, 0);
    PNS_LAUNCH(k_node_flows, blocks_for(n), kBlock, (cudaStream_t)stream, c);
    return launched("k_node_flows");
}

int pns_link_update(const pns_net* net, const pns_state* st, const pns_step_io* io, int t, int rng_mode,
                    void* stream) {
    if (check_common(net, st, t)) return 1;
    const size_t n = (size_t)(net->n_links / 2) * net->replicas;
    if (n == 0) return 0;
    const Ctx c = make_ctx(net, st, io, t, rng_mode, 0);
    PNS_LAUNCH(k_link_update, blocks_for(n), kBlock, (cudaStream_t)stream, c);
    return launched("k_link_update");
}

int pns_step(const pns_net* net, const pns_state* st, const pns_step_io* io, int t0, int n_steps, int rng_mode,
             void* stream) {
    if (rng_mode != PNS_RNG_TABLE && rng_mode != PNS_RNG_PHILOX) return fail("pns_step: TABLE or PHILOX mode only");
    if (check_common(net, st, t0) || check_common(net, st, t0 + n_steps - 1)) return 1;
    if (rng_mode == PNS_RNG_TABLE && !io->draw_b) return fail("TABLE mode needs draw_b");
    if (net->n_demand_rows > 0 && !io->demand) return fail("demand table missing");
    if (net->n_routed > 0 && !io->od_w) return fail("od weight table missing");
    const cudaStream_t s = (cudaStream_t)stream;
    const size_t n_pair = (size_t)(net->n_links / 2) * net->replicas;
    const size_t n_grp = (size_t)net->n_groups * net->replicas;
    const size_t n_node = (size_t)net->n_nodes * net->replicas;
    for (int k = 0; k < n_steps; ++k) {
        const Ctx c = make_ctx(net, st, io, t0 + k, rng_mode, k);
        if (n_pair) PNS_LAUNCH(k_link_flows, blocks_for(n_pair), kBlock, s, c);
        if (n_grp) PNS_LAUNCH(k_route_probs, blocks_for(n_grp), kBlock, s, c);
        if (n_node) PNS_LAUNCH(k_node_flows, blocks_for(n_node), kBlock, s, c);
        if (n_pair) PNS_LAUNCH(k_link_update, blocks_for(n_pair), kBlock, s, c);
    }
    return launched("pns_step");
}

int pns_step_profiled(const pns_net* net, const pns_state* st, const pns_step_io* io, int t0, int n_steps,
                      int rng_mode, void* stream, double* ms, int64_t* launches) {
#ifdef PNS_HOST_EMULATION
    (void)ms; (void)launches;
    return pns_step(net, st, io, t0, n_steps, rng_mode, stream);
#else
    if (rng_mode != PNS_RNG_TABLE && rng_mode != PNS_RNG_PHILOX) return fail("pns_step_profiled: TABLE or PHILOX mode only");
    if (check_common(net, st, t0) || check_common(net, st, t0 + n_steps - 1)) return 1;
    const cudaStream_t s = (cudaStream_t)stream;
    const size_t n_pair = (size_t)(net->n_links / 2) * net->replicas;
    const size_t n_grp = (size_t)net->n_groups * net->replicas;
    const size_t n_node = (size_t)net->n_nodes * net->replicas;
    const int per_step = 5;
    const int n_ev = per_step * n_steps;
    cudaEvent_t* ev = (cudaEvent_t*)malloc(sizeof(cudaEvent_t) * n_ev);
    for (int i = 0; i < n_ev; ++i) cudaEventCreate(&ev[i]);
    for (int k = 0; k < n_steps; ++k) {
        const Ctx c = make_ctx(net, st, io, t0 + k, rng_mode, k);
        cudaEvent_t* e = ev + per_step * k;
        cudaEventRecord(e[0], s);
        if (n_pair) PNS_LAUNCH(k_link_flows, blocks_for(n_pair), kBlock, s, c);
        cudaEventRecord(e[1], s);
        if (n_grp) PNS_LAUNCH(k_route_probs, blocks_for(n_grp), kBlock, s, c);
        cudaEventRecord(e[2], s);
        if (n_node) PNS_LAUNCH(k_node_flows, blocks_for(n_node), kBlock, s, c);
        cudaEventRecord(e[3], s);
        if (n_pair) PNS_LAUNCH(k_link_update, blocks_for(n_pair), kBlock, s, c);
        cudaEventRecord(e[4], s);
    }
    cudaError_t err = cudaStreamSynchronize(s);
    const size_t counts[4] = {n_pair, n_grp, n_node, n_pair};
    for (int k = 0; k < n_steps && err == cudaSuccess; ++k)
        for (int j = 0; j < 4; ++j) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ev[per_step * k + j], ev[per_step * k + j + 1]);
            if (counts[j]) { ms[j] += t; launches[j] += 1; }
        }
    for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i]);
    free(ev);
    if (err != cudaSuccess) return fail("pns_step_profiled", err);
    return launched("pns_step_profiled");
#endif
}

int pns_rng_selftest(int kind, int n, const int32_t* n_trials, const double* p, uint64_t seed, int t, int site,
                     int32_t* out_i, double* out_d, void* stream) {
    if (n <= 0) return 0;
    PNS_LAUNCH(k_rng_selftest, (n + 127) / 128, 128, (cudaStream_t)stream, kind, n, n_trials, p, seed, t, site, out_i,
                                                                       out_d);
    return launched("k_rng_selftest");
}

}  // extern "C"
