// The per-node linear program of RegularNode.solve(type='optimal') (reference src/LTM/node.py:249-271, constraint
// matrices node.py:73-104 `get_matrix_A` and :110-137 `update_matrix_A_eq`), solved by a dense-tableau primal simplex
// that one warp runs in shared memory.
//
// For a node with m link slots and E = m(m-1) turns e = (i, j != i) the reference hands scipy.optimize.linprog
//
//     minimise    - sum_e x_e  +  w sum_e (p_e + n_e)                     (w = Node.w = 0.01, node.py:14)
//     subject to  sum_j x_ij <= s_i                 (sending flow of incoming slot i,   rows 0 .. m-1)
//                 sum_i x_ij <= r_j                 (receiving flow of outgoing slot j, rows m .. 2m-1)
//                 phi_e X_i - x_e + p_e - n_e = 0   (X_i = sum_j x_ij; rows 2m .. 2m+E-1)
//                 x, p, n >= 0                      (linprog's default bounds)
//
// i.e. it maximises the total flow through the node with an L1 penalty on deviations from the turning fractions.
// Slack variables of the inequality rows and the p_e of the equality rows (right-hand side 0) form a feasible
// starting basis, so there is no phase one.  Pivoting: most negative reduced cost (lowest index on ties), minimum
// ratio with the lowest basic variable on ties; after kLpBlandAfter pivots Bland's rule (termination guaranteed in
// the presence of the problem's many degenerate rows).
//
// The routine is written once for LANES cooperating threads: LANES = 32 on the device (a warp; columns of the
// tableau are spread over the lanes, so a pivot is rows x cols / 32 multiply-subtracts per lane with conflict-free
// shared-memory rows), LANES = 1 in the sequential host build used by the CPU-side tests.  Arithmetic is identical
// in both (no contraction: the file is compiled with -fmad=false), so the two builds return the same bits.
#pragma once
#include <stdint.h>

namespace pns {

constexpr int kLpMaxSlots = 8;             // = PNS_MAX_DEGREE
constexpr int kLpColsPerPass = 3;          // tableau columns a lane updates together in a pivot (m <= 5: all of them)
constexpr int kLpBlandAfter = 400;
constexpr int kLpMaxPivots = 20000;
constexpr double kLpCostTol = 1e-9;        // a reduced cost below -tol (x the column's scale, if that is below 1) enters
constexpr double kLpPivotTol = 1e-9;       // the ratio test pivots on entries above tol x the largest of the column
constexpr double kLpSmallEntry = 1e-9;     // HiGHS (the solver behind linprog) drops matrix entries below its
                                           // small_matrix_value = 1e-9: the logit produces such fractions
constexpr double kLpZero = 1e-13;          // tableau entries below this are rounding residue of exact zeros

enum { LP_NOT_UNIQUE = 1 << 28, LP_UNBOUNDED = 1 << 29, LP_PIVOT_LIMIT = 1 << 30 };   // info bits above the pivot count

struct LpShape { int m, E, rows, cols, ld; };   // rows: constraints (objective row is number `rows`); cols incl. rhs
__host__ __device__ inline LpShape lp_shape(int m) {
    LpShape d;
    d.m = m; d.E = m * (m - 1); d.rows = 2 * m + d.E; d.cols = 3 * d.E + 2 * m + 1;
    d.ld = d.cols | 1;                           // odd row stride: rows of one column fall into different banks
    return d;
}
// doubles of scratch one solve needs: tableau (rows+1) x ld, pivot column (rows+1), x (E); ints: basis (rows)
__host__ __device__ inline size_t lp_scratch_doubles(int m) {
    const LpShape d = lp_shape(m);
    return (size_t)(d.rows + 1) * d.ld + (d.rows + 1) + d.E;
}
__host__ __device__ inline size_t lp_scratch_bytes(int m) {
    return (lp_scratch_doubles(m) + (lp_shape(m).rows + 1) / 2 + 1) * sizeof(double);
}

#ifdef PNS_HOST_EMULATION
#define PNS_LP_SYNC() do { } while (0)
#else
#define PNS_LP_SYNC() __syncwarp()
#endif

// max over the cooperating lanes
template <int LANES>
__device__ __forceinline__ double lp_max(double v) {
#ifndef PNS_HOST_EMULATION
    if (LANES > 1) {
#pragma unroll
        for (int d = LANES / 2; d > 0; d >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, v, d);
            v = o > v ? o : v;
        }
    }
#endif
    return v;
}

// argmin over the cooperating lanes of (key, tie); lanes without a candidate pass idx < 0
template <int LANES>
__device__ __forceinline__ void lp_argmin(double& key, int& tie, int& idx) {
#ifndef PNS_HOST_EMULATION
    if (LANES > 1) {
#pragma unroll
        for (int d = LANES / 2; d > 0; d >>= 1) {
            const double k2 = __shfl_xor_sync(0xffffffffu, key, d);
            const int t2 = __shfl_xor_sync(0xffffffffu, tie, d);
            const int i2 = __shfl_xor_sync(0xffffffffu, idx, d);
            const bool take = i2 >= 0 && (idx < 0 || k2 < key || (k2 == key && t2 < tie));
            if (take) { key = k2; tie = t2; idx = i2; }
        }
    }
#endif
}

// Solves the program for one node.  s, r: [m] (>= 0), phi: [E] with element stride phi_stride.  On return x[0..E)
// holds the turn flows (scratch memory, all lanes may read it after the trailing sync), *objective the optimum;
// the return value is the pivot count plus LP_* bits.  `scratch` = lp_scratch_bytes(m) bytes private to the caller.
template <int LANES>
__device__ inline int lp_node_solve(int m, const double* s, const double* r, const double* phi, size_t phi_stride,
                                    double w, int lane, void* scratch, double** x_out, double* objective) {
    const LpShape d = lp_shape(m);
    const int E = d.E, rows = d.rows, cols = d.cols, ld = d.ld, rhs = d.cols - 1, obj = d.rows;
    double* T = reinterpret_cast<double*>(scratch);
    double* pcol = T + (size_t)(rows + 1) * ld;
    double* x = pcol + rows + 1;
    int* basis = reinterpret_cast<int*>(x + E);

    for (int k = lane; k < (rows + 1) * ld; k += LANES) T[k] = 0.0;
    PNS_LP_SYNC();
    for (int e = lane; e < E; e += LANES) {
        const int i = e / (m - 1), jj = e - i * (m - 1), j = jj < i ? jj : jj + 1;
        double f = phi[(size_t)e * phi_stride];
        if (f < kLpSmallEntry && f > -kLpSmallEntry) f = 0.0;         // as the reference's solver does (see kLpSmallEntry)
        T[(size_t)i * ld + e] = 1.0;                                   // node.py:86-89
        T[(size_t)(m + j) * ld + e] = 1.0;                             // node.py:92-97
        double* q = T + (size_t)(2 * m + e) * ld;                      // node.py:127-135
        for (int k = 0; k < m - 1; ++k) q[i * (m - 1) + k] = f;
        q[e] = f - 1.0;
        q[E + e] = 1.0;
        q[2 * E + e] = -1.0;
        basis[2 * m + e] = E + e;
    }
    for (int k = lane; k < 2 * m; k += LANES) {
        T[(size_t)k * ld + 3 * E + k] = 1.0;
        T[(size_t)k * ld + rhs] = k < m ? s[k] : r[k - m];
        basis[k] = 3 * E + k;
    }
    PNS_LP_SYNC();
    // reduced costs with the p_e (cost w) basic in their rows: c_j - w * (column sum over the equality rows)
    for (int j = lane; j < 3 * E; j += LANES) {
        double colsum = 0.0;
        for (int e = 0; e < E; ++e) colsum += T[(size_t)(2 * m + e) * ld + j];
        T[(size_t)obj * ld + j] = (j < E ? -1.0 : w) - w * colsum;
    }
    PNS_LP_SYNC();

    int pivots = 0, flags = 0;
    for (;;) {
        const bool bland = pivots >= kLpBlandAfter;
        double zk = 0.0; int zt = 0, zj = -1;
        for (int j = lane; j < rhs; j += LANES) {
            const double z = T[(size_t)obj * ld + j];
            if (z < -kLpCostTol) {
                const double key = bland ? 0.0 : z;
                if (zj < 0 || key < zk) { zk = key; zt = j; zj = j; }
            }
        }
        lp_argmin<LANES>(zk, zt, zj);
        if (zj < 0) {
            // Nothing enters on the absolute test.  Turning fractions from the logit reach down to 1e-9 and below, and
            // a variable whose column has only entries of that size moves the solution by O(1) per 1e9 of its units:
            // its reduced cost is tiny although the step it allows is huge.  Such columns are judged relative to
            // their own scale (largest entry, between kLpZero and 1).
            for (int j = lane; j < rhs; j += LANES) {
                const double z = T[(size_t)obj * ld + j];
                if (z < 0.0) {
                    double cmax = 0.0;
                    for (int i = 0; i < rows; ++i) { const double a = fabs(T[(size_t)i * ld + j]); cmax = a > cmax ? a : cmax; }
                    if (cmax > kLpZero && cmax < 1.0 && z < -kLpCostTol * cmax) {
                        const double key = bland ? 0.0 : z / cmax;
                        if (zj < 0 || key < zk) { zk = key; zt = j; zj = j; }
                    }
                }
            }
            lp_argmin<LANES>(zk, zt, zj);
            if (zj < 0) break;                                         // optimal
        }
        double cmax = 0.0;
        for (int i = lane; i < rows; i += LANES) { const double a = T[(size_t)i * ld + zj]; cmax = a > cmax ? a : cmax; }
        cmax = lp_max<LANES>(cmax);
        const double amin = cmax * kLpPivotTol > kLpZero ? cmax * kLpPivotTol : kLpZero;
        // minimum ratio; among the rows that attain it (the program is degenerate by construction: many share
        // ratio 0) the largest pivot element, for stability -- a 1e-9 pivot next to entries of 1 turned the
        // tableau into 1e8s -- and, once Bland's rule is on, the lowest basic variable
        double qmin = 1e300;
        for (int i = lane; i < rows; i += LANES) {
            const double a = T[(size_t)i * ld + zj];
            if (a > amin) {
                const double b = T[(size_t)i * ld + rhs];
                const double q = (b > 0.0 ? b : 0.0) / a;
                qmin = q < qmin ? q : qmin;
            }
        }
        qmin = -lp_max<LANES>(-qmin);
        double rk = 0.0; int rt = 0, ri = -1;
        if (qmin < 1e300) {
            const double qtie = qmin + 1e-13 * qmin + 1e-13;
            for (int i = lane; i < rows; i += LANES) {
                const double a = T[(size_t)i * ld + zj];
                if (a > amin) {
                    const double b = T[(size_t)i * ld + rhs];
                    const double q = (b > 0.0 ? b : 0.0) / a;
                    if (q <= qtie) {
                        const double key = bland ? 0.0 : -a;
                        const int bi = basis[i];
                        if (ri < 0 || key < rk || (key == rk && bi < rt)) { rk = key; rt = bi; ri = i; }
                    }
                }
            }
        }
        lp_argmin<LANES>(rk, rt, ri);
        if (ri < 0) { flags |= LP_UNBOUNDED; break; }
        if (++pivots > kLpMaxPivots) { flags |= LP_PIVOT_LIMIT; break; }
        for (int i = lane; i <= rows; i += LANES) pcol[i] = T[(size_t)i * ld + zj];
        PNS_LP_SYNC();
        const double piv = pcol[ri];
        // row operations, lanes own columns lane, lane + LANES, ...: the scaled pivot row of a lane's columns stays
        // in registers, rows go by in the outer loop (one broadcast read of the pivot column per row; rows with a
        // zero there are skipped by the whole warp), the lane's columns in the unrolled inner one
        for (int j0 = lane; j0 < cols; j0 += LANES * kLpColsPerPass) {
            double pr[kLpColsPerPass];
#pragma unroll
            for (int k = 0; k < kLpColsPerPass; ++k) {
                const int j = j0 + k * LANES;
                pr[k] = j < cols ? T[(size_t)ri * ld + j] / piv : 0.0;
            }
            for (int i = 0; i <= rows; ++i) {
                const double f = pcol[i];
                if (f == 0.0 || i == ri) continue;
                double* __restrict__ row = T + (size_t)i * ld;
#pragma unroll
                for (int k = 0; k < kLpColsPerPass; ++k) {
                    const int j = j0 + k * LANES;
                    if (j < cols && pr[k] != 0.0) row[j] -= f * pr[k];
                }
            }
#pragma unroll
            for (int k = 0; k < kLpColsPerPass; ++k) {
                const int j = j0 + k * LANES;
                if (j < cols) T[(size_t)ri * ld + j] = pr[k];
            }
        }
        if (lane == 0) basis[ri] = zj;
        PNS_LP_SYNC();
    }
    // a non-basic column with zero reduced cost: the optimum may be a face, not a vertex
    {
        int zero_cost = 0;
        for (int j = lane; j < rhs; j += LANES) {
            const double z = T[(size_t)obj * ld + j];
            if (z <= kLpCostTol && z >= -kLpCostTol) ++zero_cost;
        }
#ifndef PNS_HOST_EMULATION
        if (LANES > 1)
            for (int dd = LANES / 2; dd > 0; dd >>= 1) zero_cost += __shfl_xor_sync(0xffffffffu, zero_cost, dd);
#endif
        if (zero_cost > rows) flags |= LP_NOT_UNIQUE;                 // the `rows` basic columns are zero by construction
    }
    for (int e = lane; e < E; e += LANES) x[e] = 0.0;
    PNS_LP_SYNC();
    for (int i = lane; i < rows; i += LANES) {
        const int b = basis[i];
        if (b < E) { const double v = T[(size_t)i * ld + rhs]; x[b] = v > 0.0 ? v : 0.0; }
    }
    PNS_LP_SYNC();
    // Fractions of 1e-7 and below make the tableau ill-conditioned (entries of 1e7 next to 1e-7); what is left of
    // that in x is a 1e-7 overshoot of a sending or receiving flow in about one program in a hundred of that
    // kind.  Scale the offending row / column back onto its bound (programs without such fractions are untouched:
    // their sums meet the bounds exactly).
    for (int i = lane; i < m; i += LANES) {
        double sum = 0.0;
        for (int k = 0; k < m - 1; ++k) sum += x[i * (m - 1) + k];
        if (sum > s[i]) { const double f = s[i] / sum; for (int k = 0; k < m - 1; ++k) x[i * (m - 1) + k] *= f; }
    }
    PNS_LP_SYNC();
    for (int j = lane; j < m; j += LANES) {
        double sum = 0.0;
        for (int i = 0; i < m; ++i) if (i != j) sum += x[i * (m - 1) + (j < i ? j : j - 1)];
        if (sum > r[j]) { const double f = r[j] / sum; for (int i = 0; i < m; ++i) if (i != j) x[i * (m - 1) + (j < i ? j : j - 1)] *= f; }
    }
    PNS_LP_SYNC();
    *x_out = x;
    *objective = -T[(size_t)obj * ld + rhs];
    return pivots | flags;
}

}  // namespace pns
