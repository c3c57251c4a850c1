// Device code of the batched LinMPC step (sm_100a).  See DESIGN.md for the algorithm and
// include/bmpc.h for the reference functions each stage replaces.
//
// One TEAM of threads (8/16/32 lanes of a warp, or a whole CTA of 64..256 threads) owns one
// controller instance at a time.  Per instance:
//   stage 0  TMA bulk loads (cp.async.bulk + mbarrier) of the instance's dense constraint
//            rows Pd, packed Hessian Hv and its Cholesky factor Lv from HBM into shared memory
//   stage 1  initpred! + linconstraint!: F, q, r, fx, row right-hand sides h   (execute.jl:247-277,
//            transcription.jl:811-848)
//   stage 2  unconstrained exit: x = -Hv^-1 q with the cached factor; feasible -> done
//            (ExplicitMPC, explicitmpc.jl:209)
//   stage 3  Mehrotra predictor-corrector interior point (warm-started from the previous period); per iteration
//            Phi = Hv + P' D P -- dense rows on the FP64 tensor pipe (DMMA 8x8x4 tiles, 16x16 or 32x32 blocks per
//            warp) for teams of a warp or more, a scalar pair loop for sub-warp teams; 1-/2-variable rows by a gather;
//            packed Cholesky in shared memory (CTA teams: blocked, DMMA trailing updates), two blocked solve pairs
//   stage 4  getinput!: Z = D x, u, lastu0, J, status               (execute.jl:536-546)
// Coordinates: "input levels" x = [v; eps], v_l = sum_{i<=l} DU_i  (DESIGN.md section 3).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bmpc {

constexpr int ST_OPTIMAL = 0, ST_ITERATION_LIMIT = 1, ST_INFEASIBLE = 2;  // = BMPC_STATUS_* in bmpc.h

// Study builds (-DBMPC_PHASE_CLK, tools/studies/phase_clk_general.py): thread 0 of a team accumulates the cycles between marks
#ifdef BMPC_PHASE_CLK
#define GCLK_DECL long long gclk_t = clock64()
#define GCLK(P_, tid_, i) do { if ((tid_) == 0 && (P_).gclk) { const long long t_ = clock64(); atomicAdd((unsigned long long*)&(P_).gclk[i], (unsigned long long)(t_ - gclk_t)); gclk_t = t_; } } while (0)
#define GCLK_COUNT(P_, tid_, i, v) do { if ((tid_) == 0 && (P_).gclk) atomicAdd((unsigned long long*)&(P_).gclk[i], (unsigned long long)(v)); } while (0)
#else
#define GCLK_DECL
#define GCLK(P_, tid_, i)
#define GCLK_COUNT(P_, tid_, i, v)
#endif

struct RowTables {
    int nS, nDr, nDb, m;     // sparse rows, dense rows, dense base rows, total rows (incl. eps>=0)
    const int* s_i1;         // [nS] variable with coefficient +sigma
    const int* s_i2;         // [nS] variable with coefficient -sigma, or -1
    const int* s_ch;         // [nS] input channel whose lastu0 shifts the bound, or -1
    const double* row_sig;   // [m]  +1 (max side) / -1 (min side); eps row: 0
    const double* row_c;     // [m]  softness c (coefficient of -eps); eps row: 1
    const int* var_ptr;      // [nz+1] CSR variable -> sparse rows
    const int* var_row;      //        row index
    const int* var_sgn;      //        +1 if the variable is i1 of the row, -1 if i2
    const int* db_rmax;      // [nDb] dense base row -> its max-side row index or -1
    const int* db_rmin;      // [nDb] ... min-side row index or -1
    const int* dr_base;      // [nDr] dense row -> base row k
    const int* dr_src;       // [nDr] dense row -> source: t < nY (F[t]), nY + i (fx[i]) or nY + nx + r (custom row r: Fw[r])
    const short* pair_i;     // [nz(nz+1)/2] row index of packed pair p
    const short* pair_j;     //              col index
};

struct SmemLayout {  // offsets in doubles inside a team's slice
    int Pd, Hv, Phi, x, xb, q, rd, rhs, dx, invd, F, tY, fx, yb, ybd, wd, s, lam, h, rp, t, ds, dl,
        xhat, lastu, dd, Dh, red, bar, ev, Fw, cu, tU, total;
};

struct StepParams {
    int N, nu, ny, nd, nx, Hp, Hc, nz, n, neps, nY, nU;
    int max_iter;
    double tol, tol_mu;
    RowTables rt;
    SmemLayout sm;
    int pd_in_smem, pd_is_ev, has_terminal, M_dense, has_L, L_dense;
    int hv_in_smem;          // 0: the packed Hessian stays in HBM/L2 (large n), read where it is used
    long sPd, sEv, sH, sK, sV, sB, sG, sJ, skx, svx, sbx, sgx, sjx, sM, sL, suop, syop;
    const double *Pd, *Ev, *Hv, *Lv, *Hee, *K, *V, *B, *G, *J, *kx, *vx, *bx, *gx, *jx, *Mw, *Lw,
        *uop, *yop, *sbase, *dbound;
    const int* lv_ok;
    const int* blk_of_t;  // [Hp] move block of time step t
    const int* blk_start; // [Hc+1] first time step of block l (j_l)
    // per-step io (device pointers)
    const double *xhat0, *ry, *Rhat_y, *Rhat_u, *d0, *Dhat0;
    double *lastu0, *Z, *u, *J_out, *F_out, *qt_out, *r_out, *lastu_prev;
    int *status, *iters;
    unsigned int* counters;  // [2] work counter, finished-CTA counter
    double* lam_ws;          // [N x ws_stride] multipliers of the previous period (IPM warm start), small kernel
    int* ws_flag;            // [N] 1 if lam_ws holds a converged IPM solution of the previous period
    int ws_stride, use_ws;
    int nHp2, nPd2;          // padded (even) sizes in doubles of packed Hv and of Pd: TMA needs 16-byte multiples
    // fused all-gather of Z̃ (multi-GPU): every rank's gather buffer [world x N x n], peer-mapped over NVLink; the
    // epilogue stores this instance's Z̃ straight into slot (rank, inst) of EVERY peer's buffer (world = 0: off)
    double* zg[8];
    int zg_world, zg_rank;
    long zg_base;                    // offset (doubles) of this launch's first row inside the destination buffer(s)
    unsigned long long* zg_flag[8];  // pull protocol (bmpc_set_gather_pull): peer p's flag array [2 world] = data epochs, then ack
                                     // epochs; nullptr: push protocol (bmpc_set_gather, the caller supplies the barrier)
    unsigned long long zg_epoch;     // period number this launch publishes (st.release.sys by the last CTA out)
    int zg_pull;                     // 1: Z̃ goes to THIS rank's slot buffer only (zg[zg_rank]); peers pull it over NVLink
    long long zg_need_ack;           // > 0: slot reuse needs every reader's ack epoch >= this value before the launch may store
    int* zg_timeout;                 // set to 1 if the ack wait gave up
    // fused observer (SteadyKalmanFilter, kalman.jl:284-309): correct before the step, predict after it.
    // est_on: x̂0 is STATE OF THE HANDLE (xstate, in/out); the corrected estimate used by the step goes to xcorr.
    int est_on, nym;
    long s_eA, s_eBu, s_eBd, s_eCm, s_eDdm, s_eK, s_efx;
    const double *eA, *eBu, *eBd, *eCm, *eDdm, *eK, *efx, *y0m;
    double *xstate, *xcorr;
    // custom linear constraints  Wmin <= Wy Ŷe + Wu Ue + Wd D̂e + Wr R̂e <= Wmax  (construct.jl:666-695, 1138-1160;
    // linconstraint_custom!, execute.jl:337-366): nw rows per step of the extended horizon (Hp + 1 blocks).
    // Wc per model: [Wy (nw x ny) | Wu (nw x nu) | Wd (nw x nd) | Wr (nw x ny) | Ĉ (ny x nx) | D̂d (ny x nd) | dop (nd)],
    // each piece column-major.
    int nw;
    long sW;
    const double* Wc;
#ifdef BMPC_PHASE_CLK
    long long* gclk;      // study builds: [32] accumulated cycles per phase of the general kernel's IPM
#endif
    const double* Ys;     // [N x nY] stochastic output predictions Ŷs added to F (InternalModel, predictstoch! execute.jl:321-327) or nullptr
    double* kkt_out;      // [N x 3] relative KKT residuals of the returned iterate: primal, dual, complementarity (or nullptr)
};

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Fused all-gather, pull protocol: called by ONE thread of the last CTA out, after it has observed every other CTA's
// arrival (each preceded by a fence after that CTA's stores into this rank's slot buffer).  The release store of the
// period number into entry `rank` of this rank's OWN flag array makes this launch's rows visible to a reader that acquires
// it and then loads them over NVLink.
__device__ __forceinline__ void publish_epoch(const StepParams& P) {
    if (P.zg_world <= 0 || !P.zg_flag[0]) return;
    if (P.zg_pull) {
        // pull protocol: the period number goes into THIS rank's own flag array only (entry `rank`); the readers poll
        // it over NVLink.  The launch never touches remote memory, so nothing has to drain before it can retire.
        unsigned long long* f = P.zg_flag[P.zg_rank] + P.zg_rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(P.zg_epoch) : "memory");
        return;
    }
    __threadfence_system();
    for (int pr = 0; pr < P.zg_world; ++pr) {
        unsigned long long* f = P.zg_flag[pr] + P.zg_rank;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(f), "l"(P.zg_epoch) : "memory");
    }
}

// Back-pressure of the pull protocol: before a launch overwrites slot (epoch % slots) every reader must have pulled the
// epoch that lived there (its ack epoch >= zg_need_ack).  Lanes 0..world-1 of ONE warp per CTA spin on the LOCAL ack
// entries (the readers write them remotely); normally satisfied on the first load.  Gives up after ~2 s.
__device__ __forceinline__ void gather_wait_acks(const StepParams& P, int lane) {
    if (P.zg_need_ack <= 0 || !P.zg_flag[P.zg_rank]) return;
    if (lane < P.zg_world) {
        const unsigned long long* a = P.zg_flag[P.zg_rank] + P.zg_world + lane;
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory");
            if ((long long)v >= P.zg_need_ack) break;
            if (clock64() - t0 > 4000000000LL) {
                *P.zg_timeout = 1;
                break;
            }
            __nanosleep(100);
        }
    }
    __syncwarp();
}

// Fw of the custom linear constraints (linconstraint_custom! + linconstraint_custom_outputs!, execute.jl:337-366), in
// ABSOLUTE units: block t = 0..Hp of  Wu u(k-1) + Wd d̂(k+t) + Wr r̂y(k+t) + Wy ŷ(k+t)  with ŷ(k) = Ĉ x̂0 + D̂d d0 + yop
// (evaloutput of the estimator) and ŷ(k+t) = F[t-1] + yop for t >= 1.  xh, lu, d0, Dh, F are this instance's vectors in
// shared memory; out has nw (Hp + 1) entries.  Called by tid = 0..nth-1 of the team; the caller synchronises.
__device__ __forceinline__ void custom_fw(const StepParams& P, long inst, int tid, int nth, const double* xh,
                                          const double* lu, const double* d0, const double* Dh, const double* F,
                                          double* out) {
    const int nw = P.nw, ny = P.ny, nu = P.nu, nd = P.nd, nx = P.nx, Hp = P.Hp;
    const double* W = P.Wc + inst * P.sW;
    const double* Wy = W;
    const double* Wu = Wy + nw * ny;
    const double* Wd = Wu + nw * nu;
    const double* Wr = Wd + nw * nd;
    const double* Ch = Wr + nw * ny;
    const double* Dd = Ch + ny * nx;
    const double* dop = Dd + ny * nd;
    const double* guop = P.uop + inst * P.suop;
    const double* gyop = P.yop + inst * P.syop;
    for (int idx = tid; idx < nw * (Hp + 1); idx += nth) {
        const int t = idx / nw, j = idx - t * nw;
        double a = 0.0;
        for (int c = 0; c < nu; ++c) a = fma(Wu[j + nw * c], lu[c] + guop[c], a);
        for (int e = 0; e < nd; ++e) a = fma(Wd[j + nw * e], (t == 0 ? d0[e] : Dh[(t - 1) * nd + e]) + dop[e], a);
        for (int o = 0; o < ny; ++o) {
            double r;
            if (t == 0)
                r = P.ry ? P.ry[inst * ny + o] : P.Rhat_y[inst * P.nY + o];
            else
                r = P.Rhat_y ? P.Rhat_y[inst * P.nY + (t - 1) * ny + o] : P.ry[inst * ny + o];
            a = fma(Wr[j + nw * o], r, a);
            double y;
            if (t == 0) {
                y = gyop[o];
                for (int k = 0; k < nx; ++k) y = fma(Ch[o + ny * k], xh[k], y);
                for (int e = 0; e < nd; ++e) y = fma(Dd[o + ny * e], d0[e], y);
            } else {
                y = F[(t - 1) * ny + o] + gyop[o];
            }
            a = fma(Wy[j + nw * o], y, a);
        }
        out[idx] = a;
    }
}

// ------------------------------------------------------------------------------------------
// Team: 8/16/32 lanes inside a warp (independent thread scheduling + masked sync), or a CTA.
// ------------------------------------------------------------------------------------------
template <int TEAM>
struct Team {
    static constexpr bool kSub = TEAM <= 32;
    int tid;
    unsigned mask;
    double* red;  // CTA teams: 40 doubles of scratch
    __device__ __forceinline__ void sync() const {
        if (kSub)
            __syncwarp(mask);
        else
            __syncthreads();
    }
    template <class Op>
    __device__ __forceinline__ double reduce(double v, Op op, double ident) const {
        if (kSub) {
#pragma unroll
            for (int o = TEAM / 2; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(mask, v, o));
            return v;
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
            __syncthreads();
            if ((tid & 31) == 0) red[tid >> 5] = v;
            __syncthreads();
            double r = ident;
            for (int w = 0; w < TEAM / 32; ++w) r = op(r, red[w]);
            return r;
        }
    }
    __device__ __forceinline__ double sum(double v) const {
        return reduce(v, [](double a, double b) { return a + b; }, 0.0);
    }
    __device__ __forceinline__ double max(double v) const {
        return reduce(v, [](double a, double b) { return fmax(a, b); }, -1e300);
    }
    __device__ __forceinline__ double min(double v) const {
        return reduce(v, [](double a, double b) { return fmin(a, b); }, 1e300);
    }
};

__device__ __forceinline__ int pidx(int i, int j) { return (i * (i + 1)) / 2 + j; }  // i >= j

// FP64 tensor-core tile: D(8x8) += A(8x4) * B(4x8).  Lane l = 4*g + t holds A[g][t], B[t][g], D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// In-place packed (row-major lower) Cholesky, column-Crout with redundant diagonal; invd = 1/L_jj.
// Returns number of guarded (non-positive) pivots (uniform across the team).
template <int TEAM>
__device__ __forceinline__ int chol_packed(const Team<TEAM>& T, double* A, double* invd, int n) {
    int bad = 0;
#ifndef BMPC_CHOL_BLOCKED_MIN
#define BMPC_CHOL_BLOCKED_MIN 56
#endif
    if (TEAM >= 64 && n < BMPC_CHOL_BLOCKED_MIN) {
        // CTA teams, small n: right-looking, ONE barrier per column.  Column k is used unscaled for the trailing update
        // (A_ij -= A_ik A_jk / d_k) and scaled to L during the next column's phase, when nobody reads it any more.
        // Threads form a TX x TY grid over the trailing triangle: TY rows per pass, TX threads along a row.
        constexpr int TX = TEAM >= 256 ? 8 : 4, TY = TEAM / TX;
        const int tx = T.tid % TX, ty = T.tid / TX;
        double rs_prev = 0.0, dk_prev = 0.0;
        for (int k = 0; k <= n; ++k) {
            T.sync();  // trailing updates of column k-1 are complete: column k is final
            double invdk = 0.0, rs = 0.0, dk = 0.0;
            if (k < n) {
                dk = A[pidx(k, k)];
                if (!(dk > 1e-280)) {
                    dk = 1e200;
                    ++bad;
                }
                rs = rsqrt(dk);
                invdk = rs * rs;
                const int rem = n - k - 1;
                // (a variant that enumerates the trailing triangle through the packed pair tables -- every lane busy, 1/5 of
                // the instructions -- was measured SLOWER at n = 41: 68 k vs 58 k cycles per factorisation; the phase is bound by
                // the 41 barrier + rsqrt + dependent update round trips, not by instruction issue)
                for (int a = ty; a < rem; a += TY) {
                    const int i = k + 1 + a;
                    const double lik = A[pidx(i, k)] * invdk;
                    double* row = A + pidx(i, k + 1);
                    int pj = pidx(k + 1 + tx, k);  // A[k+1+b][k], advanced by TX rows per step
                    // four entries per trip, ALL loads before the first store: `row` and the column alias in A, so a plain
                    // read-modify-write loop serialises one shared-memory round trip per entry (measured: the update, not the
                    // barriers, was 80 % of the 1400 cycles per column at n = 41)
                    for (int b = tx; b <= a; b += 4 * TX) {
                        double cv[4], rv[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int bb = b + u * TX;
                            const bool ok = bb <= a;
                            cv[u] = ok ? A[pj] : 0.0;
                            rv[u] = ok ? row[bb] : 0.0;
                            pj += TX * (k + 1 + bb) + TX * (TX + 1) / 2;  // pidx(r+TX,k) - pidx(r,k) with r = k+1+bb
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int bb = b + u * TX;
                            if (bb <= a) row[bb] = fma(-lik, cv[u], rv[u]);
                        }
                    }
                }
            }
            if (k > 0) {  // scale column k-1 (nobody reads it in this phase)
                for (int i = k + T.tid; i < n; i += TEAM) A[pidx(i, k - 1)] *= rs_prev;
                if (T.tid == 0) {
                    A[pidx(k - 1, k - 1)] = dk_prev * rs_prev;
                    invd[k - 1] = rs_prev;
                }
            }
            rs_prev = rs;
            dk_prev = dk;
        }
        T.sync();
        return bad;
    }
    if constexpr (TEAM >= 64) {
        // CTA teams: blocked right-looking Cholesky with panels of 16 columns (LAPACK potrf structure):
        //   (a) warp 0 factors the 16x16 diagonal block on registers + shuffles,
        //   (b) every thread takes one row below the block and solves it against the block (16-step forward
        //       substitution on registers, block entries broadcast from shared memory),
        //   (c) the trailing matrix is updated on the FP64 tensor pipe: A22 -= L21 L21' by 8x8x4 DMMA tiles,
        //       16x16 output blocks per warp, fragments read from the packed factor in shared memory.
        // Three barriers per panel instead of two per column.
        constexpr int NBK = 16, NW = TEAM / 32;
        const int warp = T.tid >> 5, lane = T.tid & 31, fg = lane >> 2, ft = lane & 3;
        for (int c0 = 0; c0 < n; c0 += NBK) {
            const int nb = min(NBK, n - c0);
            T.sync();
            if (warp == 0) {  // (a)
                const bool valid = lane < nb;
                const int ri = c0 + (valid ? lane : 0);
                const double* src = A + pidx(ri, c0);
                double row[NBK];
#pragma unroll
                for (int j = 0; j < NBK; ++j) row[j] = (valid && j <= lane) ? src[j] : 0.0;
                double pdiag = 1.0, ldiag = 1.0, inv = 1.0;
#pragma unroll
                for (int j = 0; j < NBK; ++j)
                    if (j == lane && valid) pdiag = row[j];
#pragma unroll
                for (int k = 0; k < NBK; ++k) {
                    double dk = __shfl_sync(0xffffffffu, pdiag, k);
                    if (!(dk > 1e-280)) {
                        dk = 1e200;
                        if (k < nb) ++bad;
                    }
                    const double rs = rsqrt(dk);
                    const double lik = (lane > k && valid) ? row[k] * rs : 0.0;
                    if (lane == k) {
                        inv = rs;
                        ldiag = dk * rs;
                    }
                    row[k] = lik;
                    pdiag = fma(-lik, lik, pdiag);
#pragma unroll
                    for (int j = k + 1; j < NBK; ++j) {
                        const double ljk = __shfl_sync(0xffffffffu, lik, j);
                        row[j] = fma(-lik, ljk, row[j]);
                    }
                }
                if (valid) {
                    double* dst = A + pidx(ri, c0);
#pragma unroll
                    for (int j = 0; j < NBK; ++j)
                        if (j < lane) dst[j] = row[j];
                    dst[lane] = ldiag;
                    invd[ri] = inv;
                }
            }
            const int r0 = c0 + nb;
            if (r0 >= n) break;  // (uniform) last panel: nothing below
            T.sync();
            for (int i = r0 + T.tid; i < n; i += TEAM) {  // (b)
                double* ai = A + pidx(i, c0);
                double a[NBK];
#pragma unroll
                for (int j = 0; j < NBK; ++j) a[j] = j < nb ? ai[j] : 0.0;
#pragma unroll
                for (int kk = 0; kk < NBK; ++kk) {
                    if (kk < nb) {
                        const double* lk = A + pidx(c0 + kk, c0);
                        double sacc = a[kk];
#pragma unroll
                        for (int pp = 0; pp < kk; ++pp) sacc = fma(-a[pp], lk[pp], sacc);
                        a[kk] = sacc * invd[c0 + kk];
                    }
                }
#pragma unroll
                for (int j = 0; j < NBK; ++j)
                    if (j < nb) ai[j] = a[j];
            }
            T.sync();
            {  // (c)
                const int nrem = n - r0;
                const int nblk = (nrem + 15) >> 4;
                const int ntask = nblk * (nblk + 1) / 2;
                for (int task = warp; task < ntask; task += NW) {
                    int bi = (int)((sqrt(8.0 * task + 1.0) - 1.0) * 0.5);
                    while ((bi + 1) * (bi + 2) / 2 <= task) ++bi;
                    while (bi * (bi + 1) / 2 > task) --bi;
                    const int bj = task - bi * (bi + 1) / 2;
                    const bool diag = bi == bj;
                    const int ia = r0 + 16 * bi + fg, ib = ia + 8, ja = r0 + 16 * bj + fg, jb = ja + 8;
                    const bool oia = ia < n, oib = ib < n, oja = ja < n, ojb = jb < n;
                    const double* pia = A + pidx(oia ? ia : r0, c0) + ft;
                    const double* pib = A + pidx(oib ? ib : r0, c0) + ft;
                    const double* pja = A + pidx(oja ? ja : r0, c0) + ft;
                    const double* pjb = A + pidx(ojb ? jb : r0, c0) + ft;
                    double c00a = 0.0, c00b = 0.0, c01a = 0.0, c01b = 0.0, c10a = 0.0, c10b = 0.0, c11a = 0.0, c11b = 0.0;
#pragma unroll
                    for (int s4 = 0; s4 < NBK; s4 += 4) {
                        const bool okc = s4 + ft < nb;
                        const double a0 = (okc && oia) ? -pia[s4] : 0.0;
                        const double a1 = (okc && oib) ? -pib[s4] : 0.0;
                        double b0, b1;
                        if (diag) {
                            b0 = -a0;
                            b1 = -a1;
                        } else {
                            b0 = (okc && oja) ? pja[s4] : 0.0;
                            b1 = (okc && ojb) ? pjb[s4] : 0.0;
                        }
                        dmma884(c00a, c00b, a0, b0);
                        dmma884(c10a, c10b, a1, b0);
                        dmma884(c11a, c11b, a1, b1);
                        if (!diag) dmma884(c01a, c01b, a0, b1);
                    }
                    auto add = [&](int i, int j, double v0, double v1) {
                        if (i < n) {
                            if (j <= i) A[pidx(i, j)] += v0;
                            if (j + 1 <= i) A[pidx(i, j + 1)] += v1;
                        }
                    };
                    const int jc = r0 + 16 * bj + 2 * ft;
                    add(ia, jc, c00a, c00b);
                    add(ib, jc, c10a, c10b);
                    add(ib, jc + 8, c11a, c11b);
                    if (!diag) add(ia, jc + 8, c01a, c01b);
                }
            }
        }
        return __syncthreads_or(bad);  // barrier + team-uniform "some pivot was guarded" flag (warp 0 counted them)
    }
    for (int j = 0; j < n; ++j) {
        const double* rj = A + pidx(j, 0);
        double dj = rj[j];
        for (int p = 0; p < j; ++p) dj = fma(-rj[p], rj[p], dj);
        if (!(dj > 1e-280)) {
            dj = 1e200;
            ++bad;
        }
        const double ljj = sqrt(dj);
        const double inv = 1.0 / ljj;
        T.sync();  // everybody has read row j (incl. A[j][j]) before it is overwritten
        for (int i = j + T.tid; i < n; i += TEAM) {
            if (i == j) {
                A[pidx(j, j)] = ljj;
                invd[j] = inv;
            } else {
                double* ri = A + pidx(i, 0);
                double a = ri[j];
                for (int p = 0; p < j; ++p) a = fma(-ri[p], rj[p], a);
                ri[j] = a * inv;
            }
        }
        T.sync();
    }
    return bad;
}

// Solve L L' x = b in place (b -> x) with packed row-major L and invd.
template <int TEAM>
__device__ __forceinline__ void chol_solve(const Team<TEAM>& T, const double* L, const double* invd, double* b,
                                           int n) {
    if constexpr (TEAM >= 64) {
        // CTA teams: blocks of 32 unknowns; warp 0 solves the diagonal block on registers + shuffles (no barriers
        // inside), then every thread takes one remaining row (forward) / column entry (backward) of the update.
        const int warp = T.tid >> 5, lane = T.tid & 31;
        T.sync();
        for (int b0 = 0; b0 < n; b0 += 32) {  // forward: L y = b
            const int nb = min(32, n - b0);
            if (warp == 0) {
                const bool ok = lane < nb;
                const int i = ok ? b0 + lane : b0;
                double bi = ok ? b[i] : 0.0;
                const double inv = ok ? invd[i] : 0.0;
                const double* ri = L + pidx(i, b0);
                for (int j = 0; j < nb; ++j) {
                    const double yj = __shfl_sync(0xffffffffu, bi * inv, j);
                    if (lane > j && ok) bi = fma(-ri[j], yj, bi);
                }
                if (ok) b[i] = bi * inv;
            }
            T.sync();
            for (int i = b0 + 32 + T.tid; i < n; i += TEAM) {
                const double* ri = L + pidx(i, b0);
                double a0 = b[i], a1 = 0.0;
#pragma unroll 4
                for (int j = 0; j < 32; j += 2) {
                    a0 = fma(-ri[j], b[b0 + j], a0);
                    a1 = fma(-ri[j + 1], b[b0 + j + 1], a1);
                }
                b[i] = a0 + a1;
            }
            T.sync();
        }
        for (int b0 = ((n - 1) / 32) * 32; b0 >= 0; b0 -= 32) {  // backward: L' x = y
            const int nb = min(32, n - b0);
            if (warp == 0) {
                const bool ok = lane < nb;
                const int i = ok ? b0 + lane : b0;
                double bi = ok ? b[i] : 0.0;
                const double inv = ok ? invd[i] : 0.0;
                for (int j = nb - 1; j >= 0; --j) {
                    const double xj = __shfl_sync(0xffffffffu, bi * inv, j);
                    if (lane < j) bi = fma(-L[pidx(b0 + j, i)], xj, bi);
                }
                if (ok) b[i] = bi * inv;
            }
            T.sync();
            for (int i = T.tid; i < b0; i += TEAM) {
                double a0 = b[i], a1 = 0.0;
                int pj = pidx(b0, i);
                for (int j = 0; j < nb; ++j) {
                    const double t = L[pj] * b[b0 + j];
                    if (j & 1) a1 -= t; else a0 -= t;
                    pj += b0 + j + 1;
                }
                b[i] = a0 + a1;
            }
            T.sync();
        }
        return;
    }
    for (int j = 0; j < n; ++j) {  // forward
        T.sync();
        const double yj = b[j] * invd[j];
        for (int i = j + 1 + T.tid; i < n; i += TEAM) b[i] = fma(-L[pidx(i, j)], yj, b[i]);
        T.sync();
        if (T.tid == 0) b[j] = yj;
    }
    for (int j = n - 1; j >= 0; --j) {  // backward
        T.sync();
        const double xj = b[j] * invd[j];
        const double* rj = L + pidx(j, 0);
        for (int i = T.tid; i < j; i += TEAM) b[i] = fma(-rj[i], xj, b[i]);
        T.sync();
        if (T.tid == 0) b[j] = xj;
    }
    T.sync();
}

// ------------------------------------------------------------------------------------------
// Fused SteadyKalmanFilter (direct form): the reference's correct_estimate_obsv! / predict_estimate_obsv!
// (src/estimator/kalman.jl:284-309) on the team's shared-memory copy of x̂0.  tid/nth: thread index and count of the
// team; sync: the team's barrier.  Matrices are column-major.
// ------------------------------------------------------------------------------------------
template <class Sync>
__device__ __forceinline__ void skf_correct(const StepParams& P, long inst, int tid, int nth, double* sxh, const double* sd0,
                                            double* sev, Sync sync) {
    const int nx = P.nx, nym = P.nym, nd = P.nd;
    const double* Cm = P.eCm + inst * P.s_eCm;
    const double* K = P.eK + inst * P.s_eK;
    for (int j = tid; j < nym; j += nth) {  // innovation v = y0m - Ĉm x̂0 - D̂dm d0
        double v = P.y0m[inst * nym + j];
        for (int k = 0; k < nx; ++k) v = fma(-Cm[j + (long)nym * k], sxh[k], v);
        if (nd > 0) {
            const double* Dm = P.eDdm + inst * P.s_eDdm;
            for (int l = 0; l < nd; ++l) v = fma(-Dm[j + (long)nym * l], sd0[l], v);
        }
        sev[j] = v;
    }
    sync();
    for (int i = tid; i < nx; i += nth) {  // x̂0 <- x̂0 + K̂ v
        double a = sxh[i];
        for (int j = 0; j < nym; ++j) a = fma(K[i + (long)nx * j], sev[j], a);
        sxh[i] = a;
        P.xcorr[inst * nx + i] = a;
    }
    sync();
}
// x̂0(k+1) = Â x̂0 + B̂u u0 + B̂d d0 + (f̂op - x̂op), written to the handle's state; u0[l] = su[l] + du[l]
__device__ __forceinline__ void skf_predict(const StepParams& P, long inst, int tid, int nth, const double* sxh,
                                            const double* su, const double* du, const double* sd0) {
    const int nx = P.nx, nu = P.nu, nd = P.nd;
    const double* A = P.eA + inst * P.s_eA;
    const double* Bu = P.eBu + inst * P.s_eBu;
    for (int i = tid; i < nx; i += nth) {
        double a = P.efx ? P.efx[inst * P.s_efx + i] : 0.0;
        for (int k = 0; k < nx; ++k) a = fma(A[i + (long)nx * k], sxh[k], a);
        for (int l = 0; l < nu; ++l) a = fma(Bu[i + (long)nx * l], su[l] + du[l], a);
        if (nd > 0) {
            const double* Bd = P.eBd + inst * P.s_eBd;
            for (int l = 0; l < nd; ++l) a = fma(Bd[i + (long)nx * l], sd0[l], a);
        }
        P.xstate[inst * nx + i] = a;
    }
}

// ------------------------------------------------------------------------------------------
// Structured row operators.  Row r:  g_r'x = sig_r * (p_r'v) - c_r * eps
// ------------------------------------------------------------------------------------------
struct Ctx {
    const StepParams* P;
    const double* Pd;  // dense base rows [nDb x nz], column-major, ld = nDb (shared or global memory)
    double *x, *xb, *q, *rd, *rhs, *dx, *invd, *F, *tY, *fx, *yb, *ybd, *wd, *s, *lam, *h, *rp, *t, *ds, *dl, *Hv, *Phi;
};

// yout[k] = sum_j Pd[k, j] * v[j]
template <int TEAM>
__device__ __forceinline__ void dense_apply(const Team<TEAM>& T, const Ctx& c, const double* v, double* yout) {
    const int nDb = c.P->rt.nDb, nz = c.P->nz;
    for (int k = T.tid; k < nDb; k += TEAM) {
        double a0 = 0.0, a1 = 0.0;
        int j = 0;
        for (; j + 1 < nz; j += 2) {
            a0 = fma(c.Pd[k + (long)nDb * j], v[j], a0);
            a1 = fma(c.Pd[k + (long)nDb * (j + 1)], v[j + 1], a1);
        }
        if (j < nz) a0 = fma(c.Pd[k + (long)nDb * j], v[j], a0);
        yout[k] = a0 + a1;
    }
}

// value of g_r'vec for row r given the dense base products ybase = Pd * vec_v
__device__ __forceinline__ double row_gx(const Ctx& c, int r, const double* vec, const double* ybase) {
    const RowTables& rt = c.P->rt;
    const double eps = c.P->neps ? vec[c.P->nz] : 0.0;
    double base;
    if (r < rt.nS) {
        const int i2 = rt.s_i2[r];
        base = vec[rt.s_i1[r]] - (i2 >= 0 ? vec[i2] : 0.0);
    } else if (r < rt.nS + rt.nDr) {
        base = ybase[rt.dr_base[r - rt.nS]];
    } else {
        base = 0.0;
    }
    return rt.row_sig[r] * base - rt.row_c[r] * eps;
}

// out[0..n) (+)= G' w  : out_v[j] = sum_r sig_r w_r p_r[j],  out_eps = -sum_r c_r w_r
// scale: result = base[j] + alpha * (G'w)[j];  base may be nullptr (treated as 0)
template <int TEAM>
__device__ __forceinline__ void gt_apply(const Team<TEAM>& T, const Ctx& c, const double* w, double alpha,
                                         const double* base, double* out) {
    const RowTables& rt = c.P->rt;
    const int nDb = rt.nDb, nz = c.P->nz;
    for (int k = T.tid; k < nDb; k += TEAM) {
        const int a = rt.db_rmax[k], b = rt.db_rmin[k];
        c.wd[k] = (a >= 0 ? w[a] : 0.0) - (b >= 0 ? w[b] : 0.0);
    }
    double ce = 0.0;
    if (c.P->neps)
        for (int r = T.tid; r < rt.m; r += TEAM) ce = fma(rt.row_c[r], w[r], ce);
    T.sync();
    if (TEAM >= 256 || (TEAM >= 64 && !c.P->pd_in_smem)) {
        // Pd in L1/L2 (long columns, or CTA teams that keep it out of shared memory): one warp per PAIR of columns of Pd,
        // lanes stride over the rows (coalesced 256-byte requests instead of 32 scattered sectors, up to 8 loads in flight
        // per lane), shuffle reduction of both sums together; the few sparse rows in a second pass.  (The one-thread-per-
        // column loop below left 88 of 128 threads idle at n = 41 and walked each column two dependent loads at a time.)
        const int warp = T.tid >> 5, lane = T.tid & 31;
        constexpr int NW = (TEAM >= 32 ? TEAM / 32 : 1);
        for (int j = warp; j < nz; j += 2 * NW) {
            const int j2 = j + NW;
            const bool two = j2 < nz;
            const double* col = c.Pd + (long)nDb * j;
            const double* col2 = c.Pd + (long)nDb * (two ? j2 : j);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
            for (int k = lane; k < nDb; k += 128) {
                const bool o1 = k + 32 < nDb, o2 = k + 64 < nDb, o3 = k + 96 < nDb;
                const double p0 = col[k], p1 = o1 ? col[k + 32] : 0.0, p2 = o2 ? col[k + 64] : 0.0, p3 = o3 ? col[k + 96] : 0.0;
                const double q0 = col2[k], q1 = o1 ? col2[k + 32] : 0.0, q2 = o2 ? col2[k + 64] : 0.0, q3 = o3 ? col2[k + 96] : 0.0;
                const double w0 = c.wd[k], w1 = o1 ? c.wd[k + 32] : 0.0, w2 = o2 ? c.wd[k + 64] : 0.0, w3 = o3 ? c.wd[k + 96] : 0.0;
                a0 = fma(p0, w0, a0); a1 = fma(p1, w1, a1); a2 = fma(p2, w2, a2); a3 = fma(p3, w3, a3);
                b0 = fma(q0, w0, b0); b1 = fma(q1, w1, b1); b2 = fma(q2, w2, b2); b3 = fma(q3, w3, b3);
            }
            double acc = (a0 + a1) + (a2 + a3), acc2 = (b0 + b1) + (b2 + b3);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
                acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
            }
            if (lane == 0) {
                out[j] = acc;
                if (two) out[j2] = acc2;
            }
        }
        T.sync();
        for (int j = T.tid; j < nz; j += TEAM) {
            double acc = out[j];
            for (int e = rt.var_ptr[j]; e < rt.var_ptr[j + 1]; ++e) {
                const int r = rt.var_row[e];
                acc = fma((double)rt.var_sgn[e] * rt.row_sig[r], w[r], acc);
            }
            out[j] = (base ? base[j] : 0.0) + alpha * acc;
        }
    } else
    for (int j = T.tid; j < nz; j += TEAM) {
        double a0 = 0.0, a1 = 0.0;
        const double* col = c.Pd + (long)nDb * j;
        int k = (nDb > 0) ? (j % nDb) : 0;  // skewed start: conflict-free shared-memory columns
        int cnt = 0;
        for (; cnt + 1 < nDb; cnt += 2) {
            a0 = fma(col[k], c.wd[k], a0);
            k = (k + 1 == nDb) ? 0 : k + 1;
            a1 = fma(col[k], c.wd[k], a1);
            k = (k + 1 == nDb) ? 0 : k + 1;
        }
        if (cnt < nDb) a0 = fma(col[k], c.wd[k], a0);
        double acc = a0 + a1;
        for (int e = rt.var_ptr[j]; e < rt.var_ptr[j + 1]; ++e) {
            const int r = rt.var_row[e];
            acc = fma((double)rt.var_sgn[e] * rt.row_sig[r], w[r], acc);
        }
        out[j] = (base ? base[j] : 0.0) + alpha * acc;
    }
    if (c.P->neps) {
        ce = T.sum(ce);
        if (T.tid == 0) out[nz] = (base ? base[nz] : 0.0) - alpha * ce;
    }
    T.sync();
}

// hx[i] = sum_j Hv(i,j) v[j] (+ Hee*eps on the slack row)
template <int TEAM>
__device__ __forceinline__ void hess_apply(const Team<TEAM>& T, const Ctx& c, double Hee, const double* v,
                                           double* hx) {
    const int nz = c.P->nz;
    for (int i = T.tid; i < nz; i += TEAM) {
        double a = 0.0;
        const double* ri = c.Hv + pidx(i, 0);
        for (int j = 0; j <= i; ++j) a = fma(ri[j], v[j], a);
        for (int j = i + 1; j < nz; ++j) a = fma(c.Hv[pidx(j, i)], v[j], a);
        hx[i] = a;
    }
    if (c.P->neps && T.tid == 0) hx[nz] = Hee * v[nz];
}

// Phi = H + G' diag(d) G   (packed lower, n = nz + neps), d in c.t
template <int TEAM>
__device__ __forceinline__ void build_phi(const Team<TEAM>& T, const Ctx& c, double Hee, const double* d) {
    const RowTables& rt = c.P->rt;
    const int nDb = rt.nDb, nz = c.P->nz, neps = c.P->neps;
#ifdef BMPC_PHASE_CLK
    const long long bp_t0 = clock64();
#endif
    // dense weights: wd[k] = d_max + d_min ; ybd[k] = sig*c*d summed (for the slack border)
    for (int k = T.tid; k < nDb; k += TEAM) {
        const int a = rt.db_rmax[k], b = rt.db_rmin[k];
        double w = 0.0, g = 0.0;
        if (a >= 0) {
            w += d[a];
            g += d[a] * rt.row_c[a];
        }
        if (b >= 0) {
            w += d[b];
            g -= d[b] * rt.row_c[b];
        }
        c.wd[k] = w;
        c.ybd[k] = g;
    }
    T.sync();
    if constexpr (TEAM >= 256) {
        // ---- large n: same DMMA product with 32x32 output blocks (4x4 tiles) per warp: 8 fragment loads feed 16
        // DMMAs (10 on diagonal blocks), halving the L1/L2 traffic per flop of the 16x16 scheme below.
        const int warp = T.tid >> 5, lane = T.tid & 31, fg = lane >> 2, ft = lane & 3;
        constexpr int NW = TEAM / 32;
        const int nblk = (nz + 31) >> 5;
        const int ntask = nblk * (nblk + 1) / 2;
        for (int task = warp; task < ntask; task += NW) {
            int bi = (int)((sqrt(8.0 * task + 1.0) - 1.0) * 0.5);
            while ((bi + 1) * (bi + 2) / 2 <= task) ++bi;
            while (bi * (bi + 1) / 2 > task) --bi;
            const int bj = task - bi * (bi + 1) / 2;
            const bool diag = bi == bj;
            const double* pa[4];
            const double* pb[4];
            bool oa[4], ob[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ia = 32 * bi + 8 * u + fg, ja = 32 * bj + 8 * u + fg;
                oa[u] = ia < nz;
                ob[u] = ja < nz;
                pa[u] = c.Pd + (long)nDb * (oa[u] ? ia : 0) + ft;
                pb[u] = c.Pd + (long)nDb * (ob[u] ? ja : 0) + ft;
            }
            double acc[4][4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v][0] = acc[u][v][1] = 0.0;
            // software pipeline: the fragments of k-step k0+4 are requested from L1/L2 BEFORE the 16 DMMAs of k-step k0
            // are issued, so that the ~700-cycle L2 latency overlaps a full k-step of tensor work
            double av[4], bv[4], an[4], bn[4], w, wn;
            auto fetch = [&](int k0, double (&a4)[4], double (&b4)[4], double& wk) {
                const bool okk = k0 + ft < nDb;
                const int kk = okk ? k0 : 0;
                wk = okk ? c.wd[kk + ft] : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) a4[u] = (okk && oa[u]) ? pa[u][kk] : 0.0;
                if (!diag) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) b4[v] = (okk && ob[v]) ? pb[v][kk] : 0.0;
                }
            };
            fetch(0, av, bv, w);
            for (int k0 = 0; k0 < nDb; k0 += 4) {
                fetch(k0 + 4, an, bn, wn);  // (masked past the end)
#pragma unroll
                for (int v = 0; v < 4; ++v) bv[v] = (diag ? av[v] : bv[v]) * w;
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        if (!diag || v <= u) dmma884(acc[u][v][0], acc[u][v][1], av[u], bv[v]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    av[u] = an[u];
                    bv[u] = bn[u];
                }
                w = wn;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (diag && v > u) continue;
                    const int i = 32 * bi + 8 * u + fg, j = 32 * bj + 8 * v + 2 * ft;
                    if (i < nz) {
                        if (j <= i && j < nz) c.Phi[pidx(i, j)] = c.Hv[pidx(i, j)] + acc[u][v][0];
                        if (j + 1 <= i && j + 1 < nz) c.Phi[pidx(i, j + 1)] = c.Hv[pidx(i, j + 1)] + acc[u][v][1];
                    }
                }
        }
    } else if constexpr (TEAM >= 32) {
        // ---- FP64 tensor pipe: Phi_vv = Hv + Pd' diag(wd) Pd by 8x8x4 DMMA tiles.  Each warp of the team takes
        // 16x16 blocks (2x2 tiles) of the lower triangle; the A fragments (Pd columns of the block's rows) and the
        // B fragments (wd * Pd columns of the block's columns) come straight from Pd (column-major: a fragment is
        // 8 columns x 4 consecutive rows = full 32-byte sectors), which sits in shared memory or in L1/L2.
        const int warp = T.tid >> 5, lane = T.tid & 31, fg = lane >> 2, ft = lane & 3;
        constexpr int NW = TEAM / 32;
        const int nblk = (nz + 15) >> 4;
        const int ntask = nblk * (nblk + 1) / 2;
        for (int task = warp; task < ntask; task += NW) {
            int bi = (int)((sqrt(8.0 * task + 1.0) - 1.0) * 0.5);
            while ((bi + 1) * (bi + 2) / 2 <= task) ++bi;
            while (bi * (bi + 1) / 2 > task) --bi;
            const int bj = task - bi * (bi + 1) / 2;
            const int ia = 16 * bi + fg, ib = ia + 8, ja = 16 * bj + fg, jb = ja + 8;
            const bool oia = ia < nz, oib = ib < nz, oja = ja < nz, ojb = jb < nz;
            const double* pia = c.Pd + (long)nDb * (oia ? ia : 0) + ft;
            const double* pib = c.Pd + (long)nDb * (oib ? ib : 0) + ft;
            const double* pja = c.Pd + (long)nDb * (oja ? ja : 0) + ft;
            const double* pjb = c.Pd + (long)nDb * (ojb ? jb : 0) + ft;
            const bool diag = bi == bj;
            double c00a = 0.0, c00b = 0.0, c01a = 0.0, c01b = 0.0, c10a = 0.0, c10b = 0.0, c11a = 0.0, c11b = 0.0;
            // batches of four k-steps: the 16 (8 on a diagonal block) fragment loads of a batch are all issued before its
            // first DMMA, so a batch costs ONE L1/L2 round trip (a k-step at a time cost one each: ~780 cycles per k-step
            // measured with Pd in L2)
            for (int k0 = 0; k0 < nDb; k0 += 16) {
                double a0[4], a1[4], b0[4], b1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ks = k0 + 4 * u;
                    const bool okk = ks + ft < nDb;
                    const int kk = okk ? ks : 0;  // clamp (the value is masked)
                    const double w = okk ? c.wd[kk + ft] : 0.0;
                    a0[u] = (okk && oia) ? pia[kk] : 0.0;
                    a1[u] = (okk && oib) ? pib[kk] : 0.0;
                    if (diag) {
                        b0[u] = a0[u] * w;
                        b1[u] = a1[u] * w;
                    } else {
                        b0[u] = ((okk && oja) ? pja[kk] : 0.0) * w;
                        b1[u] = ((okk && ojb) ? pjb[kk] : 0.0) * w;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    dmma884(c00a, c00b, a0[u], b0[u]);
                    dmma884(c10a, c10b, a1[u], b0[u]);
                    dmma884(c11a, c11b, a1[u], b1[u]);
                    if (!diag) dmma884(c01a, c01b, a0[u], b1[u]);
                }
            }
            // write back the lower-triangle entries this lane holds: rows i = 16bi + 8ti + fg, cols j = 16bj + 8tj + 2ft (+1)
            auto put = [&](int i, int j, double v0, double v1) {
                if (i < nz) {
                    if (j <= i && j < nz) c.Phi[pidx(i, j)] = c.Hv[pidx(i, j)] + v0;
                    if (j + 1 <= i && j + 1 < nz) c.Phi[pidx(i, j + 1)] = c.Hv[pidx(i, j + 1)] + v1;
                }
            };
            const int jc = 16 * bj + 2 * ft;
            put(ia, jc, c00a, c00b);
            put(ib, jc, c10a, c10b);
            put(ib, jc + 8, c11a, c11b);
            if (!diag) put(ia, jc + 8, c01a, c01b);
        }
    } else {
        const int npair = nz * (nz + 1) / 2;
        for (int p = T.tid; p < npair; p += TEAM) {
            const int i = rt.pair_i[p], j = rt.pair_j[p];
            const double* ci = c.Pd + (long)nDb * i;
            const double* cj = c.Pd + (long)nDb * j;
            double a0 = 0.0, a1 = 0.0;
            int k = 0;
            for (; k + 1 < nDb; k += 2) {
                a0 = fma(ci[k] * c.wd[k], cj[k], a0);
                a1 = fma(ci[k + 1] * c.wd[k + 1], cj[k + 1], a1);
            }
            if (k < nDb) a0 = fma(ci[k] * c.wd[k], cj[k], a0);
            c.Phi[p] = c.Hv[p] + a0 + a1;
        }
    }
    T.sync();
#ifdef BMPC_PHASE_CLK
    if (T.tid == 0 && c.P->gclk) atomicAdd((unsigned long long*)&c.P->gclk[9], (unsigned long long)(clock64() - bp_t0));
#endif
    // 1-/2-variable rows: gather per variable (each packed entry has exactly one writer)
    for (int j = T.tid; j < nz; j += TEAM) {
        double dg = 0.0, be = 0.0;
        for (int e = rt.var_ptr[j]; e < rt.var_ptr[j + 1]; ++e) {
            const int r = rt.var_row[e];
            const double dr = d[r];
            dg += dr;
            const double sg = (double)rt.var_sgn[e];
            be = fma(sg * rt.row_sig[r] * rt.row_c[r], dr, be);
            if (rt.var_sgn[e] > 0) {
                const int i2 = rt.s_i2[r];
                if (i2 >= 0) c.Phi[pidx(j, i2)] -= dr;  // j = i1 > i2
            }
        }
        c.Phi[pidx(j, j)] += dg;
        if (neps) {
            // border: Phi[eps, j] = -sum_r d_r sig_r c_r p_r[j]
            double a = 0.0;
            const double* col = c.Pd + (long)nDb * j;
            for (int k = 0; k < nDb; ++k) a = fma(col[k], c.ybd[k], a);
            c.Phi[pidx(nz, j)] = -(a + be);
        }
    }
    if (neps) {
        double cc = 0.0;
        for (int r = T.tid; r < rt.m; r += TEAM) cc = fma(rt.row_c[r] * rt.row_c[r], d[r], cc);
        cc = T.sum(cc);
        if (T.tid == 0) c.Phi[pidx(nz, nz)] = Hee + cc;
    }
    T.sync();
}

// max step in [0,1] keeping s + a ds > 0 and lam + a dl > 0
template <int TEAM>
__device__ __forceinline__ double max_step(const Team<TEAM>& T, const Ctx& c, int m) {
    double a = 1.0;
    for (int r = T.tid; r < m; r += TEAM) {
        const double ds = c.ds[r], dl = c.dl[r];
        if (ds < 0.0) a = fmin(a, -c.s[r] / ds);
        if (dl < 0.0) a = fmin(a, -c.lam[r] / dl);
    }
    return T.min(a);
}

// Mehrotra predictor-corrector on  min 1/2 x'Hx + q'x  s.t. Gx + s = h  (structured rows, see RowTables).
// In: c.x = starting point, c.yb = Pd*x_v, c.s = h - Gx (raw slacks), c.q, c.h, c.Hv; out: c.x, status, iters.
// lam0: multipliers of the previous period (warm start; c.x / c.yb / c.s then describe the previous solution), or nullptr
template <int TEAM>
__device__ __forceinline__ void ipm_solve(const Team<TEAM>& T, const Ctx& c, const StepParams& P, double Hee, double qs,
                                          double hscale, int& status, int& iters, const double* lam0 = nullptr,
                                          double* kkt3 = nullptr) {
    const RowTables& rt = P.rt;
    const int n = P.n, m = rt.m, nDb = rt.nDb;
    const double mu0 = fmax(1e-2 * qs * hscale / (double)m, 1e-8);
    const double lmin = 1e-4 * qs / hscale;
    for (int r = T.tid; r < m; r += TEAM) {
        const double sv = fmax(c.s[r], 1e-2 * hscale);
        c.s[r] = sv;
        c.lam[r] = lam0 ? fmax(lam0[r], lmin) : mu0 / sv;
    }
    T.sync();
    status = ST_ITERATION_LIMIT;
    double rp_inf = 0.0, best_merit = 1e300;
    int stall = 0, stagn = 0;
    double mu_first = 0.0, ep_chk = 1e300;
    bool ray_prev = false;
    GCLK_DECL;
    for (int it = 0; it <= P.max_iter; ++it) {
        GCLK(P, T.tid, 8);
        // residuals
        hess_apply(T, c, Hee, c.x, c.rhs);  // rhs <- H x
        T.sync();
        for (int j = T.tid; j < n; j += TEAM) c.rhs[j] += c.q[j];
        T.sync();
        gt_apply(T, c, c.lam, 1.0, c.rhs, c.rd);  // rd = Hx + q + G' lam
        GCLK(P, T.tid, 0);
        double e_d = 0.0, e_p = 0.0, musum = 0.0, dsc = 0.0;
        for (int j = T.tid; j < n; j += TEAM) {
            e_d = fmax(e_d, fabs(c.rd[j]));
            dsc = fmax(dsc, fmax(fabs(c.rhs[j]), fabs(c.rd[j] - c.rhs[j])));  // |Hx + q|, |G'lam|
        }
        for (int r = T.tid; r < m; r += TEAM) {
            const double rpv = row_gx(c, r, c.x, c.yb) + c.s[r] - c.h[r];
            c.rp[r] = rpv;
            e_p = fmax(e_p, fabs(rpv));
            musum = fma(c.s[r], c.lam[r], musum);
        }
        e_d = T.max(e_d);
        e_p = T.max(e_p);
        // the dual residual is a difference of terms that can be far larger than q (slack weight 2*Cwt,
        // large multipliers): it is judged relative to the largest of them, as in OSQP's eps_rel test
        const double qd = qs + T.max(dsc);
        const double mu = T.sum(musum) / (double)m;
        rp_inf = e_p;
        if (!(e_d == e_d) || !(e_p == e_p) || !(mu == mu) || e_d > 1e250 || e_p > 1e250) {
            status = ST_INFEASIBLE;
            break;
        }
        // merit = worst of the three scaled KKT residuals (1.0 = exactly at tolerance)
        const double merit = fmax(fmax(e_d / (P.tol * qd), e_p / (P.tol * hscale)),
                                  mu * (double)m / (P.tol_mu * qs * hscale));
        if (kkt3 && merit <= best_merit) {  // residuals of the best iterate so far (the one a non-converged exit returns)
            kkt3[0] = e_p / hscale;
            kkt3[1] = e_d / qd;
            kkt3[2] = mu * (double)m / (qs * hscale);
        }
        if (merit <= 1.0) {
            status = ST_OPTIMAL;
            break;
        }
        // fp64 floor: once inside the acceptable level (1e3 x tol, still far below the reference
        // solver's eps = 1e-3) the first iteration that no longer improves the merit ends the solve:
        // the normal equations lose accuracy as lam/s spreads over > 1e24 and the dual residual
        // starts to grow again.
        if (best_merit <= 1e3 && merit >= best_merit) {
            T.sync();
            for (int j = T.tid; j < n; j += TEAM) c.x[j] = c.xb[j];  // the best iterate, not the current (worse) one
            T.sync();
            status = ST_OPTIMAL;
            break;
        }
        // degenerate problems (many nearly active rows with vanishing multipliers): Phi's condition number passes 1e16
        // before the tolerance is met and the dual residual starts to drift.  The best iterate is kept; three
        // iterations without improvement from a KKT residual already below 1e-6 (relative) end the solve there.
        if (merit < best_merit) {
            stagn = 0;
            for (int j = T.tid; j < n; j += TEAM) c.xb[j] = c.x[j];
        } else if (++stagn >= 3 && best_merit <= 1e5) {
            T.sync();
            for (int j = T.tid; j < n; j += TEAM) c.x[j] = c.xb[j];
            T.sync();
            status = ST_OPTIMAL;
            break;
        }
        best_merit = fmin(best_merit, merit);
        if (it == P.max_iter) {
            // iteration cap (the reference's time limit): keep the best iterate, status ITERATION_LIMIT unless it is
            // already acceptable -- the reference keeps the solver's value with a warning (execute.jl:482-503)
            if (merit <= 1e3) status = ST_OPTIMAL;
            else if (merit > best_merit) {
                T.sync();
                for (int j = T.tid; j < n; j += TEAM) c.x[j] = c.xb[j];
                T.sync();
            }
            break;
        }
        iters = it + 1;
        GCLK(P, T.tid, 1);
        // Phi = H + G' D G, factor
        for (int r = T.tid; r < m; r += TEAM) c.t[r] = c.lam[r] / c.s[r];
        T.sync();
        build_phi(T, c, Hee, c.t);
        GCLK(P, T.tid, 2);
        chol_packed(T, c.Phi, c.invd, n);
        GCLK(P, T.tid, 3);
        // predictor: rhs = -rd - G'(d*rp - lam)
        for (int r = T.tid; r < m; r += TEAM) c.dl[r] = c.t[r] * c.rp[r] - c.lam[r];
        T.sync();
        gt_apply(T, c, c.dl, 1.0, c.rd, c.dx);
        for (int j = T.tid; j < n; j += TEAM) c.dx[j] = -c.dx[j];
        GCLK(P, T.tid, 4);
        chol_solve(T, c.Phi, c.invd, c.dx, n);
        GCLK(P, T.tid, 5);
        dense_apply(T, c, c.dx, c.ybd);
        T.sync();
        for (int r = T.tid; r < m; r += TEAM) {
            const double dsv = -c.rp[r] - row_gx(c, r, c.dx, c.ybd);
            c.ds[r] = dsv;
            c.dl[r] = -c.lam[r] - c.t[r] * dsv;
        }
        T.sync();
        GCLK(P, T.tid, 6);
        const double a_aff = max_step(T, c, m);
        double mua = 0.0;
        for (int r = T.tid; r < m; r += TEAM)
            mua = fma(c.s[r] + a_aff * c.ds[r], c.lam[r] + a_aff * c.dl[r], mua);
        mua = T.sum(mua) / (double)m;
        double sig = mua / mu;
        sig = sig * sig * sig;
        // corrector: rc = s*lam + ds*dl - sig*mu ; rhs = -rd - G'((lam*rp - rc)/s)
        for (int r = T.tid; r < m; r += TEAM) {
            const double rc = c.s[r] * c.lam[r] + c.ds[r] * c.dl[r] - sig * mu;
            c.ds[r] = rc;  // keep rc
            c.dl[r] = (c.lam[r] * c.rp[r] - rc) / c.s[r];
        }
        T.sync();
        GCLK(P, T.tid, 7);
        gt_apply(T, c, c.dl, 1.0, c.rd, c.dx);
        for (int j = T.tid; j < n; j += TEAM) c.dx[j] = -c.dx[j];
        GCLK(P, T.tid, 4);
        chol_solve(T, c.Phi, c.invd, c.dx, n);
        GCLK(P, T.tid, 5);
        dense_apply(T, c, c.dx, c.ybd);
        T.sync();
        for (int r = T.tid; r < m; r += TEAM) {
            const double rc = c.ds[r];
            const double dsv = -c.rp[r] - row_gx(c, r, c.dx, c.ybd);
            c.ds[r] = dsv;
            c.dl[r] = -(rc + c.lam[r] * dsv) / c.s[r];
        }
        T.sync();
        GCLK(P, T.tid, 6);
        // fraction to the boundary: 0.99, tending to 1 as the affine step closes the gap (never exactly 1)
        const double tau = fmin(fmax(0.99, 1.0 - mua / mu), 1.0 - 1e-6);
        const double a = fmin(1.0, tau * max_step(T, c, m));
        // infeasibility: on an infeasible problem the multipliers of the conflicting rows diverge and the step length
        // collapses (1e-9, 1e-15, ...) while the primal residual stays where it is; two such steps end the solve
        // (status INFEASIBLE below) instead of running to the iteration cap.  Feasible problems never step below 1e-3.
        stall = (a < 1e-8 && e_p > 1e-6 * hscale) ? stall + 1 : 0;
        if (stall >= 2) {
            status = ST_INFEASIBLE;
            break;
        }
        // ... and before the collapse: an infeasible problem's iterates drift along a Farkas ray for 20-30 iterations
        // (primal residual on a plateau, complementarity GROWING by orders of magnitude, h'lam < 0) -- measured on
        // infeasible MHE windows, tools/studies/mhe_infeas.py.  Checked every 4 iterations from the 8th on (every 8 from the
        // 16th until tools/studies/mhe_infeas2.py showed the same false-positive set and 6 fewer iterations per infeasible window); all three
        // signs at TWO consecutive checkpoints end the solve, and only for problems without a slack variable (with one,
        // the dense rows can always be satisfied and a long plateau is just a hard but feasible problem).
        if (it == 0) mu_first = mu;
        if (P.neps == 0 && (it & 3) == 0) {
            bool ray = false;
            if (it >= 8 && e_p > 0.5 * ep_chk && e_p > 1e-4 * hscale && mu > 100.0 * mu_first) {
                double hl = 0.0;
                for (int r = T.tid; r < m; r += TEAM) hl = fma(c.h[r], c.lam[r], hl);
                ray = T.sum(hl) < 0.0;
            }
            if (ray && ray_prev) {
                status = ST_INFEASIBLE;
                break;
            }
            ray_prev = ray;
            ep_chk = e_p;
        }
        for (int j = T.tid; j < n; j += TEAM) c.x[j] = fma(a, c.dx[j], c.x[j]);
        for (int k = T.tid; k < nDb; k += TEAM) c.yb[k] = fma(a, c.ybd[k], c.yb[k]);
        for (int r = T.tid; r < m; r += TEAM) {
            c.s[r] = fma(a, c.ds[r], c.s[r]);
            c.lam[r] = fma(a, c.dl[r], c.lam[r]);
        }
        T.sync();
    }
    GCLK(P, T.tid, 8);
    GCLK_COUNT(P, T.tid, 16, iters);
    GCLK_COUNT(P, T.tid, 17, 1);
    // status INFEASIBLE is reserved for the certified exits above (collapsed steps, Farkas ray, NaN): an iteration-limit
    // exit keeps its iterate (general.jl `iserror`: only INFEASIBLE / NUMERICAL_ERROR ... discard the solver's value)
    (void)rp_inf;
}

// threads per CTA: sub-warp teams share a 128-thread CTA; a one-warp team is its own CTA (many CTAs per SM, bounded by
// the shared memory per instance); larger teams are one CTA each
template <int TEAM>
struct CtaThreads { static constexpr int value = TEAM < 32 ? 128 : TEAM; };

template <int TEAM>
__global__ void __launch_bounds__(CtaThreads<TEAM>::value, TEAM == 128 ? 6 : ((TEAM == 64 || TEAM == 32) ? 8 : 1))
    step_kernel(const __grid_constant__ StepParams P) {
    extern __shared__ __align__(128) double smem[];
    constexpr int CTA = CtaThreads<TEAM>::value;
    constexpr int TEAMS = CTA / TEAM;
    const int team_id = threadIdx.x / TEAM;
    Team<TEAM> T;
    T.tid = threadIdx.x % TEAM;
    if constexpr (TEAM < 32) {
        const int lane = threadIdx.x & 31;
        T.mask = ((1u << TEAM) - 1u) << (lane & ~(TEAM - 1));
    } else {
        T.mask = 0xffffffffu;
    }
    double* base = smem + (long)team_id * P.sm.total;
    T.red = base + P.sm.red;
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + P.sm.bar);
    volatile int* slot = reinterpret_cast<volatile int*>(base + P.sm.bar + 1);

    Ctx c;
    c.P = &P;
    c.x = base + P.sm.x;
    c.xb = base + P.sm.xb;
    c.q = base + P.sm.q;
    c.rd = base + P.sm.rd;
    c.rhs = base + P.sm.rhs;
    c.dx = base + P.sm.dx;
    c.invd = base + P.sm.invd;
    c.F = base + P.sm.F;
    c.tY = base + P.sm.tY;
    c.fx = base + P.sm.fx;
    c.yb = base + P.sm.yb;
    c.ybd = base + P.sm.ybd;
    c.wd = base + P.sm.wd;
    c.s = base + P.sm.s;
    c.lam = base + P.sm.lam;
    c.h = base + P.sm.h;
    c.rp = base + P.sm.rp;
    c.t = base + P.sm.t;
    c.ds = base + P.sm.ds;
    c.dl = base + P.sm.dl;
    c.Hv = base + P.sm.Hv;
    c.Phi = base + P.sm.Phi;
    double* sm_xhat = base + P.sm.xhat;
    double* sm_lastu = base + P.sm.lastu;
    double* sm_d0 = base + P.sm.dd;
    double* sm_Dh = base + P.sm.Dh;
    double* sm_Pd = base + P.sm.Pd;

    const RowTables& rt = P.rt;
    const int nz = P.nz, n = P.n, neps = P.neps, nY = P.nY, nu = P.nu, ny = P.ny, nx = P.nx, nd = P.nd;
    const int m = rt.m, nS = rt.nS, nDr = rt.nDr, nDb = rt.nDb;

    if (T.tid == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    T.sync();
    uint32_t phase = 0;

    if (threadIdx.x < 32) gather_wait_acks(P, threadIdx.x);
    __syncthreads();
    for (;;) {
        // ---- fetch the next instance (dynamic scheduling: iteration counts differ) ----
        if (T.tid == 0) *slot = (int)atomicAdd(&P.counters[0], 1u);
        T.sync();
        const int inst = *slot;
        T.sync();
        if (inst >= P.N) break;

        // ---- stage 0: TMA bulk loads ----
        const double* gHv = P.Hv + (long)inst * P.sH;
        const double* gLv = P.Lv + (long)inst * P.sH;
        const double* gPd = P.Pd + (long)inst * P.sPd;
        const int lv_ok = P.lv_ok[P.sH ? inst : 0];
        c.Hv = P.hv_in_smem ? base + P.sm.Hv : const_cast<double*>(gHv);
        if (T.tid == 0) {
            fence_proxy_async();
            uint32_t bytes = (uint32_t)P.nHp2 * 8u * (P.hv_in_smem ? 2u : 1u);
            if (P.pd_in_smem) bytes += (uint32_t)P.nPd2 * 8u;
            mbar_arrive_expect_tx(bar, bytes);
            if (P.hv_in_smem) tma_bulk_g2s(c.Hv, gHv, (uint32_t)P.nHp2 * 8u, bar);
            tma_bulk_g2s(c.Phi, gLv, (uint32_t)P.nHp2 * 8u, bar);
            if (P.pd_in_smem) tma_bulk_g2s(sm_Pd, gPd, (uint32_t)P.nPd2 * 8u, bar);
        }
        c.Pd = P.pd_in_smem ? sm_Pd : gPd;

        // ---- stage 1: initpred! ----
        const double* gxh = P.est_on ? P.xstate : P.xhat0;
        for (int k = T.tid; k < nx; k += TEAM) sm_xhat[k] = gxh[(long)inst * nx + k];
        for (int k = T.tid; k < nu; k += TEAM) sm_lastu[k] = P.lastu0[(long)inst * nu + k];
        if (nd > 0) {
            for (int k = T.tid; k < nd; k += TEAM) sm_d0[k] = P.d0[(long)inst * nd + k];
            for (int k = T.tid; k < nd * P.Hp; k += TEAM)
                sm_Dh[k] = P.Dhat0 ? P.Dhat0[(long)inst * nd * P.Hp + k] : P.d0[(long)inst * nd + (k % nd)];
        }
        T.sync();
        if (P.est_on) skf_correct(P, inst, T.tid, TEAM, sm_xhat, sm_d0, base + P.sm.ev, [&] { T.sync(); });
        const double* gK = P.K + (long)inst * P.sK;
        const double* gV = P.V + (long)inst * P.sV;
        const double* gB = P.B + (long)inst * P.sB;
        const double* gyop = P.yop + (long)inst * P.syop;
        const double* gM = P.Mw + (long)inst * P.sM;
        double racc = 0.0;
        for (int t = T.tid; t < nY; t += TEAM) {
            double f = gB[t] + (P.Ys ? P.Ys[(long)inst * nY + t] : 0.0);
            for (int k = 0; k < nx; ++k) f = fma(gK[t + (long)nY * k], sm_xhat[k], f);
            for (int k = 0; k < nu; ++k) f = fma(gV[t + (long)nY * k], sm_lastu[k], f);
            if (nd > 0) {
                const double* gG = P.G + (long)inst * P.sG;
                const double* gJ = P.J + (long)inst * P.sJ;
                for (int k = 0; k < nd; ++k) f = fma(gG[t + (long)nY * k], sm_d0[k], f);
                for (int k = 0; k < nd * P.Hp; ++k) f = fma(gJ[t + (long)nY * k], sm_Dh[k], f);
            }
            c.F[t] = f;
            const double ryt = P.Rhat_y ? P.Rhat_y[(long)inst * nY + t] : P.ry[(long)inst * ny + (t % ny)];
            const double cy = f + gyop[t % ny] - ryt;
            if (P.M_dense) {
                c.tY[t] = cy;  // tY = M*Cy is formed below
            } else {
                const double ty = gM[t] * cy;
                c.tY[t] = ty;
                racc = fma(cy, ty, racc);
            }
            P.F_out[(long)inst * nY + t] = f;
        }
        if (P.M_dense) {
            // tY = M * Cy (dense, symmetric, lower triangle authoritative)
            T.sync();
            double* cyv = c.dl;  // scratch: the layout reserves max(m, nY) doubles for dl
            for (int t = T.tid; t < nY; t += TEAM) cyv[t] = c.tY[t];
            T.sync();
            for (int t = T.tid; t < nY; t += TEAM) {
                double a = 0.0;
                for (int k = 0; k < nY; ++k) {
                    const double mk = (t >= k) ? gM[t + (long)nY * k] : gM[k + (long)nY * t];
                    a = fma(mk, cyv[k], a);
                }
                c.tY[t] = a;
                racc = fma(cyv[t], a, racc);
            }
        }
        if (P.has_terminal) {
            const double* gkx = P.kx + (long)inst * P.skx;
            const double* gvx = P.vx + (long)inst * P.svx;
            const double* gbx = P.bx + (long)inst * P.sbx;
            for (int i = T.tid; i < nx; i += TEAM) {
                double f = gbx[i];
                for (int k = 0; k < nx; ++k) f = fma(gkx[i + (long)nx * k], sm_xhat[k], f);
                for (int k = 0; k < nu; ++k) f = fma(gvx[i + (long)nx * k], sm_lastu[k], f);
                if (nd > 0) {
                    const double* ggx = P.gx + (long)inst * P.sgx;
                    const double* gjx = P.jx + (long)inst * P.sjx;
                    for (int k = 0; k < nd; ++k) f = fma(ggx[i + (long)nx * k], sm_d0[k], f);
                    for (int k = 0; k < nd * P.Hp; ++k) f = fma(gjx[i + (long)nx * k], sm_Dh[k], f);
                }
                c.fx[i] = f;
            }
        }
        T.sync();
        if (P.nw > 0) {  // linconstraint_custom! (execute.jl:337-366)
            custom_fw(P, inst, T.tid, TEAM, sm_xhat, sm_lastu, sm_d0, sm_Dh, c.F, base + P.sm.Fw);
            T.sync();
        }
        // q_v = 2 Ev' tY (+ 2 sum_{t in block} L_t Cu_t);   Ev is Pd when pd_is_ev
        mbar_wait(bar, phase);
        phase ^= 1u;
        const double* gEv = P.Ev + (long)inst * P.sEv;
        const double* gL = P.has_L ? P.Lw + (long)inst * P.sL : nullptr;
        const double* guop = P.uop + (long)inst * P.suop;
        double* s_cu = base + P.sm.cu;
        double* s_tU = base + P.sm.tU;
        if (gL && P.L_dense) {
            // dense L_Hp (ControllerWeights, construct.jl:45-93): tU = L_Hp Cu with Cu = Tu u0(k-1) + Uop - R̂u  (execute.jl:269-271)
            for (int idx = T.tid; idx < P.nU; idx += TEAM) {
                const int ch = idx % nu;
                const double ru = P.Rhat_u ? P.Rhat_u[(long)inst * P.nU + idx] : guop[ch];
                s_cu[idx] = sm_lastu[ch] + guop[ch] - ru;
            }
            T.sync();
            for (int idx = T.tid; idx < P.nU; idx += TEAM) {
                double a = 0.0;
                for (int k = 0; k < P.nU; ++k) {  // column-major, lower triangle authoritative (Hermitian)
                    const double lv = (idx >= k) ? gL[idx + (long)P.nU * k] : gL[k + (long)P.nU * idx];
                    a = fma(lv, s_cu[k], a);
                }
                s_tU[idx] = a;
                racc = fma(s_cu[idx], a, racc);
            }
            T.sync();
        }
        for (int j = T.tid; j < nz; j += TEAM) {
            const double* col = (P.pd_is_ev && P.pd_in_smem) ? (sm_Pd + (long)nY * j) : (gEv + (long)nY * j);
            double a0 = 0.0, a1 = 0.0;
            int k = j % nY, cnt = 0;
            for (; cnt + 1 < nY; cnt += 2) {
                a0 = fma(col[k], c.tY[k], a0);
                k = (k + 1 == nY) ? 0 : k + 1;
                a1 = fma(col[k], c.tY[k], a1);
                k = (k + 1 == nY) ? 0 : k + 1;
            }
            if (cnt < nY) a0 = fma(col[k], c.tY[k], a0);
            double a = a0 + a1;
            if (gL && P.L_dense) {
                const int l = j / nu, ch = j % nu;
                for (int tt = P.blk_start[l]; tt < P.blk_start[l + 1]; ++tt) a += s_tU[tt * nu + ch];
            } else if (gL) {
                const int l = j / nu, ch = j % nu;
                for (int tt = P.blk_start[l]; tt < P.blk_start[l + 1]; ++tt) {
                    const int idx = tt * nu + ch;
                    const double ru = P.Rhat_u ? P.Rhat_u[(long)inst * P.nU + idx] : guop[ch];
                    const double cu = sm_lastu[ch] + guop[ch] - ru;
                    a = fma(gL[idx], cu, a);
                }
            }
            c.q[j] = 2.0 * a;
        }
        if (gL && !P.L_dense) {
            for (int idx = T.tid; idx < P.nU; idx += TEAM) {
                const int ch = idx % nu;
                const double ru = P.Rhat_u ? P.Rhat_u[(long)inst * P.nU + idx] : guop[ch];
                const double cu = sm_lastu[ch] + guop[ch] - ru;
                racc = fma(gL[idx] * cu, cu, racc);
            }
        }
        if (neps && T.tid == 0) c.q[nz] = 0.0;
        const double rconst = T.sum(racc);
        // ---- linconstraint!: row right-hand sides ----
        const double* gsb = P.sbase + (long)inst * nS;
        const double* gdb = P.dbound + (long)inst * nDr;
        double hmax = 0.0;
        for (int r = T.tid; r < m; r += TEAM) {
            double hv;
            if (r < nS) {
                const int ch = rt.s_ch[r];
                hv = gsb[r] - (ch >= 0 ? rt.row_sig[r] * sm_lastu[ch] : 0.0);
            } else if (r < nS + nDr) {
                const int src = rt.dr_src[r - nS];
                const double fsrc = src < nY ? c.F[src] : (src < nY + nx ? c.fx[src - nY] : (base + P.sm.Fw)[src - nY - nx]);
                hv = rt.row_sig[r] * (gdb[r - nS] - fsrc);
            } else {
                hv = 0.0;
            }
            c.h[r] = hv;
            hmax = fmax(hmax, fabs(hv));
        }
        const double hscale = 1.0 + T.max(hmax);
        double qmax = 0.0;
        T.sync();
        for (int j = T.tid; j < n; j += TEAM) qmax = fmax(qmax, fabs(c.q[j]));
        const double qs = 1.0 + T.max(qmax);
        const double Hee = neps ? P.Hee[P.sH ? inst : 0] : 0.0;

        // ---- stage 2: unconstrained minimiser with the cached factor (Lv sits in Phi) ----
        for (int j = T.tid; j < n; j += TEAM) c.x[j] = (j < nz && lv_ok) ? -c.q[j] : 0.0;
        if (lv_ok) {
            for (int j = T.tid; j < nz; j += TEAM) c.invd[j] = 1.0 / c.Phi[pidx(j, j)];
            T.sync();
            chol_solve(T, c.Phi, c.invd, c.x, nz);
        }
        T.sync();
        dense_apply(T, c, c.x, c.yb);
        T.sync();
        double smin = 1e300;
        for (int r = T.tid; r < m; r += TEAM) {
            const double sl = c.h[r] - row_gx(c, r, c.x, c.yb);
            c.s[r] = sl;
            smin = fmin(smin, sl);
        }
        smin = (m > 0) ? T.min(smin) : 0.0;
        int status = ST_OPTIMAL;
        int iters = 0;
        const bool feasible = (m == 0) || (lv_ok && smin >= -1e-12 * hscale);
        if (!feasible) {
            // ---- stage 3: Mehrotra predictor-corrector, warm-started from the previous period when it converged ----
            const double* lam0 = nullptr;
            if (P.use_ws && P.ws_flag[inst] != 0) {
                lam0 = P.lam_ws + (long)inst * P.ws_stride;
                const double* gZp = P.Z + (long)inst * n;
                for (int j = T.tid; j < nz; j += TEAM) {  // previous Z̃ in level coordinates
                    double a = 0.0;
                    for (int l = j % nu; l <= j; l += nu) a += gZp[l];
                    c.x[j] = a;
                }
                if (neps && T.tid == 0) c.x[nz] = gZp[nz];
                T.sync();
                dense_apply(T, c, c.x, c.yb);
                T.sync();
                for (int r = T.tid; r < m; r += TEAM) c.s[r] = c.h[r] - row_gx(c, r, c.x, c.yb);
                T.sync();
            }
            double kkt3[3] = {0.0, 0.0, 0.0};
            ipm_solve(T, c, P, Hee, qs, hscale, status, iters, lam0, kkt3);
            if (P.kkt_out && T.tid == 0) {
                P.kkt_out[(long)inst * 3] = kkt3[0];
                P.kkt_out[(long)inst * 3 + 1] = kkt3[1];
                P.kkt_out[(long)inst * 3 + 2] = kkt3[2];
            }
            if (P.use_ws) {
                const bool keep = status == ST_OPTIMAL && iters > 0;
                if (keep)
                    for (int r = T.tid; r < m; r += TEAM) P.lam_ws[(long)inst * P.ws_stride + r] = c.lam[r];
                if (T.tid == 0) P.ws_flag[inst] = keep ? 1 : 0;
            }
        } else {
            if (P.use_ws && T.tid == 0) P.ws_flag[inst] = 0;
            if (P.kkt_out && T.tid == 0) {  // unconstrained minimiser feasible: an exact solve
                P.kkt_out[(long)inst * 3] = 0.0;
                P.kkt_out[(long)inst * 3 + 1] = 0.0;
                P.kkt_out[(long)inst * 3 + 2] = 0.0;
            }
        }
        // ---- stage 4: getinput! ----
        double* gZ = P.Z + (long)inst * n;
        T.sync();
        if (status == ST_INFEASIBLE) {
            // shifted previous solution Z̃s (set_warmstart_mpc!, transcription.jl:997-1007), in level coordinates
            for (int j = T.tid; j < nz; j += TEAM) c.dx[j] = (j + nu < nz) ? gZ[j + nu] : 0.0;
            if (neps && T.tid == 0) c.x[nz] = gZ[nz];
            T.sync();
            for (int j = T.tid; j < nz; j += TEAM) {
                double a = 0.0;
                for (int l = j % nu; l <= j; l += nu) a += c.dx[l];
                c.x[j] = a;
            }
            T.sync();
        }
        // J = 1/2 x'Hx + q'x + r
        hess_apply(T, c, Hee, c.x, c.rhs);
        T.sync();
        double jacc = 0.0;
        for (int j = T.tid; j < n; j += TEAM) jacc += c.x[j] * (0.5 * c.rhs[j] + c.q[j]);
        jacc = T.sum(jacc) + rconst;
        for (int j = T.tid; j < nz; j += TEAM) gZ[j] = c.x[j] - (j >= nu ? c.x[j - nu] : 0.0);
        if (neps && T.tid == 0) gZ[nz] = c.x[nz];
        if (P.zg_world > 0) {  // fused all-gather: this rank's slot buffer (pull protocol) or peer stores over NVLink (push)
            const long off = P.zg_base + (long)inst * n;
            for (int pr = P.zg_pull ? P.zg_rank : 0; pr < (P.zg_pull ? P.zg_rank + 1 : P.zg_world); ++pr) {
                double* dst = P.zg[pr] + off;
                for (int j = T.tid; j < nz; j += TEAM) dst[j] = c.x[j] - (j >= nu ? c.x[j - nu] : 0.0);
                if (neps && T.tid == 0) dst[nz] = c.x[nz];
            }
        }
        // q̃ in reference coordinates: q̃[j] = sum_{l' >= l(j)} q_v[l' nu + ch]
        for (int j = T.tid; j < nz; j += TEAM) {
            double a = 0.0;
            for (int l = j; l < nz; l += nu) a += c.q[l];
            P.qt_out[(long)inst * n + j] = a;
        }
        if (P.est_on) skf_predict(P, inst, T.tid, TEAM, sm_xhat, sm_lastu, c.x, sm_d0);
        for (int k = T.tid; k < nu; k += TEAM) {
            const double du = c.x[k];  // DU_0 = v_0
            const double lu = sm_lastu[k];
            P.lastu_prev[(long)inst * nu + k] = lu;
            P.lastu0[(long)inst * nu + k] = lu + du;
            P.u[(long)inst * nu + k] = lu + du + guop[k];
        }
        if (T.tid == 0) {
            if (neps) P.qt_out[(long)inst * n + nz] = 0.0;
            P.r_out[inst] = rconst;
            if (P.J_out) P.J_out[inst] = jacc;
            P.status[inst] = status;
            P.iters[inst] = iters;
        }
        fence_proxy_async();  // generic-proxy writes to Hv/Phi/Pd slices precede the next TMA overwrite
        T.sync();
    }
    // ---- reset the work counters for the next launch (last CTA out) ----
    __syncthreads();
    if (threadIdx.x == 0) {
        if (P.zg_world > 0 && !P.zg_pull) __threadfence_system(); else __threadfence();  // this CTA's (peer) stores before its arrival
        const unsigned done = atomicAdd(&P.counters[1], 1u);
        if (done == gridDim.x - 1) {
            P.counters[0] = 0u;
            P.counters[1] = 0u;
            __threadfence();
            publish_epoch(P);
        }
    }
    (void)TEAMS;
}

}  // namespace bmpc
