// Device code of the batched linear MovingHorizonEstimator step (sm_100a).
// Replaces, per instance and per period, the reference's
//   add_data_windows!            src/estimator/mhe/execute.jl:497-547
//   correct_cov! / invert_cov!   :729-752, 782-795  (KalmanFilter correction, kalman.jl:1235-1268)
//   initpred!                    :419-457   F, fx̄, H̃ = 2(Ñ + ẼZ̃'M ẼZ̃), q̃, r   (the per-step Hessian GEMM)
//   linconstraint!               src/estimator/mhe/transcription.jl:732-781
//   optim_objective! + getstate! execute.jl:576-638
// and, in a second small kernel, update_cov! (:755-779, KalmanFilter prediction kalman.jl:1270-1290).
// One CTA of 256 threads per instance; the QP is solved by the same structured interior-point code as
// LinMPC (ipm_solve, bmpc_device.cuh) after an unconstrained exit with a fresh Cholesky of H̃.
// Kernel variable order: x = [x̂0_arr (nx); Ŵ (nx*Nk); ε]  (the reference's Z̃ = [ε; x̂0_arr; Ŵ] is
// written back in its own order).  During the growing window (Nk < He) the problem is simply smaller:
// only the leading Nk blocks of every matrix are used, as in trunc_predmat (:658-681).
#pragma once
#include "bmpc_device.cuh"
#include "bmpc_kf.cuh"

namespace bmpc {

struct MheLayout {  // shared-memory offsets in doubles
    int Hv, Phi, x, xb, q, rd, rhs, dx, invd, yb, ybd, wd, s, lam, h, rp, t, ds, dl, F, FX, wrow, RF, P, P2, K, M, red, total;
};

struct MheParams {
    int N, nx, nu, nym, nd, He, Nk, neps, moving, direct;
    int ldE, ldEX;            // leading dimensions of E (nym*He) and EX (nx*He)
    long sE, sEX, sG, sGX, sJ, sJX, sB, sBX, sCm, sCov;   // per-instance strides (0 = shared model)
    const double *E, *EX, *G, *GX, *J, *JX, *B, *BX, *Cm, *Rm, *rinv, *Qinv;
    const double* Rinvd;      // dense R̂^-1 (nym x nym per model) when R̂ is not diagonal, else nullptr
    double Cwt;
    double *Y0m, *U0, *D0, *X0old, *x0arr, *Parr, *invP, *Z, *xhat0, *lastu0;  // handle-owned state
    const double *xmin, *xmax, *wmin, *wmax, *vmin, *vmax;                    // per-instance bounds
    const int *row_kind, *row_bidx;  // per row: 0 arrival, 1 process noise, 2 state, 3 sensor noise; bound index
    const double* Pd;                // gathered dense rows [nDb x nz] col-major (shared or per instance)
    long sPd;
    const double *y0m, *d0;          // inputs of this period
    double *J_out, *Vhat_out, *X0_out;
    int *status, *iters;
    // IPM warm start (moving window only): multipliers of the previous period [N x ws_stride], 1/0 flag per instance,
    // and the row map of the one-block window shift (new row r <- old row ws_map[r], -1: no predecessor)
    double* lam_ws;
    int* ws_flag;
    const int* ws_map;
    int ws_stride, use_ws;
    double* Hscratch;       // nullptr: the Hessian lives in shared memory
    long sHs;
    unsigned int* counter;  // work queue of this launch (zeroed by the host): instances are handed out dynamically
    MheLayout L;
};

template <int TEAM, int MINB = 1>
__global__ void __launch_bounds__(TEAM, MINB) mhe_step_kernel(const __grid_constant__ StepParams P,
                                                         const __grid_constant__ MheParams Q) {
    extern __shared__ __align__(128) double smem[];
    Team<TEAM> T;
    T.tid = threadIdx.x;
    T.mask = 0xffffffffu;
    const MheLayout& L = Q.L;
    T.red = smem + L.red;
    Ctx c;
    c.P = &P;
    c.x = smem + L.x; c.xb = smem + L.xb; c.q = smem + L.q; c.rd = smem + L.rd; c.rhs = smem + L.rhs; c.dx = smem + L.dx;
    c.invd = smem + L.invd; c.yb = smem + L.yb; c.ybd = smem + L.ybd; c.wd = smem + L.wd; c.s = smem + L.s;
    c.lam = smem + L.lam; c.h = smem + L.h; c.rp = smem + L.rp; c.t = smem + L.t; c.ds = smem + L.ds;
    c.dl = smem + L.dl; c.Phi = smem + L.Phi;
    // the rebuilt Hessian: shared memory, or (two CTAs per SM) a per-CTA scratch slice that stays in L2
    c.Hv = Q.Hscratch ? Q.Hscratch + (long)blockIdx.x * Q.sHs : smem + L.Hv;
    c.F = nullptr; c.tY = nullptr; c.fx = nullptr;
    double* sF = smem + L.F;
    double* sFX = smem + L.FX;
    double* wrow = smem + L.wrow;
    double* sP = smem + L.P;
    double* sP2 = smem + L.P2;
    double* sK = smem + L.K;
    double* sM = smem + L.M;
    const RowTables& rt = P.rt;
    const int nx = Q.nx, nu = Q.nu, nym = Q.nym, nd = Q.nd, He = Q.He, Nk = Q.Nk, neps = Q.neps;
    const int nz = P.nz, n = P.n, m = rt.m, nS = rt.nS, nDr = rt.nDr, nDb = rt.nDb;
    const int nYk = nym * Nk, nXk = nx * Nk, ldE = Q.ldE, ldEX = Q.ldEX;

    // dynamic work queue: the windows that need the interior-point solver (a quarter of them on the C3 workload, 15-40
    // iterations each) would otherwise pile up on whichever CTA a static stride happens to give them
    __shared__ int s_inst;
    for (;;) {
        if (T.tid == 0) s_inst = (int)atomicAdd(Q.counter, 1u);
        T.sync();
        const int inst = s_inst;
        T.sync();
        if (inst >= Q.N) break;
        double* Y0m = Q.Y0m + (long)inst * nym * He;
        double* U0 = Q.U0 + (long)inst * nu * He;
        double* D0 = Q.D0 + (long)inst * nd * (He + 1);
        double* X0old = Q.X0old + (long)inst * nx * He;
        double* x0arr = Q.x0arr + (long)inst * nx;
        double* Parr = Q.Parr + (long)inst * nx * nx;
        double* invP = Q.invP + (long)inst * nx * nx;
        double* xhat0 = Q.xhat0 + (long)inst * nx;
        const double* lastu0 = Q.lastu0 + (long)inst * nu;
        const double* y0m = Q.y0m + (long)inst * nym;
        const double* d0 = nd ? Q.d0 + (long)inst * nd : nullptr;
        const double* gE = Q.E + (long)inst * Q.sE;
        const double* gEX = Q.EX + (long)inst * Q.sEX;
        const double* gCm = Q.Cm + (long)inst * Q.sCm;
        const double* gRm = Q.Rm + (long)inst * Q.sCov * nym * nym;
        const double* grinv = Q.rinv + (long)inst * Q.sCov * nym;
        const double* gQinv = Q.Qinv + (long)inst * Q.sCov * nx * nx;
        c.Pd = Q.Pd + (long)inst * Q.sPd;

        // ---- add_data_windows! (Nk was already advanced by the host) ----
        if (Q.moving) {
            // shift every window by one block through shared scratch (sF is free here)
            for (int w = 0; w < 4; ++w) {
                double* win = w == 0 ? Y0m : (w == 1 ? U0 : (w == 2 ? X0old : D0));
                const int blk = w == 0 ? nym : (w == 1 ? nu : (w == 2 ? nx : nd));
                const int len = w == 3 ? nd * (He + 1) : blk * He;
                if (blk == 0) continue;
                for (int k = T.tid; k < len - blk; k += TEAM) c.Phi[k] = win[k + blk];
                T.sync();
                for (int k = T.tid; k < len - blk; k += TEAM) win[k] = c.Phi[k];
                T.sync();
            }
        }
        for (int k = T.tid; k < nym; k += TEAM) Y0m[nym * (Nk - 1) + k] = y0m[k];
        for (int k = T.tid; k < nu; k += TEAM) U0[nu * (Nk - 1) + k] = lastu0[k];
        for (int k = T.tid; k < nd; k += TEAM) D0[nd * Nk + k] = d0[k];
        for (int k = T.tid; k < nx; k += TEAM) X0old[nx * (Nk - 1) + k] = xhat0[k];
        __threadfence_block();
        T.sync();
        for (int k = T.tid; k < nx; k += TEAM) x0arr[k] = X0old[k];
        T.sync();

        // ---- correct_cov!: Kalman correction of the arrival covariance, then its inverse (direct = true only:
        //      with direct = false the correction is part of update_cov!, k_mhe_update) ----
        if (Q.moving && Q.direct) {
            for (int e = T.tid; e < nx * nx; e += TEAM) sP[e] = Parr[e];
            T.sync();
            kf_correct_cov(T.tid, TEAM, [&] { T.sync(); }, nx, nym, gCm, gRm, sP, sK, sM, sP2);
            for (int e = T.tid; e < nx * nx; e += TEAM) {  // Hermitian(:L)
                const int i = e % nx, j = e / nx;
                Parr[e] = i >= j ? sP2[e] : sP2[j + nx * i];
            }
            T.sync();
            for (int e = T.tid; e < nx * nx; e += TEAM) sP[e] = Parr[e];
            T.sync();
            if (T.tid == 0) {
                if (small_spd_inverse(sP, sP2, sM + nym * nym, nx))
                    for (int e = 0; e < nx * nx; ++e) invP[e] = sP2[e];  // else: keep the old inverse (:785-793)
            }
            T.sync();
        }
        // ---- initpred!: F, FX ----
        const double* gG = Q.G + (long)inst * Q.sG;
        const double* gB = Q.B + (long)inst * Q.sB;
        for (int t = T.tid; t < nYk; t += TEAM) {
            double f = Y0m[t] + gB[t];
            for (int k = 0; k < nu * Nk; ++k) f = fma(gG[t + (long)ldE * k], U0[k], f);
            if (nd) {
                const double* gJ = Q.J + (long)inst * Q.sJ;
                for (int k = 0; k < nd * (Nk + 1); ++k) f = fma(gJ[t + (long)ldE * k], D0[k], f);
            }
            const bool bad = !(f == f);  // NaN measurement: the row leaves the objective (:436-441)
            sF[t] = bad ? 0.0 : f;
            wrow[t] = bad ? 0.0 : (Q.Rinvd ? 1.0 : grinv[t % nym]);  // dense R̂: wrow is the 0/1 row mask
        }
        const double* gGX = Q.GX + (long)inst * Q.sGX;
        const double* gBX = Q.BX + (long)inst * Q.sBX;
        for (int t = T.tid; t < nXk; t += TEAM) {
            double f = gBX[t];
            for (int k = 0; k < nu * Nk; ++k) f = fma(gGX[t + (long)ldEX * k], U0[k], f);
            if (nd) {
                const double* gJX = Q.JX + (long)inst * Q.sJX;
                for (int k = 0; k < nd * (Nk + 1); ++k) f = fma(gJX[t + (long)ldEX * k], D0[k], f);
            }
            sFX[t] = f;
        }
        for (int e = T.tid; e < nx * nx; e += TEAM) sP[e] = invP[e];
        T.sync();
        // ---- H = 2(E' Rinv E + blockdiag(invP, Qinv...)), q = 2(E' Rinv F - invP fx̄), r ----
        const int npair = nz * (nz + 1) / 2;
        const double* gRi = Q.Rinvd ? Q.Rinvd + (long)inst * Q.sCov * nym * nym : nullptr;
        for (int p = T.tid; p < npair; p += TEAM) {
            const int i = rt.pair_i[p], j = rt.pair_j[p];
            const double* ci = gE + (long)ldE * i;
            const double* cj = gE + (long)ldE * j;
            double a0 = 0.0, a1 = 0.0;
            if (gRi) {
                // non-diagonal R̂: invR̂_He = blockdiag(R̂^-1) couples the outputs of one time step
                // (src/estimator/construct.jl:60-119); rows of missing measurements are masked out on both sides
                for (int tb = 0; tb < nYk; tb += nym)
                    for (int o = 0; o < nym; ++o) {
                        const double ei = ci[tb + o] * wrow[tb + o];
                        if (ei == 0.0) continue;
                        double w = 0.0;
                        for (int o2 = 0; o2 < nym; ++o2) w = fma(gRi[o + nym * o2] * wrow[tb + o2], cj[tb + o2], w);
                        a0 = fma(ei, w, a0);
                    }
            } else {
                int t = 0;
                for (; t + 1 < nYk; t += 2) {
                    a0 = fma(ci[t] * wrow[t], cj[t], a0);
                    a1 = fma(ci[t + 1] * wrow[t + 1], cj[t + 1], a1);
                }
                if (t < nYk) a0 = fma(ci[t] * wrow[t], cj[t], a0);
            }
            double a = a0 + a1;
            if (i < nx) {
                a += sP[i + nx * j];  // arrival block: ex̄' invP̄ ex̄ = invP̄
            } else if (j >= nx && (i - nx) / nx == (j - nx) / nx) {
                a += gQinv[(i - nx) % nx + nx * ((j - nx) % nx)];
            }
            c.Hv[p] = 2.0 * a;
        }
        double racc = 0.0;
        double* sRF = smem + L.RF;  // R̂^-1-weighted F
        T.sync();
        for (int t = T.tid; t < nYk; t += TEAM) {
            double w;
            if (gRi) {
                const int tb = t - t % nym, o = t % nym;
                w = 0.0;
                for (int o2 = 0; o2 < nym; ++o2) w = fma(gRi[o + nym * o2] * wrow[tb + o2], sF[tb + o2], w);
                w *= wrow[t];
            } else {
                w = wrow[t] * sF[t];
            }
            sRF[t] = w;
        }
        T.sync();
        for (int i = T.tid; i < nz; i += TEAM) {
            const double* ci = gE + (long)ldE * i;
            double a = 0.0;
            for (int t = 0; t < nYk; ++t) a = fma(ci[t], sRF[t], a);
            if (i < nx) {
                double b = 0.0;
                for (int k = 0; k < nx; ++k) b = fma(sP[i + nx * k], x0arr[k], b);
                a -= b;
                racc = fma(b, x0arr[i], racc);
            }
            c.q[i] = 2.0 * a;
        }
        if (neps && T.tid == 0) c.q[nz] = 0.0;
        for (int t = T.tid; t < nYk; t += TEAM) racc = fma(sRF[t], sF[t], racc);
        const double rconst = T.sum(racc);
        const double Hee = neps ? 2.0 * Q.Cwt : 0.0;
        // ---- linconstraint!: row right-hand sides ----
        const double* bnd[6] = {Q.xmin + (long)inst * nx, Q.xmax + (long)inst * nx, Q.wmin + (long)inst * nx,
                                Q.wmax + (long)inst * nx, Q.vmin + (long)inst * nym, Q.vmax + (long)inst * nym};
        double hmax = 0.0;
        for (int r = T.tid; r < m; r += TEAM) {  // (the redundant eps >= 0 row is not compiled)
            const int kind = Q.row_kind[r], bi = Q.row_bidx[r];
            const double sg = rt.row_sig[r];
            const int side = sg > 0 ? 1 : 0;
            double hv;
            if (kind == 0) hv = sg * bnd[side][bi];
            else if (kind == 1) hv = sg * bnd[2 + side][bi % nx];
            else if (kind == 2) hv = sg * (bnd[side][bi % nx] - sFX[bi]);
            else hv = sg * (bnd[4 + side][bi % nym] - sF[bi]);
            c.h[r] = hv;
            hmax = fmax(hmax, fabs(hv));
        }
        const double hscale = 1.0 + T.max(hmax);
        double qmax = 0.0;
        T.sync();
        for (int j = T.tid; j < n; j += TEAM) qmax = fmax(qmax, fabs(c.q[j]));
        const double qs = 1.0 + T.max(qmax);
        // ---- unconstrained minimiser: fresh Cholesky of H (the Hessian changes every period) ----
        for (int p = T.tid; p < npair; p += TEAM) c.Phi[p] = c.Hv[p];
        for (int j = T.tid; j < n; j += TEAM) c.x[j] = j < nz ? -c.q[j] : 0.0;
        T.sync();
        const int bad = chol_packed(T, c.Phi, c.invd, nz);
        chol_solve(T, c.Phi, c.invd, c.x, nz);
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
        int status = ST_OPTIMAL, iters = 0;
        const bool feasible = (m == 0) || (bad == 0 && smin >= -1e-12 * hscale);
        double* gZ = Q.Z + (long)inst * (neps + nx + nx * He);
        if (!feasible) {
            // warm start from the previous window shifted by one block: Z̃s of set_warmstart_mhe! (arrival state =
            // x̂0arr_old, Ŵ shifted, transcription.jl:967-1001) and the multipliers of the rows that stay in the window
            const double* lam0 = nullptr;
            if (Q.use_ws && Q.moving && Q.ws_flag[inst] != 0) {
                const double* glw = Q.lam_ws + (long)inst * Q.ws_stride;
                for (int r = T.tid; r < m; r += TEAM) {
                    const int o = Q.ws_map[r];
                    c.dl[r] = o >= 0 ? glw[o] : 0.0;
                }
                for (int j = T.tid; j < nz; j += TEAM)
                    c.x[j] = j < nx ? x0arr[j] : ((j + nx < nx + nx * He) ? gZ[neps + j + nx] : 0.0);
                if (neps && T.tid == 0) c.x[nz] = gZ[0];
                T.sync();
                dense_apply(T, c, c.x, c.yb);
                T.sync();
                for (int r = T.tid; r < m; r += TEAM) c.s[r] = c.h[r] - row_gx(c, r, c.x, c.yb);
                T.sync();
                lam0 = c.dl;
            }
            ipm_solve(T, c, P, Hee, qs, hscale, status, iters, lam0);
            if (Q.use_ws) {
                const bool keep = status == ST_OPTIMAL && iters > 0;
                if (keep)
                    for (int r = T.tid; r < m; r += TEAM) Q.lam_ws[(long)inst * Q.ws_stride + r] = c.lam[r];
                if (T.tid == 0) Q.ws_flag[inst] = keep ? 1 : 0;
            }
        } else if (Q.use_ws && T.tid == 0) {
            Q.ws_flag[inst] = 0;
        }
        T.sync();
        // ---- outputs: Z̃ (reference order), getstate! ----
        if (status == ST_INFEASIBLE) {
            // warm start Z̃s (set_warmstart_mhe!, transcription.jl:967-1001): arrival = x̂0arr_old, Ŵ shifted
            for (int j = T.tid; j < nz; j += TEAM) {
                double v;
                if (j < nx) v = x0arr[j];
                else v = (j + nx < nx + nx * He) ? gZ[neps + j + nx] : 0.0;
                c.dx[j] = v;
            }
            T.sync();
            for (int j = T.tid; j < nz; j += TEAM) c.x[j] = c.dx[j];
            if (neps && T.tid == 0) c.x[nz] = gZ[0];
            T.sync();
        }
        hess_apply(T, c, Hee, c.x, c.rhs);
        T.sync();
        double jacc = 0.0;
        for (int j = T.tid; j < n; j += TEAM) jacc += c.x[j] * (0.5 * c.rhs[j] + c.q[j]);
        jacc = T.sum(jacc) + rconst;
        for (int j = T.tid; j < nx + nx * He; j += TEAM) gZ[neps + j] = j < nz ? c.x[j] : 0.0;  // fill0unused!
        if (neps && T.tid == 0) gZ[0] = c.x[nz];
        // X̂0 = EX z + FX ; x̂0 = last block
        for (int t = T.tid; t < nXk; t += TEAM) {
            double a = sFX[t];
            for (int j = 0; j < nz; ++j) a = fma(gEX[t + (long)ldEX * j], c.x[j], a);
            if (Q.X0_out) Q.X0_out[(long)inst * nx * He + t] = a;
            if (t >= nXk - nx) xhat0[t - (nXk - nx)] = a;
        }
        if (Q.Vhat_out)
            for (int t = T.tid; t < nYk; t += TEAM) {
                double a = sF[t];
                for (int j = 0; j < nz; ++j) a = fma(gE[t + (long)ldE * j], c.x[j], a);
                Q.Vhat_out[(long)inst * nym * He + t] = a;
            }
        if (T.tid == 0) {
            if (Q.J_out) Q.J_out[inst] = jacc;
            Q.status[inst] = status;
            Q.iters[inst] = iters;
        }
        T.sync();
    }
}

// update_estimate!: lastu0 <- u0 and, if the window is full, update_cov! (execute.jl:755-779):
//   direct = true :  P̄ <- Â P̄ Â' + Q̂                      (the correction ran in correct_cov!, step kernel)
//   direct = false:  P̄ <- Â [(I - K Ĉm) P̄] Â' + Q̂          (KalmanFilter update_estimate! = correct + predict,
//                                                           kalman.jl:520-525)
// then Hermitian(:L) and the inverse (kept when the factorisation fails, :785-793).
// Shared memory (doubles): P nx^2 | T1 nx^2 | P2 2 nq^2 + nx nym | K nx nym | M nym^2 + nq^2, nq = max(nx, nym).
__global__ void k_mhe_update(int N, int nx, int nu, int nym, int full, int direct, const double* __restrict__ A, long sA,
                             const double* __restrict__ Qc, long sQ, const double* __restrict__ Cm, long sCm,
                             const double* __restrict__ Rm, long sR, double* __restrict__ Parr,
                             double* __restrict__ invP, double* __restrict__ lastu0, const double* __restrict__ u0) {
    extern __shared__ double sm[];
    const int inst = blockIdx.x;
    if (inst >= N) return;
    const int nq = nx > nym ? nx : nym;
    double* P = sm;
    double* T1 = sm + nx * nx;
    double* P2 = T1 + nx * nx;
    double* K = P2 + 2 * nq * nq + nx * nym;
    double* M = K + nx * nym;
    const double* gA = A + inst * sA;
    const double* gQ = Qc + inst * sQ;
    for (int k = threadIdx.x; k < nu; k += blockDim.x) lastu0[(long)inst * nu + k] = u0[(long)inst * nu + k];
    if (!full) return;
    double* gP = Parr + (long)inst * nx * nx;
    for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) P[e] = gP[e];
    __syncthreads();
    const double* src = P;
    if (!direct) {
        kf_correct_cov((int)threadIdx.x, (int)blockDim.x, [] { __syncthreads(); }, nx, nym, Cm + inst * sCm, Rm + inst * sR,
                       P, K, M, P2);
        src = P2;
    }
    for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) {  // T1 = src A'
        const int i = e % nx, j = e / nx;
        double a = 0.0;
        for (int k = 0; k < nx; ++k) a = fma(src[i + nx * k], gA[j + nx * k], a);
        T1[e] = a;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) {  // P = A T1 + Q  (src is dead)
        const int i = e % nx, j = e / nx;
        double a = gQ[e];
        for (int k = 0; k < nx; ++k) a = fma(gA[i + nx * k], T1[k + nx * j], a);
        P[e] = a;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nx * nx; e += blockDim.x) {  // Hermitian(:L) -> gP, T1
        const int i = e % nx, j = e / nx;
        const double v = i >= j ? P[e] : P[j + nx * i];
        gP[e] = v;
        T1[e] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (small_spd_inverse(T1, P2, P, nx))
            for (int e = 0; e < nx * nx; ++e) invP[(long)inst * nx * nx + e] = P2[e];
    }
}

}  // namespace bmpc
