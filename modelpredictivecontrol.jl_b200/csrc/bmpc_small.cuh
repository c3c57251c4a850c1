// Register-resident step kernel for SMALL controllers (n = nu*Hc + neps <= 16): the shape of
// BASELINE.json's headline config (C1: n = 11, 40 dense Ŷ rows, 20 merged input-box rows).
//
// Mapping: 16 lanes per controller instance, two instances per warp running in LOCK-STEP (all
// control flow is warp-uniform; a finished instance is masked by a zero step length).
//   lane i  <->  decision variable i: it holds row i of Phi = H + G'DG (and then of its Cholesky
//               factor) in registers, plus x_i, q_i, rd_i ... as scalars;
//   lane l  <->  constraint rows l, l+16, l+32 ... (s, lambda, h in registers);
//   the instance's dense rows Pd (row-major), Hessian Hv and cached factor Lv arrive by TMA bulk
//   copies (cp.async.bulk + mbarrier) into the team's shared-memory slice;
//   Cholesky and the forward solve run on registers + warp shuffles, the backward solve reads the
//   factor's rows back from shared memory; reductions are xor-shuffles inside the 16-lane half.
// The number of move variables is a template parameter NZT (full unrolling, immediate shared-memory
// offsets, 128-bit loads); a controller with fewer variables is padded with decoupled dummy variables
// (unit Hessian diagonal, zero columns in Pd) whose solution is exactly 0.
// Same algorithm, tolerances and outputs as the general kernel in bmpc_device.cuh.
#pragma once
#include "bmpc_device.cuh"

namespace bmpc {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double half_sum(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double half_max(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double half_min(double v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

struct SmallLayout {  // per-team shared-memory offsets (doubles) + CTA table offsets
    int Pd, Hv, Lb, vbuf, wd, wp, ws, bb, F, tY, fx, xh, lu, dd, Dh, bar, team_total;
    int t_sigd, t_cd, t_srcd, t_sigs, t_cs, t_i1, t_i2, t_ch, t_vptr, t_vrow, t_vsgn, tab_total;
    int nPdR, nHS, nDbp;  // padded sizes (doubles) of the TMA sources; dense rows padded to even
};

struct SmallParams {
    SmallLayout L;
    const double *PdR, *HvS, *LvS;  // row-major padded copies (TMA sources)
    long sPdR, sHS;
    int has_pair_rows;  // some sparse row touches two variables (DU bounds)
    int nsr;            // sparse rows incl. the eps >= 0 row
};

template <int NZT>
struct SmallDims {
    static constexpr int LDP = (NZT + 1) & ~1;  // even: rows 16-byte aligned (LDS.128)
};

template <int NZT, int NEPS, int DS, int SS>
__global__ void __launch_bounds__(64, 7)
    step_small(const __grid_constant__ StepParams P, const __grid_constant__ SmallParams Q) {
    constexpr int NT = NZT + NEPS;
    constexpr int LDP = SmallDims<NZT>::LDP, LDH = LDP, LDN = NT | 1, NV2 = LDP / 2;
    extern __shared__ __align__(128) double smem[];
    const SmallLayout& L = Q.L;
    const RowTables& rt = P.rt;
    const int lane = threadIdx.x & 31, l16 = lane & 15, hb = lane & 16;
    const int team = threadIdx.x >> 4;
    const int nzr = P.nz, nr = P.n;  // real sizes (<= NZT, NT)
    const int nY = P.nY, nu = P.nu, ny = P.ny, nx = P.nx, nd = P.nd;
    const int nS = rt.nS, nDb = rt.nDb, nDbp = L.nDbp, nsr = Q.nsr, m = rt.m;
    double* tb = smem + (long)team * L.team_total;
    double* sPd = tb + L.Pd;
    double* sHv = tb + L.Hv;
    double* sLb = tb + L.Lb;
    double* vbuf = tb + L.vbuf;
    double* wd = tb + L.wd;
    double2* wp = reinterpret_cast<double2*>(tb + L.wp);
    double* ws = tb + L.ws;
    double* bb = tb + L.bb;
    double* sF = tb + L.F;
    double* stY = tb + L.tY;
    double* sfx = tb + L.fx;
    double* sxh = tb + L.xh;
    double* slu = tb + L.lu;
    double* sd0 = tb + L.dd;
    double* sDh = tb + L.Dh;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tb + L.bar);
    // ---- CTA-wide constant tables ----
    double* tabs = smem + (long)(blockDim.x >> 4) * L.team_total;
    double* t_sigd = tabs + L.t_sigd;
    double* t_cd = tabs + L.t_cd;
    int* t_srcd = reinterpret_cast<int*>(tabs + L.t_srcd);
    double* t_sigs = tabs + L.t_sigs;
    double* t_cs = tabs + L.t_cs;
    int* t_i1 = reinterpret_cast<int*>(tabs + L.t_i1);
    int* t_i2 = reinterpret_cast<int*>(tabs + L.t_i2);
    int* t_ch = reinterpret_cast<int*>(tabs + L.t_ch);
    int* t_vptr = reinterpret_cast<int*>(tabs + L.t_vptr);
    int* t_vrow = reinterpret_cast<int*>(tabs + L.t_vrow);
    int* t_vsgn = reinterpret_cast<int*>(tabs + L.t_vsgn);
    for (int k = threadIdx.x; k < nDb; k += blockDim.x) {
        t_sigd[k] = rt.row_sig[nS + k];
        t_cd[k] = rt.row_c[nS + k];
        t_srcd[k] = rt.dr_src[k];
    }
    for (int r = threadIdx.x; r < nsr; r += blockDim.x) {
        const bool isrow = r < nS;
        t_sigs[r] = isrow ? rt.row_sig[r] : 0.0;
        t_cs[r] = isrow ? rt.row_c[r] : 1.0;  // the eps >= 0 row: g = -eps
        t_i1[r] = isrow ? rt.s_i1[r] : -1;
        t_i2[r] = isrow ? rt.s_i2[r] : -1;
        t_ch[r] = isrow ? rt.s_ch[r] : -1;
    }
    for (int j = threadIdx.x; j <= NZT; j += blockDim.x) t_vptr[j] = rt.var_ptr[min(j, nzr)];
    for (int e = threadIdx.x; e < rt.var_ptr[nzr]; e += blockDim.x) {
        t_vrow[e] = rt.var_row[e];
        t_vsgn[e] = rt.var_sgn[e];
    }
    // zero the weight buffers once (covers the padding row when nDb is odd)
    for (int k = threadIdx.x; k < (int)(blockDim.x >> 4) * 16 * DS; k += blockDim.x) {
        const int tm = k / (16 * DS), kk = k % (16 * DS);
        double* b2 = smem + (long)tm * L.team_total;
        b2[L.wd + kk] = 0.0;
        b2[L.wp + 2 * kk] = 0.0;
        b2[L.wp + 2 * kk + 1] = 0.0;
    }
    if (l16 == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;
    const bool isvar = l16 < NZT;           // lane owns a (real or dummy) move variable
    const bool isreal = l16 < nzr;          // ... a real one
    const bool iseps = NEPS && l16 == NZT;  // lane owns the slack variable
    const int iv = isvar ? l16 : 0;         // clamped index for loads
    const double2* vb2 = reinterpret_cast<const double2*>(vbuf);
    const double2* hrow = reinterpret_cast<const double2*>(sHv + iv * LDH);
    auto bcast = [&](double v, int j) -> double { return __shfl_sync(FULL, v, hb | j); };

    for (;;) {
        int pair = 0;
        if (lane == 0) pair = (int)atomicAdd(&P.counters[0], 1u);
        pair = __shfl_sync(FULL, pair, 0);
        if (2 * pair >= P.N) break;
        const int inst_raw = 2 * pair + (lane >> 4);
        const bool valid = inst_raw < P.N;
        const int inst = valid ? inst_raw : P.N - 1;

        // ---- stage 0: TMA bulk loads of this instance's matrices ----
        const int lv_ok = P.lv_ok[P.sH ? inst : 0];
        if (l16 == 0) {
            fence_proxy_async();
            const uint32_t bytes = (uint32_t)(L.nPdR + 2 * L.nHS) * 8u;
            mbar_arrive_expect_tx(bar, bytes);
            tma_bulk_g2s(sHv, Q.HvS + (long)inst * Q.sHS, (uint32_t)L.nHS * 8u, bar);
            tma_bulk_g2s(sLb, Q.LvS + (long)inst * Q.sHS, (uint32_t)L.nHS * 8u, bar);
            if (L.nPdR) tma_bulk_g2s(sPd, Q.PdR + (long)inst * Q.sPdR, (uint32_t)L.nPdR * 8u, bar);
        }
        // ---- stage 1: initpred! ----
        for (int k = l16; k < nx; k += 16) sxh[k] = P.xhat0[(long)inst * nx + k];
        for (int k = l16; k < nu; k += 16) slu[k] = P.lastu0[(long)inst * nu + k];
        if (nd > 0) {
            for (int k = l16; k < nd; k += 16) sd0[k] = P.d0[(long)inst * nd + k];
            for (int k = l16; k < nd * P.Hp; k += 16)
                sDh[k] = P.Dhat0 ? P.Dhat0[(long)inst * nd * P.Hp + k] : P.d0[(long)inst * nd + (k % nd)];
        }
        __syncwarp();
        const double* gK = P.K + (long)inst * P.sK;
        const double* gV = P.V + (long)inst * P.sV;
        const double* gB = P.B + (long)inst * P.sB;
        const double* gyop = P.yop + (long)inst * P.syop;
        const double* guop = P.uop + (long)inst * P.suop;
        const double* gM = P.Mw + (long)inst * P.sM;
        double racc = 0.0;
        for (int t = l16; t < nY; t += 16) {
            double f = gB[t];
            for (int k = 0; k < nx; ++k) f = fma(gK[t + (long)nY * k], sxh[k], f);
            for (int k = 0; k < nu; ++k) f = fma(gV[t + (long)nY * k], slu[k], f);
            if (nd > 0) {
                const double* gG = P.G + (long)inst * P.sG;
                const double* gJ = P.J + (long)inst * P.sJ;
                for (int k = 0; k < nd; ++k) f = fma(gG[t + (long)nY * k], sd0[k], f);
                for (int k = 0; k < nd * P.Hp; ++k) f = fma(gJ[t + (long)nY * k], sDh[k], f);
            }
            sF[t] = f;
            const double ryt = P.Rhat_y ? P.Rhat_y[(long)inst * nY + t] : P.ry[(long)inst * ny + (t % ny)];
            const double cy = f + gyop[t % ny] - ryt;
            const double ty = gM[t] * cy;
            stY[t] = ty;
            racc = fma(cy, ty, racc);
            if (valid) P.F_out[(long)inst * nY + t] = f;
        }
        if (P.has_terminal) {
            const double* gkx = P.kx + (long)inst * P.skx;
            const double* gvx = P.vx + (long)inst * P.svx;
            const double* gbx = P.bx + (long)inst * P.sbx;
            for (int i = l16; i < nx; i += 16) {
                double f = gbx[i];
                for (int k = 0; k < nx; ++k) f = fma(gkx[i + (long)nx * k], sxh[k], f);
                for (int k = 0; k < nu; ++k) f = fma(gvx[i + (long)nx * k], slu[k], f);
                if (nd > 0) {
                    const double* ggx = P.gx + (long)inst * P.sgx;
                    const double* gjx = P.jx + (long)inst * P.sjx;
                    for (int k = 0; k < nd; ++k) f = fma(ggx[i + (long)nx * k], sd0[k], f);
                    for (int k = 0; k < nd * P.Hp; ++k) f = fma(gjx[i + (long)nx * k], sDh[k], f);
                }
                sfx[i] = f;
            }
        }
        __syncwarp();
        mbar_wait(bar, phase);
        phase ^= 1u;
        // q_i = 2 sum_t Ev[t,i] tY[t]  (+ input-setpoint term)
        double q = 0.0;
        {
            double a0 = 0.0, a1 = 0.0;
            const int ir = isreal ? l16 : 0;
            if (P.pd_is_ev) {
                int t = 0;
                for (; t + 1 < nY; t += 2) {
                    a0 = fma(sPd[t * LDP + ir], stY[t], a0);
                    a1 = fma(sPd[(t + 1) * LDP + ir], stY[t + 1], a1);
                }
                if (t < nY) a0 = fma(sPd[t * LDP + ir], stY[t], a0);
            } else {
                const double* col = P.Ev + (long)inst * P.sEv + (long)nY * ir;
                for (int t = 0; t < nY; ++t) a0 = fma(col[t], stY[t], a0);
            }
            double a = a0 + a1;
            if (P.has_L) {
                const double* gL = P.Lw + (long)inst * P.sL;
                const int l = ir / nu, ch = ir % nu;
                for (int tt = P.blk_start[l]; tt < P.blk_start[l + 1]; ++tt) {
                    const int idx = tt * nu + ch;
                    const double ru = P.Rhat_u ? P.Rhat_u[(long)inst * P.nU + idx] : guop[ch];
                    a = fma(gL[idx], slu[ch] + guop[ch] - ru, a);
                }
                for (int idx = l16; idx < P.nU; idx += 16) {
                    const int c2 = idx % nu;
                    const double ru = P.Rhat_u ? P.Rhat_u[(long)inst * P.nU + idx] : guop[c2];
                    const double cu = slu[c2] + guop[c2] - ru;
                    racc = fma(gL[idx] * cu, cu, racc);
                }
            }
            q = isreal ? 2.0 * a : 0.0;
        }
        const double rconst = half_sum(racc);
        // ---- linconstraint!: right-hand sides of this lane's rows ----
        double hD[DS], sD[DS], lamD[DS], sigD[DS], cD[DS], hS[SS], sSp[SS], lamS[SS];
        double hmax = 0.0;
        const double* gdb = P.dbound + (long)inst * rt.nDr;
        const double* gsb = P.sbase + (long)inst * nS;
#pragma unroll
        for (int t = 0; t < DS; ++t) {
            const int k = l16 + 16 * t;
            hD[t] = 0.0;
            sigD[t] = 0.0;
            cD[t] = 0.0;
            if (k < nDb) {
                const int src = t_srcd[k];
                const double fsrc = src < nY ? sF[src] : sfx[src - nY];
                sigD[t] = t_sigd[k];
                cD[t] = t_cd[k];
                hD[t] = sigD[t] * (gdb[k] - fsrc);
                hmax = fmax(hmax, fabs(hD[t]));
            }
        }
#pragma unroll
        for (int t = 0; t < SS; ++t) {
            const int r = l16 + 16 * t;
            hS[t] = 0.0;
            if (r < nS) {
                const int ch = t_ch[r];
                hS[t] = gsb[r] - (ch >= 0 ? t_sigs[r] * slu[ch] : 0.0);
                hmax = fmax(hmax, fabs(hS[t]));
            }
        }
        const double hscale = 1.0 + half_max(hmax);
        const double qs = 1.0 + half_max(fabs(q));
        const double Hee = NEPS ? P.Hee[P.sH ? inst : 0] : 0.0;

        // helpers -------------------------------------------------------------------------
        // dense base products yb[t] = Pd[k,:] v and sparse products gs[t] = v[i1]-v[i2]; v is in vbuf
        auto row_products = [&](double (&ybv)[DS], double (&gsv)[SS]) {
            const double2* rows[DS];
#pragma unroll
            for (int t = 0; t < DS; ++t) {
                ybv[t] = 0.0;
                rows[t] = reinterpret_cast<const double2*>(sPd + max(min(l16 + 16 * t, nDbp - 1), 0) * LDP);
            }
#pragma unroll
            for (int jj = 0; jj < NV2; ++jj) {
                const double2 v2 = vb2[jj];
#pragma unroll
                for (int t = 0; t < DS; ++t) {
                    const double2 p2 = rows[t][jj];
                    ybv[t] = fma(p2.x, v2.x, ybv[t]);
                    ybv[t] = fma(p2.y, v2.y, ybv[t]);
                }
            }
#pragma unroll
            for (int t = 0; t < SS; ++t) {
                const int r = l16 + 16 * t;
                gsv[t] = 0.0;
                if (r < nS) {
                    const int i2 = t_i2[r];
                    gsv[t] = vbuf[t_i1[r]] - (i2 >= 0 ? vbuf[i2] : 0.0);
                }
            }
        };
        auto gSf = [&](int t, double gst, double epsv) -> double {
            const int r = l16 + 16 * t;
            return r < nsr ? t_sigs[r] * gst - t_cs[r] * epsv : 0.0;
        };
        // out_i = (G'w)_i for lane i (variables and the slack lane); w given per slot
        auto gt_apply = [&](const double (&wDv)[DS], const double (&wSv)[SS]) -> double {
            double ce = 0.0;
#pragma unroll
            for (int t = 0; t < DS; ++t) {
                const int k = l16 + 16 * t;
                if (k < nDb) wd[k] = sigD[t] * wDv[t];
                ce = fma(cD[t], wDv[t], ce);
            }
#pragma unroll
            for (int t = 0; t < SS; ++t) {
                const int r = l16 + 16 * t;
                if (r < nsr) {
                    ws[r] = t_sigs[r] * wSv[t];
                    ce = fma(t_cs[r], wSv[t], ce);
                }
            }
            __syncwarp();
            double a0 = 0.0, a1 = 0.0;
            const double2* w2 = reinterpret_cast<const double2*>(wd);
            const double* col = sPd + iv;
#pragma unroll 4
            for (int k = 0; k < nDbp; k += 2) {
                const double2 w = w2[k >> 1];
                a0 = fma(col[k * LDP], w.x, a0);
                a1 = fma(col[(k + 1) * LDP], w.y, a1);
            }
            double acc = a0 + a1;
            for (int e = t_vptr[iv]; e < t_vptr[iv + 1]; ++e) acc = fma((double)t_vsgn[e], ws[t_vrow[e]], acc);
            ce = half_sum(ce);
            __syncwarp();
            return isvar ? acc : (iseps ? -ce : 0.0);
        };
        // (H x)_i with x in vbuf (variables) and eps
        auto hess_apply = [&](double epsv) -> double {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int jj = 0; jj < NV2; ++jj) {
                const double2 h2 = hrow[jj];
                const double2 v2 = vb2[jj];
                a = fma(h2.x, v2.x, a);
                b = fma(h2.y, v2.y, b);
            }
            return isvar ? a + b : (iseps ? Hee * epsv : 0.0);
        };

        // ---- stage 2: unconstrained minimiser with the cached factor (Lv in sLb, diagonal holds 1/L_ii) ----
        double x = 0.0;
        if (lv_ok) {
            double b = isvar ? -q : 0.0;
            const double invd0 = sLb[iv * LDH + iv];
#pragma unroll
            for (int j = 0; j < NZT; ++j) {
                const double yj = bcast(b * invd0, j);
                if (l16 > j && isvar) b = fma(-sLb[iv * LDH + j], yj, b);
            }
            b *= invd0;
#pragma unroll
            for (int j = NZT - 1; j >= 0; --j) {
                const double xj = bcast(b * invd0, j);
                if (l16 < j) b = fma(-sLb[j * LDH + iv], xj, b);
            }
            x = isvar ? b * invd0 : 0.0;
        }
        __syncwarp();
        vbuf[l16] = x;
        __syncwarp();
        double yb[DS], gs[SS];
        row_products(yb, gs);
        double eps = 0.0;
        double smin = 1e300;
#pragma unroll
        for (int t = 0; t < DS; ++t) {
            sD[t] = hD[t] - sigD[t] * yb[t];
            if (l16 + 16 * t < nDb) smin = fmin(smin, sD[t]);
        }
#pragma unroll
        for (int t = 0; t < SS; ++t) {
            sSp[t] = hS[t] - gSf(t, gs[t], eps);
            if (l16 + 16 * t < nsr) smin = fmin(smin, sSp[t]);
        }
        smin = (m > 0) ? half_min(smin) : 0.0;
        const bool feasible = (m == 0) || (lv_ok && smin >= -1e-12 * hscale);
        int status = ST_OPTIMAL, iters = 0;
        bool active = valid && !feasible;
        if (__any_sync(FULL, active)) {
            // ---- stage 3: Mehrotra predictor-corrector, both instances of the warp in lock-step ----
            const double minv = 1.0 / (double)max(m, 1);
            const double mu0 = fmax(1e-2 * qs * hscale * minv, 1e-8);
#pragma unroll
            for (int t = 0; t < DS; ++t) {
                sD[t] = fmax(sD[t], 1e-2 * hscale);
                lamD[t] = (l16 + 16 * t < nDb) ? mu0 / sD[t] : 0.0;
            }
#pragma unroll
            for (int t = 0; t < SS; ++t) {
                sSp[t] = fmax(sSp[t], 1e-2 * hscale);
                lamS[t] = (l16 + 16 * t < nsr) ? mu0 / sSp[t] : 0.0;
            }
            // ---- warm start (set_warmstart_mpc! analogue): previous period's Z̃ and multipliers ----
            const bool warm = P.use_ws && active && P.ws_flag[inst] != 0;
            if (__any_sync(FULL, warm)) {
                double xw = x;
                if (warm) {
                    const double* gZp = P.Z + (long)inst * nr;
                    double a = 0.0;
                    if (isreal)
                        for (int l = l16 % nu; l <= l16; l += nu) a += gZp[l];
                    xw = isreal ? a : ((iseps && NEPS) ? gZp[nzr] : 0.0);
                }
                __syncwarp();
                vbuf[l16] = xw;
                __syncwarp();
                double ybw[DS], gsw[SS];
                row_products(ybw, gsw);
                const double epsw = NEPS ? bcast(xw, NZT) : 0.0;
                if (warm) {
                    const double* glw = P.lam_ws + (long)inst * P.ws_stride;
                    const double lmin = 1e-4 * qs / hscale;
                    x = xw;
                    eps = epsw;
#pragma unroll
                    for (int t = 0; t < DS; ++t) {
                        const int k = l16 + 16 * t;
                        yb[t] = ybw[t];
                        sD[t] = fmax(hD[t] - (sigD[t] * ybw[t] - cD[t] * epsw), 1e-2 * hscale);
                        lamD[t] = (k < nDb) ? fmax(glw[k], lmin) : 0.0;
                    }
#pragma unroll
                    for (int t = 0; t < SS; ++t) {
                        const int r = l16 + 16 * t;
                        gs[t] = gsw[t];
                        sSp[t] = fmax(hS[t] - gSf(t, gsw[t], epsw), 1e-2 * hscale);
                        lamS[t] = (r < nsr) ? fmax(glw[nDb + r], lmin) : 0.0;
                    }
                }
                __syncwarp();
                vbuf[l16] = x;
                __syncwarp();
            }
            if (active) status = ST_ITERATION_LIMIT;
            double best_merit = 1e300, rp_inf = 0.0;
            // primal residual r_p = Gx + s - h.  Because ds = -r_p - G dx is formed from the COMPUTED dx, the
            // update s += a ds keeps r_p <- (1 - a) r_p exactly (up to rounding), whatever the accuracy of the
            // linear solve: r_p is carried by that recurrence instead of being re-evaluated every iteration.
            double rpD[DS], rpS[SS];
#pragma unroll
            for (int t = 0; t < DS; ++t)
                rpD[t] = (l16 + 16 * t < nDb) ? sigD[t] * yb[t] - cD[t] * eps + sD[t] - hD[t] : 0.0;
#pragma unroll
            for (int t = 0; t < SS; ++t)
                rpS[t] = (l16 + 16 * t < nsr) ? gSf(t, gs[t], eps) + sSp[t] - hS[t] : 0.0;
            for (int it = 0; it <= P.max_iter; ++it) {
                // residuals: rd = Hx + q + G'lam, rp = Gx + s - h
                const double Hx = hess_apply(eps);
                const double Gtl = gt_apply(lamD, lamS);
                const double rd = Hx + q + Gtl;
                double e_p = 0.0, musum = 0.0;
#pragma unroll
                for (int t = 0; t < DS; ++t) {
                    e_p = fmax(e_p, fabs(rpD[t]));
                    musum = fma(sD[t], lamD[t], musum);
                }
#pragma unroll
                for (int t = 0; t < SS; ++t) {
                    e_p = fmax(e_p, fabs(rpS[t]));
                    musum = fma(sSp[t], lamS[t], musum);
                }
                const double e_d = half_max(fabs(rd));
                const double qd = qs + half_max(fmax(fabs(Hx + q), fabs(Gtl)));  // scale of the dual residual's terms
                e_p = half_max(e_p);
                const double mu = half_sum(musum) * minv;
                if (active) {
                    rp_inf = e_p;
                    if (!(e_d == e_d) || !(e_p == e_p) || !(mu == mu) || e_d > 1e250 || e_p > 1e250) {
                        status = ST_INFEASIBLE;
                        active = false;
                    } else {
                        const double merit = fmax(fmax(e_d / (P.tol * qd), e_p / (P.tol * hscale)),
                                                  mu * (double)m / (P.tol_mu * qs * hscale));
                        if (merit <= 1.0 || (best_merit <= 1e3 && merit >= best_merit) ||
                            (it == P.max_iter && merit <= 1e3)) {
                            status = ST_OPTIMAL;
                            active = false;
                        }
                        best_merit = fmin(best_merit, merit);
                    }
                }
                if (it == P.max_iter || !__any_sync(FULL, active)) break;
                if (active) iters = it + 1;
                // ---- Phi = H + G' D G in registers (lane i = row i) ----
                // one reciprocal of s and of lambda per row and iteration; everything else is multiplies
                double dD[DS], dS[SS], isD[DS], isS[SS];
                double cc = 0.0;
#pragma unroll
                for (int t = 0; t < DS; ++t) {
                    const int k = l16 + 16 * t;
                    const bool ok = k < nDb;
                    isD[t] = ok ? __drcp_rn(sD[t]) : 0.0;
                    dD[t] = lamD[t] * isD[t];
                    if (ok) wp[k] = make_double2(dD[t], sigD[t] * cD[t] * dD[t]);
                    cc = fma(cD[t] * cD[t], dD[t], cc);
                }
#pragma unroll
                for (int t = 0; t < SS; ++t) {
                    const int r = l16 + 16 * t;
                    const bool ok = r < nsr;
                    isS[t] = ok ? __drcp_rn(sSp[t]) : 0.0;
                    dS[t] = lamS[t] * isS[t];
                    if (ok) {
                        ws[r] = dS[t];
                        cc = fma(t_cs[r] * t_cs[r], dS[t], cc);
                    }
                }
                cc = half_sum(cc);
                __syncwarp();
                double phi[2 * NV2 + 2];
                double pdiag = isvar ? sHv[iv * LDH + iv] : 0.0;
#pragma unroll
                for (int jj = 0; jj < NV2; ++jj) {
                    const double2 h2 = hrow[jj];
                    phi[2 * jj] = isvar ? h2.x : 0.0;
                    phi[2 * jj + 1] = isvar ? h2.y : 0.0;
                }
                phi[2 * NV2] = 0.0;
                double border = 0.0;
                {
                    const double* col = sPd + iv;
#pragma unroll 2
                    for (int k = 0; k < nDb; ++k) {
                        const double2 w = wp[k];
                        const double pik = col[k * LDP];
                        const double tk = isvar ? w.x * pik : 0.0;
                        border = fma(pik, w.y, border);
                        pdiag = fma(tk, pik, pdiag);
                        const double2* prow = reinterpret_cast<const double2*>(sPd + k * LDP);
#pragma unroll
                        for (int jj = 0; jj < NV2; ++jj) {
                            const double2 p2 = prow[jj];
                            phi[2 * jj] = fma(tk, p2.x, phi[2 * jj]);
                            phi[2 * jj + 1] = fma(tk, p2.y, phi[2 * jj + 1]);
                        }
                    }
                }
                // 1-/2-variable rows
                for (int e = t_vptr[iv]; e < t_vptr[iv + 1]; ++e) {
                    const int r = t_vrow[e];
                    const double dr = isvar ? ws[r] : 0.0;
                    pdiag += dr;
                    border = fma((double)t_vsgn[e] * t_sigs[r] * t_cs[r], dr, border);
                    if (Q.has_pair_rows && t_vsgn[e] > 0) {
                        const int i2 = t_i2[r];
#pragma unroll
                        for (int j = 0; j < NZT; ++j)
                            if (j == i2) phi[j] -= dr;
                    }
                }
                if (NEPS) {
                    bb[l16] = isvar ? border : 0.0;
                    __syncwarp();
                    if (iseps) {
#pragma unroll
                        for (int j = 0; j < NZT; ++j) phi[j] = -bb[j];
                        pdiag = Hee + cc;
                    }
                }
                // ---- Cholesky: right-looking, rows in registers, columns exchanged by shuffles ----
                double invd = 1.0;
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    double dk = bcast(pdiag, k);
                    if (!(dk > 1e-280)) dk = 1e200;
                    const double rs = rsqrt(dk);
                    const double lik = (l16 > k) ? phi[k] * rs : 0.0;  // column k of L
                    if (l16 == k) invd = rs;
                    phi[k] = lik;
                    pdiag = fma(-lik, lik, pdiag);
#pragma unroll
                    for (int j = k + 1; j < NT; ++j) {
                        const double ljk = bcast(lik, j);
                        phi[j] = fma(-lik, ljk, phi[j]);
                    }
                }
                // rows of L to shared memory for the backward substitutions
                __syncwarp();
                if (l16 < NT) {
#pragma unroll
                    for (int j = 0; j < NT; ++j) sLb[l16 * LDN + j] = phi[j];
                }
                __syncwarp();
                auto solve = [&](double b) -> double {
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        const double yj = bcast(b * invd, j);
                        b = fma(-((l16 > j) ? phi[j] : 0.0), yj, b);
                    }
                    b *= invd;
#pragma unroll
                    for (int j = NT - 1; j >= 0; --j) {
                        const double xj = bcast(b * invd, j);
                        const double lji = sLb[j * LDN + l16];
                        b = fma(-((l16 < j) ? lji : 0.0), xj, b);
                    }
                    return (l16 < NT) ? b * invd : 0.0;
                };
                // ---- predictor (pass 0) and corrector (pass 1) share one copy of the code ----
                double rcD[DS], rcS[SS], dsD[DS], dlD[DS], dsS[SS], dlS[SS], ybd[DS], gsd[SS];
#pragma unroll
                for (int t = 0; t < DS; ++t) rcD[t] = 0.0;
#pragma unroll
                for (int t = 0; t < SS; ++t) rcS[t] = 0.0;
                double dx = 0.0, a = 1.0;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    double wDv[DS], wSv[SS];
#pragma unroll
                    for (int t = 0; t < DS; ++t)
                        wDv[t] = pass ? (lamD[t] * rpD[t] - rcD[t]) * isD[t] : dD[t] * rpD[t] - lamD[t];
#pragma unroll
                    for (int t = 0; t < SS; ++t)
                        wSv[t] = pass ? (lamS[t] * rpS[t] - rcS[t]) * isS[t] : dS[t] * rpS[t] - lamS[t];
                    dx = solve(-(rd + gt_apply(wDv, wSv)));
                    __syncwarp();
                    vbuf[l16] = dx;
                    __syncwarp();
                    const double deps = NEPS ? bcast(dx, NZT) : 0.0;
                    row_products(ybd, gsd);
                    double rho = 0.0;  // max over rows of -ds/s and -dl/lambda  (step to the boundary = 1/rho)
#pragma unroll
                    for (int t = 0; t < DS; ++t) {
                        dsD[t] = -rpD[t] - (sigD[t] * ybd[t] - cD[t] * deps);
                        dlD[t] = pass ? -(rcD[t] + lamD[t] * dsD[t]) * isD[t] : -lamD[t] - dD[t] * dsD[t];
                        // -dl/lambda: predictor 1 + ds/s (since d/lambda = 1/s); corrector needs 1/lambda
                        const double rl = pass ? ((l16 + 16 * t < nDb) ? -dlD[t] * __drcp_rn(lamD[t]) : 0.0) : fma(dsD[t], isD[t], 1.0);
                        rho = fmax(rho, fmax(-dsD[t] * isD[t], rl));
                    }
#pragma unroll
                    for (int t = 0; t < SS; ++t) {
                        dsS[t] = -rpS[t] - gSf(t, gsd[t], deps);
                        dlS[t] = pass ? -(rcS[t] + lamS[t] * dsS[t]) * isS[t] : -lamS[t] - dS[t] * dsS[t];
                        const double rl = pass ? ((l16 + 16 * t < nsr) ? -dlS[t] * __drcp_rn(lamS[t]) : 0.0) : fma(dsS[t], isS[t], 1.0);
                        rho = fmax(rho, fmax(-dsS[t] * isS[t], rl));
                    }
                    rho = half_max(rho);
                    if (pass == 0) {
                        const double a_aff = rho > 1.0 ? 1.0 / rho : 1.0;
                        double mua = 0.0;
#pragma unroll
                        for (int t = 0; t < DS; ++t) mua = fma(sD[t] + a_aff * dsD[t], lamD[t] + a_aff * dlD[t], mua);
#pragma unroll
                        for (int t = 0; t < SS; ++t) mua = fma(sSp[t] + a_aff * dsS[t], lamS[t] + a_aff * dlS[t], mua);
                        mua = half_sum(mua) * minv;
                        double sig = mua / mu;
                        sig = sig * sig * sig;
#pragma unroll
                        for (int t = 0; t < DS; ++t) rcD[t] = sD[t] * lamD[t] + dsD[t] * dlD[t] - sig * mu;
#pragma unroll
                        for (int t = 0; t < SS; ++t) rcS[t] = sSp[t] * lamS[t] + dsS[t] * dlS[t] - sig * mu;
                    } else {
                        a = rho > 0.99 ? 0.99 / rho : 1.0;
                    }
                }
                if (!active) a = 0.0;  // finished instance: frozen
                x = fma(a, dx, x);
                eps = NEPS ? bcast(x, NZT) : 0.0;
                const double oma = 1.0 - a;
#pragma unroll
                for (int t = 0; t < DS; ++t) {
                    rpD[t] *= oma;
                    sD[t] = fma(a, dsD[t], sD[t]);
                    lamD[t] = fma(a, dlD[t], lamD[t]);
                }
#pragma unroll
                for (int t = 0; t < SS; ++t) {
                    rpS[t] *= oma;
                    sSp[t] = fma(a, dsS[t], sSp[t]);
                    lamS[t] = fma(a, dlS[t], lamS[t]);
                }
                __syncwarp();
                vbuf[l16] = x;
                __syncwarp();
            }
            if (status == ST_ITERATION_LIMIT && rp_inf > 1e-6 * hscale) status = ST_INFEASIBLE;
        }
        // ---- stage 4: getinput! ----
        double* gZ = P.Z + (long)inst * nr;
        __syncwarp();
        if (status == ST_INFEASIBLE) {
            // shifted previous solution (set_warmstart_mpc!), converted to level coordinates
            double a = 0.0;
            if (isreal)
                for (int l = l16 % nu; l <= l16; l += nu) a += (l + nu < nzr) ? gZ[l + nu] : 0.0;
            x = isreal ? a : (iseps ? gZ[nzr] : 0.0);
        }
        __syncwarp();
        vbuf[l16] = x;
        __syncwarp();
        eps = NEPS ? bcast(x, NZT) : 0.0;
        const double Hx = hess_apply(eps);
        double jacc = (l16 < NT) ? x * (0.5 * Hx + q) : 0.0;
        jacc = half_sum(jacc) + rconst;
        // q̃ in reference coordinates: q̃[j] = sum_{l' >= l(j), same input} q_v[l']
        bb[l16] = q;
        __syncwarp();
        if (valid) {
            if (isreal) {
                gZ[l16] = x - (l16 >= nu ? vbuf[l16 - nu] : 0.0);
                double a = 0.0;
                for (int l = l16; l < nzr; l += nu) a += bb[l];
                P.qt_out[(long)inst * nr + l16] = a;
            }
            if (iseps) {
                gZ[nzr] = x;
                P.qt_out[(long)inst * nr + nzr] = 0.0;
            }
            if (l16 < nu) {
                const double du = vbuf[l16];
                const double lu = slu[l16];
                P.lastu_prev[(long)inst * nu + l16] = lu;
                P.lastu0[(long)inst * nu + l16] = lu + du;
                P.u[(long)inst * nu + l16] = lu + du + guop[l16];
            }
            if (P.use_ws) {
                // multipliers for the next period's warm start (valid only after a converged IPM solve)
                const bool keep = status == ST_OPTIMAL && iters > 0;
                double* glw = P.lam_ws + (long)inst * P.ws_stride;
                if (keep) {
#pragma unroll
                    for (int t = 0; t < DS; ++t)
                        if (l16 + 16 * t < nDb) glw[l16 + 16 * t] = lamD[t];
#pragma unroll
                    for (int t = 0; t < SS; ++t)
                        if (l16 + 16 * t < nsr) glw[nDb + l16 + 16 * t] = lamS[t];
                }
                if (l16 == 0) P.ws_flag[inst] = keep ? 1 : 0;
            }
            if (l16 == 0) {
                P.r_out[inst] = rconst;
                if (P.J_out) P.J_out[inst] = jacc;
                P.status[inst] = status;
                P.iters[inst] = iters;
            }
        }
        fence_proxy_async();
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&P.counters[1], 1u);
        if (done == gridDim.x - 1) {
            P.counters[0] = 0u;
            P.counters[1] = 0u;
            __threadfence();
        }
    }
}

}  // namespace bmpc
