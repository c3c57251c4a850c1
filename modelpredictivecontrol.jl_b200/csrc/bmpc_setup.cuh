// One-off / on-demand kernels around the step: change of coordinates of the host-supplied
// prediction matrices (route B of bmpc.h), cached Cholesky of the Hessian, gather of the
// dense constraint rows, and the getinfo predictions.
#pragma once
#include <cuda_runtime.h>

#include "bmpc_device.cuh"

namespace bmpc {

// out[:, j] = in[:, j] - in[:, j+nu]   (X_v = X * D,  DU_l = v_l - v_{l-1});  column-major [rows x nz]
// in has per-instance stride rows*nz, out has stride ostride (even, TMA alignment).
__global__ void k_level_cols(const double* __restrict__ in, double* __restrict__ out, int rows, int nz, int nu,
                             long ostride, long tot) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= tot) return;
    const long per = (long)rows * nz;
    const long inst = e / per;
    const int rem = (int)(e - inst * per);
    const int j = rem / rows;
    const double nxt = (j + nu < nz) ? in[e + (long)rows * nu] : 0.0;
    out[inst * ostride + rem] = in[e] - nxt;
}

// Hv = D' H̃[0:nz,0:nz] D, packed row-major lower (p = i(i+1)/2 + j); Hee = H̃[n-1,n-1] when neps.
// H̃ is column-major n x n with the LOWER triangle authoritative (Hermitian(:L), construct.jl:842).
__global__ void k_level_hess(const double* __restrict__ Ht, double* __restrict__ Hv, double* __restrict__ Hee, int n,
                             int nz, int nu, int neps, int nHp2, int npair, long tot) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= tot) return;
    const long inst = e / npair;
    const int p = (int)(e - inst * npair);
    int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= p) ++i;
    while (i * (i + 1) / 2 > p) --i;
    const int j = p - i * (i + 1) / 2;
    const double* H = Ht + inst * (long)n * n;
    auto hs = [&](int a, int b) -> double {
        if (a >= nz || b >= nz) return 0.0;
        const int hi = a > b ? a : b, lo = a > b ? b : a;
        return H[hi + (long)n * lo];
    };
    Hv[inst * nHp2 + p] = hs(i, j) - hs(i + nu, j) - hs(i, j + nu) + hs(i + nu, j + nu);
    if (p == 0) Hee[inst] = neps ? H[(long)(n - 1) + (long)n * (n - 1)] : 0.0;
}

// Lv = chol(Hv), one thread per instance (one-off); ok = 0 if Hv is not safely positive definite.
__global__ void k_chol_serial(const double* __restrict__ Hv, double* __restrict__ Lv, int* __restrict__ ok, int nz,
                              int nHp2, int NM) {
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= NM) return;
    const double* A = Hv + (long)inst * nHp2;
    double* L = Lv + (long)inst * nHp2;
    double dmax = 0.0;
    for (int i = 0; i < nz; ++i) dmax = fmax(dmax, fabs(A[i * (i + 1) / 2 + i]));
    int good = 1;
    for (int i = 0; i < nz; ++i) {
        double* ri = L + i * (i + 1) / 2;
        const double* ai = A + i * (i + 1) / 2;
        for (int j = 0; j <= i; ++j) {
            const double* rj = L + j * (j + 1) / 2;
            double a = ai[j];
            for (int p = 0; p < j; ++p) a = fma(-ri[p], rj[p], a);
            if (j == i) {
                if (!(a > 1e-13 * dmax)) {
                    good = 0;
                    a = 1.0;
                }
                ri[j] = sqrt(a);
            } else {
                ri[j] = a / rj[j];
            }
        }
    }
    ok[inst] = good;
}

// Pd[k, j] = src_k < nY ? Ev[src_k, j] : (src_k < nY + nx ? exv[src_k - nY, j] : Ewv[src_k - nY - nx, j])
__global__ void k_gather_pd(const double* __restrict__ Ev, const double* __restrict__ exv, const double* __restrict__ Ewv,
                            double* __restrict__ Pd, const int* __restrict__ pd_src, int nY, int nx, int nFw, int nz, int nDb,
                            long sEv, int nPd2, long tot) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= tot) return;
    const long per = (long)nDb * nz;
    const long inst = e / per;
    const int rem = (int)(e - inst * per);
    const int j = rem / nDb, k = rem - j * nDb;
    const int src = pd_src[k];
    double v;
    if (src < nY)
        v = Ev[inst * sEv + src + (long)nY * j];
    else if (src < nY + nx)
        v = exv[inst * (long)nx * nz + (src - nY) + (long)nx * j];
    else
        v = Ewv[inst * (long)nFw * nz + (src - nY - nx) + (long)nFw * j];
    Pd[inst * nPd2 + rem] = v;
}

// X̂0 block of the MultipleShooting decision vector Z = [ΔU; X̂0] (src/controller/transcription.jl:28-60): the states the
// defect equations ES Z + FS = 0 (init_defectmat :373-414, linconstrainteq! :913-928) tie to the inputs --
//   x̂0(k+j+1) = Â x̂0(k+j) + B̂u u0(k+j) + B̂d d̂0(k+j) + (f̂op - x̂op),  u0(k+j) = u0(k-1) + sum of the moves up to block(j)
// -- evaluated from the optimal ΔU of the last step.  One CTA per instance, one thread per state, Hp sequential steps.
__global__ void k_ms_states(const double* __restrict__ A, long sA, const double* __restrict__ Bu, long sBu,
                            const double* __restrict__ Bd, long sBd, const double* __restrict__ f, long sf,
                            const double* __restrict__ Z, const double* __restrict__ xhat0, const double* __restrict__ lastu_prev,
                            const double* __restrict__ d0, const double* __restrict__ Dhat0, const int* __restrict__ blk_of_t,
                            double* __restrict__ X0, int nx, int nu, int nd, int Hp, int n) {
    extern __shared__ double sm[];
    double* xa = sm;
    double* xb = sm + nx;
    double* u = sm + 2 * nx;
    const long inst = blockIdx.x;
    const int i = threadIdx.x;
    const double* Ai = A + inst * sA;
    const double* Bi = Bu + inst * sBu;
    const double* Di = Bd + inst * sBd;
    const double* z = Z + inst * n;
    if (i < nx) xa[i] = xhat0[inst * nx + i];
    if (i < nu) u[i] = lastu_prev[inst * nu + i];
    int blk_done = -1;
    __syncthreads();
    for (int j = 0; j < Hp; ++j) {
        const int bl = blk_of_t[j];
        if (bl != blk_done) {  // a new move block starts at step j: u0 += Δu_block
            if (i < nu) u[i] += z[bl * nu + i];
            blk_done = bl;
            __syncthreads();
        }
        if (i < nx) {
            double a = f[inst * sf + i];
            for (int k = 0; k < nx; ++k) a = fma(Ai[i + (long)nx * k], xa[k], a);
            for (int c = 0; c < nu; ++c) a = fma(Bi[i + (long)nx * c], u[c], a);
            for (int e = 0; e < nd; ++e) {
                const double dv = j == 0 ? d0[inst * nd + e] : (Dhat0 ? Dhat0[inst * nd * Hp + (j - 1) * nd + e] : d0[inst * nd + e]);
                a = fma(Di[i + (long)nx * e], dv, a);
            }
            xb[i] = a;
            X0[inst * (long)nx * Hp + (long)j * nx + i] = a;
        }
        __syncthreads();
        double* t = xa;
        xa = xb;
        xb = t;
    }
}

// relaxW (construct.jl:1138-1160) in input-level coordinates: Ewv = W̄y [0; Ev] + W̄u [Pu; pu] D, column-major [nFw x nz]:
// row (t, j): sum_o Wy[j,o] Ev[(t-1) ny + o, :] (t >= 1)  +  Wu[j,c] on the column of input c's level at step min(t, Hp-1)
__global__ void k_build_ew(const double* __restrict__ Ev, long sEv, const double* __restrict__ Wc, long sW,
                           const int* __restrict__ blk_of_t, double* __restrict__ Ewv, int nw, int ny, int nu, int nd, int Hp,
                           int nY, int nz, long tot) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= tot) return;
    const int nFw = nw * (Hp + 1);
    const long per = (long)nFw * nz;
    const long inst = e / per;
    const int rem = (int)(e - inst * per);
    const int i = rem / nFw, r = rem - i * nFw;
    const int t = r / nw, j = r - t * nw;
    const double* Wy = Wc + inst * sW;
    const double* Wu = Wy + nw * ny;
    double a = 0.0;
    if (t >= 1)
        for (int o = 0; o < ny; ++o) a = fma(Wy[j + nw * o], Ev[inst * sEv + (t - 1) * ny + o + (long)nY * i], a);
    const int tc = t < Hp ? t : Hp - 1;
    const int l = i / nu, c = i - l * nu;
    if (l == blk_of_t[tc]) a += Wu[j + nw * c];
    Ewv[inst * per + rem] = a;
}

// getinfo predictions (predict!, transcription.jl:1136-1145; getU0!, :1115) evaluated in level
// coordinates: v = cumsum(ΔU) per input, Ŷ0 = Ev v + F, U0 = v[block(t)] + lastu0, x̂0end = exv v + fx̂.
// One CTA per instance.  fx̂ = bx̂ + kx̂ x̂0 + vx̂ u0(k-1) + gx̂ d0 + jx̂ D̂0 (linconstraint!, transcription.jl:818-822;
// D̂0 = NULL means d0 repeated over Hp, as moveinput! does by default, execute.jl:64).
__global__ void k_getinfo(const double* __restrict__ Ev, long sE, const double* __restrict__ exv, long sex,
                          const double* __restrict__ kx, long skx, const double* __restrict__ vx, long svx,
                          const double* __restrict__ bx, long sbx, const double* __restrict__ Z,
                          const double* __restrict__ F, const double* __restrict__ xhat0,
                          const double* __restrict__ lastu_prev, const int* __restrict__ blk_of_t,
                          double* __restrict__ Yhat0, double* __restrict__ U0, double* __restrict__ xend, int nY, int nz,
                          int n, int nu, int nx, int Hp, int nd, const double* __restrict__ gx, long sgx,
                          const double* __restrict__ jx, long sjx, const double* __restrict__ d0,
                          const double* __restrict__ Dhat0) {
    extern __shared__ double v[];
    const int inst = blockIdx.x;
    const double* z = Z + (long)inst * n;
    for (int j = threadIdx.x; j < nz; j += blockDim.x) {
        double a = 0.0;
        for (int l = j % nu; l <= j; l += nu) a += z[l];
        v[j] = a;
    }
    __syncthreads();
    const double* e = Ev + inst * sE;
    for (int t = threadIdx.x; t < nY; t += blockDim.x) {
        double a = F[(long)inst * nY + t];
        for (int j = 0; j < nz; ++j) a = fma(e[t + (long)nY * j], v[j], a);
        Yhat0[(long)inst * nY + t] = a;
    }
    for (int k = threadIdx.x; k < nu * Hp; k += blockDim.x) {
        const int t = k / nu, ch = k % nu;
        U0[(long)inst * nu * Hp + k] = lastu_prev[(long)inst * nu + ch] + v[blk_of_t[t] * nu + ch];
    }
    if (exv) {
        const double* exi = exv + inst * sex;
        const double* kxi = kx + inst * skx;
        const double* vxi = vx + inst * svx;
        const double* bxi = bx + inst * sbx;
        for (int i = threadIdx.x; i < nx; i += blockDim.x) {
            double a = bxi[i];
            for (int k = 0; k < nx; ++k) a = fma(kxi[i + (long)nx * k], xhat0[(long)inst * nx + k], a);
            for (int k = 0; k < nu; ++k) a = fma(vxi[i + (long)nx * k], lastu_prev[(long)inst * nu + k], a);
            for (int j = 0; j < nz; ++j) a = fma(exi[i + (long)nx * j], v[j], a);
            if (nd > 0 && d0) {
                const double* gxi = gx + inst * sgx;
                const double* jxi = jx + inst * sjx;
                const double* di = d0 + (long)inst * nd;
                for (int k = 0; k < nd; ++k) a = fma(gxi[i + (long)nx * k], di[k], a);
                for (int k = 0; k < nd * Hp; ++k)
                    a = fma(jxi[i + (long)nx * k], Dhat0 ? Dhat0[(long)inst * nd * Hp + k] : di[k % nd], a);
            }
            xend[(long)inst * nx + i] = a;
        }
    }
}

// Setup: dense row matrix Gt and extended Hessian / factor for the warp kernel (one CTA per instance).
//   Gt[p, :] = [sigma_r * p_r, -c_r] for the row r = pos_row[p] of dense position p < mD (zero rows up to GR):
//   sparse rows p_r = e_i1 - e_i2, dense rows p_r = Pd[base_r, :]  (the unit rows never enter the matrix)
//   H_ext = blockdiag(Hv, Hee, I_dummy);  L_ext = its Cholesky factor with 1/L_ii on the diagonal
__global__ void k_make_warp(const double* __restrict__ Pd, long sPd, int nDb, const double* __restrict__ Hv,
                            const double* __restrict__ Lv, int nHp2, const double* __restrict__ Hee, RowTables rt,
                            const int* __restrict__ pos_row, int mD, int nz, int neps, int NT, int LDG, int LDH, int GR,
                            double* __restrict__ Gw, long sGw, double* __restrict__ HL, long sHL) {
    const long inst = blockIdx.x;
    double* G = Gw + inst * sGw;
    for (int e = threadIdx.x; e < GR * LDG; e += blockDim.x) {
        const int p = e / LDG, j = e % LDG;
        double v = 0.0;
        if (p < mD) {
            const int r = pos_row[p];
            const double sg = rt.row_sig[r], c = rt.row_c[r];
            if (j < nz) {
                if (r < rt.nS)
                    v = (j == rt.s_i1[r]) ? sg : ((j == rt.s_i2[r]) ? -sg : 0.0);
                else
                    v = sg * Pd[inst * sPd + rt.dr_base[r - rt.nS] + (long)nDb * j];
            } else if (neps && j == nz) {
                v = -c;
            }
        }
        G[e] = v;
    }
    double* H = HL + inst * sHL;
    double* L = H + NT * LDH;
    const int n = nz + neps;
    for (int e = threadIdx.x; e < NT * LDH; e += blockDim.x) {
        const int i = e / LDH, j = e % LDH;
        double h = 0.0, l = 0.0;
        if (j < NT) {
            if (i < nz && j < nz) {
                const int hi = i > j ? i : j, lo = i > j ? j : i;
                h = Hv[inst * nHp2 + hi * (hi + 1) / 2 + lo];
                if (j < i) l = Lv[inst * nHp2 + i * (i + 1) / 2 + j];
                if (j == i) l = 1.0 / Lv[inst * nHp2 + i * (i + 1) / 2 + i];
            } else if (i == j) {
                const double hee = (i < n) ? Hee[inst] : 1.0;
                h = hee;
                l = hee > 0.0 ? rsqrt(hee) : 1.0;
            }
        }
        H[e] = h;
        L[e] = l;
    }
}

// Consumer side of the pull protocol: CTA (p, part) waits until rank p has published `epoch` (acquire on the LOCAL flag
// entry p), then copies rank p's rows of slot epoch % slots from p's buffer (peer loads over NVLink, 16-byte accesses)
// into dst rows [row_off[p], row_off[p+1]).  The last CTA out publishes this rank's ack (= epoch) to every peer: the
// slot may be overwritten.  A one-sided all-gather: no sender-side work beyond the step kernel's own epilogue.
struct PullParams {
    const double* src[8];           // every rank's slot buffer [slots x rows_p x n]
    unsigned long long* flags[8];   // every rank's flag array [2 world]
    int row_off[9];
    int world, rank, n, slots, parts;
    unsigned long long epoch;
    long long limit;
    int* timed_out;
    unsigned int* done;             // CTA counter (zero before the launch, reset by the last CTA)
};
__global__ void k_gather_pull(const __grid_constant__ PullParams Q, double* __restrict__ dst) {
    const int p = blockIdx.x / Q.parts, part = blockIdx.x % Q.parts;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const unsigned long long* f = Q.flags[p] + p;  // rank p's own data epoch, polled over NVLink (p != rank)
        const long long t0 = clock64();
        int good = 1;
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= Q.epoch) break;
            if (clock64() - t0 > Q.limit) {
                *Q.timed_out = 1;
                good = 0;
                break;
            }
            __nanosleep(200);
        }
        ok = good;
    }
    __syncthreads();
    if (ok && dst) {
        const int rows = Q.row_off[p + 1] - Q.row_off[p];
        const long cnt = (long)rows * Q.n;  // doubles
        const double* s = Q.src[p] + (long)(Q.epoch % Q.slots) * cnt;
        double* d = dst + (long)Q.row_off[p] * Q.n;
        const long per = (((cnt + Q.parts - 1) / Q.parts) + 1) & ~1L;  // even: every part keeps the 16-byte alignment
        const long lo = min(cnt, part * per), hi = min(cnt, lo + per);
        const bool al = (((size_t)(s + lo) | (size_t)(d + lo)) & 15) == 0;
        if (al) {
            // 16-byte loads, 8 in flight per thread: the copy is one or two NVLink round trips, not a dependent chain
            const long n2 = (hi - lo) >> 1;
            const double2* s2 = reinterpret_cast<const double2*>(s + lo);
            double2* d2 = reinterpret_cast<double2*>(d + lo);
            const long stride = blockDim.x;
            for (long i0 = threadIdx.x; i0 < n2; i0 += 8 * stride) {
                double2 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const long i = i0 + u * stride;
                    if (i < n2) asm volatile("ld.global.relaxed.sys.v2.f64 {%0, %1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(s2 + i) : "memory");
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const long i = i0 + u * stride;
                    if (i < n2) d2[i] = v[u];
                }
            }
            if (((hi - lo) & 1) && threadIdx.x == 0) d[hi - 1] = s[hi - 1];
        } else {
            for (long i = lo + threadIdx.x; i < hi; i += blockDim.x) d[i] = s[i];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(Q.done, 1u);
        if (prev == gridDim.x - 1) {
            *Q.done = 0u;
            __threadfence_system();
            for (int pr = 0; pr < Q.world; ++pr) {  // ack: every rank may reuse the slot of `epoch` as far as this reader goes
                unsigned long long* a = Q.flags[pr] + Q.world + Q.rank;
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(a), "l"(Q.epoch) : "memory");  // (after the fence above)
            }
        }
    }
}

}  // namespace bmpc
