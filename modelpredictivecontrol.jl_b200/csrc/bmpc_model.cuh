// On-device construction of the prediction matrices and the Hessian from the augmented model
// (route A of bmpc.h; also the batched setmodel!, reference src/controller/execute.jl:621-790).
// Restates init_predmat (LinModel + SingleShooting, src/controller/transcription.jl:115-194) and
// init_quadprog (src/controller/construct.jl:837-845), but emits the matrices directly in
// input-level coordinates:  Ev = E*D is assembled from the step-response blocks
// W(i) = Ĉ S(i) B̂u  (which are also the blocks of V):
//     Ev[step s, block l] = W(s-1-j_l) - W(s-1-j_{l+1}),   W(i<0) = 0,
// i.e. the response to a unit pulse of u held over move block l.
// One CTA per instance; all matrices column-major.
#pragma once
#include <cuda_runtime.h>

namespace bmpc {

struct ModelParams {
    int nu, ny, nd, nx, Hp, Hc, nz, nY, nU, neps, nHp2;
    long nEv2;
    long sModel;  // 1 per-instance inputs, 0 never (inputs always carry NM copies)
    double Cwt;
    const double *A, *Bu, *C, *Bd, *Dd, *f, *Mdiag, *Ndiag, *Ldiag;
    double *K, *V, *B, *G, *J, *kx, *vx, *bx, *gx, *jx, *Ev, *exv, *Hv, *Hee;
    const int* blk_start;  // [Hc+1] j_l
};

__global__ void __launch_bounds__(128) k_build_model(const __grid_constant__ ModelParams P) {
    extern __shared__ double sm[];
    const int nx = P.nx, nu = P.nu, ny = P.ny, nd = P.nd, Hp = P.Hp, Hc = P.Hc, nz = P.nz, nY = P.nY;
    const long inst = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    double* A = sm;                    // nx*nx
    double* Ap = A + nx * nx;          // Â^i
    double* Ap2 = Ap + nx * nx;
    double* Pm = Ap2 + nx * nx;        // Ĉ Â^i   (ny x nx)
    double* Pm2 = Pm + ny * nx;
    double* Xu = Pm2 + ny * nx;        // Â^i B̂u  (nx x nu)
    double* Xu2 = Xu + nx * nu;
    double* Su = Xu2 + nx * nu;        // S(i) B̂u
    double* Xf = Su + nx * nu;         // Â^i f
    double* Xf2 = Xf + nx;
    double* Sf = Xf2 + nx;             // S(i) f
    double* Xd = Sf + nx;              // Â^i B̂d  (nx x nd)
    double* Xd2 = Xd + nx * nd;
    double* Cm = Xd2 + nx * nd;        // Ĉ (ny x nx)

    const double* gA = P.A + inst * nx * nx;
    const double* gBu = P.Bu + inst * nx * nu;
    const double* gC = P.C + inst * ny * nx;
    const double* gBd = nd ? P.Bd + inst * nx * nd : nullptr;
    const double* gDd = nd ? P.Dd + inst * ny * nd : nullptr;
    const double* gf = P.f + inst * nx;
    double* K = P.K + inst * (long)nY * nx;
    double* V = P.V + inst * (long)nY * nu;
    double* B = P.B + inst * (long)nY;
    double* G = nd ? P.G + inst * (long)nY * nd : nullptr;
    double* J = nd ? P.J + inst * (long)nY * nd * Hp : nullptr;
    double* kx = P.kx + inst * (long)nx * nx;
    double* vx = P.vx + inst * (long)nx * nu;
    double* bx = P.bx + inst * (long)nx;
    double* gx = nd ? P.gx + inst * (long)nx * nd : nullptr;
    double* jx = nd ? P.jx + inst * (long)nx * nd * Hp : nullptr;
    double* Ev = P.Ev + inst * P.nEv2;
    double* exv = P.exv + inst * (long)nx * nz;
    double* Hv = P.Hv + inst * (long)P.nHp2;

    for (int e = tid; e < nx * nx; e += nt) {
        A[e] = gA[e];
        Ap[e] = (e / nx == e % nx) ? 1.0 : 0.0;
    }
    for (int e = tid; e < ny * nx; e += nt) Cm[e] = Pm[e] = gC[e];
    for (int e = tid; e < nx * nu; e += nt) Xu[e] = Su[e] = gBu[e];
    for (int e = tid; e < nx; e += nt) Xf[e] = Sf[e] = gf[e];
    for (int e = tid; e < nx * nd; e += nt) Xd[e] = gBd[e];
    for (int e = tid; e < nx * nz; e += nt) exv[e] = 0.0;
    if (nd)
        for (long e = tid; e < (long)nx * nd * Hp; e += nt) jx[e] = 0.0;
    __syncthreads();

    for (int s = 1; s <= Hp; ++s) {
        const int i = s - 1;  // Su = S(i)B̂u, Sf = S(i)f, Xd = Â^i B̂d, Pm = ĈÂ^i, Ap = Â^i
        for (int e = tid; e < ny * nu; e += nt) {  // V block s = Ĉ S(i) B̂u
            const int o = e % ny, c = e / ny;
            double a = 0.0;
            for (int k = 0; k < nx; ++k) a = fma(Cm[o + ny * k], Su[k + nx * c], a);
            V[(long)i * ny + o + (long)nY * c] = a;
        }
        for (int o = tid; o < ny; o += nt) {  // B block s = Ĉ S(i) f
            double a = 0.0;
            for (int k = 0; k < nx; ++k) a = fma(Cm[o + ny * k], Sf[k], a);
            B[(long)i * ny + o] = a;
        }
        for (int e = tid; e < ny * nd; e += nt) {  // G block s = Ĉ Â^i B̂d
            const int o = e % ny, c = e / ny;
            double a = 0.0;
            for (int k = 0; k < nx; ++k) a = fma(Cm[o + ny * k], Xd[k + nx * c], a);
            G[(long)i * ny + o + (long)nY * c] = a;
        }
        // terminal: exv[:, l] = S(Hp-1-j_l)B̂u - S(Hp-1-j_{l+1})B̂u
        for (int l = 0; l < Hc; ++l) {
            if (i == Hp - 1 - P.blk_start[l]) {
                for (int e = tid; e < nx * nu; e += nt) {
                    const int k = e % nx, c = e / nx;
                    exv[k + (long)nx * (l * nu + c)] += Su[e];
                    if (l >= 1) exv[k + (long)nx * ((l - 1) * nu + c)] -= Su[e];
                }
            }
        }
        if (nd && i <= Hp - 2) {  // jx block jj = Hp-1-i (1-based) = Â^i B̂d
            const int jj = Hp - 1 - i;
            for (int e = tid; e < nx * nd; e += nt) jx[(e % nx) + (long)nx * ((jj - 1) * nd + e / nx)] = Xd[e];
        }
        if (s == Hp) {
            for (int e = tid; e < nx * nu; e += nt) vx[e] = Su[e];
            for (int e = tid; e < nx; e += nt) bx[e] = Sf[e];
            for (int e = tid; e < nx * nd; e += nt) gx[e] = Xd[e];
        }
        // advance to index i+1
        for (int e = tid; e < ny * nx; e += nt) {  // Pm2 = Pm * A ; K block s = Ĉ Â^s
            const int o = e % ny, c = e / ny;
            double a = 0.0;
            for (int k = 0; k < nx; ++k) a = fma(Pm[o + ny * k], A[k + nx * c], a);
            Pm2[e] = a;
            K[(long)i * ny + o + (long)nY * c] = a;
        }
        for (int e = tid; e < nx * nx; e += nt) {
            const int r = e % nx, c = e / nx;
            double a = 0.0;
            for (int k = 0; k < nx; ++k) a = fma(Ap[r + nx * k], A[k + nx * c], a);
            Ap2[e] = a;
        }
        for (int e = tid; e < nx * nu; e += nt) {
            const int r = e % nx, c = e / nx;
            double a = 0.0;
            for (int k = 0; k < nx; ++k) a = fma(A[r + nx * k], Xu[k + nx * c], a);
            Xu2[e] = a;
        }
        for (int r = tid; r < nx; r += nt) {
            double a = 0.0;
            for (int k = 0; k < nx; ++k) a = fma(A[r + nx * k], Xf[k], a);
            Xf2[r] = a;
        }
        for (int e = tid; e < nx * nd; e += nt) {
            const int r = e % nx, c = e / nx;
            double a = 0.0;
            for (int k = 0; k < nx; ++k) a = fma(A[r + nx * k], Xd[k + nx * c], a);
            Xd2[e] = a;
        }
        __syncthreads();
        for (int e = tid; e < ny * nx; e += nt) Pm[e] = Pm2[e];
        for (int e = tid; e < nx * nx; e += nt) Ap[e] = Ap2[e];
        for (int e = tid; e < nx * nu; e += nt) {
            Xu[e] = Xu2[e];
            Su[e] += Xu2[e];
        }
        for (int e = tid; e < nx; e += nt) {
            Xf[e] = Xf2[e];
            Sf[e] += Xf2[e];
        }
        for (int e = tid; e < nx * nd; e += nt) Xd[e] = Xd2[e];
        __syncthreads();
    }
    for (int e = tid; e < nx * nx; e += nt) kx[e] = Ap[e];  // Â^Hp
    __syncthreads();

    // Ev from the blocks of V
    for (long e = tid; e < (long)nY * nz; e += nt) {
        const int t = (int)(e % nY), j = (int)(e / nY);
        const int s1 = t / ny, o = t % ny;  // s1 = s-1
        const int l = j / nu, c = j % nu;
        const int m0 = s1 - P.blk_start[l], m1 = s1 - P.blk_start[l + 1];
        double a = 0.0;
        if (m0 >= 0) a += V[(long)m0 * ny + o + (long)nY * c];
        if (m1 >= 0) a -= V[(long)m1 * ny + o + (long)nY * c];
        Ev[e] = a;
    }
    if (nd) {
        for (long e = tid; e < (long)nY * nd * Hp; e += nt) {
            const int t = (int)(e % nY);
            const int col = (int)(e / nY);
            const int i1 = t / ny, o = t % ny, j1 = col / nd, c = col % nd;  // 0-based steps
            double a = 0.0;
            if (i1 == j1)
                a = gDd[o + ny * c];
            else if (i1 > j1)
                a = G[(long)(i1 - j1 - 1) * ny + o + (long)nY * c];
            J[e] = a;
        }
    }
    __syncthreads();
    // Hv = 2 (Ev' M Ev + D' N D + blocksum(L)), packed row-major lower
    const double* Md = P.Mdiag + inst * nY;
    const double* Nd = P.Ndiag + inst * nz;
    const double* Ld = P.Ldiag + inst * P.nU;
    const int npair = nz * (nz + 1) / 2;
    for (int p = tid; p < npair; p += nt) {
        int a = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
        while ((a + 1) * (a + 2) / 2 <= p) ++a;
        while (a * (a + 1) / 2 > p) --a;
        const int b = p - a * (a + 1) / 2;
        double acc = 0.0;
        const double* ca = Ev + (long)nY * a;
        const double* cb = Ev + (long)nY * b;
        for (int t = 0; t < nY; ++t) acc = fma(ca[t] * Md[t], cb[t], acc);
        if (a == b) {
            acc += Nd[a] + (a + nu < nz ? Nd[a + nu] : 0.0);
            const int l = a / nu, ch = a % nu;
            for (int t = P.blk_start[l]; t < P.blk_start[l + 1]; ++t) acc += Ld[t * nu + ch];
        } else if (a == b + nu) {
            acc -= Nd[a];
        }
        Hv[p] = 2.0 * acc;
    }
    if (tid == 0) P.Hee[inst] = P.neps ? 2.0 * P.Cwt : 0.0;
}

}  // namespace bmpc
