"""Test helpers: build GPU batches from oracle controllers and generate the synthetic plants of
SURVEY.md section 8(d).  Test infrastructure only."""
import numpy as np

from oracle.linmpc import LinModel, LinMPC


def random_plant(rng, nx=4, nu=2, ny=2, rho=(0.5, 0.95)):
    """Random stable plant of the batch configs: A ~ N(0,1) scaled to spectral radius U(rho)."""
    A = rng.standard_normal((nx, nx))
    A *= rng.uniform(*rho) / np.abs(np.linalg.eigvals(A)).max()
    return LinModel(A, rng.standard_normal((nx, nu)), rng.standard_normal((ny, nx)))


def batch_from_oracle(mpcs, shared_model=False, team=0, tol=0.0, max_iter=0, with_terminal=None):
    """Create a BatchLinMPC holding exactly the controllers in ``mpcs`` (route B: host matrices)."""
    import mpc_b200
    m0 = mpcs[0]
    N = len(mpcs)
    model = m0.model
    b = mpc_b200.BatchLinMPC(N, model.nu, model.ny, m0.estim.nxhat, m0.Hp, m0.nb, nd=model.nd, Cwt=m0.Cwt,
                             shared_model=shared_model, team=team, tol=tol, max_iter=max_iter)
    src = mpcs[:1] if shared_model else mpcs
    st = lambda name: np.stack([getattr(m, name) for m in src])
    if with_terminal is None:
        with_terminal = any(np.isfinite(m.con.xhat0min).any() or np.isfinite(m.con.xhat0max).any() for m in mpcs)
    term = dict(ex=st("ex"), kx=st("kx"), vx=st("vx"), bx=st("bx")) if with_terminal else {}
    if with_terminal and model.nd:
        term.update(gx=st("gx"), jx=st("jx"))
    dist = dict(G=st("G"), J=st("J")) if model.nd else {}
    b.set_predmat(st("E"), st("K"), st("V"), st("B"), st("Htilde"), **dist, **term)
    M = st("M_Hp")
    if all(np.count_nonzero(Mi - np.diag(np.diag(Mi))) == 0 for Mi in M):
        M = np.stack([np.diag(Mi) for Mi in M])
    L = st("L_Hp")
    if all(np.count_nonzero(Li - np.diag(np.diag(Li))) == 0 for Li in L):
        L = np.stack([np.diag(Li) for Li in L])
    b.set_weights(M, L)
    b.set_oppoints(np.stack([m.model.uop for m in src]), np.stack([m.model.yop for m in src]))
    push_constraints(b, mpcs)
    return b


def push_constraints(b, mpcs):
    c0 = mpcs[0].con
    sc = lambda name: np.stack([getattr(m.con, name) for m in mpcs])
    soft = None
    if mpcs[0].neps:
        soft = dict(C_umin=c0.C_umin, C_umax=c0.C_umax, C_dumin=c0.C_dumin, C_dumax=c0.C_dumax,
                    C_ymin=c0.C_ymin, C_ymax=c0.C_ymax, c_xmin=c0.c_xmin, c_xmax=c0.c_xmax)
    b.set_constraints(sc("U0min"), sc("U0max"), sc("DUmin"), sc("DUmax"), sc("Y0min"), sc("Y0max"),
                      sc("xhat0min"), sc("xhat0max"), soft)


def c1_controllers(N, seed=1, nx=4, nu=2, ny=2, Hp=20, Hc=5):
    """Config C1/C2 recipe: random stable plants, nint_ym=1, Mwt=1 Nwt=0.1 Cwt=1e5,
    hard u in [-1,1], soft ymax = 0.8."""
    rng = np.random.default_rng(seed)
    mpcs, plants = [], []
    for _ in range(N):
        m = random_plant(rng, nx, nu, ny)
        mpc = LinMPC(m, Hp=Hp, Hc=Hc, Cwt=1e5)
        mpc.setconstraint(umin=[-1.0] * nu, umax=[1.0] * nu, ymax=[0.8] * ny)
        mpcs.append(mpc)
        plants.append(LinModel(m.A, m.Bu, m.C))
    return mpcs, plants, rng
