"""``Sensitivity`` with the reference's interface (gp_emu_uqsa/sensitivity/_sensitivityclasses.py):
MUCM "case 2" uncertainty and sensitivity analysis of a Gaussian-kernel, linear-mean emulator
against independent Gaussian inputs N(m, diag v).

Every O(n^2)/O(n^3) piece runs on the B200 through the C-ABI: the solves with the training matrix
(``gpe_solve``), the product-form n x n integrals Rtt / Pw and their contractions with A^-1, G and e
(``gpe_sens_contract``; the matrices are generated tile by tile on device and never reach the host),
and the main-effect sweep over x_w (``gpe_sens_main_effect``).  What stays here is O(n d) / O(q^2)
bookkeeping.  With B = diag(1/v), C = diag(1/delta^2) every matrix in the reference is diagonal, so
the integrals factorise over the input dimensions; reference line numbers are cited per method.
"""
import numpy as np


class Sensitivity:
    def __init__(self, emul, m, v):
        self.v, self.m = np.asarray(v, dtype=float), np.asarray(m, dtype=float)
        self.x = emul.training.inputs
        self.input_range = emul.all_data.input_range
        self.minmax = emul.all_data.minmax
        self.b = 1.0 / self.v                                         # diag(B)  (:20)
        self.c = 1.0 / np.array(emul.par.delta, dtype=float) ** 2     # diag(C)  (:24)
        self.B, self.C = np.diag(self.b), np.diag(self.c)
        self.f = emul.training.outputs
        self.H = emul.training.H
        self.beta = np.asarray(emul.par.beta, dtype=float)
        self.sigma = emul.par.sigma
        self.nugget = emul.par.nugget
        self._train = emul.training
        # factor A once on device; e = A^-1 (f - H beta), G = A^-1 H (:40-44), W = (H^T G)^-1 (:42)
        self._dev, _, _, st = emul.training.fit(beta=self.beta, r_div=emul.training._A_args[0])
        if st != 0:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        sol = self._dev.solve(np.column_stack([self.H, self.f - self.H.dot(self.beta)]))
        self.G, self.e = np.ascontiguousarray(sol[:, :-1]), np.ascontiguousarray(sol[:, -1])
        self.W = np.linalg.inv(self.H.T.dot(self.G))
        self.UPSQRT_const()
        self.done_uncertainty = self.done_sensitivity = self.done_main_effect = False
        self.done_interaction = self.done_totaleffectvar = False

    @property
    def A(self):
        """The training covariance matrix (reference attribute; built on device when read)."""
        return self._train.A

    # ------------------------------------------------------------------ constants (:519-552)
    def UPSQRT_const(self):
        b, c = self.b, self.c
        self.dx2 = (self.x - self.m) ** 2
        self.t1 = np.sqrt(b / (b + 2.0 * c))
        self.t2 = c * b / (b + 2.0 * c)
        self.Tk_b4_prod = self.t1 * np.exp(-self.t2 * self.dx2)
        self.T = (1.0 - self.nugget) * np.prod(self.Tk_b4_prod, axis=1)
        self.R = np.append([1.0], self.m)
        self.Q = np.outer(self.R, self.R)
        self.U = (1.0 - self.nugget) * np.prod(np.sqrt(b / (b + 4.0 * c)))
        self.P1 = b / (b + 2.0 * c)
        self.P4 = np.sqrt(b / (b + 4.0 * c))
        self.P5 = 0.5 / (b + 4.0 * c)
        self._AiT = None

    # ------------------------------------------------------------------ device contractions
    def _pw_contract(self, w):
        """tr(A^-1 Pw), [G|e]^T Pw [G|e] for the index set w (Pw_calc :621-626, P_prod_calc :599-607):
        i in wb contributes P1_i exp(-t2_i (dx_k^2 + dx_l^2)); i in w contributes
        P4_i exp(-P5_i (4 c_i^2 (x_k - x_l)^2 + 2 c_i b_i (dx_k^2 + dx_l^2)))."""
        d = self.m.size
        inw = np.zeros(d, dtype=bool)
        inw[list(w)] = True
        gamma = np.where(inw, 4.0 * self.P5 * self.c * self.c, 0.0)
        acoef = np.where(inw, 2.0 * self.P5 * self.c * self.b, self.t2)
        scale = (1.0 - self.nugget) ** 2 * np.prod(np.where(inw, self.P4, self.P1))
        V = np.column_stack([self.G, self.e])
        tr, M = self._dev.sens_contract(gamma, acoef, self.m, scale, V)
        q = self.G.shape[1]
        return tr, M[:q, :q], M[q, q]

    def _Ainv_T(self):
        if self._AiT is None:
            self._AiT = self._dev.solve(self.T)
        return self._AiT

    # ------------------------------------------------------------------ uncertainty (:54-203)
    def uncertainty(self):
        print("\n*** Uncertainty measures ***")
        self.done_uncertainty = True
        b, c, m, x = self.b, self.c, self.m, self.x
        nug = self.nugget
        self.w = list(range(m.size))
        self.Rh = np.append([1.0], m)
        self.Rhh = np.outer(self.Rh, self.Rh)
        self.Rhh[1:, 1:] += np.diag(1.0 / b)
        mpk = (2.0 * c * x + b * m) / (2.0 * c + b)
        Qk = (2.0 * c * (mpk - x) ** 2 + b * (mpk - m) ** 2).sum(axis=1)
        self.Rt = (1.0 - nug) * np.sqrt(np.prod(b) / np.prod(2.0 * c + b)) * np.exp(-0.5 * Qk)
        self.Rht = np.vstack([np.ones(x.shape[0]), mpk.T]) * self.Rt[None, :]
        self.U2 = (1.0 - nug) * np.prod(b) / np.sqrt(np.prod((2.0 * c + b) ** 2 - 4.0 * c * c))
        self.Utild = 1
        # Rtt (:90-102) is the all-inputs product-form matrix: contracted on device
        trRtt, GRttG, eRtte = self._pw_contract(self.w)
        AiRt = self._dev.solve(self.Rt)
        s2 = self.sigma ** 2
        GtRt = self.G.T.dot(self.Rt)
        self.uE = self.Rh.dot(self.beta) + self.Rt.dot(self.e)
        self.uV = s2 * (self.U2 - self.Rt.dot(AiRt) + (self.Rh - GtRt).dot(self.W).dot(self.Rh - GtRt))
        self.I1 = s2 * (self.Utild - trRtt
                        + np.trace(self.W.dot(self.Rhh - 2.0 * self.Rht.dot(self.G) + GRttG)))
        self.I2 = self.beta.dot(self.Rhh).dot(self.beta) + 2.0 * self.beta.dot(self.Rht).dot(self.e) + eRtte
        self.uEV = (self.I1 - self.uV) + (self.I2 - self.uE ** 2)
        print("E*[ E[f(X)] ]  :", self.uE)
        print("var*[ E[f(X)] ]:", self.uV)
        print("E*[ var[f(X)] ]:", self.uEV)

    # ------------------------------------------------------------------ helpers with the reference's names
    def setup_w_wb(self, P):
        self.w = [P]
        self.wb = [k for k in range(len(self.m)) if k not in self.w]

    def Qw_calc(self):
        """:554-582"""
        self.Qw = np.outer(self.R, self.R)
        for i in self.w:
            self.Qw[1 + i, 1 + i] += 1.0 / self.b[i]

    def Estar_calc(self):
        """:584-596"""
        n, d = self.x.shape
        self.Estar = np.ones((1 + d, n))
        for k in range(d):
            if k in self.w:
                self.Estar[1 + k] = (2 * self.c[k] * self.x[:, k] + self.b[k] * self.m[k]) / (2 * self.c[k] + self.b[k])
            else:
                self.Estar[1 + k] = self.m[k]

    def Uw_calc(self):
        """:610-613"""
        self.Uw_b4_prod = self.P4
        self.Uw = (1.0 - self.nugget) * np.prod(self.P4[self.wb])

    def Sw_calc(self):
        """:615-619"""
        self.Sw = self.Estar * self.T[None, :]

    def Tw_calc(self):
        """:628-633 (host form for single x_w values; the main-effect sweep runs on device)"""
        xw = np.atleast_1d(np.asarray(self.xw, dtype=float))
        val = np.prod(self.Tk_b4_prod[:, self.wb], axis=1)
        dq = ((xw[None, :] - self.x[:, self.w]) ** 2 * self.c[self.w]).sum(axis=1)
        self.Tw = (1.0 - self.nugget) * val * np.exp(-dq)

    def Rw_calc(self):
        """:635-638"""
        Rwno1 = np.array(self.m)
        Rwno1[self.w] = self.xw
        self.Rw = np.append([1.0], Rwno1)

    def _EVint(self):
        """EEE - EE2 for the current self.w / self.wb (:481-506)."""
        s2 = self.sigma ** 2
        self.Qw_calc(); self.Estar_calc(); self.Uw_calc(); self.Sw_calc()
        trPw, GPwG, ePwe = self._pw_contract(self.w)
        SwG = self.Sw.dot(self.G)
        self.EEE = s2 * (self.Uw - trPw + np.trace(self.W.dot(self.Qw - SwG - SwG.T + GPwG))) \
            + ePwe + 2.0 * self.beta.dot(self.Sw).dot(self.e) + self.beta.dot(self.Qw).dot(self.beta)
        TG = self.T.dot(self.G)
        self.EE2 = s2 * (self.U - self.T.dot(self._Ainv_T()) + (self.R - TG).dot(self.W).dot(self.R - TG)) \
            + (self.R.dot(self.beta) + self.T.dot(self.e)) ** 2
        return self.EEE - self.EE2

    # ------------------------------------------------------------------ sensitivity (:466-516)
    def sensitivity(self):
        print("\n*** Calculate sensitivity indices ***")
        self.done_sensitivity = True
        self.senseindex = np.zeros([self.m.size])
        for P in range(len(self.m)):
            self.setup_w_wb(P)
            self.xw = self.m[P]
            self.EVint = self._EVint()
            if self.done_uncertainty:
                print("E(V" + str(self.w) + ")/EV:", self.EVint / self.uEV)
            else:
                print("E(V" + str(self.w) + "):", self.EVint)
            self.senseindex[P] = self.EVint
        if self.done_uncertainty:
            print("Sum of Sensitivities:", np.sum(self.senseindex / self.uEV))

    # ------------------------------------------------------------------ total effect variance (:405-463)
    def totaleffectvariance(self):
        self.done_totaleffectvar = True
        print("\n*** Calculate total effect variance ***")
        self.senseindexwb = np.zeros([self.m.size])
        self.EVTw = np.zeros([self.m.size])
        self.EVf = self.uEV
        print("E*[ var[f(X)] ]:", self.EVf)
        for P in range(len(self.m)):
            # reference behaviour: Qw/Sw/Pw/Uw are built for w = [P] before w and wb are swapped
            # (:421-427), so E*[V] here is the one of w = [P]
            self.setup_w_wb(P)
            self.EVaaa = self._EVint()
            self.w, self.wb = self.wb, self.w
            self.senseindexwb[P] = self.EVaaa
            self.EVTw[P] = self.EVf - self.EVaaa
            print("E(V[T" + str(P) + "]):", self.EVTw[P])

    # ------------------------------------------------------------------ main effect (:238-324)
    def main_effect(self, plot=False, points=100, customKey=[], customLabels=[], plotShrink=0.9, w=[], black_white=False):
        print("\n*** Main effect measures ***")
        self.done_main_effect = True
        self.effect = np.zeros([self.m.size, points])
        self.mean_effect = np.zeros([self.m.size, points])
        if w == []:
            w = range(0, len(self.m))
        w = list(w)
        xw = np.array([np.linspace(self.input_range[P][0], self.input_range[P][1], points) for P in w])
        for P in w:
            print("Main effect measures for input", P, "range", self.input_range[P])
        # Tw . e for every (input, x_w) in one device sweep (Tw_calc :628-633)
        Te = self._dev.sens_main_effect(self.t1, self.t2, self.c, self.m, self.e, 1.0 - self.nugget, w, xw)
        Rb, Tdot = self.R.dot(self.beta), self.T.dot(self.e)
        for r, P in enumerate(w):
            RwB = Rb + (xw[r] - self.m[P]) * self.beta[1 + P]          # Rw . beta, Rw = R with m_P -> x_w (:635-638)
            self.mean_effect[P] = RwB + Te[r]
            self.effect[P] = (RwB - Rb) + (Te[r] - Tdot)
        if plot:
            print("Plotting main effects requires matplotlib (outside the rebuilt hot path); values are in .effect")

    def interaction_effect(self, i, j, points=25, customLabels=[]):
        """:327-401.  Values only (the contour plot needs matplotlib)."""
        print("\n*** Interaction effects ***")
        self.done_interaction = True
        self.interaction = np.zeros([points, points])
        print("Recalculating main effect with", points, "points...")
        self.main_effect(plot=False, points=points, w=[i, j])
        self.w = [i, j]
        self.wb = [k for k in range(len(self.m)) if k not in self.w]
        ra_i, ra_j = self.input_range[i], self.input_range[j]
        print("\nCalculating", points * points, "interaction effects...")
        for ic, xwi in enumerate(np.linspace(ra_i[0], ra_i[1], points)):
            for jc, xwj in enumerate(np.linspace(ra_j[0], ra_j[1], points)):
                self.xw = np.array([xwi, xwj])
                self.Tw_calc(); self.Rw_calc()
                self.IE = (self.Rw + self.R).dot(self.beta) + (self.Tw + self.T).dot(self.e) \
                    - self.mean_effect[i, ic] - self.mean_effect[j, jc]
                self.interaction[ic, jc] = self.IE

    # ------------------------------------------------------------------ results file (:641-662)
    def to_file(self, filename):
        print("Sensitivity & Uncertainty results to file...")
        with open(filename, 'w') as f:
            if self.done_uncertainty:
                f.write("EE " + str(self.uE) + "\n")
                f.write("VE " + str(self.uV) + "\n")
                f.write("EV " + str(self.uEV) + "\n")
            if self.done_sensitivity:
                f.write("EVw " + ' '.join(map(str, self.senseindex)) + "\n")
            if self.done_totaleffectvar:
                f.write("EVTw " + ' '.join(map(str, self.EVTw)) + "\n")
            if self.done_main_effect:
                f.write("xw " + ' '.join(map(str, [i for i in np.linspace(0.0, 1.0, self.effect[0].size)])) + "\n")
                for i in range(0, len(self.m)):
                    f.write("ME" + str(i) + " " + ' '.join(map(str, self.effect[i])) + "\n")
