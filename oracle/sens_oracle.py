"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's sensitivity arithmetic
(gp_emu_uqsa/sensitivity/_sensitivityclasses.py), vectorised over the training points.

Pinned against the real reference by tests/golden/sens_*.npz (tests/test_oracle_golden.py).
With B = diag(1/v), C = diag(1/delta^2) every matrix of the reference is diagonal, so each
integral factorises over the input dimensions; the functions below follow the reference's
expressions dimension by dimension (lines cited per function).
"""
import numpy as np


class SensOracle:
    """State of Sensitivity.__init__ (:8-51) + UPSQRT_const (:519-552)."""

    def __init__(self, X, f, H, A, beta, sigma, nugget, delta, m, v):
        self.x, self.f, self.H, self.A = np.asarray(X, float), np.asarray(f, float), np.asarray(H, float), np.asarray(A, float)
        self.beta, self.sigma, self.nugget = np.asarray(beta, float), float(sigma), float(nugget)
        self.m, self.v = np.asarray(m, float), np.asarray(v, float)
        self.b = 1.0 / self.v                                   # diag(B)  (:20)
        self.c = 1.0 / np.asarray(delta, float) ** 2            # diag(C)  (:24)
        self.e = np.linalg.solve(self.A, self.f - self.H @ self.beta)      # :40
        self.G = np.linalg.solve(self.A, self.H)                            # :44
        self.W = np.linalg.inv(self.H.T @ self.G)                           # :42
        b, c = self.b, self.c
        self.dx2 = (self.x - self.m) ** 2                                   # T3 = S3 = P3  (:524)
        self.t1 = np.sqrt(b / (b + 2 * c))                                  # T1 (:522)
        self.t2 = c * b / (b + 2 * c)                                       # T2 = S2 = P2 (:523)
        self.Tk_b4 = self.t1 * np.exp(-self.t2 * self.dx2)                  # :529
        self.T = (1 - self.nugget) * np.prod(self.Tk_b4, axis=1)            # :530
        self.R = np.append([1.0], self.m)                                   # :533
        self.U = (1 - self.nugget) * np.prod(np.sqrt(b / (b + 4 * c)))      # :537
        self.P1 = b / (b + 2 * c)                                           # :548
        self.P4 = np.sqrt(b / (b + 4 * c))                                  # :551
        self.P5 = 0.5 / (b + 4 * c)                                         # :552

    # ---- the n x n product-form matrix of Pw_calc (:621-626) for an index set w ----------------
    def Pw(self, w):
        d = self.m.size
        wb = [k for k in range(d) if k not in w]
        x, c, b = self.x, self.c, self.b
        out = np.full((x.shape[0], x.shape[0]), (1 - self.nugget) ** 2)
        for i in wb:      # P1 * P_prod[k,l,i]   (:602, :625)
            s = self.dx2[:, i]
            out = out * (self.P1[i] * np.exp(-self.t2[i] * (s[:, None] + s[None, :])))
        for i in w:       # P_b4_prod[k,l,i]     (:603-607, :626)
            s = self.dx2[:, i]
            dd = (x[:, i][:, None] - x[:, i][None, :]) ** 2
            out = out * (self.P4[i] * np.exp(-self.P5[i] * (4 * c[i] * c[i] * dd + 2 * c[i] * b[i] * (s[:, None] + s[None, :]))))
        return out

    def Qw(self, w):
        """:554-582 -- Q_w = R R^T with var added on the w block."""
        Q = np.outer(self.R, self.R)
        for i in w:
            Q[1 + i, 1 + i] += 1.0 / self.b[i]
        return Q

    def Estar(self, w):
        """:584-596"""
        n, d = self.x.shape
        E = np.ones((1 + d, n))
        for k in range(d):
            if k in w:
                E[1 + k] = (2 * self.c[k] * self.x[:, k] + self.b[k] * self.m[k]) / (2 * self.c[k] + self.b[k])
            else:
                E[1 + k] = self.m[k]
        return E

    def Uw(self, w):
        """:610-613"""
        wb = [k for k in range(self.m.size) if k not in w]
        return (1 - self.nugget) * np.prod(self.P4[wb])

    def Tw(self, w, xw):
        """:628-633"""
        wb = [k for k in range(self.m.size) if k not in w]
        xw = np.atleast_1d(np.asarray(xw, float))
        val = np.prod(self.Tk_b4[:, wb], axis=1)
        dq = ((xw[None, :] - self.x[:, w]) ** 2 * self.c[w]).sum(1)
        return (1 - self.nugget) * val * np.exp(-dq)

    # ---- public quantities -----------------------------------------------------------------
    def EVint(self, w):
        """E*[V_w] as computed inside sensitivity() (:481-506) for the index set w."""
        s2 = self.sigma ** 2
        Pw, Qw = self.Pw(w), self.Qw(w)
        Sw = self.Estar(w) * self.T[None, :]                    # Sw_calc :615-619
        AiPw = np.linalg.solve(self.A, Pw)
        EEE = s2 * (self.Uw(w) - np.trace(AiPw)
                    + np.trace(self.W @ (Qw - Sw @ self.G - self.G.T @ Sw.T + self.H.T @ AiPw @ self.G))) \
            + self.e @ Pw @ self.e + 2.0 * self.beta @ Sw @ self.e + self.beta @ Qw @ self.beta
        TG = self.T @ self.G
        EE2 = s2 * (self.U - self.T @ np.linalg.solve(self.A, self.T) + (self.R - TG) @ self.W @ (self.R - TG)) \
            + (self.R @ self.beta + self.T @ self.e) ** 2
        return EEE - EE2

    def sensitivity(self):
        return np.array([self.EVint([P]) for P in range(self.m.size)])

    def uncertainty(self):
        """uE, uV, uEV (:54-203)."""
        b, c, m, x = self.b, self.c, self.m, self.x
        nug = self.nugget
        Rh = np.append([1.0], m)                                                   # :61
        Rhh = np.outer(Rh, Rh)
        Rhh[1:, 1:] += np.diag(1.0 / b)                                            # :63-74
        mpk = (2 * c * x + b * m) / (2 * c + b)                                    # :80-81
        Qk = (2 * c * (mpk - x) ** 2 + b * (mpk - m) ** 2).sum(1)                  # :82-83
        Rt = (1 - nug) * np.sqrt(np.prod(b) / np.prod(2 * c + b)) * np.exp(-0.5 * Qk)   # :84-86
        Rht = np.vstack([np.ones(x.shape[0]), mpk.T]) * Rt[None, :]               # :87-88
        Rtt = self.Pw(list(range(m.size)))                                         # :90-102 (same integral)
        # U2 = (1-nu) det(B)/sqrt(det(Bbold)), Bbold = [[2C+B, -2C], [-2C, 2C+B]]  (:105-113)
        U2 = (1 - nug) * np.prod(b) / np.sqrt(np.prod((2 * c + b) ** 2 - 4 * c * c))
        s2 = self.sigma ** 2
        GtRt = self.G.T @ Rt
        uE = Rh @ self.beta + Rt @ self.e                                          # :186
        uV = s2 * (U2 - Rt @ np.linalg.solve(self.A, Rt) + (Rh - GtRt) @ self.W @ (Rh - GtRt))      # :187-189
        I1 = s2 * (1.0 - np.trace(np.linalg.solve(self.A, Rtt))
                   + np.trace(self.W @ (Rhh - 2.0 * Rht @ self.G + self.G.T @ Rtt @ self.G)))          # :190-194
        I2 = self.beta @ Rhh @ self.beta + 2.0 * self.beta @ Rht @ self.e + self.e @ Rtt @ self.e   # :195-197
        uEV = (I1 - uV) + (I2 - uE ** 2)                                           # :199
        return uE, uV, uEV

    def main_effect(self, input_range, points=100, w=None):
        """effect[P, j], mean_effect[P, j] (:238-285)."""
        d = self.m.size
        w = range(d) if w is None else w
        effect, mean_effect = np.zeros((d, points)), np.zeros((d, points))
        for P in w:
            for j, xw in enumerate(np.linspace(input_range[P][0], input_range[P][1], points)):
                Tw = self.Tw([P], xw)
                Rw = np.append([1.0], self.m.copy())
                Rw[1 + P] = xw
                mean_effect[P, j] = Rw @ self.beta + Tw @ self.e
                effect[P, j] = (Rw - self.R) @ self.beta + (Tw - self.T) @ self.e
        return effect, mean_effect

    def totaleffectvariance(self, uEV):
        """EVTw[P] (:405-463).  Reference quirk (verified against the real code): Qw/Sw/Pw/Uw are
        computed for w = [P] *before* w and wb are swapped (:421-427), so the value is
        uEV - E*[V_P], not uEV - E*[V_wb]."""
        return np.array([uEV - self.EVint([P]) for P in range(self.m.size)])
