"""Index layout of the centroidal-MPC NLP vectors at the C-ABI boundary ("CasADi order").

Decoded from the reference's generated code:
/root/reference/src/centroidal-mpc-walking/config/robots/ergoCubGazeboV1/tmp.c:62-67 (sparsity casadi_s0..s5),
x[45N+15], p[50N+27], g[53N+15].  Two contacts (0 = left_foot, 1 = right_foot), four corners each.
All 3 x (.) blocks are column major: knot k of a 3-vector trajectory is at base + 3k .. base + 3k + 2.
"""
from __future__ import annotations

NC, NJ, NF = 2, 4, 4


class Layout:
    def __init__(self, N: int):
        self.N = N
        self.n = 45 * N + 15
        self.np = 50 * N + 27
        self.m = 53 * N + 15

    # ---- x
    def x_com(self, k): return 3 * k
    def x_dcom(self, k): return 3 * (self.N + 1) + 3 * k
    def x_h(self, k): return 6 * (self.N + 1) + 3 * k
    def x_cbase(self, c): return 9 * (self.N + 1) + c * (18 * self.N + 3)
    def x_pos(self, c, k): return self.x_cbase(c) + 3 * k
    def x_vel(self, c, k): return self.x_cbase(c) + 3 * (self.N + 1) + 3 * k
    def x_frc(self, c, j, k): return self.x_cbase(c) + 6 * self.N + 3 + 3 * self.N * j + 3 * k

    # ---- p
    def p_cbase(self, c): return c * (19 * self.N + 6)
    def p_rot(self, c, k): return self.p_cbase(c) + 9 * k            # vec(R_k), column major
    def p_upper(self, c, k): return self.p_cbase(c) + 9 * self.N + 3 * k   # step-box limits (never enter f/g)
    def p_lower(self, c, k): return self.p_cbase(c) + 12 * self.N + 3 * k
    def p_en(self, c, k): return self.p_cbase(c) + 15 * self.N + k
    def p_nom(self, c, k): return self.p_cbase(c) + 16 * self.N + 3 * k
    def p_cur(self, c): return self.p_cbase(c) + 19 * self.N + 3
    def p_glob(self): return 38 * self.N + 12                         # com, dcom, h current (9)
    def p_comref(self, k): return self.p_glob() + 9 + 3 * k
    def p_href(self, k): return self.p_glob() + 9 + 3 * (self.N + 1) + 3 * k
    def p_extf(self, k): return self.p_glob() + 9 + 6 * (self.N + 1) + 3 * k
    def p_extt(self, k): return self.p_glob() + 9 + 6 * (self.N + 1) + 3 * self.N + 3 * k

    # ---- g
    def g_init(self): return 0                                         # com0, dcom0, h0 (9) then pos_L0, pos_R0 (6)
    def g_com(self, k): return 15 + 3 * k
    def g_dcom(self, k): return 15 + 3 * self.N + 3 * k
    def g_h(self, k): return 15 + 6 * self.N + 3 * k
    def g_pos(self, c, k): return 15 + 9 * self.N + 3 * self.N * c + 3 * k
    def g_box(self, c, k): return 15 + 15 * self.N + c * 19 * self.N + 3 * k
    def g_fric(self, c, j, k): return 15 + 15 * self.N + c * 19 * self.N + 3 * self.N + 16 * k + 4 * j
