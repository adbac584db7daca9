// Batched Newton-Raphson cross-check on the GPU: the reference's polar NR
// (helmpy/core/nr.py:244-436; host restatement: helmpy_b200/core/nr.py) for B load-scaling
// scenarios at once, on the same 2x2-block pattern, numeric LU (k_factor_coop) and wave solves
// (k_solve_flat / k_solve_cta) as the HELM path: the Jacobian block of (bus i, bus k) is
//   [ Re dS_i/dth_k   Re dS_i/dV_k * V_k ]      unknowns [ d theta_k, dV_k / V_k ]
//   [ Im dS_i/dth_k   Im dS_i/dV_k * V_k ]      rhs      [ dP_i, dQ_i ]
// with dS/dth = j V_i (delta_ik conj(I_i) - conj(Y_ik V_k)), dS/dV*V = V_i conj(Y_ik V_k) + delta_ik conj(I_i) V_i.
// A PV bus keeps its P row and gets the row [0 1] (dV = 0); the slack gets the identity block.
// Included by helm_b200.cu inside its anonymous namespace.  G = 1 only.
//
// Scratch (aliases of HELM work arrays that NR does not use): w.Pcur = (theta, |V|), w.Vh = complex V,
// w.Hh = complex I = Y V, w.Eh = per-scenario max |F| (double bits), w.g_n = iterations of the pass,
// w.s_fail = "pass converged", w.s_ncoef = iterations per pass.

__global__ void k_nr_init(PlanDev p, WorkDev w) {
    const int grp = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < p.N) {
        const size_t bi = (size_t)grp * p.N + i;
        const unsigned char bt = p.btype0[i];
        w.btype[bi] = bt;
        w.Qg[bi] = 0.0;
        const double vm = (bt != HELM_BUS_PQ) ? p.vsp[i] : 1.0;     // flat start with generator set-points (nr.py:191-230)
        w.Pcur[bi] = make_double2(0.0, vm);
        w.Vh[bi] = make_double2(vm, 0.0);
    }
    if (i == 0) {
        const bool real = grp < w.B;
        w.scale[grp] = real ? w.scale_in[grp] : 1.0;
        w.s_status[grp] = real ? ST_ACTIVE : ST_PAD;
        w.g_n[grp] = 0; w.g_flags[grp] = 0;
        w.s_fail[grp] = 0; w.s_viol[grp] = 0; w.s_npass[grp] = 0; w.s_nswitch[grp] = 0;
        w.o_b[grp] = -1;
        for (int k = 0; k < HELM_MAX_PASSES; ++k) w.s_ncoef[grp * HELM_MAX_PASSES + k] = 0;
        ((unsigned long long *)w.Eh)[grp] = 0ull;
    }
}

// injections, mismatch vector F (into w.rhs) and its max norm (nr.py:343-386)
__global__ void k_nr_eval(PlanDev p, WorkDev w) {
    const int grp = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N || w.s_status[grp] != ST_ACTIVE) return;
    const size_t base = (size_t)grp * p.N, bi = base + i;
    const double2 *Vc = w.Vh + base;
    double2 cur = make_double2(0.0, 0.0);
    for (int a = p.adj_ptr[i]; a < p.adj_ptr[i + 1]; ++a) cur = cadd(cur, cmul(p.yfull[a], Vc[p.adj_idx[a]]));
    w.Hh[bi] = cur;
    const double2 S = cmulc(Vc[i], cur);                                   // V conj(I)
    const double sc = w.scale[grp];
    const unsigned char bt = w.btype[bi];
    double2 F = make_double2(0.0, 0.0);
    if (bt != HELM_BUS_SLACK) {
        F.x = p.pg[i] * sc - p.pd[i] * sc - S.x;
        if (bt == HELM_BUS_PQ) F.y = w.Qg[bi] - p.qd[i] * sc - S.y;
    }
    w.rhs[bi] = F;
    const double e = fmax(fabs(F.x), fabs(F.y));
    if (!(e == 0.0)) atomicMax((unsigned long long *)w.Eh + grp, (unsigned long long)__double_as_longlong(e == e ? e : 1e300));
}

__global__ void k_nr_decide(WorkDev w, double tol, int max_iter) {
    const int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= w.ngroups) return;
    int gf = 0;
    if (w.s_status[grp] == ST_ACTIVE) {
        unsigned long long *eb = (unsigned long long *)w.Eh + grp;
        const double e = __longlong_as_double((long long)*eb);
        *eb = 0ull;
        if (e <= tol) {
            w.s_fail[grp] = 1;                                             // this pass has converged
            gf = GF_PASSEND;
        } else {
            const int it = w.g_n[grp] + 1;
            w.g_n[grp] = it;
            if (it == max_iter + 1) w.s_status[grp] = ST_DIVERGED;        // nr.py:378-381
            else gf = GF_FACTORING;                                       // takes a Newton step this round
        }
    }
    w.g_flags[grp] = gf;
}

// generator limits after a converged pass (nr.py:409-436)
__global__ void k_nr_qlim(PlanDev p, WorkDev w) {
    const int grp = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N || !(w.g_flags[grp] & GF_PASSEND)) return;
    const size_t bi = (size_t)grp * p.N + i;
    if (w.btype[bi] != HELM_BUS_PV) return;
    const double2 S = cmulc(w.Vh[bi], w.Hh[bi]);
    const double qg = S.y + p.qd[i] * w.scale[grp];
    double lim = qg;
    if (qg > p.qmax[i] || qg < p.qmin[i]) {
        lim = (qg > p.qmax[i]) ? p.qmax[i] : p.qmin[i];
        w.btype[bi] = HELM_BUS_PQ;
        w.s_viol[grp] = 1;
        atomicAdd(&w.s_nswitch[grp], 1);
    }
    w.Qg[bi] = lim;
}

__global__ void k_nr_passend(WorkDev w, int enforce_q) {
    const int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= w.ngroups) return;
    int gf = w.g_flags[grp];
    if (gf & GF_PASSEND) {
        const int np = w.s_npass[grp];
        if (np < HELM_MAX_PASSES) w.s_ncoef[grp * HELM_MAX_PASSES + np] = w.g_n[grp] + 1;   // +1: 0 marks an unused entry
        w.s_npass[grp] = np + 1;
        w.g_n[grp] = 0;
        w.s_fail[grp] = 0;
        if (enforce_q && w.s_viol[grp] && np + 1 < 4 * HELM_MAX_PASSES) {
            w.s_viol[grp] = 0;                                             // continue from the current voltages
            gf = 0;
        } else {
            w.s_status[grp] = ST_CONVERGED;
            gf = GF_DONE;
        }
        w.g_flags[grp] = gf;
    }
    if (w.s_status[grp] == ST_ACTIVE) {
        atomicAdd(w.active_groups, 1);
        if (gf & GF_FACTORING) w.fact_list[atomicAdd(&w.fact_count[0], 1)] = grp;
    }
}

// Jacobian blocks into the LU slots of the scenarios that take a step
__global__ void __launch_bounds__(256) k_nr_assemble(PlanDev p, WorkDev w) {
    const int nfact = w.fact_count[0];
    for (int fi = blockIdx.x; fi < nfact; fi += gridDim.x) {
        const int grp = w.fact_list[fi];
        const size_t base = (size_t)grp * p.N;
        const double2 *Vc = w.Vh + base, *Ic = w.Hh + base;
        double2 *lu = w.lu + (size_t)grp * p.nslots * 2;
        for (int slot = threadIdx.x; slot < p.nslots; slot += blockDim.x) {
            const int r = p.slot_row[slot], c = p.slot_col[slot], a = p.slot_src[slot];
            const unsigned char bt = w.btype[base + r];
            const bool diag = (r == c);
            double2 r0 = make_double2(0.0, 0.0), r1 = make_double2(0.0, 0.0);
            if (bt == HELM_BUS_SLACK) {
                if (diag) { r0.x = 1.0; r1.y = 1.0; }
            } else if (a >= 0) {
                const double2 vr = Vc[r];
                const double2 yv = cmul(p.yfull[a], Vc[c]);
                const double2 t = make_double2(yv.x, -yv.y);                // conj(Y_rc V_c)
                double2 d = make_double2(-t.x, -t.y);
                double2 dv = cmul(vr, t);
                if (diag) {
                    const double2 ci = make_double2(Ic[r].x, -Ic[r].y);    // conj(I_r)
                    d = cadd(d, ci);
                    dv = cadd(dv, cmul(ci, vr));
                }
                const double2 vd = cmul(vr, d);
                const double2 dth = make_double2(-vd.y, vd.x);             // j V_r (delta conj(I_r) - conj(Y V_c))
                r0 = make_double2(dth.x, dv.x);
                if (bt == HELM_BUS_PQ) r1 = make_double2(dth.y, dv.y);
                else if (diag) r1 = make_double2(0.0, 1.0);                // PV: dV = 0
            }
            lu[2 * (size_t)slot] = r0;
            lu[2 * (size_t)slot + 1] = r1;
        }
    }
}

// theta += d theta, |V| *= 1 + dV/V (nr.py:390-406)
__global__ void k_nr_update(PlanDev p, WorkDev w) {
    const int grp = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N || !(w.g_flags[grp] & GF_FACTORING) || w.s_status[grp] != ST_ACTIVE) return;
    const size_t bi = (size_t)grp * p.N + i;
    const unsigned char bt = w.btype[bi];
    if (bt == HELM_BUS_SLACK) return;
    const double2 dx = w.xb[bi];
    double2 pol = w.Pcur[bi];
    pol.x += dx.x;
    if (bt == HELM_BUS_PQ) pol.y *= 1.0 + dx.y;
    w.Pcur[bi] = pol;
    double sn, cs;
    sincos(pol.x, &sn, &cs);
    w.Vh[bi] = make_double2(pol.y * cs, pol.y * sn);
}

__global__ void k_nr_out(PlanDev p, WorkDev w, double2 *vout, int *out_status, int *out_iters, int *out_nswitch) {
    const int grp = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= w.B) return;
    const int st = w.s_status[grp];
    if (i < p.N) vout[(size_t)grp * p.N + p.perm[i]] = (st == ST_CONVERGED) ? w.Vh[(size_t)grp * p.N + i] : make_double2(0.0, 0.0);
    if (i == 0) {
        out_status[grp] = (st == ST_CONVERGED) ? HELM_ST_CONVERGED : (st == ST_SINGULAR) ? HELM_ST_SINGULAR : HELM_ST_DIVERGED;
        if (out_nswitch) out_nswitch[grp] = w.s_nswitch[grp];
        if (out_iters)
            for (int k = 0; k < HELM_MAX_PASSES; ++k) out_iters[grp * HELM_MAX_PASSES + k] = w.s_ncoef[grp * HELM_MAX_PASSES + k];
    }
}
