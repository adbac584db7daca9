// HELM variants beyond the default (pv_bus_model 2, classic slack): PV bus model 1 and the
// distributed-slack formulations, methods 1 and 2 (reference helm.py:62-103, 142-290, 493-525,
// 584-597).  Included by helm_b200.cu; G = 1 only.
//
// PV model 1: the unknown in slot 2i of a PV bus is the reactive-power coefficient Q_i[n]; Re V_i[n]
// is known before the solve from |V|^2 = Vsp^2 (Calculo_Vre_PV, helm.py:142-159).  In the matrix the
// first column of every block in a PV bus's block column is zero (moved to the right-hand side,
// helm.py:72-103) and the diagonal block gets [1][0] = 1.
//
// Distributed slack: one extra unknown Ploss[n] (helm.py:62-70) = a one-row / one-column border
// around the 2N x 2N block matrix.  It is solved as a bordered system:
//     [A c; r d] [x; p] = [b; beta]   ->   y = A^-1 b,  z = A^-1 c (once per factorization),
//     p = (beta - r.y) / (d - r.z),  x = y - z p
// so the block factorization and the triangular-solve kernels are reused unchanged.
#pragma once

// K_i of reference compute_k_factor / K_slack_1 (helm.py:493-525); `ktot` = sum of the
// contributing generations (unscaled: the load scale cancels in the ratio).
__device__ __forceinline__ double var_kfactor(const PlanDev &p, const WorkDev &w, int i, unsigned char bt,
                                              double ktot) {
    if (!w.dsb_model) return (bt == HELM_BUS_SLACK) ? 1.0 : 0.0;
    if (bt == HELM_BUS_PV) return p.pg[i] > 0.0 ? p.pg[i] / ktot : 0.0;
    if (bt == HELM_BUS_SLACK) return p.pg_imb0 > 0.0 ? p.pg_imb0 / ktot : 0.0;
    return 0.0;
}

// ---- side branch, after the factorization: participation factors and the border column c ---------
// One CTA per scenario.  c has -K_i in row 2i of every PV bus (helm.py:67-69); it is written into
// the right-hand-side buffer so that the next (side-branch) solve produces z = A^-1 c.
__global__ void __launch_bounds__(256) k_border_prep(PlanDev p, WorkDev w) {
    __shared__ double red[256];
    const int grp = blockIdx.x;
    if (!(w.g_flags[grp] & GF_FACTORING)) return;
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int N = p.N;
    const unsigned char *bt = w.btype + (size_t)grp * N;
    double acc = 0.0;
    if (w.dsb_model)
        for (int i = threadIdx.x; i < N; i += blockDim.x)
            if (bt[i] == HELM_BUS_PV && p.pg[i] > 0.0) acc += p.pg[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    double ktot = red[0] + (p.pg_imb0 > 0.0 ? p.pg_imb0 : 0.0);
    if (threadIdx.x == 0) w.ktot[grp] = ktot;
    double2 *rhs = w.rhs + (size_t)grp * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double k = (bt[i] == HELM_BUS_PV) ? var_kfactor(p, w, i, bt[i], ktot) : 0.0;
        rhs[i] = make_double2(-k, 0.0);
    }
}

// r.v for the border row r = row 2N of the matrix: (G_sj, -B_sj) over adj(slack) (helm.py:63-66);
// with PV model 1 the entry multiplying slot 2j of a PV bus j was moved to the RHS (helm.py:96-98).
__device__ __forceinline__ double border_row_dot(const PlanDev &p, const WorkDev &w, int grp, const double2 *v) {
    const int s = p.slack;
    const unsigned char *bt = w.btype + (size_t)grp * p.N;
    double acc = 0.0;
    for (int a = p.adj_ptr[s]; a < p.adj_ptr[s + 1]; ++a) {
        int j = p.adj_idx[a];
        double2 y = p.ytr[a];
        double2 vj = v[j];
        double g = (w.pv_model == 1 && bt[j] == HELM_BUS_PV) ? 0.0 : y.x;
        acc += g * vj.x - y.y * vj.y;
    }
    return acc;
}

// ---- side branch, after the z-solve: keep z and the Schur complement d - r.z -----------------------
__global__ void __launch_bounds__(256) k_border_fin(PlanDev p, WorkDev w) {
    const int grp = blockIdx.x;
    if (!(w.g_flags[grp] & GF_FACTORING)) return;
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int N = p.N;
    const double2 *xb = w.xb + (size_t)grp * N;
    double2 *zb = w.zb + (size_t)grp * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) zb[i] = xb[i];
    if (threadIdx.x == 0) {
        double ks = var_kfactor(p, w, p.slack, HELM_BUS_SLACK, w.ktot[grp]);
        w.sden[grp] = -ks - border_row_dot(p, w, grp, xb);
        w.ploss[(size_t)grp * w.K] = 0.0;
    }
}

// ---- main branch, after the solve: Ploss coefficient of this order ------------------------------------
__global__ void k_border_p(PlanDev p, WorkDev w) {
    int grp = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp >= w.ngroups) return;
    if (w.g_flags[grp] & (GF_DONE | GF_FACTORING)) return;
    if (w.s_status[grp] != ST_ACTIVE) return;
    const double2 *xb = w.xb + (size_t)grp * p.N;
    double pl = (w.beta[grp] - border_row_dot(p, w, grp, xb)) / w.sden[grp];
    w.pcur[grp] = pl;
    w.ploss[(size_t)grp * w.K + w.g_n[grp] + 1] = pl;
}

// ---- main branch: x = y - z p (distributed slack); PV model 1: split the PV unknown --------------------
// After this kernel xb holds V_i[j] for every bus and qcur holds Q_i[j] of the PV buses.
__global__ void k_var_pre(PlanDev p, WorkDev w) {
    const int grp = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    if (w.g_flags[grp] & (GF_DONE | GF_FACTORING)) return;
    if (w.s_status[grp] != ST_ACTIVE) return;
    const size_t bi = (size_t)grp * p.N + i;
    double2 v = w.xb[bi];
    if (w.dsb_method) {
        double pl = w.pcur[grp];
        double2 z = w.zb[bi];
        v.x -= z.x * pl;
        v.y -= z.y * pl;
    }
    if (w.pv_model == 1 && w.btype[bi] == HELM_BUS_PV) {
        w.qcur[bi] = v.x;                                  // coefficients[2i][n] = Q_i[n]
        v.x = w.vre[bi];                                   // helm.py:416-418
    }
    w.xb[bi] = v;
}

// phase-shifter sum over the neighbours' newest coefficients (helm.py:375-377)
__device__ __forceinline__ double2 var_phase_new(const PlanDev &p, const double2 *xb, int i) {
    double2 pp = make_double2(0.0, 0.0);
    for (int a = p.ph_ptr[i]; a < p.ph_ptr[i + 1]; ++a) pp = cadd(pp, cmul(p.ph_val[a], xb[p.ph_idx[a]]));
    return pp;
}

// PV-model-2 style P equation (helm.py:293-364 for PV buses, 191-258 for the slack, method 2).
// Returns CC; stores barras_CC[j] and (PV buses) the |V|^2 convolution through *vv_out.
__device__ double var_pv2_cc(const PlanDev &p, const WorkDev &w, int grp, int i, int j, double2 v,
                             const double2 *xb, double2 *Vb, double2 *Hb, size_t plane, bool is_slack,
                             double *vv_out) {
    const double2 zero = make_double2(0.0, 0.0);
    double2 PP = zero;
    for (int a = p.adj_ptr[i]; a < p.adj_ptr[i + 1]; ++a) PP = cadd(PP, cmul(p.ytr[a], xb[p.adj_idx[a]]));
    Hb[(size_t)j * plane] = PP;
    double2 ppp = zero;
    for (int x = 1; x <= j - 1; ++x) {
        double2 vv = (x == 1) ? v : Vb[(size_t)(j + 1 - x) * plane];
        ppp = cadd(ppp, cmulc(Hb[(size_t)x * plane], vv));
    }
    double2 v1 = (j == 1) ? v : Vb[plane];
    ppp = cadd(ppp, cmulc(PP, v1));
    double cc = 0.0 - ppp.x;
    double vvre = 0.0;                                     // sum_{k=1}^{j} V[k] conj(V[j+1-k])
    for (int k = 1; k <= j; ++k) {
        double2 a = (k == j) ? v : Vb[(size_t)k * plane];
        double2 b = (j + 1 - k == j) ? v : Vb[(size_t)(j + 1 - k) * plane];
        vvre += a.x * b.x + a.y * b.y;
    }
    double2 ysh = p.yshunt[i];
    const size_t bi = (size_t)grp * p.N + i;
    if (ysh.x != 0.0) {
        if (is_slack) {
            // helm.py:219-222 as written: VV = sum_{k=1}^{n-2} V[k] conj(V[n-k]), n = j+1
            double vq = 0.0;
            for (int k = 1; k <= j - 1; ++k) {
                double2 a = Vb[(size_t)k * plane];
                double2 b = (j + 1 - k == j) ? v : Vb[(size_t)(j + 1 - k) * plane];
                vq += a.x * b.x + a.y * b.y;
            }
            if (j >= 2) cc -= ysh.x * (vq + 2.0 * v.x);
            else cc -= ysh.x * 2.0 * v.x;
        } else {
            if (j >= 2) cc -= ysh.x * (w.VVant[bi] + 2.0 * v.x);
            else cc -= ysh.x * 2.0 * v.x;
        }
    }
    if (p.ph_ptr[i + 1] > p.ph_ptr[i]) {
        double2 acc = zero;
        for (int x = 0; x <= j; ++x) {
            double2 pp = zero;
            for (int a = p.ph_ptr[i]; a < p.ph_ptr[i + 1]; ++a) {
                int k = p.ph_idx[a];
                double2 vk = (x == 0) ? xb[k] : w.Vh[((size_t)grp * w.K + (j - x)) * p.N + k];
                pp = cadd(pp, cmul(p.ph_val[a], vk));
            }
            double2 vx = (x == j) ? v : Vb[(size_t)x * plane];
            acc = cadd(acc, cmulc(pp, vx));
        }
        cc -= acc.x;
    }
    if (!is_slack) w.VVant[bi] = vvre;
    *vv_out = vvre;
    return cc;
}

// ---- K1 for the variants: one thread per (bus, scenario), plain loops (clarity over speed) -------------
__global__ void __launch_bounds__(256) k_rhs_conv_var(PlanDev p, WorkDev w) {
    const int grp = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    const int gf = w.g_flags[grp];
    if (gf & (GF_DONE | GF_PASSEND | GF_FACTORING)) return;
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int N = p.N;
    const int j = w.g_n[grp] + 1;
    const size_t plane = (size_t)N;
    const size_t bi = (size_t)grp * N + i;
    const size_t h0 = ((size_t)grp * w.K) * N + i;
    double2 *Vb = w.Vh + h0, *Hb = w.Hh + h0, *Eb = w.Eh + h0;
    const double2 *xb = w.xb + (size_t)grp * N;
    const unsigned char bt = w.btype[bi];
    const double2 v = xb[i];
    const double2 zero = make_double2(0.0, 0.0);
    const double sc = w.scale[grp];
    const double ktot = w.dsb_method ? w.ktot[grp] : 1.0;
    const double *pl = w.dsb_method ? w.ploss + (size_t)grp * w.K : nullptr;
    Vb[(size_t)j * plane] = v;
    const bool pv1 = (bt == HELM_BUS_PV && w.pv_model == 1);
    const bool need_w = (bt == HELM_BUS_PQ) || pv1 || (bt == HELM_BUS_SLACK && w.dsb_method == 1);
    double2 wj = zero;
    if (need_w) {                                          // helm.py:423-433
        double2 aux = zero;
        for (int k = 0; k < j; ++k) {
            double2 vk = (k == 0) ? v : Vb[(size_t)(j - k) * plane];
            aux = cadd(aux, cmul(Hb[(size_t)k * plane], vk));
        }
        wj = make_double2(-aux.x, -aux.y);
        Hb[(size_t)j * plane] = wj;
    }
    // sum_{k=1}^{j} Ploss[k] conj(W[j+1-k])  (helm.py:177-179, 277-280)
    double2 auxP = zero;
    if (w.dsb_method && (pv1 || (bt == HELM_BUS_SLACK && w.dsb_method == 1))) {
        for (int k = 1; k <= j; ++k) {
            double2 wk = (j + 1 - k == j) ? wj : Hb[(size_t)(j + 1 - k) * plane];
            auxP.x += pl[k] * wk.x;
            auxP.y -= pl[k] * wk.y;
        }
    }
    if (bt == HELM_BUS_SLACK) {
        w.rhs[bi] = zero;
        if (w.dsb_method == 1) {                           // helm.py:164-188
            double Pi = (w.dsb_model ? p.pg_imb0 * sc : 0.0) - p.pd[i] * sc;
            double2 pp = var_phase_new(p, xb, i);
            double2 ysv = cmul(p.yshunt[i], v);
            double ks = var_kfactor(p, w, i, bt, ktot);
            w.beta[grp] = Pi * wj.x - ysv.x - pp.x + ks * auxP.x;
        } else if (w.dsb_method == 2) {                    // helm.py:191-258
            double vv;
            w.beta[grp] = var_pv2_cc(p, w, grp, i, j, v, xb, Vb, Hb, plane, true, &vv);
        }
        return;
    }
    // epsilon anti-diagonal + Pade check (same as the default kernel)
    EpsState es;
    es.prev1 = Eb[0];
    es.prev2 = zero;
    es.cur = cadd(es.prev1, v);
    Eb[0] = es.cur;
    for (int t = 1; t <= j; ++t) {
        double2 eo = (t < j) ? Eb[(size_t)t * plane] : zero;
        eps_step(es, t, j, v, eo, Eb, plane);
    }
    double2 out;
    if (bt == HELM_BUS_PQ) {                               // helm.py:367-385
        double Pi = p.pg[i] * sc - p.pd[i] * sc;
        double Qi = w.Qg[bi] - p.qd[i] * sc;
        double2 pp = var_phase_new(p, xb, i);
        double2 tt = cmul(make_double2(Pi, -Qi), make_double2(wj.x, -wj.y));
        double2 ysv = cmul(p.yshunt[i], v);
        out = make_double2(tt.x - ysv.x - pp.x, tt.y - ysv.y - pp.y);
    } else if (!pv1) {                                     // PV model 2 (helm.py:293-364)
        double vv;
        double cc = var_pv2_cc(p, w, grp, i, j, v, xb, Vb, Hb, plane, false, &vv);
        out = make_double2(cc, -vv * 0.5);
    } else {                                               // PV model 1 (helm.py:261-290)
        double *Qb = w.Qh + h0;
        Qb[(size_t)j * plane] = w.qcur[bi];
        double2 aux = zero;                                // sum_{k=1}^{j} Q[k] conj(W[j+1-k])
        for (int k = 1; k <= j; ++k) {
            double qk = Qb[(size_t)k * plane];
            double2 wk = (j + 1 - k == j) ? wj : Hb[(size_t)(j + 1 - k) * plane];
            aux.x += qk * wk.x;
            aux.y -= qk * wk.y;
        }
        double Pi = p.pg[i] * sc - p.pd[i] * sc;
        double2 pp = var_phase_new(p, xb, i);
        double2 ysv = cmul(p.yshunt[i], v);
        double ki = w.dsb_method ? var_kfactor(p, w, i, bt, ktot) : 0.0;
        // P conj(W[j]) - Yshunt V[j] - phase - j*aux + K*auxP
        out = make_double2(Pi * wj.x - ysv.x - pp.x + aux.y + ki * auxP.x,
                           -Pi * wj.y - ysv.y - pp.y - aux.x + ki * auxP.y);
        double vvre = 0.0;                                 // Calculo_Vre_PV for order j+1 (helm.py:152-157)
        for (int k = 1; k <= j; ++k) {
            double2 a = (k == j) ? v : Vb[(size_t)k * plane];
            double2 b = (j + 1 - k == j) ? v : Vb[(size_t)(j + 1 - k) * plane];
            vvre += a.x * b.x + a.y * b.y;
        }
        w.vre[bi] = -0.5 * vvre;
    }
    w.rhs[bi] = out;
    if ((j & 1) == 0) {
        double2 cur = es.cur;
        if (j >= 4) {
            double2 old = w.Pcur[bi];
            double mag1 = hypot(old.x, old.y), rad1 = atan2(old.y, old.x);
            double mag2 = hypot(cur.x, cur.y), rad2 = atan2(cur.y, cur.x);
            bool ok = (fabs(mag1 - mag2) <= w.mismatch) && (fabs(rad1 - rad2) <= w.mismatch);
            if (!ok) w.s_fail[grp] = 1;
        }
        w.Pcur[bi] = cur;
    }
}

// ---- order-0 germ and RHS of order 1 for the variants (helm.py:110-139, n == 1 branches) --------------
__global__ void k_init_var(PlanDev p, WorkDev w) {
    const int grp = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    if (!(w.g_flags[grp] & GF_FACTORING)) return;
    if (w.s_status[grp] != ST_ACTIVE) return;
    const int N = p.N;
    const size_t bi = (size_t)grp * N + i;
    const size_t h0 = ((size_t)grp * w.K) * N + i;
    const double sc = w.scale[grp];
    const unsigned char bt = w.btype[bi];
    const double2 one = make_double2(1.0, 0.0);
    w.Vh[h0] = one;
    w.Hh[h0] = one;
    w.Eh[h0] = one;
    w.VVant[bi] = 0.0;
    if (w.pv_model == 1) { w.Qh[h0] = 0.0; w.vre[bi] = 0.0; w.qcur[bi] = 0.0; }
    double2 pp = make_double2(0.0, 0.0);
    for (int a = p.ph_ptr[i]; a < p.ph_ptr[i + 1]; ++a) pp = cadd(pp, p.ph_val[a]);
    const double2 ysh = p.yshunt[i];
    double2 r;
    if (bt == HELM_BUS_SLACK) {
        r = make_double2(p.vsp[i] - 1.0, 0.0);
        w.Pcur[bi] = make_double2(p.vsp[i], 0.0);
        if (w.dsb_method) {                                // helm.py:187 / 225-230 at n = 1
            double Pi = (w.dsb_model ? p.pg_imb0 * sc : 0.0) - p.pd[i] * sc;
            w.beta[grp] = Pi - ysh.x - pp.x;
        }
    } else if (bt == HELM_BUS_PQ) {
        double Pi = p.pg[i] * sc - p.pd[i] * sc;
        double Qi = w.Qg[bi] - p.qd[i] * sc;
        r = make_double2(Pi - ysh.x - pp.x, -Qi - ysh.y - pp.y);
    } else if (w.pv_model == 2) {
        double vs = p.vsp[i];
        r = make_double2((p.pg[i] * sc - p.pd[i] * sc) - ysh.x - pp.x, (vs * vs - 1.0) * 0.5);
    } else {
        double vs = p.vsp[i];
        double Pi = p.pg[i] * sc - p.pd[i] * sc;
        r = make_double2(Pi - ysh.x - pp.x, -ysh.y - pp.y);
        w.vre[bi] = (vs * vs - 1.0) * 0.5;                 // Vre_PV[i][1], helm.py:158-159
    }
    w.rhs[bi] = r;
}

// ---- PV model 1: move the removed matrix columns to the right-hand side (helm.py:584-597) --------------
// rhs[row] -= Vre_PV[j][n] * A[row][2j] for every PV bus j.  side = 1: scenarios being factorized
// (order 1, after k_init_var); side = 0: scenarios of the main branch (after k_rhs_conv_var).
__global__ void k_pv1_correct(PlanDev p, WorkDev w, int side) {
    const int grp = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    const int gf = w.g_flags[grp];
    if (side ? !(gf & GF_FACTORING) : (gf & (GF_DONE | GF_PASSEND | GF_FACTORING)) != 0) return;
    if (w.s_status[grp] != ST_ACTIVE) return;
    const size_t base = (size_t)grp * p.N;
    const unsigned char bt = w.btype[base + i];
    double sx = 0.0, sy = 0.0;
    for (int a = p.adj_ptr[i]; a < p.adj_ptr[i + 1]; ++a) {
        int j = p.adj_idx[a];
        if (w.btype[base + j] != HELM_BUS_PV) continue;
        double2 y = p.ytr[a];
        double vr = w.vre[base + j];
        sx += y.x * vr;                                    // A[2i][2j]   = Re Ytrans_ij
        sy += y.y * vr;                                    // A[2i+1][2j] = Im Ytrans_ij
    }
    if (bt == HELM_BUS_SLACK) {
        if (w.dsb_method) w.beta[grp] -= sx;               // A[2N][2j] = Re Ytrans_sj (helm.py:96-98)
        return;
    }
    double2 r = w.rhs[base + i];
    w.rhs[base + i] = make_double2(r.x - sx, r.y - sy);
}
