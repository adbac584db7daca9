"""numpy emulation of k_assemble / k_factor / k_solve driven by the arrays the
host symbolic analysis produces.  Lets the CPU suite check ordering, fill
closure, schedules and the static 2x2-block pivoting against SuperLU without
a GPU.  Mirrors helmpy_b200/csrc/helm_b200.cu step for step."""
import numpy as np


def assemble(S, ytr_orig, btype_orig):
    """LU slot values [nslots, 2, 2] for bus types given in original numbering."""
    nslots = len(S['slot_src'])
    lu = np.zeros((nslots, 2, 2))
    bt = np.asarray(btype_orig)[S['perm']]
    for slot in range(nslots):
        r = S['slot_row'][slot]
        a = S['slot_src'][slot]
        diag = S['dslot'][r] == slot
        if bt[r] == 2:
            if diag:
                lu[slot] = np.eye(2)
        elif a >= 0:
            y = ytr_orig[S['adj_src'][a]]
            lu[slot, 0] = (y.real, -y.imag)
            if bt[r] == 0:
                lu[slot, 1] = (y.imag, y.real)
            elif diag:
                lu[slot, 1] = (1.0, 0.0)
    return lu


def factor(S, lu):
    N = len(S['perm'])
    lptr, lcol, dslot = S['lptr'], S['lcol'], S['dslot']
    fptr, fsrc, fdst = S['fptr'], S['fsrc'], S['fdst']
    for r in range(N):          # rows in level order == natural order
        for e in range(lptr[r], lptr[r + 1]):
            k = lcol[e]
            m = lu[e] @ lu[dslot[k]]
            lu[e] = m
            for u in range(fptr[e], fptr[e + 1]):
                lu[fdst[u]] -= m @ lu[fsrc[u]]
        lu[dslot[r]] = np.linalg.inv(lu[dslot[r]])
    return lu


def solve(S, lu, rhs_orig):
    """rhs_orig: [2N] in original ordering -> x [2N] original ordering."""
    N = len(S['perm'])
    perm = S['perm']
    nL = len(S['lcol'])
    x = np.zeros((N, 2))
    b = np.asarray(rhs_orig).reshape(N, 2)[perm]
    for r in range(N):
        acc = b[r].copy()
        for e in range(S['lptr'][r], S['lptr'][r + 1]):
            acc -= lu[e] @ x[S['lcol'][e]]
        x[r] = acc
    for q in range(N):
        r = S['urow'][q]
        acc = x[r].copy()
        u0 = S['uptr'][q]
        base = nL + u0 + q
        for t in range(S['uptr'][q + 1] - u0):
            acc -= lu[base + 1 + t] @ x[S['ucol'][u0 + t]]
        x[r] = lu[base] @ acc
    out = np.zeros((N, 2))
    out[perm] = x
    return out.reshape(-1)
