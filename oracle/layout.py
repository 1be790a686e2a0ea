"""oracle/layout.py -- TEST INFRASTRUCTURE (numpy helpers shared by the oracle wrappers and tests).

CONFLUX/COSTA tile layout (reference: src/conflux/lu/layout.cpp:95-123, lu_params.hpp:67-82):
global tile (gi, gj) of size v x v lives on rank (pi, pj) = (gi % Px, gj % Py), layer pk = 0, as local tile
(gi // Px, gj // Py) of a row-major Ml x Nl array.  rank = (pi*Py + pj)*Pz + pk (row-major cart order).
"""
import numpy as np


def dims(N, v, Px, Py, Pz):
    """lu_params::initialize (lu_params.hpp:67-82)."""
    import math
    tx = math.ceil(N / (v * Px))
    ty = math.ceil(N / (v * Py))
    M, NN = v * Px * tx, v * Py * ty
    Mt, Nt = M // v, NN // v
    return dict(M=M, N=NN, Mt=Mt, Nt=Nt, Ml=math.ceil(Mt / Px) * v, Nl=math.ceil(Nt / Py) * v,
                nlayr=(v + Pz - 1) // Pz, P=Px * Py * Pz)


def rank_of(pi, pj, pk, Px, Py, Pz):
    return (pi * Py + pj) * Pz + pk


def assemble(local_all, N, v, Px, Py, Pz):
    """local_all[rank] (Ml x Nl) for layer-0 ranks -> global M x N matrix."""
    d = dims(N, v, Px, Py, Pz)
    G = np.zeros((d["M"], d["N"]))
    for pi in range(Px):
        for pj in range(Py):
            loc = np.asarray(local_all[rank_of(pi, pj, 0, Px, Py, Pz)]).reshape(d["Ml"], d["Nl"])
            for lti in range(d["Ml"] // v):
                gi = lti * Px + pi
                for ltj in range(d["Nl"] // v):
                    gj = ltj * Py + pj
                    G[gi * v:(gi + 1) * v, gj * v:(gj + 1) * v] = loc[lti * v:(lti + 1) * v, ltj * v:(ltj + 1) * v]
    return G


def scatter(G, v, Px, Py, Pz):
    """global matrix -> list of P local arrays (layers pk != 0 are zero, lu_params.hpp:149-155)."""
    M, N = G.shape
    d = dims(N, v, Px, Py, Pz)
    out = [np.zeros((d["Ml"], d["Nl"])) for _ in range(d["P"])]
    for pi in range(Px):
        for pj in range(Py):
            loc = out[rank_of(pi, pj, 0, Px, Py, Pz)]
            for lti in range(d["Ml"] // v):
                gi = lti * Px + pi
                for ltj in range(d["Nl"] // v):
                    gj = ltj * Py + pj
                    loc[lti * v:(lti + 1) * v, ltj * v:(ltj + 1) * v] = G[gi * v:(gi + 1) * v, gj * v:(gj + 1) * v]
    return out


def residual(A_glob, LU_glob, perm):
    """||P A - L U||_F / ||A||_F with P A = A[perm, :] (perm[i] = original row at pivoted position i;
    conflux_opt.hpp:910,1822) and L unit-lower / U upper packed in LU_glob (conflux_miniapp.cpp:349-500)."""
    L = np.tril(LU_glob, -1) + np.eye(LU_glob.shape[0])
    U = np.triu(LU_glob)
    PA = A_glob[np.asarray(perm), :]
    return float(np.linalg.norm(PA - L @ U) / np.linalg.norm(A_glob))
