"""TEST INFRASTRUCTURE ONLY -- full-size per-bond fixtures from the *reference's own code*.

Run in the build container (where /root/reference exists):
    python oracle/make_golden_full.py [cfg5] [cfg4] [cfg3] [cfg1]

For the BASELINE configurations that are too large for tensor fixtures (cfg5: 1.9 GB of blocks) this
stores what decides the integer part of the parity gate -- for *every* bond of the chain the reference's
  SchmidtVectors.from_correlation_matrix(C, x, trunc, which=...)        (slater.py:702-755)
as called by C_to_MPS (slater.py:1293-1296 centre "LR", :1303-1305 "R", :1328-1330 "L"):
  k, n_filled("L")                    per bond
  e           mode eigenvalues        concatenated (offsets e_off)
  lam         schmidt_values (un-normalised, reference order)  concatenated (offsets chi_off)
  masks       occupation of the entangled modes on the left (bit i = mode i), uint64
  sec_q / sec_n   charge sectors (idx_L) as (charge, count) pairs (offsets sec_off)
The correlation matrix itself is not stored: it is regenerated from the (deterministic) Hamiltonian by the
test, and its checksum is.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def hopping_chain(L):
    H = np.zeros((L, L))
    i = np.arange(L - 1)
    H[i, i + 1] = H[i + 1, i] = -1.0
    return H


def cylinder_hamiltonian(Lx, Ly, t=-1.0):
    L = Lx * Ly
    H = np.zeros((L, L))
    for x in range(Lx):
        for y in range(Ly):
            i = x * Ly + y
            j = x * Ly + (y + 1) % Ly
            H[i, j] = H[j, i] = t
            if x + 1 < Lx:
                j = (x + 1) * Ly + y
                H[i, j] = H[j, i] = t
    return H


def entangled_sets(S):
    """bool (chi, k): entangled mode i occupied on the left, from the reference's embedded tables
    (slater.py:430-470: right_sets[:, entangled] = ~sets[:, ::-1])."""
    m = S.modes
    if S.left_sets is not None:
        return np.asarray(S.left_sets[:, m.ixL["entangled"]], dtype=bool)
    return ~np.asarray(S.right_sets[:, m.ixR["entangled"]], dtype=bool)[:, ::-1]


def dump_bonds(ref, name, C, N, trunc, oc=None):
    sl = ref.slater
    L = len(C)
    oc = oc or L // 2
    t0 = time.time()
    ks, fls, es, lams, masks, secq, secn = [], [], [], [], [], [], []
    for x in range(L + 1):
        which = "LR" if x == oc else ("R" if x > oc else "L")
        if x == 0:
            which = "L"          # C_to_MPS reaches bond 0 through the left sweep
        if x == L and oc != L:
            which = "R"
        S = sl.SchmidtVectors.from_correlation_matrix(C, x, trunc, which=which)
        sets = entangled_sets(S)
        k = S.n_entangled
        assert k <= 64
        w = (np.uint64(1) << np.arange(k, dtype=np.uint64))
        ks.append(k)
        fls.append(S.modes.n_filled("L"))
        es.append(np.asarray(S.modes.e, dtype=np.float64))
        lams.append(np.asarray(S.schmidt_values, dtype=np.float64))
        masks.append((sets.astype(np.uint64) * w[None, :]).sum(axis=1, dtype=np.uint64) if k else
                     np.zeros(len(sets), np.uint64))
        qs = sorted(S.idx_L.items(), key=lambda kv: kv[1].start)
        secq.append(np.array([q for q, _ in qs], dtype=np.int32))
        secn.append(np.array([s.stop - s.start for _, s in qs], dtype=np.int32))
        if x % 64 == 0:
            print(f"  {name}: bond {x}/{L}  chi {len(lams[-1])}  k {k}  {time.time() - t0:.0f}s", flush=True)
    off = lambda parts: np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
    data = dict(L=L, N=N, oc=oc, chi_max=-1 if trunc.get("chi_max") is None else trunc["chi_max"],
                svd_min=trunc.get("svd_min", 1e-6), C_sum=float(np.sum(C)), C_abs_sum=float(np.abs(C).sum()),
                k=np.array(ks, np.int32), filled_left=np.array(fls, np.int32),
                e_off=off(es), e=np.concatenate(es), chi_off=off(lams), lam=np.concatenate(lams),
                masks=np.concatenate(masks), sec_off=off(secq), sec_q=np.concatenate(secq),
                sec_n=np.concatenate(secn))
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    chi = np.diff(data["chi_off"])
    print(name, "L", L, "N", N, "max chi", chi.max(), "mean chi", chi.mean(), f"{time.time() - t0:.0f}s")


def main(which):
    ref = ref_shim.load("pass")
    sl = ref.slater
    if "cfg1" in which:
        C, N = sl.correlation_matrix(hopping_chain(64))
        dump_bonds(ref, "bonds_cfg1_chain_L64", C, N, {"chi_max": 64})
    if "cfg3" in which:
        # input of the Gutzwiller projection: spinful "PH" chain, 256 spins -> 512 fermion sites (slater.py:1274-1281)
        C1, N1 = sl.correlation_matrix(hopping_chain(256))
        C = sl.spinful_correlation_matrix(C1, True)
        dump_bonds(ref, "bonds_cfg3_spinful_ph_L512", C, int(round(np.trace(C).real)), {"chi_max": 256})
    if "cfg4" in which:
        C, N = sl.correlation_matrix(cylinder_hamiltonian(64, 6))
        dump_bonds(ref, "bonds_cfg4_cylinder_6x64", C, N, {"chi_max": 1024})
    if "cfg5" in which:
        C, N = sl.correlation_matrix(hopping_chain(1024))
        dump_bonds(ref, "bonds_cfg5_chain_L1024", C, N, {"chi_max": 1024, "svd_min": 1e-7})


if __name__ == "__main__":
    main(sys.argv[1:] or ["cfg1", "cfg3", "cfg4", "cfg5"])
