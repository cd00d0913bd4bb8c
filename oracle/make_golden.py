"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the *reference's own code*.

Run in the build container (where /root/reference exists):  python oracle/make_golden.py
The reference (temfpy/temfpy) is imported unmodified through oracle/ref_shim.py (tenpy / pfapack
mocked, see that file); every number stored below is produced by reference functions:
  SchmidtVectors.from_correlation_matrix  (slater.py:702-755)  -> e, schmidt_values, sets, idx_L
  MPSTensorData.from_schmidt_vectors      (slater.py:975-1104) -> sometimes_matrix, det_always,
                                                                  new_sets_bra / new_sets_ket
  _tensor_block                           (slater.py:828-869)  -> block values
Only the TeNPy packing (to_npc_array) is not executable here; the fixtures therefore store the
per-charge-block arrays, which is what to_npc_array writes into the npc.Array (slater.py:1132-1141).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def random_hamiltonian(L, seed, decay=2.0, cplx=False):
    """Pattern of the reference's examples/slater.py:15-20 (random, exponentially decaying)."""
    rng = np.random.default_rng(seed)
    H = rng.normal(size=(L, L))
    if cplx:
        H = H + 1j * rng.normal(size=(L, L))
    H = H + H.conj().T
    d = np.abs(np.subtract.outer(np.arange(L), np.arange(L)))
    return H * np.exp(-d / decay)


def hopping_chain(L):
    H = np.zeros((L, L))
    i = np.arange(L - 1)
    H[i, i + 1] = H[i + 1, i] = -1.0
    return H


def dump_case(ref, name, H, trunc, N=None):
    sl = ref.slater
    C, N = sl.correlation_matrix(H, N)
    L = len(C)
    oc = L // 2
    data = dict(H=H, C=C, N=N, L=L, oc=oc,
                chi_max=-1 if trunc.get("chi_max") is None else trunc["chi_max"],
                svd_min=trunc.get("svd_min", 1e-6))
    centre = sl.SchmidtVectors.from_correlation_matrix(C, oc, trunc)

    def put_bond(x, S):
        data[f"bond{x}_e"] = S.modes.e
        data[f"bond{x}_lam"] = S.schmidt_values
        q = np.zeros(S.n_schmidt, dtype=np.int64)
        for charge, slc in S.idx_L.items():
            q[slc] = charge
        data[f"bond{x}_charge"] = q

    def put_site(i, td):
        data[f"site{i}_S"] = td.sometimes_matrix
        data[f"site{i}_det"] = np.asarray(td.det_always)
        data[f"site{i}_sets_bra"] = td.new_sets_bra
        data[f"site{i}_sets_ket"] = td.new_sets_ket
        data[f"site{i}_qtotal"] = td.qtotal
        nb = 0
        # the reference's block loop (slater.py:1132-1141); bra rows of a pipe charge are contiguous
        chi_b = len(td.new_sets_bra) // 2
        q_bra_of_alpha = np.zeros(chi_b, dtype=np.int64)
        for charge, slc in td.idx_bra.items():
            q_bra_of_alpha[slc] = charge
        if td.mode == "left":
            q_rows = np.sort(np.concatenate([q_bra_of_alpha, q_bra_of_alpha + 1]), kind="stable")
        else:
            q_rows = np.sort(np.concatenate([q_bra_of_alpha, q_bra_of_alpha - 1]), kind="stable")
        qc = 1 if td.mode == "left" else -1
        for q_ket, slice_ket in td.idx_ket.items():
            rows = np.flatnonzero(q_rows == q_ket + td.qtotal * qc)
            if rows.size == 0:
                continue
            blk = td.det_always * sl._tensor_block(td.sometimes_matrix, td.new_sets_bra[rows],
                                                   td.new_sets_ket[slice_ket])
            data[f"site{i}_blk{nb}"] = blk
            data[f"site{i}_blk{nb}_meta"] = np.array([q_ket, rows[0], rows.size, slice_ket.start,
                                                      slice_ket.stop - slice_ket.start])
            nb += 1
        data[f"site{i}_nblocks"] = nb

    put_bond(oc, centre)
    prev = centre
    for i in range(oc, L):
        new = sl.SchmidtVectors.from_correlation_matrix(C, i + 1, trunc, which="R")
        put_bond(i + 1, new)
        put_site(i, sl.MPSTensorData.from_schmidt_vectors(new, prev, "right"))
        prev = new
    prev = centre
    for i in reversed(range(oc)):
        new = sl.SchmidtVectors.from_correlation_matrix(C, i, trunc, which="L")
        put_bond(i, new)
        put_site(i, sl.MPSTensorData.from_schmidt_vectors(new, prev, "left"))
        prev = new
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print(name, "L", L, "N", N, "max chi", max(len(data[f"bond{x}_lam"]) for x in range(L + 1)))


def dump_lowest_sums(ref):
    """Known-answer vectors for schmidt_utils.lowest_sums (schmidt_utils.py:211-324)."""
    su = ref.schmidt_utils
    rng = np.random.default_rng(7)
    out = {}
    cases = [(0, dict(chi_max=5), None, None), (1, dict(chi_max=5), 3, None),
             (6, dict(chi_max=20), 2, 4), (10, dict(chi_max=50, svd_min=1e-3), 0, None),
             (12, dict(svd_min=1e-2), None, 1), (9, dict(chi_max=30, sectors="two"), 1, 3),
             (14, dict(chi_max=200, svd_min=1e-5, sectors="one"), 0, 0)]
    for c, (k, tp, fl, fr) in enumerate(cases):
        a = rng.normal(size=k) * 3
        if isinstance(tp.get("sectors"), str):
            # the reference raises IndexError when the lowest set is filtered out, so
            # the golden cases keep the charge of the lowest set among the allowed sectors
            q0 = fl + int((a < 0).sum())
            tp = dict(tp, sectors=[q0, q0 + 1] if tp["sectors"] == "two" else q0)
        sums, sets = su.lowest_sums(a, su.to_stopping_condition(tp), filled_left=fl, filled_right=fr)
        out[f"c{c}_a"] = a
        out[f"c{c}_sums"] = sums
        out[f"c{c}_sets"] = np.asarray(sets).reshape(len(sums), k)
        out[f"c{c}_par"] = np.array([tp.get("chi_max", -1) or -1, tp.get("svd_min", 1e-6),
                                     -1 if fl is None else fl, -1 if fr is None else fr])
        sec = tp.get("sectors")
        out[f"c{c}_sectors"] = np.array([] if sec is None else np.atleast_1d(sec), dtype=np.int64)
        out[f"c{c}_has_sectors"] = np.array(sec is not None)
    out["ncases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "lowest_sums.npz"), **out)
    print("lowest_sums cases", len(cases))


def dump_pfaffian_case(ref, name, H, trunc, basis="C"):
    """Pfaffian path: every number below comes from the reference's own functions
    (pfaffian.py: correlation_matrix :302, SchmidtVectors.from_correlation_matrix :1216,
    MPSTensorData.from_schmidt_vectors :1578, _tensor_block :1429).  The pfapack routine
    (pfaffian.py:1425, absent from the image) is bound to oracle/pfaffian_oracle.pfaffian."""
    import pfaffian_oracle as po
    rp = ref.pfaffian
    rp.cpf = lambda A, **kw: po.pfaffian(A)
    C = rp.correlation_matrix(H, basis=f"{basis}->{basis}")
    L = len(C) // 2
    oc = L // 2
    data = dict(H=H, C=C, L=L, oc=oc, basis=basis,
                chi_max=-1 if trunc.get("chi_max") is None else trunc["chi_max"],
                svd_min=trunc.get("svd_min", 1e-6))

    def put_bond(x, S):
        data[f"bond{x}_e"] = S.modes.e
        data[f"bond{x}_lam"] = S.schmidt_values
        data[f"bond{x}_sets"] = S.left_sets if S.left_sets is not None else S.right_sets[:, ::-1]
        data[f"bond{x}_pL"] = -1 if S.pL is None else S.pL
        data[f"bond{x}_pR"] = -1 if S.pR is None else S.pR
        q = np.zeros(S.n_schmidt, dtype=np.int64)
        for par, slc in S.idx_parity.items():
            q[slc] = (par + S.pL) % 2
        data[f"bond{x}_charge"] = q

    def put_site(i, td, chi_bra, chi_ket):
        data[f"site{i}_N"] = td.pfaffian_matrix
        data[f"site{i}_norm"] = td.norm
        data[f"site{i}_qtotal"] = td.qtotal
        M = np.zeros((2 * chi_bra, chi_ket), dtype=complex)
        for n_bra, sb in td.idx_n_bra.items():          # the block loop of pfaffian.py:1766-1776
            for n_ket, sk in td.idx_n_ket.items():
                if (n_bra + n_ket) % 2 == 1:
                    continue
                M[td.leg_idx_bra[sb], sk] = td.norm * rp._tensor_block(td.pfaffian_matrix, td.new_sets_bra[sb],
                                                                      td.new_sets_ket[sk])
        data[f"site{i}_T"] = M.reshape(2, chi_bra, chi_ket)

    centre = rp.SchmidtVectors.from_correlation_matrix(C, oc, trunc, basis=basis)
    put_bond(oc, centre)
    parity = centre.parity()
    prev = centre
    for i in range(oc, L):
        new = rp.SchmidtVectors.from_correlation_matrix(C, i + 1, trunc, which="R", basis=basis, total_parity=parity)
        put_bond(i + 1, new)
        put_site(i, rp.MPSTensorData.from_schmidt_vectors(new, prev, "right"), new.n_schmidt, prev.n_schmidt)
        prev = new
    prev = centre
    for i in reversed(range(oc)):
        new = rp.SchmidtVectors.from_correlation_matrix(C, i, trunc, which="L", basis=basis, total_parity=parity)
        put_bond(i, new)
        put_site(i, rp.MPSTensorData.from_schmidt_vectors(new, prev, "left"), new.n_schmidt, prev.n_schmidt)
        prev = new
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    print(name, "L", L, "max chi", max(len(data[f"bond{x}_lam"]) for x in range(L + 1)))


if __name__ == "__main__":
    ref = ref_shim.load("pass")
    import pfaffian_oracle as _po
    dump_pfaffian_case(ref, "pfaffian_random_L8", _po.random_bdg(8, 3, cplx=True), {"chi_max": 1000, "svd_min": 1e-7})
    dump_pfaffian_case(ref, "pfaffian_random_L9_real", _po.random_bdg(9, 4, cplx=False), {"chi_max": 1000, "svd_min": 1e-7})
    dump_pfaffian_case(ref, "pfaffian_random_L14_chi20", _po.random_bdg(14, 7, cplx=True), {"chi_max": 20})
    dump_pfaffian_case(ref, "pfaffian_kitaev_L16", _po.bdg_chain(16, mu=0.0, delta=0.05), {"chi_max": 24})
    dump_case(ref, "slater_random_L12", random_hamiltonian(12, 1), {"chi_max": 1000, "svd_min": 1e-7})
    dump_case(ref, "slater_random_L20_chi24", random_hamiltonian(20, 2), {"chi_max": 24})
    dump_case(ref, "slater_random_L11_N4", random_hamiltonian(11, 3), {"chi_max": 64}, N=4)
    dump_case(ref, "slater_chain_L16", hopping_chain(16), {"chi_max": 64})
    dump_case(ref, "slater_random_L40", random_hamiltonian(40, 1), {"chi_max": 64})
    # complex Hamiltonians (the reference's own acceptance example, examples/slater.py:15-27, is complex)
    dump_case(ref, "slater_complex_L12", random_hamiltonian(12, 5, cplx=True), {"chi_max": 1000, "svd_min": 1e-7})
    dump_case(ref, "slater_complex_L24_chi32", random_hamiltonian(24, 6, cplx=True), {"chi_max": 32})
    dump_lowest_sums(ref)
