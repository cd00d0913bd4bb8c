"""Neutral block-sparse MPS container and the TeNPy adapter.

TeNPy is not part of this image, so the conversion always produces a :class:`BlockMPS` (blocks,
leg charges, Schmidt values, canonical form labels -- exactly the information the reference puts
into ``networks.mps.MPS``, slater.py:1106-1143 and :1348-1351).  :meth:`BlockMPS.to_tenpy` builds
the TeNPy object with the same leg labels and charges when ``tenpy`` is importable.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class BlockMPS:
    L: int
    tensors: list              # engine.SiteTensor (fermion MPS) or DenseSite (projected spin MPS)
    lams: list                 # L+1 normalised Schmidt vectors
    charges: list              # L+1 int arrays: conserved charge left of the bond for every index
    form: list                 # "A" / "B" / None per site
    unit_cell_width: int
    ortho_center: int | None = None
    bc: str = "finite"
    site_type: str = "FermionSite"          # "FermionSite" | "SpinHalfSite"
    conserve: str | None = "N"
    meta: dict = field(default_factory=dict)

    @property
    def chi(self):
        return [len(l) for l in self.lams[1:-1]] if self.bc == "finite" else [len(l) for l in self.lams[:-1]]

    def get_B_dense(self, i) -> np.ndarray:
        """Site tensor ``T[vL, p, vR]`` as a dense array."""
        return self.tensors[i].dense()

    def entanglement_entropy(self):
        out = []
        for l in self.lams[1:-1] if self.bc == "finite" else self.lams[:-1]:
            p = np.asarray(l) ** 2
            p = p[p > 0]
            out.append(float(-(p * np.log(p)).sum()))
        return np.array(out)

    def norm_sq(self):
        return abs(self.overlap(self))

    def overlap(self, other: "BlockMPS"):
        """<self|other> by transfer matrices on the host (diagnostics / tests)."""
        E = np.ones((1, 1), dtype=complex)
        for i in range(self.L):
            A, B = self._scaled(i).conj(), other._scaled(i)
            E = np.einsum("ab,apc,bpd->cd", E, A, B, optimize=True)
        return E[0, 0]

    def _scaled(self, i):
        T = self.get_B_dense(i)
        oc = self.ortho_center if self.ortho_center is not None else 0
        if i == oc:
            T = T * self.lams[i][:, None, None]
        if oc == self.L and i == self.L - 1:
            T = T * self.lams[self.L][None, None, :]
        return T

    # ------------------------------------------------------------------------------------------
    def _site_to_npc(self, i, npc, site, chinfo):
        """One fermion site tensor as ``npc.Array``, block by block -- a mirror of the reference's
        ``MPSTensorData.to_npc_array`` (slater.py:1106-1143): bra leg = ``LegPipe([fermion_leg, leg_bra])``
        (physical index more major, sorted + bunched by TeNPy exactly like our bra rows: stable sort by pipe
        charge), ket leg from the sector table, blocks assigned by charge, ``split_legs()`` at the end.  No dense
        ``chi x 2 x chi`` intermediate."""
        t = self.tensors[i]
        left = t.mode == "left"
        qconj = (+1, -1) if left else (-1, +1)
        name_bra, name_ket = ("vL", "vR") if left else ("vR", "vL")
        q_bra = np.asarray(self.charges[i] if left else self.charges[i + 1])
        q_ket = np.asarray(self.charges[i + 1] if left else self.charges[i])

        def qdict(q):
            cuts = np.flatnonzero(np.diff(q)) + 1
            starts = np.concatenate(([0], cuts, [len(q)]))
            return {(int(q[a]),): slice(int(a), int(b)) for a, b in zip(starts[:-1], starts[1:])}

        leg_bra = npc.LegCharge.from_qdict(chinfo, qdict(q_bra), qconj=qconj[0])
        leg_ket = npc.LegCharge.from_qdict(chinfo, qdict(q_ket), qconj=qconj[1])
        pipe = npc.LegPipe([site.leg, leg_bra], qconj=leg_bra.qconj)
        blocks = t.blocks
        dtype = blocks[0][5].dtype if blocks else np.float64
        B = npc.zeros([pipe, leg_ket], labels=[f"(p.{name_bra})", name_ket], dtype=dtype, qtotal=(t.qtotal,))
        qd = pipe.to_qdict()
        for (q_k, r0, nr, c0, nc, arr) in blocks:
            sl_bra = qd[(q_k + t.qtotal * qconj[0],)]
            assert sl_bra.stop - sl_bra.start == nr, "bra pipe sector does not match the block"
            B[sl_bra, slice(c0, c0 + nc)] = np.asarray(arr)
        return B.split_legs()

    def to_tenpy(self):
        """The wave function as ``tenpy.networks.mps.MPS`` (labels ``vL, p, vR``; virtual charges =
        conserved charge to the left of the bond, as in slater.py:1111-1128).  Number-conserving fermion
        tensors are packed block by block (``_site_to_npc``); other site types through ``from_ndarray``.
        The ``BlockMPS`` stays attached to the result (``._temfpy_b200``) so that ``gutzwiller.abrikosov*``
        accept either object.  (TeNPy is not part of the build image: this adapter is exercised only where
        ``physics-tenpy >= 1.1`` is installed.)"""
        try:
            import tenpy.linalg.np_conserved as npc
            from tenpy import networks
        except ImportError as err:     # pragma: no cover - tenpy is not in the build image
            raise ImportError("`to_tenpy()` needs physics-tenpy >= 1.1.0") from err
        if self.site_type == "FermionSite":
            site = networks.site.FermionSite(conserve=self.conserve)
        else:
            site = networks.site.SpinHalfSite(conserve=self.conserve)
        chinfo = site.leg.chinfo
        tensors = []
        for i in range(self.L):
            t = self.tensors[i]
            if self.site_type == "FermionSite" and self.conserve == "N" and hasattr(t, "blocks") and hasattr(t, "mode") \
                    and self.bc == "finite":
                tensors.append(self._site_to_npc(i, npc, site, chinfo))
                continue
            T = self.get_B_dense(i)
            if chinfo.qnumber:
                qL = np.asarray(self.charges[i]).reshape(-1, 1)
                qR = np.asarray(self.charges[(i + 1) % len(self.charges)]).reshape(-1, 1)
                legL = npc.LegCharge.from_qflat(chinfo, qL, qconj=+1)
                legR = npc.LegCharge.from_qflat(chinfo, qR, qconj=-1)
                qtotal = self.tensors[i].qtotal if hasattr(self.tensors[i], "qtotal") else 0
                B = npc.Array.from_ndarray(T, [legL, site.leg, legR], labels=["vL", "p", "vR"],
                                           qtotal=[qtotal], cutoff=0.0)
            else:
                B = npc.Array.from_ndarray_trivial(T, labels=["vL", "p", "vR"])
            tensors.append(B)
        out = networks.mps.MPS([site] * self.L, tensors, self.lams, bc=self.bc, form=self.form,
                               unit_cell_width=self.unit_cell_width)
        out._temfpy_b200 = self
        return out
