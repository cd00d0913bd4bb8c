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
    def to_tenpy(self):
        """The wave function as ``tenpy.networks.mps.MPS`` (labels ``vL, p, vR``; virtual charges =
        conserved charge to the left of the bond, as in slater.py:1111-1128)."""
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
        lams = self.lams if self.bc == "finite" else self.lams
        return networks.mps.MPS([site] * self.L, tensors, lams, bc=self.bc, form=self.form,
                                unit_cell_width=self.unit_cell_width)
