"""Pfaffian (BCS / Bogoliubov) states -> MPS.  Reference: src/temfpy/pfaffian.py (2242 lines).

Scheduled after the Slater path: it needs complex128 variants of
the mode extraction and of the minors kernel (batched Pfaffians, reference pfaffian.py:1413-1479).
Not available in this release; every entry point fails loudly (there is no CPU fallback)."""


def _na(name):
    def f(*args, **kwargs):
        raise NotImplementedError(f"temfpy_b200.pfaffian.{name}: the Pfaffian path is not implemented yet "
                                  "(SURVEY 8a rows a13-a19); use the reference for BCS states")
    f.__name__ = name
    return f


correlation_matrix = _na("correlation_matrix")
C_to_MPS = _na("C_to_MPS")
H_to_MPS = _na("H_to_MPS")
C_to_iMPS = _na("C_to_iMPS")
H_to_iMPS = _na("H_to_iMPS")
