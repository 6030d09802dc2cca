"""``pyscf.fci.addons.fix_spin_`` stand-in: records the penalty on the solver object (default shift 0.2,
pyscf's ``PENALTY``)."""


def fix_spin_(fciobj, shift=0.2, ss=None, **kwargs):
    fciobj._spin_penalty = (float(shift), 0.0 if ss is None else float(ss))
    return fciobj


fix_spin = fix_spin_
