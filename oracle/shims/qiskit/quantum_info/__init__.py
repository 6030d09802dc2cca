import numpy as np
class Pauli:
    def __init__(self, label):
        lab = label[::-1]
        self.x = np.array([c in "XY" for c in lab]); self.z = np.array([c in "ZY" for c in lab]); self._label = label
    def to_label(self): return self._label
class SparsePauliOp:
    def __init__(self, labels, coeffs=None):
        if isinstance(labels, str): labels = [labels]
        self.paulis = [Pauli(l) for l in labels]
        self.coeffs = np.ones(len(labels), dtype=complex) if coeffs is None else np.asarray(coeffs, dtype=complex)
    @property
    def size(self): return len(self.paulis)
