import numpy as np
class BitArray:
    def __init__(self, array, num_bits):
        self.array = np.asarray(array, dtype=np.uint8); self.num_bits = num_bits
    @property
    def num_shots(self): return self.array.shape[-2]
    @classmethod
    def from_bool_array(cls, arr, order="big"):
        arr = np.asarray(arr, dtype=bool); nb = arr.shape[-1]
        pad = (-nb) % 8
        arr = np.pad(arr, [(0,0)]*(arr.ndim-1)+[(pad,0)])
        return cls(np.packbits(arr, axis=-1), nb)
