"""dtype of the reference's `struct geodesic` (sim5kerr-geod.h:42-68) for reading the golden fixtures."""
import numpy as np


def geodesic_struct_dtype():
    return np.dtype([("a", "f8"), ("alpha", "f8"), ("beta", "f8"), ("incl", "f8"), ("cos_i", "f8"), ("l", "f8"), ("q", "f8"),
                     ("r1", "f8", 2), ("r2", "f8", 2), ("r3", "f8", 2), ("r4", "f8", 2), ("nrr", "i4"), ("type", "i4"),
                     ("m2p", "f8"), ("m2m", "f8"), ("mm", "f8"), ("mK", "f8"), ("rp", "f8"), ("dmdp_inf", "f8"),
                     ("Rpc", "f8"), ("Tpp", "f8"), ("Tip", "f8"), ("k", "f8", 4), ("p", "f8")])
