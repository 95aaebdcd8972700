"""engine presets (reference: src/liberate/fhe/presets/params.py:1-30): bronze logN14/K1, silver logN15/K2,
gold logN16/K4, platinum logN17/K6; scale_bits 40, maximal number of levels; gold/platinum use all GPUs."""


def _preset(logN, k, devices):
    return {"logN": logN, "num_special_primes": k, "devices": devices, "scale_bits": 40, "num_scales": None}


params = {
    "bronze": _preset(14, 1, [0]),
    "silver": _preset(15, 2, [0]),
    "gold": _preset(16, 4, None),
    "platinum": _preset(17, 6, None),
}
