"""Origin tags stamped into every ``data_struct``.

They are wire-format constants, not free text: objects pickled by the reference (``engine.save``) carry exactly these
strings and the engine's type checks compare against them, so a ciphertext or key written by one implementation loads
in the other (reference table: src/liberate/fhe/presets/types.py -- note the trailing colon of the rotation-key tag).
"""
_KEY_TAGS = (("sk", "secret"), ("pk", "public"), ("ksk", "key switch"), ("galk", "galois"), ("conjk", "conjugation"))

origins = {short: f"{name} key" for short, name in _KEY_TAGS}
origins["rotk"] = "rotation key:"
origins.update(ct="cipher text", ctt="cipher text triplet")
