# API-compatible with the reference release this engine drops in for (src/liberate/fhe/version.py)
VERSION: str = "v0.9.0"
