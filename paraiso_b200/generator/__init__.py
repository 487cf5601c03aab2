"""Code generation entry points (mirrors Language/Paraiso/Generator.hs:29-66)."""
