"""Stand-in for CompressAI (absent, unpinned third-party dependency of the reference).

TEST INFRASTRUCTURE ONLY. `layers` is exact (one-liners upstream); the entropy models are
structural stand-ins -> likelihood/bpp values are NOT CompressAI's ("parity unpinned" at
that boundary). x_hat / latents / rounded symbols do not depend on them
except through `_get_medians()` (zeros here, as at CompressAI init).
"""
